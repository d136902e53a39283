"""Worker of tests/test_gpu_fast_fp.py: runs in its own process with NB2_FP=fast (the library is chosen at import time) and
prints, as one JSON line, how far the FMA-contracted twin library is from the CPU oracle on the benchmark scenes."""

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import newton_b200  # noqa: E402
import oracle  # noqa: E402
from newton_b200 import _lib, scenes  # noqa: E402
from tests.helpers import rel_err, simulate  # noqa: E402


def main():
    assert _lib.FP_MODE == "fast" and _lib.LIB_PATH.endswith("libnewton_b200_fast.so"), _lib.LIB_PATH
    oracle.build()
    out = {"lib": os.path.basename(_lib.LIB_PATH)}
    cases = {
        "quadruped_xpbd": (scenes.quadruped_model(32, seed=1), "xpbd", {"iterations": 8}, 0.005),
        "quadruped_featherstone": (scenes.quadruped_model(32, seed=1), "featherstone", {}, 0.001),
        "box_stacks_xpbd": (scenes.box_stack_model(16, seed=0), "xpbd", {"iterations": 8}, 1.0 / 240),
    }
    for name, (model, solver, kw, dt) in cases.items():
        o_cls = oracle.SolverXPBD if solver == "xpbd" else oracle.SolverFeatherstone
        g_cls = newton_b200.solvers.SolverXPBD if solver == "xpbd" else newton_b200.solvers.SolverFeatherstone
        ref, _, rc = simulate(model, oracle.CollisionPipeline, o_cls, substeps=100, dt=dt, solver_kwargs=kw, record_contacts=True)
        got, _, gc = simulate(model.to("cuda:0"), newton_b200.CollisionPipeline, g_cls, substeps=100, dt=dt, solver_kwargs=kw,
                              record_contacts=True)
        out[name] = {
            "body_q_rel": rel_err(got.body_q.cpu().numpy(), ref.body_q.numpy()),
            "body_qd_abs": float(np.max(np.abs(got.body_qd.cpu().numpy() - ref.body_qd.numpy()))),
            "body_qd_scale": float(np.max(np.abs(ref.body_qd.numpy()))),
            "counts_equal": rc == gc,
            "contacts_last": gc[-1],
            "bit_equal": bool(np.array_equal(got.body_q.cpu().numpy(), ref.body_q.numpy())),
        }
    print("FAST_FP_RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()

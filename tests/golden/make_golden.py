"""Regenerates tests/golden/*.npz.

The reference itself cannot run in this environment (NVIDIA Warp is not installable), so these vectors are
produced by the CPU oracle - the restatement that tests/test_oracle_known_answers.py pins against the reference's own
known-answer checks.  They freeze the oracle's outputs: a later change of compiler, flags or oracle source that alters a
single bit shows up in tests/test_golden.py, and the CUDA path is compared with them on the GPU box without running the
oracle at all.

    python tests/golden/make_golden.py        # from the repository root
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from newton_b200 import scenes  # noqa: E402
from tests.helpers import canonical_contacts, simulate  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

def _standing_quadrupeds():
    import newton_b200

    m = scenes.quadruped_model(2, seed=1)
    m.joint_q.view(2, -1)[:, 2] = 0.46  # feet in contact from the first substep
    scenes.host_fk(m, m.joint_q, m.joint_qd, m)
    return m


# name -> (scene factory, solver, solver kwargs, substeps, dt)
CASES = {
    "quadruped_xpbd": (lambda: scenes.quadruped_model(2, seed=1), "SolverXPBD", {"iterations": 4}, 30, 0.005),
    "box_stack_xpbd": (lambda: scenes.box_stack_model(1, seed=0), "SolverXPBD", {"iterations": 4}, 30, 1.0 / 240),
    "convex_pile_xpbd": (lambda: scenes.convex_pile_model(1, seed=5), "SolverXPBD", {"iterations": 4}, 40, 1.0 / 240),
    "quadruped_featherstone": (_standing_quadrupeds, "SolverFeatherstone", {}, 40, 0.001),
    "pendulum_featherstone": (lambda: scenes.pendulum_model(), "SolverFeatherstone", {}, 100, 0.001),
}


def run(name, pipeline_cls, solver_pkg, to_device=None):
    factory, solver, kw, substeps, dt = CASES[name]
    model = factory()
    if to_device is not None:
        model = model.to(to_device)
    state, contacts, counts = simulate(model, pipeline_cls, getattr(solver_pkg, solver), substeps=substeps, dt=dt, solver_kwargs=kw,
                                       record_contacts=True)
    out = {"contact_counts": np.asarray(counts, dtype=np.int32)}
    for k in ("body_q", "body_qd", "joint_q", "joint_qd"):
        v = getattr(state, k, None)
        if v is not None and (k.startswith("body") or solver == "SolverFeatherstone"):
            out[k] = v.detach().cpu().numpy()
    n, cc = canonical_contacts(contacts, model)
    for k, v in cc.items():
        out["contact_" + k] = v
    return out


if __name__ == "__main__":
    oracle.build()
    for name in CASES:
        out = run(name, oracle.CollisionPipeline, oracle)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})

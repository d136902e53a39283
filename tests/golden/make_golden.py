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


def _speculative_heap():
    import torch

    m = scenes.free_bodies_model(3, seed=5)
    g = torch.Generator().manual_seed(2)
    m.body_qd.copy_((torch.rand(m.body_qd.shape, generator=g) * 2.0 - 1.0) * torch.tensor([6.0, 6.0, 6.0, 3.0, 3.0, 3.0]))
    return m


def _universal_joint():
    import torch

    from tests.test_d6_two_angular_axes import _universal_model

    m, _ = _universal_model(gravity=-9.81)
    m.joint_q.copy_(torch.tensor([0.5, 0.4, -0.3]))
    m.joint_qd.copy_(torch.tensor([1.5, -2.0, 1.0]))
    scenes.host_fk(m, m.joint_q, m.joint_qd, m)  # initial body state (the solver recomputes FK from joint_q every step)
    return m


def _mesh_on_plane():
    from tests.test_mesh_plane import mixed_mesh_model

    return mixed_mesh_model(2, seed=4)


def _speculative_kwargs():
    from newton_b200 import SpeculativeContactConfig

    return {"broad_phase": "sap", "speculative_config": SpeculativeContactConfig(max_speculative_extension=0.3)}


# name -> (scene factory, solver, solver kwargs, substeps, dt[, pipeline kwargs factory, collide dt, collide])
CASES = {
    "quadruped_xpbd": (lambda: scenes.quadruped_model(2, seed=1), "SolverXPBD", {"iterations": 4}, 30, 0.005),
    "box_stack_xpbd": (lambda: scenes.box_stack_model(1, seed=0), "SolverXPBD", {"iterations": 4}, 30, 1.0 / 240),
    "convex_pile_xpbd": (lambda: scenes.convex_pile_model(1, seed=5), "SolverXPBD", {"iterations": 4}, 40, 1.0 / 240),
    "quadruped_featherstone": (_standing_quadrupeds, "SolverFeatherstone", {}, 40, 0.001),
    "pendulum_featherstone": (lambda: scenes.pendulum_model(), "SolverFeatherstone", {}, 100, 0.001),
    # round 2: speculative contacts through the SAP broad phase (60 ms horizon), and a D6 joint with two angular axes
    "heap_speculative_xpbd": (_speculative_heap, "SolverXPBD", {"iterations": 4}, 40, 0.004, _speculative_kwargs, 0.06, True),
    "universal_joint_featherstone": (_universal_joint, "SolverFeatherstone", {"angular_damping": 0.0}, 150, 0.001, None, None, False),
    # round 2: triangle-mesh shapes against the ground plane (one contact per vertex), boxes / a sphere around them
    "mesh_plane_xpbd": (_mesh_on_plane, "SolverXPBD", {"iterations": 4}, 60, 1.0 / 240, lambda: {"reduce_contacts": False}, None, True),
}


def run(name, pipeline_cls, solver_pkg, to_device=None):
    factory, solver, kw, substeps, dt = CASES[name][:5]
    pipe_kw, collide_dt, collide = (CASES[name][5:] + (None, None, True))[:3] if len(CASES[name]) > 5 else (None, None, True)
    model = factory()
    if to_device is not None:
        model = model.to(to_device)
    state, contacts, counts = simulate(model, pipeline_cls, getattr(solver_pkg, solver), substeps=substeps, dt=dt, solver_kwargs=kw,
                                       record_contacts=collide, pipeline_kwargs=pipe_kw() if pipe_kw else None, collide_dt=collide_dt,
                                       collide=collide)
    out = {"contact_counts": np.asarray(counts, dtype=np.int32)}
    for k in ("body_q", "body_qd", "joint_q", "joint_qd"):
        v = getattr(state, k, None)
        if v is not None and (k.startswith("body") or solver == "SolverFeatherstone"):
            out[k] = v.detach().cpu().numpy()
    if contacts is not None:
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

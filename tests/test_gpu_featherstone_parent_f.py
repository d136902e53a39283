"""GPU parity test for ``State.body_parent_f`` from ``featherstone_step_kernel<L, true>`` (reference compute_body_parent_f,
featherstone/kernels.py:2371-2416).  The oracle side is pinned on the CPU by
``tests/test_oracle_known_answers_more.py::test_parent_force_static_pendulum_xpbd_and_featherstone_agree``."""

import numpy as np
import pytest

import newton_b200
from newton_b200 import scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scene", ["quadrupeds", "pendulum", "mixed_worlds"])
def test_featherstone_body_parent_f_bit_exact(cuda_lib, oracle_lib, scene):
    import oracle

    cpu = {"quadrupeds": lambda: scenes.quadruped_model(5, seed=4), "pendulum": scenes.pendulum_model,
           "mixed_worlds": lambda: scenes.mixed_worlds_model(2)}[scene]()
    cpu.request_state_attributes("body_parent_f")
    gpu = cpu.to("cuda:0")
    gpu.request_state_attributes("body_parent_f")

    def run(model, pipeline_cls, solver_cls):
        solver, pipe = solver_cls(model, angular_damping=0.05), pipeline_cls(model)
        s0, s1, ctrl, contacts = model.state(), model.state(), model.control(), pipe.contacts()
        for _ in range(40):
            s0.clear_forces()
            pipe.collide(s0, contacts)
            solver.step(s0, s1, ctrl, contacts, 1e-3)
            s0, s1 = s1, s0
        return s0

    ref = run(cpu, oracle.CollisionPipeline, oracle.SolverFeatherstone)
    out = run(gpu, newton_b200.CollisionPipeline, newton_b200.solvers.SolverFeatherstone)
    assert np.abs(ref.body_parent_f.numpy()).max() > 0.0
    for name in ("body_parent_f", "body_q", "joint_q", "joint_qd"):
        assert np.array_equal(getattr(out, name).cpu().numpy().view(np.uint32), getattr(ref, name).numpy().view(np.uint32)), name

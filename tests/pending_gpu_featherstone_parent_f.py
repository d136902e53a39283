"""NOT COLLECTED (no ``test_`` prefix): GPU parity test for ``State.body_parent_f`` from ``featherstone_step_kernel<L, true>``.

The kernel code was written at the end of round 1 without GPU time left to run it; the plain-step instantiation is byte-identical
to the validated one (SASS diff: 0 lines), this one is unexercised.  Next step: run
``python -m pytest tests/pending_gpu_featherstone_parent_f.py -m gpu -q -o python_files=pending_*.py`` on a B200; when green, rename the
file to ``test_gpu_featherstone_parent_f.py`` and drop ``SolverFeatherstone._parent_f_validated`` together with the guard in
``step()``.  The oracle side (``oracle_featherstone.h``, compute_body_parent_f) is pinned on the CPU already
(``tests/test_oracle_known_answers_more.py::test_parent_force_static_pendulum_xpbd_and_featherstone_agree``).
"""

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

    def run(model, pipeline_cls, solver_cls, unlock):
        solver, pipe = solver_cls(model, angular_damping=0.05), pipeline_cls(model)
        if unlock:
            solver._parent_f_validated = True
        s0, s1, ctrl, contacts = model.state(), model.state(), model.control(), pipe.contacts()
        for _ in range(40):
            s0.clear_forces()
            pipe.collide(s0, contacts)
            solver.step(s0, s1, ctrl, contacts, 1e-3)
            s0, s1 = s1, s0
        return s0

    ref = run(cpu, oracle.CollisionPipeline, oracle.SolverFeatherstone, False)
    out = run(gpu, newton_b200.CollisionPipeline, newton_b200.solvers.SolverFeatherstone, True)
    assert np.abs(ref.body_parent_f.numpy()).max() > 0.0
    for name in ("body_parent_f", "body_q", "joint_q", "joint_qd"):
        assert np.array_equal(getattr(out, name).cpu().numpy().view(np.uint32), getattr(ref, name).numpy().view(np.uint32)), name

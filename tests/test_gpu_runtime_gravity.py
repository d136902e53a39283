"""Per-world gravity changed at run time (reference ``Model.set_gravity`` + ``notify_model_changed``,
``newton/tests/test_runtime_gravity.py``): the CUDA solvers read ``gravity[body_world]`` live, bit for bit like the oracle."""

import numpy as np
import pytest

import newton_b200
from newton_b200 import ModelFlags, scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("solver_name", ["xpbd", "featherstone"])
def test_set_gravity_between_steps_bit_exact(cuda_lib, oracle_lib, solver_name):
    import oracle

    W = 5
    cpu = scenes.quadruped_model(W, seed=2)
    gpu = cpu.to("cuda:0")
    dt = 0.005 if solver_name == "xpbd" else 0.001
    kw = {"iterations": 2} if solver_name == "xpbd" else {"angular_damping": 0.05}

    def run(model, pipeline_cls, solver_cls):
        solver, pipe = solver_cls(model, **kw), pipeline_cls(model)
        s0, s1, ctrl, contacts = model.state(), model.state(), model.control(), pipe.contacts()
        schedule = {
            0: lambda: None,
            1: lambda: model.set_gravity((0.0, 0.0, -1.0), world=1),                                  # one world
            2: lambda: model.set_gravity(np.linspace(0.0, -9.81, W)[:, None] * np.array([[0.0, 0.0, 1.0]])),  # curriculum array
            3: lambda: model.set_gravity((0.3, 0.0, -9.81)),                                          # every world + global slot
        }
        for phase in range(4):
            schedule[phase]()
            solver.notify_model_changed(ModelFlags.MODEL_PROPERTIES)
            for _ in range(8):
                s0.clear_forces()
                pipe.collide(s0, contacts)
                solver.step(s0, s1, ctrl, contacts, dt)
                s0, s1 = s1, s0
        return s0

    solver_cls = (oracle.SolverXPBD, newton_b200.solvers.SolverXPBD) if solver_name == "xpbd" else \
        (oracle.SolverFeatherstone, newton_b200.solvers.SolverFeatherstone)
    ref = run(cpu, oracle.CollisionPipeline, solver_cls[0])
    out = run(gpu, newton_b200.CollisionPipeline, solver_cls[1])
    assert np.array_equal(gpu.numpy("gravity"), cpu.numpy("gravity"))
    for name in ("body_q", "body_qd") + (("joint_q", "joint_qd") if solver_name == "featherstone" else ()):
        assert np.array_equal(getattr(out, name).cpu().numpy().view(np.uint32), getattr(ref, name).numpy().view(np.uint32)), name
    # the worlds really saw different gravity: base heights differ between world 0 (g = 0 during phase 2) and world 4
    z = ref.body_q.view(W, 13, 7)[:, 0, 2].numpy()
    assert z[0] != z[4]

"""GPU parity of the fused Featherstone kernel vs the CPU oracle (bit-identical state after 100 substeps)."""

import numpy as np
import pytest
import torch

import newton_b200
from newton_b200 import scenes
from newton_b200.sim.builder import JointDofConfig, ModelBuilder
from newton_b200.utils import xform as X
from tests.helpers import simulate

pytestmark = pytest.mark.gpu


def _compare(model, oracle, substeps, dt, kw, collide=True):
    rs, _, rc = simulate(model, oracle.CollisionPipeline, oracle.SolverFeatherstone, substeps=substeps, dt=dt, solver_kwargs=kw,
                         collide=collide, record_contacts=collide)
    mg = model.to("cuda:0")
    gs, _, gc = simulate(mg, newton_b200.CollisionPipeline, newton_b200.solvers.SolverFeatherstone, substeps=substeps, dt=dt,
                         solver_kwargs=kw, collide=collide, record_contacts=collide)
    torch.cuda.synchronize()
    assert rc == gc
    for name in ("joint_q", "joint_qd", "body_q", "body_qd"):
        a, b = getattr(gs, name).cpu().numpy(), getattr(rs, name).numpy()
        assert np.isfinite(b).all(), name
        np.testing.assert_array_equal(a, b, err_msg=name)
    return rs


def test_config1_pendulum(oracle_lib, cuda_lib):
    """BASELINE.json configs[0]: example_basic_pendulum scene, SolverFeatherstone, 100 fps x 10 substeps."""
    model = scenes.pendulum_model()
    rs = _compare(model, oracle_lib, 300, 1.0 / 100 / 10, {"angular_damping": 0.05})
    q = rs.body_q.numpy()
    assert np.all(np.abs(q[:, 0]) < 1e-5) and np.all(q[:, 2] < 5.0 + 1e-4)  # example_basic_pendulum.py:114-122


@pytest.mark.parametrize("world_count,interval", [(1, 1), (8, 1), (5, 3)])
def test_config4_quadruped_with_penalty_contacts(oracle_lib, cuda_lib, world_count, interval):
    """BASELINE.json configs[3] at reduced env count: quadrupeds, SolverFeatherstone(angular_damping=0.05), dt = 1 ms,
    penalty contacts against the ground, floating base (solve origin at the root COM)."""
    model = scenes.quadruped_model(world_count, seed=1)
    model.joint_q.view(world_count, -1)[:, 2] = 0.47
    scenes.host_fk(model, model.joint_q, model.joint_qd, model)
    _compare(model, oracle_lib, 100, 1e-3, {"angular_damping": 0.05, "update_mass_matrix_interval": interval})


def test_mixed_joint_types(oracle_lib, cuda_lib):
    """Prismatic + revolute + ball + fixed + D6 chain hanging from the world, with joint_f and PD targets."""
    b = ModelBuilder()
    prev = -1
    links = []
    for i in range(5):
        link = b.add_link(xform=X.transform((0.0, 0.0, 2.0 - 0.3 * i)))
        b.add_shape_box(link, hx=0.05, hy=0.05, hz=0.12)
        links.append(link)
    px = X.transform((0.0, 0.0, -0.15))
    cx = X.transform((0.0, 0.0, 0.15))
    j = [
        b.add_joint_prismatic(-1, links[0], axis=(0.0, 0.0, 1.0), parent_xform=X.transform((0.0, 0.0, 2.15)), child_xform=cx,
                              target_ke=500.0, target_kd=5.0, limit_lower=-0.5, limit_upper=0.5),
        b.add_joint_revolute(links[0], links[1], axis=(0.0, 1.0, 0.0), parent_xform=px, child_xform=cx, target_ke=20.0, target_kd=0.5,
                             limit_lower=-1.0, limit_upper=1.0),
        b.add_joint_ball(links[1], links[2], parent_xform=px, child_xform=cx),
        b.add_joint_fixed(links[2], links[3], parent_xform=px, child_xform=cx),
        b.add_joint_d6(links[3], links[4], parent_xform=px, child_xform=cx,
                       linear_axes=[JointDofConfig(axis=(1.0, 0.0, 0.0), limit_lower=-0.1, limit_upper=0.1, target_ke=100.0, target_kd=1.0)],
                       angular_axes=[JointDofConfig(axis=(1.0, 0.0, 0.0)), JointDofConfig(axis=(0.0, 1.0, 0.0)),
                                     JointDofConfig(axis=(0.0, 0.0, 1.0))]),
    ]
    b.add_articulation(j)
    model = b.finalize()
    g = torch.Generator().manual_seed(3)
    model.joint_q[1] = 0.3
    model.joint_qd.copy_(torch.rand(model.joint_qd.shape, generator=g) * 0.4 - 0.2)
    model.joint_f.copy_(torch.rand(model.joint_f.shape, generator=g) * 0.2 - 0.1)
    _compare(model, oracle_lib, 100, 1e-3, {"angular_damping": 0.0}, collide=False)


def test_eval_fk_kernel_bit_exact(oracle_lib, cuda_lib):
    """nb2_eval_fk (one thread per articulation) vs the oracle's newton.eval_fk, random joint velocities."""
    for model in (scenes.quadruped_model(37, seed=3), scenes.pendulum_model()):
        g = torch.Generator().manual_seed(1)
        model.joint_qd.copy_(torch.rand(model.joint_qd.shape, generator=g) - 0.5)
        mg = model.to("cuda:0")
        mg.body_q.zero_()
        mg.body_qd.zero_()
        newton_b200.eval_fk(mg, mg.joint_q, mg.joint_qd, mg)
        oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, model)
        np.testing.assert_array_equal(mg.body_q.cpu().numpy(), model.body_q.numpy())
        np.testing.assert_array_equal(mg.body_qd.cpu().numpy(), model.body_qd.numpy())


def test_heterogeneous_worlds_bit_exact(oracle_lib, cuda_lib):
    """Featherstone over worlds of different sizes (18-dof quadruped, three free boxes, an empty world, a 1-dof link)."""
    model = scenes.mixed_worlds_model(3)
    ref, _, rc = simulate(model, oracle_lib.CollisionPipeline, oracle_lib.SolverFeatherstone, substeps=80, dt=1.0 / 480,
                          record_contacts=True)
    mg = model.to("cuda:0")
    gpu, _, gc = simulate(mg, newton_b200.CollisionPipeline, newton_b200.solvers.SolverFeatherstone, substeps=80, dt=1.0 / 480,
                          record_contacts=True)
    assert rc == gc
    for name in ("joint_q", "joint_qd", "body_q", "body_qd"):
        np.testing.assert_array_equal(getattr(gpu, name).cpu().numpy(), getattr(ref, name).numpy(), err_msg=name)


def test_eval_ik_kernel_bit_exact(oracle_lib, cuda_lib):
    """nb2_eval_ik vs the oracle's newton.eval_ik on states produced by 40 XPBD substeps (maximal-coordinate solver:
    the joint coordinates are only recoverable through eval_ik), plus the fk -> ik round trip on the device."""
    for model in (scenes.quadruped_model(9, seed=3), scenes.mixed_worlds_model(2)):
        state, _, _ = simulate(model, oracle_lib.CollisionPipeline, oracle_lib.SolverXPBD, substeps=40, dt=0.004,
                               solver_kwargs={"iterations": 3})
        q_ref, qd_ref = torch.zeros_like(model.joint_q), torch.zeros_like(model.joint_qd)
        oracle_lib.eval_ik(model, state, q_ref, qd_ref)
        mg = model.to("cuda:0")
        sg = mg.state()
        sg.body_q.copy_(state.body_q)
        sg.body_qd.copy_(state.body_qd)
        q, qd = torch.zeros_like(mg.joint_q), torch.zeros_like(mg.joint_qd)
        newton_b200.eval_ik(mg, sg, q, qd)
        np.testing.assert_array_equal(q.cpu().numpy(), q_ref.numpy())
        np.testing.assert_array_equal(qd.cpu().numpy(), qd_ref.numpy())
        newton_b200.eval_fk(mg, mg.joint_q, mg.joint_qd, sg)
        newton_b200.eval_ik(mg, sg, q, qd)
        np.testing.assert_allclose(q.cpu().numpy(), mg.joint_q.cpu().numpy(), atol=1e-6)

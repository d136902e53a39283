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


def test_kinematic_bodies_and_in_place_stepping_bit_exact(oracle_lib, cuda_lib):
    """Kinematic links under SolverFeatherstone (solver_featherstone.py:212-281, kernels.py:55-63, 1933-1976): a kinematic free base
    with prescribed joint state pushing a dynamic probe, a kinematic revolute root carrying a dynamic pendulum and a kinematic
    fixed root - effective armature 1e10, zeroed accelerations, joint state copied through, body forces ignored.  The GPU run
    steps IN PLACE (state_in is state_out, solver_featherstone.py:472) and must still equal the oracle's two-state run."""
    import torch

    from newton_b200 import ModelBuilder
    from newton_b200.utils import xform as X

    b = ModelBuilder(gravity=(0.0, 0.0, 0.0))  # like the reference scenes (test_kinematic_links.py:304-398)
    b.default_shape_cfg.ke, b.default_shape_cfg.kd, b.default_shape_cfg.kf = 1.0e4, 500.0, 0.5
    for w in range(3):
        b.begin_world()
        kin = b.add_body(xform=X.transform((-0.3, 0.0, 0.5)), mass=1.0, is_kinematic=True)
        b.add_shape_box(kin, hx=0.25, hy=0.15, hz=0.15)
        probe = b.add_body(xform=X.transform((0.1 + 0.02 * w, 0.0, 0.5)), mass=1.0)
        b.add_shape_sphere(probe, radius=0.1)
        root = b.add_link(xform=X.transform((0.0, 1.0, 1.0)), mass=1.0, inertia=np.eye(3) * 0.1, is_kinematic=True)
        pend = b.add_link(xform=X.transform((0.45, 1.0, 1.0)), mass=1.0, inertia=np.eye(3) * 0.1)
        j0 = b.add_joint_revolute(-1, root, axis=(0.0, 1.0, 0.0), parent_xform=X.transform((0.0, 1.0, 1.0)))
        j1 = b.add_joint_revolute(root, pend, axis=(0.0, 1.0, 0.0), parent_xform=X.transform((0.45, 0.0, 0.0)))
        b.add_articulation([j0, j1])
        b.add_shape_sphere(pend, xform=X.transform((0.3, 0.0, 0.0)), radius=0.1)
        fixed = b.add_link(xform=X.transform((0.0, -1.0, 0.3)), mass=1.0, inertia=np.eye(3) * 0.1, is_kinematic=True)
        b.add_articulation([b.add_joint_fixed(-1, fixed, parent_xform=X.transform((0.0, -1.0, 0.3)))])
        b.add_shape_box(fixed, hx=0.25, hy=0.25, hz=0.25)
        b.end_world()
    b.add_ground_plane()
    model = b.finalize()
    scenes.host_fk(model, model.joint_q, model.joint_qd, model)
    mg = model.to("cuda:0")
    kin_q0 = int(model.joint_q_start[0])
    dt = 1.0 / 480.0

    def run(m, pipeline_cls, solver_cls, in_place):
        solver, pipe = solver_cls(m, angular_damping=0.0), pipeline_cls(m)
        s0, s1, contacts = m.state(), m.state(), pipe.contacts()
        per_world_q = m.joint_coord_count // 3
        per_world_qd = m.joint_dof_count // 3
        for i in range(120):
            for w in range(3):  # prescribed motion of the kinematic free base (generalized coordinates, like the reference test)
                s0.joint_q[w * per_world_q + kin_q0] = -0.3 + 1.0 * i * dt
                s0.joint_qd[w * per_world_qd] = 1.0
                s0.joint_qd[w * per_world_qd + 12] = 2.0  # the kinematic revolute root keeps turning
            s0.clear_forces()
            s0.body_f[0] = torch.tensor([20.0, -15.0, 10.0, 0.5, -0.4, 0.3], device=s0.body_f.device)  # ignored: body 0 is kinematic
            pipe.collide(s0, contacts)
            if in_place:
                solver.step(s0, s0, None, contacts, dt)
            else:
                solver.step(s0, s1, None, contacts, dt)
                s0, s1 = s1, s0
        return s0

    ref = run(model, oracle_lib.CollisionPipeline, oracle_lib.SolverFeatherstone, False)
    assert int(model.joint_qd_start[2]) == 12  # dof 12 = the revolute root joint of world 0 (two 6-dof free joints before it)
    for in_place in (False, True):
        got = run(mg, newton_b200.CollisionPipeline, newton_b200.solvers.SolverFeatherstone, in_place)
        for name in ("joint_q", "joint_qd", "body_q", "body_qd"):
            np.testing.assert_array_equal(getattr(got, name).cpu().numpy(), getattr(ref, name).numpy(), err_msg=f"{name} in_place={in_place}")
    q = ref.body_q.numpy().reshape(3, -1, 7)
    assert abs(q[0, 0, 0] - (-0.3 + 119 * dt)) < 0.05 and abs(q[0, 0, 2] - 0.5) < 1e-4  # prescribed, does not fall
    assert q[0, 1, 0] > 0.15  # the probe was pushed


def test_eval_fk_body_flag_filter(oracle_lib, cuda_lib):
    """newton.eval_fk(..., body_flag_filter=BodyFlags.KINEMATIC / DYNAMIC) (sim/articulation.py:254, 421): only bodies whose flags
    intersect the filter are written, the others keep their (stale) values and their descendants are computed from those."""
    import torch

    from newton_b200 import BodyFlags

    model = scenes.quadruped_model(6, seed=5)
    model.body_flags[::13] = int(BodyFlags.KINEMATIC)  # every trunk kinematic, the legs dynamic
    mg = model.to("cuda:0")
    rng = np.random.default_rng(0)
    jq = model.joint_q.clone() + torch.from_numpy(rng.normal(scale=0.05, size=model.joint_q.shape).astype(np.float32))
    jqd = torch.from_numpy(rng.normal(scale=0.3, size=model.joint_qd.shape).astype(np.float32))
    for flt in (int(BodyFlags.KINEMATIC), int(BodyFlags.DYNAMIC), 3):
        so, sg = model.state(), mg.state()
        so.body_q += 0.01  # stale values that the filtered-out bodies must keep
        sg.body_q += 0.01
        oracle_lib.eval_fk(model, jq, jqd, so, body_flag_filter=flt)
        newton_b200.eval_fk(mg, jq.to("cuda:0"), jqd.to("cuda:0"), sg, body_flag_filter=flt)
        np.testing.assert_array_equal(sg.body_q.cpu().numpy(), so.body_q.numpy(), err_msg=str(flt))
        np.testing.assert_array_equal(sg.body_qd.cpu().numpy(), so.body_qd.numpy(), err_msg=str(flt))

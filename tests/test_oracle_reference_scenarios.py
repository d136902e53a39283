"""More of the reference's own known answers, restated for the oracle (CPU): closed forms and equivalences from test files that drive
the same joint / FK arithmetic through solvers this library does not replace (SolverMuJoCo, newton.ik).  The expectation is the
reference test's; the solver under test is the oracle's Featherstone / XPBD, whose kernels the CUDA path reproduces bit for bit.

* ``newton/tests/test_joint_drive.py:14-260`` - explicit PD joint drive, one step: qd' = qd + (ke (q* - q) + kd (qd* - qd) + M g) dt / M
* ``newton/tests/test_pendulum_revolute_vs_d6.py:14-102`` - a D6 joint with one angular axis swings like the revolute joint
* ``newton/tests/test_ik_fk_kernels.py:53-216`` - eval_fk of the five joint kinds with offset joint frames against an independent
  walk (there: ``IKSolver._fk_two_pass``; here: the float64 NumPy walk of ``utils/host_fk.py``), 1e-6, and eval_ik inverts it
"""

import math

import numpy as np
import pytest
import torch

import newton_b200
from newton_b200.sim.builder import JointDofConfig, ModelBuilder
from newton_b200.sim.enums import JointType
from newton_b200.utils import xform as X
from newton_b200.utils.host_fk import host_fk


# ---- test_joint_drive.py ---------------------------------------------------------------------------------------------------------
def _expected_qd(g, dt, M, pos_t, vel_t, ke, kd, q, qd):
    F = ke * (pos_t - q) + kd * (vel_t - qd) + M * g
    return qd + F * dt / M


@pytest.mark.parametrize("is_prismatic", [True, False])
@pytest.mark.parametrize("up_axis", ["x", "y", "z"])
@pytest.mark.parametrize("motion_axis", [0, 1, 2])
def test_joint_drive_no_limits(oracle_lib, is_prismatic, up_axis, motion_axis):
    up = "xyz".index(up_axis)
    g = 5.0 if (is_prismatic and up == motion_axis) else 0.0
    dt = 0.01
    masses, inertias = [10.0, 20.0], [4.0, 8.0]
    q0, qd0 = [100.0, 205.0], [10.0, 25.0]
    pos_t, vel_t = [200.0, 300.0], [0.0, 0.0]
    kes, kds = [100.0, 200.0], [10.0, 20.0]
    gravity = tuple(g if k == up else 0.0 for k in range(3))
    axis = tuple(1.0 if k == motion_axis else 0.0 for k in range(3))
    main = ModelBuilder(up_axis=up_axis, gravity=gravity)
    cfg = newton_b200.ShapeConfig()
    cfg.density, cfg.has_shape_collision = 0.0, False
    for i in range(2):
        main.begin_world()
        body = main.add_link(mass=masses[i], inertia=np.eye(3) * inertias[i], com=(0.0, 0.0, 0.0))
        main.add_shape_sphere(body, radius=1.0, cfg=cfg)
        kw = dict(parent=-1, child=body, axis=axis, target_pos=pos_t[i], target_vel=vel_t[i], target_ke=kes[i], target_kd=kds[i],
                  armature=0.0, friction=0.0, limit_lower=-1.0e10, limit_upper=1.0e10, limit_ke=0.0, limit_kd=0.0)
        j = main.add_joint_prismatic(**kw) if is_prismatic else main.add_joint_revolute(**kw)
        main.add_articulation([j])
        main.end_world()
        main.joint_q[i], main.joint_qd[i] = q0[i], qd0[i]
    model = main.finalize()
    M = masses if is_prismatic else inertias
    solver = oracle_lib.SolverFeatherstone(model, angular_damping=0.0)
    s0, s1, control = model.state(), model.state(), model.control()

    def one_step(q, qd, tq, tqd):
        s0.joint_q.copy_(torch.tensor(q, dtype=torch.float32))
        s0.joint_qd.copy_(torch.tensor(qd, dtype=torch.float32))
        control.joint_target_q.copy_(torch.tensor(tq, dtype=torch.float32))
        control.joint_target_qd.copy_(torch.tensor(tqd, dtype=torch.float32))
        oracle_lib.eval_fk(model, s0.joint_q, s0.joint_qd, s0)
        s0.clear_forces()
        solver.step(s0, s1, control, None, dt)
        return s1.joint_qd.numpy().copy()

    # the signed gravity component along the motion axis is +g (test_joint_drive.py:84: gravity vector = +g * up)
    got = one_step(q0, qd0, pos_t, vel_t)
    for i in range(2):
        assert got[i] == pytest.approx(_expected_qd(g, dt, M[i], pos_t[i], vel_t[i], kes[i], kds[i], q0[i], qd0[i]), abs=2e-3)
    # gains changed in place (the reference calls notify_model_changed; the arrays are read live here)
    kes2, kds2 = [kes[0] * 2.0, kes[1] * 2.5], [kds[0] * 2.75, kds[1] * 3.5]
    model.joint_target_ke.copy_(torch.tensor(kes2, dtype=torch.float32))
    model.joint_target_kd.copy_(torch.tensor(kds2, dtype=torch.float32))
    got = one_step(q0, qd0, pos_t, vel_t)
    for i in range(2):
        assert got[i] == pytest.approx(_expected_qd(g, dt, M[i], pos_t[i], vel_t[i], kes2[i], kds2[i], q0[i], qd0[i]), abs=5e-3)
    # velocity control from rest
    model.joint_target_ke.zero_()
    model.joint_target_kd.copy_(torch.tensor(kds, dtype=torch.float32))
    vt = [20.0, 300.0]
    got = one_step([0.0, 0.0], [0.0, 0.0], [0.0, 0.0], vt)
    for i in range(2):
        assert got[i] == pytest.approx(_expected_qd(g, dt, M[i], 0.0, vt[i], 0.0, kds[i], 0.0, 0.0), abs=1e-4)


# ---- test_pendulum_revolute_vs_d6.py ---------------------------------------------------------------------------------------------
def _pendulum(joint_kind):
    b = ModelBuilder(up_axis="z", gravity=-9.81)
    link = b.add_link(mass=1.0, inertia=np.diag([0.02, 0.02, 0.002]), com=(0.0, 0.0, 0.0))
    pxf, cxf = X.transform((0.0, 0.0, 2.0)), X.transform((0.0, 0.0, 0.5))  # COM half a metre below the pivot
    if joint_kind == "revolute":
        j = b.add_joint_revolute(-1, link, parent_xform=pxf, child_xform=cxf, axis=(0.0, 1.0, 0.0), armature=0.0, limit_ke=0.0, limit_kd=0.0,
                                 target_ke=0.0, target_kd=0.0)
    else:
        j = b.add_joint_d6(-1, link, angular_axes=[JointDofConfig(axis=(0.0, 1.0, 0.0), armature=0.0, limit_ke=0.0, limit_kd=0.0,
                                                                  target_ke=0.0, target_kd=0.0)],
                           parent_xform=pxf, child_xform=cxf)
    b.add_articulation([j])
    b.joint_q[0] = 0.2
    return b.finalize()


@pytest.mark.parametrize("solver_name", ["featherstone", "xpbd"])
def test_pendulum_revolute_vs_d6(oracle_lib, solver_name):
    traj = {}
    for kind in ("revolute", "d6"):
        model = _pendulum(kind)
        assert int(model.joint_type[0]) == (JointType.REVOLUTE if kind == "revolute" else JointType.D6)
        solver = oracle_lib.SolverFeatherstone(model, angular_damping=0.0) if solver_name == "featherstone" else \
            oracle_lib.SolverXPBD(model, iterations=8, angular_damping=0.0)
        s0, s1, control = model.state(), model.state(), model.control()
        oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s0)
        out = np.zeros(480)
        for i in range(480):
            s0.clear_forces()
            solver.step(s0, s1, control, None, 1.0 / 240.0)
            s0, s1 = s1, s0
            if solver_name == "xpbd":
                oracle_lib.eval_ik(model, s0, s0.joint_q, s0.joint_qd)
            out[i] = float(s0.joint_q[0])
        traj[kind] = out
    for kind in traj:  # they moved and oscillated (test_pendulum_revolute_vs_d6.py:90-97)
        assert traj[kind].max() - traj[kind].min() > 0.1
    assert np.mean(np.abs(traj["revolute"] - traj["d6"])) < 0.1  # :100-101; the two joint codes agree far better than that:
    assert np.max(np.abs(traj["revolute"] - traj["d6"])) < 1e-3


# ---- test_ik_fk_kernels.py -------------------------------------------------------------------------------------------------------
def _single_joint_model(kind):
    b = ModelBuilder()
    parent_xf = X.transform((0.1, 0.2, 0.3), X.quat_from_axis_angle((0.0, 1.0, 0.0), 0.0))
    child_xf = X.transform((-0.05, 0.0, 0.0), X.quat_from_axis_angle((1.0, 0.0, 0.0), 0.5))
    child = b.add_link(mass=0.1)
    b.add_shape_box(child, hx=0.05, hy=0.05, hz=0.05)
    kw = dict(parent_xform=parent_xf, child_xform=child_xf)
    if kind == JointType.REVOLUTE:
        j = b.add_joint_revolute(-1, child, axis=(0.0, 0.0, 1.0), **kw)
    elif kind == JointType.PRISMATIC:
        j = b.add_joint_prismatic(-1, child, axis=(1.0, 0.0, 0.0), **kw)
    elif kind == JointType.BALL:
        j = b.add_joint_ball(-1, child, **kw)
    elif kind == JointType.D6:
        j = b.add_joint_d6(-1, child, linear_axes=[JointDofConfig(axis=a) for a in ((1, 0, 0), (0, 1, 0), (0, 0, 1))],
                           angular_axes=[JointDofConfig(axis=a) for a in ((1, 0, 0), (0, 1, 0), (0, 0, 1))], **kw)
    else:
        j = b.add_joint_free(child, **kw)
    b.add_articulation([j])
    return b.finalize()


def _randomize_joint_q(model, kind, seed=0):
    rng = np.random.default_rng(seed)
    q = model.joint_q.numpy().copy()

    def small_quat():
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis) + 1e-8
        angle = rng.uniform(-np.pi / 6, np.pi / 6)
        return (*(axis * np.sin(angle / 2.0)), np.cos(angle / 2.0))

    if kind == JointType.REVOLUTE:
        q[0] = rng.uniform(-np.pi / 2, np.pi / 2)
    elif kind == JointType.PRISMATIC:
        q[0] = rng.uniform(-0.2, 0.2)
    elif kind == JointType.BALL:
        q[0:4] = small_quat()
    elif kind == JointType.D6:
        q[0:3] = rng.uniform(-0.1, 0.1, size=3)
        q[3:6] = rng.uniform(-np.pi / 8, np.pi / 8, size=3)
    else:
        q[0:3] = rng.uniform(-0.3, 0.3, size=3)
        q[3:7] = small_quat()
    model.joint_q.copy_(torch.from_numpy(q.astype(np.float32)))


@pytest.mark.parametrize("kind", [JointType.REVOLUTE, JointType.PRISMATIC, JointType.BALL, JointType.D6, JointType.FREE])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fk_of_every_joint_kind_matches_an_independent_walk_and_ik_inverts_it(oracle_lib, kind, seed):
    model = _single_joint_model(kind)
    _randomize_joint_q(model, kind, seed)
    g = torch.Generator().manual_seed(seed)
    model.joint_qd.copy_(torch.rand(model.joint_qd.shape, generator=g) - 0.5)
    ref = model.state()
    host_fk(model, model.joint_q, model.joint_qd, ref)  # float64 walk
    out = model.state()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, out)
    np.testing.assert_allclose(out.body_q.numpy(), ref.body_q.numpy(), atol=1e-6)  # assert_np_equal(..., tol=1e-6), :214
    if kind != JointType.D6:  # the host walk adds the three D6 rates about the FIXED axes; the reference (and the oracle) about the
        # successively rotated ones (compute_3d_rotational_dofs) - that arithmetic is pinned by finite differences in
        # tests/test_d6_two_angular_axes.py, and by the eval_ik round trip below
        np.testing.assert_allclose(out.body_qd.numpy(), ref.body_qd.numpy(), atol=2e-6)
    # eval_ik recovers the coordinates (quaternion coordinates up to sign)
    q_back, qd_back = torch.zeros_like(model.joint_q), torch.zeros_like(model.joint_qd)
    oracle_lib.eval_ik(model, out, q_back, qd_back)
    q, qb = model.joint_q.numpy(), q_back.numpy()
    if kind == JointType.BALL:
        assert min(np.abs(q - qb).max(), np.abs(q + qb).max()) < 2e-6
    elif kind == JointType.FREE:
        np.testing.assert_allclose(qb[:3], q[:3], atol=2e-6)
        assert min(np.abs(q[3:] - qb[3:]).max(), np.abs(q[3:] + qb[3:]).max()) < 2e-6
    else:
        np.testing.assert_allclose(qb, q, atol=5e-6)
    np.testing.assert_allclose(qd_back.numpy(), model.joint_qd.numpy(), atol=5e-6)

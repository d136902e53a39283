"""D6 joints with exactly TWO angular axes (a universal joint) in SolverFeatherstone / eval_fk.

Reference: `transform_2d_rotational_axes` / `compute_2d_rotational_dofs` (newton/_src/sim/articulation.py:37-82), used by
`jcalc_transform` (featherstone/kernels.py:212-220) and `jcalc_motion` (:301-311, incl. the apparent-derivative term
a0 x a1 * qd0 * qd1).  The reference exercises the path in test_inverse_dynamics.py:2252-2264 ("d6_2ang": axes X, Z); it holds no
closed-form numbers for it, so the oracle is pinned here by the properties that path must satisfy:

* the joint rotation is the intrinsic composition R(axis_0, q0) . R(axis_1, q1),
* the twist `eval_fk` reports is the time derivative of the pose it reports (finite differences),
* a pendulum on the joint conserves energy (wrong motion subspace or a missing apparent-derivative term makes it drift),
* non-orthogonal axes are orthonormalised the way `quat_from_matrix` does it (axis_0 kept, axis_1 projected).

The CUDA path is compared with the oracle bit for bit in the GPU tests at the bottom.
"""

import math

import numpy as np
import pytest
import torch

from newton_b200.sim.builder import JointDofConfig, ModelBuilder
from newton_b200.utils import xform as X


def _f32(a):
    return torch.tensor(np.asarray(a, dtype=np.float32))


def _universal_model(axes=((1.0, 0.0, 0.0), (0.0, 0.0, 1.0)), gravity=0.0, with_tip=True):
    b = ModelBuilder(gravity=gravity)
    link0 = b.add_link(mass=1.3, com=(0.05, -0.02, -0.4), inertia=np.diag([0.3, 0.25, 0.08]))
    b.add_shape_box(link0, hx=0.05, hy=0.05, hz=0.4)
    j0 = b.add_joint_d6(parent=-1, child=link0, parent_xform=X.transform((0.0, 0.0, 2.0), X.quat_from_axis_angle((0.0, 1.0, 0.0), 0.3)),
                        child_xform=X.transform((0.02, 0.0, 0.1)), angular_axes=[JointDofConfig.create_unlimited(a) for a in axes])
    joints = [j0]
    links = [link0]
    if with_tip:
        link1 = b.add_link(mass=0.4, com=(0.0, 0.03, -0.2), inertia=np.diag([0.02, 0.02, 0.004]))
        b.add_shape_box(link1, hx=0.03, hy=0.03, hz=0.2)
        joints.append(b.add_joint_revolute(parent=link0, child=link1, axis=(0.0, 1.0, 0.0), parent_xform=X.transform((0.0, 0.0, -0.8)),
                                           child_xform=X.transform((0.0, 0.0, 0.0))))
        links.append(link1)
    b.add_articulation(joints)
    return b.finalize(), links


def test_joint_rotation_is_the_intrinsic_x_then_z_composition(oracle_lib):
    model, (link0,) = _universal_model(with_tip=False)
    q0, q1 = 0.7, -1.1
    s = model.state()
    s.joint_q.copy_(_f32([q0, q1]))
    oracle_lib.eval_fk(model, s.joint_q, s.joint_qd, s)
    Xp = model.joint_X_p.numpy()[0].astype(np.float64)
    Xc = model.joint_X_c.numpy()[0].astype(np.float64)
    # R(x, q0) . R(z, q1): the second axis rides on the first rotation
    rot = X.quat_mul(X.quat_from_axis_angle((1.0, 0.0, 0.0), q0), X.quat_from_axis_angle((0.0, 0.0, 1.0), q1))
    expect = X.transform_mul(X.transform_mul(Xp, np.array([0.0, 0.0, 0.0, *rot])), X.transform_inverse(Xc))
    got = s.body_q.numpy()[link0].astype(np.float64)
    np.testing.assert_allclose(got[:3], expect[:3], atol=2e-6)
    assert abs(abs(float(np.dot(got[3:], expect[3:]))) - 1.0) < 1e-6


def test_non_orthogonal_axes_are_orthonormalised(oracle_lib):
    """axis_1 with a component along axis_0: q_off = quat_from_matrix([a0 a1 a0 x a1]) keeps a valid rotation only approximately;
    for orthogonal-but-permuted axes (Z, X) the result must be R(z, q0) . R(x, q1) exactly like the (X, Z) case."""
    model, (link0,) = _universal_model(axes=((0.0, 0.0, 1.0), (1.0, 0.0, 0.0)), with_tip=False)
    q0, q1 = -0.4, 0.9
    s = model.state()
    s.joint_q.copy_(_f32([q0, q1]))
    oracle_lib.eval_fk(model, s.joint_q, s.joint_qd, s)
    Xp = model.joint_X_p.numpy()[0].astype(np.float64)
    Xc = model.joint_X_c.numpy()[0].astype(np.float64)
    rot = X.quat_mul(X.quat_from_axis_angle((0.0, 0.0, 1.0), q0), X.quat_from_axis_angle((1.0, 0.0, 0.0), q1))
    expect = X.transform_mul(X.transform_mul(Xp, np.array([0.0, 0.0, 0.0, *rot])), X.transform_inverse(Xc))
    got = s.body_q.numpy()[link0].astype(np.float64)
    np.testing.assert_allclose(got[:3], expect[:3], atol=2e-6)
    assert abs(abs(float(np.dot(got[3:], expect[3:]))) - 1.0) < 1e-6


def _ang_vel_fd(qa, qb, dt):
    """world angular velocity from two successive orientations (right-trivialised)"""
    dq = X.quat_mul(qb, X.quat_inverse(qa))
    if dq[3] < 0:
        dq = -dq
    return 2.0 * dq[:3] / dt


def test_fk_twist_is_the_derivative_of_the_fk_pose(oracle_lib):
    """test_kinematics.py:191-257 style finite-difference check, on the universal joint and on its descendant."""
    model, links = _universal_model()
    q = np.array([0.6, -0.8, 0.5], dtype=np.float64)
    qd = np.array([1.2, -0.7, 0.9], dtype=np.float64)
    dt = 1.0e-3
    sa, sb, s0 = model.state(), model.state(), model.state()
    # central difference (in float32 the forward difference at 1e-4 is noise-limited)
    for s, qq in ((sa, q - 0.5 * dt * qd), (sb, q + 0.5 * dt * qd), (s0, q)):
        s.joint_q.copy_(_f32(qq))
        s.joint_qd.copy_(_f32(qd))
        oracle_lib.eval_fk(model, s.joint_q, s.joint_qd, s)
    com = model.body_com.numpy().astype(np.float64)
    for link in links:
        a, b, c = (s.body_q.numpy()[link].astype(np.float64) for s in (sa, sb, s0))
        twist = s0.body_qd.numpy()[link].astype(np.float64)
        w_fd = _ang_vel_fd(a[3:], b[3:], dt)
        np.testing.assert_allclose(twist[3:], w_fd, atol=3e-3)
        com_a, com_b = a[:3] + X.quat_rotate(a[3:], com[link]), b[:3] + X.quat_rotate(b[3:], com[link])
        np.testing.assert_allclose(twist[:3], (com_b - com_a) / dt, atol=3e-3)  # body_qd is the COM twist


def _energy(model, state, g):
    bq, bqd = state.body_q.numpy().astype(np.float64), state.body_qd.numpy().astype(np.float64)
    m, I, com = model.body_mass.numpy().astype(np.float64), model.body_inertia.numpy().astype(np.float64), model.body_com.numpy().astype(np.float64)
    e = 0.0
    for i in range(model.body_count):
        R = X.quat_to_matrix(bq[i, 3:])
        Iw = R @ I[i] @ R.T
        c = bq[i, :3] + R @ com[i]
        e += 0.5 * m[i] * bqd[i, :3] @ bqd[i, :3] + 0.5 * bqd[i, 3:] @ Iw @ bqd[i, 3:] + m[i] * g * c[2]
    return e


@pytest.mark.parametrize("gravity", [0.0, -9.81])
def test_universal_joint_pendulum_conserves_energy(oracle_lib, gravity):
    model, _ = _universal_model(gravity=gravity)
    s0, s1 = model.state(), model.state()
    s0.joint_q.copy_(_f32([0.5, 0.4, -0.3]))
    s0.joint_qd.copy_(_f32([1.5, -2.0, 1.0]))
    oracle_lib.eval_fk(model, s0.joint_q, s0.joint_qd, s0)
    e0 = _energy(model, s0, -gravity)
    ke0 = _energy(model, s0, 0.0)
    solver = oracle_lib.SolverFeatherstone(model, angular_damping=0.0)
    dt, q_start = 2.0e-4, s0.joint_q.numpy().copy()
    for _ in range(1500):  # 0.3 s
        s0.clear_forces()
        solver.step(s0, s1, None, None, dt)
        s0, s1 = s1, s0
    assert np.abs(s0.joint_q.numpy() - q_start).max() > 0.2  # it moved
    # measured drift: 3e-6 J without gravity, 3e-3 J with it (kinetic energy 3.3 J; semi-implicit Euler at dt = 0.2 ms)
    assert abs(_energy(model, s0, -gravity) - e0) < (1e-4 if gravity == 0.0 else 0.01) * ke0
    # the solver's joint_q / joint_qd and its body state stay consistent with eval_fk
    chk = model.state()
    oracle_lib.eval_fk(model, s0.joint_q, s0.joint_qd, chk)
    np.testing.assert_allclose(chk.body_q.numpy(), s0.body_q.numpy(), atol=1e-5)
    np.testing.assert_allclose(chk.body_qd.numpy(), s0.body_qd.numpy(), atol=1e-4)


# ---- eval_ik on multi-axis D6 joints (invert_2d / invert_3d_rotational_dofs, sim/articulation.py:85-126, 177-236) ---------------------


def _d6_model(axes, parent_rot=None):
    b = ModelBuilder(gravity=0.0)
    child = b.add_link(mass=1.0, inertia=np.eye(3) * 0.1)
    px = X.transform((0.1, -0.2, 0.3), parent_rot) if parent_rot is not None else None
    j = b.add_joint_d6(parent=-1, child=child, parent_xform=px, angular_axes=[JointDofConfig.create_unlimited(a) for a in axes])
    b.add_articulation([j])
    return b.finalize(), child


def test_fk_ik_d6_left_handed_angular_axes(oracle_lib):
    """newton/tests/test_kinematics.py:1057-1104, verbatim numbers: axes (X, Z, Y) form a left-handed triple; the body rotation is the
    intrinsic product qfa(X, q0) qfa(Z, q1) qfa(Y, q2) (1e-6), the angular velocity is the transported-axes sum (1e-6), and
    eval_ik recovers joint_q and joint_qd (1e-6)."""
    model, child = _d6_model(((1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0)))
    q_vals, qd_vals = np.array([0.5, -0.4, 0.7], np.float32), np.array([0.9, -0.6, 0.3], np.float32)
    state = model.state()
    state.joint_q.copy_(_f32(q_vals))
    state.joint_qd.copy_(_f32(qd_vals))
    oracle_lib.eval_fk(model, state.joint_q, state.joint_qd, state)
    rot = state.body_q.numpy()[child][3:].astype(np.float64)
    qa = X.quat_from_axis_angle
    expected = X.quat_mul(X.quat_mul(qa((1, 0, 0), 0.5), qa((0, 0, 1), -0.4)), qa((0, 1, 0), 0.7))
    np.testing.assert_allclose(rot, expected, atol=1e-6)
    q_0 = qa((1, 0, 0), 0.5)
    axis_1_w = X.quat_rotate(q_0, np.array([0.0, 0.0, 1.0]))
    q_1 = qa(axis_1_w, -0.4)
    axis_2_w = X.quat_rotate(X.quat_mul(q_1, q_0), np.array([0.0, 1.0, 0.0]))
    expected_w = np.array([1.0, 0.0, 0.0]) * 0.9 + axis_1_w * -0.6 + axis_2_w * 0.3
    np.testing.assert_allclose(state.body_qd.numpy()[child][3:], expected_w, atol=1e-6)
    q_ik, qd_ik = torch.zeros_like(model.joint_q), torch.zeros_like(model.joint_qd)
    oracle_lib.eval_ik(model, state, q_ik, qd_ik)
    np.testing.assert_allclose(q_ik.numpy(), q_vals, atol=1e-6)
    np.testing.assert_allclose(qd_ik.numpy(), qd_vals, atol=1e-6)


@pytest.mark.parametrize("axes", [((1.0, 0.0, 0.0), (0.0, 0.0, 1.0)), ((0.0, 0.0, 1.0), (1.0, 0.0, 0.0)), ((0.0, 1.0, 0.0), (0.0, 0.0, 1.0)),
                                  ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)), ((0.0, 0.0, 1.0), (0.0, 1.0, 0.0), (1.0, 0.0, 0.0))])
def test_ik_inverts_fk_on_multi_axis_d6(oracle_lib, axes):
    """eval_ik(eval_fk(q, qd)) == (q, qd) for two- and three-axis D6 joints under a rotated, offset parent frame"""
    model, _ = _d6_model(axes, parent_rot=X.quat_from_axis_angle((0.3, -0.5, 0.8), 0.9))
    rng = np.random.default_rng(len(axes))
    for _ in range(8):
        q = rng.uniform(-1.2, 1.2, len(axes)).astype(np.float32)
        qd = rng.uniform(-2.0, 2.0, len(axes)).astype(np.float32)
        state = model.state()
        state.joint_q.copy_(_f32(q))
        state.joint_qd.copy_(_f32(qd))
        oracle_lib.eval_fk(model, state.joint_q, state.joint_qd, state)
        q_ik, qd_ik = torch.zeros_like(model.joint_q), torch.zeros_like(model.joint_qd)
        oracle_lib.eval_ik(model, state, q_ik, qd_ik)
        np.testing.assert_allclose(q_ik.numpy(), q, atol=2e-6)
        np.testing.assert_allclose(qd_ik.numpy(), qd, atol=1e-5)


# ---- GPU parity -------------------------------------------------------------------------------------------------------------------


@pytest.mark.gpu
def test_gpu_eval_ik_multi_axis_d6_bit_exact(oracle_lib, cuda_lib):
    import newton_b200

    for axes in (((1.0, 0.0, 0.0), (0.0, 0.0, 1.0)), ((1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0)), ((0.0, 0.0, 1.0), (0.0, 1.0, 0.0), (1.0, 0.0, 0.0))):
        model, _ = _d6_model(axes, parent_rot=X.quat_from_axis_angle((0.3, -0.5, 0.8), 0.9))
        rng = np.random.default_rng(7 + len(axes))
        mg = model.to("cuda:0")
        for _ in range(6):
            q = rng.uniform(-1.2, 1.2, len(axes)).astype(np.float32)
            qd = rng.uniform(-2.0, 2.0, len(axes)).astype(np.float32)
            state = model.state()
            state.joint_q.copy_(_f32(q))
            state.joint_qd.copy_(_f32(qd))
            oracle_lib.eval_fk(model, state.joint_q, state.joint_qd, state)
            q_ref, qd_ref = torch.zeros_like(model.joint_q), torch.zeros_like(model.joint_qd)
            oracle_lib.eval_ik(model, state, q_ref, qd_ref)
            sg = mg.state()
            sg.body_q.copy_(state.body_q)
            sg.body_qd.copy_(state.body_qd)
            q_g, qd_g = torch.zeros_like(mg.joint_q), torch.zeros_like(mg.joint_qd)
            newton_b200.eval_ik(mg, sg, q_g, qd_g)
            np.testing.assert_array_equal(q_g.cpu().numpy(), q_ref.numpy())
            np.testing.assert_array_equal(qd_g.cpu().numpy(), qd_ref.numpy())


@pytest.mark.gpu
def test_gpu_universal_joint_bit_exact(oracle_lib, cuda_lib):
    import newton_b200
    from tests.helpers import simulate

    model, _ = _universal_model(gravity=-9.81)
    model.joint_q.copy_(_f32([0.5, 0.4, -0.3]))
    model.joint_qd.copy_(_f32([1.5, -2.0, 1.0]))
    g = torch.Generator().manual_seed(11)
    model.joint_f.copy_(torch.rand(model.joint_f.shape, generator=g) * 0.4 - 0.2)
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, model)
    kw = {"angular_damping": 0.0}
    rs, _, _ = simulate(model, oracle_lib.CollisionPipeline, oracle_lib.SolverFeatherstone, substeps=200, dt=1e-3, solver_kwargs=kw, collide=False)
    mg = model.to("cuda:0")
    sg = mg.state()
    newton_b200.eval_fk(mg, mg.joint_q, mg.joint_qd, sg)
    np.testing.assert_array_equal(sg.body_q.cpu().numpy(), model.body_q.numpy())
    np.testing.assert_array_equal(sg.body_qd.cpu().numpy(), model.body_qd.numpy())
    gs, _, _ = simulate(mg, newton_b200.CollisionPipeline, newton_b200.solvers.SolverFeatherstone, substeps=200, dt=1e-3, solver_kwargs=kw,
                        collide=False)
    torch.cuda.synchronize()
    for name in ("joint_q", "joint_qd", "body_q", "body_qd"):
        a, b = getattr(gs, name).cpu().numpy(), getattr(rs, name).numpy()
        assert np.isfinite(b).all(), name
        np.testing.assert_array_equal(a, b, err_msg=name)

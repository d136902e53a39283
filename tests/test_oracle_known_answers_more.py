"""More of the reference's own closed-form / known-answer checks for the hot path, run on the CPU oracle.

* ``newton/tests/test_up_axis.py``          - one step of free fall along the configured up axis, both solvers
* ``newton/tests/test_control_force.py``    - ``Control.joint_f`` on a FREE root, on a 3-prismatic D6 joint, with a child-frame
                                              moment arm (XPBD ``apply_joint_forces``)
* ``newton/tests/test_runtime_gravity.py``  - ``Model.set_gravity`` (all worlds / one world / array / global slot / implicit
                                              world / invalid world), per-world gravity in the solvers, global-world bodies
"""

import numpy as np
import pytest
import torch

from newton_b200 import ModelFlags
from newton_b200.sim.builder import JointDofConfig, ModelBuilder
from newton_b200.utils import xform as X
from newton_b200.utils.host_fk import host_fk


def _solver(oracle_lib, name, model):
    cls = oracle_lib.SolverFeatherstone if name == "featherstone" else oracle_lib.SolverXPBD
    return cls(model, angular_damping=0.0)


def _quat_between_z_and(axis):
    return {"y": X.quat_from_axis_angle(np.array([1.0, 0.0, 0.0]), -np.pi / 2), "z": np.array([0.0, 0.0, 0.0, 1.0])}[axis]


def _quat_rpy(r, p, y):
    """wp.quat_rpy (roll about x, pitch about y, yaw about z)."""
    cy, sy, cr, sr, cp, sp = np.cos(y * 0.5), np.sin(y * 0.5), np.cos(r * 0.5), np.sin(r * 0.5), np.cos(p * 0.5), np.sin(p * 0.5)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy])


# ---- test_up_axis.py:16-37 -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver_name", ["featherstone", "xpbd"])
@pytest.mark.parametrize("axis", ["y", "z"])
def test_gravity_follows_up_axis(oracle_lib, solver_name, axis):
    g = [0.0, 0.0, 0.0]
    g["xyz".index(axis)] = -9.81
    builder = ModelBuilder(up_axis=axis, gravity=tuple(g))
    b = builder.add_body()
    builder.add_shape_capsule(b, xform=X.transform((0.0, 0.0, 0.0), _quat_between_z_and(axis)))
    model = builder.finalize()
    s0, s1 = model.state(), model.state()
    _solver(oracle_lib, solver_name, model).step(s0, s1, model.control(), None, 1.0 / 10.0)
    assert s1.body_qd.numpy()[0, "xyz".index(axis)] == pytest.approx(-0.981, abs=1e-5)


# ---- test_control_force.py:48-84 -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver_name", ["featherstone", "xpbd"])
def test_control_force_floating_body_angular(oracle_lib, solver_name):
    builder = ModelBuilder(up_axis="y", gravity=(0.0, 0.0, 0.0))
    b = builder.add_body()
    builder.add_shape_box(b)
    builder.joint_q = [1.0, 2.0, 3.0, *_quat_rpy(-1.3, 0.8, 2.4)]
    model = builder.finalize()
    solver = _solver(oracle_lib, solver_name, model)
    s0, s1 = model.state(), model.state()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s0)
    control = model.control()
    control.joint_f.copy_(torch.tensor([0.0, 0.0, 0.0, 0.0, 0.0, 100.0]))
    for _ in range(4):
        solver.step(s0, s1, control, None, 1.0 / 10.0)
        s0, s1 = s1, s0
    qd = s0.body_qd.numpy()[0]
    assert 0.04 < qd[5] < 0.4
    assert np.abs(qd[:5]).max() <= 2e-6


# ---- test_control_force.py:87-136 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver_name", ["featherstone", "xpbd"])
def test_control_force_3d_articulation(oracle_lib, solver_name):
    builder = ModelBuilder(gravity=(0.0, 0.0, 0.0))
    builder.default_shape_cfg.density = 100.0
    b = builder.add_link()
    builder.add_shape_sphere(b)
    j = builder.add_joint_d6(-1, b, linear_axes=[JointDofConfig(axis=a, armature=0.0) for a in ((1, 0, 0), (0, 1, 0), (0, 0, 1))])
    builder.add_articulation([j])
    model = builder.finalize()
    assert model.joint_dof_count == 3
    for dim in range(3):
        solver = _solver(oracle_lib, solver_name, model)
        s0, s1 = model.state(), model.state()
        control = model.control()
        f = np.zeros(3, dtype=np.float32)
        f[dim] = 100.0
        control.joint_f.copy_(torch.from_numpy(f))
        for _ in range(4):
            solver.step(s0, s1, control, None, 1.0 / 10.0)
            s0, s1 = s1, s0
        if solver_name == "xpbd":  # maximal-coordinate solver: recover joint_qd from body_qd
            assert oracle_lib.eval_ik(model, s0, s0.joint_q, s0.joint_qd) in (0, None)
        qd = s0.joint_qd.numpy()
        assert 0.009 < qd[dim] < 0.4
        assert np.abs(np.delete(qd, dim)).max() <= 1e-6


# ---- test_control_force.py:139-197 (regression of the reference's issue #1261) ---------------------------------------------
def test_control_force_child_xform_moment_arm_xpbd(oracle_lib):
    offset_y = 2.0
    builder = ModelBuilder(gravity=(0.0, 0.0, 0.0))
    builder.default_shape_cfg.density = 100.0
    b = builder.add_link()
    builder.add_shape_sphere(b)
    axes = ((1, 0, 0), (0, 1, 0), (0, 0, 1))
    j = builder.add_joint_d6(-1, b, child_xform=X.transform((0.0, offset_y, 0.0)),
                             linear_axes=[JointDofConfig(axis=a, armature=0.0) for a in axes],
                             angular_axes=[JointDofConfig(axis=a, armature=0.0) for a in axes])
    builder.add_articulation([j])
    model = builder.finalize()
    solver = _solver(oracle_lib, "xpbd", model)
    s0, s1 = model.state(), model.state()
    control = model.control()
    f = np.zeros(model.joint_dof_count, dtype=np.float32)
    f[0] = 100.0  # force along X at the joint anchor, 2 m above the COM
    control.joint_f.copy_(torch.from_numpy(f))
    for _ in range(4):
        solver.step(s0, s1, control, None, 1.0 / 10.0)
        s0, s1 = s1, s0
    qd = s0.body_qd.numpy()[0]
    assert qd[0] > 0.001
    assert qd[5] < -0.001  # cross((0, offset_y, 0), (F, 0, 0)) = (0, 0, -F offset_y)


# ---- test_runtime_gravity.py -----------------------------------------------------------------------------------------------
def _box_world():
    w = ModelBuilder(gravity=(0.0, 0.0, -9.81))
    w.default_shape_cfg.density = 1000.0
    b = w.add_body()
    w.add_shape_box(b, hx=0.5, hy=0.5, hz=0.5)
    return w


def _run(solver, model, s0, s1, n, dt=0.01):
    control = model.control()
    for _ in range(n):
        s0.clear_forces()
        solver.step(s0, s1, control, None, dt)
        s0, s1 = s1, s0
    return s0, s1


def test_runtime_gravity_bodies_xpbd(oracle_lib):
    """:124-163"""
    model = _box_world().finalize()
    solver = oracle_lib.SolverXPBD(model)
    s0, s1 = _run(solver, model, model.state(), model.state(), 10)
    assert s0.body_qd.numpy()[0, 2] < -0.5
    model.set_gravity((9.81, 0.0, 0.0))
    solver.notify_model_changed(ModelFlags.MODEL_PROPERTIES)
    s0, s1 = _run(solver, model, s0, s1, 20)
    assert s0.body_qd.numpy()[0, 0] > 0.5


@pytest.mark.parametrize("solver_name", ["xpbd", "featherstone"])
def test_per_world_gravity_bodies(oracle_lib, solver_name):
    """:266-318 (the reference runs it with XPBD; the Featherstone oracle reads the same gravity[body_world])."""
    main = ModelBuilder(gravity=(0.0, 0.0, -9.81))
    main.replicate(_box_world(), 3)
    model = main.finalize()
    solver = oracle_lib.SolverXPBD(model) if solver_name == "xpbd" else oracle_lib.SolverFeatherstone(model, angular_damping=0.0)
    assert model.gravity.shape[0] == 4
    model.set_gravity((0.0, 0.0, 0.0), world=0)
    model.set_gravity((0.0, 0.0, -4.905), world=1)
    model.set_gravity((0.0, 0.0, -9.81), world=2)
    solver.notify_model_changed(ModelFlags.MODEL_PROPERTIES)
    s0, _ = _run(solver, model, model.state(), model.state(), 10)
    vz = s0.body_qd.numpy()[:, 2]
    assert vz[0] == pytest.approx(0.0, abs=5e-5)
    assert vz[2] < vz[1] < 0.0 and vz[2] < -0.5
    assert vz[1] == pytest.approx(-0.4905, abs=1e-5) and vz[2] == pytest.approx(-0.981, abs=1e-5)


def test_global_gravity_bodies_xpbd(oracle_lib):
    """:362-381 - a body of the global world -1 falls with the LAST gravity entry (negative-index wrap)."""
    builder = ModelBuilder(gravity=(0.0, 0.0, -2.0))
    global_body = builder.add_body(mass=1.0, inertia=np.eye(3))
    builder.begin_world(gravity=(0.0, 0.0, -5.0))
    local_body = builder.add_body(mass=1.0, inertia=np.eye(3))
    builder.end_world()
    model = builder.finalize()
    assert model.numpy("body_world").tolist() == [-1, 0]
    np.testing.assert_allclose(model.numpy("gravity"), ((0.0, 0.0, -5.0), (0.0, 0.0, -2.0)), atol=1e-6)
    s_in, s_out = model.state(), model.state()
    oracle_lib.SolverXPBD(model).step(s_in, s_out, model.control(), None, 0.1)
    qd = s_out.body_qd.numpy()
    assert qd[global_body, 2] == pytest.approx(-0.2, abs=1e-6)
    assert qd[local_body, 2] == pytest.approx(-0.5, abs=1e-6)


def test_set_gravity_spellings():
    """:461-489, :513-538, :559-579, :582-595, :603-613 (bodies instead of particles)."""
    builder = ModelBuilder(gravity=(0.0, 0.0, -9.81))
    for _ in range(2):
        builder.begin_world()
        builder.add_body(mass=1.0, inertia=np.eye(3))
        builder.end_world()
    model = builder.finalize()
    g = model.numpy("gravity")
    assert len(g) == 3 and np.allclose(g[:, 2], -9.81, atol=1e-4)
    ptr = model.gravity.data_ptr()
    model.set_gravity((0.0, 0.0, 0.0), world=0)
    assert model.gravity.data_ptr() == ptr  # updated in place: the native model keeps borrowing the same array
    g = model.numpy("gravity")
    assert g[0, 2] == 0.0 and g[1, 2] == pytest.approx(-9.81, abs=1e-4) and g[-1, 2] == pytest.approx(-9.81, abs=1e-4)
    curriculum = np.array([[0.0, 0.0, s * -9.81] for s in np.linspace(0.0, 1.0, 2)], dtype=np.float32)
    model.set_gravity(curriculum)
    g = model.numpy("gravity")
    np.testing.assert_allclose(g[:2], curriculum, atol=1e-6)
    assert g[-1, 2] == pytest.approx(-9.81, abs=1e-4)  # local-world-only input keeps the global entry

    builder = ModelBuilder(gravity=(0.0, 0.0, -9.81))
    builder.begin_world(gravity=(0.0, 0.0, -1.0))
    builder.add_body(mass=1.0, inertia=np.eye(3))
    builder.end_world()
    model = builder.finalize()
    np.testing.assert_allclose(model.numpy("gravity"), ((0.0, 0.0, -1.0), (0.0, 0.0, -9.81)), atol=1e-6)
    model.set_gravity((0.0, 0.0, -2.0), world=-1)
    np.testing.assert_allclose(model.numpy("gravity"), ((0.0, 0.0, -1.0), (0.0, 0.0, -2.0)), atol=1e-6)
    model.set_gravity(np.array(((0.0, 0.0, -3.0),), dtype=np.float32))
    np.testing.assert_allclose(model.numpy("gravity"), ((0.0, 0.0, -3.0), (0.0, 0.0, -2.0)), atol=1e-6)
    model.set_gravity(np.array(((0.0, 0.0, -4.0), (0.0, 0.0, -5.0)), dtype=np.float32))
    np.testing.assert_allclose(model.numpy("gravity"), ((0.0, 0.0, -4.0), (0.0, 0.0, -5.0)), atol=1e-6)
    model.set_gravity((0.0, 0.0, -6.0))
    np.testing.assert_allclose(model.numpy("gravity"), ((0.0, 0.0, -6.0), (0.0, 0.0, -6.0)), atol=1e-6)
    with pytest.raises(IndexError):
        model.set_gravity((0.0, 0.0, 0.0), world=1)
    with pytest.raises(IndexError):
        model.set_gravity((0.0, 0.0, 0.0), world=-2)
    with pytest.raises(ValueError):
        model.set_gravity((0.0, 0.0), world=0)
    with pytest.raises(ValueError):
        model.set_gravity(np.zeros((5, 3)))

    implicit = ModelBuilder()
    implicit.add_body(mass=1.0, inertia=np.eye(3))
    model = implicit.finalize()
    assert model.world_count == 1 and tuple(model.gravity.shape) == (1, 3)
    model.set_gravity((0.0, 0.0, -2.0), world=0)
    np.testing.assert_allclose(model.numpy("gravity"), ((0.0, 0.0, -2.0),), atol=1e-6)
    model.set_gravity(np.array(((0.0, 0.0, -3.0),), dtype=np.float32))
    np.testing.assert_allclose(model.numpy("gravity"), ((0.0, 0.0, -3.0),), atol=1e-6)


def test_implicit_world_gravity_drives_the_solver(oracle_lib):
    """:582-600 with a body: gravity[0] is shared by world 0 and the global world of an implicit single-world model."""
    implicit = ModelBuilder()
    implicit.add_body(mass=1.0, inertia=np.eye(3))
    model = implicit.finalize()
    model.set_gravity((0.0, 0.0, -4.0))
    s_in, s_out = model.state(), model.state()
    oracle_lib.SolverXPBD(model).step(s_in, s_out, model.control(), None, 0.1)
    assert s_out.body_qd.numpy()[0, 2] == pytest.approx(-0.4, abs=1e-6)


# ---- test_rigid_friction_ramp.py (XPBD rows of its solver matrix: SolverXPBD(iterations=10)) ------------------------------------
import math  # noqa: E402

import newton_b200  # noqa: E402

SIM_DT, SIM_SUBSTEPS = 1.0 / 60.0, 30


def _simulate_frames(oracle_lib, solver, pipe, contacts, model, s0, s1, frames):
    control = model.control()
    for _ in range(frames * SIM_SUBSTEPS):
        s0.clear_forces()
        pipe.collide(s0, contacts)
        solver.step(s0, s1, control, contacts, SIM_DT / SIM_SUBSTEPS)
        s0, s1 = s1, s0
    return s0, s1


def test_xpbd_friction_ramp_grid(oracle_lib):
    """:113-223 - a 5 x 5 grid of slabs on tilted ramps, mu in {0.1 .. 1.0} x angle in {3 .. 40 deg}: below the critical angle
    atan(mu) (minus 2 deg) a slab stays put (|v| < 0.1 m/s, drift < 2 cm over 0.25 s after settling), above it (plus 2 deg) it
    slides (>= 2 cm).  Box-on-box contacts: MPR manifolds + the XPBD friction projection."""
    mus, angles = (0.10, 0.30, 0.50, 0.70, 1.00), (3.0, 10.0, 20.0, 30.0, 40.0)
    builder = ModelBuilder(up_axis="z", gravity=(0.0, 0.0, -9.81))
    box_ids = []
    for row, mu in enumerate(mus):
        ids = []
        for col, angle_deg in enumerate(angles):
            cfg = newton_b200.ShapeConfig()
            cfg.collision_group = col + len(mus) * row + 1
            cfg.mu, cfg.ke, cfg.kd, cfg.kf, cfg.gap = mu, 1.0e5, 1.0e3, 0.0, 0.0
            q = X.quat_from_axis_angle(np.array([1.0, 0.0, 0.0]), math.radians(angle_deg))
            center = np.array([col * 2.5, row * 6.0, 2.0])
            builder.add_shape_box(-1, xform=X.transform(center, q), hx=0.5, hy=2.5, hz=0.05, cfg=cfg)
            up = X.quat_rotate(q, np.array([0.0, 0.0, 1.0]))
            body = builder.add_body(xform=X.transform(center + (0.05 + 0.05 + 0.001) * up, q), label=f"box_r{row}_c{col}")
            builder.add_shape_box(body, hx=0.2, hy=0.2, hz=0.05, cfg=cfg)
            ids.append(body)
        box_ids.append(ids)
    model = builder.finalize()
    solver = oracle_lib.SolverXPBD(model, iterations=10)
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    s0, s1 = _simulate_frames(oracle_lib, solver, pipe, contacts, model, model.state(), model.state(), 30)
    settle_q = s0.body_q.numpy().copy()
    s0, s1 = _simulate_frames(oracle_lib, solver, pipe, contacts, model, s0, s1, 15)
    final_q, final_qd = s0.body_q.numpy(), s0.body_qd.numpy()
    assert np.isfinite(final_q).all() and np.isfinite(final_qd).all()
    failures, static_cells, sliding_cells = [], 0, 0
    for row, mu in enumerate(mus):
        crit = math.degrees(math.atan(mu))
        for col, theta in enumerate(angles):
            bid = box_ids[row][col]
            v = float(np.linalg.norm(final_qd[bid, :3]))
            disp = float(np.linalg.norm(final_q[bid, :3] - settle_q[bid, :3]))
            tag = f"(mu={mu:.2f}, theta={theta:.1f}, crit={crit:.1f})"
            if theta < crit - 2.0:
                static_cells += 1
                if v >= 0.10 or disp >= 0.02:
                    failures.append(f"{tag}: expected static, |v|={v:.4f} disp={disp:.4f}")
            elif theta > crit + 2.0:
                sliding_cells += 1
                if disp < 0.02:
                    failures.append(f"{tag}: expected sliding, disp={disp:.4f}")
    assert static_cells == 15 and sliding_cells == 10  # no cell of this grid falls inside the +-2 deg dead band
    assert not failures, "\n".join(failures)


def test_xpbd_friction_stopping_distance(oracle_lib):
    """:226-367 - a box launched at v0 = 2 m/s on a patch of matching mu stops after d = v0^2 / (2 mu g), within 1 %, and then
    stays at rest (drift < 0.05 m/s)."""
    mus, v0, g = (0.20, 0.40, 0.70), 2.0, 9.81
    builder = ModelBuilder(up_axis="z", gravity=(0.0, 0.0, -g))
    box_ids = []
    for i, mu in enumerate(mus):
        cfg = newton_b200.ShapeConfig()
        cfg.collision_group = i + 1
        cfg.mu, cfg.ke, cfg.kd, cfg.kf, cfg.gap = mu, 1.0e5, 0.0, 0.0, 0.0
        y = i * 5.0
        builder.add_shape_box(-1, xform=X.transform((5.0 - 0.5, y, -0.05)), hx=5.0, hy=0.6, hz=0.05, cfg=cfg)
        body = builder.add_body(xform=X.transform((0.0, y, 0.25 + 0.001)), label=f"box_mu{mu:.2f}")
        builder.add_shape_box(body, hx=0.25, hy=0.25, hz=0.25, cfg=cfg)
        box_ids.append(body)
    model = builder.finalize()
    solver = oracle_lib.SolverXPBD(model, iterations=10)
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    s0, s1 = _simulate_frames(oracle_lib, solver, pipe, contacts, model, model.state(), model.state(), 30)
    initial_q = s0.body_q.numpy().copy()
    s0.body_qd[box_ids] = torch.tensor([v0, 0.0, 0.0, 0.0, 0.0, 0.0])
    frames = int(math.ceil(1.5 * v0 / (min(mus) * g) / SIM_DT))
    s0, s1 = _simulate_frames(oracle_lib, solver, pipe, contacts, model, s0, s1, frames)
    final_q = s0.body_q.numpy().copy()
    s0, s1 = _simulate_frames(oracle_lib, solver, pipe, contacts, model, s0, s1, 15)
    rest_q = s0.body_q.numpy()
    assert np.isfinite(final_q).all() and np.isfinite(rest_q).all()
    for bid, mu in zip(box_ids, mus):
        d_expected = v0 * v0 / (2.0 * mu * g)
        d = float(np.linalg.norm(final_q[bid, :2] - initial_q[bid, :2]))
        assert abs(d - d_expected) / d_expected <= 0.01, (mu, d, d_expected)
        assert float(np.linalg.norm(rest_q[bid, :2] - final_q[bid, :2])) / (15 * SIM_DT) < 0.05, mu


# ---- test_joint_controllers.py:35-101, 627-665 (revolute PD targets; XPBD(iterations=5) and Featherstone rows) --------------------
@pytest.mark.parametrize("solver_name", ["featherstone", "xpbd"])
@pytest.mark.parametrize("pos_target,vel_target,expected_pos,expected_vel,ke,kd",
                         [(math.pi / 2.0, 0.0, math.pi / 2.0, 0.0, 2000.0, 500.0), (0.0, math.pi / 2.0, None, math.pi / 2.0, 0.0, 500.0)])
def test_revolute_joint_controller(oracle_lib, solver_name, pos_target, vel_target, expected_pos, expected_vel, ke, kd):
    builder = ModelBuilder(up_axis="y", gravity=(0.0, 0.0, 0.0))
    b = builder.add_link(inertia=np.eye(3), mass=1.0)
    cfg = newton_b200.ShapeConfig()
    cfg.density = 1.0
    builder.add_shape_box(b, hx=0.2, hy=0.2, hz=0.2, cfg=cfg)
    j = builder.add_joint_revolute(-1, b, parent_xform=X.transform((0.0, 2.0, 0.0)), child_xform=X.transform((0.0, 2.0, 0.0)), axis=(0.0, 0.0, 1.0),
                                   target_pos=pos_target, target_vel=vel_target, armature=0.0, limit_ke=0.0, limit_kd=0.0, target_ke=ke, target_kd=kd)
    builder.add_articulation([j])
    model = builder.finalize()
    solver = oracle_lib.SolverFeatherstone(model, angular_damping=0.0) if solver_name == "featherstone" else \
        oracle_lib.SolverXPBD(model, angular_damping=0.0, iterations=5)
    s0, s1 = model.state(), model.state()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s0)
    control = model.control()
    control.joint_target_q.copy_(torch.tensor([pos_target], dtype=torch.float32))
    control.joint_target_qd.copy_(torch.tensor([vel_target], dtype=torch.float32))
    for _ in range(100):
        s0.clear_forces()
        solver.step(s0, s1, control, None, 1.0 / 60.0)
        s0, s1 = s1, s0
    if solver_name == "xpbd":
        oracle_lib.eval_ik(model, s0, s0.joint_q, s0.joint_qd)
    if expected_pos is not None:
        assert float(s0.joint_q[0]) == pytest.approx(expected_pos, abs=1e-2)
    if expected_vel is not None:
        assert float(s0.joint_qd[0]) == pytest.approx(expected_vel, abs=1e-2)


# ---- test_joint_damping.py:17-146, 190-212 (Featherstone rows) ---------------------------------------------------------------------
def _damped_spin(oracle_lib, kind, damping):
    builder = ModelBuilder(gravity=(0.0, 0.0, 0.0), up_axis="y")
    body = builder.add_link(mass=1.0, inertia=np.eye(3), lock_inertia=True)
    if kind == "revolute":
        joint = builder.add_joint_revolute(-1, body, axis=(0.0, 0.0, 1.0), target_ke=0.0, target_kd=0.0, damping=damping, limit_lower=-1.0e6,
                                           limit_upper=1.0e6, limit_ke=0.0, limit_kd=0.0, armature=0.0, friction=0.0)
    else:
        joint = builder.add_joint_ball(-1, body, damping=damping, armature=0.0, friction=0.0)
    builder.add_articulation([joint])
    builder.joint_qd[0] = 1.0
    model = builder.finalize()
    solver = oracle_lib.SolverFeatherstone(model, angular_damping=0.0)
    s0, s1 = model.state(), model.state()
    control = model.control()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s0)
    n = 1 if kind == "revolute" else 3
    initial = float(np.linalg.norm(s0.joint_qd.numpy()[:n]))
    for _ in range(8):
        solver.step(s0, s1, control, None, 0.01)
        s0, s1 = s1, s0
    return model, initial, float(np.linalg.norm(s0.joint_qd.numpy()[:n]))


@pytest.mark.parametrize("kind", ["revolute", "ball"])
def test_featherstone_joint_damping_decays_velocity(oracle_lib, kind):
    _, undamped_initial, undamped_final = _damped_spin(oracle_lib, kind, 0.0)
    model, damped_initial, damped_final = _damped_spin(oracle_lib, kind, 3.0)
    assert undamped_final == pytest.approx(undamped_initial, abs=1e-5, rel=1e-5)
    assert damped_final < damped_initial * 0.85
    if kind == "ball":
        np.testing.assert_allclose(model.numpy("joint_damping")[:3], [3.0, 3.0, 3.0])  # add_joint_ball(damping=...) reaches every axis


# ---- test_body_velocity.py:124-371, 848-951 (free body with an offset centre of mass; XPBD and Featherstone rows) -----------------------
COM_OFFSETS = [(0.5, 0.0, 0.0), (0.0, 0.3, 0.0), (0.0, 0.0, 0.4), (0.2, 0.3, 0.1)]
AXES = [(0.0, 0.0, 1.0), (0.0, 1.0, 0.0), (1.0, 0.0, 0.0)]
BODY_VELOCITY_SOLVERS = [("featherstone", True, 1e-3), ("xpbd", False, 1e-4)]  # (name, state set through joint_qd, CoM tolerance [m])


def _com_world(state, model, body=0):
    q = state.body_q.numpy()[body].astype(np.float64)
    return q[:3] + X.quat_rotate(q[3:], model.numpy("body_com")[body].astype(np.float64))


def _free_body_run(oracle_lib, solver_name, generalized, com, velocity, initial_pos):
    builder = ModelBuilder(gravity=(0.0, 0.0, 0.0))
    b = builder.add_body(xform=X.transform(initial_pos))
    builder.add_shape_box(b, hx=0.1, hy=0.1, hz=0.1)
    builder.body_com[b] = np.asarray(com, dtype=np.float64)
    model = builder.finalize()
    solver = _solver(oracle_lib, solver_name, model)
    s0, s1 = model.state(), model.state()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s0)
    v = torch.tensor(velocity, dtype=torch.float32)
    if generalized:  # Featherstone integrates joint_qd; body_qd follows through FK
        s0.joint_qd.copy_(v)
        oracle_lib.eval_fk(model, s0.joint_q, s0.joint_qd, s0)
    else:  # XPBD integrates body_qd
        s0.body_qd.copy_(v.reshape(1, 6))
    q0, com0 = s0.body_q.numpy()[0].copy(), _com_world(s0, model)
    for _ in range(10):
        solver.step(s0, s1, None, None, 0.01)
        s0, s1 = s1, s0
    return q0, com0, s0.body_q.numpy()[0].copy(), _com_world(s0, model)


@pytest.mark.parametrize("solver_name,generalized,tol", BODY_VELOCITY_SOLVERS)
@pytest.mark.parametrize("com", COM_OFFSETS)
@pytest.mark.parametrize("omega", AXES)
def test_angular_velocity_keeps_com_stationary(oracle_lib, solver_name, generalized, tol, com, omega):
    q0, com0, q1, com1 = _free_body_run(oracle_lib, solver_name, generalized, com, (0.0, 0.0, 0.0, *omega), (1.0, 2.0, 3.0))
    assert np.linalg.norm(com1 - com0) < tol
    assert abs(np.dot(q0[3:], q1[3:])) < 0.9999  # it did rotate


@pytest.mark.parametrize("solver_name,generalized,tol", BODY_VELOCITY_SOLVERS)
@pytest.mark.parametrize("com", COM_OFFSETS)
@pytest.mark.parametrize("direction", AXES)
def test_linear_velocity_moves_com(oracle_lib, solver_name, generalized, tol, com, direction):
    v = tuple(0.7 * d for d in direction[::-1])  # (0.7, 0, 0), (0, 0.7, 0), (0, 0, 0.7)
    _, com0, _, com1 = _free_body_run(oracle_lib, solver_name, generalized, com, (*v, 0.0, 0.0, 0.0), (0.0, 0.0, 1.0))
    assert np.linalg.norm((com1 - com0) - np.array(v) * 0.1) < tol


@pytest.mark.parametrize("solver_name,generalized,tol", BODY_VELOCITY_SOLVERS)
@pytest.mark.parametrize("com", COM_OFFSETS)
def test_combined_velocity_com_follows_linear_part(oracle_lib, solver_name, generalized, tol, com):
    q0, com0, q1, com1 = _free_body_run(oracle_lib, solver_name, generalized, com, (0.1, 0.0, 0.0, 0.0, 0.0, 1.0), (0.0, 0.0, 1.0))
    assert np.linalg.norm((com1 - com0) - np.array([0.1, 0.0, 0.0]) * 0.1) < tol
    assert abs(np.dot(q0[3:], q1[3:])) < 0.9999


def test_featherstone_root_free_joint_under_rotated_parent_xform_reports_parent_frame_qd(oracle_lib):
    """:374-405 - joint_qd of a root FREE joint lives in the (rotated) parent joint frame, body_qd in the world frame at the COM."""
    builder = ModelBuilder(gravity=(0.0, 0.0, -10.0), up_axis="z")
    parent_xform = X.transform((0.5, 0.6, 0.7), X.quat_from_axis_angle(np.array([1.0, 0.0, 0.0]), np.pi / 2.0))
    body = builder.add_link(mass=1.0, inertia=np.eye(3))
    builder.add_articulation([builder.add_joint_free(parent=-1, child=body, parent_xform=parent_xform)])
    model = builder.finalize()
    solver = oracle_lib.SolverFeatherstone(model, angular_damping=0.0)
    s0, s1 = model.state(), model.state()
    oracle_lib.eval_fk(model, s0.joint_q, s0.joint_qd, s0)
    dt = 1e-2
    pipe = oracle_lib.CollisionPipeline(model)
    solver.step(s0, s1, model.control(), pipe.contacts(), dt)
    np.testing.assert_allclose(s1.joint_qd.numpy()[0:3], (0.0, -10.0 * dt, 0.0), atol=1e-5)  # world -Z is the parent frame's -Y
    np.testing.assert_allclose(s1.joint_qd.numpy()[3:6], (0.0, 0.0, 0.0), atol=1e-5)
    np.testing.assert_allclose(s1.body_qd.numpy()[body, 0:3], (0.0, 0.0, -10.0 * dt), atol=1e-5)
    np.testing.assert_allclose(s1.body_qd.numpy()[body, 3:6], (0.0, 0.0, 0.0), atol=1e-5)


# ---- test_collision_plane_halfspace_aabb.py (half-space AABB of infinite planes, sim/collide.py:352-377) ---------------------------
TILT = math.radians(0.05)
FAR_X = -600.0


def _surface_z(normal, plane_d, x):
    return (-plane_d - normal[0] * x) / normal[2]


def _collide_once(oracle_lib, builder):
    model = builder.finalize()
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    state = model.state()
    pipe.collide(state, contacts)
    lo, hi = oracle_lib.shape_aabbs(model, state.body_q)
    return int(contacts.rigid_contact_count.item()), lo, hi


def _sphere_on_plane(oracle_lib, tilted, z_offset=0.0):
    builder = ModelBuilder()
    normal = (math.sin(TILT), 0.0, math.cos(TILT)) if tilted else (0.0, 0.0, 1.0)
    builder.add_shape_plane(plane=(*normal, 0.0), width=0.0, length=0.0)
    center_z = _surface_z(normal, 0.001, FAR_X) + 0.5 / normal[2]  # resting, 1 mm into the surface
    body = builder.add_body(xform=X.transform((FAR_X, 0.0, center_z + z_offset)))
    builder.add_shape_sphere(body, radius=0.5)
    return _collide_once(oracle_lib, builder)


def test_plane_halfspace_aabb_sphere_cases(oracle_lib):
    """:106-147"""
    normal = (math.sin(TILT), 0.0, math.cos(TILT))
    count, _, hi = _sphere_on_plane(oracle_lib, tilted=True)
    assert count > 0  # resting contact 600 m from the anchor of a 0.05-degree floor survives the broad phase
    assert hi[0][2] > _surface_z(normal, 0.0, FAR_X)  # the clamped bound clears the surface it can reach
    count, _, hi = _sphere_on_plane(oracle_lib, tilted=False)
    assert count > 0 and hi[0][2] < 1.0  # an exactly aligned ground is clamped at its surface
    count, _, hi = _sphere_on_plane(oracle_lib, tilted=False, z_offset=50.0)
    assert count == 0 and hi[0][2] < 1.0  # ... and prunes a shape hovering 50 m above it


@pytest.mark.parametrize("plane_d", [0.0, -25.0])
@pytest.mark.parametrize("tilt_deg", [0.0, 0.05, 0.081, 0.1, 1.0])
@pytest.mark.parametrize("offset", [-100.0, -600.0])
def test_plane_halfspace_aabb_resting_box_tilt_table(oracle_lib, plane_d, tilt_deg, offset):
    """:150-196"""
    builder = ModelBuilder()
    tilt = math.radians(tilt_deg)
    normal = (math.sin(tilt), 0.0, math.cos(tilt))
    builder.add_shape_plane(plane=(*normal, plane_d), width=0.0, length=0.0)
    half = 0.5
    support = half * (abs(normal[0]) + abs(normal[2]))
    center_z = (-plane_d + support - 0.001 - normal[0] * offset) / normal[2]
    body = builder.add_body(xform=X.transform((offset, 0.0, center_z)))
    builder.add_shape_box(body, hx=half, hy=half, hz=half)
    count, _, hi = _collide_once(oracle_lib, builder)
    assert count > 0, "resting box lost its ground contact (fall-through)"
    if tilt_deg == 0.0:
        assert hi[0][2] < _surface_z(normal, plane_d, 0.0) + 1.0
    else:
        assert hi[0][2] > _surface_z(normal, plane_d, offset)


def test_plane_halfspace_aabb_non_z_normal_still_clamps(oracle_lib):
    """:199-217 - a +X plane built through quat_between_vectors carries ~1 ulp of lateral residue; the clamp must still engage."""
    builder = ModelBuilder()
    builder.add_shape_plane(plane=(1.0, 0.0, 0.0, 0.0), width=0.0, length=0.0)
    builder.add_body(mass=1.0, inertia=np.eye(3))
    _, _, hi = _collide_once(oracle_lib, builder)
    assert hi[0][0] < 1.0


# ---- test_cone_orientation.py (cone mass properties as the builder computes them; host-side input construction) -------------------
def _cone_model(axis="z"):
    builder = ModelBuilder()
    body = builder.add_body()
    cfg = newton_b200.ShapeConfig()
    cfg.density = 1000.0
    rot = {"x": X.quat_from_axis_angle(np.array([0.0, 1.0, 0.0]), np.pi / 2), "y": X.quat_from_axis_angle(np.array([1.0, 0.0, 0.0]), -np.pi / 2),
           "z": np.array([0.0, 0.0, 0.0, 1.0])}[axis]
    builder.add_shape_cone(body, xform=X.transform((0.0, 0.0, 0.0), rot), radius=1.0, half_height=2.0, cfg=cfg)
    return builder.finalize()


def test_cone_mass_properties():
    """:22-43, :81-164 - COM a quarter of the height above the base (apex along +axis), m = rho pi r^2 h / 3, I_xx = I_yy != I_zz."""
    model = _cone_model()
    np.testing.assert_allclose(model.numpy("body_com")[0], (0.0, 0.0, -1.0), atol=1e-6)
    assert float(model.body_mass[0]) == pytest.approx(1000.0 * np.pi * 1.0 * 4.0 / 3.0, abs=1e-3 * 4189.0 * 1e-3 + 1e-2)
    inertia = model.numpy("body_inertia")[0]
    assert inertia[0, 0] == pytest.approx(inertia[1, 1], rel=1e-6) and abs(inertia[0, 0] - inertia[2, 2]) > 1e-2
    assert np.abs(inertia - np.diag(np.diag(inertia))).max() < 1e-3
    for axis, expected in (("x", (-1.0, 0.0, 0.0)), ("y", (0.0, -1.0, 0.0)), ("z", (0.0, 0.0, -1.0))):
        np.testing.assert_allclose(_cone_model(axis).numpy("body_com")[0], expected, atol=1e-5)


# ---- test_collision_pipeline.py:85-262, 331-386 (head-on collision of two free bodies, XPBD defaults, primitive pairs) ----------------
from newton_b200 import GeoType  # noqa: E402

VX, VYZ, ANG = 1, 2, 4
HEAD_ON = [  # (shape a, shape b, checks on a, checks on b) - the rows of collision_pipeline_contact_tests without meshes
    (GeoType.SPHERE, GeoType.SPHERE, VYZ, VX | VYZ | ANG),
    (GeoType.SPHERE, GeoType.BOX, VYZ, VX | VYZ | ANG),
    (GeoType.SPHERE, GeoType.CAPSULE, VYZ, VX | VYZ | ANG),
    (GeoType.SPHERE, GeoType.CYLINDER, VYZ, VX | VYZ | ANG),
    (GeoType.SPHERE, GeoType.CONE, VYZ, VYZ),
    (GeoType.BOX, GeoType.BOX, VYZ, VX | VYZ),
    (GeoType.CAPSULE, GeoType.CAPSULE, VYZ, VX | VYZ),
]


def _add_head_on_shape(builder, shape_type, body):
    if shape_type == GeoType.BOX:
        builder.add_shape_box(body)
    elif shape_type == GeoType.SPHERE:
        builder.add_shape_sphere(body, radius=0.5)
    elif shape_type == GeoType.CAPSULE:
        builder.add_shape_capsule(body, radius=0.25, half_height=0.3)
    elif shape_type == GeoType.CYLINDER:
        builder.add_shape_cylinder(body, radius=0.25, half_height=0.4)
    elif shape_type == GeoType.CONE:  # flat base towards the incoming body
        builder.add_shape_cone(body, xform=X.transform((0.0, 0.0, 0.0), X.quat_from_axis_angle(np.array([0.0, 1.0, 0.0]), -np.pi / 2.0)),
                               radius=0.25, half_height=0.4)


@pytest.mark.parametrize("broad_phase", ["explicit", "nxn", "sap"])
@pytest.mark.parametrize("type_a,type_b,level_a,level_b", HEAD_ON)
def test_head_on_collision_transfers_momentum_along_x(oracle_lib, type_a, type_b, level_a, level_b, broad_phase):
    builder = ModelBuilder(gravity=(0.0, 0.0, 0.0))
    builder.rigid_gap = 0.005
    body_a = builder.add_body(xform=X.transform((-1.0, 0.0, 0.0)))
    _add_head_on_shape(builder, type_a, body_a)
    builder.joint_qd[0] = 5.0
    builder.body_qd[-1] = np.array([5.0, 0.0, 0.0, 0.0, 0.0, 0.0])
    body_b = builder.add_body(xform=X.transform((1.0, 0.0, 0.0)))
    _add_head_on_shape(builder, type_b, body_b)
    model = builder.finalize()
    pipe = oracle_lib.CollisionPipeline(model, broad_phase=broad_phase)
    contacts = pipe.contacts()
    solver = oracle_lib.SolverXPBD(model)
    s0, s1, control = model.state(), model.state(), model.control()
    dt = 1.0 / 60.0 / 10
    for _ in range(100):
        pipe.collide(s0, contacts)  # once per frame, like the reference test
        for _ in range(10):
            s0.clear_forces()
            solver.step(s0, s1, control, contacts, dt)
            s0, s1 = s1, s0
    qd = s0.body_qd.numpy()
    for body, level in ((body_a, level_a), (body_b, level_b)):
        if level & VX:
            assert 0.03 < qd[body, 0] <= 5.0, (body, qd[body])
        if level & VYZ:
            assert abs(qd[body, 1]) < 3e-3 and abs(qd[body, 2]) < 3e-3, (body, qd[body])
        if level & ANG:
            assert np.abs(qd[body, 3:]).max() < 3e-3, (body, qd[body])


# ---- explicit pair list == what the NxN / SAP broad phases may emit (sim/builder.py:12796-12797 vs broad_phase_nxn.py:124-216) -------
def _friction_grid_model():
    builder = ModelBuilder(up_axis="z", gravity=(0.0, 0.0, -9.81))
    for cell in range(6):
        cfg = newton_b200.ShapeConfig()
        cfg.collision_group = cell + 1
        builder.add_shape_box(-1, xform=X.transform((cell * 2.5, 0.0, 2.0)), hx=0.5, hy=2.5, hz=0.05, cfg=cfg)
        body = builder.add_body(xform=X.transform((cell * 2.5, 0.0, 2.2)))
        builder.add_shape_box(body, hx=0.2, hy=0.2, hz=0.05, cfg=cfg)
    return builder.finalize()


def _group_zoo_model():
    """Positive / negative / zero collision groups, a visual-only shape, an explicit filter pair, global + per-world shapes."""
    builder = ModelBuilder()
    builder.add_ground_plane()
    world = ModelBuilder()
    for k, group in enumerate((1, 1, 2, -1, -2, 0, -1)):
        cfg = newton_b200.ShapeConfig()
        cfg.collision_group = group
        if k == 1:
            cfg.has_shape_collision = False  # visual only: no COLLIDE_SHAPES flag
        body = world.add_body(xform=X.transform((0.3 * k, 0.0, 0.5)))
        world.add_shape_sphere(body, radius=0.2, cfg=cfg)
    world.add_shape_collision_filter_pair(2, 3)
    builder.replicate(world, 3)
    cfg = newton_b200.ShapeConfig()
    cfg.collision_group = -3
    builder.add_shape_box(-1, xform=X.transform((0.0, 0.0, -1.0)), cfg=cfg)  # a second global shape, after the worlds
    return builder.finalize()


SCENES = {
    "quadrupeds": lambda: scenes_mod().quadruped_model(3), "box_stacks": lambda: scenes_mod().box_stack_model(2),
    "convex_pile": lambda: scenes_mod().convex_pile_model(2), "mixed_worlds": lambda: scenes_mod().mixed_worlds_model(2),
    "platform": lambda: scenes_mod().platform_model(2), "ants": lambda: scenes_mod().ants_model(2, 2),
    "shapes_on_plane": lambda: scenes_mod().shapes_on_plane_model(2), "pendulum": lambda: scenes_mod().pendulum_model(),
    "friction_grid": _friction_grid_model, "group_zoo": _group_zoo_model,
}


def scenes_mod():
    from newton_b200 import scenes

    return scenes


@pytest.mark.parametrize("name", list(SCENES))
def test_explicit_pair_list_equals_nxn_filter(oracle_lib, name):
    from oracle import broad_phase as bp

    model = SCENES[name]()
    explicit = [tuple(p) for p in model.numpy("shape_contact_pairs").tolist()]
    assert all(a < b for a, b in explicit) and len(set(explicit)) == len(explicit)
    assert set(explicit) == bp.model_nxn_pairs(model, model.shape_collision_filter_pairs)
    if name == "group_zoo":
        assert len(explicit) > 0 and model.shape_collision_filter_pairs
        # ... and with the immovable filter on (include_static_kinematic_pairs=False) only static-static pairs disappear
        pruned = bp.model_nxn_pairs(model, model.shape_collision_filter_pairs, include_static_kinematic_pairs=False)
        body = model.numpy("shape_body")
        assert set(explicit) - pruned == {p for p in explicit if body[p[0]] < 0 and body[p[1]] < 0}


def test_world_map_layout(oracle_lib):
    """precompute_world_map docstring example (broad_phase_common.py:271-306): regular worlds each followed by the shared shapes,
    then the dedicated shared-only segment; visual-only shapes dropped."""
    from oracle import broad_phase as bp

    index_map, slice_ends = bp.precompute_world_map([-1, 0, 0, 1, -1, 1], [2, 2, 0, 2, 2, 2])
    assert index_map.tolist() == [1, 0, 4, 3, 5, 0, 4, 0, 4] and slice_ends.tolist() == [3, 7, 9]


def test_nxn_filter_equals_reference_brute_force(oracle_lib):
    """The segment-wise enumeration of the NxN kernel (world map + dedicated shared segment) admits exactly the pairs of the
    reference's own host check ``find_overlapping_pairs_np`` (newton/tests/test_broad_phase.py:92-149) when every AABB overlaps:
    i < j, both with COLLIDE_SHAPES, test_world_and_group_pair - each pair once, shared-vs-shared pairs included."""
    from oracle import broad_phase as bp

    rng = np.random.default_rng(4)
    for trial in range(30):
        n = int(rng.integers(1, 40))
        worlds = rng.integers(-1, 4, size=n)
        if trial % 3 == 0:
            worlds = np.sort(worlds)
        groups = rng.integers(-3, 4, size=n)
        flags = np.where(rng.random(n) < 0.8, 3, 1)  # bit 1 = COLLIDE_SHAPES
        expected = set()
        for i in range(n):
            for j in range(i + 1, n):
                if (flags[i] & 2) == 0 or (flags[j] & 2) == 0:
                    continue
                wi, wj, gi, gj = int(worlds[i]), int(worlds[j]), int(groups[i]), int(groups[j])
                if wi != -1 and wj != -1 and wi != wj:
                    continue
                if gi == 0 or gj == 0:
                    continue
                if not ((gi == gj or gj < 0) if gi > 0 else gi != gj):
                    continue
                expected.add((i, j))
        assert bp.nxn_filter_pairs(worlds, flags, groups) == expected, trial


# ---- test_environment_group_collision.py (world / collision-group filtering in the explicit pair list) ---------------------------------
def _cfg(group, collide=True):
    cfg = newton_b200.ShapeConfig()
    cfg.collision_group = group
    cfg.has_shape_collision = collide
    return cfg


def test_shapes_of_different_worlds_are_not_paired():
    """:20-75"""
    builder = ModelBuilder()
    bodies = [builder.add_body() for _ in range(3)]
    for k, x in enumerate((0.0, 0.8)):
        builder.begin_world()
        builder.add_shape_box(bodies[k], xform=X.transform((x, 0.0, 0.0)), hx=0.5, hy=0.5, hz=0.5, cfg=_cfg(1))
        builder.end_world()
    builder.add_shape_box(bodies[2], xform=X.transform((0.4, 1.0, 0.0)), hx=0.5, hy=0.5, hz=0.5, cfg=_cfg(-1))  # global, collides with all
    model = builder.finalize()
    pairs = {tuple(sorted(p)) for p in model.numpy("shape_contact_pairs").tolist()}
    assert model.shape_contact_pair_count == 2 and pairs == {(0, 2), (1, 2)}


def test_add_world_assigns_world_indices_and_keeps_groups():
    """:116-169"""
    robot = ModelBuilder()
    robot.add_body(label="base")
    robot.add_shape_box(0, hx=0.5, hy=0.5, hz=0.5, cfg=_cfg(1))
    robot.add_body(label="link1")
    robot.add_shape_capsule(1, radius=0.1, half_height=0.5, cfg=_cfg(2))
    main = ModelBuilder()
    main.add_shape_box(-1, xform=X.transform((0.0, -1.0, 0.0)), hx=10, hy=0.1, hz=10, cfg=_cfg(-1))
    main.add_world(robot)
    main.add_world(robot)
    model = main.finalize()
    assert model.numpy("shape_world").tolist() == [-1, 0, 0, 1, 1]
    assert model.numpy("body_world").tolist() == [0, 0, 1, 1]
    assert model.numpy("shape_collision_group").tolist() == [-1, 1, 2, 1, 2]


def test_mixed_collision_and_world_groups_known_pairs(oracle_lib):
    """:171-259"""
    from oracle import broad_phase as bp

    builder = ModelBuilder()
    bodies = [builder.add_body() for _ in range(7)]
    builder.begin_world()
    builder.add_shape_sphere(bodies[0], xform=X.transform((-1.0, 0.0, 0.0)), radius=0.5, cfg=_cfg(1))
    builder.add_shape_sphere(bodies[1], xform=X.transform((0.0, 0.0, 0.0)), radius=0.5, cfg=_cfg(2))
    builder.add_shape_sphere(bodies[2], xform=X.transform((1.0, 0.0, 0.0)), radius=0.5, cfg=_cfg(-1))
    builder.end_world()
    builder.begin_world()
    builder.add_shape_sphere(bodies[3], xform=X.transform((-1.0, 2.0, 0.0)), radius=0.5, cfg=_cfg(1))
    builder.add_shape_sphere(bodies[4], xform=X.transform((0.0, 2.0, 0.0)), radius=0.5, cfg=_cfg(2))
    builder.end_world()
    builder.add_shape_sphere(bodies[5], xform=X.transform((0.0, 2.0, 0.0)), radius=0.5, cfg=_cfg(2, collide=False))  # global, visual only
    builder.add_shape_sphere(bodies[6], xform=X.transform((0.0, 4.0, 0.0)), radius=0.5, cfg=_cfg(1))  # global
    model = builder.finalize()
    pairs = {tuple(sorted(p)) for p in model.numpy("shape_contact_pairs").tolist()}
    assert pairs == {(0, 2), (1, 2), (0, 6), (2, 6), (3, 6)}
    assert pairs == bp.model_nxn_pairs(model, model.shape_collision_filter_pairs)


def test_collision_filter_pairs_are_canonical_after_merging_builders():
    """:261-316 - child shapes created before the parent's: the parent/child filter pairs must still be (low, high)."""
    builder = ModelBuilder()
    child = builder.add_link()
    builder.add_shape_box(child, hx=0.5, hy=0.5, hz=0.5)
    builder.add_shape_box(child, hx=0.5, hy=0.5, hz=0.5)
    parent = builder.add_link(xform=X.transform((2.0, 0.0, 0.0)))
    builder.add_shape_box(parent, hx=0.5, hy=0.5, hz=0.5)
    builder.add_shape_box(parent, hx=0.5, hy=0.5, hz=0.5)
    joint = builder.add_joint_revolute(parent, child, axis=(0.0, 0.0, 1.0), collision_filter_parent=True)
    builder.add_articulation([joint])
    sub = ModelBuilder()
    sub.add_shape_box(sub.add_body(), hx=0.5, hy=0.5, hz=0.5)
    builder.add_shape_box(child, hx=0.5, hy=0.5, hz=0.5)  # index 4
    builder.add_builder(sub)
    model = builder.finalize()
    pairs = {tuple(sorted(p)) for p in model.numpy("shape_contact_pairs").tolist()}
    for parent_shape in (2, 3):
        for child_shape in (0, 1, 4):
            assert (min(parent_shape, child_shape), max(parent_shape, child_shape)) not in pairs
    assert all(a < b for a, b in model.shape_collision_filter_pairs)


@pytest.mark.parametrize("world_a,world_b,col_a,col_b,expected", [
    (0, 0, 1, 1, True), (1, 1, -1, 2, True), (2, 2, 0, 1, False), (0, 1, 1, 1, False), (2, 3, -1, -1, False),
    (-1, 0, 1, 1, True), (1, -1, 2, 2, True), (-1, -1, 1, 2, False), (-1, -1, -1, 1, True)])
def test_world_and_group_pair_table(oracle_lib, world_a, world_b, col_a, col_b, expected):
    """:321-356"""
    from oracle import broad_phase as bp

    assert bp.test_world_and_group_pair(world_a, world_b, col_a, col_b) == expected
    assert ModelBuilder._test_group_pair(col_a, col_b) == bp.test_group_pair(col_a, col_b)


# ---- test_solver_xpbd.py: joint projection under large errors, contact-force reporting edge cases, articulation drift -----------------
def _xf_point(q7, p):
    return q7[:3].astype(np.float64) + X.quat_rotate(q7[3:].astype(np.float64), np.asarray(p, dtype=np.float64))


def test_xpbd_distance_joint_limits(oracle_lib):
    """:151-181 - one step projects the anchor distance into [min, max] from separated and from coincident anchors."""
    def solve(initial_distance, min_distance, max_distance):
        builder = ModelBuilder(gravity=(0.0, 0.0, 0.0))
        body = builder.add_link(xform=X.transform((initial_distance, 0.0, 0.0)))
        builder.add_shape_sphere(body, radius=0.1)
        joint = builder.add_joint_distance(-1, body, parent_xform=X.transform((0.0, 0.0, 0.0), X.quat_from_axis_angle(np.array([0.0, 0.0, 1.0]), np.pi / 2)),
                                           min_distance=min_distance, max_distance=max_distance)
        builder.add_articulation([joint])
        model = builder.finalize()
        s_in, s_out = model.state(), model.state()
        oracle_lib.SolverXPBD(model, iterations=10).step(s_in, s_out, None, None, 1.0 / 60.0)
        return s_out.body_q.numpy()[body, :3]

    assert np.linalg.norm(solve(0.25, 1.0, -1.0)) >= 0.99
    np.testing.assert_allclose(solve(0.0, 1.0, -1.0), (0.0, 1.0, 0.0), atol=0.01)  # coincident anchors: pushed along the joint's x axis
    assert np.linalg.norm(solve(2.0, -1.0, 1.0)) <= 1.01


def _two_capsules(builder):
    shape_xform = X.transform((0.0, 0.0, 0.0), X.quat_from_axis_angle(np.array([0.0, 1.0, 0.0]), 0.5 * np.pi))
    parent = builder.add_link()
    builder.add_shape_capsule(parent, xform=shape_xform, radius=0.0625, half_height=0.25)
    return parent, shape_xform


def test_xpbd_ball_joint_recovers_from_large_anchor_separation(oracle_lib):
    """:184-248"""
    half_extent = 0.0625 + 0.25
    builder = ModelBuilder(gravity=(0.0, 0.0, 0.0))
    parent, shape_xform = _two_capsules(builder)
    child = builder.add_link(xform=X.transform((2.0 * half_extent, 0.0, 0.0)))
    builder.add_shape_capsule(child, xform=shape_xform, radius=0.0625, half_height=0.25)
    root = builder.add_joint_free(child=parent)
    ball = builder.add_joint_ball(parent, child, parent_xform=X.transform((half_extent, 0.0, 0.0)), child_xform=X.transform((-half_extent, 0.0, 0.0)))
    builder.add_articulation([root, ball])
    model = builder.finalize()
    s0, s1 = model.state(), model.state()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s0)
    s0.body_q[child, :3] += torch.tensor([1.0, 1.0, 0.0])
    oracle_lib.SolverXPBD(model, iterations=2).step(s0, s1, None, None, 1.0 / 240.0)
    q = s1.body_q.numpy()
    gap = np.linalg.norm(_xf_point(q[child], (-half_extent, 0.0, 0.0)) - _xf_point(q[parent], (half_extent, 0.0, 0.0)))
    assert gap < 0.5


def test_xpbd_prismatic_joint_recovers_from_large_transverse_separation(oracle_lib):
    """:251-321"""
    builder = ModelBuilder(gravity=(0.0, 0.0, 0.0))
    parent, shape_xform = _two_capsules(builder)
    child = builder.add_link()
    builder.add_shape_capsule(child, xform=shape_xform, radius=0.0625, half_height=0.25)
    root = builder.add_joint_free(child=parent)
    slider = builder.add_joint_prismatic(parent, child, axis=(1.0, 0.0, 0.0), limit_lower=-2.0, limit_upper=2.0)
    builder.add_articulation([root, slider])
    model = builder.finalize()
    s0, s1 = model.state(), model.state()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s0)
    s0.body_q[child, :3] += torch.tensor([0.5, 1.0, 1.0])
    oracle_lib.SolverXPBD(model, iterations=2).step(s0, s1, None, None, 1.0 / 240.0)
    q = s1.body_q.numpy().astype(np.float64)
    rel = X.quat_rotate(X.quat_inverse(q[parent, 3:]), q[child, :3] - q[parent, :3])
    assert np.hypot(rel[1], rel[2]) < 0.5 and -2.0 <= rel[0] <= 2.0


def test_xpbd_prismatic_joint_retains_extension_in_parent_moment_arm(oracle_lib):
    """:324-363 - the correction's torque on the parent must use the moment arm out to the (valid) joint extension."""
    builder = ModelBuilder(gravity=(0.0, 0.0, 0.0))
    parent = builder.add_link()
    builder.add_shape_sphere(parent, radius=0.25)
    child = builder.add_link()
    builder.add_shape_sphere(child, radius=0.25)
    root = builder.add_joint_free(child=parent)
    slider = builder.add_joint_prismatic(parent, child, axis=(1.0, 0.0, 0.0), limit_lower=-2.0, limit_upper=2.0, damping=0.0)
    builder.add_articulation([root, slider])
    model = builder.finalize()
    s0, s1 = model.state(), model.state()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s0)
    s0.body_q[child, :2] += torch.tensor([0.5, 0.1])
    oracle_lib.SolverXPBD(model, iterations=2, angular_damping=0.0).step(s0, s1, None, None, 1.0 / 240.0)
    assert abs(float(s1.body_q[parent, 5])) > 0.01


def test_xpbd_contact_force_is_zero_without_contact_or_touch(oracle_lib):
    """:1038-1108"""
    builder = ModelBuilder()
    body = builder.add_body(xform=X.transform((0.0, 0.0, 5.0)))
    builder.add_shape_sphere(body, radius=0.25)
    model = builder.finalize()
    model.request_contact_attributes("force")
    solver = oracle_lib.SolverXPBD(model, iterations=2)
    s_in, s_out = model.state(), model.state()
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    pipe.collide(s_in, contacts)
    solver.step(s_in, s_out, model.control(), contacts, 1.0 / 60.0)
    solver.update_contacts(contacts, s_out)
    n = int(contacts.rigid_contact_count.item())
    assert n == 0 or np.abs(contacts.force.numpy()[:n]).max() <= 1e-6

    builder = ModelBuilder()
    builder.default_shape_cfg.gap = 1.0  # a contact pair is generated 0.5 m above the surface ...
    builder.add_ground_plane()
    body = builder.add_body(xform=X.transform((0.0, 0.0, 0.25 + 0.5)))
    builder.add_shape_sphere(body, radius=0.25)
    model = builder.finalize()
    model.set_gravity((0.0, 0.0, 0.0))
    model.request_contact_attributes("force")
    solver = oracle_lib.SolverXPBD(model, iterations=2)
    s_in, s_out = model.state(), model.state()
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    pipe.collide(s_in, contacts)
    n = int(contacts.rigid_contact_count.item())
    assert n > 0
    solver.step(s_in, s_out, model.control(), contacts, 1.0 / 60.0)
    solver.update_contacts(contacts, s_out)
    np.testing.assert_allclose(contacts.force.numpy()[:n, :3], 0.0, atol=1e-6)  # ... but reports no force


def test_xpbd_update_contacts_requires_force_attribute(oracle_lib):
    """:1111-1132"""
    builder = ModelBuilder()
    builder.add_ground_plane()
    body = builder.add_body(xform=X.transform((0.0, 0.0, 0.25)))
    builder.add_shape_sphere(body, radius=0.25)
    model = builder.finalize()
    solver = oracle_lib.SolverXPBD(model, iterations=2)
    s_in, s_out = model.state(), model.state()
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    pipe.collide(s_in, contacts)
    solver.step(s_in, s_out, model.control(), contacts, 1.0 / 60.0)
    assert contacts.force is None
    with pytest.raises(ValueError):
        solver.update_contacts(contacts)


def test_xpbd_articulation_contact_drift(oracle_lib):
    """:750-845 (the reference's issue #2030) - a quadruped lying on its side must not creep: contacts are solved before joints in
    every iteration; solving joints first leaves ~6 mm/s of lateral drift.  < 1 cm over 3 s after 2 s of settling."""
    from newton_b200 import scenes

    builder = ModelBuilder()
    builder.default_joint_cfg.armature = 0.01
    builder.default_joint_cfg.target_ke = 2000.0
    builder.default_joint_cfg.target_kd = 1.0
    builder.default_shape_cfg.ke, builder.default_shape_cfg.kd, builder.default_shape_cfg.kf, builder.default_shape_cfg.mu = 1.0e4, 1.0e2, 1.0e2, 1.0
    rot = X.quat_from_axis_angle(np.array([1.0, 0.0, 0.0]), np.pi * 0.5)
    builder.add_urdf(scenes.quadruped_urdf(), xform=X.transform((0.0, 0.0, 0.3), rot), floating=True, enable_self_collisions=False,
                     ignore_inertial_definitions=True)
    for i in range(builder.body_count):
        builder.body_inertia[i] = builder.body_inertia[i] + np.eye(3) * 0.01
        builder.body_inv_inertia[i] = np.linalg.inv(builder.body_inertia[i])
    builder.joint_q[-12:] = [0.2, 0.4, -0.6, -0.2, -0.4, 0.6, -0.2, 0.4, -0.6, 0.2, -0.4, 0.6]
    builder.joint_target_q[-12:] = builder.joint_q[-12:]
    builder.add_ground_plane()
    model = builder.finalize()
    solver = oracle_lib.SolverXPBD(model)
    s0, s1, control = model.state(), model.state(), model.control()
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s0)

    def run(frames, s0, s1):
        for _ in range(frames * 10):
            s0.clear_forces()
            pipe.collide(s0, contacts)
            solver.step(s0, s1, control, contacts, 1.0 / 1000.0)
            s0, s1 = s1, s0
        return s0, s1

    s0, s1 = run(200, s0, s1)
    x0, y0 = float(s0.body_q[0, 0]), float(s0.body_q[0, 1])
    s0, s1 = run(300, s0, s1)
    drift = float(np.hypot(float(s0.body_q[0, 0]) - x0, float(s0.body_q[0, 1]) - y0))
    assert np.isfinite(s0.body_q.numpy()).all() and drift < 0.01, drift


# ---- test_solver_xpbd.py:1283-1526 (State.body_parent_f beyond the single-body cases of test_oracle_known_answers.py) ------------------
def test_xpbd_parent_force_chain_weight_propagation(oracle_lib):
    """:1283-1357 - two links hanging from two revolute joints: the upper joint carries both weights, the lower one only its link
    (time-averaged over one second after two seconds of settling, 10 %)."""
    g = 9.81
    builder = ModelBuilder(gravity=(0.0, 0.0, -g), up_axis="z")
    link0 = builder.add_link()
    builder.add_shape_box(link0, hx=0.1, hy=0.1, hz=0.1)
    joint0 = builder.add_joint_revolute(-1, link0, child_xform=X.transform((0.0, 0.0, 1.0)), axis=(0.0, 1.0, 0.0))
    link1 = builder.add_link()
    builder.add_shape_box(link1, hx=0.1, hy=0.1, hz=0.1)
    joint1 = builder.add_joint_revolute(link0, link1, parent_xform=X.transform((0.0, 0.0, -1.0)), child_xform=X.transform((0.0, 0.0, 1.0)),
                                        axis=(0.0, 1.0, 0.0))
    builder.add_articulation([joint0, joint1])
    model = builder.finalize()
    model.request_state_attributes("body_parent_f")
    solver = oracle_lib.SolverXPBD(model, iterations=32)
    s_in, s_out = model.state(), model.state()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s_in)
    masses = model.numpy("body_mass")
    sub_dt = 1.0 / 60.0 / 8
    for _ in range(120 * 8):
        solver.step(s_in, s_out, None, None, sub_dt)
        s_in, s_out = s_out, s_in
    avg = np.zeros((2, 6))
    for _ in range(60):
        for _ in range(8):
            solver.step(s_in, s_out, None, None, sub_dt)
            s_in, s_out = s_out, s_in
        avg += s_in.body_parent_f.numpy()
    avg /= 60
    assert avg[0, 2] == pytest.approx(float(masses[0] + masses[1]) * g, rel=0.10)
    assert avg[1, 2] == pytest.approx(float(masses[1]) * g, rel=0.10)


def test_xpbd_parent_force_allocation_and_free_body(oracle_lib):
    """:1360-1416 - not requested: stays None and step() runs; a FREE-jointed body reports a zero parent wrench."""
    builder = ModelBuilder()
    link = builder.add_link()
    builder.add_shape_sphere(link, radius=0.1)
    builder.add_articulation([builder.add_joint_revolute(-1, link, axis=(0.0, 1.0, 0.0))])
    model = builder.finalize()
    s_in, s_out = model.state(), model.state()
    assert s_in.body_parent_f is None and s_out.body_parent_f is None
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s_in)
    oracle_lib.SolverXPBD(model, iterations=2).step(s_in, s_out, None, None, 1.0 / 60.0)
    assert s_out.body_parent_f is None

    builder = ModelBuilder()
    link = builder.add_link()
    builder.add_shape_sphere(link, radius=0.1)
    builder.add_articulation([builder.add_joint_free(child=link)])
    model = builder.finalize()
    model.request_state_attributes("body_parent_f")
    s_in, s_out = model.state(), model.state()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s_in)
    oracle_lib.SolverXPBD(model, iterations=2).step(s_in, s_out, None, None, 1.0 / 60.0)
    np.testing.assert_allclose(s_out.body_parent_f.numpy()[0], 0.0, atol=1e-6)


def test_xpbd_parent_force_centripetal_zero_g(oracle_lib):
    """:1419-1526 - two bodies on a hinge spinning rigidly at 5 rad/s without gravity: the time-averaged reaction on the child is
    the centripetal force m omega^2 r (10 %), with negligible torque about its COM."""
    omega = 5.0
    builder = ModelBuilder(gravity=(0.0, 0.0, 0.0), up_axis="z")
    body_1 = builder.add_link()
    builder.add_shape_box(body_1, hx=0.25, hy=0.05, hz=0.05)
    body_2 = builder.add_link()
    builder.add_shape_box(body_2, hx=0.25, hy=0.05, hz=0.05)
    joint_free = builder.add_joint_free(child=body_1)
    joint_rev = builder.add_joint_revolute(body_1, body_2, parent_xform=X.transform((0.5, 0.0, 0.0)), child_xform=X.transform((-0.5, 0.0, 0.0)),
                                           axis=(0.0, 1.0, 0.0))
    builder.add_articulation([joint_free, joint_rev])
    model = builder.finalize()
    model.request_state_attributes("body_parent_f")
    solver = oracle_lib.SolverXPBD(model, iterations=16, joint_linear_relaxation=1.0, joint_angular_relaxation=1.0, joint_linear_compliance=0.0,
                                   joint_angular_compliance=0.0, angular_damping=0.0, enable_restitution=False)
    s_in, s_out = model.state(), model.state()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s_in)
    s_in.body_qd[body_1] = torch.tensor([0.0, 0.0, 0.5 * omega, 0.0, omega, 0.0])
    s_in.body_qd[body_2] = torch.tensor([0.0, 0.0, -0.5 * omega, 0.0, omega, 0.0])
    sub_dt = 1.0 / 240.0 / 4
    f_lin, f_tau = [], []
    for _ in range(240):
        for _ in range(4):
            solver.step(s_in, s_out, None, None, sub_dt)
            s_in, s_out = s_out, s_in
        pf = s_in.body_parent_f.numpy()[body_2]
        f_lin.append(np.linalg.norm(pf[:3]))
        f_tau.append(np.linalg.norm(pf[3:]))
    expected = float(model.body_mass[body_2]) * omega * omega * 0.5
    assert float(np.mean(f_lin)) == pytest.approx(expected, rel=0.10)
    assert float(np.mean(f_tau)) < 0.10 * expected * 0.5


def _quat_to_R(q):
    x, y, z, w = (float(c) for c in q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


@pytest.mark.parametrize("joint_kind,ic", [
    ("revolute", {"v_child": (0.0, 1.0, 0.0)}), ("revolute", {"w_child": (0.0, 0.0, 2.0)}),
    ("revolute", {"w_parent": (0.0, 0.0, 2.0), "w_child": (0.0, 0.0, 2.0)}), ("ball", {"v_child": (0.0, 1.0, 0.0)}),
    ("ball", {"w_child": (0.5, 0.5, 0.5)}), ("fixed", {"v_child": (0.0, 1.0, 0.0)}),
    ("fixed", {"w_parent": (0.0, 0.0, 1.0), "w_child": (0.0, 0.0, 1.0)}), ("prismatic", {"v_child": (0.0, 1.0, 0.0)})])
def test_xpbd_parent_force_obeys_newtons_second_law(oracle_lib, joint_kind, ic):
    """:1589-1788 - two free-floating bodies, one joint, no gravity: body_parent_f[child] dt equals the change of the child's linear
    and angular momentum over the step, and the pair's total linear momentum is conserved."""
    builder = ModelBuilder(gravity=(0.0, 0.0, 0.0), up_axis="z")
    parent = builder.add_link()
    builder.add_shape_box(parent, hx=0.2, hy=0.1, hz=0.1)
    child = builder.add_link()
    builder.add_shape_box(child, hx=0.2, hy=0.1, hz=0.1)
    j_free = builder.add_joint_free(child=parent)
    pxf, cxf = X.transform((0.5, 0.0, 0.0)), X.transform((-0.5, 0.0, 0.0))
    if joint_kind == "revolute":
        j_inner = builder.add_joint_revolute(parent, child, parent_xform=pxf, child_xform=cxf, axis=(0.0, 0.0, 1.0))
    elif joint_kind == "ball":
        j_inner = builder.add_joint_ball(parent, child, parent_xform=pxf, child_xform=cxf)
    elif joint_kind == "fixed":
        j_inner = builder.add_joint_fixed(parent, child, parent_xform=pxf, child_xform=cxf)
    else:
        j_inner = builder.add_joint_prismatic(parent, child, parent_xform=pxf, child_xform=cxf, axis=(0.0, 0.0, 1.0))
    builder.add_articulation([j_free, j_inner])
    model = builder.finalize()
    model.request_state_attributes("body_parent_f")
    solver = oracle_lib.SolverXPBD(model, iterations=32, joint_linear_relaxation=1.0, joint_angular_relaxation=1.0, joint_linear_compliance=0.0,
                                   joint_angular_compliance=0.0, angular_damping=0.0, enable_restitution=False)
    s_in, s_out = model.state(), model.state()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s_in)
    for key, (body, sl) in {"v_parent": (parent, slice(0, 3)), "w_parent": (parent, slice(3, 6)), "v_child": (child, slice(0, 3)),
                            "w_child": (child, slice(3, 6))}.items():
        if key in ic:
            s_in.body_qd[body, sl] = torch.tensor(ic[key])
    q_before, qd_before = s_in.body_q.numpy().copy(), s_in.body_qd.numpy().copy()
    dt = 1e-3
    solver.step(s_in, s_out, None, None, dt)
    q_out, qd_out = s_out.body_q.numpy(), s_out.body_qd.numpy()
    mass, inertia = model.numpy("body_mass"), model.numpy("body_inertia")
    f_expected = mass[child] * (qd_out[child, :3] - qd_before[child, :3]) / dt
    r_in, r_out = _quat_to_R(q_before[child, 3:]), _quat_to_R(q_out[child, 3:])
    tau_expected = ((r_out @ inertia[child] @ r_out.T) @ qd_out[child, 3:] - (r_in @ inertia[child] @ r_in.T) @ qd_before[child, 3:]) / dt
    parent_f = s_out.body_parent_f.numpy()[child]
    np.testing.assert_allclose(parent_f[:3], f_expected, rtol=1e-4, atol=1.0)
    np.testing.assert_allclose(parent_f[3:], tau_expected, rtol=0.01, atol=1.0)
    dp = sum(mass[i] * (qd_out[i, :3] - qd_before[i, :3]) for i in range(model.body_count))
    np.testing.assert_allclose(dp, 0.0, atol=1e-5)


def test_parent_force_static_pendulum_xpbd_and_featherstone_agree(oracle_lib):
    """:1529-1586 without the MuJoCo leg - one step of a hanging link: both solvers report F_z = m g within 5 %."""
    results = {}
    for name in ("xpbd", "featherstone"):
        builder = ModelBuilder(gravity=(0.0, 0.0, -9.81), up_axis="z")
        link = builder.add_link()
        builder.add_shape_box(link, hx=0.1, hy=0.1, hz=0.1)
        builder.add_articulation([builder.add_joint_revolute(-1, link, child_xform=X.transform((0.0, 0.0, 1.0)), axis=(0.0, 1.0, 0.0))])
        model = builder.finalize()
        model.request_state_attributes("body_parent_f")
        solver = oracle_lib.SolverXPBD(model, iterations=8) if name == "xpbd" else oracle_lib.SolverFeatherstone(model)
        s0, s1 = model.state(), model.state()
        oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s0)
        solver.step(s0, s1, None, None, 5e-3)
        if s1.body_parent_f is None:
            pytest.skip("the Featherstone oracle does not report body_parent_f")
        results[name] = s1.body_parent_f.numpy()[0]
        assert results[name][2] == pytest.approx(float(model.body_mass[0]) * 9.81, rel=0.05), name


# ---- test_narrow_phase.py:2739-3160 (box-box depth accuracy across the three MPR inflation branches) --------------------------------
def _signed_distance_to_box(p, pos, half):
    d = np.abs(np.asarray(p, dtype=np.float64) - np.asarray(pos)) - np.asarray(half)
    return float(np.linalg.norm(np.maximum(d, 0.0)) + min(max(d[0], d[1], d[2]), 0.0))


def _box_pair(oracle_lib, z, thickness, gap_sum=0.2):
    I7 = np.array([0, 0, 0, 0, 0, 0, 1], dtype=np.float32)
    xb = np.array([0, 0, z, 0, 0, 0, 1], dtype=np.float32)
    return oracle_lib.convex_pair(GeoType.BOX, (0.5, 0.5, 0.5), I7, GeoType.BOX, (0.5, 0.5, 0.5), xb, gap_sum, "oracle", thickness, thickness)


@pytest.mark.parametrize("thickness", [0.0, 2.5e-5, 0.005])  # margin_sum = 0 (inflate 1e-4), in (0, 1e-4) (inflate 2e-4), >= 1e-4 (no inflation)
def test_box_box_depth_accuracy_over_inflation_branches(oracle_lib, thickness):
    """The anti-flicker inflation of the MPR support points (collision_convex.py:154-161) must be taken back out of the reported
    depth: overlap 0.01 -> -0.01, coincident faces -> 0, both to 5e-5 (the contact writer then subtracts the margins, which the
    reference test adds back), and centre -+ n d / 2 must land on the two box surfaces to 5e-5."""
    cnt, dist, pos, n = _box_pair(oracle_lib, 1.0 - 0.01, thickness)
    assert cnt > 0 and float(dist[:cnt].min()) == pytest.approx(-0.01, abs=5e-5)
    cnt, dist, pos, n = _box_pair(oracle_lib, 1.0, thickness, gap_sum=0.02)
    assert cnt > 0 and float(dist[:cnt].min()) == pytest.approx(0.0, abs=5e-5)
    cnt, dist, pos, n = _box_pair(oracle_lib, 1.0 - 0.05, thickness)
    validated = 0
    for i in range(cnt):
        if dist[i] >= 0.0:
            continue
        a = pos[i].astype(np.float64) - n[i] * (dist[i] / 2.0)
        b = pos[i].astype(np.float64) + n[i] * (dist[i] / 2.0)
        assert _signed_distance_to_box(a, (0, 0, 0), (0.5,) * 3) == pytest.approx(0.0, abs=5e-5)
        assert _signed_distance_to_box(b, (0, 0, 0.95), (0.5,) * 3) == pytest.approx(0.0, abs=5e-5)
        validated += 1
    assert validated > 0


# ---- test_kinematic_links.py:304-715 (kinematic bodies under XPBD(iterations=5): prescribed motion, immunity to wrenches, contact
# response of dynamic bodies, runtime toggling) ---------------------------------------------------------------------------------------
from newton_b200 import BodyFlags  # noqa: E402

KINEMATIC_TEST_WRENCH = torch.tensor([20.0, -15.0, 10.0, 0.5, -0.4, 0.3])


def _contact_defaults(builder):
    builder.default_shape_cfg.ke, builder.default_shape_cfg.kd, builder.default_shape_cfg.kf = 1.0e4, 500.0, 0.5


def _kin_solver(oracle_lib, model, name="xpbd"):
    """The `solvers` table of the reference test (:716-724), in-scope rows."""
    if name == "featherstone":
        return oracle_lib.SolverFeatherstone(model, angular_damping=0.0)
    return oracle_lib.SolverXPBD(model, iterations=5, angular_damping=0.0)


def _quat_close(qa, qb, min_dot):
    assert abs(float(np.dot(qa, qb))) > min_dot


@pytest.mark.parametrize("solver_name", ["xpbd", "featherstone"])
def test_kinematic_free_base_prescribed_motion_xpbd(oracle_lib, solver_name):
    """:400-481 (SolverFeatherstone: effective armature 1e10, zeroed qdd and the copied-through joint state of the kinematic joint)"""
    dt, steps, x0, vx = 1.0 / 240.0, 100, -0.3, 1.0

    def run_once(apply_force):
        builder = ModelBuilder(gravity=(0.0, 0.0, 0.0))
        _contact_defaults(builder)
        kin = builder.add_body(xform=X.transform((-0.3, 0.0, 0.0)), mass=1.0, is_kinematic=True, label="kinematic_free")
        builder.add_shape_box(kin, hx=0.25, hy=0.15, hz=0.15)
        probe = builder.add_body(xform=X.transform((0.45, 0.0, 0.0)), mass=1.0, label="probe")
        builder.add_shape_sphere(probe, radius=0.1)
        model = builder.finalize()
        joint = int(np.flatnonzero(model.numpy("joint_child") == kin)[0])
        qs, qds = int(model.joint_q_start[joint]), int(model.joint_qd_start[joint])
        solver = _kin_solver(oracle_lib, model, solver_name)
        pipe = oracle_lib.CollisionPipeline(model)
        contacts = pipe.contacts()
        s0, s1 = model.state(), model.state()
        p0 = s0.body_q.numpy()[probe, :3].copy()
        max_speed = 0.0
        for i in range(steps):
            s0.joint_q[qs : qs + 7] = torch.tensor([x0 + vx * i * dt, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0])
            s0.joint_qd[qds : qds + 6] = torch.tensor([vx, 0.0, 0.0, 0.0, 0.0, 0.0])
            oracle_lib.eval_fk(model, s0.joint_q, s0.joint_qd, s0, body_flag_filter=int(BodyFlags.KINEMATIC))
            s0.clear_forces()
            if apply_force:
                s0.body_f[kin] = KINEMATIC_TEST_WRENCH
            pipe.collide(s0, contacts)
            solver.step(s0, s1, None, contacts, dt)
            s0, s1 = s1, s0
            max_speed = max(max_speed, float(np.linalg.norm(s0.body_qd.numpy()[probe, :3])))
        q, qd = s0.body_q.numpy(), s0.body_qd.numpy()
        return dict(kin_pos=q[kin, :3].copy(), kin_quat=q[kin, 3:].copy(), kin_qd=qd[kin].copy(), probe_qd=qd[probe].copy(), probe_max_speed=max_speed,
                    probe_displacement=float(np.linalg.norm(q[probe, :3] - p0)))

    free, forced = run_once(False), run_once(True)
    assert free["kin_pos"][0] == pytest.approx(x0 + vx * (steps - 1) * dt, abs=4e-2)
    assert abs(free["kin_pos"][1]) < 2e-2 and abs(free["kin_pos"][2]) < 2e-2 and free["kin_qd"][0] > 2e-1
    np.testing.assert_allclose(forced["kin_pos"], free["kin_pos"], atol=8e-3)  # the applied wrench does not perturb prescribed motion
    _quat_close(forced["kin_quat"], free["kin_quat"], 0.9995)
    assert np.linalg.norm(forced["kin_qd"] - free["kin_qd"]) < 4e-1
    assert free["probe_max_speed"] > 3e-2  # the dynamic probe is pushed
    assert np.linalg.norm(free["probe_qd"][:3]) > 5e-3 or free["probe_displacement"] > 1e-3


def test_kinematic_revolute_root_pendulum_prescribed_motion_xpbd(oracle_lib):
    """:484-575 - maximal-coordinate solvers get the root's body_q / body_qd prescribed directly."""
    dt, steps, theta0, omega = 1.0 / 240.0, 120, -1.0, 2.0

    def run_once(apply_force):
        builder = ModelBuilder(gravity=(0.0, 0.0, 0.0))
        _contact_defaults(builder)
        root = builder.add_link(mass=1.0, inertia=np.eye(3) * 0.1, is_kinematic=True, label="kinematic_root")
        pendulum = builder.add_link(mass=1.0, inertia=np.eye(3) * 0.1, label="pendulum")
        j_root = builder.add_joint_revolute(-1, root, axis=(0.0, 1.0, 0.0), label="kinematic_root_joint")
        j_pend = builder.add_joint_revolute(root, pendulum, axis=(0.0, 1.0, 0.0), parent_xform=X.transform((0.45, 0.0, 0.0)), label="pendulum_joint")
        builder.add_articulation([j_root, j_pend])
        builder.add_shape_box(root, xform=X.transform((0.6, 0.0, 0.0)), hx=0.15, hy=0.08, hz=0.08)
        builder.add_shape_sphere(pendulum, xform=X.transform((0.3, 0.0, 0.0)), radius=0.1)
        probe = builder.add_body(xform=X.transform((0.72, 0.0, 0.0)), mass=0.6, label="probe")
        builder.add_shape_sphere(probe, radius=0.1)
        model = builder.finalize()
        solver = _kin_solver(oracle_lib, model)
        pipe = oracle_lib.CollisionPipeline(model)
        contacts = pipe.contacts()
        s0, s1 = model.state(), model.state()
        p0 = s0.body_q.numpy()[probe, :3].copy()
        max_probe = max_pend = 0.0
        for i in range(steps):
            theta = theta0 + omega * i * dt
            s0.body_q[root] = torch.tensor([0.0, 0.0, 0.0, 0.0, np.sin(theta / 2), 0.0, np.cos(theta / 2)], dtype=torch.float32)
            s0.body_qd[root] = torch.tensor([0.0, 0.0, 0.0, 0.0, omega, 0.0])
            s0.clear_forces()
            if apply_force:
                s0.body_f[root] = KINEMATIC_TEST_WRENCH
            pipe.collide(s0, contacts)
            solver.step(s0, s1, None, contacts, dt)
            s0, s1 = s1, s0
            qd = s0.body_qd.numpy()
            max_probe, max_pend = max(max_probe, float(np.linalg.norm(qd[probe, :3]))), max(max_pend, float(np.linalg.norm(qd[pendulum])))
        q, qd = s0.body_q.numpy(), s0.body_qd.numpy()
        return dict(root_pos=q[root, :3].copy(), root_quat=q[root, 3:].copy(), root_qd=qd[root].copy(), probe_max_speed=max_probe,
                    pendulum_max_speed=max_pend, probe_displacement=float(np.linalg.norm(q[probe, :3] - p0)))

    free, forced = run_once(False), run_once(True)
    theta_end = theta0 + omega * (steps - 1) * dt
    _quat_close(free["root_quat"], np.array([0.0, np.sin(theta_end / 2), 0.0, np.cos(theta_end / 2)]), 0.99)
    assert np.linalg.norm(free["root_qd"][3:]) > 5e-1
    np.testing.assert_allclose(forced["root_pos"], free["root_pos"], atol=1.5e-2)
    _quat_close(forced["root_quat"], free["root_quat"], 0.995)
    assert np.linalg.norm(forced["root_qd"] - free["root_qd"]) < 8e-1
    assert free["probe_max_speed"] > 2e-2 and free["probe_displacement"] > 1e-2 and free["pendulum_max_speed"] > 2e-2


@pytest.mark.parametrize("solver_name", ["xpbd", "featherstone"])
def test_kinematic_fixed_root_is_immune_to_forces_and_stops_the_probe_xpbd(oracle_lib, solver_name):
    """:578-638"""
    dt, steps, probe_vx = 1.0 / 240.0, 140, 3.0

    def run_once(apply_force):
        builder = ModelBuilder(gravity=(0.0, 0.0, 0.0))
        _contact_defaults(builder)
        static = builder.add_link(xform=X.transform((0.0, 0.0, 0.0)), mass=1.0, inertia=np.eye(3) * 0.1, is_kinematic=True, label="static_root")
        builder.add_articulation([builder.add_joint_fixed(-1, static, label="static_root_joint")])
        builder.add_shape_box(static, hx=0.25, hy=0.25, hz=0.25)
        probe = builder.add_body(xform=X.transform((-1.0, 0.0, 0.0)), mass=1.0, label="probe")
        builder.add_shape_sphere(probe, radius=0.12)
        model = builder.finalize()
        joint = int(np.flatnonzero(model.numpy("joint_child") == probe)[0])
        qds = int(model.joint_qd_start[joint])
        solver = _kin_solver(oracle_lib, model, solver_name)
        pipe = oracle_lib.CollisionPipeline(model)
        contacts = pipe.contacts()
        s0, s1 = model.state(), model.state()
        s0.joint_qd[qds : qds + 6] = torch.tensor([probe_vx, 0.0, 0.0, 0.0, 0.0, 0.0])
        oracle_lib.eval_fk(model, s0.joint_q, s0.joint_qd, s0)
        q_init = s0.body_q.numpy()[static].copy()
        for _ in range(steps):
            s0.clear_forces()
            if apply_force:
                s0.body_f[static] = KINEMATIC_TEST_WRENCH
            pipe.collide(s0, contacts)
            solver.step(s0, s1, None, contacts, dt)
            s0, s1 = s1, s0
        q, qd = s0.body_q.numpy(), s0.body_qd.numpy()
        return dict(q_init=q_init, static_pos=q[static, :3].copy(), static_quat=q[static, 3:].copy(), static_qd=qd[static].copy(),
                    probe_pos=q[probe, :3].copy(), probe_qd=qd[probe].copy())

    free, forced = run_once(False), run_once(True)
    np.testing.assert_allclose(free["static_pos"], free["q_init"][:3], atol=2e-3)
    _quat_close(free["static_quat"], free["q_init"][3:], 0.9999)
    assert np.linalg.norm(free["static_qd"]) < 4e-1
    np.testing.assert_allclose(forced["static_pos"], free["static_pos"], atol=2e-3)
    _quat_close(forced["static_quat"], free["static_quat"], 0.9999)
    assert np.linalg.norm(forced["static_qd"] - free["static_qd"]) < 6e-2
    assert free["probe_pos"][0] < 0.25 and abs(free["probe_qd"][0] - probe_vx) > 2.5e-1  # the probe hit the box


@pytest.mark.parametrize("solver_name", ["xpbd", "featherstone"])
def test_kinematic_runtime_toggle_xpbd(oracle_lib, solver_name):
    """:641-715 - body_flags edited at run time + notify_model_changed(BODY_PROPERTIES)."""
    builder = ModelBuilder(gravity=(0.0, 0.0, 0.0))
    body = builder.add_body(xform=X.transform((0.0, 0.0, 0.0)), mass=1.0, is_kinematic=True, label="toggle_body")
    builder.add_shape_sphere(body, radius=0.1)
    model = builder.finalize()
    solver = _kin_solver(oracle_lib, model, solver_name)
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    s0, s1 = model.state(), model.state()

    def phase(s0, s1):
        for _ in range(60):
            s0.clear_forces()
            s0.body_f[body] = torch.tensor([10.0, 0.0, 0.0, 0.0, 0.0, 0.0])
            pipe.collide(s0, contacts)
            solver.step(s0, s1, None, contacts, 1.0 / 240.0)
            s0, s1 = s1, s0
        return s0, s1

    s0, s1 = phase(s0, s1)
    assert np.linalg.norm(s0.body_q.numpy()[body, :3]) < 1e-3  # kinematic: does not move under the force
    model.body_flags[body] = int(BodyFlags.DYNAMIC)
    solver.notify_model_changed(ModelFlags.BODY_PROPERTIES)
    s0, s1 = phase(s0, s1)
    assert s0.body_q.numpy()[body, 0] > 0.05  # dynamic: accelerates
    model.body_flags[body] = int(BodyFlags.KINEMATIC)
    solver.notify_model_changed(ModelFlags.BODY_PROPERTIES)
    before = s0.body_q.numpy()[body, :3].copy()
    s0, s1 = phase(s0, s1)
    assert np.linalg.norm(s0.body_q.numpy()[body, :3] - before) < 1e-3  # kinematic again: stays put


# ---- test_parent_force.py (the SolverFeatherstone rows: State.body_parent_f from the RNEA backward pass, rtol 1e-4) -------------------
def _hanging_link(joint_axis, child_offset, parent_xform=None):
    builder = ModelBuilder(gravity=(0.0, 0.0, -9.81), up_axis="z")
    link = builder.add_link()
    builder.add_shape_box(link, hx=0.1, hy=0.1, hz=0.1)
    joint = builder.add_joint_revolute(-1, link, parent_xform=parent_xform, child_xform=X.transform(child_offset), axis=joint_axis)
    builder.add_articulation([joint])
    model = builder.finalize()
    model.request_state_attributes("body_parent_f")
    return model


@pytest.mark.parametrize("parent_xform", [None, X.transform((5.0, 3.0, -2.0)),
                                          X.transform((1.0, 2.0, 3.0), X.quat_from_axis_angle(np.array([1.0, 0.0, 0.0]), np.pi * 0.5))])
def test_featherstone_parent_force_static_pendulum(oracle_lib, parent_xform):
    """:52-84 - the joint carries exactly the weight, whatever the joint frame."""
    model = _hanging_link((0.0, 1.0, 0.0), (0.0, 0.0, 1.0), parent_xform)
    s0, s1 = model.state(), model.state()
    assert s0.body_parent_f is not None
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s0)
    oracle_lib.SolverFeatherstone(model).step(s0, s1, None, None, 5e-3)
    pf = s1.body_parent_f.numpy()[0]
    np.testing.assert_allclose(pf[:3], [0.0, 0.0, float(model.body_mass[0]) * 9.81], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(pf[3:], 0.0, atol=1e-2)


def test_featherstone_parent_force_centrifugal(oracle_lib):
    """:87-114 - horizontal pendulum spinning about Z at 5 rad/s: weight in +Z plus m omega^2 r towards the axis."""
    model = _hanging_link((0.0, 0.0, 1.0), (-1.0, 0.0, 0.0))
    s0, s1 = model.state(), model.state()
    s0.joint_qd[0] = 5.0
    oracle_lib.eval_fk(model, s0.joint_q, s0.joint_qd, s0)
    oracle_lib.SolverFeatherstone(model).step(s0, s1, None, None, 5e-3)
    pf, mass = s1.body_parent_f.numpy()[0], float(model.body_mass[0])
    np.testing.assert_allclose(pf[:3], [-mass * 25.0, 0.0, mass * 9.81], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(pf[3:], 0.0, atol=1e-2)


def test_featherstone_body_f_propagates_into_parent_force(oracle_lib):
    """:117-193 - a force / torque on the second link of a chain shows up in the first joint's reaction (non-compliant directions)."""
    builder = ModelBuilder(gravity=(0.0, 0.0, -9.81), up_axis="z")
    link0 = builder.add_link()
    builder.add_shape_box(link0, hx=0.1, hy=0.1, hz=0.1)
    joint0 = builder.add_joint_revolute(-1, link0, child_xform=X.transform((0.0, 0.0, 1.0)), axis=(0.0, 1.0, 0.0))
    link1 = builder.add_link()
    builder.add_shape_box(link1, hx=0.1, hy=0.1, hz=0.1)
    joint1 = builder.add_joint_revolute(link0, link1, parent_xform=X.transform((0.0, 0.0, -1.0)), child_xform=X.transform((0.0, 0.0, 1.0)),
                                        axis=(0.0, 1.0, 0.0))
    builder.add_articulation([joint0, joint1])
    model = builder.finalize()
    model.request_state_attributes("body_parent_f")
    solver = oracle_lib.SolverFeatherstone(model)
    total_weight = float(model.body_mass[0] + model.body_mass[1]) * 9.81
    for wrench, f_expected, tau_expected in (((0.0, 10.0, 0.0, 0.0, 0.0, 0.0), (0.0, -10.0, total_weight), (-20.0, 0.0, 0.0)),
                                             ((0.0, 0.0, 0.0, 5.0, 0.0, 0.0), (0.0, 0.0, total_weight), (-5.0, 0.0, 0.0))):
        s0, s1 = model.state(), model.state()
        oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s0)
        s0.body_f[1] = torch.tensor(wrench)
        solver.step(s0, s1, None, None, 5e-3)
        pf = s1.body_parent_f.numpy()[0]
        np.testing.assert_allclose(pf[:3], f_expected, rtol=1e-4, atol=1e-3)
        np.testing.assert_allclose(pf[3:], tau_expected, atol=1e-2)


# ---- test_rigid_contact.py:775-843 (test_box_drop, XPBD(iterations=2)) ----------------------------------------------------------------
def test_xpbd_box_drop_energy_bound(oracle_lib):
    """Two boxes (one tilted 0.5 rad) dropped onto the ground and onto each other: contacts never inject energy - |v_z| stays below
    sqrt(2 g h_max) - and after one second both rest above the ground near the origin."""
    builder = ModelBuilder()
    builder.add_ground_plane()
    half = 0.5
    body_1 = builder.add_body(xform=X.transform((0.0, 0.0, half * 1.2)))
    builder.add_shape_box(body_1, hx=half, hy=half, hz=half)
    body_2 = builder.add_body(xform=X.transform((0.0, 0.0, half * 4.2), X.quat_from_axis_angle(np.array([1.0, 0.0, 0.0]), 0.5)))
    builder.add_shape_box(body_2, hx=half, hy=half, hz=half)
    model = builder.finalize()
    solver = oracle_lib.SolverXPBD(model, iterations=2)
    s0, s1 = model.state(), model.state()
    oracle_lib.eval_fk(model, model.joint_q, model.joint_qd, s0)
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    v_max = np.sqrt(2 * 9.81 * half * 3.2)
    worst = 0.0
    for _ in range(60):
        for _ in range(8):
            s0.clear_forces()
            pipe.collide(s0, contacts)
            solver.step(s0, s1, None, contacts, 1.0 / 60.0 / 8)
            s0, s1 = s1, s0
        worst = max(worst, float(np.abs(s0.body_qd.numpy()[:, 2]).max()))
    assert worst < v_max
    q, qd = s0.body_q.numpy(), s0.body_qd.numpy()
    for b in range(model.body_count):
        assert abs(q[b, 0]) < 1.0 and abs(q[b, 1]) < 1.0 and q[b, 2] > half * 0.5
        assert np.linalg.norm(qd[b, :3]) < 1.0


# ---- mass properties of every primitive against brute-force integration (the reference checks its closed forms against
# compute_inertia_mesh of fine meshes, newton/tests/test_inertia.py:25-134, 307-409; here: a 160^3 midpoint grid) -------------------------
def _grid_mass_properties(inside, bounds, n=160):
    lo, hi = np.asarray(bounds[0], dtype=np.float64), np.asarray(bounds[1], dtype=np.float64)
    axes = [lo[k] + (np.arange(n) + 0.5) * (hi[k] - lo[k]) / n for k in range(3)]
    x, y, z = np.meshgrid(*axes, indexing="ij")
    mask = inside(x, y, z)
    dv = np.prod((hi - lo) / n)
    pts = np.stack([x[mask], y[mask], z[mask]], axis=1)
    mass = dv * len(pts)
    com = pts.mean(axis=0)
    d = pts - com
    inertia = dv * (np.einsum("ij,ij->", d, d) * np.eye(3) - d.T @ d)
    return mass, com, inertia


def _barrel_profile(r, hh, rb):
    return lambda z: r + (hh * hh - z * z) / (np.sqrt(np.maximum(rb * rb - z * z, 0.0)) + np.sqrt(rb * rb - hh * hh))


PRIMITIVES = {
    "sphere": (GeoType.SPHERE, (0.7, 0.0, 0.0), lambda x, y, z: x * x + y * y + z * z <= 0.49, 0.7),
    "box": (GeoType.BOX, (0.3, 0.5, 0.7), lambda x, y, z: (np.abs(x) <= 0.3) & (np.abs(y) <= 0.5) & (np.abs(z) <= 0.7), (0.3, 0.5, 0.7)),
    "capsule": (GeoType.CAPSULE, (0.4, 0.6, 0.0), lambda x, y, z: x * x + y * y + np.maximum(np.abs(z) - 0.6, 0.0) ** 2 <= 0.16, 1.0),
    "cylinder": (GeoType.CYLINDER, (0.5, 0.8, 0.0), lambda x, y, z: (x * x + y * y <= 0.25) & (np.abs(z) <= 0.8), 0.8),
    "barrel": (GeoType.CYLINDER, (0.4, 0.6, 1.0), lambda x, y, z: (np.abs(z) <= 0.6) & (x * x + y * y <= _barrel_profile(0.4, 0.6, 1.0)(z) ** 2), 0.8),
    "cone": (GeoType.CONE, (0.6, 0.9, 0.0), lambda x, y, z: (np.abs(z) <= 0.9) & (x * x + y * y <= (0.6 * (0.9 - z) / 1.8) ** 2), 0.9),
    "ellipsoid": (GeoType.ELLIPSOID, (0.7, 0.5, 0.3), lambda x, y, z: (x / 0.7) ** 2 + (y / 0.5) ** 2 + (z / 0.3) ** 2 <= 1.0, 0.7),
}


@pytest.mark.parametrize("name", list(PRIMITIVES))
def test_primitive_mass_properties_against_brute_force(name):
    from newton_b200.geometry.inertia import compute_inertia_shape

    geo_type, scale, inside, extent = PRIMITIVES[name]
    mass, com, inertia = compute_inertia_shape(geo_type, scale, 1000.0)
    half = np.asarray(extent if isinstance(extent, tuple) else (extent,) * 3, dtype=np.float64)  # grid bounds (tight for the box)
    m_ref, com_ref, i_ref = _grid_mass_properties(inside, (-half, half))
    assert mass == pytest.approx(1000.0 * m_ref, rel=5e-3)
    np.testing.assert_allclose(com, com_ref, atol=2e-3 * half.max())
    np.testing.assert_allclose(np.asarray(inertia), 1000.0 * i_ref, rtol=1e-2, atol=1e-3 * np.trace(1000.0 * i_ref))


HOLLOW_T = 0.1


def _shrunk(name):
    """The inner solid the reference removes from a hollow primitive: every dimension minus the wall thickness
    (geometry/inertia.py:634-722), same local frame."""
    t = HOLLOW_T
    return {
        "sphere": lambda x, y, z: x * x + y * y + z * z <= (0.7 - t) ** 2,
        "box": lambda x, y, z: (np.abs(x) <= 0.3 - t) & (np.abs(y) <= 0.5 - t) & (np.abs(z) <= 0.7 - t),
        "capsule": lambda x, y, z: x * x + y * y + np.maximum(np.abs(z) - (0.6 - t), 0.0) ** 2 <= (0.4 - t) ** 2,
        "cylinder": lambda x, y, z: (x * x + y * y <= (0.5 - t) ** 2) & (np.abs(z) <= 0.8 - t),
        "cone": lambda x, y, z: (np.abs(z) <= 0.9 - t) & (x * x + y * y <= ((0.6 - t) * (0.9 - t - z) / (1.8 - 2 * t)) ** 2),
        "ellipsoid": lambda x, y, z: (x / (0.7 - t)) ** 2 + (y / (0.5 - t)) ** 2 + (z / (0.3 - t)) ** 2 <= 1.0,
    }[name]


@pytest.mark.parametrize("name", ["sphere", "box", "capsule", "cylinder", "cone", "ellipsoid"])
def test_hollow_primitive_mass_properties_against_brute_force(name):
    """is_solid=False: outer minus inner solid, integrated on the grid - the cone's shell has its own centre of mass."""
    from newton_b200.geometry.inertia import compute_inertia_shape

    geo_type, scale, inside, extent = PRIMITIVES[name]
    inner = _shrunk(name)
    mass, com, inertia = compute_inertia_shape(geo_type, scale, 1000.0, is_solid=False, thickness=HOLLOW_T)
    half = np.asarray(extent if isinstance(extent, tuple) else (extent,) * 3, dtype=np.float64)
    m_ref, com_ref, i_ref = _grid_mass_properties(lambda x, y, z: inside(x, y, z) & ~inner(x, y, z), (-half, half), n=200)
    assert mass == pytest.approx(1000.0 * m_ref, rel=1.5e-2)
    np.testing.assert_allclose(com, com_ref, atol=4e-3 * half.max())
    np.testing.assert_allclose(np.asarray(inertia), 1000.0 * i_ref, rtol=2e-2, atol=2e-3 * np.trace(1000.0 * i_ref))
    if name == "cone":
        assert com[2] < -1.8 / 4.0 + 1e-9 and abs(com[2] + 0.45) > 1e-3  # not the solid cone's centre of mass


def test_hollow_thickness_is_validated():
    """reference geometry/inertia.py:48-76: negative, non-finite or too-thick walls raise; zero gives a massless shell + warning."""
    from newton_b200.geometry.inertia import compute_inertia_shape

    for bad in (-0.1, float("nan"), 0.7, 1.0):
        with pytest.raises(ValueError):
            compute_inertia_shape(GeoType.SPHERE, (0.7, 0.0, 0.0), 1000.0, is_solid=False, thickness=bad)
    with pytest.raises(TypeError):
        compute_inertia_shape(GeoType.BOX, (0.3, 0.5, 0.7), 1000.0, is_solid=False, thickness="thin")
    with pytest.raises(ValueError):
        compute_inertia_shape(GeoType.CONE, (0.6, 0.9, 0.0), 1000.0, is_solid=False, thickness=0.6)
    with pytest.warns(UserWarning):
        m, _, inertia = compute_inertia_shape(GeoType.BOX, (0.3, 0.5, 0.7), 1000.0, is_solid=False, thickness=0.0)
    assert m == 0.0 and not np.asarray(inertia).any()


def test_barrel_cylinder_body_can_be_built():
    """A dynamic body carrying a barrel cylinder gets the barrel's mass properties (it used to be refused by the builder)."""
    builder = ModelBuilder()
    body = builder.add_body()
    cfg = newton_b200.ShapeConfig()
    cfg.density = 500.0
    shape = builder.add_shape_cylinder(body, radius=0.4, half_height=0.6, cfg=cfg)
    builder.shape_scale[shape] = (0.4, 0.6, 1.0)  # barrel radius in scale.z (sim/builder.py:7089)
    plain = float(builder.finalize().body_mass[0])
    from newton_b200.geometry.inertia import compute_inertia_cylinder

    m_barrel, _, inertia = compute_inertia_cylinder(500.0, 0.4, 0.6, 1.0)
    assert m_barrel > plain and inertia[2, 2] > 0.5 * plain * 0.16  # bulging sides: heavier, larger axial inertia than the straight one
    with pytest.raises(ValueError):
        compute_inertia_cylinder(500.0, 0.4, 0.6, 0.5)

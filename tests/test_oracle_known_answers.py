"""Pins the CPU oracle against the known-answer checks the REFERENCE's own tests hold for this path.

The reference ships no golden files for collision / XPBD; what it does hold are
closed-form expectation tables and example end-state assertions.  Those are restated here with citations:

* analytic collider tables   - newton/tests/test_collision_primitives.py:429-1500 (distance to 1e-5)
* example_basic_urdf final   - newton/examples/basic/example_basic_urdf.py:145-161 via tests/test_examples.py:349-357
* free fall                  - newton/tests/test_physics_verification.py (semi-implicit Euler closed form)
"""

import math

import numpy as np
import pytest

from newton_b200 import GeoType, scenes
from newton_b200.sim.builder import ModelBuilder
from newton_b200.utils import xform as X
from newton_b200.utils.host_fk import host_fk

I7 = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]


def _xf(p, q=(0.0, 0.0, 0.0, 1.0)):
    return [*p, *q]


def _plane_xf(normal, pos):
    return _xf(pos, X.quat_between_vectors((0.0, 0.0, 1.0), normal))


def _axis_q(axis):
    return X.quat_between_vectors((0.0, 0.0, 1.0), axis)


# reference test_collision_primitives.py:436-447
PLANE_SPHERE = [
    ([0, 0, 1], [0, 0, 0], [0, 0, 2.0], 1.0, 1.0), ([0, 0, 1], [0, 0, 0], [0, 0, 1.5], 1.0, 0.5),
    ([0, 0, 1], [0, 0, 0], [0, 0, 1.0], 1.0, 0.0), ([0, 0, 1], [0, 0, 0], [0, 0, 0.8], 1.0, -0.2),
    ([0, 0, 1], [0, 0, 0], [0, 0, 0.5], 1.0, -0.5), ([0, 0, 1], [0, 0, 0], [0, 0, 0.2], 1.0, -0.8),
    ([1, 0, 0], [1, 0, 0], [2.0, 0, 0], 0.5, 0.5), ([1, 0, 0], [1, 0, 0], [1.5, 0, 0], 0.5, 0.0),
    ([1, 0, 0], [1, 0, 0], [1.3, 0, 0], 0.5, -0.2),
]


@pytest.mark.parametrize("n,pp,sp,r,expected", PLANE_SPHERE)
def test_plane_sphere(oracle_lib, n, pp, sp, r, expected):
    ok, dist, pos, normal = oracle_lib.primitive_pair(GeoType.PLANE, (0, 0, 0), _plane_xf(n, pp), GeoType.SPHERE, (r, 0, 0), _xf(sp))
    assert ok
    assert dist[0] == pytest.approx(expected, abs=1e-5)
    np.testing.assert_allclose(normal, n, atol=1e-6)
    if expected < 0:  # contact lies between the sphere surface and the plane (reference :470-497)
        assert np.linalg.norm(pos[0] - np.array(sp)) < r + 0.01


# reference test_collision_primitives.py:512-527
SPHERE_SPHERE = [
    ([0, 0, 0], 1.0, [3.5, 0, 0], 1.0, 1.5), ([0, 0, 0], 1.0, [3.0, 0, 0], 1.0, 1.0), ([0, 0, 0], 1.0, [2.5, 0, 0], 1.0, 0.5),
    ([0, 0, 0], 1.0, [2.0, 0, 0], 1.0, 0.0), ([0, 0, 0], 1.0, [1.8, 0, 0], 1.0, -0.2), ([0, 0, 0], 1.0, [1.5, 0, 0], 1.0, -0.5),
    ([0, 0, 0], 1.0, [1.2, 0, 0], 1.0, -0.8), ([0, 0, 0], 0.5, [2.0, 0, 0], 1.0, 0.5), ([0, 0, 0], 0.5, [1.5, 0, 0], 1.0, 0.0),
    ([0, 0, 0], 0.5, [1.2, 0, 0], 1.0, -0.3),
]


@pytest.mark.parametrize("p1,r1,p2,r2,expected", SPHERE_SPHERE)
def test_sphere_sphere(oracle_lib, p1, r1, p2, r2, expected):
    ok, dist, pos, normal = oracle_lib.primitive_pair(GeoType.SPHERE, (r1, 0, 0), _xf(p1), GeoType.SPHERE, (r2, 0, 0), _xf(p2))
    assert ok
    assert dist[0] == pytest.approx(expected, abs=1e-5)
    assert np.linalg.norm(normal) == pytest.approx(1.0, abs=1e-5)
    assert np.dot(normal, np.array(p2) - np.array(p1)) > 0  # points from geom 0 into geom 1
    # contact position is the midpoint between the two surfaces (reference :570-585)
    a = np.array(p1) + normal * r1
    b = np.array(p2) - normal * r2
    np.testing.assert_allclose(pos[0], 0.5 * (a + b), atol=1e-5)


# reference test_collision_primitives.py:601-612 (capsule at origin, axis Z, r=0.5, half-length 1; sphere r=0.5)
SPHERE_CAPSULE = [([0, 1.5, 0], 0.5), ([0, 1.0, 0], 0.0), ([0, 0.9, 0], -0.1), ([0, 0.8, 0], -0.2),
                  ([0, 0, 2.5], 0.5), ([0, 0, 2.0], 0.0), ([0, 0, 1.9], -0.1), ([0, 0, 1.8], -0.2)]


@pytest.mark.parametrize("sp,expected", SPHERE_CAPSULE)
def test_sphere_capsule(oracle_lib, sp, expected):
    ok, dist, _, normal = oracle_lib.primitive_pair(GeoType.SPHERE, (0.5, 0, 0), _xf(sp), GeoType.CAPSULE, (0.5, 1.0, 0), I7)
    assert ok
    assert dist[0] == pytest.approx(expected, abs=1e-5)


# reference test_collision_primitives.py:712-790 (parallel capsules along X, r=0.5, half-length 1)
@pytest.mark.parametrize("y,expected", [(2.0, 1.0), (1.5, 0.5), (1.0, 0.0), (0.9, -0.1), (0.8, -0.2), (0.7, -0.3)])
def test_capsule_capsule_parallel(oracle_lib, y, expected):
    q = _axis_q((1.0, 0.0, 0.0))
    ok, dist, _, normal = oracle_lib.primitive_pair(GeoType.CAPSULE, (0.5, 1.0, 0), _xf((0, 0, 0), q), GeoType.CAPSULE,
                                                    (0.5, 1.0, 0), _xf((0, y, 0), q))
    assert ok
    assert dist[0] == pytest.approx(expected, abs=1e-4)
    assert dist[1] == pytest.approx(expected, abs=1e-4)  # parallel axes produce two contacts
    np.testing.assert_allclose(normal, [0, 1, 0], atol=1e-4)


# reference test_collision_primitives.py:904-913 (ellipsoid half-axes (1,1,1.5))
@pytest.mark.parametrize("z,expected", [(2.5, 1.0), (2.0, 0.5), (1.5, 0.0), (1.4, -0.1), (1.3, -0.2), (1.0, -0.5)])
def test_plane_ellipsoid(oracle_lib, z, expected):
    ok, dist, _, normal = oracle_lib.primitive_pair(GeoType.PLANE, (0, 0, 0), I7, GeoType.ELLIPSOID, (1.0, 1.0, 1.5), _xf((0, 0, z)))
    assert ok
    assert dist[0] == pytest.approx(expected, abs=1e-5)


# reference test_collision_primitives.py:978-1060 (cylinder r=1, half-height 1 along Z; sphere r=0.5)
@pytest.mark.parametrize("sp,expected", [([2.0, 0, 0], 0.5), ([1.5, 0, 0], 0.0), ([1.4, 0, 0], -0.1), ([1.3, 0, 0], -0.2),
                                         ([0, 0, 2.0], 0.5), ([0, 0, 1.5], 0.0), ([0, 0, 1.4], -0.1)])
def test_sphere_cylinder(oracle_lib, sp, expected):
    ok, dist, _, _ = oracle_lib.primitive_pair(GeoType.SPHERE, (0.5, 0, 0), _xf(sp), GeoType.CYLINDER, (1.0, 1.0, 0.0), I7)
    assert ok
    assert dist[0] == pytest.approx(expected, abs=1e-5)


# reference test_collision_primitives.py:1149-1165 (unit half-extent box at origin)
SPHERE_BOX = [([2.5, 0, 0], 0.5, 1.0), ([2.0, 0, 0], 0.5, 0.5), ([1.5, 0, 0], 0.5, 0.0), ([1.4, 0, 0], 0.5, -0.1),
              ([1.3, 0, 0], 0.5, -0.2), ([1.2, 0, 0], 0.5, -0.3), ([0, 0, 2.0], 0.5, 0.5), ([0, 0, 1.5], 0.5, 0.0),
              ([0, 0, 1.3], 0.5, -0.2), ([0, 0, 0.4], 0.3, -0.9)]


@pytest.mark.parametrize("sp,r,expected", SPHERE_BOX)
def test_sphere_box(oracle_lib, sp, r, expected):
    ok, dist, _, _ = oracle_lib.primitive_pair(GeoType.SPHERE, (r, 0, 0), _xf(sp), GeoType.BOX, (1.0, 1.0, 1.0), I7)
    assert ok
    assert dist[0] == pytest.approx(expected, abs=1e-5)


# reference test_collision_primitives.py:1265-1274 (horizontal capsule r=0.5, half-length 1)
@pytest.mark.parametrize("z,expected", [(2.0, 1.5), (1.5, 1.0), (1.0, 0.5), (0.5, 0.0), (0.4, -0.1), (0.3, -0.2), (0.2, -0.3)])
def test_plane_capsule(oracle_lib, z, expected):
    ok, dist, pos, _ = oracle_lib.primitive_pair(GeoType.PLANE, (0, 0, 0), I7, GeoType.CAPSULE, (0.5, 1.0, 0),
                                                 _xf((0, 0, z), _axis_q((1.0, 0.0, 0.0))))
    assert ok
    assert dist[0] == pytest.approx(expected, abs=1e-5) and dist[1] == pytest.approx(expected, abs=1e-5)
    assert abs(abs(pos[0][0]) - 1.0) < 1e-5 and pos[0][0] * pos[1][0] < 0  # one contact under each end cap


# reference test_collision_primitives.py:1349-1360 (unit box): (center z, expected contact count, expected distance)
@pytest.mark.parametrize("z,count,expected", [(2.0, 0, 1.0), (1.0, 4, 0.0), (0.9, 4, -0.1), (0.8, 4, -0.2), (0.7, 4, -0.3), (0.5, 4, -0.5)])
def test_plane_box(oracle_lib, z, count, expected):
    ok, dist, pos, _ = oracle_lib.primitive_pair(GeoType.PLANE, (0, 0, 0), I7, GeoType.BOX, (1.0, 1.0, 1.0), _xf((0, 0, z)),
                                                 plane_box_margin=0.0)
    assert ok
    valid = dist < 1e9
    assert int(valid.sum()) == count
    for d in dist[valid]:
        assert d == pytest.approx(expected, abs=1e-5)


# plane-cylinder: upright cylinder (flat mode: deepest rim point + fixed tripod), reference collision_primitive.py:534-683,
# expectations in the style of test_collision_primitives.py:1444-1470 (lowest rim point = center_z - half_height)
@pytest.mark.parametrize("z", [2.0, 1.0, 0.9, 0.6])
def test_plane_cylinder_upright(oracle_lib, z):
    ok, dist, pos, normal = oracle_lib.primitive_pair(GeoType.PLANE, (0, 0, 0), I7, GeoType.CYLINDER, (0.5, 1.0, 0.0), _xf((0, 0, z)))
    assert ok
    valid = dist < 1e9
    assert int(valid.sum()) == 3  # one tripod point coincides with the "deepest" point and is merged
    for d in dist[valid]:
        assert d == pytest.approx(z - 1.0, abs=1e-5)
    r = np.linalg.norm(pos[valid][:, :2], axis=1)
    np.testing.assert_allclose(r, 0.5, atol=1e-5)  # contacts sit on the cap rim


def test_plane_cylinder_rolling(oracle_lib):
    """Cylinder lying on its side: deepest generator end + opposite end + one cap-rim point (rolling mode)."""
    ok, dist, pos, _ = oracle_lib.primitive_pair(GeoType.PLANE, (0, 0, 0), I7, GeoType.CYLINDER, (0.5, 1.0, 0.0),
                                                 _xf((0, 0, 0.45), _axis_q((1.0, 0.0, 0.0))))
    assert ok
    valid = dist < 1e9
    assert int(valid.sum()) >= 2
    assert dist[0] == pytest.approx(-0.05, abs=1e-5) and dist[1] == pytest.approx(-0.05, abs=1e-5)
    assert abs(pos[0][0] - pos[1][0]) == pytest.approx(2.0, abs=1e-5)  # the two ends of the contact generator


def test_example_basic_urdf_final_state(oracle_lib):
    """The reference example's own end-state assertion (example_basic_urdf.py:145-161; 200 frames x 10 substeps,
    SolverXPBD defaults): every link slower than 0.15, every root within 0.01 of z = 0.46."""
    worlds = 4
    model = scenes.quadruped_model(worlds, seed=None)
    pipe = oracle_lib.CollisionPipeline(model)
    solver = oracle_lib.SolverXPBD(model)
    s0, s1, ctrl, contacts = model.state(), model.state(), model.control(), pipe.contacts()
    dt = 1.0 / 100 / 10
    for _ in range(200 * 10):
        s0.clear_forces()
        pipe.collide(s0, contacts)
        solver.step(s0, s1, ctrl, contacts, dt)
        s0, s1 = s1, s0
    q, qd = s0.body_q.numpy(), s0.body_qd.numpy()
    assert np.abs(qd).max() < 0.15
    per = model.body_count // worlds
    for w in range(worlds):
        assert abs(q[w * per, 2] - 0.46) < 0.01


def test_free_fall_matches_semi_implicit_closed_form(oracle_lib):
    """A free body under gravity: v_n = g n dt, z_n = z_0 + g dt^2 n(n+1)/2 (semi-implicit Euler, solver.py:64-107)."""
    b = ModelBuilder()
    body = b.add_body(xform=X.transform((0.0, 0.0, 10.0)))
    b.add_shape_sphere(body, radius=0.1)
    model = b.finalize()
    solver = oracle_lib.SolverXPBD(model, iterations=1)
    s0, s1 = model.state(), model.state()
    dt, n = 1e-3, 200
    for _ in range(n):
        s0.clear_forces()
        solver.step(s0, s1, None, None, dt)
        s0, s1 = s1, s0
    g = -9.81
    assert s0.body_qd[0, 2].item() == pytest.approx(g * n * dt, rel=1e-5)
    assert s0.body_q[0, 2].item() == pytest.approx(10.0 + g * dt * dt * n * (n + 1) / 2, rel=1e-6)


def test_sphere_rests_on_ground(oracle_lib):
    """Sphere dropped on the plane settles at z = radius (reference tests/test_rigid_contact.py shapes-on-plane)."""
    b = ModelBuilder()
    body = b.add_body(xform=X.transform((0.0, 0.0, 0.6)))
    b.add_shape_sphere(body, radius=0.5)
    b.add_ground_plane()
    model = b.finalize()
    pipe = oracle_lib.CollisionPipeline(model)
    solver = oracle_lib.SolverXPBD(model, iterations=4)
    s0, s1, contacts = model.state(), model.state(), pipe.contacts()
    for _ in range(600):
        s0.clear_forces()
        pipe.collide(s0, contacts)
        solver.step(s0, s1, None, contacts, 1.0 / 600)
        s0, s1 = s1, s0
    assert s0.body_q[0, 2].item() == pytest.approx(0.5, abs=2e-3)
    assert abs(s0.body_qd[0, 2].item()) < 1e-2


def _pendulum(sphere_radius=0.01):
    """Single pendulum of reference tests/test_physics_verification.py:112-139 (g = -10 along Y, L = 1, theta0 = 0.05)."""
    import newton_b200

    b = ModelBuilder(gravity=(0.0, -10.0, 0.0), up_axis="y")
    link = b.add_link()
    b.add_shape_sphere(link, radius=sphere_radius)
    j = b.add_joint_revolute(parent=-1, child=link, axis=(0.0, 0.0, 1.0), parent_xform=X.transform(),
                             child_xform=X.transform((0.0, 1.0, 0.0)), armature=0.0)
    b.add_articulation([j])
    model = b.finalize()
    model.joint_q[0] = 0.05
    host_fk(model, model.joint_q, model.joint_qd, model)
    return model


@pytest.mark.parametrize("solver_name,sim_dt,gen", [("featherstone", 1e-3, True), ("xpbd", 3e-4, False)])
def test_pendulum_period(oracle_lib, solver_name, sim_dt, gen):
    """theta(t) = theta0 cos(2 pi t / T), T = 2 pi sqrt(I_pivot / (m g d)); mean trajectory error < 1 % of the amplitude
    (reference tests/test_physics_verification.py:105-185, same dt per solver as :1574-1582)."""
    model = _pendulum()
    solver = oracle_lib.SolverFeatherstone(model, angular_damping=0.0) if solver_name == "featherstone" else \
        oracle_lib.SolverXPBD(model, iterations=20, angular_damping=0.0)  # reference :1552
    mass = model.body_mass[0].item()
    I_pivot = model.body_inertia[0, 2, 2].item() + mass * 1.0
    T = 2.0 * np.pi * np.sqrt(I_pivot / (mass * 10.0 * 1.0))
    n = int(3.5 * T / sim_dt)
    s0, s1 = model.state(), model.state()
    angles = np.empty(n)
    for i in range(n):
        s0.clear_forces()
        solver.step(s0, s1, None, None, sim_dt)
        s0, s1 = s1, s0
        if gen:
            angles[i] = s0.joint_q[0].item()
        else:
            bq = s0.body_q[0]
            angles[i] = math.atan2(bq[0].item(), -bq[1].item())
    t = np.arange(1, n + 1) * sim_dt
    err = np.mean(np.abs(angles - 0.05 * np.cos(2.0 * np.pi / T * t))) / 0.05
    assert err < 0.01, err


def test_featherstone_double_pendulum_energy(oracle_lib):
    """Config 1 (example_basic_pendulum scene, SolverFeatherstone, no damping): total energy drift < 1 % over 1 s."""
    model = scenes.pendulum_model()
    solver = oracle_lib.SolverFeatherstone(model, angular_damping=0.0)
    s0, s1 = model.state(), model.state()
    m = model.body_mass.numpy()
    I = model.body_inertia.numpy()

    def energy(s):
        q, qd = s.body_q.numpy(), s.body_qd.numpy()
        e = 0.0
        for b in range(2):
            R = X.quat_to_matrix(q[b, 3:].astype(np.float64))
            Iw = R @ I[b] @ R.T
            com = X.transform_point(q[b].astype(np.float64), model.body_com[b].numpy().astype(np.float64))
            e += 0.5 * m[b] * qd[b, :3] @ qd[b, :3] + 0.5 * qd[b, 3:] @ Iw @ qd[b, 3:] + m[b] * 9.81 * com[2]
        return e

    solver.step(s0, s1, None, None, 1e-3)
    s0, s1 = s1, s0
    e0 = energy(s0)
    for _ in range(1000):
        s0.clear_forces()
        solver.step(s0, s1, None, None, 1e-3)
        s0, s1 = s1, s0
    e1 = energy(s0)
    assert abs(e1 - e0) / abs(e0) < 0.01
    q = s0.body_q.numpy()
    assert np.all(np.abs(q[:, 0]) < 1e-5) and np.all(q[:, 2] < 5.0 + 1e-4)  # stays in its plane, below the pivot


# ---- generic convex path: MPR / GJK cores + manifold ------------------------------------------------------------
# (the oracle compiles newton_b200/csrc/nb2_convex.cuh for the host; these reference known answers pin that the
#  algorithm written there is the reference's - see oracle/oracle_gjk.h)

def test_mpr_box_support_tie_boundary(oracle_lib):
    """newton/tests/test_mpr.py:173-190: witnesses stay valid just below / above the box support tie threshold."""
    tie_eps = 1.0e-6  # _CENTERED_BOX_SUPPORT_TIE_EPSILON (support_function.py:41)
    angles = tie_eps * np.array([0.5, 2.0], dtype=np.float32)
    for a in angles:
        q = [math.sin(0.5 * a), 0.0, 0.0, math.cos(0.5 * a)]
        hit, pa, pb, n, pen = oracle_lib.mpr_core(GeoType.BOX, [0.5] * 3, GeoType.BOX, [0.5] * 3, [0.0, 0.999, 0.0], q)
        assert hit == 1
        assert np.linalg.norm(n) == pytest.approx(1.0, abs=1e-6)
        np.testing.assert_allclose(n, [0.0, 1.0, 0.0], atol=1e-5)
        assert pen == pytest.approx(0.5 + 0.5 * (math.cos(a) + math.sin(a)) - 0.999, abs=2e-5)
        assert float(np.dot(pb - pa, n)) == pytest.approx(-pen, abs=1e-6)


def test_gjk_cylinder_barrel_support(oracle_lib):
    """newton/tests/test_gjk.py:128-158."""
    r, hh, br = 0.5, 1.0, 2.0
    eq = r + br - math.sqrt(br**2 - hh**2)
    np.testing.assert_allclose(oracle_lib.support_map(GeoType.CYLINDER, (r, hh, 0.0), (1, 0, 0)), (r, 0, 0), atol=1e-6)
    np.testing.assert_allclose(oracle_lib.support_map(GeoType.CYLINDER, (r, hh, br), (1, 0, 0)), (eq, 0, 0), atol=1e-6)
    np.testing.assert_allclose(oracle_lib.support_map(GeoType.CYLINDER, (r, hh, br), (0, 0, 1)), (r, 0, hh), atol=1e-6)
    d = np.array([1.0, 0.0, 0.25])
    sz = br * d[2] / np.linalg.norm(d)
    sr = r - math.sqrt(br**2 - hh**2) + math.sqrt(br**2 - sz**2)
    np.testing.assert_allclose(oracle_lib.support_map(GeoType.CYLINDER, (r, hh, br), d), (sr, 0.0, sz), atol=1e-6)


def test_gjk_distance_cases(oracle_lib):
    """newton/tests/test_gjk.py:160-235 (COLLIDE_EPSILON 1e-6 as in the reference's kernel; B's pose relative to A)."""
    q = [0, 0, 0, 1]
    sep, _, _, _, dist = oracle_lib.gjk_core(GeoType.SPHERE, (1, 0, 0), GeoType.SPHERE, (1, 0, 0), [3, 0, 0], q, eps=1e-6)
    assert sep == 1 and dist == pytest.approx(1.0, abs=1e-5)
    sep, _, _, _, dist = oracle_lib.gjk_core(GeoType.SPHERE, (1, 0, 0), GeoType.SPHERE, (1, 0, 0), [2, 0, 0], q, eps=1e-6)
    assert dist == pytest.approx(0.0, abs=1e-5)
    sep, _, _, _, dist = oracle_lib.gjk_core(GeoType.SPHERE, (3, 0, 0), GeoType.SPHERE, (3, 0, 0), [4, 0, 0], q, eps=1e-6)
    assert sep == 0 and dist == pytest.approx(0.0, abs=1e-5)  # overlapping -> "collision"
    sep, _, _, n, dist = oracle_lib.gjk_core(GeoType.BOX, (1, 1, 1), GeoType.BOX, (1, 1, 1), [4.5, 0, 0], q, eps=1e-6)
    assert sep == 1 and dist == pytest.approx(2.5, abs=1e-5)
    np.testing.assert_allclose(n, [1, 0, 0], atol=1e-5)


def _dist_to_box(p, pos, half):
    d = np.abs(np.asarray(p) - np.asarray(pos)) - np.asarray(half)
    return float(np.linalg.norm(np.maximum(d, 0.0)) + min(max(d[0], d[1], d[2]), 0.0))


def test_narrow_phase_box_box_face(oracle_lib):
    """newton/tests/test_narrow_phase.py:707-778: overlap 0.2 along X, normal A->B, midpoint surface reconstruction."""
    cnt, dist, pos, n = oracle_lib.convex_pair(GeoType.BOX, (1, 1, 1), _xf([0, 0, 0]), GeoType.BOX, (1, 1, 1), _xf([1.8, 0, 0]))
    assert cnt == 4  # full face manifold
    for i in range(cnt):
        assert np.linalg.norm(n[i]) == pytest.approx(1.0, abs=1e-5)
        assert n[i][0] > 0.9
        assert dist[i] == pytest.approx(-0.2, abs=1e-4)  # reference bar: places=1
        # check_surface_reconstruction: centre +- n*d/2 lie on the two surfaces
        assert abs(_dist_to_box(pos[i] - n[i] * dist[i] * 0.5, [0, 0, 0], [1, 1, 1])) < 1e-4
        assert abs(_dist_to_box(pos[i] + n[i] * dist[i] * 0.5, [1.8, 0, 0], [1, 1, 1])) < 1e-4


def test_narrow_phase_box_box_edge(oracle_lib):
    """newton/tests/test_narrow_phase.py:780-799: box rotated 45 deg about Z at x = 1.2 -> edge contact, unit normal."""
    a = math.pi / 4.0
    cnt, dist, pos, n = oracle_lib.convex_pair(GeoType.BOX, (0.5,) * 3, _xf([0, 0, 0]), GeoType.BOX, (0.5,) * 3,
                                               _xf([1.2, 0, 0], (0.0, 0.0, math.sin(a / 2), math.cos(a / 2))))
    assert cnt > 0
    assert np.linalg.norm(n[0]) == pytest.approx(1.0, abs=1e-5)
    np.testing.assert_allclose(n[0], [1, 0, 0], atol=1e-4)
    assert dist[0] == pytest.approx(1.2 - 0.5 - 0.5 * math.sqrt(2.0), abs=1e-4)  # edge x = 1.2 - 0.7071 vs face x = 0.5


def test_narrow_phase_ellipsoids(oracle_lib):
    """newton/tests/test_narrow_phase.py:1677-1950 (ellipsoid pairs always take the GJK/MPR route)."""
    E = GeoType.ELLIPSOID
    # separated: no contact or positive distance (:1677-1704)
    cnt, dist, _, n = oracle_lib.convex_pair(E, (1.0, 0.5, 0.3), _xf([0, 0, 0]), E, (1.0, 0.5, 0.3), _xf([3.0, 0, 0]), gap_sum=0.0)
    assert cnt == 0 or dist[0] > 0.0
    # penetrating 0.2 along X (:1706-1741)
    cnt, dist, _, n = oracle_lib.convex_pair(E, (1.0, 0.5, 0.3), _xf([0, 0, 0]), E, (1.0, 0.5, 0.3), _xf([1.8, 0, 0]))
    assert cnt == 1 and dist[0] == pytest.approx(-0.2, abs=1e-3)
    assert np.linalg.norm(n[0]) == pytest.approx(1.0, abs=1e-5) and n[0][0] > 0.99
    # sphere-like ellipsoids behave like spheres (:1913-1950)
    cnt, dist, _, n = oracle_lib.convex_pair(E, (1, 1, 1), _xf([0, 0, 0]), E, (1, 1, 1), _xf([1.8, 0, 0]))
    assert cnt == 1 and dist[0] == pytest.approx(-0.2, abs=1e-3) and n[0][0] == pytest.approx(1.0, abs=1e-3)
    # ellipsoid vs sphere: type-sorted so the sphere is shape A (:1743-1787), normal from the sphere to the ellipsoid
    cnt, dist, _, n = oracle_lib.convex_pair(GeoType.SPHERE, (0.5, 0.5, 0.5), _xf([1.4, 0, 0]), E, (1.0, 0.5, 0.3), _xf([0, 0, 0]))
    assert cnt == 1 and dist[0] == pytest.approx(-0.1, abs=1e-3) and n[0][0] < -0.99
    # ellipsoid vs box / capsule produce unit normals (:1789-1880)
    cnt, dist, _, n = oracle_lib.convex_pair(E, (1.0, 0.5, 0.3), _xf([0, 0, 0]), GeoType.BOX, (0.5, 0.5, 0.5), _xf([1.3, 0, 0]))
    assert cnt >= 1 and dist[0] == pytest.approx(-0.2, abs=1e-3) and np.linalg.norm(n[0]) == pytest.approx(1.0, abs=1e-5)
    cnt, dist, _, n = oracle_lib.convex_pair(GeoType.CAPSULE, (0.5, 1.0, 0.0), _xf([1.3, 0, 0]), E, (1.0, 0.5, 0.3), _xf([0, 0, 0]))
    assert cnt >= 1 and np.linalg.norm(n[0]) == pytest.approx(1.0, abs=1e-5)


def test_xpbd_aligned_box_stack_remains_stable(oracle_lib):
    """newton/tests/test_solver_xpbd.py:1791-1840: an aligned five-box stack (up axis Y) stays upright for 180 frames;
    MPR must keep four-point face manifolds as solver-scale quaternion drift crosses zero."""
    b = ModelBuilder(up_axis="Y")
    b.add_ground_plane()
    for i in range(5):
        body = b.add_body(xform=X.transform((0.0, 0.5 + i, 0.0)))
        b.add_shape_box(body, hx=0.5, hy=0.5, hz=0.5)
    model = b.finalize()
    solver = oracle_lib.SolverXPBD(model, iterations=4)
    s0, s1, ctl = model.state(), model.state(), model.control()
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    dt = 1.0 / 60.0 / 4
    for _ in range(180 * 4):
        s0.clear_forces()
        pipe.collide(s0, contacts)
        solver.step(s0, s1, ctl, contacts, dt)
        s0, s1 = s1, s0
    bq = s0.body_q.numpy()
    assert np.all(np.isfinite(bq))
    assert int(contacts.rigid_contact_count[0]) == 20  # 4-point manifolds on all five interfaces
    np.testing.assert_allclose(bq[:, 1], 0.5 + np.arange(5), atol=2.0e-2)
    assert float(np.max(np.linalg.norm(bq[:, (0, 2)], axis=1))) < 1.0e-2
    assert float(np.max(np.linalg.norm(bq[:, 3:][:, (0, 2)], axis=1))) < 1.0e-3


# ---- XPBD row a17: restitution, contact forces, joint reaction (body_parent_f) --------------------------------------------------

@pytest.mark.parametrize("e", [0.5, 0.8])
def test_xpbd_restitution_rebound_height(oracle_lib, e):
    """newton/tests/test_physics_verification.py:612-676 (test_restitution, SolverXPBD(iterations=10,
    enable_restitution=True)): a sphere dropped from 1 m rebounds to e^2 * h within 1 %."""
    from newton_b200.sim.builder import ShapeConfig

    g, h_drop, radius = -10.0, 1.0, 0.05
    cfg = ShapeConfig(mu=0.0, restitution=e, ke=1e4, kd=100.0, kf=0.0, margin=0.001, gap=0.0)
    b = ModelBuilder(up_axis="Y", gravity=g)
    b.add_ground_plane(cfg=cfg)
    body = b.add_body(xform=X.transform((0.0, radius + h_drop, 0.0)))
    b.add_shape_sphere(body, radius=radius, cfg=cfg)
    model = b.finalize()
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    solver = oracle_lib.SolverXPBD(model, iterations=10, angular_damping=0.0, enable_restitution=True)
    s0, s1 = model.state(), model.state()
    sim_dt = 1e-3
    ys = []
    for _ in range(int(3.0 * math.sqrt(2.0 * h_drop / abs(g)) / sim_dt)):
        s0.clear_forces()
        pipe.collide(s0, contacts)
        solver.step(s0, s1, None, contacts, sim_dt)
        s0, s1 = s1, s0
        ys.append(float(s0.body_q[0, 1]))
    y = np.array(ys)
    assert y.min() > -0.01
    impact = next(i for i in range(1, len(y) - 1) if y[i] < y[i - 1] and y[i] <= y[i + 1])
    h_rebound = y[impact:].max() - radius
    assert h_rebound == pytest.approx(e * e * h_drop, rel=0.01)


def test_xpbd_contact_force_static_equilibrium(oracle_lib):
    """newton/tests/test_solver_xpbd.py:848-1035: time-averaged ``contacts.force`` of resting bodies equals their weight
    (sphere, heavy sphere, 4-contact box with 1/N weighting, two-cube base of a 3-cube pyramid carrying 1.5 mg each)."""
    from newton_b200.sim.builder import ShapeConfig

    grav = 9.81
    b = ModelBuilder()
    b.add_ground_plane()
    masses = {}

    def body(pos, density, **shape):
        b.default_shape_cfg.density = density
        i = b.add_body(xform=X.transform(pos))
        if "radius" in shape:
            b.add_shape_sphere(i, radius=shape["radius"])
            masses[i] = density * 4.0 / 3.0 * math.pi * shape["radius"] ** 3
        else:
            b.add_shape_box(i, hx=shape["h"], hy=shape["h"], hz=shape["h"])
            masses[i] = density * (2.0 * shape["h"]) ** 3
        return i

    sphere = body((0.0, 0.0, 0.25), 1000.0, radius=0.25)
    heavy = body((10.0, 0.0, 0.5), 2000.0, radius=0.5)
    box = body((20.0, 0.0, 0.5), 1000.0, h=0.5)
    left = body((29.5, 0.0, 0.5), 1000.0, h=0.5)
    right = body((30.5, 0.0, 0.5), 1000.0, h=0.5)
    body((30.0, 0.0, 1.5), 1000.0, h=0.5)
    model = b.finalize()
    model.request_contact_attributes("force")
    solver = oracle_lib.SolverXPBD(model, iterations=32, rigid_contact_con_weighting=True)
    s0, s1, ctl = model.state(), model.state(), model.control()
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    assert contacts.force is not None
    sub_dt = 1.0 / 60.0 / 8
    acc = {i: np.zeros(3) for i in (sphere, heavy, box, left, right)}
    shape_body = model.shape_body.numpy()
    avg_steps = 60
    for frame in range(200 + avg_steps):
        for _ in range(8):
            s0.clear_forces()
            pipe.collide(s0, contacts)
            solver.step(s0, s1, ctl, contacts, sub_dt)
            s0, s1 = s1, s0
        if frame < 200:
            continue
        solver.update_contacts(contacts, s0)
        nc = int(contacts.rigid_contact_count[0])
        f = contacts.force.numpy()[:nc, :3]
        sh0, sh1 = contacts.rigid_contact_shape0.numpy()[:nc], contacts.rigid_contact_shape1.numpy()[:nc]
        for ci in range(nc):  # contacts.force = force on body0 by body1; ground is shape 0
            if sh0[ci] == 0:
                other, fo = sh1[ci], f[ci]
            elif sh1[ci] == 0:
                other, fo = sh0[ci], -f[ci]
            else:
                continue
            ob = shape_body[other]
            if ob in acc:
                acc[ob] += fo
    for i in acc:
        acc[i] /= avg_steps
    assert acc[sphere][2] == pytest.approx(-masses[sphere] * grav, rel=0.05)
    assert acc[heavy][2] == pytest.approx(-masses[heavy] * grav, rel=0.05)
    assert acc[box][2] == pytest.approx(-masses[box] * grav, rel=0.10)  # N contacts must not report N * mg
    for i in (sphere, heavy):
        assert abs(acc[i][0]) < 0.5 and abs(acc[i][1]) < 0.5
    assert -acc[left][2] == pytest.approx(1.5 * masses[left] * grav, rel=0.15)
    assert -acc[right][2] == pytest.approx(1.5 * masses[right] * grav, rel=0.15)


@pytest.mark.parametrize("joint_kind", ["revolute", "ball", "fixed"])
@pytest.mark.parametrize("parent_kinematic", [False, True])
def test_xpbd_parent_force_single_body(oracle_lib, joint_kind, parent_kinematic):
    """newton/tests/test_solver_xpbd.py:1120-1280: a body hanging 1 m below a joint to the world / a kinematic body:
    the time-averaged ``body_parent_f`` is its weight along +Z within 1 %, lateral force and torque ~ 0."""
    from newton_b200 import BodyFlags

    b = ModelBuilder(gravity=-9.81)
    if parent_kinematic:
        parent = b.add_body(xform=X.transform((0.0, 0.0, 0.0)))
        b.add_shape_box(parent, hx=0.05, hy=0.05, hz=0.05)
        b.body_flags[parent] = int(BodyFlags.KINEMATIC)
    else:
        parent = -1
    child = b.add_link()
    b.add_shape_box(child, hx=0.1, hy=0.1, hz=0.1)
    kw = dict(parent_xform=X.transform((0.0, 0.0, 0.0)), child_xform=X.transform((0.0, 0.0, 1.0)))
    if joint_kind == "revolute":
        j = b.add_joint_revolute(parent, child, axis=(0.0, 1.0, 0.0), **kw)
    elif joint_kind == "ball":
        j = b.add_joint_ball(parent, child, **kw)
    else:
        j = b.add_joint_fixed(parent, child, **kw)
    b.add_articulation([j])
    model = b.finalize()
    model.request_state_attributes("body_parent_f")
    import newton_b200

    host_fk(model, model.joint_q, model.joint_qd, model)
    solver = oracle_lib.SolverXPBD(model, iterations=8)
    s0, s1 = model.state(), model.state()
    assert s0.body_parent_f is not None
    sub_dt = 1.0 / 60.0 / 8
    avg = np.zeros(6)
    for frame in range(60 + 30):
        for _ in range(8):
            solver.step(s0, s1, None, None, sub_dt)
            s0, s1 = s1, s0
        if frame >= 60:
            avg += s0.body_parent_f.numpy()[child]
    avg /= 30
    weight = float(model.body_mass[child]) * 9.81
    assert avg[2] == pytest.approx(weight, rel=0.01)
    np.testing.assert_allclose(avg[:2], 0.0, atol=0.1)
    np.testing.assert_allclose(avg[3:], 0.0, atol=0.1)


@pytest.mark.parametrize("solver_name", ["featherstone", "xpbd"])
def test_example_basic_pendulum_final_state(oracle_lib, solver_name):
    """BASELINE config 1: ``example_basic_pendulum.py`` (2 box links, 2 revolute joints about Y, fps 100, 10 substeps,
    100 frames) and its ``test_final`` (:113-137): links stay in the x = 0 swing plane, |y| < 1, 0 < z < 5, bounded
    velocities.  Run with SolverFeatherstone (the config) and with the example's own SolverXPBD."""
    model = scenes.pendulum_model()
    solver = oracle_lib.SolverFeatherstone(model) if solver_name == "featherstone" else oracle_lib.SolverXPBD(model)
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    s0, s1, ctl = model.state(), model.state(), model.control()
    for _ in range(100 * 10):
        s0.clear_forces()
        pipe.collide(s0, contacts)
        solver.step(s0, s1, ctl, contacts, 1.0 / 100 / 10)
        s0, s1 = s1, s0
    q, qd = s0.body_q.numpy(), s0.body_qd.numpy()
    for b in (0, 1):
        assert abs(q[b, 0]) < 1e-5 and abs(q[b, 1]) < 1.0 and 0.0 < q[b, 2] < 5.0
        assert abs(qd[b, 0]) < 1e-4
        assert abs(qd[b, 1]) < 10.0 and abs(qd[b, 2]) < 5.0 and abs(qd[b, 3]) < 10.0 and abs(qd[b, 4]) < 10.0
    assert np.abs(qd).max() > 0.1  # it is swinging


# ---- more closed-form checks of newton/tests/test_physics_verification.py (both solvers where the reference runs both) ------------

def _solver(oracle_lib, name, model):
    if name == "featherstone":
        return oracle_lib.SolverFeatherstone(model, angular_damping=0.0)
    return oracle_lib.SolverXPBD(model, iterations=20, angular_damping=0.0)


@pytest.mark.parametrize("solver_name", ["featherstone", "xpbd"])
def test_projectile_motion(oracle_lib, solver_name):
    """test_physics_verification.py:289-353: ballistic flight, position / velocity against the closed form at six instants."""
    import newton_b200

    g, p0, v0 = -10.0, np.array([0.0, 10.0, 0.0]), np.array([5.0, 10.0, 0.0])
    b = ModelBuilder(up_axis="Y", gravity=g)
    body = b.add_body(xform=X.transform(p0))
    b.add_shape_sphere(body, radius=0.1)
    model = b.finalize()
    s0, s1 = model.state(), model.state()
    if solver_name == "featherstone":
        s0.joint_qd[:3] = torch_f32(v0)
        host_fk(model, s0.joint_q, s0.joint_qd, s0)
    else:
        s0.body_qd[0, :3] = torch_f32(v0)
    solver = _solver(oracle_lib, solver_name, model)
    dt = 1e-3
    for i in range(1, 101):
        s0.clear_forces()
        solver.step(s0, s1, None, None, dt)
        s0, s1 = s1, s0
        if i in (10, 20, 30, 50, 70, 100):
            t = i * dt
            pos, vel = s0.body_q.numpy()[0, :3], s0.body_qd.numpy()[0, :3]
            exp_p = p0 + v0 * t + np.array([0.0, 0.5 * g * t * t, 0.0])
            exp_v = v0 + np.array([0.0, g * t, 0.0])
            np.testing.assert_allclose(pos, exp_p, atol=max(2.0 * 0.5 * abs(g) * dt * t, 1e-3))
            np.testing.assert_allclose(vel, exp_v, atol=max(abs(g) * dt, 1e-3))


def torch_f32(a):
    import torch

    return torch.tensor(np.asarray(a), dtype=torch.float32)


def test_joint_actuation_featherstone(oracle_lib):
    """test_physics_verification.py:356-447: constant joint_f on a revolute / prismatic joint without gravity:
    omega = tau t / I_zz and v = F t / m within one step's increment."""
    import newton_b200
    from newton_b200.sim.builder import JointDofConfig  # noqa: F401

    b = ModelBuilder(up_axis="Y", gravity=0.0)
    link_rev = b.add_link()
    b.add_shape_box(link_rev, hx=0.2, hy=0.2, hz=0.2)
    j_rev = b.add_joint_revolute(parent=-1, child=link_rev, axis=(0.0, 0.0, 1.0), armature=0.0)
    b.add_articulation([j_rev])
    link_pri = b.add_link()
    b.add_shape_sphere(link_pri, radius=0.1)
    j_pri = b.add_joint_prismatic(parent=-1, child=link_pri, axis=(1.0, 0.0, 0.0), parent_xform=X.transform((0.0, 5.0, 0.0)), armature=0.0)
    b.add_articulation([j_pri])
    model = b.finalize()
    host_fk(model, model.joint_q, model.joint_qd, model)
    I_zz = float(model.body_inertia[0, 2, 2])
    mass = float(model.body_mass[1])
    solver = oracle_lib.SolverFeatherstone(model, angular_damping=0.0)
    s0, s1, ctl = model.state(), model.state(), model.control()
    qd_start = model.joint_qd_start.numpy()
    ctl.joint_f[qd_start[0]] = 5.0
    ctl.joint_f[qd_start[1]] = 5.0
    dt, n = 1e-3, 300
    for _ in range(n):
        s0.clear_forces()
        solver.step(s0, s1, ctl, None, dt)
        s0, s1 = s1, s0
    t = n * dt
    assert float(s0.joint_qd[qd_start[0]]) == pytest.approx(5.0 * t / I_zz, abs=max(5.0 / I_zz * dt, 1e-3))
    assert float(s0.joint_qd[qd_start[1]]) == pytest.approx(5.0 * t / mass, abs=max(5.0 / mass * dt, 1e-3))


@pytest.mark.parametrize("solver_name", ["featherstone", "xpbd"])
def test_momentum_conservation(oracle_lib, solver_name):
    """test_physics_verification.py:450-537: four free boxes without gravity keep linear and angular momentum (drift < 5e-4)."""
    import newton_b200

    positions = [(0.0, 0.0, 0.0), (100.0, 0.0, 0.0), (0.0, 100.0, 0.0), (0.0, 0.0, 100.0)]
    velocities = np.array([(1.0, 0.0, 0.0, 0.0, 0.0, 0.5), (0.0, -1.0, 0.0, 0.3, 0.0, 0.0), (0.0, 0.0, 1.5, 0.0, -0.2, 0.0),
                           (-0.5, 0.5, -0.5, 0.0, 0.0, -0.3)], dtype=np.float32)
    b = ModelBuilder(up_axis="Y", gravity=0.0)
    for p in positions:
        body = b.add_body(xform=X.transform(p))
        b.add_shape_box(body, hx=0.5, hy=0.5, hz=0.5)
    model = b.finalize()
    s0, s1 = model.state(), model.state()
    if solver_name == "featherstone":
        s0.joint_qd.copy_(torch_f32(velocities.reshape(-1)))
        host_fk(model, s0.joint_q, s0.joint_qd, s0)
    else:
        s0.body_qd.copy_(torch_f32(velocities))
    masses, inertias = model.body_mass.numpy(), model.body_inertia.numpy()

    def momenta(state):
        q, qd = state.body_q.numpy().astype(np.float64), state.body_qd.numpy().astype(np.float64)
        p, L = np.zeros(3), np.zeros(3)
        for i in range(4):
            R = X.quat_to_matrix(q[i, 3:])
            p += masses[i] * qd[i, :3]
            L += np.cross(q[i, :3], masses[i] * qd[i, :3]) + (R @ inertias[i] @ R.T) @ qd[i, 3:]
        return p, L

    p0, L0 = momenta(s0)
    assert np.linalg.norm(p0) > 0.1 and np.linalg.norm(L0) > 0.1
    solver = _solver(oracle_lib, solver_name, model)
    for _ in range(1000):
        s0.clear_forces()
        solver.step(s0, s1, None, None, 1e-3)
        s0, s1 = s1, s0
    p1, L1 = momenta(s0)
    assert np.linalg.norm(p1 - p0) / np.linalg.norm(p0) < 5e-4
    assert np.linalg.norm(L1 - L0) / np.linalg.norm(L0) < 5e-4
    assert np.linalg.norm(s0.body_q.numpy()[:, :3] - np.array(positions)) > 0.1


def test_torque_free_precession_featherstone(oracle_lib):
    """test_physics_verification.py:540-603: anisotropic body on a 3-angular-axis D6 joint, no torque: world angular
    momentum drifts < 5e-3 over 20 steps of 10 ms while the body precesses."""
    import newton_b200
    from newton_b200.sim.builder import JointDofConfig

    b = ModelBuilder(gravity=0.0)
    link = b.add_link(mass=1.0, com=(0.0, 0.0, 0.0), inertia=np.diag([0.2, 0.5, 0.4]))
    j = b.add_joint_d6(parent=-1, child=link, angular_axes=[JointDofConfig.create_unlimited(a) for a in "xyz"])
    b.add_articulation([j])
    model = b.finalize()
    s0, s1 = model.state(), model.state()
    s0.joint_qd[:3] = torch_f32([0.7, -0.5, 0.9])
    host_fk(model, s0.joint_q, s0.joint_qd, s0)
    I_body = model.body_inertia.numpy()[0].astype(np.float64)

    def L_world(state):
        q = state.body_q.numpy()[0].astype(np.float64)
        R = X.quat_to_matrix(q[3:])
        return (R @ I_body @ R.T) @ state.body_qd.numpy()[0, 3:].astype(np.float64)

    L0, quat0 = L_world(s0), s0.body_q.numpy()[0, 3:].copy()
    assert np.linalg.norm(L0) > 0.1
    solver = oracle_lib.SolverFeatherstone(model, angular_damping=0.0)
    for _ in range(20):
        s0.clear_forces()
        solver.step(s0, s1, None, None, 1e-2)
        s0, s1 = s1, s0
    assert np.linalg.norm(L_world(s0) - L0) / np.linalg.norm(L0) < 5e-3
    assert 2.0 * math.acos(min(abs(float(np.dot(quat0, s0.body_q.numpy()[0, 3:]))), 1.0)) > 0.1


def _ramp_scene():
    """Scene of newton/tests/test_rigid_contact.py:236-432 (objects resting on a 30-degree ramp against an end wall),
    including the two cubes written as convex hulls of their 8 corners (:402-425), furthest up the ramp."""
    L, TH, ANG, WALL_H, CUBE = 10.0, 0.5, math.radians(30.0), 2.0, 0.99
    W = CUBE * 2.01
    b = ModelBuilder()
    b.default_shape_cfg.ke, b.default_shape_cfg.kd, b.default_shape_cfg.kf = 2e4, 500.0, 0.5
    centre = np.array([0.0, L / 2 * math.cos(ANG), L / 2 * math.sin(ANG)])
    rq = X.quat_from_axis_angle((1.0, 0.0, 0.0), ANG)
    b.add_shape_plane(body=-1, xform=X.transform(centre, rq), width=0.0, length=0.0)
    fwd, up, right = (X.quat_rotate(rq, np.array(v)) for v in ((0.0, -1.0, 0.0), (0.0, 0.0, 1.0), (1.0, 0.0, 0.0)))
    gh, gt = 0.3, 0.1
    for sgn in (1.0, -1.0):
        c = centre + sgn * (W / 2 + gt / 2) * right + (gh / 2) * up
        b.add_shape_box(-1, xform=X.transform(c, rq), hx=gt / 2, hy=L / 2, hz=gh / 2)
    shift = 0.6 * L
    tmp = centre + 0.5 * CUBE * (up + shift * fwd)
    b.add_shape_box(-1, xform=X.transform((0.0, tmp[1] - CUBE / 2 * 1.4 - TH / 2, tmp[2])), hx=W / 2, hy=TH / 2, hz=WALL_H / 2)

    def at(side, k):
        return centre + 0.5 * CUBE * (up + side * right + (shift - k) * fwd)

    for side in (1.0, -1.0):
        body = b.add_body(xform=X.transform(at(side, 0.0), rq))
        b.add_shape_box(body, hx=CUBE / 2, hy=CUBE / 2, hz=CUBE / 2)
    for side in (1.0, -1.0):
        body = b.add_body(xform=X.transform(at(side, 2.01), rq))
        b.add_shape_sphere(body, radius=CUBE / 2)
    z_to_x = X.quat_between_vectors((0.0, 0.0, 1.0), (1.0, 0.0, 0.0))
    lying = X.quat_mul(rq, z_to_x)
    body = b.add_body(xform=X.transform(at(0.0, 4.02), lying))
    b.add_shape_capsule(body, radius=CUBE / 2, half_height=CUBE / 2)
    body = b.add_body(xform=X.transform(at(0.0, 6.03), lying))
    b.add_shape_cylinder(body, radius=CUBE / 2, half_height=CUBE)
    for side in (1.0, -1.0):
        body = b.add_body(xform=X.transform(at(side, 8.04), rq))
        b.add_shape_box(body, hx=CUBE / 2, hy=CUBE / 2, hz=CUBE / 2)
    for side in (1.0, -1.0):  # cones, axis along the ramp normal: plane-cone goes through the plane -> box proxy + MPR
        body = b.add_body(xform=X.transform(at(side, 10.05), rq))
        b.add_shape_cone(body, radius=CUBE / 2, half_height=CUBE / 2)
    for side in (1.0, -1.0):
        body = b.add_body(xform=X.transform(at(side, 12.06), rq))
        b.add_shape_box(body, hx=CUBE / 2, hy=CUBE / 2, hz=CUBE / 2)
    from newton_b200.geometry.mesh import Mesh

    cube_mesh = Mesh.create_box(CUBE / 2, CUBE / 2, CUBE / 2)
    for side in (1.0, -1.0):  # convex-hull cubes (test_rigid_contact.py:402-425)
        body = b.add_body(xform=X.transform(at(side, 14.07), rq))
        b.add_shape_convex_hull(body, mesh=cube_mesh, scale=(1.0, 1.0, 1.0))
    b.add_ground_plane()
    return b.finalize(), CUBE


def test_shapes_on_ramp_multicontact_stay_put(oracle_lib):
    """test_rigid_contact.py:236-510 (test_shape_collisions_gjk_mpr_multicontact): with correct MPR / GJK manifolds - cube
    vs static box walls, capsule / cylinder vs cube, the axial rolling post-process - nothing on the ramp moves more than
    0.15 cube sizes or turns more than 10 degrees in 100 frames x 10 substeps of SolverXPBD(iterations=2)."""
    model, cube = _ramp_scene()
    solver = oracle_lib.SolverXPBD(model, iterations=2)
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    s0, s1, ctl = model.state(), model.state(), model.control()
    q_init = s0.body_q.numpy().copy()
    dt = 1.0 / 60.0 / 10
    for _ in range(100 * 10):
        s0.clear_forces()
        pipe.collide(s0, contacts)
        solver.step(s0, s1, ctl, contacts, dt)
        s0, s1 = s1, s0
    q = s0.body_q.numpy()
    for i in range(model.body_count):
        disp = np.linalg.norm(q[i, :3] - q_init[i, :3])
        ang = 2.0 * math.acos(min(abs(float(np.dot(q[i, 3:], q_init[i, 3:]))), 1.0))
        assert disp < 0.15 * cube, f"body {i} moved {disp:.4f}"
        assert ang < math.radians(10.0), f"body {i} turned {math.degrees(ang):.2f} deg"


@pytest.mark.parametrize("solver_name", ["xpbd", "featherstone"])
def test_shapes_dropped_on_plane_come_to_rest(oracle_lib, solver_name):
    """newton/tests/test_rigid_contact.py:28-235 (test_shapes_on_plane), primitive shapes only (the triangle mesh is out of
    scope): spheres, lying capsules, boxes and upright cylinders of two sizes dropped from 0.5 m come to rest on the plane
    at their resting heights, upright, |v| and |w| < 0.2 - for SolverXPBD and for SolverFeatherstone (penalty contacts,
    angular_damping 0.15, friction_smoothing 2.0), 120 frames x 30 substeps."""
    b = ModelBuilder()
    cfg = b.default_shape_cfg
    cfg.ke, cfg.kd, cfg.kf, cfg.gap = 1e4, 1000.0, 0.0, 0.1
    expected = []
    z_to_y = X.quat_between_vectors((0.0, 0.0, 1.0), (0.0, 1.0, 0.0))
    for i, s in enumerate((0.5, 1.0)):
        y = 1.5 * i
        body = b.add_body(xform=X.transform((0.0, y, 0.5)))
        b.add_shape_sphere(body, radius=0.1 * s)
        expected.append((0.0, y, 0.1 * s))
        body = b.add_body(xform=X.transform((2.0, y, 0.5)))
        b.add_shape_capsule(body, xform=X.transform((0.0, 0.0, 0.0), z_to_y), radius=0.1 * s, half_height=0.3 * s)
        expected.append((2.0, y, 0.1 * s))
        body = b.add_body(xform=X.transform((4.0, y, 0.5)))
        b.add_shape_box(body, hx=0.2 * s, hy=0.25 * s, hz=0.3 * s)
        expected.append((4.0, y, 0.3 * s))
        body = b.add_body(xform=X.transform((5.0, y, 0.5)))
        b.add_shape_cylinder(body, radius=0.1 * s, half_height=0.3 * s)
        expected.append((5.0, y, 0.3 * s))
    b.add_ground_plane()
    model = b.finalize()
    if solver_name == "featherstone":
        solver = oracle_lib.SolverFeatherstone(model, angular_damping=0.15, friction_smoothing=2.0)
    else:
        solver = oracle_lib.SolverXPBD(model)
    pipe = oracle_lib.CollisionPipeline(model)
    contacts = pipe.contacts()
    s0, s1, ctl = model.state(), model.state(), model.control()
    dt = 1.0 / 60.0 / 30
    for _ in range(120 * 30):
        s0.clear_forces()
        pipe.collide(s0, contacts)
        solver.step(s0, s1, ctl, contacts, dt)
        s0, s1 = s1, s0
    q, qd = s0.body_q.numpy(), s0.body_qd.numpy()
    assert np.isfinite(q).all() and np.isfinite(qd).all()
    assert np.abs(qd[:, :3]).max() < 0.2 and np.abs(qd[:, 3:]).max() < 0.2
    np.testing.assert_allclose(q[:, :3], np.array(expected), atol=0.25)
    np.testing.assert_allclose(q[:, 3:], np.tile([0.0, 0.0, 0.0, 1.0], (model.body_count, 1)), atol=1e-1)


# ---- newton/tests/test_body_force.py: wrenches in body_f / Control.joint_f ----------------------------------------------

@pytest.mark.parametrize("solver_name", ["featherstone", "xpbd"])
@pytest.mark.parametrize("angular", [False, True])
@pytest.mark.parametrize("use_control", [False, True])
def test_body_force_floating_body(oracle_lib, solver_name, angular, use_control):
    """test_body_force.py:31-95: one step of dt = 0.1 with a 1000 N force / 1000 Nm torque on a rotated free cube: the
    driven velocity component is F/m dt (tau/I dt) within 5 %, every other component < 1e-3."""
    import newton_b200

    b = ModelBuilder(up_axis="Y", gravity=0.0)
    rot = X.quat_from_axis_angle((1.0, 0.0, 0.0), 0.5 * math.pi)
    body = b.add_body(xform=X.transform((1.0, 2.0, 3.0), rot))
    b.add_shape_box(body, hx=0.5, hy=0.5, hz=0.5)
    model = b.finalize()
    host_fk(model, model.joint_q, model.joint_qd, model)
    solver = _solver(oracle_lib, solver_name, model) if solver_name == "featherstone" else oracle_lib.SolverXPBD(model, angular_damping=0.0)
    s0, s1 = model.state(), model.state()
    ctl = model.control() if use_control else None
    idx = 5 if angular else 1
    expected = 1000.0 / (float(model.body_inertia[0, 2, 2]) if angular else float(model.body_mass[0])) * 0.1
    if use_control:
        ctl.joint_f[idx] = 1000.0
    else:
        s0.body_f[0, idx] = 1000.0
        s1.body_f[0, idx] = 1000.0
    solver.step(s0, s1, ctl, None, 0.1)
    qd = s1.body_qd.numpy()[0]
    assert qd[idx] == pytest.approx(expected, rel=5e-2)
    for i in range(6):
        if i != idx:
            assert abs(qd[i]) < 1e-3


@pytest.mark.parametrize("angular", [False, True])
def test_body_force_3d_articulation_featherstone(oracle_lib, angular):
    """test_body_force.py:98-158: a box on a 6-dof D6 joint (3 prismatic + 3 revolute axes); 1000 N (Nm) for 0.1 s along
    each axis gives exactly 0.1 m/s (0.24 / 0.282353 / 0.96 rad/s) within 1e-4 - an exact check of S, H and the solve."""
    from newton_b200.sim.builder import JointDofConfig

    b = ModelBuilder(up_axis="Y", gravity=0.0)
    b.default_shape_cfg.density = 1000.0
    link = b.add_link()
    b.add_shape_box(link, hx=0.25, hy=0.5, hz=1.0)
    j = b.add_joint_d6(-1, link, linear_axes=[JointDofConfig(axis=a) for a in "xyz"], angular_axes=[JointDofConfig(axis=a) for a in "xyz"])
    b.add_articulation([j])
    model = b.finalize()
    assert model.joint_dof_count == 6
    for dim, ang_value in enumerate((0.24, 0.282353, 0.96)):
        solver = oracle_lib.SolverFeatherstone(model, angular_damping=0.0)
        s0, s1 = model.state(), model.state()
        idx = dim + 3 if angular else dim
        s0.body_f[0, idx] = 1000.0
        s1.body_f[0, idx] = 1000.0
        solver.step(s0, s1, None, None, 0.1)
        qd = s1.body_qd.numpy()[0]
        assert qd[idx] == pytest.approx(ang_value if angular else 0.1, abs=1e-4)
        for i in range(6):
            if i != idx:
                assert abs(qd[i]) < 1e-2


@pytest.mark.parametrize("solver_name", ["featherstone", "xpbd"])
@pytest.mark.parametrize("com", [(0.5, 0.0, 0.0), (0.0, 0.3, -0.2)])
def test_body_force_at_com_offset_causes_no_rotation(oracle_lib, solver_name, com):
    """test_body_force.py:356-444: a pure force acts at the (offset) centre of mass: linear acceleration F/m, no spin."""
    import newton_b200

    b = ModelBuilder(gravity=0.0)
    body = b.add_body(xform=X.transform((0.0, 0.0, 1.0), X.quat_from_axis_angle((1.0, 0.0, 0.0), 0.5 * math.pi)))
    b.add_shape_box(body, hx=0.1, hy=0.1, hz=0.1)
    b.body_com[body] = np.array(com)
    model = b.finalize()
    host_fk(model, model.joint_q, model.joint_qd, model)
    solver = _solver(oracle_lib, solver_name, model) if solver_name == "featherstone" else oracle_lib.SolverXPBD(model, angular_damping=0.0)
    s0, s1 = model.state(), model.state()
    direction = np.array([0.0, 1.0, 0.0], dtype=np.float32)
    wrench = np.concatenate([10.0 * direction, np.zeros(3)]).astype(np.float32)
    for _ in range(5):
        s0.body_f[0] = torch_f32(wrench)
        s1.body_f[0] = torch_f32(wrench)
        solver.step(s0, s1, None, None, 0.01)
        s0, s1 = s1, s0
    qd = s0.body_qd.numpy()[0]
    expected = 10.0 / float(model.body_mass[0]) * 0.05
    assert np.abs(qd[3:]).max() < 1e-3
    assert float(np.dot(direction, qd[:3])) == pytest.approx(expected, rel=5e-2)


# ---- newton/tests/test_kinematics.py: eval_fk / eval_ik ----------------------------------------------------------------------

def test_fk_ik_revolute_small_and_large_angles(oracle_lib):
    """test_kinematics.py:95-111: eval_ik(eval_fk(q)) == q to 1e-6 for q in {+-4.0, +-5e-4, +-1e-4} rad - pins the range and
    small-angle behaviour of wp.quat_twist_angle_signed."""
    import torch

    b = ModelBuilder(gravity=0.0)
    child = b.add_link()
    j = b.add_joint_revolute(parent=-1, child=child, axis=(0.0, 0.0, 1.0))
    b.add_articulation([j])
    model = b.finalize()
    state = model.state()
    q_ik, qd_ik = torch.zeros_like(model.joint_q), torch.zeros_like(model.joint_qd)
    for angle in (-4.0, -5.0e-4, -1.0e-4, 1.0e-4, 5.0e-4, 4.0):
        state.joint_q[0] = angle
        oracle_lib.eval_fk(model, state.joint_q, state.joint_qd, state)
        oracle_lib.eval_ik(model, state, q_ik, qd_ik)
        assert float(q_ik[0]) == pytest.approx(float(np.float32(angle)), abs=1e-6)


def test_fk_ik_two_link_arm_analytic(oracle_lib):
    """test_kinematics.py:114-188: planar 2-link arm (L1 = 1.0, L2 = 0.8): FK positions against the closed form, IK recovers
    the joint angles, tolerance 1e-4."""
    import torch

    L1, L2 = 1.0, 0.8
    b = ModelBuilder(up_axis="Y", gravity=0.0)
    link0 = b.add_link()
    b.add_shape_sphere(link0, radius=0.01)
    link1 = b.add_link()
    b.add_shape_sphere(link1, radius=0.01)
    j0 = b.add_joint_revolute(parent=-1, child=link0, axis=(0.0, 0.0, 1.0), child_xform=X.transform((0.0, L1, 0.0)))
    j1 = b.add_joint_revolute(parent=link0, child=link1, axis=(0.0, 0.0, 1.0), child_xform=X.transform((0.0, L2, 0.0)))
    b.add_articulation([j0, j1])
    model = b.finalize()
    for t1, t2 in ((0.0, 0.0), (0.3, 0.0), (0.0, -0.5), (math.pi / 4, math.pi / 4), (0.3, -0.2)):
        state = model.state()
        state.joint_q[0], state.joint_q[1] = t1, t2
        oracle_lib.eval_fk(model, state.joint_q, state.joint_qd, state)
        q = state.body_q.numpy()
        np.testing.assert_allclose(q[0, :3], [L1 * math.sin(t1), -L1 * math.cos(t1), 0.0], atol=1e-4)
        np.testing.assert_allclose(q[1, :3], [L1 * math.sin(t1) + L2 * math.sin(t1 + t2), -L1 * math.cos(t1) - L2 * math.cos(t1 + t2), 0.0],
                                   atol=1e-4)
        q_ik, qd_ik = torch.zeros_like(model.joint_q), torch.zeros_like(model.joint_qd)
        oracle_lib.eval_ik(model, state, q_ik, qd_ik)
        np.testing.assert_allclose(q_ik.numpy(), [t1, t2], atol=1e-4)


def test_fk_descendant_velocity_matches_finite_difference(oracle_lib):
    """test_kinematics.py:191-257: the COM twist eval_fk reports for the tip of a 2-revolute chain with offset anchors and an
    offset parent COM agrees with the finite difference of its origin position (5e-3)."""
    b = ModelBuilder(up_axis="Y", gravity=0.0)
    link0, link1 = b.add_link(), b.add_link()
    b.body_com[link0] = np.array([0.35, 0.0, 0.0])
    j0 = b.add_joint_revolute(parent=-1, child=link0, axis=(0.0, 0.0, 1.0))
    j1 = b.add_joint_revolute(parent=link0, child=link1, axis=(0.0, 0.0, 1.0), parent_xform=X.transform((1.0, 0.0, 0.0)),
                              child_xform=X.transform((0.2, 0.0, -0.15)))
    b.add_articulation([j0, j1])
    model = b.finalize()
    s0, s1 = model.state(), model.state()
    q, qd = np.array([0.7, -0.35], dtype=np.float32), np.array([1.1, -0.45], dtype=np.float32)
    dt = 1.0e-4
    s0.joint_q.copy_(torch_f32(q)); s0.joint_qd.copy_(torch_f32(qd))
    s1.joint_q.copy_(torch_f32(q + qd * dt)); s1.joint_qd.copy_(torch_f32(qd))
    oracle_lib.eval_fk(model, s0.joint_q, s0.joint_qd, s0)
    oracle_lib.eval_fk(model, s1.joint_q, s1.joint_qd, s1)
    bq, bq1, bqd = s0.body_q.numpy(), s1.body_q.numpy(), s0.body_qd.numpy()
    fd = (bq1[link1, :3] - bq[link1, :3]) / dt
    com_world = X.quat_rotate(bq[link1, 3:].astype(np.float64), model.body_com.numpy()[link1].astype(np.float64))
    origin_vel = bqd[link1, :3] - np.cross(bqd[link1, 3:], com_world)
    np.testing.assert_allclose(fd, origin_vel, atol=5e-3)


def _free_or_distance(b, kind, parent, child, parent_xform, child_xform):
    if kind == "free":
        return b.add_joint_free(child, parent=parent, parent_xform=parent_xform, child_xform=child_xform)
    return b.add_joint_distance(parent, child, parent_xform=parent_xform, child_xform=child_xform, min_distance=-1.0, max_distance=-1.0)


@pytest.mark.parametrize("kind", ["free", "distance"])
def test_fk_free_root_descendant_velocity_matches_finite_difference(oracle_lib, kind):
    """test_kinematics.py:377-517: FREE / DISTANCE-rooted chain with offset COMs on the root and on a revolute descendant
    with non-trivial anchors; the root joint_qd is (v_com_world, omega_world).  The origin velocities recovered from
    eval_fk's body_qd agree with forward differences of body_q for both bodies (5e-3)."""
    b = ModelBuilder(up_axis="Y", gravity=0.0)
    base, child = b.add_link(), b.add_link()
    b.body_com[base] = np.array([0.15, -0.1, 0.05])
    b.body_com[child] = np.array([0.3, 0.1, -0.15])
    j0 = _free_or_distance(b, kind, -1, base, X.transform((0.0, 0.0, 0.0)), X.transform((0.0, 0.0, 0.0)))
    j1 = b.add_joint_revolute(parent=base, child=child, axis=(0.0, 0.0, 1.0), parent_xform=X.transform((0.75, 0.25, -0.1)),
                              child_xform=X.transform((0.2, -0.05, 0.15)))
    b.add_articulation([j0, j1])
    model = b.finalize()
    q, qd = model.joint_q.numpy().astype(np.float64), model.joint_qd.numpy().astype(np.float64)
    root_rot = X.quat_from_axis_angle(np.array([0.3, 1.0, -0.2]) / np.linalg.norm([0.3, 1.0, -0.2]), 0.55)
    q[0:3], q[3:7], q[7] = [0.2, -0.1, 0.15], root_rot, 0.35
    v_com, omega = np.array([0.6, -0.3, 0.2]), np.array([0.4, 0.25, -0.5])
    qd[0:3], qd[3:6], qd[6] = v_com, omega, -0.7
    dt = 1.0e-4
    com_local = np.array([0.15, -0.1, 0.05])
    rot_next = X.quat_mul(X.quat_from_axis_angle(omega / np.linalg.norm(omega), np.linalg.norm(omega) * dt), root_rot)
    com_next = q[0:3] + X.quat_rotate(root_rot, com_local) + v_com * dt
    q_next = q.copy()
    q_next[0:3], q_next[3:7], q_next[7] = com_next - X.quat_rotate(rot_next, com_local), rot_next, q[7] + qd[6] * dt
    s0, s1 = model.state(), model.state()
    s0.joint_q.copy_(torch_f32(q)); s0.joint_qd.copy_(torch_f32(qd))
    s1.joint_q.copy_(torch_f32(q_next)); s1.joint_qd.copy_(torch_f32(qd))
    oracle_lib.eval_fk(model, s0.joint_q, s0.joint_qd, s0)
    oracle_lib.eval_fk(model, s1.joint_q, s1.joint_qd, s1)
    bq, bq1, bqd = s0.body_q.numpy(), s1.body_q.numpy(), s0.body_qd.numpy()
    for body in (base, child):
        fd = (bq1[body, :3] - bq[body, :3]) / dt
        com_w = X.quat_rotate(bq[body, 3:].astype(np.float64), model.body_com.numpy()[body].astype(np.float64))
        np.testing.assert_allclose(fd, bqd[body, :3] - np.cross(bqd[body, 3:], com_w), atol=5e-3)


@pytest.mark.parametrize("kind", ["free", "distance"])
def test_ik_free_descendant_recovers_joint_state(oracle_lib, kind):
    """test_kinematics.py:520-575: a FREE / DISTANCE joint hanging off a revolute-jointed parent with offset COMs and anchors:
    eval_ik(eval_fk(q, qd)) recovers q and qd to 1e-5."""
    import torch

    b = ModelBuilder(up_axis="Y", gravity=0.0)
    base, child = b.add_link(), b.add_link()
    b.body_com[base] = np.array([0.25, -0.1, 0.0])
    b.body_com[child] = np.array([0.3, 0.15, -0.2])
    j0 = b.add_joint_revolute(parent=-1, child=base, axis=(0.0, 0.0, 1.0))
    j1 = _free_or_distance(b, kind, base, child, X.transform((1.0, 0.2, 0.3)), X.transform((0.1, -0.05, 0.2)))
    b.add_articulation([j0, j1])
    model = b.finalize()
    state = model.state()
    q, qd = state.joint_q.numpy().copy(), state.joint_qd.numpy().copy()
    q[0] = 0.35
    q[1:4] = [0.4, -0.2, 0.3]
    q[4:8] = X.quat_from_axis_angle(np.array([1.0, 2.0, -1.0]) / math.sqrt(6.0), 0.45)
    qd[0] = 0.9
    qd[1:7] = [0.2, -0.15, 0.1, 0.4, -0.3, 0.25]
    state.joint_q.copy_(torch_f32(q)); state.joint_qd.copy_(torch_f32(qd))
    oracle_lib.eval_fk(model, state.joint_q, state.joint_qd, state)
    rq, rqd = torch.zeros_like(state.joint_q), torch.zeros_like(state.joint_qd)
    oracle_lib.eval_ik(model, state, rq, rqd)
    np.testing.assert_allclose(rq.numpy(), q, atol=1e-5)
    np.testing.assert_allclose(rqd.numpy(), qd, atol=1e-5)



def test_narrow_phase_barrel_cylinder_and_cone(oracle_lib):
    """newton/tests/test_narrow_phase.py:1268-1327: barrel cylinders through the generic path - a sphere against the curved
    barrel, an infinite plane against a barrel lying on its side (plane -> box proxy, collision_core.py:566-626) and against
    its cap (>= 3 manifold points, all penetrating); plus cones resting on a plane on their base / on their side."""
    s = math.sqrt(0.5)
    cnt, dist, _, _ = oracle_lib.convex_pair(GeoType.SPHERE, (0.25, 0, 0), _xf([1.7, 0, 0]), GeoType.CYLINDER, (0.5, 1.0, 1.0), I7)
    assert cnt > 0 and dist[0] < 0.0
    cnt, dist, _, n = oracle_lib.convex_pair(GeoType.PLANE, (0, 0, 0), I7, GeoType.CYLINDER, (0.5, 1.0, 1.0), _xf([0, 0, 1.45], (0, s, 0, s)))
    assert cnt > 0 and dist[0] < 0.0
    assert dist[0] == pytest.approx(1.45 - 1.5, abs=1e-3)  # equatorial radius 0.5 + 1.0 - sqrt(1.0^2 - 1.0^2)
    np.testing.assert_allclose(n[0], [0, 0, 1], atol=1e-4)
    cnt, dist, _, _ = oracle_lib.convex_pair(GeoType.PLANE, (0, 0, 0), I7, GeoType.CYLINDER, (0.5, 1.0, 1.5), _xf([0, 0, 0.99]))
    assert cnt >= 3 and np.all(dist[:cnt] < 0.0)
    cnt, dist, pos, n = oracle_lib.convex_pair(GeoType.PLANE, (0, 0, 0), I7, GeoType.CONE, (0.5, 0.5, 0), _xf([0, 0, 0.49]))
    assert cnt >= 3 and np.allclose(dist[:cnt], -0.01, atol=1e-4)  # base disc 1 cm into the plane
    assert np.allclose(np.linalg.norm(pos[:cnt, :2], axis=1), 0.5, atol=5e-3)  # manifold points on the base rim
    cnt, dist, _, n = oracle_lib.convex_pair(GeoType.PLANE, (0, 0, 0), I7, GeoType.CONE, (0.5, 0.5, 0), _xf([0, 0, 0.3], (0, s, 0, s)))
    assert cnt >= 1 and dist[0] == pytest.approx(-0.2, abs=1e-3)  # axis horizontal: the base rim reaches 0.5 below the centre

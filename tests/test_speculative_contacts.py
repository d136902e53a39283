"""Speculative contacts: `CollisionPipeline(speculative_config=SpeculativeContactConfig(...))` + `collide(..., dt=...)`.

Reference: `newton/_src/sim/collide.py:257-280` (`write_contact_speculative`), `:475-541` (`compute_shape_velocities`), `:1076-1102`
(`SpeculativeContactConfig`), `:1823-1836, 1877-1962` (the `collide()` flow), `geometry/contact_data.py:92-233` (approach speed,
predictive score, admission), `geometry/broad_phase_common.py:41-80` (swept AABB overlap), `broad_phase_sap.py:44-78` (projection
extended by the displacement), `narrow_phase.py:241-246, 885-888, 1170-1175`.

The CPU part pins the ORACLE with the known answers of the reference's own `newton/tests/test_speculative_contacts.py` (rigid primitive /
GJK rows; the mesh-SDF and contact-reducer rows are outside the path); the GPU part compares the CUDA path with the oracle.
"""

import numpy as np
import pytest
import torch

import newton_b200
from newton_b200 import SpeculativeContactConfig
from newton_b200.sim.builder import ModelBuilder
from newton_b200.utils import xform as X

BROAD_PHASES = ("nxn", "sap", "explicit")


def _build_spheres(velocity, separation=0.3, gap=0.0):
    """test_speculative_contacts.py:538-548: two spheres of radius 0.1 along X, the first one moving"""
    b = ModelBuilder(gravity=0.0)
    b.rigid_gap = gap
    a = b.add_body(xform=X.transform((0.0, 0.0, 0.0)))
    b.add_shape_sphere(a, radius=0.1)
    b.body_qd[a] = np.array([velocity, 0.0, 0.0, 0.0, 0.0, 0.0])
    c = b.add_body(xform=X.transform((separation, 0.0, 0.0)))
    b.add_shape_sphere(c, radius=0.1)
    return b.finalize()


def _count(contacts):
    return int(contacts.rigid_contact_count.cpu().numpy()[0])


def _collide(lib, model, speculative, dt=0.02, broad_phase="nxn", ext=0.25, state=None):
    cfg = SpeculativeContactConfig(max_speculative_extension=ext) if speculative else None
    pipe = lib.CollisionPipeline(model, broad_phase=broad_phase, speculative_config=cfg)
    contacts = pipe.contacts()
    pipe.collide(state or model.state(), contacts, dt=dt)
    return contacts, pipe


def test_candidates_are_opt_in(oracle_lib):  # :612-616
    model = _build_spheres(10.0)
    assert _count(_collide(oracle_lib, model, False)[0]) == 0
    assert _count(_collide(oracle_lib, model, True)[0]) > 0


@pytest.mark.parametrize("velocity", [0.0, -10.0])
def test_candidates_require_approach(oracle_lib, velocity):  # :619-624
    assert _count(_collide(oracle_lib, _build_spheres(velocity), True)[0]) == 0


def test_candidates_require_dt(oracle_lib):  # :627-643
    model = _build_spheres(10.0)
    pipe = oracle_lib.CollisionPipeline(model, broad_phase="nxn", speculative_config=SpeculativeContactConfig(max_speculative_extension=0.25))
    contacts = pipe.contacts()
    with pytest.raises(ValueError, match="dt must be provided"):
        pipe.collide(model.state(), contacts)
    pipe.collide(model.state(), contacts, dt=0.02)  # 10 m/s * 0.02 s = 0.2 m >= the 0.1 m clearance
    assert _count(contacts) > 0
    pipe.collide(model.state(), contacts, dt=0.005)  # 0.05 m < 0.1 m
    assert _count(contacts) == 0


def test_gap_uses_larger_of_fixed_and_velocity_distance(oracle_lib):  # :646-660
    model = _build_spheres(1.0, separation=0.33, gap=0.05)  # clearance 0.13, authored pair gap 0.1
    pipe = oracle_lib.CollisionPipeline(model, broad_phase="nxn", speculative_config=SpeculativeContactConfig(max_speculative_extension=0.25))
    contacts = pipe.contacts()
    pipe.collide(model.state(), contacts, dt=0.05)  # extension 0.05: neither 0.1 nor 0.05 reaches 0.13 (they are not added)
    assert _count(contacts) == 0
    pipe.collide(model.state(), contacts, dt=0.15)  # extension 0.15 >= 0.13
    assert _count(contacts) > 0


@pytest.mark.parametrize("dt", [-0.01, float("nan"), float("inf"), float("-inf")])
def test_invalid_dt_is_rejected(oracle_lib, dt):  # :663-674
    model = _build_spheres(10.0)
    pipe = oracle_lib.CollisionPipeline(model, broad_phase="nxn", speculative_config=SpeculativeContactConfig())
    with pytest.raises(ValueError, match="dt must be a non-negative finite number"):
        pipe.collide(model.state(), pipe.contacts(), dt=dt)


def _common_motion_model():
    b = ModelBuilder(gravity=0.0)
    b.rigid_gap = 0.0
    a = b.add_body(xform=X.transform_identity())
    b.add_shape_sphere(a, radius=0.1)
    b.body_qd[a] = np.array([20.0, 0.0, 0.0, 0.0, 0.0, 0.0])
    c = b.add_body(xform=X.transform((0.4, 0.0, 0.0)))
    b.add_shape_sphere(c, radius=0.1)
    b.body_qd[c] = np.array([20.0, 0.0, 0.0, 0.0, 0.0, 0.0])
    return b.finalize()


@pytest.mark.parametrize("broad_phase", BROAD_PHASES)
def test_common_motion_is_rejected_by_the_swept_broad_phase(oracle_lib, broad_phase):  # :677-704
    """both spheres move 2 m in 0.1 s: their swept unions overlap, their RELATIVE sweep does not"""
    model = _common_motion_model()
    contacts, pipe = _collide(oracle_lib, model, True, dt=0.1, broad_phase=broad_phase)
    assert pipe.candidate_count == 0
    assert _count(contacts) == 0


def test_candidates_preserve_physical_geometry(oracle_lib):  # :707-713
    contacts, _ = _collide(oracle_lib, _build_spheres(10.0), True)
    assert _count(contacts) > 0
    assert abs(float(contacts.rigid_contact_margin0.numpy()[0]) - 0.1) < 1e-6
    assert abs(float(contacts.rigid_contact_margin1.numpy()[0]) - 0.1) < 1e-6
    # the stored points are the PHYSICAL surface points: still 0.1 m apart along the normal
    n = contacts.rigid_contact_normal.numpy()[0]
    np.testing.assert_allclose(n, [1.0, 0.0, 0.0], atol=1e-6)


def _angular_model():
    b = ModelBuilder(gravity=0.0)
    b.rigid_gap = 0.0
    body = b.add_body(xform=X.transform_identity(), mass=1.0, inertia=np.eye(3), lock_inertia=True)
    b.add_shape_sphere(body, radius=0.1, xform=X.transform((0.0, 1.0, 0.0)))
    b.body_qd[body] = np.array([0.0, 0.0, 0.0, 0.0, 0.0, -10.0])  # the shape at (0, 1, 0) moves at +10 m/s along x
    b.add_shape_sphere(-1, radius=0.1, xform=X.transform((0.3, 1.0, 0.0)))
    return b.finalize()


@pytest.mark.parametrize("broad_phase", ["nxn", "sap"])
def test_angular_motion_of_an_offset_shape_counts(oracle_lib, broad_phase):  # :716-740
    contacts, pipe = _collide(oracle_lib, _angular_model(), True, dt=0.02, broad_phase=broad_phase)
    assert pipe.candidate_count > 0
    assert _count(contacts) > 0


def _cone_model():
    b = ModelBuilder(gravity=0.0)
    b.rigid_gap = 0.0
    body = b.add_body(xform=X.transform((0.0, 0.0, 0.5)))
    b.add_shape_cone(body, radius=0.1, half_height=0.1)
    b.body_qd[body] = np.array([0.0, 0.0, -20.0, 0.0, 0.0, 0.0])
    b.add_shape_plane(width=0.0, length=0.0)
    return b.finalize()


def test_cone_reaches_infinite_plane_through_gjk(oracle_lib):  # :743-763
    """plane-cone goes through the plane -> box proxy + GJK; 0.4 m of clearance, 20 m/s * 0.03 s = 0.6 m"""
    contacts, pipe = _collide(oracle_lib, _cone_model(), True, dt=0.03, ext=0.75)
    assert pipe.candidate_count > 0
    assert _count(contacts) > 0
    assert _count(_collide(oracle_lib, _cone_model(), False, dt=0.03)[0]) == 0


def _two_boxes():
    b = ModelBuilder(gravity=0.0)
    b.rigid_gap = 0.0
    a = b.add_body()
    b.add_shape_box(a, hx=0.1, hy=0.1, hz=0.1)
    c = b.add_body(xform=X.transform((0.15, 0.0, 0.0)))
    b.add_shape_box(c, hx=0.1, hy=0.1, hz=0.1)
    return b.finalize()


def test_stationary_contacts_match_the_regular_pipeline(oracle_lib):  # :818-864
    model = _two_boxes()
    regular, _ = _collide(oracle_lib, model, False, dt=0.03)
    spec, _ = _collide(oracle_lib, model, True, dt=0.03)
    n = _count(regular)
    assert n > 0 and _count(spec) == n
    for name in ("shape0", "shape1", "point0", "point1", "normal", "margin0", "margin1"):
        a, b = getattr(regular, "rigid_contact_" + name).numpy()[:n], getattr(spec, "rigid_contact_" + name).numpy()[:n]
        np.testing.assert_allclose(b, a, rtol=0.0, atol=1e-6, err_msg=name)


def _tunnel_model():
    b = ModelBuilder(gravity=0.0)
    b.rigid_gap = 0.0
    body = b.add_body(xform=X.transform((0.0, 0.0, 0.5)))
    b.add_shape_sphere(body, radius=0.05)
    b.body_qd[body] = np.array([0.0, 0.0, -20.0, 0.0, 0.0, 0.0])
    b.add_shape_plane(width=0.0, length=0.0)
    return b.finalize(), body


def _tunnel_step(lib, solvers, model, body, speculative, dt=0.03):
    cfg = SpeculativeContactConfig(max_speculative_extension=0.75) if speculative else None
    pipe = lib.CollisionPipeline(model, broad_phase="nxn", speculative_config=cfg)
    contacts = pipe.contacts()
    s_in, s_out = model.state(), model.state()
    pipe.collide(s_in, contacts, dt=dt)
    solvers.SolverXPBD(model, iterations=5).step(s_in, s_out, None, contacts, dt)
    return float(s_out.body_q.cpu().numpy()[body, 2]), s_out


def test_speculative_contacts_prevent_tunnelling(oracle_lib):  # :867-893
    """a 5 cm sphere at 20 m/s crosses the ground in one 30 ms XPBD step unless the predicted contact is there"""
    model, body = _tunnel_model()
    assert _tunnel_step(oracle_lib, oracle_lib, model, body, False)[0] < 0.0
    assert _tunnel_step(oracle_lib, oracle_lib, model, body, True)[0] >= 0.04


def test_config_validation():  # sim/collide.py:1095-1102
    for bad in (-0.1, float("nan"), float("inf")):
        with pytest.raises(ValueError, match="max_speculative_extension"):
            SpeculativeContactConfig(max_speculative_extension=bad)
    assert SpeculativeContactConfig().max_speculative_extension == 0.1
    assert newton_b200.CollisionPipeline.SpeculativeContactConfig is SpeculativeContactConfig


# ---- GPU parity: CUDA == oracle on the same scenes, every broad phase ------------------------------------------------------------------


def _same_contacts(ref, got):
    n = _count(ref)
    assert _count(got) == n
    for name in ("shape0", "shape1", "point0", "point1", "offset0", "offset1", "normal", "margin0", "margin1"):
        a, b = getattr(ref, "rigid_contact_" + name).numpy()[:n], getattr(got, "rigid_contact_" + name).cpu().numpy()[:n]
        np.testing.assert_array_equal(b, a, err_msg=name)


GPU_SCENES = {
    "spheres_hit": (lambda: _build_spheres(10.0), 0.02, 0.25),
    "spheres_miss": (lambda: _build_spheres(10.0), 0.005, 0.25),
    "spheres_diverge": (lambda: _build_spheres(-10.0), 0.02, 0.25),
    "spheres_gap": (lambda: _build_spheres(1.0, separation=0.33, gap=0.05), 0.15, 0.25),
    "common_motion": (_common_motion_model, 0.1, 0.25),
    "angular": (_angular_model, 0.02, 0.25),
    "cone_plane": (_cone_model, 0.03, 0.75),
    "boxes": (_two_boxes, 0.03, 0.25),
    "inactive_dt0": (lambda: _build_spheres(10.0, separation=0.19), 0.0, 0.25),
}


@pytest.mark.gpu
@pytest.mark.parametrize("broad_phase", BROAD_PHASES)
@pytest.mark.parametrize("scene", sorted(GPU_SCENES))
def test_gpu_speculative_collide_matches_oracle(oracle_lib, cuda_lib, scene, broad_phase):
    make, dt, ext = GPU_SCENES[scene]
    model = make()
    ref, rp = _collide(oracle_lib, model, True, dt=dt, broad_phase=broad_phase, ext=ext)
    mg = model.to("cuda:0")
    got, gp = _collide(newton_b200, mg, True, dt=dt, broad_phase=broad_phase, ext=ext)
    torch.cuda.synchronize()
    _same_contacts(ref, got)


@pytest.mark.gpu
def test_gpu_speculative_heap_simulation_matches_oracle(oracle_lib, cuda_lib):
    """fast free bodies (spheres, boxes, capsules; mixed collision groups) thrown at each other and at the ground: 40 substeps with speculative
    contacts, contact counts per substep and the final state equal the oracle's"""
    from newton_b200 import scenes
    from tests.helpers import simulate

    model = scenes.free_bodies_model(3, seed=5)
    g = torch.Generator().manual_seed(2)
    model.body_qd.copy_((torch.rand(model.body_qd.shape, generator=g) * 2.0 - 1.0) * torch.tensor([6.0, 6.0, 6.0, 3.0, 3.0, 3.0]))
    cfg = SpeculativeContactConfig(max_speculative_extension=0.3)
    for bp in BROAD_PHASES:
        kw = {"broad_phase": bp, "speculative_config": cfg}
        rs, _, rc = simulate(model, oracle_lib.CollisionPipeline, oracle_lib.SolverXPBD, substeps=40, dt=0.004, solver_kwargs={"iterations": 4},
                             pipeline_kwargs=kw, record_contacts=True, collide_dt=0.06)
        gs, _, gc = simulate(model.to("cuda:0"), newton_b200.CollisionPipeline, newton_b200.solvers.SolverXPBD, substeps=40, dt=0.004,
                             solver_kwargs={"iterations": 4}, pipeline_kwargs=kw, record_contacts=True, collide_dt=0.06)
        torch.cuda.synchronize()
        assert rc == gc, bp
        assert sum(rc) == 2062  # 2012 without speculation (horizon 60 ms at up to 6 m/s: 50 predicted contacts over the run)
        np.testing.assert_array_equal(gs.body_q.cpu().numpy(), rs.body_q.numpy(), err_msg=bp)
        np.testing.assert_array_equal(gs.body_qd.cpu().numpy(), rs.body_qd.numpy(), err_msg=bp)


@pytest.mark.gpu
def test_gpu_speculative_contacts_prevent_tunnelling(oracle_lib, cuda_lib):
    model, body = _tunnel_model()
    z_ref, s_ref = _tunnel_step(oracle_lib, oracle_lib, model, body, True)
    mg = model.to("cuda:0")
    z, s = _tunnel_step(newton_b200, newton_b200.solvers, mg, body, True)
    assert z >= 0.04
    np.testing.assert_array_equal(s.body_q.cpu().numpy(), s_ref.body_q.numpy())
    assert _tunnel_step(newton_b200, newton_b200.solvers, mg, body, False)[0] < 0.0

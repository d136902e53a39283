"""MESH shapes against infinite planes (reference ``narrow_phase_process_mesh_plane_contacts_kernel``, ``narrow_phase.py:1761-1861``,
``reduce_contacts=False``): one contact per mesh vertex within gap + margin of the plane.

CPU part: the oracle restatement pinned by closed forms (which vertices make contacts, where they land in the body frames), by the
analytic plane-box collider on the same cube, and by the reference's ``test_mesh_box_on_ground`` scenario
(``newton/tests/test_rigid_contact.py:513-620``: a box mesh set on the ground stays put under XPBD).
GPU part: the CUDA path (group-cooperative vertex walk inside ``collide_kernel``) against the oracle, bit for bit."""

import numpy as np
import pytest

import newton_b200
from newton_b200.geometry.mesh import Mesh
from newton_b200.sim.builder import ModelBuilder
from newton_b200.utils import xform as X
from tests.helpers import canonical_contacts, simulate

BOX_TRIS = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4],
                     [1, 5, 7], [1, 7, 3]], dtype=np.int32)


def box_mesh(h=0.5):
    c = np.array([[sx * h, sy * h, sz * h] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float32)
    return Mesh(c, BOX_TRIS)


def blob_mesh(seed, n=37, radius=0.4):
    """Random points on a squashed sphere + their hull faces: a closed, outward-oriented triangle mesh with n vertices."""
    rng = np.random.default_rng(seed)
    p = rng.normal(size=(n, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    p *= radius * np.array([1.0, 0.8, 0.6])
    return Mesh(p.astype(np.float32))  # no indices: Mesh takes the hull faces for the mass properties


def single_box_model(z, rot=None, gap=None, plane_first=False):
    b = ModelBuilder()
    if gap is not None:
        b.rigid_gap = gap
    if plane_first:
        b.add_ground_plane()
    body = b.add_body(xform=X.transform((0.0, 0.0, z), rot if rot is not None else X.quat_identity()))
    b.add_shape_mesh(body, mesh=box_mesh())
    if not plane_first:
        b.add_ground_plane()
    return b.finalize()


def test_only_vertices_near_the_plane_make_contacts(oracle_lib):
    m = single_box_model(0.5)  # bottom face on the plane, top face one metre above the gap
    pipe = oracle_lib.CollisionPipeline(m)
    c = pipe.contacts()
    pipe.collide(m.state(), c)
    n, cc = canonical_contacts(c, m)
    assert n == 4
    mesh_shape, plane_shape = 0, 1
    assert (cc["shape0"] == mesh_shape).all() and (cc["shape1"] == plane_shape).all()  # stored (mesh, plane), narrow_phase.py:628
    np.testing.assert_array_equal(cc["normal"], np.tile([0.0, 0.0, -1.0], (4, 1)).astype(np.float32))  # from the mesh to the plane
    # vertex order of the mesh = sub key order: the four z = -0.5 corners, in file order
    verts = m.hull_points.numpy()
    bottom = verts[verts[:, 2] < 0]
    np.testing.assert_allclose(cc["point0"], bottom, atol=1e-7)             # body frame of the box: the vertex itself (distance 0)
    np.testing.assert_allclose(cc["point1"][:, :2], bottom[:, :2], atol=1e-7)  # world frame (static plane): its projection
    np.testing.assert_allclose(cc["point1"][:, 2], 0.0, atol=1e-7)
    assert (cc["margin0"] == 0).all() and (cc["margin1"] == 0).all()


@pytest.mark.parametrize("lift,expected", [(0.09, 4), (0.11, 0), (-0.2, 4)])
def test_gap_threshold(oracle_lib, lift, expected):
    m = single_box_model(0.5 + lift, gap=0.05)  # gap_sum = 0.1 (shape + plane)
    pipe = oracle_lib.CollisionPipeline(m)
    c = pipe.contacts()
    pipe.collide(m.state(), c)
    assert int(c.rigid_contact_count[0]) == expected


def test_tilted_cube_matches_the_analytic_box_collider(oracle_lib):
    """Same cube as a BOX primitive: plane_box reports the same corner points (there as shape b of a (plane, box) pair)."""
    rot = X.quat_from_axis_angle((1.0, 0.3, 0.0), 0.02)
    mm = single_box_model(0.52, rot)
    b = ModelBuilder()
    body = b.add_body(xform=X.transform((0.0, 0.0, 0.52), rot))
    b.add_shape_box(body, hx=0.5, hy=0.5, hz=0.5)
    b.add_ground_plane()
    mb = b.finalize()
    out = []
    for m in (mm, mb):
        pipe = oracle_lib.CollisionPipeline(m)
        c = pipe.contacts()
        pipe.collide(m.state(), c)
        out.append(canonical_contacts(c, m))
    (nm, cm), (nb, cbx) = out
    assert nm == nb == 4
    key = lambda p: np.lexsort((p[:, 1].round(4), p[:, 0].round(4)))  # noqa: E731
    pm, pb = cm["point0"], cbx["point1"]  # box-frame contact points: side a of the mesh pair, side b of the box pair
    np.testing.assert_allclose(pm[key(pm)], pb[key(pb)], atol=2e-6)
    np.testing.assert_allclose(cm["normal"], -cbx["normal"], atol=1e-7)


def test_mesh_box_rests_on_the_ground_under_xpbd(oracle_lib):
    """newton/tests/test_rigid_contact.py:513-620 (test_mesh_box_on_ground): 60 frames x 10 substeps, SolverXPBD(iterations=2)."""
    b = ModelBuilder()
    b.default_shape_cfg.ke, b.default_shape_cfg.kd, b.default_shape_cfg.mu = 1.0e5, 1.0e3, 0.5
    b.add_ground_plane()
    body = b.add_body(xform=X.transform((0.0, 0.0, 0.5), X.quat_identity()))
    b.add_shape_mesh(body, mesh=box_mesh())
    m = b.finalize()
    out, _, counts = simulate(m, oracle_lib.CollisionPipeline, oracle_lib.SolverXPBD, substeps=600, dt=1.0 / 600.0,
                              solver_kwargs={"iterations": 2}, record_contacts=True)
    q, qd = out.body_q.numpy()[0], out.body_qd.numpy()[0]
    assert 0.45 < q[2] < 0.55
    assert np.abs(qd).max() < 0.01
    assert counts[-1] == 4


def test_mesh_against_other_shapes_is_refused(oracle_lib):
    b = ModelBuilder()
    b0 = b.add_body(xform=X.transform((0.0, 0.0, 0.5), X.quat_identity()))
    b.add_shape_mesh(b0, mesh=box_mesh())
    b1 = b.add_body(xform=X.transform((0.0, 0.0, 1.6), X.quat_identity()))
    b.add_shape_box(b1, hx=0.5, hy=0.5, hz=0.5)
    m = b.finalize()
    pipe = oracle_lib.CollisionPipeline(m)
    with pytest.raises(NotImplementedError):
        pipe.collide(m.state(), pipe.contacts())


def test_pipeline_asks_for_unreduced_mesh_contacts():
    m = single_box_model(0.5)
    with pytest.raises(NotImplementedError, match="reduce_contacts=False"):
        newton_b200.CollisionPipeline(m)


def mixed_mesh_model(world_count, seed=0, plane_first=False):
    """Per world: a box, a blob mesh, a sphere and a second box, dropped on a shared ground plane.  The mesh collides with the plane
    only (mesh-vs-shape pairs filtered); with the plane added last its pair sits in the MIDDLE of the world's key-ordered pair list."""
    rng = np.random.default_rng(seed)
    scene = ModelBuilder()
    if plane_first:
        scene.add_ground_plane()
    mesh = blob_mesh(5)
    for _ in range(world_count):
        scene.begin_world()
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        ba = scene.add_body(xform=X.transform((0.0, 0.0, 0.6), X.quat_from_axis_angle((0.0, 0.0, 1.0), float(rng.uniform(-0.3, 0.3)))))
        sa = scene.add_shape_box(ba, hx=0.3, hy=0.25, hz=0.2)
        bm = scene.add_body(xform=X.transform((1.5, float(rng.uniform(-0.1, 0.1)), 0.45), q))
        sm = scene.add_shape_mesh(bm, mesh=mesh, scale=(1.0, 1.2, 0.9))
        bs = scene.add_body(xform=X.transform((-1.5, 0.0, 0.5), X.quat_identity()))
        ssp = scene.add_shape_sphere(bs, radius=0.3)
        bb = scene.add_body(xform=X.transform((0.05, 0.0, 1.1), X.quat_identity()))
        sb = scene.add_shape_box(bb, hx=0.2, hy=0.2, hz=0.2)
        for other in (sa, ssp, sb):
            scene.add_shape_collision_filter_pair(sm, other)
        scene.end_world()
    if not plane_first:
        scene.add_ground_plane()
    return scene.finalize()


@pytest.mark.parametrize("plane_first", [False, True])
def test_oracle_mixed_scene_counts(oracle_lib, plane_first):
    m = mixed_mesh_model(3, plane_first=plane_first)
    out, contacts, counts = simulate(m, oracle_lib.CollisionPipeline, oracle_lib.SolverXPBD, substeps=80, dt=1.0 / 240,
                                     solver_kwargs={"iterations": 4}, record_contacts=True)
    assert np.isfinite(out.body_q.numpy()).all()
    st = m.numpy("shape_type")
    s0 = contacts.rigid_contact_shape0.numpy()[: counts[-1]]
    assert (st[s0] == 8).sum() >= 3  # every blob touches the ground by then, several vertices each
    assert out.body_q.numpy().reshape(3, 4, 7)[:, 1, 2].min() > 0.1  # the blobs did not sink through


# ---- newton/tests/test_mesh_aabb.py:17-154: the broad-phase AABB of a mesh is its rotated LOCAL box, not a bounding sphere -----------
def _box_mesh(hx, hy, hz):
    c = np.array([[sx * hx, sy * hy, sz * hz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float32)
    return Mesh(c, BOX_TRIS)


def _mesh_model(mesh, pos, rot=None, scale=None):
    b = ModelBuilder()
    body = b.add_body(xform=X.transform(pos, rot if rot is not None else X.quat_identity()))
    b.add_shape_mesh(body, mesh=mesh, **({"scale": scale} if scale is not None else {}))
    b.add_ground_plane()
    return b.finalize()


def _aabb(oracle_lib, m, shape=0):
    lo, hi = oracle_lib.shape_aabbs(m, m.body_q)
    margin = float(m.shape_margin[shape] + m.shape_gap[shape])
    return np.asarray(lo)[shape], np.asarray(hi)[shape], margin


def test_mesh_aabb_axis_aligned_is_tight(oracle_lib):
    hx, hy, hz = 0.2, 0.2, 0.05
    pos = (0.0, 0.0, 1.0)
    lo, hi, margin = _aabb(oracle_lib, _mesh_model(_box_mesh(hx, hy, hz), pos))
    np.testing.assert_allclose(lo, np.array(pos) - (hx, hy, hz) - margin, atol=1e-4)
    np.testing.assert_allclose(hi, np.array(pos) + (hx, hy, hz) + margin, atol=1e-4)


def test_mesh_aabb_flat_table_does_not_reach_the_gripper(oracle_lib):
    b = ModelBuilder()
    table = b.add_body(xform=X.transform((0.0, 0.0, 0.05), X.quat_identity()))
    st = b.add_shape_mesh(table, mesh=_box_mesh(0.2, 0.2, 0.05))
    gripper = b.add_body(xform=X.transform((0.0, 0.0, 0.375), X.quat_identity()))
    sg = b.add_shape_mesh(gripper, mesh=_box_mesh(0.03, 0.02, 0.04))
    b.add_shape_collision_filter_pair(st, sg)  # (the pair itself is a mesh-mesh route; only the boxes are looked at here)
    b.add_ground_plane()
    m = b.finalize()
    lo, hi = oracle_lib.shape_aabbs(m, m.body_q)
    margin = float(m.shape_margin[0] + m.shape_gap[0])
    assert hi[0][2] < 0.1 + margin + 0.01
    assert lo[1][2] > hi[0][2]


def test_mesh_aabb_rotated(oracle_lib):
    hx, hy, hz = 1.0, 0.1, 0.1
    rot = X.quat_from_axis_angle((0.0, 0.0, 1.0), np.pi / 2.0)
    lo, hi, margin = _aabb(oracle_lib, _mesh_model(_box_mesh(hx, hy, hz), (0.0, 0.0, 2.0), rot))
    np.testing.assert_allclose(hi - lo, 2 * np.array([hy, hx, hz]) + 2 * margin, atol=0.02)


def test_mesh_aabb_nonuniform_scale(oracle_lib):
    pos, sc = (0.0, 0.0, 5.0), (2.0, 0.5, 3.0)
    lo, hi, margin = _aabb(oracle_lib, _mesh_model(_box_mesh(1.0, 1.0, 1.0), pos, scale=sc))
    np.testing.assert_allclose(lo, np.array(pos) - sc - margin, atol=1e-4)
    np.testing.assert_allclose(hi, np.array(pos) + sc + margin, atol=1e-4)


def test_world_shards_carry_their_meshes(oracle_lib):
    """Model.shard (N > 1 path): a shard keeps indexing the shared vertex pool, so shard-by-shard == monolithic bit for bit."""
    m = mixed_mesh_model(4, seed=3)
    kw = {"iterations": 4}
    full, _, _ = simulate(m, oracle_lib.CollisionPipeline, oracle_lib.SolverXPBD, substeps=40, dt=1.0 / 240, solver_kwargs=kw)
    parts = [simulate(m.shard(r, 2), oracle_lib.CollisionPipeline, oracle_lib.SolverXPBD, substeps=40, dt=1.0 / 240, solver_kwargs=kw)[0]
             for r in range(2)]
    np.testing.assert_array_equal(np.concatenate([p.body_q.numpy() for p in parts]), full.body_q.numpy())


# ---- GPU ----------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("plane_first", [False, True])
@pytest.mark.parametrize("solver", ["xpbd", "featherstone_free"])
def test_gpu_mesh_plane_matches_oracle(oracle_lib, cuda_lib, plane_first, solver):
    m = mixed_mesh_model(9, seed=2, plane_first=plane_first)
    if solver == "xpbd":
        o_cls, g_cls, kw, dt = oracle_lib.SolverXPBD, newton_b200.solvers.SolverXPBD, {"iterations": 4}, 1.0 / 240
    else:
        o_cls, g_cls, kw, dt = oracle_lib.SolverFeatherstone, newton_b200.solvers.SolverFeatherstone, {}, 1.0 / 1000
    n = 80
    ref, rc_, rcounts = simulate(m, oracle_lib.CollisionPipeline, o_cls, substeps=n, dt=dt, solver_kwargs=kw, record_contacts=True)
    out, gc_, gcounts = simulate(m.to("cuda:0"), newton_b200.CollisionPipeline, g_cls, substeps=n, dt=dt, solver_kwargs=kw,
                                 record_contacts=True, pipeline_kwargs={"reduce_contacts": False, "deterministic": True})
    assert rcounts == gcounts
    nr, cr = canonical_contacts(rc_, m)
    ng, cg = canonical_contacts(gc_, m)
    assert nr == ng
    for k in cr:
        np.testing.assert_array_equal(cg[k], cr[k], err_msg=k)
    np.testing.assert_array_equal(out.body_q.cpu().numpy(), ref.body_q.numpy())
    np.testing.assert_array_equal(out.body_qd.cpu().numpy(), ref.body_qd.numpy())


@pytest.mark.gpu
def test_gpu_mesh_pairs_with_other_shapes_are_refused(cuda_lib):
    b = ModelBuilder()
    b0 = b.add_body(xform=X.transform((0.0, 0.0, 0.5), X.quat_identity()))
    b.add_shape_mesh(b0, mesh=box_mesh())
    b1 = b.add_body(xform=X.transform((0.0, 0.0, 1.6), X.quat_identity()))
    b.add_shape_box(b1, hx=0.5, hy=0.5, hz=0.5)
    m = b.finalize().to("cuda:0")
    with pytest.raises(Exception, match="MESH"):
        newton_b200.CollisionPipeline(m, reduce_contacts=False)

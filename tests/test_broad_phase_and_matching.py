"""Run-time broad phases ("nxn" / "sap") and frame-to-frame contact matching - SURVEY.md §8(f) rank 3 / row a6'.
CPU part: the oracle's restatements (oracle/broad_phase.py, oracle/contact_match.py) against the reference's own test expectations
(newton/tests/test_broad_phase.py pair sets vs brute force; newton/tests/test_contact_matching.py known answers).
GPU part: the CUDA broad-phase / matching kernels against those restatements."""

import numpy as np
import pytest
import torch

import newton_b200
from newton_b200 import ModelBuilder, scenes
from newton_b200.utils import xform as X
from tests.helpers import canonical_contacts, simulate


# ------------------------------------------------------------------------------------------------ broad phase, CPU
def _random_box_model(rng, n_shapes, n_worlds):
    """test_broad_phase.py:92-150 style: random boxes, random collision groups, a few shared (world -1) shapes."""
    b = ModelBuilder(gravity=(0.0, 0.0, 0.0))
    per = n_shapes // n_worlds
    for _w in range(n_worlds):
        b.begin_world()
        for _ in range(per):
            cfg = newton_b200.ShapeConfig()
            cfg.collision_group = int(rng.choice([1, 1, 2, -1, -2, 0]))
            body = b.add_body(xform=X.transform(rng.uniform(-1.0, 1.0, size=3)))
            h = rng.uniform(0.05, 0.4, size=3)
            b.add_shape_box(body, hx=h[0], hy=h[1], hz=h[2], cfg=cfg)
        b.end_world()
    for _ in range(2):
        cfg = newton_b200.ShapeConfig()
        cfg.collision_group = -5
        b.add_shape_box(-1, xform=X.transform(rng.uniform(-1.0, 1.0, size=3)), hx=0.5, hy=0.5, hz=0.1, cfg=cfg)
    return b.finalize()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_nxn_and_sap_equal_brute_force(oracle_lib, seed):
    from oracle import broad_phase as bp

    rng = np.random.default_rng(seed)
    m = _random_box_model(rng, 24, 3)
    lo, hi = oracle_lib.shape_aabbs(m, m.body_q)
    world, group, flags = m.numpy("shape_world"), m.numpy("shape_collision_group"), m.numpy("shape_flags")
    brute = set()
    for i in range(m.shape_count):
        for j in range(i + 1, m.shape_count):
            if not (flags[i] & 2 and flags[j] & 2):
                continue
            if not bp.test_world_and_group_pair(int(world[i]), int(world[j]), int(group[i]), int(group[j])):
                continue
            if (i, j) in {(min(a, c), max(a, c)) for a, c in m.shape_collision_filter_pairs}:
                continue
            if np.all(lo[i] <= hi[j]) and np.all(hi[i] >= lo[j]):
                brute.add((i, j))
    nxn = set(bp.nxn_candidate_pairs(m, lo, hi, m.shape_collision_filter_pairs))
    sap = set(bp.sap_candidate_pairs(m, lo, hi, m.shape_collision_filter_pairs))
    assert nxn == brute and sap == brute and len(brute) > 5
    # include_static_kinematic_pairs=False drops the pairs of two immovable shapes (broad_phase_common.py:166-201)
    pruned = set(bp.nxn_candidate_pairs(m, lo, hi, m.shape_collision_filter_pairs, include_static_kinematic_pairs=False))
    body = m.numpy("shape_body")
    assert pruned == {p for p in brute if not (body[p[0]] < 0 and body[p[1]] < 0)}


def test_oracle_pipelines_agree_on_free_bodies(oracle_lib):
    """explicit (builder's list), nxn and sap must give the same contacts, step for step (the reference builds the explicit list
    with 'the exact same filtering logic as the broad phase kernels', sim/builder.py:12796-12797)."""
    m = scenes.free_bodies_model(2)
    res = {}
    for mode in ("explicit", "nxn", "sap"):
        res[mode] = simulate(m, lambda mm, mode=mode: oracle_lib.CollisionPipeline(mm, broad_phase=mode), oracle_lib.SolverXPBD, substeps=40,
                             dt=1.0 / 240, solver_kwargs={"iterations": 4}, record_contacts=True)
    for mode in ("nxn", "sap"):
        assert res[mode][2] == res["explicit"][2]
        np.testing.assert_array_equal(res[mode][0].body_q.numpy(), res["explicit"][0].body_q.numpy())
    # a model without the precomputed list: explicit finds nothing, the run-time broad phases are unaffected
    bare = scenes.free_bodies_model(2, drop_pairs=True)
    none = simulate(bare, lambda mm: oracle_lib.CollisionPipeline(mm, broad_phase="explicit"), oracle_lib.SolverXPBD, substeps=5, dt=1.0 / 240,
                    record_contacts=True)
    assert none[2] == [0] * 5
    dyn = simulate(bare, lambda mm: oracle_lib.CollisionPipeline(mm, broad_phase="sap"), oracle_lib.SolverXPBD, substeps=40, dt=1.0 / 240,
                   solver_kwargs={"iterations": 4}, record_contacts=True)
    assert dyn[2] == res["explicit"][2]


# ------------------------------------------------------------------------------------------------ contact matching, CPU
def _three_spheres():
    """test_contact_matching.py:29-44: spheres at x = -0.5, 0, 0.5 touching the ground plane."""
    b = ModelBuilder()
    b.add_ground_plane()
    for x in (-0.5, 0.0, 0.5):
        body = b.add_body(xform=X.transform((x, 0.0, 0.1)))
        b.add_shape_sphere(body, radius=0.1)
    return b.finalize()


class _OracleMatchingPipeline:
    """oracle.CollisionPipeline(deterministic=True) + oracle.contact_match.ContactMatcher, like CollisionPipeline(contact_matching=...)."""

    def __init__(self, oracle, model, **kw):
        from oracle.contact_match import ContactMatcher

        self.pipe = oracle.CollisionPipeline(model, broad_phase="nxn", deterministic=True)
        self.matcher = ContactMatcher(model, **kw)

    def contacts(self):
        return self.pipe.contacts()

    def collide(self, state, contacts):
        contacts.clear()
        self.pipe.collide(state, contacts)
        return self.matcher.match(contacts, state.body_q.numpy())


def test_matching_first_frame_identity_and_thresholds(oracle_lib):
    m = _three_spheres()
    pipe = _OracleMatchingPipeline(oracle_lib, m)
    c, s = pipe.contacts(), m.state()
    first = pipe.collide(s, c)
    assert len(first) == 3 and np.all(first == -1)  # :79-93 first frame: MATCH_NOT_FOUND everywhere
    for _ in range(2):  # :96-142 stable scene: identity across frames
        np.testing.assert_array_equal(pipe.collide(s, c), np.arange(3))
    # :287-340 every contact moved by more than pos_threshold (0.5 mm) -> MATCH_BROKEN; :343-374 a smaller move still matches
    s.body_q[:, 0] += 0.002
    np.testing.assert_array_equal(pipe.collide(s, c), [-2, -2, -2])
    s.body_q[:, 0] += 0.0002
    np.testing.assert_array_equal(pipe.collide(s, c), np.arange(3))
    # :244-284 a body that was out of reach comes back: its pair had no contacts last frame -> MATCH_NOT_FOUND, others keep theirs
    s.body_q[1, 2] = 5.0
    assert len(pipe.collide(s, c)) == 2
    s.body_q[1, 2] = 0.1
    np.testing.assert_array_equal(pipe.collide(s, c), [0, -1, 1])
    # masked reset (:165-198): only the selected world's contacts restart
    pipe.matcher.reset(np.array([True, False]))  # world 0 selected: nothing lives there (a model without begin_world() is all world -1)
    np.testing.assert_array_equal(pipe.collide(s, c), np.arange(3))
    pipe.matcher.reset(np.array([False, True]))  # the global slot: every contact restarts
    assert np.all(pipe.collide(s, c) == -1)


def test_matching_normal_threshold_and_box_manifold(oracle_lib):
    # :377-421 same position, normal turned by more than acos(0.995): broken
    b = ModelBuilder()
    b.add_ground_plane()
    body = b.add_body(xform=X.transform((0.0, 0.0, 0.099)))
    b.add_shape_box(body, hx=0.1, hy=0.1, hz=0.1)
    m = b.finalize()
    pipe = _OracleMatchingPipeline(oracle_lib, m)
    c, s = pipe.contacts(), m.state()
    assert np.all(pipe.collide(s, c) == -1)
    n = int(c.rigid_contact_count[0])
    assert n == 4  # :640-670 box on plane: four manifold contacts, identity match on the next frame
    np.testing.assert_array_equal(pipe.collide(s, c), np.arange(4))
    pipe.matcher.prev_normal[:] = np.array([np.sin(0.2), 0.0, np.cos(0.2)], dtype=np.float32)  # as if last frame's normal was tilted 11 deg
    np.testing.assert_array_equal(pipe.collide(s, c), [-2] * 4)


def _snapshot(c, n):
    return {f: getattr(c, f"rigid_contact_{f}")[:n].numpy().copy() for f in ("point0", "point1", "offset0", "offset1", "normal")}


def test_sticky_replays_matched_rows_and_passes_new_ones_through(oracle_lib):
    """test_contact_matching.py:704-773 - after a 0.1 mm shift (below the 0.5 mm threshold) the fresh narrow-phase record differs, the
    sticky record equals frame 1 byte for byte; :776-830 - a contact of a newly arrived shape is not overwritten."""
    b = ModelBuilder()
    b.add_ground_plane()
    for x in (-0.5, 0.5):
        b.add_shape_box(b.add_body(xform=X.transform((x, 0.0, 0.0995), X.quat_from_axis_angle((0.0, 0.0, 1.0), 0.3))), hx=0.1, hy=0.1, hz=0.1)
    late = b.add_shape_sphere(b.add_body(xform=X.transform((0.0, 0.0, 10.0))), radius=0.1)
    m = b.finalize()
    sticky, fresh = _OracleMatchingPipeline(oracle_lib, m, sticky=True), _OracleMatchingPipeline(oracle_lib, m)
    c, cf, s = sticky.contacts(), fresh.contacts(), m.state()
    assert np.all(sticky.collide(s, c) == -1)
    n1 = int(c.rigid_contact_count[0])
    assert n1 == 8
    first = _snapshot(c, n1)
    s.body_q[:, 0] += 0.0001
    s.body_q[:2, 2] -= 0.0001
    fresh.collide(s, cf)
    assert not np.array_equal(cf.rigid_contact_point1[:n1].numpy(), first["point1"])  # the fresh record really moved
    np.testing.assert_array_equal(sticky.collide(s, c), np.arange(n1))
    for f, want in first.items():
        np.testing.assert_array_equal(getattr(c, f"rigid_contact_{f}")[:n1].numpy(), want, err_msg=f)
    # a matched contact that has separated (fresh gap > 0) keeps the fresh record instead
    s.body_q[0, 2] += 0.0009  # 0.3 mm above the ground now: still inside the contact margin, midpoint moved by 0.45 mm
    match = sticky.collide(s, c)
    np.testing.assert_array_equal(match, np.arange(n1))
    lifted = _snapshot(c, n1)
    assert not np.array_equal(lifted["point0"][:4], first["point0"][:4]) or not np.array_equal(lifted["point1"][:4], first["point1"][:4])
    np.testing.assert_array_equal(lifted["point1"][4:], first["point1"][4:])
    # the parked sphere lands: its contact is new and carries its own shape
    s.body_q[2, :3] = torch.tensor([0.0, 0.0, 0.1])
    match = sticky.collide(s, c)
    n3 = int(c.rigid_contact_count[0])
    assert n3 == n1 + 1
    shape1 = c.rigid_contact_shape1[:n3].numpy()
    assert np.all(match[shape1 == late] < 0) and (shape1 == late).sum() == 1 and np.all(match[shape1 != late] >= 0)


def test_contact_report_lists_new_and_broken_rows(oracle_lib):
    """test_contact_matching.py:424-454, 493-528: first frame all new; stable frame neither new nor broken; a sphere that flies away
    leaves a broken row (index into the OLD sorted buffer); a masked reset silences the broken rows of the reset worlds."""
    m = _three_spheres()
    pipe = _OracleMatchingPipeline(oracle_lib, m)
    c, s = pipe.contacts(), m.state()
    pipe.collide(s, c)
    np.testing.assert_array_equal(pipe.matcher.new_indices, [0, 1, 2])
    assert len(pipe.matcher.broken_indices) == 0
    pipe.collide(s, c)
    assert len(pipe.matcher.new_indices) == 0 and len(pipe.matcher.broken_indices) == 0
    s.body_q[1, 2] = 10.0
    np.testing.assert_array_equal(pipe.collide(s, c), [0, 2])
    np.testing.assert_array_equal(pipe.matcher.broken_indices, [1])
    assert len(pipe.matcher.new_indices) == 0
    s.body_q[1, 2] = 0.1
    s.body_q[0, 2] = 10.0
    pipe.matcher.reset(np.array([False, True]))  # the global slot: every old row belongs to a reset world -> nothing is "broken"
    match = pipe.collide(s, c)
    assert np.all(match == -1) and len(pipe.matcher.broken_indices) == 0
    np.testing.assert_array_equal(pipe.matcher.new_indices, [0, 1])


# ------------------------------------------------------------------------------------------------ GPU
def _contacts_equal(cg, co, model):
    n_g, g = canonical_contacts(cg, model)
    n_o, o = canonical_contacts(co, model)
    assert n_g == n_o and n_o > 0
    for k in o:
        np.testing.assert_array_equal(g[k], o[k], err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["nxn", "sap"])
def test_gpu_runtime_broad_phase_matches_oracle(oracle_lib, cuda_lib, mode):
    """Model WITHOUT a precomputed pair list: candidates come from broadphase_kernel (filters + AABB test on the device).  Contact
    counts every substep, the final contact arrays and the final state equal the oracle's nxn / sap run bit for bit - and equal the
    explicit-list run of the full model."""
    bare = scenes.free_bodies_model(3, drop_pairs=True)
    kw = {"iterations": 4}
    ref = simulate(bare, lambda mm: oracle_lib.CollisionPipeline(mm, broad_phase=mode), oracle_lib.SolverXPBD, substeps=80, dt=1.0 / 240,
                   solver_kwargs=kw, record_contacts=True)
    got = simulate(bare.to("cuda:0"), lambda mm: newton_b200.CollisionPipeline(mm, broad_phase=mode), newton_b200.solvers.SolverXPBD, substeps=80,
                   dt=1.0 / 240, solver_kwargs=kw, record_contacts=True)
    assert got[2] == ref[2] and ref[2][-1] > 30
    _contacts_equal(got[1], ref[1], bare)
    np.testing.assert_array_equal(got[0].body_q.cpu().numpy(), ref[0].body_q.numpy())
    np.testing.assert_array_equal(got[0].body_qd.cpu().numpy(), ref[0].body_qd.numpy())
    full = scenes.free_bodies_model(3)
    exp = simulate(full.to("cuda:0"), newton_b200.CollisionPipeline, newton_b200.solvers.SolverXPBD, substeps=80, dt=1.0 / 240, solver_kwargs=kw,
                   record_contacts=True)
    assert exp[2] == got[2]
    assert torch.equal(exp[0].body_q, got[0].body_q)
    empty = simulate(bare.to("cuda:0"), newton_b200.CollisionPipeline, newton_b200.solvers.SolverXPBD, substeps=3, dt=1.0 / 240, record_contacts=True)
    assert empty[2] == [0, 0, 0]  # explicit mode really depends on the list


@pytest.mark.gpu
def test_gpu_runtime_broad_phase_on_benchmark_scenes(oracle_lib, cuda_lib):
    """quadrupeds (13 plane-cylinder pairs per world, everything else filtered) and box stacks: sap == explicit on the GPU."""
    for model, dt, n in ((scenes.quadruped_model(8, seed=1), 0.005, 40), (scenes.box_stack_model(4, seed=0), 1.0 / 240, 40)):
        mg = model.to("cuda:0")
        a = simulate(mg, newton_b200.CollisionPipeline, newton_b200.solvers.SolverXPBD, substeps=n, dt=dt, solver_kwargs={"iterations": 4},
                     record_contacts=True)
        mg2 = model.to("cuda:0")
        b = simulate(mg2, lambda mm: newton_b200.CollisionPipeline(mm, broad_phase="sap"), newton_b200.solvers.SolverXPBD, substeps=n, dt=dt,
                     solver_kwargs={"iterations": 4}, record_contacts=True)
        assert a[2] == b[2]
        assert torch.equal(a[0].body_q, b[0].body_q) and torch.equal(a[0].body_qd, b[0].body_qd)


@pytest.mark.gpu
def test_gpu_include_static_kinematic_pairs(oracle_lib, cuda_lib):
    """A kinematic body resting on a static box: the pair exists by default and is pruned with include_static_kinematic_pairs=False
    (both broad-phase flavours), while pairs with a dynamic body stay."""
    b = ModelBuilder()
    b.add_ground_plane()
    b.add_shape_box(-1, xform=X.transform((0.0, 0.0, 0.1)), hx=0.5, hy=0.5, hz=0.1)
    kin = b.add_body(xform=X.transform((0.0, 0.0, 0.29)), is_kinematic=True)
    b.add_shape_box(kin, hx=0.1, hy=0.1, hz=0.1)
    dyn = b.add_body(xform=X.transform((0.3, 0.0, 0.29)))
    b.add_shape_sphere(dyn, radius=0.1)
    m = b.finalize()
    mg = m.to("cuda:0")
    for mode in ("explicit", "nxn"):
        counts = {}
        for flag in (True, False):
            pg = newton_b200.CollisionPipeline(mg, broad_phase=mode, include_static_kinematic_pairs=flag)
            po = oracle_lib.CollisionPipeline(m, broad_phase=mode, include_static_kinematic_pairs=flag)
            cg, co = pg.contacts(), po.contacts()
            pg.collide(mg.state(), cg)
            po.collide(m.state(), co)
            counts[flag] = int(cg.rigid_contact_count.item())
            assert counts[flag] == int(co.rigid_contact_count[0])
        assert counts[True] > counts[False] > 0


@pytest.mark.gpu
def test_gpu_contact_matching_matches_oracle(oracle_lib, cuda_lib):
    """CollisionPipeline(contact_matching="latest") on a moving heap: match_index of every frame equals the oracle matcher's."""
    from oracle.contact_match import ContactMatcher

    m = scenes.free_bodies_model(2)
    mg = m.to("cuda:0")
    pg = newton_b200.CollisionPipeline(mg, broad_phase="nxn", contact_matching="latest", contact_matching_pos_threshold=0.002)
    assert pg.deterministic
    po = oracle_lib.CollisionPipeline(m, broad_phase="nxn", deterministic=True)
    matcher = ContactMatcher(m, pos_threshold=0.002)
    sg, so = newton_b200.solvers.SolverXPBD(mg, iterations=4), oracle_lib.SolverXPBD(m, iterations=4)
    g0, g1, o0, o1 = mg.state(), mg.state(), m.state(), m.state()
    cg, co = pg.contacts(), po.contacts()
    seen = set()
    for frame in range(30):
        pg.collide(g0, cg)
        po.collide(o0, co)
        want = matcher.match(co, o0.body_q.numpy())
        n = int(cg.rigid_contact_count.item())
        assert n == len(want)
        got = cg.rigid_contact_match_index[:n].cpu().numpy()
        np.testing.assert_array_equal(got, want, err_msg=f"frame {frame}")
        seen.update(np.unique(want).tolist())
        if frame == 14:  # masked reset of world 1 through the public API
            mask = torch.zeros(3, dtype=torch.bool, device="cuda:0")
            mask[1] = True
            pg.reset(mask)
            matcher.reset(np.array([False, True, False]))
        for _ in range(3):
            g0.clear_forces()
            sg.step(g0, g1, None, cg, 1.0 / 240)
            g0, g1 = g1, g0
            o0.clear_forces()
            so.step(o0, o1, None, co, 1.0 / 240)
            o0, o1 = o1, o0
    assert -1 in seen and -2 in seen and max(seen) > 5  # the run saw new, broken and matched contacts


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["latest", "sticky"])
def test_gpu_contact_report_and_sticky_match_oracle(oracle_lib, cuda_lib, mode):
    """contact_report=True with "latest" / "sticky" on the moving heap: match_index, the new / broken lists and (sticky) the replayed
    Contacts arrays of every frame equal the oracle matcher's; the solver consumes the replayed buffer (re-imported), so the states stay
    bit-identical too."""
    from oracle.contact_match import ContactMatcher

    m = scenes.free_bodies_model(2)
    mg = m.to("cuda:0")
    pg = newton_b200.CollisionPipeline(mg, broad_phase="nxn", contact_matching=mode, contact_report=True, contact_matching_pos_threshold=0.002)
    po = oracle_lib.CollisionPipeline(m, broad_phase="nxn", deterministic=True)
    matcher = ContactMatcher(m, pos_threshold=0.002, sticky=mode == "sticky")
    sg, so = newton_b200.solvers.SolverXPBD(mg, iterations=4), oracle_lib.SolverXPBD(m, iterations=4)
    g0, g1, o0, o1 = mg.state(), mg.state(), m.state(), m.state()
    cg, co = pg.contacts(), po.contacts()
    assert cg.contact_matching_mode == mode and cg.rigid_contact_new_indices is not None
    replayed = broken_seen = 0
    for frame in range(30):
        pg.collide(g0, cg)
        po.collide(o0, co)
        fresh_point1 = co.rigid_contact_point1.numpy().copy()
        want = matcher.match(co, o0.body_q.numpy())
        n = int(cg.rigid_contact_count.item())
        assert n == len(want)
        np.testing.assert_array_equal(cg.rigid_contact_match_index[:n].cpu().numpy(), want, err_msg=f"frame {frame}")
        nn, nb = int(cg.rigid_contact_new_count.item()), int(cg.rigid_contact_broken_count.item())
        np.testing.assert_array_equal(cg.rigid_contact_new_indices[:nn].cpu().numpy(), matcher.new_indices, err_msg=f"new, frame {frame}")
        np.testing.assert_array_equal(cg.rigid_contact_broken_indices[:nb].cpu().numpy(), matcher.broken_indices, err_msg=f"broken, frame {frame}")
        broken_seen += nb
        for f in ("point0", "point1", "offset0", "offset1", "normal", "margin0", "margin1"):
            np.testing.assert_array_equal(getattr(cg, f"rigid_contact_{f}")[:n].cpu().numpy(), getattr(co, f"rigid_contact_{f}")[:n].numpy(),
                                          err_msg=f"{f}, frame {frame}")
        replayed += int((fresh_point1[:n] != co.rigid_contact_point1[:n].numpy()).any(axis=1).sum())
        if frame == 14:
            mask = torch.zeros(3, dtype=torch.bool, device="cuda:0")
            mask[1] = True
            pg.reset(mask)
            matcher.reset(np.array([False, True, False]))
        for _ in range(3):
            g0.clear_forces()
            sg.step(g0, g1, None, cg, 1.0 / 240)
            g0, g1 = g1, g0
            o0.clear_forces()
            so.step(o0, o1, None, co, 1.0 / 240)
            o0, o1 = o1, o0
        np.testing.assert_array_equal(g0.body_q.cpu().numpy(), o0.body_q.numpy(), err_msg=f"state, frame {frame}")
    assert broken_seen > 0
    assert (replayed > 0) == (mode == "sticky")

"""The oracle's own restatement of the convex path (oracle/oracle_convex.h, written from the reference) against the
host compilation of the product's routine (newton_b200/csrc/nb2_convex.cuh), bit for bit, on randomised inputs.

Both are CPU code (g++ -ffp-contract=off); the GPU parity tests then compare the CUDA compilation with the oracle's
restatement, so an error in the product routine would have to be made identically in two independently written
translations to go unnoticed.  The restatement itself is pinned by the reference's known answers
(tests/test_oracle_known_answers.py)."""

import itertools

import numpy as np
import pytest

T = dict(plane=1, sphere=3, capsule=4, ellipsoid=5, cylinder=6, box=7, cone=9)
SOLID = {k: v for k, v in T.items() if k != "plane"}


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _rand_scale(rng, t):
    if t == T["plane"]:  # infinite (0, 0) or finite
        return (rng.choice([0.0, 1.0]) * rng.uniform(0.5, 3.0, 3) * np.array([1, 1, 0])).astype(np.float32)
    if t == T["cylinder"] and rng.random() < 0.3:  # barrel: scale.z = barrel radius > half height
        r, h = rng.uniform(0.1, 0.6), rng.uniform(0.1, 0.6)
        return np.array([r, h, max(h, r) * rng.uniform(1.2, 3.0)], np.float32)
    s = rng.uniform(0.1, 0.7, 3).astype(np.float32)
    if t == T["cylinder"]:
        s[2] = 0
    return s


def _rand_quat(rng):
    if rng.random() < 0.25:  # quarter turns about a coordinate axis: exercises the tie-breaking paths
        k, ax = rng.integers(0, 4), rng.integers(0, 3)
        q = np.zeros(4)
        q[ax], q[3] = np.sin(k * np.pi / 4), np.cos(k * np.pi / 4)
        return q.astype(np.float32)
    q = rng.normal(size=4)
    return (q / np.linalg.norm(q)).astype(np.float32)


def _rand_pose_pair(rng):
    pa = rng.uniform(-0.2, 0.2, 3).astype(np.float32)
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    pb = pa + d * rng.uniform(0.0, 1.6)
    if rng.random() < 0.2:  # stacked along z
        pb = pa + np.array([0, 0, rng.uniform(0.1, 1.0)])
    return (np.concatenate([pa, _rand_quat(rng)]).astype(np.float32), np.concatenate([pb, _rand_quat(rng)]).astype(np.float32))


@pytest.mark.parametrize("name_a,name_b", list(itertools.product(T, T)))
def test_pair_contacts_bit_equal(oracle_lib, name_a, name_b):
    import oracle

    rng = np.random.default_rng(1000 * T[name_a] + T[name_b])
    ta, tb = T[name_a], T[name_b]
    produced = 0
    for _ in range(60):
        sa, sb = _rand_scale(rng, ta), _rand_scale(rng, tb)
        xa, xb = _rand_pose_pair(rng)
        gap = float(rng.choice([0.0, 0.02, 0.2]))
        ma = float(rng.choice([0.0, 0.0, 5e-5, 0.01]))
        mb = float(rng.choice([0.0, 0.0, 0.01]))
        r0 = oracle.convex_pair(ta, sa, xa, tb, sb, xb, gap, "oracle", ma, mb)
        r1 = oracle.convex_pair(ta, sa, xa, tb, sb, xb, gap, "product_host", ma, mb)
        assert r0[0] == r1[0], (sa, sb, xa, xb, gap, ma, mb)
        for a, b in zip(r0[1:], r1[1:]):
            assert np.array_equal(_bits(a), _bits(b)), (sa, sb, xa, xb, gap, ma, mb)
        produced += r0[0]
    if not (ta == T["plane"] and tb == T["plane"]):
        assert produced > 0  # the sample must actually reach the contact-producing branches


@pytest.mark.parametrize("name_a,name_b", list(itertools.product(T, T)))
def test_pair_contacts_with_the_speculative_writer_bit_equal(oracle_lib, name_a, name_b):
    """write_contact_speculative (predictive admission, search gap for the pair, plane-proxy overlap radius) in both translations"""
    import oracle

    rng = np.random.default_rng(31 + 1000 * T[name_a] + T[name_b])
    ta, tb = T[name_a], T[name_b]
    produced = extra = 0
    for _ in range(40):
        sa, sb = _rand_scale(rng, ta), _rand_scale(rng, tb)
        xa, xb = _rand_pose_pair(rng)
        base = float(rng.choice([0.0, 0.02]))
        search = base + float(rng.uniform(0.0, 0.6))
        dt, ext = float(rng.choice([0.0, 0.01, 0.05])), float(rng.choice([0.1, 0.5]))
        approach = (xb[:3] - xa[:3]) * rng.uniform(-2.0, 12.0)  # A moving towards (or away from) B
        la, lb = approach.astype(np.float32), rng.normal(size=3).astype(np.float32)
        wa, wb = rng.normal(size=3).astype(np.float32) * 2.0, rng.normal(size=3).astype(np.float32) * 2.0
        r0 = oracle.convex_pair_speculative(ta, sa, xa, tb, sb, xb, search, base, dt, ext, la, wa, lb, wb, impl="oracle")
        r1 = oracle.convex_pair_speculative(ta, sa, xa, tb, sb, xb, search, base, dt, ext, la, wa, lb, wb, impl="product_host")
        assert r0[0] == r1[0], (sa, sb, xa, xb, base, search, dt, ext)
        for a, b in zip(r0[1:], r1[1:]):
            assert np.array_equal(_bits(a), _bits(b)), (sa, sb, xa, xb, base, search, dt, ext)
        produced += r0[0]
        extra += r0[0] - oracle.convex_pair(ta, sa, xa, tb, sb, xb, base, "oracle")[0]
    if not (ta == T["plane"] and tb == T["plane"]):
        assert produced > 0


@pytest.mark.parametrize("name_a,name_b", list(itertools.product(SOLID, SOLID)))
def test_mpr_and_gjk_cores_bit_equal(oracle_lib, name_a, name_b):
    import oracle

    rng = np.random.default_rng(77 + 1000 * T[name_a] + T[name_b])
    ta, tb = T[name_a], T[name_b]
    hits = seps = 0
    for _ in range(40):
        sa, sb = _rand_scale(rng, ta), _rand_scale(rng, tb)
        xa, xb = _rand_pose_pair(rng)
        pos_b, quat_b = xb[:3] - xa[:3], xb[3:]
        extend = float(rng.choice([0.0, 1e-4, 2e-4]))
        m0 = oracle.mpr_core(ta, sa, tb, sb, pos_b, quat_b, extend, impl="oracle")
        m1 = oracle.mpr_core(ta, sa, tb, sb, pos_b, quat_b, extend, impl="product_host")
        assert m0[0] == m1[0]
        if m0[0]:  # the outputs of a miss are scratch values in both
            hits += 1
            assert all(np.array_equal(_bits(a), _bits(b)) for a, b in zip(m0[1:4], m1[1:4])) and _bits(m0[4]) == _bits(m1[4])
        g0 = oracle.gjk_core(ta, sa, tb, sb, pos_b, quat_b, 0.0, impl="oracle")
        g1 = oracle.gjk_core(ta, sa, tb, sb, pos_b, quat_b, 0.0, impl="product_host")
        assert g0[0] == g1[0]
        seps += g0[0]
        assert all(np.array_equal(_bits(a), _bits(b)) for a, b in zip(g0[1:4], g1[1:4])) and _bits(g0[4]) == _bits(g1[4])
    assert hits > 0 and seps > 0


@pytest.mark.parametrize("name", list(T))
def test_support_map_bit_equal(oracle_lib, name):
    import oracle

    rng = np.random.default_rng(5 + T[name])
    for i in range(200):
        s = _rand_scale(rng, T[name])
        d = rng.normal(size=3).astype(np.float32)
        if i % 5 == 0:
            d[rng.integers(0, 3)] = 0.0  # axis-aligned / planar directions
        if i % 17 == 0:
            d[:] = 0.0
        assert np.array_equal(_bits(oracle.support_map(T[name], s, d, "oracle")), _bits(oracle.support_map(T[name], s, d, "product_host")))


@pytest.mark.parametrize("name", list(T))
def test_tight_aabb_bit_equal(oracle_lib, name):
    import oracle

    rng = np.random.default_rng(9 + T[name])
    for _ in range(100):
        s = _rand_scale(rng, T[name])
        x = np.concatenate([rng.uniform(-2, 2, 3), _rand_quat(rng)]).astype(np.float32)
        a, b = oracle.tight_aabb(T[name], s, x, "oracle"), oracle.tight_aabb(T[name], s, x, "product_host")
        assert np.array_equal(_bits(a[0]), _bits(b[0])) and np.array_equal(_bits(a[1]), _bits(b[1]))


@pytest.mark.parametrize("name_a,name_b", list(itertools.product(T, T)))
def test_degenerate_configurations_bit_equal(oracle_lib, name_a, name_b):
    """Configurations a random sweep rarely produces: coincident centres (MPR's interior-point fallbacks), faces exactly touching or
    1e-6 / 1e-4 apart, axis-aligned and quarter-turn orientations, millimetre- and hundred-metre-sized shapes, the three margin
    regimes of the MPR inflation.  Also: whatever comes out is finite."""
    import oracle

    rng = np.random.default_rng(31 + 1000 * T[name_a] + T[name_b])
    ta, tb = T[name_a], T[name_b]
    axes = np.eye(3)

    def scale(t, mode):
        if t == T["plane"]:
            return (rng.choice([0.0, 1.0]) * np.array([rng.choice([0.5, 2.0]), rng.choice([0.5, 2.0]), 0.0])).astype(np.float32)
        base = {0: rng.choice([0.25, 0.5, 1.0], size=3), 1: rng.uniform(1e-3, 1e-2, 3), 2: rng.uniform(10.0, 100.0, 3)}[mode].astype(np.float32)
        if t == T["cylinder"]:
            base[2] = 0.0 if rng.random() < 0.7 else max(base[0], base[1]) * 1.5
        return base

    def quat(mode):
        if mode == 0:
            return np.array([0, 0, 0, 1], dtype=np.float32)
        if mode == 1:
            k, ax = rng.integers(0, 4), axes[rng.integers(0, 3)]
            return np.array([*(ax * np.sin(k * np.pi / 4)), np.cos(k * np.pi / 4)], dtype=np.float32)
        return _rand_quat(rng)

    for _ in range(40):
        mode = int(rng.integers(0, 3))
        sa, sb = scale(ta, mode), scale(tb, mode)
        size = float(max(np.abs(sa).max(), np.abs(sb).max(), 1e-3))
        reach = float(sa.max() + sb.max())
        kind = int(rng.integers(0, 4))
        if kind == 0:
            off = np.zeros(3)
        elif kind == 1:
            off = axes[rng.integers(0, 3)] * reach
        elif kind == 2:
            off = axes[rng.integers(0, 3)] * reach * (1.0 + rng.choice([-1e-6, 1e-6, -1e-4, 1e-4]))
        else:
            off = np.array([0.0, 0.0, 1.0]) * size * rng.uniform(0.5, 2.5)
        xa = np.concatenate([np.zeros(3), quat(int(rng.integers(0, 3)))]).astype(np.float32)
        xb = np.concatenate([off, quat(int(rng.integers(0, 3)))]).astype(np.float32)
        gap = float(rng.choice([0.0, 0.01, 0.2])) * (size if mode else 1.0)
        ma, mb = float(rng.choice([0.0, 2.5e-5, 0.005])), float(rng.choice([0.0, 2.5e-5, 0.005]))
        r0 = oracle.convex_pair(ta, sa, xa, tb, sb, xb, gap, "oracle", ma, mb)
        r1 = oracle.convex_pair(ta, sa, xa, tb, sb, xb, gap, "product_host", ma, mb)
        assert r0[0] == r1[0], (sa, sb, xa, xb, gap, ma, mb)
        for a, b in zip(r0[1:], r1[1:]):
            assert np.array_equal(_bits(a), _bits(b)), (sa, sb, xa, xb, gap, ma, mb)
        assert np.isfinite(r0[1][: r0[0]]).all() and np.isfinite(r0[2][: r0[0]]).all() and np.isfinite(r0[3][: r0[0]]).all()

// nb2_convex.cuh - convex-convex contact generation: support maps, MPR (XenoCollide) penetration query,
// GJK closest-distance query and the <=5-point contact manifold.  One lane processes one shape pair.
//
// Restates the reference's generic convex path:
//   support_map / box support dead-band / MPR tie centring   geometry/support_function.py:121-431
//   solve_mpr_core                                           geometry/mpr.py:188-403  (zlib-licensed XenoCollide, altered upstream)
//   solve_closest_distance_core (GJK, Johnson sub-algorithm) geometry/simplex_solver.py:44-470
//   solve_convex_multi_contact                               geometry/collision_convex.py:107-231
//   build_manifold / polygon clipping / rotating calipers    geometry/multicontact.py:27-958
//   post_process_axial_on_discrete_contact                   geometry/collision_core.py:174-277
//
// HOST + DEVICE: this header compiles for the device (libnewton_b200.so) and, with g++ -ffp-contract=off, for the
// host.  The host compilation exists only for tests/test_oracle_convex_independence.py, which compares it bit for bit with
// the oracle's own, separately written restatement (oracle/oracle_convex.h); nothing in the product runs it on the CPU.
#pragma once
#include "nb2_math.cuh"

namespace nb2 {

enum { CG_PLANE = 1, CG_SPHERE = 3, CG_CAPSULE = 4, CG_ELLIPSOID = 5, CG_CYLINDER = 6, CG_BOX = 7, CG_CONE = 9, CG_CONVEX_MESH = 10 };

// GenericShapeData (support_function.py:84-118).  `center` is the Minkowski-centre seed: the local origin for every primitive,
// the centre of the scaled local AABB for a convex hull (narrow_phase.py:1102-1105).  `hull` / `hull_count` are the hull's
// UNSCALED vertices (wp.Mesh.points of shape_source; the per-shape scale is applied by the support map) - null for primitives.
struct ConvexGeom {
    int type;
    V3 scale;
    V3 center;
    const float* hull;
    int hull_count;
    NB2_DEV ConvexGeom() : type(0), hull(nullptr), hull_count(0) {}
    NB2_DEV ConvexGeom(int t, V3 s) : type(t), scale(s), hull(nullptr), hull_count(0) {}
    NB2_DEV ConvexGeom(int t, V3 s, V3 c, const float* h, int n) : type(t), scale(s), center(c), hull(h), hull_count(n) {}
};

NB2_DEV float rsqrt_rn(float v) { return 1.0f / sqrtf(v); }  // _support_rsqrt_rn, host flavour (support_function.py:44-53)

NB2_DEV V3 support_box(V3 scale, V3 dir) {  // support_function.py:121-129
    const float ds = fmax_w(fabsf(dir.x), fmax_w(fabsf(dir.y), fabsf(dir.z)));
    const float th = 1.0e-10f * ds;
    return V3((dir.x >= -th ? 1.0f : -1.0f) * scale.x, (dir.y >= -th ? 1.0f : -1.0f) * scale.y, (dir.z >= -th ? 1.0f : -1.0f) * scale.z);
}

NB2_CALL V3 support_map(const ConvexGeom& g, V3 d) {  // support_function.py:133-352
    if (g.type == CG_CONVEX_MESH) {  // vertex scan, first maximum wins (support_function.py:153-172)
        const V3 sd = cmul(d, g.scale);  // dot(scale * v, d) == dot(v, scale * d): one scaling per query instead of per vertex
        float max_dot = -1.0e10f;
        int best = 0;
        for (int i = 0; i < g.hull_count; ++i) {
            const float dv = dot(V3(g.hull[3 * i], g.hull[3 * i + 1], g.hull[3 * i + 2]), sd);
            if (dv > max_dot) {
                max_dot = dv;
                best = i;
            }
        }
        if (g.hull_count == 0) return V3();
        return cmul(V3(g.hull[3 * best], g.hull[3 * best + 1], g.hull[3 * best + 2]), g.scale);
    }
    const float eps = 1.0e-12f;
    if (g.type == CG_BOX) return support_box(g.scale, d);
    if (g.type == CG_SPHERE) {
        const float l2 = len2(d);
        const V3 n = l2 > eps ? d * rsqrt_rn(l2) : V3(1.f, 0.f, 0.f);
        return n * g.scale.x;
    }
    if (g.type == CG_CAPSULE) {
        const float l2 = len2(d);
        const V3 n = l2 > eps ? d * rsqrt_rn(l2) : V3(1.f, 0.f, 0.f);
        V3 r = n * g.scale.x;
        return r + V3(0.f, 0.f, d.z >= 0.0f ? g.scale.y : -g.scale.y);
    }
    if (g.type == CG_ELLIPSOID) {
        const float a = g.scale.x, b = g.scale.y, c = g.scale.z;
        if (len2(d) > eps) {
            const float adx = a * d.x, bdy = b * d.y, cdz = c * d.z;
            const float den = adx * adx + bdy * bdy + cdz * cdz;
            if (den > eps) {
                const float inv = rsqrt_rn(den);
                return V3((a * a) * d.x * inv, (b * b) * d.y * inv, (c * c) * d.z * inv);
            }
        }
        return V3(a, 0.f, 0.f);
    }
    if (g.type == CG_CYLINDER) {
        const float radius = g.scale.x, hh = g.scale.y, br = g.scale.z;
        const V3 dxy(d.x, d.y, 0.f);
        const float l2 = len2(dxy);
        if (br == 0.0f) {
            V3 lat(radius, 0.f, 0.f);
            if (l2 > eps) {
                const V3 n = dxy * rsqrt_rn(l2);
                lat = V3(n.x * radius, n.y * radius, 0.f);
            }
            if (d.z > 0.0f) return V3(lat.x, lat.y, hh);
            if (d.z < 0.0f) return V3(lat.x, lat.y, -hh);
            return lat;
        }
        V3 nxy(1.f, 0.f, 0.f);
        if (l2 > eps) nxy = dxy / sqrtf(l2);
        const float dl = sqrtf(l2 + d.z * d.z);
        float sz = 0.0f;
        if (dl > eps) sz = clamp_w(br * d.z / dl, -hh, hh);
        const float br2 = br * br, hh2 = hh * hh, sz2 = sz * sz;
        const float end_off = sqrtf(br2 - hh2), sup_off = sqrtf(fmax_w(br2 - sz2, 0.0f));
        const float off_sum = sup_off + end_off;
        float sr = radius;
        if (off_sum > eps) sr += (hh2 - sz2) / off_sum;
        return V3(nxy.x * sr, nxy.y * sr, sz);
    }
    if (g.type == CG_CONE) {
        const float radius = g.scale.x, hh = g.scale.y;
        const V3 apex(0.f, 0.f, hh);
        const V3 dxy(d.x, d.y, 0.f);
        const float l = len(dxy);
        const float k = hh > eps ? radius / (2.0f * hh) : 0.0f;
        if (l <= eps) return d.z >= 0.0f ? apex : V3(radius, 0.f, -hh);
        if (d.z >= k * l) return apex;
        const V3 n = dxy / l;
        return V3(n.x * radius, n.y * radius, -hh);
    }
    if (g.type == CG_PLANE) return V3((d.x >= 0.0f ? 1.0f : -1.0f) * g.scale.x, (d.y >= 0.0f ? 1.0f : -1.0f) * g.scale.y, 0.f);
    return V3();
}

// shape_support with center_ties=True (support_function.py:397-431): MPR only
NB2_CALL V3 support_ties(const ConvexGeom& g, V3 d) {
    if (g.type != CG_BOX) return support_map(g, d);
    V3 r = support_box(g.scale, d);
    const V3 c = cmul(vabs(d), g.scale);
    const float th = 1.0e-6f * (c.x + c.y + c.z);
    if (c.x <= th) r.x = 0.0f;
    if (c.y <= th) r.y = 0.0f;
    if (c.z <= th) r.z = 0.0f;
    return r;
}

struct MVert {
    V3 B, BtoA;
};
NB2_DEV V3 mvert_a(const MVert& v) { return v.B + v.BtoA; }

// minkowski_support (mpr.py:110-150); TIES selects the MPR flavour of the shape support
template <bool TIES>
NB2_CALL MVert mink_support(const ConvexGeom& ga, const ConvexGeom& gb, V3 dir, Q4 qb, V3 pb, float extend) {
    MVert v;
    V3 pa = TIES ? support_ties(ga, dir) : support_map(ga, dir);
    const V3 nd = -dir;
    const V3 tmp = qrot_inv(qb, nd);
    V3 r = TIES ? support_ties(gb, tmp) : support_map(gb, tmp);
    r = qrot(qb, r);
    v.B = r + pb;
    if (extend != 0.0f) {
        const V3 e = unit(dir) * extend * 0.5f;
        pa = pa + e;
        v.B = v.B - e;
    }
    v.BtoA = pa - v.B;
    return v;
}

// solve_mpr_core (mpr.py:188-403).  Frame: shape A at the origin, B at (qb, pb).
NB2_DEV bool mpr_core(const ConvexGeom& ga, const ConvexGeom& gb, Q4 qb, V3 pb, float extend, V3& point_a, V3& point_b, V3& normal,
                      float& penetration) {
    const int MAX_ITER = 30;
    const float COLLIDE_EPSILON = 1e-5f, NUM_EPS = 1e-16f;
    penetration = 0.0f;
    point_a = V3();
    point_b = V3();
    MVert v0;
    v0.B = pb + qrot(qb, gb.center);  // shape_center (support_function.py:563-594): zero for primitives, AABB centre for hulls
    v0.BtoA = ga.center - v0.B;
    normal = v0.BtoA;
    if (len2(normal) < NUM_EPS) {
        v0.BtoA = V3();  // fallback() of non-triangle shapes
        float best = -1.0e30f;
        V3 best_dir(1.f, 0.f, 0.f);
        for (int ax = 0; ax < 3; ++ax) {
            V3 probe;
            probe.set(ax, 1.0f);
            MVert sv = mink_support<true>(ga, gb, probe, qb, pb, extend);
            float dd = dot(sv.BtoA, probe);
            if (dd > best) {
                best = dd;
                best_dir = probe;
            }
        }
        v0.BtoA = best_dir * 1e-05f;
    }
    normal = -v0.BtoA;
    MVert v1 = mink_support<true>(ga, gb, normal, qb, pb, extend);
    point_a = mvert_a(v1);
    point_b = v1.B;
    if (dot(v1.BtoA, normal) <= 0.0f) return false;
    normal = cross(v1.BtoA, v0.BtoA);
    if (len2(normal) < NUM_EPS * NUM_EPS) {
        normal = unit(v1.BtoA - v0.BtoA);
        penetration = dot(v1.BtoA, normal);
        return true;
    }
    MVert v2 = mink_support<true>(ga, gb, normal, qb, pb, extend);
    if (dot(v2.BtoA, normal) <= 0.0f) return false;
    V3 t1 = v1.BtoA - v0.BtoA, t2 = v2.BtoA - v0.BtoA;
    normal = cross(t1, t2);
    if (dot(normal, v0.BtoA) > 0.0f) {
        MVert tmp = v1;
        v1 = v2;
        v2 = tmp;
        normal = -normal;
    }
    int phase1 = 0, phase2 = 0;
    bool hit = false;
    MVert v3;
    while (true) {
        if (phase1 > MAX_ITER) return false;
        phase1 += 1;
        v3 = mink_support<true>(ga, gb, normal, qb, pb, extend);
        if (dot(v3.BtoA, normal) <= 0.0f) return false;
        t1 = cross(v1.BtoA, v3.BtoA);
        if (dot(t1, v0.BtoA) < 0.0f) {
            v2 = v3;
            t1 = v1.BtoA - v0.BtoA;
            t2 = v3.BtoA - v0.BtoA;
            normal = cross(t1, t2);
            continue;
        }
        t1 = cross(v3.BtoA, v2.BtoA);
        if (dot(t1, v0.BtoA) < 0.0f) {
            v1 = v3;
            t1 = v3.BtoA - v0.BtoA;
            t2 = v2.BtoA - v0.BtoA;
            normal = cross(t1, t2);
            continue;
        }
        break;
    }
    while (true) {
        phase2 += 1;
        t1 = v2.BtoA - v1.BtoA;
        t2 = v3.BtoA - v1.BtoA;
        normal = cross(t1, t2);
        const float nsq = len2(normal);
        if (nsq < NUM_EPS * NUM_EPS) return false;
        if (!hit) hit = dot(normal, v1.BtoA) >= 0.0f;
        MVert v4 = mink_support<true>(ga, gb, normal, qb, pb, extend);
        V3 t3 = v4.BtoA - v3.BtoA;
        const float delta = dot(t3, normal);
        penetration = dot(v4.BtoA, normal);
        if (delta * delta <= COLLIDE_EPSILON * COLLIDE_EPSILON * nsq || penetration <= 0.0f || phase2 > MAX_ITER) {
            if (hit) {
                const float inv_n = 1.0f / sqrtf(nsq);
                penetration *= inv_n;
                normal = normal * inv_n;
                t3 = cross(v1.BtoA, t1);
                const float gamma = dot(t3, normal) * inv_n;
                t3 = cross(t2, v1.BtoA);
                const float beta = dot(t3, normal) * inv_n;
                const float alpha = 1.0f - gamma - beta;
                point_a = alpha * mvert_a(v1) + beta * mvert_a(v2) + gamma * mvert_a(v3);
                point_b = alpha * v1.B + beta * v2.B + gamma * v3.B;
            }
            return hit;
        }
        t1 = cross(v4.BtoA, v0.BtoA);
        float dd = dot(t1, v1.BtoA);
        if (dd >= 0.0f) {
            dd = dot(t1, v2.BtoA);
            if (dd >= 0.0f) v1 = v4;
            else v3 = v4;
        } else {
            dd = dot(t1, v3.BtoA);
            if (dd >= 0.0f) v2 = v4;
            else v1 = v4;
        }
    }
}

// ---- GJK closest distance (simplex_solver.py) -----------------------------------------------------------
struct Simplex {
    V3 B[4], D[4];  // D = BtoA (Minkowski difference vertices)
    float bc[4];
    unsigned mask;
};
NB2_DEV void bc_clear(float* bc) { bc[0] = bc[1] = bc[2] = bc[3] = 0.0f; }

NB2_CALL V3 closest_segment(const Simplex& s, int i0, int i1, float* bc, unsigned& mask) {  // :104-150
    const float EPS = 1e-8f;
    const V3 a = s.D[i0], b = s.D[i1];
    const V3 edge = b - a;
    const float vsq = len2(edge);
    const bool degenerate = vsq < EPS;
    const float t = -dot(a, edge) / (degenerate ? EPS : vsq);
    float l0 = 1.0f - t, l1 = t;
    mask = (1u << i0) | (1u << i1);
    bc_clear(bc);
    if (l0 < 0.0f || degenerate) {
        mask = 1u << i1;
        l0 = 0.0f;
        l1 = 1.0f;
    } else if (l1 < 0.0f) {
        mask = 1u << i0;
        l0 = 1.0f;
        l1 = 0.0f;
    }
    bc[i0] = l0;
    bc[i1] = l1;
    return l0 * a + l1 * b;
}

NB2_CALL V3 closest_triangle(const Simplex& s, int i0, int i1, int i2, float* bc, unsigned& mask) {  // :152-230
    const float EPS = 1e-8f;
    const V3 a = s.D[i0], b = s.D[i1], c = s.D[i2];
    const V3 u = a - b, w = a - c;
    const V3 n = cross(u, w);
    const float t = len2(n);
    const bool degenerate = t < EPS;
    const float it = 1.0f / (degenerate ? EPS : t);
    const V3 c1 = cross(u, a), c2 = cross(a, w);
    const float l2 = dot(c1, n) * it, l1 = dot(c2, n) * it;
    const float l0 = 1.0f - l2 - l1;
    float best = 1e30f;
    V3 cp;
    bc_clear(bc);
    mask = 0u;
    float tb[4];
    unsigned tm;
    if (l0 < 0.0f || degenerate) {
        V3 cl = closest_segment(s, i1, i2, tb, tm);
        float d2 = len2(cl);
        if (d2 < best) {
            for (int k = 0; k < 4; ++k) bc[k] = tb[k];
            mask = tm;
            best = d2;
            cp = cl;
        }
    }
    if (l1 < 0.0f || degenerate) {
        V3 cl = closest_segment(s, i0, i2, tb, tm);
        float d2 = len2(cl);
        if (d2 < best) {
            for (int k = 0; k < 4; ++k) bc[k] = tb[k];
            mask = tm;
            best = d2;
            cp = cl;
        }
    }
    if (l2 < 0.0f || degenerate) {
        V3 cl = closest_segment(s, i0, i1, tb, tm);
        float d2 = len2(cl);
        if (d2 < best) {
            for (int k = 0; k < 4; ++k) bc[k] = tb[k];
            mask = tm;
            cp = cl;
        }
    }
    if (mask != 0u) return cp;
    bc[i0] = l0;
    bc[i1] = l1;
    bc[i2] = l2;
    mask = (1u << i0) | (1u << i1) | (1u << i2);
    return l0 * a + l1 * b + l2 * c;
}

NB2_DEV float det4(V3 a, V3 b, V3 c, V3 d) { return dot(b - a, cross(c - a, d - a)); }

NB2_DEV V3 closest_tetrahedron(const Simplex& s, float* bc, unsigned& mask) {  // :236-320
    const float EPS = 1e-8f;
    const V3 v0 = s.D[0], v1 = s.D[1], v2 = s.D[2], v3 = s.D[3];
    const float det_t = det4(v0, v1, v2, v3);
    const bool degenerate = fabsf(det_t) < EPS;
    const float inv = 1.0f / (degenerate ? EPS : det_t);
    const V3 z;
    const float l0 = det4(z, v1, v2, v3) * inv, l1 = det4(v0, z, v2, v3) * inv, l2 = det4(v0, v1, z, v3) * inv;
    const float l3 = 1.0f - l0 - l1 - l2;
    float best = 1e30f;
    V3 cp;
    bc_clear(bc);
    mask = 0u;
    float tb[4];
    unsigned tm;
    if (l0 < 0.0f || degenerate) {
        V3 cl = closest_triangle(s, 1, 2, 3, tb, tm);
        float d2 = len2(cl);
        if (d2 < best) { for (int k = 0; k < 4; ++k) bc[k] = tb[k]; mask = tm; best = d2; cp = cl; }
    }
    if (l1 < 0.0f || degenerate) {
        V3 cl = closest_triangle(s, 0, 2, 3, tb, tm);
        float d2 = len2(cl);
        if (d2 < best) { for (int k = 0; k < 4; ++k) bc[k] = tb[k]; mask = tm; best = d2; cp = cl; }
    }
    if (l2 < 0.0f || degenerate) {
        V3 cl = closest_triangle(s, 0, 1, 3, tb, tm);
        float d2 = len2(cl);
        if (d2 < best) { for (int k = 0; k < 4; ++k) bc[k] = tb[k]; mask = tm; best = d2; cp = cl; }
    }
    if (l3 < 0.0f || degenerate) {
        V3 cl = closest_triangle(s, 0, 1, 2, tb, tm);
        float d2 = len2(cl);
        if (d2 < best) { for (int k = 0; k < 4; ++k) bc[k] = tb[k]; mask = tm; cp = cl; }
    }
    if (mask != 0u) return cp;
    bc[0] = l0; bc[1] = l1; bc[2] = l2; bc[3] = l3;
    mask = 15u;
    return z;
}

NB2_DEV void simplex_closest(const Simplex& s, V3& pa, V3& pb) {  // simplex_get_closest
    pa = V3();
    pb = V3();
    for (int i = 0; i < 4; ++i) {
        if ((s.mask & (1u << i)) == 0u) continue;
        pa = pa + s.bc[i] * (s.B[i] + s.D[i]);
        pb = pb + s.bc[i] * s.B[i];
    }
}

// solve_closest_distance_core (simplex_solver.py:322-470); returns `separated`
NB2_DEV bool gjk_core(const ConvexGeom& ga, const ConvexGeom& gb, Q4 qb, V3 pb, float extend, V3& point_a, V3& point_b, V3& normal,
                      float& distance, float CE = 1e-4f) {
    const float EPS = 1e-8f;
    distance = 0.0f;
    normal = V3();
    Simplex s;
    for (int i = 0; i < 4; ++i) { s.B[i] = V3(); s.D[i] = V3(); }
    bc_clear(s.bc);
    s.mask = 0u;
    int iter = 30;
    V3 v = ga.center - (pb + qrot(qb, gb.center));  // geometric_center(...).BtoA (simplex_solver.py:346-348)
    float dist_sq = len2(v);
    V3 last_dir(1.f, 0.f, 0.f);
    while (iter > 0) {
        iter -= 1;
        if (dist_sq < CE * CE) {
            simplex_closest(s, point_a, point_b);
            return false;
        }
        const V3 sd = -v;
        last_dir = sd;
        const MVert w = mink_support<false>(ga, gb, sd, qb, pb, extend);
        const V3 wv = w.BtoA;
        const float dd = dot(v, v - wv);
        if (dd <= 0.0f || dd * dd < (CE * CE * dist_sq)) break;
        bool dup = false;
        for (int i = 0; i < 4; ++i)
            if ((s.mask & (1u << i)) != 0u && len2(s.D[i] - wv) < CE * CE) {
                dup = true;
                break;
            }
        if (dup) break;
        int use = 0, free_slot = 0, idx[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            if ((s.mask & (1u << i)) != 0u) idx[use++] = i;
            else free_slot = i;
        }
        idx[use++] = free_slot;
        s.B[free_slot] = w.B;
        s.D[free_slot] = w.BtoA;
        V3 closest;
        bool success = true;
        if (use == 1) {
            closest = s.D[idx[0]];
            s.mask = 1u << idx[0];
            s.bc[idx[0]] = 1.0f;
        } else if (use == 2) {
            float bc[4];
            unsigned m;
            closest = closest_segment(s, idx[0], idx[1], bc, m);
            for (int k = 0; k < 4; ++k) s.bc[k] = bc[k];
            s.mask = m;
        } else if (use == 3) {
            float bc[4];
            unsigned m;
            closest = closest_triangle(s, idx[0], idx[1], idx[2], bc, m);
            for (int k = 0; k < 4; ++k) s.bc[k] = bc[k];
            s.mask = m;
        } else {
            float bc[4];
            unsigned m;
            closest = closest_tetrahedron(s, bc, m);
            for (int k = 0; k < 4; ++k) s.bc[k] = bc[k];
            s.mask = m;
            success = m != 15u;
        }
        if (!success) {
            simplex_closest(s, point_a, point_b);
            return false;
        }
        v = closest;
        dist_sq = len2(v);
    }
    simplex_closest(s, point_a, point_b);
    const V3 delta = point_b - point_a;
    const float dl2 = len2(delta);
    if (dl2 > EPS * EPS) {
        distance = sqrtf(dl2);
        normal = delta * (1.0f / distance);
    } else {
        distance = sqrtf(dist_sq);
        if (distance > CE) normal = v * (-1.0f / distance);
        else {
            const float nsq = len2(last_dir);
            normal = nsq > 0.0f ? last_dir * (1.0f / sqrtf(nsq)) : V3(1.f, 0.f, 0.f);
        }
    }
    return true;
}

// ---- manifold (multicontact.py) ----------------------------------------------------------------------------------
struct P2 {
    float x, y;
};
NB2_DEV float signed_area2(P2 a, P2 b, P2 q) { return (b.x - a.x) * (q.y - a.y) - (b.y - a.y) * (q.x - a.x); }
NB2_DEV float len2p(P2 a, P2 b) {
    const float dx = a.x - b.x, dy = a.y - b.y;
    return dx * dx + dy * dy;
}

NB2_CALL int trim_in_place(P2 s0, P2 s1, P2* loop, int loop_count) {  // multicontact.py:339-411
    if (loop_count < 3) return loop_count;
    P2 ia{0.f, 0.f}, ib{0.f, 0.f};
    int change_a = -1, change_b = -1;
    bool keep = false;
    bool prev_outside = signed_area2(s0, s1, loop[0]) <= 0.0f;
    for (int i = 0; i < loop_count; ++i) {
        const int nx = (i + 1) % loop_count;
        const bool outside = signed_area2(s0, s1, loop[nx]) <= 0.0f;
        if (outside != prev_outside) {
            const float sa = signed_area2(s0, s1, loop[i]), sb = signed_area2(s0, s1, loop[nx]);
            const float t = fabsf(sa) / fabsf(sa - sb);
            const P2 inter{(1.0f - t) * loop[i].x + t * loop[nx].x, (1.0f - t) * loop[i].y + t * loop[nx].y};
            if (change_a < 0) {
                change_a = i;
                keep = !prev_outside;
                ia = inter;
            } else {
                change_b = i;
                ib = inter;
            }
        }
        prev_outside = outside;
    }
    int new_count;
    if (change_a >= 0 && change_b >= 0) {
        int indexer = -1;
        new_count = loop_count;
        int i = 0;
        while (i < loop_count) {
            if (keep) {
                indexer += 1;
                loop[indexer] = loop[i];
            }
            if (i == change_a || i == change_b) {
                const P2 pt = i == change_a ? ia : ib;
                if (indexer == i && !keep) {
                    indexer += 1;
                    for (int k = new_count; k > indexer; --k) loop[k] = loop[k - 1];  // insert_vec2
                    loop[indexer] = pt;
                    new_count += 1;
                    i += 1;
                    change_b += 1;
                    loop_count += 1;
                } else {
                    indexer += 1;
                    loop[indexer] = pt;
                }
                keep = !keep;
            }
            i += 1;
        }
        new_count = indexer + 1;
    } else if (prev_outside) {
        new_count = 0;
    } else {
        new_count = loop_count;
    }
    return new_count;
}

NB2_DEV bool expand_segment(P2* poly) {  // 2-point polygon -> thin quad (multicontact.py:433-470)
    const P2 p0 = poly[0], p1 = poly[1];
    const float dx = p1.x - p0.x, dy = p1.y - p0.y;
    const float dl = sqrtf(dx * dx + dy * dy);
    if (!(dl > 1e-10f)) return false;
    const float inv = 1.0f / dl;
    const float ox = (-dy * inv) * 1e-5f, oy = (dx * inv) * 1e-5f;
    poly[0] = P2{p0.x - ox, p0.y - oy};
    poly[1] = P2{p1.x - ox, p1.y - oy};
    poly[2] = P2{p1.x + ox, p1.y + oy};
    poly[3] = P2{p0.x + ox, p0.y + oy};
    return true;
}

NB2_DEV int trim_all_in_place(P2* trim, int trim_count, P2* loop, int loop_count) {  // :414-483
    if (trim_count <= 1) return loop_count < 1 ? loop_count : 1;
    if (trim_count == 2) {
        if (!expand_segment(trim)) return loop_count < 1 ? loop_count : 1;
        trim_count = 4;
    }
    if (loop_count == 2) {
        if (!expand_segment(loop)) return loop_count < 1 ? loop_count : 1;
        loop_count = 4;
    }
    int cur = loop_count;
    const P2 t0 = trim[0];
    for (int i = 0; i < trim_count; ++i) {
        const P2 a = trim[i];
        const P2 b = i == trim_count - 1 ? t0 : trim[i + 1];
        cur = trim_in_place(a, b, loop, cur);
    }
    return cur;
}

NB2_CALL void calipers_quad(const P2* hull, int n, int out[4]) {  // approx_max_quadrilateral_area_with_calipers :486-560
    int p1 = 0, p3 = 1;
    float max_d2 = len2p(hull[p1], hull[p3]);
    const float tie = 1.0e-3f;
    int j = 1;
    for (int i = 0; i < n; ++i) {
        const P2 hi = hull[i], hi1 = hull[(i + 1) % n];
        while (true) {
            const float a1 = signed_area2(hi, hi1, hull[(j + 1) % n]);
            const float a0 = signed_area2(hi, hi1, hull[j]);
            if (a1 > a0) j = (j + 1) % n;
            else break;
        }
        const float d1 = len2p(hull[i], hull[j]);
        if (d1 > max_d2 * (1.0f + tie)) {
            max_d2 = d1;
            p1 = i;
            p3 = j;
        }
        const float d2 = len2p(hull[(i + 1) % n], hull[j]);
        if (d2 > max_d2 * (1.0f + tie)) {
            max_d2 = d2;
            p1 = (i + 1) % n;
            p3 = j;
        }
    }
    int p2 = 0, p4 = 0;
    float m1 = 0.0f, m2 = 0.0f;
    for (int i = 0; i < n; ++i) {
        const float area = signed_area2(hull[p1], hull[p3], hull[i]);
        if (area > m1 * (1.0f + tie)) {
            m1 = area;
            p2 = i;
        } else if (-area > m2 * (1.0f + tie)) {
            m2 = -area;
            p4 = i;
        }
    }
    out[0] = p1; out[1] = p2; out[2] = p3; out[3] = p4;
}

NB2_DEV int remove_zero_length_edges(P2* loop, int count, float eps) {  // :563-596
    if (count < 2) return 0;
    int w = 0;
    for (int r = 1; r < count; ++r)
        if (len2p(loop[r], loop[w]) > eps) {
            w += 1;
            loop[w] = loop[r];
        }
    int nc;
    if (w > 0) nc = len2p(loop[w], loop[0]) < eps ? w : w + 1;
    else nc = w + 1;
    if (nc < 2) nc = 0;
    return nc;
}

struct PlaneTracker {
    V3 ref, prev, normal;
    float largest;
};
NB2_DEV void tracker_update(PlaneTracker& t, V3 p, int id) {  // :165-186
    if (id == 0) {
        t.ref = p;
        t.largest = 0.0f;
    } else if (id == 1) {
        t.prev = p;
    } else {
        const V3 c = cross(t.prev - t.ref, p - t.ref);
        const float a2 = dot(c, c);
        if (a2 > t.largest) {
            t.largest = a2;
            t.normal = c;
        }
        t.prev = p;
    }
}
NB2_DEV V3 segment_projector_normal(V3 seg, V3 ref_n) {  // :189-209
    const V3 right = cross(seg, ref_n);
    const V3 n = cross(right, seg);
    const float l = len(n);
    return l > 1.0e-12f ? n * (1.0f / l) : ref_n;
}
struct Projector {
    float d;
    V3 n;
};
NB2_DEV V3 ray_plane(V3 o, V3 dir, float pd, V3 pn) {  // :107-127
    const float den = dot(dir, pn);
    if (fabsf(den) < 1.0e-12f) return o;
    const float t = -(dot(o, pn) + pd) / den;
    return o + dir * t;
}

// post_process_axial_on_discrete_contact (collision_core.py:174-277)
NB2_CALL void post_process_contact(V3& center, float& dist, V3 normal, float reff_a, float reff_b, int ta, V3 sca, V3 pos_a, Q4 rot_a, int tb,
                                  V3 scb, V3 pos_b, Q4 rot_b) {
    if (ta == CG_SPHERE || ta == CG_CAPSULE) {
        center = center + normal * (reff_a * 0.5f);
        dist = dist - reff_a;
    }
    if (tb == CG_SPHERE || tb == CG_CAPSULE) {
        center = center - normal * (reff_b * 0.5f);
        dist = dist - reff_b;
    }
    // is_discrete_shape (collision_core.py:40-48)
    const bool disc_a = ta == CG_BOX || ta == CG_PLANE || ta == CG_CONVEX_MESH, disc_b = tb == CG_BOX || tb == CG_PLANE || tb == CG_CONVEX_MESH;
    const bool ax_a = ta == CG_CYLINDER || ta == CG_CONE, ax_b = tb == CG_CYLINDER || tb == CG_CONE;
    if ((disc_a && ax_b) || (disc_b && ax_a)) {
        V3 axis, spos, an;
        float radius, hh;
        bool is_cone;
        if (disc_a && ax_b) {
            axis = qrot(rot_b, V3(0.f, 0.f, 1.f));
            radius = scb.x;
            hh = scb.y;
            is_cone = tb == CG_CONE;
            spos = pos_b;
            an = normal;
        } else {
            axis = qrot(rot_a, V3(0.f, 0.f, 1.f));
            radius = sca.x;
            hh = sca.y;
            is_cone = ta == CG_CONE;
            spos = pos_a;
            an = -normal;
        }
        const float and_ = fabsf(dot(axis, an));
        bool rolling = false;
        if (is_cone) {
            const float half = (float)atan2((double)radius, (double)(2.0f * hh));
            const float tol = 0.03490658503988659f;  // 2 deg
            if (and_ >= sin_w(half - tol) && and_ <= sin_w(half + tol)) rolling = true;
        } else if (and_ <= 0.03489949670250097f) {  // sin 2 deg
            rolling = true;
        }
        if (rolling) {
            const V3 pn = unit(cross(axis, an));
            center = center - pn * dot(center - spos, pn);  // project_point_onto_plane
        }
    }
}

// write_contact_speculative (sim/collide.py:257-280, contact_data.py:92-233): the writer admits a contact that is present now
// (separation <= the AUTHORED pair gap) or predicted to close within the collision-update interval.  `gap_sum` of the pair is then
// the velocity-extended SEARCH gap (collide.py:1832, 2010).  Origins are the shapes' own world origins (geom_transform), also for
// an infinite plane that is replaced by its box proxy.
struct ConvexSpec {
    float base_gap_sum, dt, max_extension;
    V3 origin_a, origin_b, lin_a, lin_b, ang_a, ang_b;
};
struct ConvexPairIn {
    int type_a, type_b;
    V3 scale_a, scale_b;
    Xf Xa, Xb;
    float margin_a, margin_b, gap_sum;
    const ConvexSpec* spec = nullptr;
    // CONVEX_MESH only: unscaled hull vertices (shape_source) and the centre of the scaled local AABB (Minkowski-centre seed)
    const float* hull_a = nullptr;
    const float* hull_b = nullptr;
    int hull_count_a = 0, hull_count_b = 0;
    V3 center_a, center_b;
};

// compute_gjk_mpr_contacts -> solve_convex_multi_contact -> build_manifold (+ the writer's own gap test).
// Outputs up to 5 contacts (world centre, normal, signed distance) in emission order; returns their count.
// radius_eff_* are the Minkowski radii the writer needs.
NB2_DEV int convex_contacts(const ConvexPairIn& in, float* odist, V3* opos, V3* onorm, float& reff_a, float& reff_b) {
    ConvexGeom ga(in.type_a, in.scale_a, in.center_a, in.hull_a, in.hull_count_a), gb(in.type_b, in.scale_b, in.center_b, in.hull_b, in.hull_count_b);
    reff_a = 0.0f;
    reff_b = 0.0f;
    if (ga.type == CG_SPHERE || ga.type == CG_CAPSULE) {
        reff_a = ga.scale.x;
        ga.scale.x = 0.0001f;
    }
    if (gb.type == CG_SPHERE || gb.type == CG_CAPSULE) {
        reff_b = gb.scale.x;
        gb.scale.x = 0.0001f;
    }
    const float threshold = in.gap_sum + reff_a + reff_b + in.margin_a + in.margin_b;
    const bool skip_multi = ga.type == CG_SPHERE || gb.type == CG_SPHERE || ga.type == CG_ELLIPSOID || gb.type == CG_ELLIPSOID;
    const Q4 qa = in.Xa.q;
    const V3 pa_w = in.Xa.p;
    const Q4 rq = qmul(qconj(qa), in.Xb.q);
    const V3 rp = qrot_inv(qa, in.Xb.p - in.Xa.p);
    const float margin_sum = in.margin_a + in.margin_b;
    const float eps = 1.0e-4f;
    const float enlarge = margin_sum <= 0.0f ? eps : (margin_sum < eps ? 2.0f * eps : 0.0f);
    V3 p_a, p_b, n;
    float pen, sd;
    if (mpr_core(ga, gb, rq, rp, enlarge, p_a, p_b, n, pen)) {
        sd = -pen + enlarge;
        const float he = enlarge * 0.5f;
        p_a = p_a - n * he;
        p_b = p_b + n * he;
    } else {
        gjk_core(ga, gb, rq, rp, 0.0f, p_a, p_b, n, sd);
    }
    const float total_sep = reff_a + reff_b + in.margin_a + in.margin_b;
    int count = 0;
    // one contact: post-process + the writer's gap test (write_contact, collide.py:210-254)
    auto emit = [&](V3 center_w, V3 normal_w, float dist, V3 pos_b_w, Q4 rot_b_w) {
        post_process_contact(center_w, dist, normal_w, reff_a, reff_b, ga.type, ga.scale, in.Xa.p, in.Xa.q, gb.type, gb.scale, pos_b_w, rot_b_w);
        const V3 nn = unit(normal_w);
        const V3 a_w = center_w - nn * (0.5f * dist + reff_a);
        const V3 b_w = center_w + nn * (0.5f * dist + reff_b);
        const float dd = dot(b_w - a_w, nn) - total_sep;
        if (in.spec) {
            const ConvexSpec& w = *in.spec;
            if (!(dd <= w.base_gap_sum)) {  // contact_passes_speculative_gap_check: predictive score >= 0
                const V3 va = w.lin_a + cross(w.ang_a, a_w - w.origin_a), vb = w.lin_b + cross(w.ang_b, b_w - w.origin_b);
                const float approach = fmax_w(-dot(vb - va, nn), 0.0f);
                const float extension = fmin_w(approach * w.dt, w.max_extension);
                if (!(extension - dd >= 0.0f)) return;
            }
        } else if (dd > in.gap_sum) return;
        odist[count] = dist;
        opos[count] = center_w;
        onorm[count] = normal_w;
        count += 1;
    };
    if (skip_multi || sd > threshold) {
        V3 pt = 0.5f * (p_a + p_b);
        pt = qrot(qa, pt) + pa_w;
        emit(pt, qrot(qa, n), sd, in.Xb.p, in.Xb.q);
        return count;
    }
    // ---- build_manifold (multicontact.py:779-956), in A's local frame ----
    const float PC[5] = {1.0f, 0.30901699437494745f, -0.8090169943749473f, -0.8090169943749476f, 0.30901699437494723f};
    const float PS[5] = {0.0f, 0.9510565162951535f, 0.5877852522924732f, -0.587785252292473f, -0.9510565162951536f};
    const float SIN_T = 0.03489949670250097f, COS_T = 0.9993908270190958f, EPSM = 0.00001f;
    int a_count = 0, b_count = 0;
    V3 ta, tb;  // orthonormal_basis(normal) (math/__init__.py:243-276)
    if (n.z < 0.0f) {
        const float a = 1.0f / (1.0f - n.z), b = n.x * n.y * a;
        ta = V3(1.0f - n.x * n.x * a, -b, n.x);
        tb = V3(b, n.y * n.y * a - 1.0f, -n.y);
    } else {
        const float a = 1.0f / (1.0f + n.z), b = -n.x * n.y * a;
        ta = V3(1.0f - n.x * n.x * a, b, -n.x);
        tb = V3(b, 1.0f - n.y * n.y * a, -n.y);
    }
    PlaneTracker tr_a, tr_b;
    tr_a.largest = 0.0f;
    tr_b.largest = 0.0f;
    const V3 center = 0.5f * (p_a + p_b);
    P2 buf[10];  // b polygon in buf[0..], a polygon aliased at buf[5..] exactly like the reference buffers
    for (int i = 0; i < 10; ++i) buf[i] = P2{0.f, 0.f};
    P2* bbuf = buf;
    P2* abuf = buf + 5;
    const V3 ln_b = qrot_inv(rq, -n), lta_b = qrot_inv(rq, -ta), ltb_b = qrot_inv(rq, -tb);
    for (int e = 0; e < 5; ++e) {
        const float c_sin = PC[e] * SIN_T, s_sin = PS[e] * SIN_T;
        const V3 dir_a = n * COS_T + c_sin * ta + s_sin * tb;
        const V3 pa3 = support_map(ga, dir_a);
        const V3 pra = pa3 - center;
        const P2 a2{dot(ta, pra), dot(tb, pra)};
        bool add_a = true;  // add_avoid_duplicates_vec2 (:599-617)
        if (a_count > 0 && len2p(abuf[0], a2) < EPSM) add_a = false;
        if (add_a && a_count > 1 && len2p(abuf[a_count - 1], a2) < EPSM) add_a = false;
        if (add_a) {
            abuf[a_count] = a2;
            a_count += 1;
            tracker_update(tr_a, pa3, a_count - 1);
        }
        const V3 ldb = ln_b * COS_T + c_sin * lta_b + s_sin * ltb_b;
        const V3 pbl = support_map(gb, ldb);
        const V3 pb3 = qrot(rq, pbl) + rp;
        const V3 prb = pb3 - center;
        const P2 b2{dot(ta, prb), dot(tb, prb)};
        bool add_b = true;
        if (b_count > 0 && len2p(bbuf[0], b2) < EPSM) add_b = false;
        if (add_b && b_count > 1 && len2p(bbuf[b_count - 1], b2) < EPSM) add_b = false;
        if (add_b) {
            bbuf[b_count] = b2;
            b_count += 1;
            tracker_update(tr_b, pb3, b_count - 1);
        }
    }
    const V3 normal_w = qrot(qa, n);
    const V3 pos_b_ws = qrot(qa, rp) + pa_w;  // the manifold path re-derives B's world pose from the relative one
    const Q4 rot_b_ws = qmul(qa, rq);
    int count_out = 0;
    float normal_dot = 0.0f;
    if (!(a_count < 2 || b_count < 2)) {
        Projector pja, pjb;  // create_body_projectors (:212-262)
        pja.d = 0.f; pjb.d = 0.f;
        if (tr_a.largest == 0.0f && tr_b.largest == 0.0f) {
            const V3 da = tr_a.prev - tr_a.ref, db = tr_b.prev - tr_b.ref;
            const V3 ma = 0.5f * (tr_a.ref + tr_a.prev), mb = 0.5f * (tr_b.ref + tr_b.prev);
            pja.n = segment_projector_normal(da, n);
            pja.d = -dot(ma, pja.n);
            pjb.n = segment_projector_normal(db, n);
            pjb.d = -dot(mb, pjb.n);
        } else {
            if (tr_a.largest > 0.0f) {
                const float inv = 1.0f / sqrtf(fmax_w(1.0e-12f, tr_a.largest));
                pja.n = tr_a.normal * inv;
                pja.d = -dot(p_a, pja.n);
            }
            if (tr_b.largest > 0.0f) {
                const float inv = 1.0f / sqrtf(fmax_w(1.0e-12f, tr_b.largest));
                pjb.n = tr_b.normal * inv;
                pjb.d = -dot(p_b, pjb.n);
            }
            if (tr_a.largest == 0.0f) {
                const V3 dr = tr_a.prev - tr_a.ref, mid = 0.5f * (tr_a.ref + tr_a.prev);
                pja.n = segment_projector_normal(dr, pjb.n);
                pja.d = -dot(mid, pja.n);
            }
            if (tr_b.largest == 0.0f) {
                const V3 dr = tr_b.prev - tr_b.ref, mid = 0.5f * (tr_b.ref + tr_b.prev);
                pjb.n = segment_projector_normal(dr, pja.n);
                pjb.d = -dot(mid, pjb.n);
            }
        }
        const bool dev = fabsf(dot(n, pja.n)) < COS_T || fabsf(dot(n, pjb.n)) < COS_T;  // excess_normal_deviation
        if (!dev) {
            // extract_4_point_contact_manifolds (:655-777)
            normal_dot = fabsf(dot(pja.n, pjb.n));
            int loop_count = trim_all_in_place(abuf, a_count, bbuf, b_count);
            loop_count = remove_zero_length_edges(bbuf, loop_count, EPSM);
            if (loop_count > 1) {
                int sel[4] = {0, 1, 2, 3};
                if (loop_count > 4) {
                    calipers_quad(bbuf, loop_count, sel);
                    loop_count = 4;
                }
                for (int i = 0; i < loop_count; ++i) {
                    const P2 q = bbuf[sel[i]];
                    const V3 pl = q.x * ta + q.y * tb + center;
                    const V3 a = ray_plane(pl, n, pja.d, pja.n), b = ray_plane(pl, n, pjb.d, pjb.n);
                    const V3 cl = 0.5f * (a + b);
                    const float sdist = dot(b - a, n);
                    emit(qrot(qa, cl) + pa_w, normal_w, sdist, pos_b_ws, rot_b_ws);
                }
                count_out = loop_count < 4 ? loop_count : 4;
            } else {
                normal_dot = 0.0f;
            }
        }
    }
    if (normal_dot < 0.9999984769132877f || count_out == 0) {  // should_include_deepest_contact: cos(0.1 deg)
        const V3 dc = 0.5f * (p_a + p_b);
        emit(qrot(qa, dc) + pa_w, normal_w, dot(p_b - p_a, n), pos_b_ws, rot_b_ws);
    }
    return count;
}

// compute_tight_aabb_from_support, generic branch (collision_core.py:454-548): six support evaluations along the world axes
// expressed in the shape frame.  Used for the shapes compute_shape_aabbs has no closed form for (cones).
NB2_DEV void tight_aabb_from_support(const ConvexGeom& g, Q4 q, V3 pos, V3& lo, V3& hi) {
    const M33 R = qmat(q);
    const V3 lx(R.at(0, 0), R.at(0, 1), R.at(0, 2)), ly(R.at(1, 0), R.at(1, 1), R.at(1, 2)), lz(R.at(2, 0), R.at(2, 1), R.at(2, 2));
    if (g.type == CG_CONVEX_MESH) {  // single pass over the vertices (collision_core.py:492-523)
        const V3 sx = cmul(lx, g.scale), sy = cmul(ly, g.scale), sz = cmul(lz, g.scale);
        V3 mn(1.0e10f, 1.0e10f, 1.0e10f), mx(-1.0e10f, -1.0e10f, -1.0e10f);
        for (int i = 0; i < g.hull_count; ++i) {
            const V3 p(g.hull[3 * i], g.hull[3 * i + 1], g.hull[3 * i + 2]);
            const V3 v(dot(p, sx), dot(p, sy), dot(p, sz));
            mn = vmin(mn, v);
            mx = vmax(mx, v);
        }
        lo = mn + pos;
        hi = mx + pos;
        return;
    }
    const float max_x = dot(lx, support_map(g, lx)), max_y = dot(ly, support_map(g, ly)), max_z = dot(lz, support_map(g, lz));
    const float min_x = dot(lx, support_map(g, -lx)), min_y = dot(ly, support_map(g, -ly)), min_z = dot(lz, support_map(g, -lz));
    lo = V3(min_x, min_y, min_z) + pos;
    hi = V3(max_x, max_y, max_z) + pos;
}

struct ConvexPairAabbs {  // the shapes' world AABBs as compute_shape_aabbs wrote them (external_aabb=True path)
    V3 lo_a, hi_a, lo_b, hi_b;
};

// narrow_phase_kernel_gjk_mpr + find_contacts (narrow_phase.py:1098-1216, collision_core.py:700-790): an infinite plane that
// reaches the generic path (plane-cone, plane-barrel-cylinder) is first tested against the other shape's bounding sphere and
// then replaced by a box proxy whose top face is the plane (convert_infinite_plane_to_cube, collision_core.py:566-626).
NB2_DEV int convex_contacts_any(ConvexPairIn in, const ConvexPairAabbs& bb, float* odist, V3* opos, V3* onorm, float& reff_a, float& reff_b) {
    reff_a = 0.0f;
    reff_b = 0.0f;
    // geom_data keeps HALF extents for finite planes (collide.py:452-453); infinite planes stay (0, 0)
    if (in.type_a == CG_PLANE) in.scale_a = V3(in.scale_a.x * 0.5f, in.scale_a.y * 0.5f, 0.0f);
    if (in.type_b == CG_PLANE) in.scale_b = V3(in.scale_b.x * 0.5f, in.scale_b.y * 0.5f, 0.0f);
    const bool inf_a = in.type_a == CG_PLANE && in.scale_a.x == 0.0f && in.scale_a.y == 0.0f;
    const bool inf_b = in.type_b == CG_PLANE && in.scale_b.x == 0.0f && in.scale_b.y == 0.0f;
    if (inf_a && inf_b) return 0;
    if (inf_a || inf_b) {
        const V3 ca = 0.5f * (bb.lo_a + bb.hi_a), cb = 0.5f * (bb.lo_b + bb.hi_b);  // compute_bounding_sphere_from_aabb
        const float ra = len(0.5f * (bb.hi_a - bb.lo_a)), rb = len(0.5f * (bb.hi_b - bb.lo_b));
        const Xf plane = inf_a ? in.Xa : in.Xb;
        const V3 other_center = inf_a ? cb : ca;
        const float other_radius = inf_a ? rb : ra;
        const V3 n = qrot(plane.q, V3(0.f, 0.f, 1.f));
        // check_infinite_plane_bsphere_overlap; speculative mode adds the pair's search extension to the non-plane shape's overlap
        // radius (narrow_phase.py:1170-1175) - the proxy below is sized from the plain radius + one pair gap
        if (dot(other_center - plane.p, n) > (in.spec ? other_radius + in.gap_sum : other_radius)) return 0;
        const V3 other_pos = inf_a ? in.Xb.p : in.Xa.p;
        const float size = (other_radius + in.gap_sum) * 10.0f;  // lateral_size == depth
        const float dist_n = dot(other_pos - plane.p, n);
        const V3 surface = other_pos - n * dist_n;
        const V3 adjusted = surface - n * size;
        if (inf_a) {
            in.type_a = CG_BOX;
            in.scale_a = V3(size, size, size);
            in.Xa.p = adjusted;
        } else {
            in.type_b = CG_BOX;
            in.scale_b = V3(size, size, size);
            in.Xb.p = adjusted;
        }
    }
    return convex_contacts(in, odist, opos, onorm, reff_a, reff_b);
}

}  // namespace nb2

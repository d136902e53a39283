// oracle_gjk.h - TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Generic convex-convex contacts (MPR / GJK + manifold; reference geometry/narrow_phase.py:1041-1216,
// collision_core.py:337-450, collision_convex.py:107-231, mpr.py:188-403, simplex_solver.py:322-495,
// multicontact.py:779-956).
//
// SINGLE SOURCE: unlike the rest of the oracle, these ~900 lines are NOT restated a second time.  The routine is
// written once as host+device code in newton_b200/csrc/nb2_convex.cuh and compiled here by g++ with
// -ffp-contract=off (the strict-fp arithmetic of the product build).  Consequently GPU-vs-oracle equality on these
// rows only proves that the device compilation computes what the host compilation does; that the algorithm is the
// reference's is pinned separately by the reference's own known answers for this path
// (tests/test_oracle_known_answers.py: box-box face / edge / separated cases from
// newton/tests/test_collision_primitives.py + test_narrow_phase.py, 5-box stack from test_solver_xpbd.py).
#pragma once
#include "../newton_b200/csrc/nb2_convex.cuh"
#include "oracle_collide.h"
namespace orc {

inline nb2::V3 to_nb2(vec3 v) { return nb2::V3(v.x, v.y, v.z); }
inline vec3 from_nb2(nb2::V3 v) { return vec3(v.x, v.y, v.z); }
inline nb2::Xf to_nb2(const transform& t) { return nb2::Xf(to_nb2(t.p), nb2::Q4(t.q.x, t.q.y, t.q.z, t.q.w)); }

// One pair through compute_gjk_mpr_contacts; contacts come back already gap-tested, in sort_sub_key order.
inline int convex_pair(int type_a, vec3 scale_a, const transform& Xa, float margin_a, int type_b, vec3 scale_b, const transform& Xb,
                       float margin_b, float gap_sum, float* dist, vec3* pos, vec3* normal, float& reff_a, float& reff_b,
                       vec3 lo_a = vec3(), vec3 hi_a = vec3(), vec3 lo_b = vec3(), vec3 hi_b = vec3()) {
    reff_a = reff_b = 0.0f;
    nb2::ConvexPairIn in;
    in.type_a = type_a;
    in.type_b = type_b;
    in.scale_a = to_nb2(scale_a);
    in.scale_b = to_nb2(scale_b);
    in.Xa = to_nb2(Xa);
    in.Xb = to_nb2(Xb);
    in.margin_a = margin_a;
    in.margin_b = margin_b;
    in.gap_sum = gap_sum;
    nb2::V3 p[5], n[5];
    nb2::ConvexPairAabbs bb{to_nb2(lo_a), to_nb2(hi_a), to_nb2(lo_b), to_nb2(hi_b)};
    int cnt = nb2::convex_contacts_any(in, bb, dist, p, n, reff_a, reff_b);
    for (int i = 0; i < cnt; ++i) {
        pos[i] = from_nb2(p[i]);
        normal[i] = from_nb2(n[i]);
    }
    return cnt;
}

// The GJK/MPR kernel pass over the pairs the primitive kernel forwarded (narrow_phase.py:1041-1216).
inline void gjk_mpr_pairs(const nb2_model_desc& m, const float* body_q, CollideResult& res) {
    if (res.gjk_pairs.empty()) return;
    std::vector<ShapeGeom> geom;
    compute_shape_aabbs(m, body_q, geom);
    for (auto& pr : res.gjk_pairs) {
        const int shape_a = pr.first, shape_b = pr.second;
        const ShapeGeom& A = geom[shape_a];
        const ShapeGeom& B = geom[shape_b];
        float dist[5], ra, rb;
        vec3 pos[5], normal[5];
        const float gap_sum = m.shape_gap[shape_a] + m.shape_gap[shape_b];
        // raw model scales: convex_contacts_any applies the geom_data halving of finite planes itself
        int cnt = convex_pair(m.shape_type[shape_a], load3(m.shape_scale + 3 * shape_a), A.X_ws, A.margin, m.shape_type[shape_b],
                              load3(m.shape_scale + 3 * shape_b), B.X_ws, B.margin, gap_sum,
                              dist, pos, normal, ra, rb, A.aabb_lower, A.aabb_upper, B.aabb_lower, B.aabb_upper);
        for (int i = 0; i < cnt; ++i) {
            RawContact rc;
            // emission order == sort_sub_key order; dropped manifold points only leave holes in the sub-key sequence
            write_contact(m, body_q, shape_a, shape_b, pos[i], normal[i], dist[i], ra, rb, A.margin, B.margin, i, false, rc);
            res.contacts.push_back(rc);
        }
    }
}

// Direct solver cores in A's frame (what newton/tests/test_mpr.py and test_gjk.py launch).
inline int mpr_core_test(int type_a, vec3 scale_a, int type_b, vec3 scale_b, vec3 pos_b, quat quat_b, float extend, float* out11) {
    nb2::ConvexGeom ga{type_a, to_nb2(scale_a)}, gb{type_b, to_nb2(scale_b)};
    nb2::V3 pa, pb, n;
    float pen;
    bool hit = nb2::mpr_core(ga, gb, nb2::Q4(quat_b.x, quat_b.y, quat_b.z, quat_b.w), to_nb2(pos_b), extend, pa, pb, n, pen);
    const float o[10] = {pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, n.x, n.y, n.z, pen};
    for (int i = 0; i < 10; ++i) out11[i] = o[i];
    return hit ? 1 : 0;
}
inline int gjk_core_test(int type_a, vec3 scale_a, int type_b, vec3 scale_b, vec3 pos_b, quat quat_b, float extend, float eps,
                         float* out11) {
    nb2::ConvexGeom ga{type_a, to_nb2(scale_a)}, gb{type_b, to_nb2(scale_b)};
    nb2::V3 pa, pb, n;
    float dist;
    bool separated = nb2::gjk_core(ga, gb, nb2::Q4(quat_b.x, quat_b.y, quat_b.z, quat_b.w), to_nb2(pos_b), extend, pa, pb, n, dist, eps);
    const float o[10] = {pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, n.x, n.y, n.z, dist};
    for (int i = 0; i < 10; ++i) out11[i] = o[i];
    return separated ? 1 : 0;
}
inline vec3 support_map_test(int type, vec3 scale, vec3 dir) { return from_nb2(nb2::support_map(nb2::ConvexGeom{type, to_nb2(scale)}, to_nb2(dir))); }

inline int convex_pair_test(int type_a, vec3 scale_a, const transform& Xa, int type_b, vec3 scale_b, const transform& Xb, float gap_sum,
                            float* dist5, float* pos15, float* normal15) {
    float ra, rb;
    vec3 pos[5], normal[5];
    // AABBs as the stand-alone NarrowPhase computes them (narrow_phase.py:1120-1150): tight support AABB +- the shape's gap;
    // only the bounding-sphere radius of the shape facing an infinite plane depends on them
    auto aabb = [&](int type, vec3 scale, const transform& X, vec3& lo, vec3& hi) {
        if (type == GEO_PLANE) {
            lo = X.p - vec3(1.0e6f);
            hi = X.p + vec3(1.0e6f);
            return;
        }
        nb2::V3 l, h;
        nb2::tight_aabb_from_support(nb2::ConvexGeom{type, to_nb2(scale)}, nb2::Q4(X.q.x, X.q.y, X.q.z, X.q.w), to_nb2(X.p), l, h);
        const vec3 g(0.5f * gap_sum);
        lo = from_nb2(l) - g;
        hi = from_nb2(h) + g;
    };
    vec3 lo_a, hi_a, lo_b, hi_b;
    aabb(type_a, scale_a, Xa, lo_a, hi_a);
    aabb(type_b, scale_b, Xb, lo_b, hi_b);
    int cnt = convex_pair(type_a, scale_a, Xa, 0.0f, type_b, scale_b, Xb, 0.0f, gap_sum, dist5, pos, normal, ra, rb, lo_a, hi_a, lo_b, hi_b);
    for (int i = 0; i < cnt; ++i) {
        store3(pos15 + 3 * i, pos[i]);
        store3(normal15 + 3 * i, normal[i]);
    }
    return cnt;
}
}  // namespace orc

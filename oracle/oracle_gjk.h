// oracle_gjk.h - TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Generic convex-convex contacts (MPR / GJK + manifold; reference geometry/narrow_phase.py:1041-1216,
// collision_core.py:337-450, collision_convex.py:107-231, mpr.py:188-403, simplex_solver.py:322-495,
// multicontact.py:779-956) as the oracle's collide pipeline runs them.
//
// TWO implementations are reachable from here:
//   impl 0 (default, what orc_collide uses): oracle_convex.h, the oracle's own restatement written from the reference;
//   impl 1: the PRODUCT's host+device routine newton_b200/csrc/nb2_convex.cuh compiled by g++ with -ffp-contract=off.
// tests/test_oracle_known_answers.py pins impl 0 with the reference's known answers; tests/test_abi_and_host.py checks
// impl 0 == impl 1 bit for bit on randomised pairs of every shape-type combination (CPU only), and the GPU parity
// tests compare the CUDA compilation with impl 0.
#pragma once
#include "../newton_b200/csrc/nb2_convex.cuh"
#include "oracle_collide.h"
#include "oracle_convex.h"
namespace orc {

inline nb2::V3 to_nb2(vec3 v) { return nb2::V3(v.x, v.y, v.z); }
inline vec3 from_nb2(nb2::V3 v) { return vec3(v.x, v.y, v.z); }
inline nb2::Xf to_nb2(const transform& t) { return nb2::Xf(to_nb2(t.p), nb2::Q4(t.q.x, t.q.y, t.q.z, t.q.w)); }

// One pair through compute_gjk_mpr_contacts; contacts come back already gap-tested, in sort_sub_key order.
inline int convex_pair(int type_a, vec3 scale_a, const transform& Xa, float margin_a, int type_b, vec3 scale_b, const transform& Xb,
                       float margin_b, float gap_sum, float* dist, vec3* pos, vec3* normal, float& reff_a, float& reff_b,
                       vec3 lo_a = vec3(), vec3 hi_a = vec3(), vec3 lo_b = vec3(), vec3 hi_b = vec3(), int impl = 0,
                       const cvx::HullRef& hull_a = cvx::HullRef(), const cvx::HullRef& hull_b = cvx::HullRef(),
                       const cvx::SpeculativeWriter* spec = nullptr) {
    reff_a = reff_b = 0.0f;
    if (impl == 0) {
        cvx::ContactOut out[5];
        int cnt = cvx::gjk_mpr_pair(type_a, scale_a, Xa, margin_a, lo_a, hi_a, type_b, scale_b, Xb, margin_b, lo_b, hi_b, gap_sum, out, reff_a, reff_b,
                                    hull_a, hull_b, spec);
        for (int i = 0; i < cnt; ++i) {
            dist[i] = out[i].distance;
            pos[i] = out[i].center;
            normal[i] = out[i].normal;
        }
        return cnt;
    }
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
    in.hull_a = hull_a.points;
    in.hull_count_a = hull_a.count;
    in.center_a = to_nb2(0.5f * (hull_a.local_aabb_lower + hull_a.local_aabb_upper));
    in.hull_b = hull_b.points;
    in.hull_count_b = hull_b.count;
    in.center_b = to_nb2(0.5f * (hull_b.local_aabb_lower + hull_b.local_aabb_upper));
    nb2::ConvexSpec cs;
    if (spec && spec->enabled) {
        cs.base_gap_sum = spec->base_gap_sum;
        cs.dt = spec->dt;
        cs.max_extension = spec->max_extension;
        cs.origin_a = to_nb2(spec->origin_a);
        cs.origin_b = to_nb2(spec->origin_b);
        cs.lin_a = to_nb2(spec->linear_velocity_a);
        cs.lin_b = to_nb2(spec->linear_velocity_b);
        cs.ang_a = to_nb2(spec->angular_velocity_a);
        cs.ang_b = to_nb2(spec->angular_velocity_b);
        in.spec = &cs;
    }
    nb2::V3 p[5], n[5];
    nb2::ConvexPairAabbs bb{to_nb2(lo_a), to_nb2(hi_a), to_nb2(lo_b), to_nb2(hi_b)};
    int cnt = nb2::convex_contacts_any(in, bb, dist, p, n, reff_a, reff_b);
    for (int i = 0; i < cnt; ++i) {
        pos[i] = from_nb2(p[i]);
        normal[i] = from_nb2(n[i]);
    }
    return cnt;
}

inline cvx::HullRef hull_ref(const nb2_model_desc& m, int shape) {
    cvx::HullRef h;
    if (m.shape_type[shape] == GEO_CONVEX_MESH && m.hull_points && m.shape_hull_start) {
        h.points = m.hull_points + 3 * size_t(m.shape_hull_start[shape]);
        h.count = m.shape_hull_count[shape];
        h.local_aabb_lower = load3(m.shape_collision_aabb_lower + 3 * shape);
        h.local_aabb_upper = load3(m.shape_collision_aabb_upper + 3 * shape);
    }
    return h;
}

// The GJK/MPR kernel pass over the pairs the primitive kernel forwarded (narrow_phase.py:1041-1216).
inline void gjk_mpr_pairs(const nb2_model_desc& m, const float* body_q, CollideResult& res,
                          const SpeculativeParams& sp = SpeculativeParams()) {
    if (res.gjk_pairs.empty()) return;
    std::vector<ShapeGeom> geom;
    compute_shape_aabbs(m, body_q, geom);
    if (sp.active) compute_shape_velocities(m, body_q, sp, geom);
    for (auto& pr : res.gjk_pairs) {
        const int shape_a = pr.first, shape_b = pr.second;
        const ShapeGeom& A = geom[shape_a];
        const ShapeGeom& B = geom[shape_b];
        float dist[5], ra, rb;
        vec3 pos[5], normal[5];
        const float base_gap_sum = m.shape_gap[shape_a] + m.shape_gap[shape_b];
        // speculative: the kernel sees the velocity-extended search gaps (collide.py:1832, 2010), the writer the authored ones
        const float gap_sum = sp.active ? A.search_gap + B.search_gap : base_gap_sum;
        cvx::SpeculativeWriter w;
        if (sp.enabled) {
            w.enabled = true;
            w.base_gap_sum = base_gap_sum;
            w.dt = sp.dt;
            w.max_extension = sp.max_extension;
            w.origin_a = A.X_ws.p;
            w.origin_b = B.X_ws.p;
            w.linear_velocity_a = A.linear_velocity;
            w.linear_velocity_b = B.linear_velocity;
            w.angular_velocity_a = A.angular_velocity;
            w.angular_velocity_b = B.angular_velocity;
        }
        // raw model scales: convex_contacts_any applies the geom_data halving of finite planes itself
        int cnt = convex_pair(m.shape_type[shape_a], load3(m.shape_scale + 3 * shape_a), A.X_ws, A.margin, m.shape_type[shape_b],
                              load3(m.shape_scale + 3 * shape_b), B.X_ws, B.margin, gap_sum,
                              dist, pos, normal, ra, rb, A.aabb_lower, A.aabb_upper, B.aabb_lower, B.aabb_upper, 0, hull_ref(m, shape_a),
                              hull_ref(m, shape_b), sp.enabled ? &w : nullptr);
        for (int i = 0; i < cnt; ++i) {
            RawContact rc;
            // emission order == sort_sub_key order; dropped manifold points only leave holes in the sub-key sequence
            write_contact(m, body_q, shape_a, shape_b, pos[i], normal[i], dist[i], ra, rb, A.margin, B.margin, i, false, rc);
            res.contacts.push_back(rc);
        }
    }
}

// Direct solver cores in A's frame (what newton/tests/test_mpr.py and test_gjk.py launch).
inline cvx::GenericShapeData cvx_geom(int type, vec3 scale) {
    cvx::GenericShapeData g;
    g.shape_type = type;
    g.scale = scale;
    return g;
}
inline int mpr_core_test(int type_a, vec3 scale_a, int type_b, vec3 scale_b, vec3 pos_b, quat quat_b, float extend, float* out11, int impl = 0) {
    if (impl == 0) {
        cvx::MprResult r = cvx::solve_mpr_core(cvx_geom(type_a, scale_a), cvx_geom(type_b, scale_b), quat_b, pos_b, extend);
        const float o[10] = {r.point_a.x, r.point_a.y, r.point_a.z, r.point_b.x, r.point_b.y, r.point_b.z, r.normal.x, r.normal.y, r.normal.z, r.penetration};
        for (int i = 0; i < 10; ++i) out11[i] = o[i];
        return r.collision ? 1 : 0;
    }
    nb2::ConvexGeom ga{type_a, to_nb2(scale_a)}, gb{type_b, to_nb2(scale_b)};
    nb2::V3 pa, pb, n;
    float pen;
    bool hit = nb2::mpr_core(ga, gb, nb2::Q4(quat_b.x, quat_b.y, quat_b.z, quat_b.w), to_nb2(pos_b), extend, pa, pb, n, pen);
    const float o[10] = {pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, n.x, n.y, n.z, pen};
    for (int i = 0; i < 10; ++i) out11[i] = o[i];
    return hit ? 1 : 0;
}
inline int gjk_core_test(int type_a, vec3 scale_a, int type_b, vec3 scale_b, vec3 pos_b, quat quat_b, float extend, float eps,
                         float* out11, int impl = 0) {
    if (impl == 0) {
        cvx::GjkResult r = cvx::solve_closest_distance_core(cvx_geom(type_a, scale_a), cvx_geom(type_b, scale_b), quat_b, pos_b, extend, 30, eps);
        const float o[10] = {r.point_a.x, r.point_a.y, r.point_a.z, r.point_b.x, r.point_b.y, r.point_b.z, r.normal.x, r.normal.y, r.normal.z, r.distance};
        for (int i = 0; i < 10; ++i) out11[i] = o[i];
        return r.separated ? 1 : 0;
    }
    nb2::ConvexGeom ga{type_a, to_nb2(scale_a)}, gb{type_b, to_nb2(scale_b)};
    nb2::V3 pa, pb, n;
    float dist;
    bool separated = nb2::gjk_core(ga, gb, nb2::Q4(quat_b.x, quat_b.y, quat_b.z, quat_b.w), to_nb2(pos_b), extend, pa, pb, n, dist, eps);
    const float o[10] = {pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, n.x, n.y, n.z, dist};
    for (int i = 0; i < 10; ++i) out11[i] = o[i];
    return separated ? 1 : 0;
}
inline vec3 support_map_test(int type, vec3 scale, vec3 dir, int impl = 0) {
    if (impl == 0) return cvx::support_map(cvx_geom(type, scale), dir);
    return from_nb2(nb2::support_map(nb2::ConvexGeom{type, to_nb2(scale)}, to_nb2(dir)));
}

inline cvx::HullRef hull_of_points(const float* points, int count, vec3 scale) {  // what the builder derives for a CONVEX_MESH shape
    cvx::HullRef h;
    h.points = points;
    h.count = count;
    if (count > 0) {  // scaled local AABB (sim/builder.py:11605-11610)
        vec3 lo = cw_mul(load3(points), scale), hi = lo;
        for (int i = 1; i < count; ++i) {
            vec3 p = cw_mul(load3(points + 3 * i), scale);
            lo = vec3(minf(lo.x, p.x), minf(lo.y, p.y), minf(lo.z, p.z));
            hi = vec3(maxf(hi.x, p.x), maxf(hi.y, p.y), maxf(hi.z, p.z));
        }
        h.local_aabb_lower = lo;
        h.local_aabb_upper = hi;
    }
    return h;
}
inline cvx::GenericShapeData cvx_geom_hull(int type, vec3 scale, const cvx::HullRef& h) {
    cvx::GenericShapeData g = cvx_geom(type, scale);
    if (type == GEO_CONVEX_MESH) {
        g.mesh_points = h.points;
        g.mesh_point_count = h.count;
        g.center = 0.5f * (h.local_aabb_lower + h.local_aabb_upper);
    }
    return g;
}
inline nb2::ConvexGeom nb2_geom_hull(int type, vec3 scale, const cvx::HullRef& h) {
    if (type != GEO_CONVEX_MESH) return nb2::ConvexGeom(type, to_nb2(scale));
    return nb2::ConvexGeom(type, to_nb2(scale), to_nb2(0.5f * (h.local_aabb_lower + h.local_aabb_upper)), h.points, h.count);
}

inline int convex_pair_test(int type_a, vec3 scale_a, const transform& Xa, int type_b, vec3 scale_b, const transform& Xb, float gap_sum,
                            float* dist5, float* pos15, float* normal15, int impl = 0, float margin_a = 0.0f, float margin_b = 0.0f,
                            const cvx::HullRef& hull_a = cvx::HullRef(), const cvx::HullRef& hull_b = cvx::HullRef(),
                            const cvx::SpeculativeWriter* spec = nullptr) {
    float ra, rb;
    vec3 pos[5], normal[5];
    // AABBs as the stand-alone NarrowPhase computes them (narrow_phase.py:1120-1150): tight support AABB +- the shape's gap;
    // only the bounding-sphere radius of the shape facing an infinite plane depends on them
    auto aabb = [&](int type, vec3 scale, const transform& X, const cvx::HullRef& hull, vec3& lo, vec3& hi) {
        if (type == GEO_PLANE) {
            lo = X.p - vec3(1.0e6f);
            hi = X.p + vec3(1.0e6f);
            return;
        }
        vec3 l, h;
        cvx::compute_tight_aabb_from_support(cvx_geom_hull(type, scale, hull), X.q, X.p, l, h);
        const vec3 g(0.5f * gap_sum);
        lo = l - g;
        hi = h + g;
    };
    vec3 lo_a, hi_a, lo_b, hi_b;
    aabb(type_a, scale_a, Xa, hull_a, lo_a, hi_a);
    aabb(type_b, scale_b, Xb, hull_b, lo_b, hi_b);
    int cnt = convex_pair(type_a, scale_a, Xa, margin_a, type_b, scale_b, Xb, margin_b, gap_sum, dist5, pos, normal, ra, rb, lo_a, hi_a, lo_b, hi_b, impl,
                          hull_a, hull_b, spec);
    for (int i = 0; i < cnt; ++i) {
        store3(pos15 + 3 * i, pos[i]);
        store3(normal15 + 3 * i, normal[i]);
    }
    return cnt;
}
}  // namespace orc

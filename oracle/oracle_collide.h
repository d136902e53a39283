// oracle_collide.h - TEST INFRASTRUCTURE ONLY.
// CPU restatement of the reference collision pipeline for primitive shapes:
//   compute_shape_aabbs            reference sim/collide.py:283-472
//   explicit broad phase           reference geometry/broad_phase_nxn.py:29-69, broad_phase_common.py:20-37
//   narrow_phase_primitive_kernel  reference geometry/narrow_phase.py:459-1014
//   analytic colliders             reference geometry/collision_primitive.py
//   write_contact                  reference sim/collide.py:166-254
//   deterministic sort key         reference geometry/contact_data.py:59-87
#pragma once
#include "oracle_convex.h"  // support maps for the generic AABB branch
#include <algorithm>
#include <vector>

#include "../include/newton_b200.h"
#include "oracle_math.h"

namespace orc {

enum { GEO_PLANE = 1, GEO_HFIELD = 2, GEO_SPHERE = 3, GEO_CAPSULE = 4, GEO_ELLIPSOID = 5, GEO_CYLINDER = 6, GEO_BOX = 7, GEO_MESH = 8, GEO_CONE = 9, GEO_CONVEX_MESH = 10 };
static const float MAXVAL = 1.0e10f;
static const float MINVAL = 1e-15f;
static const float CYLINDER_FLAT_MODE_COS = 0.92387953251128673848f;  // cos(22.5 deg), collision_primitive.py:44-45

inline vec3 normalize_with_norm(vec3 x, float& norm) {  // reference math/__init__.py:297-312
    norm = length(x);
    if (norm == 0.0f) return x;
    return x / norm;
}
inline float safe_div(float x, float y) { return x / (y != 0.0f ? y : 1e-15f); }

// ---- analytic colliders (reference geometry/collision_primitive.py) -------------------------
inline vec3 closest_segment_point(vec3 a, vec3 b, vec3 pt) {  // :48-53
    vec3 ab = b - a;
    float t = dot(pt - a, ab) / (dot(ab, ab) + 1e-6f);
    return a + clampf(t, 0.0f, 1.0f) * ab;
}
inline void collide_plane_sphere(vec3 n, vec3 plane_pos, vec3 sphere_pos, float r, float& dist, vec3& pos) {  // :100-107
    dist = dot(sphere_pos - plane_pos, n) - r;
    pos = sphere_pos - n * (r + 0.5f * dist);
}
inline void collide_sphere_sphere(vec3 pos1, float r1, vec3 pos2, float r2, float& dist, vec3& pos, vec3& n) {  // :110-140
    vec3 dir = pos2 - pos1;
    dist = length(dir);
    if (dist == 0.0f) n = vec3(1.f, 0.f, 0.f);
    else n = dir / dist;
    dist = dist - (r1 + r2);
    pos = pos1 + n * (r1 + 0.5f * dist);
}
inline void collide_sphere_capsule(vec3 sp, float sr, vec3 cp, vec3 axis, float cr, float chl, float& dist, vec3& pos,
                                   vec3& n) {  // :143-177
    vec3 segment = axis * chl;
    vec3 pt = closest_segment_point(cp - segment, cp + segment, sp);
    collide_sphere_sphere(sp, sr, pt, cr, dist, pos, n);
}
inline void collide_capsule_capsule(vec3 p1, vec3 a1, float r1, float hl1, vec3 p2, vec3 a2, float r2, float hl2,
                                    float dists[2], vec3 poss[2], vec3& normal) {  // :180-275
    dists[0] = dists[1] = MAXVAL;
    vec3 axis1 = a1 * hl1, axis2 = a2 * hl2;
    vec3 dif = p1 - p2;
    float ma = dot(axis1, axis1);
    float mb = -dot(axis1, axis2);
    float mc = dot(axis2, axis2);
    float u = -dot(axis1, dif);
    float v = dot(axis2, dif);
    float det = ma * mc - mb * mb;
    if (std::fabs(det) >= MINVAL) {
        float inv_det = 1.0f / det;
        float x1 = (mc * u - mb * v) * inv_det;
        float x2 = (ma * v - mb * u) * inv_det;
        if (x1 > 1.0f) { x1 = 1.0f; x2 = (v - mb) / mc; }
        else if (x1 < -1.0f) { x1 = -1.0f; x2 = (v + mb) / mc; }
        if (x2 > 1.0f) { x2 = 1.0f; x1 = clampf((u - mb) / ma, -1.0f, 1.0f); }
        else if (x2 < -1.0f) { x2 = -1.0f; x1 = clampf((u + mb) / ma, -1.0f, 1.0f); }
        vec3 vec1 = p1 + axis1 * x1;
        vec3 vec2 = p2 + axis2 * x2;
        collide_sphere_sphere(vec1, r1, vec2, r2, dists[0], poss[0], normal);
    } else {
        vec3 vec1 = p1 + axis1;
        float x2 = clampf((v - mb) / mc, -1.0f, 1.0f);
        vec3 vec2 = p2 + axis2 * x2;
        collide_sphere_sphere(vec1, r1, vec2, r2, dists[0], poss[0], normal);
        vec1 = p1 - axis1;
        x2 = clampf((v + mb) / mc, -1.0f, 1.0f);
        vec2 = p2 + axis2 * x2;
        vec3 n2;
        collide_sphere_sphere(vec1, r1, vec2, r2, dists[1], poss[1], n2);
    }
}
inline void collide_plane_ellipsoid(vec3 n, vec3 plane_pos, vec3 epos, const mat33& erot, vec3 size, float& dist, vec3& pos,
                                    vec3& normal) {  // :352-381
    vec3 sphere_support = -normalize(cw_mul(transpose(erot) * n, size));
    pos = epos + erot * cw_mul(sphere_support, size);
    dist = dot(n, pos - plane_pos);
    pos = pos - n * dist * 0.5f;
    normal = n;
}
inline void collide_plane_box(vec3 n, vec3 plane_pos, vec3 box_pos, const mat33& box_rot, vec3 box_size, float margin,
                              float dist[4], vec3 pos[4], vec3& normal) {  // :384-458
    float center_dist = dot(box_pos - plane_pos, n);
    for (int i = 0; i < 4; ++i) { dist[i] = MAXVAL; pos[i] = vec3(); }
    int ncontact = 0, worst_idx = 0;
    for (int i = 0; i < 8; ++i) {
        vec3 corner((i & 1) != 0 ? box_size.x : -box_size.x, (i & 2) != 0 ? box_size.y : -box_size.y,
                    (i & 4) != 0 ? box_size.z : -box_size.z);
        corner = box_rot * corner;
        float ldist = dot(n, corner);
        float cdist = center_dist + ldist;
        if (cdist > margin) continue;
        vec3 cpos = corner + box_pos - 0.5f * n * cdist;
        if (ncontact < 4) {
            dist[ncontact] = cdist;
            pos[ncontact] = cpos;
            if (ncontact == 0 || cdist > dist[worst_idx]) worst_idx = ncontact;
            ncontact += 1;
        } else if (cdist < dist[worst_idx]) {
            dist[worst_idx] = cdist;
            pos[worst_idx] = cpos;
            worst_idx = 0;
            if (dist[1] > dist[worst_idx]) worst_idx = 1;
            if (dist[2] > dist[worst_idx]) worst_idx = 2;
            if (dist[3] > dist[worst_idx]) worst_idx = 3;
        }
    }
    normal = n;
}
inline void collide_sphere_cylinder(vec3 sp, float sr, vec3 cp, vec3 axis, float cr, float chh, float& dist, vec3& pos,
                                    vec3& normal) {  // :461-531
    vec3 vec = sp - cp;
    float x = dot(vec, axis);
    vec3 a_proj = axis * x;
    vec3 p_proj = vec - a_proj;
    float p_proj_sqr = dot(p_proj, p_proj);
    bool collide_side = std::fabs(x) < chh;
    bool collide_cap = p_proj_sqr < (cr * cr);
    if (collide_side && collide_cap) {
        float dist_cap = chh - std::fabs(x);
        float dist_radius = cr - std::sqrt(p_proj_sqr);
        if (dist_cap < dist_radius) collide_side = false;
        else collide_cap = false;
    }
    if (collide_side) {
        collide_sphere_sphere(sp, sr, cp + a_proj, cr, dist, pos, normal);
    } else if (collide_cap) {
        vec3 pos_cap, plane_normal;
        if (x > 0.0f) { pos_cap = cp + axis * chh; plane_normal = axis; }
        else { pos_cap = cp - axis * chh; plane_normal = -axis; }
        collide_plane_sphere(plane_normal, pos_cap, sp, sr, dist, pos);
        normal = -plane_normal;
    } else {
        float inv_len = safe_div(1.0f, std::sqrt(p_proj_sqr));
        p_proj = p_proj * (cr * inv_len);
        vec3 cap_offset = axis * (signf(x) * chh);
        collide_sphere_sphere(sp, sr, cp + cap_offset + p_proj, 0.0f, dist, pos, normal);
    }
}
inline void collide_plane_cylinder(vec3 n, vec3 plane_pos, vec3 cpos, vec3 axis, float radius, float half_height,
                                   float contact_dist[4], vec3 contact_pos[4], vec3& normal) {  // :534-683
    for (int i = 0; i < 4; ++i) { contact_dist[i] = MAXVAL; contact_pos[i] = vec3(); }
    float dot_na = dot(n, axis);
    if (dot_na > 0.0f) { axis = -axis; dot_na = -dot_na; }
    vec3 cap_center = cpos + axis * half_height;
    vec3 perp_align = -n + axis * dot_na;
    float perp_align_len_sq = dot(perp_align, perp_align);
    bool has_align = perp_align_len_sq > 1e-10f;
    if (has_align) perp_align = perp_align * (1.0f / std::sqrt(perp_align_len_sq));
    float abs_dot = -dot_na;
    bool in_flat_surface_mode = abs_dot >= CYLINDER_FLAT_MODE_COS;
    vec3 perp_fixed;
    if (in_flat_surface_mode || !has_align) {
        vec3 ref(1.f, 0.f, 0.f);
        if (std::fabs(dot(axis, ref)) > 0.9f) ref = vec3(0.f, 1.f, 0.f);
        perp_fixed = ref - axis * dot(axis, ref);
        perp_fixed = normalize(perp_fixed);
    }
    vec3 deepest_perp = has_align ? perp_align : perp_fixed;
    vec3 deepest_pt = cap_center + deepest_perp * radius;
    float deepest_d = dot(deepest_pt - plane_pos, n);
    vec3 deepest_pos = deepest_pt - n * (deepest_d * 0.5f);
    contact_dist[0] = deepest_d;
    contact_pos[0] = deepest_pos;
    int ncontact = 1;
    float merge_threshold = 0.01f * maxf(radius, half_height);
    float merge_threshold_sq = merge_threshold * merge_threshold;
    if (in_flat_surface_mode) {
        vec3 u_fixed = perp_fixed * radius;
        vec3 v_fixed = cross(axis, perp_fixed) * radius;
        const float c120 = -0.5f, s120 = 0.8660254f;
        vec3 pts[3] = {cap_center + u_fixed, cap_center + c120 * u_fixed + s120 * v_fixed,
                       cap_center + c120 * u_fixed - s120 * v_fixed};
        for (int k = 0; k < 3; ++k) {
            float d = dot(pts[k] - plane_pos, n);
            vec3 p = pts[k] - n * (d * 0.5f);
            if (ncontact < 4 && length_sq(p - deepest_pos) > merge_threshold_sq) {
                contact_dist[ncontact] = d;
                contact_pos[ncontact] = p;
                ncontact += 1;
            }
        }
    } else {
        vec3 perp_roll = has_align ? perp_align : perp_fixed;
        vec3 u = perp_roll * radius;
        vec3 v = cross(axis, perp_roll) * radius;
        vec3 pt = cpos - axis * half_height + u;
        float d = dot(pt - plane_pos, n);
        vec3 pos = pt - n * (d * 0.5f);
        if (ncontact < 4 && length_sq(pos - deepest_pos) > merge_threshold_sq) {
            contact_dist[ncontact] = d;
            contact_pos[ncontact] = pos;
            ncontact += 1;
        }
        vec3 pt_pos_v = cap_center + v;
        float d_pos_v = dot(pt_pos_v - plane_pos, n);
        vec3 pt_neg_v = cap_center - v;
        float d_neg_v = dot(pt_neg_v - plane_pos, n);
        bool use_pos_v = d_pos_v <= d_neg_v;
        pt = use_pos_v ? pt_pos_v : pt_neg_v;
        d = use_pos_v ? d_pos_v : d_neg_v;
        pos = pt - n * (d * 0.5f);
        if (ncontact < 4 && length_sq(pos - deepest_pos) > merge_threshold_sq) {
            contact_dist[ncontact] = d;
            contact_pos[ncontact] = pos;
            ncontact += 1;
        }
    }
    normal = n;
}
inline void collide_sphere_box(vec3 sp, float sr, vec3 box_pos, const mat33& box_rot, vec3 box_size, float& dist, vec3& position,
                               vec3& normal) {  // :1176-1230
    vec3 center = transpose(box_rot) * (sp - box_pos);
    vec3 clamped = vmax(-box_size, vmin(box_size, center));
    float d;
    vec3 clamped_dir = normalize_with_norm(clamped - center, d);
    vec3 pos;
    if (d <= 1e-6f) {
        float closest = 2.0f * (box_size.x + box_size.y + box_size.z);
        int k = 0;
        for (int i = 0; i < 6; ++i) {
            float face_dist = std::fabs(((i % 2) ? 1.0f : -1.0f) * box_size[i / 2] - center[i / 2]);
            if (closest > face_dist) { closest = face_dist; k = i; }
        }
        vec3 nearest(0.f);
        nearest[k / 2] = (k % 2) ? -1.0f : 1.0f;
        pos = center + nearest * (sr - closest) / 2.0f;
        normal = box_rot * nearest;
        dist = -closest - sr;
    } else {
        vec3 deepest = center + clamped_dir * sr;
        pos = 0.5f * (clamped + deepest);
        normal = box_rot * clamped_dir;
        dist = d - sr;
    }
    position = box_pos + box_rot * pos;
}

// ---- pipeline ---------------------------------------------------------------------------------
struct ShapeGeom {  // geom_data / geom_xform written by compute_shape_aabbs
    vec3 scale;
    float margin;
    transform X_ws;
    vec3 aabb_lower, aabb_upper;
    // compute_shape_velocities (speculative contacts only): shape-origin velocity, angular velocity, velocity-extended gap,
    // displacement over the collision-update interval
    vec3 linear_velocity, angular_velocity, displacement;
    float search_gap = 0.0f;
};

// CollisionPipeline(speculative_config=...) + collide(dt=...) (reference sim/collide.py:1076-1102, 1823-1836)
struct SpeculativeParams {
    bool active = false;   // speculative_active: enabled and dt > 0 and max_speculative_extension > 0
    bool enabled = false;  // the pipeline was built with a speculative_config (the writer is write_contact_speculative even when inactive)
    const float* body_qd = nullptr;
    float dt = 0.0f;
    float max_extension = 0.0f;
};

// reference sim/collide.py:283-472 (primitive branches)
inline void compute_shape_aabbs(const nb2_model_desc& m, const float* body_q, std::vector<ShapeGeom>& out) {
    out.resize(m.shape_count);
    for (int sid = 0; sid < m.shape_count; ++sid) {
        int rigid_id = m.shape_body[sid];
        int geo_type = m.shape_type[sid];
        transform X_ws = transform::load(m.shape_transform + 7 * sid);
        if (rigid_id != -1) X_ws = transform::load(body_q + 7 * rigid_id) * X_ws;
        vec3 pos = X_ws.p;
        quat orientation = X_ws.q;
        float margin = m.shape_margin[sid];
        float effective_gap = margin + m.shape_gap[sid];
        vec3 margin_vec(effective_gap, effective_gap, effective_gap);
        vec3 scale = load3(m.shape_scale + 3 * sid);
        bool is_infinite_plane = (geo_type == GEO_PLANE) && (scale.x == 0.0f && scale.y == 0.0f);
        vec3 geom_scale = scale;
        vec3 lo, hi;
        if (is_infinite_plane) {
            vec3 normal = quat_rotate(orientation, vec3(0.f, 0.f, 1.f));
            const float HALF_SPACE_EXTENT = 1.0e6f;
            vec3 half_extents(HALF_SPACE_EXTENT, HALF_SPACE_EXTENT, HALF_SPACE_EXTENT);
            lo = pos - half_extents - margin_vec;
            hi = pos + half_extents + margin_vec;
            for (int i = 0; i < 3; ++i) {
                float n_i = normal[i];
                if (std::fabs(n_i) > 0.5f) {
                    float lateral = std::fabs(normal[(i + 1) % 3]) + std::fabs(normal[(i + 2) % 3]);
                    float rise = lateral * HALF_SPACE_EXTENT / std::fabs(n_i);
                    if (n_i > 0.0f) hi[i] = minf(hi[i], pos[i] + rise + effective_gap);
                    else lo[i] = maxf(lo[i], pos[i] - rise - effective_gap);
                }
            }
        } else if (geo_type == GEO_SPHERE) {
            vec3 he(scale.x, scale.x, scale.x);
            lo = pos - he - margin_vec;
            hi = pos + he + margin_vec;
        } else if (geo_type == GEO_BOX) {
            vec3 r0 = quat_rotate(orientation, vec3(1.f, 0.f, 0.f));
            vec3 r1 = quat_rotate(orientation, vec3(0.f, 1.f, 0.f));
            vec3 r2 = quat_rotate(orientation, vec3(0.f, 0.f, 1.f));
            vec3 he(std::fabs(r0.x) * scale.x + std::fabs(r1.x) * scale.y + std::fabs(r2.x) * scale.z,
                    std::fabs(r0.y) * scale.x + std::fabs(r1.y) * scale.y + std::fabs(r2.y) * scale.z,
                    std::fabs(r0.z) * scale.x + std::fabs(r1.z) * scale.y + std::fabs(r2.z) * scale.z);
            lo = pos - he - margin_vec;
            hi = pos + he + margin_vec;
        } else if (geo_type == GEO_CAPSULE) {
            vec3 axis = quat_rotate(orientation, vec3(0.f, 0.f, 1.f));
            vec3 he = vec3(scale.x, scale.x, scale.x) + vabs(axis) * scale.y;
            lo = pos - he - margin_vec;
            hi = pos + he + margin_vec;
        } else if (geo_type == GEO_CYLINDER) {
            float radius = scale.x, half_height = scale.y, barrel_radius = scale.z;
            if (barrel_radius >= half_height && barrel_radius > 0.0f)
                radius += (half_height * half_height) /
                          (barrel_radius + std::sqrt(barrel_radius * barrel_radius - half_height * half_height));
            vec3 r0 = quat_rotate(orientation, vec3(1.f, 0.f, 0.f));
            vec3 r1 = quat_rotate(orientation, vec3(0.f, 1.f, 0.f));
            vec3 r2 = quat_rotate(orientation, vec3(0.f, 0.f, 1.f));
            vec3 he(radius * std::sqrt(r0.x * r0.x + r1.x * r1.x) + half_height * std::fabs(r2.x),
                    radius * std::sqrt(r0.y * r0.y + r1.y * r1.y) + half_height * std::fabs(r2.y),
                    radius * std::sqrt(r0.z * r0.z + r1.z * r1.z) + half_height * std::fabs(r2.z));
            lo = pos - he - margin_vec;
            hi = pos + he + margin_vec;
        } else if (geo_type == GEO_CONVEX_MESH || geo_type == GEO_MESH) {
            // has_local_aabb (collide.py:348, 420-444): the builder's scaled local AABB, rotated into the world frame
            vec3 local_lo = load3(m.shape_collision_aabb_lower + 3 * sid), local_hi = load3(m.shape_collision_aabb_upper + 3 * sid);
            vec3 center = (local_lo + local_hi) * 0.5f;
            vec3 half = (local_hi - local_lo) * 0.5f;
            vec3 world_center = quat_rotate(orientation, center) + pos;
            vec3 r0 = quat_rotate(orientation, vec3(1.f, 0.f, 0.f));
            vec3 r1 = quat_rotate(orientation, vec3(0.f, 1.f, 0.f));
            vec3 r2 = quat_rotate(orientation, vec3(0.f, 0.f, 1.f));
            vec3 world_half(std::fabs(r0.x) * half.x + std::fabs(r1.x) * half.y + std::fabs(r2.x) * half.z,
                            std::fabs(r0.y) * half.x + std::fabs(r1.y) * half.y + std::fabs(r2.y) * half.z,
                            std::fabs(r0.z) * half.x + std::fabs(r1.z) * half.y + std::fabs(r2.z) * half.z);
            lo = world_center - world_half - margin_vec;
            hi = world_center + world_half + margin_vec;
        } else if (geo_type == GEO_CONE || geo_type == GEO_PLANE) {
            // generic branch (collide.py:447-468): compute_tight_aabb_from_support;
            // finite planes carry HALF extents in geom_scale (collide.py:452-453)
            if (geo_type == GEO_PLANE) geom_scale = vec3(scale.x * 0.5f, scale.y * 0.5f, 0.0f);
            cvx::GenericShapeData sd;
            sd.shape_type = geo_type;
            sd.scale = geom_scale;
            vec3 l, h;
            cvx::compute_tight_aabb_from_support(sd, orientation, pos, l, h);
            lo = l - margin_vec;
            hi = h + margin_vec;
        } else if (geo_type == GEO_ELLIPSOID) {
            // tight AABB from the support map (collide.py:447-468): extent_i = |R_i * diag(scale)|
            mat33 R = quat_to_matrix(orientation);
            vec3 he;
            for (int i = 0; i < 3; ++i)
                he[i] = std::sqrt(R.m[i][0] * scale.x * R.m[i][0] * scale.x + R.m[i][1] * scale.y * R.m[i][1] * scale.y +
                                  R.m[i][2] * scale.z * R.m[i][2] * scale.z);
            lo = pos - he - margin_vec;
            hi = pos + he + margin_vec;
        } else {
            // finite plane / unsupported: conservative bounding sphere
            float r = m.shape_collision_radius[sid];
            vec3 he(r, r, r);
            if (geo_type == GEO_PLANE) geom_scale = vec3(scale.x * 0.5f, scale.y * 0.5f, 0.0f);
            lo = pos - he - margin_vec;
            hi = pos + he + margin_vec;
        }
        out[sid].aabb_lower = lo;
        out[sid].aabb_upper = hi;
        out[sid].scale = geom_scale;
        out[sid].margin = margin;
        out[sid].X_ws = X_ws;
    }
}

inline bool check_aabb_overlap(vec3 l1, vec3 u1, vec3 l2, vec3 u2) {  // broad_phase_common.py:20-37 with cutoffs 0
    const float c = 0.0f + 0.0f;
    return l1.x <= u2.x + c && u1.x >= l2.x - c && l1.y <= u2.y + c && u1.y >= l2.y - c && l1.z <= u2.z + c &&
           u1.z >= l2.z - c;
}

// reference sim/collide.py:475-541 (compute_shape_velocities): runs after compute_shape_aabbs when speculative contacts are active
inline void compute_shape_velocities(const nb2_model_desc& m, const float* body_q, const SpeculativeParams& sp, std::vector<ShapeGeom>& g) {
    for (int sid = 0; sid < m.shape_count; ++sid) {
        ShapeGeom& o = g[sid];
        const int body_id = m.shape_body[sid];
        if (body_id == -1) {
            o.linear_velocity = vec3(0.f);
            o.angular_velocity = vec3(0.f);
            o.search_gap = m.shape_gap[sid];
            o.displacement = vec3(0.f);
            continue;
        }
        transform X_wb = transform::load(body_q + 7 * body_id);
        transform X_ws = X_wb * transform::load(m.shape_transform + 7 * sid);
        vec3 shape_origin_world = X_ws.p;
        vec3 com_world = transform_point(X_wb, load3(m.body_com + 3 * body_id));
        spatial twist = spatial::load(sp.body_qd + 6 * body_id);
        vec3 com_velocity = twist.top, angular_velocity = twist.bot;
        vec3 shape_origin_velocity = com_velocity + cross(angular_velocity, shape_origin_world - com_world);
        o.linear_velocity = shape_origin_velocity;
        o.angular_velocity = angular_velocity;
        vec3 local_lower = load3(m.shape_collision_aabb_lower + 3 * sid), local_upper = load3(m.shape_collision_aabb_upper + 3 * sid);
        vec3 al = vabs(local_lower), au = vabs(local_upper);
        vec3 furthest(maxf(al.x, au.x), maxf(al.y, au.y), maxf(al.z, au.z));
        float angular_radius = maxf(length(furthest), m.shape_collision_radius[sid]);
        float angular_speed_bound = length(angular_velocity) * angular_radius;
        float search_extension = minf((length(shape_origin_velocity) + angular_speed_bound) * sp.dt, sp.max_extension);
        o.search_gap = m.shape_gap[sid] + search_extension;
        vec3 displacement = shape_origin_velocity * sp.dt;
        float angular_extension = angular_speed_bound * sp.dt;
        o.displacement = displacement;
        float ae = minf(angular_extension, sp.max_extension);
        o.aabb_lower = o.aabb_lower - vec3(ae, ae, ae);
        o.aabb_upper = o.aabb_upper + vec3(ae, ae, ae);
    }
}

// broad_phase_common.py:41-80 (check_aabb_overlap_moving, cutoffs 0): swept overlap over the relative displacement
inline bool check_aabb_overlap_moving(const ShapeGeom& a, const ShapeGeom& b) {
    const float cutoff_combined = 0.0f + 0.0f;
    vec3 rel = a.displacement - b.displacement;
    float enter = 0.0f, exit_time = 1.0f;
    for (int axis = 0; axis < 3; ++axis) {
        float lower1 = a.aabb_lower[axis], upper1 = a.aabb_upper[axis];
        float lower2 = b.aabb_lower[axis] - cutoff_combined, upper2 = b.aabb_upper[axis] + cutoff_combined;
        float delta = rel[axis];
        if (delta == 0.0f) {
            if (lower1 > upper2 || upper1 < lower2) return false;
        } else {
            float axis_enter = (lower2 - upper1) / delta, axis_exit = (upper2 - lower1) / delta;
            if (axis_enter > axis_exit) std::swap(axis_enter, axis_exit);
            enter = maxf(enter, axis_enter);
            exit_time = minf(exit_time, axis_exit);
            if (enter > exit_time) return false;
        }
    }
    return true;
}

// contact_data.py:187-233: prepare_speculative_contact + contact_passes_speculative_gap_check.  `gap_sum` is the AUTHORED pair gap.
inline bool speculative_admit(const ShapeGeom& A, const ShapeGeom& B, vec3 center, vec3 n_ab, float distance, float radius_eff_a,
                              float radius_eff_b, float margin_a, float margin_b, float gap_sum, const SpeculativeParams& sp) {
    vec3 normal = normalize(n_ab);
    vec3 point_a = center - normal * (0.5f * distance + radius_eff_a);
    vec3 point_b = center + normal * (0.5f * distance + radius_eff_b);
    float total_separation_needed = radius_eff_a + radius_eff_b + margin_a + margin_b;
    float physical_separation = dot(point_b - point_a, normal) - total_separation_needed;
    if (physical_separation <= gap_sum) return true;
    // compute_contact_approach_speed / compute_contact_predictive_score (:92-135); geom_transform holds the shape origins
    vec3 velocity_a = A.linear_velocity + cross(A.angular_velocity, point_a - A.X_ws.p);
    vec3 velocity_b = B.linear_velocity + cross(B.angular_velocity, point_b - B.X_ws.p);
    float approach_speed = maxf(-dot(velocity_b - velocity_a, normal), 0.0f);
    float extension = minf(approach_speed * sp.dt, sp.max_extension);
    return extension - physical_separation >= 0.0f;
}

struct RawContact {  // ContactData + output of write_contact
    int shape0, shape1;
    vec3 point0, point1, offset0, offset1, normal;
    float margin0, margin1;
    int64_t key;
};

inline int64_t make_contact_sort_key(int shape_a, int shape_b, int sub) {  // contact_data.py:59-87
    return ((int64_t(shape_a) & 0xFFFFF) << 43) | ((int64_t(shape_b) & 0xFFFFF) << 23) | (int64_t(sub) & 0x7FFFFF);
}

// reference sim/collide.py:210-254 + :166-207; returns false if dropped by the gap test
inline bool write_contact(const nb2_model_desc& m, const float* body_q, int shape_a, int shape_b, vec3 center, vec3 n_ab,
                          float distance, float radius_eff_a, float radius_eff_b, float margin_a, float margin_b,
                          int sort_sub_key, bool check_gap, RawContact& out) {
    float total_separation_needed = radius_eff_a + radius_eff_b + margin_a + margin_b;
    vec3 n = normalize(n_ab);
    vec3 a_w = center - n * (0.5f * distance + radius_eff_a);
    vec3 b_w = center + n * (0.5f * distance + radius_eff_b);
    vec3 diff = b_w - a_w;
    float dist = dot(diff, n);
    float d = dist - total_separation_needed;
    float contact_gap = m.shape_gap[shape_a] + m.shape_gap[shape_b];
    if (check_gap && d > contact_gap) return false;
    out.shape0 = shape_a;
    out.shape1 = shape_b;
    int body0 = m.shape_body[shape_a];
    int body1 = m.shape_body[shape_b];
    transform X_bw_a = body0 == -1 ? transform_identity() : transform_inverse(transform::load(body_q + 7 * body0));
    transform X_bw_b = body1 == -1 ? transform_identity() : transform_inverse(transform::load(body_q + 7 * body1));
    out.point0 = transform_point(X_bw_a, a_w);
    out.point1 = transform_point(X_bw_b, b_w);
    float offset_mag_a = radius_eff_a + margin_a;
    float offset_mag_b = radius_eff_b + margin_b;
    out.offset0 = transform_vector(X_bw_a, offset_mag_a * n);
    out.offset1 = transform_vector(X_bw_b, -offset_mag_b * n);
    out.normal = n;
    out.margin0 = offset_mag_a;
    out.margin1 = offset_mag_b;
    out.key = make_contact_sort_key(shape_a, shape_b, sort_sub_key);
    return true;
}

// contact_data.py:138-156
inline bool gap_check_precomputed(vec3 center, float distance, float radius_eff_a, float radius_eff_b, vec3 n,
                                  float total_separation_needed, float gap_sum) {
    vec3 a_w = center - n * (0.5f * distance + radius_eff_a);
    vec3 b_w = center + n * (0.5f * distance + radius_eff_b);
    vec3 diff = b_w - a_w;
    float dist = dot(diff, n);
    float d = dist - total_separation_needed;
    return d <= gap_sum;
}

// Analytic dispatch of narrow_phase_primitive_kernel (narrow_phase.py:657-864).  Returns true when the pair
// is one of the analytic combinations (contacts may still be zero); false -> GJK/MPR queue.
inline bool primitive_pair(int type_a, vec3 scale_a, const transform& X_a, int type_b, vec3 scale_b, const transform& X_b,
                           float plane_box_margin, float dist[4], vec3 pos[4], vec3& normal) {
    for (int i = 0; i < 4; ++i) { dist[i] = MAXVAL; pos[i] = vec3(); }
    normal = vec3();
    vec3 pos_a = X_a.p, pos_b = X_b.p;
    quat quat_a = X_a.q, quat_b = X_b.q;
    bool is_plane_a = type_a == GEO_PLANE;
    bool is_sphere_a = type_a == GEO_SPHERE, is_sphere_b = type_b == GEO_SPHERE;
    bool is_capsule_a = type_a == GEO_CAPSULE, is_capsule_b = type_b == GEO_CAPSULE;
    bool is_ellipsoid_b = type_b == GEO_ELLIPSOID, is_cylinder_b = type_b == GEO_CYLINDER, is_box_b = type_b == GEO_BOX;
    bool use_plane_cylinder = is_plane_a && is_cylinder_b;
    if (use_plane_cylinder && scale_b.z > 0.0f) {
        vec3 pn = quat_rotate(quat_a, vec3(0.f, 0.f, 1.f));
        vec3 ca = quat_rotate(quat_b, vec3(0.f, 0.f, 1.f));
        use_plane_cylinder = std::fabs(dot(pn, ca)) * scale_b.z >= scale_b.y;
    }
    if (is_plane_a && is_sphere_b) {
        vec3 pn = quat_rotate(quat_a, vec3(0.f, 0.f, 1.f));
        collide_plane_sphere(pn, pos_a, pos_b, scale_b.x, dist[0], pos[0]);
        normal = pn;
    } else if (is_plane_a && is_ellipsoid_b) {
        vec3 pn = quat_rotate(quat_a, vec3(0.f, 0.f, 1.f));
        collide_plane_ellipsoid(pn, pos_a, pos_b, quat_to_matrix(quat_b), scale_b, dist[0], pos[0], normal);
    } else if (is_plane_a && is_box_b) {
        vec3 pn = quat_rotate(quat_a, vec3(0.f, 0.f, 1.f));
        collide_plane_box(pn, pos_a, pos_b, quat_to_matrix(quat_b), scale_b, plane_box_margin, dist, pos, normal);
    } else if (is_sphere_a && is_sphere_b) {
        collide_sphere_sphere(pos_a, scale_a.x, pos_b, scale_b.x, dist[0], pos[0], normal);
    } else if (is_plane_a && is_capsule_b) {
        vec3 pn = quat_rotate(quat_a, vec3(0.f, 0.f, 1.f));
        vec3 ca = quat_rotate(quat_b, vec3(0.f, 0.f, 1.f));
        vec3 segment = ca * scale_b.y;
        collide_plane_sphere(pn, pos_a, pos_b + segment, scale_b.x, dist[0], pos[0]);
        collide_plane_sphere(pn, pos_a, pos_b - segment, scale_b.x, dist[1], pos[1]);
        normal = pn;
    } else if (use_plane_cylinder) {
        vec3 pn = quat_rotate(quat_a, vec3(0.f, 0.f, 1.f));
        vec3 ca = quat_rotate(quat_b, vec3(0.f, 0.f, 1.f));
        collide_plane_cylinder(pn, pos_a, pos_b, ca, scale_b.x, scale_b.y, dist, pos, normal);
    } else if (is_sphere_a && is_capsule_b) {
        vec3 ca = quat_rotate(quat_b, vec3(0.f, 0.f, 1.f));
        collide_sphere_capsule(pos_a, scale_a.x, pos_b, ca, scale_b.x, scale_b.y, dist[0], pos[0], normal);
    } else if (is_capsule_a && is_capsule_b) {
        vec3 aa = quat_rotate(quat_a, vec3(0.f, 0.f, 1.f));
        vec3 ab = quat_rotate(quat_b, vec3(0.f, 0.f, 1.f));
        collide_capsule_capsule(pos_a, aa, scale_a.x, scale_a.y, pos_b, ab, scale_b.x, scale_b.y, dist, pos, normal);
    } else if (is_sphere_a && is_cylinder_b && scale_b.z == 0.0f) {
        vec3 ca = quat_rotate(quat_b, vec3(0.f, 0.f, 1.f));
        collide_sphere_cylinder(pos_a, scale_a.x, pos_b, ca, scale_b.x, scale_b.y, dist[0], pos[0], normal);
    } else if (is_sphere_a && is_box_b) {
        collide_sphere_box(pos_a, scale_a.x, pos_b, quat_to_matrix(quat_b), scale_b, dist[0], pos[0], normal);
    } else {
        return false;
    }
    return true;
}

// Routing test at narrow_phase.py:642-655: pairs sent to GJK/MPR before the analytic chain.
inline bool routes_to_gjk_early(int type_a, int type_b) {
    return type_a >= GEO_ELLIPSOID || type_b == GEO_CONE || (type_a == GEO_CAPSULE && type_b > GEO_CAPSULE);
}

struct CollideResult {
    std::vector<RawContact> contacts;            // in reference CPU emission order for candidate order = pair order
    std::vector<std::pair<int, int>> gjk_pairs;  // (shape_a, shape_b) type-sorted pairs for the GJK/MPR kernel
    std::vector<std::pair<int, int>> mesh_plane_pairs;  // (mesh, plane) as narrow_phase.py:620-631 stores them
    int unsupported_mesh_pairs = 0;              // mesh-mesh / mesh-convex / mesh-finite-plane routes (BVH / SDF paths: out of scope)
    int candidate_count = 0;
};

// narrow_phase_process_mesh_plane_contacts_kernel (narrow_phase.py:1761-1861), reduce_contacts=False: every mesh vertex within
// gap + margin of an INFINITE plane becomes one contact (shape_a = mesh, normal from the mesh to the plane, sub key = vertex
// index), written through write_contact with its own gap test (output_index = -1, collide.py:210-254).  Mesh vertices are the
// shape's range of model.hull_points (the reference reads wp.Mesh.points behind shape_source).
inline void mesh_plane_contacts(const nb2_model_desc& m, const float* body_q, const std::vector<ShapeGeom>& geom, CollideResult& res) {
    for (const auto& pr : res.mesh_plane_pairs) {
        const int mesh_shape = pr.first, plane_shape = pr.second;
        const int v0 = m.shape_hull_start[mesh_shape], nv = m.shape_hull_count[mesh_shape];
        const transform X_mesh_ws = geom[mesh_shape].X_ws, X_plane_ws = geom[plane_shape].X_ws;
        const transform X_plane_sw = transform_inverse(X_plane_ws);
        const vec3 plane_normal = transform_vector(X_plane_ws, vec3(0.f, 0.f, 1.f));
        const vec3 mesh_scale = geom[mesh_shape].scale;
        const float margin_mesh = geom[mesh_shape].margin, margin_plane = geom[plane_shape].margin;
        const float total_margin_offset = margin_mesh + margin_plane;
        const float gap_sum = m.shape_gap[mesh_shape] + m.shape_gap[plane_shape];
        for (int vi = 0; vi < nv; ++vi) {
            const vec3 vertex_local = cw_mul(load3(m.hull_points + 3 * size_t(v0 + vi)), mesh_scale);
            const vec3 vertex_world = transform_point(X_mesh_ws, vertex_local);
            const vec3 in_plane = transform_point(X_plane_sw, vertex_world);
            const vec3 point_on_plane = transform_point(X_plane_ws, vec3(in_plane.x, in_plane.y, 0.0f));
            const vec3 diff = vertex_world - point_on_plane;
            const float distance = dot(diff, plane_normal);
            if (distance < gap_sum + total_margin_offset) {
                const vec3 contact_pos = (vertex_world + point_on_plane) * 0.5f;
                RawContact rc;
                if (write_contact(m, body_q, mesh_shape, plane_shape, contact_pos, -plane_normal, distance, 0.0f, 0.0f, margin_mesh,
                                  margin_plane, vi, true, rc))
                    res.contacts.push_back(rc);
            }
        }
    }
}

// compute_shape_aabbs + explicit broad phase + primitive narrow phase.  Candidate pairs are visited in
// shape_contact_pairs order (the reference CPU device visits them in (t mod 256, t div 256) order, which only
// permutes slots; compare under the deterministic sort key).
inline void collide_primitives(const nb2_model_desc& m, const float* body_q, CollideResult& res,
                               const SpeculativeParams& sp = SpeculativeParams()) {
    std::vector<ShapeGeom> geom;
    compute_shape_aabbs(m, body_q, geom);
    if (sp.active) compute_shape_velocities(m, body_q, sp, geom);
    res.contacts.clear();
    res.gjk_pairs.clear();
    res.mesh_plane_pairs.clear();
    res.unsupported_mesh_pairs = 0;
    res.candidate_count = 0;
    for (int t = 0; t < m.shape_pair_count; ++t) {
        int s1 = m.shape_contact_pairs[2 * t + 0], s2 = m.shape_contact_pairs[2 * t + 1];
        // is_shape_pair_immovable_filtered: include_static_kinematic_pairs defaults to True -> never filtered
        if (sp.active ? !check_aabb_overlap_moving(geom[s1], geom[s2])
                      : !check_aabb_overlap(geom[s1].aabb_lower, geom[s1].aabb_upper, geom[s2].aabb_lower, geom[s2].aabb_upper))
            continue;
        res.candidate_count += 1;
        int shape_a = s1, shape_b = s2;
        if (shape_a == shape_b || shape_a < 0 || shape_b < 0) continue;
        int type_a = m.shape_type[shape_a], type_b = m.shape_type[shape_b];
        if (type_a > type_b) { std::swap(shape_a, shape_b); std::swap(type_a, type_b); }
        if (type_a == GEO_MESH || type_b == GEO_MESH) {  // mesh routing (narrow_phase.py:594-640), ahead of the analytic chain
            const vec3 sa = load3(m.shape_scale + 3 * shape_a);
            const bool infinite_plane_a = type_a == GEO_PLANE && sa.x == 0.0f && sa.y == 0.0f;
            if (infinite_plane_a && type_b == GEO_MESH) res.mesh_plane_pairs.emplace_back(shape_b, shape_a);
            else res.unsupported_mesh_pairs += 1;
            continue;
        }
        if (routes_to_gjk_early(type_a, type_b)) { res.gjk_pairs.emplace_back(shape_a, shape_b); continue; }
        const ShapeGeom& A = geom[shape_a];
        const ShapeGeom& B = geom[shape_b];
        float margin_offset_a = A.margin, margin_offset_b = B.margin;
        // the kernel's shape_gap argument is the velocity-extended search gap when speculation is active (collide.py:1832, 2010);
        // the admission test below uses the authored gaps (writer_data.shape_gap, narrow_phase.py:885-888)
        const float base_gap_sum = m.shape_gap[shape_a] + m.shape_gap[shape_b];
        float gap_sum = sp.active ? A.search_gap + B.search_gap : base_gap_sum;
        float radius_eff_a = 0.f, radius_eff_b = 0.f;
        if (type_a == GEO_SPHERE || type_a == GEO_CAPSULE) radius_eff_a = A.scale.x;
        if (type_b == GEO_SPHERE || type_b == GEO_CAPSULE) radius_eff_b = B.scale.x;
        float dist[4];
        vec3 pos[4], normal;
        bool analytic = primitive_pair(type_a, A.scale, A.X_ws, type_b, B.scale, B.X_ws,
                                       gap_sum + margin_offset_a + margin_offset_b, dist, pos, normal);
        int num_contacts = 0;
        for (int i = 0; i < 4; ++i) num_contacts += dist[i] < MAXVAL;
        if (num_contacts > 0) {
            float total_separation_needed = radius_eff_a + radius_eff_b + margin_offset_a + margin_offset_b;
            vec3 nn = normalize(normal);
            for (int i = 0; i < 4; ++i) {
                if (!(dist[i] < MAXVAL)) continue;
                if (sp.enabled) {
                    if (!speculative_admit(A, B, pos[i], normal, dist[i], radius_eff_a, radius_eff_b, margin_offset_a, margin_offset_b,
                                           base_gap_sum, sp))
                        continue;
                } else if (!gap_check_precomputed(pos[i], dist[i], radius_eff_a, radius_eff_b, nn, total_separation_needed, gap_sum))
                    continue;
                RawContact rc;
                write_contact(m, body_q, shape_a, shape_b, pos[i], normal, dist[i], radius_eff_a, radius_eff_b,
                              margin_offset_a, margin_offset_b, i, false, rc);
                res.contacts.push_back(rc);
            }
            continue;
        }
        if (analytic) continue;
        res.gjk_pairs.emplace_back(shape_a, shape_b);
    }
}

}  // namespace orc

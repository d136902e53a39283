// oracle_convex.h - TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// INDEPENDENT restatement of the reference's generic convex-convex contact path, written from the reference sources and
// NOT from newton_b200/csrc/nb2_convex.cuh: one C++ function per Warp function, serial, orc:: math (oracle_math.h),
// compiled with -ffp-contract=off.  tests/test_abi_and_host.py compares it with the host compilation of the product's
// single-source routine bit for bit on randomised pairs of every shape-type combination; the CPU oracle's collide
// pipeline (oracle_gjk.h) uses THIS file, so the GPU parity tests compare the CUDA kernels with an independent translation.
//
//   support maps                       geometry/support_function.py:121-352, 393-446, 449-461
//   minkowski_support / geometric_center  geometry/mpr.py:76-150, support_function.py:541-597
//   solve_mpr_core                     geometry/mpr.py:188-403   (XenoCollide, zlib licence, altered upstream)
//   closest_* / solve_closest_distance_core  geometry/simplex_solver.py:83-495
//   multicontact helpers + build_manifold    geometry/multicontact.py:27-956
//   solve_convex_multi_contact         geometry/collision_convex.py:107-231
//   compute_gjk_mpr_contacts           geometry/collision_core.py:337-450
//   post_process_axial_on_discrete_contact   geometry/collision_core.py:174-277
//   compute_tight_aabb_from_support    geometry/collision_core.py:454-548
//   infinite plane handling            geometry/narrow_phase.py:1098-1165, collision_core.py:552-626, 640-682
//   orthonormal_basis                  math/__init__.py:236-275
#pragma once
#include <cstdint>

#include "oracle_math.h"

namespace orc {
namespace cvx {

enum { T_PLANE = 1, T_SPHERE = 3, T_CAPSULE = 4, T_ELLIPSOID = 5, T_CYLINDER = 6, T_BOX = 7, T_CONE = 9, T_CONVEX_MESH = 10 };

struct GenericShapeData {  // support_function.py:107-118
    int shape_type = 0;
    vec3 scale;
    vec3 center;  // zero for every primitive (_shape_center, support_function.py:449-461); local AABB centre for CONVEX_MESH
    // what `auxiliary` (the packed wp.Mesh pointer) leads to for CONVEX_MESH: mesh.points, unscaled
    const float* mesh_points = nullptr;
    int mesh_point_count = 0;
};

struct vec2 {
    float x = 0.f, y = 0.f;
    vec2() {}
    vec2(float a, float b) : x(a), y(b) {}
};
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator*(float s, vec2 a) { return vec2(a.x * s, a.y * s); }
inline float length_sq(vec2 a) { return a.x * a.x + a.y * a.y; }

// _support_rsqrt_rn, CPU branch (support_function.py:44-53)
inline float support_rsqrt_rn(float value) { return 1.0f / std::sqrt(value); }

// _support_map_box (support_function.py:121-129)
inline vec3 support_map_box(const GenericShapeData& geom, vec3 direction) {
    const float BOX_SUPPORT_DEADBAND = 1.0e-10f;
    float direction_scale = maxf(std::fabs(direction.x), maxf(std::fabs(direction.y), std::fabs(direction.z)));
    float threshold = BOX_SUPPORT_DEADBAND * direction_scale;
    float sx = direction.x >= -threshold ? 1.0f : -1.0f;
    float sy = direction.y >= -threshold ? 1.0f : -1.0f;
    float sz = direction.z >= -threshold ? 1.0f : -1.0f;
    return vec3(sx * geom.scale.x, sy * geom.scale.y, sz * geom.scale.z);
}

// support_map (support_function.py:133-352), primitive branches
inline vec3 support_map(const GenericShapeData& geom, vec3 direction) {
    const float eps = 1.0e-12f;
    vec3 result(0.f, 0.f, 0.f);
    const int t = geom.shape_type;
    if (t == T_CONVEX_MESH) {  // support_function.py:153-172: furthest hull vertex, strict '>' keeps the first maximum
        vec3 mesh_scale = geom.scale;
        int num_verts = geom.mesh_point_count;
        vec3 scaled_dir = cw_mul(direction, mesh_scale);
        float max_dot = -1.0e10f;
        int best_idx = 0;
        for (int i = 0; i < num_verts; ++i) {
            float dot_val = dot(load3(geom.mesh_points + 3 * i), scaled_dir);
            if (dot_val > max_dot) {
                max_dot = dot_val;
                best_idx = i;
            }
        }
        if (num_verts > 0) result = cw_mul(load3(geom.mesh_points + 3 * best_idx), mesh_scale);
    } else if (t == T_BOX) {
        result = support_map_box(geom, direction);
    } else if (t == T_SPHERE) {
        float radius = geom.scale.x;
        float dir_len_sq = length_sq(direction);
        vec3 n = dir_len_sq > eps ? direction * support_rsqrt_rn(dir_len_sq) : vec3(1.f, 0.f, 0.f);
        result = n * radius;
    } else if (t == T_CAPSULE) {
        float radius = geom.scale.x, half_height = geom.scale.y;
        float dir_len_sq = length_sq(direction);
        vec3 n = dir_len_sq > eps ? direction * support_rsqrt_rn(dir_len_sq) : vec3(1.f, 0.f, 0.f);
        result = n * radius;
        if (direction.z >= 0.0f) result = result + vec3(0.f, 0.f, half_height);
        else result = result + vec3(0.f, 0.f, -half_height);
    } else if (t == T_ELLIPSOID) {
        float a = geom.scale.x, b = geom.scale.y, c = geom.scale.z;
        float dir_len_sq = length_sq(direction);
        if (dir_len_sq > eps) {
            float adx = a * direction.x, bdy = b * direction.y, cdz = c * direction.z;
            float denom_sq = adx * adx + bdy * bdy + cdz * cdz;
            if (denom_sq > eps) {
                float inv_denom = support_rsqrt_rn(denom_sq);
                result = vec3((a * a) * direction.x * inv_denom, (b * b) * direction.y * inv_denom, (c * c) * direction.z * inv_denom);
            } else {
                result = vec3(a, 0.f, 0.f);
            }
        } else {
            result = vec3(a, 0.f, 0.f);
        }
    } else if (t == T_CYLINDER) {
        float radius = geom.scale.x, half_height = geom.scale.y, barrel_radius = geom.scale.z;
        vec3 dir_xy(direction.x, direction.y, 0.f);
        float dir_xy_len_sq = length_sq(dir_xy);
        if (barrel_radius == 0.0f) {
            vec3 lateral_point;
            if (dir_xy_len_sq > eps) {
                vec3 n_xy = dir_xy * support_rsqrt_rn(dir_xy_len_sq);
                lateral_point = vec3(n_xy.x * radius, n_xy.y * radius, 0.f);
            } else {
                lateral_point = vec3(radius, 0.f, 0.f);
            }
            if (direction.z > 0.0f) result = vec3(lateral_point.x, lateral_point.y, half_height);
            else if (direction.z < 0.0f) result = vec3(lateral_point.x, lateral_point.y, -half_height);
            else result = lateral_point;
        } else {
            vec3 n_xy;
            if (dir_xy_len_sq > eps) {
                float dir_xy_len = std::sqrt(dir_xy_len_sq);
                n_xy = dir_xy / dir_xy_len;
            } else {
                n_xy = vec3(1.f, 0.f, 0.f);
            }
            float direction_len = std::sqrt(dir_xy_len_sq + direction.z * direction.z);
            float support_z = 0.0f;
            if (direction_len > eps) support_z = clampf(barrel_radius * direction.z / direction_len, -half_height, half_height);
            float barrel_radius_sq = barrel_radius * barrel_radius;
            float half_height_sq = half_height * half_height;
            float support_z_sq = support_z * support_z;
            float end_offset = std::sqrt(barrel_radius_sq - half_height_sq);
            float support_offset = std::sqrt(maxf(barrel_radius_sq - support_z_sq, 0.0f));
            float offset_sum = support_offset + end_offset;
            float support_radius = radius;
            if (offset_sum > eps) support_radius += (half_height_sq - support_z_sq) / offset_sum;
            result = vec3(n_xy.x * support_radius, n_xy.y * support_radius, support_z);
        }
    } else if (t == T_CONE) {
        float radius = geom.scale.x, half_height = geom.scale.y;
        vec3 apex(0.f, 0.f, half_height);
        vec3 dir_xy(direction.x, direction.y, 0.f);
        float dir_xy_len = length(dir_xy);
        float k = half_height > eps ? radius / (2.0f * half_height) : 0.0f;
        if (dir_xy_len <= eps) {
            if (direction.z >= 0.0f) result = apex;
            else result = vec3(radius, 0.f, -half_height);
        } else {
            if (direction.z >= k * dir_xy_len) {
                result = apex;
            } else {
                vec3 n_xy = dir_xy / dir_xy_len;
                result = vec3(n_xy.x * radius, n_xy.y * radius, -half_height);
            }
        }
    } else if (t == T_PLANE) {
        float half_width = geom.scale.x, half_length = geom.scale.y;
        float sx = direction.x >= 0.0f ? 1.0f : -1.0f;
        float sy = direction.y >= 0.0f ? 1.0f : -1.0f;
        result = vec3(sx * half_width, sy * half_length, 0.f);
    }
    return result;
}

// create_shape_support_function (support_function.py:393-446)
inline vec3 shape_support(const GenericShapeData& geom, vec3 direction, bool center_ties) {
    if (!center_ties) return support_map(geom, direction);
    vec3 result(0.f, 0.f, 0.f);
    if (geom.shape_type == T_BOX) {
        const float TIE_EPSILON = 1.0e-6f;  // _CENTERED_BOX_SUPPORT_TIE_EPSILON
        vec3 abs_direction(std::fabs(direction.x), std::fabs(direction.y), std::fabs(direction.z));
        result = support_map_box(geom, direction);
        vec3 contribution = cw_mul(abs_direction, geom.scale);
        float threshold = TIE_EPSILON * (contribution.x + contribution.y + contribution.z);
        if (contribution.x <= threshold) result.x = 0.0f;
        if (contribution.y <= threshold) result.y = 0.0f;
        if (contribution.z <= threshold) result.z = 0.0f;
    } else {
        result = support_map(geom, direction);
    }
    return result;
}

struct Vert {  // mpr.py:41-55
    vec3 B, BtoA;
};
inline vec3 vert_a(const Vert& v) { return v.B + v.BtoA; }

// support_map_b + minkowski_support (mpr.py:76-150)
inline Vert minkowski_support(const GenericShapeData& geom_a, const GenericShapeData& geom_b, vec3 direction, quat orientation_b,
                              vec3 position_b, float extend, bool center_ties) {
    Vert v;
    vec3 point_a = shape_support(geom_a, direction, center_ties);
    vec3 tmp_direction = -direction;
    {
        vec3 tmp = quat_rotate_inv(orientation_b, tmp_direction);
        vec3 result = shape_support(geom_b, tmp, center_ties);
        result = quat_rotate(orientation_b, result);
        result = result + position_b;
        v.B = result;
    }
    if (extend != 0.0f) {
        vec3 d = normalize(direction) * extend * 0.5f;
        point_a = point_a + d;
        v.B = v.B - d;
    }
    v.BtoA = point_a - v.B;
    return v;
}

// shape_center (support_function.py:563-594) for primitives: no triangle adjustment
inline Vert geometric_center(const GenericShapeData& geom_a, const GenericShapeData& geom_b, quat orientation_b, vec3 position_b) {
    Vert c;
    c.B = position_b + quat_rotate(orientation_b, geom_b.center);
    c.BtoA = geom_a.center - c.B;
    return c;
}

struct MprResult {
    bool collision;
    vec3 point_a, point_b, normal;
    float penetration;
};

// solve_mpr_core (mpr.py:188-403)
inline MprResult solve_mpr_core(const GenericShapeData& geom_a, const GenericShapeData& geom_b, quat orientation_b, vec3 position_b,
                                float extend) {
    const int MAX_ITER = 30;
    const float COLLIDE_EPSILON = 1e-5f;
    const float NUMERIC_EPSILON = 1e-16f;
    MprResult r;
    r.collision = false;
    r.penetration = 0.0f;
    Vert v0 = geometric_center(geom_a, geom_b, orientation_b, position_b);
    vec3 normal = v0.BtoA;
    if (length_sq(normal) < NUMERIC_EPSILON) {
        v0.BtoA = vec3(0.f);  // _minkowski_center_fallback: zero for non-triangle shapes
        if (length_sq(v0.BtoA) < NUMERIC_EPSILON) {
            float best_dot = -1.0e30f;
            vec3 best_dir(1.f, 0.f, 0.f);
            for (int axis_idx = 0; axis_idx < 3; ++axis_idx) {
                vec3 probe(0.f, 0.f, 0.f);
                probe[axis_idx] = 1.0f;
                Vert sv = minkowski_support(geom_a, geom_b, probe, orientation_b, position_b, extend, true);
                float d = dot(sv.BtoA, probe);
                if (d > best_dot) {
                    best_dot = d;
                    best_dir = probe;
                }
            }
            v0.BtoA = best_dir * 1e-05f;
        }
    }
    normal = -v0.BtoA;
    Vert v1 = minkowski_support(geom_a, geom_b, normal, orientation_b, position_b, extend, true);
    r.point_a = vert_a(v1);
    r.point_b = v1.B;
    if (dot(v1.BtoA, normal) <= 0.0f) {
        r.normal = normal;
        return r;
    }
    normal = cross(v1.BtoA, v0.BtoA);
    if (length_sq(normal) < NUMERIC_EPSILON * NUMERIC_EPSILON) {
        normal = v1.BtoA - v0.BtoA;
        normal = normalize(normal);
        vec3 temp1 = v1.BtoA;
        r.penetration = dot(temp1, normal);
        r.collision = true;
        r.normal = normal;
        return r;
    }
    Vert v2 = minkowski_support(geom_a, geom_b, normal, orientation_b, position_b, extend, true);
    if (dot(v2.BtoA, normal) <= 0.0f) {
        r.normal = normal;
        return r;
    }
    vec3 temp1 = v1.BtoA - v0.BtoA;
    vec3 temp2 = v2.BtoA - v0.BtoA;
    normal = cross(temp1, temp2);
    float dist = dot(normal, v0.BtoA);
    if (dist > 0.0f) {
        Vert tmp = v1;
        v1 = v2;
        v2 = tmp;
        normal = -normal;
    }
    int phase1 = 0, phase2 = 0;
    bool hit = false;
    Vert v3;
    while (true) {
        if (phase1 > MAX_ITER) {
            r.normal = normal;
            return r;
        }
        phase1 += 1;
        v3 = minkowski_support(geom_a, geom_b, normal, orientation_b, position_b, extend, true);
        if (dot(v3.BtoA, normal) <= 0.0f) {
            r.normal = normal;
            return r;
        }
        temp1 = cross(v1.BtoA, v3.BtoA);
        if (dot(temp1, v0.BtoA) < 0.0f) {
            v2 = v3;
            temp1 = v1.BtoA - v0.BtoA;
            temp2 = v3.BtoA - v0.BtoA;
            normal = cross(temp1, temp2);
            continue;
        }
        temp1 = cross(v3.BtoA, v2.BtoA);
        if (dot(temp1, v0.BtoA) < 0.0f) {
            v1 = v3;
            temp1 = v3.BtoA - v0.BtoA;
            temp2 = v2.BtoA - v0.BtoA;
            normal = cross(temp1, temp2);
            continue;
        }
        break;
    }
    while (true) {
        phase2 += 1;
        temp1 = v2.BtoA - v1.BtoA;
        temp2 = v3.BtoA - v1.BtoA;
        normal = cross(temp1, temp2);
        float normal_sq = length_sq(normal);
        if (normal_sq < NUMERIC_EPSILON * NUMERIC_EPSILON) {
            r.normal = normal;
            return r;  // collision == False
        }
        if (!hit) {
            float d = dot(normal, v1.BtoA);
            hit = d >= 0.0f;
        }
        Vert v4 = minkowski_support(geom_a, geom_b, normal, orientation_b, position_b, extend, true);
        vec3 temp3 = v4.BtoA - v3.BtoA;
        float delta = dot(temp3, normal);
        r.penetration = dot(v4.BtoA, normal);
        if (delta * delta <= COLLIDE_EPSILON * COLLIDE_EPSILON * normal_sq || r.penetration <= 0.0f || phase2 > MAX_ITER) {
            if (hit) {
                float inv_normal = 1.0f / std::sqrt(normal_sq);
                r.penetration *= inv_normal;
                normal = normal * inv_normal;
                temp3 = cross(v1.BtoA, temp1);
                float gamma = dot(temp3, normal) * inv_normal;
                temp3 = cross(temp2, v1.BtoA);
                float beta = dot(temp3, normal) * inv_normal;
                float alpha = 1.0f - gamma - beta;
                r.point_a = alpha * vert_a(v1) + beta * vert_a(v2) + gamma * vert_a(v3);
                r.point_b = alpha * v1.B + beta * v2.B + gamma * v3.B;
            }
            r.collision = hit;
            r.normal = normal;
            return r;
        }
        temp1 = cross(v4.BtoA, v0.BtoA);
        float dt = dot(temp1, v1.BtoA);
        if (dt >= 0.0f) {
            dt = dot(temp1, v2.BtoA);
            if (dt >= 0.0f) v1 = v4;
            else v3 = v4;
        } else {
            dt = dot(temp1, v3.BtoA);
            if (dt >= 0.0f) v2 = v4;
            else v1 = v4;
        }
    }
}

// ---- simplex_solver.py ---------------------------------------------------------------------------------------------
struct Simplex {  // Mat83f: v[2*i] = B, v[2*i+1] = BtoA
    vec3 v[8];
};
struct Closest {
    vec3 point;
    float bc[4];
    uint32_t mask;
};
inline Closest closest_zero() {
    Closest c;
    c.bc[0] = c.bc[1] = c.bc[2] = c.bc[3] = 0.0f;
    c.mask = 0u;
    return c;
}
const float GJK_EPSILON = 1e-8f;

inline Closest closest_segment(const Simplex& s, int i0, int i1) {  // :104-150
    vec3 a = s.v[2 * i0 + 1], b = s.v[2 * i1 + 1];
    vec3 edge = b - a;
    float vsq = length_sq(edge);
    bool degenerate = vsq < GJK_EPSILON;
    float denom = vsq;
    if (degenerate) denom = GJK_EPSILON;
    float t = -dot(a, edge) / denom;
    float lambda0 = 1.0f - t, lambda1 = t;
    Closest c = closest_zero();
    c.mask = (1u << i0) | (1u << i1);
    if (lambda0 < 0.0f || degenerate) {
        c.mask = 1u << i1;
        lambda0 = 0.0f;
        lambda1 = 1.0f;
    } else if (lambda1 < 0.0f) {
        c.mask = 1u << i0;
        lambda0 = 1.0f;
        lambda1 = 0.0f;
    }
    c.bc[i0] = lambda0;
    c.bc[i1] = lambda1;
    c.point = lambda0 * a + lambda1 * b;
    return c;
}

inline Closest closest_triangle(const Simplex& s, int i0, int i1, int i2) {  // :152-230
    vec3 a = s.v[2 * i0 + 1], b = s.v[2 * i1 + 1], c = s.v[2 * i2 + 1];
    vec3 u = a - b, w = a - c;
    vec3 normal = cross(u, w);
    float t = length_sq(normal);
    bool degenerate = t < GJK_EPSILON;
    float denom = t;
    if (degenerate) denom = GJK_EPSILON;
    float it = 1.0f / denom;
    vec3 c1 = cross(u, a), c2 = cross(a, w);
    float lambda2 = dot(c1, normal) * it;
    float lambda1 = dot(c2, normal) * it;
    float lambda0 = 1.0f - lambda2 - lambda1;
    float best_distance = 1e30f;
    Closest best = closest_zero();
    best.point = vec3(0.f);
    if (lambda0 < 0.0f || degenerate) {
        Closest k = closest_segment(s, i1, i2);
        float dist = length_sq(k.point);
        if (dist < best_distance) {
            best = k;
            best_distance = dist;
        }
    }
    if (lambda1 < 0.0f || degenerate) {
        Closest k = closest_segment(s, i0, i2);
        float dist = length_sq(k.point);
        if (dist < best_distance) {
            best = k;
            best_distance = dist;
        }
    }
    if (lambda2 < 0.0f || degenerate) {
        Closest k = closest_segment(s, i0, i1);
        float dist = length_sq(k.point);
        if (dist < best_distance) best = k;
    }
    if (best.mask != 0u) return best;
    Closest r = closest_zero();
    r.bc[i0] = lambda0;
    r.bc[i1] = lambda1;
    r.bc[i2] = lambda2;
    r.mask = (1u << i0) | (1u << i1) | (1u << i2);
    r.point = lambda0 * a + lambda1 * b + lambda2 * c;
    return r;
}

inline float determinant(vec3 a, vec3 b, vec3 c, vec3 d) { return dot(b - a, cross(c - a, d - a)); }

inline Closest closest_tetrahedron(const Simplex& s) {  // :236-320
    vec3 v0 = s.v[1], v1 = s.v[3], v2 = s.v[5], v3 = s.v[7];
    float det_t = determinant(v0, v1, v2, v3);
    bool degenerate = std::fabs(det_t) < GJK_EPSILON;
    float denom = det_t;
    if (degenerate) denom = GJK_EPSILON;
    float inverse_det_t = 1.0f / denom;
    vec3 zero(0.f, 0.f, 0.f);
    float lambda0 = determinant(zero, v1, v2, v3) * inverse_det_t;
    float lambda1 = determinant(v0, zero, v2, v3) * inverse_det_t;
    float lambda2 = determinant(v0, v1, zero, v3) * inverse_det_t;
    float lambda3 = 1.0f - lambda0 - lambda1 - lambda2;
    float best_distance = 1e30f;
    Closest best = closest_zero();
    best.point = vec3(0.f);
    const int faces[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
    const float lambdas[4] = {lambda0, lambda1, lambda2, lambda3};
    for (int f = 0; f < 4; ++f) {
        if (lambdas[f] < 0.0f || degenerate) {
            Closest k = closest_triangle(s, faces[f][0], faces[f][1], faces[f][2]);
            float dist = length_sq(k.point);
            if (dist < best_distance) {
                best = k;
                if (f < 3) best_distance = dist;  // the fourth test of the reference does not update best_distance
            }
        }
    }
    if (best.mask != 0u) return best;
    Closest r = closest_zero();
    r.bc[0] = lambda0;
    r.bc[1] = lambda1;
    r.bc[2] = lambda2;
    r.bc[3] = lambda3;
    r.mask = 15u;
    r.point = zero;
    return r;
}

inline void simplex_get_closest(const Simplex& s, const float* barycentric, uint32_t usage_mask, vec3& point_a, vec3& point_b) {
    point_a = vec3(0.f);
    point_b = vec3(0.f);
    for (int i = 0; i < 4; ++i) {
        if ((usage_mask & (1u << i)) == 0u) continue;
        vec3 B = s.v[2 * i], BtoA = s.v[2 * i + 1];
        float bc_val = barycentric[i];
        point_a = point_a + bc_val * (B + BtoA);
        point_b = point_b + bc_val * B;
    }
}

struct GjkResult {
    bool separated;
    vec3 point_a, point_b, normal;
    float distance;
};

// solve_closest_distance_core (simplex_solver.py:322-495)
inline GjkResult solve_closest_distance_core(const GenericShapeData& geom_a, const GenericShapeData& geom_b, quat orientation_b,
                                             vec3 position_b, float extend, int MAX_ITER = 30, float COLLIDE_EPSILON = 1e-4f) {
    GjkResult r;
    r.distance = 0.0f;
    r.normal = vec3(0.f);
    Simplex simplex_v;
    float simplex_barycentric[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t simplex_usage_mask = 0u;
    int iter_count = MAX_ITER;
    Vert center = geometric_center(geom_a, geom_b, orientation_b, position_b);
    vec3 v = center.BtoA;
    float dist_sq = length_sq(v);
    vec3 last_search_dir(1.f, 0.f, 0.f);
    while (iter_count > 0) {
        iter_count -= 1;
        if (dist_sq < COLLIDE_EPSILON * COLLIDE_EPSILON) {
            r.distance = 0.0f;
            r.normal = vec3(0.f);
            simplex_get_closest(simplex_v, simplex_barycentric, simplex_usage_mask, r.point_a, r.point_b);
            r.separated = false;
            return r;
        }
        vec3 search_dir = -v;
        last_search_dir = search_dir;
        Vert w = minkowski_support(geom_a, geom_b, search_dir, orientation_b, position_b, extend, false);
        vec3 w_v = w.BtoA;
        float delta_dist = dot(v, v - w_v);
        if (delta_dist <= 0.0f || delta_dist * delta_dist < (COLLIDE_EPSILON * COLLIDE_EPSILON * dist_sq)) break;
        bool is_duplicate = false;
        for (int i = 0; i < 4; ++i)
            if ((simplex_usage_mask & (1u << i)) != 0u)
                if (length_sq(simplex_v.v[2 * i + 1] - w_v) < COLLIDE_EPSILON * COLLIDE_EPSILON) {
                    is_duplicate = true;
                    break;
                }
        if (is_duplicate) break;
        int use_count = 0, free_slot = 0;
        int indices[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            if ((simplex_usage_mask & (1u << i)) != 0u) {
                indices[use_count] = i;
                use_count += 1;
            } else {
                free_slot = i;
            }
        }
        indices[use_count] = free_slot;
        use_count += 1;
        simplex_v.v[2 * free_slot] = w.B;
        simplex_v.v[2 * free_slot + 1] = w.BtoA;
        vec3 closest(0.f, 0.f, 0.f);
        bool success = true;
        if (use_count == 1) {
            int i0 = indices[0];
            closest = simplex_v.v[2 * i0 + 1];
            simplex_usage_mask = 1u << i0;
            simplex_barycentric[i0] = 1.0f;
        } else if (use_count == 2) {
            Closest k = closest_segment(simplex_v, indices[0], indices[1]);
            closest = k.point;
            for (int i = 0; i < 4; ++i) simplex_barycentric[i] = k.bc[i];
            simplex_usage_mask = k.mask;
        } else if (use_count == 3) {
            Closest k = closest_triangle(simplex_v, indices[0], indices[1], indices[2]);
            closest = k.point;
            for (int i = 0; i < 4; ++i) simplex_barycentric[i] = k.bc[i];
            simplex_usage_mask = k.mask;
        } else if (use_count == 4) {
            Closest k = closest_tetrahedron(simplex_v);
            closest = k.point;
            for (int i = 0; i < 4; ++i) simplex_barycentric[i] = k.bc[i];
            simplex_usage_mask = k.mask;
            bool inside_tetrahedron = k.mask == 15u;
            success = !inside_tetrahedron;
        } else {
            success = false;
        }
        vec3 new_v = closest;
        if (!success) {
            r.distance = 0.0f;
            r.normal = vec3(0.f);
            simplex_get_closest(simplex_v, simplex_barycentric, simplex_usage_mask, r.point_a, r.point_b);
            r.separated = false;
            return r;
        }
        v = new_v;
        dist_sq = length_sq(v);
    }
    simplex_get_closest(simplex_v, simplex_barycentric, simplex_usage_mask, r.point_a, r.point_b);
    vec3 delta = r.point_b - r.point_a;
    float delta_len_sq = length_sq(delta);
    if (delta_len_sq > GJK_EPSILON * GJK_EPSILON) {
        r.distance = std::sqrt(delta_len_sq);
        r.normal = delta * (1.0f / r.distance);
    } else {
        r.distance = std::sqrt(dist_sq);
        if (r.distance > COLLIDE_EPSILON) {
            r.normal = v * (-1.0f / r.distance);
        } else {
            float nsq = length_sq(last_search_dir);
            if (nsq > 0.0f) r.normal = last_search_dir * (1.0f / std::sqrt(nsq));
            else r.normal = vec3(1.f, 0.f, 0.f);
        }
    }
    r.separated = true;
    return r;
}

// ---- multicontact.py -----------------------------------------------------------------------------------------------
const float MC_EPS = 0.00001f;
const float SIN_TILT_ANGLE = 0.03489949670250097f;           // sin(2 deg)
const float COS_TILT_ANGLE = 0.9993908270190958f;            // cos(2 deg)
const float COS_DEEPEST_CONTACT_THRESHOLD_ANGLE = 0.9999984769132877f;  // cos(0.1 deg)

inline float signed_area(vec2 a, vec2 b, vec2 q) { return (b.x - a.x) * (q.y - a.y) - (b.y - a.y) * (q.x - a.x); }

inline vec3 ray_plane_intersection(vec3 ray_origin, vec3 ray_direction, float plane_d, vec3 plane_normal) {  // :83-114
    float denom = dot(ray_direction, plane_normal);
    if (std::fabs(denom) < 1.0e-12f) return ray_origin;
    float t = -(dot(ray_origin, plane_normal) + plane_d) / denom;
    return ray_origin + ray_direction * t;
}

struct BodyProjector {
    float plane_d = 0.f;
    vec3 normal;
};
struct IncrementalPlaneTracker {
    vec3 reference_point, previous_point, normal;
    float largest_area_sq = 0.f;
};

inline void update_incremental_plane_tracker(IncrementalPlaneTracker& tracker, vec3 current_point, int current_point_id) {  // :140-162
    if (current_point_id == 0) {
        tracker.reference_point = current_point;
        tracker.largest_area_sq = 0.0f;
    } else if (current_point_id == 1) {
        tracker.previous_point = current_point;
    } else {
        vec3 edge1 = tracker.previous_point - tracker.reference_point;
        vec3 edge2 = current_point - tracker.reference_point;
        vec3 cr = cross(edge1, edge2);
        float area_sq = dot(cr, cr);
        if (area_sq > tracker.largest_area_sq) {
            tracker.largest_area_sq = area_sq;
            tracker.normal = cr;
        }
        tracker.previous_point = current_point;
    }
}

inline vec3 compute_line_segment_projector_normal(vec3 segment_dir, vec3 reference_normal) {  // :165-184
    vec3 right = cross(segment_dir, reference_normal);
    vec3 normal = cross(right, segment_dir);
    float len = length(normal);
    return len > 1.0e-12f ? normal * (1.0f / len) : reference_normal;
}

inline void create_body_projectors(const IncrementalPlaneTracker& ta, vec3 anchor_point_a, const IncrementalPlaneTracker& tb,
                                   vec3 anchor_point_b, vec3 contact_normal, BodyProjector& pa, BodyProjector& pb) {  // :187-238
    pa = BodyProjector();
    pb = BodyProjector();
    if (ta.largest_area_sq == 0.0f && tb.largest_area_sq == 0.0f) {
        vec3 dir_a = ta.previous_point - ta.reference_point;
        vec3 dir_b = tb.previous_point - tb.reference_point;
        vec3 point_on_plane_a = 0.5f * (ta.reference_point + ta.previous_point);
        pa.normal = compute_line_segment_projector_normal(dir_a, contact_normal);
        pa.plane_d = -dot(point_on_plane_a, pa.normal);
        vec3 point_on_plane_b = 0.5f * (tb.reference_point + tb.previous_point);
        pb.normal = compute_line_segment_projector_normal(dir_b, contact_normal);
        pb.plane_d = -dot(point_on_plane_b, pb.normal);
        return;
    }
    if (ta.largest_area_sq > 0.0f) {
        float inv_len_n = 1.0f / std::sqrt(maxf(1.0e-12f, ta.largest_area_sq));
        pa.normal = ta.normal * inv_len_n;
        pa.plane_d = -dot(anchor_point_a, pa.normal);
    }
    if (tb.largest_area_sq > 0.0f) {
        float inv_len_n = 1.0f / std::sqrt(maxf(1.0e-12f, tb.largest_area_sq));
        pb.normal = tb.normal * inv_len_n;
        pb.plane_d = -dot(anchor_point_b, pb.normal);
    }
    if (ta.largest_area_sq == 0.0f) {
        vec3 dir = ta.previous_point - ta.reference_point;
        vec3 point_on_plane_a = 0.5f * (ta.reference_point + ta.previous_point);
        pa.normal = compute_line_segment_projector_normal(dir, pb.normal);
        pa.plane_d = -dot(point_on_plane_a, pa.normal);
    }
    if (tb.largest_area_sq == 0.0f) {
        vec3 dir = tb.previous_point - tb.reference_point;
        vec3 point_on_plane_b = 0.5f * (tb.reference_point + tb.previous_point);
        pb.normal = compute_line_segment_projector_normal(dir, pa.normal);
        pb.plane_d = -dot(point_on_plane_b, pb.normal);
    }
}

inline vec2 intersection_point(vec2 trim_seg_start, vec2 trim_seg_end, vec2 a, vec2 b) {  // :263-287
    float signed_a = signed_area(trim_seg_start, trim_seg_end, a);
    float signed_b = signed_area(trim_seg_start, trim_seg_end, b);
    float interp_ab = std::fabs(signed_a) / std::fabs(signed_a - signed_b);
    return (1.0f - interp_ab) * a + interp_ab * b;
}

inline void insert_vec2(vec2* arr, int arr_count, int index, vec2 element) {  // :290-305
    int i = arr_count;
    while (i > index) {
        arr[i] = arr[i - 1];
        i -= 1;
    }
    arr[index] = element;
}

inline int trim_in_place(vec2 trim_seg_start, vec2 trim_seg_end, vec2* loop, int loop_count) {  // :308-411
    if (loop_count < 3) return loop_count;
    vec2 intersection_a(0.f, 0.f), intersection_b(0.f, 0.f);
    int change_a = -1, change_b = -1;
    bool keep = false;
    bool prev_outside = signed_area(trim_seg_start, trim_seg_end, loop[0]) <= 0.0f;
    for (int i = 0; i < loop_count; ++i) {
        int next_idx = (i + 1) % loop_count;
        bool outside = signed_area(trim_seg_start, trim_seg_end, loop[next_idx]) <= 0.0f;
        if (outside != prev_outside) {
            vec2 intersection = intersection_point(trim_seg_start, trim_seg_end, loop[i], loop[next_idx]);
            if (change_a < 0) {
                change_a = i;
                keep = !prev_outside;
                intersection_a = intersection;
            } else {
                change_b = i;
                intersection_b = intersection;
            }
        }
        prev_outside = outside;
    }
    int new_loop_count;
    if (change_a >= 0 && change_b >= 0) {
        int loop_indexer = -1;
        new_loop_count = loop_count;
        int i = 0;
        while (i < loop_count) {
            if (keep) {
                loop_indexer += 1;
                loop[loop_indexer] = loop[i];
            }
            if (i == change_a || i == change_b) {
                vec2 pt = i == change_a ? intersection_a : intersection_b;
                if (loop_indexer == i && !keep) {
                    loop_indexer += 1;
                    insert_vec2(loop, new_loop_count, loop_indexer, pt);
                    new_loop_count += 1;
                    i += 1;
                    change_b += 1;
                    loop_count += 1;
                } else {
                    loop_indexer += 1;
                    loop[loop_indexer] = pt;
                }
                keep = !keep;
            }
            i += 1;
        }
        new_loop_count = loop_indexer + 1;
    } else if (prev_outside) {
        new_loop_count = 0;
    } else {
        new_loop_count = loop_count;
    }
    return new_loop_count;
}

// segment -> thin rectangle (both occurrences inside trim_all_in_place, :433-470)
inline bool segment_to_rectangle(vec2* poly) {
    const float move_distance = 1e-5f;
    vec2 p0 = poly[0], p1 = poly[1];
    float dir_x = p1.x - p0.x, dir_y = p1.y - p0.y;
    float dir_len = std::sqrt(dir_x * dir_x + dir_y * dir_y);
    if (dir_len > 1e-10f) {
        float inv_dir_len = 1.0f / dir_len;
        float perp_x = -dir_y * inv_dir_len, perp_y = dir_x * inv_dir_len;
        float offset_x = perp_x * move_distance, offset_y = perp_y * move_distance;
        poly[0] = vec2(p0.x - offset_x, p0.y - offset_y);
        poly[1] = vec2(p1.x - offset_x, p1.y - offset_y);
        poly[2] = vec2(p1.x + offset_x, p1.y + offset_y);
        poly[3] = vec2(p0.x + offset_x, p0.y + offset_y);
        return true;
    }
    return false;
}

inline int trim_all_in_place(vec2* trim_poly, int trim_poly_count, vec2* loop, int loop_count) {  // :414-483
    if (trim_poly_count <= 1) return loop_count < 1 ? loop_count : 1;
    if (trim_poly_count == 2) {
        if (!segment_to_rectangle(trim_poly)) return loop_count < 1 ? loop_count : 1;
        trim_poly_count = 4;
    }
    if (loop_count == 2) {
        if (!segment_to_rectangle(loop)) return loop_count < 1 ? loop_count : 1;
        loop_count = 4;
    }
    int current_loop_count = loop_count;
    vec2 trim_poly_0 = trim_poly[0];
    for (int i = 0; i < trim_poly_count; ++i) {
        vec2 trim_seg_start = trim_poly[i];
        vec2 trim_seg_end = i == trim_poly_count - 1 ? trim_poly_0 : trim_poly[i + 1];
        current_loop_count = trim_in_place(trim_seg_start, trim_seg_end, loop, current_loop_count);
    }
    return current_loop_count;
}

inline void approx_max_quadrilateral_area_with_calipers(const vec2* hull, int hull_count, int out[4]) {  // :486-560
    int n = hull_count;
    int p1 = 0, p3 = 1;
    vec2 diff(hull[p1].x - hull[p3].x, hull[p1].y - hull[p3].y);
    float max_dist_sq = diff.x * diff.x + diff.y * diff.y;
    const float tie_epsilon_rel = 1.0e-3f;
    int j = 1;
    for (int i = 0; i < n; ++i) {
        vec2 hull_i = hull[i], hull_i_plus_1 = hull[(i + 1) % n];
        while (true) {
            vec2 hull_j = hull[j], hull_j_plus_1 = hull[(j + 1) % n];
            float area_j_plus_1 = signed_area(hull_i, hull_i_plus_1, hull_j_plus_1);
            float area_j = signed_area(hull_i, hull_i_plus_1, hull_j);
            if (area_j_plus_1 > area_j) j = (j + 1) % n;
            else break;
        }
        vec2 hi = hull[i], hj = hull[j];
        vec2 d1(hi.x - hj.x, hi.y - hj.y);
        float dist_sq_1 = d1.x * d1.x + d1.y * d1.y;
        if (dist_sq_1 > max_dist_sq * (1.0f + tie_epsilon_rel)) {
            max_dist_sq = dist_sq_1;
            p1 = i;
            p3 = j;
        }
        vec2 hip1 = hull[(i + 1) % n];
        vec2 d2(hip1.x - hj.x, hip1.y - hj.y);
        float dist_sq_2 = d2.x * d2.x + d2.y * d2.y;
        if (dist_sq_2 > max_dist_sq * (1.0f + tie_epsilon_rel)) {
            max_dist_sq = dist_sq_2;
            p1 = (i + 1) % n;
            p3 = j;
        }
    }
    int p2 = 0, p4 = 0;
    float max_area_1 = 0.0f, max_area_2 = 0.0f;
    vec2 hull_p1 = hull[p1], hull_p3 = hull[p3];
    for (int i = 0; i < n; ++i) {
        float area = signed_area(hull_p1, hull_p3, hull[i]);
        if (area > max_area_1 * (1.0f + tie_epsilon_rel)) {
            max_area_1 = area;
            p2 = i;
        } else if (-area > max_area_2 * (1.0f + tie_epsilon_rel)) {
            max_area_2 = -area;
            p4 = i;
        }
    }
    out[0] = p1;
    out[1] = p2;
    out[2] = p3;
    out[3] = p4;
}

inline int remove_zero_length_edges(vec2* loop, int loop_count, float eps) {  // :563-596
    if (loop_count < 2) return 0;
    int write_idx = 0;
    for (int read_idx = 1; read_idx < loop_count; ++read_idx) {
        vec2 diff = loop[read_idx] - loop[write_idx];
        if (length_sq(diff) > eps) {
            write_idx += 1;
            loop[write_idx] = loop[read_idx];
        }
    }
    int new_loop_count;
    if (write_idx > 0) {
        vec2 diff = loop[write_idx] - loop[0];
        if (length_sq(diff) < eps) new_loop_count = write_idx;
        else new_loop_count = write_idx + 1;
    } else {
        new_loop_count = write_idx + 1;
    }
    if (new_loop_count < 2) new_loop_count = 0;
    return new_loop_count;
}

inline bool add_avoid_duplicates_vec2(vec2* arr, int& arr_count, vec2 vec, float eps) {  // :599-621
    if (arr_count > 0)
        if (length_sq(arr[0] - vec) < eps) return false;
    if (arr_count > 1)
        if (length_sq(arr[arr_count - 1] - vec) < eps) return false;
    arr[arr_count] = vec;
    arr_count += 1;
    return true;
}

inline void orthonormal_basis(vec3 n, vec3& b1, vec3& b2) {  // math/__init__.py:236-275
    if (n.z < 0.0f) {
        float a = 1.0f / (1.0f - n.z);
        float b = n.x * n.y * a;
        b1 = vec3(1.0f - n.x * n.x * a, -b, n.x);
        b2 = vec3(b, n.y * n.y * a - 1.0f, -n.y);
    } else {
        float a = 1.0f / (1.0f + n.z);
        float b = -n.x * n.y * a;
        b1 = vec3(1.0f - n.x * n.x * a, b, -n.x);
        b2 = vec3(b, 1.0f - n.y * n.y * a, -n.y);
    }
}

// ---- contact data + post-processing + writer gate ------------------------------------------------------------------------
struct ContactOut {  // what write_contact receives: centre, normal A->B, signed distance (already gap-tested)
    vec3 center, normal;
    float distance;
};
// write_contact_speculative (sim/collide.py:257-280): the writer admits a contact that is present now (separation <= the AUTHORED
// pair gap) or predicted within the collision-update horizon (contact_data.py:187-233)
struct SpeculativeWriter {
    bool enabled = false;
    float base_gap_sum = 0.f, dt = 0.f, max_extension = 0.f;
    vec3 origin_a, origin_b, linear_velocity_a, linear_velocity_b, angular_velocity_a, angular_velocity_b;
};
struct PairCtx {
    float radius_eff_a = 0.f, radius_eff_b = 0.f, margin_a = 0.f, margin_b = 0.f, gap_sum = 0.f;
    ContactOut out[5];
    int count = 0;
    SpeculativeWriter spec;
};

inline bool is_discrete_shape(int t) { return t == T_BOX || t == T_CONVEX_MESH || t == T_PLANE; }  // collision_core.py:40-48 (triangles n/a)

// post_process_axial_on_discrete_contact (collision_core.py:174-277) followed by the writer's gap test
// (sim/collide.py:210-254 with contact_passes_gap_check): emits into ctx
inline void post_process_and_write(PairCtx& ctx, vec3 contact_point_center, vec3 contact_normal_a_to_b, float contact_distance,
                                   const GenericShapeData& shape_a, vec3 pos_a_adjusted, quat rot_a, const GenericShapeData& shape_b,
                                   vec3 pos_b_adjusted, quat rot_b) {
    const int type_a = shape_a.shape_type, type_b = shape_b.shape_type;
    const vec3 normal = contact_normal_a_to_b;
    if (type_a == T_SPHERE || type_a == T_CAPSULE) {
        contact_point_center = contact_point_center + normal * (ctx.radius_eff_a * 0.5f);
        contact_distance = contact_distance - ctx.radius_eff_a;
    }
    if (type_b == T_SPHERE || type_b == T_CAPSULE) {
        contact_point_center = contact_point_center - normal * (ctx.radius_eff_b * 0.5f);
        contact_distance = contact_distance - ctx.radius_eff_b;
    }
    bool is_discrete_a = is_discrete_shape(type_a), is_discrete_b = is_discrete_shape(type_b);
    bool is_axial_a = type_a == T_CYLINDER || type_a == T_CONE;
    bool is_axial_b = type_b == T_CYLINDER || type_b == T_CONE;
    if ((is_discrete_a && is_axial_b) || (is_discrete_b && is_axial_a)) {
        vec3 shape_axis, shape_pos, axial_normal;
        float shape_radius, shape_half_height;
        bool is_cone;
        if (is_discrete_a && is_axial_b) {
            shape_axis = quat_rotate(rot_b, vec3(0.f, 0.f, 1.f));
            shape_radius = shape_b.scale.x;
            shape_half_height = shape_b.scale.y;
            is_cone = type_b == T_CONE;
            shape_pos = pos_b_adjusted;
            axial_normal = normal;
        } else {
            shape_axis = quat_rotate(rot_a, vec3(0.f, 0.f, 1.f));
            shape_radius = shape_a.scale.x;
            shape_half_height = shape_a.scale.y;
            is_cone = type_a == T_CONE;
            shape_pos = pos_a_adjusted;
            axial_normal = -normal;
        }
        float axis_normal_dot = std::fabs(dot(shape_axis, axial_normal));
        bool is_rolling = false;
        if (is_cone) {
            float cone_half_angle = (float)std::atan2((double)shape_radius, (double)(2.0f * shape_half_height));
            const float tolerance_angle = 0.03490658503988659f;  // 2 deg
            float lower_threshold = sin_w(cone_half_angle - tolerance_angle);
            float upper_threshold = sin_w(cone_half_angle + tolerance_angle);
            if (axis_normal_dot >= lower_threshold && axis_normal_dot <= upper_threshold) is_rolling = true;
        } else {
            const float perpendicular_threshold = 0.03489949670250097f;  // sin(2 deg)
            if (axis_normal_dot <= perpendicular_threshold) is_rolling = true;
        }
        if (is_rolling) {
            vec3 projection_plane_normal = normalize(cross(shape_axis, axial_normal));
            vec3 to_point = contact_point_center - shape_pos;  // project_point_onto_plane (collision_core.py:51-67)
            float distance_to_plane = dot(to_point, projection_plane_normal);
            contact_point_center = contact_point_center - projection_plane_normal * distance_to_plane;
        }
    }
    // writer gate: _contact_passes_gap_check (contact_data.py:118-156)
    float total_separation_needed = ctx.radius_eff_a + ctx.radius_eff_b + ctx.margin_a + ctx.margin_b;
    vec3 n = normalize(contact_normal_a_to_b);
    vec3 a_w = contact_point_center - n * (0.5f * contact_distance + ctx.radius_eff_a);
    vec3 b_w = contact_point_center + n * (0.5f * contact_distance + ctx.radius_eff_b);
    float d = dot(b_w - a_w, n) - total_separation_needed;
    if (ctx.spec.enabled) {
        const SpeculativeWriter& w = ctx.spec;
        if (!(d <= w.base_gap_sum)) {
            vec3 velocity_a = w.linear_velocity_a + cross(w.angular_velocity_a, a_w - w.origin_a);
            vec3 velocity_b = w.linear_velocity_b + cross(w.angular_velocity_b, b_w - w.origin_b);
            float approach_speed = maxf(-dot(velocity_b - velocity_a, n), 0.0f);
            float extension = minf(approach_speed * w.dt, w.max_extension);
            if (!(extension - d >= 0.0f)) return;
        }
    } else if (d > ctx.gap_sum) return;
    ContactOut& o = ctx.out[ctx.count++];
    o.center = contact_point_center;
    o.normal = contact_normal_a_to_b;
    o.distance = contact_distance;
}

// build_manifold + extract_4_point_contact_manifolds (multicontact.py:655-956)
inline void build_manifold(PairCtx& ctx, const GenericShapeData& geom_a, const GenericShapeData& geom_b, quat orientation_a,
                           vec3 position_a_world, quat relative_orientation_b, vec3 relative_position_b, vec3 p_a, vec3 p_b, vec3 normal) {
    const float PENT_COS[5] = {1.0f, 0.30901699437494745f, -0.8090169943749473f, -0.8090169943749476f, 0.30901699437494723f};
    const float PENT_SIN[5] = {0.0f, 0.9510565162951535f, 0.5877852522924732f, -0.587785252292473f, -0.9510565162951536f};
    int a_count = 0, b_count = 0;
    vec3 tangent_a, tangent_b;
    orthonormal_basis(normal, tangent_a, tangent_b);
    IncrementalPlaneTracker plane_tracker_a, plane_tracker_b;
    vec3 center = 0.5f * (p_a + p_b);
    vec2 b_buffer[10];  // a_buffer aliases b_buffer + 5, exactly like the reference's pointer arithmetic
    vec2* a_buffer = b_buffer + 5;
    vec3 local_normal_b = quat_rotate_inv(relative_orientation_b, -normal);
    vec3 local_ta_b = quat_rotate_inv(relative_orientation_b, -tangent_a);
    vec3 local_tb_b = quat_rotate_inv(relative_orientation_b, -tangent_b);
    for (int e = 0; e < 5; ++e) {
        float c = PENT_COS[e], s = PENT_SIN[e];
        float cos_tilt = COS_TILT_ANGLE;
        float c_sin = c * SIN_TILT_ANGLE, s_sin = s * SIN_TILT_ANGLE;
        vec3 dir_a = normal * cos_tilt + c_sin * tangent_a + s_sin * tangent_b;
        vec3 pt_a_3d = support_map(geom_a, dir_a);
        vec3 projected_a = pt_a_3d - center;
        vec2 pt_a_2d(dot(tangent_a, projected_a), dot(tangent_b, projected_a));
        if (add_avoid_duplicates_vec2(a_buffer, a_count, pt_a_2d, MC_EPS)) update_incremental_plane_tracker(plane_tracker_a, pt_a_3d, a_count - 1);
        vec3 local_dir_b = local_normal_b * cos_tilt + c_sin * local_ta_b + s_sin * local_tb_b;
        vec3 pt_b_local = support_map(geom_b, local_dir_b);
        vec3 pt_b_3d = quat_rotate(relative_orientation_b, pt_b_local) + relative_position_b;
        vec3 projected_b = pt_b_3d - center;
        vec2 pt_b_2d(dot(tangent_a, projected_b), dot(tangent_b, projected_b));
        if (add_avoid_duplicates_vec2(b_buffer, b_count, pt_b_2d, MC_EPS)) update_incremental_plane_tracker(plane_tracker_b, pt_b_3d, b_count - 1);
    }
    vec3 normal_world = quat_rotate(orientation_a, normal);
    vec3 position_a_ws = position_a_world;
    vec3 position_b_ws = quat_rotate(orientation_a, relative_position_b) + position_a_world;
    quat quaternion_a_ws = orientation_a;
    quat quaternion_b_ws = orientation_a * relative_orientation_b;
    int count_out = 0;
    float normal_dot = 0.0f;
    if (!(a_count < 2 || b_count < 2)) {
        BodyProjector projector_a, projector_b;
        create_body_projectors(plane_tracker_a, p_a, plane_tracker_b, p_b, normal, projector_a, projector_b);
        bool excess_a = std::fabs(dot(normal, projector_a.normal)) < COS_TILT_ANGLE;
        bool excess_b = std::fabs(dot(normal, projector_b.normal)) < COS_TILT_ANGLE;
        if (!(excess_a || excess_b)) {
            // extract_4_point_contact_manifolds
            normal_dot = std::fabs(dot(projector_a.normal, projector_b.normal));
            int loop_count = trim_all_in_place(a_buffer, a_count, b_buffer, b_count);
            loop_count = remove_zero_length_edges(b_buffer, loop_count, MC_EPS);
            if (loop_count > 1) {
                int result[4] = {0, 1, 2, 3};
                if (loop_count > 4) {
                    approx_max_quadrilateral_area_with_calipers(b_buffer, loop_count, result);
                    loop_count = 4;
                }
                for (int i = 0; i < loop_count; ++i) {
                    int ia = result[i];
                    vec3 p_local = b_buffer[ia].x * tangent_a + b_buffer[ia].y * tangent_b + center;
                    vec3 a = ray_plane_intersection(p_local, normal, projector_a.plane_d, projector_a.normal);
                    vec3 b = ray_plane_intersection(p_local, normal, projector_b.plane_d, projector_b.normal);
                    vec3 contact_point_local = 0.5f * (a + b);
                    float signed_distance = dot(b - a, normal);
                    vec3 contact_point_world = quat_rotate(orientation_a, contact_point_local) + position_a_world;
                    post_process_and_write(ctx, contact_point_world, normal_world, signed_distance, geom_a, position_a_ws, quaternion_a_ws, geom_b,
                                           position_b_ws, quaternion_b_ws);
                }
                count_out = loop_count < 4 ? loop_count : 4;
            } else {
                normal_dot = 0.0f;
                count_out = 0;
            }
        }
    }
    if (normal_dot < COS_DEEPEST_CONTACT_THRESHOLD_ANGLE || count_out == 0) {
        vec3 deepest_center_local = 0.5f * (p_a + p_b);
        float deepest_signed_distance = dot(p_b - p_a, normal);
        vec3 deepest_center_world = quat_rotate(orientation_a, deepest_center_local) + position_a_world;
        post_process_and_write(ctx, deepest_center_world, normal_world, deepest_signed_distance, geom_a, position_a_ws, quaternion_a_ws, geom_b,
                               position_b_ws, quaternion_b_ws);
    }
}

// solve_convex_multi_contact (collision_convex.py:107-231)
inline void solve_convex_multi_contact(PairCtx& ctx, const GenericShapeData& geom_a, const GenericShapeData& geom_b, quat orientation_a,
                                       quat orientation_b, vec3 position_a, vec3 position_b, float contact_threshold, bool skip_multi_contact) {
    quat relative_orientation_b = quat_inverse(orientation_a) * orientation_b;
    vec3 relative_position_b = quat_rotate_inv(orientation_a, position_b - position_a);
    float margin_sum = ctx.margin_a + ctx.margin_b;
    const float eps = 1.0e-4f;
    float enlarge;
    if (margin_sum <= 0.0f) enlarge = eps;
    else if (margin_sum < eps) enlarge = 2.0f * eps;
    else enlarge = 0.0f;
    MprResult m = solve_mpr_core(geom_a, geom_b, relative_orientation_b, relative_position_b, enlarge);
    vec3 point_a = m.point_a, point_b = m.point_b, normal = m.normal;
    float signed_distance;
    if (m.collision) {
        signed_distance = -m.penetration + enlarge;
        float half_enlarge = enlarge * 0.5f;
        point_a = point_a - normal * half_enlarge;
        point_b = point_b + normal * half_enlarge;
    } else {
        GjkResult g = solve_closest_distance_core(geom_a, geom_b, relative_orientation_b, relative_position_b, 0.0f);
        point_a = g.point_a;
        point_b = g.point_b;
        normal = g.normal;
        signed_distance = g.distance;
    }
    if (skip_multi_contact || signed_distance > contact_threshold) {
        vec3 point = 0.5f * (point_a + point_b);
        point = quat_rotate(orientation_a, point) + position_a;
        vec3 normal_ws = quat_rotate(orientation_a, normal);
        post_process_and_write(ctx, point, normal_ws, signed_distance, geom_a, position_a, orientation_a, geom_b, position_b, orientation_b);
        return;
    }
    build_manifold(ctx, geom_a, geom_b, orientation_a, position_a, relative_orientation_b, relative_position_b, point_a, point_b, normal);
}

// compute_gjk_mpr_contacts (collision_core.py:337-450)
inline void compute_gjk_mpr_contacts(PairCtx& ctx, GenericShapeData shape_a_data, GenericShapeData shape_b_data, quat rot_a, quat rot_b,
                                     vec3 pos_a_adjusted, vec3 pos_b_adjusted, float rigid_gap, float margin_a, float margin_b) {
    float radius_eff_a = 0.0f, radius_eff_b = 0.0f;
    const float small_radius = 0.0001f;
    int type_a = shape_a_data.shape_type, type_b = shape_b_data.shape_type;
    if (type_a == T_SPHERE || type_a == T_CAPSULE) {
        radius_eff_a = shape_a_data.scale.x;
        shape_a_data.scale.x = small_radius;
    }
    if (type_b == T_SPHERE || type_b == T_CAPSULE) {
        radius_eff_b = shape_b_data.scale.x;
        shape_b_data.scale.x = small_radius;
    }
    ctx.radius_eff_a = radius_eff_a;
    ctx.radius_eff_b = radius_eff_b;
    ctx.margin_a = margin_a;
    ctx.margin_b = margin_b;
    ctx.gap_sum = rigid_gap;
    solve_convex_multi_contact(ctx, shape_a_data, shape_b_data, rot_a, rot_b, pos_a_adjusted, pos_b_adjusted,
                               rigid_gap + radius_eff_a + radius_eff_b + margin_a + margin_b,
                               type_a == T_SPHERE || type_b == T_SPHERE || type_a == T_ELLIPSOID || type_b == T_ELLIPSOID);
}

// compute_tight_aabb_from_support, generic branch (collision_core.py:454-548): six support evaluations along the world
// axes expressed in the shape's frame
inline void compute_tight_aabb_from_support(const GenericShapeData& shape_data, quat orientation, vec3 center_pos, vec3& aabb_min, vec3& aabb_max) {
    mat33 rot_mat = quat_to_matrix(orientation);
    mat33 rot_mat_t = transpose(rot_mat);
    vec3 local_x(rot_mat_t.m[0][0], rot_mat_t.m[1][0], rot_mat_t.m[2][0]);
    vec3 local_y(rot_mat_t.m[0][1], rot_mat_t.m[1][1], rot_mat_t.m[2][1]);
    vec3 local_z(rot_mat_t.m[0][2], rot_mat_t.m[1][2], rot_mat_t.m[2][2]);
    if (shape_data.shape_type == T_CONVEX_MESH) {  // collision_core.py:492-523: one pass over the vertices
        vec3 mesh_scale = shape_data.scale;
        vec3 scaled_x = cw_mul(local_x, mesh_scale), scaled_y = cw_mul(local_y, mesh_scale), scaled_z = cw_mul(local_z, mesh_scale);
        float mnx = 1.0e10f, mxx = -1.0e10f, mny = 1.0e10f, mxy = -1.0e10f, mnz = 1.0e10f, mxz = -1.0e10f;
        for (int i = 0; i < shape_data.mesh_point_count; ++i) {
            vec3 p = load3(shape_data.mesh_points + 3 * i);
            float vx = dot(p, scaled_x), vy = dot(p, scaled_y), vz = dot(p, scaled_z);
            mnx = minf(mnx, vx); mxx = maxf(mxx, vx);
            mny = minf(mny, vy); mxy = maxf(mxy, vy);
            mnz = minf(mnz, vz); mxz = maxf(mxz, vz);
        }
        aabb_min = vec3(mnx, mny, mnz) + center_pos;
        aabb_max = vec3(mxx, mxy, mxz) + center_pos;
        return;
    }
    float max_x = dot(local_x, support_map(shape_data, local_x));
    float max_y = dot(local_y, support_map(shape_data, local_y));
    float max_z = dot(local_z, support_map(shape_data, local_z));
    float min_x = dot(local_x, support_map(shape_data, -local_x));
    float min_y = dot(local_y, support_map(shape_data, -local_y));
    float min_z = dot(local_z, support_map(shape_data, -local_z));
    aabb_min = vec3(min_x, min_y, min_z) + center_pos;
    aabb_max = vec3(max_x, max_y, max_z) + center_pos;
}

struct HullRef {  // model.shape_source (unscaled hull vertices) + model.shape_collision_aabb_lower / _upper of one CONVEX_MESH shape
    const float* points = nullptr;
    int count = 0;
    vec3 local_aabb_lower, local_aabb_upper;
};

// One pair of narrow_phase_kernel_gjk_mpr (narrow_phase.py:1066-1216) with external AABBs + find_contacts
// (collision_core.py:700-790).  `scale_*` are the MODEL's shape scales; finite planes are halved like geom_data.
inline int gjk_mpr_pair(int type_a, vec3 scale_a, const transform& X_a, float margin_a, vec3 aabb_lo_a, vec3 aabb_hi_a, int type_b, vec3 scale_b,
                        const transform& X_b, float margin_b, vec3 aabb_lo_b, vec3 aabb_hi_b, float gap_sum, ContactOut* out,
                        float& radius_eff_a, float& radius_eff_b, const HullRef& hull_a = HullRef(), const HullRef& hull_b = HullRef(),
                        const SpeculativeWriter* spec = nullptr) {
    GenericShapeData shape_data_a, shape_data_b;
    shape_data_a.shape_type = type_a;
    shape_data_b.shape_type = type_b;
    if (type_a == T_CONVEX_MESH) {  // narrow_phase.py:1096-1105: mesh pointer + cached centre of the local collision AABB
        shape_data_a.mesh_points = hull_a.points;
        shape_data_a.mesh_point_count = hull_a.count;
        shape_data_a.center = 0.5f * (hull_a.local_aabb_lower + hull_a.local_aabb_upper);
    }
    if (type_b == T_CONVEX_MESH) {
        shape_data_b.mesh_points = hull_b.points;
        shape_data_b.mesh_point_count = hull_b.count;
        shape_data_b.center = 0.5f * (hull_b.local_aabb_lower + hull_b.local_aabb_upper);
    }
    shape_data_a.scale = type_a == T_PLANE ? vec3(scale_a.x * 0.5f, scale_a.y * 0.5f, 0.0f) : scale_a;  // geom_data (collide.py:452-453)
    shape_data_b.scale = type_b == T_PLANE ? vec3(scale_b.x * 0.5f, scale_b.y * 0.5f, 0.0f) : scale_b;
    vec3 pos_a = X_a.p, pos_b = X_b.p;
    quat quat_a = X_a.q, quat_b = X_b.q;
    radius_eff_a = radius_eff_b = 0.0f;
    bool is_infinite_plane_a = type_a == T_PLANE && shape_data_a.scale.x == 0.0f && shape_data_a.scale.y == 0.0f;
    bool is_infinite_plane_b = type_b == T_PLANE && shape_data_b.scale.x == 0.0f && shape_data_b.scale.y == 0.0f;
    if (is_infinite_plane_a && is_infinite_plane_b) return 0;
    float bsphere_radius_a = 0.0f, bsphere_radius_b = 0.0f;
    if (is_infinite_plane_a || is_infinite_plane_b) {
        // compute_bounding_sphere_from_aabb (collision_core.py:552-562)
        vec3 bsphere_center_a = 0.5f * (aabb_lo_a + aabb_hi_a);
        bsphere_radius_a = length(0.5f * (aabb_hi_a - aabb_lo_a));
        vec3 bsphere_center_b = 0.5f * (aabb_lo_b + aabb_hi_b);
        bsphere_radius_b = length(0.5f * (aabb_hi_b - aabb_lo_b));
        // check_infinite_plane_bsphere_overlap (collision_core.py:640-682)
        vec3 plane_pos = is_infinite_plane_a ? pos_a : pos_b;
        quat plane_quat = is_infinite_plane_a ? quat_a : quat_b;
        vec3 other_center = is_infinite_plane_a ? bsphere_center_b : bsphere_center_a;
        float other_radius = is_infinite_plane_a ? bsphere_radius_b : bsphere_radius_a;
        // speculative + external AABBs (narrow_phase.py:1170-1175): the AABBs describe the current geometry, so the pair's search
        // extension (here gap_sum = the velocity-extended gaps) is added to the non-plane shape's overlap radius
        if (spec && spec->enabled) other_radius += gap_sum;
        vec3 plane_normal = quat_rotate(plane_quat, vec3(0.f, 0.f, 1.f));
        float center_dist = dot(other_center - plane_pos, plane_normal);
        if (!(center_dist <= other_radius)) return 0;
    }
    // convert_infinite_plane_to_cube (collision_core.py:566-626)
    auto to_cube = [](GenericShapeData& sd, quat plane_rotation, vec3 plane_position, vec3 other_position, float other_radius) {
        sd.shape_type = T_BOX;
        float lateral_size = other_radius * 10.0f;
        float depth = other_radius * 10.0f;
        sd.scale = vec3(lateral_size, lateral_size, depth);
        sd.center = vec3(0.f);
        vec3 plane_normal = quat_rotate(plane_rotation, vec3(0.f, 0.f, 1.f));
        vec3 to_other = other_position - plane_position;
        float distance_along_normal = dot(to_other, plane_normal);
        vec3 plane_surface_point = other_position - plane_normal * distance_along_normal;
        return plane_surface_point - plane_normal * depth;
    };
    vec3 pos_a_adjusted = pos_a, pos_b_adjusted = pos_b;
    if (is_infinite_plane_a) pos_a_adjusted = to_cube(shape_data_a, quat_a, pos_a, pos_b, bsphere_radius_b + gap_sum);
    if (is_infinite_plane_b) pos_b_adjusted = to_cube(shape_data_b, quat_b, pos_b, pos_a, bsphere_radius_a + gap_sum);
    PairCtx ctx;
    if (spec) ctx.spec = *spec;
    compute_gjk_mpr_contacts(ctx, shape_data_a, shape_data_b, quat_a, quat_b, pos_a_adjusted, pos_b_adjusted, gap_sum, margin_a, margin_b);
    for (int i = 0; i < ctx.count; ++i) out[i] = ctx.out[i];
    radius_eff_a = ctx.radius_eff_a;
    radius_eff_b = ctx.radius_eff_b;
    return ctx.count;
}

}  // namespace cvx
}  // namespace orc

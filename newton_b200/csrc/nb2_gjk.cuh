// nb2_gjk.cuh - device entry of the generic convex-convex path (MPR / GJK + manifold) for shape pairs without an
// analytic collider (reference geometry/narrow_phase.py:1041-1216 -> collision_core.py:700-790 find_contacts).
// The algorithm itself lives in nb2_convex.cuh (host+device, shared with the CPU oracle).
#pragma once
#include "nb2_convex.cuh"
namespace nb2 {
struct ConvexShape {
    int type;
    V3 scale;
    Xf X;
    float margin, gap;
    V3 lo, hi;  // world AABB (margin + gap included), as the broad phase used it
    // CONVEX_MESH: unscaled hull vertices + centre of the scaled local AABB (zero / null for primitives)
    const float* hull = nullptr;
    int hull_count = 0;
    V3 center;
};
// Returns a bit mask of valid entries in dist/pos/normal (up to 5 manifold contacts, emission order = sort_sub_key order).
NB2_DEV unsigned convex_pair_contacts(const ConvexShape& a, const ConvexShape& b, float* dist, V3* pos, V3* normal, float& reff_a,
                                      float& reff_b, const ConvexSpec* spec = nullptr) {
    ConvexPairIn in;
    in.spec = spec;
    in.type_a = a.type;
    in.type_b = b.type;
    in.scale_a = a.scale;
    in.scale_b = b.scale;
    in.Xa = a.X;
    in.Xb = b.X;
    in.margin_a = a.margin;
    in.margin_b = b.margin;
    in.gap_sum = a.gap + b.gap;
    in.hull_a = a.hull;
    in.hull_count_a = a.hull_count;
    in.center_a = a.center;
    in.hull_b = b.hull;
    in.hull_count_b = b.hull_count;
    in.center_b = b.center;
    ConvexPairAabbs bb{a.lo, a.hi, b.lo, b.hi};
    const int n = convex_contacts_any(in, bb, dist, pos, normal, reff_a, reff_b);
    return (1u << n) - 1u;
}
}  // namespace nb2

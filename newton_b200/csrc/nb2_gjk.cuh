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
};
// Returns a bit mask of valid entries in dist/pos/normal (up to 5 manifold contacts, emission order = sort_sub_key order).
// Pairs with a PLANE reach the generic path only for cones / barrel cylinders (narrow_phase.py:1098-1165 converts the
// infinite plane to a cube first); that conversion is not built, nb2_model_create rejects such models (nb2_api.cu).
NB2_DEV unsigned convex_pair_contacts(const ConvexShape& a, const ConvexShape& b, float* dist, V3* pos, V3* normal, float& reff_a,
                                      float& reff_b) {
    reff_a = 0.0f;
    reff_b = 0.0f;
    if (a.type == CG_PLANE || b.type == CG_PLANE) return 0u;
    ConvexPairIn in;
    in.type_a = a.type;
    in.type_b = b.type;
    in.scale_a = a.scale;
    in.scale_b = b.scale;
    in.Xa = a.X;
    in.Xb = b.X;
    in.margin_a = a.margin;
    in.margin_b = b.margin;
    in.gap_sum = a.gap + b.gap;
    const int n = convex_contacts(in, dist, pos, normal, reff_a, reff_b);
    return (1u << n) - 1u;
}
}  // namespace nb2

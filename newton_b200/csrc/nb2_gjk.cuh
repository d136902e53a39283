// nb2_gjk.cuh - convex-convex contacts (MPR / GJK + manifold) for shape pairs without an analytic collider.
// Placeholder: filled in by the GJK/MPR milestone; until then convex pairs produce no contacts.
#pragma once
#include "nb2_math.cuh"
namespace nb2 {
struct ConvexShape {
    int type;
    V3 scale;
    Xf X;
    float margin, gap, radius;
};
// Returns a bit mask of valid entries in dist/pos/normal (up to 5 manifold contacts).
NB2_DEV unsigned convex_pair_contacts(const ConvexShape&, const ConvexShape&, float*, V3*, V3*, float&, float&) { return 0u; }
}  // namespace nb2

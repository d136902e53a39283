// oracle_gjk.h - TEST INFRASTRUCTURE ONLY.  (placeholder until the GJK/MPR restatement lands)
#pragma once
#include "oracle_collide.h"
namespace orc {
inline void gjk_mpr_pairs(const nb2_model_desc&, const float*, CollideResult& res) {
    if (!res.gjk_pairs.empty()) std::fprintf(stderr, "oracle: %zu convex pairs need GJK/MPR (not implemented yet)\n", res.gjk_pairs.size());
}
inline int convex_pair_test(int, vec3, const transform&, int, vec3, const transform&, float, float*, float*, float*) { return -1; }
}  // namespace orc

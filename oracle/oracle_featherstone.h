// oracle_featherstone.h - TEST INFRASTRUCTURE ONLY.  (placeholder until the Featherstone restatement lands)
#pragma once
#include <cstdio>
#include "../include/newton_b200.h"
#include "oracle_math.h"
namespace orc {
inline void featherstone_step(const nb2_model_desc&, const nb2_featherstone_params&, const nb2_state_view&, const nb2_state_view&,
                              const nb2_control_view&, const nb2_contacts_view*, float) {
    std::fprintf(stderr, "oracle: featherstone_step not implemented yet\n");
}
}  // namespace orc

// nb2_featherstone.cu - placeholder until the Featherstone milestone lands.
#include "nb2_internal.cuh"
namespace nb2 {
nb2_status launch_featherstone_step(nb2_model*, const nb2_featherstone_params&, const nb2_state_view&, const nb2_state_view&,
                                    const nb2_control_view&, int, float, cudaStream_t) {
    set_error("nb2_featherstone_step: not implemented yet");
    return NB2_ERR_UNSUPPORTED;
}
nb2_status launch_eval_fk(nb2_model*, const float*, const float*, float*, float*, cudaStream_t) {
    set_error("nb2_eval_fk: not implemented yet");
    return NB2_ERR_UNSUPPORTED;
}
}  // namespace nb2

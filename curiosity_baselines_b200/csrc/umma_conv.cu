// tcgen05 engine entry points (placeholder until the kernels land: reports "unsupported" so
// CB_ENGINE_AUTO falls through to the SIMT engine and CB_ENGINE_UMMA fails loudly).
#include "common.cuh"
namespace cb {
bool umma_supports_fwd(const cb_conv_desc*, const cb_epilogue*) { return false; }
bool umma_supports_dgrad(const cb_conv_desc*) { return false; }
bool umma_supports_wgrad(const cb_conv_desc*) { return false; }
int umma_conv_fwd(const cb_conv_desc*, const void*, const void*, const cb_epilogue*, cudaStream_t) {
    return fail(CB_E_ARG, "umma_conv_fwd: not built");
}
int umma_conv_dgrad(const cb_conv_desc*, const void*, const void*, const void*, int, const void*, void*, cudaStream_t) {
    return fail(CB_E_ARG, "umma_conv_dgrad: not built");
}
int umma_conv_wgrad(const cb_conv_desc*, const void*, const void*, const float*, const float*, float*, float*,
                    float, cudaStream_t) {
    return fail(CB_E_ARG, "umma_conv_wgrad: not built");
}
}  // namespace cb

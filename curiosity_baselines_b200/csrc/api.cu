// Engine dispatch for the GEMM-shaped entry points + library-level queries.
#include "common.cuh"

namespace cb {
thread_local char g_err[512] = "";

int simt_conv_fwd(const cb_conv_desc*, const void*, const void*, const cb_epilogue*, cudaStream_t);
int simt_conv_dgrad(const cb_conv_desc*, const void*, const void*, const void*, int, const void*, void*, cudaStream_t);
int simt_conv_wgrad(const cb_conv_desc*, const void*, const void*, const float*, const float*, float*, float*,
                    float, cudaStream_t);
// tcgen05 engine (umma_conv.cu): return CB_E_ARG when the shape has no specialisation
int umma_conv_fwd(const cb_conv_desc*, const void*, const void*, const cb_epilogue*, cudaStream_t);
int umma_conv_dgrad(const cb_conv_desc*, const void*, const void*, const void*, int, const void*, void*, cudaStream_t);
int umma_conv_wgrad(const cb_conv_desc*, const void*, const void*, const float*, const float*, float*, float*,
                    float, cudaStream_t);
int umma_conv_dgrad_s2(const cb_conv_desc*, const void*, const void*, const void*, int, const void*, void*,
                       cudaStream_t);
bool umma_supports_dgrad_s2(const cb_conv_desc*);
bool umma_supports_fwd(const cb_conv_desc*, const cb_epilogue*);
bool umma_supports_dgrad(const cb_conv_desc*);
bool umma_supports_wgrad(const cb_conv_desc*);
int umma_grouped_fwd(int, int, int, int, int, const void*, const void*, int, const cb_epilogue*, cudaStream_t);
int umma_grouped_dgrad(int, int, int, int, const void*, const void*, const void*, int, const void*, void*, cudaStream_t);
int umma_grouped_wgrad(int, int, int, int, int, const void*, const void*, float*, float*, float, cudaStream_t);
int colsum_bf16(const void* dy, int64_t rows, int c, float* out, float beta, cudaStream_t s);
}  // namespace cb

extern "C" int cb_colsum_bf16(const void* dy, int64_t rows, int32_t c, float* out, float beta, void* stream) {
    CB_REQUIRE(dy && out && rows >= 1 && c >= 1, CB_E_ARG, "cb_colsum_bf16: bad argument");
    return cb::colsum_bf16(dy, rows, c, out, beta, cb::as_stream(stream));
}

extern "C" int cb_grouped_linear_fwd(int32_t n, int32_t groups, int32_t k, int32_t f, int32_t x_shared, const void* x,
                                     const void* w_all, int32_t act, const cb_epilogue* e, void* stream) {
    return cb::umma_grouped_fwd(n, groups, k, f, x_shared, x, w_all, act, e, cb::as_stream(stream));
}
extern "C" int cb_grouped_linear_dgrad(int32_t n, int32_t groups, int32_t k, int32_t f, const void* dy,
                                       const void* w_t_all, const void* x_saved, int32_t x_act, const void* dx_add,
                                       void* dx, void* stream) {
    return cb::umma_grouped_dgrad(n, groups, k, f, dy, w_t_all, x_saved, x_act, dx_add, dx, cb::as_stream(stream));
}
extern "C" int cb_grouped_linear_wgrad(int32_t n, int32_t groups, int32_t k, int32_t f, int32_t x_shared, const void* x,
                                       const void* dy, float* dw_all, float* dbias_all, float beta, void* stream) {
    return cb::umma_grouped_wgrad(n, groups, k, f, x_shared, x, dy, dw_all, dbias_all, beta, cb::as_stream(stream));
}

extern "C" const char* cb_last_error(void) { return cb::g_err; }
extern "C" int cb_version(void) { return 100; }

extern "C" int cb_device_supports_umma(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

extern "C" int cb_conv_fwd(const cb_conv_desc* d, const void* x, const void* w, const cb_epilogue* e,
                           void* stream) {
    CB_REQUIRE(d && e, CB_E_ARG, "cb_conv_fwd: null descriptor");
    cudaStream_t s = cb::as_stream(stream);
    CB_REQUIRE(!e->side_bf16 || (d->engine != CB_ENGINE_SIMT && cb::umma_supports_fwd(d, e) &&
                                 ((long long)d->k_h * d->k_w * d->in_c) % 64 == 0 && d->k_h == 1 && d->in_h == 1),
               CB_E_ARG, "cb_conv_fwd: side_bf16 needs a tcgen05 Linear layer with K % 64 == 0");
    if (d->engine == CB_ENGINE_SIMT) return cb::simt_conv_fwd(d, x, w, e, s);
    if (cb::umma_supports_fwd(d, e)) return cb::umma_conv_fwd(d, x, w, e, s);
    CB_REQUIRE(d->engine != CB_ENGINE_UMMA, CB_E_ARG, "cb_conv_fwd: no tcgen05 specialisation for this shape");
    return cb::simt_conv_fwd(d, x, w, e, s);
}

extern "C" int cb_conv_dgrad(const cb_conv_desc* d, const void* dy, const void* w_t, const void* x_saved,
                             int32_t x_act, const void* dx_add, void* dx, void* stream) {
    CB_REQUIRE(d, CB_E_ARG, "cb_conv_dgrad: null descriptor");
    cudaStream_t s = cb::as_stream(stream);
    if (d->engine == CB_ENGINE_SIMT) return cb::simt_conv_dgrad(d, dy, w_t, x_saved, x_act, dx_add, dx, s);
    if (cb::umma_supports_dgrad(d)) return cb::umma_conv_dgrad(d, dy, w_t, x_saved, x_act, dx_add, dx, s);
    CB_REQUIRE(d->engine != CB_ENGINE_UMMA, CB_E_ARG, "cb_conv_dgrad: no tcgen05 specialisation for this shape");
    return cb::simt_conv_dgrad(d, dy, w_t, x_saved, x_act, dx_add, dx, s);
}

extern "C" int cb_conv_wgrad(const cb_conv_desc* d, const void* x, const void* dy, const float* norm_mean,
                             const float* norm_rstd, float* dw, float* dbias, float beta, void* stream) {
    CB_REQUIRE(d, CB_E_ARG, "cb_conv_wgrad: null descriptor");
    cudaStream_t s = cb::as_stream(stream);
    if (d->engine == CB_ENGINE_SIMT) return cb::simt_conv_wgrad(d, x, dy, norm_mean, norm_rstd, dw, dbias, beta, s);
    if (cb::umma_supports_wgrad(d)) return cb::umma_conv_wgrad(d, x, dy, norm_mean, norm_rstd, dw, dbias, beta, s);
    CB_REQUIRE(d->engine != CB_ENGINE_UMMA, CB_E_ARG, "cb_conv_wgrad: no tcgen05 specialisation for this shape");
    return cb::simt_conv_wgrad(d, x, dy, norm_mean, norm_rstd, dw, dbias, beta, s);
}

// Strided host->device staging (pinned sampler buffer -> dense device tensor): lets the RND hook
// upload only the newest frame of every stack (7056 of 28224 bytes per sample) with one DMA.
extern "C" int cb_h2d_2d(void* dst, int64_t dst_pitch, const void* src_host, int64_t src_pitch,
                         int64_t width_bytes, int64_t rows, void* stream) {
    CB_REQUIRE(dst && src_host && width_bytes > 0 && rows > 0, CB_E_ARG, "cb_h2d_2d: bad argument");
    CB_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src_host, (size_t)src_pitch, (size_t)width_bytes,
                              (size_t)rows, cudaMemcpyHostToDevice, cb::as_stream(stream)));
    return CB_OK;
}

extern "C" int cb_conv_dgrad_s2_supported(const cb_conv_desc* d) {
    return d && cb::umma_supports_dgrad_s2(d) ? 1 : 0;
}

extern "C" int cb_conv_dgrad_s2(const cb_conv_desc* d, const void* dy, const void* w_shuffle, const void* x_saved,
                                int32_t x_act, const void* dx_add, void* dx, void* stream) {
    CB_REQUIRE(d, CB_E_ARG, "cb_conv_dgrad_s2: null descriptor");
    return cb::umma_conv_dgrad_s2(d, dy, w_shuffle, x_saved, x_act, dx_add, dx, cb::as_stream(stream));
}

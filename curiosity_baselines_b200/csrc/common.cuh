// Shared helpers for libcuriosity_b200: status codes, error text, bf16 conversions.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <algorithm>
#include "../../include/curiosity_b200.h"

namespace cb {

extern thread_local char g_err[512];

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CB_REQUIRE(cond, code, ...) \
    do { if (!(cond)) return cb::fail((code), __VA_ARGS__); } while (0)

#define CB_CUDA(expr)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            return cb::fail(CB_E_CUDA, "%s failed: %s (%s:%d)", #expr,                    \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                  \
    } while (0)

#define CB_LAUNCH_CHECK() CB_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float apply_act(float x, int act) {
    switch (act) {
        case CB_ACT_RELU:  return x > 0.f ? x : 0.f;
        case CB_ACT_LRELU: return x > 0.f ? x : 0.01f * x;
        case CB_ACT_ELU:   return x > 0.f ? x : expm1f(x);
        default:           return x;
    }
}
// derivative of the activation expressed through its OUTPUT y (what the forward saved)
__device__ __forceinline__ float act_grad_from_output(float y, int act) {
    switch (act) {
        case CB_ACT_RELU:  return y > 0.f ? 1.f : 0.f;
        case CB_ACT_LRELU: return y > 0.f ? 1.f : 0.01f;
        case CB_ACT_ELU:   return y > 0.f ? 1.f : y + 1.f;
        default:           return 1.f;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace cb

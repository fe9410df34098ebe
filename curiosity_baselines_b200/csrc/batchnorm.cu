// BatchNorm2d of the curiosity encoders (rlpyt/models/curiosity/encoders.py:23-25, 87-96: the optional
// nn.BatchNorm2d after every LeakyReLU / ELU, `-batch_norm`), on bf16 NHWC activations seen as [rows, C] with
// rows = N*H*W: torch's training-mode arithmetic (batch mean and biased variance over all rows of a channel) split
// into HBM-bound passes.  The few per-channel scalars in between (mean, 1/sqrt(var+eps), the running-statistics
// update, the three backward coefficients) are computed by the host side on C-element tensors.
//
//   cb_bn_moments      sum_r x, sum_r x^2                         per channel, accumulated in double
//   cb_bn_apply        z = x * scale[c] + shift[c]                (scale = gamma*invstd, shift = beta - mean*scale)
//   cb_bn_bwd_moments  sum_r dz, sum_r dz * (x - mean) * invstd   (= dbeta, dgamma)
//   cb_bn_bwd_apply    dx = (dz*a[c] + x*b[c] + c[c]) * act'(x)   (x is the OUTPUT of the activation in front of the
//                      normalisation, so the activation's derivative is folded in, as the conv dgrad epilogues do)
//
// Each thread owns one 8-channel group (one 16-byte load per row) and walks the rows with a grid stride; per-thread
// fp32 partial sums cover at most a few thousand rows before they are combined in double.
#include "common.cuh"

namespace {

using namespace cb;

constexpr int kThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        v[2 * e] = f.x; v[2 * e + 1] = f.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    return u;
}

// BWD = false: (x, x^2);  BWD = true: (dz, dz * (x - mean) * invstd)
template <bool BWD>
__global__ void __launch_bounds__(kThreads) bn_moments_kernel(const uint4* __restrict__ a, const uint4* __restrict__ x,
                                                              long long rows, int groups, const float* __restrict__ mean,
                                                              const float* __restrict__ invstd, double* __restrict__ s0,
                                                              double* __restrict__ s1) {
    __shared__ float sh0[kThreads * 8], sh1[kThreads * 8];
    const int g = threadIdx.x % groups;                      // this thread's 8-channel group
    const int rows_per_block = kThreads / groups;
    const long long r0 = (long long)blockIdx.x * rows_per_block + threadIdx.x / groups;
    const long long stride = (long long)gridDim.x * rows_per_block;
    float p0[8], p1[8], m[8], is[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        p0[e] = p1[e] = 0.f;
        m[e] = BWD ? __ldg(mean + g * 8 + e) : 0.f;
        is[e] = BWD ? __ldg(invstd + g * 8 + e) : 0.f;
    }
    for (long long r = r0; r < rows; r += stride) {
        float v[8];
        unpack8(__ldg(a + r * groups + g), v);
        if (BWD) {
            float xv[8];
            unpack8(__ldg(x + r * groups + g), xv);
#pragma unroll
            for (int e = 0; e < 8; ++e) { p0[e] += v[e]; p1[e] = fmaf(v[e], (xv[e] - m[e]) * is[e], p1[e]); }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) { p0[e] += v[e]; p1[e] = fmaf(v[e], v[e], p1[e]); }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { sh0[threadIdx.x * 8 + e] = p0[e]; sh1[threadIdx.x * 8 + e] = p1[e]; }
    __syncthreads();
    // one thread per channel adds the block's partial sums (rows_per_block of them) in double
    const int C = groups * 8;
    if ((int)threadIdx.x < C) {
        const int c = threadIdx.x, cg = c >> 3, ce = c & 7;
        double t0 = 0.0, t1 = 0.0;
        for (int i = 0; i < rows_per_block; ++i) {
            t0 += (double)sh0[(i * groups + cg) * 8 + ce];
            t1 += (double)sh1[(i * groups + cg) * 8 + ce];
        }
        atomicAdd(s0 + c, t0);
        atomicAdd(s1 + c, t1);
    }
}

__global__ void __launch_bounds__(kThreads) bn_apply_kernel(const uint4* __restrict__ x, long long items, int groups,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            uint4* __restrict__ out_bf16, float4* __restrict__ out_f32) {
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < items; i += (long long)gridDim.x * kThreads) {
        const int g = (int)(i % groups);
        float v[8];
        unpack8(__ldg(x + i), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], __ldg(scale + g * 8 + e), __ldg(shift + g * 8 + e));
        if (out_bf16) out_bf16[i] = pack8(v);
        if (out_f32) {
            out_f32[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
            out_f32[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
}

__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ x,
                                                                long long items, int groups, const float* __restrict__ ca,
                                                                const float* __restrict__ cb_, const float* __restrict__ cc,
                                                                int act, uint4* __restrict__ dx) {
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < items; i += (long long)gridDim.x * kThreads) {
        const int g = (int)(i % groups);
        float d[8], xv[8];
        unpack8(__ldg(dz + i), d);
        unpack8(__ldg(x + i), xv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = g * 8 + e;
            d[e] = (fmaf(d[e], __ldg(ca + c), fmaf(xv[e], __ldg(cb_ + c), __ldg(cc + c)))) * act_grad_from_output(xv[e], act);
        }
        dx[i] = pack8(d);
    }
}

int check(const void* p, long long rows, int c, const char* what) {
    CB_REQUIRE(p, CB_E_ARG, "%s: null pointer", what);
    CB_REQUIRE(rows > 0 && c >= 8 && c % 8 == 0 && c <= kThreads && kThreads % (c / 8) == 0, CB_E_ARG,
               "%s: rows %lld, channels %d (need 8 * a power of two, <= %d)", what, rows, c, kThreads);
    CB_REQUIRE(((uintptr_t)p & 15) == 0, CB_E_ALIGN, "%s: activations must be 16-byte aligned", what);
    return CB_OK;
}

int grid_for(long long work_items) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (work_items + kThreads - 1) / kThreads;
    const long long cap = (long long)sms * 8;
    return (int)std::max(1ll, std::min(blocks, cap));
}

}  // namespace

extern "C" int cb_bn_moments(const void* x_bf16, int64_t rows, int32_t c, double* sum, double* sumsq, void* stream) {
    int rc = check(x_bf16, rows, c, "cb_bn_moments"); if (rc) return rc;
    CB_REQUIRE(sum && sumsq, CB_E_ARG, "cb_bn_moments: null output");
    const int groups = c / 8;
    bn_moments_kernel<false><<<grid_for(rows * groups), kThreads, 0, as_stream(stream)>>>(
        (const uint4*)x_bf16, nullptr, rows, groups, nullptr, nullptr, sum, sumsq);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_bn_bwd_moments(const void* dz_bf16, const void* x_bf16, int64_t rows, int32_t c, const float* mean,
                                 const float* invstd, double* sum_dz, double* sum_dz_xhat, void* stream) {
    int rc = check(dz_bf16, rows, c, "cb_bn_bwd_moments"); if (rc) return rc;
    rc = check(x_bf16, rows, c, "cb_bn_bwd_moments"); if (rc) return rc;
    CB_REQUIRE(mean && invstd && sum_dz && sum_dz_xhat, CB_E_ARG, "cb_bn_bwd_moments: null pointer");
    const int groups = c / 8;
    bn_moments_kernel<true><<<grid_for(rows * groups), kThreads, 0, as_stream(stream)>>>(
        (const uint4*)dz_bf16, (const uint4*)x_bf16, rows, groups, mean, invstd, sum_dz, sum_dz_xhat);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_bn_apply(const void* x_bf16, int64_t rows, int32_t c, const float* scale, const float* shift,
                           void* out_bf16, float* out_f32, void* stream) {
    int rc = check(x_bf16, rows, c, "cb_bn_apply"); if (rc) return rc;
    CB_REQUIRE(scale && shift && (out_bf16 || out_f32), CB_E_ARG, "cb_bn_apply: null pointer");
    CB_REQUIRE((((uintptr_t)out_bf16 | (uintptr_t)out_f32) & 15) == 0, CB_E_ALIGN, "cb_bn_apply: outputs must be 16-byte aligned");
    const int groups = c / 8;
    bn_apply_kernel<<<grid_for(rows * groups), kThreads, 0, as_stream(stream)>>>(
        (const uint4*)x_bf16, rows * groups, groups, scale, shift, (uint4*)out_bf16, (float4*)out_f32);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_bn_bwd_apply(const void* dz_bf16, const void* x_bf16, int64_t rows, int32_t c, const float* coef_dz,
                               const float* coef_x, const float* coef_1, int32_t x_act, void* dx_bf16, void* stream) {
    int rc = check(dz_bf16, rows, c, "cb_bn_bwd_apply"); if (rc) return rc;
    rc = check(x_bf16, rows, c, "cb_bn_bwd_apply"); if (rc) return rc;
    rc = check(dx_bf16, rows, c, "cb_bn_bwd_apply"); if (rc) return rc;
    CB_REQUIRE(coef_dz && coef_x && coef_1, CB_E_ARG, "cb_bn_bwd_apply: null pointer");
    const int groups = c / 8;
    bn_bwd_apply_kernel<<<grid_for(rows * groups), kThreads, 0, as_stream(stream)>>>(
        (const uint4*)dz_bf16, (const uint4*)x_bf16, rows * groups, groups, coef_dz, coef_x, coef_1, x_act, (uint4*)dx_bf16);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

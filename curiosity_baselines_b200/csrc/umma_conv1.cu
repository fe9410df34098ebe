// First convolution of the Burda / RND encoders on tcgen05: 8x8, stride 4, uint8 NCHW frames.
//   encoders.py:83-84, rnd.py:55-59 (forward);  its weight gradient for the backward pass.
//
// The frames stay uint8 in HBM (7056 B per channel).  Converter warps read each output position's
// 8x8 patch (two 4-byte loads per patch row), apply the optional RND normalisation
// clamp((x-mean)*rstd, -5, 5) (rnd.py:169-170), convert to bf16 and write the row straight into the
// 128-byte-swizzled K-major layout tcgen05 reads -- one 16-byte store per patch row.  K order is
// (c, kh, kw), i.e. torch's own weight layout, so one k-block of 64 is one input channel.
//
//   forward : D[128 positions, 32] = A[128, 64*cin] * W^T        (+bias, LeakyReLU -> bf16 NHWC)
//   wgrad   : dW[32][cin*64] += sum over positions: the same patch tile read MN-major
//             (M = 64 patch elements, K = positions) times dy rows staged the same way.
//
// Warp roles: 0-3 converters (one tile row per thread), 4 = TMEM owner + MMA issuer, 5-8 epilogue.
#include "common.cuh"
#include "umma_common.cuh"

namespace {

constexpr int kThreads = 288;
constexpr int kMaxStages1 = 4;
constexpr int kMaxCin = 4;

struct Conv1Params {
    const uint8_t* x; long long stride_n, stride_c; int h, w, cin, oh, ow;
    long long total_rows; int tiles, stages;
    const float* mean; const float* rstd;
    // forward
    const __nv_bfloat16* w_bf16;        // [32][cin*64] K-major (torch layout cast to bf16)
    const float* bias; int act; __nv_bfloat16* out;
    // wgrad
    const __nv_bfloat16* dy; float* dw;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// builds row r of the A tile (all k-blocks) in shared memory; a_stage = generic pointer to the stage
__device__ __forceinline__ void convert_row(const Conv1Params& p, uint8_t* a_stage, int r, long long grow) {
    const bool valid = grow < p.total_rows;
    int n = 0, oy = 0, ox = 0;
    if (valid) {
        const int pp = p.oh * p.ow;
        n = (int)(grow / pp);
        const int rem = (int)(grow - (long long)n * pp);
        oy = rem / p.ow; ox = rem - oy * p.ow;
    }
    const int px0 = (4 * oy) * p.w + 4 * ox;
    for (int c = 0; c < p.cin; ++c) {
        const uint8_t* src = p.x + n * p.stride_n + c * p.stride_c + px0;
        uint8_t* dst_row = a_stage + c * 16384 + r * 128;
        // all loads of the patch first (8 rows x 8 bytes, plus the statistics), so their latencies overlap
        uint32_t lo[8], hi[8];
#pragma unroll
        for (int kh = 0; kh < 8; ++kh) {
            lo[kh] = valid ? __ldg(reinterpret_cast<const uint32_t*>(src + kh * p.w)) : 0u;
            hi[kh] = valid ? __ldg(reinterpret_cast<const uint32_t*>(src + kh * p.w + 4)) : 0u;
        }
        if (p.mean && valid) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float4 m[8], sd[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int kh = half * 4 + q;
                    const float4* m4 = reinterpret_cast<const float4*>(p.mean + px0 + kh * p.w);
                    const float4* r4 = reinterpret_cast<const float4*>(p.rstd + px0 + kh * p.w);
                    m[2 * q] = __ldg(m4); m[2 * q + 1] = __ldg(m4 + 1);
                    sd[2 * q] = __ldg(r4); sd[2 * q + 1] = __ldg(r4 + 1);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int kh = half * 4 + q;
                    const float mm[8] = {m[2 * q].x, m[2 * q].y, m[2 * q].z, m[2 * q].w,
                                         m[2 * q + 1].x, m[2 * q + 1].y, m[2 * q + 1].z, m[2 * q + 1].w};
                    const float ss[8] = {sd[2 * q].x, sd[2 * q].y, sd[2 * q].z, sd[2 * q].w,
                                         sd[2 * q + 1].x, sd[2 * q + 1].y, sd[2 * q + 1].z, sd[2 * q + 1].w};
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v[j] = (float)((lo[kh] >> (8 * j)) & 0xffu);
                        v[4 + j] = (float)((hi[kh] >> (8 * j)) & 0xffu);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fminf(fmaxf((v[j] - mm[j]) * ss[j], -5.f), 5.f);
                    *reinterpret_cast<uint4*>(dst_row + ((kh ^ (r & 7)) << 4)) =
                        make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                }
            }
        } else {
#pragma unroll
            for (int kh = 0; kh < 8; ++kh) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = (float)((lo[kh] >> (8 * j)) & 0xffu);
                    v[4 + j] = (float)((hi[kh] >> (8 * j)) & 0xffu);
                }
                *reinterpret_cast<uint4*>(dst_row + ((kh ^ (r & 7)) << 4)) =
                    make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
            }
        }
    }
}

template <bool WGRAD>
__global__ void __launch_bounds__(kThreads, 1) umma_conv1_kernel(const Conv1Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (smem_base - umma::smem_u32(smem_raw));
    // layout: [W: cin*4096 (fwd)] [stages: A cin*16384 (+ B 16384 for wgrad)] [barriers]
    const uint32_t w_bytes = WGRAD ? 0u : (uint32_t)p.cin * 4096u;
    const uint32_t a_bytes = (uint32_t)p.cin * 16384u;
    const uint32_t stage_bytes = a_bytes + (WGRAD ? 16384u : 0u);
    const uint32_t st_base = smem_base + w_bytes;
    const int kStages = p.stages;
    const uint32_t bar_base = st_base + kStages * stage_bytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages1 + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages1 + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages1 + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages1 + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // weights (forward): converter threads copy W[32][cin*64] into the swizzled K-major layout
    if (!WGRAD) {
        for (int i = threadIdx.x; i < p.cin * 32 * 8; i += blockDim.x) {      // 16-byte chunks
            const int chunk = i & 7, row = (i >> 3) & 31, c = i >> 8;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.w_bf16 + (long long)row * p.cin * 64 + c * 64) + chunk);
            *reinterpret_cast<uint4*>(gen + c * 4096 + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
        }
    } else {      // zero the dy stages once: their upper 64 bytes of every row stay zero (N padded 32 -> 64)
        for (int s = 0; s < kStages; ++s) {
            uint4* z = reinterpret_cast<uint4*>(gen + w_bytes + s * stage_bytes + a_bytes);
            for (int i = threadIdx.x; i < 1024; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { umma::mbar_init(full_bar(s), 128); umma::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { umma::mbar_init(tfull_bar(a), 1); umma::mbar_init(tempty_bar(a), 4); }
        umma::fence_barrier_init();
    }
    if (warp == 4) umma::tmem_alloc(tmem_slot, 256);
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < 4) {
        // ------------------------------------------------------------ converters
        const int r = threadIdx.x;
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            umma::mbar_wait(empty_bar(stage), phase ^ 1u);
            uint8_t* a_stage = gen + w_bytes + stage * stage_bytes;
            const long long grow = (long long)tile * 128 + r;
            convert_row(p, a_stage, r, grow);
            if (WGRAD) {
                uint8_t* b_row = a_stage + a_bytes + r * 128;
                const bool valid = grow < p.total_rows;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (valid) v = __ldg(reinterpret_cast<const uint4*>(p.dy + grow * 32) + j);
                    *reinterpret_cast<uint4*>(b_row + ((j ^ (r & 7)) << 4)) = v;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            umma::mbar_arrive(full_bar(stage));
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
    } else if (warp == 4) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            bool first = true;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                if (!WGRAD) umma::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                umma::mbar_wait(full_bar(stage), phase);
                umma::fence_after_thread_sync();
                const uint32_t sa = st_base + stage * stage_bytes;
                if (!WGRAD) {
                    constexpr uint32_t idesc = umma::make_idesc(128, 32, 0, 0);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 32);
                    for (int c = 0; c < p.cin; ++c)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma::mma_f16_ss(d_tmem, umma::make_smem_desc(sa + c * 16384 + k * 32, 16, 1024),
                                             umma::make_smem_desc(smem_base + c * 4096 + k * 32, 16, 1024), idesc,
                                             (c | k) ? 1u : 0u);
                    umma::mma_commit(empty_bar(stage));
                    umma::mma_commit(tfull_bar(acc));
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                } else {
                    constexpr uint32_t idesc = umma::make_idesc(64, 64, 1, 1);
                    for (int c = 0; c < p.cin; ++c)
#pragma unroll
                        for (int k = 0; k < 8; ++k)     // 128 positions / 16
                            umma::mma_f16_ss(tmem_base + (uint32_t)(c * 64),
                                             umma::make_smem_desc(sa + c * 16384 + k * 2048, 8192, 1024),
                                             umma::make_smem_desc(sa + p.cin * 16384 + k * 2048, 8192, 1024), idesc,
                                             (first && k == 0) ? 0u : 1u);
                    first = false;
                    umma::mma_commit(empty_bar(stage));
                }
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            if (WGRAD) umma::mma_commit(tfull_bar(0));
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 5..8)
        const int quarter = warp & 3;
        if (!WGRAD) {
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                umma::mbar_wait(tfull_bar(acc), acc_phase);
                umma::fence_after_thread_sync();
                uint32_t raw[32];
                umma::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 32), raw);
                umma::tmem_ld_wait();
                const long long grow = (long long)tile * 128 + quarter * 32 + lane;
                if (grow < p.total_rows) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
                    if (p.bias) {
                        const float4* b4 = reinterpret_cast<const float4*>(p.bias);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 b = __ldg(b4 + q);
                            v[q * 4] += b.x; v[q * 4 + 1] += b.y; v[q * 4 + 2] += b.z; v[q * 4 + 3] += b.w;
                        }
                    }
                    if (p.act == CB_ACT_LRELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.01f * v[j]);
                    } else if (p.act == CB_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                    } else if (p.act == CB_ACT_ELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : expm1f(v[j]);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(p.out + grow * 32);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        dst[q] = make_uint4(pack2(v[q * 8], v[q * 8 + 1]), pack2(v[q * 8 + 2], v[q * 8 + 3]),
                                            pack2(v[q * 8 + 4], v[q * 8 + 5]), pack2(v[q * 8 + 6], v[q * 8 + 7]));
                }
                umma::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        } else if (blockIdx.x < p.tiles) {
            // M = 64 accumulator: patch element i lives in TMEM lane (i/16)*32 + i%16
            umma::mbar_wait(tfull_bar(0), 0);
            umma::fence_after_thread_sync();
            const int K = p.cin * 64;
            for (int c = 0; c < p.cin; ++c) {
                uint32_t raw[32];
                umma::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 64), raw);
                umma::tmem_ld_wait();
                if (lane < 16) {
                    const int i = quarter * 16 + lane;
#pragma unroll
                    for (int co = 0; co < 32; ++co) atomicAdd(p.dw + (long long)co * K + c * 64 + i, __uint_as_float(raw[co]));
                }
            }
        }
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    if (warp == 4) { umma::fence_after_thread_sync(); umma::tmem_dealloc(tmem_base, 256); }
}

bool shape_ok(const cb_conv_desc* d) {
    return d->in_dtype == CB_U8 && d->k_h == 8 && d->k_w == 8 && d->stride == 4 && d->pad == 0 && d->out_c == 32 &&
           d->in_c >= 1 && d->in_c <= kMaxCin && d->in_stride_w == 1 && d->in_stride_h == d->in_w && d->in_w % 4 == 0 &&
           d->in_stride_n % 4 == 0 && d->in_stride_c % 4 == 0;
}

int fill(Conv1Params& p, const cb_conv_desc* d, const void* x, const float* mean, const float* rstd) {
    CB_REQUIRE(((uintptr_t)x & 3) == 0, CB_E_ALIGN, "conv1: frames must be 4-byte aligned");
    CB_REQUIRE(!mean || ((((uintptr_t)mean | (uintptr_t)rstd) & 15) == 0), CB_E_ALIGN, "conv1: mean/rstd 16-byte aligned");
    p.x = (const uint8_t*)x; p.stride_n = d->in_stride_n; p.stride_c = d->in_stride_c;
    p.h = d->in_h; p.w = d->in_w; p.cin = d->in_c; p.oh = d->out_h; p.ow = d->out_w;
    p.total_rows = (long long)d->n * d->out_h * d->out_w;
    p.tiles = (int)((p.total_rows + 127) / 128);
    p.mean = mean; p.rstd = rstd;
    return CB_OK;
}

template <bool WGRAD>
int launch(Conv1Params& p, cudaStream_t s) {
    const int stage_bytes = p.cin * 16384 + (WGRAD ? 16384 : 0);
    const int fixed = (WGRAD ? 0 : p.cin * 4096) + 1024 + 256;
    p.stages = (220 * 1024 - fixed) / stage_bytes;
    if (p.stages > kMaxStages1) p.stages = kMaxStages1;
    CB_REQUIRE(p.stages >= 2, CB_E_ARG, "conv1: tile does not fit shared memory");
    const int smem = fixed + p.stages * stage_bytes;
    static bool attr = false;
    if (!attr) {
        CB_CUDA(cudaFuncSetAttribute(umma_conv1_kernel<WGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = p.tiles < sms ? p.tiles : sms;
    umma_conv1_kernel<WGRAD><<<grid, kThreads, smem, s>>>(p);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // namespace

namespace cb {

int scale_buffer(float* x, int64_t n, float beta, cudaStream_t s);
int colsum_bf16(const void* dy, int64_t rows, int c, float* out, float beta, cudaStream_t s);

bool conv1_supports(const cb_conv_desc* d) { return cb_device_supports_umma() && shape_ok(d); }

// w_native: bf16 copy of the torch weight [32][cin][8][8] (K order c, kh, kw)
int conv1_fwd(const cb_conv_desc* d, const void* x, const void* w_native, const cb_epilogue* e, cudaStream_t s) {
    CB_REQUIRE(x && w_native && e && e->out_bf16 && !e->out_f32 && !e->residual && !e->side_dim, CB_E_ARG,
               "conv1 forward: needs a bf16 output and no residual/side operands");
    Conv1Params p{};
    int rc = fill(p, d, x, e->norm_mean, e->norm_rstd); if (rc) return rc;
    p.w_bf16 = (const __nv_bfloat16*)w_native; p.bias = e->bias; p.act = d->act; p.out = (__nv_bfloat16*)e->out_bf16;
    return launch<false>(p, s);
}

// dw: fp32 [32][cin*64] in the torch layout (no re-ordering needed)
int conv1_wgrad(const cb_conv_desc* d, const void* x, const void* dy, const float* mean, const float* rstd, float* dw,
                float* dbias, float beta, cudaStream_t s) {
    CB_REQUIRE(x && dy && dw, CB_E_ARG, "conv1 wgrad: null pointer");
    Conv1Params p{};
    int rc = fill(p, d, x, mean, rstd); if (rc) return rc;
    p.dy = (const __nv_bfloat16*)dy; p.dw = dw;
    rc = scale_buffer(dw, (int64_t)32 * d->in_c * 64, beta, s); if (rc) return rc;
    if (dbias) { rc = colsum_bf16(dy, p.total_rows, 32, dbias, beta, s); if (rc) return rc; }
    return launch<true>(p, s);
}

}  // namespace cb

extern "C" int cb_conv1_supported(const cb_conv_desc* d) { return d && cb::conv1_supports(d) ? 1 : 0; }

extern "C" int cb_conv1_fwd(const cb_conv_desc* d, const void* x, const void* w_native, const cb_epilogue* e,
                            void* stream) {
    CB_REQUIRE(d && cb::conv1_supports(d), CB_E_ARG, "cb_conv1_fwd: unsupported shape or device");
    return cb::conv1_fwd(d, x, w_native, e, cb::as_stream(stream));
}

extern "C" int cb_conv1_wgrad(const cb_conv_desc* d, const void* x, const void* dy, const float* norm_mean,
                              const float* norm_rstd, float* dw_native, float* dbias, float beta, void* stream) {
    CB_REQUIRE(d && cb::conv1_supports(d), CB_E_ARG, "cb_conv1_wgrad: unsupported shape or device");
    return cb::conv1_wgrad(d, x, dy, norm_mean, norm_rstd, dw_native, dbias, beta, cb::as_stream(stream));
}

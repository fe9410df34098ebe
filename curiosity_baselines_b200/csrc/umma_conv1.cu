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
#include <stdlib.h>

namespace cb {
int encode_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle);
}

namespace {

constexpr int kConvWarps = 8;           // converter warps
constexpr int kConvThreads = kConvWarps * 32;
constexpr int kMmaWarp = kConvWarps;   // MMA issuer / TMEM owner
constexpr int kStagePitch = 80;         // bytes per staged output row (64 + 16: conflict-free 16-byte stores)
constexpr int kEpiWarps = 8;            // two per TMEM lane quarter, splitting the (row tile, net) items
constexpr int kThreads = (kConvWarps + 1 + kEpiWarps) * 32;
constexpr int kMaxRaw = 8;             // raw uint8 frames in flight per CTA
constexpr int kMaxStages1 = 4;
constexpr int kMaxCin = 4;
constexpr int kConvBar = 1;            // named barrier of the converter threads

struct Conv1Params {
    const uint8_t* x; long long stride_n, stride_c; int h, w, cin, oh, ow, n;
    int rows, row_tiles;               // oh*ow output positions per sample, ceil(rows/128)
    int pitch;                         // bytes per row of the normalised bf16 frame in shared memory
    int units, stages, a_stage_bytes;  // units = n*cin (one channel of one sample each)
    int raw_slots;                     // s2d kernels: depth of the raw-frame ring
    int s2d_stage_bytes;               // s2d forward: bytes of one pipeline stage (CPK channels of one sample)
    const float* mean; const float* rstd;
    // forward: one or two networks sharing the input (RND target + predictor)
    int nets; const __nv_bfloat16* w_bf16[2]; const float* bias[2]; int act; __nv_bfloat16* out[2];
    // wgrad
    const __nv_bfloat16* dy; float* dw; float* dbias;
    int debug;                         // CB_CONV1_DEBUG bits (profiling only): 1 no stats, 2 no stores, 4 no assemble, 8 no frame loads
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// Four uint8 pixels -> fp32 without I2F (a quarter-rate conversion-pipe instruction: 7056 of them per frame
// cost ~440 cycles per SM).  PRMT builds the bits of 2^23 + x (exact in fp32), one FADD removes 2^23.
__device__ __forceinline__ void u8x4_to_f32(uint32_t px, float (&v)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __uint_as_float(__byte_perm(px, 0x4B000000u, 0x7540u | j)) - 8388608.f;
}

__device__ __forceinline__ void conv_bar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(kConvBar), "n"(kConvThreads) : "memory"); }

// global -> shared bulk copy (TMA, no tensor map): `bytes` % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kConvGroups = 2;                         // s2d forward: converter warp groups working on alternate units
constexpr int kGrpThreads = kConvThreads / kConvGroups;
constexpr int kGrpMaxGroups = 14;                      // 4-pixel groups per thread of a group (84x84 -> 1764 / 128)
constexpr int kMaxGroups = 8;          // 4-pixel groups per converter thread: ceil(h*w/4/256) (84x84 -> 7)

// one channel of one sample: global uint8 -> registers (issued early so the latency overlaps other work)
__device__ __forceinline__ void frame_prefetch(const Conv1Params& p, int unit, int tid, uint32_t (&px)[kMaxGroups]) {
    const int groups = p.h * p.w / 4;
    const int s = unit / p.cin, c = unit - s * p.cin;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.x + s * p.stride_n + c * p.stride_c);
#pragma unroll
    for (int i = 0; i < kMaxGroups; ++i) {
        const int g = tid + i * kConvThreads;
        px[i] = (unit < p.units && g < groups && !(p.debug & 8)) ? __ldg(src + g) : 0u;
    }
}

// registers -> normalised bf16 frame F[h][pitch] in shared memory
__device__ __forceinline__ void frame_store(const Conv1Params& p, int tid, const uint32_t (&px)[kMaxGroups], uint8_t* F,
                                            const float4* s_mean, const float4* s_rstd) {
    const int groups = p.h * p.w / 4, w4 = p.w / 4;
    int y = tid / w4, x4 = tid - y * w4;
    const int dy = kConvThreads / w4, dx = kConvThreads - dy * w4;
#pragma unroll
    for (int i = 0; i < kMaxGroups; ++i) {
        const int g = tid + i * kConvThreads;
        if (g < groups) {
            float v[4];
            u8x4_to_f32(px[i], v);
            if (p.mean && !(p.debug & 1)) {
                const float4 m = s_mean[g], r = s_rstd[g];        // statistics staged in shared memory
                v[0] = fminf(fmaxf((v[0] - m.x) * r.x, -5.f), 5.f);
                v[1] = fminf(fmaxf((v[1] - m.y) * r.y, -5.f), 5.f);
                v[2] = fminf(fmaxf((v[2] - m.z) * r.z, -5.f), 5.f);
                v[3] = fminf(fmaxf((v[3] - m.w) * r.w, -5.f), 5.f);
            }
            *reinterpret_cast<uint2*>(F + y * p.pitch + x4 * 8) = make_uint2(pack2(v[0], v[1]), pack2(v[2], v[3]));
        }
        y += dy; x4 += dx;
        if (x4 >= w4) { x4 -= w4; ++y; }
    }
}

// F -> patch tile: row (oy, ox), 16-byte chunk kh = 8 bf16 of frame row 4*oy+kh starting at pixel 4*ox
__device__ __forceinline__ void assemble(const Conv1Params& p, int tid, const uint8_t* F, uint8_t* A) {
    const int kh = tid & 7;
    int row = tid >> 3;
    int oy = row / p.ow, ox = row - oy * p.ow;
    constexpr int RS = kConvThreads / 8;       // rows covered per sweep
    const int step_y = RS / p.ow, step_x = RS - step_y * p.ow;
    for (; row < p.rows; row += RS) {
        const uint2* src = reinterpret_cast<const uint2*>(F + (4 * oy + kh) * p.pitch + 8 * ox);
        const uint2 lo = src[0], hi = src[1];
        *reinterpret_cast<uint4*>(A + row * 128 + ((kh ^ (row & 7)) << 4)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
        oy += step_y; ox += step_x;
        if (ox >= p.ow) { ox -= p.ow; ++oy; }
    }
}

template <bool WGRAD>
__global__ void __launch_bounds__(kThreads, 1) umma_conv1_kernel(const Conv1Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (smem_base - umma::smem_u32(smem_raw));
    // layout: [W (fwd): cin * N*128] [stages: A (+ dy rows for wgrad)] [F] [barriers]
    const int N = WGRAD ? 64 : 32 * p.nets;
    const uint32_t w_bytes = WGRAD ? 0u : (uint32_t)(p.cin * N * 128);
    const uint32_t stage_bytes = (uint32_t)p.a_stage_bytes * (WGRAD ? 2u : 1u);
    const uint32_t st_base = smem_base + w_bytes;
    const uint32_t f_off = w_bytes + p.stages * stage_bytes;
    const uint32_t f_bytes = (uint32_t)((p.h * p.pitch + 1023) & ~1023);
    const uint32_t t_off = f_off + f_bytes;                                   // per-pixel mean | rstd (fp32)
    const uint32_t t_bytes = p.mean ? (uint32_t)((p.h * p.w * 8 + 1023) & ~1023) : 0u;
    const uint32_t o_off = t_off + t_bytes;                                   // epilogue staging: 32 rows x 80 B per warp
    const uint32_t o_bytes = WGRAD ? 0u : (uint32_t)(kEpiWarps * kStagePitch * 32);
    const uint32_t bar_base = smem_base + o_off + ((o_bytes + 1023u) & ~1023u);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages1 + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages1 + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages1 + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages1 + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - smem_base));
    const uint32_t tmem_cols = WGRAD ? 256u : (N == 64 ? 512u : 256u);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // zero all stages once (rows past `rows` and the padded half of the dy rows stay zero / finite)
    {
        uint4* z = reinterpret_cast<uint4*>(gen + w_bytes);
        for (uint32_t i = threadIdx.x; i < p.stages * stage_bytes / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    }
    if (!WGRAD) {     // weights: [net][32][cin*64] -> per channel a K-major swizzled [N rows][64] block
        for (int i = threadIdx.x; i < p.cin * N * 8; i += blockDim.x) {
            const int chunk = i & 7, row = (i >> 3) % N, c = (i >> 3) / N;
            const __nv_bfloat16* wsrc = p.w_bf16[row >> 5] + (long long)(row & 31) * p.cin * 64 + c * 64;
            *reinterpret_cast<uint4*>(gen + c * N * 128 + row * 128 + ((chunk ^ (row & 7)) << 4)) =
                __ldg(reinterpret_cast<const uint4*>(wsrc) + chunk);
        }
    }
    if (p.mean) {
        float4* tm = reinterpret_cast<float4*>(gen + t_off);
        const int g4 = p.h * p.w / 4;
        for (int i = threadIdx.x; i < g4; i += blockDim.x) {
            tm[i] = __ldg(reinterpret_cast<const float4*>(p.mean) + i);
            tm[g4 + i] = __ldg(reinterpret_cast<const float4*>(p.rstd) + i);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { umma::mbar_init(full_bar(s), kConvThreads); umma::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { umma::mbar_init(tfull_bar(a), 1); umma::mbar_init(tempty_bar(a), kEpiWarps); }
        umma::fence_barrier_init();
    }
    if (warp == kMmaWarp) umma::tmem_alloc(tmem_slot, tmem_cols);
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // CTA b owns the samples b, b+grid, ... ; a unit is one input channel of one sample
    if (warp < kConvWarps) {
        // ------------------------------------------------------------ converters
        const int tid = threadIdx.x;
        uint8_t* F = gen + f_off;
        const float4* s_mean = reinterpret_cast<const float4*>(gen + t_off);
        const float4* s_rstd = s_mean + p.h * p.w / 4;
        int stage = 0; uint32_t phase = 0;
        uint32_t px[kMaxGroups];
        int s = blockIdx.x, c = 0;
        frame_prefetch(p, s * p.cin + c, tid, px);
        while (s < p.n) {
            frame_store(p, tid, px, F, s_mean, s_rstd);
            conv_bar_sync();                                   // F complete
            int s2 = s, c2 = c + 1;
            if (c2 == p.cin) { c2 = 0; s2 += gridDim.x; }
            frame_prefetch(p, s2 < p.n ? s2 * p.cin + c2 : p.units, tid, px);      // next unit's pixels in flight
            umma::mbar_wait(empty_bar(stage), phase ^ 1u);
            uint8_t* A = gen + w_bytes + stage * stage_bytes;
            if (!(p.debug & 4)) assemble(p, tid, F, A);
            if (WGRAD) {       // dy rows of this sample: 64 B each into the lower half of 128-byte rows
                uint8_t* Bs = A + p.a_stage_bytes;
                const __nv_bfloat16* dys = p.dy + (long long)s * p.rows * 32;
                for (int i = tid; i < p.rows * 4; i += kConvThreads) {
                    const int row = i >> 2, j = i & 3;
                    *reinterpret_cast<uint4*>(Bs + row * 128 + ((j ^ (row & 7)) << 4)) =
                        __ldg(reinterpret_cast<const uint4*>(dys + row * 32) + j);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            umma::mbar_arrive(full_bar(stage));
            conv_bar_sync();                                   // everyone has finished reading F
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            s = s2; c = c2;
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            bool seen[kMaxCin] = {false, false, false, false};
            const uint32_t idesc_f = N == 64 ? umma::make_idesc(128, 64, 0, 0) : umma::make_idesc(128, 32, 0, 0);
            constexpr uint32_t idesc_w = umma::make_idesc(64, 64, 1, 1);
            for (int s = blockIdx.x; s < p.n; s += gridDim.x) {
                if (!WGRAD) umma::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                for (int c = 0; c < p.cin; ++c) {
                    umma::mbar_wait(full_bar(stage), phase);
                    umma::fence_after_thread_sync();
                    const uint32_t sa = st_base + stage * stage_bytes;
                    if (!WGRAD) {
                        const uint32_t wb = smem_base + c * N * 128;
                        for (int mt = 0; mt < p.row_tiles; ++mt) {
                            const uint32_t d_tmem = tmem_base + (uint32_t)((acc * 4 + mt) * N);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma::mma_f16_ss(d_tmem, umma::make_smem_desc(sa + mt * 16384 + k * 32, 16, 1024),
                                                 umma::make_smem_desc(wb + k * 32, 16, 1024), idesc_f, (c | k) ? 1u : 0u);
                        }
                    } else {
                        const int ksteps = (p.rows + 15) / 16;
                        for (int k = 0; k < ksteps; ++k)
                            umma::mma_f16_ss(tmem_base + (uint32_t)(c * 64),
                                             umma::make_smem_desc(sa + k * 2048, 8192, 1024),
                                             umma::make_smem_desc(sa + p.a_stage_bytes + k * 2048, 8192, 1024), idesc_w,
                                             (!seen[c] && k == 0) ? 0u : 1u);
                        seen[c] = true;
                    }
                    umma::mma_commit(empty_bar(stage));
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                if (!WGRAD) {
                    umma::mma_commit(tfull_bar(acc));
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                }
            }
            if (WGRAD) umma::mma_commit(tfull_bar(0));
        }
    } else {
        // ------------------------------------------------------------ epilogue (4 warps, one per TMEM lane quarter)
        const int quarter = warp & 3;
        const int half = (warp - kMmaWarp - 1) >> 2;          // which of the two warps sharing this quarter
        if (!WGRAD) {
            int acc = 0; uint32_t acc_phase = 0;
            for (int s = blockIdx.x; s < p.n; s += gridDim.x) {
                umma::mbar_wait(tfull_bar(acc), acc_phase);
                umma::fence_after_thread_sync();
#pragma unroll 1
                for (int it = half; it < p.row_tiles * p.nets; it += kEpiWarps / 4) {
                    const int mt = it / p.nets, net = it - mt * p.nets;
                    uint32_t raw[32];
                    umma::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * 4 + mt) * N + net * 32),
                                        raw);
                    umma::tmem_ld_wait();
                    const int row = mt * 128 + quarter * 32 + lane;
                    if (row < p.rows && !(p.debug & 2)) {
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
                        const float* bsrc = net ? p.bias[1] : p.bias[0];
                        if (bsrc) {
                            const float4* b4 = reinterpret_cast<const float4*>(bsrc);
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
                        // this lane owns one 64-byte NHWC output row: two full-sector 32-byte stores
                        __nv_bfloat16* o = (net ? p.out[1] : p.out[0]) + ((long long)s * p.rows + row) * 32;
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            umma::U256 u;
#pragma unroll
                            for (int e = 0; e < 8; ++e) u.w[e] = pack2(v[q * 16 + e * 2], v[q * 16 + e * 2 + 1]);
                            umma::st_global_256(o + q * 16, u);
                        }
                    }
                }
                umma::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        } else if (blockIdx.x < p.n && half == 0) {
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
    if (warp == kMmaWarp) { umma::fence_after_thread_sync(); umma::tmem_dealloc(tmem_base, tmem_cols); }
}

// ------------------------------------------------------------------------------------------------
// Forward, space-to-depth form: the 8x8/stride-4 convolution is a 2x2/stride-1 convolution over 4x4
// pixel blocks.  The converter writes the normalised frame ONCE as [block row = by*(w/4)+bx][16 px]
// (32-byte rows, SWIZZLE_32B) and each of the four taps is a UMMA descriptor shifted by
// bh*(w/4)+bw rows into that image -- no im2col copy in shared memory at all.  Block positions in the
// last block row / column produce rows that are not stored (400 of 441 are kept for 84x84).
// ------------------------------------------------------------------------------------------------
constexpr int kS2dStageBytes = 18 * 1024;       // (4*128 + 22) rows x 32 B, rounded
constexpr int kS2dMaxStages = 6;

// Block rows may interleave CPK input channels (rows of 32*CPK bytes): SWIZZLE_32B for one channel,
// SWIZZLE_128B for four.  Address-based XOR swizzle of such an image and the matching descriptor constants.
template <int CPK> struct Pack {
    static_assert(CPK == 1 || CPK == 4, "CPK");
    static constexpr int RB = 32 * CPK;                         // bytes per block row
    static constexpr uint32_t MASK = CPK == 1 ? 1u : 7u;
    static constexpr uint32_t LAYOUT = CPK == 1 ? 6u : 2u;      // descriptor swizzle code
    static constexpr uint32_t SBO = 8u * RB;
    __device__ static __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & MASK) << 4); }
};
__device__ __forceinline__ uint64_t desc_any(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = umma::make_smem_desc(saddr, lbo, sbo);
    return (d & ~((uint64_t)7 << 61)) | ((uint64_t)layout << 61);
}

// CG converter groups of four warps (alternate units) and kEW epilogue warps (a multiple of 4).
template <int CPK, int CG, int kEW>
__global__ void __launch_bounds__((4 * CG + kEW + 2) * 32, 1) umma_conv1s_fwd_kernel(const Conv1Params p) {
    using PK = Pack<CPK>;
    constexpr int RB = PK::RB;
    constexpr int kCW = 4 * CG, kMma = kCW, kEpi0 = kCW + 1, kLoad = kCW + 1 + kEW, kEG = kEW / 4;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (smem_base - umma::smem_u32(smem_raw));
    const int N = 32 * p.nets;
    const int bwc = p.w / 4;                                   // blocks per frame row
    const int brows = (p.h / 4) * bwc;                         // block positions per sample
    const uint32_t w_bytes = (uint32_t)(p.cin * 4 * N * 32);
    const uint32_t t_off = w_bytes;
    const uint32_t t_bytes = p.mean ? (uint32_t)((p.h * p.w * 8 + 1023) & ~1023) : 0u;
    const uint32_t st_off = t_off + t_bytes;
    const int ups = p.cin / CPK;                               // pipeline units (stages) per sample
    const uint32_t stage_bytes = (uint32_t)p.s2d_stage_bytes;
    const uint32_t raw_off = st_off + p.stages * stage_bytes;                 // ring of raw uint8 frames
    const uint32_t raw_bytes = (uint32_t)((p.h * p.w + 127) & ~127);
    // The bias is added by the tensor cores: one more K = 16 MMA per row tile whose A block has ones in its first
    // three elements of every row and whose B block holds bias[n] split into three bf16 terms (hi + mid + lo, exact
    // to fp32 rounding) -- instead of 32 FADDs and 8 shared-memory loads per 32x32 block in the epilogue.
    const uint32_t ones_off = (raw_off + p.raw_slots * raw_bytes + 1023u) & ~1023u;     // [128 rows][RB]
    const uint32_t bblk_off = ones_off + 128u * RB;                                     // [N rows][RB]
    const uint32_t bar_base = smem_base + bblk_off + 64u * RB;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kS2dMaxStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kS2dMaxStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kS2dMaxStages + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kS2dMaxStages + 4);
    auto rfull_bar = [&](int s) { return bar_base + 8u * (2 * kS2dMaxStages + 5 + s); };
    auto rempty_bar = [&](int s) { return bar_base + 8u * (2 * kS2dMaxStages + 5 + kMaxRaw + s); };
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - smem_base));
    const uint32_t tmem_cols = N == 64 ? 512u : 256u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    {   // zero the stages once (block rows past the frame stay zero)
        uint4* z = reinterpret_cast<uint4*>(gen + st_off);
        for (uint32_t i = threadIdx.x; i < (uint32_t)p.stages * stage_bytes / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    }
    // weights: per (unit, tap) a K-major [N rows][CPK channels x 16 = (py,px)] block in the image's swizzle
    for (int i = threadIdx.x; i < p.cin * 4 * N * 16; i += blockDim.x) {
        const int e = i & 15, n = (i >> 4) % N, t = ((i >> 4) / N) & 3, c = (i >> 4) / (4 * N);
        const int kh = 4 * (t >> 1) + (e >> 2), kw = 4 * (t & 1) + (e & 3);
        const __nv_bfloat16 v = p.w_bf16[n >> 5][(long long)(n & 31) * p.cin * 64 + c * 64 + kh * 8 + kw];
        const int u = c / CPK, cc = c - u * CPK;
        *reinterpret_cast<__nv_bfloat16*>(gen + PK::swz((uint32_t)(((u * 4 + t) * N + n) * RB + cc * 32 + e * 2))) = v;
    }
    if (p.mean) {
        float4* tm = reinterpret_cast<float4*>(gen + t_off);
        const int g4 = p.h * p.w / 4;
        for (int i = threadIdx.x; i < g4; i += blockDim.x) {
            tm[i] = __ldg(reinterpret_cast<const float4*>(p.mean) + i);
            tm[g4 + i] = __ldg(reinterpret_cast<const float4*>(p.rstd) + i);
        }
    }
    for (uint32_t i = threadIdx.x; i < (128u + 64u) * RB / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(gen + ones_off)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (threadIdx.x < 128) {            // ones: elements 0..2 of every row
        const uint32_t o = PK::swz((uint32_t)(threadIdx.x * RB));
        *reinterpret_cast<uint2*>(gen + ones_off + o) = make_uint2(0x3f803f80u, 0x00003f80u);
    }
    if ((int)threadIdx.x < N && p.bias[threadIdx.x >> 5]) {
        const float b = __ldg(p.bias[threadIdx.x >> 5] + (threadIdx.x & 31));
        const __nv_bfloat16 hi = __float2bfloat16_rn(b);
        const float r1 = b - __bfloat162float(hi);
        const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
        const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
        const uint32_t o = PK::swz((uint32_t)(threadIdx.x * RB));
        *reinterpret_cast<uint2*>(gen + bblk_off + o) =
            make_uint2((uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16),
                       (uint32_t)__bfloat16_as_ushort(lo));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { umma::mbar_init(full_bar(s), kGrpThreads / 32); umma::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { umma::mbar_init(tfull_bar(a), 1); umma::mbar_init(tempty_bar(a), kEW); }
        for (int r = 0; r < p.raw_slots; ++r) { umma::mbar_init(rfull_bar(r), 1); umma::mbar_init(rempty_bar(r), kGrpThreads / 32); }
        umma::fence_barrier_init();
    }
    if (warp == kMma) umma::tmem_alloc(tmem_slot, tmem_cols);
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int row_tiles = (brows + 127) / 128;

    if (warp < kCW) {
        // ------------------------------------------------------------ converters: uint8 -> s2d bf16 image
        // Two groups of four warps take alternate units (one channel of one sample each), so that one group's
        // wait -> read -> convert -> fence -> arrive chain overlaps the other's (a single group of eight warps
        // marching in lock-step was latency-bound at ~1200 cycles per unit).
        const int grp = warp >> 2;
        const int tid = threadIdx.x - grp * kGrpThreads;
        const float4* s_mean = reinterpret_cast<const float4*>(gen + t_off);
        const float4* s_rstd = s_mean + p.h * p.w / 4;
        const int groups = p.h * p.w / 4;
        // Thread -> pixel-group map, fixed for the whole kernel: item i of a thread is the 4-pixel group
        // g = tid + i*threads of the raw frame; soff = where it lands in the s2d image (computed once).
        uint32_t soff[kGrpMaxGroups];
#pragma unroll
        for (int i = 0; i < kGrpMaxGroups; ++i) {
            const int g = tid + i * kGrpThreads, y = g / bwc, x4 = g - y * bwc;
            soff[i] = (uint32_t)(((y >> 2) * bwc + x4) * RB + (y & 3) * 8);          // + 32*channel, then swizzled
        }
        const bool norm = p.mean && !(p.debug & 1);
        // unit k (in this CTA's order: CPK channels of one sample) uses pipeline stage k % stages; its frames
        // are k*CPK .. k*CPK+CPK-1 of the raw ring
        const int my_samples = (p.n - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int my_units = my_samples * ups;
        for (int k = grp; k < my_units; k += CG) {
            const int stage = k % p.stages;
            const uint32_t phase = (uint32_t)(k / p.stages) & 1u;
            uint8_t* A = gen + st_off + stage * stage_bytes;
#pragma unroll 1
            for (int cc = 0; cc < CPK; ++cc) {
                const int fr = k * CPK + cc, rslot = fr % p.raw_slots;
                const uint32_t rphase = (uint32_t)(fr / p.raw_slots) & 1u;
                // this frame's raw pixels: staged by the loader warp several frames ahead (HBM latency hidden)
                umma::mbar_wait(rfull_bar(rslot), rphase);
                // all of this thread's pixels first (14 independent shared-memory loads in flight), so the raw
                // slot is released early and the conversion below never waits on a load
                uint32_t px[kGrpMaxGroups];
                {
                    const uint32_t* raw = reinterpret_cast<const uint32_t*>(gen + raw_off + rslot * raw_bytes);
#pragma unroll
                    for (int i = 0; i < kGrpMaxGroups; ++i) {
                        const int g = tid + i * kGrpThreads;
                        px[i] = g < groups ? raw[g] : 0u;
                    }
                }
                umma::loads_landed(bar_base + 320u + 4u * warp, px);   // the slot is refilled as soon as it is released
                __syncwarp();                                 // one arrival per warp
                if (lane == 0) umma::mbar_arrive(rempty_bar(rslot));
                if (cc == 0) umma::mbar_wait(empty_bar(stage), phase ^ 1u);
#pragma unroll
                for (int i = 0; i < kGrpMaxGroups; ++i) {
                    const int g = tid + i * kGrpThreads;
                    if (g < groups) {
                        float v[4];
                        u8x4_to_f32(px[i], v);
                        if (norm) {
                            const float4 m = s_mean[g], r = s_rstd[g];
                            v[0] = fminf(fmaxf((v[0] - m.x) * r.x, -5.f), 5.f);
                            v[1] = fminf(fmaxf((v[1] - m.y) * r.y, -5.f), 5.f);
                            v[2] = fminf(fmaxf((v[2] - m.z) * r.z, -5.f), 5.f);
                            v[3] = fminf(fmaxf((v[3] - m.w) * r.w, -5.f), 5.f);
                        }
                        *reinterpret_cast<uint2*>(A + PK::swz(soff[i] + cc * 32)) = make_uint2(pack2(v[0], v[1]), pack2(v[2], v[3]));
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(full_bar(stage));
        }
    } else if (warp == kLoad) {
        // ------------------------------------------------------------ raw-frame loader
        if (lane == 0) {
            int rslot = 0; uint32_t rphase = 0;
            for (int s = blockIdx.x; s < p.n; s += gridDim.x)
                for (int c = 0; c < p.cin; ++c) {
                    umma::mbar_wait(rempty_bar(rslot), rphase ^ 1u);
                    umma::mbar_arrive_expect_tx(rfull_bar(rslot), (uint32_t)(p.h * p.w));
                    bulk_load(smem_base + raw_off + rslot * raw_bytes, p.x + s * p.stride_n + c * p.stride_c,
                              (uint32_t)(p.h * p.w), rfull_bar(rslot));
                    if (++rslot == p.raw_slots) { rslot = 0; rphase ^= 1u; }
                }
        }
    } else if (warp == kMma) {
        // All 32 lanes run the loops (warp-uniform state); one elected lane issues the tcgen05 instructions.
        const bool leader = umma::elect_one();
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        const uint32_t idesc = N == 64 ? umma::make_idesc(128, 64, 0, 0) : umma::make_idesc(128, 32, 0, 0);
        for (int s = blockIdx.x; s < p.n; s += gridDim.x) {
            umma::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            for (int u = 0; u < ups; ++u) {
                umma::mbar_wait(full_bar(stage), phase);
                umma::fence_after_thread_sync();
                const uint64_t da0 = desc_any(smem_base + st_off + stage * stage_bytes, 16, PK::SBO, PK::LAYOUT);
                const uint64_t db0 = desc_any(smem_base + u * 4 * N * RB, 16, PK::SBO, PK::LAYOUT);
                const uint32_t d0 = tmem_base + (uint32_t)(acc * 4 * N);
                const uint64_t d_ones = desc_any(smem_base + ones_off, 16, PK::SBO, PK::LAYOUT);
                const uint64_t d_bias = desc_any(smem_base + bblk_off, 16, PK::SBO, PK::LAYOUT);
                if (leader) {
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        if (mt < row_tiles && !((p.debug & 16) && mt)) {
                            if (u == 0) umma::mma_f16_ss(d0 + (uint32_t)(mt * N), d_ones, d_bias, idesc, 0u);   // D = bias
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const int shift = (t >> 1) * bwc + (t & 1);
#pragma unroll
                                for (int kc = 0; kc < CPK; ++kc)     // one 16-pixel k-step per interleaved channel
                                    umma::mma_f16_ss(d0 + (uint32_t)(mt * N),
                                                     da0 + (uint64_t)((((mt * 128 + shift) * RB) >> 4) + kc * 2),
                                                     db0 + (uint64_t)(((t * N * RB) >> 4) + kc * 2), idesc, 1u);
                            }
                        }
                    }
                    umma::mma_commit(empty_bar(stage));
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            if (leader) umma::mma_commit(tfull_bar(acc));
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else {
        // ------------------------------------------------------------ epilogue
        // TMEM reads run at 64 B/clk/SM (a 32x32 fp32 block takes 64 cycles), so each warp keeps the NEXT
        // block's tcgen05.ld in flight while it applies bias/activation to the current one and stores it.
        const int quarter = warp & 3;
        const int sub = (warp - kEpi0) >> 2;                   // which of the kEG warps sharing this TMEM lane quarter
        // The block-grid row of this lane in each row tile (and whether it is stored) never changes: hoisted
        // out of the sample loop so the per-block work is tcgen05.ld + bias/activation + pack + two stores.
        int out_row[4];                                  // by*ow + bx, or -1
        bool tile_live[4];                               // any of the warp's 32 rows inside the block grid
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int r = mt * 128 + quarter * 32 + lane;
            const int by = r / bwc, bx = r - by * bwc;
            out_row[mt] = (mt < row_tiles && by < p.oh && bx < p.ow) ? by * p.ow + bx : -1;
            tile_live[mt] = mt < row_tiles && mt * 128 + quarter * 32 < brows;
        }
        const int act = p.act;
        int acc = 0; uint32_t acc_phase = 0;
        for (int s = blockIdx.x; s < p.n; s += gridDim.x) {
            umma::mbar_wait(tfull_bar(acc), acc_phase);
            umma::fence_after_thread_sync();
            const long long obase = (long long)s * p.rows;
            // (row tile, net) blocks alternate between the two warps of a quarter; mt is unrolled so that
            // out_row / tile_live stay in registers (a dynamically indexed copy lands in local memory: +30 % time).
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                if (!tile_live[mt]) continue;
#pragma unroll 1
                for (int net = 0; net < p.nets; ++net) {
                    if ((mt * p.nets + net) % kEG != sub) continue;
                    uint32_t raw[32];
                    umma::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * 4 + mt) * N + net * 32), raw);
                    umma::tmem_ld_wait();
                    if (out_row[mt] < 0 || (p.debug & 2)) continue;
                    // accumulator = conv + bias already; LeakyReLU / ReLU on packed bf16 pairs (half the instructions of
                    // the fp32 form; the negative branch is rounded twice, 2^-9 relative on values scaled by 0.01)
                    uint32_t pk[16];
                    if (act == CB_ACT_ELU) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            float a = __uint_as_float(raw[2 * e]), b = __uint_as_float(raw[2 * e + 1]);
                            a = a > 0.f ? a : expm1f(a); b = b > 0.f ? b : expm1f(b);
                            pk[e] = pack2(a, b);
                        }
                    } else {
                        const __nv_bfloat162 slope = __floats2bfloat162_rn(0.01f, 0.01f), zero = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(raw[2 * e]), __uint_as_float(raw[2 * e + 1]));
                            if (act == CB_ACT_LRELU) h = __hmax2(h, __hmul2(h, slope));
                            else if (act == CB_ACT_RELU) h = __hmax2(h, zero);
                            pk[e] = *reinterpret_cast<uint32_t*>(&h);
                        }
                    }
                    // this lane owns one 64-byte NHWC output row: two full-sector 32-byte stores
                    __nv_bfloat16* o = (net ? p.out[1] : p.out[0]) + (obase + out_row[mt]) * 32;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        umma::U256 u;
#pragma unroll
                        for (int e = 0; e < 8; ++e) u.w[e] = pk[q * 8 + e];
                        umma::st_global_256(o + q * 16, u);
                    }
                }
            }
            umma::fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    if (warp == kMma) { umma::fence_after_thread_sync(); umma::tmem_dealloc(tmem_base, tmem_cols); }
}

template <int CPK, int CG, int EW>
int launch_s2d_fwd_t(Conv1Params& p, int smem, cudaStream_t s) {
    static bool attr = false;
    if (!attr) {
        CB_CUDA(cudaFuncSetAttribute(umma_conv1s_fwd_kernel<CPK, CG, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = p.n < sms ? p.n : sms;
    umma_conv1s_fwd_kernel<CPK, CG, EW><<<grid, (4 * CG + EW + 2) * 32, smem, s>>>(p);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int launch_s2d_fwd(Conv1Params& p, cudaStream_t s) {
    const int N = 32 * p.nets;
    const int raw_bytes = (p.h * p.w + 127) & ~127;
    const int bwc = p.w / 4, row_tiles = ((p.h / 4) * bwc + 127) / 128;
    const int fixed = p.cin * 4 * N * 32 + (p.mean ? ((p.h * p.w * 8 + 1023) & ~1023) : 0) + 1024 + 1024 + 192 * 128 + 512;
    // four channels interleaved per block row when they fit (128-byte rows feed the MMA far better than 32-byte ones)
    int cpk = p.cin % 4 == 0 ? 4 : 1;
    for (;; cpk = 1) {
        p.s2d_stage_bytes = ((row_tiles * 128 + bwc + 2) * 32 * cpk + 1023) & ~1023;
        p.stages = (224 * 1024 - fixed - (cpk + 2) * raw_bytes) / p.s2d_stage_bytes;
        if (p.stages >= 2 || cpk == 1) break;
    }
    if (p.stages > 4) p.stages = 4;
    CB_REQUIRE(p.stages >= 2, CB_E_ARG, "conv1: does not fit shared memory");
    p.raw_slots = (224 * 1024 - fixed - p.stages * p.s2d_stage_bytes) / raw_bytes;
    if (p.raw_slots > kMaxRaw) p.raw_slots = kMaxRaw;
    const int smem = fixed + p.stages * p.s2d_stage_bytes + p.raw_slots * raw_bytes;
    if (cpk == 4) return launch_s2d_fwd_t<4, 2, 8>(p, smem, s);
    // (measured on the twin pass: 4 converter + 12 epilogue warps 2.17 ms, 8 + 12 at 80 registers 2.06 ms, 8 + 8 1.95 ms)
    return launch_s2d_fwd_t<1, 2, 8>(p, smem, s);
}

// ------------------------------------------------------------------------------------------------
// Weight gradient, space-to-depth form.  With block rows r = by*(w/4)+bx and tap shifts {0, 1, G, G+1}
// (G = w/4):   dW[n][tap, e] = sum_r dy_grid[r][n] * image[r + shift_tap][e].
// ONE tcgen05.mma per 16 block rows yields all four taps from a single copy of each operand, because
// the two 32-wide halves of M and the two 16-wide halves of N may be overlapping windows of the same
// shared-memory image (descriptor LBO = window distance; pinned by tools/probe_wgrad1.py):
//   A = dyP, the sample's dy rows on the block grid with G zero rows in front, written by ONE TMA tensor
//       load (out-of-range grid positions are zero-filled): MN-major SWIZZLE_64B, M = 64, LBO = G rows;
//   B = the s2d frame written once by the converters: MN-major SWIZZLE_32B, N = 32, LBO = 1 row;
//   D[(j,n)][(i,e)] = sum_q dyP[q + G j][n] * img[q + i][e]   ->   tap (bh, bw) = (1 - j, i).
// Per input channel one [64 x 32] TMEM accumulator over all samples of the CTA.  Raw uint8 frames and dy
// arrive through asynchronous copies several units ahead, so HBM latency is never exposed.
// ------------------------------------------------------------------------------------------------
constexpr int kWgConvWarps = 8;
constexpr int kWgConvThreads = kWgConvWarps * 32;
constexpr int kWgMmaWarp = kWgConvWarps;            // + 4 epilogue warps, raw-frame loader, dy loader
constexpr int kWgRawWarp = kWgConvWarps + 5;
constexpr int kWgDyWarp = kWgConvWarps + 6;
constexpr int kWgThreads = (kWgConvWarps + 7) * 32; // 15 warps: 128 registers per thread
constexpr int kWgMaxStages = 4;

struct WgGeo { int dy_rows, dy_bytes, img_bytes, stage_bytes, ksteps; };

__host__ __device__ inline WgGeo wg_geo(int h, int w, int cpk = 1) {        // cpk: channels interleaved per block row
    WgGeo g;
    const int G = w / 4, brows = (h / 4) * G;
    g.ksteps = (brows + G + 15) / 16;
    g.dy_rows = G * (h / 4 + 3);
    g.dy_bytes = (g.dy_rows * 64 + 1023) & ~1023;
    g.img_bytes = ((g.ksteps * 16 + 1) * 32 * cpk + 1023) & ~1023;
    g.stage_bytes = g.dy_bytes + g.img_bytes;
    return g;
}

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = umma::make_smem_desc(saddr, lbo, sbo);
    return (d & ~((uint64_t)7 << 61)) | ((uint64_t)layout << 61);
}

// CPK = 4: the four channels of a stack interleaved in 128-byte block rows (SWIZZLE_128B); one N = 128 MMA per 16
// block rows then covers all channels and both window columns, and a pipeline unit is a whole sample.
template <int CPK>
__global__ void __launch_bounds__(kWgThreads, 1)
umma_conv1s_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const Conv1Params p) {
    using PK = Pack<CPK>;
    constexpr int RB = PK::RB;
    const int ups = p.cin / CPK;                               // pipeline units per sample
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (smem_base - umma::smem_u32(smem_raw));
    const int bwc = p.w / 4;
    const WgGeo geo = wg_geo(p.h, p.w, CPK);
    // layout: [stages: dyP | image] [mean|rstd table] [raw ring] [barriers]
    const uint32_t t_off = (uint32_t)(p.stages * geo.stage_bytes);
    const uint32_t t_bytes = p.mean ? (uint32_t)((p.h * p.w * 8 + 1023) & ~1023) : 0u;
    const uint32_t raw_off = t_off + t_bytes;
    const uint32_t raw_bytes = (uint32_t)((p.h * p.w + 127) & ~127);
    const uint32_t ones_off = (raw_off + p.raw_slots * raw_bytes + 1023u) & ~1023u;   // 16 K rows x 16 bf16 ones (bias gradient)
    const uint32_t bar_base = smem_base + ones_off + 512u;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kWgMaxStages + s); };
    auto rfull_bar = [&](int s) { return bar_base + 8u * (2 * kWgMaxStages + s); };
    auto rempty_bar = [&](int s) { return bar_base + 8u * (2 * kWgMaxStages + kMaxRaw + s); };
    const uint32_t done_bar = bar_base + 8u * (2 * kWgMaxStages + 2 * kMaxRaw);
    const uint32_t tmem_slot = done_bar + 8u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - smem_base));
    const uint32_t tmem_need = (uint32_t)(p.cin * 32 + 16);         // + the bias-gradient accumulator [64 x 16]
    const uint32_t tmem_cols = tmem_need <= 32 ? 32u : tmem_need <= 64 ? 64u : tmem_need <= 128 ? 128u : 256u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 128) reinterpret_cast<uint32_t*>(gen + ones_off)[threadIdx.x] = 0x3f803f80u;   // bf16 1.0 pairs
    {   // zero the stages once: image rows past the frame are read by the last k-steps and must stay zero
        uint4* z = reinterpret_cast<uint4*>(gen);
        for (uint32_t i = threadIdx.x; i < t_off / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    }
    if (p.mean) {
        float4* tm = reinterpret_cast<float4*>(gen + t_off);
        const int g4 = p.h * p.w / 4;
        for (int i = threadIdx.x; i < g4; i += blockDim.x) {
            tm[i] = __ldg(reinterpret_cast<const float4*>(p.mean) + i);
            tm[g4 + i] = __ldg(reinterpret_cast<const float4*>(p.rstd) + i);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        umma::prefetch_tmap(&tmDy);
        for (int s = 0; s < p.stages; ++s) { umma::mbar_init(full_bar(s), kWgConvWarps + 1); umma::mbar_init(empty_bar(s), 1); }
        for (int r = 0; r < p.raw_slots; ++r) { umma::mbar_init(rfull_bar(r), 1); umma::mbar_init(rempty_bar(r), kWgConvWarps); }
        umma::mbar_init(done_bar, 1);
        umma::fence_barrier_init();
    }
    if (warp == kWgMmaWarp) umma::tmem_alloc(tmem_slot, tmem_cols);
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < kWgConvWarps) {
        // ------------------------------------------------------------ converters: uint8 -> s2d bf16 image
        const int tid = threadIdx.x;
        const float4* s_mean = reinterpret_cast<const float4*>(gen + t_off);
        const float4* s_rstd = s_mean + p.h * p.w / 4;
        const int groups = p.h * p.w / 4;
        int stage = 0; uint32_t phase = 0;
        int rslot = 0; uint32_t rphase = 0;
        // thread -> pixel-group map, computed once (see the forward kernel)
        uint32_t soff[kMaxGroups]; int gidx[kMaxGroups];
#pragma unroll
        for (int i = 0; i < kMaxGroups; ++i) {
            const int g = tid + i * kWgConvThreads, y = g / bwc, x4 = g - y * bwc;
            gidx[i] = g < groups ? g : -1;
            soff[i] = (uint32_t)(((y >> 2) * bwc + x4) * RB + (y & 3) * 8);       // + 32*channel, then swizzled
        }
        for (int s = blockIdx.x; s < p.n; s += gridDim.x)
            for (int u = 0; u < ups; ++u) {
                uint8_t* A = gen + stage * geo.stage_bytes + geo.dy_bytes;
#pragma unroll 1
                for (int cc = 0; cc < CPK; ++cc) {
                    umma::mbar_wait(rfull_bar(rslot), rphase);
                    uint32_t px[kMaxGroups];                   // all loads first: the slot is released early
                    {
                        const uint32_t* raw = reinterpret_cast<const uint32_t*>(gen + raw_off + rslot * raw_bytes);
#pragma unroll
                        for (int i = 0; i < kMaxGroups; ++i) px[i] = gidx[i] >= 0 ? raw[gidx[i]] : 0u;
                    }
                    umma::loads_landed(bar_base + 256u + 4u * warp, px);   // the slot is refilled as soon as it is released
                    __syncwarp();
                    if (lane == 0) umma::mbar_arrive(rempty_bar(rslot));       // one arrival per warp
                    if (++rslot == p.raw_slots) { rslot = 0; rphase ^= 1u; }
                    if (cc == 0) umma::mbar_wait(empty_bar(stage), phase ^ 1u);
#pragma unroll
                    for (int i = 0; i < kMaxGroups; ++i) {
                        const int g = gidx[i];
                        if (g >= 0) {
                            float v[4];
                            u8x4_to_f32(px[i], v);
                            if (p.mean) {
                                const float4 m = s_mean[g], r = s_rstd[g];
                                v[0] = fminf(fmaxf((v[0] - m.x) * r.x, -5.f), 5.f);
                                v[1] = fminf(fmaxf((v[1] - m.y) * r.y, -5.f), 5.f);
                                v[2] = fminf(fmaxf((v[2] - m.z) * r.z, -5.f), 5.f);
                                v[3] = fminf(fmaxf((v[3] - m.w) * r.w, -5.f), 5.f);
                            }
                            *reinterpret_cast<uint2*>(A + PK::swz(soff[i] + cc * 32)) = make_uint2(pack2(v[0], v[1]), pack2(v[2], v[3]));
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(full_bar(stage));
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
    } else if (warp == kWgRawWarp) {
        // ------------------------------------------------------------ raw-frame loader (cp.async.bulk ring)
        if (lane == 0) {
            int rslot = 0; uint32_t rphase = 0;
            for (int s = blockIdx.x; s < p.n; s += gridDim.x)
                for (int c = 0; c < p.cin; ++c) {
                    umma::mbar_wait(rempty_bar(rslot), rphase ^ 1u);
                    umma::mbar_arrive_expect_tx(rfull_bar(rslot), (uint32_t)(p.h * p.w));
                    bulk_load(smem_base + raw_off + rslot * raw_bytes, p.x + s * p.stride_n + c * p.stride_c,
                              (uint32_t)(p.h * p.w), rfull_bar(rslot));
                    if (++rslot == p.raw_slots) { rslot = 0; rphase ^= 1u; }
                }
        }
    } else if (warp == kWgDyWarp) {
        // ------------------------------------------------------------ dy loader: one tensor box per unit
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int s = blockIdx.x; s < p.n; s += gridDim.x)
                for (int u = 0; u < ups; ++u) {
                    umma::mbar_wait(empty_bar(stage), phase ^ 1u);
                    umma::mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(geo.dy_rows * 64));
                    umma::tma_load_4d(smem_base + stage * geo.stage_bytes, &tmDy, full_bar(stage), 0, 0, -1, s);
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
        }
    } else if (warp == kWgMmaWarp) {
        // all 32 lanes run the loops (warp-uniform state, descriptors in uniform registers); one elected lane issues
        const bool leader = umma::elect_one();
        int stage = 0; uint32_t phase = 0;
        uint32_t seen = 0;
        constexpr uint32_t idesc = umma::make_idesc(64, 32 * CPK, 1, 1);
        constexpr uint32_t idesc_b = umma::make_idesc(64, 16, 1, 1);
        for (int s = blockIdx.x; s < p.n; s += gridDim.x) {
            for (int u = 0; u < ups; ++u) {
                umma::mbar_wait(full_bar(stage), phase);
                umma::fence_after_thread_sync();
                const uint32_t st = smem_base + stage * geo.stage_bytes;
                const uint64_t da0 = desc_mn(st, (uint32_t)(bwc * 64), 512, 4);              // SWIZZLE_64B, windows G rows apart
                const uint64_t db0 = desc_mn(st + geo.dy_bytes, RB, PK::SBO, PK::LAYOUT);     // windows 1 row apart
                const uint64_t d1 = desc_mn(smem_base + ones_off, 32, 256, 6);
                const uint32_t first = (seen >> u) & 1u;
                if (leader) {
#pragma unroll 4
                    for (int k = 0; k < geo.ksteps; ++k)
                        umma::mma_f16_ss(tmem_base + (uint32_t)(u * 32 * CPK), da0 + (uint64_t)(k * 64), db0 + (uint64_t)(k * RB), idesc,
                                         (first | (uint32_t)k) ? 1u : 0u);
                    if (u == 0 && p.dbias) {     // column sums of dy (bias gradient): the same A windows times a block of ones
#pragma unroll 4
                        for (int k = 0; k < geo.ksteps; ++k)
                            umma::mma_f16_ss(tmem_base + (uint32_t)(p.cin * 32), da0 + (uint64_t)(k * 64), d1, idesc_b,
                                             (first | (uint32_t)k) ? 1u : 0u);
                    }
                    umma::mma_commit(empty_bar(stage));
                }
                __syncwarp();
                seen |= 1u << u;
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
        if (leader) umma::mma_commit(done_bar);
        __syncwarp();
    } else if (warp < kWgRawWarp && blockIdx.x < p.n) {
        // ------------------------------------------------------------ epilogue: TMEM -> atomics into dW
        // M = 64 accumulator: row i = (j, n) lives in TMEM lane (i/16)*32 + i%16
        const int quarter = warp & 3;
        umma::mbar_wait(done_bar, 0);
        umma::fence_after_thread_sync();
        const int K = p.cin * 64;
        for (int cb = 0; cb < p.cin; ++cb) {               // 32 accumulator columns at a time: col = u*32*CPK + (i, cc, e)
            uint32_t raw[32];
            umma::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cb * 32), raw);
            umma::tmem_ld_wait();
            if (lane < 16) {
                const int i = quarter * 16 + lane, j = i >> 5, n = i & 31, bh = 1 - j;
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const int col = cb * 32 + q, u = col / (32 * CPK), rem = col - u * 32 * CPK;
                    const int bw = rem / (16 * CPK), cc = (rem / 16) % CPK, e = rem & 15;
                    const int c = u * CPK + cc, kh = 4 * bh + (e >> 2), kw = 4 * bw + (e & 3);
                    atomicAdd(p.dw + (long long)n * K + c * 64 + kh * 8 + kw, __uint_as_float(raw[q]));
                }
            }
        }
        if (p.dbias) {      // rows (j = 1, n): the window that covers every real dy row
            uint32_t raw[16];
            umma::tmem_ld_32x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(p.cin * 32), raw);
            umma::tmem_ld_wait();
            if (lane < 16 && quarter >= 2) atomicAdd(p.dbias + (quarter - 2) * 16 + lane, __uint_as_float(raw[0]));
        }
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    if (warp == kWgMmaWarp) { umma::fence_after_thread_sync(); umma::tmem_dealloc(tmem_base, tmem_cols); }
}

bool s2d_wgrad_ok(const cb_conv_desc* d) {
    const int G = d->in_w / 4;
    const WgGeo g = wg_geo(d->in_h, d->in_w);
    return G >= 15 && G <= 256 && d->in_h / 4 + 3 <= 256 && g.ksteps * 16 + G <= g.dy_rows &&
           d->out_h == d->in_h / 4 - 1 && d->out_w == G - 1;
}

template <int CPK>
int launch_s2d_wgrad_t(const CUtensorMap& tmDy, Conv1Params& p, int smem, cudaStream_t s) {
    static bool attr = false;
    if (!attr) {
        CB_CUDA(cudaFuncSetAttribute(umma_conv1s_wgrad_kernel<CPK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = p.n < sms ? p.n : sms;
    umma_conv1s_wgrad_kernel<CPK><<<grid, kWgThreads, smem, s>>>(tmDy, p);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int launch_s2d_wgrad(Conv1Params& p, cudaStream_t s) {
    const int raw_bytes = (p.h * p.w + 127) & ~127;
    const int fixed = (p.mean ? ((p.h * p.w * 8 + 1023) & ~1023) : 0) + 1024 + 1024 + 512 + 512;
    int cpk = p.cin % 4 == 0 ? 4 : 1;
    WgGeo g;
    for (;; cpk = 1) {         // four channels per block row when two stages of them fit
        g = wg_geo(p.h, p.w, cpk);
        p.stages = (224 * 1024 - fixed - (cpk + 1) * raw_bytes) / g.stage_bytes;
        if (p.stages >= 2 || cpk == 1) break;
    }
    if (p.stages > kWgMaxStages) p.stages = kWgMaxStages;
    CB_REQUIRE(p.stages >= 2, CB_E_ARG, "conv1 wgrad: does not fit shared memory");
    p.raw_slots = (224 * 1024 - fixed - p.stages * g.stage_bytes) / raw_bytes;
    if (p.raw_slots > kMaxRaw) p.raw_slots = kMaxRaw;
    const int smem = fixed + p.stages * g.stage_bytes + p.raw_slots * raw_bytes;
    // dy [n][oh][ow][32] bf16; one box = a sample's rows on the (w/4)-wide block grid with a zero row block in front
    CUtensorMap tmDy;
    {
        uint64_t dims[4] = {32, (uint64_t)p.ow, (uint64_t)p.oh, (uint64_t)p.n};
        uint64_t strides[3] = {64, (uint64_t)p.ow * 64, (uint64_t)p.oh * p.ow * 64};
        uint32_t box[4] = {32, (uint32_t)(p.w / 4), (uint32_t)(p.h / 4 + 3), 1};
        int rc = cb::encode_tmap_bf16(&tmDy, p.dy, 4, dims, strides, box, 2); if (rc) return rc;
    }
    return cpk == 4 ? launch_s2d_wgrad_t<4>(tmDy, p, smem, s) : launch_s2d_wgrad_t<1>(tmDy, p, smem, s);
}

bool s2d_ok(const cb_conv_desc* d) {      // block grid must fit four M tiles plus the largest tap shift
    const int bwc = d->in_w / 4, brows = (d->in_h / 4) * bwc;
    return d->in_h % 4 == 0 && brows <= 512 && (brows + 127) / 128 * 128 + bwc + 1 <= kS2dStageBytes / 32 &&
           (d->in_h * d->in_w / 4 + kGrpThreads - 1) / kGrpThreads <= kGrpMaxGroups &&
           (d->in_h * d->in_w) % 16 == 0 && d->in_stride_n % 16 == 0 && d->in_stride_c % 16 == 0;   // cp.async.bulk
}

bool shape_ok(const cb_conv_desc* d) {
    return d->in_dtype == CB_U8 && d->k_h == 8 && d->k_w == 8 && d->stride == 4 && d->pad == 0 && d->out_c == 32 &&
           d->in_c >= 1 && d->in_c <= kMaxCin && d->in_stride_w == 1 && d->in_stride_h == d->in_w && d->in_w % 4 == 0 &&
           d->in_stride_n % 4 == 0 && d->in_stride_c % 4 == 0 && d->out_h * d->out_w <= 512 &&
           (d->in_h * d->in_w / 4 + kConvThreads - 1) / kConvThreads <= kMaxGroups && d->out_w <= 32;
}

int fill(Conv1Params& p, const cb_conv_desc* d, const void* x, const float* mean, const float* rstd) {
    CB_REQUIRE(((uintptr_t)x & 3) == 0, CB_E_ALIGN, "conv1: frames must be 4-byte aligned");
    CB_REQUIRE(!mean || ((((uintptr_t)mean | (uintptr_t)rstd) & 15) == 0), CB_E_ALIGN, "conv1: mean/rstd 16-byte aligned");
    p.x = (const uint8_t*)x; p.stride_n = d->in_stride_n; p.stride_c = d->in_stride_c;
    p.h = d->in_h; p.w = d->in_w; p.cin = d->in_c; p.oh = d->out_h; p.ow = d->out_w; p.n = d->n;
    p.rows = d->out_h * d->out_w; p.row_tiles = (p.rows + 127) / 128;
    p.pitch = (d->in_w * 2 + 15) / 16 * 16;
    if (p.pitch % 128 == 0) p.pitch += 16;
    p.units = d->n * d->in_c;
    p.mean = mean; p.rstd = rstd;
    p.debug = 0;
    return CB_OK;
}

template <bool WGRAD>
int launch(Conv1Params& p, cudaStream_t s) {
    // forward: rows padded to whole M=128 tiles; wgrad: K = rows rounded to 16, patches and dy rows per stage
    p.a_stage_bytes = ((p.rows + 15) / 16 * 16 * 128 + 1023) / 1024 * 1024;   // the last M tile may read past it (finite bytes)
    const int stage_bytes = p.a_stage_bytes * (WGRAD ? 2 : 1);
    const int fixed = (WGRAD ? 0 : p.cin * 32 * p.nets * 128) + ((p.h * p.pitch + 1023) & ~1023) +
                      (p.mean ? ((p.h * p.w * 8 + 1023) & ~1023) : 0) +
                      (WGRAD ? 0 : ((kEpiWarps * kStagePitch * 32 + 1023) & ~1023)) + 1024 + 256;
    p.stages = (224 * 1024 - fixed) / stage_bytes;
    if (p.stages > kMaxStages1) p.stages = kMaxStages1;
    CB_REQUIRE(p.stages >= 1, CB_E_ARG, "conv1: tile does not fit shared memory");
    const int smem = fixed + p.stages * stage_bytes;
    static bool attr = false;
    if (!attr) {
        CB_CUDA(cudaFuncSetAttribute(umma_conv1_kernel<WGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = p.n < sms ? p.n : sms;
    umma_conv1_kernel<WGRAD><<<grid, kThreads, smem, s>>>(p);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // namespace

namespace cb {

int scale_buffer(float* x, int64_t n, float beta, cudaStream_t s);
int colsum_bf16(const void* dy, int64_t rows, int c, float* out, float beta, cudaStream_t s);

bool conv1_supports(const cb_conv_desc* d) { return cb_device_supports_umma() && shape_ok(d); }

// w_native: bf16 copy of the torch weight [32][cin][8][8] (K order c, kh, kw).  A second network
// (w2/bias2/out2) may share the same input pass (RND target + predictor, rnd.py:174-177).
int conv1_fwd(const cb_conv_desc* d, const void* x, const void* w_native, const cb_epilogue* e, const void* w2,
              const float* bias2, void* out2, cudaStream_t s) {
    CB_REQUIRE(x && w_native && e && e->out_bf16 && !e->out_f32 && !e->residual && !e->side_dim, CB_E_ARG,
               "conv1 forward: needs a bf16 output and no residual/side operands");
    CB_REQUIRE(!w2 || out2, CB_E_ARG, "conv1 forward: second network needs an output");
    Conv1Params p{};
    int rc = fill(p, d, x, e->norm_mean, e->norm_rstd); if (rc) return rc;
    p.nets = w2 ? 2 : 1;
    p.w_bf16[0] = (const __nv_bfloat16*)w_native; p.bias[0] = e->bias; p.out[0] = (__nv_bfloat16*)e->out_bf16;
    p.w_bf16[1] = (const __nv_bfloat16*)w2; p.bias[1] = bias2; p.out[1] = (__nv_bfloat16*)out2;
    p.act = d->act;
    if (s2d_ok(d) && ((uintptr_t)x & 15) == 0) return launch_s2d_fwd(p, s);
    return launch<false>(p, s);
}

// dw: fp32 [32][cin*64] in the torch layout (no re-ordering needed)
int conv1_wgrad(const cb_conv_desc* d, const void* x, const void* dy, const float* mean, const float* rstd, float* dw,
                float* dbias, float beta, cudaStream_t s) {
    CB_REQUIRE(x && dy && dw, CB_E_ARG, "conv1 wgrad: null pointer");
    Conv1Params p{};
    int rc = fill(p, d, x, mean, rstd); if (rc) return rc;
    p.dy = (const __nv_bfloat16*)dy; p.dw = dw; p.nets = 1;
    rc = scale_buffer(dw, (int64_t)32 * d->in_c * 64, beta, s); if (rc) return rc;
    const bool s2d = s2d_ok(d) && s2d_wgrad_ok(d) && (((uintptr_t)x | (uintptr_t)dy) & 15) == 0;
    // Bias gradient: in the s2d kernel dy is summed on the tensor cores (the same A windows times an N = 16 block
    // of ones, 29 more small MMAs per sample) instead of a second pass over the 25.6 KB/sample of dy.
    if (dbias) {
        if (s2d) { p.dbias = dbias; rc = scale_buffer(dbias, 32, beta, s); }
        else rc = colsum_bf16(dy, (int64_t)d->n * p.rows, 32, dbias, beta, s);
        if (rc) return rc;
    }
    if (s2d)
        return launch_s2d_wgrad(p, s);
    return launch<true>(p, s);
}

}  // namespace cb

extern "C" int cb_conv1_supported(const cb_conv_desc* d) { return d && cb::conv1_supports(d) ? 1 : 0; }

extern "C" int cb_conv1_fwd(const cb_conv_desc* d, const void* x, const void* w_native, const cb_epilogue* e,
                            const void* w2_native, const float* bias2, void* out2_bf16, void* stream) {
    CB_REQUIRE(d && cb::conv1_supports(d), CB_E_ARG, "cb_conv1_fwd: unsupported shape or device");
    return cb::conv1_fwd(d, x, w_native, e, w2_native, bias2, out2_bf16, cb::as_stream(stream));
}

extern "C" int cb_conv1_wgrad(const cb_conv_desc* d, const void* x, const void* dy, const float* norm_mean,
                              const float* norm_rstd, float* dw_native, float* dbias, float beta, void* stream) {
    CB_REQUIRE(d && cb::conv1_supports(d), CB_E_ARG, "cb_conv1_wgrad: unsupported shape or device");
    return cb::conv1_wgrad(d, x, dy, norm_mean, norm_rstd, dw_native, dbias, beta, cb::as_stream(stream));
}

// Persistent, warp-specialised tcgen05 GEMM with TMA-fed K-major operands (sm_100a).
//
//   D[rows, n] = epilogue( sum_k A[rows, k] * B[n, k] )
//
// A rows are *gathered by TMA*: the A tensor map is a rank-2..5 view of the activation tensor
// and every k-block is one box of <=256 rows x 64 bf16 (128 B, SWIZZLE_128B), so an implicit-GEMM
// convolution is the same kernel as a Linear layer with a different view (see umma_conv.cu for
// the views).  B is the [N, K] bf16 weight matrix.  fp32 accumulators live in TMEM, double
// buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM -> registers -> bias/side/activation/residual -> global).
#pragma once
#include "common.cuh"
#include "umma_common.cuh"

namespace umma {

// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2.. = epilogue.  A warp may
// only read the TMEM lane quarter (warp % 4), so the epilogue warps come in groups of four; the
// kEpiGroups warps sharing a quarter split a tile's (sub-tile, 32-column chunk) items between them.
// With one warp per scheduler every tcgen05.ld / global-load latency is exposed (ncu: issue slots 22 %
// busy, tensor pipe 46 %); several warps per scheduler hide it.
constexpr int kEpiWarp0 = 2;
// SIMPLE needs few registers: 16 warps.  The epilogues with global side operands run 12 warps (3 per quarter, 128
// registers each) and fetch those operands on demand instead of 8 warps with a two-deep register prefetch (167
// registers): conv2 dgrad 2.41 -> 1.85 ms.  (A producer-side cp.async.bulk.prefetch.L2 of the tile's side
// operands helped the 8-warp version but, with 12 warps, only added 2.4 GB of duplicate DRAM reads -- ncu: 7.07 GB
// read for 4.7 GB algorithmic -- to a kernel that already runs at 80 % of HBM peak; removed.)
template <int EPI> struct EpiWarps { static constexpr int GROUPS = (EPI == 0) ? 4 : 3; };
template <int EPI> constexpr int gemm_threads() { return (kEpiWarp0 + 4 * EpiWarps<EPI>::GROUPS) * 32; }
// The plain GEMM kernel's full / dgrad epilogues keep 8 warps and the two-deep register prefetch (its tiles' side
// operands are strided blocks with longer exposed latency).
template <int EPI> struct GemmEpiWarps { static constexpr int GROUPS = (EPI == 0) ? 4 : 2; };
template <int EPI> constexpr int plain_gemm_threads() { return (kEpiWarp0 + 4 * GemmEpiWarps<EPI>::GROUPS) * 32; }

constexpr int kBiasSmemFloats = 4096;   // bias vectors up to this length are staged in shared memory

struct CoordMap {          // coordinate = ((index / div) % mod) * mul   (mod == 0: no modulo)
    int div[5], mod[5], mul[5];
};

struct GemmParams {
    // operand A view
    int a_rank;
    CoordMap a_tile;       // from the m-tile index
    CoordMap a_kb;         // from the k-block index
    int a_off[5];          // constant coordinate offset (e.g. -pad)
    int a_box_bytes;       // bytes one A box deposits in shared memory
    int kb_count;          // K / 64
    // tiling
    int m_tiles, n_tiles;
    // grouped (block-diagonal) GEMM: n-tiles [g*group_tiles, (g+1)*group_tiles) belong to problem g, whose
    // A operand starts a_group_koff elements further along K (0: all problems share A).  group_tiles == 0: off.
    int group_tiles, a_group_koff;
    // side input folded into the GEMM: the LAST k-block of A comes from the tmS tensor map (bf16 [rows, 64],
    // the one-hot action zero-padded to 64 columns) and B carries the matching 64 weight columns -- the
    // torch.cat([x, action]) of icm.py:19-21 done by the tensor cores instead of 32*side_dim FMAs per
    // output chunk in the epilogue.
    int side_kb;
    // forward-simple epilogue only: the bias was already added by the tensor cores (an extra K = 16 MMA of a ones
    // block against a hi/mid/lo bf16 split of the bias), so the epilogue is activation + pack + store
    int bias_in_acc;
    int rows_per_tile;     // valid rows of each tile (<= MT*128)
    long long m_total;     // total output rows
    int n_total;           // total output columns (row stride of all row-major operands below)
    // epilogue
    const float* bias;
    const float* side; const float* side_w; int side_dim;
    const __nv_bfloat16* add_pre;      // added before the activation (dgrad skip gradient)
    int act;                           // activation on the result
    const __nv_bfloat16* actgrad_src;  // multiply by act'(this), derivative from the saved output
    int actgrad_act;
    const __nv_bfloat16* residual;     // added after the activation
    __nv_bfloat16* out_bf16; float* out_f32;
    int debug;                         // CB_EPI_DEBUG (profiling only): 1 = skip the epilogue math and stores
    // optional pixel shuffle of the output (stride-2 dgrad): row = (n, a, b) over an (sh_ph x sh_pw)
    // grid per sample, 32-column chunk q = (ph, pw) -> NHWC pixel (2a+ph, 2b+pw) of a [*, sh_h, sh_w, 32]
    // tensor.  0 = plain row-major [*, n_total].
    int shuffle; int sh_ph, sh_pw, sh_h, sh_w;
};

template <int BN, int MT>
struct GemmConfig {
    static constexpr int A_BYTES = MT * 128 * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
    static constexpr int ACC_COLS = MT * BN;                       // one accumulator stage
    // two accumulator stages (epilogue of tile i under the MMAs of tile i+1) when they fit; a 256x256 tile fills
    // all 512 TMEM columns and runs single-buffered (its MMAs last ~50k cycles, the drain ~2k)
    static constexpr int ACC_STAGES = 2 * ACC_COLS <= 512 ? 2 : 1;
    static constexpr int TMEM_COLS = ACC_STAGES * ACC_COLS <= 32 ? 32 : ACC_STAGES * ACC_COLS <= 64 ? 64
                                     : ACC_STAGES * ACC_COLS <= 128 ? 128 : ACC_STAGES * ACC_COLS <= 256 ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + kBiasSmemFloats * 4;
    static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
    static_assert(STAGE_BYTES % 1024 == 0, "stages must stay 1024-byte aligned for SWIZZLE_128B");
    static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
};

__device__ __forceinline__ int map_coord(const CoordMap& m, int i, int idx) {
    int v = idx / m.div[i];
    if (m.mod[i]) v %= m.mod[i];
    return v * m.mul[i];
}

// Element offset of the 32-column chunk starting at column n0 of an output row = row_base + chunk_delta:
//   plain layout    row_base = grow * n_total,                       delta = n0
//   depth-to-space  row_base = pixel (2a, 2b) of the NHWC output,    delta = ((ph * sh_w) + pw) * 32, chunk q = (ph, pw)
__device__ __forceinline__ long long row_base(const GemmParams& p, long long grow) {
    if (!p.shuffle) return grow * (long long)p.n_total;
    const int pp = p.sh_ph * p.sh_pw;
    const long long n = grow / pp;
    const int rem = (int)(grow - n * pp);
    const int a = rem / p.sh_pw, b = rem - a * p.sh_pw;
    return ((n * p.sh_h + 2 * a) * p.sh_w + 2 * b) * 32;
}
__device__ __forceinline__ int chunk_delta(const GemmParams& p, int n0) {
    if (!p.shuffle) return n0;
    const int q = n0 >> 5;
    return ((q >> 1) * p.sh_w + (q & 1)) * 32;
}

// 32 consecutive bf16 of one row -> fp32 (four 16-byte loads)
__device__ __forceinline__ void load_bf16_row(float (&g)[32], const __nv_bfloat16* src) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 u = __ldg(s4 + q);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h[e]);
            g[q * 8 + e * 2] = f.x; g[q * 8 + e * 2 + 1] = f.y;
        }
    }
}
__device__ __forceinline__ void add_bf16_row(float (&v)[32], const __nv_bfloat16* src) {
    float g[32];
    load_bf16_row(g, src);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += g[j];
}

// Row-major side operands of one 32-column chunk, fetched ahead of use (they are global-memory
// reads on the epilogue's critical path: skip gradient, saved activation, residual).
struct AuxRegs { uint4 other[4], grad[4]; };   // other = add_pre (dgrad) or residual (forward), never both

__device__ __forceinline__ void aux_load(const GemmParams& p, long long co, AuxRegs& a) {
    const __nv_bfloat16* other = p.add_pre ? p.add_pre : p.residual;     // skip gradient (dgrad) or residual (forward)
    if (other) {
        const U256 lo = ld_global_nc_256(other + co), hi = ld_global_nc_256(other + co + 16);
        a.other[0] = make_uint4(lo.w[0], lo.w[1], lo.w[2], lo.w[3]); a.other[1] = make_uint4(lo.w[4], lo.w[5], lo.w[6], lo.w[7]);
        a.other[2] = make_uint4(hi.w[0], hi.w[1], hi.w[2], hi.w[3]); a.other[3] = make_uint4(hi.w[4], hi.w[5], hi.w[6], hi.w[7]);
    }
    if (p.actgrad_src) {
        const U256 lo = ld_global_nc_256(p.actgrad_src + co), hi = ld_global_nc_256(p.actgrad_src + co + 16);
        a.grad[0] = make_uint4(lo.w[0], lo.w[1], lo.w[2], lo.w[3]); a.grad[1] = make_uint4(lo.w[4], lo.w[5], lo.w[6], lo.w[7]);
        a.grad[2] = make_uint4(hi.w[0], hi.w[1], hi.w[2], hi.w[3]); a.grad[3] = make_uint4(hi.w[4], hi.w[5], hi.w[6], hi.w[7]);
    }
}

__device__ __forceinline__ void unpack_bf16x32(const uint4 (&u)[4], float (&g)[32]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[q]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h[e]);
            g[q * 8 + e * 2] = f.x; g[q * 8 + e * 2 + 1] = f.y;
        }
    }
}

// Copies the bias vector into shared memory (all threads; visible after the kernel's first
// __syncthreads) and returns the pointer the epilogue should read: smem copy, global, or null.
__device__ __forceinline__ const float* stage_bias(const GemmParams& p, uint8_t* smem_area) {
    if (!p.bias) return nullptr;
    if (p.n_total > kBiasSmemFloats) return p.bias;
    float* sb = reinterpret_cast<float*>(smem_area);
    for (int i = threadIdx.x; i < p.n_total; i += blockDim.x) sb[i] = __ldg(p.bias + i);
    return sb;
}

// Epilogue specialisations (compile-time pruning keeps the hot loop small enough for the I-cache):
//   EPI_SIMPLE  bias + activation -> bf16                       (conv layers, most Linears)
//   EPI_FULL    bias + side input + activation + residual -> bf16 and/or fp32
//   EPI_DGRAD   (+ skip gradient) * act'(saved output) -> bf16   (optionally depth-to-space)
enum { EPI_SIMPLE = 0, EPI_FULL = 1, EPI_DGRAD = 2 };

// One 32-column chunk of one output row.  `co` = element offset of the chunk in every row-major
// operand, n0 = first column (for bias / side weights), sv = the row's side inputs.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const float* bias, const uint32_t (&raw)[32],
                                               long long co, int n0, const float (&sv)[16], const AuxRegs& aux) {
    if (EPI == EPI_SIMPLE && p.bias_in_acc && p.act != CB_ACT_ELU) {
        // accumulator = conv + bias: activation on packed bf16 pairs (the negative LeakyReLU branch is rounded twice,
        // 2^-9 relative on values scaled by 0.01), then two 32-byte stores
        const __nv_bfloat162 slope = __floats2bfloat162_rn(0.01f, 0.01f), zero = __floats2bfloat162_rn(0.f, 0.f);
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(raw[2 * e]), __uint_as_float(raw[2 * e + 1]));
            if (p.act == CB_ACT_LRELU) h = __hmax2(h, __hmul2(h, slope));
            else if (p.act == CB_ACT_RELU) h = __hmax2(h, zero);
            pk[e] = *reinterpret_cast<uint32_t*>(&h);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            U256 u;
#pragma unroll
            for (int e = 0; e < 8; ++e) u.w[e] = pk[q * 8 + e];
            st_global_256(p.out_bf16 + co + q * 16, u);
        }
        return;
    }
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
    if (EPI != EPI_DGRAD && bias && !(EPI == EPI_SIMPLE && p.bias_in_acc)) {   // `bias`: the shared-memory copy when it fits
        const float4* bsrc = reinterpret_cast<const float4*>(bias + n0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 b4 = bsrc[q];
            v[q * 4] += b4.x; v[q * 4 + 1] += b4.y; v[q * 4 + 2] += b4.z; v[q * 4 + 3] += b4.w;
        }
    }
    if (EPI == EPI_FULL && p.side_dim) {
        const float* sw = p.side_w + (long long)n0 * p.side_dim;
        for (int s = 0; s < p.side_dim; ++s) {
            const float a = sv[s];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(a, __ldg(sw + j * p.side_dim + s), v[j]);
        }
    }
    if (EPI == EPI_DGRAD && p.add_pre) {
        float g[32];
        unpack_bf16x32(aux.other, g);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += g[j];
    }
    if (EPI != EPI_DGRAD) {
        if (p.act == CB_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (p.act == CB_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.01f * v[j]);
        } else if (p.act == CB_ACT_ELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : expm1f(v[j]);
        }
    }
    if (EPI == EPI_DGRAD && p.actgrad_src) {
        float g[32];
        unpack_bf16x32(aux.grad, g);
        if (p.actgrad_act == CB_ACT_ELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= g[j] > 0.f ? 1.f : g[j] + 1.f;
        } else if (p.actgrad_act != CB_ACT_NONE) {
            const float slope = p.actgrad_act == CB_ACT_LRELU ? 0.01f : 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= g[j] > 0.f ? 1.f : slope;
        }
    }
    if (EPI == EPI_FULL && p.residual) {
        float g[32];
        unpack_bf16x32(aux.other, g);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += g[j];
    }
    if (EPI != EPI_FULL || p.out_bf16) {        // 64 B per row chunk: two 32-byte stores
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            U256 u;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[q * 16 + e * 2], v[q * 16 + e * 2 + 1]);
                u.w[e] = *reinterpret_cast<uint32_t*>(&h);
            }
            st_global_256(p.out_bf16 + co + q * 16, u);
        }
    }
    if (EPI == EPI_FULL && p.out_f32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            U256 u;
#pragma unroll
            for (int e = 0; e < 8; ++e) u.w[e] = __float_as_uint(v[q * 8 + e]);
            st_global_256(p.out_f32 + co + q * 8, u);
        }
    }
}

// Epilogue of one accumulator stage for one warp: MT*BN/32 (sub-tile, 32-column chunk) items, rolled
// (an unrolled epilogue overflows the instruction cache).  For the modes with global side operands
// the loop is software-pipelined two items deep with alternating register buffers: the operands of
// item i+1 are requested before item i is processed, and of item 0 before the accumulator is awaited,
// so their HBM latency overlaps the MMAs and the math.  (Copying a prefetch buffer would stall on the
// pending loads -- the buffers are only ever consumed.)
// row_fn(mt, valid, grow, base): output row of this lane in sub-tile mt, and its element offset (row_base).
template <int BN, int MT, int EPI, int G, bool PIPE, class RowFn>      // G warps per quarter: this warp handles items sub, sub + G, ...
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const float* bias, uint32_t tmem_acc, int quarter,
                                              int sub, int lane, int n_base, uint32_t tfull, uint32_t parity, RowFn row_fn) {
    constexpr int NCH = BN / 32;
    constexpr int ITEMS = MT * NCH;
    bool valid[MT]; long long grow[MT], base[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) row_fn(mt, valid[mt], grow[mt], base[mt]);
    float sv[16];
    int sv_mt = -1;
#pragma unroll
    for (int s = 0; s < 16; ++s) sv[s] = 0.f;
    auto row_valid = [&](int mt) { return MT == 1 ? valid[0] : (mt ? valid[MT - 1] : valid[0]); };
    auto row_index = [&](int mt) { return MT == 1 ? grow[0] : (mt ? grow[MT - 1] : grow[0]); };
    auto row_off = [&](int mt) { return MT == 1 ? base[0] : (mt ? base[MT - 1] : base[0]); };
    auto request = [&](int it, AuxRegs& buf) {
        if (EPI == EPI_SIMPLE || it >= ITEMS) return;
        const int mt = it / NCH, cb = it - mt * NCH;
        if (row_valid(mt)) aux_load(p, row_off(mt) + chunk_delta(p, n_base + cb * 32), buf);
    };
    auto process = [&](int it, const AuxRegs& buf) {
        const int mt = it / NCH, cb = it - mt * NCH;
        const bool v_now = row_valid(mt);
        const long long g_now = row_index(mt);
        if (EPI == EPI_FULL && p.side_dim && mt != sv_mt) {      // the row's side inputs, once per sub-tile
            sv_mt = mt;
#pragma unroll
            for (int s = 0; s < 16; ++s)
                sv[s] = (v_now && s < p.side_dim) ? __ldg(p.side + g_now * p.side_dim + s) : 0.f;
        }
        uint32_t raw[32];
        tmem_ld_32x32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * BN + cb * 32), raw);
        tmem_ld_wait();
        if (v_now && !(p.debug & 1))
            epilogue_chunk<EPI>(p, bias, raw, row_off(mt) + chunk_delta(p, n_base + cb * 32), n_base + cb * 32, sv, buf);
    };
    AuxRegs b0;
    request(sub, b0);                       // first item's side operands: in flight while the accumulator is awaited
    mbar_wait(tfull, parity);
    fence_after_thread_sync();
    if (EPI == EPI_SIMPLE || !PIPE) {
#pragma unroll 1
        for (int it = sub; it < ITEMS; it += G) {
            if (EPI != EPI_SIMPLE && it != sub) request(it, b0);
            process(it, b0);
        }
    } else {                                // two-deep software pipeline with alternating register buffers
        AuxRegs b1;
#pragma unroll 1
        for (int it = sub; it < ITEMS; it += 2 * G) {
            request(it + G, b1);
            process(it, b0);
            if (it + G < ITEMS) {
                request(it + 2 * G, b0);
                process(it + G, b1);
            }
        }
    }
}

template <int BN, int MT, int EPI>
__global__ void __launch_bounds__(plain_gemm_threads<EPI>(), 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmS, const GemmParams p) {
    using Cfg = GemmConfig<BN, MT>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base word
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* bias = stage_bias(p, smem_raw + (bar_base + 256u - smem_u32(smem_raw)));
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if (p.side_kb) prefetch_tmap(&tmS);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4 * GemmEpiWarps<EPI>::GROUPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int total_tiles = p.m_tiles * p.n_tiles;
    const int kb_count = p.kb_count;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int tm = tile / p.n_tiles, tn = tile % p.n_tiles;
                int tc[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) tc[i] = p.a_off[i] + map_coord(p.a_tile, i, tm);
                if (p.group_tiles) tc[0] += (tn / p.group_tiles) * p.a_group_koff;
                for (int kb = 0; kb < kb_count; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(p.a_box_bytes + Cfg::B_BYTES));
                    int c[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) c[i] = tc[i] + map_coord(p.a_kb, i, kb);
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    if (p.side_kb && kb == kb_count - 1) tma_load_2d(sa, &tmS, full_bar(stage), 0, tc[1]);
                    else tma_load(p.a_rank, sa, &tmA, full_bar(stage), c);
                    tma_load_2d(sa + Cfg::A_BYTES, &tmB, full_bar(stage), kb * 64, tn * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // (One lane runs the loop here.  The window / weight-gradient / conv1 kernels instead keep the whole
        // warp converged and elect a lane per instruction -- elect_one() -- which removes the per-MMA operand
        // "waterfall" and is worth 10-35 % where many short MMAs are issued per stage; with 4-8 MMAs per
        // k-block the extra 32-lane barrier polling costs more than it saves: measured 7 % slower.)
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                fence_after_thread_sync();
                for (int kb = 0; kb < kb_count; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    fence_after_thread_sync();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    // descriptors differ only in the start-address field: add (bytes >> 4) to the low word
                    const uint64_t da0 = make_smem_desc(sa, 16, 1024);
                    const uint64_t db0 = make_smem_desc(sa + Cfg::A_BYTES, 16, 1024);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)((acc * MT + mt) * BN);
#pragma unroll
                        for (int k = 0; k < 4; ++k)             // 4 x UMMA_K(16) = 64
                            mma_f16_ss(d_tmem, da0 + (uint64_t)(mt * 1024 + k * 2), db0 + (uint64_t)(k * 2), idesc,
                                       (kb | k) ? 1u : 0u);
                    }
                    mma_commit(empty_bar(stage));             // frees the smem stage when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                mma_commit(tfull_bar(acc));                   // accumulator ready for the epilogue
                if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ------------------------------------------------------------------ epilogue
        const int quarter = warp & 3;                         // TMEM lane quarter this warp may read
        const int sub = (warp - kEpiWarp0) >> 2;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int tm = tile / p.n_tiles, tn = tile % p.n_tiles;
            epilogue_tile<BN, MT, EPI, GemmEpiWarps<EPI>::GROUPS, true>(p, bias, tmem_base + (uint32_t)(acc * MT * BN), quarter, sub, lane, tn * BN,
                                  tfull_bar(acc), acc_phase,
                                  [&](int mt, bool& valid, long long& grow, long long& base) {
                                      const int r = mt * 128 + quarter * 32 + lane;
                                      grow = (long long)tm * p.rows_per_tile + r;
                                      valid = r < p.rows_per_tile && grow < p.m_total;
                                      base = row_base(p, grow);
                                  });
            fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        }
    }

    fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_thread_sync();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace umma

// "Window" kernels: convolutions whose input slab is loaded into shared memory ONCE per tile and
// whose filter taps are read as row-shifted windows of that slab.
//
// On B200 the 128-byte swizzle is a pure function of the shared-memory address (pinned by
// tests/test_gpu_umma_probe.py), so a UMMA descriptor may start at any 128-byte row of a
// TMA-written slab.  For a stride-1 convolution over a gh x gw pixel grid, the A operand of tap
// (kh, kw) for "output position = every grid pixel" is the slab itself shifted by kh*gw + kw rows;
// grid pixels whose window would leave the image produce rows that are simply not stored.  This
// trades (gh*gw)/(oh*ow) extra MMA rows for a k_h*k_w-fold cut of L2->SM traffic, which is what
// bounds these skinny-N (32/64 channel) layers.  The weights stay resident in shared memory.
//
//   window_conv : K-major A (slab rows = pixels), B = resident [N][K] weights        fwd / dgrad
//   window_wgrad: MN-major A = slab (M = channel chunk of one or two taps, K = pixels),
//                 MN-major B = dy on the same grid (zero outside the valid outputs);
//                 accumulates over all slabs of a CTA in TMEM, then atomically adds to dW
#pragma once
#include "umma_gemm.cuh"

namespace umma {

constexpr int kMaxPlanes = 2;
constexpr int kMaxTaps = 16;
constexpr int kMaxStages = 8;

struct WindowParams {
    // slab: nplanes TMA boxes per tile, each `plane_rows` rows of 128 B
    int nplanes, plane_rows, plane_bytes, stage_bytes, stages, slack_bytes;
    int a_rank, a_sample_dim, S;           // coord[a_sample_dim] = tile * S
    int a_coord[kMaxPlanes][5];            // constant start coordinates of each plane
    // taps (k-blocks of 64): plane and row shift
    int kb_count, kb_plane[kMaxTaps], kb_shift[kMaxTaps];
    int kb_off16[kMaxTaps];                // (plane*plane_bytes + shift*128) >> 4, filled by the host
    // pixel grid per sample and its valid (stored) part
    int gh, gw, vh, vw, n_samples, tiles;
    // flat mode (flat_rows > 0): a tile is `flat_rows` consecutive rows of the (sample, grid row) axis instead of
    // S whole samples, so the 256 MMA rows are filled whatever gh*gw is (10x10 grid: 250 of 256 instead of 200)
    int flat_rows;
    GemmParams epi;                        // epilogue operands (n_total == BN, shuffle fields honoured)
};

template <int BN, int EPI>
__global__ void __launch_bounds__(gemm_threads<EPI>(), 1)
umma_window_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                        const WindowParams p) {
    constexpr int MT = 2;
    constexpr int TMEM_COLS = (2 * MT * BN <= 256) ? 256 : 512;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t w_bytes = (uint32_t)p.kb_count * BN * 128;
    const uint32_t a_base = smem_base + w_bytes;
    const uint32_t a_bytes = (uint32_t)p.stages * p.stage_bytes + p.slack_bytes;
    const uint32_t bar_base = a_base + a_bytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 2 + a); };
    const uint32_t w_bar = bar_base + 8u * (2 * kMaxStages + 4);
    const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 5);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* bias = stage_bias(p.epi, gen_base + (bar_base + 256u - smem_base));
    // bias-by-MMA operands (SWIZZLE_32B, rows of 32 B): ones [128 rows], bias split [BN rows]
    const uint32_t ones_base = bar_base + 2048u;
    const uint32_t bblk_base = ones_base + 4096u;
    if (EPI == EPI_SIMPLE && p.epi.bias_in_acc) {
        uint8_t* ob = gen_base + (ones_base - smem_base);
        for (uint32_t i = threadIdx.x; i < (4096u + BN * 32u) / 16; i += blockDim.x)
            reinterpret_cast<uint4*>(ob)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        auto swz32 = [](uint32_t off) { return off ^ (((off >> 7) & 1u) << 4); };
        if (threadIdx.x < 128)
            *reinterpret_cast<uint2*>(ob + swz32(threadIdx.x * 32u)) = make_uint2(0x3f803f80u, 0x00003f80u);
        if ((int)threadIdx.x < BN) {
            const float b = __ldg(p.epi.bias + threadIdx.x);
            const __nv_bfloat16 hi = __float2bfloat16_rn(b);
            const float r1 = b - __bfloat162float(hi);
            const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
            const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
            *reinterpret_cast<uint2*>(ob + 4096u + swz32(threadIdx.x * 32u)) =
                make_uint2((uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16),
                           (uint32_t)__bfloat16_as_ushort(lo));
        }
    }
    // one-time zero of the slab area: rows a shifted window reads past a slab are then always finite
    {
        uint4* z = reinterpret_cast<uint4*>(gen_base + w_bytes);
        for (uint32_t i = threadIdx.x; i < a_bytes / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmW);
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4 * EpiWarps<EPI>::GROUPS); }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // resident weights: kb_count boxes of {64, BN}
            mbar_arrive_expect_tx(w_bar, w_bytes);
            for (int kb = 0; kb < p.kb_count; ++kb)
                tma_load_2d(smem_base + kb * BN * 128, &tmW, w_bar, kb * 64, 0);
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(p.nplanes * p.plane_rows * 128));
                for (int pl = 0; pl < p.nplanes; ++pl) {
                    int c[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) c[i] = p.a_coord[pl][i];
                    c[p.a_sample_dim] += tile * (p.flat_rows ? p.flat_rows : p.S);
                    tma_load(p.a_rank, a_base + stage * p.stage_bytes + pl * p.plane_bytes, &tmA, full_bar(stage), c);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // all lanes run the loops (warp-uniform state); one elected lane issues (see umma_gemm_kernel)
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
        mbar_wait(w_bar, 0);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            mbar_wait(full_bar(stage), phase);
            fence_after_thread_sync();
            const uint64_t da0 = make_smem_desc(a_base + stage * p.stage_bytes, 16, 1024);
            const uint64_t db0 = make_smem_desc(smem_base, 16, 1024);
            const bool bias_mma = EPI == EPI_SIMPLE && p.epi.bias_in_acc;
            // SWIZZLE_32B descriptors (layout code 6) of the ones / bias blocks
            const uint64_t d_ones = (make_smem_desc(ones_base, 16, 256) & ~((uint64_t)7 << 61)) | ((uint64_t)6 << 61);
            const uint64_t d_bias = (make_smem_desc(bblk_base, 16, 256) & ~((uint64_t)7 << 61)) | ((uint64_t)6 << 61);
            if (leader) {
                if (bias_mma) {
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
                        mma_f16_ss(tmem_base + (uint32_t)((acc * MT + mt) * BN), d_ones, d_bias, idesc, 0u);     // D = bias
                }
                for (int kb = 0; kb < p.kb_count; ++kb) {
                    const uint64_t da = da0 + (uint64_t)p.kb_off16[kb];
                    const uint64_t db = db0 + (uint64_t)(kb * BN * 8);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)((acc * MT + mt) * BN);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_f16_ss(d_tmem, da + (uint64_t)(mt * 1024 + k * 2), db + (uint64_t)(k * 2), idesc,
                                       (bias_mma || (kb | k)) ? 1u : 0u);
                    }
                }
                mma_commit(empty_bar(stage));
                mma_commit(tfull_bar(acc));
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else if (warp >= kEpiWarp0) {
        const int quarter = warp & 3;
        const int sub = (warp - kEpiWarp0) >> 2;
        const int grid_px = p.gh * p.gw;
        // Which output pixel a lane's row of each sub-tile is never changes from tile to tile: the divisions
        // are done once here, per tile only the sample base is added.
        bool ok[MT]; int smp[MT]; long long rel_row[MT], rel_base[MT];
        int fq[MT], fx[MT];                                // flat mode: grid row within the tile, grid column
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int r = mt * 128 + quarter * 32 + lane;
            const int s = r / grid_px, rem = r - s * grid_px;
            const int gy = rem / p.gw, gx = rem - gy * p.gw;
            fq[mt] = r / p.gw; fx[mt] = r - fq[mt] * p.gw;
            ok[mt] = p.flat_rows ? (fq[mt] < p.flat_rows && fx[mt] < p.vw) : (s < p.S && gy < p.vh && gx < p.vw);
            smp[mt] = s;
            rel_row[mt] = ((long long)s * p.vh + gy) * p.vw + gx;
            rel_base[mt] = p.epi.shuffle ? (((long long)s * p.epi.sh_h + 2 * gy) * p.epi.sh_w + 2 * gx) * 32
                                         : rel_row[mt] * p.epi.n_total;
        }
        const long long tile_rows = (long long)p.S * p.vh * p.vw;
        const long long tile_elems = p.epi.shuffle ? (long long)p.S * p.epi.sh_h * p.epi.sh_w * 32 : tile_rows * p.epi.n_total;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            epilogue_tile<BN, MT, EPI, EpiWarps<EPI>::GROUPS, false>(p.epi, bias, tmem_base + (uint32_t)(acc * MT * BN), quarter, sub, lane, 0, tfull_bar(acc),
                                  acc_phase, [&](int mt, bool& valid, long long& grow, long long& base) {
                                      const int m = MT == 1 ? 0 : (mt ? MT - 1 : 0);
                                      if (p.flat_rows) {       // (sample, grid row) from the flattened row index
                                          const int G = tile * p.flat_rows + fq[m];
                                          const int sample = G / p.gh, gy = G - sample * p.gh;
                                          valid = ok[m] && gy < p.vh && sample < p.n_samples;
                                          grow = ((long long)sample * p.vh + gy) * p.vw + fx[m];
                                          base = grow * p.epi.n_total;
                                          return;
                                      }
                                      valid = ok[m] && tile * p.S + smp[m] < p.n_samples;
                                      grow = tile * tile_rows + rel_row[m];
                                      base = tile * tile_elems + rel_base[m];
                                  });
            fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) { fence_after_thread_sync(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient: D_g[M=128][N=BN] += sum over K rows of A_g^T B, both operands MN-major, K = rows
// (samples or grid pixels) of the stage.  One CTA owns a set of accumulators for its whole K range
// (split-K across CTAs) and finally atomically adds them into dW.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxAcc = 8;

struct WgradParams {
    // stage contents: na + nb TMA boxes
    int na, nb;
    int a_rank, b_rank;
    int a_coord[4][5], b_coord[4][5];      // constant start coordinates per box
    int a_step_dim, b_step_dim, a_step, b_step;     // coord[step_dim] += slab * step
    int a_box_bytes, b_box_bytes;          // bytes deposited per box
    int a_box_stride, b_box_stride;        // smem distance between consecutive boxes of an operand
    int stage_bytes, stages, slack_bytes, b_offset;  // B boxes start at stage + b_offset
    int k_rows;                            // K rows consumed per slab (multiple of 16)
    // work partition: CTA -> (unit, split); unit selects coordinate offsets for A/B columns
    int units, splits, slabs_total;
    int unit_div;                          // a_unit = unit / unit_div, b_unit = unit % unit_div
    int a_unit_dim, a_unit_step, b_unit_dim, b_unit_step;
    // grouped weight gradient: a_units [g*group_units, ...) belong to problem g, whose B (input) columns start
    // b_group_koff further (0: shared input).  group_units == 0: off.
    int group_units, b_group_koff;
    // accumulators
    int n_acc;
    int acc_a_off[kMaxAcc];                // byte offset of the accumulator's A window inside the stage
    int acc_a_lbo[kMaxAcc];                // distance between its two 64-wide M chunks
    // output: dw index = base(lane<64 ? 0 : 1) + (lane%64)*m_stride + col*n_stride (+ unit offsets)
    long long acc_base0[kMaxAcc], acc_base1[kMaxAcc];      // base1 < 0: upper half unused
    long long m_stride, n_stride, unit_m_off, unit_n_off;
    int n_cols_total, n_unit_cols;         // valid columns: min(BN, n_cols_total - unit_n*n_unit_cols)
    int m_total, m_unit_rows;              // row-major weight gradients whose output dimension is not a multiple of 128:
                                           // accumulator row a_unit*m_unit_rows + g*128 + row is stored only if < m_total
    float* dw;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
umma_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t a_bytes = (uint32_t)p.stages * p.stage_bytes + p.slack_bytes;
    const uint32_t bar_base = smem_base + a_bytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
    const uint32_t done_bar = bar_base + 8u * (2 * kMaxStages);
    const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 1);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        uint4* z = reinterpret_cast<uint4*>(gen_base);
        for (uint32_t i = threadIdx.x; i < a_bytes / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int unit = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
    const int a_unit = unit / p.unit_div, b_unit = unit % p.unit_div;
    // slabs of this split: split, split + splits, ...
    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int slab = split; slab < p.slabs_total; slab += p.splits) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(p.na * p.a_box_bytes + p.nb * p.b_box_bytes));
                const uint32_t st = smem_base + stage * p.stage_bytes;
                for (int i = 0; i < p.na; ++i) {
                    int c[5];
#pragma unroll
                    for (int d = 0; d < 5; ++d) c[d] = p.a_coord[i][d];
                    c[p.a_step_dim] += slab * p.a_step;
                    c[p.a_unit_dim] += a_unit * p.a_unit_step;
                    tma_load(p.a_rank, st + i * p.a_box_stride, &tmA, full_bar(stage), c);
                }
                for (int i = 0; i < p.nb; ++i) {
                    int c[5];
#pragma unroll
                    for (int d = 0; d < 5; ++d) c[d] = p.b_coord[i][d];
                    c[p.b_step_dim] += slab * p.b_step;
                    c[p.b_unit_dim] += b_unit * p.b_unit_step;
                    if (p.group_units) c[p.b_unit_dim] += (a_unit / p.group_units) * p.b_group_koff;
                    tma_load(p.b_rank, st + p.b_offset + i * p.b_box_stride, &tmB, full_bar(stage), c);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();          // all lanes loop, one issues (see umma_gemm_kernel)
        constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);
        int stage = 0; uint32_t phase = 0;
        uint32_t first = 1;
        for (int slab = split; slab < p.slabs_total; slab += p.splits) {
            mbar_wait(full_bar(stage), phase);
            fence_after_thread_sync();
            const uint32_t st = smem_base + stage * p.stage_bytes;
            const int ksteps = p.k_rows / 16;
            if (leader) {
                for (int g = 0; g < p.n_acc; ++g) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(g * BN);
                    const uint64_t da0 = make_smem_desc(st + p.acc_a_off[g], (uint32_t)p.acc_a_lbo[g], 1024);
                    const uint64_t db0 = make_smem_desc(st + p.b_offset, (uint32_t)p.b_box_stride, 1024);
                    for (int k = 0; k < ksteps; ++k)          // one k-step = 16 rows of 128 B = 2048 B
                        mma_f16_ss(d_tmem, da0 + (uint64_t)(k * 128), db0 + (uint64_t)(k * 128), idesc,
                                   (first && k == 0) ? 0u : 1u);
                }
                mma_commit(empty_bar(stage));
            }
            __syncwarp();
            first = 0;
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (leader) mma_commit(done_bar);
        __syncwarp();
    } else {
        // epilogue: wait for the CTA's whole K range, then atomically add into dW
        const int quarter = warp & 3;
        const bool any = split < p.slabs_total;
        if (any) {
            mbar_wait(done_bar, 0);
            fence_after_thread_sync();
            const int row = quarter * 32 + lane;                 // accumulator row (TMEM lane)
            int ncols = BN;
            if (p.n_cols_total) {
                int left = p.n_cols_total - b_unit * p.n_unit_cols;
                ncols = left < BN ? left : BN;
            }
            for (int g = 0; g < p.n_acc; ++g) {
                // (the tcgen05.ld below is warp-collective: a row past the end only skips its stores)
                const bool row_ok = !p.m_total || a_unit * p.m_unit_rows + g * 128 + row < p.m_total;
                const long long base = row < 64 ? p.acc_base0[g] : p.acc_base1[g];
                const long long o = base + (long long)(row & 63) * p.m_stride + (long long)a_unit * p.unit_m_off +
                                    (long long)b_unit * p.unit_n_off;
#pragma unroll
                for (int cb = 0; cb < BN / 32; ++cb) {
                    uint32_t raw[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * BN + cb * 32), raw);
                    tmem_ld_wait();
                    if (base >= 0 && row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = cb * 32 + j;
                            if (col < ncols) atomicAdd(p.dw + o + (long long)col * p.n_stride, __uint_as_float(raw[j]));
                        }
                    }
                }
            }
        }
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) { fence_after_thread_sync(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace umma

// RND trunk, first two convolutions (rnd.py:40-46 / 55-61) without the 20x20x32 activation between them ever
// leaving the SM:
//
//   cb_rnd_normalize_s2d   uint8 newest frame [84x84] --normalise, clamp +-5 (rnd.py:169-170)--> bf16 frame in
//                          4x4-pixel-block order ([21x21 blocks][16 px], 32-byte rows, pre-swizzled), ONCE per
//                          sample for both networks (target and predictor read the same normalised frame)
//   cb_rnd_conv12_fwd      block image --conv 8x8 s4 (1->32) + LeakyReLU--> slab in shared memory
//                                      --conv 4x4 s2 (32->64) + LeakyReLU--> a2 [n,9,9,64] bf16          (one network)
//
// Both layers run on tcgen05:
//   layer 1  the space-to-depth form of umma_conv1.cu: the 8x8/stride-4 convolution is four K = 16 MMAs per 128-row
//            tile, each reading a row-shifted window of the block image (which arrives by one bulk copy per sample);
//            its accumulators (TMEM) are drained by the producer warps, which apply LeakyReLU and deposit bf16 pixels in
//            the two-plane, two-pixels-per-128-byte-row swizzled slab that
//   layer 2  the window convolution of umma_window.cuh reads with shifted UMMA descriptors (same taps, same epilogue
//            code), one sample per 128-row tile.  No bias MMAs: the tensor pipe is the bottleneck, the biases are added
//            by the drain / epilogue warps (fp32, before the rounding to bf16).
// When the caller needs the intermediate activation (predictor in the loss pass: conv2's weight gradient reads it),
// the slab is also written out by TMA, two plane stores per sample.
//
// Measured (131 072 samples): the one-network kernel takes 1.45-1.55 ms (+ 0.54 ms per pass for the normalisation)
// against 0.78 + 0.98 ms for the separate kernels -- no gain: tcgen05.mma with both operands in shared memory costs
// ~(M + N) / 4 cycles per K = 16 step here (46 cycles at N = 32, 52 at N = 64, against 16 / 32 for the math), so these
// skinny layers are bound by the operand fetch from shared memory, not by the HBM round trip the fusion removes.  The
// TWIN kernel further down shares that fetch between target and predictor (one image, first layer as one N = 64
// problem): 2.5 ms for both networks against 3.5 ms, and is what RND runs; this one serves the predictor-only
// (feature reuse) path with identical numerics.  Earlier versions (layer 1 by mma.sync: 2.1 ms; uint8 conversion inside
// the kernel with three different producer arrangements: 1.5 ms each) and the ablations are in
// profiles/r2_fused_conv12_study.txt.
//
// Warp roles (576 threads): 0 = loader (conv2 weights by TMA, block images by bulk copy), 1 = TMEM owner + tcgen05
// issuer for both layers (and the TMA store of the slab), 2-5 = conv2 epilogue, 6-17 = first-layer drain (three per
// TMEM lane quarter).  Block images: ring of four; first-layer accumulators and slab stages: double-buffered.
#include "common.cuh"
#include "umma_window.cuh"
#include <type_traits>

namespace cb {
int encode_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle);
}

namespace {

using namespace umma;

constexpr int kH = 84, kPx = kH * kH;                 // 7056 pixels
constexpr int kBW = 21, kBRows = kBW * kBW;           // 4x4-pixel blocks per frame row / per frame
constexpr int kO1 = 20, kC1 = 32, kG = 10, kO2 = 9, kC2 = 64;
constexpr int kTaps = 8;
constexpr int kEpiGroups = 1;                         // conv2 epilogue warps per TMEM lane quarter (two 32-column chunks each)
constexpr int kProdWarp0 = kEpiWarp0 + 4 * kEpiGroups;
constexpr int kProdWarps = 12;                        // three per TMEM lane quarter
constexpr int kThreads = (kProdWarp0 + kProdWarps) * 32;
constexpr int kImgStages = 4;

constexpr uint32_t kW2Bytes = kTaps * kC2 * 128;      // 65536
constexpr uint32_t kW1Bytes = 4 * kC1 * 32;           // four taps of [32 channels][16 pixels]
constexpr uint32_t kPlaneBytes = 13312;               // 100 rows of 128 B, rounded up to the 1024-byte swizzle atom
constexpr uint32_t kStageBytes = 2 * kPlaneBytes;
constexpr uint32_t kImgBytes = kBRows * 32;           // 14112: one normalised frame in block order
constexpr uint32_t kS2dBytes = 14336;                 // its shared-memory stage (multiple of the 256-byte swizzle atom)
constexpr uint32_t kW1Off = kW2Bytes;
constexpr uint32_t kSlabOff = kW1Off + kW1Bytes;
constexpr uint32_t kS2dOff = kSlabOff + 2 * kStageBytes;      // shifted windows of the last slab plane / block image read
constexpr uint32_t kPadOff = kS2dOff + kImgStages * kS2dBytes;   // on into whatever follows: rows that are never stored
constexpr uint32_t kBarOff = (kPadOff + 4096u + 1023u) & ~1023u;  // (4 KB of zeros after the last image for its windows)
constexpr uint32_t kSmemBytes = kBarOff + 256u + (kC2 + kC1) * 4u + 1024u /*alignment*/;     // barriers, staged biases
static_assert(kSmemBytes <= 227u * 1024u, "shared memory");
static_assert(kSlabOff % 1024 == 0 && kS2dOff % 1024 == 0 && kImgBytes % 16 == 0, "alignment");

struct Rnd12Params {
    const uint8_t* img; int n;                        // normalised block images, kImgBytes per sample (cb_rnd_normalize_s2d)
    const __nv_bfloat16* w1; const float* b1;         // conv1 weights [32][64] in torch's (kh, kw) order, bias [32]
    int store_a1;
    GemmParams epi;                                   // conv2 epilogue (bias, activation, out_bf16)
};

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ uint32_t swz32(uint32_t off) { return off ^ (((off >> 7) & 1u) << 4); }
// SWIZZLE_32B descriptor: rows of 32 B, 8-row atoms of 256 B
__device__ __forceinline__ uint64_t desc32(uint32_t saddr) {
    return (make_smem_desc(saddr, 16, 256) & ~((uint64_t)7 << 61)) | ((uint64_t)6 << 61);
}

__global__ void __launch_bounds__(kThreads, 1)
umma_rnd12_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA1, const Rnd12Params p) {
    constexpr int BN = kC2;
    // TMEM columns: [0,128) second-layer accumulators (two stages of 64), [128,384) first-layer accumulators (two
    // stages x four row tiles x 32 channels)
    constexpr uint32_t TMEM_COLS = 512, kAcc1Col = 128;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + kBarOff;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };           // slab stage written            (drain -> MMA)
    auto empty_bar = [&](int s) { return bar_base + 8u * (2 + s); };    // slab stage consumed           (MMA -> drain)
    auto tfull_bar = [&](int a) { return bar_base + 8u * (4 + a); };    // conv2 accumulator ready       (MMA -> epilogue)
    auto tempty_bar = [&](int a) { return bar_base + 8u * (6 + a); };   // conv2 accumulator drained     (epilogue -> MMA)
    auto c1full_bar = [&](int s) { return bar_base + 8u * (8 + s); };   // conv1 accumulators ready      (MMA -> drain)
    auto c1empty_bar = [&](int s) { return bar_base + 8u * (10 + s); }; // conv1 accumulators drained    (drain -> MMA)
    auto ifull_bar = [&](int s) { return bar_base + 8u * (12 + s); };   // block image landed            (bulk copy -> MMA)
    auto iempty_bar = [&](int s) { return bar_base + 8u * (12 + kImgStages + s); };  // block image consumed (MMA -> loader)
    const uint32_t w_bar = bar_base + 8u * (12 + 2 * kImgStages);
    const uint32_t tmem_slot = bar_base + 8u * (13 + 2 * kImgStages);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + kBarOff + 8u * (13 + 2 * kImgStages));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* bias2 = stage_bias(p.epi, gen + kBarOff + 256u);
    // first-layer weights: per tap (bh, bw) a K-major [32 channels][16 = (py, px)] block, SWIZZLE_32B
    for (int i = threadIdx.x; i < 4 * kC1 * 16; i += blockDim.x) {
        const int e = i & 15, n = (i >> 4) & 31, t = i >> 9;
        const int kh = 4 * (t >> 1) + (e >> 2), kw = 4 * (t & 1) + (e & 3);
        *reinterpret_cast<__nv_bfloat16*>(gen + kW1Off + swz32((uint32_t)((t * kC1 + n) * 32 + e * 2))) = p.w1[n * 64 + kh * 8 + kw];
    }
    if (threadIdx.x < kC1) reinterpret_cast<float*>(gen + kBarOff + 256u + kC2 * 4u)[threadIdx.x] = __ldg(p.b1 + threadIdx.x);
    // slab padding rows, block-image tails and everything a shifted window may touch start out as zeros
    {
        uint4* z = reinterpret_cast<uint4*>(gen + kSlabOff);
        for (uint32_t i = threadIdx.x; i < (kBarOff - kSlabOff) / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmW);
        if (p.store_a1) prefetch_tmap(&tmA1);
        for (int s = 0; s < 2; ++s) {
            // drain warps arrive one by one (no CTA-wide barrier keeps them in lock-step); a slab stage is free again
            // when the second-layer MMAs AND (if it is written out) the TMA store have finished reading it
            mbar_init(full_bar(s), kProdWarps); mbar_init(empty_bar(s), p.store_a1 ? 2 : 1);
            mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4 * kEpiGroups);
            mbar_init(c1full_bar(s), 1); mbar_init(c1empty_bar(s), kProdWarps);
        }
        for (int s = 0; s < kImgStages; ++s) { mbar_init(ifull_bar(s), 1); mbar_init(iempty_bar(s), 1); }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int my_tiles = (p.n - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // this CTA's samples

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(w_bar, kW2Bytes);
            for (int kb = 0; kb < kTaps; ++kb) tma_load_2d(smem_base + kb * BN * 128, &tmW, w_bar, kb * 64, 0);
            int slot = 0; uint32_t ph = 0;
            for (int j = 0; j < my_tiles; ++j) {
                const long long tile = blockIdx.x + (long long)j * gridDim.x;
                mbar_wait(iempty_bar(slot), ph ^ 1u);
                mbar_arrive_expect_tx(ifull_bar(slot), kImgBytes);
                bulk_load(smem_base + kS2dOff + slot * kS2dBytes, p.img + tile * kImgBytes, kImgBytes, ifull_bar(slot));
                if (++slot == kImgStages) { slot = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();              // all lanes loop, one issues (see umma_gemm_kernel)
        constexpr uint32_t idesc2 = make_idesc(128, BN, 0, 0), idesc1 = make_idesc(128, kC1, 0, 0);
        mbar_wait(w_bar, 0);
        const uint64_t db2 = make_smem_desc(smem_base, 16, 1024);
        const uint64_t db1 = desc32(smem_base + kW1Off);
        // first layer of sample j is issued before the second layer of sample j-1: the producers drain its
        // accumulators while the (longer) second-layer MMAs of sample j-1 run
        int islot = 0; uint32_t iph = 0;
        for (int j = 0; j <= my_tiles; ++j) {
            if (j < my_tiles) {
                const int g = j & 1;
                const uint32_t ph = ((uint32_t)j >> 1) & 1u;
                mbar_wait(c1empty_bar(g), ph ^ 1u);
                mbar_wait(ifull_bar(islot), iph);
                fence_after_thread_sync();
                const uint64_t da0 = desc32(smem_base + kS2dOff + islot * kS2dBytes);
                if (leader) {
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        const uint32_t d = tmem_base + kAcc1Col + (uint32_t)((g * 4 + mt) * kC1);
#pragma unroll
                        for (int t = 0; t < 4; ++t)      // tap (bh, bw): the block image shifted by bh*21 + bw rows
                            mma_f16_ss(d, da0 + (uint64_t)(((mt * 128 + (t >> 1) * kBW + (t & 1)) * 32) >> 4),
                                       db1 + (uint64_t)((t * kC1 * 32) >> 4), idesc1, t ? 1u : 0u);
                    }
                    mma_commit(iempty_bar(islot));
                    mma_commit(c1full_bar(g));
                }
                __syncwarp();
                if (++islot == kImgStages) { islot = 0; iph ^= 1u; }
            }
            if (j >= 1) {
                const int jj = j - 1, st = jj & 1;    // slab stage == accumulator stage == producer group
                const uint32_t ph = ((uint32_t)jj >> 1) & 1u;
                mbar_wait(tempty_bar(st), ph ^ 1u);
                mbar_wait(full_bar(st), ph);
                fence_after_thread_sync();
                const uint64_t da0 = make_smem_desc(smem_base + kSlabOff + st * kStageBytes, 16, 1024);
                const uint32_t d_tmem = tmem_base + (uint32_t)(st * BN);
                if (leader) {
                    if (p.store_a1) {                 // the slab is complete: write the first-layer activation out, plane by plane
                        const long long tile = blockIdx.x + (long long)jj * gridDim.x;
                        const uint32_t slab_s = smem_base + kSlabOff + st * kStageBytes;
                        tma_store_5d(&tmA1, slab_s, 0, 0, 0, 0, (int)tile);
                        tma_store_5d(&tmA1, slab_s + kPlaneBytes, 0, 0, 1, 0, (int)tile);
                        bulk_commit();
                    }
#pragma unroll
                    for (int kb = 0; kb < kTaps; ++kb) {
                        // tap (kh, kw pair j2): plane kh & 1, shifted by (kh >> 1) grid rows and j2 grid columns
                        const uint32_t off16 = ((uint32_t)((kb >> 1) & 1) * kPlaneBytes + (uint32_t)((kb >> 2) * kG + (kb & 1)) * 128u) >> 4;
                        const uint64_t da = da0 + (uint64_t)off16;
                        const uint64_t db = db2 + (uint64_t)(kb * BN * 8);
#pragma unroll
                        for (int k = 0; k < 4; ++k) mma_f16_ss(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc2, (kb | k) ? 1u : 0u);
                    }
                    mma_commit(empty_bar(st));
                    mma_commit(tfull_bar(st));
                    if (p.store_a1) {                 // (the queued MMAs keep the tensor pipe busy meanwhile)
                        bulk_wait_read0();
                        mbar_arrive(empty_bar(st));
                    }
                }
                __syncwarp();
            }
        }
        if (leader && p.store_a1) bulk_wait0();
        __syncwarp();
    } else if (warp < kProdWarp0) {
        const int quarter = warp & 3;
        const int sub = (warp - kEpiWarp0) >> 2;
        const int r = quarter * 32 + lane;
        const int gy = r / kG, gx = r - gy * kG;
        const bool ok = r < kG * kG && gy < kO2 && gx < kO2;
        const long long rel_row = (long long)gy * kO2 + gx;
        for (int j = 0; j < my_tiles; ++j) {
            const long long tile = blockIdx.x + (long long)j * gridDim.x;
            const int acc = j & 1;
            epilogue_tile<BN, 1, EPI_SIMPLE, kEpiGroups, false>(
                p.epi, bias2, tmem_base + (uint32_t)(acc * BN), quarter, sub, lane, 0, tfull_bar(acc), ((uint32_t)j >> 1) & 1u,
                [&](int, bool& valid, long long& grow, long long& base) {
                    valid = ok;
                    grow = tile * (kO2 * kO2) + rel_row;
                    base = grow * BN;
                });
            fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
    } else {
        // ---- first-layer drain (twelve warps, three per TMEM lane quarter, each at its own pace: every hand-over is
        // an mbarrier with one arrival per warp): accumulators of sample j -> LeakyReLU -> bf16 -> slab stage j&1 in
        // the layout the second layer's windows read
        const int quarter = warp & 3, sub = (warp - kProdWarp0) >> 2;
        const __nv_bfloat162 slope = __floats2bfloat162_rn(0.01f, 0.01f);
        // the row tiles this warp drains: sub-warp s of a quarter takes tile s, sub-warp 0 also tile 3 (only the first
        // two quarters of it hold block rows of the frame); and where this lane's accumulator row goes in the slab:
        // block (by, bx) = output pixel, or none
        int srow[2]; uint32_t sx[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int mt = k ? 3 : sub;
            const int r = mt * 128 + quarter * 32 + lane;
            const int by = r / kBW, bx = r - by * kBW;
            const bool ok = r < kBRows && by < kO1 && bx < kO1 && (k == 0 || sub == 0);
            const uint32_t row = (uint32_t)(by & 1) * kPlaneBytes + (uint32_t)((by >> 1) * kG + (bx >> 1)) * 128u;
            srow[k] = ok ? (int)row : -1;
            sx[k] = ((row >> 7) & 7u) ^ ((uint32_t)(bx & 1) * 4u);        // 16-byte chunk c of the pixel goes to chunk c ^ sx
        }
        const int nblk = (sub == 0 && 3 * 128 + quarter * 32 < kBRows) ? 2 : 1;
        for (int j = 0; j < my_tiles; ++j) {
            const int b = j & 1;
            const uint32_t ph = ((uint32_t)j >> 1) & 1u;
            uint8_t* slab = gen + kSlabOff + b * kStageBytes;
            mbar_wait(empty_bar(b), ph ^ 1u);     // second-layer MMAs (and the store) of the sample two back are done with this slab stage
            mbar_wait(c1full_bar(b), ph);
            fence_after_thread_sync();
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (k >= nblk) continue;          // (warp-uniform)
                const int mt = k ? 3 : sub;
                uint32_t acc[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + kAcc1Col + (uint32_t)((b * 4 + mt) * kC1), acc);
                tmem_ld_wait();
                if (srow[k] < 0) continue;
                // accumulator = conv + bias; round to bf16, LeakyReLU on packed pairs (as the stand-alone kernel),
                // then the pixel's 64 bytes as four 16-byte chunks of its swizzled 128-byte slab row
                uint32_t pk[16];
                const float4* bsrc = reinterpret_cast<const float4*>(gen + kBarOff + 256u + kC2 * 4u);     // staged first-layer bias
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 b4 = bsrc[q];
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(acc[4 * q]) + b4.x, __uint_as_float(acc[4 * q + 1]) + b4.y);
                    __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(acc[4 * q + 2]) + b4.z, __uint_as_float(acc[4 * q + 3]) + b4.w);
                    h0 = __hmax2(h0, __hmul2(h0, slope));
                    h1 = __hmax2(h1, __hmul2(h1, slope));
                    pk[2 * q] = *reinterpret_cast<uint32_t*>(&h0); pk[2 * q + 1] = *reinterpret_cast<uint32_t*>(&h1);
                }
                uint8_t* dst = slab + srow[k];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4*>(dst + (((uint32_t)c ^ sx[k]) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            }
            fence_before_thread_sync();
            fence_proxy_async();                  // generic-proxy slab writes -> visible to tcgen05 / TMA
            __syncwarp();
            if (lane == 0) { mbar_arrive(c1empty_bar(b)); mbar_arrive(full_bar(b)); }
        }
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) { fence_after_thread_sync(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ================================================================================================================
// Twin variant: target AND predictor of one sample in the same CTA.  The block image is loaded once, the first layer
// is ONE set of N = 64 MMAs for both networks (the operand fetch of the 128 image rows, which is what these MMAs
// cost, is shared), the two second layers alternate on the tensor pipe so that the drain of one network's first-layer
// accumulators always runs under the other network's second-layer MMAs:
//     pipe:   conv1(j) | conv2_p(j-1) | conv2_t(j) | conv1(j+1) | conv2_p(j) | conv2_t(j+1) | ...
//     drain:            drain_t(j)     drain_p(j)                 drain_t(j+1) ...
// Shared memory: both second-layer weight sets resident (128 KB), twin first-layer weights (8 KB), one slab per
// network, two block images.  No bias MMAs: the biases are added by the drain / epilogue warps.
// ================================================================================================================
constexpr uint32_t kTwW1Off = 2 * kW2Bytes;
constexpr uint32_t kTwW1Bytes = 4 * 2 * kC1 * 32;
constexpr uint32_t kTwSlabOff = kTwW1Off + kTwW1Bytes;            // slab of network 0, then of network 1
constexpr uint32_t kTwImgOff = kTwSlabOff + 2 * kStageBytes;
constexpr int kTwImgStages = 2;
constexpr uint32_t kTwBarOff = kTwImgOff + kTwImgStages * kS2dBytes;
constexpr uint32_t kTwBiasOff = kTwBarOff + 256u;                 // bias2[2][64], bias1[2][32] (fp32)
// (the first layer's shifted M = 128 windows of the LAST image read up to block row 4*128 + 22: the allocation covers it)
constexpr uint32_t kTwWindowEnd = kTwImgOff + (kTwImgStages - 1) * kS2dBytes + (4 * 128 + kBW + 2) * 32u;
constexpr uint32_t kTwDataEnd = kTwBiasOff + (2 * kC2 + 2 * kC1) * 4u;
constexpr uint32_t kTwSmemBytes = (kTwWindowEnd > kTwDataEnd ? kTwWindowEnd : kTwDataEnd) + 1024u /*alignment*/;
static_assert(kTwSmemBytes <= 227u * 1024u, "shared memory");
static_assert(kTwSlabOff % 1024 == 0 && kTwImgOff % 1024 == 0 && kTwBarOff % 8 == 0, "alignment");

struct Rnd12TwinParams {
    const uint8_t* img; int n;
    const __nv_bfloat16* w1[2]; const float* b1[2];   // network 0 = target, 1 = predictor
    int store_a1;                                     // write the predictor's first-layer activation out
    GemmParams epi[2];
};

__global__ void __launch_bounds__(kThreads, 1)
umma_rnd12_twin_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
                       const __grid_constant__ CUtensorMap tmA1, const Rnd12TwinParams p) {
    constexpr int BN = kC2;
    // TMEM columns: [0,64) / [64,128) second-layer accumulators of network 0 / 1, [128,384) first-layer accumulators
    // (four row tiles x [32 channels of network 0 | 32 of network 1])
    constexpr uint32_t TMEM_COLS = 512, kAcc1Col = 128;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + kTwBarOff;
    auto full_bar = [&](int net) { return bar_base + 8u * net; };          // slab written           (drain -> MMA)
    auto empty_bar = [&](int net) { return bar_base + 8u * (2 + net); };   // slab consumed          (MMA -> drain)
    auto tfull_bar = [&](int net) { return bar_base + 8u * (4 + net); };   // conv2 accumulator      (MMA -> epilogue)
    auto tempty_bar = [&](int net) { return bar_base + 8u * (6 + net); };  //                        (epilogue -> MMA)
    const uint32_t c1full_bar = bar_base + 8u * 8, c1empty_bar = bar_base + 8u * 9;
    auto ifull_bar = [&](int s) { return bar_base + 8u * (10 + s); };
    auto iempty_bar = [&](int s) { return bar_base + 8u * (12 + s); };
    const uint32_t w_bar = bar_base + 8u * 14;
    const uint32_t tmem_slot = bar_base + 8u * 15;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + kTwBarOff + 8u * 15);
    float* s_bias2 = reinterpret_cast<float*>(gen + kTwBiasOff);           // [2][64]
    float* s_bias1 = s_bias2 + 2 * kC2;                                    // [2][32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 2 * kC2) s_bias2[threadIdx.x] = __ldg(p.epi[threadIdx.x >> 6].bias + (threadIdx.x & 63));
    if (threadIdx.x < 2 * kC1) s_bias1[threadIdx.x] = __ldg(p.b1[threadIdx.x >> 5] + (threadIdx.x & 31));
    // first-layer weights: per tap a K-major [64 = network x channel][16 = (py, px)] block, SWIZZLE_32B
    for (int i = threadIdx.x; i < 4 * 2 * kC1 * 16; i += blockDim.x) {
        const int e = i & 15, n = (i >> 4) & 63, t = i >> 10;
        const int kh = 4 * (t >> 1) + (e >> 2), kw = 4 * (t & 1) + (e & 3);
        *reinterpret_cast<__nv_bfloat16*>(gen + kTwW1Off + swz32((uint32_t)((t * 2 * kC1 + n) * 32 + e * 2))) =
            p.w1[n >> 5][(n & 31) * 64 + kh * 8 + kw];
    }
    {   // slabs (padding rows) and images start out as zeros
        uint4* z = reinterpret_cast<uint4*>(gen + kTwSlabOff);
        for (uint32_t i = threadIdx.x; i < (kTwBarOff - kTwSlabOff) / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmW0); prefetch_tmap(&tmW1);
        if (p.store_a1) prefetch_tmap(&tmA1);
        for (int net = 0; net < 2; ++net) {
            mbar_init(full_bar(net), kProdWarps);
            mbar_init(empty_bar(net), (net == 1 && p.store_a1) ? 2 : 1);
            mbar_init(tfull_bar(net), 1); mbar_init(tempty_bar(net), 4 * kEpiGroups);
        }
        mbar_init(c1full_bar, 1); mbar_init(c1empty_bar, kProdWarps);
        for (int s = 0; s < kTwImgStages; ++s) { mbar_init(ifull_bar(s), 1); mbar_init(iempty_bar(s), 1); }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int my_tiles = (p.n - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(w_bar, 2 * kW2Bytes);
            for (int kb = 0; kb < kTaps; ++kb) {
                tma_load_2d(smem_base + kb * BN * 128, &tmW0, w_bar, kb * 64, 0);
                tma_load_2d(smem_base + kW2Bytes + kb * BN * 128, &tmW1, w_bar, kb * 64, 0);
            }
            for (int j = 0; j < my_tiles; ++j) {
                const long long tile = blockIdx.x + (long long)j * gridDim.x;
                const int slot = j & 1;
                mbar_wait(iempty_bar(slot), (((uint32_t)j >> 1) & 1u) ^ 1u);
                mbar_arrive_expect_tx(ifull_bar(slot), kImgBytes);
                bulk_load(smem_base + kTwImgOff + slot * kS2dBytes, p.img + tile * kImgBytes, kImgBytes, ifull_bar(slot));
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        constexpr uint32_t idesc2 = make_idesc(128, BN, 0, 0), idesc1 = make_idesc(128, 2 * kC1, 0, 0);
        mbar_wait(w_bar, 0);
        const uint64_t db1 = desc32(smem_base + kTwW1Off);
        auto conv2 = [&](int net, int jj) {           // second layer of network `net` for this CTA's sample jj
            const uint32_t ph = (uint32_t)jj & 1u;
            mbar_wait(tempty_bar(net), ph ^ 1u);
            mbar_wait(full_bar(net), ph);
            fence_after_thread_sync();
            const uint32_t slab_s = smem_base + kTwSlabOff + net * kStageBytes;
            const uint64_t da0 = make_smem_desc(slab_s, 16, 1024);
            const uint64_t db0 = make_smem_desc(smem_base + net * kW2Bytes, 16, 1024);
            const uint32_t d_tmem = tmem_base + (uint32_t)(net * BN);
            if (leader) {
                const bool store = net == 1 && p.store_a1;
                if (store) {                          // the slab is complete: write the first-layer activation out, plane by plane
                    const long long tile = blockIdx.x + (long long)jj * gridDim.x;
                    tma_store_5d(&tmA1, slab_s, 0, 0, 0, 0, (int)tile);
                    tma_store_5d(&tmA1, slab_s + kPlaneBytes, 0, 0, 1, 0, (int)tile);
                    bulk_commit();
                }
#pragma unroll
                for (int kb = 0; kb < kTaps; ++kb) {
                    const uint32_t off16 = ((uint32_t)((kb >> 1) & 1) * kPlaneBytes + (uint32_t)((kb >> 2) * kG + (kb & 1)) * 128u) >> 4;
                    const uint64_t da = da0 + (uint64_t)off16;
                    const uint64_t db = db0 + (uint64_t)(kb * BN * 8);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_f16_ss(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc2, (kb | k) ? 1u : 0u);
                }
                mma_commit(empty_bar(net));
                mma_commit(tfull_bar(net));
                if (store) { bulk_wait_read0(); mbar_arrive(empty_bar(net)); }
            }
            __syncwarp();
        };
        for (int j = 0; j <= my_tiles; ++j) {
            if (j < my_tiles) {
                const int slot = j & 1;
                mbar_wait(c1empty_bar, ((uint32_t)j & 1u) ^ 1u);
                mbar_wait(ifull_bar(slot), ((uint32_t)j >> 1) & 1u);
                fence_after_thread_sync();
                const uint64_t da0 = desc32(smem_base + kTwImgOff + slot * kS2dBytes);
                if (leader) {
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        const uint32_t d = tmem_base + kAcc1Col + (uint32_t)(mt * 2 * kC1);
#pragma unroll
                        for (int t = 0; t < 4; ++t)      // tap (bh, bw): the block image shifted by bh*21 + bw rows
                            mma_f16_ss(d, da0 + (uint64_t)(((mt * 128 + (t >> 1) * kBW + (t & 1)) * 32) >> 4),
                                       db1 + (uint64_t)((t * 2 * kC1 * 32) >> 4), idesc1, t ? 1u : 0u);
                    }
                    mma_commit(iempty_bar(slot));
                    mma_commit(c1full_bar);
                }
                __syncwarp();
            }
            if (j >= 1) conv2(1, j - 1);
            if (j < my_tiles) conv2(0, j);
        }
        if (leader && p.store_a1) bulk_wait0();
        __syncwarp();
    } else if (warp < kProdWarp0) {
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const int gy = r / kG, gx = r - gy * kG;
        const bool ok = r < kG * kG && gy < kO2 && gx < kO2;
        const long long rel_row = (long long)gy * kO2 + gx;
        auto epilogue = [&](auto NET, int jj) {       // (the network index is a compile-time constant: p.epi[net] stays a
            constexpr int net = decltype(NET)::value; //  constant-bank operand instead of a local-memory copy of the struct)
            const long long tile = blockIdx.x + (long long)jj * gridDim.x;
            epilogue_tile<BN, 1, EPI_SIMPLE, kEpiGroups, false>(
                p.epi[net], s_bias2 + net * kC2, tmem_base + (uint32_t)(net * BN), quarter, 0, lane, 0, tfull_bar(net),
                (uint32_t)jj & 1u, [&](int, bool& valid, long long& grow, long long& base) {
                    valid = ok;
                    grow = tile * (kO2 * kO2) + rel_row;
                    base = grow * BN;
                });
            fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(net));
        };
        for (int j = 0; j <= my_tiles; ++j) {
            if (j >= 1) epilogue(std::integral_constant<int, 1>{}, j - 1);
            if (j < my_tiles) epilogue(std::integral_constant<int, 0>{}, j);
        }
    } else {
        // ---- first-layer drain (twelve warps, three per TMEM lane quarter): target half, then predictor half
        const int quarter = warp & 3, sub = (warp - kProdWarp0) >> 2;
        const __nv_bfloat162 slope = __floats2bfloat162_rn(0.01f, 0.01f);
        int srow[2]; uint32_t sx[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int mt = k ? 3 : sub;
            const int r = mt * 128 + quarter * 32 + lane;
            const int by = r / kBW, bx = r - by * kBW;
            const bool ok = r < kBRows && by < kO1 && bx < kO1 && (k == 0 || sub == 0);
            const uint32_t row = (uint32_t)(by & 1) * kPlaneBytes + (uint32_t)((by >> 1) * kG + (bx >> 1)) * 128u;
            srow[k] = ok ? (int)row : -1;
            sx[k] = ((row >> 7) & 7u) ^ ((uint32_t)(bx & 1) * 4u);
        }
        const int nblk = (sub == 0 && 3 * 128 + quarter * 32 < kBRows) ? 2 : 1;
        for (int j = 0; j < my_tiles; ++j) {
            const uint32_t ph = (uint32_t)j & 1u;
            mbar_wait(c1full_bar, ph);
            fence_after_thread_sync();
#pragma unroll
            for (int net = 0; net < 2; ++net) {
                uint8_t* slab = gen + kTwSlabOff + net * kStageBytes;
                const float4* bsrc = reinterpret_cast<const float4*>(s_bias1 + net * kC1);
                mbar_wait(empty_bar(net), ph ^ 1u);   // this network's second-layer MMAs (and the store) of the previous sample are done
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (k >= nblk) continue;          // (warp-uniform)
                    const int mt = k ? 3 : sub;
                    uint32_t acc[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + kAcc1Col + (uint32_t)(mt * 2 * kC1 + net * kC1), acc);
                    tmem_ld_wait();
                    if (srow[k] < 0) continue;
                    uint32_t pk[16];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b4 = bsrc[q];
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(acc[4 * q]) + b4.x, __uint_as_float(acc[4 * q + 1]) + b4.y);
                        __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(acc[4 * q + 2]) + b4.z, __uint_as_float(acc[4 * q + 3]) + b4.w);
                        h0 = __hmax2(h0, __hmul2(h0, slope));
                        h1 = __hmax2(h1, __hmul2(h1, slope));
                        pk[2 * q] = *reinterpret_cast<uint32_t*>(&h0); pk[2 * q + 1] = *reinterpret_cast<uint32_t*>(&h1);
                    }
                    uint8_t* dst = slab + srow[k];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(dst + (((uint32_t)c ^ sx[k]) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                }
                fence_before_thread_sync();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { mbar_arrive(full_bar(net)); if (net == 1) mbar_arrive(c1empty_bar); }
            }
        }
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) { fence_after_thread_sync(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}


// uint8 frame -> normalised bf16 block image.  One thread per block row (4x4 pixels): four 4-byte pixel loads
// (coalesced along the frame row), the per-pixel statistics through L1, one 32-byte store in the address-swizzled
// order the first layer's UMMA descriptors read after a plain bulk copy into shared memory.
__global__ void __launch_bounds__(256) rnd_normalize_s2d_kernel(const uint8_t* __restrict__ x, long long stride_n, long long total_rows,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 uint8_t* __restrict__ out) {
    const __nv_bfloat162 five = __floats2bfloat162_rn(5.f, 5.f), nfive = __floats2bfloat162_rn(-5.f, -5.f);
    for (long long t = blockIdx.x * 256ll + threadIdx.x; t < total_rows; t += gridDim.x * 256ll) {
        const long long sidx = t / kBRows;
        const int row = (int)(t - sidx * kBRows);
        const int by = row / kBW, bx = row - by * kBW;
        const int pix0 = (4 * by) * kH + 4 * bx;
        const uint8_t* src = x + sidx * stride_n + pix0;
        uint32_t o[8];
#pragma unroll
        for (int py = 0; py < 4; ++py) {
            const uint32_t px = __ldg(reinterpret_cast<const uint32_t*>(src + py * kH));
            const float4 m = __ldg(reinterpret_cast<const float4*>(mean + pix0 + py * kH));
            const float4 r = __ldg(reinterpret_cast<const float4*>(rstd + pix0 + py * kH));
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (float)((px >> (8 * e)) & 0xffu);
            __nv_bfloat162 lo = __floats2bfloat162_rn((v[0] - m.x) * r.x, (v[1] - m.y) * r.y);
            __nv_bfloat162 hi = __floats2bfloat162_rn((v[2] - m.z) * r.z, (v[3] - m.w) * r.w);
            // clamp after the rounding: 5 is a bf16 value and rounding is monotonic, so the result is the same
            lo = __hmin2(__hmax2(lo, nfive), five);
            hi = __hmin2(__hmax2(hi, nfive), five);
            o[2 * py] = *reinterpret_cast<uint32_t*>(&lo); o[2 * py + 1] = *reinterpret_cast<uint32_t*>(&hi);
        }
        uint4* dst = reinterpret_cast<uint4*>(out + sidx * kImgBytes + (long long)row * 32);
        const int sw = (row >> 2) & 1;            // SWIZZLE_32B: the two 16-byte halves of rows 4..7 of every 8 swap
        dst[sw] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[sw ^ 1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

}  // namespace

extern "C" int cb_rnd_conv12_supported(void) { return cb_device_supports_umma(); }

extern "C" int cb_rnd_normalize_s2d(const uint8_t* x, int64_t stride_n, int32_t n, const float* mean, const float* rstd,
                                    void* out_s2d, void* stream) {
    CB_REQUIRE(x && mean && rstd && out_s2d, CB_E_ARG, "cb_rnd_normalize_s2d: null pointer");
    CB_REQUIRE(n > 0, CB_E_ARG, "cb_rnd_normalize_s2d: n = %d", n);
    CB_REQUIRE(((uintptr_t)x & 3) == 0 && stride_n % 4 == 0 && (((uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)out_s2d) & 15) == 0,
               CB_E_ALIGN, "cb_rnd_normalize_s2d: frames must be 4-byte, statistics and output 16-byte aligned");
    const long long rows = (long long)n * kBRows;
    long long blocks = (rows + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    rnd_normalize_s2d_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, stride_n, rows, mean, rstd, (uint8_t*)out_s2d);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_rnd_conv12_fwd(const void* img_s2d, int32_t n, const void* w1_bf16, const float* b1, const void* w2_bf16,
                                 const float* b2, void* a1_out_bf16, void* a2_out_bf16, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CB_REQUIRE(img_s2d && w1_bf16 && b1 && w2_bf16 && b2 && a2_out_bf16, CB_E_ARG, "cb_rnd_conv12_fwd: null pointer");
    CB_REQUIRE(n > 0, CB_E_ARG, "cb_rnd_conv12_fwd: n = %d", n);
    CB_REQUIRE(((uintptr_t)img_s2d & 15) == 0, CB_E_ALIGN, "cb_rnd_conv12_fwd: block images must be 16-byte aligned");
    CB_REQUIRE(cb_device_supports_umma() == 1, CB_E_ARCH, "cb_rnd_conv12_fwd needs an sm_100 device");
    CUtensorMap tmW, tmA1;
    {
        uint64_t dims[2] = {(uint64_t)(16 * kC1), (uint64_t)kC2};
        uint64_t strides[1] = {(uint64_t)(16 * kC1) * 2};
        uint32_t box[2] = {64, (uint32_t)kC2};
        int rc = cb::encode_tmap_bf16(&tmW, w2_bf16, 2, dims, strides, box, 3); if (rc) return rc;
    }
    if (a1_out_bf16) {      // [n][20][20][32] seen as [n][10][2][10][64]: one plane of one sample per store
        uint64_t dims[5] = {64, (uint64_t)kG, 2, (uint64_t)kG, (uint64_t)n};
        uint64_t strides[4] = {128, (uint64_t)kO1 * 64, (uint64_t)kO1 * 128, (uint64_t)kO1 * kO1 * 64};
        uint32_t box[5] = {64, (uint32_t)kG, 1, (uint32_t)kG, 1};
        int rc = cb::encode_tmap_bf16(&tmA1, a1_out_bf16, 5, dims, strides, box, 3); if (rc) return rc;
    } else {
        tmA1 = tmW;
    }
    Rnd12Params p{};
    p.img = (const uint8_t*)img_s2d; p.n = n;
    p.w1 = (const __nv_bfloat16*)w1_bf16; p.b1 = b1; p.store_a1 = a1_out_bf16 ? 1 : 0;
    p.epi.n_total = kC2; p.epi.bias = b2; p.epi.bias_in_acc = 0; p.epi.act = CB_ACT_LRELU;
    p.epi.out_bf16 = (__nv_bfloat16*)a2_out_bf16;
    static bool attr = false;
    if (!attr) {
        CB_CUDA(cudaFuncSetAttribute(umma_rnd12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        attr = true;
    }
    const int grid = n < sm_count() ? n : sm_count();
    umma_rnd12_kernel<<<grid, kThreads, kSmemBytes, s>>>(tmW, tmA1, p);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_rnd_conv12_twin_fwd(const void* img_s2d, int32_t n, const void* w1_t, const float* b1_t, const void* w2_t,
                                      const float* b2_t, const void* w1_p, const float* b1_p, const void* w2_p,
                                      const float* b2_p, void* a2_t_out, void* a1_p_out, void* a2_p_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CB_REQUIRE(img_s2d && w1_t && b1_t && w2_t && b2_t && w1_p && b1_p && w2_p && b2_p && a2_t_out && a2_p_out, CB_E_ARG,
               "cb_rnd_conv12_twin_fwd: null pointer");
    CB_REQUIRE(n > 0, CB_E_ARG, "cb_rnd_conv12_twin_fwd: n = %d", n);
    CB_REQUIRE(((uintptr_t)img_s2d & 15) == 0, CB_E_ALIGN, "cb_rnd_conv12_twin_fwd: block images must be 16-byte aligned");
    CB_REQUIRE(cb_device_supports_umma() == 1, CB_E_ARCH, "cb_rnd_conv12_twin_fwd needs an sm_100 device");
    CUtensorMap tmW[2], tmA1;
    const void* w2[2] = {w2_t, w2_p};
    for (int net = 0; net < 2; ++net) {
        uint64_t dims[2] = {(uint64_t)(16 * kC1), (uint64_t)kC2};
        uint64_t strides[1] = {(uint64_t)(16 * kC1) * 2};
        uint32_t box[2] = {64, (uint32_t)kC2};
        int rc = cb::encode_tmap_bf16(&tmW[net], w2[net], 2, dims, strides, box, 3); if (rc) return rc;
    }
    if (a1_p_out) {
        uint64_t dims[5] = {64, (uint64_t)kG, 2, (uint64_t)kG, (uint64_t)n};
        uint64_t strides[4] = {128, (uint64_t)kO1 * 64, (uint64_t)kO1 * 128, (uint64_t)kO1 * kO1 * 64};
        uint32_t box[5] = {64, (uint32_t)kG, 1, (uint32_t)kG, 1};
        int rc = cb::encode_tmap_bf16(&tmA1, a1_p_out, 5, dims, strides, box, 3); if (rc) return rc;
    } else {
        tmA1 = tmW[0];
    }
    Rnd12TwinParams p{};
    p.img = (const uint8_t*)img_s2d; p.n = n;
    p.w1[0] = (const __nv_bfloat16*)w1_t; p.w1[1] = (const __nv_bfloat16*)w1_p; p.b1[0] = b1_t; p.b1[1] = b1_p;
    p.store_a1 = a1_p_out ? 1 : 0;
    const float* b2[2] = {b2_t, b2_p};
    void* out[2] = {a2_t_out, a2_p_out};
    for (int net = 0; net < 2; ++net) {
        p.epi[net].n_total = kC2; p.epi[net].bias = b2[net]; p.epi[net].bias_in_acc = 0; p.epi[net].act = CB_ACT_LRELU;
        p.epi[net].out_bf16 = (__nv_bfloat16*)out[net];
    }
    static bool attr = false;
    if (!attr) {
        CB_CUDA(cudaFuncSetAttribute(umma_rnd12_twin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTwSmemBytes));
        attr = true;
    }
    const int grid = n < sm_count() ? n : sm_count();
    umma_rnd12_twin_kernel<<<grid, kThreads, kTwSmemBytes, s>>>(tmW[0], tmW[1], tmA1, p);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// Persistent GRU recurrence for NDIGO's belief state (rlpyt/models/curiosity/ndigo.py:79,127,145; torch.nn.GRU, gate
// order r,z,n) -- SURVEY 8f rank 3.
//
// The recurrence is independent across environment columns, so the batch is partitioned over CTAs by ROWS: a CTA
// owns 32 rows for all T steps and never talks to another CTA -- no grid barrier, one launch for the whole sequence
// instead of two launches per step (a GEMM whose epilogue had no next tile to overlap with, plus the cell kernel).
// W_hh (3G x G bf16, 96 KB for G = 128) is loaded ONCE into registers as tensor-core B fragments (96 registers per
// thread); the hidden state lives in registers (fp32) and, as the next step's bf16 A operand, in a double-buffered
// shared-memory tile; one __syncthreads per step.  The per-step product is 32 x 384 x 128: latency, not throughput,
// is what matters here, so it is issued as warp-level mma.sync (m16n8k16, bf16 operands, fp32 accumulation) whose
// accumulator fragments land in the registers of the threads that apply the cell -- tcgen05's TMEM round trip and
// single-issuer model buy nothing at this size.
//
// Backward (BPTT) mirrors it: dL/dh is carried between steps in fp32 registers (the two-launch version carried it in
// bf16), the cell gradient writes d gi / d gh (bf16 operands of the later weight-gradient GEMMs) and the 32 x 128 x 384
// product d gh W_hh feeds the next step.
#include "common.cuh"

namespace {

constexpr int kRows = 32;          // batch rows per CTA
constexpr int kG = 128;            // hidden units (ndigo.py:47 gru_size default)
constexpr int kThreads = 256;      // 8 warps: warp w owns hidden units [16w, 16w + 16)

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float sigmoid_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ---------------------------------------------------------------------------------------------------- forward
// gates fp32 [T,B,3G]: x W_ih^T + b_ih on entry, the activated (r,z,n) on exit.  h_all fp32 / h_bf bf16 [T+1,B,G] with
// row 0 = the initial state.  ghn fp32 [T,B,G] = h W_hn^T + b_hn (needed by the backward pass).
__global__ void __launch_bounds__(kThreads, 1)
gru_persistent_fwd_kernel(int T, int B, float* __restrict__ gates, const __nv_bfloat16* __restrict__ w_hh,
                          const float* __restrict__ b_hh, float* __restrict__ h_all, __nv_bfloat16* __restrict__ h_bf,
                          float* __restrict__ ghn_all) {
    constexpr int LD = kG + 8;                                  // padded row (bank-conflict-free fragment loads)
    __shared__ __align__(16) __nv_bfloat16 hs[2][kRows * LD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int row0 = blockIdx.x * kRows;
    const size_t bg = (size_t)B * kG;
    // W_hh as B fragments: [gate s][n-tile q][k-step][2]
    uint32_t wf[3][2][8][2];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const __nv_bfloat16* wrow = w_hh + (size_t)(s * kG + 16 * warp + 8 * q + g) * kG;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                wf[s][q][kk][0] = *reinterpret_cast<const uint32_t*>(wrow + 16 * kk + 2 * tq);
                wf[s][q][kk][1] = *reinterpret_cast<const uint32_t*>(wrow + 16 * kk + 8 + 2 * tq);
            }
        }
    float bias[3][2][2];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            bias[s][q][0] = __ldg(b_hh + s * kG + 16 * warp + 8 * q + 2 * tq);
            bias[s][q][1] = __ldg(b_hh + s * kG + 16 * warp + 8 * q + 2 * tq + 1);
        }
    // element (m, half, q, e): row = 16m + g + 8*half, hidden j = 16*warp + 8q + 2tq + e
    float hprev[2][2][2][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int r = row0 + 16 * m + g + 8 * hf;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int j = 16 * warp + 8 * q + 2 * tq;
                float2 v = make_float2(0.f, 0.f);
                if (r < B) v = *reinterpret_cast<const float2*>(h_all + (size_t)r * kG + j);
                hprev[m][hf][q][0] = v.x; hprev[m][hf][q][1] = v.y;
            }
        }
    for (int i = threadIdx.x; i < kRows * kG / 2; i += kThreads) {         // bf16 initial state -> A tile 0
        const int r = i / (kG / 2), c = (i % (kG / 2)) * 2;
        uint32_t v = 0;
        if (row0 + r < B) v = *reinterpret_cast<const uint32_t*>(h_bf + (size_t)(row0 + r) * kG + c);
        *reinterpret_cast<uint32_t*>(&hs[0][r * LD + c]) = v;
    }
    __syncthreads();

    for (int t = 0; t < T; ++t) {
        const __nv_bfloat16* a_tile = hs[t & 1];
        __nv_bfloat16* next_tile = hs[(t + 1) & 1];
        float* gt = gates + (size_t)t * 3 * bg;
        // accumulators start from gi + b_hh for r and z (only their sum matters), from b_hn for n (gh_n is needed alone)
        float acc[3][2][2][4];
        float gin[2][2][2][2];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int r = row0 + 16 * m + g + 8 * hf;
                const float* grow = gt + (size_t)r * 3 * kG;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int j = 16 * warp + 8 * q + 2 * tq;
                    float2 vr = make_float2(0.f, 0.f), vz = vr, vn = vr;
                    if (r < B) {
                        vr = *reinterpret_cast<const float2*>(grow + j);
                        vz = *reinterpret_cast<const float2*>(grow + kG + j);
                        vn = *reinterpret_cast<const float2*>(grow + 2 * kG + j);
                    }
                    acc[0][q][m][2 * hf] = vr.x + bias[0][q][0]; acc[0][q][m][2 * hf + 1] = vr.y + bias[0][q][1];
                    acc[1][q][m][2 * hf] = vz.x + bias[1][q][0]; acc[1][q][m][2 * hf + 1] = vz.y + bias[1][q][1];
                    acc[2][q][m][2 * hf] = bias[2][q][0];        acc[2][q][m][2 * hf + 1] = bias[2][q][1];
                    gin[m][hf][q][0] = vn.x; gin[m][hf][q][1] = vn.y;
                }
            }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            uint32_t af[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const __nv_bfloat16* ap = a_tile + (16 * m + g) * LD + 16 * kk + 2 * tq;
                af[m][0] = *reinterpret_cast<const uint32_t*>(ap);
                af[m][1] = *reinterpret_cast<const uint32_t*>(ap + 8 * LD);
                af[m][2] = *reinterpret_cast<const uint32_t*>(ap + 8);
                af[m][3] = *reinterpret_cast<const uint32_t*>(ap + 8 * LD + 8);
            }
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int m = 0; m < 2; ++m) mma_bf16(acc[s][q][m], af[m], wf[s][q][kk]);
        }
        // ---- cell (torch.nn.GRU): r, z, n, h = (1 - z) n + z h_prev
        float* hn_t = ghn_all + (size_t)t * bg;
        float* hout = h_all + (size_t)(t + 1) * bg;
        __nv_bfloat16* hbout = h_bf + (size_t)(t + 1) * bg;
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int rl = 16 * m + g + 8 * hf, r = row0 + rl;
                float* grow = gt + (size_t)r * 3 * kG;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int j = 16 * warp + 8 * q + 2 * tq;
                    float rr[2], zz[2], nn[2], hh[2], ghn[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        rr[e] = sigmoid_(acc[0][q][m][2 * hf + e]);
                        zz[e] = sigmoid_(acc[1][q][m][2 * hf + e]);
                        ghn[e] = acc[2][q][m][2 * hf + e];
                        nn[e] = tanhf(gin[m][hf][q][e] + rr[e] * ghn[e]);
                        hh[e] = (1.f - zz[e]) * nn[e] + zz[e] * hprev[m][hf][q][e];
                        hprev[m][hf][q][e] = hh[e];
                    }
                    const uint32_t pk = pack2(hh[0], hh[1]);
                    *reinterpret_cast<uint32_t*>(next_tile + rl * LD + j) = r < B ? pk : 0u;
                    if (r < B) {
                        *reinterpret_cast<float2*>(grow + j) = make_float2(rr[0], rr[1]);
                        *reinterpret_cast<float2*>(grow + kG + j) = make_float2(zz[0], zz[1]);
                        *reinterpret_cast<float2*>(grow + 2 * kG + j) = make_float2(nn[0], nn[1]);
                        *reinterpret_cast<float2*>(hn_t + (size_t)r * kG + j) = make_float2(ghn[0], ghn[1]);
                        *reinterpret_cast<float2*>(hout + (size_t)r * kG + j) = make_float2(hh[0], hh[1]);
                        *reinterpret_cast<uint32_t*>(hbout + (size_t)r * kG + j) = pk;
                    }
                }
            }
        __syncthreads();            // the next step's A tile is complete; the tile of step t-1 may be overwritten
    }
}

// ---------------------------------------------------------------------------------------------------- backward
// dout fp32 [T,B,G]: gradient arriving at every h_t from the layers above.  w_hh_t bf16 [G][3G] = W_hh transposed.
// Writes dgi / dgh bf16 [T,B,3G] (gradients w.r.t. the two pre-activation sums).
__global__ void __launch_bounds__(kThreads, 1)
gru_persistent_bwd_kernel(int T, int B, const float* __restrict__ gates, const float* __restrict__ ghn_all,
                          const float* __restrict__ h_all, const float* __restrict__ dout,
                          const __nv_bfloat16* __restrict__ w_hh_t, __nv_bfloat16* __restrict__ dgi,
                          __nv_bfloat16* __restrict__ dgh) {
    constexpr int K = 3 * kG, LD = K + 8;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __nv_bfloat16* tiles = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // [2][kRows * LD]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int row0 = blockIdx.x * kRows;
    const size_t bg = (size_t)B * kG;
    // W_hh^T as B fragments: output n = hidden j (this warp's 16), k = the 3G gate columns
    uint32_t wf[2][24][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const __nv_bfloat16* wrow = w_hh_t + (size_t)(16 * warp + 8 * q + g) * K;
#pragma unroll
        for (int kk = 0; kk < 24; ++kk) {
            wf[q][kk][0] = *reinterpret_cast<const uint32_t*>(wrow + 16 * kk + 2 * tq);
            wf[q][kk][1] = *reinterpret_cast<const uint32_t*>(wrow + 16 * kk + 8 + 2 * tq);
        }
    }
    float dh[2][2][2][2];                  // recurrent part of dL/dh_t (fp32, carried in registers)
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int q = 0; q < 2; ++q) dh[m][hf][q][0] = dh[m][hf][q][1] = 0.f;

    for (int t = T - 1; t >= 0; --t) {
        __nv_bfloat16* tile = tiles + (size_t)(t & 1) * kRows * LD;
        const float* gt = gates + (size_t)t * 3 * bg;
        float carry[2][2][2][2];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int rl = 16 * m + g + 8 * hf, r = row0 + rl;
                const bool live = r < B;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int j = 16 * warp + 8 * q + 2 * tq;
                    float2 vr = make_float2(0.f, 0.f), vz = vr, vn = vr, vhn = vr, vhp = vr, vd = vr;
                    if (live) {
                        const float* grow = gt + (size_t)r * 3 * kG;
                        vr = *reinterpret_cast<const float2*>(grow + j);
                        vz = *reinterpret_cast<const float2*>(grow + kG + j);
                        vn = *reinterpret_cast<const float2*>(grow + 2 * kG + j);
                        vhn = *reinterpret_cast<const float2*>(ghn_all + (size_t)t * bg + (size_t)r * kG + j);
                        vhp = *reinterpret_cast<const float2*>(h_all + (size_t)t * bg + (size_t)r * kG + j);
                        vd = *reinterpret_cast<const float2*>(dout + (size_t)t * bg + (size_t)r * kG + j);
                    }
                    const float rr[2] = {vr.x, vr.y}, zz[2] = {vz.x, vz.y}, nn[2] = {vn.x, vn.y}, hn[2] = {vhn.x, vhn.y},
                                hp[2] = {vhp.x, vhp.y}, dd[2] = {vd.x, vd.y};
                    float dar[2], daz[2], dan[2], dhn[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float d = dd[e] + dh[m][hf][q][e];
                        dan[e] = d * (1.f - zz[e]) * (1.f - nn[e] * nn[e]);
                        daz[e] = d * (hp[e] - nn[e]) * zz[e] * (1.f - zz[e]);
                        dar[e] = dan[e] * hn[e] * rr[e] * (1.f - rr[e]);
                        dhn[e] = dan[e] * rr[e];
                        carry[m][hf][q][e] = d * zz[e];
                    }
                    const uint32_t pr = pack2(dar[0], dar[1]), pz = pack2(daz[0], daz[1]), pn = pack2(dan[0], dan[1]),
                                   ph = pack2(dhn[0], dhn[1]);
                    __nv_bfloat16* trow = tile + rl * LD;
                    *reinterpret_cast<uint32_t*>(trow + j) = live ? pr : 0u;
                    *reinterpret_cast<uint32_t*>(trow + kG + j) = live ? pz : 0u;
                    *reinterpret_cast<uint32_t*>(trow + 2 * kG + j) = live ? ph : 0u;
                    if (live) {
                        const size_t o = (size_t)t * 3 * bg + (size_t)r * 3 * kG;
                        *reinterpret_cast<uint32_t*>(dgi + o + j) = pr;
                        *reinterpret_cast<uint32_t*>(dgi + o + kG + j) = pz;
                        *reinterpret_cast<uint32_t*>(dgi + o + 2 * kG + j) = pn;
                        *reinterpret_cast<uint32_t*>(dgh + o + j) = pr;
                        *reinterpret_cast<uint32_t*>(dgh + o + kG + j) = pz;
                        *reinterpret_cast<uint32_t*>(dgh + o + 2 * kG + j) = ph;
                    }
                }
            }
        if (t == 0) break;
        __syncthreads();                                   // d gh_t is complete in the tile
        float acc[2][2][4];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[q][m][e] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 24; ++kk) {
            uint32_t af[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const __nv_bfloat16* ap = tile + (16 * m + g) * LD + 16 * kk + 2 * tq;
                af[m][0] = *reinterpret_cast<const uint32_t*>(ap);
                af[m][1] = *reinterpret_cast<const uint32_t*>(ap + 8 * LD);
                af[m][2] = *reinterpret_cast<const uint32_t*>(ap + 8);
                af[m][3] = *reinterpret_cast<const uint32_t*>(ap + 8 * LD + 8);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int m = 0; m < 2; ++m) mma_bf16(acc[q][m], af[m], wf[q][kk]);
        }
        // dL/dh_{t-1} (recurrent part) = d gh_t W_hh + dh_t * z_t
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int e = 0; e < 2; ++e) dh[m][hf][q][e] = acc[q][m][2 * hf + e] + carry[m][hf][q][e];
    }
}

}  // namespace

extern "C" int cb_gru_persistent_supported(int32_t G) { return G == kG ? 1 : 0; }

extern "C" int cb_gru_persistent_fwd(int32_t T, int32_t B, int32_t G, float* gates, const void* w_hh_bf16,
                                     const float* b_hh, float* h_all, void* h_all_bf16, float* ghn_all, void* stream) {
    CB_REQUIRE(G == kG, CB_E_ARG, "cb_gru_persistent_fwd: hidden size %d not specialised (128)", G);
    CB_REQUIRE(T >= 1 && B >= 1 && gates && w_hh_bf16 && b_hh && h_all && h_all_bf16 && ghn_all, CB_E_ARG,
               "cb_gru_persistent_fwd: bad argument");
    gru_persistent_fwd_kernel<<<cb::ceil_div(B, kRows), kThreads, 0, cb::as_stream(stream)>>>(
        T, B, gates, (const __nv_bfloat16*)w_hh_bf16, b_hh, h_all, (__nv_bfloat16*)h_all_bf16, ghn_all);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_gru_persistent_bwd(int32_t T, int32_t B, int32_t G, const float* gates, const float* ghn_all,
                                     const float* h_all, const float* dout, const void* w_hh_t_bf16, void* dgi_bf16,
                                     void* dgh_bf16, void* stream) {
    CB_REQUIRE(G == kG, CB_E_ARG, "cb_gru_persistent_bwd: hidden size %d not specialised (128)", G);
    CB_REQUIRE(T >= 1 && B >= 1 && gates && ghn_all && h_all && dout && w_hh_t_bf16 && dgi_bf16 && dgh_bf16, CB_E_ARG,
               "cb_gru_persistent_bwd: bad argument");
    const int smem = 2 * kRows * (3 * kG + 8) * (int)sizeof(__nv_bfloat16);
    static bool attr = false;
    if (!attr) {
        CB_CUDA(cudaFuncSetAttribute(gru_persistent_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    gru_persistent_bwd_kernel<<<cb::ceil_div(B, kRows), kThreads, smem, cb::as_stream(stream)>>>(
        T, B, gates, ghn_all, h_all, dout, (const __nv_bfloat16*)w_hh_t_bf16, (__nv_bfloat16*)dgi_bf16,
        (__nv_bfloat16*)dgh_bf16);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

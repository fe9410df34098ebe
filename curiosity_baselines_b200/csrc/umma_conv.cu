// tcgen05 engine: host side.  Maps a cb_conv_desc onto the TMA-gather GEMM of umma_gemm.cuh by
// choosing a tensor-map *view* of the activation tensor under which every 64-wide k-block of the
// implicit GEMM is one TMA box:
//
//   linear / flatten-linear   [rows, K]                        rank 2, box {64, rows}
//   stride-1 conv, C%64==0    NHWC as (c, w, h, n)             rank 4, box {64, out_w, out_h, S}
//                             start (c0, kw-pad, kh-pad, n0): out-of-range pixels are zero-filled
//                             by TMA, which is exactly the zero padding (and the "full" correlation
//                             of a stride-1 dgrad, run as a pad = k-1 forward over dy)
//   stride-2 conv, C==32      NHWC [n][h][w][32] re-read as [n][h/2][2][w/2][64]: two horizontally
//                             adjacent pixels form one 128-byte k-block row (rank 5, box
//                             {64, out_w, 1, out_h, S})
//
// S = samples per tile so that S*out_h*out_w <= 256 rows (two M=128 UMMA sub-tiles).
#include "umma_gemm.cuh"
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace cb {

namespace {

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    });
    return fn;
}

// bf16 tensor map; dims[0] is the contiguous dimension; strides_bytes has rank-1 entries
int encode_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box) {
    EncodeFn enc = get_encode();
    CB_REQUIRE(enc, CB_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    CB_REQUIRE(((uintptr_t)base & 15) == 0, CB_E_ALIGN, "TMA base address not 16-byte aligned");
    cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) {
        CB_REQUIRE(strides_bytes[i] % 16 == 0, CB_E_ALIGN, "TMA stride %d (%llu B) not a multiple of 16", i,
                   (unsigned long long)strides_bytes[i]);
        gs[i] = strides_bytes[i];
    }
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CB_REQUIRE(r == CUDA_SUCCESS, CB_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CB_OK;
}

}  // namespace

// bf16 tensor map with a chosen swizzle (0 = none, 1 = 32 B, 2 = 64 B, 3 = 128 B); used by umma_conv1.cu
int encode_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle) {
    EncodeFn enc = get_encode();
    CB_REQUIRE(enc, CB_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    CB_REQUIRE(((uintptr_t)base & 15) == 0, CB_E_ALIGN, "TMA base address not 16-byte aligned");
    cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) {
        CB_REQUIRE(strides_bytes[i] % 16 == 0, CB_E_ALIGN, "TMA stride %d not a multiple of 16", i);
        gs[i] = strides_bytes[i];
    }
    const CUtensorMapSwizzle sw = swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CB_REQUIRE(r == CUDA_SUCCESS, CB_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CB_OK;
}

namespace {

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

bool device_ok() {
    static int ok = -1;
    if (ok < 0) ok = cb_device_supports_umma();
    return ok == 1;
}

void identity_map(umma::CoordMap& m) {
    for (int i = 0; i < 5; ++i) { m.div[i] = 1; m.mod[i] = 1; m.mul[i] = 0; }   // (idx/1 % 1)*0 == 0
}

struct Plan {
    CUtensorMap tmA, tmB, tmS;      // tmS: side-input block (GemmParams::side_kb), else unused
    umma::GemmParams p;
    int bn, mt;
};

int epi_mode(const umma::GemmParams& p) {
    if (p.add_pre || p.actgrad_src) return umma::EPI_DGRAD;
    if (p.side_dim || p.residual || p.out_f32 || !p.out_bf16) return umma::EPI_FULL;
    return umma::EPI_SIMPLE;
}

template <int BN, int MT, int EPI>
int launch_epi(const Plan& pl, cudaStream_t s) {
    using Cfg = umma::GemmConfig<BN, MT>;
    static bool attr_set = false;
    if (!attr_set) {
        CB_CUDA(cudaFuncSetAttribute(umma::umma_gemm_kernel<BN, MT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::SMEM_BYTES));
        attr_set = true;
    }
    int tiles = pl.p.m_tiles * pl.p.n_tiles;
    int grid = tiles < sm_count() ? tiles : sm_count();
    umma::umma_gemm_kernel<BN, MT, EPI><<<grid, umma::plain_gemm_threads<EPI>(), Cfg::SMEM_BYTES, s>>>(
        pl.tmA, pl.tmB, pl.p.side_kb ? pl.tmS : pl.tmA, pl.p);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

template <int BN, int MT>
int launch(const Plan& pl, cudaStream_t s) {
    switch (epi_mode(pl.p)) {
        case umma::EPI_DGRAD: return launch_epi<BN, MT, umma::EPI_DGRAD>(pl, s);
        case umma::EPI_FULL:  return launch_epi<BN, MT, umma::EPI_FULL>(pl, s);
        default:              return launch_epi<BN, MT, umma::EPI_SIMPLE>(pl, s);
    }
}

int dispatch(const Plan& pl, cudaStream_t s) {
    switch (pl.bn * 10 + pl.mt) {
        case 2562: return launch<256, 2>(pl, s);
        case 2561: return launch<256, 1>(pl, s);
        case 2242: return launch<224, 2>(pl, s);
        case 2241: return launch<224, 1>(pl, s);
        case 1282: return launch<128, 2>(pl, s);
        case 642:  return launch<64, 2>(pl, s);
        case 322:  return launch<32, 2>(pl, s);
        default:   return fail(CB_E_ARG, "umma: no kernel for BN=%d MT=%d", pl.bn, pl.mt);
    }
}

// A 128x256 tile needs 48 KB of operands per 64-wide k-block -- 96 B/clk/SM at full MMA rate, more than L2 delivers
// (~42 B/clk/SM): the plain GEMMs are L2-bound at ~45 % of peak.  A 256x256 tile (MT = 2) moves 64 KB for twice
// the math.  Used when there are enough rows to keep every SM busy with the larger tiles.
int pick_bn(int n_total, int* mt, long long rows = 0) {
    if (n_total % 256 == 0) {
        *mt = rows >= 256LL * 2 * 148 ? 2 : 1;
        return 256;
    }
    if (n_total % 128 == 0) { *mt = 2; return 128; }
    if (n_total % 224 == 0) { *mt = 1; return 224; }   // (256x224 single-buffered tiles measured slower: fc dgrad has K = 512 only)
    if (n_total % 64 == 0)  { *mt = 2; return 64; }
    if (n_total % 32 == 0)  { *mt = 2; return 32; }
    return 0;
}

// ---- the three A views ---------------------------------------------------------------------
enum View { VIEW_NONE = 0, VIEW_ROWS, VIEW_CONV_S1, VIEW_CONV_S2 };

// geometry of an implicit GEMM whose A operand is a bf16 NHWC tensor [n, h, w, c]
struct Geo {
    int n, h, w, c;            // tensor read by TMA
    int kh, kw, stride, pad;   // filter walk
    int oh, ow;                // rows per sample = oh*ow
};

View classify(const Geo& g) {
    if (g.kh == g.h && g.kw == g.w && g.pad == 0 && g.oh == 1 && g.ow == 1 && (g.h * g.w * g.c) % 8 == 0)
        return VIEW_ROWS;                                   // linear, or Linear over a flattened map
    if (g.stride == 1 && g.c % 64 == 0 && g.oh * g.ow <= 256 && g.ow <= 256 && g.oh <= 256) return VIEW_CONV_S1;
    if (g.stride == 2 && g.c == 32 && g.pad == 0 && g.kw % 2 == 0 && g.w % 2 == 0 && g.h % 2 == 0 &&
        g.oh * g.ow <= 256)
        return VIEW_CONV_S2;
    return VIEW_NONE;
}

// fills tmA and the A-related params; `x` is the bf16 NHWC tensor
int plan_a(Plan& pl, const Geo& g, const void* x, int mt) {
    umma::GemmParams& p = pl.p;
    identity_map(p.a_tile);
    identity_map(p.a_kb);
    View v = classify(g);
    const int max_rows = mt * 128;
    if (v == VIEW_ROWS) {
        const uint64_t K = (uint64_t)g.h * g.w * g.c;
        uint64_t dims[2] = {K, (uint64_t)g.n};
        uint64_t strides[1] = {K * 2};
        uint32_t box[2] = {64, (uint32_t)max_rows};
        int rc = encode_bf16(&pl.tmA, x, 2, dims, strides, box); if (rc) return rc;
        p.a_rank = 2;
        p.a_kb.div[0] = 1; p.a_kb.mod[0] = 0; p.a_kb.mul[0] = 64;
        p.a_tile.div[1] = 1; p.a_tile.mod[1] = 0; p.a_tile.mul[1] = max_rows;
        p.kb_count = (int)((K + 63) / 64);
        p.rows_per_tile = max_rows;
        p.a_box_bytes = max_rows * 128;
        p.m_total = g.n;
        p.m_tiles = (g.n + max_rows - 1) / max_rows;
        return CB_OK;
    }
    const int P = g.oh * g.ow;
    int S = max_rows / P; if (S < 1) return fail(CB_E_ARG, "umma: %d rows per sample exceed the tile", P);
    if (S > 256) S = 256;
    p.rows_per_tile = S * P;
    p.a_box_bytes = S * P * 128;
    p.m_total = (long long)g.n * P;
    p.m_tiles = (g.n + S - 1) / S;
    if (v == VIEW_CONV_S1) {
        const int cb = g.c / 64;
        uint64_t dims[4] = {(uint64_t)g.c, (uint64_t)g.w, (uint64_t)g.h, (uint64_t)g.n};
        uint64_t strides[3] = {(uint64_t)g.c * 2, (uint64_t)g.w * g.c * 2, (uint64_t)g.h * g.w * g.c * 2};
        uint32_t box[4] = {64, (uint32_t)g.ow, (uint32_t)g.oh, (uint32_t)S};
        int rc = encode_bf16(&pl.tmA, x, 4, dims, strides, box); if (rc) return rc;
        p.a_rank = 4;
        // k-block kb = ((kh*KW)+kw)*cb + c64   ->  (c64*64, kw-pad, kh-pad)
        p.a_kb.div[0] = 1;            p.a_kb.mod[0] = cb;   p.a_kb.mul[0] = 64;
        p.a_kb.div[1] = cb;           p.a_kb.mod[1] = g.kw; p.a_kb.mul[1] = 1;
        p.a_kb.div[2] = cb * g.kw;    p.a_kb.mod[2] = 0;    p.a_kb.mul[2] = 1;
        p.a_tile.div[3] = 1; p.a_tile.mod[3] = 0; p.a_tile.mul[3] = S;
        p.kb_count = g.kh * g.kw * cb;
        p.a_off[1] = -g.pad; p.a_off[2] = -g.pad;
        return CB_OK;
    }
    if (v == VIEW_CONV_S2) {
        // [n][h][w][32] == [n][h/2][2][w/2][64]
        uint64_t dims[5] = {64, (uint64_t)g.w / 2, 2, (uint64_t)g.h / 2, (uint64_t)g.n};
        uint64_t strides[4] = {128, (uint64_t)g.w * 64, (uint64_t)g.w * 128, (uint64_t)g.h * g.w * 64};
        uint32_t box[5] = {64, (uint32_t)g.ow, 1, (uint32_t)g.oh, (uint32_t)S};
        int rc = encode_bf16(&pl.tmA, x, 5, dims, strides, box); if (rc) return rc;
        p.a_rank = 5;
        const int kw2 = g.kw / 2;
        // k-block kb = kh*kw2 + j  (j = kw pair)  ->  (0, j, kh%2, kh/2)
        p.a_kb.div[1] = 1;        p.a_kb.mod[1] = kw2; p.a_kb.mul[1] = 1;
        p.a_kb.div[2] = kw2;      p.a_kb.mod[2] = 2;   p.a_kb.mul[2] = 1;
        p.a_kb.div[3] = kw2 * 2;  p.a_kb.mod[3] = 0;   p.a_kb.mul[3] = 1;
        p.a_tile.div[4] = 1; p.a_tile.mod[4] = 0; p.a_tile.mul[4] = S;
        p.kb_count = g.kh * kw2;
        return CB_OK;
    }
    return fail(CB_E_ARG, "umma: unsupported A view");
}

int plan_b(Plan& pl, const void* w, int n_total, long long k_total) {
    uint64_t dims[2] = {(uint64_t)k_total, (uint64_t)n_total};
    uint64_t strides[1] = {(uint64_t)k_total * 2};
    uint32_t box[2] = {64, (uint32_t)pl.bn};
    return encode_bf16(&pl.tmB, w, 2, dims, strides, box);
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// Fold the side input into the GEMM as one more k-block (see GemmParams::side_kb).  side_bf16: [rows, 64].
// Returns the epilogue's side_dim (0 when folded).
int plan_side(Plan& pl, const cb_epilogue* e, int rows, int* side_dim_out) {
    *side_dim_out = e->side_dim;
    if (!e->side_bf16 || !e->side_dim) return CB_OK;
    CB_REQUIRE(e->side_dim <= 64 && pl.p.a_rank == 2, CB_E_ARG, "side_bf16: side_dim must be <= 64 on a plain Linear");
    uint64_t dims[2] = {64, (uint64_t)rows};
    uint64_t strides[1] = {128};
    uint32_t box[2] = {64, (uint32_t)(pl.mt * 128)};
    int rc = encode_bf16(&pl.tmS, e->side_bf16, 2, dims, strides, box); if (rc) return rc;
    pl.p.side_kb = 1;
    pl.p.kb_count += 1;
    *side_dim_out = 0;
    return CB_OK;
}

Geo fwd_geo(const cb_conv_desc* d) {
    return Geo{d->n, d->in_h, d->in_w, d->in_c, d->k_h, d->k_w, d->stride, d->pad, d->out_h, d->out_w};
}

bool dense_nhwc(const cb_conv_desc* d) {
    return d->in_dtype == CB_BF16 && d->in_stride_c == 1 && d->in_stride_w == d->in_c &&
           d->in_stride_h == (int64_t)d->in_w * d->in_c && d->in_stride_n == (int64_t)d->in_h * d->in_w * d->in_c;
}

}  // namespace

// ------------------------------------------------------------------------------ forward
bool window_supports_fwd(const cb_conv_desc* d, const cb_epilogue* e);
int window_conv_fwd(const cb_conv_desc* d, const void* x, const void* w, const cb_epilogue* e, cudaStream_t s);
bool window_supports_dgrad(const cb_conv_desc* d);
int window_conv_dgrad(const cb_conv_desc* d, const void* dy, const void* w_t, const void* x_saved, int x_act,
                      const void* dx_add, void* dx, cudaStream_t s);
int window_conv_dgrad_s2(const cb_conv_desc* d, const void* dy, const void* w_shuffle, const void* x_saved,
                         int x_act, const void* dx_add, void* dx, cudaStream_t s);
static constexpr bool use_window() { return true; }      // (the TMA-gather GEMM remains the path for shapes the window kernels do not tile)

bool umma_supports_fwd(const cb_conv_desc* d, const cb_epilogue* e) {
    if (use_window() && window_supports_fwd(d, e)) return true;
    if (!device_ok() || !dense_nhwc(d) || e->norm_mean) return false;
    int mt;
    if (!pick_bn(d->out_c, &mt)) return false;
    if (e->side_dim > 16 && !e->side_bf16) return false;      // the epilogue path keeps at most 16 side inputs in registers
    Geo g = fwd_geo(d);
    View v = classify(g);
    if (v == VIEW_NONE) return false;
    if (v != VIEW_ROWS && g.oh * g.ow > mt * 128) return false;
    return true;
}

int umma_conv_fwd(const cb_conv_desc* d, const void* x, const void* w, const cb_epilogue* e, cudaStream_t s) {
    CB_REQUIRE(x && w && e && (e->out_bf16 || e->out_f32), CB_E_ARG, "cb_conv_fwd: null pointer");
    if (use_window() && window_supports_fwd(d, e)) return window_conv_fwd(d, x, w, e, s);
    Plan pl{};
    pl.bn = pick_bn(d->out_c, &pl.mt, (d->k_h == d->in_h && d->k_w == d->in_w && d->pad == 0) ? d->n : 0);
    Geo g = fwd_geo(d);
    int rc = plan_a(pl, g, x, pl.mt); if (rc) return rc;
    const long long K = (long long)d->k_h * d->k_w * d->in_c;
    int side_dim = e->side_dim;
    if (e->side_bf16 && K % 64 == 0) { rc = plan_side(pl, e, d->n, &side_dim); if (rc) return rc; }
    rc = plan_b(pl, w, d->out_c, pl.p.side_kb ? K + 64 : K); if (rc) return rc;
    umma::GemmParams& p = pl.p;
    p.n_total = d->out_c;
    p.n_tiles = d->out_c / pl.bn;
    p.bias = e->bias;
    p.side = e->side; p.side_w = e->side_w; p.side_dim = side_dim;
    p.act = d->act;
    p.residual = (const __nv_bfloat16*)e->residual;
    p.out_bf16 = (__nv_bfloat16*)e->out_bf16; p.out_f32 = e->out_f32;
    CB_REQUIRE(aligned16(p.out_bf16) && aligned16(p.out_f32) && aligned16(p.residual), CB_E_ALIGN,
               "cb_conv_fwd: outputs must be 16-byte aligned");
    return dispatch(pl, s);
}

// ------------------------------------------------------------------------------ dgrad
// dx[rows=(n,ih,iw), ci] = sum_{kh,kw,co} dy[n, ih+pad-kh, iw+pad-kw, co] * w_t[ci][kh][kw][co]   (stride 1)
// which is a stride-1 forward over dy with the filter walked backwards; linear layers are the 1x1 case.
bool umma_supports_dgrad(const cb_conv_desc* d) {
    if (!device_ok()) return false;
    if (use_window() && window_supports_dgrad(d)) return true;
    int mt;
    if (!pick_bn(d->in_c, &mt)) return false;
    if (d->k_h == 1 && d->k_w == 1 && d->in_h == 1 && d->in_w == 1 && d->out_c % 8 == 0) return true;
    if (d->stride == 1 && d->out_c % 64 == 0 && d->in_h * d->in_w <= mt * 128 &&
        !(d->k_h == d->in_h && d->k_w == d->in_w))
        return true;
    return false;
}

int umma_conv_dgrad(const cb_conv_desc* d, const void* dy, const void* w_t, const void* x_saved, int x_act,
                    const void* dx_add, void* dx, cudaStream_t s) {
    CB_REQUIRE(dy && w_t && dx, CB_E_ARG, "cb_conv_dgrad: null pointer");
    if (use_window() && window_supports_dgrad(d)) return window_conv_dgrad(d, dy, w_t, x_saved, x_act, dx_add, dx, s);
    Plan pl{};
    pl.bn = pick_bn(d->in_c, &pl.mt, (d->k_h == 1 && d->k_w == 1 && d->in_h == 1 && d->in_w == 1) ? d->n : 0);
    umma::GemmParams& p = pl.p;
    int rc;
    if (d->k_h == 1 && d->k_w == 1 && d->in_h == 1 && d->in_w == 1) {
        Geo g{d->n, 1, 1, d->out_c, 1, 1, 1, 0, 1, 1};
        rc = plan_a(pl, g, dy, pl.mt); if (rc) return rc;
    } else {
        // rows are input pixels (ih, iw); tap (kh, kw) reads dy at (ih + pad - kh, iw + pad - kw)
        Geo g{d->n, d->out_h, d->out_w, d->out_c, d->k_h, d->k_w, 1, 0, d->in_h, d->in_w};
        rc = plan_a(pl, g, dy, pl.mt); if (rc) return rc;
        // coordinate = pad - k: walk the filter backwards from a constant offset
        p.a_kb.mul[1] = -1; p.a_kb.mul[2] = -1;
        p.a_off[1] = d->pad; p.a_off[2] = d->pad;
    }
    const long long K = (long long)d->k_h * d->k_w * d->out_c;
    rc = plan_b(pl, w_t, d->in_c, K); if (rc) return rc;
    p.n_total = d->in_c;
    p.n_tiles = d->in_c / pl.bn;
    p.add_pre = (const __nv_bfloat16*)dx_add;
    p.act = CB_ACT_NONE;
    p.actgrad_src = (const __nv_bfloat16*)x_saved; p.actgrad_act = x_act;
    p.out_bf16 = (__nv_bfloat16*)dx;
    CB_REQUIRE(aligned16(dx) && aligned16(dx_add) && aligned16(x_saved), CB_E_ALIGN,
               "cb_conv_dgrad: operands must be 16-byte aligned");
    return dispatch(pl, s);
}

// ------------------------------------------------------------------------------ grouped linear layers
// E independent Linear(K -> F) problems over the same n rows as ONE launch (Disagreement's ensemble of
// forward models, disagreement.py:97-100,127-130).  Activations of all problems live side by side in one
// row-major tensor [n, E*F]; weights are stacked along their output dimension.  The product is then a
// block-diagonal GEMM: output column block g reads A columns [g*K, (g+1)*K) (or [0, K) when every
// problem reads the same input), so only the producer's K coordinate depends on the group.
bool umma_supports_grouped(int groups, int k, int f) {
    int mt;
    return device_ok() && groups >= 1 && k % 64 == 0 && f % 64 == 0 && pick_bn(f, &mt) != 0 && f % pick_bn(f, &mt) == 0;
}

int umma_grouped_fwd(int n, int groups, int k, int f, int x_shared, const void* x, const void* w_all, int act,
                     const cb_epilogue* e, cudaStream_t s) {
    CB_REQUIRE(x && w_all && e && (e->out_bf16 || e->out_f32), CB_E_ARG, "cb_grouped_linear_fwd: null pointer");
    // (side inputs folded in as one more k-block -- side_bf16 -- may be up to 64 wide; the epilogue path keeps 16)
    CB_REQUIRE(umma_supports_grouped(groups, k, f) && (e->side_dim <= 16 || (e->side_bf16 && e->side_dim <= 64)) &&
               !e->norm_mean, CB_E_ARG, "cb_grouped_linear_fwd: unsupported shape or device");
    Plan pl{};
    pl.bn = pick_bn(f, &pl.mt, n);
    Geo g{n, 1, 1, x_shared ? k : groups * k, 1, 1, 1, 0, 1, 1};
    int rc = plan_a(pl, g, x, pl.mt); if (rc) return rc;
    umma::GemmParams& p = pl.p;
    p.kb_count = k / 64;
    int side_dim = e->side_dim;
    rc = plan_side(pl, e, n, &side_dim); if (rc) return rc;
    rc = plan_b(pl, w_all, groups * f, p.side_kb ? k + 64 : k); if (rc) return rc;
    p.n_total = groups * f;
    p.n_tiles = groups * f / pl.bn;
    p.group_tiles = f / pl.bn;
    p.a_group_koff = x_shared ? 0 : k;
    p.bias = e->bias;
    p.side = e->side; p.side_w = e->side_w; p.side_dim = side_dim;
    p.act = act;
    p.residual = (const __nv_bfloat16*)e->residual;
    p.out_bf16 = (__nv_bfloat16*)e->out_bf16; p.out_f32 = e->out_f32;
    CB_REQUIRE(aligned16(p.out_bf16) && aligned16(p.out_f32) && aligned16(p.residual), CB_E_ALIGN,
               "cb_grouped_linear_fwd: outputs must be 16-byte aligned");
    return dispatch(pl, s);
}

// dx[n, E*K] block g = (dy[:, g*F:(g+1)*F] W_g + dx_add) .* act'(x_saved);  w_t_all: bf16 [E*K, F]
int umma_grouped_dgrad(int n, int groups, int k, int f, const void* dy, const void* w_t_all, const void* x_saved,
                       int x_act, const void* dx_add, void* dx, cudaStream_t s) {
    CB_REQUIRE(dy && w_t_all && dx, CB_E_ARG, "cb_grouped_linear_dgrad: null pointer");
    CB_REQUIRE(umma_supports_grouped(groups, f, k), CB_E_ARG, "cb_grouped_linear_dgrad: unsupported shape or device");
    Plan pl{};
    pl.bn = pick_bn(k, &pl.mt, n);
    Geo g{n, 1, 1, groups * f, 1, 1, 1, 0, 1, 1};
    int rc = plan_a(pl, g, dy, pl.mt); if (rc) return rc;
    rc = plan_b(pl, w_t_all, groups * k, f); if (rc) return rc;
    umma::GemmParams& p = pl.p;
    p.kb_count = f / 64;
    p.n_total = groups * k;
    p.n_tiles = groups * k / pl.bn;
    p.group_tiles = k / pl.bn;
    p.a_group_koff = f;
    p.add_pre = (const __nv_bfloat16*)dx_add;
    p.act = CB_ACT_NONE;
    p.actgrad_src = (const __nv_bfloat16*)x_saved; p.actgrad_act = x_act;
    p.out_bf16 = (__nv_bfloat16*)dx;
    CB_REQUIRE(aligned16(dx) && aligned16(dx_add) && aligned16(x_saved), CB_E_ALIGN,
               "cb_grouped_linear_dgrad: operands must be 16-byte aligned");
    return dispatch(pl, s);
}

// ------------------------------------------------------------------------------ stride-2 dgrad
// k = 2*stride = 4, pad 0 (Burda conv2).  With ih = 2a+ph, kh = 2u+ph the four output parities share
// the same 2x2 taps of dy, so one GEMM computes all of them:
//   D[(n,a,b), (ph,pw,ci)] = sum_{u,v,co} dy[n, a-u, b-v, co] * W[co, 2u+ph, 2v+pw, ci]
// (M = n*(in_h/2)*(in_w/2), K = 4*out_c, N = 4*in_c) and the epilogue scatters the four 32-channel
// chunks of a row to the pixels (2a+ph, 2b+pw) -- a depth-to-space store.
bool umma_supports_dgrad_s2(const cb_conv_desc* d) {
    return device_ok() && d->stride == 2 && d->k_h == 4 && d->k_w == 4 && d->pad == 0 && d->in_c == 32 &&
           d->out_c % 64 == 0 && d->in_h % 2 == 0 && d->in_w % 2 == 0 && (d->in_h / 2) * (d->in_w / 2) <= 256;
}

int umma_conv_dgrad_s2(const cb_conv_desc* d, const void* dy, const void* w_shuffle, const void* x_saved,
                       int x_act, const void* dx_add, void* dx, cudaStream_t s) {
    CB_REQUIRE(dy && w_shuffle && dx, CB_E_ARG, "cb_conv_dgrad_s2: null pointer");
    CB_REQUIRE(umma_supports_dgrad_s2(d), CB_E_ARG, "cb_conv_dgrad_s2: unsupported shape or device");
    if (use_window()) return window_conv_dgrad_s2(d, dy, w_shuffle, x_saved, x_act, dx_add, dx, s);
    Plan pl{};
    pl.bn = 128; pl.mt = 2;
    umma::GemmParams& p = pl.p;
    Geo g{d->n, d->out_h, d->out_w, d->out_c, 2, 2, 1, 0, d->in_h / 2, d->in_w / 2};
    int rc = plan_a(pl, g, dy, pl.mt); if (rc) return rc;
    p.a_kb.mul[1] = -1; p.a_kb.mul[2] = -1;
    rc = plan_b(pl, w_shuffle, 4 * d->in_c, 4LL * d->out_c); if (rc) return rc;
    p.n_total = 4 * d->in_c;
    p.n_tiles = 1;
    p.add_pre = (const __nv_bfloat16*)dx_add;
    p.act = CB_ACT_NONE;
    p.actgrad_src = (const __nv_bfloat16*)x_saved; p.actgrad_act = x_act;
    p.out_bf16 = (__nv_bfloat16*)dx;
    p.shuffle = 1; p.sh_ph = d->in_h / 2; p.sh_pw = d->in_w / 2; p.sh_h = d->in_h; p.sh_w = d->in_w;
    CB_REQUIRE(aligned16(dx) && aligned16(dx_add) && aligned16(x_saved), CB_E_ALIGN,
               "cb_conv_dgrad_s2: operands must be 16-byte aligned");
    return dispatch(pl, s);
}

}  // namespace cb

// =================================================================================================
// Window kernels (umma_window.cuh): host-side planning
// =================================================================================================
#include "umma_window.cuh"

namespace cb {
namespace {

constexpr int kSmemBudget = 224 * 1024;      // dynamic shared memory we allow ourselves per CTA

int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct WindowGeo {
    // tensor read through TMA (bf16 NHWC) and its view
    const void* x; int n, h, w, c;
    bool s2;                    // stride-2 k4 c32 "two pixels per row" view with two row-parity planes
    int gh, gw, vh, vw;         // pixel grid per sample and stored part
    int y0, x0;                 // grid origin in tensor coordinates (negative = zero padding)
    int ntaps; int tap_plane[umma::kMaxTaps], tap_shift[umma::kMaxTaps];
};

// encode the slab tensor map (box = one plane of S samples) and fill the slab part of WindowParams
int plan_slab(CUtensorMap* tm, const WindowGeo& g, int S, umma::WindowParams& p) {
    p.S = S; p.gh = g.gh; p.gw = g.gw; p.vh = g.vh; p.vw = g.vw; p.n_samples = g.n;
    p.tiles = (g.n + S - 1) / S;
    p.plane_rows = S * g.gh * g.gw;
    p.plane_bytes = round_up(p.plane_rows * 128, 1024);
    memset(p.a_coord, 0, sizeof(p.a_coord));
    if (!g.s2) {
        uint64_t dims[4] = {(uint64_t)g.c, (uint64_t)g.w, (uint64_t)g.h, (uint64_t)g.n};
        uint64_t strides[3] = {(uint64_t)g.c * 2, (uint64_t)g.w * g.c * 2, (uint64_t)g.h * g.w * g.c * 2};
        uint32_t box[4] = {64, (uint32_t)g.gw, (uint32_t)g.gh, (uint32_t)S};
        int rc = encode_bf16(tm, g.x, 4, dims, strides, box); if (rc) return rc;
        p.a_rank = 4; p.a_sample_dim = 3; p.nplanes = 1;
        p.a_coord[0][1] = g.x0; p.a_coord[0][2] = g.y0;
    } else {
        uint64_t dims[5] = {64, (uint64_t)g.w / 2, 2, (uint64_t)g.h / 2, (uint64_t)g.n};
        uint64_t strides[4] = {128, (uint64_t)g.w * 64, (uint64_t)g.w * 128, (uint64_t)g.h * g.w * 64};
        uint32_t box[5] = {64, (uint32_t)g.gw, 1, (uint32_t)g.gh, (uint32_t)S};
        int rc = encode_bf16(tm, g.x, 5, dims, strides, box); if (rc) return rc;
        p.a_rank = 5; p.a_sample_dim = 4; p.nplanes = 2;
        p.a_coord[1][2] = 1;                                 // second plane: odd rows
    }
    p.stage_bytes = p.nplanes * p.plane_bytes;
    p.kb_count = g.ntaps;
    int max_shift = 0;
    for (int i = 0; i < g.ntaps; ++i) {
        p.kb_plane[i] = g.tap_plane[i]; p.kb_shift[i] = g.tap_shift[i];
        p.kb_off16[i] = (g.tap_plane[i] * p.plane_bytes + g.tap_shift[i] * 128) >> 4;
        if (g.tap_shift[i] > max_shift) max_shift = g.tap_shift[i];
    }
    p.slack_bytes = round_up((256 + max_shift) * 128, 1024);   // a shifted M=256 window may run this far past a plane
    return CB_OK;
}

void taps_s1(WindowGeo& g, int k, bool flipped) {
    g.ntaps = k * k;
    for (int kh = 0; kh < k; ++kh)
        for (int kw = 0; kw < k; ++kw) {
            int i = kh * k + kw;
            g.tap_plane[i] = 0;
            g.tap_shift[i] = flipped ? ((k - 1 - kh) * g.gw + (k - 1 - kw)) : (kh * g.gw + kw);
        }
}

template <int BN, int EPI>
int launch_window_epi(const CUtensorMap& tmA, const CUtensorMap& tmW, umma::WindowParams& p, cudaStream_t s) {
    p.epi.debug = 0;
    const int w_bytes = p.kb_count * BN * 128;
    // forward-simple epilogue: let the tensor cores add the bias when its operand blocks (4 KB + BN*32 B) fit
    int extra = 0;
    if (EPI == umma::EPI_SIMPLE && p.epi.bias && kSmemBudget - w_bytes - p.slack_bytes - 2048 - 4096 - BN * 32 >= 2 * p.stage_bytes) {
        p.epi.bias_in_acc = 1;
        extra = 4096 + BN * 32;
    }
    int avail = kSmemBudget - w_bytes - p.slack_bytes - 1024 - 1024 - extra;
    p.stages = avail / p.stage_bytes;
    if (p.stages > umma::kMaxStages) p.stages = umma::kMaxStages;
    CB_REQUIRE(p.stages >= 2, CB_E_ARG, "window conv: slab does not fit shared memory (stage %d B)", p.stage_bytes);
    const int smem = w_bytes + p.stages * p.stage_bytes + p.slack_bytes + 1024 + 1024 + extra;   // + barriers + bias[BN] (+ bias blocks)
    static bool attr = false;
    if (!attr) {
        CB_CUDA(cudaFuncSetAttribute(umma::umma_window_conv_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kSmemBudget));
        attr = true;
    }
    int grid = p.tiles < sm_count() ? p.tiles : sm_count();
    umma::umma_window_conv_kernel<BN, EPI><<<grid, umma::gemm_threads<EPI>(), smem, s>>>(tmA, tmW, p);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

template <int BN>
int launch_window(const CUtensorMap& tmA, const CUtensorMap& tmW, umma::WindowParams& p, cudaStream_t s) {
    switch (epi_mode(p.epi)) {
        case umma::EPI_DGRAD: return launch_window_epi<BN, umma::EPI_DGRAD>(tmA, tmW, p, s);
        case umma::EPI_FULL:  return launch_window_epi<BN, umma::EPI_FULL>(tmA, tmW, p, s);
        default:              return launch_window_epi<BN, umma::EPI_SIMPLE>(tmA, tmW, p, s);
    }
}

int plan_weights(CUtensorMap* tm, const void* w, int bn, long long k_total) {
    uint64_t dims[2] = {(uint64_t)k_total, (uint64_t)bn};
    uint64_t strides[1] = {(uint64_t)k_total * 2};
    uint32_t box[2] = {64, (uint32_t)bn};
    return encode_bf16(tm, w, 2, dims, strides, box);
}

// two stages of one-plane slabs + resident weights + slack must fit the shared-memory budget
bool window_fits(int ntaps, int bn, int plane_rows, int nplanes, int max_shift) {
    int w_bytes = ntaps * bn * 128;
    int stage = nplanes * round_up(plane_rows * 128, 1024);
    int slack = round_up((256 + max_shift) * 128, 1024);
    return w_bytes + 2 * stage + slack + 2048 <= kSmemBudget;
}

bool window_shape_s1(int c, int k, int gh, int gw, int n_out) {
    if (!(c == 64 && k * k <= umma::kMaxTaps && gh * gw <= 256 && gw <= 256 && gh <= 256 &&
          (n_out == 64 || n_out == 128)))
        return false;
    int S = 256 / (gh * gw);
    return window_fits(k * k, n_out, S * gh * gw, 1, (k - 1) * gw + k - 1);
}

}  // namespace

// ---- forward ----------------------------------------------------------------------------------
bool window_supports_fwd(const cb_conv_desc* d, const cb_epilogue* e) {
    if (!device_ok() || !dense_nhwc(d) || e->norm_mean || e->side_dim) return false;
    if (d->stride == 1 && d->k_h == d->k_w && d->k_h > 1 && !(d->k_h == d->in_h && d->pad == 0))
        return window_shape_s1(d->in_c, d->k_h, d->in_h + 2 * d->pad, d->in_w + 2 * d->pad, d->out_c);
    if (d->stride == 2 && d->k_h == 4 && d->k_w == 4 && d->pad == 0 && d->in_c == 32 && d->in_h % 2 == 0 &&
        d->in_w % 2 == 0 && (d->in_h / 2) * (d->in_w / 2) <= 128 && (d->out_c == 64 || d->out_c == 128))
        return true;
    return false;
}

// Stride-2 k4 c32 forward (Burda conv2) in flat mode: the tensor is viewed as [n*h/2][2][w/2][64] and a tile is
// 25 consecutive grid rows of the flattened (sample, row) axis plus a two-row halo for the taps.
int window_conv_fwd_s2_flat(const cb_conv_desc* d, const void* x, const void* w, const cb_epilogue* e, cudaStream_t s) {
    const int gh = d->in_h / 2, gw = d->in_w / 2;
    const int RT = 256 / gw, halo = 2;                     // grid rows per tile; taps reach one grid row + one column further
    umma::WindowParams p{};
    CUtensorMap tmA, tmW;
    p.S = 1; p.gh = gh; p.gw = gw; p.vh = d->out_h; p.vw = d->out_w; p.n_samples = d->n;
    p.flat_rows = RT;
    p.tiles = (int)(((long long)d->n * gh + RT - 1) / RT);
    p.plane_rows = (RT + halo) * gw;
    p.plane_bytes = round_up(p.plane_rows * 128, 1024);
    uint64_t dims[4] = {64, (uint64_t)gw, 2, (uint64_t)d->n * gh};
    uint64_t strides[3] = {128, (uint64_t)d->in_w * 64, (uint64_t)d->in_w * 128};
    uint32_t box[4] = {64, (uint32_t)gw, 1, (uint32_t)(RT + halo)};
    int rc = encode_bf16(&tmA, x, 4, dims, strides, box); if (rc) return rc;
    p.a_rank = 4; p.a_sample_dim = 3; p.nplanes = 2;
    p.a_coord[1][2] = 1;                                   // second plane: odd rows
    p.stage_bytes = p.nplanes * p.plane_bytes;
    p.kb_count = 8;
    int max_shift = 0;
    for (int kh = 0; kh < 4; ++kh)
        for (int j = 0; j < 2; ++j) {
            const int i = kh * 2 + j, shift = (kh >> 1) * gw + j;
            p.kb_plane[i] = kh & 1; p.kb_shift[i] = shift;
            p.kb_off16[i] = ((kh & 1) * p.plane_bytes + shift * 128) >> 4;
            if (shift > max_shift) max_shift = shift;
        }
    const int over = 256 + max_shift - p.plane_rows;       // rows a shifted M=256 window reads past a plane
    p.slack_bytes = over > 0 ? round_up(over * 128, 1024) : 0;
    rc = plan_weights(&tmW, w, d->out_c, (long long)d->k_h * d->k_w * d->in_c); if (rc) return rc;
    p.epi.n_total = d->out_c;
    p.epi.bias = e->bias; p.epi.act = d->act;
    p.epi.residual = (const __nv_bfloat16*)e->residual;
    p.epi.out_bf16 = (__nv_bfloat16*)e->out_bf16; p.epi.out_f32 = e->out_f32;
    return d->out_c == 64 ? launch_window<64>(tmA, tmW, p, s) : launch_window<128>(tmA, tmW, p, s);
}

int window_conv_fwd(const cb_conv_desc* d, const void* x, const void* w, const cb_epilogue* e, cudaStream_t s) {
    CB_REQUIRE(x && w && e && (e->out_bf16 || e->out_f32), CB_E_ARG, "cb_conv_fwd: null pointer");
    {
        const int gw = d->in_w / 2, gh = d->in_h / 2;
        // worth it when whole samples leave many of the 256 tile rows empty
        if (d->stride == 2 && !e->residual && gw <= 128 && (256 / (gh * gw)) * gh * gw < (256 / gw) * gw - gw &&
            (long long)d->n * gh < (1ll << 31))
            return window_conv_fwd_s2_flat(d, x, w, e, s);
    }
    WindowGeo g{};
    g.x = x; g.n = d->n; g.h = d->in_h; g.w = d->in_w; g.c = d->in_c;
    g.vh = d->out_h; g.vw = d->out_w;
    if (d->stride == 1) {
        g.s2 = false; g.gh = d->in_h + 2 * d->pad; g.gw = d->in_w + 2 * d->pad; g.y0 = g.x0 = -d->pad;
        taps_s1(g, d->k_h, false);
    } else {
        g.s2 = true; g.gh = d->in_h / 2; g.gw = d->in_w / 2;
        g.ntaps = 8;
        for (int kh = 0; kh < 4; ++kh)
            for (int j = 0; j < 2; ++j) { g.tap_plane[kh * 2 + j] = kh & 1; g.tap_shift[kh * 2 + j] = (kh >> 1) * g.gw + j; }
    }
    umma::WindowParams p{};
    CUtensorMap tmA, tmW;
    int S = 256 / (g.gh * g.gw);
    int rc = plan_slab(&tmA, g, S, p); if (rc) return rc;
    rc = plan_weights(&tmW, w, d->out_c, (long long)d->k_h * d->k_w * d->in_c); if (rc) return rc;
    p.epi.n_total = d->out_c;
    p.epi.bias = e->bias; p.epi.act = d->act;
    p.epi.residual = (const __nv_bfloat16*)e->residual;
    p.epi.out_bf16 = (__nv_bfloat16*)e->out_bf16; p.epi.out_f32 = e->out_f32;
    return d->out_c == 64 ? launch_window<64>(tmA, tmW, p, s) : launch_window<128>(tmA, tmW, p, s);
}

// ---- dgrad -------------------------------------------------------------------------------------
bool window_supports_dgrad(const cb_conv_desc* d) {
    if (!device_ok()) return false;
    if (d->stride == 1 && d->k_h == d->k_w && d->k_h > 1 && !(d->k_h == d->in_h && d->pad == 0))
        return window_shape_s1(d->out_c, d->k_h, d->in_h + d->k_h - 1, d->in_w + d->k_w - 1, d->in_c);
    return false;
}

int window_conv_dgrad(const cb_conv_desc* d, const void* dy, const void* w_t, const void* x_saved, int x_act,
                      const void* dx_add, void* dx, cudaStream_t s) {
    CB_REQUIRE(dy && w_t && dx, CB_E_ARG, "cb_conv_dgrad: null pointer");
    const int k = d->k_h;
    WindowGeo g{};
    g.x = dy; g.n = d->n; g.h = d->out_h; g.w = d->out_w; g.c = d->out_c; g.s2 = false;
    g.gh = d->in_h + k - 1; g.gw = d->in_w + k - 1; g.vh = d->in_h; g.vw = d->in_w;
    g.y0 = g.x0 = d->pad - (k - 1);
    taps_s1(g, k, true);
    umma::WindowParams p{};
    CUtensorMap tmA, tmW;
    int S = 256 / (g.gh * g.gw);
    int rc = plan_slab(&tmA, g, S, p); if (rc) return rc;
    rc = plan_weights(&tmW, w_t, d->in_c, (long long)k * k * d->out_c); if (rc) return rc;
    p.epi.n_total = d->in_c;
    p.epi.add_pre = (const __nv_bfloat16*)dx_add;
    p.epi.actgrad_src = (const __nv_bfloat16*)x_saved; p.epi.actgrad_act = x_act;
    p.epi.out_bf16 = (__nv_bfloat16*)dx;
    return d->in_c == 64 ? launch_window<64>(tmA, tmW, p, s) : launch_window<128>(tmA, tmW, p, s);
}

int window_conv_dgrad_s2(const cb_conv_desc* d, const void* dy, const void* w_shuffle, const void* x_saved,
                         int x_act, const void* dx_add, void* dx, cudaStream_t s) {
    WindowGeo g{};
    g.x = dy; g.n = d->n; g.h = d->out_h; g.w = d->out_w; g.c = d->out_c; g.s2 = false;
    g.vh = d->in_h / 2; g.vw = d->in_w / 2; g.gh = g.vh + 1; g.gw = g.vw + 1;
    g.y0 = g.x0 = -1;
    g.ntaps = 4;
    for (int u = 0; u < 2; ++u)
        for (int v = 0; v < 2; ++v) { g.tap_plane[u * 2 + v] = 0; g.tap_shift[u * 2 + v] = (1 - u) * g.gw + (1 - v); }
    umma::WindowParams p{};
    CUtensorMap tmA, tmW;
    int S = 256 / (g.gh * g.gw);
    CB_REQUIRE(S >= 1 && d->out_c == 64 && d->in_c == 32, CB_E_ARG, "cb_conv_dgrad_s2: unsupported shape");
    int rc = plan_slab(&tmA, g, S, p); if (rc) return rc;
    rc = plan_weights(&tmW, w_shuffle, 128, 4LL * d->out_c); if (rc) return rc;
    p.epi.n_total = 128;
    p.epi.add_pre = (const __nv_bfloat16*)dx_add;
    p.epi.actgrad_src = (const __nv_bfloat16*)x_saved; p.epi.actgrad_act = x_act;
    p.epi.out_bf16 = (__nv_bfloat16*)dx;
    p.epi.shuffle = 1; p.epi.sh_ph = g.vh; p.epi.sh_pw = g.vw; p.epi.sh_h = d->in_h; p.epi.sh_w = d->in_w;
    return launch_window<128>(tmA, tmW, p, s);
}

}  // namespace cb

// =================================================================================================
// Weight gradients on tcgen05 (umma_wgrad_kernel): MN-major operands, split-K over CTAs
// =================================================================================================
namespace cb {

int scale_buffer(float* x, int64_t n, float beta, cudaStream_t s);
int colsum_bf16(const void* dy, int64_t rows, int c, float* out, float beta, cudaStream_t s);

namespace {

enum WgradKind { WG_NONE = 0, WG_ROWS, WG_CONV_S1, WG_CONV_S2 };

WgradKind wgrad_kind(const cb_conv_desc* d) {
    if (!device_ok() || !dense_nhwc(d)) return WG_NONE;
    const long long K = (long long)d->k_h * d->k_w * d->in_c;
    if (d->k_h == d->in_h && d->k_w == d->in_w && d->pad == 0 && d->out_h == 1 && d->out_w == 1 &&
        d->out_c % 8 == 0 && K % 8 == 0)
        return WG_ROWS;        // (an output dimension that is not a multiple of 128 runs a partly empty last row tile)
    if (d->stride == 1 && d->k_h == d->k_w && d->k_h > 1 && d->k_h * d->k_w <= 2 * umma::kMaxAcc &&
        d->in_c == 64 && d->out_c == 64 && (d->in_h + 2 * d->pad) * (d->in_w + 2 * d->pad) <= 81)
        return WG_CONV_S1;
    if (d->stride == 2 && d->k_h == 4 && d->k_w == 4 && d->pad == 0 && d->in_c == 32 && d->out_c == 64 &&
        d->in_h % 2 == 0 && d->in_w % 2 == 0 && (d->in_h / 2) * (d->in_w / 2) <= 100)
        return WG_CONV_S2;
    return WG_NONE;
}

template <int BN>
int launch_wgrad(const CUtensorMap& tmA, const CUtensorMap& tmB, const umma::WgradParams& p, cudaStream_t s) {
    const int smem = p.stages * p.stage_bytes + p.slack_bytes + 1024 + 512;
    CB_REQUIRE(smem <= kSmemBudget && p.stages >= 2, CB_E_ARG, "wgrad: stages do not fit shared memory");
    static bool attr = false;
    if (!attr) {
        CB_CUDA(cudaFuncSetAttribute(umma::umma_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kSmemBudget));
        attr = true;
    }
    umma::umma_wgrad_kernel<BN><<<p.units * p.splits, 192, smem, s>>>(tmA, tmB, p);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// dW[g][out][K] = dy_g^T x_g for `groups` problems stacked as dy [n, groups*out], x [n, groups*K] (or the
// shared [n, K]), dw [groups*out, K]:  A = dy (M = 128-row chunks of the stacked output dimension),
// B = x (N = 256 columns of the group's K range), split-K over CTAs.
int wgrad_rows(int n, int out_c, int K, int groups, int x_shared, const void* x, const void* dy, float* dw,
               cudaStream_t s) {
    umma::WgradParams p{};
    CUtensorMap tmA, tmB;
    p.dw = dw;
    const int out_all = groups * out_c, k_all = (groups > 1 && !x_shared) ? groups * K : K;
    uint64_t da[2] = {(uint64_t)out_all, (uint64_t)n}; uint64_t sa[1] = {(uint64_t)out_all * 2};
    uint64_t db[2] = {(uint64_t)k_all, (uint64_t)n};   uint64_t sb[1] = {(uint64_t)k_all * 2};
    uint32_t box[2] = {64, 64};
    int rc = encode_bf16(&tmA, dy, 2, da, sa, box); if (rc) return rc;
    rc = encode_bf16(&tmB, x, 2, db, sb, box); if (rc) return rc;
    // MC = 128-row output chunks per CTA: two of them share each x slab (64 KB of operands per 256x256x64 MACs
    // instead of 48 KB per 128x256x64 -- these GEMMs are L2-bandwidth-bound)
    // small problems (512 x 512): the atomic epilogue of the doubled tile would dominate -- measured slower
    const int MC = (out_c % 256 == 0 && (K >= 1024 || out_all >= 2048)) ? 2 : 1;
    p.na = 2 * MC; p.nb = 4; p.a_rank = p.b_rank = 2;
    for (int i = 0; i < 4; ++i) { p.a_coord[i][0] = i * 64; p.b_coord[i][0] = i * 64; }
    p.a_step_dim = p.b_step_dim = 1; p.a_step = p.b_step = 64;
    p.a_unit_dim = p.b_unit_dim = 0; p.a_unit_step = 128 * MC; p.b_unit_step = 256;
    p.a_box_bytes = p.b_box_bytes = 8192; p.a_box_stride = p.b_box_stride = 8192;
    p.b_offset = 2 * MC * 8192; p.stage_bytes = (2 * MC + 4) * 8192; p.stages = MC == 2 ? 3 : 4; p.slack_bytes = 0; p.k_rows = 64;
    const int n_tiles = (K + 255) / 256;
    const int m_units = (out_all + 128 * MC - 1) / (128 * MC);
    p.units = m_units * n_tiles; p.unit_div = n_tiles;
    if (out_all % (128 * MC)) { p.m_total = out_all; p.m_unit_rows = 128 * MC; }      // last row tile partly past the end: TMA zero-fills, stores are guarded
    if (groups > 1) { p.group_units = out_c / (128 * MC); p.b_group_koff = x_shared ? 0 : K; }
    p.slabs_total = (n + 63) / 64;
    p.n_acc = MC;
    for (int g = 0; g < MC; ++g) {
        p.acc_a_off[g] = g * 2 * 8192; p.acc_a_lbo[g] = 8192;
        p.acc_base0[g] = (long long)g * 128 * K; p.acc_base1[g] = (long long)g * 128 * K + 64LL * K;
    }
    p.m_stride = K; p.n_stride = 1; p.unit_m_off = 128LL * MC * K; p.unit_n_off = 256;
    p.n_cols_total = K; p.n_unit_cols = 256;
    int splits = (2 * sm_count()) / p.units; if (splits < 1) splits = 1;
    if (splits > p.slabs_total) splits = p.slabs_total;
    p.splits = splits;
    return launch_wgrad<256>(tmA, tmB, p, s);
}

}  // namespace

bool umma_supports_wgrad(const cb_conv_desc* d) { return wgrad_kind(d) != WG_NONE; }

// dw_all [groups*F, K] = beta*dw_all + per-group dy_g^T x_g;  dbias_all [groups*F] likewise
int umma_grouped_wgrad(int n, int groups, int k, int f, int x_shared, const void* x, const void* dy, float* dw_all,
                       float* dbias_all, float beta, cudaStream_t s) {
    CB_REQUIRE(x && dy && dw_all, CB_E_ARG, "cb_grouped_linear_wgrad: null pointer");
    CB_REQUIRE(device_ok() && f % 128 == 0 && k % 8 == 0 && (x_shared || k % 64 == 0), CB_E_ARG,
               "cb_grouped_linear_wgrad: unsupported shape or device");
    int rc = scale_buffer(dw_all, (int64_t)groups * f * k, beta, s); if (rc) return rc;
    if (dbias_all) { rc = colsum_bf16(dy, n, groups * f, dbias_all, beta, s); if (rc) return rc; }
    return wgrad_rows(n, f, k, groups, x_shared, x, dy, dw_all, s);
}

int umma_conv_wgrad(const cb_conv_desc* d, const void* x, const void* dy, const float* mean, const float* rstd,
                    float* dw, float* dbias, float beta, cudaStream_t s) {
    CB_REQUIRE(x && dy && dw, CB_E_ARG, "cb_conv_wgrad: null pointer");
    CB_REQUIRE(!mean && !rstd, CB_E_ARG, "cb_conv_wgrad: normalised uint8 input has no tcgen05 path");
    const WgradKind kind = wgrad_kind(d);
    const long long K = (long long)d->k_h * d->k_w * d->in_c;
    int rc = scale_buffer(dw, (int64_t)d->out_c * K, beta, s); if (rc) return rc;
    if (dbias) {
        rc = colsum_bf16(dy, (int64_t)d->n * d->out_h * d->out_w, d->out_c, dbias, beta, s); if (rc) return rc;
    }
    umma::WgradParams p{};
    CUtensorMap tmA, tmB;
    p.dw = dw;
    p.unit_div = 1;
    if (kind == WG_ROWS) return wgrad_rows(d->n, d->out_c, (int)K, 1, 0, x, dy, dw, s);
    if (kind == WG_CONV_S1) {
        const int k = d->k_h, gh = d->in_h + 2 * d->pad, gw = d->in_w + 2 * d->pad;
        const int S = 256 / (gh * gw) > 0 ? 256 / (gh * gw) : 1;
        const int rows = S * gh * gw;
        uint64_t dx4[4] = {64, (uint64_t)d->in_w, (uint64_t)d->in_h, (uint64_t)d->n};
        uint64_t sx[3] = {128, (uint64_t)d->in_w * 128, (uint64_t)d->in_h * d->in_w * 128};
        uint64_t dy4[4] = {64, (uint64_t)d->out_w, (uint64_t)d->out_h, (uint64_t)d->n};
        uint64_t sy[3] = {128, (uint64_t)d->out_w * 128, (uint64_t)d->out_h * d->out_w * 128};
        uint32_t box[4] = {64, (uint32_t)gw, (uint32_t)gh, (uint32_t)S};
        rc = encode_bf16(&tmA, x, 4, dx4, sx, box); if (rc) return rc;
        rc = encode_bf16(&tmB, dy, 4, dy4, sy, box); if (rc) return rc;     // beyond out_h/out_w: zero fill
        p.na = 1; p.nb = 1; p.a_rank = p.b_rank = 4;
        p.a_coord[0][1] = -d->pad; p.a_coord[0][2] = -d->pad;
        p.a_step_dim = p.b_step_dim = 3; p.a_step = p.b_step = S;
        p.a_unit_dim = p.b_unit_dim = 0; p.a_unit_step = p.b_unit_step = 0;
        p.k_rows = round_up(rows, 16);
        const int max_shift = (k - 1) * gw + (k - 1);
        const int a_region = round_up((p.k_rows + max_shift + 1) * 128, 1024);
        const int b_region = round_up(p.k_rows * 128, 1024);
        p.a_box_bytes = p.b_box_bytes = rows * 128;
        p.a_box_stride = a_region; p.b_box_stride = b_region;
        p.b_offset = a_region; p.stage_bytes = a_region + b_region; p.slack_bytes = 0;
        p.stages = (kSmemBudget - 1536) / p.stage_bytes; if (p.stages > 4) p.stages = 4;
        p.units = 1; p.slabs_total = (d->n + S - 1) / S;
        p.splits = p.slabs_total < sm_count() ? p.slabs_total : sm_count();
        const int taps = k * k;
        p.n_acc = (taps + 1) / 2;
        for (int g = 0; g < p.n_acc; ++g) {
            const int t0 = 2 * g, t1 = 2 * g + 1;
            const int s0 = (t0 / k) * gw + (t0 % k);
            p.acc_a_off[g] = s0 * 128;
            p.acc_base0[g] = (long long)t0 * 64;
            if (t1 < taps) {
                const int s1 = (t1 / k) * gw + (t1 % k);
                p.acc_a_lbo[g] = (s1 - s0) * 128;
                p.acc_base1[g] = (long long)t1 * 64;
            } else {
                p.acc_a_lbo[g] = 128;            // upper half computed from a neighbouring window and discarded
                p.acc_base1[g] = -1;
            }
        }
        p.m_stride = 1; p.n_stride = K;
        return launch_wgrad<64>(tmA, tmB, p, s);
    }
    if (kind == WG_CONV_S2) {
        const int gh = d->in_h / 2, gw = d->in_w / 2;
        const int S = 2 * gh * gw <= 208 ? 2 : 1;
        const int rows = S * gh * gw;
        uint64_t dx5[5] = {64, (uint64_t)gw, 2, (uint64_t)gh, (uint64_t)d->n};
        uint64_t sx[4] = {128, (uint64_t)d->in_w * 64, (uint64_t)d->in_w * 128, (uint64_t)d->in_h * d->in_w * 64};
        uint32_t boxa[5] = {64, (uint32_t)gw, 1, (uint32_t)gh, (uint32_t)S};
        uint64_t dy4[4] = {64, (uint64_t)d->out_w, (uint64_t)d->out_h, (uint64_t)d->n};
        uint64_t sy[3] = {128, (uint64_t)d->out_w * 128, (uint64_t)d->out_h * d->out_w * 128};
        uint32_t boxb[4] = {64, (uint32_t)gw, (uint32_t)gh, (uint32_t)S};
        rc = encode_bf16(&tmA, x, 5, dx5, sx, boxa); if (rc) return rc;
        rc = encode_bf16(&tmB, dy, 4, dy4, sy, boxb); if (rc) return rc;
        p.na = 2; p.nb = 1; p.a_rank = 5; p.b_rank = 4;
        p.a_coord[1][2] = 1;
        p.a_step_dim = 4; p.b_step_dim = 3; p.a_step = p.b_step = S;
        p.a_unit_dim = p.b_unit_dim = 0; p.a_unit_step = p.b_unit_step = 0;
        p.k_rows = round_up(rows, 16);
        const int max_shift = gw + 1;
        const int a_region = round_up((p.k_rows + max_shift + 1) * 128, 1024);
        const int b_region = round_up(p.k_rows * 128, 1024);
        p.a_box_bytes = p.b_box_bytes = rows * 128;
        p.a_box_stride = a_region; p.b_box_stride = b_region;
        p.b_offset = 2 * a_region; p.stage_bytes = 2 * a_region + b_region; p.slack_bytes = 0;
        p.stages = (kSmemBudget - 1536) / p.stage_bytes; if (p.stages > 4) p.stages = 4;
        p.units = 1; p.slabs_total = (d->n + S - 1) / S;
        p.splits = p.slabs_total < sm_count() ? p.slabs_total : sm_count();
        p.n_acc = 4;
        for (int kh = 0; kh < 4; ++kh) {
            p.acc_a_off[kh] = (kh & 1) * a_region + (kh >> 1) * gw * 128;
            p.acc_a_lbo[kh] = 128;                      // kw pairs j = 0, 1 are one grid column apart
            p.acc_base0[kh] = kh * 128; p.acc_base1[kh] = kh * 128 + 64;
        }
        p.m_stride = 1; p.n_stride = K;
        return launch_wgrad<64>(tmA, tmB, p, s);
    }
    return fail(CB_E_ARG, "cb_conv_wgrad: no tcgen05 specialisation for this shape");
}

}  // namespace cb

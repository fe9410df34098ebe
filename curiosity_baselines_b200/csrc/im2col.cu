// Explicit im2col / col2im for convolution shapes that have no implicit-GEMM specialisation on the tcgen05 engine
// (UniverseHead's 3x3 stride-2 pad-1 layers, encoders.py:7-35; the 16/32-channel policy encoder of
// models/conv2d.py:67-118).  The patch matrix goes through HBM once per pass, so these layers run their
// MACs on the tensor cores through the row GEMM kernels (umma_gemm.cuh) at the price of extra traffic:
//   forward   col = im2col(x)            y  = act(col W^T + b)            (cb_conv_fwd on [M, K_pad] rows)
//   wgrad     col = im2col(x)            dW = dy^T col                    (cb_conv_wgrad on rows)
//   dgrad     dcol = dy W                dx = col2im(dcol) .* act'(x)     (cb_conv_dgrad on rows + this file)
// Both kernels are HBM-bound gathers: one thread per 8 consecutive k (16-byte stores / loads).
#include "common.cuh"

namespace {

struct ColGeo {
    int in_h, in_w, in_c, k_h, k_w, stride, pad, out_h, out_w;
    int64_t sn, sh, sw, sc;       // element strides of x
    int K, k_pad;
};

__device__ __forceinline__ float load_x(const void* x, int dtype, int64_t off) {
    return dtype == CB_U8 ? (float)static_cast<const uint8_t*>(x)[off]
                          : cb::bf2f(static_cast<const __nv_bfloat16*>(x)[off]);
}

// col[(n - n0)*P + p][k] for k = (kh*KW + kw)*C + ci; zero outside the image and for k >= K
template <bool VEC>
__global__ void im2col_kernel(const void* __restrict__ x, int dtype, ColGeo g, __nv_bfloat16* __restrict__ col,
                              int n0, int64_t rows) {
    const int chunks = g.k_pad / 8;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * chunks) return;
    const int64_t row = idx / chunks;
    const int k0 = (int)(idx - row * chunks) * 8;
    const int P = g.out_h * g.out_w;
    const int n = n0 + (int)(row / P), p = (int)(row % P);
    const int oy = p / g.out_w, ox = p - oy * g.out_w;
    uint4 out = make_uint4(0, 0, 0, 0);
    if (VEC) {      // bf16, channel stride 1, C % 8 == 0: the 8 k are 8 channels of one tap
        if (k0 < g.K) {
            const int tap = k0 / g.in_c, c0 = k0 - tap * g.in_c;
            const int kh = tap / g.k_w, kw = tap - kh * g.k_w;
            const int iy = oy * g.stride + kh - g.pad, ix = ox * g.stride + kw - g.pad;
            if (iy >= 0 && iy < g.in_h && ix >= 0 && ix < g.in_w)
                out = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(x) + n * g.sn + iy * g.sh +
                                                           ix * g.sw + c0));
        }
    } else {
        __nv_bfloat16 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + j;
            float f = 0.f;
            if (k < g.K) {
                const int tap = k / g.in_c, ci = k - tap * g.in_c;
                const int kh = tap / g.k_w, kw = tap - kh * g.k_w;
                const int iy = oy * g.stride + kh - g.pad, ix = ox * g.stride + kw - g.pad;
                if (iy >= 0 && iy < g.in_h && ix >= 0 && ix < g.in_w)
                    f = load_x(x, dtype, n * g.sn + iy * g.sh + ix * g.sw + ci * g.sc);
            }
            v[j] = cb::f2bf(f);
        }
        out = *reinterpret_cast<uint4*>(v);
    }
    *reinterpret_cast<uint4*>(col + row * g.k_pad + k0) = out;
}

// scalar form for channel counts that are not a multiple of 8 (a first layer's 3 or 4 channels)
__global__ void col2im_scalar_kernel(const __nv_bfloat16* __restrict__ dcol, ColGeo g,
                                     const __nv_bfloat16* __restrict__ x_saved, int x_act,
                                     const __nv_bfloat16* __restrict__ dx_add, __nv_bfloat16* __restrict__ dx,
                                     int64_t dn, int64_t dh, int64_t dw, int n0, int64_t pixels) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= pixels * g.in_c) return;
    const int64_t pix = idx / g.in_c;
    const int c = (int)(idx - pix * g.in_c);
    const int hw = g.in_h * g.in_w;
    const int nl = (int)(pix / hw), r = (int)(pix % hw);
    const int iy = r / g.in_w, ix = r - iy * g.in_w;
    const int P = g.out_h * g.out_w;
    float acc = 0.f;
    for (int kh = 0; kh < g.k_h; ++kh) {
        const int ty = iy + g.pad - kh;
        if (ty < 0 || ty % g.stride || ty / g.stride >= g.out_h) continue;
        for (int kw = 0; kw < g.k_w; ++kw) {
            const int tx = ix + g.pad - kw;
            if (tx < 0 || tx % g.stride || tx / g.stride >= g.out_w) continue;
            acc += cb::bf2f(dcol[((int64_t)nl * P + (ty / g.stride) * g.out_w + tx / g.stride) * g.k_pad +
                                 (kh * g.k_w + kw) * g.in_c + c]);
        }
    }
    const int64_t o = (int64_t)(n0 + nl) * dn + iy * dh + ix * dw + c;
    if (dx_add) acc += cb::bf2f(dx_add[o]);
    if (x_saved && x_act != CB_ACT_NONE) acc *= cb::act_grad_from_output(cb::bf2f(x_saved[o]), x_act);
    dx[o] = cb::f2bf(acc);
}

// dx[n, iy, ix, c0..c0+7] = (sum over the taps that read this pixel of dcol + dx_add) * act'(x_saved)
__global__ void col2im_kernel(const __nv_bfloat16* __restrict__ dcol, ColGeo g, const __nv_bfloat16* __restrict__ x_saved,
                              int x_act, const __nv_bfloat16* __restrict__ dx_add, __nv_bfloat16* __restrict__ dx,
                              int64_t dn, int64_t dh, int64_t dw, int n0, int64_t pixels) {
    const int cch = g.in_c / 8;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= pixels * cch) return;
    const int64_t pix = idx / cch;
    const int c0 = (int)(idx - pix * cch) * 8;
    const int hw = g.in_h * g.in_w;
    const int nl = (int)(pix / hw), r = (int)(pix % hw);
    const int iy = r / g.in_w, ix = r - iy * g.in_w;
    const int P = g.out_h * g.out_w;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int kh = 0; kh < g.k_h; ++kh) {
        const int ty = iy + g.pad - kh;
        if (ty < 0 || ty % g.stride) continue;
        const int oy = ty / g.stride;
        if (oy >= g.out_h) continue;
        for (int kw = 0; kw < g.k_w; ++kw) {
            const int tx = ix + g.pad - kw;
            if (tx < 0 || tx % g.stride) continue;
            const int ox = tx / g.stride;
            if (ox >= g.out_w) continue;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(
                dcol + ((int64_t)nl * P + oy * g.out_w + ox) * g.k_pad + (kh * g.k_w + kw) * g.in_c + c0));
            const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += cb::bf2f(b[j]);
        }
    }
    const int64_t o = (int64_t)(n0 + nl) * dn + iy * dh + ix * dw + c0;
    if (dx_add) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(dx_add + o));
        const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += cb::bf2f(b[j]);
    }
    if (x_saved && x_act != CB_ACT_NONE) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x_saved + o));
        const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= cb::act_grad_from_output(cb::bf2f(b[j]), x_act);
    }
    __nv_bfloat16 outv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) outv[j] = cb::f2bf(acc[j]);
    *reinterpret_cast<uint4*>(dx + o) = *reinterpret_cast<uint4*>(outv);
}

int fill_geo(ColGeo& g, const cb_conv_desc* d, int k_pad) {
    CB_REQUIRE(d && d->in_c >= 1 && d->k_h >= 1 && d->k_w >= 1 && d->stride >= 1 && d->pad >= 0, CB_E_ARG,
               "im2col: bad descriptor");
    CB_REQUIRE(d->out_h == (d->in_h + 2 * d->pad - d->k_h) / d->stride + 1 &&
               d->out_w == (d->in_w + 2 * d->pad - d->k_w) / d->stride + 1, CB_E_ARG, "im2col: inconsistent output extent");
    g.in_h = d->in_h; g.in_w = d->in_w; g.in_c = d->in_c; g.k_h = d->k_h; g.k_w = d->k_w; g.stride = d->stride;
    g.pad = d->pad; g.out_h = d->out_h; g.out_w = d->out_w;
    g.sn = d->in_stride_n; g.sh = d->in_stride_h; g.sw = d->in_stride_w; g.sc = d->in_stride_c;
    g.K = d->k_h * d->k_w * d->in_c; g.k_pad = k_pad;
    CB_REQUIRE(k_pad >= g.K && k_pad % 8 == 0, CB_E_ARG, "im2col: k_pad=%d must be a multiple of 8 and >= K=%d", k_pad, g.K);
    return CB_OK;
}

}  // namespace

extern "C" int cb_im2col(const cb_conv_desc* d, const void* x, void* col, int32_t k_pad, int32_t n0, int32_t n1,
                         void* stream) {
    CB_REQUIRE(x && col && n0 >= 0 && n1 > n0, CB_E_ARG, "cb_im2col: bad argument");
    CB_REQUIRE(d && (d->in_dtype == CB_U8 || d->in_dtype == CB_BF16), CB_E_ARG, "cb_im2col: input must be uint8 or bf16");
    CB_REQUIRE(((uintptr_t)col & 15) == 0, CB_E_ALIGN, "cb_im2col: col not 16-byte aligned");
    ColGeo g;
    int rc = fill_geo(g, d, k_pad); if (rc) return rc;
    const int64_t rows = (int64_t)(n1 - n0) * g.out_h * g.out_w;
    const int64_t total = rows * (k_pad / 8);
    const unsigned blocks = (unsigned)cb::ceil_div<int64_t>(total, 256);
    const bool vec = d->in_dtype == CB_BF16 && g.sc == 1 && g.in_c % 8 == 0 && ((uintptr_t)x & 15) == 0 &&
                     g.sn % 8 == 0 && g.sh % 8 == 0 && g.sw % 8 == 0;
    if (vec) im2col_kernel<true><<<blocks, 256, 0, cb::as_stream(stream)>>>(x, d->in_dtype, g, (__nv_bfloat16*)col, n0, rows);
    else im2col_kernel<false><<<blocks, 256, 0, cb::as_stream(stream)>>>(x, d->in_dtype, g, (__nv_bfloat16*)col, n0, rows);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_col2im(const cb_conv_desc* d, const void* dcol, int32_t k_pad, const void* x_saved, int32_t x_act,
                         const void* dx_add, void* dx, int64_t dx_stride_n, int64_t dx_stride_h, int64_t dx_stride_w,
                         int32_t n0, int32_t n1, void* stream) {
    CB_REQUIRE(dcol && dx && n0 >= 0 && n1 > n0, CB_E_ARG, "cb_col2im: bad argument");
    ColGeo g;
    int rc = fill_geo(g, d, k_pad); if (rc) return rc;
    const int64_t pixels = (int64_t)(n1 - n0) * g.in_h * g.in_w;
    const bool vec = g.in_c % 8 == 0 && (((uintptr_t)dcol | (uintptr_t)dx | (uintptr_t)x_saved | (uintptr_t)dx_add) & 15) == 0 &&
                     dx_stride_n % 8 == 0 && dx_stride_h % 8 == 0 && dx_stride_w % 8 == 0;
    if (!vec) {
        col2im_scalar_kernel<<<(unsigned)cb::ceil_div<int64_t>(pixels * g.in_c, 256), 256, 0, cb::as_stream(stream)>>>(
            (const __nv_bfloat16*)dcol, g, (const __nv_bfloat16*)x_saved, x_act, (const __nv_bfloat16*)dx_add,
            (__nv_bfloat16*)dx, dx_stride_n, dx_stride_h, dx_stride_w, n0, pixels);
        CB_LAUNCH_CHECK();
        return CB_OK;
    }
    const int64_t total = pixels * (g.in_c / 8);
    col2im_kernel<<<(unsigned)cb::ceil_div<int64_t>(total, 256), 256, 0, cb::as_stream(stream)>>>(
        (const __nv_bfloat16*)dcol, g, (const __nv_bfloat16*)x_saved, x_act, (const __nv_bfloat16*)dx_add,
        (__nv_bfloat16*)dx, dx_stride_n, dx_stride_h, dx_stride_w, n0, pixels);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

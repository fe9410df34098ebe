// Generic CUDA-core implicit-GEMM engine (CB_ENGINE_SIMT): any filter size / stride / padding,
// uint8-NCHW or bf16-NHWC input.  It is the shape-complete path (UniverseHead, MazeHead, NDIGO
// layers, odd sizes) and the on-device cross-check for the tcgen05 kernels in umma_*.cu, which
// take over the Burda/RND shapes.  Operands are rounded to bf16 and accumulated in fp32, i.e. the
// same numerics contract as the tensor-core path.
//   forward  : encoders.py:7-119, icm.py:9-43,88-92, rnd.py:54-117
//   backward : what torch.autograd derives for those modules (dgrad / wgrad / bias grad)
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

struct GemmParams {
    cb_conv_desc d;
    const void* x;              // forward input (u8 / bf16) or dy for dgrad
    const __nv_bfloat16* w;     // [N][K] K-contiguous
    const __nv_bfloat16* dy;    // wgrad: [positions][out_c]
    cb_epilogue e;
    // dgrad
    const __nv_bfloat16* x_saved; int x_act; const __nv_bfloat16* dx_add; __nv_bfloat16* dx;
    // wgrad
    float* dw; int64_t k_per_split;
    int64_t M, N, K;
};

__device__ __forceinline__ float round_bf16(float v) { return cb::bf2f(cb::f2bf(v)); }

// value of the forward input at (n, ih, iw, ci); zero outside the image
__device__ __forceinline__ float load_input(const cb_conv_desc& d, const void* x, const float* mean,
                                            const float* rstd, int n, int ih, int iw, int ci) {
    if (ih < 0 || ih >= d.in_h || iw < 0 || iw >= d.in_w) return 0.f;
    int64_t off = n * d.in_stride_n + ih * d.in_stride_h + iw * d.in_stride_w + ci * d.in_stride_c;
    if (d.in_dtype == CB_U8) {
        float v = (float)reinterpret_cast<const uint8_t*>(x)[off];
        if (mean) {
            int px = ih * d.in_w + iw;
            v = (v - mean[px]) * rstd[px];
            v = fminf(fmaxf(v, -5.f), 5.f);
            v = round_bf16(v);
        }
        return v;
    }
    return cb::bf2f(reinterpret_cast<const __nv_bfloat16*>(x)[off]);
}

enum { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

template <int MODE>
__device__ __forceinline__ float load_a(const GemmParams& p, int64_t m, int64_t k) {
    const cb_conv_desc& d = p.d;
    if (MODE == MODE_FWD) {
        int ci = (int)(k % d.in_c); int64_t r = k / d.in_c;
        int kw = (int)(r % d.k_w); int kh = (int)(r / d.k_w);
        int ow = (int)(m % d.out_w); int64_t q = m / d.out_w;
        int oh = (int)(q % d.out_h); int n = (int)(q / d.out_h);
        return load_input(d, p.x, p.e.norm_mean, p.e.norm_rstd, n, oh * d.stride - d.pad + kh,
                          ow * d.stride - d.pad + kw, ci);
    } else if (MODE == MODE_DGRAD) {
        int co = (int)(k % d.out_c); int64_t r = k / d.out_c;
        int kw = (int)(r % d.k_w); int kh = (int)(r / d.k_w);
        int iw = (int)(m % d.in_w); int64_t q = m / d.in_w;
        int ih = (int)(q % d.in_h); int n = (int)(q / d.in_h);
        int a = ih + d.pad - kh, b = iw + d.pad - kw;
        if (a < 0 || b < 0 || a % d.stride || b % d.stride) return 0.f;
        int oh = a / d.stride, ow = b / d.stride;
        if (oh >= d.out_h || ow >= d.out_w) return 0.f;
        return cb::bf2f(p.dy[(((int64_t)n * d.out_h + oh) * d.out_w + ow) * d.out_c + co]);
    } else {  // WGRAD: A[m=co][k=position]
        return cb::bf2f(p.dy[k * d.out_c + m]);
    }
}

template <int MODE>
__device__ __forceinline__ float load_b(const GemmParams& p, int64_t n, int64_t k) {
    const cb_conv_desc& d = p.d;
    if (MODE == MODE_WGRAD) {   // B[n=patch index][k=position] = im2col(position, patch index)
        int ci = (int)(n % d.in_c); int64_t r = n / d.in_c;
        int kw = (int)(r % d.k_w); int kh = (int)(r / d.k_w);
        int ow = (int)(k % d.out_w); int64_t q = k / d.out_w;
        int oh = (int)(q % d.out_h); int s = (int)(q / d.out_h);
        return load_input(d, p.x, p.e.norm_mean, p.e.norm_rstd, s, oh * d.stride - d.pad + kh,
                          ow * d.stride - d.pad + kw, ci);
    }
    return cb::bf2f(p.w[n * p.K + k]);
}

template <int MODE>
__global__ void __launch_bounds__(THREADS) simt_gemm_kernel(GemmParams p) {
    __shared__ float As[BK][BM + 1];
    __shared__ float Bs[BK][BN + 1];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
    int64_t kbeg = 0, kend = p.K;
    if (MODE == MODE_WGRAD) {
        kbeg = (int64_t)blockIdx.z * p.k_per_split;
        kend = min(p.K, kbeg + p.k_per_split);
    }
    const int tx = tid % 16, ty = tid / 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int i = 0; i < (BM * BK) / THREADS; ++i) {
            int e = tid + i * THREADS;
            int mm, kk;
            if (MODE == MODE_WGRAD) { mm = e % BM; kk = e / BM; } else { mm = e / BK; kk = e % BK; }
            int64_t m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < p.M && k < kend) ? load_a<MODE>(p, m, k) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < (BN * BK) / THREADS; ++i) {
            int e = tid + i * THREADS;
            int nn, kk;
            if (MODE == MODE_WGRAD) { nn = e % BN; kk = e / BN; } else { nn = e / BK; kk = e % BK; }
            int64_t n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < p.N && k < kend) ? load_b<MODE>(p, n, k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            if (MODE == MODE_FWD) {
                if (p.e.bias) v += p.e.bias[n];
                for (int s = 0; s < p.e.side_dim; ++s)
                    v = fmaf(p.e.side[m * p.e.side_dim + s], p.e.side_w[n * p.e.side_dim + s], v);
                v = cb::apply_act(v, p.d.act);
                if (p.e.residual)
                    v += cb::bf2f(reinterpret_cast<const __nv_bfloat16*>(p.e.residual)[m * p.N + n]);
                if (p.e.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.e.out_bf16)[m * p.N + n] = cb::f2bf(v);
                if (p.e.out_f32) p.e.out_f32[m * p.N + n] = v;
            } else if (MODE == MODE_DGRAD) {
                if (p.dx_add) v += cb::bf2f(p.dx_add[m * p.N + n]);
                if (p.x_saved) v *= cb::act_grad_from_output(cb::bf2f(p.x_saved[m * p.N + n]), p.x_act);
                p.dx[m * p.N + n] = cb::f2bf(v);
            } else {
                atomicAdd(p.dw + m * p.N + n, v);
            }
        }
    }
}

__global__ void scale_kernel(float* __restrict__ x, int64_t n, float beta) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = beta == 0.f ? 0.f : x[i] * beta;
}

// column sums of dy [rows][c] -> dbias[c] (atomic over row chunks)
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ dy, int64_t rows, int c,
                              float* __restrict__ out, int64_t rows_per_block) {
    int col = blockIdx.y * blockDim.x + threadIdx.x;
    if (col >= c) return;
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s += cb::bf2f(dy[r * c + col]);
    atomicAdd(out + col, s);
}

// Vectorised column sums for c % 8 == 0 and c <= 2048: c/8 threads cover one row with 16-byte loads,
// 256/(c/8) rows per sweep; per-thread fp32 partials are folded through shared memory, one atomic
// per column per block.  HBM-bound (reads dy once).
__global__ void __launch_bounds__(256) colsum_vec_kernel(const __nv_bfloat16* __restrict__ dy, int64_t rows, int c,
                                                         float* __restrict__ out, int64_t rows_per_block) {
    const int tpr = c >> 3;                         // threads per row
    const int rps = 256 / tpr;                      // rows per sweep
    const int tr = threadIdx.x / tpr, tc = threadIdx.x - tr * tpr;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (tr < rps) {
        // four rows per iteration, loads first: one 16-byte load in flight per thread left the kernel latency-bound
        // (2 TB/s); the accumulation order per thread is unchanged
        for (int64_t r = r0 + tr; r < r1; r += 4 * (int64_t)rps) {
            uint4 u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t rr = r + (int64_t)k * rps;
                u[k] = rr < r1 ? __ldg(reinterpret_cast<const uint4*>(dy + rr * c) + tc) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[k]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h[e]);
                    acc[2 * e] += f.x; acc[2 * e + 1] += f.y;
                }
            }
        }
    }
    __shared__ float sm[256 * 8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[threadIdx.x * 8 + j] = acc[j];
    __syncthreads();
    for (int col = threadIdx.x; col < c; col += 256) {
        const int t = col >> 3, j = col & 7;
        float s = 0.f;
        for (int rr = 0; rr < rps; ++rr) s += sm[(rr * tpr + t) * 8 + j];
        atomicAdd(out + col, s);
    }
}

// dsw[n][s] += sum_m dy[m][n]*side[m][s]
__global__ void side_wgrad_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ side,
                                  float* __restrict__ dsw, int64_t m, int out_c, int side_dim,
                                  int64_t rows_per_block) {
    int n = blockIdx.y * blockDim.x + threadIdx.x;
    if (n >= out_c) return;
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(m, r0 + rows_per_block);
    for (int s = 0; s < side_dim; ++s) {
        float acc = 0.f;
        for (int64_t r = r0; r < r1; ++r) acc = fmaf(cb::bf2f(dy[r * out_c + n]), side[r * side_dim + s], acc);
        atomicAdd(dsw + (int64_t)n * side_dim + s, acc);
    }
}

// Same product with each thread owning 8 consecutive output columns (one 16-byte load per row) and 8 side columns
// at a time in registers: dy is read once per 8 side inputs instead of once per side input, with 2-byte loads.
// block = (cgs column groups) x (256 / cgs row lanes); grid = (row blocks, column-group blocks, side chunks of 8).
__global__ void side_wgrad_vec_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ side,
                                      float* __restrict__ dsw, int64_t m, int out_c, int side_dim, int64_t rows_per_block) {
    const int cg = blockIdx.y * blockDim.x + threadIdx.x;             // 8-column group
    if (cg * 8 >= out_c) return;
    const int s0 = blockIdx.z * 8;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(m, r0 + rows_per_block);
    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[j][k] = 0.f;
    // two rows per iteration with all loads issued first (the single-row loop was latency-bound at ~0.8 TB/s)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 2 * (int64_t)blockDim.y) {
        uint4 v[2];
        float sv[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t rr = r + (int64_t)u * blockDim.y;
            const bool ok = rr < r1;
            v[u] = ok ? __ldg(reinterpret_cast<const uint4*>(dy + rr * out_c) + cg) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 8; ++k) sv[u][k] = (ok && s0 + k < side_dim) ? __ldg(side + rr * side_dim + s0 + k) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
            float x[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h[e]); x[2 * e] = f.x; x[2 * e + 1] = f.y; }
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[j][k] = fmaf(x[j], sv[u][k], acc[j][k]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (s0 + k < side_dim && acc[j][k] != 0.f)
                atomicAdd(dsw + (int64_t)(cg * 8 + j) * side_dim + s0 + k, acc[j][k]);
}

int check_desc(const cb_conv_desc* d, const char* who) {
    CB_REQUIRE(d, CB_E_ARG, "%s: null descriptor", who);
    CB_REQUIRE(d->n >= 1 && d->in_h >= 1 && d->in_w >= 1 && d->in_c >= 1 && d->k_h >= 1 && d->k_w >= 1 &&
               d->stride >= 1 && d->pad >= 0 && d->out_c >= 1, CB_E_ARG, "%s: bad extent", who);
    int oh = (d->in_h + 2 * d->pad - d->k_h) / d->stride + 1;
    int ow = (d->in_w + 2 * d->pad - d->k_w) / d->stride + 1;
    CB_REQUIRE(oh == d->out_h && ow == d->out_w, CB_E_ARG, "%s: out extent %dx%d != expected %dx%d", who,
               d->out_h, d->out_w, oh, ow);
    CB_REQUIRE(d->in_dtype == CB_U8 || d->in_dtype == CB_BF16, CB_E_ARG, "%s: in_dtype %d", who, d->in_dtype);
    return CB_OK;
}

}  // namespace

namespace cb {

int colsum_bf16(const void* dy, int64_t rows, int c, float* out, float beta, cudaStream_t s);

// x = beta * x (beta == 0 stores zeros); used to pre-scale atomically accumulated outputs
int scale_buffer(float* x, int64_t n, float beta, cudaStream_t s) {
    scale_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, s>>>(x, n, beta);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// out[c] = beta*out[c] + sum_rows dy[rows][c]   (bias gradient)
int colsum_bf16(const void* dy, int64_t rows, int c, float* out, float beta, cudaStream_t s) {
    int rc = scale_buffer(out, c, beta, s); if (rc) return rc;
    if (c % 8 == 0 && c <= 2048 && (((uintptr_t)dy) & 15) == 0) {
        int64_t rpb = std::max<int64_t>(256, ceil_div<int64_t>(rows, 148 * 8));
        colsum_vec_kernel<<<(unsigned)ceil_div<int64_t>(rows, rpb), 256, 0, s>>>((const __nv_bfloat16*)dy, rows, c, out, rpb);
        CB_LAUNCH_CHECK();
        return CB_OK;
    }
    int64_t rpb = std::max<int64_t>(64, ceil_div<int64_t>(rows, 148 * 4));
    dim3 g2((unsigned)ceil_div<int64_t>(rows, rpb), (unsigned)ceil_div(c, 128));
    colsum_kernel<<<g2, 128, 0, s>>>((const __nv_bfloat16*)dy, rows, c, out, rpb);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int simt_conv_fwd(const cb_conv_desc* d, const void* x, const void* w, const cb_epilogue* e, cudaStream_t s) {
    int rc = check_desc(d, "cb_conv_fwd"); if (rc) return rc;
    CB_REQUIRE(x && w && e && (e->out_bf16 || e->out_f32), CB_E_ARG, "cb_conv_fwd: null pointer");
    CB_REQUIRE(e->side_dim == 0 || (e->side && e->side_w), CB_E_ARG, "cb_conv_fwd: side input missing");
    GemmParams p{};
    p.d = *d; p.x = x; p.w = (const __nv_bfloat16*)w; p.e = *e;
    p.M = (int64_t)d->n * d->out_h * d->out_w; p.N = d->out_c; p.K = (int64_t)d->k_h * d->k_w * d->in_c;
    dim3 grid((unsigned)ceil_div<int64_t>(p.M, BM), (unsigned)ceil_div<int64_t>(p.N, BN));
    simt_gemm_kernel<MODE_FWD><<<grid, THREADS, 0, s>>>(p);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int simt_conv_dgrad(const cb_conv_desc* d, const void* dy, const void* w_t, const void* x_saved, int x_act,
                    const void* dx_add, void* dx, cudaStream_t s) {
    int rc = check_desc(d, "cb_conv_dgrad"); if (rc) return rc;
    CB_REQUIRE(dy && w_t && dx, CB_E_ARG, "cb_conv_dgrad: null pointer");
    GemmParams p{};
    p.d = *d; p.dy = (const __nv_bfloat16*)dy; p.w = (const __nv_bfloat16*)w_t;
    p.x_saved = (const __nv_bfloat16*)x_saved; p.x_act = x_act; p.dx_add = (const __nv_bfloat16*)dx_add; p.dx = (__nv_bfloat16*)dx;
    p.M = (int64_t)d->n * d->in_h * d->in_w; p.N = d->in_c; p.K = (int64_t)d->k_h * d->k_w * d->out_c;
    dim3 grid((unsigned)ceil_div<int64_t>(p.M, BM), (unsigned)ceil_div<int64_t>(p.N, BN));
    simt_gemm_kernel<MODE_DGRAD><<<grid, THREADS, 0, s>>>(p);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int simt_conv_wgrad(const cb_conv_desc* d, const void* x, const void* dy, const float* mean, const float* rstd,
                    float* dw, float* dbias, float beta, cudaStream_t s) {
    int rc = check_desc(d, "cb_conv_wgrad"); if (rc) return rc;
    CB_REQUIRE(x && dy && dw, CB_E_ARG, "cb_conv_wgrad: null pointer");
    GemmParams p{};
    p.d = *d; p.x = x; p.dy = (const __nv_bfloat16*)dy; p.dw = dw;
    p.e.norm_mean = mean; p.e.norm_rstd = rstd;
    p.M = d->out_c; p.N = (int64_t)d->k_h * d->k_w * d->in_c; p.K = (int64_t)d->n * d->out_h * d->out_w;
    int64_t tiles = ceil_div<int64_t>(p.M, BM) * ceil_div<int64_t>(p.N, BN);
    int64_t want = std::max<int64_t>(1, (148 * 4) / tiles);
    int64_t kper = std::max<int64_t>(BK * 8, ceil_div<int64_t>(ceil_div<int64_t>(p.K, want), BK) * BK);
    int splits = (int)ceil_div<int64_t>(p.K, kper);
    p.k_per_split = kper;
    int64_t nw = p.M * p.N;
    scale_kernel<<<(unsigned)ceil_div<int64_t>(nw, 256), 256, 0, s>>>(dw, nw, beta);
    CB_LAUNCH_CHECK();
    dim3 grid((unsigned)ceil_div<int64_t>(p.M, BM), (unsigned)ceil_div<int64_t>(p.N, BN), (unsigned)splits);
    simt_gemm_kernel<MODE_WGRAD><<<grid, THREADS, 0, s>>>(p);
    CB_LAUNCH_CHECK();
    if (dbias) return colsum_bf16(dy, p.K, d->out_c, dbias, beta, s);
    return CB_OK;
}

}  // namespace cb

extern "C" int cb_side_wgrad(const void* dy, const float* side, float* dsw, int64_t m, int32_t out_c,
                             int32_t side_dim, float beta, void* stream) {
    CB_REQUIRE(dy && side && dsw && m >= 1 && out_c >= 1 && side_dim >= 1, CB_E_ARG, "cb_side_wgrad: bad argument");
    cudaStream_t s = cb::as_stream(stream);
    int64_t nw = (int64_t)out_c * side_dim;
    scale_kernel<<<(unsigned)cb::ceil_div<int64_t>(nw, 256), 256, 0, s>>>(dsw, nw, beta);
    CB_LAUNCH_CHECK();
    if (out_c % 8 == 0 && (((uintptr_t)dy) & 15) == 0) {
        const int cgs_total = out_c / 8;
        const int bx = cgs_total < 256 ? cgs_total : 256;                  // column groups per block
        int by = 256 / bx; if (by < 1) by = 1;                             // row lanes per block
        const int col_blocks = cb::ceil_div(cgs_total, bx), side_chunks = cb::ceil_div(side_dim, 8);
        int64_t row_blocks = std::max<int64_t>(1, (148 * 4) / ((int64_t)col_blocks * side_chunks));
        int64_t rpb = std::max<int64_t>(64, cb::ceil_div<int64_t>(m, row_blocks));
        dim3 grid((unsigned)cb::ceil_div<int64_t>(m, rpb), (unsigned)col_blocks, (unsigned)side_chunks);
        side_wgrad_vec_kernel<<<grid, dim3(bx, by), 0, s>>>((const __nv_bfloat16*)dy, side, dsw, m, out_c, side_dim, rpb);
        CB_LAUNCH_CHECK();
        return CB_OK;
    }
    int64_t rpb = std::max<int64_t>(64, cb::ceil_div<int64_t>(m, 148 * 2));
    dim3 grid((unsigned)cb::ceil_div<int64_t>(m, rpb), (unsigned)cb::ceil_div(out_c, 128));
    side_wgrad_kernel<<<grid, 128, 0, s>>>((const __nv_bfloat16*)dy, side, dsw, m, out_c, side_dim, rpb);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

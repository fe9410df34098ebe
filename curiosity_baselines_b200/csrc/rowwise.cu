// Per-sample reductions over the feature dimension (one warp per row, shuffle reductions),
// loss weighting, and the fused clip+Adam optimiser step.
//   icm.py:131-145, disagreement.py:136-167, rnd.py:181-212, utils/tensor.py:39-46,
//   algos/pg/ppo.py:147-150.
#include "common.cuh"

namespace {

constexpr int kRowThreads = 256;                   // 8 rows per block

__global__ void rowwise_sqerr_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                     const float* __restrict__ elem_mask, float elem_scale,
                                     float scale, float* __restrict__ out, int64_t m, int f) {
    int64_t row = (int64_t)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
    if (row >= m) return;
    int lane = threadIdx.x & 31;
    const float* p = pred + row * f;
    const float* t = target + row * f;
    float acc = 0.f;
    for (int j = lane; j < f; j += 32) {
        float d = p[j] - t[j];
        float e = d * d;
        if (elem_mask) e *= elem_mask[row * f + j] * elem_scale;
        acc += e;
    }
    acc = cb::warp_sum(acc);
    if (lane == 0) out[row] = acc * scale;
}

__global__ void sqerr_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                  const float* __restrict__ coef, float scale,
                                  const float* __restrict__ elem_mask, float elem_scale,
                                  __nv_bfloat16* __restrict__ dpred, int64_t total, int f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t row = i / f;
    float g = 2.f * scale * coef[row] * (pred[i] - target[i]);
    if (elem_mask) g *= elem_mask[i] * elem_scale;
    dpred[i] = cb::f2bf(g);
}

__global__ void ensemble_variance_kernel(const float* const* __restrict__ preds, int e, float beta,
                                         float* __restrict__ out, int64_t m, int f) {
    int64_t row = (int64_t)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
    if (row >= m) return;
    int lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int j = lane; j < f; j += 32) {
        float mean = 0.f;
        for (int k = 0; k < e; ++k) mean += preds[k][row * f + j];
        mean /= (float)e;
        float v = 0.f;
        for (int k = 0; k < e; ++k) { float d = preds[k][row * f + j] - mean; v += d * d; }
        acc += v / (float)(e - 1);                          // unbiased (torch.var default)
    }
    acc = cb::warp_sum(acc);
    if (lane == 0) out[row] = beta * acc / (float)f;
}

// Ensemble predictions packed side by side, pred [m, e*f] (the grouped-GEMM layout):
//   out[k][row] = scale * sum_j mask_k[row,j]*elem_scale * (pred[row, k*f+j] - target[row,j])^2
__global__ void ensemble_sqerr_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                      const float* __restrict__ masks, float elem_scale, float scale,
                                      float* __restrict__ out, int64_t m, int f, int e) {
    int64_t row = (int64_t)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
    if (row >= m) return;
    int lane = threadIdx.x & 31;
    const float* t = target + row * f;
    for (int k = 0; k < e; ++k) {
        const float* p = pred + (row * e + k) * f;
        const float* mk = masks ? masks + ((int64_t)k * m + row) * f : nullptr;
        float acc = 0.f;
        for (int j = lane; j < f; j += 32) {
            float d = p[j] - t[j];
            float v = d * d;
            if (mk) v *= mk[j] * elem_scale;
            acc += v;
        }
        acc = cb::warp_sum(acc);
        if (lane == 0) out[(int64_t)k * m + row] = acc * scale;
    }
}

__global__ void ensemble_sqerr_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                           const float* __restrict__ coef, float scale,
                                           const float* __restrict__ masks, float elem_scale,
                                           __nv_bfloat16* __restrict__ dpred, int64_t m, int f, int e) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // over m * e * f
    if (i >= m * e * f) return;
    const int64_t row = i / ((int64_t)e * f);
    const int rem = (int)(i - row * e * f), k = rem / f, j = rem - k * f;
    float g = 2.f * scale * coef[row] * (pred[i] - target[row * f + j]);
    if (masks) g *= masks[((int64_t)k * m + row) * f + j] * elem_scale;
    dpred[i] = cb::f2bf(g);
}

__global__ void ensemble_variance_packed_kernel(const float* __restrict__ pred, int e, float beta,
                                                float* __restrict__ out, int64_t m, int f) {
    int64_t row = (int64_t)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
    if (row >= m) return;
    int lane = threadIdx.x & 31;
    const float* p = pred + row * e * f;
    float acc = 0.f;
    for (int j = lane; j < f; j += 32) {
        float mean = 0.f;
        for (int k = 0; k < e; ++k) mean += p[k * f + j];
        mean /= (float)e;
        float v = 0.f;
        for (int k = 0; k < e; ++k) { float d = p[k * f + j] - mean; v += d * d; }
        acc += v / (float)(e - 1);                          // unbiased (torch.var default)
    }
    acc = cb::warp_sum(acc);
    if (lane == 0) out[row] = beta * acc / (float)f;
}

// binary_cross_entropy(sigmoid(z), y) as torch computes it (fp32 sigmoid, log clamped at -100):
// row[m] = scale * sum_f bce ; dz = gscale * (p - y)           (ndigo.py:197-206, 265-268)
__global__ void bce_logits_kernel(const float* __restrict__ z, const float* __restrict__ y, float scale,
                                  float* __restrict__ row, float gscale, float* __restrict__ dz, int64_t m, int f) {
    int64_t r = (int64_t)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
    if (r >= m) return;
    int lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int j = lane; j < f; j += 32) {
        const float zz = z[r * f + j], yy = y[r * f + j];
        const float p = 1.f / (1.f + expf(-zz));
        const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
        acc -= yy * lp + (1.f - yy) * lq;
        if (dz) dz[r * f + j] = gscale * (p - yy);
    }
    acc = cb::warp_sum(acc);
    if (lane == 0 && row) row[r] = acc * scale;
}

// All of NDIGO's horizons in one pass (ndigo.py:154-163, 194-206, 234-268).  logits [n, G*fp] holds predictor g+1's
// outputs in columns [g*fp, g*fp + d) (the grouped-GEMM layout, fp = d rounded up to the tile); row (t,b) of predictor
// k = g+1 is scored against obs[t+k, b]; rows with t + k >= T belong to no pair and contribute nothing.
//   row_loss[g][t*B+b] = row_scale_g * sum_f bce      dz[n, G*fp] (bf16) = grad_scale_g * (sigmoid - target), 0 elsewhere
// with row_scale_g = 1/d (bonus) or 1/((T-k)*B*d) (loss, so that the sum over rows is the mean over the pairs),
// grad_scale_g = upstream / ((T-k)*B*d).
struct HorizonOffsets { int off[16]; };
__global__ void bce_horizons_kernel(const float* __restrict__ z, const float* __restrict__ obs, int T, int B, int d,
                                    int G, int fp, int mean_over_pairs, float* __restrict__ row_loss, float upstream,
                                    __nv_bfloat16* __restrict__ dz, HorizonOffsets ho) {
    const int64_t n = (int64_t)T * B;
    const int64_t r = (int64_t)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
    if (r >= n) return;
    const int lane = threadIdx.x & 31;
    const int t = (int)(r / B), b = (int)(r - (int64_t)t * B);
    for (int g = 0; g < G; ++g) {
        const int k = g + 1;
        const bool live = t + k < T;
        const float pairs = (float)(T - k) * (float)B * (float)d;
        const float rs = mean_over_pairs ? 1.f / pairs : 1.f / (float)d, gs = upstream / pairs;
        const float* zr = z + (r * G + g) * fp;
        const float* y = obs + ((int64_t)(t + ho.off[g]) * B + b) * d;
        __nv_bfloat16* dr = dz ? dz + (r * G + g) * fp : nullptr;
        float acc = 0.f;
        for (int j = lane; j < fp; j += 32) {
            float grad = 0.f;
            if (live && j < d) {
                // torch's formula (fp32 sigmoid, both logs clamped at -100) on the fast SFU intrinsics: ~2 ulp, far inside
                // the 1e-2 gate; this pass is otherwise bound by its 2 logs + exp + divide per element, not by HBM
                const float zz = zr[j], yy = y[j];
                const float p = __fdividef(1.f, 1.f + __expf(-zz));
                const float lp = fmaxf(__logf(p), -100.f), lq = fmaxf(__logf(1.f - p), -100.f);
                acc -= yy * lp + (1.f - yy) * lq;
                grad = gs * (p - yy);
            }
            if (dr) dr[j] = cb::f2bf(grad);
        }
        acc = cb::warp_sum(acc);
        if (lane == 0 && row_loss) row_loss[(int64_t)g * n + r] = live ? acc * rs : 0.f;
    }
}

// Sliding action windows of NDIGO (ndigo.py:154-163, 236-240: the reference builds them with a Python double loop):
// win[(t,b), j*A + a] = onehot[t+j, b, a] for j < W and t + j < T, zero otherwise and in the padding columns.
// Written twice: bf16 [n, 64] (the extra k-block of the grouped first layer) and fp32 [n, W*A] (side weight gradient).
__global__ void action_windows_kernel(const float* __restrict__ onehot, int T, int B, int A, int W,
                                      __nv_bfloat16* __restrict__ win_bf16, int ld, float* __restrict__ win_f32) {
    const int64_t n = (int64_t)T * B;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * 64) return;
    const int64_t r = idx >> 6;
    const int c = (int)(idx & 63);
    const int t = (int)(r / B), b = (int)(r - (int64_t)t * B);
    float v = 0.f;
    if (c < W * A) {
        const int j = c / A, a = c - j * A;
        if (t + j < T) v = __ldg(onehot + ((int64_t)(t + j) * B + b) * A + a);
        if (win_f32) win_f32[r * (W * A) + c] = v;
    }
    win_bf16[r * ld + c] = cb::f2bf(v);
}

// one thread per row; A (actions) is small
__global__ void cross_entropy_kernel(const float* __restrict__ logits, const float* __restrict__ onehot,
                                     const float* __restrict__ coef, float* __restrict__ loss,
                                     float* __restrict__ dlogits, int64_t m, int a) {
    int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= m) return;
    const float* z = logits + row * a;
    const float* oh = onehot + row * a;
    int label = 0; float best = oh[0], zmax = z[0];
    for (int j = 1; j < a; ++j) {
        if (oh[j] > best) { best = oh[j]; label = j; }      // first max, like torch.max
        zmax = fmaxf(zmax, z[j]);
    }
    float se = 0.f;
    for (int j = 0; j < a; ++j) se += expf(z[j] - zmax);
    float lse = logf(se) + zmax;
    if (loss) loss[row] = lse - z[label];
    if (dlogits) {
        float c = coef[row];
        for (int j = 0; j < a; ++j)
            dlogits[row * a + j] = c * (expf(z[j] - lse) - (j == label ? 1.f : 0.f));
    }
}

__global__ void weighted_mean_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                     float* __restrict__ out, int64_t n) {
    // single block of 1024 threads, fixed traversal order -> deterministic
    double sx = 0, sw = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        float wi = w ? w[i] : 1.f;
        sx += (double)(x[i] * wi);
        sw += (double)wi;
    }
    __shared__ double sm[2][32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    sx = cb::warp_sum(sx); sw = cb::warp_sum(sw);
    if (lane == 0) { sm[0][wid] = sx; sm[1][wid] = sw; }
    __syncthreads();
    if (wid == 0) {
        sx = lane < (blockDim.x >> 5) ? sm[0][lane] : 0;
        sw = lane < (blockDim.x >> 5) ? sm[1][lane] : 0;
        sx = cb::warp_sum(sx); sw = cb::warp_sum(sw);
        if (lane == 0) { out[0] = (float)(sx / sw); out[1] = (float)sw; }
    }
}

__global__ void dot_kernel(const float* __restrict__ x, const float* __restrict__ w,
                           float* __restrict__ out, int64_t n) {
    double s = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i] * (double)w[i];
    __shared__ double sm[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s = cb::warp_sum(s);
    if (lane == 0) sm[wid] = s;
    __syncthreads();
    if (wid == 0) {
        s = lane < (blockDim.x >> 5) ? sm[lane] : 0;
        s = cb::warp_sum(s);
        if (lane == 0) out[0] = (float)s;
    }
}

__global__ void loss_coef_kernel(const float* __restrict__ valid, const float* __restrict__ mask,
                                 float upstream, const float* __restrict__ denom,
                                 float* __restrict__ coef, int64_t n) {
    // single block: sum(valid) then scale -- n is T*B (<= a few hundred thousand)
    double s = 0;
    if (!denom)
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)valid[i];
    __shared__ double sm[32];
    __shared__ float inv;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s = cb::warp_sum(s);
    if (lane == 0) sm[wid] = s;
    __syncthreads();
    if (wid == 0) {
        s = lane < (blockDim.x >> 5) ? sm[lane] : 0;
        s = cb::warp_sum(s);
        if (lane == 0) inv = (float)((double)upstream / (denom ? (double)*denom : s));
    }
    __syncthreads();
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
        coef[i] = valid[i] * (mask ? mask[i] : 1.f) * inv;
}

__global__ void sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
    double s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double v = (double)g[i];
        s += v * v;
    }
    __shared__ double sm[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s = cb::warp_sum(s);
    if (lane == 0) sm[wid] = s;
    __syncthreads();
    if (wid == 0) {
        s = lane < (blockDim.x >> 5) ? sm[lane] : 0;
        s = cb::warp_sum(s);
        if (lane == 0) atomicAdd(out, s);
    }
}

__global__ void adam_clip_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                 const float* __restrict__ g, int64_t n,
                                 const double* __restrict__ sumsq, float max_norm, float lr,
                                 float beta1, float beta2, float eps, float step_size, float bc2_sqrt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // torch.nn.utils.clip_grad_norm_: coef = max_norm/(norm+1e-6), clamped to 1
    float norm = (float)sqrt(*sumsq);
    float coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
    float gi = g[i] * coef;
    float mi = beta1 * m[i] + (1.f - beta1) * gi;
    float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    // torch Adam: denom = sqrt(v)/sqrt(bc2) + eps ; p -= lr/bc1 * m/denom
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
}

// The same update with the learning rate and the step counter read from DEVICE memory (the counter is bumped by a
// one-thread launch first): nothing step-dependent is baked into the launch parameters, so a captured CUDA graph of a
// whole training step replays correctly while the schedule moves on.
__global__ void bump_step_kernel(int32_t* step) { *step += 1; }
__global__ void adam_clip_dev_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                     const float* __restrict__ g, int64_t n, const double* __restrict__ sumsq,
                                     float max_norm, float beta1, float beta2, float eps, const float* __restrict__ lr,
                                     const int32_t* __restrict__ step) {
    __shared__ float s_step_size, s_bc2;
    if (threadIdx.x == 0) {       // bias corrections in double, like torch's Python scalars
        const double st = (double)(*step);
        const double bc1 = 1.0 - pow((double)beta1, st), bc2 = 1.0 - pow((double)beta2, st);
        s_step_size = (float)((double)(*lr) / bc1);
        s_bc2 = (float)sqrt(bc2);
    }
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float norm = (float)sqrt(*sumsq);
    float coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
    float gi = g[i] * coef;
    float mi = beta1 * m[i] + (1.f - beta1) * gi;
    float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / s_bc2 + eps;
    p[i] -= s_step_size * (mi / denom);
}

__global__ void prepare_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                           __nv_bfloat16* __restrict__ wt, int o, int c, int kh, int kw) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)o * c * kh * kw;
    if (idx >= total) return;
    // idx enumerates the torch layout [o][c][kh][kw]
    int x = idx % kw; int64_t r = idx / kw;
    int y = r % kh; r /= kh;
    int ci = r % c; int oi = (int)(r / c);
    __nv_bfloat16 val = cb::f2bf(w[idx]);
    if (wf) wf[(((int64_t)oi * kh + y) * kw + x) * c + ci] = val;
    if (wt) wt[(((int64_t)ci * kh + y) * kw + x) * o + oi] = val;
}

__global__ void unprepare_conv_grad_kernel(const float* __restrict__ g, float* __restrict__ dst, int o,
                                           int c, int kh, int kw, float beta) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)o * c * kh * kw;
    if (idx >= total) return;
    int x = idx % kw; int64_t r = idx / kw;
    int y = r % kh; r /= kh;
    int ci = r % c; int oi = (int)(r / c);
    float s = g[(((int64_t)oi * kh + y) * kw + x) * c + ci];
    dst[idx] = beta == 0.f ? s : beta * dst[idx] + s;
}

}  // namespace

extern "C" int cb_rowwise_sqerr(const float* pred, const float* target, float scale, float* out,
                                int64_t m, int32_t f, void* stream) {
    CB_REQUIRE(pred && target && out && m >= 1 && f >= 1, CB_E_ARG, "cb_rowwise_sqerr: bad argument");
    rowwise_sqerr_kernel<<<(unsigned)cb::ceil_div<int64_t>(m, kRowThreads / 32), kRowThreads, 0,
                           cb::as_stream(stream)>>>(pred, target, nullptr, 1.f, scale, out, m, f);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_rowwise_sqerr_masked(const float* pred, const float* target, const float* elem_mask,
                                       float elem_scale, float scale, float* out, int64_t m, int32_t f,
                                       void* stream) {
    CB_REQUIRE(pred && target && out && m >= 1 && f >= 1, CB_E_ARG, "cb_rowwise_sqerr_masked: bad argument");
    rowwise_sqerr_kernel<<<(unsigned)cb::ceil_div<int64_t>(m, kRowThreads / 32), kRowThreads, 0,
                           cb::as_stream(stream)>>>(pred, target, elem_mask, elem_scale, scale, out, m, f);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_sqerr_grad(const float* pred, const float* target, const float* coef, float scale,
                             const float* elem_mask, float elem_scale, void* dpred_bf16, int64_t m,
                             int32_t f, void* stream) {
    CB_REQUIRE(pred && target && coef && dpred_bf16 && m >= 1 && f >= 1, CB_E_ARG, "cb_sqerr_grad: bad argument");
    int64_t total = m * f;
    sqerr_grad_kernel<<<(unsigned)cb::ceil_div<int64_t>(total, 256), 256, 0, cb::as_stream(stream)>>>(
        pred, target, coef, scale, elem_mask, elem_scale, (__nv_bfloat16*)dpred_bf16, total, f);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_ensemble_variance(const float* const* preds, int32_t e, float beta, float* out,
                                    int64_t m, int32_t f, void* stream) {
    CB_REQUIRE(preds && out && e >= 2 && m >= 1 && f >= 1, CB_E_ARG, "cb_ensemble_variance: bad argument");
    ensemble_variance_kernel<<<(unsigned)cb::ceil_div<int64_t>(m, kRowThreads / 32), kRowThreads, 0,
                               cb::as_stream(stream)>>>(preds, e, beta, out, m, f);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_ensemble_sqerr(const float* pred, const float* target, const float* masks, float elem_scale,
                                 float scale, float* out, int64_t m, int32_t f, int32_t e, void* stream) {
    CB_REQUIRE(pred && target && out && m >= 1 && f >= 1 && e >= 1, CB_E_ARG, "cb_ensemble_sqerr: bad argument");
    ensemble_sqerr_kernel<<<(unsigned)cb::ceil_div<int64_t>(m, kRowThreads / 32), kRowThreads, 0,
                            cb::as_stream(stream)>>>(pred, target, masks, elem_scale, scale, out, m, f, e);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_ensemble_sqerr_grad(const float* pred, const float* target, const float* coef, float scale,
                                      const float* masks, float elem_scale, void* dpred_bf16, int64_t m, int32_t f,
                                      int32_t e, void* stream) {
    CB_REQUIRE(pred && target && coef && dpred_bf16 && m >= 1 && f >= 1 && e >= 1, CB_E_ARG,
               "cb_ensemble_sqerr_grad: bad argument");
    ensemble_sqerr_grad_kernel<<<(unsigned)cb::ceil_div<int64_t>(m * e * f, 256), 256, 0, cb::as_stream(stream)>>>(
        pred, target, coef, scale, masks, elem_scale, (__nv_bfloat16*)dpred_bf16, m, f, e);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_ensemble_variance_packed(const float* pred, int32_t e, float beta, float* out, int64_t m,
                                           int32_t f, void* stream) {
    CB_REQUIRE(pred && out && e >= 2 && m >= 1 && f >= 1, CB_E_ARG, "cb_ensemble_variance_packed: bad argument");
    ensemble_variance_packed_kernel<<<(unsigned)cb::ceil_div<int64_t>(m, kRowThreads / 32), kRowThreads, 0,
                                      cb::as_stream(stream)>>>(pred, e, beta, out, m, f);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_cross_entropy(const float* logits, const float* onehot, const float* coef, float* loss,
                                float* dlogits, int64_t m, int32_t a, void* stream) {
    CB_REQUIRE(logits && onehot && m >= 1 && a >= 1, CB_E_ARG, "cb_cross_entropy: bad argument");
    CB_REQUIRE(!dlogits || coef, CB_E_ARG, "cb_cross_entropy: dlogits needs coef");
    cross_entropy_kernel<<<(unsigned)cb::ceil_div<int64_t>(m, 128), 128, 0, cb::as_stream(stream)>>>(
        logits, onehot, coef, loss, dlogits, m, a);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_bce_logits(const float* logits, const float* target, float scale, float* row_loss, float gscale,
                             float* dlogits, int64_t m, int32_t f, void* stream) {
    CB_REQUIRE(logits && target && (row_loss || dlogits) && m >= 1 && f >= 1, CB_E_ARG, "cb_bce_logits: bad argument");
    bce_logits_kernel<<<(unsigned)cb::ceil_div<int64_t>(m, kRowThreads / 32), kRowThreads, 0, cb::as_stream(stream)>>>(
        logits, target, scale, row_loss, gscale, dlogits, m, f);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_bce_horizons(const float* logits, const float* obs, int32_t T, int32_t B, int32_t d, int32_t groups,
                               int32_t fp, int32_t mean_over_pairs, const int32_t* target_offset_host, float* row_loss,
                               float upstream, void* dlogits_bf16, void* stream) {
    CB_REQUIRE(logits && obs && (row_loss || dlogits_bf16) && T >= 1 && B >= 1 && d >= 1 && groups >= 1 && groups <= 16 &&
               fp >= d, CB_E_ARG, "cb_bce_horizons: bad argument");
    HorizonOffsets ho;
    for (int g = 0; g < 16; ++g) ho.off[g] = g + 1;
    if (target_offset_host)
        for (int g = 0; g < groups; ++g) {
            CB_REQUIRE(target_offset_host[g] >= 0 && target_offset_host[g] <= g + 1, CB_E_ARG,
                       "cb_bce_horizons: target offset of predictor k must be in [0, k]");
            ho.off[g] = target_offset_host[g];
        }
    const int64_t n = (int64_t)T * B;
    bce_horizons_kernel<<<(unsigned)cb::ceil_div<int64_t>(n, kRowThreads / 32), kRowThreads, 0, cb::as_stream(stream)>>>(
        logits, obs, T, B, d, groups, fp, mean_over_pairs, row_loss, upstream, (__nv_bfloat16*)dlogits_bf16, ho);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_action_windows(const float* onehot, int32_t T, int32_t B, int32_t A, int32_t W, void* win_bf16,
                                 int32_t ld_bf16, float* win_f32, void* stream) {
    CB_REQUIRE(onehot && win_bf16 && T >= 1 && B >= 1 && A >= 1 && W >= 1 && W * A <= 64 && ld_bf16 >= 64, CB_E_ARG,
               "cb_action_windows: bad argument (W*A must be <= 64, ld_bf16 >= 64)");
    const int64_t total = (int64_t)T * B * 64;
    action_windows_kernel<<<(unsigned)cb::ceil_div<int64_t>(total, 256), 256, 0, cb::as_stream(stream)>>>(
        onehot, T, B, A, W, (__nv_bfloat16*)win_bf16, ld_bf16, win_f32);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_weighted_mean(const float* x, const float* w, float* out, int64_t n, void* stream) {
    CB_REQUIRE(x && out && n >= 1, CB_E_ARG, "cb_weighted_mean: bad argument");
    weighted_mean_kernel<<<1, 1024, 0, cb::as_stream(stream)>>>(x, w, out, n);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_dot(const float* x, const float* w, float* out, int64_t n, void* stream) {
    CB_REQUIRE(x && w && out && n >= 1, CB_E_ARG, "cb_dot: bad argument");
    dot_kernel<<<1, 1024, 0, cb::as_stream(stream)>>>(x, w, out, n);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_loss_coef(const float* valid, const float* mask, float upstream, const float* denom,
                            float* coef, int64_t n, void* stream) {
    CB_REQUIRE(valid && coef && n >= 1, CB_E_ARG, "cb_loss_coef: bad argument");
    loss_coef_kernel<<<1, 1024, 0, cb::as_stream(stream)>>>(valid, mask, upstream, denom, coef, n);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_sumsq(const float* g, int64_t n, double* out, void* stream) {
    CB_REQUIRE(g && out && n >= 1, CB_E_ARG, "cb_sumsq: bad argument");
    int blocks = (int)std::min<int64_t>(148 * 4, cb::ceil_div<int64_t>(n, 256));
    sumsq_kernel<<<blocks, 256, 0, cb::as_stream(stream)>>>(g, n, out);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_adam_clip_step(float* p, float* m, float* v, const float* g, int64_t n,
                                 const double* sumsq, float max_norm, float lr, float beta1, float beta2,
                                 float eps, int32_t step, void* stream) {
    CB_REQUIRE(p && m && v && g && sumsq && n >= 1 && step >= 1, CB_E_ARG, "cb_adam_clip_step: bad argument");
    // bias corrections in double like torch's Python scalars (1 - powf(0.999f, 1) loses ~1e-5 relative in fp32)
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1), bc2s = (float)sqrt(bc2);
    adam_clip_kernel<<<(unsigned)cb::ceil_div<int64_t>(n, 256), 256, 0, cb::as_stream(stream)>>>(
        p, m, v, g, n, sumsq, max_norm, lr, beta1, beta2, eps, step_size, bc2s);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_adam_clip_step_dev(float* p, float* m, float* v, const float* g, int64_t n, const double* sumsq,
                                     float max_norm, float beta1, float beta2, float eps, const float* lr_dev,
                                     int32_t* step_dev, void* stream) {
    CB_REQUIRE(p && m && v && g && sumsq && lr_dev && step_dev && n >= 1, CB_E_ARG, "cb_adam_clip_step_dev: bad argument");
    cudaStream_t s = cb::as_stream(stream);
    bump_step_kernel<<<1, 1, 0, s>>>(step_dev);
    CB_LAUNCH_CHECK();
    adam_clip_dev_kernel<<<(unsigned)cb::ceil_div<int64_t>(n, 256), 256, 0, s>>>(p, m, v, g, n, sumsq, max_norm, beta1,
                                                                                  beta2, eps, lr_dev, step_dev);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_prepare_conv_weight(const float* w, void* w_fwd, void* w_t, int32_t o, int32_t i,
                                      int32_t kh, int32_t kw, void* stream) {
    CB_REQUIRE(w && (w_fwd || w_t), CB_E_ARG, "cb_prepare_conv_weight: null pointer");
    int64_t total = (int64_t)o * i * kh * kw;
    prepare_conv_weight_kernel<<<(unsigned)cb::ceil_div<int64_t>(total, 256), 256, 0, cb::as_stream(stream)>>>(
        w, (__nv_bfloat16*)w_fwd, (__nv_bfloat16*)w_t, o, i, kh, kw);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_unprepare_conv_grad(const float* g, float* dst, int32_t o, int32_t i, int32_t kh,
                                      int32_t kw, float beta, void* stream) {
    CB_REQUIRE(g && dst, CB_E_ARG, "cb_unprepare_conv_grad: null pointer");
    int64_t total = (int64_t)o * i * kh * kw;
    unprepare_conv_grad_kernel<<<(unsigned)cb::ceil_div<int64_t>(total, 256), 256, 0, cb::as_stream(stream)>>>(
        g, dst, o, i, kh, kw, beta);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

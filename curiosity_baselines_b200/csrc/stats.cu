// Running statistics of the RND path, kept entirely on the device.
//   rlpyt/models/curiosity/rnd.py:150-203, rlpyt/utils/averages.py:41-89.
// The observation moments are accumulated as exact integers (uint8 pixels -> uint64 sums),
// so the batch mean/variance equal numpy's float64 result to rounding.  HBM-bound.
#include "common.cuh"

namespace {

// ---- n_valid[b] = sum_t |done-1| ; total over b ------------------------------------------
__global__ void valid_lengths_kernel(const float* __restrict__ done, int32_t* __restrict__ n_valid,
                                     unsigned long long* __restrict__ n_total, int T, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    int n = 0;
    if (b < B) {
        float s = 0.f;
        for (int t = 0; t < T; ++t) s += fabsf(__ldg(done + (size_t)t * B + b) - 1.f);
        n = (int)s;                       // int(done[i].item()) in the reference
        n = min(max(n, 0), T);
        n_valid[b] = n;
    }
    // block reduce
    __shared__ int sm[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int v = n;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < (blockDim.x >> 5) ? sm[lane] : 0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(n_total, (unsigned long long)v);
    }
}

// ---- per-pixel integer moments -----------------------------------------------------------
// grid.x = frame groups, block = ceil(pixels/16) threads (one 16-byte column per thread).
constexpr int kFramesPerBlock = 128;     // sumsq <= 65025*128 fits uint32

__global__ void obs_moments_kernel(const uint8_t* __restrict__ frames, int64_t frame_stride,
                                   int pixels, const int32_t* __restrict__ n_valid, int T, int B,
                                   unsigned long long* __restrict__ sum,
                                   unsigned long long* __restrict__ sumsq) {
    int col = blockIdx.y * blockDim.x + threadIdx.x;       // 16-pixel column
    if (col * 16 >= pixels) return;
    int64_t f0 = (int64_t)blockIdx.x * kFramesPerBlock;
    int64_t f1 = min(f0 + (int64_t)kFramesPerBlock, (int64_t)T * B);
    uint32_t s[16], q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { s[i] = 0; q[i] = 0; }
    // four frames per trip: their 16-byte loads are all in flight before the first is consumed
    for (int64_t f = f0; f < f1; f += 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t ff = f + u;
            bool live = ff < f1;
            if (live) { const int t = (int)(ff / B), b = (int)(ff % B); live = t < __ldg(n_valid + b); }   // block-uniform
            v[u] = live ? __ldg(reinterpret_cast<const uint4*>(frames + ff * frame_stride) + col) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t p = (w[j] >> (8 * k)) & 0xffu;
                    s[j * 4 + k] += p;
                    q[j * 4 + k] += p * p;
                }
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        int px = col * 16 + i;
        if (s[i]) atomicAdd(sum + px, (unsigned long long)s[i]);
        if (q[i]) atomicAdd(sumsq + px, (unsigned long long)q[i]);
    }
}

// differ[0] += number of 16-byte words in which frame f of a differs from frame f of b (both uint8, `pixels` bytes
// per frame, frame strides in bytes).  Guard for reusing features computed on "the same" batch (rnd.py call sites:
// compute_bonus and compute_loss both receive samples.env.next_observation, algos/pg/base.py:72, ppo.py:107-110).
__global__ void frames_differ_kernel(const uint8_t* __restrict__ a, int64_t a_stride, const uint8_t* __restrict__ b,
                                     int64_t b_stride, int pixels, int64_t frames, unsigned int* __restrict__ differ) {
    const int words = pixels / 16;
    const int64_t total = frames * words;
    unsigned int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = i / words;
        const int w = (int)(i - f * words);
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(a + f * a_stride) + w);
        const uint4 y = __ldg(reinterpret_cast<const uint4*>(b + f * b_stride) + w);
        bad += (x.x != y.x) | (x.y != y.y) | (x.z != y.z) | (x.w != y.w);
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(differ, bad);
}

// ---- Chan merge (averages.py:54-68) --------------------------------------------------------
__device__ __forceinline__ void chan_merge(double& mean, double& var, double count, double b_mean,
                                           double b_var, double b_count) {
    double delta = b_mean - mean;
    double tot = count + b_count;
    double new_mean = mean + delta * b_count / tot;
    double m2 = var * count + b_var * b_count + delta * delta * count * b_count / (count + b_count);
    mean = new_mean;
    var = m2 / (count + b_count);
}

__global__ void rms_merge_from_sums_kernel(const unsigned long long* __restrict__ sum,
                                           const unsigned long long* __restrict__ sumsq,
                                           const int64_t* __restrict__ n_total,
                                           double* __restrict__ mean, double* __restrict__ var,
                                           const double* __restrict__ count,
                                           float* __restrict__ mean_f32, float* __restrict__ rstd_f32,
                                           int pixels) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    double n = (double)(*n_total);
    double m = mean[i], v = var[i];
    if (n > 0) {
        double bm = (double)sum[i] / n;
        // population variance from exact integer sums: (n*sumsq - sum^2)/n^2
        double bv = ((double)sumsq[i] - (double)sum[i] * bm) / n;
        if (bv < 0) bv = 0;
        chan_merge(m, v, *count, bm, bv, n);
        mean[i] = m; var[i] = v;
    }
    if (mean_f32) mean_f32[i] = (float)m;
    if (rstd_f32) rstd_f32[i] = 1.0f / sqrtf((float)v);
}

__global__ void bump_count_kernel(double* count, const int64_t* n_total) {
    *count += (double)(*n_total);
}

// ---- reward forward filter -----------------------------------------------------------------
__global__ void reward_filter_kernel(const float* __restrict__ rews, const float* __restrict__ done,
                                     float gamma, float* __restrict__ rewems,
                                     const int32_t* __restrict__ initialized,
                                     float* __restrict__ total, int T, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    bool init = *initialized != 0;
    float e = init ? rewems[b] : 0.f;
    for (int t = 0; t < T; ++t) {
        size_t o = (size_t)t * B + b;
        float r = __ldg(rews + o);
        if (!init) { e = r; init = true; }                    // first call seeds unmasked
        else if (fabsf(__ldg(done + o) - 1.f) == 1.0f) e = __fadd_rn(__fmul_rn(e, gamma), r);
        total[o] = e;
    }
    rewems[b] = e;
}
__global__ void set_flag_kernel(int32_t* f) { *f = 1; }

// ---- scalar moments + merge ------------------------------------------------------------------
// Two deterministic stages: per-block f64 partials, then one warp folds them into
// sums = {sum x, sum x^2, n}.  Kept separate from the merge so that a data-parallel host can
// all-reduce `sums` across ranks in between (the running state then equals the single-process one).
__global__ void moments_partial_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ ws) {
    double s = 0, q = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double d = (double)__ldg(x + i);
        s += d; q += d * d;
    }
    __shared__ double sm[2][32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s = cb::warp_sum(s); q = cb::warp_sum(q);
    if (lane == 0) { sm[0][wid] = s; sm[1][wid] = q; }
    __syncthreads();
    if (wid == 0) {
        s = lane < (blockDim.x >> 5) ? sm[0][lane] : 0;
        q = lane < (blockDim.x >> 5) ? sm[1][lane] : 0;
        s = cb::warp_sum(s); q = cb::warp_sum(q);
        if (lane == 0) { ws[blockIdx.x] = s; ws[1024 + blockIdx.x] = q; }
    }
}
__global__ void moments_finish_kernel(int64_t n, int nblocks, const double* __restrict__ ws,
                                      double* __restrict__ sums) {
    double s = 0, q = 0;
    for (int i = threadIdx.x; i < nblocks; i += 32) { s += ws[i]; q += ws[1024 + i]; }
    s = cb::warp_sum(s); q = cb::warp_sum(q);
    if (threadIdx.x == 0) { sums[0] = s; sums[1] = q; sums[2] = (double)n; }
}
__global__ void rms_merge_moments_kernel(const double* __restrict__ sums, const double* count_num,
                                         double count_den, double* mean, double* var, double* count) {
    double n = sums[2];
    if (n <= 0) return;
    double bm = sums[0] / n;
    double bv = sums[1] / n - bm * bm;
    if (bv < 0) bv = 0;
    double bc = count_num ? (*count_num) / count_den : n;
    double m = *mean, v = *var, c = *count;
    chan_merge(m, v, c, bm, bv, bc);
    *mean = m; *var = v; *count = c + bc;
}

__global__ void valid_count_kernel(const float* __restrict__ done, double* out, int64_t n) {
    double s = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)fabsf(__ldg(done + i) - 1.f);
    __shared__ double sm[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s = cb::warp_sum(s);
    if (lane == 0) sm[wid] = s;
    __syncthreads();
    if (wid == 0) {
        s = lane < (blockDim.x >> 5) ? sm[lane] : 0;
        s = cb::warp_sum(s);
        if (lane == 0) out[0] = s;
    }
}

// (x - mean_valid) / max(std_valid, floor) over ALL elements; mean/std over mask>0, unbiased std
__global__ void masked_normalize_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                        float floor_std, float* __restrict__ out, int64_t n) {
    __shared__ double sm[2][32];
    __shared__ double s_mean, s_cnt, s_std;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double s = 0, c = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
        if (!mask || mask[i] > 0.f) { s += (double)x[i]; c += 1.0; }
    s = cb::warp_sum(s); c = cb::warp_sum(c);
    if (lane == 0) { sm[0][wid] = s; sm[1][wid] = c; }
    __syncthreads();
    if (wid == 0) {
        s = lane < (blockDim.x >> 5) ? sm[0][lane] : 0;
        c = lane < (blockDim.x >> 5) ? sm[1][lane] : 0;
        s = cb::warp_sum(s); c = cb::warp_sum(c);
        if (lane == 0) { s_mean = s / c; s_cnt = c; }
    }
    __syncthreads();
    double mean = s_mean, q = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
        if (!mask || mask[i] > 0.f) { double d = (double)x[i] - mean; q += d * d; }
    q = cb::warp_sum(q);
    if (lane == 0) sm[0][wid] = q;
    __syncthreads();
    if (wid == 0) {
        q = lane < (blockDim.x >> 5) ? sm[0][lane] : 0;
        q = cb::warp_sum(q);
        if (lane == 0) s_std = sqrt(q / (s_cnt - 1.0));
    }
    __syncthreads();
    float mf = (float)mean, sd = fmaxf((float)s_std, floor_std);
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) out[i] = __fdiv_rn(__fsub_rn(x[i], mf), sd);
}

// ---- generic RunningMeanStd.update(x) (averages.py:48-52): x f64 [n, d], one block per column -----------------
// batch mean and population variance are computed in two passes like numpy, then Chan-merged with the OLD count;
// the count itself is bumped by a second one-thread launch once every column has been merged.
__device__ __forceinline__ double block_sum(double v, double* sm) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = cb::warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
    if (wid == 0) { v = cb::warp_sum(v); if (lane == 0) sm[0] = v; }
    __syncthreads();
    return sm[0];
}
__global__ void rms_update_columns_kernel(const double* __restrict__ x, int64_t n, int d, double* __restrict__ mean,
                                          double* __restrict__ var, const double* __restrict__ count) {
    __shared__ double sm[32];
    const int c = blockIdx.x;
    double s = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i * d + c];
    const double bm = block_sum(s, sm) / (double)n;
    double q = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const double t = x[i * d + c] - bm; q += t * t; }
    const double bv = block_sum(q, sm) / (double)n;
    if (threadIdx.x == 0) {
        double m = mean[c], v = var[c];
        chan_merge(m, v, *count, bm, bv, (double)n);
        mean[c] = m; var[c] = v;
    }
}
__global__ void add_count_kernel(double* count, double n) { *count += n; }

// ---- advantage normalisation in two stages (algos/pg/base.py:95-103) ------------------------------------------
// stage 1: per-block f64 partials of {sum x, sum x^2, count} over mask > 0, folded into sums[3]; a data-parallel
// host all-reduces `sums`; stage 2 applies (x - mean) / max(std_unbiased, floor) to every element.
__global__ void masked_moments_partial_kernel(const float* __restrict__ x, const float* __restrict__ mask, int64_t n,
                                              double* __restrict__ ws) {
    double s = 0, q = 0, c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (!mask || __ldg(mask + i) > 0.f) { const double d = (double)__ldg(x + i); s += d; q += d * d; c += 1.0; }
    __shared__ double sm[32];
    s = block_sum(s, sm); q = block_sum(q, sm); c = block_sum(c, sm);
    if (threadIdx.x == 0) { ws[blockIdx.x] = s; ws[512 + blockIdx.x] = q; ws[1024 + blockIdx.x] = c; }
}
__global__ void masked_moments_finish_kernel(int nblocks, const double* __restrict__ ws, double* __restrict__ sums) {
    double s = 0, q = 0, c = 0;
    for (int i = threadIdx.x; i < nblocks; i += 32) { s += ws[i]; q += ws[512 + i]; c += ws[1024 + i]; }
    s = cb::warp_sum(s); q = cb::warp_sum(q); c = cb::warp_sum(c);
    if (threadIdx.x == 0) { sums[0] = s; sums[1] = q; sums[2] = c; }
}
__global__ void normalize_apply_kernel(const float* __restrict__ x, const double* __restrict__ sums, float floor_std,
                                       float* __restrict__ out, int64_t n) {
    const double c = sums[2], mean = sums[0] / c;
    double m2 = sums[1] - sums[0] * mean;                  // sum (x - mean)^2
    if (m2 < 0) m2 = 0;
    const float mf = (float)mean, sd = fmaxf((float)sqrt(m2 / (c - 1.0)), floor_std);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __fdiv_rn(__fsub_rn(x[i], mf), sd);
}

__global__ void rnd_scale_bonus_kernel(const float* __restrict__ err, const float* __restrict__ done,
                                       const double* __restrict__ rew_var, float beta,
                                       float* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float sd = sqrtf((float)(*rew_var));
    out[i] = beta * (__fdiv_rn(err[i], sd) * fabsf(done[i] - 1.f));
}

}  // namespace

extern "C" int cb_valid_lengths(const float* done, int32_t* n_valid, int64_t* n_total, int T, int B,
                                void* stream) {
    CB_REQUIRE(done && n_valid && n_total, CB_E_ARG, "cb_valid_lengths: null pointer");
    CB_REQUIRE(T >= 1 && B >= 1, CB_E_ARG, "cb_valid_lengths: T=%d B=%d", T, B);
    cudaStream_t s = cb::as_stream(stream);
    CB_CUDA(cudaMemsetAsync(n_total, 0, sizeof(int64_t), s));
    valid_lengths_kernel<<<cb::ceil_div(B, 128), 128, 0, s>>>(
        done, n_valid, reinterpret_cast<unsigned long long*>(n_total), T, B);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_obs_moments_u8(const uint8_t* frames, int64_t frame_stride, int pixels,
                                 const int32_t* n_valid, int T, int B, unsigned long long* sum,
                                 unsigned long long* sumsq, void* stream) {
    CB_REQUIRE(frames && n_valid && sum && sumsq, CB_E_ARG, "cb_obs_moments_u8: null pointer");
    CB_REQUIRE(pixels > 0 && pixels % 16 == 0 && frame_stride % 16 == 0, CB_E_ALIGN,
               "cb_obs_moments_u8: pixels=%d frame_stride=%lld must be multiples of 16", pixels,
               (long long)frame_stride);
    CB_REQUIRE(((uintptr_t)frames & 15) == 0, CB_E_ALIGN, "cb_obs_moments_u8: frames not 16B aligned");
    int cols = pixels / 16;
    int threads = 128;
    dim3 grid((unsigned)cb::ceil_div((int64_t)T * B, (int64_t)kFramesPerBlock),
              (unsigned)cb::ceil_div(cols, threads));
    obs_moments_kernel<<<grid, threads, 0, cb::as_stream(stream)>>>(frames, frame_stride, pixels,
                                                                    n_valid, T, B, sum, sumsq);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_frames_differ(const void* a, int64_t a_stride, const void* b, int64_t b_stride, int32_t pixels,
                                int64_t frames, uint32_t* differ, void* stream) {
    CB_REQUIRE(a && b && differ && pixels >= 16 && pixels % 16 == 0 && frames >= 1, CB_E_ARG, "cb_frames_differ: bad argument");
    CB_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)a_stride | (uintptr_t)b_stride) & 15) == 0, CB_E_ALIGN,
               "cb_frames_differ: frames must be 16-byte aligned");
    frames_differ_kernel<<<148 * 8, 256, 0, cb::as_stream(stream)>>>((const uint8_t*)a, a_stride, (const uint8_t*)b, b_stride,
                                                                   pixels, frames, differ);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_rms_merge_from_sums(const unsigned long long* sum, const unsigned long long* sumsq,
                                      const int64_t* n_total, double* mean, double* var,
                                      double* count, float* mean_f32, float* rstd_f32, int pixels,
                                      void* stream) {
    CB_REQUIRE(sum && sumsq && n_total && mean && var && count, CB_E_ARG,
               "cb_rms_merge_from_sums: null pointer");
    cudaStream_t s = cb::as_stream(stream);
    rms_merge_from_sums_kernel<<<cb::ceil_div(pixels, 256), 256, 0, s>>>(
        sum, sumsq, n_total, mean, var, count, mean_f32, rstd_f32, pixels);
    CB_LAUNCH_CHECK();
    bump_count_kernel<<<1, 1, 0, s>>>(count, n_total);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_reward_filter(const float* rews, const float* done, float gamma, float* rewems,
                                int32_t* initialized, float* total, int T, int B, void* stream) {
    CB_REQUIRE(rews && done && rewems && initialized && total, CB_E_ARG,
               "cb_reward_filter: null pointer");
    cudaStream_t s = cb::as_stream(stream);
    reward_filter_kernel<<<cb::ceil_div(B, 128), 128, 0, s>>>(rews, done, gamma, rewems,
                                                              initialized, total, T, B);
    CB_LAUNCH_CHECK();
    set_flag_kernel<<<1, 1, 0, s>>>(initialized);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_moments_partial(const float* x, int64_t n, double* sums, double* ws, void* stream) {
    CB_REQUIRE(x && sums && ws, CB_E_ARG, "cb_moments_partial: null pointer");
    CB_REQUIRE(n >= 1, CB_E_ARG, "cb_moments_partial: n=%lld", (long long)n);
    cudaStream_t s = cb::as_stream(stream);
    int nblocks = (int)std::min<int64_t>(1024, cb::ceil_div<int64_t>(n, 256));
    moments_partial_kernel<<<nblocks, 256, 0, s>>>(x, n, ws);
    CB_LAUNCH_CHECK();
    moments_finish_kernel<<<1, 32, 0, s>>>(n, nblocks, ws, sums);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_rms_merge_moments(const double* sums, const double* count_num, double count_den,
                                    double* mean, double* var, double* count, void* stream) {
    CB_REQUIRE(sums && mean && var && count, CB_E_ARG, "cb_rms_merge_moments: null pointer");
    CB_REQUIRE(!count_num || count_den > 0, CB_E_ARG, "cb_rms_merge_moments: count_den must be > 0");
    rms_merge_moments_kernel<<<1, 1, 0, cb::as_stream(stream)>>>(sums, count_num, count_den, mean, var, count);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_valid_count(const float* done, double* out, int64_t n, void* stream) {
    CB_REQUIRE(done && out && n >= 1, CB_E_ARG, "cb_valid_count: bad argument");
    valid_count_kernel<<<1, 1024, 0, cb::as_stream(stream)>>>(done, out, n);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_masked_normalize(const float* x, const float* mask, float floor_std, float* out,
                                   int64_t n, void* stream) {
    CB_REQUIRE(x && out && n >= 2, CB_E_ARG, "cb_masked_normalize: bad argument");
    masked_normalize_kernel<<<1, 1024, 0, cb::as_stream(stream)>>>(x, mask, floor_std, out, n);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_rms_update_columns(const double* x, int64_t n, int32_t d, double* mean, double* var, double* count,
                                     void* stream) {
    CB_REQUIRE(x && mean && var && count && n >= 1 && d >= 1, CB_E_ARG, "cb_rms_update_columns: bad argument");
    cudaStream_t s = cb::as_stream(stream);
    rms_update_columns_kernel<<<d, 256, 0, s>>>(x, n, d, mean, var, count);
    CB_LAUNCH_CHECK();
    add_count_kernel<<<1, 1, 0, s>>>(count, (double)n);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_masked_moments(const float* x, const float* mask, int64_t n, double* sums, double* ws, void* stream) {
    CB_REQUIRE(x && sums && ws && n >= 1, CB_E_ARG, "cb_masked_moments: bad argument");
    cudaStream_t s = cb::as_stream(stream);
    int nblocks = (int)std::min<int64_t>(512, cb::ceil_div<int64_t>(n, 1024));
    masked_moments_partial_kernel<<<nblocks, 256, 0, s>>>(x, mask, n, ws);
    CB_LAUNCH_CHECK();
    masked_moments_finish_kernel<<<1, 32, 0, s>>>(nblocks, ws, sums);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_normalize_apply(const float* x, const double* sums, float floor_std, float* out, int64_t n,
                                  void* stream) {
    CB_REQUIRE(x && sums && out && n >= 1, CB_E_ARG, "cb_normalize_apply: bad argument");
    int nblocks = (int)std::min<int64_t>(148 * 8, cb::ceil_div<int64_t>(n, 256));
    normalize_apply_kernel<<<nblocks, 256, 0, cb::as_stream(stream)>>>(x, sums, floor_std, out, n);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_rnd_scale_bonus(const float* err, const float* done, const double* rew_var,
                                  float beta, float* out, int64_t n, void* stream) {
    CB_REQUIRE(err && done && rew_var && out, CB_E_ARG, "cb_rnd_scale_bonus: null pointer");
    rnd_scale_bonus_kernel<<<(unsigned)cb::ceil_div<int64_t>(n, 256), 256, 0, cb::as_stream(stream)>>>(
        err, done, rew_var, beta, out, n);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

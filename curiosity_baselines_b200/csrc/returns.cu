// Discounted returns, GAE and the valid mask: one thread per environment column, serial
// over T, coalesced over B.  Mirrors rlpyt/algos/utils.py:8-40,104-112 with the reference's
// fp32 operation order (explicit _rn intrinsics: no FMA contraction), so results are
// bit-identical to the CPU reference.
#include "common.cuh"

namespace {

constexpr int kChunk = 8;   // timesteps prefetched per thread before the serial recurrence

__global__ void gae_kernel(const float* __restrict__ reward, const float* __restrict__ value,
                           const float* __restrict__ done, const float* __restrict__ bootstrap,
                           float discount, float gl, float* __restrict__ adv,
                           float* __restrict__ ret, int T, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float next_v = bootstrap[b];
    float next_adv = 0.f;
    bool last = true;
    for (int t1 = T; t1 > 0; t1 -= kChunk) {
        int t0 = max(t1 - kChunk, 0);
        float r[kChunk], v[kChunk], d[kChunk];
#pragma unroll
        for (int i = 0; i < kChunk; ++i) {
            int t = t1 - 1 - i;
            if (t >= t0) {
                size_t o = (size_t)t * B + b;
                r[i] = __ldg(reward + o); v[i] = __ldg(value + o); d[i] = __ldg(done + o);
            }
        }
#pragma unroll
        for (int i = 0; i < kChunk; ++i) {
            int t = t1 - 1 - i;
            if (t < t0) break;
            float nd = __fsub_rn(1.f, d[i]);
            // delta = r + discount*v' * nd - v
            float delta = __fsub_rn(__fadd_rn(r[i], __fmul_rn(__fmul_rn(discount, next_v), nd)), v[i]);
            float a = last ? delta
                           : __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nd), next_adv));
            size_t o = (size_t)t * B + b;
            adv[o] = a;
            ret[o] = __fadd_rn(a, v[i]);
            next_adv = a; next_v = v[i]; last = false;
        }
    }
}

__global__ void discount_return_kernel(const float* __restrict__ reward,
                                       const float* __restrict__ done,
                                       const float* __restrict__ bootstrap, float discount,
                                       float* __restrict__ ret, int T, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float next = bootstrap[b];
    bool last = true;
    for (int t1 = T; t1 > 0; t1 -= kChunk) {
        int t0 = max(t1 - kChunk, 0);
        float r[kChunk], d[kChunk];
#pragma unroll
        for (int i = 0; i < kChunk; ++i) {
            int t = t1 - 1 - i;
            if (t >= t0) {
                size_t o = (size_t)t * B + b;
                r[i] = __ldg(reward + o); d[i] = __ldg(done + o);
            }
        }
#pragma unroll
        for (int i = 0; i < kChunk; ++i) {
            int t = t1 - 1 - i;
            if (t < t0) break;
            float nd = __fsub_rn(1.f, d[i]);
            // last row: r + discount*bootstrap*nd ; others: r + ret'*discount*nd
            float x = last ? __fmul_rn(__fmul_rn(discount, next), nd)
                           : __fmul_rn(__fmul_rn(next, discount), nd);
            float out = __fadd_rn(r[i], x);
            ret[(size_t)t * B + b] = out;
            next = out; last = false;
        }
    }
}

__global__ void valid_from_done_kernel(const float* __restrict__ done, float* __restrict__ valid,
                                       int T, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float cum = 0.f;
    for (int t = 0; t < T; ++t) {
        size_t o = (size_t)t * B + b;
        valid[o] = __fsub_rn(1.f, fminf(cum, 1.f));
        cum = __fadd_rn(cum, __ldg(done + o));
    }
}

}  // namespace

extern "C" int cb_gae(const float* reward, const float* value, const float* done,
                      const float* bootstrap, float discount, float gae_lambda, float* advantage,
                      float* return_, int T, int B, void* stream) {
    CB_REQUIRE(reward && value && done && bootstrap && advantage && return_, CB_E_ARG, "cb_gae: null pointer");
    CB_REQUIRE(T >= 1 && B >= 1, CB_E_ARG, "cb_gae: T=%d B=%d", T, B);
    // the reference multiplies the Python floats first: discount * gae_lambda in double
    float gl = (float)((double)discount * (double)gae_lambda);
    int threads = 128;
    gae_kernel<<<cb::ceil_div(B, threads), threads, 0, cb::as_stream(stream)>>>(
        reward, value, done, bootstrap, discount, gl, advantage, return_, T, B);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_discount_return(const float* reward, const float* done, const float* bootstrap,
                                  float discount, float* return_, int T, int B, void* stream) {
    CB_REQUIRE(reward && done && bootstrap && return_, CB_E_ARG, "cb_discount_return: null pointer");
    CB_REQUIRE(T >= 1 && B >= 1, CB_E_ARG, "cb_discount_return: T=%d B=%d", T, B);
    int threads = 128;
    discount_return_kernel<<<cb::ceil_div(B, threads), threads, 0, cb::as_stream(stream)>>>(
        reward, done, bootstrap, discount, return_, T, B);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_valid_from_done(const float* done, float* valid, int T, int B, void* stream) {
    CB_REQUIRE(done && valid, CB_E_ARG, "cb_valid_from_done: null pointer");
    CB_REQUIRE(T >= 1 && B >= 1, CB_E_ARG, "cb_valid_from_done: T=%d B=%d", T, B);
    int threads = 128;
    valid_from_done_kernel<<<cb::ceil_div(B, threads), threads, 0, cb::as_stream(stream)>>>(
        done, valid, T, B);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

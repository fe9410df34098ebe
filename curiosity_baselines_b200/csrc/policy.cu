// Policy-side pieces of the PPO update (rlpyt/models/pg/atari_lstm_model.py:114-154,
// rlpyt/algos/pg/ppo.py:192-225, rlpyt/distributions/categorical.py:33-46):
//   * the LSTM cell's gate arithmetic, forward and backward (the two GEMMs of each timestep
//     run on the tcgen05 Linear kernels; these kernels are the fused elementwise part),
//   * softmax over the action logits (+ its backward for the autograd drop-in),
//   * the clipped-surrogate / value / entropy loss with its gradient w.r.t. logits and value.
// All of them are HBM-bound elementwise kernels: 16-byte accesses, one pass over each operand.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// gates layout per env row: [i | f | g | o], each H wide (torch.nn.LSTM chunk order).
// One thread handles 4 consecutive hidden units of one env.
// gx and gates may alias (in-place activation): neither is __restrict__.
__global__ void lstm_cell_fwd_kernel(const float* gx, const float* __restrict__ gh,
                                     const float* __restrict__ c_prev, float* gates,
                                     float* __restrict__ c_out, __nv_bfloat16* __restrict__ h_bf16,
                                     float* __restrict__ h_f32, int B, int H) {
    const int q = H >> 2;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * q) return;
    const int b = (int)(idx / q), j = (int)(idx % q) * 4;
    const size_t row = (size_t)b * 4 * H;
    float4 a[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float4 x = *reinterpret_cast<const float4*>(gx + row + (size_t)g * H + j);
        float4 h = *reinterpret_cast<const float4*>(gh + row + (size_t)g * H + j);
        a[g] = make_float4(x.x + h.x, x.y + h.y, x.z + h.z, x.w + h.w);
    }
    float4 cp = *reinterpret_cast<const float4*>(c_prev + (size_t)b * H + j);
    const float* pi = &a[0].x; const float* pf = &a[1].x; const float* pg = &a[2].x; const float* po = &a[3].x;
    const float* pc = &cp.x;
    float gi[4], gf[4], gg[4], go[4], c[4], h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        gi[k] = sigmoidf_(pi[k]); gf[k] = sigmoidf_(pf[k]); gg[k] = tanhf(pg[k]); go[k] = sigmoidf_(po[k]);
        c[k] = gf[k] * pc[k] + gi[k] * gg[k];
        h[k] = go[k] * tanhf(c[k]);
    }
    *reinterpret_cast<float4*>(gates + row + 0 * (size_t)H + j) = make_float4(gi[0], gi[1], gi[2], gi[3]);
    *reinterpret_cast<float4*>(gates + row + 1 * (size_t)H + j) = make_float4(gf[0], gf[1], gf[2], gf[3]);
    *reinterpret_cast<float4*>(gates + row + 2 * (size_t)H + j) = make_float4(gg[0], gg[1], gg[2], gg[3]);
    *reinterpret_cast<float4*>(gates + row + 3 * (size_t)H + j) = make_float4(go[0], go[1], go[2], go[3]);
    *reinterpret_cast<float4*>(c_out + (size_t)b * H + j) = make_float4(c[0], c[1], c[2], c[3]);
    __nv_bfloat162 h01 = __floats2bfloat162_rn(h[0], h[1]), h23 = __floats2bfloat162_rn(h[2], h[3]);
    uint2 packed;
    packed.x = *reinterpret_cast<uint32_t*>(&h01);
    packed.y = *reinterpret_cast<uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(h_bf16 + (size_t)b * H + j) = packed;
    if (h_f32) *reinterpret_cast<float4*>(h_f32 + (size_t)b * H + j) = make_float4(h[0], h[1], h[2], h[3]);
}

// dh: bf16 gradient arriving at h_t (loss heads + recurrence, already summed); dc_next: fp32 gradient
// arriving at c_t from step t+1 (NULL at the last step).  Writes the pre-activation gate gradients
// (bf16 [B,4H], the operand of the recurrent dgrad and of the batched weight gradients) and dc_prev.
__global__ void lstm_cell_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const float* __restrict__ dc_next,
                                     const float* __restrict__ gates, const float* __restrict__ c_prev,
                                     const float* __restrict__ c, __nv_bfloat16* __restrict__ dgates,
                                     float* __restrict__ dc_prev, int B, int H) {
    const int q = H >> 2;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * q) return;
    const int b = (int)(idx / q), j = (int)(idx % q) * 4;
    const size_t row = (size_t)b * 4 * H;
    float4 vi = *reinterpret_cast<const float4*>(gates + row + 0 * (size_t)H + j);
    float4 vf = *reinterpret_cast<const float4*>(gates + row + 1 * (size_t)H + j);
    float4 vg = *reinterpret_cast<const float4*>(gates + row + 2 * (size_t)H + j);
    float4 vo = *reinterpret_cast<const float4*>(gates + row + 3 * (size_t)H + j);
    float4 vcp = *reinterpret_cast<const float4*>(c_prev + (size_t)b * H + j);
    float4 vc = *reinterpret_cast<const float4*>(c + (size_t)b * H + j);
    float4 vdc = dc_next ? *reinterpret_cast<const float4*>(dc_next + (size_t)b * H + j) : make_float4(0, 0, 0, 0);
    uint2 pk = *reinterpret_cast<const uint2*>(dh + (size_t)b * H + j);
    __nv_bfloat162 d01 = *reinterpret_cast<__nv_bfloat162*>(&pk.x), d23 = *reinterpret_cast<__nv_bfloat162*>(&pk.y);
    float dhv[4] = {__low2float(d01), __high2float(d01), __low2float(d23), __high2float(d23)};
    const float *gi = &vi.x, *gf = &vf.x, *gg = &vg.x, *go = &vo.x, *cp = &vcp.x, *cc = &vc.x, *dcn = &vdc.x;
    float di[4], df[4], dg[4], dO[4], dcp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float tc = tanhf(cc[k]);
        float dc = dcn[k] + dhv[k] * go[k] * (1.f - tc * tc);
        dO[k] = dhv[k] * tc * go[k] * (1.f - go[k]);
        di[k] = dc * gg[k] * gi[k] * (1.f - gi[k]);
        df[k] = dc * cp[k] * gf[k] * (1.f - gf[k]);
        dg[k] = dc * gi[k] * (1.f - gg[k] * gg[k]);
        dcp[k] = dc * gf[k];
    }
    auto store4 = [&](int g, const float* v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), c2 = __floats2bfloat162_rn(v[2], v[3]);
        uint2 p;
        p.x = *reinterpret_cast<uint32_t*>(&a);
        p.y = *reinterpret_cast<uint32_t*>(&c2);
        *reinterpret_cast<uint2*>(dgates + row + (size_t)g * H + j) = p;
    };
    store4(0, di); store4(1, df); store4(2, dg); store4(3, dO);
    *reinterpret_cast<float4*>(dc_prev + (size_t)b * H + j) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
}

constexpr int kMaxA = 64;

// F.softmax(logits, -1) as torch evaluates it in fp32: exp(x - max) / sum.
__global__ void softmax_fwd_kernel(const float* __restrict__ logits, float* __restrict__ prob, int64_t n, int A) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const float* x = logits + m * A;
    float mx = x[0];
    for (int a = 1; a < A; ++a) mx = fmaxf(mx, x[a]);
    float s = 0.f;
    for (int a = 0; a < A; ++a) s += expf(x[a] - mx);
    float inv = 1.f / s;
    for (int a = 0; a < A; ++a) prob[m * A + a] = expf(x[a] - mx) * inv;
}

// dlogits = p .* (dprob - sum_a dprob*p)
__global__ void softmax_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ dprob,
                                   float* __restrict__ dlogits, int64_t n, int A) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    float dot = 0.f;
    for (int a = 0; a < A; ++a) dot += prob[m * A + a] * dprob[m * A + a];
    for (int a = 0; a < A; ++a) dlogits[m * A + a] = prob[m * A + a] * (dprob[m * A + a] - dot);
}

// ppo.py:192-225 per sample.  terms[0..3][m] = surrogate, 0.5*(v-R)^2, entropy, perplexity;
// gradients of  -sum coef*surr + vcoef*sum coef*verr - ecoef*sum coef*entropy  w.r.t. logits / value.
__global__ void ppo_loss_kernel(const float* __restrict__ logits, const float* __restrict__ value,
                                const int64_t* __restrict__ action, const float* __restrict__ old_prob,
                                const float* __restrict__ advantage, const float* __restrict__ return_,
                                const float* __restrict__ coef, float ratio_clip, float vcoef, float ecoef,
                                float* __restrict__ prob_out, float* __restrict__ terms,
                                float* __restrict__ dlogits, float* __restrict__ dvalue, int64_t n, int A) {
    constexpr float EPS = 1e-8f;     // distributions/categorical.py:8
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    float p[kMaxA];
    const float* x = logits + m * A;
    float mx = x[0];
    for (int a = 1; a < A; ++a) mx = fmaxf(mx, x[a]);
    float s = 0.f;
    for (int a = 0; a < A; ++a) { p[a] = expf(x[a] - mx); s += p[a]; }
    float inv = 1.f / s;
    float ent = 0.f;
    for (int a = 0; a < A; ++a) {
        p[a] *= inv;
        ent -= p[a] * logf(p[a] + EPS);
        if (prob_out) prob_out[m * A + a] = p[a];
    }
    const int act = min(max((int)action[m], 0), A - 1);     // an out-of-range index must not fault the context
    const float den = old_prob[m * A + act] + EPS;
    const float ratio = (p[act] + EPS) / den;
    const float adv = advantage[m];
    const float s1 = ratio * adv;
    const float lo = 1.f - ratio_clip, hi = 1.f + ratio_clip;
    const float s2 = fminf(fmaxf(ratio, lo), hi) * adv;
    const float v = value[m], err = v - return_[m];
    terms[m] = fminf(s1, s2);
    terms[n + m] = 0.5f * err * err;
    terms[2 * n + m] = ent;
    terms[3 * n + m] = expf(ent);
    if (!dlogits) return;
    const float c = coef[m];
    // torch.min splits a tie evenly between its operands and clamp passes the gradient inside
    // [lo, hi] (bounds included): inside the range both halves reach `ratio`, outside only surr_1 can.
    const bool inside = ratio >= lo && ratio <= hi;
    float dratio = 0.f;
    if (s1 < s2) dratio = adv;
    else if (s1 == s2) dratio = inside ? adv : 0.5f * adv;
    // dL/dp_a
    float g[kMaxA];
    float dot = 0.f;
    for (int a = 0; a < A; ++a) {
        float ga = ecoef * c * (logf(p[a] + EPS) + p[a] / (p[a] + EPS));     // -ecoef * d(entropy)/dp
        if (a == act) ga += -c * dratio / den;
        g[a] = ga;
        dot += ga * p[a];
    }
    for (int a = 0; a < A; ++a) dlogits[m * A + a] = p[a] * (g[a] - dot);
    dvalue[m] = vcoef * c * err;
}

// x (fp32) -> bf16, optionally added to a second fp32 tensor first (dh = d_pi-path + d_value-path)
__global__ void add_to_bf16_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                   __nv_bfloat16* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = cb::f2bf(a[i] + (b ? b[i] : 0.f));
}

}  // namespace

extern "C" int cb_lstm_cell_fwd(const float* gx, const float* gh, const float* c_prev, float* gates, float* c_out,
                                void* h_bf16, float* h_f32, int32_t B, int32_t H, void* stream) {
    CB_REQUIRE(gx && gh && c_prev && gates && c_out && h_bf16 && B >= 1 && H >= 4 && H % 4 == 0, CB_E_ARG,
               "cb_lstm_cell_fwd: bad argument");
    int64_t total = (int64_t)B * (H / 4);
    lstm_cell_fwd_kernel<<<(unsigned)cb::ceil_div<int64_t>(total, 256), 256, 0, cb::as_stream(stream)>>>(
        gx, gh, c_prev, gates, c_out, (__nv_bfloat16*)h_bf16, h_f32, B, H);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_lstm_cell_bwd(const void* dh_bf16, const float* dc_next, const float* gates, const float* c_prev,
                                const float* c, void* dgates_bf16, float* dc_prev, int32_t B, int32_t H, void* stream) {
    CB_REQUIRE(dh_bf16 && gates && c_prev && c && dgates_bf16 && dc_prev && B >= 1 && H >= 4 && H % 4 == 0, CB_E_ARG,
               "cb_lstm_cell_bwd: bad argument");
    int64_t total = (int64_t)B * (H / 4);
    lstm_cell_bwd_kernel<<<(unsigned)cb::ceil_div<int64_t>(total, 256), 256, 0, cb::as_stream(stream)>>>(
        (const __nv_bfloat16*)dh_bf16, dc_next, gates, c_prev, c, (__nv_bfloat16*)dgates_bf16, dc_prev, B, H);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_softmax_fwd(const float* logits, float* prob, int64_t n, int32_t a, void* stream) {
    CB_REQUIRE(logits && prob && n >= 1 && a >= 1, CB_E_ARG, "cb_softmax_fwd: bad argument");
    softmax_fwd_kernel<<<(unsigned)cb::ceil_div<int64_t>(n, 128), 128, 0, cb::as_stream(stream)>>>(logits, prob, n, a);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_softmax_bwd(const float* prob, const float* dprob, float* dlogits, int64_t n, int32_t a,
                              void* stream) {
    CB_REQUIRE(prob && dprob && dlogits && n >= 1 && a >= 1, CB_E_ARG, "cb_softmax_bwd: bad argument");
    softmax_bwd_kernel<<<(unsigned)cb::ceil_div<int64_t>(n, 128), 128, 0, cb::as_stream(stream)>>>(prob, dprob, dlogits,
                                                                                                  n, a);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_ppo_loss(const float* logits, const float* value, const int64_t* action, const float* old_prob,
                           const float* advantage, const float* return_, const float* coef, float ratio_clip,
                           float value_coeff, float entropy_coeff, float* prob_out, float* terms, float* dlogits,
                           float* dvalue, int64_t n, int32_t a, void* stream) {
    CB_REQUIRE(logits && value && action && old_prob && advantage && return_ && terms && n >= 1 && a >= 1 && a <= kMaxA,
               CB_E_ARG, "cb_ppo_loss: bad argument (1 <= action_size <= 64)");
    CB_REQUIRE(!dlogits || (coef && dvalue), CB_E_ARG, "cb_ppo_loss: gradients need coef and dvalue");
    ppo_loss_kernel<<<(unsigned)cb::ceil_div<int64_t>(n, 128), 128, 0, cb::as_stream(stream)>>>(
        logits, value, action, old_prob, advantage, return_, coef, ratio_clip, value_coeff, entropy_coeff, prob_out,
        terms, dlogits, dvalue, n, a);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_add_to_bf16(const float* a, const float* b, void* out_bf16, int64_t n, void* stream) {
    CB_REQUIRE(a && out_bf16 && n >= 1, CB_E_ARG, "cb_add_to_bf16: bad argument");
    add_to_bf16_kernel<<<(unsigned)cb::ceil_div<int64_t>(n, 256), 256, 0, cb::as_stream(stream)>>>(
        a, b, (__nv_bfloat16*)out_bf16, n);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

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


// ---------------------------------------------------------------------------------------------------------
// pi / value heads (atari_lstm_model.py:144-145): N = A and N = 1 Linear layers over the same [n,H] LSTM output.
// As GEMMs they are 5..8 columns wide -- HBM-bound on reading h -- so they run fused on CUDA cores: one warp per
// row, bf16 weights (the rounding the GEMM engines apply) in shared memory, fp32 accumulation, softmax in the
// same pass.  Algorithmic bytes: forward n*H*2 read; backward n*H*2 read + n*H*2 written.
constexpr int kHeadsMaxO = 8;       // A + 1 <= 8 outputs (larger action sets use the generic Linear path)

__device__ __forceinline__ void bf16x8_to_f32(const uint4& v, float* f) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(p[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

__device__ __forceinline__ void stage_head_weights(__nv_bfloat16* sw, const float* w_pi, const float* w_v, int H, int A) {
    for (int i = threadIdx.x; i < (A + (w_v ? 1 : 0)) * H; i += blockDim.x) {
        int o = i / H, j = i - o * H;
        sw[i] = cb::f2bf(o < A ? w_pi[(size_t)o * H + j] : w_v[j]);
    }
}

__global__ void __launch_bounds__(512) policy_heads_fwd_kernel(
        const __nv_bfloat16* __restrict__ h, const float* __restrict__ w_pi, const float* __restrict__ b_pi,
        const float* __restrict__ w_v, const float* __restrict__ b_v, float* __restrict__ logits,
        float* __restrict__ value, float* __restrict__ prob, int64_t n, int H, int A) {
    extern __shared__ __align__(16) unsigned char heads_smem[];
    __nv_bfloat16* sw = reinterpret_cast<__nv_bfloat16*>(heads_smem);
    stage_head_weights(sw, w_pi, w_v, H, A);
    __syncthreads();
    const int AO = A + (w_v ? 1 : 0), lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const float bias = lane < A ? b_pi[lane] : ((lane == A && w_v) ? b_v[0] : 0.f);
    constexpr int R = 2;            // rows in flight per warp (4 rows at 110 registers / 16 warps per SM measured slower:
                                    // 0.194 vs 0.166 ms at n = 262 144)
    for (int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R; row0 < n; row0 += warps * R) {
        float acc[R][kHeadsMaxO];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int o = 0; o < kHeadsMaxO; ++o) acc[r][o] = 0.f;
        for (int j0 = lane * 8; j0 < H; j0 += 256) {
            uint4 hv[R];
#pragma unroll
            for (int r = 0; r < R; ++r)
                hv[r] = (row0 + r < n) ? __ldg(reinterpret_cast<const uint4*>(h + (size_t)(row0 + r) * H + j0))
                                       : make_uint4(0, 0, 0, 0);
            float hf[R][8];
#pragma unroll
            for (int r = 0; r < R; ++r) bf16x8_to_f32(hv[r], hf[r]);
#pragma unroll
            for (int o = 0; o < kHeadsMaxO; ++o) {
                if (o < AO) {
                    float wf[8];
                    bf16x8_to_f32(*reinterpret_cast<const uint4*>(sw + (size_t)o * H + j0), wf);
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc[r][o] = fmaf(hf[r][k], wf[k], acc[r][o]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float mine = 0.f;
#pragma unroll
            for (int o = 0; o < kHeadsMaxO; ++o) {
                float s = cb::warp_sum(acc[r][o]);
                if (lane == o) mine = s;
            }
            mine += bias;
            // softmax over lanes 0..A-1
            float mx = lane < A ? mine : -INFINITY;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float e = lane < A ? expf(mine - mx) : 0.f;
            float se = cb::warp_sum(e);
            const int64_t row = row0 + r;
            if (row < n) {
                if (lane < A) {
                    logits[row * A + lane] = mine;
                    if (prob) prob[row * A + lane] = e / se;
                } else if (lane == A && w_v) {
                    value[row] = mine;
                }
            }
        }
    }
}

// dh = dlogits W_pi + dvalue w_v (bf16);  dW_pi += dlogits^T h, dw_v += dvalue^T h, db likewise.
// A warp owns one 256-column segment of the rows it visits, so its slice of dW stays in registers.
__global__ void __launch_bounds__(256, 2) policy_heads_bwd_kernel(
        const __nv_bfloat16* __restrict__ h, const float* __restrict__ dlogits, const float* __restrict__ dvalue,
        const float* __restrict__ w_pi, const float* __restrict__ w_v, __nv_bfloat16* __restrict__ dh,
        float* __restrict__ dw_pi, float* __restrict__ db_pi, float* __restrict__ dw_v, float* __restrict__ db_v,
        int64_t n, int H, int A, int h_relu) {
    extern __shared__ __align__(16) unsigned char heads_smem[];
    const int AO = A + (w_v ? 1 : 0), lane = threadIdx.x & 31;
    __nv_bfloat16* sw = reinterpret_cast<__nv_bfloat16*>(heads_smem);
    float* sdw = reinterpret_cast<float*>(heads_smem + (((size_t)AO * H * 2 + 15) & ~(size_t)15));
    float* sdb = sdw + (size_t)AO * H;
    stage_head_weights(sw, w_pi, w_v, H, A);
    for (int i = threadIdx.x; i < AO * H; i += blockDim.x) sdw[i] = 0.f;
    if (threadIdx.x < kHeadsMaxO) sdb[threadIdx.x] = 0.f;
    __syncthreads();
    const int nseg = H >> 8;
    const int64_t gwarp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);      // a multiple of nseg (host guarantees)
    const int seg = (int)(gwarp % nseg);
    const int j0 = seg * 256 + lane * 8;
    float accw[kHeadsMaxO][8], accb[kHeadsMaxO];
#pragma unroll
    for (int o = 0; o < kHeadsMaxO; ++o) {
        accb[o] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) accw[o][k] = 0.f;
    }
    constexpr int R = 2;
    for (int64_t row0 = (gwarp / nseg) * R; row0 < n; row0 += (warps / nseg) * R) {
        uint4 hv[R];
        float d[R][kHeadsMaxO];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = row0 + r;
            const bool ok = row < n;
            hv[r] = ok ? __ldg(reinterpret_cast<const uint4*>(h + (size_t)row * H + j0)) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int o = 0; o < kHeadsMaxO; ++o)
                d[r][o] = !ok ? 0.f : (o < A ? __ldg(dlogits + row * A + o) : ((o == A && w_v) ? __ldg(dvalue + row) : 0.f));
        }
        float hf[R][8], g[R][8];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            bf16x8_to_f32(hv[r], hf[r]);
#pragma unroll
            for (int k = 0; k < 8; ++k) g[r][k] = 0.f;
        }
#pragma unroll
        for (int o = 0; o < kHeadsMaxO; ++o) {
            if (o < AO) {                    // weights of my 8 columns: conflict-free 16-byte shared loads
                float wf[8];
                bf16x8_to_f32(*reinterpret_cast<const uint4*>(sw + (size_t)o * H + j0), wf);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    accb[o] += d[r][o];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        g[r][k] = fmaf(d[r][o], wf[k], g[r][k]);
                        accw[o][k] = fmaf(d[r][o], hf[r][k], accw[o][k]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = row0 + r;
            if (row < n) {
                if (h_relu) {            // h is a ReLU output: the gradient passes only where it was positive
#pragma unroll
                    for (int k = 0; k < 8; ++k) g[r][k] = hf[r][k] > 0.f ? g[r][k] : 0.f;
                }
                __nv_bfloat162 p0 = __floats2bfloat162_rn(g[r][0], g[r][1]), p1 = __floats2bfloat162_rn(g[r][2], g[r][3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(g[r][4], g[r][5]), p3 = __floats2bfloat162_rn(g[r][6], g[r][7]);
                uint4 out;
                out.x = *reinterpret_cast<uint32_t*>(&p0); out.y = *reinterpret_cast<uint32_t*>(&p1);
                out.z = *reinterpret_cast<uint32_t*>(&p2); out.w = *reinterpret_cast<uint32_t*>(&p3);
                *reinterpret_cast<uint4*>(dh + (size_t)row * H + j0) = out;
            }
        }
    }
#pragma unroll
    for (int o = 0; o < kHeadsMaxO; ++o) {
        if (o < AO) {
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(sdw + (size_t)o * H + j0 + k, accw[o][k]);
            if (seg == 0 && lane == 0) atomicAdd(sdb + o, accb[o]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < AO * H; i += blockDim.x) {
        int o = i / H, j = i - o * H;
        atomicAdd(o < A ? dw_pi + (size_t)o * H + j : dw_v + j, sdw[i]);
    }
    if (threadIdx.x < AO) atomicAdd(threadIdx.x < A ? db_pi + threadIdx.x : db_v, sdb[threadIdx.x]);
}


// ---------------------------------------------------------------------------------------------------------
// torch.nn.GRU cell (ndigo.py:79,127; gate order r,z,n):  r = s(gi_r+gh_r), z = s(gi_z+gh_z),
// n = tanh(gi_n + r*gh_n), h' = (1-z)*n + z*h.  gi = x W_ih^T + b_ih (all T at once), gh = h W_hh^T + b_hh (per step).
// gates (may alias gi) <- (r, z, n); ghn <- gh_n (needed by the backward); one thread per 4 hidden units.
__global__ void gru_cell_fwd_kernel(const float* gi, const float* __restrict__ gh, const float* __restrict__ h_prev,
                                    float* gates, float* __restrict__ ghn, float* __restrict__ h_out,
                                    __nv_bfloat16* __restrict__ h_bf16, int B, int G) {
    const int q = G >> 2;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * q) return;
    const int b = (int)(idx / q), j = (int)(idx % q) * 4;
    const size_t row = (size_t)b * 3 * G;
    float4 ir = *reinterpret_cast<const float4*>(gi + row + j), iz = *reinterpret_cast<const float4*>(gi + row + G + j);
    float4 in = *reinterpret_cast<const float4*>(gi + row + 2 * (size_t)G + j);
    float4 hr = *reinterpret_cast<const float4*>(gh + row + j), hz = *reinterpret_cast<const float4*>(gh + row + G + j);
    float4 hn = *reinterpret_cast<const float4*>(gh + row + 2 * (size_t)G + j);
    float4 hp = *reinterpret_cast<const float4*>(h_prev + (size_t)b * G + j);
    const float *pir = &ir.x, *piz = &iz.x, *pin = &in.x, *phr = &hr.x, *phz = &hz.x, *phn = &hn.x, *php = &hp.x;
    float r[4], z[4], n[4], h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        r[k] = sigmoidf_(pir[k] + phr[k]);
        z[k] = sigmoidf_(piz[k] + phz[k]);
        n[k] = tanhf(pin[k] + r[k] * phn[k]);
        h[k] = (1.f - z[k]) * n[k] + z[k] * php[k];
    }
    *reinterpret_cast<float4*>(gates + row + j) = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float4*>(gates + row + G + j) = make_float4(z[0], z[1], z[2], z[3]);
    *reinterpret_cast<float4*>(gates + row + 2 * (size_t)G + j) = make_float4(n[0], n[1], n[2], n[3]);
    *reinterpret_cast<float4*>(ghn + (size_t)b * G + j) = hn;
    *reinterpret_cast<float4*>(h_out + (size_t)b * G + j) = make_float4(h[0], h[1], h[2], h[3]);
    __nv_bfloat162 h01 = __floats2bfloat162_rn(h[0], h[1]), h23 = __floats2bfloat162_rn(h[2], h[3]);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h01); pk.y = *reinterpret_cast<uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(h_bf16 + (size_t)b * G + j) = pk;
}

// dh = dout (fp32, gradient from the layers above at this step) + dh_rec (bf16, from step t+1; NULL at the end).
// dgi / dgh: bf16 [B,3G] gradients w.r.t. the two pre-activation sums; dh_carry (bf16) = dh*z, the part of
// dL/dh_{t-1} that does not go through W_hh (handed to the recurrent dgrad as its dx_add).
__global__ void gru_cell_bwd_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ dh_rec,
                                    const float* __restrict__ gates, const float* __restrict__ ghn,
                                    const float* __restrict__ h_prev, __nv_bfloat16* __restrict__ dgi,
                                    __nv_bfloat16* __restrict__ dgh, __nv_bfloat16* __restrict__ dh_carry, int B, int G) {
    const int q = G >> 2;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * q) return;
    const int b = (int)(idx / q), j = (int)(idx % q) * 4;
    const size_t row = (size_t)b * 3 * G, col = (size_t)b * G + j;
    float4 vr = *reinterpret_cast<const float4*>(gates + row + j), vz = *reinterpret_cast<const float4*>(gates + row + G + j);
    float4 vn = *reinterpret_cast<const float4*>(gates + row + 2 * (size_t)G + j);
    float4 vhn = *reinterpret_cast<const float4*>(ghn + col), vhp = *reinterpret_cast<const float4*>(h_prev + col);
    float4 vd = dout ? *reinterpret_cast<const float4*>(dout + col) : make_float4(0, 0, 0, 0);
    float dh[4] = {vd.x, vd.y, vd.z, vd.w};
    if (dh_rec) {
        uint2 pk = *reinterpret_cast<const uint2*>(dh_rec + col);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&pk.x), c = *reinterpret_cast<__nv_bfloat162*>(&pk.y);
        dh[0] += __low2float(a); dh[1] += __high2float(a); dh[2] += __low2float(c); dh[3] += __high2float(c);
    }
    const float *r = &vr.x, *z = &vz.x, *n = &vn.x, *hn = &vhn.x, *hp = &vhp.x;
    float dar[4], daz[4], dan[4], dhn[4], carry[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        dan[k] = dh[k] * (1.f - z[k]) * (1.f - n[k] * n[k]);
        daz[k] = dh[k] * (hp[k] - n[k]) * z[k] * (1.f - z[k]);
        dar[k] = dan[k] * hn[k] * r[k] * (1.f - r[k]);
        dhn[k] = dan[k] * r[k];
        carry[k] = dh[k] * z[k];
    }
    auto store4 = [&](__nv_bfloat16* dst, const float* v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), c = __floats2bfloat162_rn(v[2], v[3]);
        uint2 p;
        p.x = *reinterpret_cast<uint32_t*>(&a); p.y = *reinterpret_cast<uint32_t*>(&c);
        *reinterpret_cast<uint2*>(dst) = p;
    };
    store4(dgi + row + j, dar); store4(dgi + row + G + j, daz); store4(dgi + row + 2 * (size_t)G + j, dan);
    store4(dgh + row + j, dar); store4(dgh + row + G + j, daz); store4(dgh + row + 2 * (size_t)G + j, dhn);
    store4(dh_carry + col, carry);
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

// `outputs` = all output columns (A logits + 1 value, or A when there is no value head)
extern "C" int cb_policy_heads_supported(int32_t H, int32_t outputs) {
    return (outputs >= 1 && outputs <= kHeadsMaxO && H >= 256 && H % 256 == 0 && H <= 2048) ? 1 : 0;
}

extern "C" int cb_policy_heads_fwd(const void* h_bf16, const float* w_pi, const float* b_pi, const float* w_v,
                                   const float* b_v, float* logits, float* value, float* prob, int64_t n, int32_t H,
                                   int32_t A, void* stream) {
    CB_REQUIRE(h_bf16 && w_pi && b_pi && logits && n >= 1 && A >= 1, CB_E_ARG, "cb_policy_heads_fwd: bad argument");
    CB_REQUIRE(!w_v || (b_v && value), CB_E_ARG, "cb_policy_heads_fwd: the value head needs b_v and value");
    CB_REQUIRE(cb_policy_heads_supported(H, A + (w_v ? 1 : 0)), CB_E_ARG,
               "cb_policy_heads_fwd: needs <= 8 output columns and H a multiple of 256");
    CB_REQUIRE((((uintptr_t)h_bf16) & 15) == 0, CB_E_ALIGN, "cb_policy_heads_fwd: h must be 16-byte aligned");
    size_t smem = (size_t)(A + (w_v ? 1 : 0)) * H * 2;
    int blocks = (int)std::min<int64_t>(148 * 2, cb::ceil_div<int64_t>(n, 16 * 2));
    policy_heads_fwd_kernel<<<blocks, 512, smem, cb::as_stream(stream)>>>(
        (const __nv_bfloat16*)h_bf16, w_pi, b_pi, w_v, b_v, logits, value, prob, n, H, A);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_policy_heads_bwd(const void* h_bf16, const float* dlogits, const float* dvalue, const float* w_pi,
                                   const float* w_v, void* dh_bf16, float* dw_pi, float* db_pi, float* dw_v,
                                   float* db_v, int64_t n, int32_t H, int32_t A, int32_t h_act, void* stream) {
    CB_REQUIRE(h_bf16 && dlogits && w_pi && dh_bf16 && dw_pi && db_pi && n >= 1 && A >= 1, CB_E_ARG,
               "cb_policy_heads_bwd: bad argument");
    CB_REQUIRE(!w_v || (dvalue && dw_v && db_v), CB_E_ARG, "cb_policy_heads_bwd: the value head needs dvalue, dw_v, db_v");
    CB_REQUIRE(h_act == CB_ACT_NONE || h_act == CB_ACT_RELU, CB_E_ARG, "cb_policy_heads_bwd: h_act must be none or relu");
    CB_REQUIRE(cb_policy_heads_supported(H, A + (w_v ? 1 : 0)), CB_E_ARG,
               "cb_policy_heads_bwd: needs <= 8 output columns and H a multiple of 256");
    CB_REQUIRE(((((uintptr_t)h_bf16) | ((uintptr_t)dh_bf16)) & 15) == 0, CB_E_ALIGN,
               "cb_policy_heads_bwd: h / dh must be 16-byte aligned");
    cudaStream_t s = cb::as_stream(stream);
    CB_CUDA(cudaMemsetAsync(dw_pi, 0, sizeof(float) * (size_t)A * H, s));
    CB_CUDA(cudaMemsetAsync(db_pi, 0, sizeof(float) * (size_t)A, s));
    if (w_v) {
        CB_CUDA(cudaMemsetAsync(dw_v, 0, sizeof(float) * (size_t)H, s));
        CB_CUDA(cudaMemsetAsync(db_v, 0, sizeof(float), s));
    }
    const int AO = A + (w_v ? 1 : 0), nseg = H / 256;
    size_t smem = (((size_t)AO * H * 2 + 15) & ~(size_t)15) + sizeof(float) * ((size_t)AO * H + kHeadsMaxO);
    static bool attr_set = false;
    if (!attr_set) {
        CB_CUDA(cudaFuncSetAttribute(policy_heads_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    CB_REQUIRE(smem <= 100 * 1024, CB_E_ARG, "cb_policy_heads_bwd: H too large");
    // 8 warps per block; 8 % nseg == 0 for nseg in {1,2,4,8} so every block covers whole rows
    CB_REQUIRE(8 % nseg == 0, CB_E_ARG, "cb_policy_heads_bwd: H/256 must divide 8");
    int blocks = (int)std::min<int64_t>(148 * 2, cb::ceil_div<int64_t>(n * nseg, 8 * 2));
    policy_heads_bwd_kernel<<<blocks, 256, smem, s>>>((const __nv_bfloat16*)h_bf16, dlogits, dvalue, w_pi, w_v,
                                                      (__nv_bfloat16*)dh_bf16, dw_pi, db_pi, dw_v, db_v, n, H, A,
                                                      h_act == CB_ACT_RELU ? 1 : 0);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// The recurrences over T as native host loops: 2 launches per timestep issued from C++ (a Python / ctypes round
// trip per launch costs ~15 us, more than either kernel runs).  The per-step GEMMs go through the library's own
// Linear entry points (tcgen05 where the shape has a specialisation).
static void linear_desc(cb_conv_desc* d, int n, int in_c, int out_c, int engine) {
    *d = cb_conv_desc{};
    d->n = n; d->in_h = d->in_w = 1; d->in_c = in_c;
    d->k_h = d->k_w = d->stride = 1; d->pad = 0;
    d->out_h = d->out_w = 1; d->out_c = out_c;
    d->in_dtype = CB_BF16;
    d->in_stride_n = d->in_stride_h = d->in_stride_w = in_c; d->in_stride_c = 1;
    d->act = CB_ACT_NONE; d->engine = engine;
}

extern "C" int cb_lstm_sequence_fwd(int32_t T, int32_t B, int32_t H, float* gates, const void* w_hh_bf16,
                                    const float* b_hh, void* h_all_bf16, float* c_all, float* gh_ws, float* hn,
                                    int32_t engine, void* stream) {
    CB_REQUIRE(T >= 1 && B >= 1 && H >= 8 && H % 8 == 0 && gates && w_hh_bf16 && b_hh && h_all_bf16 && c_all && gh_ws,
               CB_E_ARG, "cb_lstm_sequence_fwd: bad argument");
    cb_conv_desc d; linear_desc(&d, B, H, 4 * H, engine);
    __nv_bfloat16* h_all = (__nv_bfloat16*)h_all_bf16;
    const size_t bh = (size_t)B * H;
    for (int t = 0; t < T; ++t) {
        cb_epilogue e{};
        e.bias = b_hh; e.out_f32 = gh_ws;
        int rc = cb_conv_fwd(&d, h_all + t * bh, w_hh_bf16, &e, stream);
        if (rc) return rc;
        float* g = gates + (size_t)t * 4 * bh;
        rc = cb_lstm_cell_fwd(g, gh_ws, c_all + t * bh, g, c_all + (t + 1) * bh, h_all + (t + 1) * bh,
                              (hn && t == T - 1) ? hn : nullptr, B, H, stream);
        if (rc) return rc;
    }
    return CB_OK;
}

extern "C" int cb_lstm_sequence_bwd(int32_t T, int32_t B, int32_t H, const float* gates, const float* c_all,
                                    const void* dh_out_bf16, const void* w_hh_t_bf16, void* dgates_bf16,
                                    float* dc_ws, void* dh_ws_bf16, int32_t engine, void* stream) {
    CB_REQUIRE(T >= 1 && B >= 1 && H >= 8 && H % 8 == 0 && gates && c_all && dh_out_bf16 && w_hh_t_bf16 && dgates_bf16 &&
               dc_ws && dh_ws_bf16, CB_E_ARG, "cb_lstm_sequence_bwd: bad argument");
    cb_conv_desc d; linear_desc(&d, B, H, 4 * H, engine);
    const __nv_bfloat16* dh_out = (const __nv_bfloat16*)dh_out_bf16;
    __nv_bfloat16* dgates = (__nv_bfloat16*)dgates_bf16;
    __nv_bfloat16* dh_ws = (__nv_bfloat16*)dh_ws_bf16;
    const size_t bh = (size_t)B * H;
    const __nv_bfloat16* dh = dh_out + (size_t)(T - 1) * bh;
    for (int t = T - 1; t >= 0; --t) {
        int rc = cb_lstm_cell_bwd(dh, t < T - 1 ? dc_ws + (size_t)((t + 1) & 1) * bh : nullptr, gates + (size_t)t * 4 * bh,
                                  c_all + t * bh, c_all + (t + 1) * bh, dgates + (size_t)t * 4 * bh,
                                  dc_ws + (size_t)(t & 1) * bh, B, H, stream);
        if (rc) return rc;
        if (t > 0) {            // dL/dh_{t-1} = dgates_t W_hh + (gradient from the heads at t-1)
            __nv_bfloat16* out = dh_ws + (size_t)(t & 1) * bh;
            rc = cb_conv_dgrad(&d, dgates + (size_t)t * 4 * bh, w_hh_t_bf16, nullptr, CB_ACT_NONE,
                               dh_out + (size_t)(t - 1) * bh, out, stream);
            if (rc) return rc;
            dh = out;
        }
    }
    return CB_OK;
}

extern "C" int cb_gru_cell_fwd(const float* gi, const float* gh, const float* h_prev, float* gates, float* ghn,
                               float* h_out, void* h_bf16, int32_t B, int32_t G, void* stream) {
    CB_REQUIRE(gi && gh && h_prev && gates && ghn && h_out && h_bf16 && B >= 1 && G >= 4 && G % 4 == 0, CB_E_ARG,
               "cb_gru_cell_fwd: bad argument");
    int64_t total = (int64_t)B * (G / 4);
    gru_cell_fwd_kernel<<<(unsigned)cb::ceil_div<int64_t>(total, 256), 256, 0, cb::as_stream(stream)>>>(
        gi, gh, h_prev, gates, ghn, h_out, (__nv_bfloat16*)h_bf16, B, G);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_gru_cell_bwd(const float* dout, const void* dh_rec_bf16, const float* gates, const float* ghn,
                               const float* h_prev, void* dgi_bf16, void* dgh_bf16, void* dh_carry_bf16, int32_t B,
                               int32_t G, void* stream) {
    CB_REQUIRE((dout || dh_rec_bf16) && gates && ghn && h_prev && dgi_bf16 && dgh_bf16 && dh_carry_bf16 && B >= 1 &&
               G >= 4 && G % 4 == 0, CB_E_ARG, "cb_gru_cell_bwd: bad argument");
    int64_t total = (int64_t)B * (G / 4);
    gru_cell_bwd_kernel<<<(unsigned)cb::ceil_div<int64_t>(total, 256), 256, 0, cb::as_stream(stream)>>>(
        dout, (const __nv_bfloat16*)dh_rec_bf16, gates, ghn, h_prev, (__nv_bfloat16*)dgi_bf16, (__nv_bfloat16*)dgh_bf16,
        (__nv_bfloat16*)dh_carry_bf16, B, G);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

extern "C" int cb_gru_sequence_fwd(int32_t T, int32_t B, int32_t G, float* gates, const void* w_hh_bf16,
                                   const float* b_hh, float* h_all, void* h_all_bf16, float* ghn_all, float* gh_ws,
                                   int32_t engine, void* stream) {
    CB_REQUIRE(T >= 1 && B >= 1 && G >= 8 && G % 8 == 0 && gates && w_hh_bf16 && b_hh && h_all && h_all_bf16 && ghn_all &&
               gh_ws, CB_E_ARG, "cb_gru_sequence_fwd: bad argument");
    cb_conv_desc d; linear_desc(&d, B, G, 3 * G, engine);
    __nv_bfloat16* hb = (__nv_bfloat16*)h_all_bf16;
    const size_t bg = (size_t)B * G;
    for (int t = 0; t < T; ++t) {
        cb_epilogue e{};
        e.bias = b_hh; e.out_f32 = gh_ws;
        int rc = cb_conv_fwd(&d, hb + t * bg, w_hh_bf16, &e, stream);
        if (rc) return rc;
        float* g = gates + (size_t)t * 3 * bg;
        rc = cb_gru_cell_fwd(g, gh_ws, h_all + t * bg, g, ghn_all + t * bg, h_all + (t + 1) * bg, hb + (t + 1) * bg, B, G,
                             stream);
        if (rc) return rc;
    }
    return CB_OK;
}

extern "C" int cb_gru_sequence_bwd(int32_t T, int32_t B, int32_t G, const float* gates, const float* ghn_all,
                                   const float* h_all, const float* dout, const void* w_hh_t_bf16, void* dgi_bf16,
                                   void* dgh_bf16, void* dh_ws_bf16, int32_t engine, void* stream) {
    CB_REQUIRE(T >= 1 && B >= 1 && G >= 8 && G % 8 == 0 && gates && ghn_all && h_all && dout && w_hh_t_bf16 && dgi_bf16 &&
               dgh_bf16 && dh_ws_bf16, CB_E_ARG, "cb_gru_sequence_bwd: bad argument");
    cb_conv_desc d; linear_desc(&d, B, G, 3 * G, engine);
    __nv_bfloat16* dgi = (__nv_bfloat16*)dgi_bf16;
    __nv_bfloat16* dgh = (__nv_bfloat16*)dgh_bf16;
    __nv_bfloat16* ws = (__nv_bfloat16*)dh_ws_bf16;           // [3,B,G]: carry + two ping-pong dh buffers
    const size_t bg = (size_t)B * G;
    const __nv_bfloat16* dh_rec = nullptr;
    for (int t = T - 1; t >= 0; --t) {
        int rc = cb_gru_cell_bwd(dout + t * bg, dh_rec, gates + (size_t)t * 3 * bg, ghn_all + t * bg, h_all + t * bg,
                                 dgi + (size_t)t * 3 * bg, dgh + (size_t)t * 3 * bg, ws, B, G, stream);
        if (rc) return rc;
        if (t > 0) {            // dL/dh_{t-1} (recurrent part) = dgh_t W_hh + dh_t * z_t
            __nv_bfloat16* out = ws + (size_t)(1 + (t & 1)) * bg;
            rc = cb_conv_dgrad(&d, dgh + (size_t)t * 3 * bg, w_hh_t_bf16, nullptr, CB_ACT_NONE, ws, out, stream);
            if (rc) return rc;
            dh_rec = out;
        }
    }
    return CB_OK;
}

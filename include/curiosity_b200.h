/*
 * curiosity_b200.h -- C ABI of libcuriosity_b200.so: the B200 (sm_100a) implementation of
 * the intrinsic-reward hot path of Improbable-AI/curiosity_baselines.
 *
 * The reference has no native ABI for this path: it is a Python object protocol
 * (`curiosity_model.compute_bonus/compute_loss`, the free functions in
 * rlpyt/algos/utils.py).  Each entry point below names the reference lines whose
 * arithmetic it replaces; the Python host in curiosity_baselines_b200/ sequences them
 * behind the reference's own class/function signatures.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - plain pointers and sizes only, no framework types; `stream` is a cudaStream_t
 *     passed as void* (NULL = default stream);
 *   - the library never allocates or frees caller memory; scratch space is passed in
 *     (`ws`, sized by the matching *_workspace_bytes query);
 *   - returns 0 on success, <0 on error (CB_E_*); cb_last_error() gives the message of
 *     the last failing call on the calling thread;
 *   - launches are asynchronous on `stream`; no global mutable state, re-entrant per
 *     stream; one CUDA context (the primary context) per process;
 *   - layouts: rollout tensors are time-major [T,B,...]; frames are uint8 NCHW exactly as
 *     the sampler writes them; activations between layers are bf16 NHWC; "prepared"
 *     weights are bf16 [C_out][k_h][k_w][C_in] (see cb_prepare_* in the Python host).
 */
#ifndef CURIOSITY_B200_H
#define CURIOSITY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_OK            0
#define CB_E_ARG        -1   /* bad shape / null pointer / unsupported combination */
#define CB_E_ALIGN      -2   /* pointer or stride not aligned as the kernel requires */
#define CB_E_CUDA       -3   /* a CUDA runtime / driver call failed */
#define CB_E_ARCH       -4   /* device is not sm_100 (tcgen05 kernels) */
#define CB_E_WORKSPACE  -5   /* workspace too small */

/* element types of gathered inputs */
#define CB_U8    0
#define CB_BF16  1
#define CB_F32   2

/* activations (encoders.py / icm.py / rnd.py) */
#define CB_ACT_NONE   0
#define CB_ACT_RELU   1
#define CB_ACT_LRELU  2      /* negative slope 0.01, torch default */
#define CB_ACT_ELU    3      /* alpha 1 */

/* execution engine for the GEMM-shaped ops */
#define CB_ENGINE_AUTO  0    /* tcgen05 where a specialisation exists, else SIMT */
#define CB_ENGINE_SIMT  1    /* generic CUDA-core implicit GEMM (any shape) */
#define CB_ENGINE_UMMA  2    /* tcgen05/TMEM/TMA; CB_E_ARG if shape unsupported */

const char* cb_last_error(void);
int         cb_version(void);
/* 1 if the current device can run the tcgen05 kernels (compute capability 10.x) */
int         cb_device_supports_umma(void);

/* ------------------------------------------------------------------------------------
 * Returns / advantages over T, parallel over B.     rlpyt/algos/utils.py
 * All tensors fp32 [T,B] row-major (b contiguous); bootstrap [B]; done is 0/1 float.
 * Bit-exact with the reference's fp32 operation order (no FMA contraction).
 * ---------------------------------------------------------------------------------- */
/* algos/utils.py:8-21 */
int cb_discount_return(const float* reward, const float* done, const float* bootstrap,
                       float discount, float* return_, int T, int B, void* stream);
/* algos/utils.py:24-40 */
int cb_gae(const float* reward, const float* value, const float* done, const float* bootstrap,
           float discount, float gae_lambda, float* advantage, float* return_,
           int T, int B, void* stream);
/* algos/utils.py:104-112 */
int cb_valid_from_done(const float* done, float* valid, int T, int B, void* stream);

/* ------------------------------------------------------------------------------------
 * Running statistics.     rlpyt/utils/averages.py:41-89, models/curiosity/rnd.py:150-203
 * A RunningMeanStd state is three device arrays: mean[n] f64, var[n] f64, count[1] f64.
 * ---------------------------------------------------------------------------------- */
/* rnd.py:150-152: n_valid[b] = sum_t |done[t,b]-1| (int32 [B]); also total -> *n_total */
int cb_valid_lengths(const float* done, int32_t* n_valid, int64_t* n_total, int T, int B,
                     void* stream);
/* rnd.py:153-160 + averages.py:48-52: exact integer per-pixel sums over the frames
 * (t,b) with t < n_valid[b].  frames: uint8, frame (t,b) starts at
 * frames + (t*B+b)*frame_stride bytes and has `pixels` bytes (pixels % 16 == 0,
 * frame_stride % 16 == 0).  sum/sumsq: uint64 [pixels], must be zeroed by the caller. */
int cb_obs_moments_u8(const uint8_t* frames, int64_t frame_stride, int pixels,
                      const int32_t* n_valid, int T, int B,
                      unsigned long long* sum, unsigned long long* sumsq, void* stream);
/* averages.py:54-68 (Chan merge) from the integer sums; also emits fp32 mean and
 * 1/sqrt(var) (rnd.py:162-169) for the conv-1 loader. */
int cb_rms_merge_from_sums(const unsigned long long* sum, const unsigned long long* sumsq,
                           const int64_t* n_total, double* mean, double* var, double* count,
                           float* mean_f32, float* rstd_f32, int pixels, void* stream);
/* differ[0] += number of 16-byte words in which the uint8 frames a[f] and b[f] differ (f < frames; strides in bytes,
 * pixels % 16 == 0).  Guard for reusing features of a batch: the reference hands the SAME next_observation to
 * RND.compute_bonus and RND.compute_loss (algos/pg/base.py:72-73, ppo.py:107-110). */
int cb_frames_differ(const void* a, int64_t a_stride, const void* b, int64_t b_stride, int32_t pixels,
                     int64_t frames, uint32_t* differ, void* stream);
/* averages.py:75-89 driven as rnd.py:186-189: per-env discounted running sum over T with
 * the not-done mask |done-1|; rewems[B] persists across calls; *initialized is 0 before the first
 * call (first row seeds all envs unmasked) and set to 1.  total[T,B] receives the copy
 * returned at each step. */
int cb_reward_filter(const float* rews, const float* done, float gamma, float* rewems,
                     int32_t* initialized, float* total, int T, int B, void* stream);
/* rnd.py:190-193 / algos/pg/base.py:77-81, stage 1: sums = {sum x, sum x^2, n} in f64
 * (deterministic; ws >= 2048 f64).  A data-parallel host all-reduces `sums` before stage 2. */
int cb_moments_partial(const float* x, int64_t n, double* sums, double* ws, void* stream);
/* stage 2, averages.py:54-68: merge mean/var from `sums` into the scalar running state with
 * batch_count = *count_num / count_den (rnd.py:190-193: mean valid length), or n when
 * count_num is NULL (algos/pg/base.py:81). */
int cb_rms_merge_moments(const double* sums, const double* count_num, double count_den,
                         double* mean, double* var, double* count, void* stream);
/* out[0] = sum |done-1| over n elements (numerator of the mean valid length, rnd.py:190) */
int cb_valid_count(const float* done, double* out, int64_t n, void* stream);
/* algos/pg/base.py:95-103: out = (x - mean_m) / max(std_m, floor) with mean/unbiased std over
 * the elements where mask > 0 (all when mask is NULL) */
int cb_masked_normalize(const float* x, const float* mask, float floor_std, float* out,
                        int64_t n, void* stream);
/* averages.py:48-68 for a vector-shaped state: x f64 [n, d] row-major; batch mean / population variance per column
 * (two passes, like numpy) Chan-merged into mean[d], var[d]; count[1] += n. */
int cb_rms_update_columns(const double* x, int64_t n, int32_t d, double* mean, double* var, double* count,
                          void* stream);
/* algos/pg/base.py:95-103 in two stages so that a data-parallel host can all-reduce in between:
 * sums = {sum x, sum x^2, count} over the elements with mask > 0 (all when mask is NULL), f64, deterministic
 * (ws >= 1536 f64); then out = (x - mean) / max(std_unbiased, floor) from `sums`. */
int cb_masked_moments(const float* x, const float* mask, int64_t n, double* sums, double* ws, void* stream);
int cb_normalize_apply(const float* x, const double* sums, float floor_std, float* out, int64_t n, void* stream);
/* rnd.py:198-203: r = beta * err / sqrt(var) * |done-1| */
int cb_rnd_scale_bonus(const float* err, const float* done, const double* rew_var,
                       float beta, float* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------
 * GEMM-shaped layers (convolution as implicit GEMM; Linear is the 1x1 case).
 * encoders.py:7-119, icm.py:9-43,88-92, rnd.py:54-117.
 * ---------------------------------------------------------------------------------- */
typedef struct cb_conv_desc {
    int32_t n;                       /* samples (T*B) */
    int32_t in_h, in_w, in_c;        /* input extent */
    int32_t k_h, k_w, stride, pad;   /* filter */
    int32_t out_h, out_w, out_c;     /* output extent (NHWC bf16) */
    int32_t in_dtype;                /* CB_U8 or CB_BF16 */
    int64_t in_stride_n, in_stride_h, in_stride_w, in_stride_c;  /* in elements */
    int32_t act;                     /* activation applied to the output */
    int32_t engine;                  /* CB_ENGINE_* */
} cb_conv_desc;

typedef struct cb_epilogue {
    const float*   bias;             /* [out_c] or NULL */
    const void*    residual;         /* bf16 [M,out_c] added after the activation, or NULL */
    const float*   side;             /* fp32 [M,side_dim]: extra input columns (one-hot action) */
    const float*   side_w;           /* fp32 [out_c,side_dim]: their weights (icm.py:19-21) */
    int32_t        side_dim;
    const float*   norm_mean;        /* CB_U8 input only: per-pixel mean  [in_h*in_w] (rnd.py:169) */
    const float*   norm_rstd;        /* per-pixel 1/sqrt(var); input clamped to [-5,5] */
    void*          out_bf16;         /* bf16 [M,out_c] or NULL */
    float*         out_f32;          /* fp32 [M,out_c] or NULL */
    const void*    side_bf16;        /* optional, Linear layers on tcgen05 with K % 64 == 0: the side input as bf16
                                        [M,64] (zero padded).  When given, `w` must be [out_c, K+64] with the side
                                        weights in columns K..K+side_dim-1 (zero padded): the concatenation of
                                        icm.py:19-21 then runs on the tensor cores as one more k-block and
                                        side / side_w are not read. */
} cb_epilogue;

/* y = act(conv(x, w) + bias + side*side_w) + residual.   w: bf16 [out_c][k_h][k_w][in_c] */
int cb_conv_fwd(const cb_conv_desc* d, const void* x, const void* w, const cb_epilogue* e,
                void* stream);
/* dx = (dy (*) w + dx_add) .* act'(x_saved).  dy: bf16 [n,out_h,out_w,out_c];
 * w_t: bf16 [in_c][k_h][k_w][out_c]; x_saved: the bf16 NHWC *output* of the layer that
 * produced x (its activation x_act decides act'), or NULL; dx_add: bf16 gradient arriving
 * at x over a skip connection (icm.py:22), or NULL; dx: bf16 NHWC [n,in_h,in_w,in_c] */
int cb_conv_dgrad(const cb_conv_desc* d, const void* dy, const void* w_t, const void* x_saved,
                  int32_t x_act, const void* dx_add, void* dx, void* stream);
/* tcgen05-only dgrad of a k=4, stride=2, pad=0, in_c=32 convolution (Burda conv2): all four
 * output parities in one GEMM with a depth-to-space store.  w_shuffle: bf16
 * [(ph,pw,ci)][(u,v,co)] = W[co][2u+ph][2v+pw][ci].  cb_conv_dgrad_s2_supported() tells whether
 * the shape/device qualifies (otherwise use cb_conv_dgrad). */
int cb_conv_dgrad_s2_supported(const cb_conv_desc* d);
int cb_conv_dgrad_s2(const cb_conv_desc* d, const void* dy, const void* w_shuffle,
                     const void* x_saved, int32_t x_act, const void* dx_add, void* dx, void* stream);
/* tcgen05 first convolution of the Burda / RND encoders: 8x8 stride 4, uint8 NCHW frames read in
 * place (in_stride_w == 1, in_stride_h == in_w), 1..4 input channels, 32 output channels; the RND
 * normalisation (e->norm_mean / norm_rstd) is applied while the patch tile is built.
 * w_native / dw_native use torch's own [32][in_c][8][8] order (bf16 / fp32).  A second network
 * sharing the input (w2_native, bias2, out2_bf16; NULL to skip) is computed in the same pass:
 * RND's target and predictor read the same normalised frame (rnd.py:174-177). */
int cb_conv1_supported(const cb_conv_desc* d);
int cb_conv1_fwd(const cb_conv_desc* d, const void* x, const void* w_native, const cb_epilogue* e,
                 const void* w2_native, const float* bias2, void* out2_bf16, void* stream);
/* BatchNorm2d of the curiosity encoders (rlpyt/models/curiosity/encoders.py:23-25, 87-96, `-batch_norm`) on bf16 NHWC
 * activations seen as [rows, c] (rows = N*H*W, c = 8 * a power of two <= 256), torch's training-mode arithmetic as
 * HBM-bound passes; the per-channel scalars in between are the caller's (c-element tensors).
 *   cb_bn_moments      sum[c] += sum_r x, sumsq[c] += sum_r x^2              (double; caller zeroes)
 *   cb_bn_apply        out = x * scale[c] + shift[c]                         (bf16 and / or fp32 output)
 *   cb_bn_bwd_moments  sum_dz[c] += sum_r dz, sum_dz_xhat[c] += sum_r dz * (x - mean[c]) * invstd[c]
 *   cb_bn_bwd_apply    dx = (dz*coef_dz[c] + x*coef_x[c] + coef_1[c]) * act'(x), x being the OUTPUT of activation
 *                      x_act in front of the normalisation (CB_ACT_*) */
int cb_bn_moments(const void* x_bf16, int64_t rows, int32_t c, double* sum, double* sumsq, void* stream);
int cb_bn_apply(const void* x_bf16, int64_t rows, int32_t c, const float* scale, const float* shift,
                void* out_bf16, float* out_f32, void* stream);
int cb_bn_bwd_moments(const void* dz_bf16, const void* x_bf16, int64_t rows, int32_t c, const float* mean,
                      const float* invstd, double* sum_dz, double* sum_dz_xhat, void* stream);
int cb_bn_bwd_apply(const void* dz_bf16, const void* x_bf16, int64_t rows, int32_t c, const float* coef_dz,
                    const float* coef_x, const float* coef_1, int32_t x_act, void* dx_bf16, void* stream);
/* RND trunk, first TWO convolutions (rnd.py:40-46 / 55-61: Conv 8x8 s4 1->32, LeakyReLU, Conv 4x4 s2 32->64,
 * LeakyReLU) with the 20x20x32 activation between them kept in shared memory.
 *   cb_rnd_normalize_s2d: the newest 84x84 uint8 frame of each sample (read in place at x + i*stride_n) is normalised
 *     with mean / rstd [7056], clamped to +-5 (rnd.py:169-170) and written as a bf16 image in 4x4-pixel-block order,
 *     n * 14112 bytes (21 x 21 blocks of 16 bf16) -- once for both networks (target and predictor read the same normalised frame).
 *   cb_rnd_conv12_fwd: one network over those images.  a1_out_bf16 (NULL to skip) additionally receives the
 *     intermediate activation ([n,20,20,32], needed by the second layer's weight gradient).  w1: bf16 [32][64] in
 *     torch's (kh, kw) order, w2: bf16 [64][(kh,kw,c) = 512] (the cb_prepare_conv_weight forward layout),
 *     a2_out_bf16: [n,9,9,64]. */
int cb_rnd_conv12_supported(void);
int cb_rnd_normalize_s2d(const uint8_t* x, int64_t stride_n, int32_t n, const float* mean, const float* rstd,
                         void* out_s2d, void* stream);
int cb_rnd_conv12_fwd(const void* img_s2d, int32_t n, const void* w1_bf16, const float* b1, const void* w2_bf16,
                      const float* b2, void* a1_out_bf16, void* a2_out_bf16, void* stream);
/* Target and predictor of every sample in one launch: the block image is read once and the first layer is a single
 * N = 64 problem for both networks.  a1_p_out (NULL to skip): the predictor's first-layer activation. */
int cb_rnd_conv12_twin_fwd(const void* img_s2d, int32_t n, const void* w1_t, const float* b1_t, const void* w2_t,
                           const float* b2_t, const void* w1_p, const float* b1_p, const void* w2_p, const float* b2_p,
                           void* a2_t_out, void* a1_p_out, void* a2_p_out, void* stream);
int cb_conv1_wgrad(const cb_conv_desc* d, const void* x, const void* dy, const float* norm_mean,
                   const float* norm_rstd, float* dw_native, float* dbias, float beta, void* stream);
/* dw[out_c][k_h][k_w][in_c] (fp32) = beta*dw + sum_positions dy^T x;  dbias[out_c] likewise
 * (NULL to skip).  x addressed as in cb_conv_fwd (norm_* honoured for CB_U8). */
int cb_conv_wgrad(const cb_conv_desc* d, const void* x, const void* dy, const float* norm_mean,
                  const float* norm_rstd, float* dw, float* dbias, float beta, void* stream);
/* Grouped Linear layers: `groups` independent Linear(k -> f) problems over the same n rows in ONE tcgen05
 * launch (the ensemble of forward models of Disagreement, disagreement.py:97-100,127-130 -- the reference runs
 * them one after the other).  Activations of all problems are packed side by side, row-major [n, groups*f];
 * weights are stacked along the output dimension: w_all bf16 [groups*f, k], w_t_all bf16 [groups*k, f],
 * dw_all fp32 [groups*f, k].  x_shared != 0: every problem reads the same x [n, k] (first layer).
 * The epilogue operands of cb_epilogue are the packed / stacked ones (bias [groups*f], side_w [groups*f, side_dim],
 * residual / outputs [n, groups*f]; `side` [n, side_dim] is shared).  tcgen05 only (k, f multiples of 64). */
int cb_grouped_linear_fwd(int32_t n, int32_t groups, int32_t k, int32_t f, int32_t x_shared, const void* x,
                          const void* w_all, int32_t act, const cb_epilogue* e, void* stream);
/* dx [n, groups*k] block g = (dy[:, g*f:(g+1)*f] W_g + dx_add) .* act'(x_saved) */
int cb_grouped_linear_dgrad(int32_t n, int32_t groups, int32_t k, int32_t f, const void* dy, const void* w_t_all,
                            const void* x_saved, int32_t x_act, const void* dx_add, void* dx, void* stream);
int cb_grouped_linear_wgrad(int32_t n, int32_t groups, int32_t k, int32_t f, int32_t x_shared, const void* x,
                            const void* dy, float* dw_all, float* dbias_all, float beta, void* stream);
/* Explicit im2col for convolution shapes without an implicit-GEMM specialisation (UniverseHead's 3x3 stride-2 pad-1
 * layers, encoders.py:7-35; the policy's Conv2dHeadModel, models/conv2d.py:67-118): their MACs then run on the tensor
 * cores through cb_conv_fwd / cb_conv_dgrad / cb_conv_wgrad on the [rows, k_pad] patch matrix.
 * col (bf16, row stride k_pad, k = (kh*k_w + kw)*in_c + ci, columns >= K zero) holds the out_h*out_w patch rows of
 * samples [n0, n1); x is addressed through the descriptor's strides (CB_U8 or CB_BF16). */
int cb_im2col(const cb_conv_desc* d, const void* x, void* col, int32_t k_pad, int32_t n0, int32_t n1, void* stream);
/* inverse gather for the data gradient: dx[n,ih,iw,ci] = (sum over the taps reading that pixel of dcol + dx_add)
 * .* act'(x_saved) for samples [n0, n1); dcol bf16 [(n1-n0)*out_h*out_w, k_pad]; dx / x_saved / dx_add bf16 with element
 * strides dx_stride_{n,h,w} (channel stride 1; 16-byte vector path when in_c % 8 == 0). */
int cb_col2im(const cb_conv_desc* d, const void* dcol, int32_t k_pad, const void* x_saved, int32_t x_act,
              const void* dx_add, void* dx, int64_t dx_stride_n, int64_t dx_stride_h, int64_t dx_stride_w,
              int32_t n0, int32_t n1, void* stream);
/* bias gradient: out[c] = beta*out[c] + sum_rows dy[row, c]  (dy bf16 [rows, c]) */
int cb_colsum_bf16(const void* dy, int64_t rows, int32_t c, float* out, float beta, void* stream);
/* side-input weight gradient: dsw[out_c,side_dim] = beta*dsw + dy^T side */
int cb_side_wgrad(const void* dy, const float* side, float* dsw, int64_t m, int32_t out_c,
                  int32_t side_dim, float beta, void* stream);

/* ------------------------------------------------------------------------------------
 * Per-sample losses / bonuses (row-wise reductions).
 * ---------------------------------------------------------------------------------- */
/* icm.py:133, rnd.py:183: out[m] = scale * sum_f (pred[m,f]-target[m,f])^2  (fp32 in) */
int cb_rowwise_sqerr(const float* pred, const float* target, float scale, float* out,
                     int64_t m, int32_t f, void* stream);
/* d/dpred of sum_m coef[m]*out[m]: dpred[m,f] (bf16) = 2*scale*coef[m]*(pred-target);
 * optional elementwise keep-mask/(1-p) (disagreement.py:155) */
int cb_sqerr_grad(const float* pred, const float* target, const float* coef, float scale,
                  const float* elem_mask, float elem_scale, void* dpred_bf16,
                  int64_t m, int32_t f, void* stream);
/* disagreement.py:155-164: out[m] = scale*sum_f mask[m,f]*elem_scale*(pred-target)^2 */
int cb_rowwise_sqerr_masked(const float* pred, const float* target, const float* elem_mask,
                            float elem_scale, float scale, float* out, int64_t m, int32_t f,
                            void* stream);
/* disagreement.py:136-140: out[m] = beta * mean_f var_e(pred[e][m,f]) (unbiased, two-pass);
 * preds: E pointers to fp32 [m,f] packed in a device array of pointers */
int cb_ensemble_variance(const float* const* preds, int32_t e, float beta, float* out,
                         int64_t m, int32_t f, void* stream);
/* The same three reductions for ensemble predictions packed as pred [m, e*f] (the grouped layout):
 * out [e, m] per-model masked squared errors (masks fp32 [e, m, f] or NULL); their gradient dpred bf16 [m, e*f];
 * and the unbiased ensemble variance. */
int cb_ensemble_sqerr(const float* pred, const float* target, const float* masks, float elem_scale, float scale,
                      float* out, int64_t m, int32_t f, int32_t e, void* stream);
int cb_ensemble_sqerr_grad(const float* pred, const float* target, const float* coef, float scale,
                           const float* masks, float elem_scale, void* dpred_bf16, int64_t m, int32_t f, int32_t e,
                           void* stream);
int cb_ensemble_variance_packed(const float* pred, int32_t e, float beta, float* out, int64_t m, int32_t f,
                                void* stream);
/* icm.py:141: per-row cross entropy of logits[m,a] against argmax(onehot[m,:]);
 * dlogits (fp32, may be NULL) = coef[m]*(softmax-onehot_of_argmax) */
int cb_cross_entropy(const float* logits, const float* onehot, const float* coef, float* loss,
                     float* dlogits, int64_t m, int32_t a, void* stream);
/* ndigo.py:197-206, 265-268: binary_cross_entropy(sigmoid(logits), target) per element as torch
 * evaluates it (fp32 sigmoid, logs clamped at -100); row_loss[m] = scale * sum_f bce (NULL to skip),
 * dlogits[m,f] = gscale * (sigmoid - target) (NULL to skip). */
int cb_bce_logits(const float* logits, const float* target, float scale, float* row_loss, float gscale,
                  float* dlogits, int64_t m, int32_t f, void* stream);
/* All of NDIGO's prediction horizons in one pass (ndigo.py:154-163, 194-206, 234-268): logits fp32 [T*B, groups*fp]
 * holds predictor k = g+1 in columns [g*fp, g*fp+d) (the grouped-GEMM layout); row (t,b) of predictor k is scored
 * against obs[t+off_g, b] (obs fp32 [T*B, d]) with off_g = k (the loss, ndigo.py:265-268) unless target_offset_host
 * (HOST array of `groups` ints, each in [0, k]) says otherwise -- the bonus scores predictor H+1 against obs[t+H]
 * (ndigo.py:199-206); rows with t+k >= T belong to no pair (loss 0, gradient 0).  groups <= 16.
 * row_loss [groups, T*B] = scale * sum_f bce with scale 1/d, or 1/((T-k)*B*d) when mean_over_pairs (the sum over rows
 * is then torch's mean over the pairs); dlogits_bf16 [T*B, groups*fp] = upstream/((T-k)*B*d) * (sigmoid - target),
 * zero in the padding columns.  Either output may be NULL. */
int cb_bce_horizons(const float* logits, const float* obs, int32_t T, int32_t B, int32_t d, int32_t groups, int32_t fp,
                    int32_t mean_over_pairs, const int32_t* target_offset_host, float* row_loss, float upstream,
                    void* dlogits_bf16, void* stream);
/* ndigo.py:154-163, 236-240 (a Python double loop in the reference): sliding windows of the one-hot actions,
 * win[(t,b), j*A + a] = onehot[t+j, b, a] for j < W, t+j < T, else 0.  win_bf16: 64 columns per row (zero padded) with
 * row stride ld_bf16 elements -- e.g. columns [G, G+64) of the predictors' packed [belief | window] input matrix;
 * win_f32 [T*B, W*A] or NULL.  W*A <= 64. */
int cb_action_windows(const float* onehot, int32_t T, int32_t B, int32_t A, int32_t W, void* win_bf16, int32_t ld_bf16,
                      float* win_f32, void* stream);
/* utils/tensor.py:39-46: out[0] = sum(x*w)/sum(w) (w NULL -> mean); deterministic order.
 * Also writes out[1] = sum(w). */
int cb_weighted_mean(const float* x, const float* w, float* out, int64_t n, void* stream);
/* out[0] = sum_i x[i]*w[i] (f64 accumulation, fixed order) */
int cb_dot(const float* x, const float* w, float* out, int64_t n, void* stream);
/* coef[m] = upstream * mask[m]*valid[m]/D with D = *denom when given (the all-reduced
 * sum(valid) of a sharded batch) else the local sum(valid); mask may be NULL */
int cb_loss_coef(const float* valid, const float* mask, float upstream, const float* denom,
                 float* coef, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------
 * Policy network + PPO loss.   models/pg/atari_lstm_model.py:114-154, algos/pg/ppo.py:192-225,
 * distributions/categorical.py:33-46.  The LSTM's two GEMMs per timestep (x W_ih^T for all T at once,
 * h W_hh^T per step) run through cb_conv_fwd / cb_conv_dgrad / cb_conv_wgrad as Linear layers.
 * ---------------------------------------------------------------------------------- */
/* torch.nn.LSTM cell (gate order i,f,g,o): a = gx + gh (fp32 [B,4H] each, biases already added by the GEMM
 * epilogues); gates <- (sigmoid(a_i), sigmoid(a_f), tanh(a_g), sigmoid(a_o)) (fp32 [B,4H], may alias gx);
 * c_out = f*c_prev + i*g (fp32 [B,H]); h = o*tanh(c_out) as bf16 [B,H] (next GEMM operand) and, when h_f32 is
 * given, fp32.  H % 4 == 0, rows 16-byte aligned. */
int cb_lstm_cell_fwd(const float* gx, const float* gh, const float* c_prev, float* gates, float* c_out,
                     void* h_bf16, float* h_f32, int32_t B, int32_t H, void* stream);
/* backward of the cell: dh bf16 [B,H] (all gradient arriving at h_t), dc_next fp32 [B,H] or NULL;
 * dgates bf16 [B,4H] = gradient w.r.t. the pre-activations a; dc_prev fp32 [B,H]. */
int cb_lstm_cell_bwd(const void* dh_bf16, const float* dc_next, const float* gates, const float* c_prev,
                     const float* c, void* dgates_bf16, float* dc_prev, int32_t B, int32_t H, void* stream);
/* The whole recurrence of atari_lstm_model.py:142 as ONE call (the host loop over T runs inside the library: two
 * launches per step without a Python round trip).  gates fp32 [T,B,4H]: x W_ih^T + b_ih on entry, the activated
 * gates on exit; w_hh bf16 [4H,H]; h_all bf16 / c_all fp32 [T+1,B,H] with row 0 = the initial state; gh_ws fp32
 * [B,4H] scratch; hn fp32 [B,H] or NULL receives h_T.  `engine` as in cb_conv_desc. */
int cb_lstm_sequence_fwd(int32_t T, int32_t B, int32_t H, float* gates, const void* w_hh_bf16, const float* b_hh,
                         void* h_all_bf16, float* c_all, float* gh_ws, float* hn, int32_t engine, void* stream);
/* Back-propagation through time: dh_out bf16 [T,B,H] = gradient arriving at every h_t from the layers above;
 * w_hh_t bf16 [H,4H]; dgates bf16 [T,B,4H] out; dc_ws fp32 [2,B,H] and dh_ws bf16 [2,B,H] scratch. */
int cb_lstm_sequence_bwd(int32_t T, int32_t B, int32_t H, const float* gates, const float* c_all,
                         const void* dh_out_bf16, const void* w_hh_t_bf16, void* dgates_bf16, float* dc_ws,
                         void* dh_ws_bf16, int32_t engine, void* stream);
/* torch.nn.GRU (ndigo.py:79,127; gate order r,z,n).  Cell: gi = x W_ih^T + b_ih, gh = h W_hh^T + b_hh (fp32 [B,3G]);
 * gates (may alias gi) <- (r, z, n); ghn <- gh_n; h_out = (1-z) n + z h_prev (fp32 + bf16 copies). */
int cb_gru_cell_fwd(const float* gi, const float* gh, const float* h_prev, float* gates, float* ghn, float* h_out,
                    void* h_bf16, int32_t B, int32_t G, void* stream);
/* dh = dout (fp32, may be NULL) + dh_rec (bf16, may be NULL); dgi / dgh bf16 [B,3G]; dh_carry bf16 [B,G] = dh*z */
int cb_gru_cell_bwd(const float* dout, const void* dh_rec_bf16, const float* gates, const float* ghn,
                    const float* h_prev, void* dgi_bf16, void* dgh_bf16, void* dh_carry_bf16, int32_t B, int32_t G,
                    void* stream);
/* The recurrence over T in one call.  gates fp32 [T,B,3G] (gi on entry, (r,z,n) on exit); w_hh bf16 [3G,G];
 * h_all fp32 and h_all_bf16 [T+1,B,G] with row 0 = the initial state (zeros for NDIGO); ghn_all fp32 [T,B,G];
 * gh_ws fp32 [B,3G] scratch. */
int cb_gru_sequence_fwd(int32_t T, int32_t B, int32_t G, float* gates, const void* w_hh_bf16, const float* b_hh,
                        float* h_all, void* h_all_bf16, float* ghn_all, float* gh_ws, int32_t engine, void* stream);
/* BPTT: dout fp32 [T,B,G] = gradient arriving at every h_t; w_hh_t bf16 [G,3G]; dgi / dgh bf16 [T,B,3G] out;
 * dh_ws bf16 [3,B,G] scratch. */
int cb_gru_sequence_bwd(int32_t T, int32_t B, int32_t G, const float* gates, const float* ghn_all, const float* h_all,
                        const float* dout, const void* w_hh_t_bf16, void* dgi_bf16, void* dgh_bf16, void* dh_ws_bf16,
                        int32_t engine, void* stream);
/* The same recurrence as ONE persistent launch (csrc/recurrent.cu): the batch is partitioned over CTAs by rows (the
 * recurrence is independent across environment columns), W_hh stays in registers as tensor-core fragments, the hidden
 * state in registers / shared memory; no per-step launches, no fp32 gh round trip through HBM.  BPTT likewise, with
 * dL/dh carried in fp32.  Same operands as cb_gru_sequence_fwd / _bwd minus the scratch buffers.
 * cb_gru_persistent_supported(G): 1 for the specialised hidden size (128, ndigo.py:47). */
int cb_gru_persistent_supported(int32_t G);
int cb_gru_persistent_fwd(int32_t T, int32_t B, int32_t G, float* gates, const void* w_hh_bf16, const float* b_hh,
                          float* h_all, void* h_all_bf16, float* ghn_all, void* stream);
int cb_gru_persistent_bwd(int32_t T, int32_t B, int32_t G, const float* gates, const float* ghn_all, const float* h_all,
                          const float* dout, const void* w_hh_t_bf16, void* dgi_bf16, void* dgh_bf16, void* stream);
/* atari_lstm_model.py:144: F.softmax over the last dimension, fp32 [n,a]; and its backward */
int cb_softmax_fwd(const float* logits, float* prob, int64_t n, int32_t a, void* stream);
int cb_softmax_bwd(const float* prob, const float* dprob, float* dlogits, int64_t n, int32_t a, void* stream);
/* ppo.py:209-223 per sample m (a <= 64): p = softmax(logits[m]); ratio = (p[act]+1e-8)/(old_prob[m,act]+1e-8);
 * terms[0*n+m] = min(ratio*adv, clamp(ratio, 1-clip, 1+clip)*adv); terms[1*n+m] = 0.5*(value-return)^2;
 * terms[2*n+m] = -sum p log(p+1e-8); terms[3*n+m] = exp(entropy).  prob_out [n,a] optional.
 * When dlogits is given: gradients of  -sum_m coef*surr + value_coeff*sum_m coef*verr - entropy_coeff*sum_m coef*ent
 * w.r.t. logits (fp32 [n,a]) and value (fp32 [n]); coef[m] = valid[m]/sum(valid) (cb_loss_coef).
 * torch.min's tie rule and clamp's inclusive bounds are reproduced. */
int cb_ppo_loss(const float* logits, const float* value, const int64_t* action, const float* old_prob,
                const float* advantage, const float* return_, const float* coef, float ratio_clip,
                float value_coeff, float entropy_coeff, float* prob_out, float* terms, float* dlogits,
                float* dvalue, int64_t n, int32_t a, void* stream);
/* atari_lstm_model.py:144-145 fused: logits = h W_pi^T + b_pi (fp32 [n,A]), value = h w_v + b_v (fp32 [n]),
 * prob = softmax(logits) (optional) in ONE pass over h (bf16 [n,H], 16-byte aligned rows); weights are the fp32
 * masters (rounded to bf16 on the fly like the GEMM engines do), fp32 accumulation.  HBM-bound: n*H*2 bytes read.
 * The value head is optional (w_v = b_v = value = NULL): the same kernel then serves any narrow Linear(H -> A)
 * over bf16 rows, e.g. the last layer of ICM's inverse model (icm.py:88-92).
 * cb_policy_heads_supported(H, outputs): outputs (= A + 1, or A without a value head) <= 8 and
 * H in {256, 512, 1024, 2048}; other shapes use cb_conv_fwd. */
int cb_policy_heads_supported(int32_t H, int32_t outputs);
int cb_policy_heads_fwd(const void* h_bf16, const float* w_pi, const float* b_pi, const float* w_v,
                        const float* b_v, float* logits, float* value, float* prob, int64_t n, int32_t H,
                        int32_t A, void* stream);
/* backward of both heads in one pass: dh (bf16 [n,H]) = (dlogits W_pi + dvalue w_v) .* act'(h) with h_act =
 * CB_ACT_NONE (LSTM output) or CB_ACT_RELU (h is a ReLU output, icm.py:90);  dw_pi [A,H], db_pi [A], dw_v [H],
 * db_v [1] (fp32, overwritten) = the weight / bias gradients.  n*H*2 bytes read + n*H*2 written.
 * The value-head operands may be NULL together (see above). */
int cb_policy_heads_bwd(const void* h_bf16, const float* dlogits, const float* dvalue, const float* w_pi,
                        const float* w_v, void* dh_bf16, float* dw_pi, float* db_pi, float* dw_v, float* db_v,
                        int64_t n, int32_t H, int32_t A, int32_t h_act, void* stream);
/* out (bf16) = a + b (fp32, b may be NULL): gradient hand-over between fp32 loss heads and bf16 GEMM operands */
int cb_add_to_bf16(const float* a, const float* b, void* out_bf16, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------
 * Optimiser.   ppo.py:147-150 (clip_grad_norm_ over all params, then Adam.step)
 * ---------------------------------------------------------------------------------- */
/* out[0] += sum g^2 over n elements (caller zeroes out; call once per flat buffer) */
int cb_sumsq(const float* g, int64_t n, double* out, void* stream);
/* scale = min(1, max_norm/(sqrt(*sumsq)+1e-6)); Adam (torch defaults: no amsgrad, L2 wd=0)
 * p,m,v,g fp32 flat; step is the 1-based step count for bias correction. */
int cb_adam_clip_step(float* p, float* m, float* v, const float* g, int64_t n,
                      const double* sumsq, float max_norm, float lr, float beta1, float beta2,
                      float eps, int32_t step, void* stream);

/* The same update with the learning rate (fp32 [1]) and the step counter (int32 [1], incremented by this call before
 * use) in DEVICE memory: no step-dependent launch parameter, so a CUDA graph of a whole training step can be replayed
 * while the linear learning-rate schedule (ppo.py:66-70) and the bias correction move on. */
int cb_adam_clip_step_dev(float* p, float* m, float* v, const float* g, int64_t n, const double* sumsq, float max_norm,
                          float beta1, float beta2, float eps, const float* lr_dev, int32_t* step_dev, void* stream);

/* ------------------------------------------------------------------------------------
 * Layout helpers for prepared weights.
 * ---------------------------------------------------------------------------------- */
/* torch conv weight fp32 [o][i][kh][kw] -> bf16 [o][kh][kw][i] and bf16 [i][kh][kw][o] */
int cb_prepare_conv_weight(const float* w, void* w_fwd, void* w_t, int32_t o, int32_t i,
                           int32_t kh, int32_t kw, void* stream);
/* inverse for gradients: fp32 [o][kh][kw][i] -> fp32 [o][i][kh][kw] (dst = beta*dst + src) */
int cb_unprepare_conv_grad(const float* g, float* dst, int32_t o, int32_t i, int32_t kh,
                           int32_t kw, float beta, void* stream);

/* ------------------------------------------------------------------------------------
 * Staging.  agents/pg/categorical.py:89-99 (buffer_to(device)) for the uint8 frame stacks.
 * ---------------------------------------------------------------------------------- */
/* rows x width_bytes from (pinned) host memory with pitch src_pitch to device memory with
 * pitch dst_pitch, asynchronously on `stream`. */
int cb_h2d_2d(void* dst, int64_t dst_pitch, const void* src_host, int64_t src_pitch,
              int64_t width_bytes, int64_t rows, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CURIOSITY_B200_H */

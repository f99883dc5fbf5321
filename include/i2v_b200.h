/* i2v_b200 — C ABI of the B200-native I2V / ENS-I2V / AENS-I2V per-step attack loop.
 *
 * The reference (zhipeng-wei/Image-to-Video-I2V-attack) is pure Python over PyTorch; it has no FFI of
 * its own.  Each entry point below therefore cites the reference *Python* lines whose arithmetic it
 * replaces (paths relative to the reference root).  The host side (Python, `i2v_b200/capi.py`) binds
 * these through ctypes, passing `tensor.data_ptr()` and `torch.cuda.current_stream().cuda_stream`.
 *
 * Conventions
 *  - All pointers are DEVICE pointers owned by the caller (torch allocations) unless stated otherwise.
 *  - Kernels are asynchronous on `stream` (a cudaStream_t passed as void*). No hidden allocation.
 *  - Return value: 0 on success, negative I2V_E* code on failure; i2v_last_error() gives the text.
 *  - sm_100a only.  There is no CPU fallback and no other-arch fallback: on a different device the
 *    launch fails and the error is reported.
 *  - "channel layout": element i of a tensor belongs to channel ((i / inner) % channels).
 *        [N,3,H,W] planar  -> inner = H*W,   channels = 3   (image_attacks.py:59  [3,1,1] broadcast)
 *        [B,3,T,H,W]       -> inner = T*H*W, channels = 3   (base_attacks.py:154 [3,1,1,1] broadcast)
 *        [N,H,W,4] NHWC4   -> inner = 1,     channels = 4   (native engine; channel 3 is padding and is
 *                                                            kept at exactly 0 by every kernel)
 *    mean = {0.485,0.456,0.406}, std = {0.229,0.224,0.225} rounded to f32 (image_attacks.py:33-34).
 */
#ifndef I2V_B200_H_
#define I2V_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I2V_VERSION 100

#define I2V_OK            0
#define I2V_EINVAL       -1   /* bad argument (null pointer, bad layout, unsupported size)          */
#define I2V_ECUDA        -2   /* CUDA runtime error (launch failure, wrong architecture, ...)      */
#define I2V_EUNSUPPORTED -3   /* shape not supported by the tensor-core path (caller must not hide) */

typedef void* i2v_stream_t;

int         i2v_version(void);
const char* i2v_last_error(void);
/* Verifies that `device` is compute capability 10.x and makes it current. */
int         i2v_device_check(int device);

/* ---------------------------------------------------------------------------------------------
 * K3 family — fused per-pixel update kernels (HBM-bound, float4 streaming)
 * ------------------------------------------------------------------------------------------- */

/* x = inp * std_c + mean_c          (mul then add, two roundings, no FMA)
 * replaces `_transform_video(video, 'back')`: image_attacks.py:60-62, base_attacks.py:155-157.   */
int i2v_denorm_f32(const float* inp, float* x, int64_t n, int64_t inner, int channels, i2v_stream_t stream);

/* out = (x - mean_c) / std_c          (sub then true division, two roundings)
 * replaces `_transform_video(video, 'forward')`: image_attacks.py:57-59, base_attacks.py:152-154. */
int i2v_normalize_f32(const float* x, float* out, int64_t n, int64_t inner, int channels, i2v_stream_t stream);

/* out = (clamp(x + clamp(mod,-eps,eps), 0, 1) - mean_c) / std_c
 * replaces image_attacks.py:331-332 / 360-361 (also 465-466, TPAMI_attack.py:268-269, 316-317).  */
int i2v_compose_norm_f32(const float* x, const float* mod, float* out, int64_t n, int64_t inner,
                         int channels, float eps, i2v_stream_t stream);

int i2v_fill_f32(float* p, float value, int64_t n, i2v_stream_t stream);

/* K3a: one Adam step on the unconstrained modifier, fused with the backward of the compose/normalise
 * block in front of it and with the compose/normalise of the NEXT step behind it.
 *   g        : dcost/d(true_image), normalised space                     (read)
 *   m, v     : Adam exp_avg / exp_avg_sq                                  (read+write)
 *   mod      : modifier                                                   (read+write)
 *   x        : clean frames in [0,1]                                      (read)
 *   next_img : (clamp(x + clamp(mod',±eps),0,1) - mean_c)/std_c           (write)
 * Arithmetic (every op one f32 rounding; `fma` = fused):
 *   gm   = (g / std_c) * 1[0 <= x+clamp(mod) <= 1] * 1[-eps <= mod <= eps]   autograd of 331-332
 *   m    = fma(f32(1-beta1), gm - m, m)                                      torch.optim.Adam lerp_
 *   v    = fma(f32(1-beta2), gm*gm, v*f32(beta2))                            mul_ + addcmul_      [CUDA arithmetic]
 *   den  = sqrt(v) / f32(sqrt(1-beta2^step)) + f32(adam_eps)
 *   mod  = fma(f32(-lr/(1-beta1^step)), m / den, mod)                        addcdiv_             [CUDA arithmetic]
 * replaces image_attacks.py:351-353 (+331-332 of the next iteration); `step` is 1-based.
 * torch's CPU kernels group the last two differently — v = fma(f32(1-beta2)*gm, gm, v*beta2), mod = mod + (ss*m)/den —
 * and K3a reproduces either bit for bit: i2v_set_adam_arithmetic(1) = torch's CUDA foreach Adam (default: the reference
 * hard-codes .cuda(), image_attacks.py:304), 0 = torch's CPU Adam (what the committed CPU fixtures were made with).   */
int i2v_set_adam_arithmetic(int cuda_arith);
int i2v_get_adam_arithmetic(void);
int i2v_adam_compose_f32(const float* g, float* m, float* v, float* mod, const float* x,
                         float* next_img, int64_t n, int64_t inner, int channels, float eps,
                         double lr, double beta1, double beta2, double adam_eps, int step,
                         i2v_stream_t stream);

/* CUDA-graph friendly form of K3a: the two step-dependent scalars come from a device table
 * step_table[2*k+0] = f32(sqrt(1-beta2^(k+1))), step_table[2*k+1] = f32(-lr/(1-beta1^(k+1)))
 * indexed by the device counter *step_idx (0-based; advanced by i2v_step_advance).
 * i2v_adam_step_table fills a HOST table with exactly the scalars i2v_adam_compose_f32 would use. */
int i2v_adam_step_table(float* host_table, int steps, double lr, double beta1, double beta2);
int i2v_adam_compose_table_f32(const float* g, float* m, float* v, float* mod, const float* x,
                               float* next_img, int64_t n, int64_t inner, int channels, float eps,
                               float w1, float beta2, float a2, float adam_eps,
                               const float* step_table, const int* step_idx, i2v_stream_t stream);
int i2v_step_advance(int* step_idx, i2v_stream_t stream);

/* K3b: BIM update block.  adv is normalised on entry and exit.
 *   a = adv*std_c; a = a + mean_c; a = a + step_size*sign(g); d = clamp(a - x, ±eps);
 *   a = clamp(x + d, 0, 1); adv = (a - mean_c)/std_c        — every op one f32 rounding, sign(0)=0.
 * replaces base_attacks.py:289-293 (same block at 334-338, 405-409, 473-477, 546-550, 605-609,
 * 677-681, 806-810, video_attacks.py:224-228).  `project`=0 gives FGSM's block (254-257: no
 * eps-projection, x unused and may be NULL).                                                       */
int i2v_sign_step_project_f32(float* adv, const float* g, const float* x, int64_t n, int64_t inner,
                              int channels, float step_size, float eps, int project,
                              i2v_stream_t stream);

/* K3c: MI-FGSM.  g is [B,C,T,HW] (C = 3).
 *   i2v_frame_absmean_f32: norm[b,t] = mean_{c,h,w} |g|            utils.py:63 (frame_level=True)
 *                          or, clip_level!=0, norm[b] = mean_{c,t,h,w}|g|   utils.py:65
 *   i2v_mi_sign_step_project_f32: gn = g / norm; gn = gn + momentum*decay; momentum = gn;
 *                          then the K3b block with sign(gn)          base_attacks.py:328-338       */
int i2v_frame_absmean_f32(const float* g, float* norm, int B, int C, int T, int64_t HW,
                          int clip_level, i2v_stream_t stream);
int i2v_mi_sign_step_project_f32(float* adv, const float* g, float* momentum, const float* norm,
                                 const float* x, int B, int C, int T, int64_t HW, int clip_level,
                                 float decay, float step_size, float eps, i2v_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K1 — per-frame cosine feature loss and its analytic gradient (HBM-bound, two passes, the second
 *      served from L2; one 8-CTA cluster per frame with a DSMEM reduction)
 * ------------------------------------------------------------------------------------------- */

/*   cos[n]    = sum_i (a/max(|a|,1e-8))*(b/max(|b|,1e-8))             F.cosine_similarity, dim=1
 *   grad_a[n] = w * ( b/(|a||b|) - cos*a/|a|^2 )  [* 1[a>0] if relu_mask]   autograd of the above
 * a, b, grad_a: [N, D] contiguous (any permutation of the D axis gives the same cos — the reduction
 * is over the whole frame — so NCHW and NHWC feature maps are both accepted).
 * w = *w_dev if w_dev != NULL else w_host (ENS-I2V: 1; AENS: coeffs[l]/L, TPAMI_attack.py:289-291).
 * Sums are accumulated in FP64; the gradient is formed in FP64 and rounded once.
 * grad_a may be NULL (loss only: the init-feature pass never needs it).
 * replaces image_attacks.py:341-347 + autograd (also 475-480; TPAMI_attack.py:282-286).           */
int i2v_cosine_loss_grad_f32(const float* a, const float* b, float* grad_a, float* cos_out,
                             int64_t N, int64_t D, const float* w_dev, float w_host, int relu_mask,
                             i2v_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K2 — adaptive layer re-weighting (latency-bound, one warp)
 * ------------------------------------------------------------------------------------------- */

/* coeffs <- softmax(softmax(prev) + momentum*coeffs)      TPAMI_attack.py:265      (L <= 32)
 * w_out[l] = coeffs[l] / L  (the upstream dcost/dcos of `mean_l(coeffs[l]*sum_n cos[l,n])`, 289-291)
 * weights_log (nullable): row *step_idx of a [steps, L] device log — `self.weights`, 266.         */
int i2v_layer_reweight_f32(float* coeffs, const float* prev, int L, float momentum, float* w_out,
                           float* weights_log, const int* step_idx, i2v_stream_t stream);

/* From cos[L,N]:  s[l] = sum_n cos[l,n] (fixed order);
 *   mode 0 (I2V / ENS-I2V): cost = sum_l s[l]                         image_attacks.py:347, 480
 *   mode 1 (AENS):          cost = mean_l(coeffs[l]*s[l]);            TPAMI_attack.py:290-291
 *                           prev[l] = coef_CE ? coeffs[l]*s[l] : s[l] TPAMI_attack.py:293-297
 * cost_log[*step_idx] = cost (device log; `loss_info` / `cost_saved`, one D2H after the loop).     */
int i2v_layer_sums_f32(const float* cos, const float* coeffs, float* prev, float* cost_log,
                       const int* step_idx, int L, int64_t N, int mode, int coef_CE,
                       i2v_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K4 / K5 — truncated image-backbone forward and data-gradient (NHWC f32 activations)
 * ------------------------------------------------------------------------------------------- */

typedef struct i2v_conv_desc {
    int32_t N, H, W, Cin;          /* conv input  x [N,H,W,Cin]                                     */
    int32_t Cout, R, S;            /* filter [Cout,R,S,Cin]                                          */
    int32_t stride, pad;           /* square stride / zero padding                                   */
    int32_t P, Q;                  /* conv output y [N,P,Q,Cout]                                     */
} i2v_conv_desc;

#define I2V_EPI_RELU        1   /* y = max(y, 0) after bias and residual                            */
#define I2V_LAYOUT_X_NCHW  16   /* the conv INPUT-side tensor (x for fwd, dx for dgrad) is [N,C,H,W]:
                                   lets the stem read the [N,3,H,W] image and its data-gradient write
                                   dcost/dimage in the layout the update kernels use                 */

/* CUDA-core (exact FP32 FFMA) gather-GEMM convolution — every shape, the validation reference of the
 * tensor-core path and the odd-shape path (7x7/s2 Cin=3 stem, AlexNet 11x11/s4, strided dgrad).
 *   fwd  : y  = relu?( conv(x, w)*bn_scale + bias [+ residual] )
 *          bmat = [(r,s,ci), co] = w[co,ci,r,s]*bn_scale[co]  (leading dimension = Cout rounded up to 4)
 *          eval-mode BatchNorm folded into bmat/bias: torchvision `bn(conv(x))` with BN in eval,
 *          image_attacks.py:253-256; forward of image_attacks.py:334.
 *   dgrad: dx = ( conv_transpose(dy, w*bn_scale) [+ addend] ) * 1[mask_src > 0]
 *          bmat = [(r,s,co), ci] = w[co,ci,r,s]*bn_scale[co]  (leading dimension = Cin rounded up to 4)
 *          mask_src = the forward activation this gradient flows into (ReLU backward), or NULL.
 *          The data gradient only: the reference's cost.backward() (image_attacks.py:352) also computes
 *          weight gradients that nothing reads.                                       */
int i2v_conv_fwd_simt_f32(const i2v_conv_desc* d, const float* x, const float* bmat, const float* bias,
                          const float* residual, float* y, int flags, i2v_stream_t stream);
int i2v_conv_dgrad_simt_f32(const i2v_conv_desc* d, const float* dy, const float* bmat, const float* addend,
                            const float* mask_src, float* dx, int flags, i2v_stream_t stream);

/* ---- K7: depth-wise "same" stencil, the gradient smoothing of the translation-invariant attacks -------------
 * base_attacks.py:438-449 (TIFGSM: 15x15 Gaussian per frame, kt = 1) and 636-648 (TIFGSM3D: 15x15x15).
 * src/dst = `volumes` = B*C independent [T,H,W] volumes (a contiguous [B,C,T,H,W] gradient), k = [kt,kh,kw]
 * (odd extents, <= 8192 taps), zero padding: out[v,t,h,w] = sum k[a,b,c] * in[v,t+a-kt/2,h+b-kh/2,w+c-kw/2]. */
int i2v_depthwise_stencil_f32(const float* src, float* dst, int64_t volumes, int T, int H, int W, const float* k,
                              int kt, int kh, int kw, i2v_stream_t stream);

/* ---- K8: temporal translation (reference video_attacks.py, TemporalTranslation) ------------------------------
 * i2v_temporal_shift_stack_f32: the D cyclically shifted copies of the clip fed to the model each step
 *   (93-105 `_cycle_move`, 192-200): out[d][bc][(t + moves[d]) mod T][hw] = adv[bc][t][hw]; adv = [B*C, T, HW].
 * i2v_temporal_combine_f32: the gradient augmentation (163-177 with `_conv1d_frame` 80-91) over grads = [D][B*C,T,HW]:
 *   out = f32(1-weight) * sum_d k[d]*G_d[t] + f32(weight) * sum_d k[d]*G_d[(t + moves[d]) mod T]   (d-ordered FMAs).
 * `moves` and `kernel` are HOST arrays of length D <= 32 (they travel as kernel parameters).                  */
int i2v_temporal_shift_stack_f32(const float* adv, float* out, int64_t BC, int T, int64_t HW, const int* moves, int D,
                                 i2v_stream_t stream);
int i2v_temporal_combine_f32(const float* grads, const float* kernel, const int* moves, int D, double weight, float* out,
                             int64_t BC, int T, int64_t HW, i2v_stream_t stream);

/* ---- K9: ILAF intermediate-level loss (reference image_attacks.py:586-612) and K3d, its update block -------
 * Per hooked layer, delta = f - f_ori, n = |delta|_2, d0 = the unit displacement of the example being fine-tuned,
 * init_norm = its length:  loss = -(0.5 n / init_norm + <d0, delta / n>).
 *   i2v_ila_loss_f32  one HBM pass (FP64 block sums, fixed order) + finalize: stats = {cA, cB, loss, n} with
 *                     d loss / d f = cA * delta + cB * d0; cost_log[*step_idx] (+)= loss.  workspace >=
 *                     i2v_ila_workspace_doubles() doubles.
 *   i2v_ila_grad_f32  grad = cA * (f - f_ori) + cB * d0.
 *   i2v_sign_descent_compose_f32 (615-617, 582-585): gm = (g / std_c) * 1[0 <= x+clamp(mod) <= 1] * 1[|mod| <= eps];
 *                     mod -= step_size * sign(gm); next_img = (clamp(x + clamp(mod,+-eps),0,1) - mean_c)/std_c.      */
int i2v_ila_workspace_doubles(void);
int i2v_ila_loss_f32(const float* f, const float* f_ori, const float* d0, int64_t n, float init_norm, double* workspace,
                     float* stats, float* cost_log, const int* step_idx, int add_to_cost, i2v_stream_t stream);
int i2v_ila_grad_f32(const float* f, const float* f_ori, const float* d0, float* grad, int64_t n, const float* stats,
                     i2v_stream_t stream);
int i2v_sign_descent_compose_f32(const float* g, float* mod, const float* x, float* next_img, int64_t n, int64_t inner,
                                 int channels, float eps, float step_size, i2v_stream_t stream);

/* ---- K6: Dispersion-Reduction loss (reference image_attacks.py:129-234, ImageGuidedStd_Adam) ---------------
 * cost = activations.std() over the WHOLE hooked feature map [N,C,h,w], unbiased (image_attacks.py:216-220);
 * d cost / d x_i = (x_i - mean) / ((n - 1) * std).  The map may be fed in slices (frame chunks):
 *   i2v_std_accumulate_f32  acc[0] += sum(a), acc[1] += sum(a*a) over n elements (FP64, fixed order);
 *                           workspace >= i2v_std_workspace_doubles() doubles; the caller zeroes acc per step
 *   i2v_std_finalize_f32    stats = {mean, std, 1/((n_total-1)*std)} (f32) from acc; cost_log[*step_idx] = std
 *                           (add_to_cost: += , the reference sums the per-layer stds, image_attacks.py:220)
 *   i2v_std_grad_f32        grad = (a - mean) * stats[2], zeroed where a <= 0 if relu_mask (pre-activation
 *                           gradient convention of the native engine)                                      */
int i2v_std_workspace_doubles(void);
int i2v_std_accumulate_f32(const float* a, int64_t n, double* workspace, double* acc, i2v_stream_t stream);
int i2v_std_finalize_f32(const double* acc, int64_t n_total, float* stats, float* cost_log, const int* step_idx,
                         int add_to_cost, i2v_stream_t stream);
int i2v_std_grad_f32(const float* a, float* grad, int64_t n, const float* stats, int relu_mask, i2v_stream_t stream);

/* First layer (Cin = 3, Cout = 64: ResNet 7x7/s2, AlexNet 11x11/s4, VGG 3x3/s1, SqueezeNet 3x3/s2).
 *   fwd  : x [N,3,H,W] (the layout the update kernels keep the image in) -> y [N,P,Q,64] NHWC, + bias, ReLU;
 *          w = [(c,r,s), 64] = weight[co,c,r,s]*bn_scale[co]
 *   dgrad: dy [N,P,Q,64] -> dx [N,3,H,W] = dcost/dimage;  w = [(r,s), c, 64]
 * torchvision `conv1/bn1/relu` (resnet) or `features[0:2]`; forward / backward of image_attacks.py:334 / 352. */
int i2v_conv_stem_supported(const i2v_conv_desc* d);
int i2v_conv_stem_fwd_f32(const i2v_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                          int flags, i2v_stream_t stream);
int i2v_conv_stem_dgrad_f32(const i2v_conv_desc* d, const float* dy, const float* w, float* dx, i2v_stream_t stream);

/* Tensor-core path: the same convolution as an implicit GEMM on tcgen05 (kind::tf32, accumulator in TMEM,
 * operands staged by TMA — im2col-mode tensor maps for R > 1 or stride > 1 — behind an mbarrier pipeline).
 *   w_hi, w_lo : weights [Cout, R*S*Cin] K-major (tap-major, channel-minor), BN scale folded.  w_hi = the f32
 *                weights themselves (the tensor core reads the upper 19 bits of an f32 pattern, i.e.
 *                hi = trunc_tf32(w)); w_lo = w - trunc_tf32(w), exact in f32.  w_lo != NULL selects FP32-parity
 *                mode (3xTF32: a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, FP32 accumulate; A_lo is derived on the fly
 *                from the raw activation tile); w_lo == NULL is plain TF32 (pass weights rounded to nearest
 *                TF32 as w_hi).
 *   dgrad != 0 : data gradient of a stride-1 convolution: src = dy, dst = dx, and w_* is the flipped,
 *                transposed filter [Cin, R*S*Cout] (w[co,ci,R-1-r,S-1-s]*bn_scale[co]).
 *   epilogue   : dst = relu?( acc + bias [+ residual] ) * 1[mask_src > 0]   (each part optional)
 * i2v_conv_tc_supported() says whether a shape is implemented (channels % 32 / % 64, stride-1 dgrad);
 * unsupported shapes return I2V_EINVAL — the caller chooses the CUDA-core kernel explicitly.            */
int i2v_conv_tc_supported(const i2v_conv_desc* d, int dgrad);
int i2v_conv_tc_f32(const i2v_conv_desc* d, int dgrad, const float* src, const float* w_hi, const float* w_lo,
                    const float* bias, const float* residual, const float* mask_src, float* dst, int flags,
                    i2v_stream_t stream);
/* Same, with the ReLU-backward mask as BITS instead of an f32 tensor (1/32 of the mask traffic; the backward of
 * torch's ReLU, image_attacks.py:352 through autograd).  mask_bits = [C_dst/32][M] uint32 words, M = rows of dst,
 * bit j of word (w, m) <-> element (m, 32w + j).  dgrad = 0: OUTPUT, the activity bits 1[dst > 0] of this
 * convolution's result (optional);  dgrad = 1: INPUT, dst is zeroed where the bit is clear.  mask_src and
 * mask_bits are mutually exclusive.  With dense output rows and no f32 mask this entry point runs the TMA
 * epilogue (TMEM -> swizzled shared-memory slot -> cp.async.bulk.tensor store; residual by TMA load).       */
int i2v_conv_tc_bits_f32(const i2v_conv_desc* d, int dgrad, const float* src, const float* w_hi, const float* w_lo,
                         const float* bias, const float* residual, const float* mask_src, uint32_t* mask_bits,
                         float* dst, int flags, i2v_stream_t stream);

/* First-layer data gradient on the tensor cores (replaces i2v_conv_stem_dgrad_f32 when Cout % 32 == 0):
 * Z[(n,p,q), (c,r,s)] = sum_co dy * W is one tcgen05 GEMM whose result is stored as planes Z^T[(c,r,s)][m] into
 * z_scratch, then a col2im gather produces dx [N,3,H,W].
 * wz_hi / wz_lo = [NZ, Cout] K-major: rows (c,r,s) of weight[co,c,r,s]*bn_scale[co], zero-padded to
 * NZ = ceil(3*R*S/64)*64, split into TF32 hi / lo (wz_lo = NULL: plain TF32).  N*P*Q % 4 == 0.            */
/* rows of wz_* the first-layer dgrad GEMM expects for `cols` = 3*R*S filter columns (zero-padded to a multiple of 64,
 * or of 128 with $I2V_STEM_NZ=128) */
int i2v_conv_stem_dgrad_tc_rows(int cols);
int i2v_conv_stem_dgrad_tc_f32(const i2v_conv_desc* d, const float* dy, const float* wz_hi, const float* wz_lo,
                               float* z_scratch, float* dx, i2v_stream_t stream);
/* Frames are processed in groups of i2v_conv_stem_dgrad_tc_group(d) (bounds the scratch; $I2V_STEM_GROUP_MB):
 * z_scratch needs ceil(3*R*S/32)*32 * group*P*Q floats.                                                     */
int i2v_conv_stem_dgrad_tc_group(const i2v_conv_desc* d);

/* First-layer forward on the tensor cores (replaces i2v_conv_stem_fwd_f32 when Cout % 64 == 0): an im2col pass
 * writes the patch matrix col[(n,p,q)][Kp], k = (c,r,s), Kp = ceil(3*R*S/32)*32, for a group of
 * i2v_conv_stem_fwd_tc_group(d) frames, and a tcgen05 GEMM with K = Kp applies the filters, bias and
 * ReLU.  wk_hi / wk_lo = [Cout, Kp] K-major (zero-padded), TF32 hi / lo split (wk_lo = NULL: plain TF32);
 * col_scratch holds group * P*Q*Kp floats; y = [N,P,Q,Cout] NHWC.                                           */
int i2v_conv_stem_fwd_tc_group(const i2v_conv_desc* d);
/* EXPERIMENTAL alternative without the patch matrix (engine: $I2V_STEM_DIRECT=1; S in 5..8, stride 2, Cout = 64, 3xTF32
 * only): the image is packed into a zero-padded NHWC4 copy (xp_scratch, i2v_conv_stem_fwd_direct_scratch_floats(d)
 * floats) and every filter row is a 32-float window of it read by a 4-D tiled TMA box of 16 x 8 output pixels.
 * wr_hi / wr_lo = [Cout, R*32] K-major, k = r*32 + s*4 + c, zero where s >= S or c == 3.                           */
int     i2v_conv_stem_fwd_direct_supported(const i2v_conv_desc* d);
int64_t i2v_conv_stem_fwd_direct_scratch_floats(const i2v_conv_desc* d);
int     i2v_conv_stem_fwd_direct_f32(const i2v_conv_desc* d, const float* x, const float* wr_hi, const float* wr_lo,
                                     const float* bias, float* xp_scratch, float* y, int flags, i2v_stream_t stream);
/* First-layer data gradient WITHOUT scratch (7x7 / stride 2 / pad 3, Cout = 64, Q <= 128 — ResNet's and DenseNet's stem;
 * image_attacks.py:352 for that layer): one dy row per tile, contraction over the output channels on the tensor cores, the
 * col2im on chip (warp shuffles + a register window of the 7 image rows a dy row touches), dcost/dimage [N,3,H,W] written
 * once.  wd_hi / wd_lo = [160, 64] K-major: the 147 taps k = (c,r,s) followed by 13 zero rows (hi = w, lo = w -
 * trunc_tf32(w)).                                                                                                 */
int i2v_conv_stem_dgrad_direct_supported(const i2v_conv_desc* d);
int i2v_conv_stem_dgrad_direct_f32(const i2v_conv_desc* d, const float* dy, const float* wd_hi, const float* wd_lo, float* dx,
                                   i2v_stream_t stream);
/* The same with ResNet's / DenseNet's 3x3 / stride-2 / pad-1 max pooling (floor mode) fused in front: dy_pooled and argmax are
 * the [N, P2, Q2, 64] gradient of the POOLED map and the argmax plane i2v_maxpool_fwd_flags_f32 wrote (mark_dead mode, so the
 * stem's ReLU-backward mask is already in it).  The rows of the stem activation's gradient are rebuilt on chip from the one or
 * two pooled rows that cover them, in the window order of i2v_maxpool_bwd_f32: bit-identical to that call followed by
 * i2v_conv_stem_dgrad_direct_f32, without the 4x larger intermediate ever reaching HBM (image_attacks.py:352 for
 * model.maxpool + model.conv1).                                                                                    */
int i2v_conv_stem_dgrad_pool_supported(const i2v_conv_desc* d, int P2, int Q2);
int i2v_conv_stem_dgrad_pool_f32(const i2v_conv_desc* d, int P2, int Q2, const float* dy_pooled, const uint8_t* argmax,
                                 const float* wd_hi, const float* wd_lo, float* dx, i2v_stream_t stream);
/* First-layer forward WITHOUT the patch matrix (7x7 / stride 2 / pad 3, Cout = 64, Q <= 128, W % 4 == 0; image_attacks.py:334
 * for that layer): one output row per tile, the 7 x 3 input rows staged by TMA (zero filled outside the image), the
 * 128 x 160 patch tile assembled on chip, weights resident; bias + ReLU in the epilogue.  wk_hi / wk_lo as for
 * i2v_conv_stem_fwd_tc_f32: [64, 160] K-major, k = (c,r,s) + 13 zero columns.                                      */
int i2v_conv_stem_fwd_rows_supported(const i2v_conv_desc* d);
int i2v_conv_stem_fwd_rows_f32(const i2v_conv_desc* d, const float* x, const float* wk_hi, const float* wk_lo,
                               const float* bias, float* y, int flags, i2v_stream_t stream);
/* The same with the 3x3 / stride-2 / pad-1 max pooling behind it (torchvision resnet.py: conv1 -> bn1 -> relu -> maxpool, run by
 * image_attacks.py:334) fused into the epilogue: pooled [N, P2, Q2, 64] and its argmax plane (window position r*3+s, first
 * maximum in window order; 255 = no element > 0 when flags has I2V_POOL_MARK_DEAD) come out, bit-identical to
 * i2v_conv_stem_fwd_rows_f32 followed by i2v_maxpool_fwd_flags_f32; the 4x larger stem activation never reaches HBM.
 * P and Q even, P2 = P/2, Q2 = Q/2.  flags: I2V_EPI_RELU | I2V_POOL_MARK_DEAD.                                       */
int i2v_conv_stem_fwd_pool_supported(const i2v_conv_desc* d, int P2, int Q2);
int i2v_conv_stem_fwd_pool_f32(const i2v_conv_desc* d, int P2, int Q2, const float* x, const float* wk_hi, const float* wk_lo,
                               const float* bias, float* pooled, uint8_t* argmax, int flags, i2v_stream_t stream);
int i2v_conv_stem_fwd_tc_f32(const i2v_conv_desc* d, const float* x, const float* wk_hi, const float* wk_lo,
                             const float* bias, float* col_scratch, float* y, int flags, i2v_stream_t stream);

/* Two 1x1 convolutions whose outputs are added — a Bottleneck's downsample branch (d: 1x1 / stride s / pad 0 over x) and its
 * last convolution (1x1 / stride 1 over t [N,P,Q,C2]; torchvision resnet.py Bottleneck.forward `out += identity`, run by
 * image_attacks.py:334) — as ONE implicit GEMM over K = Cin followed by C2: the k-steps of the second range read their
 * activation tile from t through a second tensor map.  w_hi / w_lo = [Cout, Cin + C2] K-major (the two BN-folded weight
 * matrices side by side, TF32 hi / lo split; w_lo = NULL: plain TF32), bias = the sum of the two folded biases.  The
 * downsample output (written once and read back as the residual otherwise) never exists.  Forward only; bits_out as the
 * forward mask_bits of i2v_conv_tc_bits_f32.                                                                        */
int i2v_conv_tc_dual_f32(const i2v_conv_desc* d, const float* x, int C2, const float* t, const float* w_hi, const float* w_lo,
                         const float* bias, uint32_t* bits_out, float* dst, int flags, i2v_stream_t stream);
/* Strided data gradient on the tensor cores, one stride-parity class (ph, pw) per call: image rows
 * h = stride*i + ph receive only the taps r = r0 + stride*a, r0 = (ph + pad) mod stride, from dy row
 * i + (ph + pad - r0)/stride - a — a dense stride-1 implicit GEMM over dy whose output rows are scattered
 * with pitch `stride` into dx.  w_* = [Cin, (tap_h, tap_w, co)], tap_h = A_h-1-a (hi/lo split as above).
 * Classes without taps are a no-op (their pixels get no gradient from this conv).                       */
int i2v_conv_tc_dgrad_class_f32(const i2v_conv_desc* d, int ph, int pw, const float* dy, const float* w_hi,
                                const float* w_lo, const float* addend, const float* mask_src, float* dx,
                                i2v_stream_t stream);
/* The same with the ReLU-backward mask as BITS ([Cin/32][N*H*W] words: the bits_out of the forward launch that produced the
 * tensor dx is the gradient of) instead of the f32 activation.  With mask_src == NULL the class runs the TMA epilogue: its
 * rows scatter into dx through an im2col-mode TMA store over the class's strided view of dx (the addend is read the same
 * way) — no per-thread global access.                                                                              */
int i2v_conv_tc_dgrad_class_bits_f32(const i2v_conv_desc* d, int ph, int pw, const float* dy, const float* w_hi,
                                     const float* w_lo, const float* addend, const float* mask_src, const uint32_t* mask_bits,
                                     float* dx, i2v_stream_t stream);

/* k x k max pooling (stride, -inf padding), NHWC, C % 4 == 0.  argmax[N,P,Q,C] = r*k+s of the FIRST
 * maximum in window scan order (torch.nn.MaxPool2d); backward is the gather form (no atomics) and can
 * apply the ReLU-backward mask of the pooled tensor's producer.  torchvision resnet.maxpool (3,2,1),
 * vgg (2,2,0), alexnet / squeezenet (3,2,0; ceil_mode via P,Q).                                      */
int i2v_maxpool_fwd_f32(const float* x, float* y, uint8_t* argmax, int N, int H, int W, int C, int P, int Q,
                        int k, int stride, int pad, i2v_stream_t stream);
/* flags: bit 2 (I2V_POOL_MARK_DEAD): the pooled tensor is a ReLU output — windows whose maximum is not > 0 are marked
 * argmax = 255 (no winner), so the backward pass routes nothing through them and needs no ReLU-backward mask.     */
#define I2V_POOL_MARK_DEAD 4
int i2v_maxpool_fwd_flags_f32(const float* x, float* y, uint8_t* argmax, int N, int H, int W, int C, int P, int Q,
                              int k, int stride, int pad, int flags, i2v_stream_t stream);
/* flags: bit 0 (I2V_POOL_ACCUMULATE) dx += ; bit 1 (I2V_POOL_MASK_POOLED) mask_src is the POOLED output y
 * [N,P,Q,C] instead of the pooled tensor's input x [N,H,W,C]: the winner of a window is y itself, so
 * 1[x[argmax] > 0] = 1[y > 0], and y is stride^2 times smaller than x.                                   */
#define I2V_POOL_ACCUMULATE 1
#define I2V_POOL_MASK_POOLED 2
int i2v_maxpool_bwd_f32(const float* dy, const uint8_t* argmax, const float* mask_src, float* dx, int N, int H,
                        int W, int C, int P, int Q, int k, int stride, int pad, int flags,
                        i2v_stream_t stream);

/* DenseNet (reference image_attacks.py:95-98 constructs it; hook `features.denseblock{d}`).
 * Pre-activation: dst[m, c] = max(0, scale[c]*src[m*src_ld + c] + shift[c]) for c < C, 0 for C <= c < Cp — eval-mode
 * BatchNorm (scale = gamma/sqrt(var+eps), shift = beta - mean*scale) + ReLU of the first C channels of a concat buffer
 * with row pitch src_ld, written densely as [M, Cp] (torchvision densenet._DenseLayer.norm1/relu1, _Transition.norm/relu).
 * All channel counts are multiples of 4.                                                                        */
int i2v_bn_relu_f32(const float* src, int64_t M, int C, int Cp, int src_ld, const float* scale, const float* shift,
                    float* dst, i2v_stream_t stream);
/* 2x2 / stride-2 average pooling of a transition (floor mode), NHWC: forward writes channels [dst_off, dst_off+C) of
 * rows with pitch dst_ld (the next block's concat buffer); backward reads such a slice and writes dx [N,H,W,C] densely. */
int i2v_avgpool2_fwd_f32(const float* x, float* y, int N, int H, int W, int C, int dst_ld, int dst_off, i2v_stream_t stream);
int i2v_avgpool2_bwd_f32(const float* dy, float* dx, int N, int H, int W, int C, int src_ld, int src_off, i2v_stream_t stream);

/* dst[m, dst_off : dst_off+Ccopy] (=|+=) src[m, src_off : src_off+Ccopy] — channel concat of SqueezeNet's
 * Fire modules (torch.cat([expand1x1, expand3x3], 1)) and its backward split.                        */
int i2v_copy_channels_f32(const float* src, float* dst, int64_t M, int Csrc, int src_off, int Cdst, int dst_off,
                          int Ccopy, int accumulate, i2v_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* I2V_B200_H_ */

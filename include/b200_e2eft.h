/* b200_e2eft.h — C ABI of the B200-native single-step denoising engine (libb200_e2eft.so).
 *
 * Drop-in boundary for the hot path named by BASELINE.json `north_star`:
 *   VAE.encode -> UNet2DConditionModel forward (t=999) -> x0 -> VAE.decode.
 * The reference (VisualComputingInstitute/diffusion-e2e-ft) is pure Python and has no FFI of its
 * own; every entry point below replaces the third-party library kernel (cuDNN / cuBLAS / xformers /
 * ATen) behind one leaf operator of the reference's model graph.  The citation after each
 * prototype is the reference call site (relative to /root/reference) the function serves.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are CUDA device pointers owned by the caller
 *     (PyTorch allocates inputs, outputs and workspaces); the library never allocates or syncs;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return 0 = ok; <0 = invalid argument (nothing launched); >0 = cudaError_t of the launch;
 *     b200_last_error_string() describes the last failure on the calling thread;
 *   - activations are NHWC ("channels last") fp16 operands; the residual stream may be fp16 or fp32
 *     (`*_f32` flags); accumulation, normalisation statistics and softmax are always fp32;
 *   - there is no CPU fallback: without an sm_100a device every launch returns an error.
 */
#ifndef B200_E2EFT_H_
#define B200_E2EFT_H_

#ifdef __cplusplus
extern "C" {
#endif

const char* b200_last_error_string(void);
int b200_abi_version(void);

/* epilogue activations */
#define B200_ACT_NONE 0
#define B200_ACT_SILU 1
#define B200_ACT_GEGLU 2 /* W rows pre-interleaved per tile: [value half | gate half] */
#define B200_ACT_GELU 3  /* erf GELU */
#define B200_ACT_EXP2 4  /* exp2(alpha * acc + bias): softmax probabilities recomputed from the log-sum-exp */

/* out[b][m][n] = act( alpha * sum_k A[b][m][k] * W[(b)][n][k] + bias + residual[b][m][n] )
 * tcgen05 GEMM, fp16 operands (K contiguous), fp32 accumulate in TMEM.
 * Replaces: nn.Linear of proj_in/proj_out (GeoWizard/geowizard/models/transformer_2d.py:152-155,
 * 214-217), to_q/to_k/to_v/to_out (attention.py:470-478,501), GEGLU proj + FF out
 * (attention.py:755,765), TimestepEmbedding / time_emb_proj / class_embedding
 * (unet_2d_condition.py:974-1000); batched form = QK^T and PV of the VAE mid attention
 * (unet_2d_blocks.py:589-601). */
int b200_linear(const void* A, long long lda, long long a_batch_stride,
                const void* W, long long ldw, long long w_batch_stride /* 0 = shared */,
                int M, int N, int K, int batch,
                const float* bias, int bias_row,
                const void* residual, long long ld_res, long long res_batch_stride,
                void* out, long long ldo, long long out_batch_stride, int out_f32,
                int act, float alpha,
                double* chan_stats /* optional [M/rows_per_img][N][2]: per-channel sum / sum of squares of
                                      the stored values (caller zeroes): each thread accumulates shifted fp32
                                      partial sums (no cancellation when |mean| >> std) and merges them with
                                      fp64 atomics, so the result is order-independent to ~1e-16 */,
                int rows_per_img,
                void* out2_f16 /* optional fp16 copy of `out` (same strides) for a following GEMM operand */,
                int res_mul /* 1: out = act(alpha*acc + bias) * residual (GEGLU as gate GEMM + value GEMM) */,
                int a_mn /* 1: A is stored [K][M] (row pitch lda >= M): out = A^T-as-stored x W^T without a transposition
                            pass (MN-major UMMA operand).  Weight gradients dW = dY^T X, attention backward dK = dS^T Q */,
                int w_mn /* 1: W is stored [K][N] (row pitch ldw >= N): data gradients dX = dY W, dQ = dS K */,
                long long bias_batch_stride /* bias_row with batch > 1: bias of batch b starts at bias + b * stride */,
                void* stream);

/* GEGLU tile width for packed width N (weights/bias rows are interleaved per tile of this width:
 * [value rows of the tile | gate rows of the tile]); 0 = not tileable. */
int b200_geglu_block_n(int N);

/* Implicit-GEMM convolution on NHWC fp16 input, weights packed [Cout][tap][Cin] (+[C2] shortcut
 * columns), tcgen05 + TMA, no im2col buffer.  Tap t reads input pixel
 * (ho*stride + tap_dy[t], wo*stride + tap_dx[t]); out-of-range pixels read as zero (padding).
 * Output pixel (ho,wo) is written at (ho*out_mul+out_oy, wo*out_mul+out_ox) of an
 * (Ho*out_mul x Wo*out_mul) NHWC (or NCHW when out_nchw) tensor.
 *   out = act( conv(X) + conv1x1(X2) + bias[c] + rowvec[img][c] + residual )
 * Replaces: ResnetBlock2D.conv1/conv2/conv_shortcut (instantiated at
 * GeoWizard/geowizard/models/unet_2d_blocks.py:1064-1076,1211-1223,2242-2254,2400-2412,667-679),
 * Downsample2D.conv (:1107-1113,1228-1234, VAE :1315-1321), Upsample2D.conv (:2285,2417),
 * conv_out (unet_2d_condition.py:617-619) and the VAE encoder/decoder convs
 * (Marigold/marigold/marigold_pipeline.py:493,516). */
int b200_conv2d_nhwc(const void* X, int NB, int H, int W, int Cin,
                     const void* X2, int C2,
                     const void* Wp, int Cout, int num_taps, const int* tap_dy, const int* tap_dx,
                     int stride, int Ho, int Wo, int out_mul, int out_oy, int out_ox,
                     const float* bias, const float* rowvec, long long ld_rowvec,
                     const void* residual, void* out, int out_f32, int out_nchw, int act,
                     double* chan_stats /* optional [NB][Cout][2], see b200_linear */,
                     void* out2_f16 /* optional fp16 NHWC copy of `out` */, void* stream);

/* 3x3 / stride 1 / pad 1 convolution with Cout <= 8 (the `conv_out` layers: unet_2d_condition.py:617-619
 * and the VAE encoder/decoder conv_out): NHWC fp16 in (C % 64 == 0), NCHW fp32 out, input read once.
 * wq: fp16 [C/64][9 taps][4 k-steps][8 n][16 k] (zero padded to 8 output channels). */
int b200_conv3x3_small_cout(const void* x, int NB, int H, int W, int C, const void* wq,
                            const float* bias, int Cout, float* out, void* stream);

/* Patch matrix for the small-Cin input convolutions (conv_in: 8->320, 3->128, 4->512):
 * out[pixel][tap*Cin + c] fp16, row length Kpad (zero padded).  `x` is NCHW (x_f32 ? fp32 : fp16).
 * Serves unet_2d_condition.py:294-296,1084 and the VAE conv_in. */
int b200_im2col3x3_nchw(const void* x, int x_f32, int NB, int C, int H, int W, void* out,
                        int Kpad, void* stream);

/* GroupNorm statistics over NHWC input that is the channel-concatenation of up to two tensors
 * (skip-connection concat, unet_2d_blocks.py:2328,2456, is never materialised).
 * sums[n][g][2] (double) must be zeroed by the caller. */
int b200_group_norm_stats(const void* x1, int C1, const void* x2, int C2, int in_f32, int NB,
                          int HW, int groups, double* sums, void* stream);
/* y = [silu]( (x-mean)*rstd*gamma + beta ) as fp16 NHWC; optional raw fp16 copy of the
 * (concatenated) input for the 1x1 shortcut operand.
 * Replaces GroupNorm+SiLU of ResnetBlock2D norm1/norm2, conv_norm_out
 * (unet_2d_condition.py:605-610,1209-1211) and Transformer2DModel.norm (transformer_2d.py:151,331). */
int b200_group_norm_apply(const void* x1, int C1, const void* x2, int C2, int in_f32, int NB, int HW,
                          int groups, const double* sums, const float* gamma, const float* beta,
                          float eps, int silu, void* y, void* raw_copy, void* stream);

/* Same, but the statistics come from per-channel sums produced by the epilogue of the kernel that wrote
 * each source (chan_stats of b200_linear / b200_conv2d_nhwc): cs1 [NB][C1][2], cs2 [NB][C2][2] (fp64). */
int b200_group_norm_apply_cs(const void* x1, int C1, const double* cs1, const void* x2, int C2,
                             const double* cs2, int in_f32, int NB, int HW, int groups,
                             const float* gamma, const float* beta, float eps, int silu, void* y,
                             void* raw_copy, void* stream);

/* LayerNorm over the last dim of [rows][C] (in_f32 ? fp32 : fp16) -> fp16.
 * Replaces BasicTransformerBlock.norm1/2/3 (attention.py:205,237,264). */
int b200_layer_norm(const void* x, int in_f32, long long rows, int C, const float* gamma,
                    const float* beta, float eps, void* y, void* stream);

/* Flash attention, head_dim 64, fp16 Q/K/V read in place from (possibly fused) projection
 * buffers: element (b, l, h, d) of Q is q[b*q_bs + l*q_ls + h*64 + d] (same for K, V).
 * `kv_segments` = 2 implements GeoWizard's joint self-attention: batch element b attends to the
 * keys/values of b%(B/2) and b%(B/2)+B/2 concatenated (attention.py:482-491).
 * out[b][l][h*64+d] fp16 with row stride o_ls.   softmax(QK^T*scale)V, fp32 softmax.
 * Replaces xformers.ops.memory_efficient_attention (attention.py:497) / attn1, attn2
 * (attention.py:338-343,375-380). */
int b200_attention_d64(const void* q, long long q_bs, long long q_ls,
                       const void* k, long long k_bs, long long k_ls,
                       const void* v, long long v_bs, long long v_ls,
                       void* out, long long o_bs, long long o_ls,
                       int B, int heads, int Lq, int Lk, int kv_segments, float scale,
                       float* lse /* optional [B][heads][Lq] fp32: log2-domain log-sum-exp of the scaled scores, so that
                                     P_ij = exp2(scale * log2(e) * S_ij - lse_i) — the backward pass recomputes P from it */,
                       void* stream);

/* Row softmax: P[r][:] = softmax(scale * S[r][:]) fp32 -> fp16 (VAE mid-block attention, d=512). */
int b200_softmax_rows(const float* S, long long lds, void* P, long long ldp, long long rows,
                      int cols, float scale, void* stream);

/* Grouped softmax for the constant-context cross-attention specialisation (SURVEY.md §8 f1;
 * Marigold/marigold/marigold_pipeline.py:428-432 repeats ONE [1,2,1024] empty-text embedding over the batch,
 * GeoWizard/geowizard/models/attention.py:375-380 attends to it): logits [rows][ld_in] fp32, column
 * head*S + s; softmax over the S keys of every head -> fp16 [rows][ld_out] (columns >= heads*S zeroed). */
int b200_softmax_groups(const float* logits, int ld_in, long long rows, int heads, int S, void* P, int ld_out,
                        void* stream);

/* Nearest-neighbour resize NHWC (in_f32 ? fp32 : fp16) -> fp16 (Upsample2D interpolate,
 * exact 2x or explicit `size=`, unet_2d_condition.py:1185-1186). */
int b200_upsample_nearest_nhwc(const void* x, int in_f32, int NB, int H, int W, int C, int OH,
                               int OW, void* y, void* stream);

/* Sinusoidal timestep embedding (flip_sin_to_cos, freq_shift 0) -> fp16 [B][dim].
 * t is a device array of B floats.  unet_2d_condition.py:974. */
int b200_timestep_embedding(const float* t, int B, int dim, void* out, void* stream);

/* CLIP text embeddings: out[r][c] = tok[ids[r]][c] + pos[r % L][c], fp32 [rows][C]; ids int64 (clamped to the table);
 * tables fp16 or fp32 (w_f32).  transformers==4.37.2 models/clip/modeling_clip.py CLIPTextEmbeddings.forward, reached
 * from Marigold/marigold/marigold_pipeline.py:369 (`self.text_encoder(text_input_ids)[0]`). */
int b200_embed_tokens(const long long* ids, const void* tok, const void* pos, int w_f32, long long rows, int L, int C,
                      int vocab, float* out, void* stream);

/* Small dense per-pixel channel mix on NCHW fp32 (Cin, Cout <= 8):
 * out[n][co][p] = sum_ci Wm[co][ci] * (a1*in1[n][ci][p] + a2*in2[n][ci][p]) + bias[co].
 * Serves quant_conv + latent scaling (marigold_pipeline.py:494-497) and
 * pred_original_sample + /0.18215 + post_quant_conv (:457-465,513-515). */
int b200_pointwise_nchw(const float* in1, float a1, const float* in2, float a2, int in_cstride,
                        const float* Wm, const float* bias, int NB, int Cin, int Cout, long long HW,
                        float* out, void* stream);

/* Decode post-ops on NCHW fp32 [B][3][HW]: mode 0 = depth: (clip(mean_c, -1, 1)+1)/2 -> [B][1][HW];
 * mode 1 = normals: x/(||x||_2+1e-5) * sign -> [B][3][HW] (marigold_pipeline.py:467-478);
 * mode 2 / 3 = the training variants: clip(mean_c) without the affine map / normalised then clamped
 * (training/train.py:532-540). */
int b200_decode_post(const float* x, int NB, long long HW, int mode, float sign, float* out,
                     void* stream);

/* Task losses of the E2E fine-tuning step, forward only (training/util/loss.py:13-67, training/train.py:542-556).
 * pred/target NCHW fp32 ([B][1][HW] depth, [B][3][HW] normals), mask [B][HW] bytes, workspace: zeroed doubles
 * (5*B + 2 for SSI, 2 for angular), out: one float (mean over masked pixels; nan for an empty mask). */
int b200_ssi_loss(const float* pred, const float* target, const unsigned char* mask, int B, long long HW,
                  double* workspace, float* out, void* stream);
int b200_angular_loss(const float* pred, const float* target, const unsigned char* mask, int B, long long HW,
                      double* workspace, float* out, void* stream);

/* Optimizer side of the fine-tuning step on flat fp32 buffers (training/train.py:346-353,564-566):
 * sum of squares (gradient norm; `out` is a zeroed double accumulated with atomics) and a fused
 * clip_grad_norm_ + torch.optim.AdamW update (decoupled weight decay, bias correction by `step` >= 1).
 * grad_norm_sq (device, may be NULL) and max_grad_norm give the clip coefficient without a host sync. */
int b200_sumsq(const float* x, long long n, double* out, void* stream);
int b200_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                    float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    const double* grad_norm_sq, float max_grad_norm, void* stream);
/* Same with a loss-scaled gradient buffer: grad holds S*g and grad_norm_sq = |S*g|^2; grad_unscale = 1/S. */
int b200_adamw_step_scaled(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                           float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                           const double* grad_norm_sq, float max_grad_norm, float grad_unscale, void* stream);

/* Same update driven by a device-side state block (fp32[8]: loss scale, growth tracker, applied steps, skipped
 * steps, skipped flag, gradient multiplier, bc1, bc2) so that neither the skip decision nor dynamic loss scaling
 * needs a host sync: the step is skipped (parameters and moments untouched, as torch.optim.AdamW leaves parameters
 * whose .grad is None) when the gradient norm is non-finite (fp16 overflow of the loss-scaled backward: the scale is
 * halved) or exactly zero (empty validity masks / NaN loss: training/train.py:503,546-551 back-propagates 0).
 * `grad` holds (loss scale x world size) x the mean gradient; inv_world = 1 / data-parallel ranks. */
int b200_adamw_step_state(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                          float lr, float beta1, float beta2, float eps, float weight_decay,
                          const double* grad_norm_sq, float max_grad_norm, float inv_world, float* state,
                          int dynamic_scale, float growth_interval, float min_scale, float max_scale, void* stream);

/* ---- Backward pass (training/train.py:545-566, `accelerator.backward(loss)`).  The GEMM-shaped halves run on
 * b200_linear / b200_conv2d_nhwc with re-packed operands; these are the streaming / reduction kernels around them.
 *
 * b200_gather_planar: out[c][q] (fp16, row stride ldo >= NB*Ho*Wo, zero-filled past the last pixel) =
 *   x[n][(stride*o + oy)/up][(stride*p + ox)/up][c] (pixel stride ldx >= C) with q = (n*Ho + o)*Wo + p, zero outside the (up-sampled)
 *   image: the K-major (K = pixels) operand of a weight-gradient GEMM for one kernel tap of a stride-1/-2 or
 *   nearest-2x-upsampled 3x3 conv; with Ho=H, Wo=W, stride=1, up=1, oy=ox=0 a plain [rows][C] -> [C][rows] transpose.
 * b200_col_sum: out[c] += sum_rows x[row][c] (bias gradients).
 * GroupNorm backward (diffusers GroupNorm(32) + optional SiLU, NHWC): mean_rstd [NB][groups][2] from the
 *   forward's statistics; pass 1 accumulates S [NB][Ctot][2] = (sum dz, sum dz*xhat) per channel (zeroed by
 *   the caller; call once per concatenated input with its channel offset), pass 2 writes
 *   dx = rstd*(dz*gamma - mean_g(gamma dz) - xhat*mean_g(gamma dz xhat)) (+ add).  d_gamma = sum_n S[..1],
 *   d_beta = sum_n S[..0].  dy is fp16 [NB][HW][Ctot].
 * b200_layer_norm_bwd: dx (+ add) and d_gamma/d_beta (accumulated into zeroed fp32 [C]).
 * b200_softmax_bwd_rows: dS = scale * P o (dP - rowsum(dP o P)), P/dS fp16, dP fp32.
 * b200_act_bwd: dx = dy * act'(x), act = B200_ACT_SILU | B200_ACT_GELU (fp16).
 * b200_geglu_bwd: y = h*gelu(g): dh = dy*gelu(g), dg = dy*h*gelu'(g). */
int b200_gather_planar(const void* x, int in_f32, long long ldx, int NB, int H, int W, int C, int Ho, int Wo, int stride, int up,
                       int oy, int ox, void* out, long long ldo, void* stream);
int b200_col_sum(const void* x, int in_f32, long long rows, int C, long long ld, float* out, void* stream);
/* delta[b][h][t] = sum_d a[b,t,h*64+d] * c[b,t,h*64+d] (fp16 in, fp32 out): the row term of the softmax backward,
 * dS = scale * P o (dP - delta), with a = dO and c = O (attention backward, attention.py:497). */
int b200_rowdot_heads(const void* a, long long a_bs, long long a_ls, const void* c, long long c_bs, long long c_ls,
                      int B, int L, int heads, float* out, void* stream);
int b200_group_norm_mean_rstd(const double* sums, const double* cs1, int C1, const double* cs2, int C2, int NB, int HW,
                              int groups, float eps, float* mean_rstd, void* stream);
int b200_group_norm_bwd_sums(const void* x, int in_f32, int Cx, int c_off, int Ctot, const void* dy, int NB, int HW,
                             int groups, const float* mean_rstd, const float* gamma, const float* beta, int silu,
                             float* S, void* stream);
int b200_group_norm_bwd_apply(const void* x, int in_f32, int Cx, int c_off, int Ctot, const void* dy, int NB, int HW,
                              int groups, const float* mean_rstd, const float* gamma, const float* beta, int silu,
                              const float* S, const void* add, void* dx, int out_f32, void* stream);
int b200_layer_norm_bwd(const void* x, int in_f32, long long rows, int C, const float* gamma, const void* dy,
                        float eps, const void* add, void* dx, int out_f32, float* dgamma, float* dbeta, void* stream);
int b200_softmax_bwd_rows(const void* P, long long ldp, const float* dP, long long ldd, void* dS, long long rows,
                          int cols, float scale, void* stream);
int b200_act_bwd(const void* x, const void* dy, long long n, int act, void* dx, void* stream);
int b200_geglu_bwd(const void* h, const void* g, long long ld_hg, const void* dy, long long rows, int inner, void* dh,
                   void* dg, long long ld_d, void* stream);

/* Loss / post-op backward of the fine-tuning step (training/train.py:532-556, training/util/loss.py):
 * d(loss)/d(pred) * grad_out[0] (device scalar: the upstream gradient, i.e. the loss scale).  The SSI gradient
 * flows through the per-image least-squares scale/shift, as torch.autograd does in the reference.
 * workspace: zeroed doubles, 7*B (ssi) / 1 (angular).  decode_post_bwd: modes 2 (depth) / 3 (normals). */
int b200_ssi_loss_bwd(const float* pred, const float* target, const unsigned char* mask, int B, long long HW,
                      double* workspace, const float* grad_out, float* dpred, void* stream);
int b200_angular_loss_bwd(const float* pred, const float* target, const unsigned char* mask, int B, long long HW,
                          double* workspace, const float* grad_out, float* dpred, void* stream);
int b200_decode_post_bwd(const float* x, const float* dout, int NB, long long HW, int mode, float* dx, void* stream);
/* Backward of b200_upsample_nearest_nhwc (explicit output size, unet_2d_condition.py:1185-1186): fp32 NHWC,
 * dx[n,h,w,:] = (add) + sum of dy over the output pixels whose nearest source is (h,w). */
int b200_upsample_nearest_bwd(const float* dy, int NB, int H, int W, int C, int OH, int OW, const float* add,
                              float* dx, void* stream);

/* fp32 <-> fp16 casts / layout helpers used at module boundaries. */
int b200_cast_f32_to_f16(const float* x, void* y, long long n, void* stream);
int b200_nhwc_to_nchw_f32(const void* x, int in_f32, int NB, int C, long long HW, float* y, void* stream);

/* debugging: force a BLOCK_N (0 = automatic) */
void b200_debug_force_block_n(int bn);
/* debugging / perf experiments (results are wrong when set): 1 = skip epilogue stores, 2 = skip A loads,
 * 4 = skip W loads */
void b200_debug_set_flags(int flags);
/* 1 (default) = swap operands automatically when Cout % 128 == 0; 0 = never */
void b200_debug_set_swap(int mode);
void b200_debug_set_halo(int mode);   /* 1 = automatic halo-resident stride-1 3x3 conv (default), 0 = per-tap boxes */
void b200_debug_set_attention_version(int v);      /* 2 = two-pass softmax, O in registers; 3 = S read once, O in TMEM */
int b200_debug_last_path(void);       /* path of the last b200_conv2d_nhwc call: 1 = halo-resident, 0 = per-tap boxes */

/* torchvision resize(x, size, BICUBIC, antialias=True) of [planes][H][W] fp32 (aten _upsample_bicubic2d_aa, Keys
 * cubic a = -0.5): the CLIP image-encoder input of GeoWizard/geowizard/models/geowizard_pipeline.py:239-243.
 * tmp: [planes][H][OW] fp32 scratch. */
int b200_resize_bicubic_aa(const float* x, long long planes, int H, int W, int OH, int OW, float* tmp, float* out,
                           void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Host-pipeline post/pre-processing on the device.
 *
 * b200_ensemble_normals: Marigold/marigold/marigold_pipeline.py:59-71 == GeoWizard/geowizard/utils/
 *   normal_ensemble.py:6-22.  preds [E][3][HW] fp32 (E <= 32) -> out [3][HW] = preds[index] / (|.| + 1e-5) with
 *   index = argmin_e sum_pixels acos(clip(cos(mean-angle normal, n_e), +-0.999)); err_ws = double[E] scratch.
 * b200_ensemble_depths_objective / _reduce: Marigold/marigold/util/ensemble.py:40-132.  imgs [E][HW] fp32, s/t [E]
 *   device fp32.  objective: ws (double[2] scratch) [0] = sum over pixels and pairs i<j of (v_i - v_j)^2,
 *   out3[1], out3[2] = min / max of the reduced map (the scipy-BFGS closure :74-98 finishes on the host);
 *   reduce: aligned / uncertainty [HW] = median + MAD (reduction 0) or mean + std (1), scaled to [0, 1] (:110-130).
 * b200_minmax_rows: per-row (min, max) of [rows][cols] fp32 -> out[rows][2] (ws: uint32[2*rows]) (:66-69 init guess).
 * b200_minmax_normalise: x = (x - min x) / (max x - min x) in place (marigold_pipeline.py:305-312); minmax_out[2] opt.
 * b200_rgb_normalise: uint8 / fp32 [0,255] -> fp32 x / 255 * 2 - 1 (:245-247).
 * b200_resize_bilinear_aa: torchvision resize(..., BILINEAR, antialias=True) of [planes][H][W] fp32 (:237-242,315-321),
 *   separable (width then height), tmp = [planes][H][OW] scratch.   b200_resize_nearest: geowizard normals. */
int b200_ensemble_normals(const float* preds, int E, long long HW, double* err_ws, float* out, int* index, void* stream);
int b200_ensemble_depths_objective(const float* imgs, const float* s, const float* t, int E, long long HW,
                                   int reduction, double* ws, float* out3, void* stream);
int b200_ensemble_depths_reduce(const float* imgs, const float* s, const float* t, int E, long long HW, int reduction,
                                double* ws, float* aligned, float* uncertainty, void* stream);
int b200_minmax_rows(const float* x, int rows, long long cols, unsigned int* ws, float* out, void* stream);
int b200_minmax_normalise(float* x, long long n, unsigned int* ws, float* minmax_out, void* stream);
int b200_rgb_normalise(const void* x, int in_u8, long long n, int round_u8, float* out, void* stream);
int b200_resize_bilinear_aa(const float* x, long long planes, int H, int W, int OH, int OW, float* tmp, float* out,
                            void* stream);
int b200_resize_nearest(const float* x, long long planes, int H, int W, int OH, int OW, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_E2EFT_H_ */

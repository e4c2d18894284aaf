/*
 * multivae_b200 C-ABI — the drop-in boundary of the B200-native multimodal-VAE training step.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a CUDA stream handle
 * (cudaStream_t passed as void*), allocates nothing (all workspaces are caller-provided) and
 * returns an int status: 0 = ok, !=0 = error (message via mv_last_error(), thread-local).
 * No exception crosses this boundary.  The reference (MultiVae, pure Python/PyTorch) has no FFI of
 * its own; each function cites the reference code it replaces (paths relative to
 * /root/reference/src/multivae/).  The Python binding a MultiVae maintainer would add is the
 * ctypes stub in multivae_b200/_cabi.py.
 */
#ifndef MULTIVAE_B200_H
#define MULTIVAE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MV_OK 0
#define MV_ERR_ARG 1
#define MV_ERR_CUDA 2
#define MV_ERR_UNSUPPORTED 3

/* element types of activation / reconstruction buffers */
#define MV_F32 0
#define MV_BF16 1

/* decoder output distributions: models/base/base_utils.py:62-87 (set_decoder_dist) */
#define MV_DIST_NORMAL 0
#define MV_DIST_LAPLACE 1
#define MV_DIST_BERNOULLI 2
#define MV_DIST_CATEGORICAL 3   /* own entry points (mv_moe_lpx_cat_*): the log-prob needs a softmax over the class dimension */

/* latent prior/posterior families: models/mmvaePlus/mmvaePlus_model.py:57-73 */
#define MV_LATENT_LAPLACE 0
#define MV_LATENT_NORMAL 1

/* IWAE objectives: mmvaePlus_model.py:305-363, mmvae_model.py:238-292 */
#define MV_LOSS_IWAE 0
#define MV_LOSS_DREG 1

/* prior-expert modes of the PoE family: mvtcae_model.py:166 (never), mvae_model.py:75-79 (always,
 * stable_poe), mopoe_model.py:252-261 (only for the full subset) */
#define MV_PRIOR_NEVER 0
#define MV_PRIOR_ALWAYS_STABLE 1
#define MV_PRIOR_FULL_SUBSET 2

/* fused epilogues of the tensor-core GEMM / implicit-GEMM convolution */
#define MV_ACT_NONE 0
#define MV_ACT_RELU 1
#define MV_ACT_LRELU02 2
#define MV_ACT_SIGMOID 3

const char* mv_last_error(void);
/* library/version probe; returns the compiled SM architecture (100 for sm_100a) */
int mv_version(int* major, int* minor, int* sm_arch);
/* number of CUDA kernels this library has launched in this process (bench.py's "gpu_launches") */
unsigned long long mv_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Fused ELBO path, MoE family (MMVAE, MMVAE+).
 * ------------------------------------------------------------------------------------------- */

/* lpx[c,k,b] (+)= rescale * mask_r[b] * sum_d log p(x[b,d] | recon[c,k,b,d])
 * Replaces: recon_log_probs[recon_mod](x_recon, x).view(K,B,-1).mul(rescale).sum(-1) accumulated
 * over recon_mod — mmvaePlus_model.py:277-292, mmvae_model.py:208-225, base_utils.py:62-87.
 * recon: [C,K,B,D] (dtype MV_F32|MV_BF16), x: [B,D] f32, lpx: [C,K,B] f32, mask_r: [B] u8 or NULL. */
int mv_moe_lpx_fwd(const void* recon, int recon_dtype, const float* x, float* lpx, int C, int K, int B,
                   int64_t D, int dist, float dist_scale, float rescale, const uint8_t* mask_r,
                   int accumulate, void* stream);

/* g_recon[c,k,b,d] = g_loss * coef[c,k,b] * rescale * mask_r[b] * d/d(recon) log p(x|recon)
 * Backward of the above through the IWAE/DReG weights (coef = d loss / d lw, from mv_moe_lw_fwd). */
int mv_moe_lpx_bwd(const void* recon, int recon_dtype, const float* x, const float* coef,
                   const float* g_loss, void* g_recon, int C, int K, int B, int64_t D, int dist,
                   float dist_scale, float rescale, const uint8_t* mask_r, void* stream);

/* The same two kernels for ALL reconstructed modalities in one launch (host arrays of n_mod <= 8 device pointers /
 * per-modality scales; every modality has the same D, element type and distribution family — the PolyMNIST case).
 * lpx is overwritten with the sum over modalities.  Shapes that do not qualify run as n_mod single launches. */
int mv_moe_lpx_fwd_multi(int n_mod, const void* const* recon, int recon_dtype, const float* const* x, float* lpx, int C,
                         int K, int B, int64_t D, int dist, const float* dist_scale, const float* rescale,
                         const uint8_t* const* mask_r, void* stream);
int mv_moe_lpx_bwd_multi(int n_mod, const void* const* recon, int recon_dtype, const float* const* x, const float* coef,
                         const float* g_loss, void* const* g_recon, int C, int K, int B, int64_t D, int dist,
                         const float* dist_scale, const float* rescale, const uint8_t* const* mask_r, void* stream);

/* Categorical decoders (base_utils.py:28-59 cross_entropy_, :81-87): recon [C,K,B,P,V] holds logits over V classes at P positions,
 * x [B,P,V] the target probabilities (one-hot):
 *   lpx[c,k,b] (+)= rescale * mask_r[b] * sum_p sum_v x[b,p,v] * log_softmax(recon[c,k,b,p,:] + 1e-6)[v]
 *   g_recon[c,k,b,p,v] = g_loss * coef[c,k,b] * rescale * mask_r[b] * (x[b,p,v] - softmax(recon[c,k,b,p,:])[v] * sum_v' x[b,p,v']) */
int mv_moe_lpx_cat_fwd(const void* recon, int recon_dtype, const float* x, float* lpx, int C, int K, int B, int P, int V,
                       float rescale, const uint8_t* mask_r, int accumulate, void* stream);
int mv_moe_lpx_cat_bwd(const void* recon, int recon_dtype, const float* x, const float* coef, const float* g_loss, void* g_recon,
                       int C, int K, int B, int P, int V, float rescale, const uint8_t* mask_r, void* stream);

/* out[b] = logsumexp_r lw[r,b] - log R: the importance-sampled log-likelihood estimate over R = K (or n_modalities * K) samples
 * (compute_joint_nll / compute_cond_nll: base_ae_model.py:396-442, mmvaePlus_model.py:478-533, mmvae_model.py:366-444,
 * mopoe_model.py:468-595, mvae_model.py:241-317, mvtcae_model.py:214-289).  lw [R,B] f32, out [B] f32. */
int mv_logmeanexp(const float* lw, int R, int B, float* out, void* stream);

/* General Gaussian KL of base_utils.py:90-119, summed over the last dimension:
 *   out[r] = sum_l 0.5 * (plv - lv + exp(lv - plv) + (mu - pm)^2 / exp(plv) - 1)
 * mu, lv [rows, L]; prior_mu, prior_lv [prior_rows, L] with prior_rows = rows or 1 (broadcast).  The backward returns the
 * gradients of all four inputs given g_out [rows] (g_prior_* may be NULL; for a broadcast prior they are summed over rows). */
int mv_gauss_kl_fwd(const float* mu, const float* lv, const float* prior_mu, const float* prior_lv, float* out, int64_t rows, int L,
                    int prior_rows, void* stream);
int mv_gauss_kl_bwd(const float* mu, const float* lv, const float* prior_mu, const float* prior_lv, const float* g_out, float* g_mu,
                    float* g_lv, float* g_prior_mu, float* g_prior_lv, int64_t rows, int L, int prior_rows, void* stream);

/* Latent terms + importance weights + loss for one batch, and their unit gradients.
 * Replaces _compute_k_lws + _dreg_looser/_iwae_looser (mmvaePlus_model.py:230-363) and
 * compute_k_lws + dreg_looser/iwae_looser (mmvae_model.py:160-292).
 *   u [C,K,B,L], w [C,K,B,Lw] (Lw may be 0: MMVAE), posterior mu_u/sig_u [C,B,L], mu_w/sig_w [C,B,Lw],
 *   prior pz_mean/pz_std [L+Lw], lpx [C,K,B], masks [C,B] u8 or NULL (mask of modality c for sample b).
 * Outputs: lw, wk, coef [C,K,B]; loss_b [B] (per-sample loss, already negated and divided by n_mods);
 *   unit gradients (for d loss = 1): g_u [C,K,B,L], g_w [C,K,B,Lw] (direct terms only, NOT yet multiplied
 *   by the DReG wk); g_mu_u,g_sig_u [C,B,L], g_mu_w,g_sig_w [C,B,Lw] (zero-filled when detach_post != 0);
 *   g_pz_std [B,L+Lw] per-sample partials of d loss / d prior std.
 * skip_u_prior != 0 leaves log p(u) out of lw (and of g_u / g_pz_std): CMVAE's mixture-of-clusters prior over the shared code is
 * added by the caller through `lpx` (cmvae_model.py:305-340); the private code keeps the fixed prior pz[L:]. */
int mv_moe_lw_fwd(const float* u, const float* w, const float* mu_u, const float* sig_u, const float* mu_w,
                  const float* sig_w, const float* pz_mean, const float* pz_std, const float* lpx,
                  const uint8_t* masks, float* lw, float* wk, float* coef, float* loss_b, float* g_u,
                  float* g_w, float* g_mu_u, float* g_sig_u, float* g_mu_w, float* g_sig_w, float* g_pz_std,
                  int C, int K, int B, int L, int Lw, int latent_kind, int loss_kind, float beta,
                  int detach_post, int skip_u_prior, void* stream);

/* Reparameterised sampling of the MoE family in one launch per direction (replaces _log_var_to_std + the rsample / stack / cat glue
 * of mmvaePlus_model.py:113-186 and mmvae_model.py:66-130, and their autograd twins):
 *   sig_u = std(lv_u) [C,B,L], sig_w = std(lv_w) [C,B,Lw]    std_kind 0: softmax(lv) * dim + 1e-6, 1: exp(lv / 2), 2: softplus(lv) + 1e-6
 *   U [C,K,B,L] = mu_u + sig_u * noise_u,  W [C,K,B,Lw] = mu_w + sig_w * noise_w
 *   Z [C(recon r),C(cond c),K,B,L+Lw] = cat(U[c], r == c ? W[c] : prior_mean[r] + prior_std[r] * noise_x[c, j(r)])   decoder inputs
 * noise_x [C,C-1,K,B,Lw] holds, for every conditioning modality, one standard draw per OTHER modality (in modality order).
 * Lw = 0 (MMVAE): w / prior / Z pointers are NULL.  The backward sums the gradients reaching the samples (g_U, g_W: direct
 * terms of the ELBO; g_Z: the decoders), multiplies them by the DReG weights wk [C,K,B] when given (mmvaePlus_model.py:330-338;
 * prior samples are not weighted), adds the gradients arriving at sig_u / sig_w themselves and applies the derivative of std():
 * outputs g_mu_*, g_lv_* (same shapes as the inputs) and g_prior_std [C,Lw] (summed over c, k, b). */
int mv_moe_sample_fwd(const float* mu_u, const float* lv_u, const float* mu_w, const float* lv_w, const float* prior_mean,
                      const float* prior_std, const float* noise_u, const float* noise_w, const float* noise_x, int std_kind,
                      float* sig_u, float* sig_w, float* U, float* W, float* Z, int C, int K, int B, int L, int Lw, void* stream);
int mv_moe_sample_bwd(const float* lv_u, const float* lv_w, const float* sig_u, const float* sig_w, const float* noise_u,
                      const float* noise_w, const float* noise_x, const float* g_U, const float* g_W, const float* g_Z,
                      const float* g_sig_u, const float* g_sig_w, const float* wk, int std_kind, float* g_mu_u, float* g_lv_u,
                      float* g_mu_w, float* g_lv_w, float* g_prior_std, int C, int K, int B, int L, int Lw, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused posterior aggregation, PoE family (MVTCAE, MVAE, MoPoE).
 *   mu, lv      [M,B,L] f32  unimodal posterior parameters (encoder outputs)
 *   masks       [M,B] u8 or NULL: unavailable experts are excluded (the reference sets their log-variance to
 *               +inf: mvtcae_model.py:126-130, mvae_model.py:65-70)
 *   subsets     [S] u32 bitmasks (bit m = modality m), the table of mopoe_model.py:76-106 / mvae_model.py:159-172
 *   sel         [B] i32 index into `subsets` of the posterior each sample is drawn from
 *               (deterministic_mixture_component_selection, mopoe_model.py:435-465), NULL = subset 0
 *   w           [S,B] f32 per-sample subset weights of the KL sum, or NULL = w_uniform for all
 *   prior_mode  MV_PRIOR_*; stable != 0 selects stable_poe (base_utils.py:133-147), else poe with `eps` (:122-130)
 * Outputs: z [B,L] = mu_sel + exp(0.5*lv_sel)*noise (rsample_from_gaussian, base_utils.py:150-172), optional
 *   joint_mu/joint_lv [B,L], kl_b [B] = sum_s w_s * KL(q_s || N(0,I)) (mopoe_model.py:108-145, mvae_model.py:105,
 *   mvtcae_model.py:52-54), kldm_b [M,B] = KL(q_sel || q_m) (mvtcae_model.py:82-88) or NULL.
 * mv_poe_bwd recomputes the aggregation and returns d/d(mu,lv) given d/dz, d/dkl_b, d/dkldm_b.
 * ------------------------------------------------------------------------------------------- */
int mv_poe_fwd(const float* mu, const float* lv, const uint8_t* masks, const uint32_t* subsets, int S,
               const int32_t* sel, const float* w, float w_uniform, const float* noise, int prior_mode, int stable,
               float eps, float* z, float* joint_mu, float* joint_lv, float* kl_b, float* kldm_b, int M, int B,
               int L, void* stream);
int mv_poe_bwd(const float* mu, const float* lv, const uint8_t* masks, const uint32_t* subsets, int S,
               const int32_t* sel, const float* w, float w_uniform, const float* noise, int prior_mode, int stable,
               float eps, const float* g_z, const float* g_kl, const float* g_kldm, float* g_mu, float* g_lv, int M,
               int B, int L, void* stream);


/* ---------------------------------------------------------------------------------------------
 * Tensor-core contractions of the encoders / decoders (tcgen05 + TMA, bf16 operands, fp32 accumulate).
 *
 * Activations of a convolutional stage live in a "shared-halo" NHWC matrix: rows = flat pixels,
 * columns = channels (bf16).  Image i of height H, width W occupies rows [i*S, (i+1)*S), S = (H+1)*(W+1):
 * first one zero row of W+1 pixels, then H rows of W real pixels followed by one zero pixel; after the
 * last image comes one more zero row, so the matrix has n_img*S + (W+1) rows.  In this layout filter tap
 * (r,s) of a 3x3 / stride 1 / pad 1 convolution reads the same matrix shifted by (r-1)*(W+1) + (s-1) rows.
 *
 * mv_tapgemm:  out[p, n] = epilogue( sum_t sum_c A[p + tap_off[t], c] * Wt[t*N_total + n, c] )
 *   Replaces (forward and data-gradient) nn.Conv2d 3x3 / 1x1 and nn.Linear of models/nn/mmnist.py:214-366
 *   (ResnetBlock :229-241, fc :289-295,339, conv_img :287,352-354) and default_architectures.py:31-39,237-241.
 *   epilogue:  y = act(acc + bias[n]);  if dact1 / dmask1: y *= (dact1[p,n] > 0 ? 1 : slope1)
 *              if out2 && out2_pre: out2[p,n] = y
 *              o = alpha*y + (res ? res[p,n] : 0);  out[p,n] = o
 *              if out2 && !out2_pre: out2[p,n] = alpha2 * o * (dact2 ? (dact2[p,n] > 0 ? 1 : slope2) : 1)
 *              (out2_mask / dmask2: the same with the activation-derivative source stored as one sign bit per element)
 *   rows that are halo positions (img_stride > 0) are written as zeros.  out_mode 1 scatters the first
 *   n_valid columns of the valid pixels to a dense NCHW bf16 image tensor [n_img, n_valid, H, W].
 * ------------------------------------------------------------------------------------------- */
typedef struct mv_tapgemm_args {
  const void* A;        /* bf16 [a_rows, a_ld] (first Cin columns used) */
  int64_t a_rows;
  int32_t a_ld, Cin;    /* Cin: 16 or a multiple of 64 */
  const void* Wt;       /* bf16 [T * N_total, Cin], K-major */
  int32_t T;
  int32_t tap_off[9];
  int32_t N_total, BN;  /* output columns, columns per CTA tile (16/32/64/128) */
  int64_t P;            /* output rows */
  const float* bias;    /* [N_total] or NULL */
  int32_t act;          /* MV_ACT_* */
  float alpha;
  const void* res; int32_t res_ld;
  const void* dact1; int32_t dact1_ld; float slope1;
  void* out; int32_t out_ld;
  void* out2; int32_t out2_ld; int32_t out2_pre; float alpha2;
  const void* dact2; int32_t dact2_ld; float slope2;
  int32_t img_stride, Wp, W, H, n_img;   /* halo geometry (img_stride = 0: plain GEMM, no mask) */
  int32_t out_mode, n_valid;
  /* sign masks of an activation, one bit per element, 64-bit word per row ([rows padded to a multiple of 126][64 columns]):
   *   out2_mask  written INSTEAD of out2 (out2 must be NULL): bit n = (act(acc + bias)[p, n] > 0); 3x3 / 64-output layers only
   *   dmask2     read INSTEAD of dact2: the second output is alpha2 * o * (bit ? 1 : slope2); N_total = 64 only */
  void* out2_mask;
  const void* dmask2;
  const void* dmask1;   /* read INSTEAD of dact1 (y *= bit ? 1 : slope1); 64-output layers only */
  /* res_mask (3x3 / 64-output layers only): the residual is read as res[p, n] * (bit n of res_mask[p] ? res_scale_pos : res_scale_neg).
   * Lets a producer store ONE tensor t = a * g * lrelu'(d) and the consumer recover g = t / (a * lrelu'(d)) from d's sign bits
   * (the image head's data gradient writes only the pre-scaled gradient; the block's skip connection un-scales it here). */
  const void* res_mask; float res_scale_pos, res_scale_neg;
  /* fused 1x1 term (plain 3x3 convolutions with 64 inputs and 64 outputs only): out[p, n] additionally gets sum_c A2[p, c] * W2[n, c]
   * (A2 bf16 [P][a2_ld], 64 columns used; W2 bf16 [64][64] K-major).  The data gradient of a ResnetBlock with a learned shortcut
   * (models/nn/mmnist.py:243-251), conv_0^T(g_h) + shortcut^T(g_out), in one launch: no separate 1x1 kernel, no residual tensor. */
  const void* A2; int32_t a2_ld; const void* W2;
} mv_tapgemm_args;
int mv_tapgemm(const mv_tapgemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * General tensor-core GEMM of the fully connected layers (tcgen05 + TMA, bf16 operands, fp32 accumulate):
 *
 *   out[m, n] (+)= epilogue( alpha * sum_k A(m, k) * B(n, k) )
 *
 * A is stored [M][K] (a_mn = 0, "K-major") or [K][M] (a_mn = 1, "MN-major"); B is stored [N][K] (b_mn = 0) or [K][N] (b_mn = 1);
 * a_ld / b_ld are the row pitches in elements (multiples of 8), base pointers 16-byte aligned.  No padding of M, N, K is needed
 * (out-of-range elements are zero-filled by TMA).  One entry point covers the three products of a Linear layer
 * (models/nn/default_architectures.py:21-130,225-258; nn.Linear of mmnist.py:98,181,289-295,339 and autograd's addmm backward):
 *   forward         Y = X W^T + b      A = X  [B][K],  B = W [N][K]
 *   data gradient   dX = dY W          A = dY [B][N],  B = W [N][K] with b_mn = 1 (the reduction runs over W's rows)
 *   weight gradient dW += dY^T X       A = dY [B][N] with a_mn = 1,  B = X [B][K] with b_mn = 1,  out_kind = MV_OUT_F32_ADD
 * epilogue (MV_OUT_BF16 / MV_OUT_F32):  y = act(alpha * acc + bias[n]);  if dact: y *= (dact[m, n] > 0 ? 1 : dslope)
 *   (dact = the saved OUTPUT of the ReLU / LeakyReLU that produced this layer's input: the activation derivative of the
 *   layer below fused into the data-gradient GEMM).  MV_OUT_F32_ADD adds alpha * acc (+ bias) into fp32 `out` with vector
 *   reductions and splits the reduction over CTAs when the output has fewer tiles than the GPU has SMs.
 * ------------------------------------------------------------------------------------------- */
#define MV_OUT_BF16 0
#define MV_OUT_F32 1
#define MV_OUT_F32_ADD 2
typedef struct mv_gemm_args {
  const void* A; int64_t M; int32_t a_ld, a_mn;
  const void* B; int32_t N, b_ld, b_mn;
  int32_t K;
  const float* bias;    /* [N] or NULL */
  int32_t act;          /* MV_ACT_* */
  float alpha;
  const void* dact; int32_t dact_ld; float dslope;   /* bf16 [M][dact_ld] or NULL */
  void* out; int32_t out_ld, out_kind;               /* bf16 or fp32 [M][out_ld] */
} mv_gemm_args;
int mv_gemm(const mv_gemm_args* args, void* stream);
/* out[n] += sum_p G[p, n] for any N % 8 == 0 (bias gradients of the fully connected layers; G bf16 [P][ld], out fp32) */
int mv_colsum_any(const void* G, int64_t P, int ld, int N, float* out, void* stream);
/* out = g * act'(.) elementwise over n elements (n % 8 == 0): the derivative is taken from the SAVED OUTPUT y of the activation
 * (Sigmoid: y (1 - y); ReLU / LeakyReLU: y > 0 ? 1 : slope); g fp32 or bf16 (MV_F32 / MV_BF16), y and out bf16.
 * Replaces autograd's sigmoid_backward / threshold_backward in front of the gradient GEMMs of a layer. */
int mv_act_bwd(const void* g, int g_dtype, const void* y, void* out, int64_t n, int act, float slope, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Strided and transposed convolutions of the small convolutional networks (models/nn/svhn.py:7-70, mmnist.py:78-110,173-207) as
 * gather + mv_gemm.  The patch matrix `cols` is [n_img * grid_h * grid_w][ld] with columns ordered tap-major (t*C + c, channels
 * contiguous: vector gathers on NHWC tensors) or channel-major (c*T + t = a torch Conv2d / ConvTranspose2d weight flattened over
 * its last three dimensions), see tc_order; columns >= C*kh*kw are zero padding.  `H, W, C, nchw` describe the image-side tensor (NHWC when nchw = 0).
 *   mv_im2col  cols[(n, gy, gx), c*T + t] = src[n, gy*stride - pad + ky, gx*stride - pad + kx, c]  (0 outside), src fp32 or bf16
 *              (nn.Conv2d forward: grid = the convolution's output; nn.ConvTranspose2d backward: src = gradient of its output,
 *              grid = its input)
 *   mv_col2im  dst[n, y, x, c] = act(bias[c] + sum over the taps t that reach (y, x) of cols[(n, gy, gx), c*T + t])
 *              * (dact ? (dact[n, y, x, c] > 0 ? 1 : dslope) : 1), dst / dact bf16 in the image-side layout; cols fp32 or bf16
 *              (the GEMM that produces it writes fp32, so the sum over the taps sees unrounded partial products)
 *              (nn.ConvTranspose2d forward after `cols = x W`; nn.Conv2d data gradient after `cols = dY W`)
 *   mv_chan_sum_nchw  out[c] += sum_{n, y, x} g[n, c, y, x]  (bias gradient of an NCHW layer, g bf16, out fp32)
 * ------------------------------------------------------------------------------------------- */
typedef struct mv_conv_geom {
  int32_t n_img, H, W, C, nchw;
  int32_t kh, kw, stride, pad;
  int32_t grid_h, grid_w, ld;
  int32_t tc_order;   /* columns of the patch matrix: 1 = t*C + c (tap-major, channels contiguous), 0 = c*T + t (a torch weight flattened) */
} mv_conv_geom;
int mv_im2col(const void* src, int src_dtype, void* cols, const mv_conv_geom* geom, void* stream);
int mv_col2im(const void* cols, int cols_dtype, void* dst, const mv_conv_geom* geom, const float* bias, int act, const void* dact,
              float dslope, void* stream);
int mv_chan_sum_nchw(const void* g, int n_img, int C, int HW, float* out, void* stream);

/* Weight gradient of a tap-GEMM layer:  dW[t, n, c] += sum_p G[p, n] * X[p + tap_off[t], c]   (fp32, accumulating).
 * Replaces the weight-gradient half of autograd's convolution_backward / addmm backward for the layers above
 * (the reference reaches it through loss.backward(), trainers/base/base_trainer.py:359).
 *   X bf16 [x_rows, x_ld] (Cin = 64 / 128 / 256 columns used), G bf16 [g_rows, g_ld] (N = 16 / 64 / 128 columns),
 *   dW fp32 [T, N, Cin], must be initialised by the caller (zeros, or a running gradient).
 *   db (optional, fp32 [N]): bias gradient db[n] += sum_p G[p, n], computed from the same shared-memory tiles. */
int mv_wgrad(const void* X, int64_t x_rows, int x_ld, int Cin, const void* G, int64_t g_rows, int g_ld, int N, int T,
             const int* tap_off, int64_t P, float* dW, float* db, void* stream);
/* same, for a slice of N columns [n_offset, n_offset + N) of a layer with N_total output channels: G points at the
 * slice's first column, dW is the full [T, N_total, Cin] tensor */
int mv_wgrad_slice(const void* X, int64_t x_rows, int x_ld, int Cin, const void* G, int64_t g_rows, int g_ld, int N, int T,
                   const int* tap_off, int64_t P, float* dW, int N_total, int n_offset, float* db, void* stream);

/* same contraction, accumulated straight into a gradient tensor in the torch Conv2d layout: dW[(n_offset + n), c, t] (fp32
 * [N_out, Cin, T], e.g. `weight.grad` itself) += ..., only for n < n_valid (output columns beyond are padding); db[n_offset + n]
 * likewise.  Lets the training step skip the permuted copy and the separate `grad += ...` kernel per parameter. */
int mv_wgrad_nct(const void* X, int64_t x_rows, int x_ld, int Cin, const void* G, int64_t g_rows, int g_ld, int N, int T,
                 const int* tap_off, int64_t P, float* dW, int n_offset, int n_valid, float* db, void* stream);

/* HBM-bound helpers of the shared-halo layout (all tensors bf16 unless noted):
 *   mv_upsample2x_fwd  nn.Upsample(scale_factor=2) (models/nn/mmnist.py:345): in (H x W, C ch) -> out (2H x 2W, C ch)
 *   mv_upsample2x_bwd  its gradient: g_in = sum of the 2x2 children of g_out; optional g_pre = alpha*g_in*lrelu'(act)
 *                      (the residual-branch gradient of the ResnetBlock below, mmnist.py:243-246)
 *   mv_head_grad_pack  dense NCHW gradient [n_img, ch, H, W] times lrelu'(y_out) -> halo matrix [P, 16]
 *   mv_colsum          out[n] += sum_p G[p, n] (bias gradients; out fp32) */
int mv_upsample2x_fwd(const void* in, void* out, int n_img, int H, int W, int C, void* stream);
int mv_upsample2x_bwd(const void* g_out, const void* act, void* g_in, void* g_pre, int n_img, int H, int W, int C, float alpha,
                      float slope, void* stream);
int mv_head_grad_pack(const void* g, const void* y_out, void* out, int n_img, int H, int W, int ch, float slope, void* stream);
int mv_colsum(const void* G, int64_t P, int ld, int N, float* out, void* stream);
/*   mv_avgpool3s2_fwd  nn.AvgPool2d(3, stride=2, padding=1) of the ResNet encoder (models/nn/mmnist.py:278): (H x W) -> (H/2 x W/2)
 *   mv_avgpool3s2_bwd  its gradient (+ optional g_pre = alpha*g_in*lrelu'(act), like mv_upsample2x_bwd)
 *   mv_scale_dact      out = alpha * g * lrelu'(act) on [P, C] matrices */
int mv_avgpool3s2_fwd(const void* in, void* out, int n_img, int H, int W, int C, void* stream);
int mv_avgpool3s2_bwd(const void* g_out, const void* act, void* g_in, void* g_pre, int n_img, int H, int W, int C, float alpha,
                      float slope, void* stream);
int mv_scale_dact(const void* g, const void* act, void* out, int64_t P, int C, float alpha, float slope, void* stream);
/*   mv_lrelu_fwd       out = leaky_relu(x, slope) on a [P, C] matrix: the pre-activation `actvn(x)` of the CUB ResNet blocks
 *                      (models/nn/cub.py:281-283, 296-299), whose shortcut reads the raw x */
int mv_lrelu_fwd(const void* x, void* out, int64_t P, int C, float slope, void* stream);
/* Index gathers of the fully connected layers that read / write the halo layout (the `fc` of DecoderResnetMMNIST, models/nn/mmnist.py:339,
 * and the `fc_mu / fc_lv` heads of EncoderResnetMMNIST, :289-295; the same layers of nn/cub.py).  idx: device int64.
 *   mode 0 (rows):    out[j, c] = src[idx[j], c]  (0 where idx[j] >= src_rows)      mode 1 (columns): out[r, j] = src[r, idx[j]]
 *   mv_gather_cast  fp32 -> bf16 GEMM operand; columns [out_cols, out_ld) are written as zeros
 *   mv_gather_f32   fp32 -> fp32, out = beta * out + gathered (beta = 0: plain store): gradients back in the parameter's layout */
int mv_gather_cast(const float* src, int64_t src_rows, int64_t src_cols, int64_t src_ld, const int64_t* idx, int64_t out_rows,
                   int64_t out_cols, int mode, void* out_bf16, int64_t out_ld, void* stream);
int mv_gather_f32(const float* src, int64_t src_rows, int64_t src_cols, int64_t src_ld, const int64_t* idx, int64_t out_rows,
                  int64_t out_cols, int mode, float* out, int64_t out_ld, float beta, void* stream);

/* All convolution-weight packs of a network in one launch.  Each item turns an fp32 Conv2d weight [N, C, kh, kw] (T = kh*kw)
 * into the bf16 operand matrices of mv_tapgemm: dst_fwd [T * Npad, Cpad] (row t*Npad + n, column c) for the forward pass and
 * dst_dgrad [T * Cpad, Npad] (row (T-1-t)*Cpad + c, column n: taps flipped, channel roles swapped) for the data gradient;
 * either destination may be NULL.  Padding rows / columns (Npad > N, Cpad > C) are written as zeros.  Replaces the per-layer
 * permute / flip / cast of the reference-side autograd graph (nn.Conv2d weights of models/nn/mmnist.py:229-241,287,352). */
#define MV_PACK_MAX_ITEMS 24
typedef struct mv_pack_item {
  const void* src;      /* fp32 [N, C, T] */
  void* dst_fwd;        /* bf16 [T * Npad, Cpad] or NULL */
  void* dst_dgrad;      /* bf16 [T * Cpad, Npad] or NULL */
  int32_t N, C, T, Npad, Cpad;
} mv_pack_item;
int mv_pack_conv_weights(const mv_pack_item* items, int n_items, void* stream);

/* Weight hand-over for the tap-major column order, all layers of a network in one launch (mv_pack_item: src, dst_fwd, N, C, T,
 * Cpad = row pitch of the packed matrix >= T*C; dst_dgrad / Npad unused):
 *   mv_pack_tc        dst_fwd bf16 [N][Cpad]: dst[n, t*C + c] = src fp32 [N][C][T], columns >= T*C zero
 *   mv_unpack_tc_add  dst_fwd fp32 [N][C][T] (a parameter's .grad) += src fp32 [N][Cpad] at [n, t*C + c] */
int mv_pack_tc(const mv_pack_item* items, int n_items, void* stream);
int mv_unpack_tc_add(const mv_pack_item* items, int n_items, void* stream);

/* The reverse hand-over for the gradients of a whole network in one launch: dst[n, c, t] += src[t, n, c] for every item, where
 * src is the fp32 [T, Npad, Cpad] buffer mv_wgrad accumulated into (swapped != 0: [T, Cpad, Npad], the role-swapped image
 * convolution) and dst the parameter's own gradient in the torch Conv2d layout [N, C, T] (`weight.grad`; a bias gradient is an
 * item with C = T = 1).  Replaces one permuted `grad += dW` kernel per parameter (autograd's AccumulateGrad). */
typedef struct mv_unpack_item {
  const void* src;      /* fp32 [T, Npad, Cpad] (or [T, Cpad, Npad]) */
  void* dst;            /* fp32 [N, C, T], accumulated into */
  int32_t N, C, T, Npad, Cpad, swapped;
} mv_unpack_item;
int mv_unpack_wgrad_add(const mv_unpack_item* items, int n_items, void* stream);

#ifdef __cplusplus
}
#endif
#endif

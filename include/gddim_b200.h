/*
 * gddim_b200 -- C ABI of the B200-native gDDIM sampling hot path.
 *
 * The reference (qsh-zh/gDDIM) has no FFI layer: its boundary is plain Python calls
 * (SURVEY.md 8b).  Each entry point below names the reference function it stands in for
 * (paths relative to the reference repository root).  All pointers are caller-owned; the library never
 * frees caller memory.  Functions return 0 on success and a negative value on error;
 * gddim_last_error() returns a message for the calling thread.
 *
 * Layouts: images are NHWC.  The CLD state in *reference layout* is [B,H,W,C,2] (last axis = (x, v),
 * cld_jax/sde_lib.py:270-274); in *net layout* it is [B,H,W,2C] with channel g*C+d
 * (cld_jax/models/utils.py:153).  "dev" pointers are CUDA device pointers on the context's device;
 * `stream` is a cudaStream_t passed as void* (NULL = default stream).
 */
#ifndef GDDIM_B200_H
#define GDDIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDDIM_ABI_VERSION 1

typedef struct gddim_ctx gddim_ctx;         /* one per GPU: weights, workspace, CUDA graphs */
typedef struct gddim_cld gddim_cld;         /* host-side CLD SDE tables (fp64) */
typedef struct gddim_blur gddim_blur;       /* host-side blur SDE tables (fp64) */
typedef struct gddim_sampler gddim_sampler; /* a configured sampler bound to a ctx */

/* config.model.* / config.data.* fields read by cld_jax/models/ncsnpp.py:43-67 */
typedef struct {
  int image_size;          /* config.data.image_size */
  int data_channels;       /* config.data.num_channels */
  int state_mult;          /* 2 for CLD (x and v stacked on channels), 1 for blur */
  int nf;
  int n_levels;            /* len(ch_mult) */
  int ch_mult[8];
  int num_res_blocks;
  int n_attn;
  int attn_resolutions[8];
  int fir;                 /* FIR [1,3,3,1] resampling (1) or naive repeat / mean (0) */
  int skip_rescale;
  int progressive_input;   /* 0 = none, 1 = residual */
  int embedding_type;      /* 0 = fourier, 1 = positional */
  int conditional;
  int centered;            /* config.data.centered (if 0 the net maps x -> 2x-1 first, ncsnpp.py:136-138) */
} gddim_model_cfg;

const char* gddim_last_error(void);
int gddim_abi_version(void);
/* 1 if a CUDA device is usable from this process, else 0 (never throws) */
int gddim_cuda_available(void);

/* ---- context / parameters: models/utils.py:109-125 init_model, run_lib.py:707-711 restore + replicate ---- */
int gddim_ctx_create(int device, const gddim_model_cfg* cfg, int max_batch, gddim_ctx** out);
/* flags: GDDIM_CTX_PRECISE_WEIGHTS = every 3x3 / shortcut convolution keeps its weights as an fp16 (hi, lo) pair and runs
 * its K loop twice (A x W_hi + A x W_lo): removes the weight half of the operand rounding (parity mode; the reference
 * computes in fp32).  Twice the weight memory and tensor work of those layers. */
#define GDDIM_CTX_PRECISE_WEIGHTS 1u
int gddim_ctx_create_ex(int device, const gddim_model_cfg* cfg, int max_batch, unsigned flags, gddim_ctx** out);
void gddim_ctx_destroy(gddim_ctx* ctx);
/* parameter inventory in Flax naming (e.g. "ResnetBlockBigGANpp_3/Conv_0/kernel"), Flax layouts
 * (conv HWIO, dense (in,out)).  kind: 0 variance_scaling(fan_avg, uniform), 1 zeros, 2 ones, 3 normal */
int gddim_param_count(const gddim_ctx* ctx);
int gddim_param_spec(const gddim_ctx* ctx, int index, char* name_buf, int name_buf_len, int shape[4], int* ndim,
                     int* kind, float* scale);
int gddim_param_set(gddim_ctx* ctx, const char* name, const float* host_data, size_t n_elem);
/* packs the loaded parameters into device layouts (fp16 K-major GEMM operands) and builds the launch plan */
int gddim_ctx_finalize(gddim_ctx* ctx);
/* 0 = tcgen05/TMA kernels (default), 1 = CUDA-core reference kernels (on-GPU validation only) */
int gddim_ctx_set_gemm_impl(gddim_ctx* ctx, int impl);
size_t gddim_ctx_workspace_bytes(const gddim_ctx* ctx);
/* Static launch plan of the network (available right after gddim_ctx_create, no GPU needed): number of planned ops, and
 * for op `index` its tag ("ResnetBlockBigGANpp_3/conv1", "AttnBlockpp_0/gn_qkv", ...) and kind (0 stem, 1 groupnorm,
 * 2 conv/GEMM, 3 head, 4 im2col, 5 V transpose, 6 small attention, 7 row softmax, 8 fused attention, 9 GroupNorm+qkv). */
int gddim_ctx_plan_size(const gddim_ctx* ctx);
int gddim_ctx_plan_op(const gddim_ctx* ctx, int index, char* tag_buf, int tag_buf_len, int* kind);
long long gddim_ctx_launch_count(const gddim_ctx* ctx);   /* kernels launched by this ctx so far */

/* per-op CUDA-event timing of the network evaluation (eager launches; CUDA graphs are bypassed while on).
 * ms_by_kind[8]: accumulated ms per op kind {0 stem, 1 groupnorm, 2 conv/gemm (tcgen05), 3 head, 4 im2col,
 * 5 transpose, 6 small attention}; gemm_flops: algorithmic FLOPs (2 M N K) of the GEMM launches timed. */
int gddim_ctx_set_profile(gddim_ctx* ctx, int on);
int gddim_ctx_get_profile(const gddim_ctx* ctx, double* ms_by_kind, double* gemm_flops, long long* gemm_launches);
int gddim_ctx_dump_profile(const gddim_ctx* ctx, const char* path);   /* per-op CSV */
/* algorithmic HBM bytes of all GroupNorm(+swish, +resample) ops over the profiled forwards (fp32 in, fp16 out) */
int gddim_ctx_get_profile_hbm(const gddim_ctx* ctx, double* norm_bytes);

/* ---- score network: NCSNpp.apply / get_eps_fn's model call (models/utils.py:128-166; ncsnpp.py:41-243) ----
 * x_dev, out_dev: fp32 [batch, S, S, data_channels*state_mult] in net layout; t = diffusion time (labels = 999 t
 * are formed inside, models/utils.py:172).  One t for the whole batch (as on the sampling path). */
int gddim_unet_forward(gddim_ctx* ctx, const float* x_dev, float t, float* out_dev, int batch, void* stream);

/* ---- CLD SDE tables: cld_jax/sde_lib.py:45-118 (CLD.__init__), :182-253, :289-319 ---- */
int gddim_cld_create(double m_inv, double beta_0, double beta_1, double vv_gamma, double numerical_eps, double R_dt,
                     int is_R_rk, gddim_cld** out);
void gddim_cld_destroy(gddim_cld* cld);
int gddim_cld_R(const gddim_cld* cld, const double* t, int n, double* out /*[n,2,2]*/);
int gddim_cld_psi(const gddim_cld* cld, const double* s, const double* t, int n, double* out /*[n,2,2]*/);
int gddim_cld_F(const gddim_cld* cld, double t, double* out /*[2,2]*/);
int gddim_cld_G(const gddim_cld* cld, double t, double* out /*[2,2]*/);
int gddim_cld_eps_integrand(const gddim_cld* cld, const double* t, int n, double* out /*[n,2,2]*/);
/* CLD.get_deis_coef(order, rev_ts): out [n_ts-1, order+3, 2, 2] */
int gddim_cld_deis_coef(const gddim_cld* cld, int order, const double* rev_ts, int n_ts, double* out);
/* CLD.prepare_order0_coef(rev_ts): mean_out, eps_out [n_ts-1, 2, 2] */
int gddim_cld_order0_coef(const gddim_cld* cld, const double* rev_ts, int n_ts, double* mean_out, double* eps_out);
/* LambdaSDE(sde, lambda_coef, use_order0).get_deis_coef(order, rev_ts) (sde_lib.py:435-454): out [n_ts-1, order+4, 2, 2]
 * = (x_coef, order+2 eps slots, conditional reverse covariance) */
int gddim_cld_sdeis_coef(const gddim_cld* cld, double lambda_coef, int use_order0, int order, const double* rev_ts,
                         int n_ts, double* out);
/* the factor applied to standard normals by jax.random.multivariate_normal(method='svd'): U sqrt(S), sign-normalised */
int gddim_mvn_factor_svd(const double* cov /*[2,2]*/, double* out /*[2,2]*/);
/* sampling.get_rev_ts (cld_jax/sampling.py:241-249; blur_jax/sampling.py:42-51): out [num_step+1] */
int gddim_rev_ts(double T, double eps, int ts_order, int num_step, double* out);

/* ---- blur SDE tables: blur_jax/sde_lib.py:18-163 ---- */
int gddim_blur_create(double sigma_blur_max, double sampling_eps, gddim_blur** out);
void gddim_blur_destroy(gddim_blur* b);
double gddim_blur_sampling_T(const gddim_blur* b);
int gddim_blur_y_mean_coef(const gddim_blur* b, double t, double* out /*[32,32]*/);
double gddim_blur_y_std_coef(const gddim_blur* b, double t);
double gddim_blur_t2alpha(const gddim_blur* b, double t);

/* ---- update operators on device arrays ---- */
/* deis.multistep_ab_step (cld_jax/deis.py:141-151), reference layout: x, new_eps, x_out [n_pairs,2];
 * eps_pred, eps_pred_out [order+1, n_pairs, 2]; deis_coef host [order+3,2,2] fp32 */
int gddim_multistep_ab_step(const float* x_dev, const float* deis_coef_host, const float* new_eps_dev,
                            const float* eps_pred_dev, float* x_out_dev, float* eps_pred_out_dev, int order,
                            long long n_pairs, void* stream);
/* blur_jax/multistep.py:94-98 ab_step: ei_coef host [n_hist+2] */
int gddim_scalar_ab_step(const float* x_dev, const float* ei_coef_host, const float* new_eps_dev,
                         const float* eps_pred_dev, float* x_out_dev, float* eps_pred_out_dev, int n_hist, long long n,
                         void* stream);
/* '(b ... d g) <-> b ... (g d)' (models/utils.py:153,158) */
int gddim_relayout(const float* src_dev, float* dst_dev, long long n_pix, int C, int to_net, void* stream);
/* blur.batch_img_dct / batch_img_idct (blur_jax/blur.py:99-107) on [B,32,32,C] */
int gddim_dct2d_32(const float* in_dev, float* out_dev, int batch, int C, int forward, void* stream);

/* ---- operator level (layers.py:66-107 ddpm_conv1x1/3x3, :467-478 NIN; layerspp.py:74-78 attention einsums;
 *      flax nn.GroupNorm + swish + up_or_down_sampling resamplers, layerspp.py:196-213) ----
 * out[m,n] = (sum_seg sum_tap sum_c A_seg[pixel(m)+tap, c] Wt[n, k] * rowscale[m] + bias[n] + bias2[n]
 *             + residual[m,n]) * scale, m over the [B,H,W] pixel grid.  A*, w, out16: fp16; K-major weights
 * Wt [N, w_ld] with k = w_koff + seg-major, tap-major, channel-minor.  epi = 1: row softmax (N == 256).
 * epi = 2: the GroupNorm (+ swish) that FOLLOWS the convolution (layerspp.py:218 h = act(GroupNorm_1(h))) applied by
 * the epilogue itself: out16 = act(GN(out)) with flax semantics (eps, contiguous groups, statistics per image);
 * gn_gamma / gn_beta [N], out32 / residual / rowscale must be NULL; geometries: gddim_gemm_gnf_supported.
 * Channel counts (a*_c): multiples of 64 run on the tcgen05 kernel; other multiples of 16 run on the CUDA-core kernel
 * whatever `impl` says (linear epilogue only).  The network itself never needs that: its 32- / 96-channel layers are
 * planned pixel-paired. */
typedef struct {
  const void* a0; int a0_ctot, a0_coff, a0_c, a0_taps;
  const void* a1; int a1_ctot, a1_coff, a1_c, a1_taps;      /* a1 = NULL: single segment */
  int B, H, W;
  const void* w; int N, w_ld, w_koff;
  long long w_batch_stride; int w_rows_per_batch;           /* per-image B operand (attention); 0 = shared */
  const float* bias; const float* bias2; const float* residual; const float* rowscale;
  float scale;
  float* out32; void* out16; float* row_out; int ldo;
  int epi;                                                  /* 0 linear, 1 softmax, 2 linear + GroupNorm(+swish) of the output */
  int impl;                                                 /* 0 tcgen05, 1 CUDA-core reference */
  int force_block_n;                                        /* 0 = heuristic */
  int force_m_sub;                                          /* with force_block_n: 2 = 256-row CTA tiles */
  int n_store;                                              /* 0 = all N; else store only columns < n_store (fp32) */
  int force_cta_pairs;                                      /* 0 = heuristic, 1 = single-CTA MMA, 2 = cta_group::2 pairs (block_n 256) */
  int reverse;                                              /* 1 = tiles in descending order (L2 reuse along producer -> consumer chains); same results */
  const float* gn_gamma; const float* gn_beta;              /* epi = 2 */
  float gn_eps; int gn_groups; int gn_silu;
  int wsplit;                                               /* 0 / 1: plain; 2: w rows hold W_hi then W_lo (K columns apart, W = W_hi + W_lo): two K passes */
} gddim_gemm_desc;
int gddim_conv_gemm(const gddim_gemm_desc* d, void* stream);
int gddim_gemm_gnf_supported(int H, int W, int N, int groups);   /* 1 if epi = 2 is available for this output geometry */

/* resample: 0 none, 1 FIR down, 2 FIR up, 3 naive (mean) down, 4 naive (repeat) up.  dst16 = act(GN(x)) resampled,
 * raw16 = x resampled (either may be NULL).  Sources are fp32 NHWC, channel-concatenated (src2 optional). */
typedef struct {
  const float* src1; int c1; const float* src2; int c2;
  int B, H, W; int groups;
  const float* gamma; const float* beta; float eps;
  int silu; int resample;
  void* dst16; void* raw16;
  float raw_scale;            /* raw16 = x * raw_scale */
  int reverse;                /* 1 = apply pass walks the batch in descending order; same results */
} gddim_norm_desc;
int gddim_group_norm(const gddim_norm_desc* d, void* stream);

/* Single-head self-attention of AttnBlockpp (cld_jax/models/layerspp.py:74-78: the two einsums around the softmax):
 * out16[b,t,:] = softmax_s(q[b,t,:] . k[b,s,:] * scale) v[b,s,:], with q, k, v the channel thirds of qkv16 [B,T,3C]
 * (fp16, device).  T = 256, C = 256: one fused tcgen05 kernel (scores never leave the SM); T <= 64: CUDA-core
 * kernel; other shapes return an error (the network composes them from gddim_conv_gemm calls). */
int gddim_attention(const void* qkv16_dev, void* out16_dev, int B, int T, int C, float scale, int reverse, void* stream);
/* The same plus the block's output projection and residual (NIN_3 and `x + h`, layerspp.py:79-83) in one kernel
 * (T = 256, C = 256 only): out32 = (attention(qkv) @ w3^T + residual) * out_scale + bias3 * out_scale, and the
 * per-32-row column statistics of out32 (colstats [B*T/32][2][C]: sums, sums of squares) for the next GroupNorm.
 * w3: fp16 [C_out][C_in] (K-major), bias3 [C], residual / out32 fp32 [B,T,C]; all device pointers. */
int gddim_attention_proj(const void* qkv16_dev, const void* w3_16_dev, const float* bias3_dev, const float* residual_dev,
                         float* out32_dev, float* colstats_dev, int B, int T, int C, float scale, float out_scale,
                         int reverse, void* stream);

/* GroupNorm (no activation) fused into a following pointwise projection, the head of AttnBlockpp (layerspp.py:69-72):
 * out16 = fp16(GroupNorm(x)) @ w^T + bias, x fp32 [B,T,C], w fp16 [N][C] (K-major), out16 fp16 [B,T,N].
 * C = 256, N = 768, T a multiple of 128; the normalised tensor never reaches HBM. */
int gddim_gn_qkv(const float* x_dev, const float* gamma_dev, const float* beta_dev, int groups, float eps,
                 const void* w16_dev, const float* bias_dev, void* out16_dev, int B, int T, int C, int N, int reverse,
                 void* stream);

/* ---- samplers ----
 * kind 0: CLD deis (sampling.py:204-253 _impl_deis_sampler / get_deis_sampler)
 * kind 1: CLD order0 (sampling.py:156-202 get_order0_sampler, is_em = 0)
 * kind 2: blur order0 (blur_jax/sampling.py:53-90) */
#define GDDIM_CLD_DEIS 0
#define GDDIM_CLD_ORDER0 1
#define GDDIM_BLUR_ORDER0 2
#define GDDIM_CLD_PROGRAM 4 /* explicit step list (gddim_sampler_create_program) */
#define GDDIM_CLD_SDEIS 3   /* stochastic gDDIM: sampling.py:380-427 _impl_sdeis_sampler on sde_lib.py:334-466 LambdaSDE */
typedef struct {
  int kind;
  int nfe;
  int deis_order;
  int ts_order;
  int denoising;     /* config.sampling.noise_removal */
  int mixed_score;   /* config.model.mixed_score (CLD) */
  int use_graph;     /* CUDA graphs: deterministic samplers capture the WHOLE sample call (second call at a batch size onwards),
                        calls with per-step noise / traces one graph per network evaluation */
  float x_mul, x_add; /* inverse_scaler as an affine map: x_out = x * x_mul + x_add  ((x+1)/2 -> 0.5, 0.5) */
  float lambda_coef;  /* sdeis: config.sampling.lambda_coef */
  int sdeis_use_order0; /* sdeis: config.sampling.sdeis_use_order0 */
  unsigned long long seed; /* sdeis: Philox key for the injected noise when no explicit noise is given */
} gddim_sampler_cfg;

/* exactly one of cld / blur is used, according to kind */
int gddim_sampler_create(gddim_ctx* ctx, const gddim_sampler_cfg* cfg, const gddim_cld* cld, const gddim_blur* blur,
                         gddim_sampler** out);
/* same with an explicit time grid rev_ts[num_step+1] (cld_jax/sampling.py:204 _impl_deis_sampler, :255 hybdeis) */
int gddim_sampler_create_ts(gddim_ctx* ctx, const gddim_sampler_cfg* cfg, const gddim_cld* cld, const gddim_blur* blur,
                            const double* rev_ts, int n_ts, gddim_sampler** out);
/* A sampler as an explicit list of affine steps on the (x, v) pairs -- the form every CLD sampler of
 * cld_jax/sampling.py reduces to:  u <- A u + sum_j C_j eps_j + F z,  z ~ N(0, I_2).
 * eps_0 is this step's network evaluation at diffusion time t (t < 0: no evaluation; then n_eps must be 0 or refer
 * only to earlier evaluations through `first_eps` = 1), eps_j the j-th most recent earlier evaluation.
 * Used for 'ldeis' (sampling.py:497-540), 'em' (624-669), 'sscs' (542-622); tables come from the gddim_cld_* calls. */
typedef struct {
  double t;          /* time of the network evaluation made by this step, or < 0 for none */
  int n_eps;         /* number of eps terms (<= 6) */
  int first_eps;     /* 0: eps_0 is this step's evaluation; 1: terms start at the most recent earlier evaluation */
  float A[4];
  float C[6][4];
  float F[4];        /* noise factor, all zero = deterministic step */
  float M[4];        /* mixed score: eps_0 += M u (applied when cfg.mixed_score) */
  int trace;         /* 1: the state after this step is written to trace_dev (in order) */
  int has_P;         /* 1: the network is evaluated on P u instead of u (rotating-frame samplers: mldeis) */
  float P[4];
} gddim_step;
int gddim_sampler_create_program(gddim_ctx* ctx, const gddim_sampler_cfg* cfg, const gddim_step* steps, int n_steps,
                                 int history, gddim_sampler** out);
/* LSDE(sde).get_deis_coef(order, rev_ts) (sde_lib.py:469-519: Cholesky L_t in place of R_t): out [n_ts-1, order+3, 2, 2] */
int gddim_cld_ldeis_coef(const gddim_cld* cld, int order, const double* rev_ts, int n_ts, double* out);
/* MLCLD(sde).get_deis_coef(order, rev_ts) (sampling.py:286-325): out [n_ts-1, order+3, 2, 2]; psi1: expm(int_0^t F_1)
 * (inverse = 1: expm(int_t^0 F_1)), sde_lib.py:120-156 */
int gddim_cld_mldeis_coef(const gddim_cld* cld, int order, const double* rev_ts, int n_ts, double* out);
int gddim_cld_psi1(const gddim_cld* cld, double t, int inverse, double* out /*[2,2]*/);
/* Lifetime: a sampler is built on one gddim_ctx (weights, workspace and CUDA graphs are baked in).  gddim_ctx_destroy
 * releases the device state of every sampler still attached to it and orphans them: an orphaned handle stays valid for
 * gddim_sampler_destroy / gddim_sampler_alive, every gddim_sample* call on it fails with an error.  Destroy order is
 * therefore free (the reference has no such notion: JAX closures keep their arrays alive). */
void gddim_sampler_destroy(gddim_sampler* s);
int gddim_sampler_alive(const gddim_sampler* s);   /* 1 while the context the sampler was built on exists */
/* Philox key of the internally drawn noise (sdeis / em / sscs) for the following gddim_sample* calls -- the role of the
 * `rng` argument of the reference's sampler(rng, state, ...) (cld_jax/sampling.py:380-427).  No rebuild, graphs stay. */
int gddim_sampler_set_seed(gddim_sampler* s, unsigned long long seed);
/* The table the sampler steps through (for index-exact parity checks): fp32 [n_steps, order+3, 2, 2] for CLD
 * deis.  Returns the number of floats written (or needed when out == NULL). */
long long gddim_sampler_coef(const gddim_sampler* s, float* out, long long cap);
int gddim_sampler_num_steps(const gddim_sampler* s);
int gddim_sampler_rev_ts(const gddim_sampler* s, double* out, int cap);
/* One call = the reference's sampler(rng, state, u) body.
 *  CLD:  u [batch,S,S,C,2] reference layout -> x [batch,S,S,C] (after the affine inverse scaler),
 *        v [batch,S,S,C];  blur: u = y [batch,32,32,C] -> x; v may be NULL.
 *  host_buffers = 1: u/x/v are host pointers (pinned or pageable); copies are issued on `stream` and the
 *  call returns after synchronising it.  host_buffers = 0: device pointers, asynchronous on `stream`.
 *  trace_dev (optional, device): receives the state after every multistep update, [n_steps, batch, ...]
 *  in reference layout. */
int gddim_sample(gddim_sampler* s, const float* u, float* x, float* v, int batch, int host_buffers, float* trace_dev,
                 void* stream);
/* gddim_sample with explicit standard normals for the stochastic sampler: noise_dev (device) [n_steps, batch, S, S,
 * C, 2] in reference layout, or NULL for the internal Philox4x32-10 stream (key = seed, counter = (element, step)):
 * like a jax PRNGKey, the same seed reproduces the same noise. */
int gddim_sample_noise(gddim_sampler* s, const float* u, float* x, float* v, int batch, int host_buffers,
                       float* trace_dev, const float* noise_dev, void* stream);
/* kernels launched (or replayed through CUDA graphs) by this sampler so far */
long long gddim_sampler_launch_count(const gddim_sampler* s);
/* measurement hook: the sampler's per-step update KERNEL (deis.multistep_ab_step at full order / the blur DCT-update-IDCT
 * step; deterministic CLD calls apply the same update inside the head convolution's epilogue instead) launched `iters` times on the sampler's own buffers; average launch time (CUDA events on `stream`) and the
 * algorithmic bytes of one launch, (order+3) state arrays for CLD, 4 for blur */
int gddim_sampler_time_update(gddim_sampler* s, int batch, int iters, double* ms_per_launch, double* bytes_per_launch,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GDDIM_B200_H */

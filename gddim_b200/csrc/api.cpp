// C ABI (include/gddim_b200.h): contexts, host tables, update operators and the samplers.
//
// Sampler loops restate (index-exact) cld_jax/sampling.py:204-253 (_impl_deis_sampler, get_deis_sampler),
// :156-202 (get_order0_sampler), :30-39 (denoising step) and blur_jax/sampling.py:53-90.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/gddim_b200.h"
#include "kernels.h"
#include "tables.h"
#include "unet.h"

using namespace gddim;

static thread_local std::string g_err;
static int set_err(const std::string& m) { g_err = m; return -1; }

struct gddim_sampler;
struct gddim_ctx {
  int device;
  std::unique_ptr<UNet> net;
  std::vector<gddim_sampler*> samplers;   // samplers built on this context: orphaned (not freed) by gddim_ctx_destroy
};
struct gddim_cld { std::unique_ptr<CldTables> t; };
struct gddim_blur { std::unique_ptr<BlurTables> t; };

struct gddim_sampler {
  gddim_ctx* ctx = nullptr;
  gddim_sampler_cfg cfg;
  int n_steps = 0, order = 0, C = 0, S = 0;
  bool is_blur = false;
  std::vector<double> rev_ts;
  std::vector<float> coef;            // CLD: [n_steps][per][4], per = order+3 (deis/order0) or order+4 (sdeis)
  int per = 0;
  std::vector<float> nfac;            // sdeis: [n_steps][4] factor applied to the standard normals
  int capacity = 0;                   // images the sampler's own buffers were sized for (= the context's max_batch then)
  std::vector<gddim_step> program;    // GDDIM_CLD_PROGRAM: explicit step list
  int history = 1;
  float den_A[4], den_C[4];
  std::vector<float> mixm;            // [n_steps + 1][4]  R(t)^-1 [[0,0],[0,1]] when mixed_score
  float* d_temb_all = nullptr;        // [n_steps (+1 denoise)][temb_total]
  float* d_u = nullptr;               // state, net layout (CLD) / y (blur)
  std::vector<float*> d_eps;          // ring of network outputs
  float* d_xin = nullptr;             // blur: network input x = IDCT(y)
  float* d_stage = nullptr;           // host-buffer staging / reference-layout scratch
  float *d_x = nullptr, *d_v = nullptr;
  float *d_blur_a = nullptr, *d_blur_b = nullptr;
  std::vector<cudaGraphExec_t> graphs;   // per ring slot (trace / explicit-noise / stochastic calls)
  int graph_batch = 0;
  // deterministic samplers: ONE graph per sample call -- every network evaluation, time-embedding copy, update kernel,
  // the layout changes at both ends -- over the sampler's own buffers (inputs staged in d_stage, results in d_x / d_v)
  cudaGraphExec_t whole_graph = nullptr;
  int whole_batch = 0;                   // batch the graph was captured for
  int warm_batch = 0;                    // batch that has run once eagerly (lazy initialisations done: capturable)
  long long launches_per_sample = 0;
  bool in_whole = false;                 // inside the body of a whole-sample run: no per-slot graphs, no re-entry
  long long kernels_per_forward = 0;
  long long launches = 0;
  cudaStream_t own_stream = nullptr;   // used when the caller passes the legacy default stream (not capturable)
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
};

extern "C" {

const char* gddim_last_error(void) { return g_err.c_str(); }
int gddim_abi_version(void) { return GDDIM_ABI_VERSION; }
int gddim_cuda_available(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n > 0 ? 1 : 0;
}

// ---- context ------------------------------------------------------------------------------------------------
int gddim_ctx_create(int device, const gddim_model_cfg* cfg, int max_batch, gddim_ctx** out) {
  return gddim_ctx_create_ex(device, cfg, max_batch, 0, out);
}
int gddim_ctx_create_ex(int device, const gddim_model_cfg* cfg, int max_batch, unsigned flags, gddim_ctx** out) {
  if (!cfg || !out || max_batch < 1) return set_err("gddim_ctx_create: bad arguments");
  std::unique_ptr<gddim_ctx> c(new gddim_ctx);
  c->device = device;
  c->net.reset(new UNet(*cfg, max_batch, (flags & GDDIM_CTX_PRECISE_WEIGHTS) != 0));
  if (!c->net->error().empty()) return set_err("gddim_ctx_create: " + c->net->error());
  *out = c.release();
  return 0;
}
static void free_sampler_buffers(gddim_sampler* s);
static void orphan_sampler(gddim_sampler* s);
void gddim_ctx_destroy(gddim_ctx* ctx) {
  if (!ctx) return;
  if (ctx->net && ctx->net->finalized()) cudaSetDevice(ctx->device);
  // samplers hold a raw ctx*, buffers sized for this context's max_batch and CUDA graphs with its weight / workspace
  // pointers baked in: release their device state now and orphan them, so that a stale handle fails loudly in
  // gddim_sample* instead of replaying freed memory (the handle itself stays valid until gddim_sampler_destroy)
  for (gddim_sampler* s : ctx->samplers) orphan_sampler(s);
  delete ctx;
}
int gddim_param_count(const gddim_ctx* ctx) { return ctx ? (int)ctx->net->specs().size() : -1; }
int gddim_param_spec(const gddim_ctx* ctx, int index, char* name_buf, int name_buf_len, int shape[4], int* ndim,
                     int* kind, float* scale) {
  if (!ctx || index < 0 || index >= (int)ctx->net->specs().size()) return set_err("gddim_param_spec: bad index");
  const ParamSpec& s = ctx->net->specs()[index];
  if (name_buf && name_buf_len > 0) {
    strncpy(name_buf, s.name.c_str(), name_buf_len - 1);
    name_buf[name_buf_len - 1] = 0;
  }
  if (shape) for (int i = 0; i < 4; ++i) shape[i] = i < (int)s.shape.size() ? s.shape[i] : 1;
  if (ndim) *ndim = (int)s.shape.size();
  if (kind) *kind = s.kind;
  if (scale) *scale = s.scale;
  return 0;
}
int gddim_param_set(gddim_ctx* ctx, const char* name, const float* host_data, size_t n_elem) {
  if (!ctx || !name || !host_data) return set_err("gddim_param_set: bad arguments");
  if (ctx->net->finalized()) return set_err("gddim_param_set: context already finalized");
  if (ctx->net->set_param(name, host_data, n_elem)) return set_err(ctx->net->error());
  return 0;
}
int gddim_ctx_finalize(gddim_ctx* ctx) {
  if (!ctx) return set_err("gddim_ctx_finalize: null ctx");
  if (!gddim_cuda_available()) return set_err("gddim_ctx_finalize: no CUDA device available (the hot path has no CPU fallback)");
  if (cudaSetDevice(ctx->device) != cudaSuccess) return set_err("cudaSetDevice failed");
  if (ctx->net->finalize()) return set_err(ctx->net->error());
  return 0;
}
int gddim_ctx_set_gemm_impl(gddim_ctx* ctx, int impl) {
  if (!ctx || (impl != 0 && impl != 1)) return set_err("gddim_ctx_set_gemm_impl: bad arguments");
  ctx->net->gemm_impl = impl;
  return 0;
}
size_t gddim_ctx_workspace_bytes(const gddim_ctx* ctx) { return ctx ? ctx->net->workspace_bytes() : 0; }
int gddim_ctx_plan_size(const gddim_ctx* ctx) { return ctx ? ctx->net->num_ops() : -1; }
int gddim_ctx_plan_op(const gddim_ctx* ctx, int index, char* tag_buf, int tag_buf_len, int* kind) {
  if (!ctx || index < 0 || index >= ctx->net->num_ops()) return set_err("gddim_ctx_plan_op: bad index");
  const Op& op = ctx->net->op_at(index);
  if (tag_buf && tag_buf_len > 0) {
    strncpy(tag_buf, op.tag.c_str(), tag_buf_len - 1);
    tag_buf[tag_buf_len - 1] = 0;
  }
  if (kind) *kind = (int)op.kind;
  return 0;
}
long long gddim_ctx_launch_count(const gddim_ctx* ctx) { return ctx ? ctx->net->launch_count() : -1; }

int gddim_ctx_set_profile(gddim_ctx* ctx, int on) {
  if (!ctx) return set_err("gddim_ctx_set_profile: null ctx");
  ctx->net->set_profile(on != 0);
  return 0;
}
int gddim_ctx_get_profile(const gddim_ctx* ctx, double* ms_by_kind, double* gemm_flops, long long* gemm_launches) {
  if (!ctx || !ms_by_kind) return set_err("gddim_ctx_get_profile: bad arguments");
  ctx->net->get_profile(ms_by_kind, gemm_flops, gemm_launches);
  return 0;
}
int gddim_ctx_get_profile_hbm(const gddim_ctx* ctx, double* norm_bytes) {
  if (!ctx || !norm_bytes) return set_err("gddim_ctx_get_profile_hbm: bad arguments");
  *norm_bytes = ctx->net->profile_norm_bytes();
  return 0;
}
int gddim_ctx_dump_profile(const gddim_ctx* ctx, const char* path) {
  if (!ctx || !path) return set_err("gddim_ctx_dump_profile: bad arguments");
  if (ctx->net->dump_profile(path)) return set_err("gddim_ctx_dump_profile: cannot write file");
  return 0;
}

int gddim_unet_forward(gddim_ctx* ctx, const float* x_dev, float t, float* out_dev, int batch, void* stream) {
  if (!ctx || !x_dev || !out_dev) return set_err("gddim_unet_forward: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (ctx->net->time_projections((double)t, ctx->net->temb_cur(), st)) return set_err(ctx->net->error());
  if (ctx->net->forward(x_dev, out_dev, batch, st)) return set_err(ctx->net->error());
  return 0;
}

// ---- CLD tables ---------------------------------------------------------------------------------------------
int gddim_cld_create(double m_inv, double beta_0, double beta_1, double vv_gamma, double numerical_eps, double R_dt,
                     int is_R_rk, gddim_cld** out) {
  if (!out || !(R_dt > 0) || R_dt > 1e-2) return set_err("gddim_cld_create: bad arguments");
  gddim_cld* c = new gddim_cld;
  c->t.reset(new CldTables(m_inv, beta_0, beta_1, vv_gamma, numerical_eps, R_dt, is_R_rk != 0));
  *out = c;
  return 0;
}
void gddim_cld_destroy(gddim_cld* cld) { delete cld; }
static void put(double* o, const Mat2& m) { o[0] = m.a; o[1] = m.b; o[2] = m.c; o[3] = m.d; }
int gddim_cld_R(const gddim_cld* cld, const double* t, int n, double* out) {
  if (!cld || !t || !out) return set_err("gddim_cld_R: bad arguments");
  for (int i = 0; i < n; ++i) put(out + 4 * i, cld->t->R(t[i]));
  return 0;
}
int gddim_cld_psi(const gddim_cld* cld, const double* s, const double* t, int n, double* out) {
  if (!cld || !s || !t || !out) return set_err("gddim_cld_psi: bad arguments");
  for (int i = 0; i < n; ++i) put(out + 4 * i, cld->t->psi(s[i], t[i]));
  return 0;
}
int gddim_cld_F(const gddim_cld* cld, double t, double* out) {
  if (!cld || !out) return set_err("gddim_cld_F: bad arguments");
  put(out, cld->t->F(t));
  return 0;
}
int gddim_cld_G(const gddim_cld* cld, double t, double* out) {
  if (!cld || !out) return set_err("gddim_cld_G: bad arguments");
  put(out, cld->t->G(t));
  return 0;
}
int gddim_cld_eps_integrand(const gddim_cld* cld, const double* t, int n, double* out) {
  if (!cld || !t || !out) return set_err("gddim_cld_eps_integrand: bad arguments");
  for (int i = 0; i < n; ++i) put(out + 4 * i, cld->t->eps_integrand(t[i]));
  return 0;
}
int gddim_cld_deis_coef(const gddim_cld* cld, int order, const double* rev_ts, int n_ts, double* out) {
  if (!cld || !rev_ts || !out || order < 0 || order > 4 || n_ts < 2) return set_err("gddim_cld_deis_coef: bad arguments");
  if (n_ts - 1 < order) return set_err("gddim_cld_deis_coef: fewer steps than the multistep order");
  cld->t->deis_coef(order, rev_ts, n_ts, out);
  return 0;
}
int gddim_cld_order0_coef(const gddim_cld* cld, const double* rev_ts, int n_ts, double* mean_out, double* eps_out) {
  if (!cld || !rev_ts || !mean_out || !eps_out || n_ts < 2) return set_err("gddim_cld_order0_coef: bad arguments");
  cld->t->order0_coef(rev_ts, n_ts, mean_out, eps_out);
  return 0;
}
int gddim_cld_sdeis_coef(const gddim_cld* cld, double lambda_coef, int use_order0, int order, const double* rev_ts,
                         int n_ts, double* out) {
  if (!cld || !rev_ts || !out || order < 0 || order > 4 || n_ts < 2 || n_ts - 1 < order)
    return set_err("gddim_cld_sdeis_coef: bad arguments");
  LambdaTables lt(*cld->t, lambda_coef, use_order0 != 0);
  lt.deis_coef(order, rev_ts, n_ts, out);
  return 0;
}
int gddim_mvn_factor_svd(const double* cov, double* out) {
  if (!cov || !out) return set_err("gddim_mvn_factor_svd: bad arguments");
  put(out, mvn_factor_svd(Mat2{cov[0], cov[1], cov[2], cov[3]}));
  return 0;
}
int gddim_cld_ldeis_coef(const gddim_cld* cld, int order, const double* rev_ts, int n_ts, double* out) {
  if (!cld || !rev_ts || !out || order < 0 || order > 4 || n_ts < 2 || n_ts - 1 < order)
    return set_err("gddim_cld_ldeis_coef: bad arguments");
  cld->t->ldeis_coef(order, rev_ts, n_ts, out);
  return 0;
}
int gddim_cld_mldeis_coef(const gddim_cld* cld, int order, const double* rev_ts, int n_ts, double* out) {
  if (!cld || !rev_ts || !out || order < 0 || order > 4 || n_ts < 2 || n_ts - 1 < order)
    return set_err("gddim_cld_mldeis_coef: bad arguments");
  if (cld->t->beta_1 != 0.0) return set_err("gddim_cld_mldeis_coef: MLCLD requires beta_1 == 0 (sampling.py:289)");
  cld->t->mldeis_coef(order, rev_ts, n_ts, out);
  return 0;
}
int gddim_cld_psi1(const gddim_cld* cld, double t, int inverse, double* out) {
  if (!cld || !out) return set_err("gddim_cld_psi1: bad arguments");
  put(out, inverse ? cld->t->f1_psi(t, 0.0) : cld->t->f1_psi(0.0, t));
  return 0;
}
int gddim_rev_ts(double T, double eps, int ts_order, int num_step, double* out) {
  if (!out || num_step < 1 || ts_order < 1) return set_err("gddim_rev_ts: bad arguments");
  rev_timesteps(T, eps, ts_order, num_step, out);
  return 0;
}

// ---- blur tables --------------------------------------------------------------------------------------------
int gddim_blur_create(double sigma_blur_max, double sampling_eps, gddim_blur** out) {
  if (!out) return set_err("gddim_blur_create: bad arguments");
  gddim_blur* b = new gddim_blur;
  b->t.reset(new BlurTables(sigma_blur_max, sampling_eps));
  *out = b;
  return 0;
}
void gddim_blur_destroy(gddim_blur* b) { delete b; }
double gddim_blur_sampling_T(const gddim_blur* b) { return b->t->sampling_T(); }
int gddim_blur_y_mean_coef(const gddim_blur* b, double t, double* out) {
  if (!b || !out) return set_err("gddim_blur_y_mean_coef: bad arguments");
  b->t->y_mean_coef(t, out);
  return 0;
}
double gddim_blur_y_std_coef(const gddim_blur* b, double t) { return b->t->y_std_coef(t); }
double gddim_blur_t2alpha(const gddim_blur* b, double t) { return b->t->t2alpha(t); }

// ---- update operators -----------------------------------------------------------------------------------------
static int need_cuda(const char* who) {
  if (!gddim_cuda_available()) return set_err(std::string(who) + ": no CUDA device available (no CPU fallback)");
  return 0;
}

int gddim_multistep_ab_step(const float* x_dev, const float* deis_coef_host, const float* new_eps_dev,
                            const float* eps_pred_dev, float* x_out_dev, float* eps_pred_out_dev, int order,
                            long long n_pairs, void* stream) {
  if (need_cuda("gddim_multistep_ab_step")) return -1;
  if (!x_dev || !deis_coef_host || !new_eps_dev || !eps_pred_dev || !x_out_dev || !eps_pred_out_dev || order < 0 ||
      order > 5)
    return set_err("gddim_multistep_ab_step: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  float* d_coef = nullptr;
  const size_t nb = (size_t)(order + 3) * 4 * sizeof(float);
  if (cudaMallocAsync(&d_coef, nb, st) != cudaSuccess) return set_err("cudaMallocAsync failed");
  cudaMemcpyAsync(d_coef, deis_coef_host, nb, cudaMemcpyHostToDevice, st);
  int rc = ab_step_ref_layout_launch(x_dev, d_coef, new_eps_dev, eps_pred_dev, x_out_dev, eps_pred_out_dev, order,
                                     n_pairs, st);
  cudaFreeAsync(d_coef, st);
  if (rc) return set_err("gddim_multistep_ab_step: launch failed");
  return 0;
}

int gddim_scalar_ab_step(const float* x_dev, const float* ei_coef_host, const float* new_eps_dev,
                         const float* eps_pred_dev, float* x_out_dev, float* eps_pred_out_dev, int n_hist, long long n,
                         void* stream) {
  if (need_cuda("gddim_scalar_ab_step")) return -1;
  if (!x_dev || !ei_coef_host || !new_eps_dev || !x_out_dev || n_hist < 0 || n_hist > 8)
    return set_err("gddim_scalar_ab_step: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  float* d_coef = nullptr;
  const size_t nb = (size_t)(n_hist + 2) * sizeof(float);
  if (cudaMallocAsync(&d_coef, nb, st) != cudaSuccess) return set_err("cudaMallocAsync failed");
  cudaMemcpyAsync(d_coef, ei_coef_host, nb, cudaMemcpyHostToDevice, st);
  int rc = scalar_ab_step_launch(x_dev, d_coef, new_eps_dev, eps_pred_dev, x_out_dev, eps_pred_out_dev, n_hist, n, st);
  cudaFreeAsync(d_coef, st);
  if (rc) return set_err("gddim_scalar_ab_step: launch failed");
  return 0;
}

int gddim_relayout(const float* src_dev, float* dst_dev, long long n_pix, int C, int to_net, void* stream) {
  if (need_cuda("gddim_relayout")) return -1;
  if (relayout_launch(src_dev, dst_dev, n_pix, C, to_net, (cudaStream_t)stream)) return set_err("gddim_relayout: launch failed");
  return 0;
}

int gddim_dct2d_32(const float* in_dev, float* out_dev, int batch, int C, int forward, void* stream) {
  if (need_cuda("gddim_dct2d_32")) return -1;
  if (dct32_launch(in_dev, out_dev, batch, C, forward, (cudaStream_t)stream)) return set_err("gddim_dct2d_32: launch failed");
  return 0;
}

// ---- operator level ---------------------------------------------------------------------------------------------
int gddim_conv_gemm(const gddim_gemm_desc* d, void* stream) {
  if (need_cuda("gddim_conv_gemm")) return -1;
  if (!d || !d->a0 || !d->w) return set_err("gddim_conv_gemm: bad arguments");
  GemmOp g;
  memset(&g, 0, sizeof(g));
  g.nseg = d->a1 ? 2 : 1;
  g.seg[0] = {(const __half*)d->a0, d->a0_ctot, d->a0_coff, d->a0_c, d->a0_taps};
  if (d->a1) g.seg[1] = {(const __half*)d->a1, d->a1_ctot, d->a1_coff, d->a1_c, d->a1_taps};
  g.B = d->B; g.H = d->H; g.W = d->W;
  g.w = (const __half*)d->w; g.N = d->N; g.w_ld = d->w_ld; g.w_koff = d->w_koff;
  g.w_batch_stride = d->w_batch_stride; g.w_rows_per_batch = d->w_rows_per_batch;
  g.bias = d->bias; g.bias2 = d->bias2; g.residual = d->residual; g.rowscale = d->rowscale; g.scale = d->scale;
  g.out32 = d->out32; g.out16 = (__half*)d->out16; g.row_out = d->row_out; g.ldo = d->ldo; g.epi = d->epi; g.n_store = d->n_store;
  g.reverse = d->reverse;
  g.wsplit = d->wsplit;
  g.gn_gamma = d->gn_gamma; g.gn_beta = d->gn_beta; g.gn_eps = d->gn_eps; g.gn_groups = d->gn_groups; g.gn_silu = d->gn_silu;
  if (d->impl == 0 && gemm_prepare(&g, d->force_block_n, d->force_m_sub, d->force_cta_pairs)) return set_err(std::string("gddim_conv_gemm: ") + gemm_last_error());
  if (gemm_launch(&g, d->impl, (cudaStream_t)stream)) return set_err(std::string("gddim_conv_gemm: ") + gemm_last_error());
  return 0;
}

int gddim_gemm_gnf_supported(int H, int W, int N, int groups) { return gemm_gnf_supported(H, W, N, groups); }

int gddim_attention(const void* qkv16_dev, void* out16_dev, int B, int T, int C, float scale, int reverse, void* stream) {
  if (need_cuda("gddim_attention")) return -1;
  if (!qkv16_dev || !out16_dev || B < 1) return set_err("gddim_attention: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (attn_fused_supported(T, C)) {
    AttnOp a;
    memset(&a, 0, sizeof(a));
    a.qkv = (const __half*)qkv16_dev; a.out16 = (__half*)out16_dev; a.B = B; a.T = T; a.C = C; a.scale = scale;
    a.reverse = reverse;
    if (attn_fused_prepare(&a)) return set_err(std::string("gddim_attention: ") + gemm_last_error());
    if (attn_fused_launch(&a, B, st)) return set_err("gddim_attention: launch failed");
    return 0;
  }
  if (T <= 64) {
    if (small_attn_launch((const __half*)qkv16_dev, (__half*)out16_dev, B, T, C, scale, st)) return set_err("gddim_attention: launch failed");
    return 0;
  }
  return set_err("gddim_attention: unsupported shape (T = 256 with C = 256, or T <= 64)");
}

int gddim_attention_proj(const void* qkv16_dev, const void* w3_16_dev, const float* bias3_dev, const float* residual_dev,
                         float* out32_dev, float* colstats_dev, int B, int T, int C, float scale, float out_scale,
                         int reverse, void* stream) {
  if (need_cuda("gddim_attention_proj")) return -1;
  if (!qkv16_dev || !w3_16_dev || !bias3_dev || !residual_dev || !out32_dev || !colstats_dev || B < 1)
    return set_err("gddim_attention_proj: bad arguments");
  if (!attn_fused_supported(T, C)) return set_err("gddim_attention_proj: unsupported shape (T = 256 with C = 256)");
  AttnOp a;
  memset(&a, 0, sizeof(a));
  a.qkv = (const __half*)qkv16_dev; a.B = B; a.T = T; a.C = C; a.scale = scale; a.reverse = reverse;
  a.w3 = (const __half*)w3_16_dev; a.bias3 = bias3_dev; a.residual = residual_dev; a.out32 = out32_dev;
  a.colstats = colstats_dev; a.out_scale = out_scale;
  if (attn_fused_prepare(&a)) return set_err(std::string("gddim_attention_proj: ") + gemm_last_error());
  if (attn_fused_launch(&a, B, (cudaStream_t)stream)) return set_err("gddim_attention_proj: launch failed");
  return 0;
}

int gddim_gn_qkv(const float* x_dev, const float* gamma_dev, const float* beta_dev, int groups, float eps,
                 const void* w16_dev, const float* bias_dev, void* out16_dev, int B, int T, int C, int N, int reverse,
                 void* stream) {
  if (need_cuda("gddim_gn_qkv")) return -1;
  if (!x_dev || !gamma_dev || !beta_dev || !w16_dev || !bias_dev || !out16_dev || B < 1 || groups < 1)
    return set_err("gddim_gn_qkv: bad arguments");
  if (!gn_qkv_supported(T, C, N)) return set_err("gddim_gn_qkv: unsupported shape (C = 256, N = 768, T % 128 == 0)");
  cudaStream_t st = (cudaStream_t)stream;
  NormOp n;
  memset(&n, 0, sizeof(n));
  n.src1 = x_dev; n.c1 = C; n.B = B; n.H = T; n.W = 1; n.groups = groups; n.gamma = gamma_dev; n.beta = beta_dev; n.eps = eps;
  n.coef_only = 1;
  n.splits = norm_splits(B, T, 1);
  char* scratch = nullptr;
  const size_t nb_part = (size_t)B * n.splits * groups * 2 * sizeof(float);
  const size_t nb_coef = (size_t)B * 2 * C * sizeof(float);
  const size_t nb_tick = (size_t)B * sizeof(unsigned int);
  if (cudaMallocAsync(&scratch, nb_part + nb_coef + nb_tick, st) != cudaSuccess) return set_err("gddim_gn_qkv: scratch allocation failed");
  n.partial = (float*)scratch;
  n.coef = (float*)(scratch + nb_part);
  n.ticket = (unsigned int*)(scratch + nb_part + nb_coef);
  cudaMemsetAsync(n.ticket, 0, nb_tick, st);
  int rc = norm_launch(&n, st);
  GnQkvOp q;
  memset(&q, 0, sizeof(q));
  q.x = x_dev; q.coef = n.coef; q.w = (const __half*)w16_dev; q.bias = bias_dev; q.out16 = (__half*)out16_dev;
  q.B = B; q.T = T; q.C = C; q.N = N; q.reverse = reverse;
  if (!rc) rc = gn_qkv_prepare(&q) ? -100 : 0;
  if (!rc) rc = gn_qkv_launch(&q, B, st) ? -101 : 0;
  cudaFreeAsync(scratch, st);
  if (rc) return set_err("gddim_gn_qkv: launch failure (rc=" + std::to_string(rc) + ")");
  return 0;
}

int gddim_group_norm(const gddim_norm_desc* d, void* stream) {
  if (need_cuda("gddim_group_norm")) return -1;
  if (!d || !d->src1) return set_err("gddim_group_norm: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  NormOp n;
  memset(&n, 0, sizeof(n));
  n.src1 = d->src1; n.c1 = d->c1; n.src2 = d->src2; n.c2 = d->c2;
  n.B = d->B; n.H = d->H; n.W = d->W; n.groups = d->groups; n.gamma = d->gamma; n.beta = d->beta; n.eps = d->eps;
  n.silu = d->silu; n.resample = d->resample; n.dst16 = (__half*)d->dst16; n.raw16 = (__half*)d->raw16;
  n.raw_scale = d->raw_scale;
  n.reverse = d->reverse;
  n.splits = norm_splits(d->B, d->H, d->W);
  char* scratch = nullptr;
  const size_t nb_part = (size_t)d->B * n.splits * (d->groups > 0 ? d->groups : 1) * 2 * sizeof(float);
  const size_t nb_coef = (size_t)d->B * 2 * (d->c1 + d->c2) * sizeof(float);
  const size_t nb_tick = (size_t)d->B * sizeof(unsigned int);
  if (cudaMallocAsync(&scratch, nb_part + nb_coef + nb_tick, st) != cudaSuccess)
    return set_err("gddim_group_norm: scratch allocation failed");
  n.partial = (float*)scratch;
  n.coef = (float*)(scratch + nb_part);
  n.ticket = (unsigned int*)(scratch + nb_part + nb_coef);
  cudaMemsetAsync(n.ticket, 0, nb_tick, st);
  const int rc = norm_launch(&n, st);
  cudaFreeAsync(scratch, st);
  if (rc) return set_err("gddim_group_norm: unsupported shape or launch failure (rc=" + std::to_string(rc) + ")");
  return 0;
}

// ---- samplers ---------------------------------------------------------------------------------------------------
static void free_sampler_buffers(gddim_sampler* s) {
  for (auto g : s->graphs) if (g) cudaGraphExecDestroy(g);
  s->graphs.clear();
  if (s->whole_graph) cudaGraphExecDestroy(s->whole_graph);
  s->whole_graph = nullptr; s->whole_batch = s->warm_batch = 0;
  cudaFree(s->d_temb_all); cudaFree(s->d_u); cudaFree(s->d_xin); cudaFree(s->d_stage); cudaFree(s->d_x);
  cudaFree(s->d_v); cudaFree(s->d_blur_a); cudaFree(s->d_blur_b);
  s->d_temb_all = s->d_u = s->d_xin = s->d_stage = s->d_x = s->d_v = s->d_blur_a = s->d_blur_b = nullptr;
  for (auto p : s->d_eps) cudaFree(p);
  s->d_eps.clear();
  if (s->own_stream) cudaStreamDestroy(s->own_stream);
  if (s->ev_in) cudaEventDestroy(s->ev_in);
  if (s->ev_out) cudaEventDestroy(s->ev_out);
  s->own_stream = nullptr; s->ev_in = s->ev_out = nullptr;
}
static void orphan_sampler(gddim_sampler* s) {
  free_sampler_buffers(s);
  s->ctx = nullptr;
  s->capacity = 0;
}
static void attach_sampler(gddim_ctx* ctx, gddim_sampler* s) {
  s->capacity = ctx->net->max_batch();
  ctx->samplers.push_back(s);
}

int gddim_sampler_create(gddim_ctx* ctx, const gddim_sampler_cfg* cfg, const gddim_cld* cld, const gddim_blur* blur,
                         gddim_sampler** out) {
  return gddim_sampler_create_ts(ctx, cfg, cld, blur, nullptr, 0, out);
}

int gddim_sampler_create_ts(gddim_ctx* ctx, const gddim_sampler_cfg* cfg, const gddim_cld* cld, const gddim_blur* blur,
                            const double* rev_ts_in, int n_ts, gddim_sampler** out) {
  if (!ctx || !cfg || !out) return set_err("gddim_sampler_create: bad arguments");
  if (!ctx->net->finalized()) return set_err("gddim_sampler_create: context not finalized");
  if (cudaSetDevice(ctx->device) != cudaSuccess) return set_err("cudaSetDevice failed");
  std::unique_ptr<gddim_sampler> s(new gddim_sampler);
  s->ctx = ctx;
  s->cfg = *cfg;
  UNet& net = *ctx->net;
  s->S = net.image_size();
  s->C = net.cfg().data_channels;
  const int B = net.max_batch();
  const size_t state_elems = (size_t)B * s->S * s->S * net.net_channels();
  std::vector<double> eval_ts;       // the time of every network evaluation, in order

  if (cfg->kind == GDDIM_CLD_DEIS || cfg->kind == GDDIM_CLD_ORDER0 || cfg->kind == GDDIM_CLD_SDEIS) {
    if (!cld) return set_err("gddim_sampler_create: CLD sampler needs a gddim_cld");
    if (net.cfg().state_mult != 2) return set_err("gddim_sampler_create: CLD sampler needs a state_mult=2 network");
    const CldTables& t = *cld->t;
    const bool is_o0 = cfg->kind == GDDIM_CLD_ORDER0;
    s->order = is_o0 ? 0 : cfg->deis_order;
    if (s->order < 0 || s->order > 4) return set_err("gddim_sampler_create: deis_order must be in 0..4");
    s->n_steps = cfg->denoising ? cfg->nfe - 1 : cfg->nfe;      // sampling.py:205 / :160
    if (s->n_steps < 1 || s->n_steps < s->order) return set_err("gddim_sampler_create: nfe too small for this order");
    const int ts_order = is_o0 ? 2 : cfg->ts_order;             // sampling.py:162 hard-codes 2 for order0
    s->rev_ts.resize(s->n_steps + 1);
    if (rev_ts_in != nullptr) {
      if (is_o0 || n_ts != s->n_steps + 1) return set_err("gddim_sampler_create_ts: rev_ts must have num_step + 1 entries (deis sampler only)");
      for (int i = 0; i <= s->n_steps; ++i) s->rev_ts[i] = rev_ts_in[i];
    } else {
      rev_timesteps(t.T, t.sampling_eps, ts_order, s->n_steps, s->rev_ts.data());
    }
    const bool is_sdeis = cfg->kind == GDDIM_CLD_SDEIS;
    const int per = s->order + (is_sdeis ? 4 : 3);
    s->per = per;
    std::vector<double> c((size_t)s->n_steps * per * 4, 0.0);
    if (is_sdeis) {
      if (rev_ts_in != nullptr) return set_err("gddim_sampler_create_ts: custom time grids are supported by the deis sampler only");
      LambdaTables lt(t, (double)cfg->lambda_coef, cfg->sdeis_use_order0 != 0);
      lt.deis_coef(s->order, s->rev_ts.data(), s->n_steps + 1, c.data());
      // sampling.py:420: the last covariance is zeroed ("avoid numerical error")
      for (int k = 0; k < 4; ++k) c[((size_t)(s->n_steps - 1) * per + (per - 1)) * 4 + k] = 0.0;
      s->nfac.resize((size_t)s->n_steps * 4);
      for (int i = 0; i < s->n_steps; ++i) {
        const double* q = &c[((size_t)i * per + (per - 1)) * 4];
        const Mat2 f = mvn_factor_svd(Mat2{q[0], q[1], q[2], q[3]});
        s->nfac[i * 4 + 0] = (float)f.a; s->nfac[i * 4 + 1] = (float)f.b;
        s->nfac[i * 4 + 2] = (float)f.c; s->nfac[i * 4 + 3] = (float)f.d;
      }
    } else if (is_o0) {
      std::vector<double> mean((size_t)s->n_steps * 4), eps((size_t)s->n_steps * 4);
      t.order0_coef(s->rev_ts.data(), s->n_steps + 1, mean.data(), eps.data());
      for (int i = 0; i < s->n_steps; ++i) {
        memcpy(&c[(size_t)i * per * 4], &mean[(size_t)i * 4], 4 * sizeof(double));
        memcpy(&c[(size_t)i * per * 4 + 4], &eps[(size_t)i * 4], 4 * sizeof(double));
      }
    } else {
      t.deis_coef(s->order, s->rev_ts.data(), s->n_steps + 1, c.data());
    }
    s->coef.resize(c.size());
    for (size_t i = 0; i < c.size(); ++i) s->coef[i] = (float)c[i];
    for (int i = 0; i < s->n_steps; ++i) eval_ts.push_back(s->rev_ts[i]);
    if (cfg->denoising) {
      Mat2 A, Cm;
      t.denoise_coef(t.sampling_eps, &A, &Cm);
      s->den_A[0] = (float)A.a; s->den_A[1] = (float)A.b; s->den_A[2] = (float)A.c; s->den_A[3] = (float)A.d;
      s->den_C[0] = (float)Cm.a; s->den_C[1] = (float)Cm.b; s->den_C[2] = (float)Cm.c; s->den_C[3] = (float)Cm.d;
      eval_ts.push_back(t.sampling_eps);
    }
    s->mixm.assign(eval_ts.size() * 4, 0.f);
    if (cfg->mixed_score) {
      for (size_t i = 0; i < eval_ts.size(); ++i) {
        const Mat2 ri = inv(t.R(eval_ts[i]));      // eps += R^-1 [0, v]  (models/utils.py:174-176)
        s->mixm[i * 4 + 0] = 0.f; s->mixm[i * 4 + 1] = (float)ri.b;
        s->mixm[i * 4 + 2] = 0.f; s->mixm[i * 4 + 3] = (float)ri.d;
      }
    }
    s->d_eps.resize(s->order + 1, nullptr);
  } else if (cfg->kind == GDDIM_BLUR_ORDER0) {
    if (!blur) return set_err("gddim_sampler_create: blur sampler needs a gddim_blur");
    if (rev_ts_in != nullptr) return set_err("gddim_sampler_create_ts: custom time grids are supported by the CLD deis sampler only");
    if (net.cfg().state_mult != 1 || s->S != 32) return set_err("gddim_sampler_create: blur sampler needs a 32x32 state_mult=1 network");
    if (s->C > 5) return set_err("gddim_sampler_create: blur sampler supports up to 5 channels");
    s->is_blur = true;
    const BlurTables& t = *blur->t;
    s->n_steps = cfg->nfe;
    s->rev_ts.resize(s->n_steps + 1);
    rev_timesteps(t.sampling_T(), t.sampling_eps, cfg->ts_order, s->n_steps, s->rev_ts.data());
    std::vector<double> a((size_t)s->n_steps * 1024), b((size_t)s->n_steps * 1024);
    t.order0_coef(s->rev_ts.data(), s->n_steps + 1, a.data(), b.data());
    std::vector<float> af(a.size()), bf(b.size());
    for (size_t i = 0; i < a.size(); ++i) { af[i] = (float)a[i]; bf[i] = (float)b[i]; }
    if (cudaMalloc(&s->d_blur_a, af.size() * 4) != cudaSuccess || cudaMalloc(&s->d_blur_b, bf.size() * 4) != cudaSuccess)
      return set_err("cudaMalloc failed");
    cudaMemcpy(s->d_blur_a, af.data(), af.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(s->d_blur_b, bf.data(), bf.size() * 4, cudaMemcpyHostToDevice);
    for (int i = 0; i < s->n_steps; ++i) eval_ts.push_back(s->rev_ts[i]);
    s->d_eps.resize(1, nullptr);
    if (cudaMalloc(&s->d_xin, state_elems * 4) != cudaSuccess) return set_err("cudaMalloc failed");
  } else {
    return set_err("gddim_sampler_create: unknown sampler kind");
  }

  bool ok = cudaMalloc(&s->d_u, state_elems * 4) == cudaSuccess;
  for (auto& p : s->d_eps) ok = ok && cudaMalloc(&p, state_elems * 4) == cudaSuccess;
  ok = ok && cudaMalloc(&s->d_stage, state_elems * 4) == cudaSuccess;
  const size_t img_elems = (size_t)B * s->S * s->S * s->C;
  ok = ok && cudaMalloc(&s->d_x, img_elems * 4) == cudaSuccess;
  ok = ok && cudaMalloc(&s->d_v, img_elems * 4) == cudaSuccess;
  const int tt = net.temb_total();
  if (tt > 0) {
    ok = ok && cudaMalloc(&s->d_temb_all, eval_ts.size() * (size_t)tt * 4) == cudaSuccess;
    if (ok)
      for (size_t i = 0; i < eval_ts.size(); ++i)
        if (net.time_projections(eval_ts[i], s->d_temb_all + i * (size_t)tt, 0)) { free_sampler_buffers(s.get()); return set_err(net.error()); }
  }
  if (!ok || cudaDeviceSynchronize() != cudaSuccess) {
    free_sampler_buffers(s.get());
    return set_err(std::string("gddim_sampler_create: device allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
  }
  attach_sampler(ctx, s.get());
  *out = s.release();
  return 0;
}

int gddim_sampler_create_program(gddim_ctx* ctx, const gddim_sampler_cfg* cfg, const gddim_step* steps, int n_steps,
                                 int history, gddim_sampler** out) {
  if (!ctx || !cfg || !steps || !out || n_steps < 1 || history < 1 || history > 6)
    return set_err("gddim_sampler_create_program: bad arguments");
  if (!ctx->net->finalized()) return set_err("gddim_sampler_create_program: context not finalized");
  if (ctx->net->cfg().state_mult != 2) return set_err("gddim_sampler_create_program: needs a state_mult=2 (CLD) network");
  if (cudaSetDevice(ctx->device) != cudaSuccess) return set_err("cudaSetDevice failed");
  std::unique_ptr<gddim_sampler> s(new gddim_sampler);
  s->ctx = ctx;
  s->cfg = *cfg;
  s->cfg.kind = GDDIM_CLD_PROGRAM;
  UNet& net = *ctx->net;
  s->S = net.image_size();
  s->C = net.cfg().data_channels;
  s->program.assign(steps, steps + n_steps);
  s->history = history;
  s->n_steps = 0;
  std::vector<double> eval_ts;
  int n_evals = 0;
  for (int i = 0; i < n_steps; ++i) {
    const gddim_step& st = steps[i];
    if (st.n_eps < 0 || st.n_eps > 6) return set_err("gddim_sampler_create_program: n_eps out of range");
    if (st.t >= 0) { eval_ts.push_back(st.t); ++n_evals; }
    if (st.n_eps > 0 && !st.first_eps && st.t < 0) return set_err("gddim_sampler_create_program: eps_0 used by a step without evaluation");
    if (st.n_eps > 0) {
      const int back_max = st.n_eps - 1 + ((st.t >= 0 && st.first_eps) ? 1 : 0);   // oldest evaluation referenced
      if (back_max >= history) return set_err("gddim_sampler_create_program: eps term older than the history ring");
      if (back_max > n_evals - 1) return set_err("gddim_sampler_create_program: eps term refers to an evaluation not made yet");
    }
    if (st.trace) s->n_steps += 1;
  }
  const size_t state_elems = (size_t)net.max_batch() * s->S * s->S * net.net_channels();
  const size_t img_elems = (size_t)net.max_batch() * s->S * s->S * s->C;
  s->d_eps.resize(history, nullptr);
  bool ok = cudaMalloc(&s->d_u, state_elems * 4) == cudaSuccess;
  ok = ok && cudaMalloc(&s->d_xin, state_elems * 4) == cudaSuccess;
  for (auto& p : s->d_eps) ok = ok && cudaMalloc(&p, state_elems * 4) == cudaSuccess;
  ok = ok && cudaMalloc(&s->d_stage, state_elems * 4) == cudaSuccess;
  ok = ok && cudaMalloc(&s->d_x, img_elems * 4) == cudaSuccess && cudaMalloc(&s->d_v, img_elems * 4) == cudaSuccess;
  const int tt = net.temb_total();
  if (tt > 0 && !eval_ts.empty()) {
    ok = ok && cudaMalloc(&s->d_temb_all, eval_ts.size() * (size_t)tt * 4) == cudaSuccess;
    if (ok)
      for (size_t i = 0; i < eval_ts.size(); ++i)
        if (net.time_projections(eval_ts[i], s->d_temb_all + i * (size_t)tt, 0)) { free_sampler_buffers(s.get()); return set_err(net.error()); }
  }
  if (!ok || cudaDeviceSynchronize() != cudaSuccess) {
    free_sampler_buffers(s.get());
    return set_err("gddim_sampler_create_program: device allocation failed");
  }
  attach_sampler(ctx, s.get());
  *out = s.release();
  return 0;
}

void gddim_sampler_destroy(gddim_sampler* s) {
  if (!s) return;
  if (s->ctx) {                                  // still attached: detach from the context, then free
    cudaSetDevice(s->ctx->device);
    auto& v = s->ctx->samplers;
    v.erase(std::remove(v.begin(), v.end(), s), v.end());
    free_sampler_buffers(s);
  }
  delete s;
}

int gddim_sampler_alive(const gddim_sampler* s) { return (s && s->ctx) ? 1 : 0; }

int gddim_sampler_set_seed(gddim_sampler* s, unsigned long long seed) {
  if (!s) return set_err("gddim_sampler_set_seed: bad arguments");
  s->cfg.seed = seed;                            // read at launch time: no rebuild, graphs stay valid
  return 0;
}

long long gddim_sampler_coef(const gddim_sampler* s, float* out, long long cap) {
  if (!s) return -1;
  const long long n = (long long)s->coef.size();
  if (out && cap >= n) memcpy(out, s->coef.data(), n * sizeof(float));
  return n;
}
int gddim_sampler_num_steps(const gddim_sampler* s) { return s ? s->n_steps : -1; }
int gddim_sampler_rev_ts(const gddim_sampler* s, double* out, int cap) {
  if (!s) return -1;
  const int n = (int)s->rev_ts.size();
  if (out && cap >= n) memcpy(out, s->rev_ts.data(), n * sizeof(double));
  return n;
}

static bool whole_graph_enabled() {
  static int on = -1;                       // GDDIM_NO_WHOLE_GRAPH=1: A/B switch back to one graph per network evaluation
  if (on < 0) { const char* e = getenv("GDDIM_NO_WHOLE_GRAPH"); on = (e && e[0] == '1') ? 0 : 1; }
  return on == 1;
}
// does a call draw per-step noise (a key / explicit normals: arguments of the call, not properties of the sampler)?
static bool sampler_draws_noise(const gddim_sampler* s) {
  if (s->cfg.kind == GDDIM_CLD_SDEIS) return true;
  for (const gddim_step& ps : s->program)
    if (ps.F[0] != 0.f || ps.F[1] != 0.f || ps.F[2] != 0.f || ps.F[3] != 0.f) return true;
  return false;
}

// one network evaluation #e (time index e) reading s->d_u / d_xin and writing ring slot `slot`
// upd: the CLD update that consumes this evaluation.  Where the evaluation is launched directly (eager calls and the body of
// a whole-sample graph) it travels with the forward pass (the head convolution's epilogue applies it); per-slot graphs are
// replayed for many steps and cannot hold per-step coefficients, so there the update kernel follows the graph launch.
static int eval_net(gddim_sampler* s, int e, int slot, int batch, cudaStream_t st, const CldStepArgs* upd = nullptr) {
  UNet& net = *s->ctx->net;
  const int tt = net.temb_total();
  if (tt > 0)
    cudaMemcpyAsync(net.temb_cur(), s->d_temb_all + (size_t)e * tt, (size_t)tt * 4, cudaMemcpyDeviceToDevice, st);
  const float* in = (s->is_blur || s->cfg.kind == GDDIM_CLD_PROGRAM) ? s->d_xin : s->d_u;
  if (s->cfg.use_graph && !net.profiling() && !s->in_whole) {
    if (s->graph_batch != batch) {
      for (auto g : s->graphs) if (g) cudaGraphExecDestroy(g);
      s->graphs.assign(s->d_eps.size(), nullptr);
      s->graph_batch = batch;
    }
    if (!s->graphs[slot]) {
      // eager warm-up (sets function attributes), then capture
      if (net.forward(in, s->d_eps[slot], batch, st)) return set_err(net.error());
      cudaStreamSynchronize(st);
      const long long before = net.launch_count();
      cudaGraph_t g = nullptr;
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return set_err("cudaStreamBeginCapture failed");
      int rc = net.forward(in, s->d_eps[slot], batch, st);
      cudaError_t ce = cudaStreamEndCapture(st, &g);
      if (rc || ce != cudaSuccess) return set_err(rc ? net.error() : std::string("graph capture failed: ") + cudaGetErrorString(ce));
      s->kernels_per_forward = net.launch_count() - before;
      ce = cudaGraphInstantiate(&s->graphs[slot], g, 0);
      cudaGraphDestroy(g);
      if (ce != cudaSuccess) return set_err(std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(ce));
    }
    if (cudaGraphLaunch(s->graphs[slot], st) != cudaSuccess) return set_err("cudaGraphLaunch failed");
    s->launches += s->kernels_per_forward;
    if (upd != nullptr) {
      if (cld_step_launch(upd, st)) return set_err("cld_step launch failed");
      s->launches += 1;
    }
  } else {
    const long long before = net.launch_count();
    if (net.forward(in, s->d_eps[slot], batch, st, upd)) return set_err(net.error());
    s->launches += net.launch_count() - before;
  }
  return 0;
}

int gddim_sample(gddim_sampler* s, const float* u, float* x, float* v, int batch, int host_buffers, float* trace_dev,
                 void* stream) {
  return gddim_sample_noise(s, u, x, v, batch, host_buffers, trace_dev, nullptr, stream);
}

int gddim_sample_noise(gddim_sampler* s, const float* u, float* x, float* v, int batch, int host_buffers,
                       float* trace_dev, const float* noise_dev, void* stream) {
  if (!s || !u || !x) return set_err("gddim_sample: bad arguments");
  if (noise_dev != nullptr && s->cfg.kind != GDDIM_CLD_SDEIS && s->cfg.kind != GDDIM_CLD_PROGRAM)
    return set_err("gddim_sample_noise: explicit noise is for the stochastic samplers");
  if (!s->ctx) return set_err("gddim_sample: the network context this sampler was built on has been destroyed");
  UNet& net = *s->ctx->net;
  if (batch < 1 || batch > s->capacity || batch > net.max_batch())
    return set_err("gddim_sample: batch exceeds the capacity the sampler was created with");
  if (cudaSetDevice(s->ctx->device) != cudaSuccess) return set_err("cudaSetDevice failed");
  cudaStream_t caller = (cudaStream_t)stream;
  cudaStream_t st = caller;
  const bool legacy = (caller == nullptr || caller == cudaStreamLegacy);
  if (legacy) {
    if (!s->own_stream) {
      if (cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&s->ev_in, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&s->ev_out, cudaEventDisableTiming) != cudaSuccess)
        return set_err("gddim_sample: stream creation failed");
    }
    st = s->own_stream;
    cudaEventRecord(s->ev_in, caller);          // order after the caller's pending work ...
    cudaStreamWaitEvent(st, s->ev_in, 0);
  }
  const long long n_pix = (long long)batch * s->S * s->S;
  const size_t state_elems = (size_t)n_pix * net.net_channels();
  const size_t img_elems = (size_t)n_pix * s->C;

  // ---- deterministic samplers: the whole call as one CUDA graph ----------------------------------------------------
  // (per-step noise keys, explicit noise and traces are arguments of the call and would be baked into the graph: those
  // calls keep one graph per network evaluation with the update kernels launched in between)
  if (whole_graph_enabled() && s->cfg.use_graph && !net.profiling() && !s->in_whole && !trace_dev && !noise_dev &&
      !sampler_draws_noise(s)) {
    if (!s->is_blur && !v) return set_err("gddim_sample: CLD sampler needs a v output");
    if (cudaMemcpyAsync(s->d_stage, u, state_elems * 4, host_buffers ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return set_err("gddim_sample: input copy failed");
    int rc = 0;
    s->in_whole = true;
    if (s->whole_graph && s->whole_batch == batch) {
      if (cudaGraphLaunch(s->whole_graph, st) != cudaSuccess) rc = set_err("cudaGraphLaunch failed");
      s->launches += s->launches_per_sample;
    } else if (s->warm_batch != batch) {
      // first call at this batch: eager (function attributes, exchange buffers, the DCT matrix are set up lazily)
      if (s->whole_graph) { cudaGraphExecDestroy(s->whole_graph); s->whole_graph = nullptr; s->whole_batch = 0; }
      rc = gddim_sample_noise(s, s->d_stage, s->d_x, s->d_v, batch, 0, nullptr, nullptr, st);
      if (!rc) s->warm_batch = batch;
    } else {
      const long long before = s->launches;
      cudaGraph_t g = nullptr;
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        s->in_whole = false;
        return set_err("cudaStreamBeginCapture failed");
      }
      rc = gddim_sample_noise(s, s->d_stage, s->d_x, s->d_v, batch, 0, nullptr, nullptr, st);
      const cudaError_t ce = cudaStreamEndCapture(st, &g);
      if (!rc && ce != cudaSuccess) rc = set_err(std::string("graph capture failed: ") + cudaGetErrorString(ce));
      if (!rc) {
        s->launches_per_sample = s->launches - before;
        const cudaError_t ci = cudaGraphInstantiate(&s->whole_graph, g, 0);
        if (ci != cudaSuccess) rc = set_err(std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(ci));
        else s->whole_batch = batch;
      }
      if (g) cudaGraphDestroy(g);
      if (!rc && cudaGraphLaunch(s->whole_graph, st) != cudaSuccess) rc = set_err("cudaGraphLaunch failed");
    }
    s->in_whole = false;
    if (rc) return rc;
    const cudaMemcpyKind back = host_buffers ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    cudaMemcpyAsync(x, s->d_x, img_elems * 4, back, st);
    if (!s->is_blur) cudaMemcpyAsync(v, s->d_v, img_elems * 4, back, st);
    if (host_buffers) {
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return set_err(std::string("gddim_sample: ") + cudaGetErrorString(e));
    }
    if (legacy) {
      cudaEventRecord(s->ev_out, st);
      cudaStreamWaitEvent(caller, s->ev_out, 0);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(std::string("gddim_sample: ") + cudaGetErrorString(e));
    return 0;
  }

  if (!s->is_blur) {
    if (!v) return set_err("gddim_sample: CLD sampler needs a v output");
    const float* u_dev = u;
    if (host_buffers) {
      if (cudaMemcpyAsync(s->d_stage, u, state_elems * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) return set_err("H2D copy failed");
      u_dev = s->d_stage;
    }
    if (relayout_launch(u_dev, s->d_u, n_pix, s->C, 1, st)) return set_err("relayout failed");
    s->launches += 1;
    if (s->cfg.kind == GDDIM_CLD_PROGRAM) {
      const int ring = s->history;
      int ne = -1;            // index of the latest evaluation
      int nn = 0, ntr = 0;    // noise draws / traced states so far
      for (size_t i = 0; i < s->program.size(); ++i) {
        const gddim_step& ps = s->program[i];
        int slot = ne >= 0 ? ne % ring : 0;
        if (ps.t >= 0) {
          ++ne;
          slot = ne % ring;
          // network input: the state, or P * state for rotating-frame samplers
          CldStepArgs pin;
          memset(&pin, 0, sizeof(pin));
          pin.u = s->d_u; pin.u_out = s->d_xin; pin.n_pix = n_pix; pin.C = s->C;
          const float ident[4] = {1.f, 0.f, 0.f, 1.f};
          memcpy(pin.coef[0], ps.has_P ? ps.P : ident, 16);
          if (cld_step_launch(&pin, st)) return set_err("cld_step launch failed");
          s->launches += 1;
          if (eval_net(s, ne, slot, batch, st)) return -1;
        }
        CldStepArgs a;
        memset(&a, 0, sizeof(a));
        a.u = s->d_u; a.u_out = s->d_u;
        a.n_pix = n_pix; a.C = s->C;
        a.mixed = (s->cfg.mixed_score && ps.t >= 0) ? 1 : 0;
        a.eps_store = s->d_eps[slot];
        memcpy(a.mixm, ps.M, 16);
        memcpy(a.coef[0], ps.A, 16);
        a.n_eps = ps.n_eps;
        for (int j = 0; j < ps.n_eps; ++j) {
          memcpy(a.coef[1 + j], ps.C[j], 16);
          const int back = j + (ps.first_eps ? 1 : 0) - ((ps.t >= 0) ? 0 : (ps.first_eps ? 1 : 0));
          if (ne - back < 0) return set_err("gddim_sample: program refers to an evaluation that was not made yet");
          a.eps[j] = s->d_eps[(ne - back) % ring];
        }
        if (a.mixed && (ps.n_eps == 0 || ps.first_eps)) {
          // the mixed-score term must be folded into this evaluation even if the step does not use it directly
          return set_err("gddim_sample: mixed_score steps must consume their own evaluation as eps_0");
        }
        if (ps.F[0] != 0.f || ps.F[1] != 0.f || ps.F[2] != 0.f || ps.F[3] != 0.f) {
          memcpy(a.nfac, ps.F, 16);
          if (noise_dev != nullptr) { a.noise_mode = 1; a.noise = noise_dev + (size_t)nn * state_elems; }
          else { a.noise_mode = 2; a.seed = s->cfg.seed; a.stream_id = (unsigned long long)nn; }
          ++nn;
        }
        if (cld_step_launch(&a, st)) return set_err("cld_step launch failed");
        s->launches += 1;
        if (trace_dev && ps.trace) {
          if (relayout_launch(s->d_u, trace_dev + (size_t)ntr * state_elems, n_pix, s->C, 0, st)) return set_err("trace relayout failed");
          s->launches += 1;
          ++ntr;
        }
      }
    }
    const int ring = s->order + 1;
    const int per = s->per;
    const int n_evals = s->cfg.kind == GDDIM_CLD_PROGRAM ? 0 : s->n_steps + (s->cfg.denoising ? 1 : 0);
    for (int e = 0; e < n_evals; ++e) {
      const int slot = e % ring;
      CldStepArgs a;
      memset(&a, 0, sizeof(a));
      a.u = s->d_u; a.u_out = s->d_u;
      a.n_pix = n_pix; a.C = s->C;
      a.mixed = s->cfg.mixed_score; a.eps_store = s->d_eps[slot];
      memcpy(a.mixm, &s->mixm[(size_t)e * 4], 16);
      if (e < s->n_steps) {
        const float* c = &s->coef[(size_t)e * per * 4];
        const int r = e < s->order ? e : s->order;          // rows i < order run at order i (deis.py:75)
        a.n_eps = r + 1;
        memcpy(a.coef[0], c, 16);
        for (int j = 0; j <= r; ++j) {
          memcpy(a.coef[1 + j], c + (1 + j) * 4, 16);
          a.eps[j] = s->d_eps[((e - j) % ring + ring) % ring];
        }
        if (s->cfg.kind == GDDIM_CLD_SDEIS) {
          memcpy(a.nfac, &s->nfac[(size_t)e * 4], 16);
          if (noise_dev != nullptr) { a.noise_mode = 1; a.noise = noise_dev + (size_t)e * state_elems; }
          else { a.noise_mode = 2; a.seed = s->cfg.seed; a.stream_id = (unsigned long long)e; }   // same key -> same noise, like a jax PRNGKey
        }
      } else {
        a.n_eps = 1;
        memcpy(a.coef[0], s->den_A, 16);
        memcpy(a.coef[1], s->den_C, 16);
        a.eps[0] = s->d_eps[slot];
      }
      if (eval_net(s, e, slot, batch, st, &a)) return -1;       // network evaluation + the update that consumes it
      if (trace_dev && e < s->n_steps) {
        if (relayout_launch(s->d_u, trace_dev + (size_t)e * state_elems, n_pix, s->C, 0, st)) return set_err("trace relayout failed");
        s->launches += 1;
      }
    }
    float* xd = host_buffers ? s->d_x : x;
    float* vd = host_buffers ? s->d_v : v;
    if (cld_split_launch(s->d_u, xd, vd, n_pix, s->C, s->cfg.x_mul, s->cfg.x_add, st)) return set_err("split failed");
    s->launches += 1;
    if (host_buffers) {
      cudaMemcpyAsync(x, s->d_x, img_elems * 4, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(v, s->d_v, img_elems * 4, cudaMemcpyDeviceToHost, st);
    }
  } else {
    if (host_buffers) {
      if (cudaMemcpyAsync(s->d_u, u, state_elems * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) return set_err("H2D copy failed");
    } else {
      cudaMemcpyAsync(s->d_u, u, state_elems * 4, cudaMemcpyDeviceToDevice, st);
    }
    if (dct32_launch(s->d_u, s->d_xin, batch, s->C, 0, st)) return set_err("idct failed");
    s->launches += 1;
    for (int e = 0; e < s->n_steps; ++e) {
      if (eval_net(s, e, 0, batch, st)) return -1;
      if (blur_step_launch(s->d_u, s->d_eps[0], s->d_blur_a + (size_t)e * 1024, s->d_blur_b + (size_t)e * 1024, s->d_u,
                           s->d_xin, batch, s->C, st))
        return set_err("blur_step launch failed");
      s->launches += 1;
      if (trace_dev) cudaMemcpyAsync(trace_dev + (size_t)e * state_elems, s->d_u, state_elems * 4, cudaMemcpyDeviceToDevice, st);
    }
    float* xd = host_buffers ? s->d_x : x;
    if (scale_shift_launch(s->d_xin, xd, (long long)img_elems, s->cfg.x_mul, s->cfg.x_add, st)) return set_err("final scale failed");
    s->launches += 1;
    if (host_buffers) cudaMemcpyAsync(x, s->d_x, img_elems * 4, cudaMemcpyDeviceToHost, st);
  }
  if (host_buffers) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return set_err(std::string("gddim_sample: ") + cudaGetErrorString(e));
  }
  if (legacy) {                                   // ... and make the caller's stream wait for the results
    cudaEventRecord(s->ev_out, st);
    cudaStreamWaitEvent(caller, s->ev_out, 0);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(std::string("gddim_sample: ") + cudaGetErrorString(e));
  return 0;
}

long long gddim_sampler_launch_count(const gddim_sampler* s) { return s ? s->launches : -1; }

// Measurement hook (bench.py's HBM roofline leg): the sampler's own per-step update kernel -- cld_step at full order
// (reads u and order+1 ring slots, writes u) or blur_step (reads y and eps, writes y and the next network input) --
// launched `iters` times on the sampler's buffers, each from a flushed L2, timed one by one with CUDA events on `stream`.
int gddim_sampler_time_update(gddim_sampler* s, int batch, int iters, double* ms_per_launch, double* bytes_per_launch,
                              void* stream) {
  if (!s || !s->ctx || iters < 1 || !ms_per_launch || !bytes_per_launch) return set_err("gddim_sampler_time_update: bad arguments");
  if (batch < 1 || batch > s->capacity) return set_err("gddim_sampler_time_update: batch exceeds the sampler's capacity");
  if (s->cfg.kind == GDDIM_CLD_PROGRAM) return set_err("gddim_sampler_time_update: not available for step programs");
  if (cudaSetDevice(s->ctx->device) != cudaSuccess) return set_err("cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  UNet& net = *s->ctx->net;
  const long long n_pix = (long long)batch * s->S * s->S;
  const size_t state_bytes = (size_t)n_pix * net.net_channels() * 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = 0;
  // The launches run back to back inside ONE event pair on rotating sets of buffers whose total size is several times
  // the L2 (126 MB), so that every launch finds its operands in HBM -- as in the sampler, where a whole network
  // evaluation runs between two updates -- without the per-launch event / launch overhead in the average.
  const int arrays = s->is_blur ? 3 : s->order + 2;                 // distinct arrays a launch touches (u is updated in place)
  const size_t set_bytes = (size_t)arrays * state_bytes;
  int sets = (int)(((size_t)400 << 20) / set_bytes) + 1;
  if (sets > 64) sets = 64;
  if (sets < 2) sets = 2;
  char* scratch = nullptr;
  if (cudaMalloc(&scratch, (size_t)sets * set_bytes) != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return set_err("gddim_sampler_time_update: cudaMalloc failed"); }
  cudaMemsetAsync(scratch, 0, (size_t)sets * set_bytes, st);
  auto one = [&](int k) -> int {
    float* base = reinterpret_cast<float*>(scratch + (size_t)(k % sets) * set_bytes);
    const size_t stride = state_bytes / 4;
    if (s->is_blur)
      return blur_step_launch(base, base + stride, s->d_blur_a, s->d_blur_b, base, base + 2 * stride, batch, s->C, st);
    CldStepArgs a;
    memset(&a, 0, sizeof(a));
    a.u = base; a.u_out = base; a.n_pix = n_pix; a.C = s->C;
    a.n_eps = s->order + 1;
    const float ident[4] = {1.f, 0.f, 0.f, 1.f};
    memcpy(a.coef[0], ident, 16);
    for (int j = 0; j <= s->order; ++j) a.eps[j] = base + (size_t)(1 + j) * stride;
    return cld_step_launch(&a, st);
  };
  for (int i = 0; i < sets && !rc; ++i) rc = one(i);
  cudaEventRecord(e0, st);
  for (int i = 0; i < iters && !rc; ++i) rc = one(i);
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float total = 0.f;
  cudaEventElapsedTime(&total, e0, e1);
  cudaFree(scratch);
  const float ms = (float)total;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (rc) return set_err("gddim_sampler_time_update: launch failed");
  *ms_per_launch = ms / iters;
  // algorithmic bytes: CLD reads u + (order+1) eps, writes u' = (order+3) state arrays;
  // blur reads y and eps, writes y' and the next network input = 4 arrays
  *bytes_per_launch = s->is_blur ? 4.0 * (double)state_bytes : (double)(s->order + 3) * (double)state_bytes;
  return 0;
}

}  // extern "C"

// NCSN++ / DDPM++ score network: parameter inventory, static launch plan and executor.
// Mirrors the control flow of cld_jax/models/ncsnpp.py:41-243 (see unet.cpp for per-block citations).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/gddim_b200.h"
#include "kernels.h"

namespace gddim {

struct ParamSpec {
  std::string name;
  std::vector<int> shape;
  int kind;      // 0 variance_scaling, 1 zeros, 2 ones, 3 normal
  float scale;
};

// (OP_STEM / OP_HEAD are never planned -- stem and head run as GEMM ops; the names keep the numbering of gddim_ctx_plan_op stable)
enum OpKind { OP_STEM, OP_NORM, OP_GEMM, OP_HEAD, OP_IM2COL, OP_TRANSPOSE_V, OP_SMALL_ATTN, OP_SOFTMAX_ROWS, OP_ATTN_FUSED, OP_GN_QKV };

struct Op {
  OpKind kind;
  NormOp norm;
  GemmOp gemm;
  AttnOp attn;
  GnQkvOp gq;
  // generic fields for the small kernels
  const float* f_in = nullptr;
  const __half* h_in = nullptr;
  float* f_out = nullptr;
  __half* h_out = nullptr;
  const float* w = nullptr;
  const float* bias = nullptr;
  int H = 0, W = 0, cin = 0, cout = 0, kpad = 0, T = 0, ld = 0, voff = 0, use_fir = 0;
  float scale = 1.f;
  bool in_is_external = false;     // reads the caller's x
  bool out_is_external = false;    // writes the caller's out
  std::string tag;
};

class UNet {
 public:
  // precise: conv weights as fp16 hi + lo pairs (two K passes per convolution; parity mode, ~2x slower convolutions)
  UNet(const gddim_model_cfg& cfg, int max_batch, bool precise = false);
  ~UNet();
  const std::vector<ParamSpec>& specs() const { return specs_; }
  int set_param(const std::string& name, const float* host, size_t n);
  int finalize();
  bool finalized() const { return finalized_; }
  // time conditioning -------------------------------------------------------------------------------
  // computes the per-block time projections for diffusion time t into `dst_dev` ([temb_total] floats)
  int time_projections(double t, float* dst_dev, cudaStream_t st);
  int temb_total() const { return temb_total_; }
  float* temb_cur() const { return temb_cur_; }     // the buffer the conv epilogues read
  // forward: uses whatever temb_cur() currently holds
  // head_update (CLD samplers): the update u' = A u + sum_j C_j eps_j that consumes this evaluation (eps[0] = out_dev) runs
  // as part of the forward pass -- inside the head convolution's epilogue where that layer qualifies
  // (gemm_head_update_supported), else as the update kernel right behind it
  int forward(const float* x_dev, float* out_dev, int batch, cudaStream_t st, const CldStepArgs* head_update = nullptr);
  int net_channels() const { return cfg_.data_channels * cfg_.state_mult; }
  int image_size() const { return cfg_.image_size; }
  int max_batch() const { return max_batch_; }
  size_t workspace_bytes() const { return finalized_ ? arena_bytes_ + weight_bytes_ : arena_peak_ + weight_top_; }
  long long launch_count() const { return launches_; }
  int gemm_impl = 0;
  const std::string& error() const { return err_; }
  const gddim_model_cfg& cfg() const { return cfg_; }
  int num_ops() const { return (int)ops_.size(); }
  const Op& op_at(int i) const { return ops_[i]; }
  // per-op CUDA-event timing of forward() (eager launches only); accumulates until reset
  void set_profile(bool on);
  bool profiling() const { return profile_; }
  void get_profile(double ms_by_kind[8], double* gemm_flops, long long* gemm_launches) const;
  int dump_profile(const char* path) const;
  double profile_norm_bytes() const;      // algorithmic HBM bytes of all GroupNorm ops over the profiled forwards

 private:
  // prod_op: index in ops_ of the GEMM whose linear epilogue wrote this tensor (-1: other producers) -- lets the next
  // block attach its GroupNorm_0 to that epilogue (dual GroupNorm epilogue, resblock())
  struct T32 { float* p; int C, H, W; size_t bytes; float* stats; size_t stats_bytes; bool stats_valid; int prod_op = -1; };
  struct T16 { __half* p; int C, H, W; size_t bytes; };
  struct Scope;
  gddim_model_cfg cfg_;
  int max_batch_;
  bool dry_ = true;
  bool precise_ = false;
  bool finalized_ = false;
  std::string err_;
  std::vector<ParamSpec> specs_;
  std::map<std::string, std::vector<float>> host_params_;
  std::vector<Op> ops_;
  // buffers released since the last allocation: returned to the free list by the next ordinary allocation, but kept out of
  // it for an allocation that becomes an extra OUTPUT of the op pushed last (it must not alias that op's inputs)
  std::vector<std::pair<void*, size_t>> pending_free_;
  bool hold_pending_ = false;
  size_t last_flush_at_ = 0;            // ops_.size() when released buffers last went back to the free list
  void flush_pending();
  long long launches_ = 0;
  bool profile_ = false;
  std::vector<cudaEvent_t> prof_ev_;
  std::vector<double> prof_op_ms_;        // per op, accumulated
  std::vector<double> prof_op_flops_;     // per op per forward (GEMMs)
  std::vector<double> prof_op_bytes_;     // per op per forward (GroupNorm family: algorithmic HBM bytes)
  long long prof_forwards_ = 0;

  // workspace arena (activations) with a first-fit free list, offsets assigned during the walk
  char* arena_ = nullptr;
  size_t arena_bytes_ = 0, arena_peak_ = 0;
  std::vector<std::pair<size_t, size_t>> free_;   // (offset, size)
  size_t arena_top_ = 0;
  void* a_alloc(size_t bytes);
  void a_free(void* p, size_t bytes);

  // weight arena (bump)
  char* wts_ = nullptr;
  size_t weight_bytes_ = 0, weight_top_ = 0;
  void* w_alloc(size_t bytes);
  float* upload_f32(const std::vector<float>& v);
  __half* upload_f16(const std::vector<__half>& v);

  // time embedding
  int temb_dim_ = 0, emb_in_dim_ = 0;
  int temb_total_ = 0;
  std::vector<float> fourier_w_;
  float *d_dense0_w_ = nullptr, *d_dense0_b_ = nullptr, *d_dense1_w_ = nullptr, *d_dense1_b_ = nullptr;
  float *d_proj_w_ = nullptr, *d_proj_b_ = nullptr;     // [temb_dim, temb_total], [temb_total]
  std::vector<float> proj_w_host_, proj_b_host_;
  float *d_temb0_ = nullptr, *d_temb1_ = nullptr, *d_temb2_ = nullptr;
  float* temb_cur_ = nullptr;
  float* gn_partial_ = nullptr;
  float* gn_coef_ = nullptr;
  unsigned int* gn_ticket_ = nullptr;

  // walk
  const std::vector<float>* param(Scope& s, const std::string& name, std::vector<int> shape, int kind, float scale);
  int walk();
  T32 new32(int C, int H, int W);
  T16 new16(int C, int H, int W);
  void rel(T32& t);
  void rel(T16& t);
  T32 resblock(Scope& top, const T32& in1, const T32* in2, int out_ch, bool up, bool down);
  T32 attnblock(Scope& top, const T32& x);
  void add_norm(const T32& in1, const T32* in2, const float* gamma, const float* beta, bool silu, int resample,
                T16* dst, T16* raw, const std::string& tag);
  // copies > 1: the same projection again in the following column blocks (pixel-paired convolutions read [b | b])
  int add_temb_proj(const std::vector<float>* w, const std::vector<float>* b, const std::vector<float>* conv_b, int out_ch, int copies = 1);
  std::pair<const float*, const float*> gn_params(Scope& s, int C);
  int fail(const std::string& m) { err_ = m; return -1; }
};

}  // namespace gddim

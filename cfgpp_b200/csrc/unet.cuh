// cfgpp_b200 — UNet2DConditionModel executor: weight registry + repacking, static launch plan over the hand-written
// kernels, CUDA-graph replay of one fused sampler step. Structure follows SURVEY.md Appendix A (diffusers 0.27.1).
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/cfgpp_b200.h"
#include "attention.cuh"
#include "gemm.cuh"
#include "ops.cuh"

namespace cfgpp {

struct DevTensor {
  __half* p = nullptr;
  std::vector<int64_t> shape;
  size_t numel() const {
    size_t n = 1;
    for (auto d : shape) n *= static_cast<size_t>(d);
    return n;
  }
};

struct PlanStep {
  std::string name;
  double flops = 0.0;  // algorithmic FLOPs of this launch group (0 for non-contraction kernels)
  int launches = 1;
  int kind = 3;  // 0 linear GEMM, 1 conv3x3 (implicit GEMM), 2 attention, 3 other (norm / elementwise)
  std::function<void(cudaStream_t)> fn;
};

class Unet {
 public:
  Unet(const cfgpp_model_desc& d, int device);
  ~Unet();

  void load_weight(const std::string& key, const void* data, const int64_t* shape, int ndim, int dtype,
                   cudaStream_t stream);
  void finalize_weights(cudaStream_t stream);
  void prepare(int batch, int h_lat, int w_lat);
  size_t workspace_bytes() const { return workspace_bytes_; }
  double forward_flops() const { return forward_flops_; }
  int launches_per_step() const { return launches_per_step_; }
  double prompt_flops() const { return prompt_flops_; }
  int prompt_launches() const { return prompt_launches_; }

  void set_prompt(const __half* ctx, int n_ctx, const __half* pooled, const float* time_ids, int add_rows,
                  cudaStream_t stream);
  void unet_forward(const void* z, int z_dtype, float t, float in_scale, __half* eps_uc, __half* eps_c,
                    cudaStream_t stream);
  void set_schedule(int method, int state_dtype, const cfgpp_step_state* steps, int nsteps, cudaStream_t stream);
  void set_state(const void* z, int z_dtype, cudaStream_t stream);
  // ancestral samplers: fp16 noise table [slots][B,4,H,W] (device), copied into a handle-owned buffer
  void set_noise(const __half* noise, int slots, cudaStream_t stream);
  void run_steps(int first_step, int nsteps, cudaStream_t stream);
  void get_state(int which, void* out, cudaStream_t stream);
  void apply_step(int step, const __half* eps_uc, const __half* eps_c, cudaStream_t stream);
  // Eager un-fused forward with a CUDA-event pair around every plan entry (profiling aid for bench.py).
  struct ProfEntry {
    std::string name;
    int kind;
    double flops;
    float ms;
  };
  std::vector<ProfEntry> profile_forward(const void* z, int z_dtype, float t, float in_scale, cudaStream_t stream);

 private:
  // ---- weights ----
  const DevTensor& raw(const std::string& key) const;
  __half* alloc_weight(size_t numel);
  __half* packed_conv3x3(const std::string& key);  // (Cout,Cin,3,3) -> [Cout][9][Cin]
  __half* packed_cat_rows(const std::vector<std::string>& keys);
  __half* packed_geglu(const std::string& key, bool is_bias);
  __half* packed_heads_rows(const std::vector<std::string>& keys, int heads, int hd, int hdp);
  __half* packed_heads_cols(const std::string& key, int heads, int hd, int hdp);
  struct FoldedLN {
    __half* w;
    float* s;
    float* t;
  };
  FoldedLN folded_ln(const std::string& cache_key, const __half* w_packed, int N, int K, const std::string& norm_prefix,
                     const __half* bias_packed);
  std::map<std::string, FoldedLN> fold_cache_;
  static bool lnfold_disabled();
  __half* plain(const std::string& key);

  // ---- workspace ----
  __half* alloc_act(size_t numel);
  void* alloc_bytes(size_t bytes);
  struct Scratch {
    size_t need = 0;
    __half* p = nullptr;
  };
  // ---- plan building ----
  struct Act {
    __half* p;
    int C;
  };
  Act build_resnet(const std::string& prefix, Act x1, const Act* x2, int Cout, int H, int W, int temb_off);
  Act build_transformer(const std::string& prefix, Act x, int H, int W, int layers, int heads);
  Act build_downsample(const std::string& prefix, Act x, int H, int W);
  Act build_upsample(const std::string& prefix, Act x, int H, int W);
  void add_gemm(const std::string& name, const GemmOp& op, double algorithmic_flops = -1.0);
  void add_attn(const std::string& name, const AttnOp& op);
  void add_step(const std::string& name, std::function<void(cudaStream_t)> fn, int launches = 1);
  void run_plan(const std::vector<PlanStep>& plan, cudaStream_t stream);
  void run_body(cudaStream_t stream, bool concurrent);
  static bool split_disabled();
  void build_final(int mode_with_dtype, bool expose_eps);
  void ensure_graph(cudaStream_t stream);

  cfgpp_model_desc d_;
  int device_;
  bool finalized_ = false, prepared_ = false;
  std::map<std::string, DevTensor> raw_;
  std::vector<void*> weight_allocs_;
  std::vector<void*> act_allocs_;
  size_t workspace_bytes_ = 0, weight_bytes_ = 0;

  // packed weights: resolved lazily during plan building (finalize just validates + packs what is shape-independent)
  std::map<std::string, __half*> packed_cache_;
  __half* temb_w_all_ = nullptr;  // [sumCout][time_embed_dim]
  __half* temb_b_all_ = nullptr;
  int temb_total_ = 0;
  std::vector<std::pair<std::string, int>> temb_order_;  // resnet prefix -> offset
  int time_embed_dim_ = 0;

  // plan
  int B_ = 0, NB_ = 0, H_ = 0, W_ = 0;
  std::vector<PlanStep>* cur_plan_ = nullptr;
  std::vector<PlanStep> prologue_plan_;  // timestep embedding -> temb for all resnets
  std::vector<PlanStep> branch_plan_[2];  // conv_in output .. last up block, one plan per CFG half (uncond / cond)
  std::vector<PlanStep> tail_plan_;       // conv_norm_out + SiLU over the whole batch
  int n_branches_ = 1;
  // branch currently being built (rows [brow0_, brow0_ + bnb_) of the UNet batch)
  int bnb_ = 0, brow0_ = 0;
  std::string btag_;
  float* bgn_partial_ = nullptr;
  __half* out_override_ = nullptr;
  std::vector<PlanStep> prompt_plan_;    // cross-attention K/V projections + add-embedding
  double forward_flops_ = 0.0, prompt_flops_ = 0.0;
  int launches_per_step_ = 0, prompt_launches_ = 0;

  // scratch (sized as the max over all uses while building, allocated afterwards; closures hold Scratch*)
  std::map<std::string, std::unique_ptr<Scratch>> scratch_;
  Scratch* scratch(const std::string& name, size_t numel_half);

  // conditioning / per-step device state
  const __half* ctx_ = nullptr;  // caller-owned, valid while the prompt is set
  __half* ctx_copy_ = nullptr;
  int n_ctx_ = 77;
  __half* add_in_ = nullptr;    // [NB][proj_in_dim]
  __half* add_h1_ = nullptr;
  __half* aug_emb_ = nullptr;   // [NB][time_embed_dim]
  __half* pooled_copy_ = nullptr;
  float* time_ids_copy_ = nullptr;
  int add_rows_ = 0;
  bool has_aug_ = false;
  __half *t_sin_ = nullptr, *t_h1_ = nullptr, *emb_ = nullptr, *semb_ = nullptr, *temb_all_ = nullptr;
  float* gn_partial_ = nullptr;
  StepState* cur_state_ = nullptr;    // device
  StepState* step_table_ = nullptr;   // device [nsteps]
  int* step_counter_ = nullptr;       // device
  int nsteps_ = 0, method_ = 0, state_dtype_ = CFGPP_F32;
  std::vector<cfgpp_step_state> steps_host_;
  void* z_state_ = nullptr;   // (B,4,H,W) fp32-sized buffer (holds fp16 or fp32)
  void* aux_state_ = nullptr;
  void* z0t_state_ = nullptr;
  const __half** noise_slot_ = nullptr;  // device word holding noise_buf_ (read by the step kernel: graph-stable)
  __half* noise_buf_ = nullptr;          // owned, survives prepare(); re-allocated when a larger table arrives
  size_t noise_cap_ = 0;                 // elements
  const void* fwd_z_ = nullptr;  // input of the un-fused forward
  int fwd_z_dtype_ = CFGPP_F32;
  __half *fwd_eps_uc_ = nullptr, *fwd_eps_c_ = nullptr;
  Act final_norm_{nullptr, 0};
  __half* conv_in_out_ = nullptr;
  __half *conv_in_w_ = nullptr, *conv_in_b_ = nullptr, *conv_out_w_ = nullptr, *conv_out_b_ = nullptr;

  float* sk_ws_ = nullptr;       // this handle's stream-K workspace (gemm.cuh StreamKScope)
  unsigned* sk_flags_ = nullptr;
  cudaGraph_t graph_ = nullptr;
  cudaGraphExec_t graph_exec_ = nullptr;
  bool graph_valid_ = false;
  cudaStream_t capture_stream_ = nullptr;
  cudaStream_t capture_stream2_ = nullptr;
  cudaEvent_t fork_ev_ = nullptr, join_ev_ = nullptr;
};

}  // namespace cfgpp

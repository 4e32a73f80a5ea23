// cfgpp_b200 — CLIP text tower executor: the prompt conditioning of the reference —
// `pipe.text_encoder(ids)[0]` (SD v1.5, latent_diffusion.py:93-115) and `text_enc(ids, output_hidden_states=True)` ->
// `hidden_states[-2]` / `[0]` for the two SDXL encoders (latent_sdxl.py:77-93) — i.e. transformers `CLIPTextModel` /
// `CLIPTextModelWithProjection`: token + position embedding, N pre-LN layers (causal self-attention with 64-wide
// heads, MLP with quick_gelu or gelu), final LayerNorm, pooled <|endoftext|> row (+ bias-free text_projection).
// Activations are [batch * tokens][hidden] fp16; the q/k/v projections run as ONE GEMM on a concatenated weight; every
// projection / MLP GEMM is the tcgen05 kernel of gemm.cu (residual adds in its epilogue), attention / activation /
// embedding are the kernels of text_kernels.cu, LayerNorm is norm.cu's.
#pragma once
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/cfgpp_b200.h"
#include "gemm.cuh"
#include "ops.cuh"

namespace cfgpp {

void run_f32_to_f16(const float* in, __half* out, size_t n, cudaStream_t stream);
// text_kernels.cu
void run_clip_embed(const int* ids, const __half* tok, const __half* pos, __half* out, int M, int T, int D, int vocab,
                    cudaStream_t stream);
void run_clip_attention(const __half* qkv, __half* out, int B, int T, int heads, int D, cudaStream_t stream);
void run_clip_activation(__half* x, size_t n, int mode, cudaStream_t stream);
void run_clip_gather_rows(const __half* x, const int* index, __half* out, int B, int T, int D, cudaStream_t stream);

class ClipTextEncoder {
 public:
  ClipTextEncoder(const cfgpp_clip_desc& d, int device);
  ~ClipTextEncoder();
  void load_weight(const std::string& key, const void* data, const int64_t* shape, int ndim, int dtype,
                   cudaStream_t stream);
  void finalize_weights(cudaStream_t stream);
  // ids [batch][tokens] int32 (device), pooled_index [batch] int32 (device; may be null when pooled_out is null).
  // hidden_out = hidden_states[num_layers - skip] (no final LayerNorm), last_out = final_layer_norm(hidden_states[-1]),
  // pooled_out = last_out[b, pooled_index[b]] (x text_projection^T when projection_dim > 0). Null outputs are skipped.
  void encode(const int* ids, const int* pooled_index, int batch, int tokens, int skip, __half* hidden_out,
              __half* last_out, __half* pooled_out, cudaStream_t stream);
  double flops() const { return flops_; }
  size_t workspace_bytes() const { return workspace_bytes_; }

 private:
  struct Tensor {
    __half* p = nullptr;
    std::vector<int64_t> shape;
    size_t numel() const {
      size_t n = 1;
      for (auto d : shape) n *= static_cast<size_t>(d);
      return n;
    }
  };
  const Tensor& raw(const std::string& key) const;
  __half* plain(const std::string& key, size_t expect_numel) const;
  void* alloc_bytes(size_t bytes, bool weight);
  void prepare(int batch, int tokens);

  cfgpp_clip_desc d_;
  int device_;
  bool finalized_ = false;
  std::map<std::string, Tensor> raw_;
  std::vector<__half*> qkv_w_, qkv_b_;  // per layer: [3D][D], [3D]
  std::vector<void*> weight_allocs_, act_allocs_;
  size_t workspace_bytes_ = 0;
  double flops_ = 0.0;
  int B_ = 0, T_ = 0;
  using Step = std::function<void(cudaStream_t)>;
  std::vector<std::vector<Step>> layer_plan_;  // one group of launches per encoder layer
  const int* ids_in_ = nullptr;                 // set per encode() call
  __half *x0_ = nullptr, *x1_ = nullptr, *ln_ = nullptr, *qkv_ = nullptr, *att_ = nullptr, *mlp_ = nullptr;
  __half *last_ = nullptr, *pool_ = nullptr;
  float* sk_ws_ = nullptr;
  unsigned* sk_flags_ = nullptr;
};

}  // namespace cfgpp

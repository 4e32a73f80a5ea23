// cfgpp_b200 — UNet2DConditionModel executor (see unet.cuh). Host-side orchestration only; every FLOP runs in the
// hand-written kernels of gemm.cu / attention.cu / norm.cu / elementwise.cu.
#include "unet.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace cfgpp {

namespace {

__global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = __float2half_rn(in[i]);
}

// (Cout, Cin, 3, 3) -> [Cout][tap][Cin]
__global__ void pack_conv3x3_kernel(const __half* __restrict__ in, __half* __restrict__ out, int Cout, int Cin) {
  const size_t n = static_cast<size_t>(Cout) * Cin * 9;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = i % Cin;
    const size_t t = i / Cin;
    const int tap = t % 9;
    const int co = t / 9;
    out[i] = in[(static_cast<size_t>(co) * Cin + ci) * 9 + tap];
  }
}

// GEGLU proj rows (2*inner, K): per 128 rows interleave value / gate halves into 256-row tiles
__global__ void pack_geglu_kernel(const __half* __restrict__ in, __half* __restrict__ out, int inner, int K) {
  const size_t n = static_cast<size_t>(2) * inner * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = i % K;
    const int r = i / K;  // packed row
    const int tile = r / 256, w = r % 256;
    const int src_row = (w < 128) ? (tile * 128 + w) : (inner + tile * 128 + (w - 128));
    out[i] = in[static_cast<size_t>(src_row) * K + k];
  }
}

// rows of `nmat` stacked (heads*hd, K) matrices -> [(mat, head, hdp)][K], rows hd..hdp-1 of every head zero
__global__ void pack_heads_rows_kernel(const __half* const* __restrict__ mats, __half* __restrict__ out, int nmat,
                                       int heads, int hd, int hdp, int K) {
  const size_t n = static_cast<size_t>(nmat) * heads * hdp * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = i % K;
    size_t t = i / K;
    const int r = t % hdp;
    t /= hdp;
    const int h = t % heads;
    const int m = t / heads;
    out[i] = (r < hd) ? mats[m][(static_cast<size_t>(h) * hd + r) * K + k] : __float2half(0.f);
  }
}

// (N, heads*hd) -> (N, heads*hdp) with zero columns hd..hdp-1 per head
__global__ void pack_heads_cols_kernel(const __half* __restrict__ in, __half* __restrict__ out, int N, int heads, int hd,
                                       int hdp) {
  const size_t n = static_cast<size_t>(N) * heads * hdp;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = i % hdp;
    size_t t = i / hdp;
    const int h = t % heads;
    const int row = t / heads;
    out[i] = (c < hd) ? in[(static_cast<size_t>(row) * heads + h) * hd + c] : __float2half(0.f);
  }
}

// LayerNorm fold (see GemmParams): one warp per (packed) weight row n
__global__ void fold_ln_kernel(const __half* __restrict__ w, const __half* __restrict__ gamma,
                               const __half* __restrict__ beta, const __half* __restrict__ bias,
                               __half* __restrict__ wf, float* __restrict__ s_out, float* __restrict__ t_out, int N,
                               int K) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float s = 0.f, t = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = __half2float(w[static_cast<size_t>(n) * K + k]);
    const __half wg = __float2half_rn(wv * __half2float(gamma[k]));
    wf[static_cast<size_t>(n) * K + k] = wg;
    s += __half2float(wg);
    t += __half2float(beta[k]) * wv;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  if (lane == 0) {
    s_out[n] = s;
    t_out[n] = t + (bias ? __half2float(bias[n]) : 0.f);
  }
}

int grid_for(size_t n) { return static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 8)); }

}  // namespace

void gemm_configure();
void attn_configure();

// shared with vae.cu (weight ingestion helpers)
void run_f32_to_f16(const float* in, __half* out, size_t n, cudaStream_t stream) {
  f32_to_f16_kernel<<<grid_for(n), 256, 0, stream>>>(in, out, n);
  CFGPP_CHECK_CUDA(cudaGetLastError());
}
void run_pack_conv3x3(const __half* in, __half* out, int Cout, int Cin, cudaStream_t stream) {
  pack_conv3x3_kernel<<<grid_for(static_cast<size_t>(Cout) * Cin * 9), 256, 0, stream>>>(in, out, Cout, Cin);
  CFGPP_CHECK_CUDA(cudaGetLastError());
}

Unet::Unet(const cfgpp_model_desc& d, int device) : d_(d), device_(device) {
  CFGPP_CHECK_CUDA(cudaSetDevice(device));
  CFGPP_REQUIRE(d.num_levels >= 2 && d.num_levels <= CFGPP_MAX_LEVELS, "num_levels must be 2..4");
  CFGPP_REQUIRE(d.norm_num_groups == 32, "only GroupNorm(32) is implemented");
  CFGPP_REQUIRE(d.in_channels == 4 && d.out_channels == 4, "latent channels must be 4");
  time_embed_dim_ = d.block_out_channels[0] * 4;
  has_aug_ = d.addition_time_embed_dim > 0;
  gemm_configure();
  attn_configure();
  streamk_alloc(&sk_ws_, &sk_flags_);
  CFGPP_CHECK_CUDA(cudaStreamCreateWithFlags(&capture_stream_, cudaStreamNonBlocking));
  CFGPP_CHECK_CUDA(cudaStreamCreateWithFlags(&capture_stream2_, cudaStreamNonBlocking));
  CFGPP_CHECK_CUDA(cudaEventCreateWithFlags(&fork_ev_, cudaEventDisableTiming));
  CFGPP_CHECK_CUDA(cudaEventCreateWithFlags(&join_ev_, cudaEventDisableTiming));
}

Unet::~Unet() {
  if (graph_exec_) cudaGraphExecDestroy(graph_exec_);
  if (graph_) cudaGraphDestroy(graph_);
  if (capture_stream_) cudaStreamDestroy(capture_stream_);
  if (capture_stream2_) cudaStreamDestroy(capture_stream2_);
  if (fork_ev_) cudaEventDestroy(fork_ev_);
  if (join_ev_) cudaEventDestroy(join_ev_);
  for (auto& kv : raw_) cudaFree(kv.second.p);
  for (void* p : weight_allocs_) cudaFree(p);
  for (void* p : act_allocs_) cudaFree(p);
  if (noise_buf_) cudaFree(noise_buf_);
  streamk_free(sk_ws_, sk_flags_);
}

// ------------------------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------------------------
void Unet::load_weight(const std::string& key, const void* data, const int64_t* shape, int ndim, int dtype,
                       cudaStream_t stream) {
  CFGPP_REQUIRE(!finalized_, "weights already finalized");
  CFGPP_REQUIRE(dtype == CFGPP_F16 || dtype == CFGPP_F32, "weight dtype must be fp16 or fp32");
  DevTensor t;
  t.shape.assign(shape, shape + ndim);
  const size_t n = t.numel();
  CFGPP_CHECK_CUDA(cudaMalloc(&t.p, std::max<size_t>(n, 8) * sizeof(__half)));
  weight_bytes_ += n * sizeof(__half);
  if (dtype == CFGPP_F16) {
    CFGPP_CHECK_CUDA(cudaMemcpyAsync(t.p, data, n * sizeof(__half), cudaMemcpyDeviceToDevice, stream));
  } else {
    f32_to_f16_kernel<<<grid_for(n), 256, 0, stream>>>(static_cast<const float*>(data), t.p, n);
    CFGPP_CHECK_CUDA(cudaGetLastError());
  }
  auto it = raw_.find(key);
  if (it != raw_.end()) {
    cudaFree(it->second.p);
    raw_.erase(it);
  }
  raw_[key] = t;
}

const DevTensor& Unet::raw(const std::string& key) const {
  auto it = raw_.find(key);
  if (it == raw_.end()) throw Error(-10, "missing weight: " + key);
  return it->second;
}

__half* Unet::alloc_weight(size_t numel) {
  void* p = nullptr;
  CFGPP_CHECK_CUDA(cudaMalloc(&p, std::max<size_t>(numel, 8) * sizeof(__half)));
  weight_allocs_.push_back(p);
  weight_bytes_ += numel * sizeof(__half);
  return static_cast<__half*>(p);
}

__half* Unet::plain(const std::string& key) { return raw(key).p; }

__half* Unet::packed_conv3x3(const std::string& key) {
  auto it = packed_cache_.find(key);
  if (it != packed_cache_.end()) return it->second;
  const DevTensor& t = raw(key);
  CFGPP_REQUIRE(t.shape.size() == 4 && t.shape[2] == 3 && t.shape[3] == 3, "expected (Cout,Cin,3,3): " + key);
  const int Cout = static_cast<int>(t.shape[0]), Cin = static_cast<int>(t.shape[1]);
  __half* out = alloc_weight(t.numel());
  pack_conv3x3_kernel<<<grid_for(t.numel()), 256>>>(t.p, out, Cout, Cin);
  CFGPP_CHECK_CUDA(cudaGetLastError());
  packed_cache_[key] = out;
  return out;
}

__half* Unet::packed_cat_rows(const std::vector<std::string>& keys) {
  std::string name = "cat:";
  for (auto& k : keys) name += k + "|";
  auto it = packed_cache_.find(name);
  if (it != packed_cache_.end()) return it->second;
  size_t total = 0;
  for (auto& k : keys) total += raw(k).numel();
  __half* out = alloc_weight(total);
  size_t off = 0;
  for (auto& k : keys) {
    const DevTensor& t = raw(k);
    CFGPP_CHECK_CUDA(cudaMemcpy(out + off, t.p, t.numel() * sizeof(__half), cudaMemcpyDeviceToDevice));
    off += t.numel();
  }
  packed_cache_[name] = out;
  return out;
}

__half* Unet::packed_geglu(const std::string& key, bool is_bias) {
  auto it = packed_cache_.find("geglu:" + key);
  if (it != packed_cache_.end()) return it->second;
  const DevTensor& t = raw(key);
  const int rows = static_cast<int>(t.shape[0]);
  const int K = is_bias ? 1 : static_cast<int>(t.shape[1]);
  CFGPP_REQUIRE(rows % 256 == 0, "GEGLU width must be a multiple of 256: " + key);
  __half* out = alloc_weight(t.numel());
  pack_geglu_kernel<<<grid_for(t.numel()), 256>>>(t.p, out, rows / 2, K);
  CFGPP_CHECK_CUDA(cudaGetLastError());
  packed_cache_["geglu:" + key] = out;
  return out;
}

__half* Unet::packed_heads_rows(const std::vector<std::string>& keys, int heads, int hd, int hdp) {
  if (hd == hdp) return keys.size() == 1 ? plain(keys[0]) : packed_cat_rows(keys);
  std::string name = "heads_rows:";
  for (auto& k : keys) name += k + "|";
  auto it = packed_cache_.find(name);
  if (it != packed_cache_.end()) return it->second;
  const int K = static_cast<int>(raw(keys[0]).shape[1]);
  std::vector<const __half*> ptrs;
  for (auto& k : keys) {
    const DevTensor& t = raw(k);
    CFGPP_REQUIRE(t.shape.size() >= 2 && t.shape[0] == heads * hd && t.shape[1] == K, "unexpected projection shape: " + k);
    ptrs.push_back(t.p);
  }
  const __half** dptrs = nullptr;
  CFGPP_CHECK_CUDA(cudaMalloc(&dptrs, ptrs.size() * sizeof(__half*)));
  CFGPP_CHECK_CUDA(cudaMemcpy(dptrs, ptrs.data(), ptrs.size() * sizeof(__half*), cudaMemcpyHostToDevice));
  const size_t total = keys.size() * static_cast<size_t>(heads) * hdp * K;
  __half* out = alloc_weight(total);
  pack_heads_rows_kernel<<<grid_for(total), 256>>>(dptrs, out, static_cast<int>(keys.size()), heads, hd, hdp, K);
  CFGPP_CHECK_CUDA(cudaGetLastError());
  CFGPP_CHECK_CUDA(cudaDeviceSynchronize());
  cudaFree(dptrs);
  packed_cache_[name] = out;
  return out;
}

__half* Unet::packed_heads_cols(const std::string& key, int heads, int hd, int hdp) {
  if (hd == hdp) return plain(key);
  auto it = packed_cache_.find("heads_cols:" + key);
  if (it != packed_cache_.end()) return it->second;
  const DevTensor& t = raw(key);
  const int N = static_cast<int>(t.shape[0]);
  CFGPP_REQUIRE(t.shape[1] == heads * hd, "unexpected to_out shape: " + key);
  const size_t total = static_cast<size_t>(N) * heads * hdp;
  __half* out = alloc_weight(total);
  pack_heads_cols_kernel<<<grid_for(total), 256>>>(t.p, out, N, heads, hd, hdp);
  CFGPP_CHECK_CUDA(cudaGetLastError());
  packed_cache_["heads_cols:" + key] = out;
  return out;
}

Unet::FoldedLN Unet::folded_ln(const std::string& cache_key, const __half* w_packed, int N, int K,
                               const std::string& norm_prefix, const __half* bias_packed) {
  auto it = fold_cache_.find(cache_key);
  if (it != fold_cache_.end()) return it->second;
  FoldedLN f;
  f.w = alloc_weight(static_cast<size_t>(N) * K);
  f.s = reinterpret_cast<float*>(alloc_weight(static_cast<size_t>(N) * 2));
  f.t = reinterpret_cast<float*>(alloc_weight(static_cast<size_t>(N) * 2));
  const int warps = 8;
  fold_ln_kernel<<<(N + warps - 1) / warps, warps * 32>>>(w_packed, plain(norm_prefix + ".weight"),
                                                           plain(norm_prefix + ".bias"), bias_packed, f.w, f.s, f.t, N, K);
  CFGPP_CHECK_CUDA(cudaGetLastError());
  fold_cache_[cache_key] = f;
  return f;
}

bool Unet::lnfold_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CFGPP_NO_LNFOLD");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

void Unet::finalize_weights(cudaStream_t stream) {
  CFGPP_CHECK_CUDA(cudaStreamSynchronize(stream));
  // structural validation: every key the plan will touch must exist (dry walk at a nominal size)
  finalized_ = true;
  try {
    // the smallest latent for which every level keeps a spatial extent (H, W >= 1 at the deepest level)
    const int s = 1 << (d_.num_levels - 1);
    prepare(1, std::max(8, s * 8), std::max(8, s * 8));
    prepared_ = false;
  } catch (...) {
    finalized_ = false;
    throw;
  }
}

// ------------------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------------------
void* Unet::alloc_bytes(size_t bytes) {
  void* p = nullptr;
  bytes = (bytes + 255) & ~static_cast<size_t>(255);
  CFGPP_CHECK_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 256)));
  act_allocs_.push_back(p);
  workspace_bytes_ += bytes;
  return p;
}

__half* Unet::alloc_act(size_t numel) { return static_cast<__half*>(alloc_bytes(numel * sizeof(__half))); }

Unet::Scratch* Unet::scratch(const std::string& name, size_t numel_half) {
  auto& s = scratch_[name];
  if (!s) s.reset(new Scratch());
  if (s->p == nullptr) {
    s->need = std::max(s->need, numel_half);
  } else {
    CFGPP_REQUIRE(numel_half <= s->need, "scratch undersized: " + name);
  }
  return s.get();
}

// ------------------------------------------------------------------------------------------------------------
// plan building. The structure is walked twice by prepare(): a sizing pass (scratch buffers unallocated: only
// sizes are recorded, no ops are created) and the real pass.
// ------------------------------------------------------------------------------------------------------------
namespace {
bool g_dry = false;
}

void Unet::add_step(const std::string& name, std::function<void(cudaStream_t)> fn, int launches) {
  if (g_dry) return;
  PlanStep s;
  s.name = name;
  s.fn = std::move(fn);
  s.launches = launches;
  cur_plan_->push_back(std::move(s));
}

void Unet::add_gemm(const std::string& name, const GemmOp& op, double algorithmic_flops) {
  PlanStep s;
  s.name = name;
  s.flops = algorithmic_flops >= 0 ? algorithmic_flops : op.flops();
  s.kind = op.p.conv ? 1 : 0;
  s.fn = [op](cudaStream_t st) { run_gemm_op(op, st); };
  cur_plan_->push_back(std::move(s));
}

void Unet::add_attn(const std::string& name, const AttnOp& op) {
  PlanStep s;
  s.name = name;
  s.flops = op.flops();
  s.kind = 2;
  s.fn = [op](cudaStream_t st) { run_attn_op(op, st); };
  cur_plan_->push_back(std::move(s));
}

Unet::Act Unet::build_resnet(const std::string& prefix, Act x1, const Act* x2, int Cout, int H, int W, int temb_off) {
  const int C1 = x1.C, C2 = x2 ? x2->C : 0, Cin = C1 + C2;
  const int HW = H * W;
  const size_t M = static_cast<size_t>(bnb_) * HW;
  Scratch* s_norm = scratch(btag_ + "norm", M * std::max(Cin, Cout));
  Scratch* s_h1 = scratch(btag_ + "h1", M * Cout);
  Scratch* s_sc = (Cin != Cout) ? scratch(btag_ + "shortcut", M * Cout) : nullptr;
  __half* out = g_dry ? nullptr : (out_override_ ? out_override_ : alloc_act(M * Cout));
  out_override_ = nullptr;
  if (g_dry) {
    // validate keys
    raw(prefix + ".norm1.weight"); raw(prefix + ".norm1.bias"); raw(prefix + ".conv1.weight");
    raw(prefix + ".conv1.bias"); raw(prefix + ".time_emb_proj.weight"); raw(prefix + ".time_emb_proj.bias");
    raw(prefix + ".norm2.weight"); raw(prefix + ".norm2.bias"); raw(prefix + ".conv2.weight");
    raw(prefix + ".conv2.bias");
    if (Cin != Cout) { raw(prefix + ".conv_shortcut.weight"); raw(prefix + ".conv_shortcut.bias"); }
    workspace_bytes_ += M * Cout * sizeof(__half);
    return Act{nullptr, Cout};
  }
  const __half* x2p = x2 ? x2->p : nullptr;
  const __half *g1 = plain(prefix + ".norm1.weight"), *b1 = plain(prefix + ".norm1.bias");
  const __half *g2 = plain(prefix + ".norm2.weight"), *b2 = plain(prefix + ".norm2.bias");
  const float eps = d_.norm_eps;
  float* partial = bgn_partial_;
  const int NB = bnb_;
  __half* normp = s_norm->p;
  __half* h1p = s_h1->p;
  const __half* x1p = x1.p;
  add_step(prefix + ".norm1+silu", [=](cudaStream_t st) {
    run_groupnorm(x1p, C1, x2p, C2, NB, HW, g1, b1, eps, true, partial, normp, st);
  }, 2);
  add_gemm(prefix + ".conv1", make_conv3x3_op(normp, bnb_, H, W, Cin, packed_conv3x3(prefix + ".conv1.weight"), Cout,
                                              plain(prefix + ".conv1.bias"), temb_all_ + static_cast<size_t>(brow0_) * temb_total_ + temb_off, temb_total_, HW,
                                              h1p));
  add_step(prefix + ".norm2+silu", [=](cudaStream_t st) {
    run_groupnorm(h1p, Cout, nullptr, 0, NB, HW, g2, b2, eps, true, partial, normp, st);
  }, 2);
  const __half* residual = x1p;
  if (Cin != Cout) {
    add_gemm(prefix + ".conv_shortcut",
             make_linear_op(x1p, C1, x2p, C2, C1, plain(prefix + ".conv_shortcut.weight"), static_cast<int>(M), Cout,
                            Cin, plain(prefix + ".conv_shortcut.bias"), nullptr, 0, 1, s_sc->p, Cout, false));
    residual = s_sc->p;
  }
  add_gemm(prefix + ".conv2", make_conv3x3_op(normp, bnb_, H, W, Cout, packed_conv3x3(prefix + ".conv2.weight"), Cout,
                                              plain(prefix + ".conv2.bias"), residual, Cout, 1, out));
  return Act{out, Cout};
}

Unet::Act Unet::build_transformer(const std::string& prefix, Act x, int H, int W, int layers, int heads) {
  const int C = x.C;
  const int HW = H * W;
  const int Mi = bnb_ * HW;
  const size_t M = static_cast<size_t>(Mi);
  const int D = d_.cross_attention_dim;
  CFGPP_REQUIRE(C % heads == 0 && C / heads <= 192,
                "attention supports head_dim <= 192 (got " + std::to_string(C / std::max(heads, 1)) + ") at " + prefix);
  const int hd = C / heads;
  const int hdp = attn_padded_head_dim(hd);  // heads are zero-padded to a multiple of 64 channels (SD v1.5: 40/80/160)
  const int Cp = heads * hdp;
  Scratch* s_norm = scratch(btag_ + "norm", M * C);
  Scratch* s_tok = scratch(btag_ + "tokens", M * C);
  Scratch* s_qkv = scratch(btag_ + "qkv", M * 3 * Cp);
  Scratch* s_attn = scratch(btag_ + "attn", M * Cp);
  Scratch* s_q = scratch(btag_ + "q", M * Cp);
  Scratch* s_ff = scratch(btag_ + "ff", M * 4 * C);
  Scratch* s_stats[3];
  for (int i = 0; i < 3; ++i)  // [16 N blocks][M] float2 partial row statistics (LayerNorm fold)
    s_stats[i] = scratch(btag_ + "lnstats" + std::to_string(i), static_cast<size_t>(32) * M * 2 * 2);
  __half* out = g_dry ? nullptr : alloc_act(M * C);
  const int Mkv = bnb_ * n_ctx_;
  if (g_dry) {
    raw(prefix + ".norm.weight"); raw(prefix + ".norm.bias"); raw(prefix + ".proj_in.weight");
    raw(prefix + ".proj_in.bias"); raw(prefix + ".proj_out.weight"); raw(prefix + ".proj_out.bias");
    for (int k = 0; k < layers; ++k) {
      const std::string b = prefix + ".transformer_blocks." + std::to_string(k);
      for (const char* n : {".norm1.weight", ".norm1.bias", ".norm2.weight", ".norm2.bias", ".norm3.weight",
                            ".norm3.bias", ".attn1.to_q.weight", ".attn1.to_k.weight", ".attn1.to_v.weight",
                            ".attn1.to_out.0.weight", ".attn1.to_out.0.bias", ".attn2.to_q.weight",
                            ".attn2.to_k.weight", ".attn2.to_v.weight", ".attn2.to_out.0.weight",
                            ".attn2.to_out.0.bias", ".ff.net.0.proj.weight", ".ff.net.0.proj.bias",
                            ".ff.net.2.weight", ".ff.net.2.bias"})
        raw(b + n);
      workspace_bytes_ += static_cast<size_t>(Mkv) * 2 * Cp * sizeof(__half);
    }
    workspace_bytes_ += M * C * sizeof(__half);
    return Act{nullptr, C};
  }
  const int NB = bnb_;
  float* partial = bgn_partial_;
  __half *normp = s_norm->p, *tok = s_tok->p, *qkv = s_qkv->p, *attn = s_attn->p, *qb = s_q->p, *ff = s_ff->p;
  // LayerNorm fold: the GEMMs that write the residual stream `tok` also emit per-row partial statistics, and the
  // GEMMs that read LN(tok) run on `tok` with gamma folded into the weight (no LayerNorm launches at all)
  const bool fold = !lnfold_disabled();
  float* stats[3] = {nullptr, nullptr, nullptr};
  int parts[3] = {0, 0, 0};
  if (fold)
    for (int i = 0; i < 3; ++i) stats[i] = reinterpret_cast<float*>(s_stats[i]->p);
  // a producer's N blocks must tile C exactly so that every column contributes to the row statistics
  auto producer = [&](const std::function<GemmOp(int)>& make, int slot) {
    GemmOp op = make(0);
    if (fold) {
      if (C % op.bn != 0) {
        for (int bn : {160, 128, 64})
          if (C % bn == 0) {
            op = make(bn);
            break;
          }
      }
      CFGPP_REQUIRE(C % op.bn == 0 && op.p.num_n_blocks <= 16, "LayerNorm fold: no tile width divides C");
      // two partial sums per N block: the epilogue splits a tile's columns between two warps per row
      op.p.stats_out = stats[slot];
      parts[slot] = 2 * op.p.num_n_blocks;
    }
    return op;
  };
  auto consumer = [&](GemmOp op, const FoldedLN& f, int slot) {
    op.p.stats_in = stats[slot];
    op.p.ln_parts = parts[slot];
    op.p.ln_inv_c = 1.0f / static_cast<float>(C);
    op.p.ln_eps = 1e-5f;
    op.p.ln_s = f.s;
    op.p.ln_t = f.t;
    return op;
  };
  {
    const __half *g = plain(prefix + ".norm.weight"), *b = plain(prefix + ".norm.bias");
    const __half* xp = x.p;
    add_step(prefix + ".norm", [=](cudaStream_t st) {
      run_groupnorm(xp, C, nullptr, 0, NB, HW, g, b, 1e-6f, false, partial, normp, st);
    }, 2);
  }
  add_gemm(prefix + ".proj_in", producer([&](int bn) {
             return make_linear_op(normp, C, nullptr, 0, 0, plain(prefix + ".proj_in.weight"), Mi, C, C,
                                   plain(prefix + ".proj_in.bias"), nullptr, 0, 1, tok, C, false, bn);
           }, 0));
  for (int k = 0; k < layers; ++k) {
    const std::string b = prefix + ".transformer_blocks." + std::to_string(k);
    auto add_ln = [&](const std::string& n) {
      const __half *g = plain(b + n + ".weight"), *be = plain(b + n + ".bias");
      add_step(b + n, [=](cudaStream_t st) { run_layernorm(tok, Mi, C, g, be, 1e-5f, normp, st); });
    };
    // --- self-attention ---
    __half* wqkv = packed_heads_rows({b + ".attn1.to_q.weight", b + ".attn1.to_k.weight", b + ".attn1.to_v.weight"},
                                     heads, hd, hdp);
    if (fold) {
      const FoldedLN f = folded_ln(b + ".attn1.qkv", wqkv, 3 * Cp, C, b + ".norm1", nullptr);
      add_gemm(b + ".attn1.to_qkv(+norm1)",
               consumer(make_linear_op(tok, C, nullptr, 0, 0, f.w, Mi, 3 * Cp, C, nullptr, nullptr, 0, 1, qkv, 3 * Cp, false),
                        f, 0),
               2.0 * Mi * 3.0 * C * C);
    } else {
      add_ln(".norm1");
      add_gemm(b + ".attn1.to_qkv",
               make_linear_op(normp, C, nullptr, 0, 0, wqkv, Mi, 3 * Cp, C, nullptr, nullptr, 0, 1, qkv, 3 * Cp, false),
               2.0 * Mi * 3.0 * C * C);
    }
    add_attn(b + ".attn1.sdpa",
             make_attn_op(qkv, 3 * Cp, qkv + Cp, 3 * Cp, qkv + 2 * Cp, 3 * Cp, attn, Cp, bnb_, heads, HW, HW, hd));
    add_gemm(b + ".attn1.to_out", producer([&](int bn) {
               return make_linear_op(attn, Cp, nullptr, 0, 0,
                                     packed_heads_cols(b + ".attn1.to_out.0.weight", heads, hd, hdp), Mi, C, Cp,
                                     plain(b + ".attn1.to_out.0.bias"), tok, C, 1, tok, C, false, bn);
             }, 1),
             2.0 * Mi * static_cast<double>(C) * C);
    // --- cross-attention (K/V projected once per prompt by the prompt plan) ---
    __half* wq2 = packed_heads_rows({b + ".attn2.to_q.weight"}, heads, hd, hdp);
    if (fold) {
      const FoldedLN f = folded_ln(b + ".attn2.q", wq2, Cp, C, b + ".norm2", nullptr);
      add_gemm(b + ".attn2.to_q(+norm2)",
               consumer(make_linear_op(tok, C, nullptr, 0, 0, f.w, Mi, Cp, C, nullptr, nullptr, 0, 1, qb, Cp, false), f, 1),
               2.0 * Mi * static_cast<double>(C) * C);
    } else {
      add_ln(".norm2");
      add_gemm(b + ".attn2.to_q",
               make_linear_op(normp, C, nullptr, 0, 0, wq2, Mi, Cp, C, nullptr, nullptr, 0, 1, qb, Cp, false),
               2.0 * Mi * static_cast<double>(C) * C);
    }
    __half* kv = alloc_act(static_cast<size_t>(Mkv) * 2 * Cp);
    {
      __half* wkv = packed_heads_rows({b + ".attn2.to_k.weight", b + ".attn2.to_v.weight"}, heads, hd, hdp);
      std::vector<PlanStep>* save = cur_plan_;
      cur_plan_ = &prompt_plan_;
      add_gemm(b + ".attn2.to_kv",
               make_linear_op(ctx_copy_ + static_cast<size_t>(brow0_) * n_ctx_ * D, D, nullptr, 0, 0, wkv, Mkv, 2 * Cp, D,
                              nullptr, nullptr, 0, 1, kv, 2 * Cp, false),
               2.0 * Mkv * 2.0 * C * D);
      cur_plan_ = save;
    }
    add_attn(b + ".attn2.sdpa", make_attn_op(qb, Cp, kv, 2 * Cp, kv + Cp, 2 * Cp, attn, Cp, bnb_, heads, HW, n_ctx_, hd));
    add_gemm(b + ".attn2.to_out", producer([&](int bn) {
               return make_linear_op(attn, Cp, nullptr, 0, 0,
                                     packed_heads_cols(b + ".attn2.to_out.0.weight", heads, hd, hdp), Mi, C, Cp,
                                     plain(b + ".attn2.to_out.0.bias"), tok, C, 1, tok, C, false, bn);
             }, 2),
             2.0 * Mi * static_cast<double>(C) * C);
    // --- GEGLU feed-forward ---
    __half* wg = packed_geglu(b + ".ff.net.0.proj.weight", false);
    __half* bg = packed_geglu(b + ".ff.net.0.proj.bias", true);
    if (fold) {
      const FoldedLN f = folded_ln(b + ".ff.geglu", wg, 8 * C, C, b + ".norm3", bg);
      add_gemm(b + ".ff.geglu(+norm3)",
               consumer(make_linear_op(tok, C, nullptr, 0, 0, f.w, Mi, 8 * C, C, nullptr, nullptr, 0, 1, ff, 4 * C, true), f, 2));
    } else {
      add_ln(".norm3");
      add_gemm(b + ".ff.geglu",
               make_linear_op(normp, C, nullptr, 0, 0, wg, Mi, 8 * C, C, bg, nullptr, 0, 1, ff, 4 * C, true));
    }
    add_gemm(b + ".ff.out", producer([&](int bn) {
               return make_linear_op(ff, 4 * C, nullptr, 0, 0, plain(b + ".ff.net.2.weight"), Mi, C, 4 * C,
                                     plain(b + ".ff.net.2.bias"), tok, C, 1, tok, C, false, bn);
             }, 0));
  }
  add_gemm(prefix + ".proj_out", make_linear_op(tok, C, nullptr, 0, 0, plain(prefix + ".proj_out.weight"), Mi, C, C,
                                                plain(prefix + ".proj_out.bias"), x.p, C, 1, out, C, false));
  return Act{out, C};
}

Unet::Act Unet::build_downsample(const std::string& prefix, Act x, int H, int W) {
  const int C = x.C;
  const int Ho = H / 2, Wo = W / 2;
  const size_t Mo = static_cast<size_t>(bnb_) * Ho * Wo;
  static const bool use_im2col = [] {  // CFGPP_NO_S2TMA=1: the round-1 path (materialised stride-2 im2col + plain GEMM)
    const char* e = getenv("CFGPP_NO_S2TMA");
    return e != nullptr && e[0] == '1';
  }();
  Scratch* s_col = use_im2col ? scratch(btag_ + "im2col", Mo * 9 * C) : nullptr;
  __half* out = g_dry ? nullptr : alloc_act(Mo * C);
  if (g_dry) {
    raw(prefix + ".conv.weight"); raw(prefix + ".conv.bias");
    workspace_bytes_ += Mo * C * sizeof(__half);
    return Act{nullptr, C};
  }
  if (!use_im2col) {
    // stride-2 conv as an implicit GEMM: the A tile of every tap comes through a tensor map with element strides 2
    add_gemm(prefix + ".conv", make_conv3x3_op(x.p, bnb_, H, W, C, packed_conv3x3(prefix + ".conv.weight"), C,
                                               plain(prefix + ".conv.bias"), nullptr, 0, 1, out, 0, 2));
    return Act{out, C};
  }
  const int NB = bnb_;
  const __half* xp = x.p;
  __half* col = s_col->p;
  add_step(prefix + ".im2col", [=](cudaStream_t st) { run_im2col_s2(xp, col, NB, H, W, C, st); });
  add_gemm(prefix + ".conv", make_linear_op(col, 9 * C, nullptr, 0, 0, packed_conv3x3(prefix + ".conv.weight"),
                                            static_cast<int>(Mo), C, 9 * C, plain(prefix + ".conv.bias"), nullptr, 0, 1,
                                            out, C, false));
  return Act{out, C};
}

Unet::Act Unet::build_upsample(const std::string& prefix, Act x, int H, int W) {
  const int C = x.C;
  const size_t Mo = static_cast<size_t>(bnb_) * 4 * H * W;
  Scratch* s_up = scratch(btag_ + "upsampled", Mo * C);
  __half* out = g_dry ? nullptr : alloc_act(Mo * C);
  if (g_dry) {
    raw(prefix + ".conv.weight"); raw(prefix + ".conv.bias");
    workspace_bytes_ += Mo * C * sizeof(__half);
    return Act{nullptr, C};
  }
  const int NB = bnb_;
  const __half* xp = x.p;
  __half* up = s_up->p;
  add_step(prefix + ".nearest2x", [=](cudaStream_t st) { run_upsample2x(xp, up, NB, H, W, C, st); });
  add_gemm(prefix + ".conv", make_conv3x3_op(up, bnb_, 2 * H, 2 * W, C, packed_conv3x3(prefix + ".conv.weight"), C,
                                             plain(prefix + ".conv.bias"), nullptr, 0, 1, out));
  return Act{out, C};
}

void Unet::prepare(int batch, int h_lat, int w_lat) {
  CFGPP_REQUIRE(finalized_, "call cfgpp_finalize_weights first");
  // validate BEFORE anything is freed: a rejected shape must leave the previous plan usable
  CFGPP_REQUIRE(batch >= 1 && 2 * batch <= 16, "batch must be 1..8 (UNet batch 2*batch <= 16)");
  {
    const int down = 1 << (d_.num_levels - 1);
    CFGPP_REQUIRE(h_lat >= down && w_lat >= down && h_lat % down == 0 && w_lat % down == 0,
                  "latent H, W must be multiples of 2^(num_levels-1)");
    for (int i = 0, h = h_lat, w = w_lat; i < d_.num_levels; ++i, h /= 2, w /= 2)
      CFGPP_REQUIRE(conv3x3_geometry_supported(h, w),
                    "conv3x3 tiler: unsupported level geometry " + std::to_string(h) + "x" + std::to_string(w));
  }
  CFGPP_CHECK_CUDA(cudaSetDevice(device_));
  CFGPP_CHECK_CUDA(cudaDeviceSynchronize());
  StreamKScope sk_scope(sk_ws_, sk_flags_);  // every GEMM op built below parks its stream-K partials in OUR workspace
  // from here on the old plan is gone: a throw below must not leave the handle looking prepared
  prepared_ = false;
  nsteps_ = 0;
  graph_valid_ = false;
  // drop the previous plan / workspace
  for (void* p : act_allocs_) cudaFree(p);
  act_allocs_.clear();
  scratch_.clear();
  prologue_plan_.clear();
  branch_plan_[0].clear();
  branch_plan_[1].clear();
  tail_plan_.clear();
  prompt_plan_.clear();
  workspace_bytes_ = 0;
  graph_valid_ = false;
  B_ = batch; NB_ = 2 * batch; H_ = h_lat; W_ = w_lat;
  const int L = d_.num_levels;
  const int C0 = d_.block_out_channels[0];
  const int TE = time_embed_dim_;

  // resnet order (= temb offsets) is fixed by the structure walk below; compute it first
  temb_order_.clear();
  temb_total_ = 0;
  std::vector<std::string> temb_w_keys, temb_b_keys;
  auto reg_resnet = [&](const std::string& prefix, int Cout) {
    temb_order_.push_back({prefix, temb_total_});
    temb_total_ += Cout;
    temb_w_keys.push_back(prefix + ".time_emb_proj.weight");
    temb_b_keys.push_back(prefix + ".time_emb_proj.bias");
  };
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < d_.layers_per_block; ++j)
      reg_resnet("down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), d_.block_out_channels[i]);
  reg_resnet("mid_block.resnets.0", d_.block_out_channels[L - 1]);
  reg_resnet("mid_block.resnets.1", d_.block_out_channels[L - 1]);
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < d_.layers_per_block + 1; ++j)
      reg_resnet("up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), d_.block_out_channels[L - 1 - i]);
  auto temb_off = [&](const std::string& prefix) {
    for (auto& p : temb_order_)
      if (p.first == prefix) return p.second;
    throw Error(-11, "internal: unknown resnet " + prefix);
  };

  for (int pass = 0; pass < 2; ++pass) {
    g_dry = (pass == 0);
    if (!g_dry) {
      workspace_bytes_ = 0;
      // allocate scratch + fixed buffers now that sizes are known
      for (auto& kv : scratch_) kv.second->p = alloc_act(kv.second->need);
      gn_partial_ = static_cast<float*>(alloc_bytes(static_cast<size_t>(NB_) * 128 * 64 * sizeof(float)));
      t_sin_ = alloc_act(C0);
      t_h1_ = alloc_act(TE);
      emb_ = alloc_act(static_cast<size_t>(NB_) * TE);
      semb_ = alloc_act(static_cast<size_t>(NB_) * TE);
      temb_all_ = alloc_act(static_cast<size_t>(NB_) * temb_total_);
      ctx_copy_ = alloc_act(static_cast<size_t>(NB_) * n_ctx_ * d_.cross_attention_dim);
      if (has_aug_) {
        add_in_ = alloc_act(static_cast<size_t>(NB_) * d_.projection_class_embeddings_input_dim);
        add_h1_ = alloc_act(static_cast<size_t>(NB_) * TE);
        aug_emb_ = alloc_act(static_cast<size_t>(NB_) * TE);
        pooled_copy_ = alloc_act(static_cast<size_t>(NB_) * d_.pooled_dim);
        time_ids_copy_ = static_cast<float*>(alloc_bytes(static_cast<size_t>(NB_) * 6 * sizeof(float)));
      }
      cur_state_ = static_cast<StepState*>(alloc_bytes(sizeof(StepState)));
      step_counter_ = static_cast<int*>(alloc_bytes(sizeof(int)));
      step_table_ = static_cast<StepState*>(alloc_bytes(sizeof(StepState) * 1024));
      const size_t lat = static_cast<size_t>(B_) * 4 * H_ * W_;
      z_state_ = alloc_bytes(lat * sizeof(float));
      aux_state_ = alloc_bytes(lat * sizeof(float));
      z0t_state_ = alloc_bytes(lat * sizeof(float));
      noise_slot_ = static_cast<const __half**>(alloc_bytes(sizeof(__half*)));
      CFGPP_CHECK_CUDA(cudaMemcpy(noise_slot_, &noise_buf_, sizeof(__half*), cudaMemcpyHostToDevice));
      fwd_eps_uc_ = alloc_act(lat);
      fwd_eps_c_ = alloc_act(lat);
      temb_w_all_ = packed_cat_rows(temb_w_keys);
      temb_b_all_ = packed_cat_rows(temb_b_keys);
      conv_in_w_ = plain("conv_in.weight");
      conv_in_b_ = plain("conv_in.bias");
      conv_out_w_ = packed_conv3x3("conv_out.weight");
      conv_out_b_ = plain("conv_out.bias");
    } else {
      for (const char* k : {"conv_in.weight", "conv_in.bias", "conv_out.weight", "conv_out.bias",
                            "conv_norm_out.weight", "conv_norm_out.bias", "time_embedding.linear_1.weight",
                            "time_embedding.linear_1.bias", "time_embedding.linear_2.weight",
                            "time_embedding.linear_2.bias"})
        raw(k);
      if (has_aug_)
        for (const char* k : {"add_embedding.linear_1.weight", "add_embedding.linear_1.bias",
                              "add_embedding.linear_2.weight", "add_embedding.linear_2.bias"})
          raw(k);
    }

    // ---- prologue: timestep embedding -> per-resnet time_emb_proj (ResnetBlock2D temb) ----
    cur_plan_ = &prologue_plan_;
    if (!g_dry) {
      const int NB = NB_;
      const __half *w1 = plain("time_embedding.linear_1.weight"), *b1 = plain("time_embedding.linear_1.bias");
      const __half *w2 = plain("time_embedding.linear_2.weight"), *b2 = plain("time_embedding.linear_2.bias");
      __half *t_sin = t_sin_, *t_h1 = t_h1_, *emb = emb_, *semb = semb_, *temb_all = temb_all_;
      const __half* aug = has_aug_ ? aug_emb_ : nullptr;
      const StepState* cur = cur_state_;
      const __half *wa = temb_w_all_, *ba = temb_b_all_;
      const int ttot = temb_total_;
      add_step("time_proj", [=](cudaStream_t st) { run_sincos_embed(&cur->t, 1, 1, C0, t_sin, C0, 0, st); });
      add_step("time_embedding.linear_1+silu", [=](cudaStream_t st) {
        run_small_linear(t_sin, C0, w1, b1, nullptr, 0, t_h1, TE, nullptr, 1, TE, C0, true, st);
      });
      add_step("time_embedding.linear_2(+aug_emb)", [=](cudaStream_t st) {
        run_small_linear(t_h1, 0, w2, b2, aug, TE, emb, TE, semb, NB, TE, TE, false, st);
      });
      add_step("resnets.time_emb_proj", [=](cudaStream_t st) {
        run_small_linear(semb, TE, wa, ba, nullptr, 0, temb_all, ttot, nullptr, NB, ttot, TE, false, st);
      });
    }

    // ---- prompt plan: add-embedding (SDXL text_time) ----
    cur_plan_ = &prompt_plan_;
    if (!g_dry && has_aug_) {
      const int NB = NB_;
      const int ATE = d_.addition_time_embed_dim, PD = d_.pooled_dim, AIN = d_.projection_class_embeddings_input_dim;
      CFGPP_REQUIRE(AIN == PD + 6 * ATE, "projection_class_embeddings_input_dim != pooled_dim + 6*addition_time_embed_dim");
      const __half *w1 = plain("add_embedding.linear_1.weight"), *b1 = plain("add_embedding.linear_1.bias");
      const __half *w2 = plain("add_embedding.linear_2.weight"), *b2 = plain("add_embedding.linear_2.bias");
      __half *add_in = add_in_, *add_h1 = add_h1_, *aug = aug_emb_, *pooled = pooled_copy_;
      float* tids = time_ids_copy_;
      add_step("add_embedding.assemble", [=](cudaStream_t st) {
        run_copy_rows(pooled, NB, PD, add_in, AIN, 0, NB, st);
        for (int j = 0; j < 6; ++j) run_sincos_embed(tids + j, 6, NB, ATE, add_in, AIN, PD + j * ATE, st);
      }, 7);
      add_step("add_embedding.linear_1+silu", [=](cudaStream_t st) {
        run_small_linear(add_in, AIN, w1, b1, nullptr, 0, add_h1, TE, nullptr, NB, TE, AIN, true, st);
      });
      add_step("add_embedding.linear_2", [=](cudaStream_t st) {
        run_small_linear(add_h1, TE, w2, b2, nullptr, 0, aug, TE, nullptr, NB, TE, TE, false, st);
      });
    }

    // ---- body ----
    // The unconditional and the conditional halves of the UNet batch never interact before the CFG++ mix, so the body
    // can be built as two launch plans (rows [0,B) and [B,2B)) that run on forked streams inside the captured graph,
    // one half's GEMM fill / drain, norms and attention overlapping the other half's tensor work (CFGPP_SPLIT=1). That
    // won 4 % while a GEMM launch lost ~10 us to fill / drain / epilogue; since the issue-loop and epilogue rewrites the
    // full-batch kernels are faster per row and the single plan is ahead by ~1 % (tools/ab_split.sh), so it is the
    // default.
    const int HW0 = H_ * W_;
    Act h0{g_dry ? nullptr : alloc_act(static_cast<size_t>(NB_) * HW0 * C0), C0};
    if (g_dry) workspace_bytes_ += 2 * static_cast<size_t>(NB_) * HW0 * C0 * sizeof(__half);
    __half* final_h = g_dry ? nullptr : alloc_act(static_cast<size_t>(NB_) * HW0 * C0);
    n_branches_ = (NB_ >= 2 && !split_disabled()) ? 2 : 1;
    for (int br = 0; br < n_branches_; ++br) {
      bnb_ = NB_ / n_branches_;
      brow0_ = br * bnb_;
      btag_ = "b" + std::to_string(br) + ".";
      bgn_partial_ = g_dry ? nullptr : gn_partial_ + static_cast<size_t>(br) * bnb_ * 128 * 64;
      cur_plan_ = &branch_plan_[br];
      int H = H_, W = W_;
      Act h{g_dry ? nullptr : h0.p + static_cast<size_t>(brow0_) * HW0 * C0, C0};
      std::vector<Act> skips{h};
      for (int i = 0; i < L; ++i) {
        const std::string blk = "down_blocks." + std::to_string(i);
        const int Cout = d_.block_out_channels[i];
        for (int j = 0; j < d_.layers_per_block; ++j) {
          const std::string rp = blk + ".resnets." + std::to_string(j);
          h = build_resnet(rp, h, nullptr, Cout, H, W, temb_off(rp));
          if (d_.down_has_attn[i])
            h = build_transformer(blk + ".attentions." + std::to_string(j), h, H, W, d_.transformer_layers[i],
                                  d_.num_heads[i]);
          skips.push_back(h);
        }
        if (i != L - 1) {
          h = build_downsample(blk + ".downsamplers.0", h, H, W);
          H /= 2; W /= 2;
          skips.push_back(h);
        }
      }
      {
        const int Cm = d_.block_out_channels[L - 1];
        h = build_resnet("mid_block.resnets.0", h, nullptr, Cm, H, W, temb_off("mid_block.resnets.0"));
        h = build_transformer("mid_block.attentions.0", h, H, W, d_.transformer_layers[L - 1], d_.num_heads[L - 1]);
        h = build_resnet("mid_block.resnets.1", h, nullptr, Cm, H, W, temb_off("mid_block.resnets.1"));
      }
      for (int i = 0; i < L; ++i) {
        const std::string blk = "up_blocks." + std::to_string(i);
        const int rev = L - 1 - i;
        const int Cout = d_.block_out_channels[rev];
        for (int j = 0; j < d_.layers_per_block + 1; ++j) {
          Act skip = skips.back();
          skips.pop_back();
          const std::string rp = blk + ".resnets." + std::to_string(j);
          const bool last = (i == L - 1) && (j == d_.layers_per_block) && !d_.up_has_attn[i];
          if (last && !g_dry) out_override_ = final_h + static_cast<size_t>(brow0_) * HW0 * C0;
          h = build_resnet(rp, h, &skip, Cout, H, W, temb_off(rp));
          if (d_.up_has_attn[i])
            h = build_transformer(blk + ".attentions." + std::to_string(j), h, H, W, d_.transformer_layers[rev],
                                  d_.num_heads[rev]);
        }
        if (i != L - 1) {
          h = build_upsample(blk + ".upsamplers.0", h, H, W);
          H *= 2; W *= 2;
        }
      }
      CFGPP_REQUIRE(skips.empty() && H == H_ && W == W_, "internal: skip stack mismatch");
      if (!g_dry && h.p != final_h + static_cast<size_t>(brow0_) * HW0 * C0) {
        // the last block ended with a transformer: gather its output into the shared tail input
        const __half* src = h.p;
        __half* dst = final_h + static_cast<size_t>(brow0_) * HW0 * C0;
        const size_t bytes = static_cast<size_t>(bnb_) * HW0 * C0 * sizeof(__half);
        add_step(btag_ + "gather_tail", [=](cudaStream_t st) {
          CFGPP_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
        });
      }
    }
    // ---- tail: conv_norm_out + SiLU over the whole batch feeds the fused conv_out / CFG++ step kernel ----
    cur_plan_ = &tail_plan_;
    bnb_ = NB_; brow0_ = 0; btag_ = "tail."; bgn_partial_ = gn_partial_;
    Scratch* s_norm = scratch("tail.norm", static_cast<size_t>(NB_) * HW0 * C0);
    if (!g_dry) {
      const __half *g = plain("conv_norm_out.weight"), *b = plain("conv_norm_out.bias");
      const int NB = NB_, HW = HW0;
      const float eps = d_.norm_eps;
      float* partial = gn_partial_;
      __half* normp = s_norm->p;
      const __half* hp = final_h;
      add_step("conv_norm_out+silu", [=](cudaStream_t st) {
        run_groupnorm(hp, C0, nullptr, 0, NB, HW, g, b, eps, true, partial, normp, st);
      }, 2);
      final_norm_ = Act{normp, C0};
      conv_in_out_ = h0.p;
    }
    if (g_dry) {
      // discard everything the sizing pass pushed (it pushes nothing) and keep the scratch sizes
      prologue_plan_.clear(); branch_plan_[0].clear(); branch_plan_[1].clear(); tail_plan_.clear();
      prompt_plan_.clear();
    }
  }
  g_dry = false;

  // FLOP / launch accounting (the reference executes the K/V projections every step: count them per forward)
  forward_flops_ = 0.0;
  launches_per_step_ = 3;  // select_step + conv_in + conv_out_step
  for (auto* pl : {&branch_plan_[0], &branch_plan_[1], &tail_plan_})
    for (auto& s : *pl) { forward_flops_ += s.flops; launches_per_step_ += s.launches; }
  for (auto& s : prologue_plan_) launches_per_step_ += s.launches;
  prompt_flops_ = 0.0;
  prompt_launches_ = 0;
  for (auto& s : prompt_plan_) { prompt_flops_ += s.flops; prompt_launches_ += s.launches; }
  forward_flops_ += prompt_flops_;
  const double px = static_cast<double>(NB_) * H_ * W_;
  forward_flops_ += 2.0 * px * (36.0 * C0 + 36.0 * C0);  // conv_in + conv_out
  forward_flops_ += 2.0 * NB_ * (static_cast<double>(C0) * TE + static_cast<double>(TE) * TE +
                                 static_cast<double>(temb_total_) * TE);
  if (has_aug_) {  // the add-embedding MLP runs in the prompt plan
    const double f = 2.0 * NB_ * (static_cast<double>(d_.projection_class_embeddings_input_dim) * TE +
                                  static_cast<double>(TE) * TE);
    forward_flops_ += f;
    prompt_flops_ += f;
  }
  CFGPP_CHECK_CUDA(cudaDeviceSynchronize());
  prepared_ = true;
}

void Unet::run_plan(const std::vector<PlanStep>& plan, cudaStream_t stream) {
  for (const auto& s : plan) s.fn(stream);
}

// Body of the forward. `concurrent`: fork the two CFG halves onto a second stream (used under graph capture, where
// the fork / join become graph edges); otherwise the halves run back to back on the caller's stream.
void Unet::run_body(cudaStream_t stream, bool concurrent) {
  if (n_branches_ == 2 && concurrent) {
    CFGPP_CHECK_CUDA(cudaEventRecord(fork_ev_, stream));
    CFGPP_CHECK_CUDA(cudaStreamWaitEvent(capture_stream2_, fork_ev_, 0));
    run_plan(branch_plan_[0], stream);
    run_plan(branch_plan_[1], capture_stream2_);
    CFGPP_CHECK_CUDA(cudaEventRecord(join_ev_, capture_stream2_));
    CFGPP_CHECK_CUDA(cudaStreamWaitEvent(stream, join_ev_, 0));
  } else {
    for (int br = 0; br < n_branches_; ++br) run_plan(branch_plan_[br], stream);
  }
  run_plan(tail_plan_, stream);
}

bool Unet::split_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CFGPP_SPLIT");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

// ------------------------------------------------------------------------------------------------------------
// conditioning
// ------------------------------------------------------------------------------------------------------------
void Unet::set_prompt(const __half* ctx, int n_ctx, const __half* pooled, const float* time_ids, int add_rows,
                      cudaStream_t stream) {
  CFGPP_REQUIRE(prepared_, "call cfgpp_prepare first");
  CFGPP_REQUIRE(n_ctx == n_ctx_, "context length differs from the prepared plan (77)");
  CFGPP_CHECK_CUDA(cudaMemcpyAsync(ctx_copy_, ctx, static_cast<size_t>(NB_) * n_ctx_ * d_.cross_attention_dim * 2,
                                   cudaMemcpyDeviceToDevice, stream));
  if (has_aug_) {
    CFGPP_REQUIRE(pooled && time_ids, "SDXL add-embedding needs pooled text embeds and time ids");
    CFGPP_REQUIRE(add_rows == NB_ || add_rows == B_, "add_rows must be batch or 2*batch");
    // broadcast rows r -> r % add_rows (the un-duplicated case of latent_sdxl.py:249-252)
    for (int r0 = 0; r0 < NB_; r0 += add_rows) {
      CFGPP_CHECK_CUDA(cudaMemcpyAsync(pooled_copy_ + static_cast<size_t>(r0) * d_.pooled_dim, pooled,
                                       static_cast<size_t>(add_rows) * d_.pooled_dim * 2, cudaMemcpyDeviceToDevice,
                                       stream));
      CFGPP_CHECK_CUDA(cudaMemcpyAsync(time_ids_copy_ + static_cast<size_t>(r0) * 6, time_ids,
                                       static_cast<size_t>(add_rows) * 6 * sizeof(float), cudaMemcpyDeviceToDevice,
                                       stream));
    }
  }
  run_plan(prompt_plan_, stream);
}

// ------------------------------------------------------------------------------------------------------------
// un-fused forward == predict_noise
// ------------------------------------------------------------------------------------------------------------
void Unet::unet_forward(const void* z, int z_dtype, float t, float in_scale, __half* eps_uc, __half* eps_c,
                        cudaStream_t stream) {
  CFGPP_REQUIRE(prepared_, "call cfgpp_prepare first");
  StepState s{};
  s.t = t;
  s.in_scale = in_scale;
  // cudaMemcpyAsync from pageable memory stages the 48 bytes before returning: `s` may go out of scope
  CFGPP_CHECK_CUDA(cudaMemcpyAsync(cur_state_, &s, sizeof(s), cudaMemcpyHostToDevice, stream));
  run_plan(prologue_plan_, stream);
  run_conv_in(z, z_dtype == CFGPP_F16 ? 1 : 0, &cur_state_->in_scale, conv_in_w_, conv_in_b_, conv_in_out_, B_, H_, W_,
              d_.block_out_channels[0], 2, stream);
  run_body(stream, true);  // the two CFG halves fork onto a side stream and join before the tail
  run_conv_out_step(final_norm_.p, conv_out_w_, conv_out_b_, B_, H_, W_, final_norm_.C, STEP_NONE, nullptr, nullptr,
                    nullptr, nullptr, eps_uc, eps_c, stream);
}

std::vector<Unet::ProfEntry> Unet::profile_forward(const void* z, int z_dtype, float t, float in_scale,
                                                   cudaStream_t stream) {
  CFGPP_REQUIRE(prepared_, "call cfgpp_prepare first");
  std::vector<ProfEntry> out;
  std::vector<cudaEvent_t> evs;
  auto mark = [&]() {
    cudaEvent_t e;
    CFGPP_CHECK_CUDA(cudaEventCreate(&e));
    CFGPP_CHECK_CUDA(cudaEventRecord(e, stream));
    evs.push_back(e);
  };
  StepState s{};
  s.t = t;
  s.in_scale = in_scale;
  CFGPP_CHECK_CUDA(cudaMemcpyAsync(cur_state_, &s, sizeof(s), cudaMemcpyHostToDevice, stream));
  mark();
  for (const auto& st : prologue_plan_) {
    st.fn(stream);
    mark();
    out.push_back({st.name, st.kind, st.flops, 0.f});
  }
  run_conv_in(z, z_dtype == CFGPP_F16 ? 1 : 0, &cur_state_->in_scale, conv_in_w_, conv_in_b_, conv_in_out_, B_, H_, W_,
              d_.block_out_channels[0], 2, stream);
  mark();
  out.push_back({"conv_in", 3, 2.0 * NB_ * H_ * W_ * 36.0 * d_.block_out_channels[0], 0.f});
  for (auto* pl : {&branch_plan_[0], &branch_plan_[1], &tail_plan_})
    for (const auto& st : *pl) {
      st.fn(stream);
      mark();
      out.push_back({st.name, st.kind, st.flops, 0.f});
    }
  run_conv_out_step(final_norm_.p, conv_out_w_, conv_out_b_, B_, H_, W_, final_norm_.C, STEP_NONE, nullptr, nullptr,
                    nullptr, nullptr, fwd_eps_uc_, fwd_eps_c_, stream);
  mark();
  out.push_back({"conv_out+step", 3, 2.0 * NB_ * H_ * W_ * 36.0 * d_.block_out_channels[0], 0.f});
  CFGPP_CHECK_CUDA(cudaStreamSynchronize(stream));
  for (size_t i = 0; i < out.size(); ++i) CFGPP_CHECK_CUDA(cudaEventElapsedTime(&out[i].ms, evs[i], evs[i + 1]));
  for (auto e : evs) cudaEventDestroy(e);
  return out;
}

// ------------------------------------------------------------------------------------------------------------
// fused trajectory
// ------------------------------------------------------------------------------------------------------------
void Unet::set_schedule(int method, int state_dtype, const cfgpp_step_state* steps, int nsteps, cudaStream_t stream) {
  CFGPP_REQUIRE(prepared_, "call cfgpp_prepare first");
  CFGPP_REQUIRE(method >= CFGPP_STEP_DDIM_CFGPP && method <= CFGPP_STEP_DDIM_CFG, "unknown method");
  CFGPP_REQUIRE(nsteps >= 1 && nsteps <= 1024, "nsteps must be 1..1024");
  static_assert(sizeof(cfgpp_step_state) == sizeof(StepState), "ABI struct mismatch");
  static_assert(sizeof(cfgpp_step_coef) == sizeof(StepCoef), "ABI struct mismatch");
  if (method != method_ || state_dtype != state_dtype_) graph_valid_ = false;
  method_ = method;
  state_dtype_ = state_dtype;
  nsteps_ = nsteps;
  steps_host_.assign(steps, steps + nsteps);
  CFGPP_CHECK_CUDA(cudaMemcpyAsync(step_table_, steps_host_.data(), sizeof(StepState) * nsteps, cudaMemcpyHostToDevice,
                                   stream));
}

void Unet::set_state(const void* z, int z_dtype, cudaStream_t stream) {
  CFGPP_REQUIRE(prepared_, "call cfgpp_prepare first");
  CFGPP_REQUIRE(z_dtype == state_dtype_, "state dtype differs from the schedule's state dtype");
  const size_t lat = static_cast<size_t>(B_) * 4 * H_ * W_;
  const size_t es = (state_dtype_ == CFGPP_F16) ? 2 : 4;
  CFGPP_CHECK_CUDA(cudaMemcpyAsync(z_state_, z, lat * es, cudaMemcpyDeviceToDevice, stream));
}

void Unet::set_noise(const __half* noise, int slots, cudaStream_t stream) {
  CFGPP_REQUIRE(prepared_, "call cfgpp_prepare first");
  CFGPP_REQUIRE(noise != nullptr && slots >= 1 && slots <= 1024, "noise table: 1..1024 slots");
  const size_t n = static_cast<size_t>(slots) * B_ * 4 * H_ * W_;
  if (n > noise_cap_) {
    CFGPP_CHECK_CUDA(cudaStreamSynchronize(stream));  // a replay in flight may still read the old table
    if (noise_buf_) cudaFree(noise_buf_);
    noise_buf_ = nullptr;
    noise_cap_ = 0;
    CFGPP_CHECK_CUDA(cudaMalloc(&noise_buf_, n * sizeof(__half)));
    noise_cap_ = n;
    CFGPP_CHECK_CUDA(cudaMemcpyAsync(noise_slot_, &noise_buf_, sizeof(__half*), cudaMemcpyHostToDevice, stream));
    CFGPP_CHECK_CUDA(cudaStreamSynchronize(stream));  // &noise_buf_ is host memory of this object: do not let it race
  }
  CFGPP_CHECK_CUDA(cudaMemcpyAsync(noise_buf_, noise, n * sizeof(__half), cudaMemcpyDeviceToDevice, stream));
}

void Unet::ensure_graph(cudaStream_t stream) {
  if (graph_valid_) return;
  if (graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
  if (graph_) { cudaGraphDestroy(graph_); graph_ = nullptr; }
  const int mode = method_ | (state_dtype_ == CFGPP_F16 ? 0x100 : 0);
  CFGPP_CHECK_CUDA(cudaStreamBeginCapture(capture_stream_, cudaStreamCaptureModeRelaxed));
  try {
    run_select_step(step_table_, step_counter_, cur_state_, capture_stream_);
    run_plan(prologue_plan_, capture_stream_);
    run_conv_in(z_state_, state_dtype_ == CFGPP_F16 ? 1 : 0, &cur_state_->in_scale, conv_in_w_, conv_in_b_,
                conv_in_out_, B_, H_, W_, d_.block_out_channels[0], 2, capture_stream_);
    run_body(capture_stream_, true);
    run_conv_out_step(final_norm_.p, conv_out_w_, conv_out_b_, B_, H_, W_, final_norm_.C, mode, &cur_state_->coef,
                      z_state_, aux_state_, z0t_state_, nullptr, nullptr, capture_stream_, noise_slot_);
  } catch (...) {
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(capture_stream_, &g);
    if (g) cudaGraphDestroy(g);
    throw;
  }
  CFGPP_CHECK_CUDA(cudaStreamEndCapture(capture_stream_, &graph_));
  CFGPP_CHECK_CUDA(cudaGraphInstantiate(&graph_exec_, graph_, 0));
  graph_valid_ = true;
}

void Unet::run_steps(int first_step, int nsteps, cudaStream_t stream) {
  CFGPP_REQUIRE(prepared_ && nsteps_ > 0, "call cfgpp_set_schedule first");
  CFGPP_REQUIRE(first_step >= 0 && first_step + nsteps <= nsteps_, "step range outside the schedule");
  ensure_graph(stream);
  CFGPP_CHECK_CUDA(cudaMemcpyAsync(step_counter_, &first_step, sizeof(int), cudaMemcpyHostToDevice, stream));
  for (int i = 0; i < nsteps; ++i) CFGPP_CHECK_CUDA(cudaGraphLaunch(graph_exec_, stream));
}

void Unet::get_state(int which, void* out, cudaStream_t stream) {
  CFGPP_REQUIRE(prepared_, "call cfgpp_prepare first");
  const size_t lat = static_cast<size_t>(B_) * 4 * H_ * W_;
  const size_t es = (state_dtype_ == CFGPP_F16) ? 2 : 4;
  const void* src = which == 0 ? z_state_ : (which == 1 ? z0t_state_ : aux_state_);
  CFGPP_CHECK_CUDA(cudaMemcpyAsync(out, src, lat * es, cudaMemcpyDeviceToDevice, stream));
}

void Unet::apply_step(int step, const __half* eps_uc, const __half* eps_c, cudaStream_t stream) {
  CFGPP_REQUIRE(prepared_ && step >= 0 && step < nsteps_, "step outside the schedule");
  const int mode = method_ | (state_dtype_ == CFGPP_F16 ? 0x100 : 0);
  const int n = B_ * 4 * H_ * W_;
  run_step_only(eps_uc, eps_c, n, mode, &step_table_[step].coef, z_state_, aux_state_, z0t_state_, stream, noise_slot_);
}

}  // namespace cfgpp

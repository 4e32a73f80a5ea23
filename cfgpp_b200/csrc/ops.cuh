// cfgpp_b200 — host entry points of the non-GEMM kernels (norms, embeddings, conv_in / conv_out + fused CFG++ step,
// resampling helpers). All enqueue on the given stream, never synchronise, and are CUDA-graph capturable.
#pragma once
#include "host.h"

namespace cfgpp {

// ---- norm.cu -----------------------------------------------------------------------------------------------
int gn_px_per_block(int HW);
int gn_num_chunks(int HW);
size_t gn_partial_floats(int B, int HW);
// GroupNorm(32) over channel-concat [x1 | x2] (x2 may be null), NHWC, optional SiLU. out [B,HW,C1+C2] fp16.
void run_groupnorm(const __half* x1, int C1, const __half* x2, int C2, int B, int HW, const __half* gamma,
                   const __half* beta, float eps, bool silu, float* partial, __half* out, cudaStream_t stream);
void run_layernorm(const __half* x, int M, int C, const __half* gamma, const __half* beta, float eps, __half* out,
                   cudaStream_t stream);

// ---- elementwise.cu ----------------------------------------------------------------------------------------
// diffusers get_timestep_embedding(flip_sin_to_cos=True, shift 0): out[i, col_off + (cos | sin)], fp16.
// value i is read at vals[i * val_stride] and written to row i.
void run_sincos_embed(const float* vals, int val_stride, int n, int dim, __half* out, int ld, int col_off,
                      cudaStream_t stream);
// out[r, n] = fp16(acc + bias[n]) (+ addend[r, n] in fp16 arithmetic); optional out_silu (replace by SiLU) and
// out2 = SiLU(out) copy. R <= 16 rows; weight [N][K] fp16, K % 8 == 0.
void run_small_linear(const __half* in, int ld_in, const __half* w, const __half* bias, const __half* addend,
                      int ld_add, __half* out, int ld_out, __half* out2, int R, int N, int K, bool out_silu,
                      cudaStream_t stream);
// copy rows: dst[r, col_off + c] = src[r % src_rows, c]  (fp16) — used to assemble the add-embedding input
void run_copy_rows(const __half* src, int src_rows, int cols, __half* dst, int ld_dst, int col_off, int R,
                   cudaStream_t stream);

// conv_in 3x3 pad 1, Cin = 4: z [B,4,H,W] (fp32 or fp16 NCHW, optionally scaled by in_scale in fp16 arithmetic)
// -> NHWC fp16 [reps*B, H, W, Cout]; the same result is written `reps` times (uncond and cond halves share z).
// in_scale (device pointer, may be null): model input is z * (*in_scale) — the DPM++ `x * c_in` (latent_sdxl.py:901).
void run_conv_in(const void* z, int z_is_half, const float* in_scale, const __half* w /*[Cout][36]*/,
                 const __half* bias, __half* out, int B, int H, int W, int Cout, int reps, cudaStream_t stream);

enum StepMode : int {
  STEP_NONE = 0,        // only emit eps_uc / eps_c (the predict_noise seam)
  STEP_DDIM_CFGPP = 1,  // latent_diffusion.py:660-666, latent_sdxl.py:738-744 (fp32 state)
  STEP_DDIM_INV_CFGPP = 2,  // latent_diffusion.py:904-908 (fp32 state)
  STEP_DPMPP2M_CFGPP = 3,   // latent_sdxl.py:902-919 (fp16 state, keeps old_denoised)
  STEP_DDIM_CFG = 4,        // plain-CFG DDIM step / inversion step: Tweedie AND renoise with the guided eps
                            // (latent_diffusion.py:283-287, :176-177; latent_sdxl.py:451-455, :321-322)
};

struct StepCoef {  // per-step scalars, computed on the host in fp32 exactly as the reference does
  float lambda;    // cfg_guidance
  float c0, c1, c2, c3;  // DDIM: sqrt(1-at), sqrt(at), sqrt(at_next), sqrt(1-at_next)
                         // DDIM-inv: sqrt(1-at_prev), sqrt(at_prev), sqrt(at), sqrt(1-at)
                         // DPM++: c_out(-sigma_i), 1/sigma_i, sigma_{i+1}, unused
  float d0, d1, d2;      // DPM++ 2M branch: -exp(-h), expm1(-h), 1/(2r) ; d3 = exp(-h)
  float d3;
  int second_order;      // DPM++ / Euler family bits: 1 = 2M update (else Euler), 2 = extrapolate with the guided
                         // estimate (plain CFG), 4 = 2M difference term on the guided estimate (SD v1.5 dpm++_2m_cfg++),
                         // 8 = ancestral noise: + noise[slot c3] * d3 (sigma_up), 16 / 32 = midpoint / final call of a
                         // DPM-Solver++(2S) step (16: d0 = sigma_s / sigma_t, d1 = expm1(-h r); 32: d0 = exp(-h),
                         // d1 = sigma_down / sigma_t, d2 = expm1(-h))
};

// One sampler step's device-resident scalars; a table of these lives in HBM and a 1-thread kernel selects the
// current entry, so a single CUDA graph replays for every step without host involvement.
struct StepState {
  float t;         // timestep fed to the UNet
  float in_scale;  // c_in (1.0 for DDIM)
  StepCoef coef;
};
void run_select_step(const StepState* table, int* counter, StepState* cur, cudaStream_t stream);

// conv_out 3x3 (Cin -> 4) on the GroupNorm+SiLU'ed NHWC input x [2B,H,W,Cin] fused with the CFG++ guidance mix and
// the scheduler update. noise_slot (may be null): device word holding the base of the ancestral noise table
// [slots][B,4,H,W] fp16. z is the sampler state (NCHW, fp32 for DDIM modes, fp16 for DPM++), updated in place.
void run_conv_out_step(const __half* x, const __half* w /*[4][9][Cin]*/, const __half* bias, int B, int H, int W,
                       int Cin, int mode, const StepCoef* coef_dev, void* z, void* aux /*old_denoised*/,
                       void* z0t_out, __half* eps_uc, __half* eps_c, cudaStream_t stream,
                       const __half* const* noise_slot = nullptr);

// standalone fused CFG++ update from given eps (used when a per-step callback needs the un-fused seam)
void run_step_only(const __half* eps_uc, const __half* eps_c, int n, int mode, const StepCoef* coef_dev, void* z,
                   void* aux, void* z0t_out, cudaStream_t stream, const __half* const* noise_slot = nullptr);

void run_upsample2x(const __half* x, __half* out, int B, int H, int W, int C, cudaStream_t stream);
// stride-2 pad-1 3x3 im2col: x [B,H,W,C] -> out [B*(H/2)*(W/2), 9*C] (tap-major, matches the packed weight)
void run_im2col_s2(const __half* x, __half* out, int B, int H, int W, int C, cudaStream_t stream);

}  // namespace cfgpp

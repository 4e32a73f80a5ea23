/* cfgpp_b200 — C ABI of the Blackwell-native CFG++ sampling hot path (libcfgpp_b200.so).
 *
 * The reference (CFGpp-diffusion/CFGpp) has no FFI: its extension point is the Python solver registry
 * (latent_diffusion.py:13-26, latent_sdxl.py:15-28) and, inside a solver, the seam
 *     predict_noise(zt, t, uc, c[, added_cond_kwargs]) -> (eps_uc, eps_c)          latent_diffusion.py:131-158
 *                                                                                 latent_sdxl.py:167-185
 * which calls diffusers' UNet2DConditionModel.forward, followed by the hand-written CFG++ update of each solver
 * (latent_diffusion.py:660-666, 904-908; latent_sdxl.py:738-744, 902-919). This library replaces exactly that:
 * the batched (uncond+cond) UNet forward plus the guidance mix and scheduler update, as hand-written sm_100a CUDA.
 * The Python mirror of the solver API (cfgpp_b200/latent_diffusion.py, latent_sdxl.py) binds these symbols with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: every function returns 0 on success or a negative status; cfgpp_last_error() returns the message of
 * the calling thread's last failure. No C++ exception crosses the boundary. All pointers named *_dev are CUDA
 * device pointers owned by the caller; the library owns only the opaque handle, its packed-weight arena and its
 * activation workspace. All work is enqueued asynchronously on the given cudaStream_t (passed as void*), with no
 * internal host synchronisation on the hot path. A handle is not thread-safe; use one per (process, GPU).
 * Layouts at the boundary are the reference's: latents NCHW contiguous (fp32 or fp16), context (rows,77,D) fp16.
 */
#ifndef CFGPP_B200_H_
#define CFGPP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFGPP_MAX_LEVELS 4

/* UNet2DConditionModel structure (diffusers config fields). */
typedef struct cfgpp_model_desc {
  int in_channels;                         /* 4 */
  int out_channels;                        /* 4 */
  int num_levels;                          /* len(block_out_channels): 4 (SD v1.5) / 3 (SDXL) */
  int block_out_channels[CFGPP_MAX_LEVELS];
  int down_has_attn[CFGPP_MAX_LEVELS];     /* CrossAttnDownBlock2D -> 1, DownBlock2D -> 0 */
  int up_has_attn[CFGPP_MAX_LEVELS];       /* in up_blocks order */
  int layers_per_block;                    /* 2 */
  int transformer_layers[CFGPP_MAX_LEVELS];/* per down level; mid uses the last; up uses reversed */
  int num_heads[CFGPP_MAX_LEVELS];         /* diffusers `attention_head_dim` (= number of heads) */
  int cross_attention_dim;                 /* 768 / 2048 */
  int use_linear_projection;               /* 0: 1x1-conv proj_in/out (SD v1.5), 1: Linear (SDXL) */
  int norm_num_groups;                     /* 32 */
  float norm_eps;                          /* 1e-5 (Transformer2DModel's GroupNorm uses 1e-6) */
  int addition_time_embed_dim;             /* 0: no add-embedding; 256: SDXL text_time */
  int projection_class_embeddings_input_dim; /* 2816 */
  int pooled_dim;                          /* 1280 */
} cfgpp_model_desc;

/* dtype codes */
#define CFGPP_F16 0
#define CFGPP_F32 1

/* sampler update fused behind the UNet (cfgpp_set_schedule `method`) */
#define CFGPP_STEP_NONE 0            /* no update: expose eps_uc / eps_c only */
#define CFGPP_STEP_DDIM_CFGPP 1      /* ddim_cfg++ (+_lightning)  latent_diffusion.py:660-666, latent_sdxl.py:738-744 */
#define CFGPP_STEP_DDIM_INV_CFGPP 2  /* inversion of ddim_inversion_cfg++  latent_diffusion.py:904-908 */
#define CFGPP_STEP_DPMPP2M_CFGPP 3   /* dpm++_2m_cfgpp  latent_sdxl.py:902-919 */
#define CFGPP_STEP_DDIM_CFG 4        /* plain-CFG ddim step and inversion step (baselines)  latent_diffusion.py:283-287 */

/* Per-step scalars, computed by the host exactly as the reference computes them (fp32 torch CPU ops). */
typedef struct cfgpp_step_coef {
  float lambda_;         /* cfg_guidance */
  float c0, c1, c2, c3;  /* DDIM: sqrt(1-a_t), sqrt(a_t), sqrt(a_next), sqrt(1-a_next)
                            DDIM-inv: sqrt(1-a_prev), sqrt(a_prev), sqrt(a_t), sqrt(1-a_t)
                            DPM++2M: c_out = -sigma_i, 1/sigma_i, sigma_{i+1}, unused */
  float d0, d1, d2, d3;  /* DPM++2M 2nd-order branch: -exp(-h), expm1(-h), 1/(2r), exp(-h) */
  int second_order;      /* VE-cast family bits: 1 = 2M update (else the Euler-CFG++ update: first step / sigma_next == 0 /
                            euler solvers), 2 = extrapolate with the guided estimate (plain-CFG euler, dpm++_2m),
                            4 = 2M difference term on the guided estimate (SD v1.5 dpm++_2m_cfg++),
                            8 = ancestral (euler_a, dpm++_2s_a: latent_diffusion.py:757-760, :823): after the update add
                                noise[slot c3] * d3 (sigma_up), the table comes from cfgpp_set_noise,
                            16 / 32 = the two UNet calls of a DPM-Solver++(2S) step (latent_diffusion.py:796-821), one
                                schedule entry each: 16 = midpoint, d0 = sigma_s / sigma_t, d1 = expm1(-h r) (x is parked,
                                the state becomes x_2); 32 = final, d0 = exp(-h), d1 = sigma_down / sigma_t,
                                d2 = expm1(-h) (plain-CFG form when bit 2 is set) */
} cfgpp_step_coef;

typedef struct cfgpp_step_state {
  float t;         /* timestep fed to the UNet (DDIM: t; DPM++: sigma_to_t(sigma_i) = t-1) */
  float in_scale;  /* model input scale c_in (1.0 for DDIM) */
  cfgpp_step_coef coef;
} cfgpp_step_state;

typedef struct cfgpp_handle cfgpp_handle;

int cfgpp_version(void);
const char* cfgpp_last_error(void);

/* ---- model lifetime: replaces pipe.unet obtained at latent_diffusion.py:67 / latent_sdxl.py:50,391 ---------- */
int cfgpp_create(const cfgpp_model_desc* desc, int device, cfgpp_handle** out);
int cfgpp_destroy(cfgpp_handle* h);
/* One call per state-dict entry under its diffusers key, fp16 or fp32 device tensor. */
int cfgpp_load_weight(cfgpp_handle* h, const char* diffusers_key, const void* data_dev, const int64_t* shape, int ndim,
                      int dtype, void* stream);
/* Repack into kernel-native layouts (conv [Cout][9][Cin], fused QKV / KV, interleaved GEGLU, concatenated
 * time_emb_proj). Fails listing the first missing key if the state dict is incomplete. */
int cfgpp_finalize_weights(cfgpp_handle* h, void* stream);
/* Build the launch plan + workspace for `batch` images of latent size (h_lat, w_lat); UNet batch is 2*batch. */
int cfgpp_prepare(cfgpp_handle* h, int batch, int h_lat, int w_lat);
int cfgpp_workspace_bytes(cfgpp_handle* h, size_t* bytes);
/* Algorithmic FLOPs (2*MACs of conv/linear/QK^T/PV as the reference executes them) of one 2*batch UNet forward,
 * and the number of kernel launches one fused step enqueues. */
int cfgpp_forward_flops(cfgpp_handle* h, double* flops);
int cfgpp_launches_per_step(cfgpp_handle* h, int* n);
/* Accounting of the prepared plan, all per 2*batch UNet forward: step_flops = FLOPs the fused step EXECUTES every step
 * (excludes what runs once per prompt); prompt_flops / prompt_launches = the once-per-set_prompt part (cross-attention
 * K/V projections, SDXL add-embedding). cfgpp_forward_flops == step_flops + prompt_flops is the reference-equivalent
 * algorithmic figure (diffusers recomputes the K/V projections every step). */
int cfgpp_plan_stats(cfgpp_handle* h, double* step_flops, double* prompt_flops, int* prompt_launches);

/* ---- per-prompt conditioning: the tensors predict_noise concatenates (latent_sdxl.py:178-182, 249-257) ------ */
/* ctx_dev: (2*batch, 77, cross_dim) fp16 = cat([uc, c]); pooled_dev: (add_rows, pooled_dim) fp16;
 * time_ids_dev: (add_rows, 6) fp32; add_rows is 2*batch, or batch when the reference does not duplicate the added
 * conditions (cfg_guidance in {0,1}: latent_sdxl.py:249-252 — rows then broadcast over both halves).
 * pooled/time_ids are ignored (may be NULL) for models without add-embedding. n_ctx = tokens per row (77). */
int cfgpp_set_prompt(cfgpp_handle* h, const void* ctx_dev, int n_ctx, const void* pooled_dev, const float* time_ids_dev,
                     int add_rows, void* stream);

/* ---- un-fused seam == predict_noise ------------------------------------------------------------------------- */
/* z_dev: (batch,4,h,w) NCHW of z_dtype; model input is z * in_scale; outputs (batch,4,h,w) fp16 each. */
int cfgpp_unet_forward(cfgpp_handle* h, const void* z_dev, int z_dtype, float t, float in_scale, void* eps_uc_dev,
                       void* eps_c_dev, void* stream);

/* Profiling aid: the same un-fused forward with a CUDA-event pair around every plan entry. Arrays are caller-owned
 * host buffers of max_n entries; kind: 0 linear GEMM, 1 conv3x3, 2 attention, 3 other; names_host holds max_n
 * NUL-terminated strings of name_stride bytes each (may be NULL). Synchronises the stream (not a hot-path call). */
int cfgpp_profile_forward(cfgpp_handle* h, const void* z_dev, int z_dtype, float t, float in_scale, int max_n,
                          int* n_out, float* ms_host, double* flops_host, int* kind_host, char* names_host,
                          int name_stride, void* stream);

/* ---- fused trajectory: UNet + CFG++ mix + scheduler update per step, one CUDA graph replayed per step -------- */
int cfgpp_set_schedule(cfgpp_handle* h, int method, int state_dtype, const cfgpp_step_state* steps_host, int nsteps,
                       void* stream);
/* Copy the caller's initial state into the library's state buffer (z: zT / x0; aux: old_denoised or NULL). */
int cfgpp_set_state(cfgpp_handle* h, const void* z_dev, int z_dtype, void* stream);
/* Ancestral samplers: the trajectory's fresh noise, drawn up front in the order the reference's loop would draw it
 * (`torch.randn_like(x)` once per step with sigma_next > 0). noise_dev: fp16 [slots][batch,4,h,w]; copied. */
int cfgpp_set_noise(cfgpp_handle* h, const void* noise_dev, int slots, void* stream);
/* Run `nsteps` consecutive steps starting at schedule index `first_step` on the internal state. */
int cfgpp_run_steps(cfgpp_handle* h, int first_step, int nsteps, void* stream);
/* which: 0 = state z (same dtype as the state), 1 = z0t of the last executed step. */
int cfgpp_get_state(cfgpp_handle* h, int which, void* out_dev, void* stream);
/* Standalone update from caller-provided eps (callback path: the caller may have modified nothing, it just needs
 * z0t / zt materialised between UNet calls). Applies schedule entry `step` to the internal state. */
int cfgpp_apply_step(cfgpp_handle* h, int step, const void* eps_uc_dev, const void* eps_c_dev, void* stream);

/* ---- AutoencoderKL decoder: replaces `self.vae.decode(zt / scaling_factor).sample` of
 * latent_sdxl.py:155-164 (VAE madebyollin/sdxl-vae-fp16-fix, :44) and latent_diffusion.py:123-129 on the same conv /
 * GEMM / GroupNorm kernels. Weights under the diffusers AutoencoderKL keys (`post_quant_conv.*`, `decoder.*`). ----- */
typedef struct cfgpp_vae_desc {
  int latent_channels;                       /* 4 */
  int out_channels;                          /* 3 */
  int num_levels;                            /* len(block_out_channels): 4 */
  int block_out_channels[CFGPP_MAX_LEVELS];  /* (128, 256, 512, 512) */
  int layers_per_block;                      /* 2 (the decoder's up blocks hold layers_per_block + 1 resnets) */
  int norm_num_groups;                       /* 32 */
  float scaling_factor;                      /* 0.13025 (SDXL) / 0.18215 (SD v1.5) */
} cfgpp_vae_desc;
typedef struct cfgpp_vae_handle cfgpp_vae_handle;
int cfgpp_vae_create(const cfgpp_vae_desc* desc, int device, cfgpp_vae_handle** out);
int cfgpp_vae_destroy(cfgpp_vae_handle* h);
int cfgpp_vae_load_weight(cfgpp_vae_handle* h, const char* diffusers_key, const void* data_dev, const int64_t* shape,
                          int ndim, int dtype, void* stream);
int cfgpp_vae_finalize_weights(cfgpp_vae_handle* h, void* stream);
/* zt_dev: (batch,4,h,w) NCHW of z_dtype — the SCALED latent the samplers return (the division by scaling_factor
 * happens inside, in zt's dtype, as in the reference); image_dev: (batch,3,8h,8w) NCHW fp16 (num_levels = 4).
 * The plan / workspace for (batch,h,w) is built on first use and cached. */
int cfgpp_vae_decode(cfgpp_vae_handle* h, const void* zt_dev, int z_dtype, int batch, int h_lat, int w_lat,
                     void* image_dev, void* stream);
/* ENCODER half — replaces `self.vae.encode(x).latent_dist.sample() * scaling_factor` (latent_sdxl.py:151-152,
 * latent_diffusion.py:117-121), the front end of the inversion / editing solvers. Needs the `encoder.*` and
 * `quant_conv.*` weights to have been loaded. image_dev: (batch,3,H,W) NCHW of image_dtype in [-1, 1];
 * noise_dev: (batch,4,H/8,W/8) fp16 — the `randn` draw of DiagonalGaussianDistribution.sample, made by the caller —
 * or NULL for the posterior mean; latent_out: (batch,4,H/8,W/8) fp32 (the fp16 module's output under the reference's
 * autocast: `exp` promotes the posterior's std to fp32), already multiplied by scaling_factor. */
int cfgpp_vae_encode(cfgpp_vae_handle* h, const void* image_dev, int image_dtype, int batch, int height, int width,
                     const void* noise_dev, void* latent_out, void* stream);
/* Algorithmic FLOPs of one decode of the prepared shape, and its activation workspace. */
int cfgpp_vae_stats(cfgpp_vae_handle* h, double* flops, size_t* workspace_bytes);

/* ---- CLIP text encoder: replaces `self.text_encoder(ids)[0]` of latent_diffusion.py:93-115 and
 * `text_enc(ids, output_hidden_states=True)` -> `hidden_states[-2]` / `[-(clip_skip + 2)]` / `[0]` of
 * latent_sdxl.py:77-93 (transformers CLIPTextModel: openai/clip-vit-large-patch14; CLIPTextModelWithProjection:
 * OpenCLIP ViT-bigG). Weights under the transformers keys (`text_model.*`, `text_projection.weight`). Tokenisation
 * stays on the host (cfgpp_b200/tokenizer.py). ----- */
typedef struct cfgpp_clip_desc {
  int vocab_size;        /* 49408 */
  int max_positions;     /* 77 */
  int hidden_size;       /* 768 (CLIP-L) / 1280 (bigG); heads are 64 wide */
  int intermediate_size; /* 3072 / 5120 */
  int num_layers;        /* 12 / 32 */
  int num_heads;         /* 12 / 20 */
  int hidden_act;        /* 0 = quick_gelu (CLIP-L), 1 = gelu (bigG) */
  int projection_dim;    /* 0 = CLIPTextModel; > 0 = CLIPTextModelWithProjection (1280) */
  float layer_norm_eps;  /* 1e-5 */
} cfgpp_clip_desc;
typedef struct cfgpp_clip_handle cfgpp_clip_handle;
int cfgpp_clip_create(const cfgpp_clip_desc* desc, int device, cfgpp_clip_handle** out);
int cfgpp_clip_destroy(cfgpp_clip_handle* h);
int cfgpp_clip_load_weight(cfgpp_clip_handle* h, const char* transformers_key, const void* data_dev, const int64_t* shape,
                           int ndim, int dtype, void* stream);
int cfgpp_clip_finalize_weights(cfgpp_clip_handle* h, void* stream);
/* input_ids_dev: (batch, n_tokens) int32 token ids; pooled_index_dev: (batch) int32 row of the pooled token (the
 * <|endoftext|> position the model's eos rule selects; may be null when pooled_out is null). Outputs, fp16, each may
 * be null: hidden_out (batch, n_tokens, hidden) = hidden_states[num_layers - skip] (skip = 1: the penultimate layer SDXL
 * conditions on; skip = clip_skip + 1 in general); last_hidden_out = final_layer_norm(hidden_states[-1]) (what SD v1.5
 * conditions on); pooled_out (batch, projection_dim or hidden) = text_embeds / pooler_output. */
int cfgpp_clip_encode(cfgpp_clip_handle* h, const int32_t* input_ids_dev, const int32_t* pooled_index_dev, int batch,
                      int n_tokens, int skip, void* hidden_out, void* last_hidden_out, void* pooled_out, void* stream);
int cfgpp_clip_stats(cfgpp_clip_handle* h, double* flops, size_t* workspace_bytes);

/* ---- operator-level entry points (one kernel each; used by the kernel parity tests and micro-benchmarks) ----- */
int cfgpp_op_linear(const void* a, int lda, const void* a2, int lda2, int k_split, const void* w, int M, int N, int K,
                    const void* bias, const void* addend, int ld_add, int add_rows_per_group, void* out, int ldc,
                    int geglu, int force_bn, void* stream);
int cfgpp_op_conv3x3(const void* x, int B, int H, int W, int Cin, const void* w, int Cout, const void* bias,
                     const void* addend, int ld_add, int add_rows_per_group, void* out, int force_bn, void* stream);
/* Downsample2D: 3x3, stride 2 on NHWC x [B,H,W,Cin] (even H, W) -> [B,H/2,W/2,Cout]; the A tile is fetched by TMA with
 * element strides 2 (no im2col copy). pad = 1: the UNet's (symmetric zero padding); pad = 0: the AutoencoderKL encoder's
 * (one zero row / column AFTER the image, F.pad(x, (0, 1, 0, 1)) + un-padded convolution). */
int cfgpp_op_conv3x3_s2(const void* x, int B, int H, int W, int Cin, const void* w, int Cout, const void* bias, int pad,
                        void* out, void* stream);
/* head h of q / k / v / out occupies columns [h*P, h*P + head_dim) with P = head_dim rounded up to a multiple of 64
 * (columns head_dim..P-1 must be zero in q / k / v and come back zero in out). */
int cfgpp_op_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                       int H, int Nq, int Nkv, int head_dim, void* stream);
int cfgpp_op_groupnorm(const void* x1, int C1, const void* x2, int C2, int B, int HW, const void* gamma,
                       const void* beta, float eps, int silu, void* out, void* stream);
int cfgpp_op_layernorm(const void* x, int M, int C, const void* gamma, const void* beta, float eps, void* out,
                       void* stream);
/* noise_dev (may be null): fp16 ancestral-noise table [slots][n]; the slot is coef_host->c3 (second_order bit 8). */
int cfgpp_op_cfgpp_step(const void* eps_uc, const void* eps_c, int n, int method, int state_dtype,
                        const cfgpp_step_coef* coef_host, void* z, void* aux, void* z0t_out, const void* noise_dev,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CFGPP_B200_H_ */

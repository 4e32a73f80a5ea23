"""SDXL / SDXL-Lightning CFG++ solvers on the Blackwell-native backend.

Mirror of the reference's `latent_sdxl.py` solver API for the hot path named by BASELINE.json — same registry
(`register_solver` / `get_solver`, latent_sdxl.py:15-28), same class / method names and argument meaning
(`SDXL.sample` :200-266, `reverse_process` :715-755 / :843-858 / :864-930, `predict_noise` :167-185,
`initialize_latent` :268-299, `sigma_to_t` :333-346), same errors (ValueError for unknown / duplicate solver,
AssertionError for Lightning with cfg_guidance != 1, NotImplementedError for unknown init methods) — but the UNet
forward, the CFG++ guidance mix and the scheduler update run in hand-written sm_100a CUDA behind the C ABI
(include/cfgpp_b200.h). With `callback_fn=None` a whole trajectory is enqueued as NFE replays of one CUDA graph with
no host synchronisation; with a callback the un-fused seam (`predict_noise` + `apply_step`) is used so that `z0t` /
`zt` are materialised and may be replaced by the callback, exactly like the reference loop.

Registered here: ddim_cfg++, ddim_cfg++_lightning, dpm++_2m_cfgpp (the solvers of SURVEY.md §8a) and, from §8 f1,
dpm++_2m_cfgpp_lightning (:932-952), ddim_edit_cfg++ (:954-1025, both loops on the fused step modes), euler_cfg++ and
euler_cfg++_lightning (:757-836: fused VE-cast trajectories, kdiffusion.py; the op-by-op torch form only with a callback).
The VAE (decode and encode, vae.py) and the two CLIP text towers (text_encoder.py, §8 f3) run on the
native backend too; `text_encoders=` / `vae=` accept replacements.
"""
from __future__ import annotations

import warnings
from typing import Any, Optional, Tuple

import torch

from . import kdiffusion as K
from . import schedule as S
from .conditioning import SyntheticTextEncoder
from .text_encoder import CLIPTextConfig, ClipConditioner, get_conditioner
from .config import UNetConfig, sdxl_config
from .engine import NativeUNet
from .weights import load_safetensors_state_dict, synthetic_state_dict

####### Factory #######
__SOLVER__ = {}


def register_solver(name: str):
    def wrapper(cls):
        if __SOLVER__.get(name, None) is not None:
            raise ValueError(f"Solver {name} already registered.")
        __SOLVER__[name] = cls
        return cls
    return wrapper


def get_solver(name: str, **kwargs):
    if name not in __SOLVER__:
        raise ValueError(f"Solver {name} does not exist.")
    return __SOLVER__[name](**kwargs)

########################

_ENGINES = {}


def resolve_state_dict(model_key: str, cfg: UNetConfig, device):
    """`*.safetensors` path -> real weights; 'synthetic[:seed]' -> seeded synthetic; anything else (an HF hub id:
    nothing can be downloaded here) -> synthetic with a warning."""
    if model_key.endswith(".safetensors"):
        return load_safetensors_state_dict(model_key, device=device)
    seed = 1234
    if model_key.startswith("synthetic"):
        if ":" in model_key:
            seed = int(model_key.split(":", 1)[1])
    else:
        warnings.warn(f"no checkpoint for '{model_key}' is available offline; using seeded synthetic UNet weights "
                      f"(pass a diffusers-format *.safetensors path as model_key for real weights)")
    return synthetic_state_dict(cfg, seed=seed, device=device)


def get_engine(model_key: str, cfg: UNetConfig, device, state_dict=None) -> NativeUNet:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("cfgpp_b200 solvers run on CUDA (sm_100a) only — there is no CPU fallback on the product "
                           "path; the CPU eager baseline lives in oracle/ and bench.py")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (model_key, cfg.name, idx)
    ent = _ENGINES.get(key)
    # an explicit state_dict is part of the identity: the entry keeps a strong reference to it and compares with `is`
    # (an id() of a dead dict can be recycled by a different one)
    if ent is not None and ent[1] is state_dict:
        return ent[0]
    # same key, other weights: the cache entry is replaced; the old engine (5 GB) is destroyed by NativeUNet.__del__
    # as soon as the last solver holding it goes away
    sd = state_dict if state_dict is not None else resolve_state_dict(model_key, cfg, torch.device("cuda", idx))
    eng = NativeUNet(cfg, sd, torch.device("cuda", idx))
    _ENGINES[key] = (eng, state_dict)
    return eng


def release_engines():
    """Destroy every cached engine (packed weights + workspace) — the cache otherwise lives as long as the process."""
    for eng, _ in _ENGINES.values():
        eng.close()
    _ENGINES.clear()


class _Scheduler:
    """The two attributes of the diffusers scheduler object the reference touches."""
    def __init__(self, sch: S.Schedule, device):
        self.timesteps = sch.timesteps.to(device)
        self.alphas_cumprod = sch.alphas_cumprod
        self.final_alpha_cumprod = sch.final_alpha_cumprod


def default_text_encoders(cfg: UNetConfig, device):
    """(text_enc_1, text_enc_2) for a UNet config: CLIP-L + OpenCLIP bigG for the real SDXL widths (768 + 1280 = 2048,
    pooled 1280), proportionally narrow towers for the test-sized configs; widths that are not multiples of the 64-wide
    CLIP head fall back to the shape-only stand-in of conditioning.py."""
    d2 = cfg.pooled_dim
    d1 = cfg.cross_attention_dim - d2
    if d1 <= 0:
        d1 = cfg.cross_attention_dim // 2
        d2 = cfg.cross_attention_dim - d1
    if (d1, d2, cfg.pooled_dim) == (768, 1280, 1280):
        return (get_conditioner("clip_l", device, "sdxl"), get_conditioner("clip_bigg", device, "sdxl"))
    if d1 % 64 or d2 % 64 or cfg.pooled_dim % 8:
        return (SyntheticTextEncoder(d1, 0), SyntheticTextEncoder(d2, cfg.pooled_dim))

    def small(name, d, proj, act, pad):
        return CLIPTextConfig(name=name, vocab_size=1024, hidden_size=d, intermediate_size=4 * d, num_hidden_layers=2,
                              num_attention_heads=d // 64, hidden_act=act, projection_dim=proj, pad_token_id=pad)
    return (get_conditioner("", device, "sdxl", cfg=small(f"clip_{d1}", d1, 0, "quick_gelu", 1023)),
            get_conditioner("", device, "sdxl", cfg=small(f"clip_{d2}_proj", d2, cfg.pooled_dim, "gelu", 0)))


class SDXL(K.KDiffusionMixin):
    schedule_kind = "ddim"
    quantize = True

    def __init__(self,
                 solver_config,
                 model_key: str = "stabilityai/stable-diffusion-xl-base-1.0",
                 dtype=torch.float16,
                 device='cuda',
                 unet_config: Optional[UNetConfig] = None,
                 state_dict=None,
                 text_encoders=None,
                 vae=None):
        self.device = device
        self.dtype = dtype
        self.cfg = unet_config or sdxl_config()
        self.unet = get_engine(model_key, self.cfg, device, state_dict)

        # CLIP text towers on the native backend (text_encoder.py; the reference takes pipe.text_encoder /
        # pipe.text_encoder_2, latent_sdxl.py:46-49). Pass `text_encoders=(fn1, fn2)`, prompt -> (hidden, pooled), to override.
        self.text_enc_1, self.text_enc_2 = text_encoders or default_text_encoders(self.cfg, device)
        if vae is None:
            # AutoencoderKL decoder on the native backend (vae.py; the reference loads madebyollin/sdxl-vae-fp16-fix,
            # latent_sdxl.py:44). Pass `vae=` (any object with decode(zt) / encode(x, dtype)) to override.
            from .vae import get_vae
            vae = get_vae("sdxl_vae", device)
        self.vae = vae
        self.vae_scale_factor = self.cfg.vae_scale_factor
        self.default_sample_size = self.cfg.sample_size

        # sampling parameters (latent_sdxl.py:56-67 / :407-418)
        self._sch = S.Schedule.make(solver_config.num_sampling, self.schedule_kind)
        self.total_alphas = self._sch.total_alphas
        self.sigmas = self._sch.sigmas
        self.log_sigmas = self._sch.log_sigmas
        self.skip = self._sch.skip
        self.final_alpha_cumprod = self._sch.final_alpha_cumprod
        self.scheduler = _Scheduler(self._sch, device)

    def __call__(self, *args: Any, **kwargs: Any) -> Any:
        self.sample(*args, **kwargs)

    def alpha(self, t):
        at = self.scheduler.alphas_cumprod[t] if t >= 0 else self.final_alpha_cumprod
        return at

    @torch.no_grad()
    def _text_embed(self, prompt, text_enc, clip_skip):
        prompt = prompt[0] if isinstance(prompt, (list, tuple)) else prompt
        if isinstance(text_enc, ClipConditioner):
            return text_enc(prompt, self.device, clip_skip=clip_skip)
        return text_enc(prompt, self.device)

    @torch.no_grad()
    def get_text_embed(self, null_prompt_1, prompt_1, null_prompt_2=None, prompt_2=None, clip_skip=None):
        prompt_embed_1, pool_prompt_embed = self._text_embed(prompt_1, self.text_enc_1, clip_skip)
        if prompt_2 is None:
            prompt_embed = [prompt_embed_1]
        else:
            prompt_embed_2, pool_prompt_embed = self._text_embed(prompt_2, self.text_enc_2, clip_skip)
            prompt_embed = [prompt_embed_1, prompt_embed_2]
        null_embed_1, pool_null_embed = self._text_embed(null_prompt_1, self.text_enc_1, clip_skip)
        if null_prompt_2 is None:
            null_embed = [null_embed_1]
        else:
            null_embed_2, pool_null_embed = self._text_embed(null_prompt_2, self.text_enc_2, clip_skip)
            null_embed = [null_embed_1, null_embed_2]
        null_prompt_embeds = torch.concat(null_embed, dim=-1)
        prompt_embeds = torch.concat(prompt_embed, dim=-1)
        return null_prompt_embeds, prompt_embeds, pool_null_embed, pool_prompt_embed

    @torch.no_grad()
    def encode(self, x):
        return self.vae.encode(x, self.dtype)

    def decode(self, zt):
        return self.vae.decode(zt).float()

    # ---- the seam: batched (uncond + cond) UNet forward on the native backend -----------------------------------
    def _prepare(self, zt, uc, c, added_cond_kwargs, force: bool = False):
        b, _, h, w = zt.shape
        self.unet.prepare(b, h, w)
        if self.cfg.addition_embed_type == "text_time":
            self.unet.bind_prompt(uc, c, added_cond_kwargs['text_embeds'], added_cond_kwargs['time_ids'], force=force)
        else:
            self.unet.bind_prompt(uc, c, force=force)

    def predict_noise(self, zt, t, uc, c, added_cond_kwargs, in_scale: float = 1.0):
        if uc is None or c is None:
            # single-branch paths of the reference (latent_sdxl.py:169-176) are never taken by the CFG++ solvers
            uc = c if uc is None else uc
            c = uc if c is None else c
        self._prepare(zt, uc, c, added_cond_kwargs)
        return self.unet.predict_noise(zt, float(t), in_scale)

    def _get_add_time_ids(self, original_size, crops_coords_top_left, target_size, dtype, text_encoder_projection_dim):
        add_time_ids = list(original_size + crops_coords_top_left + target_size)
        passed_add_embed_dim = self.cfg.addition_time_embed_dim * len(add_time_ids) + text_encoder_projection_dim
        expected_add_embed_dim = self.cfg.projection_class_embeddings_input_dim
        assert expected_add_embed_dim == passed_add_embed_dim, (
            f"Model expects an added time embedding vector of length {expected_add_embed_dim}, but a vector of "
            f"{passed_add_embed_dim} was created. The model has an incorrect config.")
        return torch.tensor([add_time_ids], dtype=dtype)

    def sample(self,
               prompt1=["", ""],
               prompt2=["", ""],
               cfg_guidance: float = 5.0,
               original_size: Optional[Tuple[int, int]] = None,
               crops_coords_top_left: Tuple[int, int] = (0, 0),
               target_size: Optional[Tuple[int, int]] = None,
               negative_original_size: Optional[Tuple[int, int]] = None,
               negative_crops_coords_top_left: Tuple[int, int] = (0, 0),
               negative_target_size: Optional[Tuple[int, int]] = None,
               clip_skip: Optional[int] = None,
               **kwargs):
        height = self.default_sample_size * self.vae_scale_factor
        width = self.default_sample_size * self.vae_scale_factor
        original_size = original_size or (height, width)
        target_size = target_size or (height, width)

        (null_prompt_embeds, prompt_embeds, pool_null_embed, pool_prompt_embed) = self.get_text_embed(
            prompt1[0], prompt1[1], prompt2[0], prompt2[1], clip_skip)

        add_text_embeds = pool_prompt_embed
        add_time_ids = self._get_add_time_ids(original_size, crops_coords_top_left, target_size,
                                              dtype=prompt_embeds.dtype,
                                              text_encoder_projection_dim=int(pool_prompt_embed.shape[-1]))
        if negative_original_size is not None and negative_target_size is not None:
            negative_add_time_ids = self._get_add_time_ids(negative_original_size, negative_crops_coords_top_left,
                                                           negative_target_size, dtype=prompt_embeds.dtype,
                                                           text_encoder_projection_dim=int(pool_prompt_embed.shape[-1]))
        else:
            negative_add_time_ids = add_time_ids
        negative_text_embeds = pool_null_embed

        if cfg_guidance != 0.0 and cfg_guidance != 1.0:
            add_text_embeds = torch.cat([negative_text_embeds, add_text_embeds], dim=0)
            add_time_ids = torch.cat([negative_add_time_ids, add_time_ids], dim=0)

        add_cond_kwargs = {'text_embeds': add_text_embeds.to(self.device), 'time_ids': add_time_ids.to(self.device)}

        zt = self.reverse_process(null_prompt_embeds, prompt_embeds, cfg_guidance, add_cond_kwargs, target_size,
                                  **kwargs)
        with torch.no_grad():
            img = self.decode(zt)
        img = (img / 2 + 0.5).clamp(0, 1)
        return img.detach().cpu()

    def initialize_latent(self, method: str = 'random', src_img: Optional[torch.Tensor] = None,
                          add_cond_kwargs: Optional[dict] = None, **kwargs):
        if method == 'random':
            size = kwargs.get('size', (1, 4, 128, 128))
            z = torch.randn(size).to(self.device)  # CPU generator, then H2D — latent_sdxl.py:288-289
        elif method == 'random_kdiffusion':
            size = kwargs.get('latent_dim', (1, 4, 128, 128))
            sigmas = kwargs.get('sigmas', [14.6146])
            z = torch.randn(size).to(self.device)
            z = z * (sigmas[0] ** 2 + 1) ** 0.5
        elif method == 'ddim':
            assert src_img is not None, "src_img must be provided for inversion"
            z = self.inversion(self.encode(src_img.to(self.dtype).to(self.device)), kwargs.get('uc'), kwargs.get('c'),
                               kwargs.get('cfg_guidance', 0.0), add_cond_kwargs)
        elif method == 'npi':
            assert src_img is not None, "src_img must be provided for inversion"
            z = self.inversion(self.encode(src_img.to(self.dtype).to(self.device)), kwargs.get('c'), kwargs.get('c'),
                               1.0, add_cond_kwargs)
        else:
            raise NotImplementedError
        return z

    def reverse_process(self, *args, **kwargs):
        raise NotImplementedError

    @torch.no_grad()
    def inversion(self, z0, uc, c, cfg_guidance, add_cond_kwargs):
        """Plain-CFG DDIM inversion (latent_sdxl.py:301-324): Tweedie and renoise both with the guided eps; fused
        trajectory on the STEP_DDIM_CFG mode with the fp16 VAE latent as state."""
        if cfg_guidance == 0.0 or cfg_guidance == 1.0:
            add_cond_kwargs['text_embeds'] = add_cond_kwargs['text_embeds'][-1].unsqueeze(0)
            add_cond_kwargs['time_ids'] = add_cond_kwargs['time_ids'][-1].unsqueeze(0)
        steps = S.ddim_inversion_cfgpp_steps(self._sch, cfg_guidance)
        z0 = z0.clone().to(self.device)
        return self._run_trajectory(S.STEP_DDIM_CFG, z0.dtype, steps, z0, uc, c, add_cond_kwargs, None, 'zt')

    def sigma_to_t(self, sigma, quantize=None):
        quantize = self.quantize if quantize is None else quantize
        return S.sigma_to_t(self._sch, sigma, quantize)

    # ---- shared trajectory driver -------------------------------------------------------------------------------
    def _run_trajectory(self, method, state_dtype, steps, z_init, uc, c, add_cond_kwargs, callback_fn, result):
        """`result`: 'z0t' (DDIM family returns the Tweedie estimate of the last step) or 'zt' (DPM++ returns x)."""
        self._prepare(z_init, uc, c, add_cond_kwargs, force=True)  # every trajectory re-binds its prompt
        eng = self.unet
        eng.set_schedule(method, state_dtype, steps)
        eng.set_state(z_init)
        if callback_fn is None:
            eng.run_steps(0, len(steps))
        else:
            for i, st in enumerate(steps):
                zt = eng.get_state(0)
                eps_uc, eps_c = eng.predict_noise(zt, st.t, st.in_scale)
                eng.apply_step(i, eps_uc, eps_c)
                kw = {'z0t': eng.get_state(1).detach(), 'zt': eng.get_state(0).detach(), 'decode': self.decode}
                kw = callback_fn(i, torch.tensor(int(st.t), device=self.device), kw)
                eng.set_state(kw['zt'])
                self._cb_z0t = kw['z0t']
            if result == 'z0t':
                return self._cb_z0t
        return eng.get_state(1 if result == 'z0t' else 0)


class SDXLLightning(SDXL):
    schedule_kind = "lightning"

    def __init__(self,
                 solver_config,
                 base_model_key: str = "stabilityai/stable-diffusion-xl-base-1.0",
                 light_model_ckpt: str = "ckpt/sdxl_lightning_4step_unet.safetensors",
                 dtype=torch.float16,
                 device='cuda',
                 **kwargs):
        import os
        key = light_model_ckpt if os.path.exists(light_model_ckpt) else "synthetic:4321"
        if key.startswith("synthetic"):
            warnings.warn(f"Lightning checkpoint '{light_model_ckpt}' not found; using seeded synthetic UNet weights")
        SDXL.__init__(self, solver_config, model_key=key, dtype=dtype, device=device, **kwargs)


###########################################
# Base version (plain CFG — the baselines the paper compares against)
###########################################

@register_solver('ddim')
class BaseDDIM(SDXL):
    """latent_sdxl.py:425-467: fp32 state, fused trajectory, renoise with the guided eps."""
    step_mode = S.STEP_DDIM_CFG

    def reverse_process(self, null_prompt_embeds, prompt_embeds, cfg_guidance, add_cond_kwargs, shape=(1024, 1024),
                        callback_fn=None, **kwargs):
        b = null_prompt_embeds.shape[0]
        zt = kwargs.get('zT')
        if zt is None:
            zt = self.initialize_latent(size=(b, 4, shape[1] // self.vae_scale_factor, shape[0] // self.vae_scale_factor))
        steps = S.ddim_cfgpp_steps(self._sch, cfg_guidance, sdxl_indexing=True,
                                   tables_on_device=(self.schedule_kind == "lightning"))
        return self._run_trajectory(self.step_mode, torch.float32, steps, zt.float(), null_prompt_embeds,
                                    prompt_embeds, add_cond_kwargs, callback_fn, 'z0t')


@register_solver('euler')
class Euler(SDXL):
    """Karras Euler (VE casted), plain CFG, Karras sigmas (latent_sdxl.py:469-517)."""
    quantize = True

    @torch.no_grad()
    def reverse_process(self, null_prompt_embeds, prompt_embeds, cfg_guidance, add_cond_kwargs, shape=(1024, 1024),
                        callback_fn=None, **kwargs):
        ts = self.total_sigmas()
        sigmas = K.get_sigmas_karras(len(self.scheduler.timesteps), ts.min(), ts.max(), rho=7.)
        zt = kwargs.get('xT')
        if zt is None and kwargs.get('zT') is not None:  # an N(0,1) draw, as every other solver accepts it
            zt = kwargs['zT'].to(self.device) * (sigmas[0] ** 2 + 1) ** 0.5
        if zt is None:
            zt_dim = (1, 4, shape[1] // self.vae_scale_factor, shape[0] // self.vae_scale_factor)
            zt = self.initialize_latent(method="random_kdiffusion", latent_dim=zt_dim, sigmas=sigmas)
        z0t, _ = K.euler_cfgpp_loop(self, zt.to(torch.float16), sigmas, cfg_guidance,
                                    (null_prompt_embeds, prompt_embeds, add_cond_kwargs), callback_fn, cfgpp=False)
        return z0t


@register_solver('ddim_lightning')
class BaseDDIMLight(BaseDDIM, SDXLLightning):
    def __init__(self, **kwargs):
        SDXLLightning.__init__(self, **kwargs)

    def reverse_process(self, null_prompt_embeds, prompt_embeds, cfg_guidance, add_cond_kwargs, shape=(1024, 1024),
                        callback_fn=None, **kwargs):
        assert cfg_guidance == 1.0, "CFG should be turned off in the lightning version"
        return super().reverse_process(null_prompt_embeds, prompt_embeds, cfg_guidance, add_cond_kwargs, shape,
                                       callback_fn, **kwargs)


@register_solver('euler_lightning')
class EulerLight(Euler, SDXLLightning):
    quantize = True

    def __init__(self, **kwargs):
        SDXLLightning.__init__(self, **kwargs)

    def reverse_process(self, null_prompt_embeds, prompt_embeds, cfg_guidance, add_cond_kwargs, shape=(1024, 1024),
                        callback_fn=None, **kwargs):
        assert cfg_guidance == 1.0, "CFG should be turned off in the lightning version"
        return super().reverse_process(null_prompt_embeds, prompt_embeds, cfg_guidance, add_cond_kwargs, shape,
                                       callback_fn, **kwargs)


###########################################
# CFG++ version
###########################################

@register_solver("ddim_cfg++")
class BaseDDIMCFGpp(SDXL):
    def reverse_process(self,
                        null_prompt_embeds,
                        prompt_embeds,
                        cfg_guidance,
                        add_cond_kwargs,
                        shape=(1024, 1024),
                        callback_fn=None,
                        **kwargs):
        b = null_prompt_embeds.shape[0]
        zt = kwargs.get('zT')
        if zt is None:
            zt = self.initialize_latent(size=(b, 4, shape[1] // self.vae_scale_factor, shape[0] // self.vae_scale_factor))
        steps = S.ddim_cfgpp_steps(self._sch, cfg_guidance, sdxl_indexing=True,
                                   tables_on_device=(self.schedule_kind == "lightning"))
        # fp32 state: zt comes from torch.randn (fp32) and promotes every update (latent_sdxl.py:289, 741-744)
        return self._run_trajectory(S.STEP_DDIM_CFGPP, torch.float32, steps, zt.float(), null_prompt_embeds,
                                    prompt_embeds, add_cond_kwargs, callback_fn, 'z0t')


@register_solver('ddim_cfg++_lightning')
class BaseDDIMCFGppLight(BaseDDIMCFGpp, SDXLLightning):
    def __init__(self, **kwargs):
        SDXLLightning.__init__(self, **kwargs)

    def reverse_process(self,
                        null_prompt_embeds,
                        prompt_embeds,
                        cfg_guidance,
                        add_cond_kwargs,
                        shape=(1024, 1024),
                        callback_fn=None,
                        **kwargs):
        assert cfg_guidance == 1.0, "CFG should be turned off in the lightning version"
        return super().reverse_process(null_prompt_embeds, prompt_embeds, cfg_guidance, add_cond_kwargs, shape,
                                       callback_fn, **kwargs)


@register_solver('dpm++_2m_cfgpp')
class DPMpp2mCFGppSolver(SDXL):
    quantize = True

    def reverse_process(self,
                        null_prompt_embeds,
                        prompt_embeds,
                        cfg_guidance,
                        add_cond_kwargs,
                        shape=(1024, 1024),
                        callback_fn=None,
                        **kwargs):
        b = null_prompt_embeds.shape[0]
        steps, sigma0 = S.dpmpp_2m_cfgpp_steps(self._sch, cfg_guidance)
        x = kwargs.get('zT')
        if x is None:
            x = self.initialize_latent(method='random', size=(b, 4, shape[1] // self.vae_scale_factor,
                                                              shape[0] // self.vae_scale_factor))
        x = x.to(torch.float16)
        x = x * sigma0  # fp16 tensor x 0-dim fp32 -> fp16 (latent_sdxl.py:882-884)
        return self._run_trajectory(S.STEP_DPMPP2M_CFGPP, torch.float16, steps, x, null_prompt_embeds, prompt_embeds,
                                    add_cond_kwargs, callback_fn, 'zt')


@register_solver('dpm++_2m_cfgpp_lightning')
class DPMpp2mCFGppLightningSolver(DPMpp2mCFGppSolver, SDXLLightning):
    def __init__(self, **kwargs):
        SDXLLightning.__init__(self, **kwargs)

    def reverse_process(self, null_prompt_embeds, prompt_embeds, cfg_guidance, add_cond_kwargs, shape=(1024, 1024),
                        callback_fn=None, **kwargs):
        assert cfg_guidance == 1.0, "CFG should be turned off in the lightning version"
        return super().reverse_process(null_prompt_embeds, prompt_embeds, cfg_guidance, add_cond_kwargs, shape,
                                       callback_fn, **kwargs)


@register_solver('euler_cfg++')
class EulerCFGpp(SDXL):
    """Karras Euler (VE casted) with CFG++ on the sampling timesteps' own sigmas (latent_sdxl.py:757-808): the native
    UNet behind `predict_noise`, the Euler update in torch (kdiffusion.py)."""
    quantize = True

    @torch.no_grad()
    def reverse_process(self, null_prompt_embeds, prompt_embeds, cfg_guidance, add_cond_kwargs, shape=(1024, 1024),
                        callback_fn=None, **kwargs):
        total_sigmas = self.total_sigmas()
        sigmas = total_sigmas[torch.round(self.scheduler.timesteps.cpu()).int()]
        sigmas = torch.cat([sigmas, torch.tensor([0.0])])
        zt = kwargs.get('xT')
        if zt is None and kwargs.get('zT') is not None:  # an N(0,1) draw, as every other solver accepts it
            zt = kwargs['zT'].to(self.device) * (sigmas[0] ** 2 + 1) ** 0.5
        if zt is None:
            zt_dim = (1, 4, shape[1] // self.vae_scale_factor, shape[0] // self.vae_scale_factor)
            zt = self.initialize_latent(method="random_kdiffusion", latent_dim=zt_dim, sigmas=sigmas)
        z0t, _ = K.euler_cfgpp_loop(self, zt.to(torch.float16), sigmas, cfg_guidance,
                                    (null_prompt_embeds, prompt_embeds, add_cond_kwargs), callback_fn)
        return z0t


@register_solver('euler_cfg++_lightning')
class EulerCFGppLight(EulerCFGpp, SDXLLightning):
    quantize = True

    def __init__(self, **kwargs):
        SDXLLightning.__init__(self, **kwargs)

    def reverse_process(self, null_prompt_embeds, prompt_embeds, cfg_guidance, add_cond_kwargs, shape=(1024, 1024),
                        callback_fn=None, **kwargs):
        assert cfg_guidance == 1.0, "CFG should be turned off in the lightning version"
        return super().reverse_process(null_prompt_embeds, prompt_embeds, cfg_guidance, add_cond_kwargs, shape,
                                       callback_fn, **kwargs)


@register_solver("ddim_edit")
class EditWardSwapDDIM(SDXL):
    """Three-prompt front end of the editing solvers (prompt = [null, source, target]) — latent_sdxl.py:570-655 — and
    the plain-CFG edit loop (:656-707): plain inversion under the source prompt, plain DDIM under the target prompt,
    both through `alpha()` with the fp16 VAE latent as state (fused step mode STEP_DDIM_CFG)."""

    def reverse_process(self, null_prompt_embeds, src_prompt_embeds, tgt_prompt_embed, cfg_guidance,
                        add_src_cond_kwargs, add_tgt_cond_kwargs, callback_fn=None, **kwargs):
        zt = self.initialize_latent(method='ddim', src_img=kwargs.get('src_img', None), uc=null_prompt_embeds,
                                    c=src_prompt_embeds, cfg_guidance=cfg_guidance,
                                    add_cond_kwargs=add_src_cond_kwargs)
        steps = S.ddim_cfgpp_steps(self._sch, cfg_guidance, sdxl_indexing=False)
        return self._run_trajectory(S.STEP_DDIM_CFG, zt.dtype, steps, zt, null_prompt_embeds, tgt_prompt_embed,
                                    add_tgt_cond_kwargs, callback_fn, 'z0t')

    def sample(self,
               prompt1=["", "", ""],
               prompt2=["", "", ""],
               cfg_guidance: float = 5.0,
               original_size: Optional[Tuple[int, int]] = None,
               crops_coords_top_left: Tuple[int, int] = (0, 0),
               target_size: Optional[Tuple[int, int]] = None,
               negative_original_size: Optional[Tuple[int, int]] = None,
               negative_crops_coords_top_left: Tuple[int, int] = (0, 0),
               negative_target_size: Optional[Tuple[int, int]] = None,
               clip_skip: Optional[int] = None,
               **kwargs):
        height = self.default_sample_size * self.vae_scale_factor
        width = self.default_sample_size * self.vae_scale_factor
        original_size = original_size or (height, width)
        target_size = target_size or (height, width)

        (null_prompt_embeds, src_prompt_embeds, pool_null_embed, pool_src) = self.get_text_embed(
            prompt1[0], prompt1[1], prompt2[0], prompt2[1], clip_skip)
        (_, tgt_prompt_embeds, _, pool_tgt) = self.get_text_embed(prompt1[0], prompt1[2], prompt2[0], prompt2[2],
                                                                  clip_skip)
        proj_dim = int(pool_src.shape[-1])
        add_time_ids = self._get_add_time_ids(original_size, crops_coords_top_left, target_size,
                                              dtype=src_prompt_embeds.dtype, text_encoder_projection_dim=proj_dim)
        if negative_original_size is not None and negative_target_size is not None:
            negative_add_time_ids = self._get_add_time_ids(negative_original_size, negative_crops_coords_top_left,
                                                           negative_target_size, dtype=src_prompt_embeds.dtype,
                                                           text_encoder_projection_dim=proj_dim)
        else:
            negative_add_time_ids = add_time_ids
        add_src, add_tgt = pool_src, pool_tgt
        if cfg_guidance != 0.0 and cfg_guidance != 1.0:
            add_src = torch.cat([pool_null_embed, add_src], dim=0)
            add_tgt = torch.cat([pool_null_embed, add_tgt], dim=0)
            add_time_ids = torch.cat([negative_add_time_ids, add_time_ids], dim=0)
        add_src_cond_kwargs = {'text_embeds': add_src.to(self.device), 'time_ids': add_time_ids.to(self.device)}
        add_tgt_cond_kwargs = {'text_embeds': add_tgt.to(self.device), 'time_ids': add_time_ids.to(self.device)}

        zt = self.reverse_process(null_prompt_embeds, src_prompt_embeds, tgt_prompt_embeds, cfg_guidance,
                                  add_src_cond_kwargs, add_tgt_cond_kwargs, **kwargs)
        with torch.no_grad():
            img = self.decode(zt)
        img = (img / 2 + 0.5).clamp(0, 1)
        return img.detach().cpu()


@register_solver("ddim_edit_cfg++")
class EditWardSwapDDIMCFGpp(EditWardSwapDDIM):
    """CFG++ inversion under the source prompt (Tweedie with eps_uc, renoise with the guided eps), then CFG++ DDIM
    sampling under the target prompt — latent_sdxl.py:955-1025. Both loops index the schedule through `alpha()`
    (negative t -> final_alpha_cumprod) and carry the fp16 VAE latent, so they run as the two fused step modes the
    SD v1.5 `ddim_inversion_cfg++` uses."""

    @torch.no_grad()
    def inversion(self, z0, uc, c, cfg_guidance, add_cond_kwargs):
        if cfg_guidance == 0.0 or cfg_guidance == 1.0:
            add_cond_kwargs['text_embeds'] = add_cond_kwargs['text_embeds'][-1].unsqueeze(0)
            add_cond_kwargs['time_ids'] = add_cond_kwargs['time_ids'][-1].unsqueeze(0)
        steps = S.ddim_inversion_cfgpp_steps(self._sch, cfg_guidance)
        z0 = z0.clone().to(self.device)
        return self._run_trajectory(S.STEP_DDIM_INV_CFGPP, z0.dtype, steps, z0, uc, c, add_cond_kwargs, None, 'zt')

    def reverse_process(self, null_prompt_embeds, src_prompt_embeds, tgt_prompt_embed, cfg_guidance,
                        add_src_cond_kwargs, add_tgt_cond_kwargs, callback_fn=None, **kwargs):
        zt = self.initialize_latent(method='ddim', src_img=kwargs.get('src_img', None), uc=null_prompt_embeds,
                                    c=src_prompt_embeds, cfg_guidance=cfg_guidance,
                                    add_cond_kwargs=add_src_cond_kwargs)
        steps = S.ddim_cfgpp_steps(self._sch, cfg_guidance, sdxl_indexing=False)
        return self._run_trajectory(S.STEP_DDIM_CFGPP, zt.dtype, steps, zt, null_prompt_embeds, tgt_prompt_embed,
                                    add_tgt_cond_kwargs, callback_fn, 'z0t')


if __name__ == "__main__":
    print(f"Possble solvers: {[x for x in __SOLVER__.keys()]}")

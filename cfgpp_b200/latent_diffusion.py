"""Stable Diffusion v1.5 CFG++ solvers on the Blackwell-native backend.

Mirror of the reference's `latent_diffusion.py` solver API for the hot path: registry (:13-26), `StableDiffusion`
base (:54-241: `alpha`, `get_text_embed`, `encode`, `decode`, `predict_noise`, `inversion`, `initialize_latent`),
`ddim_cfg++` (:621-679) and `ddim_inversion_cfg++` (:882-957). Same names, argument meaning and errors; the UNet
forward, CFG++ mix and DDIM update run in hand-written sm_100a CUDA behind include/cfgpp_b200.h.
SURVEY §8 f1 (the rest of the CFG++ `--method` surface): `ddim_edit_cfg++` (:959-1010) on the same fused step modes;
`euler_cfg++` (:682-724), `euler_a_cfg++` (:727-768), `dpm++_2s_a_cfg++` (:771-827), `dpm++_2m_cfg++` (:830-879) as
fused VE-cast trajectories (kdiffusion.py: the ancestral ones with their noise drawn up front); the op-by-op torch
form over `predict_noise` runs when a callback is installed.

dtype note (reference promotion rules): `ddim_cfg++` keeps an fp32 latent state (zT is a fp32
`torch.randn`); `ddim_inversion_cfg++` starts from the fp16 VAE latent, so both its inversion loop and the following
sampling loop run with an fp16 state — every update op rounds to fp16, which the fused step kernel reproduces.
"""
from __future__ import annotations

from typing import Any, Optional

import torch

from . import kdiffusion as K
from . import schedule as S
from .conditioning import SyntheticTextEncoder
from .text_encoder import CLIPTextConfig, get_conditioner
from .config import UNetConfig, sd15_config
from .latent_sdxl import _Scheduler, get_engine

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


def default_text_encoder(cfg: UNetConfig, device):
    """CLIP-L (hidden 768) for SD v1.5; a proportionally narrow tower for the test-sized configs."""
    d = cfg.cross_attention_dim
    if d == 768:
        return get_conditioner("clip_l", device, "sd15")
    if d % 64:
        return SyntheticTextEncoder(d, 0)
    small = CLIPTextConfig(name=f"clip_{d}", vocab_size=1024, hidden_size=d, intermediate_size=4 * d, num_hidden_layers=2,
                           num_attention_heads=d // 64, pad_token_id=1023)
    return get_conditioner("", device, "sd15", cfg=small)


class StableDiffusion(K.KDiffusionMixin):
    def __init__(self,
                 solver_config,
                 model_key: str = "runwayml/stable-diffusion-v1-5",
                 device: Optional[torch.device] = None,
                 **kwargs):
        self.device = device
        self.dtype = kwargs.get("pipe_dtype", torch.float16)
        self.cfg: UNetConfig = kwargs.get("unet_config") or sd15_config()
        self.unet = get_engine(model_key, self.cfg, device, kwargs.get("state_dict"))
        # CLIP-L text tower on the native backend (text_encoder.py; the reference takes pipe.text_encoder,
        # latent_diffusion.py:65-66). Pass `text_encoder=fn`, prompt -> (hidden (1,77,D), None), to override.
        self.text_encoder = kwargs.get("text_encoder") or default_text_encoder(self.cfg, device)
        self.vae = kwargs.get("vae")
        if self.vae is None:
            # AutoencoderKL decoder on the native backend (vae.py; the reference uses pipe.vae, latent_diffusion.py:64)
            from .vae import get_vae
            self.vae = get_vae("sd15_vae", device)

        self._sch = S.Schedule.make(solver_config.num_sampling, "ddim")
        self.total_alphas = self._sch.total_alphas
        self.sigmas = self._sch.sigmas
        self.log_sigmas = self._sch.log_sigmas
        self.skip = self._sch.skip
        self.final_alpha_cumprod = self._sch.final_alpha_cumprod
        self.scheduler = _Scheduler(self._sch, device)

    def __call__(self, *args: Any, **kwargs: Any) -> Any:
        self.sample(*args, **kwargs)

    def sample(self, *args: Any, **kwargs: Any) -> Any:
        raise NotImplementedError("Solver must implement sample() method.")

    def alpha(self, t):
        return self._sch.alpha(t)

    @torch.no_grad()
    def get_text_embed(self, null_prompt, prompt):
        null_text_embed, _ = self.text_encoder(null_prompt, self.device)
        text_embed, _ = self.text_encoder(prompt, self.device)
        return null_text_embed, text_embed

    def encode(self, x):
        return self.vae.encode(x, self.dtype)

    def decode(self, zt):
        return self.vae.decode(zt).float()

    def _prepare(self, zt, uc, c, force: bool = False):
        b, _, h, w = zt.shape
        self.unet.prepare(b, h, w)
        self.unet.bind_prompt(uc, c, force=force)

    def predict_noise(self, zt: torch.Tensor, t: torch.Tensor, uc: torch.Tensor, c: torch.Tensor):
        if uc is None or c is None:
            uc = c if uc is None else uc
            c = uc if c is None else c
        self._prepare(zt, uc, c)
        return self.unet.predict_noise(zt, float(t))

    def _run(self, method, steps, z_init, uc, c, callback_fn=None):
        self._prepare(z_init, uc, c, force=True)  # every trajectory re-binds its prompt
        eng = self.unet
        eng.set_schedule(method, z_init.dtype, steps)
        eng.set_state(z_init)
        z0t = None
        if callback_fn is None:
            eng.run_steps(0, len(steps))
            z0t = eng.get_state(1)
        else:
            for i, st in enumerate(steps):
                eps_uc, eps_c = eng.predict_noise(eng.get_state(0), st.t)
                eng.apply_step(i, eps_uc, eps_c)
                kw = {'z0t': eng.get_state(1).detach(), 'zt': eng.get_state(0).detach(), 'decode': self.decode}
                kw = callback_fn(i, torch.tensor(int(st.t), device=self.device), kw)
                z0t = kw['z0t']
                eng.set_state(kw['zt'])
        return z0t, eng.get_state(0)

    @torch.no_grad()
    def inversion(self, z0: torch.Tensor, uc: torch.Tensor, c: torch.Tensor, cfg_guidance: float = 1.0):
        """Plain-CFG DDIM inversion (latent_diffusion.py:160-182): Tweedie and renoise both with the guided eps. Same
        scalars as the CFG++ inversion; the fused step kernel's STEP_DDIM_CFG mode picks the eps."""
        steps = S.ddim_inversion_cfgpp_steps(self._sch, cfg_guidance)
        _, zt = self._run(S.STEP_DDIM_CFG, steps, z0.clone().to(self.device), uc, c)
        return zt

    def initialize_latent(self, method: str = 'random', src_img: Optional[torch.Tensor] = None, **kwargs):
        if method == 'ddim':
            z = self.inversion(self.encode(src_img.to(self.dtype).to(self.device)), kwargs.get('uc'), kwargs.get('c'),
                               cfg_guidance=kwargs.get('cfg_guidance', 0.0))
        elif method == 'npi':
            z = self.inversion(self.encode(src_img.to(self.dtype).to(self.device)), kwargs.get('c'), kwargs.get('c'),
                               cfg_guidance=1.0)
        elif method == 'random':
            size = kwargs.get('latent_dim', (1, 4, self.cfg.sample_size, self.cfg.sample_size))
            z = torch.randn(size).to(self.device)  # CPU generator, then H2D — latent_diffusion.py:199-200
        elif method == 'random_kdiffusion':
            size = kwargs.get('latent_dim', (1, 4, self.cfg.sample_size, self.cfg.sample_size))
            sigmas = kwargs.get('sigmas', [14.6146])
            z = torch.randn(size).to(self.device)
            z = z * (sigmas[0] ** 2 + 1) ** 0.5
        else:
            raise NotImplementedError
        return z


###########################################
# Base version (plain CFG — the baselines the paper compares against)
###########################################

@register_solver("ddim")
class BaseDDIM(StableDiffusion):
    """Basic DDIM solver for SD with plain CFG (latent_diffusion.py:247-299), fused trajectory."""

    def reverse_process(self, uc, c, cfg_guidance, zt, callback_fn=None):
        steps = S.ddim_cfgpp_steps(self._sch, cfg_guidance, sdxl_indexing=False)
        z0t, _ = self._run(S.STEP_DDIM_CFG, steps, zt, uc, c, callback_fn)
        return z0t

    def sample(self, cfg_guidance=7.5, prompt=["", ""], callback_fn=None, **kwargs):
        uc, c = self.get_text_embed(null_prompt=prompt[0], prompt=prompt[1])
        zt = kwargs.get('zT')
        if zt is None:
            zt = self.initialize_latent()
        z0t = self.reverse_process(uc, c, cfg_guidance, zt, callback_fn)
        img = self.decode(z0t)
        img = (img / 2 + 0.5).clamp(0, 1)
        return img.detach().cpu()


@register_solver("ddim_inversion")
class InversionDDIM(BaseDDIM):
    """Reconstruction / editing after plain-CFG inversion (latent_diffusion.py:506-558)."""

    def sample(self, src_img, cfg_guidance=7.5, prompt=["", "", ""], callback_fn=None, **kwargs):
        uc, c = self.get_text_embed(null_prompt=prompt[0], prompt=prompt[1])
        zt = self.initialize_latent(method='ddim', src_img=src_img, uc=uc, c=c, cfg_guidance=cfg_guidance)
        z0t = self.reverse_process(uc, c, cfg_guidance, zt, callback_fn)
        img = self.decode(z0t)
        img = (img / 2 + 0.5).clamp(0, 1)
        return img.detach().cpu()


@register_solver("ddim_edit")
class EditWordSwapDDIM(InversionDDIM):
    """Editing via WordSwap after plain-CFG inversion (latent_diffusion.py:561-612)."""

    def sample(self, src_img, cfg_guidance=7.5, prompt=["", "", ""], callback_fn=None, **kwargs):
        uc, src_c = self.get_text_embed(null_prompt=prompt[0], prompt=prompt[1])
        _, tgt_c = self.get_text_embed(null_prompt=prompt[0], prompt=prompt[2])
        zt = self.initialize_latent(method='ddim', src_img=src_img, uc=uc, c=src_c, cfg_guidance=cfg_guidance)
        z0t = self.reverse_process(uc, tgt_c, cfg_guidance, zt, callback_fn)
        img = self.decode(z0t)
        img = (img / 2 + 0.5).clamp(0, 1)
        return img.detach().cpu()


###########################################
# CFG++ version
###########################################

@register_solver("ddim_cfg++")
class BaseDDIMCFGpp(StableDiffusion):
    """DDIM solver for SD with CFG++ (text-to-image)."""

    def reverse_process(self, uc, c, cfg_guidance, zt, callback_fn=None):
        steps = S.ddim_cfgpp_steps(self._sch, cfg_guidance, sdxl_indexing=False)
        z0t, _ = self._run(S.STEP_DDIM_CFGPP, steps, zt, uc, c, callback_fn)
        return z0t

    def sample(self, cfg_guidance=7.5, prompt=["", ""], callback_fn=None, **kwargs):
        uc, c = self.get_text_embed(null_prompt=prompt[0], prompt=prompt[1])
        zt = kwargs.get('zT')
        if zt is None:
            zt = self.initialize_latent()
        z0t = self.reverse_process(uc, c, cfg_guidance, zt, callback_fn)
        img = self.decode(z0t)
        img = (img / 2 + 0.5).clamp(0, 1)
        return img.detach().cpu()


@register_solver("ddim_inversion_cfg++")
class InversionDDIMCFGpp(BaseDDIMCFGpp):
    """Editing via WordSwap after inversion (CFG++ inversion: Tweedie with eps_uc, renoise with the guided eps)."""

    @torch.no_grad()
    def inversion(self, z0: torch.Tensor, uc: torch.Tensor, c: torch.Tensor, cfg_guidance: float = 1.0):
        steps = S.ddim_inversion_cfgpp_steps(self._sch, cfg_guidance)
        _, zt = self._run(S.STEP_DDIM_INV_CFGPP, steps, z0.clone().to(self.device), uc, c)
        return zt

    def sample(self, src_img, cfg_guidance=7.5, prompt=["", ""], callback_fn=None, **kwargs):
        uc, c = self.get_text_embed(null_prompt=prompt[0], prompt=prompt[1])
        zt = self.initialize_latent(method='ddim', src_img=src_img, uc=uc, c=c, cfg_guidance=cfg_guidance)
        z0t = self.reverse_process(uc, c, cfg_guidance, zt, callback_fn)
        img = self.decode(z0t)
        img = (img / 2 + 0.5).clamp(0, 1)
        return img.detach().cpu()


@register_solver("ddim_edit_cfg++")
class EditWordSwapDDIMCFGpp(InversionDDIMCFGpp):
    """Editing via WordSwap after inversion: CFG++ inversion under the source prompt, CFG++ sampling under the target
    prompt (latent_diffusion.py:959-1010). Both loops run as fused trajectories with an fp16 state."""

    def sample(self, src_img, cfg_guidance=7.5, prompt=["", "", ""], callback_fn=None, **kwargs):
        uc, src_c = self.get_text_embed(null_prompt=prompt[0], prompt=prompt[1])
        _, tgt_c = self.get_text_embed(null_prompt=prompt[0], prompt=prompt[2])
        zt = self.initialize_latent(method='ddim', src_img=src_img, uc=uc, c=src_c, cfg_guidance=cfg_guidance)
        z0t = self.reverse_process(uc, tgt_c, cfg_guidance, zt, callback_fn)
        img = self.decode(z0t)
        img = (img / 2 + 0.5).clamp(0, 1)
        return img.detach().cpu()


class _KarrasCFGpp(StableDiffusion):
    """Shared front / back end of the VE-cast samplers: Karras sigmas over the NFE steps, x ~ N(0, sigma_0^2 + 1) in
    fp16, decode of either the last Tweedie estimate or the final state (latent_diffusion.py:688-697, 720-724)."""
    adopt_callback = True
    decode_state = False  # True: decode x (dpm++ variants), False: decode the last denoised (euler variants)

    def _loop(self, x, sigmas, cfg_guidance, cond, callback_fn):
        raise NotImplementedError

    def karras_sigmas(self):
        ts = self.total_sigmas()
        return K.get_sigmas_karras(len(self.scheduler.timesteps), ts.min(), ts.max(), rho=7.)

    @torch.no_grad()
    def reverse_process(self, uc, c, cfg_guidance, x=None, callback_fn=None, noise=None):
        """`x`: a ready (scaled) start state; `noise`: an N(0,1) draw to scale (the `zT` every solver accepts)."""
        sigmas = self.karras_sigmas()
        if x is None and noise is not None:
            x = noise.to(self.device) * (sigmas[0] ** 2 + 1) ** 0.5
        if x is None:
            x = self.initialize_latent(method="random_kdiffusion", sigmas=sigmas,
                                       latent_dim=(1, 4, self.cfg.sample_size, self.cfg.sample_size))
        return self._loop(x.to(torch.float16), sigmas, cfg_guidance, (uc, c), callback_fn)

    def sample(self, cfg_guidance, prompt=["", ""], callback_fn=None, **kwargs):
        uc, c = self.get_text_embed(null_prompt=prompt[0], prompt=prompt[1])
        denoised, x = self.reverse_process(uc, c, cfg_guidance, kwargs.get('xT'), callback_fn, kwargs.get('zT'))
        img = self.decode(x if self.decode_state else denoised)
        img = (img / 2 + 0.5).clamp(0, 1)
        return img.detach().cpu()


@register_solver("euler_cfg++")
class EulerCFGppSolver(_KarrasCFGpp):
    def _loop(self, x, sigmas, cfg_guidance, cond, callback_fn):
        # the reference binds the callback's return values to unused names here (:713-719): nothing is adopted
        return K.euler_cfgpp_loop(self, x, sigmas, cfg_guidance, cond, callback_fn, adopt_callback=False)


@register_solver("euler_a_cfg++")
class EulerAncestralCFGppSolver(_KarrasCFGpp):
    """Karras Euler (VE casted) + ancestral sampling."""
    def _loop(self, x, sigmas, cfg_guidance, cond, callback_fn):
        return K.euler_cfgpp_loop(self, x, sigmas, cfg_guidance, cond, callback_fn, ancestral=True,
                                  adopt_callback=False)


@register_solver("dpm++_2s_a_cfg++")
class DPMpp2sAncestralCFGppSolver(_KarrasCFGpp):
    decode_state = True

    def _loop(self, x, sigmas, cfg_guidance, cond, callback_fn):
        return K.dpmpp_2s_a_cfgpp_loop(self, x, sigmas, cfg_guidance, cond, callback_fn)


@register_solver("dpm++_2m_cfg++")
class DPMpp2mCFGppSolver(_KarrasCFGpp):
    decode_state = True

    def _loop(self, x, sigmas, cfg_guidance, cond, callback_fn):
        return K.dpmpp_2m_cfgpp_karras_loop(self, x, sigmas, cfg_guidance, cond, callback_fn)


@register_solver("euler")
class EulerCFGSolver(_KarrasCFGpp):
    """Karras Euler (VE casted), plain CFG (latent_diffusion.py:302-346)."""
    def _loop(self, x, sigmas, cfg_guidance, cond, callback_fn):
        return K.euler_cfgpp_loop(self, x, sigmas, cfg_guidance, cond, callback_fn, adopt_callback=False, cfgpp=False)


@register_solver("euler_a")
class EulerAncestralCFGSolver(_KarrasCFGpp):
    """Karras Euler + ancestral sampling, plain CFG (latent_diffusion.py:349-390)."""
    def _loop(self, x, sigmas, cfg_guidance, cond, callback_fn):
        return K.euler_cfgpp_loop(self, x, sigmas, cfg_guidance, cond, callback_fn, ancestral=True,
                                  adopt_callback=False, cfgpp=False)


@register_solver("dpm++_2s_a")
class DPMpp2sAncestralCFGSolver(_KarrasCFGpp):
    """DPM-Solver++(2S) ancestral, plain CFG (latent_diffusion.py:393-451)."""
    decode_state = True

    def _loop(self, x, sigmas, cfg_guidance, cond, callback_fn):
        return K.dpmpp_2s_a_cfgpp_loop(self, x, sigmas, cfg_guidance, cond, callback_fn, cfgpp=False)


@register_solver("dpm++_2m")
class DPMpp2mCFGSolver(_KarrasCFGpp):
    """DPM-Solver++(2M), plain CFG (latent_diffusion.py:454-503)."""
    decode_state = True

    def _loop(self, x, sigmas, cfg_guidance, cond, callback_fn):
        return K.dpmpp_2m_cfgpp_karras_loop(self, x, sigmas, cfg_guidance, cond, callback_fn, cfgpp=False)


if __name__ == "__main__":
    print(f"Possble solvers: {[x for x in __SOLVER__.keys()]}")

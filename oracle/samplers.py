"""ORACLE (test infrastructure — never imported by the product path).

Restatement of the reference's CFG++ sampler loops, semantics (incl. dtype promotion, index shift and the
quirks of SURVEY.md Appendix C) kept line for line:

  predict_noise                 latent_diffusion.py:131-158, latent_sdxl.py:167-185
  SD v1.5  ddim_cfg++           latent_diffusion.py:621-679   (alpha(): :88-90)
  SD v1.5  ddim_inversion_cfg++ latent_diffusion.py:882-957
  SDXL     ddim_cfg++           latent_sdxl.py:713-755
  SDXL     ddim_cfg++_lightning latent_sdxl.py:838-858 (asserts cfg_guidance == 1.0)
  SDXL     dpm++_2m_cfgpp       latent_sdxl.py:860-930 (sigma_to_t :333-346, to_d :353-355)

PARITY UNPINNED: the reference has no tests / golden vectors and cannot be imported here (diffusers is
absent), so these loops are pinned only by the algebraic identities in tests/test_oracle_samplers.py
(lambda=0 => unconditional DDIM; eps_uc == eps_c => CFG++ == DDIM; inversion step inverts the sampling step;
DPM++2M first step == Euler-CFG++).

The loops take the conditioning tensors directly (text encoders / VAE are outside the hot path and have no
weights offline): `uc`, `c` (1,77,D) and, for SDXL, `add_cond_kwargs`; zT is passed explicitly (the
reference draws it from the CPU generator: latent_diffusion.py:200, latent_sdxl.py:289).
All functions run under the caller's autocast context exactly like the decorated reference methods.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .schedule import ScheduleTables


def predict_noise(unet, zt, t, uc, c, added_cond_kwargs=None):
    """latent_diffusion.py:131-158 / latent_sdxl.py:167-185 (only the batched uc+c branch is used by CFG++)."""
    t_in = t.unsqueeze(0)
    if uc is None:
        noise_c = unet(zt, t_in, encoder_hidden_states=c, added_cond_kwargs=added_cond_kwargs)["sample"]
        noise_uc = noise_c
    elif c is None:
        noise_uc = unet(zt, t_in, encoder_hidden_states=uc, added_cond_kwargs=added_cond_kwargs)["sample"]
        noise_c = noise_uc
    else:
        c_embed = torch.cat([uc, c], dim=0)
        z_in = torch.cat([zt] * 2)
        t_in = torch.cat([t_in] * 2)
        if zt.shape[0] > 1:  # batch extension (the reference is batch-1): B independent trajectories share t
            t_in = t.reshape(1).expand(z_in.shape[0])
        noise_pred = unet(z_in, t_in, encoder_hidden_states=c_embed, added_cond_kwargs=added_cond_kwargs)["sample"]
        noise_uc, noise_c = noise_pred.chunk(2)
    return noise_uc, noise_c


def _alpha_sd15(tb: ScheduleTables, t):
    """StableDiffusion.alpha (latent_diffusion.py:88-90): negative t -> final_alpha_cumprod."""
    return tb.alphas_cumprod[int(t)] if int(t) >= 0 else tb.final_alpha_cumprod


@torch.no_grad()
def sd15_ddim_cfgpp(unet, tb: ScheduleTables, zT, uc, c, cfg_guidance: float,
                    callback_fn: Optional[Callable] = None, record: Optional[list] = None):
    """latent_diffusion.py:634-679 up to (not including) the VAE decode. Returns z0t of the last step."""
    zt = zT
    z0t = None
    for step, t in enumerate(tb.timesteps):
        at = _alpha_sd15(tb, t)            # CPU 0-dim tensor, as in the reference (acts as an fp32 scalar)
        at_prev = _alpha_sd15(tb, t - tb.skip)
        noise_uc, noise_c = predict_noise(unet, zt, t.to(zt.device), uc, c)
        noise_pred = noise_uc + cfg_guidance * (noise_c - noise_uc)
        if record is not None:
            record.append({"zt": zt.clone(), "noise_uc": noise_uc.clone(), "noise_c": noise_c.clone()})
        z0t = (zt - (1 - at).sqrt() * noise_pred) / at.sqrt()
        zt = at_prev.sqrt() * z0t + (1 - at_prev).sqrt() * noise_uc
        if callback_fn is not None:
            kw = callback_fn(step, t, {"z0t": z0t.detach(), "zt": zt.detach(), "decode": None})
            z0t, zt = kw["z0t"], kw["zt"]
    return z0t


@torch.no_grad()
def sd15_inversion_cfgpp(unet, tb: ScheduleTables, z0, uc, c, cfg_guidance: float):
    """InversionDDIMCFGpp.inversion, latent_diffusion.py:888-910 (Tweedie with eps_uc, renoise with guided eps)."""
    zt = z0.clone()
    for t in reversed(tb.timesteps):
        at = _alpha_sd15(tb, t)
        at_prev = _alpha_sd15(tb, t - tb.skip)
        noise_uc, noise_c = predict_noise(unet, zt, t.to(zt.device), uc, c)
        noise_pred = noise_uc + cfg_guidance * (noise_c - noise_uc)
        z0t = (zt - (1 - at_prev).sqrt() * noise_uc) / at_prev.sqrt()
        zt = at.sqrt() * z0t + (1 - at).sqrt() * noise_pred
    return zt


@torch.no_grad()
def sd15_ddim_inversion_cfgpp(unet, tb, z0_src, uc, c, cfg_guidance, callback_fn=None):
    """InversionDDIMCFGpp.sample, latent_diffusion.py:912-957 with the VAE-encoded source latent given."""
    zT = sd15_inversion_cfgpp(unet, tb, z0_src, uc, c, cfg_guidance)
    return sd15_ddim_cfgpp(unet, tb, zT, uc, c, cfg_guidance, callback_fn)


@torch.no_grad()
def sdxl_ddim_cfgpp(unet, tb: ScheduleTables, zT, uc, c, cfg_guidance: float, add_cond_kwargs,
                    callback_fn: Optional[Callable] = None, record: Optional[list] = None,
                    tables_on_device: bool = False):
    """BaseDDIMCFGpp.reverse_process, latent_sdxl.py:715-755. `at_next` of the last step is a negative-index
    lookup (t - skip < 0) whose result only feeds the discarded last zt (Appendix C.2).
    The table lives on the CPU for SDXL (:67) — a CPU 0-dim operand enters CUDA tensor ops as an fp32 scalar — and
    on the device for Lightning (:418), where a 0-dim CUDA operand of an fp16 tensor op is first cast to fp16."""
    zt = zT
    z0t = None
    acp = tb.alphas_cumprod.to(zt.device) if tables_on_device else tb.alphas_cumprod
    for step, t in enumerate(tb.timesteps.int()):
        next_t = t - tb.skip
        at = acp[t]
        at_next = acp[next_t]
        noise_uc, noise_c = predict_noise(unet, zt, t.to(zt.device), uc, c, add_cond_kwargs)
        noise_pred = noise_uc + cfg_guidance * (noise_c - noise_uc)
        if record is not None:
            record.append({"zt": zt.clone(), "noise_uc": noise_uc.clone(), "noise_c": noise_c.clone()})
        z0t = (zt - (1 - at).sqrt() * noise_pred) / at.sqrt()
        zt = at_next.sqrt() * z0t + (1 - at_next).sqrt() * noise_uc
        if callback_fn is not None:
            kw = callback_fn(step, t, {"z0t": z0t.detach(), "zt": zt.detach(), "decode": None})
            z0t, zt = kw["z0t"], kw["zt"]
    return z0t


@torch.no_grad()
def sdxl_ddim_cfgpp_lightning(unet, tb, zT, uc, c, cfg_guidance, add_cond_kwargs, callback_fn=None, record=None):
    """BaseDDIMCFGppLight.reverse_process, latent_sdxl.py:843-858."""
    assert cfg_guidance == 1.0, "CFG should be turned off in the lightning version"
    return sdxl_ddim_cfgpp(unet, tb, zT, uc, c, cfg_guidance, add_cond_kwargs, callback_fn, record,
                           tables_on_device=True)


def sigma_to_t(tb: ScheduleTables, sigma: torch.Tensor) -> torch.Tensor:
    """SDXL.sigma_to_t with quantize=True, latent_sdxl.py:333-339: nearest index in the UN-shifted sigma table."""
    total_sigmas = (1 - tb.total_alphas).sqrt() / tb.total_alphas.sqrt()
    dists = sigma - total_sigmas[:, None]
    return dists.abs().argmin(dim=0).view(sigma.shape)


@torch.no_grad()
def sdxl_dpmpp_2m_cfgpp(unet, tb: ScheduleTables, noise, uc, c, cfg_guidance: float, add_cond_kwargs,
                        callback_fn=None, record: Optional[list] = None):
    """DPMpp2mCFGppSolver.reverse_process, latent_sdxl.py:864-930. `noise` is the N(0,1) draw of
    initialize_latent (cast to fp16 and scaled by sigmas[0] here as in :882-884). fp16 state throughout."""
    alphas = tb.alphas_cumprod[tb.timesteps.int().cpu()].cpu()
    sigmas = (1 - alphas).sqrt() / alphas.sqrt()
    x = noise.to(torch.float16)
    x = x * sigmas[0]
    t_fn = lambda sigma: sigma.log().neg()  # noqa: E731
    old_denoised = None
    for i, _ in enumerate(tb.timesteps[:-1].int()):
        at = alphas[i]
        sigma = sigmas[i]
        c_in = at.clone().sqrt()
        c_out = -sigma.clone()
        new_t = sigma_to_t(tb, sigma).to(x.device)
        noise_uc, noise_c = predict_noise(unet, x * c_in, new_t, uc, c, add_cond_kwargs)
        noise_pred = noise_uc + cfg_guidance * (noise_c - noise_uc)
        if record is not None:
            record.append({"x": x.clone(), "noise_uc": noise_uc.clone(), "noise_c": noise_c.clone(),
                           "old_denoised": None if old_denoised is None else old_denoised.clone()})
        denoised = x + c_out * noise_pred
        uncond_denoised = x + c_out * noise_uc
        t, t_next = t_fn(sigmas[i]), t_fn(sigmas[i + 1])
        h = t_next - t
        if old_denoised is None or sigmas[i + 1] == 0:
            x = denoised + (x - uncond_denoised) / sigmas[i].item() * sigmas[i + 1]
        else:
            h_last = t - t_fn(sigmas[i - 1])
            r = h_last / h
            extra1 = -torch.exp(-h) * uncond_denoised - (-h).expm1() * (uncond_denoised - old_denoised) / (2 * r)
            extra2 = torch.exp(-h) * x
            x = denoised + extra1 + extra2
        old_denoised = uncond_denoised
        if callback_fn is not None:
            kw = callback_fn(i, new_t, {"z0t": denoised.detach(), "zt": x.detach(), "decode": None})
            denoised, x = kw["z0t"], kw["zt"]
    return x

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
absent), so these loops are pinned only by the algebraic identities in tests/test_oracle.py and the golden vectors
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


# ------------------------------------------------------------------------------------------------------------------
# the VE-cast ("k-diffusion") CFG++ samplers and the CFG++ editing loops.
#   helpers                       latent_diffusion.py:30-37 (ancestral step), :211-241 (timestep / to_d / x->denoised)
#   SD v1.5 euler_cfg++           latent_diffusion.py:682-724      euler_a_cfg++   :727-768
#   SD v1.5 dpm++_2s_a_cfg++      latent_diffusion.py:771-827      dpm++_2m_cfg++  :830-879
#   SD v1.5 ddim_edit_cfg++       latent_diffusion.py:959-1010
#   SDXL    euler_cfg++           latent_sdxl.py:757-808           ddim_edit_cfg++ :954-1025
# Same conventions as above: conditioning tensors and the N(0,1) draw are passed in; CPU 0-dim sigmas.
# ------------------------------------------------------------------------------------------------------------------

def ancestral_step(sigma_from, sigma_to, eta=1.0):
    """latent_diffusion.py:30-37."""
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


def kd_timestep(tb: ScheduleTables, sigma: torch.Tensor) -> torch.Tensor:
    """StableDiffusion.timestep / SDXL.timestep (latent_diffusion.py:211-214): nearest training level in log-sigma."""
    dists = sigma.log() - tb.log_sigmas[:, None]
    return dists.abs().argmin(dim=0).view(sigma.shape)


def kd_denoised(unet, x, sigma, t, uc, c, cfg_guidance, add_cond_kwargs=None):
    """kdiffusion_x_to_denoised (latent_diffusion.py:232-241) / kdiffusion_zt_to_denoised (latent_sdxl.py:357-363)."""
    xc = x / (sigma ** 2 + 1) ** 0.5
    noise_uc, noise_c = predict_noise(unet, xc, t, uc, c, add_cond_kwargs)
    noise_pred = noise_uc + cfg_guidance * (noise_c - noise_uc)
    return x - noise_pred * sigma, x - noise_uc * sigma


def karras_sigmas(tb: ScheduleTables):
    from .schedule import get_sigmas_karras
    total_sigmas = (1 - tb.total_alphas).sqrt() / tb.total_alphas.sqrt()
    return get_sigmas_karras(len(tb.timesteps), total_sigmas.min(), total_sigmas.max(), rho=7.0)


@torch.no_grad()
def kd_euler_cfgpp(unet, tb, x, sigmas, uc, c, cfg_guidance, add_cond_kwargs=None, ancestral=False, plus=True):
    """The loop body shared by euler_cfg++ (latent_diffusion.py:701-711, latent_sdxl.py:787-799) and euler_a_cfg++
    (latent_diffusion.py:745-755). x is the scaled fp16 start state. Returns (last denoised, x).
    plus=False: the plain-CFG `euler` / `euler_a` (latent_diffusion.py:322-330, :366-379; latent_sdxl.py:497-507) —
    the ODE derivative takes the guided estimate."""
    denoised = None
    for i in range(len(sigmas) - 1):
        sigma = sigmas[i]
        t = kd_timestep(tb, sigma).to(x.device)
        denoised, uncond_denoised = kd_denoised(unet, x, sigma, t, uc, c, cfg_guidance, add_cond_kwargs)
        d = (x - (uncond_denoised if plus else denoised)) / sigma.item()
        if ancestral:
            sigma_down, sigma_up = ancestral_step(sigmas[i], sigmas[i + 1])
            x = denoised + d * sigma_down
            if sigmas[i + 1] > 0:
                x = x + torch.randn_like(x) * sigma_up
        else:
            x = denoised + d * sigmas[i + 1]
    return denoised, x


@torch.no_grad()
def kd_dpmpp_2s_a_cfgpp(unet, tb, x, sigmas, uc, c, cfg_guidance, plus=True):
    """latent_diffusion.py:786-817; plus=False: plain `dpm++_2s_a`, latent_diffusion.py:410-437."""
    t_fn = lambda sigma: sigma.log().neg()  # noqa: E731
    sigma_fn = lambda t: t.neg().exp()      # noqa: E731
    denoised = None
    for i in range(len(sigmas) - 1):
        sigma = sigmas[i]
        new_t = kd_timestep(tb, sigma).to(x.device)
        denoised, uncond_denoised = kd_denoised(unet, x, sigma, new_t, uc, c, cfg_guidance)
        sigma_down, sigma_up = ancestral_step(sigmas[i], sigmas[i + 1])
        first = uncond_denoised if plus else denoised
        if sigma_down == 0:
            d = (x - first) / sigmas[i].item()
            x = denoised + d * sigma_down
        else:
            t, t_next = t_fn(sigmas[i]), t_fn(sigma_down)
            r = 1 / 2
            h = t_next - t
            s = t + r * h
            x_2 = (sigma_fn(s) / sigma_fn(t)) * x - (-h * r).expm1() * first
            sigma_s = sigma_fn(s)
            t_2 = kd_timestep(tb, sigma_s).to(x.device)
            denoised_2, uncond_denoised_2 = kd_denoised(unet, x_2, sigma_s, t_2, uc, c, cfg_guidance)
            if plus:
                x = denoised_2 - torch.exp(-h) * uncond_denoised_2 + (sigma_fn(t_next) / sigma_fn(t)) * x
            else:
                x = (sigma_fn(t_next) / sigma_fn(t)) * x - (-h).expm1() * denoised_2
        if sigmas[i + 1] > 0:
            x = x + torch.randn_like(x) * sigma_up
    return denoised, x


@torch.no_grad()
def kd_dpmpp_2m_cfgpp_sd15(unet, tb, x, sigmas, uc, c, cfg_guidance, plus=True):
    """latent_diffusion.py:848-866 — second-order term on (denoised - old_denoised) (the SDXL file uses
    uncond_denoised there, see sdxl_dpmpp_2m_cfgpp above). plus=False: plain `dpm++_2m`, latent_diffusion.py:470-487."""
    t_fn = lambda sigma: sigma.log().neg()  # noqa: E731
    old_denoised = None
    denoised = None
    for i in range(len(sigmas) - 1):
        sigma = sigmas[i]
        new_t = kd_timestep(tb, sigma).to(x.device)
        denoised, uncond_denoised = kd_denoised(unet, x, sigma, new_t, uc, c, cfg_guidance)
        t, t_next = t_fn(sigmas[i]), t_fn(sigmas[i + 1])
        h = t_next - t
        lead = uncond_denoised if plus else denoised
        if old_denoised is None or sigmas[i + 1] == 0:
            x = denoised + (x - lead) / sigmas[i].item() * sigmas[i + 1]
        else:
            h_last = t - t_fn(sigmas[i - 1])
            r = h_last / h
            extra1 = -torch.exp(-h) * lead - (-h).expm1() * (denoised - old_denoised) / (2 * r)
            extra2 = torch.exp(-h) * x
            x = denoised + extra1 + extra2
        old_denoised = lead
    return denoised, x


def sdxl_euler_sigmas(tb: ScheduleTables):
    """latent_sdxl.py:773-778: the sampling timesteps' own sigmas (not Karras) plus a trailing 0."""
    total_sigmas = (1 - tb.total_alphas).sqrt() / tb.total_alphas.sqrt()
    sigmas = total_sigmas[torch.round(tb.timesteps.cpu()).int()]
    return torch.cat([sigmas, torch.tensor([0.0])])


def kd_start_state(noise, sigmas):
    """initialize_latent('random_kdiffusion') + the fp16 cast of the samplers (latent_diffusion.py:203-207, :695-697)."""
    return (noise * (sigmas[0] ** 2 + 1) ** 0.5).to(torch.float16)


@torch.no_grad()
def ddim_edit_cfgpp(unet, tb, z0_src, uc, c_src, c_tgt, cfg_guidance, add_src=None, add_tgt=None):
    """ddim_edit_cfg++: CFG++ inversion under the source prompt, CFG++ DDIM under the target prompt, both through
    alpha() (latent_diffusion.py:959-1010 with :888-910; latent_sdxl.py:955-1025). fp16 state throughout."""
    if add_src is not None and (cfg_guidance == 0.0 or cfg_guidance == 1.0):
        add_src = {k: v[-1].unsqueeze(0) for k, v in add_src.items()}
    zt = z0_src.clone()
    for t in reversed(tb.timesteps):
        at, at_prev = _alpha_sd15(tb, t), _alpha_sd15(tb, t - tb.skip)
        noise_uc, noise_c = predict_noise(unet, zt, t.to(zt.device), uc, c_src, add_src)
        noise_pred = noise_uc + cfg_guidance * (noise_c - noise_uc)
        z0t = (zt - (1 - at_prev).sqrt() * noise_uc) / at_prev.sqrt()
        zt = at.sqrt() * z0t + (1 - at).sqrt() * noise_pred
    zT = zt
    z0t = None
    for t in tb.timesteps:
        at, at_next = _alpha_sd15(tb, t), _alpha_sd15(tb, t - tb.skip)
        noise_uc, noise_c = predict_noise(unet, zt, t.to(zt.device), uc, c_tgt, add_tgt)
        noise_pred = noise_uc + cfg_guidance * (noise_c - noise_uc)
        z0t = (zt - (1 - at).sqrt() * noise_pred) / at.sqrt()
        zt = at_next.sqrt() * z0t + (1 - at_next).sqrt() * noise_uc
    return zT, z0t


# ------------------------------------------------------------------------------------------------------------------
# plain-CFG DDIM baselines (renoise with the guided eps) — latent_diffusion.py:160-182 (inversion),
# :247-299 (ddim), :506-612 (ddim_inversion / ddim_edit); latent_sdxl.py:301-324, :425-467, :656-707.
# ------------------------------------------------------------------------------------------------------------------

@torch.no_grad()
def ddim_plain(unet, tb, zT, uc, c, cfg_guidance, add_cond_kwargs=None, sdxl_indexing=False, record=None):
    """BaseDDIM: SD v1.5 indexes through alpha() (latent_diffusion.py:275-287), SDXL through the raw table with the
    negative-index wrap on the last step (latent_sdxl.py:443-455)."""
    zt, z0t = zT, None
    ts = tb.timesteps.int() if sdxl_indexing else tb.timesteps
    for t in ts:
        if sdxl_indexing:
            at, at_next = tb.alphas_cumprod[t], tb.alphas_cumprod[t - tb.skip]
        else:
            at, at_next = _alpha_sd15(tb, t), _alpha_sd15(tb, t - tb.skip)
        noise_uc, noise_c = predict_noise(unet, zt, t.to(zt.device), uc, c, add_cond_kwargs)
        noise_pred = noise_uc + cfg_guidance * (noise_c - noise_uc)
        if record is not None:
            record.append({"zt": zt.clone(), "noise_uc": noise_uc.clone(), "noise_c": noise_c.clone()})
        z0t = (zt - (1 - at).sqrt() * noise_pred) / at.sqrt()
        zt = at_next.sqrt() * z0t + (1 - at_next).sqrt() * noise_pred
    return z0t


@torch.no_grad()
def ddim_inversion_plain(unet, tb, z0, uc, c, cfg_guidance, add_cond_kwargs=None):
    """StableDiffusion.inversion / SDXL.inversion (latent_diffusion.py:160-182, latent_sdxl.py:301-324)."""
    if add_cond_kwargs is not None and (cfg_guidance == 0.0 or cfg_guidance == 1.0):
        add_cond_kwargs = {k: v[-1].unsqueeze(0) for k, v in add_cond_kwargs.items()}
    zt = z0.clone()
    for t in reversed(tb.timesteps):
        at, at_prev = _alpha_sd15(tb, t), _alpha_sd15(tb, t - tb.skip)
        noise_uc, noise_c = predict_noise(unet, zt, t.to(zt.device), uc, c, add_cond_kwargs)
        noise_pred = noise_uc + cfg_guidance * (noise_c - noise_uc)
        z0t = (zt - (1 - at_prev).sqrt() * noise_pred) / at_prev.sqrt()
        zt = at.sqrt() * z0t + (1 - at).sqrt() * noise_pred
    return zt


@torch.no_grad()
def ddim_edit_plain(unet, tb, z0_src, uc, c_src, c_tgt, cfg_guidance, add_src=None, add_tgt=None):
    """ddim_edit (c_tgt != c_src) / ddim_inversion (c_tgt == c_src): plain inversion, then plain DDIM through alpha()."""
    zT = ddim_inversion_plain(unet, tb, z0_src, uc, c_src, cfg_guidance, add_src)
    return zT, ddim_plain(unet, tb, zT, uc, c_tgt, cfg_guidance, add_tgt, sdxl_indexing=False)

"""ORACLE (test infrastructure — never imported by the product path).

Restatement of the scheduler *tables* the reference reads from diffusers 0.27.1 (`scheduler.step()` is never
called by the reference; only `alphas_cumprod`, `final_alpha_cumprod`, `timesteps`, `set_timesteps`):
  latent_diffusion.py:69-80, latent_sdxl.py:56-67 (DDIMScheduler, hub config: scaled_linear betas
  0.00085..0.012, 1000 train steps, steps_offset=1, timestep_spacing="leading", set_alpha_to_one=False)
  latent_sdxl.py:407-418 (EulerDiscreteScheduler, timestep_spacing="trailing", Lightning).

PARITY UNPINNED against diffusers itself (not installable offline); pinned by the known-answer constants
of SURVEY.md Appendix B in tests/test_schedule.py.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

NUM_TRAIN_TIMESTEPS = 1000


def alphas_cumprod_table() -> torch.Tensor:
    """DDIMScheduler.__init__: betas = linspace(sqrt(b0), sqrt(b1), T, fp32) ** 2; cumprod(1 - betas)."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, NUM_TRAIN_TIMESTEPS, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_leading_timesteps(num_inference_steps: int, steps_offset: int = 1) -> torch.Tensor:
    """DDIMScheduler.set_timesteps, timestep_spacing='leading'."""
    step_ratio = NUM_TRAIN_TIMESTEPS // num_inference_steps
    ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
    return torch.from_numpy(ts + steps_offset)


def euler_trailing_timesteps(num_inference_steps: int) -> torch.Tensor:
    """EulerDiscreteScheduler.set_timesteps, timestep_spacing='trailing' (float32 timesteps)."""
    ts = np.round(np.arange(NUM_TRAIN_TIMESTEPS, 0, -NUM_TRAIN_TIMESTEPS / num_inference_steps)) - 1
    return torch.from_numpy(ts.astype(np.float32))


@dataclass
class ScheduleTables:
    """What StableDiffusion.__init__/SDXL.__init__/SDXLLightning.__init__ leave on `self` / `self.scheduler`."""
    total_alphas: torch.Tensor        # (1000,) un-shifted abar
    sigmas: torch.Tensor              # sqrt(1-abar)/sqrt(abar)
    log_sigmas: torch.Tensor
    timesteps: torch.Tensor           # scheduler.timesteps after set_timesteps(NFE)
    skip: int                         # 1000 // NFE
    alphas_cumprod: torch.Tensor      # (1001,) = cat([1.0], abar)  -> index t == original t-1
    final_alpha_cumprod: torch.Tensor  # abar[0] (set_alpha_to_one=False); absent for Lightning


def make_tables(num_sampling: int, kind: str = "ddim") -> ScheduleTables:
    abar = alphas_cumprod_table()
    sigmas = (1 - abar).sqrt() / abar.sqrt()
    if kind == "ddim":
        ts = ddim_leading_timesteps(num_sampling)
    elif kind == "lightning":
        ts = euler_trailing_timesteps(num_sampling)
    else:
        raise ValueError(kind)
    return ScheduleTables(total_alphas=abar.clone(), sigmas=sigmas, log_sigmas=sigmas.log(), timesteps=ts,
                          skip=NUM_TRAIN_TIMESTEPS // num_sampling,
                          alphas_cumprod=torch.cat([torch.tensor([1.0]), abar]), final_alpha_cumprod=abar[0].clone())


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0):
    """latent_diffusion.py:43-50."""
    ramp = torch.linspace(0, 1, n + 1)[:-1]
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])])

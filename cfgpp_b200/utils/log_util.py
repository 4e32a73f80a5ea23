"""`set_seed` / `create_workdir` with the reference's semantics (utils/log_util.py:44-50): the CPU generator seed
defines zT (`torch.randn(size).to(device)`, latent_diffusion.py:200, latent_sdxl.py:289)."""
from pathlib import Path

import numpy as np
import torch


def create_workdir(workdir: Path):
    workdir.joinpath('result').mkdir(parents=True, exist_ok=True)


def set_seed(seed: int):
    torch.random.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    np.random.seed(seed)

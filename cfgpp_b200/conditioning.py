"""Shape-only stand-ins for components either side of the hot path (since round 2 only fall-backs: the native VAE
decoder of vae.py and the native CLIP text towers of text_encoder.py are the defaults of every solver).

The reference takes them from the diffusers pipeline: `pipe.tokenizer/text_encoder` (latent_diffusion.py:65-66,
latent_sdxl.py:46-49) and `pipe.vae` / `madebyollin/sdxl-vae-fp16-fix` (latent_diffusion.py:64, latent_sdxl.py:44).
Neither diffusers nor any checkpoint is available offline, so the solver classes accept pluggable objects with the
same call contract; the defaults below are deterministic stand-ins that produce tensors of the real shapes/dtypes
(they are NOT models). A user with diffusers installed passes the real encoders / VAE instead.
"""
from __future__ import annotations

import hashlib
from typing import Optional, Tuple

import torch


def _seed_from(text: str, salt: str) -> int:
    return int.from_bytes(hashlib.sha256((salt + "\x00" + text).encode()).digest()[:8], "little") & ((1 << 62) - 1)


class SyntheticTextEncoder:
    """Maps a prompt string to a fixed pseudo-random embedding: (1, 77, dim) hidden states and (1, pooled_dim) pooled
    output, fp16 — the shapes `get_text_embed` returns (latent_diffusion.py:93-115, latent_sdxl.py:96-128)."""

    def __init__(self, dim: int, pooled_dim: int = 0, n_ctx: int = 77, dtype=torch.float16):
        self.dim, self.pooled_dim, self.n_ctx, self.dtype = dim, pooled_dim, n_ctx, dtype

    def __call__(self, prompt: str, device) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        g = torch.Generator(device="cpu").manual_seed(_seed_from(prompt, f"ctx{self.dim}"))
        hidden = torch.randn(1, self.n_ctx, self.dim, generator=g).to(self.dtype).to(device)
        pooled = None
        if self.pooled_dim:
            pooled = torch.randn(1, self.pooled_dim, generator=g).to(self.dtype).to(device)
        return hidden, pooled


class LatentPreviewDecoder:
    """Stand-in for `vae.decode(z / scaling_factor).sample` (latent_diffusion.py:123-129, latent_sdxl.py:155-164):
    a fixed 4->3 linear map of the latent followed by nearest x8 upsampling, giving an image-shaped tensor in roughly
    [-1, 1]. VAE decode is outside the hot path and stays on the reference path when a real VAE is supplied."""

    # a commonly used latent->RGB preview projection for SD-family latents
    _W = torch.tensor([[0.298, 0.207, 0.208], [0.187, 0.286, 0.173], [-0.158, 0.189, 0.264], [-0.184, -0.271, -0.473]])

    def __init__(self, scale_factor: int = 8):
        self.scale_factor = scale_factor

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        w = self._W.to(z.device, torch.float32)
        img = torch.einsum("bchw,cr->brhw", z.float(), w)
        return torch.nn.functional.interpolate(img, scale_factor=self.scale_factor, mode="nearest")

    def encode(self, x: torch.Tensor, dtype=torch.float16) -> torch.Tensor:
        """Stand-in for `vae.encode(x).latent_dist.sample() * scaling` (latent_diffusion.py:117-121): 8x8 average
        pooling of a fixed 3->4 projection (deterministic; the real posterior sample would consume CUDA RNG)."""
        w = torch.linalg.pinv(self._W).to(x.device, torch.float32)  # (3,4)
        lat = torch.einsum("brhw,rc->bchw", x.float(), w)
        return torch.nn.functional.avg_pool2d(lat, self.scale_factor).to(dtype)

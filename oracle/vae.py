"""ORACLE (test infrastructure, not product code): pure-PyTorch restatement of diffusers 0.27.1 `AutoencoderKL` — the
DECODER behind `self.vae.decode(zt / scaling_factor).sample` (reference latent_sdxl.py:155-164, VAE =
madebyollin/sdxl-vae-fp16-fix :44; latent_diffusion.py:123-129, VAE of runwayml/stable-diffusion-v1-5 :64) and the
ENCODER + posterior behind `self.vae.encode(x).latent_dist.sample() * scaling_factor` (latent_sdxl.py:151-152,
latent_diffusion.py:117-121; keys `encoder.*`, `quant_conv.*`).

PARITY UNPINNED against the real library: diffusers is not installable offline and the reference ships no golden
vectors. Module names equal the diffusers state-dict keys (`post_quant_conv`, `decoder.conv_in`,
`decoder.mid_block.resnets.N`, `decoder.mid_block.attentions.0.{group_norm,to_q,to_k,to_v,to_out.0}`,
`decoder.up_blocks.N.resnets.M`, `decoder.up_blocks.N.upsamplers.0.conv`, `decoder.conv_norm_out`,
`decoder.conv_out`), so a real `vae/diffusion_pytorch_model.safetensors` loads with strict=True on the decoder subset.
Structural checksum: the SDXL / SD v1.5 decoder (+ post_quant_conv) has 49,490,199 parameters
(AutoencoderKL total 83,653,863 = encoder 34,163,592 + quant_conv 72 + this).

Only tests/, __graft_entry__.smoke() and bench.py may import this package.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class VAEConfig:
    name: str = "sdxl_vae"
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.13025   # SDXL; SD v1.5: 0.18215


def sdxl_vae_config() -> VAEConfig:
    return VAEConfig()


def sd15_vae_config() -> VAEConfig:
    return VAEConfig(name="sd15_vae", scaling_factor=0.18215)


def tiny_vae_config() -> VAEConfig:
    return VAEConfig(name="tiny_vae", block_out_channels=(64, 64, 128, 128), layers_per_block=1)


class VaeResnetBlock2D(nn.Module):
    """ResnetBlock2D with temb_channels=None (vae.py Decoder: resnet_eps=1e-6, output_scale_factor=1)."""
    def __init__(self, cin: int, cout: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class VaeAttention(nn.Module):
    """UNetMidBlock2D attention: Attention(heads = C / attention_head_dim = 1, bias=True, norm_num_groups=32, eps 1e-6,
    residual_connection=True, rescale_output_factor=1, upcast_softmax) driven by AttnProcessor2_0."""
    def __init__(self, c: int, groups: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        residual = x
        t = x.view(b, c, h * w).transpose(1, 2)
        t = self.group_norm(t.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]   # one head of width C
        o = self.to_out[0](o.to(q.dtype))
        return o.transpose(-1, -2).reshape(b, c, h, w) + residual


class VaeMidBlock(nn.Module):
    def __init__(self, c: int, groups: int):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnetBlock2D(c, c, groups), VaeResnetBlock2D(c, c, groups)])
        self.attentions = nn.ModuleList([VaeAttention(c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class VaeUpsample(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class VaeUpBlock(nn.Module):
    def __init__(self, cin: int, cout: int, layers: int, groups: int, add_upsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.upsamplers = nn.ModuleList([VaeUpsample(cout)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.latent_channels, boc[-1], 3, padding=1)
        self.mid_block = VaeMidBlock(boc[-1], g)
        rev = list(reversed(boc))
        ups, cout = [], rev[0]
        for i, c in enumerate(rev):
            cin, cout = cout, c
            ups.append(VaeUpBlock(cin, cout, cfg.layers_per_block + 1, g, add_upsample=(i != len(rev) - 1)))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for blk in self.up_blocks:
            x = blk(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKLDecoder(nn.Module):
    """`AutoencoderKL.decode(z).sample`: post_quant_conv (1x1) then the decoder. The caller divides by
    scaling_factor first, as the reference does (latent_sdxl.py:163, latent_diffusion.py:127)."""
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.cfg = cfg
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)
        self.decoder = Decoder(cfg)

    def forward(self, z):
        return self.decoder(self.post_quant_conv(z))


class VaeDownsample(nn.Module):
    """Downsample2D(use_conv=True, padding=0): zero-pad right / bottom by one pixel, then a stride-2 3x3 convolution."""
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class VaeDownBlock(nn.Module):
    """DownEncoderBlock2D: `layers` resnets, then the downsampler (all but the last block)."""
    def __init__(self, cin: int, cout: int, layers: int, groups: int, add_downsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([VaeDownsample(cout)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class Encoder(nn.Module):
    """diffusers 0.27.1 `Encoder` (double_z=True): conv_in -> down_blocks -> mid_block -> GroupNorm + SiLU -> conv_out
    (2 * latent_channels moments)."""
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.out_channels, boc[0], 3, padding=1)
        downs, cout = [], boc[0]
        for i, c in enumerate(boc):
            cin, cout = cout, c
            downs.append(VaeDownBlock(cin, cout, cfg.layers_per_block, g, add_downsample=(i != len(boc) - 1)))
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = VaeMidBlock(boc[-1], g)
        self.conv_norm_out = nn.GroupNorm(g, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for blk in self.down_blocks:
            x = blk(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKLEncoder(nn.Module):
    """`AutoencoderKL.encode(x).latent_dist`: encoder, quant_conv (1x1), DiagonalGaussianDistribution. `forward`
    returns (mean, std) with the distribution's clamp of the log-variance to [-30, 20]."""
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.cfg = cfg
        self.encoder = Encoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)

    def forward(self, x):
        moments = self.quant_conv(self.encoder(x))
        mean, logvar = torch.chunk(moments, 2, dim=1)
        logvar = torch.clamp(logvar, -30.0, 20.0)
        return mean, torch.exp(0.5 * logvar)


def build_vae_encoder(cfg: VAEConfig, state_dict=None, dtype=torch.float32, device="cpu") -> AutoencoderKLEncoder:
    if state_dict is None:
        return AutoencoderKLEncoder(cfg).to(device=device, dtype=dtype).eval().requires_grad_(False)
    with torch.device("meta"):
        m = AutoencoderKLEncoder(cfg)
    m.load_state_dict({k: v.to(device=device, dtype=dtype) for k, v in state_dict.items()}, strict=True, assign=True)
    return m.eval().requires_grad_(False)


@torch.no_grad()
def encode(vae: AutoencoderKLEncoder, x: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """SDXL.encode / StableDiffusion.encode: `vae.encode(x).latent_dist.sample() * scaling_factor`
    (latent_sdxl.py:151-152, latent_diffusion.py:117-121); `noise` is the `randn_tensor(mean.shape, dtype=param dtype)`
    draw of DiagonalGaussianDistribution.sample, made by the caller so both sides consume the same values."""
    mean, std = vae(x)
    return (mean + std * noise.to(mean.dtype)) * vae.cfg.scaling_factor


def count_params(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())


def build_vae_decoder(cfg: VAEConfig, state_dict=None, dtype=torch.float32, device="cpu") -> AutoencoderKLDecoder:
    if state_dict is None:
        return AutoencoderKLDecoder(cfg).to(device=device, dtype=dtype).eval().requires_grad_(False)
    with torch.device("meta"):
        m = AutoencoderKLDecoder(cfg)
    m.load_state_dict({k: v.to(device=device, dtype=dtype) for k, v in state_dict.items()}, strict=True, assign=True)
    return m.eval().requires_grad_(False)


@torch.no_grad()
def decode(vae: AutoencoderKLDecoder, zt: torch.Tensor) -> torch.Tensor:
    """SDXL.decode / StableDiffusion.decode: `vae.decode(zt / scaling_factor).sample.float()`."""
    return vae(zt / vae.cfg.scaling_factor).float()

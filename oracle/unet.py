"""ORACLE (test infrastructure — never imported by the product path).

Pure-PyTorch restatement of `diffusers==0.27.1` `UNet2DConditionModel.forward` as the reference calls it
(reference call sites: latent_diffusion.py:146,149,155 and latent_sdxl.py:170-171,174-175,181-182; the
dependency is pinned at environment.yaml:87 and is NOT vendored in /root/reference, so the algorithm is
restated here from the published 0.27.1 sources: models/unets/unet_2d_condition.py, unet_2d_blocks.py,
resnet.py, transformer_2d.py, attention.py, attention_processor.py (AttnProcessor2_0), embeddings.py,
activations.py (GEGLU), upsampling.py / downsampling.py).

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path and diffusers /
model weights are not installable offline, so this restatement cannot be pinned against the real library.
Pins we do have: the parameter-count checksums 859,520,964 (SD v1.5) / 2,567,463,684 (SDXL) which only
the exact published layer structure reproduces, and strict `load_state_dict` under diffusers key names.

Module attribute names equal the diffusers ones so that `state_dict()` keys are the diffusers keys.
The same code is (a) the fp32 truth, (b) under `torch.autocast('cuda', fp16)` the stand-in for the
reference's CUDA path (it issues the same torch ops), (c) on CPU the "repo's own CPU eager path".
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass(frozen=True)
class UNetConfig:
    name: str
    sample_size: int
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",)
    up_block_types: Tuple[str, ...] = ("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 1, 1, 1)
    num_attention_heads: Tuple[int, ...] = (8, 8, 8, 8)  # diffusers' `attention_head_dim` misnomer
    cross_attention_dim: int = 768
    use_linear_projection: bool = False
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    addition_embed_type: Optional[str] = None  # "text_time" for SDXL
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816
    pooled_dim: int = 1280

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4


def sd15_config() -> UNetConfig:
    return UNetConfig(name="sd15", sample_size=64)


def sdxl_config() -> UNetConfig:
    return UNetConfig(
        name="sdxl", sample_size=128, block_out_channels=(320, 640, 1280),
        down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
        up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
        transformer_layers_per_block=(1, 2, 10), num_attention_heads=(5, 10, 20), cross_attention_dim=2048,
        use_linear_projection=True, addition_embed_type="text_time")


def tiny_sdxl_config(sample_size: int = 32) -> UNetConfig:
    """Same topology as SDXL (incl. text_time embedding, linear projections, head_dim 64), small widths."""
    return UNetConfig(
        name="tiny_sdxl", sample_size=sample_size, block_out_channels=(64, 128, 256),
        down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
        up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
        transformer_layers_per_block=(1, 1, 2), num_attention_heads=(1, 2, 4), cross_attention_dim=128,
        use_linear_projection=True, addition_embed_type="text_time", addition_time_embed_dim=32,
        projection_class_embeddings_input_dim=6 * 32 + 64, pooled_dim=64)


def tiny_sd15_config(sample_size: int = 32) -> UNetConfig:
    """Same topology as SD v1.5 (1x1-conv projections, 4 levels), small widths, head_dim 64."""
    return UNetConfig(name="tiny_sd15", sample_size=sample_size, block_out_channels=(64, 128, 256, 256),
                      num_attention_heads=(1, 2, 4, 4), cross_attention_dim=128)


# ------------------------------------------------------------------------------------------------
# embeddings.py
# ------------------------------------------------------------------------------------------------
def get_timestep_embedding(timesteps: torch.Tensor, embedding_dim: int, flip_sin_to_cos: bool = True,
                           downscale_freq_shift: float = 0.0, scale: float = 1.0,
                           max_period: int = 10000) -> torch.Tensor:
    half_dim = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half_dim, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half_dim - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half_dim:], emb[:, :half_dim]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(F.silu(self.linear_1(sample)))


# ------------------------------------------------------------------------------------------------
# resnet.py / downsampling.py / upsampling.py
# ------------------------------------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, stride=1, padding=1)
        self.conv_shortcut = (nn.Conv2d(in_channels, out_channels, 1, stride=1, padding=0)
                              if in_channels != out_channels else None)

    def forward(self, input_tensor, temb):
        hidden_states = self.conv1(F.silu(self.norm1(input_tensor)))
        temb = self.time_emb_proj(F.silu(temb))[:, :, None, None]
        hidden_states = hidden_states + temb
        hidden_states = self.conv2(F.silu(self.norm2(hidden_states)))  # dropout p=0
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + hidden_states) / 1.0  # output_scale_factor = 1.0


class Downsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


# ------------------------------------------------------------------------------------------------
# attention_processor.py (AttnProcessor2_0) / attention.py / transformer_2d.py
# ------------------------------------------------------------------------------------------------
class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, hidden_states, encoder_hidden_states=None):
        b = hidden_states.shape[0]
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q, k, v = self.to_q(hidden_states), self.to_k(ctx), self.to_v(ctx)
        hd = q.shape[-1] // self.heads
        q = q.view(b, -1, self.heads, hd).transpose(1, 2)
        k = k.view(b, -1, self.heads, hd).transpose(1, 2)
        v = v.view(b, -1, self.heads, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, self.heads * hd).to(q.dtype)
        return self.to_out[0](o) / 1.0  # rescale_output_factor = 1.0


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, hidden_states, encoder_hidden_states):
        hidden_states = self.attn1(self.norm1(hidden_states)) + hidden_states
        hidden_states = self.attn2(self.norm2(hidden_states), encoder_hidden_states) + hidden_states
        hidden_states = self.ff(self.norm3(hidden_states)) + hidden_states
        return hidden_states


class Transformer2DModel(nn.Module):
    def __init__(self, heads: int, dim_head: int, in_channels: int, num_layers: int, cross_attention_dim: int,
                 groups: int, use_linear_projection: bool):
        super().__init__()
        inner = heads * dim_head
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = nn.Linear(in_channels, inner)
            self.proj_out = nn.Linear(inner, in_channels)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner, 1)
            self.proj_out = nn.Conv2d(inner, in_channels, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])

    def forward(self, hidden_states, encoder_hidden_states):
        b, _, h, w = hidden_states.shape
        residual = hidden_states
        hidden_states = self.norm(hidden_states)
        if not self.use_linear_projection:
            hidden_states = self.proj_in(hidden_states)
            inner = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(b, h * w, inner)
        else:
            inner = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(b, h * w, inner)
            hidden_states = self.proj_in(hidden_states)
        for blk in self.transformer_blocks:
            hidden_states = blk(hidden_states, encoder_hidden_states)
        if not self.use_linear_projection:
            hidden_states = hidden_states.reshape(b, h, w, inner).permute(0, 3, 1, 2).contiguous()
            hidden_states = self.proj_out(hidden_states)
        else:
            hidden_states = self.proj_out(hidden_states)
            hidden_states = hidden_states.reshape(b, h, w, inner).permute(0, 3, 1, 2).contiguous()
        return hidden_states + residual


# ------------------------------------------------------------------------------------------------
# unet_2d_blocks.py
# ------------------------------------------------------------------------------------------------
class DownBlock(nn.Module):  # DownBlock2D / CrossAttnDownBlock2D
    def __init__(self, cfg: UNetConfig, i: int, in_ch: int, out_ch: int, is_final: bool):
        super().__init__()
        has_attn = cfg.down_block_types[i] == "CrossAttnDownBlock2D"
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_ch if j == 0 else out_ch, out_ch, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps)
            for j in range(cfg.layers_per_block)])
        if has_attn:
            heads = cfg.num_attention_heads[i]
            self.attentions = nn.ModuleList([
                Transformer2DModel(heads, out_ch // heads, out_ch, cfg.transformer_layers_per_block[i],
                                   cfg.cross_attention_dim, cfg.norm_num_groups, cfg.use_linear_projection)
                for _ in range(cfg.layers_per_block)])
        else:
            self.attentions = None
        self.downsamplers = None if is_final else nn.ModuleList([Downsample2D(out_ch)])

    def forward(self, hidden_states, temb, encoder_hidden_states):
        outputs = ()
        for j, resnet in enumerate(self.resnets):
            hidden_states = resnet(hidden_states, temb)
            if self.attentions is not None:
                hidden_states = self.attentions[j](hidden_states, encoder_hidden_states)
            outputs += (hidden_states,)
        if self.downsamplers is not None:
            hidden_states = self.downsamplers[0](hidden_states)
            outputs += (hidden_states,)
        return hidden_states, outputs


class MidBlock(nn.Module):  # UNetMidBlock2DCrossAttn
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        ch = cfg.block_out_channels[-1]
        heads = cfg.num_attention_heads[-1]
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps)
                                      for _ in range(2)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(heads, ch // heads, ch, cfg.transformer_layers_per_block[-1], cfg.cross_attention_dim,
                               cfg.norm_num_groups, cfg.use_linear_projection)])

    def forward(self, hidden_states, temb, encoder_hidden_states):
        hidden_states = self.resnets[0](hidden_states, temb)
        hidden_states = self.attentions[0](hidden_states, encoder_hidden_states)
        return self.resnets[1](hidden_states, temb)


class UpBlock(nn.Module):  # UpBlock2D / CrossAttnUpBlock2D
    def __init__(self, cfg: UNetConfig, i: int, in_ch: int, out_ch: int, prev_out_ch: int, is_final: bool):
        super().__init__()
        has_attn = cfg.up_block_types[i] == "CrossAttnUpBlock2D"
        n = cfg.layers_per_block + 1
        resnets = []
        for j in range(n):
            res_skip = in_ch if j == n - 1 else out_ch
            res_in = prev_out_ch if j == 0 else out_ch
            resnets.append(ResnetBlock2D(res_in + res_skip, out_ch, cfg.time_embed_dim, cfg.norm_num_groups,
                                         cfg.norm_eps))
        self.resnets = nn.ModuleList(resnets)
        if has_attn:
            rev = len(cfg.block_out_channels) - 1 - i
            heads = cfg.num_attention_heads[rev]
            self.attentions = nn.ModuleList([
                Transformer2DModel(heads, out_ch // heads, out_ch, cfg.transformer_layers_per_block[rev],
                                   cfg.cross_attention_dim, cfg.norm_num_groups, cfg.use_linear_projection)
                for _ in range(n)])
        else:
            self.attentions = None
        self.upsamplers = None if is_final else nn.ModuleList([Upsample2D(out_ch)])

    def forward(self, hidden_states, res_samples, temb, encoder_hidden_states):
        for j, resnet in enumerate(self.resnets):
            res = res_samples[-1]
            res_samples = res_samples[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = resnet(hidden_states, temb)
            if self.attentions is not None:
                hidden_states = self.attentions[j](hidden_states, encoder_hidden_states)
        if self.upsamplers is not None:
            hidden_states = self.upsamplers[0](hidden_states)
        return hidden_states, res_samples


# ------------------------------------------------------------------------------------------------
# unet_2d_condition.py
# ------------------------------------------------------------------------------------------------
class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.cfg = cfg
        boc = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], cfg.time_embed_dim)
        if cfg.addition_embed_type == "text_time":
            self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim, cfg.time_embed_dim)
        down = []
        out_ch = boc[0]
        for i in range(len(boc)):
            in_ch, out_ch = out_ch, boc[i]
            down.append(DownBlock(cfg, i, in_ch, out_ch, is_final=(i == len(boc) - 1)))
        self.down_blocks = nn.ModuleList(down)
        self.mid_block = MidBlock(cfg)
        rev = list(reversed(boc))
        up = []
        out_ch = rev[0]
        for i in range(len(boc)):
            prev_out = out_ch
            out_ch = rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            up.append(UpBlock(cfg, i, in_ch, out_ch, prev_out, is_final=(i == len(boc) - 1)))
        self.up_blocks = nn.ModuleList(up)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs=None):
        cfg = self.cfg
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.int64, device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        t_emb = get_timestep_embedding(timesteps, cfg.block_out_channels[0]).to(dtype=sample.dtype)
        emb = self.time_embedding(t_emb)
        if cfg.addition_embed_type == "text_time":
            text_embeds = added_cond_kwargs["text_embeds"]
            time_ids = added_cond_kwargs["time_ids"]
            time_embeds = get_timestep_embedding(time_ids.flatten(), cfg.addition_time_embed_dim)
            time_embeds = time_embeds.reshape((text_embeds.shape[0], -1))
            add_embeds = torch.concat([text_embeds, time_embeds], dim=-1).to(emb.dtype)
            emb = emb + self.add_embedding(add_embeds)

        sample = self.conv_in(sample)
        down_res = (sample,)
        for blk in self.down_blocks:
            sample, res = blk(sample, emb, encoder_hidden_states)
            down_res += res
        sample = self.mid_block(sample, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            res = down_res[-n:]
            down_res = down_res[:-n]
            sample, _ = blk(sample, res, emb, encoder_hidden_states)
        sample = self.conv_out(F.silu(self.conv_norm_out(sample)))
        return {"sample": sample}


def count_params(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())


def build_unet(cfg: UNetConfig, state_dict=None, dtype=torch.float32, device="cpu") -> UNet2DConditionModel:
    """Construct on the meta device when a state dict is given (no wasted init of 2.6 B params)."""
    if state_dict is None:
        return UNet2DConditionModel(cfg).to(device=device, dtype=dtype).eval().requires_grad_(False)
    with torch.device("meta"):
        m = UNet2DConditionModel(cfg)
    m.load_state_dict({k: v.to(device=device, dtype=dtype) for k, v in state_dict.items()}, strict=True, assign=True)
    return m.eval().requires_grad_(False)

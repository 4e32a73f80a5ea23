"""ORACLE (test infrastructure, not product code): pure-PyTorch restatement of the CLIP text towers behind the
reference's prompt conditioning — `pipe.text_encoder` (transformers `CLIPTextModel`: openai/clip-vit-large-patch14 for
SD v1.5, latent_diffusion.py:65-66, 93-115, and SDXL encoder 1, latent_sdxl.py:46-47) and `pipe.text_encoder_2`
(`CLIPTextModelWithProjection`, OpenCLIP ViT-bigG, latent_sdxl.py:48-49), consumed by `_text_embed`
(latent_sdxl.py:77-93: `hidden_states[-2]` / `[-(clip_skip + 2)]` and output `[0]`).

PARITY PINNED: unlike the UNet / VAE oracles this one is checked against the real library. `transformers` 5.5.0 is in
the image, so tests/test_text_encoder_cpu.py loads one random-init `CLIPTextModel` / `CLIPTextModelWithProjection`
state dict into both implementations and requires equal `hidden_states`, `last_hidden_state`, `pooler_output` and
`text_embeds` (fp32, <= 2e-5), and tests/golden/r02_clip_golden.pt holds outputs produced by transformers itself
(tests/golden/make_golden.py clip). Module names equal the transformers state-dict keys (`text_model.embeddings.*`,
`text_model.encoder.layers.N.{self_attn.{q,k,v,out}_proj,layer_norm1,layer_norm2,mlp.fc1,mlp.fc2}`,
`text_model.final_layer_norm`, `text_projection`), so a real checkpoint loads with strict=True.

Only tests/, __graft_entry__.smoke() and bench.py may import this package.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class CLIPTextCfg:
    name: str = "clip_l"
    vocab_size: int = 49408
    max_position_embeddings: int = 77
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    hidden_act: str = "quick_gelu"        # "quick_gelu" (OpenAI CLIP) | "gelu" (OpenCLIP bigG)
    layer_norm_eps: float = 1e-5
    projection_dim: int = 0               # > 0: CLIPTextModelWithProjection (bias-free text_projection)
    eos_token_id: int = 2                 # the SD / SDXL text-encoder configs keep the legacy value (argmax pooling)
    bos_token_id: int = 49406
    pad_token_id: int = 49407             # tokenizer-side: CLIP-L pads with <|endoftext|>, SDXL tokenizer_2 with "!" (0)


def clip_l_config() -> CLIPTextCfg:
    """openai/clip-vit-large-patch14 text tower (SD v1.5 text_encoder, SDXL text_encoder)."""
    return CLIPTextCfg()


def clip_bigg_config() -> CLIPTextCfg:
    """OpenCLIP ViT-bigG/14 text tower with projection (SDXL text_encoder_2)."""
    return CLIPTextCfg(name="clip_bigg", hidden_size=1280, intermediate_size=5120, num_hidden_layers=32,
                       num_attention_heads=20, hidden_act="gelu", projection_dim=1280, pad_token_id=0)


def tiny_clip_config(projection_dim: int = 0, act: str = "quick_gelu") -> CLIPTextCfg:
    return CLIPTextCfg(name="tiny_clip", vocab_size=256, hidden_size=128, intermediate_size=256, num_hidden_layers=3,
                       num_attention_heads=2, hidden_act=act, projection_dim=projection_dim, bos_token_id=254,
                       pad_token_id=255 if not projection_dim else 0)


def _act(name: str):
    if name == "quick_gelu":
        return lambda x: x * torch.sigmoid(1.702 * x)
    if name == "gelu":
        return F.gelu
    raise ValueError(f"unsupported hidden_act {name}")


class _Embeddings(nn.Module):
    def __init__(self, c: CLIPTextCfg):
        super().__init__()
        self.token_embedding = nn.Embedding(c.vocab_size, c.hidden_size)
        self.position_embedding = nn.Embedding(c.max_position_embeddings, c.hidden_size)

    def forward(self, ids):
        pos = torch.arange(ids.shape[-1], device=ids.device)
        return self.token_embedding(ids) + self.position_embedding(pos)[None]


class _Attention(nn.Module):
    def __init__(self, c: CLIPTextCfg):
        super().__init__()
        d = c.hidden_size
        self.heads = c.num_attention_heads
        self.k_proj = nn.Linear(d, d)
        self.v_proj = nn.Linear(d, d)
        self.q_proj = nn.Linear(d, d)
        self.out_proj = nn.Linear(d, d)

    def forward(self, x):
        b, t, d = x.shape
        hd = d // self.heads
        q = self.q_proj(x).view(b, t, self.heads, hd).transpose(1, 2)
        k = self.k_proj(x).view(b, t, self.heads, hd).transpose(1, 2)
        v = self.v_proj(x).view(b, t, self.heads, hd).transpose(1, 2)
        w = torch.matmul(q, k.transpose(-1, -2)) * hd ** -0.5
        mask = torch.full((t, t), float("-inf"), device=x.device, dtype=torch.float32).triu(1)   # causal, no padding mask
        w = F.softmax(w.float() + mask, dim=-1).to(q.dtype)
        o = torch.matmul(w, v).transpose(1, 2).reshape(b, t, d)
        return self.out_proj(o)


class _MLP(nn.Module):
    def __init__(self, c: CLIPTextCfg):
        super().__init__()
        self.fc1 = nn.Linear(c.hidden_size, c.intermediate_size)
        self.fc2 = nn.Linear(c.intermediate_size, c.hidden_size)
        self.act = _act(c.hidden_act)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _Layer(nn.Module):
    def __init__(self, c: CLIPTextCfg):
        super().__init__()
        self.self_attn = _Attention(c)
        self.layer_norm1 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.mlp = _MLP(c)
        self.layer_norm2 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)

    def forward(self, x):
        x = x + self.self_attn(self.layer_norm1(x))
        return x + self.mlp(self.layer_norm2(x))


class _Encoder(nn.Module):
    def __init__(self, c: CLIPTextCfg):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(c) for _ in range(c.num_hidden_layers)])


class _TextTransformer(nn.Module):
    def __init__(self, c: CLIPTextCfg):
        super().__init__()
        self.embeddings = _Embeddings(c)
        self.encoder = _Encoder(c)
        self.final_layer_norm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


def eos_positions(ids: torch.Tensor, eos_token_id: int) -> torch.Tensor:
    """Row of the pooled token (modeling_clip.CLIPTextTransformer.forward): legacy configs (eos_token_id == 2) take
    the arg-max token id (<|endoftext|> is the largest id), newer ones the first occurrence of eos_token_id."""
    if eos_token_id == 2:
        return ids.to(torch.int).argmax(dim=-1)
    return (ids.to(torch.int) == eos_token_id).int().argmax(dim=-1)


class CLIPText(nn.Module):
    """forward(ids) -> (hidden_states list of L+1 tensors, last_hidden_state, pooler_output, text_embeds | None)."""
    def __init__(self, c: CLIPTextCfg):
        super().__init__()
        self.cfg = c
        self.text_model = _TextTransformer(c)
        self.text_projection = nn.Linear(c.hidden_size, c.projection_dim, bias=False) if c.projection_dim else None

    @torch.no_grad()
    def forward(self, ids: torch.Tensor):
        tm = self.text_model
        x = tm.embeddings(ids)
        hs: List[torch.Tensor] = [x]
        for layer in tm.encoder.layers:
            x = layer(x)
            hs.append(x)
        last = tm.final_layer_norm(x)
        pos = eos_positions(ids, self.cfg.eos_token_id)
        pooled = last[torch.arange(last.shape[0], device=last.device), pos]
        embeds = self.text_projection(pooled) if self.text_projection is not None else None
        return hs, last, pooled, embeds


def build_clip_text(cfg: CLIPTextCfg, state_dict=None, dtype=torch.float32, device="cpu") -> CLIPText:
    if state_dict is None:
        return CLIPText(cfg).to(device=device, dtype=dtype).eval().requires_grad_(False)
    with torch.device("meta"):
        m = CLIPText(cfg)
    sd = {k: v.to(device=device, dtype=dtype) for k, v in state_dict.items() if not k.endswith("position_ids")}
    m.load_state_dict(sd, strict=True, assign=True)
    return m.eval().requires_grad_(False)


@torch.no_grad()
def text_embed(model: CLIPText, ids: torch.Tensor, clip_skip: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """SDXL._text_embed (latent_sdxl.py:77-93): (hidden_states[-2] or [-(clip_skip+2)], output[0]). Output [0] is
    `text_embeds` for the projection model and `last_hidden_state` for the plain one."""
    hs, last, _pooled, embeds = model(ids)
    h = hs[-2] if clip_skip is None else hs[-(clip_skip + 2)]
    return h, (embeds if embeds is not None else last)


@torch.no_grad()
def sd15_text_embed(model: CLIPText, ids: torch.Tensor) -> torch.Tensor:
    """StableDiffusion.get_text_embed (latent_diffusion.py:93-115): `text_encoder(ids)[0]` = last_hidden_state."""
    return model(ids)[1]


def count_params(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())

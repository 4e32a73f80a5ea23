"""Pins for the oracle itself (the reference has no tests / golden vectors -> PARITY UNPINNED vs diffusers):
structural checksums of the UNet restatement, and the algebraic identities of the CFG++ loops."""
import dataclasses

import pytest
import torch

from cfgpp_b200 import config as C
from cfgpp_b200 import weights as Wt
from oracle import samplers as OSm
from oracle import schedule as OS
from oracle import unet as O


def oracle_cfg(cfg):
    return O.UNetConfig(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(O.UNetConfig)})


@pytest.mark.parametrize("name,expect", [("sd15", 859_520_964), ("sdxl", 2_567_463_684)])
def test_param_count_checksum(name, expect):
    cfg = C.CONFIGS[name]()
    with torch.device("meta"):
        m = O.UNet2DConditionModel(oracle_cfg(cfg))
    assert O.count_params(m) == expect
    assert Wt.num_params(cfg) == expect
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: s for k, s, _ in Wt.unet_param_specs(cfg)}


@pytest.mark.parametrize("name", ["tiny_sdxl", "tiny_sd15"])
def test_tiny_forward_and_strict_load(name):
    cfg = C.CONFIGS[name]()
    sd = Wt.synthetic_state_dict(cfg, seed=3)
    m = O.build_unet(oracle_cfg(cfg), sd, dtype=torch.float32)
    g = torch.Generator().manual_seed(0)
    B, hw = 2, 16 if name == "tiny_sd15" else 16
    hw = 16
    x = torch.randn(B, 4, hw, hw, generator=g)
    ctx = torch.randn(B, 77, cfg.cross_attention_dim, generator=g)
    add = None
    if cfg.addition_embed_type:
        add = {"text_embeds": torch.randn(B, cfg.pooled_dim, generator=g),
               "time_ids": torch.tensor([[128., 128, 0, 0, 128, 128]] * B)}
    y = m(x, torch.tensor(500), ctx, add)["sample"]
    assert y.shape == x.shape and torch.isfinite(y).all() and 0.05 < y.std() < 20
    # batch rows are independent (the cond/uncond halves do not interact)
    y0 = m(x[:1], torch.tensor(500), ctx[:1], None if add is None else {k: v[:1] for k, v in add.items()})["sample"]
    assert torch.allclose(y0, y[:1], atol=2e-4, rtol=1e-4)


class FakeUNet:
    """eps(z, t, ctx) = a(t) * z + ctx-dependent offset; cheap, deterministic, batch-row independent."""
    def __init__(self, same=False):
        self.same = same

    def __call__(self, z, t, encoder_hidden_states=None, added_cond_kwargs=None):
        off = encoder_hidden_states.float().mean(dim=(1, 2)).view(-1, 1, 1, 1)
        if self.same:
            off = off * 0
        a = 0.3 + 0.0005 * t.float().view(-1, 1, 1, 1)
        return {"sample": (a * z.float() + off).to(z.dtype)}


def _ctx(seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, 77, 8, generator=g) + seed


def test_lambda0_is_unconditional_ddim_and_equal_eps_is_ddim():
    tb = OS.make_tables(10)
    g = torch.Generator().manual_seed(1)
    zT = torch.randn(1, 4, 8, 8, generator=g)
    uc, c = _ctx(1), _ctx(2)
    # plain DDIM with eps_uc only
    z = zT.clone()
    for t in tb.timesteps:
        at, an = tb.alphas_cumprod[int(t)], (tb.alphas_cumprod[int(t) - tb.skip] if int(t) - tb.skip >= 0 else tb.final_alpha_cumprod)
        e = FakeUNet()(z, t[None], uc)["sample"]
        z0 = (z - (1 - at).sqrt() * e) / at.sqrt()
        z = an.sqrt() * z0 + (1 - an).sqrt() * e
    got = OSm.sd15_ddim_cfgpp(FakeUNet(), tb, zT, uc, c, 0.0)
    assert torch.allclose(got, z0, atol=1e-5)
    # eps_uc == eps_c  =>  CFG++ == DDIM for any lambda
    a = OSm.sd15_ddim_cfgpp(FakeUNet(same=True), tb, zT, uc, c, 0.6)
    b = OSm.sd15_ddim_cfgpp(FakeUNet(same=True), tb, zT, uc, c, 0.0)
    assert torch.allclose(a, b, atol=1e-5)


def test_inversion_step_inverts_sampling_step():
    """One CFG++ sampling step followed by the CFG++ inversion step at the same (t, eps) is the identity."""
    tb = OS.make_tables(50)
    g = torch.Generator().manual_seed(2)
    zt = torch.randn(1, 4, 8, 8, generator=g)
    eu, ec = torch.randn(1, 4, 8, 8, generator=g), torch.randn(1, 4, 8, 8, generator=g)
    lam, t = 0.6, 501
    at, ap = tb.alphas_cumprod[t], tb.alphas_cumprod[t - tb.skip]
    npred = eu + lam * (ec - eu)
    z0 = (zt - (1 - at).sqrt() * npred) / at.sqrt()
    z_prev = ap.sqrt() * z0 + (1 - ap).sqrt() * eu          # sampling  (latent_diffusion.py:663-666)
    z0i = (z_prev - (1 - ap).sqrt() * eu) / ap.sqrt()        # inversion (latent_diffusion.py:907-908)
    z_back = at.sqrt() * z0i + (1 - at).sqrt() * npred
    assert torch.allclose(z_back, zt, atol=1e-5)


def test_sdxl_loop_matches_sd15_loop_except_last_step_discard():
    tb = OS.make_tables(10)
    g = torch.Generator().manual_seed(3)
    zT = torch.randn(1, 4, 8, 8, generator=g)
    uc, c = _ctx(3), _ctx(4)
    a = OSm.sd15_ddim_cfgpp(FakeUNet(), tb, zT, uc, c, 0.6)
    b = OSm.sdxl_ddim_cfgpp(FakeUNet(), tb, zT, uc, c, 0.6, None)
    assert torch.allclose(a, b, atol=1e-6)  # z0t of the last step never sees at_next


def test_dpmpp_first_step_is_euler_cfgpp_and_runs_nfe_minus_1():
    tb = OS.make_tables(25)
    g = torch.Generator().manual_seed(4)
    noise = torch.randn(1, 4, 8, 8, generator=g)
    uc, c = _ctx(5), _ctx(6)
    rec = []
    x = OSm.sdxl_dpmpp_2m_cfgpp(FakeUNet(), tb, noise, uc, c, 0.6, None, record=rec)
    assert len(rec) == 24 and x.dtype == torch.float16 and rec[0]["old_denoised"] is None
    assert torch.isfinite(x.float()).all()
    # UNet sees t-1 (sigma_to_t on the un-shifted table)
    alphas = tb.alphas_cumprod[tb.timesteps.int()]
    sig = (1 - alphas).sqrt() / alphas.sqrt()
    assert OSm.sigma_to_t(tb, sig[:3]).tolist() == [960, 920, 880]


def test_lightning_requires_lambda_1():
    tb = OS.make_tables(4, "lightning")
    with pytest.raises(AssertionError):
        OSm.sdxl_ddim_cfgpp_lightning(FakeUNet(), tb, torch.zeros(1, 4, 8, 8), _ctx(1), _ctx(2), 0.6, None)

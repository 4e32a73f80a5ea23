"""CPU tests of the VE-cast CFG++ samplers: the product loops (cfgpp_b200/kdiffusion.py) against
the oracle restatements (oracle/samplers.py) on a deterministic stand-in UNet, plus the algebraic identities that pin
the oracle (the reference has no tests of its own): eta = 0 ancestral == plain step, eps_uc == eps_c => CFG++ Euler ==
Euler, first DPM++(2M) step == Euler step, Karras schedule shape."""
import math

import pytest
import torch

from cfgpp_b200 import kdiffusion as K
from oracle import samplers as OSm, schedule as OS


class FakeUNet:
    """Deterministic, nonlinear stand-in with the diffusers call signature: eps depends on z, t and the context."""
    def __call__(self, z, t, encoder_hidden_states=None, added_cond_kwargs=None):
        t = t.reshape(-1, 1, 1, 1).to(z.dtype)
        ctx = encoder_hidden_states.float().mean(dim=(1, 2)).reshape(-1, 1, 1, 1).to(z.dtype)
        eps = torch.tanh(z * (0.5 + t / 1000)) * 0.8 + 0.3 * ctx
        if added_cond_kwargs is not None:
            eps = eps + 0.05 * added_cond_kwargs["text_embeds"].float().mean().to(z.dtype)
        return {"sample": eps}


class StubSolver(K.KDiffusionMixin):
    """What the product loops need from a solver object, with FakeUNet behind the predict_noise seam (the real classes
    put the native engine there)."""
    def __init__(self, tb, unet):
        self.log_sigmas, self.total_alphas, self.device = tb.log_sigmas, tb.total_alphas, torch.device("cpu")
        self.unet, self.decode = unet, None

    def predict_noise(self, zt, t, uc, c, added_cond_kwargs=None):
        tt = t.reshape(1)
        e_uc = self.unet(zt, tt, uc, added_cond_kwargs)["sample"]
        e_c = self.unet(zt, tt, c, added_cond_kwargs)["sample"]
        return e_uc, e_c


def _setup(nfe=6, hw=8, seed=3):
    g = torch.Generator().manual_seed(seed)
    tb = OS.make_tables(nfe)
    noise = torch.randn(1, 4, hw, hw, generator=g)
    uc = torch.randn(1, 77, 16, generator=g).half()
    c = torch.randn(1, 77, 16, generator=g).half()
    return tb, noise, uc, c, FakeUNet()


def _close16(a, b, ulps=2):
    a, b = a.float(), b.float()
    tol = ulps * torch.pow(2.0, torch.floor(torch.log2(b.abs().clamp_min(6.1e-5))) - 10)
    return bool(((a - b).abs() <= tol).all())


@pytest.mark.parametrize("plus", [True, False])
@pytest.mark.parametrize("ancestral", [False, True])
def test_product_euler_loops_match_oracle(ancestral, plus):
    tb, noise, uc, c, unet = _setup()
    sigmas = OSm.karras_sigmas(tb)
    assert torch.equal(sigmas, K.get_sigmas_karras(len(tb.timesteps), tb.sigmas.min(), tb.sigmas.max(), rho=7.))
    x0 = OSm.kd_start_state(noise, sigmas)
    torch.manual_seed(11)
    d_ref, x_ref = OSm.kd_euler_cfgpp(unet, tb, x0.clone(), sigmas, uc, c, 0.6, ancestral=ancestral, plus=plus)
    torch.manual_seed(11)
    d, x = K.euler_cfgpp_loop(StubSolver(tb, unet), x0.clone(), sigmas, 0.6, (uc, c), ancestral=ancestral, cfgpp=plus)
    assert x.dtype == torch.float16 and _close16(x, x_ref) and _close16(d, d_ref)


@pytest.mark.parametrize("plus", [True, False])
def test_product_dpmpp_loops_match_oracle(plus):
    tb, noise, uc, c, unet = _setup(nfe=7)
    sigmas = OSm.karras_sigmas(tb)
    x0 = OSm.kd_start_state(noise, sigmas)
    s = StubSolver(tb, unet)
    torch.manual_seed(5)
    d_ref, x_ref = OSm.kd_dpmpp_2s_a_cfgpp(unet, tb, x0.clone(), sigmas, uc, c, 0.6, plus=plus)
    torch.manual_seed(5)
    d, x = K.dpmpp_2s_a_cfgpp_loop(s, x0.clone(), sigmas, 0.6, (uc, c), cfgpp=plus)
    assert _close16(x, x_ref) and _close16(d, d_ref)
    d_ref, x_ref = OSm.kd_dpmpp_2m_cfgpp_sd15(unet, tb, x0.clone(), sigmas, uc, c, 0.6, plus=plus)
    d, x = K.dpmpp_2m_cfgpp_karras_loop(s, x0.clone(), sigmas, 0.6, (uc, c), cfgpp=plus)
    assert _close16(x, x_ref) and _close16(d, d_ref)


def test_sdxl_euler_uses_the_sampling_timesteps_sigmas():
    tb, noise, uc, c, unet = _setup(nfe=5)
    sigmas = OSm.sdxl_euler_sigmas(tb)
    assert len(sigmas) == 6 and sigmas[-1] == 0 and bool((sigmas[:-1] == tb.sigmas[tb.timesteps.long()]).all())
    add = {"text_embeds": torch.ones(2, 4).half(), "time_ids": torch.zeros(2, 6).half()}
    x0 = OSm.kd_start_state(noise, sigmas)
    d_ref, _ = OSm.kd_euler_cfgpp(unet, tb, x0.clone(), sigmas, uc, c, 0.6, add)
    d, _ = K.euler_cfgpp_loop(StubSolver(tb, unet), x0.clone(), sigmas, 0.6, (uc, c, add))
    assert _close16(d, d_ref)
    # the nearest-level lookup returns the sampling timestep itself
    for i, t in enumerate(tb.timesteps.long()):
        assert int(OSm.kd_timestep(tb, sigmas[i])) == int(t)


def test_ancestral_step_identities():
    for (a, b) in [(14.6, 9.1), (2.0, 0.5), (0.3, 0.0)]:
        a, b = torch.tensor(a), torch.tensor(b)
        down, up = OSm.ancestral_step(a, b)
        assert math.isclose(float(down ** 2 + up ** 2), float(b ** 2), rel_tol=1e-5, abs_tol=1e-7)
        assert OSm.ancestral_step(a, b, eta=0.0) == (b, 0.0)
        d2, u2 = K.get_ancestral_step(a, b)
        assert float(d2) == float(down) and float(u2) == float(up)


def test_cfgpp_euler_reduces_to_euler_without_guidance_gap():
    """eps_uc == eps_c: denoised == uncond_denoised, so x' = D + (x - D) * sigma' / sigma for any lambda."""
    tb, noise, uc, _, unet = _setup()
    sigmas = OSm.karras_sigmas(tb)
    x = OSm.kd_start_state(noise, sigmas).float()
    for lam in (0.0, 0.6, 1.0):
        _, x1 = OSm.kd_euler_cfgpp(unet, tb, x.clone(), sigmas[:2], uc, uc, lam)
        t = OSm.kd_timestep(tb, sigmas[0])
        den = x - unet(x / (sigmas[0] ** 2 + 1) ** 0.5, t.reshape(1), uc)["sample"] * sigmas[0]
        assert torch.allclose(x1, den + (x - den) * (sigmas[1] / sigmas[0]), rtol=1e-5, atol=1e-5)


def test_first_dpmpp2m_step_is_an_euler_step_and_karras_shape():
    tb, noise, uc, c, unet = _setup(nfe=6)
    sigmas = OSm.karras_sigmas(tb)
    assert len(sigmas) == 7 and sigmas[-1] == 0 and bool((sigmas[:-2] > sigmas[1:-1]).all())
    assert math.isclose(float(sigmas[0]), float(tb.sigmas.max()), rel_tol=1e-6)
    x0 = OSm.kd_start_state(noise, sigmas)
    _, xe = OSm.kd_euler_cfgpp(unet, tb, x0.clone(), sigmas[:2], uc, c, 0.6)
    _, xm = OSm.kd_dpmpp_2m_cfgpp_sd15(unet, tb, x0.clone(), sigmas[:2], uc, c, 0.6)
    assert torch.equal(xe, xm)


def test_edit_loop_round_trip_with_identical_prompts():
    """CFG++ inversion followed by CFG++ sampling under the SAME prompt returns near the source latent (the DDIM
    inversion approximation: eps evaluated one step off) — a sanity pin of the two loops' pairing and indexing."""
    tb, noise, uc, c, unet = _setup(nfe=50, hw=4)
    z0 = (0.3 * noise).half()
    zT, z0t = OSm.ddim_edit_cfgpp(unet, tb, z0, uc, c, c, 0.6)
    assert zT.dtype == torch.float16 and z0t.dtype == torch.float16
    assert float((z0t.float() - z0.float()).norm() / z0.float().norm()) < 0.25
    assert not torch.equal(zT, z0)


def test_plain_and_plus_coincide_without_guidance_gap_and_differ_otherwise():
    """eps_uc == eps_c: CFG and CFG++ are the same sampler; with a gap and lambda != 0 they are not."""
    tb, noise, uc, c, unet = _setup()
    sigmas = OSm.karras_sigmas(tb)
    x0 = OSm.kd_start_state(noise, sigmas)
    for fn in (OSm.kd_euler_cfgpp, OSm.kd_dpmpp_2m_cfgpp_sd15):
        a = fn(unet, tb, x0.clone(), sigmas, uc, uc, 0.6, plus=True)[1]
        b = fn(unet, tb, x0.clone(), sigmas, uc, uc, 0.6, plus=False)[1]
        assert _close16(a, b, ulps=4)
        a = fn(unet, tb, x0.clone(), sigmas, uc, c, 0.6, plus=True)[1]
        b = fn(unet, tb, x0.clone(), sigmas, uc, c, 0.6, plus=False)[1]
        assert not _close16(a, b, ulps=4)
    # DDIM: plain CFG == CFG++ when lambda = 0 ... no: CFG++ renoises with eps_uc, plain with the guided eps, which
    # coincide exactly when lambda = 0
    z = noise
    a = OSm.sd15_ddim_cfgpp(unet, tb, z, uc, c, 0.0)
    b = OSm.ddim_plain(unet, tb, z, uc, c, 0.0)
    assert torch.equal(a, b)
    assert not torch.equal(OSm.sd15_ddim_cfgpp(unet, tb, z, uc, c, 0.6), OSm.ddim_plain(unet, tb, z, uc, c, 0.6))

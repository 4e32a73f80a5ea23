"""Known-answer tests for the scheduler tables (SURVEY.md Appendix B constants) — for the oracle restatement and for
the product-side tables, plus the per-step coefficient tables against the oracle loops' own scalar arithmetic."""
import pytest
import torch

from cfgpp_b200 import schedule as PS
from oracle import schedule as OS


def close(a, b, tol=2e-7):
    return abs(float(a) - float(b)) <= tol * max(1.0, abs(float(b)))


@pytest.mark.parametrize("mod", [OS, PS])
def test_alphas_cumprod_known_answers(mod):
    abar = mod.alphas_cumprod_table()
    assert abar.dtype == torch.float32 and abar.shape == (1000,)
    assert close(abar[0], 0.9991499782) and close(abar[980], 0.0058437791) and close(abar[999], 0.0046600951)
    sig = (1 - abar).sqrt() / abar.sqrt()
    assert close(sig.min(), 0.0291675329, 1e-6) and close(sig.max(), 14.6146469116, 1e-6)


@pytest.mark.parametrize("mod", [OS, PS])
def test_timesteps(mod):
    t50 = mod.ddim_leading_timesteps(50)
    assert t50[0] == 981 and t50[1] == 961 and t50[-2] == 21 and t50[-1] == 1 and len(t50) == 50
    t10 = mod.ddim_leading_timesteps(10)
    assert t10.tolist() == [901, 801, 701, 601, 501, 401, 301, 201, 101, 1]
    t25 = mod.ddim_leading_timesteps(25)
    assert t25[0] == 961 and t25[1] == 921 and t25[-1] == 1
    assert mod.euler_trailing_timesteps(4).tolist() == [999.0, 749.0, 499.0, 249.0]


def test_shifted_table_and_lightning_constants():
    tb = OS.make_tables(50)
    sch = PS.Schedule.make(50)
    for acp in (tb.alphas_cumprod, sch.alphas_cumprod):
        assert acp.shape == (1001,) and acp[0] == 1.0
        assert close(acp[981], 0.00584378, 1e-6) and close(acp[961], 0.00736524, 1e-6)
    assert tb.skip == sch.skip == 20
    assert close(sch.final_alpha_cumprod, 0.99914998, 1e-7)
    tl = PS.Schedule.make(4, "lightning")
    vals = [float(tl.alphas_cumprod[int(t)]) for t in tl.timesteps]
    for v, e in zip(vals, [0.00471670, 0.05707992, 0.27900973, 0.67707050]):
        assert close(v, e, 2e-6)
    assert tl.skip == 250


def test_dpmpp_tables_known_answers():
    sch = PS.Schedule.make(25)
    steps, sigma0 = PS.dpmpp_2m_cfgpp_steps(sch, 0.6)
    assert len(steps) == 24  # loop runs over timesteps[:-1]  (latent_sdxl.py:890)
    assert close(sigma0, 11.60917377, 1e-6)
    assert [int(s.t) for s in steps[:3]] == [960, 920, 880]  # the UNet is fed t-1 (sigma_to_t on the un-shifted table)
    assert close(-steps[0].coef.c0, 11.60917377, 1e-6) and close(steps[0].coef.c2, 9.28758049, 1e-6)
    assert steps[0].coef.second_order == 0 and all(s.coef.second_order == 1 for s in steps[1:])
    assert close(steps[-1].coef.c2, 0.02916753, 1e-5)
    k = PS.get_sigmas_karras if hasattr(PS, "get_sigmas_karras") else OS.get_sigmas_karras
    ks = OS.get_sigmas_karras(50, 0.0291675329, 14.6146469116)
    assert close(ks[0], 14.61464310, 1e-5) and close(ks[1], 13.45211220, 1e-5) and ks[-1] == 0


def test_ddim_step_tables_match_reference_indexing():
    sch = PS.Schedule.make(50)
    sx = PS.ddim_cfgpp_steps(sch, 0.6, sdxl_indexing=True)
    sd = PS.ddim_cfgpp_steps(sch, 0.6, sdxl_indexing=False)
    acp = sch.alphas_cumprod
    assert len(sx) == len(sd) == 50 and int(sx[0].t) == 981 and int(sx[-1].t) == 1
    assert close(sx[0].coef.c1, acp[981].sqrt()) and close(sx[0].coef.c2, acp[961].sqrt())
    # last step: SDXL wraps to a negative index (t - skip = -19 -> acp[-19]); SD1.5 uses final_alpha_cumprod
    assert close(sx[-1].coef.c2, acp[-19].sqrt()) and close(sd[-1].coef.c2, sch.final_alpha_cumprod.sqrt())
    for a, b in zip(sx[:-1], sd[:-1]):
        assert (a.coef.c0, a.coef.c1, a.coef.c2, a.coef.c3) == (b.coef.c0, b.coef.c1, b.coef.c2, b.coef.c3)
    inv = PS.ddim_inversion_cfgpp_steps(sch, 0.6)
    assert int(inv[0].t) == 1 and int(inv[-1].t) == 981
    assert close(inv[0].coef.c1, sch.final_alpha_cumprod.sqrt()) and close(inv[0].coef.c2, acp[1].sqrt())


def test_sigma_to_t_quantized_and_interpolated():
    """latent_sdxl.py:333-346: nearest table index, or the fractional index between the two bracketing sigmas."""
    from cfgpp_b200 import schedule as S
    sch = S.Schedule.make(50)
    ts = (1 - sch.total_alphas).sqrt() / sch.total_alphas.sqrt()
    exact = ts[[3, 500, 998]]
    assert S.sigma_to_t(sch, exact).tolist() == [3, 500, 998]
    assert torch.allclose(S.sigma_to_t(sch, exact, quantize=False), torch.tensor([3., 500., 998.]), atol=1e-3)
    mid = (ts[200] + ts[201]) / 2
    t = S.sigma_to_t(sch, mid[None], quantize=False)
    assert 200.0 < t.item() < 201.0 and abs(t.item() - 200.5) < 0.05
    assert S.sigma_to_t(sch, torch.tensor([1e6]), quantize=False).item() == 999.0   # clamped above the table


def test_kd_ancestral_schedule_layout():
    """euler_a: one entry per step; dpm++_2s_a: (midpoint, final) pairs while sigma_down > 0, the Euler form on the last
    step; one noise slot per step with sigma_next > 0, numbered in loop order (latent_diffusion.py:744-762, :782-825)."""
    from cfgpp_b200 import kdiffusion as K, schedule as S
    sigmas = K.get_sigmas_karras(5, 0.03, 14.6, rho=7.)
    tfn = lambda s: torch.tensor(321)  # noqa: E731
    for cfgpp in (True, False):
        base = 0 if cfgpp else S.KD_EXTRAP_GUIDED
        steps, slots = S.kd_ancestral_steps(sigmas, tfn, 0.7, cfgpp, two_s=False)
        assert slots == 4 and len(steps) == 5
        for i, st in enumerate(steps):
            down, up = K.get_ancestral_step(sigmas[i], sigmas[i + 1])
            assert st.coef.second_order == base | (S.KD_NOISE if i < 4 else 0)
            assert st.coef.c2 == pytest.approx(float(down)) and st.coef.d3 == pytest.approx(float(up) if i < 4 else 0.0)
            assert st.coef.c0 == pytest.approx(-float(sigmas[i])) and (i == 4 or int(st.coef.c3) == i)
        steps, slots = S.kd_ancestral_steps(sigmas, tfn, 0.7, cfgpp, two_s=True)
        assert slots == 4 and len(steps) == 9
        bits = [st.coef.second_order for st in steps]
        assert bits == [base | S.KD_2S_MID, base | S.KD_2S_FINAL | S.KD_NOISE] * 4 + [base]
        for i in range(4):
            mid, fin = steps[2 * i], steps[2 * i + 1]
            down, up = K.get_ancestral_step(sigmas[i], sigmas[i + 1])
            h = -torch.log(down) + torch.log(sigmas[i])
            assert fin.coef.d0 == pytest.approx(float(torch.exp(-h)), rel=1e-5)
            assert fin.coef.d1 == pytest.approx(float(down / sigmas[i]), rel=1e-5) and int(fin.coef.c3) == i
            sigma_s = float((sigmas[i] * down) ** 0.5)                  # geometric midpoint in log-sigma (r = 1/2)
            assert -fin.coef.c0 == pytest.approx(sigma_s, rel=1e-5) and mid.coef.d0 == pytest.approx(sigma_s / float(sigmas[i]), rel=1e-5)
            assert fin.in_scale == pytest.approx(1.0 / (sigma_s ** 2 + 1) ** 0.5, rel=1e-5)

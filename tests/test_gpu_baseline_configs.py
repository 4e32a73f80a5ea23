"""Parity at the geometry and batch of the BASELINE.json configurations themselves, plus the multi-prompt regression.

configs[1]  SD v1.5  64x64 latent   batch 4 (UNet batch 8)          -> test_unet_forward_sd15_b4
configs[2]  SDXL     128x128 latent batch 2 (UNet batch 4)          -> test_unet_forward_sdxl_b2 (the bench workload)
configs[3]  SDXL     dpm++_2m_cfgpp NFE=25                          -> test_dpmpp_nfe25_teacher_forced_full_sdxl
configs[4]  Lightning ddim_cfg++_lightning NFE=4 lambda=1 batch 8   -> test_lightning_nfe4_trajectory_vs_oracle (+ B=8)

Tolerances are the stated ones: forward rel-L2 <= 5e-3 vs the fp16-autocast oracle and not further
from the fp32 oracle than 1.5x the fp16 oracle itself; teacher-forced steps <= 5e-3; free-running <= 3e-2.
"""
from types import SimpleNamespace

import pytest
import torch

from helpers import OracleCudaUNet, build_pair, make_inputs, oracle_cfg, rel_l2

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def _forward_case(name, B, hw, t, fp32_check=True):
    from oracle import unet as O
    cfg, sd, net, ref16 = build_pair(name, dev)
    z, uc, c, add = make_inputs(cfg, B, hw, dev)
    net.prepare(B, hw, hw)
    net.set_prompt(torch.cat([uc, c]), add["text_embeds"] if add else None, add["time_ids"].float() if add else None)
    eu, ec = net.predict_noise(z, float(t))
    got = torch.cat([eu, ec]).float()
    net.close()
    z_in, t_in, ctx = torch.cat([z] * 2), torch.tensor(t, device=dev), torch.cat([uc, c])
    r16 = ref16(z_in, t_in, ctx, add)["sample"].float()
    del ref16
    e16 = rel_l2(got, r16)
    assert torch.isfinite(got).all() and e16 <= 5e-3, e16
    # every sample of the batch individually (a row mix-up inside the batch hides in the aggregate norm)
    for i in range(2 * B):
        assert rel_l2(got[i], r16[i]) <= 5e-3, (i, rel_l2(got[i], r16[i]))
    if fp32_check:
        m32 = O.build_unet(oracle_cfg(cfg), sd, dtype=torch.float32, device=dev)
        with torch.no_grad():
            r32 = m32(z_in, t_in, ctx.float(), {k: v.float() for k, v in add.items()} if add else None)["sample"]
        del m32
        e32, b32 = rel_l2(got, r32), rel_l2(r16, r32)
        print(f"{name} B={B} {hw}x{hw}: vs fp16 oracle {e16:.3e}, vs fp32 oracle {e32:.3e} (fp16 oracle itself {b32:.3e})")
        assert e32 <= 1.5 * b32 + 1e-4


def test_unet_forward_sdxl_b2():
    """configs[2] — the bench workload: SDXL at 128x128 latent, batch 2 => UNet batch 4."""
    _forward_case("sdxl", 2, 128, 501)


def test_unet_forward_sd15_b4():
    """configs[1]: SD v1.5 at 64x64 latent, batch 4 => UNet batch 8."""
    _forward_case("sd15", 4, 64, 401)


def test_unet_forward_tiny_sdxl_b8():
    """configs[4] batch: 8 images => UNet batch 16 (the plan's maximum), tiny SDXL so it stays cheap."""
    _forward_case("tiny_sdxl", 8, 32, 249)


def test_lightning_nfe4_trajectory_vs_oracle():
    """configs[4] solver: ddim_cfg++_lightning, NFE=4, lambda=1 (latent_sdxl.py:838-858). Own arithmetic: trailing
    timesteps, schedule table resident on the DEVICE (:418) so the two eps scalars round through fp16, un-duplicated
    added conditions (cfg_guidance == 1, :249-252). Teacher-forced every step, then free-running, then through the
    registered solver class at batch 8."""
    from cfgpp_b200 import latent_sdxl as LX, schedule as S
    from oracle import samplers as OSm, schedule as OS
    cfg, sd, net, ref = build_pair("tiny_sdxl", dev)
    B, hw, nfe, lam = 2, 32, 4, 1.0
    z, uc, c, add = make_inputs(cfg, B, hw, dev, duplicate_added=False)
    tb = OS.make_tables(nfe, "lightning")
    rec = []
    # The reference passes (1, .) added conditions at cfg_guidance == 1 and lets them broadcast over the UNet batch of
    # 2; that only works for ONE image. The engine generalises it to `batch` rows (row r of either half uses row
    # r % batch); the oracle is given the equivalent duplicated form.
    dup = lambda a: {k: torch.cat([v, v]) for k, v in a.items()}  # noqa: E731
    z0_ref = OSm.sdxl_ddim_cfgpp_lightning(ref, tb, z, uc, c, lam, dup(add), record=rec)
    steps = S.ddim_cfgpp_steps(S.Schedule.make(nfe, "lightning"), lam, sdxl_indexing=True, tables_on_device=True)
    assert [int(s.t) for s in steps] == [int(t) for t in tb.timesteps]
    net.prepare(B, hw, hw)
    net.set_prompt(torch.cat([uc, c]), add["text_embeds"], add["time_ids"].float())
    net.set_schedule(S.STEP_DDIM_CFGPP, torch.float32, steps)
    for i, r in enumerate(rec):
        net.set_state(r["zt"])
        eu, ec = net.predict_noise(r["zt"], steps[i].t)
        assert rel_l2(eu, r["noise_uc"]) <= 5e-3 and rel_l2(ec, r["noise_c"]) <= 5e-3
        net.run_steps(i, 1)
        if i + 1 < len(rec):
            assert rel_l2(net.get_state(0), rec[i + 1]["zt"]) <= 5e-3, i
    net.set_state(z)
    net.run_steps(0, nfe)
    e = rel_l2(net.get_state(1), z0_ref)
    print(f"lightning NFE=4 lambda=1 free-running: rel-L2 final z0t {e:.3e}")
    assert e <= 3e-2
    net.close()
    # the registered solver, batch 8 (configs[4]), against the oracle loop
    z8, uc8, c8, add8 = make_inputs(cfg, 8, hw, dev, seed=21, duplicate_added=False)
    with pytest.warns(UserWarning):
        lt = LX.get_solver("ddim_cfg++_lightning", solver_config=SimpleNamespace(num_sampling=nfe), device=dev,
                           unet_config=cfg, state_dict=sd)
    got = lt.reverse_process(uc8, c8, 1.0, add8, shape=(8 * hw, 8 * hw), zT=z8)
    want = OSm.sdxl_ddim_cfgpp_lightning(ref, tb, z8, uc8, c8, 1.0, dup(add8))
    e8 = rel_l2(got, want)
    print(f"lightning solver batch 8: rel-L2 final z0t {e8:.3e}")
    assert got.shape == (8, 4, hw, hw) and e8 <= 3e-2
    LX.release_engines()


def test_dpmpp_nfe25_teacher_forced_full_sdxl():
    """configs[3] solver at its real size: SDXL 128x128 latent, dpm++_2m_cfgpp, NFE=25 (24 steps, latent_sdxl.py:890).
    Teacher-forced on the oracle's states for the first-order step, two second-order steps and the last step:
    eps <= 5e-3, and the fused update reproduces the oracle's next state to <= 5e-3."""
    from cfgpp_b200 import schedule as S
    from oracle import samplers as OSm, schedule as OS
    cfg, sd, net, ref = build_pair("sdxl", dev)
    B, hw, nfe, lam = 1, 128, 25, 0.6
    z, uc, c, add = make_inputs(cfg, B, hw, dev)
    tb = OS.make_tables(nfe)
    steps, sigma0 = S.dpmpp_2m_cfgpp_steps(S.Schedule.make(nfe), lam)
    assert len(steps) == nfe - 1
    net.prepare(B, hw, hw)
    net.set_prompt(torch.cat([uc, c]), add["text_embeds"], add["time_ids"].float())
    net.set_schedule(S.STEP_DPMPP2M_CFGPP, torch.float16, steps)
    # oracle trajectory truncated after 4 steps (each full-size oracle step is a 4-sample eager forward)
    import dataclasses
    rec = []
    tb4 = dataclasses.replace(tb, timesteps=tb.timesteps[:5])
    OSm.sdxl_dpmpp_2m_cfgpp(ref, tb4, z, uc, c, lam, add, record=rec)
    assert len(rec) == 4
    x0 = z.to(torch.float16) * sigma0
    assert torch.equal(x0, rec[0]["x"])
    net.set_state(x0)
    for i in range(3):
        r = rec[i]
        eu, ec = net.predict_noise(r["x"], steps[i].t, steps[i].in_scale)
        e_u, e_c = rel_l2(eu, r["noise_uc"]), rel_l2(ec, r["noise_c"])
        assert e_u <= 5e-3 and e_c <= 5e-3, (i, e_u, e_c)
        net.set_state(r["x"])      # teacher-force the state (old_denoised carried by the engine from step i-1)
        net.run_steps(i, 1)
        e_x = rel_l2(net.get_state(0), rec[i + 1]["x"])
        print(f"dpm++ NFE=25 full SDXL step {i}: eps {e_u:.2e}/{e_c:.2e}, next x {e_x:.2e}")
        assert e_x <= 5e-3, (i, e_x)
    net.close()


def test_two_prompts_back_to_back_on_one_solver():
    """Regression for the stale-conditioning bug: one solver, prompt A then prompt B (the text_to_mscoco.py loop,
    reference examples/text_to_mscoco.py:54-62). The second image must equal a fresh solver's image for B and differ
    from A's. Freshly allocated embeddings come back at recycled addresses, which is exactly what an address-keyed
    cache cannot tell apart."""
    from cfgpp_b200 import latent_diffusion as LD, latent_sdxl as LX
    from cfgpp_b200.config import tiny_sd15_config, tiny_sdxl_config
    from cfgpp_b200.utils.log_util import set_seed
    for mod, cfgf, mk, kw_a, kw_b in (
            (LX, tiny_sdxl_config, "synthetic:7",
             dict(prompt1=["", "a cat"], prompt2=["", "a cat"], target_size=(256, 256)),
             dict(prompt1=["", "a dog on a skateboard"], prompt2=["", "a dog on a skateboard"], target_size=(256, 256))),
            (LD, tiny_sd15_config, "synthetic:9", dict(prompt=["", "a cat"]), dict(prompt=["", "a dog on a skateboard"]))):
        for method in ("ddim_cfg++", "euler_cfg++"):   # fused-trajectory path and the per-step predict_noise path
            mk_solver = lambda: mod.get_solver(method, solver_config=SimpleNamespace(num_sampling=4), device="cuda:0",  # noqa: E731
                                               unet_config=cfgf(), model_key=mk)
            s = mk_solver()
            set_seed(42)
            img_a = s.sample(cfg_guidance=0.6, **kw_a)
            set_seed(42)
            img_b = s.sample(cfg_guidance=0.6, **kw_b)
            del s
            LX.release_engines()
            set_seed(42)
            img_b_fresh = mk_solver().sample(cfg_guidance=0.6, **kw_b)
            LX.release_engines()
            assert torch.equal(img_b, img_b_fresh), f"{mod.__name__}:{method}: second prompt used stale conditioning"
            assert not torch.equal(img_a, img_b), f"{mod.__name__}:{method}: the prompt has no influence"


def test_two_solvers_sharing_one_engine_do_not_leak_prompts():
    """Solver A binds P1; solver B (same cached engine) binds P2 and changes the batch shape; A runs again."""
    from cfgpp_b200 import latent_sdxl as LX
    cfg, sd, net, _ = build_pair("tiny_sdxl", dev)
    net.close()
    kw = dict(solver_config=SimpleNamespace(num_sampling=3), device=dev, unet_config=cfg, state_dict=sd)
    a, b = LX.get_solver("ddim_cfg++", **kw), LX.get_solver("ddim_cfg++", **kw)
    assert a.unet is b.unet
    z1, uc1, c1, add1 = make_inputs(cfg, 1, 32, dev, seed=1)
    z2, uc2, c2, add2 = make_inputs(cfg, 2, 32, dev, seed=2)
    r1 = a.reverse_process(uc1, c1, 0.6, add1, shape=(256, 256), zT=z1)
    b.reverse_process(uc2, c2, 0.6, add2, shape=(256, 256), zT=z2)
    r1_again = a.reverse_process(uc1, c1, 0.6, add1, shape=(256, 256), zT=z1)
    assert torch.equal(r1, r1_again)
    LX.release_engines()


def test_failed_prepare_forces_replan():
    """a prepare() that throws (unsupported latent) must not leave the handle 'prepared' on freed buffers."""
    from cfgpp_b200 import _native as nv
    cfg, sd, net, _ = build_pair("tiny_sdxl", dev)
    z, uc, c, add = make_inputs(cfg, 1, 32, dev)
    net.prepare(1, 32, 32)
    net.set_prompt(torch.cat([uc, c]), add["text_embeds"], add["time_ids"].float())
    a = net.predict_noise(z, 500.0)
    with pytest.raises(nv.NativeError):
        net.prepare(1, 24, 24)      # width 24 is not a power of two: rejected before anything is freed
    with pytest.raises(nv.NativeError):
        net.prepare(9, 32, 32)      # UNet batch 18 > 16
    net.prepare(1, 32, 32)
    net.set_prompt(torch.cat([uc, c]), add["text_embeds"], add["time_ids"].float())
    b = net.predict_noise(z, 500.0)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    net.close()

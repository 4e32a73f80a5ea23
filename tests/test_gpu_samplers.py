"""Sampler-level parity: the fused CFG++ step kernel against the oracle loops' arithmetic, teacher-forced and
free-running trajectories, fused-graph vs un-fused seam equivalence, and the solver API end to end.

Tolerances (stated): step kernel given identical eps — DDIM fp32 state rel <= 1e-6 (observed bit-exact up to the
final divide), DPM++ fp16 state <= 2 fp16 ulp; teacher-forced per-step z_{t-1} rel-L2 <= 5e-3; free-running final
latent rel-L2 reported and loosely gated (chaotic divergence at fp16 noise level, SURVEY §7 hard part 3)."""
from types import SimpleNamespace

import pytest
import torch

from helpers import OracleCudaUNet, build_pair, make_inputs, rel_l2

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def _ulp16(x):
    x = x.float().abs().clamp_min(6.1e-5)
    return torch.pow(2.0, torch.floor(torch.log2(x)) - 10)


def test_step_kernel_ddim_matches_reference_arithmetic():
    from cfgpp_b200 import _native as nv, schedule as S
    g = torch.Generator().manual_seed(0)
    sch = S.Schedule.make(50)
    steps = S.ddim_cfgpp_steps(sch, 0.6, True)
    for idx in (0, 17, 49):
        st = steps[idx]
        zt = torch.randn(2, 4, 32, 32, generator=g).to(dev)
        eu = torch.randn(2, 4, 32, 32, generator=g).half().to(dev)
        ec = torch.randn(2, 4, 32, 32, generator=g).half().to(dev)
        t = int(st.t)
        at, an = sch.alphas_cumprod[t], sch.alphas_cumprod[t - sch.skip]  # CPU 0-dim tensors, as in the reference
        npred = eu + 0.6 * (ec - eu)                             # latent_sdxl.py:738 (fp16 tensor ops)
        z0_ref = (zt - (1 - at).sqrt() * npred) / at.sqrt()      # :741
        zn_ref = an.sqrt() * z0_ref + (1 - an).sqrt() * eu       # :744
        z = zt.clone()
        z0 = nv.op_cfgpp_step(eu, ec, S.STEP_DDIM_CFGPP, st.coef, z)
        assert z0.dtype == torch.float32
        assert (z0 - z0_ref).abs().max() <= 1e-6 * z0_ref.abs().max()
        assert (z - zn_ref).abs().max() <= 1e-6 * zn_ref.abs().max()
    # inversion: Tweedie with eps_uc, renoise with the guided eps (latent_diffusion.py:907-908), fp16 state
    inv = S.ddim_inversion_cfgpp_steps(sch, 0.6)[5]
    t = int(inv.t)
    at, ap = sch.alpha(t), sch.alpha(t - sch.skip)
    zt = torch.randn(1, 4, 32, 32, generator=g).half().to(dev)
    eu = torch.randn(1, 4, 32, 32, generator=g).half().to(dev)
    ec = torch.randn(1, 4, 32, 32, generator=g).half().to(dev)
    npred = eu + 0.6 * (ec - eu)
    z0_ref = (zt - (1 - ap).sqrt() * eu) / ap.sqrt()
    zn_ref = at.sqrt() * z0_ref + (1 - at).sqrt() * npred
    z = zt.clone()
    nv.op_cfgpp_step(eu, ec, S.STEP_DDIM_INV_CFGPP, inv.coef, z)
    assert z.dtype == torch.float16 and ((z.float() - zn_ref.float()).abs() <= 2 * _ulp16(zn_ref)).all()


def test_step_kernel_dpmpp_matches_reference_arithmetic():
    from cfgpp_b200 import _native as nv, schedule as S
    g = torch.Generator().manual_seed(1)
    sch = S.Schedule.make(25)
    steps, sigma0 = S.dpmpp_2m_cfgpp_steps(sch, 0.6)
    alphas = sch.alphas_cumprod[sch.timesteps.int()]
    sigmas = (1 - alphas).sqrt() / alphas.sqrt()
    t_fn = lambda s: s.log().neg()  # noqa: E731
    x = (torch.randn(1, 4, 32, 32, generator=g).half() * sigma0).to(dev)
    old = None
    xs_native = x.clone()
    aux = torch.zeros_like(x)
    for i in range(4):
        eu = torch.randn(1, 4, 32, 32, generator=g).half().to(dev)
        ec = torch.randn(1, 4, 32, 32, generator=g).half().to(dev)
        # reference arithmetic, latent_sdxl.py:902-919
        npred = eu + 0.6 * (ec - eu)
        c_out = -sigmas[i].clone()
        den, ud = x + c_out * npred, x + c_out * eu
        t, tn = t_fn(sigmas[i]), t_fn(sigmas[i + 1])
        h = tn - t
        if old is None or sigmas[i + 1] == 0:
            xr = den + (x - ud) / sigmas[i].item() * sigmas[i + 1]
        else:
            r = (t - t_fn(sigmas[i - 1])) / h
            xr = den + (-torch.exp(-h) * ud - (-h).expm1() * (ud - old) / (2 * r)) + torch.exp(-h) * x
        old = ud
        nv.op_cfgpp_step(eu, ec, S.STEP_DPMPP2M_CFGPP, steps[i].coef, xs_native, aux, want_z0t=False)
        assert xs_native.dtype == torch.float16
        assert ((xs_native.float() - xr.float()).abs() <= 2 * _ulp16(xr)).all(), f"step {i}"
        assert ((aux.float() - ud.float()).abs() <= 1 * _ulp16(ud)).all()
        x = xr
        xs_native.copy_(xr)  # teacher-force so the per-step bound is meaningful
        aux.copy_(ud)


@pytest.mark.parametrize("name", ["tiny_sdxl", "tiny_sd15"])
def test_teacher_forced_and_free_running_ddim(name):
    from cfgpp_b200 import schedule as S
    from oracle import samplers as OSm, schedule as OS
    cfg, sd, net, ref = build_pair(name, dev)
    B, hw, nfe, lam = 1, 32, 10, 0.6
    z, uc, c, add = make_inputs(cfg, B, hw, dev)
    tb = OS.make_tables(nfe)
    rec = []
    if cfg.addition_embed_type:
        z0_ref = OSm.sdxl_ddim_cfgpp(ref, tb, z, uc, c, lam, add, record=rec)
    else:
        z0_ref = OSm.sd15_ddim_cfgpp(ref, tb, z, uc, c, lam, record=rec)
    sch = S.Schedule.make(nfe)
    steps = S.ddim_cfgpp_steps(sch, lam, sdxl_indexing=bool(cfg.addition_embed_type))
    net.prepare(B, hw, hw)
    net.set_prompt(torch.cat([uc, c]), add["text_embeds"] if add else None, add["time_ids"].float() if add else None)
    net.set_schedule(S.STEP_DDIM_CFGPP, torch.float32, steps)
    # teacher-forced: feed the oracle's z_t, compare eps and z_{t-1}
    for i, r in enumerate(rec):
        net.set_state(r["zt"])
        eu, ec = net.predict_noise(r["zt"], steps[i].t)
        assert rel_l2(eu, r["noise_uc"]) <= 5e-3 and rel_l2(ec, r["noise_c"]) <= 5e-3
        net.run_steps(i, 1)
        if i + 1 < len(rec):
            assert rel_l2(net.get_state(0), rec[i + 1]["zt"]) <= 5e-3
    # free-running fused trajectory
    net.set_state(z)
    net.run_steps(0, nfe)
    z0 = net.get_state(1)
    e = rel_l2(z0, z0_ref)
    print(f"free-running {name} NFE={nfe}: rel-L2(final z0t) = {e:.3e}")
    assert e <= 3e-2
    net.close()


def test_fused_graph_equals_unfused_seam_bit_exact():
    from cfgpp_b200 import schedule as S
    cfg, sd, net, _ = build_pair("tiny_sdxl", dev)
    z, uc, c, add = make_inputs(cfg, 2, 32, dev)
    sch = S.Schedule.make(6)
    steps = S.ddim_cfgpp_steps(sch, 0.6, True)
    net.prepare(2, 32, 32)
    net.set_prompt(torch.cat([uc, c]), add["text_embeds"], add["time_ids"].float())
    net.set_schedule(S.STEP_DDIM_CFGPP, torch.float32, steps)
    net.set_state(z)
    net.run_steps(0, 6)
    a0, a1 = net.get_state(0).clone(), net.get_state(1).clone()
    net.set_state(z)
    for i, st in enumerate(steps):
        eu, ec = net.predict_noise(net.get_state(0), st.t)
        net.apply_step(i, eu, ec)
    assert torch.equal(a0, net.get_state(0)) and torch.equal(a1, net.get_state(1))
    net.close()


def test_dpmpp_trajectory_vs_oracle():
    from cfgpp_b200 import schedule as S
    from oracle import samplers as OSm, schedule as OS
    cfg, sd, net, ref = build_pair("tiny_sdxl", dev)
    B, hw, nfe, lam = 1, 32, 8, 0.6
    z, uc, c, add = make_inputs(cfg, B, hw, dev)
    tb = OS.make_tables(nfe)
    rec = []
    x_ref = OSm.sdxl_dpmpp_2m_cfgpp(ref, tb, z, uc, c, lam, add, record=rec)
    sch = S.Schedule.make(nfe)
    steps, sigma0 = S.dpmpp_2m_cfgpp_steps(sch, lam)
    assert len(steps) == nfe - 1 == len(rec)
    net.prepare(B, hw, hw)
    net.set_prompt(torch.cat([uc, c]), add["text_embeds"], add["time_ids"].float())
    net.set_schedule(S.STEP_DPMPP2M_CFGPP, torch.float16, steps)
    x0 = z.to(torch.float16) * sigma0
    assert torch.equal(x0, rec[0]["x"])
    for i, r in enumerate(rec):   # teacher-forced eps (UNet sees x * c_in at t-1)
        eu, ec = net.predict_noise(r["x"], steps[i].t, steps[i].in_scale)
        assert rel_l2(eu, r["noise_uc"]) <= 5e-3 and rel_l2(ec, r["noise_c"]) <= 5e-3
    net.set_state(x0)
    net.run_steps(0, len(steps))
    e = rel_l2(net.get_state(0), x_ref)
    print(f"free-running dpm++_2m_cfgpp NFE={nfe}: rel-L2(final x) = {e:.3e}")
    assert e <= 3e-2
    net.close()


def test_inversion_then_sampling_sd15_vs_oracle():
    from cfgpp_b200 import schedule as S
    from oracle import samplers as OSm, schedule as OS
    cfg, sd, net, ref = build_pair("tiny_sd15", dev)
    B, hw, nfe, lam = 1, 32, 6, 0.6
    z, uc, c, _ = make_inputs(cfg, B, hw, dev)
    z0_src = (0.5 * z).half()  # the VAE latent is fp16 -> fp16 state throughout
    tb = OS.make_tables(nfe)
    zT_ref = OSm.sd15_inversion_cfgpp(ref, tb, z0_src, uc, c, lam)
    sch = S.Schedule.make(nfe)
    net.prepare(B, hw, hw)
    net.set_prompt(torch.cat([uc, c]))
    net.set_schedule(S.STEP_DDIM_INV_CFGPP, torch.float16, S.ddim_inversion_cfgpp_steps(sch, lam))
    net.set_state(z0_src)
    net.run_steps(0, nfe)
    zT = net.get_state(0)
    assert zT.dtype == torch.float16 and rel_l2(zT, zT_ref) <= 3e-2
    net.close()


def test_solver_api_end_to_end_and_callback_path():
    from cfgpp_b200 import latent_diffusion as LD, latent_sdxl as LX
    from cfgpp_b200.config import tiny_sd15_config, tiny_sdxl_config
    from cfgpp_b200.utils.callback_util import get_callback
    from cfgpp_b200.utils.log_util import set_seed
    conf = SimpleNamespace(num_sampling=5)
    kw = dict(solver_config=conf, device="cuda:0", unet_config=tiny_sdxl_config(), model_key="synthetic:7")
    s = LX.get_solver("ddim_cfg++", **kw)
    set_seed(42)
    img = s.sample(prompt1=["", "a cat"], prompt2=["", "a cat"], cfg_guidance=0.6, target_size=(256, 256))
    assert img.shape == (1, 3, 256, 256) and img.device.type == "cpu" and 0 <= img.min() and img.max() <= 1
    set_seed(42)
    cb = get_callback("record", frequency=1)
    img_cb = s.sample(prompt1=["", "a cat"], prompt2=["", "a cat"], cfg_guidance=0.6, target_size=(256, 256),
                      callback_fn=cb)
    assert len(cb.records) == 5 and torch.equal(img, img_cb)  # same kernels on both paths
    # lambda = 0: the cond branch must not influence the result (CFG++ == unconditional DDIM)
    set_seed(42)
    a = s.sample(prompt1=["", "a cat"], prompt2=["", "a cat"], cfg_guidance=0.0, target_size=(256, 256))
    # dpm++ runs NFE-1 steps and returns x
    d = LX.get_solver("dpm++_2m_cfgpp", **kw)
    set_seed(42)
    img_d = d.sample(prompt1=["", "a cat"], prompt2=["", "a cat"], cfg_guidance=0.6, target_size=(256, 256))
    assert img_d.shape == (1, 3, 256, 256) and torch.isfinite(img_d).all()
    with pytest.warns(UserWarning):
        lt = LX.get_solver("ddim_cfg++_lightning", solver_config=SimpleNamespace(num_sampling=4), device="cuda:0",
                           unet_config=tiny_sdxl_config())
    with pytest.raises(AssertionError, match="CFG should be turned off"):
        lt.sample(prompt1=["", "x"], prompt2=["", "x"], cfg_guidance=0.6, target_size=(256, 256))
    set_seed(42)
    assert torch.isfinite(lt.sample(prompt1=["", "x"], prompt2=["", "x"], cfg_guidance=1.0, target_size=(256, 256))).all()
    # SD v1.5 family
    sd = LD.get_solver("ddim_cfg++", solver_config=conf, device="cuda:0", unet_config=tiny_sd15_config(),
                       model_key="synthetic:9")
    set_seed(42)
    im = sd.sample(prompt=["", "a dog"], cfg_guidance=0.6)
    assert im.shape == (1, 3, 256, 256) and torch.isfinite(im).all()
    inv = LD.get_solver("ddim_inversion_cfg++", solver_config=conf, device="cuda:0", unet_config=tiny_sd15_config(),
                        model_key="synthetic:9")
    src = torch.rand(1, 3, 256, 256) * 2 - 1
    im2 = inv.sample(src_img=src, prompt=["", "a dog"], cfg_guidance=0.6)
    assert im2.shape == (1, 3, 256, 256) and torch.isfinite(im2).all()
    assert a.shape == img.shape


# ---- the VE-cast CFG++ samplers and the CFG++ editing loops on the native UNet seam ----------

@pytest.mark.parametrize("method", ["euler_cfg++", "euler_a_cfg++", "dpm++_2s_a_cfg++", "dpm++_2m_cfg++"])
def test_sd15_kdiffusion_cfgpp_solvers_vs_oracle(method):
    """Free-running trajectories (NFE=6, Karras sigmas, fp16 state): product solver (native UNet behind predict_noise)
    vs the oracle loop on the fp16-autocast oracle UNet, same start state, same CUDA noise stream for the ancestral
    variants. Stated tolerance: rel-L2 of the final state / Tweedie estimate <= 3e-2 (as for the DDIM trajectories)."""
    from cfgpp_b200 import latent_diffusion as LD
    from oracle import samplers as OSm, schedule as OS
    cfg, sd, net, ref = build_pair("tiny_sd15", dev)
    net.close()
    nfe, lam, hw = 6, 0.6, 32
    z, uc, c, _ = make_inputs(cfg, 1, hw, dev)
    solver = LD.get_solver(method, solver_config=SimpleNamespace(num_sampling=nfe), device=dev, unet_config=cfg,
                           state_dict=sd)
    tb = OS.make_tables(nfe)
    sigmas = solver.karras_sigmas()
    assert torch.equal(sigmas, OSm.karras_sigmas(tb))
    x0 = OSm.kd_start_state(z, sigmas)
    oracle_loop = {"euler_cfg++": lambda: OSm.kd_euler_cfgpp(ref, tb, x0.clone(), sigmas, uc, c, lam),
                   "euler_a_cfg++": lambda: OSm.kd_euler_cfgpp(ref, tb, x0.clone(), sigmas, uc, c, lam, ancestral=True),
                   "dpm++_2s_a_cfg++": lambda: OSm.kd_dpmpp_2s_a_cfgpp(ref, tb, x0.clone(), sigmas, uc, c, lam),
                   "dpm++_2m_cfg++": lambda: OSm.kd_dpmpp_2m_cfgpp_sd15(ref, tb, x0.clone(), sigmas, uc, c, lam)}[method]
    torch.manual_seed(123)
    d_ref, x_ref = oracle_loop()
    torch.manual_seed(123)
    d, x = solver.reverse_process(uc, c, lam, x0.clone())
    assert x.dtype == torch.float16 and d.dtype == torch.float16
    e_x, e_d = rel_l2(x, x_ref), rel_l2(d, d_ref)
    print(f"{method}: rel-L2 final x {e_x:.3e}, last denoised {e_d:.3e}")
    assert e_x <= 3e-2 and e_d <= 3e-2
    img = solver.sample(cfg_guidance=lam, prompt=["", "a dog"])
    assert img.shape == (1, 3, 8 * cfg.sample_size, 8 * cfg.sample_size) and torch.isfinite(img).all()


def test_sdxl_euler_and_edit_cfgpp_vs_oracle():
    from cfgpp_b200 import latent_sdxl as LX
    from oracle import samplers as OSm, schedule as OS
    cfg, sd, net, ref = build_pair("tiny_sdxl", dev)
    net.close()
    nfe, lam, hw = 5, 0.6, 32
    z, uc, c, add = make_inputs(cfg, 1, hw, dev)
    g = torch.Generator().manual_seed(5)
    c_tgt = torch.randn(1, 77, cfg.cross_attention_dim, generator=g).half().to(dev)
    add_tgt = {"text_embeds": torch.randn(2, cfg.pooled_dim, generator=g).half().to(dev), "time_ids": add["time_ids"]}
    tb = OS.make_tables(nfe)
    kw = dict(solver_config=SimpleNamespace(num_sampling=nfe), device=dev, unet_config=cfg, state_dict=sd)
    # euler_cfg++ on the sampling timesteps' sigmas
    eul = LX.get_solver("euler_cfg++", **kw)
    sigmas = OSm.sdxl_euler_sigmas(tb)
    x0 = OSm.kd_start_state(z, sigmas)
    z0_ref, _ = OSm.kd_euler_cfgpp(ref, tb, x0.clone(), sigmas, uc, c, lam, add)
    z0 = eul.reverse_process(uc, c, lam, add, shape=(8 * hw, 8 * hw), xT=x0.clone())
    e = rel_l2(z0, z0_ref)
    print(f"sdxl euler_cfg++: rel-L2 last z0t {e:.3e}")
    assert z0.dtype == torch.float16 and e <= 3e-2
    # ddim_edit_cfg++: CFG++ inversion under the source prompt, CFG++ DDIM under the target prompt (fused step modes)
    ed = LX.get_solver("ddim_edit_cfg++", **kw)
    z0_src = (0.4 * z).half()
    ed.encode = lambda img: z0_src  # the VAE is outside the path: inject its latent
    zT_ref, z0t_ref = OSm.ddim_edit_cfgpp(ref, tb, z0_src, uc, c, c_tgt, lam, dict(add), dict(add_tgt))
    zT = ed.inversion(z0_src, uc, c, lam, dict(add))
    assert zT.dtype == torch.float16 and rel_l2(zT, zT_ref) <= 3e-2
    z0t = ed.reverse_process(uc, c, c_tgt, lam, dict(add), dict(add_tgt), src_img=torch.zeros(1, 3, 8 * hw, 8 * hw))
    e = rel_l2(z0t, z0t_ref)
    print(f"sdxl ddim_edit_cfg++: rel-L2 zT {rel_l2(zT, zT_ref):.3e}, edited z0t {e:.3e}")
    assert z0t.dtype == torch.float16 and e <= 3e-2
    # Lightning flavours: registry + the guidance assertion
    with pytest.warns(UserWarning):
        lt = LX.get_solver("dpm++_2m_cfgpp_lightning", solver_config=SimpleNamespace(num_sampling=4), device=dev,
                           unet_config=cfg)
    with pytest.raises(AssertionError, match="CFG should be turned off"):
        lt.reverse_process(uc, c, 0.6, add)
    with pytest.warns(UserWarning):
        le = LX.get_solver("euler_cfg++_lightning", solver_config=SimpleNamespace(num_sampling=4), device=dev,
                           unet_config=cfg)
    out = le.reverse_process(uc, c, 1.0, {k: v[-1:].clone() for k, v in add.items()}, shape=(8 * hw, 8 * hw))
    assert out.shape == (1, 4, hw, hw) and torch.isfinite(out).all()


# ---- plain-CFG baselines (fused STEP_DDIM_CFG mode, plain k-diffusion loops) -----------------

def test_step_kernel_plain_cfg_matches_reference_arithmetic():
    from cfgpp_b200 import _native as nv, schedule as S
    g = torch.Generator().manual_seed(3)
    sch = S.Schedule.make(50)
    st = S.ddim_cfgpp_steps(sch, 7.5, True)[11]
    t = int(st.t)
    at, an = sch.alphas_cumprod[t], sch.alphas_cumprod[t - sch.skip]
    zt = torch.randn(2, 4, 32, 32, generator=g).to(dev)
    eu = torch.randn(2, 4, 32, 32, generator=g).half().to(dev)
    ec = torch.randn(2, 4, 32, 32, generator=g).half().to(dev)
    npred = eu + 7.5 * (ec - eu)                              # latent_sdxl.py:449
    z0_ref = (zt - (1 - at).sqrt() * npred) / at.sqrt()       # :452
    zn_ref = an.sqrt() * z0_ref + (1 - an).sqrt() * npred     # :455 (guided eps, not eps_uc)
    z = zt.clone()
    z0 = nv.op_cfgpp_step(eu, ec, S.STEP_DDIM_CFG, st.coef, z)
    assert (z0 - z0_ref).abs().max() <= 1e-6 * z0_ref.abs().max()
    assert (z - zn_ref).abs().max() <= 1e-6 * zn_ref.abs().max()
    # fp16 state (the inversion / edit loops start from the VAE latent)
    inv = S.ddim_inversion_cfgpp_steps(sch, 7.5)[5]
    t = int(inv.t)
    at, ap = sch.alpha(t), sch.alpha(t - sch.skip)
    zh = torch.randn(1, 4, 32, 32, generator=g).half().to(dev)
    eu, ec = eu[:1].contiguous(), ec[:1].contiguous()
    npred = eu + 7.5 * (ec - eu)
    z0_ref = (zh - (1 - ap).sqrt() * npred) / ap.sqrt()       # latent_diffusion.py:179
    zn_ref = at.sqrt() * z0_ref + (1 - at).sqrt() * npred     # :180
    z = zh.clone()
    nv.op_cfgpp_step(eu, ec, S.STEP_DDIM_CFG, inv.coef, z)
    assert z.dtype == torch.float16 and ((z.float() - zn_ref.float()).abs() <= 2 * _ulp16(zn_ref)).all()


def test_plain_cfg_solvers_vs_oracle():
    from cfgpp_b200 import latent_diffusion as LD, latent_sdxl as LX
    from oracle import samplers as OSm, schedule as OS
    # SDXL: fused plain ddim (fp32 state) and Karras euler
    cfg, sd, net, ref = build_pair("tiny_sdxl", dev)
    net.close()
    nfe, lam, hw = 6, 2.0, 32
    z, uc, c, add = make_inputs(cfg, 1, hw, dev)
    tb = OS.make_tables(nfe)
    kw = dict(solver_config=SimpleNamespace(num_sampling=nfe), device=dev, unet_config=cfg, state_dict=sd)
    z0_ref = OSm.ddim_plain(ref, tb, z, uc, c, lam, add, sdxl_indexing=True)
    z0 = LX.get_solver("ddim", **kw).reverse_process(uc, c, lam, add, shape=(8 * hw, 8 * hw), zT=z)
    e = rel_l2(z0, z0_ref)
    print(f"sdxl ddim (plain CFG, lambda={lam}): rel-L2 final z0t {e:.3e}")
    assert z0.dtype == torch.float32 and e <= 3e-2
    sig = OSm.karras_sigmas(tb)
    x0 = OSm.kd_start_state(z, sig)
    d_ref, _ = OSm.kd_euler_cfgpp(ref, tb, x0.clone(), sig, uc, c, lam, add, plus=False)
    d = LX.get_solver("euler", **kw).reverse_process(uc, c, lam, add, shape=(8 * hw, 8 * hw), xT=x0.clone())
    e = rel_l2(d, d_ref)
    print(f"sdxl euler (plain CFG): rel-L2 last z0t {e:.3e}")
    assert e <= 3e-2
    # SD v1.5: plain inversion + edit on the fused mode (fp16 state), one plain k-diffusion sampler
    cfg, sd, net, ref = build_pair("tiny_sd15", dev)
    net.close()
    z, uc, c, _ = make_inputs(cfg, 1, hw, dev)
    g = torch.Generator().manual_seed(9)
    c_tgt = torch.randn(1, 77, cfg.cross_attention_dim, generator=g).half().to(dev)
    kw = dict(solver_config=SimpleNamespace(num_sampling=nfe), device=dev, unet_config=cfg, state_dict=sd)
    ed = LD.get_solver("ddim_edit", **kw)
    z0_src = (0.4 * z).half()
    zT_ref, z0t_ref = OSm.ddim_edit_plain(ref, tb, z0_src, uc, c, c_tgt, lam)
    zT = ed.inversion(z0_src, uc, c, lam)
    z0t = ed.reverse_process(uc, c_tgt, lam, zT)
    print(f"sd15 ddim_edit (plain CFG): rel-L2 zT {rel_l2(zT, zT_ref):.3e}, edited z0t {rel_l2(z0t, z0t_ref):.3e}")
    assert zT.dtype == torch.float16 and rel_l2(zT, zT_ref) <= 3e-2 and rel_l2(z0t, z0t_ref) <= 3e-2
    s2 = LD.get_solver("dpm++_2s_a", **kw)
    sig = s2.karras_sigmas()
    x0 = OSm.kd_start_state(z, sig)
    torch.manual_seed(77)
    _, x_ref = OSm.kd_dpmpp_2s_a_cfgpp(ref, tb, x0.clone(), sig, uc, c, lam, plus=False)
    torch.manual_seed(77)
    _, x = s2.reverse_process(uc, c, lam, x0.clone())
    print(f"sd15 dpm++_2s_a (plain CFG): rel-L2 final x {rel_l2(x, x_ref):.3e}")
    assert rel_l2(x, x_ref) <= 3e-2


def test_step_kernel_kdiffusion_family_matches_reference_arithmetic():
    """The selector bits of the fused VE-cast step (cfgpp_step_coef.second_order): Euler with the unconditional /
    guided extrapolation (latent_diffusion.py:699-719 / :326-330) and the Karras-sigma DPM++(2M) of SD v1.5 whose
    difference term uses the GUIDED estimate (:863), CFG++ and plain — against the torch ops of kdiffusion.py on the
    same eps, <= 2 fp16 ulp per step (teacher-forced)."""
    from cfgpp_b200 import _native as nv, kdiffusion as K, schedule as S
    g = torch.Generator().manual_seed(4)
    sigmas = K.get_sigmas_karras(6, 0.03, 14.6, rho=7.)
    timestep_fn = lambda s: torch.tensor(500)  # noqa: E731 — irrelevant for the update arithmetic
    t_fn = lambda s: s.log().neg()  # noqa: E731
    for cfgpp in (True, False):
        for second in (False, True):
            steps = S.kd_steps(sigmas, timestep_fn, 0.6, cfgpp, second_order=second, diff_guided=second)
            x = (torch.randn(1, 4, 32, 32, generator=g) * sigmas[0]).half().to(dev)
            xs, aux, old = x.clone(), torch.zeros_like(x), None
            for i in range(len(sigmas) - 1):
                eu = torch.randn(1, 4, 32, 32, generator=g).half().to(dev)
                ec = torch.randn(1, 4, 32, 32, generator=g).half().to(dev)
                npred = eu + 0.6 * (ec - eu)
                den, ud = x - npred * sigmas[i], x - eu * sigmas[i]
                ex = ud if cfgpp else den
                if not second or old is None or sigmas[i + 1] == 0:
                    xr = den + (x - ex) / sigmas[i].item() * sigmas[i + 1]
                else:
                    h = t_fn(sigmas[i + 1]) - t_fn(sigmas[i])
                    r = (t_fn(sigmas[i]) - t_fn(sigmas[i - 1])) / h
                    xr = den + (-torch.exp(-h) * ex - (-h).expm1() * (den - old) / (2 * r)) + torch.exp(-h) * x
                old = ex
                z0 = nv.op_cfgpp_step(eu, ec, S.STEP_DPMPP2M_CFGPP, steps[i].coef, xs, aux)
                tag = f"cfgpp={cfgpp} second={second} step {i}"
                assert ((xs.float() - xr.float()).abs() <= 2 * _ulp16(xr)).all(), tag
                assert ((z0.float() - den.float()).abs() <= 1 * _ulp16(den)).all(), tag
                assert ((aux.float() - ex.float()).abs() <= 1 * _ulp16(ex)).all(), tag
                x = xr
                xs.copy_(xr)
                aux.copy_(ex)


def test_step_kernel_ancestral_family_matches_reference_arithmetic():
    """Selector bits 8 / 16 / 32 of the fused VE-cast step: the ancestral Euler update and the two calls of a
    DPM-Solver++(2S) ancestral step (latent_diffusion.py:744-762, :782-825; plain forms :372-379, :408-437) against the
    torch ops of kdiffusion.py on the same eps and the same pre-drawn noise, <= 2 fp16 ulp per call (teacher-forced)."""
    from cfgpp_b200 import _native as nv, kdiffusion as K, schedule as S
    g = torch.Generator().manual_seed(9)
    sigmas = K.get_sigmas_karras(6, 0.03, 14.6, rho=7.)
    timestep_fn = lambda s: torch.tensor(500)  # noqa: E731 — irrelevant for the update arithmetic
    t_fn = lambda s: s.log().neg()      # noqa: E731
    sigma_fn = lambda t: t.neg().exp()  # noqa: E731
    shape = (2, 4, 16, 16)
    draw = lambda: torch.randn(shape, generator=g).half().to(dev)  # noqa: E731
    for cfgpp in (True, False):
        for two_s in (False, True):
            steps, slots = S.kd_ancestral_steps(sigmas, timestep_fn, 0.6, cfgpp, two_s)
            assert slots == len(sigmas) - 2                       # every step but the last adds noise
            assert len(steps) == (2 * (len(sigmas) - 2) + 1 if two_s else len(sigmas) - 1)
            noise = torch.stack([draw() for _ in range(slots)])
            x = (torch.randn(shape, generator=g) * sigmas[0]).half().to(dev)
            xs, aux, k = x.clone(), torch.zeros_like(x), 0

            def call(xin, sigma, coef, tag):
                """One UNet call's worth of update on the native kernel and in torch; returns (den, ud)."""
                eu, ec = draw(), draw()
                npred = eu + 0.6 * (ec - eu)
                z0 = nv.op_cfgpp_step(eu, ec, S.STEP_DPMPP2M_CFGPP, coef, xs, aux, noise=noise)
                den, ud = xin - npred * sigma, xin - eu * sigma
                assert ((z0.float() - den.float()).abs() <= 1 * _ulp16(den)).all(), tag
                return den, ud

            for i in range(len(sigmas) - 1):
                tag = f"cfgpp={cfgpp} two_s={two_s} step {i}"
                sigma_down, sigma_up = K.get_ancestral_step(sigmas[i], sigmas[i + 1])
                if not two_s or sigma_down == 0:
                    den, ud = call(x, sigmas[i], steps[k].coef, tag)
                    k += 1
                    ex = ud if cfgpp else den
                    xr = den + (x - ex) / sigmas[i].item() * sigma_down
                else:
                    den, ud = call(x, sigmas[i], steps[k].coef, tag + " mid")
                    ex = ud if cfgpp else den
                    t, t_next = t_fn(sigmas[i]), t_fn(sigma_down)
                    h = t_next - t
                    s = t + 0.5 * h
                    x_2 = (sigma_fn(s) / sigma_fn(t)) * x - (-h * 0.5).expm1() * ex
                    assert ((xs.float() - x_2.float()).abs() <= 2 * _ulp16(x_2)).all(), tag + " x_2"
                    assert torch.equal(aux, x), tag + " parked x"
                    xs.copy_(x_2)
                    den2, ud2 = call(x_2, sigma_fn(s), steps[k + 1].coef, tag + " final")
                    k += 2
                    if cfgpp:
                        xr = den2 - torch.exp(-h) * ud2 + (sigma_fn(t_next) / sigma_fn(t)) * x
                    else:
                        xr = (sigma_fn(t_next) / sigma_fn(t)) * x - (-h).expm1() * den2
                if sigmas[i + 1] > 0:
                    xr = xr + noise[i] * sigma_up
                assert ((xs.float() - xr.float()).abs() <= 2 * _ulp16(xr)).all(), tag
                x = xr
                xs.copy_(xr)
            assert k == len(steps)


@pytest.mark.parametrize("method", ["euler_a_cfg++", "dpm++_2s_a_cfg++", "euler_a", "dpm++_2s_a"])
def test_ancestral_fused_trajectory_equals_callback_path(method):
    """The ancestral samplers run as fused CUDA-graph trajectories (noise drawn up front); with a callback installed
    they take the op-by-op loop over the `predict_noise` seam, drawing noise step by step. Same seed => the same noise
    values, and the same state up to fp16 rounding-order noise of the last ulp."""
    from cfgpp_b200 import latent_diffusion as LD
    cfg, sd, net, ref = build_pair("tiny_sd15", dev)
    net.close()
    del ref
    nfe, lam, hw = 6, 0.6, 32
    z, uc, c, _ = make_inputs(cfg, 1, hw, dev)
    solver = LD.get_solver(method, solver_config=SimpleNamespace(num_sampling=nfe), device=dev, unet_config=cfg,
                           state_dict=sd)
    x0 = (z * (solver.karras_sigmas()[0] ** 2 + 1) ** 0.5).half()
    seen = []
    torch.manual_seed(77)
    d_cb, x_cb = solver.reverse_process(uc, c, lam, x0.clone(), callback_fn=lambda i, t, kw: (seen.append(i), kw)[1])
    torch.manual_seed(77)
    d_f, x_f = solver.reverse_process(uc, c, lam, x0.clone())
    assert seen == list(range(nfe))
    e_x, e_d = rel_l2(x_f, x_cb), rel_l2(d_f, d_cb)
    print(f"{method}: fused vs callback path — final x {e_x:.3e}, last denoised {e_d:.3e}")
    assert e_x <= 2e-3 and e_d <= 2e-3
    torch.manual_seed(78)
    _, x_other = solver.reverse_process(uc, c, lam, x0.clone())
    assert rel_l2(x_other, x_f) > 1e-2        # a different seed really changes the injected noise

"""Generates tests/golden/r01_golden.pt — run from the repo root:  python tests/golden/make_golden.py

The reference cannot be imported in this container (diffusers / munch are absent, no checkpoints: SURVEY.md section
8c), so these vectors are NOT outputs of the reference: they are outputs of the oracle restatement (oracle/), frozen
so that (a) the oracle itself cannot drift unnoticed, (b) the CUDA path is compared with committed numbers and not only
with an oracle evaluated at test time, (c) the Appendix-B schedule constants are pinned as numbers.
Contents: schedule tables (sampled), tiny-UNet fp32 forwards on seeded CPU weights, every oracle sampler on a closed-form
stand-in UNet."""
import dataclasses
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from cfgpp_b200 import config as C, weights as Wt          # noqa: E402
from oracle import samplers as OSm, schedule as OS, unet as O  # noqa: E402


class ClosedFormUNet:
    """eps = tanh(z * (0.5 + t/1000)) * 0.8 + 0.3 * mean(ctx) (+ 0.05 * mean(pooled)) — same as tests/test_kdiffusion_cpu."""
    def __call__(self, z, t, encoder_hidden_states=None, added_cond_kwargs=None):
        t = t.reshape(-1, 1, 1, 1).to(z.dtype)
        ctx = encoder_hidden_states.float().mean(dim=(1, 2)).reshape(-1, 1, 1, 1).to(z.dtype)
        eps = torch.tanh(z * (0.5 + t / 1000)) * 0.8 + 0.3 * ctx
        if added_cond_kwargs is not None:
            eps = eps + 0.05 * added_cond_kwargs["text_embeds"].float().mean().to(z.dtype)
        return {"sample": eps}


def unet_case(name, seed=1234, hw=16):
    cfg = C.CONFIGS[name]()
    sd = Wt.synthetic_state_dict(cfg, seed=seed, device="cpu")           # CPU generator: box independent
    ocfg = O.UNetConfig(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(O.UNetConfig)})
    m = O.build_unet(ocfg, sd, dtype=torch.float32)
    g = torch.Generator().manual_seed(99)
    z = torch.randn(1, 4, hw, hw, generator=g)
    ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=g).half()
    add = None
    if cfg.addition_embed_type:
        add = {"text_embeds": torch.randn(2, cfg.pooled_dim, generator=g).half(),
               "time_ids": torch.tensor([[hw * 8., hw * 8, 0, 0, hw * 8, hw * 8]] * 2).half()}
    t = 481
    with torch.no_grad():
        out = m(torch.cat([z, z]), torch.tensor([t, t]), ctx.float(),
                None if add is None else {k: v.float() for k, v in add.items()})["sample"]
    return {"seed": seed, "hw": hw, "t": t, "z": z, "ctx": ctx, "add": add, "eps_uc": out[:1].clone(), "eps_c": out[1:].clone()}


def sampler_cases():
    g = torch.Generator().manual_seed(3)
    nfe, hw = 6, 8
    tb = OS.make_tables(nfe)
    noise = torch.randn(1, 4, hw, hw, generator=g)
    uc = torch.randn(1, 77, 16, generator=g).half()
    c = torch.randn(1, 77, 16, generator=g).half()
    c2 = torch.randn(1, 77, 16, generator=g).half()
    add = {"text_embeds": torch.ones(2, 4).half(), "time_ids": torch.zeros(2, 6).half()}
    u = ClosedFormUNet()
    ks = OSm.karras_sigmas(tb)
    x0 = OSm.kd_start_state(noise, ks)
    xs = OSm.kd_start_state(noise, OSm.sdxl_euler_sigmas(tb))
    z0 = (0.3 * noise).half()
    out = {"inputs": {"nfe": nfe, "noise": noise, "uc": uc, "c": c, "c_tgt": c2, "add": add, "lam": 0.6}}

    def seeded(fn):
        torch.manual_seed(17)
        return fn()
    out["sd15_ddim_cfgpp"] = OSm.sd15_ddim_cfgpp(u, tb, noise, uc, c, 0.6)
    out["sd15_inversion_cfgpp"] = OSm.sd15_inversion_cfgpp(u, tb, z0, uc, c, 0.6)
    out["sdxl_ddim_cfgpp"] = OSm.sdxl_ddim_cfgpp(u, tb, noise, uc, c, 0.6, add)
    out["sdxl_dpmpp_2m_cfgpp"] = OSm.sdxl_dpmpp_2m_cfgpp(u, tb, noise, uc, c, 0.6, add)
    out["ddim_edit_cfgpp"] = OSm.ddim_edit_cfgpp(u, tb, z0, uc, c, c2, 0.6)[1]
    out["ddim_plain_sd15"] = OSm.ddim_plain(u, tb, noise, uc, c, 0.6)
    out["ddim_plain_sdxl"] = OSm.ddim_plain(u, tb, noise, uc, c, 0.6, add, sdxl_indexing=True)
    out["ddim_edit_plain"] = OSm.ddim_edit_plain(u, tb, z0, uc, c, c2, 0.6)[1]
    for plus in (True, False):
        k = "cfgpp" if plus else "cfg"
        out[f"euler_{k}"] = OSm.kd_euler_cfgpp(u, tb, x0.clone(), ks, uc, c, 0.6, plus=plus)[1]
        out[f"euler_a_{k}"] = seeded(lambda: OSm.kd_euler_cfgpp(u, tb, x0.clone(), ks, uc, c, 0.6, ancestral=True, plus=plus)[1])
        out[f"dpmpp_2s_a_{k}"] = seeded(lambda: OSm.kd_dpmpp_2s_a_cfgpp(u, tb, x0.clone(), ks, uc, c, 0.6, plus=plus)[1])
        out[f"dpmpp_2m_sd15_{k}"] = OSm.kd_dpmpp_2m_cfgpp_sd15(u, tb, x0.clone(), ks, uc, c, 0.6, plus=plus)[1]
    out["sdxl_euler_cfgpp"] = OSm.kd_euler_cfgpp(u, tb, xs.clone(), OSm.sdxl_euler_sigmas(tb), uc, c, 0.6, add)[0]
    return out


def schedule_cases():
    t50, t10, l4 = OS.make_tables(50), OS.make_tables(10), OS.make_tables(4, "lightning")
    idx = torch.tensor([0, 1, 2, 250, 500, 750, 999, 1000])
    return {"alphas_cumprod_at": (idx, t50.alphas_cumprod[idx].double()), "final_alpha_cumprod": t50.final_alpha_cumprod.double(),
            "timesteps_50": t50.timesteps.clone(), "timesteps_10": t10.timesteps.clone(), "timesteps_lightning_4": l4.timesteps.clone(),
            "skip": (t50.skip, t10.skip, l4.skip), "sigma_min_max": (t50.sigmas.min().double(), t50.sigmas.max().double()),
            "karras_6": OSm.karras_sigmas(OS.make_tables(6)).double()}


def vae_case(seed=4242, hw=16):
    """fp32-oracle decode of the tiny AutoencoderKL decoder on seeded CPU weights (tests/golden/r02_vae_golden.pt)."""
    from cfgpp_b200 import vae as V
    from oracle import vae as OV
    cfg = V.tiny_vae_config()
    sd = V.synthetic_vae_state_dict(cfg, seed=seed, device="cpu")
    m = OV.build_vae_decoder(OV.VAEConfig(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(OV.VAEConfig)}), sd,
                             dtype=torch.float32)
    g = torch.Generator().manual_seed(77)
    zt = torch.randn(1, 4, hw, hw, generator=g) * cfg.scaling_factor * 6.0
    return {"seed": seed, "hw": hw, "zt": zt, "image": OV.decode(m, zt).half()}


def clip_ids(cfg, batch=3, seed=5):
    """Well-formed prompts: <bos> body <eos> pad..., one of them filling all 77 positions."""
    g = torch.Generator().manual_seed(seed)
    T, bos, eos = cfg.max_position_embeddings, cfg.vocab_size - 2, cfg.vocab_size - 1
    ids = torch.full((batch, T), cfg.pad_token_id, dtype=torch.int64)
    for b, n in enumerate([7, 30, T - 2][:batch]):
        ids[b, 0] = bos
        ids[b, 1:1 + n] = torch.randint(1, cfg.vocab_size - 3, (n,), generator=g)
        ids[b, 1 + n] = eos
    return ids


def clip_case(seed=777):
    """Outputs of the REAL transformers CLIPTextModel / CLIPTextModelWithProjection (the classes the reference's
    pipeline instantiates) on seeded tiny towers (tests/golden/r02_clip_golden.pt). The weights are regenerated from
    the seed by text_encoder.synthetic_clip_state_dict, so only ids and outputs are stored."""
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection
    from cfgpp_b200 import text_encoder as TE
    out = {"transformers": transformers.__version__, "seed": seed, "cases": {}}
    for name, proj, act in (("plain_quick_gelu", 0, "quick_gelu"), ("proj_gelu", 64, "gelu")):
        cfg = TE.tiny_clip_config(proj, act)
        sd = TE.synthetic_clip_state_dict(cfg, seed=seed, device="cpu")
        hc = CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                            num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                            max_position_embeddings=cfg.max_position_embeddings, hidden_act=cfg.hidden_act,
                            projection_dim=proj or 8, layer_norm_eps=cfg.layer_norm_eps, eos_token_id=cfg.eos_token_id,
                            bos_token_id=cfg.vocab_size - 2, pad_token_id=cfg.pad_token_id)
        m = (CLIPTextModelWithProjection if proj else CLIPTextModel)(hc).eval()
        missing, unexpected = m.load_state_dict({k: v.float() for k, v in sd.items()}, strict=False)
        assert not unexpected and all(k.endswith("position_ids") for k in missing), (missing, unexpected)
        ids = clip_ids(cfg)
        with torch.no_grad():
            o = m(ids, output_hidden_states=True)
        out["cases"][name] = {"proj": proj, "act": act, "ids": ids.to(torch.int32),
                              "hidden_states": [h.half() for h in o.hidden_states],
                              "last_hidden_state": o.last_hidden_state.half(),
                              "pooled": (o.text_embeds if proj else o.pooler_output).half()}
    return out


if __name__ == "__main__":
    torch.set_num_threads(1)  # summation order of the CPU kernels is part of the pin
    if len(sys.argv) > 1 and sys.argv[1] == "clip":
        out = Path(__file__).with_name("r02_clip_golden.pt")
        torch.save(clip_case(), out)
        print(out, out.stat().st_size, "bytes")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "vae":
        out = Path(__file__).with_name("r02_vae_golden.pt")
        torch.save({"vae": vae_case(), "torch": torch.__version__}, out)
        print(out, out.stat().st_size, "bytes")
        sys.exit(0)
    blob = {"schedule": schedule_cases(), "unet": {n: unet_case(n) for n in ("tiny_sdxl", "tiny_sd15")},
            "samplers": sampler_cases(), "torch": torch.__version__}
    out = Path(__file__).with_name("r01_golden.pt")
    torch.save(blob, out)
    print(out, out.stat().st_size, "bytes")

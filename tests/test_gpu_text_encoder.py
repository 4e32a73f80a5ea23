"""CLIP text towers through the C ABI against (a) the committed fixture of outputs produced by the real
transformers classes, (b) the oracle restatement (oracle/text_encoder.py, itself pinned to transformers on CPU) on the
same seeded weights, and (c) transformers itself on the GPU when it imports: `hidden_states[-2]` / `[-(clip_skip+2)]`,
`last_hidden_state`, pooled `text_embeds` / `pooler_output` (latent_sdxl.py:77-93, latent_diffusion.py:93-115).

Stated tolerance (same rule as the UNet forward): rel-L2 <= 5e-3 against the fp16 model — what the reference runs
(`torch_dtype=float16`) — AND the error against the fp32 model must not exceed 1.5x the fp16 model's own."""
import dataclasses
from pathlib import Path

import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")
GOLDEN = Path(__file__).parent / "golden" / "r02_clip_golden.pt"


def _ocfg(cfg):
    from oracle import text_encoder as OT
    names = {f.name for f in dataclasses.fields(OT.CLIPTextCfg)}
    return OT.CLIPTextCfg(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(cfg) if f.name in names})


def _ids(cfg, batch, seed=3):
    g = torch.Generator().manual_seed(seed)
    T, bos, eos = cfg.max_position_embeddings, cfg.vocab_size - 2, cfg.vocab_size - 1
    ids = torch.full((batch, T), cfg.pad_token_id, dtype=torch.int32)
    for b in range(batch):
        n = [5, T - 2, 23, 0, 40, 11, 60, 1][b % 8]
        ids[b, 0] = bos
        ids[b, 1:1 + n] = torch.randint(1, cfg.vocab_size - 3, (n,), generator=g, dtype=torch.int32)
        ids[b, 1 + n] = eos
    return ids


def test_native_text_encoder_against_transformers_fixture():
    from cfgpp_b200 import text_encoder as TE
    blob = torch.load(GOLDEN)
    for name, case in blob["cases"].items():
        cfg = TE.tiny_clip_config(case["proj"], case["act"])
        sd = TE.synthetic_clip_state_dict(cfg, seed=blob["seed"], device="cpu")
        enc = TE.NativeCLIPTextEncoder(cfg, sd, dev)
        L = cfg.num_hidden_layers
        for skip in range(L + 1):
            hidden, last, pooled = enc.encode(case["ids"], skip=skip)
            e = rel_l2(hidden.float().cpu(), case["hidden_states"][L - skip].float())
            assert e <= 3e-3, f"{name} hidden_states[{L - skip}] {e:.3e}"
        e_last = rel_l2(last.float().cpu(), case["last_hidden_state"].float())
        e_pool = rel_l2(pooled.float().cpu(), case["pooled"].float())
        print(f"{name}: vs transformers fixture — last {e_last:.3e}, pooled {e_pool:.3e}")
        assert e_last <= 3e-3 and e_pool <= 3e-3
        enc.close()


def _case(cfg, batch, seed=21, skips=(1,)):
    from cfgpp_b200 import text_encoder as TE
    from oracle import text_encoder as OT
    sd = TE.synthetic_clip_state_dict(cfg, seed=seed, device=dev)
    ids = _ids(cfg, batch)
    enc = TE.NativeCLIPTextEncoder(cfg, sd, dev)
    m16 = OT.build_clip_text(_ocfg(cfg), sd, dtype=torch.float16, device=dev)
    hs16, last16, pooled16, emb16 = m16(ids.long().to(dev))
    del m16
    m32 = OT.build_clip_text(_ocfg(cfg), sd, dtype=torch.float32, device=dev)
    hs32, last32, pooled32, emb32 = m32(ids.long().to(dev))
    del m32
    L = cfg.num_hidden_layers
    out = {}
    for skip in skips:
        hidden, last, pooled = enc.encode(ids, skip=skip)
        again = enc.encode(ids, skip=skip)
        assert all(torch.equal(a, b) for a, b in zip((hidden, last, pooled), again))  # deterministic, plan reuse
        p16, p32 = (emb16, emb32) if cfg.projection_dim else (pooled16, pooled32)
        for what, got, r16, r32 in (("hidden", hidden, hs16[L - skip], hs32[L - skip]), ("last", last, last16, last32),
                                    ("pooled", pooled, p16, p32)):
            e16, e32, b32 = rel_l2(got, r16), rel_l2(got, r32), rel_l2(r16, r32)
            print(f"{cfg.name} B={batch} skip={skip} {what}: vs fp16 model {e16:.3e}, vs fp32 model {e32:.3e} "
                  f"(fp16 model itself {b32:.3e})")
            assert torch.isfinite(got).all()
            assert e16 <= 5e-3 and e32 <= 1.5 * b32 + 1e-4, (cfg.name, what, e16, e32, b32)
        out[skip] = (hidden, last, pooled)
    st = enc.stats
    assert st["flops"] > 0 and st["workspace_bytes"] > 0
    enc.close()
    return out


@pytest.mark.parametrize("proj,act,batch", [(0, "quick_gelu", 1), (64, "gelu", 3), (0, "gelu", 8), (128, "quick_gelu", 16)])
def test_tiny_towers_vs_oracle(proj, act, batch):
    from cfgpp_b200 import text_encoder as TE
    _case(TE.tiny_clip_config(proj, act), batch, skips=(0, 1, 2, 3))


def test_clip_l_full_size_vs_oracle():
    """openai/clip-vit-large-patch14 geometry (SD v1.5 text_encoder, SDXL text_encoder): 12 layers, 768 wide."""
    from cfgpp_b200 import text_encoder as TE
    _case(TE.clip_l_config(), batch=2, skips=(0, 1))


def test_clip_bigg_full_size_vs_oracle():
    """OpenCLIP ViT-bigG geometry (SDXL text_encoder_2): 32 layers, 1280 wide, gelu, 1280-d text projection."""
    from cfgpp_b200 import text_encoder as TE
    _case(TE.clip_bigg_config(), batch=2, skips=(1, 2))


def test_clip_l_against_transformers_on_gpu():
    """The library class itself (fp16, as the reference loads it) on the same weights and ids."""
    tr = pytest.importorskip("transformers")
    from cfgpp_b200 import text_encoder as TE
    cfg = TE.clip_l_config()
    sd = TE.synthetic_clip_state_dict(cfg, seed=5, device=dev)
    hc = tr.CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                           num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                           max_position_embeddings=77, hidden_act=cfg.hidden_act, projection_dim=768, eos_token_id=2,
                           bos_token_id=49406, pad_token_id=1)
    ref = tr.CLIPTextModel(hc).eval()
    ref.load_state_dict({k: v.float().cpu() for k, v in sd.items()}, strict=False)
    ids = _ids(cfg, 2)
    ref16 = ref.to(device=dev, dtype=torch.float16)
    with torch.no_grad():
        o = ref16(ids.long().to(dev), output_hidden_states=True)
    enc = TE.NativeCLIPTextEncoder(cfg, sd, dev)
    hidden, last, pooled = enc.encode(ids, skip=1)
    for what, got, r in (("hidden_states[-2]", hidden, o.hidden_states[-2]), ("last_hidden_state", last, o.last_hidden_state),
                         ("pooler_output", pooled, o.pooler_output)):
        e = rel_l2(got, r)
        print(f"clip_l vs transformers fp16 {what}: {e:.3e}")
        assert e <= 5e-3
    enc.close()


def test_conditioners_feed_the_solvers():
    """`get_text_embed` of both solver families runs the native towers (default path) with the reference's shapes."""
    from types import SimpleNamespace
    from cfgpp_b200 import latent_diffusion as LD
    from cfgpp_b200 import latent_sdxl as LX
    from cfgpp_b200 import text_encoder as TE
    from cfgpp_b200.config import tiny_sd15_config, tiny_sdxl_config
    s = LX.get_solver("ddim_cfg++", solver_config=SimpleNamespace(num_sampling=4), device="cuda:0",
                      unet_config=tiny_sdxl_config(), model_key="synthetic:7")
    assert isinstance(s.text_enc_1, TE.ClipConditioner) and isinstance(s.text_enc_2, TE.ClipConditioner)
    uc, c, pn, pc = s.get_text_embed("", "a photo of a cat", "", "a photo of a cat")
    assert uc.shape == c.shape == (1, 77, s.cfg.cross_attention_dim) and pn.shape == pc.shape == (1, s.cfg.pooled_dim)
    assert uc.dtype == torch.float16 and not torch.equal(uc, c) and torch.isfinite(c).all()
    uc2, c2, _, pc2 = s.get_text_embed("", "a photo of a dog", "", "a photo of a dog")
    assert torch.equal(uc, uc2) and not torch.equal(c, c2) and not torch.equal(pc, pc2)
    # causal attention: the shared prefix "<bos> a photo of a" conditions identically
    assert torch.equal(c[:, :5], c2[:, :5]) and not torch.equal(c[:, 5], c2[:, 5])
    _, c3, _, _ = s.get_text_embed("", "a photo of a cat", "", "a photo of a cat", clip_skip=1)
    assert not torch.equal(c, c3)
    # many prompts per call: row i of the batched encode == the single-prompt result
    prompts = [f"a photo of object number {i} " + "very " * (i % 7) + "nice" for i in range(37)] + [""]
    hb, pb = s.text_enc_2.encode_batch(prompts)
    assert hb.shape == (38, 77, s.text_enc_2.encoder.cfg.hidden_size) and pb.shape == (38, s.cfg.pooled_dim)
    for i in (0, 15, 16, 36, 37):
        h1, p1 = s.text_enc_2(prompts[i])
        assert rel_l2(hb[i:i + 1], h1) <= 1e-3 and rel_l2(pb[i:i + 1], p1) <= 1e-3
    d = LD.get_solver("ddim_cfg++", solver_config=SimpleNamespace(num_sampling=4), device="cuda:0",
                      unet_config=tiny_sd15_config(), model_key="synthetic:7")
    assert isinstance(d.text_encoder, TE.ClipConditioner)
    u, t = d.get_text_embed("", "a photo of a cat")
    assert u.shape == t.shape == (1, 77, d.cfg.cross_attention_dim) and torch.isfinite(t).all() and not torch.equal(u, t)
    LX.release_engines()
    TE.release_text_encoders()

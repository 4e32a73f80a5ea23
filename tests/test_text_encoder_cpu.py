"""CPU checks of the text-conditioning row: the oracle restatement of the CLIP text towers is pinned
against the real `transformers` classes the reference's pipeline instantiates (live when transformers imports, and
through the committed fixture tests/golden/r02_clip_golden.pt either way), the BPE tokenizer against transformers'
`CLIPTokenizer`, and the host-side tables of cfgpp_b200/text_encoder.py against both."""
import collections
import dataclasses
import json
from pathlib import Path

import pytest
import torch

from cfgpp_b200 import text_encoder as TE
from cfgpp_b200 import tokenizer as TK
from oracle import text_encoder as OT

GOLDEN = Path(__file__).parent / "golden" / "r02_clip_golden.pt"


def _ocfg(cfg: TE.CLIPTextConfig) -> OT.CLIPTextCfg:
    names = {f.name for f in dataclasses.fields(OT.CLIPTextCfg)}
    return OT.CLIPTextCfg(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(cfg) if f.name in names})


def test_param_counts_match_the_published_models():
    # openai/clip-vit-large-patch14 text tower: 123,060,480; OpenCLIP bigG text tower + projection: 694,659,840
    assert TE.num_clip_params(TE.clip_l_config()) == 123_060_480
    assert TE.num_clip_params(TE.clip_bigg_config()) == 694_659_840
    for cfg in (TE.clip_l_config(), TE.clip_bigg_config()):
        with torch.device("meta"):
            m = OT.CLIPText(_ocfg(cfg))
        assert OT.count_params(m) == TE.num_clip_params(cfg)
        assert set(m.state_dict().keys()) == {k for k, _, _ in TE.clip_param_specs(cfg)}


def test_oracle_matches_transformers_fixture():
    blob = torch.load(GOLDEN)
    for name, case in blob["cases"].items():
        cfg = TE.tiny_clip_config(case["proj"], case["act"])
        sd = TE.synthetic_clip_state_dict(cfg, seed=blob["seed"], device="cpu")
        m = OT.build_clip_text(_ocfg(cfg), sd, dtype=torch.float32)
        hs, last, pooled, emb = m(case["ids"].long())
        assert len(hs) == len(case["hidden_states"]) == cfg.num_hidden_layers + 1
        for a, b in zip(hs, case["hidden_states"]):
            assert (a - b.float()).abs().max() <= 2e-3 * max(1.0, b.float().abs().max()), name  # fixture stored as fp16
        assert (last - case["last_hidden_state"].float()).abs().max() <= 4e-3
        got = emb if case["proj"] else pooled
        assert (got - case["pooled"].float()).abs().max() <= 4e-3


@pytest.mark.parametrize("proj,act", [(0, "quick_gelu"), (64, "gelu")])
def test_oracle_matches_transformers_live(proj, act):
    tr = pytest.importorskip("transformers")
    cfg = TE.tiny_clip_config(proj, act)
    hc = tr.CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                           num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                           max_position_embeddings=77, hidden_act=act, projection_dim=proj or 8, eos_token_id=2,
                           bos_token_id=cfg.vocab_size - 2, pad_token_id=cfg.pad_token_id)
    ref = (tr.CLIPTextModelWithProjection if proj else tr.CLIPTextModel)(hc).eval()
    sd = {k: v.float() for k, v in TE.synthetic_clip_state_dict(cfg, seed=11, device="cpu").items()}
    ref.load_state_dict(sd, strict=False)
    m = OT.build_clip_text(_ocfg(cfg), sd, dtype=torch.float32)
    ids = torch.randint(1, cfg.vocab_size - 3, (2, 77))
    ids[:, 0] = cfg.vocab_size - 2
    ids[0, 12] = cfg.vocab_size - 1
    ids[0, 13:] = cfg.pad_token_id
    ids[1, 76] = cfg.vocab_size - 1
    with torch.no_grad():
        o = ref(ids, output_hidden_states=True)
    hs, last, pooled, emb = m(ids)
    for a, b in zip(hs, o.hidden_states):
        assert torch.allclose(a, b, atol=2e-5, rtol=1e-5)
    assert torch.allclose(last, o.last_hidden_state, atol=2e-5, rtol=1e-5)
    assert torch.allclose(emb if proj else pooled, o.text_embeds if proj else o.pooler_output, atol=2e-5, rtol=1e-5)
    # the reference's selection of outputs (latent_sdxl.py:77-93)
    h, out0 = OT.text_embed(m, ids, clip_skip=None)
    assert torch.equal(h, hs[-2]) and torch.equal(out0, emb if proj else last)
    h2, _ = OT.text_embed(m, ids, clip_skip=1)
    assert torch.equal(h2, hs[-3])


def _toy_vocab(tmp_path):
    corpus = ("a photo of an astronaut riding a horse on mars. a cat's whiskers, the dog's bone! don't stop 123 believing "
              "professional photograph of a sunset over the mountains, highly detailed, 8k resolution, trending on "
              "artstation; café naïve résumé — über cool ... a a a the the of of")
    b2u = TK.bytes_to_unicode()
    words = collections.Counter()
    for tok in TK._WORD_PATTERN.findall(TK.clean_text(corpus)):
        sym = [b2u[b] for b in tok.encode()]
        sym[-1] += "</w>"
        words[tuple(sym)] += 1
    merges = []
    for _ in range(120):
        pairs = collections.Counter()
        for w, c in words.items():
            for p in zip(w[:-1], w[1:]):
                pairs[p] += c
        if not pairs:
            break
        best = max(sorted(pairs), key=lambda p: pairs[p])
        merges.append(best)
        nw = collections.Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i < len(w) - 1 and (w[i], w[i + 1]) == best:
                    out.append(w[i] + w[i + 1])
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            nw[tuple(out)] += c
        words = nw
    vocab = list(b2u.values()) + [v + "</w>" for v in b2u.values()] + ["".join(m) for m in merges] + [TK.BOS_TOKEN, TK.EOS_TOKEN]
    vf, mf = tmp_path / "vocab.json", tmp_path / "merges.txt"
    vf.write_text(json.dumps({t: i for i, t in enumerate(vocab)}), encoding="utf-8")
    mf.write_text("#version: 0.2\n" + "\n".join(" ".join(m) for m in merges) + "\n", encoding="utf-8")
    return str(vf), str(mf)


PROMPTS = ["a photo of an astronaut riding a horse on mars", "", "A Cat's   whiskers!!  don't STOP",
           "café naïve 12345 über-cool ... the end", "x " * 100, "red pandas & bamboo; 8k, trending"]


def test_bpe_tokenizer_framing_and_merges(tmp_path):
    vf, mf = _toy_vocab(tmp_path)
    tok = TK.ClipBPETokenizer(vf, mf)
    rows = tok(PROMPTS)
    assert all(len(r) == 77 and r[0] == tok.bos_token_id for r in rows)
    assert rows[1] == [tok.bos_token_id, tok.eos_token_id] + [tok.pad_token_id] * 75       # the null prompt
    assert rows[4][76] == tok.eos_token_id and rows[4].count(tok.eos_token_id) == 1        # truncated to 75 + specials
    assert tok.tokenize("the the") == tok.tokenize("THE   the")                             # lower-casing, whitespace
    assert tok.pooled_index(rows)[1] == 1 and tok.pooled_index(rows)[4] == 76


def test_bpe_tokenizer_matches_transformers(tmp_path):
    tr = pytest.importorskip("transformers")
    vf, mf = _toy_vocab(tmp_path)
    mine, ref = TK.ClipBPETokenizer(vf, mf), tr.CLIPTokenizer(vf, mf)
    assert (mine.bos_token_id, mine.eos_token_id, mine.pad_token_id) == (ref.bos_token_id, ref.eos_token_id, ref.pad_token_id)
    for p in PROMPTS:
        assert mine([p])[0] == ref(p, padding="max_length", max_length=77, truncation=True).input_ids, p


def test_hash_tokenizer_is_a_well_formed_stand_in():
    for pad in (49407, 0):
        tok = TK.HashTokenizer(49408, pad)
        rows = tok(["a photo of a cat", "", "word " * 200])
        assert all(len(r) == 77 and r[0] == 49406 for r in rows)
        assert rows[1][1] == 49407 and rows[1][2:] == [pad] * 75
        assert rows[2][76] == 49407 and all(1 <= t <= 49405 for t in rows[2][1:76])
        assert tok(["a photo of a cat"]) == tok(["A  photo of a CAT"])
        assert tok.pooled_index(rows) == [rows[0].index(49407), 1, 76]


def test_native_encoder_refuses_cpu():
    from cfgpp_b200._native import NativeError
    with pytest.raises(NativeError):
        TE.NativeCLIPTextEncoder(TE.tiny_clip_config(), {}, device="cpu")

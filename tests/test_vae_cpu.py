"""CPU side of the AutoencoderKL decoder row: the oracle's structural checksum, the key surface shared by
oracle and product, the committed golden image, and the no-CUDA failure mode of the product path."""
import dataclasses
from pathlib import Path

import pytest
import torch


def _ocfg(cfg):
    from oracle import vae as OV
    return OV.VAEConfig(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(OV.VAEConfig)})


def test_decoder_param_count_reproduces_the_published_autoencoderkl_size():
    """AutoencoderKL of SD v1.5 / SDXL has 83,653,863 parameters; encoder 34,163,592 + quant_conv 72 + what the oracle
    restates (post_quant_conv + decoder) = 49,490,199 — only the published structure reproduces the total."""
    from cfgpp_b200 import vae as V
    from oracle import vae as OV
    m = OV.build_vae_decoder(OV.sdxl_vae_config())
    assert OV.count_params(m) == 49_490_199 == V.num_vae_decoder_params(V.sdxl_vae_config())
    assert 34_163_592 + 72 + OV.count_params(m) == 83_653_863
    # product spec and oracle modules expose the same diffusers keys and shapes
    spec = {k: tuple(s) for k, s, _ in V.vae_decoder_param_specs(V.sdxl_vae_config())}
    assert spec == {k: tuple(v.shape) for k, v in m.state_dict().items()}


def test_oracle_vae_against_committed_golden():
    gold = torch.load(Path(__file__).parent / "golden" / "r02_vae_golden.pt", weights_only=False)["vae"]
    from cfgpp_b200 import vae as V
    from oracle import vae as OV
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        cfg = V.tiny_vae_config()
        m = OV.build_vae_decoder(_ocfg(cfg), V.synthetic_vae_state_dict(cfg, seed=gold["seed"], device="cpu"), dtype=torch.float32)
        img = OV.decode(m, gold["zt"])
    finally:
        torch.set_num_threads(n)
    assert img.shape == (1, 3, 8 * gold["hw"], 8 * gold["hw"])
    assert (img - gold["image"].float()).abs().max() <= 2e-3 * gold["image"].float().abs().max()


def test_oracle_decode_divides_by_the_scaling_factor():
    from oracle import vae as OV
    m = OV.build_vae_decoder(OV.tiny_vae_config())
    z = torch.randn(1, 4, 8, 8)
    assert torch.allclose(OV.decode(m, z), m(z / 0.13025), atol=1e-6)
    assert OV.sd15_vae_config().scaling_factor == 0.18215  # latent_diffusion.py:127


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_native_vae_fails_loudly_without_cuda():
    from cfgpp_b200 import _native as nv, vae as V
    cfg = V.tiny_vae_config()
    with pytest.raises(nv.NativeError, match="CUDA"):
        V.NativeVAEDecoder(cfg, V.synthetic_vae_state_dict(cfg), "cpu")


def test_encoder_param_count_and_posterior_semantics():
    """Encoder 34,163,592 + quant_conv 72 parameters (with the decoder's 49,490,199: the published 83,653,863); the
    posterior clamps the log-variance to [-30, 20] and `encode` is (mean + std * noise) * scaling_factor."""
    import dataclasses
    from cfgpp_b200 import vae as V
    from oracle import vae as OV
    for cfg in (V.sdxl_vae_config(), V.sd15_vae_config()):
        assert V.num_vae_encoder_params(cfg) == 34_163_592 + 72
        assert V.num_vae_encoder_params(cfg) + V.num_vae_decoder_params(cfg) == 83_653_863
        with torch.device("meta"):
            m = OV.AutoencoderKLEncoder(OV.VAEConfig(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(OV.VAEConfig)}))
        assert OV.count_params(m) == V.num_vae_encoder_params(cfg) and OV.count_params(m.encoder) == 34_163_592
        assert set(m.state_dict().keys()) == {k for k, _, _ in V.vae_encoder_param_specs(cfg)}
    cfg = V.tiny_vae_config()
    sd = V.synthetic_vae_state_dict(cfg, seed=3, with_encoder=True)
    assert {k for k in sd if k.startswith(("decoder.", "post_quant_conv."))} == set(V.synthetic_vae_state_dict(cfg, seed=3))
    ocfg = OV.VAEConfig(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(OV.VAEConfig)})
    m = OV.build_vae_encoder(ocfg, {k: v for k, v in sd.items() if k.startswith(("encoder.", "quant_conv."))})
    x = torch.rand(1, 3, 64, 64) * 2 - 1
    mean, std = m(x)
    noise = torch.randn_like(mean)
    assert torch.allclose(OV.encode(m, x, noise), (mean + std * noise) * cfg.scaling_factor)
    m.quant_conv.bias.data[4:] = 100.0      # log-variance far above the clamp
    _, std_hi = m(x)
    assert torch.allclose(std_hi, torch.full_like(std_hi, float(torch.exp(torch.tensor(10.0)))))
    m.quant_conv.bias.data[4:] = -100.0
    _, std_lo = m(x)
    assert torch.allclose(std_lo, torch.full_like(std_lo, float(torch.exp(torch.tensor(-15.0)))))

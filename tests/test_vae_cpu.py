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

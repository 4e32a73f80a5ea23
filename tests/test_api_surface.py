"""Host-side contract: solver registries (names, errors), the C-ABI library (loads, exports every declared symbol),
loud failure without CUDA, struct layouts, weight-spec coverage. No GPU compute here."""
import ctypes
import re
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_registries_mirror_reference_names_and_errors():
    from cfgpp_b200 import latent_diffusion as LD
    from cfgpp_b200 import latent_sdxl as LX
    assert {"ddim_cfg++", "ddim_inversion_cfg++"} <= set(LD.__SOLVER__)
    assert {"ddim_cfg++", "ddim_cfg++_lightning", "dpm++_2m_cfgpp"} <= set(LX.__SOLVER__)
    # the rest of the CFG++ --method surface
    assert {"euler_cfg++", "euler_a_cfg++", "dpm++_2s_a_cfg++", "dpm++_2m_cfg++", "ddim_edit_cfg++"} <= set(LD.__SOLVER__)
    assert {"euler_cfg++", "euler_cfg++_lightning", "dpm++_2m_cfgpp_lightning", "ddim_edit_cfg++"} <= set(LX.__SOLVER__)
    # section 8 f4: the plain-CFG baselines -> the registries now equal the reference's --method surface
    assert set(LD.__SOLVER__) == {"ddim", "euler", "euler_a", "dpm++_2s_a", "dpm++_2m", "ddim_inversion", "ddim_edit",
                                  "ddim_cfg++", "euler_cfg++", "euler_a_cfg++", "dpm++_2s_a_cfg++", "dpm++_2m_cfg++",
                                  "ddim_inversion_cfg++", "ddim_edit_cfg++"}
    assert set(LX.__SOLVER__) == {"ddim", "euler", "ddim_lightning", "euler_lightning", "ddim_edit", "ddim_cfg++",
                                  "euler_cfg++", "euler_cfg++_lightning", "ddim_cfg++_lightning", "dpm++_2m_cfgpp",
                                  "dpm++_2m_cfgpp_lightning", "ddim_edit_cfg++"}
    with pytest.raises(ValueError, match="does not exist"):
        LX.get_solver("no_such_solver")
    with pytest.raises(ValueError, match="already registered"):
        LX.register_solver("ddim_cfg++")(object)
    with pytest.raises(ValueError, match="already registered"):
        LD.register_solver("ddim_cfg++")(object)


def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "cfgpp_b200.h").read_text()
    declared = set(re.findall(r"\b(cfgpp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    from cfgpp_b200 import _native as nv
    lib = nv.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"symbols declared in include/cfgpp_b200.h but not exported: {missing}"
    assert lib.cfgpp_version() >= 100


def test_abi_struct_layouts():
    from cfgpp_b200.config import ModelDescC
    from cfgpp_b200.schedule import StepCoefC, StepStateC
    assert ctypes.sizeof(StepCoefC) == 40 and ctypes.sizeof(StepStateC) == 48
    assert ctypes.sizeof(ModelDescC) == 4 * (3 + 4 * 3 + 1 + 4 * 2 + 7)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    from cfgpp_b200 import latent_sdxl as LX
    from cfgpp_b200.config import tiny_sdxl_config
    with pytest.raises(RuntimeError, match="CUDA"):
        LX.get_solver("ddim_cfg++", solver_config=SimpleNamespace(num_sampling=4), device="cpu",
                      unet_config=tiny_sdxl_config(), model_key="synthetic:1")


def test_product_does_not_import_oracle():
    for f in (ROOT / "cfgpp_b200").rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
        assert "from .. import oracle" not in src


def test_step_coef_tables_are_float32_exact():
    """The host evaluates the per-step scalars with fp32 torch CPU ops, as the reference loops do."""
    from cfgpp_b200 import schedule as S
    sch = S.Schedule.make(50)
    st = S.ddim_cfgpp_steps(sch, 0.6, True)[7]
    t = int(st.t)
    at = sch.alphas_cumprod[t]
    assert st.coef.c0 == float((1 - at).sqrt()) and st.coef.c1 == float(at.sqrt())
    assert st.coef.lambda_ == float(torch.tensor(0.6, dtype=torch.float32))


def test_draw_callbacks_dump_decoded_latents(tmp_path):
    """utils/callback_util.py:39-65 of the reference: decode z0t / zt of the step, write under <workdir>/record/."""
    import torch
    from cfgpp_b200.utils.callback_util import ComposeCallback
    cb = ComposeCallback(workdir=tmp_path, frequency=1, callbacks=["draw_noisy", "draw_tweedie"])
    kw = {"z0t": torch.zeros(1, 4, 8, 8), "zt": torch.ones(1, 4, 8, 8),
          "decode": lambda z: z[:, :3].repeat_interleave(8, 2).repeat_interleave(8, 3)}
    out = cb(0, torch.tensor(901), kw)
    assert out is kw
    assert len(list((tmp_path / "record" / "tweedie").glob("x0_901.*"))) == 1
    assert len(list((tmp_path / "record" / "noisy").glob("xt_901.*"))) == 1


def test_checkpoint_resolution_round_trip(tmp_path):
    """model_key = a diffusers-format *.safetensors file -> those exact weights; 'synthetic:<seed>' -> seeded weights;
    an HF hub id (nothing can be downloaded here) -> synthetic with a warning."""
    import pytest
    import torch
    from safetensors.torch import save_file
    from cfgpp_b200 import weights as Wt
    from cfgpp_b200.config import tiny_sdxl_config
    from cfgpp_b200.latent_sdxl import resolve_state_dict
    cfg = tiny_sdxl_config()
    sd = Wt.synthetic_state_dict(cfg, seed=77, device="cpu")
    path = tmp_path / "unet.safetensors"
    save_file({k: v.contiguous() for k, v in sd.items()}, str(path))
    back = resolve_state_dict(str(path), cfg, "cpu")
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    assert [k for k, _, _ in Wt.unet_param_specs(cfg)] == list(sd)       # diffusers key names, spec order
    again = resolve_state_dict("synthetic:77", cfg, "cpu")
    assert all(torch.equal(again[k], sd[k]) for k in sd)
    with pytest.warns(UserWarning, match="no checkpoint"):
        other = resolve_state_dict("stabilityai/stable-diffusion-xl-base-1.0", cfg, "cpu")
    assert set(other) == set(sd)


def test_pipeline_directory_resolver(tmp_path):
    """cfgpp_b200.checkpoints: the diffusers pipeline layout the reference's from_pretrained reads (latent_sdxl.py:41-49)."""
    from cfgpp_b200.checkpoints import find_pipeline_files
    import pytest as _pt
    with _pt.raises(FileNotFoundError) as e:
        find_pipeline_files(tmp_path, "sdxl")
    assert "text_encoder_2" in str(e.value) and "tokenizer_2" in str(e.value) and "unet" in str(e.value)
    for d, name in (("unet", "diffusion_pytorch_model.fp16.safetensors"), ("vae", "diffusion_pytorch_model.safetensors"),
                    ("text_encoder", "model.safetensors"), ("text_encoder_2", "model.fp16.safetensors")):
        (tmp_path / d).mkdir()
        (tmp_path / d / name).write_bytes(b"")
    for t in ("tokenizer", "tokenizer_2"):
        (tmp_path / t).mkdir()
        (tmp_path / t / "vocab.json").write_text("{}")
        (tmp_path / t / "merges.txt").write_text("#version: 0.2\n")
    f = find_pipeline_files(tmp_path, "sdxl")
    assert f["unet"].name.endswith(".fp16.safetensors") and f["text_encoder_2"].parent.name == "text_encoder_2"
    assert set(find_pipeline_files(tmp_path, "sd15")) == {"unet", "vae", "text_encoder", "tokenizer/vocab.json", "tokenizer/merges.txt"}
    with _pt.raises(ValueError):
        find_pipeline_files(tmp_path, "sd3")

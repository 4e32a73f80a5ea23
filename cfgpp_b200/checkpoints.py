"""Where a diffusers-format pipeline directory keeps the files the native components load — the layout
`StableDiffusionPipeline.from_pretrained` / `StableDiffusionXLPipeline.from_pretrained` read in the reference
(latent_diffusion.py:62-66, latent_sdxl.py:41-49):

    <dir>/unet/diffusion_pytorch_model[.fp16].safetensors
    <dir>/vae/diffusion_pytorch_model[.fp16].safetensors        (SDXL: the reference swaps in madebyollin/sdxl-vae-fp16-fix)
    <dir>/text_encoder/model[.fp16].safetensors      + <dir>/tokenizer/{vocab.json, merges.txt}
    <dir>/text_encoder_2/model[.fp16].safetensors    + <dir>/tokenizer_2/{vocab.json, merges.txt}     (SDXL only)

`solver_components(dir, family, device)` turns it into the keyword arguments of `get_solver(...)`. Nothing can be
downloaded here; without a directory every component falls back to seeded synthetic weights."""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional


def _weights(folder: Path, stems) -> Optional[Path]:
    for stem in stems:
        for name in (f"{stem}.fp16.safetensors", f"{stem}.safetensors"):
            if (folder / name).is_file():
                return folder / name
    return None


def find_pipeline_files(ckpt_dir, family: str) -> Dict[str, Path]:
    """family: 'sd15' | 'sdxl'. Raises FileNotFoundError naming every missing piece."""
    root = Path(ckpt_dir)
    want = {"unet": (root / "unet", ("diffusion_pytorch_model",)), "vae": (root / "vae", ("diffusion_pytorch_model",)),
            "text_encoder": (root / "text_encoder", ("model",))}
    toks = ["tokenizer"]
    if family == "sdxl":
        want["text_encoder_2"] = (root / "text_encoder_2", ("model",))
        toks.append("tokenizer_2")
    elif family != "sd15":
        raise ValueError(f"unknown model family {family!r}")
    found: Dict[str, Path] = {}
    missing = []
    for key, (folder, stems) in want.items():
        p = _weights(folder, stems)
        if p is None:
            missing.append(f"{folder}/{stems[0]}[.fp16].safetensors")
        else:
            found[key] = p
    for t in toks:
        for name in ("vocab.json", "merges.txt"):
            p = root / t / name
            if p.is_file():
                found[f"{t}/{name}"] = p
            else:
                missing.append(str(p))
    if missing:
        raise FileNotFoundError("pipeline directory is incomplete, missing: " + ", ".join(missing))
    return found


def solver_components(ckpt_dir, family: str, device) -> dict:
    """Keyword arguments for `latent_sdxl.get_solver` / `latent_diffusion.get_solver` that load every component of the
    pipeline directory on the native backend (UNet, VAE encoder + decoder, CLIP text towers + BPE tokenizers)."""
    from .text_encoder import get_conditioner
    from .vae import get_vae
    f = find_pipeline_files(ckpt_dir, family)
    kw = {"model_key": str(f["unet"])}
    if family == "sdxl":
        kw["vae"] = get_vae("sdxl_vae", device, str(f["vae"]))
        kw["text_encoders"] = (
            get_conditioner("clip_l", device, "sdxl", str(f["text_encoder"]), str(f["tokenizer/vocab.json"]),
                            str(f["tokenizer/merges.txt"])),
            get_conditioner("clip_bigg", device, "sdxl", str(f["text_encoder_2"]), str(f["tokenizer_2/vocab.json"]),
                            str(f["tokenizer_2/merges.txt"])))
    else:
        kw["vae"] = get_vae("sd15_vae", device, str(f["vae"]))
        kw["text_encoder"] = get_conditioner("clip_l", device, "sd15", str(f["text_encoder"]),
                                             str(f["tokenizer/vocab.json"]), str(f["tokenizer/merges.txt"]))
    return kw

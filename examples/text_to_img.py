"""CLI with the reference's flag surface (examples/text_to_img.py:14-24):
    python -m examples.text_to_img --model sdxl --method ddim_cfg++ --cfg_guidance 0.6 --NFE 50 --prompt "..."
Runs on the Blackwell-native backend. Without checkpoints (offline) the UNet weights are seeded synthetic and the
text encoder / VAE are stand-ins (cfgpp_b200/conditioning.py), so the PNG is only a plumbing check."""
import argparse
from pathlib import Path
from types import SimpleNamespace

import torch

from cfgpp_b200.checkpoints import solver_components
from cfgpp_b200.latent_diffusion import get_solver
from cfgpp_b200.latent_sdxl import get_solver as get_solver_sdxl
from cfgpp_b200.utils.log_util import create_workdir, set_seed


def main():
    parser = argparse.ArgumentParser(description="Latent Diffusion")
    parser.add_argument("--workdir", type=Path, default="examples/workdir/t2i")
    parser.add_argument("--device", type=str, default="cuda")
    parser.add_argument("--null_prompt", type=str, default="low quality,jpeg artifacts,blurry,poorly drawn,ugly,worst quality,")
    parser.add_argument("--prompt", type=str, default="")
    parser.add_argument("--cfg_guidance", type=float, default=7.5)
    parser.add_argument("--method", type=str, default='ddim_cfg++')
    parser.add_argument("--model", type=str, default='sd15', choices=["sd15", "sd20", "sdxl", "sdxl_lightning"])
    parser.add_argument("--NFE", type=int, default=50)
    parser.add_argument("--seed", type=int, default=42)
    parser.add_argument("--ckpt_dir", type=Path, default=None,
                        help="diffusers-format pipeline directory (unet/, vae/, text_encoder[_2]/, tokenizer[_2]/); "
                             "default: seeded synthetic weights (nothing can be downloaded here)")
    args = parser.parse_args()

    set_seed(args.seed)
    create_workdir(args.workdir)
    solver_config = SimpleNamespace(num_sampling=args.NFE)  # the reference munchifies {'num_sampling': NFE}
    callback = None

    if args.model in ("sdxl", "sdxl_lightning"):
        extra = solver_components(args.ckpt_dir, "sdxl", args.device) if args.ckpt_dir else {}
        solver = get_solver_sdxl(args.method, solver_config=solver_config, device=args.device, **extra)
        result = solver.sample(prompt1=[args.null_prompt, args.prompt], prompt2=[args.null_prompt, args.prompt],
                               cfg_guidance=args.cfg_guidance, target_size=(1024, 1024), callback_fn=callback)
    else:
        extra = solver_components(args.ckpt_dir, "sd15", args.device) if args.ckpt_dir else {}
        solver = get_solver(args.method, solver_config=solver_config, device=args.device, **extra)
        result = solver.sample(prompt=[args.null_prompt, args.prompt], cfg_guidance=args.cfg_guidance,
                               callback_fn=callback)

    out = args.workdir.joinpath('result/generated.pt')
    torch.save(result, out)
    try:
        from torchvision.utils import save_image
        save_image(result, args.workdir.joinpath('result/generated.png'), normalize=True)
    except Exception:  # torchvision is optional here
        pass
    print(f"saved {out}")


if __name__ == "__main__":
    main()

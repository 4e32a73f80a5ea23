"""Many-prompt generation with the reference's flag surface (examples/text_to_mscoco.py:14-26) —
    python -m examples.text_to_mscoco --model sdxl --method ddim_cfg++ --cfg_guidance 0.6 --prompt_dir prompts.txt
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 -m examples.text_to_mscoco --model sdxl ...
The reference loops over the prompts on one GPU. Here prompt i goes to rank i % world: every rank
holds a full UNet replica, rank 0's weights are broadcast once over NCCL so the replicas are bit-identical, and there
is no collective inside the sampling loop. Every rank draws the zT of EVERY image from the seeded CPU generator, in
order, and keeps its own - so image i is the same picture whatever the world size."""
import argparse
import os
from pathlib import Path
from types import SimpleNamespace

import torch

from cfgpp_b200 import dist as D
from cfgpp_b200 import weights as Wt
from cfgpp_b200.config import sd15_config, sdxl_config
from cfgpp_b200.latent_diffusion import get_solver
from cfgpp_b200.latent_sdxl import get_solver as get_solver_sdxl
from cfgpp_b200.utils.log_util import create_workdir, set_seed


def read_prompts(path: Path, limit: int = 10000):
    if not Path(path).exists():
        return [f"synthetic prompt {i}" for i in range(8)]  # offline: no MS-COCO caption file
    with open(path, 'r') as f:
        return [ln.strip() for ln in f if ln.strip()][:limit]


def main():
    parser = argparse.ArgumentParser(description="Latent Diffusion")
    parser.add_argument("--workdir", type=Path, default="examples/workdir/mscoco")
    parser.add_argument('--prompt_dir', type=Path, default=Path('examples/assets/coco_v2.txt'))
    parser.add_argument("--device", type=str, default="cuda")
    parser.add_argument("--null_prompt", type=str, default="")
    parser.add_argument("--prompt", type=str, default="")
    parser.add_argument("--cfg_guidance", type=float, default=7.5)
    parser.add_argument("--method", type=str, default='ddim')
    parser.add_argument("--model", type=str, default='sd15', choices=["sd15", "sd20", "sdxl", "sdxl_lightning"])
    parser.add_argument("--NFE", type=int, default=50)
    parser.add_argument("--seed", type=int, default=42)
    parser.add_argument("--ckpt_dir", type=Path, default=None,
                        help="diffusers-format pipeline directory (unet/, vae/, text_encoder[_2]/, tokenizer[_2]/); "
                             "default: seeded synthetic weights (nothing can be downloaded here)")
    args = parser.parse_args()

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local) if world > 1 else torch.device(args.device)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=device)

    set_seed(args.seed)
    if rank == 0:
        create_workdir(args.workdir)
    text_list = read_prompts(args.prompt_dir)
    solver_config = SimpleNamespace(num_sampling=args.NFE)
    sdxl = args.model in ("sdxl", "sdxl_lightning")
    cfg = sdxl_config() if sdxl else sd15_config()

    kw = {}
    if args.ckpt_dir:  # every rank reads the pipeline directory itself (UNet, VAE, text towers, tokenizers)
        from cfgpp_b200.checkpoints import solver_components
        kw = solver_components(args.ckpt_dir, "sdxl" if sdxl else "sd15", device)
    elif world > 1:  # one bucketed NCCL broadcast of rank 0's weights; afterwards the ranks never talk again
        sd = Wt.synthetic_state_dict(cfg, seed=1234, device=device) if rank == 0 else None
        kw["state_dict"] = D.broadcast_state_dict(sd, Wt.unet_param_specs(cfg), device, src=0)
    solver = (get_solver_sdxl if sdxl else get_solver)(args.method, solver_config=solver_config, device=device, **kw)

    mine = set(D.shard_indices(len(text_list), rank, world))
    latent = (1, 4, cfg.sample_size, cfg.sample_size)
    for i, text in enumerate(text_list):
        zT = torch.randn(latent)  # CPU generator, advanced for every image on every rank (see module docstring)
        if i not in mine:
            continue
        print(f'[rank {rank}] processing {i + 1}/{len(text_list)}: {text}', flush=True)
        if sdxl:
            result = solver.sample(prompt1=[args.null_prompt, text], prompt2=[args.null_prompt, text],
                                   cfg_guidance=args.cfg_guidance, target_size=(1024, 1024), zT=zT)
        else:
            result = solver.sample(prompt=[args.null_prompt, text], cfg_guidance=args.cfg_guidance, zT=zT)
        torch.save(result, args.workdir.joinpath(f'{str(i).zfill(5)}.pt'))
        try:
            from torchvision.utils import save_image
            save_image(result, args.workdir.joinpath(f'{str(i).zfill(5)}.png'), normalize=True)
        except Exception:  # torchvision is optional here
            pass
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""N>1 host logic on CPU: gloo, world size 2 (the GPU path uses the same code over NCCL): round-robin prompt
sharding, the bucketed weight broadcast (bit-identical replicas) and the final latent gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cfgpp_b200 import dist as D
    from cfgpp_b200 import weights as Wt
    from cfgpp_b200.config import tiny_sdxl_config
    cfg = tiny_sdxl_config()
    specs = Wt.unet_param_specs(cfg)
    sd = Wt.synthetic_state_dict(cfg, seed=5) if rank == 0 else None
    got = D.broadcast_state_dict(sd, specs, torch.device("cpu"), src=0, bucket_elems=1 << 20)  # several buckets
    ref = Wt.synthetic_state_dict(cfg, seed=5)
    same = all(torch.equal(got[k], ref[k]) for k in ref) and set(got) == set(ref)
    n_items = 7
    mine = D.shard_indices(n_items, rank, world)
    local = torch.stack([torch.full((4, 2, 2), float(i)) for i in mine])
    full = D.gather_latents(local, n_items, dst=0)
    ok_gather = True
    if rank == 0:
        ok_gather = all(float(full[i].mean()) == float(i) for i in range(n_items))
    out.put((rank, same, mine, ok_gather))
    dist.destroy_process_group()


def test_broadcast_shard_gather_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1], "broadcast replicas differ from the source state dict"
    assert res[0][2] == [0, 2, 4, 6] and res[1][2] == [1, 3, 5]
    assert res[0][3]


def test_shard_indices_cover_everything_once():
    from cfgpp_b200.dist import shard_indices
    for world in (1, 2, 4, 8):
        allidx = sorted(i for r in range(world) for i in shard_indices(64, r, world))
        assert allidx == list(range(64))
        assert max(len(shard_indices(64, r, world)) for r in range(world)) == 64 // world

"""Profiling target for ncu (one GPU): builds the SDXL engine (batch 2, 1024x1024), warms up, then brackets exactly one
fused CFG++ step (CUDA graph replay) with cudaProfilerStart/Stop.
    ncu --profile-from-start off ... python tools/ncu_target.py [nsteps]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cfgpp_b200 import config as C, schedule as S, weights as Wt  # noqa: E402
from cfgpp_b200.engine import NativeUNet  # noqa: E402

dev = torch.device("cuda:0")
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = C.sdxl_config()
sd = Wt.synthetic_state_dict(cfg, seed=1234, device=dev)
g = torch.Generator().manual_seed(0)
B, hw = 2, 128
z = torch.randn(B, 4, hw, hw, generator=g).to(dev)
ctx = torch.randn(2 * B, 77, cfg.cross_attention_dim, generator=g).half().to(dev)
pooled = torch.randn(2 * B, cfg.pooled_dim, generator=g).half().to(dev)
tid = torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * (2 * B)).to(dev)
net = NativeUNet(cfg, sd, dev)
del sd
net.prepare(B, hw, hw)
net.set_prompt(ctx, pooled, tid)
net.set_schedule(S.STEP_DDIM_CFGPP, torch.float32, S.ddim_cfgpp_steps(S.Schedule.make(50), 0.6, True))
net.set_state(z)
net.run_steps(0, 2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
net.run_steps(2, nsteps)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", nsteps, "fused step(s);", net.launches_per_step, "launches/step")

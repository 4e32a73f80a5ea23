"""bench.py — headline benchmark of the CFG++ sampling hot path (driver contract in the task statement).

    python bench.py --gpus N --steps K --warmup W            # ours: hand-written sm_100a path behind the C ABI
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the repo's CPU eager path (oracle)

Metric (BASELINE.json): images/sec, device-timed, SDXL 1024x1024 NFE=50 ddim_cfg++ lambda=0.6, batch 2 per GPU
(= configs[2]); N>1 shards independent prompts over ranks (weak scaling, no per-step collective, one NCCL broadcast
of the UNet weights at init). One "step" = one full sampling trajectory of one batch (NFE fused UNet+CFG++ steps):
from zT resident in HBM to the final latent z0t — text encoding and VAE decode stay on the reference path and are
outside the metric.

Synthetic data: no checkpoint / tokenizer exists offline, so weights are seeded synthetic under the diffusers key
names (random-init of the real SDXL architecture, 2,567,463,684 params) and the conditioning tensors are seeded
random embeddings of the real shapes.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

NFE = 50
LAMBDA = 0.6
BATCH = 2
LATENT = 128
METRIC = "images/sec (device-timed) SDXL 1024x1024 NFE=50 ddim_cfg++"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, mean over the 12 consecutive
# gemm_kernel launches of a 1280-channel transformer block in one `ncu --set full` capture (cold L2: an upper bound;
# it equals the algorithmic A + W + residual bytes, i.e. no re-reads) — profiles/r01_v3_ncu_full.md
NCU_GEMM_DRAM_BYTES_PER_LAUNCH = 33.45e6
NCU_TRAFFIC_SOURCE = "ncu --set full, gpurun_out/prof_gemm.ncu-rep (round 1 v3), summarised in profiles/r01_v3_ncu_full.md"


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops": d.get("bf16_tflops_sustained", 1445.3), "hbm": d.get("hbm_gbs", 6587.7), "source": "measured"}
    return {"tflops": 1400.0, "hbm": 6650.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "power_w_median": statistics.median(pw) if pw else None, "power_w_max": max(pw) if pw else None}


# ----------------------------------------------------------------------------------------------------------------
# synthetic workload
# ----------------------------------------------------------------------------------------------------------------
def synthetic_conditioning(cfg, batch, seed, pin=True):
    """Host-side (pinned) conditioning + zT for one batch, shapes of latent_sdxl.py:222-257 / :289."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    t = {
        "uc": torch.randn(batch, 77, cfg.cross_attention_dim, generator=g).half(),
        "c": torch.randn(batch, 77, cfg.cross_attention_dim, generator=g).half(),
        "pooled": torch.randn(2 * batch, cfg.pooled_dim, generator=g).half(),
        "time_ids": torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]] * (2 * batch)).half(),
    }
    g2 = torch.Generator(device="cpu").manual_seed(42 + seed)
    t["zT"] = torch.randn(batch, 4, LATENT, LATENT, generator=g2)
    if pin and torch.cuda.is_available():
        t = {k: v.pin_memory() for k, v in t.items()}
    return t


def nbytes(*ts):
    return int(sum(x.numel() * x.element_size() for x in ts))


# ----------------------------------------------------------------------------------------------------------------
# CPU baseline = the oracle ("repo's own CPU eager path"), bounded sample
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(sd_provider, max_seconds=150.0):
    """Times ONE SDXL UNet sample-forward (batch 1 at 128x128 latent, fp32, all host threads) of the oracle and
    extrapolates to images/sec: one image = NFE x 2 such forwards (uncond + cond). Bounded: a GEMM probe first
    decides whether the full forward fits the time budget; if not, a 64x64-latent forward is timed and scaled by the
    algorithmic FLOP ratio (stated in `sample`)."""
    import dataclasses
    from cfgpp_b200 import config as C
    from oracle import unet as O
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cfg = C.sdxl_config()
    ocfg = O.UNetConfig(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(O.UNetConfig)})
    # throughput probe (fp32 GEMM shaped like the dominant FF layer)
    a, b = torch.randn(1024, 5120), torch.randn(5120, 1280)
    torch.mm(a, b)
    t0 = time.perf_counter()
    for _ in range(3):
        torch.mm(a, b)
    gflops = 3 * 2 * 1024 * 5120 * 1280 / (time.perf_counter() - t0) / 1e9
    full_flops = 6.7612e12
    latent = LATENT if full_flops / (gflops * 1e9 * 0.6) < max_seconds else 64
    sd = sd_provider()
    m = O.build_unet(ocfg, sd, dtype=torch.float32, device="cpu")
    del sd
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, latent, latent, generator=g)
    ctx = torch.randn(1, 77, cfg.cross_attention_dim, generator=g)
    add = {"text_embeds": torch.randn(1, cfg.pooled_dim, generator=g),
           "time_ids": torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]])}
    with torch.no_grad():
        t0 = time.perf_counter()
        m(x, torch.tensor(501), ctx, add)
        dt = time.perf_counter() - t0
    if latent == LATENT:
        t_fwd, how = dt, "1 SDXL UNet sample-forward (batch 1, 128x128 latent, fp32) = 1/100 of one image"
    else:
        # algorithmic FLOPs per sample-forward: 6.7612 T at 128x128 of which self-attention 0.7516 T;
        # at 64x64 convs / token GEMMs shrink 4x and self-attention 16x
        f128 = 6.7612
        f64 = (f128 - 0.7516) / 4 + 0.7516 / 16
        t_fwd = dt * f128 / f64
        how = ("1 SDXL UNet sample-forward at 64x64 latent scaled to 128x128 by algorithmic FLOPs "
               "(the full-size forward would exceed the time budget) = 1/100 of one image")
    value = 1.0 / (2 * NFE * t_fwd)
    return {"value": value, "unit": "images/sec", "cores": cores, "kind": "port", "sample": how,
            "seconds_per_unet_forward": t_fwd, "probe_gflops": gflops}


# ----------------------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path = the oracle port (the reference itself
    cannot be imported: diffusers is absent and not installable offline), all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cfgpp_b200 import config as C, weights as Wt
    cfg = C.sdxl_config()
    vals = []
    for _ in range(max(1, min(args.steps, 2))):  # each "step" is one bounded sample; keep the run to a few minutes
        r = cpu_reference_sample(lambda: Wt.synthetic_state_dict(cfg, seed=1234, device="cpu", dtype=torch.float16))
        vals.append(r)
    best = max(vals, key=lambda r: r["value"])
    line = {"impl": "reference", "metric": METRIC, "value": best["value"], "unit": "images/sec", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * BATCH / best["value"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SDXL 1024x1024 ddim_cfg++ lambda=0.6 NFE=50 batch=2 (configs[2])",
                       "global_batch": BATCH * args.gpus, "parallelism": f"dp{args.gpus}"},
            "cpu_baseline": best,
            "e2e": {"value": best["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from cfgpp_b200 import config as C, schedule as S, weights as Wt
    from cfgpp_b200 import dist as D
    from cfgpp_b200.engine import NativeUNet
    from cfgpp_b200.latent_sdxl import get_solver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours) needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line (NCCL prints its version banner there)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = C.sdxl_config()
    # ---- weights: rank 0 generates, ONE bucketed NCCL broadcast at init makes replicas bit-identical -------------
    sd = Wt.synthetic_state_dict(cfg, seed=1234, device=dev) if rank == 0 else None
    if world > 1:
        sd = D.broadcast_state_dict(sd, Wt.unet_param_specs(cfg), dev, src=0)
    solver = get_solver("ddim_cfg++", solver_config=argparse.Namespace(num_sampling=NFE), device=dev,
                        model_key="synthetic:1234", state_dict=sd)
    eng: NativeUNet = solver.unet
    eng.prepare(BATCH, LATENT, LATENT)
    steps = S.ddim_cfgpp_steps(S.Schedule.make(NFE), LAMBDA, sdxl_indexing=True)

    # every trajectory uses its own prompt / zT (rank r owns items r, r+W, ...: D.shard_indices)
    n_traj = args.warmup + args.steps
    items = D.shard_indices(world * n_traj, rank, world)
    host = [synthetic_conditioning(cfg, BATCH, seed=it) for it in items]
    dev_in = [{k: v.to(dev) for k, v in h.items()} for h in host]
    torch.cuda.synchronize()

    def trajectory_device(i):
        """inputs already resident in HBM (the `value` leg)."""
        d = dev_in[i]
        eng.set_prompt(torch.cat([d["uc"], d["c"]]), d["pooled"], d["time_ids"].float())
        eng.set_schedule(S.STEP_DDIM_CFGPP, torch.float32, steps)
        eng.set_state(d["zT"])
        eng.run_steps(0, NFE)
        return eng.get_state(1)

    def trajectory_e2e(i):
        """public solver API with HOST buffers: H2D of this step's inputs, D2H of the result (the `e2e` leg)."""
        h = host[i]
        uc, c = h["uc"].to(dev, non_blocking=True), h["c"].to(dev, non_blocking=True)
        add = {"text_embeds": h["pooled"].to(dev, non_blocking=True), "time_ids": h["time_ids"].to(dev, non_blocking=True)}
        zT = h["zT"].to(dev, non_blocking=True)
        z0t = solver.reverse_process(uc, c, LAMBDA, add, (1024, 1024), zT=zT)
        return z0t.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, first, count):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(first, first + count):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for i in range(args.warmup):
        trajectory_device(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(trajectory_device, args.warmup, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # e2e leg (reuses the same warm engine; its own warm-up trajectory first)
    trajectory_e2e(0)
    ms_e2e = timed(trajectory_e2e, args.warmup, args.steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    imgs = BATCH * args.steps * world
    value = imgs / (ms_dev / 1e3)
    e2e_val = imgs / (ms_e2e / 1e3)
    peaks = measured_peaks()

    # ---- roofline of the dominant kernel (tcgen05 GEMM / implicit-GEMM conv), measured live with CUDA events ----
    d = dev_in[0]
    eng.set_prompt(torch.cat([d["uc"], d["c"]]), d["pooled"], d["time_ids"].float())
    eng.profile_forward(d["zT"], 501.0)
    prof = eng.profile_forward(d["zT"], 501.0)
    by_kind = {0: [0.0, 0.0, 0], 1: [0.0, 0.0, 0], 2: [0.0, 0.0, 0], 3: [0.0, 0.0, 0]}
    for _, kind, fl, ms in prof:
        by_kind[kind][0] += fl
        by_kind[kind][1] += ms
        by_kind[kind][2] += 1
    gemm_fl = by_kind[0][0] + by_kind[1][0]
    gemm_ms = by_kind[0][1] + by_kind[1][1]
    gemm_n = by_kind[0][2] + by_kind[1][2]
    tot_ms = sum(v[1] for v in by_kind.values())
    achieved = gemm_fl / (gemm_ms / 1e3) / 1e12
    roofline = {"bound": "tensor", "kernel": "gemm_kernel<BN,GEGLU,CL> (tcgen05 GEMM + implicit-GEMM conv3x3)",
                "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                "traffic": NCU_GEMM_DRAM_BYTES_PER_LAUNCH, "traffic_source": NCU_TRAFFIC_SOURCE,
                "peak_source": peaks["source"] + " (bf16 sustained)",
                "flops_per_launch": gemm_fl / max(gemm_n, 1), "launches_per_forward": gemm_n,
                "avg_launch_us": 1e3 * gemm_ms / max(gemm_n, 1), "share_of_step": gemm_ms / tot_ms,
                "by_kind_ms": {"linear_gemm": by_kind[0][1], "conv3x3": by_kind[1][1], "attention": by_kind[2][1],
                               "norm_elementwise": by_kind[3][1]},
                "by_kind_tflops": {"linear_gemm": by_kind[0][0] / max(by_kind[0][1], 1e-9) / 1e9,
                                   "conv3x3": by_kind[1][0] / max(by_kind[1][1], 1e-9) / 1e9,
                                   "attention": by_kind[2][0] / max(by_kind[2][1], 1e-9) / 1e9}}
    step_tflops = eng.forward_flops * NFE * args.steps * world / (ms_dev / 1e3) / 1e12 / world

    line = {"metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "SDXL 1024x1024 ddim_cfg++ lambda=0.6 NFE=50 batch=2 per GPU (configs[2])",
                       "global_batch": BATCH * world, "parallelism": f"dp{world} (independent prompts, no per-step collective)",
                       "step": "one full NFE=50 trajectory of one batch (UNet uncond+cond + fused CFG++/DDIM update per step)",
                       "l2": "inputs larger than L2 (5.1 GB fp16 weights streamed every UNet forward; 126 MB L2)"},
            "e2e": {"value": e2e_val, "unit": "images/sec",
                    "h2d_bytes_per_step": nbytes(host[0]["uc"], host[0]["c"], host[0]["pooled"], host[0]["time_ids"],
                                                 host[0]["zT"]),
                    "d2h_bytes_per_step": BATCH * 4 * LATENT * LATENT * 4,
                    "api": "cfgpp_b200.latent_sdxl.get_solver('ddim_cfg++').reverse_process(...) with pinned host inputs"},
            "gpu_launches": int((eng.launches_per_step * NFE + 80) * args.steps),
            "clocks": clocks, "roofline": roofline,
            "unet_tflops_per_gpu": step_tflops, "unet_frac_of_peak": step_tflops / peaks["tflops"],
            "forward_tflop": eng.forward_flops / 1e12}

    # ---- baselines measured beside it (rank 0, N=1 only) ---------------------------------------------------------
    if world == 1 and not args.no_baselines:
        del solver, eng
        torch.cuda.empty_cache()
        line["gpu_eager_baseline"] = gpu_eager_baseline(cfg, sd, dev)
        line["speedup_vs_gpu_eager"] = value / line["gpu_eager_baseline"]["value"]
        sd_cpu = {k: v.cpu() for k, v in sd.items()}
        del sd
        torch.cuda.empty_cache()
        line["cpu_baseline"] = cpu_reference_sample(lambda: sd_cpu)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def gpu_eager_baseline(cfg, sd, dev, n_time=6, n_warm=2):
    """The north-star comparator ("reference CUDA path" stand-in): the restated diffusers op sequence
    + the reference-style Python step loop (incl. its per-step host syncs) under torch.autocast('cuda', fp16) on the
    same GPU and inputs. Times `n_time` steps after `n_warm` and scales to NFE (every step costs the same)."""
    import dataclasses
    from oracle import samplers as OSm, schedule as OS, unet as O
    ocfg = O.UNetConfig(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(O.UNetConfig)})
    m = O.build_unet(ocfg, sd, dtype=torch.float16, device=dev)
    h = synthetic_conditioning(cfg, BATCH, seed=0, pin=False)
    uc, c, zT = h["uc"].to(dev), h["c"].to(dev), h["zT"].to(dev)
    add = {"text_embeds": h["pooled"].to(dev), "time_ids": h["time_ids"].to(dev)}
    tb = OS.make_tables(NFE)

    def run(nsteps):
        tb_n = dataclasses.replace(tb, timesteps=tb.timesteps[:nsteps])
        with torch.autocast("cuda", dtype=torch.float16):
            return OSm.sdxl_ddim_cfgpp(m, tb_n, zT, uc, c, LAMBDA, add)

    run(n_warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(n_time)
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / n_time
    return {"value": BATCH / (NFE * ms_step / 1e3), "unit": "images/sec", "ms_per_unet_step": ms_step,
            "what": "restated diffusers UNet + reference step loop, torch eager, autocast fp16, same GPU"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-baselines", action="store_true", help="skip the GPU-eager and CPU baselines at N=1")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

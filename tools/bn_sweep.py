"""Tile-width sweep: time every GEMM / conv shape of the SDXL UNet (full batch and per-branch half batch) for each
BN the kernel supports, 20 back-to-back launches inside a CUDA graph (no host launch overhead, L2-warm operands as in
the step). Prints the table that choose_bn()'s cost model is fitted to."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cfgpp_b200 import _native as nv  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
IT = 20


def timed(fn):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(IT):
            fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * IT) * 1e3  # us


def linear_case(M, N, K, res, geglu=False):
    a = torch.randn(M, K, generator=g).half().to(dev)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).half().to(dev)
    bias = torch.randn(N, generator=g).half().to(dev)
    add = torch.randn(M, N, generator=g).half().to(dev) if res else None
    return lambda bn: (lambda: nv.op_linear(a, w, bias, add, geglu=geglu, force_bn=bn))


def conv_case(B, H, Cin, Cout):
    x = torch.randn(B, H, H, Cin, generator=g).half().to(dev)
    w = (torch.randn(Cout, 9 * Cin, generator=g) * (9 * Cin) ** -0.5).half().to(dev)
    bias = torch.randn(Cout, generator=g).half().to(dev)
    return lambda bn: (lambda: nv.op_conv3x3(x, w, bias, force_bn=bn))


cases = []
for B in (4, 2):  # full CFG batch, and one branch of the split graph
    T2, T1, T0 = B * 1024, B * 4096, B * 16384
    cases += [
        (f"L2 to_qkv   {T2}x3840x1280", linear_case(T2, 3840, 1280, False), 2.0 * T2 * 3840 * 1280),
        (f"L2 to_out   {T2}x1280x1280+r", linear_case(T2, 1280, 1280, True), 2.0 * T2 * 1280 * 1280),
        (f"L2 ff.out   {T2}x1280x5120+r", linear_case(T2, 1280, 5120, True), 2.0 * T2 * 1280 * 5120),
        (f"L2 geglu    {T2}x10240x1280", linear_case(T2, 10240, 1280, False, True), 2.0 * T2 * 10240 * 1280),
        (f"L1 to_qkv   {T1}x1920x640", linear_case(T1, 1920, 640, False), 2.0 * T1 * 1920 * 640),
        (f"L1 to_out   {T1}x640x640+r", linear_case(T1, 640, 640, True), 2.0 * T1 * 640 * 640),
        (f"L1 ff.out   {T1}x640x2560+r", linear_case(T1, 640, 2560, True), 2.0 * T1 * 640 * 2560),
        (f"L1 geglu    {T1}x5120x640", linear_case(T1, 5120, 640, False, True), 2.0 * T1 * 5120 * 640),
        (f"L2 shortcut {T2}x1280x2560", linear_case(T2, 1280, 2560, False), 2.0 * T2 * 1280 * 2560),
        (f"L1 shortcut {T1}x640x1280", linear_case(T1, 640, 1280, False), 2.0 * T1 * 640 * 1280),
        (f"L0 shortcut {T0}x320x640", linear_case(T0, 320, 640, False), 2.0 * T0 * 320 * 640),
        (f"L2 conv 1280->1280 B{B}", conv_case(B, 32, 1280, 1280), 2.0 * T2 * 1280 * 9 * 1280),
        (f"L2 conv 2560->1280 B{B}", conv_case(B, 32, 2560, 1280), 2.0 * T2 * 1280 * 9 * 2560),
        (f"L1 conv 640->640 B{B}", conv_case(B, 64, 640, 640), 2.0 * T1 * 640 * 9 * 640),
        (f"L1 conv 1280->640 B{B}", conv_case(B, 64, 1280, 640), 2.0 * T1 * 640 * 9 * 1280),
        (f"L0 conv 320->320 B{B}", conv_case(B, 128, 320, 320), 2.0 * T0 * 320 * 9 * 320),
        (f"L0 conv 640->320 B{B}", conv_case(B, 128, 640, 320), 2.0 * T0 * 320 * 9 * 640),
    ]

print(f"{'case':34s} " + " ".join(f"{'bn' + str(b):>9s}" for b in (0, 64, 128, 160, 256)) + "   best   TF/s(best)")
for name, mk, flops in cases:
    row = []
    geglu = "geglu" in name
    for bn in (0, 64, 128, 160, 256):
        if geglu and bn not in (0, 256):
            row.append(float("nan")); continue
        try:
            row.append(timed(mk(bn)))
        except Exception:  # unsupported width for this N
            row.append(float("nan"))
    valid = [(t, b) for t, b in zip(row[1:], (64, 128, 160, 256)) if t == t]
    bt, bb = min(valid)
    print(f"{name:34s} " + " ".join(f"{t:9.1f}" for t in row) + f"   {bb:4d}   {flops / bt / 1e6:6.0f}"
          + ("   <-- heuristic off by %.0f%%" % ((row[0] / bt - 1) * 100) if row[0] > bt * 1.03 else ""), flush=True)

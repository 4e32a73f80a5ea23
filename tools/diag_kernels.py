"""GPU diagnostic sweep of the operator-level kernels against torch references (run under gpurun).

Prints one line per case and never aborts on a mismatch, so that a single GPU call yields the full picture.
    python tools/diag_kernels.py [gemm] [conv] [attn] [norm] ...
"""
from __future__ import annotations

import sys
import time
import traceback
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cfgpp_b200 import _native as nv  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
RESULTS = []


def report(name, got, ref, tol=2e-3):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    rel = err.max().item() / denom
    relL2 = (got - ref).norm().item() / (ref.norm().item() + 1e-12)
    ok = bool(torch.isfinite(got).all()) and relL2 < tol
    line = f"[{'OK ' if ok else 'BAD'}] {name}: relL2={relL2:.3e} maxerr/max={rel:.3e} max|ref|={denom:.3f}"
    print(line, flush=True)
    if not ok:
        # error structure: which 8-row / 8-col residues are wrong (swizzle / descriptor diagnosis)
        e2 = err.reshape(-1, err.shape[-1])
        bad = e2 > (0.02 * denom)
        print(f"      bad frac={bad.float().mean().item():.4f}  rows bad (first 16 of {e2.shape[0]}): "
              f"{bad.any(1)[:16].int().tolist()}  cols bad (first 32): {bad.any(0)[:32].int().tolist()}")
        rb = bad.any(1).reshape(-1).cpu()
        cb = bad.any(0).reshape(-1).cpu()
        ridx = torch.arange(rb.numel())
        print(f"      bad rows by (row%128)//32: {[int(rb[(ridx % 128) // 32 == i].sum()) for i in range(4)]}"
              f"  by row%8: {[int(rb[ridx % 8 == i].sum()) for i in range(8)]}")
        cidx = torch.arange(cb.numel())
        print(f"      bad cols by (col%64)//8: {[int(cb[(cidx % 64) // 8 == i].sum()) for i in range(8)]}"
              f"  by col//32 (first 8): {[int(cb[32*i:32*i+32].sum()) for i in range(min(8, cb.numel() // 32))]}")
        print(f"      got[0,:8]={got.reshape(-1, got.shape[-1])[0, :8].tolist()}")
        print(f"      ref[0,:8]={ref.reshape(-1, ref.shape[-1])[0, :8].tolist()}")
    RESULTS.append((name, ok))
    return ok


def guarded(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"[EXC] {name}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()
        RESULTS.append((name, False))


def ref_linear(a, w, bias, addend, rows_per_group):
    acc = a.float() @ w.float().t()
    if bias is not None:
        acc = acc + bias.float()
    t = acc.half()
    if addend is not None:
        ad = addend.float()
        if rows_per_group > 1:
            ad = ad.repeat_interleave(rows_per_group, dim=0)[: a.shape[0]]
        t = (t.float() + ad).half()
    return t


def diag_gemm():
    g = torch.Generator(device="cpu").manual_seed(0)

    def rnd(*s, scale=1.0):
        return (torch.randn(*s, generator=g) * scale).half().to(dev)

    cases = [
        # M, N, K, bias, addend(0 none / 1 full / >1 rows-per-group), force_bn
        (128, 64, 64, False, 0, 64),
        (128, 128, 64, False, 0, 128),
        (128, 128, 128, False, 0, 128),
        (256, 256, 256, True, 0, 256),
        (256, 320, 320, True, 1, 160),
        (308, 1280, 2048, False, 0, 0),
        (4096, 1280, 1280, True, 1, 0),
        (4096, 640, 640, True, 1, 0),
        (2048, 320, 960, True, 1024, 0),
        (1000, 200, 192, True, 1, 128),
        (16384, 1920, 640, False, 0, 0),
        (4096, 1280, 5120, True, 1, 0),
    ]
    for (M, N, K, hb, ha, bn) in cases:
        name = f"linear M={M} N={N} K={K} bias={hb} add={ha} bn={bn}"

        def run():
            a = rnd(M, K)
            w = rnd(N, K, scale=K ** -0.5)
            bias = rnd(N) if hb else None
            addend = None
            if ha == 1:
                addend = rnd(M, N)
            elif ha > 1:
                addend = rnd((M + ha - 1) // ha, N)
            out = nv.op_linear(a, w, bias, addend, ha if ha > 1 else 1, force_bn=bn)
            report(name, out, ref_linear(a, w, bias, addend, ha))

        guarded(name, run)

    # dual-source A (K concat)
    def run_dual():
        M, K1, K2, N = 1024, 640, 320, 320
        a1, a2 = rnd(M, K1), rnd(M, K2)
        w = rnd(N, K1 + K2, scale=(K1 + K2) ** -0.5)
        bias = rnd(N)
        out = nv.op_linear(a1, w, bias, None, 1, a2=a2)
        report("linear dual-source", out, ref_linear(torch.cat([a1, a2], 1), w, bias, None, 1))

    guarded("linear dual-source", run_dual)

    # GEGLU: weight rows interleaved per 256-row tile as 128 value + 128 gate
    def run_geglu():
        M, C = 512, 640
        inner = 4 * C
        a = rnd(M, C)
        w = rnd(2 * inner, C, scale=C ** -0.5)  # torch layout: rows [0,inner) value, [inner, 2 inner) gate
        b = rnd(2 * inner)
        idx = []
        for t in range(inner // 128):
            idx += list(range(t * 128, t * 128 + 128)) + list(range(inner + t * 128, inner + t * 128 + 128))
        idx = torch.tensor(idx, device=dev)
        out = nv.op_linear(a, w[idx].contiguous(), b[idx].contiguous(), geglu=True)
        h = (a.float() @ w.float().t() + b.float()).half()
        val, gate = h[:, :inner], h[:, inner:]
        ref = (val.float() * torch.nn.functional.gelu(gate.float()).half().float()).half()
        report("linear GEGLU", out, ref)

    guarded("linear GEGLU", run_geglu)


def diag_conv():
    g = torch.Generator(device="cpu").manual_seed(1)

    def rnd(*s, scale=1.0):
        return (torch.randn(*s, generator=g) * scale).half().to(dev)

    cases = [
        # B, H, W, Cin, Cout, bias, temb, residual, bn
        (1, 32, 32, 64, 64, False, False, False, 64),
        (2, 32, 32, 64, 128, True, False, False, 128),
        (1, 128, 128, 64, 64, False, False, False, 64),
        (2, 64, 64, 128, 128, True, True, False, 0),
        (4, 16, 16, 128, 256, True, False, True, 0),
        (2, 8, 8, 128, 128, True, True, False, 0),
        (1, 128, 128, 320, 320, True, True, False, 0),
        (2, 64, 64, 640, 320, True, False, True, 0),
        (4, 32, 32, 1280, 1280, True, True, False, 0),
    ]
    for (B, H, W, Cin, Cout, hb, ht, hr, bn) in cases:
        name = f"conv3x3 B={B} H={H} W={W} Cin={Cin} Cout={Cout} bias={hb} temb={ht} res={hr} bn={bn}"

        def run():
            x = rnd(B, Cin, H, W)
            w = rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5)
            bias = rnd(Cout) if hb else None
            x_nhwc = x.permute(0, 2, 3, 1).contiguous()
            w_packed = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
            addend, rpg = None, 1
            if ht:
                addend, rpg = rnd(B, Cout), H * W
            elif hr:
                addend = rnd(B * H * W, Cout)
            out = nv.op_conv3x3(x_nhwc, w_packed, bias, addend, rpg)
            ref = torch.nn.functional.conv2d(x.float(), w.float(), bias.float() if hb else None, padding=1)
            ref = ref.half().permute(0, 2, 3, 1).reshape(B * H * W, Cout)
            if ht:
                ref = (ref.float() + addend.float().repeat_interleave(H * W, 0)).half()
            elif hr:
                ref = (ref.float() + addend.float()).half()
            report(name, out.reshape(B * H * W, Cout), ref)

        guarded(name, run)


def bench_gemm():
    print("--- GEMM timing (CUDA events, 20 iters) ---", flush=True)
    g = torch.Generator(device="cpu").manual_seed(2)
    shapes = [(4096, 1280, 1280, 0), (4096, 3840, 1280, 0), (4096, 1280, 5120, 0), (16384, 640, 640, 0),
              (16384, 1920, 640, 0), (16384, 640, 2560, 0), (65536, 320, 320, 0), (8192, 8192, 8192, 256),
              (4096, 1280, 1280, 128), (4096, 1280, 1280, 160), (4096, 1280, 1280, 256)]
    for (M, N, K, bn) in shapes:
        try:
            a = (torch.randn(M, K, generator=g)).half().to(dev)
            w = (torch.randn(N, K, generator=g) * K ** -0.5).half().to(dev)
            for _ in range(3):
                nv.op_linear(a, w, force_bn=bn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                nv.op_linear(a, w, force_bn=bn)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            e0.record()
            for _ in range(20):
                torch.matmul(a, w.t())
            e1.record()
            torch.cuda.synchronize()
            ms_t = e0.elapsed_time(e1) / 20
            tf = 2.0 * M * N * K / ms / 1e9
            print(f"gemm M={M} N={N} K={K} bn={bn}: {ms*1e3:.1f} us  {tf:.0f} TFLOP/s   (torch.matmul {ms_t*1e3:.1f} us "
                  f"{2.0*M*N*K/ms_t/1e9:.0f} TFLOP/s)", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"[EXC] bench gemm {M},{N},{K}: {e}", flush=True)

    print("--- conv3x3 timing ---", flush=True)
    for (B, H, W, Cin, Cout) in [(4, 128, 128, 320, 320), (4, 64, 64, 640, 640), (4, 32, 32, 1280, 1280),
                                 (4, 32, 32, 2560, 1280)]:
        try:
            x = torch.randn(B, Cin, H, W, generator=g).half().to(dev)
            w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (9 * Cin) ** -0.5).half().to(dev)
            x_nhwc = x.permute(0, 2, 3, 1).contiguous()
            w_packed = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
            xcl = x.contiguous(memory_format=torch.channels_last)
            wcl = w.contiguous(memory_format=torch.channels_last)
            for _ in range(3):
                nv.op_conv3x3(x_nhwc, w_packed)
                torch.nn.functional.conv2d(x, w, padding=1)
                torch.nn.functional.conv2d(xcl, wcl, padding=1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            res = []
            for fn in (lambda: nv.op_conv3x3(x_nhwc, w_packed), lambda: torch.nn.functional.conv2d(x, w, padding=1),
                       lambda: torch.nn.functional.conv2d(xcl, wcl, padding=1)):
                e0.record()
                for _ in range(10):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res.append(e0.elapsed_time(e1) / 10)
            fl = 2.0 * B * H * W * Cout * Cin * 9
            print(f"conv B={B} {H}x{W} {Cin}->{Cout}: ours {res[0]*1e3:.1f} us {fl/res[0]/1e9:.0f} TF/s | cudnn nchw "
                  f"{res[1]*1e3:.1f} us {fl/res[1]/1e9:.0f} | cudnn nhwc {res[2]*1e3:.1f} us {fl/res[2]/1e9:.0f}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"[EXC] bench conv: {e}", flush=True)


if __name__ == "__main__":
    which = set(sys.argv[1:]) or {"gemm", "conv", "bench"}
    t0 = time.time()
    print(f"device: {torch.cuda.get_device_name(0)}  lib: {nv.lib_path()}", flush=True)
    if "gemm" in which:
        diag_gemm()
    if "conv" in which:
        diag_conv()
    if "bench" in which:
        bench_gemm()
    nbad = sum(1 for _, ok in RESULTS if not ok)
    print(f"=== {len(RESULTS) - nbad}/{len(RESULTS)} cases OK in {time.time() - t0:.1f}s ===", flush=True)

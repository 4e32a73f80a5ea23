"""Average board power / SM clock while looping one kernel class for ~1.5 s each (is the step power-capped?)."""
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cfgpp_b200 import _native as nv  # noqa: E402

dev = torch.device("cuda:0")


class Sampler:
    def __init__(self):
        self.lines = []
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=power.draw,clocks.sm,clocks_event_reasons.sw_power_cap",
                                   "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._r, daemon=True).start()

    def _r(self):
        for ln in self.p.stdout:
            self.lines.append((time.time(), ln.strip()))

    def window(self, t0, t1):
        pw, ck = [], []
        for t, ln in self.lines:
            if t0 + 0.3 <= t <= t1:
                f = [x.strip() for x in ln.split(",")]
                try:
                    pw.append(float(f[0])); ck.append(float(f[1]))
                except ValueError:
                    pass
        return (sum(pw) / max(len(pw), 1), sum(ck) / max(len(ck), 1), len(pw))


def loop(name, fn, seconds=1.5, flops=0.0):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); n = 0
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    t1 = time.time()
    ms = e0.elapsed_time(e1) / n
    pw, ck, k = S.window(t0, t1)
    tf = flops / ms / 1e9 if flops else 0
    print(f"{name:44s} {ms*1e3:9.1f} us/iter  {tf:7.0f} TF/s  power {pw:6.0f} W  sm clock {ck:5.0f} MHz  ({k} samples)"
          f"  energy/iter {pw*ms/1e3:8.4f} J  {pw*ms/1e3/max(flops,1)*1e12:6.3f} pJ/FLOP", flush=True)
    time.sleep(0.5)


S = Sampler()
time.sleep(1.0)
g = torch.Generator().manual_seed(0)


def mk(M, N, K, res=False, geglu=False):
    a = torch.randn(M, K, generator=g).half().to(dev)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).half().to(dev)
    bias = torch.randn(N, generator=g).half().to(dev)
    add = torch.randn(M, N, generator=g).half().to(dev) if res else None
    return lambda: nv.op_linear(a, w, bias, add, geglu=geglu)


loop("idle (sleep)", lambda: time.sleep(0.01), 1.0)
loop("gemm 8192^3", mk(8192, 8192, 8192), flops=2.0 * 8192 ** 3)
loop("gemm geglu 4096x10240x1280", mk(4096, 10240, 1280, geglu=True), flops=2.0 * 4096 * 10240 * 1280)
loop("gemm ff.out 4096x1280x5120 +res", mk(4096, 1280, 5120, True), flops=2.0 * 4096 * 1280 * 5120)
loop("gemm to_out 4096x1280x1280 +res", mk(4096, 1280, 1280, True), flops=2.0 * 4096 * 1280 * 1280)
for (M, N, K) in [(4096, 10240, 1280), (4096, 1280, 5120), (4096, 3840, 1280), (4096, 1280, 1280)]:
    am = torch.randn(M, K, generator=g).half().to(dev)
    wm = (torch.randn(N, K, generator=g) * K ** -0.5).half().to(dev)
    loop(f"torch.matmul {M}x{N}x{K} (cuBLAS)", lambda: torch.matmul(am, wm.t()), flops=2.0 * M * N * K)
    loop(f"ours plain   {M}x{N}x{K}", lambda: nv.op_linear(am, wm), flops=2.0 * M * N * K)
a8 = torch.randn(8192, 8192, generator=g).half().to(dev)
loop("torch.matmul 8192^3 (cuBLAS)", lambda: torch.matmul(a8, a8), flops=2.0 * 8192 ** 3)
qkv = (torch.randn(4, 1024, 3 * 1280, generator=g) * 1.2).half().to(dev)
loop("attention self N=1024 H=20 B=4", lambda: nv.op_attention(qkv[:, :, :1280], qkv[:, :, 1280:2560], qkv[:, :, 2560:], 20),
     flops=4.0 * 4 * 20 * 1024 * 1024 * 64)
qkv4 = (torch.randn(4, 4096, 3 * 640, generator=g) * 1.2).half().to(dev)
loop("attention self N=4096 H=10 B=4", lambda: nv.op_attention(qkv4[:, :, :640], qkv4[:, :, 640:1280], qkv4[:, :, 1280:], 10),
     flops=4.0 * 4 * 10 * 4096 * 4096 * 64)
x = torch.randn(4096, 1280, generator=g).half().to(dev)
gm, bt = torch.ones(1280).half().to(dev), torch.zeros(1280).half().to(dev)
loop("layernorm 4096x1280", lambda: nv.op_layernorm(x, gm, bt))
xc = torch.randn(4, 1280, 32, 32, generator=g).half().to(dev)
wc = (torch.randn(1280, 1280, 3, 3, generator=g) * 0.01).half().to(dev)
xn = xc.permute(0, 2, 3, 1).contiguous(); wp = wc.permute(0, 2, 3, 1).reshape(1280, -1).contiguous()
loop("conv3x3 4x32x32 1280->1280", lambda: nv.op_conv3x3(xn, wp), flops=2.0 * 4096 * 1280 * 1280 * 9)
S.p.terminate()

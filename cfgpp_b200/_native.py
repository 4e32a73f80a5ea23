"""ctypes binding of libcfgpp_b200.so (the C-ABI boundary declared in include/cfgpp_b200.h).

There is deliberately no fallback: if the library is absent or a call fails, this raises.
PyTorch is used only as the owner of device memory and streams (raw pointers cross the boundary).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_void_p
from pathlib import Path

import torch

# CFGPP_B200_LIB points at an alternative build of the same library (A/B experiments with tools/build_variant.sh)
_LIB_PATH = Path(os.environ.get("CFGPP_B200_LIB") or Path(__file__).resolve().parent / "lib" / "libcfgpp_b200.so")
_lib = None


class NativeError(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise NativeError(
                f"{_LIB_PATH} not found — build it with `python -m cfgpp_b200.build` "
                "(there is no CPU/eager fallback on the product path)")
        _lib = ctypes.CDLL(str(_LIB_PATH))
        _lib.cfgpp_last_error.restype = c_char_p
        _lib.cfgpp_version.restype = c_int
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = load().cfgpp_last_error()
        raise NativeError(f"cfgpp native call failed ({status}): {msg.decode() if msg else '?'}")


def ptr(t: torch.Tensor | None) -> c_void_p:
    if t is None:
        return c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "native ops take contiguous CUDA tensors"
    return c_void_p(t.data_ptr())


def stream_ptr() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


# ------------------------------------------------------------------------------------------------
# operator-level wrappers (one kernel launch each) — used by tests and micro-benchmarks
# ------------------------------------------------------------------------------------------------
def op_linear(a: torch.Tensor, w: torch.Tensor, bias=None, addend=None, add_rows_per_group: int = 1,
              a2: torch.Tensor | None = None, geglu: bool = False, force_bn: int = 0) -> torch.Tensor:
    """out = epilogue(cat([a, a2], -1) @ w.T); a [M,K1] fp16, w [N,K] fp16 (already packed for GEGLU)."""
    lib = load()
    M, K1 = a.shape
    K = K1 + (a2.shape[1] if a2 is not None else 0)
    N = w.shape[0]
    assert w.shape[1] == K and a.dtype == torch.float16 and w.dtype == torch.float16
    n_out = N // 2 if geglu else N
    out = torch.empty((M, n_out), dtype=torch.float16, device=a.device)
    check(lib.cfgpp_op_linear(ptr(a), c_int(a.stride(0)), ptr(a2), c_int(a2.stride(0) if a2 is not None else 0),
                              c_int(K1), ptr(w), c_int(M), c_int(N), c_int(K), ptr(bias), ptr(addend),
                              c_int(addend.stride(0) if addend is not None else 0), c_int(add_rows_per_group),
                              ptr(out), c_int(n_out), c_int(1 if geglu else 0), c_int(force_bn), stream_ptr()))
    return out


def op_conv3x3(x_nhwc: torch.Tensor, w_packed: torch.Tensor, bias=None, addend=None,
               add_rows_per_group: int = 1, force_bn: int = 0) -> torch.Tensor:
    """x [B,H,W,Cin] fp16 NHWC, w_packed [Cout, 9*Cin] (tap-major), returns [B,H,W,Cout]."""
    lib = load()
    B, H, W, Cin = x_nhwc.shape
    Cout = w_packed.shape[0]
    assert w_packed.shape[1] == 9 * Cin
    out = torch.empty((B, H, W, Cout), dtype=torch.float16, device=x_nhwc.device)
    check(lib.cfgpp_op_conv3x3(ptr(x_nhwc), c_int(B), c_int(H), c_int(W), c_int(Cin), ptr(w_packed), c_int(Cout),
                               ptr(bias), ptr(addend), c_int(addend.stride(0) if addend is not None else 0),
                               c_int(add_rows_per_group), ptr(out), c_int(force_bn), stream_ptr()))
    return out


def op_conv3x3_s2(x_nhwc: torch.Tensor, w_packed: torch.Tensor, bias=None, pad: int = 1) -> torch.Tensor:
    """Downsample2D conv: x [B,H,W,Cin] fp16 NHWC (even H, W), w_packed [Cout, 9*Cin] (tap-major) -> [B,H/2,W/2,Cout].
    pad=1: symmetric zero padding (UNet); pad=0: one zero row / column after the image (AutoencoderKL encoder)."""
    lib = load()
    B, H, W, Cin = x_nhwc.shape
    Cout = w_packed.shape[0]
    out = torch.empty((B, H // 2, W // 2, Cout), dtype=torch.float16, device=x_nhwc.device)
    check(lib.cfgpp_op_conv3x3_s2(ptr(x_nhwc), c_int(B), c_int(H), c_int(W), c_int(Cin), ptr(w_packed), c_int(Cout),
                                  ptr(bias), c_int(pad), ptr(out), stream_ptr()))
    return out


def op_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, head_dim: int = 64) -> torch.Tensor:
    """q [B,Nq,H*P], k/v [B,Nkv,H*P] fp16 (views with a row stride are fine) -> [B,Nq,H*P]; P = head_dim rounded up
    to a multiple of 64, the padding columns of every head being zero."""
    lib = load()
    B, Nq, C = q.shape
    Nkv = k.shape[1]
    out = torch.empty((B, Nq, C), dtype=torch.float16, device=q.device)
    for t in (q, k, v):
        assert t.is_cuda and t.stride(2) == 1 and t.stride(0) == t.shape[1] * t.stride(1)
    check(lib.cfgpp_op_attention(c_void_p(q.data_ptr()), c_int(q.stride(1)), c_void_p(k.data_ptr()), c_int(k.stride(1)),
                                 c_void_p(v.data_ptr()), c_int(v.stride(1)), ptr(out), c_int(C), c_int(B),
                                 c_int(heads), c_int(Nq), c_int(Nkv), c_int(head_dim), stream_ptr()))
    return out


def op_groupnorm(x1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, silu: bool,
                 x2: torch.Tensor | None = None) -> torch.Tensor:
    """x1 [B,HW,C1] (+ x2 [B,HW,C2]) NHWC fp16 -> GroupNorm(32) over the channel concat, optional SiLU."""
    lib = load()
    B, HW, C1 = x1.shape
    C2 = x2.shape[2] if x2 is not None else 0
    out = torch.empty((B, HW, C1 + C2), dtype=torch.float16, device=x1.device)
    check(lib.cfgpp_op_groupnorm(ptr(x1), c_int(C1), ptr(x2), c_int(C2), c_int(B), c_int(HW), ptr(gamma), ptr(beta),
                                 c_float(eps), c_int(1 if silu else 0), ptr(out), stream_ptr()))
    return out


def op_layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    lib = load()
    M, C = x.shape
    out = torch.empty_like(x)
    check(lib.cfgpp_op_layernorm(ptr(x), c_int(M), c_int(C), ptr(gamma), ptr(beta), c_float(eps), ptr(out),
                                 stream_ptr()))
    return out


def op_cfgpp_step(eps_uc: torch.Tensor, eps_c: torch.Tensor, method: int, coef, z: torch.Tensor,
                  aux: torch.Tensor | None = None, want_z0t: bool = True, noise: torch.Tensor | None = None):
    """In-place CFG++ update of z (fp32 or fp16 state) from given eps; returns z0t (or None). `noise`: fp16 table
    [slots, *z.shape] of the ancestral samplers (slot = coef.c3)."""
    from ctypes import byref
    lib = load()
    z0t = torch.empty_like(z) if want_z0t else None
    code = 0 if z.dtype == torch.float16 else 1
    check(lib.cfgpp_op_cfgpp_step(ptr(eps_uc), ptr(eps_c), c_int(z.numel()), c_int(method), c_int(code), byref(coef),
                                  ptr(z), ptr(aux), ptr(z0t), ptr(noise), stream_ptr()))
    return z0t

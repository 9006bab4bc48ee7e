"""Helpers for the -m gpu tests: call the C ABI (include/hvit.h) with torch device tensors."""
import torch

import hvit_b200
from hvit_b200 import _lib


def lib():
    return _lib.load()


def stream():
    return _lib.current_stream_ptr()


def P(t):
    return None if t is None else t.data_ptr()


def sync():
    torch.cuda.synchronize()


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def gemm_16(a, w, scale=None, shift=None, act=0, residual=None, out_f32=False, out=None, ldc=None):
    """tcgen05 GEMM; the 16-bit operand type (bf16 / fp16) follows a.dtype."""
    M, K = a.shape
    N = w.shape[0]
    assert a.dtype == w.dtype and a.dtype in (torch.bfloat16, torch.float16)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32 if out_f32 else a.dtype, device=a.device)
    _lib.check(lib().hvit_gemm_16(P(a), a.stride(0), P(w), P(scale), P(shift), act, P(residual),
                                  0 if residual is None else residual.stride(0), P(out), ldc or out.stride(0),
                                  1 if out_f32 else 0, M, N, K, 1 if a.dtype == torch.float16 else 0, stream()),
               "hvit_gemm_16")
    return out


def gemm_f32(a, w, scale=None, shift=None, act=0, residual=None):
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    _lib.check(lib().hvit_gemm_f32(P(a), a.stride(0), P(w), P(scale), P(shift), act, P(residual),
                                   0 if residual is None else residual.stride(0), P(out), N, M, N, K, stream()),
               "hvit_gemm_f32")
    return out

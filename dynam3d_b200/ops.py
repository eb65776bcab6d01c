"""Thin Python wrappers over the C ABI (one function per exported kernel family).

All tensors are CUDA tensors owned by the caller; nothing here computes on the host or falls back to torch ops.
"""
import ctypes

import torch

from . import _lib as L


def gemm(a, w, out=None, bias=None, act=L.ACT_NONE, residual=None, out_dtype=None, simt=False):
    """out = epi(a @ w.T).  a [M,K], w [N,K] fp16/bf16 row-major; see d3d_gemm in include/dynam3d_b200.h."""
    assert a.is_cuda and w.is_cuda and a.dtype == w.dtype and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    n_out = N // 2 if act == L.ACT_SWIGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=out_dtype or a.dtype)
    assert out.stride(1) == 1 and out.shape[0] == M and out.shape[1] == n_out
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(1) == 1
    args = L.GemmArgs(
        L.ptr(a), a.stride(0), L.ptr(w), w.stride(0), L.ptr(out), out.stride(0), M, N, K,
        L.kind_of(a.dtype), L.kind_of(out.dtype), L.ptr(bias), int(act),
        L.ptr(residual), residual.stride(0) if residual is not None else 0)
    fn = L.lib().d3d_gemm_simt if simt else L.lib().d3d_gemm
    L.check(fn(ctypes.byref(args), L.stream_ptr()))
    return out

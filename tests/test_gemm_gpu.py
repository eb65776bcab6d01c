"""tcgen05 GEMM vs a plain PyTorch fp32 reference of the same op (floating-point kernel; tolerance stated per case)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias, act, residual):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    elif act == 3:
        y = torch.nn.functional.silu(y)
    elif act == 4:
        y = torch.nn.functional.silu(y[:, 0::2]) * y[:, 1::2]
    if residual is not None:
        y = y + residual
    return y


CASES = [
    # M, N, K, dtype, bias, act, residual, out_dtype
    (128, 128, 64, torch.float16, False, 0, False, torch.float32),
    (128, 128, 256, torch.bfloat16, False, 0, False, torch.float32),
    (577, 1024, 1024, torch.float16, True, 0, True, torch.float32),
    (6924, 3072, 1024, torch.float16, True, 0, False, torch.float16),
    (6924, 4096, 1024, torch.float16, True, 1, False, torch.float16),
    (600, 768, 3072, torch.float16, True, 2, False, torch.float16),
    (600, 16384, 3072, torch.bfloat16, False, 4, False, torch.bfloat16),
    (37, 2304, 768, torch.float16, True, 0, False, torch.float16),
    (300, 3072, 1544, torch.float16, True, 0, False, torch.float32),
    (8, 32064, 3072, torch.bfloat16, False, 0, False, torch.float32),
    (6912, 1024, 592, torch.float16, False, 0, False, torch.float32),
    (1, 768, 768, torch.float16, True, 3, False, torch.float32),
    (4800, 8192, 3072, torch.bfloat16, False, 0, True, torch.float32),
    # CTA-pair (cta_group::2) shapes: >= 148 pair tiles of 256x256; ragged M (not a multiple of 256 / 128) and N = 768
    (20000, 4096, 1024, torch.float16, True, 1, False, torch.float16),
    (19999, 1024, 4096, torch.float16, True, 0, True, torch.float32),
    (13849, 768, 1024, torch.float16, False, 0, False, torch.float16),
    (6001, 16384, 3072, torch.float16, False, 4, False, torch.float16),
    (6100, 9216, 3072, torch.bfloat16, False, 0, False, torch.bfloat16),
]


@pytest.mark.parametrize("M,N,K,dtype,use_bias,act,use_res,out_dtype", CASES)
def test_gemm_matches_torch(M, N, K, dtype, use_bias, act, use_res, out_dtype):
    from dynam3d_b200 import ops, _lib
    _lib.require_device(0)
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5)).to(dtype)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1 if use_bias else None
    n_out = N // 2 if act == 4 else N
    res = torch.randn(M, n_out, device="cuda", generator=g) if use_res else None
    out = ops.gemm(a, w, bias=bias, act=act, residual=res, out_dtype=out_dtype)
    torch.cuda.synchronize()
    ref = _ref(a, w, bias, act, res)
    # fp32 accumulate over K products of 16-bit operands: only summation order differs; 16-bit outputs add one rounding
    tol = 2e-3 if out_dtype == torch.float32 else (8e-3 if out_dtype == torch.float16 else 3e-2)
    err = (out.float() - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), f"max abs err {err}"
    # in-place residual (C aliases residual) is what the transformer blocks use
    if use_res and out_dtype == torch.float32:
        x = res.clone()
        ops.gemm(a, w, out=x, bias=bias, act=act, residual=x)
        torch.cuda.synchronize()
        assert torch.equal(x, out)


def test_gemm_simt_agrees():
    from dynam3d_b200 import ops
    a = torch.randn(200, 320, device="cuda").half()
    w = torch.randn(136, 320, device="cuda").half()
    b = torch.randn(136, device="cuda")
    y1 = ops.gemm(a, w, bias=b, act=2, out_dtype=torch.float32)
    y2 = ops.gemm(a, w, bias=b, act=2, out_dtype=torch.float32, simt=True)
    torch.cuda.synchronize()
    assert (y1 - y2).abs().max().item() < 2e-3


def test_gemm_pair_and_single_cta_agree_bitwise():
    """The CTA-pair kernel accumulates every output element over K in the same order as the single-CTA kernel."""
    from dynam3d_b200 import ops, _lib
    a = (torch.randn(16000, 1024, device="cuda") * 0.5).half()
    w = (torch.randn(4096, 1024, device="cuda") / 32).half()
    b = torch.randn(4096, device="cuda")
    try:
        _lib.lib().d3d_gemm_set_pair_mode(0)
        y0 = ops.gemm(a, w, bias=b, act=1)
        _lib.lib().d3d_gemm_set_pair_mode(1)
        y1 = ops.gemm(a, w, bias=b, act=1)
    finally:
        _lib.lib().d3d_gemm_set_pair_mode(-1)
    torch.cuda.synchronize()
    assert torch.equal(y0, y1)

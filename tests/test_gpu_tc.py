"""GPU: the tcgen05 GEMM (bf16 operands, fp32 accumulate) against an fp32 torch matmul of the same
bf16-rounded operands.  Both sides see identical inputs, so the tolerance only covers fp32 summation order."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _linear_bf16(x, w, b, relu=False, residual=None):
    from puzzlenet_b200 import _lib
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty(M, N, device=DEV, dtype=torch.float32)
    _lib.call("pz_linear_bf16", x.data_ptr(), x.stride(0), w.data_ptr(), b.data_ptr() if b is not None else None, M, N, K,
              1 if relu else 0, residual.data_ptr() if residual is not None else None,
              residual.stride(0) if residual is not None else 0, y.data_ptr(), N, _lib.stream_ptr())
    return y


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (256, 128, 128), (512, 256, 256), (1024, 384, 256), (2048, 1024, 1280),
                                   (256 * 150, 128, 64)])
def test_linear_bf16_matches_fp32_matmul(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    x = (torch.randn(M, K, generator=g)).to(torch.bfloat16).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    y = _linear_bf16(x, w, b)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + b
    err = (y - ref).abs().max().item()
    if err > 1e-3:
        bad = ((y - ref).abs() > 1e-3).nonzero()
        print(f"M={M} N={N} K={K} max err {err}; {bad.shape[0]} bad of {M * N}; first bad {bad[:8].tolist()}")
        print("y[0,:8]", y[0, :8].tolist(), "ref[0,:8]", ref[0, :8].tolist())
    assert err < 1e-3


def test_linear_bf16_relu_residual_and_strides():
    g = torch.Generator().manual_seed(3)
    M, N, K = 512, 256, 256
    xfull = torch.randn(M, 1280, generator=g).to(torch.bfloat16).to(DEV)
    x = xfull[:, 256:512]                                   # strided view: ldx = 1280
    w = (torch.randn(N, K, generator=g) / 16).to(torch.bfloat16).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    y = _linear_bf16(x, w, b, relu=True, residual=res)
    ref = res + torch.relu(x.float() @ w.float().t() + b)
    assert (y - ref).abs().max().item() < 1e-3

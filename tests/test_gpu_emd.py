"""GPU parity: approximate EMD vs the C oracle (known-answer pinned) and, where it was built in the build
container, vs the reference's own kernels compiled for sm_100a (oracle/_ref/libemd_ref.so)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import emd_oracle as eo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libemd_ref.so")


def test_known_answer():
    """PyTorchEMD/test_emd_loss.py:8-33: cost 0.71 per item and the closed-form gradients."""
    from puzzlenet_b200.emd import earth_mover_distance
    p1 = torch.tensor([[[1.7, -0.1, 0.1], [0.1, 1.2, 0.3]]]).repeat(3, 1, 1).to(DEV).requires_grad_()
    p2 = torch.tensor([[[0.3, 1.8, 0.2], [1.2, -0.2, 0.3]]]).repeat(3, 1, 1).to(DEV).requires_grad_()
    d = earth_mover_distance(p1, p2, transpose=False)
    np.testing.assert_allclose(d.detach().cpu().numpy(), [0.71] * 3, rtol=1e-5)
    (d[0] / 2 + d[1] * 2 + d[2] / 3).backward()
    q1 = p1.detach().clone().requires_grad_()
    q2 = p2.detach().clone().requires_grad_()
    gt = sum(w * (((q1[i, 0] - q2[i, 1]) ** 2).sum() + ((q1[i, 1] - q2[i, 0]) ** 2).sum())
             for i, w in enumerate((0.5, 2.0, 1 / 3)))
    gt.backward()
    np.testing.assert_allclose(p1.grad.cpu().numpy(), q1.grad.cpu().numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(p2.grad.cpu().numpy(), q2.grad.cpu().numpy(), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("b,n,m", [(3, 64, 64), (2, 128, 128), (2, 100, 40), (2, 48, 96), (2, 1024, 1024), (1, 1500, 700)])
def test_vs_c_oracle(b, n, m):
    from puzzlenet_b200 import emd_cuda
    rng = np.random.default_rng(n + m)
    x1 = rng.standard_normal((b, n, 3)).astype(np.float32) * 0.5
    x2 = rng.standard_normal((b, m, 3)).astype(np.float32) * 0.5
    t1, t2 = torch.from_numpy(x1).to(DEV), torch.from_numpy(x2).to(DEV)
    match = emd_cuda.approxmatch_forward(t1, t2)
    assert match.shape == (b, m, n)
    cost = emd_cuda.matchcost_forward(t1, t2, match)
    ref_match = eo.approxmatch(x1, x2)
    ref_cost = eo.matchcost(x1, x2, ref_match)
    # __expf (GPU) vs expf (C oracle): 1e-4 relative on the cost, 1e-3 on match/gradients
    np.testing.assert_allclose(cost.cpu().numpy(), ref_cost, rtol=1e-4)
    assert np.abs(match.cpu().numpy() - ref_match).max() < 1e-3
    gc = torch.linspace(0.5, 1.5, b, device=DEV)
    g1, g2 = emd_cuda.matchcost_backward(gc, t1, t2, match)
    r1, r2 = eo.matchcost_grad(gc.cpu().numpy(), x1, x2, match.cpu().numpy())
    np.testing.assert_allclose(g1.cpu().numpy(), r1, rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(g2.cpu().numpy(), r2, rtol=1e-3, atol=1e-4)


def test_transpose_default_and_2d_inputs():
    from puzzlenet_b200.emd import earth_mover_distance
    a = torch.randn(2, 64, 3, device=DEV)
    b = torch.randn(2, 64, 3, device=DEV)
    c1 = earth_mover_distance(a.transpose(1, 2), b.transpose(1, 2))          # default transpose=True: (b,3,n)
    c2 = earth_mover_distance(a, b, transpose=False)
    assert torch.equal(c1, c2)
    assert earth_mover_distance(a[0], b[0], transpose=False).shape == (1,)
    with pytest.raises(AssertionError):
        earth_mover_distance(a.cpu(), b.cpu(), transpose=False)              # emd.py:10


@pytest.mark.skipif(not os.path.isfile(REF_SO), reason="oracle/_ref/libemd_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("b,n,m", [(4, 1024, 1024), (3, 128, 128), (2, 96, 200)])
def test_vs_reference_kernels_on_gpu(b, n, m):
    """Same GPU, same __expf: the reference's own kernels (unmodified, sm_100a build) vs ours."""
    from puzzlenet_b200 import emd_cuda
    ref = ctypes.CDLL(REF_SO)
    g = torch.Generator().manual_seed(n)
    x1 = (torch.randn(b, n, 3, generator=g) * 0.5).to(DEV)
    x2 = (torch.randn(b, m, 3, generator=g) * 0.5).to(DEV)
    rmatch = torch.empty(b, m, n, device=DEV)
    temp = torch.empty(32 * (n + m) * 2, device=DEV)
    rcost = torch.empty(b, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    vp = ctypes.c_void_p
    assert ref.emd_ref_approxmatch(b, n, m, vp(x1.data_ptr()), vp(x2.data_ptr()), vp(rmatch.data_ptr()), vp(temp.data_ptr()), vp(st)) == 0
    assert ref.emd_ref_matchcost(b, n, m, vp(x1.data_ptr()), vp(x2.data_ptr()), vp(rmatch.data_ptr()), vp(rcost.data_ptr()), vp(st)) == 0
    match = emd_cuda.approxmatch_forward(x1, x2)
    cost = emd_cuda.matchcost_forward(x1, x2, match)
    torch.cuda.synchronize()
    np.testing.assert_allclose(cost.cpu().numpy(), rcost.cpu().numpy(), rtol=1e-4)
    assert (match - rmatch).abs().max().item() < 1e-4
    gc = torch.ones(b, device=DEV)
    g1, g2 = emd_cuda.matchcost_backward(gc, x1, x2, rmatch)
    r1, r2 = torch.empty(b, n, 3, device=DEV), torch.empty(b, m, 3, device=DEV)
    assert ref.emd_ref_matchcost_grad(b, n, m, vp(gc.data_ptr()), vp(x1.data_ptr()), vp(x2.data_ptr()), vp(rmatch.data_ptr()),
                                      vp(r1.data_ptr()), vp(r2.data_ptr()), vp(st)) == 0
    torch.cuda.synchronize()
    np.testing.assert_allclose(g1.cpu().numpy(), r1.cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(g2.cpu().numpy(), r2.cpu().numpy(), rtol=1e-4, atol=1e-5)

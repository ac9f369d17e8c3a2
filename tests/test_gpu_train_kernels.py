"""GPU parity of the individual training kernels (csrc/train.cu) against torch-CPU autograd on small, odd-shaped
problems -- the end-to-end gradient tests (test_gpu_training.py) only exercise the shapes of the model."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _st():
    return torch.cuda.current_stream().cuda_stream


_KEEP = []


def _d(t):
    """device pointer of a copy of t that stays alive (a temporary's memory would be handed to the next temporary)"""
    c = t.detach().to(DEV).contiguous()
    _KEEP.append(c)
    return c.data_ptr()


@pytest.fixture(scope="module")
def lib():
    from puzzlenet_b200 import _lib
    return _lib


@pytest.mark.parametrize("B,P,C", [(3, 7, 5), (4, 33, 64), (1, 16, 8)])
def test_bn_point_train_forward_backward(lib, B, P, C):
    g = torch.Generator().manual_seed(B * 100 + P)
    x = torch.randn(B, P, C, generator=g, requires_grad=True)
    gamma = (torch.rand(P, generator=g) + 0.5).requires_grad_()
    beta = torch.randn(P, generator=g).requires_grad_()
    rm, rv = torch.randn(P, generator=g) * 0.1, torch.rand(P, generator=g) + 0.5
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y_ref = torch.relu(F.batch_norm(x, rm_ref, rv_ref, gamma, beta, training=True, momentum=0.1, eps=1e-5))
    dy = torch.randn(B, P, C, generator=g)
    y_ref.backward(dy)
    xd, gd, bd, rmd, rvd = (t.detach().to(DEV).contiguous() for t in (x, gamma, beta, rm, rv))
    y, sm, si = torch.empty(B, P, C, device=DEV), torch.empty(P, device=DEV), torch.empty(P, device=DEV)
    lib.call("pz_bn_point_train_forward", xd.data_ptr(), B, P, C, gd.data_ptr(), bd.data_ptr(), rmd.data_ptr(),
             rvd.data_ptr(), 0.1, 1e-5, 1, y.data_ptr(), sm.data_ptr(), si.data_ptr(), _st())
    np.testing.assert_allclose(y.cpu().numpy(), y_ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    if B * C > 1:
        np.testing.assert_allclose(rmd.cpu().numpy(), rm_ref.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(rvd.cpu().numpy(), rv_ref.numpy(), rtol=1e-5, atol=1e-6)
    dx, dg, db = torch.empty_like(xd), torch.full((P,), 7.0, device=DEV), torch.full((P,), 7.0, device=DEV)
    lib.call("pz_bn_point_train_backward", xd.data_ptr(), y.data_ptr(), _d(dy), B, P, C, gd.data_ptr(),
             sm.data_ptr(), si.data_ptr(), 1, 0, dx.data_ptr(), dg.data_ptr(), db.data_ptr(), _st())
    np.testing.assert_allclose(dx.cpu().numpy(), x.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(dg.cpu().numpy(), gamma.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(db.cpu().numpy(), beta.grad.numpy(), rtol=1e-4, atol=1e-5)
    lib.call("pz_bn_point_train_backward", xd.data_ptr(), y.data_ptr(), _d(dy), B, P, C, gd.data_ptr(),
             sm.data_ptr(), si.data_ptr(), 1, 1, dx.data_ptr(), dg.data_ptr(), db.data_ptr(), _st())       # accumulate
    np.testing.assert_allclose(dg.cpu().numpy(), 2 * gamma.grad.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("G,K,C", [(5, 32, 128), (3, 7, 10), (2, 256, 1024), (4, 32, 6), (3, 5, 260)])
def test_maxpool_forward_backward(lib, G, K, C):
    g = torch.Generator().manual_seed(G + K + C)
    x = torch.relu(torch.randn(G, K, C, generator=g)).requires_grad_()       # ReLU outputs, as in the model
    y_ref, arg_ref = x.max(dim=1)
    dy = torch.randn(G, C, generator=g)
    y_ref.backward(dy)
    xd = x.detach().to(DEV)
    y, arg = torch.empty(G, C, device=DEV), torch.empty(G, C, device=DEV, dtype=torch.int32)
    lib.call("pz_maxpool_forward", xd.data_ptr(), G, K, C, y.data_ptr(), arg.data_ptr(), _st())
    np.testing.assert_array_equal(y.cpu().numpy(), y_ref.detach().numpy())
    picked = torch.gather(x.detach(), 1, arg.cpu().long().unsqueeze(1)).squeeze(1)
    np.testing.assert_array_equal(picked.numpy(), y_ref.detach().numpy())     # arg points at a maximiser (ties: any)
    dx = torch.full((G, K, C), 9.0, device=DEV)
    lib.call("pz_maxpool_backward", _d(dy), y.data_ptr(), arg.data_ptr(), G, K, C, 0, dx.data_ptr(), _st())
    ref = torch.zeros(G, K, C).scatter_(1, arg.cpu().long().unsqueeze(1), dy.unsqueeze(1))
    np.testing.assert_array_equal(dx.cpu().numpy(), ref.numpy())
    lib.call("pz_maxpool_backward", _d(dy), y.data_ptr(), arg.data_ptr(), G, K, C, 1, dx.data_ptr(), _st())
    gate = (y_ref.detach() > 0).float()
    np.testing.assert_array_equal(dx.cpu().numpy(), (ref * gate.unsqueeze(1)).numpy())


def test_scatter_add_rows_and_group_kernels(lib):
    g = torch.Generator().manual_seed(4)
    clouds, N, S, K, C, ld, c0 = 3, 20, 6, 5, 12, 17, 3
    M = clouds * S * K
    src = torch.randn(M, ld, generator=g)
    idx = torch.randint(0, N, (M,), generator=g)
    dst = torch.randn(clouds * N, C, generator=g)
    ref = dst.clone()
    for m in range(M):
        ref[(m // (S * K)) * N + idx[m]] += src[m, c0:c0 + C]
    d = dst.to(DEV)
    lib.call("pz_scatter_add_rows", _d(src), ld, c0, C, _d(idx), M, S * K, N, d.data_ptr(), C, _st())
    np.testing.assert_allclose(d.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)
    # gather_sub_relu / group_scatter_grad
    P, Q = torch.randn(clouds * N, C, generator=g), torch.randn(clouds * S, C, generator=g)
    out = torch.empty(M, C, device=DEV)
    lib.call("pz_gather_sub_relu", _d(P), _d(Q), _d(idx), clouds * S, K, S, N, C,
             out.data_ptr(), _st())
    rows = (torch.arange(M) // (S * K)) * N + idx
    ref_out = torch.relu(P[rows] - Q[torch.arange(M) // K])
    np.testing.assert_allclose(out.cpu().numpy(), ref_out.numpy(), rtol=1e-6, atol=1e-6)
    dpre = torch.randn(M, C, generator=g) * (ref_out > 0)
    dP, dQ = torch.zeros(clouds * N, C, device=DEV), torch.empty(clouds * S, C, device=DEV)
    lib.call("pz_group_scatter_grad", _d(dpre), _d(idx), clouds * S, K, S, N, C, dP.data_ptr(),
             dQ.data_ptr(), _st())
    ref_dP = torch.zeros(clouds * N, C).index_add_(0, rows, dpre)
    ref_dQ = -dpre.view(clouds * S, K, C).sum(1)
    np.testing.assert_allclose(dP.cpu().numpy(), ref_dP.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dQ.cpu().numpy(), ref_dQ.numpy(), rtol=1e-5, atol=1e-5)
    # broadcast_rows / group_sum / colsum / axpby / relu_gate / bias_act
    srcb = torch.randn(4, 9, generator=g)
    dstb = torch.zeros(4 * 5, 13, device=DEV)
    lib.call("pz_broadcast_rows", _d(srcb), 4, 5, 9, dstb.data_ptr(), 13, _st())
    np.testing.assert_array_equal(dstb[:, :9].cpu().numpy(), srcb.repeat_interleave(5, 0).numpy())
    gs = torch.empty(4, 9, device=DEV)
    lib.call("pz_group_sum", dstb.data_ptr(), 13, 4, 5, 9, gs.data_ptr(), _st())
    np.testing.assert_allclose(gs.cpu().numpy(), 5 * srcb.numpy(), rtol=1e-6)
    big = torch.randn(5000, 37, generator=g)
    cs = torch.ones(37, device=DEV)
    lib.call("pz_colsum", _d(big), 37, 5000, 37, 2.0, cs.data_ptr(), _st())
    np.testing.assert_allclose(cs.cpu().numpy(), 2.0 + big.double().sum(0).numpy(), rtol=1e-4, atol=1e-3)
    xa, ya = torch.randn(11, 7, generator=g), torch.randn(11, 7, generator=g)
    oa = torch.empty(11, 9, device=DEV)
    lib.call("pz_axpby", 11, 7, 2.0, _d(xa), 7, -0.5, _d(ya), 7, oa.data_ptr(), 9, _st())
    np.testing.assert_allclose(oa[:, :7].cpu().numpy(), (2 * xa - 0.5 * ya).numpy(), rtol=1e-6, atol=1e-6)
    og = torch.empty(11, 7, device=DEV)
    lib.call("pz_relu_gate", 11, 7, _d(xa), 7, _d(ya), 7, og.data_ptr(), 7, _st())
    np.testing.assert_array_equal(og.cpu().numpy(), torch.where(ya > 0, xa, torch.zeros(())).numpy())
    xb = xa.to(DEV).clone()
    bias = torch.randn(7, generator=g)
    lib.call("pz_bias_act", 11, 7, xb.data_ptr(), 7, _d(bias), 1, _st())
    np.testing.assert_allclose(xb.cpu().numpy(), torch.relu(xa + bias).numpy(), rtol=1e-6, atol=1e-6)


def test_softmax_cross_entropy_pose_grad_adam(lib):
    g = torch.Generator().manual_seed(6)
    # softmax backward
    S = (torch.randn(10, 33, generator=g) * 3).requires_grad_()
    A = torch.softmax(S * 0.125, -1)
    dA = torch.randn(10, 33, generator=g)
    A.backward(dA)
    dS = torch.empty(10, 33, device=DEV)
    lib.call("pz_softmax_backward", _d(A), _d(dA), 10, 33, 0.125, dS.data_ptr(), _st())
    np.testing.assert_allclose(dS.cpu().numpy(), S.grad.numpy(), rtol=1e-4, atol=1e-6)
    # cross entropy, both layouts
    B, N = 3, 50
    logits = torch.randn(B, 2, N, generator=g).requires_grad_()
    target = (torch.rand(B, N, generator=g) < 0.3).float()
    loss = F.cross_entropy(logits, target.long())
    loss.backward()
    for point_major in (0, 1):
        ld = logits.detach().permute(0, 2, 1).contiguous() if point_major else logits.detach()
        out, dl = torch.zeros(1, device=DEV), torch.empty_like(ld, device=DEV)
        lib.call("pz_cross_entropy", _d(ld), _d(target), B, N, point_major, 1.0, out.data_ptr(),
                 dl.data_ptr(), _st())
        np.testing.assert_allclose(out.item(), loss.item(), rtol=1e-5)
        got = dl.cpu().permute(0, 2, 1) if point_major else dl.cpu()
        np.testing.assert_allclose(got.numpy(), logits.grad.numpy(), rtol=1e-4, atol=1e-7)
    # pose gradient through se3.exp + transform + comp, incl. the small-angle (Taylor) branch and omega = 0
    from oracle import puzzle_oracle as po
    tw = torch.randn(4, 6, generator=g) * 0.4
    tw[1, :3] *= 1e-3
    tw[2, :3] = 0
    tw.requires_grad_()
    pts = torch.randn(4, 30, 3, generator=g)
    dq = torch.randn(4, 30, 3, generator=g)
    igt = po.se3_exp(torch.randn(4, 6, generator=g) * 0.3)
    mat = po.se3_exp(tw)
    q = po.se3_transform(mat, pts.permute(0, 2, 1)).permute(0, 2, 1)
    ((q * dq).sum() + 0.7 * po.comp(mat, igt)).backward()
    d6 = torch.empty(4, 6, device=DEV)
    lib.call("pz_pose_grad", _d(tw), _d(pts), _d(dq), 30,
             _d(igt), ctypes.c_float(0.7), 4, 0.0, d6.data_ptr(), _st())
    np.testing.assert_allclose(d6.cpu().numpy(), tw.grad.numpy(), rtol=2e-4, atol=2e-5)
    # chamfer gradient is covered in test_gpu_losses.py; Adam in test_gpu_training.py

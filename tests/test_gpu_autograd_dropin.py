"""The drop-in ``index_points`` / ``sample_and_group`` carry the gradients the reference's gather / subtract / cat
sequence has (pointnet_util.py:39-50, :115-130): with ``dropin.install(model=False)`` the reference's own network
trains through them, and mlp1/mlp2/bn1/bn2 (stage 1) and mlp3/mlp4 (stage 2) receive gradient ONLY through these
functions (model5_b.py:449-461).  Checked against torch-CPU autograd over the oracle's restatement."""
import pytest
import torch

from oracle import puzzle_oracle as po
from puzzlenet_b200 import pointnet_util as pu

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return ((a.detach().cpu() - b.detach()).abs().max() / b.detach().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("returnfps", [False, True])
def test_sample_and_group_backward_matches_oracle_autograd(returnfps):
    g = torch.Generator().manual_seed(3)
    B, N, D, S, K = 3, 200, 16, 24, 8
    xyz = torch.rand(B, N, 3, generator=g) - 0.5
    feat_in = torch.randn(B, N, 12, generator=g)
    pre = torch.nn.Linear(12, D)
    post = torch.nn.Linear(3 + D, 5)
    w_out = torch.randn(B, S, K, 5, generator=g)
    w_xyz = torch.randn(B, S, 3, generator=g)

    def run(dev, sg, **kw):
        pre_d, post_d = [torch.nn.Linear(m.in_features, m.out_features).to(dev) for m in (pre, post)]
        pre_d.load_state_dict(pre.state_dict()); post_d.load_state_dict(post.state_dict())
        x = xyz.to(dev).clone().requires_grad_(True)
        f = feat_in.to(dev).clone().requires_grad_(True)
        torch.manual_seed(17)                                       # the FPS start draw (CPU generator)
        r = sg(S, 0, K, x, pre_d(f), returnfps, True, **kw)
        loss = (torch.relu(post_d(r[1])) * w_out.to(dev)).sum() + (r[0] * w_xyz.to(dev)).sum()
        if returnfps:
            loss = loss + (r[2] ** 2).sum()
        loss.backward()
        return dict(loss=loss.detach(), x=x.grad, f=f.grad, pre_w=pre_d.weight.grad, pre_b=pre_d.bias.grad,
                    post_w=post_d.weight.grad)

    got = run(DEV, pu.sample_and_group)
    ref = run("cpu", lambda S_, r_, K_, x, p, rf, knn: po.sample_and_group(S_, r_, K_, x, p, returnfps=rf, knn=knn))
    errs = {k: _rel(got[k], ref[k]) for k in ref}
    print("sample_and_group autograd vs oracle:", errs)
    assert max(errs.values()) < 1e-4, errs


def test_index_points_backward_and_dtype_passthrough():
    g = torch.Generator().manual_seed(5)
    pts = torch.randn(2, 50, 7, generator=g)
    idx = torch.randint(0, 50, (2, 9, 4), generator=g)
    w = torch.randn(2, 9, 4, 7, generator=g)
    a = pts.to(DEV).requires_grad_(True)
    (pu.index_points(a, idx.to(DEV)) * w.to(DEV)).sum().backward()
    b = pts.clone().requires_grad_(True)
    (po.index_points(b, idx) * w).sum().backward()
    assert _rel(a.grad, b.grad) < 1e-6
    # non-float payloads (no gradient) still gather, and no graph is recorded under no_grad
    u8 = torch.randint(0, 255, (2, 50, 3), generator=g, dtype=torch.uint8)
    assert torch.equal(pu.index_points(u8.to(DEV), idx.to(DEV)).cpu(), po.index_points(u8, idx))
    with torch.no_grad():
        assert pu.index_points(a, idx.to(DEV)).grad_fn is None


def test_group_mlp_maxpool_refuses_autograd():
    g = torch.Generator().manual_seed(1)
    xyz = (torch.rand(1, 64, 3, generator=g) - 0.5).to(DEV)
    feat = torch.randn(1, 64, 64, generator=g).to(DEV).requires_grad_(True)
    new_xyz = xyz[:, :8].contiguous()
    idx = pu.knn_point(32, xyz, new_xyz)
    w1, b1 = torch.randn(128, 67, device=DEV), torch.randn(128, device=DEV)
    w2, b2 = torch.randn(128, 128, device=DEV), torch.randn(128, device=DEV)
    with pytest.raises(RuntimeError):
        pu.group_mlp_maxpool(xyz, feat, new_xyz, idx, w1, b1, w2, b2)
    with torch.no_grad():
        assert pu.group_mlp_maxpool(xyz, feat, new_xyz, idx, w1, b1, w2, b2).shape == (1, 8, 128)

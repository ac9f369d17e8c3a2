"""GPU parity for BASELINE configs[0] (the reference's test.py path: a batch of 4 piece pairs cut from 11 000-point
vase models): SURVEY.md §8(d) C1 stand-in data -- points on a vase-like surface of revolution, random plane cut,
FPS of both halves to 1024, random rigid motion -- built with the CPU oracle, pushed through ``test_step`` on the
GPU (fp32 and bf16 paths) and compared with the oracle's forward + epilogue."""
import numpy as np
import pytest
import torch

from oracle import puzzle_oracle as po

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def vase_model(seed, n=11000):
    rng = np.random.default_rng(seed)
    z = rng.uniform(-0.5, 0.5, n)
    th = rng.uniform(0, 2 * np.pi, n)
    r = 0.25 + 0.15 * np.sin(3 * z)
    pts = np.stack([r * np.cos(th), r * np.sin(th), z], 1) + rng.normal(0, 0.002, (n, 3))
    pts -= pts.mean(0)
    return (pts / np.linalg.norm(pts, axis=1).max()).astype(np.float32)


def c1_batch(B=4):
    items = []
    for s in range(B):
        pc = vase_model(2024 + s)
        np.random.seed(s)
        up, down = po.plane_split(pc)
        while up.shape[0] < 1024 or down.shape[0] < 1024:
            up, down = po.plane_split(pc)
        up, down = torch.from_numpy(po.dataset_fps(up, 1024)), torch.from_numpy(po.dataset_fps(down, 1024))
        fpcb, rpcb, fpc_idx, rpc_idx = po.get_boundary(down, up)
        torch.manual_seed(s)
        x = torch.randn(1, 6)
        x = x / x.norm(p=2, dim=1, keepdim=True) * 0.8
        g = po.se3_exp(x)
        mup = po.se3_transform(g, up.T[None])[0].T.contiguous()
        items.append((down, mup, g[0], up, fpcb, rpcb, fpc_idx, rpc_idx))
    return [torch.stack([it[k] for it in items]) for k in range(8)]


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_config1_test_step(cuda_model, state_dict, precision, tol):
    batch = c1_batch(4)
    starts = torch.stack([torch.randint(0, n, (4,), generator=torch.Generator().manual_seed(7 + i))
                          for i, n in enumerate((1024, 512, 1024, 512))])
    o = po.predict5(state_dict, batch[0], batch[1], starts=((starts[0], starts[1]), (starts[2], starts[3])))
    ref = po.test_step_scores(o["out"], o["de_fpcb"], o["de_mrpcb"], batch[0], batch[3], batch[4], batch[5], batch[6],
                              batch[7], batch[2])
    cuda_model.precision = precision
    dev_batch = [t.to(DEV) for t in batch]
    out, _, de_f, de_m = cuda_model.predict5(dev_batch, 4, starts=starts)
    rel = lambda a, b: ((a.cpu() - b).abs().max() / b.abs().max()).item()      # noqa: E731
    assert rel(out, o["out"]) < tol and rel(de_f, o["de_fpcb"]) < tol and rel(de_m, o["de_mrpcb"]) < tol
    from puzzlenet_b200 import losses
    s = losses.pair_score(out, de_f, de_m, dev_batch[0], dev_batch[3], dev_batch[4], dev_batch[5], dev_batch[6],
                          dev_batch[7], dev_batch[2]).cpu()
    if precision == "fp32":
        np.testing.assert_allclose(s[:, 0].numpy(), ref["r_iso"].numpy(), atol=0.01)            # degrees
        np.testing.assert_allclose(s[:, 1].numpy(), ref["t_iso"].numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(s[:, 8].numpy(), ref["cd_fpc"].numpy(), rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(s[:, 9].numpy(), ref["cd_rpc"].numpy(), rtol=1e-3, atol=1e-6)
        # IoU counts can differ by a point or two when two probabilities tie to the last bit at the 128th place
        assert (s[:, 4] - ref["inter_f"]).abs().max() <= 2 and (s[:, 6] - ref["inter_m"]).abs().max() <= 2
    cuda_model.precision = "fp32"
    scores = cuda_model.test_step(dev_batch, 0)
    assert scores.shape == (1, 10) and torch.isfinite(scores).all()

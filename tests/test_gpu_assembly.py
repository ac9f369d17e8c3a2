"""GPU: the assembly driver on the CUDA path -- batched pair scoring equals per-pair scoring, piece down-sampling
equals the oracle's dataset FPS, and the greedy loop runs end to end."""
import numpy as np
import pytest
import torch

from oracle import puzzle_oracle as po

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pieces(P, n, seed):
    g = torch.Generator().manual_seed(seed)
    return [(torch.rand(n + 37 * i, 3, generator=g) - 0.5).numpy() for i in range(P)]


def test_downsample_pieces_matches_dataset_fps():
    from puzzlenet_b200 import assembly
    pieces = _pieces(3, 3000, 1)
    starts = [5, 100, 2999]
    got = assembly.downsample_pieces(pieces, 1024, starts=starts, device=torch.device(DEV)).cpu().numpy()
    for i, p in enumerate(pieces):
        np.testing.assert_array_equal(got[i], po.dataset_fps(p, 1024, start=starts[i]))


def test_score_all_pairs_equals_single_pairs(cuda_model):
    from puzzlenet_b200 import assembly, losses
    from puzzlenet_b200.weights import make_batch
    cuda_model.precision = "fp32"
    clouds = torch.rand(5, 1024, 3, generator=torch.Generator().manual_seed(2)).to(DEV) - 0.5
    starts = torch.randint(0, 512, (4, 64), generator=torch.Generator().manual_seed(3))

    class FixedStartScorer(assembly.ModelScorer):         # fixed FPS starts so batching cannot change the result
        def __call__(self, fpc, mrpc):
            b = fpc.shape[0]
            out6, _, de_f, de_m = self.model.predict5(make_batch(fpc, mrpc), b, starts=starts[:, :b].contiguous())
            s = losses.pair_score(out6, de_f, de_m, fpc, mrpc)
            return torch.cat([out6, s[:, 10:11], torch.ones(b, 1, device=fpc.device)], dim=1)

    scorer = FixedStartScorer(cuda_model)
    pairs, rows = assembly.score_all_pairs(clouds, scorer, batch=4)
    assert pairs.shape == (10, 2) and rows.shape == (10, assembly.ROW_COLS)
    for r, (i, j) in enumerate(pairs.tolist()):
        b = r % 4                                            # position inside its batch decides the FPS start used
        one = cuda_model.predict5(make_batch(clouds[i:i + 1], clouds[j:j + 1]), 1, starts=starts[:, b:b + 1].contiguous())
        np.testing.assert_allclose(rows[r, :6].cpu().numpy(), one[0][0].cpu().numpy(), rtol=1e-5, atol=1e-6)
    assert torch.all(rows[:, 7] == 1) and torch.isfinite(rows).all()


@pytest.mark.parametrize("rescore", [False, True])
def test_assemble_end_to_end(cuda_model, rescore):
    from puzzlenet_b200 import assembly
    cuda_model.precision = "bf16"
    clouds = torch.rand(6, 1024, 3, generator=torch.Generator().manual_seed(4)).to(DEV) - 0.5
    torch.manual_seed(0)
    poses, merges = assembly.assemble(clouds, assembly.ModelScorer(cuda_model), batch=64, rescore=rescore)
    assert poses.shape == (6, 4, 4) and len(merges) == 5
    for k in range(6):                                       # every pose is a rigid motion
        R = poses[k][:3, :3]
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-5)
        np.testing.assert_allclose(poses[k][3], [0, 0, 0, 1], atol=1e-12)
    if rescore:      # pairs are i < j and the fixed piece keeps its frame: piece 0 ends up as the root
        np.testing.assert_allclose(poses[0], np.eye(4), atol=1e-12)
    else:            # Kruskal may re-express piece 0's component; exactly one piece (the root) keeps the identity
        assert sum(np.allclose(poses[k], np.eye(4), atol=1e-12) for k in range(6)) == 1
    cuda_model.precision = "fp32"

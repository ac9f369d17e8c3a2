"""GPU: the rows around the forward wired together the way train.py / test.py use them -- raw pieces ->
dataset.make_pair_batch -> training_step (pretraining branch, then the full branch) -> test_step."""
import types

import numpy as np
import pytest
import torch

from puzzlenet_b200.weights import synthetic_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_raw_pieces_to_training_and_evaluation():
    from puzzlenet_b200 import dataset as D
    from puzzlenet_b200.model5_b import TouchedRegraster
    from puzzlenet_b200.training import Trainer
    g = torch.Generator().manual_seed(3)
    pieces = [(torch.randn(5000 + 100 * i, 3, generator=g) * 0.3).numpy() for i in range(8)]
    np.random.seed(1)
    torch.manual_seed(2)
    batch = D.make_pair_batch(pieces, 0.8)
    model = TouchedRegraster(types.SimpleNamespace(dataset="vase", loss_mode=1, loss_sum=False, lr=1e-5,
                                                   pretrain_epochs=700))
    model.load_state_dict(synthetic_state_dict(0))
    model.to(DEV)
    model._trainer_state = Trainer(model, model.C, precision="tf32")
    model.current_epoch = 0                                    # < pretrain_epochs -> predict6 branch (model5_b.py:928)
    pre = [float(model.training_step(batch, 0)["loss"]) for _ in range(3)]
    assert all(np.isfinite(pre)) and pre[-1] < pre[0]
    enc2_before = model.Encoder2.out.weight.detach().clone()
    model.current_epoch = 700                                  # full branch
    full = [model.training_step(batch, 0) for _ in range(3)]
    assert all(np.isfinite(float(o["loss"])) for o in full)
    assert {"ce_f", "ce_m", "loss_mrpcb", "loss_emd", "lr"} <= set(full[0]["terms"])
    assert not torch.equal(model.Encoder2.out.weight, enc2_before)       # Encoder2 only trains in the full branch
    model.eval()
    model.precision = "bf16"
    scores = model.test_step(list(batch), 0)
    assert scores.shape == (1, 10) and torch.isfinite(scores).all()
    assert 0.0 <= scores[0, 6].item() <= 1.0 and 0.0 <= scores[0, 7].item() <= 1.0      # the two IoUs

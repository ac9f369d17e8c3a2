"""CPU: oracle/train_oracle.py (the restated training_step + autograd) against the frozen loss / gradient digests of
the UNMODIFIED reference training_step (tests/golden/reference_training.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import puzzle_oracle as po
from oracle import train_oracle as to
from tests.golden_inputs import FPS_SEED, training_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_training.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLDEN))


def test_training_step_loss_and_grads(gold, state_dict):
    sd = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in state_dict.items()}
    batch = training_inputs(2, po.se3_exp)
    bn_state = {}
    torch.manual_seed(FPS_SEED)
    r = to.training_loss(sd, batch, bn_state=bn_state)
    np.testing.assert_allclose(r["loss"].detach().numpy(), gold["loss"], rtol=1e-5)
    r["loss"].backward()
    seen = 0
    for key, ref in gold.items():
        if key.startswith("grad/"):
            g = sd[key[5:]].grad
            assert g is not None, key
            d = to.grad_digest(g)
            scale = max(np.sqrt(ref[1]), 1e-12)
            if key.endswith("mlpk.bias"):      # mathematically zero (a constant per softmax row): rounding noise
                qref = gold[key.replace("mlpk", "mlpq")][1]
                assert d[1] < 1e-6 * qref and ref[1] < 1e-6 * qref, key
                seen += 1
                continue
            np.testing.assert_allclose(d[1], ref[1], rtol=2e-3, err_msg=key)                 # sum of squares
            np.testing.assert_allclose(d[2:], ref[2:], rtol=0, atol=2e-3 * scale, err_msg=key)
            seen += 1
        elif key.startswith("buf/"):
            np.testing.assert_allclose(bn_state[key[4:]].numpy()[:16], ref, rtol=1e-5, atol=1e-6, err_msg=key)
    assert seen > 100
    # parameters the reference leaves without gradient (unused decoders, dt) have none here either
    for k, v in sd.items():
        if k.startswith(("fpc_decoder", "rpc_decoder", "dt")):
            assert v.grad is None


def test_pretraining_branch_loss_and_grads(gold, state_dict):
    """current_epoch < pretrain_epochs: predict6 (Encoder on both clouds) + pose losses only (model5_b.py:928-931,
    :1048-1050).  Encoder2 and the boundary heads receive no gradient."""
    sd = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in state_dict.items()}
    batch = training_inputs(2, po.se3_exp)
    bn_state = {}
    torch.manual_seed(FPS_SEED)
    r = to.training_loss(sd, batch, bn_state=bn_state, pretrain=True)
    np.testing.assert_allclose(r["loss"].detach().numpy(), gold["pre/loss"], rtol=1e-5)
    r["loss"].backward()
    seen = 0
    for key, ref in gold.items():
        if key.startswith("pre/grad/"):
            name = key[9:]
            g = sd[name].grad
            assert g is not None, key
            if name.endswith("mlpk.bias"):
                continue
            d = to.grad_digest(g)
            np.testing.assert_allclose(d[1], ref[1], rtol=2e-3, err_msg=key)
            np.testing.assert_allclose(d[2:], ref[2:], rtol=0, atol=2e-3 * max(np.sqrt(ref[1]), 1e-12), err_msg=key)
            seen += 1
        elif key.startswith("pre/buf/"):
            got = bn_state.get(key[8:], state_dict[key[8:]])          # Encoder2's statistics stay untouched
            np.testing.assert_allclose(got.numpy()[:16], ref, rtol=1e-5, atol=1e-6, err_msg=key)
    assert 30 < seen < 60
    assert all(sd[k].grad is None for k in sd if k.startswith(("Encoder2.", "MLP")) and sd[k].is_floating_point()
               and "running" not in k)

"""TEST INFRASTRUCTURE ONLY -- freeze the loss and gradients of the UNMODIFIED reference ``training_step``
(model5_b.py:912-1155) at B=2 on the CPU.  The reference's EMD is CUDA-only (PyTorchEMD/emd.py:10), so
``model5_b.earth_mover_distance`` is pointed at the C-oracle Function of oracle/train_oracle.py (same
forward/backward contract as emd.py:5-21); logging / visualisation hooks are no-ops.

    python oracle/make_golden_training.py      -> tests/golden/reference_training.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim, train_oracle  # noqa: E402
from puzzlenet_b200.weights import synthetic_pairs, synthetic_state_dict  # noqa: E402
from tests.golden_inputs import FPS_SEED, training_inputs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_training.npz")


def main():
    ns = ref_shim.load_reference()
    m5 = ns.model5_b
    m5.earth_mover_distance = train_oracle.earth_mover_distance
    model = m5.TouchedRegraster(ref_shim.reference_config())
    model.load_state_dict(synthetic_state_dict(0), strict=True)
    model.device = torch.device("cpu")
    model.current_epoch = 1000
    model.vis = lambda *a, **k: None
    model.vis_attention = lambda *a, **k: None
    model.logger = ref_shim._Anything()

    class _Sched:
        def get_last_lr(self):
            return [0.0]
    model.scheduler = _Sched()
    batch = training_inputs(2, ns.se3.exp)
    model.train()
    torch.manual_seed(FPS_SEED)
    loss = model.training_step(batch, 0)["loss"]
    loss.backward()
    out = {"loss": loss.detach().numpy()}
    for name, p in model.named_parameters():
        if p.grad is not None:
            out["grad/" + name] = train_oracle.grad_digest(p.grad)
    for name, b in model.named_buffers():
        if "running" in name:
            out["buf/" + name] = b.detach().numpy()[:16].copy()
    # ---- the pretraining branch (current_epoch < pretrain_epochs -> predict6, pose losses only)
    model2 = m5.TouchedRegraster(ref_shim.reference_config())
    model2.load_state_dict(synthetic_state_dict(0), strict=True)
    model2.device = torch.device("cpu")
    model2.C.pretrain_epochs = 700
    model2.current_epoch = 0
    model2.vis = lambda *a, **k: None
    model2.vis_attention = lambda *a, **k: None
    model2.logger = ref_shim._Anything()
    model2.scheduler = _Sched()
    model2.train()
    torch.manual_seed(FPS_SEED)
    loss2 = model2.training_step(batch, 0)["loss"]
    loss2.backward()
    out["pre/loss"] = loss2.detach().numpy()
    for name, p in model2.named_parameters():
        if p.grad is not None:
            out["pre/grad/" + name] = train_oracle.grad_digest(p.grad)
    for name, b in model2.named_buffers():
        if "running" in name:
            out["pre/buf/" + name] = b.detach().numpy()[:16].copy()
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, f"{os.path.getsize(GOLDEN) / 1024:.0f} KiB;", len(out), "arrays; loss", float(loss))


if __name__ == "__main__":
    main()

"""Parameter layout of ``TouchedRegraster`` and deterministic synthetic weights.

The reference ships no checkpoint, so benchmarks and parity
tests use seeded random parameters with exactly the reference's ``state_dict``
keys and shapes (model5_b.py:417-441 encoder, :84-90 attention layer,
:530-599 heads).  The values do **not** depend on the
reference's constructor order: every tensor is drawn from its own generator
keyed by (seed, position in ``PARAM_SPECS``), so the same dictionary can be
loaded into the reference model (``load_state_dict``), into the oracle and
into :class:`puzzlenet_b200.model5_b.TouchedRegraster`.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

NUM_POINTS = 1024


def _encoder_specs(prefix: str):
    specs = [
        (f"{prefix}.mlp1", 64, 3), (f"{prefix}.mlp2", 64, 64),
        (f"{prefix}.mlp3", 128, 67), (f"{prefix}.mlp4", 128, 128),
        (f"{prefix}.mlp5", 256, 131), (f"{prefix}.mlp6", 256, 256),
    ]
    for i in (1, 2, 3, 4):
        specs += [(f"{prefix}.atten{i}.mlpq", 64, 256), (f"{prefix}.atten{i}.mlpk", 64, 256),
                  (f"{prefix}.atten{i}.mlpv", 256, 256), (f"{prefix}.atten{i}.out", 256, 256)]
    specs.append((f"{prefix}.out", 1024, 1280))
    return specs


def linear_specs():
    """[(module path, out_features, in_features)] for every nn.Linear of the model."""
    specs = _encoder_specs("Encoder") + _encoder_specs("Encoder2")
    for dec in ("fpc_decoder", "rpc_decoder"):            # unused by predict5, kept for ckpt compat
        specs += [(f"{dec}.mlp1", 512, 512), (f"{dec}.mlp2", 256, 512), (f"{dec}.mlp3", 2, 256)]
    specs += [("tfMLP.0", 1024, 2048), ("tfMLP.2", 512, 1024), ("tfMLP.4", 512, 512),
              ("tfMLP.6", 256, 512), ("tfMLP.8", 6, 256)]
    for name in ("MLPLocalPreRpc", "MLPLocalPreFpc"):
        specs += [(f"{name}.0", 64, 64), (f"{name}.2", 64, 64), (f"{name}.4", 64, 64)]
    for name in ("MLPRpcb", "MLPFpcb"):
        specs += [(f"{name}.0", 64, 128), (f"{name}.2", 32, 64), (f"{name}.4", 2, 32)]
    return specs


def bn_names():
    return ["Encoder.bn1", "Encoder.bn2", "Encoder2.bn1", "Encoder2.bn2"]


def synthetic_state_dict(seed: int = 0, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Seeded parameters with non-trivial BatchNorm running statistics.

    Linear weights are U(-g/sqrt(fan_in), g/sqrt(fan_in)) with gain g = 2 (4 for the
    attention q/k projections) and biases U(-1/sqrt(fan_in), 1/sqrt(fan_in)): with
    nn.Linear's default gain of 1 the 20-layer stack shrinks every activation until the
    outputs are input-independent biases, which would make parity tests vacuous.
    BN gamma ~ U(0.5, 1.5), beta ~ N(0, 0.1), running_mean ~ N(0, 0.1),
    running_var ~ U(0.5, 1.5) so eval-mode BN is not the identity.
    """
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["dt"] = torch.full((1, 6), 1.0e-2, dtype=dtype)
    slot = 0

    def gen():
        nonlocal slot
        slot += 1
        return torch.Generator().manual_seed(seed * 100003 + slot)

    for name, fout, fin in linear_specs():
        bound = 1.0 / math.sqrt(fin)
        gain = 4.0 if name.endswith((".mlpq", ".mlpk")) else 2.0
        sd[f"{name}.weight"] = ((torch.rand(fout, fin, generator=gen()) * 2 - 1) * bound * gain).to(dtype)
        sd[f"{name}.bias"] = ((torch.rand(fout, generator=gen()) * 2 - 1) * bound).to(dtype)
    for name in bn_names():
        sd[f"{name}.weight"] = (torch.rand(NUM_POINTS, generator=gen()) + 0.5).to(dtype)
        sd[f"{name}.bias"] = (torch.randn(NUM_POINTS, generator=gen()) * 0.1).to(dtype)
        sd[f"{name}.running_mean"] = (torch.randn(NUM_POINTS, generator=gen()) * 0.1).to(dtype)
        sd[f"{name}.running_var"] = (torch.rand(NUM_POINTS, generator=gen()) + 0.5).to(dtype)
        sd[f"{name}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd


def synthetic_pairs(batch: int, seed: int = 64, n: int = NUM_POINTS):
    """C2 clouds of SURVEY.md §8(d): uniform cube, fpc and mrpc from one generator."""
    g = torch.Generator().manual_seed(seed)
    fpc = torch.rand(batch, n, 3, generator=g) - 0.5
    mrpc = torch.rand(batch, n, 3, generator=g) - 0.5
    return fpc, mrpc


def make_batch(fpc: torch.Tensor, mrpc: torch.Tensor):
    """The 8-tuple predict5 unpacks (model5_b.py:691-699); only [0] and [1] are computed on."""
    b, n = fpc.shape[0], fpc.shape[1]
    dev = fpc.device
    z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
    igt = torch.eye(4, device=dev).repeat(b, 1, 1)
    return [fpc, mrpc, igt, z(b, n, 3), z(b, 128, 3), z(b, 128, 3), z(b, n), z(b, n)]

"""CPU: host-side mirror of the reference interface -- parameter layout, RNG contract, loud failure
without CUDA, drop-in module registration."""
import sys
import types

import pytest
import torch

from puzzlenet_b200 import dropin, weights
from puzzlenet_b200.model5_b import PCTransformer_nonsort, TouchedRegraster


def _cfg():
    return types.SimpleNamespace(dataset="vase")


def test_state_dict_layout_matches_reference_spec(state_dict):
    model = TouchedRegraster(_cfg())
    own = model.state_dict()
    assert list(sorted(own.keys())) == list(sorted(state_dict.keys()))
    for k, v in state_dict.items():
        assert tuple(own[k].shape) == tuple(v.shape), k
    n_params = sum(p.numel() for p in model.parameters())
    assert n_params == 8_059_220
    model.load_state_dict(state_dict, strict=True)


def test_synthetic_weights_are_deterministic():
    a, b = weights.synthetic_state_dict(0), weights.synthetic_state_dict(0)
    assert all(torch.equal(a[k], b[k]) for k in a)
    c = weights.synthetic_state_dict(1)
    assert not torch.equal(a["tfMLP.0.weight"], c["tfMLP.0.weight"])


def test_cpu_tensors_fail_loudly(state_dict):
    """No CPU fallback: a CPU tensor must raise, not silently compute."""
    from puzzlenet_b200 import pointnet_util as pu
    x = torch.rand(1, 64, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        pu.farthest_point_sample(x, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        pu.square_distance(x, x)
    model = TouchedRegraster(_cfg()).eval()
    fpc, mrpc = weights.synthetic_pairs(1)
    with pytest.raises(RuntimeError, match="CUDA"):
        model.predict5(weights.make_batch(fpc, mrpc), 0)
    with pytest.raises(RuntimeError, match="CUDA"):          # the train-mode forward is CUDA-only as well
        model.predict5(weights.make_batch(fpc, mrpc), 0, training=True)
    enc = PCTransformer_nonsort(_cfg())
    with pytest.raises(NotImplementedError):          # train mode
        enc(torch.rand(1, 1024, 3))


def test_predict5_rng_contract_is_four_cpu_draws():
    """(1024,B), (512,B), (1024,B), (512,B) from the CPU default generator."""
    B = 3
    torch.manual_seed(99)
    expect = [torch.randint(0, n, (B,), dtype=torch.long) for n in (1024, 512, 1024, 512)]
    after = torch.rand(1)
    torch.manual_seed(99)
    got = [torch.randint(0, n, (B,), dtype=torch.long) for n in (1024, 512, 1024, 512)]
    assert all(torch.equal(a, b) for a, b in zip(expect, got))
    assert torch.equal(after, torch.rand(1))


def test_dropin_registers_reference_module_names():
    dropin.install()
    try:
        import pointnet_util as pu          # noqa: F401  (what model5_b.py:39 does)
        from PyTorchEMD.emd import earth_mover_distance   # model5_b.py:48
        import emd_cuda                     # PyTorchEMD/emd.py:2
        import model5_b
        assert pu.__name__ == "puzzlenet_b200.pointnet_util"
        assert callable(earth_mover_distance) and hasattr(emd_cuda, "approxmatch_forward")
        assert hasattr(model5_b, "TouchedRegraster")
        for fn in ("square_distance", "index_points", "farthest_point_sample", "query_ball_point",
                   "sample_and_group", "sample_and_group_all"):
            assert hasattr(pu, fn)
    finally:
        dropin.uninstall()
    assert "pointnet_util" not in sys.modules or not sys.modules["pointnet_util"].__name__.startswith("puzzlenet_b200")


def test_signatures_match_reference():
    import inspect
    from puzzlenet_b200 import emd, pointnet_util as pu
    assert list(inspect.signature(pu.sample_and_group).parameters) == \
        ["npoint", "radius", "nsample", "xyz", "points", "returnfps", "knn"]
    assert list(inspect.signature(pu.query_ball_point).parameters) == ["radius", "nsample", "xyz", "new_xyz"]
    assert list(inspect.signature(pu.farthest_point_sample).parameters) == ["xyz", "npoint"]
    assert list(inspect.signature(pu.square_distance).parameters) == ["src", "dst"]
    assert list(inspect.signature(pu.index_points).parameters) == ["points", "idx"]
    sig = inspect.signature(emd.earth_mover_distance)
    assert list(sig.parameters) == ["xyz1", "xyz2", "transpose"] and sig.parameters["transpose"].default is True
    p5 = inspect.signature(TouchedRegraster.predict5).parameters
    assert list(p5)[:5] == ["self", "batch", "batch_indic", "need", "training"]
    assert p5["need"].default is False and p5["training"].default is False


def test_flat_parameter_buffer_and_gradient_buckets():
    """training._Flat: live parameters become 256-byte aligned views of one buffer; the three all-reduce buckets
    (Encoder, Encoder2, heads) tile it exactly; the unused decoders and dt stay outside."""
    from puzzlenet_b200.training import _Flat
    model = TouchedRegraster(_cfg())
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    flat = _Flat(model)
    assert set(flat.segments) == {"Encoder", "Encoder2", "heads"}
    spans = sorted(flat.segments.values())
    assert spans[0][0] == 0 and spans[-1][1] == flat.n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    live = 0
    for n, p in model.named_parameters():
        assert torch.equal(p.detach(), before[n])                       # values survive the re-homing
        if n.startswith(("fpc_decoder", "rpc_decoder")) or n == "dt":
            assert id(p) not in flat.grad_of
            continue
        live += p.numel()
        assert flat.g(p).shape == p.shape
        assert (p.data_ptr() - flat.params.data_ptr()) % 256 == 0        # cudaMalloc bases are 256-byte aligned
        assert (flat.g(p).data_ptr() - flat.grads.data_ptr()) == (p.data_ptr() - flat.params.data_ptr())
        lo = (p.data_ptr() - flat.params.data_ptr()) // 4
        assert 0 <= lo and lo + p.numel() <= flat.n
    assert live == 7_270_218


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the driver's reference arm: the oracle port on the host cores) writes exactly one
    JSON line to stdout, with the keys of the bench contract; anything else a library prints goes to stderr."""
    import json
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0

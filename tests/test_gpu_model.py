"""GPU parity: encoder blocks and the pair-matching forward vs the CPU oracle / frozen reference outputs.
Tolerances are the ones north_star states: indices bit-exact; features and boundary logits within 1e-4
relative (fp32 path; relative to the tensor's max magnitude); rotation within 0.01 deg and translation
within 1e-4 of the reference."""
import numpy as np
import pytest
import torch

from oracle import puzzle_oracle as po
from puzzlenet_b200.weights import make_batch, synthetic_pairs
from tests.golden_inputs import FPS_SEED, golden_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REL_FP32 = 1e-4


def rel_err(got, ref):
    ref = torch.as_tensor(ref)
    got = torch.as_tensor(got).cpu().to(ref.dtype)
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def test_se3_exp(goldens):
    from puzzlenet_b200 import se3
    _, _, _, twist = golden_inputs()
    got = se3.exp(twist.to(DEV))
    np.testing.assert_allclose(got.cpu().numpy(), goldens["se3_exp"], rtol=0, atol=2e-6)


def test_linear_and_attention_blocks(state_dict):
    from puzzlenet_b200.model5_b import layerAttention, scaled_dot_production
    g = torch.Generator().manual_seed(1)
    q, k, v = torch.randn(3, 256, 64, generator=g), torch.randn(3, 256, 64, generator=g), torch.randn(3, 256, 256, generator=g)
    vals, attn = scaled_dot_production(q.to(DEV), k.to(DEV), v.to(DEV))
    rv, ra = po.scaled_dot_production(q, k, v)
    assert rel_err(vals, rv) < REL_FP32 and rel_err(attn, ra) < REL_FP32
    np.testing.assert_allclose(attn.sum(-1).cpu().numpy(), 1.0, atol=1e-5)
    layer = layerAttention(None, 256)
    sd = {kk[len("Encoder.atten3."):]: vv for kk, vv in state_dict.items() if kk.startswith("Encoder.atten3.")}
    layer.load_state_dict(sd)
    x = torch.randn(2, 256, 256, generator=g) * 0.5
    out, a = layer.to(DEV)(x.to(DEV))
    ro, ra = po.layer_attention(state_dict, "Encoder.atten3", x)
    assert rel_err(out, ro) < REL_FP32 and rel_err(a, ra) < REL_FP32


@pytest.mark.parametrize("B", [1, 3, 150])
def test_layer_attention_bf16_fused_kernel(state_dict, B):
    """The fused tcgen05 layer kernel (attention_layer_tc.cu) through ``layerAttention(precision='bf16')``: 2e-2 of the
    oracle's fp32 layer, and ~1e-3 of a torch emulation that rounds the same operands (x, weights, q, k, v, P, r) to
    bf16 -- the second bound is what catches a mis-placed tile or a dropped k-block.  B = 150 > 148 SMs."""
    from puzzlenet_b200.model5_b import layerAttention
    g = torch.Generator().manual_seed(5)
    layer = layerAttention(None, 256)
    pre = "Encoder2.atten2."
    sd = {kk[len(pre):]: vv for kk, vv in state_dict.items() if kk.startswith(pre)}
    layer.load_state_dict(sd)
    layer.precision = "bf16"
    x = torch.randn(B, 256, 256, generator=g) * 0.5
    out, a = layer.to(DEV)(x.to(DEV))
    torch.cuda.synchronize()
    ro, ra = po.layer_attention(state_dict, pre[:-1], x)
    assert rel_err(out, ro) < 2e-2
    assert rel_err(a, ra) < 0.25
    np.testing.assert_allclose(a.sum(-1).cpu().numpy(), 1.0, atol=1e-4)

    def bf(t):
        return t.to(torch.bfloat16).float()
    xb = bf(x)
    q = bf(xb @ bf(sd["mlpq.weight"]).t() + sd["mlpq.bias"])
    k = bf(xb @ bf(sd["mlpk.weight"]).t() + sd["mlpk.bias"])
    v = bf(xb @ bf(sd["mlpv.weight"]).t() + sd["mlpv.bias"])
    s_ = q @ k.transpose(1, 2) / 8.0
    pexp = torch.exp(s_ - s_.max(-1, keepdim=True).values)
    r = bf(xb - (bf(pexp) @ v) / pexp.sum(-1, keepdim=True))
    emu = xb + torch.relu(r @ bf(sd["out.weight"]).t() + sd["out.bias"])
    assert rel_err(out, emu) < 5e-3
    assert rel_err(a, pexp / pexp.sum(-1, keepdim=True)) < 2e-2


def test_group_mlp_maxpool_matches_materialised_path(state_dict):
    """Fused gather+MLP+max-pool == sample_and_group(...) -> mlp3 -> relu -> mlp4 -> relu -> max (oracle)."""
    from puzzlenet_b200 import pointnet_util as pu
    g = torch.Generator().manual_seed(2)
    xyz = torch.rand(2, 1024, 3, generator=g) - 0.5
    feat = torch.randn(2, 1024, 64, generator=g)
    torch.manual_seed(8)
    nx, npts, _, _, idx = po.sample_and_group(128, 0, 32, xyz, feat, knn=True, return_idx=True)
    ref = torch.relu(po._lin(state_dict, "Encoder.mlp4", torch.relu(po._lin(state_dict, "Encoder.mlp3", npts)))).max(-2).values
    got = pu.group_mlp_maxpool(xyz.to(DEV), feat.to(DEV), nx.to(DEV), idx.to(DEV),
                               state_dict["Encoder.mlp3.weight"].to(DEV), state_dict["Encoder.mlp3.bias"].to(DEV),
                               state_dict["Encoder.mlp4.weight"].to(DEV), state_dict["Encoder.mlp4.bias"].to(DEV))
    assert got.shape == ref.shape
    assert rel_err(got, ref) < REL_FP32


def test_group_mlp_maxpool_bf16(state_dict):
    """Same fused op on the tcgen05 path (layer-1 split P - Q, gathered operand, max-pool epilogue): 2e-2."""
    from puzzlenet_b200 import pointnet_util as pu
    g = torch.Generator().manual_seed(2)
    xyz = torch.rand(2, 1024, 3, generator=g) - 0.5
    feat = torch.randn(2, 1024, 64, generator=g)
    torch.manual_seed(8)
    nx, npts, _, _, idx = po.sample_and_group(128, 0, 32, xyz, feat, knn=True, return_idx=True)
    ref = torch.relu(po._lin(state_dict, "Encoder.mlp4", torch.relu(po._lin(state_dict, "Encoder.mlp3", npts)))).max(-2).values
    got = pu.group_mlp_maxpool(xyz.to(DEV), feat.to(DEV), nx.to(DEV), idx.to(DEV),
                               state_dict["Encoder.mlp3.weight"].to(DEV), state_dict["Encoder.mlp3.bias"].to(DEV),
                               state_dict["Encoder.mlp4.weight"].to(DEV), state_dict["Encoder.mlp4.bias"].to(DEV), precision=1)
    assert rel_err(got, ref) < 2e-2


def test_encoder_intermediates(cuda_model, state_dict, goldens):
    fpc, _ = synthetic_pairs(2, seed=64)
    torch.manual_seed(FPS_SEED)
    got = cuda_model.Encoder(fpc.to(DEV), return_intermediates=True)
    torch.manual_seed(FPS_SEED)
    ref = po.encoder_forward(state_dict, "Encoder", fpc)
    # indices: bit-exact, against the oracle AND the frozen reference
    for name in ("fps1", "knn1", "fps2", "knn2"):
        assert torch.equal(got[name].cpu(), ref[name]), name
        assert np.array_equal(got[name].cpu().numpy(), goldens["enc_" + name]), name
    assert torch.equal(got["x2"].cpu(), ref["x2"])
    errs = {n: rel_err(got[n], ref[n]) for n in ("x_feature", "f1f", "f2f", "attention", "out", "f_global")}
    errs["att_cat"] = rel_err(got["att_cat"], torch.cat(ref["att"] + [ref["f2f"]], -1))
    print("encoder rel errors:", errs)
    assert max(errs.values()) < REL_FP32, errs
    assert rel_err(got["f_global"], goldens["enc_f_global"]) < REL_FP32
    assert rel_err(got["out"][:, ::32], goldens["enc_out_rows"]) < REL_FP32


def test_encoder_api_tuple(cuda_model):
    fpc, _ = synthetic_pairs(1, seed=3)
    out = cuda_model.Encoder2(fpc.to(DEV))
    assert [tuple(t.shape) for t in out] == [(1, 1024), (1, 256, 3), (1, 256, 256), (1, 256, 1024), (1, 1024, 64)]
    with pytest.raises(ValueError):
        cuda_model.Encoder(torch.zeros(1, 512, 3, device=DEV))


def test_predict5_goldens(cuda_model, goldens):
    """B=2 against the frozen outputs of the unmodified reference (need=True tuple)."""
    fpc, mrpc = synthetic_pairs(2, seed=64)
    torch.manual_seed(FPS_SEED)
    r = cuda_model.predict5(make_batch(fpc.to(DEV), mrpc.to(DEV)), 0, need=True)
    assert len(r) == 8 and r[1] == [0]
    assert np.array_equal(r[2].cpu().numpy(), goldens["p5_x2_fpc"])
    assert np.array_equal(r[4].cpu().numpy(), goldens["p5_x2_mrpc"])
    errs = dict(out=rel_err(r[0], goldens["p5_out"]), de_fpcb=rel_err(r[6], goldens["p5_de_fpcb"]),
                de_mrpcb=rel_err(r[7], goldens["p5_de_mrpcb"]),
                attn_f=rel_err(r[3][:, ::16], goldens["p5_attn_fpc_rows"]),
                attn_m=rel_err(r[5][:, ::16], goldens["p5_attn_mrpc_rows"]))
    print("predict5 rel errors vs reference goldens:", errs)
    assert max(errs.values()) < REL_FP32, errs
    # pose: rotation within 0.01 deg, translation within 1e-4 of the reference
    from puzzlenet_b200 import se3
    g = se3.exp(r[0]).cpu()
    gref = torch.from_numpy(goldens["p5_mat"])
    assert po.rotation_error_deg(g[:, :3, :3], gref[:, :3, :3]).max().item() < 0.01
    assert po.translation_error(g[:, :3, 3], gref[:, :3, 3]).max().item() < 1e-4


@pytest.mark.parametrize("B", [1, 5])
def test_predict5_vs_oracle(cuda_model, state_dict, B):
    fpc, mrpc = synthetic_pairs(B, seed=100 + B)
    torch.manual_seed(B)
    out, out_again, de_f, de_m = cuda_model.predict5(make_batch(fpc.to(DEV), mrpc.to(DEV)), 0)
    assert out_again is out and de_f.shape == (B, 2, 1024)
    torch.manual_seed(B)
    ref = po.predict5(state_dict, fpc, mrpc)
    errs = dict(out=rel_err(out, ref["out"]), de_f=rel_err(de_f, ref["de_fpcb"]), de_m=rel_err(de_m, ref["de_mrpcb"]))
    print(f"B={B} rel errors:", errs)
    assert max(errs.values()) < REL_FP32, errs


def test_predict5_2d_input_and_forward_alias(cuda_model, state_dict):
    fpc, mrpc = synthetic_pairs(1, seed=9)
    batch = [fpc[0].to(DEV), mrpc[0].to(DEV)] + make_batch(fpc, mrpc)[2:]
    torch.manual_seed(4)
    out = cuda_model.forward(batch, 0)[0]                    # forward -> predict5; 2-D clouds get a batch dim
    torch.manual_seed(4)
    ref = po.predict5(state_dict, fpc[0], mrpc[0])
    assert out.shape == (1, 6) and rel_err(out, ref["out"]) < REL_FP32


def test_predict5_full_batch_properties(cuda_model):
    """BASELINE config 2 size (B=64): permutation equivariance over pairs and determinism -- properties
    that hold at any size, where the CPU oracle would take ~20 s."""
    B = 64
    fpc, mrpc = synthetic_pairs(B, seed=64)
    starts = torch.stack([torch.randint(0, n, (B,), generator=torch.Generator().manual_seed(i))
                          for i, n in enumerate((1024, 512, 1024, 512))])
    fpc, mrpc = fpc.to(DEV), mrpc.to(DEV)
    a = cuda_model.predict5(make_batch(fpc, mrpc), 0, starts=starts)
    b = cuda_model.predict5(make_batch(fpc, mrpc), 0, starts=starts)
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])                  # run-to-run deterministic
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1))
    c = cuda_model.predict5(make_batch(fpc[perm.to(DEV)], mrpc[perm.to(DEV)]), 0, starts=starts[:, perm])
    assert torch.equal(c[0], a[0][perm.to(DEV)]) and torch.equal(c[3], a[3][perm.to(DEV)])
    assert torch.isfinite(a[0]).all() and torch.isfinite(a[2]).all()


# ----------------------------------------------------------------------------- bf16 (tcgen05) path
REL_BF16 = 2e-2


@pytest.fixture()
def bf16_model(cuda_model):
    cuda_model.precision = cuda_model.Encoder.precision = cuda_model.Encoder2.precision = "bf16"
    yield cuda_model
    cuda_model.precision = cuda_model.Encoder.precision = cuda_model.Encoder2.precision = "fp32"


def test_encoder_bf16_intermediates(bf16_model, state_dict):
    fpc, _ = synthetic_pairs(2, seed=64)
    torch.manual_seed(FPS_SEED)
    got = bf16_model.Encoder(fpc.to(DEV), return_intermediates=True)
    torch.manual_seed(FPS_SEED)
    ref = po.encoder_forward(state_dict, "Encoder", fpc)
    for name in ("fps1", "knn1", "fps2", "knn2"):           # geometry stays fp32: still bit-exact
        assert torch.equal(got[name].cpu(), ref[name]), name
    errs = {n: rel_err(got[n], ref[n]) for n in ("x_feature", "f1f", "f2f", "out", "f_global")}
    errs["att_cat"] = rel_err(got["att_cat"], torch.cat(ref["att"] + [ref["f2f"]], -1))
    # the attention MAP is a softmax of logits up to ~10 with these (deliberately peaky, q/k gain 4) weights: a
    # 2^-9 operand rounding moves a dominant probability by several percent of itself; it is reported, and
    # bounded loosely, separately from the features (which are what north_star bounds at 2e-2)
    attn_err = rel_err(got["attention"], ref["attention"])
    print("bf16 encoder rel errors:", errs, "attention map:", attn_err)
    assert max(errs.values()) < REL_BF16, errs
    assert attn_err < 0.25
    # the tensor-core stem (stem_tc_kernel) is a split-bf16 product: x_feature, an output in its own right, stays at
    # fp32 accuracy in the bf16 path too
    assert errs["x_feature"] < 1e-4, errs["x_feature"]


def test_attention_timeline_hook(bf16_model):
    """pz_profile_attention_timeline: off by default, and when a buffer is registered CTA 0 of the fused attention
    layer leaves increasing SM clock stamps at its phase boundaries (the instrumented bench script relies on it)."""
    from puzzlenet_b200 import _lib
    fpc, mrpc = synthetic_pairs(2, seed=3)
    batch = make_batch(fpc.to(DEV), mrpc.to(DEV))
    tl = torch.zeros(64 + 2 * 4096, device=DEV, dtype=torch.int64)
    bf16_model.predict5(batch, 0)
    torch.cuda.synchronize()
    assert int(tl.abs().sum()) == 0
    _lib.call("pz_profile_attention_timeline", tl.data_ptr(), tl.numel())
    try:
        bf16_model.predict5(batch, 0)
        torch.cuda.synchronize()
    finally:
        _lib.call("pz_profile_attention_timeline", None, 0)
    t = tl.cpu().tolist()
    order = [15, 0, 2, 3, 4, 5, 6, 7, 10, 11, 12, 13, 17, 14]     # entry, start, q|k ... out stored (epilogue thread 0)
    stamps = [t[i] for i in order]
    assert all(b > a for a, b in zip(stamps, stamps[1:])), stamps
    assert t[64 + 1] > t[64] > 0                                   # %globaltimer at entry / exit of CTA 0
    assert t[1024] > 0                                             # the stage-1 gather GEMM stamps slots 1024..1455
    # a buffer of exactly the documented minimum (64 + 2 * clouds, B = 2 -> 4 clouds): the attention stamps land, the
    # gather GEMM's slots 1024.. are out of range and must be skipped -- nothing may be written behind the buffer
    n = 64 + 2 * 4
    guard = torch.zeros(n + 4096, device=DEV, dtype=torch.int64)
    _lib.call("pz_profile_attention_timeline", guard.data_ptr(), n)
    try:
        bf16_model.predict5(batch, 0)
        torch.cuda.synchronize()
    finally:
        _lib.call("pz_profile_attention_timeline", None, 0)
    assert int(guard[15]) > 0 and int(guard[n:].abs().sum()) == 0
    with pytest.raises(Exception):
        _lib.call("pz_profile_attention_timeline", guard.data_ptr(), 8)


@pytest.mark.parametrize("B", [2, 3])
def test_predict5_bf16_vs_oracle(bf16_model, state_dict, B):
    fpc, mrpc = synthetic_pairs(B, seed=64)
    torch.manual_seed(FPS_SEED)
    out, _, de_f, de_m = bf16_model.predict5(make_batch(fpc.to(DEV), mrpc.to(DEV)), 0)
    torch.manual_seed(FPS_SEED)
    ref = po.predict5(state_dict, fpc, mrpc)
    errs = dict(out=rel_err(out, ref["out"]), de_f=rel_err(de_f, ref["de_fpcb"]), de_m=rel_err(de_m, ref["de_mrpcb"]))
    g, gref = po.se3_exp(out.cpu()), po.se3_exp(ref["out"])
    rot = po.rotation_error_deg(g[:, :3, :3], gref[:, :3, :3]).max().item()
    tr = po.translation_error(g[:, :3, 3], gref[:, :3, 3]).max().item()
    print(f"bf16 predict5 B={B}: rel errors {errs}; pose deviation {rot:.4f} deg, {tr:.2e}")
    assert max(errs.values()) < REL_BF16, errs


def test_predict5_bf16_need_and_determinism(bf16_model):
    fpc, mrpc = synthetic_pairs(4, seed=7)
    starts = torch.stack([torch.randint(0, n, (4,), generator=torch.Generator().manual_seed(i))
                          for i, n in enumerate((1024, 512, 1024, 512))])
    batch = make_batch(fpc.to(DEV), mrpc.to(DEV))
    a = bf16_model.predict5(batch, 0, need=True, starts=starts)
    b = bf16_model.predict5(batch, 0, need=True, starts=starts)
    assert all(torch.equal(x, y) for x, y in zip(a[2:], b[2:])) and torch.equal(a[0], b[0])
    np.testing.assert_allclose(a[3].sum(-1).cpu().numpy(), 1.0, atol=1e-4)     # mean of 4 softmax maps


def test_cuda_graph_mode_matches_eager(cuda_model):
    """CUDA-graph replay (fork/join onto the internal side stream captured too) == eager, fp32 and bf16."""
    fpc, mrpc = synthetic_pairs(4, seed=21)
    starts = torch.stack([torch.randint(0, n, (4,), generator=torch.Generator().manual_seed(i))
                          for i, n in enumerate((1024, 512, 1024, 512))])
    batch = make_batch(fpc.to(DEV), mrpc.to(DEV))
    try:
        for prec in ("fp32", "bf16", "split"):
            cuda_model.precision = prec
            cuda_model.cuda_graphs = False
            eager = [t.clone() for t in cuda_model.predict5(batch, 0, starts=starts)[1:]]
            cuda_model.cuda_graphs = True
            for _ in range(3):                                   # capture, then two replays
                graphed = cuda_model.predict5(batch, 0, starts=starts)[1:]
            torch.cuda.synchronize()
            assert all(torch.equal(a, b) for a, b in zip(eager, graphed)), prec
            # different inputs through the same graph; results are FRESH tensors (as in eager mode and the reference):
            # the first result must survive the second call on the same stream
            fpc2, mrpc2 = synthetic_pairs(4, seed=22)
            b2 = make_batch(fpc2.to(DEV), mrpc2.to(DEV))
            g2 = cuda_model.predict5(b2, 0, starts=starts)[1:]
            assert all(a.data_ptr() != b.data_ptr() for a, b in zip(graphed, g2))
            torch.cuda.synchronize()
            assert all(torch.equal(a, b) for a, b in zip(eager, graphed)), prec
            cuda_model.cuda_graphs = False
            e2 = cuda_model.predict5(b2, 0, starts=starts)[1:]
            assert all(torch.equal(a, b) for a, b in zip(e2, g2)), prec
    finally:
        cuda_model.cuda_graphs = False
        cuda_model.precision = "fp32"

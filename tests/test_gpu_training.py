"""GPU parity: the training step (train-mode forward, losses, backward, Adam) through the C ABI vs the CPU training
oracle (oracle/train_oracle.py, itself pinned to the unmodified reference training_step) and the frozen reference
gradient digests (tests/golden/reference_training.npz)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import puzzle_oracle as po
from oracle import train_oracle as to
from puzzlenet_b200.weights import synthetic_state_dict
from tests.golden_inputs import FPS_SEED, training_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_training.npz")


def _fresh_model():
    from puzzlenet_b200.model5_b import TouchedRegraster
    model = TouchedRegraster(types.SimpleNamespace(dataset="vase", loss_mode=1, loss_sum=False, lr=0.9e-3))
    model.load_state_dict(synthetic_state_dict(0), strict=True)
    return model.to(DEV)


def _starts(B):
    torch.manual_seed(FPS_SEED)
    return torch.stack([torch.randint(0, 1024, (B,)), torch.randint(0, 512, (B,)),
                        torch.randint(0, 1024, (B,)), torch.randint(0, 512, (B,))])


@pytest.fixture(scope="module")
def step_result():
    """one forward_backward at B=2 on the GPU and the same step through the CPU oracle's autograd"""
    from puzzlenet_b200.training import Trainer
    model = _fresh_model()
    tr = Trainer(model)
    batch = training_inputs(2, po.se3_exp)
    terms = tr.forward_backward([t.to(DEV) for t in batch], starts=_starts(2))
    torch.cuda.synchronize()
    sd = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in synthetic_state_dict(0).items()}
    bn_state = {}
    torch.manual_seed(FPS_SEED)
    ref = to.training_loss(sd, batch, bn_state=bn_state)
    ref["loss"].backward()
    return model, tr, terms, sd, ref, bn_state


def test_loss_terms_match_oracle(step_result):
    _, tr, terms, _, ref, _ = step_result
    for k in ("loss_re", "loss_g", "loss_emd", "ce_f", "ce_m", "loss_fpcb", "loss_mrpcb", "emd_fpcb", "emd_mrpcb", "loss"):
        np.testing.assert_allclose(terms[k], float(ref[k]), rtol=2e-4, err_msg=k)
    np.testing.assert_allclose(tr.last["out"].cpu().numpy(), ref["out"].detach().numpy(), rtol=1e-4, atol=1e-5)


def test_gradients_match_oracle_autograd(step_result):
    """every live parameter: relative L2 error of the gradient tensor <= 1e-3 (fp32 accumulation-order noise of
    ~1e-5 on the large tensors; the split-K atomics and the EMD's __expf vs expf are the slack)"""
    model, tr, _, sd, _, _ = step_result
    worst = {}
    for name, p in model.named_parameters():
        ref = sd[name].grad
        if name.startswith(("fpc_decoder", "rpc_decoder")) or name == "dt":
            assert ref is None
            continue
        got = tr.flat.g(p).cpu()
        if name.endswith("mlpk.bias"):            # mathematically zero (constant per softmax row)
            q = sd[name.replace("mlpk", "mlpq")].grad
            assert got.norm() < 1e-3 * q.norm()
            continue
        err = (got - ref).norm() / ref.norm().clamp_min(1e-20)
        worst[name] = err.item()
    bad = {k: v for k, v in worst.items() if not v < 1e-3}
    assert not bad, bad
    assert len(worst) > 100


def test_gradients_match_reference_goldens(step_result):
    model, tr, terms, _, _, _ = step_result
    gold = dict(np.load(GOLDEN))
    np.testing.assert_allclose(terms["loss"], gold["loss"], rtol=2e-4)
    for name, p in model.named_parameters():
        key = "grad/" + name
        if key not in gold or name.endswith("mlpk.bias"):
            continue
        d = to.grad_digest(tr.flat.g(p).cpu())
        np.testing.assert_allclose(d[1], gold[key][1], rtol=5e-3, err_msg=name)
        np.testing.assert_allclose(d[2:], gold[key][2:], rtol=0, atol=3e-3 * np.sqrt(gold[key][1]), err_msg=name)


def test_bn_running_stats_updated(step_result):
    model, _, _, _, _, bn_state = step_result
    for name in ("Encoder.bn1", "Encoder.bn2", "Encoder2.bn1", "Encoder2.bn2"):
        mod = model.get_submodule(name)
        np.testing.assert_allclose(mod.running_mean.cpu().numpy(), bn_state[name + ".running_mean"].numpy(),
                                   rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(mod.running_var.cpu().numpy(), bn_state[name + ".running_var"].numpy(),
                                   rtol=1e-4, atol=1e-6)


def test_adam_step_matches_torch_adam():
    """three optimizer steps on the flat buffer vs torch.optim.Adam + StepLR on a CPU copy with the same grads"""
    from puzzlenet_b200.training import Trainer
    model = _fresh_model()
    tr = Trainer(model, lr=1e-3)
    ref_p = tr.flat.params.detach().cpu().clone().requires_grad_()
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, 50, 0.999)
    g = torch.Generator().manual_seed(0)
    for _ in range(3):
        grad = torch.randn(tr.flat.n, generator=g) * 0.1
        tr.flat.grads.copy_(grad)
        tr.optimizer_step()
        ref_p.grad = grad.clone()
        opt.step()
        sched.step()
    np.testing.assert_allclose(tr.flat.params.cpu().numpy(), ref_p.detach().numpy(), rtol=1e-5, atol=1e-7)
    # parameters are views of the flat buffer: the modules see the update
    assert model.tfMLP[0].weight.data_ptr() >= tr.flat.params.data_ptr()


def test_training_reduces_loss():
    """a few full steps on one fixed batch with a small learning rate (the synthetic weights put the EMD term in the
    thousands, where Adam's unit-size first steps at lr 1e-3 overshoot): the loss goes down step after step, and
    eval-mode inference still runs afterwards"""
    from puzzlenet_b200.training import Trainer
    from puzzlenet_b200.weights import make_batch
    model = _fresh_model()
    tr = Trainer(model, lr=1e-5)
    batch = [t.to(DEV) for t in training_inputs(4, po.se3_exp)]
    st = _starts(4)
    losses_seen = [tr.training_step(batch, starts=st)["loss"] for _ in range(6)]
    assert all(np.isfinite(losses_seen)) and losses_seen[-1] < losses_seen[1] < losses_seen[0], losses_seen
    model.eval()
    out = model.predict5(make_batch(batch[0], batch[1]), 4)[0]
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_sgemm_variants(ta, tb):
    from puzzlenet_b200 import training as T
    g = torch.Generator().manual_seed(5)
    M, N, K = 197, 131, 67
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    ref = (A.T if ta else A) @ (B.T if tb else B)
    Ad, Bd = A.to(DEV), B.to(DEV)
    C = torch.empty(M, N, device=DEV)
    T.gemm(Ad, Bd, C, M, N, K, ta=ta, tb=tb, lda=A.shape[1], ldb=B.shape[1], ldc=N, bias=bias.to(DEV), relu=True)
    np.testing.assert_allclose(C.cpu().numpy(), torch.relu(ref + bias).numpy(), rtol=1e-4, atol=1e-4)
    C.zero_()
    T.gemm(Ad, Bd, C, M, N, K, ta=ta, tb=tb, lda=A.shape[1], ldb=B.shape[1], ldc=N, splitk=5, alpha=0.5)
    np.testing.assert_allclose(C.cpu().numpy(), 0.5 * ref.numpy(), rtol=1e-4, atol=1e-4)
    # batched, with beta / mask / residual
    Ab = torch.randn(3, *A.shape, generator=g)
    Bb = torch.randn(3, *B.shape, generator=g)
    C0 = torch.randn(3, M, N, generator=g)
    mask = torch.randn(3, M, N, generator=g)
    res = torch.randn(3, M, N, generator=g)
    refb = (Ab.transpose(1, 2) if ta else Ab) @ (Bb.transpose(1, 2) if tb else Bb) + 2.0 * C0
    refb = torch.where(mask > 0, refb, torch.zeros(())) + res
    Cd = C0.to(DEV)
    T.gemm(Ab.to(DEV), Bb.to(DEV), Cd, M, N, K, ta=ta, tb=tb, lda=A.shape[1], ldb=B.shape[1], ldc=N, beta=2.0, batch=3,
           sa=A.numel(), sb=B.numel(), sc=M * N, mask=mask.to(DEV), ldmask=N, residual=res.to(DEV), ldres=N)
    np.testing.assert_allclose(Cd.cpu().numpy(), refb.numpy(), rtol=1e-4, atol=1e-4)


def test_predict5_training_mode_and_model_training_step():
    """predict5(training=True, need=True) = the train-mode forward (batch-stat BN) vs the oracle; model.training_step
    = Trainer.training_step behind the reference's method name."""
    model = _fresh_model()
    batch = training_inputs(2, po.se3_exp)
    st = _starts(2)
    r = model.predict5([t.to(DEV) for t in batch], 2, need=True, training=True, starts=st)
    sd = synthetic_state_dict(0)
    o = po.predict5(sd, batch[0], batch[1], need=True, starts=((st[0], st[1]), (st[2], st[3])), train_bn=True)
    assert len(r) == 8 and r[1] == [0]
    np.testing.assert_allclose(r[0].cpu().numpy(), o["out"].numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(r[2].cpu().numpy(), o["enc_fpc"]["x2"].numpy())
    for got, key in ((r[3], "enc_fpc"), (r[5], "enc_mrpc")):        # 1e-4 of the map's max magnitude, as for features
        ref_a = o[key]["attention"].numpy()
        assert np.abs(got.cpu().numpy() - ref_a).max() / np.abs(ref_a).max() < 1e-4
    ref = o["de_fpcb"].numpy()
    assert np.abs(r[6].cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-4
    out = model.training_step([t.to(DEV) for t in batch], 0, starts=st)
    assert np.isfinite(float(out["loss"])) and "loss_emd" in out["terms"]


def _tf32_round(x):
    """round-to-nearest-even to a 10-bit mantissa (what the MMA does to its operands, up to the rounding mode)"""
    i = x.contiguous().view(torch.int32)
    i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (384, 256, 96), (256, 128, 4096), (256, 64, 64), (128, 192, 32)])
def test_gemm_tf32_layouts(a_mn, b_mn, M, N, K):
    """every operand-major combination of the tcgen05 TF32 GEMM vs a float64 product of the fp32 inputs: the error
    must be explained by operand rounding to 10 mantissa bits (|err| <= 2^-10 * sum |a||b|), and far from the O(1)
    error a wrong shared-memory layout / descriptor would give"""
    from puzzlenet_b200 import _lib
    g = torch.Generator().manual_seed(M + N + K + a_mn * 2 + b_mn)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    ref = A.double() @ B.double().T
    bound = (A.abs().double() @ B.abs().double().T) * 2.0 ** -10
    Ad = (A.T.contiguous() if a_mn else A).to(DEV)
    Bd = (B.T.contiguous() if b_mn else B).to(DEV)
    C = torch.full((M, N), 7.0, device=DEV)
    _lib.call("pz_gemm_tf32", a_mn, b_mn, M, N, K, Ad.data_ptr(), Ad.shape[1], Bd.data_ptr(), Bd.shape[1], C.data_ptr(), N,
              1, None, 0, None, 0, 0, torch.cuda.current_stream().cuda_stream)
    err = (C.cpu().double() - ref).abs()
    assert (err <= bound + 1e-6).all(), (err.max().item(), bound.max().item())
    # exact agreement with a product of pre-rounded operands, up to fp32 accumulation order
    ref_t = _tf32_round(A).double() @ _tf32_round(B).double().T
    assert (C.cpu().double() - ref_t).abs().max() <= 2e-3 * ref_t.abs().max()
    # split-K (atomics into zeros) and the epilogue options
    C2 = torch.zeros(M, N, device=DEV)
    _lib.call("pz_gemm_tf32", a_mn, b_mn, M, N, K, Ad.data_ptr(), Ad.shape[1], Bd.data_ptr(), Bd.shape[1], C2.data_ptr(), N,
              5, None, 0, None, 0, 0, torch.cuda.current_stream().cuda_stream)
    assert (C2 - C).abs().max().item() <= 1e-3 * C.abs().max().item()
    bias, mask = torch.randn(N, generator=g).to(DEV), torch.randn(M, N, generator=g).to(DEV)
    C3 = torch.ones(M, N, device=DEV)
    _lib.call("pz_gemm_tf32", a_mn, b_mn, M, N, K, Ad.data_ptr(), Ad.shape[1], Bd.data_ptr(), Bd.shape[1], C3.data_ptr(), N,
              1, bias.data_ptr(), 1, mask.data_ptr(), N, 1, torch.cuda.current_stream().cuda_stream)
    want = torch.where(mask > 0, torch.relu(C + bias + 1.0), torch.zeros((), device=DEV))
    assert (C3 - want).abs().max().item() <= 1e-5 * want.abs().max().item() + 1e-5


def test_tf32_training_step_close_to_fp32():
    """precision='tf32' (tensor-core GEMMs for every large Linear, forward and backward; B=8 so that the row counts
    reach the tensor-core path): loss terms within 3e-2, every large gradient tensor within 1e-1 relative L2 and
    0.99 cosine of the fp32 path on the same batch.  Measured 3-5 % on the early layers: TF32 operand rounding
    (2^-11) is amplified by the deliberately peaky attention of the synthetic weights (logits ~10) and
    by max-pool winners that flip; the GEMM itself is pinned to rounding error by test_gemm_tf32_layouts."""
    from puzzlenet_b200.training import Trainer
    batch = [t.to(DEV) for t in training_inputs(8, po.se3_exp)]
    st = _starts(8)
    res = {}
    for prec in ("fp32", "tf32"):
        model = _fresh_model()
        tr = Trainer(model, precision=prec)
        n0 = __import__("puzzlenet_b200")._lib.load().pz_launch_count()
        terms = tr.forward_backward(batch, starts=st)
        torch.cuda.synchronize()
        res[prec] = (terms, {n: tr.flat.g(p).clone() for n, p in model.named_parameters() if id(p) in tr.flat.grad_of})
    t32, ttf = res["fp32"][0], res["tf32"][0]
    for k in ("loss", "loss_re", "loss_emd", "ce_f", "ce_m"):
        np.testing.assert_allclose(ttf[k], t32[k], rtol=3e-2, err_msg=k)
    worst = {}
    for n, g32 in res["fp32"][1].items():
        if n.endswith("mlpk.bias") or g32.numel() < 4096:
            continue
        gtf = res["tf32"][1][n]
        worst[n] = ((gtf - g32).norm() / g32.norm().clamp_min(1e-20)).item()
        cos = torch.dot(gtf.flatten(), g32.flatten()) / (gtf.norm() * g32.norm()).clamp_min(1e-30)
        # the q / k projection gradients are ill-conditioned here: with near-one-hot attention rows the softmax
        # Jacobian A*(dA - sum dA*A) subtracts nearly equal numbers, amplifying the 1e-3 input perturbation
        qk = ".mlpq." in n or ".mlpk." in n
        assert cos.item() > (0.9 if qk else 0.99), (n, cos.item())
        if qk:
            worst.pop(n)
    bad = {k: v for k, v in worst.items() if not v < 1e-1}
    assert not bad, bad
    assert max(worst.values()) > 1e-6          # the tensor-core path really ran (fp32 vs fp32 would be bit-equal)


def test_pretraining_branch_matches_oracle_and_reference():
    """training_step(pretrain=True): predict6 + pose losses; Encoder's gradients are the sum of two backward passes;
    checked against the oracle's autograd and the reference's own pretraining-step digests."""
    from puzzlenet_b200.training import Trainer
    model = _fresh_model()
    tr = Trainer(model)
    batch = training_inputs(2, po.se3_exp)
    terms = tr.forward_backward_pretrain([t.to(DEV) for t in batch], starts=_starts(2))
    torch.cuda.synchronize()
    sd = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in synthetic_state_dict(0).items()}
    bn_state = {}
    torch.manual_seed(FPS_SEED)
    ref = to.training_loss(sd, batch, bn_state=bn_state, pretrain=True)
    ref["loss"].backward()
    gold = dict(np.load(GOLDEN))
    np.testing.assert_allclose(terms["loss"], float(ref["loss"]), rtol=2e-4)
    np.testing.assert_allclose(terms["loss"], gold["pre/loss"], rtol=2e-4)
    checked = 0
    for name, p in model.named_parameters():
        if id(p) not in tr.flat.grad_of:
            continue
        got, want = tr.flat.g(p).cpu(), sd[name].grad
        if want is None:                                         # Encoder2 and the boundary heads
            assert got.abs().max().item() == 0.0, name
            continue
        if name.endswith("mlpk.bias"):
            continue
        # 1e-2, not the 1e-3 of the predict5 test: on this batch one arg-max of the tail max-pool (two candidates
        # equal to ~1e-7) resolves differently on the GPU, which re-routes one of 2048 gradient entries --
        # Encoder.out.bias (a column sum, routing-independent) still agrees to 1e-6, every pass-1 gradient to 3e-4
        assert ((got - want).norm() / want.norm().clamp_min(1e-20)).item() < 1e-2, name
        d = to.grad_digest(got)
        np.testing.assert_allclose(d[1], gold["pre/grad/" + name][1], rtol=2e-2, err_msg=name)
        checked += 1
    assert checked > 30
    for name in ("Encoder.bn1", "Encoder.bn2"):                  # two train-mode passes -> two momentum updates
        mod = model.get_submodule(name)
        np.testing.assert_allclose(mod.running_var.cpu().numpy(), bn_state[name + ".running_var"].numpy(), rtol=1e-4, atol=1e-6)
    out = model.training_step([t.to(DEV) for t in batch], 0, starts=_starts(2), pretrain=True)
    assert np.isfinite(float(out["loss"]))


def test_attention_peak_loss_terms_and_loss_modes():
    """use_emd2 / use_cd2 (model5_b.py:937-942, :1001-1043: value-only terms on the [B,B,3] attention-peak points),
    use_emd3 and another loss_mode / loss_sum combination vs the oracle."""
    from puzzlenet_b200.training import Trainer
    batch = training_inputs(3, po.se3_exp)
    st = _starts(3)
    for cfg in (dict(loss_mode=1, loss_sum=False, use_emd2=True, use_cd2=True, use_emd3=True),
                dict(loss_mode=4, loss_sum=True, use_emd2=False, use_cd2=True, use_emd3=False)):
        model = _fresh_model()
        tr = Trainer(model, types.SimpleNamespace(lr=1e-3, **cfg))
        terms = tr.forward_backward([t.to(DEV) for t in batch], starts=st)
        sd = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v.clone())
              for k, v in synthetic_state_dict(0).items()}
        ref = to.training_loss(sd, batch, starts=((st[0], st[1]), (st[2], st[3])), **cfg)
        for k in ("loss", "loss_cd2", "emd2", "loss_re", "loss_emd"):
            np.testing.assert_allclose(terms[k], float(ref[k]), rtol=3e-4, err_msg=f"{k} {cfg}")
        ref["loss"].backward()
        for name in ("tfMLP.8.weight", "Encoder2.out.weight", "MLPFpcb.4.weight"):
            got, want = tr.flat.g(model.get_parameter(name)).cpu(), sd[name].grad
            assert ((got - want).norm() / want.norm()).item() < 1e-3, (name, cfg)


@pytest.mark.parametrize("a_mn,b_mn,N,K", [(0, 0, 128, 128), (0, 1, 256, 256), (1, 0, 128, 64), (0, 1, 128, 256)])
def test_gemm_tf32_weights_resident_mode(a_mn, b_mn, N, K):
    """large-M forward / data-gradient shapes take the weights-resident path of pz_gemm_tf32 (B loaded once per CTA);
    result vs a float64 product within operand-rounding error, and identical to the streaming path (forced by a
    split-K of 1 on a zeroed output... i.e. by splitk=2 which disables residency) up to accumulation order."""
    from puzzlenet_b200 import _lib
    g = torch.Generator().manual_seed(N + K + a_mn + 2 * b_mn)
    M = 128 * 640
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    bias = torch.randn(N, generator=g)
    Ad = (A.T.contiguous() if a_mn else A).to(DEV)
    Bd = (B.T.contiguous() if b_mn else B).to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    C = torch.empty(M, N, device=DEV)
    bias_d = bias.to(DEV)
    _lib.call("pz_gemm_tf32", a_mn, b_mn, M, N, K, Ad.data_ptr(), Ad.shape[1], Bd.data_ptr(), Bd.shape[1], C.data_ptr(), N,
              1, bias_d.data_ptr(), 1, None, 0, 0, st)
    ref = torch.relu(A.double() @ B.double().T + bias.double())
    bound = (A.abs().double() @ B.abs().double().T) * 2.0 ** -10 + 1e-6
    assert ((C.cpu().double() - ref).abs() <= bound).all()
    C2 = torch.zeros(M, N, device=DEV)
    _lib.call("pz_gemm_tf32", a_mn, b_mn, M, N, K, Ad.data_ptr(), Ad.shape[1], Bd.data_ptr(), Bd.shape[1], C2.data_ptr(), N,
              2, None, 0, None, 0, 0, st)
    want = torch.relu(C2 + bias_d)
    assert (C - want).abs().max().item() <= 2e-3 * want.abs().max().item()


def test_gemm_tf32_batched_and_softmax_forward():
    """the per-cloud attention products: batched TF32 GEMM in the four operand layouts used by the attention
    backward, and the softmax forward kernel vs torch"""
    from puzzlenet_b200 import training as T
    g = torch.Generator().manual_seed(9)
    Bt, L = 5, 256
    for a_mn, b_mn, N, K in ((0, 0, 256, 64), (0, 1, 256, 256), (1, 1, 64, 256), (0, 1, 64, 256), (1, 1, 256, 256)):
        A = torch.randn(Bt, L, K, generator=g)
        Bm = torch.randn(Bt, N, K, generator=g)
        ref = A.double() @ Bm.double().transpose(1, 2)
        Ad = (A.transpose(1, 2).contiguous() if a_mn else A).to(DEV)
        Bd = (Bm.transpose(1, 2).contiguous() if b_mn else Bm).to(DEV)
        C = torch.empty(Bt, L, N, device=DEV)
        T.bgemm_tf32(a_mn, b_mn, L, N, K, Ad, Ad.shape[2], Bd, Bd.shape[2], C, N, Bt, Ad[0].numel(), Bd[0].numel(), L * N)
        bound = (A.abs().double() @ Bm.abs().double().transpose(1, 2)) * 2.0 ** -10 + 1e-6
        assert ((C.cpu().double() - ref).abs() <= bound).all(), (a_mn, b_mn, N, K)
    S = torch.randn(Bt * L, L, generator=g) * 5
    out = torch.empty(Bt * L, L, device=DEV)
    Sd = S.to(DEV)
    from puzzlenet_b200 import _lib
    _lib.call("pz_softmax_forward", Sd.data_ptr(), Bt * L, L, 0.125, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    np.testing.assert_allclose(out.cpu().numpy(), torch.softmax(S * 0.125, -1).numpy(), rtol=1e-5, atol=1e-7)


def test_checkpoint_resume_continues_the_run():
    """model.state_dict() + Trainer.state_dict() after 2 steps, loaded into a fresh model / trainer, continue with
    the same third step as the uninterrupted run (same learning rate and step count, loss to 2e-5)."""
    from puzzlenet_b200.training import Trainer
    batch = [t.to(DEV) for t in training_inputs(2, po.se3_exp)]
    st = _starts(2)
    a = _fresh_model()
    ta = Trainer(a, lr=1e-5)
    for _ in range(2):
        ta.training_step(batch, starts=st)
    ck_model = {k: v.detach().clone() for k, v in a.state_dict().items()}
    ck_opt = ta.state_dict()
    ref = ta.training_step(batch, starts=st)
    b = _fresh_model()
    b.load_state_dict(ck_model)
    tb = Trainer(b, lr=123.0)                  # wrong lr on purpose: the checkpoint restores it
    tb.load_state_dict(ck_opt)
    got = tb.training_step(batch, starts=st)
    assert got["lr"] == ref["lr"] and tb.step_count == 3
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=2e-5)      # split-K atomics in the pose-MLP forward
    # Adam normalises every gradient entry by its own magnitude: entries whose gradient is pure fp32-atomics noise
    # (|g| ~ 1e-12) still move by up to +-lr, in a direction the summation order decides -> atol = 2 lr
    np.testing.assert_allclose(tb.flat.params.cpu().numpy(), ta.flat.params.cpu().numpy(), rtol=1e-5, atol=2e-5)
    for name in ("Encoder.bn1.running_mean", "Encoder2.bn2.running_var"):
        np.testing.assert_allclose(b.state_dict()[name].cpu().numpy(), a.state_dict()[name].cpu().numpy(), rtol=1e-6)

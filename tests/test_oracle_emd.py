"""CPU: the C restatement of the reference's EMD kernels against the only golden vector the reference
holds for this path (PyTorchEMD/test_emd_loss.py:8-33: optimal matching p1[0]<->p2[1], p1[1]<->p2[0],
per-item cost 0.30 + 0.41 = 0.71, and the autograd gradients of d0/2 + 2 d1 + d2/3)."""
import numpy as np

from oracle import emd_oracle as eo

P1 = np.array([[[1.7, -0.1, 0.1], [0.1, 1.2, 0.3]]], np.float32).repeat(3, 0)
P2 = np.array([[[0.3, 1.8, 0.2], [1.2, -0.2, 0.3]]], np.float32).repeat(3, 0)


def test_known_answer_cost_and_match():
    match = eo.approxmatch(P1, P2)
    assert match.shape == (3, 2, 2)
    np.testing.assert_allclose(match[0], [[0, 1], [1, 0]], atol=1e-6)     # match[l][k]: xyz2 l <-> xyz1 k
    np.testing.assert_allclose(eo.matchcost(P1, P2, match), [0.71] * 3, rtol=1e-5)


def test_known_answer_gradients():
    w = np.array([0.5, 2.0, 1.0 / 3.0], np.float32)         # d loss / d cost_i of the reference's test loss
    match = eo.approxmatch(P1, P2)
    g1, g2 = eo.matchcost_grad(w, P1, P2, match)
    exp1 = np.stack([2 * (P1[i] - P2[i][::-1]) * w[i] for i in range(3)])
    exp2 = np.stack([2 * (P2[i] - P1[i][::-1]) * w[i] for i in range(3)])
    np.testing.assert_allclose(g1, exp1, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(g2, exp2, rtol=1e-5, atol=1e-6)


def test_match_is_a_transport_plan():
    rng = np.random.default_rng(0)
    for n, m in ((64, 64), (128, 64), (48, 96)):
        a = rng.standard_normal((2, n, 3)).astype(np.float32)
        b = rng.standard_normal((2, m, 3)).astype(np.float32)
        match = eo.approxmatch(a, b)                       # [b, m, n]
        mult_l = 1 if n >= m else m // n                   # integer division, emd_kernel.cu:29-35
        mult_r = n // m if n >= m else 1
        assert (match >= 0).all()
        assert (match.sum(1) <= mult_l + 1e-3).all()       # every xyz1 point ships at most multiL
        assert (match.sum(2) <= mult_r + 1e-3).all()
        assert match.sum() > 0.9 * 2 * min(n * mult_l, m * mult_r)

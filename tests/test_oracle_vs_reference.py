"""CPU, build container only: run the UNMODIFIED reference (through oracle/ref_shim.py) next to the
oracle on fresh seeds.  Skipped where /root/reference does not exist (the GPU box)."""
import pytest
import torch

from oracle import puzzle_oracle as po
from oracle import ref_shim
from puzzlenet_b200.weights import make_batch, synthetic_pairs, synthetic_state_dict

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load_reference()


def test_operators_bit_exact(ref):
    pu = ref.pointnet_util
    g = torch.Generator().manual_seed(21)
    x = torch.rand(3, 500, 3, generator=g) - 0.5
    torch.manual_seed(5)
    a = pu.farthest_point_sample(x, 100)
    torch.manual_seed(5)
    assert torch.equal(a, po.farthest_point_sample(x, 100))
    q = x[:, :77]
    assert torch.equal(pu.square_distance(q, x), po.square_distance(q, x))
    assert torch.equal(pu.query_ball_point(0.2, 16, x, q), po.query_ball_point(0.2, 16, x, q))
    tw = torch.randn(6, 6, generator=g)
    assert torch.allclose(ref.se3.exp(tw), po.se3_exp(tw), atol=1e-6)


def test_predict5_matches_reference(ref):
    sd = synthetic_state_dict(3)
    model = ref.model5_b.TouchedRegraster(ref_shim.reference_config())
    model.load_state_dict(sd, strict=True)
    model.eval()
    fpc, mrpc = synthetic_pairs(2, seed=5)
    with torch.no_grad():
        torch.manual_seed(77)
        r = model.predict5(make_batch(fpc, mrpc), 0, need=True)
    torch.manual_seed(77)
    o = po.predict5(sd, fpc, mrpc)
    assert torch.equal(r[2], o["enc_fpc"]["x2"]) and torch.equal(r[4], o["enc_mrpc"]["x2"])
    for a, b in ((r[0], o["out"]), (r[6], o["de_fpcb"]), (r[7], o["de_mrpcb"]), (r[3], o["enc_fpc"]["attention"])):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)

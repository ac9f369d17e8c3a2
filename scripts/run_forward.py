"""A few B=64 predict5 forwards in one precision on one stream, eager launches: the short command `ncu` profiles.
    python scripts/run_forward.py [--precision split|bf16|fp32] [--iters 3]"""
import argparse, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from puzzlenet_b200.model5_b import TouchedRegraster
from puzzlenet_b200.weights import make_batch, synthetic_pairs, synthetic_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="split")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
dev = torch.device("cuda:0")
model = TouchedRegraster(types.SimpleNamespace(dataset="vase"))
model.load_state_dict(synthetic_state_dict(0))
model.to(dev).eval()
model.precision = a.precision
fpc, mrpc = synthetic_pairs(a.batch, seed=64)
batch = make_batch(fpc.to(dev), mrpc.to(dev))
starts = torch.stack([torch.randint(0, n, (a.batch,), generator=torch.Generator().manual_seed(i))
                      for i, n in enumerate((1024, 512, 1024, 512))]).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.iters):
    if i == a.iters - 1:
        e0.record()
    out = model.predict5(batch, 0, starts=starts)
e1.record()
torch.cuda.synchronize()
print(f"{a.precision} B={a.batch}: last forward {e0.elapsed_time(e1):.3f} ms, twist[0] = {out[0][0].tolist()}")

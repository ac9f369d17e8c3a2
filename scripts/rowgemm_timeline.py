"""Kernel-internal timeline of the CTA-pair split row GEMM (first pair), taken inside a B=64 split forward.
    PZ_RG_TIMELINE=outproj|qk|vt|p1|tail python scripts/rowgemm_timeline.py"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from puzzlenet_b200 import _lib
from puzzlenet_b200.model5_b import TouchedRegraster
from puzzlenet_b200.weights import make_batch, synthetic_pairs, synthetic_state_dict

dev = torch.device("cuda:0")
model = TouchedRegraster(types.SimpleNamespace(dataset="vase"))
model.load_state_dict(synthetic_state_dict(0), strict=True)
model.to(dev).eval()
model.precision = "split"
fpc, mrpc = synthetic_pairs(64, seed=64)
batch = make_batch(fpc.to(dev), mrpc.to(dev))
for _ in range(3):
    model.predict5(batch, 0)
torch.cuda.synchronize()
tl = torch.zeros(4096, device=dev, dtype=torch.int64)
_lib.call("pz_profile_attention_timeline", tl.data_ptr(), tl.numel())
model.predict5(batch, 0)
torch.cuda.synchronize()
_lib.call("pz_profile_attention_timeline", None, 0)
t = tl.cpu().tolist()
if t[2048] or t[2048 + 512]:
    t0 = min(x for x in (t[2048], t[2048 + 512]) if x)
    print(f"timeline of launch '{os.environ.get('PZ_RG_TIMELINE')}' (us since the first CTA's entry)")
    for crank in (0, 1):
        b = 2048 + 512 * crank
        g = lambda i: (t[b + i] - t0) / 1e3 if t[b + i] else None
        fmt = lambda v: "   -  " if v is None else f"{v:6.2f}"
        print(f"CTA rank {crank}: entry {fmt(g(0))}  synced {fmt(g(1))}  weights {fmt(g(2))}  exit {fmt(g(3))}")
        print("  producer issue/arrive per job:", " ".join(f"{fmt(g(16 + 2 * j))}/{fmt(g(17 + 2 * j))}" for j in range(16) if t[b + 16 + 2 * j]))
        print("  MMAs issued per job:          ", " ".join(fmt(g(128 + j)) for j in range(16) if t[b + 128 + j]))
        print("  epilogue accf/done per tile:  ", " ".join(f"{fmt(g(192 + 2 * j))}/{fmt(g(193 + 2 * j))}" for j in range(8) if t[b + 192 + 2 * j]))

a = [t[3072 + i] for i in range(16)]
if a[0]:
    names = ["entry", "q|k landed", "S in TMEM", "P written", "O half 0", "r half 0 stored", "O half 1", "r half 1 stored",
             "out half 0 accumulated", "out half 0 stored", "out half 1 accumulated", "out half 1 stored",
             "next q|k accumulated", "next v block 0 accumulated", "next v block 1 accumulated", "next q|k|v stored"]
    a = [v for v in a if v]
    print("attention_split_kernel, CTA 0 (us since entry): " + ", ".join(f"{n} {(v - a[0]) / 1e3:.2f}" for n, v in zip(names, a)))


"""Host-side cost of one predict5 call in CUDA-graph mode (the pipelined bench legs are only as fast as the host can
issue replays): wall time per call at B=1 (GPU work negligible) and at B=64 without synchronising."""
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from puzzlenet_b200.model5_b import TouchedRegraster  # noqa: E402
from puzzlenet_b200.weights import make_batch, synthetic_pairs, synthetic_state_dict  # noqa: E402

dev = torch.device("cuda:0")
model = TouchedRegraster(types.SimpleNamespace(dataset="vase"))
model.load_state_dict(synthetic_state_dict(0), strict=True)
model.to(dev).eval()
model.precision = "bf16"
model.cuda_graphs = True
for B in (1, 64):
    fpc, mrpc = synthetic_pairs(B, seed=1)
    batch = make_batch(fpc.to(dev), mrpc.to(dev))
    starts = torch.zeros(4, B, dtype=torch.int64, device=dev)
    for _ in range(5):
        model.predict5(batch, 0, starts=starts)
    torch.cuda.synchronize()
    n = 300
    t0 = time.perf_counter()
    for _ in range(n):
        model.predict5(batch, 0, starts=starts)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"B={B}: host issue {1e6 * (t1 - t0) / n:.1f} us/call, incl. drain {1e6 * (t2 - t0) / n:.1f} us/call")

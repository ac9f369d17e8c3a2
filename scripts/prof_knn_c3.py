"""kNN at the C3 shape (64 clouds x 11000 points, 1024 queries, k = 32) alone, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from puzzlenet_b200 import pointnet_util as pu
g = torch.Generator().manual_seed(0)
xyz = (torch.rand(64, 11000, 3, generator=g) - 0.5).cuda()
q = xyz[:, :1024].contiguous()
for _ in range(3):
    idx = pu.knn_point(32, xyz, q)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); idx = pu.knn_point(32, xyz, q); e1.record(); torch.cuda.synchronize()
print("knn C3 ms", e0.elapsed_time(e1))

"""Run the geometry kernels alone at the C2 shapes (128 clouds) so that ncu can capture them cheaply."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from puzzlenet_b200 import pointnet_util as pu

dev = "cuda:0"
g = torch.Generator().manual_seed(0)
xyz = (torch.rand(128, 1024, 3, generator=g) - 0.5).to(dev)
for _ in range(3):
    torch.manual_seed(1)
    idx = pu.farthest_point_sample(xyz, 512)
    new_xyz = pu.index_points(xyz, idx)
    knn = pu.knn_point(32, xyz, new_xyz)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); idx = pu.farthest_point_sample(xyz, 512); e1.record(); knn = pu.knn_point(32, xyz, new_xyz); e2.record()
torch.cuda.synchronize()
print("fps1 ms", e0.elapsed_time(e1), "knn1 ms", e1.elapsed_time(e2))

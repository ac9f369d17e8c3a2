import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from puzzlenet_b200 import pointnet_util as pu
from puzzlenet_b200.weights import synthetic_state_dict
B, N, S, K, D = 64, 11000, 1024, 32, 64
dev = "cuda:0"
xyz = (torch.rand(B, N, 3, generator=torch.Generator().manual_seed(3)) - 0.5).to(dev)
feat = torch.randn(B, N, D, generator=torch.Generator().manual_seed(4)).to(dev)
sd = synthetic_state_dict(0)
w1, b1 = sd["Encoder.mlp3.weight"].to(dev), sd["Encoder.mlp3.bias"].to(dev)
w2, b2 = sd["Encoder.mlp4.weight"].to(dev), sd["Encoder.mlp4.bias"].to(dev)
torch.manual_seed(5)
fps_idx = pu.farthest_point_sample(xyz, S)
new_xyz = pu.index_points(xyz, fps_idx)
idx = pu.knn_point(K, xyz, new_xyz)
for _ in range(2):
    pu.group_mlp_maxpool(xyz, feat, new_xyz, idx, w1, b1, w2, b2, precision=1)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pu.group_mlp_maxpool(xyz, feat, new_xyz, idx, w1, b1, w2, b2, precision=1)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))

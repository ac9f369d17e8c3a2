import sys, torch
sys.path.insert(0, '/root/repo')
from puzzlenet_b200 import pointnet_util as pu
dev='cuda:0'
for seed in (3,4,5,6):
    xyz = (torch.rand(64, 11000, 3, generator=torch.Generator().manual_seed(seed)) - 0.5).to(dev)
    torch.manual_seed(5)
    idx = pu.farthest_point_sample(xyz, 1024)
    nx = pu.index_points(xyz, idx)
    for _ in range(2): pu.knn_point(32, xyz, nx)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): pu.knn_point(32, xyz, nx)
    e1.record(); torch.cuda.synchronize()
    print('seed',seed,'knn ms',e0.elapsed_time(e1)/5)

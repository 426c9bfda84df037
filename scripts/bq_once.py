import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth, numpy as np
from tsmdet_b200 import pointnet2_utils as pu
dev = torch.device("cuda:0")
xyz = torch.from_numpy(synth.cloud_ground_objects(16, 16384, 1)).to(dev)
i = pu.farthest_point_sample(xyz, 4096); new = torch.gather(xyz, 1, i.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
for _ in range(3):
    pu.ball_query(0.2, 16, xyz, new)
torch.cuda.synchronize()
x2 = new; i2 = pu.farthest_point_sample(x2, 1024); new2 = torch.gather(x2, 1, i2.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
for _ in range(3):
    pu.ball_query(0.8, 32, x2, new2)
torch.cuda.synchronize()
print("ok")

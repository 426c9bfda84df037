import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth
from tsmdet_b200 import pointnet2_utils as pu
b, n, m = 16, int(os.environ.get("N", 16384)), int(os.environ.get("M", 1024))
xyz = torch.from_numpy(synth.cloud_ground_objects(b, n, 1)).cuda()
for _ in range(2):
    idx = pu.farthest_point_sample(xyz, m)
torch.cuda.synchronize()
print("ok", idx.shape)

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["TSMDET_LIB"] = os.path.join(ROOT, "tsm-det-pointcloud-_b200", "build", "prof", "libtsmdet_b200_prof.so")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth
from tsmdet_b200 import pointnet2_utils as pu
dev = torch.device("cuda:0")
os.environ["TSMDET_FPS_ALGO"] = "bucket"
for (n, m, T, P) in [(16384, 4096, 1024, 16), (4096, 1024, 512, 8)]:
    os.environ["TSMDET_FPSB_T"] = str(T); os.environ["TSMDET_FPSB_P"] = str(P)
    xyz = torch.from_numpy(synth.cloud_ground_objects(2, n, 1)).to(dev)
    pu.farthest_point_sample(xyz, m); torch.cuda.synchronize()

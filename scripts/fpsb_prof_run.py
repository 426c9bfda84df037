"""Per-phase clock() totals of the multi-pick bucketed FPS (library built with -DFPSB_PROF; TSMDET_LIB points at it)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth
from tsmdet_b200 import pointnet2_utils as pu
dev = torch.device("cuda:0")
os.environ["TSMDET_FPS_ALGO"] = "bucket"
for (n, m) in ((16384, 4096),):
    xyz = torch.from_numpy(getattr(synth, os.environ.get("GEN", "cloud_ground_objects"))(16, n, 1)).to(dev)
    for K in (4, 8):
        os.environ["TSMDET_FPSB_K"] = str(K)
        pu.farthest_point_sample(xyz, m); torch.cuda.synchronize()

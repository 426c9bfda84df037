import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
from tsmdet_b200 import pointnet2_utils as pu
from tsmdet_b200.pipeline import SABackboneNMS
from tsmdet_b200.pointnet2_modules import gather_xyz, sa_mlp_maxpool
dev = torch.device("cuda:0")
xyz = torch.from_numpy(bench.make_inputs(16, 0)[0]).to(dev)
eng = SABackboneNMS(precision="bf16", use_graph=False).to(dev)
idx = pu.farthest_point_sample(xyz, 1024); x2 = gather_xyz(xyz, idx); x3 = x2[:, :512].contiguous()
l3 = eng.backbone.layers[2]; f2 = torch.rand(16, 128, 1024, device=dev)
c3, i3 = pu.ball_query(1.6, 32, x2, x3); out3 = torch.empty(16, 256, 512, device=dev); fl = l3._folded_layers()[0]
for _ in range(3):
    sa_mlp_maxpool(x2, x3, f2, i3, c3, fl, out3, 0, precision="bf16")
torch.cuda.synchronize(); print("ok")

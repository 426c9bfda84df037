"""One un-graphed pass of the benchmarked step (for ncu launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from tsmdet_b200.pipeline import SABackboneNMS
dev = torch.device("cuda:0")
eng = SABackboneNMS(precision="bf16", use_graph=False).to(dev)
d = [torch.from_numpy(a).to(dev) for a in bench.make_inputs(16, 0)]
for _ in range(int(os.environ.get("REPS", "3"))):
    r = eng.forward_device(*d)
torch.cuda.synchronize()
print("ok", {k: tuple(v.shape) for k, v in r.items()})

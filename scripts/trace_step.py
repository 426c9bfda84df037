import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
from tsmdet_b200.pipeline import SABackboneNMS
dev = torch.device("cuda:0")
eng = SABackboneNMS(precision="bf16", use_graph=False).to(dev)
d = [torch.from_numpy(a).to(dev) for a in bench.make_inputs(16, 0)]
for _ in range(3):
    tr = eng.trace_step(*d)
for n, t in tr: print(f"{t:8.3f} ms  {n}")

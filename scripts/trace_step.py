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
if os.environ.get("NO_NMS"):
    dummy = (torch.zeros(16, 512, dtype=torch.int64, device=dev), torch.zeros(16, dtype=torch.int32, device=dev))
    eng._nms_two_pass = lambda b, s: dummy
    for _ in range(3):
        tr = eng.trace_step(*d)
    print("--- without NMS")
    for n, t in tr: print(f"{t:8.3f} ms  {n}")

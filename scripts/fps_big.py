import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, synth
from tsmdet_b200 import pointnet2_utils as pu
from oracle import build_ref
dev = torch.device("cuda:0")
def timeit(fn, reps=2):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
ref = build_ref.load_ref("pointnet2_batch_cuda")
for (b, n, m) in [(16, 20000, 4096), (8, 65536, 16384), (2, 180000, 16384), (8, 180000, 16384)]:
    xyz = torch.from_numpy(synth.cloud_uniform(b, n, 1, synth.WAYMO_RANGE)).to(dev)
    ms = timeit(lambda: pu.farthest_point_sample(xyz, m))
    out = dict(b=b, n=n, m=m, ms=round(ms, 3), us_per_iter=round(1000 * ms / (m - 1), 3))
    if ref is not None and os.environ.get("WITH_REF"):
        temp = torch.full((b, n), 1e10, device=dev); idx = torch.zeros((b, m), dtype=torch.int32, device=dev)
        out["ref_ms"] = round(timeit(lambda: ref.farthest_point_sampling_wrapper(b, n, m, xyz, temp, idx), reps=1), 2)
        got = pu.farthest_point_sample(xyz, m); torch.cuda.synchronize()
        out["equal_ref"] = bool(torch.equal(got, idx))
    print(json.dumps(out), flush=True)

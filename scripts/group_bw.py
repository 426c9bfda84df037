"""HBM-bound ops of the path: achieved GB/s against algorithmic bytes, graph-timed."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth
from tsmdet_b200 import pointnet2_utils as pu
dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6546.6) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6546.6
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()  # cold L2: the roofline is HBM
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(side); g.replay(); e.record(side); e.synchronize(); tot += s.elapsed_time(e)
    torch.cuda.synchronize()
    return tot / reps
rows = []
for (name, b, c, n, m, s, r) in [("K-L1", 16, 1, 16384, 4096, 16, 0.2), ("K-L2", 16, 32, 4096, 1024, 32, 0.8), ("K-L3", 16, 128, 1024, 512, 32, 1.6)]:
    xyz = torch.from_numpy(synth.cloud_ground_objects(b, n, 1)).to(dev)
    feats = torch.rand(b, c, n, device=dev)
    new = torch.gather(xyz, 1, pu.farthest_point_sample(xyz, m).long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    cnt, idx = pu.ball_query(r, s, xyz, new)
    ms = t(lambda: pu.grouping_operation(feats, idx))
    by = b * (4 * c * n + 4 * m * s + 4 * c * m * s)
    rows.append(dict(op="grouping_operation", shape=name, ms=round(ms, 4), alg_MB=round(by / 1e6, 2), GBs=round(by / ms / 1e6, 1), frac=round(by / ms / 1e6 / PEAK, 3)))
    qg = pu.QueryAndGroup(r, s, use_xyz=True)
    ms = t(lambda: qg(xyz, new, feats))
    by = b * (12 * n + 12 * m + 4 * m * s + 4 * m) + b * (4 * c * n + 4 * (3 + c) * m * s + 4 * 3 * m * s)
    rows.append(dict(op="QueryAndGroup (materialised)", shape=name, ms=round(ms, 4), alg_MB=round(by / 1e6, 2), GBs=round(by / ms / 1e6, 1), frac=round(by / ms / 1e6 / PEAK, 3)))
# three_interpolate, Waymo FP shape
b, c, m, n = 8, 128, 16384, 65536
feats = torch.rand(b, c, m, device=dev); idx = torch.randint(0, m, (b, n, 3), device=dev, dtype=torch.int32)
idx = torch.sort(idx, dim=1)[0].contiguous()  # neighbouring unknowns interpolate from neighbouring knowns
w = torch.rand(b, n, 3, device=dev)
ms = t(lambda: pu.three_interpolate(feats, idx, w))
by = b * (4 * c * m + 24 * n + 4 * c * n)
rows.append(dict(op="three_interpolate", shape="W C=128", ms=round(ms, 4), alg_MB=round(by / 1e6, 2), GBs=round(by / ms / 1e6, 1), frac=round(by / ms / 1e6 / PEAK, 3)))
unk = torch.from_numpy(synth.cloud_uniform(b, n, 3, synth.WAYMO_RANGE)).to(dev); kn = unk[:, ::4, :].contiguous()
ms = t(lambda: pu.three_nn(unk, kn), reps=3)
rows.append(dict(op="three_nn (grid)", shape="W 65536x16384", ms=round(ms, 4), tests_per_s=round(b * n * m / ms * 1e3 / 1e12, 2)))
for r in rows: print(json.dumps(r), flush=True)

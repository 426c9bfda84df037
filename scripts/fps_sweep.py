"""Times FPS for every (cluster size, block size) launch shape, plus the reference CUDA kernels."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth
from tsmdet_b200 import _lib, pointnet2_utils as pu
from oracle import build_ref

dev = torch.device("cuda:0")
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps

only = sys.argv[1] if len(sys.argv) > 1 else None
res = []
shapes = [(16, 16384, 4096), (16, 4096, 1024), (16, 1024, 512), (8, 65536, 4096), (32, 16384, 1024)]
if os.environ.get("QUICK"): shapes = shapes[:2]
for (b, n, m) in shapes:
    xyz = torch.from_numpy(synth.cloud_ground_objects(b, n, 1)).to(dev)
    for c in ((8,) if os.environ.get("QUICK") else (1, 2, 4, 8, 16)):
        for t in ((128,) if os.environ.get("QUICK") else (128, 256, 512, 1024)):
            os.environ["TSMDET_FPS_CLUSTER"] = str(c); os.environ["TSMDET_FPS_THREADS"] = str(t)
            cc, tt, pp, ss = (ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int())
            _lib.call("tsmdet_fps_plan", b, n, ctypes.byref(cc), ctypes.byref(tt), ctypes.byref(pp), ctypes.byref(ss))
            if cc.value != c or tt.value != t: continue
            try:
                ms = timeit(lambda: pu.farthest_point_sample(xyz, m))
            except Exception as ex:
                print("fail", b, n, c, t, ex); continue
            r = dict(b=b, n=n, m=m, cluster=c, threads=t, P=pp.value, smem=ss.value, ms=round(ms, 4), us_per_iter=round(1000 * ms / (m - 1), 4))
            res.append(r); print(json.dumps(r), flush=True)
    os.environ.pop("TSMDET_FPS_CLUSTER"); os.environ.pop("TSMDET_FPS_THREADS")
    ms = timeit(lambda: pu.farthest_point_sample(xyz, m)); print(json.dumps(dict(b=b, n=n, m=m, impl="default_plan", ms=round(ms, 4), us_per_iter=round(1000 * ms / (m - 1), 4))), flush=True)
    ms = timeit(lambda: pu.farthest_point_sample_chained(xyz, m)); print(json.dumps(dict(b=b, n=n, m=m, impl="chained_root", ms=round(ms, 4), us_per_iter=round(1000 * ms / (m - 1), 4))), flush=True)
    ref = build_ref.load_ref("pointnet2_batch_cuda")
    if ref is not None and n <= 16384:
        temp = torch.full((b, n), 1e10, device=dev); idx = torch.zeros((b, m), dtype=torch.int32, device=dev)
        ms = timeit(lambda: ref.farthest_point_sampling_wrapper(b, n, m, xyz, temp, idx), reps=2)
        print(json.dumps(dict(b=b, n=n, m=m, impl="reference_cuda", ms=round(ms, 3), us_per_iter=round(1000 * ms / (m - 1), 3))), flush=True)

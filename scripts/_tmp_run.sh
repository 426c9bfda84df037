timeout 600 python -m pytest tests/test_fps_gpu.py -m gpu -x -q -k "bucket_cluster" 2>&1 | tail -15
timeout 300 python - <<'PY'
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, numpy as np, synth
from tsmdet_b200 import _lib, pointnet2_utils as pu
dev = torch.device('cuda:0')
def T(a): return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cases = [("20000->4096 B16", T(synth.cloud_ground_objects(16, 20000, 5)), 4096),
         ("65536->16384 B8", T(synth.cloud_uniform(8, 65536, 7, synth.WAYMO_RANGE)), 16384),
         ("180000->16384 B2", T(synth.cloud_uniform(2, 180000, 9, synth.WAYMO_RANGE)), 16384)]
for k in ("1", "2", "4", None):
    if k is None: os.environ.pop("TSMDET_FPSC_K", None)
    else: os.environ["TSMDET_FPSC_K"] = k
    _lib.reload_options()
    for name, x, m in cases:
        pu.farthest_point_sample(x, m); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); idx = pu.farthest_point_sample(x, m); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        print(f"FPSC_K={k}  {name}: {min(ts):.3f} ms  {1000*min(ts)/(m-1):.3f} us/pick  checksum {int(idx.long().sum())}")
PY

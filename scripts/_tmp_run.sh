timeout 900 python -m pytest tests/test_fps_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python - <<'PY'
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, numpy as np, synth
from tsmdet_b200 import _lib, pointnet2_utils as pu
dev = torch.device('cuda:0')
def T(a): return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cases = [("16384->4096 B16 objects", T(synth.cloud_ground_objects(16, 16384, 5)), 4096),
         ("16384->4096 B16 dup", T(synth.cloud_dup_padded(16, 16384, 6)), 4096),
         ("20000->4096 B16", T(synth.cloud_ground_objects(16, 20000, 5)), 4096),
         ("65536->16384 B8", T(synth.cloud_uniform(8, 65536, 7, synth.WAYMO_RANGE)), 16384)]
_lib.call("tsmdet_fps_configure", 2)
for name, x, m in cases:
    pu.farthest_point_sample(x, m); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); idx = pu.farthest_point_sample(x, m); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    print(f"{name}: {min(ts):.3f} ms  {1000*min(ts)/(m-1):.3f} us/pick")
PY

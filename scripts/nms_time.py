import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
from tsmdet_b200 import iou3d_nms_utils
dev = torch.device("cuda:0")
d = [torch.from_numpy(a).to(dev) for a in bench.make_inputs(16, 0)]
boxes, scores = d[2], d[3]
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        fn(); g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side): fn()
        g.replay(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(side)
        for _ in range(reps): g.replay()
        e.record(side)
    torch.cuda.synchronize(); return s.elapsed_time(e) / reps
for thr in (0.01, 0.1):
    ms = t(lambda: iou3d_nms_utils.nms_gpu_batch(boxes, scores, thr))
    sel, num = iou3d_nms_utils.nms_gpu_batch(boxes, scores, thr)
    print(json.dumps(dict(ctas=os.environ.get("TSMDET_NMS_CTAS_PER_SM", "2"), thresh=thr, ms=round(ms, 4), kept=num.tolist()[:4])), flush=True)

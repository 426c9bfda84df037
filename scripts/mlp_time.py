import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
from tsmdet_b200 import pointnet2_utils
from tsmdet_b200.pipeline import SABackboneNMS
from tsmdet_b200.pointnet2_modules import gather_xyz, sa_mlp_maxpool
dev = torch.device("cuda:0")
eng = SABackboneNMS(precision="bf16", use_graph=False).to(dev)
xyz, feats, boxes, scores = [torch.from_numpy(a).to(dev) for a in bench.make_inputs(16, 0)]
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
cur_xyz, cur_f = xyz, feats
out_ms = []
with torch.no_grad():
    for li, layer in enumerate(eng.backbone.layers):
        m = layer.npoint_list[0]; g = layer.groupers[0]
        idx = pointnet2_utils.farthest_point_sample(cur_xyz, m); new_xyz = gather_xyz(cur_xyz, idx)
        cnt, bidx = pointnet2_utils.ball_query(g.radius, g.nsample, cur_xyz, new_xyz)
        layers = layer._folded_layers()[0]
        out = torch.empty((16, layers[-1][0].shape[0], m), device=dev)
        ms = t(lambda: sa_mlp_maxpool(cur_xyz, new_xyz, cur_f, bidx, cnt, layers, out, 0, precision=layer.precision))
        out_ms.append(round(ms, 4)); cur_xyz, cur_f = new_xyz, out
print(json.dumps(dict(alt=os.environ.get("TSMDET_MLP_TMEM_ALT"), occ=os.environ.get("TSMDET_MLP_OCC"), ms=out_ms, chk=float(cur_f.sum()))))

"""Which kernels slow the FPS chain down when they run beside it?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
from tsmdet_b200 import pointnet2_utils as pu, iou3d_nms_utils as iu
from tsmdet_b200.pipeline import SABackboneNMS
from tsmdet_b200.pointnet2_modules import gather_xyz, sa_mlp_maxpool
dev = torch.device("cuda:0")
xyz, feats, boxes, scores = [torch.from_numpy(a).to(dev) for a in bench.make_inputs(16, 0)]
eng = SABackboneNMS(precision="bf16", use_graph=False).to(dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def fps(): pu.farthest_point_sample(xyz, 4096)
idx = pu.farthest_point_sample(xyz, 4096); nx = gather_xyz(xyz, idx)
def nms(): eng._nms_two_pass(boxes, scores)
def bq(): pu.ball_query(0.2, 16, xyz, nx)
l3 = eng.backbone.layers[2]; f2 = torch.rand(16, 128, 1024, device=dev); x2 = nx[:, :1024].contiguous(); x3 = x2[:, :512].contiguous()
c3, i3 = pu.ball_query(1.6, 32, x2, x3); out3 = torch.empty(16, 256, 512, device=dev); fl = l3._folded_layers()[0]
def mlp3(): sa_mlp_maxpool(x2, x3, f2, i3, c3, fl, out3, 0, precision="bf16")
def timed(fa, fb=None, reps=3):
    for _ in range(2):
        fa(); 
        if fb: fb()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1): fa()
        if fb:
            with torch.cuda.stream(s2): fb()
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
print("fps alone      %.3f ms" % timed(fps))
print("nms alone      %.3f ms" % timed(nms))
print("bq1 alone      %.3f ms" % timed(bq))
print("mlp3 alone     %.3f ms" % timed(mlp3))
print("fps || fps     %.3f ms" % timed(fps, fps))
print("fps || nms     %.3f ms" % timed(fps, nms))
print("fps || bq1     %.3f ms" % timed(fps, bq))
print("fps || mlp3x4  %.3f ms" % timed(fps, lambda: [mlp3() for _ in range(4)]))

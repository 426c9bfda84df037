"""BASELINE config 2b: the reference-true layer 0 of fast_cpc.yaml -- N = 20000 -> 4096 (d-fps), three
dilated scales (0-0.2, 0.2-0.4, 0.4-0.8; nsample 32; MLPs [1+3,16,16,32] x2, [1+3,32,32,64]), aggregation Conv1d 128 -> 64,
then score-weighted FPS 4096 -> 512 -- B = 16, timed from CUDA graphs; checked against the eager fp32 stack."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth
from tsmdet_b200 import pointnet2_utils as pu
from tsmdet_b200.pointnet2_modules import PointnetSAModuleFSMSG
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
B, N = 16, 20000
xyz = torch.from_numpy(np.concatenate([synth.cloud_ground_objects(B // 2, N, 0), synth.cloud_dup_padded(B - B // 2, N, 1)], 0)).to(dev)
feats = torch.rand(B, 1, N, device=dev)
torch.manual_seed(0)
def make(precision):
    torch.manual_seed(0)
    return PointnetSAModuleFSMSG(npoint_list=[4096], sample_range_list=[[0, N]], sample_method_list=['d-fps'],
                                 radii=[0.2, 0.4, 0.8], nsamples=[32, 32, 32], mlps=[[1, 16, 16, 32], [1, 16, 16, 32], [1, 32, 32, 64]],
                                 dilated_radius_group=True, aggregation_mlp=[64], fused=True, precision=precision).to(dev).eval()
layer = make("bf16")
def step():
    with torch.no_grad():
        new_xyz, nf, idx = layer(xyz, feats)
        w = torch.sigmoid(nf[:, 0, :]).contiguous()           # stand-in for the head's per-point score
        idx2 = pu.furthest_point_sample_weights(new_xyz, w, 512)
    return new_xyz, nf, idx, idx2
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
res = {}
for algo in ("cluster", "auto"):
    if algo == "cluster": os.environ["TSMDET_FPS_ALGO"] = "cluster"
    else: os.environ.pop("TSMDET_FPS_ALGO", None)
    res[f"step_ms_{algo}"] = round(t(step), 4)
with torch.no_grad():
    new_xyz, nf, idx, idx2 = step()
    res["fps_20000_4096_ms"] = round(t(lambda: pu.farthest_point_sample(xyz, 4096)), 4)
    res["sfps_4096_512_ms"] = round(t(lambda: pu.furthest_point_sample_weights(new_xyz, torch.sigmoid(nf[:, 0, :]).contiguous(), 512)), 4)
    for i, g in enumerate(layer.groupers):
        res[f"query_scale{i}_ms"] = round(t(lambda: pu.ball_query_dilated(g.radius_in, g.radius_out, g.nsample, xyz, new_xyz)), 4)
    # parity of the fused bf16 layer vs the eager fp32 stack
    ref = make("fp32"); ref.fused = False
    _, nf_ref, idx_ref = ref(xyz[:4], feats[:4])
    assert torch.equal(idx[:4], idx_ref)
    err = float((nf[:4] - nf_ref).abs().max()); scale = float(nf_ref.abs().max())
    res["bf16_vs_eager_fp32_max_abs"] = err; res["out_scale"] = scale
res["frames_per_s"] = round(B / res["step_ms_auto"] * 1e3, 1)
print(json.dumps(res))

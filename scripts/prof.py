"""One parametrised runner for profiling / timing single ops of the path (replaces the one-off scripts of round 1).

    python scripts/prof.py mlp1|mlp2|mlp3|fp|fps|bq|nms|voxel|step [--reps 3] [--time]

Runs the op `reps` times on BASELINE-shaped inputs (config 2 / 3; `fp` = config 4's FP layer) so that
`ncu -k regex:<kernel> ... python scripts/prof.py <op>` captures exactly that kernel; --time prints CUDA-event times.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from tsmdet_b200 import _lib, iou3d_nms_utils  # noqa: E402
from tsmdet_b200 import pointnet2_utils as pu  # noqa: E402
from tsmdet_b200.pipeline import SABackboneNMS  # noqa: E402
from tsmdet_b200.pointnet2_modules import PointnetFPModule, gather_xyz, sa_mlp_maxpool  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("op")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    xyz_np, feats_np, boxes_np, scores_np = bench.make_inputs(16, 0)
    xyz, feats = torch.from_numpy(xyz_np).to(dev), torch.from_numpy(feats_np).to(dev)
    eng = SABackboneNMS(precision=a.precision, use_graph=False).to(dev)
    _lib.call("tsmdet_fps_configure", 2)
    levels = [(xyz, feats)]
    with torch.no_grad():
        outs = eng.backbone(xyz, feats)
    for o in outs:
        levels.append((o[0], o[1]))
    fn = None
    if a.op in ("mlp1", "mlp2", "mlp3"):
        li = int(a.op[-1]) - 1
        layer = eng.backbone.layers[li]
        src_xyz, src_f = levels[li]
        new_xyz = levels[li + 1][0]
        g = layer.groupers[0]
        cnt, bidx = pu.ball_query(g.radius, g.nsample, src_xyz, new_xyz)
        fl = layer._folded_layers()[0]
        img = layer._packed_layers(src_f.shape[1], True)[0]
        out = torch.empty((16, fl[-1][0].shape[0], new_xyz.shape[1]), device=dev)
        assert img is not None, "the tensor kernel does not take this layer at this precision"
        fn = lambda: sa_mlp_maxpool(src_xyz, new_xyz, src_f, bidx, cnt, fl, out, 0, precision=a.precision, packed=img)  # noqa: E731
    elif a.op == "fp":
        import synth
        x = torch.from_numpy(synth.cloud_uniform(8, 65536, 7, synth.WAYMO_RANGE)).to(dev)
        nx = gather_xyz(x, pu.farthest_point_sample(x, 16384))
        f2, kf = torch.rand((8, 2, 65536), device=dev), torch.rand((8, 128, 16384), device=dev)
        torch.manual_seed(0)
        fp = PointnetFPModule(mlp=[130, 128, 128], precision=a.precision).to(dev).eval()
        fn = lambda: fp(x, nx, f2, kf)  # noqa: E731
    elif a.op == "fps":
        fn = lambda: pu.farthest_point_sample(xyz, 4096)  # noqa: E731
    elif a.op == "bq":
        nx = levels[1][0]
        fn = lambda: pu.ball_query(0.2, 16, xyz, nx)  # noqa: E731
    elif a.op == "nms":
        b, s = torch.from_numpy(boxes_np).to(dev), torch.from_numpy(scores_np).to(dev)
        fn = lambda: iou3d_nms_utils.nms_gpu_batch(b, s, 0.01)  # noqa: E731
    elif a.op == "voxel":
        from tsmdet_b200 import voxel_aggregation_utils as vau
        nx, nf = levels[1][0], torch.rand((16, 64, 4096), device=dev)
        fn = lambda: vau.voxelize_centroids(nx, nf, [0.05, 0.05, 0.1], [0, -40, -3, 70.4, 40, 1])  # noqa: E731
    elif a.op == "step":  # one whole un-graphed step (for `ncu --metrics gpu__time_duration.sum` launch lists)
        b, s = torch.from_numpy(boxes_np).to(dev), torch.from_numpy(scores_np).to(dev)
        eng.chain_fps = True
        fn = lambda: eng.forward_device(xyz, feats, b, s)  # noqa: E731
    else:
        raise SystemExit(f"unknown op {a.op}")
    with torch.no_grad():
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            fn()
            ev1.record()
            torch.cuda.synchronize()
            ts.append(ev0.elapsed_time(ev1))
    if a.time:
        print(a.op, "ms:", [round(t, 4) for t in ts])
    print("ok")


if __name__ == "__main__":
    main()

"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

    # GPU box (reference CUDA kernels; needs oracle/_ref/*.so built by oracle/build_ref.py):
    python tests/golden/make_golden.py --out gpurun_out/golden
    # this container (reference CPU IoU only):
    python tests/golden/make_golden.py --cpu-only --out tests/golden

The fixtures pin the CPU oracle (``oracle/*.c``) to the reference's own outputs: there are no
golden vectors or tests in the reference itself.  Inputs are stored next to the
outputs so nothing depends on RNG stability.  Files are small (< 1 MB total, compressed).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402
from oracle import build_ref  # noqa: E402


def cpu_fixtures(out):
    import torch

    ref = build_ref.load_ref("iou3d_nms_cuda")
    assert ref is not None, "oracle/_ref/iou3d_nms_cuda.so missing: run python oracle/build_ref.py"
    a = np.concatenate([synth.boxes_random(96, seed=1), synth.boxes_clustered(160, seed=2, centres=24)], 0)
    b = np.concatenate([synth.boxes_clustered(128, seed=3, centres=24), synth.boxes_random(64, seed=4)], 0)
    # exact duplicates, shared edges and axis-aligned boxes: the degenerate branches of the clipping code
    a[5] = b[7]
    a[6] = b[8]; a[6, 0] += b[8, 3]          # touching along an edge
    a[7, 6] = 0.0; b[9] = a[7]; b[9, 0] += 0.5
    ans = torch.zeros((a.shape[0], b.shape[0]), dtype=torch.float32)
    ref.boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(b), ans)
    np.savez_compressed(os.path.join(out, "iou_bev_cpu.npz"), boxes_a=a, boxes_b=b, iou=ans.numpy())
    print("iou_bev_cpu", ans.shape, float(ans.max()))


def gpu_fixtures(out):
    import torch

    pn = build_ref.load_ref("pointnet2_batch_cuda")
    iou = build_ref.load_ref("iou3d_nms_cuda")
    assert pn is not None and iou is not None, "oracle/_ref not built"
    dev = torch.device("cuda:0")

    def T(x):
        return torch.from_numpy(np.ascontiguousarray(x)).to(dev)

    # ---- FPS (farthest_point_sampling_wrapper) on uniform / duplicate-padded / lattice clouds
    fps = {}
    for name, xyz, m in [
        ("uniform", synth.cloud_uniform(2, 2048, seed=0), 256),
        ("dup", synth.cloud_dup_padded(2, 2048, seed=1), 512),
        ("lattice", synth.cloud_lattice(2, 1500, seed=2), 700),
        ("small", synth.cloud_dup_padded(3, 300, seed=3), 300),
        ("kitti", synth.cloud_ground_objects(1, 16384, seed=4), 4096),
    ]:
        b, n, _ = xyz.shape
        x = T(xyz)
        temp = torch.full((b, n), 1e10, device=dev)
        idx = torch.zeros((b, m), dtype=torch.int32, device=dev)
        pn.farthest_point_sampling_wrapper(b, n, m, x, temp, idx)
        torch.cuda.synchronize()
        fps[f"{name}_xyz"] = xyz
        fps[f"{name}_idx"] = idx.cpu().numpy()
        fps[f"{name}_temp"] = temp.cpu().numpy()
    np.savez_compressed(os.path.join(out, "fps.npz"), **fps)

    # ---- weighted FPS
    xyz = synth.cloud_dup_padded(2, 4096, seed=5)
    w = np.random.default_rng(6).uniform(0, 1, size=(2, 4096)).astype(np.float32) ** 2
    w[:, ::97] = 0.0  # exercises the max(w, 1e-12) clamp
    idx = torch.zeros((2, 512), dtype=torch.int32, device=dev)
    temp = torch.full((2, 4096), 1e10, device=dev)
    pn.furthest_point_sampling_weights_wrapper(2, 4096, 512, T(xyz), T(w), temp, idx)
    np.savez_compressed(os.path.join(out, "fps_weights.npz"), xyz=xyz, weights=w, idx=idx.cpu().numpy())

    # ---- matrix FPS (+ weighted)
    xyz = synth.cloud_dup_padded(2, 384, seed=7)
    mat = torch.cdist(T(xyz), T(xyz)).contiguous()
    idx = torch.zeros((2, 128), dtype=torch.int32, device=dev)
    temp = torch.full((2, 384), 1e10, device=dev)
    pn.furthest_point_sampling_matrix_wrapper(2, 384, 128, mat, temp, idx)
    w = np.random.default_rng(8).uniform(0, 1, size=(2, 384)).astype(np.float32)
    idx2 = torch.zeros((2, 128), dtype=torch.int32, device=dev)
    temp2 = torch.full((2, 384), 1e10, device=dev)
    pn.furthest_point_sampling_with_weighted_dist_wrapper(2, 384, 128, mat, T(w), temp2, idx2)
    np.savez_compressed(os.path.join(out, "fps_matrix.npz"), matrix=mat.cpu().numpy(), weights=w,
                        idx=idx.cpu().numpy(), idx_weighted=idx2.cpu().numpy())

    # ---- ball query (+ dilated)
    xyz = synth.cloud_ground_objects(2, 4096, seed=9)
    new_xyz = np.ascontiguousarray(xyz[:, ::8, :])
    bq = {"xyz": xyz, "new_xyz": new_xyz}
    for tag, r_in, r_out, ns in [("r08", None, 0.8, 32), ("r02", None, 0.2, 16), ("d0408", 0.4, 0.8, 32), ("r30", None, 3.0, 8)]:
        b, n, _ = xyz.shape
        m = new_xyz.shape[1]
        idx = torch.zeros((b, m, ns), dtype=torch.int32, device=dev)
        cnt = torch.zeros((b, m), dtype=torch.int32, device=dev)
        if r_in is None:
            pn.ball_query_wrapper(b, n, m, r_out, ns, T(new_xyz), T(xyz), cnt, idx)
        else:
            pn.ball_query_dilated_wrapper(b, n, m, r_in, r_out, ns, T(new_xyz), T(xyz), cnt, idx)
        bq[f"{tag}_idx"] = idx.cpu().numpy()
        bq[f"{tag}_cnt"] = cnt.cpu().numpy()
    np.savez_compressed(os.path.join(out, "ball_query.npz"), **bq)

    # ---- three_nn / three_interpolate
    unknown = synth.cloud_ground_objects(2, 1024, seed=10)
    known = np.ascontiguousarray(unknown[:, ::4, :]) + np.float32(0.01)
    known[:, 5] = known[:, 4]  # duplicate known point: tie in the cascade
    d2 = torch.zeros((2, 1024, 3), device=dev)
    idx = torch.zeros((2, 1024, 3), dtype=torch.int32, device=dev)
    pn.three_nn_wrapper(2, 1024, 256, T(unknown), T(known), d2, idx)
    feats = np.random.default_rng(11).normal(size=(2, 16, 256)).astype(np.float32)
    dist = torch.sqrt(d2)
    dr = 1.0 / (dist + 1e-8)
    wgt = (dr / dr.sum(2, keepdim=True)).contiguous()
    outp = torch.zeros((2, 16, 1024), device=dev)
    pn.three_interpolate_wrapper(2, 16, 256, 1024, T(feats), idx, wgt, outp)
    np.savez_compressed(os.path.join(out, "interpolate.npz"), unknown=unknown, known=known, dist2=d2.cpu().numpy(),
                        idx=idx.cpu().numpy(), feats=feats, weight=wgt.cpu().numpy(), out=outp.cpu().numpy())

    # ---- rotated IoU / overlap / NMS on the GPU
    a = np.concatenate([synth.boxes_random(64, seed=12), synth.boxes_clustered(192, seed=13, centres=24)], 0)
    ans = torch.zeros((256, 256), device=dev)
    iou.boxes_iou_bev_gpu(T(a), T(a), ans)
    ov = torch.zeros((256, 256), device=dev)
    iou.boxes_overlap_bev_gpu(T(a), T(a), ov)
    nms = {"boxes": a, "iou": ans.cpu().numpy(), "overlap": ov.cpu().numpy()}
    bx = synth.boxes_clustered(2048, seed=14, centres=150)
    sc = synth.scores_random(2048, seed=15)
    order = np.argsort(-sc, kind="stable")
    bs = np.ascontiguousarray(bx[order])
    nms["nms_boxes_sorted"] = bs
    for th in (0.01, 0.1, 0.5, 0.7):
        keep = torch.zeros(2048, dtype=torch.int64)
        k = iou.nms_gpu(T(bs), keep, th)
        nms[f"keep_{th}"] = keep[:k].numpy()
        keep = torch.zeros(2048, dtype=torch.int64)
        k = iou.nms_normal_gpu(T(bs), keep, th)
        nms[f"keepn_{th}"] = keep[:k].numpy()
    full = torch.zeros((2048, 2048), device=dev)
    iou.boxes_iou_bev_gpu(T(bs), T(bs), full)
    # store the IoU of the sorted boxes sparsely (non-zeros only) to keep the fixture small
    f = full.cpu().numpy()
    nz = np.nonzero(f)
    nms["full_nz_i"] = nz[0].astype(np.int32)
    nms["full_nz_j"] = nz[1].astype(np.int32)
    nms["full_nz_v"] = f[nz]
    np.savez_compressed(os.path.join(out, "iou_nms_gpu.npz"), **nms)
    print("gpu fixtures written to", out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.dirname(os.path.abspath(__file__)))
    ap.add_argument("--cpu-only", action="store_true")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    cpu_fixtures(args.out)
    if not args.cpu_only:
        gpu_fixtures(args.out)

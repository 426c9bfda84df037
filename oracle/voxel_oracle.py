"""CPU restatement (torch, CPU tensors) of the centroid voxelisation step after SA layer 0.

TEST INFRASTRUCTURE ONLY -- nothing on the product path imports this file.

Follows /root/reference/pcdet/utils/voxel_aggregation_utils.py:48-83 (get_voxel_indices), :132-161
(get_centroid_per_voxel), /root/reference/pcdet/utils/common_utils.py:248-265 (scatter_point_inds /
generate_voxel2pinds) and the call site /root/reference/pcdet/ops/pointnet2/pointnet2_batch/pointnet2_modules.py:
1323-1355.  Pinned: tests/golden/voxel_centroids.npz holds the outputs of the reference's OWN functions (imported
from /root/reference by tests/golden/make_golden_voxel.py, CPU tensors) and tests/test_oracle_cpu.py checks this
restatement against them bit for bit.
"""
from __future__ import annotations

import numpy as np
import torch


def get_voxel_indices(point_coords: torch.Tensor, voxel_size, point_cloud_range) -> torch.Tensor:
    vs = torch.as_tensor(voxel_size, dtype=torch.float32)
    r0 = torch.as_tensor(point_cloud_range, dtype=torch.float32)[0:3]
    return ((point_coords.float() - r0) / vs).long()


def get_centroid_per_voxel(points: torch.Tensor, voxel_idxs: torch.Tensor, num_points_in_voxel=None):
    """Sorted unique rows of voxel_idxs (lexicographic), inverse, counts; per-voxel sums accumulated sequentially in
    row order in fp32 (what torch's CPU scatter_add_ does), then one fp32 division."""
    v = voxel_idxs.numpy().astype(np.int64)
    p = points.numpy().astype(np.float32)
    order = np.lexsort(v.T[::-1], axis=0)             # stable: ties keep ascending row order
    vs = v[order]
    head = np.ones(len(vs), dtype=bool)
    head[1:] = (vs[1:] != vs[:-1]).any(axis=1)
    seg = np.cumsum(head) - 1
    nu = int(seg[-1]) + 1 if len(vs) else 0
    inverse = np.empty(len(vs), dtype=np.int64)
    inverse[order] = seg
    counts = np.bincount(seg, minlength=nu).astype(np.int64)
    sums = np.zeros((nu, p.shape[1]), dtype=np.float32)
    if num_points_in_voxel is None:
        for row in range(len(p)):                      # ascending row order, fp32 accumulation
            sums[inverse[row]] += p[row]
        cent = sums / counts.astype(np.float32)[:, None]
    else:
        w = num_points_in_voxel.numpy().astype(np.int64)
        wsum = np.zeros(nu, dtype=np.int64)
        for row in range(len(p)):
            sums[inverse[row]] += p[row] * np.float32(w[row])
            wsum[inverse[row]] += w[row]
        cent = sums / wsum.astype(np.float32)[:, None]
    return (torch.from_numpy(cent), torch.from_numpy(vs[head]), torch.from_numpy(counts), torch.from_numpy(inverse))


def voxelize_centroids(new_xyz: torch.Tensor, new_features: torch.Tensor, voxel_size, point_cloud_range):
    """pointnet2_modules.py:1323-1355 on CPU tensors: new_xyz (B,M,3), new_features (B,C,M)."""
    b, c, m = new_features.shape
    vi = get_voxel_indices(new_xyz.clone().view(-1, 3).contiguous(), voxel_size, point_cloud_range)
    batch_idx = torch.arange(b).view(b, 1).expand(b, m).reshape(-1, 1).long()
    voxel_idxs = torch.cat((batch_idx, torch.flip(vi, dims=[1])), dim=-1)
    xyz_for_voxel = torch.cat([batch_idx, new_xyz.view(-1, 3)], dim=-1)
    feats = new_features.permute(0, 2, 1).contiguous().reshape(b * m, c)
    point_for_voxel = torch.cat([xyz_for_voxel, feats], dim=-1)
    cent, cvi, counts, inverse = get_centroid_per_voxel(point_for_voxel, voxel_idxs)
    return {"voxel_idxs": voxel_idxs, "centroids_coords_features": cent, "centroid_voxel_idxs": cvi,
            "num_points_in_voxel": counts, "unique_idxs": inverse}


def generate_voxel2pinds(indices: torch.Tensor, batch_size: int, spatial_shape) -> torch.Tensor:
    out = -torch.ones((batch_size, *[int(s) for s in spatial_shape]), dtype=torch.int32)
    idx = indices.long()
    out[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]] = torch.arange(idx.shape[0], dtype=torch.int32)
    return out

"""Centroid voxelisation after SA layer 0 -- the functions of
``/root/reference/pcdet/utils/voxel_aggregation_utils.py`` that the layer-0 branch calls
(``get_voxel_indices`` :48-83, ``get_centroid_per_voxel`` :132-161), ``generate_voxel2pinds`` of
``pcdet/utils/common_utils.py:248-265``, and the whole tail of the branch as one call
(``voxelize_centroids`` = ``pointnet2_modules.py:1323-1355``), on the kernels of ``csrc/voxel_centroid.cu``.

The reference builds the result from ``torch.unique(dim=0, return_inverse, return_counts)`` (a device-wide
lexicographic sort), two ``scatter_add_`` with atomics and about ten small copies; here a frame's points are sorted
by voxel in shared memory by one CTA, and every voxel's points are added in ascending point order -- deterministic,
and bit-equal to the reference functions evaluated on the CPU (sequential ``scatter_add_``), which is what pins
the parity tests (``tests/golden/voxel_centroids.npz`` holds the reference functions' own outputs).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from ._lib import call, ptr, stream_ptr

_ERR_RANGE, _ERR_BATCH = 1, 2


def get_voxel_indices(point_coords: torch.Tensor, voxel_size, point_cloud_range) -> torch.Tensor:
    """ref :48-83 -- (N,3) points -> (N,3) int64 voxel indices in x,y,z order: ``((p - range[0:3]) / voxel_size).long()``
    (no range check, like the reference).  Plain elementwise arithmetic: stays in torch (the fused path below computes
    the same expression inside its sort kernel)."""
    assert point_coords.shape[1] == 3
    vs = torch.as_tensor(voxel_size, dtype=torch.float32, device=point_coords.device)
    r0 = torch.as_tensor(point_cloud_range, dtype=torch.float32, device=point_coords.device)[0:3]
    return ((point_coords - r0) / vs).long()


def _check(err: torch.Tensor, what: str) -> None:
    e = int(err.item())
    if e & _ERR_RANGE:
        raise ValueError(f"{what}: a voxel coordinate lies outside [-32768, 32767] (points far outside the range?)")
    if e & _ERR_BATCH:
        raise ValueError(f"{what}: rows must be grouped frame after frame, the same number of rows per frame")


def get_centroid_per_voxel(points: torch.Tensor, voxel_idxs: torch.Tensor, num_points_in_voxel: Optional[torch.Tensor] = None,
                           batch_size: Optional[int] = None):
    """ref :132-161 -- points (N, 4+f) [b,x,y,z,f...], voxel_idxs (N,4) int64 [b,z,y,x] ->
    ``(centroids (N',4+f), centroid_voxel_idxs (N',4), labels_count (N'), unique_idxs (N))``: the sorted unique voxels,
    per-voxel means (weighted by ``num_points_in_voxel`` when given), point counts and the inverse mapping.
    Rows must be grouped frame after frame with equal counts (what the reference's call sites pass, <= 16384 rows per
    frame); ``batch_size`` saves the host read of the last row's batch index."""
    assert points.shape[0] == voxel_idxs.shape[0] and voxel_idxs.shape[1] == 4
    assert points.is_cuda and voxel_idxs.is_cuda, "tsmdet_b200 ops need CUDA tensors (there is no CPU fallback)"
    n, w = points.shape
    dev = points.device
    if n == 0:
        return (points.new_zeros((0, w)), voxel_idxs.new_zeros((0, 4)), voxel_idxs.new_zeros((0,)),
                voxel_idxs.new_zeros((0,)))
    if batch_size is None:
        batch_size = int(voxel_idxs[-1, 0].item()) + 1
    assert n % batch_size == 0, "every frame must hold the same number of rows"
    m = n // batch_size
    points = points.contiguous().float()
    voxel_idxs = voxel_idxs.contiguous().long()
    weights = None if num_points_in_voxel is None else num_points_in_voxel.contiguous().long()
    centroids = torch.empty((n, w), dtype=torch.float32, device=dev)
    cvi = torch.empty((n, 4), dtype=torch.int64, device=dev)
    counts = torch.empty((n,), dtype=torch.int64, device=dev)
    inverse = torch.empty((n,), dtype=torch.int64, device=dev)
    num = torch.zeros((1,), dtype=torch.int32, device=dev)
    err = torch.zeros((1,), dtype=torch.int32, device=dev)
    call("tsmdet_centroid_per_voxel", batch_size, m, w - 4, ptr(points), ptr(voxel_idxs), ptr(weights), ptr(centroids),
         ptr(cvi), ptr(counts), ptr(inverse), ptr(num), ptr(err), stream_ptr(dev))
    k = int(num.item())  # the one host read (torch.unique synchronises for its output size too)
    _check(err, "get_centroid_per_voxel")
    return centroids[:k], cvi[:k], counts[:k], inverse


def voxelize_centroids(new_xyz: torch.Tensor, new_features: Optional[torch.Tensor], voxel_size: Sequence[float],
                       point_cloud_range: Sequence[float]):
    """The tail of the layer-0 branch (pointnet2_modules.py:1323-1355) in one call: new_xyz (B,M,3), new_features
    (B,C,M) -> dict with ``voxel_idxs (B*M,4) [b,z,y,x]``, ``centroids (N',4) [b,x,y,z]``, ``centroids_features (N',C)``,
    ``centroid_voxel_idxs (N',4)`` (``.int()`` of it = SparseConvTensor.indices), ``num_points_in_voxel (N')``,
    ``unique_idxs (B*M)``.  None of the reference's intermediate copies (clone/view/flip/cat/permute) is made."""
    b, m, _ = new_xyz.shape
    c = 0 if new_features is None else new_features.shape[1]
    assert new_xyz.is_cuda and new_xyz.is_contiguous() and (new_features is None or new_features.is_contiguous())
    dev = new_xyz.device
    n = b * m
    voxel_idxs = torch.empty((n, 4), dtype=torch.int64, device=dev)
    rows = torch.empty((n, 4 + c), dtype=torch.float32, device=dev)
    cvi = torch.empty((n, 4), dtype=torch.int64, device=dev)
    counts = torch.empty((n,), dtype=torch.int64, device=dev)
    inverse = torch.empty((n,), dtype=torch.int64, device=dev)
    num = torch.zeros((1,), dtype=torch.int32, device=dev)
    err = torch.zeros((1,), dtype=torch.int32, device=dev)
    vs = [float(v) for v in voxel_size]
    r0 = [float(v) for v in point_cloud_range[0:3]]
    # the reference holds both as float32 tensors: round the Python doubles the same way
    vs = torch.tensor(vs, dtype=torch.float32).tolist()
    r0 = torch.tensor(r0, dtype=torch.float32).tolist()
    call("tsmdet_voxel_centroids", b, m, c, ptr(new_xyz), ptr(new_features), vs[0], vs[1], vs[2], r0[0], r0[1], r0[2],
         ptr(voxel_idxs), ptr(rows), ptr(cvi), ptr(counts), ptr(inverse), ptr(num), ptr(err), stream_ptr(dev))
    k = int(num.item())
    _check(err, "voxelize_centroids")
    return {"voxel_idxs": voxel_idxs, "centroids": rows[:k, 0:4].contiguous(), "centroids_features": rows[:k, 4:],
            "centroids_coords_features": rows[:k], "centroid_voxel_idxs": cvi[:k], "num_points_in_voxel": counts[:k],
            "unique_idxs": inverse}


class Voxel2PointIndex:
    """``generate_voxel2pinds`` (common_utils.py:248-265) with a persistent table: the reference allocates and fills a
    dense (B,Z,Y,X) int32 tensor per call (369 MB per KITTI frame at full resolution); here the table lives on, and an
    update resets only the entries the previous call set before writing the new ones."""

    def __init__(self, batch_size: int, spatial_shape: Sequence[int], device):
        self.shape = (int(batch_size), *[int(v) for v in spatial_shape])
        self.table = torch.empty(self.shape, dtype=torch.int32, device=device)
        self._prev = None

    def update(self, indices: torch.Tensor) -> torch.Tensor:
        """indices (n,4) int32 [b,z,y,x] (SparseConvTensor.indices) -> the dense table (valid until the next update)."""
        assert indices.is_cuda and indices.dtype == torch.int32 and indices.shape[1] == 4
        indices = indices.contiguous()
        err = torch.zeros((1,), dtype=torch.int32, device=indices.device)
        nb, nz, ny, nx = self.shape
        prev = self._prev
        call("tsmdet_voxel2pinds", indices.shape[0], ptr(indices), 0 if prev is None else prev.shape[0], ptr(prev), nb, nz, ny,
             nx, ptr(self.table), ptr(err), stream_ptr(indices.device))
        self._prev = indices
        self._err = err
        return self.table


def generate_voxel2pinds(indices: torch.Tensor, batch_size: int, spatial_shape: Sequence[int]) -> torch.Tensor:
    """ref common_utils.py:257-265 on (indices, batch_size, spatial_shape) of a SparseConvTensor: a fresh dense table."""
    return Voxel2PointIndex(batch_size, spatial_shape, indices.device).update(indices.int())

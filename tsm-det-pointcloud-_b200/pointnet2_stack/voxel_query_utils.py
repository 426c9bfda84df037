"""``VoxelQuery`` / ``VoxelQueryDilated`` and the ``VoxelQueryAndGrouping(Dilated)`` modules of
``/root/reference/pcdet/ops/pointnet2/pointnet2_stack/voxel_query_utils.py`` (:11-114, 117-236) on the sm_100a kernels.
Same signatures and returns; the one deliberate difference is inside ``VoxelQueryAndGrouping*.forward``: the reference
converts global point rows to frame-local ones with a Python loop over frames that reads ``xyz_batch_cnt`` on the host
(:85-90, 205-210); here the same subtraction is one broadcast, no host synchronisation."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function

from . import pointnet2_stack_cuda as pointnet2
from . import pointnet2_utils


class VoxelQuery(Function):
    """ref :11-51"""

    @staticmethod
    def forward(ctx, max_range, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor,
                new_coords: torch.Tensor, point_indices: torch.Tensor):
        assert new_xyz.is_contiguous()
        assert xyz.is_contiguous()
        assert new_coords.is_contiguous()
        assert point_indices.is_contiguous()
        M = new_coords.shape[0]
        B, Z, Y, X = point_indices.shape
        idx = torch.zeros((M, nsample), dtype=torch.int32, device=xyz.device)
        cnt_unique = torch.zeros((M, 1), dtype=torch.int32, device=xyz.device)
        z_range, y_range, x_range = max_range
        pointnet2.voxel_query_wrapper(M, Z, Y, X, nsample, radius, z_range, y_range, x_range, new_xyz, xyz, new_coords,
                                      point_indices, idx, cnt_unique)
        empty_ball_mask = (idx[:, 0] == -1)
        idx[empty_ball_mask] = 0
        volume = (x_range * 2 + 1) * (y_range * 2 + 1) * (z_range * 2 + 1)
        density = cnt_unique / volume
        return idx, empty_ball_mask, density

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


voxel_query = VoxelQuery.apply


class VoxelQueryDilated(Function):
    """ref :117-163"""

    @staticmethod
    def forward(ctx, max_range, stride, former_radius: float, radius: float, nsample: int, xyz: torch.Tensor,
                new_xyz: torch.Tensor, new_coords: torch.Tensor, point_indices: torch.Tensor):
        assert new_xyz.is_contiguous()
        assert xyz.is_contiguous()
        assert new_coords.is_contiguous()
        assert point_indices.is_contiguous()
        M = new_coords.shape[0]
        B, Z, Y, X = point_indices.shape
        idx = torch.zeros((M, nsample), dtype=torch.int32, device=xyz.device)
        cnt_unique = torch.zeros((M, 1), dtype=torch.int32, device=xyz.device)
        idx_cnt = torch.zeros((M, 1), dtype=torch.int32, device=xyz.device)
        z_range, y_range, x_range = max_range
        z_stride, y_stride, x_stride = stride
        pointnet2.voxel_query_dilated_wrapper(M, Z, Y, X, nsample, former_radius, radius, z_range, y_range, x_range,
                                              z_stride, y_stride, x_stride, new_xyz, xyz, new_coords, point_indices, idx,
                                              cnt_unique, idx_cnt)
        empty_ball_mask = (idx[:, 0] == -1)
        idx[empty_ball_mask] = 0
        density = cnt_unique / nsample
        density_score = torch.clamp(density, max=1.0)
        return idx, empty_ball_mask, density_score

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


voxel_query_dilated = VoxelQueryDilated.apply


def _to_frame_local(idx: torch.Tensor, xyz_batch_cnt: torch.Tensor, batch_size: int, nsample: int, empty_ball_mask):
    """idx1.view(B,-1,nsample)[b] -= sum(xyz_batch_cnt[:b]); idx1[empty] = 0  (ref :83-91) without the host loop."""
    starts = torch.cumsum(xyz_batch_cnt.to(torch.int64), 0) - xyz_batch_cnt.to(torch.int64)
    idx = (idx.view(batch_size, -1, nsample) - starts.view(batch_size, 1, 1).to(idx.dtype)).view(-1, nsample)
    idx[empty_ball_mask] = 0
    return idx.contiguous()


class VoxelQueryAndGrouping(nn.Module):
    """ref :54-114"""

    def __init__(self, max_range, radius: float, nsample: int):
        super().__init__()
        self.max_range, self.radius, self.nsample = max_range, radius, nsample

    def forward(self, new_coords: torch.Tensor, xyz: torch.Tensor, xyz_batch_cnt: torch.Tensor, new_xyz: torch.Tensor,
                new_xyz_batch_cnt: torch.Tensor, features: torch.Tensor, voxel2point_indices: torch.Tensor):
        """-> (grouped_features (M,C,nsample), grouped_xyz (M,3,nsample), empty_ball_mask (M), density (M,1))"""
        assert xyz.shape[0] == xyz_batch_cnt.sum(), 'xyz: %s, xyz_batch_cnt: %s' % (str(xyz.shape), str(new_xyz_batch_cnt))
        assert new_coords.shape[0] == new_xyz_batch_cnt.sum(), \
            'new_coords: %s, new_xyz_batch_cnt: %s' % (str(new_coords.shape), str(new_xyz_batch_cnt))
        batch_size = xyz_batch_cnt.shape[0]
        idx1, empty_ball_mask1, density_score = voxel_query(self.max_range, self.radius, self.nsample, xyz, new_xyz,
                                                            new_coords, voxel2point_indices)
        idx1 = _to_frame_local(idx1, xyz_batch_cnt, batch_size, self.nsample, empty_ball_mask1)
        grouped_xyz = pointnet2_utils.grouping_operation(xyz, xyz_batch_cnt, idx1, new_xyz_batch_cnt)
        grouped_features = pointnet2_utils.grouping_operation(features, xyz_batch_cnt, idx1, new_xyz_batch_cnt)
        return grouped_features, grouped_xyz, empty_ball_mask1, density_score


class VoxelQueryAndGroupingDilated(nn.Module):
    """ref :166-236"""

    def __init__(self, max_range, stride, former_radius: float, radius: float, nsample: int):
        super().__init__()
        self.max_range, self.stride, self.former_radius, self.radius, self.nsample = \
            max_range, stride, former_radius, radius, nsample

    def forward(self, new_coords: torch.Tensor, xyz: torch.Tensor, xyz_batch_cnt: torch.Tensor, new_xyz: torch.Tensor,
                new_xyz_batch_cnt: torch.Tensor, features: torch.Tensor, voxel2point_indices: torch.Tensor):
        assert xyz.shape[0] == xyz_batch_cnt.sum(), 'xyz: %s, xyz_batch_cnt: %s' % (str(xyz.shape), str(new_xyz_batch_cnt))
        assert new_coords.shape[0] == new_xyz_batch_cnt.sum(), \
            'new_coords: %s, new_xyz_batch_cnt: %s' % (str(new_coords.shape), str(new_xyz_batch_cnt))
        batch_size = xyz_batch_cnt.shape[0]
        idx1, empty_ball_mask1, density_score = voxel_query_dilated(
            self.max_range, self.stride, self.former_radius, self.radius, self.nsample, xyz, new_xyz, new_coords,
            voxel2point_indices)
        idx1 = _to_frame_local(idx1, xyz_batch_cnt, batch_size, self.nsample, empty_ball_mask1)
        grouped_xyz = pointnet2_utils.grouping_operation(xyz, xyz_batch_cnt, idx1, new_xyz_batch_cnt)
        grouped_features = pointnet2_utils.grouping_operation(features, xyz_batch_cnt, idx1, new_xyz_batch_cnt)
        return grouped_features, grouped_xyz, empty_ball_mask1, density_score

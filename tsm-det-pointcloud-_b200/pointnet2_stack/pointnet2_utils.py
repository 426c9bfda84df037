"""``GroupingOperation`` / ``StackFarthestPointSampling`` / ``FarthestPointSampling`` of
``/root/reference/pcdet/ops/pointnet2/pointnet2_stack/pointnet2_utils.py`` (:48-105, 158-222) on the sm_100a kernels:
same signatures, argument checks and returns (the ball-query, 3-NN and vector-pool ops of that file are not on the
path of the shipped model and are not part of this package)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import pointnet2_stack_cuda as pointnet2


class GroupingOperation(Function):
    """ref :48-103"""

    @staticmethod
    def forward(ctx, features: torch.Tensor, features_batch_cnt: torch.Tensor, idx: torch.Tensor,
                idx_batch_cnt: torch.Tensor):
        """features (N1+N2..., C), features_batch_cnt (B), idx (M1+M2..., nsample) frame-local rows, idx_batch_cnt (B)
        -> (M1+M2..., C, nsample)"""
        assert features.is_contiguous()
        assert features_batch_cnt.is_contiguous()
        assert idx.is_contiguous()
        assert idx_batch_cnt.is_contiguous()
        assert features.shape[0] == features_batch_cnt.sum(), \
            'features: %s, features_batch_cnt: %s' % (str(features.shape), str(features_batch_cnt))
        assert idx.shape[0] == idx_batch_cnt.sum(), 'idx: %s, idx_batch_cnt: %s' % (str(idx.shape), str(idx_batch_cnt))
        M, nsample = idx.size()
        N, C = features.size()
        B = idx_batch_cnt.shape[0]
        output = torch.empty((M, C, nsample), dtype=torch.float32, device=features.device)
        pointnet2.group_points_wrapper(B, M, C, nsample, features, features_batch_cnt, idx, idx_batch_cnt, output)
        ctx.for_backwards = (B, N, idx, features_batch_cnt, idx_batch_cnt)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        B, N, idx, features_batch_cnt, idx_batch_cnt = ctx.for_backwards
        M, C, nsample = grad_out.size()
        grad_features = torch.zeros((N, C), dtype=torch.float32, device=grad_out.device)
        pointnet2.group_points_grad_wrapper(B, M, C, N, nsample, grad_out.data.contiguous(), idx, idx_batch_cnt,
                                            features_batch_cnt, grad_features)
        return grad_features, None, None, None


grouping_operation = GroupingOperation.apply


class FarthestPointSampling(Function):
    """ref :158-184 -- the dense-batch sampler under its stack-module name"""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int):
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        output = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        pointnet2.farthest_point_sampling_wrapper(B, N, npoint, xyz, temp, output)
        return output

    @staticmethod
    def backward(xyz, a=None):
        return None, None


farthest_point_sample = furthest_point_sample = FarthestPointSampling.apply


class StackFarthestPointSampling(Function):
    """ref :187-219"""

    @staticmethod
    def forward(ctx, xyz, xyz_batch_cnt, npoint):
        """xyz (N1+N2..., 3), xyz_batch_cnt [N1, N2, ...], npoint int | list | tensor -> (sum npoint) int32 GLOBAL rows"""
        assert xyz.is_contiguous() and xyz.shape[1] == 3
        batch_size = xyz_batch_cnt.__len__()
        if not isinstance(npoint, torch.Tensor):
            if not isinstance(npoint, list):
                npoint = [npoint for i in range(batch_size)]
            npoint = torch.tensor(npoint, device=xyz.device).int()
        N, _ = xyz.size()
        temp = torch.full((N,), 1e10, dtype=torch.float32, device=xyz.device)
        output = torch.empty((int(npoint.sum().item()),), dtype=torch.int32, device=xyz.device)
        pointnet2.stack_farthest_point_sampling_wrapper(xyz, temp, xyz_batch_cnt.int().contiguous(), output,
                                                        npoint.contiguous())
        return output

    @staticmethod
    def backward(xyz, a=None):
        return None, None


stack_farthest_point_sample = StackFarthestPointSampling.apply

"""Python op surface of the dense-batch PointNet++ ops -- same names, arguments and results as
``/root/reference/pcdet/ops/pointnet2/pointnet2_batch/pointnet2_utils.py``, on the
sm_100a kernels behind ``libtsmdet_b200.so``.

Kept from the reference, including its quirks:
  * index tensors are ``torch.int32``; inputs must be contiguous (asserts as in :97, :234-235, ...);
  * ``farthest_point_sample`` and ``furthest_point_sample`` are the same ``Function.apply``
    (:111 rebinds the name defined at :21);
  * ``QueryAndGroup*.forward`` return the 3-tuple ``(idx_cnt, new_features, grouped_xyz)`` (:530, :568);
  * ``ThreeNN`` returns ``sqrt(dist2)`` (:282).
Changed on purpose: outputs are allocated on the input's device (not "the current CUDA device"),
kernels run on the current stream, and QueryAndGroup materialises its outputs with one fused
kernel instead of transpose + 2 gathers + subtract + cat.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from . import pointnet2_batch_cuda as pointnet2
from ._lib import call, ptr, stream_ptr


def _i32(shape, device):
    return torch.empty(shape, dtype=torch.int32, device=device)


def _f32(shape, device):
    return torch.empty(shape, dtype=torch.float32, device=device)


def _scratch_min_dist(b: int, n: int, device):
    """``temp`` of the reference callers: (B,N) filled with 1e10 (:35, :101)."""
    return torch.full((b, n), 1e10, dtype=torch.float32, device=device)


@torch.no_grad()
def calc_dist_matrix_for_sampling(xyz: torch.Tensor, features: torch.Tensor = None, gamma: float = 1.0):
    """ref :9-17 -- pairwise distances for 'f-fps' (torch.cdist, optionally + gamma * feature distance)."""
    dist = torch.cdist(xyz, xyz)
    if features is not None:
        dist += torch.cdist(features, features) * gamma
    return dist


@torch.no_grad()
def furthest_point_sample_matrix(matrix: torch.Tensor, npoint: int) -> torch.Tensor:
    """ref :41-58 -- FPS over a (B,N,N) distance matrix -> (B,npoint) int32."""
    assert matrix.is_contiguous()
    b, n, _ = matrix.size()
    out = _i32((b, npoint), matrix.device)
    temp = _scratch_min_dist(b, n, matrix.device)
    pointnet2.furthest_point_sampling_matrix_wrapper(b, n, npoint, matrix, temp, out)
    return out


@torch.no_grad()
def furthest_point_sample_weights(xyz: torch.Tensor, weights: torch.Tensor, npoint: int) -> torch.Tensor:
    """ref :61-81 -- score-weighted FPS ('s-fps'): xyz (B,N,3), weights (B,N) -> (B,npoint) int32."""
    assert xyz.is_contiguous()
    assert weights.is_contiguous()
    b, n, _ = xyz.size()
    out = _i32((b, npoint), xyz.device)
    temp = _scratch_min_dist(b, n, xyz.device)
    pointnet2.furthest_point_sampling_weights_wrapper(b, n, npoint, xyz, weights, temp, out)
    return out


class FarthestPointSampling(Function):
    """ref :85-108"""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        assert xyz.is_contiguous()
        b, n, _ = xyz.size()
        out = _i32((b, npoint), xyz.device)
        temp = _scratch_min_dist(b, n, xyz.device)
        pointnet2.farthest_point_sampling_wrapper(b, n, npoint, xyz, temp, out)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None


farthest_point_sample = furthest_point_sample = FarthestPointSampling.apply


class FpsChainState:
    """What an FPS run leaves behind for an FPS over its own output (see ``tsmdet_fps_chain``)."""

    __slots__ = ("tie_iter", "vals", "m")

    def __init__(self, tie_iter, vals, m):
        self.tie_iter, self.vals, self.m = tie_iter, vals, m


@torch.no_grad()
def farthest_point_sample_chained(xyz: torch.Tensor, npoint: int, parent: "FpsChainState" = None):
    """``farthest_point_sample`` for stacked SA layers: returns ``(idx, state)``.

    Pass the previous layer's ``state`` when ``xyz`` is exactly that layer's sampled centres in sampling
    order (``new_xyz``); the result is bit-identical to ``farthest_point_sample(xyz, npoint)`` -- FPS over
    the first picks of an FPS run re-selects them in order -- and costs microseconds whenever the parent
    run recorded no tie between distinct points (otherwise the full algorithm runs for that cloud)."""
    assert xyz.is_contiguous()
    b, n, _ = xyz.size()
    out = _i32((b, npoint), xyz.device)
    tie = _i32((b,), xyz.device)
    vals = _f32((b, npoint), xyz.device)
    use_parent = parent is not None and parent.m >= n
    call("tsmdet_fps_chain", b, n, npoint, ptr(xyz), None, ptr(out), ptr(tie), ptr(vals),
         ptr(parent.tie_iter) if use_parent else None, ptr(parent.vals) if use_parent else None,
         parent.m if use_parent else 0, stream_ptr(xyz.device))
    return out, FpsChainState(tie, vals, npoint)


class FurthestPointSamplingWithDist(Function):
    """ref :114-137 -- xyz here is a (B,N,N) distance matrix."""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        assert xyz.is_contiguous()
        b, n, _ = xyz.size()
        out = _i32((b, npoint), xyz.device)
        temp = _scratch_min_dist(b, n, xyz.device)
        pointnet2.furthest_point_sampling_with_dist_wrapper(b, n, npoint, xyz, temp, out)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample_with_dist = FurthestPointSamplingWithDist.apply


class FurthestPointSamplingWithWeightedDist(Function):
    """ref :143-166"""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, weights: torch.Tensor, npoint: int) -> torch.Tensor:
        assert xyz.is_contiguous()
        b, n, _ = xyz.size()
        out = _i32((b, npoint), xyz.device)
        temp = _scratch_min_dist(b, n, xyz.device)
        pointnet2.furthest_point_sampling_with_weighted_dist_wrapper(b, n, npoint, xyz, weights.contiguous(), temp, out)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None


furthest_point_sample_with_weighted_dist = FurthestPointSamplingWithWeightedDist.apply


class GatherOperation(Function):
    """ref :223-254 -- features (B,C,N), idx (B,npoint) -> (B,C,npoint)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        b, npoint = idx.size()
        _, c, n = features.size()
        out = _f32((b, c, npoint), features.device)
        pointnet2.gather_points_wrapper(b, c, n, npoint, features, idx, out)
        ctx.for_backwards = (idx, c, n)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, c, n = ctx.for_backwards
        b, npoint = idx.size()
        grad_features = torch.zeros((b, c, n), dtype=torch.float32, device=grad_out.device)
        pointnet2.gather_points_grad_wrapper(b, c, n, npoint, grad_out.data.contiguous(), idx, grad_features)
        return grad_features, None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    """ref :260-286 -- unknown (B,N,3), known (B,M,3) -> (dist (B,N,3) = sqrt(dist2), idx (B,N,3) int32)."""

    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert unknown.is_contiguous()
        assert known.is_contiguous()
        b, n, _ = unknown.size()
        m = known.size(1)
        dist2 = _f32((b, n, 3), unknown.device)
        idx = _i32((b, n, 3), unknown.device)
        pointnet2.three_nn_wrapper(b, n, m, unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """ref :292-334 -- features (B,C,M), idx (B,n,3), weight (B,n,3) -> (B,C,n)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        assert weight.is_contiguous()
        b, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        out = _f32((b, c, n), features.device)
        pointnet2.three_interpolate_wrapper(b, c, m, n, features, idx, weight, out)
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.three_interpolate_for_backward
        b, c, n = grad_out.size()
        grad_features = torch.zeros((b, c, m), dtype=torch.float32, device=grad_out.device)
        pointnet2.three_interpolate_grad_wrapper(b, c, n, m, grad_out.data.contiguous(), idx, weight, grad_features)
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    """ref :340-378 -- features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        b, nfeatures, nsample = idx.size()
        _, c, n = features.size()
        out = _f32((b, c, nfeatures, nsample), features.device)
        pointnet2.group_points_wrapper(b, c, n, nfeatures, nsample, features, idx, out)
        ctx.for_backwards = (idx, n)
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, n = ctx.for_backwards
        b, c, npoint, nsample = grad_out.size()
        grad_features = torch.zeros((b, c, n), dtype=torch.float32, device=grad_out.device)
        pointnet2.group_points_grad_wrapper(b, c, n, npoint, nsample, grad_out.data.contiguous(), idx, grad_features)
        return grad_features, None


grouping_operation = GroupingOperation.apply


@torch.no_grad()
def ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor):
    """ref :413-433 -> (idx_cnt (B,npoint) int32, idx (B,npoint,nsample) int32)."""
    assert new_xyz.is_contiguous()
    assert xyz.is_contiguous()
    b, n, _ = xyz.size()
    npoint = new_xyz.size(1)
    idx = torch.zeros((b, npoint, nsample), dtype=torch.int32, device=xyz.device)
    idx_cnt = torch.zeros((b, npoint), dtype=torch.int32, device=xyz.device)
    pointnet2.ball_query_wrapper(b, n, npoint, radius, nsample, new_xyz, xyz, idx_cnt, idx)
    return idx_cnt, idx


@torch.no_grad()
def ball_query_dilated(radius_in: float, radius_out: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor):
    """ref :436-457"""
    assert new_xyz.is_contiguous()
    assert xyz.is_contiguous()
    b, n, _ = xyz.size()
    npoint = new_xyz.size(1)
    idx_cnt = torch.zeros((b, npoint), dtype=torch.int32, device=xyz.device)
    idx = torch.zeros((b, npoint, nsample), dtype=torch.int32, device=xyz.device)
    pointnet2.ball_query_dilated_wrapper(b, n, npoint, radius_in, radius_out, nsample, new_xyz, xyz, idx_cnt, idx)
    return idx_cnt, idx


class _GroupConcat(Function):
    """Fused materialisation of (new_features, grouped_xyz) from idx; differentiable w.r.t. features
    (the reference gets the same gradient through GroupingOperation.backward + cat)."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, features, idx, use_xyz: bool):
        b, n, _ = xyz.size()
        _, m, s = idx.size()
        c = 0 if features is None else features.size(1)
        ctot = (3 if use_xyz else 0) + c
        new_features = _f32((b, ctot, m, s), xyz.device)
        grouped_xyz = _f32((b, 3, m, s), xyz.device)
        call("tsmdet_group_concat", b, c, n, m, s, int(use_xyz), ptr(xyz), ptr(new_xyz), ptr(features), ptr(idx),
             ptr(new_features), ptr(grouped_xyz), stream_ptr(xyz.device))
        ctx.saved = (idx, n, c, use_xyz, features is not None)
        ctx.mark_non_differentiable(grouped_xyz)
        return new_features, grouped_xyz

    @staticmethod
    def backward(ctx, g_feat, g_xyz=None):
        idx, n, c, use_xyz, has_f = ctx.saved
        g_features = None
        if has_f and ctx.needs_input_grad[2]:
            b, _, m, s = g_feat.size()
            gf = g_feat[:, (3 if use_xyz else 0):].contiguous()
            g_features = torch.zeros((b, c, n), dtype=torch.float32, device=g_feat.device)
            pointnet2.group_points_grad_wrapper(b, c, n, m, s, gf, idx, g_features)
        return None, None, g_features, None, None


def _query_and_group(idx_cnt, idx, xyz, new_xyz, features, use_xyz):
    if features is not None:
        assert features.is_contiguous()
    else:
        assert use_xyz, "Cannot have not features and not use xyz as a feature!"
    new_features, grouped_xyz = _GroupConcat.apply(xyz, new_xyz, features, idx, bool(use_xyz or features is None))
    return idx_cnt, new_features, grouped_xyz


class QueryAndGroup(nn.Module):
    """ref :496-530 -- forward(xyz (B,N,3), new_xyz (B,npoint,3), features (B,C,N)|None)
    -> (idx_cnt (B,npoint), new_features (B,3+C,npoint,nsample), grouped_xyz (B,3,npoint,nsample))."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        idx_cnt, idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        return _query_and_group(idx_cnt, idx, xyz, new_xyz, features, self.use_xyz)


class QueryAndGroupDilated(nn.Module):
    """ref :533-568 -- annulus query radius_in <= d < radius_out."""

    def __init__(self, radius_in: float, radius_out: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius_in, self.radius_out, self.nsample, self.use_xyz = radius_in, radius_out, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        idx_cnt, idx = ball_query_dilated(self.radius_in, self.radius_out, self.nsample, xyz, new_xyz)
        return _query_and_group(idx_cnt, idx, xyz, new_xyz, features, self.use_xyz)


class GroupAll(nn.Module):
    """ref :571-594 -- pure view/cat logic, no kernel."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            if self.use_xyz:
                new_features = torch.cat([grouped_xyz, grouped_features], dim=1)  # (B, 3 + C, 1, N)
            else:
                new_features = grouped_features
        else:
            new_features = grouped_xyz
        return new_features

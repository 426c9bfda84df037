"""Native surface of the reference's ``pointnet2_batch_cuda`` pybind module, re-hosted on the C ABI.

Same function names, positional arguments and return values as
``/root/reference/pcdet/ops/pointnet2/pointnet2_batch/src/pointnet2_api.cpp:10-33`` so that the
reference's own ``pointnet2_utils.py`` works unchanged when this module is bound in place of the
compiled extension (``from . import pointnet2_batch_cuda as pointnet2``).  Differences, all
deliberate: kernels run on torch's CURRENT stream (the reference uses the legacy default
stream), and bad inputs raise instead of ``exit(-1)`` (ball_query.cpp:20-32).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def _chk(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise ValueError("must be a CUDA tensor")
        if not t.is_contiguous():
            raise ValueError("must be a contiguous tensor")


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx_cnt, idx):
    """ref: ball_query.cpp:47-58"""
    _chk(new_xyz, xyz)
    call("tsmdet_ball_query", b, n, m, float(radius), nsample, ptr(new_xyz), ptr(xyz), ptr(idx_cnt), ptr(idx),
         stream_ptr(xyz.device))
    return 1


def ball_query_dilated_wrapper(b, n, m, radius_in, radius_out, nsample, new_xyz, xyz, idx_cnt, idx):
    """ref: ball_query.cpp:60-71"""
    _chk(new_xyz, xyz)
    call("tsmdet_ball_query_dilated", b, n, m, float(radius_in), float(radius_out), nsample, ptr(new_xyz), ptr(xyz),
         ptr(idx_cnt), ptr(idx), stream_ptr(xyz.device))
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    """ref: group_points.cpp:27-41"""
    call("tsmdet_group_points", b, c, n, npoints, nsample, ptr(points), ptr(idx), ptr(out), stream_ptr(points.device))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    """ref: group_points.cpp:13-24"""
    call("tsmdet_group_points_grad", b, c, n, npoints, nsample, ptr(grad_out), ptr(idx), ptr(grad_points),
         stream_ptr(grad_out.device))
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    """ref: sampling.cpp:17-27"""
    call("tsmdet_gather_points", b, c, n, npoints, ptr(points), ptr(idx), ptr(out), stream_ptr(points.device))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    """ref: sampling.cpp:30-40"""
    call("tsmdet_gather_points_grad", b, c, n, npoints, ptr(grad_out), ptr(idx), ptr(grad_points),
         stream_ptr(grad_out.device))
    return 1


def farthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    """ref: sampling.cpp:43-52"""
    call("tsmdet_farthest_point_sampling", b, n, m, ptr(points), ptr(temp), ptr(idx), stream_ptr(points.device))
    return 1


furthest_point_sampling_wrapper = farthest_point_sampling_wrapper  # sampling.cpp:58-67: same algorithm


def furthest_point_sampling_with_dist_wrapper(b, n, m, points, temp, idx):
    """ref: sampling.cpp:70-80 (returns 2 there)"""
    call("tsmdet_furthest_point_sampling_matrix", b, n, m, ptr(points), ptr(temp), ptr(idx), stream_ptr(points.device))
    return 2


def furthest_point_sampling_matrix_wrapper(b, n, m, matrix, temp, idx):
    """ref: sampling.cpp:98-109"""
    call("tsmdet_furthest_point_sampling_matrix", b, n, m, ptr(matrix), ptr(temp), ptr(idx), stream_ptr(matrix.device))
    return 1


def furthest_point_sampling_with_weighted_dist_wrapper(b, n, m, points, weights, temp, idx):
    """ref: sampling.cpp:83-95 (returns 2 there)"""
    call("tsmdet_furthest_point_sampling_with_weighted_dist", b, n, m, ptr(points), ptr(weights), ptr(temp), ptr(idx),
         stream_ptr(points.device))
    return 2


def furthest_point_sampling_weights_wrapper(b, n, m, xyz, weights, temp, idx):
    """ref: sampling.cpp:112-122"""
    call("tsmdet_furthest_point_sampling_weights", b, n, m, ptr(xyz), ptr(weights), ptr(temp), ptr(idx),
         stream_ptr(xyz.device))
    return 1


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    """ref: interpolate.cpp:20-30"""
    call("tsmdet_three_nn", b, n, m, ptr(unknown), ptr(known), ptr(dist2), ptr(idx), stream_ptr(unknown.device))


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    """ref: interpolate.cpp:33-44"""
    call("tsmdet_three_interpolate", b, c, m, n, ptr(points), ptr(idx), ptr(weight), ptr(out),
         stream_ptr(points.device))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    """ref: interpolate.cpp:47-60"""
    call("tsmdet_three_interpolate_grad", b, c, n, m, ptr(grad_out), ptr(idx), ptr(weight), ptr(grad_points),
         stream_ptr(grad_out.device))


__all__ = [
    "ball_query_wrapper", "ball_query_dilated_wrapper", "group_points_wrapper", "group_points_grad_wrapper",
    "gather_points_wrapper", "gather_points_grad_wrapper", "farthest_point_sampling_wrapper",
    "furthest_point_sampling_with_dist_wrapper", "furthest_point_sampling_with_weighted_dist_wrapper",
    "furthest_point_sampling_wrapper", "furthest_point_sampling_matrix_wrapper",
    "furthest_point_sampling_weights_wrapper", "three_nn_wrapper", "three_interpolate_wrapper",
    "three_interpolate_grad_wrapper",
]

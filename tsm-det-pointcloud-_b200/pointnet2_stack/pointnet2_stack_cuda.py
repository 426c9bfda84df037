"""Native surface of the reference's ``pointnet2_stack_cuda`` pybind module for the ops this package implements,
re-hosted on the C ABI: same names / positional arguments / returns as
``/root/reference/pcdet/ops/pointnet2/pointnet2_stack/src/pointnet2_api.cpp:13-20``."""
from __future__ import annotations

from .._lib import call, ptr, stream_ptr
from ..pointnet2_batch_cuda import farthest_point_sampling_wrapper  # noqa: F401  (:16, the dense-batch sampler)


def _chk(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise ValueError("must be a CUDA tensor")
        if not t.is_contiguous():
            raise ValueError("must be a contiguous tensor")


def voxel_query_wrapper(M, R1, R2, R3, nsample, radius, z_range, y_range, x_range, new_xyz, xyz, new_coords,
                        point_indices, idx, cnt_unique):
    """ref voxel_query.cpp:27-46"""
    _chk(new_coords, point_indices, new_xyz, xyz)
    call("tsmdet_voxel_query", M, R1, R2, R3, nsample, float(radius), int(z_range), int(y_range), int(x_range), ptr(new_xyz),
         ptr(xyz), ptr(new_coords), ptr(point_indices), ptr(idx), ptr(cnt_unique), stream_ptr(xyz.device))
    return 1


def voxel_query_dilated_wrapper(M, R1, R2, R3, nsample, former_radius, radius, z_range, y_range, x_range, z_stride,
                                y_stride, x_stride, new_xyz, xyz, new_coords, point_indices, idx, cnt_unique, idx_cnt):
    """ref voxel_query.cpp:48-75"""
    _chk(new_coords, point_indices, new_xyz, xyz)
    call("tsmdet_voxel_query_dilated", M, R1, R2, R3, nsample, float(former_radius), float(radius), int(z_range),
         int(y_range), int(x_range), int(z_stride), int(y_stride), int(x_stride), ptr(new_xyz), ptr(xyz), ptr(new_coords),
         ptr(point_indices), ptr(idx), ptr(cnt_unique), ptr(idx_cnt), stream_ptr(xyz.device))
    return 1


def group_points_wrapper(B, M, C, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out):
    """ref group_points.cpp (stack): features (N,C), idx (M,nsample) -> out (M,C,nsample)"""
    _chk(features, features_batch_cnt, idx, idx_batch_cnt, out)
    call("tsmdet_stack_group_points", B, M, C, nsample, ptr(features), ptr(features_batch_cnt), ptr(idx),
         ptr(idx_batch_cnt), ptr(out), stream_ptr(features.device))
    return 1


def group_points_grad_wrapper(B, M, C, N, nsample, grad_out, idx, idx_batch_cnt, features_batch_cnt, grad_features):
    _chk(grad_out, idx, idx_batch_cnt, features_batch_cnt, grad_features)
    call("tsmdet_stack_group_points_grad", B, M, C, N, nsample, ptr(grad_out), ptr(idx), ptr(idx_batch_cnt),
         ptr(features_batch_cnt), ptr(grad_features), stream_ptr(grad_out.device))
    return 1


def stack_farthest_point_sampling_wrapper(points, temp, xyz_batch_cnt, idx, num_sampled_points):
    """ref sampling.cpp (stack): points (N,3), temp (N) filled with 1e10, counts on the device -> idx (sum M)"""
    _chk(points, temp, xyz_batch_cnt, idx, num_sampled_points)
    call("tsmdet_stack_farthest_point_sampling", points.shape[0], xyz_batch_cnt.shape[0], ptr(points), ptr(temp),
         ptr(xyz_batch_cnt), ptr(idx), ptr(num_sampled_points), stream_ptr(points.device))
    return 1

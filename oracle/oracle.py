"""numpy front-end of the CPU oracle (``liboracle.so``).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module; the
product package (``tsmdet_b200``) never does and has no CPU fallback.

Each wrapper mirrors the argument order of the reference's pybind function it
restates (``pointnet2_api.cpp:11-32``, ``iou3d_nms_api.cpp:12-16``) but allocates the
outputs the way the reference's Python callers do (``pointnet2_utils.py:100-101,
429-430``): ``temp`` filled with 1e10, ``idx``/``idx_cnt`` zeroed.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force: bool = False) -> str:
    """Compile ``liboracle.so`` with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("pointnet2_oracle.c", "iou3d_oracle.c", "stack_oracle.c", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_opt_n_threads.restype = ctypes.c_int
        _lib.orc_nms.restype = ctypes.c_int
        _lib.orc_nms_sweep.restype = ctypes.c_int
        _lib.orc_nms_from_iou.restype = ctypes.c_int
    return _lib


def set_threads(n: int) -> None:
    """Bound the OpenMP team used by the batch/centre loops (1 = scalar port)."""
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(n))
    except OSError:  # pragma: no cover
        pass


def _f(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def opt_n_threads(n: int) -> int:
    return int(lib().orc_opt_n_threads(int(n)))


# --------------------------------------------------------------------------- FPS
def fps(xyz, npoint: int, return_temp: bool = False):
    """sampling_gpu.cu:100-216 via pointnet2_utils.py:85-111."""
    xyz = _f(xyz)
    b, n, _ = xyz.shape
    temp = np.full((b, n), 1e10, dtype=np.float32)
    idx = np.zeros((b, npoint), dtype=np.int32)
    lib().orc_fps(b, n, npoint, _p(xyz, _f32p), _p(temp, _f32p), _p(idx, _i32p))
    return (idx, temp) if return_temp else idx


def fps_weights(xyz, weights, npoint: int):
    """sampling_gpu.cu:901-1022 via pointnet2_utils.py:61-81."""
    xyz, weights = _f(xyz), _f(weights)
    b, n, _ = xyz.shape
    temp = np.full((b, n), 1e10, dtype=np.float32)
    idx = np.zeros((b, npoint), dtype=np.int32)
    lib().orc_fps_weights(b, n, npoint, _p(xyz, _f32p), _p(weights, _f32p), _p(temp, _f32p), _p(idx, _i32p))
    return idx


def fps_matrix(matrix, npoint: int):
    """sampling_gpu.cu:750-855 (== :262-377) via pointnet2_utils.py:41-58, 114-140."""
    matrix = _f(matrix)
    b, n, _ = matrix.shape
    temp = np.full((b, n), 1e10, dtype=np.float32)
    idx = np.zeros((b, npoint), dtype=np.int32)
    lib().orc_fps_matrix(b, n, npoint, _p(matrix, _f32p), _p(temp, _f32p), _p(idx, _i32p))
    return idx


def fps_weighted_matrix(matrix, weights, npoint: int):
    """sampling_gpu.cu:424-541 via pointnet2_utils.py:143-169."""
    matrix, weights = _f(matrix), _f(weights)
    b, n, _ = matrix.shape
    temp = np.full((b, n), 1e10, dtype=np.float32)
    idx = np.zeros((b, npoint), dtype=np.int32)
    lib().orc_fps_weighted_matrix(
        b, n, npoint, _p(matrix, _f32p), _p(weights, _f32p), _p(temp, _f32p), _p(idx, _i32p)
    )
    return idx


# ------------------------------------------------------------- gather / group
def gather_points(points, idx):
    points, idx = _f(points), _i(idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = np.empty((b, c, m), dtype=np.float32)
    lib().orc_gather_points(b, c, n, m, _p(points, _f32p), _p(idx, _i32p), _p(out, _f32p))
    return out


def gather_points_grad(grad_out, idx, n: int):
    grad_out, idx = _f(grad_out), _i(idx)
    b, c, m = grad_out.shape
    g = np.zeros((b, c, n), dtype=np.float32)
    lib().orc_gather_points_grad(b, c, n, m, _p(grad_out, _f32p), _p(idx, _i32p), _p(g, _f32p))
    return g


def group_points(points, idx):
    points, idx = _f(points), _i(idx)
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = np.empty((b, c, npoints, nsample), dtype=np.float32)
    lib().orc_group_points(b, c, n, npoints, nsample, _p(points, _f32p), _p(idx, _i32p), _p(out, _f32p))
    return out


def group_points_grad(grad_out, idx, n: int):
    grad_out, idx = _f(grad_out), _i(idx)
    b, c, npoints, nsample = grad_out.shape
    g = np.zeros((b, c, n), dtype=np.float32)
    lib().orc_group_points_grad(b, c, n, npoints, nsample, _p(grad_out, _f32p), _p(idx, _i32p), _p(g, _f32p))
    return g


# ------------------------------------------------------------------ ball query
def ball_query(radius: float, nsample: int, xyz, new_xyz):
    """ball_query_gpu.cu:75-112 via pointnet2_utils.py:413-433 -> (idx_cnt, idx)."""
    xyz, new_xyz = _f(xyz), _f(new_xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = np.zeros((b, m, nsample), dtype=np.int32)
    cnt = np.zeros((b, m), dtype=np.int32)
    lib().orc_ball_query(
        b, n, m, ctypes.c_float(radius), nsample, _p(new_xyz, _f32p), _p(xyz, _f32p), _p(cnt, _i32p), _p(idx, _i32p)
    )
    return cnt, idx


def ball_query_dilated(radius_in: float, radius_out: float, nsample: int, xyz, new_xyz):
    """ball_query_gpu.cu:138-176 via pointnet2_utils.py:436-457."""
    xyz, new_xyz = _f(xyz), _f(new_xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = np.zeros((b, m, nsample), dtype=np.int32)
    cnt = np.zeros((b, m), dtype=np.int32)
    lib().orc_ball_query_dilated(
        b, n, m, ctypes.c_float(radius_in), ctypes.c_float(radius_out), nsample,
        _p(new_xyz, _f32p), _p(xyz, _f32p), _p(cnt, _i32p), _p(idx, _i32p),
    )
    return cnt, idx


def query_and_group(xyz, new_xyz, features, radius: float, nsample: int, radius_in: float | None = None,
                    use_xyz: bool = True):
    """QueryAndGroup(.Dilated).forward, pointnet2_utils.py:496-568 -> (idx_cnt, new_features, grouped_xyz)."""
    xyz, new_xyz = _f(xyz), _f(new_xyz)
    if radius_in is None:
        cnt, idx = ball_query(radius, nsample, xyz, new_xyz)
    else:
        cnt, idx = ball_query_dilated(radius_in, radius, nsample, xyz, new_xyz)
    xyz_t = np.ascontiguousarray(xyz.transpose(0, 2, 1))
    grouped_xyz = group_points(xyz_t, idx) - new_xyz.transpose(0, 2, 1)[..., None]
    if features is not None:
        gf = group_points(_f(features), idx)
        new_features = np.concatenate([grouped_xyz, gf], axis=1) if use_xyz else gf
    else:
        new_features = grouped_xyz
    return cnt, new_features, grouped_xyz, idx


# --------------------------------------------------------------- interpolation
def three_nn(unknown, known):
    """interpolate_gpu.cu:16-59 -> (dist2, idx); the Python caller takes sqrt (pointnet2_utils.py:282)."""
    unknown, known = _f(unknown), _f(known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.empty((b, n, 3), dtype=np.float32)
    idx = np.empty((b, n, 3), dtype=np.int32)
    lib().orc_three_nn(b, n, m, _p(unknown, _f32p), _p(known, _f32p), _p(d2, _f32p), _p(idx, _i32p))
    return d2, idx


def three_interpolate(points, idx, weight):
    points, idx, weight = _f(points), _i(idx), _f(weight)
    b, c, m = points.shape
    n = idx.shape[1]
    out = np.empty((b, c, n), dtype=np.float32)
    lib().orc_three_interpolate(b, c, m, n, _p(points, _f32p), _p(idx, _i32p), _p(weight, _f32p), _p(out, _f32p))
    return out


def three_interpolate_grad(grad_out, idx, weight, m: int):
    grad_out, idx, weight = _f(grad_out), _i(idx), _f(weight)
    b, c, n = grad_out.shape
    g = np.zeros((b, c, m), dtype=np.float32)
    lib().orc_three_interpolate_grad(b, c, n, m, _p(grad_out, _f32p), _p(idx, _i32p), _p(weight, _f32p), _p(g, _f32p))
    return g


# ------------------------------------------------------------------- IoU / NMS
def boxes_iou_bev(boxes_a, boxes_b):
    """iou3d_cpu.cpp:232-252."""
    a, b = _f(boxes_a), _f(boxes_b)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    lib().orc_boxes_iou_bev(a.shape[0], _p(a, _f32p), b.shape[0], _p(b, _f32p), _p(out, _f32p))
    return out


def boxes_overlap_bev(boxes_a, boxes_b):
    a, b = _f(boxes_a), _f(boxes_b)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    lib().orc_boxes_overlap_bev(a.shape[0], _p(a, _f32p), b.shape[0], _p(b, _f32p), _p(out, _f32p))
    return out


def nms_sorted(boxes_sorted, thresh: float, normal: bool = False):
    """Native nms_gpu / nms_normal_gpu (iou3d_nms.cpp:90-186): boxes already in score order."""
    bx = _f(boxes_sorted)
    n = bx.shape[0]
    keep = np.zeros((max(n, 1),), dtype=np.int64)
    k = lib().orc_nms(n, _p(bx, _f32p), ctypes.c_float(thresh), int(normal), _p(keep, _i64p))
    return keep[:k].copy()


def nms_from_iou(iou, thresh: float):
    """The reference's greedy sweep replayed over a dense (N,N) IoU matrix."""
    iou = _f(iou)
    n = iou.shape[0]
    keep = np.zeros((max(n, 1),), dtype=np.int64)
    k = lib().orc_nms_from_iou(n, _p(iou, _f32p), ctypes.c_float(thresh), _p(keep, _i64p))
    return keep[:k].copy()


def nms(boxes, scores, thresh: float, normal: bool = False, order=None):
    """Python-level nms_gpu (iou3d_nms_utils.py:84-99): returns order[keep].

    ``order`` may be supplied so that oracle and candidate share one sort (torch's
    sort is not stable)."""
    boxes = _f(boxes)
    if order is None:
        order = np.argsort(-np.asarray(scores, dtype=np.float32), kind="stable")
    order = np.asarray(order, dtype=np.int64)
    keep = nms_sorted(boxes[order], thresh, normal)
    return order[keep]


# --------------------------------------------------------------- pointnet2_stack
def voxel_query(max_range, radius: float, nsample: int, xyz, new_xyz, new_coords, point_indices, stride=(1, 1, 1),
                former_radius: float | None = None):
    """voxel_query_gpu.cu:10-98 (former_radius None) / :125-215 (dilated) -> (idx (M,nsample), cnt_unique (M),
    idx_cnt (M) | None), raw kernel outputs (idx[:,0] == -1 marks an empty ball)."""
    xyz, new_xyz = _f(xyz), _f(new_xyz)
    new_coords, point_indices = _i(new_coords), _i(point_indices)
    m = new_coords.shape[0]
    _, r1, r2, r3 = point_indices.shape
    idx = np.zeros((m, nsample), dtype=np.int32)
    cnt_unique = np.zeros((m,), dtype=np.int32)
    idx_cnt = np.zeros((m,), dtype=np.int32) if former_radius is not None else None
    lib().orc_voxel_query(
        m, r1, r2, r3, nsample, ctypes.c_float(-1.0 if former_radius is None else former_radius), ctypes.c_float(radius),
        int(max_range[0]), int(max_range[1]), int(max_range[2]), int(stride[0]), int(stride[1]), int(stride[2]),
        _p(new_xyz, _f32p), _p(xyz, _f32p), _p(new_coords, _i32p), _p(point_indices, _i32p), _p(idx, _i32p),
        _p(cnt_unique, _i32p), _p(idx_cnt, _i32p) if idx_cnt is not None else None)
    return idx, cnt_unique, idx_cnt


def stack_group_points(features, features_batch_cnt, idx, idx_batch_cnt):
    features, idx = _f(features), _i(idx)
    fb, ib = _i(features_batch_cnt), _i(idx_batch_cnt)
    m, nsample = idx.shape
    c = features.shape[1]
    out = np.empty((m, c, nsample), dtype=np.float32)
    lib().orc_stack_group_points(len(ib), m, c, nsample, _p(features, _f32p), _p(fb, _i32p), _p(idx, _i32p),
                                 _p(ib, _i32p), _p(out, _f32p))
    return out


def stack_group_points_grad(grad_out, idx, idx_batch_cnt, features_batch_cnt, n: int):
    grad_out, idx = _f(grad_out), _i(idx)
    fb, ib = _i(features_batch_cnt), _i(idx_batch_cnt)
    m, c, nsample = grad_out.shape
    g = np.zeros((n, c), dtype=np.float32)
    lib().orc_stack_group_points_grad(len(ib), m, c, n, nsample, _p(grad_out, _f32p), _p(idx, _i32p), _p(ib, _i32p),
                                      _p(fb, _i32p), _p(g, _f32p))
    return g


def stack_fps(xyz, xyz_batch_cnt, npoints):
    """sampling_gpu.cu:188-316: xyz (sum N,3), xyz_batch_cnt (B), npoints (B) -> (sum M) global row numbers."""
    xyz = _f(xyz)
    cnt, npts = _i(xyz_batch_cnt), _i(npoints)
    temp = np.full((xyz.shape[0],), 1e10, dtype=np.float32)
    out = np.zeros((max(int(npts.sum()), 1),), dtype=np.int32)
    lib().orc_stack_fps(len(cnt), _p(xyz, _f32p), _p(temp, _f32p), _p(cnt, _i32p), _p(out, _i32p), _p(npts, _i32p))
    return out[: int(npts.sum())]

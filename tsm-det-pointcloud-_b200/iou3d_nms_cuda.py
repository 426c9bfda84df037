"""Native surface of the reference's ``iou3d_nms_cuda`` pybind module, re-hosted on the C ABI.

Same names / arguments / returns as ``/root/reference/pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-17``.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import call, ptr, stream_ptr


def _chk(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise ValueError("must be a CUDA tensor")
        if not t.is_contiguous():
            raise ValueError("must be a contiguous tensor")


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    """ref: iou3d_nms.cpp:49-68"""
    _chk(boxes_a, boxes_b, ans_overlap)
    call("tsmdet_boxes_overlap_bev", boxes_a.size(0), ptr(boxes_a), boxes_b.size(0), ptr(boxes_b), ptr(ans_overlap),
         stream_ptr(boxes_a.device))
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    """ref: iou3d_nms.cpp:70-88"""
    _chk(boxes_a, boxes_b, ans_iou)
    call("tsmdet_boxes_iou_bev", boxes_a.size(0), ptr(boxes_a), boxes_b.size(0), ptr(boxes_b), ptr(ans_iou),
         stream_ptr(boxes_a.device))
    return 1


def _nms(fn, boxes, keep, thresh):
    _chk(boxes)
    if keep.is_cuda or keep.dtype != torch.int64 or not keep.is_contiguous():
        raise ValueError("keep must be a contiguous CPU int64 tensor (as in the reference)")
    num = ctypes.c_int(0)
    call(fn, boxes.size(0), ptr(boxes), float(thresh), ptr(keep), ctypes.byref(num), stream_ptr(boxes.device))
    return int(num.value)


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """ref: iou3d_nms.cpp:90-136 -- boxes (N,7) CUDA in score order, keep (N) CPU int64, returns num kept."""
    return _nms("tsmdet_nms_gpu", boxes, keep, nms_overlap_thresh)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    """ref: iou3d_nms.cpp:139-186"""
    return _nms("tsmdet_nms_normal_gpu", boxes, keep, nms_overlap_thresh)


def boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou):
    """ref: iou3d_cpu.cpp:232-252 -- host tensors in and out."""
    for t in (boxes_a, boxes_b, ans_iou):
        if t.is_cuda:
            raise ValueError("boxes_iou_bev_cpu takes CPU tensors")
        if not t.is_contiguous():
            raise ValueError("must be a contiguous tensor")
    call("tsmdet_boxes_iou_bev_cpu", boxes_a.size(0), ptr(boxes_a), boxes_b.size(0), ptr(boxes_b), ptr(ans_iou))
    return 1


__all__ = ["boxes_overlap_bev_gpu", "boxes_iou_bev_gpu", "nms_gpu", "nms_normal_gpu", "boxes_iou_bev_cpu"]

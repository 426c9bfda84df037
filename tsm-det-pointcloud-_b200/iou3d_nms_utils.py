"""Rotated-box IoU / NMS helpers -- same names, arguments and results as
``/root/reference/pcdet/ops/iou3d_nms/iou3d_nms_utils.py:12-116``.

``nms_gpu`` keeps the reference contract (returns ``(order[keep], None)``, kept indices in score
order, ``pre_maxsize`` honoured, extra kwargs ignored -- so the upper-case config keys the callers
splat in are harmless, SURVEY.md section 0 bug 3), but the suppression mask never leaves the device: the sweep
runs in a kernel and only the kept indices are produced.  ``nms_gpu_batch`` is the device-resident
batched form used by the post-processing path.
"""
from __future__ import annotations

import torch

from . import common_utils, iou3d_nms_cuda
from ._lib import call, ptr, stream_ptr


def boxes_bev_iou_cpu(boxes_a, boxes_b):
    """ref :12-28 -- numpy or CPU tensors in, same kind out."""
    boxes_a, is_numpy = common_utils.check_numpy_to_torch(boxes_a)
    boxes_b, is_numpy = common_utils.check_numpy_to_torch(boxes_b)
    assert not (boxes_a.is_cuda or boxes_b.is_cuda), 'Only support CPU tensors'
    assert boxes_a.shape[1] == 7 and boxes_b.shape[1] == 7
    ans_iou = boxes_a.new_zeros(torch.Size((boxes_a.shape[0], boxes_b.shape[0])))
    iou3d_nms_cuda.boxes_iou_bev_cpu(boxes_a.contiguous(), boxes_b.contiguous(), ans_iou)
    return ans_iou.numpy() if is_numpy else ans_iou


def boxes_iou_bev(boxes_a, boxes_b):
    """ref :31-45 -- (N,7),(M,7) CUDA -> (N,M) rotated BEV IoU."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    ans_iou = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_nms_cuda.boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans_iou)
    return ans_iou


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """ref :48-81 -- BEV overlap from the kernel, height overlap and volumes in torch."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7

    boxes_a_height_max = (boxes_a[:, 2] + boxes_a[:, 5] / 2).view(-1, 1)
    boxes_a_height_min = (boxes_a[:, 2] - boxes_a[:, 5] / 2).view(-1, 1)
    boxes_b_height_max = (boxes_b[:, 2] + boxes_b[:, 5] / 2).view(1, -1)
    boxes_b_height_min = (boxes_b[:, 2] - boxes_b[:, 5] / 2).view(1, -1)

    overlaps_bev = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_nms_cuda.boxes_overlap_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), overlaps_bev)

    max_of_min = torch.max(boxes_a_height_min, boxes_b_height_min)
    min_of_max = torch.min(boxes_a_height_max, boxes_b_height_max)
    overlaps_h = torch.clamp(min_of_max - max_of_min, min=0)

    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-6)


def _nms_device(fn_name: str, boxes_sorted: torch.Tensor, thresh: float):
    """Run the device-side mask + sweep for one frame; returns (keep (n,) int64 CUDA, num (1,) int32 CUDA)."""
    n = boxes_sorted.size(0)
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=boxes_sorted.device)
    num = torch.zeros((1,), dtype=torch.int32, device=boxes_sorted.device)
    call(fn_name, 1, n, ptr(boxes_sorted), boxes_sorted.size(1), None, float(thresh), ptr(keep), ptr(num),
         stream_ptr(boxes_sorted.device))
    return keep, num


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """ref :84-99 -- boxes (N,7), scores (N) -> (LongTensor of kept original indices, None)."""
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order].contiguous()
    keep, num = _nms_device("tsmdet_nms_batch", boxes, thresh)
    num_out = int(num.item())  # the one sync the reference also has (it copies the whole mask instead)
    return order[keep[:num_out]].contiguous(), None


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    """ref :102-116 -- axis-aligned variant."""
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    boxes = boxes[order].contiguous()
    keep, num = _nms_device("tsmdet_nms_normal_batch", boxes, thresh)
    num_out = int(num.item())
    return order[keep[:num_out]].contiguous(), None


@torch.no_grad()
def nms_gpu_batch(boxes, scores, thresh, counts=None, normal: bool = False, presorted: bool = False):
    """Device-resident NMS over a batch of frames -- no host synchronisation.

    boxes (F,N,>=7), scores (F,N) [padded entries should carry -inf], counts (F,) int32 valid boxes per
    frame or None; presorted=True promises each frame is already in descending score order.  Returns
    (selected (F,N) int64: original indices of the kept boxes in score order, padded with -1;
    num_keep (F,) int32)."""
    f, n = scores.shape
    if presorted:
        order = None
        sorted_boxes = boxes.contiguous()
    else:
        order = scores.sort(1, descending=True)[1]
        sorted_boxes = torch.gather(boxes, 1, order.unsqueeze(-1).expand(-1, -1, boxes.size(2))).contiguous()
    keep = torch.empty((f, max(n, 1)), dtype=torch.int64, device=boxes.device)
    num = torch.zeros((f,), dtype=torch.int32, device=boxes.device)
    call("tsmdet_nms_normal_batch" if normal else "tsmdet_nms_batch", f, n, ptr(sorted_boxes), sorted_boxes.size(2),
         ptr(counts), float(thresh), ptr(keep), ptr(num), stream_ptr(boxes.device))
    ar = torch.arange(n, device=boxes.device).unsqueeze(0)
    valid = ar < num.unsqueeze(1)
    sel = torch.where(valid, keep[:, :n], torch.zeros_like(keep[:, :n]))
    if order is not None:
        sel = torch.gather(order, 1, sel)
    return torch.where(valid, sel, torch.full_like(sel, -1)), num

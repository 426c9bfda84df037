"""Post-processing NMS drivers -- same functions and results as
``/root/reference/pcdet/models/model_utils/model_nms_utils.py:6-127`` on top of this package's
``iou3d_nms_utils`` (``NMS_TYPE`` is still looked up by name with ``getattr``)."""
from __future__ import annotations

import torch

from . import iou3d_nms_utils


def _nms_fn(nms_config):
    return getattr(iou3d_nms_utils, nms_config.NMS_TYPE)


def class_agnostic_nms(box_scores, box_preds, nms_config, score_thresh=None, depth_score=None):
    """ref :6-28"""
    src_box_scores = box_scores
    if score_thresh is not None:
        scores_mask = (box_scores >= score_thresh)
        box_scores = box_scores[scores_mask]
        box_preds = box_preds[scores_mask]
    if depth_score is not None:
        depth_score = depth_score[scores_mask]
        box_scores = box_scores * depth_score

    selected = []
    if box_scores.shape[0] > 0:
        box_scores_nms, indices = torch.topk(box_scores, k=min(nms_config.NMS_PRE_MAXSIZE, box_scores.shape[0]))
        boxes_for_nms = box_preds[indices]
        keep_idx, _ = _nms_fn(nms_config)(boxes_for_nms[:, 0:7], box_scores_nms, nms_config.NMS_THRESH, **nms_config)
        selected = indices[keep_idx[:nms_config.NMS_POST_MAXSIZE]]

    if score_thresh is not None:
        original_idxs = scores_mask.nonzero().view(-1)
        selected = original_idxs[selected]
    return selected, src_box_scores[selected]


def class_agnostic_nms_v2(box_scores, box_labels, box_preds, nms_config, score_thresh=None):
    """ref :31-49 -- per-class (labels 0..2) pre/post sizes and thresholds given as lists."""
    src_box_scores = box_scores
    selected_idx = []
    for i in range(0, 3):
        mask = (box_labels == i)
        box_scores_mask = box_scores[mask]
        box_preds_mask = box_preds[mask]
        if box_scores.shape[0] > 0:
            box_scores_nms, indices = torch.topk(
                box_scores_mask, k=min((nms_config.NMS_PRE_MAXSIZE)[i], box_scores_mask.shape[0]))
            boxes_for_nms = box_preds_mask[indices]
            keep_idx, _ = _nms_fn(nms_config)(
                boxes_for_nms[:, 0:7], box_scores_nms, (nms_config.NMS_THRESH)[i], **nms_config)
            selected = indices[keep_idx[:(nms_config.NMS_POST_MAXSIZE)[i]]]
        original_idxs = mask.nonzero().view(-1)
        selected = original_idxs[selected]
        selected_idx.append(selected)
    selected_idx = torch.cat(selected_idx, dim=-1)
    return selected_idx, src_box_scores[selected_idx]


def multi_thresh(box_scores, box_labels, box_preds, nms_config, score_thresh=None):
    """ref :52-87 -- the live one: per-class score threshold -> top-k -> NMS, then one cross-class NMS."""
    src_box_scores = box_scores
    selected = []
    selected_end = []
    if score_thresh is not None:
        for i, cur_thresh in enumerate(score_thresh):
            mask = ((i + 1) == box_labels)
            cur_box_scores = box_scores[mask]
            cur_box_preds = box_preds[mask]
            score_mask = (cur_box_scores >= cur_thresh)
            cur_box_scores = cur_box_scores[score_mask]
            cur_box_preds = cur_box_preds[score_mask]

            if cur_box_scores.shape[0] > 0:
                cur_box_scores_nms, indices = torch.topk(
                    cur_box_scores, k=min(nms_config.NMS_PRE_MAXSIZE, cur_box_scores.shape[0]))
                cur_box_for_nms = cur_box_preds[indices]
                keep_idx, _ = _nms_fn(nms_config)(
                    cur_box_for_nms, cur_box_scores_nms, nms_config.NMS_THRESH, **nms_config)
                cur_selected = indices[keep_idx[:nms_config.NMS_POST_MAXSIZE]]

                score_idxs = score_mask.nonzero().view(-1)
                original_idxs = mask.nonzero().view(-1)
                cur_selected = score_idxs[cur_selected]
                cur_selected = original_idxs[cur_selected]
                selected.append(cur_selected)
    if len(selected):
        selected = torch.cat(selected, dim=0)
        box_for_nms_end = box_preds[selected]
        box_scores_nms_end = box_scores[selected]
        keep_idx_end, _ = _nms_fn(nms_config)(
            box_for_nms_end, box_scores_nms_end, nms_config.NMS_THRESH, **nms_config)
        selected_end = selected[keep_idx_end]
    return selected_end, src_box_scores[selected_end]


@torch.no_grad()
def multi_thresh_batch(box_scores, box_labels, box_preds, nms_config, score_thresh):
    """`multi_thresh` for a whole batch of frames on the device, with no host synchronisation: what the
    reference's per-frame / per-class Python loop (``detector3d_template.py:228-303`` calling ref :52-87, up to
    num_class + 1 ``nms_gpu`` calls per frame, each ending in a blocking copy and ``nonzero()`` syncs) computes,
    as ``num_class + 1`` batched NMS launches in total.

    box_scores (F,P), box_labels (F,P) int (classes are 1-based, as in the reference), box_preds (F,P,>=7),
    score_thresh: one threshold per class.  Returns ``(selected (F,K) int64, num (F,) int32, scores (F,K))``:
    per frame the indices ``multi_thresh`` returns (score order), padded with -1 / -inf; K = num_class *
    NMS_POST_MAXSIZE.  Rotated NMS only (``NMS_TYPE == 'nms_gpu'``); scores are assumed distinct where order
    matters (torch.sort / topk leave the order of ties unspecified in the reference too)."""
    if nms_config.NMS_TYPE not in ("nms_gpu", "nms_normal_gpu"):
        raise ValueError(f"multi_thresh_batch: unsupported NMS_TYPE {nms_config.NMS_TYPE}")
    normal = nms_config.NMS_TYPE == "nms_normal_gpu"
    f, p = box_scores.shape
    dev = box_scores.device
    pre = min(int(nms_config.NMS_PRE_MAXSIZE), p)
    post = int(nms_config.NMS_POST_MAXSIZE)
    thresh = float(nms_config.NMS_THRESH)
    boxes7 = box_preds[:, :, :7]
    neg = torch.full_like(box_scores, float("-inf"))
    parts_idx, parts_score = [], []
    for i, cur_thresh in enumerate(score_thresh):
        live = (box_labels == (i + 1)) & (box_scores >= cur_thresh)
        vals, order = torch.topk(torch.where(live, box_scores, neg), k=pre, dim=1)  # descending; dead entries last
        counts = (vals > float("-inf")).sum(dim=1).to(torch.int32)
        cand = torch.gather(boxes7, 1, order.unsqueeze(-1).expand(-1, -1, 7))
        sel, num = iou3d_nms_utils.nms_gpu_batch(cand, vals, thresh, counts=counts, normal=normal, presorted=True)
        k = min(post, pre)
        sel = sel[:, :k]
        ok = sel >= 0
        orig = torch.gather(order, 1, torch.where(ok, sel, torch.zeros_like(sel)))
        parts_idx.append(torch.where(ok, orig, torch.full_like(orig, -1)))
        parts_score.append(torch.where(ok, torch.gather(box_scores, 1, torch.where(ok, orig, torch.zeros_like(orig))),
                                       torch.full((f, k), float("-inf"), device=dev, dtype=box_scores.dtype)))
    if not parts_idx:
        return (torch.full((f, 0), -1, dtype=torch.int64, device=dev), torch.zeros((f,), dtype=torch.int32, device=dev),
                torch.full((f, 0), float("-inf"), device=dev, dtype=box_scores.dtype))
    cat_idx = torch.cat(parts_idx, dim=1)        # (F, K)
    cat_score = torch.cat(parts_score, dim=1)
    ok = cat_idx >= 0
    cat_boxes = torch.gather(boxes7, 1, torch.where(ok, cat_idx, torch.zeros_like(cat_idx)).unsqueeze(-1).expand(-1, -1, 7))
    counts = ok.sum(dim=1).to(torch.int32)
    # the cross-class pass: nms_gpu sorts by score itself (padding carries -inf and sorts last)
    sel, num = iou3d_nms_utils.nms_gpu_batch(cat_boxes, cat_score, thresh, counts=counts, normal=normal)
    ok2 = sel >= 0
    safe = torch.where(ok2, sel, torch.zeros_like(sel))
    out_idx = torch.where(ok2, torch.gather(cat_idx, 1, safe), torch.full_like(sel, -1))
    out_score = torch.where(ok2, torch.gather(cat_score, 1, safe), torch.full_like(cat_score, float("-inf")))
    return out_idx, num, out_score


def multi_classes_nms(cls_scores, box_preds, nms_config, score_thresh=None):
    """ref :89-127 -- cls_scores (N,num_class), box_preds (N,7+C) -> (scores, labels, boxes)."""
    pred_scores, pred_labels, pred_boxes = [], [], []
    for k in range(cls_scores.shape[1]):
        if score_thresh is not None:
            scores_mask = (cls_scores[:, k] >= score_thresh)
            box_scores = cls_scores[scores_mask, k]
            cur_box_preds = box_preds[scores_mask]
        else:
            box_scores = cls_scores[:, k]
            cur_box_preds = box_preds

        selected = []
        if box_scores.shape[0] > 0:
            box_scores_nms, indices = torch.topk(box_scores, k=min(nms_config.NMS_PRE_MAXSIZE, box_scores.shape[0]))
            boxes_for_nms = cur_box_preds[indices]
            keep_idx, _ = _nms_fn(nms_config)(boxes_for_nms[:, 0:7], box_scores_nms, nms_config.NMS_THRESH, **nms_config)
            selected = indices[keep_idx[:nms_config.NMS_POST_MAXSIZE]]

        pred_scores.append(box_scores[selected])
        pred_labels.append(box_scores.new_ones(len(selected)).long() * k)
        pred_boxes.append(cur_box_preds[selected])

    pred_scores = torch.cat(pred_scores, dim=0)
    pred_labels = torch.cat(pred_labels, dim=0)
    pred_boxes = torch.cat(pred_boxes, dim=0)
    return pred_scores, pred_labels, pred_boxes


class NmsConfig(dict):
    """Minimal attribute-dict stand-in for the reference's EasyDict NMS_CONFIG (keys are upper-case and
    are splatted into the nms function as kwargs, exactly as the reference does)."""

    __getattr__ = dict.__getitem__

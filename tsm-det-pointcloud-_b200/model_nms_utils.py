"""Post-processing NMS drivers -- the functions of
``/root/reference/pcdet/models/model_utils/model_nms_utils.py:6-127`` (same names, arguments and results) on top
of this package's ``iou3d_nms_utils``.

The reference's own file runs unchanged over this package's ``iou3d_nms_utils`` (INTEGRATION.md section 1;
``tests/test_dropin_reference_py_gpu.py`` does exactly that), so nothing here is a transcription: every driver is
the same three steps -- restrict to a subset, keep the ``NMS_PRE_MAXSIZE`` best, run ``NMS_TYPE`` and keep
``NMS_POST_MAXSIZE`` -- expressed ONCE (``_subset_topk_nms``) with device-side masks instead of the reference's
boolean-index compactions and ``nonzero()`` calls, and ``multi_thresh_batch`` runs the live driver for a whole batch
of frames without any host synchronisation.  Scores are assumed distinct where order matters (``torch.topk`` /
``sort`` leave the order of ties unspecified in the reference too).
"""
from __future__ import annotations

import torch

from . import iou3d_nms_utils

_NEG_INF = float("-inf")


def _subset_topk_nms(rank_scores, boxes7, subset, pre_max, post_max, thresh, nms_config):
    """Indices (into the full arrays, best first) that survive: restrict to ``subset`` (bool mask or None), take the
    ``pre_max`` highest ``rank_scores``, run ``nms_config.NMS_TYPE`` at ``thresh``, keep ``post_max``."""
    p = rank_scores.shape[0]
    if p == 0:
        return rank_scores.new_zeros((0,), dtype=torch.int64)
    masked = rank_scores if subset is None else torch.where(subset, rank_scores, rank_scores.new_full((), _NEG_INF))
    vals, order = torch.topk(masked, k=min(int(pre_max), p))          # descending; excluded entries sort last
    if nms_config.NMS_TYPE in ("nms_gpu", "nms_normal_gpu"):
        # one batched launch with a device-side count: no boolean indexing, one sync for the variable-length result
        counts = (vals > _NEG_INF).sum().to(torch.int32).view(1)
        sel, num = iou3d_nms_utils.nms_gpu_batch(boxes7[order].unsqueeze(0), vals.unsqueeze(0), float(thresh), counts=counts,
                                                 normal=nms_config.NMS_TYPE == "nms_normal_gpu", presorted=True)
        kept = sel[0, :min(int(num.item()), int(post_max))]
    else:  # any other NMS_TYPE the caller's iou3d_nms_utils offers, looked up by name as the reference does
        live = int((vals > _NEG_INF).sum().item())
        kept, _ = getattr(iou3d_nms_utils, nms_config.NMS_TYPE)(boxes7[order[:live]], vals[:live], thresh, **nms_config)
        kept = kept[:int(post_max)]
    return order[kept]


def class_agnostic_nms(box_scores, box_preds, nms_config, score_thresh=None, depth_score=None):
    """ref :6-28 -- optional score threshold, optional depth re-weighting of the ranking score."""
    subset = None if score_thresh is None else box_scores >= score_thresh
    rank = box_scores
    if depth_score is not None:
        if score_thresh is None:  # the reference indexes depth_score with a mask that only exists with a threshold
            raise UnboundLocalError("class_agnostic_nms: depth_score needs score_thresh (reference :13-14)")
        rank = box_scores * depth_score
    selected = _subset_topk_nms(rank, box_preds[:, 0:7], subset, nms_config.NMS_PRE_MAXSIZE,
                                nms_config.NMS_POST_MAXSIZE, nms_config.NMS_THRESH, nms_config)
    return selected, box_scores[selected]


def class_agnostic_nms_v2(box_scores, box_labels, box_preds, nms_config, score_thresh=None):
    """ref :31-49 -- labels 0..2, with per-class lists for the pre / post sizes and the threshold."""
    parts = [_subset_topk_nms(box_scores, box_preds[:, 0:7], box_labels == c, nms_config.NMS_PRE_MAXSIZE[c],
                              nms_config.NMS_POST_MAXSIZE[c], nms_config.NMS_THRESH[c], nms_config) for c in range(3)]
    selected = torch.cat(parts, dim=-1)
    return selected, box_scores[selected]


def multi_thresh(box_scores, box_labels, box_preds, nms_config, score_thresh=None):
    """ref :52-87 -- the live driver: per class (labels 1..C) score threshold -> top-k -> NMS, then one cross-class
    NMS over the union.  Returns ``([], empty)`` when nothing survives, like the reference."""
    parts = []
    for c, cur_thresh in enumerate(score_thresh if score_thresh is not None else ()):
        subset = (box_labels == (c + 1)) & (box_scores >= cur_thresh)
        part = _subset_topk_nms(box_scores, box_preds[:, 0:7], subset, nms_config.NMS_PRE_MAXSIZE,
                                nms_config.NMS_POST_MAXSIZE, nms_config.NMS_THRESH, nms_config)
        if part.numel():
            parts.append(part)
    if not parts:
        return [], box_scores[[]]
    union = torch.cat(parts, dim=0)
    final = _subset_topk_nms(box_scores[union], box_preds[union][:, 0:7], None, union.numel(), union.numel(),
                             nms_config.NMS_THRESH, nms_config)
    selected = union[final]
    return selected, box_scores[selected]


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
    """ref :89-127 -- cls_scores (N,num_class), box_preds (N,7+C) -> (scores, labels, boxes), class after class."""
    out_scores, out_labels, out_boxes = [], [], []
    for k in range(cls_scores.shape[1]):
        col = cls_scores[:, k]
        subset = None if score_thresh is None else col >= score_thresh
        sel = _subset_topk_nms(col, box_preds[:, 0:7], subset, nms_config.NMS_PRE_MAXSIZE, nms_config.NMS_POST_MAXSIZE,
                               nms_config.NMS_THRESH, nms_config)
        out_scores.append(col[sel])
        out_labels.append(torch.full((sel.numel(),), k, dtype=torch.int64, device=col.device))
        out_boxes.append(box_preds[sel])
    return torch.cat(out_scores, dim=0), torch.cat(out_labels, dim=0), torch.cat(out_boxes, dim=0)


class NmsConfig(dict):
    """Minimal attribute-dict stand-in for the reference's EasyDict NMS_CONFIG (keys are upper-case and
    are splatted into the nms function as kwargs, exactly as the reference does)."""

    __getattr__ = dict.__getitem__

"""Checkers shared by the -m gpu tests, ``__graft_entry__.smoke()`` and ``bench.py``'s post-run verification
(TEST INFRASTRUCTURE: this module calls the CPU oracle; nothing on the product path imports it).

Floating-point metrics, stated as measured instead of one loose bound.  For ``got`` vs ``want``:
  max_abs_over_scale   max|got - want| / max(max|want|, 1)             -- the bound of round 1
  rel_l2               ||got - want||_2 / ||want||_2                    -- the "1e-2 rel" of north_star, as a norm
  max_rel_big          max over |want| > 0.1 * max|want| of |got - want| / |want|   -- elementwise, large outputs
  max_rel_all          max over |want| > 1e-3 * max|want| of |got - want| / |want|  -- elementwise, reported only:
                       after ReLU a near-zero output is a cancellation of O(scale) terms, so bf16 operands
                       (2^-9 relative each) give it an O(2^-9 * scale) ABSOLUTE error whatever the kernel does.
Bars (north_star: "max abs 1e-5 for fp32, 1e-2 rel for bf16 MLP"): fp32 path max_abs <= 1e-5 ABSOLUTE (measured on
B200: <= 1.2e-6 over every test shape), rel_l2 <= 1e-5; bf16 path rel_l2 <= 1e-2 (the "1e-2 rel", as a norm; measured
<= 2.8e-3), max_abs_over_scale <= 1e-2 (measured <= 4.3e-3) and max_rel_big <= 5e-2 (measured <= 3.1e-2: an output
above a tenth of the largest one is within 5 % elementwise).
"""
from __future__ import annotations

import numpy as np
import torch

BARS = {
    "fp32": {"max_abs": 1e-5, "rel_l2": 1e-5, "max_rel_big": 1e-4},
    "bf16": {"max_abs_over_scale": 1e-2, "rel_l2": 1e-2, "max_rel_big": 5e-2},
    "tf32": {"max_abs_over_scale": 2e-3, "rel_l2": 2e-3, "max_rel_big": 6e-3},
}


def err_metrics(got: torch.Tensor, want: torch.Tensor) -> dict:
    got, want = got.double(), want.double()
    diff = (got - want).abs()
    scale = float(want.abs().max())
    big = want.abs() > 0.1 * scale
    anyv = want.abs() > 1e-3 * scale
    return {
        "scale": scale,
        "max_abs": float(diff.max()),
        "max_abs_over_scale": float(diff.max()) / max(scale, 1.0),
        "rel_l2": float(diff.norm() / want.norm().clamp_min(1e-30)),
        "max_rel_big": float((diff[big] / want.abs()[big]).max()) if bool(big.any()) else 0.0,
        "max_rel_all": float((diff[anyv] / want.abs()[anyv]).max()) if bool(anyv.any()) else 0.0,
    }


def check_metrics(m: dict, precision: str, what: str = "") -> None:
    for k, bar in BARS[precision].items():
        assert m[k] <= bar, f"{what}: {k} = {m[k]:.3e} exceeds the {precision} bar {bar:g} ({m})"


def oracle_fps_chain(orc, xyz_np: np.ndarray, npoints):
    """[(idx, new_xyz)] per level: the oracle's FPS run level after level on the centres it picked."""
    out, cur = [], xyz_np
    for m in npoints:
        idx = orc.fps(cur, m)
        cur = np.take_along_axis(cur, idx.astype(np.int64)[..., None], axis=1)
        out.append((idx, cur))
    return out


def oracle_two_pass_keep(orc, iou_sorted: np.ndarray, order: np.ndarray, pre: float, post: float, k_post: int):
    """Keep-list of 'NMS at `pre` over all proposals, then at `post` over the survivors' (score order), replayed by
    the oracle's greedy sweep over the IoU matrix of the score-sorted boxes."""
    k1 = orc.nms_from_iou(iou_sorted, pre)
    k2 = orc.nms_from_iou(np.ascontiguousarray(iou_sorted[np.ix_(k1, k1)]), post)
    return order[k1[k2]][:k_post]


@torch.no_grad()
def verify_step(engine, orc, xyz_np, feats_np, boxes_np, scores_np, got: dict, precision: str, frames=None) -> dict:
    """One step's results (``got``: xyz / features / det_idx or det / det_num, device or host tensors) against
    (i) the oracle's FPS chain (final centres bit-exact => every level's indices are the oracle's),
    (ii) the eager fp32 Conv2d/BatchNorm2d/ReLU/max_pool2d stack on the same centres (stated tolerance),
    (iii) the oracle's greedy sweep over the rotated-IoU matrix (keep-lists bit-exact).  Returns the metrics."""
    from tsmdet_b200 import iou3d_nms_utils

    dev = next(engine.parameters()).device
    frames = list(range(xyz_np.shape[0])) if frames is None else list(frames)
    layers = engine.backbone.layers
    chain = oracle_fps_chain(orc, xyz_np[frames], [l.npoint_list[0] for l in layers])
    got_xyz = got["xyz"].cpu().numpy()[frames]
    assert np.array_equal(got_xyz, chain[-1][1]), "sampled centres differ from the oracle's FPS chain"
    # (ii) eager fp32 reference stack (TF32 off), same weights
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        for l in layers:
            l.fused = False
        outs = engine.backbone(torch.from_numpy(xyz_np[frames]).to(dev), torch.from_numpy(feats_np[frames]).to(dev))
    finally:
        for l in layers:
            l.fused = True
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    for (idx_o, _), (_, _, idx_e) in zip(chain, outs):
        assert np.array_equal(idx_e.cpu().numpy(), idx_o), "FPS indices differ from the oracle"
    want = outs[-1][1]
    m = err_metrics(got["features"].to(dev)[frames], want)
    check_metrics(m, precision, "SA features vs eager fp32 stack")
    # (iii) keep-lists
    k_post = got["det"].shape[1] if "det" in got else got["det_idx"].shape[1]
    kept_total = 0
    for f in frames:
        tb = torch.from_numpy(boxes_np[f]).to(dev)
        ts = torch.from_numpy(scores_np[f]).to(dev)
        order = ts.sort(0, descending=True)[1]
        sb = tb[order].contiguous()
        iou = iou3d_nms_utils.boxes_iou_bev(sb, sb).cpu().numpy()
        want_idx = oracle_two_pass_keep(orc, iou, order.cpu().numpy(), engine.nms_pre, engine.nms_post, k_post)
        n = int(got["det_num"][f])
        assert n == want_idx.size, f"frame {f}: kept {n}, oracle keeps {want_idx.size}"
        if "det_idx" in got:
            assert np.array_equal(got["det_idx"][f, :n].cpu().numpy(), want_idx), f"frame {f}: keep-list differs"
        rec = got["det"][f].cpu().numpy()
        assert np.array_equal(rec[:n, :7], boxes_np[f][want_idx]) and np.array_equal(rec[:n, 7], scores_np[f][want_idx])
        assert not rec[n:].any()
        kept_total += n
    m["frames_checked"] = len(frames)
    m["detections_checked"] = kept_total
    return m

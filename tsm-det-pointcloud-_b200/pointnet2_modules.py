"""Set-abstraction (layer-0 branch) and feature-propagation modules on the sm_100a kernels.

Mirrors, for the part of the reference that runs on ``pointnet2_batch`` ops,
``/root/reference/pcdet/ops/pointnet2/pointnet2_batch/pointnet2_modules.py``:

  * ``PointnetFPModule``                     :130-178  (three_nn -> inverse-distance weights -> three_interpolate -> MLP)
  * ``PointnetSAModuleFSMSG`` (this file)    the ``sa_layer_idx == 0`` / ``sp_tensor is None`` branch of
    ``_VoxelPointnetSAModuleFS(Distillation)Base.forward`` :1143, 1153-1221, 1259-1268, 1297-1321 with the
    MLPs built as in ``VoxelPointnetSAModuleFSMSGDistillation.__init__`` :1517-1558, 1589-1603:
    sample (d-fps / f-fps / s-fps / s-topk) -> gather centres -> per scale {ball query (plain or
    dilated) -> group -> mask empty balls -> [Conv2d 1x1, BN, ReLU]* -> max over nsample}
    -> concat scales -> aggregation MLP.

What happens after the aggregation MLP in the reference (centroid voxelisation, spconv) is a
different algorithm family and outside this package.

Two execution modes give the same numbers within the stated tolerances:
  ``fused=False``  the reference's eager sequence (materialised (B,C,npoint,nsample) tensors, torch
                   Conv2d/BatchNorm2d/ReLU/max_pool2d) -- kept as the behavioural reference;
  ``fused=True``   (eval mode only) BatchNorm folded into the 1x1 convs, one fused kernel per scale
                   (gather + MLP + max-pool, nothing materialised).  ``precision='fp32'`` is the
                   1e-5 parity mode, ``precision='bf16'`` / ``'tf32'`` run the contraction on tcgen05 tensor cores
                   (kind::f16 / kind::tf32).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils
from ._lib import call, ptr, stream_ptr


def build_shared_mlp(spec: Sequence[int], bn: bool = True, dims: int = 2) -> nn.Sequential:
    """[Conv{dims}d 1x1 (bias iff no BN), BatchNorm, ReLU] per consecutive pair of ``spec`` (ref :1549-1555)."""
    conv = nn.Conv2d if dims == 2 else nn.Conv1d
    norm = nn.BatchNorm2d if dims == 2 else nn.BatchNorm1d
    layers: List[nn.Module] = []
    for k in range(len(spec) - 1):
        layers.append(conv(spec[k], spec[k + 1], kernel_size=1, bias=not bn))
        if bn:
            layers.append(norm(spec[k + 1]))
        layers.append(nn.ReLU())
    return nn.Sequential(*layers)


@torch.no_grad()
def fold_conv_bn(seq: nn.Sequential):
    """Collapse an eval-mode [Conv1x1, (BN), ReLU]* stack into per-layer (W (cout,cin), b (cout)) fp32."""
    out = []
    mods = list(seq)
    i = 0
    while i < len(mods):
        conv = mods[i]
        assert isinstance(conv, (nn.Conv1d, nn.Conv2d)) and all(k == 1 for k in conv.kernel_size)
        w = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).float()
        b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(conv.out_channels, device=w.device)
        i += 1
        if i < len(mods) and isinstance(mods[i], (nn.BatchNorm1d, nn.BatchNorm2d)):
            bn = mods[i]
            scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            w = w * scale[:, None]
            b = (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
            i += 1
        assert i < len(mods) and isinstance(mods[i], nn.ReLU), "fused path expects Conv-(BN)-ReLU triples"
        i += 1
        out.append((w.contiguous(), b.contiguous()))
    return out


PRECISIONS = {"fp32": 0, "bf16": 1, "tf32": 2}  # C ABI codes; bf16 / tf32 = tcgen05 tensor cores


def pack_mlp(layers, *, dense: bool, nsample: int = 1, c_feat: int = 0, c1: int = 0, use_xyz: bool = True,
             precision: str = "bf16"):
    """The tensor path's weight image for folded ``layers`` (C ABI ``tsmdet_mlp_pack_p``; ``precision`` 'bf16' or 'tf32' --
    an image only fits calls of the same precision): a uint8 CUDA tensor to pass as
    ``packed=`` to ``sa_mlp_maxpool`` / ``pointwise_mlp``, or None when the second-generation tcgen05 kernel does not
    take the shape (the unpacked calls then fall back to the first-generation / fp32 kernels).  Build it once per set
    of weights: the packing kernel costs 5-8 us per call, as much as a small layer."""
    from ._lib import TsmdetError

    nl = len(layers)
    cin0 = layers[0][0].shape[1]
    chans = [cin0] + [w.shape[0] for w, _ in layers]
    ch_arr = (ctypes.c_int * (nl + 1))(*chans)
    w_arr = (ctypes.c_void_p * nl)(*[w.data_ptr() for w, _ in layers])
    b_arr = (ctypes.c_void_p * nl)(*[bb.data_ptr() for _, bb in layers])
    size = ctypes.c_longlong(0)
    dev = layers[0][0].device
    args = (PRECISIONS[precision], int(dense), int(nsample), int(c_feat), int(c1), int(use_xyz), nl, ch_arr, w_arr, b_arr)
    try:
        call("tsmdet_mlp_pack_p", *args, None, ctypes.byref(size), stream_ptr(dev))
    except TsmdetError as e:
        if e.code == 1000001:
            return None
        raise
    packed = torch.empty((int(size.value),), dtype=torch.uint8, device=dev)
    call("tsmdet_mlp_pack_p", *args, ptr(packed), ctypes.byref(size), stream_ptr(dev))
    return packed


def _effective_precision(precision: str, packed) -> str:
    """tf32 operands take twice the shared memory of bf16: an MLP whose weights do not fit (pack_mlp -> None, e.g.
    [131,128,128,256]) runs in the fp32 FMA kernel instead -- tighter, slower, never looser than asked for.  (bf16
    without an image goes to the unpacked call, which falls back to the first-generation tcgen05 kernel.)"""
    return "fp32" if precision == "tf32" and packed is None else precision


def sa_mlp_maxpool(xyz, new_xyz, features, idx, idx_cnt, layers, out, out_c0: int, use_xyz: bool = True,
                   precision: str = "fp32", packed: Optional[torch.Tensor] = None,
                   feat_t: Optional[torch.Tensor] = None, out_t: Optional[torch.Tensor] = None):
    """One fused set-abstraction scale (C ABI ``tsmdet_sa_mlp_maxpool`` / ``_packed``): writes
    ``out[:, out_c0:out_c0+cout, :]`` (out is (B, Ctot, npoint) fp32 contiguous).  ``packed`` = ``pack_mlp(layers,
    dense=False, nsample=..., c_feat=..., use_xyz=...)`` skips the per-call weight packing (bf16 path only).
    Stacked layers on the packed bf16 path can chain without transposes: ``out_t`` (B, npoint, round_up(cout, 8)) bf16
    receives the pooled features as rows, and is the next layer's ``feat_t`` (then ``features`` may be None -- pass
    ``c_feat`` through ``feat_t.shape[2]``-compatible layers -- and ``out`` may be None when only the rows are needed)."""
    b, n, _ = xyz.shape
    _, m, s = idx.shape
    if features is not None:
        c_feat = features.shape[1]
    elif feat_t is not None:
        c_feat = layers[0][0].shape[1] - (3 if use_xyz else 0)
        row_dt, row_q = (torch.bfloat16, 8) if precision == "bf16" else (torch.float32, 4)
        assert feat_t.dtype == row_dt and feat_t.shape == (b, n, (c_feat + row_q - 1) // row_q * row_q) and feat_t.is_contiguous()
    else:
        c_feat = 0
    nl = len(layers)
    chans = [(3 if use_xyz else 0) + c_feat] + [w.shape[0] for w, _ in layers]
    for l, (w, bias) in enumerate(layers):
        assert w.shape == (chans[l + 1], chans[l]) and w.is_contiguous() and bias.is_contiguous()
    ch_arr = (ctypes.c_int * (nl + 1))(*chans)
    if packed is not None and precision in ("bf16", "tf32"):
        if out_t is not None:
            row_dt, row_q = (torch.bfloat16, 8) if precision == "bf16" else (torch.float32, 4)
            assert out_t.dtype == row_dt and out_t.is_contiguous() and out_t.shape == (b, m, (chans[-1] + row_q - 1) // row_q * row_q)
        call("tsmdet_sa_mlp_maxpool_packed_p", PRECISIONS[precision], b, n, m, s, c_feat, int(use_xyz), ptr(xyz), ptr(new_xyz), ptr(features),
             ptr(feat_t), ptr(idx), ptr(idx_cnt), nl, ch_arr, ptr(packed), ptr(out), ptr(out_t),
             out.shape[1] if out is not None else 0, out_c0, stream_ptr(xyz.device))
        return out
    assert feat_t is None and out_t is None and out is not None, "row-chained layers need the packed tensor path"
    w_arr = (ctypes.c_void_p * nl)(*[w.data_ptr() for w, _ in layers])
    b_arr = (ctypes.c_void_p * nl)(*[bb.data_ptr() for _, bb in layers])
    prec = PRECISIONS[precision]
    call("tsmdet_sa_mlp_maxpool", b, n, m, s, c_feat, int(use_xyz), ptr(xyz), ptr(new_xyz), ptr(features), ptr(idx),
         ptr(idx_cnt), nl, ch_arr, w_arr, b_arr, ptr(out), out.shape[1], out_c0, prec, stream_ptr(xyz.device))
    return out


def pointwise_mlp(src0: torch.Tensor, src1: Optional[torch.Tensor], layers, out: Optional[torch.Tensor] = None,
                  out_c0: int = 0, precision: str = "fp32", packed: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Point-wise shared MLP, no pooling (C ABI ``tsmdet_pointwise_mlp``): ``out[:, out_c0:out_c0+cout, :] =
    MLP(cat([src0, src1], dim=1))`` for folded ``layers`` [(W (cout,cin), b (cout))]; src0 (B,c0,n), src1 (B,c1,n) |
    None.  The concatenation is never materialised.  ``precision='bf16'`` runs on tcgen05 tensor cores."""
    b, c0, n = src0.shape
    c1 = 0 if src1 is None else src1.shape[1]
    assert src0.is_contiguous() and src0.dtype == torch.float32
    assert src1 is None or (src1.is_contiguous() and src1.shape[0] == b and src1.shape[2] == n)
    nl = len(layers)
    chans = [c0 + c1] + [w.shape[0] for w, _ in layers]
    for l, (w, bias) in enumerate(layers):
        assert w.shape == (chans[l + 1], chans[l]) and w.is_contiguous() and bias.is_contiguous()
    if out is None:
        out = torch.empty((b, chans[-1], n), dtype=torch.float32, device=src0.device)
    assert out.is_contiguous() and out.shape[0] == b and out.shape[2] == n
    ch_arr = (ctypes.c_int * (nl + 1))(*chans)
    if packed is not None and precision in ("bf16", "tf32"):
        call("tsmdet_pointwise_mlp_packed_p", PRECISIONS[precision], b, n, c0, c1, ptr(src0), ptr(src1), nl, ch_arr, ptr(packed), ptr(out),
             out.shape[1], out_c0, stream_ptr(src0.device))
        return out
    w_arr = (ctypes.c_void_p * nl)(*[w.data_ptr() for w, _ in layers])
    b_arr = (ctypes.c_void_p * nl)(*[bb.data_ptr() for _, bb in layers])
    prec = PRECISIONS[precision]
    call("tsmdet_pointwise_mlp", b, n, c0, c1, ptr(src0), ptr(src1), nl, ch_arr, w_arr, b_arr, ptr(out), out.shape[1],
         out_c0, prec, stream_ptr(src0.device))
    return out


def gather_xyz(xyz: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """new_xyz[b,p,:] = xyz[b,idx[b,p],:] -- the transpose/gather_operation/transpose chain (:1143, 1212-1215)."""
    b, n, _ = xyz.shape
    m = idx.shape[1]
    if out is None:
        out = torch.empty((b, m, 3), dtype=torch.float32, device=xyz.device)
    assert out.shape == (b, m, 3) and out.is_contiguous() and out.dtype == torch.float32
    call("tsmdet_gather_xyz", b, n, m, ptr(xyz), ptr(idx), ptr(out), stream_ptr(xyz.device))
    return out


def stage_points(points: torch.Tensor, batch_size: int, xyz: Optional[torch.Tensor] = None,
                 features: Optional[torch.Tensor] = None, bad: Optional[torch.Tensor] = None):
    """Input staging: the collated device batch ``points (B*N, 4+C)`` = [batch_idx, x, y, z,
    features...] -> ``(xyz (B,N,3), features (B,C,N) | None)`` in ONE kernel -- what the reference does with
    ``break_up_pc`` + ``view`` + ``permute(0,2,1).contiguous()`` (pointnet2_backbone.py:796-800, 814-823) and the
    ``xyz.transpose(1,2).contiguous()`` of pointnet2_modules.py:1143.  ``bad`` (1,) int32, if given, is incremented
    on the device for every tile holding a row whose batch index is not its frame (the reference asserts equal
    per-frame counts, :819); without it the check is skipped.  C ABI ``tsmdet_stage_points``."""
    assert points.is_cuda and points.dtype == torch.float32 and points.is_contiguous() and points.dim() == 2
    rows, w = points.shape
    assert w >= 4 and rows % batch_size == 0, "every frame must hold the same number of points"
    n, c = rows // batch_size, w - 4
    if xyz is None:
        xyz = torch.empty((batch_size, n, 3), dtype=torch.float32, device=points.device)
    if features is None and c > 0:
        features = torch.empty((batch_size, c, n), dtype=torch.float32, device=points.device)
    assert xyz.shape == (batch_size, n, 3) and xyz.is_contiguous()
    assert c == 0 or (features.shape == (batch_size, c, n) and features.is_contiguous())
    call("tsmdet_stage_points", batch_size, n, c, ptr(points), ptr(xyz), ptr(features) if c > 0 else None,
         ptr(bad), stream_ptr(points.device))
    return xyz, (features if c > 0 else None)


class PointnetFPModule(nn.Module):
    r"""Propagates the features of one set to another (ref :130-178)."""

    def __init__(self, *, mlp: List[int], bn: bool = True, fused: bool = True, precision: str = "fp32"):
        """``fused`` (eval mode only): BatchNorm folded into the 1x1 convs and the whole MLP -- including the
        concatenation of the interpolated and the skip features (:171) -- in ONE kernel (``tsmdet_pointwise_mlp``:
        fp32 FMA, or bf16 tcgen05 with ``precision='bf16'``); otherwise the reference's eager Conv2d/BatchNorm2d."""
        super().__init__()
        self.mlp = build_shared_mlp(mlp, bn=True) if bn else build_shared_mlp(mlp, bn=False)
        self.fused = fused
        self.precision = precision
        self._folded = None
        self._packed = None

    def train(self, mode: bool = True):
        self._folded = None
        return super().train(mode)

    def _load_from_state_dict(self, *args, **kwargs):
        self._folded = None  # weights changed under a cached fold / packed image
        return super()._load_from_state_dict(*args, **kwargs)

    def forward(self, unknown, known, unknow_feats, known_feats):
        """unknown (B,n,3), known (B,m,3), unknow_feats (B,C1,n)|None, known_feats (B,C2,m) -> (B,mlp[-1],n)"""
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_recip, dim=2, keepdim=True)
            weight = dist_recip / norm
            interpolated_feats = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated_feats = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))

        if self.fused and not self.training:
            c0 = interpolated_feats.shape[1]
            c1 = 0 if unknow_feats is None else unknow_feats.shape[1]
            if self._folded is None:
                self._folded = fold_conv_bn(self.mlp)
                self._packed = (pack_mlp(self._folded, dense=True, c_feat=c0, c1=c1, precision=self.precision)
                                if self.precision in ("bf16", "tf32") else None)
            return pointwise_mlp(interpolated_feats.contiguous(),
                                 None if unknow_feats is None else unknow_feats.contiguous(), self._folded,
                                 precision=_effective_precision(self.precision, self._packed), packed=self._packed)
        if unknow_feats is not None:
            new_features = torch.cat([interpolated_feats, unknow_feats], dim=1)
        else:
            new_features = interpolated_feats
        new_features = self.mlp(new_features.unsqueeze(-1))
        return new_features.squeeze(-1)


class PointnetSAModuleFSMSG(nn.Module):
    """Fusion-sampling, multi-scale-grouping set abstraction on the dense-batch ops (reference layer 0).

    Constructor arguments follow ``VoxelPointnetSAModuleFSMSGDistillation.__init__`` (:1442-1466) for the
    options that the layer-0 branch reads; ``mlps[i]`` lists the channels WITHOUT the +3 for xyz, as in
    the reference's config (it adds 3 itself when ``use_xyz``, :1538-1540).
    """

    def __init__(self, *, npoint_list: List[int] = None, sample_range_list: List[List[int]] = None,
                 sample_method_list: List[str] = None, radii: List[float], nsamples: List[int],
                 mlps: List[List[int]], bn: bool = True, use_xyz: bool = True, pool_method: str = 'max_pool',
                 dilated_radius_group: bool = False, skip_connection: bool = False, weight_gamma: float = 1.0,
                 aggregation_mlp: Optional[List[int]] = None, fused: bool = True, precision: str = "fp32"):
        super().__init__()
        assert npoint_list is None or len(npoint_list) == len(sample_range_list) == len(sample_method_list)
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint_list = npoint_list
        self.sample_range_list = sample_range_list
        self.sample_method_list = sample_method_list
        self.radii, self.nsamples = list(radii), list(nsamples)
        self.use_xyz = use_xyz
        self.pool_method = pool_method
        self.dilated_radius_group = dilated_radius_group
        self.skip_connection = skip_connection
        self.weight_gamma = weight_gamma
        self.fused = fused
        self.precision = precision

        self.groupers = nn.ModuleList()
        self.point_mlps = nn.ModuleList()
        former_radius = 0.0
        in_channels, out_channels = 0, 0
        for radius, nsample, spec in zip(radii, nsamples, mlps):
            if dilated_radius_group:
                self.groupers.append(pointnet2_utils.QueryAndGroupDilated(former_radius, radius, nsample, use_xyz=use_xyz))
            else:
                self.groupers.append(pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz))
            former_radius = radius
            spec = list(spec)
            in_channels = spec[0]
            if use_xyz:
                spec[0] += 3
            self.point_mlps.append(build_shared_mlp(spec, bn=bn))
            out_channels += spec[-1]
        self.feature_channels = in_channels
        if skip_connection:
            out_channels += in_channels
        if aggregation_mlp:
            self.aggregation_mlp = build_shared_mlp([out_channels] + list(aggregation_mlp), bn=bn, dims=1)
            out_channels = aggregation_mlp[-1]
        else:
            self.aggregation_mlp = None
        self.out_channels = out_channels
        self._folded = None  # cache of folded (W, b) per scale, built lazily in eval mode
        self._folded_agg = None
        self._packed = None
        self._packed_agg = None

    def _load_from_state_dict(self, *args, **kwargs):
        # a checkpoint loaded after a warm-up forward must not leave the fused path on stale folded weights
        # (ADVICE r1); CUDA graphs captured before the load hold pointers to the old folds and must be re-captured
        self._folded = None
        self._folded_agg = None
        return super()._load_from_state_dict(*args, **kwargs)

    # ------------------------------------------------------------------ sampling (ref :1153-1210)
    def sample(self, xyz, features=None, scores=None):
        sample_idx_list = []
        for (lo, hi), method, npoint in zip(self.sample_range_list, self.sample_method_list, self.npoint_list):
            xyz_slice = xyz[:, lo:hi, :].contiguous()
            if method == 'd-fps':
                sample_idx = pointnet2_utils.furthest_point_sample(xyz_slice, npoint)
            elif method == 'f-fps':
                features_slice = features[:, :, lo:hi]
                dist_matrix = pointnet2_utils.calc_dist_matrix_for_sampling(
                    xyz_slice, features_slice.permute(0, 2, 1), self.weight_gamma)
                sample_idx = pointnet2_utils.furthest_point_sample_matrix(dist_matrix.contiguous(), npoint)
            elif method == 's-fps':
                assert scores is not None
                scores_slice = scores[:, lo:hi].contiguous()
                scores_slice = scores_slice.sigmoid() ** self.weight_gamma
                sample_idx = pointnet2_utils.furthest_point_sample_weights(xyz_slice, scores_slice, npoint)
            elif method == 's-topk':
                assert scores is not None
                _, sample_idx = torch.topk(scores[:, lo:hi], k=npoint, dim=-1)
                sample_idx = sample_idx.int()
            else:
                raise NotImplementedError(method)
            sample_idx_list.append(sample_idx + lo)
        return torch.cat(sample_idx_list, dim=-1).contiguous()

    def _folded_layers(self):
        if self._folded is None:
            self._folded = [fold_conv_bn(m) for m in self.point_mlps]
            self._packed = None
        return self._folded

    def _packed_layers(self, c_feat: int, use_xyz: bool):
        """Per scale: the tensor path's weight image of the folded MLP (None: shape not taken / fp32 mode), built once
        per fold -- CUDA graphs captured afterwards hold its address, like the folded weights'."""
        folded = self._folded_layers()
        if self._packed is None or self._packed[0] != (c_feat, use_xyz, self.precision):
            imgs = [pack_mlp(layers, dense=False, nsample=g.nsample, c_feat=c_feat, use_xyz=use_xyz, precision=self.precision)
                    if self.precision in ("bf16", "tf32") else None for layers, g in zip(folded, self.groupers)]
            self._packed = ((c_feat, use_xyz, self.precision), imgs)
        return self._packed[1]

    def train(self, mode: bool = True):
        self._folded = None
        self._folded_agg = None
        return super().train(mode)

    # ------------------------------------------------------------------ forward
    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, new_xyz: torch.Tensor = None,
                scores: torch.Tensor = None):
        """xyz (B,N,3), features (B,C,N)|None, scores (B,N)|None
        -> (new_xyz (B,npoint,3), new_features (B,Cout,npoint), sample_idx (B,npoint) int32 | None)"""
        sample_idx = None
        old_features = None
        if new_xyz is None:
            sample_idx = self.sample(xyz, features, scores)
            new_xyz = gather_xyz(xyz.contiguous(), sample_idx)
            if self.skip_connection and features is not None:
                old_features = pointnet2_utils.gather_operation(features.contiguous(), sample_idx)

        use_fused = self.fused and not self.training and self.pool_method == 'max_pool'
        b, npoint, _ = new_xyz.shape
        if use_fused:
            folded = self._folded_layers()
            widths = [layers[-1][0].shape[0] for layers in folded]
            extra = old_features.shape[1] if old_features is not None else 0
            new_features = torch.empty((b, sum(widths) + extra, npoint), dtype=torch.float32, device=xyz.device)
            c0 = 0
            eff_xyz = self.use_xyz or features is None
            packed = self._packed_layers(0 if features is None else features.shape[1], eff_xyz)
            for grouper, layers, w, img in zip(self.groupers, folded, widths, packed):
                if isinstance(grouper, pointnet2_utils.QueryAndGroupDilated):
                    idx_cnt, idx = pointnet2_utils.ball_query_dilated(
                        grouper.radius_in, grouper.radius_out, grouper.nsample, xyz, new_xyz)
                else:
                    idx_cnt, idx = pointnet2_utils.ball_query(grouper.radius, grouper.nsample, xyz, new_xyz)
                sa_mlp_maxpool(xyz, new_xyz, features, idx, idx_cnt, layers, new_features, c0,
                               use_xyz=eff_xyz, precision=_effective_precision(self.precision, img), packed=img)
                c0 += w
            if old_features is not None:
                new_features[:, c0:] = old_features
        else:
            new_features_list = []
            for grouper, mlp in zip(self.groupers, self.point_mlps):
                idx_cnt, grouped_features, _ = grouper(xyz, new_xyz, features)  # (B, C, npoint, nsample)
                idx_cnt_mask = (idx_cnt > 0).float().unsqueeze(1).unsqueeze(-1)   # (B, 1, npoint, 1)
                grouped_features = grouped_features * idx_cnt_mask                # mask the INPUT (ref :1265-1267)
                y = mlp(grouped_features)
                if self.pool_method == 'max_pool':
                    y = F.max_pool2d(y, kernel_size=[1, y.size(3)])
                elif self.pool_method == 'avg_pool':
                    y = F.avg_pool2d(y, kernel_size=[1, y.size(3)])
                else:
                    raise NotImplementedError(self.pool_method)
                new_features_list.append(y.squeeze(-1))
            if old_features is not None:
                new_features_list.append(old_features)
            new_features = torch.cat(new_features_list, dim=1)

        if self.aggregation_mlp is not None:
            if use_fused:  # ref :1320-1321 as one kernel (BN folded; tensor cores in bf16 mode)
                if self._folded_agg is None:
                    self._folded_agg = fold_conv_bn(self.aggregation_mlp)
                    self._packed_agg = (pack_mlp(self._folded_agg, dense=True, c_feat=new_features.shape[1],
                                                 precision=self.precision)
                                        if self.precision in ("bf16", "tf32") else None)
                new_features = pointwise_mlp(new_features.contiguous(), None, self._folded_agg,
                                             precision=_effective_precision(self.precision, self._packed_agg),
                                             packed=self._packed_agg)
            else:
                new_features = self.aggregation_mlp(new_features)
        return new_xyz.contiguous(), new_features.contiguous(), sample_idx


# The name the reference's backbone config resolves for this layer (sa_layer_idx == 0 only).
VoxelPointnetSAModuleFSMSGDistillation = PointnetSAModuleFSMSG


class PointNet2SAStack(nn.Module):
    """A plain stack of single- or multi-scale SA layers -- BASELINE.json config 2
    (16384 -> 4096 -> 1024 -> 512, radii 0.2/0.8/1.6, nsample 16/32/32)."""

    def __init__(self, npoints: Sequence[int], radii: Sequence[Sequence[float]], nsamples: Sequence[Sequence[int]],
                 mlps: Sequence[Sequence[Sequence[int]]], in_channels: int = 1, fused: bool = True,
                 precision: str = "fp32", chain_fps: bool = True):
        super().__init__()
        self.chain_fps = chain_fps
        self.layers = nn.ModuleList()
        c = in_channels
        for npoint, r, ns, specs in zip(npoints, radii, nsamples, mlps):
            specs = [[c] + list(s) for s in specs]
            layer = PointnetSAModuleFSMSG(
                npoint_list=[npoint], sample_range_list=[[0, None]], sample_method_list=['d-fps'],
                radii=list(r), nsamples=list(ns), mlps=specs, fused=fused, precision=precision)
            self.layers.append(layer)
            c = layer.out_channels
        self.out_channels = c

    def forward(self, xyz, features=None):
        outs = []
        state = None
        for layer in self.layers:
            # every layer here is one d-fps over all of the previous layer's centres: chain the samplers
            if self.chain_fps:
                idx, state = pointnet2_utils.farthest_point_sample_chained(xyz.contiguous(), layer.npoint_list[0], state)
            else:
                idx = pointnet2_utils.farthest_point_sample(xyz.contiguous(), layer.npoint_list[0])
            new_xyz = gather_xyz(xyz.contiguous(), idx)
            xyz, features, _ = layer(xyz, features, new_xyz=new_xyz)
            outs.append((xyz, features, idx))
        return outs


def kitti_sa_stack(fused: bool = True, precision: str = "fp32") -> PointNet2SAStack:
    """L1 16384->4096 r=0.2 ns=16 [1+3,16,16,32]; L2 4096->1024 r=0.8 ns=32
    [32+3,64,64,128]; L3 1024->512 r=1.6 ns=32 [128+3,128,128,256]."""
    return PointNet2SAStack(
        npoints=[4096, 1024, 512], radii=[[0.2], [0.8], [1.6]], nsamples=[[16], [32], [32]],
        mlps=[[[16, 16, 32]], [[64, 64, 128]], [[128, 128, 256]]], in_channels=1, fused=fused, precision=precision)

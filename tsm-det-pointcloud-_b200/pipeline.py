"""The benchmarked hot path as one callable: SA backbone (BASELINE.json config 2) + rotated NMS
(config 3: 4096 proposals per frame, IoU 0.01 then 0.1 on the survivors), frames sharded per GPU.

Execution model (B200-first: streams + CUDA graphs instead of a tracing compiler):

  * the three FPS launches form a serial latency chain (FPS L2 samples the centres FPS L1 picked),
    but they occupy only one 128-thread CTA per SM, so ball query, the tcgen05 MLP kernels and the NMS
    kernels run BESIDE them on other streams:
        stream A : [stage] fps1 -> gather1 -> fps2 -> gather2 -> fps3 -> gather3
        stream B :          [e1] query1 -> mlp1 -> [e2] query2 -> mlp2 -> [e3] query3 -> mlp3
        stream C : sort -> nms(0.01) -> compact -> nms(0.1) -> records
  * the whole DAG is captured once into a CUDA graph and replayed per step (one launch, no Python
    between kernels, no allocator traffic).

Two front doors:

``forward_device``  device-resident SoA tensors ``xyz (F,N,3), feats (F,C,N), boxes (F,P,7), scores (F,P)``
                    (bench ``value``).
``forward_host``    the reference-facing call with HOST buffers (bench ``e2e``): the batch arrives the way the
                    reference's data loader delivers it -- ONE collated array ``points (F*N, 1+3+C)`` =
                    [batch_idx, x, y, z, features...] (pcdet/models/__init__.py:23-34) -- next to the proposals,
                    in one pinned buffer (``HostIO``).  Per step: ONE H2D copy, the graph (whose first node is the
                    staging kernel, csrc/staging.cu: collated rows -> xyz (F,N,3) + features (F,C,N), replacing
                    break_up_pc / view / permute / transpose of pointnet2_backbone.py:796-826 and
                    pointnet2_modules.py:1143), and ONE D2H copy of the packed results
                    [features | xyz | detection records + counts] (rank 0 adds one copy of the gathered records,
                    as the reference's merge_results_dist returns them on rank 0 only, common_utils.py:224-245).
"""
from __future__ import annotations

import os
import sys
from typing import Dict, Optional

import torch

from . import _lib, iou3d_nms_utils, pointnet2_utils
from .pointnet2_modules import _effective_precision, gather_xyz, kitti_sa_stack, sa_mlp_maxpool, stage_points
from .sharding import PeerGather, gather_detections, gather_packed, pack_detections


class HostIO:
    """Pinned host buffers of one step, laid out so that a step costs ONE copy per direction.

    ``inp`` (flat f32)  = [points (F*N, 4+C) | boxes (F,P,7) | scores (F,P)]  -- fill the views in place
    ``out`` (flat f32)  = [features (F,Cout,M) | xyz (F,M,3) | det (F,K,9) | det_num (F) as i32 bits]
    ``all`` (flat f32)  = the gathered records of every rank (world, F*K*9 + F) -- copied on rank 0 only
    """

    def __init__(self, frames: int, n_points: int, c_feat: int, n_props: int, c_out: int, m_out: int, k_post: int,
                 world: int = 1):
        self.frames, self.n_points, self.c_feat, self.n_props = frames, n_points, c_feat, n_props
        w = 4 + c_feat
        n_pts, n_box, n_sc = frames * n_points * w, frames * n_props * 7, frames * n_props
        self.inp = torch.empty((n_pts + n_box + n_sc,), dtype=torch.float32).pin_memory()
        self.points = self.inp[:n_pts].view(frames * n_points, w)
        self.boxes = self.inp[n_pts:n_pts + n_box].view(frames, n_props, 7)
        self.scores = self.inp[n_pts + n_box:].view(frames, n_props)
        n_f, n_x, n_d = frames * c_out * m_out, frames * m_out * 3, frames * k_post * 9
        self.out = torch.empty((n_f + n_x + n_d + frames,), dtype=torch.float32).pin_memory()
        self.features = self.out[:n_f].view(frames, c_out, m_out)
        self.xyz = self.out[n_f:n_f + n_x].view(frames, m_out, 3)
        self.det = self.out[n_f + n_x:n_f + n_x + n_d].view(frames, k_post, 9)
        self.det_num = self.out[n_f + n_x + n_d:].view(torch.int32)
        self.det_packed = self.out[n_f + n_x:]
        self.all = torch.empty((world, n_d + frames), dtype=torch.float32).pin_memory() if world > 1 else None
        self.all_det = self.all[:, :n_d].view(world, frames, k_post, 9) if world > 1 else self.det.unsqueeze(0)
        self.all_num = self.all[:, n_d:].view(torch.int32) if world > 1 else self.det_num.unsqueeze(0)

    def fill(self, xyz, feats, boxes, scores):
        """Collate SoA host arrays into the reference's ``points`` layout (what its data loader hands over)."""
        f, n, _ = xyz.shape
        pts = self.points.view(f, n, -1)
        pts[:, :, 0] = torch.arange(f, dtype=torch.float32).view(f, 1)
        pts[:, :, 1:4] = torch.as_tensor(xyz)
        if self.c_feat:
            pts[:, :, 4:] = torch.as_tensor(feats).permute(0, 2, 1)
        self.boxes.copy_(torch.as_tensor(boxes)[:, :, :7])
        self.scores.copy_(torch.as_tensor(scores))
        return self

    @property
    def h2d_bytes(self) -> int:
        return self.inp.numel() * 4

    def d2h_bytes(self, rank0: bool) -> int:
        return self.out.numel() * 4 + (self.all.numel() * 4 if (self.all is not None and rank0) else 0)


class SABackboneNMS(torch.nn.Module):
    def __init__(self, precision: str = "bf16", nms_pre: float = 0.01, nms_post: float = 0.1, k_post: int = 512,
                 seed: int = 0, use_graph: bool = True, chain_fps: bool = False, gather_group=None):
        super().__init__()
        # torch.distributed group this engine's detection gather runs on, INSIDE the step (and so inside its
        # CUDA graph: no per-step host-side collective launch).  Engines that run concurrently need a group
        # each -- collectives of one communicator must execute in one order on every rank.
        self.gather_group = gather_group
        torch.manual_seed(seed)
        self.backbone = kitti_sa_stack(fused=True, precision=precision)
        g = torch.Generator().manual_seed(seed + 1)
        for m in self.backbone.modules():  # randomised BN running stats
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
        self.precision = precision
        self.nms_pre, self.nms_post, self.k_post = nms_pre, nms_post, k_post
        self.use_graph = use_graph
        # chained samplers (tsmdet_fps_chain) skip levels 2/3 when the level-1 run proves it exact, but the
        # bookkeeping costs ~17 % of level 1 and one tied cloud in the batch forfeits the saving: measured a wash
        # at 16 clouds per step, a win for small batches -- so it is opt-in here.
        self.chain_fps = chain_fps
        self.nms_after_fps = os.environ.get("TSMDET_NMS_AFTER_FPS", "0") != "0"
        self._graphs: Dict[tuple, dict] = {}
        self._streams = None
        self._trace = None  # list of (name, event) when tracing (see trace_step)
        self._pg = None
        self._pg_tried = False
        self._cap_stream = None
        self.eval()

    # ------------------------------------------------------------------ the DAG
    def _side_streams(self, dev):
        if self._streams is None:
            self._streams = [torch.cuda.Stream(dev) for _ in range(2)]
        return self._streams

    @torch.no_grad()
    def _nms_two_pass(self, boxes, scores):
        """IoU `nms_pre` over all proposals, then `nms_post` over the survivors; (F,k) indices + counts."""
        f, p = scores.shape
        sel1, num1 = iou3d_nms_utils.nms_gpu_batch(boxes, scores, self.nms_pre)
        valid = sel1 >= 0
        safe = torch.where(valid, sel1, torch.zeros_like(sel1))
        boxes2 = torch.gather(boxes, 1, safe.unsqueeze(-1).expand(-1, -1, boxes.size(2)))
        scores2 = torch.where(valid, torch.gather(scores, 1, safe), torch.full_like(scores, float("-inf")))
        # survivors are already in score order: no second sort needed
        sel2, num2 = iou3d_nms_utils.nms_gpu_batch(boxes2, scores2, self.nms_post, counts=num1, presorted=True)
        valid2 = sel2 >= 0
        final = torch.where(valid2, torch.gather(safe, 1, torch.where(valid2, sel2, torch.zeros_like(sel2))),
                            torch.full_like(sel2, -1))
        k = min(self.k_post, p)
        return final[:, :k].contiguous(), torch.clamp(num2, max=k)

    def _out_layout(self, frames: int, p: int):
        """Sections of the packed result buffer: (n_features, n_xyz, n_det, c_out, m_out, k)."""
        last = self.backbone.layers[-1]
        c_out, m_out = last.out_channels, last.npoint_list[0]
        k = min(self.k_post, p)
        return frames * c_out * m_out, frames * m_out * 3, frames * k * 9, c_out, m_out, k

    @torch.no_grad()
    def _run(self, xyz, feats, boxes, scores, points=None):
        """Multi-stream DAG; everything it launches is ordered after / joined back into the current stream.
        With ``points`` (the collated (F*N, 4+C) batch) the first node is the staging kernel, which fills
        ``xyz`` / ``feats``.  The results land in ONE packed buffer [features | xyz | det | det_num]."""
        dev = xyz.device
        main = torch.cuda.current_stream(dev)
        s_sa, s_nms = self._side_streams(dev)

        trace = self._trace
        def mark(name, stream):
            if trace is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                trace.append((name, ev))
        mark("start", main)
        layers = self.backbone.layers
        b = xyz.shape[0]
        n_f, n_x, n_d, c_out, m_out, k = self._out_layout(b, scores.shape[1])
        pack = torch.empty((n_f + n_x + n_d + b,), dtype=torch.float32, device=dev)
        out_features = pack[:n_f].view(b, c_out, m_out)
        out_xyz = pack[n_f:n_f + n_x].view(b, m_out, 3)
        out_det = pack[n_f + n_x:n_f + n_x + n_d].view(b, k, 9)
        out_num = pack[n_f + n_x + n_d:].view(torch.int32)
        bad = None
        s_nms.wait_stream(main)  # NMS needs only the proposals: it starts before the staging kernel is done
        if points is not None:
            bad = torch.zeros((1,), dtype=torch.int32, device=dev)
            stage_points(points, b, xyz, feats, bad)
            mark("stage", main)
        s_sa.wait_stream(main)
        cur_xyz = xyz
        centres, ready = [], []
        # stream A (current): the FPS chain
        state = None
        for li, layer in enumerate(layers):
            if self.chain_fps:
                idx, state = pointnet2_utils.farthest_point_sample_chained(cur_xyz, layer.npoint_list[0], state)
            else:
                idx = pointnet2_utils.farthest_point_sample(cur_xyz, layer.npoint_list[0])
            new_xyz = gather_xyz(cur_xyz, idx, out=out_xyz if li == len(layers) - 1 else None)
            mark(f"fps{len(ready) + 1}", main)
            ev = torch.cuda.Event()
            ev.record(main)
            centres.append((cur_xyz, new_xyz))
            ready.append(ev)
            cur_xyz = new_xyz
        # stream B: query + fused MLP per layer, as soon as that layer's centres exist
        with torch.cuda.stream(s_sa):
            cur_f, cur_t = feats, None  # fp32 (B,C,N) planes | bf16 (B,N,C) rows written by the previous layer
            imgs = [l._packed_layers(l._folded_layers()[0][0][0].shape[1] - 3, True)[0] for l in layers]  # built once
            for li, (layer, (src_xyz, new_xyz), ev) in enumerate(zip(layers, centres, ready)):
                s_sa.wait_event(ev)
                g = layer.groupers[0]
                cnt, bidx = pointnet2_utils.ball_query(g.radius, g.nsample, src_xyz, new_xyz)
                mark(f"query{src_xyz.shape[1]}", s_sa)
                folded = layer._folded_layers()[0]
                img = imgs[li]
                last = li == len(layers) - 1
                cout = folded[-1][0].shape[0]
                # Intermediate layers of the packed bf16 path hand their features on as bf16 ROWS -- exactly what the
                # next layer's gather reads -- so no (B,C,M) fp32 tensor is written and no transpose kernel runs.
                chain = img is not None and not last and imgs[li + 1] is not None
                out = out_features if last else (None if chain else torch.empty((b, cout, new_xyz.shape[1]),
                                                                              dtype=torch.float32, device=dev))
                row_dt, row_q = (torch.bfloat16, 8) if layer.precision == "bf16" else (torch.float32, 4)  # tf32: fp32 rows
                out_t = torch.empty((b, new_xyz.shape[1], (cout + row_q - 1) // row_q * row_q), dtype=row_dt,
                                    device=dev) if chain else None
                sa_mlp_maxpool(src_xyz, new_xyz, cur_f, bidx, cnt, folded, out, 0,
                               precision=_effective_precision(layer.precision, img), packed=img, feat_t=cur_t, out_t=out_t)
                mark(f"mlp{src_xyz.shape[1]}", s_sa)
                cur_f, cur_t = out, out_t
        # stream C: NMS is independent of the backbone.  (Measured inside the captured graph: letting it run
        # beside the sampling chain is faster -- 4.98 vs 5.61 ms/step -- than holding it back until the chain
        # is done; TSMDET_NMS_AFTER_FPS=1 selects the latter for experiments.)
        with torch.cuda.stream(s_nms):
            if self.nms_after_fps:
                s_nms.wait_event(ready[-1])
            det_idx, det_num = self._nms_two_pass(boxes, scores)
            mark("nms", s_nms)
            # fixed-size detection records (F, K, 9) = box(7), score, 0 -- what the ranks exchange at N > 1
            live = det_idx >= 0
            safe_k = torch.where(live, det_idx, torch.zeros_like(det_idx))
            out_det.zero_()
            out_det[:, :, :7] = torch.gather(boxes[:, :, :7], 1, safe_k.unsqueeze(-1).expand(-1, -1, 7))
            out_det[:, :, 7] = torch.gather(scores, 1, safe_k)
            out_det.mul_(live.unsqueeze(-1))
            out_num.copy_(det_num)
        main.wait_stream(s_sa)
        main.wait_stream(s_nms)
        res = {"xyz": out_xyz, "features": out_features, "det_idx": det_idx, "det_num": out_num, "det": out_det,
               "det_packed": pack[n_f + n_x:],  # records + counts: the collective's send buffer, contiguous
               "packed": pack}
        if bad is not None:
            res["bad_rows"] = bad
        if self.gather_group is not None:
            # the one collective of the path: every rank receives every rank's records (equal shards, so the
            # sizes are known without a size exchange or a host sync)
            world = torch.distributed.get_world_size(self.gather_group)
            res["all_det"], res["all_num"] = gather_detections(out_det, out_num, frames_total=b * world,
                                                               group=self.gather_group)
        mark("end", main)
        return res

    def trace_step(self, xyz, feats, boxes, scores):
        """Runs the DAG eagerly with timing events; returns [(name, ms since start)] (diagnostics)."""
        self._run(xyz, feats, boxes, scores)
        torch.cuda.synchronize(xyz.device)
        self._trace = []
        try:
            self._run(xyz, feats, boxes, scores)
            torch.cuda.synchronize(xyz.device)
            t0 = self._trace[0][1]
            return [(n, t0.elapsed_time(e)) for n, e in self._trace]
        finally:
            self._trace = None

    # ------------------------------------------------------------------ graph capture / replay
    def _capture(self, key, static_in, run, keep=()):
        dev = static_in[0].device
        # the sampling chain is captured on the capture stream itself: TSMDET_FPS_PRIORITY=1 makes it a high-priority
        # stream (its kernels' graph nodes inherit the priority), so FPS CTAs are placed first when SMs free up.
        # ONE capture stream per engine, like its side streams: the library's scratch buffers are keyed by stream, and
        # torch hands out streams round-robin from a pool of 32 -- a fresh stream per capture would eventually alias
        # another lane's stream and share its scratch with a concurrently replaying graph.
        if self._cap_stream is None:
            self._cap_stream = torch.cuda.Stream(dev, priority=-1 if os.environ.get("TSMDET_FPS_PRIORITY", "0") == "1" else 0)
        cap = self._cap_stream
        cap.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap):
            for _ in range(2):  # warm-up on the capture stream: grows every scratch buffer, folds BN, plans FPS
                run()
        cap.synchronize()
        graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count
        with torch.cuda.graph(graph, stream=cap):
            out = run()
        torch.cuda.current_stream(dev).wait_stream(cap)
        # kernels of this library inside one replay (the graph also holds a few torch sort/gather kernels)
        # `keep`: tensors the captured kernels address that nothing else references (the graph holds raw pointers)
        ent = {"graph": graph, "in": static_in, "out": out, "launches": _lib.launch_count - l0, "keep": keep}
        self._graphs[key] = ent
        return ent

    def _graph_for(self, xyz, feats, boxes, scores):
        key = (xyz.device.index, tuple(xyz.shape), tuple(feats.shape), tuple(boxes.shape), tuple(scores.shape))
        ent = self._graphs.get(key)
        if ent is not None:
            return ent
        static_in = [torch.empty_like(t) for t in (xyz, feats, boxes, scores)]
        for s, t in zip(static_in, (xyz, feats, boxes, scores)):
            s.copy_(t)
        return self._capture(key, static_in, lambda: self._run(*static_in))

    def _graph_for_host(self, io: HostIO, dev):
        """The staged variant: ONE flat device input [points | boxes | scores]; the graph starts with the staging
        kernel."""
        key = ("host", dev.index, io.frames, io.n_points, io.c_feat, io.n_props)
        ent = self._graphs.get(key)
        if ent is not None:
            return ent
        flat = torch.empty(io.inp.shape, dtype=torch.float32, device=dev)
        flat.copy_(io.inp)
        f, n, c, p = io.frames, io.n_points, io.c_feat, io.n_props
        n_pts, n_box = f * n * (4 + c), f * p * 7
        points = flat[:n_pts].view(f * n, 4 + c)
        boxes = flat[n_pts:n_pts + n_box].view(f, p, 7)
        scores = flat[n_pts + n_box:].view(f, p)
        xyz = torch.empty((f, n, 3), dtype=torch.float32, device=dev)
        feats = torch.empty((f, c, n), dtype=torch.float32, device=dev)
        return self._capture(key, [flat], lambda: self._run(xyz, feats, boxes, scores, points=points),
                             keep=(xyz, feats, points, boxes, scores))

    def _gather(self, res):
        """The one collective of the path, issued after the replay: equal shards, so a single exchange of the packed
        buffer the graph produced -- no size exchange, no host sync.  Returns once the CURRENT STREAM is ordered
        after the arrival of every rank's records (both transports)."""
        if not (torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1):
            res["all_det"], res["all_num"] = res["det"].unsqueeze(0), res["det_num"].unsqueeze(0)
            return res
        f, k = res["det"].shape[0], res["det"].shape[1]
        pg = self._peer_gather(res["det_packed"])
        if pg is not None:  # one-sided stores into every rank's receive ring (sharding.PeerGather)
            res["all_packed"] = pg.put(res["det_packed"])[:, :pg.numel]
            pg.wait_stream()  # the views below are only valid once the peers' stores have landed
            res["all_det"], res["all_num"] = pg.views(f, k)
        else:
            res["all_det"], res["all_num"], self._gather_out = gather_packed(
                res["det_packed"], f, k, out=getattr(self, "_gather_out", None))
            res["all_packed"] = self._gather_out
        return res

    @torch.no_grad()
    def forward_device(self, xyz, feats, boxes, scores, gather: bool = False) -> Dict[str, torch.Tensor]:
        """xyz (F,N,3), feats (F,C,N), boxes (F,P,7), scores (F,P) on the GPU.  The returned tensors are
        owned by the engine and overwritten by the next call with the same shapes (graph replay); with
        ``gather=True`` they include every rank's records (``all_det`` / ``all_num``), complete in stream order."""
        if self.use_graph:
            ent = self._graph_for(xyz, feats, boxes, scores)
            for s, t in zip(ent["in"], (xyz, feats, boxes, scores)):
                if s.data_ptr() != t.data_ptr():
                    s.copy_(t, non_blocking=True)
            ent["graph"].replay()
            _lib.launch_count += ent["launches"]
            res = dict(ent["out"])
        else:
            res = self._run(xyz, feats, boxes, scores)
        if gather and "all_det" not in res:
            self._gather(res)
        return res

    def _peer_gather(self, packed):
        """The transport of the detection gather: peer-memory stores (TSMDET_GATHER=peer, the default on GPUs) or the
        NCCL all_gather (TSMDET_GATHER=nccl, and whenever the IPC set-up fails on ANY rank -- PeerGather.create is
        collective and every rank takes the same branch)."""
        if not self._pg_tried:
            self._pg_tried = True
            if os.environ.get("TSMDET_GATHER", "peer") == "peer" and packed.is_cuda:
                self._pg, err = PeerGather.create(packed.numel(), packed.device)
                if self._pg is None:
                    print(f"[tsmdet] peer-memory gather unavailable ({err}); using the NCCL all_gather", file=sys.stderr)
        return self._pg

    @staticmethod
    def check_staged(res):
        """The reference asserts equal per-frame point counts (pointnet2_backbone.py:819); the staging kernel counts
        violations on the device -- this reads the counter (a host sync)."""
        bad = int(res["bad_rows"].item())
        if bad != 0:
            raise ValueError(f"points: {bad} tile(s) hold rows whose batch index is not their frame")

    def static_inputs(self, xyz, feats, boxes, scores):
        """The engine-owned input buffers for these shapes (write into them to skip the D2D copy)."""
        return self._graph_for(xyz, feats, boxes, scores)["in"]

    def host_io(self, frames: int, n_points: int, c_feat: int, n_props: int) -> HostIO:
        world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        _, _, _, c_out, m_out, k = self._out_layout(frames, n_props)
        return HostIO(frames, n_points, c_feat, n_props, c_out, m_out, k, world)

    @torch.no_grad()
    def forward_host(self, io: HostIO, gather: bool = False, sync: bool = True) -> HostIO:
        """Reference-facing call: pinned host buffers in (``io.points / boxes / scores``), pinned host results out
        (``io.features / xyz / det / det_num`` and, on rank 0, ``io.all_det / all_num``).  ONE H2D copy, the
        captured step, ONE D2H copy (+ one on rank 0 for the gathered records)."""
        dev = next(self.parameters()).device
        ent = self._graph_for_host(io, dev)
        ent["in"][0].copy_(io.inp, non_blocking=True)
        ent["graph"].replay()
        _lib.launch_count += ent["launches"]
        res = dict(ent["out"])
        if gather:
            self._gather(res)
        io.out.copy_(res["packed"], non_blocking=True)
        if gather and io.all is not None and torch.distributed.get_rank() == 0:
            io.all.copy_(res["all_packed"], non_blocking=True)
        if sync:
            torch.cuda.current_stream(dev).synchronize()
            self.check_staged(res)
        return io


class PipelinedRunner:
    """Keeps `depth` steps in flight: step i runs on lane i % depth (its own stream, engine replica, CUDA
    graph and buffers), so the serial FPS latency chain of one batch overlaps the next batch's.
    Every step still performs the complete hot path on its own batch; results of lane l are valid until
    that lane is used again (depth steps later)."""

    def __init__(self, depth: int = 2, device=None, **engine_kwargs):
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth = depth
        # The lanes share the default communicator: their gathers are issued after each replay, in step order on
        # every rank.  (Capturing a per-lane communicator's all_gather into the lane graphs deadlocked on 2 GPUs:
        # concurrently replaying graphs let the ranks start the lanes' collectives in different orders.)
        self.engines = [SABackboneNMS(**engine_kwargs).to(self.dev) for _ in range(depth)]  # same seed: same weights
        self.lanes = [torch.cuda.Stream(self.dev) for _ in range(depth)]
        self._i = 0
        self._io = [None] * depth

    def _configure(self, frames: int):
        # `depth` batches sample concurrently: once their clouds outnumber the 8-CTA clusters the GPU can host,
        # the one-SM-per-cloud sampler (fps_bucket.cu) is the one that keeps every batch running (identical
        # results; the choice is baked into the captured graphs).  TSMDET_FPS_ALGO still overrides.
        sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        crowded = self.depth * frames * 8 > sms
        _lib.call("tsmdet_fps_configure", 2 if crowded else 0)
        for eng in self.engines:  # that sampler records the chaining facts for free: levels 2/3 become look-ups
            eng.chain_fps = eng.chain_fps or crowded
        return crowded

    def prepare(self, xyz, feats, boxes, scores):
        """Capture every lane's graph; returns each lane's device-resident input buffers."""
        self._configure(xyz.shape[0])
        multi = torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
        ins = []
        for eng, lane in zip(self.engines, self.lanes):
            lane.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(lane):
                res = eng.forward_device(xyz, feats, boxes, scores)
                ins.append(eng.static_inputs(xyz, feats, boxes, scores))
                if multi:
                    eng._peer_gather(res["det_packed"])  # map the peers' receive rings now, not in a timed step
        self.sync()
        return ins

    def prepare_host(self, h_xyz, h_feats, h_boxes, h_scores):
        """Capture every lane's staged graph and fill every lane's pinned HostIO with this batch (collated the way
        the reference's data loader delivers it); returns the HostIO objects."""
        f, n, _ = h_xyz.shape
        self._configure(f)
        multi = torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
        for i, (eng, lane) in enumerate(zip(self.engines, self.lanes)):
            io = eng.host_io(f, n, h_feats.shape[1], h_scores.shape[1]).fill(h_xyz, h_feats, h_boxes, h_scores)
            self._io[i] = io
            lane.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(lane):
                eng.forward_host(io, gather=False, sync=False)
                if multi:
                    eng._peer_gather(eng._graph_for_host(io, self.dev)["out"]["det_packed"])
        self.sync()
        return self._io

    def submit_device(self, inputs, gather: bool = False, pre=None):
        lane_id = self._i % self.depth
        self._i += 1
        lane = self.lanes[lane_id]
        with torch.cuda.stream(lane):
            if pre is not None:
                pre()
            return lane_id, self.engines[lane_id].forward_device(*inputs[lane_id], gather=gather)

    def submit_host(self, gather: bool = False, pre=None):
        """One step from the lane's pinned HostIO (fill ``runner.io(lane)`` beforehand) -> its pinned results
        (asynchronous; call sync() before reading them)."""
        lane_id = self._i % self.depth
        self._i += 1
        lane, eng = self.lanes[lane_id], self.engines[lane_id]
        with torch.cuda.stream(lane):
            if pre is not None:
                pre()
            eng.forward_host(self._io[lane_id], gather=gather, sync=False)
        return lane_id, self._io[lane_id]

    def io(self, lane_id: int) -> HostIO:
        return self._io[lane_id]

    def fork(self):
        """Order every lane after the current stream (call before the first submit of a timed region)."""
        cur = torch.cuda.current_stream(self.dev)
        for lane in self.lanes:
            lane.wait_stream(cur)

    def join(self):
        """Order the current stream after every lane (call before recording the end event)."""
        cur = torch.cuda.current_stream(self.dev)
        for lane in self.lanes:
            cur.wait_stream(lane)

    def sync(self):
        for lane in self.lanes:
            lane.synchronize()

    def check(self):
        """After sync(): raises if any lane's staging kernel saw a malformed collated batch."""
        for eng, io in zip(self.engines, self._io):
            if io is not None:
                eng.check_staged(eng._graph_for_host(io, self.dev)["out"])

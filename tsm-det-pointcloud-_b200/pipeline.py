"""The benchmarked hot path as one callable: SA backbone (BASELINE.json config 2) + rotated NMS
(config 3: 4096 proposals per frame, IoU 0.01 then 0.1 on the survivors), frames sharded per GPU.

Execution model (B200-first: streams + CUDA graphs instead of a tracing compiler):

  * the three FPS launches form a serial latency chain (FPS L2 samples the centres FPS L1 picked),
    but they occupy only one 128-thread CTA per SM, so ball query, the tcgen05 MLP kernels and the NMS
    kernels run BESIDE them on other streams:
        stream A : fps1 -> gather1 -> fps2 -> gather2 -> fps3 -> gather3
        stream B :          [e1] query1 -> mlp1 -> [e2] query2 -> mlp2 -> [e3] query3 -> mlp3
        stream C : sort -> nms(0.01) -> compact -> nms(0.1)
  * the whole DAG is captured once into a CUDA graph and replayed per step (one launch, no Python
    between kernels, no allocator traffic).

``forward_device`` takes device-resident tensors (bench ``value``); ``forward_host`` is the
reference-facing call with HOST buffers: pinned H2D copies in, results copied back (bench ``e2e``).
"""
from __future__ import annotations

import os
import sys
from typing import Dict, Optional

import torch

from . import _lib, iou3d_nms_utils, pointnet2_utils
from .pointnet2_modules import gather_xyz, kitti_sa_stack, sa_mlp_maxpool
from .sharding import PeerGather, gather_detections, gather_packed, pack_detections


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

    @torch.no_grad()
    def _run(self, xyz, feats, boxes, scores):
        """Multi-stream DAG; everything it launches is ordered after / joined back into the current stream."""
        dev = xyz.device
        main = torch.cuda.current_stream(dev)
        s_sa, s_nms = self._side_streams(dev)
        s_sa.wait_stream(main)
        s_nms.wait_stream(main)

        trace = self._trace
        def mark(name, stream):
            if trace is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                trace.append((name, ev))
        mark("start", main)
        layers = self.backbone.layers
        b = xyz.shape[0]
        cur_xyz = xyz
        centres, ready = [], []
        # stream A (current): the FPS chain
        state = None
        for layer in layers:
            if self.chain_fps:
                idx, state = pointnet2_utils.farthest_point_sample_chained(cur_xyz, layer.npoint_list[0], state)
            else:
                idx = pointnet2_utils.farthest_point_sample(cur_xyz, layer.npoint_list[0])
            new_xyz = gather_xyz(cur_xyz, idx)
            mark(f"fps{len(ready) + 1}", main)
            ev = torch.cuda.Event()
            ev.record(main)
            centres.append((cur_xyz, new_xyz))
            ready.append(ev)
            cur_xyz = new_xyz
        # stream B: query + fused MLP per layer, as soon as that layer's centres exist
        with torch.cuda.stream(s_sa):
            cur_f = feats
            for layer, (src_xyz, new_xyz), ev in zip(layers, centres, ready):
                s_sa.wait_event(ev)
                g = layer.groupers[0]
                cnt, bidx = pointnet2_utils.ball_query(g.radius, g.nsample, src_xyz, new_xyz)
                mark(f"query{src_xyz.shape[1]}", s_sa)
                folded = layer._folded_layers()[0]
                out = torch.empty((b, folded[-1][0].shape[0], new_xyz.shape[1]), dtype=torch.float32, device=dev)
                sa_mlp_maxpool(src_xyz, new_xyz, cur_f, bidx, cnt, folded, out, 0, precision=layer.precision)
                mark(f"mlp{src_xyz.shape[1]}", s_sa)
                cur_f = out
        # stream C: NMS is independent of the backbone.  (Measured inside the captured graph: letting it run
        # beside the sampling chain is faster -- 4.98 vs 5.61 ms/step -- than holding it back until the chain
        # is done; TSMDET_NMS_AFTER_FPS=1 selects the latter for experiments.)
        with torch.cuda.stream(s_nms):
            if self.nms_after_fps:
                s_nms.wait_event(ready[-1])
            det_idx, det_num = self._nms_two_pass(boxes, scores)
            mark("nms", s_nms)
            # fixed-size detection records (F, K, 9) = box(7), score, 0 -- what the ranks exchange at N > 1
            f, k = det_idx.shape
            live = det_idx >= 0
            safe_k = torch.where(live, det_idx, torch.zeros_like(det_idx))
            rec = torch.zeros((f, k, 9), dtype=torch.float32, device=dev)
            rec[:, :, :7] = torch.gather(boxes[:, :, :7], 1, safe_k.unsqueeze(-1).expand(-1, -1, 7))
            rec[:, :, 7] = torch.gather(scores, 1, safe_k)
            rec = rec * live.unsqueeze(-1)
        main.wait_stream(s_sa)
        main.wait_stream(s_nms)
        res = {"xyz": cur_xyz, "features": cur_f, "det_idx": det_idx, "det_num": det_num, "det": rec,
               "det_packed": pack_detections(rec, det_num)}  # the collective's send buffer, built inside the graph
        if self.gather_group is not None:
            # the one collective of the path: every rank receives every rank's records (equal shards, so the
            # sizes are known without a size exchange or a host sync)
            world = torch.distributed.get_world_size(self.gather_group)
            res["all_det"], res["all_num"] = gather_detections(rec, det_num, frames_total=rec.shape[0] * world,
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
    def _graph_for(self, xyz, feats, boxes, scores):
        key = (xyz.device.index, tuple(xyz.shape), tuple(feats.shape), tuple(boxes.shape), tuple(scores.shape))
        ent = self._graphs.get(key)
        if ent is not None:
            return ent
        dev = xyz.device
        static_in = [torch.empty_like(t) for t in (xyz, feats, boxes, scores)]
        for s, t in zip(static_in, (xyz, feats, boxes, scores)):
            s.copy_(t)
        # the sampling chain is captured on the capture stream itself: TSMDET_FPS_PRIORITY=1 makes it a high-priority
        # stream (its kernels' graph nodes inherit the priority), so FPS CTAs are placed first when SMs free up
        cap = torch.cuda.Stream(dev, priority=-1 if os.environ.get("TSMDET_FPS_PRIORITY", "0") == "1" else 0)
        cap.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap):
            for _ in range(2):  # warm-up on the capture stream: grows every scratch buffer, folds BN, plans FPS
                self._run(*static_in)
        cap.synchronize()
        graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count
        with torch.cuda.graph(graph, stream=cap):
            out = self._run(*static_in)
        torch.cuda.current_stream(dev).wait_stream(cap)
        # kernels of this library inside one replay (the graph also holds a few torch sort/gather kernels)
        ent = {"graph": graph, "in": static_in, "out": out, "launches": _lib.launch_count - l0}
        self._graphs[key] = ent
        return ent

    @torch.no_grad()
    def forward_device(self, xyz, feats, boxes, scores, gather: bool = False) -> Dict[str, torch.Tensor]:
        """xyz (F,N,3), feats (F,C,N), boxes (F,P,7), scores (F,P) on the GPU.  The returned tensors are
        owned by the engine and overwritten by the next call with the same shapes (graph replay)."""
        if self.use_graph:
            ent = self._graph_for(xyz, feats, boxes, scores)
            for s, t in zip(ent["in"], (xyz, feats, boxes, scores)):
                if s.data_ptr() != t.data_ptr():
                    s.copy_(t, non_blocking=True)
            ent["graph"].replay()
            _lib.launch_count += ent["launches"]
            res = dict(ent["out"])
            boxes, scores = ent["in"][2], ent["in"][3]
        else:
            res = self._run(xyz, feats, boxes, scores)
        if gather and "all_det" not in res:
            # the one collective of the path, issued after the replay: equal shards, so a single all_gather of the
            # packed buffer the graph produced -- no size exchange, no host sync, no other kernels
            if torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
                f, k = res["det"].shape[0], res["det"].shape[1]
                pg = self._peer_gather(res["det_packed"])
                if pg is not None:  # one-sided copies into every rank's receive buffer (sharding.PeerGather)
                    pg.put(res["det_packed"])
                    res["all_det"], res["all_num"] = pg.views(f, k)
                else:
                    res["all_det"], res["all_num"], self._gather_out = gather_packed(
                        res["det_packed"], f, k, out=getattr(self, "_gather_out", None))
            else:
                res["all_det"], res["all_num"] = res["det"].unsqueeze(0), res["det_num"].unsqueeze(0)
        return res

    def _peer_gather(self, packed):
        """The transport of the detection gather: peer-memory copies (TSMDET_GATHER=peer, the default on GPUs) or the
        NCCL all_gather (TSMDET_GATHER=nccl, and whenever the IPC set-up fails -- every rank takes the same branch)."""
        if not hasattr(self, "_pg"):
            self._pg = None
            want = os.environ.get("TSMDET_GATHER", "peer") == "peer" and packed.is_cuda
            ok = torch.zeros((1,), dtype=torch.int32, device=packed.device)
            if want:
                try:
                    self._pg = PeerGather(packed.numel(), packed.device)
                    ok += 1
                except Exception as e:  # noqa: BLE001 -- fall back to NCCL together with everybody else
                    print(f"[tsmdet] peer-memory gather unavailable ({e}); using the NCCL all_gather", file=sys.stderr)
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
            if int(ok.item()) == 0:
                self._pg = None
        return self._pg

    def static_inputs(self, xyz, feats, boxes, scores):
        """The engine-owned input buffers for these shapes (write into them to skip the D2D copy)."""
        return self._graph_for(xyz, feats, boxes, scores)["in"]

    @torch.no_grad()
    def forward_host(self, h_xyz, h_feats, h_boxes, h_scores, h_out: Optional[dict] = None, gather: bool = False):
        """Pinned host tensors in, pinned host results out (returns after the copies finish)."""
        dev = next(self.parameters()).device
        if self.use_graph:
            key = (dev.index, *[tuple(t.shape) for t in (h_xyz, h_feats, h_boxes, h_scores)])
            ent = self._graphs.get(key)
            if ent is None:
                ent = self._graph_for(*[t.to(dev) for t in (h_xyz, h_feats, h_boxes, h_scores)])
            d_in = ent["in"]
            for s, t in zip(d_in, (h_xyz, h_feats, h_boxes, h_scores)):
                s.copy_(t, non_blocking=True)
        else:
            d_in = [t.to(dev, non_blocking=True) for t in (h_xyz, h_feats, h_boxes, h_scores)]
        res = self.forward_device(*d_in, gather=gather)
        if h_out is None:
            h_out = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in res.items()}
        for k, v in res.items():
            h_out[k].copy_(v, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return h_out


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
        self._h_out = [None] * depth

    def prepare(self, xyz, feats, boxes, scores):
        """Capture every lane's graph; returns each lane's device-resident input buffers."""
        # `depth` batches sample concurrently: once their clouds outnumber the 8-CTA clusters the GPU can host,
        # the one-SM-per-cloud sampler (fps_bucket.cu) is the one that keeps every batch running (identical
        # results; the choice is baked into the captured graphs).  TSMDET_FPS_ALGO still overrides.
        sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        crowded = self.depth * xyz.shape[0] * 8 > sms
        _lib.call("tsmdet_fps_configure", 2 if crowded else 0)
        for eng in self.engines:  # that sampler records the chaining facts for free: levels 2/3 become look-ups
            eng.chain_fps = eng.chain_fps or crowded
        ins = []
        for eng, lane in zip(self.engines, self.lanes):
            lane.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(lane):
                res = eng.forward_device(xyz, feats, boxes, scores)
                ins.append(eng.static_inputs(xyz, feats, boxes, scores))
                if torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
                    eng._peer_gather(res["det_packed"])  # map the peers' receive buffers now, not in a timed step
        self.sync()
        return ins

    def submit_device(self, inputs, gather: bool = False, pre=None):
        lane_id = self._i % self.depth
        self._i += 1
        lane = self.lanes[lane_id]
        with torch.cuda.stream(lane):
            if pre is not None:
                pre()
            return lane_id, self.engines[lane_id].forward_device(*inputs[lane_id], gather=gather)

    def submit_host(self, h_in, gather: bool = False, pre=None):
        """Pinned host tensors in -> pinned host results (asynchronous; call sync() before reading them)."""
        lane_id = self._i % self.depth
        self._i += 1
        lane, eng = self.lanes[lane_id], self.engines[lane_id]
        with torch.cuda.stream(lane):
            if pre is not None:
                pre()
            ent = eng._graphs.get((self.dev.index, *[tuple(t.shape) for t in h_in]))
            d_in = ent["in"]
            for s, t in zip(d_in, h_in):
                s.copy_(t, non_blocking=True)
            res = eng.forward_device(*d_in, gather=gather)
            if gather and getattr(eng, "_pg", None) is not None:
                eng._pg.wait_stream()  # the copies below read the gathered records: wait for the peers' stores
            if self._h_out[lane_id] is None:
                self._h_out[lane_id] = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in res.items()}
            for k, v in res.items():
                self._h_out[lane_id][k].copy_(v, non_blocking=True)
        return lane_id, self._h_out[lane_id]

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

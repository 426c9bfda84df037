#!/usr/bin/env python
"""bench.py -- frames/s of the SA backbone + rotated NMS hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 50 --warmup 8                  # this repo's sm_100a path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                             # reference CPU arm (oracle port)

A "step" = one pass of the hot path over one batch of synthetic frames per GPU:
  config 2  KITTI SA stack 16384 -> 4096 -> 1024 -> 512 (radii 0.2/0.8/1.6, nsample 16/32/32,
            MLPs [4,16,16,32] [35,64,64,128] [131,128,128,256]), batch 16 frames per GPU, plus
  config 3  rotated NMS on 4096 proposals per frame, IoU 0.01 then 0.1 on the survivors; at N > 1 the
            padded detections are gathered on every rank (frames are sharded, weak scaling): one kernel of
            peer-memory stores per step by default, the NCCL all_gather with TSMDET_GATHER=nccl.
Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same
through the host-buffer API (pinned H2D of the inputs + D2H of the results inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FRAMES_PER_GPU = 16
N_POINTS = 16384
N_PROPOSALS = 4096
METRIC = "frames/sec (SA backbone+NMS, 16384 pts/frame)"
WORKLOAD = ("KITTI SA stack 16384->4096->1024->512 (r 0.2/0.8/1.6, ns 16/32/32) batch 16/GPU + rotated NMS "
            "4096 proposals/frame IoU 0.01 then 0.1")


# ------------------------------------------------------------------------------------ inputs
def make_inputs(frames: int, seed: int):
    import numpy as np
    import synth

    xyz = np.concatenate([synth.cloud_ground_objects(frames // 2, N_POINTS, seed),
                          synth.cloud_dup_padded(frames - frames // 2, N_POINTS, seed + 1)], 0)
    feats = np.random.default_rng(seed + 2).uniform(0, 1, size=(frames, 1, N_POINTS)).astype(np.float32)
    boxes = np.stack([synth.boxes_clustered(N_PROPOSALS, seed + 10 + i, centres=200) for i in range(frames)])
    scores = np.stack([synth.scores_random(N_PROPOSALS, seed + 100 + i) for i in range(frames)])
    return xyz, feats, boxes, scores


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an in-process NVML thread (about 250 Hz;
    only the samples whose timestamps fall inside [start(), stop()] are kept), nvidia-smi -lms as the fallback."""

    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, device):
        self.device, self.rows, self.proc, self.nvml, self.handle = device, [], None, None, None
        self._stop = threading.Event()
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            props = torch.cuda.get_device_properties(device)
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(props.uuid)).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(device.index or 0)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _loop(self):
        nv, h = self.nvml, self.handle
        while not self._stop.is_set():
            try:
                self.rows.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                  int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
            except Exception:
                break
            time.sleep(0.004)  # ~250 Hz: NVML calls hold the GIL, the timed loop needs it

    def start(self):
        self.t0 = time.perf_counter()
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.device.index or 0)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        t1 = time.perf_counter()
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=1.0)
            nv = self.nvml
            inside = [r for r in self.rows if self.t0 <= r[0] <= t1]
            sm = sorted(r[1] for r in inside)
            bits = 0
            for r in inside:
                bits |= r[2]
            try:
                mx = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            except Exception:
                mx = None
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                    "reasons": sorted(n for n, b in self.REASONS if bits & b), "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_frames_per_second(frames: int, threads: int, seed: int = 0):
    """The oracle port (this repo's C restatement of the reference kernels + torch-CPU fp32 MLP) on the
    host cores, over `frames` frames of the same workload.  Returns (frames/s, seconds)."""
    import numpy as np
    import torch

    from oracle import oracle as orc

    orc.build()
    orc.set_threads(threads)
    torch.set_num_threads(threads)
    from tsmdet_b200.pointnet2_modules import kitti_sa_stack  # module definitions only; no CUDA call below

    xyz, feats, boxes, scores = make_inputs(max(frames, 2), seed)
    xyz, feats, boxes, scores = xyz[:frames], feats[:frames], boxes[:frames], scores[:frames]
    torch.manual_seed(0)
    net = kitti_sa_stack(fused=False).eval()
    t0 = time.perf_counter()
    cur_xyz, cur_f = xyz, feats
    with torch.no_grad():
        for layer in net.layers:
            npoint = layer.npoint_list[0]
            idx = orc.fps(cur_xyz, npoint)
            new_xyz = np.take_along_axis(cur_xyz, idx.astype(np.int64)[..., None], axis=1)
            g = layer.groupers[0]
            cnt, nf, _, _ = orc.query_and_group(cur_xyz, new_xyz, cur_f, g.radius, g.nsample)
            x = torch.from_numpy(nf) * torch.from_numpy((cnt > 0).astype(np.float32))[:, None, :, None]
            y = layer.point_mlps[0](x).max(dim=3)[0]
            cur_xyz, cur_f = new_xyz, y.numpy()
    for i in range(frames):
        order = np.argsort(-scores[i], kind="stable")
        k1 = orc.nms_sorted(boxes[i][order], 0.01)
        orc.nms_sorted(boxes[i][order][k1], 0.1)
    dt = time.perf_counter() - t0
    return frames / dt, dt


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (the pointnet2 ops are CUDA-only in
    the reference, so this is the oracle port of those kernels + its rotated-IoU CPU code), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames_per_step = FRAMES_PER_GPU  # one full per-GPU batch, so the OpenMP loops over clouds use every core
    for _ in range(args.warmup and 1):
        cpu_frames_per_second(frames_per_step, cores)
    vals = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        fps, _ = cpu_frames_per_second(frames_per_step, cores)
        vals.append(fps)
        if time.perf_counter() - t_all > 240:
            break
    value = len(vals) * frames_per_step / sum(frames_per_step / v for v in vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 * frames_per_step / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": frames_per_step},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{frames_per_step} frame(s) of the same workload per step, {len(vals)} steps, "
                                   f"OpenMP C oracle + torch-CPU fp32 MLP on {cores} threads"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("TSMDET_BENCH_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--depth", type=int, default=int(os.environ.get("TSMDET_BENCH_DEPTH", "8")),
                    help="steps kept in flight (each on its own stream / CUDA graph / buffers)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-kernels", action="store_true", help="per-kernel CUDA-event breakdown to stderr")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from tsmdet_b200 import _lib
    from tsmdet_b200.pipeline import PipelinedRunner

    depth = max(1, args.depth)
    xyz_np, feats_np, boxes_np, scores_np = make_inputs(FRAMES_PER_GPU, seed=1000 * rank)
    h = [torch.from_numpy(a).pin_memory() for a in (xyz_np, feats_np, boxes_np, scores_np)]
    d0 = [t.to(dev) for t in h]
    try:
        runner = PipelinedRunner(depth=depth, device=dev, precision=args.precision)
        lane_inputs = runner.prepare(*d0)  # captures one CUDA graph per lane; inputs stay resident in HBM
    except _lib.TsmdetError as e:
        if args.precision == "bf16" and e.code == 1000001:
            args.precision = "fp32"
            runner = PipelinedRunner(depth=depth, device=dev, precision="fp32")
            lane_inputs = runner.prepare(*d0)
        else:
            raise
    engine = runner.engines[0]
    gather = world > 1 and os.environ.get("TSMDET_BENCH_NO_GATHER", "0") == "0"  # (experiments only)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def flush():
        flush_buf.zero_()  # L2 flush before every step, inside the timed region (on the step's own stream)

    def timed(submit, steps):
        """K steps, `depth` in flight; device time between one start and one end event on the main stream."""
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        runner.fork()
        for _ in range(steps):
            submit()
        runner.join()
        e.record()
        torch.cuda.synchronize(dev)
        return s.elapsed_time(e)

    dev_step = lambda: runner.submit_device(lane_inputs, gather=gather, pre=flush)  # noqa: E731
    host_step = lambda: runner.submit_host(h, gather=gather, pre=flush)  # noqa: E731

    for _ in range(depth):  # prime every lane once (graph upload, first collective) whatever --warmup is
        dev_step()
    runner.sync()
    for _ in range(args.warmup):
        dev_step()
    runner.sync()
    barrier()
    gather_transport = None
    if gather:
        # outside the timed region: whatever transport the gather uses, every rank must now hold exactly what an NCCL
        # all_gather of the last step's packed records returns
        from tsmdet_b200.sharding import gather_packed

        for eng in runner.engines[:2]:
            res = eng.forward_device(*lane_inputs[0], gather=True)
            torch.cuda.synchronize(dev)
            pg = getattr(eng, "_pg", None)
            gather_transport = "peer-memory stores (tsmdet_peer_put), one kernel per step" if pg is not None else "nccl all_gather"
            if pg is not None:
                pg.wait()
            f, k = res["det"].shape[0], res["det"].shape[1]
            ref_det, ref_num, _ = gather_packed(res["det_packed"], f, k)
            torch.cuda.synchronize(dev)
            if not (torch.equal(res["all_det"], ref_det) and torch.equal(res["all_num"], ref_num)):
                raise RuntimeError("detection gather differs from the NCCL all_gather of the same records")
        barrier()
    sampler = ClockSampler(dev)
    sampler.start()
    l0 = _lib.launch_count
    t_wall = time.perf_counter()
    ms_total = timed(dev_step, args.steps)
    barrier()
    wall = time.perf_counter() - t_wall
    launches = _lib.launch_count - l0
    clocks = sampler.stop()
    for _ in range(depth + args.warmup):
        host_step()
    runner.sync()
    barrier()
    ms_e2e_total = timed(host_step, args.steps)
    barrier()
    _, h_out = runner.submit_host(h, gather=gather)
    runner.sync()

    # single-step latency (one step in flight, no flush inside the events)
    lat = []
    for _ in range(5):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        engine.forward_device(*lane_inputs[0])
        e.record()
        torch.cuda.synchronize(dev)
        lat.append(s.elapsed_time(e))
    lat.sort()

    tot = torch.tensor([ms_total, ms_e2e_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)  # max over ranks
    t_dev, t_e2e = float(tot[0]) / 1000.0, float(tot[1]) / 1000.0
    frames_total = FRAMES_PER_GPU * world * args.steps
    value = frames_total / t_dev
    e2e = frames_total / t_e2e
    d = lane_inputs[0]

    # ---- per-kernel breakdown + roofline of the dominant kernel (CUDA events on the launching stream)
    roof, kernels = kernel_breakdown(engine, d, dev, args)

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * t_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu": FRAMES_PER_GPU, "points_per_frame": N_POINTS,
                   "proposals_per_frame": N_PROPOSALS, "parallelism": f"frames sharded x{world}",
                   "mlp_precision": args.precision,
                   "l2": "256 MB buffer written before every step, inside the timed region",
                   "execution": "one CUDA graph per step (FPS chain, query+MLP, NMS on 3 concurrent streams); "
                                f"{depth} step(s) in flight on separate streams/buffers",
                   "pipeline_depth": depth, "ms_per_step_single_in_flight": lat[len(lat) // 2],
                   "gather": gather_transport},
        "e2e": {"value": e2e, "unit": "frames/s",
                "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in h)),
                "d2h_bytes_per_step": int(sum(t.numel() * t.element_size() for t in h_out.values()))},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": kernels,
        "wall_s_timed_region": wall,
    }
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        n_cpu = 32  # two per-GPU batches: about 10 s of host work, every core busy in the OpenMP loops over clouds
        v, dt = cpu_frames_per_second(n_cpu, cores)
        line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"{n_cpu} frames of the same workload ({dt:.1f} s): OpenMP C oracle of the reference "
                                          f"kernels + torch-CPU fp32 MLP + oracle rotated NMS"}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernel_breakdown(engine, d, dev, args):
    """Times each stage of one step on its own (CUDA events, warm, 5 repeats) and derives the roofline
    entry for the kernel with the largest share (algorithmic bytes: SURVEY.md 8d / DESIGN.md)."""
    import torch

    from tsmdet_b200 import iou3d_nms_utils, pointnet2_utils
    from tsmdet_b200.pointnet2_modules import gather_xyz, sa_mlp_maxpool

    peaks = {"hbm_gbs": 6650.0, "src": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = dict(json.load(f), src="measured")
    except OSError:
        pass

    xyz, feats, boxes, scores = d
    b = xyz.shape[0]
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def t(fn, reps=5):
        """Device time of one stage: captured into a CUDA graph (as the timed step runs it -- no per-launch host
        cost between its kernels), replayed `reps` times between two events on the launching stream."""
        fn()
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            fn()  # warm on the capture stream: grows that stream's scratch buffers
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                fn()
            graph.replay()
            total = 0.0
            for _ in range(reps):
                flush_buf.zero_()  # cold L2 for every replay, as in the timed step
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(side)
                graph.replay()
                e.record(side)
                e.synchronize()
                total += s.elapsed_time(e)
        torch.cuda.synchronize(dev)
        return total / reps

    kernels = []
    kernels_ctx = {}
    cur_xyz, cur_f = xyz, feats
    with torch.no_grad():
        for li, layer in enumerate(engine.backbone.layers):
            n = cur_xyz.shape[1]
            m = layer.npoint_list[0]
            g = layer.groupers[0]
            ns = g.nsample
            c = cur_f.shape[1]
            idx = pointnet2_utils.farthest_point_sample(cur_xyz, m)
            ms_fps = t(lambda: pointnet2_utils.farthest_point_sample(cur_xyz, m))
            new_xyz = gather_xyz(cur_xyz, idx)
            cnt, bidx = pointnet2_utils.ball_query(g.radius, ns, cur_xyz, new_xyz)
            ms_bq = t(lambda: pointnet2_utils.ball_query(g.radius, ns, cur_xyz, new_xyz))
            layers = layer._folded_layers()[0]
            out = torch.empty((b, layers[-1][0].shape[0], m), device=dev)
            ms_mlp = t(lambda: sa_mlp_maxpool(cur_xyz, new_xyz, cur_f, bidx, cnt, layers, out, 0, precision=layer.precision))
            fps_bytes = b * (12 * n + 4 * m)
            bq_bytes = b * (12 * n + 12 * m + 4 * m * ns + 4 * m)
            chans = [3 + c] + [w.shape[0] for w, _ in layers]
            flops = b * 2 * m * ns * sum(chans[i] * chans[i + 1] for i in range(len(chans) - 1))
            sa_bytes = b * (12 * n + 4 * c * n + 12 * m + 4 * chans[-1] * m + 4 * m * ns)
            kernels += [
                {"name": f"fps_L{li + 1}", "ms": ms_fps, "alg_bytes": fps_bytes, "gbs": fps_bytes / ms_fps / 1e6,
                 "us_per_iter": 1000.0 * ms_fps / max(m - 1, 1), "point_updates_per_s": b * n * (m - 1) / ms_fps * 1e3},
                {"name": f"ball_query_L{li + 1}", "ms": ms_bq, "alg_bytes": bq_bytes, "gbs": bq_bytes / ms_bq / 1e6,
                 "tests_per_s": b * n * m / ms_bq * 1e3},
                {"name": f"sa_mlp_maxpool_L{li + 1}", "ms": ms_mlp, "alg_bytes": sa_bytes, "gbs": sa_bytes / ms_mlp / 1e6,
                 "tflops": flops / ms_mlp / 1e9},
            ]
            kernels_ctx = {"xyz": cur_xyz, "f": cur_f, "idx": bidx}
            cur_xyz, cur_f = new_xyz, out
        # the API-level (materialising) grouping op on the last layer's shape: the HBM-bound kernel of the path
        lay = engine.backbone.layers[-1]
        g3 = lay.groupers[0]
        src_xyz, src_f = kernels_ctx["xyz"], kernels_ctx["f"]
        c3, n3, m3, s3 = src_f.shape[1], src_xyz.shape[1], lay.npoint_list[0], g3.nsample
        ms_grp = t(lambda: pointnet2_utils.grouping_operation(src_f, kernels_ctx["idx"]))
        grp_bytes = b * (4 * c3 * n3 + 4 * m3 * s3 + 4 * c3 * m3 * s3)
        kernels.append({"name": f"grouping_operation_L{len(engine.backbone.layers)} (API op, not on the fused path)",
                        "ms": ms_grp, "alg_bytes": grp_bytes, "gbs": grp_bytes / ms_grp / 1e6,
                        "hbm_frac": grp_bytes / ms_grp / 1e6 / float(peaks.get("hbm_gbs", 6650.0))})
        ms_nms = t(lambda: iou3d_nms_utils.nms_gpu_batch(boxes, scores, 0.01))
        p = boxes.shape[1]
        kernels.append({"name": "nms_batch(0.01)", "ms": ms_nms, "alg_bytes": b * 36 * p, "gbs": b * 36 * p / ms_nms / 1e6,
                        "pairs_per_s": b * p * (p - 1) / 2 / ms_nms * 1e3})
    top = max(kernels, key=lambda k: k["ms"])
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None  # DRAM bytes per launch of that kernel from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.jsonl")) as f:
            for ln in f:
                rec = json.loads(ln)
                if top["name"].startswith("fps") and "fps_bucket_kernel" in rec["kernel"]:
                    traffic = rec["dram_bytes"]
    except OSError:
        pass
    roof = {"kernel": top["name"], "bound": "hbm", "achieved": top["gbs"], "peak": peak, "unit": "GB/s",
            "frac": top["gbs"] / peak, "traffic": traffic, "peak_source": peaks["src"],
            "note": "FPS is a serial-latency chain (one argmax per selected point); its HBM traffic is the "
                    "compulsory 12N+4M bytes per cloud, so the HBM fraction is tiny by construction -- see us_per_iter"}
    if args.profile_kernels:
        for k in kernels:
            print(json.dumps(k), file=sys.stderr)
    return roof, kernels


if __name__ == "__main__":
    main()

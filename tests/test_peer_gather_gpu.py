"""GPU suite (needs two GPUs; skipped on a one-GPU box): the peer-memory detection gather (sharding.PeerGather ->
tsmdet_peer_put / tsmdet_peer_wait) against the NCCL all_gather of the same packed records.

The ranks are deliberately SKEWED (one of them is delayed on the device by ``torch.cuda._sleep`` every step, the other
one every third step), the data changes every step, and there is NO barrier between steps: the credit-based flow
control of csrc/peer_put.cu alone has to keep a fast rank from overwriting records a slow rank has not read yet
(the single-buffer version passed only because the test re-synchronised the ranks after every step)."""
import multiprocessing as mp
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _records(rank, step, frames, k, dev):
    g = torch.Generator().manual_seed(1000 * rank + step)
    rec = torch.rand((frames, k, 9), generator=g).to(dev)
    cnt = torch.randint(0, k + 1, (frames,), generator=g, dtype=torch.int32).to(dev)
    return rec, cnt


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist

    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    ok, why = True, ""
    try:
        from tsmdet_b200 import _lib
        from tsmdet_b200.sharding import PeerGather, pack_detections

        steps = 24
        for frames, k, slots in ((3, 5, 2), (16, 512, 2), (16, 512, 3)):  # 3*5*9+3 = 138 floats: scalar tail, padded rows
            pg, err = PeerGather.create(frames * k * 9 + frames, dev, slots=slots)
            assert pg is not None, err
            got = []
            for step in range(1, steps + 1):
                # skew: rank 0 is slow every step, rank 1 every third step (~5 ms of device time each)
                if rank == 0 or step % 3 == 0:
                    torch.cuda._sleep(10_000_000)
                rec, cnt = _records(rank, step, frames, k, dev)
                pg.put(pack_detections(rec, cnt))
                pg.wait_stream()
                det, num = pg.views(frames, k)
                got.append((det.clone(), num.clone()))  # the read is stream-ordered before the next put (= its ack)
            torch.cuda.synchronize(dev)
            for step, (det, num) in enumerate(got, start=1):
                for r in range(world):
                    rec, cnt = _records(r, step, frames, k, dev)
                    if not (torch.equal(det[r], rec) and torch.equal(num[r], cnt)):
                        ok, why = False, f"frames {frames} slots {slots} step {step}: rank {r}'s records torn or stale"
            ok = ok and bool((pg.flags == steps).all())
            if _lib.read_status() != 0:
                ok, why = False, "watchdog status set"
            dist.barrier()  # before the ring of this shape is torn down
        q.put((rank, ok, why))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_gather_skewed_ranks_no_barrier():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), [w for _, _, w in res]


def _worker_engine(rank, world, port, q):
    """forward_device(gather=True) hands out the gathered views already ordered after the peers' stores (ADVICE r1:
    it used to return live views of the receive buffer without waiting)."""
    import sys

    import numpy as np

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import synth
    import torch.distributed as dist

    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    ok = True
    try:
        from tsmdet_b200.pipeline import SABackboneNMS
        from tsmdet_b200.sharding import gather_packed

        eng = SABackboneNMS(precision="bf16").to(dev)
        f = 2
        for step in range(4):
            seed = 100 * rank + step
            xyz = torch.from_numpy(synth.cloud_ground_objects(f, 16384, seed)).to(dev)
            feats = torch.rand((f, 1, 16384), generator=torch.Generator().manual_seed(seed)).to(dev)
            boxes = torch.from_numpy(np.stack([synth.boxes_clustered(1024, seed + i, centres=80) for i in range(f)])).to(dev)
            scores = torch.from_numpy(np.stack([synth.scores_random(1024, seed + 10 + i) for i in range(f)])).to(dev)
            if rank == 1:
                torch.cuda._sleep(20_000_000)
            res = eng.forward_device(xyz, feats, boxes, scores, gather=True)
            all_det, all_num = res["all_det"].clone(), res["all_num"].clone()
            ref_det, ref_num, _ = gather_packed(res["det_packed"].clone(), f, res["det"].shape[1])
            torch.cuda.synchronize(dev)
            ok = ok and torch.equal(all_det, ref_det) and torch.equal(all_num, ref_num)
            ok = ok and int(all_num.sum()) > 0
        q.put((rank, ok, ""))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_forward_device_gather_waits_for_peers():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_engine, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)

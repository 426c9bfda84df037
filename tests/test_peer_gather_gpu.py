"""GPU suite (needs two GPUs; skipped on a one-GPU box): the peer-memory detection gather (sharding.PeerGather ->
tsmdet_peer_put / tsmdet_peer_wait) against the NCCL all_gather of the same packed records, several steps, odd sizes."""
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


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist

    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    ok = True
    try:
        from tsmdet_b200.sharding import PeerGather, gather_packed, pack_detections

        for frames, k in ((3, 5), (16, 512)):  # 3*5*9+3 = 138 floats: not a multiple of 4 (scalar tail, padded rows)
            pg = PeerGather(frames * k * 9 + frames, dev)
            for step in range(1, 4):
                g = torch.Generator().manual_seed(100 * rank + step)
                rec = torch.rand((frames, k, 9), generator=g).to(dev)
                cnt = torch.randint(0, k + 1, (frames,), generator=g, dtype=torch.int32).to(dev)
                packed = pack_detections(rec, cnt)
                pg.put(packed)
                pg.wait_stream()
                torch.cuda.synchronize(dev)
                pg.wait()
                det, num = pg.views(frames, k)
                ref_det, ref_num, _ = gather_packed(packed, frames, k)
                torch.cuda.synchronize(dev)
                ok = ok and torch.equal(det, ref_det) and torch.equal(num, ref_num)
                ok = ok and bool((pg.flags == step).all())
                dist.barrier()  # nobody starts the next put while a peer still compares this one
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2 or os.environ.get("TSMDET_TEST_MULTI_GPU", "0") != "1",
                    reason="needs two GPUs and TSMDET_TEST_MULTI_GPU=1 (bench.py checks the same equality at N > 1)")
def test_peer_gather_equals_nccl_all_gather():
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
    assert all(ok for _, ok in res)

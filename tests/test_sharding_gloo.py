"""CPU suite: the N>1 path (frame sharding + the detection all_gather) on gloo, world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_frames, k_post, q):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tsmdet_b200.sharding import gather_detections, pad_detections, shard_bounds

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(total_frames, world, rank)
        boxes, scores, labels = [], [], []
        for f in range(lo, hi):
            g = torch.Generator().manual_seed(f)
            n = (f * 7) % (k_post + 5)  # some frames overflow k_post, some are empty
            boxes.append(torch.rand((n, 7), generator=g) + f)
            scores.append(torch.rand((n,), generator=g))
            labels.append(torch.randint(1, 4, (n,), generator=g))
        padded, cnt = pad_detections(boxes, scores, labels, k_post, device=torch.device("cpu"))
        allp, allc = gather_detections(padded, cnt, frames_total=total_frames)
        allp2, allc2 = gather_detections(padded, cnt)  # sizes discovered with an extra tiny collective
        ok = torch.equal(allp, allp2) and torch.equal(allc, allc2)
        q.put((rank, ok, allp.numpy().copy(), allc.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_gather_detections_world2():
    total, k_post, world = 5, 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, k_post, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    assert all(r[1] for r in res)
    a, b = res[0], res[1]
    assert (a[2] == b[2]).all() and (a[3] == b[3]).all()  # every rank holds the full result
    allp, allc = torch.from_numpy(a[2]), torch.from_numpy(a[3])
    assert allp.shape == (total, k_post, 9) and allc.shape == (total,)
    for f in range(total):
        g = torch.Generator().manual_seed(f)
        n = (f * 7) % (k_post + 5)
        bx = torch.rand((n, 7), generator=g) + f
        k = min(n, k_post)
        assert int(allc[f]) == k
        assert torch.equal(allp[f, :k, :7], bx[:k])
        assert float(allp[f, k:].abs().sum()) == 0.0


def _worker_packed(rank, world, port, frames, k_post, q):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tsmdet_b200.sharding import gather_packed, pack_detections

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        rec = torch.rand((frames, k_post, 9), generator=g) + rank
        cnt = torch.randint(0, k_post + 1, (frames,), generator=g, dtype=torch.int32)
        out = None
        for _ in range(2):  # second call reuses the receive buffer
            all_det, all_num, out = gather_packed(pack_detections(rec, cnt), frames, k_post, out=out)
        q.put((rank, rec.numpy().copy(), cnt.numpy().copy(), all_det.numpy().copy(), all_num.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_gather_packed_world2():
    """The pipeline's collective (equal shards, one all_gather of the packed records + counts)."""
    frames, k_post, world = 3, 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_packed, args=(r, world, port, frames, k_post, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    for r in res:
        assert r[3].shape == (world, frames, k_post, 9) and r[4].shape == (world, frames)
        for src in res:
            assert (r[3][src[0]] == src[1]).all() and (r[4][src[0]] == src[2]).all()

"""Frame sharding across GPUs and the one collective of this path: gathering detections.

Frames are independent (every op here is per cloud; NMS is per frame), so a batch is split by
frame across ranks exactly like the reference's DistributedSampler splits the dataset
(``/root/reference/pcdet/datasets/__init__.py:24-44``), with no collective inside the SA stack or NMS.
The reference merges results through pickle files on a shared tmpdir plus two barriers
(``pcdet/utils/common_utils.py:224-245``); here the per-frame detections are padded to a fixed
``(K, 9)`` record [x,y,z,dx,dy,dz,heading,score,label] and exchanged with ONE all_gather (NCCL over
NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_frames: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of frames owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(num_frames, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_frames(x: torch.Tensor, world_size: int, rank: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], world_size, rank)
    return x[lo:hi]


def pad_detections(boxes: Sequence[torch.Tensor], scores: Sequence[torch.Tensor], labels: Sequence[torch.Tensor],
                   k_post: int, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-frame variable-length detections -> (F, k_post, 9) float32 zero-padded + (F,) int32 counts."""
    f = len(boxes)
    device = device if device is not None else (boxes[0].device if f else torch.device("cpu"))
    out = torch.zeros((f, k_post, 9), dtype=torch.float32, device=device)
    cnt = torch.zeros((f,), dtype=torch.int32, device=device)
    for i in range(f):
        k = min(int(boxes[i].shape[0]), k_post)
        if k:
            out[i, :k, :7] = boxes[i][:k, :7]
            out[i, :k, 7] = scores[i][:k]
            out[i, :k, 8] = labels[i][:k].to(torch.float32)
        cnt[i] = k
    return out, cnt


def gather_detections(padded: torch.Tensor, counts: torch.Tensor, frames_total: Optional[int] = None,
                      group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All ranks receive every rank's (F_local, K, 9) records and counts, concatenated in rank order
    (= original frame order for ``shard_bounds`` sharding).  Ranks may own different numbers of
    frames; shards are padded to the largest one for the fixed-size collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return padded, counts
    world = dist.get_world_size(group)
    k = padded.shape[1]
    if frames_total is None:  # unknown shard sizes: one small exchange (synchronises the host)
        f_local = torch.tensor([padded.shape[0]], dtype=torch.int64, device=padded.device)
        f_all = [torch.zeros_like(f_local) for _ in range(world)]
        dist.all_gather(f_all, f_local, group=group)
        sizes = [int(t.item()) for t in f_all]
    else:
        sizes = [shard_bounds(frames_total, world, r)[1] - shard_bounds(frames_total, world, r)[0] for r in range(world)]
    fmax = max(sizes)
    # one flat buffer per rank: records then counts (as float32 bit patterns) -> a single collective
    buf = torch.zeros((fmax * k * 9 + fmax,), dtype=torch.float32, device=padded.device)
    buf[: padded.numel()] = padded.reshape(-1)
    buf[fmax * k * 9: fmax * k * 9 + counts.numel()] = counts.to(torch.int32).view(torch.float32)
    out = torch.empty((world, buf.numel()), dtype=torch.float32, device=padded.device)
    if hasattr(dist, "all_gather_into_tensor") and padded.is_cuda:
        dist.all_gather_into_tensor(out, buf, group=group)
    else:
        dist.all_gather(list(out.unbind(0)), buf, group=group)
    recs, cnts = [], []
    for r in range(world):
        recs.append(out[r, : sizes[r] * k * 9].view(sizes[r], k, 9))
        cnts.append(out[r, fmax * k * 9: fmax * k * 9 + sizes[r]].contiguous().view(torch.int32))
    return torch.cat(recs, 0), torch.cat(cnts, 0)


def pack_detections(padded: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """(F,K,9) records + (F,) int32 counts -> one flat float32 buffer (records, then the counts' bit patterns): the
    send buffer of ``gather_packed``.  Pure tensor ops, so it can live inside a captured CUDA graph."""
    return torch.cat([padded.reshape(-1), counts.to(torch.int32).view(torch.float32)])


def gather_packed(packed: torch.Tensor, frames: int, k: int, out: Optional[torch.Tensor] = None, group=None):
    """Equal shards (every rank owns ``frames`` frames): ONE collective and no other device or host work --
    returns ``(all_det (world, frames, k, 9), all_num (world, frames) int32, out)`` as views of the receive buffer
    ``out`` (pass it back in to reuse it)."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world, packed.numel()), dtype=torch.float32, device=packed.device)
    if packed.is_cuda:
        dist.all_gather_into_tensor(out, packed, group=group)
    else:
        dist.all_gather(list(out.unbind(0)), packed, group=group)
    n = frames * k * 9
    return out[:, :n].view(world, frames, k, 9), out[:, n:].view(torch.int32), out


class PeerGather:
    """The same gather as ``gather_packed`` without a rendezvous: every rank owns a receive buffer ``(world, stride)``
    in its HBM, maps every peer's buffer through CUDA IPC once, and per step ONE kernel (``csrc/peer_put.cu``) stores
    its packed records into row ``rank`` of every rank's buffer over NVLink and then publishes the step number into
    that rank's flag word.  Nothing waits for a peer: an NCCL all_gather kernel holds SMs until all ranks have launched
    theirs, which at one collective per 0.6 ms step cost ~15 % of the 8-GPU throughput.  The records of step ``s``
    from rank ``r`` are complete once ``flags[r] >= s``; ``wait`` checks that on the host.  One process per GPU on one
    node; every rank must construct its PeerGather objects in the same order (the IPC handles travel through
    ``all_gather_object``)."""

    def __init__(self, numel: int, device, group=None):
        import ctypes

        from torch.multiprocessing.reductions import reduce_tensor

        from . import _lib

        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.numel = numel
        self.stride = (numel + 3) // 4 * 4  # rows start 16-byte aligned
        self.recv = torch.zeros((self.world, self.stride), dtype=torch.float32, device=device)
        self.flags = torch.zeros((self.world,), dtype=torch.int64, device=device)
        self._sync = torch.zeros((2,), dtype=torch.int32, device=device)  # [CTAs done, step]: owned by the kernel
        self.step = 0
        handles = [None] * self.world
        dist.all_gather_object(handles, [reduce_tensor(self.recv), reduce_tensor(self.flags)], group=group)
        self._peers = []  # keep the mappings alive
        rows, flags = (ctypes.c_void_p * self.world)(), (ctypes.c_void_p * self.world)()
        for r in range(self.world):
            if r == self.rank:
                pr, pf = self.recv, self.flags
            else:
                # torch's rebuild would open the IPC handle in the context of the PRODUCER's device index; kernels on
                # my device fault on such a mapping (measured, scripts/peer_diag.py).  Opening it with my own device
                # current (argument 6 of rebuild_cuda_tensor = the device to open under) maps the peer's memory into my
                # device's address space with peer access -- the pattern NCCL itself uses.
                (f0, a0), (f1, a1) = handles[r]
                a0, a1 = list(a0), list(a1)
                assert isinstance(a0[6], int) and isinstance(a1[6], int), "unexpected torch IPC handle layout"
                peer_dev = a0[6]
                a0[6] = a1[6] = self.recv.device.index
                with torch.cuda.device(self.recv.device):
                    _lib.call("tsmdet_enable_peer_access", peer_dev)
                    pr, pf = f0(*a0), f1(*a1)
            self._peers.append((pr, pf))
            rows[r] = pr.data_ptr() + self.rank * self.stride * 4
            flags[r] = pf.data_ptr() + self.rank * 8
        self._rows, self._flags = rows, flags

    def put(self, packed: torch.Tensor):
        """Stream-ordered, asynchronous (one kernel on the current stream); returns this rank's receive buffer, whose
        rows fill as the peers' stores land."""
        from . import _lib

        assert packed.is_cuda and packed.dtype == torch.float32 and packed.is_contiguous() and packed.numel() == self.numel
        self.step += 1
        _lib.call("tsmdet_peer_put", _lib.ptr(packed), self.numel, self.world, self._rows, self._flags,
                  _lib.ptr(self._sync), _lib.stream_ptr(packed.device))
        return self.recv

    def wait_stream(self):
        """Order the current stream after the arrival of every rank's records of this rank's last ``put`` (a one-warp
        kernel watching the flag words) -- what a consumer on the device, or a D2H copy of the result, needs."""
        from . import _lib

        _lib.call("tsmdet_peer_wait", _lib.ptr(self.flags), self.world, _lib.ptr(self._sync), -1,
                  _lib.stream_ptr(self.flags.device))

    def wait(self, step: Optional[int] = None, timeout_s: float = 30.0):
        """Block the host until every rank's records of ``step`` (default: the last ``put``) are in ``recv``."""
        import time

        want = self.step if step is None else step
        t0 = time.perf_counter()
        while not bool((self.flags >= want).all()):
            if time.perf_counter() - t0 > timeout_s:
                raise RuntimeError(f"PeerGather.wait: flags {self.flags.tolist()} never reached {want}")

    def views(self, frames: int, k: int):
        n = frames * k * 9
        return self.recv[:, :n].view(self.world, frames, k, 9), self.recv[:, n:self.numel].view(torch.int32)

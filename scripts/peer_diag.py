"""2-GPU diagnostic of the peer-memory gather: which way of mapping the peers' buffers lets a kernel store into them."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
from tsmdet_b200 import _lib
n = 4096
for mode in ("own_device", "peer_device"):
    recv = torch.zeros((world, n), dtype=torch.float32, device=dev)
    flags = torch.zeros((world,), dtype=torch.int64, device=dev)
    sync = torch.zeros((2,), dtype=torch.int32, device=dev)
    handles = [None] * world
    dist.all_gather_object(handles, [reduce_tensor(recv), reduce_tensor(flags)])
    rows, fl = (ctypes.c_void_p * world)(), (ctypes.c_void_p * world)()
    keep = []
    for r in range(world):
        if r == rank:
            pr, pf = recv, flags
        else:
            (f0, a0), (f1, a1) = handles[r]
            if mode == "own_device":  # open the IPC handle in MY device's context
                a0, a1 = list(a0), list(a1)
                a0[6] = dev.index; a1[6] = dev.index
            pr, pf = f0(*a0), f1(*a1)
            _lib.call("tsmdet_enable_peer_access", r)
        keep.append((pr, pf))
        rows[r] = pr.data_ptr() + rank * n * 4
        fl[r] = pf.data_ptr() + rank * 8
        print(f"[{mode}] rank {rank}: peer {r} tensor on {pr.device} ptr {pr.data_ptr():#x}", flush=True)
    src = torch.full((n,), float(rank + 1), device=dev)
    torch.cuda.synchronize(); dist.barrier()
    try:
        _lib.call("tsmdet_peer_put", _lib.ptr(src), n, world, rows, fl, _lib.ptr(sync), _lib.stream_ptr(dev))
        _lib.call("tsmdet_peer_wait", _lib.ptr(flags), world, _lib.ptr(sync), -1, _lib.stream_ptr(dev))
        torch.cuda.synchronize(); dist.barrier()
        want = torch.arange(1, world + 1, dtype=torch.float32, device=dev).unsqueeze(1).expand(world, n)
        print(f"[{mode}] rank {rank}: put ok, rows correct = {bool(torch.equal(recv, want))}, flags {flags.tolist()}", flush=True)
    except Exception as e:
        print(f"[{mode}] rank {rank}: FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)
        os._exit(0)
dist.barrier()
os._exit(0)

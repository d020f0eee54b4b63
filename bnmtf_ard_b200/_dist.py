"""Process-group plumbing (torch.distributed) for the one-process-per-GPU layout.
torch is only touched when the caller has already initialised a process group."""
import sys

import numpy as np


def _pg():
    td = sys.modules.get("torch.distributed")
    if td is None or not td.is_available() or not td.is_initialized():
        return None
    return td


def info():
    """(rank, world_size); (0, 1) when no process group is initialised."""
    td = _pg()
    if td is None:
        return 0, 1
    return td.get_rank(), td.get_world_size()


def shard_bounds(n, rank, world):
    """Equal contiguous shards of size ceil(n/world); the tail ranks may be short or empty.
    Returns (start, stop, shard_size)."""
    size = -(-n // world)
    start = min(n, rank * size)
    stop = min(n, start + size)
    return start, stop, size


def _device_for_backend(td):
    import torch
    if td.get_backend() == "nccl":
        from . import _lib
        return torch.device("cuda", _lib.current_device())
    return torch.device("cpu")


def allreduce_sum(values):
    """Sum a small float64 vector over all ranks (identity when world == 1)."""
    values = np.asarray(values, dtype=np.float64)
    td = _pg()
    if td is None or td.get_world_size() == 1:
        return values.copy()
    import torch
    t = torch.from_numpy(values.copy()).to(_device_for_backend(td))
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return t.cpu().numpy()


def allreduce_max(values):
    """Element-wise MAX of a small int64 vector over all ranks."""
    values = np.asarray(values, dtype=np.int64)
    td = _pg()
    if td is None or td.get_world_size() == 1:
        return values.copy()
    import torch
    t = torch.from_numpy(values.copy()).to(_device_for_backend(td))
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return t.cpu().numpy()


def broadcast_bytes(payload, src=0):
    """Broadcast a fixed-length bytes object from `src` to every rank."""
    td = _pg()
    if td is None or td.get_world_size() == 1:
        return payload
    import torch
    dev = _device_for_backend(td)
    if td.get_rank() == src:
        t = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(len(payload), dtype=torch.uint8, device=dev)
    td.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def allgather_bytes(payload):
    """Every rank's fixed-length bytes object, concatenated in rank order."""
    td = _pg()
    if td is None or td.get_world_size() == 1:
        return payload
    import torch
    dev = _device_for_backend(td)
    mine = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    parts = [torch.zeros_like(mine) for _ in range(td.get_world_size())]
    td.all_gather(parts, mine)
    return b"".join(bytes(t.cpu().tolist()) for t in parts)


def broadcast_int(value, src=0):
    """A 64-bit integer from rank `src` to every rank (identity when world == 1)."""
    td = _pg()
    if td is None or td.get_world_size() == 1:
        return int(value)
    import torch
    t = torch.tensor([int(value)], dtype=torch.int64, device=_device_for_backend(td))
    td.broadcast(t, src=src)
    return int(t.item())


def same_on_all_ranks(values):
    """True when every rank holds bitwise the same float64 vector (MAX == MIN of the bit patterns)."""
    td = _pg()
    if td is None or td.get_world_size() == 1:
        return True
    import torch
    bits = np.ascontiguousarray(np.asarray(values, dtype=np.float64)).view(np.int64)
    hi = torch.from_numpy(bits.copy()).to(_device_for_backend(td))
    lo = hi.clone()
    td.all_reduce(hi, op=td.ReduceOp.MAX)
    td.all_reduce(lo, op=td.ReduceOp.MIN)
    return bool(torch.equal(hi, lo))


def barrier():
    td = _pg()
    if td is not None and td.get_world_size() > 1:
        td.barrier()

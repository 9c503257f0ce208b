"""Multi-GPU sharding of the scan when one PROCESS per GPU is used (torchrun / bench.py --gpus N).

Reads are independent units: the input is cut into contiguous, record-aligned
ranges balanced BY BASES (ultra-long reads make read counts meaningless), rank r scans range r
on its own GPU, and there is no data-path collective.  The only exchanges are the small ones
here: a gather of the 40-byte result rows to rank 0 (file order = rank order, because the
ranges are contiguous) and the max / sum reductions of the timing.  Backend: NCCL on GPUs,
gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_by_bases(offsets: np.ndarray, world: int) -> list:
    """Contiguous read ranges [(lo, hi)] * world with (nearly) equal bases; offsets[n+1] = read starts."""
    n = len(offsets) - 1
    total = int(offsets[n])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        i = int(np.searchsorted(offsets, target, side="left"))
        cuts.append(min(max(i, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_by_count(n_reads: int, world: int) -> list:
    """Equal read counts per rank (synthetic workloads whose reads are iid)."""
    return [(n_reads * r // world, n_reads * (r + 1) // world) for r in range(world)]


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _device():
    import torch
    dist = _dist()
    if dist is not None and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def reduce_scalar(x: float, op: str = "max") -> float:
    """max / sum of a scalar over ranks (identity without a process group)."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def gather_rows(rows: np.ndarray, dst: int = 0):
    """Gather every rank's `tps_row` array on rank `dst`, concatenated in rank order.
    Returns the concatenation on `dst`, None elsewhere."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return rows.copy()
    world, rank, dev = dist.get_world_size(), dist.get_rank(), _device()
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[rank] = len(rows)
    dist.all_reduce(counts)
    cap = int(counts.max().item())
    item = rows.dtype.itemsize
    buf = np.zeros(cap * item, dtype=np.uint8)
    buf[:len(rows) * item] = np.frombuffer(rows.tobytes(), dtype=np.uint8)
    mine = torch.from_numpy(buf).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == dst else None
    dist.gather(mine, parts, dst=dst)
    if rank != dst:
        return None
    out = [np.frombuffer(p.cpu().numpy().tobytes()[:int(counts[r].item()) * item], dtype=rows.dtype)
           for r, p in enumerate(parts)]
    return np.concatenate(out)

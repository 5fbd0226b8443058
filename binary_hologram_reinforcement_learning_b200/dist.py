"""One process per GPU: sharding helpers and the single collective of the path.

Environments, targets and candidate ranges are independent, so the
hot path has no exchange step.  ``torch.distributed`` (NCCL over NVLink on the
GPU box, gloo in CPU tests) is used only to all-gather per-episode statistics and
to reduce the decile histograms of a sharded sweep.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np


def env_info() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (defaults 0,1,0)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
            int(os.environ.get("LOCAL_RANK", 0)))


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(local_rank: int):
    """NUMA node of GPU `local_rank` (sysfs), or None when the platform does not say."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa_node(local_rank: int):
    """Run this rank on the cores of its GPU's NUMA node, BEFORE it allocates pinned host memory.

    The observation blocks of the eager path are pinned host memory the GPU writes over PCIe every
    step; first-touch places them on the node of the allocating thread, and a block on the other
    socket sends every write across the inter-socket link.  Returns the node, or None (unknown
    topology, or the cpuset of the container has no core of that node: nothing is changed).
    """
    node = gpu_numa_node(local_rank)
    if node is None or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n units for this rank (candidate sweeps)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def shard_indices(n: int, rank: int, world: int) -> np.ndarray:
    """Round-robin env / target indices e = rank (mod world)."""
    return np.arange(rank, n, world, dtype=np.int64)


def init_process_group(backend: str = None):
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return
    rank, world, local = env_info()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend=backend, rank=rank, world_size=world)


def _device():
    import torch
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_episode_stats(stats: np.ndarray) -> np.ndarray:
    """All-gather a (k_rank, C) float64 array of per-episode rows; returns all rows, rank order.

    Rows are [episode_reward, steps, flip_count, initial_psnr, final_psnr] in the
    vectorised env, but any fixed column count works.  Ranks may hold different
    row counts (padded to the maximum for the collective).
    """
    import torch
    import torch.distributed as dist
    stats = np.asarray(stats, dtype=np.float64)
    if stats.ndim == 1:
        stats = stats.reshape(0, 0) if stats.size == 0 else stats.reshape(1, -1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats
    world, dev = dist.get_world_size(), _device()
    shape = torch.tensor([stats.shape[0], stats.shape[1] if stats.ndim == 2 else 0],
                         dtype=torch.int64, device=dev)
    shapes = [torch.zeros_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape)
    rows = [int(s[0]) for s in shapes]
    cols = max(int(s[1]) for s in shapes)
    mx = max(max(rows), 1)
    buf = torch.zeros((mx, max(cols, 1)), dtype=torch.float64, device=dev)
    if stats.size:
        buf[:stats.shape[0], :stats.shape[1]] = torch.from_numpy(stats).to(dev)
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    parts = [o[:r, :cols].cpu().numpy() for o, r in zip(outs, rows)]
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, cols))


def reduce_histograms(*arrays: np.ndarray) -> List[np.ndarray]:
    """Sum decile counters (attempted / improved / gain) over ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [np.asarray(a) for a in arrays]
    dev = _device()
    out = []
    for a in arrays:
        a = np.asarray(a)
        t = torch.from_numpy(a.astype(np.float64)).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out.append(t.cpu().numpy().astype(a.dtype))
    return out


def max_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()

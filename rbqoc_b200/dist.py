"""Multi-GPU plumbing for the Monte-Carlo sweeps: one process per GPU, samples sharded by contiguous
global index range, no data-path collective; the per-rank `rbq_stats` blocks (2.1 KB) meet in ONE
all-gather and are merged with `rbq_stats_merge`.  torch.distributed is plumbing only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import Stats
from .engine import STATS_INT64_WORDS, stats_from_words, stats_merge


def shard_range(S: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [first, first+count) of rank's samples; sizes differ by at most one."""
    base, rem = divmod(int(S), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def stats_to_words(st: Stats) -> np.ndarray:
    out = np.empty(STATS_INT64_WORDS, dtype=np.int64)
    C.memmove(out.ctypes.data, C.byref(st), C.sizeof(Stats))
    return out


def allgather_merge_stats(words, group=None) -> Stats:
    """words: torch int64 tensor (STATS_INT64_WORDS,) holding this rank's rbq_stats, on the device for
    NCCL or on the CPU for gloo.  One all_gather_into_tensor, then a host merge in rank order
    (deterministic for a given world size)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    gathered = torch.empty(world * STATS_INT64_WORDS, dtype=torch.int64, device=words.device)
    dist.all_gather_into_tensor(gathered, words.contiguous(), group=group)
    host = gathered.cpu().numpy().reshape(world, STATS_INT64_WORDS)
    total = stats_from_words(host[0])
    for r in range(1, world):
        stats_merge(total, stats_from_words(host[r]))
    return total


def bind_to_gpu_numa_node(device: int) -> list[int] | None:
    """Pin this process to the CPUs NVML reports as local to `device` (what `numactl --cpunodebind` would do), so that
    page-locked buffers allocated afterwards are first-touched on the GPU's own NUMA node and host<->device traffic does not
    cross the socket interconnect.  One process per GPU makes this the natural placement; returns the CPU list, or None
    when NVML / the affinity call is unavailable (nothing is changed then)."""
    import os

    try:
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(int(device))
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None

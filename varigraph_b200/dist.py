"""Multi-GPU plumbing of the count phase (one process per GPU, torch.distributed).

The path shards over reads: every rank counts its own shard against a replica of the index, and
one integer all-reduce per sample combines the per-rank count vectors.  Counts are saturating
sums, so min(255, sum_r min(255, c_r)) == min(255, total occurrences): exact.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def bind_to_gpu_numa(device_index: int) -> str:
    """Pin this process to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function)
    BEFORE it allocates pinned staging memory, so that first touch puts the buffers on the GPU's own
    NUMA node and the H2D DMA of eight ranks does not funnel through one socket.  Returns what it did."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "no local cpus in the affinity mask"
        os.sched_setaffinity(0, cpus)
        return f"gpu {device_index} ({bdf}): cpus {spec}"
    except Exception as ex:  # plumbing only: never fatal
        return f"not bound ({type(ex).__name__}: {ex})"


def shard_bounds(lines: np.ndarray, world: int) -> list:
    """Cut a staged chunk ('read\\n' records) into `world` contiguous shards at read boundaries.
    -> [(begin, end)] * world covering the buffer exactly; a shard may be empty."""
    n = int(lines.size)
    cuts = [0]
    for r in range(1, world):
        p = min(max(n * r // world, cuts[-1]), n)
        while p < n and p > 0 and lines[p - 1] != 10:  # advance to just after a newline
            p += 1
        cuts.append(p)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def reduce_counts(counts32: torch.Tensor, group=None) -> torch.Tensor:
    """Sum per-rank u8-valued counts held in an int32 tensor across ranks, clamp to 255 -> uint8."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts32, op=dist.ReduceOp.SUM, group=group)
    return counts32.clamp_(max=255).to(torch.uint8)

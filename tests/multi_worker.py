"""Worker of the multi-GPU tests: one process per rank (torch.multiprocessing.spawn), handles carried
by torch.distributed (gloo).  With fewer GPUs than ranks the ranks share a device -- CUDA IPC and the
peer-flag barrier work between processes on one GPU as well (the contexts are time-sliced), which is
how the 1-GPU test box still exercises the world-size-2 code."""
from __future__ import annotations

import os

import numpy as np


def exchange_with_dist(dist):
    def _x(mine: bytes):
        got = [None] * dist.get_world_size()
        dist.all_gather_object(got, mine)
        return got
    return _x


def submit_in_rounds(ix, lines: np.ndarray, dist) -> int:
    """Feed a rank's staged reads to a sharded index: as many collective rounds as the slowest rank needs."""
    off, rounds = 0, 0
    while True:
        more = 1 if off < lines.size else 0
        if dist is not None:
            import torch
            t = torch.tensor([more])
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if int(t) == 0:
                break
        elif not more:
            break
        room = ix.room()
        end = min(lines.size, off + room)
        if end < lines.size:  # cut at a read boundary
            nl = np.flatnonzero(lines[off:end] == 10)
            end = off + (int(nl[-1]) + 1 if nl.size else 0)
        if end > off:
            ix.submit(lines[off:end])
        off = end
        ix.flush()
        rounds += 1
    return rounds


def worker(rank: int, world: int, port: int, scenario: str, path: str, env: dict) -> None:
    os.environ.update(env)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from varigraph_b200 import capi
    from varigraph_b200 import dist as vdist
    import datetime
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=120))
    try:
        z = np.load(path + ".in.npz")
        keys, lines, k = z["keys"], z["lines"], int(z["k"])
        dev = rank % torch.cuda.device_count()
        ctx = capi.Context(dev, buffer_mb=1)
        comm = capi.Comm(ctx, rank, world, int(z["arena"]), exchange_with_dist(dist))
        b, e = vdist.shard_bounds(lines, world)[rank]
        mine = lines[b:e]
        out = {}
        if scenario == "replicated":
            ix = capi.Index(ctx, keys, k)
            for rep in range(2):  # twice: the arena scratch and the barrier epochs are reused
                ix.begin()
                ix.submit(mine)
                counts = comm.allreduce_counts(ix)
                pos, hits = ix.stats()
                ix.end(want_counts=False)
            out = dict(counts=counts, pos=pos, hits=hits)
        elif scenario == "replica":
            # rank 0 alone builds the index; the other rank receives table, slot order and pre-filter over peer memory
            ix = comm.replicate(0, capi.Index(ctx, keys, k) if rank == 0 else None)
            perm = ix.slot_perm()
            for rep in range(2):
                ix.begin()
                ix.submit(mine)
                slots, _ = comm.allreduce_slots(ix)
                pos, hits = ix.stats()
                ix.end(want_counts=False)
            ix.begin()
            ix.submit(mine)
            counts = comm.allreduce_counts(ix)  # the key-order call on a replica: slot-order reduce + one gather
            ix.end(want_counts=False)
            out = dict(counts=counts, slot_counts=slots[perm], perm=perm, pos=pos, hits=hits, parts=ix.partitions, n=ix.n)
        elif scenario in ("sharded", "sharded_device"):
            if scenario == "sharded_device":  # the keys already sit on this rank's GPU
                dk = torch.from_numpy(keys.view(np.int64)).cuda(dev)
                torch.cuda.synchronize(dev)
                ix = capi.Index(ctx, (dk.data_ptr(), keys.size), k, comm=comm, round_bytes=int(z["round_bytes"]))
            else:
                ix = capi.Index(ctx, keys, k, comm=comm, round_bytes=int(z["round_bytes"]))
            for rep in range(2):
                ix.begin()
                rounds = submit_in_rounds(ix, mine, dist)
                counts, pos, hits = ix.end()
            out = dict(counts=counts, pos=pos, hits=hits, rounds=rounds, own=ix.own_keys, parts=ix.partitions)
        else:
            raise ValueError(scenario)
        comm.check()
        out["device"] = dev
        np.savez(f"{path}.out{rank}.npz", **out)
        dist.barrier()  # nobody unmaps a peer's arena while it is still in use
        ix.close()
        comm.close()
        ctx.close()
    finally:
        dist.destroy_process_group()

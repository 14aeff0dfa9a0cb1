"""CPU suite, part 3: the N>1 host logic (read sharding + the one integer all-reduce) with gloo,
world_size 2.  The per-rank count here comes from the oracle -- in the test only, standing in for
the CUDA kernel -- so what is checked is the sharding and the reduce: min(255, sum of per-rank
saturated counts) must equal the single-process result, including for saturating k-mers."""
import datetime
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers
from varigraph_b200 import dist as vdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, keys, lines, k, out_path):
    from tests import oracle_binding as ob
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=120))
    b, e = vdist.shard_bounds(lines, world)[rank]
    mine, pos, hits = ob.Oracle().count_lines(keys, lines[b:e], k)
    total = vdist.reduce_counts(torch.from_numpy(mine.astype(np.int32)))
    tp = torch.tensor([pos, hits], dtype=torch.int64)
    dist.all_reduce(tp)
    if rank == 0:
        np.savez(out_path, counts=total.numpy(), pos=int(tp[0]), hits=int(tp[1]))
    dist.destroy_process_group()


def test_shard_bounds_cover_and_respect_reads():
    rng = np.random.default_rng(0)
    reads = [bytes(rng.choice(list(b"ACGT"), size=rng.integers(1, 40)).tolist()) for _ in range(101)]
    buf = np.frombuffer(b"\n".join(reads) + b"\n", dtype=np.uint8)
    for world in (1, 2, 3, 8, 200):
        b = vdist.shard_bounds(buf, world)
        assert b[0][0] == 0 and b[-1][1] == buf.size and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        for s, e in b:
            assert s == e or (buf[e - 1] == 10 and (s == 0 or buf[s - 1] == 10))


@pytest.mark.parametrize("saturate", [False, True])
def test_two_rank_reduce_matches_single_process(tmp_path, oracle, saturate):
    t = helpers.tiny()
    lines = t["lines"]
    if saturate:  # 300 copies of one read: its k-mers exceed 255 only after the ranks are combined
        best = max(range(300), key=lambda i: oracle.count_lines(t["keys"], lines[i * 151:(i + 1) * 151], t["k"])[2])
        hot = np.tile(lines[best * 151:(best + 1) * 151], 300)
        lines = np.concatenate([hot[: 150 * 151], lines, hot[150 * 151:]])
    want, wpos, whits = oracle.count_lines(t["keys"], lines, t["k"])
    out = str(tmp_path / "r0.npz")
    helpers.spawn_ranks(_worker, lambda port: (2, port, t["keys"], lines, t["k"], out), nprocs=2)
    z = np.load(out)
    assert np.array_equal(z["counts"], want)
    assert (int(z["pos"]), int(z["hits"])) == (wpos, whits)
    if saturate:
        assert want.max() == 255


class _FakeShardedIndex:
    """Stands in for a sharded vg_index on the CPU: a round takes `round_bytes`, flush() is the collective."""
    def __init__(self, round_bytes):
        self.round_bytes, self.used, self.flushes, self.got = round_bytes, 0, 0, []

    def room(self):
        return self.round_bytes - self.used

    def submit(self, b):
        assert 0 < b.size <= self.room() and b[-1] == 10
        self.used += b.size
        self.got.append(bytes(b))

    def flush(self):
        self.used = 0
        self.flushes += 1


def _rounds_worker(rank, world, port, sizes, out_path):
    from tests import multi_worker
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=120))
    rng = np.random.default_rng(rank)
    reads = [bytes(rng.choice(list(b"ACGT"), size=rng.integers(1, 90)).tolist()) + b"\n" for _ in range(sizes[rank])]
    lines = np.frombuffer(b"".join(reads), dtype=np.uint8)
    ix = _FakeShardedIndex(1000)
    rounds = multi_worker.submit_in_rounds(ix, lines, dist)
    ok = b"".join(ix.got) == lines.tobytes() and rounds == ix.flushes
    np.savez(f"{out_path}.{rank}.npz", rounds=rounds, ok=ok, nbytes=lines.size)
    dist.destroy_process_group()


def test_sharded_rounds_are_collective(tmp_path):
    """Ranks with very different amounts of reads still make the same number of collective flushes, and
    every rank submits all of its reads, cut at read boundaries, never beyond the room of a round."""
    out = str(tmp_path / "rounds")
    helpers.spawn_ranks(_rounds_worker, lambda port: (2, port, (400, 7), out), nprocs=2)
    z = [np.load(f"{out}.{r}.npz") for r in range(2)]
    assert bool(z[0]["ok"]) and bool(z[1]["ok"])
    assert int(z[0]["rounds"]) == int(z[1]["rounds"]) >= int(z[0]["nbytes"]) // 1000

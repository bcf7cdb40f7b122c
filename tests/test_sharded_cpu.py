"""The N > 1 path on CPU: world_size-2 (and 3) process groups over gloo run ttv_b200.sharded.ttv_sharded with the
kernel call replaced by the oracle, which exercises the partition arithmetic (slab offsets, uneven splits, b slices) and
the collective (reduce / all-reduce of the partial C) without a GPU.  The GPU run of the same driver is bench.py under
torchrun."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ttv_b200.sharded import make_shard, split_range, ttv_sharded  # noqa: E402

CASES = [
    # (na, pia, dtype)
    ((5, 4, 6), (1, 2, 3), np.float64),
    ((5, 4, 7), (3, 1, 2), np.int64),          # slowest mode is 2; uneven split of 4 over 3 ranks too
    ((3, 8, 2, 5), (2, 4, 1, 3), np.float32),
    ((6, 5), (2, 1), np.complex128),
]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, failures):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import Oracle
        oracle = Oracle()

        def compute(q, a, na, pia, b, c):       # the checker stands in for the kernel; c is a torch CPU tensor
            out = oracle.ttv(q, a.numpy(), na, pia, b.numpy())
            c.copy_(torch.from_numpy(out))

        rng = np.random.default_rng(42)        # same data on every rank
        for na, pia, dtype in CASES:
            n = int(np.prod(na))
            a_full = rng.integers(-5, 6, n).astype(dtype)
            for q in range(1, len(na) + 1):
                b = rng.integers(-5, 6, na[q - 1]).astype(dtype)
                want = oracle.ttv(q, a_full, na, pia, b)
                sh = make_shard(q, na, pia, rank, world)
                a_local = torch.from_numpy(a_full[sh.a_offset: sh.a_offset + sh.a_count].copy())
                for reduce_to in (0, None):
                    c, sh2 = ttv_sharded(q, a_local, na, pia, torch.from_numpy(b), rank=rank, world=world,
                                         reduce_to=reduce_to, compute=compute)
                    assert sh2 == sh
                    if sh.kind == "free":
                        got = c.numpy()
                        ok = np.array_equal(got, want[sh.c_offset: sh.c_offset + sh.c_count])
                    elif reduce_to is None or rank == 0:
                        ok = np.array_equal(c.numpy(), want)
                    else:
                        ok = True
                    if not ok:
                        failures.put((rank, na, pia, q, str(dtype), reduce_to))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_partition_and_collective_over_gloo(world):
    ctx = mp.get_context("spawn")
    failures = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, failures)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert failures.empty(), failures.get()


def test_split_range_and_shards_cover_everything():
    for extent in (1, 2, 7, 8, 2048):
        for world in (1, 2, 3, 4, 8):
            ranges = [split_range(extent, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and sum(c for _, c in ranges) == extent
            assert all(ranges[i][0] + ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in ranges) - min(c for _, c in ranges) <= 1
    na, pia = (2048, 2048, 2048), (1, 2, 3)
    for q, kind in ((1, "free"), (2, "free"), (3, "nq")):
        shards = [make_shard(q, na, pia, r, 8) for r in range(8)]
        assert all(s.kind == kind and s.mode == 3 and s.count == 256 for s in shards)
        assert sum(s.a_count for s in shards) == 2048 ** 3
        assert [s.a_offset for s in shards] == [r * 256 * 2048 * 2048 for r in range(8)]
        if kind == "free":
            assert sum(s.c_count for s in shards) == 2048 ** 2 and shards[3].c_offset == 3 * 256 * 2048
        else:
            assert all(s.c_count == 2048 ** 2 and s.c_offset == 0 for s in shards)
    # last-order layout: the slowest mode is mode 1
    s = make_shard(1, (2048, 64, 32), (3, 2, 1), 1, 2)
    assert s.mode == 1 and s.kind == "nq" and s.begin == 1024 and s.na_local == (1024, 64, 32)

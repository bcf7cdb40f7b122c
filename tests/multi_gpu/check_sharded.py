"""Multi-GPU parity of ttv_b200.sharded, one process per GPU (NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multi_gpu/check_sharded.py
Every rank builds the same global integer-valued tensor, owns its slab, and the three exchange forms -- free split (none),
n_q split + NCCL reduce / all-reduce, n_q split fused with the exchange over peer memory (PeerExchange, as one kernel per
GPU with an in-kernel barrier and as scatter kernel + library barrier + reduce kernel) -- are compared
bit for bit with the oracle on the global problem.  Prints "multi-gpu parity ok" on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.oracle import Oracle  # noqa: E402
from ttv_b200.sharded import PeerExchange, make_shard, ttv_sharded  # noqa: E402

CASES = [((37, 21, 40), (1, 2, 3)), ((16, 33, 9, 12), (2, 1, 4, 3)), ((300, 17), (1, 2)), ((17, 300), (2, 1)), ((5, 6, 7, 19), (4, 3, 2, 1))]


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    oracle = Oracle()
    checked = 0
    for dtype in (np.float32, np.float64, np.int32, np.complex64):
        tdt = torch.from_numpy(np.zeros(1, dtype)).dtype
        ex = PeerExchange(40000, tdt, dev)                          # ONE kernel per GPU: scatter + in-kernel barrier + sum
        ex3 = PeerExchange(40000, tdt, dev, single_kernel=False)   # scatter kernel + symmetric-memory barrier + reduce kernel
        rng = np.random.default_rng(99)                     # the same data on every rank
        for na, pia in CASES:
            n = int(np.prod(na))
            a_full = rng.integers(-6, 7, n).astype(dtype)
            for q in range(1, len(na) + 1):
                b = rng.integers(-6, 7, na[q - 1]).astype(dtype)
                want = oracle.ttv(q, a_full, na, pia, b)
                sh = make_shard(q, na, pia, rank, world)
                a_loc = torch.from_numpy(a_full[sh.a_offset: sh.a_offset + sh.a_count].copy()).to(dev)
                tb = torch.from_numpy(b).to(dev)
                # plain: free split, or n_q split + NCCL all-reduce
                c, s2 = ttv_sharded(q, a_loc, na, pia, tb, rank=rank, world=world, reduce_to=None)
                got = c.cpu().numpy()
                if s2.kind == "free":
                    assert np.array_equal(got, want[s2.c_offset: s2.c_offset + s2.c_count]), (na, pia, q, "free", rank)
                else:
                    assert np.array_equal(got, want), (na, pia, q, "nccl", rank)
                    # fused exchange over peer memory: this rank's block of C
                    for rnd in range(4):                    # several rounds: both halves of the workspace and their reuse
                        form = ex if rnd != 2 else ex3
                        c2, s3 = ttv_sharded(q, a_loc, na, pia, tb, rank=rank, world=world, exchange=form)
                        if na[s3.mode - 1] >= world:
                            assert s3.kind == "nq-scattered"
                            assert np.array_equal(c2.cpu().numpy(), want[s3.c_offset: s3.c_offset + s3.c_count]), (na, pia, q, "fused", rank, rnd)
                    assert not ex.timed_out(), (na, pia, q, "a single-kernel exchange timed out", rank)
                checked += 1
    t = torch.tensor([checked], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print(f"multi-gpu parity ok: world={world}, {checked} products per rank, {int(t.item())} in total", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

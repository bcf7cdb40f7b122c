"""Multi-GPU TTV on one node: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch) for the plumbing.

The reference is a single-process shared-memory library (OpenMP); it has no distributed path.  This module is the
B200-native counterpart of its outer-loop parallelism (reference detail/tensor_times_vector.h:583-1351): instead of
OpenMP threads over the free modes, GPUs over the slowest mode of the layout.

Partitioning of a global tensor (na, pia) over G ranks, always along the SLOWEST mode pi_p of the layout,
so that every rank owns one contiguous slab of A:

  q != pi_p   free-mode split: the slab is itself a packed tensor with extent na[pi_p]/G in mode pi_p; every rank
              computes its slab of C with the full b.  NO communication.
  q == pi_p   n_q split: rank r holds rows [r*nq/G, (r+1)*nq/G) of the contraction mode and the matching slice of b,
              computes a full-size partial C, and the partials are summed with ONE reduce / all-reduce (NCCL).
              Integer sums are exact; float sums change order, which the n_q*eps tolerance covers.

              With a PeerExchange (symmetric memory over NVLink / NVSwitch) the exchange is FUSED into the kernel: every
              rank's kernel stores its partial block by block straight into the owners' memory, a device-side barrier
              follows, and each rank sums the slots it received -- a deterministic reduce-scatter that leaves C
              distributed like the free split does (ttv_b200_view_scatter / ttv_b200_reduce_slots).

The arithmetic is always the C-ABI kernel on the local slab; nothing here computes on the CPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Sequence


def gpu_numa_cpus(device_index: int):
    """CPUs of the NUMA node the GPU's PCIe root port hangs off (sysfs local_cpulist of the device), or None when the
    platform does not say (single node, virtualised PCI topology)."""
    import torch
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            text = f.read().strip()
    except (OSError, AttributeError):
        return None
    cpus = set()
    for part in text.split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus or None


def bind_host_to_gpu(device_index: int) -> bool:
    """Pins the calling process to the CPUs next to its GPU, so that the pinned staging buffers it allocates afterwards
    (first touch) live in the memory of that socket and the H2D / D2H copies of its slab do not cross the socket
    interconnect.  With one process per GPU on a two-socket 8-GPU box every rank otherwise allocates wherever the
    scheduler happened to start it.  Returns False when nothing was changed."""
    import os
    cpus = gpu_numa_cpus(device_index)
    if not cpus or not hasattr(os, "sched_setaffinity"):
        return False
    allowed = os.sched_getaffinity(0)
    want = cpus & allowed
    if not want or want == allowed:
        return False
    os.sched_setaffinity(0, want)
    return True


@dataclass(frozen=True)
class Shard:
    """what one rank owns of the global problem"""
    rank: int
    world: int
    mode: int                 # the split mode pi_p (1-based)
    kind: str                 # "free" (no communication) or "nq" (partial C + reduce)
    begin: int                # first index of the split mode on this rank
    count: int                # number of indices of the split mode on this rank (may be 0 when world > extent)
    na_local: tuple           # shape of the local slab
    a_offset: int             # element offset of the slab inside the global A (packed strides)
    a_count: int              # elements of the slab
    c_offset: int             # element offset of the local C inside the global C (free split); 0 for the n_q split
    c_count: int              # elements of the local C


class PeerExchange:
    """Workspaces in symmetric memory for the fused n_q-split exchange: on every GPU two halves (rounds alternate, which
    orders the reuse of a half behind the barrier of the round in between) of [world][blk_cap] elements, all mapped into
    every process of the group (torch.distributed._symmetric_memory: peer pointers over NVLink)."""

    def __init__(self, max_c_elems: int, dtype, device, group=None, single_kernel: bool = True):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.dtype = dtype
        self.itemsize = torch.empty(0, dtype=dtype).element_size()
        self.blk_cap = self.block(max_c_elems, self.world)
        self.buf = symm_mem.empty(2 * self.world * self.blk_cap, dtype=dtype, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group.group_name)
        self.round = 0
        # single-kernel form (ttv_b200_view_exchange): one flag word per rank on every GPU, mapped into every process, plus
        # 16 bytes of local scratch (arrival counter + error flag).  Tokens = round numbers, they only grow.
        self.single_kernel = single_kernel
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=device)
        self.flags.zero_()
        self.flags_hdl = symm_mem.rendezvous(self.flags, self.group.group_name)
        self.scratch = torch.zeros(4, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        self.flags_hdl.barrier(channel=0)                    # every rank's flags are zero before anybody writes a token

    @staticmethod
    def block(n: int, world: int) -> int:
        """elements of C's flat index space per rank: equal blocks, rounded up to 256 elements (whole 16-byte vectors)"""
        return -(-(-(-n // world)) // 256) * 256

    def timed_out(self) -> bool:
        """True when a single-kernel exchange gave up waiting for a peer (reads the error word: synchronises)"""
        return bool(int(self.scratch[2].item()))

    def exchange(self, outer: int, nq: int, inner: int, a_local, b_local):
        """partial product of this rank -> peers' slots -> barrier -> sum of the received slots.  Returns (c_block, first,
        count): this rank's block [first, first + count) of the flat C.  single_kernel (default): all of it is ONE kernel
        launch per GPU (in-kernel flag barrier over NVLink); otherwise scatter kernel + symmetric-memory barrier + reduce kernel."""
        import torch
        from . import api
        n = outer * inner
        blk = self.block(n, self.world)
        if blk > self.blk_cap or a_local.dtype != self.dtype:
            raise ValueError("PeerExchange: workspace too small or element type differs")
        half = (self.round % 2) * self.world * self.blk_cap
        self.round += 1
        peers = [int(p) + half * self.itemsize for p in self.hdl.buffer_ptrs]
        first = min(n, self.rank * blk)
        count = max(0, min(blk, n - first))
        if self.single_kernel:
            # ONE launch: the kernel scatters its partials, tells every GPU that it has delivered round `self.round`, waits for
            # all deliveries into its own workspace and sums them
            c_block = torch.empty(count, dtype=self.dtype, device=a_local.device)
            api.ttv_view_exchange(outer, nq, inner, a_local, b_local, peers, [int(p) for p in self.flags_hdl.buffer_ptrs], self.rank, blk,
                                  c_block if count else None, self.round, self.scratch)
            return c_block, first, count
        api.ttv_view_scatter(outer, nq, inner, a_local, b_local, peers, self.rank, blk)
        self.hdl.barrier(channel=0)                      # every rank's partials have landed in every owner's slots
        c_block = torch.empty(count, dtype=self.dtype, device=a_local.device)
        if count:
            api.reduce_slots(self.buf[half: half + self.world * blk], c_block, count, blk, self.world)
        return c_block, first, count


def split_range(extent: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous balanced ranges: the first extent % world ranks get one more"""
    base, extra = divmod(extent, world)
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def make_shard(q: int, na: Sequence[int], pia: Sequence[int], rank: int, world: int) -> Shard:
    p = len(na)
    if p < 2 or len(pia) != p or not (1 <= q <= p):
        raise ValueError("make_shard: need order >= 2, a layout of the same length and 1 <= q <= p")
    mode = int(pia[-1])
    extent = int(na[mode - 1])
    begin, count = split_range(extent, world, rank)
    slab = 1
    for m in pia[:-1]:
        slab *= int(na[m - 1])          # elements per index of the slowest mode
    na_local = list(int(x) for x in na)
    na_local[mode - 1] = count
    kind = "nq" if mode == q else "free"
    if kind == "free":
        c_per_index = slab // int(na[q - 1])
        c_offset, c_count = begin * c_per_index, count * c_per_index
    else:
        c_offset, c_count = 0, slab
    return Shard(rank, world, mode, kind, begin, count, tuple(na_local), begin * slab, count * slab, c_offset, c_count)


def ttv_sharded(q: int, a_local, na: Sequence[int], pia: Sequence[int], b, *, rank: int, world: int, c_local=None,
                group=None, reduce_to: int | None = 0, compute: Callable | None = None, exchange: PeerExchange | None = None,
                asynchronous: bool = False):
    """One sharded TTV.  a_local: this rank's slab (flat, packed, see make_shard); b: the FULL vector (every rank
    holds it; it is tiny).  Returns (c_local, shard):
      free split  -> this rank's slab of C
      n_q split   -> the reduced C on rank `reduce_to` (or on every rank when reduce_to is None); other ranks get
                     their partial back
    `compute(q, a, na_local, pia, b, c)` defaults to the C-ABI kernel; the gloo CPU tests inject a checker here to
    exercise the partition arithmetic and the collective without a GPU.
    asynchronous=True only enqueues the local kernel on the current torch stream (TTV_B200_FLAG_ASYNC): back-to-back
    products then run without a host round trip between them, which a 1.3 ms kernel on eight ranks does feel."""
    import torch
    import torch.distributed as dist

    sh = make_shard(q, na, pia, rank, world)
    if compute is None:
        from . import api

        def compute(q_, a_, na_, pia_, b_, c_):
            nc = api.generate_output_shape(na_, q_); pic = api.generate_output_layout(pia_, q_)
            api.ttv_lowlevel(q_, len(na_), a_, na_, api.generate_strides(na_, pia_), pia_, b_, [int(b_.shape[0])], c_, nc,
                             api.generate_strides(nc, pic), pic, flags=api.FLAG_ASYNC if asynchronous else 0)

    if c_local is None:
        c_local = torch.empty(sh.c_count, dtype=a_local.dtype, device=a_local.device)
    if sh.kind == "free":
        if sh.count:
            compute(q, a_local, list(sh.na_local), list(pia), b, c_local)
        return c_local, sh

    # n_q split
    b_part = b[sh.begin: sh.begin + sh.count]
    if exchange is not None and world > 1 and int(na[sh.mode - 1]) >= world:
        # fused: the kernel's stores ARE the exchange; C comes back distributed (this rank's block of the flat C)
        import dataclasses
        c_block, first, count = exchange.exchange(1, sh.count, sh.c_count, a_local, b_part)
        return c_block, dataclasses.replace(sh, kind="nq-scattered", c_offset=first, c_count=count)
    if sh.count:
        compute(q, a_local, list(sh.na_local), list(pia), b_part, c_local)
    else:
        c_local.zero_()
    if world > 1:
        view = torch.view_as_real(c_local) if c_local.is_complex() else c_local     # complex reduces as 2x real
        if reduce_to is None:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(view, dst=reduce_to, op=dist.ReduceOp.SUM, group=group)
    return c_local, sh

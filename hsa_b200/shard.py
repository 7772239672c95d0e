"""Host-side read sharding for the multi-GPU path.

The search shards by reads: every rank holds a replica of the index and searches a contiguous block of the
input read array, so there is no collective on the data path.  The only exchanges are the one-time index
broadcast (device blocks, NCCL) and -- when one process wants the whole result -- an ordered gather of the
per-rank results.  These helpers are backend-agnostic `torch.distributed` code (nccl on the GPU box, gloo
in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

# The reference processes reads in batches of 0x186A0 (bwtaln.c:477) and two per-batch quantities leak into
# the per-read results: the N filter uses bwa_cal_maxdiff(max_len of the BATCH) (bwtaln.c:267-274, 314-317),
# and the option switch after the first splice fallback is per batch.  Aligning shard
# boundaries to REF_BATCH makes every rank see whole reference batches, so ragged-length inputs give the
# same results sharded as unsharded; with fixed-length reads any alignment does.
REF_BATCH = 0x186A0


def shard_bounds(n_reads: int, world: int, align: int = 1) -> list[tuple[int, int]]:
    """Contiguous [lo, hi) per rank, sizes differing by less than 2 x `align`, boundaries multiples of `align`."""
    if world < 1:
        raise ValueError("world must be >= 1")
    units = (n_reads + align - 1) // align
    out, lo = [], 0
    for r in range(world):
        u = units // world + (1 if r < units % world else 0)
        hi = min(n_reads, lo + u * align)
        out.append((lo, hi))
        lo = hi
    return out


def shard_reads(codes: np.ndarray, off: np.ndarray, lens: np.ndarray, rank: int, world: int, align: int = 1):
    """This rank's slice of a concatenated read set: (codes, off rebased to 0, lens, (lo, hi))."""
    lo, hi = shard_bounds(int(lens.shape[0]), world, align)[rank]
    if hi == lo:
        return codes[:0], off[:0], lens[:0], (lo, hi)
    b0 = int(off[lo])
    b1 = int(off[hi - 1]) + int(lens[hi - 1])
    return codes[b0:b1], (off[lo:hi] - off[lo]).astype(off.dtype), lens[lo:hi], (lo, hi)


def broadcast_bytes(t: torch.Tensor | None, nbytes: int, src: int, device) -> torch.Tensor:
    """One-time index replication: rank `src` passes its uint8 view of the device blocks, everyone else
    receives into a fresh buffer."""
    if dist.get_rank() != src:
        t = torch.empty(nbytes, dtype=torch.uint8, device=device)
    assert t is not None and t.numel() == nbytes
    dist.broadcast(t, src)
    return t


def gather_in_input_order(n_aln: np.ndarray, aln: np.ndarray, dst: int = 0):
    """Gather per-rank results (n_aln[n_local], hits[(sum n_aln), 9] in item order) to rank `dst`, concatenated
    in input order (= rank order, because shards are contiguous).  Returns (n_aln_all, aln_all) on dst, else
    (None, None).  Uses gather_object-free fixed-width exchanges: sizes first, then padded payloads."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = torch.tensor([int(n_aln.shape[0]), int(aln.shape[0])], dtype=torch.int64)
    all_sizes = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    max_n = max(int(s[0]) for s in all_sizes)
    max_a = max(int(s[1]) for s in all_sizes)
    pn = torch.zeros(max(max_n, 1), dtype=torch.int32)
    pn[: n_aln.shape[0]] = torch.from_numpy(np.ascontiguousarray(n_aln, dtype=np.int32))
    pa = torch.zeros((max(max_a, 1), 9), dtype=torch.int64)
    if aln.shape[0]:
        pa[: aln.shape[0]] = torch.from_numpy(np.ascontiguousarray(aln).astype(np.int64))
    gn = [torch.zeros_like(pn) for _ in range(world)] if rank == dst else None
    ga = [torch.zeros_like(pa) for _ in range(world)] if rank == dst else None
    dist.gather(pn, gn, dst=dst)
    dist.gather(pa, ga, dst=dst)
    if rank != dst:
        return None, None
    n_all = np.concatenate([gn[r][: int(all_sizes[r][0])].numpy() for r in range(world)])
    a_all = np.concatenate([ga[r][: int(all_sizes[r][1])].numpy().astype(np.uint32) for r in range(world)])
    return n_all, a_all

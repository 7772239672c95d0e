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


def gather_in_input_order(n_aln: np.ndarray, aln: np.ndarray, dst: int = 0, device=None):
    """Gather per-rank results (n_aln[n_local], hits[(sum n_aln), 9] in item order) to rank `dst`, concatenated
    in input order (= rank order, because shards are contiguous).  Returns (n_aln_all, aln_all) on dst, else
    (None, None).  Fixed-width exchanges: sizes first, then payloads padded to the largest shard.
    `device`: stage the payloads through this device (the NCCL backend moves device tensors: host results -> HBM ->
    NVLink -> rank dst -> host); None = host tensors (gloo)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cpu") if device is None else torch.device(device)
    sizes = torch.tensor([int(n_aln.shape[0]), int(aln.shape[0])], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [s.cpu().tolist() for s in all_sizes]
    max_n = max(1, max(s[0] for s in all_sizes))
    max_a = max(1, max(s[1] for s in all_sizes))
    pn = torch.zeros(max_n, dtype=torch.int32, device=dev)
    pn[: n_aln.shape[0]].copy_(torch.from_numpy(np.ascontiguousarray(n_aln, dtype=np.int32)), non_blocking=True)
    pa = torch.zeros((max_a, 9), dtype=torch.int32, device=dev)
    if aln.shape[0]:
        pa[: aln.shape[0]].copy_(torch.from_numpy(np.ascontiguousarray(aln, dtype=np.uint32).view(np.int32)), non_blocking=True)
    gn = [torch.zeros_like(pn) for _ in range(world)] if rank == dst else None
    ga = [torch.zeros_like(pa) for _ in range(world)] if rank == dst else None
    dist.gather(pn, gn, dst=dst)
    dist.gather(pa, ga, dst=dst)
    if rank != dst:
        return None, None
    n_all = torch.cat([gn[r][: all_sizes[r][0]] for r in range(world)]).cpu().numpy()
    a_all = torch.cat([ga[r][: all_sizes[r][1]] for r in range(world)]).cpu().numpy().view(np.uint32)
    return n_all, a_all


_pinned_cache: dict = {}


def _pinned(name: str, n: int, dtype, cols: int = 0) -> torch.Tensor:
    """A re-used pinned host buffer of at least n rows (D2H into pageable memory runs at a fraction of the PCIe rate)."""
    t = _pinned_cache.get(name)
    if t is None or t.shape[0] < n or t.dtype != dtype:
        shape = (n + n // 8 + 16, cols) if cols else (n + n // 8 + 16,)
        t = torch.empty(shape, dtype=dtype)
        if torch.cuda.is_available():
            t = t.pin_memory()
        _pinned_cache[name] = t
    return t[:n]


def gather_results(n_aln: np.ndarray, aln_off: np.ndarray, aln: np.ndarray, device=None, dst: int = 0):
    """Ordered gather of the flat result of a batch as the C ABI returns it (hsa_result_t: n_aln[n_local], aln_off[n_local]
    = first hit of the item in this rank's hit arena, aln[(hits), 9]) to rank `dst`: (n_aln of all reads in input order,
    aln_off rebased into the concatenated arena, the concatenated arena) on dst, (None, None, None) elsewhere.  The arenas
    are moved as they are -- hits of one item stay contiguous and in discovery order, items are addressed through aln_off.
    `device`: stage through this device (NCCL); on dst the outputs land in re-used pinned host buffers (valid until the
    next call)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cpu") if device is None else torch.device(device)
    sizes = torch.tensor([int(n_aln.shape[0]), int(aln.shape[0])], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [s.cpu().tolist() for s in all_sizes]
    max_n = max(1, max(s[0] for s in all_sizes))
    max_a = max(1, max(s[1] for s in all_sizes))
    pn = torch.zeros(max_n, dtype=torch.int32, device=dev)
    po = torch.zeros(max_n, dtype=torch.int64, device=dev)
    pa = torch.zeros((max_a, 9), dtype=torch.int32, device=dev)
    nl, al = int(n_aln.shape[0]), int(aln.shape[0])
    if nl:
        pn[:nl].copy_(torch.from_numpy(np.ascontiguousarray(n_aln, dtype=np.int32)), non_blocking=True)
        po[:nl].copy_(torch.from_numpy(np.ascontiguousarray(aln_off).view(np.int64)), non_blocking=True)
    if al:
        pa[:al].copy_(torch.from_numpy(np.ascontiguousarray(aln, dtype=np.uint32).view(np.int32)), non_blocking=True)
    gn = [torch.zeros_like(pn) for _ in range(world)] if rank == dst else None
    go = [torch.zeros_like(po) for _ in range(world)] if rank == dst else None
    ga = [torch.zeros_like(pa) for _ in range(world)] if rank == dst else None
    dist.gather(pn, gn, dst=dst)
    dist.gather(po, go, dst=dst)
    dist.gather(pa, ga, dst=dst)
    if rank != dst:
        return None, None, None
    base, offs = 0, []
    for r in range(world):                                   # rebase every rank's offsets into the concatenated arena
        offs.append(go[r][: all_sizes[r][0]] + base)
        base += all_sizes[r][1]
    n_tot = sum(s[0] for s in all_sizes)
    out_n = _pinned("n_aln", n_tot, torch.int32)
    out_o = _pinned("aln_off", n_tot, torch.int64)
    out_a = _pinned("aln", max(base, 1), torch.int32, 9)[:base]
    out_n.copy_(torch.cat([gn[r][: all_sizes[r][0]] for r in range(world)]), non_blocking=True)
    out_o.copy_(torch.cat(offs), non_blocking=True)
    if base:
        out_a.copy_(torch.cat([ga[r][: all_sizes[r][1]] for r in range(world)]), non_blocking=True)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    return out_n.numpy(), out_o.numpy().view(np.uint64), out_a.numpy().view(np.uint32)


class HostGather:
    """Ordered gather of the per-rank results on the HOST, through one POSIX shared-memory segment ("per-GPU
    n_aln[] + packed bwt_aln1_t[] D2H, host concatenates in input order").

    Why not NCCL for this: the search kernels are persistent and fill every SM, so a collective's kernels queue behind them
    and the gather serialises with the next batch (measured at 8 GPUs: 123 M reads/s end to end against 158 M device-
    resident).  Results are already in pinned host memory after the D2H copy of hsa_job_wait; every rank copies its three
    arrays into its slot of the segment (ranks in parallel, no GPU involved) and rank `dst` reads the whole structure in
    place: n_aln / aln_off of all reads in input order, aln_off pointing into one arena in which rank r's hits start at
    r * max_hits.  `ctl` is a gloo process group (host-side barrier; an NCCL barrier would queue behind the kernels too)."""

    def __init__(self, max_items: int, max_hits: int, ctl=None, dst: int = 0):
        from multiprocessing import shared_memory
        self.world, self.rank, self.dst, self.ctl = dist.get_world_size(), dist.get_rank(), dst, ctl
        self.max_items, self.max_hits = int(max_items), int(max_hits)
        per_rank = self.max_items * 12 + self.max_hits * 36
        name = [None]
        if self.rank == dst:
            self.shm = shared_memory.SharedMemory(create=True, size=self.world * per_rank + 16 * self.world)
            name[0] = self.shm.name
        dist.broadcast_object_list(name, src=dst, group=ctl)
        if self.rank != dst:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:                                        # the segment belongs to dst: keep this process's tracker out of it
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        buf = self.shm.buf
        W, I, H = self.world, self.max_items, self.max_hits
        self.counts = np.ndarray((W, 2), dtype=np.int64, buffer=buf, offset=0)
        o = 16 * W
        self.n_aln = np.ndarray((W, I), dtype=np.int32, buffer=buf, offset=o); o += W * I * 4
        self.aln_off = np.ndarray((W, I), dtype=np.uint64, buffer=buf, offset=o); o += W * I * 8
        self.aln = np.ndarray((W, H, 9), dtype=np.uint32, buffer=buf, offset=o)

    def gather(self, n_aln: np.ndarray, aln_off: np.ndarray, aln: np.ndarray):
        """Every rank calls it with its batch result; on `dst` returns (n_aln, aln_off, arena) of all ranks' reads in input
        order as views of the shared segment (valid until the next call), elsewhere None."""
        n, h = int(n_aln.shape[0]), int(aln.shape[0])
        if n > self.max_items or h > self.max_hits:
            raise ValueError("result larger than the gather slot")
        r = self.rank
        self.n_aln[r, :n] = n_aln
        np.add(aln_off, np.uint64(r * self.max_hits), out=self.aln_off[r, :n])
        if h:
            self.aln[r, :h] = aln
        self.counts[r] = (n, h)
        dist.barrier(group=self.ctl)
        if r != self.dst:
            dist.barrier(group=self.ctl)                    # dst has consumed the segment: slots may be rewritten
            return None
        if all(int(self.counts[q, 0]) == self.max_items for q in range(self.world)):
            out = (self.n_aln.reshape(-1), self.aln_off.reshape(-1), self.aln.reshape(-1, 9))
        else:
            out = (np.concatenate([self.n_aln[q, : int(self.counts[q, 0])] for q in range(self.world)]),
                   np.concatenate([self.aln_off[q, : int(self.counts[q, 0])] for q in range(self.world)]),
                   self.aln.reshape(-1, 9))
        return out

    def release(self):
        """dst calls it when it is done with the views of the last gather()."""
        if self.rank == self.dst:
            dist.barrier(group=self.ctl)

    def close(self):
        self.n_aln = self.aln_off = self.aln = self.counts = None
        try:
            self.shm.close()
            if self.rank == self.dst:
                self.shm.unlink()
        except Exception:
            pass

"""Throughput of the other BASELINE.json configs through the public API (GPU box helper; bench.py measures configs[2] and
configs[1]).  Prints one JSON line per config:
  cfg0  4.6 Mb genome, 100 k x 75 bp, max_diff 2 / max_gapo 1            (hsa_whole_reads)
  cfg3  spliced reads on --genome: the six seed searches per read alone    (hsa_splice_seeds), and the whole path a spliced
        read takes -- whole-read search, then bwt_splice_match for what found nothing (hsa_whole_reads + hsa_splice_match_batch)
  cfg4  150 bp reads, 2 % substitutions, up to two indels, max_diff 5 / max_gapo 2 (hsa_whole_reads)
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from hsa_b200 import api, build, index_build, synth_torch  # noqa: E402


def pinned(t):
    return t.cpu().pin_memory()


def run(name, fn, n, reps=3):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        res = fn()
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, res)
    dt, res = best
    print(json.dumps({"config": name, "reads": n, "wall_ms": dt * 1e3, "kernel_ms": res.kernel_ms,
                      "reads_per_s_e2e": n / dt, "reads_per_s_kernel": n / (res.kernel_ms * 1e-3),
                      "occ_lookups": res.occ_lookups, "items_with_hits": int((res.n_aln > 0).sum()),
                      "heavy": res.n_strict}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=46_000_003)
    ap.add_argument("--reads", type=int, default=2_000_000)
    a = ap.parse_args()
    build.build_native()
    dev = torch.device("cuda", 0)

    # cfg0
    g0 = synth_torch.make_genome(4_600_003, 3, dev)
    ix0 = api.Index.upload(index_build.build_index(g0, device=dev), 0)
    n = 100_000
    r = synth_torch.simulate_reads(g0, n, 75, 11)
    codes, off, lens = pinned(r.reshape(-1)), pinned(torch.arange(n, dtype=torch.int64) * 75), pinned(torch.full((n,), 75, dtype=torch.int32))
    opt = api.gap_init_opt(fnr=0.0, max_diff=2, max_gapo=1)
    run("cfg0: 4.6 Mb, 100k x 75 bp, n2 o1", lambda: ix0.whole_reads(codes, off, lens, opt, copy=False), n)
    ix0.close()

    g = synth_torch.make_genome(a.genome, 1, dev)
    introns = synth_torch.plant_introns(g, max(100, a.genome // 150_000), 2)      # motif-carrying introns for cfg3
    ix = api.Index.upload(index_build.build_full_index(g, device=dev), 0)
    # cfg3: spliced reads = two exons across an intron
    for L in (100, 75):
        spliced_seeds(a, ix, g, dev, L)
        spliced_pipeline(a, ix, g, introns, L)
    # cfg4 stress
    n = max(a.reads // 4, 100_000)
    r = synth_torch.simulate_reads(g, n, 150, 21, sub_rate=0.02, indel_frac=0.10)
    codes, off, lens = pinned(r.reshape(-1)), pinned(torch.arange(n, dtype=torch.int64) * 150), pinned(torch.full((n,), 150, dtype=torch.int32))
    opt = api.gap_init_opt(fnr=0.0, max_diff=5, max_gapo=2)
    run(f"cfg4: {a.genome} bp, {n} x 150 bp, 2% subs, 10% indel reads, n5 o2", lambda: ix.whole_reads(codes, off, lens, opt, copy=False), n, reps=2)


def spliced_pipeline(a, ix, g, introns, L):
    """cfg3, the whole path of an RNA-seq read: hsa_whole_reads, then hsa_splice_match_batch (bwt_splice_match: seeds,
    correlation, motif scan, extension) for the reads that found nothing -- 3/4 junction reads over planted introns, 1/4
    ordinary reads."""
    n = a.reads
    nj = 3 * n // 4
    reads = torch.cat([synth_torch.simulate_junction_reads(g, introns, nj, L, 31 + L), synth_torch.simulate_reads(g, n - nj, L, 32 + L)])
    reads = reads[torch.randperm(n, device=reads.device, generator=torch.Generator(device=reads.device).manual_seed(3))]
    codes, off, lens = pinned(reads.reshape(-1)), pinned(torch.arange(n, dtype=torch.int64) * L), pinned(torch.full((n,), L, dtype=torch.int32))
    opt = api.gap_init_opt()
    sopt = api.gap_init_opt(mode=2, max_diff=api.bwa_cal_maxdiff(L))          # local_opt as the driver holds it from batch 2 on
    rd = codes.numpy().reshape(n, L)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        res = ix.whole_reads(codes, off, lens, opt, copy=False)
        t1 = time.perf_counter()
        un = np.nonzero(res.n_aln == 0)[0]
        sub = np.ascontiguousarray(rd[un]).reshape(-1)
        n_aln, _ = ix.splice_match(sub, (np.arange(un.shape[0], dtype=np.uint64) * L), np.full(un.shape[0], L, dtype=np.uint32), sopt)
        t2 = time.perf_counter()
        if best is None or t2 - t0 < best[0]:
            best = (t2 - t0, t1 - t0, t2 - t1, int(un.shape[0]), int((n_aln == 2).sum()), int((n_aln > 0).sum()), ix.last_splice_lookups)
    dt, t_whole, t_splice, n_un, two, any_, lk = best
    print(json.dumps({"config": f"cfg3 whole path: {a.genome} bp, {n} x {L} bp reads (75 % across planted introns) -> hsa_whole_reads, then "
                                f"hsa_splice_match_batch for the {n_un} reads that found nothing", "reads": n, "wall_ms": dt * 1e3,
                      "reads_per_s_e2e": n / dt, "whole_reads_ms": t_whole * 1e3, "splice_ms": t_splice * 1e3,
                      "splice_reads_per_s": n_un / t_splice, "spliced_two_part": two, "spliced_any": any_,
                      "splice_occ_lookups_per_read": lk / max(n_un, 1)}), flush=True)


def spliced_seeds(a, ix, g, dev, L):
    """cfg3: spliced reads = two exons across an intron; the six seed searches per read (len/3 bp each: 33/33/34 for
    100 bp reads, 25/25/25 for 75 bp reads -- the 25 bp figure of BASELINE.json's config) are what the hot path sees."""
    n = a.reads
    reads = synth_torch.simulate_spliced_reads(g, n, L, 5)
    codes, off, lens = pinned(reads.reshape(-1)), pinned(torch.arange(n, dtype=torch.int64) * L), pinned(torch.full((n,), L, dtype=torch.int32))
    opt = api.gap_init_opt()
    run(f"cfg3: {a.genome} bp, {n} spliced {L} bp reads -> 6 seed searches each ({L // 3} bp segments)",
        lambda: ix.splice_seeds(codes, off, lens, opt), n, reps=2)
    # the same, double-buffered through the asynchronous job pair (results left in the library's pinned buffers)
    def pipelined(steps):
        job = ix.splice_seeds_submit(codes, off, lens, opt)
        for k in range(steps):
            nxt = ix.splice_seeds_submit(codes, off, lens, opt) if k + 1 < steps else None
            last = job.wait(copy=False)
            job = nxt
        torch.cuda.synchronize()
        return last
    pipelined(4)
    t0 = time.perf_counter()
    last = pipelined(4)
    dt = time.perf_counter() - t0
    print(json.dumps({"config": f"cfg3 pipelined ({L} bp reads): 4 batches through hsa_splice_seeds_submit / hsa_job_wait", "reads": 4 * n,
                      "wall_ms": dt * 1e3, "reads_per_s_e2e": 4 * n / dt, "kernel_ms_per_batch": last.kernel_ms}), flush=True)


if __name__ == "__main__":
    main()

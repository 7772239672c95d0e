"""The real drop-in, measured (GPU box): the reference's own batch loop (bwtaln.c:477, 506) with bwa_cal_sa_reg_gap replaced by
shim/hsa_gpu_shim.c's bwa_cal_sa_reg_gap_gpu (oracle/_ref/hsa_ref_gpu gpudriver) next to the stock driver
(oracle/_ref/hsa_ref driver), splice fallback included, on a genome indexed by the reference builder.  Also the splice
path alone (hsa_splice_match_batch vs oracle/_ref/hsa_ref splice).  Prints one JSON line per measurement.
    python tools/bench_shim.py [--genome 46000003] [--reads 1000000]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from hsa_b200 import api, index_io, synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "hsa_ref")
REF_GPU = os.path.join(ROOT, "oracle", "_ref", "hsa_ref_gpu")


def run(cmd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    t0 = time.time()
    out = subprocess.run(cmd, check=True, capture_output=True, text=True, env=e).stdout
    j = json.loads(out.strip().splitlines()[-1])
    j["wall_secs"] = time.time() - t0
    return j


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=46_000_003)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--unaligned-frac", type=float, default=0.01, help="share of junction / junk reads (splice fallback)")
    a = ap.parse_args()
    procs = os.cpu_count() or 1
    n = a.reads
    g, introns = synth.make_intron_genome(a.genome, 501, max(100, a.genome // 4000))
    with tempfile.TemporaryDirectory() as td:
        t0 = time.time()
        synth.write_fasta(os.path.join(td, "g.fa"), g)
        subprocess.run([REF, "index", "g", "g.fa"], cwd=td, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        print(json.dumps({"what": "reference builder (HSA index)", "genome_bp": a.genome, "secs": time.time() - t0}), flush=True)
        prefix = os.path.join(td, "g")
        n_j = int(n * a.unaligned_frac * 0.8)
        n_junk = int(n * a.unaligned_frac * 0.2)
        rng = np.random.default_rng(3)
        parts = [synth.simulate_reads(g, n - n_j - n_junk, 100, 1).codes.reshape(-1, 100),
                 synth.simulate_junction_reads(g, introns, n_j, 100, 2).codes.reshape(-1, 100),
                 rng.integers(0, 4, size=(n_junk, 100), dtype=np.uint8)]
        codes = np.concatenate(parts)
        codes = codes[rng.permutation(n)]
        rs = synth.ReadSet(np.full(n, 100, dtype=np.uint32), np.ascontiguousarray(codes).reshape(-1))
        rp = os.path.join(td, "r.reads")
        synth.write_reads_bin(rp, rs)
        n1 = min(n, 100_000)
        synth.write_reads_bin(os.path.join(td, "r1.reads"), rs.subset(0, n1))
        base = dict(workload=f"{a.genome} bp genome (reference-built index), {n} x 100 bp reads, {a.unaligned_frac:.1%} of them "
                             f"junction / junk reads that take the splice fallback, default options, batches of 100 000 (bwtaln.c:477)")
        j = run([REF, "driver", prefix, os.path.join(td, "r1.reads"), "x", "nout=1", "procs=1"])
        print(json.dumps(dict(base, what="stock driver, 1 thread (the reference as it ships)", reads=n1, reads_per_s=n1 / j["secs"],
                              aligned_any=j["aligned_any"], aligned_whole=j["aligned_whole"])), flush=True)
        j = run([REF, "driver", prefix, rp, "x", "nout=1", f"procs={procs}"])
        cpu_rate = n / j["secs"]
        print(json.dumps(dict(base, what=f"stock driver, {procs} forked processes", reads=n, reads_per_s=cpu_rate,
                              aligned_any=j["aligned_any"], aligned_whole=j["aligned_whole"])), flush=True)
        for splice, batch in (("1", 100_000), ("0", 100_000), ("1", 1_000_000)):
            j = run([REF_GPU, "gpudriver", prefix, rp, "x", "nout=1", f"batch={batch}"], env={"HSA_GPU_SPLICE": splice})
            print(json.dumps(dict(base, what=f"GPU shim (bwa_cal_sa_reg_gap_gpu), splice fallback on the {'GPU' if splice == '1' else 'host'}, "
                                            f"batches of {batch}", reads=n, reads_per_s=n / j["secs"], vs_stock_all_cores=n / j["secs"] / cpu_rate,
                                  aligned_any=j["aligned_any"], aligned_whole=j["aligned_whole"], driver_secs=j["secs"])), flush=True)
        # the splice path alone
        m = min(n, 400_000)
        jr = synth.simulate_junction_reads(g, introns, m, 100, 7, sub_rate=0.015)
        synth.write_reads_bin(os.path.join(td, "j.reads"), jr)
        j = run([REF, "splice", prefix, os.path.join(td, "j.reads"), "x", "nout=1", f"procs={procs}", "clear_gape=1"])
        ix = index_io.load_index(prefix)
        dev = api.Index.upload(ix, 0)
        opt = api.gap_init_opt(max_diff=api.bwa_cal_maxdiff(100), mode=2)
        off = jr.offsets[:-1].astype(np.uint64)
        dev.splice_match(jr.codes[: 100 * 1000], off[:1000], jr.lens[:1000], opt)
        t0 = time.time()
        n_aln, _ = dev.splice_match(jr.codes, off, jr.lens, opt)
        dt = time.time() - t0
        print(json.dumps({"what": "splice path alone: hsa_splice_match_batch (host buffers) vs oracle/_ref/hsa_ref splice", "reads": m,
                          "gpu_reads_per_s": m / dt, "cpu_reads_per_s": m / j["secs"], "cpu_procs": procs, "two_part": int((n_aln == 2).sum()),
                          "cpu_two_part": j["two_parts"], "occ_lookups_per_read": dev.last_splice_lookups / m}), flush=True)
        dev.close()


if __name__ == "__main__":
    main()

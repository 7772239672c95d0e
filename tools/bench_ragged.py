"""Throughput of a ragged batch (1 M reads, 33..140 bp) through hsa_whole_reads with the work order pre-binned by
(length, max_diff) and in input order (HSA_B200_PREBIN=0).  GPU box helper; one JSON line per setting.
    python tools/bench_ragged.py [--genome 46000003] [--reads 1000000]"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(genome, n):
    import numpy as np
    import torch
    from hsa_b200 import api, index_build, synth_torch
    dev = torch.device("cuda", 0)
    g = synth_torch.make_genome(genome, 1, dev)
    ix = api.Index.upload(index_build.build_index(g, device=dev, sa_interval=0), 0)
    reads = synth_torch.simulate_reads(g, n, 140, 77).cpu().numpy()
    lens = np.random.default_rng(5).integers(33, 141, size=n).astype(np.uint32)
    keep = np.arange(140)[None, :] < lens[:, None]
    codes = np.ascontiguousarray(reads[keep])
    off = np.concatenate([[0], np.cumsum(lens.astype(np.int64))[:-1]]).astype(np.uint64)
    opt = api.gap_init_opt()
    ix.whole_reads(codes, off, lens, opt)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        res = ix.whole_reads(codes, off, lens, opt, copy=False)
        dt = time.perf_counter() - t0
        best = (dt, res.kernel_ms, int((res.n_aln > 0).sum()), res.occ_lookups) if best is None or dt < best[0] else best
    print(json.dumps({"prebin": os.environ.get("HSA_B200_PREBIN", "1"), "genome_bp": genome, "reads": n, "lengths": "uniform 33..140 bp",
                      "wall_ms": best[0] * 1e3, "kernel_ms": best[1], "reads_per_s_kernel": n / (best[1] * 1e-3),
                      "aligned": best[2], "occ_lookups": best[3]}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=46_000_003)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        child(a.genome, a.reads)
    else:
        for pb in ("1", "0"):
            subprocess.run([sys.executable, __file__, "--child", "--genome", str(a.genome), "--reads", str(a.reads)],
                           env=dict(os.environ, HSA_B200_PREBIN=pb), check=True)

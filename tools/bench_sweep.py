"""Run tools/exp_tail.py-style timing under a list of environment variants (GPU box helper).

    python tools/bench_sweep.py --reads 4000000 "HSA_B200_MINB=5" "HSA_B200_MINB=6 HSA_B200_POP_BIAS=-8"
Each variant runs in a fresh process (the library reads its environment once per workspace).
"""
import os
import subprocess
import sys

CHILD = r'''
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from hsa_b200 import api, build, index_build, synth_torch
n = int(sys.argv[1]); G = int(os.environ.get("EXP_GENOME", 46000003)); L = int(os.environ.get("EXP_LEN", 100))
dev = torch.device("cuda", 0)
genome = synth_torch.make_genome(G, 1, dev)
index = api.Index.upload(index_build.build_index(genome, device=dev), 0)
reads = synth_torch.simulate_reads(genome, n, L, 1000)
codes = reads.reshape(-1).cpu().pin_memory()
off = (torch.arange(n, dtype=torch.int64) * L).pin_memory()
lens = torch.full((n,), L, dtype=torch.int32).pin_memory()
opt = api.gap_init_opt()
best = None
for rep in range(3):
    t0 = time.perf_counter()
    res = index.whole_reads(codes, off, lens, opt, copy=False)
    dt = time.perf_counter() - t0
    if best is None or res.kernel_ms < best[0]:
        best = (res.kernel_ms, dt)
print(f"kernel={best[0]:.1f}ms wall={best[1]*1e3:.1f}ms reads/s(kernel)={n/best[0]/1e3:.2f}M strict={res.n_strict} "
      f"lookups={res.occ_lookups} hits={int(res.n_aln.sum())}")
'''


def main():
    args = sys.argv[1:]
    reads = "4000000"
    if args and args[0] == "--reads":
        reads, args = args[1], args[2:]
    for variant in args or [""]:
        env = dict(os.environ)
        for kv in variant.split():
            k, v = kv.split("=", 1)
            env[k] = v
        p = subprocess.run([sys.executable, "-c", CHILD, reads], env=env, capture_output=True, text=True)
        out = p.stdout.strip().splitlines()
        print(f"[{variant or 'default'}] {out[-1] if out else 'FAILED rc=%d %s' % (p.returncode, p.stderr[-800:])}", flush=True)
        for line in p.stderr.splitlines():
            if line.startswith("[hsa_b200 trace]"):
                print("    " + line[:1400], flush=True)
                break


if __name__ == "__main__":
    main()

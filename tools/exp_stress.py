"""GPU-box experiment: the stress configuration (150 bp, n5 o2) and the default batch under env variants."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hsa_b200 import api, build, index_build, synth_torch
build.build_native()
dev = torch.device("cuda", 0)
G = int(os.environ.get("EXP_GENOME", 46_000_003))
genome = synth_torch.make_genome(G, 1, dev)
index = api.Index.upload(index_build.build_index(genome, device=dev, sa_interval=0), 0)
def run(name, n, L, opt, **kw):
    r = synth_torch.simulate_reads(genome, n, L, 21, **kw)
    codes = r.reshape(-1).cpu().pin_memory(); off = (torch.arange(n, dtype=torch.int64) * L).pin_memory(); lens = torch.full((n,), L, dtype=torch.int32).pin_memory()
    best = None
    for _ in range(3):
        res = index.whole_reads(codes, off, lens, opt, copy=False)
        best = res.kernel_ms if best is None else min(best, res.kernel_ms)
    print(f"{name}: kernel={best:.1f}ms reads/s={n / best / 1e3:.2f}M heavy={res.n_strict} hits={int(res.n_aln.sum())}", flush=True)
N_STRESS = int(os.environ.get("EXP_N", 500_000))
run(f"stress {N_STRESS} x 150bp n5o2", N_STRESS, 150, api.gap_init_opt(fnr=0.0, max_diff=5, max_gapo=2), sub_rate=0.02, indel_frac=0.10)
if not os.environ.get("EXP_N"):
    run("default 2M x 100bp", 2_000_000, 100, api.gap_init_opt())
    run("default 10M x 100bp", 10_000_000, 100, api.gap_init_opt())

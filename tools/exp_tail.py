"""GPU-box experiment: per-launch times and lane occupancy of the whole-read path at several batch sizes
(HSA_B200_TRACE=1 makes the library print one trace line per batch on stderr)."""
import os
import sys
import time

os.environ.setdefault("HSA_B200_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from hsa_b200 import api, build, index_build, synth_torch  # noqa: E402

build.build_native()
dev = torch.device("cuda", 0)
G = int(os.environ.get("EXP_GENOME", 46_000_003))
L = int(os.environ.get("EXP_LEN", 100))
sizes = [int(x) for x in (sys.argv[1:] or ["100000", "500000", "2000000", "10000000"])]
genome = synth_torch.make_genome(G, 1, dev)
index = api.Index.upload(index_build.build_index(genome, device=dev), 0)
opt = api.gap_init_opt()
for n in sizes:
    reads = synth_torch.simulate_reads(genome, n, L, 1000)
    codes = reads.reshape(-1).cpu().pin_memory()
    off = (torch.arange(n, dtype=torch.int64) * L).pin_memory()
    lens = torch.full((n,), L, dtype=torch.int32).pin_memory()
    for rep in range(2):
        t0 = time.perf_counter()
        res = index.whole_reads(codes, off, lens, opt, copy=False)
        dt = time.perf_counter() - t0
        print(f"n={n} rep={rep} wall={dt * 1e3:.1f}ms kernel={res.kernel_ms:.1f}ms steps={res.steps} pops={res.pops} "
              f"lookups={res.occ_lookups} strict={res.n_strict}", flush=True)

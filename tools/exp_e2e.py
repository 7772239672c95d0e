"""GPU-box experiment: where the end-to-end (host-buffer, double-buffered) path spends its wall time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hsa_b200 import api, build, index_build, synth_torch
build.build_native()
dev = torch.device("cuda", 0)
G = int(os.environ.get("EXP_GENOME", 46_000_003)); L = 100; n = int(os.environ.get("EXP_READS", 10_000_000))
genome = synth_torch.make_genome(G, 1, dev)
index = api.Index.upload(index_build.build_index(genome, device=dev), 0)
reads = synth_torch.simulate_reads(genome, n, L, 1000)
codes = reads.reshape(-1).cpu().pin_memory()
off = (torch.arange(n, dtype=torch.int64) * L).pin_memory()
lens = torch.full((n,), L, dtype=torch.int32).pin_memory()
opt = api.gap_init_opt()
def loop(steps, verbose):
    t00 = time.perf_counter()
    job = index.whole_reads_submit(codes, off, lens, opt)
    for k in range(steps):
        t0 = time.perf_counter()
        nxt = index.whole_reads_submit(codes, off, lens, opt) if k + 1 < steps else None
        t1 = time.perf_counter()
        last = job.wait(copy=False)
        t2 = time.perf_counter()
        if verbose:
            print(f"step {k}: submit {1e3 * (t1 - t0):.1f} ms, wait {1e3 * (t2 - t1):.1f} ms, kernel_ms {last.kernel_ms:.1f}", flush=True)
        job = nxt
    torch.cuda.synchronize()
    return time.perf_counter() - t00
loop(3, False)
dt = loop(6, True)
print(f"6 steps: {dt * 1e3:.1f} ms -> {6 * n / dt / 1e6:.2f} M reads/s")

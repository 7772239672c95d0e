import os, sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
os.environ["HSA_B200_TRACE"] = "1"
from hsa_b200 import api, index_build, synth_torch
dev = torch.device("cuda", 0)
G = int(sys.argv[1])
g = synth_torch.make_genome(G, 1, dev)
introns = synth_torch.plant_introns(g, max(100, G // 150_000), 2)
ix = api.Index.upload(index_build.build_index(g, device=dev, sa_interval=0), 0)
n, L = 1_000_000, 100
reads = synth_torch.simulate_junction_reads(g, introns, n, L, 131)
codes = reads.reshape(-1).cpu().numpy(); off = np.arange(n, dtype=np.uint64) * L; lens = np.full(n, L, dtype=np.uint32)
opt = api.gap_init_opt()
for _ in range(2):
    t0 = time.perf_counter(); res = ix.whole_reads(codes, off, lens, opt); dt = time.perf_counter() - t0
    print("junction reads: wall %.1f ms kernel %.1f ms heavy %d aligned %d lookups/read %.0f" % (dt*1e3, res.kernel_ms, res.n_strict, int((res.n_aln>0).sum()), res.occ_lookups / n), flush=True)

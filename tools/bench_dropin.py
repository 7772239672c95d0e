"""The drop-in leg alone, with the shim's per-phase seconds (HSA_GPU_SHIM_TIMING=1): the reference's batch loop with
bwa_cal_sa_reg_gap_gpu on a product-written 46 Mb index, 1 % of the reads through the splice fallback.
    python tools/bench_dropin.py [--reads 3000000] [--batches 100000,262144,1000000]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=46_000_003)
    ap.add_argument("--reads", type=int, default=3_000_000)
    ap.add_argument("--batches", default="100000,1000000")
    ap.add_argument("--stock", action="store_true", help="also time the stock driver on all host cores")
    ap.add_argument("--sam", type=int, default=0, help="also run the SAM stage (sam / gpusam) on the first N reads")
    ap.add_argument("--trace", action="store_true", help="HSA_B200_TRACE=1: per-launch completion times of every batch (stderr)")
    a = ap.parse_args()
    import numpy as np
    import torch
    from hsa_b200 import index_build, index_io, synth, synth_torch
    dev = torch.device("cuda:0")
    ref, ref_gpu = os.path.join(ROOT, "oracle", "_ref", "hsa_ref"), os.path.join(ROOT, "oracle", "_ref", "hsa_ref_gpu")
    G, n, L = a.genome, a.reads, 100
    genome = synth_torch.make_genome(G, 20240611, dev)
    introns = synth_torch.plant_introns(genome, 300, 7)
    n_j = n // 100
    reads = torch.cat([synth_torch.simulate_reads(genome, n - n_j, L, 41), synth_torch.simulate_junction_reads(genome, introns, n_j, L, 42)])
    reads = reads[torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(5))].cpu().numpy()
    with tempfile.TemporaryDirectory() as td:
        index_io.save_index(index_build.build_full_index(genome, device=dev), os.path.join(td, "g"))
        del genome
        torch.cuda.empty_cache()
        rs = synth.ReadSet(np.full(n, L, dtype=np.uint32), np.ascontiguousarray(reads).reshape(-1))
        synth.write_reads_bin(os.path.join(td, "r.reads"), rs)
        env = dict(os.environ, HSA_GPU_SHIM_TIMING="1")
        if a.trace:
            env["HSA_B200_TRACE"] = "1"
        if a.stock:
            procs = os.cpu_count() or 1
            out = subprocess.run([ref, "driver", os.path.join(td, "g"), os.path.join(td, "r.reads"), "x", "nout=1", f"procs={procs}"],
                                 check=True, capture_output=True, text=True).stdout
            j = json.loads(out.strip().splitlines()[-1])
            print(json.dumps({"what": f"stock driver, {procs} processes", "reads_per_s": n / j["secs"], "aligned_any": j["aligned_any"]}), flush=True)
        if a.sam:
            n2 = min(n, a.sam)
            synth.write_reads_bin(os.path.join(td, "r2.reads"), rs.subset(0, n2))
            out = subprocess.run([ref, "sam", os.path.join(td, "g"), os.path.join(td, "r2.reads"), os.path.join(td, "c.bin"), os.path.join(td, "c.sam")],
                                 check=True, capture_output=True, text=True).stdout
            js = json.loads(out.strip().splitlines()[-1])
            print(json.dumps({"what": "reference sam (one thread)", "reads": n2, "secs_sam": js["secs_sam"], "secs_search": js["secs_search"]}), flush=True)
            for rep in range(2):
                p = subprocess.run([ref_gpu, "gpusam", os.path.join(td, "g"), os.path.join(td, "r2.reads"), os.path.join(td, "d.bin"), os.path.join(td, "d.sam")],
                                   check=True, capture_output=True, text=True, env=env)
                jg = json.loads(p.stdout.strip().splitlines()[-1])
                print(json.dumps({"what": "gpusam", "rep": rep, "reads": n2, "secs_sam": jg["secs_sam"], "secs_search": jg["secs_search"],
                                  "phases": [ln for ln in p.stderr.splitlines() if "seconds:" in ln or "host]" in ln]}), flush=True)
        for b in [int(x) for x in a.batches.split(",") if x]:
            for rep in range(2):
                p = subprocess.run([ref_gpu, "gpudriver", os.path.join(td, "g"), os.path.join(td, "r.reads"), "x", "nout=1", f"batch={b}"],
                                   check=True, capture_output=True, text=True, env=env)
                j = json.loads(p.stdout.strip().splitlines()[-1])
                phases = [ln for ln in p.stderr.splitlines() if "seconds:" in ln]
                if a.trace:
                    sys.stderr.write(p.stderr)
                print(json.dumps({"what": "gpudriver", "batch": b, "rep": rep, "reads": n, "reads_per_s": n / j["secs"], "driver_secs": j["secs"],
                                  "aligned_any": j["aligned_any"], "phases": phases[-1] if phases else None}), flush=True)


if __name__ == "__main__":
    main()

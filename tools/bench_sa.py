"""GPU-box measurement of the SA index -> text position path (hsa_sa_values_device).

One step = N random SA indices -> N text positions, indices and results resident in HBM.  Reported: positions/s, the
PsiMinus steps walked (one index sector each), achieved algorithmic GB/s (64 B per step as the reference touches a BWT
window + an occ row, + one 32-B sector for the SA sample) against the live random-sector probe, and the reference's own
BWTSaValue timed on one host core over a sample (oracle/_ref/hsa_ref sa).
    python tools/bench_sa.py [--genome 46000003] [--n 50000000]"""
import argparse, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from hsa_b200 import api, build, index_build, index_io, synth_torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=46_000_003)
    ap.add_argument("--n", type=int, default=50_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000)
    a = ap.parse_args()
    build.build_native()
    dev = torch.device("cuda", 0)
    genome = synth_torch.make_genome(a.genome, 1, dev)
    t0 = time.time()
    host = index_build.build_index(genome, device=dev)
    build_s = time.time() - t0
    ix = api.Index.upload(host, 0)
    g = torch.Generator(device=dev); g.manual_seed(7)
    idx = torch.randint(0, a.genome + 1, (a.n,), generator=g, device=dev, dtype=torch.int64).to(torch.int32)
    out = torch.zeros(a.n, dtype=torch.int32, device=dev)
    steps_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    s = torch.cuda.current_stream()
    for _ in range(3):
        ix.sa_values_device(idx.data_ptr(), a.n, out.data_ptr(), steps_dev.data_ptr(), s.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(a.steps):
        ix.sa_values_device(idx.data_ptr(), a.n, out.data_ptr(), steps_dev.data_ptr(), s.cuda_stream)
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    walked = int(steps_dev.item())
    # parity spot check against the host-buffer entry and, on a sample, against the oracle-free definition: text order
    sample = idx[:100000].cpu().numpy().astype(np.uint32)
    assert np.array_equal(ix.sa_values(sample), out[:100000].cpu().numpy().astype(np.uint32))
    idx_bytes = sum(ix.blocks(w)[1] for w in (0,))
    peak = api.random_sector_probe(0, max(idx_bytes, 1 << 20), 64)
    algo = walked * 64 + a.n * 32
    line = {"metric": "sa_positions_per_sec", "value": a.n / (ms * 1e-3), "unit": "positions/s", "n": a.n, "ms_per_step": ms,
            "genome_bp": a.genome, "psi_minus_steps": walked, "steps_per_query": walked / a.n,
            "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": algo / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": algo,
                         "unit_def": "64 B per PsiMinus step (BWT window + occ row of BWTOccValueOnSpot) + 32 B per query (SA sample)",
                         "physical_sector_gbs": (walked + a.n) * 32 / (ms * 1e-3) / 1e9,
                         "peak_kind": f"live random 32-B-sector probe over the forward index's footprint ({idx_bytes / 1e6:.0f} MB)"},
            "index_build_secs_with_sa": build_s}
    # CPU: the reference's BWTSaValue, one core, bounded sample
    ref = os.path.join(ROOT, "oracle", "_ref", "hsa_ref")
    if os.path.exists(ref) and a.cpu_sample:
        with tempfile.TemporaryDirectory() as td:
            p = os.path.join(td, "g.index")
            index_io.save_bwt(host.fwd, p + ".bwt", p + ".fmv"); index_io.save_bwt(host.rev, p + ".rev.bwt", p + ".rev.fmv")
            index_io.save_sa(host.fwd, p + ".sa")
            smp = idx[:a.cpu_sample].cpu().numpy().astype(np.uint32)
            with open(os.path.join(td, "i.bin"), "wb") as f:
                np.asarray([smp.shape[0]], dtype=np.uint32).tofile(f); smp.tofile(f)
            r = subprocess.run([ref, "sa", os.path.join(td, "g"), os.path.join(td, "i.bin"), os.path.join(td, "o.bin")],
                               capture_output=True, text=True)
            if r.returncode == 0:
                j = json.loads(r.stdout.strip().splitlines()[-1])
                got = np.fromfile(os.path.join(td, "o.bin"), dtype=np.uint32)[1:].reshape(-1, 2)[:, 0]
                line["cpu_baseline"] = {"value": smp.shape[0] / j["secs"], "unit": "positions/s", "cores": 1, "kind": "reference",
                                        "sample": f"{smp.shape[0]} of the step's SA indices, BWTSaValue in a loop",
                                        "identical_to_gpu": bool(np.array_equal(got, out[:a.cpu_sample].cpu().numpy().astype(np.uint32)))}
            else:
                line["cpu_baseline"] = {"unavailable": r.stderr[-300:]}
    print(json.dumps(line))


if __name__ == "__main__":
    main()

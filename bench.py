#!/usr/bin/env python
"""bench.py -- aligned reads/s of the B200 inexact-search path on BASELINE.json's configs[2]: GRCh38-sized (3.1 Gb)
synthetic genome, index replicated per GPU, ONE set of 100 M simulated 100 bp reads sharded over the GPUs (default
gap_opt_t), next to the reference's own CPU path on the host cores.  At N = 1 the line also carries a `secondary` block
for configs[1] (46 Mb genome, L2-resident index, 10 M reads).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--reads-total R] [--genome G]

A step = one pass of the whole-read path over the whole read set.  N > 1 is launched by torchrun (one rank per GPU):
rank 0 builds the index and broadcasts its device blocks over NCCL once; the host partitions the read set into
contiguous shards (hsa_b200.shard.shard_bounds); every rank searches its shard in batches of <= --batch reads; in the
end-to-end leg the per-rank results are gathered to rank 0 in input order.  Total work is fixed as N grows ("strong").
One JSON line is printed by rank 0.  See DESIGN.md section 6 for the definitions.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned_reads_per_sec"
UNIT = "reads/s"
ALGO_BYTES_PER_LOOKUP = 64          # BWT window sector + minor-occ sector of the REFERENCE layout
DEVICE_BYTES_PER_LOOKUP = 32        # the device layout: one 32-byte sector holds the counts and the 64 symbols
GEN_BLOCK = 2_500_000               # reads are generated in blocks of this many: block b has seed READ_SEED + b
READ_SEED = 1000
GENOME_SEED = 1


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads-total", type=int, default=100_000_000, help="reads of the whole job per step (configs[2]: 100 M)")
    ap.add_argument("--batch", type=int, default=12_500_000, help="reads per device batch (one hsa_whole_reads call)")
    ap.add_argument("--genome", type=int, default=3_100_000_003, help="synthetic genome length (configs[2]: 3.1 Gb)")
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (0 = min(steps, 5), 10 with N > 1)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[1] block (N = 1 only)")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workload
def workload_name(genome: int) -> str:
    return {46_000_003: "configs[1]", 3_100_000_003: "configs[2]", 4_600_003: "configs[0]-sized genome"}.get(genome, "custom")


def gen_reads(genome_t, lo: int, hi: int, L: int):
    """Reads [lo, hi) of the job's global read set: block b = reads [b * GEN_BLOCK, (b + 1) * GEN_BLOCK) is drawn with seed
    READ_SEED + b, so every rank (and the reference arm) can make exactly its own part."""
    import torch
    from hsa_b200 import synth_torch
    parts = []
    for b in range(lo // GEN_BLOCK, (hi + GEN_BLOCK - 1) // GEN_BLOCK):
        blk = synth_torch.simulate_reads(genome_t, GEN_BLOCK, L, READ_SEED + b)
        parts.append(blk[max(lo - b * GEN_BLOCK, 0): min(hi - b * GEN_BLOCK, GEN_BLOCK)])
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)


def write_index_files(index, prefix: str):
    from hsa_b200 import index_io
    p = prefix + ".index"
    index_io.save_bwt(index.fwd, p + ".bwt", p + ".fmv")
    index_io.save_bwt(index.rev, p + ".rev.bwt", p + ".rev.fmv")


class CpuReference:
    """The reference's own CPU implementation of the whole-read path: oracle/_ref/hsa_ref `whole` = the unmodified
    bwt_cal_width / bwt_match_gap from the reference tree, looped per read exactly as the whole-read part of
    bwa_cal_sa_reg_gap does (oracle/ref_harness.c: run_whole_range; like-for-like with hsa_whole_reads, no splice
    fallback), P forked processes over contiguous shards.  Test / baseline infrastructure: never on the product path."""

    def __init__(self, index, td: str):
        self.bin = os.path.join(ROOT, "oracle", "_ref", "hsa_ref")
        if not os.path.exists(self.bin):
            raise RuntimeError("oracle/_ref/hsa_ref is missing (built by __graft_entry__.build() where the reference sources exist)")
        self.td = td
        self.prefix = os.path.join(td, "g")
        write_index_files(index, self.prefix)

    def run(self, reads_np, procs: int, name: str, with_output: bool):
        import numpy as np
        from hsa_b200 import synth
        n, L = reads_np.shape
        rs = synth.ReadSet(np.full(n, L, dtype=np.uint32), np.ascontiguousarray(reads_np).reshape(-1))
        rp = os.path.join(self.td, name + ".reads")
        if not os.path.exists(rp):
            synth.write_reads_bin(rp, rs)
        out = os.path.join(self.td, name + ".aln")
        args = [self.bin, "whole", self.prefix, rp, out, f"procs={procs}"] + ([] if with_output else ["nout=1"])
        j = json.loads(subprocess.run(args, check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1])
        dump = synth.read_aln_dump(out) if with_output else None
        return j, dump


def cpu_baseline_block(cpu: CpuReference, reads_np, procs: int, gpu_n_aln=None, gpu_rows12=None):
    """cpu_baseline + parity: the sample runs WITH output on all host cores; the same reads' GPU results are compared
    with the dump bit for bit (n_aln and the 12 words of every hit, in hit order)."""
    import numpy as np
    n = reads_np.shape[0]
    j, dump = cpu.run(reads_np, procs, "sample", with_output=gpu_n_aln is not None)
    n1 = min(n, 20_000)
    j1, _ = cpu.run(reads_np[:n1], 1, "single", with_output=False)
    per_core = n / j["secs"] / procs
    block = {"value": n / j["secs"], "unit": UNIT, "cores": procs, "kind": "reference",
             "sample": f"the first {n} reads of the job, whole-read path (bwt_cal_width + bwt_match_gap per read and strand as "
                       f"bwa_cal_sa_reg_gap drives them; no splice fallback), {procs} forked processes",
             "aligned": j["aligned"], "secs": j["secs"], "per_core": per_core,
             "single_thread": {"value": n1 / j1["secs"], "sample": f"{n1} reads, 1 process (the reference as it ships)"}}
    parity = None
    if dump is not None:
        ref_n, ref_rows = dump
        bad_n = int((ref_n != gpu_n_aln).sum())
        same_rows = ref_rows.shape == gpu_rows12.shape and bool(np.array_equal(ref_rows, gpu_rows12))
        bad_rows = 0 if same_rows else (int((ref_rows != gpu_rows12).any(axis=1).sum()) if ref_rows.shape == gpu_rows12.shape
                                        else abs(ref_rows.shape[0] - gpu_rows12.shape[0]) or 1)
        parity = {"reads": n, "hits": int(ref_rows.shape[0]), "n_aln_mismatch": bad_n, "row_mismatch": bad_rows,
                  "against": "oracle/_ref/hsa_ref whole: the unmodified reference's bwt_cal_width + bwt_match_gap on the same "
                             "index files and the same reads, n_aln and all 12 words per hit in hit order"}
    return block, parity


class _CudaView:
    """Zero-copy torch view of raw device memory (for NCCL broadcast of the index blocks)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class DeviceLeg:
    """Reads and results resident in HBM: hsa_whole_reads_device per batch of the shard, all on one stream."""

    def __init__(self, index, reads_t, batch: int):
        import torch
        from hsa_b200 import api
        self.torch, self.api = torch, api
        dev = reads_t.device
        self.n, self.L = int(reads_t.shape[0]), int(reads_t.shape[1])
        self.codes = reads_t.reshape(-1)
        self.batches = [(lo, min(lo + batch, self.n)) for lo in range(0, self.n, batch)]
        nb_max = max(hi - lo for lo, hi in self.batches)
        self.off = torch.arange(nb_max, device=dev, dtype=torch.int64) * self.L          # offsets inside a batch
        self.len = torch.full((nb_max,), self.L, dtype=torch.int32, device=dev)
        self.cap = 2 * nb_max + 1024
        self.n_aln = torch.zeros(self.n, dtype=torch.int32, device=dev)
        self.aln_off = torch.zeros(self.n, dtype=torch.int64, device=dev)
        self.aln = torch.zeros((len(self.batches), self.cap * 9), dtype=torch.int32, device=dev)
        self.stats = torch.zeros((len(self.batches), 8), dtype=torch.int64, device=dev)
        self.opt = api.gap_init_opt()
        self.ws = api.DeviceWorkspace(index)
        self.stream = torch.cuda.current_stream()

    def batch(self, b: int):
        lo, hi = self.batches[b]
        self.ws.whole_reads_device(self.codes.data_ptr() + lo * self.L, self.off.data_ptr(), self.len.data_ptr(), hi - lo,
                                   [self.L], self.opt, self.n_aln.data_ptr() + 4 * lo, self.aln_off.data_ptr() + 8 * lo,
                                   self.aln[b].data_ptr(), self.cap, self.stats[b].data_ptr(), self.stream.cuda_stream)
        return self.ws.last_launches()

    def step(self, check_every_batch: bool) -> int:
        launches = 0
        for b in range(len(self.batches)):
            launches += self.batch(b)
            if check_every_batch or b == len(self.batches) - 1:
                self.ws.check()                      # raises unless the batch was searched to the end
        return launches

    def verify_stats(self):
        st = self.stats.cpu().tolist()
        for b, s in enumerate(st):
            if s[1] > self.cap or s[4] or s[7]:
                raise RuntimeError(f"batch {b} left work undone / overflowed its buffers: {s}")
        return st

    def results_of(self, lo: int, hi: int):
        """(n_aln, rows12 in item order) of reads [lo, hi) -- must lie inside batch 0 (the parity sample)."""
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        assert hi <= self.batches[0][1]
        n_aln = self.n_aln[lo:hi].cpu().numpy()
        off = self.aln_off[lo:hi].cpu().numpy().astype(np.int64)
        arena = self.aln[0].cpu().numpy().view(np.uint32).reshape(-1, 9)
        nz = np.nonzero(n_aln)[0]
        cnt = n_aln[nz].astype(np.int64)
        idx = np.repeat(off[nz] - np.concatenate([[0], np.cumsum(cnt)[:-1]]), cnt) + np.arange(int(cnt.sum()))
        a = arena[idx]
        r = np.zeros((a.shape[0], 12), dtype=np.uint32)          # the 12-word dump layout of oracle/ref_harness.c
        r[:, 0], r[:, 1], r[:, 2] = a[:, 0] & 0xFFFF, (a[:, 0] >> 16) & 0xFF, (a[:, 0] >> 24) & 0xFF
        r[:, 3:7] = a[:, 1:5]
        r[:, 7], r[:, 8] = a[:, 5] & 0x3FFFFFFF, a[:, 5] >> 30
        r[:, 9:12] = a[:, 6:9]
        return n_aln, r


def timed_device_steps(leg: DeviceLeg, steps: int, warmup: int, sync_all, world: int, dev, sampler=None):
    import torch
    import torch.distributed as dist
    for _ in range(warmup):
        leg.step(check_every_batch=True)
    sync_all()
    if sampler is not None:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record(leg.stream)
    for _ in range(steps):
        launches += leg.step(check_every_batch=False)
    e1.record(leg.stream)
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    leg.verify_stats()
    return float(ms.item()), launches


def kernel_config(leg):
    """Launch configuration of the per-lane search kernel for this batch size and option set (reporting only)."""
    try:
        return leg.ws.last_config()
    except Exception as e:                                     # an older library: the line stays valid without it
        return {"unavailable": str(e)[:80]}


def dominant_kernel(leg: DeviceLeg, peak_gbs: float, bound: str):
    """The per-lane search kernel on its own: one more (untimed) batch with an event behind every launch."""
    leg.ws.launch_timing(True)
    leg.batch(0)
    launch_ms, fast_lookups = leg.ws.launch_times()
    leg.ws.launch_timing(False)
    step_ms = sum(t for _, t in launch_ms) or 1e-9
    search_ms = sum(t for nm, t in launch_ms if nm in ("search1", "search2"))
    n_search = sum(1 for nm, _ in launch_ms if nm in ("search1", "search2")) or 1
    dom_bytes = fast_lookups * ALGO_BYTES_PER_LOOKUP
    achieved = dom_bytes / (search_ms * 1e-3) / 1e9 if search_ms > 0 else 0.0
    phys = fast_lookups * DEVICE_BYTES_PER_LOOKUP / (search_ms * 1e-3) / 1e9 if search_ms > 0 else 0.0
    lo, hi = leg.batches[0]
    return {"bound": bound, "achieved": phys, "peak": peak_gbs, "unit": "GB/s", "frac": phys / peak_gbs,
            "accounting": "achieved = occ lookups x 32 B / kernel time: the device layout answers one occ lookup from ONE 32-byte "
                          "sector (counts + 64 symbols), so 32 B per lookup is what the path as built must move (DESIGN.md section 3); "
                          "reference_layout below is the same with SURVEY.md 8d's 64 B (the reference's two sectors per lookup)",
            "reference_layout": {"bytes_per_lookup": ALGO_BYTES_PER_LOOKUP, "achieved": achieved, "frac": achieved / peak_gbs,
                                 "note": "exceeds 1 where the index is HBM-bound: the one-sector layout, not the kernel, beats the "
                                         "reference's two-sector roofline"},
            "physical": {"bytes_per_lookup": DEVICE_BYTES_PER_LOOKUP, "achieved": phys, "frac": phys / peak_gbs},
            "kernel": "search_kernel: the per-lane search kernel, pass-1 + pass-2 launches of one batch",
            "kernel_config": kernel_config(leg),
            "batch_reads": hi - lo, "kernel_launches_per_batch": n_search, "kernel_ms_per_launch": search_ms / n_search,
            "kernel_share_of_step": search_ms / step_ms,
            "algorithmic_bytes_per_launch": fast_lookups * DEVICE_BYTES_PER_LOOKUP / n_search, "algorithmic_bytes_per_lookup": DEVICE_BYTES_PER_LOOKUP,
            "kernel_occ_lookups_per_batch": fast_lookups,
            "launch_ms": [[nm, round(t, 3)] for nm, t in launch_ms]}


def ncu_traffic(genome: int, batch_reads: int):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of the same workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        for e in reversed(tj if isinstance(tj, list) else [tj]):          # the latest capture of a workload wins
            if int(e.get("genome_bp", 0)) == genome and int(e.get("reads", 0)) == batch_reads:
                return e.get("dram_bytes_per_launch"), e.get("source"), e
    except Exception:
        pass
    return None, None, None


def secondary_configs1(local_rank: int, dev, peaks: dict, args):
    """configs[1]: 46 Mb genome (both directions L2-resident), 10 M reads, device-resident, with its own roofline."""
    import torch
    from hsa_b200 import api, index_build, synth_torch
    G, n, L = 46_000_003, 10_000_000, args.read_len
    genome = synth_torch.make_genome(G, GENOME_SEED, dev)
    host_index = index_build.build_index(genome, device=dev, sa_interval=0)
    index = api.Index.upload(host_index, local_rank)
    reads = synth_torch.simulate_reads(genome, n, L, READ_SEED)
    leg = DeviceLeg(index, reads, n)
    ms, launches = timed_device_steps(leg, 5, 3, torch.cuda.synchronize, 1, dev)
    idx_bytes = sum(index.blocks(w)[1] for w in (0, 1))
    variants = api.random_sector_probe_ex(local_rank, idx_bytes, 64) if not args.no_probe else None
    peak = max(variants.values()) if variants else peaks.get("hbm_gbs", 6650.0)
    roof = dominant_kernel(leg, peak, "l2")
    roof["random_sector_probe"] = {"footprint_index_bytes": idx_bytes, "variants_gbs": variants}
    roof["peak_kind"] = f"measured live: random 32-byte-sector loads over {idx_bytes / 1e6:.1f} MB (the uploaded index, L2-resident)"
    tr, src, _ = ncu_traffic(G, n)
    roof["traffic"], roof["traffic_source"] = tr, src
    st = leg.stats.cpu().tolist()[0]
    out = {"workload": f"configs[1]: {G} bp i.i.d. synthetic genome, {n} simulated {L} bp reads, default gap_opt_t, 1 GPU",
           "value": n * 5 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / 5, "steps": 5, "warmup": 3, "gpu_launches": launches,
           "aligned_fraction": float((leg.n_aln > 0).sum().item()) / n, "occ_lookups_per_read": st[2] / n,
           "heavy_searches_handed_to_cooperative_kernel": st[3], "roofline": roof}
    leg.ws.close()
    index.close()
    return out


def dropin_block(dev, args):
    """The drop-in as HSA would use it (N = 1): the reference's own batch loop (bwtaln.c:477, 506; 100 000-read batches)
    with bwa_cal_sa_reg_gap replaced by shim/hsa_gpu_shim.c's bwa_cal_sa_reg_gap_gpu -- whole-read searches and the
    spliced-read fallback on the GPU -- next to the stock driver on the host cores.  oracle/_ref/hsa_ref_gpu is the
    unmodified reference linked with the shim; the index files are written by the product (index_io.save_index)."""
    import numpy as np
    import torch
    from hsa_b200 import index_build, index_io, synth, synth_torch
    ref, ref_gpu = os.path.join(ROOT, "oracle", "_ref", "hsa_ref"), os.path.join(ROOT, "oracle", "_ref", "hsa_ref_gpu")
    if not (os.path.exists(ref) and os.path.exists(ref_gpu)):
        return {"unavailable": "oracle/_ref/hsa_ref[_gpu] not built"}
    G, n, L = 46_000_003, 3_000_000, args.read_len         # 30 driver batches: the first one's allocations are amortised
    procs = os.cpu_count() or 1
    genome = synth_torch.make_genome(G, GENOME_SEED, dev)
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
        n1 = 50_000
        synth.write_reads_bin(os.path.join(td, "r1.reads"), rs.subset(0, n1))

        def run(cmd):
            out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
            return json.loads(out.strip().splitlines()[-1])
        j1 = run([ref, "driver", os.path.join(td, "g"), os.path.join(td, "r1.reads"), "x", "nout=1", "procs=1"])
        jp = run([ref, "driver", os.path.join(td, "g"), os.path.join(td, "r.reads"), "x", "nout=1", f"procs={procs}"])
        def best_of_two(cmd, key):
            # a GPU process that starts right after another one released gigabytes of device memory can stall for seconds in
            # its first allocations (seen in both orders, tools/bench_dropin.py): two runs, the better one counts
            a, b = run(cmd), run(cmd)
            return a if a[key] <= b[key] else b
        jg = best_of_two([ref_gpu, "gpudriver", os.path.join(td, "g"), os.path.join(td, "r.reads"), "x", "nout=1"], "secs")
        jg1m = best_of_two([ref_gpu, "gpudriver", os.path.join(td, "g"), os.path.join(td, "r.reads"), "x", "nout=1", "batch=1000000"], "secs")
        # the stage after the search (generate_sam_se_core, bwtse.c:884: selection, positions, CIGAR, MD, printing to a file)
        n2 = 300_000
        synth.write_reads_bin(os.path.join(td, "r2.reads"), rs.subset(0, n2))
        js = run([ref, "sam", os.path.join(td, "g"), os.path.join(td, "r2.reads"), os.path.join(td, "c.bin"), os.path.join(td, "c.sam")])
        jsg = best_of_two([ref_gpu, "gpusam", os.path.join(td, "g"), os.path.join(td, "r2.reads"), os.path.join(td, "d.bin"), os.path.join(td, "d.sam")], "secs_sam")
        with open(os.path.join(td, "c.bin"), "rb") as fa, open(os.path.join(td, "d.bin"), "rb") as fb:
            sam_identical = fa.read() == fb.read()
    return {"workload": f"{G} bp genome with planted introns, {n} x {L} bp reads, 1 % of them across introns (splice fallback), default "
                        f"options, the reference's 100 000-read batches; index files written by the product, loaded by the reference",
            "value": n / jg["secs"], "unit": UNIT, "aligned_any": jg["aligned_any"], "runs": "best of two processes",
            "path": "oracle/_ref/hsa_ref_gpu gpudriver: the unmodified reference program with bwa_cal_sa_reg_gap_gpu (shim/hsa_gpu_shim.c)",
            "with_1M_read_batches": {"value": n / jg1m["secs"], "aligned_any": jg1m["aligned_any"],
                                     "note": "the same program with the batch constant of bwtaln.c:477 raised from 100 000 to 1 000 000"},
            "sam_stage": {"reads": n2, "gpu_reads_per_s": n2 / jsg["secs_sam"], "reference_single_thread_reads_per_s": n2 / js["secs_sam"],
                          "fields_identical": sam_identical,
                          "whole_program_reads_per_s": {"gpu": n2 / (jsg["secs_sam"] + jsg["secs_search"]),
                                                        "reference_single_thread": n2 / (js["secs_sam"] + js["secs_search"])},
                          "path": "generate_sam_se_core_gpu (hsa_sam_se_batch, then the SAM text formatted on the shim's helper threads and "
                                  "written to a file) vs generate_sam_se_core, inside the reference's batch loop; whole_program = "
                                  "bwa_cal_sa_reg_gap[_gpu] + generate_sam_se_core[_gpu] per batch, as bwa_aln_core runs them; best of two GPU runs"},
            "stock_driver_all_cores": {"value": n / jp["secs"], "cores": procs, "aligned_any": jp["aligned_any"]},
            "stock_driver_single_thread": {"value": n1 / j1["secs"], "sample": f"{n1} reads (the reference as it ships)"}}


# ------------------------------------------------------------------------------------------------ main
def main():
    args = parse_args()
    rank, local_rank, world = dist_env()
    import numpy as np

    L = args.read_len
    cfg_name = workload_name(args.genome)
    workload = (f"{cfg_name}: {args.genome} bp i.i.d. synthetic genome, index replicated per GPU, one set of {args.reads_total} "
                f"simulated {L} bp reads sharded over the GPUs, default gap_opt_t")
    config = {"workload": workload, "reads_total_per_step": args.reads_total, "batch_reads": args.batch, "genome_bp": args.genome,
              "read_len": L, "options": "gap_init_opt defaults (fnr 0.04 -> max_diff 5 @100bp, max_gapo 1)",
              "parallelism": f"one read set, contiguous host shards x{world} (shard_bounds), index replicated (NCCL broadcast)",
              "l2_policy": "inputs larger than L2 (per batch: 1.25 GB of reads, 6.9 GB of per-item rows, ~2 GB of per-worker stack "
                           "arenas rewritten; the 3.1 GB index itself is 25x the L2); no explicit flush",
              "step": "whole-read part of bwa_cal_sa_reg_gap (both strand passes) over the whole read set; reads without a hit "
                      "are what the host hands to bwt_splice_match (not timed, in either arm)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        import torch
        from hsa_b200 import index_build, synth_torch
        procs = os.cpu_count() or 1
        dev = "cuda" if torch.cuda.is_available() else "cpu"     # index construction is plumbing, not the timed path
        genome = synth_torch.make_genome(args.genome, GENOME_SEED, dev)
        index = index_build.build_index(genome, device=dev, sa_interval=0)
        n_sample = args.cpu_sample or min(args.reads_total, (20_000 if args.genome > 1_000_000_000 else 40_000) * procs)
        reads = gen_reads(genome, 0, n_sample, L).cpu().numpy()       # the FIRST n_sample reads of the b200 arm's job
        del genome
        with tempfile.TemporaryDirectory() as td:
            cpu = CpuReference(index, td)
            runs = [cpu.run(reads, procs, "sample", with_output=False)[0] for _ in range(args.warmup + args.steps)]
            j1, _ = cpu.run(reads[: min(n_sample, 20_000)], 1, "single", with_output=False)
        timed = runs[args.warmup:]
        secs = sum(r["secs"] for r in timed)
        value = n_sample * len(timed) / secs
        cb = {"value": value, "unit": UNIT, "cores": procs, "kind": "reference", "per_core": value / procs,
              "sample": f"each step = the first {n_sample} reads of the job through oracle/_ref/hsa_ref whole (the unmodified reference's "
                        f"bwt_cal_width + bwt_match_gap per read and strand as bwa_cal_sa_reg_gap drives them, no splice fallback), "
                        f"{procs} forked processes", "aligned": timed[-1]["aligned"],
              "single_thread": {"value": min(n_sample, 20_000) / j1["secs"], "sample": "20000 reads, 1 process (the reference as it ships)"}}
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(timed),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": config, "sample_reads_per_step": n_sample, "cpu_baseline": cb,
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from hsa_b200 import api, build, index_build, shard, synth_torch

    build.build_native()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctl = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ctl = dist.new_group(backend="gloo")          # host-side control plane of the result gather (no GPU kernels)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- the drop-in inside the reference program (N = 1), before this process fills the GPU and pins host memory ----
    dropin, sam_stage = None, None
    if world == 1 and not args.no_secondary and not args.no_cpu_baseline:
        dropin = dropin_block(dev, args)
        torch.cuda.empty_cache()
        # the stage after the search on its own: tools/bench_sam.py, 46 Mb genome, 2 M reads
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_sam
            sam_stage = bench_sam.measure(argparse.Namespace(genome=46_000_003, n=2_000_000, steps=3, sample=100_000))
        except Exception as e:                                  # a side block must never take the headline measurement down
            sam_stage = {"unavailable": repr(e)[:300]}
        torch.cuda.empty_cache()

    # ---- index: built once on rank 0 (torch ops on the GPU), broadcast as device blocks over NCCL ----
    t_idx = time.time()
    if rank == 0:
        genome = synth_torch.make_genome(args.genome, GENOME_SEED, dev)
        host_index = index_build.build_index(genome, device=dev, sa_interval=0)
        idx0 = api.Index.upload(host_index, local_rank)
        metas = [idx0.meta(0), idx0.meta(1)]
    else:
        genome, host_index, idx0, metas = None, None, None, [[0] * 7, [0] * 7]
    index_build_secs = time.time() - t_idx
    bcast = None
    if world > 1:
        mt = torch.tensor(metas, dtype=torch.int64, device=dev)
        dist.broadcast(mt, 0)
        metas = mt.cpu().tolist()
        sync_all()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        blocks, nbytes_total = [], 0
        for which in (0, 1):
            nbytes = (metas[which][0] // 64 + 1) * 32
            nbytes_total += nbytes
            if rank == 0:
                ptr, nb = idx0.blocks(which)
                assert nb == nbytes
                t = torch.as_tensor(_CudaView(ptr, nb), device=dev)
            else:
                t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            dist.broadcast(t, 0)
            blocks.append(t)
        b1.record()
        torch.cuda.synchronize()
        bms = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device=dev)
        dist.all_reduce(bms, op=dist.ReduceOp.MAX)
        bcast = {"index_broadcast_ms": float(bms.item()), "index_broadcast_bytes": nbytes_total,
                 "index_broadcast_gbs": nbytes_total / (float(bms.item()) * 1e-3) / 1e9}
        index = idx0 if rank == 0 else api.Index.from_blocks(metas[0], metas[1], blocks[0].data_ptr(), blocks[1].data_ptr(), local_rank)
        # synthetic reads are drawn from the genome: every rank needs it to make its shard (not part of the search path)
        if rank != 0:
            genome = torch.empty(args.genome, dtype=torch.uint8, device=dev)
        dist.broadcast(genome, 0)
    else:
        index = idx0

    # ---- this rank's contiguous shard of the one read set (host partition) -----------------------------
    lo, hi = shard.shard_bounds(args.reads_total, world, align=min(GEN_BLOCK, max(1, args.reads_total // world)))[rank]
    reads = gen_reads(genome, lo, hi, L)
    n_local = hi - lo
    if rank != 0:
        del genome
    leg = DeviceLeg(index, reads, args.batch)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, launches = timed_device_steps(leg, args.steps, args.warmup, sync_all, world, dev, sampler)
    stats = leg.stats.cpu().numpy()
    aligned = int((leg.n_aln > 0).sum().item())
    value = args.reads_total * args.steps / (ms_total * 1e-3)
    ms_per_step = ms_total / args.steps

    # ---- end to end through the host-buffer C ABI (pinned host memory in, pinned host results out) ---------
    # hsa_whole_reads_submit / hsa_job_wait per batch, double-buffered: while batch i runs, the H2D copy of batch i+1 and
    # the D2H copy of batch i-1 use the copy engines.  Every batch moves its own inputs host->device and its own results
    # device->host inside the timed region; with N > 1 every batch's results are also gathered to rank 0 (input order).
    codes_host = reads.reshape(-1).cpu().pin_memory()
    off_host = leg.off.cpu().pin_memory()
    len_host = leg.len.cpu().pin_memory()
    opt = leg.opt
    e2e_steps = args.e2e_steps or min(args.steps, 5 if world == 1 else 10)     # N = 1: a step is the whole 100 M-read set (5 s)
    gather_bytes = [0]
    nb_max = max(bhi - blo for blo, bhi in leg.batches)
    gatherer = shard.HostGather(nb_max, 2 * nb_max + 1024, ctl=ctl) if world > 1 else None

    def submit(b):
        blo, bhi = leg.batches[b]
        return index.whole_reads_submit(codes_host[blo * L: bhi * L], off_host[: bhi - blo], len_host[: bhi - blo], opt)

    def e2e_loop(steps):
        jobs = [b for _ in range(steps) for b in range(len(leg.batches))]
        n_launch, d2h, last0 = 0, 0, None
        job = submit(jobs[0])
        for k, b in enumerate(jobs):
            nxt = submit(jobs[k + 1]) if k + 1 < len(jobs) else None
            res = job.wait(copy=False)
            n_launch += res.kernel_launches
            d2h += res.n_aln.shape[0] * 12 + int(res.aln.shape[0]) * 36
            if b == 0:
                last0 = (res.n_aln.copy(), int(res.occ_lookups), int(res.n_strict))
            if world > 1:
                # results of batch b of every rank -> rank 0, input order, through host shared memory (shard.HostGather)
                got = gatherer.gather(res.n_aln, res.aln_off, res.aln)
                if rank == 0:
                    gather_bytes[0] = int(gatherer.counts[:, 0].sum()) * 12 + int(gatherer.counts[:, 1].sum()) * 36
                    assert got[0].shape[0] == int(gatherer.counts[:, 0].sum())
                    gatherer.release()
            job = nxt
        torch.cuda.synchronize()
        return n_launch, d2h, last0

    e2e_loop(1)                        # warm-up: job slots' device buffers, the round-robin pinned result buffers, gather buffers
    sync_all()
    t0 = time.perf_counter()
    e2e_launches, d2h_total, last0 = e2e_loop(e2e_steps)
    e2e_secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_secs, op=dist.ReduceOp.MAX)
    e2e_value = args.reads_total * e2e_steps / float(e2e_secs.item())
    h2d = n_local * (L + 8 + 4)
    d2h = d2h_total // e2e_steps
    clocks = sampler.stop() if rank == 0 else None

    # the two entry points agree on batch 0 (n_aln of every read); a difference is a bug, not a statistic
    blo, bhi = leg.batches[0]
    if not np.array_equal(leg.n_aln[blo:bhi].cpu().numpy(), last0[0]):
        raise RuntimeError("hsa_whole_reads_device and hsa_whole_reads disagree on batch 0")

    tot = torch.tensor([aligned, int(stats[:, 2].sum()), int(stats[:, 3].sum()), h2d, d2h], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    aligned_all, lookups_all, heavy_all, h2d_all, d2h_all = [int(x) for x in tot.cpu().tolist()]

    if gatherer is not None:
        gatherer.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    idx_bytes = sum(index.blocks(w)[1] for w in (0, 1))
    fwd_bytes = index.blocks(0)[1]
    probe = None
    if not args.no_probe:
        # the dominant kernel (search_kernel) walks the forward direction's blocks only (hsa_core.cuh: lookup_a reads P.ix.fwd;
        # the reverse direction belongs to the width pass), so its roofline is the random-sector rate over THAT footprint
        variants = api.random_sector_probe_ex(local_rank, max(fwd_bytes, 1 << 20), 64)
        both = api.random_sector_probe_ex(local_rank, max(idx_bytes, 1 << 20), 64)
        probe = {"footprint_index_gbs": max(variants.values()), "footprint_index_bytes": fwd_bytes, "variants_gbs": variants,
                 "both_directions": {"bytes": idx_bytes, "gbs": max(both.values()), "variants_gbs": both}}
    peak = probe["footprint_index_gbs"] if probe else peaks.get("hbm_gbs", 6650.0)
    l2_resident = idx_bytes < 100e6
    roofline = dominant_kernel(leg, peak, "l2" if l2_resident else "hbm")
    tr, src, _ = ncu_traffic(args.genome, leg.batches[0][1] - leg.batches[0][0])
    roofline.update({
        "traffic": tr, "traffic_source": src,
        "whole_step": {"achieved": lookups_all * DEVICE_BYTES_PER_LOOKUP / (ms_per_step * 1e-3) / 1e9,
                       "frac": lookups_all * DEVICE_BYTES_PER_LOOKUP / (ms_per_step * 1e-3) / 1e9 / (peak * world),
                       "reference_layout_frac": lookups_all * ALGO_BYTES_PER_LOOKUP / (ms_per_step * 1e-3) / 1e9 / (peak * world),
                       "occ_lookups_per_step": lookups_all, "lookups_per_read": lookups_all / args.reads_total,
                       "ms_per_step": ms_per_step,
                       "note": "all kernels of the step (width + search + cooperative stage) over the timed region, all GPUs"},
        "peak_kind": ("measured live: best of six random 32-byte-sector probe shapes over the footprint the kernel walks, the "
                      f"forward direction's blocks ({fwd_bytes / 1e6:.1f} MB, {'L2-resident' if l2_resident else 'HBM-resident, 12x the 126 MB L2'}; "
                      f"both directions together: {idx_bytes / 1e6:.1f} MB)"
                      if probe else "MEASURED_PEAKS.json streaming copy"),
        "hbm_stream_peak_gbs": peaks.get("hbm_gbs"), "random_sector_probe": probe,
        "deduplicated": (None if l2_resident else
                         {"index_sectors_per_lookup": 0.79, "pairs_sharing_one_sector": 0.43,
                          "source": "ncu, profiles/r02_search_3g_summary.txt: 8.0 G index sectors delivered for 10.16 G lookups -- the two "
                                    "lookups of a pair (k and l + 1) fall into one 64-symbol block for 43 % of the pairs (SURVEY.md 8d's "
                                    "deduplicated figure); frac counts every lookup as one sector, so the sectors that really cross L2 are "
                                    "0.79 x achieved"}),
        "limit": (None if l2_resident else
                  "what caps the probe (and the kernel) past L2 is address translation, not DRAM: every SM's TLB reaches 128 x 2 MB "
                  "pages; the random-sector rate is a function of the PAGES touched, not of the bytes (96 MB spread over 6 GiB: 37 G "
                  "sectors/s, L2-resident; the same pages private to each SM: 264 G/s), and DRAM itself delivers 50 G sectors/s "
                  "(128 B fetched per 32 B missed = the streaming bandwidth) -- profiles/r02_probe_pages[23]?.jsonl, DESIGN.md section 3")})

    cpu_baseline, parity = None, None
    if not args.no_cpu_baseline and world == 1:
        procs = os.cpu_count() or 1
        n_sample = args.cpu_sample or min(leg.batches[0][1], (20_000 if args.genome > 1_000_000_000 else 40_000) * procs)
        gpu_n, gpu_rows = leg.results_of(0, n_sample)
        with tempfile.TemporaryDirectory() as td:
            cpu_baseline, parity = cpu_baseline_block(CpuReference(host_index, td), reads[:n_sample].cpu().numpy(), procs, gpu_n, gpu_rows)
        if parity["n_aln_mismatch"] or parity["row_mismatch"]:
            print(json.dumps({"error": "GPU results differ from the reference", "parity": parity}), file=sys.stderr)
            raise RuntimeError("parity failure against oracle/_ref/hsa_ref")

    secondary = None
    if world == 1 and not args.no_secondary and args.genome != 46_000_003:
        leg.ws.close()
        del leg, reads, codes_host
        index.close()
        torch.cuda.empty_cache()
        secondary = secondary_configs1(local_rank, dev, peaks, args)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                    "steps": e2e_steps, "gathered_to_rank0_bytes_per_batch": gather_bytes[0] if world > 1 else 0,
                    "path": "hsa_whole_reads_submit / hsa_job_wait per batch from pinned host buffers, double-buffered"
                            + ("; every batch's results gathered to rank 0 in input order through host shared memory "
                               "(shard.HostGather: no GPU kernels, so nothing queues behind the persistent search kernels)" if world > 1 else "")},
            "gpu_launches": launches, "e2e_gpu_launches": e2e_launches, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "parity": parity, "secondary": secondary, "dropin": dropin, "sam_stage": sam_stage,
            "aligned_fraction": aligned_all / args.reads_total,
            "heavy_searches_handed_to_cooperative_kernel": heavy_all,
            "index_build_secs": index_build_secs, "index_broadcast": bcast}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

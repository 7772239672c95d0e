#!/usr/bin/env python
"""bench.py -- aligned reads/s of the B200 inexact-search path on BASELINE.json's configs[1]
(46 Mb synthetic genome, 10 M simulated 100 bp reads, default gap_opt_t), next to the reference's own
CPU path on the host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--reads R] [--genome G]

N > 1 is launched by torchrun (one rank per GPU): rank 0 builds the index and broadcasts its device
blocks over NCCL once; every rank then searches its own shard of R reads (weak scaling, no collective
on the data path).  One JSON line is printed by rank 0.  See DESIGN.md section 6 for definitions.
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
ALGO_BYTES_PER_LOOKUP = 64          # BWT window sector + minor-occ sector of the reference layout


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step (configs[1]: 10 M)")
    ap.add_argument("--genome", type=int, default=46_000_003, help="synthetic genome length (configs[1]: 46 Mb)")
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-probe", action="store_true")
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


def write_index_files(index, prefix: str):
    from hsa_b200 import index_io
    p = prefix + ".index"
    index_io.save_bwt(index.fwd, p + ".bwt", p + ".fmv")
    index_io.save_bwt(index.rev, p + ".rev.bwt", p + ".rev.fmv")


def cpu_reference_run(index, reads_np, n_sample: int, procs: int) -> dict:
    """Time the reference's own CPU implementation of the whole-read path (oracle/_ref/hsa_ref `whole`:
    the unmodified bwt_cal_width / bwt_match_gap from /root/reference, P forked processes over contiguous
    shards) or, where that binary is absent, the oracle port (1 thread).  Bounded sample of the workload."""
    import numpy as np
    from hsa_b200 import synth
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "hsa_ref")
    sample = reads_np[:n_sample]
    L = sample.shape[1]
    rs = synth.ReadSet(np.full(sample.shape[0], L, dtype=np.uint32), np.ascontiguousarray(sample).reshape(-1))
    if os.path.exists(ref_bin):
        with tempfile.TemporaryDirectory() as td:
            write_index_files(index, os.path.join(td, "g"))
            synth.write_reads_bin(os.path.join(td, "r.reads"), rs)
            out = subprocess.run([ref_bin, "whole", os.path.join(td, "g"), os.path.join(td, "r.reads"), "x", "nout=1",
                                  f"procs={procs}"], check=True, capture_output=True, text=True).stdout
            j = json.loads(out.strip().splitlines()[-1])
            one = None
            if procs > 1:                    # the reference as it ships: single-threaded
                n1 = min(rs.n, 40_000)
                synth.write_reads_bin(os.path.join(td, "r1.reads"), rs.subset(0, n1))
                o1 = subprocess.run([ref_bin, "whole", os.path.join(td, "g"), os.path.join(td, "r1.reads"), "x", "nout=1",
                                     "procs=1"], check=True, capture_output=True, text=True).stdout
                j1 = json.loads(o1.strip().splitlines()[-1])
                one = {"value": n1 / j1["secs"], "sample": f"{n1} reads, 1 process"}
        return {"value": rs.n / j["secs"], "unit": UNIT, "cores": procs, "kind": "reference",
                "sample": f"{rs.n} of the step's reads, whole-read path (no splice fallback), {procs} processes",
                "aligned": j["aligned"], "secs": j["secs"], "single_thread": one}
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    o = ol.Oracle(index)
    t0 = time.time()
    n_aln, _ = o.whole(rs, ol.default_opt())
    secs = time.time() - t0
    return {"value": rs.n / secs, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{rs.n} of the step's reads, oracle port, 1 thread", "aligned": int((n_aln > 0).sum()), "secs": secs}


class _CudaView:
    """Zero-copy torch view of raw device memory (for NCCL broadcast of the index blocks)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def main():
    args = parse_args()
    rank, local_rank, world = dist_env()
    import numpy as np

    cfg_name = {46_000_003: "configs[1]", 3_100_000_003: "configs[2] (10 M-read batches)", 4_600_003: "configs[0]-sized genome"}.get(
        args.genome, "custom")
    workload = (f"{cfg_name}: {args.genome} bp i.i.d. synthetic genome, {args.reads} simulated {args.read_len} bp "
                f"reads per GPU per step, default gap_opt_t")
    config = {"workload": workload, "reads_per_gpu_per_step": args.reads, "genome_bp": args.genome,
              "read_len": args.read_len, "options": "gap_init_opt defaults (fnr 0.04 -> max_diff 5 @100bp, max_gapo 1)",
              "parallelism": f"read-sharded x{world}, index replicated", "l2_policy": "inputs larger than L2 "
              "(reads 1 GB/step + 5.5 GB of per-item rows and ~2 GB of per-worker stack arenas rewritten every step); "
              "no explicit flush", "step": "whole-read part of bwa_cal_sa_reg_gap (both strand passes); reads without a "
              "hit are what the host hands to bwt_splice_match (not timed, in either arm)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        from hsa_b200 import index_build, synth
        import torch
        procs = os.cpu_count() or 1
        # same generator / seeds as the b200 arm, bounded sample per step
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        from hsa_b200 import synth_torch
        genome = synth_torch.make_genome(args.genome, 1, dev)
        index = index_build.build_index(genome, device=dev)
        n_sample = args.cpu_sample or min(args.reads, 40_000 * procs)
        reads = synth_torch.simulate_reads(genome, n_sample, args.read_len, 1000).cpu().numpy()
        vals = []
        for _ in range(args.warmup + args.steps):
            vals.append(cpu_reference_run(index, reads, n_sample, procs))
        timed = vals[args.warmup:]
        secs = sum(v["secs"] for v in timed)
        value = n_sample * len(timed) / secs
        cb = dict(timed[-1]); cb["value"] = value
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(timed),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": dict(config, sample_reads_per_step=n_sample), "cpu_baseline": cb,
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from hsa_b200 import api, build, index_build, synth_torch

    build.build_native()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- index: built once on rank 0 (torch ops on the GPU), broadcast as device blocks over NCCL ----
    t_idx = time.time()
    if rank == 0:
        genome = synth_torch.make_genome(args.genome, 1, dev)
        host_index = index_build.build_index(genome, device=dev)
        idx0 = api.Index.upload(host_index, local_rank)
        metas = [idx0.meta(0), idx0.meta(1)]
    else:
        genome, host_index, idx0, metas = None, None, None, [[0] * 7, [0] * 7]
    if world > 1:
        mt = torch.tensor(metas, dtype=torch.int64, device=dev)
        dist.broadcast(mt, 0)
        metas = mt.cpu().tolist()
        blocks = []
        for which in (0, 1):
            nbytes = (metas[which][0] // 64 + 1) * 32
            if rank == 0:
                ptr, nb = idx0.blocks(which)
                assert nb == nbytes
                t = torch.as_tensor(_CudaView(ptr, nb), device=dev)
            else:
                t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            dist.broadcast(t, 0)
            blocks.append(t)
        index = idx0 if rank == 0 else api.Index.from_blocks(metas[0], metas[1], blocks[0].data_ptr(), blocks[1].data_ptr(), local_rank)
        # every rank generates its reads from the same genome: broadcast it too (46 MB)
        if rank != 0:
            genome = torch.empty(args.genome, dtype=torch.uint8, device=dev)
        dist.broadcast(genome, 0)
    else:
        index = idx0
    index_secs = time.time() - t_idx

    # ---- synthetic reads of this rank's shard ------------------------------------------------------------
    n, L = args.reads, args.read_len
    reads = synth_torch.simulate_reads(genome, n, L, 1000 + rank)
    codes_dev = reads.reshape(-1)
    off_dev = (torch.arange(n, device=dev, dtype=torch.int64) * L)
    len_dev = torch.full((n,), L, dtype=torch.int32, device=dev)
    opt = api.gap_init_opt()
    ws = api.DeviceWorkspace(index)
    aln_cap = 2 * n + 1024
    n_aln_dev = torch.zeros(n, dtype=torch.int32, device=dev)
    aln_off_dev = torch.zeros(n, dtype=torch.int64, device=dev)
    aln_dev = torch.zeros(aln_cap * 9, dtype=torch.int32, device=dev)
    stats_dev = torch.zeros(8, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def step_device():
        ws.whole_reads_device(codes_dev.data_ptr(), off_dev.data_ptr(), len_dev.data_ptr(), n, [L], opt,
                              n_aln_dev.data_ptr(), aln_off_dev.data_ptr(), aln_dev.data_ptr(), aln_cap,
                              stats_dev.data_ptr(), stream.cuda_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
        launches += ws.last_launches()
    e1.record(stream)
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    stats = stats_dev.cpu().tolist()         # {-, hits, lookups, heavy, bad, pops, steps, unprocessed} of the last step
    aligned = int((n_aln_dev > 0).sum().item())
    if stats[1] > aln_cap or stats[4] or stats[7] or stats[3] > n // 4:
        raise RuntimeError(f"device run left work undone / overflowed its buffers: {stats}")
    value = world * n * args.steps / (ms_total * 1e-3)
    kernel_ms = ms_total / args.steps

    # ---- end to end through the host-buffer C ABI (pinned host memory in, pinned host results out) ---------
    # The user-facing call pair hsa_whole_reads_submit / hsa_job_wait, double-buffered: while batch i runs, the
    # H2D copy of batch i+1 and the D2H copy of batch i-1 use the copy engines.  Every step moves its own inputs
    # host->device and its own results device->host inside the timed region.
    codes_host = reads.reshape(-1).cpu().pin_memory()
    off_host = off_dev.cpu().pin_memory()
    len_host = len_dev.cpu().pin_memory()
    def e2e_loop(steps):
        launches, last = 0, None
        job = index.whole_reads_submit(codes_host, off_host, len_host, opt)
        for k in range(steps):
            nxt = index.whole_reads_submit(codes_host, off_host, len_host, opt) if k + 1 < steps else None
            last = job.wait(copy=False)
            launches += last.kernel_launches
            job = nxt
        torch.cuda.synchronize()
        return launches, last

    e2e_loop(max(4, args.warmup))     # warm-up: the job slots' device buffers and the four round-robin pinned result buffers
    sync_all()
    t0 = time.perf_counter()
    e2e_launches, res = e2e_loop(args.steps)
    d2h = n * 4 + n * 8 + int(res.aln.shape[0]) * 36
    e2e_secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_secs, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / float(e2e_secs.item())
    h2d = codes_host.numel() + off_host.numel() * 8 + len_host.numel() * 4
    clocks = sampler.stop() if rank == 0 else None
    strict_reads = int(res.n_strict)

    # ---- parity spot-check inside the bench: device-resident and host-buffer runs agree ------------------
    same = bool(np.array_equal(n_aln_dev.cpu().numpy(), res.n_aln))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    lookups = int(res.occ_lookups)
    algo_bytes = lookups * ALGO_BYTES_PER_LOOKUP
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    idx_bytes = sum(index.blocks(w)[1] for w in (0, 1))
    probe = None
    if not args.no_probe:
        probe = {"footprint_index_gbs": api.random_sector_probe(local_rank, max(idx_bytes, 1 << 20), 64),
                 "footprint_8GiB_gbs": api.random_sector_probe(local_rank, 8 << 30, 64)}
    peak = probe["footprint_index_gbs"] if probe else peaks.get("hbm_gbs", 6650.0)
    # ---- the dominant kernel on its own: one more (untimed) step with an event behind every launch -----------
    ws.launch_timing(True)
    step_device()
    launch_ms, fast_lookups = ws.launch_times()
    ws.launch_timing(False)
    step_ms = sum(t for _, t in launch_ms) or 1e-9
    search_ms = sum(t for nm, t in launch_ms if nm in ("search1", "search2"))
    n_search = sum(1 for nm, _ in launch_ms if nm in ("search1", "search2")) or 1
    dom_bytes = fast_lookups * ALGO_BYTES_PER_LOOKUP
    dom_achieved = dom_bytes / (search_ms * 1e-3) / 1e9 if search_ms > 0 else 0.0
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if int(tj.get("genome_bp", 0)) == args.genome and int(tj.get("reads", 0)) == args.reads:   # same workload only
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    except Exception:
        pass
    resident = "L2-resident" if idx_bytes < 100e6 else "HBM-resident, far larger than the 126 MB L2"
    roofline = {"bound": "hbm", "achieved": dom_achieved, "peak": peak, "unit": "GB/s", "frac": dom_achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "search_kernel<128,5,u32,true>: the per-lane search kernel, pass-1 + pass-2 launches of the step",
                "kernel_launches_per_step": n_search, "kernel_ms_per_launch": search_ms / n_search,
                "kernel_share_of_step": search_ms / step_ms,
                "algorithmic_bytes_per_launch": dom_bytes / n_search, "algorithmic_bytes_per_lookup": ALGO_BYTES_PER_LOOKUP,
                "kernel_occ_lookups_per_step": fast_lookups,
                "launch_ms": [[nm, round(t, 3)] for nm, t in launch_ms],
                "whole_step": {"achieved": achieved, "frac": achieved / peak, "occ_lookups_per_step": lookups,
                               "lookups_per_read": lookups / n, "ms_per_step": kernel_ms,
                               "note": "all kernels of the step (width + search + cooperative stage) over the timed region"},
                "peak_kind": ("measured live: random 32-byte-sector loads over a footprint equal to the uploaded index "
                              f"({idx_bytes / 1e6:.1f} MB, {resident})" if probe else "MEASURED_PEAKS.json streaming copy"),
                "hbm_stream_peak_gbs": peaks.get("hbm_gbs"),
                "frac_of_hbm_stream": dom_achieved / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                "random_sector_probe": probe}
    cpu_baseline = None
    if not args.no_cpu_baseline:
        procs = os.cpu_count() or 1
        n_sample = args.cpu_sample or min(n, 40_000 * procs)
        cpu_baseline = cpu_reference_run(host_index, reads[:n_sample].cpu().numpy(), n_sample, procs)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "e2e_gpu_launches": e2e_launches, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "aligned_fraction": aligned / n, "heavy_searches_handed_to_cooperative_kernel": strict_reads,
            "device_vs_host_path_identical": same, "index_build_secs": index_secs}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

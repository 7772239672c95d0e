"""CPU: the host-side logic of bench.py that the measured numbers rest on -- the one global read set is a pure function of
(genome, read index), so every rank's contiguous shard and the reference arm's sample are slices of the SAME set."""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_read_set_is_shardable(monkeypatch):
    bench = load_bench()
    from hsa_b200 import shard, synth_torch
    monkeypatch.setattr(bench, "GEN_BLOCK", 700)
    g = synth_torch.make_genome(50003, 3, "cpu")
    total, L = 5000, 60
    whole = bench.gen_reads(g, 0, total, L)
    assert whole.shape == (total, L) and whole.dtype == torch.uint8 and int(whole.max()) <= 4
    for world in (1, 2, 3, 8):
        bounds = shard.shard_bounds(total, world, align=min(bench.GEN_BLOCK, max(1, total // world)))
        assert bounds[0][0] == 0 and bounds[-1][1] == total
        parts = [bench.gen_reads(g, lo, hi, L) for lo, hi in bounds if hi > lo]
        assert torch.equal(torch.cat(parts), whole), f"shards of world {world} are not slices of the one read set"
    # the reference arm's sample = the first reads of the job
    assert torch.equal(bench.gen_reads(g, 0, 1234, L), whole[:1234])
    # unaligned ranges too
    assert torch.equal(bench.gen_reads(g, 650, 2150, L), whole[650:2150])


def test_workload_names_and_parity_block_shape():
    bench = load_bench()
    assert bench.workload_name(3_100_000_003) == "configs[2]" and bench.workload_name(46_000_003) == "configs[1]"
    assert bench.ALGO_BYTES_PER_LOOKUP == 64 and bench.DEVICE_BYTES_PER_LOOKUP == 32

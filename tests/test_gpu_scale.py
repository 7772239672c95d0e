"""GPU (-m gpu): bit-exact parity with the UNMODIFIED reference at the genome sizes the bench is measured on.

Every test here runs the reference's own bwt_cal_width / bwt_match_gap (oracle/_ref/hsa_ref_count: the reference
objects compiled in place by oracle/Makefile, linked with our harness; forked over the box's host cores) WITH output on
the same index files and the same reads as the CUDA path, and compares n_aln, the twelve words of every hit, the hit
order and the count of occ lookups:

  * 46 Mb genome (configs[1]):  200 k x 100 bp reads, default gap_opt_t, through hsa_whole_reads AND through the
                                device-resident hsa_whole_reads_device (single- and multi-chunk, two pipes, keep_gape);
                                60 k spliced reads -> six seed searches each, 33/33/34 bp and 25 bp seeds (configs[3]);
                                24 k x 150 bp stress reads, max_diff 5, two gap opens (configs[4]);
  * 3.1 Gb genome (configs[2]): 100 k x 100 bp reads, default gap_opt_t (textLength at 72 % of 2^32, 48 M blocks);
  * the 46 Mb index the product's builder makes == the arrays `HSA index` (2BWT-Builder) writes for the same text.

The reference binaries are test infrastructure that travels to the GPU box with the snapshot; where they are missing
these tests FAIL (they never skip): a GPU run without them would prove nothing about parity.
"""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

import emu_lib as el
import oracle_lib as ol
from hsa_b200 import api, index_build, index_io, synth, synth_torch

pytestmark = pytest.mark.gpu

PROCS = os.cpu_count() or 1


def need_ref():
    assert os.path.exists(ol.REF_BIN_COUNT), (
        "oracle/_ref/hsa_ref_count is missing: build it with `make -C oracle ref` where the reference sources exist "
        "(python -c 'import __graft_entry__ as g; g.build()'); the binaries travel to the GPU box with the snapshot")


def run_ref(args):
    need_ref()
    out = subprocess.run([ol.REF_BIN_COUNT] + args, check=True, capture_output=True, text=True).stdout
    return json.loads(out.strip().splitlines()[-1])


class Scale:
    """A synthetic genome indexed on the GPU by the product's builder, written where the reference binary loads it."""

    def __init__(self, length: int, seed: int, full: bool = False, n_introns: int = 0):
        """full: everything bwa_cal_sa_reg_gap reads (SA samples, annotation, packed text) for the splice path, with
        `n_introns` motif-carrying introns planted in the text; all seven files are written in the reference's formats
        by the product's own writer (index_io.save_index)."""
        self.dev = torch.device("cuda", 0)
        self.td = tempfile.mkdtemp(prefix="hsa_scale_")
        self.genome = synth_torch.make_genome(length, seed, self.dev)
        self.introns = synth_torch.plant_introns(self.genome, n_introns, seed + 1) if n_introns else None
        self.prefix = os.path.join(self.td, "g")
        if full:
            self.index = index_build.build_full_index(self.genome, device=self.dev)
        else:
            self.index = index_build.build_index(self.genome, device=self.dev, sa_interval=0)
        index_io.save_index(self.index, self.prefix)
        self.ix = api.Index.upload(self.index, 0)

    def reads_file(self, name: str, reads_t: torch.Tensor):
        reads = reads_t.cpu().numpy()
        n, L = reads.shape
        rs = synth.ReadSet(np.full(n, L, dtype=np.uint32), np.ascontiguousarray(reads).reshape(-1))
        path = os.path.join(self.td, name + ".reads")
        synth.write_reads_bin(path, rs)
        return rs, path

    def reference(self, mode: str, reads_path: str, opt_args):
        out = reads_path + "." + mode + ".aln"
        j = run_ref([mode, self.prefix, reads_path, out, f"procs={PROCS}"] + list(opt_args))
        n_aln, rows = synth.read_aln_dump(out)
        os.remove(out)
        return n_aln, rows, j["occ4"] + j["occ1"]

    def close(self):
        self.ix.close()
        import shutil
        shutil.rmtree(self.td, ignore_errors=True)


def assert_same(res, exp_n, exp_rows, exp_lookups, what):
    bad_n = int((res.n_aln != exp_n).sum())
    assert bad_n == 0, f"{what}: n_aln differs from the reference on {bad_n} of {exp_n.shape[0]} items"
    rows = el.aln9_to_rows12(res.ordered())
    assert rows.shape == exp_rows.shape and np.array_equal(rows, exp_rows), f"{what}: hit rows differ from the reference"
    assert res.occ_lookups == exp_lookups, f"{what}: occ lookups {res.occ_lookups} != reference {exp_lookups}"


@pytest.fixture(scope="module")
def g46():
    s = Scale(46_000_003, 1)
    yield s
    s.close()


def test_whole_reads_46mb_vs_reference_binary(g46):
    """configs[1] shape: 200 k x 100 bp, defaults -- host-buffer entry point."""
    rs, path = g46.reads_file("cfg2", synth_torch.simulate_reads(g46.genome, 200_000, 100, 77))
    exp_n, exp_rows, exp_lk = g46.reference("whole", path, [])
    assert int((exp_n > 0).sum()) > 0.98 * rs.n
    res = g46.ix.whole_reads(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, api.gap_init_opt())
    assert_same(res, exp_n, exp_rows, exp_lk, "hsa_whole_reads @46 Mb")


class _DeviceRun:
    def __init__(self, ix, reads_t, opt, keep_gape=0):
        dev = reads_t.device
        n, L = reads_t.shape
        self.n, self.L = n, L
        self.codes = reads_t.reshape(-1).contiguous()
        self.off = torch.arange(n, device=dev, dtype=torch.int64) * L
        self.len = torch.full((n,), L, dtype=torch.int32, device=dev)
        self.cap = 2 * n + 1024
        self.n_aln = torch.zeros(n, dtype=torch.int32, device=dev)
        self.aln_off = torch.zeros(n, dtype=torch.int64, device=dev)
        self.aln = torch.zeros(self.cap * 9, dtype=torch.int32, device=dev)
        self.stats = torch.zeros(8, dtype=torch.int64, device=dev)
        self.ws = api.DeviceWorkspace(ix)
        self.opt, self.keep_gape = opt, keep_gape

    def run(self):
        st = torch.cuda.current_stream()
        self.ws.whole_reads_device(self.codes.data_ptr(), self.off.data_ptr(), self.len.data_ptr(), self.n, [self.L], self.opt,
                                   self.n_aln.data_ptr(), self.aln_off.data_ptr(), self.aln.data_ptr(), self.cap,
                                   self.stats.data_ptr(), st.cuda_stream, keep_gape=self.keep_gape)
        stats = self.ws.check()
        n_aln = self.n_aln.cpu().numpy()
        rows9 = self.aln.cpu().numpy().view(np.uint32).reshape(-1, 9)
        rows = el.gather_rows(n_aln, self.aln_off.cpu().numpy().astype(np.uint64), rows9)
        return n_aln, rows, stats


@pytest.mark.parametrize("env", [{}, {"HSA_B200_CHUNK": "49152", "HSA_B200_PIPES": "2"}, {"HSA_B200_MINB": "6", "HSA_B200_NB_FAST": "40"}],
                         ids=["one_chunk", "multi_chunk_two_pipes", "six_blocks_40_buckets"])
def test_device_resident_entry_point_46mb_vs_reference_binary(g46, monkeypatch, env):
    """The entry point behind bench.py's `value` (reads and results resident in HBM): same reference comparison, also
    with the batch cut into several chunks on two internal streams, and with the configuration large batches get by
    themselves (six resident blocks per SM, 40 score buckets per lane: searches that push a record of score >= 40 go
    through the cooperative kernel)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    reads_t = synth_torch.simulate_reads(g46.genome, 200_000, 100, 77)
    rs, path = g46.reads_file("cfg2dev", reads_t)
    exp_n, exp_rows, exp_lk = g46.reference("whole", path, [])
    run = _DeviceRun(g46.ix, reads_t, api.gap_init_opt())
    try:
        n_aln, rows, stats = run.run()
        assert np.array_equal(n_aln, exp_n)
        assert np.array_equal(rows, exp_rows)
        assert stats[2] == exp_lk and stats[7] == 0
    finally:
        run.ws.close()


def test_large_batch_configuration_46mb_vs_reference_binary(g46):
    """420 000 reads: above the size from which a batch runs with six blocks per SM and 40 score buckets on its own
    (hsa_b200.cu: dense_fast) -- host-buffer entry point, every hit compared with the reference binary."""
    reads_t = synth_torch.simulate_reads(g46.genome, 420_000, 100, 79)
    rs, path = g46.reads_file("cfg2dense", reads_t)
    exp_n, exp_rows, exp_lk = g46.reference("whole", path, [])
    res = g46.ix.whole_reads(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, api.gap_init_opt())
    assert_same(res, exp_n, exp_rows, exp_lk, "hsa_whole_reads @46 Mb, 420 k reads (six blocks per SM, 40 buckets)")


def test_device_resident_keep_gape_46mb_vs_reference_binary(g46):
    """keep_gape = 1 (the driver's option state after the first splice fallback of a process): gap
    extensions count against max_diff.  Reads with a 3-base deletion exercise it; reference run with clear_gape=0."""
    base = synth_torch.simulate_reads(g46.genome, 60_000, 103, 91, indel_frac=0.0)
    reads_t = torch.cat([base[:, :50], base[:, 53:]], dim=1).contiguous()       # 100 bp with a 3-base gap vs the genome
    rs, path = g46.reads_file("gape", reads_t)
    exp_n, exp_rows, exp_lk = g46.reference("whole", path, ["clear_gape=0"])
    exp_n0, _, _ = g46.reference("whole", path, [])
    assert int((exp_n != exp_n0).sum()) > 1000        # the option state matters on these reads
    run = _DeviceRun(g46.ix, reads_t, api.gap_init_opt(), keep_gape=1)
    try:
        n_aln, rows, stats = run.run()
        assert np.array_equal(n_aln, exp_n) and np.array_equal(rows, exp_rows) and stats[2] == exp_lk
    finally:
        run.ws.close()


def test_device_resident_entry_point_when_most_searches_are_heavy(g46, monkeypatch):
    """A 24-record stack arena sends most searches to the cooperative stage, far more than one round of it holds (a quarter
    of the batch): the queued rounds must cover them all, with results unchanged."""
    monkeypatch.setenv("HSA_B200_ARENA_CAP", "24")
    reads_t = synth_torch.simulate_reads(g46.genome, 200_000, 100, 77)
    rs, path = g46.reads_file("cfg2heavy", reads_t)
    exp_n, exp_rows, exp_lk = g46.reference("whole", path, [])
    run = _DeviceRun(g46.ix, reads_t, api.gap_init_opt())
    try:
        n_aln, rows, stats = run.run()
        assert stats[3] > 200_000 // 4 and stats[7] == 0          # heavy searches beyond one round, none left over
        assert np.array_equal(n_aln, exp_n) and np.array_equal(rows, exp_rows) and stats[2] == exp_lk
    finally:
        run.ws.close()


def test_device_entry_point_reports_unprocessed_searches(g46, monkeypatch):
    """Without the cooperative stage and with a tiny stack arena the fast kernel hands searches on that nothing finishes:
    hsa_workspace_check must fail loudly instead of leaving n_aln = 0 behind."""
    monkeypatch.setenv("HSA_B200_COOP", "0")
    monkeypatch.setenv("HSA_B200_ARENA_CAP", "24")
    reads_t = synth_torch.simulate_reads(g46.genome, 50_000, 100, 5)
    run = _DeviceRun(g46.ix, reads_t, api.gap_init_opt())
    try:
        with pytest.raises(api.HsaError, match="left unprocessed"):
            run.run()
        assert int(run.stats.cpu()[7]) > 0
    finally:
        run.ws.close()


@pytest.mark.parametrize("L", [100, 75], ids=["33bp_seeds", "25bp_seeds"])
def test_splice_seeds_46mb_vs_reference_binary(g46, L):
    """configs[3] shape: spliced reads -> the six seed calls of bwt_splice_match (bwtgap.c:797-820), prefix-width quirk
    included, 60 k reads = 360 k searches."""
    rs, path = g46.reads_file(f"spl{L}", synth_torch.simulate_spliced_reads(g46.genome, 60_000, L, 13 + L))
    exp_n, exp_rows, exp_lk = g46.reference("seeds", path, [])
    assert int((exp_n > 0).sum()) > 60_000
    res = g46.ix.splice_seeds(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, api.gap_init_opt())
    assert_same(res, exp_n, exp_rows, exp_lk, f"hsa_splice_seeds @46 Mb, {L // 3} bp seeds")


def test_stress_46mb_vs_reference_binary(g46):
    """configs[4] shape: 150 bp, 2 % substitutions, indels in 10 % of the reads, max_diff 5, two gap opens: the heavy
    searches (cooperative kernel, large stacks) at bench genome size."""
    rs, path = g46.reads_file("stress", synth_torch.simulate_reads(g46.genome, 24_000, 150, 21, sub_rate=0.02, indel_frac=0.10))
    args = ["fnr=0", "max_diff=5", "max_gapo=2"]
    exp_n, exp_rows, exp_lk = g46.reference("whole", path, args)
    res = g46.ix.whole_reads(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, api.gap_init_opt(fnr=0.0, max_diff=5, max_gapo=2))
    assert res.n_strict > 100                          # the cooperative stage was exercised
    assert_same(res, exp_n, exp_rows, exp_lk, "stress @46 Mb")


def test_index_46mb_equals_reference_builder(g46):
    """The search arrays the product's builder makes for the 46 Mb text == what `HSA index` (2BWT-Builder.c:215, run here
    through oracle/_ref/hsa_ref index) writes: bwtCode, both occ tables, inverseSa0, cumulative frequencies, both
    directions."""
    need_ref()
    td = g46.td
    synth.write_fasta(os.path.join(td, "ref.fa"), g46.genome.cpu().numpy())
    subprocess.run([ol.REF_BIN, "index", "ref", "ref.fa"], cwd=td, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ref = index_io.load_index(os.path.join(td, "ref"), with_sa=False)
    import common  # noqa: F401  (puts tests/golden on the path)
    import make_golden
    assert make_golden.index_digest(g46.index) == make_golden.index_digest(ref)


def test_whole_reads_3gb_vs_reference_binary():
    """configs[2]: GRCh38-sized 3.1 Gb genome (textLength at 72 % of 2^32, 48 M blocks, 64-bit block offsets), 100 k x
    100 bp reads with default options, against the reference binary on the same index files."""
    s = Scale(3_100_000_003, 1, full=True, n_introns=20_000)
    try:
        reads_t = synth_torch.simulate_reads(s.genome, 100_000, 100, 1000)
        rs, path = s.reads_file("cfg3g", reads_t)
        exp_n, exp_rows, exp_lk = s.reference("whole", path, [])
        assert int((exp_n > 0).sum()) > 0.98 * rs.n
        res = s.ix.whole_reads(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, api.gap_init_opt())
        assert_same(res, exp_n, exp_rows, exp_lk, "hsa_whole_reads @3.1 Gb")
        run = _DeviceRun(s.ix, reads_t, api.gap_init_opt())
        try:
            n_aln, rows, stats = run.run()
            assert np.array_equal(n_aln, exp_n) and np.array_equal(rows, exp_rows) and stats[2] == exp_lk
        finally:
            run.ws.close()
        # spliced reads at this size too (configs[3] names the 3.1 Gb genome): 20 k reads -> 120 k seed searches
        rs2, path2 = s.reads_file("spl3g", synth_torch.simulate_spliced_reads(s.genome, 20_000, 75, 3))
        exp_n, exp_rows, exp_lk = s.reference("seeds", path2, [])
        res = s.ix.splice_seeds(rs2.codes, rs2.offsets[:-1].astype(np.uint64), rs2.lens, api.gap_init_opt())
        assert_same(res, exp_n, exp_rows, exp_lk, "hsa_splice_seeds @3.1 Gb, 25 bp seeds")
        # the whole spliced-read fallback at this size (configs[3]): bwt_splice_match on the reference, loading the index
        # files the product wrote, against hsa_splice_match_batch: junction reads over planted introns + random introns
        jr = torch.cat([synth_torch.simulate_junction_reads(s.genome, s.introns, 16_000, 100, 5, sub_rate=0.015),
                        synth_torch.simulate_spliced_reads(s.genome, 4_000, 100, 6)])
        rs3, path3 = s.reads_file("junc3g", jr)
        out3 = path3 + ".splice.aln"
        j = run_ref(["splice", s.prefix, path3, out3, f"procs={PROCS}", "clear_gape=1"])
        exp_n, exp_rows = synth.read_aln_dump(out3)
        assert j["two_parts"] > 3_000
        ro = api.GapOpt.from_buffer_copy(bytes(el.resolve_read_opt(ol.default_opt(), 100, 1)))
        n_aln, aln = s.ix.splice_match(rs3.codes, rs3.offsets[:-1].astype(np.uint64), rs3.lens, ro)
        keep = np.arange(2)[None, :] < n_aln[:, None]
        assert np.array_equal(n_aln, exp_n), "hsa_splice_match_batch @3.1 Gb: n_aln differs from the reference"
        assert np.array_equal(el.aln9_to_rows12(aln[keep]), exp_rows), "hsa_splice_match_batch @3.1 Gb: rows differ from the reference"
    finally:
        s.close()

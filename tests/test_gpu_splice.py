"""GPU (-m gpu): the spliced-read fallback on the device (splice_kernel, hsa_splice_match_batch) against the reference's
bwt_splice_match: the committed goldens (tests/golden/golden_splice.*), and -- at a genome size where every SA / text
access leaves the caches -- the reference binary itself on the same index files (oracle/_ref/hsa_ref splice)."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import emu_lib as el
import oracle_lib as ol
import splice_common
from hsa_b200 import api, index_io, synth

pytestmark = pytest.mark.gpu

CASES = ["junction_100", "junction_100_gape_kept", "junction_75", "junction_150_n3o2", "junction_50_loggap", "random_introns",
         "unalignable_mix", "ragged"]


def to_api_opts(opts):
    return [api.GapOpt.from_buffer_copy(bytes(o)) for o in opts]


def rows_of(n_aln, aln):
    keep = np.arange(2)[None, :] < n_aln[:, None]
    return el.aln9_to_rows12(aln[keep])


@pytest.fixture(scope="module")
def gs():
    return splice_common.GoldenSplice()


@pytest.fixture(scope="module")
def dev(gs):
    ix = api.Index.upload(gs.index("cuda"), 0)
    yield ix
    ix.close()


@pytest.mark.parametrize("case", CASES)
def test_splice_match_vs_golden(gs, dev, case):
    rs = gs.reads(case)
    opts, oi = gs.opts(case, rs)
    n_aln, aln = dev.splice_match(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, to_api_opts(opts), oi)
    exp_n, exp_rows = gs.expected(case)
    assert np.array_equal(n_aln, exp_n)
    assert np.array_equal(rows_of(n_aln, aln), exp_rows)
    assert dev.last_splice_lookups > 0


def test_reads_that_outgrow_their_scratch_are_rerun(gs, dev, monkeypatch):
    """Tiny per-thread stack arenas and hit lists: the first pass flags most reads, the large-capacity pass finishes them;
    results are unchanged."""
    monkeypatch.setenv("HSA_B200_SPLICE_ARENA", "64")
    monkeypatch.setenv("HSA_B200_SPLICE_ALNS", "16")
    rs = gs.reads("junction_100")
    opts, oi = gs.opts("junction_100", rs)
    n_aln, aln = dev.splice_match(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, to_api_opts(opts), oi)
    exp_n, exp_rows = gs.expected("junction_100")
    assert np.array_equal(n_aln, exp_n) and np.array_equal(rows_of(n_aln, aln), exp_rows)


def test_bad_inputs(gs, dev, golden_index):
    rs = gs.reads("junction_100").subset(0, 8)
    opt = api.gap_init_opt(max_diff=5)
    off = rs.offsets[:-1].astype(np.uint64)
    n_aln, aln = dev.splice_match(rs.codes[:0], off[:0], rs.lens[:0], opt)
    assert n_aln.shape == (0,)
    with pytest.raises(api.HsaError, match="36 bases"):
        dev.splice_match(rs.codes, off, np.full(8, 30, dtype=np.uint32), opt)
    with pytest.raises(api.HsaError, match="opt_idx"):
        dev.splice_match(rs.codes, off, rs.lens, [opt], np.full(8, 3, dtype=np.uint32))
    bare = api.Index.upload(golden_index, 0)                       # no packed text / block list attached
    try:
        with pytest.raises(api.HsaError, match="attach"):
            bare.splice_match(rs.codes, off, rs.lens, opt)
    finally:
        bare.close()


def test_splice_match_12mb_vs_reference_binary():
    """A 12 Mb genome with 3 000 planted introns, indexed by the reference's own builder (its .bwt / .sa / .ann / .pac are
    what both sides load): 40 k junction reads + 10 k reads across random introns + 10 k diverged / junk reads through
    oracle/_ref/hsa_ref splice (all host cores) and through hsa_splice_match_batch."""
    assert os.path.exists(ol.REF_BIN), "oracle/_ref/hsa_ref is missing (built where the reference sources exist; travels with the snapshot)"
    g, introns = synth.make_intron_genome(12_000_011, 401, 3000)
    rng = np.random.default_rng(9)
    with tempfile.TemporaryDirectory() as td:
        synth.write_fasta(os.path.join(td, "g.fa"), g)
        subprocess.run([ol.REF_BIN, "index", "g", "g.fa"], cwd=td, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ix = index_io.load_index(os.path.join(td, "g"))
        dev = api.Index.upload(ix, 0)
        try:
            a = synth.simulate_junction_reads(g, introns, 40_000, 100, 1, sub_rate=0.015)
            b = synth.simulate_spliced_reads(g, 10_000, 100, 2)[0]
            c = synth.simulate_reads(g, 8_000, 100, 3, sub_rate=0.08, indel_frac=0.3)
            junk = rng.integers(0, 4, size=2_000 * 100, dtype=np.uint8)
            rs = synth.ReadSet(np.full(60_000, 100, dtype=np.uint32), np.concatenate([a.codes, b.codes, c.codes, junk]))
            rp, outp = os.path.join(td, "r.reads"), os.path.join(td, "r.aln")
            synth.write_reads_bin(rp, rs)
            for cg in (1, 0):
                opt = ol.default_opt()
                out = subprocess.run([ol.REF_BIN, "splice", os.path.join(td, "g"), rp, outp, f"procs={os.cpu_count() or 1}",
                                      f"clear_gape={cg}"] + ol.opt_args(opt), check=True, capture_output=True, text=True).stdout
                j = json.loads(out.strip().splitlines()[-1])
                exp_n, exp_rows = synth.read_aln_dump(outp)
                ro = api.GapOpt.from_buffer_copy(bytes(el.resolve_read_opt(opt, 100, cg)))
                n_aln, aln = dev.splice_match(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, ro)
                assert int((exp_n == 2).sum()) > 10_000 and j["aligned"] == int((exp_n > 0).sum())
                assert np.array_equal(n_aln, exp_n)
                assert np.array_equal(rows_of(n_aln, aln), exp_rows)
        finally:
            dev.close()

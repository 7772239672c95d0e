"""CPU: the SAM-field stage (hsa_b200/csrc/hsa_sam.cuh, compiled for the host by tests/emu) against the reference.

* goldens made by the unmodified reference (tests/golden/make_golden_sam.py: bwa_cal_sa_reg_gap + generate_sam_se_core in one
  process): hit selection with the drand48 stream, positions, mapQ, the alternative-hit lists, the pairing of spliced parts,
  CIGARs from the banded DP, MD / NM -- every field of every read;
* live against oracle/_ref/hsa_ref on fresh seeds (container only; skipped where the reference binary is absent).
"""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import common  # noqa: F401
import emu_lib as el
import oracle_lib as ol
import sam_common as sc
from hsa_b200 import index_io, synth
from splice_common import GoldenSplice

HERE = os.path.dirname(os.path.abspath(__file__))


class GoldenSam:
    def __init__(self):
        import make_golden_sam as mg
        self.mg = mg
        with open(os.path.join(HERE, "golden", "golden_sam.json")) as f:
            self.meta = json.load(f)
        self.arr = np.load(os.path.join(HERE, "golden", "golden_sam.npz"))
        self.base = GoldenSplice()
        assert self.meta["genome_digest"] == self.base.meta["genome_digest"]

    def case(self, name):
        c = self.meta["cases"][name]
        rs = self.mg.make_reads(self.base.genome, self.base.introns, c["reads"])
        assert self.mg.mgs.digest(rs.codes) == c["reads_digest"], "read generator drifted"
        want = sc.parse_ref_dump(self.arr[name + ".dump"])
        return c, rs, self.arr[name + ".n_aln"], self.arr[name + ".rows"], want


@pytest.fixture(scope="module")
def gs():
    return GoldenSam()


@pytest.fixture(scope="module")
def emu(gs):
    return el.Emu(gs.base.index())


CASES = ["dna_100", "dna_and_junctions", "dna_150_n5o2", "ragged_nocc6", "repeats_75"]


@pytest.mark.parametrize("name", CASES)
def test_sam_fields_vs_golden(gs, emu, name):
    c, rs, n_aln, rows, want = gs.case(name)
    na, off, a9 = sc.hits_input(n_aln, rows)
    got, state, counts = sc.emu_sam(emu, rs, na, off, a9, ol.default_opt(**c["opt"]), n_occ=c["n_occ"])
    bad = sc.diff(got, want)
    assert not bad, "\n".join(bad)
    assert int(counts[3]) >= c["with_cigar"]            # every CIGAR came out of the dynamic programme
    assert sum(1 for g in got if g[0]["type"]) == c["matched"]


def test_rng_stream_continues_across_batches(gs, emu):
    """Two half batches with the state handed on == one batch (generate_sam_se_core is called per 100 000-read batch and
    the drand48 stream runs on, bwtaln.c:477-514)."""
    c, rs, n_aln, rows, want = gs.case("repeats_75")
    half = rs.n // 2
    off_rows = np.concatenate([[0], np.cumsum(n_aln)])
    got, state = [], 0
    for lo, hi in ((0, half), (half, rs.n)):
        na, off, a9 = sc.hits_input(n_aln[lo:hi], rows[off_rows[lo]:off_rows[hi]])
        g, state, _ = sc.emu_sam(emu, rs.subset(lo, hi), na, off, a9, ol.default_opt(), n_occ=3, rng_state=state)
        got += g
    assert not sc.diff(got, want)


def test_empty_and_unmatched(emu, gs):
    rs = synth.ReadSet(np.asarray([100, 40], dtype=np.uint32), np.zeros(140, dtype=np.uint8))
    na, off, a9 = sc.hits_input(np.zeros(2, dtype=np.int32), np.zeros((0, 12), dtype=np.uint32))
    got, state, counts = sc.emu_sam(emu, rs, na, off, a9, ol.default_opt())
    assert all(g[0]["type"] == 0 for g in got) and state == 0 and int(counts[1]) == 0


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/hsa_ref not built (reference sources absent)")
@pytest.mark.parametrize("seed,length,okw", [(101, 100, {}), (102, 60, {}), (103, 120, dict(fnr=0.0, max_diff=4, max_gapo=2, max_gape=3))])
def test_sam_fields_vs_reference_live(gs, emu, seed, length, okw):
    genome, introns = gs.base.genome, gs.base.introns
    a = synth.simulate_reads(genome, 1500, length, seed, sub_rate=0.015, indel_frac=0.4)
    j = synth.simulate_junction_reads(genome, introns, 300, length, seed + 1, sub_rate=0.01)
    rs = synth.ReadSet(np.concatenate([a.lens, j.lens]), np.concatenate([a.codes, j.codes]))
    opt = ol.default_opt(**okw)
    with tempfile.TemporaryDirectory() as td:
        synth.write_fasta(os.path.join(td, "g.fa"), genome)
        subprocess.run([ol.REF_BIN, "index", "g", "g.fa"], cwd=td, check=True, stdout=subprocess.DEVNULL)
        rp = os.path.join(td, "r.reads")
        synth.write_reads_bin(rp, rs)
        ol.run_ref(["sam", os.path.join(td, "g"), rp, rp + ".bin", rp + ".sam"] + ol.opt_args(opt))
        ol.run_ref(["driver", os.path.join(td, "g"), rp, rp + ".aln"] + ol.opt_args(opt))
        n_aln, rows = synth.read_aln_dump(rp + ".aln")
        want = sc.parse_ref_dump(rp + ".bin")
    na, off, a9 = sc.hits_input(n_aln, rows)
    got, _, _ = sc.emu_sam(emu, rs, na, off, a9, opt)
    bad = sc.diff(got, want)
    assert not bad, "\n".join(bad)

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


def _fabricated_hits(n_reads, text_length, seed, emu):
    """Hit lists with many ties for the best score (several intervals of widths 1..3 at one score), which i.i.d.-genome searches
    rarely produce: the data-dependent part of the drand48 stream.  Intervals are drawn from SA indices whose suffixes start at
    least 200 bases before the end of the text (a real full-length hit cannot hang over the end; MD would read past it)."""
    rng = np.random.default_rng(seed)
    cand = rng.integers(1, text_length - 4, size=40 * n_reads).astype(np.uint32)
    ok = np.ones(cand.shape[0], dtype=bool)
    for d in range(3):
        ok &= emu.sa_values(cand + d)[0] < text_length - 200
    pool = cand[ok]
    n_aln = rng.integers(0, 6, size=n_reads).astype(np.int32)
    rows = []
    for c in n_aln.tolist():
        if c == 0:
            continue
        m = int(rng.integers(1, c + 1))                      # hits sharing the best score
        best = int(rng.integers(0, 4)) * 3
        for j in range(c):
            k = int(pool[int(rng.integers(0, pool.shape[0]))]); w = int(rng.integers(1, 4))
            score = best if j < m else best + 3 * int(rng.integers(1, 3))
            rows.append([score // 3, 0, 0, k, k + w - 1, 0, 0, 0, int(rng.integers(0, 2)), 0, 99 if j == 0 else 0, score])
    return n_aln, np.asarray(rows, dtype=np.uint32).reshape(-1, 12)


@pytest.mark.parametrize("seed,n_occ", [(1, 3), (2, 0), (3, 8)])
def test_cut_selection_equals_the_sequential_statement(gs, emu, seed, n_occ):
    """The selection as the kernels run it (prefix sums + LCG jumps for reads with one best hit, a single chain over the reads
    with several) against sam_select, the read-after-read statement, on hit lists full of best-score ties: same records, same
    stream state at the end."""
    import ctypes as C
    n = 3000
    rs = synth.simulate_reads(gs.base.genome, n, 100, 50 + seed)
    n_aln, rows = _fabricated_hits(n, int(gs.base.genome.shape[0]), seed, emu)
    na, off, a9 = sc.hits_input(n_aln, rows)
    out = {}
    try:
        for seq in (1, 0):
            el.lib().emu_sam_set_sequential(C.c_int(seq))
            got, state, _ = sc.emu_sam(emu, rs, na, off, a9, ol.default_opt(), n_occ=n_occ, rng_state=0x1234ABCD330E)
            out[seq] = (got, state)
        el.lib().emu_sam_dependent_reads.restype = C.c_uint32
        assert el.lib().emu_sam_dependent_reads() > 500          # the chain over reads with several best hits is really walked
    finally:
        el.lib().emu_sam_set_sequential(C.c_int(0))
    assert out[0][1] == out[1][1] and not sc.diff(out[0][0], out[1][0])
    assert sum(1 for g in out[0][0] if g[0]["type"] == 2) > 500


def test_rng48_jump_is_n_single_steps():
    """Rng48::jump (hsa_sam.cuh) against n applications of drand48's recurrence, python integers."""
    A, Cc, M = 0x5DEECE66D, 0xB, (1 << 48) - 1
    def jump(x, n):
        ca, cc, aa, ac = A, Cc, 1, 0
        while n:
            if n & 1:
                aa, ac = (aa * ca) & ((1 << 64) - 1), (ac * ca + cc) & ((1 << 64) - 1)
            cc, ca = ((ca + 1) * cc) & ((1 << 64) - 1), (ca * ca) & ((1 << 64) - 1)
            n >>= 1
        return (aa * x + ac) & M
    for x0 in (0, 0x1234ABCD330E, M):
        x = x0
        for n in range(1, 300):
            x = (x * A + Cc) & M
            assert jump(x0, n) == x
    x = 0
    for _ in range(100000):
        x = (x * A + Cc) & M
    assert jump(0, 100000) == x


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/hsa_ref not built (reference sources absent)")
def test_banded_dp_vs_aln_global_core_on_odd_geometries(tmp_path):
    """dp_global (hsa_sam.cuh) against the reference's aln_global_core + bwa_aln_path2cigar (harness mode `dp`) on window / read
    pairs the pipeline tests rarely produce: windows clipped at the end of the text (shorter than the read), windows much
    longer than the read, lengths below the band width, reads with N, unrelated sequences -- score and CIGAR identical."""
    import ctypes as C
    rng = np.random.default_rng(9)
    pairs = []
    for _ in range(600):
        len2 = int(rng.choice([1, 2, 7, 30, 49, 50, 51, 75, 100, 101, 150, 260]))
        kind = int(rng.integers(0, 5))
        len1 = max(1, len2 + int(rng.integers(-min(len2 - 1, 60), 12))) if kind < 3 else int(rng.integers(1, len2 + 70))
        read = rng.integers(0, 4, size=len2).astype(np.uint8)
        ref = rng.integers(0, 4, size=len1).astype(np.uint8)
        if kind != 4:                                        # related sequences: the window is the read with edits
            src = list(read)
            for _e in range(int(rng.integers(0, 6))):
                p = int(rng.integers(0, max(1, len(src))))
                op = int(rng.integers(0, 3))
                if op == 0 and src:
                    src[p] = int(rng.integers(0, 4))
                elif op == 1:
                    src.insert(p, int(rng.integers(0, 4)))
                elif src:
                    del src[p]
            src = (src + list(rng.integers(0, 4, size=len1)))[:len1]
            ref = np.asarray(src, dtype=np.uint8)
        if rng.random() < 0.2:
            read[rng.integers(0, len2)] = 4
        pairs.append((ref, read))
    with open(tmp_path / "p.bin", "wb") as f:
        np.asarray([len(pairs)], dtype=np.uint32).tofile(f)
        for ref, read in pairs:
            np.asarray([ref.shape[0], read.shape[0]], dtype=np.uint32).tofile(f)
            ref.tofile(f); read.tofile(f)
    ol.run_ref(["dp", str(tmp_path / "p.bin"), str(tmp_path / "o.bin")])
    w = np.fromfile(tmp_path / "o.bin", dtype=np.uint32)
    L = el.lib()
    L.emu_dp.restype = C.c_int
    L.emu_dp.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_uint32, C.POINTER(C.c_int32), C.c_void_p]
    p, narrow_refused = 1, 0
    for ref, read in pairs:
        score, n_cig = int(w[p].view(np.int32) if hasattr(w[p], "view") else w[p]), int(w[p + 1])
        want = w[p + 2:p + 2 + n_cig].tolist()
        p += 2 + n_cig
        out = np.zeros(ref.shape[0] + read.shape[0] + 4, dtype=np.uint32)
        sc_ = C.c_int32(0)
        # first with the width the library sizes for a batch without gaps beyond 8, then full width if that is refused
        n = L.emu_dp(ref.ctypes.data, ref.shape[0], read.ctypes.data, read.shape[0], min(2 * 50 + 8 + 1, ref.shape[0] + 1), C.byref(sc_), out.ctypes.data)
        if n < 0:
            narrow_refused += 1
            n = L.emu_dp(ref.ctypes.data, ref.shape[0], read.ctypes.data, read.shape[0], ref.shape[0] + 1, C.byref(sc_), out.ctypes.data)
        assert n == n_cig and out[:n].tolist() == want and (sc_.value & 0xFFFFFFFF) == (score & 0xFFFFFFFF), (ref.shape[0], read.shape[0])
    assert narrow_refused > 0        # the clipped-window guard was exercised

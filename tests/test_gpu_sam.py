"""GPU (-m gpu): the SAM-field stage through the C ABI (hsa_sam_se_batch / hsa_sam_format) against the reference.

* goldens made by the unmodified reference's generate_sam_se_core (tests/golden/golden_sam.*): every field of every read,
  and the SAM text bwa_print_sam1 printed for them (sha256 of the printable lines);
* end to end on the GPU: hsa_whole_reads + hsa_splice_match_batch -> hsa_sam_se_batch == the same goldens (the stage consumes
  the search's own output, as generate_sam_se_core consumes bwa_cal_sa_reg_gap's);
* at bench scale against the reference binary run on the box: 46 Mb genome, 120 k reads (gapped, repeated, spliced).
"""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest
import torch

import emu_lib as el
import oracle_lib as ol
import sam_common as sc
from hsa_b200 import api, synth, synth_torch
from test_sam_emu import CASES, GoldenSam, _fabricated_hits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gs():
    return GoldenSam()


@pytest.fixture(scope="module")
def index(gs):
    ix = api.Index.upload(gs.base.index(), 0)
    yield ix
    ix.close()


def _opt(okw):
    return api.gap_init_opt(**okw)


@pytest.mark.parametrize("name", CASES)
def test_sam_fields_and_text_vs_golden(gs, index, name):
    c, rs, n_aln, rows, want = gs.case(name)
    na, off, a9 = sc.hits_input(n_aln, rows)
    opt = _opt(c["opt"])
    res = index.sam_se(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, na, off, a9, opt, n_occ=c["n_occ"])
    got = sc.unpack_result(res.rec, res.multi, res.cigar, res.md)
    bad = sc.diff(got, want)
    assert not bad, "\n".join(bad)
    assert res.n_refined >= c["with_cigar"]
    opt.mode &= ~1                                          # the driver has cleared GAPE in the caller's options (bwtaln.c:261)
    text = res.format(["synth"])
    assert text.count(b"\n") == c["lines"]
    assert hashlib.sha256(sc.printable_lines(text)).hexdigest() == c["text_sha256"]


def _driver_on_gpu(index, rs, okw):
    """bwa_cal_sa_reg_gap's results for one batch from the GPU entry points, option leak included -- the logic of
    shim/hsa_gpu_shim.c in Python: whole-read pass A with the caller's options (GAPE cleared, bwtaln.c:261); from the first
    read that falls through to the splice path on, aux->opt is local_opt (copied BEFORE the clear, :254, with max_diff of the
    longest read and max_gapo clamped, :273-276), so the reads after it are searched again with that (pass B); reads without
    a hit go through bwt_splice_match with local_opt as it stands for them."""
    opt = api.gap_init_opt(**okw)
    off = rs.offsets[:-1].astype(np.uint64)
    max_len = int(rs.lens.max())
    local = api.GapOpt.from_buffer_copy(bytes(opt))
    if opt.fnr > 0:
        local.max_diff = api.bwa_cal_maxdiff(max_len, 0.02, opt.fnr)
    if local.max_diff < local.max_gapo:
        local.max_gapo = local.max_diff

    def filtered(i):
        r = rs.read(i)
        return int((r > 3).sum()) > local.max_diff or not r[:15].any() or bool((r[:15] == 3).all())

    a = index.whole_reads(rs.codes, off, rs.lens, opt)
    n_aln = a.n_aln.copy()
    rows = [el.aln9_to_rows12(a.item(i)) for i in range(rs.n)]
    first = next((i for i in range(rs.n) if n_aln[i] == 0 and not filtered(i)), None)
    if first is not None and first + 1 < rs.n:
        sub = rs.subset(first + 1, rs.n)
        b = index.whole_reads(sub.codes, sub.offsets[:-1].astype(np.uint64), sub.lens, local, keep_gape=local.mode & 1)
        for j in range(sub.n):
            n_aln[first + 1 + j] = b.n_aln[j]
            rows[first + 1 + j] = el.aln9_to_rows12(b.item(j))
    miss = [i for i in range(rs.n) if n_aln[i] == 0 and not filtered(i)]
    if miss:
        sub = synth.ReadSet(rs.lens[miss], np.concatenate([rs.read(i) for i in miss]))
        opts, oi = [], []
        for i in miss:
            o = api.GapOpt.from_buffer_copy(bytes(local))
            if i != first:                                  # written through aux->opt == &local_opt (bwtaln.c:330-332)
                L = int(rs.lens[i])
                if opt.fnr > 0:
                    o.max_diff = api.bwa_cal_maxdiff(L, 0.02, opt.fnr)
                o.seed_len = opt.seed_len if opt.seed_len < L else 0x7FFFFFFF
            key = bytes(o)
            if key not in [bytes(x) for x in opts]:
                opts.append(o)
            oi.append([bytes(x) for x in opts].index(key))
        sn, sa = index.splice_match(sub.codes, sub.offsets[:-1].astype(np.uint64), sub.lens, opts, np.asarray(oi, dtype=np.uint32))
        for j, i in enumerate(miss):
            n_aln[i] = sn[j]
            rows[i] = el.aln9_to_rows12(sa[j, :sn[j]])
    return n_aln, (np.concatenate(rows) if rows else np.zeros((0, 12), np.uint32))


@pytest.mark.parametrize("name", ["dna_and_junctions", "dna_150_n5o2"])
def test_search_then_sam_end_to_end(gs, index, name):
    """hits straight from the GPU search (whole reads, then the splice fallback for the reads that found nothing), then
    the SAM stage on them: equal to what the reference program computes from the reads alone."""
    c, rs, n_aln_ref, rows_ref, want = gs.case(name)
    n_aln, rows = _driver_on_gpu(index, rs, c["opt"])
    assert np.array_equal(n_aln, n_aln_ref) and np.array_equal(rows, rows_ref), "the GPU search's hits differ from the driver's"
    na, ao, a9 = sc.hits_input(n_aln, rows)
    out = index.sam_se(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, na, ao, a9, _opt(c["opt"]), n_occ=c["n_occ"])
    bad = sc.diff(sc.unpack_result(out.rec, out.multi, out.cigar, out.md), want)
    assert not bad, "\n".join(bad)


def test_arena_overflow_is_retried_not_truncated(gs, index, monkeypatch):
    """CIGAR / MD arenas that start far too small (HSA_B200_SAM_TINY): the batch is run again with more room until it fits;
    results are the same."""
    monkeypatch.setenv("HSA_B200_SAM_TINY", "1")
    c, rs, n_aln, rows, want = gs.case("dna_and_junctions")
    na, off, a9 = sc.hits_input(n_aln, rows)
    res = index.sam_se(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, na, off, a9, _opt(c["opt"]), n_occ=c["n_occ"])
    assert not sc.diff(sc.unpack_result(res.rec, res.multi, res.cigar, res.md), want)


@pytest.mark.parametrize("sequential", ["0", "1"], ids=["cut_selection", "sequential_statement_on_one_thread"])
def test_selection_with_many_best_score_ties(gs, index, sequential, monkeypatch):
    """Hit lists full of ties for the best score (the data-dependent part of the drand48 stream): the kernels' cut selection
    (prefix sums, LCG jumps, one chain over the reads with several best hits) and the fallback that runs the sequential
    statement on one device thread both equal the host build of the sequential statement, stream state included."""
    import ctypes as C
    monkeypatch.setenv("HSA_B200_SAM_SEQUENTIAL", sequential)
    n = 3000
    rs = synth.simulate_reads(gs.base.genome, n, 100, 51)
    emu = el.Emu(gs.base.index())
    n_aln, rows = _fabricated_hits(n, int(gs.base.genome.shape[0]), 1, emu)
    na, off, a9 = sc.hits_input(n_aln, rows)
    try:
        el.lib().emu_sam_set_sequential(C.c_int(1))
        want, want_state, _ = sc.emu_sam(emu, rs, na, off, a9, ol.default_opt(), n_occ=3, rng_state=0x1234ABCD330E)
    finally:
        el.lib().emu_sam_set_sequential(C.c_int(0))
    res = index.sam_se(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, na, off, a9, _opt({}), n_occ=3, rng48_state=0x1234ABCD330E)
    assert res.rng48_state == want_state
    bad = sc.diff(sc.unpack_result(res.rec, res.multi, res.cigar, res.md), want)
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("name", ["dna_and_junctions", "ragged_nocc6"])
def test_device_resident_entry_point(gs, index, name):
    """hsa_sam_se_device: reads and hits as device tensors (what hsa_whole_reads_device leaves in HBM), results left on the
    device -- fetched here only to compare them with the goldens."""
    c, rs, n_aln, rows, want = gs.case(name)
    na, off, a9 = sc.hits_input(n_aln, rows)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    codes = torch.cat([t(rs.codes), torch.zeros(16, dtype=torch.uint8, device=dev)])
    roff, rlen = t(rs.offsets[:-1].astype(np.int64)), t(rs.lens.astype(np.int32))
    dna, dao, da9 = t(na), t(off.astype(np.int64)), t(a9.view(np.int32))
    torch.cuda.synchronize()
    out, state = index.sam_se_device(codes.data_ptr(), roff.data_ptr(), rlen.data_ptr(), rs.n, int(rs.lens.max()), dna.data_ptr(),
                                     dao.data_ptr(), da9.data_ptr(), _opt(c["opt"]), n_occ=c["n_occ"])
    rec = index.copy_from_device(out.rec_dev, (rs.n, api.SAM_REC_WORDS), np.uint32)
    multi = index.copy_from_device(out.multi_dev, (out.n_multi, api.SAM_MULTI_WORDS), np.uint32)
    cigar = index.copy_from_device(out.cigar_dev, (out.n_cigar,), np.uint32)
    md = index.copy_from_device(out.md_dev, (out.md_bytes,), np.uint8).tobytes()
    bad = sc.diff(sc.unpack_result(rec, multi, cigar, md), want)
    assert not bad, "\n".join(bad)
    assert out.n_refined >= c["with_cigar"] and out.kernel_ms > 0


def test_window_clipped_at_the_end_of_the_text(gs, index):
    """A gapped hit whose reference window runs past the end of the text: refine_gapped_core clips the window (bwtse.c:393-394), the
    band gets wider than the batch's scratch rows, and the library runs the batch again with full-width rows.  GPU == the host
    build of the same code (which tests/test_sam_emu.py pins to aln_global_core on such geometries)."""
    emu = el.Emu(gs.base.index())
    n_text = int(gs.base.genome.shape[0])
    cand = np.arange(1, 60000, dtype=np.uint32)
    pos = emu.sa_values(cand)[0]
    ks = cand[(pos > n_text - 140) & (pos < n_text - 110)][:3]
    assert ks.shape[0] == 3
    L = 150
    rs = synth.simulate_reads(gs.base.genome, 40, L, 77)
    n_aln = np.zeros(40, dtype=np.int32); rows = []
    for i, k in zip((3, 17, 30), ks.tolist()):
        n_aln[i] = 1
        rows.append([1, 1, 0, k, k, 0, 0, 0, 0, 0, L - 1, 14])      # one mismatch, one gap open
    na, off, a9 = sc.hits_input(n_aln, np.asarray(rows, dtype=np.uint32))
    want, _, counts = sc.emu_sam(emu, rs, na, off, a9, ol.default_opt())
    res = index.sam_se(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, na, off, a9, _opt({}))
    bad = sc.diff(sc.unpack_result(res.rec, res.multi, res.cigar, res.md), want)
    assert not bad, "\n".join(bad)
    assert res.n_refined == 3 and sum(1 for g in want if g[1]) == 3


def test_rng_state_chains_batches(gs, index):
    c, rs, n_aln, rows, want = gs.case("repeats_75")
    half = rs.n // 2
    off_rows = np.concatenate([[0], np.cumsum(n_aln)])
    got, state = [], 0
    for lo, hi in ((0, half), (half, rs.n)):
        sub = rs.subset(lo, hi)
        na, off, a9 = sc.hits_input(n_aln[lo:hi], rows[off_rows[lo]:off_rows[hi]])
        r = index.sam_se(sub.codes, sub.offsets[:-1].astype(np.uint64), sub.lens, na, off, a9, _opt({}), rng48_state=state)
        state = r.rng48_state
        got += sc.unpack_result(r.rec, r.multi, r.cigar, r.md)
    assert not sc.diff(got, want)


def test_sam_46mb_vs_reference_binary():
    """Bench scale: 46 Mb genome with planted introns, 120 k reads (100 k DNA reads with indels in 20 %, 20 k junction reads),
    the reference binary (bwa_cal_sa_reg_gap + generate_sam_se_core, one process: one drand48 stream) on the index files
    the product wrote, against hsa_sam_se_batch fed with the reference driver's hits; SAM text compared line by line."""
    from test_gpu_scale import Scale, need_ref
    need_ref()
    s = Scale(46_000_003, 5, full=True, n_introns=2000)
    try:
        reads = torch.cat([synth_torch.simulate_reads(s.genome, 100_000, 100, 91, indel_frac=0.20),
                           synth_torch.simulate_junction_reads(s.genome, s.introns, 20_000, 100, 92, sub_rate=0.01)])
        reads = reads[torch.randperm(reads.shape[0], generator=torch.Generator().manual_seed(3)).to(reads.device)]
        rs, path = s.reads_file("sam", reads)
        j = json.loads(subprocess.run([ol.REF_BIN, "sam", s.prefix, path, path + ".bin", path + ".sam"], check=True, capture_output=True,
                                      text=True).stdout.strip().splitlines()[-1])
        subprocess.run([ol.REF_BIN, "driver", s.prefix, path, path + ".aln"], check=True, capture_output=True)
        n_aln, rows = synth.read_aln_dump(path + ".aln")
        want = sc.parse_ref_dump(path + ".bin")
        assert j["with_cigar"] > 15_000 and j["splicing"] > 5_000
        na, off, a9 = sc.hits_input(n_aln, rows)
        opt = api.gap_init_opt()
        res = s.ix.sam_se(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, na, off, a9, opt)
        bad = sc.diff(sc.unpack_result(res.rec, res.multi, res.cigar, res.md), want)
        assert not bad, "\n".join(bad)
        opt.mode &= ~1
        text = sc.printable_lines(res.format(["synth"]))
        ref_text = sc.printable_lines(open(path + ".sam", "rb").read())
        assert text == ref_text, "SAM text differs from the reference's"
        print(f"\n[sam @46 Mb] {rs.n} reads, {res.n_refined} refined, kernels {res.kernel_ms:.2f} ms; reference generate_sam_se_core "
              f"{j['secs_sam']:.2f} s on one core")
    finally:
        s.close()

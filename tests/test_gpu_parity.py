"""GPU (-m gpu): the compiled CUDA path, called through the C ABI (libhsa_b200.so via hsa_b200.api), against
  * the committed golden vectors produced by the unmodified reference, and
  * the oracle on fresh seeded inputs,
bit for bit: n_aln, {n_mm,n_gapo,n_gape,k,l,rev_k,rev_l,strand,start,end,score} per hit, hit order, and
the count of occ lookups the reference algorithm issues."""
import os

import numpy as np
import pytest

import emu_lib as el
import oracle_lib as ol
from hsa_b200 import api, index_build, synth

pytestmark = pytest.mark.gpu

CASES = ["cfg1_75bp_n2o1", "cfg2_100bp_default", "cfg5_150bp_n5o2", "ragged_nonstop", "ragged_loggap_gape",
         "short_entries", "exact_only", "noskip_gaps"]


def to_api_opt(o: ol.GapOpt) -> api.GapOpt:
    return api.GapOpt.from_buffer_copy(bytes(o))


@pytest.fixture(scope="module")
def dev_index(golden_index):
    ix = api.Index.upload(golden_index, 0)
    yield ix
    ix.close()


def percall_tasks(rs, opt):
    """The task list equivalent to ref_harness.c's percall mode (both strands of every read)."""
    lens = sorted(set(rs.lens.tolist()))
    opts = [el.resolve_read_opt(opt, L, 1) for L in lens]
    oi = {L: i for i, L in enumerate(lens)}
    off = rs.offsets
    t = np.zeros(rs.n * 2, dtype=api.TASK_DTYPE)
    L = rs.lens.astype(np.uint32)
    for s_i, s in enumerate((1, 0)):
        v = t[s_i::2]
        v["read_off"] = off[:-1]
        v["read_len"] = L
        v["strand"] = s
        v["len"] = L
        v["seed_mode"] = [1 if int(x) > opts[oi[int(x)]].seed_len else 0 for x in L]
        v["opt_idx"] = [oi[int(x)] for x in L]
    return t, [to_api_opt(o) for o in opts]


def test_rank_both_layouts(golden, dev_index):
    idx, occ = golden.arr["occ_idx"], golden.arr["occ"]
    for which, cols in ((0, slice(0, 4)), (1, slice(4, 8))):
        assert np.array_equal(dev_index.occ(which, idx, layout=0), occ[:, cols])
        assert np.array_equal(dev_index.occ(which, idx, layout=1), occ[:, cols])


def test_rank_exhaustive_small(golden_index, dev_index):
    """Every index of a prefix and a suffix of the SA range, both layouts == oracle."""
    n = golden_index.fwd.text_length
    idx = np.concatenate([np.arange(0, 3000), np.arange(n - 3000, n + 2)]).astype(np.uint32)
    o = ol.Oracle(golden_index)
    for which in (0, 1):
        exp, _ = o.occ(which, idx)
        assert np.array_equal(dev_index.occ(which, idx, layout=1), exp)
        assert np.array_equal(dev_index.occ(which, idx, layout=0), exp)


def test_sa_values_golden(golden, dev_index):
    """hsa_sa_values == the reference's BWTSaValue on the golden SA indices (incl. 0, textLength, inverseSa0)."""
    idx = golden.arr["sa_idx"]
    assert np.array_equal(dev_index.sa_values(idx), golden.arr["sa_val"])
    assert dev_index.last_sa_steps == int(golden.arr["sa_steps"].astype(np.int64).sum())


def test_sa_values_every_index_is_a_permutation(golden_index, dev_index):
    """All n + 1 SA indices: values == the oracle's, and together they are every text position exactly once
    (SA[0], the '$' suffix, reads -1 as the reference keeps it)."""
    n = golden_index.fwd.text_length
    idx = np.arange(0, n + 1, dtype=np.uint32)
    got = dev_index.sa_values(idx)
    exp, steps = ol.Oracle(golden_index).sa_values(idx)
    assert np.array_equal(got, exp) and dev_index.last_sa_steps == int(steps.astype(np.int64).sum())
    assert got[0] == 0xFFFFFFFF and np.array_equal(np.sort(got[1:]), np.arange(n, dtype=np.uint32))


def test_sa_values_bad_and_empty(golden_index, dev_index):
    assert dev_index.sa_values(np.zeros(0, dtype=np.uint32)).shape == (0,)
    with pytest.raises(api.HsaError):
        dev_index.sa_values(np.asarray([golden_index.fwd.text_length + 1], dtype=np.uint32))


def test_sa_locate_matches_reference(golden):
    """hsa_sa_locate == BWTRetrievePositionFromSAIndex on the multi-record golden genome (every SA index but 0), and
    SA index 0 (SA[0] = -1) reports no block."""
    from hsa_b200 import index_io
    recs, text = golden.genome2()
    index = index_build.build_index(text, device="cuda")
    index.blocks = index_io.blocks_of_records([r.shape[0] for r in recs])
    ix = api.Index.upload(index, 0)
    try:
        idx = np.arange(1, text.shape[0] + 1, dtype=np.uint32)
        assert np.array_equal(ix.sa_locate(idx), golden.arr["loc_out"])
        z = ix.sa_locate(np.zeros(1, dtype=np.uint32))
        assert z.tolist() == [[0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF]]
        with pytest.raises(api.HsaError):
            ix.attach_blocks(np.asarray([[0, 10, 5, 0]], dtype=np.uint32))        # start > end
    finally:
        ix.close()


def test_hit_intervals_to_positions(golden, golden_index, dev_index):
    """The consumer's view (bwa_cal_pac_pos bwtse.c:350-369, bwt_aln_corelate_check bwtgap.c:669-742): every SA index of
    the hit intervals of a whole-read batch -> text position; for gap-free hits the strand-resolved read lies there
    with exactly n_mm mismatches.  (Not the max_diff = 0 case: there the reference's bwt_match_exact leaves k = 0 in
    the hit, 2BWT-Interface.c:383, so its own interval is not the match's.)"""
    case = "cfg1_75bp_n2o1"
    rs = golden.reads(case)
    opt = to_api_opt(ol.default_opt(**golden.opt_kwargs(case)))
    res = dev_index.whole_reads(rs.codes, rs.offsets[:-1], rs.lens, opt)
    text = golden.genome
    checked = 0
    for r in range(rs.n):
        for h in range(int(res.n_aln[r])):
            a = res.aln[int(res.aln_off[r]) + h]
            n_mm, n_gapo, k, l, strand = int(a[0]) & 0xFFFF, (int(a[0]) >> 16) & 0xFF, int(a[1]), int(a[2]), int(a[5]) >> 30
            if n_gapo or l - k > 50:
                continue
            pos = dev_index.sa_values(np.arange(k, l + 1, dtype=np.uint32))
            read = rs.read(r)
            seq = np.where(read[::-1] < 4, 3 - read[::-1], read[::-1]) if strand else read
            for p in pos.tolist():
                assert p + seq.shape[0] <= text.shape[0]
                assert int((text[p: p + seq.shape[0]] != seq).sum()) == n_mm
                checked += 1
        if checked > 400:
            break
    assert checked > 200


def test_width(golden, dev_index):
    case = "ragged_nonstop"
    rs = golden.reads(case).subset(0, 64)
    bid, w = dev_index.cal_width(rs.codes, rs.offsets[:-1], rs.lens)
    off = rs.offsets
    for r, (b, ww) in enumerate(golden.widths(case)):
        assert bid[r] == b
        assert np.array_equal(w[off[r] + r: off[r] + r + int(rs.lens[r]) + 1], ww)


@pytest.mark.parametrize("case", CASES)
def test_match_gap_tasks_vs_golden(golden, dev_index, case):
    rs = golden.reads(case)
    opt = ol.default_opt(**golden.opt_kwargs(case))
    tasks, opts = percall_tasks(rs, opt)
    res = dev_index.match_gap_batch(rs.codes, tasks, opts)
    exp_n, exp_rows = golden.expected(case, "percall")
    assert np.array_equal(res.n_aln, exp_n)
    assert np.array_equal(el.aln9_to_rows12(res.ordered()), exp_rows)
    assert res.occ_lookups == golden.lookups(case, "percall")
    assert res.kernel_launches >= 1


@pytest.mark.parametrize("case", CASES)
def test_whole_reads_vs_golden(golden, dev_index, case):
    rs = golden.reads(case)
    opt = to_api_opt(ol.default_opt(**golden.opt_kwargs(case)))
    res = dev_index.whole_reads(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, opt)
    exp_n, exp_rows = golden.expected(case, "whole")
    assert np.array_equal(res.n_aln, exp_n)
    assert np.array_equal(el.aln9_to_rows12(res.ordered()), exp_rows)
    assert res.occ_lookups == golden.lookups(case, "whole")


@pytest.mark.parametrize("case", CASES)
def test_splice_seeds_vs_golden(golden, dev_index, case):
    rs = golden.reads(case)
    opt = to_api_opt(ol.default_opt(**golden.opt_kwargs(case)))
    res = dev_index.splice_seeds(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, opt)
    exp_n, exp_rows = golden.expected(case, "seeds")
    assert np.array_equal(res.n_aln, exp_n)
    assert np.array_equal(el.aln9_to_rows12(res.ordered()), exp_rows)
    assert res.occ_lookups == golden.lookups(case, "seeds")


@pytest.mark.parametrize("case", CASES)
def test_cooperative_kernel_path(golden, golden_index, monkeypatch, case):
    """A step budget of one pop makes the fast kernel hand every search that pops anything to the warp-cooperative
    kernel (hsa_coop.cuh); whole-read, per-call and seed results and the lookup count must not change."""
    monkeypatch.setenv("HSA_B200_STEP_BUDGET", "1")
    ix = api.Index.upload(golden_index, 0)
    try:
        rs = golden.reads(case)
        opt0 = ol.default_opt(**golden.opt_kwargs(case))
        res = ix.whole_reads(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, to_api_opt(opt0))
        exp_n, exp_rows = golden.expected(case, "whole")
        if case != "exact_only":
            assert res.n_strict > 0
        assert np.array_equal(res.n_aln, exp_n)
        assert np.array_equal(el.aln9_to_rows12(res.ordered()), exp_rows)
        assert res.occ_lookups == golden.lookups(case, "whole")
        res = ix.splice_seeds(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, to_api_opt(opt0))
        exp_n, exp_rows = golden.expected(case, "seeds")
        assert np.array_equal(res.n_aln, exp_n)
        assert np.array_equal(el.aln9_to_rows12(res.ordered()), exp_rows)
        assert res.occ_lookups == golden.lookups(case, "seeds")
        tasks, opts = percall_tasks(rs, opt0)
        res = ix.match_gap_batch(rs.codes, tasks, opts)
        exp_n, exp_rows = golden.expected(case, "percall")
        assert np.array_equal(res.n_aln, exp_n)
        assert np.array_equal(el.aln9_to_rows12(res.ordered()), exp_rows)
        assert res.occ_lookups == golden.lookups(case, "percall")
    finally:
        ix.close()


def test_strict_rerun_path(golden, golden_index, monkeypatch):
    """A deliberately tiny fast-kernel arena forces the large-capacity re-run; results must not change."""
    monkeypatch.setenv("HSA_B200_ARENA_CAP", "96")
    monkeypatch.setenv("HSA_B200_COOP", "0")                  # straight to the large-capacity kernel
    ix = api.Index.upload(golden_index, 0)
    try:
        case = "cfg5_150bp_n5o2"
        rs = golden.reads(case)
        opt = to_api_opt(ol.default_opt(**golden.opt_kwargs(case)))
        res = ix.whole_reads(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, opt)
        exp_n, exp_rows = golden.expected(case, "whole")
        assert res.n_strict > 0 and res.kernel_launches >= 2
        assert np.array_equal(res.n_aln, exp_n)
        assert np.array_equal(el.aln9_to_rows12(res.ordered()), exp_rows)
        assert res.occ_lookups == golden.lookups(case, "whole")
    finally:
        ix.close()


def test_async_jobs_double_buffered(golden, dev_index):
    """hsa_whole_reads_submit / hsa_job_wait: three batches in flight, results identical to the blocking call
    and delivered per job; a fourth submit is refused until one job has been waited for."""
    cases = ["cfg2_100bp_default", "cfg1_75bp_n2o1", "ragged_nonstop"]
    jobs = []
    for case in cases:
        rs = golden.reads(case)
        opt = to_api_opt(ol.default_opt(**golden.opt_kwargs(case)))
        jobs.append(dev_index.whole_reads_submit(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, opt))
    rs0 = golden.reads(cases[0])
    with pytest.raises(api.HsaError):
        dev_index.whole_reads_submit(rs0.codes, rs0.offsets[:-1].astype(np.uint64), rs0.lens, api.gap_init_opt())
    for case, job in zip(cases, jobs):
        res = job.wait()
        exp_n, exp_rows = golden.expected(case, "whole")
        assert np.array_equal(res.n_aln, exp_n)
        assert np.array_equal(el.aln9_to_rows12(res.ordered()), exp_rows)
        assert res.occ_lookups == golden.lookups(case, "whole")
    with pytest.raises(api.HsaError):
        jobs[0].wait()


def test_empty_and_bad_inputs(dev_index):
    opt = api.gap_init_opt()
    res = dev_index.whole_reads(np.zeros(0, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32), opt)
    assert res.n_aln.shape[0] == 0
    with pytest.raises(api.HsaError):      # empty read
        dev_index.whole_reads(np.zeros(4, np.uint8), np.zeros(1, np.uint64), np.zeros(1, np.uint32), opt)
    with pytest.raises(api.HsaError):      # score range beyond the bucket table must be refused, not truncated
        dev_index.whole_reads(np.zeros(40, np.uint8), np.zeros(1, np.uint64), np.full(1, 40, np.uint32),
                              api.gap_init_opt(s_gapo=60, max_gapo=3))


def test_long_reads_use_the_rows_variant(golden, golden_index, dev_index):
    """Reads of 300-700 bp do not fit the per-lane shared-memory bound bytes: the fast kernel then reads them from
    the rows (search_kernel<..., BIDS_SMEM=false>).  Checked against the oracle, whole reads and seeds."""
    rs = synth.ragged_reads(golden.genome, 160, 300, 700, seed=77, sub_rate=0.004)
    opt = api.gap_init_opt(fnr=0.0, max_diff=3, max_gapo=1)
    oopt = ol.default_opt(fnr=0.0, max_diff=3, max_gapo=1)
    o = ol.Oracle(golden_index)
    res = dev_index.whole_reads(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, opt)
    n_ref, rows_ref = o.whole(rs, oopt)
    assert np.array_equal(res.n_aln, n_ref)
    assert np.array_equal(el.aln9_to_rows12(res.ordered()), rows_ref)
    assert int((n_ref > 0).sum()) > 100
    res = dev_index.splice_seeds(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, opt)
    n_ref, rows_ref = o.seeds(rs, oopt)
    assert np.array_equal(res.n_aln, n_ref)
    assert np.array_equal(el.aln9_to_rows12(res.ordered()), rows_ref)


def test_seeds_async_job_equals_blocking_call(golden, dev_index):
    case = "cfg2_100bp_default"
    rs = golden.reads(case)
    opt = to_api_opt(ol.default_opt(**golden.opt_kwargs(case)))
    a = dev_index.splice_seeds(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, opt)
    j1 = dev_index.splice_seeds_submit(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, opt)
    j2 = dev_index.splice_seeds_submit(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, opt)
    b1, b2 = j1.wait(), j2.wait()
    for b in (b1, b2):
        assert np.array_equal(a.n_aln, b.n_aln) and np.array_equal(a.ordered(), b.ordered()) and a.occ_lookups == b.occ_lookups
    exp_n, exp_rows = golden.expected(case, "seeds")
    assert np.array_equal(b1.n_aln, exp_n) and np.array_equal(el.aln9_to_rows12(b1.ordered()), exp_rows)


def test_medium_batch_vs_oracle_and_properties():
    """A fresh 2 Mb genome, 60k reads (fills the whole grid): bit-exact against the oracle on a sample,
    plus size-independent properties on everything: every hit interval is non-empty and inside the SA range,
    rev interval has the same width, exact reads have a score-0 hit whose interval contains their origin."""
    g = synth.make_genome(2000003, 31)
    index = index_build.build_index(g, device="cuda")
    ix = api.Index.upload(index, 0)
    try:
        rs = synth.simulate_reads(g, 60000, 100, 32)
        opt = api.gap_init_opt()
        res = ix.whole_reads(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, opt)
        f = api.aln_fields(res.aln)
        n = index.fwd.text_length
        assert (f["k"] <= f["l"]).all() and (f["l"] <= n).all()
        assert ((f["l"] - f["k"]) == (f["rev_l"] - f["rev_k"])).all()
        assert (f["score"] == 3 * f["n_mm"].astype(np.int64) + 11 * f["n_gapo"] + 4 * f["n_gape"]).all()
        # per read, scores are non-decreasing except for the top2 tail (discovery order)
        sub = rs.subset(0, 4000)
        n_ref, rows_ref = ol.Oracle(index).whole(sub, ol.default_opt())
        assert np.array_equal(res.n_aln[:4000], n_ref)
        idx = np.concatenate([np.arange(int(o), int(o) + int(c)) for o, c in zip(res.aln_off[:4000], res.n_aln[:4000]) if c])
        assert np.array_equal(el.aln9_to_rows12(res.aln[idx]), rows_ref)
        # idempotence: a second run gives identical per-read results
        res2 = ix.whole_reads(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, opt)
        assert np.array_equal(res.n_aln, res2.n_aln) and np.array_equal(res.ordered(), res2.ordered())
        assert res.occ_lookups == res2.occ_lookups
    finally:
        ix.close()


def test_million_reads_properties_against_the_text():
    """BASELINE-scale batch shape (1 M x 100 bp, defaults) on a 4.6 Mb genome, checked without any reference run:
      * splitting the batch in two and shuffling the reads leave every read's hits unchanged (searches are independent),
      * EVERY gap-free hit interval, resolved to text positions through hsa_sa_values, holds the strand-resolved read
        with exactly n_mm mismatches, and every position of the text where the read's origin lies is inside the
        score-0 interval for error-free reads,
      * a gapped hit's interval is non-empty and its rev interval has the same width."""
    from hsa_b200 import synth_torch
    import torch
    dev = torch.device("cuda", 0)
    G, n, L = 4_600_003, 1_000_000, 100
    genome_t = synth_torch.make_genome(G, 11, dev)
    index = index_build.build_index(genome_t, device=dev)
    ix = api.Index.upload(index, 0)
    try:
        text = genome_t.cpu().numpy()
        reads = synth_torch.simulate_reads(genome_t, n, L, 2024).cpu().numpy()
        off = (np.arange(n, dtype=np.uint64) * L)
        lens = np.full(n, L, dtype=np.uint32)
        opt = api.gap_init_opt()
        res = ix.whole_reads(reads.reshape(-1), off, lens, opt)
        assert int((res.n_aln > 0).sum()) > 0.98 * n
        # split + shuffle invariance
        perm = np.random.default_rng(3).permutation(n)
        res_p = ix.whole_reads(reads[perm].reshape(-1), off, lens, opt)
        assert np.array_equal(res_p.n_aln, res.n_aln[perm])
        half = n // 2
        res_a = ix.whole_reads(reads[:half].reshape(-1), off[:half], lens[:half], opt)
        assert np.array_equal(res_a.n_aln, res.n_aln[:half]) and np.array_equal(res_a.ordered(), res.ordered()[: int(res.n_aln[:half].sum())])
        # hits of the shuffled run, per read, equal the original read's hits
        first = np.concatenate([[0], np.cumsum(res.n_aln.astype(np.int64))[:-1]])
        ordered, ordered_p = res.ordered(), res_p.ordered()
        first_p = np.concatenate([[0], np.cumsum(res_p.n_aln.astype(np.int64))[:-1]])
        one = np.nonzero(res.n_aln == 1)[0]
        inv = np.empty(n, dtype=np.int64); inv[perm] = np.arange(n)
        assert np.array_equal(ordered[first[one]], ordered_p[first_p[inv[one]]])
        # every gap-free hit against the text
        f = api.aln_fields(ordered)
        owner = np.repeat(np.arange(n), res.n_aln)
        width = (f["l"].astype(np.int64) - f["k"].astype(np.int64) + 1)
        assert (width >= 1).all() and ((f["rev_l"].astype(np.int64) - f["rev_k"]) == width - 1).all()
        sel = np.nonzero((f["n_gapo"] == 0) & (width <= 4))[0]
        assert sel.shape[0] > 0.9 * n
        sa_idx = np.concatenate([f["k"][sel] + j for j in range(4)])          # up to four positions per interval
        keep = np.concatenate([width[sel] > j for j in range(4)])
        hit_of = np.concatenate([sel] * 4)[keep]
        pos = ix.sa_values(sa_idx[keep].astype(np.uint32)).astype(np.int64)
        assert (pos + L <= G).all()
        rd = reads[owner[hit_of]]
        rc = np.where(rd[:, ::-1] < 4, 3 - rd[:, ::-1], rd[:, ::-1])
        seq = np.where((f["strand"][hit_of] == 1)[:, None], rc, rd)
        win = text[pos[:, None] + np.arange(L)[None, :]]
        assert np.array_equal((win != seq).sum(axis=1), f["n_mm"][hit_of].astype(np.int64))
    finally:
        ix.close()


def test_short_reads_and_filters_vs_oracle(golden, golden_index, dev_index):
    """Edges of the per-read driver logic (bwtaln.c:314-332): reads no longer than seed_len (the reference then reads
    width_seed out of bounds, SURVEY.md hazard 3; the oracle and the library both search them without seeding), reads
    the N filter drops, and reads with a 15-base poly-A / poly-T prefix (dropped before any search)."""
    o = ol.Oracle(golden_index)
    opt_o = ol.default_opt()
    opt = to_api_opt(opt_o)
    rs = synth.ragged_reads(golden.genome, 600, 12, 40, 9, sub_rate=0.02)
    codes = rs.codes.copy()
    off = rs.offsets
    for r in range(0, 60, 3):                       # a run of N at the start: beyond the N budget of short reads
        codes[off[r]: off[r] + 6] = 4
    for r in range(1, 60, 3):                       # poly-A / poly-T prefixes (only reads >= 15 bp can have one)
        if rs.lens[r] >= 15:
            codes[off[r]: off[r] + 15] = 0 if r % 2 else 3
    rs2 = synth.ReadSet(rs.lens, codes)
    n_ref, rows_ref = o.whole(rs2, opt_o)
    res = dev_index.whole_reads(rs2.codes, rs2.offsets[:-1].astype(np.uint64), rs2.lens, opt)
    assert np.array_equal(res.n_aln, n_ref)
    assert np.array_equal(el.aln9_to_rows12(res.ordered()), rows_ref)
    assert res.occ_lookups == o.last_lookups
    assert int(n_ref.sum()) > 300 and int((n_ref[1:60:3] == 0).sum()) >= 15


@pytest.mark.parametrize("prebin", ["1", "0"], ids=["prebinned_by_length", "input_order"])
def test_ragged_batch_vs_oracle(monkeypatch, prebin):
    """A ragged batch (33..140 bp, so max_diff 3..6 and both seeded and unseeded reads) with and without the pre-binning
    of the work order by (length, max_diff): bit-exact against the oracle either way, results in input order."""
    monkeypatch.setenv("HSA_B200_PREBIN", prebin)
    g = synth.make_genome(2000003, 51)
    index = index_build.build_index(g, device="cuda")
    ix = api.Index.upload(index, 0)
    try:
        rs = synth.ragged_reads(g, 24000, 33, 140, 52, sub_rate=0.012)
        res = ix.whole_reads(rs.codes, rs.offsets[:-1].astype(np.uint64), rs.lens, api.gap_init_opt())
        o = ol.Oracle(index)
        n_ref, rows_ref = o.whole(rs, ol.default_opt())
        assert np.array_equal(res.n_aln, n_ref)
        assert np.array_equal(el.aln9_to_rows12(res.ordered()), rows_ref)
        assert res.occ_lookups == o.last_lookups
        assert int((n_ref > 0).sum()) > 0.9 * rs.n
    finally:
        ix.close()

"""CPU: the oracle (oracle/hsa_oracle.c) against the golden vectors the unmodified reference produced."""
import numpy as np
import pytest

import oracle_lib as ol


@pytest.fixture(scope="module")
def oracle(golden_index):
    return ol.Oracle(golden_index)


def test_maxdiff_table(golden):
    for L, v in golden.meta["maxdiff"].items():
        assert ol.lib().hsao_cal_maxdiff(int(L), 0.02, 0.04) == v


def test_rank_matches_reference(golden, oracle):
    idx, occ = golden.arr["occ_idx"], golden.arr["occ"]
    o4f, o1f = oracle.occ(0, idx)
    o4r, o1r = oracle.occ(1, idx)
    assert np.array_equal(o4f, occ[:, 0:4]) and np.array_equal(o4r, occ[:, 4:8])
    assert np.array_equal(o1f, occ[:, 8:12]) and np.array_equal(o1r, occ[:, 12:16])


def test_sa_value_matches_reference(golden, oracle):
    """BWTSaValue (BWT.c:1195-1225): SA index -> text position, with the number of PsiMinus steps walked."""
    idx = golden.arr["sa_idx"]
    val, steps = oracle.sa_values(idx)
    assert np.array_equal(val, golden.arr["sa_val"]) and np.array_equal(steps, golden.arr["sa_steps"])
    # what the values mean: SA[i] is the start of the i-th smallest suffix; check a few against the text itself
    n = golden.genome.shape[0]
    assert val[0] == 0xFFFFFFFF                        # SA[0] (the '$' suffix) is kept as -1, BWT.c:222
    assert val[6] == 0 and idx[6] == golden.meta["index"]["fwd"]["inverse_sa0"]   # the whole text sits at inverseSa0
    ok = idx != 0
    assert int(val[ok].max()) < n


def test_locate_matches_reference(golden):
    """BWTRetrievePositionFromSAIndex (2BWT-Interface.c:329-362) on the multi-record genome: every SA index but 0 ->
    {occ_pos, chromosome id, 1-based position in the chromosome}, against the reference's own output; the product's
    builder and block list reproduce the reference's index and annotation for that FASTA."""
    from hsa_b200 import index_build, index_io
    import make_golden
    recs, text = golden.genome2()
    ix2 = index_build.build_index(text, device="cpu")
    assert make_golden.index_digest(ix2) == golden.meta["index2"]
    blocks = index_io.blocks_of_records([r.shape[0] for r in recs])
    assert np.array_equal(blocks.table(), golden.arr["loc_blocks"])
    idx = np.arange(1, text.shape[0] + 1, dtype=np.uint32)
    out = ol.Oracle(ix2).locate(idx, blocks)
    assert np.array_equal(out, golden.arr["loc_out"])
    # what it means: record seq_id, 1-based offset ori_pos, holds the suffix that starts at occ_pos
    for q in range(0, idx.shape[0], 997):
        occ, sid, op = (int(x) for x in out[q])
        assert np.array_equal(recs[sid][op - 1: op + 7], text[occ: occ + 8][: recs[sid].shape[0] - (op - 1)])


@pytest.mark.parametrize("case", ["cfg1_75bp_n2o1", "cfg2_100bp_default", "ragged_nonstop"])
def test_width_matches_reference(golden, oracle, case):
    rs = golden.reads(case)
    for r, (bid, w) in enumerate(golden.widths(case)):
        b, ww = oracle.cal_width(rs.read(r))
        assert b == bid and np.array_equal(ww, w)


@pytest.mark.parametrize("mode", ["percall", "whole", "seeds"])
@pytest.mark.parametrize("case", ["cfg1_75bp_n2o1", "cfg2_100bp_default", "cfg5_150bp_n5o2", "ragged_nonstop",
                                  "ragged_loggap_gape", "short_entries", "exact_only", "noskip_gaps"])
def test_search_matches_reference(golden, oracle, case, mode):
    rs = golden.reads(case)
    opt = ol.default_opt(**golden.opt_kwargs(case))
    n_aln, rows = getattr(oracle, mode)(rs, opt)
    exp_n, exp_rows = golden.expected(case, mode)
    assert np.array_equal(n_aln, exp_n)
    assert np.array_equal(rows, exp_rows)
    assert oracle.last_lookups == golden.lookups(case, mode)

"""CPU: the product's index builder (hsa_b200/index_build.py) reproduces the reference builder bit for bit."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import common
import oracle_lib as ol
from hsa_b200 import index_build, index_io, synth


def _digest(ix):
    import make_golden
    return make_golden.index_digest(ix)


def test_matches_reference_builder_digest(golden, golden_index):
    assert _digest(golden_index) == golden.meta["index"]


def test_file_roundtrip(golden_index, tmp_path):
    p = str(tmp_path / "g.index")
    index_io.save_bwt(golden_index.fwd, p + ".bwt", p + ".fmv")
    index_io.save_bwt(golden_index.rev, p + ".rev.bwt", p + ".rev.fmv")
    index_io.save_sa(golden_index.fwd, p + ".sa")
    back = index_io.load_index(str(tmp_path / "g"))
    assert back.fwd.sa_interval == golden_index.fwd.sa_interval == 8
    assert np.array_equal(back.fwd.sa_value, golden_index.fwd.sa_value) and back.rev.sa_value is None
    for name in ("fwd", "rev"):
        a, b = getattr(golden_index, name), getattr(back, name)
        assert a.inverse_sa0 == b.inverse_sa0 and np.array_equal(a.cumulative_freq, b.cumulative_freq)
        assert np.array_equal(a.bwt_code, b.bwt_code) and np.array_equal(a.occ_value, b.occ_value)
        assert np.array_equal(a.occ_value_major, b.occ_value_major)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("length,seed,repeat", [(1027, 1, False), (4097, 2, False), (65536 + 257, 3, False),
                                                (131072 + 5, 4, True), (256 * 37 + 1, 5, True)])
def test_against_live_reference_builder(length, seed, repeat):
    """Edge sizes around the 256 / 65536 sampling intervals, plain and repeat-rich."""
    g = synth.make_repeat_genome(length, seed, n_dups=5, dup_len=40, tandem=3) if repeat else synth.make_genome(length, seed)
    mine = index_build.build_index(g, device="cpu")
    with tempfile.TemporaryDirectory() as td:
        synth.write_fasta(os.path.join(td, "g.fa"), g)
        subprocess.run([ol.REF_BIN, "index", "g", "g.fa"], cwd=td, check=True, stdout=subprocess.DEVNULL)
        ref = index_io.load_index(os.path.join(td, "g"))
    assert _digest(mine) == _digest(ref)


def test_annotation_roundtrip(tmp_path):
    """load_ann parses what the reference builder writes (HSP.c:330-345) for a multi-record FASTA."""
    b = index_io.blocks_of_records([3001, 777, 150])
    with open(tmp_path / "x.ann", "w") as f:
        f.write("3928\t3\t0\n")
        for i in range(3):
            f.write(f"4\trec{i}\n")
        f.write("3\n")
        for row in b.table().tolist():
            f.write("%d\t%u\t%u\t%u\n" % tuple(row))
    back = index_io.load_ann(str(tmp_path / "x.ann"))
    assert np.array_equal(back.table(), b.table()) and back.names == ["rec0", "rec1", "rec2"]

"""CPU, build container only: the oracle against the LIVE unmodified reference (oracle/_ref) on fresh seeds,
larger than the committed golden vectors.  Skipped where /root/reference was never compiled."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_lib as ol
from hsa_b200 import index_io, synth

pytestmark = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def live():
    g = synth.make_repeat_genome(300007, 901)
    td = tempfile.mkdtemp(prefix="hsa_live_")
    synth.write_fasta(os.path.join(td, "g.fa"), g)
    subprocess.run([ol.REF_BIN, "index", "g", "g.fa"], cwd=td, check=True, stdout=subprocess.DEVNULL)
    ix = index_io.load_index(os.path.join(td, "g"))
    return g, td, ol.Oracle(ix)


@pytest.mark.parametrize("length,okw,seed", [(75, dict(max_diff=2, max_gapo=1, fnr=0.0), 1), (100, {}, 2),
                                             (150, dict(max_diff=5, max_gapo=2, fnr=0.0), 3)])
@pytest.mark.parametrize("mode", ["percall", "whole", "seeds"])
def test_oracle_equals_live_reference(live, length, okw, seed, mode):
    g, td, oracle = live
    rs = synth.simulate_reads(g, 800, length, seed, sub_rate=0.015, indel_frac=0.08)
    rp = os.path.join(td, f"r{seed}.reads")
    synth.write_reads_bin(rp, rs)
    opt = ol.default_opt(**okw)
    outp = os.path.join(td, f"r{seed}.{mode}.aln")
    j = ol.run_ref([mode, os.path.join(td, "g"), rp, outp] + ol.opt_args(opt), count=True)
    exp_n, exp_rows = synth.read_aln_dump(outp)
    n_aln, rows = getattr(oracle, mode)(rs, opt)
    assert np.array_equal(n_aln, exp_n) and np.array_equal(rows, exp_rows)
    assert oracle.last_lookups == j["occ4"] + j["occ1"]

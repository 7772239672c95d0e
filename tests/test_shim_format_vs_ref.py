"""CPU, build container only: the shim's SAM formatter (shim/hsa_gpu_shim.c: sam_line / sam_print_batch, the replacement of the
print loop of generate_sam_se_core, bwtse.c:922-926) against the reference's own bwa_print_sam1 -- on fields computed by the
reference itself, so no GPU is involved.  `hsa_ref sam` is the stock program; `hsa_ref_gpu samfmt` is the same program with only
the print loop replaced (ref_harness.c).  qual=1 gives every read a quality string, every third a barcode and every fifth a clipped
length, which the GPU-side tests (reads without qualities) cannot reach: quality reversal on the reverse strand, BC:Z and XC:i."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
import sam_common as sc
from hsa_b200 import synth

REF_GPU = os.path.join(ol.ROOT, "oracle", "_ref", "hsa_ref_gpu")
pytestmark = pytest.mark.skipif(not (ol.have_ref() and os.path.exists(REF_GPU)), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("shimfmt")
    g = synth.make_repeat_genome(400009, seed=61)                 # repeats: reads with several hits (X0 / X1 / XA)
    synth.write_fasta(str(d / "g.fa"), g)
    subprocess.run([ol.REF_BIN, "index", "g", "g.fa"], cwd=d, check=True, capture_output=True)
    a = synth.simulate_reads(g, 2500, 100, seed=62, indel_frac=0.2)
    b, _ = synth.simulate_spliced_reads(g, 700, 100, seed=63, min_intron=60, max_intron=3000)
    rng = np.random.default_rng(64)
    junk = rng.integers(0, 4, size=(40, 100), dtype=np.uint8)
    codes = np.concatenate([a.codes.reshape(-1, 100), b.codes.reshape(-1, 100), junk])
    codes = codes[rng.permutation(codes.shape[0])]
    rs = synth.ReadSet(np.full(codes.shape[0], 100, np.uint32), np.ascontiguousarray(codes).reshape(-1))
    synth.write_reads_bin(str(d / "r.reads"), rs)
    return d


@pytest.mark.parametrize("opts", [["qual=1", "batch=5000"], ["qual=0", "batch=900"], ["qual=1", "batch=700", "mode=0", "max_top2=1"]],
                         ids=["qualities_barcodes_clipping", "plain_small_batches", "CM_tag_and_no_X1"])
@pytest.mark.parametrize("threads", ["1", "6"])
def test_formatter_prints_what_bwa_print_sam1_prints(workdir, opts, threads):
    env = dict(os.environ, HSA_GPU_SHIM_THREADS=threads)
    ref = subprocess.run([ol.REF_BIN, "sam", "g", "r.reads", "a.bin", "a.sam"] + opts, cwd=workdir, capture_output=True, text=True)
    assert ref.returncode == 0, ref.stderr[-2000:]
    fmt = subprocess.run([REF_GPU, "samfmt", "g", "r.reads", "b.bin", "b.sam"] + opts, cwd=workdir, capture_output=True, text=True, env=env)
    assert fmt.returncode == 0, fmt.stderr[-2000:]
    a, b = (open(workdir / f, "rb").read() for f in ("a.sam", "b.sam"))
    pa, pb = sc.printable_lines(a), sc.printable_lines(b)
    assert pa.count(b"\n") > 2500 and b"XT:A:S" in pa and b"XA:Z:" in pa and b"MD:Z:" in pa
    if "qual=1" in opts:
        assert b"BC:Z:TTAGGC" in pa and b"XC:i:97" in pa and b"\t*\tXT" not in pa
    assert pa == pb, "SAM text differs from bwa_print_sam1's"
    assert a.count(b"\n") == b.count(b"\n")
    assert open(workdir / "a.bin", "rb").read() == open(workdir / "b.bin", "rb").read()

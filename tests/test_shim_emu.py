"""CPU, build container only: the drop-in END TO END without a GPU.

oracle/_ref/hsa_ref_emu = the unmodified reference objects + our harness + shim/hsa_gpu_shim.c, linked with
tests/emu/hsa_emu_backend.cpp instead of libhsa_b200.so: the C ABI entry points the shim calls (hsa_whole_reads,
hsa_splice_match_batch, hsa_sam_se_batch, ...) run the device sources on the host (TEST INFRASTRUCTURE -- the product has no CPU
path).  So everything that is host logic in the boundary -- the option switch after the first fallback read, the two GPU passes
of a process's first batch, the order-dependent filter pass, helper threads, the splice batch on its own thread, the SAM
formatter -- is checked here against the stock program, byte for byte, the way tests/test_gpu_shim.py checks it on a B200."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
import sam_common as sc
from hsa_b200 import synth

REF_EMU = os.path.join(ol.ROOT, "oracle", "_ref", "hsa_ref_emu")
pytestmark = pytest.mark.skipif(not (ol.have_ref() and os.path.exists(REF_EMU)),
                                reason="oracle/_ref/hsa_ref_emu not built (needs /root/reference; make -C oracle shim_emu)")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("shimemu")
    g = synth.make_genome(700001, seed=71)
    synth.write_fasta(str(d / "g.fa"), g)
    subprocess.run([ol.REF_BIN, "index", "g", "g.fa"], cwd=d, check=True, capture_output=True)
    a = synth.simulate_reads(g, 2600, 100, seed=72, indel_frac=0.15)
    b, _ = synth.simulate_spliced_reads(g, 700, 100, seed=73, min_intron=60, max_intron=3000)
    rng = np.random.default_rng(74)
    junk = rng.integers(0, 4, size=(40, 100), dtype=np.uint8)
    codes = np.concatenate([a.codes.reshape(-1, 100), b.codes.reshape(-1, 100), junk])
    codes = codes[rng.permutation(codes.shape[0])]
    rs = synth.ReadSet(np.full(codes.shape[0], 100, np.uint32), np.ascontiguousarray(codes).reshape(-1))
    synth.write_reads_bin(str(d / "r.reads"), rs)
    return d


@pytest.mark.parametrize("splice", ["1", "0"], ids=["splice_batch_emulated", "splice_by_the_reference"])
@pytest.mark.parametrize("opts", [["batch=1500"], ["mode=2", "batch=1500"], ["fnr=0", "max_diff=3", "max_gapo=2", "batch=1500"], ["batch=700"],
                                  ["batch=100000"]],
                         ids=["default_gape_switch", "mode_without_gape", "fixed_maxdiff", "small_batches", "one_batch"])
def test_driver_with_the_shim_matches_the_stock_driver(workdir, opts, splice):
    env = dict(os.environ, HSA_GPU_SPLICE=splice, HSA_GPU_SHIM_THREADS="5")
    cpu = subprocess.run([ol.REF_BIN, "driver", "g", "r.reads", "cpu.aln"] + opts, cwd=workdir, capture_output=True, text=True)
    assert cpu.returncode == 0, cpu.stderr[-2000:]
    emu = subprocess.run([REF_EMU, "gpudriver", "g", "r.reads", "emu.aln"] + opts, cwd=workdir, capture_output=True, text=True, env=env)
    assert emu.returncode == 0, emu.stderr[-2000:]
    n_c, rows_c = synth.read_aln_dump(str(workdir / "cpu.aln"))
    n_e, rows_e = synth.read_aln_dump(str(workdir / "emu.aln"))
    assert (n_c > 0).sum() > 2500 and (n_c == 0).sum() > 30 and int((n_c == 2).sum()) > 100     # hits, misses and spliced pairs
    assert np.array_equal(n_c, n_e)
    assert np.array_equal(rows_c, rows_e)


@pytest.mark.parametrize("opts", [["batch=1500"], ["batch=700", "fnr=0", "max_diff=3", "max_gapo=2", "qual=1"]], ids=["default", "fixed_maxdiff_qualities"])
def test_whole_program_with_both_shim_stages(workdir, opts):
    """bwa_cal_sa_reg_gap_gpu + generate_sam_se_core_gpu (selection on the drand48 stream carried across batches, positions,
    banded DP, MD, the shim's own SAM formatter) against the stock program: SAM text and every bwa_seq_t field."""
    cpu = subprocess.run([ol.REF_BIN, "sam", "g", "r.reads", "c.bin", "c.sam"] + opts, cwd=workdir, capture_output=True, text=True)
    assert cpu.returncode == 0, cpu.stderr[-2000:]
    emu = subprocess.run([REF_EMU, "gpusam", "g", "r.reads", "e.bin", "e.sam"] + opts, cwd=workdir, capture_output=True, text=True,
                         env=dict(os.environ, HSA_GPU_SHIM_THREADS="5"))
    assert emu.returncode == 0, emu.stderr[-2000:]
    a, b = (sc.printable_lines(open(workdir / f, "rb").read()) for f in ("c.sam", "e.sam"))
    assert a.count(b"\n") > 2500 and b"XT:A:S" in a and b"M1I" in a
    assert a == b, "SAM text differs from the stock program's"
    assert open(workdir / "c.bin", "rb").read() == open(workdir / "e.bin", "rb").read()


def test_sa_values_through_the_shim(workdir):
    """`gpusa`: the reference's BWTSaValue against shim/hsa_gpu_shim.c's hsa_gpu_sa_values (sa_value_dev on the host) for 50 000 SA
    indices, compared in C."""
    rng = np.random.default_rng(75)
    idx = rng.integers(0, 700001 + 1, size=50_000).astype(np.uint32)
    idx[:4] = [0, 1, 700001, 700000]
    with open(workdir / "sa.bin", "wb") as f:
        np.asarray([idx.shape[0]], dtype=np.uint32).tofile(f)
        idx.tofile(f)
    r = subprocess.run([REF_EMU, "gpusa", "g", "sa.bin"], cwd=workdir, capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-2000:])
    assert '"mismatches":0' in r.stdout


@pytest.mark.parametrize("mode,opts", [("percall", []), ("percall", ["clear_gape=0", "max_gapo=2"]), ("seeds", [])],
                         ids=["whole_read_frames", "gape_counts_two_gap_opens", "splice_seed_frames"])
def test_per_call_symbol(workdir, mode, opts):
    """bwt_match_gap_gpu(bwt_aux_t*, int*) (bwtgap.h:26) through hsa_match_gap_call with the CALLER's width arrays: every call of
    the harness's percall / seeds loop goes through the reference and through the symbol on the same frame; hits are compared with
    memcmp and width_back after gap_shadow's rewrite word for word, in C; the dumps are compared here."""
    rs = synth.read_reads_bin(str(workdir / "r.reads")).subset(0, 1000)
    synth.write_reads_bin(str(workdir / "r_small.reads"), rs)
    cpu = subprocess.run([ol.REF_BIN, mode, "g", "r_small.reads", "pc_cpu.aln"] + opts, cwd=workdir, capture_output=True, text=True)
    assert cpu.returncode == 0, cpu.stderr[-2000:]
    emu = subprocess.run([REF_EMU, "gpu" + mode, "g", "r_small.reads", "pc_emu.aln"] + opts, cwd=workdir, capture_output=True, text=True)
    assert emu.returncode == 0, (emu.stdout[-500:], emu.stderr[-2000:])
    assert '"call_mismatches":0,"width_mismatches":0' in emu.stdout
    n_c, rows_c = synth.read_aln_dump(str(workdir / "pc_cpu.aln"))
    n_e, rows_e = synth.read_aln_dump(str(workdir / "pc_emu.aln"))
    assert int((n_c > 0).sum()) > 500 and np.array_equal(n_c, n_e) and np.array_equal(rows_c, rows_e)

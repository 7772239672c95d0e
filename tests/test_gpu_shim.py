"""GPU (-m gpu): the reference-side binding as real C code (shim/hsa_gpu_shim.c).

oracle/_ref/hsa_ref_gpu links the UNMODIFIED reference objects, our harness and the shim against libhsa_b200.so.
Its `gpudriver` mode runs the reference's batch loop with bwa_cal_sa_reg_gap replaced by bwa_cal_sa_reg_gap_gpu
(whole-read searches on the GPU; bwt_splice_match and everything after it stay reference host code); `driver` runs
the stock CPU bwa_cal_sa_reg_gap.  Both dump every read's bwt_aln1_t array; the dumps must be identical -- including
the reads that went through the splice fallback and the reference's option switch after the first of them
(SURVEY.md section 3.2).  The index is built by the reference's own builder (the binaries travel with the snapshot)."""
import os
import subprocess

import numpy as np
import pytest

from hsa_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "hsa_ref")
REF_GPU = os.path.join(ROOT, "oracle", "_ref", "hsa_ref_gpu")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    # never skip on a GPU run: without these binaries the drop-in boundary would go unchecked
    assert os.path.exists(REF) and os.path.exists(REF_GPU), (
        "oracle/_ref/hsa_ref[_gpu] are missing: build them where the reference sources exist "
        "(python -c 'import __graft_entry__ as g; g.build()' runs make -C oracle ref shim); they travel with the snapshot")
    d = tmp_path_factory.mktemp("shim")
    g = synth.make_genome(1200011, seed=41)
    synth.write_fasta(str(d / "g.fa"), g)
    subprocess.run([REF, "index", "g", "g.fa"], cwd=d, check=True, capture_output=True)
    # ordinary reads, reads with a 1-bp indel (gapped hits), spliced reads and a few hopeless ones: the last two
    # kinds fall through to bwt_splice_match and flip the driver's option switch
    a = synth.simulate_reads(g, 5000, 100, seed=42, indel_frac=0.15)
    b, _ = synth.simulate_spliced_reads(g, 1500, 100, seed=43, min_intron=60, max_intron=3000)
    rng = np.random.default_rng(44)
    junk = synth.ReadSet(np.full(60, 100, np.uint32), rng.integers(0, 4, size=6000, dtype=np.uint8))
    codes = np.concatenate([a.codes.reshape(-1, 100), b.codes.reshape(-1, 100), junk.codes.reshape(-1, 100)])
    codes = codes[rng.permutation(codes.shape[0])]
    rs = synth.ReadSet(np.full(codes.shape[0], 100, np.uint32), np.ascontiguousarray(codes).reshape(-1))
    synth.write_reads_bin(str(d / "r.reads"), rs)
    return d


@pytest.mark.parametrize("splice_on_gpu", ["1", "0"], ids=["splice_on_gpu", "splice_on_host"])
@pytest.mark.parametrize("opts", [[], ["mode=2"], ["fnr=0", "max_diff=3", "max_gapo=2"], ["batch=1700"]],
                         ids=["default_gape_switch", "mode_without_gape", "fixed_maxdiff", "small_batches"])
def test_gpu_driver_matches_stock_driver(workdir, opts, splice_on_gpu, monkeypatch):
    """splice_on_gpu: the reads that found nothing go through hsa_splice_match_batch (bwt_splice_match on the GPU, one
    batch per driver call, each read with the option state the driver holds for it); splice_on_host: through the
    reference's own bwt_splice_match.  Either way every read's output equals the stock driver's."""
    monkeypatch.setenv("HSA_GPU_SPLICE", splice_on_gpu)
    args = list(opts)
    if not any(o.startswith("batch=") for o in args):
        args.append("batch=3000")
    cpu = subprocess.run([REF, "driver", "g", "r.reads", "cpu.aln"] + args, cwd=workdir, check=True, capture_output=True, text=True)
    gpu = subprocess.run([REF_GPU, "gpudriver", "g", "r.reads", "gpu.aln"] + args, cwd=workdir, capture_output=True, text=True)
    assert gpu.returncode == 0, gpu.stderr[-2000:]
    n_c, rows_c = synth.read_aln_dump(str(workdir / "cpu.aln"))
    n_g, rows_g = synth.read_aln_dump(str(workdir / "gpu.aln"))
    assert (n_c > 0).sum() > 4000 and (n_c == 0).sum() > 50        # the case really covers hits, splices and misses
    assert np.array_equal(n_c, n_g)
    assert np.array_equal(rows_c, rows_g)
    assert '"aligned_any"' in cpu.stdout and '"aligned_any"' in gpu.stdout


@pytest.mark.parametrize("threads", ["1", "8"], ids=["no_helper_threads", "eight_threads"])
def test_gpu_driver_helper_threads_do_not_change_the_output(workdir, threads, monkeypatch):
    """The shim spreads the order-independent host work (packing, N counts, the per-read hit arrays) over helper threads and
    keeps the splice batch in flight meanwhile; one batch of all 6 560 reads is large enough for every threaded loop to be
    cut into several chunks.  Output equals the stock driver's with and without helpers."""
    monkeypatch.setenv("HSA_GPU_SHIM_THREADS", threads)
    monkeypatch.setenv("HSA_GPU_SHIM_TIMING", "1")
    args = ["batch=100000"]
    subprocess.run([REF, "driver", "g", "r.reads", "cpu_t.aln"] + args, cwd=workdir, check=True, capture_output=True, text=True)
    gpu = subprocess.run([REF_GPU, "gpudriver", "g", "r.reads", "gpu_t.aln"] + args, cwd=workdir, capture_output=True, text=True)
    assert gpu.returncode == 0, gpu.stderr[-2000:]
    assert "[hsa_gpu] seconds:" in gpu.stderr
    n_c, rows_c = synth.read_aln_dump(str(workdir / "cpu_t.aln"))
    n_g, rows_g = synth.read_aln_dump(str(workdir / "gpu_t.aln"))
    assert (n_c > 0).sum() > 4000 and np.array_equal(n_c, n_g) and np.array_equal(rows_c, rows_g)


def test_gpu_sa_values_match_BWTSaValue_inside_the_reference_program(workdir):
    """`gpusa`: the reference's own BWTSaValue (host) against shim/hsa_gpu_shim.c's hsa_gpu_sa_values (GPU batch call on
    the reference-loaded saValue array) for 200 000 SA indices of the reference-built index, compared in C."""
    rng = np.random.default_rng(45)
    idx = rng.integers(0, 1200011 + 1, size=200_000).astype(np.uint32)
    idx[:4] = [0, 1, 1200011, 1200010]
    with open(workdir / "sa.bin", "wb") as f:
        np.asarray([idx.shape[0]], dtype=np.uint32).tofile(f)
        idx.tofile(f)
    r = subprocess.run([REF_GPU, "gpusa", "g", "sa.bin"], cwd=workdir, capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-2000:])
    assert '"mismatches":0' in r.stdout


@pytest.mark.parametrize("mode,opts", [("percall", []), ("percall", ["clear_gape=0", "max_gapo=2"]), ("seeds", [])],
                         ids=["whole_read_frames", "gape_counts_two_gap_opens", "splice_seed_frames"])
def test_per_call_symbol_matches_bwt_match_gap(workdir, mode, opts):
    """bwt_match_gap_gpu(bwt_aux_t*, int*) -- the per-call drop-in symbol (bwtgap.h:26), a batch of one through
    hsa_match_gap_call with the CALLER's width arrays: every call of the harness's percall / seeds loop is made through
    the reference and through the GPU symbol on the same frame; hits must be byte-identical (memcmp of the bwt_aln1_t
    arrays) and so must width_back after gap_shadow's in-place rewrite.  The dump (GPU results) equals the CPU dump."""
    rs = synth.read_reads_bin(str(workdir / "r.reads")).subset(0, 1500)
    synth.write_reads_bin(str(workdir / "r_small.reads"), rs)
    cpu = subprocess.run([REF, mode, "g", "r_small.reads", "pc_cpu.aln"] + opts, cwd=workdir, check=True, capture_output=True, text=True)
    gpu = subprocess.run([REF_GPU, "gpu" + mode, "g", "r_small.reads", "pc_gpu.aln"] + opts, cwd=workdir, capture_output=True, text=True)
    assert gpu.returncode == 0, (gpu.stdout[-500:], gpu.stderr[-2000:])
    assert '"call_mismatches":0,"width_mismatches":0' in gpu.stdout
    n_c, rows_c = synth.read_aln_dump(str(workdir / "pc_cpu.aln"))
    n_g, rows_g = synth.read_aln_dump(str(workdir / "pc_gpu.aln"))
    assert int((n_c > 0).sum()) > 1000 and np.array_equal(n_c, n_g) and np.array_equal(rows_c, rows_g)


@pytest.mark.parametrize("opts", [["batch=3000"], ["batch=1700", "fnr=0", "max_diff=3", "max_gapo=2"]], ids=["default", "fixed_maxdiff_small_batches"])
def test_whole_program_sam_output_is_byte_identical(workdir, opts):
    """The reference's batch loop (bwtaln.c:477-522) with BOTH stages from the shim -- bwa_cal_sa_reg_gap_gpu and
    generate_sam_se_core_gpu (hit selection on the drand48 stream, positions, CIGAR from the banded DP, MD / NM from
    hsa_sam_se_batch; lines printed by the reference's own bwa_print_sam1) -- against the stock program: the SAM text on
    stdout and the dump of every read's bwa_seq_t fields must be byte-identical, across batches."""
    cpu = subprocess.run([REF, "sam", "g", "r.reads", "cpu.bin", "cpu.sam"] + opts, cwd=workdir, check=True, capture_output=True, text=True)
    gpu = subprocess.run([REF_GPU, "gpusam", "g", "r.reads", "gpu.bin", "gpu.sam"] + opts, cwd=workdir, capture_output=True, text=True)
    assert gpu.returncode == 0, (gpu.stdout[-500:], gpu.stderr[-2000:])
    import sam_common as sc
    # (a spliced hit with a negative intron length makes bwa_print_sam1 index past its CIGAR table, bwtse.c:712: such lines
    # print whatever byte the binary holds there; they are compared through the field dump below instead)
    a, b = (sc.printable_lines(open(workdir / f, "rb").read()) for f in ("cpu.sam", "gpu.sam"))
    assert a.count(b"\n") > 5000 and b"XT:A:S" in a and b"M1I" in a                      # gapped and spliced reads among them
    assert a == b, "SAM text differs from the stock program's"
    assert open(workdir / "cpu.bin", "rb").read() == open(workdir / "gpu.bin", "rb").read()
    assert '"secs_sam"' in cpu.stdout and '"secs_sam"' in gpu.stdout

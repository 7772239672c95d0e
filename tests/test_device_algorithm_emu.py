"""CPU: the CUDA worker's source (hsa_b200/csrc/hsa_core.cuh) compiled for the host, against the golden
vectors and the oracle.  This checks the device ALGORITHM where no GPU exists; the GPU tests
(test_gpu_parity.py) check the compiled kernels themselves through the C ABI."""
import numpy as np
import pytest

import emu_lib as el
import oracle_lib as ol


@pytest.fixture(scope="module")
def emu(golden_index):
    return el.Emu(golden_index)


def test_rank_both_layouts(golden, emu):
    idx, occ = golden.arr["occ_idx"], golden.arr["occ"]
    for which, cols in ((0, slice(0, 4)), (1, slice(4, 8))):
        assert np.array_equal(emu.occ(which, 0, idx), occ[:, cols])     # reference layout
        assert np.array_equal(emu.occ(which, 1, idx), occ[:, cols])     # re-packed device layout


def test_sa_value_device_code(golden, emu):
    """psi_minus_dev / sa_value_dev (one sector per PsiMinus step on the re-packed layout) against BWTSaValue."""
    val, steps = emu.sa_values(golden.arr["sa_idx"])
    assert np.array_equal(val, golden.arr["sa_val"]) and np.array_equal(steps, golden.arr["sa_steps"])


def test_sa_value_is_the_suffix_position(golden, emu):
    """Independent of the reference: consecutive SA indices name lexicographically increasing suffixes of the text."""
    text = golden.genome
    n = text.shape[0]
    idx = np.arange(5000, 5064, dtype=np.uint32)
    pos, _ = emu.sa_values(idx)
    sufs = [bytes(text[int(p): int(p) + 64]) for p in pos]
    assert sufs == sorted(sufs) and len(set(int(p) for p in pos)) == 64 and int(pos.max()) < n


def test_locate_device_code(golden):
    """locate_dev (the block search of BWTRetrievePositionFromSAIndex) on the reference's block list and positions."""
    from hsa_b200 import index_io
    recs, _ = golden.genome2()
    blocks = index_io.blocks_of_records([r.shape[0] for r in recs])
    exp = golden.arr["loc_out"]
    sid, op = el.Emu.locate(blocks, exp[:, 0])
    assert np.array_equal(sid, exp[:, 1]) and np.array_equal(op, exp[:, 2])
    sid, op = el.Emu.locate(blocks, np.asarray([0xFFFFFFFF, int(blocks.end[-1]) + 1], dtype=np.uint32))
    assert (sid == 0xFFFFFFFF).all() and (op == 0xFFFFFFFF).all()      # outside every block: outputs untouched


def test_width(golden, emu):
    case = "ragged_nonstop"
    rs = golden.reads(case).subset(0, 64)
    bid, w = emu.width(rs)
    off = rs.offsets
    for r, (b, ww) in enumerate(golden.widths(case)):
        assert bid[r] == b
        assert np.array_equal(w[off[r] + r: off[r] + r + int(rs.lens[r]) + 1], ww)


@pytest.mark.parametrize("arena_cap", [1022, 65535], ids=["fast16", "large32"])
@pytest.mark.parametrize("mode", ["percall", "whole", "seeds"])
@pytest.mark.parametrize("case", ["cfg1_75bp_n2o1", "cfg2_100bp_default", "cfg5_150bp_n5o2", "ragged_nonstop",
                                  "ragged_loggap_gape", "short_entries", "exact_only", "noskip_gaps"])
def test_search_matches_reference(golden, emu, case, mode, arena_cap):
    """fast16: the fast configuration (16-bit link halves, 64 score buckets, bound bytes in shared memory), with
    the items it cannot hold re-run by the large-capacity one exactly as hsa_b200.cu's run_batch does;
    large32: the large-capacity configuration (32-bit halves, bound bytes in the rows) for everything."""
    rs = golden.reads(case)
    opt = ol.default_opt(**golden.opt_kwargs(case))
    n_aln, rows, status = getattr(emu, mode)(rs, opt, arena_cap=arena_cap, hit_cap=4096, rerun_cap=65535)
    exp_n, exp_rows = golden.expected(case, mode)
    assert int((status != 0).sum()) == 0
    assert np.array_equal(n_aln, exp_n)
    assert np.array_equal(rows, exp_rows)
    assert emu.last_lookups == golden.lookups(case, mode)


def test_capacity_overflow_is_flagged_not_silent(golden, emu):
    """With a tiny stack arena the worker must flag the read for the strict re-run, never emit hits for it."""
    case = "cfg5_150bp_n5o2"
    rs = golden.reads(case)
    opt = ol.default_opt(**golden.opt_kwargs(case))
    n_aln, rows, status = emu.whole(rs, opt, arena_cap=128, hit_cap=32)
    exp_n, _ = golden.expected(case, "whole")
    flagged = status == 1
    assert flagged.any() and emu.last_strict == int(flagged.sum())
    assert np.array_equal(n_aln[~flagged], exp_n[~flagged])
    assert (n_aln[flagged] == 0).all()


@pytest.mark.parametrize("mode", ["percall", "whole", "seeds"])
@pytest.mark.parametrize("case", ["cfg1_75bp_n2o1", "cfg2_100bp_default", "cfg5_150bp_n5o2", "ragged_nonstop",
                                  "ragged_loggap_gape", "short_entries", "exact_only", "noskip_gaps"])
def test_cooperative_kernel_matches_reference(golden, emu, case, mode):
    """Every search is handed to the warp-cooperative kernel (hsa_coop.cuh: speculative waves over the lowest
    bucket, ordered commit) by giving the fast configuration a step budget of one pop; what the cooperative
    kernel cannot hold (scores >= 64 in NONSTOP mode) falls through to the large-capacity configuration."""
    rs = golden.reads(case)
    opt = ol.default_opt(**golden.opt_kwargs(case))
    n_aln, rows, status = getattr(emu, mode)(rs, opt, arena_cap=1022, hit_cap=32, rerun_cap=65535, coop=True, step_budget=1)
    exp_n, exp_rows = golden.expected(case, mode)
    if case != "exact_only":                   # (max_diff 0: nothing is ever popped, so nothing runs out of budget)
        assert emu.last_flagged_first > 0      # the fast configuration really handed searches on
    assert int((status != 0).sum()) == 0
    assert np.array_equal(n_aln, exp_n)
    assert np.array_equal(rows, exp_rows)
    assert emu.last_lookups == golden.lookups(case, mode)


def test_device_sources_under_address_and_ub_sanitizers():
    """compute-sanitizer is not available on the GPU pool, so the memory checker for the device algorithm is the host
    build of the same sources with -fsanitize=address,undefined (emu_lib.SANITIZE): a slice of this file, the splice path
    and the SAM stage re-run in a child process under it (the whole three files pass under it too: DESIGN.md section 5)."""
    import os
    import subprocess
    import sys
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not installed")
    if os.environ.get("HSA_EMU_SANITIZE"):
        return                                                   # already inside the sanitized child
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, HSA_EMU_SANITIZE="1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-s", "-m", "not gpu", "-p", "no:cacheprovider",
                        os.path.join(here, "test_device_algorithm_emu.py"), os.path.join(here, "test_splice_emu.py"),
                        os.path.join(here, "test_sam_emu.py"),
                        "-k", "(ragged_nonstop and fast16) or (cooperative and cfg5) or (splice_match_vs_golden and 100) or "
                              "(sam_fields_vs_golden) or test_width or capacity_overflow"],
                       env=env, capture_output=True, text=True, cwd=os.path.dirname(here))
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and "AddressSanitizer" not in tail and "runtime error" not in tail, tail
    assert " passed" in r.stdout

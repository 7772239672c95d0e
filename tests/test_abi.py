"""CPU: the C-ABI library loads, exports every symbol include/hsa_b200.h declares, and its struct mirrors
have the reference's sizes.  No compute call is made (there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

from hsa_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hsa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hsa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = api.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"libhsa_b200.so does not export {s}"


def test_struct_sizes_match_reference_abi():
    # measured on the reference with gcc x86-64: sizeof(gap_opt_t)=64, bwt_aln1_t=36, bwt_width_t=8
    assert C.sizeof(api.GapOpt) == 64
    assert api.ALN_WORDS * 4 == 36
    assert C.sizeof(api.Task) == 40


def test_defaults_and_maxdiff(golden):
    o = api.gap_init_opt()
    assert (o.s_mm, o.s_gapo, o.s_gape, o.max_diff, o.max_gapo, o.max_gape) == (3, 11, 4, -1, 1, 6)
    assert (o.indel_end_skip, o.max_del_occ, o.max_entries, o.seed_len, o.max_seed_diff, o.max_top2) == (5, 10, 2000000, 32, 2, 30)
    assert o.mode == 3 and abs(o.fnr - 0.04) < 1e-7
    for L, v in golden.meta["maxdiff"].items():
        assert api.bwa_cal_maxdiff(int(L)) == v


def test_compute_fails_loudly_without_gpu(golden_index):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.HsaError):
        api.Index.upload(golden_index, 0)


def test_product_never_imports_oracle():
    """The product path must not reference oracle/ or the host emulation."""
    pkg = os.path.join(ROOT, "hsa_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert "oracle_lib" not in src and "liboracle" not in src and "hsa_emu" not in src, fn
                assert "emu_lib" not in src, fn

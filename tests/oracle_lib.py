"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE; never imported by hsa_b200/)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "hsa_ref")
REF_BIN_COUNT = os.path.join(ORACLE_DIR, "_ref", "hsa_ref_count")


class GapOpt(C.Structure):  # == hsa_gap_opt_t == gap_opt_t (bwtaln.h:133-143)
    _fields_ = [("s_mm", C.c_int), ("s_gapo", C.c_int), ("s_gape", C.c_int), ("mode", C.c_int),
                ("indel_end_skip", C.c_int), ("max_del_occ", C.c_int), ("max_entries", C.c_int),
                ("fnr", C.c_float), ("max_diff", C.c_int), ("max_gapo", C.c_int), ("max_gape", C.c_int),
                ("max_seed_diff", C.c_int), ("seed_len", C.c_int), ("n_threads", C.c_int),
                ("max_top2", C.c_int), ("trim_qual", C.c_int)]


class BwtView(C.Structure):  # == hsa_bwt_view_t
    _fields_ = [("textLength", C.c_uint32), ("inverseSa0", C.c_uint32), ("cumulativeFreq", C.c_uint32 * 5),
                ("bwtCode", C.c_void_p), ("bwtSizeInWord", C.c_uint32),
                ("occValue", C.c_void_p), ("occSizeInWord", C.c_uint32),
                ("occValueMajor", C.c_void_p), ("occMajorSizeInWord", C.c_uint32)]


class OracleIndex(C.Structure):
    _fields_ = [("fwd", BwtView), ("rev", BwtView)]


def _view(arr) -> BwtView:
    v = BwtView()
    v.textLength = arr.text_length
    v.inverseSa0 = arr.inverse_sa0
    for i in range(5):
        v.cumulativeFreq[i] = int(arr.cumulative_freq[i])
    v.bwtCode = arr.bwt_code.ctypes.data
    v.bwtSizeInWord = arr.bwt_code.shape[0]
    v.occValue = arr.occ_value.ctypes.data
    v.occSizeInWord = arr.occ_value.shape[0]
    v.occValueMajor = arr.occ_value_major.ctypes.data
    v.occMajorSizeInWord = arr.occ_value_major.shape[0]
    return v


def build_oracle() -> None:
    src = os.path.join(ORACLE_DIR, "hsa_oracle.c")
    if (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(LIB_PATH)
        L.hsao_occ4.argtypes = [C.POINTER(BwtView), C.c_uint32, C.POINTER(C.c_uint32)]
        L.hsao_occ1.argtypes = [C.POINTER(BwtView), C.c_uint32, C.c_uint32]
        L.hsao_occ1.restype = C.c_uint32
        L.hsao_sa_value.argtypes = [C.POINTER(BwtView), C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.hsao_sa_value.restype = C.c_uint32
        L.hsao_locate.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.hsao_cal_maxdiff.argtypes = [C.c_int, C.c_double, C.c_double]
        L.hsao_cal_maxdiff.restype = C.c_int
        L.hsao_gap_opt_default.argtypes = [C.POINTER(GapOpt)]
        L.hsao_cal_width.argtypes = [C.POINTER(OracleIndex), C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.hsao_cal_width.restype = C.c_int
        for name in ("hsao_percall", "hsao_whole"):
            f = getattr(L, name)
            f.argtypes = [C.POINTER(OracleIndex), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                          C.POINTER(GapOpt), C.c_int, C.c_void_p, C.POINTER(C.c_size_t)]
            f.restype = C.c_void_p
        L.hsao_seeds.argtypes = [C.POINTER(OracleIndex), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                 C.POINTER(GapOpt), C.c_void_p, C.POINTER(C.c_size_t)]
        L.hsao_seeds.restype = C.c_void_p
        L.hsao_free.argtypes = [C.c_void_p]
        L.hsao_occ4_calls.restype = C.c_uint64
        L.hsao_occ1_calls.restype = C.c_uint64
        _lib = L
    return _lib


def default_opt(**kw) -> GapOpt:
    o = GapOpt()
    lib().hsao_gap_opt_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Oracle:
    """The oracle over one loaded index (hsa_b200.index_io.Index2BWT)."""

    def __init__(self, index):
        self.index = index            # keeps the numpy arrays alive
        self.ix = OracleIndex(_view(index.fwd), _view(index.rev))

    def occ(self, which: int, indices: np.ndarray):
        L = lib()
        v = self.ix.fwd if which == 0 else self.ix.rev
        n = indices.shape[0]
        o4 = np.zeros((n, 4), dtype=np.uint32)
        o1 = np.zeros((n, 4), dtype=np.uint32)
        buf = (C.c_uint32 * 4)()
        for i, x in enumerate(indices.tolist()):
            L.hsao_occ4(C.byref(v), x, buf)
            o4[i] = buf[:]
            for c in range(4):
                o1[i, c] = L.hsao_occ1(C.byref(v), x, c)
        return o4, o1

    def sa_values(self, indices: np.ndarray):
        """(SA values, PsiMinus steps walked) of SA indices on the forward BWT (BWTSaValue)."""
        L = lib()
        b = self.index.fwd
        out = np.zeros(indices.shape[0], dtype=np.uint32)
        steps = np.zeros(indices.shape[0], dtype=np.uint32)
        st = C.c_uint32(0)
        for i, x in enumerate(indices.tolist()):
            out[i] = L.hsao_sa_value(C.byref(self.ix.fwd), b.sa_value.ctypes.data, b.sa_interval, x, C.byref(st))
            steps[i] = st.value
        return out, steps

    def locate(self, indices: np.ndarray, blocks):
        """BWTRetrievePositionFromSAIndex: rows {occ_pos, seq_id, ori_pos} (-1, -1 where no block holds the position)."""
        L = lib()
        pos, _ = self.sa_values(indices)
        t = blocks.table()
        out = np.full((indices.shape[0], 3), 0xFFFFFFFF, dtype=np.uint32)
        out[:, 0] = pos
        a, b = C.c_uint32(0), C.c_uint32(0)
        for i, x in enumerate(pos.tolist()):
            if L.hsao_locate(t.ctypes.data, t.shape[0], x, C.byref(a), C.byref(b)):
                out[i, 1], out[i, 2] = a.value, b.value
        return out

    def cal_width(self, seq: np.ndarray, type_: int = 1):
        L = lib()
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        w = np.zeros((seq.shape[0] + 1, 2), dtype=np.uint32)
        bid = L.hsao_cal_width(C.byref(self.ix), seq.shape[0], seq.ctypes.data, w.ctypes.data, type_)
        return bid, w

    def _batch(self, fn, rs, opt, extra):
        L = lib()
        codes = np.ascontiguousarray(rs.codes, dtype=np.uint8)
        off = np.ascontiguousarray(rs.offsets[:-1], dtype=np.uint64)
        lens = np.ascontiguousarray(rs.lens, dtype=np.uint32)
        mult = {"hsao_percall": 2, "hsao_whole": 1, "hsao_seeds": 6}[fn]
        n_aln = np.zeros(rs.n * mult, dtype=np.int32)
        total = C.c_size_t(0)
        L.hsao_reset_counters()
        args = [C.byref(self.ix), codes.ctypes.data, off.ctypes.data, lens.ctypes.data, rs.n, C.byref(opt)]
        args += extra + [n_aln.ctypes.data, C.byref(total)]
        p = getattr(L, fn)(*args)
        rows = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(max(total.value, 1) * 12,))
        rows = rows[: total.value * 12].reshape(total.value, 12).copy()
        L.hsao_free(p)
        self.last_lookups = int(L.hsao_occ4_calls()) + int(L.hsao_occ1_calls())
        return n_aln, rows

    def percall(self, rs, opt, clear_gape: int = 1):
        return self._batch("hsao_percall", rs, opt, [clear_gape])

    def whole(self, rs, opt, clear_gape: int = 1):
        return self._batch("hsao_whole", rs, opt, [clear_gape])

    def seeds(self, rs, opt):
        return self._batch("hsao_seeds", rs, opt, [])


def have_ref() -> bool:
    return os.path.exists(REF_BIN)


def run_ref(args, count: bool = False, cwd=None) -> dict:
    """Run the compiled reference harness; returns its one-line JSON."""
    import json
    out = subprocess.run([REF_BIN_COUNT if count else REF_BIN] + [str(a) for a in args], cwd=cwd,
                         check=True, capture_output=True, text=True).stdout.strip().splitlines()
    return json.loads(out[-1])


def opt_args(opt: GapOpt):
    """key=value arguments that make ref_harness.c's gap_opt_t equal to `opt`."""
    keys = ["s_mm", "s_gapo", "s_gape", "mode", "indel_end_skip", "max_del_occ", "max_entries", "max_diff",
            "max_gapo", "max_gape", "max_seed_diff", "seed_len", "max_top2"]
    return [f"{k}={getattr(opt, k)}" for k in keys] + [f"fnr={opt.fnr:.9g}"]

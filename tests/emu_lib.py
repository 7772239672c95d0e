"""ctypes binding of tests/emu/libhsa_emu.so: the device algorithm (hsa_core.cuh) compiled for the host.

TEST INFRASTRUCTURE ONLY -- lets the CPU suite check the CUDA worker's logic against the oracle without
a GPU.  The product never loads it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
# HSA_EMU_SANITIZE=1: the same sources with -fsanitize=address,undefined (run the suite with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
# ASAN_OPTIONS=detect_leaks=0): the only memory checker this pool offers for the device algorithm -- compute-sanitizer is closed
SANITIZE = os.environ.get("HSA_EMU_SANITIZE", "") not in ("", "0")
LIB = os.path.join(EMU_DIR, "libhsa_emu_asan.so" if SANITIZE else "libhsa_emu.so")
CORE = os.path.join(ol.ROOT, "hsa_b200", "csrc", "hsa_core.cuh")


class Task(C.Structure):  # == hsa_task_t
    _fields_ = [("read_off", C.c_uint64), ("read_len", C.c_uint32), ("strand", C.c_uint32),
                ("sub_off", C.c_uint32), ("len", C.c_uint32), ("wsrc_off", C.c_uint32),
                ("seed_mode", C.c_uint32), ("opt_idx", C.c_uint32), ("reserved", C.c_uint32)]


def build():
    src = os.path.join(EMU_DIR, "hsa_emu.cpp")
    csrc = os.path.dirname(CORE)
    deps = [src, os.path.join(ol.ROOT, "include", "hsa_b200.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    newest = max(os.path.getmtime(d) for d in deps)
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < newest:
        extra = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"] if SANITIZE else ["-O2"]
        subprocess.check_call(["g++"] + extra + ["-fPIC", "-shared", "-std=c++17", "-o", LIB, src])


_lib = None


def _padded(codes):
    """The read codes with 16 readable bytes behind them, as the product's device copy has (hsa_b200.cu: upload_reads):
    the device code loads the bases as aligned 32-bit words (include/hsa_b200.h)."""
    c = np.zeros(codes.shape[0] + 16, dtype=np.uint8)
    c[: codes.shape[0]] = codes
    return c


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.emu_index_new.restype = C.c_void_p
        L.emu_index_new.argtypes = [C.POINTER(ol.BwtView), C.POINTER(ol.BwtView)]
        L.emu_index_free.argtypes = [C.c_void_p]
        L.emu_occ.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(ol.BwtView), C.c_void_p, C.c_size_t, C.c_void_p]
        L.emu_sa_values.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.emu_run.restype = C.c_long
        L.emu_set_rerun.argtypes = [C.c_uint32]
        L.emu_set_coop.argtypes = [C.c_int, C.c_uint32]
        L.emu_coop_waves.restype = C.c_uint64
        L.emu_flagged_first.restype = C.c_uint64
        L.emu_run.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                              C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int32, C.c_uint32, C.c_uint32,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                              C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.emu_splice.restype = C.c_uint64
        L.emu_splice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_uint32,
                                 C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def resolve_read_opt(opt: ol.GapOpt, length: int, clear_gape: int = 1) -> ol.GapOpt:
    """Per-read option resolution of bwa_cal_sa_reg_gap before any local_opt switch (bwtaln.c:260-261, 330-332)."""
    o = ol.GapOpt.from_buffer_copy(bytes(opt))
    if clear_gape:
        o.mode &= ~1
    if opt.fnr > 0:
        o.max_diff = ol.lib().hsao_cal_maxdiff(length, 0.02, opt.fnr)
    o.seed_len = opt.seed_len if opt.seed_len < length else 0x7FFFFFFF
    return o


def aln9_to_rows12(aln9: np.ndarray) -> np.ndarray:
    """hsa_aln1_t words -> the 12-word dump layout of oracle/ref_harness.c."""
    n = aln9.shape[0]
    r = np.zeros((n, 12), dtype=np.uint32)
    r[:, 0] = aln9[:, 0] & 0xFFFF
    r[:, 1] = (aln9[:, 0] >> 16) & 0xFF
    r[:, 2] = (aln9[:, 0] >> 24) & 0xFF
    r[:, 3:7] = aln9[:, 1:5]
    r[:, 7] = aln9[:, 5] & 0x3FFFFFFF
    r[:, 8] = aln9[:, 5] >> 30
    r[:, 9] = aln9[:, 6]
    r[:, 10] = aln9[:, 7]
    r[:, 11] = aln9[:, 8]
    return r


def gather_rows(n_aln: np.ndarray, aln_off: np.ndarray, aln9: np.ndarray) -> np.ndarray:
    """Hits of all items in item order (the arena order is arbitrary on the GPU)."""
    idx = [np.arange(int(o), int(o) + int(c)) for c, o in zip(n_aln.tolist(), aln_off.tolist()) if c]
    if not idx:
        return np.zeros((0, 12), dtype=np.uint32)
    return aln9_to_rows12(aln9[np.concatenate(idx)])


class Emu:
    def __init__(self, index):
        self.index = index
        self.vf, self.vr = ol._view(index.fwd), ol._view(index.rev)
        self.h = lib().emu_index_new(C.byref(self.vf), C.byref(self.vr))

    def __del__(self):
        try:
            lib().emu_index_free(self.h)
        except Exception:
            pass

    def occ(self, which: int, layout: int, idx: np.ndarray) -> np.ndarray:
        out = np.zeros((idx.shape[0], 4), dtype=np.uint32)
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        lib().emu_occ(self.h, which, layout, C.byref(self.vf if which == 0 else self.vr), idx.ctypes.data,
                      idx.shape[0], out.ctypes.data)
        return out

    @staticmethod
    def locate(blocks, pos: np.ndarray):
        t = blocks.table()
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        a = np.zeros(pos.shape[0], dtype=np.uint32)
        b = np.zeros(pos.shape[0], dtype=np.uint32)
        lib().emu_locate(C.c_void_p(t.ctypes.data), C.c_uint32(t.shape[0]), C.c_void_p(pos.ctypes.data), C.c_size_t(pos.shape[0]),
                         C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data))
        return a, b

    def splice(self, rs, opts, opt_idx=None, arena_cap=1 << 16, aln_cap=512):
        """bwt_splice_match for every read (hsa_splice.cuh on the host): (n_aln[n], rows12 of the n_aln entries in read
        order, status[n]).  Needs index.fwd.sa_value, index.blocks and index.packed_dna."""
        ix = self.index
        opts = list(opts) if isinstance(opts, (list, tuple)) else [opts]
        oa = (ol.GapOpt * len(opts))(*opts)
        t = np.ascontiguousarray(ix.blocks.table(), dtype=np.uint32)
        off = np.ascontiguousarray(rs.offsets[:-1], dtype=np.uint64)
        lens = np.ascontiguousarray(rs.lens, dtype=np.uint32)
        n = rs.n
        n_aln = np.zeros(n, dtype=np.int32)
        aln = np.zeros((n, 2, 9), dtype=np.uint32)
        status = np.zeros(n, dtype=np.uint8)
        oi = None if opt_idx is None else np.ascontiguousarray(opt_idx, dtype=np.uint32)
        codes = _padded(rs.codes)
        self.last_lookups = lib().emu_splice(self.h, ix.fwd.sa_value.ctypes.data, ix.fwd.sa_interval, t.ctypes.data, t.shape[0],
                                             ix.packed_dna.ctypes.data, ix.dna_length, codes.ctypes.data, off.ctypes.data,
                                             lens.ctypes.data, n, C.cast(oa, C.c_void_p), len(opts),
                                             None if oi is None else oi.ctypes.data, arena_cap, aln_cap,
                                             n_aln.ctypes.data, aln.ctypes.data, status.ctypes.data)
        keep = np.arange(2)[None, :] < n_aln[:, None]
        return n_aln, aln9_to_rows12(aln[keep]), status

    def sa_values(self, idx: np.ndarray):
        b = self.index.fwd
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        out = np.zeros(idx.shape[0], dtype=np.uint32)
        steps = np.zeros(idx.shape[0], dtype=np.uint32)
        lib().emu_sa_values(self.h, b.sa_value.ctypes.data, b.sa_interval, idx.ctypes.data, idx.shape[0], out.ctypes.data,
                            steps.ctypes.data)
        return out, steps

    def run(self, kind, rs, opts, tasks=None, len2opt=None, filter_max_n=0, arena_cap=1022, hit_cap=32,
            n_items=None, want_width=False, rerun_cap=0, coop=False, step_budget=0):
        """rerun_cap: 0 = items the configuration cannot hold are reported in `status` (1); otherwise they are
        re-run with that arena capacity (the large-capacity configuration), as the product's host code does."""
        codes = _padded(rs.codes)
        off = np.ascontiguousarray(rs.offsets[:-1], dtype=np.uint64)
        lens = np.ascontiguousarray(rs.lens, dtype=np.uint32)
        max_len = int(lens.max()) if rs.n else 0
        n_groups = rs.n if tasks is None else len(tasks)
        if n_items is None:
            n_items = n_groups * (6 if kind == 2 else 1)
        n_aln = np.zeros(n_items, dtype=np.int32)
        aln_off = np.zeros(n_items, dtype=np.uint64)
        status = np.full(n_items, 0xFF, dtype=np.uint8)
        cap = max(n_items * 8, 1024)
        aln = np.zeros((cap, 9), dtype=np.uint32)
        optarr = (ol.GapOpt * len(opts))(*opts)
        tarr = (Task * len(tasks))(*tasks) if tasks is not None else None
        l2o = np.ascontiguousarray(len2opt, dtype=np.uint16) if len2opt is not None else None
        wout = np.zeros((int(lens.sum()) + rs.n, 2), dtype=np.uint32) if want_width else None
        bid = np.zeros(rs.n, dtype=np.int32) if want_width else None
        lk, ns, pops = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        lib().emu_set_rerun(C.c_uint32(rerun_cap))
        lib().emu_set_coop(C.c_int(1 if coop else 0), C.c_uint32(step_budget))
        total = lib().emu_run(self.h, kind, codes.ctypes.data, C.cast(tarr, C.c_void_p) if tarr is not None else None,
                              off.ctypes.data, lens.ctypes.data, n_groups, C.cast(optarr, C.c_void_p), len(opts),
                              l2o.ctypes.data if l2o is not None else None, max_len, filter_max_n, arena_cap, hit_cap,
                              n_aln.ctypes.data, aln_off.ctypes.data, status.ctypes.data, aln.ctypes.data, cap,
                              wout.ctypes.data if want_width else None, bid.ctypes.data if want_width else None,
                              C.byref(lk), C.byref(ns), C.byref(pops))
        assert total >= 0
        self.last_lookups, self.last_strict, self.last_pops = lk.value, ns.value, pops.value
        self.last_flagged_first = int(lib().emu_flagged_first())
        if want_width:
            return bid, wout
        return n_aln, aln_off, status, aln[:total]

    # convenience wrappers with the same semantics as Oracle.percall / whole / seeds
    def percall(self, rs, opt, clear_gape=1, **kw):
        lens = sorted(set(rs.lens.tolist()))
        opts = [resolve_read_opt(opt, L, clear_gape) for L in lens]
        oi = {L: i for i, L in enumerate(lens)}
        off = rs.offsets
        tasks = []
        for r in range(rs.n):
            L = int(rs.lens[r])
            sm = 1 if L > opts[oi[L]].seed_len else 0
            for s in (1, 0):
                tasks.append(Task(int(off[r]), L, s, 0, L, 0, sm, oi[L], 0))
        n_aln, aln_off, status, aln = self.run(0, rs, opts, tasks=tasks, **kw)
        return n_aln, gather_rows(n_aln, aln_off, aln), status

    def whole(self, rs, opt, clear_gape=1, **kw):
        lens = sorted(set(rs.lens.tolist()))
        opts = [resolve_read_opt(opt, L, clear_gape) for L in lens]
        max_len = max(lens) if lens else 0
        l2o = np.zeros(max_len + 1, dtype=np.uint16)
        for i, L in enumerate(lens):
            l2o[L] = i
        fmax = ol.lib().hsao_cal_maxdiff(max_len, 0.02, opt.fnr) if opt.fnr > 0 else opt.max_diff
        n_aln, aln_off, status, aln = self.run(1, rs, opts, len2opt=l2o, filter_max_n=fmax, **kw)
        return n_aln, gather_rows(n_aln, aln_off, aln), status

    def seeds(self, rs, opt, **kw):
        so = ol.GapOpt.from_buffer_copy(bytes(opt))
        so.mode &= ~1
        so.max_gapo = 0
        so.max_gape = 0
        so.max_diff = opt.max_seed_diff
        n_aln, aln_off, status, aln = self.run(2, rs, [so], **kw)
        return n_aln, gather_rows(n_aln, aln_off, aln), status

    def width(self, rs):
        return self.run(3, rs, [ol.default_opt()], want_width=True)

"""Python host mirror of the reference's interface for the inexact-search path, over the C ABI.

Names follow the reference (gap_opt_t -> GapOpt, bwt_aln1_t -> numpy structured ALN_DTYPE,
bwa_cal_sa_reg_gap's whole-read part -> Index.whole_reads, bwt_match_gap -> Index.match_gap_batch,
bwt_splice_match's seed calls -> Index.splice_seeds, bwt_cal_width -> Index.cal_width,
BWTAllOccValue -> Index.occ).  All compute goes through `libhsa_b200.so` (hand-written CUDA for
sm_100a); if the library is missing or there is no CUDA device the calls raise -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# HSA_B200_LIB selects another build of the same library (e.g. the -DHSA_PHASE_PROF diagnostics build)
LIB_PATH = os.environ.get("HSA_B200_LIB") or os.path.join(_HERE, "libhsa_b200.so")

MODE_GAPE, MODE_COMPREAD, MODE_LOGGAP, MODE_NONSTOP = 0x01, 0x02, 0x04, 0x10
SEED_NONE, SEED_TAIL, SEED_ALIAS = 0, 1, 2


class HsaError(RuntimeError):
    pass


class GapOpt(C.Structure):
    """gap_opt_t (bwtaln.h:133-143)."""
    _fields_ = [("s_mm", C.c_int), ("s_gapo", C.c_int), ("s_gape", C.c_int), ("mode", C.c_int),
                ("indel_end_skip", C.c_int), ("max_del_occ", C.c_int), ("max_entries", C.c_int),
                ("fnr", C.c_float), ("max_diff", C.c_int), ("max_gapo", C.c_int), ("max_gape", C.c_int),
                ("max_seed_diff", C.c_int), ("seed_len", C.c_int), ("n_threads", C.c_int),
                ("max_top2", C.c_int), ("trim_qual", C.c_int)]


class BwtView(C.Structure):
    _fields_ = [("textLength", C.c_uint32), ("inverseSa0", C.c_uint32), ("cumulativeFreq", C.c_uint32 * 5),
                ("bwtCode", C.c_void_p), ("bwtSizeInWord", C.c_uint32),
                ("occValue", C.c_void_p), ("occSizeInWord", C.c_uint32),
                ("occValueMajor", C.c_void_p), ("occMajorSizeInWord", C.c_uint32)]


class Task(C.Structure):
    """hsa_task_t: one bwt_match_gap call."""
    _fields_ = [("read_off", C.c_uint64), ("read_len", C.c_uint32), ("strand", C.c_uint32),
                ("sub_off", C.c_uint32), ("len", C.c_uint32), ("wsrc_off", C.c_uint32),
                ("seed_mode", C.c_uint32), ("opt_idx", C.c_uint32), ("reserved", C.c_uint32)]


TASK_DTYPE = np.dtype([("read_off", "<u8"), ("read_len", "<u4"), ("strand", "<u4"), ("sub_off", "<u4"),
                       ("len", "<u4"), ("wsrc_off", "<u4"), ("seed_mode", "<u4"), ("opt_idx", "<u4"),
                       ("reserved", "<u4")])
assert TASK_DTYPE.itemsize == C.sizeof(Task) == 40


class _Result(C.Structure):
    _fields_ = [("n_items", C.c_size_t), ("n_aln", C.POINTER(C.c_int32)), ("aln_off", C.POINTER(C.c_uint64)),
                ("aln", C.c_void_p), ("n_aln_total", C.c_size_t), ("occ_lookups", C.c_uint64),
                ("n_strict", C.c_uint64), ("pops", C.c_uint64), ("steps", C.c_uint64),
                ("kernel_ms", C.c_float), ("kernel_launches", C.c_uint32),
                ("cap_items", C.c_size_t), ("cap_aln", C.c_size_t)]


# bwt_aln1_t (bwtaln.h:41-50) as 9 little-endian words; bit-fields unpacked by helpers below
ALN_WORDS = 9


class _SamResult(C.Structure):        # == hsa_sam_result_t
    _fields_ = [("n_reads", C.c_size_t), ("rec", C.c_void_p), ("multi", C.c_void_p), ("n_multi", C.c_size_t),
                ("cigar", C.c_void_p), ("n_cigar", C.c_size_t), ("md", C.c_void_p), ("md_bytes", C.c_size_t),
                ("n_refined", C.c_uint64), ("kernel_ms", C.c_float),
                ("cap_rec", C.c_size_t), ("cap_multi", C.c_size_t), ("cap_cigar", C.c_size_t), ("cap_md", C.c_size_t)]


class _SamDevice(C.Structure):        # == hsa_sam_device_t
    _fields_ = [("rec_dev", C.c_void_p), ("multi_dev", C.c_void_p), ("cigar_dev", C.c_void_p), ("md_dev", C.c_void_p),
                ("n_multi", C.c_size_t), ("n_cigar", C.c_size_t), ("md_bytes", C.c_size_t),
                ("n_refined", C.c_uint64), ("n_several_best", C.c_uint64), ("kernel_ms", C.c_float)]


SAM_REC_WORDS, SAM_MULTI_WORDS = 22, 12      # hsa_sam1_t / hsa_multi1_t in 32-bit words


def aln_fields(aln9: np.ndarray) -> dict:
    """Unpack hsa_aln1_t words into named uint32/int32 columns."""
    return dict(n_mm=aln9[:, 0] & 0xFFFF, n_gapo=(aln9[:, 0] >> 16) & 0xFF, n_gape=(aln9[:, 0] >> 24) & 0xFF,
                k=aln9[:, 1], l=aln9[:, 2], rev_k=aln9[:, 3], rev_l=aln9[:, 4], type=aln9[:, 5] & 0x3FFFFFFF,
                strand=aln9[:, 5] >> 30, start=aln9[:, 6].view(np.int32), end=aln9[:, 7].view(np.int32),
                score=aln9[:, 8].view(np.int32))


_lib = None


def lib():
    """Load the CUDA library; raises if it has not been built (python -m hsa_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HsaError(f"{LIB_PATH} is missing: build it with `python -m hsa_b200.build` "
                       "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.hsa_b200_abi_version.restype = C.c_int
    L.hsa_last_error.restype = C.c_char_p
    L.hsa_gap_opt_default.argtypes = [C.POINTER(GapOpt)]
    L.hsa_cal_maxdiff.argtypes = [C.c_int, C.c_double, C.c_double]
    L.hsa_cal_maxdiff.restype = C.c_int
    L.hsa_index_upload.argtypes = [C.c_int, C.POINTER(BwtView), C.POINTER(BwtView), C.POINTER(C.c_void_p)]
    L.hsa_index_from_device.argtypes = [C.c_int, C.POINTER(BwtView), C.POINTER(BwtView), C.POINTER(C.c_void_p)]
    L.hsa_index_blocks.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.hsa_index_from_blocks.argtypes = [C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p,
                                        C.c_int, C.POINTER(C.c_void_p)]
    L.hsa_index_meta.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32)]
    L.hsa_index_free.argtypes = [C.c_void_p]
    L.hsa_occ_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.hsa_cal_width_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                      C.c_void_p, C.c_void_p]
    L.hsa_match_gap_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                      C.c_size_t, C.POINTER(_Result)]
    L.hsa_whole_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(GapOpt),
                                  C.c_int, C.POINTER(_Result)]
    L.hsa_whole_reads_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(GapOpt),
                                         C.c_int, C.POINTER(C.c_void_p)]
    L.hsa_job_wait.argtypes = [C.c_void_p, C.POINTER(_Result)]
    L.hsa_splice_seeds_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(GapOpt),
                                          C.POINTER(C.c_void_p)]
    L.hsa_splice_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(GapOpt),
                                   C.POINTER(_Result)]
    L.hsa_result_free.argtypes = [C.POINTER(_Result)]
    L.hsa_workspace_create.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_size_t, C.POINTER(C.c_void_p)]
    L.hsa_workspace_free.argtypes = [C.c_void_p]
    L.hsa_workspace_check.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.hsa_workspace_last_launches.argtypes = [C.c_void_p]
    L.hsa_workspace_last_launches.restype = C.c_uint32
    L.hsa_workspace_last_config.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    L.hsa_workspace_last_config.restype = None
    L.hsa_workspace_launch_timing.argtypes = [C.c_void_p, C.c_int]
    L.hsa_workspace_launch_times.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_float), C.c_size_t,
                                             C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
    L.hsa_whole_reads_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.c_void_p, C.c_size_t, C.POINTER(GapOpt), C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.hsa_random_sector_probe.argtypes = [C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
    L.hsa_random_sector_probe_ex.argtypes = [C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_double), C.c_int]
    L.hsa_index_attach_sa.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32]
    L.hsa_sa_values.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_uint64)]
    L.hsa_sa_values_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hsa_index_attach_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.hsa_sa_locate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hsa_index_attach_packed_dna.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.hsa_splice_match_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.hsa_sam_se_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(GapOpt), C.c_int, C.POINTER(C.c_uint64), C.POINTER(_SamResult)]
    L.hsa_sam_result_free.argtypes = [C.POINTER(_SamResult)]
    L.hsa_copy_from_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.hsa_sam_se_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(GapOpt), C.c_int, C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(_SamDevice)]
    L.hsa_sam_format.argtypes = [C.POINTER(_SamResult), C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.POINTER(C.c_char_p), C.c_size_t, C.POINTER(GapOpt), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.hsa_match_gap_call.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(GapOpt),
                                     C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise HsaError(f"hsa_b200 error {rc}: {lib().hsa_last_error().decode()}")


def gap_init_opt(**overrides) -> GapOpt:
    """gap_init_opt (bwtaln.c:21-44) with keyword overrides."""
    o = GapOpt()
    lib().hsa_gap_opt_default(C.byref(o))
    for k, v in overrides.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def bwa_cal_maxdiff(length: int, err: float = 0.02, thres: float = 0.04) -> int:
    """bwa_cal_maxdiff (bwtaln.c:46-58)."""
    return lib().hsa_cal_maxdiff(length, err, thres)


def _view_host(arr) -> BwtView:
    v = BwtView()
    v.textLength, v.inverseSa0 = arr.text_length, arr.inverse_sa0
    for i in range(5):
        v.cumulativeFreq[i] = int(arr.cumulative_freq[i])
    v.bwtCode, v.bwtSizeInWord = arr.bwt_code.ctypes.data, arr.bwt_code.shape[0]
    v.occValue, v.occSizeInWord = arr.occ_value.ctypes.data, arr.occ_value.shape[0]
    v.occValueMajor, v.occMajorSizeInWord = arr.occ_value_major.ctypes.data, arr.occ_value_major.shape[0]
    return v


class BatchResult:
    """Flat result of a batch: per item n_aln and its hits (rows of 9 hsa_aln1_t words), in item order."""

    def __init__(self, r: _Result, copy: bool = True):
        n = r.n_items
        tot = r.n_aln_total
        n_aln = np.ctypeslib.as_array(r.n_aln, shape=(max(n, 1),))[:n]
        off = np.ctypeslib.as_array(r.aln_off, shape=(max(n, 1),))[:n]
        if tot:
            aln = np.ctypeslib.as_array(C.cast(r.aln, C.POINTER(C.c_uint32)), shape=(tot * ALN_WORDS,))
            aln = aln.reshape(tot, ALN_WORDS)
        else:
            aln = np.zeros((0, ALN_WORDS), dtype=np.uint32)
        self.n_aln = n_aln.copy() if copy else n_aln
        self.aln_off = off.copy() if copy else off
        self.aln = aln.copy() if copy else aln
        self.occ_lookups = int(r.occ_lookups)
        self.n_strict = int(r.n_strict)
        self.pops, self.steps = int(r.pops), int(r.steps)
        self.kernel_ms = float(r.kernel_ms)
        self.kernel_launches = int(r.kernel_launches)

    def item(self, i: int) -> np.ndarray:
        o, c = int(self.aln_off[i]), int(self.n_aln[i])
        return self.aln[o:o + c]

    def ordered(self) -> np.ndarray:
        """All hits concatenated in item order (the device arena order is arbitrary)."""
        if self.aln.shape[0] == 0:
            return self.aln
        nz = np.nonzero(self.n_aln)[0]
        cnt = self.n_aln[nz].astype(np.int64)
        starts = self.aln_off[nz].astype(np.int64)
        idx = np.repeat(starts - np.concatenate([[0], np.cumsum(cnt)[:-1]]), cnt) + np.arange(int(cnt.sum()))
        return self.aln[idx]


class Index:
    """Device-resident 2BWT search arrays (replaces the in-memory result of BWTLoad2BWT, 2BWT-Interface.c:13)."""

    def __init__(self, handle: int, device: int):
        self._h = C.c_void_p(handle)
        self.device = device
        self._res = _Result()
        self._job_res = [_Result() for _ in range(4)]     # result buffers of the asynchronous jobs, round-robin
        self._job_seq = 0

    @classmethod
    def upload(cls, index2bwt, device: int = 0) -> "Index":
        """From host arrays (hsa_b200.index_io.Index2BWT: loaded reference files or index_build output)."""
        vf, vr = _view_host(index2bwt.fwd), _view_host(index2bwt.rev)
        h = C.c_void_p()
        _check(lib().hsa_index_upload(device, C.byref(vf), C.byref(vr), C.byref(h)))
        ix = cls(h.value, device)
        if getattr(index2bwt.fwd, "sa_value", None) is not None:
            ix.attach_sa(index2bwt.fwd.sa_value, index2bwt.fwd.sa_interval)
        if getattr(index2bwt, "blocks", None) is not None:
            ix.attach_blocks(index2bwt.blocks.table())
        if getattr(index2bwt, "packed_dna", None) is not None:
            ix.attach_packed_dna(index2bwt.packed_dna, index2bwt.dna_length)
        return ix

    @classmethod
    def from_prefix(cls, prefix: str, device: int = 0) -> "Index":
        from .index_io import load_index
        return cls.upload(load_index(prefix), device)

    @classmethod
    def from_blocks(cls, meta_fwd, meta_rev, blocks_fwd_ptr: int, blocks_rev_ptr: int, device: int) -> "Index":
        """Wrap device-layout blocks that already live on `device` (e.g. received by an NCCL broadcast).
        The caller keeps the memory alive."""
        mf = (C.c_uint32 * 7)(*[int(x) for x in meta_fwd])
        mr = (C.c_uint32 * 7)(*[int(x) for x in meta_rev])
        h = C.c_void_p()
        _check(lib().hsa_index_from_blocks(device, mf, mr, blocks_fwd_ptr, blocks_rev_ptr, 0, C.byref(h)))
        return cls(h.value, device)

    def meta(self, which: int):
        m = (C.c_uint32 * 7)()
        _check(lib().hsa_index_meta(self._h, which, m))
        return list(m)

    def blocks(self, which: int):
        p, b = C.c_void_p(), C.c_size_t()
        _check(lib().hsa_index_blocks(self._h, which, C.byref(p), C.byref(b)))
        return p.value, b.value

    def close(self):
        if self._h:
            lib().hsa_result_free(C.byref(self._res))
            for r in self._job_res:
                lib().hsa_result_free(C.byref(r))
            lib().hsa_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- SA index -> text position -------------------------------------------------------------------
    def attach_sa(self, sa_value: np.ndarray, sa_interval: int) -> None:
        """The forward BWT's loaded saValue array (BWT.c:205-223), copied to the device once."""
        sa = np.ascontiguousarray(sa_value, dtype=np.uint32)
        _check(lib().hsa_index_attach_sa(self._h, sa.ctypes.data, sa.shape[0], int(sa_interval)))

    def sa_values(self, sa_indices: np.ndarray) -> np.ndarray:
        """BWTSaValue (BWT.c:1195-1225) for every SA index; self.last_sa_steps = PsiMinus steps walked in total."""
        idx = np.ascontiguousarray(sa_indices, dtype=np.uint32)
        out = np.zeros(idx.shape[0], dtype=np.uint32)
        st = C.c_uint64(0)
        _check(lib().hsa_sa_values(self._h, idx.ctypes.data, idx.shape[0], out.ctypes.data, C.byref(st)))
        self.last_sa_steps = int(st.value)
        return out

    def attach_blocks(self, blocks4: np.ndarray) -> None:
        """HSP::blockList rows {chrID, blockStart, blockEnd, ori} (index_io.Blocks.table())."""
        t = np.ascontiguousarray(blocks4, dtype=np.uint32)
        _check(lib().hsa_index_attach_blocks(self._h, t.ctypes.data, t.shape[0]))

    def sa_locate(self, sa_indices: np.ndarray) -> np.ndarray:
        """BWTRetrievePositionFromSAIndex (2BWT-Interface.c:329-362): rows {occ_pos, seq_id, ori_pos}."""
        idx = np.ascontiguousarray(sa_indices, dtype=np.uint32)
        o = [np.zeros(idx.shape[0], dtype=np.uint32) for _ in range(3)]
        _check(lib().hsa_sa_locate(self._h, idx.ctypes.data, idx.shape[0], o[0].ctypes.data, o[1].ctypes.data, o[2].ctypes.data))
        return np.stack(o, axis=1)

    def attach_packed_dna(self, packed_dna: np.ndarray, dna_length: int) -> None:
        """HSP::packedDNA / HSP::dnaLength (index_io.load_pac or index_io.pack_dna): the text the splice path scans."""
        p = np.ascontiguousarray(packed_dna, dtype=np.uint32)
        if p.shape[0] < (int(dna_length) + 15) // 16:
            raise HsaError("packed_dna shorter than dna_length / 16 words")
        _check(lib().hsa_index_attach_packed_dna(self._h, p.ctypes.data, int(dna_length)))

    def splice_match(self, codes, off, lens, opts, opt_idx=None):
        """bwt_splice_match (bwtgap.c:748) for every read: (n_aln[n] in 0..2, aln[n, 2, 9] hsa_aln1_t words).  opts = one
        GapOpt or a list of them (the aux->opt the driver holds per read); opt_idx[n] selects per read."""
        opts = list(opts) if isinstance(opts, (list, tuple)) else [opts]
        oa = (GapOpt * len(opts))(*opts)
        cp, op, lp, n = _ptr(codes, np.uint8), _ptr(off, np.uint64), _ptr(lens, np.uint32), _len(lens)
        oi = None if opt_idx is None else np.ascontiguousarray(opt_idx, dtype=np.uint32)
        n_aln = np.zeros(n, dtype=np.int32)
        aln = np.zeros((n, 2, ALN_WORDS), dtype=np.uint32)
        lk = C.c_uint64(0)
        _check(lib().hsa_splice_match_batch(self._h, cp[0], op[0], lp[0], n, C.cast(oa, C.c_void_p), len(opts),
                                            None if oi is None else oi.ctypes.data, n_aln.ctypes.data, aln.ctypes.data, C.byref(lk)))
        self.last_splice_lookups = int(lk.value)
        return n_aln, aln

    def sam_se(self, codes, off, lens, n_aln, aln_off, aln9, opt: GapOpt, n_occ: int = 3, rng48_state: int = 0,
               copy: bool = True, into: "SamResult | None" = None) -> "SamResult":
        """What generate_sam_se_core (bwtse.c:884) computes for a batch before printing: hit selection (host, the reference's
        drand48 stream from `rng48_state`), then on the GPU positions, pairing of spliced parts, banded DP -> CIGAR, MD / NM.
        n_aln / aln_off / aln9 = the reads' hits as hsa_whole_reads (+ hsa_splice_match_batch) leave them."""
        cp, op, lp, n = _ptr(codes, np.uint8), _ptr(off, np.uint64), _ptr(lens, np.uint32), _len(lens)
        na = np.ascontiguousarray(n_aln, dtype=np.int32)
        ao = np.ascontiguousarray(aln_off, dtype=np.uint64)
        a9 = np.ascontiguousarray(aln9, dtype=np.uint32)
        res = _SamResult()
        if into is not None and into._r is not None:       # re-use the library-managed host arrays of an earlier result
            res, into._r = into._r, None
        st = C.c_uint64(rng48_state)
        rc = lib().hsa_sam_se_batch(self._h, cp[0], op[0], lp[0], n, na.ctypes.data, ao.ctypes.data, a9.ctypes.data, C.byref(opt),
                                    n_occ, C.byref(st), C.byref(res))
        if rc:
            lib().hsa_sam_result_free(C.byref(res))
            _check(rc)
        return SamResult(res, st.value, (cp[1], op[1], lp[1]), opt, copy)

    def copy_from_device(self, ptr: int, shape, dtype) -> np.ndarray:
        """numpy copy of a device array the library returned a raw pointer to."""
        out = np.zeros(shape, dtype=dtype)
        _check(lib().hsa_copy_from_device(self._h, out.ctypes.data, ptr, out.nbytes))
        return out

    def sam_se_device(self, codes_ptr: int, off_ptr: int, len_ptr: int, n_reads: int, max_len: int, n_aln_ptr: int, aln_off_ptr: int,
                      aln_ptr: int, opt: GapOpt, n_occ: int = 3, rng48_state: int = 0, stream_ptr: int = 0):
        """hsa_sam_se_device: reads and hits resident in HBM (raw device pointers), results left on the device.  Returns
        (_SamDevice with the result pointers / counts, new rng48 state)."""
        out = _SamDevice()
        st = C.c_uint64(rng48_state)
        _check(lib().hsa_sam_se_device(self._h, codes_ptr, off_ptr, len_ptr, n_reads, max_len, n_aln_ptr, aln_off_ptr, aln_ptr, C.byref(opt),
                                       n_occ, C.byref(st), stream_ptr, C.byref(out)))
        return out, st.value

    def sa_values_device(self, idx_ptr: int, n: int, out_ptr: int, steps_ptr: int = 0, stream_ptr: int = 0) -> None:
        _check(lib().hsa_sa_values_device(self._h, idx_ptr, n, out_ptr, steps_ptr, stream_ptr))

    # ---- rank ---------------------------------------------------------------------------------------
    def occ(self, which: int, indices: np.ndarray, layout: int = 1) -> np.ndarray:
        """BWTAllOccValue (BWT.c:793) for every index; which: 0 fwd / 1 rev; layout 0 reference, 1 device."""
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        o4 = np.zeros((idx.shape[0], 4), dtype=np.uint32)
        o1 = np.zeros((idx.shape[0], 4), dtype=np.uint32)
        _check(lib().hsa_occ_batch(self._h, which, layout, idx.ctypes.data, idx.shape[0], o4.ctypes.data, o1.ctypes.data))
        return o4

    # ---- bwt_cal_width --------------------------------------------------------------------------------
    def cal_width(self, codes: np.ndarray, off: np.ndarray, lens: np.ndarray):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        n = lens.shape[0]
        total = int(lens.sum()) + n
        w = np.zeros((total, 2), dtype=np.uint32)
        bid = np.zeros(n, dtype=np.int32)
        _check(lib().hsa_cal_width_batch(self._h, codes.ctypes.data, off.ctypes.data, lens.ctypes.data, n, 1,
                                         w.ctypes.data, bid.ctypes.data))
        return bid, w

    # ---- bwt_match_gap ----------------------------------------------------------------------------------
    def match_gap_batch(self, codes: np.ndarray, tasks: np.ndarray, opts) -> BatchResult:
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
        optarr = (GapOpt * len(opts))(*opts)
        _check(lib().hsa_match_gap_batch(self._h, codes.ctypes.data, codes.shape[0], tasks.ctypes.data, tasks.shape[0],
                                         C.cast(optarr, C.c_void_p), len(opts), C.byref(self._res)))
        return BatchResult(self._res)

    def whole_reads(self, codes, off, lens, opt: GapOpt, keep_gape: int = 0, copy: bool = True) -> BatchResult:
        """Whole-read part of bwa_cal_sa_reg_gap (bwtaln.c:303-360, 371-372). Accepts numpy arrays or raw
        (pointer, ...) triples of pinned host buffers via objects exposing .ctypes / data_ptr()."""
        cp, op, lp, n = _ptr(codes, np.uint8), _ptr(off, np.uint64), _ptr(lens, np.uint32), _len(lens)
        _check(lib().hsa_whole_reads(self._h, cp[0], op[0], lp[0], n, C.byref(opt), keep_gape, C.byref(self._res)))
        return BatchResult(self._res, copy=copy)

    def whole_reads_submit(self, codes, off, lens, opt: GapOpt, keep_gape: int = 0) -> "Job":
        """Asynchronous whole_reads: enqueue H2D + kernels and return; Job.wait() finishes and fetches.  Keep
        the host buffers alive (pinned memory recommended) until then; at most three jobs in flight."""
        cp, op, lp, n = _ptr(codes, np.uint8), _ptr(off, np.uint64), _ptr(lens, np.uint32), _len(lens)
        h = C.c_void_p()
        _check(lib().hsa_whole_reads_submit(self._h, cp[0], op[0], lp[0], n, C.byref(opt), keep_gape, C.byref(h)))
        res = self._job_res[self._job_seq % len(self._job_res)]
        self._job_seq += 1
        return Job(h, res, (cp, op, lp))

    def splice_seeds_submit(self, codes, off, lens, opt: GapOpt) -> "Job":
        """Asynchronous splice_seeds (see whole_reads_submit)."""
        cp, op, lp, n = _ptr(codes, np.uint8), _ptr(off, np.uint64), _ptr(lens, np.uint32), _len(lens)
        h = C.c_void_p()
        _check(lib().hsa_splice_seeds_submit(self._h, cp[0], op[0], lp[0], n, C.byref(opt), C.byref(h)))
        res = self._job_res[self._job_seq % len(self._job_res)]
        self._job_seq += 1
        return Job(h, res, (cp, op, lp))

    def splice_seeds(self, codes, off, lens, opt: GapOpt) -> BatchResult:
        """The six seed searches of bwt_splice_match (bwtgap.c:797-820) per read."""
        cp, op, lp, n = _ptr(codes, np.uint8), _ptr(off, np.uint64), _ptr(lens, np.uint32), _len(lens)
        _check(lib().hsa_splice_seeds(self._h, cp[0], op[0], lp[0], n, C.byref(opt), C.byref(self._res)))
        return BatchResult(self._res)


class Job:
    """One batch in flight (hsa_job_t)."""

    def __init__(self, handle, res, keepalive):
        self._h, self._res, self._keep = handle, res, keepalive

    def wait(self, copy: bool = True) -> BatchResult:
        """Finish the batch and return its results (copy=False: views of the library's pinned result buffers,
        valid until four more jobs have been submitted on the same Index)."""
        if self._h is None:
            raise HsaError("job already waited for")
        h, self._h = self._h, None
        _check(lib().hsa_job_wait(h, C.byref(self._res)))
        self._keep = None
        return BatchResult(self._res, copy=copy)


def _libc_free(p) -> None:
    f = C.CDLL(None).free
    f.argtypes, f.restype = [C.c_void_p], None
    f(p)


class SamResult:
    """hsa_sam_result_t: rec[n, 22] (hsa_sam1_t words), multi[m, 12], cigar[...] (bwa_cigar_t), md bytes; format() = the SAM
    text of bwa_print_sam1 for the batch."""

    def __init__(self, r: _SamResult, rng48_state: int, reads, opt: GapOpt, copy: bool = True):
        """copy=False: rec / multi / cigar are views of the library's arrays (valid until close()); md stays in the library."""
        self._r, self.rng48_state, self._reads, self._opt = r, rng48_state, reads, opt
        n = r.n_reads
        own = (lambda a: a.copy()) if copy else (lambda a: a)
        self.rec = own(np.ctypeslib.as_array(C.cast(r.rec, C.POINTER(C.c_uint32)), shape=(n, SAM_REC_WORDS))) if n else np.zeros((0, SAM_REC_WORDS), np.uint32)
        self.multi = (own(np.ctypeslib.as_array(C.cast(r.multi, C.POINTER(C.c_uint32)), shape=(r.n_multi, SAM_MULTI_WORDS)))
                      if r.n_multi else np.zeros((0, SAM_MULTI_WORDS), np.uint32))
        self.cigar = own(np.ctypeslib.as_array(C.cast(r.cigar, C.POINTER(C.c_uint32)), shape=(r.n_cigar,))) if r.n_cigar else np.zeros(0, np.uint32)
        self.md = (C.string_at(r.md, r.md_bytes) if r.md_bytes else b"") if copy else None
        self.md_bytes = int(r.md_bytes)
        self.n_refined, self.kernel_ms = int(r.n_refined), float(r.kernel_ms)

    def format(self, chr_names, first: int = 0, count: int | None = None, names=None) -> bytes:
        codes, off, lens = self._reads
        n = self._r.n_reads
        count = n - first if count is None else count
        ca = (C.c_char_p * len(chr_names))(*[c.encode() if isinstance(c, str) else c for c in chr_names])
        na = None
        if names is not None:
            na = (C.c_char_p * n)(*[x.encode() if isinstance(x, str) else x for x in names])
        text, nb = C.c_void_p(), C.c_size_t()
        cp, op, lp = _ptr(codes, np.uint8), _ptr(off, np.uint64), _ptr(lens, np.uint32)
        _check(lib().hsa_sam_format(C.byref(self._r), first, count, cp[0], op[0], lp[0], C.cast(na, C.c_void_p) if na is not None else None,
                                    ca, len(chr_names), C.byref(self._opt), C.byref(text), C.byref(nb)))
        out = C.string_at(text, nb.value) if nb.value else b""
        _libc_free(text)
        return out

    def close(self):
        if self._r is not None:
            lib().hsa_sam_result_free(C.byref(self._r))
            self._r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr(x, dtype):
    """(address, keepalive) of a numpy array (made contiguous / typed) or a torch tensor."""
    if hasattr(x, "data_ptr"):
        return x.data_ptr(), x
    a = np.ascontiguousarray(x, dtype=dtype)
    return a.ctypes.data, a


def _len(x) -> int:
    return int(x.shape[0])


class DeviceWorkspace:
    """Scratch for the device-resident entry point (reads and results stay in HBM)."""

    def __init__(self, index: Index):
        self.index = index
        h = C.c_void_p()
        _check(lib().hsa_workspace_create(index._h, 0, 0, 0, C.byref(h)))
        self._h = h

    def whole_reads_device(self, codes_ptr, off_ptr, len_ptr, n_reads, lens_present, opt, n_aln_ptr, aln_off_ptr,
                           aln_ptr, aln_capacity, stats_ptr, stream_ptr=0, keep_gape=0):
        lp = np.ascontiguousarray(lens_present, dtype=np.uint32)
        _check(lib().hsa_whole_reads_device(self.index._h, self._h, codes_ptr, off_ptr, len_ptr, n_reads,
                                            lp.ctypes.data, lp.shape[0], C.byref(opt), keep_gape, n_aln_ptr,
                                            aln_off_ptr, aln_ptr, aln_capacity, stats_ptr, stream_ptr))

    def check(self) -> list:
        """Wait for the last whole_reads_device call and verify that its results are complete (raises HsaError if any
        search was left unprocessed or the hit arena overflowed).  Returns the statistics block
        {-, hits, lookups, heavy, bad, pops, steps, unprocessed}."""
        st = (C.c_uint64 * 8)()
        _check(lib().hsa_workspace_check(self._h, st))
        return [int(x) for x in st]

    def last_launches(self) -> int:
        return int(lib().hsa_workspace_last_launches(self._h))

    def last_config(self) -> dict:
        """How the per-lane search kernel of the last call was configured (hsa_workspace_last_config)."""
        o = (C.c_uint32 * 4)()
        lib().hsa_workspace_last_config(self._h, o)
        return {"resident_blocks_per_sm_bound": int(o[0]), "score_buckets_per_lane": int(o[1]), "shared_memory_per_block_bytes": int(o[2]),
                "bound_bytes_in_shared_memory": bool(o[3])}

    def launch_timing(self, enable: bool) -> None:
        _check(lib().hsa_workspace_launch_timing(self._h, 1 if enable else 0))

    def launch_times(self):
        """([(launch name, ms), ...], occ lookups of the searches the per-lane search kernel completed) of the last
        call made with launch_timing(True); synchronises the device."""
        names = C.create_string_buffer(4096)
        ms = (C.c_float * 64)()
        n, lk = C.c_size_t(0), C.c_uint64(0)
        _check(lib().hsa_workspace_launch_times(self._h, names, 4096, ms, 64, C.byref(n), C.byref(lk)))
        nm = names.value.decode().split(";") if n.value else []
        return [(nm[i], float(ms[i])) for i in range(n.value)], int(lk.value)

    def close(self):
        if self._h:
            lib().hsa_workspace_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def random_sector_probe(device: int, footprint_bytes: int, iters: int = 64) -> float:
    """Measured random 32-byte-sector load throughput in GB/s over `footprint_bytes`."""
    g = C.c_double()
    _check(lib().hsa_random_sector_probe(device, footprint_bytes, iters, C.byref(g)))
    return g.value


PROBE_VARIANTS = ("chains4_2x128b", "mlp4_256b", "mlp8_256b", "mlp16_256b", "mix8_256b+32b", "lf_walk_256b+4b")


def random_sector_probe_ex(device: int, footprint_bytes: int, iters: int = 64) -> dict:
    """Every probe variant's GB/s (see include/hsa_b200.h); the roofline peak is the best of them."""
    v = (C.c_double * len(PROBE_VARIANTS))()
    _check(lib().hsa_random_sector_probe_ex(device, footprint_bytes, iters, v, len(PROBE_VARIANTS)))
    return {k: float(x) for k, x in zip(PROBE_VARIANTS, v)}

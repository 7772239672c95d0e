"""Shared pieces of the SAM-stage tests: the reference's per-read dump ('HSAM', written by
oracle/ref_harness.c mode `sam`), the conversion of hit dumps into the library's input, the comparison, and the CPU build of
hsa_b200/csrc/hsa_sam.cuh (tests/emu).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C

import numpy as np

import emu_lib as el
import oracle_lib as ol

REC_WORDS = 22      # hsa_sam1_t
MULTI_WORDS = 12    # hsa_multi1_t
REC_FIELDS = ["type", "strand", "n_mm", "n_gapo", "n_gape", "mapQ", "score", "sa", "seq_id", "ori_pos", "occ_pos", "c1", "c2",
              "start", "end", "n_cigar", "cigar_off", "nm", "md_len", "md_off", "n_multi", "multi_off"]
MULTI_FIELDS = ["n_cigar", "cigar_off", "gap", "mm", "strand", "sa", "ori_pos", "occ_pos", "seq_id", "aln_id", "start", "end"]
# dump order of ref_harness.c: put_sam_rec
DUMP_REC = ["type", "strand", "n_mm", "n_gapo", "n_gape", "mapQ", "score", "sa", "seq_id", "ori_pos", "occ_pos", "c1", "c2",
            "start", "end", "n_cigar", "nm", "md_len", "n_multi"]
DUMP_MULTI = ["n_cigar", "gap", "mm", "strand", "sa", "ori_pos", "occ_pos", "seq_id", "aln_id", "start", "end"]


def parse_ref_dump(src):
    """HSAM dump (path or uint32 array) -> list of (rec dict, cigar tuple, md bytes, [(multi dict, cigar tuple), ...]) per read."""
    w = np.fromfile(src, dtype=np.uint32) if isinstance(src, str) else np.ascontiguousarray(src, dtype=np.uint32)
    assert w[0] == 0x4D415348, "bad HSAM magic"
    n, p, out = int(w[1]), 2, []
    raw = w.tobytes()
    for _ in range(n):
        rec = dict(zip(DUMP_REC, w[p:p + 19].tolist()))
        p += 20
        cig, md, multi = (), b"", []
        if rec["type"] != 0:
            cig = tuple(w[p:p + rec["n_cigar"]].tolist()); p += rec["n_cigar"]
            md = raw[4 * p:4 * p + rec["md_len"]]; p += (rec["md_len"] + 3) // 4
            for _j in range(rec["n_multi"]):
                m = dict(zip(DUMP_MULTI, w[p:p + 11].tolist())); p += 12
                mc = tuple(w[p:p + m["n_cigar"]].tolist()); p += m["n_cigar"]
                multi.append((m, mc))
        out.append((rec, cig, md, multi))
    assert p == w.shape[0], "trailing words in the HSAM dump"
    return out


def rows12_to_aln9(rows: np.ndarray) -> np.ndarray:
    """12-word dump rows of ref_harness.c -> hsa_aln1_t words."""
    a = np.zeros((rows.shape[0], 9), dtype=np.uint32)
    a[:, 0] = (rows[:, 0] & 0xFFFF) | ((rows[:, 1] & 0xFF) << 16) | ((rows[:, 2] & 0xFF) << 24)
    a[:, 1:5] = rows[:, 3:7]
    a[:, 5] = (rows[:, 7] & 0x3FFFFFFF) | (rows[:, 8] << 30)
    a[:, 6] = rows[:, 9]; a[:, 7] = rows[:, 10]; a[:, 8] = rows[:, 11]
    return a


def hits_input(n_aln: np.ndarray, rows12: np.ndarray):
    """(n_aln int32, aln_off uint64, aln9) as hsa_sam_se_batch takes them, from a driver dump."""
    n_aln = np.ascontiguousarray(n_aln, dtype=np.int32)
    off = np.zeros(n_aln.shape[0], dtype=np.uint64)
    if n_aln.shape[0]:
        off[1:] = np.cumsum(n_aln[:-1].astype(np.uint64))
    return n_aln, off, np.ascontiguousarray(rows12_to_aln9(rows12))


def unpack_result(rec: np.ndarray, multi: np.ndarray, cigar: np.ndarray, md: bytes):
    """library / emulation arrays -> the same per-read structure parse_ref_dump returns."""
    out = []
    for r in rec.tolist():
        d = dict(zip(REC_FIELDS, r))
        d["score"] = d["score"] & 0xFFFFFFFF
        if d["type"] == 0:
            out.append(({k: 0 for k in DUMP_REC}, (), b"", []))
            continue
        cig = tuple(cigar[d["cigar_off"]:d["cigar_off"] + d["n_cigar"]].tolist())
        m_md = md[d["md_off"]:d["md_off"] + d["md_len"]]
        ml = []
        for j in range(d["n_multi"]):
            m = dict(zip(MULTI_FIELDS, multi[d["multi_off"] + j].tolist()))
            mc = tuple(cigar[m["cigar_off"]:m["cigar_off"] + m["n_cigar"]].tolist())
            ml.append(({k: m[k] & 0xFFFFFFFF for k in DUMP_MULTI}, mc))
        out.append(({k: d[k] & 0xFFFFFFFF for k in DUMP_REC}, cig, bytes(m_md), ml))
    return out


def diff(got, want, limit=5):
    """-> list of human-readable differences (empty = identical)."""
    bad = []
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        if g != w:
            bad.append(f"read {i}:\n  got  {g}\n  want {w}")
            if len(bad) >= limit:
                break
    return bad


def emu_sam(emu: el.Emu, rs, n_aln, aln_off, aln9, opt: ol.GapOpt, n_occ=3, rng_state=0):
    """hsa_sam.cuh on the host for a whole batch -> (per-read structure, new rng state, counts)."""
    ix = emu.index
    L = el.lib()
    L.emu_sam.restype = C.c_int
    L.emu_sam.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                          C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int,
                          C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                          C.c_void_p]
    t = np.ascontiguousarray(ix.blocks.table(), dtype=np.uint32)
    off = np.ascontiguousarray(rs.offsets[:-1], dtype=np.uint64)
    lens = np.ascontiguousarray(rs.lens, dtype=np.uint32)
    n = rs.n
    max_len = int(lens.max()) if n else 0
    md_tab = None
    if opt.fnr > 0:
        md_tab = np.asarray([ol.lib().hsao_cal_maxdiff(l, 0.02, opt.fnr) for l in range(max_len + 1)], dtype=np.int32)
    rec = np.zeros((n, REC_WORDS), dtype=np.uint32)
    multi_cap = 128 * n + 1024
    multi = np.zeros((multi_cap, MULTI_WORDS), dtype=np.uint32)
    cigar = np.zeros(64 * n + 1024, dtype=np.uint32)
    md = np.zeros(512 * n + 1024, dtype=np.uint8)
    counts = np.zeros(5, dtype=np.uint64)
    st = C.c_uint64(rng_state)
    aln9 = np.ascontiguousarray(aln9, dtype=np.uint32)
    rc = L.emu_sam(emu.h, ix.fwd.sa_value.ctypes.data, ix.fwd.sa_interval, t.ctypes.data, t.shape[0], ix.packed_dna.ctypes.data,
                   ix.dna_length, rs.codes.ctypes.data, off.ctypes.data, lens.ctypes.data, n, n_aln.ctypes.data, aln_off.ctypes.data,
                   aln9.ctypes.data, None if md_tab is None else md_tab.ctypes.data, opt.max_diff, n_occ, C.byref(st),
                   rec.ctypes.data, multi.ctypes.data, multi_cap, cigar.ctypes.data, cigar.shape[0], md.ctypes.data, md.shape[0],
                   counts.ctypes.data)
    assert rc == 0 and int(counts[4]) == 0, (rc, counts)
    return unpack_result(rec, multi, cigar, md.tobytes()), st.value, counts


def printable_lines(text: bytes) -> bytes:
    """SAM text without the lines whose CIGAR holds an operation code beyond the reference's "MIDNSHP=X" table (a negative
    intron length, bwtse.c:593: the reference prints whatever byte follows its string literal; hsa_sam_format prints '?')."""
    keep = [l for l in text.split(b"\n") if l and b"?" not in l and all(32 <= c < 127 or c == 9 for c in l)]
    return b"\n".join(keep) + b"\n"

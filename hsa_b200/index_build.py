"""2BWT index construction with torch tensor ops (runs on the GPU when one is present, else on the CPU).

This is plumbing, not the hot path: the north star keeps index construction outside the accelerated
path ("built once, uploaded once").  The arrays produced here are bit-identical to what the
reference's `HSA index` writes (BWTConstruct.c:929-1207 occ tables, :1209-1239 file layout), which
tests/test_index_build.py checks against the reference builder and against committed digests, so a
synthetic genome of any size can be indexed on the GPU box where /root/reference does not exist.

Algorithm (memory-lean enough for a 3.1 Gb text on one 180 GB GPU; the suffix array itself is never held):
  * suffixes are bucketed by their first `b` symbols (b grows with the text so that a bucket holds <= 2^27
    suffixes); buckets are visited in lexicographic order;
  * inside a bucket the suffixes are sorted by a 63-bit key of their next 21 symbols (3 bits each, '$' and
    everything past the end = 0, so '$' sorts first), ties are broken by successive 21-symbol keys read straight
    from the text (two rounds suffice for an i.i.d. genome; repeats just take more rounds);
  * each sorted bucket emits its slice of the BWT, bwt[r] = text[sa[r] - 1]; the rank of suffix 0 is inverseSa0
    and its placeholder is dropped from the packed stream ('$' is not stored, BWT.c:804);
  * occ = running symbol counts sampled every 256 symbols (16-bit, relative) and 65536 (32-bit).
The reverse BWT is the same construction on the reversed text (2BWT-Builder.c:116-213).
"""
from __future__ import annotations

import numpy as np
import torch

from .index_io import (BWTArrays, Index2BWT, OCC_INTERVAL, OCC_INTERVAL_MAJOR, SA_INTERVAL, bwt_resident_words,
                       occ_major_words, occ_minor_words, sa_value_words)

KEY_SYMS = 21                       # symbols per sort key (3 bits each -> 63 bits, non-negative int64)
BUCKET_TARGET = 1 << 27             # suffixes per bucket the sort is sized for


def _key_at(text: torch.Tensor, pos: torch.Tensor, start: int, n: int) -> torch.Tensor:
    """63-bit key of symbols [start, start + KEY_SYMS) of the suffixes at `pos` ('$' / past the end = 0)."""
    key = torch.zeros_like(pos)
    for j in range(KEY_SYMS):
        p = pos + (start + j)
        inside = p < n
        sym = torch.where(inside, text[torch.clamp(p, max=n - 1)].to(torch.int64) + 1, torch.zeros_like(p))
        key = (key << 3) | sym
    return key


def _sort_bucket(text: torch.Tensor, pos: torch.Tensor, skip: int, n: int) -> torch.Tensor:
    """Suffix positions of one bucket (all sharing their first `skip` symbols) in lexicographic order."""
    if pos.numel() <= 1:
        return pos
    key = _key_at(text, pos, skip, n)
    key, idx = torch.sort(key, stable=True)
    pos = pos[idx]
    del idx
    tie = key[1:] == key[:-1]
    del key
    h = skip + KEY_SYMS
    # tied[i]: element i belongs to a run of equal prefixes (length >= 2); grp: run id, increasing along the order
    while bool(tie.any()):
        m = pos.numel()
        tied = torch.zeros(m, dtype=torch.bool, device=pos.device)
        tied[1:] |= tie
        tied[:-1] |= tie
        start_of_run = torch.ones(m, dtype=torch.bool, device=pos.device)
        start_of_run[1:] = ~tie
        grp_all = torch.cumsum(start_of_run.to(torch.int64), 0)
        where = torch.nonzero(tied).squeeze(1)          # ascending; runs are contiguous in it
        sub = pos[where]
        grp = grp_all[where]
        del tied, start_of_run, grp_all
        k2 = _key_at(text, sub, h, n)
        k2s, o1 = torch.sort(k2, stable=True)           # by next key ...
        g2, o2 = torch.sort(grp[o1], stable=True)       # ... then (stably) by run: ordered by (run, next key)
        perm = o1[o2]
        k2s = k2s[o2]
        pos[where] = sub[perm]                          # runs keep their slots; only their inside is reordered
        new_tie_sub = (g2[1:] == g2[:-1]) & (k2s[1:] == k2s[:-1])
        tie = torch.zeros(m - 1, dtype=torch.bool, device=pos.device)
        # a tie between slots where[i] and where[i+1] is a tie between adjacent elements (same run => adjacent)
        tie[where[:-1][new_tie_sub]] = True
        h += KEY_SYMS
    return pos


def bwt_of(text: torch.Tensor, sa_interval: int = 0):
    """(bwt symbols with '$' dropped: uint8[n], inverseSa0, SA samples or None) of text + '$' ('$' smallest).
    sa_interval > 0: also the suffix-array samples SA[0], SA[sa_interval], ... (int64; BWTGenerateSaValue,
    BWTConstruct.c:1310-1371), taken from the sorted buckets as they go by."""
    n = int(text.shape[0])
    dev = text.device
    m = n + 1
    sa = torch.zeros(sa_value_words(n, sa_interval), dtype=torch.int64, device=dev) if sa_interval else None
    # symbols the buckets are keyed on
    b = 0
    while (m >> (2 * b)) > BUCKET_TARGET:
        b += 1
    out = torch.empty(m, dtype=torch.uint8, device=dev)
    inverse_sa0 = -1
    if b == 0:
        buckets = [None]
    else:
        code = torch.zeros(m, dtype=torch.int16, device=dev)
        for j in range(b):                              # base-5 code of the first b symbols, '$'/past the end = 0
            sym = torch.zeros(m, dtype=torch.int16, device=dev)
            if n - j > 0:
                sym[: n - j] = text[j:].to(torch.int16) + 1
            code = code * 5 + sym
            del sym
        buckets = list(range(5 ** b))
        counts = torch.bincount(code.to(torch.int64), minlength=5 ** b).cpu().tolist() if m <= (1 << 28) else None
    filled = 0
    for v in buckets:
        if v is None:
            pos = torch.arange(m, dtype=torch.int64, device=dev)
        else:
            if counts is not None and counts[v] == 0:
                continue
            pos = torch.nonzero(code == v).squeeze(1)
            if pos.numel() == 0:
                continue
        order = _sort_bucket(text, pos, b, n)
        del pos
        k = int(order.numel())
        z = torch.nonzero(order == 0)
        if z.numel():
            inverse_sa0 = filled + int(z.item())
        out[filled:filled + k] = text[torch.clamp(order - 1, min=0)]   # entry at inverse_sa0 is a placeholder
        if sa is not None:
            first = (-filled) % sa_interval                 # first rank of this bucket that is a multiple of the interval
            if first < k:
                sa[(filled + first) // sa_interval: (filled + first) // sa_interval + (k - first + sa_interval - 1) // sa_interval] = \
                    order[first::sa_interval]
        filled += k
        del order
    assert filled == m and inverse_sa0 >= 0
    bwt = torch.cat([out[:inverse_sa0], out[inverse_sa0 + 1:]])       # n symbols, '$' removed (BWT.c:804)
    if sa is not None:
        sa[0] = 0xFFFFFFFF                                            # SA[0] = textLength is kept as -1 (BWT.c:222)
    return bwt, inverse_sa0, sa


def _pack_2bit_msb(symbols: torch.Tensor, n_words: int) -> torch.Tensor:
    """Pack 2-bit symbols 16 per uint32 word, first symbol in the two most significant bits (BWT.c:954).
    Returns int64 values < 2^32."""
    dev = symbols.device
    padded = torch.zeros(n_words * 16, dtype=torch.uint8, device=dev)
    padded[: symbols.shape[0]] = symbols
    q = padded.view(n_words * 4, 4)
    byts = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]          # 4 symbols per byte, first in the MSBs
    del padded, q
    w = byts.view(n_words, 4).to(torch.int64)
    return (w[:, 0] << 24) | (w[:, 1] << 16) | (w[:, 2] << 8) | w[:, 3]


def build_bwt(text: torch.Tensor, sa_interval: int = 0) -> dict:
    """One direction.  Returns torch tensors (int64 holding uint32 values) + scalars."""
    n = int(text.shape[0])
    dev = text.device
    bwt, inverse_sa0, sa = bwt_of(text, sa_interval)
    counts = torch.bincount(bwt, minlength=4)[:4].to(torch.int64) if n < (1 << 30) else \
        torch.stack([(bwt == c).sum() for c in range(4)]).to(torch.int64)
    cum = torch.zeros(5, dtype=torch.int64, device=dev)
    cum[1:] = torch.cumsum(counts, 0)
    code = _pack_2bit_msb(bwt, bwt_resident_words(n))
    # occ samples: counts of each symbol in bwt[0 : e*256), e = 0 .. ceil(n/256)
    num_occ = (n + OCC_INTERVAL - 1) // OCC_INTERVAL + 1
    n_pad = (num_occ - 1) * OCC_INTERVAL
    occ_abs = torch.zeros((num_occ, 4), dtype=torch.int64, device=dev)
    full = n // OCC_INTERVAL                             # complete 256-symbol intervals
    for c in range(4):
        per = torch.zeros(num_occ - 1, dtype=torch.int64, device=dev)
        if full:
            per[:full] = (bwt[: full * OCC_INTERVAL].view(full, OCC_INTERVAL) == c).sum(dim=1)
        if full < num_occ - 1:                           # the last, partial interval
            tail = int((bwt[full * OCC_INTERVAL:] == c).sum().item())
            if c == 0:
                tail += n_pad - n    # the zero padding past textLength counts as 'A' in the last sample, exactly
                                     # as BWTDecodeAll's A = span - C - G - T does (BWT.c:677)
            per[full] = tail
        occ_abs[1:, c] = torch.cumsum(per, 0)
        del per
    per_major = OCC_INTERVAL_MAJOR // OCC_INTERVAL
    e = torch.arange(num_occ, device=dev)
    major_rows = occ_abs[(e // per_major) * per_major]   # absolute count at the enclosing major sample
    minor = occ_abs - major_rows                         # < 65536 (BWTConstruct.c:1097-1104)
    n_minor_words = occ_minor_words(n)
    n_pairs = n_minor_words // 4
    minor_pad = torch.zeros((n_pairs * 2, 4), dtype=torch.int64, device=dev)
    minor_pad[:num_occ] = minor
    if num_occ < n_pairs * 2:
        minor_pad[num_occ:] = minor[-1]      # the unused half of the last word repeats the last sample, as the
                                             # reference builder leaves it (never read: e <= num_occ - 1)
    mp = minor_pad.view(n_pairs, 2, 4)
    occ_value = ((mp[:, 0, :] << 16) | mp[:, 1, :]).reshape(-1)      # even sample -> high half (BWT.c:1045)
    n_major_words = occ_major_words(n)
    occ_major = torch.zeros(n_major_words, dtype=torch.int64, device=dev)
    maj = occ_abs[::per_major].reshape(-1)
    occ_major[: maj.shape[0]] = maj
    return dict(text_length=n, inverse_sa0=inverse_sa0, cum=cum, code=code, occ_value=occ_value, occ_major=occ_major,
                sa=sa, sa_interval=sa_interval)


def _to_arrays(d: dict) -> BWTArrays:
    u32 = lambda t: t.to("cpu").numpy().astype(np.uint32)   # noqa: E731
    arr = BWTArrays(d["text_length"], d["inverse_sa0"], u32(d["cum"]), u32(d["code"]), u32(d["occ_value"]),
                    u32(d["occ_major"]))
    if d.get("sa") is not None:
        arr.sa_value, arr.sa_interval = u32(d["sa"]), int(d["sa_interval"])
    arr.check()
    return arr


def build_index(genome_codes, device: str | torch.device | None = None, sa_interval: int = SA_INTERVAL) -> Index2BWT:
    """Build both BWTs of a genome given as base codes 0..3 (numpy uint8 or torch tensor), and the forward text's
    suffix-array samples every `sa_interval` SA indices (0 = none)."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    text = torch.as_tensor(np.ascontiguousarray(genome_codes) if isinstance(genome_codes, np.ndarray) else genome_codes)
    text = text.to(device=device, dtype=torch.uint8)
    if int(text.max().item()) > 3:
        raise ValueError("genome must contain only A/C/G/T codes 0..3")
    fwd = _to_arrays(build_bwt(text, sa_interval))
    rev = _to_arrays(build_bwt(torch.flip(text, dims=[0])))
    return Index2BWT(fwd, rev)


def pack_dna_words(text: torch.Tensor, chunk: int = 1 << 28) -> np.ndarray:
    """HSP::packedDNA (16 symbols per 32-bit word, first symbol in the two MSBs; index_io.pack_dna) of a text held as a
    torch tensor, packed where the tensor lives (a 3.1 Gb text is packed on the GPU in chunks)."""
    n = int(text.shape[0])
    words = (n + 15) // 16 + 1
    out = np.zeros(words, dtype=np.uint32)
    sh = (30 - 2 * torch.arange(16, device=text.device, dtype=torch.int64))
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        t = text[lo:hi].to(torch.int64)
        pad = (-t.shape[0]) % 16
        if pad:
            t = torch.cat([t, torch.zeros(pad, dtype=torch.int64, device=text.device)])
        w = (t.reshape(-1, 16) << sh[None, :]).sum(dim=1)
        out[lo // 16: lo // 16 + w.shape[0]] = w.to("cpu").numpy().astype(np.uint32)
    return out


def build_full_index(genome_codes, device=None, names=None) -> Index2BWT:
    """Everything bwa_cal_sa_reg_gap reads for a one-record (or multi-record: pass the record lengths as `names` = list of
    (name, length)) ACGT-only genome: both BWTs, the forward SA samples, the block list and the packed text."""
    from .index_io import Blocks, blocks_of_records
    ix = build_index(genome_codes, device=device)
    text = torch.as_tensor(np.ascontiguousarray(genome_codes) if isinstance(genome_codes, np.ndarray) else genome_codes)
    n = int(text.shape[0])
    if names is None:
        ix.blocks = blocks_of_records([n])
        ix.blocks.names = ["synth"]
    else:
        ix.blocks = blocks_of_records([ln for _, ln in names])
        ix.blocks.names = [nm for nm, _ in names]
    dev = device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
    ix.packed_dna, ix.dna_length = pack_dna_words(text.to(dev)), n
    return ix

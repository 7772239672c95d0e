"""2BWT index construction with torch tensor ops (runs on the GPU when one is present, else on the CPU).

This is plumbing, not the hot path: the north star keeps index construction outside the accelerated
path ("built once, uploaded once").  The arrays produced here are bit-identical to what the
reference's `HSA index` writes (BWTConstruct.c:929-1207 occ tables, :1209-1239 file layout), which
tests/test_index_build.py checks against the reference builder and against committed digests, so a
synthetic genome of any size can be indexed on the GPU box where /root/reference does not exist.

Algorithm: suffix array of text$ by prefix doubling on packed integer keys (torch.sort), then
  bwt[r]  = text[sa[r]-1]              ('$' at r = inverseSa0 is dropped from the packed stream)
  occ     = running symbol counts sampled every 256 symbols (16-bit, relative) and 65536 (32-bit)
The reverse BWT is the same construction on the reversed text (2BWT-Builder.c:116-213).
"""
from __future__ import annotations

import numpy as np
import torch

from .index_io import (BWTArrays, Index2BWT, OCC_INTERVAL, OCC_INTERVAL_MAJOR, bwt_resident_words,
                       occ_major_words, occ_minor_words)


def suffix_array(text: torch.Tensor) -> torch.Tensor:
    """Suffix array of text + '$' ('$' smallest): int64[n+1], sa[0] == n."""
    n = int(text.shape[0])
    dev = text.device
    m = n + 1
    K = 20                                              # symbols per initial key, 3 bits each
    sym = torch.zeros(m + K, dtype=torch.int64, device=dev)
    sym[:n] = text.to(torch.int64) + 1                  # '$' and everything past it = 0
    key = torch.zeros(m, dtype=torch.int64, device=dev)
    for j in range(K):
        key = (key << 3) | sym[j:j + m]
    del sym
    order = torch.argsort(key, stable=True)
    skey = key[order]
    del key
    diff = torch.ones(m, dtype=torch.int64, device=dev)
    diff[1:] = (skey[1:] != skey[:-1]).to(torch.int64)
    del skey
    h = K
    while True:
        grp = torch.cumsum(diff, 0) - 1                 # dense rank of each sorted suffix under its first h symbols
        if int(grp[-1].item()) + 1 == m:
            return order
        rank = torch.empty(m, dtype=torch.int64, device=dev)
        rank[order] = grp
        # only suffixes inside tied groups need the second key (a random genome has a handful)
        same_next = diff[1:] == 0
        tied = torch.zeros(m, dtype=torch.bool, device=dev)
        tied[1:] |= same_next
        tied[:-1] |= same_next
        pos = torch.nonzero(tied).squeeze(1)            # positions in sorted order, ascending
        suf = order[pos]
        nxt = suf + h
        r2 = torch.where(nxt < m, rank[torch.clamp(nxt, max=m - 1)], torch.zeros_like(nxt))
        comp = grp[pos] * (m + 1) + r2                  # (group, rank of the next h symbols)
        comp_sorted, sub = torch.sort(comp, stable=True)
        order[pos] = suf[sub]                           # groups are contiguous: positions stay in place
        nd = torch.ones(pos.shape[0], dtype=torch.int64, device=dev)
        nd[1:] = (comp_sorted[1:] != comp_sorted[:-1]).to(torch.int64)
        diff = torch.ones(m, dtype=torch.int64, device=dev)
        diff[pos] = nd
        h *= 2


def _pack_2bit_msb(symbols: torch.Tensor, n_words: int) -> torch.Tensor:
    """Pack 2-bit symbols 16 per uint32 word, first symbol in the two most significant bits (BWT.c:954)."""
    dev = symbols.device
    padded = torch.zeros(n_words * 16, dtype=torch.int64, device=dev)
    padded[: symbols.shape[0]] = symbols.to(torch.int64)
    shifts = (30 - 2 * torch.arange(16, device=dev, dtype=torch.int64))
    words = (padded.view(n_words, 16) << shifts).sum(dim=1)
    return words


def build_bwt(text: torch.Tensor) -> dict:
    """One direction.  Returns torch tensors (int64 holding uint32 values) + scalars."""
    n = int(text.shape[0])
    dev = text.device
    sa = suffix_array(text)
    inverse_sa0 = int(torch.nonzero(sa == 0).item())
    prev = torch.clamp(sa - 1, min=0)
    bwt_full = text[prev]                                # entry at inverse_sa0 is a placeholder
    keep = torch.ones(n + 1, dtype=torch.bool, device=dev)
    keep[inverse_sa0] = False
    bwt = bwt_full[keep]                                 # n symbols, '$' removed (BWT.c:804)
    del sa, prev, bwt_full, keep
    counts = torch.bincount(bwt.to(torch.int64), minlength=4)
    cum = torch.zeros(5, dtype=torch.int64, device=dev)
    cum[1:] = torch.cumsum(counts, 0)
    code = _pack_2bit_msb(bwt, bwt_resident_words(n))
    # occ samples: counts of each symbol in bwt[0 : e*256), e = 0 .. ceil(n/256)
    num_occ = (n + OCC_INTERVAL - 1) // OCC_INTERVAL + 1
    n_pad = (num_occ - 1) * OCC_INTERVAL
    occ_abs = torch.zeros((num_occ, 4), dtype=torch.int64, device=dev)
    for c in range(4):
        ind = torch.zeros(n_pad, dtype=torch.int64, device=dev)
        ind[:n] = (bwt == c).to(torch.int64)
        if c == 0:
            ind[n:] = 1          # the zero padding past textLength counts as 'A' in the last sample, exactly
                                 # as BWTDecodeAll's A = span - C - G - T does (BWT.c:677)
        per = ind.view(num_occ - 1, OCC_INTERVAL).sum(dim=1)
        occ_abs[1:, c] = torch.cumsum(per, 0)
        del ind, per
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
    return dict(text_length=n, inverse_sa0=inverse_sa0, cum=cum, code=code, occ_value=occ_value, occ_major=occ_major)


def _to_arrays(d: dict) -> BWTArrays:
    u32 = lambda t: t.to("cpu").numpy().astype(np.uint32)   # noqa: E731
    arr = BWTArrays(d["text_length"], d["inverse_sa0"], u32(d["cum"]), u32(d["code"]), u32(d["occ_value"]),
                    u32(d["occ_major"]))
    arr.check()
    return arr


def build_index(genome_codes, device: str | torch.device | None = None) -> Index2BWT:
    """Build both BWTs of a genome given as base codes 0..3 (numpy uint8 or torch tensor)."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    text = torch.as_tensor(np.ascontiguousarray(genome_codes) if isinstance(genome_codes, np.ndarray) else genome_codes)
    text = text.to(device=device, dtype=torch.uint8)
    if int(text.max().item()) > 3:
        raise ValueError("genome must contain only A/C/G/T codes 0..3")
    fwd = _to_arrays(build_bwt(text))
    rev = _to_arrays(build_bwt(torch.flip(text, dims=[0])))
    return Index2BWT(fwd, rev)

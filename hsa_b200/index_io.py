"""Host-side 2BWT index arrays: read the reference builder's files, or hold arrays built in-process.

File formats follow the reference loader BWTLoad (BWT.c:107-223):
  <prefix>.index.bwt      : inverseSa0, cumulativeFreq[1..4], then ceil(n/16) words of 2-bit codes
  <prefix>.index.fmv      : inverseSa0, cumulativeFreq[1..4], occValue (minor, 16-bit pairs), occValueMajor
  <prefix>.index.rev.bwt / .rev.fmv : the same for the BWT of the reversed text
  <prefix>.index.sa       : inverseSa0, cumulativeFreq[1..4], saInterval, then (n + saInterval) / saInterval suffix-array
                            samples SA[0], SA[saInterval], ... of the forward text (BWTConstruct.c:1373-1392; the loader
                            overwrites entry 0 with -1, BWT.c:218-222)
Sizes follow BWTResidentSizeInWord / BWTOccValueMinorSizeInWord / BWTOccValueMajorSizeInWord
(BWT.c:1079-1116); the resident bwtCode is padded to a multiple of 256 symbols plus 8 words (BWT.c:176).
"""
from __future__ import annotations

import dataclasses
import numpy as np

OCC_INTERVAL = 256
OCC_INTERVAL_MAJOR = 65536
CHAR_PER_WORD = 16
SA_INTERVAL = 8                     # the reference builder's compiled default (2BWT-Builder.c:97)


def sa_value_words(n: int, interval: int = SA_INTERVAL) -> int:
    """saValueSizeInWord, BWT.c:219."""
    return (n + interval) // interval


def bwt_resident_words(n: int) -> int:
    """BWTResidentSizeInWord(n) + WORD_BETWEEN_OCC/2 (BWT.c:176, :1079-1088)."""
    rounded = (n + OCC_INTERVAL - 1) // OCC_INTERVAL * OCC_INTERVAL
    return (rounded + CHAR_PER_WORD - 1) // CHAR_PER_WORD + 8


def bwt_file_words(n: int) -> int:
    return (n + CHAR_PER_WORD - 1) // CHAR_PER_WORD


def occ_minor_words(n: int) -> int:
    """BWTOccValueMinorSizeInWord, BWT.c:1097-1104."""
    num = (n + OCC_INTERVAL - 1) // OCC_INTERVAL + 1
    return (num + 1) // 2 * 4


def occ_major_words(n: int) -> int:
    """BWTOccValueMajorSizeInWord, BWT.c:1106-1116."""
    num = (n + OCC_INTERVAL - 1) // OCC_INTERVAL + 1
    per = OCC_INTERVAL_MAJOR // OCC_INTERVAL
    return (num + per - 1) // per * 4


@dataclasses.dataclass
class BWTArrays:
    """The fields of the reference's `BWT` struct that the search path reads (BWT.h:61-83)."""
    text_length: int
    inverse_sa0: int
    cumulative_freq: np.ndarray   # uint32[5]
    bwt_code: np.ndarray          # uint32[bwt_resident_words]
    occ_value: np.ndarray         # uint32[occ_minor_words]
    occ_value_major: np.ndarray   # uint32[occ_major_words]
    sa_value: np.ndarray | None = None   # uint32[sa_value_words]: SA samples as LOADED (entry 0 = 0xFFFFFFFF); forward only
    sa_interval: int = 0

    def check(self) -> None:
        n = self.text_length
        assert self.cumulative_freq.dtype == np.uint32 and self.cumulative_freq.shape == (5,)
        assert int(self.cumulative_freq[4]) == n
        assert self.bwt_code.dtype == np.uint32 and self.bwt_code.shape[0] >= bwt_file_words(n)
        assert self.occ_value.shape[0] == occ_minor_words(n)
        assert self.occ_value_major.shape[0] == occ_major_words(n)
        if self.sa_value is not None:
            assert self.sa_interval > 0 and self.sa_value.dtype == np.uint32
            assert self.sa_value.shape[0] == sa_value_words(n, self.sa_interval)


@dataclasses.dataclass
class Blocks:
    """HSP::blockList (HSP.h:41-46, loaded from <prefix>.index.ann by HSPLoad, HSP.c:85-106): the N-free stretches of
    the chromosomes in packed-text coordinates.  chr_id, start, end (inclusive), ori: uint32[n_blocks], ascending."""
    chr_id: np.ndarray
    start: np.ndarray
    end: np.ndarray
    ori: np.ndarray
    names: list = dataclasses.field(default_factory=list)

    @property
    def n(self) -> int:
        return int(self.start.shape[0])

    def table(self) -> np.ndarray:
        return np.ascontiguousarray(np.stack([self.chr_id, self.start, self.end, self.ori], axis=1).astype(np.uint32))


def load_ann(path: str) -> Blocks:
    """<prefix>.index.ann as HSPLoad reads it (HSP.c:85-106)."""
    with open(path) as f:
        tok = f.read().split()
    n_chr = int(tok[1])
    p = 3
    names = []
    for _ in range(n_chr):
        names.append(tok[p + 1]); p += 2
    nb = int(tok[p]); p += 1
    t = np.asarray(tok[p: p + 4 * nb], dtype=np.int64).reshape(nb, 4).astype(np.uint32)
    return Blocks(t[:, 0].copy(), t[:, 1].copy(), t[:, 2].copy(), t[:, 3].copy(), names)


def blocks_of_records(lengths) -> Blocks:
    """The block list the reference builder writes for a FASTA of ACGT-only records (one block per record, ori 0;
    HSP.c:222-310)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    ends = np.cumsum(lengths)
    return Blocks(np.arange(lengths.shape[0], dtype=np.uint32), (ends - lengths).astype(np.uint32),
                  (ends - 1).astype(np.uint32), np.zeros(lengths.shape[0], dtype=np.uint32),
                  [f"r{i}" for i in range(lengths.shape[0])])


@dataclasses.dataclass
class Index2BWT:
    fwd: BWTArrays
    rev: BWTArrays
    blocks: Blocks | None = None
    packed_dna: np.ndarray | None = None   # HSP::packedDNA as DNALoadPacked leaves it: 16 symbols per word, first in the MSBs
    dna_length: int = 0


def load_bwt(bwt_path: str, fmv_path: str) -> BWTArrays:
    with open(bwt_path, "rb") as f:
        hdr = np.fromfile(f, dtype=np.uint32, count=5)
        inverse_sa0 = int(hdr[0])
        cum = np.zeros(5, dtype=np.uint32)
        cum[1:] = hdr[1:]
        n = int(cum[4])
        code = np.zeros(bwt_resident_words(n), dtype=np.uint32)
        body = np.fromfile(f, dtype=np.uint32, count=bwt_file_words(n))
        if body.shape[0] != bwt_file_words(n):
            raise ValueError(f"{bwt_path}: truncated bwt code")
        code[: body.shape[0]] = body
        # BWTClearTrailingBwtCode (BWT.c:1118-1150): symbols past textLength are zeroed
        tail = n % CHAR_PER_WORD
        if tail:
            keep = np.uint32((0xFFFFFFFF << (32 - 2 * tail)) & 0xFFFFFFFF)
            code[n // CHAR_PER_WORD] &= keep
    with open(fmv_path, "rb") as f:
        hdr2 = np.fromfile(f, dtype=np.uint32, count=5)
        if int(hdr2[0]) != inverse_sa0 or not np.array_equal(hdr2[1:], cum[1:]):
            raise ValueError(f"{fmv_path}: header does not match {bwt_path}")
        occ = np.fromfile(f, dtype=np.uint32, count=occ_minor_words(n))
        major = np.fromfile(f, dtype=np.uint32, count=occ_major_words(n))
        if occ.shape[0] != occ_minor_words(n) or major.shape[0] != occ_major_words(n):
            raise ValueError(f"{fmv_path}: truncated occ tables")
    arr = BWTArrays(n, inverse_sa0, cum, code, occ, major)
    arr.check()
    return arr


def load_sa(arr: BWTArrays, sa_path: str) -> None:
    """Attach `<prefix>.index.sa` to the forward BWT's arrays the way BWTLoad does (BWT.c:205-223)."""
    with open(sa_path, "rb") as f:
        hdr = np.fromfile(f, dtype=np.uint32, count=6)
        if int(hdr[0]) != arr.inverse_sa0 or not np.array_equal(hdr[1:5], arr.cumulative_freq[1:]):
            raise ValueError(f"{sa_path}: header does not match the BWT")
        interval = int(hdr[5])
        sa = np.fromfile(f, dtype=np.uint32, count=sa_value_words(arr.text_length, interval))
        if sa.shape[0] != sa_value_words(arr.text_length, interval):
            raise ValueError(f"{sa_path}: truncated SA samples")
    sa[0] = 0xFFFFFFFF                                   # BWT.c:222
    arr.sa_value, arr.sa_interval = sa, interval


def save_sa(arr: BWTArrays, sa_path: str) -> None:
    """BWTSaveSaValue, BWTConstruct.c:1373-1392 (entry 0 is written as textLength)."""
    hdr = np.concatenate([np.asarray([arr.inverse_sa0], dtype=np.uint32), arr.cumulative_freq[1:],
                          np.asarray([arr.sa_interval, arr.text_length], dtype=np.uint32)])
    with open(sa_path, "wb") as f:
        hdr.tofile(f)
        arr.sa_value[1:].tofile(f)


def pack_dna(text: np.ndarray) -> np.ndarray:
    """The in-memory HSP::packedDNA of a text given as codes 0..3: symbol k sits in word k >> 4 at bit (~k & 15) << 1
    (DNALoadPacked with convertToWordPacked, TextConverter.c:677-725), one spare zero word at the end."""
    n = int(text.shape[0])
    pad = np.zeros(((n + 15) // 16 + 1) * 16, dtype=np.uint32)
    pad[:n] = text
    sh = (30 - 2 * np.arange(16)).astype(np.uint32)
    return (pad.reshape(-1, 16) << sh[None, :]).sum(axis=1, dtype=np.uint64).astype(np.uint32)


def load_pac(path: str):
    """(packed words, dnaLength) of `<prefix>.index.pac` as written by the reference builder (HSP.c:311-323): four symbols
    per byte, first in the MSBs, then [a zero byte if length % 4 == 0 and] a byte holding length % 4."""
    raw = np.fromfile(path, dtype=np.uint8)
    n_data = raw.shape[0] - 1
    n = (n_data - 1) * 4 + int(raw[-1])
    words = (n + 15) // 16
    buf = np.zeros((words + 1) * 4, dtype=np.uint8)
    buf[:n_data] = raw[:n_data]
    return buf.view(">u4").astype(np.uint32), n


def save_pac(text: np.ndarray, path: str) -> None:
    """Write a text (codes 0..3) in the reference's .pac format."""
    n = int(text.shape[0])
    pad = np.zeros((n + 3) // 4 * 4, dtype=np.uint8)
    pad[:n] = text
    q = pad.reshape(-1, 4)
    body = (q[:, 0] << 6 | q[:, 1] << 4 | q[:, 2] << 2 | q[:, 3]).astype(np.uint8)
    with open(path, "wb") as f:
        body.tofile(f)
        if n % 4 == 0:
            f.write(b"\0")
        f.write(bytes([n % 4]))


def save_ann(blocks: Blocks, n_chars: int, path: str) -> None:
    """`<prefix>.index.ann` as the reference builder writes it (HSP.c:325-339): total characters, record count and the
    FASTA random seed (0), one line per record name, the block count, one line per block."""
    with open(path, "w") as f:
        f.write("%u\t%d\t%d\n" % (n_chars, len(blocks.names), 0))
        for nm in blocks.names:
            f.write("%d\t%s\n" % (len(nm), nm))
        f.write("%d\n" % blocks.n)
        for row in blocks.table().tolist():
            f.write("%d\t%u\t%u\t%u\n" % tuple(row))


def pack_bytes_of_words(words: np.ndarray, n: int) -> np.ndarray:
    """The .pac body (four symbols per byte, first in the MSBs) from the in-memory packed words."""
    return np.ascontiguousarray(words[: (n + 15) // 16]).astype(">u4").view(np.uint8)[: (n + 3) // 4]


def save_index(ix: Index2BWT, prefix: str) -> None:
    """Write every file BWTLoad2BWT reads (2BWT-Interface.c:13-66) in the reference's own formats, so that an index made
    by the product's builder loads into the unmodified reference (`HSA aln <prefix> ...`): .index.{bwt,fmv,rev.bwt,rev.fmv}
    and, when the index carries them, .index.sa (BWTConstruct.c:1373-1392), .index.pac (HSP.c:311-323), .index.ann."""
    p = prefix + ".index"
    save_bwt(ix.fwd, p + ".bwt", p + ".fmv")
    save_bwt(ix.rev, p + ".rev.bwt", p + ".rev.fmv")
    if ix.fwd.sa_value is not None:
        save_sa(ix.fwd, p + ".sa")
    if ix.packed_dna is not None:
        n = int(ix.dna_length)
        with open(p + ".pac", "wb") as f:
            pack_bytes_of_words(ix.packed_dna, n).tofile(f)
            if n % 4 == 0:
                f.write(b"\0")
            f.write(bytes([n % 4]))
    if ix.blocks is not None:
        save_ann(ix.blocks, int(ix.fwd.text_length), p + ".ann")


def load_index(prefix: str, with_sa: bool = True) -> Index2BWT:
    """Load `<prefix>.index.{bwt,fmv,rev.bwt,rev.fmv}` (and `.sa` if present) as written by `HSA index <prefix> <fasta>`."""
    import os
    p = prefix + ".index"
    ix = Index2BWT(load_bwt(p + ".bwt", p + ".fmv"), load_bwt(p + ".rev.bwt", p + ".rev.fmv"))
    if with_sa and os.path.exists(p + ".sa"):
        load_sa(ix.fwd, p + ".sa")
    if os.path.exists(p + ".ann"):
        ix.blocks = load_ann(p + ".ann")
    if os.path.exists(p + ".pac"):
        ix.packed_dna, ix.dna_length = load_pac(p + ".pac")
    return ix


def save_bwt(arr: BWTArrays, bwt_path: str, fmv_path: str) -> None:
    """Write the two files in the reference's on-disk format (BWTConstruct.c:1209-1239)."""
    n = arr.text_length
    hdr = np.concatenate([np.asarray([arr.inverse_sa0], dtype=np.uint32), arr.cumulative_freq[1:]])
    with open(bwt_path, "wb") as f:
        hdr.tofile(f)
        arr.bwt_code[: bwt_file_words(n)].tofile(f)
    with open(fmv_path, "wb") as f:
        hdr.tofile(f)
        arr.occ_value.tofile(f)
        arr.occ_value_major.tofile(f)

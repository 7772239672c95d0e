"""Seeded synthetic genomes and simulated reads of the shapes BASELINE.json names.

Base codes follow the reference's read encoding (bwaseqio.c: A,C,G,T = 0..3, N = 4).  Genomes are i.i.d.
uniform A/C/G/T with length % 16 != 0 and no N, because the reference index builder corrupts the
reverse text when textLength % 16 == 0 (2BWT-Builder.c:189-208) and mishandles N runs (HSP.c:311-323).
"""
from __future__ import annotations

import dataclasses
import numpy as np

READS_MAGIC = 0x52415348  # 'HSAR'
ALN_MAGIC = 0x41415348    # 'HSAA'
ALN_WORDS = 12            # n_mm n_gapo n_gape k l rev_k rev_l type strand start end score


def make_genome(length: int, seed: int) -> np.ndarray:
    if length % 16 == 0:
        raise ValueError("genome length % 16 must be != 0 (reference builder bug, 2BWT-Builder.c:189-208)")
    rng = np.random.default_rng(seed)
    return rng.integers(0, 4, size=length, dtype=np.uint8)


def write_fasta(path: str, codes: np.ndarray, name: str = "synth", width: int = 80) -> None:
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    n = codes.shape[0]
    full = n // width
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        if full:
            body = np.empty((full, width + 1), dtype=np.uint8)
            body[:, :width] = lut[codes[: full * width]].reshape(full, width)
            body[:, width] = ord("\n")
            f.write(body.tobytes())
        if n % width:
            f.write(lut[codes[full * width:]].tobytes() + b"\n")


def revcomp(codes: np.ndarray) -> np.ndarray:
    """seq_reverse(len, seq, 1) of bwaseqio.c:73-90: reverse, 3-c for c<4, codes >= 4 unchanged."""
    r = codes[..., ::-1].copy()
    m = r < 4
    r[m] = 3 - r[m]
    return r


@dataclasses.dataclass
class ReadSet:
    lens: np.ndarray      # uint32 [n]
    codes: np.ndarray     # uint8 [sum(lens)]

    @property
    def n(self) -> int:
        return int(self.lens.shape[0])

    @property
    def offsets(self) -> np.ndarray:
        off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(self.lens, out=off[1:])
        return off

    def read(self, i: int) -> np.ndarray:
        off = self.offsets
        return self.codes[off[i]:off[i + 1]]

    def subset(self, lo: int, hi: int) -> "ReadSet":
        off = self.offsets
        return ReadSet(self.lens[lo:hi].copy(), self.codes[off[lo]:off[hi]].copy())


def simulate_reads(genome: np.ndarray, n: int, length: int, seed: int, sub_rate: float = 0.01,
                   indel_frac: float = 0.05, max_indels: int = 1, n_rate: float = 0.001,
                   indel_margin: int = 8) -> ReadSet:
    """Fixed-length reads: uniform start, 50 % reverse-complemented, per-base substitutions, a 1-bp
    indel (up to `max_indels`) in `indel_frac` of the reads placed >= `indel_margin` from the ends,
    `n_rate` of bases turned into N."""
    rng = np.random.default_rng(seed)
    G = genome.shape[0]
    pad = max_indels + 1
    starts = rng.integers(0, G - length - pad, size=n)
    idx = starts[:, None] + np.arange(length + pad)[None, :]
    win = genome[idx]                                   # [n, length+pad]
    reads = win[:, :length].copy()
    # indels (loop only over the affected reads)
    sel = np.nonzero(rng.random(n) < indel_frac)[0]
    for r in sel:
        src = win[r].tolist()
        k = int(rng.integers(1, max_indels + 1))
        for _ in range(k):
            pos = int(rng.integers(indel_margin, length - indel_margin))
            if rng.random() < 0.5:
                del src[pos]                            # deletion from the read
            else:
                src.insert(pos, int(rng.integers(0, 4)))  # insertion into the read
        reads[r] = np.asarray(src[:length], dtype=np.uint8)
    # substitutions
    sub = rng.random((n, length)) < sub_rate
    shift = rng.integers(1, 4, size=(n, length), dtype=np.uint8)
    reads = np.where(sub, (reads + shift) & 3, reads).astype(np.uint8)
    # strand
    rc = rng.random(n) < 0.5
    reads[rc] = revcomp(reads[rc])
    # N
    if n_rate > 0:
        reads[rng.random((n, length)) < n_rate] = 4
    return ReadSet(np.full(n, length, dtype=np.uint32), reads.reshape(-1))


def simulate_spliced_reads(genome: np.ndarray, n: int, length: int, seed: int, sub_rate: float = 0.01,
                           min_intron: int = 50, max_intron: int = 50000) -> tuple[ReadSet, np.ndarray]:
    """cfg-4 style reads: two exons joined across a synthetic intron.  Returns the reads and a
    GT..AG-patched copy of the genome positions is NOT made (the genome stays i.i.d.); the seed-search
    hot path does not look at motifs, only the host-side splice logic would."""
    rng = np.random.default_rng(seed)
    G = genome.shape[0]
    reads = np.empty((n, length), dtype=np.uint8)
    split = rng.integers(length // 3, length - length // 3, size=n)
    intron = rng.integers(min_intron, max_intron + 1, size=n)
    starts = rng.integers(0, G - length - max_intron - 1, size=n)
    for r in range(n):
        a = int(split[r]); s = int(starts[r]); g = int(intron[r])
        reads[r, :a] = genome[s:s + a]
        reads[r, a:] = genome[s + a + g:s + g + length]
    sub = rng.random((n, length)) < sub_rate
    shift = rng.integers(1, 4, size=(n, length), dtype=np.uint8)
    reads = np.where(sub, (reads + shift) & 3, reads).astype(np.uint8)
    rc = rng.random(n) < 0.5
    reads[rc] = revcomp(reads[rc])
    return ReadSet(np.full(n, length, dtype=np.uint32), reads.reshape(-1)), split


def ragged_reads(genome: np.ndarray, n: int, min_len: int, max_len: int, seed: int,
                 sub_rate: float = 0.02) -> ReadSet:
    """Ragged-length reads for edge-case tests."""
    rng = np.random.default_rng(seed)
    G = genome.shape[0]
    lens = rng.integers(min_len, max_len + 1, size=n).astype(np.uint32)
    parts = []
    for L in lens.tolist():
        s = int(rng.integers(0, G - L))
        r = genome[s:s + L].copy()
        m = rng.random(L) < sub_rate
        r[m] = (r[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
        if rng.random() < 0.5:
            r = revcomp(r)
        parts.append(r)
    return ReadSet(lens, np.concatenate(parts) if parts else np.zeros(0, np.uint8))


def write_reads_bin(path: str, rs: ReadSet) -> None:
    with open(path, "wb") as f:
        np.asarray([READS_MAGIC, rs.n], dtype=np.uint32).tofile(f)
        rs.lens.astype(np.uint32).tofile(f)
        rs.codes.astype(np.uint8).tofile(f)


def read_reads_bin(path: str) -> ReadSet:
    with open(path, "rb") as f:
        hdr = np.fromfile(f, dtype=np.uint32, count=2)
        assert int(hdr[0]) == READS_MAGIC
        lens = np.fromfile(f, dtype=np.uint32, count=int(hdr[1]))
        codes = np.fromfile(f, dtype=np.uint8, count=int(lens.sum()))
    return ReadSet(lens, codes)


def read_aln_dump(path: str) -> tuple[np.ndarray, np.ndarray]:
    """Parse an 'HSAA' dump -> (n_aln[int32 n_items], alns[uint32 total, 12])."""
    raw = np.fromfile(path, dtype=np.uint32)
    assert int(raw[0]) == ALN_MAGIC
    n_items = int(raw[1])
    n_aln = np.zeros(n_items, dtype=np.int32)
    rows = []
    p = 2
    for i in range(n_items):
        c = int(raw[p]); p += 1
        n_aln[i] = c
        if c:
            rows.append(raw[p:p + c * ALN_WORDS].reshape(c, ALN_WORDS))
            p += c * ALN_WORDS
    assert p == raw.shape[0]
    alns = np.concatenate(rows) if rows else np.zeros((0, ALN_WORDS), dtype=np.uint32)
    return n_aln, alns


def make_repeat_genome(length: int, seed: int, n_dups: int = 40, dup_len: int = 300, tandem: int = 12) -> np.ndarray:
    """A random genome with dispersed near-identical duplications and tandem repeats, so that SA intervals
    with many occurrences exist (exercises best_cnt / max_top2 / the (k,l) dedupe of bwtgap.c:201-213),
    which an i.i.d. genome almost never does."""
    g = make_genome(length, seed)
    rng = np.random.default_rng(seed + 1000)
    for _ in range(n_dups):
        src = int(rng.integers(0, length - dup_len))
        seg = g[src:src + dup_len].copy()
        for _copy in range(int(rng.integers(1, 5))):
            dst = int(rng.integers(0, length - dup_len))
            s2 = seg.copy()
            m = rng.random(dup_len) < 0.01
            s2[m] = (s2[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
            g[dst:dst + dup_len] = s2
    for _ in range(tandem):
        unit = rng.integers(0, 4, size=int(rng.integers(1, 7)), dtype=np.uint8)
        reps = int(rng.integers(20, 80))
        t = np.tile(unit, reps)
        dst = int(rng.integers(0, length - t.shape[0]))
        g[dst:dst + t.shape[0]] = t
    return g


def make_intron_genome(length: int, seed: int, n_introns: int, min_intron: int = 60, max_intron: int = 20000,
                       base: np.ndarray | None = None):
    """A genome (i.i.d., or `base`) with `n_introns` planted introns whose ends carry the motifs the reference's splice
    path looks for (bwtgap.c:536-537: GT..AG, GC..AG, AT..AC on the forward strand): returns (genome, introns[n, 2] =
    first and one-past-last intron base).  Introns do not overlap and keep 200 bases of exon around them."""
    rng = np.random.default_rng(seed)
    g = make_genome(length, seed) if base is None else base.copy()
    introns = []
    pos = 300
    slot = (length - 600) // n_introns
    motifs = [((2, 3), (0, 2)), ((2, 1), (0, 2)), ((0, 3), (0, 1))]          # GT-AG, GC-AG, AT-AC
    for i in range(n_introns):
        ilen = int(rng.integers(min_intron, min(max_intron, slot - 420) + 1))
        a = pos + 200 + int(rng.integers(0, slot - ilen - 400))
        b = a + ilen
        m = motifs[int(rng.choice(3, p=[0.8, 0.1, 0.1]))]
        g[a], g[a + 1] = m[0]
        g[b - 2], g[b - 1] = m[1]
        introns.append((a, b))
        pos += slot
    return g, np.asarray(introns, dtype=np.int64)


def simulate_junction_reads(genome: np.ndarray, introns: np.ndarray, n: int, length: int, seed: int,
                            sub_rate: float = 0.01, min_anchor: int = 8) -> ReadSet:
    """RNA-seq-style reads across the planted introns of make_intron_genome: `left` exon bases ending at the intron
    start joined to length - left bases from the intron end, left uniform in [min_anchor, length - min_anchor];
    substitutions, 50 % reverse-complemented."""
    rng = np.random.default_rng(seed)
    which = rng.integers(0, introns.shape[0], size=n)
    left = rng.integers(min_anchor, length - min_anchor + 1, size=n)
    reads = np.empty((n, length), dtype=np.uint8)
    for r in range(n):
        a, b = int(introns[which[r], 0]), int(introns[which[r], 1])
        le = int(left[r])
        reads[r, :le] = genome[a - le:a]
        reads[r, le:] = genome[b:b + length - le]
    sub = rng.random((n, length)) < sub_rate
    shift = rng.integers(1, 4, size=(n, length), dtype=np.uint8)
    reads = np.where(sub, (reads + shift) & 3, reads).astype(np.uint8)
    rc = rng.random(n) < 0.5
    reads[rc] = revcomp(reads[rc])
    return ReadSet(np.full(n, length, dtype=np.uint32), reads.reshape(-1))

"""Device-side (torch) generators for the large synthetic workloads of BASELINE.json: a 46 Mb / 3.1 Gb
i.i.d. genome and tens of millions of simulated reads are produced directly in HBM.  Plumbing only."""
from __future__ import annotations

import torch


def make_genome(length: int, seed: int, device) -> torch.Tensor:
    if length % 16 == 0:
        raise ValueError("genome length % 16 must be != 0 (reference builder bug, 2BWT-Builder.c:189-208)")
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randint(0, 4, (length,), dtype=torch.uint8, device=device, generator=g)


def simulate_reads(genome: torch.Tensor, n: int, length: int, seed: int, sub_rate: float = 0.01,
                   indel_frac: float = 0.05, n_rate: float = 0.001, indel_margin: int = 8,
                   chunk: int = 2_000_000) -> torch.Tensor:
    """uint8 [n, length] reads (codes 0..3, N = 4): uniform start, 50 % reverse-complemented, per-base
    substitutions, one 1-bp indel in `indel_frac` of the reads >= indel_margin from the ends, n_rate N."""
    dev = genome.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    G = genome.shape[0]
    out = torch.empty((n, length), dtype=torch.uint8, device=dev)
    j = torch.arange(length, device=dev)[None, :]
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        start = torch.randint(0, G - length - 2, (m, 1), device=dev, generator=gen)
        kind = torch.rand((m, 1), device=dev, generator=gen)
        has_indel = kind < indel_frac
        is_del = has_indel & (kind < indel_frac * 0.5)
        is_ins = has_indel & ~is_del
        pos = torch.randint(indel_margin, length - indel_margin, (m, 1), device=dev, generator=gen)
        shift = torch.where(is_del & (j >= pos), 1, 0) - torch.where(is_ins & (j > pos), 1, 0)
        r = genome[start + j + shift]
        ins_base = torch.randint(0, 4, (m, 1), dtype=torch.uint8, device=dev, generator=gen)
        r = torch.where(is_ins & (j == pos), ins_base, r)
        sub = torch.rand((m, length), device=dev, generator=gen) < sub_rate
        sh = torch.randint(1, 4, (m, length), dtype=torch.uint8, device=dev, generator=gen)
        r = torch.where(sub, (r + sh) & 3, r)
        rc = torch.rand((m, 1), device=dev, generator=gen) < 0.5
        r = torch.where(rc, 3 - torch.flip(r, dims=[1]), r)
        if n_rate > 0:
            isn = torch.rand((m, length), device=dev, generator=gen) < n_rate
            r = torch.where(isn, torch.full_like(r, 4), r)
        out[lo:lo + m] = r
    return out


def simulate_spliced_reads(genome: torch.Tensor, n: int, length: int, seed: int, sub_rate: float = 0.01,
                           min_intron: int = 50, max_intron: int = 50_000) -> torch.Tensor:
    """uint8 [n, length] RNA-seq-style reads (BASELINE.json configs[3]): two exons joined across an intron of
    min_intron..max_intron bases, the junction in the middle third of the read, substitutions, 50 % reverse-complemented.
    bwt_splice_match cuts such a read into three seed segments of length // 3 bases (bwtgap.c:763-764, 800-802)."""
    dev = genome.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    G = genome.shape[0]
    start = torch.randint(0, G - max_intron - 2 * length, (n, 1), device=dev, generator=gen)
    split = torch.randint(length // 3, length - length // 3, (n, 1), device=dev, generator=gen)
    intron = torch.randint(min_intron, max_intron, (n, 1), device=dev, generator=gen)
    j = torch.arange(length, device=dev)[None, :]
    reads = genome[start + j + torch.where(j >= split, intron, torch.zeros_like(intron))]
    sub = torch.rand((n, length), device=dev, generator=gen) < sub_rate
    reads = torch.where(sub, (reads + torch.randint(1, 4, (n, length), dtype=torch.uint8, device=dev, generator=gen)) & 3, reads)
    rc = torch.rand((n, 1), device=dev, generator=gen) < 0.5
    return torch.where(rc, 3 - torch.flip(reads, dims=[1]), reads)


def plant_introns(genome: torch.Tensor, n_introns: int, seed: int, min_intron: int = 60, max_intron: int = 20000) -> torch.Tensor:
    """Write the splice motifs the reference looks for (bwtgap.c:536-537) at the ends of `n_introns` non-overlapping
    intervals of `genome` (in place): 80 % GT..AG, 10 % GC..AG, 10 % AT..AC.  Returns introns[n, 2] = first and
    one-past-last intron base (int64, on the genome's device)."""
    dev = genome.device
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    G = int(genome.shape[0])
    slot = (G - 600) // n_introns
    assert slot > max_intron + 900, "too many introns for this genome"
    ilen = torch.randint(min_intron, max_intron + 1, (n_introns,), generator=gen)
    a = 300 + torch.arange(n_introns) * slot + 200 + (torch.rand(n_introns, generator=gen) * (slot - ilen - 400).float()).long()
    b = a + ilen
    kind = torch.rand(n_introns, generator=gen)
    d0 = torch.where(kind < 0.9, torch.tensor(2), torch.tensor(0))                    # G G A
    d1 = torch.where(kind < 0.8, torch.tensor(3), torch.where(kind < 0.9, torch.tensor(1), torch.tensor(3)))   # T C T
    a1 = torch.where(kind < 0.9, torch.tensor(2), torch.tensor(1))                    # AG AG AC
    g = genome
    g[a.to(dev)] = d0.to(dev, torch.uint8)
    g[(a + 1).to(dev)] = d1.to(dev, torch.uint8)
    g[(b - 2).to(dev)] = torch.zeros(n_introns, dtype=torch.uint8, device=dev)
    g[(b - 1).to(dev)] = a1.to(dev, torch.uint8)
    return torch.stack([a, b], dim=1).to(dev)


def simulate_junction_reads(genome: torch.Tensor, introns: torch.Tensor, n: int, length: int, seed: int,
                            sub_rate: float = 0.01, min_anchor: int = 8) -> torch.Tensor:
    """uint8 [n, length] reads across the planted introns: `left` exon bases ending at the intron start joined to
    length - left bases from the intron end; substitutions; 50 % reverse-complemented."""
    dev = genome.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    which = torch.randint(0, introns.shape[0], (n,), device=dev, generator=gen)
    left = torch.randint(min_anchor, length - min_anchor + 1, (n, 1), device=dev, generator=gen)
    a, b = introns[which, 0][:, None], introns[which, 1][:, None]
    j = torch.arange(length, device=dev)[None, :]
    pos = torch.where(j < left, a - left + j, b + (j - left))
    reads = genome[pos]
    sub = torch.rand((n, length), device=dev, generator=gen) < sub_rate
    reads = torch.where(sub, (reads + torch.randint(1, 4, (n, length), dtype=torch.uint8, device=dev, generator=gen)) & 3, reads)
    rc = torch.rand((n, 1), device=dev, generator=gen) < 0.5
    return torch.where(rc, 3 - torch.flip(reads, dims=[1]), reads)

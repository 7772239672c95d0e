"""Generates tests/golden/golden_splice.{json,npz}: outputs of the UNMODIFIED reference's bwt_splice_match (bwtgap.c:748)
on seeded inputs, through oracle/_ref/hsa_ref `splice` (oracle/ref_harness.c: the frame bwa_cal_sa_reg_gap builds for a
read that found nothing on either strand).  Run in the build container (needs /root/reference compiled by
`make -C oracle ref`):   python tests/golden/make_golden_splice.py
The fixtures are what the CPU suite checks the host build of hsa_b200/csrc/hsa_splice.cuh against and what the -m gpu
suite checks the CUDA path against on a box without the reference sources."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from hsa_b200 import index_io, synth  # noqa: E402

GENOME = dict(length=300007, seed=77, n_introns=110, min_intron=60, max_intron=2200,
              base=dict(length=300007, seed=78, n_dups=30, dup_len=300, tandem=10))

# name -> (read spec, gap_opt_t overrides, clear_gape)
CASES = {
    "junction_100": (dict(kind="junction", n=700, length=100, seed=1, sub_rate=0.01), {}, 1),
    "junction_100_gape_kept": (dict(kind="junction", n=500, length=100, seed=2, sub_rate=0.02), {}, 0),
    "junction_75": (dict(kind="junction", n=500, length=75, seed=3, sub_rate=0.01), {}, 1),
    "junction_150_n3o2": (dict(kind="junction", n=400, length=150, seed=4, sub_rate=0.02), dict(fnr=0.0, max_diff=3, max_gapo=2), 1),
    "junction_50_loggap": (dict(kind="junction", n=400, length=50, seed=5, sub_rate=0.01), dict(mode=0x07), 0),
    "random_introns": (dict(kind="spliced", n=500, length=100, seed=6), {}, 1),
    "unalignable_mix": (dict(kind="mix", n=500, length=100, seed=7), {}, 1),
    "ragged": (dict(kind="ragged", n=450, seed=8), {}, 1),
}


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_genome():
    base = synth.make_repeat_genome(**GENOME["base"])
    kw = {k: v for k, v in GENOME.items() if k != "base"}
    return synth.make_intron_genome(base=base, **kw)


def make_reads(genome, introns, spec) -> synth.ReadSet:
    k = spec["kind"]
    if k == "junction":
        return synth.simulate_junction_reads(genome, introns, spec["n"], spec["length"], spec["seed"], sub_rate=spec["sub_rate"])
    if k == "spliced":
        return synth.simulate_spliced_reads(genome, spec["n"], spec["length"], spec["seed"], min_intron=60, max_intron=20000)[0]
    if k == "mix":          # what really reaches the splice path: diverged reads, reads with N, junk
        n, L, rng = spec["n"], spec["length"], np.random.default_rng(spec["seed"])
        a = synth.simulate_reads(genome, n // 2, L, spec["seed"] + 100, sub_rate=0.08, indel_frac=0.3)
        j = synth.simulate_junction_reads(genome, introns, n // 4, L, spec["seed"] + 200, sub_rate=0.03)
        c = j.codes.copy()
        c[rng.random(c.shape[0]) < 0.004] = 4
        junk = rng.integers(0, 4, size=(n - n // 2 - n // 4) * L, dtype=np.uint8)
        return synth.ReadSet(np.full(n, L, dtype=np.uint32), np.concatenate([a.codes, c, junk]))
    if k == "ragged":
        parts = [synth.simulate_junction_reads(genome, introns, spec["n"] // 3, L, spec["seed"] + L) for L in (90, 101, 76)]
        return synth.ReadSet(np.concatenate([p.lens for p in parts]), np.concatenate([p.codes for p in parts]))
    raise ValueError(k)


def main():
    assert ol.have_ref(), "build oracle/_ref first: make -C oracle ref"
    genome, introns = make_genome()
    meta, arrays = dict(genome=GENOME, genome_digest=digest(genome), cases={}), {}
    with tempfile.TemporaryDirectory() as td:
        synth.write_fasta(os.path.join(td, "g.fa"), genome)
        subprocess.run([ol.REF_BIN, "index", "g", "g.fa"], cwd=td, check=True, stdout=subprocess.DEVNULL)
        prefix = os.path.join(td, "g")
        ix = index_io.load_index(prefix)
        meta["packed_dna_digest"] = digest(ix.packed_dna[: (ix.dna_length + 15) // 16])
        meta["dna_length"] = ix.dna_length
        meta["blocks"] = ix.blocks.table().tolist()
        for name, (spec, okw, clear_gape) in CASES.items():
            rs = make_reads(genome, introns, spec)
            rp = os.path.join(td, name + ".reads")
            synth.write_reads_bin(rp, rs)
            opt = ol.default_opt(**okw)
            outp = os.path.join(td, name + ".aln")
            ol.run_ref(["splice", prefix, rp, outp] + ol.opt_args(opt) + [f"clear_gape={clear_gape}"])
            n_aln, rows = synth.read_aln_dump(outp)
            arrays[name + ".n_aln"] = n_aln.astype(np.int32)
            arrays[name + ".rows"] = rows
            meta["cases"][name] = dict(reads=spec, opt=okw, clear_gape=clear_gape, reads_digest=digest(rs.codes),
                                       outcomes=np.bincount(n_aln, minlength=3).tolist())
    np.savez_compressed(os.path.join(HERE, "golden_splice.npz"), **arrays)
    with open(os.path.join(HERE, "golden_splice.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote golden_splice.npz", os.path.getsize(os.path.join(HERE, "golden_splice.npz")), "bytes")
    for k, v in meta["cases"].items():
        print(k, v["outcomes"])


if __name__ == "__main__":
    main()

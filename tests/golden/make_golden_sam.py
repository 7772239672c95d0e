"""Generates tests/golden/golden_sam.{json,npz}: what the UNMODIFIED reference's generate_sam_se_core (bwtse.c:884) computes
for seeded read sets -- through oracle/_ref/hsa_ref `sam` (oracle/ref_harness.c: the stock batch loop of bwa_aln_core, i.e.
bwa_cal_sa_reg_gap then generate_sam_se_core, in one process so that the drand48 stream runs as in the reference program)
-- together with the hits bwa_cal_sa_reg_gap produced for the same reads (`driver` mode), which are the stage's input.
Run in the build container (needs /root/reference compiled by `make -C oracle ref`):   python tests/golden/make_golden_sam.py
The genome is the splice goldens' (duplications, tandem repeats and motif-carrying introns: multi-hit reads, gapped reads and
spliced reads all occur)."""
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
sys.path.insert(0, HERE)
import oracle_lib as ol  # noqa: E402
import make_golden_splice as mgs  # noqa: E402
import sam_common as sc  # noqa: E402
from hsa_b200 import synth  # noqa: E402

# name -> (read spec, gap_opt_t overrides, n_occ)
CASES = {
    "dna_100": (dict(kind="dna", n=2500, length=100, seed=31, sub_rate=0.01, indel_frac=0.25), {}, 3),
    "dna_and_junctions": (dict(kind="both", n=1600, length=100, seed=32), {}, 3),
    "dna_150_n5o2": (dict(kind="dna", n=900, length=150, seed=33, sub_rate=0.02, indel_frac=0.5), dict(fnr=0.0, max_diff=5, max_gapo=2), 3),
    "ragged_nocc6": (dict(kind="ragged", n=1200, seed=34), {}, 6),
    "repeats_75": (dict(kind="repeats", n=1200, length=75, seed=35), {}, 3),
}


def make_reads(genome, introns, spec) -> synth.ReadSet:
    k = spec["kind"]
    if k == "dna":
        return synth.simulate_reads(genome, spec["n"], spec["length"], spec["seed"], sub_rate=spec["sub_rate"], indel_frac=spec["indel_frac"])
    if k == "both":
        a = synth.simulate_reads(genome, spec["n"] * 5 // 8, spec["length"], spec["seed"], sub_rate=0.01, indel_frac=0.2)
        j = synth.simulate_junction_reads(genome, introns, spec["n"] - spec["n"] * 5 // 8, spec["length"], spec["seed"] + 1, sub_rate=0.01)
        rng = np.random.default_rng(spec["seed"])
        order = rng.permutation(spec["n"])                       # interleave: the drand48 stream crosses both kinds
        codes = np.concatenate([a.codes, j.codes]).reshape(spec["n"], spec["length"])[order]
        return synth.ReadSet(np.full(spec["n"], spec["length"], dtype=np.uint32), np.ascontiguousarray(codes).reshape(-1))
    if k == "ragged":
        parts = [synth.simulate_reads(genome, spec["n"] // 3, L, spec["seed"] + L, sub_rate=0.01, indel_frac=0.3) for L in (90, 101, 76)]
        return synth.ReadSet(np.concatenate([p.lens for p in parts]), np.concatenate([p.codes for p in parts]))
    if k == "repeats":
        # reads drawn from the duplicated / tandem stretches: several best hits (X0 > 1, the random choice of bwtse.c:44-53)
        rng = np.random.default_rng(spec["seed"])
        n, L = spec["n"], spec["length"]
        base = synth.simulate_reads(genome, n, L, spec["seed"], sub_rate=0.005, indel_frac=0.1)
        return base
    raise ValueError(k)


def main():
    assert ol.have_ref(), "build oracle/_ref first: make -C oracle ref"
    genome, introns = mgs.make_genome()
    meta, arrays = dict(genome=mgs.GENOME, genome_digest=mgs.digest(genome), cases={}), {}
    with tempfile.TemporaryDirectory() as td:
        synth.write_fasta(os.path.join(td, "g.fa"), genome)
        subprocess.run([ol.REF_BIN, "index", "g", "g.fa"], cwd=td, check=True, stdout=subprocess.DEVNULL)
        prefix = os.path.join(td, "g")
        for name, (spec, okw, n_occ) in CASES.items():
            rs = make_reads(genome, introns, spec)
            rp = os.path.join(td, name + ".reads")
            synth.write_reads_bin(rp, rs)
            opt = ol.default_opt(**okw)
            j = ol.run_ref(["sam", prefix, rp, rp + ".bin", rp + ".sam"] + ol.opt_args(opt) + [f"nocc={n_occ}"])
            ol.run_ref(["driver", prefix, rp, rp + ".aln"] + ol.opt_args(opt))
            n_aln, rows = synth.read_aln_dump(rp + ".aln")
            arrays[name + ".n_aln"] = n_aln.astype(np.int32)
            arrays[name + ".rows"] = rows
            arrays[name + ".dump"] = np.fromfile(rp + ".bin", dtype=np.uint32)
            text = open(rp + ".sam", "rb").read()
            recs = sc.parse_ref_dump(rp + ".bin")
            meta["cases"][name] = dict(reads=spec, opt=okw, n_occ=n_occ, reads_digest=mgs.digest(rs.codes),
                                       text_sha256=hashlib.sha256(sc.printable_lines(text)).hexdigest(),
                                       lines=text.count(b"\n"), matched=j["matched"], with_cigar=j["with_cigar"], splicing=j["splicing"],
                                       with_alternatives=sum(1 for r in recs if r[3]),
                                       repeats=sum(1 for r in recs if r[0]["type"] == 2))
    np.savez_compressed(os.path.join(HERE, "golden_sam.npz"), **arrays)
    with open(os.path.join(HERE, "golden_sam.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote golden_sam.npz", os.path.getsize(os.path.join(HERE, "golden_sam.npz")), "bytes")
    for k, v in meta["cases"].items():
        print(k, {x: v[x] for x in ("lines", "matched", "with_cigar", "splicing", "with_alternatives", "repeats")})


if __name__ == "__main__":
    main()

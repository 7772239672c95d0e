"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference (oracle/_ref).

Run in the build container, where /root/reference exists and `make -C oracle ref` has been done:
    python tests/golden/make_golden.py
The committed fixtures pin the oracle (and through it the CUDA path) to the reference's own outputs on
seeded inputs; the tests regenerate the inputs from the seeds recorded in golden.json.
"""
from __future__ import annotations

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

from hsa_b200 import index_io, synth  # noqa: E402
import oracle_lib as ol  # noqa: E402

CASES = {
    # name: (reads spec, option overrides)
    "cfg1_75bp_n2o1": (dict(kind="sim", n=1500, length=75, seed=101), dict(max_diff=2, max_gapo=1, fnr=0.0)),
    "cfg2_100bp_default": (dict(kind="sim", n=1500, length=100, seed=102), dict()),
    "cfg5_150bp_n5o2": (dict(kind="sim", n=400, length=150, seed=103, sub_rate=0.02, indel_frac=0.1, max_indels=2),
                        dict(max_diff=5, max_gapo=2, fnr=0.0)),
    "ragged_nonstop": (dict(kind="ragged", n=600, min_len=33, max_len=140, seed=104), dict(mode=0x12)),
    "ragged_loggap_gape": (dict(kind="ragged", n=600, min_len=33, max_len=140, seed=105), dict(mode=0x07)),
    "short_entries": (dict(kind="sim", n=600, length=60, seed=106), dict(max_entries=60)),
    "exact_only": (dict(kind="sim", n=600, length=50, seed=107, sub_rate=0.0, indel_frac=0.0), dict(max_diff=0, fnr=0.0)),
    "noskip_gaps": (dict(kind="sim", n=500, length=70, seed=108, indel_frac=0.3),
                    dict(max_diff=3, fnr=0.0, max_gapo=2, max_gape=3, indel_end_skip=0)),
}
GENOME = dict(length=120011, seed=77)
GENOME2 = dict(lengths=[3001, 777, 12007, 150, 9001, 75, 20011], seed=78)     # several FASTA records -> several blocks


def make_genome2():
    """The records of the multi-record golden genome (ACGT only) and their concatenation = the indexed text."""
    rng = np.random.default_rng(GENOME2["seed"])
    recs = [rng.integers(0, 4, size=n, dtype=np.uint8) for n in GENOME2["lengths"]]
    return recs, np.concatenate(recs)


def make_reads(genome, spec):
    spec = dict(spec)
    kind = spec.pop("kind")
    if kind == "sim":
        return synth.simulate_reads(genome, **spec)
    rs = synth.ragged_reads(genome, spec["n"], spec["min_len"], spec["max_len"], spec["seed"], sub_rate=0.03)
    rng = np.random.default_rng(spec["seed"] + 1)
    codes = rs.codes.copy()
    codes[rng.random(codes.shape[0]) < 0.01] = 4
    return synth.ReadSet(rs.lens, codes)


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def index_digest(ix) -> dict:
    out = {}
    for name in ("fwd", "rev"):
        b = getattr(ix, name)
        nfile = index_io.bwt_file_words(b.text_length)
        out[name] = dict(text_length=b.text_length, inverse_sa0=b.inverse_sa0,
                         cumulative_freq=[int(x) for x in b.cumulative_freq],
                         bwt_code=digest(b.bwt_code[:nfile]), occ_value=digest(b.occ_value),
                         occ_value_major=digest(b.occ_value_major))
        if b.sa_value is not None:
            out[name].update(sa_interval=int(b.sa_interval), sa_value=digest(b.sa_value))
    return out


def main():
    assert ol.have_ref(), "build oracle/_ref first: make -C oracle ref"
    meta = dict(genome=GENOME, cases={}, maxdiff={})
    genome = synth.make_repeat_genome(**GENOME)
    with tempfile.TemporaryDirectory() as td:
        synth.write_fasta(os.path.join(td, "g.fa"), genome)
        subprocess.run([ol.REF_BIN, "index", "g", "g.fa"], cwd=td, check=True, stdout=subprocess.DEVNULL)
        prefix = os.path.join(td, "g")
        ix = index_io.load_index(prefix)
        meta["index"] = index_digest(ix)
        # rank
        rng = np.random.default_rng(5)
        idx = rng.integers(0, ix.fwd.text_length + 2, size=4000).astype(np.uint32)
        idx[:6] = [0, 1, ix.fwd.text_length, ix.fwd.text_length + 1, ix.fwd.inverse_sa0, ix.fwd.inverse_sa0 + 1]
        with open(os.path.join(td, "idx.bin"), "wb") as f:
            np.asarray([idx.shape[0]], dtype=np.uint32).tofile(f)
            idx.tofile(f)
        ol.run_ref(["occ", prefix, os.path.join(td, "idx.bin"), os.path.join(td, "occ.out")])
        occ = np.fromfile(os.path.join(td, "occ.out"), dtype=np.uint32)[1:].reshape(-1, 16)
        arrays = dict(occ_idx=idx, occ=occ)
        # SA index -> text position (BWTSaValue) on the forward BWT: every residue class, the ends, inverseSa0
        sidx = rng.integers(0, ix.fwd.text_length + 1, size=6000).astype(np.uint32)
        sidx[:8] = [0, 1, 7, 8, ix.fwd.text_length, ix.fwd.text_length - 1, ix.fwd.inverse_sa0, ix.fwd.inverse_sa0 + 1]
        with open(os.path.join(td, "sidx.bin"), "wb") as f:
            np.asarray([sidx.shape[0]], dtype=np.uint32).tofile(f)
            sidx.tofile(f)
        ol.run_ref(["sa", prefix, os.path.join(td, "sidx.bin"), os.path.join(td, "sa.out")])
        sa = np.fromfile(os.path.join(td, "sa.out"), dtype=np.uint32)[1:].reshape(-1, 2)
        arrays.update(sa_idx=sidx, sa_val=sa[:, 0].copy(), sa_steps=sa[:, 1].copy())
        for name, (spec, okw) in CASES.items():
            rs = make_reads(genome, spec)
            rp = os.path.join(td, name + ".reads")
            synth.write_reads_bin(rp, rs)
            opt = ol.default_opt(**okw)
            info = dict(reads=spec, opt=okw, reads_digest=digest(rs.codes))
            for mode in ("percall", "whole", "seeds"):
                outp = os.path.join(td, f"{name}.{mode}.aln")
                j = ol.run_ref([mode, prefix, rp, outp] + ol.opt_args(opt), count=True)
                n_aln, rows = synth.read_aln_dump(outp)
                arrays[f"{name}.{mode}.n_aln"] = n_aln.astype(np.int32)
                arrays[f"{name}.{mode}.rows"] = rows
                info[mode] = dict(lookups=int(j["occ4"] + j["occ1"]), hits=int(n_aln.sum()))
            # widths of the first 64 reads (type 1)
            sub = rs.subset(0, min(64, rs.n))
            wp = os.path.join(td, name + ".w.reads")
            synth.write_reads_bin(wp, sub)
            ol.run_ref(["width", prefix, wp, 1, os.path.join(td, name + ".w")])
            arrays[f"{name}.width"] = np.fromfile(os.path.join(td, name + ".w"), dtype=np.uint32)
            meta["cases"][name] = info
        # multi-record genome: BWTRetrievePositionFromSAIndex = BWTSaValue + the block search over the annotation
        recs, text2 = make_genome2()
        td2 = os.path.join(td, "g2"); os.mkdir(td2)
        with open(os.path.join(td2, "m.fa"), "wb") as f:
            for i, r in enumerate(recs):
                f.write(f">rec{i}\n".encode() + np.frombuffer(b"ACGT", dtype=np.uint8)[r].tobytes() + b"\n")
        subprocess.run([ol.REF_BIN, "index", "m", "m.fa"], cwd=td2, check=True, stdout=subprocess.DEVNULL)
        ix2 = index_io.load_index(os.path.join(td2, "m"))
        assert ix2.fwd.text_length == text2.shape[0]
        meta["genome2"] = GENOME2
        meta["index2"] = index_digest(ix2)
        lidx = np.arange(1, ix2.fwd.text_length + 1, dtype=np.uint32)          # every SA index but 0 (SA[0] = -1: no block)
        with open(os.path.join(td2, "lidx.bin"), "wb") as f:
            np.asarray([lidx.shape[0]], dtype=np.uint32).tofile(f)
            lidx.tofile(f)
        ol.run_ref(["locate", os.path.join(td2, "m"), os.path.join(td2, "lidx.bin"), os.path.join(td2, "loc.out")])
        arrays["loc_out"] = np.fromfile(os.path.join(td2, "loc.out"), dtype=np.uint32)[1:].reshape(-1, 3)
        arrays["loc_blocks"] = ix2.blocks.table()
        out = subprocess.run([ol.REF_BIN, "maxdiff", "400", "0.04"], check=True, capture_output=True, text=True).stdout
        meta["maxdiff"] = {int(a): int(b) for a, b in (ln.split() for ln in out.strip().splitlines())}
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote golden.npz", os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes")
    for k, v in meta["cases"].items():
        print(k, {m: v[m] for m in ("percall", "whole", "seeds")})


if __name__ == "__main__":
    main()

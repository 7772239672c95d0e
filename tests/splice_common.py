"""Shared fixtures of the splice-path tests: the golden genome / reads / expected outputs of tests/golden/golden_splice.*"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import emu_lib as el  # noqa: E402
import oracle_lib as ol  # noqa: E402
from hsa_b200 import index_build, index_io  # noqa: E402


class GoldenSplice:
    def __init__(self):
        import make_golden_splice as mg
        self.mg = mg
        with open(os.path.join(HERE, "golden", "golden_splice.json")) as f:
            self.meta = json.load(f)
        assert self.meta["genome"] == mg.GENOME, "golden genome spec drifted: regenerate tests/golden/golden_splice.*"
        self.arr = np.load(os.path.join(HERE, "golden", "golden_splice.npz"))
        self.genome, self.introns = mg.make_genome()
        assert mg.digest(self.genome) == self.meta["genome_digest"], "genome generator drifted"
        self._index = None

    @property
    def cases(self):
        return list(self.meta["cases"].keys())

    def index(self, device="cpu"):
        """The full index the splice path needs, made by the product's own code: both BWTs + SA samples (index_build),
        the block list of a one-record FASTA, the packed text -- checked against what the reference builder wrote."""
        if self._index is None:
            ix = index_build.build_index(self.genome, device=device)
            ix.blocks = index_io.blocks_of_records([self.genome.shape[0]])
            ix.packed_dna, ix.dna_length = index_io.pack_dna(self.genome), int(self.genome.shape[0])
            assert ix.blocks.table().tolist() == self.meta["blocks"] and ix.dna_length == self.meta["dna_length"]
            assert self.mg.digest(ix.packed_dna[: (ix.dna_length + 15) // 16]) == self.meta["packed_dna_digest"]
            self._index = ix
        return self._index

    def reads(self, name):
        c = self.meta["cases"][name]
        rs = self.mg.make_reads(self.genome, self.introns, c["reads"])
        assert self.mg.digest(rs.codes) == c["reads_digest"], "read generator drifted"
        return rs

    def opts(self, name, rs):
        """(per-length resolved option sets, opt_idx per read): aux->opt as the driver holds it for each read."""
        c = self.meta["cases"][name]
        opt = ol.default_opt(**c["opt"])
        lens = sorted(set(rs.lens.tolist()))
        opts = [el.resolve_read_opt(opt, L, c["clear_gape"]) for L in lens]
        return opts, np.asarray([lens.index(int(x)) for x in rs.lens], dtype=np.uint32)

    def expected(self, name):
        return self.arr[name + ".n_aln"], self.arr[name + ".rows"]

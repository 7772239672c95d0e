"""Shared helpers of the test-suite: golden fixtures and seeded inputs."""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "golden"))

from hsa_b200 import synth  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class Golden:
    """tests/golden/golden.{json,npz}: outputs of the unmodified reference on seeded inputs."""

    def __init__(self):
        with open(os.path.join(HERE, "golden", "golden.json")) as f:
            self.meta = json.load(f)
        self.arr = np.load(os.path.join(HERE, "golden", "golden.npz"))
        self.genome = synth.make_repeat_genome(**self.meta["genome"])
        self._reads = {}

    def genome2(self):
        """(records, concatenated text) of the multi-record golden genome (make_golden.make_genome2)."""
        import make_golden
        assert make_golden.GENOME2 == self.meta["genome2"], "multi-record genome spec drifted"
        return make_golden.make_genome2()

    @property
    def cases(self):
        return list(self.meta["cases"].keys())

    def reads(self, name: str) -> synth.ReadSet:
        if name not in self._reads:
            import make_golden
            rs = make_golden.make_reads(self.genome, self.meta["cases"][name]["reads"])
            assert digest(rs.codes) == self.meta["cases"][name]["reads_digest"], "input generator drifted"
            self._reads[name] = rs
        return self._reads[name]

    def opt_kwargs(self, name: str) -> dict:
        return dict(self.meta["cases"][name]["opt"])

    def expected(self, name: str, mode: str):
        return self.arr[f"{name}.{mode}.n_aln"], self.arr[f"{name}.{mode}.rows"]

    def lookups(self, name: str, mode: str) -> int:
        return int(self.meta["cases"][name][mode]["lookups"])

    def widths(self, name: str):
        """[(bid, width[len+1, 2]) ...] of the first reads of the case, from the reference's bwt_cal_width."""
        raw = self.arr[f"{name}.width"]
        n = int(raw[1])
        out, p = [], 2
        for _ in range(n):
            bid, m = int(raw[p]), int(raw[p + 1])
            p += 2
            out.append((bid, raw[p:p + 2 * m].reshape(m, 2)))
            p += 2 * m
        return out

"""CPU (tests/emu): share of the search's occ-lookup pairs whose two lookups (at k and at l + 1) fall into the same 32-byte
index sector -- SURVEY.md section 8d's "deduplicated figure".  TEST INFRASTRUCTURE (host build of hsa_core.cuh).
    python tools/dedup_model.py 46000003 4000"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import emu_lib as el, oracle_lib as ol
from hsa_b200 import index_build, synth
G, n = int(sys.argv[1]), int(sys.argv[2])
genome = synth.make_genome(G, 1)
emu = el.Emu(index_build.build_index(genome, device="cpu", sa_interval=0))
rs = synth.simulate_reads(genome, n, 100, 1000)
o = (C.c_uint64 * 2)()
el.lib().emu_pair_stats(o)
emu.whole(rs, ol.default_opt(), arena_cap=1022, hit_cap=32, rerun_cap=65535)
el.lib().emu_pair_stats(o)
print(f"genome {G} bp, {n} reads: {o[0]} lookup pairs, {o[1]} with both lookups in one sector = {100.0 * o[1] / o[0]:.1f} %; "
      f"distinct sectors per pair = {2 - o[1] / o[0]:.3f}")

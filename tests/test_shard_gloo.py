"""World-size-2 gloo test of the multi-GPU host logic: contiguous read sharding, one-time index
block broadcast, ordered gather.  The search itself is stood in for by the host emulation of the device
state machine (tests/emu, test infrastructure) -- the point here is the plumbing around it: sharded results,
re-assembled in input order, must equal the single-process results."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from hsa_b200 import shard  # noqa: E402


def test_shard_bounds_cover_and_align():
    for n in (0, 1, 7, 100000, 250001, 1000003):
        for world in (1, 2, 3, 8):
            for align in (1, shard.REF_BATCH):
                b = shard.shard_bounds(n, world, align)
                assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
                assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
                assert all(lo % align == 0 for lo, hi in b if lo < n)
                sizes = [hi - lo for lo, hi in b]
                assert max(sizes) - min(sizes) < 2 * align


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import common
        import emu_lib as el
        import oracle_lib as ol
        from hsa_b200 import index_build
        g = common.Golden()
        rs = g.reads("cfg2_100bp_default")
        opt = ol.default_opt(**g.opt_kwargs("cfg2_100bp_default"))
        # "index broadcast": rank 0 owns the packed bytes, everyone receives a replica
        index = index_build.build_index(g.genome, device="cpu") if rank == 0 else None
        payload = torch.from_numpy(np.concatenate([index.fwd.bwt_code, index.rev.bwt_code]).view(np.uint8).copy()) if rank == 0 else None
        nbytes = torch.tensor([payload.numel() if rank == 0 else 0])
        dist.broadcast(nbytes, 0)
        got = shard.broadcast_bytes(payload, int(nbytes.item()), 0, "cpu")
        if rank != 0:
            index = index_build.build_index(g.genome, device="cpu")       # replica; check it equals the broadcast bytes
        mine = np.concatenate([index.fwd.bwt_code, index.rev.bwt_code]).view(np.uint8)
        assert np.array_equal(got.numpy(), mine)
        off = rs.offsets[:-1].astype(np.uint64)
        codes, off_l, lens_l, (lo, hi) = shard.shard_reads(rs.codes, off, rs.lens, rank, world)
        from hsa_b200 import synth
        sub = synth.ReadSet(lens_l.copy(), codes.copy())
        n_aln, rows, status = el.Emu(index).whole(sub, opt)
        rows9 = np.zeros((rows.shape[0], 9), dtype=np.uint32)
        rows9[:, :] = rows[:, [0, 3, 4, 5, 6, 8, 9, 10, 11]]            # any fixed 9 columns: payload identity is what is checked
        n_all, a_all = shard.gather_in_input_order(n_aln, rows9, dst=0)
        # the C ABI's flat result form (arena + per-item offsets; arena order arbitrary): reverse this rank's arena
        cnt = n_aln.astype(np.int64)
        first = np.concatenate([[0], np.cumsum(cnt)[:-1]])
        arena = rows9[::-1].copy()
        aln_off = (rows9.shape[0] - first - cnt).astype(np.uint64)
        for i in np.nonzero(cnt > 1)[0]:                               # hits of one item stay in discovery order
            o, c = int(aln_off[i]), int(cnt[i])
            arena[o:o + c] = arena[o:o + c][::-1]
        n2, off2, a2 = shard.gather_results(n_aln, aln_off, arena, dst=0)
        # the host-side gather through shared memory (what bench.py's end-to-end leg uses)
        hg = shard.HostGather(max_items=rs.n, max_hits=2 * rs.n + 8)
        got = hg.gather(n_aln, aln_off, arena)
        if rank == 0:
            n3, off3, a3 = got
            exp_n3, exp_rows3 = g.expected("cfg2_100bp_default", "whole")
            assert np.array_equal(n3, exp_n3)
            idx3 = np.concatenate([np.arange(int(o), int(o) + int(c)) for o, c in zip(off3, n3) if c])
            assert np.array_equal(a3[idx3], exp_rows3[:, [0, 3, 4, 5, 6, 8, 9, 10, 11]])
            hg.release()
        hg.close()
        if rank == 0:
            exp_n, exp_rows = g.expected("cfg2_100bp_default", "whole")
            exp9 = exp_rows[:, [0, 3, 4, 5, 6, 8, 9, 10, 11]]
            assert np.array_equal(n_all, exp_n)
            assert np.array_equal(a_all, exp9)
            assert np.array_equal(n2, exp_n)
            idx = np.concatenate([np.arange(int(o), int(o) + int(c)) for o, c in zip(off2, n2) if c])
            assert np.array_equal(a2[idx], exp9)
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()

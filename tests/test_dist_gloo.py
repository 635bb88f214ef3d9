"""world_size-2 (and 3) gloo runs of the sharded matcher's host logic on CPU: shard bounds,
the fixed-size record all-gather, overflow regrowth and catalogue-order merge.  The per-rank
matcher is a CPU stand-in built on the oracle (tests may use oracle/; the product may not)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from tvidz_b200 import synth
from tvidz_b200.dist import ShardedCatalogue, merge_records, shard_bounds, take_shard


class OracleShard:
    """CPU stand-in for Catalogue.match_async, same record format."""

    def __init__(self, ts, off, vid):
        self.ts, self.off, self.vid = ts, off, vid

    def match_async(self, q, min_match, out):
        counts = oracle.match_counts(self.ts, self.off, np.asarray(q, np.float64))
        keep = np.nonzero(counts >= min_match)[0]
        cap = out.shape[0] - 1
        out.zero_()
        out[0, 0] = min(len(keep), 2**31 - 1)
        out[0, 1] = int(len(keep) > cap)
        n = min(len(keep), cap)
        out[1:1 + n, 0] = torch.from_numpy(self.vid[keep[:n]].astype(np.int32))
        out[1:1 + n, 1] = torch.from_numpy(counts[keep[:n]].astype(np.int32))


class OracleFragmentShard:
    """CPU stand-in for FragmentCatalogue.match_async, same int32 [3 * (cap + 1)] record."""

    def __init__(self, ts, off, vid):
        self.ts, self.off, self.vid = ts, off, vid

    def match_async(self, q, min_match, out, **kw):
        score, delta = oracle.fragment_rows(self.ts, self.off, np.asarray(q, np.float64), **kw)
        keep = np.nonzero(score >= min_match)[0]
        cap = out.numel() // 3 - 1
        out.zero_()
        out[0] = min(len(keep), 2**31 - 1)
        out[1] = int(len(keep) > cap)
        n = min(len(keep), cap)
        rec = out[2:2 + 2 * n].view(n, 2)
        rec[:, 0] = torch.from_numpy(self.vid[keep[:n]].astype(np.int32))
        rec[:, 1] = torch.from_numpy(score[keep[:n]].astype(np.int32))
        out[2 * (cap + 1) + 1: 2 * (cap + 1) + 1 + n] = torch.from_numpy(delta[keep[:n]].astype(np.int32))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q_out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ts, off, vid = synth.synth_catalogue(5000, seed=21)
        sc = ShardedCatalogue(ts, off, vid, hit_capacity=8, local_factory=OracleShard)
        results = []
        for r, mm in ((17, 2), (4999, 5), (2500, 0), (0, 1)):
            q = ts[off[r]:off[r + 1]]
            got = sc.find_duplicates(q, mm)
            want = oracle.find_duplicates_csr(ts, off, vid, q, mm)
            results.append(got == want)
        results.append(sc.cap >= 5000 // world)           # grew past the tiny initial capacity
        # fragment mode through the same gather: 30 s clip out of row 123 of a long-row catalogue
        from tvidz_b200.dist import ShardedFragmentCatalogue
        from tvidz_b200.fragment import clip_query, rank_fragments
        fts, foff, fvid = synth.synth_catalogue(240, len_range=(300, 700), gap_range=(15, 150), seed=5)
        fq = clip_query(fts[foff[123]:foff[124]], 9000)
        fc = ShardedFragmentCatalogue(fts, foff, fvid, hit_capacity=4, local_factory=OracleFragmentShard)
        for mm, k in ((3, None), (len(fq), 1), (2, 5)):
            got = fc.find_fragments(fq, mm, top_k=k)
            sc_, dl_ = oracle.fragment_rows(fts, foff, np.asarray(fq))
            keep = np.nonzero(sc_ >= mm)[0]
            want = rank_fragments(fvid[keep], sc_[keep], dl_[keep], 1000.0, k)
            results.append(got == want and len(got) >= 1)
        results.append(fc.find_fragments(fq, len(fq), top_k=1)[0][0] == int(fvid[123]))
        q_out.put((rank, results, sc.bounds))
    except Exception as e:  # surface the failure instead of letting the parent time out
        q_out.put((rank, [False, repr(e)], []))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_find_duplicates_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    outs = [q.get(timeout=180) for _ in range(world)]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    bounds = outs[0][2]
    for rank, results, b in outs:
        assert all(results), (rank, results)
        assert b == bounds
    assert bounds[0][0] == 0 and bounds[-1][1] == 5000
    assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))


def test_shard_bounds_balance_by_values():
    ts, off, vid = synth.synth_catalogue(10_000, seed=3)
    for world in (1, 2, 4, 8):
        b = shard_bounds(off, world)
        assert b[0][0] == 0 and b[-1][1] == 10_000
        sizes = [int(off[hi] - off[lo]) for lo, hi in b]
        assert max(sizes) - min(sizes) <= 2 * 120          # within a couple of rows of each other
        parts = [take_shard(ts, off, vid, lo, hi) for lo, hi in b]
        assert np.array_equal(np.concatenate([p[0] for p in parts]), ts)
        assert np.array_equal(np.concatenate([p[2] for p in parts]), vid)
        assert all(p[1][0] == 0 and p[1][-1] == p[0].shape[0] for p in parts)
    # degenerate: more ranks than rows, empty catalogue
    assert shard_bounds(np.array([0, 3]), 4)[-1][1] == 1
    assert shard_bounds(np.array([0]), 2) == [(0, 0), (0, 0)]


def test_merge_records_order_and_overflow():
    cap = 3
    g = np.zeros((2, cap + 1, 2), np.int32)
    g[0, 0] = (2, 0); g[0, 1] = (5, 9); g[0, 2] = (6, 2)
    g[1, 0] = (1, 0); g[1, 1] = (70, 4)
    pairs, over, need = merge_records(g, cap)
    assert pairs.tolist() == [[5, 9], [6, 2], [70, 4]] and not over and need == 2
    g[1, 0] = (9, 1)
    pairs, over, need = merge_records(g, cap)
    assert over and need == 9

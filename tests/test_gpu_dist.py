"""The sharded matcher on real GPUs: NCCL all-gather and the fused peer-store gather, against the
C oracle.  Runs with as many ranks as there are GPUs (1 on the round-end box, where the fused path
still exercises its kernels: symmetric memory, last-block record store, flag release, wait kernel)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q_out):
    import numpy as np
    import torch.distributed as dist

    import oracle
    from tvidz_b200 import synth
    from tvidz_b200.dist import ShardedCatalogue, ShardedFragmentCatalogue
    from tvidz_b200.fragment import clip_query
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok = []
    try:
        ts, off, vid = synth.synth_catalogue(60_000, seed=33)
        for gather in ("nccl", "fused"):
            sc = ShardedCatalogue(ts, off, vid, hit_capacity=64, device=rank, gather=gather)
            for r, mm in ((17, 2), (59_999, 5), (30_000, 1), (4, 2), (4, 2)):
                q = ts[off[r]:off[r + 1]]
                ok.append((gather, r, mm, sc.find_duplicates(q, mm) == oracle.find_duplicates_csr(ts, off, vid, q, mm)))
            ok.append((gather, "grew", sc.cap > 64))
            for _ in range(20):                                   # back-to-back epochs without host reads
                g = sc.enqueue(ts[off[9]:off[10]], 2)
            torch.cuda.synchronize()
            want = oracle.find_duplicates_csr(ts, off, vid, ts[off[9]:off[10]], 2)
            ok.append((gather, "pipelined", int(g[:, 0, 0].sum().item()) == len(want)))
            # 8 queries per pass over every shard (and a second, ragged group), full lists against the oracle
            picks = [3, 17, 4000, 29_999, 30_000, 45_678, 59_999, 12, 31, 50_000, 7]
            qs = [ts[off[r]:off[r + 1]] for r in picks] + [np.zeros(0)]
            many = sc.find_duplicates_many(qs, 2)
            ok.append((gather, "match_many", many == [oracle.find_duplicates_csr(ts, off, vid, q, 2) for q in qs]))
        fts, foff, fvid = synth.synth_catalogue(3000, len_range=(600, 1400), gap_range=(15, 150), seed=34)
        fq = clip_query(fts[foff[1234]:foff[1235]], 20_000)
        want = oracle.find_fragments_csr(fts, foff, fvid, fq, min_match=4)
        for gather in ("nccl", "fused"):
            fc = ShardedFragmentCatalogue(fts, foff, fvid, hit_capacity=8, device=rank, gather=gather)
            for _ in range(3):                                    # consecutive epochs alternate buffer sets
                got = fc.find_fragments(fq, 4)
                ok.append(("fragment-" + gather, [(v, s, round(o * 1000)) for v, s, o in got] == want and len(want) >= 1))
            ok.append(("fragment-" + gather + "-anchor1",
                       [(v, s, round(o * 1000)) for v, s, o in fc.find_fragments(fq, 4, anchor=1)]
                       == oracle.find_fragments_csr(fts, foff, fvid, fq, min_match=4, anchor=1)))
        q_out.put((rank, ok))
    except Exception as e:
        q_out.put((rank, [("exception", repr(e), False)]))
        raise
    finally:
        dist.destroy_process_group()


def test_sharded_matcher_nccl_and_fused(cuda):
    world = min(torch.cuda.device_count(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    outs = [q.get(timeout=300) for _ in range(world)]
    [p.join(60) for p in procs]
    for rank, ok in outs:
        assert all(item[-1] for item in ok), (rank, [item for item in ok if not item[-1]])
    assert all(p.exitcode == 0 for p in procs)

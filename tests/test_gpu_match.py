"""Stage 2 parity on the GPU: Catalogue / Inspector (through the C ABI) against the golden
vectors produced by the reference's own code and against the CPU oracle."""
import threading

import numpy as np
import pytest

import oracle
from oracle import match_oracle
from tvidz_b200 import synth
from tvidz_b200.catalog import Catalogue, rows_to_csr
from tvidz_b200.inspector import Inspector

pytestmark = pytest.mark.gpu


def _rows(case):
    return [(vid, ts) for vid, (_, ts) in zip(case["video_ids"], case["catalogue"])]


def test_golden_vectors_through_catalogue(cuda, match_golden):
    for c in match_golden:
        cat = Catalogue.from_rows(_rows(c))
        mm = 5 if c["min_match"] is None else c["min_match"]
        got = cat.find_duplicates(c["query"], mm)
        assert [list(x) for x in got] == c["expected"], c["name"]
        assert all(type(v) is int and type(n) is int for v, n in got)
        cat.close()


def test_reference_test_through_dropin_api(cuda):
    """inspector/test_app.py:66-84 verbatim against the drop-in module."""
    from tvidz_b200 import inspector as db
    db.clear_db()
    v1 = db.add_video('a.mp4')
    v2 = db.add_video('b.mp4')
    db.add_timestamps(v1.id, [1.0, 2.0, 3.0, 4.0, 5.0])
    db.add_timestamps(v2.id, [10.0, 20.0, 30.0, 40.0, 50.0])
    dups = db.find_duplicates([10.0, 20.0, 30.0, 40.0, 50.0], min_match=5)
    assert (v1.id, 0) not in dups
    assert (v2.id, 5) in dups
    v3 = db.add_video('c.mp4')
    db.add_timestamps(v3.id, [1.0, 2.0, 3.0, 4.0, 5.0])
    dups = db.find_duplicates([1.0, 2.0, 3.0, 4.0, 5.0], min_match=5)
    assert (v1.id, 5) in dups
    assert (v3.id, 5) in dups
    assert dups == [(v1.id, 5), (v3.id, 5)]
    db.add_timestamps(v1.id, [9.0])                      # upsert replaces the row (db.py:54-57)
    assert db.find_duplicates([1.0, 2.0, 3.0, 4.0, 5.0]) == [(v3.id, 5)]
    assert db.find_duplicates([1.2, 5.7, 12.3, 18.9], min_match=2) == []
    db.clear_db()


def test_streaming_golden_through_inspector(cuda, stream_golden):
    for c in stream_golden:
        ins = Inspector()
        for fn, ts in c["catalogue"]:
            v = ins.add_video(fn)
            ins.add_timestamps(v.id, ts)
        me = ins.add_video("upload")
        assert me.id == c["self_video_id"]
        cuts, ids = ins.analyze_cuts(me.id, c["tokens"], min_match=2)
        res = c["result"]
        assert cuts == res["scene_cuts"] and len(cuts) == res["total_cuts"], c["name"]
        assert ids == c["stored_duplicates"] == ins.get_video_by_id(me.id).duplicates, c["name"]
        assert {ins.get_video_by_id(i).filename for i in ids} == set(res["duplicates"]), c["name"]
        assert ins._rows[me.id] == res["scene_cuts"]


@pytest.mark.parametrize("n_rows,seed", [(1, 0), (1000, 1), (200_000, 2)])
def test_random_catalogue_vs_c_oracle(cuda, n_rows, seed):
    ts, off, vid = synth.synth_catalogue(n_rows, seed=seed)
    cat = Catalogue(ts, off, vid)
    rng = np.random.default_rng(seed)
    for mm in (1, 2, 5, 0):
        r = int(rng.integers(n_rows))
        q = ts[off[r]:off[r + 1]].copy()
        want = oracle.find_duplicates_csr(ts, off, vid, q, mm)
        got = cat.find_duplicates(q, mm)
        assert got == want
        assert (int(vid[r]), len(q)) in got
    cat.close()


def test_kth_matches_oracle(cuda):
    ts, off, vid = synth.synth_catalogue(50_000, seed=4, gap_range=(1, 60))
    cat = Catalogue(ts, off, vid)
    rng = np.random.default_rng(4)
    for mm in (1, 2, 3):
        r = int(rng.integers(50_000))
        q = ts[off[r]:off[r + 1]].copy()
        v, c, k = cat.match(q, mm, with_kth=True)
        counts = oracle.match_counts(ts, off, q)
        kth = oracle.match_kth(ts, off, q, mm)
        keep = np.nonzero(counts >= mm)[0]
        assert np.array_equal(v, vid[keep]) and np.array_equal(c, counts[keep]) and np.array_equal(k, kth[keep])
        assert len(keep) > 1
    cat.close()


def test_unsorted_rows_repeats_and_specials(cuda):
    nan = float("nan")
    rows = [(7, [3.0, 1.0, 3.0, 3.0, 2.0]), (8, [nan, -0.0, 0.0, nan]), (9, []), (10, [float("inf"), 5e-324]),
            (11, [2.0, 2.0])]
    q = [3.0, 3.0, 0.0, -0.0, nan, 2.0, float("inf"), 5e-324, 4.0]
    cat = Catalogue.from_rows(rows)
    for mm in (-1, 0, 1, 2, 3, 4):
        assert cat.find_duplicates(q, mm) == match_oracle.find_duplicates(match_oracle.hydrate(rows), q, mm)
    assert cat.n_values == 3 + 1 + 0 + 2 + 1           # canonical rows: repeats, NaN and -0.0 folded
    v, c, k = cat.match(q, 2, with_kth=True)
    ts, off, vid = rows_to_csr(rows)
    assert np.array_equal(k, oracle.match_kth(ts, off, q, 2)[np.isin(vid, v)])
    cat.close()


def test_query_longer_than_one_launch(cuda):
    """More distinct query values than one launch holds (2048): the count accumulates."""
    ts, off, vid = synth.synth_catalogue(3000, len_range=(1, 50), gap_range=(1, 300), seed=6)
    cat = Catalogue(ts, off, vid)
    q = np.unique(ts)[:5000]
    assert q.shape[0] > 4096
    q = np.concatenate([q, q[:100]])                    # plus repeats
    assert cat.find_duplicates(q, 30) == oracle.find_duplicates_csr(ts, off, vid, q, 30)
    cat.close()


def test_hit_list_grows_past_capacity(cuda):
    ts, off, vid = synth.synth_catalogue(20_000, seed=8)
    cat = Catalogue(ts, off, vid, hit_capacity=16)
    q = ts[off[5]:off[6]]
    got = cat.find_duplicates(q, 0)                     # min_match <= 0 returns every row (B.2)
    assert len(got) == 20_000 and got == oracle.find_duplicates_csr(ts, off, vid, q, 0)
    assert cat.find_duplicates(q, 5) == oracle.find_duplicates_csr(ts, off, vid, q, 5)
    cat.close()


def test_empty_catalogue_and_empty_query(cuda):
    cat = Catalogue.from_rows([])
    assert cat.find_duplicates([1.0, 2.0], 0) == [] and cat.find_duplicates([], 5) == []
    cat.close()
    cat = Catalogue.from_rows([(1, [1.0]), (2, [])])
    assert cat.find_duplicates([], 0) == [(1, 0), (2, 0)] and cat.find_duplicates([], 1) == []
    cat.close()


def test_concurrent_queries_from_threads(cuda):
    """One analysis thread per upload (app.py:43,472): calls are independent."""
    ts, off, vid = synth.synth_catalogue(30_000, seed=9)
    cat = Catalogue(ts, off, vid)
    want, got = {}, {}
    rows = [11, 222, 3333, 4444, 25_000, 29_999]
    for r in rows:
        want[r] = oracle.find_duplicates_csr(ts, off, vid, ts[off[r]:off[r + 1]], 2)

    def work(r):
        for _ in range(5):
            got[r] = cat.find_duplicates(ts[off[r]:off[r + 1]], 2)

    th = [threading.Thread(target=work, args=(r,)) for r in rows]
    [t.start() for t in th]
    [t.join() for t in th]
    assert got == want
    cat.close()


def test_full_size_million_rows(cuda):
    """BASELINE config 4: one query against 1M rows.  Checked against the C oracle (seconds)
    and through size-independent properties (self-match, min_match monotonicity)."""
    ts, off, vid = synth.synth_catalogue(1_000_000, seed=0)
    cat = Catalogue(ts, off, vid, hit_capacity=1 << 17)
    r = 123_456
    q = ts[off[r]:off[r + 1]].copy()
    got2 = cat.find_duplicates(q, 2)
    assert got2 == oracle.find_duplicates_csr(ts, off, vid, q, 2)
    got5 = cat.find_duplicates(q, 5)
    assert (int(vid[r]), len(q)) in got2 and (int(vid[r]), len(q)) in got5
    assert set(got5) <= set(got2) and len(got2) > len(got5) >= 1
    assert cat.find_duplicates(q, len(q)) == [(int(vid[r]), len(q))]
    cat.close()


def test_reference_style_per_cut_loop_is_a_device_upsert(cuda, stream_golden):
    """The unmodified shape of app.py:228-255 -- add_timestamps() then find_duplicates() after every
    new cut -- against a 30k-row catalogue: same verdicts as the oracle's streaming loop, and the
    catalogue is packed exactly once (row rewrites are device-side upserts into the tail)."""
    ts, off, vid = synth.synth_catalogue(30_000, seed=12)
    ins = Inspector()
    rows = []
    for r in range(30_000):
        row = ts[off[r]:off[r + 1]].tolist()
        rows.append((r + 1, row))
    ins.load_rows(rows)
    assert ins._next_id == 30_001
    assert ins.find_duplicates([1.0], 1) is not None and ins.repacks == 1
    for src in (10, 20_000, 29_999):
        me = ins.add_video("upload-%d.mp4" % src)
        tokens = ["%.6g" % t for t in rows[src][1]]
        scene, dups = [], []
        for tok in tokens:                                   # app.py:228-255, verbatim shape
            t = float(tok)
            if not scene or t != scene[-1]:
                scene.append(t)
                ins.add_timestamps(me.id, scene)
                dups = [d for d in ins.find_duplicates(scene, min_match=2) if d[0] != me.id]
                if dups:
                    break
        o_scene, o_ids, _ = match_oracle.streaming_analysis(rows, me.id, tokens, 2)
        assert scene == o_scene and [d[0] for d in dups] == o_ids and src + 1 in o_ids
        rows.append((me.id, list(scene)))                    # the truncated row stays in the catalogue (Q5)
    # rewriting a packed row hides its old content and serves the new one
    ins.add_timestamps(11, [123456.5, 123457.5])
    assert ins.find_duplicates(rows[10][1], 5) == [d for d in
                                                   match_oracle.find_duplicates(rows, rows[10][1], 5) if d[0] != 11]
    assert (11, 2) in ins.find_duplicates([123456.5, 123457.5], 2)
    assert ins.repacks == 1
    info = ins._cat.tail_info()
    assert info["rows"] == 4 and info["replaced_packed_rows"] == 1      # 3 uploads (each rewritten in place) + row 11
    small = Inspector(tail_values=512, hit_capacity=64)     # a tail of one unit: fills, is compacted, finally repacks
    for i in range(300):
        v = small.add_video("v%d" % i)
        small.add_timestamps(v.id, [float(i), float(i) + 0.5])
        assert small.find_duplicates([float(i), float(i) + 0.5], 2) == [(v.id, 2)]
    assert small.repacks == 2 and small.find_duplicates([0.0, 0.5, 3.0, 3.5, 299.0, 299.5], 2) == [(1, 2), (4, 2), (300, 2)]


def _model_find(rows, q, mm, python=False):
    """rows: dict video_id -> list, in order of last write (what the Inspector promises)."""
    if python:
        return match_oracle.find_duplicates(match_oracle.hydrate(list(rows.items())), q, mm)
    return oracle.find_duplicates_csr(*rows_to_csr(list(rows.items())), q, mm)


@pytest.mark.parametrize("tail_values,seed", [(1 << 16, 1), (2048, 2)])
def test_device_upserts_against_a_host_model(cuda, tail_values, seed):
    """Random add_timestamps traffic (db.py:43-64: replace-or-append) straight on the mutable device
    catalogue -- packed rows replaced, tail rows replaced, the last row rewritten in place, rows longer
    than the kernel-parameter path, empty rows -- against a dict that replays the same writes.  The small
    tail forces tail compactions on the way."""
    rng = np.random.default_rng(seed)
    ts, off, vid = synth.synth_catalogue(5000, seed=40 + seed, len_range=(0, 40))
    model = {int(vid[r]): ts[off[r]:off[r + 1]].tolist() for r in range(5000)}
    cat = Catalogue(ts, off, vid, mutable=True, tail_values=tail_values, hit_capacity=8192)
    pool = np.unique(ts)
    last, repacks = None, 0
    for step in range(400):
        kind = rng.integers(0, 5)
        if kind == 0 or last is None:
            v = int(rng.integers(1, 5001))                               # a packed (or already moved) row
        elif kind == 1:
            v = last                                                     # the row written last: in place
        elif kind == 2:
            v = 6000 + int(rng.integers(0, 50))                          # new videos / earlier tail rows
        else:
            v = int(rng.choice(list(model.keys())))
        n = int(rng.choice([0, 1, 3, 17, 60, 300]))
        row = np.sort(rng.choice(pool, n, replace=False)).tolist() if n else []
        if n and rng.integers(0, 4) == 0:
            row = row + row[:2] + [float("nan"), -0.0]                   # repeats and specials are canonicalised
        ok = cat.upsert(v, row)
        model.pop(v, None)
        model[v] = row
        last = v
        if not ok:                                                       # tail full of live rows: repack, as the Inspector does
            cat.close()
            cat = Catalogue.from_rows(list(model.items()), mutable=True, tail_values=tail_values, hit_capacity=8192)
            repacks += 1
        if step % 7 == 0 or step > 390:
            probe = model[last] if model[last] else model[int(rng.choice(list(model.keys())))]
            for mm in (2, 1, 0):
                assert cat.find_duplicates(probe, mm) == _model_find(model, probe, mm), (step, mm)
    v, c, k = cat.match(model[last] or [1.0], 1, with_kth=True)          # the early-exit kernel reads tail rows too
    items = list(model.items())
    mts, moff, mvid = rows_to_csr(items)
    counts = oracle.match_counts(mts, moff, model[last] or [1.0])
    keep = np.nonzero(counts >= 1)[0]
    assert np.array_equal(v, mvid[keep]) and np.array_equal(k, oracle.match_kth(mts, moff, model[last] or [1.0], 1)[keep])
    got = cat.find_duplicates_many([model[last], items[3][1], []], 1)    # the batched kernel sees the same catalogue
    assert got == [_model_find(model, q, 1) for q in (model[last], items[3][1], [])]
    assert cat.find_duplicates(items[3][1], 1) == _model_find(model, items[3][1], 1, python=True)
    info = cat.tail_info()
    assert info["values"] <= info["capacity"] and (repacks > 0) == (tail_values == 2048)
    cat.close()


def test_upsert_overflow_reports_and_immutable_refuses(cuda):
    cat = Catalogue.from_rows([(1, [1.0, 2.0])], mutable=True, tail_values=512)
    for i in range(8):                                                   # 8 live rows x 64 values fill the 512-value tail
        assert cat.upsert(100 + i, [float(1000 * i + k) for k in range(64)])
    assert cat.upsert(200, [5.0]) is False                               # nothing to drop: the caller must repack
    assert cat.upsert(100, [7.0, 8.0]) is True                           # replacing a live row frees room (tail compaction)
    assert cat.find_duplicates([7.0, 8.0, 1.0, 2.0], 2) == [(1, 2), (100, 2)]
    cat.close()
    frozen = Catalogue.from_rows([(1, [1.0])])
    with pytest.raises(Exception):
        frozen.upsert(1, [2.0])
    frozen.close()


@pytest.mark.parametrize("shape", ["empty_rows", "giant_row", "short_rows", "two_values"])
def test_tile_boundaries(cuda, shape):
    """Catalogues whose rows fall awkwardly on the tile grid: thousands of empty rows, one row far larger
    than a tile's share, more rows than 4096 per tile would allow, and the smallest catalogue."""
    rng = np.random.default_rng(5)
    if shape == "empty_rows":
        lens = np.zeros(30_000, np.int64)
        lens[rng.integers(0, 30_000, 500)] = rng.integers(1, 90, 500)
    elif shape == "giant_row":
        lens = rng.integers(1, 30, 3000)
        lens[1500] = 400_000
    elif shape == "short_rows":
        lens = rng.integers(0, 3, 700_000)
    else:
        lens = np.array([1, 1], np.int64)
    off = np.zeros(lens.shape[0] + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    ts = np.round(rng.integers(0, 3_000_000, int(off[-1])) / 30.0, 4)
    vid = np.arange(1, lens.shape[0] + 1, dtype=np.int32)
    cat = Catalogue(ts, off, vid, hit_capacity=1 << 12)
    r = int(np.argmax(lens))
    queries = [ts[off[r]:off[r + 1]][:200], ts[rng.integers(0, ts.shape[0], 50)], ts[:1]]
    for q in queries:
        for mm in (1, 2, 0):
            assert cat.find_duplicates(q, mm) == oracle.find_duplicates_csr(ts, off, vid, q, mm), (shape, mm)
    got = cat.find_duplicates_many(queries, 1)
    assert got == [oracle.find_duplicates_csr(ts, off, vid, q, 1) for q in queries]
    cat.close()


def test_concurrent_callers_are_combined_into_batched_passes(cuda):
    """8 analysis threads (app.py:43,472) calling find_duplicates at once on one Inspector: every caller
    gets exactly its own answer, and the device answered several of them per catalogue pass."""
    ts, off, vid = synth.synth_catalogue(200_000, seed=17)
    ins = Inspector()
    ins.load_rows([(int(vid[r]), ts[off[r]:off[r + 1]].tolist()) for r in range(200_000)])
    ins.find_duplicates([1.0], 1)                                        # pack once, outside the race
    picks = [3, 1999, 50_000, 77_777, 120_000, 150_001, 180_000, 199_999]
    want = {r: oracle.find_duplicates_csr(ts, off, vid, ts[off[r]:off[r + 1]], 2) for r in picks}
    got, errs = {}, []
    start = threading.Barrier(len(picks))

    def work(r):
        try:
            start.wait()
            for i in range(20):
                res = ins.find_duplicates(ts[off[r]:off[r + 1]].tolist(), 2)
                if i == 7 and r == picks[0]:                             # an upsert in the middle of the traffic
                    ins.add_timestamps(999_999, [0.25, 0.75])
                got[r] = res
        except Exception as e:                                           # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(r,)) for r in picks]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs and got == want
    assert sum(ins.batches) == 1 + 20 * len(picks) and max(ins.batches) > 1 and ins.repacks == 1
    assert ins.find_duplicates([0.25, 0.75], 2) == [(999_999, 2)]


def test_batched_queries_equal_single_queries(cuda, match_golden):
    """Up to 8 queries per catalogue pass: identical to one find_duplicates call per query."""
    ts, off, vid = synth.synth_catalogue(120_000, seed=14)
    cat = Catalogue(ts, off, vid, hit_capacity=64)              # tiny capacity: exercises the regrowth path
    rng = np.random.default_rng(14)
    queries = [ts[off[r]:off[r + 1]].copy() for r in rng.integers(0, 120_000, 19)]
    queries += [np.zeros(0), np.array([float("nan"), 1.0, -0.0, 0.0]), np.array([5.0, 5.0, 5.0]),
                np.unique(ts)[:400]]                             # empty, specials, repeats, too long for a batch
    for mm in (2, 5, 0):
        qs = queries if mm else queries[:3]
        got = cat.find_duplicates_many(qs, mm)
        assert len(got) == len(qs)
        for q, g in zip(qs, got):
            assert g == oracle.find_duplicates_csr(ts, off, vid, q, mm)
    assert cat.find_duplicates_many([], 2) == []
    cat.close()
    for n_rows in (150, 20_001):                                 # row counts that are no multiple of 4 (or 16)
        t2, o2, v2 = synth.synth_catalogue(n_rows, seed=n_rows)
        cat = Catalogue(t2, o2, v2, hit_capacity=64)
        qs = [t2[o2[r]:o2[r + 1]].copy() for r in (0, n_rows // 2, n_rows - 1)]
        assert cat.find_duplicates_many(qs, 2) == [oracle.find_duplicates_csr(t2, o2, v2, q, 2) for q in qs]
        cat.close()
    for c in match_golden[:12]:                                  # the reference's own vectors, batched together
        rows = _rows(c)
        cat = Catalogue.from_rows(rows)
        mm = 5 if c["min_match"] is None else c["min_match"]
        got = cat.find_duplicates_many([c["query"]] * 3 + [[]], mm)
        assert [list(x) for x in got[0]] == c["expected"] and got[0] == got[1] == got[2]
        cat.close()


def test_pipelined_async_queries_keep_their_own_results(cuda):
    """Different queries enqueued back to back on one stream (no host wait in between; the kernels
    overlap their neighbours' tails through programmatic dependent launch) each fill their own
    record with exactly the oracle's hit list -- including dense queries that drain the survivor
    queue mid-stream and queries too long for the kernel parameters."""
    import torch
    ts, off, vid = synth.synth_catalogue(150_000, seed=21)
    cat = Catalogue(ts, off, vid, hit_capacity=1 << 17)
    rng = np.random.default_rng(21)
    dense = np.unique(ts[rng.integers(0, ts.shape[0], 1500)])           # ~1500 keys, 115k hit rows: upload path
    queries = [ts[off[r]:off[r + 1]] for r in (5, 77_000, 149_999)] + [dense, np.zeros(0), ts[off[9]:off[10]]]
    queries = queries * 3
    recs = [torch.zeros(((1 << 17) + 1, 2), dtype=torch.int32, device="cuda") for _ in queries]
    for q, rec in zip(queries, recs):
        cat.match_async(q, 2, rec)
    torch.cuda.synchronize()
    for q, rec in zip(queries, recs):
        r = rec.cpu().numpy()
        n = int(r[0, 0])
        assert r[0, 1] == 0
        assert [tuple(x) for x in r[1:1 + n].tolist()] == oracle.find_duplicates_csr(ts, off, vid, q, 2)
    cat.close()


def test_batches_in_both_parameter_block_sizes(cuda):
    """A batch's keys ride in the kernel parameters: 96 distinct values per query in the small block, up to 224
    in the large one; one long query moves the whole batch to the large block; 225 leaves the batch."""
    ts, off, vid = synth.synth_catalogue(60_000, seed=41)
    cat = Catalogue(ts, off, vid, hit_capacity=1 << 14)
    pool = np.unique(ts)
    rng = np.random.default_rng(41)
    pick = lambda k: pool[rng.choice(pool.shape[0], k, replace=False)]
    for sizes in ([96] * 8, [97] + [10] * 7, [224] * 8, [1, 96, 97, 223, 224, 0, 50, 150], [225, 224, 3]):
        qs = [np.concatenate([pick(k), pick(k)[:k // 3]]) if k else np.zeros(0) for k in sizes]   # some repeats on top
        qs = [np.concatenate([q, q[:5]]) for q in qs]                                           # ... and exact duplicates
        assert cat.find_duplicates_many(qs, 2) == [oracle.find_duplicates_csr(ts, off, vid, q, 2) for q in qs], sizes
    cat.close()


def test_exchange_words_survive_the_sequence_wrap(cuda):
    """Tile totals are exchanged as {16-bit query sequence, total}.  A batch of 8 leaves words for slots 1..7 that
    single queries never rewrite; 65535 queries later the sequence comes round again and those words must not
    look current (the state is reset when the sequence wraps)."""
    import torch
    ts, off, vid = synth.synth_catalogue(20_001, seed=43)
    cat = Catalogue(ts, off, vid, hit_capacity=1 << 12)
    rng = np.random.default_rng(43)
    qa = [ts[off[r]:off[r + 1]].copy() for r in rng.integers(0, 20_001, 8)]
    qb = [ts[off[r]:off[r + 1]].copy() for r in rng.integers(0, 20_001, 8)]
    rec8 = torch.zeros((8, (1 << 12) + 1, 2), dtype=torch.int32, device="cuda")
    rec = torch.zeros(((1 << 12) + 1, 2), dtype=torch.int32, device="cuda")
    cat.match_batch_async(qa, 1, rec8)                                  # sequence s on the asynchronous workspace
    q1 = qa[0][:3]
    for _ in range(65_534):
        cat.match_async(q1, 5, rec)
    cat.match_batch_async(qb, 1, rec8)                                  # sequence s again
    torch.cuda.synchronize()
    host = rec8.cpu().numpy()
    for b, q in enumerate(qb):
        n = int(host[b, 0, 0])
        assert [tuple(x) for x in host[b, 1:1 + n].tolist()] == oracle.find_duplicates_csr(ts, off, vid, q, 1)
    cat.close()


def test_fused_gather_loopback_with_eight_peers(cuda):
    """One GPU plays all 8 ranks of the fused gather: the kernel stores its tagged record into 8 slots of an
    ordinary device buffer (one "peer" each) and its last CTAs wait for 8 x n_queries lists -- the shape of an
    8-GPU run (64 lists for a batch), without the 8 GPUs.  Two epochs, so the second one meets stale tags."""
    import torch
    ts, off, vid = synth.synth_catalogue(20_001, seed=47)
    cap = 1 << 12
    cat = Catalogue(ts, off, vid, hit_capacity=cap)
    rng = np.random.default_rng(47)
    world = 8
    for nq in (1, 8):
        rec_ints = nq * (cap + 1) * 4
        buf = torch.zeros(world * rec_ints, dtype=torch.int32, device="cuda")
        peers = np.asarray([buf.data_ptr() + 4 * p * rec_ints for p in range(world)], np.uint64)
        for epoch in (1, 2):
            qs = [ts[off[r]:off[r + 1]].copy() for r in rng.integers(0, 20_001, nq)]
            if nq == 1:
                cat.match_gather_async(qs[0], 1, world, peers, buf.data_ptr(), rec_ints, cap, epoch)
            else:
                cat.match_batch_gather_async(qs, 1, world, peers, buf.data_ptr(), rec_ints, cap, epoch)
            torch.cuda.synchronize()
            host = buf.cpu().numpy().reshape(world, nq, cap + 1, 4)
            for b, q in enumerate(qs):
                want = oracle.find_duplicates_csr(ts, off, vid, q, 1)
                for p in range(world):
                    n = int(host[p, b, 0, 0])
                    assert n == len(want) and host[p, b, 0, 2] == 0
                    assert (host[p, b, :n + 1, 1::2] == epoch).all()
                    assert [tuple(x) for x in host[p, b, 1:1 + n, 0::2].tolist()] == want
    cat.close()

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


def test_reference_style_per_cut_loop_uses_the_overlay(cuda, stream_golden):
    """The unmodified shape of app.py:228-255 -- add_timestamps() then find_duplicates() after every
    new cut -- against a 30k-row catalogue: same verdicts as the oracle's streaming loop, and the big
    catalogue is packed exactly once (row rewrites go to the overlay + tombstones)."""
    ts, off, vid = synth.synth_catalogue(30_000, seed=12)
    ins = Inspector()
    rows = []
    for r in range(30_000):
        row = ts[off[r]:off[r + 1]].tolist()
        ins._rows[r + 1] = row
        rows.append((r + 1, row))
    ins._next_id = 30_001
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
    small = Inspector(overlay_limit=2)                      # the overlay folds into a fresh pack when it outgrows its limit
    for i in range(6):
        v = small.add_video("v%d" % i)
        small.add_timestamps(v.id, [float(i), float(i) + 0.5])
        assert small.find_duplicates([float(i), float(i) + 0.5], 2) == [(v.id, 2)]
    assert small.repacks == 2 and small.find_duplicates([0.0, 0.5, 3.0, 3.5], 2) == [(1, 2), (4, 2)]


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

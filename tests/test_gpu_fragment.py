"""Fragment mode on the GPU against this repo's own oracle (parity unpinned: the reference has no
fragment matcher), plus the one reference-anchored property: at offset 0 with zero tolerance the
score is find_duplicates' match_count."""
import numpy as np
import pytest

import oracle
from tvidz_b200 import synth
from tvidz_b200.catalog import Catalogue
from tvidz_b200.fragment import FragmentCatalogue, clip_query

pytestmark = pytest.mark.gpu


def _long_catalogue(n_rows, seed):
    return synth.synth_catalogue(n_rows, len_range=(600, 1400), gap_range=(15, 150), seed=seed)


def test_clip_is_found_at_its_offset(cuda):
    ts, off, vid = _long_catalogue(400, seed=11)
    cat = FragmentCatalogue(ts, off, vid)
    rng = np.random.default_rng(11)
    for _ in range(6):
        r = int(rng.integers(400))
        row = ts[off[r]:off[r + 1]]
        f0 = int(rng.integers(0, int(row[-1] * 30) - 900))
        q = clip_query(row, f0)
        if len(q) < 5:
            continue
        got = cat.find_fragments(q, min_match=len(q))
        want = oracle.find_fragments_csr(ts, off, vid, q, min_match=len(q))
        assert [(v, s) for v, s, _ in got] == [(v, s) for v, s, _ in want]
        hit = [g for g in got if g[0] == int(vid[r])]
        assert hit and hit[0][1] == len(q) and abs(hit[0][2] - f0 / 30.0) <= 0.008
    cat.close()


@pytest.mark.parametrize("anchor", [0, 1, 2, 3])          # 0 = every offset C[j] - Q[i] (SURVEY.md B.4 as written)
@pytest.mark.parametrize("min_match", [2, 3, 5])
def test_matches_oracle_all_rows(cuda, min_match, anchor):
    ts, off, vid = _long_catalogue(1500, seed=5)
    cat = FragmentCatalogue(ts, off, vid, hit_capacity=8)          # forces the capacity regrowth path
    row = ts[off[77]:off[78]]
    q = clip_query(row, 12_345)
    v, s, d = cat.match(q, min_match, anchor=anchor)
    want = oracle.find_fragments_csr(ts, off, vid, q, min_match=min_match, anchor=anchor)
    assert list(zip(v.tolist(), s.tolist(), d.tolist())) == want
    assert len(want) > 1 or min_match == 5 or anchor > 1
    if anchor == 0:      # the anchored modes can only lose candidates: their hits are a subset with no higher score
        for a in (1, 2, 3):
            sub = dict((v2, s2) for v2, s2, _ in oracle.find_fragments_csr(ts, off, vid, q, min_match=min_match, anchor=a))
            full = dict((v2, s2) for v2, s2, _ in want)
            assert all(v2 in full and full[v2] >= s2 for v2, s2 in sub.items())
    assert (int(vid[77]), len(q)) in list(zip(v.tolist(), s.tolist()))
    cat.close()


@pytest.mark.parametrize("anchor", [2, 3])
def test_streaming_kernel_ragged_rows_and_long_queries(cuda, anchor):
    """The streaming kernel reads the shard as one flat array: rows of every length (empty, shorter
    than an anchor, straddling warp / CTA chunk boundaries), intervals beyond the bucket range, dense
    true matches (every row is a shifted copy) and queries with more than 32 intervals (the
    bucket sets are indexed mod 32)."""
    rng = np.random.default_rng(40 + anchor)
    base = np.cumsum(rng.integers(15, 150, 6000)) / 30.0
    base[3000:] += 40.0                                            # one interval of > 32767 ticks
    rows = []
    for i in range(900):
        n = int(rng.choice([0, 1, 2, 3, 4, 7, 60, 300, 1100]))
        a = int(rng.integers(0, 6000 - n))
        shift = float(rng.integers(0, 100)) if i % 3 else 0.0
        rows.append((i + 1, (np.round((base[a:a + n] + shift) * 1000) / 1000).tolist()))
    from tvidz_b200.catalog import rows_to_csr
    ts, off, vid = rows_to_csr(rows)
    assert off[-1] > 6 * 8192                                     # several CTA chunks
    cat = FragmentCatalogue(ts, off, vid, hit_capacity=64)
    for a, n, mm in ((100, 8, 3), (2990, 20, 5), (500, 50, 4), (1234, 90, 0)):
        q = (np.round(base[a:a + n] * 1000) / 1000).tolist()
        v, s, d = cat.match(q, mm, anchor=anchor)
        want = oracle.find_fragments_csr(ts, off, vid, q, min_match=mm, anchor=anchor)
        assert list(zip(v.tolist(), s.tolist(), d.tolist())) == want, (a, n, mm)
        assert len(want) > 3
    cat.close()


def test_survivor_queue_overflow_resolves_in_place(cuda, monkeypatch):
    """A query that matches nearly everywhere fills a warp's survivor queue before the end of its
    stream (forced tiny here): the early drains must give identical results."""
    monkeypatch.setenv("TVZ_FRAG_QUEUE_CAP", "40")
    rng = np.random.default_rng(9)
    base = np.cumsum(rng.integers(15, 150, 3000)) / 30.0
    rows = [(i + 1, (np.round((base[a:a + 400] + i) * 1000) / 1000).tolist())
            for i, a in enumerate(rng.integers(0, 2600, 300))]
    from tvidz_b200.catalog import rows_to_csr
    ts, off, vid = rows_to_csr(rows)
    cat = FragmentCatalogue(ts, off, vid, hit_capacity=512)
    q = (np.round(base[1000:1030] * 1000) / 1000).tolist()
    for _ in range(2):                                              # the queue is reset between queries
        v, s, d = cat.match(q, 4)
        want = oracle.find_fragments_csr(ts, off, vid, q, min_match=4)
        assert list(zip(v.tolist(), s.tolist(), d.tolist())) == want
    assert len(want) > 40
    cat.close()


def test_short_rows_tolerances_and_edges(cuda):
    rows = [(1, [0.5, 1.0, 2.5, 4.0]), (2, [100.5, 101.0, 102.5, 104.0, 250.0]), (3, []), (4, [7.0]),
            (5, [10.0, 10.5]), (6, [4.0, 2.5, 1.0, 0.5, 0.5, float("nan")]),        # unsorted, repeat, NaN
            (7, [1000.503, 1001.004, 1002.498, 1003.999])]                            # 3-4 ms off
    from tvidz_b200.catalog import rows_to_csr
    ts, off, vid = rows_to_csr(rows)
    cat = FragmentCatalogue(ts, off, vid)
    q = [0.5, 1.0, 2.5, 4.0]
    for tol, tol_gap in ((0, 0), (2, 4), (7, 14)):
        for mm in (0, 1, 2, 4):
            for anchor in (0, 1, 2, 3):
                v, s, d = cat.match(q, mm, tol=tol, tol_gap=tol_gap, anchor=anchor)
                assert list(zip(v.tolist(), s.tolist(), d.tolist())) == \
                    oracle.find_fragments_csr(ts, off, vid, q, min_match=mm, tol=tol, tol_gap=tol_gap,
                                              anchor=anchor), (tol, mm, anchor)
    res = dict((v, (s, o)) for v, s, o in cat.find_fragments(q, 4))
    assert res[1] == (4, 0.0) and res[2] == (4, 100.0) and res[6] == (4, 0.0) and res[7][0] == 4
    assert 7 not in dict((v, s) for v, s, _ in cat.find_fragments(q, 4, tol=0, tol_gap=0))
    assert cat.find_fragments([], 1) == [] and cat.find_fragments([3.0], 1) == []      # no interval, no anchor
    assert cat.find_fragments(q, 4, top_k=2)[0][1] == 4
    cat.close()


def test_zero_offset_collapses_to_find_duplicates(cuda):
    """offset 0, tolerance 0 == match_count of db.py:85-89 on tick-exact data."""
    rng = np.random.default_rng(3)
    rows = [(i + 1, np.unique(rng.integers(0, 4000, int(rng.integers(0, 60))) / 8.0).tolist()) for i in range(800)]
    from tvidz_b200.catalog import rows_to_csr
    ts, off, vid = rows_to_csr(rows)
    frag = FragmentCatalogue(ts, off, vid)
    exact = Catalogue(ts, off, vid)
    for _ in range(5):
        q = np.unique(rng.integers(0, 4000, 40) / 8.0)
        for mm in (1, 2, 4):
            v, s, d = frag.match(q, mm, tol=0, tol_gap=0, zero_offset_only=True)
            assert list(zip(v.tolist(), s.tolist())) == exact.find_duplicates(q, mm) == \
                oracle.find_duplicates_csr(ts, off, vid, q, mm)
            assert not d.any()
    frag.close()
    exact.close()


def test_long_rows_read_in_place(cuda):
    """Rows longer than the 2048-tick shared-memory buffer take the in-place path."""
    ts, off, vid = synth.synth_catalogue(12, len_range=(2500, 5000), gap_range=(15, 150), seed=2)
    cat = FragmentCatalogue(ts, off, vid)
    row = ts[off[5]:off[6]]
    q = clip_query(row, 30_000, n_frames=1800)
    for anchor in (1, 2):
        v, s, d = cat.match(q, 3, anchor=anchor)
        assert list(zip(v.tolist(), s.tolist(), d.tolist())) == \
            oracle.find_fragments_csr(ts, off, vid, q, min_match=3, anchor=anchor)
        assert (int(vid[5]), len(q)) in list(zip(v.tolist(), s.tolist()))
    cat.close()


def test_full_size_config5(cuda):
    """BASELINE config 5 geometry: a 30 s clip against 100k long videos.  Oracle on a 3000-row
    slice around the source row; size-independent property on the whole catalogue (the source row
    is reported with every cut and the right offset)."""
    ts, off, vid = _long_catalogue(100_000, seed=0)
    cat = FragmentCatalogue(ts, off, vid)
    r = 54_321
    row = ts[off[r]:off[r + 1]]
    f0 = 40_000
    q = clip_query(row, f0)
    assert len(q) >= 6
    for anchor in (2, 1):
        got = cat.find_fragments(q, min_match=5, anchor=anchor)
        top = cat.find_fragments(q, min_match=5, top_k=1, anchor=anchor)[0]
        assert top[0] == int(vid[r]) and top[1] == len(q) and abs(top[2] - f0 / 30.0) <= 0.008
        lo, hi = r - 1500, r + 1500
        sub = oracle.find_fragments_csr(ts[off[lo]:off[hi]], off[lo:hi + 1] - off[lo], vid[lo:hi], q, min_match=5,
                                        anchor=anchor)
        ids = set(vid[lo:hi].tolist())
        assert [(v, s, round(o * 1000)) for v, s, o in got if v in ids] == sub
    cat.close()

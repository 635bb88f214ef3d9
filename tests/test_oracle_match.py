"""The CPU oracle for stage 2 against golden vectors produced by executing the
reference's own find_duplicates / analyze_file (tests/golden/gen_golden.py)."""
import numpy as np

import oracle
from oracle import match_oracle
from tvidz_b200.catalog import rows_to_csr


def _catalogue(case):
    # json.load maps every "NaN" to one shared object; a DB fetch never does (see hydrate)
    return match_oracle.hydrate((vid, ts) for vid, (_, ts) in zip(case["video_ids"], case["catalogue"]))


def _mm(case):
    return 5 if case["min_match"] is None else case["min_match"]


def test_reference_vectors_are_present(match_golden):
    names = {c["name"] for c in match_golden}
    assert {"test_app.py:71-78", "test_app.py:80-84", "app.py:399-408"} <= names
    c = next(c for c in match_golden if c["name"] == "test_app.py:71-78")
    assert c["expected"] == [[c["video_ids"][1], 5]]
    c = next(c for c in match_golden if c["name"] == "test_app.py:80-84")
    assert c["expected"] == [[c["video_ids"][0], 5], [c["video_ids"][2], 5]]


def test_python_oracle_equals_reference(match_golden):
    for c in match_golden:
        got = match_oracle.find_duplicates(_catalogue(c), c["query"], _mm(c))
        assert [list(x) for x in got] == c["expected"], c["name"]


def test_c_oracle_equals_reference(match_golden):
    for c in match_golden:
        ts, off, vid = rows_to_csr(_catalogue(c))
        got = oracle.find_duplicates_csr(ts, off, vid, np.asarray(c["query"], np.float64), _mm(c))
        assert [list(x) for x in got] == c["expected"], c["name"]


def test_streaming_oracle_equals_reference(stream_golden):
    for c in stream_golden:
        cat = _catalogue(c)
        names = {vid: fn for vid, (fn, _) in zip(c["video_ids"], c["catalogue"])}
        scene, ids, dup_names = match_oracle.streaming_analysis(cat, c["self_video_id"], c["tokens"], 2, names)
        res = c["result"]
        assert res["status"] == "done", c["name"]
        assert scene == res["scene_cuts"] and len(scene) == res["total_cuts"], c["name"]
        assert set(dup_names) == set(res["duplicates"]), c["name"]
        assert ids == c["stored_duplicates"], c["name"]
        rec = match_oracle.result_record(scene, dup_names, res["original_filename"], res["clean_filename"])
        assert {k: v for k, v in rec.items() if k != "duplicates"} == {k: v for k, v in res.items() if k != "duplicates"}


def test_kth_is_the_streaming_stop(stream_golden):
    """the loop stops at min over other rows of K_r."""
    for c in stream_golden:
        cat = [(v, t) for v, t in _catalogue(c) if v != c["self_video_id"]]
        cuts = []
        for tok in c["tokens"]:
            ts = float(tok)
            if not cuts or ts != cuts[-1]:
                cuts.append(ts)
        if not cat:
            continue
        ts, off, vid = rows_to_csr(cat)
        kth = oracle.match_kth(ts, off, np.asarray(cuts), 2)
        hit = kth[kth > 0]
        stop = int(hit.min()) if hit.size else len(cuts)
        assert cuts[:stop] == c["result"]["scene_cuts"], c["name"]
        assert sorted(int(v) for v in vid[kth == stop]) == sorted(c["stored_duplicates"]) or not hit.size


def test_c_oracle_equals_python_oracle_random():
    rng = np.random.default_rng(9)
    rows = []
    for i in range(300):
        L = int(rng.integers(0, 50))
        rows.append((i + 1, (rng.integers(0, 400, L) / 4.0).tolist()))
    ts, off, vid = rows_to_csr(rows)
    for _ in range(20):
        q = (rng.integers(0, 400, int(rng.integers(0, 30))) / 4.0).tolist()
        mm = int(rng.integers(-1, 6))
        assert oracle.find_duplicates_csr(ts, off, vid, q, mm) == match_oracle.find_duplicates(rows, q, mm)

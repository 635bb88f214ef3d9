"""Generate tests/golden/match_golden.json and stream_golden.json by executing
the reference's own code (see ref_harness.py).  Run in the build container:

    python tests/golden/gen_golden.py

Seeds are fixed; the fixtures are committed so the GPU box (which has no
/root/reference) checks the oracle and the CUDA path against them.
"""
from __future__ import annotations

import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_harness import ref_analyze_file, ref_find_duplicates  # noqa: E402


def g6(n, fps=30):
    """FFmpeg <=6 showinfo pts_time text for frame n at time_base 1/fps
    (libavutil/timestamp.h: "%.6g" of av_q2d(tb) * pts)."""
    return "%.6g" % ((1.0 / fps) * n)


def synth_row(rng, lo, hi, gap=(15, 600)):
    n, L, out = 0, rng.randint(lo, hi), []
    for _ in range(L):
        n += rng.randint(*gap)
        out.append(float(g6(n)))
    return out


def match_cases():
    cases = []

    def add(name, catalogue, query, min_match):
        ids, res = ref_find_duplicates(catalogue, query, min_match)
        cases.append({"name": name, "catalogue": [[f, list(t)] for f, t in catalogue],
                      "video_ids": ids, "query": list(query), "min_match": min_match,
                      "expected": [list(x) for x in res]})

    # --- the reference's own vectors -------------------
    a = [1.0, 2.0, 3.0, 4.0, 5.0]
    b = [10.0, 20.0, 30.0, 40.0, 50.0]
    add("test_app.py:71-78", [("a.mp4", a), ("b.mp4", b)], b, 5)
    add("test_app.py:80-84", [("a.mp4", a), ("b.mp4", b), ("c.mp4", a)], a, 5)
    add("app.py:399-408", [("test.mp4", [1.2, 5.7, 12.3, 18.9])], [1.2, 5.7, 12.3, 18.9], 2)
    guide = [("sample1.mp4", [1.2, 5.7, 12.3, 18.9, 25.1]), ("sample2.mp4", [2.1, 8.4, 15.7, 22.1, 28.9]),
             ("duplicate.mp4", [1.2, 5.7, 12.3])]
    add("guide.md:1269-1272 q=row3", guide, [1.2, 5.7, 12.3], 2)
    add("guide.md:1269-1272 q=row1", guide, [1.2, 5.7, 12.3, 18.9, 25.1], 3)
    add("default min_match", [("a.mp4", a), ("b.mp4", a[:4])], a, None)

    # --- edge semantics --------------------------------
    nan, inf = float("nan"), float("inf")
    edge = [("dupes_in_row", [1.0, 1.0, 2.0, 2.0, 2.0]), ("zeros", [0.0, 3.0]), ("negzero", [-0.0, 3.0]),
            ("nan_row", [nan, 1.0, nan]), ("inf_row", [inf, -inf, 1.0]), ("empty", []), ("single", [7.5]),
            ("unsorted", [9.0, 1.0, 5.0, 3.0])]
    add("query multiplicity", edge, [1.0, 1.0, 1.0, 2.0], 1)
    add("signed zero", edge, [0.0, -0.0, 3.0], 1)
    add("nan never matches", edge, [nan, 1.0, nan], 1)
    add("infinities", edge, [inf, -inf], 1)
    add("empty query min0", edge, [], 0)
    add("empty query min1", edge, [], 1)
    add("min_match negative", edge, [1.0], -3)
    add("min_match zero", edge, [1.0, 9.0], 0)
    add("ints in query", edge, [1, 2, 3, 5, 9], 2)
    add("unsorted query", edge, [5.0, 9.0, 3.0, 1.0, 5.0], 2)
    add("denormal and huge", [("x", [5e-324, 1.7976931348623157e308, 2.2250738585072014e-308])],
        [5e-324, 0.0, 1.7976931348623157e308], 1)

    # --- random frame-quantised catalogues (chance collisions included) -------
    rng = random.Random(20261018)
    for k in range(12):
        n_rows = rng.choice([1, 7, 33, 150])
        cat = [("v%d.mp4" % i, synth_row(rng, 0, 40, gap=(1, 40))) for i in range(n_rows)]
        src = rng.randrange(n_rows)
        mode = k % 4
        if mode == 0:
            q = list(cat[src][1])                       # full duplicate
        elif mode == 1:
            q = list(cat[src][1])[: rng.randint(0, 6)]  # streaming prefix
        elif mode == 2:
            q = synth_row(rng, 0, 40, gap=(1, 40))      # unrelated video
        else:
            q = list(cat[src][1]) + list(cat[src][1])[:3] + [123456.0]  # repeats + a miss
        add("random-%d" % k, cat, q, rng.choice([1, 2, 2, 5]))
    return cases


def showinfo_line(n_sel, frame, fps=30, addr="0x5581f2c3a540"):
    """A showinfo line shaped like FFmpeg 5/6 prints it (vf_showinfo.c):
    n:%4d pts:%7s pts_time:%-7s ...  (pts == frame index at time_base 1/fps)."""
    return ("[Parsed_showinfo_1 @ %s] n:%4d pts:%7s pts_time:%-7s duration:%7s duration_time:%-7s "
            "fmt:yuv420p cl:left sar:1/1 s:1920x1080 i:P iskey:%d type:%c checksum:%08X "
            "plane_checksum:[%08X %08X %08X] mean:[%d %d %d] stdev:[%.1f %.1f %.1f]"
            % (addr, n_sel, frame, g6(frame, fps), 1, g6(1, fps), n_sel % 2, "IPB"[n_sel % 3],
               0xDEADBEEF ^ frame, frame * 2654435761 % 2**32, 17, 23, 120, 128, 128, 40.5, 3.2, 2.9))


def stream_cases():
    cases = []
    rng = random.Random(77)
    noise = ["Input #0, mov,mp4,m4a,3gp,3g2,mj2, from '/tmp/x.mp4':",
             "  Stream #0:0(und): Video: h264 (High), yuv420p, 1920x1080, 30 fps, 30 tbr, 15360 tbn",
             "[Parsed_showinfo_1 @ 0x5581f2c3a540] config in time_base: 1/30, frame_rate: 30/1",
             "[Parsed_showinfo_1 @ 0x5581f2c3a540] config out time_base: 0/0, frame_rate: 0/0",
             "frame=  512 fps=0.0 q=-0.0 size=N/A time=00:00:17.06 bitrate=N/A speed=34.1x"]

    def add(name, catalogue, upload, frames, repeat_at=()):
        lines = list(noise[:4])
        for i, fr in enumerate(frames):
            lines.append(showinfo_line(i, fr))
            if i in repeat_at:                       # same pts_time twice in a row (app.py:231)
                lines.append(showinfo_line(i, fr))
            if i % 3 == 2:
                lines.append(noise[4])
        res, me, stored, ids, terminated = ref_analyze_file(catalogue, upload, lines)
        cases.append({"name": name, "catalogue": [[f, list(t)] for f, t in catalogue], "video_ids": ids,
                      "upload": upload, "self_video_id": me, "stderr": lines,
                      "tokens": [g6(fr) for fr in frames], "result": res, "stored_duplicates": stored,
                      "terminated": terminated})

    def cuts(n):
        f, out = 0, []
        for _ in range(n):
            f += rng.randint(15, 600)
            out.append(f)
        return out

    base = cuts(14)
    other = cuts(9)
    cat = [("orig.mp4", [float(g6(f)) for f in base]), ("other.mp4", [float(g6(f)) for f in other])]
    add("exact re-upload stops at cut 2", cat, "1760000000000-orig.mp4", base)
    add("no duplicate runs to the end", cat, "1760000000001-fresh.mp4", cuts(11))
    shifted = [base[0] + 1] + base[1:5] + [base[5] + 2] + base[6:]
    add("first cut differs: stops at cut 3", cat, "1760000000002-almost.mp4", shifted)
    add("repeated pts_time line is dropped", cat, "clip.mp4", base[:6], repeat_at=(0, 3))
    two = cat + [("copy-of-orig.mp4", [float(g6(f)) for f in base[:8]])]
    add("two stored copies both reported", two, "uploads/1760000000003-orig.mp4", base)
    add("single cut never matches", cat, "one.mp4", base[:1])
    add("empty catalogue", [], "first.mp4", base[:5])
    mixed = [("a.mp4", [float(g6(f)) for f in (base[0], other[1], base[3])]),
             ("b.mp4", [float(g6(f)) for f in (base[2], base[3])])]
    add("hits arrive late and tie", mixed, "1760000000004-m.mp4", base[:6])
    add("repeated lines without a duplicate", cat, "r.mp4", cuts(7), repeat_at=(1, 2, 6))
    add("long timestamps", [("long.mp4", [float(g6(f)) for f in (215999, 300000, 300037)])],
        "l.mp4", [3037, 215999, 300000, 300037])
    return cases


if __name__ == "__main__":
    m = match_cases()
    with open(os.path.join(HERE, "match_golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/gen_golden.py", "source": "inspector/db.py:76-94 executed",
                   "cases": m}, f, indent=1)
    s = stream_cases()
    with open(os.path.join(HERE, "stream_golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/gen_golden.py",
                   "source": "inspector/app.py:117-322 executed with stubbed I/O", "cases": s}, f, indent=1)
    print("match cases:", len(m), " stream cases:", len(s))
    for c in s:
        print(" ", c["name"], "->", c["result"]["status"], c["result"].get("total_cuts"),
              c["result"].get("duplicates"), "terminated" if c["terminated"] else "")

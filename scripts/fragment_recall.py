"""Recall of the interval-anchored candidate rules (anchor = 1, 2, 3) against the exhaustive SURVEY.md B.4
mode (anchor = 0: every offset C[j] - Q[i]) on a configs[4]-shaped catalogue, for clean 30 s clips and for
clips that lost 1-3 of their cuts (a missed detection removes a cut and merges two intervals) or whose cuts
jitter by a few ms.  Also times the four modes.   python scripts/fragment_recall.py [rows] [clips]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tvidz_b200 import synth
from tvidz_b200.fragment import FragmentCatalogue, clip_query

n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
n_clips = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ts, off, vid = synth.synth_catalogue(n_rows, len_range=(600, 1400), gap_range=(15, 150), seed=1)
cat = FragmentCatalogue(ts, off, vid, hit_capacity=1 << 14)
rng = np.random.default_rng(11)
mm = 5
print(f"catalogue: {n_rows} rows, {cat.n_values} ticks; {n_clips} clips per condition; min_match = {mm}")
print(f"{'condition':<28} {'clip cuts':>9} | " + " | ".join(f"anchor {a}: true-row recall, hits vs exhaustive" for a in (1, 2, 3)))
conds = [("clean", 0, 0), ("1 cut dropped", 1, 0), ("2 cuts dropped", 2, 0), ("3 cuts dropped", 3, 0),
         ("jitter +-5 ms", 0, 5), ("2 dropped + jitter +-5 ms", 2, 5)]
for name, drop, jitter in conds:
    found = {a: 0 for a in (0, 1, 2, 3)}
    hits = {a: 0 for a in (0, 1, 2, 3)}
    agree = {a: 0 for a in (1, 2, 3)}
    usable, lens = 0, []
    for _ in range(n_clips):
        r = int(rng.integers(n_rows))
        row = ts[off[r]:off[r + 1]]
        f0 = int(rng.integers(0, max(1, int(row[-1] * 30) - 900)))
        q = np.asarray(clip_query(row, f0))
        if q.shape[0] - drop < mm + 1:
            continue
        if drop:
            keep = np.sort(rng.choice(q.shape[0], q.shape[0] - drop, replace=False))
            q = q[keep]
        if jitter:
            q = q + rng.integers(-jitter, jitter + 1, q.shape[0]) / 1000.0
        usable += 1
        lens.append(q.shape[0])
        res = {}
        for a in (0, 1, 2, 3):
            v, s, d = cat.match(q, mm, anchor=a)
            res[a] = dict(zip(v.tolist(), zip(s.tolist(), d.tolist())))
            hits[a] += len(res[a])
            found[a] += int(vid[r]) in res[a]
        for a in (1, 2, 3):
            agree[a] += sum(1 for k, val in res[a].items() if res[0].get(k) == val)
    cols = []
    for a in (1, 2, 3):
        cols.append(f"{found[a] / max(1, found[0]):6.3f} ({found[a]}/{found[0]}), {hits[a]}/{hits[0]} hits, {agree[a]} identical")
    print(f"{name:<28} {np.mean(lens):9.1f} | " + " | ".join(cols) + f"   [exhaustive finds the true row in {found[0]}/{usable}]")
# timing on one clean clip
q = np.asarray(clip_query(ts[off[54_321 % n_rows]:off[54_321 % n_rows + 1]], 40_000))
rec = torch.zeros(3 * ((1 << 14) + 1), dtype=torch.int32, device="cuda")
for a in (2, 3, 1, 0):
    for _ in range(2):
        cat.match_async(q, mm, rec, anchor=a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 20 if a >= 2 else 3
    e0.record()
    for _ in range(k):
        cat.match_async(q, mm, rec, anchor=a)
    e1.record()
    torch.cuda.synchronize()
    print(f"anchor {a}: {e0.elapsed_time(e1) / k * 1e3:9.1f} us per query ({len(q)} cuts, {n_rows} rows)")
cat.close()

"""Profiling target for fragment mode (configs[4] geometry, fewer rows by default)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tvidz_b200 import synth
from tvidz_b200.fragment import FragmentCatalogue, clip_query

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ts, off, vid = synth.synth_catalogue(n, len_range=(600, 1400), gap_range=(15, 150), seed=1)
cat = FragmentCatalogue(ts, off, vid, hit_capacity=1 << 12)
r = n // 2
q = clip_query(ts[off[r]:off[r + 1]], 40_000)
rec = torch.zeros(3 * ((1 << 12) + 1), dtype=torch.int32, device="cuda:0")
for _ in range(reps):
    cat.match_async(q, 5, rec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    cat.match_async(q, 5, rec)
e1.record()
torch.cuda.synchronize()
print("hits", int(rec[0]), "cuts", len(q), "ms/query", e0.elapsed_time(e1) / reps)

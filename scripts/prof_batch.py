import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tvidz_b200 import synth
from tvidz_b200.catalog import Catalogue
ts, off, vid = synth.synth_catalogue(1_000_000, seed=0)
cat = Catalogue(ts, off, vid, hit_capacity=1 << 16)
rng = np.random.default_rng(7)
qs = [ts[off[r]:off[r + 1]].copy() for r in rng.integers(0, 1_000_000, 8)]
for _ in range(2):
    out = cat.match_many(qs, 2)
torch.cuda.synchronize()
print(sum(len(a) for a, b in out))

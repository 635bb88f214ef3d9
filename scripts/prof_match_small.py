"""ncu target: the matcher on an N=8-sized shard (125k rows) -- a few real queries, then a few empty ones
(the kernel's fixed cost alone).  python scripts/prof_match_small.py [rows]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tvidz_b200 import synth
from tvidz_b200.catalog import Catalogue

n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
ts, off, vid = synth.synth_catalogue(n, seed=0)
cat = Catalogue(ts, off, vid, hit_capacity=1 << 15)
rec = torch.zeros(((1 << 15) + 1, 2), dtype=torch.int32, device="cuda")
q = ts[off[12_345]:off[12_346]].copy()
for _ in range(6):
    cat.match_async(q, 2, rec)
torch.cuda.synchronize()
print("hits", int(rec[0, 0]))
for _ in range(4):
    cat.match_async(np.zeros(0), 2, rec)
torch.cuda.synchronize()
cat.close()

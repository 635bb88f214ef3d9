"""Small profiling target: the bench's two workloads without the CPU arms or e2e loops.
Used under ncu (after a plain run of the same command has exited 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tvidz_b200 import scene, synth
from tvidz_b200.catalog import Catalogue

what = sys.argv[1] if len(sys.argv) > 1 else "both"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
if what in ("score", "both"):
    frames = synth.synth_frames(64, 32, 1080, 1920, seed=100, scene_len=(8, 20), device=dev)
    for _ in range(reps):
        sad, score, sel = scene.score_frames(frames)
    torch.cuda.synchronize()
    print("cuts", int(sel.sum()))
if what in ("match", "both"):
    ts, off, vid = synth.synth_catalogue(1_000_000, seed=0)
    cat = Catalogue(ts, off, vid, hit_capacity=1 << 16)
    q = ts[off[123456]:off[123457]].copy()
    rec = torch.zeros(((1 << 16) + 1, 2), dtype=torch.int32, device=dev)
    for _ in range(reps):
        cat.match_async(q, 2, rec)
    torch.cuda.synchronize()
    print("hits", int(rec[0, 0]))
if what == "batch":
    import numpy as np
    ts, off, vid = synth.synth_catalogue(1_000_000, seed=0)
    cat = Catalogue(ts, off, vid, hit_capacity=1 << 15)
    qs = [ts[off[r]:off[r + 1]].copy() for r in np.random.default_rng(7).integers(0, 1_000_000, 8)]
    rec8 = torch.zeros((8, (1 << 15) + 1, 2), dtype=torch.int32, device=dev)
    for _ in range(reps):
        cat.match_batch_async(qs, 2, rec8)
    torch.cuda.synchronize()
    print("hits", rec8[:, 0, 0].tolist())

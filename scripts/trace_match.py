"""Where one matcher query spends its time inside the kernel: per-CTA phase timestamps (debug hook
tvz_debug_tile_trace) for a cold-L2 query at the given shard sizes.
  python scripts/trace_match.py [rows ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
from tvidz_b200 import synth
from tvidz_b200._lib import check, lib
from tvidz_b200.catalog import Catalogue

sizes = [int(x) for x in sys.argv[1:]] or [1_000_000, 125_000]
dev = torch.device("cuda:0")
ts, off, vid = synth.synth_catalogue(1_000_000, seed=0)
q = ts[off[123_456]:off[123_457]].copy()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sink = torch.zeros(1, dtype=torch.int64, device=dev)
mhz = 1965.0
names = ["zero map+counts (+td)", "keys -> map", "pdl wait + seq", "stream (warp 0)", "drain + sync", "count rows + publish", "poll predecessors", "excl + header", "write hits (thread 0)"]
for n in sizes:
    cat = Catalogue(ts[:off[n]], off[:n + 1], vid[:n], hit_capacity=1 << 15)
    rec = torch.zeros(((1 << 15) + 1, 2), dtype=torch.int32, device=dev)
    for _ in range(3):
        cat.match_async(q, 2, rec)
    trace = torch.zeros((cat.n_tiles, 16), dtype=torch.int64, device=dev)
    ws = cat._ws_async(0)
    for mode in ("cold", "warm"):
        check(lib().tvz_debug_tile_trace(ws.handle, trace.data_ptr()))
        if mode == "cold":
            flush.zero_()
            sink += flush.view(torch.int64).sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cat.match_async(q, 2, rec)
        e1.record()
        torch.cuda.synchronize()
        check(lib().tvz_debug_tile_trace(ws.handle, None))
        t = trace.cpu().numpy().astype(np.float64)
        start_ns = t[:, 0] - t[:, 0].min()
        d = np.diff(t[:, 1:11], axis=1) / mhz * 1.0          # cycles -> us at `mhz` MHz
        total = (t[:, 10] - t[:, 1]) / mhz
        print(f"rows {n} {mode}: event time {e0.elapsed_time(e1) * 1e3:.1f} us; CTA start spread {start_ns.max() / 1e3:.1f} us "
              f"(median {np.median(start_ns) / 1e3:.1f}); CTA lifetime median {np.median(total):.1f} max {total.max():.1f} us "
              f"(clock {mhz} MHz assumed)")
        for k, nm in enumerate(names):
            col = d[:, k]
            print(f"    {nm:<24} median {np.median(col):6.2f}  p90 {np.percentile(col, 90):6.2f}  max {col.max():6.2f} us")
    cat.close()

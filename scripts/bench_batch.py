"""The 8-query pass on ONE GPU at the shard sizes of 1- and 8-GPU runs (no gather): device time per pass back to back
and with a cold L2, the kernel alone (library events), and what the host spends enqueueing a pass when nothing
throttles it.  TVZ_LIB selects a build variant (scripts/batch_variants.sh).
  python scripts/bench_batch.py [rows ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from tvidz_b200 import synth
from tvidz_b200.catalog import Catalogue

sizes = [int(x) for x in sys.argv[1:]] or [1_000_000, 125_000]
dev = torch.device("cuda:0")
ts, off, vid = synth.synth_catalogue(1_000_000, seed=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sink = torch.zeros(1, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream()
CAP = 1 << 15
for n in sizes:
    cat = Catalogue(ts[:off[n]], off[:n + 1], vid[:n], hit_capacity=CAP)
    rec8 = torch.zeros((8, CAP + 1, 2), dtype=torch.int32, device=dev)
    qs = [ts[off[r]:off[r + 1]].copy() for r in np.random.default_rng(7).integers(0, n, 8)]
    out = {}
    for mm in (2, 5):
        for _ in range(5):
            cat.match_batch_async(qs, mm, rec8)
        torch.cuda.synchronize()
        host = rec8.cpu().numpy()
        ok = True
        if n <= 125_000:                                              # the oracle finishes in seconds at this size
            for b, q in enumerate(qs):
                want = oracle.find_duplicates_csr(ts[:off[n]], off[:n + 1], vid[:n], q, mm)
                k = int(host[b, 0, 0])
                ok = ok and k == len(want) and [tuple(x) for x in host[b, 1:1 + k].tolist()] == want
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(100):
            cat.match_batch_async(qs, mm, rec8)
        e1.record(stream)
        torch.cuda.synchronize()
        b2b = e0.elapsed_time(e1) / 100 * 1e3
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a, b in evs:
            flush.zero_()
            sink += flush.view(torch.int64).sum()
            a.record(stream)
            cat.match_batch_async(qs, mm, rec8)
            b.record(stream)
        torch.cuda.synchronize()
        cold = float(np.mean([a.elapsed_time(b) for a, b in evs])) * 1e3
        cat.debug_count_kernel_ms(True)
        ks = []
        for _ in range(10):
            flush.zero_()
            sink += flush.view(torch.int64).sum()
            cat.match_batch_async(qs, mm, rec8)
            ks.append(cat.debug_count_kernel_ms() * 1e3)
        cat.debug_count_kernel_ms(False)
        hosts = []
        for _ in range(20):                                           # bursts of 3 on an idle stream: no back-pressure
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                cat.match_batch_async(qs, mm, rec8)
            hosts.append((time.perf_counter() - t0) / 3 * 1e6)
        torch.cuda.synchronize()
        out[mm] = (b2b, cold, float(np.mean(ks)), float(np.median(hosts)), int(host[:, 0, 0].sum()), ok)
    long_qs = [np.arange(200, dtype=np.float64) * 1.5 + 0.25 * b for b in range(8)]   # > 96 distinct values: the large parameter block
    for _ in range(5):
        cat.match_batch_async(long_qs, 2, rec8)
    e0.record(stream)
    for _ in range(100):
        cat.match_batch_async(long_qs, 2, rec8)
    e1.record(stream)
    torch.cuda.synchronize()
    long_b2b = e0.elapsed_time(e1) / 100 * 1e3
    print(f"rows {n:>8} tiles {cat.n_tiles:>4}: " + " | ".join(
        f"mm{mm}: pass b2b {v[0]:6.1f} us cold {v[1]:6.1f} us kernel(cold) {v[2]:6.1f} us host enqueue {v[3]:5.1f} us hits {v[4]} parity {v[5]}"
        for mm, v in out.items()) + f" | 200-value queries b2b {long_b2b:6.1f} us", flush=True)
    cat.close()

"""Matcher timings on ONE GPU at the shard sizes of 1/2/4/8-GPU runs (no gather): whole query back to back
and with a cold L2, the kernel alone (library events), and what the host spends enqueueing one query.
  python scripts/bench_match.py [rows ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tvidz_b200 import synth
from tvidz_b200.catalog import Catalogue

sizes = [int(x) for x in sys.argv[1:]] or [1_000_000, 500_000, 250_000, 125_000]
dev = torch.device("cuda:0")
ts, off, vid = synth.synth_catalogue(1_000_000, seed=0)
r_star = 123_456
q = ts[off[r_star]:off[r_star + 1]].copy()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sink = torch.zeros(1, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream()
for n in sizes:
    cat = Catalogue(ts[:off[n]], off[:n + 1], vid[:n], hit_capacity=1 << 15)
    rec = torch.zeros(((1 << 15) + 1, 2), dtype=torch.int32, device=dev)
    rec8 = torch.zeros((8, (1 << 15) + 1, 2), dtype=torch.int32, device=dev)
    for _ in range(5):
        cat.match_async(q, 2, rec)
    torch.cuda.synchronize()
    K = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(K):
        cat.match_async(q, 2, rec)
    host_us = (time.perf_counter() - t0) / K * 1e6
    e1.record(stream)
    torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) / K * 1e3
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in evs:
        flush.zero_()
        sink += flush.view(torch.int64).sum()
        a.record(stream)
        cat.match_async(q, 2, rec)
        b.record(stream)
    torch.cuda.synchronize()
    cold = float(np.mean([a.elapsed_time(b) for a, b in evs])) * 1e3
    # kernel alone, device time: events recorded while the stream is still busy with the flush
    cat.debug_count_kernel_ms(True)
    ks = []
    for _ in range(10):
        flush.zero_()
        sink += flush.view(torch.int64).sum()
        cat.match_async(q, 2, rec)
        ks.append(cat.debug_count_kernel_ms() * 1e3)
    cat.debug_count_kernel_ms(False)
    # a selective query (the reference's default min_match = 5): most tiles have no hit and skip the exchange
    for _ in range(3):
        cat.match_async(q, 5, rec)
    e0.record(stream)
    for _ in range(K):
        cat.match_async(q, 5, rec)
    e1.record(stream)
    torch.cuda.synchronize()
    mm5_b2b = e0.elapsed_time(e1) / K * 1e3
    for a, b in evs:
        flush.zero_()
        sink += flush.view(torch.int64).sum()
        a.record(stream)
        cat.match_async(q, 5, rec)
        b.record(stream)
    torch.cuda.synchronize()
    mm5_cold = float(np.mean([a.elapsed_time(b) for a, b in evs])) * 1e3
    # the fixed cost alone: an empty query runs the same kernel without streaming anything
    qe = np.zeros(0)
    for _ in range(3):
        cat.match_async(qe, 2, rec)
    e0.record(stream)
    for _ in range(K):
        cat.match_async(qe, 2, rec)
    e1.record(stream)
    torch.cuda.synchronize()
    floor_b2b = e0.elapsed_time(e1) / K * 1e3
    for a, b in evs:
        flush.zero_()
        sink += flush.view(torch.int64).sum()
        a.record(stream)
        cat.match_async(qe, 2, rec)
        b.record(stream)
    torch.cuda.synchronize()
    floor_cold = float(np.mean([a.elapsed_time(b) for a, b in evs])) * 1e3
    # 8 queries per pass
    qs = [ts[off[r]:off[r + 1]].copy() for r in np.random.default_rng(7).integers(0, n, 8)]
    for _ in range(3):
        cat.match_batch_async(qs, 2, rec8)
    torch.cuda.synchronize()
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(50):
        cat.match_batch_async(qs, 2, rec8)
    host8_us = (time.perf_counter() - t0) / 50 * 1e6
    e1.record(stream)
    torch.cuda.synchronize()
    b8 = e0.elapsed_time(e1) / 50 * 1e3
    t0 = time.perf_counter()
    for _ in range(50):
        full = cat.find_duplicates(q, 2)
    e2e = (time.perf_counter() - t0) / 50 * 1e6
    t0 = time.perf_counter()
    for _ in range(50):
        full5 = cat.find_duplicates(q, 5)
    e2e5 = (time.perf_counter() - t0) / 50 * 1e6
    print(f"rows {n:>8} tiles {cat.n_tiles:>4} values {cat.n_values:>9}: b2b {b2b:6.1f} us  cold {cold:6.1f} us  kernel(cold) {np.mean(ks):6.1f} us  "
          f"mm5 b2b {mm5_b2b:5.1f} cold {mm5_cold:5.1f} us  host enqueue {host_us:5.1f} us  floor(empty query) b2b {floor_b2b:5.1f} cold {floor_cold:5.1f} us | batch8 pass {b8:6.1f} us ({b8 / 8:5.1f}/query) host {host8_us:5.1f} us | "
          f"e2e mm2 {e2e:6.1f} us ({len(full)} hits) mm5 {e2e5:6.1f} us ({len(full5)} hits)", flush=True)
    cat.close()

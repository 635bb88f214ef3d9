#!/usr/bin/env python
"""bench.py -- TVIDZ analysis hot path on B200 (contract in the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline metric (BASELINE.json): 1080p frames/s scene-cut scoring, on configs[1]
(64 concurrent synthetic 1080p30 streams on one B200).  One step = one pass of the
scoring path (luma byte-SAD + scene score + select) over the resident batch.  Scene
scoring does not shard ("replicas only"): at N > 1 every rank scores its own 64 streams
(weak scaling).  The second half of the metric -- video-pair matches/s, configs[3], one
query against 1M stored arrays sharded over the N GPUs with one NCCL all-gather of the
per-shard hit records -- is reported in the same JSON line under "matching".

Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 1920, 1080
N_STREAMS, N_FRAMES = 64, 32
CATALOGUE_ROWS = 1_000_000
PYTHON_MATCH_SAMPLE_ROWS = 100_000


def ncu_traffic(kernel: str):
    """dram read+write bytes per launch of `kernel` from the committed ncu --set full capture
    (profiles/traffic.json, written from the .ncu-rep by scripts/ncu_summary.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(kernel)
    except OSError:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------ CPU arms (oracle = test infrastructure)
def host_threads() -> int:
    """All the host threads this process may use (torchrun pins OMP_NUM_THREADS=1: ask the OS)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_scoring_baseline(frames_np: np.ndarray, budget_s: float = 10.0):
    """The C restatement of FFmpeg's scene score on all host cores, one stream per thread."""
    import oracle
    S, F = frames_np.shape[:2]
    nt = host_threads()
    oracle.scene_batch(np.ascontiguousarray(frames_np[: min(S, 8), : min(F, 4)]), n_threads=nt)   # warm
    done, t0, used = 0, time.perf_counter(), 1
    while True:
        _, _, _, used = oracle.scene_batch(frames_np, n_threads=nt)
        done += S * (F - 1)
        el = time.perf_counter() - t0
        if el >= budget_s or done >= 40 * S * (F - 1):
            break
    return {"value": done / el, "unit": "frames/s", "cores": int(used), "kind": "port",
            "sample": f"{S} streams x {F} frames of {W}x{H} luma in host RAM, {done // (S * (F - 1))} passes, "
                      f"{el:.1f} s; C restatement of ff_scene_sad_c + get_scene_score (gcc -O3 -march=native, "
                      f"OpenMP one stream per thread); excludes the video decode the reference also pays"}


def cpu_matching_baseline(ts, off, vid, q, min_match, sample_rows: int):
    """The reference's own algorithm: the Python loop of inspector/db.py:85-91 (one thread)."""
    from oracle import match_oracle
    n = min(sample_rows, len(vid))
    rows = [(int(vid[r]), ts[off[r]:off[r + 1]].tolist()) for r in range(n)]
    ql = [float(x) for x in q]
    t0 = time.perf_counter()
    res = match_oracle.find_duplicates(rows, ql, min_match)
    el = time.perf_counter() - t0
    return {"value": n / el, "unit": "pairs/s", "cores": 1, "kind": "port",
            "sample": f"{n} of {len(vid)} rows, one query of {len(ql)} cuts, min_match={min_match}, {el:.1f} s; "
                      f"pure-Python loop of db.py:85-91 over in-memory lists (excludes the per-call Postgres "
                      f"full-table fetch the reference also pays)", "hits_in_sample": len(res)}


def make_host_frames(S, F, seed=0):
    """Synthetic scenes built cheaply on the host: per stream a few random base
    images, per frame a cheap bounded perturbation."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((S, F, H, W), np.uint8)
    for s in range(S):
        t = 0
        while t < F:
            n = min(int(rng.integers(8, 20)), F - t)
            base = rng.integers(2, 251, (H, W), dtype=np.uint8)
            for k in range(n):
                noise = rng.integers(0, 5, (H, W), dtype=np.uint8)
                out[s, t + k] = base + noise - 2
            t += n
    return out


# ------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    S, F = N_STREAMS, 5                                   # bounded sample: 64 streams x 5 frames per step
    frames = make_host_frames(S, F, seed=1)
    import oracle
    nt = host_threads()
    for _ in range(max(args.warmup, 1)):
        oracle.scene_batch(frames, n_threads=nt)
    t0 = time.perf_counter()
    used = 1
    for _ in range(args.steps):
        _, _, _, used = oracle.scene_batch(frames, n_threads=nt)
    el = time.perf_counter() - t0
    value = args.steps * S * (F - 1) / el
    from tvidz_b200 import synth
    ts, off, vid = synth.synth_catalogue(PYTHON_MATCH_SAMPLE_ROWS, seed=0)
    r = 12_345
    q = ts[off[r]:off[r + 1]]
    m = cpu_matching_baseline(ts, off, vid, q, 2, PYTHON_MATCH_SAMPLE_ROWS)
    sample = (f"{S} streams x {F} frames of {W}x{H} luma per step ({S * (F - 1)} frame pairs), host RAM; "
              f"C restatement of FFmpeg scene_sad + get_scene_score, OpenMP one stream per thread; "
              f"excludes decode (no ffmpeg binary in the image)")
    line = {"impl": "reference", "metric": "1080p frames/s scene-cut scoring", "value": value, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: 64 concurrent synthetic 1080p30 streams, scene scoring "
                                   "(reference arm: CPU, bounded sample per step)",
                       "streams": S, "frames_per_stream": F, "width": W, "height": H, "threshold": 0.3},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": int(used), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "matching": {"metric": "video-pair matches/s", "value": m["value"], "unit": "pairs/s",
                         "cpu_baseline": m}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from tvidz_b200 import _lib, scene, synth
    from tvidz_b200.catalog import Catalogue
    from tvidz_b200.dist import ShardedCatalogue

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # Libraries (NCCL prints its version banner) may write to stdout: park fd 1 on stderr so that the
    # JSON line is the only thing rank 0 ever prints there.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: tvidz_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    peak_gbs, peak_src = peaks()
    K, Wm = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------- stage 1: scoring, resident in HBM
    S, F = N_STREAMS, N_FRAMES
    frames = synth.synth_frames(S, F, H, W, seed=100 + rank, scene_len=(8, 20), device=dev)
    stream = torch.cuda.current_stream()
    sptr = int(stream.cuda_stream)
    pitch, fstride, sstride = scene._strides(frames)
    assert lib.tvz_sad_luma_u8_path(frames.data_ptr(), W, H, pitch, fstride, sstride) == 1
    sad = torch.empty((S, F), dtype=torch.int64, device=dev)
    score = torch.empty((S, F), dtype=torch.float64, device=dev)
    sel = torch.empty((S, F), dtype=torch.uint8, device=dev)

    def step(ev=None):
        if ev is not None:
            ev[0].record(stream)
        _lib.check(lib.tvz_sad_luma_u8(frames.data_ptr(), S, F, W, H, pitch, fstride, sstride, sad.data_ptr(), sptr))
        if ev is not None:
            ev[1].record(stream)
        _lib.check(lib.tvz_scene_select(sad.data_ptr(), S, F, W, H, 8, 0.3, score.data_ptr(), sel.data_ptr(), sptr))

    for _ in range(Wm):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_mark0 = sampler.mark()
    e0.record(stream)
    for i in range(K):
        step(kev[i])
    e1.record(stream)
    barrier()
    t_mark1 = sampler.mark()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    sad_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))       # memset(16 KB) + SAD kernel
    clocks = sampler.stop(t_mark0, t_mark1) if rank == 0 else None
    pairs_per_step = S * (F - 1)                                        # frames SAD-ed and scored per step
    value = world * pairs_per_step * K / (total_ms * 1e-3)
    algo_bytes = pairs_per_step * W * H                                 # W*H bytes per scored frame (8d)
    achieved = algo_bytes / (sad_ms * 1e-3) / 1e9
    n_cuts = int(sel.sum().item())

    # ---------------------------------------------------------- configs[2]: long-form 4K, one stream, chunked
    longform = None
    if not args.no_longform:
        LF, LH, LW = 1024, 2160, 3840                               # one HBM-resident chunk of the 2-hour video
        lf = torch.randint(0, 256, (1, LF, LH, LW), dtype=torch.uint8, device=dev)
        lsad = torch.empty((1, LF), dtype=torch.int64, device=dev)
        lp, lfs, lss = scene._strides(lf)
        lstep = lambda: _lib.check(lib.tvz_sad_luma_u8(lf.data_ptr(), 1, LF, LW, LH, lp, lfs, lss,      # noqa: E731
                                                       lsad.data_ptr(), sptr))
        for _ in range(3):
            lstep()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0.record(stream)
        for _ in range(5):
            lstep()
        l1.record(stream)
        barrier()
        l_ms = max_over_ranks(l0.elapsed_time(l1)) / 5
        l_gbs = (LF - 1) * LH * LW / (l_ms * 1e-3) / 1e9
        longform = {"workload": "configs[2]: 2-hour 4K60 long-form video, scored in HBM-resident chunks of "
                                f"{LF} frames (one carry frame between chunks, scene.StreamScorer)",
                    "frames_per_s_4k": (LF - 1) / (l_ms * 1e-3), "ms_per_chunk": l_ms,
                    "roofline": {"bound": "hbm", "achieved": l_gbs, "peak": peak_gbs, "unit": "GB/s",
                                 "frac": l_gbs / peak_gbs, "frac_of_8TBs_nominal": l_gbs / 8000.0},
                    "projected_s_for_432000_frames": 432000 / ((LF - 1) / (l_ms * 1e-3))}
        del lf, lsad
        torch.cuda.empty_cache()

    # ---------------------------------------------------------- stage 1 end to end (host buffers)
    e2e = None
    host = torch.empty((S, F, H, W), dtype=torch.uint8).pin_memory()
    host.copy_(frames.cpu())
    e2e_steps = max(1, min(K, args.e2e_steps))
    hsad = hscore = hsel = None
    for _ in range(2):
        hsad, hscore, hsel = scene.score_frames_host(host, chunk_frames=args.chunk)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hsad, hscore, hsel = scene.score_frames_host(host, chunk_frames=args.chunk)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    el = max_over_ranks(el * 1e3) * 1e-3
    assert np.array_equal(hsad.astype(np.int64), sad.cpu().numpy()), "host-entry SADs differ from resident path"
    e2e = {"value": world * pairs_per_step * e2e_steps / el, "unit": "frames/s",
           "h2d_bytes_per_step": int(S * F * H * W), "d2h_bytes_per_step": int(S * F * (8 + 8 + 1)),
           "steps": e2e_steps, "ms_per_step": 1e3 * el / e2e_steps,
           "path": "tvz_scene_score_host: pinned host frames -> chunked H2D overlapped with SAD -> scores on host",
           "h2d_gbs": S * F * H * W * e2e_steps / el / 1e9}

    # ---------------------------------------------------------- stage 2: matching, sharded over N GPUs
    matching = None
    if not args.no_match:
        ts, off, vid = synth.synth_catalogue(CATALOGUE_ROWS, seed=0)
        r_star = 123_456
        q = ts[off[r_star]:off[r_star + 1]].copy()
        mm = 2
        cap = max(4096, (1 << 15) // world)      # per-shard hit record; regrown on overflow
        if world == 1:
            cat = Catalogue(ts, off, vid, device=local, hit_capacity=cap)
            record = torch.zeros((cap + 1, 2), dtype=torch.int32, device=dev)
            enqueue = lambda: cat.match_async(q, mm, record)                     # noqa: E731
            full = lambda: cat.find_duplicates(q, mm)                            # noqa: E731
            local_algo = cat.algo_bytes
            n_values = cat.n_values
            local_cat = cat
        else:
            sc = ShardedCatalogue(ts, off, vid, hit_capacity=cap, device=local, gather=args.gather)
            enqueue = lambda: sc.enqueue(q, mm)                                  # noqa: E731
            full = lambda: sc.find_duplicates(q, mm)                             # noqa: E731
            local_algo = sc.local.algo_bytes
            n_values = sc.local.n_values
            local_cat = sc.local
        hits = full()
        assert (int(vid[r_star]), len(q)) in hits
        Km = max(K, 20)
        for _ in range(Wm):
            enqueue()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        m0.record(stream)
        for _ in range(Km):
            enqueue()
        m1.record(stream)
        barrier()
        m_ms_b2b = max_over_ranks(m0.elapsed_time(m1)) / Km          # back to back: L2 may keep part of the shard
        # `value`: every query starts with a cold L2 -- a 256 MB buffer is rewritten and read back in between
        # (the read leaves clean lines, so the query does not also pay for the flush's write-backs); the
        # fingerprint array of the 1M-row catalogue (128 MB) would otherwise partly survive in the 126 MB L2
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        sink = torch.zeros(1, dtype=torch.int64, device=dev)
        qev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Km)]
        barrier()
        for a, b in qev:
            flush.zero_()
            sink += flush.view(torch.int64).sum()
            a.record(stream)
            enqueue()
            b.record(stream)
        barrier()
        m_ms = max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in qev])))
        del flush, sink
        barrier()
        t0 = time.perf_counter()
        for _ in range(Km):
            full()
        torch.cuda.synchronize()
        m_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3) / Km
        # the reference's default threshold (db.py:76 min_match=5): one hit instead of ~19k chance hits, so the
        # Python list of tuples that dominates the figure above is short
        full5 = (lambda: cat.find_duplicates(q, 5)) if world == 1 else (lambda: sc.find_duplicates(q, 5))
        hits5 = full5()
        barrier()
        t0 = time.perf_counter()
        for _ in range(Km):
            full5()
        torch.cuda.synchronize()
        m_e2e5 = max_over_ranks((time.perf_counter() - t0) * 1e3) / Km
        # the dominant kernel alone, bracketed by CUDA events on its own stream inside the library
        local_cat.debug_count_kernel_ms(True)
        kms = []
        for _ in range(10):
            enqueue()
            kms.append(local_cat.debug_count_kernel_ms())
        local_cat.debug_count_kernel_ms(False)
        count_ms = max_over_ranks(float(np.mean(kms)))
        streamed = 2 * n_values                                      # the count kernel streams 16-bit fingerprints
        matching = {"metric": "video-pair matches/s", "value": CATALOGUE_ROWS / (m_ms * 1e-3), "unit": "pairs/s",
                    "ms_per_query": m_ms, "ms_per_query_back_to_back": m_ms_b2b,
                    "l2": "flushed before every timed query (256 MB rewritten, then read); back_to_back = no flush, queries "
                          "pipelined on the stream", "scaling": "strong", "n_gpus": world,
                    "config": {"workload": "configs[3]: one full-duplicate query against 1M synthetic "
                                           "cut-timestamp arrays, rows sharded over the GPUs",
                               "rows": CATALOGUE_ROWS, "values": int(off[-1]), "query_len": int(len(q)),
                               "min_match": mm, "hits": len(hits),
                               "collective": "none" if world == 1 else (
                                   "fused: the compaction kernel stores each shard's hit record into every peer "
                                   "over NVLink (symmetric memory) and raises a flag; no NCCL call on the data path"
                                   if args.gather == "fused" else
                                   f"NCCL all_gather of int32 [{cap + 1},2] per-shard hit records"),
                               "catalogue_exceeds_l2": bool(local_algo > 126e6)},
                    "e2e": {"value": CATALOGUE_ROWS / (m_e2e * 1e-3), "unit": "pairs/s", "ms_per_query": m_e2e,
                            "h2d_bytes_per_step": int(len(q) * 8 * 2 + len(q) * 4),
                            "d2h_bytes_per_step": int((len(hits) + 1) * 8),
                            "path": "Catalogue.find_duplicates: host query in, Python list of tuples out",
                            "min_match_5": {"value": CATALOGUE_ROWS / (m_e2e5 * 1e-3), "unit": "pairs/s",
                                            "ms_per_query": m_e2e5, "hits": len(hits5)}},
                    "roofline": {"bound": "hbm", "achieved": local_algo / (count_ms * 1e-3) / 1e9,
                                 "peak": peak_gbs, "unit": "GB/s",
                                 "frac": local_algo / (count_ms * 1e-3) / 1e9 / peak_gbs,
                                 "traffic": ncu_traffic("match_count_kernel") if world == 1 else None,
                                 "kernel": "match_count_kernel (streams 16-bit fingerprints of the stored values; "
                                           "the ordered compaction runs in the same cooperative launch)",
                                 "kernel_ms": count_ms,
                                 "algorithmic_bytes_per_launch": int(local_algo),
                                 "streamed_bytes_per_launch": int(streamed),
                                 "achieved_streamed_bytes": streamed / (count_ms * 1e-3) / 1e9,
                                 "frac_streamed_bytes": streamed / (count_ms * 1e-3) / 1e9 / peak_gbs,
                                 "note": "per GPU (slowest rank); algorithmic bytes 8*values + 8*(rows+1) of "
                                         "the local shard per SURVEY.md 8d, which asks for this accounting even "
                                         "when a narrower lossless encoding is stored: the kernel reads 2 B per "
                                         "stored value (its filter_hash) and verifies survivors against the 8-byte "
                                         "values, so frac > 1 means less traffic than the accounting assumes, not "
                                         "more than the HBM can deliver; kernel_ms is the whole fused launch "
                                         "(count + compaction"
                                         + (" + peer stores of the hit record; the wait kernel behind it is in "
                                            "ms_per_query)" if world > 1 else ")"),
                                 "peak_source": peak_src},
                    "gpu_launches_per_query": 1 if world == 1 else 2}
        if world > 1 and not args.no_weak:
            # weak-scaling companion: 1M rows PER GPU (rank r holds rows r*1M.. of a world*1M-row catalogue)
            wts, woff, wvid = synth.synth_catalogue(CATALOGUE_ROWS, seed=1000 + rank)
            wvid = (wvid.astype(np.int64) + rank * CATALOGUE_ROWS).astype(np.int32)
            wsc = ShardedCatalogue(wts, woff, wvid, hit_capacity=1 << 15, device=local, gather=args.gather,
                                   presharded=True)
            for _ in range(Wm):
                wsc.enqueue(q, mm)
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            w0.record(stream)
            for _ in range(Km):
                wsc.enqueue(q, mm)
            w1.record(stream)
            barrier()
            w_ms = max_over_ranks(w0.elapsed_time(w1)) / Km
            w_hits = len(wsc.find_duplicates(q, mm))
            matching["weak_scaling"] = {"rows": CATALOGUE_ROWS * world, "rows_per_gpu": CATALOGUE_ROWS,
                                        "value": CATALOGUE_ROWS * world / (w_ms * 1e-3), "unit": "pairs/s",
                                        "ms_per_query": w_ms, "hits": w_hits, "scaling": "weak",
                                        "note": "same query against a catalogue of 1M rows per GPU"}
            del wsc
        if world == 1:
            # 64 concurrent analyses (one per stream of configs[1]) asking at once: 8 queries per catalogue pass
            rng = np.random.default_rng(7)
            qs = [ts[off[r]:off[r + 1]].copy() for r in rng.integers(0, CATALOGUE_ROWS, 64)]
            many = cat.find_duplicates_many(qs, mm)
            assert many[5] == cat.find_duplicates(qs[5], mm)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                cat.match_many(qs, mm)
            b_s = (time.perf_counter() - t0) / reps
            cat.debug_count_kernel_ms(True)
            cat.match_many(qs[:8], mm)
            b_kernel_ms = cat.debug_count_kernel_ms()
            cat.debug_count_kernel_ms(False)
            matching["batched"] = {"queries": 64, "queries_per_pass": 8, "value": 64 * CATALOGUE_ROWS / b_s,
                                   "unit": "pairs/s", "ms_per_64_queries": 1e3 * b_s,
                                   "count_kernel_ms_per_pass": b_kernel_ms,
                                   "kernel_pairs_per_s": 8 * CATALOGUE_ROWS / (b_kernel_ms * 1e-3),
                                   "path": "Catalogue.match_many: host queries in, numpy hit arrays out (host<->device "
                                           "copies and synchronisation inside the timed region)",
                                   "hits_total": int(sum(len(m) for m in many))}
        if rank == 0 and not args.no_cpu:
            matching["cpu_baseline"] = cpu_matching_baseline(ts, off, vid, q, mm, PYTHON_MATCH_SAMPLE_ROWS)

    # ---------------------------------------------------------- fragment mode (configs[4]), sharded like matching
    fragment = None
    if not args.no_match and not args.no_fragment:
        from tvidz_b200.dist import ShardedFragmentCatalogue
        from tvidz_b200.fragment import FragmentCatalogue, clip_query
        fts, foff, fvid = synth.synth_catalogue(100_000, len_range=(600, 1400), gap_range=(15, 150), seed=1)
        fr, f0 = 54_321, 40_000
        fq = clip_query(fts[foff[fr]:foff[fr + 1]], f0)
        fmm, fcap = 5, 1 << 12
        if world == 1:
            fcat = FragmentCatalogue(fts, foff, fvid, device=local, hit_capacity=fcap)
            frec = torch.zeros(3 * (fcap + 1), dtype=torch.int32, device=dev)
            f_enqueue = lambda: fcat.match_async(fq, fmm, frec)                  # noqa: E731
            f_full = lambda: fcat.find_fragments(fq, fmm, top_k=16)              # noqa: E731
            f_vals, f_rows = fcat.n_values, fcat.n_rows
        else:
            fsc = ShardedFragmentCatalogue(fts, foff, fvid, hit_capacity=fcap, device=local)
            f_enqueue = lambda: fsc.enqueue(fq, fmm)                             # noqa: E731
            f_full = lambda: fsc.find_fragments(fq, fmm, top_k=16)               # noqa: E731
            f_vals, f_rows = fsc.local.n_values, fsc.local.n_rows
        top = f_full()
        assert top[0][0] == int(fvid[fr]) and top[0][1] == len(fq) and abs(top[0][2] - f0 / 30.0) <= 0.008
        Kf = 10
        for _ in range(Wm):
            f_enqueue()
        f0e, f1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        f0e.record(stream)
        for _ in range(Kf):
            f_enqueue()
        f1e.record(stream)
        barrier()
        f_ms = max_over_ranks(f0e.elapsed_time(f1e)) / Kf
        # the most permissive candidate rule (anchor = 1: every agreeing adjacent pair), per-row kernel
        f_enqueue1 = (lambda: fcat.match_async(fq, fmm, frec, anchor=1)) if world == 1 else \
            (lambda: fsc.enqueue(fq, fmm, anchor=1))
        for _ in range(2):
            f_enqueue1()
        barrier()
        f0e.record(stream)
        for _ in range(Kf):
            f_enqueue1()
        f1e.record(stream)
        barrier()
        f_ms1 = max_over_ranks(f0e.elapsed_time(f1e)) / Kf
        barrier()
        t0 = time.perf_counter()
        for _ in range(Kf):
            f_full()
        torch.cuda.synchronize()
        f_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3) / Kf
        f_algo = 8 * f_vals + 8 * (f_rows + 1)
        f_actual = 4 * f_vals + 8 * (f_rows + 1)
        fragment = {"metric": "video-pair matches/s (fragment mode)", "value": 100_000 / (f_ms * 1e-3),
                    "unit": "pairs/s", "ms_per_query": f_ms, "scaling": "strong", "n_gpus": world,
                    "config": {"workload": "configs[4]: 30 s clip embedded at a random offset in one of 100k longer "
                                           "videos, rows sharded over the GPUs, top-16",
                               "rows": 100_000, "values": int(foff[-1]), "clip_cuts": len(fq), "min_match": fmm,
                               "anchor_intervals": 2,
                               "semantics": "builder-defined (reference has no fragment matcher): parity unpinned",
                               "top1": list(top[0])},
                    "e2e": {"value": 100_000 / (f_e2e * 1e-3), "unit": "pairs/s", "ms_per_query": f_e2e},
                    "anchor_1": {"value": 100_000 / (f_ms1 * 1e-3), "unit": "pairs/s", "ms_per_query": f_ms1,
                                 "note": "same query with anchor_intervals = 1 (per-row kernel, ~90 candidate "
                                         "offsets verified per row)"},
                    "roofline": {"bound": "hbm", "achieved": f_algo / (f_ms * 1e-3) / 1e9, "peak": peak_gbs,
                                 "unit": "GB/s", "frac": f_algo / (f_ms * 1e-3) / 1e9 / peak_gbs, "traffic": ncu_traffic("fragment_stream_kernel") if world == 1 else None,
                                 "achieved_actual_bytes": f_actual / (f_ms * 1e-3) / 1e9,
                                 "frac_actual_bytes": f_actual / (f_ms * 1e-3) / 1e9 / peak_gbs,
                                 "kernel": "fragment_stream_kernel<2>",
                                 "note": "per GPU, whole query (streaming fragment kernel + compaction"
                                         + (" + all_gather)" if world > 1 else ")")
                                         + "; accounted at 8 B per stored timestamp (SURVEY.md 8d), the kernel "
                                           "actually reads int32 ticks: %d bytes" % (4 * f_vals + 8 * (f_rows + 1)),
                                 "peak_source": peak_src}}

    # ---------------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_scoring_baseline(host.numpy(), budget_s=args.cpu_budget)

    if rank == 0:
        line = {"metric": "1080p frames/s scene-cut scoring", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": K, "warmup": Wm, "ms_per_step": total_ms / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "configs[1]: batch of 64 concurrent synthetic 1080p30 streams, scene "
                                       "scoring on 1xB200" + (" (one replica per GPU)" if world > 1 else ""),
                           "streams": S, "frames_per_stream": F, "width": W, "height": H, "threshold": 0.3,
                           "unit_of_work": "frame pairs SAD-ed and scored (streams x (frames-1)) per step",
                           "resident_bytes": int(S * F * H * W), "l2": "inputs (4.2 GB) exceed the 126 MB L2",
                           "cuts_found": n_cuts},
                "e2e": e2e, "gpu_launches": 2 * K,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                             "frac": achieved / peak_gbs, "traffic": ncu_traffic("sad_bulk_kernel"),
                             "kernel": "sad_bulk_kernel (TMA bulk ring, read-once)", "kernel_ms": sad_ms,
                             "algorithmic_bytes_per_launch": int(algo_bytes),
                             "note": "W*H bytes per scored frame (SURVEY.md 8d) x streams x (frames-1)",
                             "peak_source": peak_src, "frac_of_8TBs_nominal": achieved / 8000.0},
                "clocks": clocks}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if longform is not None:
            line["longform_4k"] = longform
        if matching is not None:
            line["matching"] = matching
        if fragment is not None:
            line["fragment"] = fragment
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--chunk", type=int, default=0, help="frames per H2D chunk in the host-buffer entry")
    ap.add_argument("--cpu-budget", type=float, default=10.0)
    ap.add_argument("--no-match", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fragment", action="store_true")
    ap.add_argument("--no-longform", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="how the sharded matcher exchanges per-shard hit records at N > 1")
    args = ap.parse_args()
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))


if __name__ == "__main__":
    main()

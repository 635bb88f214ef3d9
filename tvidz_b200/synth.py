"""Seeded synthetic inputs of the shapes BASELINE.json names: luma frames
with hard scene cuts, and frame-quantised cut-timestamp catalogues.  Used by bench.py,
the tests and smoke(); nothing here is on the product path."""
from __future__ import annotations

import numpy as np
import torch


def synth_frames(n_streams: int, n_frames: int, height: int, width: int, seed: int = 0,
                 scene_len=(15, 600), device="cpu", pitch: int | None = None) -> torch.Tensor:
    """uint8 [S, F, H, P]: per stream, scenes of U{scene_len} frames; every scene has a uniform
    random base image and each frame adds iid noise in {-2..2} (clipped).  Within a scene
    mafd ~ 1.6, across a cut ~ 85, so FFmpeg's score fires exactly once per scene change."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    P = width if pitch is None else pitch
    out = torch.empty((n_streams, n_frames, height, P), dtype=torch.uint8, device=device)
    if P > width:
        out[..., width:] = 0xA5                       # padding must never contribute
    for s in range(n_streams):
        t = 0
        while t < n_frames:
            n = int(torch.randint(scene_len[0], scene_len[1] + 1, (1,), generator=g, device=device).item())
            n = min(n, n_frames - t)
            base = torch.randint(0, 256, (1, height, width), generator=g, device=device, dtype=torch.int16)
            noise = torch.randint(-2, 3, (n, height, width), generator=g, device=device, dtype=torch.int16)
            out[s, t:t + n, :, :width] = (base + noise).clamp_(0, 255).to(torch.uint8)
            t += n
    return out


def round_g6(val: np.ndarray) -> np.ndarray:
    """Vectorised float("%.6g" % v) for positive finite v (FFmpeg <= 6 pts_time text)."""
    val = np.asarray(val, np.float64)
    out = np.zeros_like(val)
    nz = val > 0
    v = val[nz]
    mag = np.floor(np.log10(v)).astype(np.int64)
    pow10 = 10.0 ** np.arange(0, 40, dtype=np.float64)   # exact up to 1e22
    # scale by 10^(5-mag): multiply when mag <= 5, divide otherwise (both exact powers)
    lo = mag <= 5
    r = np.where(lo, v * pow10[np.clip(5 - mag, 0, 39)], v / pow10[np.clip(mag - 5, 0, 39)])
    r = np.rint(r)
    bump = r >= 1e6                                   # rounding carried into a seventh digit
    mag = np.where(bump, mag + 1, mag)
    r = np.where(bump, r / 10, r)
    lo = mag <= 5
    out[nz] = np.where(lo, r / pow10[np.clip(5 - mag, 0, 39)], r * pow10[np.clip(mag - 5, 0, 39)])
    return out


def synth_catalogue(n_rows: int, len_range=(8, 120), gap_range=(15, 600), fps: int = 30, seed: int = 0):
    """CSR catalogue of cut-timestamp rows: cumulative frame gaps -> float("%.6g" % (n/fps)).
    Returns (ts f64, off i64 [N+1], video_id i32 [N] = 1..N)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = rng.integers(len_range[0], len_range[1] + 1, size=n_rows, dtype=np.int64)
    off = np.zeros(n_rows + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    total = int(off[-1])
    gaps = rng.integers(gap_range[0], gap_range[1] + 1, size=total, dtype=np.int64)
    csum = np.cumsum(gaps)
    start = np.repeat(csum[off[:-1]] - gaps[off[:-1]], lens) if total else csum
    frames = csum - start                              # per-row cumulative frame index
    # exact pts_time text per distinct frame index, then a gather
    lut = np.array([float("%.6g" % ((1.0 / fps) * n)) for n in range(int(frames.max()) + 1 if total else 1)])
    ts = lut[frames]
    return ts, off, np.arange(1, n_rows + 1, dtype=np.int32)

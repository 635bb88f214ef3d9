"""oracle -- CPU restatement of the reference's analysis hot path.

TEST INFRASTRUCTURE ONLY.  Allowed importers: tests/, __graft_entry__.smoke()
(as the checker) and bench.py's cpu_baseline / --impl reference legs (as the
CPU arm).  Nothing under tvidz_b200/ imports this package; the product path
fails loudly when its CUDA library is missing instead of falling back here.

Parity status
  stage 2 (find_duplicates, streaming loop): PINNED -- tests/golden/match_*.json
      were produced by executing /root/reference/inspector/db.py and app.py
      themselves with stubbed I/O (tests/golden/gen_*.py).
  stage 1 (FFmpeg select scene score): PARITY UNPINNED -- arithmetic lives in
      FFmpeg, absent from the reference tree and from this image; pinned only
      by the known-answer tests of SURVEY.md A.5.
  fragment mode: PARITY UNPINNED -- not implemented by the reference at all.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess

import numpy as np

from . import match_oracle  # noqa: F401  (pure-Python restatement)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtvz_oracle.so")
_SIG = os.path.join(_HERE, "_build", "cpu.sig")
_lib = None


def _cpu_signature() -> str:
    """-march=native binaries are host specific: key the build on the CPU flags."""
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    h = hashlib.sha1(flags.encode())
    for name in ("scene_oracle.c", "match_oracle.c", "fragment_oracle.c", "Makefile"):
        p = os.path.join(_HERE, name)
        if os.path.exists(p):
            with open(p, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (oracle/Makefile); returns the .so path."""
    sig = _cpu_signature()
    have = None
    if os.path.exists(_SIG):
        with open(_SIG) as f:
            have = f.read().strip()
    if force or have != sig or not os.path.exists(_SO):
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
        r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
        with open(_SIG, "w") as f:
            f.write(sig)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        c = ctypes
        L.tvzo_scene_sad_u8.restype = c.c_uint64
        L.tvzo_scene_sad_u8.argtypes = [c.c_void_p, c.c_int64, c.c_void_p, c.c_int64, c.c_int, c.c_int]
        L.tvzo_scene_scores.restype = None
        L.tvzo_scene_scores.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_double,
                                        c.c_void_p, c.c_void_p]
        L.tvzo_scene_batch.restype = c.c_int
        L.tvzo_scene_batch.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int64, c.c_int64,
                                       c.c_int64, c.c_int, c.c_double, c.c_void_p, c.c_void_p, c.c_void_p,
                                       c.c_int]
        L.tvzo_pts_time_string.restype = c.c_int
        L.tvzo_pts_time_string.argtypes = [c.c_int64, c.c_int, c.c_int, c.c_int, c.c_char_p, c.c_int]
        L.tvzo_cut_timestamps.restype = c.c_int
        L.tvzo_cut_timestamps.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_void_p]
        L.tvzo_match_counts.restype = None
        L.tvzo_match_counts.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_void_p, c.c_int, c.c_void_p, c.c_int]
        L.tvzo_match_kth.restype = None
        L.tvzo_match_kth.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_void_p, c.c_int, c.c_int,
                                     c.c_void_p, c.c_int]
        L.tvzo_fragment_rows.restype = None
        L.tvzo_fragment_rows.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_void_p, c.c_int, c.c_double, c.c_int,
                                         c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_void_p, c.c_int]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------- stage 1
def scene_batch(luma: np.ndarray, width: int | None = None, threshold: float = 0.3,
                bitdepth: int = 8, n_threads: int = 0):
    """luma: uint8 (bitdepth 8) or uint16 (bitdepth 9..16) [S, F, H, P] (P = pitch in samples >= width).
    Returns (sad u64 [S,F], score f64 [S,F], selected u8 [S,F], threads_used)."""
    assert luma.dtype in (np.uint8, np.uint16) and luma.ndim == 4 and luma.flags.c_contiguous
    assert (luma.dtype == np.uint8) == (bitdepth == 8)
    S, F, H, P = luma.shape
    W = P if width is None else width
    b = luma.dtype.itemsize
    sad = np.zeros((S, F), np.uint64)
    score = np.zeros((S, F), np.float64)
    sel = np.zeros((S, F), np.uint8)
    used = lib().tvzo_scene_batch(_p(luma), S, F, W, H, P * b, H * P * b, F * H * P * b, bitdepth, threshold,
                                  _p(sad), _p(score), _p(sel), n_threads)
    return sad, score, sel, used


def scene_scores(sad: np.ndarray, width: int, height: int, threshold: float = 0.3, bitdepth: int = 8):
    """sad: uint64 [F] of one stream (sad[0] ignored) -> (score f64 [F], selected u8 [F])."""
    sad = np.ascontiguousarray(sad, np.uint64)
    F = sad.shape[0]
    score = np.zeros(F, np.float64)
    sel = np.zeros(F, np.uint8)
    lib().tvzo_scene_scores(_p(sad), F, width, height, bitdepth, threshold, _p(score), _p(sel))
    return score, sel


def pts_time_string(pts: int, tb_num: int = 1, tb_den: int = 30, mode: int = 0) -> str:
    buf = ctypes.create_string_buffer(64)
    lib().tvzo_pts_time_string(pts, tb_num, tb_den, mode, buf, 64)
    return buf.value.decode()


def cut_timestamps(selected: np.ndarray, tb_num: int = 1, tb_den: int = 30, mode: int = 0,
                   pts: np.ndarray | None = None) -> list[float]:
    """app.py:228-232 over one stream's selected flags."""
    selected = np.ascontiguousarray(selected, np.uint8)
    F = selected.shape[0]
    out = np.zeros(max(F, 1), np.float64)
    if pts is not None:
        pts = np.ascontiguousarray(pts, np.int64)
    n = lib().tvzo_cut_timestamps(_p(selected), F, _p(pts) if pts is not None else None,
                                  tb_num, tb_den, mode, _p(out))
    return [float(x) for x in out[:n]]


# ---------------------------------------------------------------- stage 2
def match_counts(ts: np.ndarray, off: np.ndarray, q, n_threads: int = 0) -> np.ndarray:
    """db.py:85-89 for every row of a CSR catalogue -> int32 [N]."""
    ts = np.ascontiguousarray(ts, np.float64)
    off = np.ascontiguousarray(off, np.int64)
    q = np.ascontiguousarray(q, np.float64)
    n = off.shape[0] - 1
    counts = np.zeros(n, np.int32)
    lib().tvzo_match_counts(_p(ts), _p(off), n, _p(q), q.shape[0], _p(counts), n_threads)
    return counts


def match_kth(ts, off, q, min_match: int, n_threads: int = 0) -> np.ndarray:
    ts = np.ascontiguousarray(ts, np.float64)
    off = np.ascontiguousarray(off, np.int64)
    q = np.ascontiguousarray(q, np.float64)
    n = off.shape[0] - 1
    kth = np.zeros(n, np.int32)
    lib().tvzo_match_kth(_p(ts), _p(off), n, _p(q), q.shape[0], min_match, _p(kth), n_threads)
    return kth


def find_duplicates_csr(ts, off, video_id, q, min_match=5, n_threads: int = 0):
    """db.py:76-94 over a CSR catalogue, result in catalogue order."""
    counts = match_counts(ts, off, q, n_threads)
    keep = np.nonzero(counts >= min_match)[0]
    vid = np.asarray(video_id)
    return [(int(vid[r]), int(counts[r])) for r in keep]


# ---------------------------------------------------------------- fragment mode (builder-defined spec)
def fragment_rows(ts, off, q, tick_hz: float = 1000.0, tol: int = 7, tol_gap: int = 14, anchor: int = 2,
                  zero_only: bool = False, n_threads: int = 0):
    """Best (score, offset ticks) per row under the interval-anchored spec -> (int32 [N], int64 [N])."""
    ts = np.ascontiguousarray(ts, np.float64)
    off = np.ascontiguousarray(off, np.int64)
    q = np.ascontiguousarray(q, np.float64)
    n = off.shape[0] - 1
    score = np.zeros(n, np.int32)
    delta = np.zeros(n, np.int64)
    lib().tvzo_fragment_rows(_p(ts), _p(off), n, _p(q), q.shape[0], tick_hz, tol, tol_gap, int(anchor), int(zero_only),
                             _p(score), _p(delta), n_threads)
    return score, delta


def find_fragments_csr(ts, off, video_id, q, min_match=5, **kw):
    """[(video_id, score, offset_ticks)] for rows whose best score >= min_match, catalogue order."""
    score, delta = fragment_rows(ts, off, q, **kw)
    keep = np.nonzero(score >= min_match)[0]
    vid = np.asarray(video_id)
    return [(int(vid[r]), int(score[r]), int(delta[r])) for r in keep]

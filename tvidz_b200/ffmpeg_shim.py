"""Drop-in at the reference's subprocess seam.

``inspector/app.py:202-209`` launches::

    stdbuf -oL -eL ffmpeg -hide_banner -loglevel info -i <file> -vf select=gt(scene\\,0.3),showinfo -f null -

and scrapes ``n:`` / ``pts_time:`` tokens from the ``showinfo`` lines on stderr (``app.py:216-232``).
This module accepts that argv, decodes the file's luma planes on the host, scores them on the GPU
(``scene.StreamScorer``: chunked, one carried frame, bit-identical to a one-shot pass) and prints
vf_showinfo-shaped lines for the selected frames, line-buffered, so an unmodified ``analyze_file``
can consume it::

    python -m tvidz_b200.ffmpeg_shim -hide_banner -loglevel info -i clip.mp4 \\
        -vf 'select=gt(scene\\,0.3),showinfo' -f null -

(a two-line ``ffmpeg`` wrapper script on PATH that execs the line above is all the reference needs).

Frame sources (host side, like the reference's own software decode):
  * ``.y4m`` (YUV4MPEG2, 8-bit 4:2:0 / 4:2:2 / 4:4:4 / mono): parsed here, no decoder involved;
  * anything OpenCV's bundled libavcodec opens: ``cv2.VideoCapture`` with ``CAP_PROP_CONVERT_RGB = 0``
    hands back the decoder's plane 0 -- for the planar YUV formats FFmpeg's select filter keeps as they
    are (yuv420p, yuvj420p, ...) that is exactly the luma plane the filter would SAD.
Timestamps: constant-frame-rate streams, ``pts_time = "%.6g" % (n / fps)`` (or the FFmpeg >= 7 text with
``fmt="f7"``), which is what FFmpeg prints for them whatever the container's tick.
"""
from __future__ import annotations

import re
import signal
import sys
from fractions import Fraction
from typing import Callable, Iterator

import numpy as np

from . import scene

CHUNK_FRAMES = 64


# ------------------------------------------------------------------ argv
def parse_ffmpeg_args(argv: list[str]) -> dict:
    """The subset of ffmpeg's command line the reference uses (app.py:202-208)."""
    out = {"input": None, "threshold": None, "showinfo": False, "loglevel": "info"}
    i = 0
    while i < len(argv):
        a = argv[i]
        if a == "-i" and i + 1 < len(argv):
            out["input"] = argv[i + 1]
            i += 2
        elif a in ("-vf", "-filter:v") and i + 1 < len(argv):
            graph = argv[i + 1]
            m = re.search(r"select\s*=\s*'?\s*gt\(\s*scene\s*\\?,\s*([0-9.eE+-]+)\s*\)", graph)
            if not m:
                raise ValueError(f"unsupported filter graph {graph!r}: expected select=gt(scene\\,T)[,showinfo]")
            out["threshold"] = float(m.group(1))
            out["showinfo"] = "showinfo" in graph
            i += 2
        elif a == "-loglevel" and i + 1 < len(argv):
            out["loglevel"] = argv[i + 1]
            i += 2
        elif a in ("-f", "-c:v") and i + 1 < len(argv):
            i += 2                                  # -f null: the output is discarded anyway
        else:
            i += 1                                  # -hide_banner, '-', -an ...
    if out["input"] is None:
        raise ValueError("no input file (-i)")
    if out["threshold"] is None:
        raise ValueError("no scene-selection filter (-vf select=gt(scene\\,T),showinfo)")
    return out


# ------------------------------------------------------------------ frame sources
def y4m_frames(path: str) -> tuple[int, int, Fraction, Iterator[np.ndarray]]:
    """(width, height, fps, iterator of uint8 [H, W] luma planes) of a YUV4MPEG2 file."""
    f = open(path, "rb")
    header = f.readline()
    if not header.startswith(b"YUV4MPEG2"):
        f.close()
        raise ValueError("not a YUV4MPEG2 file")
    w = h = None
    fps = Fraction(30, 1)
    chroma = "420"
    for tok in header.split()[1:]:
        t = tok.decode("ascii", "replace")
        if t[0] == "W":
            w = int(t[1:])
        elif t[0] == "H":
            h = int(t[1:])
        elif t[0] == "F":
            num, den = t[1:].split(":")
            fps = Fraction(int(num), int(den))
        elif t[0] == "C":
            chroma = t[1:]
    if not w or not h:
        f.close()
        raise ValueError("y4m header without W/H")
    if "p1" in chroma or "p9" in chroma:
        f.close()
        raise ValueError(f"y4m colourspace {chroma}: only 8-bit sources are handled by the shim")
    if chroma.startswith("mono"):
        frame_bytes = w * h
    elif chroma.startswith("444"):
        frame_bytes = 3 * w * h
    elif chroma.startswith("422"):
        frame_bytes = 2 * w * h
    elif chroma.startswith("420"):
        frame_bytes = w * h + 2 * (((w + 1) // 2) * ((h + 1) // 2))
    elif chroma.startswith("411"):
        frame_bytes = w * h + 2 * (((w + 3) // 4) * h)
    else:
        f.close()
        raise ValueError(f"unknown y4m colourspace {chroma}")

    def it():
        try:
            while True:
                line = f.readline()
                if not line:
                    return
                if not line.startswith(b"FRAME"):
                    raise ValueError("corrupt y4m: expected FRAME")
                buf = f.read(frame_bytes)
                if len(buf) < frame_bytes:
                    return
                yield np.frombuffer(buf, np.uint8, w * h).reshape(h, w)
        finally:
            f.close()

    return w, h, fps, it()


def cv2_frames(path: str, decode_threads: int | None = None) -> tuple[int, int, Fraction, Iterator[np.ndarray]]:
    """Luma planes through OpenCV's bundled libavcodec (software decode on the host, as in the reference).
    decode_threads: libavcodec threads for this file (None = OpenCV's default, all cores)."""
    import cv2
    if decode_threads is None:
        cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
    else:
        cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG, [cv2.CAP_PROP_N_THREADS, int(decode_threads)])
    if not cap.isOpened():
        raise ValueError(f"cannot open {path!r}")
    if not cap.set(cv2.CAP_PROP_CONVERT_RGB, 0):
        cap.release()
        raise ValueError("this OpenCV build cannot return undecorated decoder planes (CAP_PROP_CONVERT_RGB)")
    w, h = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    fps = Fraction(cap.get(cv2.CAP_PROP_FPS) or 30.0).limit_denominator(1001)

    def it():
        try:
            while True:
                ok, frame = cap.read()
                if not ok:
                    return
                frame = np.asarray(frame)
                if frame.ndim != 2 or frame.dtype != np.uint8 or frame.shape[1] != w or frame.shape[0] < h:
                    raise ValueError(f"decoder plane of shape {frame.shape}: not a planar 8-bit YUV source")
                yield frame[:h]
        finally:
            cap.release()

    return w, h, fps, it()


def open_frames(path: str, decode_threads: int | None = None):
    with open(path, "rb") as f:
        magic = f.read(9)
    return y4m_frames(path) if magic == b"YUV4MPEG2" else cv2_frames(path, decode_threads)


_PINNED: dict = {}                  # (chunk, h, w) -> idle pairs of pinned staging buffers
_PINNED_LOCK = __import__("threading").Lock()


def _pinned_pair(chunk: int, h: int, w: int):
    import torch
    with _PINNED_LOCK:
        idle = _PINNED.setdefault((chunk, h, w), [])
        if idle:
            return idle.pop()
    return [torch.empty((chunk, h, w), dtype=torch.uint8).pin_memory() for _ in range(2)]


def _release_pair(bufs) -> None:
    key = tuple(bufs[0].shape)
    with _PINNED_LOCK:
        idle = _PINNED.setdefault(key, [])
        if len(idle) < 64:
            idle.append(bufs)


def score_files(paths, threshold: float = 0.3, workers: int | None = None, chunk_frames: int = 32, fmt: str = "g6",
                device: int | None = None) -> list[dict]:
    """Several uploads at once, the way the reference runs one analysis thread per upload (app.py:43,472):
    every worker thread decodes its file on the host (libavcodec releases the GIL), stages luma chunks in
    its own pair of pinned buffers, copies them to the GPU on its own stream and scores them there
    (``scene.StreamScorer``), so decode, PCIe and the SAD kernel of different uploads overlap.
    -> [{cuts, frames, width, height}] in input order."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    import torch
    paths = list(paths)
    workers = max(1, min(len(paths), workers or len(os.sched_getaffinity(0))))
    per_file = max(1, len(os.sched_getaffinity(0)) // workers)
    dev = torch.cuda.current_device() if device is None else int(device)

    def one(path):
        w, h, fps, frames = open_frames(path, per_file)
        time_base = (fps.denominator, fps.numerator)
        with torch.cuda.device(dev):
            stream = torch.cuda.Stream()
            bufs = _pinned_pair(chunk_frames, h, w)                # page-locking is slow: the pairs are recycled
            views = [b.numpy() for b in bufs]
            free = [torch.cuda.Event(), torch.cuda.Event()]       # buffer i has left for the device
            scorer = scene.StreamScorer(threshold)
            sels = []
            b = k = total = 0

            def ship(n):
                nonlocal b
                with torch.cuda.stream(stream):
                    d = bufs[b][:n].to(torch.device("cuda", dev), non_blocking=True)
                    free[b].record(stream)
                    sels.append(scorer.feed(d)[2][0])
                b ^= 1
                free[b].synchronize()                              # (a never-recorded event is complete)

            for frame in frames:
                np.copyto(views[b][k], frame)
                k += 1
                if k == chunk_frames:
                    ship(k)
                    total, k = total + k, 0
            if k:
                ship(k)
                total += k
            stream.synchronize()
            sel = torch.cat(sels).cpu().numpy() if sels else np.zeros(0, np.uint8)
            _release_pair(bufs)
        return {"cuts": scene.cut_timestamps(sel, None, time_base, fmt), "frames": total, "width": w, "height": h}

    with ThreadPoolExecutor(workers) as pool:
        return list(pool.map(one, paths))


# ------------------------------------------------------------------ scoring + protocol
def gpu_chunk_scorer(threshold: float) -> Callable[[np.ndarray], np.ndarray]:
    """chunk uint8 [n, H, W] (host) -> selected uint8 [n]; state carried between calls on the GPU."""
    import torch
    scorer = scene.StreamScorer(threshold)

    def feed(chunk: np.ndarray) -> np.ndarray:
        t = torch.from_numpy(np.ascontiguousarray(chunk)).pin_memory().cuda(non_blocking=True)
        _, _, sel = scorer.feed(t)
        return sel[0].cpu().numpy()

    return feed


def run(argv: list[str], out=None, scorer_factory: Callable[[float], Callable] = gpu_chunk_scorer,
        fmt: str = "g6", chunk_frames: int = CHUNK_FRAMES) -> int:
    """Emulate the reference's ffmpeg invocation; returns the process exit code."""
    out = sys.stderr if out is None else out
    try:
        args = parse_ffmpeg_args(argv)
        w, h, fps, frames = open_frames(args["input"])
    except (OSError, ValueError) as e:
        print(f"{argv and argv[-1] or 'ffmpeg_shim'}: {e}", file=out, flush=True)
        return 1
    feed = scorer_factory(args["threshold"])
    time_base = (fps.denominator, fps.numerator)              # seconds per frame, as av_q2d(tb)
    print(f"Input #0, from '{args['input']}':  Stream #0:0: Video: rawvideo, {w}x{h}, {float(fps):g} fps", file=out,
          flush=True)
    n_out, t0 = 0, 0
    buf = np.empty((chunk_frames, h, w), np.uint8)
    k = 0

    def flush_chunk(n):
        nonlocal n_out, t0
        sel = np.asarray(feed(buf[:n]))
        if args["showinfo"]:
            for j in np.nonzero(sel)[0]:
                p = t0 + int(j)
                print("[Parsed_showinfo_1 @ 0x0] n:%4d pts:%7s pts_time:%-7s fmt:yuv420p sar:1/1 s:%dx%d"
                      % (n_out, p, scene.pts_time_string(p, time_base, fmt), w, h), file=out, flush=True)
                n_out += 1
        t0 += n

    for frame in frames:
        buf[k] = frame
        k += 1
        if k == chunk_frames:
            flush_chunk(k)
            k = 0
    if k:
        flush_chunk(k)
    print(f"frame={t0:5d} selected={n_out}", file=out, flush=True)
    return 0


def main() -> None:
    signal.signal(signal.SIGTERM, lambda *_: sys.exit(255))    # app.py:251 terminates the process at the first duplicate
    sys.exit(run(sys.argv[1:]))


if __name__ == "__main__":
    main()

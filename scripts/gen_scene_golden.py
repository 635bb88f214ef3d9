"""Pins stage 1 to a real FFmpeg the day one is reachable (none is in this image or on the
GPU boxes: `which ffmpeg` finds nothing, bench.py records it).

  python scripts/gen_scene_golden.py            # needs `ffmpeg` on PATH

writes tests/golden/scene_golden.json: for a few synthetic clips (written as YUV4MPEG2 from a recorded seed)
  * lavfi.scene_score of every frame, from  -vf "select='gte(scene,0)',metadata=print"
  * the pts_time tokens of the frames selected by the reference's own graph, the exact argv of
    inspector/app.py:202-208:  -vf select=gt(scene\\,0.3),showinfo -f null -
tests/test_oracle_scene.py::test_scene_golden_from_ffmpeg then requires oracle/scene_oracle.c (and, on the
GPU, the CUDA path) to reproduce both -- until the file exists that test reports itself as skipped, and
DESIGN.md says "stage 1: parity unpinned".
"""
from __future__ import annotations

import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "scene_golden.json")


def clip(seed: int, n: int, w: int, h: int) -> np.ndarray:
    """uint8 [n, h, w] luma: scenes of 5..25 frames, a random base image per scene + bounded noise, plus one
    slow fade (large mafd, small diff: must NOT be selected) and one frame pair tuned to the 0.3 float-cast edge."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, h, w), np.uint8)
    t = 0
    while t < n:
        k = min(int(rng.integers(5, 26)), n - t)
        base = rng.integers(8, 248, (h, w)).astype(np.int16)
        fade = rng.integers(0, 3) == 0
        for j in range(k):
            f = base + rng.integers(-2, 3, (h, w)) + (3 * j if fade else 0)
            out[t + j] = np.clip(f, 0, 255).astype(np.uint8)
        t += k
    if n >= 4:
        out[n - 2] = 0
        out[n - 1] = 30          # mafd exactly 30.0: min/100 = 0.3 as a double, 0.30000001 as the float av_clipf returns
    return out


def write_y4m(path: str, luma: np.ndarray, fps: int = 30) -> None:
    n, h, w = luma.shape
    chroma = np.full((h // 2) * (w // 2) * 2, 128, np.uint8).tobytes()
    with open(path, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F%d:1 Ip A1:1 C420jpeg\n" % (w, h, fps))
        for i in range(n):
            f.write(b"FRAME\n")
            f.write(luma[i].tobytes())
            f.write(chroma)


def run_ffmpeg(exe: str, path: str):
    """-> (scores per frame, pts_time tokens of the reference's selection, seconds the reference command took)"""
    r = subprocess.run([exe, "-hide_banner", "-loglevel", "info", "-i", path, "-vf",
                        "select='gte(scene,0)',metadata=print", "-f", "null", "-"], capture_output=True, text=True)
    scores = [float(x) for x in re.findall(r"lavfi\.scene_score=([0-9.eE+-]+)", r.stderr)]
    argv = [exe, "-hide_banner", "-loglevel", "info", "-i", path, "-vf", "select=gt(scene\\,0.3),showinfo",
            "-f", "null", "-"]                                     # app.py:202-208 (stdbuf only line-buffers)
    t0 = time.perf_counter()
    r2 = subprocess.run(argv, capture_output=True, text=True)
    el = time.perf_counter() - t0
    tokens = []
    for line in r2.stderr.splitlines():                            # app.py:216-232
        if "showinfo" in line and "pts_time:" in line:
            tokens.append(line.split("pts_time:")[1].split()[0])
    return scores, tokens, el


def time_reference_command(exe: str, luma: np.ndarray) -> dict:
    """bench.py hook: the reference's own command on one synthetic stream, timed (decode of raw y4m included)."""
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "clip.y4m")
        write_y4m(p, luma)
        scores, tokens, el = run_ffmpeg(exe, p)
    ver = subprocess.run([exe, "-version"], capture_output=True, text=True).stdout.splitlines()[0]
    return {"found": True, "version": ver, "frames": int(luma.shape[0]), "seconds": el,
            "frames_per_s": luma.shape[0] / el, "cuts": len(tokens),
            "command": "ffmpeg -hide_banner -loglevel info -i clip.y4m -vf select=gt(scene\\,0.3),showinfo -f null -"}


def main() -> int:
    exe = shutil.which("ffmpeg")
    if not exe:
        print("no ffmpeg on PATH: stage 1 stays unpinned (nothing written)")
        return 1
    ver = subprocess.run([exe, "-version"], capture_output=True, text=True).stdout.splitlines()[0]
    cases = []
    with tempfile.TemporaryDirectory() as d:
        for seed, n, w, h in ((1, 120, 320, 180), (2, 90, 1918, 1080), (3, 200, 64, 48)):
            luma = clip(seed, n, w, h)
            p = os.path.join(d, f"c{seed}.y4m")
            write_y4m(p, luma)
            scores, tokens, _ = run_ffmpeg(exe, p)
            cases.append({"seed": seed, "frames": n, "width": w, "height": h, "fps": 30, "scores": scores,
                          "pts_time_tokens": tokens})
    with open(OUT, "w") as f:
        json.dump({"ffmpeg": ver, "generator": "scripts/gen_scene_golden.py", "cases": cases}, f)
    print("wrote", OUT, "from", ver)
    return 0


if __name__ == "__main__":
    sys.exit(main())

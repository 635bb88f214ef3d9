"""Writes the synthetic test clips of the decode front-end (tests/golden/clip_*.webm|mpg): smooth random
textures that slide a few pixels per frame, with a hard cut every `scene` frames -- compressible, and the
scene filter selects exactly one frame per cut.  VP9 decoding is bit-exact across conformant decoders, so
the VP9 clip pins NVDEC's luma to libavcodec's; the MPEG-2 clip is the fast-to-encode bench input.

  python tests/golden/gen_clip.py out.webm VP90 1920 1080 120 30     (cv2 / libvpx: ~1 fps at 1080p)
"""
import sys

import cv2
import numpy as np


def frames(w, h, n, scene, seed=0):
    rng = np.random.default_rng(seed)
    base = None
    for i in range(n):
        if i % scene == 0:
            small = rng.integers(0, 256, (max(2, h // 30), max(2, w // 30), 3), dtype=np.uint8)
            base = cv2.resize(small, (w, h), interpolation=cv2.INTER_LINEAR)
        yield np.roll(base, 2 * (i % scene), axis=1)


def write(path, fourcc, w, h, n, scene, fps=30, seed=0):
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*fourcc), fps, (w, h))
    if not vw.isOpened():
        raise RuntimeError(f"OpenCV cannot encode {fourcc} here")
    for f in frames(w, h, n, scene, seed):
        vw.write(f)
    vw.release()


if __name__ == "__main__":
    out, cc, w, h, n, scene = sys.argv[1], sys.argv[2], *map(int, sys.argv[3:7])
    write(out, cc, w, h, n, scene)
    print("wrote", out)

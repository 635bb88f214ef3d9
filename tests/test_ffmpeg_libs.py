"""Stage 1 against REAL FFmpeg code where any is in the image: OpenCV's wheel bundles libavutil (60.8 = FFmpeg 8.0.1;
no libavfilter, so get_scene_score itself stays unpinned).  Three pieces of the path the reference launches at
inspector/app.py:202-209 live in libavutil and are exported, and are pinned here:
  * av_ts_make_time_string2  -- the pts_time text vf_showinfo prints on FFmpeg >= 7 (the "f7" convention);
  * av_expr_parse_and_eval   -- the evaluator behind select='gt(scene,0.3)': the verdict at the float32-cast edge (A.5 #3);
  * av_pixelutils_get_sad_fn -- FFmpeg's (SIMD) byte-SAD of a block: not the function the select filter calls
                                (that is libavfilter's ff_scene_sad), but the same arithmetic, block by block.
CPU only; skipped (saying so) where the library is not there."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import oracle
from tvidz_b200 import scene


def _avutil():
    try:
        import cv2
    except Exception:
        pytest.skip("no OpenCV wheel, hence no bundled libavutil")
    libs = glob.glob(os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv*libs", "libavutil-*.so*"))
    if not libs:
        pytest.skip("the OpenCV wheel bundles no libavutil here")
    L = C.CDLL(libs[0])
    L.av_version_info.restype = C.c_char_p
    return L


class AVRational(C.Structure):
    _fields_ = [("num", C.c_int), ("den", C.c_int)]


def test_pts_time_text_equals_libavutil():
    L = _avutil()
    if not hasattr(L, "av_ts_make_time_string2"):
        pytest.skip("libavutil older than FFmpeg 7: timestamp text is an inline %.6g there")
    f = L.av_ts_make_time_string2
    f.restype, f.argtypes = C.c_char_p, [C.c_char_p, C.c_int64, AVRational]
    buf = C.create_string_buffer(64)
    rng = np.random.default_rng(3)
    n = 0
    for tb in [(1, 30), (1, 15360), (1001, 30000), (1, 25), (1, 90000), (1, 1000), (1, 60)]:
        pts_list = list(range(0, 600)) + [3037, 215999, 300000, 512 * 37] + rng.integers(0, 2**40, 300).tolist()
        for pts in pts_list:
            want = f(buf, int(pts), AVRational(*tb)).decode()
            assert scene.pts_time_string(int(pts), tb, "f7") == want, (tb, pts)
            assert oracle.pts_time_string(int(pts), tb[0], tb[1], 1) == want, (tb, pts)
            n += 1
    assert n > 6000 and L.av_version_info()


def test_select_expression_equals_libavutil():
    """gt(scene,0.3) as libavutil's eval.c computes it, at the values get_scene_score can return around the edge:
    0.3 as a double is NOT selected, 0.3 rounded to float32 (what av_clipf hands back) IS."""
    L = _avutil()
    ev = L.av_expr_parse_and_eval
    ev.restype = C.c_int
    ev.argtypes = [C.POINTER(C.c_double), C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_double)] + [C.c_void_p] * 4 + \
                  [C.c_void_p, C.c_int, C.c_void_p]
    names = (C.c_char_p * 2)(b"scene", None)
    W, H = 1920, 1080
    sads = [0, 62_207_999, 62_208_000, 62_208_001, 255 * W * H, 31 * W * H, 12345678]
    for thr in ("0.3", "0.25", "0.8"):
        for sad in sads:
            score, sel = oracle.scene_scores(np.asarray([0, sad], np.uint64), W, H, float(thr))
            vals = (C.c_double * 1)(float(score[1]))
            res = C.c_double()
            assert ev(C.byref(res), ("gt(scene,%s)" % thr).encode(), names, vals, None, None, None, None, None, 0, None) == 0
            assert bool(res.value) == bool(sel[1]), (thr, sad, score[1])
    score, sel = oracle.scene_scores(np.asarray([0, 62_208_000], np.uint64), W, H, 0.3)
    assert sel[1] == 1 and score[1] == float(np.float32(0.3))          # A.5 #3, confirmed by the real evaluator above


def test_block_sad_equals_libavutil():
    L = _avutil()
    get = L.av_pixelutils_get_sad_fn
    get.restype, get.argtypes = C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_void_p]
    p = get(5, 5, 0, None)                                   # 32x32 blocks, unaligned variant
    if not p:
        pytest.skip("this libavutil was built without pixelutils")
    sad32 = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t)(p)
    rng = np.random.default_rng(5)
    H, W = 96, 160
    a = rng.integers(0, 256, (H, W), dtype=np.uint8)
    b = rng.integers(0, 256, (H, W), dtype=np.uint8)
    b[:32] = 255 - a[:32]
    total = 0
    for y in range(0, H, 32):
        for x in range(0, W, 32):
            total += sad32(a[y:, x:].ctypes.data, W, b[y:, x:].ctypes.data, W)
    o_sad, _, _, _ = oracle.scene_batch(np.stack([a, b])[None])
    assert int(o_sad[0, 1]) == total == int(np.abs(a.astype(np.int64) - b.astype(np.int64)).sum())

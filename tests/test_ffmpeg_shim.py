"""The subprocess-seam drop-in (tvidz_b200/ffmpeg_shim.py): the reference's own ffmpeg argv in, the
showinfo lines its parser expects out.  Host logic (argv, YUV4MPEG2 / OpenCV frame sources, line
protocol) is checked on the CPU with the oracle standing in for the GPU scorer; the GPU test runs the
real thing."""
import io

import numpy as np
import pytest

import oracle
from tvidz_b200 import ffmpeg_shim, scene

REF_ARGV = ['-hide_banner', '-loglevel', 'info', '-i', None, '-vf', 'select=gt(scene\\,0.3),showinfo', '-f', 'null', '-']


def _reference_parse(stderr_text):
    """inspector/app.py:216-232, restated: showinfo lines -> scene_timestamps."""
    scene_timestamps = []
    for line in stderr_text.splitlines():
        if 'showinfo' in line and 'pts_time:' in line:
            ts = float(line.split('pts_time:')[1].split(' ')[0])
            if not scene_timestamps or ts != scene_timestamps[-1]:
                scene_timestamps.append(ts)
    return scene_timestamps


def _synthetic_luma(n=150, h=90, w=160, seed=3):
    rng = np.random.default_rng(seed)
    frames, t = np.empty((n, h, w), np.uint8), 0
    while t < n:
        m = min(int(rng.integers(9, 31)), n - t)
        base = rng.integers(2, 251, (h, w), dtype=np.uint8)
        for k in range(m):
            frames[t + k] = base + rng.integers(0, 5, (h, w), dtype=np.uint8) - 2
        t += m
    return frames


def _write_y4m(path, luma, fps=(30, 1)):
    n, h, w = luma.shape
    with open(path, 'wb') as f:
        f.write(b'YUV4MPEG2 W%d H%d F%d:%d Ip A1:1 C420jpeg\n' % (w, h, fps[0], fps[1]))
        chroma = bytes([128]) * (2 * ((w + 1) // 2) * ((h + 1) // 2))
        for t in range(n):
            f.write(b'FRAME\n')
            f.write(luma[t].tobytes())
            f.write(chroma)


def _oracle_scorer_factory(threshold):
    seen = []

    def feed(chunk):
        seen.extend(np.array(chunk))
        _, _, sel, _ = oracle.scene_batch(np.stack(seen)[None], None, threshold)
        return sel[0, len(seen) - len(chunk):]
    return feed


def test_argv_of_the_reference_is_understood():
    argv = list(REF_ARGV)
    argv[4] = '/tmp/x.mp4'
    a = ffmpeg_shim.parse_ffmpeg_args(argv)
    assert a == {'input': '/tmp/x.mp4', 'threshold': 0.3, 'showinfo': True, 'loglevel': 'info'}
    assert ffmpeg_shim.parse_ffmpeg_args(['-i', 'a', '-vf', "select='gt(scene,0.45)',showinfo"])['threshold'] == 0.45
    with pytest.raises(ValueError):
        ffmpeg_shim.parse_ffmpeg_args(['-i', 'a', '-vf', 'scale=1:1'])
    with pytest.raises(ValueError):
        ffmpeg_shim.parse_ffmpeg_args(['-vf', 'select=gt(scene\\,0.3),showinfo'])


@pytest.mark.parametrize("chunk", [1, 7, 64])
def test_y4m_through_the_line_protocol(tmp_path, chunk):
    luma = _synthetic_luma()
    path = str(tmp_path / 'clip.y4m')
    _write_y4m(path, luma)
    argv = list(REF_ARGV)
    argv[4] = path
    err = io.StringIO()
    assert ffmpeg_shim.run(argv, out=err, scorer_factory=_oracle_scorer_factory, chunk_frames=chunk) == 0
    _, _, sel, _ = oracle.scene_batch(luma[None], None, 0.3)
    want = scene.cut_timestamps(sel[0])
    assert len(want) >= 3 and _reference_parse(err.getvalue()) == want
    # n: counts the selected frames, as vf_showinfo does behind select
    ns = [int(ln.split('n:')[1].split('pts:')[0]) for ln in err.getvalue().splitlines() if 'showinfo' in ln]
    assert ns == list(range(len(ns)))


def test_ntsc_rate_and_bad_inputs(tmp_path):
    luma = _synthetic_luma(n=40)
    path = str(tmp_path / 'ntsc.y4m')
    _write_y4m(path, luma, fps=(30000, 1001))
    argv = list(REF_ARGV)
    argv[4] = path
    err = io.StringIO()
    assert ffmpeg_shim.run(argv, out=err, scorer_factory=_oracle_scorer_factory) == 0
    _, _, sel, _ = oracle.scene_batch(luma[None], None, 0.3)
    assert _reference_parse(err.getvalue()) == scene.cut_timestamps(sel[0], time_base=(1001, 30000))
    argv[4] = str(tmp_path / 'missing.mp4')
    assert ffmpeg_shim.run(argv, out=io.StringIO(), scorer_factory=_oracle_scorer_factory) == 1


def test_opencv_decoder_hands_back_the_luma_plane(tmp_path):
    cv2 = pytest.importorskip("cv2")
    path = str(tmp_path / 'clip.avi')
    h, w = 96, 128
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*'MJPG'), 25, (w, h))
    if not vw.isOpened():
        pytest.skip("no MJPG encoder in this OpenCV build")
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 255, (h, w, 3), dtype=np.uint8), rng.integers(0, 255, (h, w, 3), dtype=np.uint8)
    for t in range(24):
        vw.write(a if t < 12 else b)
    vw.release()
    ww, hh, fps, frames = ffmpeg_shim.open_frames(path)
    planes = list(frames)
    assert (ww, hh, fps) == (w, h, 25) and len(planes) == 24 and planes[0].shape == (h, w)
    mad = lambda x, y: np.abs(x.astype(np.int32) - y.astype(np.int32)).mean()            # noqa: E731
    assert mad(planes[0], planes[5]) < 10 and mad(planes[11], planes[12]) > 40   # lossy codec: stills drift a little
    argv = list(REF_ARGV)
    argv[4] = path
    err = io.StringIO()
    assert ffmpeg_shim.run(argv, out=err, scorer_factory=_oracle_scorer_factory) == 0
    assert _reference_parse(err.getvalue()) == [12 / 25]


@pytest.mark.gpu
def test_shim_end_to_end_on_the_gpu(cuda, tmp_path):
    luma = _synthetic_luma(n=200, h=270, w=480, seed=9)
    path = str(tmp_path / 'clip.y4m')
    _write_y4m(path, luma)
    argv = list(REF_ARGV)
    argv[4] = path
    err = io.StringIO()
    assert ffmpeg_shim.run(argv, out=err, chunk_frames=48) == 0
    _, _, sel, _ = oracle.scene_batch(luma[None], None, 0.3)
    assert _reference_parse(err.getvalue()) == scene.cut_timestamps(sel[0])


@pytest.mark.gpu
def test_inspector_analyze_file_flags_the_reupload(cuda, tmp_path):
    """analyze_file from a local file (app.py:197 onwards): first upload -> all cuts, no duplicates;
    the same content uploaded again -> stops at its second cut and names the original (app.py:233-255)."""
    from tvidz_b200.inspector import Inspector
    luma = _synthetic_luma(n=200, h=180, w=320, seed=5)
    path = str(tmp_path / 'clip.y4m')
    _write_y4m(path, luma)
    _, _, sel, _ = oracle.scene_batch(luma[None], None, 0.3)
    cuts = scene.cut_timestamps(sel[0])
    ins = Inspector()
    first = ins.analyze_file('videos/1700000000-holiday.y4m', path)
    assert first == {'status': 'done', 'scene_cuts': cuts, 'progress': 1.0, 'total_cuts': len(cuts), 'duplicates': [],
                     'original_filename': '1700000000-holiday.y4m', 'clean_filename': 'holiday.y4m'}
    again = ins.analyze_file('videos/1700000001-copy.y4m', path)
    assert again['status'] == 'done' and again['scene_cuts'] == cuts[:2] and again['duplicates'] == ['holiday.y4m']
    bad = ins.analyze_file('videos/x.mp4', str(tmp_path / 'missing.mp4'))
    assert bad['status'] == 'error' and bad['total_cuts'] == 0

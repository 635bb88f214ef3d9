"""Stage 1 host side: scene-cut detection with the reference's contract.

The reference obtains its cut list by running
``ffmpeg -i f -vf select=gt(scene\\,0.3),showinfo -f null -`` (inspector/app.py:202-209)
and scraping ``pts_time:`` tokens from stderr (app.py:216-232).  Here the arithmetic of
the `select` filter runs in the sm_100a kernels behind include/tvidz_b200.h and this
module reproduces the text protocol: the formatted timestamp of every selected frame,
parsed back with ``float`` and consecutive-deduplicated, so ``detect_scene_cuts`` returns
the very list ``scene_timestamps`` the reference loop builds.

torch is used for device memory and streams only.
"""
from __future__ import annotations

import math
from typing import Iterable, Iterator, Sequence

import numpy as np
import torch

from ._lib import check, lib

DEFAULT_THRESHOLD = 0.3            # literal at inspector/app.py:206
TS_FORMATS = ("g6", "f7")          # FFmpeg <= 6.x "%.6g"  |  FFmpeg >= 7.0 trimmed "%.6f"


def _stream_ptr(stream) -> int:
    if stream is None:
        stream = torch.cuda.current_stream()
    return int(stream.cuda_stream)


def _as_4d(frames: torch.Tensor) -> torch.Tensor:
    if frames.dim() == 3:
        frames = frames.unsqueeze(0)
    if frames.dim() != 4:
        raise ValueError("frames must be [streams, frames, height, pitch] or [frames, height, pitch]")
    if frames.dtype not in (torch.uint8, torch.uint16, torch.int16):
        raise TypeError("frames must be 8-bit luma (torch.uint8) or 16-bit samples (torch.uint16 / int16 bits)")
    return frames


def _strides(frames: torch.Tensor):
    S, F, H, P = frames.shape
    ss, fs, ps, es = frames.stride()
    if es != 1 and P > 1:
        raise ValueError("luma rows must be contiguous samples")
    if S == 1:
        ss = max(ss, fs * F)
    b = frames.element_size()                     # strides in BYTES
    return int(ps) * b, int(fs) * b, int(ss) * b


def sad_luma(frames: torch.Tensor, width: int | None = None, stream=None) -> torch.Tensor:
    """uint64-valued SAD of consecutive frames, as int64 tensor [S, F] (SAD < 2^63).

    frames: uint8 CUDA tensor [S, F, H, P]; `width` <= P is the visible width (row padding
    never contributes, FFmpeg scene_sad.c).  out[s, 0] = 0.
    """
    frames = _as_4d(frames)
    if not frames.is_cuda:
        raise ValueError("frames must live on the GPU (use score_frames_host for host buffers)")
    S, F, H, P = frames.shape
    W = P if width is None else int(width)
    pitch, fstride, sstride = _strides(frames)
    out = torch.empty((S, F), dtype=torch.int64, device=frames.device)
    fn = lib().tvz_sad_luma_u8 if frames.element_size() == 1 else lib().tvz_sad_luma_u16
    with torch.cuda.device(frames.device):
        check(fn(frames.data_ptr(), S, F, W, H, pitch, fstride, sstride, out.data_ptr(), _stream_ptr(stream)))
    return out


def scene_select(sad: torch.Tensor, width: int, height: int, threshold: float = DEFAULT_THRESHOLD,
                 bitdepth: int = 8, stream=None):
    """FFmpeg get_scene_score + gt(scene, threshold) on device -> (score f64 [S,F], selected u8 [S,F])."""
    if sad.dim() == 1:
        sad = sad.unsqueeze(0)
    sad = sad.contiguous()
    S, F = sad.shape
    score = torch.empty((S, F), dtype=torch.float64, device=sad.device)
    sel = torch.empty((S, F), dtype=torch.uint8, device=sad.device)
    with torch.cuda.device(sad.device):
        check(lib().tvz_scene_select(sad.data_ptr(), S, F, int(width), int(height), int(bitdepth),
                                     float(threshold), score.data_ptr(), sel.data_ptr(), _stream_ptr(stream)))
    return score, sel


def score_frames(frames: torch.Tensor, width: int | None = None, threshold: float = DEFAULT_THRESHOLD,
                 stream=None, bitdepth: int | None = None):
    """Device-resident frames -> (sad, score, selected) device tensors [S, F].  16-bit sample tensors
    (yuv420p10 and deeper) need `bitdepth` (default 10), which scales mafd as f_select.c does."""
    frames = _as_4d(frames)
    S, F, H, P = frames.shape
    W = P if width is None else int(width)
    if bitdepth is None:
        bitdepth = 8 if frames.element_size() == 1 else 10
    if (bitdepth == 8) != (frames.element_size() == 1):
        raise ValueError("bitdepth 8 goes with uint8 frames, 9..16 with 16-bit frames")
    sad = sad_luma(frames, W, stream)
    score, sel = scene_select(sad, W, H, threshold, bitdepth, stream)
    return sad, score, sel


def score_frames_host(frames, width: int | None = None, threshold: float = DEFAULT_THRESHOLD,
                      chunk_frames: int = 0, device: int | None = None, bitdepth: int | None = None):
    """Host frames (numpy uint8 or CPU torch tensor [S,F,H,P], ideally pinned) -> numpy
    (sad u64, score f64, selected u8), each [S, F].  Copies host->device inside the call
    (tvz_scene_score_host): this is the end-to-end entry a binding in analyze_file uses."""
    if isinstance(frames, np.ndarray):
        if frames.dtype not in (np.uint8, np.uint16):
            raise TypeError("frames must be uint8 or uint16")
        if frames.ndim == 3:
            frames = frames[None]
        S, F, H, P = frames.shape
        esize = frames.dtype.itemsize
        ss, fs, ps, es = (x // esize for x in frames.strides)
        ptr = frames.ctypes.data
    else:
        frames = _as_4d(frames)
        if frames.is_cuda:
            raise ValueError("score_frames_host takes host memory")
        S, F, H, P = frames.shape
        esize = frames.element_size()
        ss, fs, ps, es = frames.stride()
        ptr = frames.data_ptr()
    if es != 1 and P > 1:
        raise ValueError("luma rows must be contiguous samples")
    if S == 1:
        ss = max(ss, fs * F)
    if bitdepth is None:
        bitdepth = 8 if esize == 1 else 10
    if (bitdepth == 8) != (esize == 1):
        raise ValueError("bitdepth 8 goes with uint8 frames, 9..16 with 16-bit frames")
    ss, fs, ps = ss * esize, fs * esize, ps * esize           # the C ABI takes byte strides
    W = P if width is None else int(width)
    sad = np.zeros((S, F), np.uint64)
    score = np.zeros((S, F), np.float64)
    sel = np.zeros((S, F), np.uint8)
    ctx = torch.cuda.device(device) if device is not None else torch.cuda.device(torch.cuda.current_device())
    with ctx:
        check(lib().tvz_scene_score_host(ptr, S, F, W, H, int(ps), int(fs), int(ss), int(bitdepth), float(threshold),
                                         int(chunk_frames), sad.ctypes.data, score.ctypes.data, sel.ctypes.data))
    return sad, score, sel


class StreamScorer:
    """Scores one or more concurrent streams chunk by chunk, as long-form video arrives (NVDEC
    surfaces, a host ring buffer, ...): the reference's ffmpeg keeps only O(1) state per stream
    -- the previous frame and prev_mafd (f_select.c get_scene_score) -- and so does this.

    feed(chunk [S, n, H, P] or [n, H, P] on the GPU) returns (sad, score, selected) for exactly
    those n frames, bit-identical to scoring the whole stream in one call.  Between chunks only
    the last frame of each stream (one D2D copy) and its SAD are carried.
    """

    def __init__(self, threshold: float = DEFAULT_THRESHOLD, width: int | None = None, bitdepth: int | None = None):
        self.threshold = float(threshold)
        self.width = width
        self.bitdepth = bitdepth    # None: 8 for uint8 chunks, 10 for 16-bit ones (as score_frames)
        self._pair = None           # [S, 2, H, P]: slot 0 = carried frame, slot 1 = first frame of the new chunk
        self._last_sad = None       # int64 [S]: SAD of the last frame fed (its mafd is the next prev_mafd)
        self.frames_seen = 0

    def feed(self, chunk: torch.Tensor):
        chunk = _as_4d(chunk)
        if not chunk.is_cuda:
            raise ValueError("StreamScorer.feed takes GPU frames")
        S, n, H, P = chunk.shape
        if n == 0:
            z = torch.zeros((S, 0), device=chunk.device)
            return z.long(), z.double(), z.to(torch.uint8)
        W = P if self.width is None else int(self.width)
        bitdepth = self.bitdepth if self.bitdepth is not None else (8 if chunk.element_size() == 1 else 10)
        if (bitdepth == 8) != (chunk.element_size() == 1):
            raise ValueError("bitdepth 8 goes with uint8 frames, 9..16 with 16-bit frames")
        if self._pair is not None and self._pair.dtype != chunk.dtype:
            raise TypeError("the sample type changed between chunks")
        sad = sad_luma(chunk, W)                                  # [S, n], column 0 = 0 for now
        if self._pair is None:                                    # first chunk: frame 0 has no predecessor
            self._pair = torch.empty((S, 2, H, P), dtype=chunk.dtype, device=chunk.device)
            score, sel = scene_select(sad, W, H, self.threshold, bitdepth)
        else:
            self._pair[:, 1].copy_(chunk[:, 0])
            sad[:, 0] = sad_luma(self._pair, W)[:, 1]             # carried frame vs first new frame
            # two leading columns: a dummy "frame 0", then the carried frame's SAD, whose mafd is the
            # prev_mafd of this chunk's first frame (prev_mafd is updated on every frame, A.3)
            ext = torch.cat([torch.zeros_like(sad[:, :1]), self._last_sad[:, None], sad], dim=1)
            score, sel = scene_select(ext, W, H, self.threshold, bitdepth)
            score, sel = score[:, 2:].contiguous(), sel[:, 2:].contiguous()
        self._pair[:, 0].copy_(chunk[:, -1])
        self._last_sad = sad[:, -1].clone()
        self.frames_seen += n
        return sad, score, sel


# ------------------------------------------------------------------ timestamp text protocol
def pts_time_string(pts: int, time_base: tuple[int, int] = (1, 30), fmt: str = "g6") -> str:
    """Text showinfo prints after ``pts_time:`` (libavutil/timestamp.h).

    g6: FFmpeg <= 6.x  ``"%.6g" % (av_q2d(tb) * pts)``.
    f7: FFmpeg >= 7.0  ``"%.*f"`` with precision 6 (more below 1.0), trailing zeros and a
        bare trailing point trimmed.
    """
    val = (time_base[0] / time_base[1]) * pts       # av_q2d(tb) * ts, in this order
    if fmt == "g6":
        return "%.6g" % val
    if fmt != "f7":
        raise ValueError(f"unknown timestamp format {fmt!r}; expected one of {TS_FORMATS}")
    lg = math.floor(math.log10(abs(val))) if val != 0 else -math.inf
    precision = int(-lg) + 5 if (math.isfinite(lg) and lg < 0) else 6
    s = "%.*f" % (precision, val)
    last = len(s) - 1
    while last and s[last] == "0":
        last -= 1
    while last and s[last] != "f" and not s[last].isdigit():
        last -= 1
    return s[: last + 1]


def cut_timestamps(selected: Iterable[int], pts: Sequence[int] | None = None,
                   time_base: tuple[int, int] = (1, 30), fmt: str = "g6") -> list[float]:
    """The list the loop at app.py:216-232 builds from one stream's selected frames:
    ``float(token)`` per selected frame, appended iff it differs from the last entry."""
    out: list[float] = []
    for t, flag in enumerate(selected):
        if not flag:
            continue
        ts = float(pts_time_string(int(pts[t]) if pts is not None else t, time_base, fmt))
        if not out or ts != out[-1]:
            out.append(ts)
    return out


def showinfo_lines(selected: Iterable[int], pts: Sequence[int] | None = None,
                   time_base: tuple[int, int] = (1, 30), fmt: str = "g6", width: int = 1920,
                   height: int = 1080) -> Iterator[str]:
    """stderr lines shaped like vf_showinfo's (``n:%4d pts:%7s pts_time:%-7s ...``) for the
    selected frames, so the unmodified parser at app.py:216-232 can consume this scorer."""
    n = 0
    for t, flag in enumerate(selected):
        if not flag:
            continue
        p = int(pts[t]) if pts is not None else t
        yield ("[Parsed_showinfo_1 @ 0x0] n:%4d pts:%7s pts_time:%-7s fmt:yuv420p sar:1/1 s:%dx%d"
               % (n, p, pts_time_string(p, time_base, fmt), width, height))
        n += 1


def detect_scene_cuts(frames, width: int | None = None, threshold: float = DEFAULT_THRESHOLD,
                      time_base: tuple[int, int] = (1, 30), pts: Sequence[int] | None = None,
                      fmt: str = "g6"):
    """Scene-cut timestamps of decoded luma frames, as the reference would have scraped them
    from ffmpeg (app.py:202-232).  [F,H,P] -> list[float]; [S,F,H,P] -> list of lists.
    CUDA tensors are scored in place; host arrays go through the host-buffer entry."""
    single = frames.ndim == 3
    if isinstance(frames, torch.Tensor) and frames.is_cuda:
        _, _, sel = score_frames(frames, width, threshold)
        sel = sel.cpu().numpy()
    else:
        _, _, sel = score_frames_host(frames, width, threshold)
    cuts = [cut_timestamps(row, pts, time_base, fmt) for row in sel]
    return cuts[0] if single else cuts

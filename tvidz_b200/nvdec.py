"""Decode front-end: file -> demuxed packets -> NVDEC -> luma frames in HBM -> scene scorer.

Replaces the software decode inside the reference's ffmpeg process (inspector/app.py:202-208).  OpenCV's
bundled libavformat demuxes (``CAP_PROP_FORMAT = -1`` hands back the raw packets, Annex-B for H.264 / HEVC);
the packets go to the GPU's hardware decoder through the C ABI (csrc/nvdec.cu, which loads the driver's
libnvcuvid.so.1 at run time); the luma planes land in a device ring and ``scene.StreamScorer`` reads them
there.  Raw frames never cross PCIe: a 1080p frame is ~10-100 KB on the bus instead of 2 MB.

There is no software fallback: without NVDEC (or for a codec this GPU's NVDEC does not decode) the calls
raise, and callers that want host decode use ``ffmpeg_shim.open_frames`` explicitly.
"""
from __future__ import annotations

import ctypes as C
from fractions import Fraction
from typing import Iterator

import numpy as np
import torch

from . import scene
from ._lib import TvzError, check, lib

CODECS = {"mpeg2": 1, "mpeg4": 2, "h264": 3, "hevc": 4, "vp8": 5, "vp9": 6, "av1": 7}
# container fourcc (as OpenCV reports it) -> codec
FOURCC = {"mpg2": "mpeg2", "MPG2": "mpeg2", "mpgv": "mpeg2", "m2v1": "mpeg2", "\x02\x00\x00\x10": "mpeg2",
          "FMP4": "mpeg4", "mp4v": "mpeg4", "MP4V": "mpeg4", "XVID": "mpeg4", "DIVX": "mpeg4",
          "avc1": "h264", "h264": "h264", "H264": "h264", "x264": "h264",
          "hev1": "hevc", "hvc1": "hevc", "HEVC": "hevc", "hevc": "hevc",
          "VP80": "vp8", "VP90": "vp9", "vp09": "vp9", "av01": "av1", "AV01": "av1"}


def available() -> bool:
    return bool(lib().tvz_nvdec_available())


def why_unavailable() -> str:
    lib().tvz_nvdec_available()
    return (lib().tvz_last_error() or b"").decode(errors="replace")


def caps(codec: str, bitdepth: int = 8) -> dict:
    """What this GPU's NVDEC does for `codec` (4:2:0): {supported, max_width, max_height, engines}."""
    out = np.zeros(4, np.int32)
    check(lib().tvz_nvdec_caps(CODECS[codec], int(bitdepth), out.ctypes.data))
    return {"supported": bool(out[0]), "max_width": int(out[1]), "max_height": int(out[2]), "engines": int(out[3])}


def all_caps() -> dict:
    res = {}
    for name in CODECS:
        try:
            res[name] = caps(name)
        except TvzError as e:
            res[name] = {"supported": False, "error": str(e)}
    return res


class _DevMem:
    """A raw device range as a zero-copy torch tensor source (__cuda_array_interface__)."""

    def __init__(self, ptr: int, shape: tuple, typestr: str):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


class NvDecoder:
    """One NVDEC session: feed() demuxed packets, read decoded luma frames out of a device ring."""

    def __init__(self, codec: str, ring_frames: int = 192, device: int | None = None):
        if codec not in CODECS:
            raise ValueError(f"unknown codec {codec!r}; expected one of {sorted(CODECS)}")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.codec = codec
        self._h = C.c_void_p()
        self._total = C.c_int64(0)
        self._ring = None
        self.width = self.height = 0
        self.bitdepth = 8
        self.ring_frames = int(ring_frames)
        with torch.cuda.device(self.device):
            check(lib().tvz_decoder_create(CODECS[codec], self.ring_frames, C.byref(self._h)))

    def feed(self, packet, pts: int = 0, end_of_stream: bool = False) -> int:
        """-> frames complete in the ring so far (display order)."""
        if packet is None:
            ptr, size = None, 0
        else:
            buf = np.ascontiguousarray(packet, np.uint8).reshape(-1)
            ptr, size = buf.ctypes.data, buf.shape[0]
        with torch.cuda.device(self.device):
            check(lib().tvz_decoder_feed(self._h, ptr, size, int(pts), int(end_of_stream), C.byref(self._total)))
        if self._ring is None and self._total.value > 0:
            self._map_ring()
        return int(self._total.value)

    def _map_ring(self) -> None:
        info = np.zeros(6, np.int64)
        check(lib().tvz_decoder_info(self._h, info.ctypes.data))
        self.width, self.height, self.bitdepth = int(info[0]), int(info[1]), int(info[2])
        ptr = int(lib().tvz_decoder_ring(self._h))
        typestr = "|u1" if self.bitdepth == 8 else "<u2"
        self._ring = torch.as_tensor(_DevMem(ptr, (self.ring_frames, self.height, self.width), typestr),
                                     device=torch.device("cuda", self.device))

    @property
    def frames_decoded(self) -> int:
        return int(self._total.value)

    def frames(self, first: int, n: int) -> torch.Tensor:
        """Frames [first, first + n) as a [n, H, W] view of the ring (they must not straddle its end)."""
        if n <= 0:
            return self._ring[:0]
        s = first % self.ring_frames
        if s + n > self.ring_frames:
            raise ValueError("the requested frames wrap around the ring: read them in two pieces")
        return self._ring[s:s + n]

    def pts(self, first: int, n: int) -> np.ndarray:
        out = np.zeros(n, np.int64)
        check(lib().tvz_decoder_pts(self._h, int(first), int(n), out.ctypes.data))
        return out

    def close(self) -> None:
        if self._h is not None and self._h.value:
            self._ring = None
            lib().tvz_decoder_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def open_packets(path: str):
    """(codec, width, height, fps, iterator of uint8 packet arrays) of a video file, demuxed on the host
    (OpenCV's bundled libavformat; no decoding)."""
    import cv2
    cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
    if not cap.isOpened():
        raise ValueError(f"cannot open {path!r}")
    cc = int(cap.get(cv2.CAP_PROP_FOURCC)).to_bytes(4, "little").decode("latin1")
    codec = FOURCC.get(cc)
    w, h = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    fps = Fraction(cap.get(cv2.CAP_PROP_FPS) or 30.0).limit_denominator(1001)
    if codec is None:
        cap.release()
        raise ValueError(f"container fourcc {cc!r}: no NVDEC codec known for it")
    if not cap.set(cv2.CAP_PROP_FORMAT, -1):
        cap.release()
        raise ValueError("this OpenCV build cannot hand back undecoded packets (CAP_PROP_FORMAT = -1)")

    def it() -> Iterator[np.ndarray]:
        try:
            while True:
                ok, pkt = cap.read()
                if not ok:
                    return
                yield np.asarray(pkt).reshape(-1)
        finally:
            cap.release()

    return codec, w, h, fps, it()


def score_file(path: str, threshold: float = scene.DEFAULT_THRESHOLD, chunk_frames: int = 64, fmt: str = "g6",
               device: int | None = None, keep_sad: bool = False):
    """The cut list FFmpeg's ``select=gt(scene,T),showinfo`` would print for `path` (app.py:202-232), with
    decode on NVDEC and scoring on the SMs.  -> dict(cuts, frames, width, height, fps[, sad])."""
    codec, w, h, fps, packets = open_packets(path)
    dev = torch.cuda.current_device() if device is None else int(device)
    dec = NvDecoder(codec, ring_frames=3 * chunk_frames, device=dev)
    scorer = None
    time_base = (fps.denominator, fps.numerator)
    cuts: list[float] = []
    sads = []
    consumed = 0

    def consume(n):
        nonlocal scorer, consumed
        if scorer is None:
            # P016 surfaces carry the sample in the HIGH bits of 16: SADs are 2^(16-depth) times those of the
            # low-bit layout, and dividing mafd by 2^(16-8) instead of 2^(depth-8) gives the identical double
            scorer = scene.StreamScorer(threshold, bitdepth=8 if dec.bitdepth == 8 else 16)
        sad, _, sel = scorer.feed(dec.frames(consumed, n))
        flags = sel[0].cpu().numpy()
        for j in np.nonzero(flags)[0]:
            ts = float(scene.pts_time_string(consumed + int(j), time_base, fmt))
            if not cuts or ts != cuts[-1]:                      # app.py:231
                cuts.append(ts)
        if keep_sad:
            sads.append(sad[0].cpu().numpy())
        consumed += n

    try:
        with torch.cuda.device(dev):
            for i, pkt in enumerate(packets):
                total = dec.feed(pkt, pts=i)
                while total - consumed >= chunk_frames:
                    consume(chunk_frames)
            total = dec.feed(None, end_of_stream=True)
            while total - consumed > 0:
                consume(min(chunk_frames, total - consumed))
        out = {"cuts": cuts, "frames": consumed, "width": dec.width, "height": dec.height, "fps": fps,
               "bitdepth": dec.bitdepth, "codec": codec}
        if keep_sad:
            out["sad"] = np.concatenate(sads) if sads else np.zeros(0, np.int64)
        return out
    finally:
        dec.close()

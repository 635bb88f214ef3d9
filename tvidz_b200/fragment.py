"""Fragment (offset-invariant) matching: a clip's cut list against longer stored videos
(BASELINE.json config 5).  The reference only advertises this (README.md:5); its matcher is
offset-0 exact membership (inspector/db.py:78-79).  The semantics are therefore this
package's own -- "interval-anchored alignment", stated in csrc/fragment.cu and restated by the
repo's CPU checker (test infrastructure; parity unpinned) -- and collapse to find_duplicates' match_count at
offset 0 with zero tolerance on tick-exact data.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import TVZ_ERR_OVERFLOW, check, lib
from .catalog import rows_to_csr

DEFAULT_TICK_HZ = 1000.0      # 1 ms ticks
DEFAULT_TOL = 7               # ticks: "%.6g" keeps 10 ms resolution above 1000 s (5 ms error) + rounding
DEFAULT_TOL_GAP = 14
DEFAULT_ANCHOR = 2            # consecutive agreeing intervals that nominate an offset (1 = most permissive, slower)


class FragmentCatalogue:
    """Device-resident int32 tick rows (sorted, unique) of one shard."""

    def __init__(self, ts, off, video_id, tick_hz: float = DEFAULT_TICK_HZ, device: int | None = None,
                 hit_capacity: int = 1 << 14):
        ts = np.ascontiguousarray(ts, np.float64)
        off = np.ascontiguousarray(off, np.int64)
        video_id = np.ascontiguousarray(video_id, np.int32)
        if off.ndim != 1 or off.shape[0] != video_id.shape[0] + 1:
            raise ValueError("off must have one more entry than video_id")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.tick_hz = float(tick_hz)
        self.n_rows = int(video_id.shape[0])
        self._cap = int(hit_capacity)
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().tvz_fragcat_create(ts.ctypes.data, off.ctypes.data, video_id.ctypes.data, self.n_rows,
                                           self.tick_hz, self._cap, C.byref(self._handle)))
        self.n_values = int(lib().tvz_fragcat_values(self._handle))
        # SURVEY.md 8d accounting (8 B per stored timestamp + 8 B per row) and what actually moves
        self.algo_bytes = 8 * self.n_values + 8 * (self.n_rows + 1)
        self.actual_bytes = 4 * self.n_values + 8 * (self.n_rows + 1)

    @classmethod
    def from_rows(cls, rows, **kw) -> "FragmentCatalogue":
        return cls(*rows_to_csr(rows), **kw)

    def match(self, clip_timestamps, min_match: int = 5, tol: int = DEFAULT_TOL, tol_gap: int = DEFAULT_TOL_GAP,
              anchor: int = DEFAULT_ANCHOR, zero_offset_only: bool = False):
        """-> (video_id i32 [n], score i32 [n], offset_ticks i32 [n]) in catalogue order."""
        q = np.ascontiguousarray(np.asarray(clip_timestamps, dtype=np.float64).reshape(-1))
        n_out = C.c_int64(0)
        while True:
            cap = self._cap
            vid, sc, dl = (np.empty(cap, np.int32) for _ in range(3))
            with torch.cuda.device(self.device):
                rc = lib().tvz_fragcat_match(self._handle, q.ctypes.data, q.shape[0], int(min_match), int(tol),
                                             int(tol_gap), int(anchor), int(zero_offset_only), vid.ctypes.data,
                                             sc.ctypes.data,
                                             dl.ctypes.data, cap, C.byref(n_out))
            if rc == TVZ_ERR_OVERFLOW and n_out.value > cap:
                self._cap = int(n_out.value)       # the library has grown its side; rerun
                continue
            check(rc)
            n = int(n_out.value)
            return vid[:n], sc[:n], dl[:n]

    def match_async(self, clip_timestamps, min_match: int, out: torch.Tensor, tol: int = DEFAULT_TOL,
                    tol_gap: int = DEFAULT_TOL_GAP, anchor: int = DEFAULT_ANCHOR, zero_offset_only: bool = False,
                    stream=None) -> None:
        """Enqueue on `stream`; `out`: int32 CUDA tensor [3 * (cap + 1)] (layout in include/tvidz_b200.h)."""
        if out.dtype != torch.int32 or out.dim() != 1 or out.numel() % 3 or not out.is_contiguous():
            raise ValueError("out must be a contiguous int32 [3 * (cap + 1)] tensor")
        cap = out.numel() // 3 - 1
        q = np.ascontiguousarray(np.asarray(clip_timestamps, dtype=np.float64).reshape(-1))
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        with torch.cuda.device(self.device):
            check(lib().tvz_fragcat_match_async(self._handle, q.ctypes.data, q.shape[0], int(min_match), int(tol),
                                                int(tol_gap), int(anchor), int(zero_offset_only), out.data_ptr(), cap,
                                                int(st.cuda_stream)))

    def match_gather_async(self, clip_timestamps, min_match: int, peer_record: np.ndarray, peer_flag: np.ndarray,
                           my_flags_ptr: int, out_cap: int, epoch: int, tol: int = DEFAULT_TOL,
                           tol_gap: int = DEFAULT_TOL_GAP, anchor: int = DEFAULT_ANCHOR, zero_offset_only: bool = False,
                           stream=None) -> None:
        """Enqueue one query whose record is stored straight into every peer's gather buffer by the
        compaction kernel (tvz_fragcat_match_gather_async)."""
        q = np.ascontiguousarray(np.asarray(clip_timestamps, dtype=np.float64).reshape(-1))
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        with torch.cuda.device(self.device):
            check(lib().tvz_fragcat_match_gather_async(self._handle, q.ctypes.data, q.shape[0], int(min_match), int(tol),
                                                       int(tol_gap), int(anchor), int(zero_offset_only),
                                                       int(peer_record.shape[0]), peer_record.ctypes.data,
                                                       peer_flag.ctypes.data, int(my_flags_ptr), int(out_cap),
                                                       int(epoch) & 0xffffffff, int(st.cuda_stream)))

    def find_fragments(self, clip_timestamps, min_match: int = 5, top_k: int | None = None, **kw):
        """[(video_id, score, offset_seconds)]: every stored video that contains at least `min_match`
        of the clip's cuts at one common offset.  Catalogue order, or best-first when top_k is given."""
        vid, sc, dl = self.match(clip_timestamps, min_match, **kw)
        return rank_fragments(vid, sc, dl, self.tick_hz, top_k)

    def close(self) -> None:
        if self._handle is not None and self._handle.value:
            lib().tvz_fragcat_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rank_fragments(vid, score, delta, tick_hz: float, top_k: int | None):
    vid, score, delta = np.asarray(vid), np.asarray(score), np.asarray(delta)
    if top_k is not None:
        order = np.lexsort((np.arange(vid.shape[0]), -score.astype(np.int64)))[:top_k]   # stable: best first
        vid, score, delta = vid[order], score[order], delta[order]
    return [(int(v), int(s), d / tick_hz) for v, s, d in zip(vid.tolist(), score.tolist(), delta.tolist())]


def clip_query(ts_row: np.ndarray, start_frame: int, n_frames: int = 900, fps: int = 30) -> list[float]:
    """The cut list a clip of `n_frames` cut out of a stored video at `start_frame` produces: the
    row's cuts inside the window, re-based to the clip start and formatted like showinfo would
    ("%.6g" of the clip-local time).  The cut at the clip's first frame has no predecessor frame and
    is therefore not detected."""
    frames = np.rint(np.asarray(ts_row) * fps).astype(np.int64)
    inside = frames[(frames > start_frame) & (frames < start_frame + n_frames)]
    return [float("%.6g" % ((1.0 / fps) * int(n - start_frame))) for n in inside]

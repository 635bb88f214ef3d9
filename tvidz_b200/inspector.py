"""Drop-in for the hot-path functions of the reference's inspector.

Same names, arguments and return formats as inspector/db.py (add_video, add_timestamps,
update_duplicates, find_duplicates, get_video_by_id, get_video_by_filename) and the
analysis loop of inspector/app.py:197-302, with the two compute stages running on the
GPU.  Only what the hot path needs is here: rows are kept in process memory where the
reference keeps them in Postgres (persistence is out of scope).

Module-level functions act on a default `Inspector`, like db.py's module-level session.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass, field
from typing import Iterable, Sequence

import numpy as np

from . import scene
from .catalog import Catalogue, rows_to_csr


@dataclass
class Video:                                   # db.py:12-20
    id: int
    filename: str
    thumbnail_path: str | None = None
    duplicates: list[int] = field(default_factory=list)


class Inspector:
    """Videos + `video_timestamps` rows in process memory, matched on the GPU.

    The packed device catalogue is immutable; row updates go to a small overlay catalogue and
    a tombstone set so that the reference's per-cut pattern -- add_timestamps() then
    find_duplicates() for every new cut (app.py:234-235) -- never repacks the big catalogue.
    The overlay is folded into a fresh pack once it outgrows `overlay_limit` rows.
    Result order: rows in order of their last write (the reference's own order is Postgres heap
    order, db.py:83, which also moves a row on UPDATE).
    """

    def __init__(self, device: int | None = None, overlay_limit: int | None = None):
        self._device = device
        self._lock = threading.RLock()
        self._videos: dict[int, Video] = {}
        self._rows: dict[int, list[float]] = {}     # video_id -> timestamps, ordered by last write
        self._next_id = 1
        self._overlay_limit = overlay_limit
        self._main: Catalogue | None = None         # packed snapshot
        self._main_ids: set[int] = set()
        self._tomb: set[int] = set()                # ids in the snapshot whose row was rewritten since
        self._overlay_rows: dict[int, list[float]] = {}
        self._overlay: Catalogue | None = None
        self._overlay_dirty = False
        self.repacks = 0                            # observability: full packs of the big catalogue

    # ---------------------------------------------------------------- db.py mirror
    def add_video(self, filename, thumbnail_path=None) -> Video:          # db.py:32-41
        with self._lock:
            v = Video(self._next_id, filename, thumbnail_path)
            self._videos[v.id] = v
            self._next_id += 1
            return v

    def add_timestamps(self, video_id, timestamps) -> None:               # db.py:43-64 (upsert)
        row = [float(x) for x in timestamps]
        with self._lock:
            self._rows.pop(video_id, None)
            self._rows[video_id] = row                                    # last write goes last
            if self._main is not None:
                if video_id in self._main_ids:
                    self._tomb.add(video_id)
                self._overlay_rows.pop(video_id, None)
                self._overlay_rows[video_id] = row
                self._overlay_dirty = True

    def update_duplicates(self, video_id, duplicate_ids) -> None:         # db.py:66-74
        with self._lock:
            v = self._videos.get(video_id)
            if v:
                v.duplicates = list(duplicate_ids)

    def get_video_by_id(self, video_id):                                  # db.py:96-102
        return self._videos.get(video_id)

    def get_video_by_filename(self, filename):                            # db.py:104-109
        with self._lock:
            for v in self._videos.values():
                if v.filename == filename:
                    return v
        return None

    def clear_db(self) -> None:                                           # /admin/clear-db
        with self._lock:
            self._videos.clear()
            self._rows.clear()
            self._next_id = 1
            self._drop_packs()

    def _drop_packs(self) -> None:
        for c in (self._main, self._overlay):
            if c is not None:
                c.close()
        self._main = self._overlay = None
        self._main_ids, self._tomb, self._overlay_rows = set(), set(), {}
        self._overlay_dirty = False

    def _packs(self):
        """-> (main catalogue, tombstoned ids as int32 array, overlay catalogue or None)."""
        with self._lock:
            limit = self._overlay_limit if self._overlay_limit is not None else max(1024, len(self._main_ids) // 64)
            if self._main is None or len(self._overlay_rows) > limit:
                self._drop_packs()
                self._main = Catalogue(*rows_to_csr(self._rows.items()), device=self._device)
                self._main_ids = set(self._rows.keys())
                self.repacks += 1
            if self._overlay_dirty:
                if self._overlay is not None:
                    self._overlay.close()
                self._overlay = Catalogue(*rows_to_csr(self._overlay_rows.items()), device=self._device,
                                          hit_capacity=4096)
                self._overlay_dirty = False
            tomb = np.fromiter(self._tomb, np.int32, len(self._tomb))
            return self._main, tomb, self._overlay

    def _match(self, new_timestamps, min_match: int, with_kth: bool = False):
        with self._lock:        # a concurrent upsert may retire the packs: queries on one Inspector serialise
            return self._match_locked(new_timestamps, min_match, with_kth)

    def _match_locked(self, new_timestamps, min_match: int, with_kth: bool):
        main, tomb, overlay = self._packs()
        parts = [main.match(new_timestamps, min_match, with_kth)]
        if tomb.size:
            keep = ~np.isin(parts[0][0], tomb)
            parts[0] = tuple(a[keep] for a in parts[0])
        if overlay is not None:
            parts.append(overlay.match(new_timestamps, min_match, with_kth))
        return tuple(np.concatenate(cols) for cols in zip(*parts))

    def find_duplicates(self, new_timestamps, min_match=5):               # db.py:76-94
        """[(video_id, match_count)] for every stored row with at least `min_match` of the
        query's timestamps (exact equality, self included)."""
        vid, cnt = self._match(new_timestamps, min_match)
        return list(zip(vid.tolist(), cnt.tolist()))

    # ---------------------------------------------------------------- app.py:216-302
    def analyze_cuts(self, video_id: int, pts_time_tokens: Iterable, min_match: int = 2):
        """The per-cut loop of app.py:228-255 in one catalogue pass.

        Feeds the `pts_time` values of the selected frames; returns (scene_timestamps,
        duplicate_ids).  Equivalent to re-running find_duplicates on every prefix and
        stopping at the first prefix on which another video reaches `min_match`.
        """
        cuts: list[float] = []
        for tok in pts_time_tokens:
            ts = float(tok)                                   # app.py:230
            if not cuts or ts != cuts[-1]:                    # app.py:231
                cuts.append(ts)
        if min_match < 1:
            raise ValueError("the streaming loop needs min_match >= 1")
        if not cuts:
            return cuts, []
        vid, cnt, kth = self._match(cuts, min_match, with_kth=True)
        keep = vid != video_id                                # app.py:237 drops self
        vid, kth = vid[keep], kth[keep]
        if vid.size == 0:
            stop = len(cuts)
            dup_ids: list[int] = []
        else:
            stop = int(kth.min())                             # first prefix with any hit
            dup_ids = vid[kth == stop].tolist()
        final = cuts[:stop]
        self.add_timestamps(video_id, final)                  # app.py:234 (last upsert wins)
        if dup_ids:
            self.update_duplicates(video_id, dup_ids)         # app.py:239
        return final, dup_ids

    def analyze_frames(self, key: str, frames, width: int | None = None, threshold: float = 0.3,
                       time_base=(1, 30), pts: Sequence[int] | None = None, fmt: str = "g6") -> dict:
        """analyze_file (app.py:117-322) for already-decoded luma frames of one video:
        returns the result record of app.py:293-302 (or :307-315 on error)."""
        filename = key.split('/')[-1] if key and '/' in key else key or 'unknown_file'   # app.py:122
        original = filename
        if '-' in filename and filename.split('-')[0].isdigit():                          # app.py:128-130
            original = '-'.join(filename.split('-')[1:])
        video = self.add_video(original)                                                  # app.py:150
        try:
            cuts = scene.detect_scene_cuts(frames, width, threshold, time_base, pts, fmt)
            final, dup_ids = self.analyze_cuts(video.id, cuts, min_match=2)               # app.py:235
            names = []
            for d in dup_ids:                                                             # app.py:241-245
                dv = self.get_video_by_id(d)
                if dv:
                    names.append(dv.filename)
            return {'status': 'done', 'scene_cuts': final, 'progress': 1.0, 'total_cuts': len(final),
                    'duplicates': list(set(names)) if names else [], 'original_filename': filename,
                    'clean_filename': original}
        except Exception as e:                                                            # app.py:303-315
            return {'status': 'error', 'error': str(e), 'progress': 0.0, 'total_cuts': 0, 'duplicates': [],
                    'original_filename': filename, 'clean_filename': original}


    def analyze_file(self, key: str, local_path: str, threshold: float = 0.3, fmt: str = "g6",
                     chunk_frames: int = 64) -> dict:
        """analyze_file (app.py:117-322) from the point where the upload sits in a local file
        (app.py:197): host decode of the luma planes (ffmpeg_shim.open_frames: YUV4MPEG2 or OpenCV's
        libavcodec), GPU scoring chunk by chunk, the cut timestamps FFmpeg's showinfo would print, the
        per-cut duplicate loop -- and the result record of app.py:293-302."""
        from . import ffmpeg_shim
        filename = key.split('/')[-1] if key and '/' in key else key or 'unknown_file'   # app.py:122
        original = filename
        if '-' in filename and filename.split('-')[0].isdigit():                          # app.py:128-130
            original = '-'.join(filename.split('-')[1:])
        video = self.add_video(original)                                                  # app.py:150
        try:
            w, h, fps, frames = ffmpeg_shim.open_frames(local_path)
            feed = ffmpeg_shim.gpu_chunk_scorer(threshold)
            time_base = (fps.denominator, fps.numerator)
            tokens, t0, k = [], 0, 0
            import numpy as np
            buf = np.empty((chunk_frames, h, w), np.uint8)

            def flush(n):
                nonlocal t0
                for j in np.nonzero(np.asarray(feed(buf[:n])))[0]:
                    tokens.append(scene.pts_time_string(t0 + int(j), time_base, fmt))   # the pts_time: token
                t0 += n

            for frame in frames:
                buf[k] = frame
                k += 1
                if k == chunk_frames:
                    flush(k)
                    k = 0
            if k:
                flush(k)
            final, dup_ids = self.analyze_cuts(video.id, tokens, min_match=2)             # app.py:228-255
            names = []
            for d in dup_ids:                                                             # app.py:241-245
                dv = self.get_video_by_id(d)
                if dv:
                    names.append(dv.filename)
            return {'status': 'done', 'scene_cuts': final, 'progress': 1.0, 'total_cuts': len(final),
                    'duplicates': list(set(names)) if names else [], 'original_filename': filename,
                    'clean_filename': original}
        except Exception as e:                                                            # app.py:303-315
            return {'status': 'error', 'error': str(e), 'progress': 0.0, 'total_cuts': 0, 'duplicates': [],
                    'original_filename': filename, 'clean_filename': original}


_default = Inspector()


def add_video(filename, thumbnail_path=None):
    return _default.add_video(filename, thumbnail_path)


def add_timestamps(video_id, timestamps):
    return _default.add_timestamps(video_id, timestamps)


def update_duplicates(video_id, duplicate_ids):
    return _default.update_duplicates(video_id, duplicate_ids)


def find_duplicates(new_timestamps, min_match=5):
    return _default.find_duplicates(new_timestamps, min_match)


def get_video_by_id(video_id):
    return _default.get_video_by_id(video_id)


def get_video_by_filename(filename):
    return _default.get_video_by_filename(filename)


def clear_db():
    return _default.clear_db()


def analyze_frames(key, frames, **kw):
    return _default.analyze_frames(key, frames, **kw)


def analyze_file(key, local_path, **kw):
    return _default.analyze_file(key, local_path, **kw)

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
import time
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


class _Request:
    __slots__ = ("q", "min_match", "with_kth", "result", "error", "event", "promoted")

    def __init__(self, q, min_match, with_kth):
        self.q, self.min_match, self.with_kth = q, min_match, with_kth
        self.result = self.error = None
        self.event = threading.Event()
        self.promoted = False


class Inspector:
    """Videos + `video_timestamps` rows in process memory, matched on the GPU.

    One mutable device catalogue: add_timestamps() is a device-side row upsert (the replaced row is
    neutralised in place, the new one lands in the catalogue's tail), so the reference's per-cut
    pattern -- add_timestamps() then find_duplicates() for every new cut (app.py:234-235) -- costs one
    small kernel plus one query kernel and never repacks.  A repack (full pack of the host rows)
    happens only when the tail is full of live rows.

    Concurrency: the reference runs one analysis thread per upload (app.py:43,472), each calling
    find_duplicates.  Concurrent callers are COMBINED: whoever arrives first leads, collects the
    requests that queued up meanwhile and answers up to 8 of them with one batched pass over the
    catalogue; then leadership moves on to the next waiting caller.  Upserts and query passes
    exclude one another (a query never sees half an upsert).
    Result order: rows in order of their last write (the reference's own order is Postgres heap
    order, db.py:83, which also moves a row on UPDATE).
    """

    def __init__(self, device: int | None = None, tail_values: int | None = None, hit_capacity: int | None = None):
        self._device = device
        self._lock = threading.RLock()              # host tables
        self._gpu = threading.Lock()                # one pass or one upsert at a time on the device catalogue
        self._qlock = threading.Lock()              # the request queue
        self._pending: list[_Request] = []
        self._leader = False
        self._crowd = 1
        self._videos: dict[int, Video] = {}
        self._rows: dict[int, list[float]] = {}     # video_id -> timestamps, ordered by last write
        self._next_id = 1
        self._tail_values = tail_values
        self._hit_capacity = hit_capacity
        self._cat: Catalogue | None = None          # packed snapshot + tail
        self._stale = True                          # the device catalogue must be (re)built before the next query
        self.repacks = 0                            # observability: full packs of the host rows
        self.batches: list[int] = []                # observability: queries answered per device pass (last 1024)

    # ---------------------------------------------------------------- db.py mirror
    def add_video(self, filename, thumbnail_path=None) -> Video:          # db.py:32-41
        with self._lock:
            v = Video(self._next_id, filename, thumbnail_path)
            self._videos[v.id] = v
            self._next_id += 1
            return v

    def add_timestamps(self, video_id, timestamps) -> None:               # db.py:43-64 (upsert)
        row = [float(x) for x in timestamps]
        with self._gpu:                                                   # no query pass while the row changes
            with self._lock:
                self._rows.pop(video_id, None)
                self._rows[video_id] = row                                # last write goes last
                cat = None if self._stale else self._cat
            if cat is not None and not cat.upsert(video_id, row):
                self._stale = True                                        # tail full: repack before the next query

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
        with self._gpu, self._lock:
            self._videos.clear()
            self._rows.clear()
            self._next_id = 1
            self._drop_pack()

    def load_rows(self, rows: Iterable[tuple[int, Sequence[float]]]) -> None:
        """Bulk ingest of existing `video_timestamps` rows ((video_id, timestamps) pairs), e.g. the
        table as fetched once at start-up; equivalent to add_timestamps per row."""
        with self._gpu, self._lock:
            for vid, ts in rows:
                self._rows.pop(vid, None)
                self._rows[vid] = ts if isinstance(ts, list) else [float(x) for x in ts]
                self._next_id = max(self._next_id, int(vid) + 1)
            self._stale = True

    def _drop_pack(self) -> None:
        if self._cat is not None:
            self._cat.close()
        self._cat = None
        self._stale = True

    def _catalogue(self) -> Catalogue:
        """The device catalogue, (re)packed from the host rows if needed.  Called with _gpu held."""
        if self._stale or self._cat is None:
            with self._lock:
                self._drop_pack()
                n = len(self._rows)
                tail = self._tail_values if self._tail_values is not None else max(1 << 18, 8 * n)
                kw = {} if self._hit_capacity is None else {"hit_capacity": self._hit_capacity}
                self._cat = Catalogue(*rows_to_csr(list(self._rows.items())), device=self._device, mutable=True,
                                      tail_values=tail, **kw)
                self._stale = False
                self.repacks += 1
        return self._cat

    # ---------------------------------------------------------------- combining front door
    def _run(self, batch: list[_Request]) -> None:
        """Answer the requests of one device pass (all with the same min_match, none with kth unless alone)."""
        try:
            with self._gpu:
                cat = self._catalogue()
                if len(batch) == 1:
                    r = batch[0]
                    r.result = cat.match(r.q, r.min_match, r.with_kth)
                else:
                    for r, res in zip(batch, cat.match_many([r.q for r in batch], batch[0].min_match)):
                        r.result = res
                self.batches.append(len(batch))
                if len(self.batches) > 1024:
                    del self.batches[:512]
        except BaseException as e:                  # every waiter must wake up, with the error
            for r in batch:
                if r.result is None:
                    r.error = e

    def _match(self, new_timestamps, min_match: int, with_kth: bool = False):
        q = np.ascontiguousarray(np.asarray(new_timestamps if isinstance(new_timestamps, np.ndarray)
                                            else list(new_timestamps), dtype=np.float64).reshape(-1))
        req = _Request(q, int(min_match), with_kth)
        with self._qlock:
            self._pending.append(req)
            lead = not self._leader
            if lead:
                self._leader = True
        if not lead:
            req.event.wait()
            lead = req.promoted and req.result is None and req.error is None
        if lead:
            batch_size = 8
            if self._crowd > 1:
                # other threads have been asking concurrently: they are runnable but need the interpreter to
                # queue their request -- yield to them (twice at most) before assembling the batch
                for _ in range(2):
                    time.sleep(0)
                    if len(self._pending) >= self._crowd:
                        break
            with self._qlock:
                # the leader's own request first, then compatible ones in arrival order
                self._pending.remove(req)
                batch = [req]
                if not with_kth and Catalogue.batchable(q):
                    for r in list(self._pending):
                        if len(batch) == batch_size:
                            break
                        if r.min_match == req.min_match and not r.with_kth and Catalogue.batchable(r.q):
                            batch.append(r)
                            self._pending.remove(r)
            self._run(batch)
            self._crowd = max(len(batch), self._crowd - 1)          # how many callers shared recent passes (decays)
            with self._qlock:
                nxt = self._pending[0] if self._pending else None
                if nxt is None:
                    self._leader = False
                else:
                    nxt.promoted = True             # leadership moves on
            for r in batch[1:]:
                r.event.set()
            if nxt is not None:
                nxt.event.set()
        if req.error is not None:
            raise req.error
        return req.result

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

"""Stage 2 host side: the packed `video_timestamps` catalogue and find_duplicates.

Mirrors inspector/db.py:22-27 (one float8[] row per video), db.py:43-64 (row upsert) and
db.py:76-94 (``find_duplicates(new_timestamps, min_match=5) -> [(video_id, match_count)]``):
the rows live on the GPU and one call streams their 16-bit fingerprints once through the
sm_100a tile kernel (csrc/match.cu).  Result order is catalogue (row) order -- upserted rows
last; the reference's own order is the unspecified heap order of ``query.all()`` (db.py:83).
"""
from __future__ import annotations

import ctypes as C
import itertools
import threading
from typing import Iterable, Sequence

import numpy as np
import torch

from ._lib import TVZ_ERR_OVERFLOW, TvzError, check, lib

DEFAULT_HIT_CAPACITY = 1 << 16
DEFAULT_TAIL_VALUES = 1 << 20


def rows_to_csr(rows: Iterable[tuple[int, Sequence[float]]]):
    """[(video_id, timestamps)] -> (ts f64, off i64 [N+1], video_id i32 [N])."""
    rows = rows if isinstance(rows, (list, tuple)) else list(rows)
    lens = np.fromiter((len(ts) for _, ts in rows), np.int64, len(rows))
    off = np.zeros(len(rows) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    ts = np.fromiter(itertools.chain.from_iterable(ts for _, ts in rows), np.float64, int(off[-1]))
    return ts, off, np.fromiter((vid for vid, _ in rows), np.int32, len(rows))


def _ptr(a: np.ndarray) -> int:
    """Data pointer of a contiguous array (cheaper than a.ctypes.data)."""
    return a.__array_interface__["data"][0]


def _as_query(q) -> np.ndarray:
    if isinstance(q, np.ndarray) and q.dtype == np.float64 and q.ndim == 1 and q.flags.c_contiguous:
        return q
    return np.ascontiguousarray(np.asarray(q if isinstance(q, np.ndarray) else list(q), dtype=np.float64).reshape(-1))


class _Workspace:
    """One tvz_match_ws plus the host arrays its synchronous calls fill."""
    __slots__ = ("handle", "cap", "vid", "cnt", "kth", "p_vid", "p_cnt", "p_kth")

    def __init__(self, handle, cap):
        self.handle, self.cap = handle, cap
        self.vid = np.empty(cap, np.int32)
        self.cnt = np.empty(cap, np.int32)
        self.kth = np.empty(cap, np.int32)
        self.p_vid, self.p_cnt, self.p_kth = _ptr(self.vid), _ptr(self.cnt), _ptr(self.kth)


class Catalogue:
    """A device-resident shard of `video_timestamps` rows.

    mutable=True reserves a tail for `upsert` (db.py:43-64 on the device); otherwise the shard is
    immutable once packed.  Synchronous calls check a workspace out of a small pool (the reference
    runs one analysis thread per upload, app.py:43,472, and threads come and go); the asynchronous
    calls use one dedicated workspace, because consecutive queries on a stream share its state.
    """

    def __init__(self, ts: np.ndarray, off: np.ndarray, video_id: np.ndarray, device: int | None = None,
                 hit_capacity: int = DEFAULT_HIT_CAPACITY, mutable: bool = False,
                 tail_values: int = DEFAULT_TAIL_VALUES):
        ts = np.ascontiguousarray(ts, np.float64)
        off = np.ascontiguousarray(off, np.int64)
        video_id = np.ascontiguousarray(video_id, np.int32)
        if off.ndim != 1 or off.shape[0] != video_id.shape[0] + 1:
            raise ValueError("off must have one more entry than video_id")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.hit_capacity = int(hit_capacity)
        self.mutable = bool(mutable)
        self._handle = C.c_void_p()
        self._pool: list[_Workspace] = []          # idle workspaces
        self._all_ws: list[_Workspace] = []
        self._async_ws: _Workspace | None = None
        self._lock = threading.Lock()
        with torch.cuda.device(self.device):
            if mutable:
                check(lib().tvz_catalog_create_mutable(ts.ctypes.data, off.ctypes.data, video_id.ctypes.data,
                                                       int(video_id.shape[0]), int(tail_values), C.byref(self._handle)))
            else:
                check(lib().tvz_catalog_create(ts.ctypes.data, off.ctypes.data, video_id.ctypes.data,
                                               int(video_id.shape[0]), C.byref(self._handle)))
        self.n_tiles = int(lib().tvz_catalog_tiles(self._handle))

    @classmethod
    def from_rows(cls, rows, **kw) -> "Catalogue":
        return cls(*rows_to_csr(rows), **kw)

    # sizes move with upserts
    @property
    def n_rows(self) -> int:
        return int(lib().tvz_catalog_rows(self._handle))

    @property
    def n_values(self) -> int:
        return int(lib().tvz_catalog_values(self._handle))

    @property
    def algo_bytes(self) -> int:
        return int(lib().tvz_catalog_algo_bytes(self._handle))

    def tail_info(self) -> dict:
        out = np.zeros(4, np.int64)
        check(lib().tvz_catalog_tail_info(self._handle, out.ctypes.data))
        return {"rows": int(out[0]), "values": int(out[1]), "capacity": int(out[2]), "replaced_packed_rows": int(out[3])}

    # ---- workspaces ------------------------------------------------------------------------
    def _new_ws(self, cap: int) -> _Workspace:
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().tvz_match_ws_create(self._handle, cap, C.byref(h)))
        ws = _Workspace(h, cap)
        with self._lock:
            self._all_ws.append(ws)
        return ws

    def _drop_ws(self, ws: _Workspace) -> None:
        with self._lock:
            self._all_ws.remove(ws)
        lib().tvz_match_ws_destroy(ws.handle)

    def _checkout(self, min_capacity: int = 0) -> _Workspace:
        if self._handle is None:
            raise RuntimeError("catalogue is closed")
        with self._lock:
            ws = self._pool.pop() if self._pool else None
        want = max(self.hit_capacity, min_capacity)
        if ws is not None and ws.cap < want:
            self._drop_ws(ws)
            ws = None
        if ws is None:
            self.hit_capacity = want               # later workspaces start at the size that was needed
            ws = self._new_ws(want)
        return ws

    def _checkin(self, ws: _Workspace) -> None:
        with self._lock:
            if self._handle is not None:           # (close() has already destroyed every workspace otherwise)
                self._pool.append(ws)

    def _ws_async(self, min_capacity: int) -> _Workspace:
        ws = self._async_ws
        if ws is None or ws.cap < min_capacity:
            if ws is not None:
                torch.cuda.synchronize(self.device)
                self._drop_ws(ws)
            ws = self._async_ws = self._new_ws(max(self.hit_capacity, min_capacity))
        return ws

    # ---- db.py:43-64 -----------------------------------------------------------------------
    def upsert(self, video_id: int, timestamps) -> bool:
        """Replace (or append) the row of `video_id` on the device.  False: the tail is full -- repack.
        The caller keeps queries out while this runs (Inspector does)."""
        q = _as_query(timestamps)
        rc = lib().tvz_catalog_upsert(self._handle, int(video_id), _ptr(q), q.shape[0])     # (runs on the catalogue's device)
        if rc == TVZ_ERR_OVERFLOW:
            return False
        check(rc)
        return True

    # ---- db.py:76-94 -----------------------------------------------------------------------
    def match(self, new_timestamps, min_match: int = 5, with_kth: bool = False):
        """-> (video_id i32 [n], match_count i32 [n][, kth i32 [n]]) in catalogue order."""
        q = _as_query(new_timestamps)
        n_out = C.c_int64(0)
        need = 0
        while True:
            ws = self._checkout(need)
            try:
                rc = lib().tvz_catalog_match(self._handle, ws.handle, _ptr(q), q.shape[0], int(min_match),
                                             ws.p_vid, ws.p_cnt, ws.p_kth if with_kth else None, ws.cap, C.byref(n_out))
                if rc == TVZ_ERR_OVERFLOW and n_out.value > ws.cap:
                    need = int(n_out.value)          # grow the workspace and run the query again
                    continue
                check(rc)
                n = int(n_out.value)
                if with_kth:
                    return ws.vid[:n].copy(), ws.cnt[:n].copy(), ws.kth[:n].copy()
                return ws.vid[:n].copy(), ws.cnt[:n].copy()
            finally:
                self._checkin(ws)

    @staticmethod
    def batchable(q: np.ndarray) -> bool:
        """Can this query ride in a batched pass (<= 224 distinct values, 16-bit counts)?"""
        if q.shape[0] > 65535:
            return False
        if q.shape[0] <= int(lib().tvz_catalog_batch_limit()):
            return True
        return np.unique(q[~np.isnan(q)]).shape[0] <= int(lib().tvz_catalog_batch_limit())

    def match_many(self, queries, min_match: int = 5):
        """Several find_duplicates queries at once -- up to 8 per pass over the catalogue.
        -> list of (video_id i32 [n_i], match_count i32 [n_i]) per query, catalogue order.
        Queries with more distinct values than the batch limit run through the single-query path."""
        qs = [_as_query(q) for q in queries]
        small = [i for i, q in enumerate(qs) if self.batchable(q)]
        results: list = [None] * len(qs)
        for i in set(range(len(qs))) - set(small):
            results[i] = self.match(qs[i], min_match)
        if small:
            q_off = np.zeros(len(small) + 1, np.int64)
            np.cumsum([qs[i].shape[0] for i in small], out=q_off[1:])
            q_all = np.concatenate([qs[i] for i in small]) if q_off[-1] else np.zeros(1, np.float64)
            out_off = np.zeros(len(small) + 1, np.int64)
            need_q, need_t = C.c_int64(0), C.c_int64(0)
            cap_total, need = max(self.hit_capacity, 4096), 0
            while True:
                ws = self._checkout(need)
                try:
                    vid = np.empty(cap_total, np.int32)
                    cnt = np.empty(cap_total, np.int32)
                    with torch.cuda.device(self.device):
                        rc = lib().tvz_catalog_match_batch(self._handle, ws.handle, q_all.ctypes.data,
                                                           q_off.ctypes.data, len(small), int(min_match),
                                                           vid.ctypes.data, cnt.ctypes.data, out_off.ctypes.data,
                                                           cap_total, C.byref(need_q), C.byref(need_t))
                finally:
                    self._checkin(ws)
                if rc == TVZ_ERR_OVERFLOW:
                    need = max(need, int(need_q.value))
                    cap_total = max(cap_total, int(need_t.value))
                    continue
                check(rc)
                break
            for k, i in enumerate(small):
                a, b = int(out_off[k]), int(out_off[k + 1])
                results[i] = (vid[a:b].copy(), cnt[a:b].copy())
        return results

    def find_duplicates_many(self, queries, min_match: int = 5) -> list[list[tuple[int, int]]]:
        """[find_duplicates(q, min_match) for q in queries], answered in batched catalogue passes."""
        return [list(zip(v.tolist(), c.tolist())) for v, c in self.match_many(queries, min_match)]

    def find_duplicates(self, new_timestamps, min_match: int = 5) -> list[tuple[int, int]]:
        """db.py:76-94: list of (video_id, match_count) tuples of Python ints."""
        vid, cnt = self.match(new_timestamps, min_match)
        return list(zip(vid.tolist(), cnt.tolist()))

    # ---- device-resident entry points (pipelines, the sharded matcher) -----------------------
    def _stream(self, stream) -> int:
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        return int(st.cuda_stream)

    def match_async(self, new_timestamps, min_match: int, out: torch.Tensor, stream=None) -> None:
        """Enqueue one query on `stream` (default: torch's current stream) and return at once.
        `out`: int32 CUDA tensor [cap + 1, 2] receiving the fixed-size record described in
        include/tvidz_b200.h (row 0 = [n_hits, overflow], then (video_id, match_count))."""
        if out.dtype != torch.int32 or out.dim() != 2 or out.shape[1] != 2 or not out.is_contiguous():
            raise ValueError("out must be a contiguous int32 [cap + 1, 2] tensor")
        cap = out.shape[0] - 1
        q = _as_query(new_timestamps)
        ws = self._ws_async(cap)
        check(lib().tvz_catalog_match_async(self._handle, ws.handle, _ptr(q), q.shape[0], int(min_match),
                                            out.data_ptr(), cap, self._stream(stream)))

    @staticmethod
    def _pack_queries(queries):
        qs = [q if (type(q) is np.ndarray and q.dtype == np.float64 and q.ndim == 1) else _as_query(q) for q in queries]
        q_off = np.empty(len(qs) + 1, np.int64)
        total = q_off[0] = 0
        for i, q in enumerate(qs):
            total += q.shape[0]
            q_off[i + 1] = total
        return (np.concatenate(qs) if total else np.zeros(1, np.float64)), q_off

    def match_batch_async(self, queries, min_match: int, out: torch.Tensor, stream=None) -> None:
        """Up to 8 queries in one pass; `out`: int32 CUDA tensor [n_queries, cap + 1, 2]."""
        if out.dtype != torch.int32 or out.dim() != 3 or out.shape[2] != 2 or not out.is_contiguous():
            raise ValueError("out must be a contiguous int32 [n_queries, cap + 1, 2] tensor")
        if out.shape[0] < len(queries):
            raise ValueError("out holds fewer records than there are queries")
        cap = out.shape[1] - 1
        q_all, q_off = self._pack_queries(queries)
        ws = self._ws_async(cap)
        with torch.cuda.device(self.device):
            check(lib().tvz_catalog_match_batch_async(self._handle, ws.handle, q_all.ctypes.data, q_off.ctypes.data,
                                                      len(queries), int(min_match), out.data_ptr(), cap,
                                                      self._stream(stream)))

    def match_gather_async(self, new_timestamps, min_match: int, n_peers: int, peer_record: np.ndarray, my_slots_ptr: int,
                           slot_stride_ints: int, out_cap: int, epoch: int, stream=None) -> None:
        """Enqueue one query whose per-shard record is stored straight into every peer's gather buffer (tagged
        entries) by the query's kernel, whose last CTA waits until all peers' records have landed here
        (tvz_catalog_match_gather_async).  peer_record: contiguous uint64 array of n_peers peer addresses, or of ONE
        multicast address that reaches all peers."""
        q = _as_query(new_timestamps)
        ws = self._ws_async(out_cap)
        check(lib().tvz_catalog_match_gather_async(self._handle, ws.handle, _ptr(q), q.shape[0], int(min_match),
                                                   int(n_peers), peer_record.shape[0], _ptr(peer_record),
                                                   my_slots_ptr, slot_stride_ints, out_cap,
                                                   int(epoch) & 0xffffffff, self._stream(stream)))

    def match_batch_gather_async(self, queries, min_match: int, n_peers: int, peer_record: np.ndarray, my_slots_ptr: int,
                                 slot_stride_ints: int, out_cap: int, epoch: int, stream=None) -> None:
        """The batched form of match_gather_async: a slot is [8, out_cap + 1, 4]."""
        q_all, q_off = self._pack_queries(queries)
        ws = self._ws_async(out_cap)
        with torch.cuda.device(self.device):
            check(lib().tvz_catalog_match_batch_gather_async(
                self._handle, ws.handle, q_all.ctypes.data, q_off.ctypes.data, len(queries), int(min_match),
                int(n_peers), int(peer_record.shape[0]), peer_record.ctypes.data, int(my_slots_ptr), int(slot_stride_ints),
                int(out_cap), int(epoch) & 0xffffffff, self._stream(stream)))

    def debug_count_kernel_ms(self, enable: bool | None = None) -> float | None:
        """Bench hook: enable event timing of the query kernel on the asynchronous workspace, or
        (enable=None) read the duration of the last query's kernel in ms."""
        ws = self._ws_async(0)
        if enable is not None:
            check(lib().tvz_debug_match_timing(ws.handle, int(enable)))
            return None
        ms = C.c_float(0)
        check(lib().tvz_debug_match_count_ms(ws.handle, C.byref(ms)))
        return float(ms.value)

    def close(self) -> None:
        with self._lock:
            h, self._handle = self._handle, None
            wss, self._all_ws, self._pool, self._async_ws = self._all_ws, [], [], None
        if h is None:
            return
        for ws in wss:
            lib().tvz_match_ws_destroy(ws.handle)
        if h.value:
            lib().tvz_catalog_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

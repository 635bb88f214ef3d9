"""Stage 2 host side: the packed `video_timestamps` catalogue and find_duplicates.

Mirrors inspector/db.py:22-27 (one float8[] row per video) and db.py:76-94
(``find_duplicates(new_timestamps, min_match=5) -> [(video_id, match_count)]``): the
rows live on the GPU in CSR form and one call streams them once through the sm_100a
compare-and-count kernel.  Result order is catalogue (row) order; the reference's own
order is the unspecified heap order of ``query.all()`` (db.py:83).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Iterable, Sequence

import numpy as np
import torch

from ._lib import TVZ_ERR_OVERFLOW, TvzError, check, lib

DEFAULT_HIT_CAPACITY = 1 << 16


def rows_to_csr(rows: Iterable[tuple[int, Sequence[float]]]):
    """[(video_id, timestamps)] -> (ts f64, off i64 [N+1], video_id i32 [N])."""
    vids, lens, chunks = [], [], []
    for vid, ts in rows:
        a = np.asarray(ts, dtype=np.float64).reshape(-1)
        vids.append(int(vid))
        lens.append(a.shape[0])
        chunks.append(a)
    off = np.zeros(len(lens) + 1, np.int64)
    if lens:
        np.cumsum(np.asarray(lens, np.int64), out=off[1:])
    ts = np.concatenate(chunks) if chunks else np.zeros(0, np.float64)
    return np.ascontiguousarray(ts, np.float64), off, np.asarray(vids, np.int32)


class Catalogue:
    """A device-resident shard of `video_timestamps` rows (immutable once packed)."""

    def __init__(self, ts: np.ndarray, off: np.ndarray, video_id: np.ndarray, device: int | None = None,
                 hit_capacity: int = DEFAULT_HIT_CAPACITY):
        ts = np.ascontiguousarray(ts, np.float64)
        off = np.ascontiguousarray(off, np.int64)
        video_id = np.ascontiguousarray(video_id, np.int32)
        if off.ndim != 1 or off.shape[0] != video_id.shape[0] + 1:
            raise ValueError("off must have one more entry than video_id")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.n_rows = int(video_id.shape[0])
        self.hit_capacity = int(hit_capacity)
        self._handle = C.c_void_p()
        self._tls = threading.local()
        self._all_ws: list[C.c_void_p] = []
        self._lock = threading.Lock()
        with torch.cuda.device(self.device):
            check(lib().tvz_catalog_create(ts.ctypes.data, off.ctypes.data, video_id.ctypes.data,
                                           self.n_rows, C.byref(self._handle)))
        self.n_values = int(lib().tvz_catalog_values(self._handle))
        self.algo_bytes = int(lib().tvz_catalog_algo_bytes(self._handle))

    @classmethod
    def from_rows(cls, rows, **kw) -> "Catalogue":
        return cls(*rows_to_csr(rows), **kw)

    # one workspace per host thread: the reference runs one analysis thread per upload
    # (app.py:43,472) and every find_duplicates call is independent (db.py:81,93-94)
    def _ws(self, min_capacity: int = 0):
        ws = getattr(self._tls, "ws", None)
        cap = getattr(self._tls, "cap", 0)
        if ws is None or cap < min_capacity:
            new_cap = max(self.hit_capacity, min_capacity)
            new = C.c_void_p()
            with torch.cuda.device(self.device):
                check(lib().tvz_match_ws_create(self._handle, new_cap, C.byref(new)))
            with self._lock:
                if ws is not None:
                    self._all_ws.remove(ws)
                    lib().tvz_match_ws_destroy(ws)
                self._all_ws.append(new)
            self._tls.ws, self._tls.cap = new, new_cap
            self._tls.vid = np.empty(new_cap, np.int32)
            self._tls.cnt = np.empty(new_cap, np.int32)
            self._tls.kth = np.empty(new_cap, np.int32)
            ws = new
        return ws

    def match(self, new_timestamps, min_match: int = 5, with_kth: bool = False):
        """-> (video_id i32 [n], match_count i32 [n][, kth i32 [n]]) in catalogue order."""
        if self._handle is None:
            raise RuntimeError("catalogue is closed")
        q = np.ascontiguousarray(np.asarray(list(new_timestamps) if not isinstance(new_timestamps, np.ndarray)
                                            else new_timestamps, dtype=np.float64).reshape(-1))
        n_out = C.c_int64(0)
        need = 0
        while True:
            ws = self._ws(need)
            t = self._tls
            with torch.cuda.device(self.device):
                rc = lib().tvz_catalog_match(self._handle, ws, q.ctypes.data, q.shape[0], int(min_match),
                                             t.vid.ctypes.data, t.cnt.ctypes.data,
                                             t.kth.ctypes.data if with_kth else None, t.cap, C.byref(n_out))
            if rc == TVZ_ERR_OVERFLOW and n_out.value > t.cap:
                need = int(n_out.value)          # grow the workspace and run the query again
                continue
            check(rc)
            break
        n = int(n_out.value)
        if with_kth:
            return t.vid[:n].copy(), t.cnt[:n].copy(), t.kth[:n].copy()
        return t.vid[:n].copy(), t.cnt[:n].copy()

    def match_many(self, queries, min_match: int = 5):
        """Several find_duplicates queries at once -- up to 8 per pass over the catalogue.
        -> list of (video_id i32 [n_i], match_count i32 [n_i]) per query, catalogue order.
        Queries with more distinct values than the batch limit run through the single-query path."""
        if self._handle is None:
            raise RuntimeError("catalogue is closed")
        qs = [np.ascontiguousarray(np.asarray(q, dtype=np.float64).reshape(-1)) for q in queries]
        limit = int(lib().tvz_catalog_batch_limit())
        small = [i for i, q in enumerate(qs) if np.unique(q[~np.isnan(q)]).shape[0] <= limit]
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
                ws = self._ws(need)
                vid = np.empty(cap_total, np.int32)
                cnt = np.empty(cap_total, np.int32)
                with torch.cuda.device(self.device):
                    rc = lib().tvz_catalog_match_batch(self._handle, ws, q_all.ctypes.data, q_off.ctypes.data,
                                                       len(small), int(min_match), vid.ctypes.data, cnt.ctypes.data,
                                                       out_off.ctypes.data, cap_total, C.byref(need_q),
                                                       C.byref(need_t))
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

    def match_async(self, new_timestamps, min_match: int, out: torch.Tensor, stream=None) -> None:
        """Enqueue one query on `stream` (default: torch's current stream) and return at once.
        `out`: int32 CUDA tensor [cap + 1, 2] receiving the fixed-size record described in
        include/tvidz_b200.h (row 0 = [n_hits, overflow], then (video_id, match_count))."""
        if out.dtype != torch.int32 or out.dim() != 2 or out.shape[1] != 2 or not out.is_contiguous():
            raise ValueError("out must be a contiguous int32 [cap + 1, 2] tensor")
        cap = out.shape[0] - 1
        q = np.ascontiguousarray(np.asarray(new_timestamps, dtype=np.float64).reshape(-1))
        ws = self._ws(cap)
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        with torch.cuda.device(self.device):
            check(lib().tvz_catalog_match_async(self._handle, ws, q.ctypes.data, q.shape[0], int(min_match),
                                                out.data_ptr(), cap, int(st.cuda_stream)))

    def match_gather_async(self, new_timestamps, min_match: int, peer_record: np.ndarray, peer_flag: np.ndarray,
                           my_flags_ptr: int, out_cap: int, epoch: int, stream=None) -> None:
        """Enqueue one query whose per-shard record is stored straight into every peer's gather
        buffer by the compaction phase of the query's kernel (tvz_catalog_match_gather_async)."""
        q = np.ascontiguousarray(np.asarray(new_timestamps, dtype=np.float64).reshape(-1))
        peer_record = np.ascontiguousarray(peer_record, np.uint64)
        peer_flag = np.ascontiguousarray(peer_flag, np.uint64)
        ws = self._ws(out_cap)
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        with torch.cuda.device(self.device):
            check(lib().tvz_catalog_match_gather_async(self._handle, ws, q.ctypes.data, q.shape[0], int(min_match),
                                                       int(peer_record.shape[0]), peer_record.ctypes.data,
                                                       peer_flag.ctypes.data, int(my_flags_ptr), int(out_cap),
                                                       int(epoch) & 0xffffffff, int(st.cuda_stream)))

    def debug_count_kernel_ms(self, enable: bool | None = None) -> float | None:
        """Bench hook: enable event timing of the count kernel on this thread's workspace, or
        (enable=None) read the duration of the last query's count kernel in ms."""
        ws = self._ws(0)
        if enable is not None:
            check(lib().tvz_debug_match_timing(ws, int(enable)))
            return None
        ms = C.c_float(0)
        check(lib().tvz_debug_match_count_ms(ws, C.byref(ms)))
        return float(ms.value)

    def find_duplicates(self, new_timestamps, min_match: int = 5) -> list[tuple[int, int]]:
        """db.py:76-94: list of (video_id, match_count) tuples of Python ints."""
        vid, cnt = self.match(new_timestamps, min_match)
        return list(zip(vid.tolist(), cnt.tolist()))

    def close(self) -> None:
        with self._lock:
            for ws in self._all_ws:
                lib().tvz_match_ws_destroy(ws)
            self._all_ws.clear()
            if self._handle is not None and self._handle.value:
                lib().tvz_catalog_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""Multi-GPU matcher: the catalogue is sharded by row across the ranks of one box, every
rank streams its own shard, and the only exchange is the gather of the fixed-size per-shard hit
records.  One process per GPU, torch.distributed for the plumbing (NCCL on GPUs;
the host logic is backend-agnostic and is tested with gloo).

Two gathers: "fused" -- the records live in symmetric (peer-mapped) memory and the query's own
kernel stores every hit into all peers over NVLink, raises a flag and waits for the peers' flags
(no collective launch on the data path) -- and "nccl", one all_gather_into_tensor.

Scene scoring does not shard -- streams are independent ("replicas only").
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(off: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous row ranges per rank, balanced by stored values (sum of row lengths), not by
    row count.  Concatenating the shards in rank order restores catalogue order."""
    off = np.asarray(off, np.int64)
    n = off.shape[0] - 1
    total = int(off[-1]) if n > 0 else 0
    cuts = [0]
    for r in range(1, world):
        target = (total * r) // world
        cuts.append(max(int(np.searchsorted(off, target, side="left")), cuts[-1]))
    cuts.append(n)
    cuts = [min(c, n) for c in cuts]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def take_shard(ts: np.ndarray, off: np.ndarray, video_id: np.ndarray, lo: int, hi: int):
    off = np.asarray(off, np.int64)
    return (np.ascontiguousarray(ts[off[lo]:off[hi]]), np.ascontiguousarray(off[lo:hi + 1] - off[lo]),
            np.ascontiguousarray(video_id[lo:hi]))


def merge_records(gathered: np.ndarray, cap: int):
    """gathered: int32 [world, cap + 1, 2] of per-shard records -> (pairs int32 [n, 2] in
    catalogue order, overflowed: bool, needed: int = largest per-shard hit count)."""
    parts, overflow, needed = [], False, 0
    for rec in gathered:
        n, flag = int(rec[0, 0]), int(rec[0, 1])
        needed = max(needed, n)
        if flag or n > cap:
            overflow = True
            continue
        parts.append(rec[1:1 + n])
    pairs = np.concatenate(parts) if parts else np.zeros((0, 2), np.int32)
    return pairs, overflow, needed


class _SymmetricRecords:
    """Double-buffered gather buffers in symmetric memory: 2 sets x [world] records of `rec_ints`
    int32 each, then 2 sets x [world] flags.  Rank r's record lives in slot r of every rank's buffer."""

    def __init__(self, rec_ints: int, device, group, multicast: bool = False):
        import torch.distributed._symmetric_memory as symm
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rec_ints = int(rec_ints)
        if self.rec_ints % 4:
            raise ValueError("record slots must be 16-byte multiples (the kernels store 8-byte pairs into them)")
        n_rec = 2 * self.world * self.rec_ints
        self.sym = symm.empty(n_rec + 2 * self.world + 32, dtype=torch.int32, device=device)
        self.sym.zero_()
        self.hdl = symm.rendezvous(self.sym, group if group is not None else dist.group.WORLD)
        torch.cuda.synchronize(device)
        self.hdl.barrier()
        base = np.asarray([int(p) for p in self.hdl.buffer_ptrs], np.uint64)
        # NVSwitch multicast mapping of the same buffer, if the platform offers one: a store to it lands in every rank's copy
        try:
            self.multicast = int(getattr(self.hdl, "multicast_ptr", 0) or 0) if multicast else 0
        except Exception:
            self.multicast = 0
        self.peer_record, self.peer_flag, self.my_flags, self.views, self.my_slots, self.mc_record = [], [], [], [], [], []
        for s in range(2):
            self.my_slots.append(int(base[self.rank]) + 4 * (s * self.world * self.rec_ints))
            if self.multicast:      # ONE address for this rank's slot in everybody's buffer (the layout is symmetric)
                self.mc_record.append(np.asarray(
                    [self.multicast + 4 * (s * self.world * self.rec_ints + self.rank * self.rec_ints)], np.uint64))
            self.peer_record.append(np.ascontiguousarray(
                base + np.uint64(4 * (s * self.world * self.rec_ints + self.rank * self.rec_ints))))
            self.peer_flag.append(np.ascontiguousarray(base + np.uint64(4 * (n_rec + s * self.world + self.rank))))
            self.my_flags.append(int(base[self.rank]) + 4 * (n_rec + s * self.world))
            self.views.append(self.sym[s * self.world * self.rec_ints:(s + 1) * self.world * self.rec_ints])
        self.epoch = 0

    def next(self):
        """-> (set index, epoch) of the next query; consecutive queries alternate buffer sets."""
        self.epoch += 1
        return self.epoch & 1, self.epoch


def _host_mirror(n_ints: int, pinned: bool) -> torch.Tensor:
    t = torch.zeros(n_ints, dtype=torch.int32)
    return t.pin_memory() if pinned else t


class ShardedCatalogue:
    """find_duplicates over a catalogue sharded across the ranks of the default process group.

    `local_factory(ts, off, video_id)` builds the per-rank matcher; it must offer
    ``match_async(q, min_match, out)`` filling an int32 [cap + 1, 2] record tensor on the
    rank's device (and ``match_batch_async(queries, min_match, out)`` for match_many).  The default
    is the CUDA `Catalogue` (no CPU path in the product; the gloo tests inject a CPU stand-in to
    exercise this host logic).
    """

    def __init__(self, ts, off, video_id, hit_capacity: int = 1 << 15, device=None,
                 local_factory: Callable | None = None, group=None, gather: str = "nccl", presharded: bool = False,
                 multicast: bool = True):
        """`presharded=True`: (ts, off, video_id) already ARE this rank's shard (catalogues too large to
        materialise on every rank); shards concatenate in rank order."""
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if presharded:
            n_local = torch.tensor([len(video_id)], dtype=torch.int64,
                                   device="cpu" if local_factory is not None else torch.device("cuda", device if device is not None else torch.cuda.current_device()))
            sizes = [torch.zeros_like(n_local) for _ in range(self.world)]
            dist.all_gather(sizes, n_local, group=group)
            edges = np.concatenate([[0], np.cumsum([int(x.item()) for x in sizes])])
            self.bounds = [(int(edges[r]), int(edges[r + 1])) for r in range(self.world)]
            lo, hi = self.bounds[self.rank]
            s_ts, s_off, s_vid = (np.ascontiguousarray(ts, np.float64), np.ascontiguousarray(off, np.int64),
                                  np.ascontiguousarray(video_id, np.int32))
        else:
            self.bounds = shard_bounds(off, self.world)
            lo, hi = self.bounds[self.rank]
            s_ts, s_off, s_vid = take_shard(np.asarray(ts), np.asarray(off), np.asarray(video_id), lo, hi)
        self._cuda = local_factory is None
        if local_factory is None:
            from .catalog import Catalogue
            dev = torch.cuda.current_device() if device is None else device
            self.device = torch.device("cuda", dev)
            self.local = Catalogue(s_ts, s_off, s_vid, device=dev, hit_capacity=hit_capacity)
        else:
            self.device = torch.device("cpu") if device is None else torch.device(device)
            self.local = local_factory(s_ts, s_off, s_vid)
        if gather not in ("nccl", "fused"):
            raise ValueError("gather must be 'nccl' or 'fused'")
        if gather == "fused" and min(h - l for l, h in self.bounds) == 0:
            gather = "nccl"                       # an empty shard cannot run the fused epilogue
        self.gather = gather
        self.multicast = bool(multicast)          # fused gather: ship through the NVSwitch multicast mapping when there is one
        self.n_rows_local = hi - lo
        self.n_values_local = int(s_off[-1]) if s_off.size else 0
        self.batch = 8
        self._alloc(hit_capacity)

    # ---- buffers -------------------------------------------------------------------------
    def _alloc(self, cap: int) -> None:
        self.cap = int(cap)
        # ints per record entry: 2 = (value0, value1) for NCCL; 4 = (value0, epoch, value1, epoch), the tagged form the
        # fused gather ships (every 8-byte half carries the query epoch: complete when all tags are there)
        self.ent = 4 if self.gather == "fused" else 2
        self.rec_ints = (self.cap + 1) * self.ent
        self._sym = self._sym_b = None
        self._host = _host_mirror(self.world * self.rec_ints, self._cuda)
        self._host_np = self._host.numpy().reshape(self.world, self.cap + 1, self.ent)
        self._host_b = self._host_b_np = None
        if self.gather == "nccl":
            self.record = torch.zeros((self.cap + 1, 2), dtype=torch.int32, device=self.device)
            self.gathered = torch.zeros((self.world * (self.cap + 1), 2), dtype=torch.int32, device=self.device)
            self.record_b = self.gathered_b = None
            return
        self._sym = _SymmetricRecords(self.rec_ints, self.device, self.group, multicast=self.multicast)

    def _alloc_batch(self) -> None:
        if self._host_b is None:
            self._host_b = _host_mirror(self.world * self.batch * self.rec_ints, self._cuda)
            self._host_b_np = self._host_b.numpy().reshape(self.world, self.batch, self.cap + 1, self.ent)
        if self.gather == "nccl":
            if self.record_b is None:
                self.record_b = torch.zeros((self.batch, self.cap + 1, 2), dtype=torch.int32, device=self.device)
                self.gathered_b = torch.zeros((self.world, self.batch, self.cap + 1, 2), dtype=torch.int32,
                                              device=self.device)
        elif self._sym_b is None:
            self._sym_b = _SymmetricRecords(self.batch * self.rec_ints, self.device, self.group, multicast=self.multicast)

    def _regrow(self, needed: int) -> None:
        if self.gather == "fused":
            torch.cuda.synchronize(self.device)
            self._sym.hdl.barrier()               # nobody still writes into the old buffers
        self._alloc(max(needed, 2 * self.cap))    # same records on every rank -> same decision

    # ---- one query -----------------------------------------------------------------------
    def enqueue(self, new_timestamps, min_match: int) -> torch.Tensor:
        """Local count + compaction + the gather, all on the current stream; returns the device
        tensor [world, cap + 1, ent] that holds every shard's record once the stream gets there
        (ent = 2: (value0, value1); fused gather: ent = 4, (value0, epoch, value1, epoch))."""
        if self.gather == "nccl":
            self.local.match_async(new_timestamps, min_match, self.record)
            dist.all_gather_into_tensor(self.gathered, self.record, group=self.group)
            return self.gathered.view(self.world, self.cap + 1, 2)
        sy = self._sym
        s, epoch = sy.next()
        dst = sy.mc_record[s] if sy.multicast else sy.peer_record[s]
        self.local.match_gather_async(new_timestamps, min_match, self.world, dst, sy.my_slots[s], self.rec_ints,
                                      self.cap, epoch)
        return sy.views[s].view(self.world, self.cap + 1, self.ent)

    def _read(self, g: torch.Tensor, host: torch.Tensor, host_np: np.ndarray, n_records: int, stride: int):
        """Every record's header + an optimistic first slice of its hits in ONE device-to-host copy; the
        rest (rare) in a second one.  -> headers int32 [n_records, 2] = (n_hits, overflow)"""
        first = min(self.cap, 1024)
        eb = 4 * self.ent                                         # bytes per entry
        flat = host_np.reshape(n_records, self.cap + 1, self.ent)
        if self._cuda:
            from ._lib import check, lib
            st = int(torch.cuda.current_stream(self.device).cuda_stream)
            check(lib().tvz_copy_records_to_host(g.data_ptr(), host.data_ptr(), n_records, 4 * stride, 0,
                                                 eb * (first + 1), 1, st))
            heads = flat[:, 0, ::self.ent // 2]
            n_max = int(heads[:, 0].max())
            if not heads[:, 1].any() and first < n_max <= self.cap:
                check(lib().tvz_copy_records_to_host(g.data_ptr(), host.data_ptr(), n_records, 4 * stride,
                                                     eb * (first + 1), eb * (n_max - first), 1, st))
        else:
            host.copy_(g.reshape(-1))
            heads = flat[:, 0, ::self.ent // 2]
        return heads

    def find_duplicates(self, new_timestamps, min_match: int = 5) -> list[tuple[int, int]]:
        """Every rank returns the full list [(video_id, match_count)] in catalogue order."""
        while True:
            g = self.enqueue(new_timestamps, min_match)
            heads = self._read(g, self._host, self._host_np, self.world, self.rec_ints)
            hl = heads.tolist()                                   # [[n_hits, overflow], ...]: plain ints from here on
            n_max = max(h[0] for h in hl) if hl else 0
            if n_max > self.cap or any(h[1] for h in hl):
                self._regrow(n_max)
                continue
            flat: list = []
            host = self._host_np
            step = self.ent // 2
            for r, (n, _) in enumerate(hl):                       # shards in rank order = catalogue order
                if n:
                    flat += host[r, 1:1 + n, ::step].ravel().tolist()
            return list(zip(flat[0::2], flat[1::2]))

    # ---- 8 queries per pass --------------------------------------------------------------
    def enqueue_many(self, queries, min_match: int) -> torch.Tensor:
        """Up to 8 queries answered by ONE pass over every shard -> device tensor [world, 8, cap + 1, ent]."""
        if len(queries) > self.batch:
            raise ValueError("a batch holds at most %d queries" % self.batch)
        self._alloc_batch()
        if self.gather == "nccl":
            self.local.match_batch_async(queries, min_match, self.record_b)
            dist.all_gather_into_tensor(self.gathered_b, self.record_b, group=self.group)
            return self.gathered_b
        sy = self._sym_b
        s, epoch = sy.next()
        dst = sy.mc_record[s] if sy.multicast else sy.peer_record[s]
        self.local.match_batch_gather_async(queries, min_match, self.world, dst, sy.my_slots[s],
                                            self.batch * self.rec_ints, self.cap, epoch)
        return sy.views[s].view(self.world, self.batch, self.cap + 1, self.ent)

    def match_many(self, queries, min_match: int = 5) -> list[np.ndarray]:
        """[int32 [n_i, 2] (video_id, match_count) in catalogue order for every query], 8 queries per
        catalogue pass; every rank returns all of them.  Each query: <= 224 distinct values."""
        queries = list(queries)
        out: list = [None] * len(queries)
        g0 = 0
        while g0 < len(queries):
            group = queries[g0:g0 + self.batch]
            g = self.enqueue_many(group, min_match)
            heads = self._read(g, self._host_b, self._host_b_np, self.world * self.batch, self.rec_ints)
            heads = heads.reshape(self.world, self.batch, 2)[:, :len(group)]
            n_max = int(heads[..., 0].max()) if heads.size else 0
            if heads[..., 1].any() or n_max > self.cap:
                self._regrow(n_max)
                continue
            for b in range(len(group)):
                pairs, _, _ = merge_records(self._host_b_np[:, b, :n_max + 1, ::self.ent // 2], self.cap)
                out[g0 + b] = pairs
            g0 += len(group)
        return out

    def find_duplicates_many(self, queries, min_match: int = 5) -> list[list[tuple[int, int]]]:
        res = []
        for pairs in self.match_many(queries, min_match):
            flat = pairs.reshape(-1).tolist()
            res.append(list(zip(flat[0::2], flat[1::2])))
        return res


class ShardedFragmentCatalogue:
    """Fragment matching over a row-sharded catalogue: per-shard fixed-size records
    (int32 [3 * (cap + 1)]) gathered -- fused peer stores by default on GPUs, or one NCCL all-gather --
    merged in catalogue order, top-k on the host."""

    def __init__(self, ts, off, video_id, tick_hz: float = 1000.0, hit_capacity: int = 1 << 12, device=None,
                 local_factory: Callable | None = None, group=None, gather: str | None = None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.tick_hz = float(tick_hz)
        self.bounds = shard_bounds(off, self.world)
        lo, hi = self.bounds[self.rank]
        s_ts, s_off, s_vid = take_shard(np.asarray(ts), np.asarray(off), np.asarray(video_id), lo, hi)
        self._cuda = local_factory is None
        if local_factory is None:
            from .fragment import FragmentCatalogue
            dev = torch.cuda.current_device() if device is None else device
            self.device = torch.device("cuda", dev)
            self.local = FragmentCatalogue(s_ts, s_off, s_vid, tick_hz=tick_hz, device=dev, hit_capacity=hit_capacity)
        else:
            self.device = torch.device("cpu") if device is None else torch.device(device)
            self.local = local_factory(s_ts, s_off, s_vid)
        if gather is None:
            gather = "fused" if self._cuda else "nccl"
        if gather == "fused" and (not self._cuda or min(h - l for l, h in self.bounds) == 0):
            gather = "nccl"
        self.gather = gather
        self._alloc(hit_capacity)

    def _alloc(self, cap: int) -> None:
        self.cap = int(cap)
        self.rec_ints = (3 * (self.cap + 1) + 3) // 4 * 4      # slot stride: the record, padded to 16 bytes
        self._host = _host_mirror(self.world * self.rec_ints, self._cuda)
        self._sym = None
        if self.gather == "nccl":
            self.record = torch.zeros(self.rec_ints, dtype=torch.int32, device=self.device)
            self.gathered = torch.zeros(self.world * self.rec_ints, dtype=torch.int32, device=self.device)
        else:
            self._sym = _SymmetricRecords(self.rec_ints, self.device, self.group)

    def enqueue(self, clip_timestamps, min_match: int, **kw) -> torch.Tensor:
        """-> device tensor [world, 3 * (cap + 1)] holding every shard's record once the stream gets there."""
        if self.gather == "nccl":
            self.local.match_async(clip_timestamps, min_match, self.record[:3 * (self.cap + 1)], **kw)
            dist.all_gather_into_tensor(self.gathered, self.record, group=self.group)
            return self.gathered.view(self.world, self.rec_ints)
        sy = self._sym
        s, epoch = sy.next()
        self.local.match_gather_async(clip_timestamps, min_match, sy.peer_record[s], sy.peer_flag[s], sy.my_flags[s],
                                      self.cap, epoch, **kw)
        return sy.views[s].view(self.world, self.rec_ints)

    def find_fragments(self, clip_timestamps, min_match: int = 5, top_k: int | None = None, **kw):
        from .fragment import rank_fragments
        while True:
            g = self.enqueue(clip_timestamps, min_match, **kw)
            if self._cuda:
                # headers + an optimistic slice of pairs and offsets: two strided copies, one wait
                from ._lib import check, lib
                st = int(torch.cuda.current_stream(self.device).cuda_stream)
                first, pitch = min(self.cap, 256), 4 * self.rec_ints
                copy = lambda off, width, sync: check(lib().tvz_copy_records_to_host(   # noqa: E731
                    g.data_ptr(), self._host.data_ptr(), self.world, pitch, off, width, sync, st))
                copy(0, 8 * (first + 1), 0)
                copy(8 * (self.cap + 1) + 4, 4 * first, 1)
                n_max = int(self._host.numpy().reshape(self.world, self.rec_ints)[:, 0].max())
                if first < n_max <= self.cap:
                    copy(8 * (first + 1), 8 * (n_max - first), 0)
                    copy(8 * (self.cap + 1) + 4 + 4 * first, 4 * (n_max - first), 1)
            else:
                self._host.copy_(g.reshape(-1))
            h = self._host.numpy().reshape(self.world, self.rec_ints)
            pairs, overflow, needed = merge_records(h[:, : 2 * (self.cap + 1)].reshape(self.world, self.cap + 1, 2),
                                                    self.cap)
            if overflow:
                if self.gather == "fused":
                    self._sym.hdl.barrier()
                self._alloc(max(needed, 2 * self.cap))
                continue
            deltas = [h[r, 2 * (self.cap + 1) + 1: 2 * (self.cap + 1) + 1 + int(h[r, 0])] for r in range(self.world)]
            delta = np.concatenate(deltas) if deltas else np.zeros(0, np.int32)
            return rank_fragments(pairs[:, 0], pairs[:, 1], delta, self.tick_hz, top_k)

"""Multi-GPU matcher: the catalogue is sharded by row across the ranks of one box, every
rank streams its own shard, and the only data-path collective is one all-gather of the
fixed-size per-shard hit records.  One process per GPU, torch.distributed
for the plumbing (NCCL on GPUs; the host logic is backend-agnostic and is tested with gloo).

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


class ShardedCatalogue:
    """find_duplicates over a catalogue sharded across the ranks of the default process group.

    `local_factory(ts, off, video_id)` builds the per-rank matcher; it must offer
    ``match_async(q, min_match, out)`` filling an int32 [cap + 1, 2] record tensor on the
    rank's device.  The default is the CUDA `Catalogue` (no CPU path in the product; the
    gloo tests inject a CPU stand-in to exercise this host logic).

    gather="nccl"  : one all_gather_into_tensor of the fixed-size records per query.
    gather="fused" : the records live in symmetric (peer-mapped) memory and the query's kernel (its compaction phase)
                     itself stores each rank's record into every peer over NVLink and raises a
                     flag; no collective launch on the data path (tvz_catalog_match_gather_async).
    """

    def __init__(self, ts, off, video_id, hit_capacity: int = 1 << 15, device=None,
                 local_factory: Callable | None = None, group=None, gather: str = "nccl", presharded: bool = False):
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
        self.n_rows_local = hi - lo
        self.n_values_local = int(s_off[-1]) if s_off.size else 0
        self.epoch = 0
        self._alloc(hit_capacity)

    # ---- buffers -------------------------------------------------------------------------
    def _alloc(self, cap: int) -> None:
        self.cap = int(cap)
        rec = (self.cap + 1) * 2
        if self.gather == "nccl":
            self.record = torch.zeros((self.cap + 1, 2), dtype=torch.int32, device=self.device)
            self.gathered = torch.zeros((self.world * (self.cap + 1), 2), dtype=torch.int32, device=self.device)
            return
        # fused: symmetric buffer = 2 sets x [world records] followed by 2 sets x [world flags]
        import torch.distributed._symmetric_memory as symm
        n_rec = 2 * self.world * rec
        self.sym = symm.empty(n_rec + 2 * self.world + 32, dtype=torch.int32, device=self.device)
        self.sym.zero_()
        self.hdl = symm.rendezvous(self.sym, self.group if self.group is not None else dist.group.WORLD)
        torch.cuda.synchronize(self.device)
        self.hdl.barrier()
        base = np.asarray([int(p) for p in self.hdl.buffer_ptrs], np.uint64)
        self._peer_record, self._peer_flag, self._my_flags, self._views = [], [], [], []
        for s in range(2):
            self._peer_record.append(base + np.uint64(4 * (s * self.world * rec + self.rank * rec)))
            self._peer_flag.append(base + np.uint64(4 * (n_rec + s * self.world + self.rank)))
            self._my_flags.append(int(base[self.rank]) + 4 * (n_rec + s * self.world))
            self._views.append(self.sym[s * self.world * rec:(s + 1) * self.world * rec].view(self.world, self.cap + 1, 2))
        self.epoch = 0

    def enqueue(self, new_timestamps, min_match: int) -> torch.Tensor:
        """Local count + compaction + the gather, all on the current stream; returns the device
        tensor [world, cap + 1, 2] that holds every shard's record once the stream gets there."""
        if self.gather == "nccl":
            self.local.match_async(new_timestamps, min_match, self.record)
            dist.all_gather_into_tensor(self.gathered, self.record, group=self.group)
            return self.gathered.view(self.world, self.cap + 1, 2)
        self.epoch += 1
        s = self.epoch & 1
        self.local.match_gather_async(new_timestamps, min_match, self._peer_record[s], self._peer_flag[s],
                                      self._my_flags[s], self.cap, self.epoch)
        return self._views[s]

    def find_duplicates(self, new_timestamps, min_match: int = 5) -> list[tuple[int, int]]:
        """Every rank returns the full list [(video_id, match_count)] in catalogue order."""
        while True:
            g = self.enqueue(new_timestamps, min_match)
            heads = g[:, 0, :].cpu().numpy()                      # {n_hits, overflow} of every shard
            n_max = int(heads[:, 0].max()) if heads.size else 0
            if heads[:, 1].any() or n_max > self.cap:
                if self.gather == "fused":
                    self.hdl.barrier()                            # nobody still writes into the old buffers
                self._alloc(max(n_max, 2 * self.cap))             # same records on every rank -> same decision
                continue
            pairs, _, _ = merge_records(g[:, :n_max + 1, :].cpu().numpy(), self.cap)
            return list(zip(pairs[:, 0].tolist(), pairs[:, 1].tolist()))


class ShardedFragmentCatalogue:
    """Fragment matching over a row-sharded catalogue: per-shard fixed-size records
    (int32 [3 * (cap + 1)]) gathered with one all-gather, merged in catalogue order, top-k on the host."""

    def __init__(self, ts, off, video_id, tick_hz: float = 1000.0, hit_capacity: int = 1 << 12, device=None,
                 local_factory: Callable | None = None, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.tick_hz = float(tick_hz)
        self.bounds = shard_bounds(off, self.world)
        lo, hi = self.bounds[self.rank]
        s_ts, s_off, s_vid = take_shard(np.asarray(ts), np.asarray(off), np.asarray(video_id), lo, hi)
        if local_factory is None:
            from .fragment import FragmentCatalogue
            dev = torch.cuda.current_device() if device is None else device
            self.device = torch.device("cuda", dev)
            self.local = FragmentCatalogue(s_ts, s_off, s_vid, tick_hz=tick_hz, device=dev, hit_capacity=hit_capacity)
        else:
            self.device = torch.device("cpu") if device is None else torch.device(device)
            self.local = local_factory(s_ts, s_off, s_vid)
        self._alloc(hit_capacity)

    def _alloc(self, cap: int) -> None:
        self.cap = int(cap)
        self.record = torch.zeros(3 * (self.cap + 1), dtype=torch.int32, device=self.device)
        self.gathered = torch.zeros(self.world * 3 * (self.cap + 1), dtype=torch.int32, device=self.device)

    def enqueue(self, clip_timestamps, min_match: int, **kw) -> None:
        self.local.match_async(clip_timestamps, min_match, self.record, **kw)
        dist.all_gather_into_tensor(self.gathered, self.record, group=self.group)

    def find_fragments(self, clip_timestamps, min_match: int = 5, top_k: int | None = None, **kw):
        from .fragment import rank_fragments
        while True:
            self.enqueue(clip_timestamps, min_match, **kw)
            g = self.gathered.view(self.world, 3 * (self.cap + 1)).cpu().numpy()
            pairs, overflow, needed = merge_records(g[:, : 2 * (self.cap + 1)].reshape(self.world, self.cap + 1, 2),
                                                    self.cap)
            if overflow:
                self._alloc(max(needed, 2 * self.cap))
                continue
            deltas = [g[r, 2 * (self.cap + 1) + 1: 2 * (self.cap + 1) + 1 + int(g[r, 0])] for r in range(self.world)]
            delta = np.concatenate(deltas) if deltas else np.zeros(0, np.int32)
            return rank_fragments(pairs[:, 0], pairs[:, 1], delta, self.tick_hz, top_k)

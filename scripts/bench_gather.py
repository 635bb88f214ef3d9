"""The sharded matcher with the fused gather, device time per query / per 8-query pass (max over ranks), for build
variants selected by TVZ_LIB.  Run under torchrun on N GPUs:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/bench_gather.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from tvidz_b200 import synth
from tvidz_b200.dist import ShardedCatalogue

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
ts, off, vid = synth.synth_catalogue(1_000_000, seed=0)
q = ts[off[123_456]:off[123_457]].copy()
qs = [ts[off[r]:off[r + 1]].copy() for r in np.random.default_rng(7).integers(0, 1_000_000, 8)]
sc = ShardedCatalogue(ts, off, vid, hit_capacity=1 << 15, device=dev.index, gather="fused")
stream = torch.cuda.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps, cold=False):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    if not cold:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    else:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            flush.zero_()
            a.record(stream)
            fn()
            b.record(stream)
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e3


res = {}
for mm in (2, 5):
    res[f"one query mm{mm} b2b"] = timed(lambda: sc.enqueue(q, mm), 200)
    res[f"one query mm{mm} cold"] = timed(lambda: sc.enqueue(q, mm), 20, cold=True)
    res[f"8 queries mm{mm} b2b"] = timed(lambda: sc.enqueue_many(qs, mm), 100)
hits = len(sc.find_duplicates(q, 2))
many = sum(len(m) for m in sc.match_many(qs, 2))
if rank == 0:
    print(f"world {world} lib {os.path.basename(os.environ.get('TVZ_LIB', 'default'))}: " +
          "  ".join(f"{k} {v:6.1f} us" for k, v in res.items()) + f"  hits {hits} / {many}", flush=True)
dist.destroy_process_group()

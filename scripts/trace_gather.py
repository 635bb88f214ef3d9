"""Phase timestamps of the fused gather's last CTA on every rank (torchrun, N >= 2):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/trace_gather.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from tvidz_b200 import synth
from tvidz_b200._lib import check, lib
from tvidz_b200.dist import ShardedCatalogue

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ts, off, vid = synth.synth_catalogue(1_000_000, seed=0)
q = ts[off[123_456]:off[123_457]].copy()
sc = ShardedCatalogue(ts, off, vid, hit_capacity=max(4096, (1 << 15) // world), device=local, gather="fused")
for mm in (2, 5):
    for _ in range(5):
        sc.enqueue(q, mm)
    torch.cuda.synchronize()
    ws = sc.local._ws_async(0)
    trace = torch.zeros((sc.local.n_tiles, 16), dtype=torch.int64, device=dev)
    rows = []
    for it in range(6):
        trace.zero_()
        dist.barrier()
        torch.cuda.synchronize()
        check(lib().tvz_debug_tile_trace(ws.handle, trace.data_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):                                  # back to back: the ranks run in lock step; the trace keeps the last query
            sc.enqueue(q, mm)
        e1.record()
        torch.cuda.synchronize()
        check(lib().tvz_debug_tile_trace(ws.handle, None))
        t = trace.cpu().numpy().astype(np.float64)
        mhz = 1965.0
        n_cta = int((t[:, 1] > 0).sum())
        t = t[:n_cta]
        life = (t[:, 10] - t[:, 1]) / mhz                    # CTA start -> its own hits written and shipped
        recv = np.nonzero(t[:, 15] > 0)[0]                   # the receiver CTAs (the last n_peers of the grid)
        wait = (t[recv, 15] - t[recv, 11]) / mhz
        start_ns = t[:, 0] - t[:, 0].min()
        rows.append((e0.elapsed_time(e1) * 1e2, np.median(life), life.max(), wait.min(), wait.max(),
                     ((t[recv, 15] - t[recv, 1]) / mhz).max(), start_ns.max() / 1e3))
    r = np.median(np.asarray(rows[1:]), axis=0)
    print(f"rank {rank} mm={mm}: event {r[0]:.1f} us | CTA start -> shipped: median {r[1]:.1f} max {r[2]:.1f} us | receivers wait for "
          f"their peer's record: {r[3]:.1f} .. {r[4]:.1f} us | slowest receiver's lifetime {r[5]:.1f} us | CTA start spread {r[6]:.1f} us", flush=True)
dist.destroy_process_group()

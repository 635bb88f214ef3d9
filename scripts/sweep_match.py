"""Compare the count-kernel variants on the bench catalogue (1M rows): kernel time via the
library's own CUDA events, whole query via torch events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tvidz_b200 import _lib, synth
from tvidz_b200.catalog import Catalogue

dev = torch.device("cuda:0")
n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ts, off, vid = synth.synth_catalogue(n_rows, seed=0)
cat = Catalogue(ts, off, vid, hit_capacity=1 << 16)
q = ts[off[n_rows // 8]:off[n_rows // 8 + 1]].copy()
if len(sys.argv) > 2 and sys.argv[2] == "nohit":
    q = q + 1e9          # same number of keys, no value of the catalogue matches
rec = torch.zeros(((1 << 16) + 1, 2), dtype=torch.int32, device=dev)
lib = _lib.lib()
ref = None
for variant, name in ((0, "LDG.256 rolling"),):
    for _ in range(3):
        cat.match_async(q, 2, rec)
    torch.cuda.synchronize()
    cat.debug_count_kernel_ms(True)
    ks = []
    for _ in range(20):
        cat.match_async(q, 2, rec)
        ks.append(cat.debug_count_kernel_ms())
    cat.debug_count_kernel_ms(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        cat.match_async(q, 2, rec)
    e1.record()
    torch.cuda.synchronize()
    whole = e0.elapsed_time(e1) / 50
    r = rec.cpu().numpy().copy()
    if ref is None:
        ref = r
    same = bool(np.array_equal(ref[: ref[0, 0] + 1], r[: r[0, 0] + 1]))
    gbs = cat.algo_bytes / (np.mean(ks) * 1e-3) / 1e9
    print(f"{name:16s} count kernel {np.mean(ks)*1e3:7.1f} us (min {np.min(ks)*1e3:6.1f})  {gbs:7.1f} GB/s  "
          f"whole query {whole*1e3:7.1f} us  hits {int(r[0,0])} same={same}", flush=True)

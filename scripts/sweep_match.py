"""Compare launch shapes / layouts of the count kernel on the bench catalogue (1M rows).

  python scripts/sweep_match.py build            # here (no GPU): one library per variant under build/variants/
  python scripts/sweep_match.py run [rows]       # on the GPU box: runs every variant (TVZ_LIB hook of _lib.py)
  python scripts/sweep_match.py [rows] [nohit]   # times the library that is loaded: kernel time via the
                                                 # library's own CUDA events, whole query via torch events
"""
import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR_DIR = os.path.join(ROOT, "build", "variants")
VARIANTS = {            # name: (threads, units, extra defines)
    "fp_t512_u2": (512, 2, []),
    "fp_t512_u3": (512, 3, []),
    "fp_t512_u4": (512, 4, []),
    "fp_t768_b2_u2": (768, 2, ["-DTVZ_FP_MINB=2", "-DTVZ_FP_QUEUE=96"]),
}
if len(sys.argv) > 1 and sys.argv[1] == "build":
    from tvidz_b200 import build as b
    os.makedirs(VAR_DIR, exist_ok=True)
    for name, (t, u, extra) in VARIANTS.items():
        out = os.path.join(VAR_DIR, f"libtvz_{name}.so")
        cmd = [b.NVCC] + b.FLAGS + [f"-DTVZ_FP_THREADS={t}", f"-DTVZ_FP_UNITS={u}"] + extra + ["-Xptxas", "-v", "-o", out] + \
            [os.path.join(b.CSRC, x) for x in b.SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        lines = r.stderr.splitlines()
        info = [lines[i + 1].strip().split("Function properties for ")[-1][:0] + lines[i + 2].strip() + " | " + lines[i + 3].strip()
                for i, ln in enumerate(lines) if "match_count_kernelILb1" in ln and "Compiling" in ln]
        print(name, "FAILED " + r.stderr[-300:] if r.returncode else info)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "run":
    for name in VARIANTS:
        lib = os.path.join(VAR_DIR, f"libtvz_{name}.so")
        if os.path.exists(lib):
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[2:], capture_output=True, text=True,
                               env=dict(os.environ, TVZ_LIB=lib))
            print(f"{name:22s}", (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)
    sys.exit(0)
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

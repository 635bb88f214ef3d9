"""Sweep the SAD ring geometry on the bench workload (64 x 32 x 1080p): prints GB/s per variant."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tvidz_b200 import _lib, scene, synth

dev = torch.device("cuda:0")
S, F, H, W = 64, 32, 1080, 1920
if len(sys.argv) > 1 and sys.argv[1] == "4k":
    S, F, H, W = 1, 256, 2160, 3840
frames = torch.randint(0, 256, (S, F, H, W), dtype=torch.uint8, device=dev)
lib = _lib.lib()
names = {0: "6x32K", 1: "4x16K", 2: "3x32K", 3: "8x16K", 4: "12x16K", 5: "4x32K"}
ref = None
for variant in range(6):
    for ctas in (1, 2, 3):
        for upc in (16,):
            _lib.check(lib.tvz_debug_sad_tuning(variant, ctas, upc, 16))
            for _ in range(3):
                sad = scene.sad_luma(frames)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(10):
                sad = scene.sad_luma(frames)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 10
            if ref is None:
                ref = sad.clone()
            ok = bool(torch.equal(ref, sad))
            gbs = S * (F - 1) * H * W / ms / 1e6
            print(f"variant {variant} ({names[variant]}) ctas/SM<={ctas}: {ms*1e3:8.1f} us  {gbs:7.1f} GB/s algorithmic  ok={ok}", flush=True)

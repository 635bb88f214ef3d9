"""Which libnvcuvid copies exist on this box, and what each one answers (run on the GPU box)."""
import glob, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cands = sorted(set(glob.glob("/usr/lib/libnvcuvid*") + glob.glob("/usr/local/nvidia/lib*/libnvcuvid*") +
                   glob.glob("/usr/lib/x86_64-linux-gnu/libnvcuvid*")))
for c in cands:
    print(c, "->", os.path.realpath(c), os.path.getsize(os.path.realpath(c)))
print("NVIDIA_DRIVER_CAPABILITIES =", os.environ.get("NVIDIA_DRIVER_CAPABILITIES"))
print(subprocess.run("ls -la /dev/nvidia* /dev/nvidia-caps 2>&1 | head -20; cat /proc/driver/nvidia/version 2>&1 | head -3; ldconfig -p | grep -E 'libcuda.so|nvcuvid'",
                     shell=True, capture_output=True, text=True).stdout)
code = """
import sys; sys.path.insert(0, %r)
import torch; torch.cuda.init(); torch.zeros(1, device='cuda')
from tvidz_b200 import nvdec, _lib
print('available', nvdec.available(), 'loaded', _lib.lib().tvz_nvdec_library())
print(nvdec.all_caps())
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for c in [None] + [x for x in cands if x.endswith(".so.1") or x[-1].isdigit()]:
    env = dict(os.environ)
    if c:
        env["TVZ_NVCUVID"] = c
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("== TVZ_NVCUVID =", c, "\n", r.stdout[-1500:], r.stderr[-500:])

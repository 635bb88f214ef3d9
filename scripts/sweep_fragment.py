"""Launch-shape sweep of the streaming fragment kernel.

  python scripts/sweep_fragment.py build     # here (no GPU): one library per variant under build/variants/
  python scripts/sweep_fragment.py run       # on the GPU box: times configs[4] with each of them

Variants are the same sources compiled with -DTVZ_FS_THREADS / _MINB / _UNITS (threads per CTA,
CTAs per SM the register budget is sized for, 256-bit loads in flight per thread) and loaded through
the TVZ_LIB hook of tvidz_b200/_lib.py.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR_DIR = os.path.join(ROOT, "build", "variants")
VARIANTS = {            # name: (threads, min blocks/SM, units, extra defines)
    "t256_b3_u4": (256, 3, 4, []),
    "t256_b2_u6": (256, 2, 6, []),
    "t256_b2_u8": (256, 2, 8, []),
    "t128_b4_u6": (128, 4, 6, []),
    "t160_b4_u5": (160, 4, 5, []),
    "t192_b3_u6": (192, 3, 6, []),
    "t192_b3_u5": (192, 3, 5, []),
    "t256_b2_u6_nopark": (256, 2, 6, ["-DTVZ_FS_NOPARK"]),
}


def build():
    from tvidz_b200 import build as b
    os.makedirs(VAR_DIR, exist_ok=True)
    for name, (t, mb, u, extra) in VARIANTS.items():
        out = os.path.join(VAR_DIR, f"libtvz_{name}.so")
        cmd = [b.NVCC] + b.FLAGS + [f"-DTVZ_FS_THREADS={t}", f"-DTVZ_FS_MINB={mb}", f"-DTVZ_FS_UNITS={u}"] + extra + \
            ["-Xptxas", "-v", "-o", out] + [os.path.join(b.CSRC, s) for s in b.SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            print(name, "FAILED", r.stderr[-400:])
            continue
        lines = r.stderr.splitlines()
        info = [lines[i + 1].strip() + " | " + lines[i + 2].strip() for i, ln in enumerate(lines)
                if "fragment_stream_kernelILi2" in ln and "Compiling" in ln]
        print(name, info)


def run():
    n = sys.argv[2] if len(sys.argv) > 2 else "100000"
    for name in VARIANTS:
        lib = os.path.join(VAR_DIR, f"libtvz_{name}.so")
        if not os.path.exists(lib):
            continue
        env = dict(os.environ, TVZ_LIB=lib, TVZ_FRAG_L2_AHEAD=os.environ.get("TVZ_FRAG_L2_AHEAD", "0"))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "prof_fragment.py"), n, "50"],
                           capture_output=True, text=True, env=env)
        print(f"{name:22s}", (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()

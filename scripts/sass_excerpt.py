"""Per-kernel counts of the SASS mnemonics that show what the kernels are built from (run anywhere: cuobjdump
reads the in-tree library).   python scripts/sass_excerpt.py > profiles/r02_sass_excerpt.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "tvidz_b200", "libtvidz_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = [("UBLKCP", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("VABSDIFF4", r"\bVABSDIFF4"), ("REDUX", r"\bREDUX"),
       ("LDG.E.*.256", r"\bLDG\.E[.\w]*\.256"), ("LDG.E.128", r"\bLDG\.E[.\w]*\.128"), ("LDG STRONG.GPU/SYS (relaxed polls)", r"\bLDG\.E[.\w]*STRONG"),
       ("LDS.U8", r"\bLDS\.U8"), ("LDS.128", r"\bLDS\.128"), ("STS.128", r"\bSTS\.128"), ("ATOMS (shared atomics)", r"\bATOMS"),
       ("RED/ATOMG (global)", r"\b(RED|ATOMG)\b"), ("VOTE", r"\bVOTE"), ("SHFL", r"\bSHFL"), ("MEMBAR", r"\bMEMBAR"),
       ("CCTL", r"\bCCTL"), ("ACQBULK/griddepcontrol", r"\bACQBULK|\bDEPBAR"), ("BAR.SYNC", r"\bBAR\.SYNC"), ("NANOSLEEP", r"\bNANOSLEEP")]
print("SASS mnemonic counts per kernel, sm_100a, from `cuobjdump -sass tvidz_b200/libtvidz_b200.so`")
print("(UBLKCP = TMA 1-D bulk copy global->shared; SYNCS = mbarrier ops; VABSDIFF4 = 4-byte SAD with accumulate;")
print(" LDG...256 = 256-bit global loads; UTMALDG is absent on purpose: the bulk copies are 1-D, no tensor map needed)\n")
arch = re.search(r"arch = (sm_\w+)", out)
print("arch:", arch.group(1) if arch else "?")
for m in re.finditer(r"Function : (\S+)\n(.*?)(?=\n\s*Function : |\Z)", out, re.S):
    name, body = m.group(1), m.group(2)
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(anonymous namespace\)::|tvz::|<unnamed>::", "", short).split("(")[0]
    n_inst = len(re.findall(r"^\s+/\*[0-9a-f]{4}\*/", body, re.M))
    counts = [(label, len(re.findall(p, body))) for label, p in PAT]
    print(f"== {short}   ({n_inst} instructions)")
    print("   " + ", ".join(f"{l}: {c}" for l, c in counts if c))

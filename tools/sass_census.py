"""SASS opcode census of the built objects (evidence that the contraction kernels are Blackwell-native: tcgen05.mma = UTC*MMA,
tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG/UTMASTG/UTMAPF/UBLKPF; legacy mma.sync = HMMA).
Usage (no GPU needed):  python tools/sass_census.py > profiles/sass_census.txt     (after `make -C tapclip_b200/csrc`)"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKPF", "UTCBAR", "HMMA", "MUFU.EX2", "RED", "ATOMG"]
objs = sorted(glob.glob(os.path.join(ROOT, "build", "obj", "*.o")))
if not objs:
    sys.exit("no objects under build/obj: run `make -C tapclip_b200/csrc` first")
print("# SASS opcode counts per object (cuobjdump -sass, sm_100a); regenerate with tools/sass_census.py")
print("# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store,")
print("# UTMAPF/UBLKPF = TMA / bulk L2 prefetch, HMMA = legacy mma.sync")
print(f"{'object':22s} " + " ".join(f"{o:>12s}" for o in OPS) + "   kernels")
for o in objs:
    sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    counts = collections.Counter()
    kernels = len(re.findall(r"^\s*Function :", sass, flags=re.M))
    for line in sass.splitlines():
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for name in OPS:
            if name == "UTCHMMA":
                if op.startswith("UTCHMMA") and ".2CTA" not in op:
                    counts[name] += 1
            elif name == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    counts[name] += 1
            elif op.startswith(name):
                counts[name] += 1
    if any(counts.values()):
        print(f"{os.path.basename(o):22s} " + " ".join(f"{counts[n]:12d}" for n in OPS) + f"   {kernels}")

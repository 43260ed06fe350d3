"""Per-kernel counts of the Blackwell instructions in the shipped library (cuobjdump -sass):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit),
UTMALDG / UTMASTG (tensor-map TMA), SYNCS (mbarrier), LDGSTS (cp.async).   python tools/sass_counts.py > profiles/r2_sass_counts.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "hy2dl_b200", "csrc", "libhy2dl_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "MUFU", "FFMA"]
counts, name = {}, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|hy2dl::", "", name).split("(")[0]
        counts[name] = dict.fromkeys(ops, 0)
        continue
    if name:
        for op in ops:
            if re.search(r"\b" + op + r"\b|\b" + op + r"\.", line):
                counts[name][op] += 1
print("library:", os.path.relpath(lib, ROOT), " (nvcc -gencode arch=compute_100a,code=sm_100a; cuobjdump -sass)")
print(f"{'kernel':58s}" + "".join(f"{o:>9s}" for o in ops))
tot = dict.fromkeys(ops, 0)
for k in sorted(counts, key=lambda k: -counts[k]["UTCHMMA"]):
    c = counts[k]
    print(f"{k[:58]:58s}" + "".join(f"{c[o]:9d}" for o in ops))
    for o in ops:
        tot[o] += c[o]
print(f"{'TOTAL':58s}" + "".join(f"{tot[o]:9d}" for o in ops))
print("\nNo UTMALDG / UTMASTG: operands move by 1-D bulk copies (UBLKCP, the TMA engine without a tensor map) and cp.async; see DESIGN.md"
      " section 3.6 for why no tensor map is used.")

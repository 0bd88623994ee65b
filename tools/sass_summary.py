"""Per-kernel SASS opcode summary of libfav_b200.so (runs without a GPU): which kernels really use tcgen05 / TMEM / TMA.
  python tools/sass_summary.py > profiles/sass_summary.txt
Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG =
TMA tensor load / store, UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, LDGSTS = cp.async, HMMA = legacy mma.sync."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "failure-aware-vision_b200", "csrc", "libfav_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCCP", "SYNCS", "ELECT", "LDGSTS", "HMMA",
       "MUFU", "ATOMS", "RED", "ATOMG", "LDL", "STL"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["_total"] += 1
        if op in OPS:
            counts[kern][op] += 1
            total[op] += 1
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  ({len(counts)} kernels); columns: instruction count per opcode, blank = 0")
print(f"# library totals: " + ", ".join(f"{k} x{v}" for k, v in sorted(total.items(), key=lambda kv: -kv[1])))
print(f"{'kernel':58s} {'instrs':>7s} " + " ".join(f"{o:>7s}" for o in OPS))
for k, c in counts.items():
    print(f"{k[:58]:58s} {c['_total']:7d} " + " ".join(f"{(c[o] or ''):>7}" for o in OPS))

"""Opcode histogram per kernel of libnerfb200.so (cuobjdump -sass), for the mnemonics that prove the sm_100a features in use:
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UTMALDG (TMA tensor copy), UBLKCP (bulk copy),
SYNCS (mbarrier), FHFMA (mixed-precision fma), F2FP (16-bit packs), plus totals.  Usage: python tools/sass_histogram.py [lib]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "nerf_b200/libnerfb200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UBLKCP", "SYNCS", "FHFMA", "F2FP", "MUFU", "LDGSTS", "STG", "LDG", "STS", "LDS", "SHFL", "BAR"]
kern, hist, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for k in KEYS:
            if op.startswith(k):
                hist[kern][k + (".2CTA" if ".2CTA" in op else "")] += 1
print(f"SASS opcode histogram of {lib} (sm_100a); columns = instruction counts in the kernel's code")
for k, h in hist.items():
    if total[k] < 40:
        continue
    print(f"\n{k}  [{total[k]} instructions]")
    print("   " + "  ".join(f"{op}={n}" for op, n in sorted(h.items())))

#!/usr/bin/env python3
"""SASS evidence without a GPU: per kernel of liborbx.so, the instruction count and the mnemonics that show which hardware
paths the kernel uses (TMA box loads = UTMALDG, mbarrier = SYNCS, integer SIMD = IDP.4A / IDP.2A / VABSDIFF4 / VIADD.16x2,
warp collectives = SHFL / VOTE / REDUX, tensor cores = none by design).
usage: sass_histogram.py LIB.so > profiles/<round>/sass_mnemonics.txt"""
import collections, re, subprocess, sys

lib = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ("UTMALDG", "UBLKCP", "SYNCS", "IDP.4A", "IDP.2A", "VABSDIFF4", "VIADD.16x2", "VIMNMX", "PRMT", "SHFL", "VOTE", "REDUX", "ATOMS", "ATOMG", "RED.",
       "LDG", "STG", "LDS", "STS", "IMAD", "DMUL", "DADD", "FMUL", "FADD", "F2I", "HMMA", "UTC", "LDTM", "BAR.SYNC")
fn, per = None, {}
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[fn] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if m and fn:
        op = m.group(1)
        per[fn]["_total"] += 1
        for k in KEY:
            if op.startswith(k) or (k.endswith(".") is False and k in op and k in ("IDP.4A", "IDP.2A", "VIADD.16x2", "BAR.SYNC")):
                per[fn][k] += 1
                break
print("# static SASS instruction counts per kernel (cuobjdump -sass %s)" % lib)
for fn in sorted(per):
    c = per[fn]
    if c["_total"] < 20:
        continue
    print(f"\n{fn}: {c['_total']} instructions")
    print("   " + "  ".join(f"{k}={c[k]}" for k in KEY if c[k]))
    tc = c["HMMA"] + c["UTC"] + c["LDTM"]
    print(f"   tensor-core instructions: {tc} (none expected: the path has no dense contraction)")

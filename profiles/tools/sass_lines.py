#!/usr/bin/env python3
"""Join an ncu SASS source page with nvdisasm line info: executed warp-instructions per CUDA source line.
usage: sass_lines.py REPORT.ncu-rep KERNEL_REGEX LIB.so SOURCE_BASENAME [units-divisor] [min-per-unit]"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, kre, lib, srcname = sys.argv[1:5]
div = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
minv = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr = rows[1]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[2:] if len(r) > ie and r[0].startswith("0x")]
base = int(body[0][ia], 16)
execd = {int(r[ia], 16) - base: (int(r[ie]), int(r[isamp])) for r in body}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
mangled_hint = re.sub(r"[^A-Za-z0-9_]", "", kname.split("(")[0].split("::")[-1].split("<")[0])
per_line = collections.Counter(); samp_line = collections.Counter()
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line, cur_file, active = None, None, None, False
    candidates = []
    for ln in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1); active = mangled_hint in cur_fn and all(f"Li{a}E" in cur_fn for a in re.findall(r"\(int\)(\d+)", kname)); 
            if active: candidates.append(cur_fn)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_file, cur_line = os.path.basename(m.group(1)), int(m.group(2)); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m and active:
            off = int(m.group(1), 16)
            if off in execd and cur_file == srcname:
                per_line[cur_line] += execd[off][0]; samp_line[cur_line] += execd[off][1]
            elif off in execd:
                per_line[(cur_file, cur_line)] += execd[off][0]; samp_line[(cur_file, cur_line)] += execd[off][1]
    if candidates and len(set(candidates)) > 1 and "<" in kname:
        pass
src = open(os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc", srcname)).read().splitlines() if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc", srcname)) else []
tot = sum(per_line.values())
print(f"kernel {kname}: {tot/div:.1f} warp-instr per unit")
for k in sorted(per_line, key=lambda x: (0, x) if isinstance(x, int) else (1, x[1])):
    v = per_line[k] / div
    if v >= minv:
        txt = src[k - 1].strip()[:100] if isinstance(k, int) and k - 1 < len(src) else str(k)
        print(f"{k!s:>8} {v:9.1f} {samp_line[k]:7d}  {txt}")

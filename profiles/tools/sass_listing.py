#!/usr/bin/env python3
"""Plain SASS listings of the hot kernels (addresses + mnemonics, encodings stripped), one file per kernel, so that the loops the design text reasons
about can be audited: FAST pass A (the compass loop), the exact-score network, the resize row step, the blur pair step, the rBRIEF rotation loop.
usage: sass_listing.py LIB.so OUTDIR [ncu-report]   — with a report, every line also carries the executed warp-instruction count of the capture."""
import csv, io, os, re, subprocess, sys

lib, outdir = sys.argv[1], sys.argv[2]
rep = sys.argv[3] if len(sys.argv) > 3 else None
KERNELS = {"k_fast_cells<64>": r"k_fast_cellsILi64E", "k_resize_s12": r"k_resize_s12", "k_blur<false>": r"k_blurILb0E", "k_describe": r"k_describeE", "k_octree<128>": r"k_octreeILi128E",
           "k_bow_descend": r"k_bow_descend", "k_search_by_projection<false>": r"k_search_by_projectionILb0E"}
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
os.makedirs(outdir, exist_ok=True)
for name, pat in KERNELS.items():
    blk = next((b for b in blocks if re.search(pat, b.split("\n", 1)[0])), None)
    if blk is None:
        continue
    counts = {}
    if rep:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + name.split("<")[0]], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) > 2:
            hdr = rows[1]
            ia, ie = hdr.index("Address"), hdr.index("Instructions Executed")
            body = [r for r in rows[2:] if len(r) > ie and r[0].startswith("0x")]
            body = body[:len(body) // 2] if len(body) > 1 and body[0][1] == body[len(body) // 2][1] else body   # (the page lists the function twice)
            base = int(body[0][ia], 16)
            counts = {int(r[ia], 16) - base: int(r[ie]) for r in body}
    lines = []
    for ln in blk.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            off = int(m.group(1), 16)
            c = f"{counts[off]:>12d}  " if off in counts else ""
            lines.append(f"{c}/*{m.group(1)}*/  {m.group(2).strip()}")
    fn = os.path.join(outdir, "sass_" + re.sub(r"[^A-Za-z0-9_]", "_", name) + ".txt")
    with open(fn, "w") as f:
        f.write(f"# {name}: {len(lines)} instructions" + (" ; first column = executed warp-instructions in " + os.path.basename(rep) if counts else "") + "\n")
        f.write("\n".join(lines) + "\n")
    print(fn, len(lines))

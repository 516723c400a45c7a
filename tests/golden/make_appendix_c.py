#!/usr/bin/env python3
"""Transcribes the six known-answer cases of SURVEY.md Appendix C into tests/golden/appendix_c.json.

Provenance of the numbers: produced during the survey by an emulation whose control flow was checked
against the verbatim reference lines (ORBextractor.cc:45-120, 454-753) and whose pixel primitives were cv2 4.13
(resize / FastFeatureDetector / GaussianBlur), i.e. they are independent of oracle/orb_oracle.cpp.
usage: make_appendix_c.py SURVEY.md > appendix_c.json
"""
import json, re, sys

def rows(text, header_start):
    lines = text.splitlines()
    i = next(k for k, l in enumerate(lines) if l.startswith(header_start))
    out = []
    for l in lines[i + 2:]:
        if not l.startswith("|"):
            break
        out.append([c.strip() for c in l.strip().strip("|").split("|")])
    return out

def ints(s):
    return [int(v) for v in re.findall(r"-?\d+", s)]

def main():
    text = open(sys.argv[1]).read()
    cases = {}
    for r in rows(text, "| case | W×H | seed |"):
        name = r[0].strip("`")
        w, h = ints(r[1])
        a = re.findall(r"[\d.]+", r[3])
        n, ret = ints(r[9])
        cases[name] = dict(width=w, height=h, seed=int(r[2]), nfeatures=int(a[0]), scale=float(a[1]), nlevels=int(a[2]), ini_th=int(a[3]),
                           min_th=int(a[4]), lapping=ints(r[4]), image_sha256_16=re.search(r"`([0-9a-f]{16})", r[5]).group(1),
                           first_pixels=ints(r[5].split("/")[1]), candidates=ints(r[6]), fallback_cells=ints(r[7]), kept=ints(r[8]), n=n, mono_index=ret)
    for r in rows(text, "| case | kp_xyor sha256 |"):
        c = cases[r[0].strip("`")]
        c["kp_xyor_sha256"], c["kp_full_sha256"], c["desc_sha256"] = (x.strip("`") for x in r[1:4])
        m = re.match(r"\((\d+), (\d+)\), size (\d+), angle ([\d.]+), resp (\d+), oct (\d+)", r[4])
        c["first_keypoint"] = dict(x=int(m.group(1)), y=int(m.group(2)), size=int(m.group(3)), angle=float(m.group(4)), response=int(m.group(5)), octave=int(m.group(6)))
        c["first_descriptor_hex"] = r[5].strip("`")
    for r in rows(text, "| case | L0 … L7 |"):
        cases[r[0].strip("`")]["level_sha256_16"] = re.findall(r"`([0-9a-f]{16})`", r[1])
    json.dump(cases, sys.stdout, indent=1)

main()

"""The Frame / ORBmatcher seams of INTEGRATION.md 2b, 2d, 2e compiled as C++ against the C ABI and the drop-in ORBextractor.h
(tests/frame_seams_demo.cpp), the way a maintainer would bind them: it must compile and link without a GPU, and on a GPU its
grid, frame-to-frame and local-map search results must equal the CPU oracle's."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tc2li_slam_b200", "build", "frame_seams_demo")


def build_demo():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.join(ROOT, "tc2li_slam_b200", "lib")
    cmd = ["g++", "-O2", "-std=c++14", "-pthread", "-o", EXE, os.path.join(ROOT, "tests", "frame_seams_demo.cpp"),
           "-I" + os.path.join(ROOT, "tc2li_slam_b200", "compat"), "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "cvstub"),
           "-L" + libdir, "-lorbx", "-L" + os.path.join(ROOT, "oracle"), "-lorb_oracle",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    subprocess.check_call(cmd)


def test_frame_seams_compile_and_link():
    build_demo()
    assert subprocess.run([EXE, "compile-only"]).returncode == 0


@pytest.mark.gpu
def test_frame_seams_match_oracle():
    build_demo()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stderr)
    lines = dict((l.split()[0], [int(v) for v in l.split()[1:]]) for l in out.stdout.strip().splitlines())
    W, H, mb, mbf = 1241, 376, 0.5372, 387.5744
    left, right = po.OracleExtractor(2000, 1.2, 8, 20, 7), po.OracleExtractor(2000, 1.2, 8, 20, 7)
    _, k, d = left(po.synth_v1(W, H, 77))
    right(po.synth_v1(W, H, 77, 16, 0x5bd1e995))
    ur, _ = po.stereo_match(left, right, mb, mbf)
    bounds = (0.0, float(W), 0.0, float(H))
    start, items = po.assign_grid(k, bounds)
    gsum = 0
    for c in range(64 * 48):
        lst = items[start[c]:start[c + 1]].astype(np.int64)
        gsum += int(((c + 1) * (lst + 1) * (np.arange(len(lst)) + 1)).sum())
    assert lines["grid"] == [int(start[-1]), gsum]
    sf = left.tables()["scale_factors"]
    n = len(k)
    q = np.zeros(n, po.PROJ_QUERY_DTYPE)
    q["u"], q["v"], q["angle"] = k["x"], k["y"], k["angle"]
    q["radius"] = np.float32(15.0) * sf[k["octave"]]
    q["ur"] = k["x"] - np.float32(1.0)
    q["min_level"], q["max_level"], q["blocks"] = k["octave"] - 1, k["octave"] + 1, 1
    q["desc"] = d
    w = np.arange(1, n + 1, dtype=np.int64)
    m, nm = po.search_by_projection(k, d, ur, (start, items), bounds, q, True)
    assert lines["frame"] == [nm, int(((m.astype(np.int64) + 2) * w).sum())]
    q["radius"] = np.float32(4.0) * np.float32(3.0) * sf[k["octave"]]
    q["min_level"], q["max_level"] = k["octave"] - 1, k["octave"]
    taken = np.zeros(n, np.uint8)
    taken[::5] = 1
    m, nm = po.search_local_points(k, d, ur, taken, (start, items), bounds, q, 0.8)
    assert lines["local"] == [nm, int(((m.astype(np.int64) + 2) * w).sum())]
    assert nm > 500
    # Frame::ComputeBoW against the same synthetic vocabulary (written in ORBvoc.txt's format for the oracle)
    kk, n_nodes = 4, 85
    text = ["4 3 0 0"]
    for i in range(1, n_nodes):
        desc = " ".join(str((i * 131 + j * 29 + (i * j) % 7) & 255) for j in range(32))
        wgt = 0.5 + (i % 13) * 0.37 if i >= 21 else 0.0
        text.append(f"{(i - 1) // kk} {1 if i >= 21 else 0} {desc} {wgt!r}")
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write("\n".join(text))
    words, values, nodes, start, feats = po.OracleVocabulary(f.name).transform(d, 2)
    os.unlink(f.name)
    vsum = int((values.view(np.uint64) >> np.uint64(20)).astype(object).sum())
    fsum = 0
    for j in range(len(nodes)):
        lst = feats[start[j]:start[j + 1]].astype(object)
        fsum += int(sum((int(nodes[j]) + 1) * (int(x) + 1) * (t + 1) for t, x in enumerate(lst)))
    assert lines["bow"] == [len(words), len(nodes), int((words.astype(np.int64) + 1).sum()), vsum, fsum]
    assert len(words) > 30

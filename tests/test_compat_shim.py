"""The drop-in C++ header tc2li_slam_b200/compat/ORBextractor.h: compiles against OpenCV-shaped headers (oracle/cvstub here,
since this image has no OpenCV SDK) and links liborbx.so; on a GPU it reproduces the oracle when driven like Frame::Frame
drives the reference (two extractor instances on two host threads, Frame.cc:139-142)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tc2li_slam_b200", "build", "compat_demo")


def build_demo():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.join(ROOT, "tc2li_slam_b200", "lib")
    cmd = ["g++", "-O2", "-std=c++14", "-pthread", "-o", EXE, os.path.join(ROOT, "tests", "compat_demo.cpp"),
           "-I" + os.path.join(ROOT, "tc2li_slam_b200", "compat"), "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "cvstub"),
           "-L" + libdir, "-lorbx", "-L" + os.path.join(ROOT, "oracle"), "-lorb_oracle",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    subprocess.check_call(cmd)


def test_shim_compiles_and_links():
    build_demo()
    assert subprocess.run([EXE, "compile-only"]).returncode == 0


@pytest.mark.gpu
def test_shim_two_threads_match_oracle():
    build_demo()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    ora = po.OracleExtractor(2000, 1.2, 8, 20, 7)
    for s, (xoff, nx) in enumerate(((0, 0), (16, 0x5bd1e995))):
        om, ok, od = ora(po.synth_v1(1241, 376, 42, xoff, nx))
        n, mono, dsum, ksum, w7, h7 = lines[s].split()
        want = int((od.astype(np.uint64) * (np.arange(od.size, dtype=np.uint64).reshape(od.shape) + 1)).sum())
        assert int(n) == len(ok) and int(mono) == om and int(dsum) == want
        ks = float((ok["x"].astype(np.float64) + 2 * ok["y"].astype(np.float64) + 3 * ok["angle"].astype(np.float64) + ok["octave"]).sum())
        assert abs(float(ksum) - ks) < 1e-3 * max(1.0, abs(ks)) * 1e-3
        assert (int(w7), int(h7)) == (346, 105)
    assert lines[2].startswith("levels 8")

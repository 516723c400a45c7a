#!/usr/bin/env python3
"""Per-stage device time for an arbitrary shape: stages_shape.py W H N_IMAGES NFEATURES"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import tc2li_slam_b200 as T
w, h, n, nf = (int(x) for x in sys.argv[1:5])
b = T.BatchExtractor(w, h, n, nf, 1.2, 8, 20, 7)
b.synth(7, n, stereo=False)
for _ in range(3):
    b.run(n)
b.check()
b.enable_timing(True)
acc = {}
for _ in range(3):
    b.run(n)
    for k, v in b.stage_ms().items():
        acc[k] = acc.get(k, 0) + v / 3
print(w, h, n, {k: round(v, 3) for k, v in acc.items()}, "total", round(sum(acc.values()), 3))
cands = [len(b.candidates(0, l)) for l in range(8)]
print("candidates image 0:", cands)

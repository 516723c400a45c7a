"""Times (CUDA events) and, under ncu, profiles the two bag-of-words kernels on a resident batch: 1024 KITTI-shaped images, synthetic k = 10 / L = 6 vocabulary."""
import sys
import numpy as np
sys.path.insert(0, ".")
import tc2li_slam_b200 as T
from tc2li_slam_b200.orbx import Event

n = 1024
bx = T.BatchExtractor(1241, 376, n, 2000, 1.2, 8, 20, 7)
bx.synth(1, n, True)
bx.run(n)
k, L = 10, 6
nn = (k ** (L + 1) - 1) // (k - 1)
rng = np.random.default_rng(5)
parent = np.zeros(nn, np.int32); parent[1:] = (np.arange(1, nn) - 1) // k
isw = (np.arange(nn) >= nn - k ** L).astype(np.uint8)
voc = T.ORBVocabulary().from_arrays(k, L, 0, 0, parent, isw, rng.integers(0, 256, (nn, 32), dtype=np.uint8), np.where(isw == 1, rng.uniform(0.5, 12, nn), 0.0))
bx.compute_bow(voc, n, 4); bx.sync()
a, b = Event(bx), Event(bx)
a.record(bx)
for _ in range(3):
    bx.compute_bow(voc, n, 4)
b.record(bx)
print("bow ms per 1024 images:", a.elapsed_ms(b) / 3)

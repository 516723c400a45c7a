"""Times (CUDA events) and, under ncu, profiles k_stereo_match on a resident batch of 512 KITTI-shaped stereo frames."""
import sys
sys.path.insert(0, ".")
import tc2li_slam_b200 as T
from tc2li_slam_b200.orbx import Event

n = 1024
bx = T.BatchExtractor(1241, 376, n, 2000, 1.2, 8, 20, 7)
bx.synth(1, n, True)
bx.run(n)
bx.stereo_match(n // 2, 0.5372, 387.5744); bx.sync()
a, b = Event(bx), Event(bx)
a.record(bx)
for _ in range(3):
    bx.stereo_match(n // 2, 0.5372, 387.5744)
b.record(bx)
print("stereo match ms per 512 frames:", a.elapsed_ms(b) / 3)

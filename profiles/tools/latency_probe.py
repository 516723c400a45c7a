"""Where the time of the reference-shaped call goes (one KITTI stereo pair): wall times of the pieces, medians over 200 calls."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import tc2li_slam_b200 as T
from oracle import pyoracle as po

W, H = 1241, 376
L, R = po.synth_v1(W, H, 7), po.synth_v1(W, H, 7, 16, 0x5bd1e995)


def med(f, n=200):
    for _ in range(10):
        f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return round(1e6 * float(np.median(ts)), 1)


ex = T.ORBextractor(2000, 1.2, 8, 20, 7)
st = T.StereoORBextractor(2000, 1.2, 8, 20, 7)
bx = T.BatchExtractor(W, H, 2, 2000, 1.2, 8, 20, 7)
pin = T.PinnedArray((2, H, W), np.uint8)
pin.array[0], pin.array[1] = L, R
dst = np.empty_like(L)
out = {"host_memcpy_one_image_us": med(lambda: np.copyto(dst, L)),
       "single_image_call_us": med(lambda: ex(L)),
       "stereo_call_us": med(lambda: st(L, R))}
bx.upload(pin.array); bx.run(2); bx.sync()
def resident(n):
    bx.run(n); bx.sync()
out["resident_run_1_image_us"] = med(lambda: resident(1))
out["resident_run_2_images_us"] = med(lambda: resident(2))
def up_run(n):
    bx.upload(pin.array[:n]); bx.run(n); bx.sync()
out["upload_run_1_image_us"] = med(lambda: up_run(1))
bx.enable_timing(True)
resident(1); resident(1)
out["stage_ms_1_image_serialised"] = bx.stage_ms()
print(out)

import sys, time
import numpy as np
sys.path.insert(0, ".")
import tc2li_slam_b200 as T
from oracle import pyoracle as po
W, H = 1241, 376
L, R = po.synth_v1(W, H, 7), po.synth_v1(W, H, 7, 16, 0x5bd1e995)
st = T.StereoORBextractor(2000, 1.2, 8, 20, 7)
for _ in range(5): st(L, R)
print("max_images of left.batch:", T.lib().orbx_batch_max_images(T.lib().orbx_extractor_batch(st.left._h)))
ts=[]
for _ in range(300):
    t0=time.perf_counter(); st(L,R); ts.append(time.perf_counter()-t0)
print("stereo us", np.median(ts)*1e6, np.percentile(ts,10)*1e6)
ex = T.ORBextractor(2000, 1.2, 8, 20, 7)
for _ in range(5): ex(L)
ts=[]
for _ in range(300):
    t0=time.perf_counter(); ex(L); ts.append(time.perf_counter()-t0)
print("single us", np.median(ts)*1e6)

"""System::TrackStereoLidar's rectification on the device (System.cc:238-248: cv::remap with the CV_32F maps of
cv::initUndistortRectifyMap, Settings.cc:523-526), in the reference's order remap -> cvtColor -> ORBextractor.  Level 0 must equal
the oracle's remap (+ grey conversion), itself pinned against cv2.remap / cv2.cvtColor (tests/test_oracle_vs_cv2.py) — and cv2's
directly where importable — and the keypoints / descriptors must equal the extraction of that image."""
import numpy as np
import pytest

import tc2li_slam_b200 as T
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

ARGS = (1000, 1.2, 8, 20, 7)


def rect_maps(dw, dh, sw, sh, tilt):
    """a radial undistortion plus a small rotation / shift, as stereoRectify + initUndistortRectifyMap produce; the corners sample
    outside the raw image (constant-0 border)"""
    y, x = np.mgrid[0:dh, 0:dw].astype(np.float32)
    cx, cy = np.float32(dw / 2), np.float32(dh / 2)
    r2 = ((x - cx) ** 2 + (y - cy) ** 2) / np.float32(cx * cx + cy * cy)
    mx = (np.float32(sw / 2) + (x - cx) * (1 + np.float32(0.08) * r2) * np.float32(sw / dw) * np.float32(1.02) + np.float32(tilt) * (y - cy) - 2).astype(np.float32)
    my = (np.float32(sh / 2) + (y - cy) * (1 + np.float32(0.08) * r2) * np.float32(sh / dh) * np.float32(1.02) - np.float32(tilt) * (x - cx) + 1).astype(np.float32)
    return mx, my


def raw_image(w, h, seed, channels):
    planes = [po.synth_v1(w, h, seed + 13 * c) for c in range(3 if channels > 1 else 1)]
    if channels == 4:
        planes.append(np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8))
    return planes[0] if channels == 1 else np.ascontiguousarray(np.stack(planes, axis=-1))


def expected_gray(raw, mx, my, rgb):
    rect = po.remap(raw, mx, my)
    if cv2 is not None:
        assert np.array_equal(rect, cv2.remap(raw, mx, my, cv2.INTER_LINEAR))
    return rect if raw.ndim == 2 else po.cvt_gray(rect, rgb)


@pytest.mark.parametrize("sw,sh,dw,dh,channels,rgb", [(752, 480, 752, 480, 1, False), (752, 480, 720, 460, 3, True), (800, 500, 752, 480, 4, False), (1241, 376, 1241, 376, 1, False)])
def test_rectified_single_image(sw, sh, dw, dh, channels, rgb):
    raw = raw_image(sw, sh, 3, channels)
    mx, my = rect_maps(dw, dh, sw, sh, 0.01)
    want = expected_gray(raw, mx, my, rgb)
    ref = po.OracleExtractor(*ARGS)
    rm, rk, rd = ref(want)
    ex = T.ORBextractor(*ARGS)
    ex.set_rectification(mx, my, (sw, sh))
    for _ in range(3):  # eager, capture, graph replay
        m, k, d = ex(raw) if channels == 1 else ex.extract_color(raw, rgb)
        assert np.array_equal(ex.level(0), want)
        assert m == rm and len(k) == len(rk) > 200
        assert k.tobytes() == rk.tobytes() and np.array_equal(d, rd)
    # a raw image of the wrong size is refused; switching rectification off returns to plain input
    with pytest.raises(T.orbx.OrbxError):
        ex(np.zeros((sh + 1, sw), np.uint8))
    ex.set_rectification(None, None, None)
    m, k, d = ex(want)
    assert k.tobytes() == rk.tobytes() and np.array_equal(d, rd)
    ex.close()


def test_rectified_stereo_pair_with_different_maps():
    """left and right extractor carry their own maps (M1l/M2l vs M1r/M2r)"""
    sw, sh = 752, 480
    rawL, rawR = raw_image(sw, sh, 21, 1), raw_image(sw, sh, 22, 1)
    mL, mR = rect_maps(sw, sh, sw, sh, 0.012), rect_maps(sw, sh, sw, sh, -0.007)
    wantL, wantR = expected_gray(rawL, *mL, False), expected_gray(rawR, *mR, False)
    refL, refR = po.OracleExtractor(*ARGS), po.OracleExtractor(*ARGS)
    wl, wr = refL(wantL), refR(wantR)
    st = T.StereoORBextractor(*ARGS)
    st.left.set_rectification(*mL, (sw, sh))
    st.right.set_rectification(*mR, (sw, sh))
    for _ in range(2):
        (ml, kl, dl), (mr, kr, dr) = st(rawL, rawR)
        assert kl.tobytes() == wl[1].tobytes() and np.array_equal(dl, wl[2])
        assert kr.tobytes() == wr[1].tobytes() and np.array_equal(dr, wr[2])
    st.close()


def test_rectified_batch():
    sw, sh, dw, dh = 800, 500, 752, 480
    mx, my = rect_maps(dw, dh, sw, sh, 0.01)
    raws = np.stack([raw_image(sw, sh, 40 + i, 3) for i in range(3)])
    b = T.BatchExtractor(dw, dh, 3, *ARGS)
    b.set_rectification(mx, my, (sw, sh))
    b.upload_raw(raws, rgb=True)
    b.run(3)
    kps, desc, cnt, mono = b.download(3)
    ref = po.OracleExtractor(*ARGS)
    for i in range(3):
        want = expected_gray(raws[i], mx, my, True)
        assert np.array_equal(b.level(i, 0), want)
        rm, rk, rd = ref(want)
        assert cnt[i] == len(rk) and kps[i, :cnt[i]].tobytes() == rk.tobytes() and np.array_equal(desc[i, :cnt[i]], rd)
    b.close()

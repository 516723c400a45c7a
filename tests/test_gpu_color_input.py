"""The caller's grayscale conversion on the device (Tracking::GrabImageStereoLidar, Tracking.cc:1646-1671): colour input ->
cv::cvtColor arithmetic -> level 0 -> the extractor.  The device grey image must equal the oracle's (pinned against cv2.cvtColor
over all 2^24 colours, tests/test_oracle_vs_cv2.py) and cv2's where importable, and keypoints / descriptors must equal the
extraction of that grey image."""
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


def colour_image(w, h, seed, channels):
    """three differently seeded textures as the colour planes (+ a noisy alpha), so that the grey image has corners"""
    planes = [po.synth_v1(w, h, seed + 17 * c) for c in range(3)]
    if channels == 4:
        planes.append(np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8))
    return np.ascontiguousarray(np.stack(planes, axis=-1))


@pytest.mark.parametrize("w,h,channels,rgb", [(1241, 376, 3, True), (1241, 376, 3, False), (752, 480, 4, True), (752, 480, 4, False), (637, 283, 3, True), (333, 263, 4, False)])
def test_colour_single_image(w, h, channels, rgb):
    img = colour_image(w, h, 5, channels)
    want_gray = po.cvt_gray(img, rgb)
    if cv2 is not None:
        code = {(3, True): cv2.COLOR_RGB2GRAY, (3, False): cv2.COLOR_BGR2GRAY, (4, True): cv2.COLOR_RGBA2GRAY, (4, False): cv2.COLOR_BGRA2GRAY}[(channels, rgb)]
        assert np.array_equal(want_gray, cv2.cvtColor(img, code))
    ref = po.OracleExtractor(*ARGS)
    rm, rk, rd = ref(want_gray)
    ex = T.ORBextractor(*ARGS)
    for _ in range(3):  # eager, capture, graph replay
        m, k, d = ex.extract_color(img, rgb)
        assert np.array_equal(ex.level(0), want_gray)
        assert m == rm and len(k) == len(rk) > 200
        assert k.tobytes() == rk.tobytes() and np.array_equal(d, rd)
    # the same handle goes back to grey input (format change re-captures the graph)
    m, k, d = ex(want_gray)
    assert k.tobytes() == rk.tobytes() and np.array_equal(d, rd)
    ex.close()


def test_colour_strided_rows():
    img = colour_image(400, 300, 9, 3)
    padded = np.zeros((300, 400 * 3 + 37), np.uint8)
    padded[:, :1200] = img.reshape(300, 1200)
    view = np.lib.stride_tricks.as_strided(padded, shape=(300, 400, 3), strides=(padded.strides[0], 3, 1))
    ex = T.ORBextractor(*ARGS)
    n, mono = T.orbx.C.c_int(0), T.orbx.C.c_int(0)
    cap = ex._outputs()
    T.orbx._check(T.orbx.lib().orbx_extract_color(ex._h, T.orbx._p(padded), 400, 300, padded.strides[0], 3, 1, 0, 0, T.orbx._p(ex._kps), T.orbx._p(ex._desc), cap,
                                                  T.orbx.C.byref(n), T.orbx.C.byref(mono)))
    assert np.array_equal(ex.level(0), po.cvt_gray(np.ascontiguousarray(view), True))
    ex.close()


def test_colour_stereo_pair_and_batch():
    w, h = 1241, 376
    imL, imR = colour_image(w, h, 31, 3), colour_image(w, h, 32, 3)
    gL, gR = po.cvt_gray(imL, False), po.cvt_gray(imR, False)
    refL, refR = po.OracleExtractor(*ARGS), po.OracleExtractor(*ARGS)
    wl, wr = refL(gL), refR(gR)
    st = T.StereoORBextractor(*ARGS)
    for _ in range(2):
        (ml, kl, dl), (mr, kr, dr) = st.extract_color(imL, imR, False)
        assert kl.tobytes() == wl[1].tobytes() and np.array_equal(dl, wl[2])
        assert kr.tobytes() == wr[1].tobytes() and np.array_equal(dr, wr[2])
    st.close()
    b = T.BatchExtractor(w, h, 4, *ARGS)
    b.upload_color(np.stack([imL, imR, imR, imL]), False)
    b.run(4)
    kps, desc, cnt, mono = b.download(4)
    for i, want in enumerate((wl, wr, wr, wl)):
        assert cnt[i] == len(want[1]) and kps[i, :cnt[i]].tobytes() == want[1].tobytes() and np.array_equal(desc[i, :cnt[i]], want[2])
    b.close()


def test_colour_argument_errors():
    ex = T.ORBextractor(*ARGS)
    img = np.zeros((64, 64, 2), np.uint8)
    n, mono = T.orbx.C.c_int(0), T.orbx.C.c_int(0)
    rc = T.orbx.lib().orbx_extract_color(ex._h, T.orbx._p(img), 64, 64, 128, 2, 1, 0, 0, None, None, 0, T.orbx.C.byref(n), T.orbx.C.byref(mono))
    assert rc == -2 or rc < 0
    rc = T.orbx.lib().orbx_extract_color(ex._h, T.orbx._p(img), 64, 64, 100, 3, 1, 0, 0, None, None, 0, T.orbx.C.byref(n), T.orbx.C.byref(mono))
    assert rc < 0  # stride smaller than width * channels
    ex.close()

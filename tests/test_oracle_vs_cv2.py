"""Pins the oracle's restated OpenCV primitives against the REAL OpenCV: the cv2 wheel of this image
(4.13).  These are the pixel primitives the reference calls (ORBextractor.cc:76, 800, 1106, 1156, 1158) and the ones the
stand-in headers of oracle/cvstub forward to when the reference's own translation unit is compiled (oracle/_ref).
Bit-exact on every check.  Skipped when cv2 is not importable."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import pyoracle as po


@pytest.fixture(autouse=True, params=[True, False], ids=["ipp_on", "ipp_off"])
def ipp(request):
    if hasattr(cv2, "ipp"):
        cv2.ipp.setUseIPP(request.param)
    cv2.setNumThreads(1)
    yield


@pytest.mark.parametrize("shape", [(376, 1241), (480, 752), (370, 1226), (200, 640), (135, 240)])
def test_resize_chain_matches_cv2(shape):
    h, w = shape
    img = po.synth_v1(w, h, 7)
    sf = 1.0
    cur_o, cur_c = img, img
    for l in range(1, 8):
        sf = np.float32(sf * 1.2)
        inv = np.float32(1.0) / np.float32(sf)
        dw, dh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))
        cur_c = cv2.resize(cur_c, (dw, dh), interpolation=cv2.INTER_LINEAR)
        cur_o = po.resize(cur_o, dw, dh)
        assert np.array_equal(cur_o, cur_c), f"level {l}"


def test_resize_other_ratios():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (301, 517), dtype=np.uint8)
    for dw, dh in [(345, 201), (259, 151), (470, 274), (130, 75), (516, 300)]:
        assert np.array_equal(po.resize(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)), (dw, dh)


@pytest.mark.parametrize("channels", [1, 3, 4])
def test_input_resize_matches_cv2(channels):
    """System::TrackStereoLidar's cv::resize(im, newImSize) (System.cc:250-252): down, up, exact 2x (OpenCV switches to its INTER_AREA
    fast path there, which is the same arithmetic), anisotropic and identity sizes, 1 / 3 / 4 channels, followed by cvtColor."""
    rng = np.random.default_rng(channels)
    src = rng.integers(0, 256, (480, 752) if channels == 1 else (480, 752, channels), dtype=np.uint8)
    for dw, dh in [(600, 350), (1241, 376), (376, 240), (752, 480), (1000, 700), (640, 480), (188, 120)]:
        ref = cv2.resize(src, (dw, dh))
        assert np.array_equal(po.input_resize(src, dw, dh), ref), (dw, dh)
        if channels > 1:
            code = cv2.COLOR_RGB2GRAY if channels == 3 else cv2.COLOR_RGBA2GRAY
            assert np.array_equal(po.input_resize(src, dw, dh, rgb=True), cv2.cvtColor(ref, code)), (dw, dh)


def test_border_reflect101():
    img = po.synth_v1(97, 53, 1)
    assert np.array_equal(np.pad(img, 19, mode="reflect"), cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))


@pytest.mark.parametrize("shape", [(376, 1241), (120, 200), (37, 41), (9, 300)])
def test_gaussian_blur_matches_cv2(shape):
    h, w = shape
    rng = np.random.default_rng(h)
    for img in (po.synth_v1(w, h, 3), rng.integers(0, 256, (h, w), dtype=np.uint8), (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)):
        want = cv2.GaussianBlur(img, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
        assert np.array_equal(po.blur(img), want)


@pytest.mark.parametrize("t", [20, 12, 7, 1])
def test_fast_matches_cv2(t):
    rng = np.random.default_rng(t)
    det = cv2.FastFeatureDetector_create(t, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    imgs = [po.synth_v1(320, 200, 9)] + [po.synth_v1(1241, 376, 2)[y:y + 45, x:x + 42] for y, x in rng.integers(0, 300, (30, 2))]
    imgs += [rng.integers(0, 256, (40, 44), dtype=np.uint8), (rng.integers(0, 2, (48, 52)) * 255).astype(np.uint8), np.full((30, 30), 9, np.uint8),
             po.synth_v1(60, 6, 1), po.synth_v1(6, 60, 1), po.synth_v1(7, 7, 1)]
    for img in imgs:
        img = np.ascontiguousarray(img)
        kps = det.detect(img)
        want = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
        got = po.fast(img, t)
        assert np.array_equal(got, want), (img.shape, len(got), len(want))


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(3)
    ys = rng.integers(-2 ** 21, 2 ** 21, 20000).astype(np.float32)
    xs = rng.integers(-2 ** 21, 2 ** 21, 20000).astype(np.float32)
    ys[:50] = 0
    xs[25:75] = 0
    for y, x in zip(ys, xs):
        assert np.float32(po.fast_atan2(y, x)) == np.float32(cv2.fastAtan2(float(y), float(x))), (y, x)


def _cv2_knn2(dl, dr):
    """What Frame::ComputeStereoFishEyeMatches sees (Frame.cc:1178-1187): BFMatcher(NORM_HAMMING).knnMatch(k=2) + ratio gate."""
    nq = len(dl)
    out = np.full((nq, 4), -1, np.int32)
    ok = np.zeros(nq, np.uint8)
    if nq == 0 or len(dr) == 0:
        return out, ok
    matches = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(dl, dr, 2)
    assert len(matches) == nq
    for q, m in enumerate(matches):
        for k, dm in enumerate(m):
            assert dm.queryIdx == q
            out[q, 2 * k], out[q, 2 * k + 1] = dm.trainIdx, int(dm.distance)
        ok[q] = len(m) >= 2 and m[0].distance < m[1].distance * 0.7
    return out, ok


@pytest.mark.parametrize("nq,nt,distinct", [(300, 280, 0), (257, 513, 0), (64, 1, 0), (5, 2, 0), (40, 0, 0), (0, 17, 0), (200, 200, 6), (150, 90, 2), (33, 700, 12)])
def test_lapping_knn_match_matches_cv2(nq, nt, distinct):
    """`distinct` > 0: descriptors drawn from a few prototypes with a few flipped bits -> many equal distances (tie order)."""
    rng = np.random.default_rng(nq * 1000 + nt)
    if distinct:
        protos = rng.integers(0, 256, (distinct, 32), dtype=np.uint8)

        def draw(n):
            d = protos[rng.integers(0, distinct, n)].copy()
            flip = rng.integers(0, 3, n)
            for i in range(n):
                for _ in range(flip[i]):
                    d[i, rng.integers(0, 32)] ^= np.uint8(1 << rng.integers(0, 8))
            return d
        dl, dr = draw(nq), draw(nt)
    else:
        dl = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        dr = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    want, want_ok = _cv2_knn2(dl, dr)
    got, got_ok = po.lapping_match(dl, dr)
    assert np.array_equal(got, want)
    assert np.array_equal(got_ok, want_ok)


def test_lapping_ratio_gate_boundary():
    """d0 * 10 == d1 * 7 sits exactly on Lowe's boundary: the double product decides (7 < 10 * 0.7 is false in IEEE double)."""
    dl = np.zeros((1, 32), np.uint8)
    dr = np.zeros((2, 32), np.uint8)
    dl[0, 0] = 0x7F                                   # 7 bits away from train 0 (all zero)
    dr[1, 0], dr[1, 1], dr[1, 2] = 0x7F, 0xFF, 0x03   # train 1: 0 + 8 + 2 = 10 bits away
    want, want_ok = _cv2_knn2(dl, dr)
    got, got_ok = po.lapping_match(dl, dr)
    assert want[0].tolist() == [0, 7, 1, 10]
    assert np.array_equal(got, want) and np.array_equal(got_ok, want_ok)
    assert bool(got_ok[0]) == (np.float32(7) < float(np.float32(10)) * 0.7)


@pytest.mark.parametrize("channels,rgb", [(3, True), (3, False), (4, True), (4, False)])
def test_cvt_gray_matches_cv2(channels, rgb):
    """Tracking::GrabImageStereoLidar's cvtColor (Tracking.cc:1646-1671), all four codes."""
    code = {(3, True): cv2.COLOR_RGB2GRAY, (3, False): cv2.COLOR_BGR2GRAY, (4, True): cv2.COLOR_RGBA2GRAY, (4, False): cv2.COLOR_BGRA2GRAY}[(channels, rgb)]
    rng = np.random.default_rng(channels * 2 + rgb)
    for h, w in [(376, 1241), (37, 53), (1, 1), (5, 1023)]:
        img = rng.integers(0, 256, (h, w, channels), dtype=np.uint8)
        assert np.array_equal(po.cvt_gray(img, rgb), cv2.cvtColor(img, code)), (h, w)


def test_cvt_gray_all_colours():
    """every one of the 2^24 (c0, c1, c2) triples, both channel orders"""
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([(v & 255), (v >> 8) & 255, (v >> 16)], axis=-1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(po.cvt_gray(img, True), cv2.cvtColor(img, cv2.COLOR_RGB2GRAY))
    assert np.array_equal(po.cvt_gray(img, False), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


def _rect_maps(dw, dh, sw, sh, seed):
    """maps shaped like a rectification (smooth warp + rotation), reaching outside the source at the corners, with a sprinkling
    of exact-grid and exact-half coordinates (the rounding boundaries of cvRound(v * 32))"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:dh, 0:dw].astype(np.float32)
    cx, cy = np.float32(dw / 2), np.float32(dh / 2)
    r2 = ((x - cx) ** 2 + (y - cy) ** 2) / np.float32(cx * cx + cy * cy)
    mx = (cx + (x - cx) * (1 + 0.12 * r2) * np.float32(sw / dw) * 1.03 + 0.02 * (y - cy) - 3).astype(np.float32)
    my = (cy + (y - cy) * (1 + 0.12 * r2) * np.float32(sh / dh) * 1.03 - 0.02 * (x - cx) + 2).astype(np.float32)
    k = rng.integers(0, dw * dh, 4000)
    mx.flat[k[:1000]] = np.round(mx.flat[k[:1000]])
    my.flat[k[1000:2000]] = np.round(my.flat[k[1000:2000]]) + np.float32(0.5)
    mx.flat[k[2000:3000]] = (np.round(mx.flat[k[2000:3000]] * 32) + np.float32(0.5)) / np.float32(32)   # ties of cvRound(v * 32)
    my.flat[k[3000:]] = (np.round(my.flat[k[3000:]] * 32) - np.float32(0.5)) / np.float32(32)
    return mx, my


@pytest.mark.parametrize("channels", [1, 3, 4])
def test_remap_matches_cv2(channels):
    """System::TrackStereoLidar's cv::remap(im, out, M1, M2, cv::INTER_LINEAR) (System.cc:247-248), CV_32F maps (Settings.cc:523-526)."""
    rng = np.random.default_rng(channels)
    for (sw, sh, dw, dh) in [(752, 480, 752, 480), (640, 400, 600, 380), (97, 61, 131, 83)]:
        shape = (sh, sw) if channels == 1 else (sh, sw, channels)
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        mx, my = _rect_maps(dw, dh, sw, sh, sw + channels)
        want = cv2.remap(img, mx, my, cv2.INTER_LINEAR)
        got = po.remap(img, mx, my)
        assert np.array_equal(got, want), (sw, sh, dw, dh, int((got != want).sum()))


def test_remap_far_outside_and_borders():
    img = np.full((50, 70), 200, np.uint8)
    y, x = np.mgrid[0:64, 0:96].astype(np.float32)
    mx = (x * 1.5 - 30).astype(np.float32)
    my = (y * 1.7 - 25).astype(np.float32)
    mx[0, 0], my[0, 0] = -1e6, 3.0
    mx[1, 1], my[1, 1] = 5.0, 4e5
    assert np.array_equal(po.remap(img, mx, my), cv2.remap(img, mx, my, cv2.INTER_LINEAR))


def undistort_cases(rng, n_trials):
    for trial in range(n_trials):
        fx, fy = rng.uniform(300, 900, 2)
        cam = np.array([fx, fy, rng.uniform(300, 700), rng.uniform(150, 300)], np.float32)
        nd = int(rng.choice([4, 5, 8, 12]))
        d = np.zeros(nd, np.float32)
        d[0], d[1] = rng.uniform(-0.4, 0.4), rng.uniform(-0.2, 0.2)
        d[2:4] = rng.uniform(-0.01, 0.01, 2)
        if nd >= 5:
            d[4] = rng.uniform(-0.05, 0.05)
        if nd >= 8:
            d[5:8] = rng.uniform(-0.05, 0.05, 3)
        if nd >= 12:
            d[8:12] = rng.uniform(-0.002, 0.002, 4)
        new = cam if trial % 2 == 0 else np.array([fx * 0.9, fy * 1.1, cam[2] + 5, cam[3] - 3], np.float32)
        yield cam, d, new


def cv2_undistort(pts, cam, d, new):
    K = np.array([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], np.float32)
    P = np.array([[new[0], 0, new[2]], [0, new[1], new[3]], [0, 0, 1]], np.float32)
    return cv2.undistortPoints(pts.reshape(-1, 1, 2), K, d, None, P).reshape(-1, 2)


def test_undistort_points_matches_cv2():
    """Frame::UndistortKeyPoints' cv::undistortPoints(mat, mat, K, mDistCoef, cv::Mat(), mK) (Frame.cc:777-810): the restatement is bit-identical to OpenCV
    for 4 / 5 / 8 / 12 coefficients, equal and different new camera matrices, points inside and far outside the image (strong distortion: the icdist < 0 exit)."""
    rng = np.random.default_rng(1)
    for cam, d, new in undistort_cases(rng, 120):
        pts = np.stack([rng.uniform(-200, 1500, 1500), rng.uniform(-200, 600, 1500)], 1).astype(np.float32)
        assert np.array_equal(cv2_undistort(pts, cam, d, new).view(np.uint32), po.undistort_points(pts, cam, d, new).view(np.uint32))
    cam = np.array([400, 400, 620, 188], np.float32)   # extreme barrel distortion: the rational term turns negative for far points
    pts = np.stack([rng.uniform(-3000, 4000, 4000), rng.uniform(-2000, 2500, 4000)], 1).astype(np.float32)
    for d in (np.array([-0.9, 0.3, 0, 0], np.float32), np.array([0.5, -2.0, 0.001, 0.001, 3.0], np.float32)):
        assert np.array_equal(cv2_undistort(pts, cam, d, cam).view(np.uint32), po.undistort_points(pts, cam, d, cam).view(np.uint32))

"""System::TrackStereoLidar's input resize on the device (System.cc:250-252: cv::resize(im, imToFeed, settings_->newImSize()) when
Settings::needToResize()), in the reference's order resize -> cvtColor -> ORBextractor.  Level 0 must equal the oracle's resize
(+ grey conversion), itself pinned against cv2.resize / cv2.cvtColor (tests/test_oracle_vs_cv2.py::test_input_resize_matches_cv2) — and
cv2's directly where importable — and the keypoints / descriptors must equal the extraction of that image."""
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


def raw_image(w, h, seed, channels):
    planes = [po.synth_v1(w, h, seed + 13 * c) for c in range(3 if channels > 1 else 1)]
    if channels == 4:
        planes.append(np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8))
    return planes[0] if channels == 1 else np.ascontiguousarray(np.stack(planes, axis=-1))


@pytest.mark.parametrize("sw,sh,dw,dh,channels,rgb", [(1241, 376, 1000, 303, 1, False), (752, 480, 600, 350, 3, True), (752, 480, 376, 240, 4, False),
                                                       (640, 400, 800, 500, 3, False), (1920, 1080, 1241, 376, 1, False), (700, 300, 700, 300, 3, True)])
def test_resized_input_matches_oracle(sw, sh, dw, dh, channels, rgb):
    n = 3
    raws = np.stack([raw_image(sw, sh, 70 + i, channels) for i in range(n)])
    bx = T.BatchExtractor(dw, dh, n, *ARGS)
    with pytest.raises(T.OrbxError):  # neither rectification nor resize configured
        bx.upload_raw(raws, rgb)
    bx.set_input_resize((sw, sh))
    bx.upload_raw(raws, rgb)
    bx.run(n)
    l0 = np.empty((n, dh, dw), np.uint8)
    bx.download_images(0, n, l0)
    kp, de, cnt, mono = bx.download(n)
    ora = po.OracleExtractor(*ARGS)
    for i in range(n):
        exp = po.input_resize(raws[i], dw, dh, rgb if channels > 1 else None)
        assert np.array_equal(l0[i], exp), i
        if cv2 is not None:
            ref = cv2.resize(raws[i], (dw, dh))
            if channels > 1:
                ref = cv2.cvtColor(ref, {(3, True): cv2.COLOR_RGB2GRAY, (3, False): cv2.COLOR_BGR2GRAY, (4, True): cv2.COLOR_RGBA2GRAY,
                                         (4, False): cv2.COLOR_BGRA2GRAY}[(channels, rgb)])
            assert np.array_equal(l0[i], ref), i
        om, ok, od = ora(exp)
        assert cnt[i] == len(ok) and mono[i] == om
        assert kp[i, :cnt[i]].tobytes() == ok.tobytes() and np.array_equal(de[i, :cnt[i]], od)
    # switching it off again restores the plain path; a rectification replaces it
    bx.set_input_resize(None)
    with pytest.raises(T.OrbxError):
        bx.upload_raw(raws, rgb)
    bx.close()


def test_single_image_handles_resize_then_extract():
    """orbx_set_input_resize on the reference-shaped handles: grey and colour raw images, eager call and CUDA-graph replays, a change of the
    fed size, switching back to plain input and on to a rectification."""
    ex = T.ORBextractor(*ARGS)
    ora = po.OracleExtractor(*ARGS)
    for (sw, sh, dw, dh, ch) in ((752, 480, 600, 350, 1), (752, 480, 600, 350, 3), (1241, 376, 1000, 303, 1)):
        ex.set_input_resize((dw, dh), (sw, sh))
        for rep in range(3):
            raw = raw_image(sw, sh, 90 + rep, ch)
            if ch == 1:
                m, k, d = ex(raw, None, (0, 0))
                exp = po.input_resize(raw, dw, dh)
            else:
                m, k, d = ex.extract_color(raw, True, (0, 0))
                exp = po.input_resize(raw, dw, dh, rgb=True)
            assert np.array_equal(ex.level(0), exp), (sw, sh, dw, dh, ch, rep)
            om, ok, od = ora(exp)
            assert m == om and k.tobytes() == ok.tobytes() and np.array_equal(d, od)
        with pytest.raises(T.OrbxError):  # an image that is not of the raw size
            ex(np.zeros((dh, dw), np.uint8), None, (0, 0))
    ex.set_input_resize(None, None)
    img = po.synth_v1(752, 480, 3)
    m, k, d = ex(img, None, (0, 0))
    om, ok, od = ora(img)
    assert m == om and k.tobytes() == ok.tobytes() and np.array_equal(d, od)
    ex.close()

"""Descriptor stage of Frame::ComputeStereoFishEyeMatches on the device (Frame.cc:1162-1187) against the
CPU oracle, which is pinned against cv2.BFMatcher(NORM_HAMMING).knnMatch (tests/test_oracle_vs_cv2.py); where cv2 is
importable the device result is also compared with OpenCV directly.  Bit-exact: both neighbours, both distances (tie order
included) and Lowe's ratio gate for every query of the left lapping tail."""
import numpy as np
import pytest

import tc2li_slam_b200 as T
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def cv2_knn2(dl, dr):
    out = np.full((len(dl), 4), -1, np.int32)
    ok = np.zeros(len(dl), np.uint8)
    if len(dl) == 0 or len(dr) == 0:
        return out, ok
    for q, m in enumerate(cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(np.ascontiguousarray(dl), np.ascontiguousarray(dr), 2)):
        for k, dm in enumerate(m):
            out[q, 2 * k], out[q, 2 * k + 1] = dm.trainIdx, int(dm.distance)
        ok[q] = len(m) >= 2 and m[0].distance < m[1].distance * 0.7
    return out, ok


def check_pair(imL, imR, args, lapL, lapR, min_queries=0):
    st = T.StereoORBextractor(*args)
    for _ in range(2):  # eager call, then the CUDA-graph replay
        (ml, kl, dl), (mr, kr, dr) = st(imL, imR, lapL, lapR)
        knn, ok = st.lapping_match()
        want, want_ok = po.lapping_match(dl[ml:], dr[mr:])
        assert len(knn) == len(dl) - ml >= min_queries
        assert np.array_equal(knn, want), int((knn != want).any(axis=1).sum())
        assert np.array_equal(ok, want_ok)
        if cv2 is not None:
            cw, cok = cv2_knn2(dl[ml:], dr[mr:])
            assert np.array_equal(knn, cw) and np.array_equal(ok, cok)
    st.close()
    return knn, ok


def test_euroc_shaped_lapping_pair():
    """BASELINE.json configs[2]: 752x480, 1200 features, non-trivial lapping windows on both sides."""
    args = (1200, 1.2, 8, 20, 7)
    imL, imR = po.synth_v1(752, 480, 21), po.synth_v1(752, 480, 21, 16, 0x5bd1e995)
    knn, ok = check_pair(imL, imR, args, (250, 751), (0, 501), min_queries=300)
    assert ok.sum() > 20  # the synthetic right image is the left one shifted: many pairs pass Lowe's gate


@pytest.mark.parametrize("lapL,lapR", [((0, 751), (0, 751)), ((300, 400), (0, 0)), ((0, 0), (100, 600)), ((700, 751), (740, 751)), ((100, 700), (745, 751))])
def test_lapping_windows_including_empty_and_tiny_tails(lapL, lapR):
    args = (1200, 1.2, 8, 20, 7)
    imL, imR = po.synth_v1(752, 480, 5), po.synth_v1(752, 480, 6)
    check_pair(imL, imR, args, lapL, lapR)


def test_tails_larger_than_one_shared_memory_tile():
    """4000 features with the whole image as lapping area: more than 2048 train rows -> several tiles are merged."""
    args = (4000, 1.2, 8, 20, 7)
    imL, imR = po.synth_v1(1920, 1080, 31), po.synth_v1(1920, 1080, 31, 16, 0x5bd1e995)
    knn, _ = check_pair(imL, imR, args, (0, 1919), (0, 1919), min_queries=3000)
    assert (knn[:, 0] >= 2048).any() and (knn[:, 0] < 2048).any()


def test_many_equal_distances():
    """A periodic texture repeats descriptors: the order among equal distances must be OpenCV's (smallest train index first)."""
    y, x = np.mgrid[0:300, 0:640]
    img = ((((x // 8) + (y // 8)) % 2) * 200 + 20).astype(np.uint8)
    args = (800, 1.2, 4, 20, 7)
    knn, ok = check_pair(img, img, args, (0, 639), (0, 639), min_queries=100)
    assert (knn[:, 1] == knn[:, 3]).sum() > 10  # ties between the nearest and the second nearest do occur


def test_lapping_match_batch():
    W, H, NF = 752, 480, 5
    b = T.BatchExtractor(W, H, 2 * NF, 1200, 1.2, 8, 20, 7)
    b.synth(777, 2 * NF, stereo=True)
    lap = np.tile(np.array([[250, 751], [0, 501]], np.int32), (NF, 1))
    b.run(2 * NF, lap)
    b.lapping_match(NF)
    knn, ok = b.download_lapping(NF)
    kps, desc, cnt, mono = b.download(2 * NF)
    for f in range(NF):
        nl, ml, nr, mr = cnt[2 * f], mono[2 * f], cnt[2 * f + 1], mono[2 * f + 1]
        want, want_ok = po.lapping_match(desc[2 * f, ml:nl], desc[2 * f + 1, mr:nr])
        assert nl - ml > 100
        assert np.array_equal(knn[f, :nl - ml], want) and np.array_equal(ok[f, :nl - ml], want_ok), f
    b.close()

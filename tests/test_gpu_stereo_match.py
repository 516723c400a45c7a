"""Frame::ComputeStereoMatches on the device against the CPU oracle, which itself is pinned against the
reference's own function bodies (tests/test_oracle_vs_ref.py).  mvuRight / mvDepth must be bit-identical, unmatched (-1)
entries included."""
import numpy as np
import pytest

import tc2li_slam_b200 as T
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def oracle_pair(w, h, seed, args, mb, mbf, xoff=16):
    left, right = po.OracleExtractor(*args), po.OracleExtractor(*args)
    imL, imR = po.synth_v1(w, h, seed), po.synth_v1(w, h, seed, xoff, 0x5bd1e995)
    left(imL)
    right(imR)
    return imL, imR, po.stereo_match(left, right, mb, mbf)


@pytest.mark.parametrize("w,h,seed,args,mb,mbf", [(1241, 376, 7, (2000, 1.2, 8, 20, 7), 0.5372, 387.5744), (1241, 376, 8, (2000, 1.2, 8, 12, 7), 0.5372, 387.5744),
                                                    (752, 480, 9, (1200, 1.2, 8, 20, 7), 0.11, 47.9), (640, 200, 10, (500, 1.2, 6, 20, 7), 0.3, 120.0),
                                                    (1241, 376, 11, (2000, 1.2, 8, 20, 7), 0.2, 60.0), (1920, 1080, 12, (4000, 1.2, 8, 20, 7), 0.5, 400.0)])
def test_stereo_match_single_pair(w, h, seed, args, mb, mbf):
    imL, imR, (ou, od) = oracle_pair(w, h, seed, args, mb, mbf)
    st = T.StereoORBextractor(*args)
    for _ in range(2):  # eager call, then the CUDA-graph replay
        st(imL, imR)
        u, d = st.stereo_match(mb, mbf)
        assert len(u) == len(ou) and (ou >= 0).sum() > 30
        assert np.array_equal(u.view(np.uint32), ou.view(np.uint32)), int((u != ou).sum())
        assert np.array_equal(d.view(np.uint32), od.view(np.uint32))
    st.close()


def test_stereo_match_various_disparities_and_no_match():
    args = (1000, 1.2, 6, 20, 7)
    for xoff, mb, mbf in ((3, 0.5, 300.0), (40, 0.5, 300.0), (0, 0.5, 300.0), (16, 1.0, 8.0)):  # last: maxD = 8 < disparity: nothing matches
        imL, imR, (ou, od) = oracle_pair(800, 300, 50 + xoff, args, mb, mbf, xoff)
        st = T.StereoORBextractor(*args)
        st(imL, imR)
        u, d = st.stereo_match(mb, mbf)
        assert np.array_equal(u.view(np.uint32), ou.view(np.uint32)) and np.array_equal(d.view(np.uint32), od.view(np.uint32)), xoff
        st.close()
    # unrelated right image: few or no matches, still identical
    left, right = po.OracleExtractor(*args), po.OracleExtractor(*args)
    imL, imR = po.synth_v1(800, 300, 1), po.synth_v1(800, 300, 999)
    left(imL)
    right(imR)
    ou, od = po.stereo_match(left, right, 0.5, 300.0)
    st = T.StereoORBextractor(*args)
    st(imL, imR)
    u, d = st.stereo_match(0.5, 300.0)
    assert np.array_equal(u.view(np.uint32), ou.view(np.uint32)) and np.array_equal(d.view(np.uint32), od.view(np.uint32))
    st.close()


def test_stereo_match_batch():
    W, H, NF = 1241, 376, 6
    b = T.BatchExtractor(W, H, 2 * NF, 2000, 1.2, 8, 20, 7)
    b.synth(4242, 2 * NF, stereo=True)
    b.run(2 * NF)
    b.stereo_match(NF, 0.5372, 387.5744)
    u, d = b.download_stereo(NF)
    kps, desc, cnt, mono = b.download(2 * NF)
    left, right = po.OracleExtractor(2000, 1.2, 8, 20, 7), po.OracleExtractor(2000, 1.2, 8, 20, 7)
    for f in range(NF):
        left(po.synth_v1(W, H, 4242 + f))
        right(po.synth_v1(W, H, 4242 + f, 16, 0x5bd1e995))
        ou, od = po.stereo_match(left, right, 0.5372, 387.5744)
        n = cnt[2 * f]
        assert n == len(ou)
        assert np.array_equal(u[f, :n].view(np.uint32), ou.view(np.uint32)) and np.array_equal(d[f, :n].view(np.uint32), od.view(np.uint32)), f
    b.close()

"""The restated oracle against the reference's OWN translation unit (oracle/_ref: slam_framework/src/ORBextractor.cc compiled
verbatim against stand-in OpenCV headers whose pixel primitives are cv2-verified).  Whole pipeline: keypoints (bit-exact,
angles included), descriptors, return value, pyramid levels and their REFLECT_101 borders.  Skipped where oracle/_ref was
never built (it is built wherever /root/reference is mounted and then travels with the repo snapshot)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from oracle import pyref

pytestmark = pytest.mark.skipif(not pyref.available(), reason="oracle/_ref not built (needs /root/reference once)")

CASES = [  # (w, h, seed, args, lapping)
    (1241, 376, 0, (2000, 1.2, 8, 20, 7), (0, 0)),
    (1241, 376, 1, (2000, 1.2, 8, 12, 7), (0, 0)),
    (1226, 370, 2, (2000, 1.2, 8, 12, 7), (0, 0)),
    (752, 480, 3, (1200, 1.2, 8, 20, 7), (250, 751)),
    (1241, 376, 4, (2000, 1.2, 8, 20, 7), (0, 1000)),
    (1920, 1080, 5, (4000, 1.2, 8, 20, 7), (0, 0)),
    (900, 500, 33, (1500, 1.5, 4, 25, 10), (100, 600)),
    (900, 500, 34, (700, 1.1, 10, 20, 7), (0, 0)),
    (640, 200, 35, (500, 1.2, 6, 20, 7), (0, 0)),
]


@pytest.mark.parametrize("w,h,seed,args,lap", CASES)
def test_whole_pipeline_matches_reference_tu(w, h, seed, args, lap):
    img = po.synth_v1(w, h, seed)
    ref, ora = pyref.RefExtractor(*args), po.OracleExtractor(*args)
    rm, rk, rd = ref(img, lap)
    om, ok, od = ora(img, lap)
    assert rm == om and len(rk) == len(ok)
    assert rk.tobytes() == ok.tobytes(), "keypoints (x, y, size, angle, response, octave, class_id)"
    assert np.array_equal(rd, od), "descriptors"
    for l in range(args[2]):
        assert np.array_equal(ref.level(l), ora.level(l)), f"pyramid level {l}"
        assert np.array_equal(ref.level(l, 19), np.pad(ora.level(l), 19, mode="reflect")), f"border of level {l}"
    rt, ot = ref.tables(), ora.tables()
    for k in rt:
        assert np.array_equal(rt[k].view(np.uint32), ot[k].view(np.uint32)), k


@pytest.mark.parametrize("kind", ["constant", "low_contrast", "checkerboard", "saturated_noise", "random"])
def test_degenerate_inputs_match_reference_tu(kind):
    W, H = 640, 200
    rng = np.random.default_rng(5)
    if kind == "constant":
        img = np.full((H, W), 128, np.uint8)
    elif kind == "low_contrast":
        img = (120 + (po.synth_v1(W, H, 3) >> 4)).astype(np.uint8)
    elif kind == "checkerboard":
        yy, xx = np.mgrid[0:H, 0:W]
        img = (((xx // 8 + yy // 8) & 1) * 200 + 20).astype(np.uint8)
    elif kind == "saturated_noise":
        img = (rng.integers(0, 2, (H, W)) * 255).astype(np.uint8)
    else:
        img = rng.integers(0, 256, (H, W), dtype=np.uint8)
    args = (500, 1.2, 6, 20, 7)
    rm, rk, rd = pyref.RefExtractor(*args)(img)
    om, ok, od = po.OracleExtractor(*args)(img)
    assert rm == om and rk.tobytes() == ok.tobytes() and np.array_equal(rd, od)


def test_empty_image_returns_minus_one():
    assert pyref.RefExtractor()(None)[0] == -1
    assert po.OracleExtractor()(None)[0] == -1


def test_many_seeds_kitti():
    args = (2000, 1.2, 8, 20, 7)
    ref, ora = pyref.RefExtractor(*args), po.OracleExtractor(*args)
    for seed in range(100, 112):
        img = po.synth_v1(1241, 376, seed, 16 * (seed & 1), 0x5bd1e995 * (seed & 1))
        rm, rk, rd = ref(img)
        om, ok, od = ora(img)
        assert rm == om and rk.tobytes() == ok.tobytes() and np.array_equal(rd, od), seed


@pytest.mark.skipif(not pyref.stereo_available(), reason="oracle/_ref/libref_stereo.so not built")
@pytest.mark.parametrize("w,h,seed,args,mb,mbf", [(1241, 376, 7, (2000, 1.2, 8, 20, 7), 0.5372, 387.5744), (1241, 376, 8, (2000, 1.2, 8, 12, 7), 0.5372, 387.5744),
                                                    (752, 480, 9, (1200, 1.2, 8, 20, 7), 0.11, 47.9), (640, 200, 10, (500, 1.2, 6, 20, 7), 0.3, 120.0),
                                                    (1241, 376, 11, (2000, 1.2, 8, 20, 7), 0.2, 60.0)])
def test_stereo_match_restatement_equals_reference_function(w, h, seed, args, mb, mbf):
    """oracle stereo_match vs the reference's own Frame::ComputeStereoMatches + ORBmatcher::DescriptorDistance (verbatim
    function bodies compiled around a stand-in Frame): uRight and depth bit-identical, including the unmatched (-1) entries."""
    left, right = po.OracleExtractor(*args), po.OracleExtractor(*args)
    ml, kl, dl = left(po.synth_v1(w, h, seed))
    mr, kr, dr = right(po.synth_v1(w, h, seed, 16, 0x5bd1e995))
    ou, od = po.stereo_match(left, right, mb, mbf)
    t = left.tables()
    nl = args[2]
    ru, rd = pyref.stereo_match(kl, dl, kr, dr, [left.level(l) for l in range(nl)], [right.level(l) for l in range(nl)],
                                t["scale_factors"], t["inv_scale_factors"], mb, mbf)
    assert (ou >= 0).sum() > 50
    assert np.array_equal(ou.view(np.uint32), ru.view(np.uint32)) and np.array_equal(od.view(np.uint32), rd.view(np.uint32))

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


# ---- the feature grid (Frame::AssignFeaturesToGrid / GetFeaturesInArea) against the reference's own function bodies ----
def _grid_keypoints(kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "extracted":
        _, k, _ = po.OracleExtractor(2000, 1.2, 8, 20, 7)(po.synth_v1(1241, 376, seed))
        return k, (0.0, 1241.0, 0.0, 376.0)
    n = 3000
    k = np.zeros(n, po.KP_DTYPE)
    if kind == "cell_edges":  # coordinates on and next to the .5 rounding boundaries of PosInGrid, and outside the image bounds
        inv_w, inv_h = np.float32(64) / np.float32(1241), np.float32(48) / np.float32(376)
        cx = (rng.integers(-2, 67, n).astype(np.float32) + np.float32(0.5)) / inv_w
        cy = (rng.integers(-2, 51, n).astype(np.float32) + np.float32(0.5)) / inv_h
        k["x"] = np.nextafter(cx, cx + rng.choice([-1, 0, 1], n).astype(np.float32) * 10).astype(np.float32)
        k["y"] = np.nextafter(cy, cy + rng.choice([-1, 0, 1], n).astype(np.float32) * 10).astype(np.float32)
        bounds = (0.0, 1241.0, 0.0, 376.0)
    else:  # "offset_bounds": undistorted-style bounds that do not start at 0, clustered points (long cell lists)
        k["x"] = (rng.normal(300, 40, n)).astype(np.float32)
        k["y"] = (rng.normal(200, 25, n)).astype(np.float32)
        bounds = (-13.25, 765.5, -9.75, 489.125)
    k["octave"] = rng.integers(0, 8, n)
    k["size"] = 31
    k["class_id"] = -1
    return k, bounds


@pytest.mark.skipif(not pyref.grid_available(), reason="oracle/_ref/libref_grid.so not built (needs /root/reference once)")
@pytest.mark.parametrize("kind,seed", [("extracted", 0), ("extracted", 7), ("cell_edges", 1), ("cell_edges", 2), ("offset_bounds", 3)])
def test_feature_grid_matches_reference_bodies(kind, seed):
    k, bounds = _grid_keypoints(kind, seed)
    ref = pyref.RefFrameGrid(k, bounds)
    rs, ri = ref.grid()
    os_, oi = po.assign_grid(k, bounds)
    assert np.array_equal(rs, os_) and np.array_equal(ri, oi)
    rng = np.random.default_rng(seed + 100)
    for q in range(400):
        if q % 4 == 0:  # around an existing keypoint, radius such that |dx| == r happens (strict compare)
            j = rng.integers(0, len(k))
            x, y = float(k["x"][j]) + float(rng.integers(-3, 4)), float(k["y"][j]) + float(rng.integers(-3, 4))
            r = float(rng.integers(1, 40))
        else:
            x = float(rng.uniform(bounds[0] - 60, bounds[1] + 60))
            y = float(rng.uniform(bounds[2] - 60, bounds[3] + 60))
            r = float(rng.choice([0.5, 3, 7.5, 15, 40, 100, 2000]))
        lv = [(-1, -1), (0, 0), (2, 5), (3, -1), (0, 7), (-1, 2)][q % 6]
        a = ref.features_in_area(x, y, r, *lv)
        b = po.features_in_area(k, (os_, oi), bounds, x, y, r, *lv)
        assert np.array_equal(a, b), (q, x, y, r, lv)


@pytest.mark.skipif(not pyref.grid_available(), reason="oracle/_ref/libref_grid.so not built (needs /root/reference once)")
def test_feature_grid_two_grid_case_is_one_grid_per_image():
    """Fisheye stereo (Nleft != -1): mGrid bins mvKeys, mGridRight bins mvKeysRight with indices i - Nleft (Frame.cc:428-441), and
    GetFeaturesInArea(..., bRight) reads the matching pair — i.e. exactly one single-grid per image, which is what the batch API builds."""
    bounds = (0.0, 752.0, 0.0, 480.0)
    ex = po.OracleExtractor(1200, 1.2, 8, 20, 7)
    _, kl, _ = ex(po.synth_v1(752, 480, 21), (250, 751))
    _, kr, _ = ex(po.synth_v1(752, 480, 21, 16, 0x5bd1e995), (0, 501))
    ref = pyref.RefFrameGrid(kl, bounds, kr)
    rng = np.random.default_rng(9)
    for right, k in ((False, kl), (True, kr)):
        rs, ri = ref.grid(right)
        os_, oi = po.assign_grid(k, bounds)
        assert np.array_equal(rs, os_) and np.array_equal(ri, oi), right
        for q in range(200):
            x, y, r = float(rng.uniform(-30, 780)), float(rng.uniform(-30, 510)), float(rng.choice([3, 10, 25, 80]))
            lv = [(-1, -1), (1, 4), (0, 2), (5, -1)][q % 4]
            assert np.array_equal(ref.features_in_area(x, y, r, *lv, right=right), po.features_in_area(k, (os_, oi), bounds, x, y, r, *lv)), (right, q)


# ---- the search of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) against the reference's own function body ----
def projection_case(seed, w=1241, h=376, args=(2000, 1.2, 8, 20, 7), shift=5):
    """Current frame = a synthetic image; 'last frame' = the same scene `shift` px to the side: its keypoints become the map points
    (position, octave, angle, descriptor), with random depth, observation counts (0 = temporal point that does not block) and noise."""
    rng = np.random.default_rng(seed)
    ex = po.OracleExtractor(*args)
    _, kc, dc = ex(po.synth_v1(w, h, seed))
    _, kl, dl = ex(po.synth_v1(w, h, seed, shift))
    sf = ex.tables()["scale_factors"]
    n = len(kl)
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = kl["x"] - np.float32(shift) + rng.normal(0, 1.5, n).astype(np.float32)
    pts[:, 1] = kl["y"] + rng.normal(0, 1.5, n).astype(np.float32)
    pts[:, 2] = rng.uniform(2, 60, n).astype(np.float32)
    pts[:, 3] = kl["angle"]
    pts[::7, 3] = (kl["angle"][::7] + rng.uniform(0, 360, len(kl[::7])).astype(np.float32)) % np.float32(360)  # wrong orientation: rotation check
    nobs = rng.choice([0, 1, 3], n, p=[0.2, 0.3, 0.5]).astype(np.int32)
    pd = dl.copy()
    flip = rng.random((n, 32)) < 0.02
    pd ^= (flip * rng.integers(1, 256, (n, 32))).astype(np.uint8)
    ur = np.where(rng.random(len(kc)) < 0.7, kc["x"] - rng.uniform(0.5, 40, len(kc)).astype(np.float32), np.float32(-1)).astype(np.float32)
    return kc, dc, ur, (0.0, float(w), 0.0, float(h)), sf, pts, kl["octave"].astype(np.int32), nobs, pd


@pytest.mark.skipif(not pyref.search_available(), reason="oracle/_ref/libref_search.so not built (needs /root/reference once)")
@pytest.mark.parametrize("seed,th,mode,check", [(1, 15.0, 0, True), (2, 7.0, 0, True), (3, 15.0, 1, True), (4, 15.0, 2, True), (5, 30.0, 0, False), (6, 15.0, 0, True)])
def test_search_by_projection_matches_reference_body(seed, th, mode, check):
    kc, dc, ur, bounds, sf, pts, octv, nobs, pd = projection_case(seed)
    mb, mbf = 0.5372, 387.5744
    tlcz = {0: 0.0, 1: 2.0 * mb, 2: -2.0 * mb}[mode]
    rm, rn = pyref.search_by_projection(kc, dc, ur, bounds, sf, mb, mbf, pts, octv, nobs, pd, th, tlcz, False, check)
    grid = po.assign_grid(kc, bounds)
    q = po.make_proj_queries(pts, octv, nobs, pd, sf, th, mbf, mode)
    om, on = po.search_by_projection(kc, dc, ur, grid, bounds, q, check)
    assert rn == on and rn > 300, (rn, on)
    assert np.array_equal(rm, om), int((rm != om).sum())


def local_map_case(seed, w=1241, h=376, args=(2000, 1.2, 8, 20, 7)):
    """Local map = the keypoints of two shifted views of the scene (so most windows hold a right and a near-duplicate candidate: the
    ratio test fires), with predicted levels off by up to one, mixed viewing cosines, depths, bad / out-of-view points."""
    rng = np.random.default_rng(seed)
    ex = po.OracleExtractor(*args)
    _, kc, dc = ex(po.synth_v1(w, h, seed))
    sf = ex.tables()["scale_factors"]
    parts = [ex(po.synth_v1(w, h, seed, s))[1:] + (s,) for s in (3, 6)]
    k = np.concatenate([p[0] for p in parts])
    d = np.concatenate([p[1] for p in parts])
    sh = np.concatenate([np.full(len(p[0]), p[2], np.float32) for p in parts])
    n = len(k)
    pts = np.zeros((n, 5), np.float32)
    pts[:, 0] = k["x"] - sh + rng.normal(0, 1.0, n).astype(np.float32)
    pts[:, 1] = k["y"] + rng.normal(0, 1.0, n).astype(np.float32)
    pts[:, 2] = pts[:, 0] - rng.uniform(0.5, 40, n).astype(np.float32)
    pts[:, 3] = rng.choice([0.9, 0.9985, 0.9999], n).astype(np.float32)
    pts[:, 4] = rng.uniform(1, 80, n).astype(np.float32)
    lv = np.clip(k["octave"] + rng.integers(0, 2, n), 0, args[2] - 1).astype(np.int32)
    nobs = rng.choice([0, 1, 3], n, p=[0.15, 0.35, 0.5]).astype(np.int32)
    flags = (rng.random(n) < 0.9).astype(np.int32) | ((rng.random(n) < 0.05).astype(np.int32) << 1)
    pd = d.copy()
    flip = rng.random((n, 32)) < 0.03
    pd ^= (flip * rng.integers(1, 256, (n, 32))).astype(np.uint8)
    order = rng.permutation(n)
    ur = np.where(rng.random(len(kc)) < 0.7, kc["x"] - rng.uniform(0.5, 40, len(kc)).astype(np.float32), np.float32(-1)).astype(np.float32)
    taken0 = rng.choice([0, 1, 2], len(kc), p=[0.7, 0.2, 0.1]).astype(np.uint8)
    return kc, dc, ur, taken0, (0.0, float(w), 0.0, float(h)), sf, pts[order], lv[order], nobs[order], flags[order], pd[order]


@pytest.mark.skipif(not pyref.search_available(), reason="oracle/_ref/libref_search.so not built (needs /root/reference once)")
@pytest.mark.parametrize("seed,th,ratio,far", [(1, 1.0, 0.8, False), (2, 3.0, 0.8, False), (3, 5.0, 0.9, True), (4, 15.0, 0.6, False), (5, 1.0, 0.8, True)])
def test_search_local_points_matches_reference_body(seed, th, ratio, far):
    kc, dc, ur, taken0, bounds, sf, pts, lv, nobs, flags, pd = local_map_case(seed)
    rm, rn = pyref.search_local_points(kc, dc, ur, taken0, bounds, sf, pts, lv, nobs, flags, pd, th, ratio, far, 40.0)
    q, sel = po.make_local_queries(pts, lv, nobs, flags, pd, sf, th, far, 40.0)
    om, on = po.search_local_points(kc, dc, ur, taken0 == 1, po.assign_grid(kc, bounds), bounds, q, ratio)
    assert rn == on and rn > 100, (rn, on)
    assert np.array_equal(rm, np.where(om >= 0, sel[np.maximum(om, 0)], -1)), int((rm != np.where(om >= 0, sel[np.maximum(om, 0)], -1)).sum())


def keyframe_queries(pts, lv, flags, pd, sf, th):
    """The map points the relocalisation search actually looks for (ORBmatcher.cc:1915-1947: non-NULL, not bad, not already found; the bounds and
    distance tests pass here), as queries."""
    sel = np.nonzero((flags & 7) == 0)[0]
    q = np.zeros(len(sel), po.PROJ_QUERY_DTYPE)
    q["u"], q["v"], q["angle"] = pts[sel, 0], pts[sel, 1], pts[sel, 3]
    q["radius"] = np.float32(th) * np.ascontiguousarray(sf, np.float32)[lv[sel]]
    q["min_level"], q["max_level"], q["blocks"] = lv[sel] - 1, lv[sel] + 1, 1
    q["desc"] = pd[sel]
    return q, sel


@pytest.mark.skipif(not pyref.search_available(), reason="oracle/_ref/libref_search.so not built (needs /root/reference once)")
@pytest.mark.parametrize("seed,th,dist,check", [(11, 10.0, 100, True), (12, 3.0, 64, True), (13, 10.0, 100, False)])
def test_search_keyframe_points_matches_reference_body(seed, th, dist, check):
    kc, dc, _, bounds, sf, pts, octv, nobs, pd = projection_case(seed)
    rng = np.random.default_rng(seed)
    flags = (rng.random(len(pts)) < 0.05).astype(np.int32) | ((rng.random(len(pts)) < 0.05).astype(np.int32) << 1) | ((rng.random(len(pts)) < 0.2).astype(np.int32) << 2)
    taken0 = (rng.random(len(kc)) < 0.3).astype(np.uint8)
    rm, rn = pyref.search_keyframe_points(kc, dc, taken0, bounds, sf, pts, octv, flags, pd, th, dist, check)
    q, sel = keyframe_queries(pts, octv, flags, pd, sf, th)
    om, on = po.search_by_projection(kc, dc, None, po.assign_grid(kc, bounds), bounds, q, check, dist, taken0)
    assert rn == on and rn > 20, (rn, on)
    assert np.array_equal(rm, np.where(om >= 0, sel[np.maximum(om, 0)], -1))


# ---- DBoW2 vocabulary transform (Frame::ComputeBoW, Frame.cc:768-775) ----
BOW_CASES = [  # (k, L, scoring, weighting): L1 / L2 / chi-square / KL / Bhattacharyya / dot-product x TF-IDF / TF / IDF / binary
    (10, 6, 0, 0), (10, 4, 1, 0), (6, 3, 0, 1), (4, 5, 5, 2), (3, 4, 2, 3), (10, 3, 4, 0), (10, 3, 3, 0), (9, 3, 5, 0), (2, 8, 0, 0),
]


@pytest.mark.skipif(not pyref.RefVocabulary.available(), reason="oracle/_ref/libref_bow.so not built")
@pytest.mark.parametrize("k,L,scoring,weighting", BOW_CASES)
def test_bow_transform_restatement_equals_vendored_dbow2(tmp_path, k, L, scoring, weighting):
    """oracle vocabulary_* == the reference's vendored DBoW2 (TemplatedVocabulary::loadFromTextFile + transform, FORB::distance,
    BowVector::addWeight / addIfNotExist / normalize) on synthetic vocabularies in ORBvoc.txt's format (the real file is absent from
    the reference checkout): word ids, values (as float64 bits), FeatureVector node ids and their feature lists."""
    path = str(tmp_path / "voc.txt")
    po.write_vocabulary_text(path, k, L, scoring, weighting, seed=100 * k + L)
    ora, ref = po.OracleVocabulary(path), pyref.RefVocabulary(path)
    assert ora.n_words == ref.n_words == k ** L
    rng = np.random.default_rng(k + L)
    ext = po.OracleExtractor(500, 1.2, 4, 20, 7)
    _, _, real = ext(po.synth_v1(640, 200, 7))
    sets = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in (0, 1, 7, 500, 2000)] + [real, np.repeat(real[:5], 40, axis=0)]
    for d in sets:
        for levelsup in (0, 2, 4, L - 1, L, L + 2):
            a, b = ora.transform(d, levelsup), ref.transform(d, levelsup)
            for x, y, what in zip(a, b, ("words", "values", "nodes", "node_start", "features")):
                assert x.tobytes() == y.tobytes(), (what, len(d), levelsup)


# ---- two-grid (fisheye stereo, Nleft != -1) branches of both searches: ORBmatcher.cc:1803-1868 and :153-218 ----
import two_grid_cases as tg  # noqa: E402


def oracle_stereo_frame(seed, kind):
    ex = po.OracleExtractor(*tg.ARGS)
    il, ir = tg.stereo_images(seed, kind)
    (_, kl, dl), (_, kr, dr) = ex(il), ex(ir)
    return ex.tables()["scale_factors"], (kl, dl), (kr, dr)


@pytest.mark.skipif(not pyref.search_available(), reason="oracle/_ref/libref_search.so not built (needs /root/reference once)")
@pytest.mark.parametrize("seed,th,mode,check,kind", [(21, 15.0, 0, True, "plain"), (22, 7.0, 0, True, "plain"), (23, 15.0, 1, True, "plain"), (24, 15.0, 2, False, "plain"),
                                                    (25, 30.0, 0, True, "plain"), (26, 7.0, 0, True, "band"), (27, 15.0, 0, False, "band")])
def test_two_grid_search_by_projection_matches_reference_body(seed, th, mode, check, kind):
    sf, (kl, dl), (kr, dr) = oracle_stereo_frame(seed, kind)
    pts, octv, nobs, pd = tg.frame_to_frame_points(seed, sf, *tg.last_frame(seed))
    mb = 0.11
    tlcz = {0: 0.0, 1: 2.0 * mb, 2: -2.0 * mb}[mode]
    rm, rn = pyref.search_by_projection_stereo(kl, kr, np.concatenate([dl, dr]), tg.BOUNDS, sf, mb, 47.9, pts, octv, nobs, pd, th, tlcz, tg.DX_R, tg.DY_R, check)
    q, side, owner = po.make_proj_queries_stereo(pts, octv, nobs, pd, sf, th, mode, tg.DX_R, tg.DY_R)
    left, right = (kl, dl, po.assign_grid(kl, tg.BOUNDS)), (kr, dr, po.assign_grid(kr, tg.BOUNDS))
    ml, mr, on = po.search_by_projection_stereo(left, right, tg.BOUNDS, q, side, check)
    om = np.concatenate([ml, mr])
    om = np.where(om >= 0, owner[np.maximum(om, 0)], -1)
    assert rn == on and rn > 300, (rn, on)
    assert np.array_equal(rm, om), int((rm != om).sum())
    assert (rm[:len(kl)] >= 0).sum() > 100 and (rm[len(kl):] >= 0).sum() > 100
    if kind == "band":   # the case is sensitive to the cancelled right search: treating the right parts as independent gives another answer
        _, mr2, on2 = po.search_by_projection_stereo(left, right, tg.BOUNDS, q, np.where(side == 1, 2, side).astype(np.uint8), check)
        assert on2 != on or not np.array_equal(mr, mr2)


@pytest.mark.skipif(not pyref.search_available(), reason="oracle/_ref/libref_search.so not built (needs /root/reference once)")
@pytest.mark.parametrize("seed,th,ratio,far,kind", [(31, 1.0, 0.8, False, "plain"), (32, 3.0, 0.8, False, "plain"), (33, 5.0, 0.9, True, "plain"), (34, 15.0, 0.6, False, "plain"),
                                                   (35, 15.0, 0.8, False, "periodic"), (36, 15.0, 0.9, True, "periodic")])
def test_two_grid_search_local_points_matches_reference_body(seed, th, ratio, far, kind):
    sf, (kl, dl), (kr, dr) = oracle_stereo_frame(seed, kind)
    taken0, l2r, r2l, pts, lv, lvr, nobs, flags, pd = tg.local_map_points(seed, len(kl), len(kr), kind)
    rm, rn = pyref.search_local_points_stereo(kl, kr, np.concatenate([dl, dr]), taken0, l2r, r2l, tg.BOUNDS, sf, pts, lv, lvr, nobs, flags, pd, th, ratio, far, 40.0)
    q, side, owner = po.make_local_queries_stereo(pts, lv, lvr, nobs, flags, pd, sf, th, far, 40.0)
    assert (side == 0).sum() > 100 and (side == 1).sum() > 100 and (side == 2).sum() > 20
    left, right = (kl, dl, po.assign_grid(kl, tg.BOUNDS)), (kr, dr, po.assign_grid(kr, tg.BOUNDS))
    tl, tr = taken0[:len(kl)] == 1, taken0[len(kl):] == 1
    ml, mr, on = po.search_local_points_stereo(left, right, tl, tr, l2r, r2l, tg.BOUNDS, q, side, ratio)
    om = np.concatenate([ml, mr])
    om = np.where(om >= 0, owner[np.maximum(om, 0)], -1)
    assert rn == on and rn > 100, (rn, on)
    assert np.array_equal(rm, om), int((rm != om).sum())
    if kind == "periodic":   # sensitive to the ratio-test `continue` that also cancels the right search, and to the left/right links
        ml2, mr2, on2 = po.search_local_points_stereo(left, right, tl, tr, l2r, r2l, tg.BOUNDS, q, np.where(side == 1, 2, side).astype(np.uint8), ratio)
        assert on2 != on
        ml3, mr3, on3 = po.search_local_points_stereo(left, right, tl, tr, None, None, tg.BOUNDS, q, side, ratio)
        assert on3 != on

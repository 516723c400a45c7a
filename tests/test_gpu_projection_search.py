"""The search of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) on the device against the
CPU oracle, which is pinned against the reference's own function body (tests/test_oracle_vs_ref.py::test_search_by_projection_*).
CurrentFrame.mvpMapPoints (as query indices) and the returned match count must be identical."""
import numpy as np
import pytest

import tc2li_slam_b200 as T
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
MB, MBF = 0.5372, 387.5744


def make_case(bx, w, h, seed, shift, rng):
    """Two images resident in `bx`: 0 = current left, 1 = current right; the map points come from the same scene shifted by `shift`."""
    cur_l, cur_r = po.synth_v1(w, h, seed), po.synth_v1(w, h, seed, 16, 0x5bd1e995)
    bx.upload(np.stack([cur_l, cur_r]))
    bx.run(2)
    bx.stereo_match(1, MB, MBF)
    bx.assign_grid(2)
    kp, de, cnt, _ = bx.download(2)
    u, _ = bx.download_stereo(1)
    kc, dc, ur = kp[0, :cnt[0]], de[0, :cnt[0]], u[0, :cnt[0]]
    return kc, dc, ur


def points_from(args, w, h, seed, shift, rng):
    ex = po.OracleExtractor(*args)
    _, kl, dl = ex(po.synth_v1(w, h, seed, shift))
    n = len(kl)
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = kl["x"] - np.float32(shift) + rng.normal(0, 1.5, n).astype(np.float32)
    pts[:, 1] = kl["y"] + rng.normal(0, 1.5, n).astype(np.float32)
    pts[:, 2] = rng.uniform(2, 60, n).astype(np.float32)
    pts[:, 3] = kl["angle"]
    pts[::7, 3] = (kl["angle"][::7] + rng.uniform(0, 360, len(kl[::7])).astype(np.float32)) % np.float32(360)
    nobs = rng.choice([0, 1, 3], n, p=[0.2, 0.3, 0.5]).astype(np.int32)
    pd = dl.copy()
    flip = rng.random((n, 32)) < 0.02
    pd ^= (flip * rng.integers(1, 256, (n, 32))).astype(np.uint8)
    return pts, kl["octave"].astype(np.int32), nobs, pd, ex.tables()["scale_factors"]


@pytest.mark.parametrize("w,h,args,seed,th,mode,check,use_ur", [(1241, 376, (2000, 1.2, 8, 20, 7), 1, 15.0, 0, True, True), (1241, 376, (2000, 1.2, 8, 20, 7), 2, 7.0, 0, True, True),
                                                                 (1241, 376, (2000, 1.2, 8, 20, 7), 3, 15.0, 1, True, True), (1241, 376, (2000, 1.2, 8, 20, 7), 4, 15.0, 2, True, False),
                                                                 (752, 480, (1200, 1.2, 8, 20, 7), 5, 30.0, 0, False, True), (640, 200, (500, 1.2, 6, 20, 7), 6, 15.0, 0, True, True),
                                                                 (1920, 1080, (4000, 1.2, 8, 20, 7), 7, 15.0, 0, True, True)])
def test_projection_search_matches_oracle(w, h, args, seed, th, mode, check, use_ur):
    rng = np.random.default_rng(seed)
    bx = T.BatchExtractor(w, h, 2, *args)
    kc, dc, ur = make_case(bx, w, h, seed, 5, rng)
    pts, octv, nobs, pd, sf = points_from(args, w, h, seed, 5, rng)
    bounds = (0.0, float(w), 0.0, float(h))
    q = po.make_proj_queries(pts, octv, nobs, pd, sf, th, MBF, mode)
    om, on = po.search_by_projection(kc, dc, ur if use_ur else None, po.assign_grid(kc, bounds), bounds, q, check)
    for _ in range(2):
        m, nm = bx.search_by_projection(0, q.view(T.PROJ_QUERY_DTYPE), use_ur, check)
        assert nm == on and on > 0.1 * len(q), (nm, on)
        assert np.array_equal(m[:len(kc)], om), int((m[:len(kc)] != om).sum())
        assert (m[len(kc):] == -1).all()
    bx.close()


def test_projection_search_edge_cases():
    w, h, args = 640, 200, (500, 1.2, 6, 20, 7)
    rng = np.random.default_rng(0)
    bx = T.BatchExtractor(w, h, 2, *args)
    q1 = np.zeros(1, T.PROJ_QUERY_DTYPE)
    q1["radius"], q1["min_level"], q1["max_level"] = 10.0, -1, -1
    with pytest.raises(T.OrbxError):  # no grid yet
        bx.search_by_projection(0, q1)
    kc, dc, ur = make_case(bx, w, h, 11, 5, rng)
    m, nm = bx.search_by_projection(0, np.zeros(0, T.PROJ_QUERY_DTYPE))
    assert nm == 0 and (m == -1).all()
    # all points on one keypoint: only the first blocking point may keep it; a non-blocking point in front is overwritten
    j = len(kc) // 2
    q = np.zeros(4, T.PROJ_QUERY_DTYPE)
    q["u"], q["v"], q["radius"], q["ur"] = kc["x"][j], kc["y"][j], 0.75, -1000.0
    q["min_level"], q["max_level"], q["angle"] = -1, -1, kc["angle"][j]
    q["desc"] = dc[j]
    q["blocks"] = [0, 1, 1, 0]
    bounds = (0.0, float(w), 0.0, float(h))
    om, on = po.search_by_projection(kc, dc, None, po.assign_grid(kc, bounds), bounds, q, False)
    m, nm = bx.search_by_projection(0, q, False, False)
    assert nm == on and np.array_equal(m[:len(kc)], om)
    assert om[j] == 1 and on >= 2
    # a window far outside the image and a point whose descriptor is too far from everything
    q2 = q[:2].copy()
    q2["u"][0] = 5000.0
    q2["desc"][1] = ~dc[j]
    m, nm = bx.search_by_projection(0, q2, False, True)
    om, on = po.search_by_projection(kc, dc, None, po.assign_grid(kc, bounds), bounds, q2, True)
    assert nm == on and np.array_equal(m[:len(kc)], om)
    bx.close()


@pytest.mark.parametrize("w,h,args,seed,th,ratio,use_ur", [(1241, 376, (2000, 1.2, 8, 20, 7), 21, 1.0, 0.8, True), (1241, 376, (2000, 1.2, 8, 20, 7), 22, 3.0, 0.8, True),
                                                            (1241, 376, (2000, 1.2, 8, 20, 7), 23, 5.0, 0.9, False), (752, 480, (1200, 1.2, 8, 20, 7), 24, 15.0, 0.6, True),
                                                            (1920, 1080, (4000, 1.2, 8, 20, 7), 25, 3.0, 0.8, True)])
def test_local_map_search_matches_oracle(w, h, args, seed, th, ratio, use_ur):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, ...): the map = keypoints of two shifted views (near-duplicate candidates make the
    ratio test fire), predicted levels off by up to one, keypoints already taken before the search."""
    rng = np.random.default_rng(seed)
    bx = T.BatchExtractor(w, h, 2, *args)
    kc, dc, ur = make_case(bx, w, h, seed, 0, rng)
    ex = po.OracleExtractor(*args)
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
    flags = (rng.random(n) < 0.9).astype(np.int32)
    order = rng.permutation(n)
    q, _ = po.make_local_queries(pts[order], lv[order], nobs[order], flags[order], d[order], sf, th)
    taken = rng.random(len(kc)) < 0.2
    bounds = (0.0, float(w), 0.0, float(h))
    om, on = po.search_local_points(kc, dc, ur if use_ur else None, taken, po.assign_grid(kc, bounds), bounds, q, ratio)
    for _ in range(2):
        m, nm = bx.search_local_points(0, q.view(T.PROJ_QUERY_DTYPE), taken, use_ur, ratio)
        assert nm == on and on > 100, (nm, on)
        assert np.array_equal(m[:len(kc)], om), int((m[:len(kc)] != om).sum())
        assert not (m[:len(kc)][taken] >= 0).any() and (m[len(kc):] == -1).all()
    # without earlier matches
    om, on = po.search_local_points(kc, dc, None, None, po.assign_grid(kc, bounds), bounds, q, ratio)
    m, nm = bx.search_local_points(0, q.view(T.PROJ_QUERY_DTYPE), None, False, ratio)
    assert nm == on and np.array_equal(m[:len(kc)], om)
    bx.close()


def test_search_through_the_single_image_handles():
    """The reference-shaped path of INTEGRATION.md 2d / 2e: two ORBextractor handles (left / right), orbx_stereo_match, then the grid and
    the projection search on the batch-of-one behind the left handle (orbx_extractor_batch)."""
    w, h, args = 1241, 376, (2000, 1.2, 8, 20, 7)
    rng = np.random.default_rng(5)
    st = T.StereoORBextractor(*args)
    (ml, kl, dl), _ = st(po.synth_v1(w, h, 31), po.synth_v1(w, h, 31, 16, 0x5bd1e995))
    ur, _ = st.stereo_match(MB, MBF)
    hb = st.left.batch
    hb.assign_grid(1)
    pts, octv, nobs, pd, sf = points_from(args, w, h, 31, 5, rng)
    bounds = (0.0, float(w), 0.0, float(h))
    q = po.make_proj_queries(pts, octv, nobs, pd, sf, 15.0, MBF, 0)
    om, on = po.search_by_projection(kl, dl, ur, po.assign_grid(kl, bounds), bounds, q, True)
    m, nm = hb.search_by_projection(0, q.view(T.PROJ_QUERY_DTYPE), True, True)
    assert nm == on and on > 300 and np.array_equal(m[:len(kl)], om)
    start, items = hb.download_grid(1)
    os_, oi = po.assign_grid(kl, bounds)
    assert np.array_equal(start[0], os_) and np.array_equal(items[0, :os_[-1]], oi)
    st.close()


@pytest.mark.parametrize("seed,th,dist,check", [(41, 10.0, 100, True), (42, 3.0, 64, True), (43, 10.0, 100, False)])
def test_relocalisation_search_matches_oracle(seed, th, dist, check):
    """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) behind its projection: every assigned keypoint blocks, keypoints
    that already hold a point are an input, no right-coordinate gate."""
    w, h, args = 1241, 376, (2000, 1.2, 8, 20, 7)
    rng = np.random.default_rng(seed)
    bx = T.BatchExtractor(w, h, 2, *args)
    kc, dc, _ = make_case(bx, w, h, seed, 5, rng)
    pts, octv, nobs, pd, sf = points_from(args, w, h, seed, 5, rng)
    q = np.zeros(len(pts), T.PROJ_QUERY_DTYPE)
    q["u"], q["v"], q["angle"] = pts[:, 0], pts[:, 1], pts[:, 3]
    q["radius"] = np.float32(th) * sf[octv]
    q["min_level"], q["max_level"], q["blocks"] = octv - 1, octv + 1, 1
    q["desc"] = pd
    taken = rng.random(len(kc)) < 0.3
    bounds = (0.0, float(w), 0.0, float(h))
    om, on = po.search_by_projection(kc, dc, None, po.assign_grid(kc, bounds), bounds, q.view(po.PROJ_QUERY_DTYPE), check, dist, taken)
    m, nm = bx.search_keyframe_points(0, q, taken, check, dist)
    assert nm == on and on > 20 and np.array_equal(m[:len(kc)], om)
    assert not (m[:len(kc)][taken] >= 0).any()
    bx.close()

"""The two-grid (fisheye stereo, Frame::Nleft != -1) branches of the projection searches on the device — ORBmatcher.cc:1803-1868 (frame to frame)
and :153-218 (local map) — against the CPU oracle, which is pinned against the reference's own function bodies on the SAME cases
(tests/test_oracle_vs_ref.py::test_two_grid_*, inputs from tests/two_grid_cases.py).  mvpMapPoints of both images (as sub-query indices) and
the returned match count must be identical."""
import numpy as np
import pytest

import tc2li_slam_b200 as T
import two_grid_cases as tg
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def resident_frame(seed, kind, extra_images=0):
    il, ir = tg.stereo_images(seed, kind)
    bx = T.BatchExtractor(tg.W, tg.H, 2 + extra_images, *tg.ARGS)
    imgs = [il, ir] + [po.synth_v1(tg.W, tg.H, 900 + i) for i in range(extra_images)]
    bx.upload(np.stack(imgs))
    bx.run(len(imgs))
    bx.assign_grid(len(imgs), tg.BOUNDS)
    kp, de, cnt, _ = bx.download(2)
    left, right = (kp[0, :cnt[0]], de[0, :cnt[0]]), (kp[1, :cnt[1]], de[1, :cnt[1]])
    sf = po.OracleExtractor(*tg.ARGS).tables()["scale_factors"]
    return bx, sf, left, right


def with_grid(side):
    return side + (po.assign_grid(side[0], tg.BOUNDS),)


@pytest.mark.parametrize("seed,th,mode,check,kind", [(21, 15.0, 0, True, "plain"), (22, 7.0, 0, True, "plain"), (23, 15.0, 1, True, "plain"), (24, 15.0, 2, False, "plain"),
                                                    (25, 30.0, 0, True, "plain"), (26, 7.0, 0, True, "band"), (27, 15.0, 0, False, "band")])
def test_two_grid_frame_to_frame_search_matches_oracle(seed, th, mode, check, kind):
    bx, sf, left, right = resident_frame(seed, kind)
    pts, octv, nobs, pd = tg.frame_to_frame_points(seed, sf, *tg.last_frame(seed))
    q, side, _ = po.make_proj_queries_stereo(pts, octv, nobs, pd, sf, th, mode, tg.DX_R, tg.DY_R)
    ol, orr, on = po.search_by_projection_stereo(with_grid(left), with_grid(right), tg.BOUNDS, q, side, check)
    for _ in range(2):
        ml, mr, nm = bx.search_by_projection_stereo(0, 1, q.view(T.PROJ_QUERY_DTYPE), side, check)
        assert nm == on and on > 300, (nm, on)
        assert np.array_equal(ml[:len(ol)], ol) and np.array_equal(mr[:len(orr)], orr), (int((ml[:len(ol)] != ol).sum()), int((mr[:len(orr)] != orr).sum()))
        assert (ml[len(ol):] == -1).all() and (mr[len(orr):] == -1).all()
    assert (ol >= 0).sum() > 100 and (orr >= 0).sum() > 100
    bx.close()


@pytest.mark.parametrize("seed,th,ratio,far,kind", [(31, 1.0, 0.8, False, "plain"), (32, 3.0, 0.8, False, "plain"), (33, 5.0, 0.9, True, "plain"), (34, 15.0, 0.6, False, "plain"),
                                                   (35, 15.0, 0.8, False, "periodic"), (36, 15.0, 0.9, True, "periodic")])
def test_two_grid_local_map_search_matches_oracle(seed, th, ratio, far, kind):
    bx, sf, left, right = resident_frame(seed, kind)
    nl, nr = len(left[0]), len(right[0])
    taken0, l2r, r2l, pts, lv, lvr, nobs, flags, pd = tg.local_map_points(seed, nl, nr, kind)
    q, side, _ = po.make_local_queries_stereo(pts, lv, lvr, nobs, flags, pd, sf, th, far, 40.0)
    tl, tr = taken0[:nl] == 1, taken0[nl:] == 1
    ol, orr, on = po.search_local_points_stereo(with_grid(left), with_grid(right), tl, tr, l2r, r2l, tg.BOUNDS, q, side, ratio)
    ml, mr, nm = bx.search_local_points_stereo(0, 1, q.view(T.PROJ_QUERY_DTYPE), side, tl, tr, l2r, r2l, ratio)
    assert nm == on and on > 100, (nm, on)
    assert np.array_equal(ml[:nl], ol) and np.array_equal(mr[:nr], orr), (int((ml[:nl] != ol).sum()), int((mr[:nr] != orr).sum()))
    # without links / without taken keypoints (NULL arrays)
    ol, orr, on = po.search_local_points_stereo(with_grid(left), with_grid(right), None, None, None, None, tg.BOUNDS, q, side, ratio)
    ml, mr, nm = bx.search_local_points_stereo(0, 1, q.view(T.PROJ_QUERY_DTYPE), side, nn_ratio=ratio)
    assert nm == on and np.array_equal(ml[:nl], ol) and np.array_equal(mr[:nr], orr)
    # only one of each pair of arrays
    ol, orr, on = po.search_local_points_stereo(with_grid(left), with_grid(right), None, tr, l2r, None, tg.BOUNDS, q, side, ratio)
    ml, mr, nm = bx.search_local_points_stereo(0, 1, q.view(T.PROJ_QUERY_DTYPE), side, None, tr, l2r, None, ratio)
    assert nm == on and np.array_equal(ml[:nl], ol) and np.array_equal(mr[:nr], orr)
    bx.close()


def test_frame_images_need_not_be_neighbours_and_single_grid_still_works():
    """Right image at index 3, left at index 1 of a four-image handle; afterwards the single-grid search on the same handle (shared scratch buffers)."""
    seed = 28
    il, ir = tg.stereo_images(seed, "plain")
    bx = T.BatchExtractor(tg.W, tg.H, 4, *tg.ARGS)
    bx.upload(np.stack([po.synth_v1(tg.W, tg.H, 901), il, po.synth_v1(tg.W, tg.H, 902), ir]))
    bx.run(4)
    bx.assign_grid(4, tg.BOUNDS)
    kp, de, cnt, _ = bx.download(4)
    left, right = (kp[1, :cnt[1]], de[1, :cnt[1]]), (kp[3, :cnt[3]], de[3, :cnt[3]])
    sf = po.OracleExtractor(*tg.ARGS).tables()["scale_factors"]
    pts, octv, nobs, pd = tg.frame_to_frame_points(seed, sf, *tg.last_frame(seed))
    q, side, _ = po.make_proj_queries_stereo(pts, octv, nobs, pd, sf, 15.0, 0, tg.DX_R, tg.DY_R)
    ol, orr, on = po.search_by_projection_stereo(with_grid(left), with_grid(right), tg.BOUNDS, q, side, True)
    ml, mr, nm = bx.search_by_projection_stereo(1, 3, q.view(T.PROJ_QUERY_DTYPE), side, True)
    assert nm == on and np.array_equal(ml[:len(ol)], ol) and np.array_equal(mr[:len(orr)], orr)
    q1 = po.make_proj_queries(pts, octv, nobs, pd, sf, 15.0, 47.9, 0)
    om, on1 = po.search_by_projection(left[0], left[1], None, po.assign_grid(left[0], tg.BOUNDS), tg.BOUNDS, q1, True)
    m, nm1 = bx.search_by_projection(1, q1.view(T.PROJ_QUERY_DTYPE), False, True)
    assert nm1 == on1 and np.array_equal(m[:len(om)], om)
    # malformed side arrays are rejected
    bad = side.copy()
    bad[0] = 1
    with pytest.raises(T.OrbxError):
        bx.search_by_projection_stereo(1, 3, q.view(T.PROJ_QUERY_DTYPE), bad, True)
    with pytest.raises(T.OrbxError):
        bx.search_by_projection_stereo(1, 1, q.view(T.PROJ_QUERY_DTYPE), side, True)
    bx.close()

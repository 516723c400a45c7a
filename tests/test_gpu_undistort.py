"""Frame::UndistortKeyPoints (Frame.cc:777-810) on the device: orbx_batch_undistort_keypoints against the CPU oracle, which is bit-identical to
cv2.undistortPoints (tests/test_oracle_vs_cv2.py::test_undistort_points_matches_cv2), and the feature grid / area queries / projection search on the
undistorted coordinates (Frame::mvKeysUn) against the oracle's grid and search on the same coordinates."""
import numpy as np
import pytest

import tc2li_slam_b200 as T
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
W, H, ARGS = 752, 480, (1200, 1.2, 8, 20, 7)
CAM = np.array([458.654, 457.296, 367.215, 248.375], np.float32)                               # EuRoC cam0 (config/.../EuRoC.yaml)
DIST = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)


def resident(n=3):
    imgs = np.stack([po.synth_v1(W, H, 70 + i) for i in range(n)])
    bx = T.BatchExtractor(W, H, n, *ARGS)
    bx.upload(imgs)
    bx.run(n)
    kp, de, cnt, _ = bx.download(n)
    return bx, kp, de, cnt


@pytest.mark.parametrize("dist,new_cam", [(DIST, None), (np.r_[DIST, np.float32(0.01)], np.array([450, 450, 370, 240], np.float32)),
                                          (np.array([-0.2, 0.05, 0.001, -0.001, 0.01, 0.02, -0.01, 0.005], np.float32), None),
                                          (np.array([-0.2, 0.05, 0.001, -0.001, 0.01, 0.02, -0.01, 0.005, 1e-3, -1e-3, 5e-4, 2e-4], np.float32), None)])
def test_undistorted_coordinates_match_oracle(dist, new_cam):
    bx, kp, _, cnt = resident()
    bx.undistort_keypoints(3, CAM, dist, new_cam)
    xy = bx.download_undistorted(3)
    for i in range(3):
        pts = np.stack([kp[i, :cnt[i]]["x"], kp[i, :cnt[i]]["y"]], 1)
        want = po.undistort_points(pts, CAM, dist, new_cam)
        assert np.array_equal(xy[i, :cnt[i]].view(np.uint32), want.view(np.uint32)), i
        assert np.abs(want - pts).max() > 1.0   # the distortion is not a no-op
    bx.close()


def test_zero_first_coefficient_means_no_undistortion():
    bx, kp, _, cnt = resident(1)
    bx.undistort_keypoints(1, CAM, np.zeros(4, np.float32))      # mDistCoef.at<float>(0) == 0.0: mvKeysUn = mvKeys
    with pytest.raises(T.OrbxError):
        bx.download_undistorted(1)
    bx.assign_grid(1)
    start, items = bx.download_grid(1)
    os_, oi = po.assign_grid(kp[0, :cnt[0]], (0.0, float(W), 0.0, float(H)))
    assert np.array_equal(start[0], os_) and np.array_equal(items[0, :os_[-1]], oi)
    with pytest.raises(T.OrbxError):
        bx.undistort_keypoints(1, CAM, np.array([0.1, 0.2, 0.3], np.float32))
    bx.close()


def test_grid_queries_and_search_use_the_undistorted_keypoints():
    bx, kp, de, cnt = resident(2)
    bx.undistort_keypoints(2, CAM, DIST)
    xy = bx.download_undistorted(2)
    # Frame::ComputeImageBounds with distortion: the undistorted image corners (Frame.cc:812-839)
    corners = po.undistort_points(np.array([[0, 0], [W, 0], [0, H], [W, H]], np.float32), CAM, DIST)
    bounds = (float(min(corners[0, 0], corners[2, 0])), float(max(corners[1, 0], corners[3, 0])), float(min(corners[0, 1], corners[1, 1])), float(max(corners[2, 1], corners[3, 1])))
    bx.assign_grid(2, bounds)
    start, items = bx.download_grid(2)
    un = []
    for i in range(2):
        k = kp[i, :cnt[i]].copy()
        k["x"], k["y"] = xy[i, :cnt[i], 0], xy[i, :cnt[i], 1]
        un.append(k)
        os_, oi = po.assign_grid(k, bounds)
        assert np.array_equal(start[i], os_) and np.array_equal(items[i, :os_[-1]], oi), i
        raw_s, _ = po.assign_grid(kp[i, :cnt[i]], bounds)
        assert not np.array_equal(raw_s, os_)   # the distorted coordinates would give another grid
    rng = np.random.default_rng(4)
    q = np.zeros(300, T.AREA_QUERY_DTYPE)
    q["image"] = rng.integers(0, 2, 300)
    q["x"], q["y"], q["r"] = rng.uniform(bounds[0], bounds[1], 300), rng.uniform(bounds[2], bounds[3], 300), rng.choice([5, 15, 40], 300)
    q["min_level"], q["max_level"] = -1, -1
    counts, idx = bx.features_in_area(q, 2000)
    for j, rec in enumerate(q):
        k = un[rec["image"]]
        ref = po.features_in_area(k, po.assign_grid(k, bounds), bounds, rec["x"], rec["y"], rec["r"], -1, -1)
        assert counts[j] == len(ref) and np.array_equal(idx[j, :counts[j]], ref), j
    # frame-to-frame search: the frame's own (undistorted) keypoints as map points
    k0 = un[0]
    sf = po.OracleExtractor(*ARGS).tables()["scale_factors"]
    pq = np.zeros(len(k0), po.PROJ_QUERY_DTYPE)
    pq["u"], pq["v"], pq["angle"] = k0["x"] + np.float32(1.5), k0["y"] - np.float32(0.75), k0["angle"]
    pq["radius"] = np.float32(7.0) * sf[k0["octave"]]
    pq["ur"], pq["blocks"] = -1.0, 1
    pq["min_level"], pq["max_level"] = k0["octave"] - 1, k0["octave"] + 1
    pq["desc"] = de[0, :cnt[0]]
    om, on = po.search_by_projection(k0, de[0, :cnt[0]], None, po.assign_grid(k0, bounds), bounds, pq, True)
    m, nm = bx.search_by_projection(0, pq.view(T.PROJ_QUERY_DTYPE), False, True)
    assert nm == on and on > 500 and np.array_equal(m[:cnt[0]], om)
    # a new run drops the undistorted coordinates: the grid is built from mvKeys again
    bx.run(2)
    bx.assign_grid(2, bounds)
    start2, _ = bx.download_grid(2)
    assert np.array_equal(start2[0], po.assign_grid(kp[0, :cnt[0]], bounds)[0])
    bx.close()

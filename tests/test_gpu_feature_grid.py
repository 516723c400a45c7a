"""Frame::AssignFeaturesToGrid / Frame::GetFeaturesInArea on the device (the feature-grid part) against the
CPU oracle, which is pinned against the reference's own function bodies (tests/test_oracle_vs_ref.py::test_feature_grid_*).
Cell lists (CSR of Frame::mGrid) and query results must be identical, order included."""
import numpy as np
import pytest

import tc2li_slam_b200 as T
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def make_queries(rng, kps_per_image, bounds, n):
    q = np.zeros(n, T.AREA_QUERY_DTYPE)
    levels = [(-1, -1), (0, 0), (2, 5), (3, -1), (0, 7), (-1, 2)]
    for i in range(n):
        img = int(rng.integers(0, len(kps_per_image)))
        k = kps_per_image[img]
        if i % 3 == 0 and len(k):  # around a keypoint with integer offsets: |dx| == r ties of the strict compare
            j = int(rng.integers(0, len(k)))
            x, y = float(k["x"][j]) + float(rng.integers(-3, 4)), float(k["y"][j]) + float(rng.integers(-3, 4))
            r = float(rng.integers(1, 40))
        else:
            x, y = float(rng.uniform(bounds[0] - 60, bounds[1] + 60)), float(rng.uniform(bounds[2] - 60, bounds[3] + 60))
            r = float(rng.choice([0.5, 3, 7.5, 15, 40, 100, 2000]))
        q[i] = (img, x, y, r) + levels[i % 6]
    return q


@pytest.mark.parametrize("w,h,args,bounds", [(1241, 376, (2000, 1.2, 8, 20, 7), None), (752, 480, (1200, 1.2, 8, 20, 7), None),
                                              (1241, 376, (2000, 1.2, 8, 20, 7), (-13.25, 1250.5, -9.75, 380.125)), (640, 200, (500, 1.2, 6, 20, 7), None),
                                              (1920, 1080, (4000, 1.2, 8, 20, 7), None)])
def test_grid_and_area_queries_match_oracle(w, h, args, bounds):
    n = 4
    imgs = np.stack([po.synth_v1(w, h, 40 + i, 16 * (i & 1), 0x5bd1e995 * (i & 1)) for i in range(n)])
    bx = T.BatchExtractor(w, h, n, *args)
    bx.upload(imgs)
    bx.run(n)
    bx.assign_grid(n, bounds)
    kp, _, cnt, _ = bx.download(n)
    start, items = bx.download_grid(n)
    bnd = bounds or (0.0, float(w), 0.0, float(h))
    kps = [kp[i, :cnt[i]] for i in range(n)]
    grids = []
    for i in range(n):
        os_, oi = po.assign_grid(kps[i], bnd)
        assert np.array_equal(start[i], os_), i
        assert np.array_equal(items[i, :os_[-1]], oi), i
        assert os_[-1] > 0.9 * cnt[i]
        grids.append((os_, oi))
    rng = np.random.default_rng(w + h)
    q = make_queries(rng, kps, bnd, 600)
    cap = 4200
    counts, idx = bx.features_in_area(q, cap)
    for i, rec in enumerate(q):
        ref = po.features_in_area(kps[rec["image"]], grids[rec["image"]], bnd, rec["x"], rec["y"], rec["r"], rec["min_level"], rec["max_level"])
        assert counts[i] == len(ref), (i, rec)
        assert np.array_equal(idx[i, :counts[i]], ref), (i, rec)
    assert (counts > 0).sum() > 100
    # a small capacity truncates the list but still reports the full count
    c2, i2 = bx.features_in_area(q, 8)
    assert np.array_equal(c2, counts)
    for i in range(len(q)):
        m = min(int(counts[i]), 8)
        assert np.array_equal(i2[i, :m], idx[i, :m])
    bx.close()


def test_grid_edge_cases():
    w, h, args = 640, 200, (500, 1.2, 6, 20, 7)
    bx = T.BatchExtractor(w, h, 2, *args)
    imgs = np.stack([np.full((h, w), 128, np.uint8), po.synth_v1(w, h, 3)])  # image 0: no keypoints at all
    bx.upload(imgs)
    bx.run(2)
    with pytest.raises(T.OrbxError):  # queries before the grid exists
        bx.features_in_area([(0, 10.0, 10.0, 5.0, -1, -1)], 4)
    with pytest.raises(T.OrbxError):
        bx.assign_grid(2, (0.0, 0.0, 0.0, 10.0))
    bx.assign_grid(2)
    start, items = bx.download_grid(2)
    assert not start[0].any()
    kp, _, cnt, _ = bx.download(2)
    os_, oi = po.assign_grid(kp[1, :cnt[1]], (0.0, float(w), 0.0, float(h)))
    assert np.array_equal(start[1], os_) and np.array_equal(items[1, :os_[-1]], oi)
    counts, idx = bx.features_in_area([(0, 100.0, 100.0, 50.0, -1, -1), (1, -500.0, 100.0, 20.0, -1, -1), (1, 100.0, 5000.0, 20.0, -1, -1),
                                       (1, 320.0, 100.0, 1e6, -1, -1)], 1024)
    assert list(counts[:3]) == [0, 0, 0] and counts[3] == os_[-1]
    assert np.array_equal(idx[3, :counts[3]], oi)  # the whole grid in (column, row, push_back) order
    with pytest.raises(T.OrbxError):  # image without a grid
        bx.features_in_area([(2, 1.0, 1.0, 1.0, -1, -1)], 4)
    bx.close()


def test_grid_large_batch_properties():
    """256 resident KITTI-shaped images: every keypoint lands in exactly the cell PosInGrid names, lists are ascending, and a
    whole-image query returns every assigned keypoint once."""
    w, h, n = 1241, 376, 256
    bx = T.BatchExtractor(w, h, n)
    bx.synth(1000, n)
    bx.run(n)
    bx.assign_grid(n)
    kp, _, cnt, _ = bx.download(n)
    start, items = bx.download_grid(n)
    inv_w, inv_h = np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)
    for i in range(0, n, 17):
        k = kp[i, :cnt[i]]
        px = np.floor(np.abs(k["x"] * inv_w) + np.float32(0.5)).astype(np.int64)  # round half away from zero (coordinates are >= 0 here)
        py = np.floor(np.abs(k["y"] * inv_h) + np.float32(0.5)).astype(np.int64)
        ok = (px < 64) & (py < 48)
        cell = px * 48 + py
        assert start[i, -1] == ok.sum()
        for c in np.unique(cell[ok])[::37]:
            lst = items[i, start[i, c]:start[i, c + 1]]
            assert np.array_equal(lst, np.nonzero(ok & (cell == c))[0])
    q = np.zeros(n, T.AREA_QUERY_DTYPE)
    q["image"] = np.arange(n)
    q["x"], q["y"], q["r"], q["min_level"], q["max_level"] = w / 2, h / 2, 1e5, -1, -1
    counts, idx = bx.features_in_area(q, bx.K)
    assert np.array_equal(counts, start[:n, -1])
    for i in range(0, n, 31):
        assert np.array_equal(idx[i, :counts[i]], items[i, :counts[i]])
    bx.close()

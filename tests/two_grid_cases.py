"""Inputs of the two-grid (fisheye stereo, Frame::Nleft != -1) search tests, shared by the oracle-vs-reference tests (CPU) and the
device-vs-oracle tests (GPU).  Everything is derived from synthetic IMAGES, so the device path can extract the same frames."""
import numpy as np

from oracle import pyoracle as po

W, H, ARGS = 752, 480, (1200, 1.2, 8, 20, 7)
BOUNDS = (0.0, float(W), 0.0, float(H))
DX_R, DY_R = -16.0, 0.5   # where the right camera sees a point, relative to the left projection


def stereo_images(seed, kind="plain"):
    """plain: a scene and the same scene 16 px to the side.  band: the left image has a featureless vertical band (left windows there are
    EMPTY while the right image has the features: ORBmatcher.cc:1744-1745 must cancel the right search).  periodic: both images repeat
    every 48 px (wide left windows hold several identical candidates: the ratio test of ORBmatcher.cc:133-134 fires and must cancel the right search)."""
    left, right = po.synth_v1(W, H, seed), po.synth_v1(W, H, seed, 16, 0x5bd1e995)
    if kind == "band":
        left = left.copy()
        left[:, 300:430] = 128
    elif kind == "periodic":
        wide = po.synth_v1(W + 64, H, seed)
        tile = wide[:, 100:148]
        rep = np.tile(tile, (1, (W + 64) // 48 + 2))
        left, right = np.ascontiguousarray(rep[:, :W]), np.ascontiguousarray(rep[:, 16:16 + W])
    return left, right


def frame_to_frame_points(seed, sf, kp, dp):
    """Map points of a 'last frame' (keypoints kp / descriptors dp of the scene 5 px to the side), projected into the current left camera."""
    rng = np.random.default_rng(seed)
    keep = rng.random(len(kp)) < 0.8
    kp, dp = kp[keep], dp[keep]
    n = len(kp)
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = kp["x"] - np.float32(5) + rng.normal(0, 1.5, n).astype(np.float32)
    pts[:, 1] = kp["y"] + rng.normal(0, 1.5, n).astype(np.float32)
    pts[:, 2] = rng.uniform(2, 60, n).astype(np.float32)
    pts[:, 3] = kp["angle"]
    pts[::6, 3] = (kp["angle"][::6] + rng.uniform(0, 360, len(kp[::6])).astype(np.float32)) % np.float32(360)
    inb = (pts[:, 0] >= BOUNDS[0]) & (pts[:, 0] <= BOUNDS[1]) & (pts[:, 1] >= BOUNDS[2]) & (pts[:, 1] <= BOUNDS[3])
    pts = pts[inb]
    octv, pd = kp["octave"].astype(np.int32)[inb], dp[inb].copy()
    nobs = rng.choice([0, 1, 3], len(pts), p=[0.2, 0.3, 0.5]).astype(np.int32)
    pd ^= ((rng.random(pd.shape) < 0.02) * rng.integers(1, 256, pd.shape)).astype(np.uint8)
    return pts, octv, nobs, pd


def last_frame(seed):
    """The frame the map points come from: the plain scene 5 px to the side (oracle extraction: an INPUT of the search, not what is tested)."""
    _, kp, dp = po.OracleExtractor(*ARGS)(po.synth_v1(W, H, seed, 5))
    return kp, dp


def local_map_points(seed, n_left, n_right, kind="plain"):
    """Local-map points for the two-grid search: the keypoints of two shifted views of the scene with their stored projections into both cameras."""
    rng = np.random.default_rng(seed)
    ex = po.OracleExtractor(*ARGS)
    if kind == "periodic":
        views = [(ex(stereo_images(seed, "periodic")[0])[1:], 0.0)]
    else:
        views = [(ex(po.synth_v1(W, H, seed, s))[1:], float(s)) for s in (3, 6)]
    k = np.concatenate([v[0][0] for v in views])
    d = np.concatenate([v[0][1] for v in views])
    sh = np.concatenate([np.full(len(v[0][0]), v[1], np.float32) for v in views])
    n = len(k)
    pts = np.zeros((n, 7), np.float32)
    pts[:, 0] = k["x"] - sh + rng.normal(0, 1.0, n).astype(np.float32)
    pts[:, 1] = k["y"] + rng.normal(0, 1.0, n).astype(np.float32)
    pts[:, 2] = rng.choice([0.9, 0.9985, 0.9999], n).astype(np.float32)
    pts[:, 3] = rng.uniform(1, 80, n).astype(np.float32)
    pts[:, 4] = pts[:, 0] - np.float32(16) + rng.normal(0, 0.7, n).astype(np.float32)
    pts[:, 5] = pts[:, 1] + rng.normal(0, 0.7, n).astype(np.float32)
    pts[:, 6] = rng.choice([0.9, 0.9985, 0.9999], n).astype(np.float32)
    lv = np.clip(k["octave"] + rng.integers(0, 2, n), 0, ARGS[2] - 1).astype(np.int32)
    lvr = np.where(rng.random(n) < 0.1, -1, np.clip(k["octave"] + rng.integers(0, 2, n), 0, ARGS[2] - 1)).astype(np.int32)
    nobs = rng.choice([0, 1, 3], n, p=[0.15, 0.35, 0.5]).astype(np.int32)
    flags = (rng.random(n) < 0.8).astype(np.int32) | ((rng.random(n) < 0.05).astype(np.int32) << 1) | ((rng.random(n) < 0.8).astype(np.int32) << 2)
    pd = d.copy()
    pd ^= ((rng.random(pd.shape) < 0.03) * rng.integers(1, 256, pd.shape)).astype(np.uint8)
    order = rng.permutation(n)
    taken0 = rng.choice([0, 1, 2], n_left + n_right, p=[0.7, 0.2, 0.1]).astype(np.uint8)
    m = min(n_left, n_right) // 3   # mvLeftToRightMatch / mvRightToLeftMatch: a partial one-to-one pairing
    li, ri = rng.permutation(n_left)[:m], rng.permutation(n_right)[:m]
    l2r, r2l = np.full(n_left, -1, np.int32), np.full(n_right, -1, np.int32)
    l2r[li], r2l[ri] = ri, li
    return taken0, l2r, r2l, pts[order], lv[order], lvr[order], nobs[order], flags[order], pd[order]

"""ctypes front-end of the CPU oracle (oracle/orb_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; the product
package (tc2li_slam_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liborb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "orb_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "liborb_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orbo_create.restype = C.c_void_p
        L.orbo_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orbo_destroy.argtypes = [C.c_void_p]
        L.orbo_set_blur_taps.argtypes = [C.c_void_p, C.c_void_p]
        L.orbo_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orbo_num_keypoints.argtypes = [C.c_void_p]
        L.orbo_get_keypoints.argtypes = [C.c_void_p, C.c_void_p]
        L.orbo_get_descriptors.argtypes = [C.c_void_p, C.c_void_p]
        L.orbo_get_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orbo_level_size.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orbo_get_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orbo_get_blurred.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orbo_num_candidates.argtypes = [C.c_void_p, C.c_int]
        L.orbo_get_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orbo_num_fallback_cells.argtypes = [C.c_void_p, C.c_int]
        L.orbo_num_selected.argtypes = [C.c_void_p, C.c_int]
        L.orbo_get_selected.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orbo_stereo_match.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orbo_assign_grid.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orbo_features_in_area.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_float] * 7 + [C.c_int, C.c_int, C.c_void_p, C.c_int]
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.orbo_search_by_projection_stereo.argtypes = [vp, ci, vp, vp, vp, vp, ci, vp, vp, vp, cf, cf, cf, cf, vp, vp, ci, ci, ci, vp, vp]
        L.orbo_search_local_points_stereo.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp, ci, vp, vp, vp, vp, vp, cf, cf, cf, cf, vp, vp, ci, cf, ci, vp, vp]
        L.orbo_search_by_projection.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orbo_search_local_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.orbo_lapping_match.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orbo_cvt_gray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orbo_remap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orbo_synth_v1.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_uint32]
        L.orbo_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orbo_blur.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.orbo_fast.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orbo_fast_atan2.restype = C.c_float
        L.orbo_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orbo_sincos.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
        L.orbo_std_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def synth_v1(w, h, seed, xoff=0, noise_xor=0):
    """orbx-synth-v1.  Right image of a stereo pair: xoff=16, noise_xor=0x5bd1e995."""
    out = np.empty((h, w), np.uint8)
    lib().orbo_synth_v1(_p(out), w, h, seed & 0xFFFFFFFF, xoff, noise_xor & 0xFFFFFFFF)
    return out


def resize(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orbo_resize(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def input_resize(image, dw, dh, rgb=None):
    """cv::resize(image, (dw, dh)) as System::TrackStereoLidar feeds its inputs (System.cc:250-252; INTER_LINEAR, per channel), followed —
    for colour input when `rgb` is given — by the cvtColor of Tracking::Grab* (Tracking.cc:1646-1671).  image: uint8 [h, w] or [h, w, 3|4]."""
    image = np.ascontiguousarray(image, np.uint8)
    if image.ndim == 2:
        return resize(image, dw, dh)
    out = np.stack([resize(np.ascontiguousarray(image[:, :, c]), dw, dh) for c in range(image.shape[2])], axis=2)
    return out if rgb is None else cvt_gray(out, rgb)


def blur(src, taps=(18, 34, 48, 56, 48, 34, 18)):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    t = np.asarray(taps, np.int32)
    lib().orbo_blur(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0], _p(t))
    return dst


def fast(img, t):
    """FAST-9/16 + NMS on a crop; returns int array [n,3] of (x, y, score) in raster order."""
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size // 4 + 16
    out = np.empty((cap, 3), np.int32)
    n = lib().orbo_fast(_p(img), img.shape[1], img.shape[0], img.strides[0], t, _p(out), cap)
    return out[:n].copy()


def fast_atan2(y, x):
    return lib().orbo_fast_atan2(float(y), float(x))


def sincos(angle):
    s, c = C.c_float(), C.c_float()
    lib().orbo_sincos(float(angle), C.byref(s), C.byref(c))
    return s.value, c.value


def std_sort_perm(nkeys, ulx):
    """Permutation produced by libstdc++ std::sort with the reference comparator on (nKeys, UL.x) pairs."""
    a = np.ascontiguousarray(nkeys, np.int32).copy()
    b = np.ascontiguousarray(ulx, np.int32).copy()
    lib().orbo_std_sort_pairs(_p(a), _p(b), len(a))
    return a


def stereo_match(left, right, mb, mbf):
    """Frame::ComputeStereoMatches (Frame.cc:841-1011) on the last images extracted by two OracleExtractors.
    Returns (uRight, depth), float32 [N_left], -1 where unmatched."""
    n = lib().orbo_num_keypoints(left._h)
    u, d = np.empty(n, np.float32), np.empty(n, np.float32)
    lib().orbo_stereo_match(left._h, right._h, float(mb), float(mbf), _p(u), _p(d))
    return u, d


def cvt_gray(image, rgb):
    """cv::cvtColor(image, COLOR_{RGB,BGR}[A]2GRAY) on an 8-bit [h, w, 3|4] image, as the reference's Grab* entry points do
    (Tracking.cc:1646-1671); rgb=True: the first channel is red."""
    image = np.ascontiguousarray(image, np.uint8)
    h, w, c = image.shape
    out = np.empty((h, w), np.uint8)
    lib().orbo_cvt_gray(_p(image), w, h, image.strides[0], c, int(bool(rgb)), _p(out), out.strides[0])
    return out


def remap(image, map_x, map_y):
    """cv::remap(image, map_x, map_y, INTER_LINEAR) with CV_32FC1 maps and the default constant-0 border, as System::TrackStereoLidar
    rectifies its inputs (System.cc:238-248).  image: uint8 [h, w] or [h, w, 3|4]; maps: float32 [dh, dw]."""
    image = np.ascontiguousarray(image, np.uint8)
    map_x, map_y = np.ascontiguousarray(map_x, np.float32), np.ascontiguousarray(map_y, np.float32)
    ch = 1 if image.ndim == 2 else image.shape[2]
    dh, dw = map_x.shape
    out = np.empty((dh, dw) if image.ndim == 2 else (dh, dw, ch), np.uint8)
    lib().orbo_remap(_p(image), image.shape[1], image.shape[0], image.strides[0], ch, _p(map_x), _p(map_y), dw, _p(out), dw, dh, out.strides[0])
    return out


def lapping_match(desc_left, desc_right):
    """Descriptor stage of Frame::ComputeStereoFishEyeMatches (Frame.cc:1162-1187) on the two lapping tails:
    returns (knn int32 [nq, 4] = trainIdx0, distance0, trainIdx1, distance1 with -1 where absent, ok uint8 [nq])."""
    dl = np.ascontiguousarray(desc_left, np.uint8).reshape(-1, 32)
    dr = np.ascontiguousarray(desc_right, np.uint8).reshape(-1, 32)
    out = np.empty((len(dl), 4), np.int32)
    ok = np.empty(len(dl), np.uint8)
    lib().orbo_lapping_match(_p(dl), len(dl), _p(dr), len(dr), _p(out), _p(ok))
    return out, ok


GRID_COLS, GRID_ROWS = 64, 48  # FRAME_GRID_COLS / FRAME_GRID_ROWS (Frame.h:59-60)


def assign_grid(kps, bounds):
    """Frame::AssignFeaturesToGrid (Frame.cc:412-443) for the single-grid, zero-distortion case as CSR: (cell_start int32 [3073],
    items int32 [n_assigned]); cell c = column * 48 + row.  bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY) (Frame::ComputeImageBounds)."""
    kps = np.ascontiguousarray(kps)
    start = np.empty(GRID_COLS * GRID_ROWS + 1, np.int32)
    items = np.empty(max(len(kps), 1), np.int32)
    lib().orbo_assign_grid(_p(kps), len(kps), *[float(v) for v in bounds], _p(start), _p(items))
    return start, items[:start[-1]]


def features_in_area(kps, grid, bounds, x, y, r, min_level=-1, max_level=-1):
    """Frame::GetFeaturesInArea (Frame.cc:687-749): keypoint indices in the reference's order."""
    kps = np.ascontiguousarray(kps)
    start, items = np.ascontiguousarray(grid[0], np.int32), np.ascontiguousarray(grid[1], np.int32)
    out = np.empty(max(len(kps), 1), np.int32)
    n = lib().orbo_features_in_area(_p(kps), _p(start), _p(items), *[float(v) for v in bounds], float(x), float(y), float(r), int(min_level), int(max_level),
                                    _p(out), len(out))
    return out[:n]


PROJ_QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("ur", "<f4"), ("min_level", "<i4"), ("max_level", "<i4"), ("angle", "<f4"),
                             ("blocks", "<i4"), ("desc", "u1", (32,))])  # orbx_proj_query (64 bytes)


def make_proj_queries(points, octaves, nobs, point_desc, scale_factors, th, mbf, mode):
    """What ORBmatcher::SearchByProjection (ORBmatcher.cc:1712-1742) knows about a map point when it starts searching, in the
    reference's float arithmetic.  points: float32 [nq, 4] = (uv(0), uv(1), x3Dc(2), kpLF.angle); mode 0 = default window
    [o-1, o+1], 1 = bForward (o, -1), 2 = bBackward (0, o).  Points behind the camera or outside the image bounds are the caller's to drop."""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 4)
    octv = np.ascontiguousarray(octaves, np.int32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    q = np.zeros(len(pts), PROJ_QUERY_DTYPE)
    q["u"], q["v"], q["angle"] = pts[:, 0], pts[:, 1], pts[:, 3]
    q["radius"] = np.float32(th) * sf[octv]
    invz = (1.0 / pts[:, 2].astype(np.float64)).astype(np.float32)  # const float invzc = 1.0/x3Dc(2)
    q["ur"] = pts[:, 0] - np.float32(mbf) * invz
    q["min_level"] = [octv - 1, octv, np.zeros_like(octv)][mode]
    q["max_level"] = [octv + 1, np.full_like(octv, -1), octv][mode]
    q["blocks"] = np.ascontiguousarray(nobs, np.int32) > 0
    q["desc"] = np.ascontiguousarray(point_desc, np.uint8).reshape(-1, 32)
    return q


def search_by_projection(kps, desc, uright, grid, bounds, queries, check_orientation=True, th_high=100, taken=None):
    """The search of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, ...) from ORBmatcher.cc:1729 on (see orb_oracle.cpp).
    uright = None stands for a frame without right coordinates; with uright = None, `taken` (keypoints that hold a point before the search),
    blocks = 1 in every query and th_high = ORBdist it is the relocalisation overload (ORBmatcher.cc:1898-2019).  Returns (match int32 [n], nmatches)."""
    kps = np.ascontiguousarray(kps)
    desc = np.ascontiguousarray(desc, np.uint8)
    start, items = np.ascontiguousarray(grid[0], np.int32), np.ascontiguousarray(grid[1], np.int32)
    q = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE)
    ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
    t0 = None if taken is None else np.ascontiguousarray(np.asarray(taken) != 0, np.uint8)
    match = np.empty(max(len(kps), 1), np.int32)
    nm = lib().orbo_search_by_projection(_p(kps), len(kps), _p(desc), None if ur is None else _p(ur), None if t0 is None else _p(t0), _p(start), _p(items),
                                         *[float(v) for v in bounds], _p(q), len(q), int(bool(check_orientation)), int(th_high), _p(match))
    return match[:len(kps)], nm


def make_local_queries(points, levels, nobs, flags, point_desc, scale_factors, th, far_points=False, th_far=50.0):
    """The map points ORBmatcher::SearchByProjection(F, vpMapPoints, th, ...) actually searches for (ORBmatcher.cc:58-81), as queries, in the
    reference's float arithmetic.  Returns (queries, index of each query in the input list)."""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 5)
    lv, fl = np.ascontiguousarray(levels, np.int32), np.ascontiguousarray(flags, np.int32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    keep = ((fl & 1) != 0) & ((fl & 2) == 0)
    if far_points:
        keep &= ~(pts[:, 4] > np.float32(th_far))
    sel = np.nonzero(keep)[0]
    q = np.zeros(len(sel), PROJ_QUERY_DTYPE)
    r = np.where(pts[sel, 3] > 0.998, np.float32(2.5), np.float32(4.0)).astype(np.float32)  # RadiusByViewingCos: float viewCos against the double 0.998
    if th != 1.0:
        r = r * np.float32(th)
    q["u"], q["v"], q["ur"] = pts[sel, 0], pts[sel, 1], pts[sel, 2]
    q["radius"] = r * sf[lv[sel]]
    q["min_level"], q["max_level"] = lv[sel] - 1, lv[sel]
    q["blocks"] = np.ascontiguousarray(nobs, np.int32)[sel] > 0
    q["desc"] = np.ascontiguousarray(point_desc, np.uint8).reshape(-1, 32)[sel]
    return q, sel


def search_local_points(kps, desc, uright, taken0, grid, bounds, queries, nn_ratio=0.8, th_high=100):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, ...) (ORBmatcher.cc:52-222) on prepared queries (see orb_oracle.cpp).
    Returns (match int32 [n]: query index assigned by this search or -1, nmatches)."""
    kps = np.ascontiguousarray(kps)
    desc = np.ascontiguousarray(desc, np.uint8)
    start, items = np.ascontiguousarray(grid[0], np.int32), np.ascontiguousarray(grid[1], np.int32)
    q = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE)
    ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
    t0 = None if taken0 is None else np.ascontiguousarray(taken0, np.uint8)
    match = np.empty(max(len(kps), 1), np.int32)
    nm = lib().orbo_search_local_points(_p(kps), len(kps), _p(desc), None if ur is None else _p(ur), None if t0 is None else _p(t0), _p(start), _p(items),
                                        *[float(v) for v in bounds], _p(q), len(q), float(nn_ratio), int(th_high), _p(match))
    return match[:len(kps)], nm


def make_proj_queries_stereo(points, octaves, nobs, point_desc, scale_factors, th, mode, dx_r, dy_r):
    """Two sub-queries per projected map point for a two-grid current frame (ORBmatcher.cc:1729-1868): its window in the left image (side 0) and,
    paired with it (side 1), the window around its projection into the right camera, here (u + dx_r, v + dy_r) in float32.  Returns (queries, side, owner)."""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 4)
    n = len(pts)
    base = make_proj_queries(pts, octaves, nobs, point_desc, scale_factors, th, 0.0, mode)
    q = np.zeros(2 * n, PROJ_QUERY_DTYPE)
    q[0::2] = base
    q[1::2] = base
    q["u"][1::2] = pts[:, 0] + np.float32(dx_r)
    q["v"][1::2] = pts[:, 1] + np.float32(dy_r)
    q["ur"] = 0
    side = np.zeros(2 * n, np.uint8)
    side[1::2] = 1
    return q, side, np.repeat(np.arange(n, dtype=np.int32), 2)


def make_local_queries_stereo(points, levels, levels_r, nobs, flags, point_desc, scale_factors, th, far_points=False, th_far=50.0):
    """Sub-queries of the two-grid local-map search (ORBmatcher.cc:58-81, :154-163).  points: float32 [nq, 7] = (mTrackProjX, mTrackProjY, mTrackViewCos,
    mTrackDepth, mTrackProjXR, mTrackProjYR, mTrackViewCosR); flags bit 0 = mbTrackInView, bit 1 = isBad(), bit 2 = mbTrackInViewR.  Returns (queries, side, owner)."""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 7)
    lv, lvr, fl = np.ascontiguousarray(levels, np.int32), np.ascontiguousarray(levels_r, np.int32), np.ascontiguousarray(flags, np.int32)
    nob = np.ascontiguousarray(nobs, np.int32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    pd = np.ascontiguousarray(point_desc, np.uint8).reshape(-1, 32)
    rows, side, owner = [], [], []
    for i in range(len(pts)):
        in_l, in_r = bool(fl[i] & 1), bool(fl[i] & 4)
        if not in_l and not in_r:
            continue
        if far_points and pts[i, 3] > np.float32(th_far):
            continue
        if fl[i] & 2:
            continue
        if in_l:
            r = np.float32(2.5) if pts[i, 2] > 0.998 else np.float32(4.0)
            if th != 1.0:
                r = r * np.float32(th)
            rows.append((pts[i, 0], pts[i, 1], r * sf[lv[i]], 0.0, lv[i] - 1, lv[i], 0.0, int(nob[i] > 0), pd[i]))
            side.append(0)
            owner.append(i)
        if in_r and lvr[i] != -1:
            r = np.float32(2.5) if pts[i, 6] > 0.998 else np.float32(4.0)
            rows.append((pts[i, 4], pts[i, 5], r * sf[lvr[i]], 0.0, lvr[i] - 1, lvr[i], 0.0, int(nob[i] > 0), pd[i]))
            side.append(1 if in_l else 2)
            owner.append(i)
    q = np.zeros(len(rows), PROJ_QUERY_DTYPE)
    for j, row in enumerate(rows):
        q[j] = row
    return q, np.array(side, np.uint8), np.array(owner, np.int32)


def _side_args(kps, desc, grid):
    kps = np.ascontiguousarray(kps)
    return kps, np.ascontiguousarray(desc, np.uint8), np.ascontiguousarray(grid[0], np.int32), np.ascontiguousarray(grid[1], np.int32)


def search_by_projection_stereo(left, right, bounds, queries, side, check_orientation=True, th_high=100):
    """Two-grid frame-to-frame search (see orb_oracle.cpp).  left / right = (kps, desc, grid).  Returns (match_left, match_right, nmatches): sub-query indices."""
    kl, dl, sl, il = _side_args(*left)
    kr, dr, sr, ir = _side_args(*right)
    q, sd = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE), np.ascontiguousarray(side, np.uint8)
    ml, mr = np.empty(max(len(kl), 1), np.int32), np.empty(max(len(kr), 1), np.int32)
    nm = lib().orbo_search_by_projection_stereo(_p(kl), len(kl), _p(dl), _p(sl), _p(il), _p(kr), len(kr), _p(dr), _p(sr), _p(ir), *[float(v) for v in bounds], _p(q), _p(sd),
                                                len(q), int(bool(check_orientation)), int(th_high), _p(ml), _p(mr))
    return ml[:len(kl)], mr[:len(kr)], nm


def search_local_points_stereo(left, right, taken_l, taken_r, l2r, r2l, bounds, queries, side, nn_ratio=0.8, th_high=100):
    """Two-grid local-map search (see orb_oracle.cpp).  left / right = (kps, desc, grid); taken_* = keypoints holding a point WITH observations before the
    search (or None); l2r / r2l = Frame::mvLeftToRightMatch / mvRightToLeftMatch (or None)."""
    kl, dl, sl, il = _side_args(*left)
    kr, dr, sr, ir = _side_args(*right)
    q, sd = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE), np.ascontiguousarray(side, np.uint8)
    tl = None if taken_l is None else np.ascontiguousarray(np.asarray(taken_l) != 0, np.uint8)
    tr = None if taken_r is None else np.ascontiguousarray(np.asarray(taken_r) != 0, np.uint8)
    a = None if l2r is None else np.ascontiguousarray(l2r, np.int32)
    b = None if r2l is None else np.ascontiguousarray(r2l, np.int32)
    o = lambda x: None if x is None else _p(x)
    ml, mr = np.empty(max(len(kl), 1), np.int32), np.empty(max(len(kr), 1), np.int32)
    nm = lib().orbo_search_local_points_stereo(_p(kl), len(kl), _p(dl), o(tl), _p(sl), _p(il), o(a), _p(kr), len(kr), _p(dr), o(tr), _p(sr), _p(ir), o(b),
                                               *[float(v) for v in bounds], _p(q), _p(sd), len(q), float(nn_ratio), int(th_high), _p(ml), _p(mr))
    return ml[:len(kl)], mr[:len(kr)], nm


def undistort_points(xy, cam, dist, new_cam=None):
    """Frame::UndistortKeyPoints' cv::undistortPoints(mat, mat, K, mDistCoef, cv::Mat(), mK) (Frame.cc:777-810) on float32 points [n, 2];
    cam / new_cam = (fx, fy, cx, cy); dist = 4, 5, 8 or 12 coefficients."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    c = np.ascontiguousarray(cam, np.float32)
    nc = c if new_cam is None else np.ascontiguousarray(new_cam, np.float32)
    d = np.ascontiguousarray(dist, np.float32)
    out = np.empty_like(xy)
    L = lib()
    L.orbo_undistort_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.orbo_undistort_points(_p(xy), len(xy), _p(c), _p(d), len(d), _p(nc), _p(out))
    return out


class OracleExtractor:
    """Mirror of TC2LI_SLAM::ORBextractor (ORBextractor.h:52-118) on the CPU oracle."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self._h = C.c_void_p(lib().orbo_create(nfeatures, scale_factor, nlevels, ini_th, min_th))

    def __del__(self):
        try:
            if self._h:
                lib().orbo_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_blur_taps(self, taps):
        t = np.asarray(taps, np.int32)
        lib().orbo_set_blur_taps(self._h, _p(t))

    def tables(self):
        n = self.nlevels
        sf, isf, s2, is2 = (np.empty(n, np.float32) for _ in range(4))
        quota = np.empty(n, np.int32)
        umax = np.empty(16, np.int32)
        lib().orbo_get_tables(self._h, _p(sf), _p(isf), _p(s2), _p(is2), _p(quota), _p(umax))
        return dict(scale_factors=sf, inv_scale_factors=isf, level_sigma2=s2, inv_level_sigma2=is2, features_per_level=quota, umax=umax)

    def __call__(self, image, lapping=(0, 0)):
        """Returns (mono_index, keypoints[N] structured, descriptors[N,32])."""
        if image is None or image.size == 0:
            return -1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        image = np.asarray(image)
        assert image.dtype == np.uint8 and image.ndim == 2
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        r = lib().orbo_extract(self._h, _p(image), image.shape[1], image.shape[0], image.strides[0], int(lapping[0]), int(lapping[1]))
        if r < 0:
            return r, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        n = lib().orbo_num_keypoints(self._h)
        kps = np.empty(n, KP_DTYPE)
        desc = np.empty((n, 32), np.uint8)
        lib().orbo_get_keypoints(self._h, _p(kps))
        lib().orbo_get_descriptors(self._h, _p(desc))
        return r, kps, desc

    # ---- stage dumps of the last call ----
    def level(self, l):
        w, h = C.c_int(), C.c_int()
        lib().orbo_level_size(self._h, l, C.byref(w), C.byref(h))
        out = np.empty((h.value, w.value), np.uint8)
        lib().orbo_get_level(self._h, l, _p(out))
        return out

    def blurred(self, l):
        out = np.empty_like(self.level(l))
        return out if lib().orbo_get_blurred(self._h, l, _p(out)) else None

    def candidates(self, l):
        """float32 [n,3] (x, y, response); x,y relative to (16,16), cell-row-major then raster order."""
        n = lib().orbo_num_candidates(self._h, l)
        out = np.empty((n, 3), np.float32)
        lib().orbo_get_candidates(self._h, l, _p(out))
        return out

    def fallback_cells(self, l):
        return lib().orbo_num_fallback_cells(self._h, l)

    def selected(self, l):
        n = lib().orbo_num_selected(self._h, l)
        out = np.empty(n, KP_DTYPE)
        lib().orbo_get_selected(self._h, l, _p(out))
        return out


# ---- DBoW2 vocabulary transform (Frame::ComputeBoW) ----
def write_vocabulary_text(path, k, L, scoring, weighting, seed=1, stop_fraction=0.05, trailing_newline=False):
    """A complete k-ary tree of depth L with random 256-bit node descriptors and idf-like weights in the text format the reference loads
    ORBvoc.txt from (TemplatedVocabulary::loadFromTextFile): header 'k L scoring weighting', then per node in creation order
    'parent is_leaf d0 .. d31 weight'.  Nodes are numbered breadth-first like DBoW2's own files; a fraction of the words is
    stopped (weight 0).  Returns the number of nodes written (without the root)."""
    rng = np.random.default_rng(seed)
    lines = [f"{k} {L} {scoring} {weighting}"]
    level, nid, n = [0], 0, 0
    for depth in range(1, L + 1):
        nxt = []
        for parent in level:
            for _ in range(k):
                nid += 1
                d = rng.integers(0, 256, 32)
                leaf = depth == L
                wgt = 0.0
                if leaf:
                    wgt = 0.0 if rng.random() < stop_fraction else float(np.round(rng.uniform(0.5, 12.0), 6))
                    if weighting in (1, 3):   # TF, BINARY: the stored weight of a word is 1
                        wgt = 0.0 if wgt == 0.0 else 1.0
                lines.append(f"{parent} {1 if leaf else 0} " + " ".join(str(int(x)) for x in d) + f" {wgt!r}")
                nxt.append(nid)
                n += 1
        level = nxt
    with open(path, "w") as f:
        f.write("\n".join(lines) + ("\n" if trailing_newline else ""))
    return n


class _BowBase:
    def transform(self, desc, levelsup=4):
        """desc [n, 32] uint8 -> (word ids ascending, values, node ids ascending, node_start [nodes + 1], feature indices)"""
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        words, values = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.float64)
        nodes, start, feats = np.zeros(max(n, 1), np.uint32), np.zeros(n + 2, np.int32), np.zeros(max(n, 1), np.uint32)
        nw, nn = C.c_int(), C.c_int()
        self._fn(self._h, _p(desc), n, levelsup, _p(words), _p(values), C.byref(nw), _p(nodes), _p(start), _p(feats), C.byref(nn))
        return words[:nw.value], values[:nw.value], nodes[:nn.value], start[:nn.value + 1], feats[:start[nn.value]]


class OracleVocabulary(_BowBase):
    """The oracle's restatement of the vendored DBoW2 (orb_oracle.cpp: vocabulary_*)."""

    def __init__(self, path):
        L = lib()
        L.orbo_voc_load_text.restype = C.c_void_p
        L.orbo_voc_load_text.argtypes = [C.c_char_p]
        L.orbo_voc_destroy.argtypes = [C.c_void_p]
        L.orbo_voc_info.argtypes = [C.c_void_p, C.c_void_p]
        L.orbo_voc_export.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.orbo_bow_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7
        self._h = L.orbo_voc_load_text(os.fsencode(path))
        if not self._h:
            raise RuntimeError("cannot load the vocabulary text file")
        self._fn = L.orbo_bow_transform
        info = np.zeros(6, np.int32)
        L.orbo_voc_info(self._h, _p(info))
        self.k, self.L, self.scoring, self.weighting, self.n_nodes, self.n_words = (int(x) for x in info)

    def export(self):
        """(parent, is_word, descriptors, weights) with node 0 = root: the arrays orbx_vocabulary_create takes."""
        n = self.n_nodes
        parent, isw, desc, w = np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros((n, 32), np.uint8), np.zeros(n, np.float64)
        lib().orbo_voc_export(self._h, _p(parent), _p(isw), _p(desc), _p(w))
        return parent, isw, desc, w

    def __del__(self):
        try:
            if self._h:
                lib().orbo_voc_destroy(self._h)
        except Exception:
            pass

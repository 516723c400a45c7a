"""ctypes front-end of oracle/_ref/libref_orb.so: the reference's OWN ORBextractor.cc (compiled where it lies under
/root/reference against the stand-in OpenCV headers of oracle/cvstub).  TEST INFRASTRUCTURE ONLY.

Built by `make -C oracle ref` where /root/reference is mounted; the .so travels to the GPU box with the gpurun snapshot.
Used to validate the restated oracle and as the CPU baseline of bench.py (`cpu_baseline.kind == "reference"`)."""
import ctypes as C
import os

import numpy as np

from .pyoracle import KP_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libref_orb.so")
_lib = None


def available():
    return os.path.exists(_LIB)


def lib():
    global _lib
    if _lib is None:
        C.CDLL(os.path.join(_HERE, "liborb_oracle.so"), mode=C.RTLD_GLOBAL)  # the stub's pixel primitives
        L = C.CDLL(_LIB)
        L.refo_create.restype = C.c_void_p
        L.refo_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.refo_destroy.argtypes = [C.c_void_p]
        L.refo_set_blur_taps.argtypes = [C.c_void_p]
        L.refo_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.refo_get_keypoints.argtypes = [C.c_void_p]
        L.refo_get_descriptors.argtypes = [C.c_void_p]
        L.refo_tables.argtypes = [C.c_void_p] * 5
        L.refo_level_size.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.refo_get_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefExtractor:
    """TC2LI_SLAM::ORBextractor itself (ORBextractor.h:52-118)."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels = nlevels
        self._h = C.c_void_p(lib().refo_create(nfeatures, scale_factor, nlevels, ini_th, min_th))

    def __del__(self):
        try:
            if self._h:
                lib().refo_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def tables(self):
        n = self.nlevels
        a, b, c, d = (np.empty(n, np.float32) for _ in range(4))
        lib().refo_tables(self._h, _p(a), _p(b), _p(c), _p(d))
        return dict(scale_factors=a, inv_scale_factors=b, level_sigma2=c, inv_level_sigma2=d)

    def __call__(self, image, lapping=(0, 0)):
        if image is None or image.size == 0:
            r = lib().refo_extract(self._h, None, 0, 0, 0, int(lapping[0]), int(lapping[1]))
            return r, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        image = np.asarray(image)
        assert image.dtype == np.uint8 and image.ndim == 2
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        r = lib().refo_extract(self._h, _p(image), image.shape[1], image.shape[0], image.strides[0], int(lapping[0]), int(lapping[1]))
        n = lib().refo_num_keypoints()
        kps = np.empty(n, KP_DTYPE)
        desc = np.empty((n, 32), np.uint8)
        if n:
            lib().refo_get_keypoints(_p(kps))
            lib().refo_get_descriptors(_p(desc))
        return r, kps, desc

    def level(self, l, border=0):
        w, h = C.c_int(), C.c_int()
        lib().refo_level_size(self._h, l, C.byref(w), C.byref(h))
        out = np.empty((h.value + 2 * border, w.value + 2 * border), np.uint8)
        lib().refo_get_level(self._h, l, border, _p(out))
        return out


_LIB_STEREO = os.path.join(_HERE, "_ref", "libref_stereo.so")
_slib = None


def stereo_available():
    return os.path.exists(_LIB_STEREO)


def stereo_match(kps_l, desc_l, kps_r, desc_r, pyr_l, pyr_r, scale_factors, inv_scale_factors, mb, mbf):
    """The reference's own Frame::ComputeStereoMatches (Frame.cc:841-1011, built by oracle/tools/gen_ref_stereo.py) on explicit
    inputs: keypoints/descriptors of both images and both pyramids (lists of contiguous uint8 levels)."""
    global _slib
    if _slib is None:
        C.CDLL(os.path.join(_HERE, "liborb_oracle.so"), mode=C.RTLD_GLOBAL)
        _slib = C.CDLL(_LIB_STEREO)
        _slib.refs_stereo_match.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    nl = len(pyr_l)
    pl = [np.ascontiguousarray(p, np.uint8) for p in pyr_l]
    pr = [np.ascontiguousarray(p, np.uint8) for p in pyr_r]
    ptr_l = (C.c_void_p * nl)(*[p.ctypes.data for p in pl])
    ptr_r = (C.c_void_p * nl)(*[p.ctypes.data for p in pr])
    w = np.array([p.shape[1] for p in pl], np.int32)
    h = np.array([p.shape[0] for p in pl], np.int32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    isf = np.ascontiguousarray(inv_scale_factors, np.float32)
    kl, kr = np.ascontiguousarray(kps_l), np.ascontiguousarray(kps_r)
    dl, dr = np.ascontiguousarray(desc_l, np.uint8), np.ascontiguousarray(desc_r, np.uint8)
    u, d = np.empty(len(kl), np.float32), np.empty(len(kl), np.float32)
    _slib.refs_stereo_match(_p(kl), len(kl), _p(dl), _p(kr), len(kr), _p(dr), nl, ptr_l, ptr_r, _p(w), _p(h), _p(sf), _p(isf), float(mb), float(mbf), _p(u), _p(d))
    return u, d


_LIB_GRID = os.path.join(_HERE, "_ref", "libref_grid.so")
_glib = None


def grid_available():
    return os.path.exists(_LIB_GRID)


class RefFrameGrid:
    """The reference's own Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (Frame.cc:412-443, 687-762, built by
    oracle/tools/gen_ref_grid.py) on explicit keypoints: the single-grid case (Nleft == -1, zero distortion) or, with kps_right, the
    two-grid case of the fisheye stereo constructor (Nleft = len(kps), mGrid over mvKeys, mGridRight over mvKeysRight)."""

    def __init__(self, kps, bounds, kps_right=None):
        global _glib
        if _glib is None:
            _glib = C.CDLL(_LIB_GRID)
            _glib.refg_create.restype = C.c_void_p
            _glib.refg_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_float] * 4
            _glib.refg_destroy.argtypes = [C.c_void_p]
            _glib.refg_grid.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
            _glib.refg_features_in_area.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
        self._kps = np.ascontiguousarray(kps)
        self._kr = None if kps_right is None else np.ascontiguousarray(kps_right)
        self._h = _glib.refg_create(_p(self._kps), len(self._kps), None if self._kr is None else _p(self._kr), -1 if self._kr is None else len(self._kr),
                                    *[float(v) for v in bounds])

    def __del__(self):
        if getattr(self, "_h", None):
            _glib.refg_destroy(self._h)
            self._h = None

    def grid(self, right=False):
        n = len(self._kr) if right else len(self._kps)
        start = np.empty(64 * 48 + 1, np.int32)
        items = np.empty(max(n, 1), np.int32)
        _glib.refg_grid(self._h, int(right), _p(start), _p(items))
        return start, items[:start[-1]]

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1, right=False):
        n = len(self._kr) if right else len(self._kps)
        out = np.empty(max(n, 1), np.int32)
        n = _glib.refg_features_in_area(self._h, int(right), float(x), float(y), float(r), int(min_level), int(max_level), _p(out), len(out))
        return out[:n]


_LIB_SEARCH = os.path.join(_HERE, "_ref", "libref_search.so")
_splib = None


def search_available():
    return os.path.exists(_LIB_SEARCH)


def search_by_projection(kps, desc, uright, bounds, scale_factors, mb, mbf, points, octaves, nobs, point_desc, th, tlcz, mono=False, check_orientation=True):
    """The reference's own ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1685-1895, built by
    oracle/tools/gen_ref_search.py) with identity pose algebra: points = float32 [nq, 4] (u, v, z, last-keypoint angle).
    Returns (match int32 [n]: map-point index per current keypoint or -1, nmatches)."""
    global _splib
    if _splib is None:
        _splib = C.CDLL(_LIB_SEARCH)
        _splib.refsp_search.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_float, C.c_float,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]
    kps = np.ascontiguousarray(kps)
    desc = np.ascontiguousarray(desc, np.uint8)
    ur = np.ascontiguousarray(uright, np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 4)
    octv, nob = np.ascontiguousarray(octaves, np.int32), np.ascontiguousarray(nobs, np.int32)
    pd = np.ascontiguousarray(point_desc, np.uint8).reshape(-1, 32)
    match = np.empty(max(len(kps), 1), np.int32)
    nm = _splib.refsp_search(_p(kps), len(kps), _p(desc), _p(ur), *[float(v) for v in bounds], _p(sf), len(sf), float(mb), float(mbf), _p(pts), _p(octv), _p(nob),
                             _p(pd), len(pts), float(th), float(tlcz), int(bool(mono)), int(bool(check_orientation)), _p(match))
    return match[:len(kps)], nm


def search_local_points(kps, desc, uright, taken0, bounds, scale_factors, points, levels, nobs, flags, point_desc, th, nn_ratio, far_points=False, th_far=50.0):
    """The reference's own ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) (ORBmatcher.cc:52-222), single-grid case.
    points: float32 [nq, 5] = (mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos, mTrackDepth); flags bit 0 = mbTrackInView, bit 1 = isBad();
    taken0[k]: 0 = F.mvpMapPoints[k] is NULL, 1 = holds a point with observations, 2 = holds one without.  Returns (match int32 [n], nmatches)."""
    search_by_projection.__doc__  # (loads nothing; the library handle is shared below)
    global _splib
    if _splib is None:
        _splib = C.CDLL(_LIB_SEARCH)
        _splib.refsp_search.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_float, C.c_float,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]
    _splib.refsp_search_local.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p]
    kps = np.ascontiguousarray(kps)
    desc = np.ascontiguousarray(desc, np.uint8)
    ur = np.ascontiguousarray(uright, np.float32)
    t0 = np.ascontiguousarray(taken0, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 5)
    lv, nob, fl = np.ascontiguousarray(levels, np.int32), np.ascontiguousarray(nobs, np.int32), np.ascontiguousarray(flags, np.int32)
    pd = np.ascontiguousarray(point_desc, np.uint8).reshape(-1, 32)
    match = np.empty(max(len(kps), 1), np.int32)
    nm = _splib.refsp_search_local(_p(kps), len(kps), _p(desc), _p(ur), _p(t0), *[float(v) for v in bounds], _p(sf), len(sf), _p(pts), _p(lv), _p(nob), _p(fl), _p(pd),
                                   len(pts), float(th), float(nn_ratio), int(bool(far_points)), float(th_far), _p(match))
    return match[:len(kps)], nm


def search_keyframe_points(kps, desc, taken0, bounds, scale_factors, points, levels, flags, point_desc, th, orb_dist, check_orientation=True):
    """The reference's own ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1898-2019) with identity pose
    algebra.  points: float32 [nq, 4] = (u, v, z, keyframe keypoint angle); flags bit 0 = NULL entry, bit 1 = isBad(), bit 2 = in sAlreadyFound;
    taken0[k] != 0: CurrentFrame.mvpMapPoints[k] is non-NULL before the search.  Returns (match int32 [n], nmatches)."""
    global _splib
    if _splib is None:
        _splib = C.CDLL(_LIB_SEARCH)
        _splib.refsp_search.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_float, C.c_float,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]
    _splib.refsp_search_kf.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p]
    kps = np.ascontiguousarray(kps)
    desc = np.ascontiguousarray(desc, np.uint8)
    t0 = np.ascontiguousarray(taken0, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 4)
    lv, fl = np.ascontiguousarray(levels, np.int32), np.ascontiguousarray(flags, np.int32)
    pd = np.ascontiguousarray(point_desc, np.uint8).reshape(-1, 32)
    match = np.empty(max(len(kps), 1), np.int32)
    nm = _splib.refsp_search_kf(_p(kps), len(kps), _p(desc), _p(t0), *[float(v) for v in bounds], _p(sf), len(sf), _p(pts), _p(lv), _p(fl), _p(pd), len(pts), float(th),
                                int(orb_dist), int(bool(check_orientation)), _p(match))
    return match[:len(kps)], nm


def _search_lib():
    global _splib
    if _splib is None:
        _splib = C.CDLL(_LIB_SEARCH)
        _splib.refsp_search.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_float, C.c_float,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]
    return _splib


def search_by_projection_stereo(kps_l, kps_r, desc_lr, bounds, scale_factors, mb, mbf, points, octaves, nobs, point_desc, th, tlcz, dx_r, dy_r, check_orientation=True):
    """The reference's own ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, false) with a TWO-GRID current frame (Nleft = len(kps_l)): the
    right camera sees a point at (u + dx_r, v + dy_r).  Returns (match int32 [nL + nR]: map-point index or -1, nmatches)."""
    L = _search_lib()
    L.refsp_search_stereo.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p]
    kl, kr = np.ascontiguousarray(kps_l), np.ascontiguousarray(kps_r)
    d = np.ascontiguousarray(desc_lr, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 4)
    octv, nob = np.ascontiguousarray(octaves, np.int32), np.ascontiguousarray(nobs, np.int32)
    pd = np.ascontiguousarray(point_desc, np.uint8).reshape(-1, 32)
    match = np.empty(max(len(kl) + len(kr), 1), np.int32)
    nm = L.refsp_search_stereo(_p(kl), len(kl), _p(kr), len(kr), _p(d), *[float(v) for v in bounds], _p(sf), len(sf), float(mb), float(mbf), _p(pts), _p(octv), _p(nob),
                               _p(pd), len(pts), float(th), float(tlcz), float(dx_r), float(dy_r), int(bool(check_orientation)), _p(match))
    return match[:len(kl) + len(kr)], nm


def search_local_points_stereo(kps_l, kps_r, desc_lr, taken0, l2r, r2l, bounds, scale_factors, points, levels, levels_r, nobs, flags, point_desc, th, nn_ratio,
                               far_points=False, th_far=50.0):
    """The reference's own ORBmatcher::SearchByProjection(F, vpMapPoints, ...) with a TWO-GRID frame.  points: float32 [nq, 7] = (mTrackProjX, mTrackProjY,
    mTrackViewCos, mTrackDepth, mTrackProjXR, mTrackProjYR, mTrackViewCosR); flags bit 0 = mbTrackInView, bit 1 = isBad(), bit 2 = mbTrackInViewR."""
    L = _search_lib()
    L.refsp_search_local_stereo.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p,
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int,
                                           C.c_float, C.c_void_p]
    kl, kr = np.ascontiguousarray(kps_l), np.ascontiguousarray(kps_r)
    d = np.ascontiguousarray(desc_lr, np.uint8)
    t0 = np.ascontiguousarray(taken0, np.uint8)
    a, b = np.ascontiguousarray(l2r, np.int32), np.ascontiguousarray(r2l, np.int32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 7)
    lv, lvr = np.ascontiguousarray(levels, np.int32), np.ascontiguousarray(levels_r, np.int32)
    nob, fl = np.ascontiguousarray(nobs, np.int32), np.ascontiguousarray(flags, np.int32)
    pd = np.ascontiguousarray(point_desc, np.uint8).reshape(-1, 32)
    match = np.empty(max(len(kl) + len(kr), 1), np.int32)
    nm = L.refsp_search_local_stereo(_p(kl), len(kl), _p(kr), len(kr), _p(d), _p(t0), _p(a), _p(b), *[float(v) for v in bounds], _p(sf), len(sf), _p(pts), _p(lv), _p(lvr),
                                     _p(nob), _p(fl), _p(pd), len(pts), float(th), float(nn_ratio), int(bool(far_points)), float(th_far), _p(match))
    return match[:len(kl) + len(kr)], nm


class RefVocabulary:
    """The reference's OWN DBoW2 (oracle/_ref/libref_bow.so, oracle/ref_bow_capi.cpp): loadFromTextFile + transform."""
    _LIB = os.path.join(_HERE, "_ref", "libref_bow.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls._LIB)

    def __init__(self, path):
        L = C.CDLL(self._LIB)
        L.refb_load_text.restype = C.c_void_p
        L.refb_load_text.argtypes = [C.c_char_p]
        L.refb_destroy.argtypes = [C.c_void_p]
        L.refb_words.argtypes = [C.c_void_p]
        L.refb_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7
        self._L = L
        self._h = L.refb_load_text(os.fsencode(path))
        if not self._h:
            raise RuntimeError("the reference could not load the vocabulary text file")
        self.n_words = L.refb_words(self._h)

    def transform(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        words, values = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.float64)
        nodes, start, feats = np.zeros(max(n, 1), np.uint32), np.zeros(n + 2, np.int32), np.zeros(max(n, 1), np.uint32)
        nw, nn = C.c_int(), C.c_int()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self._L.refb_transform(self._h, p(desc), n, levelsup, p(words), p(values), C.byref(nw), p(nodes), p(start), p(feats), C.byref(nn))
        return words[:nw.value], values[:nw.value], nodes[:nn.value], start[:nn.value + 1], feats[:start[nn.value]]

    def __del__(self):
        try:
            if self._h:
                self._L.refb_destroy(self._h)
        except Exception:
            pass

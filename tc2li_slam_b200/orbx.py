"""ctypes binding of liborbx.so (include/orbx.h) and the Python mirror of TC2LI_SLAM::ORBextractor.

`ORBextractor` keeps the reference's class surface (slam_framework/include/ORBextractor.h:52-118): the five-argument
constructor, `__call__(image, mask, vLappingArea)` returning `(monoIndex, keypoints, descriptors)`, the six getters and
`mvImagePyramid`.  `StereoORBextractor` replaces the two extraction threads of Frame::Frame (Frame.cc:139-142) by two CUDA
streams; `BatchExtractor` is the throughput path (many images of one shape per launch set).

Everything here is host-side plumbing over the C ABI; all pixel work happens in the CUDA kernels of csrc/.  There is no
CPU fallback: loading fails loudly when liborbx.so is missing, and constructing an extractor fails when no GPU is usable.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liborbx.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])  # layout of cv::KeyPoint (28 bytes)
assert KP_DTYPE.itemsize == 28

OK, ERR_EMPTY, ERR_TOO_SMALL, ERR_ARG, ERR_CUDA, ERR_CAPACITY, ERR_TOO_LARGE = 0, -1, -2, -3, -4, -5, -6
NUM_STAGES = 6
STAGE_NAMES = ("pyramid", "fast", "blur", "octree", "assemble", "orient_describe")


class OrbxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"orbx error {code}: {msg}")
        self.code = code


_lib = None
_vp = C.c_void_p


def lib():
    """Load liborbx.so (built in-tree by `__graft_entry__.build()` / `make -C tc2li_slam_b200/csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    i, f, sz, u32 = C.c_int, C.c_float, C.c_size_t, C.c_uint32
    sig = {
        "orbx_create": (i, [i, f, i, i, i, i, C.POINTER(_vp)]),
        "orbx_destroy": (None, [_vp]),
        "orbx_get_levels": (i, [_vp]),
        "orbx_get_scale_factor": (f, [_vp]),
        "orbx_get_scale_factors": (i, [_vp, _vp]),
        "orbx_get_inverse_scale_factors": (i, [_vp, _vp]),
        "orbx_get_scale_sigma_squares": (i, [_vp, _vp]),
        "orbx_get_inverse_scale_sigma_squares": (i, [_vp, _vp]),
        "orbx_get_features_per_level": (i, [_vp, _vp]),
        "orbx_max_keypoints": (i, [_vp]),
        "orbx_set_blur_taps": (i, [_vp, _vp]),
        "orbx_extract": (i, [_vp, _vp, i, i, sz, i, i, _vp, _vp, i, _vp, _vp]),
        "orbx_extract_stereo": (i, [_vp, _vp, _vp, _vp, i, i, sz, sz, _vp, _vp, _vp, _vp, i, _vp, _vp, _vp, _vp, i, _vp, _vp]),
        "orbx_get_level_size": (i, [_vp, i, _vp, _vp]),
        "orbx_get_level": (i, [_vp, i, _vp, sz, i]),
        "orbx_extractor_batch": (_vp, [_vp]),
        "orbx_batch_create": (i, [i, f, i, i, i, i, i, i, i, C.POINTER(_vp)]),
        "orbx_batch_destroy": (None, [_vp]),
        "orbx_batch_max_keypoints": (i, [_vp]),
        "orbx_batch_max_images": (i, [_vp]),
        "orbx_batch_set_blur_taps": (i, [_vp, _vp]),
        "orbx_batch_set_stream": (i, [_vp, _vp]),
        "orbx_batch_get_stream": (_vp, [_vp]),
        "orbx_batch_upload": (i, [_vp, _vp, sz, sz, i, i]),
        "orbx_batch_upload_device": (i, [_vp, _vp, sz, sz, i, i]),
        "orbx_batch_upload_color": (i, [_vp, _vp, sz, sz, i, i, i, i]),
        "orbx_batch_upload_raw": (i, [_vp, _vp, sz, sz, i, i, i, i]),
        "orbx_batch_set_rectification": (i, [_vp, _vp, _vp, sz, i, i]),
        "orbx_set_rectification": (i, [_vp, _vp, _vp, i, i, sz, i, i]),
        "orbx_set_input_resize": (i, [_vp, i, i, i, i]),
        "orbx_extract_color": (i, [_vp, _vp, i, i, sz, i, i, i, i, _vp, _vp, i, _vp, _vp]),
        "orbx_extract_stereo_color": (i, [_vp, _vp, _vp, _vp, i, i, sz, sz, i, i, _vp, _vp, _vp, _vp, i, _vp, _vp, _vp, _vp, i, _vp, _vp]),
        "orbx_batch_synth": (i, [_vp, u32, i, i, i]),
        "orbx_batch_download_images": (i, [_vp, i, i, _vp, sz, sz]),
        "orbx_batch_run": (i, [_vp, i, _vp]),
        "orbx_batch_download": (i, [_vp, i, _vp, _vp, _vp, _vp]),
        "orbx_batch_sync": (i, [_vp]),
        "orbx_batch_stereo_match": (i, [_vp, i, f, f]),
        "orbx_batch_download_stereo": (i, [_vp, i, _vp, _vp]),
        "orbx_stereo_match": (i, [_vp, _vp, f, f, _vp, _vp, i, _vp]),
        "orbx_lapping_match": (i, [_vp, _vp, _vp, _vp, i, _vp, _vp]),
        "orbx_batch_lapping_match": (i, [_vp, i]),
        "orbx_batch_download_lapping": (i, [_vp, i, _vp, _vp]),
        "orbx_batch_undistort_keypoints": (i, [_vp, i, _vp, _vp, i, _vp]),
        "orbx_batch_download_undistorted": (i, [_vp, i, _vp]),
        "orbx_batch_assign_grid": (i, [_vp, i, f, f, f, f]),
        "orbx_batch_download_grid": (i, [_vp, i, _vp, _vp]),
        "orbx_batch_features_in_area": (i, [_vp, _vp, i, i, _vp, _vp]),
        "orbx_batch_search_by_projection": (i, [_vp, i, _vp, i, i, i, i, _vp, _vp]),
        "orbx_batch_search_local_points": (i, [_vp, i, _vp, i, i, f, i, _vp, _vp, _vp]),
        "orbx_batch_set_input_resize": (i, [_vp, i, i]),
        "orbx_batch_search_keyframe_points": (i, [_vp, i, _vp, i, i, i, _vp, _vp, _vp]),
        "orbx_batch_run_range": (i, [_vp, i, i, _vp]),
        "orbx_batch_run_pipelined": (i, [_vp, i, _vp, i]),
        "orbx_batch_download_async": (i, [_vp, i, i, _vp, _vp, _vp, _vp]),
        "orbx_batch_check": (i, [_vp]),
        "orbx_event_create": (_vp, [_vp]),
        "orbx_event_destroy": (None, [_vp]),
        "orbx_event_record": (i, [_vp, _vp]),
        "orbx_event_wait": (i, [_vp, _vp]),
        "orbx_event_synchronize": (i, [_vp]),
        "orbx_event_elapsed_ms": (i, [_vp, _vp, _vp]),
        "orbx_batch_get_level_size": (i, [_vp, i, _vp, _vp]),
        "orbx_batch_get_level": (i, [_vp, i, i, _vp, sz, i]),
        "orbx_batch_enable_timing": (i, [_vp, i]),
        "orbx_batch_stage_ms": (i, [_vp, _vp]),
        "orbx_batch_kernel_launches": (i, [_vp]),
        "orbx_batch_debug_blurred": (i, [_vp, i, i, _vp, sz]),
        "orbx_batch_debug_candidates": (i, [_vp, i, i, _vp, i, _vp]),
        "orbx_batch_debug_selected": (i, [_vp, i, i, _vp, i, _vp]),
        "orbx_batch_debug_fallback_cells": (i, [_vp, i, i, _vp]),
        "orbx_batch_search_by_projection_stereo": (i, [_vp, i, i, _vp, _vp, i, i, i, _vp, _vp, _vp]),
        "orbx_batch_search_local_points_stereo": (i, [_vp, i, i, _vp, _vp, i, f, i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
        "orbx_vocabulary_load_text": (i, [i, C.c_char_p, C.POINTER(_vp)]),
        "orbx_vocabulary_create": (i, [i, i, i, i, i, i, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
        "orbx_vocabulary_destroy": (None, [_vp]),
        "orbx_vocabulary_info": (i, [_vp, _vp]),
        "orbx_vocabulary_transform": (i, [_vp, _vp, i, i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
        "orbx_batch_compute_bow": (i, [_vp, _vp, i, i]),
        "orbx_batch_download_bow": (i, [_vp, i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
        "orbx_pool_create": (i, [i, f, i, i, i, _vp, i, i, i, i, i, C.POINTER(_vp)]),
        "orbx_pool_destroy": (None, [_vp]),
        "orbx_pool_max_keypoints": (i, [_vp]),
        "orbx_pool_devices": (i, [_vp]),
        "orbx_pool_extract": (i, [_vp, _vp, sz, sz, i, _vp, _vp, _vp, _vp, _vp]),
        "orbx_pool_last_ms": (i, [_vp, _vp]),
        "orbx_pool_kernel_launches": (i, [_vp]),
        "orbx_alloc_pinned": (_vp, [sz]),
        "orbx_free_pinned": (None, [_vp]),
        "orbx_tables": (i, [i, f, i, i, i, _vp, _vp, _vp, _vp, _vp, _vp]),
        "orbx_level_geometry": (i, [i, f, i, i, i, i, _vp]),
        "orbx_debug_sort": (i, [_vp, i, _vp, i, i]),
        "orbx_debug_sincos": (i, [_vp, _vp, _vp, i]),
        "orbx_debug_atan2": (i, [_vp, _vp, _vp, i]),
        "orbx_version": (C.c_char_p, []),
        "orbx_last_error": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here == the library does not export what include/orbx.h declares
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED = None  # filled lazily by tests: names declared in include/orbx.h


def _p(a):
    return a.ctypes.data_as(_vp)


def _check(code):
    if code != OK:
        raise OrbxError(code, lib().orbx_last_error().decode())
    return code


# --------------------------------------------------------------------------------------------------------------------
# host-only helpers (no GPU)
# --------------------------------------------------------------------------------------------------------------------
def tables(nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    """Constants of ORBextractor's constructor (ORBextractor.cc:383-443)."""
    sf, isf, s2, is2 = (np.empty(nlevels, np.float32) for _ in range(4))
    quota = np.empty(nlevels, np.int32)
    umax = np.empty(16, np.int32)
    _check(lib().orbx_tables(nfeatures, scale_factor, nlevels, ini_th, min_th, _p(sf), _p(isf), _p(s2), _p(is2), _p(quota), _p(umax)))
    return dict(scale_factors=sf, inv_scale_factors=isf, level_sigma2=s2, inv_level_sigma2=is2, features_per_level=quota, umax=umax)


def level_geometry(width, height, level, nfeatures=2000, scale_factor=1.2, nlevels=8):
    out = np.empty(8, np.int32)
    _check(lib().orbx_level_geometry(nfeatures, scale_factor, nlevels, width, height, level, _p(out)))
    return dict(zip(("w", "h", "nCols", "nRows", "wCell", "hCell", "nIni", "capacity"), (int(v) for v in out)))


def debug_sort(keys, force_depth=-1, on_device=False):
    keys = np.ascontiguousarray(keys, np.uint32)
    perm = np.empty(len(keys), np.int32)
    _check(lib().orbx_debug_sort(_p(keys), len(keys), _p(perm), force_depth, int(on_device)))
    return perm


def debug_sincos(angles_deg):
    a = np.ascontiguousarray(angles_deg, np.float32)
    s, c = np.empty_like(a), np.empty_like(a)
    _check(lib().orbx_debug_sincos(_p(a), _p(s), _p(c), len(a)))
    return s, c


def debug_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    _check(lib().orbx_debug_atan2(_p(y), _p(x), _p(out), len(y)))
    return out


class Event:
    """CUDA event recorded on a BatchExtractor's stream (device-side timing / ordering between handles)."""

    def __init__(self, batch):
        self._e = lib().orbx_event_create(batch._h)   # (an event lives on the batch's device)
        if not self._e:
            raise OrbxError(ERR_CUDA, lib().orbx_last_error().decode())

    def record(self, batch):
        _check(lib().orbx_event_record(self._e, batch._h))

    def wait_on(self, batch):
        """Make `batch`'s stream wait for this event."""
        _check(lib().orbx_event_wait(batch._h, self._e))

    def synchronize(self):
        _check(lib().orbx_event_synchronize(self._e))

    def elapsed_ms(self, stop):
        ms = C.c_float(0)
        _check(lib().orbx_event_elapsed_ms(self._e, stop._e, C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            if self._e:
                lib().orbx_event_destroy(self._e)
                self._e = None
        except Exception:
            pass


class PinnedArray:
    """Page-locked host buffer exposed as a numpy array (for overlapped H2D / D2H in the batch path)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._ptr = lib().orbx_alloc_pinned(max(nbytes, 1))
        if not self._ptr:
            raise OrbxError(ERR_CUDA, lib().orbx_last_error().decode())
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._ptr:
            self.array = None
            lib().orbx_free_pinned(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------------------------------
# the ORBextractor mirror
# --------------------------------------------------------------------------------------------------------------------
def _as_gray(image):
    if image is None:
        return None
    image = np.asarray(image)
    if image.size == 0:
        return None
    assert image.dtype == np.uint8 and image.ndim == 2, "ORBextractor expects CV_8UC1 (ORBextractor.cc:1067)"
    if image.strides[1] != 1:
        image = np.ascontiguousarray(image)
    return image


def _as_color(image):
    if image is None:
        return None
    image = np.asarray(image)
    if image.size == 0:
        return None
    assert image.dtype == np.uint8 and image.ndim == 3 and image.shape[2] in (3, 4), "colour input is CV_8UC3 or CV_8UC4 (Tracking.cc:1646-1671)"
    if image.strides[2] != 1 or image.strides[1] != image.shape[2]:
        image = np.ascontiguousarray(image)
    return image


class ORBextractor:
    """Drop-in mirror of TC2LI_SLAM::ORBextractor (ORBextractor.h:52-118) over the CUDA path."""

    HARRIS_SCORE, FAST_SCORE = 0, 1  # ORBextractor.h:56 (unused by the reference as well)

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0):
        self._h = _vp()
        _check(lib().orbx_create(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST), int(device), C.byref(self._h)))
        self.nlevels = int(nlevels)
        self._kcap = 0
        self._kps = self._desc = None

    def close(self):
        if getattr(self, "_h", None):
            lib().orbx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters (ORBextractor.h:70-90) ----
    def GetLevels(self):
        return lib().orbx_get_levels(self._h)

    def GetScaleFactor(self):
        return lib().orbx_get_scale_factor(self._h)

    def _table(self, fn, dtype=np.float32):
        out = np.empty(self.nlevels, dtype)
        _check(fn(self._h, _p(out)))
        return out

    def GetScaleFactors(self):
        return self._table(lib().orbx_get_scale_factors)

    def GetInverseScaleFactors(self):
        return self._table(lib().orbx_get_inverse_scale_factors)

    def GetScaleSigmaSquares(self):
        return self._table(lib().orbx_get_scale_sigma_squares)

    def GetInverseScaleSigmaSquares(self):
        return self._table(lib().orbx_get_inverse_scale_sigma_squares)

    @property
    def mnFeaturesPerLevel(self):
        return self._table(lib().orbx_get_features_per_level, np.int32)

    def set_blur_taps(self, taps):
        t = np.asarray(taps, np.int32)
        _check(lib().orbx_set_blur_taps(self._h, _p(t)))

    def _outputs(self):
        cap = lib().orbx_max_keypoints(self._h)
        if cap > self._kcap:
            self._kps = np.empty(cap, KP_DTYPE)
            self._desc = np.empty((cap, 32), np.uint8)
            self._kcap = cap
        return self._kcap

    def __call__(self, image, mask=None, vLappingArea=(0, 0)):
        """operator() (ORBextractor.cc:1060-1141).  Returns (monoIndex, keypoints[N], descriptors[N,32]);
        monoIndex is -1 for an empty image, like the reference; `mask` is ignored, like the reference."""
        image = _as_gray(image)
        if image is None:
            return -1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        cap = self._outputs()
        n, mono = C.c_int(0), C.c_int(0)
        _check(lib().orbx_extract(self._h, _p(image), image.shape[1], image.shape[0], image.strides[0], int(vLappingArea[0]), int(vLappingArea[1]),
                                  _p(self._kps), _p(self._desc), cap, C.byref(n), C.byref(mono)))
        return mono.value, self._kps[:n.value].copy(), self._desc[:n.value].copy()

    def set_rectification(self, map_x, map_y, src_size):
        """System::TrackStereoLidar's cv::remap (System.cc:238-248) in front of this extractor: float32 maps [H, W] (Settings::M1l/M2l or
        M1r/M2r) sampling raw images of src_size = (src_width, src_height); afterwards __call__ / extract_color take the RAW image.
        map_x=None switches it off."""
        if map_x is None:
            _check(lib().orbx_set_rectification(self._h, None, None, 0, 0, 0, 0, 0))
            return
        mx, my = np.ascontiguousarray(map_x, np.float32), np.ascontiguousarray(map_y, np.float32)
        assert mx.shape == my.shape and mx.ndim == 2
        _check(lib().orbx_set_rectification(self._h, _p(mx), _p(my), mx.shape[1], mx.shape[0], mx.strides[0], int(src_size[0]), int(src_size[1])))

    def set_input_resize(self, size, src_size):
        """System::TrackStereoLidar's cv::resize(im, newImSize) (System.cc:250-252) in front of this extractor: afterwards __call__ /
        extract_color take the RAW image of src_size = (width, height) and extract from its resize to size; size=None switches it off."""
        if size is None:
            _check(lib().orbx_set_input_resize(self._h, 0, 0, 0, 0))
            return
        _check(lib().orbx_set_input_resize(self._h, int(size[0]), int(size[1]), int(src_size[0]), int(src_size[1])))

    def extract_color(self, image, rgb, vLappingArea=(0, 0)):
        """The caller's cvtColor (Tracking.cc:1646-1671: COLOR_RGB[A]2GRAY if rgb else COLOR_BGR[A]2GRAY) on the device, then
        operator().  image: uint8 [H, W, 3|4].  The grey image the reference keeps (mImGray) is self.level(0)."""
        image = _as_color(image)
        if image is None:
            return -1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        cap = self._outputs()
        n, mono = C.c_int(0), C.c_int(0)
        _check(lib().orbx_extract_color(self._h, _p(image), image.shape[1], image.shape[0], image.strides[0], image.shape[2], int(bool(rgb)),
                                        int(vLappingArea[0]), int(vLappingArea[1]), _p(self._kps), _p(self._desc), cap, C.byref(n), C.byref(mono)))
        return mono.value, self._kps[:n.value].copy(), self._desc[:n.value].copy()

    # ---- mvImagePyramid (ORBextractor.h:92): fetched lazily from the device ----
    def level(self, l, border=0):
        w, h = C.c_int(), C.c_int()
        _check(lib().orbx_get_level_size(self._h, l, C.byref(w), C.byref(h)))
        out = np.empty((h.value + 2 * border, w.value + 2 * border), np.uint8)
        _check(lib().orbx_get_level(self._h, l, _p(out), out.strides[0], border))
        return out

    @property
    def mvImagePyramid(self):
        """Levels of the last image; each is the interior view of a buffer that carries the 19-px REFLECT_101 border,
        like the reference's cv::Mat ROIs (ORBextractor.cc:1149-1165)."""
        return [self.level(l, 19)[19:-19, 19:-19] for l in range(self.nlevels)]

    @property
    def batch(self):
        """The batch-of-one behind this handle (stage dumps for tests)."""
        return BatchExtractor._wrap(lib().orbx_extractor_batch(self._h), self.nlevels)


class StereoORBextractor:
    """mpORBextractorLeft / mpORBextractorRight of Tracking (Tracking.cc:626-629) driven as Frame::Frame drives them
    (Frame.cc:139-142), with the two host threads replaced by the two handles' CUDA streams."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0):
        self.left = ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device)
        self.right = ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device)

    def __call__(self, imLeft, imRight, lapLeft=(0, 0), lapRight=(0, 0)):
        imLeft, imRight = _as_gray(imLeft), _as_gray(imRight)
        if imLeft is None or imRight is None:
            e = (-1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8))
            return e, e
        assert imLeft.shape == imRight.shape
        L, R = self.left, self.right
        capL, capR = L._outputs(), R._outputs()
        nl, ml, nr, mr = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        ll = np.asarray(lapLeft, np.int32)
        lr = np.asarray(lapRight, np.int32)
        _check(lib().orbx_extract_stereo(L._h, R._h, _p(imLeft), _p(imRight), imLeft.shape[1], imLeft.shape[0], imLeft.strides[0], imRight.strides[0],
                                         _p(ll), _p(lr), _p(L._kps), _p(L._desc), capL, C.byref(nl), C.byref(ml),
                                         _p(R._kps), _p(R._desc), capR, C.byref(nr), C.byref(mr)))
        return ((ml.value, L._kps[:nl.value].copy(), L._desc[:nl.value].copy()),
                (mr.value, R._kps[:nr.value].copy(), R._desc[:nr.value].copy()))

    def extract_color(self, imLeft, imRight, rgb, lapLeft=(0, 0), lapRight=(0, 0)):
        """Tracking::GrabImageStereoLidar's two cvtColor calls (Tracking.cc:1646-1671) on the device, then both extractions."""
        imLeft, imRight = _as_color(imLeft), _as_color(imRight)
        if imLeft is None or imRight is None:
            e = (-1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8))
            return e, e
        assert imLeft.shape == imRight.shape
        L, R = self.left, self.right
        capL, capR = L._outputs(), R._outputs()
        nl, ml, nr, mr = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        ll = np.asarray(lapLeft, np.int32)
        lr = np.asarray(lapRight, np.int32)
        _check(lib().orbx_extract_stereo_color(L._h, R._h, _p(imLeft), _p(imRight), imLeft.shape[1], imLeft.shape[0], imLeft.strides[0], imRight.strides[0],
                                               imLeft.shape[2], int(bool(rgb)), _p(ll), _p(lr), _p(L._kps), _p(L._desc), capL, C.byref(nl), C.byref(ml),
                                               _p(R._kps), _p(R._desc), capR, C.byref(nr), C.byref(mr)))
        return ((ml.value, L._kps[:nl.value].copy(), L._desc[:nl.value].copy()),
                (mr.value, R._kps[:nr.value].copy(), R._desc[:nr.value].copy()))

    def stereo_match(self, mb, mbf):
        """Frame::ComputeStereoMatches (Frame.cc:841-1011) for the pair last extracted: (mvuRight, mvDepth), -1 = no match."""
        cap = lib().orbx_max_keypoints(self.left._h)
        u, d = np.empty(cap, np.float32), np.empty(cap, np.float32)
        n = C.c_int(0)
        _check(lib().orbx_stereo_match(self.left._h, self.right._h, float(mb), float(mbf), _p(u), _p(d), cap, C.byref(n)))
        return u[:n.value].copy(), d[:n.value].copy()

    def lapping_match(self):
        """Descriptor stage of Frame::ComputeStereoFishEyeMatches (Frame.cc:1162-1187) for the pair last extracted:
        (knn int32 [nq, 4] = trainIdx0, distance0, trainIdx1, distance1 of BFmatcher.knnMatch(k=2) on the lapping tails, -1 where
        absent; ratio_ok uint8 [nq] = Lowe's gate).  Query q is left keypoint monoLeft + q, train t right keypoint monoRight + t."""
        cap = lib().orbx_max_keypoints(self.left._h)
        knn, ok = np.empty((cap, 4), np.int32), np.empty(cap, np.uint8)
        nq, nt = C.c_int(0), C.c_int(0)
        _check(lib().orbx_lapping_match(self.left._h, self.right._h, _p(knn), _p(ok), cap, C.byref(nq), C.byref(nt)))
        return knn[:nq.value].copy(), ok[:nq.value].copy()

    def close(self):
        self.left.close()
        self.right.close()


GRID_COLS, GRID_ROWS = 64, 48  # ORBX_GRID_COLS / ORBX_GRID_ROWS = FRAME_GRID_COLS / FRAME_GRID_ROWS (Frame.h:59-60)
AREA_QUERY_DTYPE = np.dtype([("image", "<i4"), ("x", "<f4"), ("y", "<f4"), ("r", "<f4"), ("min_level", "<i4"), ("max_level", "<i4")])  # orbx_area_query


PROJ_QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("ur", "<f4"), ("min_level", "<i4"), ("max_level", "<i4"), ("angle", "<f4"),
                             ("blocks", "<i4"), ("desc", "u1", (32,))])  # orbx_proj_query


class BatchExtractor:
    """Many images of one shape per launch set (BASELINE.json configs[4]); frames x {left,right} are consecutive images."""

    def __init__(self, width, height, max_images, nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, device=0):
        self._h = _vp()
        _check(lib().orbx_batch_create(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST), int(device),
                                       int(width), int(height), int(max_images), C.byref(self._h)))
        self._own = True
        self.nlevels = int(nlevels)
        self.width, self.height, self.max_images = int(width), int(height), int(max_images)
        self.K = lib().orbx_batch_max_keypoints(self._h)

    @classmethod
    def _wrap(cls, handle, nlevels):
        if not handle:
            return None
        self = cls.__new__(cls)
        self._h = _vp(handle)
        self._own = False
        self.nlevels = nlevels
        self.K = lib().orbx_batch_max_keypoints(self._h)
        self.max_images = lib().orbx_batch_max_images(self._h)
        w, h = C.c_int(0), C.c_int(0)
        lib().orbx_batch_get_level_size(self._h, 0, C.byref(w), C.byref(h))
        self.width, self.height = w.value, h.value
        return self

    def close(self):
        if getattr(self, "_h", None) and self._own:
            lib().orbx_batch_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_blur_taps(self, taps):
        t = np.asarray(taps, np.int32)
        _check(lib().orbx_batch_set_blur_taps(self._h, _p(t)))

    def set_stream(self, cuda_stream):
        _check(lib().orbx_batch_set_stream(self._h, _vp(cuda_stream)))

    @property
    def stream(self):
        return lib().orbx_batch_get_stream(self._h)

    def upload(self, images, first=0):
        """images: uint8 [n, H, W] host array (pinned memory makes the copy asynchronous)."""
        images = np.asarray(images)
        assert images.dtype == np.uint8 and images.ndim == 3 and images.strides[2] == 1
        _check(lib().orbx_batch_upload(self._h, _p(images), images.strides[1], images.strides[0], first, images.shape[0]))

    def upload_color(self, images, rgb, first=0):
        """images: uint8 [n, H, W, 3|4] host array; converted to grey on the device as cv::cvtColor does (Tracking.cc:1646-1671)."""
        images = np.asarray(images)
        assert images.dtype == np.uint8 and images.ndim == 4 and images.shape[3] in (3, 4) and images.strides[3] == 1 and images.strides[2] == images.shape[3]
        _check(lib().orbx_batch_upload_color(self._h, _p(images), images.strides[1], images.strides[0], images.shape[3], int(bool(rgb)), first, images.shape[0]))

    def set_rectification(self, map_x, map_y, src_size):
        """float32 maps [H, W] for all images of the handle (System.cc:238-248); raw images are then uploaded with upload_raw()."""
        mx, my = np.ascontiguousarray(map_x, np.float32), np.ascontiguousarray(map_y, np.float32)
        assert mx.shape == my.shape == (self.height, self.width)
        _check(lib().orbx_batch_set_rectification(self._h, _p(mx), _p(my), mx.strides[0], int(src_size[0]), int(src_size[1])))

    def set_input_resize(self, src_size):
        """System::TrackStereoLidar's cv::resize(im, newImSize) (System.cc:250-252) on the device: upload_raw() then takes raw images of
        src_size = (width, height) and resizes them into level 0; None switches it off."""
        sw, sh = (0, 0) if src_size is None else src_size
        _check(lib().orbx_batch_set_input_resize(self._h, int(sw), int(sh)))

    def upload_raw(self, images, rgb=False, first=0):
        """raw images uint8 [n, srcH, srcW] or [n, srcH, srcW, 3|4]: cv::remap (+ cvtColor) on the device into level 0."""
        images = np.ascontiguousarray(images)
        ch = 1 if images.ndim == 3 else images.shape[3]
        _check(lib().orbx_batch_upload_raw(self._h, _p(images), images.strides[1], images.strides[0], ch, int(bool(rgb)), first, images.shape[0]))

    def upload_device(self, dev_ptr, row_stride, image_stride, first, n):
        _check(lib().orbx_batch_upload_device(self._h, _vp(dev_ptr), row_stride, image_stride, first, n))

    def synth(self, seed0, n, stereo=True, first=0):
        _check(lib().orbx_batch_synth(self._h, seed0 & 0xFFFFFFFF, first, n, int(stereo)))

    def download_images(self, first, n, out):
        """Level-0 images [first, first+n) into a dense uint8 [n, H, W'] host array (W' >= W)."""
        assert out.dtype == np.uint8 and out.ndim == 3 and out.strides[2] == 1 and out.strides[0] == out.strides[1] * out.shape[1]
        _check(lib().orbx_batch_download_images(self._h, first, n, _p(out), out.strides[1], out.strides[0]))

    def run(self, n, lapping=None):
        lp = None
        if lapping is not None:
            lp = np.ascontiguousarray(lapping, np.int32).reshape(n, 2)
        _check(lib().orbx_batch_run(self._h, n, _p(lp) if lp is not None else None))

    def run_range(self, first, n, lapping=None):
        lp = None
        if lapping is not None:
            lp = np.ascontiguousarray(lapping, np.int32).reshape(n, 2)
        _check(lib().orbx_batch_run_range(self._h, first, n, _p(lp) if lp is not None else None))

    def run_pipelined(self, n, chunk, lapping=None):
        """Like run(), in chunks of `chunk` images spread over the handle's internal streams."""
        lp = None
        if lapping is not None:
            lp = np.ascontiguousarray(lapping, np.int32).reshape(n, 2)
        _check(lib().orbx_batch_run_pipelined(self._h, n, _p(lp) if lp is not None else None, int(chunk)))

    def stereo_match(self, n_frames, mb, mbf):
        """Batched Frame::ComputeStereoMatches: image 2f = left, 2f+1 = right (after run()); asynchronous."""
        _check(lib().orbx_batch_stereo_match(self._h, int(n_frames), float(mb), float(mbf)))

    def download_stereo(self, n_frames):
        u, d = np.empty((n_frames, self.K), np.float32), np.empty((n_frames, self.K), np.float32)
        _check(lib().orbx_batch_download_stereo(self._h, int(n_frames), _p(u), _p(d)))
        return u, d

    def lapping_match(self, n_frames):
        """Batched descriptor stage of Frame::ComputeStereoFishEyeMatches: image 2f = left, 2f+1 = right (after run()); asynchronous."""
        _check(lib().orbx_batch_lapping_match(self._h, int(n_frames)))

    def download_lapping(self, n_frames):
        """(knn int32 [n_frames, K, 4], ratio_ok uint8 [n_frames, K]); row f: the N_left - mono_left queries of frame f are valid."""
        knn, ok = np.empty((n_frames, self.K, 4), np.int32), np.empty((n_frames, self.K), np.uint8)
        _check(lib().orbx_batch_download_lapping(self._h, int(n_frames), _p(knn), _p(ok)))
        return knn, ok

    def undistort_keypoints(self, n_images, camera, dist, new_camera=None):
        """Batched Frame::UndistortKeyPoints (Frame.cc:777-810) on the resident keypoints; camera / new_camera = (fx, fy, cx, cy), dist = mDistCoef.  Asynchronous;
        the grid / single-grid searches built afterwards use the undistorted coordinates."""
        c = np.ascontiguousarray(camera, np.float32)
        nc = c if new_camera is None else np.ascontiguousarray(new_camera, np.float32)
        d = None if dist is None else np.ascontiguousarray(dist, np.float32)
        _check(lib().orbx_batch_undistort_keypoints(self._h, int(n_images), _p(c), None if d is None else _p(d), 0 if d is None else len(d), _p(nc)))

    def download_undistorted(self, n_images):
        xy = np.empty((int(n_images), self.K, 2), np.float32)
        _check(lib().orbx_batch_download_undistorted(self._h, int(n_images), _p(xy)))
        return xy

    def assign_grid(self, n_images, bounds=None):
        """Batched Frame::AssignFeaturesToGrid (Frame.cc:412-443) on the resident keypoints (after run()); asynchronous.
        bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY); default: (0, width, 0, height) as Frame::ComputeImageBounds gives without distortion."""
        if bounds is None:
            bounds = (0.0, float(self.width), 0.0, float(self.height))
        _check(lib().orbx_batch_assign_grid(self._h, int(n_images), *[float(v) for v in bounds]))

    def download_grid(self, n_images):
        """(cell_start int32 [n_images, 3073], items int32 [n_images, K]): Frame::mGrid[i][j] = items[cell_start[i * 48 + j] : cell_start[i * 48 + j + 1]]."""
        start, items = np.empty((n_images, GRID_COLS * GRID_ROWS + 1), np.int32), np.empty((n_images, self.K), np.int32)
        _check(lib().orbx_batch_download_grid(self._h, int(n_images), _p(start), _p(items)))
        return start, items

    def features_in_area(self, queries, capacity=256):
        """Batched Frame::GetFeaturesInArea (Frame.cc:687-749).  queries: AREA_QUERY_DTYPE records (image, x, y, r, min_level, max_level)
        or an iterable of such tuples; returns (counts int32 [nq], indices int32 [nq, capacity]) — row q holds the first
        min(counts[q], capacity) indices in the reference's order."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=AREA_QUERY_DTYPE))
        counts, idx = np.empty(len(q), np.int32), np.empty((len(q), int(capacity)), np.int32)
        _check(lib().orbx_batch_features_in_area(self._h, _p(q), len(q), int(capacity), _p(counts), _p(idx)))
        return counts, idx

    def search_by_projection(self, image, queries, use_uright=True, check_orientation=True, th_high=100):
        """The search of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) from ORBmatcher.cc:1729 on, against the resident
        image `image` (after run(), assign_grid() and — for use_uright — stereo_match()).  queries: PROJ_QUERY_DTYPE records, one per
        projected map point in LastFrame order.  Returns (match int32 [K]: query index per keypoint or -1, nmatches)."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=PROJ_QUERY_DTYPE))
        match, nm = np.empty(self.K, np.int32), C.c_int(0)
        _check(lib().orbx_batch_search_by_projection(self._h, int(image), _p(q), len(q), int(bool(use_uright)), int(bool(check_orientation)), int(th_high),
                                                     _p(match), C.byref(nm)))
        return match, nm.value

    def search_keyframe_points(self, image, queries, taken=None, check_orientation=True, orb_dist=100):
        """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1898-2019, Tracking::Relocalization) behind its
        projection; taken: per-keypoint flags (CurrentFrame.mvpMapPoints[k] non-NULL).  Returns (match int32 [K], nmatches)."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=PROJ_QUERY_DTYPE))
        t = None
        if taken is not None:
            t = np.zeros(self.K, np.uint8)
            t[:len(taken)] = np.asarray(taken) != 0
        match, nm = np.empty(self.K, np.int32), C.c_int(0)
        _check(lib().orbx_batch_search_keyframe_points(self._h, int(image), _p(q), len(q), int(bool(check_orientation)), int(orb_dist),
                                                       _p(t) if t is not None else None, _p(match), C.byref(nm)))
        return match, nm.value

    def search_local_points(self, image, queries, taken=None, use_uright=True, nn_ratio=0.8, th_high=100):
        """ORBmatcher::SearchByProjection(F, vpMapPoints, th, ...) (ORBmatcher.cc:52-222, Tracking::SearchLocalPoints) against the resident image
        `image`.  queries: PROJ_QUERY_DTYPE records in vpMapPoints order; taken: per-keypoint flags (F.mvpMapPoints[k] holds a point with
        observations).  Returns (match int32 [K]: query index written by this search or -1, nmatches)."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=PROJ_QUERY_DTYPE))
        t = None
        if taken is not None:
            t = np.zeros(self.K, np.uint8)
            t[:len(taken)] = np.asarray(taken) != 0
        match, nm = np.empty(self.K, np.int32), C.c_int(0)
        _check(lib().orbx_batch_search_local_points(self._h, int(image), _p(q), len(q), int(bool(use_uright)), float(nn_ratio), int(th_high),
                                                    _p(t) if t is not None else None, _p(match), C.byref(nm)))
        return match, nm.value

    def _padded(self, a, dtype, fill):
        if a is None:
            return None
        out = np.full(self.K, fill, dtype)
        out[:len(a)] = a
        return out

    def search_by_projection_stereo(self, image_left, image_right, queries, side, check_orientation=True, th_high=100):
        """Two-grid frame (fisheye stereo, Nleft != -1) variant of search_by_projection (ORBmatcher.cc:1729-1868): sub-queries with side[i] = 0 (left image),
        1 (right image, directly after its left part) — see include/orbx.h.  Returns (match_left [K], match_right [K], nmatches)."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=PROJ_QUERY_DTYPE))
        sd = np.ascontiguousarray(side, np.uint8)
        ml, mr, nm = np.empty(self.K, np.int32), np.empty(self.K, np.int32), C.c_int(0)
        _check(lib().orbx_batch_search_by_projection_stereo(self._h, int(image_left), int(image_right), _p(q), _p(sd), len(q), int(bool(check_orientation)), int(th_high),
                                                            _p(ml), _p(mr), C.byref(nm)))
        return ml, mr, nm.value

    def search_local_points_stereo(self, image_left, image_right, queries, side, taken_left=None, taken_right=None, left_to_right=None, right_to_left=None,
                                   nn_ratio=0.8, th_high=100):
        """Two-grid variant of search_local_points (ORBmatcher.cc:52-222 with F.Nleft != -1); left_to_right / right_to_left = Frame::mvLeftToRightMatch /
        mvRightToLeftMatch.  Returns (match_left [K], match_right [K], nmatches)."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=PROJ_QUERY_DTYPE))
        sd = np.ascontiguousarray(side, np.uint8)
        tl = self._padded(None if taken_left is None else np.asarray(taken_left) != 0, np.uint8, 0)
        tr = self._padded(None if taken_right is None else np.asarray(taken_right) != 0, np.uint8, 0)
        a, b = self._padded(left_to_right, np.int32, -1), self._padded(right_to_left, np.int32, -1)
        o = lambda x: None if x is None else _p(x)
        ml, mr, nm = np.empty(self.K, np.int32), np.empty(self.K, np.int32), C.c_int(0)
        _check(lib().orbx_batch_search_local_points_stereo(self._h, int(image_left), int(image_right), _p(q), _p(sd), len(q), float(nn_ratio), int(th_high), o(tl), o(tr),
                                                           o(a), o(b), _p(ml), _p(mr), C.byref(nm)))
        return ml, mr, nm.value

    def compute_bow(self, vocabulary, n_images, levelsup=4):
        """Batched Frame::ComputeBoW (Frame.cc:768-775) on the resident descriptors (after run()); asynchronous."""
        _check(lib().orbx_batch_compute_bow(self._h, vocabulary._h, int(n_images), int(levelsup)))

    def download_bow(self, n_images):
        """Per image i: (words[:nw], values[:nw], nodes[:nn], node_start[:nn + 1], features[:node_start[nn]]) as ORBVocabulary.transform returns them."""
        n, K = int(n_images), self.K
        words, values, nw = np.empty((n, K), np.uint32), np.empty((n, K), np.float64), np.empty(n, np.int32)
        nodes, start, feats, nn = np.empty((n, K), np.uint32), np.empty((n, K + 1), np.int32), np.empty((n, K), np.uint32), np.empty(n, np.int32)
        _check(lib().orbx_batch_download_bow(self._h, n, _p(words), _p(values), _p(nw), _p(nodes), _p(start), _p(feats), _p(nn)))
        return [(words[i, :nw[i]], values[i, :nw[i]], nodes[i, :nn[i]], start[i, :nn[i] + 1], feats[i, :start[i, nn[i]]]) for i in range(n)]

    def sync(self):
        _check(lib().orbx_batch_sync(self._h))

    def check(self):
        """Synchronise the stream and raise if a device-side capacity error was recorded."""
        _check(lib().orbx_batch_check(self._h))

    def download_async(self, first, n, keypoints, descriptors, counts, mono):
        """Enqueue D2H of images [first, first+n) into the given (ideally pinned) host arrays; no wait."""
        _check(lib().orbx_batch_download_async(self._h, first, n, _p(keypoints) if keypoints is not None else None,
                                               _p(descriptors) if descriptors is not None else None,
                                               _p(counts) if counts is not None else None, _p(mono) if mono is not None else None))

    def download(self, n, keypoints=None, descriptors=None, counts=None, mono=None):
        """Fills (or allocates) host arrays [n,K] / [n,K,32] / [n] / [n]; returns them."""
        if keypoints is None:
            keypoints = np.empty((n, self.K), KP_DTYPE)
        if descriptors is None:
            descriptors = np.empty((n, self.K, 32), np.uint8)
        if counts is None:
            counts = np.empty(n, np.int32)
        if mono is None:
            mono = np.empty(n, np.int32)
        _check(lib().orbx_batch_download(self._h, n, _p(keypoints), _p(descriptors), _p(counts), _p(mono)))
        return keypoints, descriptors, counts, mono

    def download_counts(self, n):
        counts, mono = np.empty(n, np.int32), np.empty(n, np.int32)
        _check(lib().orbx_batch_download(self._h, n, None, None, _p(counts), _p(mono)))
        return counts, mono

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        _check(lib().orbx_batch_get_level_size(self._h, l, C.byref(w), C.byref(h)))
        return w.value, h.value

    def level(self, image, l, border=0):
        w, h = self.level_size(l)
        out = np.empty((h + 2 * border, w + 2 * border), np.uint8)
        _check(lib().orbx_batch_get_level(self._h, image, l, _p(out), out.strides[0], border))
        return out

    def blurred(self, image, l):
        w, h = self.level_size(l)
        out = np.empty((h, w), np.uint8)
        _check(lib().orbx_batch_debug_blurred(self._h, image, l, _p(out), out.strides[0]))
        return out

    def _xyr(self, fn, image, l):
        n = C.c_int(0)
        _check(fn(self._h, image, l, None, 0, C.byref(n)))
        out = np.empty((n.value, 3), np.float32)
        if n.value:
            _check(fn(self._h, image, l, _p(out), n.value, C.byref(n)))
        return out

    def candidates(self, image, l):
        """FAST candidates in the reference's order: float32 [n,3] (x, y relative to (16,16), response)."""
        return self._xyr(lib().orbx_batch_debug_candidates, image, l)

    def selected(self, image, l):
        return self._xyr(lib().orbx_batch_debug_selected, image, l)

    def fallback_cells(self, image, l):
        n = C.c_int(0)
        _check(lib().orbx_batch_debug_fallback_cells(self._h, image, l, C.byref(n)))
        return n.value

    def enable_timing(self, on=True):
        _check(lib().orbx_batch_enable_timing(self._h, int(on)))

    def stage_ms(self):
        ms = np.empty(NUM_STAGES, np.float32)
        _check(lib().orbx_batch_stage_ms(self._h, _p(ms)))
        return dict(zip(STAGE_NAMES, (float(v) for v in ms)))

    @property
    def kernel_launches(self):
        return lib().orbx_batch_kernel_launches(self._h)


class ORBVocabulary:
    """The resident DBoW2 vocabulary (ORBVocabulary = TemplatedVocabulary<FORB::TDescriptor, FORB>, slam_framework/include/ORBVocabulary.h):
    `loadFromTextFile` reads ORBvoc.txt's format (TemplatedVocabulary.h:1350-1433), `transform` is
    transform(features, BowVector, FeatureVector, levelsup) (TemplatedVocabulary.h:1139-1214) for host descriptors; resident
    descriptors go through BatchExtractor.compute_bow."""

    def __init__(self, device=0):
        self.device, self._h = int(device), None

    def _adopt(self, h):
        self.close()
        self._h = h
        info = np.zeros(6, np.int32)
        _check(lib().orbx_vocabulary_info(self._h, _p(info)))
        self.k, self.L, self.scoring, self.weighting, self.n_nodes, self.n_words = (int(v) for v in info)

    def loadFromTextFile(self, path):
        h = _vp()
        _check(lib().orbx_vocabulary_load_text(self.device, os.fsencode(path), C.byref(h)))
        self._adopt(h)
        return True

    def from_arrays(self, k, L, scoring, weighting, parent, is_word, descriptors, weights):
        parent, is_word = np.ascontiguousarray(parent, np.int32), np.ascontiguousarray(is_word, np.uint8)
        descriptors, weights = np.ascontiguousarray(descriptors, np.uint8), np.ascontiguousarray(weights, np.float64)
        h = _vp()
        _check(lib().orbx_vocabulary_create(self.device, int(k), int(L), int(scoring), int(weighting), len(parent), _p(parent), _p(is_word), _p(descriptors),
                                            _p(weights), C.byref(h)))
        self._adopt(h)
        return self

    def size(self):
        return self.n_words

    def transform(self, descriptors, levelsup=4):
        """descriptors [n, 32] uint8 -> (words ascending, values, nodes ascending, node_start [nodes + 1], features)"""
        d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        n = len(d)
        words, values = np.empty(max(n, 1), np.uint32), np.empty(max(n, 1), np.float64)
        nodes, start, feats = np.empty(max(n, 1), np.uint32), np.zeros(n + 1, np.int32), np.empty(max(n, 1), np.uint32)
        nw, nn = C.c_int(), C.c_int()
        _check(lib().orbx_vocabulary_transform(self._h, _p(d), n, int(levelsup), _p(words), _p(values), C.byref(nw), _p(nodes), _p(start), _p(feats), C.byref(nn)))
        return words[:nw.value], values[:nw.value], nodes[:nn.value], start[:nn.value + 1], feats[:start[nn.value]]

    def close(self):
        if self._h:
            lib().orbx_vocabulary_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ExtractorPool:
    """orbx_pool: every GPU of the box behind one call.  Frames are sharded by chunk over the devices, one host
    thread and `handles_per_device` pipelined batch handles per device; host images in, frame-ordered host results out."""

    def __init__(self, width, height, devices, nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, chunk_images=128, handles_per_device=3):
        devs = np.ascontiguousarray(np.asarray(devices, np.int32))
        self._h = _vp()
        _check(lib().orbx_pool_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, _p(devs), len(devs), width, height, chunk_images,
                                      handles_per_device, C.byref(self._h)))
        self.width, self.height, self.n_devices = width, height, len(devs)
        self.K = lib().orbx_pool_max_keypoints(self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib().orbx_pool_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def extract(self, images, lapping=None, keypoints=None, descriptors=None, counts=None, mono=None):
        """images: [n, H, W] uint8 (pinned for speed).  Returns (keypoints [n, K], descriptors [n, K, 32], counts [n], mono [n]);
        pass preallocated (pinned) outputs to keep the downloads asynchronous."""
        n = len(images)
        assert images.dtype == np.uint8 and images.shape[1:] == (self.height, self.width) and images.strides[2] == 1
        keypoints = np.empty((n, self.K), KP_DTYPE) if keypoints is None else keypoints
        descriptors = np.empty((n, self.K, 32), np.uint8) if descriptors is None else descriptors
        counts = np.empty(n, np.int32) if counts is None else counts
        mono = np.empty(n, np.int32) if mono is None else mono
        lap = None if lapping is None else np.ascontiguousarray(lapping, np.int32)
        _check(lib().orbx_pool_extract(self._h, _p(images), images.strides[1], images.strides[0], n, None if lap is None else _p(lap), _p(keypoints),
                                       _p(descriptors), _p(counts), _p(mono)))
        return keypoints, descriptors, counts, mono

    def last_ms(self):
        ms = np.zeros(self.n_devices, np.float32)
        _check(lib().orbx_pool_last_ms(self._h, _p(ms)))
        return ms

    @property
    def kernel_launches(self):
        return lib().orbx_pool_kernel_launches(self._h)

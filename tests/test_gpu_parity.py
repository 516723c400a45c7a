"""Parity of the CUDA path (through the C ABI, tc2li_slam_b200.orbx) against the CPU oracle, stage by stage and end to
end, plus the Appendix C golden hashes.  Bit-exact everywhere: pyramid and blur bytes, FAST candidate sequences, quadtree
selection AND order, float32 angles, descriptors, output permutation.  (north_star tolerance for descriptors is <= 1e-4 of
bits / angles within 1e-3 deg; the tests below demand zero differences.)"""
import hashlib

import numpy as np
import pytest

import tc2li_slam_b200 as T
from oracle import pyoracle as po
from tc2li_slam_b200 import orbx

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def stage_compare(img, args, lapping=(0, 0), taps=None):
    """Runs oracle and CUDA path on one image and compares every stage.  Returns (mono, kps, desc) of the CUDA path."""
    ora = po.OracleExtractor(*args)
    ex = T.ORBextractor(*args)
    if taps is not None:
        ora.set_blur_taps(taps)
        ex.set_blur_taps(taps)
    om, ok, od = ora(img, lapping)
    m, k, d = ex(img, None, lapping)
    b = ex.batch
    nl = args[2]
    for l in range(nl):
        assert np.array_equal(ex.level(l), ora.level(l)), f"pyramid level {l}"
    for l in range(nl):
        oc = ora.candidates(l)
        gc = b.candidates(0, l)
        assert len(gc) == len(oc), f"level {l}: {len(gc)} candidates vs {len(oc)}"
        assert np.array_equal(gc, oc), f"level {l}: candidate sequence"
        assert b.fallback_cells(0, l) == ora.fallback_cells(l), f"level {l}: fallback cells"
    for l in range(nl):
        os_ = ora.selected(l)
        gs = b.selected(0, l)
        assert len(gs) == len(os_), f"level {l}: {len(gs)} selected vs {len(os_)}"
        assert np.array_equal(gs[:, 0] + 16, os_["x"]) and np.array_equal(gs[:, 1] + 16, os_["y"]) and np.array_equal(gs[:, 2], os_["response"]), \
            f"level {l}: selection or its order"
    for l in range(nl):
        ob = ora.blurred(l)
        if ob is not None:
            assert np.array_equal(b.blurred(0, l), ob), f"blurred level {l}"
    assert m == om and len(k) == len(ok)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(k[f], ok[f]), f
    assert np.array_equal(k["angle"].view(np.uint32), ok["angle"].view(np.uint32)), "angles must be bit-identical"
    nbad = int(np.unpackbits(d ^ od).sum())
    assert nbad == 0, f"{nbad} descriptor bits differ"
    ex.close()
    return m, k, d


@pytest.mark.parametrize("name", ["kitti_s0", "kitti_s1_th12", "kitti_imu_s2", "euroc_s3_lap", "mono_s4_lap0_1000", "hd_s5"])
def test_golden_appendix_c(golden_cases, name):
    c = golden_cases[name]
    img = po.synth_v1(c["width"], c["height"], c["seed"])
    m, k, d = stage_compare(img, (c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"]), c["lapping"])
    assert len(k) == c["n"] and m == c["mono_index"]
    assert sha(k) == c["kp_full_sha256"]
    assert sha(d) == c["desc_sha256"]


def test_kitti_stereo_pair_two_streams():
    W, H = 1241, 376
    left = po.synth_v1(W, H, 11)
    right = po.synth_v1(W, H, 11, xoff=16, noise_xor=0x5bd1e995)
    st = T.StereoORBextractor(2000, 1.2, 8, 20, 7)
    ora = po.OracleExtractor(2000, 1.2, 8, 20, 7)
    for _ in range(2):  # second call reuses every buffer
        (ml, kl, dl), (mr, kr, dr) = st(left, right)
        for img, (m, k, d) in ((left, (ml, kl, dl)), (right, (mr, kr, dr))):
            om, ok, od = ora(img)
            assert m == om and k.tobytes() == ok.tobytes() and np.array_equal(d, od)
    # mvImagePyramid with its REFLECT_101 border (ORBextractor.cc:1149-1165)
    lv = st.left.level(3, border=19)
    inner = ora_level = None
    ora(left)
    inner = ora.level(3)
    assert np.array_equal(lv[19:-19, 19:-19], inner)
    assert np.array_equal(lv, np.pad(inner, 19, mode="reflect"))
    st.close()


@pytest.mark.parametrize("kind", ["constant", "low_contrast", "checkerboard", "saturated_noise", "random", "gradient_steps"])
def test_degenerate_inputs(kind):
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
    elif kind == "random":
        img = rng.integers(0, 256, (H, W), dtype=np.uint8)
    else:
        yy, xx = np.mgrid[0:H, 0:W]
        img = ((xx // 5) * 9 + (yy // 7) * 13).astype(np.uint8)
    m, k, d = stage_compare(img, (500, 1.2, 6, 20, 7))
    if kind == "constant":
        assert m == 0 and len(k) == 0 and d.shape == (0, 32)


def test_empty_image_returns_minus_one():
    ex = T.ORBextractor(1000, 1.2, 8, 20, 7)
    m, k, d = ex(np.empty((0, 0), np.uint8))
    assert m == -1 and len(k) == 0
    m, k, d = ex(None)
    assert m == -1
    ex.close()


def test_too_small_image_is_an_error():
    ex = T.ORBextractor(1000, 1.2, 8, 20, 7)
    with pytest.raises(T.OrbxError) as e:
        ex(np.zeros((100, 100), np.uint8))
    assert e.value.code == orbx.ERR_TOO_SMALL
    ex.close()


def test_strided_input_and_blur_taps_variant():
    big = po.synth_v1(800, 300, 21)
    view = big[10:250, 20:700]  # non-contiguous rows
    stage_compare(view, (800, 1.2, 5, 20, 7))
    stage_compare(np.ascontiguousarray(view), (800, 1.2, 5, 15, 5), taps=(18, 34, 49, 55, 49, 34, 18))


def test_other_scale_and_levels():
    img = po.synth_v1(900, 500, 33)
    stage_compare(img, (1500, 1.5, 4, 25, 10), (100, 600))
    stage_compare(img, (700, 1.1, 10, 20, 7))


def test_batch_matches_single_and_oracle():
    W, H, NF = 1241, 376, 4  # 4 stereo frames = 8 images, generated on the device
    b = T.BatchExtractor(W, H, 2 * NF, 2000, 1.2, 8, 20, 7)
    b.synth(100, 2 * NF, stereo=True)
    b.run(2 * NF)
    kps, desc, cnt, mono = b.download(2 * NF)
    ora = po.OracleExtractor(2000, 1.2, 8, 20, 7)
    for i in range(2 * NF):
        seed = 100 + i // 2
        img = po.synth_v1(W, H, seed, 16, 0x5bd1e995) if i & 1 else po.synth_v1(W, H, seed)
        assert np.array_equal(b.level(i, 0), img), "device generator orbx-synth-v1"
        om, ok, od = ora(img)
        assert cnt[i] == len(ok) and mono[i] == om
        assert kps[i, :cnt[i]].tobytes() == ok.tobytes()
        assert np.array_equal(desc[i, :cnt[i]], od)
    # re-run on the same handle with per-image lapping windows
    lap = np.tile(np.array([[0, 1000]], np.int32), (2 * NF, 1))
    b.run(2 * NF, lap)
    kps2, desc2, cnt2, mono2 = b.download(2 * NF)
    om, ok, od = ora(po.synth_v1(W, H, 100), (0, 1000))
    assert mono2[0] == om and kps2[0, :cnt2[0]].tobytes() == ok.tobytes() and np.array_equal(desc2[0, :cnt2[0]], od)
    b.close()


def test_float_routines_on_device():
    rng = np.random.default_rng(1)
    deg = np.concatenate([rng.uniform(0, 360, 200000), np.arange(0, 361, 0.25), [0.0, 1e-6, 45.0, 90.0, 359.99997]]).astype(np.float32)
    s, c = orbx.debug_sincos(deg)
    factor = np.float32(np.pi / 180.0)
    rs = np.empty_like(s)
    rc = np.empty_like(c)
    for i, a in enumerate((deg * factor).astype(np.float32)):
        rs[i], rc[i] = po.sincos(a)
    assert np.array_equal(s.view(np.uint32), rs.view(np.uint32)) and np.array_equal(c.view(np.uint32), rc.view(np.uint32))
    y = rng.integers(-2 ** 20, 2 ** 20, 100000).astype(np.float32)
    x = rng.integers(-2 ** 20, 2 ** 20, 100000).astype(np.float32)
    y[:100] = 0
    x[50:150] = 0
    got = orbx.debug_atan2(y, x)
    want = np.array([po.fast_atan2(a, b) for a, b in zip(y, x)], np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_sort_emulation_on_device():
    rng = np.random.default_rng(2)
    for n in (1, 2, 16, 17, 33, 100, 333, 512, 900):
        nk = rng.integers(2, 9, n)
        ux = rng.integers(0, 40, n) * 30
        keys = ((nk << 12) | ux).astype(np.uint32)
        want = po.std_sort_perm(nk, ux)
        assert np.array_equal(orbx.debug_sort(keys, on_device=True), want), n
        assert np.array_equal(orbx.debug_sort(keys, force_depth=-2, on_device=True), want), ("CTA-parallel emulation", n)
    for trial in range(60):  # the CTA-parallel emulation on tie-heavy and tie-free inputs of many sizes
        n = int(rng.integers(1, 1200))
        nk = rng.integers(2, 2 + int(rng.integers(1, 400)), n)
        ux = rng.integers(0, int(rng.integers(1, 80)), n) * 25
        keys = ((nk << 12) | ux).astype(np.uint32)
        assert np.array_equal(orbx.debug_sort(keys, force_depth=-2, on_device=True), po.std_sort_perm(nk, ux)), (trial, n)


def test_4k_single_frame():
    """BASELINE.json configs[3]: 3840x2160, 4000 features (136 k candidates on level 0 go through the quadtree)."""
    img = po.synth_v1(3840, 2160, 9)
    m, k, d = stage_compare(img, (4000, 1.2, 8, 20, 7))
    assert 4000 <= len(k) <= 4000 + 2 * 8


def test_euroc_stereo_with_lapping_windows():
    """BASELINE.json configs[2]: 752x480 stereo, 1200 features, KB8-style lapping windows left {250,751} / right {0,501}
    (Frame.cc:1090-1091): the in-window keypoints are filled from the back, monoIndex = number outside the window."""
    W, H = 752, 480
    left = po.synth_v1(W, H, 21)
    right = po.synth_v1(W, H, 21, xoff=16, noise_xor=0x5bd1e995)
    st = T.StereoORBextractor(1200, 1.2, 8, 20, 7)
    ora = po.OracleExtractor(1200, 1.2, 8, 20, 7)
    (ml, kl, dl), (mr, kr, dr) = st(left, right, (250, 751), (0, 501))
    for img, lap, (m, k, d) in ((left, (250, 751), (ml, kl, dl)), (right, (0, 501), (mr, kr, dr))):
        om, ok, od = ora(img, lap)
        assert m == om and 0 < m < len(k)
        assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)
        assert np.all((k["x"][:m] < lap[0]) | (k["x"][:m] > lap[1])) and np.all((k["x"][m:] >= lap[0]) & (k["x"][m:] <= lap[1]))
    st.close()


def test_batch_sub_ranges_and_full_size_properties():
    """run_range on sub-ranges of a resident batch equals one run over everything; size-independent properties at the
    benchmark's scale: counts within [nfeatures, nfeatures + 2 * nlevels], coordinates inside the FAST region, octaves sorted."""
    W, H, N = 1241, 376, 24
    b = T.BatchExtractor(W, H, N, 2000, 1.2, 8, 20, 7)
    b.synth(777, N, stereo=True)
    b.run(N)
    k1, d1, c1, m1 = b.download(N)
    for first in range(0, N, 7):
        b.run_range(first, min(7, N - first))
    k2, d2, c2, m2 = b.download(N)
    assert np.array_equal(c1, c2) and np.array_equal(m1, m2)
    for i in range(N):
        assert k1[i, :c1[i]].tobytes() == k2[i, :c2[i]].tobytes() and np.array_equal(d1[i, :c1[i]], d2[i, :c2[i]])
        kk = k1[i, :c1[i]]
        assert 2000 <= c1[i] <= 2016 and m1[i] == c1[i]
        assert np.all(np.diff(kk["octave"]) >= 0) and kk["octave"].min() == 0 and kk["octave"].max() == 7
        sf = T.tables()["scale_factors"][kk["octave"]]
        assert np.all(kk["x"] >= 19 * sf - 1e-3) and np.all(kk["x"] <= (W - 19)) and np.all(kk["y"] >= 19 * sf - 1e-3) and np.all(kk["y"] <= H - 19)
        assert np.all((kk["angle"] >= 0) & (kk["angle"] < 360)) and np.all(kk["class_id"] == -1)
        assert np.all(kk["response"] >= 7) and np.all(kk["response"] <= 254)
    b.close()


def test_random_shapes_and_parameters():
    """Geometry edge cases (narrow last cells, one-row grids, odd pitches): 48 random shapes / parameter sets, final outputs
    compared bit for bit with the oracle (shapes the oracle rejects as too small must be rejected identically)."""
    rng = np.random.default_rng(2024)
    ex_cache = {}
    n_ok = 0
    for trial in range(48):
        w, h = int(rng.integers(90, 900)), int(rng.integers(90, 500))
        if w < 0.55 * h:
            w = h  # portrait shapes have no quadtree root in the reference (nIni == 0)
        args = (int(rng.integers(100, 1500)), float(rng.choice([1.2, 1.2, 1.15, 1.3, 1.5])), int(rng.integers(1, 7)), int(rng.choice([20, 12, 30])), int(rng.choice([7, 5, 12])))
        if args[4] > args[3]:
            args = args[:4] + (args[3],)
        img = po.synth_v1(w, h, 5000 + trial)
        lap = (int(rng.integers(0, w)), int(rng.integers(0, w))) if trial % 3 == 0 else (0, 0)
        om, ok, od = po.OracleExtractor(*args)(img, lap)
        ex = T.ORBextractor(*args)
        if om == -2:
            with pytest.raises(T.OrbxError) as e:
                ex(img, None, lap)
            assert e.value.code == orbx.ERR_TOO_SMALL
        else:
            m, k, d = ex(img, None, lap)
            assert m == om and len(k) == len(ok), (trial, w, h, args)
            assert k.tobytes() == ok.tobytes() and np.array_equal(d, od), (trial, w, h, args)
            n_ok += 1
        ex.close()
    assert n_ok >= 24


def test_large_resident_batch_spot_checks():
    """A resident batch at benchmark scale (1024 device-generated KITTI-shaped images, one launch set, then the pipelined
    multi-stream variant): every image within the count bounds, 16 random images compared bit for bit with the oracle."""
    W, H, N = 1241, 376, 1024
    b = T.BatchExtractor(W, H, N, 2000, 1.2, 8, 20, 7)
    b.synth(31337, N, stereo=True)
    b.run(N)
    kps, desc, cnt, mono = b.download(N)
    assert cnt.min() >= 2000 and cnt.max() <= 2016 and np.array_equal(cnt, mono)
    b.run_pipelined(N, 128)
    kps2, desc2, cnt2, _ = b.download(N)
    assert np.array_equal(cnt, cnt2)
    ora = po.OracleExtractor(2000, 1.2, 8, 20, 7)
    rng = np.random.default_rng(7)
    for i in rng.choice(N, 16, replace=False):
        seed = 31337 + i // 2
        img = po.synth_v1(W, H, seed, 16, 0x5bd1e995) if i & 1 else po.synth_v1(W, H, seed)
        om, ok, od = ora(img)
        for kk, dd, cc in ((kps, desc, cnt), (kps2, desc2, cnt2)):
            assert cc[i] == len(ok) and kk[i, :cc[i]].tobytes() == ok.tobytes() and np.array_equal(dd[i, :cc[i]], od), int(i)
    b.close()


@pytest.mark.parametrize("w", [257, 258, 259, 384, 385, 386, 387, 640, 770, 895, 896])
def test_blur_edges_at_strip_boundaries(w):
    """The blur mirrors the three columns beyond the image edge inside its shared-memory boxes (no margin columns in HBM): widths
    whose last 128-pixel strip holds 1, 2, 3 or all 128 pixels, and widths that make lower levels end at such boundaries; every
    blurred level is compared in full."""
    img = po.synth_v1(w, 150, 900 + w)
    stage_compare(img, (400, 1.2, 4, 20, 7))
    stage_compare(img, (400, 1.2, 2, 20, 7), taps=(18, 34, 49, 55, 49, 34, 18))


@pytest.mark.parametrize("ini,mn", [(150, 130), (128, 127), (200, 90), (127, 7)])
def test_high_fast_thresholds(ini, mn):
    """Thresholds on both sides of 128 (the packed byte compare has a separate path for t >= 128) on a high-contrast image."""
    rng = np.random.default_rng(ini)
    img = (rng.integers(0, 2, (200, 640)) * 255).astype(np.uint8)
    img[:, 320:] = po.synth_v1(640, 200, 5)[:, 320:]
    stage_compare(img, (500, 1.2, 4, ini, mn))

"""orbx_pool (include/orbx.h): frames sharded by chunk over the GPUs of the box, frame-ordered host results.  Every frame is compared
with the CPU oracle; the multi-device cases run when the box has the devices (the round-end box may have one)."""
import ctypes as C

import numpy as np
import pytest

import tc2li_slam_b200 as T
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

ARGS = (600, 1.2, 6, 20, 7)
W, H = 640, 200


def n_devices():
    import torch
    return torch.cuda.device_count()


def frames(n):
    return np.stack([po.synth_v1(W, H, 100 + i, xoff=16 * (i & 1)) for i in range(n)])


def check_all(pool, imgs, lapping=None):
    kps, desc, counts, mono = pool.extract(imgs, lapping)
    ora = po.OracleExtractor(*ARGS)
    for i, img in enumerate(imgs):
        om, ok, od = ora(img, tuple(lapping[i]) if lapping is not None else (0, 0))
        assert counts[i] == len(ok) and mono[i] == om, f"frame {i}: counts"
        assert kps[i, :counts[i]].tobytes() == ok.tobytes(), f"frame {i}: keypoints differ from the oracle"
        assert np.array_equal(desc[i, :counts[i]], od), f"frame {i}: descriptors differ from the oracle"
    return counts


@pytest.mark.parametrize("chunk,handles,n", [(4, 2, 11), (3, 3, 3), (8, 1, 5)])
def test_pool_one_device_every_frame_matches_oracle(chunk, handles, n):
    pool = T.ExtractorPool(W, H, [0], *ARGS, chunk_images=chunk, handles_per_device=handles)
    assert pool.n_devices == 1
    imgs = frames(n)
    c1 = check_all(pool, imgs)
    c2 = check_all(pool, imgs)   # second call reuses every handle
    assert np.array_equal(c1, c2)
    assert pool.kernel_launches > 0 and pool.last_ms()[0] > 0
    pool.close()


def test_pool_lapping_windows_pass_through():
    pool = T.ExtractorPool(W, H, [0], *ARGS, chunk_images=2, handles_per_device=2)
    imgs = frames(5)
    lap = np.array([[100 + 10 * i, 500] for i in range(5)], np.int32)
    check_all(pool, imgs, lap)
    pool.close()


def test_pool_pinned_buffers_and_strided_rows():
    pool = T.ExtractorPool(W, H, [0], *ARGS, chunk_images=4, handles_per_device=2)
    pin = T.PinnedArray((6, H, W + 16), np.uint8)
    imgs = frames(6)
    pin.array[:, :, :W] = imgs
    check_all(pool, pin.array[:, :, :W])
    pool.close()


@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_pool_multi_device_every_frame_matches_oracle(ndev):
    if n_devices() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    pool = T.ExtractorPool(W, H, list(range(ndev)), *ARGS, chunk_images=2, handles_per_device=2)
    imgs = frames(4 * ndev + 3)   # ragged: the last chunk is short, one device gets an extra chunk
    check_all(pool, imgs)
    assert (pool.last_ms() > 0).all()
    pool.close()


def test_pool_rejects_bad_arguments():
    lib = T.lib()
    h = C.c_void_p()
    devs = (C.c_int * 2)(0, 0)
    assert lib.orbx_pool_create(600, 1.2, 6, 20, 7, devs, 2, W, H, 4, 2, C.byref(h)) == T.orbx.ERR_ARG   # a device twice
    assert lib.orbx_pool_create(600, 1.2, 6, 20, 7, devs, 1, W, H, 0, 2, C.byref(h)) == T.orbx.ERR_ARG   # empty chunks
    assert lib.orbx_pool_create(600, 1.2, 6, 20, 7, devs, 1, W, H, 4, 2, C.byref(h)) == 0
    assert lib.orbx_pool_extract(h, None, W, W * H, 3, None, None, None, None, None) == T.orbx.ERR_EMPTY
    lib.orbx_pool_destroy(h)

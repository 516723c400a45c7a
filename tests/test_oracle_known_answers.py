"""The CPU oracle must reproduce the six known-answer cases of SURVEY.md Appendix C (tests/golden/appendix_c.json):
image, pyramid-level hashes, FAST candidate / fallback-cell / kept counts per level, keypoint and descriptor hashes.
Those answers were produced independently of oracle/orb_oracle.cpp (cv2 4.13 primitives + reference control flow)."""
import hashlib

import numpy as np
import pytest

from oracle import pyoracle as po

CASES = ["kitti_s0", "kitti_s1_th12", "kitti_imu_s2", "euroc_s3_lap", "mono_s4_lap0_1000", "hd_s5"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", CASES)
def test_appendix_c(golden_cases, name):
    c = golden_cases[name]
    img = po.synth_v1(c["width"], c["height"], c["seed"])
    assert sha(img)[:16] == c["image_sha256_16"]
    assert list(img[0, :8]) == c["first_pixels"]
    ex = po.OracleExtractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    mono, kps, desc = ex(img, c["lapping"])
    assert [sha(ex.level(l))[:16] for l in range(c["nlevels"])] == c["level_sha256_16"]
    assert [len(ex.candidates(l)) for l in range(c["nlevels"])] == c["candidates"]
    assert [ex.fallback_cells(l) for l in range(c["nlevels"])] == c["fallback_cells"]
    assert [len(ex.selected(l)) for l in range(c["nlevels"])] == c["kept"]
    assert len(kps) == c["n"] and mono == c["mono_index"]
    xyor = np.stack([kps["x"], kps["y"], kps["response"], kps["octave"].astype(np.float32)], axis=1).astype("<f4")
    assert sha(xyor) == c["kp_xyor_sha256"]
    assert sha(kps) == c["kp_full_sha256"]
    assert sha(desc) == c["desc_sha256"]
    if c["lapping"] == [0, 0]:
        fk = c["first_keypoint"]
        assert (kps["x"][0], kps["y"][0], kps["size"][0], kps["response"][0], kps["octave"][0]) == (fk["x"], fk["y"], fk["size"], fk["response"], fk["octave"])
        assert float(kps["angle"][0]) == fk["angle"]
        assert desc[0].tobytes().hex() == c["first_descriptor_hex"]

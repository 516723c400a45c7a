"""Frame::ComputeBoW on the device: orbx_vocabulary_* / orbx_batch_compute_bow against the CPU oracle, which is pinned
against the reference's vendored DBoW2 compiled here (tests/test_oracle_vs_ref.py::test_bow_transform_*), and against golden vectors the
reference's DBoW2 itself produced (tests/golden/bow_golden.json, oracle/tools/gen_bow_golden.py).  Word ids, node ids and feature lists
must be identical; the WordValues are compared as float64 BITS.  The real ORBvoc.txt is absent from the reference checkout
(.MISSING_LARGE_BLOBS): vocabularies are synthetic trees written in its text format, k = 10 / L = 6 (its shape) included."""
import json
import os

import numpy as np
import pytest

import tc2li_slam_b200 as T
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
NAMES = ("words", "values", "nodes", "node_start", "features")


def same(a, b, ctx):
    for x, y, what in zip(a, b, NAMES):
        assert x.dtype == y.dtype or what == "node_start", (what, x.dtype, y.dtype)
        assert np.asarray(x).tobytes() == np.asarray(y, x.dtype).tobytes(), (what,) + tuple(ctx)


def test_golden_vectors_from_the_reference_dbow2(tmp_path):
    with open(os.path.join(os.path.dirname(__file__), "golden", "bow_golden.json")) as f:
        cases = json.load(f)
    for c in cases:
        path = str(tmp_path / "v.txt")
        with open(path, "w") as f:
            f.write(c["vocabulary_text"])
        voc = T.ORBVocabulary()
        voc.loadFromTextFile(path)
        assert (voc.k, voc.L, voc.scoring, voc.weighting) == (c["k"], c["L"], c["scoring"], c["weighting"]) and voc.size() == c["k"] ** c["L"]
        got = voc.transform(np.array(c["descriptors"], np.uint8), c["levelsup"])
        want = (np.array(c["words"], np.uint32), np.array([float.fromhex(v) for v in c["values_hex"]], np.float64), np.array(c["nodes"], np.uint32),
                np.array(c["node_start"], np.int32), np.array(c["features"], np.uint32))
        same(got, want, (c["k"], c["L"], c["scoring"], c["weighting"]))
        voc.close()


@pytest.mark.parametrize("k,L,scoring,weighting", [(10, 6, 0, 0), (10, 4, 1, 0), (6, 3, 0, 1), (4, 5, 5, 2), (3, 4, 2, 3), (9, 3, 5, 0), (2, 8, 4, 1), (20, 2, 3, 0)])
def test_host_descriptors_match_oracle(tmp_path, k, L, scoring, weighting):
    path = str(tmp_path / "voc.txt")
    po.write_vocabulary_text(path, k, L, scoring, weighting, seed=7 * k + L)
    ora = po.OracleVocabulary(path)
    voc = T.ORBVocabulary()
    voc.loadFromTextFile(path)
    assert voc.size() == ora.n_words and voc.n_nodes == ora.n_nodes
    rng = np.random.default_rng(k * L)
    _, _, real = po.OracleExtractor(500, 1.2, 4, 20, 7)(po.synth_v1(640, 200, 7))
    sets = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in (0, 1, 31, 32, 33, 257, 2000, 5000, 8192)] + [real, np.repeat(real[:7], 100, axis=0)]
    for d in sets:
        for levelsup in (0, 4, L - 1, L, L + 3):
            same(voc.transform(d, levelsup), ora.transform(d, levelsup), (len(d), levelsup))
    # the same tree handed over as arrays
    v2 = T.ORBVocabulary().from_arrays(k, L, scoring, weighting, *ora.export())
    same(v2.transform(sets[6], 4), ora.transform(sets[6], 4), ("arrays",))
    with pytest.raises(T.OrbxError):
        voc.transform(rng.integers(0, 256, (8193, 32), dtype=np.uint8), 4)
    voc.close()
    v2.close()


def test_tied_distances_take_the_first_child_and_stopped_words_vanish(tmp_path):
    # every child of a node carries the SAME descriptor: the descent must keep the first one at every level (strict <); all words stopped
    # but a few: features that land on stopped words appear in neither vector
    k, L = 4, 3
    lines = [f"{k} {L} 0 0"]
    level, nid = [0], 0
    for depth in range(1, L + 1):
        nxt = []
        for parent in level:
            for c in range(k):
                nid += 1
                wgt = 0.0 if depth < L or (nid % 3 == 1) else 1.5 + nid / 64.0
                lines.append(f"{parent} {1 if depth == L else 0} " + " ".join(["170"] * 32) + f" {wgt!r}")
                nxt.append(nid)
        level = nxt
    path = str(tmp_path / "ties.txt")
    open(path, "w").write("\n".join(lines))
    ora, voc = po.OracleVocabulary(path), T.ORBVocabulary()
    voc.loadFromTextFile(path)
    d = np.random.default_rng(3).integers(0, 256, (50, 32), dtype=np.uint8)
    got = voc.transform(d, 1)
    same(got, ora.transform(d, 1), ("ties",))
    assert len(got[0]) == 1 and got[0][0] == 0 and len(got[4]) == 50   # first word of the first branch
    voc.close()


@pytest.mark.parametrize("w,h,args,n", [(1241, 376, (2000, 1.2, 8, 20, 7), 6), (752, 480, (1200, 1.2, 8, 20, 7), 4), (640, 200, (500, 1.2, 6, 20, 7), 5)])
def test_resident_descriptors_batch_matches_oracle(tmp_path, w, h, args, n):
    path = str(tmp_path / "voc.txt")
    po.write_vocabulary_text(path, 10, 5, 0, 0, seed=11)
    ora, voc = po.OracleVocabulary(path), T.ORBVocabulary()
    voc.loadFromTextFile(path)
    imgs = np.stack([po.synth_v1(w, h, 60 + i, 16 * (i & 1), 0x5bd1e995 * (i & 1)) for i in range(n)])
    imgs[n - 1] = 128   # an image without keypoints: empty vectors
    bx = T.BatchExtractor(w, h, n, *args)
    bx.upload(imgs)
    bx.run(n)
    bx.compute_bow(voc, n, 4)
    _, desc, cnt, _ = bx.download(n)
    res = bx.download_bow(n)
    assert cnt[n - 1] == 0 and len(res[n - 1][0]) == 0 and len(res[n - 1][2]) == 0
    for i in range(n):
        same(res[i], ora.transform(desc[i, :cnt[i]], 4), (i,))
        assert i == n - 1 or len(res[i][0]) > 0.5 * cnt[i]
    # again with another level for the FeatureVector, fewer images
    bx.compute_bow(voc, 2, 2)
    for i, r in enumerate(bx.download_bow(2)):
        same(r, ora.transform(desc[i, :cnt[i]], 2), (i, "levelsup 2"))
    with pytest.raises(T.OrbxError):
        bx.download_bow(3)
    bx.close()
    voc.close()


def test_single_image_handle_frame_compute_bow(tmp_path):
    """The reference-shaped sequence: extractor(image) -> ComputeBoW on what it produced (Frame.cc:768-775)."""
    path = str(tmp_path / "voc.txt")
    po.write_vocabulary_text(path, 10, 4, 0, 0, seed=5)
    ora, voc = po.OracleVocabulary(path), T.ORBVocabulary()
    voc.loadFromTextFile(path)
    ex = T.ORBextractor(1000, 1.2, 8, 20, 7)
    _, kps, desc = ex(po.synth_v1(752, 480, 9))
    b = ex.batch
    b.compute_bow(voc, 1, 4)
    same(b.download_bow(1)[0], ora.transform(desc, 4), ("single",))
    same(voc.transform(desc, 4), ora.transform(desc, 4), ("host",))
    ex.close()
    voc.close()


def test_bad_vocabulary_files_are_rejected(tmp_path):
    voc = T.ORBVocabulary()
    with pytest.raises(T.OrbxError):
        voc.loadFromTextFile(str(tmp_path / "missing.txt"))
    for text in ("21 3 0 0\n", "10 0 0 0\n", "10 3 6 0\n", "10 3 0 4\n", "10 3 0 0\n", "3 1 0 0\n5 1 " + " ".join(["1"] * 32) + " 1.0\n"):
        p = str(tmp_path / "bad.txt")
        open(p, "w").write(text)
        with pytest.raises(T.OrbxError):
            voc.loadFromTextFile(p)

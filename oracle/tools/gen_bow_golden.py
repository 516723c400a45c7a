"""Golden vectors for the bag-of-words transform: outputs of the REFERENCE's own DBoW2 (oracle/_ref/libref_bow.so = the vendored
slam_framework/Thirdparty/DBoW2 compiled where it lies) on small synthetic vocabularies -> tests/golden/bow_golden.json.
Run in the container that mounts /root/reference (after `make -C oracle`):  python oracle/tools/gen_bow_golden.py"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402  (only its vocabulary WRITER is used here)
from oracle import pyref  # noqa: E402

CASES = [(3, 3, 0, 0, 4, 1), (4, 3, 1, 0, 5, 2), (3, 4, 5, 0, 6, 4), (5, 2, 0, 2, 7, 1), (2, 5, 2, 3, 8, 3), (3, 3, 5, 1, 9, 0)]  # k, L, scoring, weighting, seed, levelsup
out = []
for k, L, s, w, seed, lu in CASES:
    path = tempfile.mktemp(suffix=".txt")
    po.write_vocabulary_text(path, k, L, s, w, seed=seed, stop_fraction=0.1)
    text = open(path).read()
    ref = pyref.RefVocabulary(path)
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (12, 32), dtype=np.uint8)
    d = np.concatenate([base, base[:6], base[:3], rng.integers(0, 256, (20, 32), dtype=np.uint8)])   # repeated descriptors: words seen 1-3 times
    words, values, nodes, start, feats = ref.transform(d, lu)
    out.append(dict(k=k, L=L, scoring=s, weighting=w, levelsup=lu, vocabulary_text=text, descriptors=d.tolist(), words=words.tolist(),
                    values_hex=[float(v).hex() for v in values], nodes=nodes.tolist(), node_start=start.tolist(), features=feats.tolist()))
    os.unlink(path)
with open(os.path.join(ROOT, "tests", "golden", "bow_golden.json"), "w") as f:
    json.dump(out, f)
print("wrote", len(out), "cases")

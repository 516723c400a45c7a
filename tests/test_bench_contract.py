"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line, and the algorithmic byte
model matches SURVEY.md 8(d) for the KITTI shape."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-pairs", "8"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_algorithmic_bytes_model():
    sys.path.insert(0, ROOT)
    import bench
    ab = bench.algorithmic_bytes(2011)
    P, P0, P7 = 1444097, 466616, 36330
    assert ab["fast"] == P and ab["blur"] == 2 * P and ab["pyramid"] == (P - P7) + (P - P0)
    assert ab["total"] == 5 * P - P0 - P7 + 1321 * 2011 == 9374070

"""The reference arm of bench.py (`--impl reference`: the oracle port timed on the host cores, no GPU) prints ONE JSON line
with the keys the measurement contract names; the numbers it reports are the ones it measured (steps honoured, value
consistent with ms_per_step and the sample it describes)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "frames/sec compressed" and d["unit"] == "frames/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32" and d["extrapolated"] is False
    assert "workload" in d["config"] and "sample" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "128x128" in cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    # value = sample pixel-frames / (512 x 512) / seconds: 128 x 128 x 4096 pixel-frames = 256 full-FOV frame equivalents
    assert abs(d["value"] - 256.0 / (d["ms_per_step"] / 1e3)) <= 1e-6 * d["value"]

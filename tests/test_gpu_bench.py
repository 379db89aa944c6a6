"""GPU: bench.py's own arm prints one JSON line with the contract's keys (short run)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"}


def test_own_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-cpu-baseline",
                          "--no-precise"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["metric"] == "visual_tokens_per_sec" and d["n_gpus"] == 1 and d["value"] > 50_000
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0.2 < r["frac"] < 1.3 and r["achieved"] > 0 and r["peak"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 16384 * 1176 and e["d2h_bytes_per_step"] > 4096 * 3584
    assert d["gpu_launches"] > 100 * d["steps"]
    assert "workload" in d["config"] and "sm_mhz" in d["clocks"]
    assert d["saturation_count"] == 0                       # the f16 headline never clamped an activation
    side = d["configs"]
    for name in ("cfg1_16f_448", "cfg3_one_clip_per_gpu_64f_448", "cfg4_128f_336x588", "cfg5_ragged_16_clips"):
        assert side[name].get("ms", 0) > 0 and side[name]["tokens_per_sec"] > 50_000, (name, side[name])
    assert "library_bar" in d

"""bench.py's own JSON line carries every key of the contract (small workload, so it runs in seconds)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--batch", "16", "--hw", "64",
                          "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-800:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "unet_deglare_512x512_images_per_sec" and d["unit"] == "images/s" and d["n_gpus"] == 1
    assert d["steps"] == 3 and d["warmup"] == 3 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "fp16" and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["value"] > 0 and abs(d["value"] - 16 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    assert d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == "images/s" and e["h2d_bytes_per_step"] == 16 * 64 * 64 * 4 and e["d2h_bytes_per_step"] == 16 * 64 * 64 * 4
    assert e["max_abs_vs_device_path"] == 0.0
    assert d["e2e_u8"]["h2d_bytes_per_step"] == 16 * 64 * 64 and d["train_step"]["value"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert "traffic" in r and set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])

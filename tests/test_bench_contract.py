"""bench.py's JSON contract: the reference arm on CPU (bounded sample, oracle port on the host cores) and, on a GPU box,
the product arm's line with roofline, cpu_baseline-free short run, e2e and clocks."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "1.0"])
    assert d["impl"] == "reference" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True and d["scaling"] == "strong"
    assert d["metric"] == "AM-chain aggregate input Msamples/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("config5") and d["config"]["channels_total"] == 65536 and d["config"]["block"] == 65536
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_product_arm_line(cuda):
    d = _run(["--steps", "3", "--warmup", "3", "--no-cpu", "--block", "8192", "--no-side", "--e2e-channels", "4096"])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "gpu", "gpu_launches", "clocks", "roofline", "e2e"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "strong"
    assert d["gpu_launches"] == 3 * len(d["gpu"]["kernels"]) and d["gpu"]["channels_per_gpu"] == 65536
    r = d["roofline"]
    assert r["kernel"].split(" ")[0] in d["gpu"]["kernels"]            # named by the library, not guessed
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["clocks"]
    assert c["samples"] >= 1 and c["sm_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}

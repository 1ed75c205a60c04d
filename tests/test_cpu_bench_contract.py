"""The bench line of the reference arm (`bench.py --impl reference`: the unmodified reference module on the host cores)
carries every key the measurement contract names, and the product arm refuses to run without a GPU instead of falling
back.  (The test uses the `tiny` workload; the driver's run uses the default one at full T and resolution.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_schema():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--workload", "tiny"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    from oracle import ref_loader
    assert cb["kind"] == ("reference" if ref_loader.available() else "port")
    assert cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert "scaled" not in cb["sample"]  # no pixel / T extrapolation
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0  # no silent CPU fallback

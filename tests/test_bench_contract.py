"""bench.py's reference arm runs on the CPU (oracle port of the stage): one tiny run checks the JSON contract the driver parses
(keys, units, impl marker, zero copy bytes) without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c5", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    line = r.stdout.strip().splitlines()[-1]
    d = json.loads(line)
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_default_arm_refuses_to_run_without_a_gpu():
    """no CPU fallback: on a box without a CUDA device the product arm must fail loudly, not fall back to the oracle"""
    sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200"))
    import b2enc
    if b2enc.lib().b2_device_count() > 0:
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout) or "GPU" in (r.stderr + r.stdout)

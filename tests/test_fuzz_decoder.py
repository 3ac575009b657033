"""Randomised decoder round trip on the CPU (scripts/cpu_fuzz_entropy.py): oracle encode stage + host slice writers (CAVLC /
CABAC, deblocking, 8x8 transform, partitions) -> libavcodec -> equals the oracle's reconstruction.  A 6,000-case run is recorded
in profiles/r1_final_fuzz.txt; the suite runs a short slice."""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_streams_decode_to_the_oracle_reconstruction():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "cpu_fuzz_entropy.py"), "120", "11"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "120 cases decoder-exact" in r.stdout


def test_host_slice_writers_under_sanitizers():
    """scripts/host_entropy_asan.py: random frames through both slice writers, dense and packed levels, built with
    -fsanitize=address,undefined: identical output, overflow reported, no finding"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "host_entropy_asan.py"), "40", "13"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "no sanitizer finding" in r.stdout


def test_oracle_under_sanitizers():
    """scripts/oracle_asan.sh: the C oracle's encode stage (all tool sets, random sizes and content) built with
    -fsanitize=address,undefined -- no out-of-bounds access, no undefined arithmetic"""
    r = subprocess.run(["bash", os.path.join(ROOT, "scripts", "oracle_asan.sh"), "24", "7"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "runtime error" not in r.stderr and "no sanitizer finding" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]

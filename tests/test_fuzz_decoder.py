"""Randomised decoder round trip on the CPU (scripts/cpu_fuzz_entropy.py): oracle encode stage + host slice writers (CAVLC /
CABAC, deblocking, 8x8 transform, partitions) -> libavcodec -> equals the oracle's reconstruction.  A 6,000-case run is recorded
in profiles/r1_final_fuzz.txt; the suite runs a short slice."""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_streams_decode_to_the_oracle_reconstruction():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "cpu_fuzz_entropy.py"), "120", "11"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "120 cases decoder-exact" in r.stdout

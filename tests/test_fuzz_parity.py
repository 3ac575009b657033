"""Randomised engine-vs-oracle sweep (scripts/gpu_fuzz.py): random picture sizes (down to one macroblock), QPs, search
ranges, tool sets (deblocking, 8x8 transform, partitions 0/1/2, packed levels), slot counts and content types; every field,
level and reconstructed plane must be bit-identical.  The suite runs a short slice; profiles/r1_final_fuzz.txt records a
1,500-case run on the B200."""
import os, subprocess, sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_configurations_bit_exact():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_fuzz.py"), "40", "7"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "40 cases bit-exact" in r.stdout


def test_random_dropin_streams_byte_identical():
    """scripts/gpu_fuzz_dropin.py: the x264-mirror call sequence with random presets, profiles, tools, GOP lengths, GOP slots and
    frame counts (partial batches) produces the oracle encoder's stream byte for byte"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_fuzz_dropin.py"), "40", "9"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "40 drop-in streams byte-identical" in r.stdout


def test_random_conversions_bit_exact():
    """scripts/gpu_fuzz_sws.py: the sws_scale mirror (K0) over all eight source formats with random sizes / strides / padding"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_fuzz_sws.py"), "160", "5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "160 conversions bit-identical" in r.stdout

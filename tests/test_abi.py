"""The C-ABI library loads and exports every function the headers in include/ declare (CPU only, no
compute calls); the host entry points that need no GPU honour the reference's error convention."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for hdr in ("b2enc.h", "b2enc_engine.h", "b2enc_kernels.h", "b2enc_filters.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(b2k?_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_all_declared_symbols_exported(b2):
    lib = b2.lib()
    fns = declared_functions()
    assert len(fns) >= 30
    missing = [f for f in fns if not hasattr(lib, f)]
    assert not missing, missing


def test_param_and_error_conventions(b2):
    L = b2._dropin_lib()
    p = b2.Param()
    assert L.b2_param_default_preset(C.byref(p), b"medium", b"film") == 0          # av_encode.c:102-103 defaults
    assert (p.i_merange, p.b_subpel, p.i_keyint_max) == (16, 1, 32)
    assert p.i_gop_slots == 32                                                       # closed GOPs in flight per GPU (include/b2enc.h)
    assert (p.b_cabac, p.b_deblocking_filter) == (1, 1)                             # x264's medium: CABAC + loop filter
    assert L.b2_param_apply_profile(C.byref(p), b"baseline") == 0 and p.b_cabac == 0  # a profile only removes tools
    assert L.b2_param_default_preset(C.byref(p), b"ultrafast", None) == 0 and (p.b_cabac, p.b_deblocking_filter) == (0, 0)
    assert L.b2_param_default_preset(C.byref(p), b"slow", None) == 0 and p.i_merange == 32
    assert p.b_me_prune == 1                                                         # lossless search pruning: on by default, same bytes
    assert L.b2_param_default_preset(C.byref(p), b"medium", b"zerolatency") == 0 and p.i_gop_slots == 1
    # tunes as in x264: film = deblock -1:-1 (the reference's default, av_encode.c:103), several tunes separated by ','
    assert L.b2_param_default_preset(C.byref(p), b"medium", b"film") == 0
    assert (p.i_deblocking_filter_alphac0, p.i_deblocking_filter_beta) == (-1, -1)
    assert L.b2_param_default_preset(C.byref(p), b"medium", b"film,zerolatency") == 0
    assert (p.i_gop_slots, p.i_deblocking_filter_alphac0) == (1, -1)
    assert L.b2_param_default_preset(C.byref(p), b"medium", b"stillimage") == 0 and p.i_deblocking_filter_beta == -3
    assert L.b2_param_default_preset(C.byref(p), b"medium", b"film,nosuchtune") != 0
    assert L.b2_param_default_preset(C.byref(p), b"warp9", None) != 0              # av_encode.c:384-386
    assert L.b2_param_default_preset(C.byref(p), b"medium", b"nosuchtune") != 0
    assert L.b2_param_apply_profile(C.byref(p), None) == 0                          # av_encode.c:403 with profile NULL
    assert L.b2_param_apply_profile(C.byref(p), b"high") == 0
    assert L.b2_param_apply_profile(C.byref(p), b"ultra") != 0                      # av_encode.c:404-405


def test_no_cpu_fallback(b2):
    """without a GPU the product refuses to run instead of silently computing on the CPU"""
    if b2.lib().b2_device_count() > 0:
        return
    L = b2._dropin_lib()
    p = b2.Param()
    L.b2_param_default_preset(C.byref(p), b"medium", None)
    p.i_width, p.i_height = 64, 64
    assert not L.b2_encoder_open(C.byref(p))                                        # NULL, like av_encode.c:408-411 expects
    assert not L.b2_sws_getContext(64, 64, 0, 64, 64, 0, 1, None, None, None)
    import pytest
    with pytest.raises(RuntimeError):
        b2.require_gpu()

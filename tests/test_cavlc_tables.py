"""Structural checks of the CAVLC tables in video-encoder_b200/host/b2h_cavlc.c (H.264 tables 9-4,
9-5, 9-7..9-10): every VLC table must be a complete prefix-free code and the coded_block_pattern
maps must be permutations.  A typo in a rarely used entry would otherwise only show up as a
corrupt stream on unusual content.  CPU only."""
import ctypes as C
import numpy as np


def _table(oracle, which):
    L = oracle.lib()
    L.b2h_table.restype = C.POINTER(C.c_uint8)
    r, c = C.c_int(), C.c_int()
    p = L.b2h_table(which, C.byref(r), C.byref(c))
    return np.ctypeslib.as_array(p, shape=(r.value, c.value)).copy()


def _check_prefix_code(lens, bits, complete=True):
    codes = [(int(l), int(b)) for l, b in zip(lens, bits) if l > 0]
    strs = [format(b, "0%db" % l) for l, b in codes]
    assert all(len(s) == l for s, (l, _) in zip(strs, codes)), "code value wider than its length"
    assert len(set(strs)) == len(strs), "duplicate code"
    for i, a in enumerate(strs):
        for j, b in enumerate(strs):
            assert i == j or not b.startswith(a), f"{a} is a prefix of {b}"
    kraft = sum(2.0 ** -l for l, _ in codes)
    if complete:
        # complete up to the reserved long all-zero prefixes (start-code emulation guard)
        assert 0 <= 1.0 - kraft <= 2.0 ** -8, f"Kraft sum {kraft}"
    else:
        assert kraft <= 1.0


def test_coeff_token(oracle):
    ln, bt = _table(oracle, 0), _table(oracle, 1)
    for t in range(4):
        valid = np.array([[tc, t1] for tc in range(17) for t1 in range(4)])
        mask = np.array([t1 <= min(tc, 3) for tc, t1 in valid])
        assert ((ln[t] > 0) == mask).all(), "coeff_token table has entries for impossible (TotalCoeff,T1)"
        # tables 0..2 are complete codes; table 3 is a 6-bit FLC with two unused words
        _check_prefix_code(ln[t], bt[t], complete=(t < 3))
    ln, bt = _table(oracle, 2)[0], _table(oracle, 3)[0]
    _check_prefix_code(ln, bt, complete=False)


def test_total_zeros_and_runs(oracle):
    ln, bt = _table(oracle, 4), _table(oracle, 5)
    for tc in range(1, 16):
        n = 16 - tc + 1                                  # total_zeros in 0..16-tc
        assert (ln[tc - 1][:n] > 0).all() and (ln[tc - 1][n:] == 0).all()
        _check_prefix_code(ln[tc - 1][:n], bt[tc - 1][:n])
    ln, bt = _table(oracle, 6), _table(oracle, 7)
    for tc in range(1, 4):
        n = 4 - tc + 1
        _check_prefix_code(ln[tc - 1][:n], bt[tc - 1][:n])
    ln, bt = _table(oracle, 8), _table(oracle, 9)
    for zl in range(1, 8):
        n = zl + 1 if zl < 7 else 15
        assert (ln[zl - 1][:n] > 0).all()
        _check_prefix_code(ln[zl - 1][:n], bt[zl - 1][:n], complete=(zl < 7))


def test_cbp_maps(oracle):
    for which in (10, 11):
        t = _table(oracle, which)[0]
        assert sorted(t.tolist()) == list(range(48))

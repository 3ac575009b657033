"""The host CABAC arithmetic coder (video-encoder_b200/host/b2h_cabac.c: 64-bit low, four bytes leaving at once, carries added
in the buffer, table-driven renormalisation) against a literal transcription of ITU-T H.264 9.3.4.2, Figures 9-7..9-12
(PutBit with firstBitFlag and bitsOutstanding, RenormE one bit at a time) on random and adversarial bin lists: long bypass
runs (carry chains and 0xff runs), skewed contexts (long MPS runs, then an LPS: the largest renormalisation shifts), terminate
bins in the middle.  CPU only; the slice-level syntax is pinned separately by libavcodec's decoder (tests/test_cabac.py)."""
import ctypes as C
import numpy as np
import pytest


class Textbook:
    """9.3.4.2, flow charts transcribed one box per statement"""
    def __init__(self, range_lps, trans_lps, state):
        self.rtab, self.ttab = range_lps, trans_lps
        self.pstate = [int(s) >> 1 for s in state]; self.mps = [int(s) & 1 for s in state]
        self.low, self.rng, self.first, self.outstanding, self.bits = 0, 510, 1, 0, []

    def put_bit(self, b):                       # Figure 9-9
        if self.first: self.first = 0
        else: self.bits.append(b)
        while self.outstanding > 0:
            self.bits.append(1 - b); self.outstanding -= 1

    def renorm(self):                           # Figure 9-8
        while self.rng < 256:
            if self.low < 256: self.put_bit(0)
            elif self.low >= 512: self.low -= 512; self.put_bit(1)
            else: self.low -= 256; self.outstanding += 1
            self.rng <<= 1; self.low <<= 1

    def decision(self, ctx, b):                 # Figure 9-7
        q = (self.rng >> 6) & 3
        rlps = int(self.rtab[self.pstate[ctx]][q])
        self.rng -= rlps
        if b != self.mps[ctx]:
            self.low += self.rng; self.rng = rlps
            if self.pstate[ctx] == 0: self.mps[ctx] = 1 - self.mps[ctx]
            self.pstate[ctx] = int(self.ttab[self.pstate[ctx]])
        else:
            self.pstate[ctx] = min(self.pstate[ctx] + 1, 62)
        self.renorm()

    def bypass(self, b):                        # Figure 9-10
        self.low <<= 1
        if b: self.low += self.rng
        if self.low >= 1024: self.put_bit(1); self.low -= 1024
        elif self.low < 512: self.put_bit(0)
        else: self.low -= 512; self.outstanding += 1

    def terminate(self, b):                     # Figures 9-11 / 9-12
        self.rng -= 2
        if b:
            self.low += self.rng
            self.rng = 2; self.renorm()
            self.put_bit((self.low >> 9) & 1)
            v = ((self.low >> 7) & 3) | 1
            self.bits += [(v >> 1) & 1, v & 1]
        else:
            self.renorm()

    def bytes(self):
        bits = self.bits + [0] * (-len(self.bits) % 8)
        return np.packbits(np.array(bits, dtype=np.uint8)).tobytes()


def _run(oracle, ops, state):
    L = oracle.lib()
    L.b2h_cabac_code_bins.restype = C.c_size_t
    L.b2h_cabac_code_bins.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
    L.b2h_cabac_table.restype = C.POINTER(C.c_uint8)
    rtab = np.ctypeslib.as_array(L.b2h_cabac_table(0), shape=(64, 4)).copy()
    ttab = np.ctypeslib.as_array(L.b2h_cabac_table(1), shape=(64,)).copy()
    ops = np.ascontiguousarray(ops, dtype=np.uint32); state = np.ascontiguousarray(state, dtype=np.uint8)
    out = np.zeros(len(ops) * 2 + 64, dtype=np.uint8)
    n = L.b2h_cabac_code_bins(ops.ctypes.data, len(ops), state.ctypes.data, out.ctypes.data, out.size)
    t = Textbook(rtab, ttab, state)
    for op in ops.tolist():
        ctx, b = op >> 2, op & 1
        if ctx < 1024: t.decision(ctx, b)
        elif ctx == 1024: t.bypass(b)
        else: t.terminate(b)
    return out[:n].tobytes(), t.bytes()


BYPASS, TERM = 1024 << 2, 1025 << 2


def _states(rng):
    return ((rng.integers(0, 63, 1024) << 1) | rng.integers(0, 2, 1024)).astype(np.uint8)


@pytest.mark.parametrize("seed", range(12))
def test_random_bins(oracle, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 6000))
    ctx = rng.integers(0, int(rng.choice([3, 40, 1024])), n)
    p1 = rng.random() if seed % 3 else 0.5
    ops = (ctx << 2) | (rng.random(n) < p1)
    kind = rng.random(n)
    ops = np.where(kind < 0.15, BYPASS | (rng.random(n) < 0.5), ops)
    ops = np.where(kind > 0.98, TERM, ops)                      # end_of_slice_flag = 0 between macroblocks
    ops = np.append(ops, TERM | 1)
    got, want = _run(oracle, ops, _states(rng))
    assert got == want


@pytest.mark.parametrize("bit", [0, 1])
@pytest.mark.parametrize("n", [1, 7, 8, 9, 23, 24, 25, 31, 32, 33, 64, 1000])
def test_bypass_runs_carry_chains(oracle, bit, n):
    # bypass ones keep low close to the top of the interval: 0xff runs; a following LPS carries through all of them
    rng = np.random.default_rng(n)
    ops = [BYPASS | bit] * n + [(5 << 2) | 1, (5 << 2) | 0] * 3 + [BYPASS | (1 - bit)] * (n // 2) + [TERM | 1]
    st = _states(rng); st[5] = (62 << 1) | 0                    # most skewed state: LPS = 1 has the smallest range
    got, want = _run(oracle, np.array(ops), st)
    assert got == want


def test_skewed_contexts_largest_shifts(oracle):
    rng = np.random.default_rng(3)
    ops = []
    for k in range(400):
        c = int(rng.integers(0, 4))
        ops += [(c << 2) | 0] * int(rng.integers(1, 60)) + [(c << 2) | 1] + [BYPASS | 1] * int(rng.integers(0, 12))
        if k % 37 == 0: ops.append(TERM)
    ops.append(TERM | 1)
    st = np.zeros(1024, dtype=np.uint8); st[:4] = [(62 << 1), (40 << 1), (0 << 1), (63 << 1)]      # 63: the terminate-like state
    got, want = _run(oracle, np.array(ops), st)
    assert got == want


def test_every_flush_phase(oracle):
    # the final flush with every number of pending bits (the four-byte output leaves 0..31 bits + the 10-bit window behind)
    rng = np.random.default_rng(9)
    st = _states(rng)
    for n in range(0, 80):
        ops = [BYPASS | int(b) for b in rng.integers(0, 2, n)] + [TERM | 1]
        got, want = _run(oracle, np.array(ops), st)
        assert got == want, n

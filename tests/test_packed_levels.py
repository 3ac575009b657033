"""Packed level stream (engine option pack_levels, kernel K9; layout in include/b2enc_types.h): the host slice writers
must produce the same NAL from the packed stream as from the dense 832 B/MB array, for both entropy coders and with
the 8x8 transform (whose 64-level blocks travel as a unit).  CPU only: the stream is built on the host here; the GPU
test (test_engine_parity.py::test_engine_packed_levels) checks that K9 produces exactly this stream."""
import numpy as np
import pytest
from test_oracle_decode import smooth_seq


@pytest.mark.parametrize("cabac,t8", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_slice_from_packed_levels_equals_dense(oracle, b2, cabac, t8):
    w, h, qp = 208, 160, 24
    frames = smooth_seq(w, h, 4, seed=21, cut=2)
    _, _, infos, coefs = oracle.encode_sequence(frames, w, h, qp=qp, merange=16, gop=32, cabac=cabac, transform8x8=t8, deblock=1)
    ent = oracle.Entropy(w, h, qp, cabac=cabac, transform8x8=t8, deblock=1)
    total_dense = total_packed = 0
    for t, (info, coef) in enumerate(zip(infos, coefs)):
        packed = b2.pack_levels(info, coef)
        assert np.array_equal(b2.unpack_levels(info, packed)["blk"], coef["blk"])          # nothing but zero blocks is dropped
        ft = 0 if t == 0 else 1
        assert ent.slice_packed(ft, t, 0, info, packed) == ent.slice(ft, t, 0, info, coef)
        total_dense += coef.nbytes; total_packed += packed.size
    assert total_packed < total_dense
    if t8:
        assert sum(int(i["transform8x8"].sum()) for i in infos) > 0

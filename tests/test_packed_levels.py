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


def test_packed_decision_records_keep_everything_the_slice_writers_read(oracle):
    """b2_mbinfo_packed_t (24 bytes per macroblock, what crosses PCIe with pack_levels): pack -> unpack keeps every field the
    CABAC / CAVLC writers read, for every tool (8x8 transform, intra 8x8, all partition shapes, I16x16 / I4x4 / inter): the slices
    written from the round-tripped records are byte-identical; `cost` / `i8_modes` / the analysis modes of inter MBs are dropped"""
    import ctypes as C
    from test_oracle_decode import smooth_seq, shear_seq, coarse_seq
    L = oracle.lib()
    w, h = 208, 160
    seqs = [(smooth_seq(w, h, 4, seed=5, cut=2), dict(transform8x8=1, partitions=2)), (shear_seq(w, h, 4, seed=6, amp=3), dict(transform8x8=1, partitions=1)),
            (coarse_seq(w, h, 3, seed=7, scale=10), dict(transform8x8=1)), (smooth_seq(w, h, 3, seed=8), dict()),
            ([oracle.synth_frame(w, h, t) for t in range(3)], dict())]                      # flat regions: I16x16
    seen_parts, seen_types = set(), set()
    for frames, kw in seqs:
        _, _, infos, coefs = oracle.encode_sequence(frames, w, h, qp=30, merange=16, gop=32, deblock=1, cabac=1, **kw)
        for cabac in (0, 1):
            ent = oracle.Entropy(w, h, 30, cabac=cabac, deblock=1, transform8x8=kw.get("transform8x8", 0))
            for t, (info, coef) in enumerate(zip(infos, coefs)):
                info = np.ascontiguousarray(info)
                packed = np.zeros(info.size * 24, np.uint8); back = np.zeros_like(info)
                L.b2h_info_pack(info.ctypes.data_as(C.c_void_p), packed.ctypes.data_as(C.c_void_p), info.size)
                L.b2h_info_unpack(packed.ctypes.data_as(C.c_void_p), back.ctypes.data_as(C.c_void_p), info.size)
                assert ent.slice(0 if t == 0 else 1, t, 0, back, coef) == ent.slice(0 if t == 0 else 1, t, 0, info, coef)
                for f in ("mvx", "mvy", "mb_type", "cbp", "nnz_mask", "part", "transform8x8", "mv8"):
                    assert np.array_equal(back[f], info[f]), f
                intra = info["mb_type"] != 0
                assert np.array_equal(back["i4_mode"][intra], info["i4_mode"][intra]) and np.array_equal(back["chroma_mode"][intra], info["chroma_mode"][intra])
                assert not back["cost"].any() and not back["i8_modes"].any() and not back["i4_mode"][~intra].any()
                seen_parts |= set(info["part"].tolist()); seen_types |= set(info["mb_type"].tolist())
    assert seen_parts == {0, 1, 2, 3} and seen_types == {0, 1, 2, 3}

"""ctypes loader for the CPU oracle (oracle/libb2oracle.so).  TEST INFRASTRUCTURE ONLY:
import from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product (video-encoder_b200/) never imports this module."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PAD = 64
PADC = 32


class Frame(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("w16", C.c_int), ("h16", C.c_int),
                ("mbw", C.c_int), ("mbh", C.c_int), ("pitch", C.c_int), ("pitchc", C.c_int),
                ("buf", C.c_void_p * 3), ("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p)]


class Params(C.Structure):
    _fields_ = [("qp", C.c_int), ("merange", C.c_int), ("subpel", C.c_int), ("intra_in_p", C.c_int)]


MV = np.dtype([("x", "<i2"), ("y", "<i2")])
MBINFO = np.dtype([("mvx", "<i2"), ("mvy", "<i2"), ("mb_type", "u1"), ("i16_mode", "u1"),
                   ("chroma_mode", "u1"), ("cbp", "u1"), ("i4_mode", "u1", (16,)),
                   ("cost", "<u4"), ("nnz_mask", "<u4")])
MBCOEF = np.dtype([("blk", "<i2", (26, 16))])
assert MBINFO.itemsize == 32 and MBCOEF.itemsize == 832


def build(force=False):
    so = os.path.join(_HERE, "libb2oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.b2o_psnr_y.restype = C.c_double
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OFrame:
    """A padded oracle frame (owns its memory)."""

    def __init__(self, w, h):
        self.f = Frame()
        if lib().b2o_frame_alloc(C.byref(self.f), w, h) != 0:
            raise MemoryError
        self.w, self.h = w, h

    def __del__(self):
        try:
            lib().b2o_frame_free(C.byref(self.f))
        except Exception:
            pass

    def load(self, y, u, v):
        y = np.ascontiguousarray(y, np.uint8); u = np.ascontiguousarray(u, np.uint8); v = np.ascontiguousarray(v, np.uint8)
        planes = (C.c_void_p * 3)(y.ctypes.data, u.ctypes.data, v.ctypes.data)
        strides = (C.c_int * 3)(y.shape[1], u.shape[1], v.shape[1])
        lib().b2o_frame_load(C.byref(self.f), planes, strides)
        return self

    def _plane(self, ptr, pitch, w, h):
        n = pitch * (h - 1) + w
        buf = (C.c_uint8 * n).from_address(ptr)
        a = np.frombuffer(buf, np.uint8)
        return np.lib.stride_tricks.as_strided(a, (h, w), (pitch, 1))

    @property
    def y(self):   # coded-size view
        return self._plane(self.f.y, self.f.pitch, self.f.w16, self.f.h16)

    @property
    def u(self):
        return self._plane(self.f.u, self.f.pitchc, self.f.w16 // 2, self.f.h16 // 2)

    @property
    def v(self):
        return self._plane(self.f.v, self.f.pitchc, self.f.w16 // 2, self.f.h16 // 2)


def synth_frame(w, h, t, stream=0):
    y = np.empty((h, w), np.uint8)
    u = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    v = np.empty_like(u)
    lib().b2o_synth_frame(w, h, t, stream, _p(y), _p(u), _p(v))
    return y, u, v


def me_fullpel(cur: OFrame, ref: OFrame, R, pmv=None, lam=0):
    n = cur.f.mbw * cur.f.mbh
    mv = np.zeros(n, MV); cost = np.zeros(n, np.uint32)
    lib().b2o_me_fullpel(C.byref(cur.f), C.byref(ref.f), R, _p(pmv) if pmv is not None else None, lam, _p(mv), _p(cost))
    return mv, cost


def me_fullpel_mb(cur, ref, R, mbx, mby, pmv=(0, 0), lam=0):
    mv = np.zeros(1, MV); cost = np.zeros(1, np.uint32)
    p = np.zeros(1, MV); p["x"] = pmv[0]; p["y"] = pmv[1]
    lib().b2o_me_fullpel_mb.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
    packed = int(p.view(np.uint32)[0])
    lib().b2o_me_fullpel_mb(C.addressof(cur.f), C.addressof(ref.f), R, mbx, mby, packed, lam, _p(mv), _p(cost))
    return (int(mv["x"][0]), int(mv["y"][0])), int(cost[0])

"""ctypes binding of libb2enc.so -- the Python-side harness over the C-ABI in include/*.h.
The library is CUDA-only: there is no CPU fallback, loading or calling fails loudly when the
shared object or a GPU is missing."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MV = np.dtype([("x", "<i2"), ("y", "<i2")])
MBINFO = np.dtype([("mvx", "<i2"), ("mvy", "<i2"), ("mb_type", "u1"), ("i16_mode", "u1"),
                   ("chroma_mode", "u1"), ("cbp", "u1"), ("i4_mode", "u1", (16,)),
                   ("cost", "<u4"), ("nnz_mask", "<u4")])
MBCOEF = np.dtype([("blk", "<i2", (26, 16))])


def so_path():
    return os.path.join(_HERE, "libb2enc.so")


def build(force=False):
    if force or not os.path.exists(so_path()):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so_path()


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(so_path()):
            raise RuntimeError("libb2enc.so is not built (run `make -C video-encoder_b200`); "
                               "there is no CPU fallback for the encode stage")
        _LIB = C.CDLL(so_path())
        _LIB.b2_bench_vabsdiff4_peak.restype = C.c_double
        _LIB.b2_bench_vabsdiff4_peak.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    return _LIB


def require_gpu():
    if lib().b2_device_count() <= 0:
        raise RuntimeError("b2enc: no CUDA device visible; the encode stage has no CPU fallback")


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def me_fullpel(cur_y, ref_y, merange, pmv=None, lam=0, iters=0):
    """cur_y/ref_y: uint8 [n,h,w] (or [h,w]).  Returns (mv[n,mbs], cost[n,mbs], kernel_ms|None)."""
    require_gpu()
    cur_y = np.ascontiguousarray(cur_y, np.uint8); ref_y = np.ascontiguousarray(ref_y, np.uint8)
    if cur_y.ndim == 2:
        cur_y = cur_y[None]; ref_y = ref_y[None]
    n, h, w = cur_y.shape
    nmb = (w // 16) * (h // 16)
    mv = np.zeros((n, nmb), MV); cost = np.zeros((n, nmb), np.uint32)
    if pmv is not None:
        pmv = np.ascontiguousarray(pmv, MV).reshape(n, nmb)
    ms = C.c_float(0)
    rc = lib().b2k_me_fullpel(_p(cur_y), _p(ref_y), w, h, n, merange, _p(pmv), lam, _p(mv), _p(cost),
                              iters, C.byref(ms) if iters > 0 else None)
    if rc != 0:
        raise RuntimeError("b2k_me_fullpel failed (%d)" % rc)
    return mv, cost, (ms.value if iters > 0 else None)


def vabsdiff4_peak(device=0, outer=256, reps=5):
    require_gpu()
    ms = C.c_double(0)
    rate = lib().b2_bench_vabsdiff4_peak(device, outer, reps, C.byref(ms))
    if rate <= 0:
        raise RuntimeError("vabsdiff4 microbenchmark failed")
    return rate, ms.value

"""Generates tests/golden/sws_golden.npz: input pictures and the I420 output of the LIVE libswscale
(9.1.100, bundled with the OpenCV wheel) called exactly like the reference does at
av_encode.c:427-430 / :545-547: sws_getContext(W,H,srcfmt, W,H,YUV420P, SWS_FAST_BILINEAR,0,0,0)
then one whole-frame sws_scale.  Run in the build container (needs cv2); the .npz is committed."""
import ctypes as C
import glob
import os
import numpy as np
import cv2  # noqa: F401

d = os.path.join(os.path.dirname(cv2.__file__), "..", "opencv_python_headless.libs")
avutil = C.CDLL(glob.glob(os.path.join(d, "libavutil-*.so*"))[0])
sws = C.CDLL(glob.glob(os.path.join(d, "libswscale-*.so*"))[0])
sws.sws_getContext.restype = C.c_void_p
sws.sws_getContext.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
sws.sws_scale.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
avutil.av_get_pix_fmt.restype = C.c_int
avutil.av_get_pix_fmt.argtypes = [C.c_char_p]
SWS_FAST_BILINEAR = 1


def run_sws(fmt_name, w, h, planes, strides):
    src_fmt = avutil.av_get_pix_fmt(fmt_name.encode()); dst_fmt = avutil.av_get_pix_fmt(b"yuv420p")
    ctx = sws.sws_getContext(w, h, src_fmt, w, h, dst_fmt, SWS_FAST_BILINEAR, None, None, None)
    assert ctx
    cw, ch = (w + 1) // 2, (h + 1) // 2
    dy = np.zeros((h, w + 16), np.uint8); du = np.zeros((ch, cw + 16), np.uint8); dv = np.zeros((ch, cw + 16), np.uint8)
    sp = (C.c_void_p * 4)(*[p.ctypes.data if p is not None else None for p in planes + [None] * (4 - len(planes))])
    ss = (C.c_int * 4)(*(strides + [0] * (4 - len(strides))))
    dp = (C.c_void_p * 4)(dy.ctypes.data, du.ctypes.data, dv.ctypes.data, None)
    ds = (C.c_int * 4)(dy.shape[1], du.shape[1], dv.shape[1], 0)
    r = sws.sws_scale(ctx, sp, ss, 0, h, dp, ds)
    assert r == h, r
    sws.sws_freeContext(C.c_void_p(ctx))
    return dy[:, :w].copy(), du[:, :cw].copy(), dv[:, :cw].copy()


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    cases = [(64, 48), (158, 82), (61, 35), (34, 18)]
    for (w, h) in cases:
        cw, ch = (w + 1) // 2, (h + 1) // 2
        # yuv420p with padded strides
        y = rng.integers(0, 256, (h, w + 7), dtype=np.uint8); u = rng.integers(0, 256, (ch, cw + 5), dtype=np.uint8)
        v = rng.integers(0, 256, (ch, cw + 3), dtype=np.uint8)
        o = run_sws("yuv420p", w, h, [y, u, v], [y.shape[1], u.shape[1], v.shape[1]])
        out[f"yuv420p_{w}x{h}_in0"] = y[:, :w]; out[f"yuv420p_{w}x{h}_in1"] = u[:, :cw]; out[f"yuv420p_{w}x{h}_in2"] = v[:, :cw]
        for i in range(3): out[f"yuv420p_{w}x{h}_out{i}"] = o[i]
        # nv12
        y = rng.integers(0, 256, (h, w), dtype=np.uint8); uv = rng.integers(0, 256, (ch, 2 * cw), dtype=np.uint8)
        o = run_sws("nv12", w, h, [y, uv], [w, 2 * cw])
        out[f"nv12_{w}x{h}_in0"] = y; out[f"nv12_{w}x{h}_in1"] = uv
        for i in range(3): out[f"nv12_{w}x{h}_out{i}"] = o[i]
        # packed 4:2:2 (even widths only)
        if w % 2 == 0:
            for fmt in ("yuyv422", "uyvy422"):
                p = rng.integers(0, 256, (h, 2 * w), dtype=np.uint8)
                o = run_sws(fmt, w, h, [p], [2 * w])
                out[f"{fmt}_{w}x{h}_in0"] = p
                for i in range(3): out[f"{fmt}_{w}x{h}_out{i}"] = o[i]
            # bgr24: dedicated converter in libswscale (even widths), bit-exact closed form in the oracle
            p = rng.integers(0, 256, (h, 3 * w), dtype=np.uint8)
            o = run_sws("bgr24", w, h, [p], [3 * w])
            out[f"bgr24_{w}x{h}_in0"] = p
            for i in range(3): out[f"bgr24_{w}x{h}_out{i}"] = o[i]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sws_golden.npz"), **out)
    print("wrote", len(out), "arrays")
    tolerance_set()


def smooth(rng, h, w, scale=6):
    """band-limited content: the tolerance-pinned formats run through libswscale's drifting fast-bilinear pass"""
    b = rng.integers(0, 256, (h // scale + 3, w // scale + 3)).astype(np.float32)
    B = cv2.resize(b, (b.shape[1] * scale, b.shape[0] * scale), interpolation=cv2.INTER_CUBIC)
    return np.clip(B[:h, :w], 0, 255).astype(np.uint8)


def tolerance_set():
    """tests/golden/sws_tolerance.npz: rgb24 / yuv422p / yuv411p on small smooth pictures, where
    the position drift of libswscale's SWS_FAST_BILINEAR pass stays below rounding (see oracle/b2o_convert.c)"""
    rng = np.random.default_rng(20261019)
    out = {}
    for (w, h) in [(64, 48), (70, 38), (34, 18), (48, 36), (61, 20), (32, 18), (96, 20)]:
        cw, cw4 = (w + 1) // 2, (w + 3) // 4
        y = smooth(rng, h, w)
        for fmt, scw in (("yuv422p", cw), ("yuv411p", cw4)):
            if fmt == "yuv411p" and w % 32:       # chroma doubling: only the x86 path (output chroma width % 16 == 0) is centre-aligned;
                continue                          # the C path maps (srcW-2)/(dstW-2), an endpoint-aligned stretch that is not reproduced
            u = smooth(rng, h, scw, 4); v = smooth(rng, h, scw, 4)
            o = run_sws(fmt, w, h, [y, u, v], [w, scw, scw])
            out[f"{fmt}_{w}x{h}_in0"] = y; out[f"{fmt}_{w}x{h}_in1"] = u; out[f"{fmt}_{w}x{h}_in2"] = v
            for i in range(3): out[f"{fmt}_{w}x{h}_out{i}"] = o[i]
        if w % 2:
            continue
        rgb = np.ascontiguousarray(np.stack([smooth(rng, h, w), smooth(rng, h, w), smooth(rng, h, w)], axis=2).reshape(h, 3 * w))
        o = run_sws("rgb24", w, h, [rgb], [3 * w])
        out[f"rgb24_{w}x{h}_in0"] = rgb
        for i in range(3): out[f"rgb24_{w}x{h}_out{i}"] = o[i]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sws_tolerance.npz"), **out)
    print("wrote", len(out), "tolerance arrays")


if __name__ == "__main__":
    main()

"""Randomised sweep of the sws_scale mirror (K0 through b2_sws_getContext / b2_sws_scale, host buffers) against the C oracle
conversion (which the committed libswscale golden vectors pin): all eight source formats, random sizes incl. odd ones where
the format allows, random source strides and destination padding, random content.   usage: gpu_fuzz_sws.py [cases] [seed]"""
import sys, os, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("video-encoder_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, d))
import numpy as np
import b2enc, b2oracle

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
FMTS = ["yuv420p", "nv12", "yuyv422", "uyvy422", "bgr24", "rgb24", "yuv422p", "yuv411p"]


def size_ok(fmt, w, h):
    if fmt in ("yuyv422", "uyvy422", "rgb24"): return w % 2 == 0 and h % 2 == 0
    if fmt == "bgr24": return w % 2 == 0
    if fmt == "yuv422p": return h % 2 == 0
    if fmt == "yuv411p": return w % 4 == 0 and h % 2 == 0
    return True


def planes(fmt, w, h, pad):
    cw, ch = (w + 1) // 2, (h + 1) // 2
    shapes = {"yuv420p": [(h, w), (ch, cw), (ch, cw)], "nv12": [(h, w), (ch, 2 * cw)], "yuyv422": [(h, 2 * w)], "uyvy422": [(h, 2 * w)],
              "bgr24": [(h, 3 * w)], "rgb24": [(h, 3 * w)], "yuv422p": [(h, w), (h, cw), (h, cw)], "yuv411p": [(h, w), (h, (w + 3) // 4), (h, (w + 3) // 4)]}[fmt]
    mode = rng.integers(0, 3)
    out = []
    for (r, c) in shapes:
        if mode == 0: a = rng.integers(0, 256, (r, c + pad), dtype=np.uint8)
        elif mode == 1: a = rng.choice(np.array([0, 1, 127, 128, 254, 255], np.uint8), (r, c + pad))
        else: a = (np.add.outer(np.arange(r) * 3, np.arange(c + pad) * 5) & 255).astype(np.uint8)
        out.append(a)
    return out


t0 = time.time()
for case in range(ncases):
    fmt = FMTS[case % len(FMTS)]
    while True:
        w = int(rng.integers(16, 700)); h = int(rng.integers(16, 500))
        if size_ok(fmt, w, h): break
    ins = planes(fmt, w, h, int(rng.integers(0, 40)))
    desc = f"case {case}: {fmt} {w}x{h} strides {[p.shape[1] for p in ins]}"
    try:
        y, u, v = b2enc.sws_convert(fmt, w, h, ins, dst_pad=int(rng.integers(0, 17)))
        oy, ou, ov = b2oracle.convert_to_i420(fmt, w, h, ins)
        assert np.array_equal(y, oy) and np.array_equal(u, ou) and np.array_equal(v, ov)
    except Exception:
        print("MISMATCH/ERROR", desc, flush=True); traceback.print_exc(); sys.exit(1)
    print("ok", desc, flush=True)
print(f"{ncases} conversions bit-identical to the oracle in {time.time() - t0:.0f} s")

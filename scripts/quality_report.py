"""Config C1 of BASELINE.json (640x480 @30, 300 synthetic frames, fixed QP): the CPU comparator (C oracle encoder; the
reference's libx264 path is not buildable here) against the GPU drop-in encoder -- bitstream size, Y-PSNR of the decoded
stream, decodability -- for every tool set of rows N1-N3 (CAVLC / CABAC / 8x8 transform / partitions), on the BASELINE
content (moving gradient + noise) and on sheared motion.  Writes gpurun_out/quality_c1.json (copied to profiles/)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, b2enc, b2oracle as o
from test_oracle_decode import shear_seq

W, H, QP, GOP, R = 640, 480, 26, 32, 16
CONTENT = {"synthetic pan (BASELINE C1)": (300, lambda n: [o.synth_frame(W, H, t) for t in range(n)]),
           "sheared motion": (20, lambda n: shear_seq(W, H, n, seed=5, stripe=72, band=56, amp=3))}
TOOLS = [("baseline: CAVLC, 4x4, 16x16 inter", dict(profile="baseline"), dict(cabac=0)),
         ("main: CABAC", dict(), dict(cabac=1)),
         ("high: CABAC + adaptive 8x8 / intra 8x8", dict(profile="high", b_transform_8x8=1), dict(cabac=1, transform8x8=1)),
         ("high + partitions (local)", dict(profile="high", b_transform_8x8=1, b_partitions=1), dict(cabac=1, transform8x8=1, partitions=1)),
         ("high + partitions (own full-pel search)", dict(profile="high", b_transform_8x8=1, b_partitions=2), dict(cabac=1, transform8x8=1, partitions=2))]


def psnr_of(bs, frames):
    dec = o.decode_yuv(o.split_access_units(bs))
    assert len(dec) == len(frames)
    sse = 0.0
    for (dy, _, _), (y, _, _) in zip(dec, frames):
        d = dy.astype(np.int32) - y.astype(np.int32)
        sse += float((d * d).sum())
    return 10 * np.log10(255.0 ** 2 * W * H * len(frames) / sse)


res = {"config": "640x480@30, QP %d, GOP %d, merange %d, deblocking on" % (QP, GOP, R),
       "comparator": "C oracle encoder (libx264 absent: parity unpinned against x264 itself)", "runs": []}
for cname, (n, gen) in CONTENT.items():
    frames = gen(n)
    for tname, dkw, okw in TOOLS:
        t0 = time.time()
        ref_bs, _, _, _ = o.encode_sequence(frames, W, H, qp=QP, merange=R, gop=GOP, fps=(30, 1), deblock=1, deblock_offsets=(-1, -1), **okw)
        t_cpu = time.time() - t0
        enc = b2enc.DropInEncoder(W, H, preset="medium", tune="film", quality=QP, fps=(30, 1), annexb=1, i_keyint_max=GOP, i_gop_slots=8, **dkw)
        t0 = time.time()
        out = []
        for t, fr in enumerate(frames):
            size, nals, pts, dts, key = enc.encode(fr, t)
            if size > 0: out.append(b"".join(d for _, d in nals))
        while enc.delayed() > 0:
            size, nals, pts, dts, key = enc.encode(None, 0)
            out.append(b"".join(d for _, d in nals))
        t_gpu = time.time() - t0
        enc.close()
        gpu_bs = b"".join(out)
        run = {"content": cname, "frames": n, "tools": tname, "gpu_bytes": len(gpu_bs), "cpu_bytes": len(ref_bs),
               "kbps": round(len(gpu_bs) * 8 * 30 / n / 1e3, 1), "gpu_psnr_y": round(psnr_of(gpu_bs, frames), 3),
               "identical_bitstream": gpu_bs == ref_bs, "cpu_oracle_seconds_1thread": round(t_cpu, 2),
               "gpu_dropin_seconds_incl_host_entropy": round(t_gpu, 3), "gpu_dropin_fps": round(n / t_gpu, 1)}
        res["runs"].append(run)
        print(json.dumps(run), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "quality_c1.json"), "w"), indent=1)

"""Config C1 of BASELINE.json (640x480 @30, 300 synthetic frames, fixed QP): the CPU comparator (C oracle encoder; the
reference's libx264 path is not buildable here) against the GPU drop-in encoder -- bitstream size, Y-PSNR of the decoded
stream, decodability.  Writes gpurun_out/quality_c1.json (copied to profiles/ by hand)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, b2enc, b2oracle as o

W, H, N, QP, GOP, R = 640, 480, 300, 26, 32, 16
frames = [o.synth_frame(W, H, t) for t in range(N)]

t0 = time.time()
ref_bs, ref_recons, _, _ = o.encode_sequence(frames, W, H, qp=QP, merange=R, gop=GOP, fps=(30, 1), deblock=1)
t_cpu = time.time() - t0

enc = b2enc.DropInEncoder(W, H, preset="medium", tune="film", quality=QP, fps=(30, 1), annexb=1, i_keyint_max=GOP, i_gop_slots=8)
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


def psnr_of(bs):
    dec = o.decode_yuv(o.split_access_units(bs))
    assert len(dec) == N
    sse = 0.0
    for (dy, _, _), (y, _, _) in zip(dec, frames):
        d = dy.astype(np.int32) - y.astype(np.int32)
        sse += float((d * d).sum())
    return 10 * np.log10(255.0 ** 2 * W * H * N / sse)

res = {"config": "C1 640x480@30, 300 synthetic frames, QP %d, GOP %d, merange %d" % (QP, GOP, R),
       "comparator": "C oracle encoder (libx264 absent: parity unpinned against x264 itself)",
       "cpu_bytes": len(ref_bs), "gpu_bytes": len(gpu_bs), "bitrate_ratio": len(gpu_bs) / len(ref_bs),
       "cpu_kbps": len(ref_bs) * 8 * 30 / N / 1e3, "cpu_psnr_y": psnr_of(ref_bs), "gpu_psnr_y": psnr_of(gpu_bs),
       "identical_bitstream": gpu_bs == ref_bs, "cpu_seconds_1thread": t_cpu, "gpu_dropin_seconds_incl_host_entropy": t_gpu,
       "gpu_dropin_fps": N / t_gpu}
res["psnr_delta_db"] = res["gpu_psnr_y"] - res["cpu_psnr_y"]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "quality_c1.json"), "w"), indent=1)
print(json.dumps(res))

"""Randomised sweep of the x264-mirror drop-in (b2_encoder_encode, include/b2enc.h) on the GPU: random sizes, QPs, presets,
profiles / tools, GOP lengths, GOP-slot counts and frame counts (partial last batch, partial last GOP); the concatenated NAL
payloads must be byte-identical to the oracle encoder's stream and decode (libavcodec) to the oracle's reconstruction.
usage: gpu_fuzz_dropin.py [cases] [seed]"""
import sys, os, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("video-encoder_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, d))
import numpy as np
import b2enc, b2oracle
from test_dropin import drive, to_annexb
from test_oracle_decode import smooth_seq, coarse_seq, shear_seq

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time()
for case in range(ncases):
    w = int(rng.integers(8, 60)) * 2; h = int(rng.integers(8, 48)) * 2
    qp = int(rng.integers(12, 50)); gop = int(rng.integers(1, 7)); slots = int(rng.integers(1, 5)); n = int(rng.integers(1, 20))
    preset = str(rng.choice(["medium", "slow", "ultrafast", "veryfast"])); profile = rng.choice([None, "baseline", "main", "high"])
    t8 = int(rng.random() < 0.5); parts = int(rng.choice([0, 1, 2])); deblock = int(rng.random() < 0.7); annexb = int(rng.random() < 0.5)
    if preset == "ultrafast": parts = 0                      # no sub-pel stage to refine partitions in
    kind = str(rng.choice(["smooth", "coarse", "shear"])); sd = int(rng.integers(0, 1 << 30))
    fr = smooth_seq(w, h, n, seed=sd, cut=(int(rng.integers(1, n)) if n > 2 and rng.random() < 0.4 else None)) if kind == "smooth" else \
        coarse_seq(w, h, n, seed=sd, scale=int(rng.integers(4, 20))) if kind == "coarse" else shear_seq(w, h, n, seed=sd, amp=int(rng.integers(1, 3)))
    desc = f"case {case}: {w}x{h} qp {qp} preset {preset} profile {profile} t8 {t8} parts {parts} deblock {deblock} annexb {annexb} gop {gop} slots {slots} frames {n} {kind} seed {sd}"
    try:
        out = drive(b2enc, fr, w, h, preset=preset, tune="film", quality=qp, profile=profile, annexb=annexb, i_keyint_max=gop,
                    i_gop_slots=slots, b_transform_8x8=t8, b_partitions=parts, b_deblocking_filter=deblock)
        assert len(out) == n, "frame count"
        bs = to_annexb(out, length_prefixed=not annexb)
        # what the parameter set resolves to (b2h_encoder.c: presets[], b2_param_apply_profile)
        merange = 32 if preset == "slow" else 16
        subpel = 0 if preset == "ultrafast" else 1; intra = 0 if preset == "ultrafast" else 1
        cabac = 0 if (preset == "ultrafast" or profile == "baseline") else 1
        eff_t8 = 0 if profile in ("baseline", "main") else t8
        eff_deblock = deblock
        prm_parts = parts
        ref_bs, recons, _, _ = b2oracle.encode_sequence(fr, w, h, qp=qp, merange=merange, subpel=subpel, intra_in_p=intra, gop=gop, fps=(30, 1),
                                                         deblock=eff_deblock, cabac=cabac, transform8x8=eff_t8, partitions=prm_parts,
                                                         deblock_offsets=(-1, -1))
        assert bs == ref_bs, "bitstream differs from the oracle encoder's"
        dec = b2oracle.decode_yuv(b2oracle.split_access_units(bs))
        assert len(dec) == n and all(np.array_equal(d[0], r.y[:h, :w]) for d, r in zip(dec, recons)), "decoder drift"
    except Exception:
        print("MISMATCH/ERROR", desc, flush=True); traceback.print_exc(); sys.exit(1)
    print("ok", desc, flush=True)
print(f"{ncases} drop-in streams byte-identical to the oracle encoder's in {time.time() - t0:.0f} s")

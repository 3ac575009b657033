"""Test helper: the smallest ISO-BMFF (MP4) file that carries one AVC video track -- stands in for the reference's
libmp4v2 muxer (enc_mp4_write_video_sample, av_encode.c:683-744) so that the drop-in encoder's b_annexb = 0 output
(4-byte length-prefixed NALs, one sample per picture) and b2_avcc_write's record can be validated by a real demuxer
(libavformat through cv2.VideoCapture).  Not part of the product."""
import struct


def box(kind, payload):
    return struct.pack(">I4s", 8 + len(payload), kind) + payload


def full(kind, version, flags, payload):
    return box(kind, struct.pack(">I", (version << 24) | flags) + payload)


def write_mp4(path, avcc, samples, sync, w, h, timescale=30, delta=1):
    """samples: list of byte strings (length-prefixed NALs of one picture); sync: list of bools (IDR)."""
    n = len(samples)
    dur = n * delta
    mat = struct.pack(">9I", 0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000)
    ftyp = box(b"ftyp", b"isom" + struct.pack(">I", 512) + b"isomiso2avc1mp41")
    mdat_payload = b"".join(samples)

    def moov(chunk_off):
        mvhd = full(b"mvhd", 0, 0, struct.pack(">IIII", 0, 0, timescale, dur) + struct.pack(">IH", 0x10000, 0x100) + b"\0" * 10 +
                    mat + b"\0" * 24 + struct.pack(">I", 2))
        tkhd = full(b"tkhd", 0, 3, struct.pack(">IIIII", 0, 0, 1, 0, dur) + b"\0" * 8 + struct.pack(">HHHH", 0, 0, 0, 0) + mat +
                    struct.pack(">II", w << 16, h << 16))
        mdhd = full(b"mdhd", 0, 0, struct.pack(">IIIIHH", 0, 0, timescale, dur, 0x55c4, 0))
        hdlr = full(b"hdlr", 0, 0, struct.pack(">I4s", 0, b"vide") + b"\0" * 12 + b"VideoHandler\0")
        vmhd = full(b"vmhd", 0, 1, b"\0" * 8)
        dinf = box(b"dinf", full(b"dref", 0, 0, struct.pack(">I", 1) + full(b"url ", 0, 1, b"")))
        avc1 = box(b"avc1", b"\0" * 6 + struct.pack(">H", 1) + b"\0" * 16 + struct.pack(">HH", w, h) +
                   struct.pack(">II", 0x480000, 0x480000) + struct.pack(">I", 0) + struct.pack(">H", 1) + b"\0" * 32 +
                   struct.pack(">Hh", 24, -1) + box(b"avcC", avcc))
        stsd = full(b"stsd", 0, 0, struct.pack(">I", 1) + avc1)
        stts = full(b"stts", 0, 0, struct.pack(">III", 1, n, delta))
        stss = full(b"stss", 0, 0, struct.pack(">I", sum(sync)) + b"".join(struct.pack(">I", i + 1) for i, s in enumerate(sync) if s))
        stsc = full(b"stsc", 0, 0, struct.pack(">IIII", 1, 1, n, 1))
        stsz = full(b"stsz", 0, 0, struct.pack(">II", 0, n) + b"".join(struct.pack(">I", len(s)) for s in samples))
        stco = full(b"stco", 0, 0, struct.pack(">II", 1, chunk_off))
        stbl = box(b"stbl", stsd + stts + stss + stsc + stsz + stco)
        minf = box(b"minf", vmhd + dinf + stbl)
        mdia = box(b"mdia", mdhd + hdlr + minf)
        return box(b"moov", mvhd + box(b"trak", tkhd + mdia))

    off = len(ftyp) + len(moov(0)) + 8
    with open(path, "wb") as f:
        f.write(ftyp + moov(off) + box(b"mdat", mdat_payload))

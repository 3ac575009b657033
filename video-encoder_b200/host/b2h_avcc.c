/*
 * b2h_avcc.c -- AVCDecoderConfigurationRecord (the payload of the MP4 `avcC` box, ISO/IEC 14496-15 5.2.4.1) from the
 * encoder's SPS and PPS NAL units.  SURVEY.md 8f row N3: in the reference this record is assembled by libmp4v2 from
 * MP4AddH264VideoTrack(..., profile, compat, level, sampleLenFieldSizeMinusOne = 3) (av_encode.c:638-640, patched with
 * the SPS bytes at :703-716) plus MP4AddH264SequenceParameterSet / MP4AddH264PictureParameterSet (:722, :727).
 * b2_avcc_write builds the same record directly, so that a muxer without libmp4v2 can consume the drop-in
 * encoder's b_annexb = 0 output (4-byte big-endian NAL lengths, lengthSizeMinusOne = 3).
 */
#include <string.h>
#include "b2enc.h"

int b2_avcc_write(const uint8_t *sps, int sps_size, const uint8_t *pps, int pps_size, uint8_t *out, int cap)
{
    if (!sps || !pps || !out || sps_size < 4 || pps_size < 1 || sps_size > 0xffff || pps_size > 0xffff) return -1;
    if ((sps[0] & 31) != B2_NAL_SPS || (pps[0] & 31) != B2_NAL_PPS) return -1;     /* NAL units without length prefix / start code */
    const int high = sps[1] == 100 || sps[1] == 110 || sps[1] == 122 || sps[1] == 144;
    const int need = 6 + 2 + sps_size + 1 + 2 + pps_size + (high ? 4 : 0);
    if (cap < need) return -1;
    uint8_t *p = out;
    *p++ = 1;                                  /* configurationVersion                              */
    *p++ = sps[1];                             /* AVCProfileIndication   (av_encode.c:703)          */
    *p++ = sps[2];                             /* profile_compatibility  (av_encode.c:704)          */
    *p++ = sps[3];                             /* AVCLevelIndication     (av_encode.c:705)          */
    *p++ = 0xfc | 3;                           /* lengthSizeMinusOne = 3 (av_encode.c:640)          */
    *p++ = 0xe0 | 1;                           /* numOfSequenceParameterSets                        */
    *p++ = (uint8_t)(sps_size >> 8); *p++ = (uint8_t)sps_size;
    memcpy(p, sps, (size_t)sps_size); p += sps_size;
    *p++ = 1;                                  /* numOfPictureParameterSets                         */
    *p++ = (uint8_t)(pps_size >> 8); *p++ = (uint8_t)pps_size;
    memcpy(p, pps, (size_t)pps_size); p += pps_size;
    if (high) {                                /* High profiles: chroma_format 4:2:0, 8-bit, no SPS extensions */
        *p++ = 0xfc | 1; *p++ = 0xf8 | 0; *p++ = 0xf8 | 0; *p++ = 0;
    }
    return (int)(p - out);
}

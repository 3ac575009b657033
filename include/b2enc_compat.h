/*
 * b2enc_compat.h -- the reference-side binding: include this in av_encode.c INSTEAD of <x264.h> (av_encode.c:27) and
 * <libswscale/swscale.h> (:28), after the libav* headers (it maps libavutil's PIX_FMT_* values).  Every libx264 /
 * libswscale name the reference's hot path uses is redirected to its libb2enc.so mirror (b2enc.h); no
 * other source line changes.  tests/test_reference_text.py compiles the reference's own text of enc_x264_open (:378-438),
 * enc_x264_close (:440-444), enc_avfilter_pull_to_x264_context (:525-560) and of the two loops (:968-975, :1076-1083)
 * through this header, links it and runs it.
 */
#ifndef B2ENC_COMPAT_H
#define B2ENC_COMPAT_H
#include "b2enc.h"

/* the reference passes the decoder's enum PixelFormat (av_encode.c:428) and PIX_FMT_YUV420P (:429); five formats are
 * bit-exact against libswscale, three more (DV / archive sources) are the drift-free closed form
 * pinned with a tolerance.  Anything else -> -1 -> b2_sws_getContext returns NULL -> the existing error path :432-435 */
static inline int b2_fmt_from_av(int f)
{
    switch (f) {
    case PIX_FMT_YUV420P: return B2_FMT_YUV420P;
    case PIX_FMT_NV12:    return B2_FMT_NV12;
    case PIX_FMT_YUYV422: return B2_FMT_YUYV422;
    case PIX_FMT_UYVY422: return B2_FMT_UYVY422;
    case PIX_FMT_BGR24:   return B2_FMT_BGR24;
    case PIX_FMT_RGB24:   return B2_FMT_RGB24;
    case PIX_FMT_YUV422P: return B2_FMT_YUV422P;
    case PIX_FMT_YUV411P: return B2_FMT_YUV411P;
    default:              return -1;
    }
}

/* libswscale subset (av_encode.c:427-430, :441, :545-547) */
#define SwsContext                 b2_sws_context
#define sws_getContext(sw, sh, sfmt, dw, dh, dfmt, flags, sf, df, par) \
    b2_sws_getContext((sw), (sh), b2_fmt_from_av(sfmt), (dw), (dh), b2_fmt_from_av(dfmt), (flags), (sf), (df), (par))
#define sws_scale                  b2_sws_scale
#define sws_freeContext            b2_sws_freeContext
#define SWS_FAST_BILINEAR          B2_SWS_FAST_BILINEAR

/* libx264 subset (av_encode.c:369-444, :543-544, :970, :1076-1078) */
#define x264_t                      b2_t
#define x264_param_t                b2_param_t
#define x264_picture_t              b2_picture_t
#define x264_nal_t                  b2_nal_t
#define x264_param_default_preset   b2_param_default_preset
#define x264_param_apply_profile    b2_param_apply_profile
#define x264_encoder_open           b2_encoder_open
#define x264_picture_alloc          b2_picture_alloc      /* page-locked host memory */
#define x264_picture_clean          b2_picture_clean
#define x264_encoder_encode         b2_encoder_encode
#define x264_encoder_delayed_frames b2_encoder_delayed_frames
#define x264_encoder_close          b2_encoder_close
#define X264_CSP_I420               B2_CSP_I420
#define X264_TYPE_AUTO              B2_TYPE_AUTO
#define X264_RC_CRF                 B2_RC_CRF
#define NAL_SPS                     B2_NAL_SPS
#define NAL_PPS                     B2_NAL_PPS
#define NAL_FILLER                  B2_NAL_FILLER

#endif

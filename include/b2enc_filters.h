/*
 * b2enc_filters.h -- C-ABI of the GPU pre-filter stage (libb2enc.so), SURVEY.md 8f row N4: the step in front of the hot
 * path.  The reference pipes every decoded frame through a user-supplied libavfilter graph (av_encode.c:69-71 option
 * "--filters", built at :451-517, fed at :962, drained at :525-560) and its author recommends "hqdn3d,yadif" for DV
 * material (av_encode.c:35).  This header mirrors that push / poll / pull interface for exactly those two filters,
 * executed by CUDA kernels (csrc/k10_filters.cu):
 *
 *   reference call (av_encode.c)                                    replacement
 *   -------------------------------------------------------------  ---------------------------
 *   avfilter_graph_alloc + create_filter + graph_parse + config     b2_filter_graph_create   (:451-517)
 *   av_vsrc_buffer_add_frame(src, frame, OVERWRITE)                  b2_filter_add_frame      (:962)
 *   avfilter_poll_frame(sink->inputs[0])                             b2_filter_poll_frame     (:529)
 *   av_vsink_buffer_get_video_buffer_ref + fill_frame + unref        b2_filter_get_frame      (:532-550)
 *   (end of input: libavfilter flushes on EOF)                       b2_filter_flush
 *   avfilter_graph_free                                              b2_filter_graph_free
 *
 * Parity: libavfilter is not available in this image, so the two filters are pinned only against the C restatement in
 * oracle/b2o_filters.c (bit-exact), which follows the published algorithms of vf_hqdn3d.c / vf_yadif.c.
 * Error convention as everywhere: NULL / negative + message on stderr; no CPU fallback.
 */
#ifndef B2ENC_FILTERS_H
#define B2ENC_FILTERS_H
#include <stdint.h>
#include "b2enc_types.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2_filter_graph b2_filter_graph_t;

/* `filters`: libavfilter chain syntax restricted to "hqdn3d[=luma_spatial[:chroma_spatial[:luma_tmp[:chroma_tmp]]]]" and
 * "yadif[=mode[:parity]]" (mode 0 only; parity -1 auto from the frame's top_field_first, 0 tff, 1 bff), joined by ','.
 * NULL or "" is a pass-through graph (av_encode.c:482-509).  fmt: B2_FMT_YUV420P, B2_FMT_YUV422P or B2_FMT_YUV411P. */
b2_filter_graph_t *b2_filter_graph_create(int width, int height, int fmt, const char *filters, int device);
/* push one frame; the source has been read when the call returns.  Returns 0, < 0 on error. */
int b2_filter_add_frame(b2_filter_graph_t *g, const uint8_t *const src[3], const int stride[3], int64_t pts, int top_field_first);
/* number of filtered frames waiting to be pulled */
int b2_filter_poll_frame(b2_filter_graph_t *g);
/* pull the oldest filtered frame into the caller's planes; returns 1, 0 when none is ready, < 0 on error */
int b2_filter_get_frame(b2_filter_graph_t *g, uint8_t *const dst[3], const int stride[3], int64_t *pts);
/* end of input: frames a filter still holds (yadif keeps one) become available */
int b2_filter_flush(b2_filter_graph_t *g);
void b2_filter_graph_free(b2_filter_graph_t *g);

#ifdef __cplusplus
}
#endif
#endif

// b2_internal.h -- internal launcher prototypes (device pointers + stream); not part of the C-ABI.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b2enc_types.h"

int b2_make_plane_tmap(CUtensorMap *tm, const void *base, int pitch, int rows, int nplanes, int bw, int bh);
int b2_launch_extend_border(uint8_t *d_planes, int pitch, int rows, int nplanes, int pad, int iw, int ih,
                            cudaStream_t st);
extern "C" int b2_k1_window_box(int R, int *bw, int *bh);
int b2_launch_me_fullpel(int R, const CUtensorMap *tm_cur, const CUtensorMap *tm_ref, int mbw, int mbh,
                         int nframes, const b2_mv_t *d_pmv, int lambda, b2_mv_t *d_mv, uint32_t *d_cost,
                         cudaStream_t st);

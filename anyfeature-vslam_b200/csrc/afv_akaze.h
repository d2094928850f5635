// afv_akaze.h -- internal interface of the akaze61 extractor (afv_akaze.cu), called from the C ABI (afv_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/afv.h"

struct AfvAkaze;
int  afv_akaze_create(AfvAkaze** out, int nfeatures, int nlevels, float scale_factor, float detect_th, int max_batch, int max_w, int max_h);
void afv_akaze_destroy(AfvAkaze* s);
uint8_t* afv_akaze_stage(AfvAkaze* s);
int  afv_akaze_run(AfvAkaze* s, const uint8_t* d_gray, int B, int w, int h, int stride, long frame_stride, afv_keypoint* d_kps,
                   uint8_t* d_desc, float* d_kpsize, int cap, int* d_n_out, cudaStream_t st);
int  afv_akaze_status(AfvAkaze* s, int B, cudaStream_t st);
int  afv_akaze_debug_read(AfvAkaze* s, int what, int frame, int level, void* out, long cap_bytes, long* n_bytes);

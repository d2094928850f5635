// afv_brisk.h -- internal interface of the brisk48 extractor (afv_brisk.cu), called from the C ABI (afv_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/afv.h"

struct AfvBrisk;
int  afv_brisk_create(AfvBrisk** out, int nfeatures, int nlevels, float scale_factor, float detect_th, int max_batch, int max_w, int max_h);
void afv_brisk_destroy(AfvBrisk* s);
uint8_t* afv_brisk_stage(AfvBrisk* s);
int  afv_brisk_run(AfvBrisk* s, const uint8_t* d_gray, int B, int w, int h, int stride, long frame_stride, afv_keypoint* d_kps,
                   uint8_t* d_desc, float* d_kpsize, int cap, int* d_n_out, cudaStream_t st);
int  afv_brisk_status(AfvBrisk* s, int B, cudaStream_t st);
int  afv_brisk_debug_read(AfvBrisk* s, int what, int frame, int level, void* out, long cap_bytes, long* n_bytes);

// afv_sift.h -- internal interface of the sift128 extractor (afv_sift.cu), called from the C ABI (afv_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/afv.h"

struct AfvSift;
int  afv_sift_create(AfvSift** out, int nfeatures, int nlevels, float scale_factor, int max_batch, int max_w, int max_h);
void afv_sift_destroy(AfvSift* s);
uint8_t* afv_sift_stage(AfvSift* s);     // device staging buffer for the host-buffer API (max_batch * max_w * max_h bytes)
int  afv_sift_run(AfvSift* s, const uint8_t* d_gray, int B, int w, int h, int stride, long frame_stride, afv_keypoint* d_kps,
                  float* d_desc, float* d_kpsize, int cap, int* d_n_out, cudaStream_t st);
int  afv_sift_status(AfvSift* s, int B, cudaStream_t st);
int  afv_sift_debug_read(AfvSift* s, int what, int frame, int level, void* out, long cap_bytes, long* n_bytes);

// afv_orbslam2.h -- internal interface of the vanilla ORB-SLAM2 extractor (afv_orbslam2.cu), called from the C ABI (afv_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/afv.h"

struct AfvOs2;
int  afv_os2_create(AfvOs2** out, int nfeatures, int nlevels, float scale_factor, float detect_th, int max_batch, int max_w, int max_h);
void afv_os2_destroy(AfvOs2* s);
uint8_t* afv_os2_stage(AfvOs2* s);
int  afv_os2_run(AfvOs2* s, const uint8_t* d_gray, int B, int w, int h, int stride, long frame_stride, afv_keypoint* d_kps,
                 uint8_t* d_desc, float* d_kpsize, int cap, int* d_n_out, cudaStream_t st);
int  afv_os2_status(AfvOs2* s, int B, cudaStream_t st);
int  afv_os2_debug_read(AfvOs2* s, int what, int frame, int level, void* out, long cap_bytes, long* n_bytes);

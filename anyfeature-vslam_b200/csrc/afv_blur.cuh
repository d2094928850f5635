// afv_blur.cuh -- separable Gaussian step on float images, shared by the sift128 and akaze61 scale spaces.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct AfvBlurTaps { float t[16]; };     // centre outward, t[0..R]

// MODE 0: float source; 1: u8 source, value / 255.0f (sift128); 2: u8 source, value * (1.0f / 255.0f) (akaze61)
// ------------------------------------------------------------------------------------------------------
// Gaussian step.  CTA = 128 x 32 outputs, 256 threads.  The (128+2R) x (32+2R) source footprint is staged with clamped
// coordinates; row pass: one thread = 4 consecutive outputs from a (4+2R)-float register window (float4 LDS); column
// pass: one thread = 8 consecutive rows of one column from an (8+2R) register window.  acc = t0*c; acc += tj*(l + r).
// ------------------------------------------------------------------------------------------------------
#define SB_W 128
#define SB_H 32
template <int R, int MODE>
__global__ void __launch_bounds__(256) k_afv_blur(const void* __restrict__ src_, int sstride, long long sfstride,
                                                  float* __restrict__ dst, float* __restrict__ dog, int w, int h,
                                                  int stride, long long istride, const __grid_constant__ AfvBlurTaps taps) {
    constexpr int PW = (SB_W + 2 * R + 3) & ~3;          // staged row pitch (floats)
    constexpr int PH = SB_H + 2 * R;
    extern __shared__ __align__(16) float sm[];
    float* in = sm;                                      // [PH][PW]
    float* mid = sm + PH * PW;                           // [PH][SB_W]
    const int tid = threadIdx.x, f = blockIdx.z;
    const int tx0 = blockIdx.x * SB_W, ty0 = blockIdx.y * SB_H;
    float tp[R + 1];
#pragma unroll
    for (int j = 0; j <= R; ++j) tp[j] = taps.t[j];
    if (MODE != 0) {
        const uint8_t* s = reinterpret_cast<const uint8_t*>(src_) + (long long)f * sfstride;
        for (int i = tid; i < PH * PW; i += 256) {
            const int ry = i / PW, rx = i - ry * PW;
            const int y = min(max(ty0 - R + ry, 0), h - 1), x = min(max(tx0 - R + rx, 0), w - 1);
            const float pv = (float)s[(long long)y * sstride + x];
            in[i] = MODE == 1 ? pv / 255.0f : pv * (1.0f / 255.0f);
        }
    } else {
        const float* s = reinterpret_cast<const float*>(src_) + (long long)f * sfstride;
        for (int i = tid; i < PH * PW; i += 256) {
            const int ry = i / PW, rx = i - ry * PW;
            const int y = min(max(ty0 - R + ry, 0), h - 1), x = min(max(tx0 - R + rx, 0), w - 1);
            in[i] = s[(long long)y * sstride + x];
        }
    }
    __syncthreads();
    // row pass
    for (int u = tid; u < PH * (SB_W / 4); u += 256) {
        const int ry = u >> 5, xg = (u & 31) * 4;
        constexpr int NW = (4 + 2 * R + 3) / 4;
        float win[NW * 4];
        const float4* p = reinterpret_cast<const float4*>(in + ry * PW + xg);
#pragma unroll
        for (int k = 0; k < NW; ++k) { const float4 v = p[k]; win[4 * k] = v.x; win[4 * k + 1] = v.y; win[4 * k + 2] = v.z; win[4 * k + 3] = v.w; }
        float4 o;
        float* op = reinterpret_cast<float*>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float acc = tp[0] * win[q + R];
#pragma unroll
            for (int j = 1; j <= R; ++j) acc = acc + tp[j] * (win[q + R - j] + win[q + R + j]);
            op[q] = acc;
        }
        *reinterpret_cast<float4*>(mid + ry * SB_W + xg) = o;
    }
    __syncthreads();
    // column pass
    for (int u = tid; u < SB_W * (SB_H / 8); u += 256) {
        const int x = u & 127, yg = (u >> 7) * 8;
        const int gx = tx0 + x;
        float win[8 + 2 * R];
#pragma unroll
        for (int k = 0; k < 8 + 2 * R; ++k) win[k] = mid[(yg + k) * SB_W + x];
        if (gx < w) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int gy = ty0 + yg + q;
                float acc = tp[0] * win[q + R];
#pragma unroll
                for (int j = 1; j <= R; ++j) acc = acc + tp[j] * (win[q + R - j] + win[q + R + j]);
                if (gy < h) {
                    const long long o = (long long)f * istride + (long long)gy * stride + gx;
                    dst[o] = acc;
                    if (dog) dog[o] = acc - in[(yg + q + R) * PW + x + R];
                }
            }
        }
    }
}


template <int R, int MODE> static size_t afv_blur_smem() {
    return sizeof(float) * (size_t)(SB_H + 2 * R) * (((SB_W + 2 * R + 3) & ~3) + SB_W);
}
template <int R, int MODE> static cudaError_t afv_blur_cfg() {
    return cudaFuncSetAttribute(k_afv_blur<R, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)afv_blur_smem<R, MODE>());
}
template <int R, int MODE>
static void afv_blur_launch(const void* src, int sstride, long long sfstride, float* dst, float* dog, int w, int h, int stride,
                            long long istride, const AfvBlurTaps& taps, int B, cudaStream_t st) {
    dim3 g((w + SB_W - 1) / SB_W, (h + SB_H - 1) / SB_H, B);
    k_afv_blur<R, MODE><<<g, 256, afv_blur_smem<R, MODE>(), st>>>(src, sstride, sfstride, dst, dog, w, h, stride, istride, taps);
}

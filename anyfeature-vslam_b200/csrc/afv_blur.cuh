// afv_blur.cuh -- separable Gaussian step on float images, shared by the sift128 and akaze61 scale spaces.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

struct AfvBlurTaps { float t[16]; };     // centre outward, t[0..R]

// MODE 0: float source; 1: u8 source, value / 255.0f (sift128); 2: u8 source, value * (1.0f / 255.0f) (akaze61)
// ------------------------------------------------------------------------------------------------------
// Gaussian step.  CTA = 128 x 32 outputs, 256 threads.  The (128+2R) x (32+2R) source footprint is staged with clamped
// coordinates; row pass: one thread = 4 consecutive outputs from a (4+2R)-float register window (float4 LDS); column
// pass: one thread = 8 consecutive rows of one column from an (8+2R) register window.  acc = t0*c; acc = fma(tj, l + r, acc).
// ------------------------------------------------------------------------------------------------------
#define SB_W 128
#define SB_H 32

// ---- TMA (cp.async.bulk.tensor) + mbarrier primitives, inline PTX (sm_90+/sm_100a) ---------------------------------
__device__ __forceinline__ uint32_t blur_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void blur_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(blur_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void blur_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(blur_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void blur_tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(blur_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(blur_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void blur_mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok = 0;
    for (int spin = 0; !ok; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(blur_smem_u32(bar)), "r"(phase) : "memory");
        if (spin > (1 << 24)) __trap();                  // a lost TMA completion must not hang the box
    }
}

// CTA = 128 x 32 outputs, 256 threads.  Staged footprint: rows ty0-R .. ty0+32+R-1, columns tx0-R4 .. tx0+128+R4-1 with
// R4 = R rounded up to 4 (16-byte aligned rows for float4 LDS and for the TMA box).  Interior tiles of float sources are
// staged by ONE TMA box load (cp.async.bulk.tensor.3d + mbarrier); border tiles and u8 sources use clamped loads.
template <int R, int MODE>
__global__ void __launch_bounds__(256) k_afv_blur(const void* __restrict__ src_, int sstride, long long sfstride,
                                                  float* __restrict__ dst, float* __restrict__ dog, int w, int h,
                                                  int stride, long long istride, const __grid_constant__ AfvBlurTaps taps,
                                                  const __grid_constant__ CUtensorMap tm, int tma_z0) {
    constexpr int R4 = (R + 3) & ~3;
    constexpr int PW = SB_W + 2 * R4;                    // staged row pitch (floats)
    constexpr int PH = SB_H + 2 * R;
    extern __shared__ __align__(128) float sm[];
    float* in = sm;                                      // [PH][PW]
    float* mid = sm + PH * PW;                           // [PH][SB_W]
    __shared__ __align__(8) uint64_t tma_bar;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, f = blockIdx.z;
    const int tx0 = blockIdx.x * SB_W, ty0 = blockIdx.y * SB_H;
    float tp[R + 1];
#pragma unroll
    for (int j = 0; j <= R; ++j) tp[j] = taps.t[j];
    const bool interior = MODE == 0 && tma_z0 >= 0 && tx0 - R >= 0 && tx0 + SB_W + R <= w && ty0 - R >= 0 && ty0 + SB_H + R <= h;
    if (interior) {
        if (tid == 0) blur_mbar_init(&tma_bar, 1);
        __syncthreads();
        if (tid == 0) {
            blur_mbar_expect_tx(&tma_bar, PH * PW * 4);
            blur_tma_load_3d(in, &tm, &tma_bar, tx0 - R4, ty0 - R, tma_z0 + f);
        }
        blur_mbar_wait(&tma_bar, 0);
    } else if (MODE != 0) {
        const uint8_t* s = reinterpret_cast<const uint8_t*>(src_) + (long long)f * sfstride;
        for (int ry = wid; ry < PH; ry += 8) {
            const uint8_t* row = s + (long long)min(max(ty0 - R + ry, 0), h - 1) * sstride;
            for (int c = lane; c < PW; c += 32) {
                const float pv = (float)row[min(max(tx0 - R4 + c, 0), w - 1)];
                in[ry * PW + c] = MODE == 1 ? pv / 255.0f : pv * (1.0f / 255.0f);
            }
        }
    } else {
        const float* s = reinterpret_cast<const float*>(src_) + (long long)f * sfstride;
        for (int ry = wid; ry < PH; ry += 8) {
            const float* row = s + (long long)min(max(ty0 - R + ry, 0), h - 1) * sstride;
            for (int c = lane; c < PW; c += 32) in[ry * PW + c] = row[min(max(tx0 - R4 + c, 0), w - 1)];
        }
    }
    __syncthreads();
    // row pass: outputs xg..xg+3 of staged row ry; output q is centred on staged column xg + q + R4
    for (int u = tid; u < PH * (SB_W / 4); u += 256) {
        const int ry = u >> 5, xg = (u & 31) * 4;
        constexpr int NW = 1 + R4 / 2;
        float win[NW * 4];
        const float4* p = reinterpret_cast<const float4*>(in + ry * PW + xg);
#pragma unroll
        for (int k = 0; k < NW; ++k) { const float4 v = p[k]; win[4 * k] = v.x; win[4 * k + 1] = v.y; win[4 * k + 2] = v.z; win[4 * k + 3] = v.w; }
        float4 o;
        float* op = reinterpret_cast<float*>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float acc = tp[0] * win[q + R4];
#pragma unroll
            for (int j = 1; j <= R; ++j) acc = __fmaf_rn(tp[j], win[q + R4 - j] + win[q + R4 + j], acc);
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
                for (int j = 1; j <= R; ++j) acc = __fmaf_rn(tp[j], win[q + R - j] + win[q + R + j], acc);
                if (gy < h) {
                    const long long o = (long long)f * istride + (long long)gy * stride + gx;
                    dst[o] = acc;
                    if (dog) dog[o] = acc - in[(yg + q + R) * PW + x + R4];
                }
            }
        }
    }
}

template <int R, int MODE> static size_t afv_blur_smem() {
    return sizeof(float) * (size_t)(SB_H + 2 * R) * ((SB_W + 2 * ((R + 3) & ~3)) + SB_W);
}
template <int R, int MODE> static cudaError_t afv_blur_cfg() {
    return cudaFuncSetAttribute(k_afv_blur<R, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)afv_blur_smem<R, MODE>());
}

// 3-D float tensor map (x, y, image) over `nimg` images of w x h floats, row pitch `stride` floats, image pitch `istride`
// floats, with the staging box of k_afv_blur<R>.  Returns false when the driver entry point is missing (callers then
// launch with tm_z0 = -1: every tile takes the clamped-load path).
typedef CUresult (*afv_blur_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool afv_blur_tmap(CUtensorMap* m, const float* base, int w, int h, int stride, long long istride, long long nimg, int R) {
    static afv_blur_encode_fn enc = nullptr;
    if (!enc) {
        void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) return false;
        enc = (afv_blur_encode_fn)fn;
    }
    const int R4 = (R + 3) & ~3;
    const cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)nimg};
    const cuuint64_t gstr[2] = {(cuuint64_t)stride * 4, (cuuint64_t)istride * 4};
    const cuuint32_t box[3] = {(cuuint32_t)(SB_W + 2 * R4), (cuuint32_t)(SB_H + 2 * R), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// tm == nullptr (or tm_z0 < 0): no TMA.  With a map, image f of the launch is image tm_z0 + f of the map.
template <int R, int MODE>
static void afv_blur_launch(const void* src, int sstride, long long sfstride, float* dst, float* dog, int w, int h, int stride,
                            long long istride, const AfvBlurTaps& taps, int B, cudaStream_t st, const CUtensorMap* tm = nullptr, int tm_z0 = -1) {
    dim3 g((w + SB_W - 1) / SB_W, (h + SB_H - 1) / SB_H, B);
    CUtensorMap dummy;
    memset(&dummy, 0, sizeof(dummy));
    k_afv_blur<R, MODE><<<g, 256, afv_blur_smem<R, MODE>(), st>>>(src, sstride, sfstride, dst, dog, w, h, stride, istride, taps,
                                                                   tm ? *tm : dummy, tm ? tm_z0 : -1);
}

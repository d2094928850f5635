// afv_brisk.cu -- hand-written sm_100a kernels of the brisk48 extraction path.
//
// Replaces, behind the C ABI of include/afv.h, what the reference's FeatureExtractor_brisk48 runs on the CPU per frame
// (reference src/Feature_brisk48.cpp:11-60: brisk::BriskFeatureDetector(int(detectTh), nOctaves/2, true).detect ->
// keypoints_level[octave] -> DistributeOctTree per level -> all levels merged -> brisk::BriskDescriptorExtractor(true, true,
// briskV2).compute -> mergeKeypointLevels; computeSize with powf(scaleFactor0, octave)).  ETH brisk v2 is not vendored by the
// reference: the arithmetic follows oracle/afv_oracle_brisk.c (the published BRISK in the BRISK authors' own structure, whose
// detector / orientation / 512-bit descriptor core is pinned to cv2 4.13.0's cv::BRISK; score contract ORC_BRISK_DENSE; the
// 48-byte rows use the documented 384-pair stand-in table).  CUDA == oracle bit for bit (tests/test_brisk_gpu.py).
//
//   k_brk_resize    layer pyramid: cv::resize INTER_AREA (2/3 sample of layer 0, half samples of layer i-2; exact 2x2 mean or
//                   the general float-weight path, <= 4 x 4 taps in OpenCV's accumulation order)
//   k_brk_score     dense AGAST/FAST 9-16 corner score of every layer pixel (smem tile, packed (d,-d) min/max chain) + unordered
//                   list of the detections with score >= threshold
//   k_brk_ismax     thread per detection: BriskScaleSpace::isMax2D on the thresholded score image (3x3 >=, equal neighbours by
//                   their 3x3 binomial sums) -> candidate list
//   k_brk_refine    thread per candidate: refine3D (maxima search in the layers above / below on bilinear score samples, 2-D
//                   quadratic sub-pixel fits, 1-D parabola over scale) -> keypoint (x, y, size, response) + raster key
//   k_brk_octree    CTA per (layer, frame): bitonic sort by raster key (detection order), DistributeOctTree (afv_octree.cuh)
//   k_brk_merge     CTA per frame: layers ascending, pattern scale index, border filter of the descriptor extractor, compaction
//   k_brk_integral* cv::integral (CV_32S) of the frame: row scan + column scan
//   k_brk_describe  warp per kept keypoint: 60 box-smoothed samples (integral image), 870 long pairs -> orientation (integer sums,
//                   double-precision polynomial atan2), rotated samples, 384 short-pair bits; merged cv::KeyPoint rows
#include "afv_common.cuh"
#include "afv_octree.cuh"
#include "afv_brisk.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include <vector>

#define BRK_MAX_LAYERS 8
#define BRK_POINTS 60
#define BRK_SCALES 64
#define BRK_NROT 1024
#define BRK_NLONG 870
#define BRK_NSHORT48 384
#define BRK_KP_CAP 4096            // keypoints per (frame, layer): sorted in shared memory
#define BRK_BASIC_SIZE 12.0f

#define BRK_ST_LIST_OVERFLOW 1
#define BRK_ST_OUT_OVERFLOW 4
#define BRK_ST_OCTREE_OVERFLOW 8

struct BrkLayerG {
    int w, h, stride; long long fstride;                 // score arena geometry (bytes); img uses img_stride / img_fstride
    const uint8_t* img; int img_stride; long long img_fstride;
    uint8_t* score;                                      // [B][h][stride] true 9-16 score (0 below 1)
    float scale, offset;
    int agast_cap, cand_cap, kp_cap;
    uint32_t* agast; uint32_t* cand;                     // [B][cap] y << 16 | x
    float4* kp; uint32_t* kkey;                          // [B][kp_cap] {x, y, size, response}, raster key of the integer maximum
    int src, exact_half;                                 // source layer of the resize
    const int* xsi; const float* xal; const int* ysi; const float* yal;     // [w][4] / [h][4] INTER_AREA taps (alpha 0 = unused)
};

struct BrkParams {
    int B, layers, W, H, nlevels, out_cap, threshold;
    int n_ini; float hX;
    int q_ext[AFV_MAX_LEVELS]; float size_norm[AFV_MAX_LEVELS];
    BrkLayerG lv[BRK_MAX_LAYERS];
    int* cnt;                  // [B][32]: 0..7 #agast, 8..15 #cand, 16..23 #kp
    int* status;               // [B]
    float* okx; float* oky; uint32_t* oresp; float4* okp; unsigned short* knode; unsigned char* kquad;   // [B][layers][BRK_KP_CAP] sorted keys
    int* keep; int keep_cap; int* keepcnt; int oct_ncap;                                                  // [B][layers][keep_cap], [B][16]
    int* integral;             // [B][(H+1)][(W+1)]
    int* kscale;               // [B][out_cap] pattern scale index of the output keypoints
    // pattern tables (device, shared by all extractors of the process)
    const float2* pat;         // [scale][rot][point] sample offsets
    const int4* ptab;          // [scale][point] {sigma bits, scaling, scaling2, 0}
    const int2* longp;         // [870] {i | j << 8, wdx & 0xffff | wdy << 16}
    const uint16_t* shortp;    // [384] i | j << 8
    const float* size_thr;     // [64] smallest keypoint size with scale index >= k
    const int* size_list;      // [64] border per scale index
};

__device__ __forceinline__ uint32_t brk_f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint32_t brk_sat_u8(float v) {
    uint32_t r;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// ------------------------------------------------------------------------------------------------------
// INTER_AREA layer from its source layer
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_brk_resize(const __grid_constant__ BrkParams P, int l) {
    const BrkLayerG& D = P.lv[l];
    const BrkLayerG& S = P.lv[D.src];
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6), f = blockIdx.z;
    if (x >= D.w || y >= D.h) return;
    const uint8_t* src = S.img + (long long)f * S.img_fstride;
    uint8_t* dst = const_cast<uint8_t*>(D.img) + (long long)f * D.img_fstride;
    if (D.exact_half) {
        const uint8_t* s = src + (long long)(2 * y) * S.img_stride + 2 * x;
        dst[(long long)y * D.img_stride + x] = (uint8_t)(((int)s[0] + (int)s[1] + (int)s[S.img_stride] + (int)s[S.img_stride + 1] + 2) >> 2);
        return;
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float beta = D.yal[4 * y + j];
        if (beta == 0.f) break;
        const uint8_t* row = src + (long long)D.ysi[4 * y + j] * S.img_stride;
        float buf = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a = D.xal[4 * x + k];
            if (a == 0.f) break;
            buf = __fadd_rn(buf, __fmul_rn((float)row[D.xsi[4 * x + k]], a));
        }
        sum = __fadd_rn(sum, __fmul_rn(beta, buf));
    }
    dst[(long long)y * D.img_stride + x] = (uint8_t)brk_sat_u8(sum);
}

// ------------------------------------------------------------------------------------------------------
// dense 9-16 corner score.  Tile 64 x 16, 256 threads, 4 pixels per thread.
// ------------------------------------------------------------------------------------------------------
#define BS_W 64
#define BS_H 16
#define BS_PW (BS_W + 8)
#define BRK_CIRC16(F) F(0, 0, 3) F(1, 1, 3) F(2, 2, 2) F(3, 3, 1) F(4, 3, 0) F(5, 3, -1) F(6, 2, -2) F(7, 1, -3) \
                      F(8, 0, -3) F(9, -1, -3) F(10, -2, -2) F(11, -3, -1) F(12, -3, 0) F(13, -3, 1) F(14, -2, 2) F(15, -1, 3)

__global__ void __launch_bounds__(256) k_brk_score(const __grid_constant__ BrkParams P, int l) {
    __shared__ uint8_t pix[BS_H + 6][BS_PW];
    const BrkLayerG& L = P.lv[l];
    const int f = blockIdx.z, tid = threadIdx.x;
    const int x0 = blockIdx.x * BS_W, y0 = blockIdx.y * BS_H;
    const uint8_t* img = L.img + (long long)f * L.img_fstride;
    for (int i = tid; i < (BS_H + 6) * (BS_W + 6); i += 256) {
        const int r = i / (BS_W + 6), c = i - r * (BS_W + 6);
        const int gx = x0 - 3 + c, gy = y0 - 3 + r;
        pix[r][c] = (gx >= 0 && gx < L.w && gy >= 0 && gy < L.h) ? img[(long long)gy * L.img_stride + gx] : 0;
    }
    __syncthreads();
    uint8_t* sc = L.score + (long long)f * L.fstride;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = q * 256 + tid;
        const int r = i / BS_W, c = i % BS_W;
        const int gx = x0 + c, gy = y0 + r;
        if (gx >= L.w || gy >= L.h) continue;
        int s = 0;
        if (gx >= 3 && gy >= 3 && gx < L.w - 3 && gy < L.h - 3) {
            const uint8_t* p = &pix[r + 3][c + 3];
            const int v = p[0];
            uint32_t e[16];
#define BDIFF(k, dx, dy) { const int dv = v - (int)p[(dy) * BS_PW + (dx)]; e[k] = ((uint32_t)dv & 0xffffu) | ((uint32_t)(-dv) << 16); }
            BRK_CIRC16(BDIFF)
#undef BDIFF
            uint32_t m2[16], m4[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) m2[k] = __vmins2(e[k], e[(k + 1) & 15]);
#pragma unroll
            for (int k = 0; k < 16; ++k) m4[k] = __vmins2(m2[k], m2[(k + 2) & 15]);
            uint32_t acc = 0x80008000u;
#pragma unroll
            for (int k = 0; k < 16; ++k) acc = __vmaxs2(acc, __vmins2(__vmins2(m4[k], m4[(k + 4) & 15]), e[(k + 8) & 15]));
            const int bd = (int)(short)(acc & 0xffffu), bb = (int)(short)(acc >> 16);
            s = (bd > bb ? bd : bb) - 1;
            if (s < 1) s = 0;
        }
        sc[(long long)gy * L.stride + gx] = (uint8_t)s;
        if (s >= P.threshold) {
            const int slot = atomicAdd(&P.cnt[f * 32 + l], 1);
            if (slot < L.agast_cap) (L.agast + (long long)f * L.agast_cap)[slot] = ((uint32_t)gy << 16) | (uint32_t)gx;
            else atomicOr(&P.status[f], BRK_ST_LIST_OVERFLOW);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// score access (BriskLayer::getAgastScore): true score inside the 3-pixel border, 0 outside / below the threshold asked for
// ------------------------------------------------------------------------------------------------------
struct BrkView { const uint8_t* sc; int w, h, stride; };
__device__ __forceinline__ BrkView brk_view(const BrkLayerG& L, int f) {
    BrkView v; v.sc = L.score + (long long)f * L.fstride; v.w = L.w; v.h = L.h; v.stride = L.stride; return v;
}
__device__ __forceinline__ int brk_sc(const BrkView& V, int x, int y, int threshold) {
    if (x < 3 || y < 3 || x >= V.w - 3 || y >= V.h - 3) return 0;
    const int s = V.sc[(long long)y * V.stride + x];
    return s >= threshold ? s : 0;
}
__device__ __forceinline__ int brk_sc_f(const BrkView& V, float xf, float yf, int threshold) {
    const int x = (int)xf; const float rx1 = __fsub_rn(xf, (float)x), rx = __fsub_rn(1.0f, rx1);
    const int y = (int)yf; const float ry1 = __fsub_rn(yf, (float)y), ry = __fsub_rn(1.0f, ry1);
    float v = __fmul_rn(__fmul_rn(rx, ry), (float)brk_sc(V, x, y, threshold));
    v = __fadd_rn(v, __fmul_rn(__fmul_rn(rx1, ry), (float)brk_sc(V, x + 1, y, threshold)));
    v = __fadd_rn(v, __fmul_rn(__fmul_rn(rx, ry1), (float)brk_sc(V, x, y + 1, threshold)));
    v = __fadd_rn(v, __fmul_rn(__fmul_rn(rx1, ry1), (float)brk_sc(V, x + 1, y + 1, threshold)));
    return (int)(uint8_t)(int)v;
}
// AGAST 5-8 score on the 8-neighbour ring (BriskLayer::getAgastScore_5_8), computed from the layer image
__device__ int brk_sc58(const uint8_t* img, int w, int h, int stride, int x, int y) {
    if (x < 2 || y < 2 || x >= w - 2 || y >= h - 2) return 0;
    const uint8_t* p = img + (long long)y * stride + x;
    const int v = p[0];
    // (d, -d) packed in the two signed 16-bit halves: one __vmins2 chain gives min(d) and min(-d) of an arc.  The scalar form
    // max(min5(d), -max5(d)) is the pattern nvcc 12.9 / ptxas miscompiles for sm_100a (tools/dbg/minmax_dbg.cu).
    uint32_t e[8];
    const int off[8] = {-1, stride - 1, stride, stride + 1, 1, -stride + 1, -stride, -stride - 1};
#pragma unroll
    for (int k = 0; k < 8; ++k) { const int dv = (int)p[off[k]] - v; e[k] = ((uint32_t)dv & 0xffffu) | ((uint32_t)(-dv) << 16); }
    uint32_t acc = 0x80008000u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t m = __vmins2(__vmins2(e[k], e[(k + 1) & 7]), __vmins2(e[(k + 2) & 7], e[(k + 3) & 7]));
        m = __vmins2(m, e[(k + 4) & 7]);
        acc = __vmaxs2(acc, m);
    }
    const int b0 = (int)(short)(acc & 0xffffu), b1 = (int)(short)(acc >> 16);
    const int best = b0 > b1 ? b0 : b1;
    const int s = best - 1;
    return s >= 1 ? s : 0;
}

// ------------------------------------------------------------------------------------------------------
// isMax2D on the thresholded score image
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_brk_ismax(const __grid_constant__ BrkParams P, int l) {
    const BrkLayerG& L = P.lv[l];
    const int f = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
    const int n = min(P.cnt[f * 32 + l], L.agast_cap);
    if (k >= n) return;
    const uint32_t pk = (L.agast + (long long)f * L.agast_cap)[k];
    const int x = pk & 0xffff, y = pk >> 16;
    const uint8_t* sc = L.score + (long long)f * L.fstride;
    const int T = P.threshold, W = L.stride;
    int v[5][5];                                      // thresholded 5x5 neighbourhood (detections are >= 3 px inside the layer)
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
            const int xx = x + dx, yy = y + dy;
            int s = 0;
            if (xx >= 0 && yy >= 0 && xx < L.w && yy < L.h) { s = sc[(long long)yy * W + xx]; if (s < T) s = 0; }
            v[dy + 2][dx + 2] = s;
        }
    const int c = v[2][2];
    bool ok = true;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) if (c < v[2 + dy][2 + dx]) ok = false;
    if (ok) {
        const int smoothed = 4 * c + 2 * (v[2][1] + v[2][3] + v[1][2] + v[3][2]) + v[1][1] + v[1][3] + v[3][1] + v[3][3];
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if ((dx || dy) && c == v[2 + dy][2 + dx]) {
                    const int cy = 2 + dy, cx = 2 + dx;
                    const int o = v[cy - 1][cx - 1] + 2 * v[cy - 1][cx] + v[cy - 1][cx + 1] + 2 * v[cy][cx - 1] + 4 * v[cy][cx] + 2 * v[cy][cx + 1] +
                                  v[cy + 1][cx - 1] + 2 * v[cy + 1][cx] + v[cy + 1][cx + 1];
                    if (o > smoothed) ok = false;
                }
            }
    }
    if (!ok) return;
    const int slot = atomicAdd(&P.cnt[f * 32 + 8 + l], 1);
    if (slot < L.cand_cap) (L.cand + (long long)f * L.cand_cap)[slot] = pk;
    else atomicOr(&P.status[f], BRK_ST_LIST_OVERFLOW);
}

// ------------------------------------------------------------------------------------------------------
// refine3D and helpers (device restatement of oracle/afv_oracle_brisk.c; float ops explicit, TU compiled --fmad=false)
// ------------------------------------------------------------------------------------------------------
__device__ float brk_subpixel2d(int s_0_0, int s_0_1, int s_0_2, int s_1_0, int s_1_1, int s_1_2, int s_2_0, int s_2_1, int s_2_2,
                                float& delta_x, float& delta_y) {
    const int tmp1 = s_0_0 + s_0_2 - 2 * s_1_1 + s_2_0 + s_2_2;
    const int coeff1 = 3 * (tmp1 + s_0_1 - ((s_1_0 + s_1_2) << 1) + s_2_1);
    const int coeff2 = 3 * (tmp1 - ((s_0_1 + s_2_1) << 1) + s_1_0 + s_1_2);
    const int tmp2 = s_0_2 - s_2_0;
    const int tmp3 = (s_0_0 + tmp2 - s_2_2);
    const int tmp4 = tmp3 - 2 * tmp2;
    const int coeff3 = -3 * (tmp3 + s_0_1 - s_2_1);
    const int coeff4 = -3 * (tmp4 + s_1_0 - s_1_2);
    const int coeff5 = (s_0_0 - s_0_2 - s_2_0 + s_2_2) * 4;
    const int coeff6 = (-(s_0_0 + s_0_2 - ((s_1_0 + s_0_1 + s_1_2 + s_2_1) * 2) - 5 * s_1_1 + s_2_0 + s_2_2)) * 2;
    const int H_det = 4 * coeff1 * coeff2 - coeff5 * coeff5;
    if (H_det == 0) { delta_x = 0.0f; delta_y = 0.0f; return __fdiv_rn((float)coeff6, 18.0f); }
    if (!(H_det > 0 && coeff1 < 0)) {
        int tmp_max = coeff3 + coeff4 + coeff5;
        delta_x = 1.0f; delta_y = 1.0f;
        int tmp = -coeff3 + coeff4 - coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; delta_x = -1.0f; delta_y = 1.0f; }
        tmp = coeff3 - coeff4 - coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; delta_x = 1.0f; delta_y = -1.0f; }
        tmp = -coeff3 - coeff4 + coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; delta_x = -1.0f; delta_y = -1.0f; }
        return __fdiv_rn((float)(tmp_max + coeff1 + coeff2 + coeff6), 18.0f);
    }
    delta_x = __fdiv_rn((float)(2 * coeff2 * coeff3 - coeff4 * coeff5), (float)(-H_det));
    delta_y = __fdiv_rn((float)(2 * coeff1 * coeff4 - coeff3 * coeff5), (float)(-H_det));
    bool tx = false, tx_ = false, ty = false, ty_ = false;
    if (delta_x > 1.0f) tx = true; else if (delta_x < -1.0f) tx_ = true;
    if (delta_y > 1.0f) ty = true;
    if (delta_y < -1.0f) ty_ = true;
    const float c1 = (float)coeff1, c2 = (float)coeff2, c3 = (float)coeff3, c4 = (float)coeff4, c5 = (float)coeff5, c6 = (float)coeff6;
    // ((((c1*dx*dx + c2*dy*dy) + c3*dx) + c4*dy) + c5*dx*dy) + c6, every product left to right
#define BRK_QUAD(dx, dy) __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(c1, dx), dx), __fmul_rn(__fmul_rn(c2, dy), dy)), \
                         __fmul_rn(c3, dx)), __fmul_rn(c4, dy)), __fmul_rn(__fmul_rn(c5, dx), dy)), c6), 18.0f)
    if (tx || tx_ || ty || ty_) {
        float dx1 = 0.0f, dx2 = 0.0f, dy1 = 0.0f, dy2 = 0.0f;
        if (tx) {
            dx1 = 1.0f; dy1 = __fdiv_rn(-(float)(coeff4 + coeff5), (float)(2 * coeff2));
            if (dy1 > 1.0f) dy1 = 1.0f; else if (dy1 < -1.0f) dy1 = -1.0f;
        } else if (tx_) {
            dx1 = -1.0f; dy1 = __fdiv_rn(-(float)(coeff4 - coeff5), (float)(2 * coeff2));
            if (dy1 > 1.0f) dy1 = 1.0f; else if (dy1 < -1.0f) dy1 = -1.0f;
        }
        if (ty) {
            dy2 = 1.0f; dx2 = __fdiv_rn(-(float)(coeff3 + coeff5), (float)(2 * coeff1));
            if (dx2 > 1.0f) dx2 = 1.0f; else if (dx2 < -1.0f) dx2 = -1.0f;
        } else if (ty_) {
            dy2 = -1.0f; dx2 = __fdiv_rn(-(float)(coeff3 - coeff5), (float)(2 * coeff1));
            if (dx2 > 1.0f) dx2 = 1.0f; else if (dx2 < -1.0f) dx2 = -1.0f;
        }
        const float max1 = BRK_QUAD(dx1, dy1);
        const float max2 = BRK_QUAD(dx2, dy2);
        if (max1 > max2) { delta_x = dx1; delta_y = dy1; return max1; }
        delta_x = dx2; delta_y = dy2; return max2;
    }
    const float dx = delta_x, dy = delta_y;
    return BRK_QUAD(dx, dy);
#undef BRK_QUAD
}

__device__ float brk_refine1d(int kind, float s_05, float s0, float s05, float& mx) {
    const int i_05 = (int)__dadd_rn(__dmul_rn(1024.0, (double)s_05), 0.5), i0 = (int)__dadd_rn(__dmul_rn(1024.0, (double)s0), 0.5),
              i05 = (int)__dadd_rn(__dmul_rn(1024.0, (double)s05), 0.5);
    int a, b, c; float lo, hi, dv;
    if (kind == 0) { a = 16 * i_05 - 24 * i0 + 8 * i05; b = -40 * i_05 + 54 * i0 - 14 * i05; c = 24 * i_05 - 27 * i0 + 6 * i05; lo = 0.75f; hi = 1.5f; dv = 3072.0f; }
    else if (kind == 1) { a = 9 * i_05 - 18 * i0 + 9 * i05; b = -21 * i_05 + 36 * i0 - 15 * i05; c = 12 * i_05 - 16 * i0 + 6 * i05; lo = 0.6666666666666666666666666667f; hi = 1.3333333333333333333333333333f; dv = 2048.0f; }
    else { a = 2 * i_05 - 4 * i0 + 2 * i05; b = -5 * i_05 + 8 * i0 - 3 * i05; c = 3 * i_05 - 3 * i0 + 1 * i05; lo = 0.7f; hi = 1.5f; dv = 1024.0f; }
    if (a >= 0) {
        if (s0 >= s_05 && s0 >= s05) { mx = s0; return 1.0f; }
        if (s_05 >= s0 && s_05 >= s05) { mx = s_05; return lo; }
        if (s05 >= s0 && s05 >= s_05) { mx = s05; return hi; }
    }
    float r = __fdiv_rn(-(float)b, (float)(2 * a));
    if (r < lo) r = lo; else if (r > hi) r = hi;
    mx = __fadd_rn(__fadd_rn((float)c, __fmul_rn(__fmul_rn((float)a, r), r)), __fmul_rn((float)b, r));
    mx = __fdiv_rn(mx, dv);
    return r;
}

__device__ __forceinline__ float brk_patch_fit(const BrkView& V, int x, int y, float& dx, float& dy) {
    const int s00 = brk_sc(V, x - 1, y - 1, 1), s10 = brk_sc(V, x, y - 1, 1), s20 = brk_sc(V, x + 1, y - 1, 1);
    const int s01 = brk_sc(V, x - 1, y, 1), s11 = brk_sc(V, x, y, 1), s21 = brk_sc(V, x + 1, y, 1);
    const int s02 = brk_sc(V, x - 1, y + 1, 1), s12 = brk_sc(V, x, y + 1, 1), s22 = brk_sc(V, x + 1, y + 1, 1);
    return brk_subpixel2d(s00, s01, s02, s10, s11, s12, s20, s21, s22, dx, dy);
}
__device__ __forceinline__ int brk_ring_sum(const BrkView& B, int x, int y) {
    return 2 * (brk_sc(B, x - 1, y, 1) + brk_sc(B, x + 1, y, 1) + brk_sc(B, x, y + 1, 1) + brk_sc(B, x, y - 1, 1)) +
           (brk_sc(B, x + 1, y + 1, 1) + brk_sc(B, x - 1, y + 1, 1) + brk_sc(B, x + 1, y - 1, 1) + brk_sc(B, x - 1, y - 1, 1));
}

// getScoreMaxAbove (below == false) / getScoreMaxBelow (below == true): scan the footprint of the 3x3 patch in the neighbouring
// layer on bilinear score samples; any sample above `threshold` (the centre score) in the first / middle rows rejects
__device__ float brk_max_neighbour(const BrkView& N, bool below, int layer, int x_layer, int y_layer, int threshold, bool& ismax, float& dx, float& dy) {
    ismax = false;
    float x_1, x1, y_1, y1;
    if (!below) {
        if (layer % 2 == 0) {
            x_1 = __fdiv_rn((float)(4 * x_layer - 1 - 2), 6.0f); x1 = __fdiv_rn((float)(4 * x_layer - 1 + 2), 6.0f);
            y_1 = __fdiv_rn((float)(4 * y_layer - 1 - 2), 6.0f); y1 = __fdiv_rn((float)(4 * y_layer - 1 + 2), 6.0f);
        } else {
            x_1 = __fdiv_rn((float)(6 * x_layer - 1 - 3), 8.0f); x1 = __fdiv_rn((float)(6 * x_layer - 1 + 3), 8.0f);
            y_1 = __fdiv_rn((float)(6 * y_layer - 1 - 3), 8.0f); y1 = __fdiv_rn((float)(6 * y_layer - 1 + 3), 8.0f);
        }
    } else {
        if (layer % 2 == 0) {
            x_1 = __fdiv_rn((float)(8 * x_layer + 1 - 4), 6.0f); x1 = __fdiv_rn((float)(8 * x_layer + 1 + 4), 6.0f);
            y_1 = __fdiv_rn((float)(8 * y_layer + 1 - 4), 6.0f); y1 = __fdiv_rn((float)(8 * y_layer + 1 + 4), 6.0f);
        } else {
            x_1 = __fdiv_rn((float)(6 * x_layer + 1 - 3), 4.0f); x1 = __fdiv_rn((float)(6 * x_layer + 1 + 3), 4.0f);
            y_1 = __fdiv_rn((float)(6 * y_layer + 1 - 3), 4.0f); y1 = __fdiv_rn((float)(6 * y_layer + 1 + 3), 4.0f);
        }
    }
    const float thr = (float)threshold;
    int max_x = (int)x_1 + 1, max_y = (int)y_1 + 1;
    float tmp_max;
    float mx = (float)brk_sc_f(N, x_1, y_1, 1);
    if (mx > thr) return 0.f;
    for (int x = (int)x_1 + 1; x <= (int)x1; ++x) {
        tmp_max = (float)brk_sc_f(N, (float)x, y_1, 1);
        if (tmp_max > thr) return 0.f;
        if (tmp_max > mx) { mx = tmp_max; max_x = x; }
    }
    tmp_max = (float)brk_sc_f(N, x1, y_1, 1);
    if (tmp_max > thr) return 0.f;
    if (tmp_max > mx) { mx = tmp_max; max_x = (int)x1; }
    for (int y = (int)y_1 + 1; y <= (int)y1; ++y) {
        tmp_max = (float)brk_sc_f(N, x_1, (float)y, 1);
        if (tmp_max > thr) return 0.f;
        if (tmp_max > mx) { mx = tmp_max; max_x = (int)__fadd_rn(x_1, 1.f); max_y = y; }
        for (int x = (int)x_1 + 1; x <= (int)x1; ++x) {
            tmp_max = (float)brk_sc(N, x, y, 1);
            if (tmp_max > thr) return 0.f;
            if (below && tmp_max == mx) {
                const int t1 = brk_ring_sum(N, x, y), t2 = brk_ring_sum(N, max_x, max_y);
                if (t1 > t2) { max_x = x; max_y = y; }
            }
            if (tmp_max > mx) { mx = tmp_max; max_x = x; max_y = y; }
        }
        tmp_max = (float)brk_sc_f(N, x1, (float)y, 1);
        if (tmp_max > thr) return 0.f;
        if (tmp_max > mx) { mx = tmp_max; max_x = (int)x1; max_y = y; }
    }
    tmp_max = (float)brk_sc_f(N, x_1, y1, 1);
    if (tmp_max > mx) { mx = tmp_max; max_x = (int)__fadd_rn(x_1, 1.f); max_y = (int)y1; }
    for (int x = (int)x_1 + 1; x <= (int)x1; ++x) {
        tmp_max = (float)brk_sc_f(N, (float)x, y1, 1);
        if (tmp_max > mx) { mx = tmp_max; max_x = x; max_y = (int)y1; }
    }
    tmp_max = (float)brk_sc_f(N, x1, y1, 1);
    if (tmp_max > mx) { mx = tmp_max; max_x = (int)x1; max_y = (int)y1; }

    float dx_1, dy_1;
    const float refined_max = brk_patch_fit(N, max_x, max_y, dx_1, dy_1);
    const float real_x = __fadd_rn((float)max_x, dx_1), real_y = __fadd_rn((float)max_y, dy_1);
    bool returnrefined = true;
    if (!below) {
        if (layer % 2 == 0) {
            dx = __fsub_rn(__fdiv_rn(__fadd_rn(__fmul_rn(real_x, 6.0f), 1.0f), 4.0f), (float)x_layer);
            dy = __fsub_rn(__fdiv_rn(__fadd_rn(__fmul_rn(real_y, 6.0f), 1.0f), 4.0f), (float)y_layer);
        } else {
            dx = __fsub_rn(__fdiv_rn(__fadd_rn(__fmul_rn(real_x, 8.0f), 1.0f), 6.0f), (float)x_layer);
            dy = __fsub_rn(__fdiv_rn(__fadd_rn(__fmul_rn(real_y, 8.0f), 1.0f), 6.0f), (float)y_layer);
        }
    } else {
        if (layer % 2 == 0) {
            dx = __fsub_rn((float)__ddiv_rn(__dadd_rn(__dmul_rn((double)real_x, 6.0), 1.0), 8.0), (float)x_layer);
            dy = __fsub_rn((float)__ddiv_rn(__dadd_rn(__dmul_rn((double)real_y, 6.0), 1.0), 8.0), (float)y_layer);
        } else {
            dx = __fsub_rn((float)__ddiv_rn(__dsub_rn(__dmul_rn((double)real_x, 4.0), 1.0), 6.0), (float)x_layer);
            dy = __fsub_rn((float)__ddiv_rn(__dsub_rn(__dmul_rn((double)real_y, 4.0), 1.0), 6.0), (float)y_layer);
        }
    }
    if (dx > 1.0f) { dx = 1.0f; returnrefined = false; }
    if (dx < -1.0f) { dx = -1.0f; returnrefined = false; }
    if (dy > 1.0f) { dy = 1.0f; returnrefined = false; }
    if (dy < -1.0f) { dy = -1.0f; returnrefined = false; }
    ismax = true;
    if (returnrefined) return refined_max > mx ? refined_max : mx;
    return mx;
}

__global__ void __launch_bounds__(128) k_brk_refine(const __grid_constant__ BrkParams P, int l) {
    const BrkLayerG& L = P.lv[l];
    const int f = blockIdx.y, k = blockIdx.x * 128 + threadIdx.x;
    const int n = min(P.cnt[f * 32 + 8 + l], L.cand_cap);
    if (k >= n) return;
    const uint32_t pk = (L.cand + (long long)f * L.cand_cap)[k];
    const int px = pk & 0xffff, py = pk >> 16;
    const BrkView T = brk_view(L, f);
    float ox, oy, osize, oresp;
    if (l == P.layers - 1) {
        const int center = brk_sc_f(T, (float)px, (float)py, P.threshold);
        const BrkView Bv = brk_view(P.lv[l - 1], f);
        bool ismax; float dx, dy;
        (void)brk_max_neighbour(Bv, true, l, px, py, center, ismax, dx, dy);
        if (!ismax) return;
        float delta_x, delta_y;
        const float mx = brk_patch_fit(T, px, py, delta_x, delta_y);
        ox = __fadd_rn(__fmul_rn(__fadd_rn((float)px, delta_x), L.scale), L.offset);
        oy = __fadd_rn(__fmul_rn(__fadd_rn((float)py, delta_y), L.scale), L.offset);
        osize = __fmul_rn(BRK_BASIC_SIZE, L.scale); oresp = mx;
    } else {
        const int center = brk_sc(T, px, py, 1);
        const BrkView Av = brk_view(P.lv[l + 1], f);
        bool ismax; float dxa = 0.f, dya = 0.f;
        const float max_above = brk_max_neighbour(Av, false, l, px, py, center, ismax, dxa, dya);
        if (!ismax) return;
        float dxb = 0.f, dyb = 0.f, max_below, mx, scale;
        if (l == 0) {
            const uint8_t* img = L.img + (long long)f * L.img_fstride;
            int q[9], mb = 0;
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int i = 0; i < 3; ++i) { q[3 * i + j] = brk_sc58(img, L.w, L.h, L.img_stride, px - 1 + i, py - 1 + j); mb = max(mb, q[3 * i + j]); }
            (void)brk_subpixel2d(q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], dxb, dyb);
            max_below = (float)mb;
        } else {
            const BrkView Bv = brk_view(P.lv[l - 1], f);
            max_below = brk_max_neighbour(Bv, true, l, px, py, center, ismax, dxb, dyb);
            if (!ismax) return;
        }
        float dxl, dyl;
        const float max_layer = brk_patch_fit(T, px, py, dxl, dyl);
        const float mid = (float)center > max_layer ? (float)center : max_layer;
        if (l % 2 == 0) {
            scale = brk_refine1d(l == 0 ? 2 : 0, max_below, mid, max_above, mx);
            if (scale > 1.0f) {
                const float r0 = __fdiv_rn(__fsub_rn(1.5f, scale), .5f), r1 = __fsub_rn(1.0f, r0);
                ox = __fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, dxl), __fmul_rn(r1, dxa)), (float)px), L.scale), L.offset);
                oy = __fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, dyl), __fmul_rn(r1, dya)), (float)py), L.scale), L.offset);
            } else if (l == 0) {
                const float r0 = __fdiv_rn(__fsub_rn(scale, 0.5f), 0.5f), r_1 = __fsub_rn(1.0f, r0);
                ox = __fadd_rn(__fadd_rn(__fmul_rn(r0, dxl), __fmul_rn(r_1, dxb)), (float)px);
                oy = __fadd_rn(__fadd_rn(__fmul_rn(r0, dyl), __fmul_rn(r_1, dyb)), (float)py);
            } else {
                const float r0 = __fdiv_rn(__fsub_rn(scale, 0.75f), 0.25f), r_1 = __fsub_rn(1.0f, r0);
                ox = __fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, dxl), __fmul_rn(r_1, dxb)), (float)px), L.scale), L.offset);
                oy = __fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, dyl), __fmul_rn(r_1, dyb)), (float)py), L.scale), L.offset);
            }
        } else {
            scale = brk_refine1d(1, max_below, mid, max_above, mx);
            if (scale > 1.0f) {
                const float r0 = __fsub_rn(4.0f, __fmul_rn(scale, 3.0f)), r1 = __fsub_rn(1.0f, r0);
                ox = __fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, dxl), __fmul_rn(r1, dxa)), (float)px), L.scale), L.offset);
                oy = __fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, dyl), __fmul_rn(r1, dya)), (float)py), L.scale), L.offset);
            } else {
                const float r0 = __fsub_rn(__fmul_rn(scale, 3.0f), 2.0f), r_1 = __fsub_rn(1.0f, r0);
                ox = __fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, dxl), __fmul_rn(r_1, dxb)), (float)px), L.scale), L.offset);
                oy = __fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, dyl), __fmul_rn(r_1, dyb)), (float)py), L.scale), L.offset);
            }
        }
        scale = __fmul_rn(scale, L.scale);
        if (!(mx > (float)P.threshold)) return;
        osize = __fmul_rn(BRK_BASIC_SIZE, scale); oresp = mx;
    }
    const int slot = atomicAdd(&P.cnt[f * 32 + 16 + l], 1);
    if (slot < L.kp_cap) {
        (L.kp + (long long)f * L.kp_cap)[slot] = make_float4(ox, oy, osize, oresp);
        (L.kkey + (long long)f * L.kp_cap)[slot] = (uint32_t)(py * L.w + px);
    } else atomicOr(&P.status[f], BRK_ST_LIST_OVERFLOW);
}

// ------------------------------------------------------------------------------------------------------
// per (layer, frame): detection order (raster key) by a bitonic sort, then DistributeOctTree
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_brk_octree(const __grid_constant__ BrkParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lv = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const BrkLayerG& L = P.lv[lv];
    const int M = min(P.cnt[f * 32 + 16 + lv], L.kp_cap);
    if (M == 0) { if (tid == 0) P.keepcnt[f * 16 + lv] = 0; return; }
    const long long base = ((long long)f * P.layers + lv) * BRK_KP_CAP;
    {
        unsigned long long* sk = reinterpret_cast<unsigned long long*>(smem_raw);
        const uint32_t* kkey = L.kkey + (long long)f * L.kp_cap;
        int np = 1;
        while (np < M) np <<= 1;
        for (int k = tid; k < np; k += 256) sk[k] = k < M ? (((unsigned long long)kkey[k] << 32) | (unsigned)k) : ~0ull;
        __syncthreads();
        for (int k = 2; k <= np; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < np; t += 256) {
                    const int txj = t ^ j;
                    if (txj > t) {
                        const unsigned long long a = sk[t], b = sk[txj];
                        const bool up = (t & k) == 0;
                        if ((a > b) == up) { sk[t] = b; sk[txj] = a; }
                    }
                }
                __syncthreads();
            }
        const float4* kp = L.kp + (long long)f * L.kp_cap;
        for (int k = tid; k < M; k += 256) {
            const float4 p = kp[(unsigned)(sk[k] & 0xffffffffu)];
            P.okp[base + k] = p; P.okx[base + k] = p.x; P.oky[base + k] = p.y; P.oresp[base + k] = brk_f2ord(p.w);
        }
        __syncthreads();
    }
    OctWork W;
    oct_carve(smem_raw, P.oct_ncap, W);
    bool overflow = false;
    int size = oct_distribute(W, P.okx + base, P.oky + base, P.knode + base, P.kquad + base, M, P.q_ext[lv], P.n_ini, P.hX, P.H,
                              P.oct_ncap, tid, overflow);
    if (overflow && tid == 0) atomicOr(&P.status[f], BRK_ST_OCTREE_OVERFLOW);
    unsigned long long* best = W.best;
    for (int p = tid; p < size; p += 256) best[p] = 0ull;
    __syncthreads();
    const unsigned short* knode = P.knode + base; const uint32_t* oresp = P.oresp + base;
    for (int k = tid; k < M; k += 256) atomicMax(&best[knode[k]], ((unsigned long long)oresp[k] << 32) | (unsigned long long)(0xffffffffu - (unsigned)k));
    __syncthreads();
    if (size > P.keep_cap) { if (tid == 0) atomicOr(&P.status[f], BRK_ST_OCTREE_OVERFLOW); size = P.keep_cap; }
    int* keep = P.keep + ((long long)f * P.layers + lv) * P.keep_cap;
    for (int p = tid; p < size; p += 256) keep[p] = (int)(0xffffffffu - (unsigned)(best[p] & 0xffffffffu));
    if (tid == 0) P.keepcnt[f * 16 + lv] = size;
}

// ------------------------------------------------------------------------------------------------------
// merge (layers ascending) + the descriptor extractor's border filter, order-preserving compaction
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_brk_merge(const __grid_constant__ BrkParams P, afv_keypoint* __restrict__ kps,
                                                   float* __restrict__ kpsize, int* __restrict__ n_out) {
    __shared__ int lstart[BRK_MAX_LAYERS + 1];
    __shared__ int wtot[8];
    __shared__ int s_base;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        int acc = 0;
        for (int l = 0; l < P.layers; ++l) { lstart[l] = acc; acc += P.keepcnt[f * 16 + l]; }
        lstart[P.layers] = acc; s_base = 0;
    }
    __syncthreads();
    const int total = lstart[P.layers];
    for (int j0 = 0; j0 < total; j0 += 256) {
        const int j = j0 + tid;
        bool keepit = false; float4 kp = make_float4(0, 0, 0, 0); int lv = 0, sidx = 0;
        if (j < total) {
            while (j >= lstart[lv + 1]) ++lv;
            const int key = P.keep[((long long)f * P.layers + lv) * P.keep_cap + (j - lstart[lv])];
            kp = P.okp[((long long)f * P.layers + lv) * BRK_KP_CAP + key];
            // pattern scale index: number of thresholds <= size (host-built from the original's log expression)
            int lo = 0, hi = BRK_SCALES - 1;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (kp.z >= P.size_thr[mid]) lo = mid; else hi = mid - 1; }
            sidx = lo;
            const int border = P.size_list[sidx];
            const float minX = (float)border, maxX = (float)(P.W - border), maxY = (float)(P.H - border);
            keepit = !((kp.x < minX) || (kp.x >= maxX) || (kp.y < minX) || (kp.y >= maxY));
        }
        const unsigned m = __ballot_sync(0xffffffffu, keepit);
        if (lane == 0) wtot[wid] = __popc(m);
        __syncthreads();
        int off = s_base, tot = 0;
        for (int w = 0; w < 8; ++w) { if (w < wid) off += wtot[w]; tot += wtot[w]; }
        off += __popc(m & ((1u << lane) - 1));
        if (keepit) {
            if (off < P.out_cap) {
                const long long o = (long long)f * P.out_cap + off;
                afv_keypoint k; k.x = kp.x; k.y = kp.y; k.size = kp.z; k.angle = -1.f; k.response = kp.w; k.octave = lv; k.class_id = -1;
                kps[o] = k;
                if (kpsize) kpsize[o] = P.size_norm[lv];
                P.kscale[o] = sidx;
            } else atomicOr(&P.status[f], BRK_ST_OUT_OVERFLOW);
        }
        __syncthreads();
        if (tid == 0) s_base += tot;
        __syncthreads();
    }
    if (tid == 0) n_out[f] = min(s_base, P.out_cap);
}

// ------------------------------------------------------------------------------------------------------
// cv::integral (CV_32S): I[y+1][x+1] = sum of img[0..y][0..x]; first row / column zero
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_brk_integral_rows(const __grid_constant__ BrkParams P) {
    __shared__ int wsum[8];
    __shared__ int carry;
    const int y = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const BrkLayerG& L = P.lv[0];
    const int iw = P.W + 1;
    int* I = P.integral + (long long)f * iw * (P.H + 1);
    if (y == P.H) { for (int x = tid; x < iw; x += 256) I[x] = 0; return; }       // row 0 of the integral
    const uint8_t* row = L.img + (long long)f * L.img_fstride + (long long)y * L.img_stride;
    int* out = I + (long long)(y + 1) * iw;
    if (tid == 0) { carry = 0; out[0] = 0; }
    __syncthreads();
    for (int x0 = 0; x0 < P.W; x0 += 1024) {
        const int x = x0 + 4 * tid;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (x + k < P.W) ? (int)row[x + k] : 0;
        v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
        int incl = v[3];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        int base = carry;
        for (int w = 0; w < wid; ++w) base += wsum[w];
        base += incl - v[3];
#pragma unroll
        for (int k = 0; k < 4; ++k) if (x + k < P.W) out[x + k + 1] = base + v[k];
        __syncthreads();
        if (tid == 255) carry = base + v[3];
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_brk_integral_cols(const __grid_constant__ BrkParams P) {
    const int x = blockIdx.x * 256 + threadIdx.x, f = blockIdx.y;
    const int iw = P.W + 1;
    if (x >= iw) return;
    int* I = P.integral + (long long)f * iw * (P.H + 1) + x;
    int acc = 0;
    for (int y = 1; y <= P.H; ++y) { acc += I[(long long)y * iw]; I[(long long)y * iw] = acc; }
}

// ------------------------------------------------------------------------------------------------------
// BRISK_Impl::smoothedIntensity
// ------------------------------------------------------------------------------------------------------
__device__ int brk_smoothed(const uint8_t* image, int istride, const int* integral, int integralcols, float key_x, float key_y,
                            float2 pp, int4 pt) {
    const float xf = __fadd_rn(pp.x, key_x), yf = __fadd_rn(pp.y, key_y);
    const float sigma_half = __int_as_float(pt.x);
    int ret_val;
    if (sigma_half < 0.5f) {
        const int x = (int)xf, y = (int)yf;
        const int r_x = (int)__fmul_rn(__fsub_rn(xf, (float)x), 1024.f), r_y = (int)__fmul_rn(__fsub_rn(yf, (float)y), 1024.f);
        const int r_x_1 = 1024 - r_x, r_y_1 = 1024 - r_y;
        const uint8_t* ptr = image + x + (long long)y * istride;
        ret_val = r_x_1 * r_y_1 * (int)ptr[0] + r_x * r_y_1 * (int)ptr[1] + r_x * r_y * (int)ptr[istride + 1] + r_x_1 * r_y * (int)ptr[istride];
        return (ret_val + 512) / 1024;
    }
    const int scaling = pt.y, scaling2 = pt.z;
    const float fs = (float)scaling;
    const float x_1 = __fsub_rn(xf, sigma_half), x1 = __fadd_rn(xf, sigma_half), y_1 = __fsub_rn(yf, sigma_half), y1 = __fadd_rn(yf, sigma_half);
    const int x_left = (int)__dadd_rn((double)x_1, 0.5), y_top = (int)__dadd_rn((double)y_1, 0.5);
    const int x_right = (int)__dadd_rn((double)x1, 0.5), y_bottom = (int)__dadd_rn((double)y1, 0.5);
    const float r_x_1 = __fadd_rn(__fsub_rn((float)x_left, x_1), 0.5f), r_y_1 = __fadd_rn(__fsub_rn((float)y_top, y_1), 0.5f);
    const float r_x1 = __fadd_rn(__fsub_rn(x1, (float)x_right), 0.5f), r_y1 = __fadd_rn(__fsub_rn(y1, (float)y_bottom), 0.5f);
    const int dx = x_right - x_left - 1, dy = y_bottom - y_top - 1;
    const int A = (int)__fmul_rn(__fmul_rn(r_x_1, r_y_1), fs), B = (int)__fmul_rn(__fmul_rn(r_x1, r_y_1), fs);
    const int C = (int)__fmul_rn(__fmul_rn(r_x1, r_y1), fs), D = (int)__fmul_rn(__fmul_rn(r_x_1, r_y1), fs);
    const int r_x_1_i = (int)__fmul_rn(r_x_1, fs), r_y_1_i = (int)__fmul_rn(r_y_1, fs);
    const int r_x1_i = (int)__fmul_rn(r_x1, fs), r_y1_i = (int)__fmul_rn(r_y1, fs);
    const uint8_t* p0 = image + x_left + (long long)y_top * istride;
    if (dx + dy > 2) {
        ret_val = A * (int)p0[0] + B * (int)p0[dx + 1] + C * (int)p0[(long long)(dy + 1) * istride + dx + 1] + D * (int)p0[(long long)(dy + 1) * istride];
        const int* pi = integral + x_left + (long long)integralcols * y_top + 1;
        const int tmp1 = pi[0], tmp2 = pi[dx];
        const int tmp3 = pi[integralcols + dx], tmp4 = pi[integralcols + dx + 1];
        const long long rb = (long long)(dy + 1) * integralcols;
        const int tmp5 = pi[rb + dx + 1], tmp6 = pi[rb + dx];
        const int tmp7 = pi[rb + integralcols + dx], tmp8 = pi[rb + integralcols];
        const int tmp9 = pi[rb], tmp10 = pi[rb - 1];
        const int tmp11 = pi[integralcols - 1], tmp12 = pi[integralcols];
        const int upper = (tmp3 - tmp2 + tmp1 - tmp12) * r_y_1_i;
        const int middle = (tmp6 - tmp3 + tmp12 - tmp9) * scaling;
        const int left = (tmp9 - tmp12 + tmp11 - tmp10) * r_x_1_i;
        const int right = (tmp5 - tmp4 + tmp3 - tmp6) * r_x1_i;
        const int bottom = (tmp7 - tmp6 + tmp9 - tmp8) * r_y1_i;
        return (ret_val + upper + middle + left + right + bottom + scaling2 / 2) / scaling2;
    }
    // small boxes: direct weighted sum, first row / middle rows / last row
    const uint8_t* ptr = p0;
    ret_val = A * (int)ptr[0];
    for (int i = 1; i <= dx; ++i) ret_val += r_y_1_i * (int)ptr[i];
    ret_val += B * (int)ptr[dx + 1];
    for (int j = 1; j <= dy; ++j) {
        ptr = p0 + (long long)j * istride;
        ret_val += r_x_1_i * (int)ptr[0];
        for (int i = 1; i <= dx; ++i) ret_val += (int)ptr[i] * scaling;
        ret_val += r_x1_i * (int)ptr[dx + 1];
    }
    ptr = p0 + (long long)(dy + 1) * istride;
    ret_val += D * (int)ptr[0];
    for (int i = 1; i <= dx; ++i) ret_val += r_y1_i * (int)ptr[i];
    ret_val += C * (int)ptr[dx + 1];
    return (ret_val + scaling2 / 2) / scaling2;
}

// atan2 in double from + - * / only (same operation sequence as orc_brisk_atan2): bit-identical on the CPU and here
__device__ double brk_atan2(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    if (ax == 0.0 && ay == 0.0) return 0.0;
    const bool swap = ay > ax;
    double t = swap ? __ddiv_rn(ax, ay) : __ddiv_rn(ay, ax);
    const double SQ3 = 1.7320508075688772, T12 = 0.2679491924311227, PI6 = 0.5235987755982989, PI2 = 1.5707963267948966, PI = 3.141592653589793;
    bool red = false;
    if (t > T12) { t = __ddiv_rn(__dsub_rn(__dmul_rn(t, SQ3), 1.0), __dadd_rn(SQ3, t)); red = true; }
    const double z = __dmul_rn(t, t);
    double p = 1.0 / 29.0;
#define BRK_AT(c) p = __dsub_rn((c), __dmul_rn(z, p));
    BRK_AT(1.0 / 27.0) BRK_AT(1.0 / 25.0) BRK_AT(1.0 / 23.0) BRK_AT(1.0 / 21.0) BRK_AT(1.0 / 19.0) BRK_AT(1.0 / 17.0) BRK_AT(1.0 / 15.0)
    BRK_AT(1.0 / 13.0) BRK_AT(1.0 / 11.0) BRK_AT(1.0 / 9.0) BRK_AT(1.0 / 7.0) BRK_AT(1.0 / 5.0) BRK_AT(1.0 / 3.0) BRK_AT(1.0)
#undef BRK_AT
    double a = __dmul_rn(t, p);
    if (red) a = __dadd_rn(a, PI6);
    if (swap) a = __dsub_rn(PI2, a);
    if (x < 0.0) a = __dsub_rn(PI, a);
    if (y < 0.0) a = -a;
    return a;
}

__global__ void __launch_bounds__(256) k_brk_describe(const __grid_constant__ BrkParams P, afv_keypoint* __restrict__ kps,
                                                      uint8_t* __restrict__ desc, const int* __restrict__ n_out) {
    __shared__ int s_val[8][64];
    const int f = blockIdx.y, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int j = blockIdx.x * 8 + wp;
    if (j >= n_out[f]) return;
    const long long o = (long long)f * P.out_cap + j;
    const float kx = kps[o].x, ky = kps[o].y;
    const int scale = P.kscale[o];
    const BrkLayerG& L0 = P.lv[0];
    const uint8_t* image = L0.img + (long long)f * L0.img_fstride;
    const int iw = P.W + 1;
    const int* integral = P.integral + (long long)f * iw * (P.H + 1);
    int* val = s_val[wp];
    const float2* pat0 = P.pat + ((long long)scale * BRK_NROT) * BRK_POINTS;
    const int4* ptab = P.ptab + scale * BRK_POINTS;
    for (int i = lane; i < BRK_POINTS; i += 32) val[i] = brk_smoothed(image, L0.img_stride, integral, iw, kx, ky, pat0[i], ptab[i]);
    __syncwarp();
    int d0 = 0, d1 = 0;
    for (int p = lane; p < BRK_NLONG; p += 32) {
        const int2 lp = P.longp[p];
        const int dt = val[lp.x & 0xff] - val[(lp.x >> 8) & 0xff];
        d0 += dt * (int)(short)(lp.y & 0xffff) / 1024;
        d1 += dt * (int)(short)(lp.y >> 16) / 1024;
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) { d0 += __shfl_xor_sync(0xffffffffu, d0, s); d1 += __shfl_xor_sync(0xffffffffu, d1, s); }
    float angle = (float)__dmul_rn(__ddiv_rn(brk_atan2((double)(float)d1, (double)(float)d0), 3.14159265358979323846), 180.0);
    int theta;
    if (angle == -1.f) theta = 0;
    else {
        theta = (int)__dadd_rn(__dmul_rn((double)BRK_NROT, __ddiv_rn((double)angle, 360.0)), 0.5);
        if (theta < 0) theta += BRK_NROT;
        if (theta >= BRK_NROT) theta -= BRK_NROT;
    }
    if (angle < 0) angle = __fadd_rn(angle, 360.f);
    __syncwarp();
    const float2* pat = pat0 + (long long)theta * BRK_POINTS;
    for (int i = lane; i < BRK_POINTS; i += 32) val[i] = brk_smoothed(image, L0.img_stride, integral, iw, kx, ky, pat[i], ptab[i]);
    __syncwarp();
    // 384 short pairs, bit p of the row = values[i] > values[j]; lane l owns byte l (and byte 32 + l for l < 16)
    for (int b = lane; b < 48; b += 32) {
        uint32_t byte = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t sp = P.shortp[8 * b + k];
            byte |= (uint32_t)(val[sp & 0xff] > val[sp >> 8]) << k;
        }
        desc[o * 48 + b] = (uint8_t)byte;
    }
    if (lane == 0) kps[o].angle = angle;
}

// debug tap: detect list of one frame in detection order (layer, raster): x, y, size, response, layer
__global__ void k_brk_tap_list(const __grid_constant__ BrkParams P, int f, float* out, int cap, int* n_total) {
    // single CTA: relies on the sorted copy written by k_brk_octree
    int acc = 0;
    for (int l = 0; l < P.layers; ++l) {
        const int M = min(P.cnt[f * 32 + 16 + l], P.lv[l].kp_cap);
        const long long base = ((long long)f * P.layers + l) * BRK_KP_CAP;
        for (int k = threadIdx.x; k < M; k += blockDim.x) {
            if (acc + k < cap) {
                const float4 p = P.okp[base + k];
                float* o = out + 5 * (long long)(acc + k);
                o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.w; o[4] = (float)l;
            }
        }
        acc += M;
    }
    if (threadIdx.x == 0) *n_total = acc;
}

// ======================================================================================================
// host side
// ======================================================================================================
struct BrkTables {                      // pattern tables: built once per process and device
    float2* pat = nullptr; int4* ptab = nullptr; int2* longp = nullptr; uint16_t* shortp = nullptr; float* size_thr = nullptr; int* size_list = nullptr;
    bool ready = false;
};
static std::mutex g_brk_mu;
static BrkTables g_brk_tab[64];

// size -> pattern scale index exactly as BRISK_Impl::computeDescriptorsAndOrOrientation (float log, see the oracle)
static int brk_scale_index_host(float size) {
    static const float log2c = 0.693147180559945f;
    const float lb_scalerange = (float)(logf(30.f) / (log2c));
    const float basicSize06 = BRK_BASIC_SIZE * 0.6f;
    int scale = (int)((float)BRK_SCALES / lb_scalerange * (logf(size / (basicSize06)) / log2c) + 0.5);
    if (scale < 0) scale = 0;
    if (scale >= BRK_SCALES) scale = BRK_SCALES - 1;
    return scale;
}

static int brk_build_tables(int dev, BrkTables** out) {
    std::lock_guard<std::mutex> lk(g_brk_mu);
    if (dev < 0 || dev >= 64) { afv_set_error("brisk48: device index out of range"); return AFV_ERR_INVALID; }
    BrkTables& T = g_brk_tab[dev];
    *out = &T;
    if (T.ready) return AFV_OK;
    // BRISK_Impl::generateKernel (radii 0.85 * {0, 2.9, 4.9, 7.4, 10.8}, 1 + 10 + 14 + 15 + 20 points, dMax 5.85, dMin 8.2)
    const float fsc = 0.85f * 1.0f;
    const float rList[5] = {(float)(fsc * 0.), (float)(fsc * 2.9), (float)(fsc * 4.9), (float)(fsc * 7.4), (float)(fsc * 10.8)};
    const int nList[5] = {1, 10, 14, 15, 20};
    const float dMax = 5.85f, dMin = 8.2f;
    std::vector<float2> pat((size_t)BRK_POINTS * BRK_SCALES * BRK_NROT);
    std::vector<float> sig((size_t)BRK_SCALES * BRK_POINTS);
    std::vector<int4> ptab((size_t)BRK_SCALES * BRK_POINTS);
    std::vector<int> size_list(BRK_SCALES);
    const float lb_scale = (float)((double)logf(30.f) / log(2.0));
    const float lb_scale_step = lb_scale / (float)BRK_SCALES;
    const float sigma_scale = 1.3f;
    size_t it = 0;
    for (unsigned scale = 0; scale < BRK_SCALES; ++scale) {
        const float sl = (float)pow(2.0, (double)((float)scale * lb_scale_step));
        unsigned maxsize = 0;
        for (unsigned rot = 0; rot < BRK_NROT; ++rot) {
            const double theta = (double)rot * 2 * M_PI / (double)BRK_NROT;
            int pt = 0;
            for (int ring = 0; ring < 5; ++ring)
                for (int num = 0; num < nList[ring]; ++num, ++pt) {
                    const double alpha = ((double)num) * 2 * M_PI / (double)nList[ring];
                    float2 p;
                    p.x = (float)((double)(sl * rList[ring]) * cos(alpha + theta));
                    p.y = (float)((double)(sl * rList[ring]) * sin(alpha + theta));
                    float sigma;
                    if (ring == 0) sigma = sigma_scale * sl * 0.5f;
                    else sigma = (float)((double)(sigma_scale * sl) * ((double)rList[ring]) * sin(M_PI / nList[ring]));
                    const unsigned size = (unsigned)((int)ceil((double)((sl * rList[ring]) + sigma)) + 1);
                    if (maxsize < size) maxsize = size;
                    pat[it++] = p;
                    if (rot == 0) {
                        sig[(size_t)scale * BRK_POINTS + pt] = sigma;
                        const float area = 4.0f * sigma * sigma;
                        int4 e; memcpy(&e.x, &sigma, 4);
                        e.y = (int)(4194304.0 / area);
                        e.z = (int)((float)e.y * area / 1024.0);
                        e.w = 0;
                        ptab[(size_t)scale * BRK_POINTS + pt] = e;
                    }
                }
        }
        size_list[scale] = (int)maxsize;
    }
    std::vector<int2> longp; std::vector<uint16_t> shortp; std::vector<std::pair<float, int>> sd;
    const float dMin_sq = dMin * dMin, dMax_sq = dMax * dMax;
    for (unsigned i = 1; i < BRK_POINTS; ++i)
        for (unsigned j = 0; j < i; ++j) {
            const float dx = pat[j].x - pat[i].x, dy = pat[j].y - pat[i].y;
            const float norm_sq = (dx * dx + dy * dy);
            if (norm_sq > dMin_sq) {
                const int wdx = (int)((double)(dx / (norm_sq)) * 2048.0 + 0.5), wdy = (int)((double)(dy / (norm_sq)) * 2048.0 + 0.5);
                int2 e; e.x = (int)(i | (j << 8)); e.y = (int)(((unsigned)wdx & 0xffffu) | ((unsigned)wdy << 16));
                longp.push_back(e);
            } else if (norm_sq < dMax_sq) {
                sd.push_back(std::make_pair(norm_sq, (int)shortp.size()));
                shortp.push_back((uint16_t)(i | (j << 8)));
            }
        }
    if ((int)longp.size() != BRK_NLONG || shortp.size() != 512) { afv_set_error("brisk48: internal: unexpected pair counts %zu / %zu", longp.size(), shortp.size()); return AFV_ERR_INVALID; }
    // 48-byte stand-in table: the 384 shortest of the 512 short pairs (ties by enumeration order), kept in enumeration order
    std::vector<std::pair<float, int>> srt = sd;
    std::stable_sort(srt.begin(), srt.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first < b.first || (a.first == b.first && a.second < b.second); });
    std::vector<char> take(shortp.size(), 0);
    for (int k = 0; k < BRK_NSHORT48; ++k) take[srt[k].second] = 1;
    std::vector<uint16_t> short48;
    for (size_t k = 0; k < shortp.size(); ++k) if (take[k]) short48.push_back(shortp[k]);
    // size thresholds: smallest float size whose scale index is >= k (the index is monotonic in the size)
    std::vector<float> thr(BRK_SCALES, 0.f);
    for (int k = 1; k < BRK_SCALES; ++k) {
        uint32_t lo = 0x3f800000u, hi = 0x49742400u;               // 1.0f .. 1e6f, positive floats order like their bit patterns
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2; float v; memcpy(&v, &mid, 4);
            if (brk_scale_index_host(v) >= k) hi = mid; else lo = mid + 1;
        }
        memcpy(&thr[k], &lo, 4);
    }
    cudaError_t e = cudaMalloc((void**)&T.pat, pat.size() * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc((void**)&T.ptab, ptab.size() * sizeof(int4));
    if (e == cudaSuccess) e = cudaMalloc((void**)&T.longp, longp.size() * sizeof(int2));
    if (e == cudaSuccess) e = cudaMalloc((void**)&T.shortp, short48.size() * sizeof(uint16_t));
    if (e == cudaSuccess) e = cudaMalloc((void**)&T.size_thr, thr.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&T.size_list, size_list.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(T.pat, pat.data(), pat.size() * sizeof(float2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(T.ptab, ptab.data(), ptab.size() * sizeof(int4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(T.longp, longp.data(), longp.size() * sizeof(int2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(T.shortp, short48.data(), short48.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(T.size_thr, thr.data(), thr.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(T.size_list, size_list.data(), size_list.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { afv_set_error("brisk48: pattern table upload failed: %s", cudaGetErrorString(e)); return AFV_ERR_CUDA; }
    T.ready = true;
    return AFV_OK;
}

struct AfvBrisk {
    int nfeatures, nlevels, max_batch, max_w, max_h, device;
    float scale_factor, detect_th;
    std::vector<void*> allocs;
    uint8_t* img[BRK_MAX_LAYERS]; uint8_t* score[BRK_MAX_LAYERS];
    uint32_t* agast[BRK_MAX_LAYERS]; uint32_t* cand[BRK_MAX_LAYERS]; float4* kp[BRK_MAX_LAYERS]; uint32_t* kkey[BRK_MAX_LAYERS];
    int* tab_i[BRK_MAX_LAYERS]; float* tab_f[BRK_MAX_LAYERS];
    int agast_cap[BRK_MAX_LAYERS], cand_cap[BRK_MAX_LAYERS], kp_cap[BRK_MAX_LAYERS], max_stride[BRK_MAX_LAYERS], max_lh[BRK_MAX_LAYERS];
    uint8_t* gray_stage;
    int* h_status;
    int cur_w, cur_h;
    size_t oct_smem;
    BrkParams P;
};

template <typename T>
static int brk_alloc(AfvBrisk* s, T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 256);
    if (e != cudaSuccess) { afv_set_error("brisk48: cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e)); return AFV_ERR_CUDA; }
    s->allocs.push_back(q);
    *p = (T*)q;
    return AFV_OK;
}

static void brk_layer_dims(int w, int h, int layers, int* lw, int* lh) {
    for (int i = 0; i < layers; ++i) {
        if (i == 0) { lw[i] = w; lh[i] = h; }
        else if (i == 1) { lw[i] = 2 * (w / 3); lh[i] = 2 * (h / 3); }
        else { lw[i] = lw[i - 2] / 2; lh[i] = lh[i - 2] / 2; }
    }
}

// cv::computeResizeAreaTab, <= 4 taps per output sample, zero-padded
static int brk_area_tab(int ssize, int dsize, std::vector<int>& si, std::vector<float>& al) {
    si.assign((size_t)dsize * 4, 0); al.assign((size_t)dsize * 4, 0.f);
    const double scale = 1.0 / ((double)dsize / (double)ssize);
    for (int dx = 0; dx < dsize; ++dx) {
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        const double cell = scale < ssize - fsx1 ? scale : ssize - fsx1;
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        if (sx2 > ssize - 1) sx2 = ssize - 1;
        if (sx1 > sx2) sx1 = sx2;
        int k = 0;
        auto push = [&](int s, float a) { if (k >= 4) return false; si[(size_t)dx * 4 + k] = s; al[(size_t)dx * 4 + k] = a; ++k; return true; };
        bool ok = true;
        if (sx1 - fsx1 > 1e-3) ok = ok && push(sx1 - 1, (float)((sx1 - fsx1) / cell));
        for (int sx = sx1; sx < sx2; ++sx) ok = ok && push(sx, (float)(1.0 / cell));
        if (fsx2 - sx2 > 1e-3) { double a = fsx2 - sx2; if (a > 1.0) a = 1.0; if (a > cell) a = cell; ok = ok && push(sx2, (float)(a / cell)); }
        if (!ok) return -1;
    }
    return 0;
}

int afv_brisk_create(AfvBrisk** out, int nfeatures, int nlevels, float scale_factor, float detect_th, int max_batch, int max_w, int max_h) {
    *out = nullptr;
    const int layers = 2 * (nlevels / 2);               // BriskFeatureDetector(th, nOctaves / 2): 2 * octaves layers
    if (layers < 2 || layers > BRK_MAX_LAYERS) { afv_set_error("brisk48: n_octaves must be 2..%d (got %d)", BRK_MAX_LAYERS + 1, nlevels); return AFV_ERR_INVALID; }
    if ((int)detect_th < 1 || (int)detect_th > 254) { afv_set_error("brisk48: detection threshold %d outside 1..254", (int)detect_th); return AFV_ERR_INVALID; }
    if (max_w > 65535 || max_h > 65535) { afv_set_error("brisk48: frame dimension > 65535 not supported"); return AFV_ERR_INVALID; }
    AfvBrisk* s = new AfvBrisk();
    s->nfeatures = nfeatures; s->nlevels = nlevels; s->max_batch = max_batch; s->max_w = max_w; s->max_h = max_h;
    s->scale_factor = scale_factor; s->detect_th = detect_th; s->cur_w = s->cur_h = 0; s->h_status = nullptr; s->gray_stage = nullptr;
    memset(&s->P, 0, sizeof(s->P));
    cudaGetDevice(&s->device);
    BrkTables* T = nullptr;
    int rc = brk_build_tables(s->device, &T);
    if (rc) { delete s; return rc; }
    BrkParams& P = s->P;
    P.layers = layers; P.nlevels = nlevels; P.threshold = (int)detect_th;
    P.pat = T->pat; P.ptab = T->ptab; P.longp = T->longp; P.shortp = T->shortp; P.size_thr = T->size_thr; P.size_list = T->size_list;
    const size_t B = (size_t)max_batch;
    int lw[BRK_MAX_LAYERS], lh[BRK_MAX_LAYERS];
    brk_layer_dims(max_w, max_h, layers, lw, lh);
    for (int i = 0; i < layers && rc == AFV_OK; ++i) {
        s->max_stride[i] = ((lw[i] + 15) & ~15) + 16; s->max_lh[i] = lh[i] + 2;
        const size_t bytes = (size_t)s->max_stride[i] * s->max_lh[i];
        int ac = lw[i] * lh[i] / 8; if (ac < 1024) ac = 1024;
        s->agast_cap[i] = ac; s->cand_cap[i] = ac / 2;
        s->kp_cap[i] = s->cand_cap[i] < BRK_KP_CAP ? s->cand_cap[i] : BRK_KP_CAP;
        s->img[i] = nullptr;
        if (i > 0) rc = brk_alloc(s, &s->img[i], bytes * B);
        if (rc == AFV_OK) rc = brk_alloc(s, &s->score[i], bytes * B);
        if (rc == AFV_OK) rc = brk_alloc(s, &s->agast[i], (size_t)s->agast_cap[i] * B);
        if (rc == AFV_OK) rc = brk_alloc(s, &s->cand[i], (size_t)s->cand_cap[i] * B);
        if (rc == AFV_OK) rc = brk_alloc(s, &s->kp[i], (size_t)s->kp_cap[i] * B);
        if (rc == AFV_OK) rc = brk_alloc(s, &s->kkey[i], (size_t)s->kp_cap[i] * B);
        if (rc == AFV_OK) rc = brk_alloc(s, &s->tab_i[i], (size_t)(lw[i] + lh[i] + 8) * 4);
        if (rc == AFV_OK) rc = brk_alloc(s, &s->tab_f[i], (size_t)(lw[i] + lh[i] + 8) * 4);
    }
    int maxq = 0;
    if (rc == AFV_OK) {
        // mnFeaturesPerLevel (reference src/FeatureExtractor.cpp:97-108) and computeSize (:132-142) with powf(scaleFactor0, octave)
        float factor = 1.0f / scale_factor;
        float nDesired = (float)nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nlevels - 1; ++l) { P.q_ext[l] = (int)lrintf(nDesired); sum += P.q_ext[l]; nDesired *= factor; }
        P.q_ext[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
        const float maxSize0 = powf(1.2f, (float)(8 - 1.0)), maxSize = maxSize0, minSize = 1.0f;
        for (int l = 0; l < nlevels; ++l) {
            const float sz = powf(scale_factor, (float)l);
            float sn = maxSize;
            if (maxSize > minSize) sn = 1.0f + (sz - minSize) * (maxSize0 - 1.0f) / (maxSize - minSize);
            P.size_norm[l] = sn;
            if (P.q_ext[l] > maxq) maxq = P.q_ext[l];
        }
        P.keep_cap = maxq + 8; P.oct_ncap = maxq + 16;
        s->oct_smem = oct_work_bytes(P.oct_ncap);
        if (s->oct_smem < 8 * (size_t)BRK_KP_CAP) s->oct_smem = 8 * (size_t)BRK_KP_CAP;
        if (s->oct_smem > 227 * 1024) { afv_set_error("brisk48: nfeatures too large for the octree workspace"); rc = AFV_ERR_INVALID; }
    }
    const size_t LK = (size_t)layers * BRK_KP_CAP;
    if (rc == AFV_OK) rc = brk_alloc(s, &P.cnt, 32 * B);
    if (rc == AFV_OK) rc = brk_alloc(s, &P.status, B);
    if (rc == AFV_OK) rc = brk_alloc(s, &P.okx, LK * B);
    if (rc == AFV_OK) rc = brk_alloc(s, &P.oky, LK * B);
    if (rc == AFV_OK) rc = brk_alloc(s, &P.oresp, LK * B);
    if (rc == AFV_OK) rc = brk_alloc(s, &P.okp, LK * B);
    if (rc == AFV_OK) rc = brk_alloc(s, &P.knode, LK * B);
    if (rc == AFV_OK) rc = brk_alloc(s, &P.kquad, LK * B);
    if (rc == AFV_OK) rc = brk_alloc(s, &P.keep, (size_t)P.keep_cap * layers * B);
    if (rc == AFV_OK) rc = brk_alloc(s, &P.keepcnt, 16 * B);
    if (rc == AFV_OK) rc = brk_alloc(s, &P.integral, (size_t)(max_w + 1) * (max_h + 1) * B);
    if (rc == AFV_OK) rc = brk_alloc(s, &P.kscale, (size_t)(nfeatures + 3 * nlevels + 64) * B);
    if (rc == AFV_OK) rc = brk_alloc(s, &s->gray_stage, (size_t)max_w * max_h * B);
    if (rc == AFV_OK) {
        cudaError_t e = cudaMallocHost((void**)&s->h_status, sizeof(int) * B);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_brk_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->oct_smem);
        if (e != cudaSuccess) { afv_set_error("brisk48: setup failed: %s", cudaGetErrorString(e)); rc = AFV_ERR_CUDA; }
    }
    if (rc != AFV_OK) { afv_brisk_destroy(s); return rc; }
    *out = s;
    return AFV_OK;
}

void afv_brisk_destroy(AfvBrisk* s) {
    if (!s) return;
    for (void* p : s->allocs) cudaFree(p);
    if (s->h_status) cudaFreeHost(s->h_status);
    delete s;
}

uint8_t* afv_brisk_stage(AfvBrisk* s) { return s->gray_stage; }

static int brk_configure(AfvBrisk* s, int w, int h) {
    if (w > s->max_w || h > s->max_h || w < 64 || h < 64) {
        afv_set_error("frame %dx%d outside the extractor's configured range (64..%d x 64..%d)", w, h, s->max_w, s->max_h);
        return AFV_ERR_INVALID;
    }
    BrkParams& P = s->P;
    if (w == s->cur_w && h == s->cur_h) return AFV_OK;
    P.W = w; P.H = h;
    P.n_ini = (int)round((double)((float)w / (float)h));
    if (P.n_ini < 1) { afv_set_error("portrait frames with w/h < 0.5 are not supported (reference divides by zero)"); return AFV_ERR_INVALID; }
    P.hX = (float)w / (float)P.n_ini;
    int lw[BRK_MAX_LAYERS], lh[BRK_MAX_LAYERS];
    brk_layer_dims(w, h, P.layers, lw, lh);
    for (int i = 0; i < P.layers; ++i) {
        if (lw[i] < 8 || lh[i] < 8) { afv_set_error("brisk48: frame %dx%d too small for %d layers", w, h, P.layers); return AFV_ERR_INVALID; }
        BrkLayerG& L = P.lv[i];
        L.w = lw[i]; L.h = lh[i]; L.stride = (lw[i] + 15) & ~15; L.fstride = (long long)L.stride * lh[i];
        if (L.stride > s->max_stride[i] || lh[i] > s->max_lh[i]) { afv_set_error("brisk48: internal: layer %d larger than its arena", i); return AFV_ERR_INVALID; }
        L.img = s->img[i]; L.img_stride = L.stride; L.img_fstride = L.fstride;
        L.score = s->score[i];
        if (i == 0) { L.scale = 1.0f; L.offset = 0.0f; }
        else if (i == 1) { L.scale = 1.5f; L.offset = 0.5f * L.scale - 0.5f; }
        else { L.scale = P.lv[i - 2].scale * 2.0f; L.offset = 0.5f * L.scale - 0.5f; }
        L.agast_cap = s->agast_cap[i]; L.cand_cap = s->cand_cap[i]; L.kp_cap = s->kp_cap[i];
        L.agast = s->agast[i]; L.cand = s->cand[i]; L.kp = s->kp[i]; L.kkey = s->kkey[i];
        L.src = i == 0 ? 0 : (i == 1 ? 0 : i - 2);
        L.exact_half = 0; L.xsi = L.ysi = nullptr; L.xal = L.yal = nullptr;
        if (i > 0) {
            const int sw = lw[L.src], sh = lh[L.src];
            if (sw == 2 * lw[i] && sh == 2 * lh[i]) L.exact_half = 1;
            else {
                std::vector<int> xs, ys; std::vector<float> xa, ya;
                if (brk_area_tab(sw, lw[i], xs, xa) || brk_area_tab(sh, lh[i], ys, ya)) { afv_set_error("brisk48: internal: INTER_AREA footprint wider than 4 taps"); return AFV_ERR_INVALID; }
                AFV_CUDA_CHECK(cudaMemcpy(s->tab_i[i], xs.data(), xs.size() * 4, cudaMemcpyHostToDevice));
                AFV_CUDA_CHECK(cudaMemcpy(s->tab_i[i] + xs.size(), ys.data(), ys.size() * 4, cudaMemcpyHostToDevice));
                AFV_CUDA_CHECK(cudaMemcpy(s->tab_f[i], xa.data(), xa.size() * 4, cudaMemcpyHostToDevice));
                AFV_CUDA_CHECK(cudaMemcpy(s->tab_f[i] + xa.size(), ya.data(), ya.size() * 4, cudaMemcpyHostToDevice));
                L.xsi = s->tab_i[i]; L.ysi = s->tab_i[i] + xs.size(); L.xal = s->tab_f[i]; L.yal = s->tab_f[i] + xa.size();
            }
        }
    }
    s->cur_w = w; s->cur_h = h;
    return AFV_OK;
}

int afv_brisk_run(AfvBrisk* s, const uint8_t* d_gray, int B, int w, int h, int stride, long frame_stride, afv_keypoint* d_kps,
                  uint8_t* d_desc, float* d_kpsize, int cap, int* d_n_out, cudaStream_t st) {
    if (B < 1 || B > s->max_batch) { afv_set_error("batch %d outside 1..%d", B, s->max_batch); return AFV_ERR_INVALID; }
    if (cap > s->nfeatures + 3 * s->nlevels + 64) { afv_set_error("brisk48: cap %d larger than the extractor's output arena (%d)", cap, s->nfeatures + 3 * s->nlevels + 64); return AFV_ERR_INVALID; }
    int rc = brk_configure(s, w, h);
    if (rc) return rc;
    BrkParams P = s->P;
    P.B = B; P.out_cap = cap;
    P.lv[0].img = d_gray; P.lv[0].img_stride = stride; P.lv[0].img_fstride = frame_stride;
    AFV_CUDA_CHECK(cudaMemsetAsync(P.cnt, 0, sizeof(int) * 32 * B, st));
    AFV_CUDA_CHECK(cudaMemsetAsync(P.status, 0, sizeof(int) * B, st));
    { AfvProfScope ps("k_brk_resize", st);
      for (int i = 1; i < P.layers; ++i) {
          const BrkLayerG& L = P.lv[i];
          k_brk_resize<<<dim3((L.w + 63) / 64, (L.h + 3) / 4, B), 256, 0, st>>>(P, i); ++g_afv_launches;
      } }
    { AfvProfScope ps("k_brk_score", st);
      for (int i = 0; i < P.layers; ++i) {
          const BrkLayerG& L = P.lv[i];
          k_brk_score<<<dim3((L.w + BS_W - 1) / BS_W, (L.h + BS_H - 1) / BS_H, B), 256, 0, st>>>(P, i); ++g_afv_launches;
      } }
    { AfvProfScope ps("k_brk_ismax", st);
      for (int i = 0; i < P.layers; ++i) { k_brk_ismax<<<dim3((P.lv[i].agast_cap + 255) / 256, B), 256, 0, st>>>(P, i); ++g_afv_launches; } }
    { AfvProfScope ps("k_brk_refine", st);
      for (int i = 0; i < P.layers; ++i) { k_brk_refine<<<dim3((P.lv[i].cand_cap + 127) / 128, B), 128, 0, st>>>(P, i); ++g_afv_launches; } }
    { AfvProfScope ps("k_brk_octree", st); k_brk_octree<<<dim3(P.layers, B), 256, s->oct_smem, st>>>(P); ++g_afv_launches; }
    { AfvProfScope ps("k_brk_merge", st); k_brk_merge<<<B, 256, 0, st>>>(P, d_kps, d_kpsize, d_n_out); ++g_afv_launches; }
    { AfvProfScope ps("k_brk_integral", st);
      k_brk_integral_rows<<<dim3(h + 1, B), 256, 0, st>>>(P); ++g_afv_launches;
      k_brk_integral_cols<<<dim3((w + 1 + 255) / 256, B), 256, 0, st>>>(P); ++g_afv_launches; }
    { AfvProfScope ps("k_brk_describe", st); k_brk_describe<<<dim3((cap + 7) / 8, B), 256, 0, st>>>(P, d_kps, d_desc, d_n_out); ++g_afv_launches; }
    AFV_CUDA_CHECK(cudaGetLastError());
    s->P.B = B; s->P.out_cap = cap;
    s->P.lv[0].img = d_gray; s->P.lv[0].img_stride = stride; s->P.lv[0].img_fstride = frame_stride;
    return AFV_OK;
}

int afv_brisk_status(AfvBrisk* s, int B, cudaStream_t st) {
    AFV_CUDA_CHECK(cudaMemcpyAsync(s->h_status, s->P.status, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    AFV_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b)
        if (s->h_status[b]) {
            afv_set_error("brisk48: capacity exceeded in frame %d (flags 0x%x: 1 detection / candidate / keypoint lists, 4 caller cap, 8 octree)", b, s->h_status[b]);
            return AFV_ERR_CAPACITY;
        }
    return AFV_OK;
}

// taps: what = 30 layer image (u8, tight), 31 true score image, 32 detect list in detection order (x, y, size, response, layer floats)
int afv_brisk_debug_read(AfvBrisk* s, int what, int frame, int level, void* out, long cap_bytes, long* n_bytes) {
    const BrkParams& P = s->P;
    if (what == 30 || what == 31) {
        if (level < 0 || level >= P.layers) { afv_set_error("afv_debug_read: bad brisk layer"); return AFV_ERR_INVALID; }
        const BrkLayerG& L = P.lv[level];
        const long need = (long)L.w * L.h;
        if (cap_bytes < need) { afv_set_error("buffer too small"); return AFV_ERR_INVALID; }
        if (what == 30) AFV_CUDA_CHECK(cudaMemcpy2D(out, L.w, L.img + (long long)frame * L.img_fstride, L.img_stride, L.w, L.h, cudaMemcpyDeviceToHost));
        else AFV_CUDA_CHECK(cudaMemcpy2D(out, L.w, L.score + (long long)frame * L.fstride, L.stride, L.w, L.h, cudaMemcpyDeviceToHost));
        *n_bytes = need;
        return AFV_OK;
    }
    if (what == 32) {
        const int cap = (int)(cap_bytes / 20);
        float* d_out = nullptr; int* d_n = nullptr;
        AFV_CUDA_CHECK(cudaMalloc((void**)&d_out, (size_t)(cap > 0 ? cap : 1) * 20));
        AFV_CUDA_CHECK(cudaMalloc((void**)&d_n, 4));
        k_brk_tap_list<<<1, 256>>>(P, frame, d_out, cap, d_n);
        int n = 0;
        cudaError_t e = cudaMemcpy(&n, d_n, 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && n <= cap) e = cudaMemcpy(out, d_out, (size_t)n * 20, cudaMemcpyDeviceToHost);
        cudaFree(d_out); cudaFree(d_n);
        if (e != cudaSuccess) { afv_set_error("tap copy failed: %s", cudaGetErrorString(e)); return AFV_ERR_CUDA; }
        if (n > cap) { afv_set_error("buffer too small (%d keypoints)", n); return AFV_ERR_INVALID; }
        *n_bytes = (long)n * 20;
        return AFV_OK;
    }
    afv_set_error("unknown brisk tap %d", what);
    return AFV_ERR_INVALID;
}

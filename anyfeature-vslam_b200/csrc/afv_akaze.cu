// afv_akaze.cu -- akaze61 extraction (FeatureExtractor_akaze61, reference src/Feature_akaze61.cpp:7-77) for sm_100a.
//
// The reference delegates the arithmetic to libAKAZE (not vendored); this file implements the published A-KAZE algorithm
// in libAKAZE's structure with the options the reference sets (omax = nOctaves/4, nsublevels = nOctaves/2, dthreshold =
// detectionTh, PM_G2, MLDB-486) plus what the reference does around it (octave := class_id, DistributeOctTree per
// evolution level, levels merged before the descriptors, computeSize).  Arithmetic contract = oracle/afv_oracle_akaze.c
// (IEEE float32 without contraction: this TU is compiled --fmad=false), results are bit-identical with that oracle.
// PARITY vs libAKAZE itself is UNPINNED (cv2.AKAZE family check in tests/test_oracle_akaze.py).
//
// Kernels (batch of B frames; every image is [frame][row][stride] floats):
//   k_afv_blur<R>      Gaussian (sigma 1.6 base, sigma 1.0 per level), shared with sift128 (afv_blur.cuh)
//   k_akz_mag / k_akz_hist / k_akz_kcontrast   contrast factor: 70th percentile of the gradient magnitude histogram
//   k_akz_half         2x2 area half-sampling at an octave change
//   k_akz_flow         Scharr gradient of the smoothed level + Perona-Malik g2 conductivity
//   k_akz_nld          one explicit (FED) diffusion step
//   k_akz_deriv1/2     scaled Scharr first / second derivatives, determinant of the Hessian
//   k_akz_extrema      3x3 maxima over threshold, border test -> candidate list; k_akz_sort raster order
//   k_akz_select       CTA per frame: libAKAZE's sequential duplicate suppression (same / lower level, then upper level),
//                      sub-pixel refinement, stable partition per level
//   k_akz_octree       DistributeOctTree per (level, frame) (afv_octree.cuh)
//   k_akz_describe     warp per kept keypoint: main orientation (109 samples, 42 sliding windows, integer sums) + MLDB
#include "afv_common.cuh"
#include "afv_octree.cuh"
#include "afv_blur.cuh"
#include "afv_akaze.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define AKZ_MAX_LV 8
#define AKZ_PI 3.14159265358979323846f
#define AKZ_2PI 6.28318530717958647692f
#define AKZ_LIST_CAP 8192          // upper bound of the per-frame candidate list (shared memory of k_akz_select)

struct AkzLevelG {
    int w, h, stride; long long istride;
    int octave, sigma_size, nsteps, cand_cap;
    float esigma, psize;
    float tau[16];
    float *Lt, *Lsm, *Lx, *Ly, *Ldet;        // [B] images
    uint2* cand; uint2* srt;                 // [B][cand_cap] {y << 16 | x, response bits}
};

struct AkzParams {
    int B, nl, W, H, nfeatures, nlevels, out_cap;
    int n_ini; float hX, dth;
    int q_ext[AFV_MAX_LEVELS]; float size_norm[AFV_MAX_LEVELS];
    AkzLevelG lv[AKZ_MAX_LV];
    unsigned* hmax; int* hist; float* kcontrast;       // [B], [B][304], [B]
    int* cnt; int* status;                             // [B][16], [B]
    int key_cap;
    float4* kpt; int* kcls;                            // [B][key_cap] refined keypoints in list order {x, y, size, response}
    float* okx; float* oky; uint32_t* oresp; int* oidx; unsigned short* knode; unsigned char* kquad;     // [B][key_cap]
    int* selinfo;                                      // [B][32]
    int* keep; int keep_cap; int* keepcnt; int oct_ncap;
    int list_cap;                                      // capacity of k_akz_select's list for this geometry (<= AKZ_LIST_CAP)
};

#define AKZ_ST_CAND_OVERFLOW 1
#define AKZ_ST_OUT_OVERFLOW 4
#define AKZ_ST_OCTREE_OVERFLOW 8

__constant__ uchar4 c_mldb_bits[488];       // (cell a, cell b, channel, 0) per descriptor bit
__constant__ signed char c_ori_ij[109][2];  // (i, j) of the 109 orientation samples, libAKAZE loop order

__device__ __forceinline__ int akz_refl(int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * n - 2 - i; return min(max(i, 0), n - 1); }
__device__ __forceinline__ int akz_fround(float v) { return (int)floorf(v + 0.5f); }

__device__ __forceinline__ float akz_exp2(float t) {
    if (t < -126.f) t = -126.f;
    if (t > 126.f) t = 126.f;
    const float n = rintf(t), f = t - n;
    float p = 0x1.430912p-13f;
    p = p * f + 0x1.5d87fep-10f; p = p * f + 0x1.3b2ab6p-7f; p = p * f + 0x1.c6b08ep-5f;
    p = p * f + 0x1.ebfbep-3f; p = p * f + 0x1.62e43p-1f; p = p * f + 1.0f;
    return p * __uint_as_float((unsigned)((int)n + 127) << 23);
}
__device__ __forceinline__ float akz_gauss25(int i, int j) {
    return 0x1.a13714p-6f * akz_exp2((float)(i * i + j * j) * (-0.08f * 0x1.715476p+0f));
}
__device__ __forceinline__ float akz_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
    if (mx == 0.f) return 0.f;
    const float a = mn / mx, z = a * a;
    float p = -0x1.394942p-8f;
    p = p * z + 0x1.9256c4p-6f; p = p * z + -0x1.eabc6cp-5f; p = p * z + 0x1.974118p-4f; p = p * z + -0x1.1f5284p-3f;
    p = p * z + 0x1.990384p-3f; p = p * z + -0x1.555216p-2f; p = p * z + 0x1.fffffep-1f;
    float r = p * a;
    if (ay > ax) r = 0x1.921fb6p+0f - r;
    if (x < 0.f) r = AKZ_PI - r;
    if (y < 0.f) r = AKZ_2PI - r;
    if (r >= AKZ_2PI) r = r - AKZ_2PI;
    if (r < 0.f) r = 0.f;
    return r;
}
__device__ __forceinline__ void akz_sincos(float a, float* sn, float* cs) {
    const float q = rintf(a * 0x1.45f306p-1f);
    const float r = a - q * 0x1.921fb6p+0f;
    const float z = r * r;
    float s = 0x1.71de3ap-19f;
    s = s * z + -0x1.a01a02p-13f; s = s * z + 0x1.111112p-7f; s = s * z + -0x1.555556p-3f; s = s * z + 1.0f; s = s * r;
    float c = 0x1.a01a02p-16f;
    c = c * z + -0x1.6c16c2p-10f; c = c * z + 0x1.555556p-5f; c = c * z + -0.5f; c = c * z + 1.0f;
    const int k = ((int)q) & 3;
    if (k == 0) { *sn = s; *cs = c; } else if (k == 1) { *sn = c; *cs = -s; } else if (k == 2) { *sn = -s; *cs = -c; } else { *sn = -c; *cs = s; }
}

#define AKZ_PIX()  const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6), f = blockIdx.z

// cv::Scharr 3x3 (un-normalised, BORDER_REFLECT_101) at (x, y)
__device__ __forceinline__ void akz_scharr3(const float* __restrict__ s, int st, int w, int h, int x, int y, float* lx, float* ly) {
    const float* r0 = s + (long long)akz_refl(y - 1, h) * st; const float* r1 = s + (long long)y * st; const float* r2 = s + (long long)akz_refl(y + 1, h) * st;
    const int xm = akz_refl(x - 1, w), xp = akz_refl(x + 1, w);
    *lx = 3.0f * ((r0[xp] - r0[xm]) + (r2[xp] - r2[xm])) + 10.0f * (r1[xp] - r1[xm]);
    *ly = 3.0f * ((r2[xm] - r0[xm]) + (r2[xp] - r0[xp])) + 10.0f * (r2[x] - r0[x]);
}

// gradient magnitude of the sigma-1 smoothed input + per-frame maximum over the interior (compute_k_percentile, part 1)
__global__ void __launch_bounds__(256) k_akz_mag(const float* __restrict__ g, float* __restrict__ mag, int w, int h, int st, long long ist,
                                                 unsigned* __restrict__ hmax) {
    AKZ_PIX();
    float m = 0.f;
    if (x >= 1 && x < w - 1 && y >= 1 && y < h - 1) {
        float lx, ly;
        akz_scharr3(g + f * ist, st, w, h, x, y, &lx, &ly);
        m = sqrtf(lx * lx + ly * ly);
        mag[f * ist + (long long)y * st + x] = m;
    }
    unsigned u = __float_as_uint(m);                    // m >= 0: the bit pattern orders like the value
    u = __reduce_max_sync(0xffffffffu, u);
    if ((threadIdx.x & 31) == 0 && u) atomicMax(&hmax[f], u);
}

__global__ void __launch_bounds__(256) k_akz_hist(const float* __restrict__ mag, int w, int h, int st, long long ist,
                                                  const unsigned* __restrict__ hmax, int* __restrict__ hist) {
    __shared__ int sh[304];
    for (int i = threadIdx.x; i < 304; i += 256) sh[i] = 0;
    __syncthreads();
    AKZ_PIX();
    int nbin = -1;
    if (x >= 1 && x < w - 1 && y >= 1 && y < h - 1) {
        const float m = mag[f * ist + (long long)y * st + x];
        if (m != 0.0f) {
            nbin = (int)floorf(300.0f * (m / __uint_as_float(hmax[f])));
            if (nbin == 300) nbin = 299;
        }
    }
    // most magnitudes fall into a few low bins: one shared atomic per distinct bin per warp
    const unsigned peers = __match_any_sync(0xffffffffu, nbin);
    if (nbin >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) { atomicAdd(&sh[nbin], __popc(peers)); }
    const unsigned valid = __ballot_sync(0xffffffffu, nbin >= 0);
    if ((threadIdx.x & 31) == 0 && valid) atomicAdd(&sh[300], __popc(valid));
    __syncthreads();
    for (int i = threadIdx.x; i < 301; i += 256) if (sh[i]) atomicAdd(&hist[f * 304 + i], sh[i]);
}

__global__ void k_akz_kcontrast(const unsigned* __restrict__ hmax, const int* __restrict__ hist, float* __restrict__ kc, int B) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= B) return;
    const int* hf = hist + f * 304;
    const int nthreshold = (int)((float)hf[300] * 0.7f);
    int k = 0, nelements = 0;
    for (k = 0; nelements < nthreshold && k < 300; ++k) nelements += hf[k];
    kc[f] = nelements < nthreshold ? 0.03f : __uint_as_float(hmax[f]) * ((float)k / 300.0f);
}

__global__ void __launch_bounds__(256) k_akz_half(const float* __restrict__ src, int sst, long long sist, float* __restrict__ dst,
                                                  int w, int h, int st, long long ist) {
    AKZ_PIX();
    if (x >= w || y >= h) return;
    const float* P = src + f * sist + (long long)(2 * y) * sst + 2 * x;
    dst[f * ist + (long long)y * st + x] = ((P[0] + P[1]) + (P[sst] + P[sst + 1])) * 0.25f;
}

__global__ void __launch_bounds__(256) k_akz_flow(const float* __restrict__ lsm, float* __restrict__ flow, int w, int h, int st,
                                                  long long ist, const float* __restrict__ kc, int octave) {
    AKZ_PIX();
    if (x >= w || y >= h) return;
    float k = kc[f];
    for (int o = 0; o < octave; ++o) k = k * 0.75f;
    const float inv_k = 1.0f / (k * k);
    float lx, ly;
    akz_scharr3(lsm + f * ist, st, w, h, x, y, &lx, &ly);
    flow[f * ist + (long long)y * st + x] = 1.0f / (1.0f + inv_k * (lx * lx + ly * ly));
}

__global__ void __launch_bounds__(256) k_akz_nld(const float* __restrict__ Ld_, const float* __restrict__ c_, float* __restrict__ out,
                                                 int w, int h, int st, long long ist, float tau) {
    AKZ_PIX();
    if (x >= w || y >= h) return;
    const long long p = f * ist + (long long)y * st + x;
    const float* Ld = Ld_ + p; const float* c = c_ + p;
    const float l0 = Ld[0], c0 = c[0];
    float xpos = 0.f, xneg = 0.f, ypos = 0.f, yneg = 0.f;
    if (x + 1 < w) xpos = (c0 + c[1]) * (Ld[1] - l0);
    if (x > 0) xneg = (c[-1] + c0) * (l0 - Ld[-1]);
    if (y + 1 < h) ypos = (c0 + c[st]) * (Ld[st] - l0);
    if (y > 0) yneg = (c[-st] + c0) * (l0 - Ld[-st]);
    out[p] = l0 + (0.5f * tau) * ((xpos - xneg) + (ypos - yneg));
}

// ------------------------------------------------------------------------------------------------------
// One whole FED cycle (NS explicit steps) of a level in ONE kernel: the Lt / conductivity footprint of a tile plus a halo of
// NS cells is staged in shared memory once and the steps run back to back on a region that shrinks by one cell per step
// (temporal blocking), so HBM sees one read of Lt and of the flow and one write instead of NS x (2 reads + 1 write).
// Per-cell arithmetic and the one-sided image borders are those of k_akz_nld.
// ------------------------------------------------------------------------------------------------------
#define FED_SW 128
#define FED_SH 48
struct AkzTaus { float t[16]; };
__global__ void __launch_bounds__(256) k_akz_fed(const float* __restrict__ Lin, const float* __restrict__ flow, float* __restrict__ Lout,
                                                 int w, int h, int st, long long ist, int NS, const __grid_constant__ AkzTaus taus) {
    extern __shared__ __align__(16) float fsm[];
    float* A = fsm; float* Bf = fsm + FED_SH * FED_SW; float* C = Bf + FED_SH * FED_SW;
    const int tid = threadIdx.x, lane = tid & 31, wrow = tid >> 5, f = blockIdx.z;
    const int ow = FED_SW - 2 * NS, oh = FED_SH - 2 * NS;
    const int gx0 = blockIdx.x * ow - NS, gy0 = blockIdx.y * oh - NS;
    const float* Lf = Lin + f * ist; const float* Cf = flow + f * ist;
    for (int i = tid; i < FED_SH * FED_SW; i += 256) {
        const int r = i / FED_SW, c = i - r * FED_SW;
        const int gy = gy0 + r, gx = gx0 + c;
        const bool in = gx >= 0 && gx < w && gy >= 0 && gy < h;
        A[i] = in ? Lf[(long long)gy * st + gx] : 0.f;
        C[i] = in ? Cf[(long long)gy * st + gx] : 0.f;
    }
    __syncthreads();
    float* src = A; float* dst = Bf;
    for (int j = 0; j < NS; ++j) {
        const float ht = 0.5f * taus.t[j];
        for (int r = j + 1 + wrow; r < FED_SH - 1 - j; r += 8) {
            const int gy = gy0 + r;
            if (gy < 0 || gy >= h) continue;
            for (int c = j + 1 + lane; c < FED_SW - 1 - j; c += 32) {
                const int gx = gx0 + c;
                if (gx < 0 || gx >= w) continue;
                const int p = r * FED_SW + c;
                const float l0 = src[p], c0 = C[p];
                float xpos = 0.f, xneg = 0.f, ypos = 0.f, yneg = 0.f;
                if (gx + 1 < w) xpos = (c0 + C[p + 1]) * (src[p + 1] - l0);
                if (gx > 0) xneg = (C[p - 1] + c0) * (l0 - src[p - 1]);
                if (gy + 1 < h) ypos = (c0 + C[p + FED_SW]) * (src[p + FED_SW] - l0);
                if (gy > 0) yneg = (C[p - FED_SW] + c0) * (l0 - src[p - FED_SW]);
                dst[p] = l0 + ht * ((xpos - xneg) + (ypos - yneg));
            }
        }
        __syncthreads();
        float* t = src; src = dst; dst = t;
    }
    float* Of = Lout + f * ist;
    for (int r = NS + wrow; r < FED_SH - NS; r += 8) {
        const int gy = gy0 + r;
        if (gy >= h) continue;
        for (int c = NS + lane; c < FED_SW - NS; c += 32) {
            const int gx = gx0 + c;
            if (gx < w) Of[(long long)gy * st + gx] = src[r * FED_SW + c];
        }
    }
}

// scaled Scharr value at (x, y): dir 0 = d/dx, 1 = d/dy (row pass then column pass of sepFilter2D, evaluated directly)
__device__ __forceinline__ float akz_dscaled(const float* __restrict__ s, int st, int w, int h, int x, int y, int dir, int sc, float norm, float wc) {
    const int xm = akz_refl(x - sc, w), xp = akz_refl(x + sc, w);
    const float* r0 = s + (long long)akz_refl(y - sc, h) * st; const float* r1 = s + (long long)y * st; const float* r2 = s + (long long)akz_refl(y + sc, h) * st;
    if (dir == 0) return norm * ((r0[xp] - r0[xm]) + (r2[xp] - r2[xm])) + wc * (r1[xp] - r1[xm]);
    return (norm * (r2[xm] + r2[xp]) + wc * r2[x]) - (norm * (r0[xm] + r0[xp]) + wc * r0[x]);
}

__global__ void __launch_bounds__(256) k_akz_deriv1(const float* __restrict__ lsm, float* __restrict__ lx0, float* __restrict__ ly0,
                                                    int w, int h, int st, long long ist, int sc, float norm, float wc) {
    AKZ_PIX();
    if (x >= w || y >= h) return;
    const long long p = f * ist + (long long)y * st + x;
    lx0[p] = akz_dscaled(lsm + f * ist, st, w, h, x, y, 0, sc, norm, wc);
    ly0[p] = akz_dscaled(lsm + f * ist, st, w, h, x, y, 1, sc, norm, wc);
}

__global__ void __launch_bounds__(256) k_akz_deriv2(const float* __restrict__ lx0, const float* __restrict__ ly0, float* __restrict__ Lx,
                                                    float* __restrict__ Ly, float* __restrict__ Ldet, int w, int h, int st, long long ist,
                                                    int sc, float norm, float wc) {
    AKZ_PIX();
    if (x >= w || y >= h) return;
    const long long p = f * ist + (long long)y * st + x;
    const float lxx = akz_dscaled(lx0 + f * ist, st, w, h, x, y, 0, sc, norm, wc);
    const float lyy = akz_dscaled(ly0 + f * ist, st, w, h, x, y, 1, sc, norm, wc);
    const float lxy = akz_dscaled(lx0 + f * ist, st, w, h, x, y, 1, sc, norm, wc);
    const float s1 = (float)sc, s2 = (float)(sc * sc);
    const float a = lxx * s2, b = lyy * s2, c = lxy * s2;
    Ldet[p] = a * b - c * c;
    Lx[p] = lx0[p] * s1; Ly[p] = ly0[p] * s1;
}

// ------------------------------------------------------------------------------------------------------
// Fused multiscale derivatives: one CTA = 64 x 32 outputs.  The smoothed level is staged once with a halo of 2*sc
// (BORDER_REFLECT_101 resolved while staging), the un-scaled first derivatives are built in shared memory on the tile plus a
// halo of sc -- for cells outside the image as the derivative AT THE REFLECTED POSITION, which is what the second sepFilter2D
// of the reference sees -- and the second derivatives / Hessian determinant come from those.  Same per-value arithmetic as
// k_akz_deriv1 + k_akz_deriv2 with ~3 global loads per pixel instead of ~32.
// ------------------------------------------------------------------------------------------------------
#define HS_TW 64
#define HS_TH 32
__host__ __device__ inline size_t akz_hess_smem(int sc) {
    return sizeof(float) * ((size_t)(HS_TW + 4 * sc) * (HS_TH + 4 * sc) + 2 * (size_t)(HS_TW + 2 * sc) * (HS_TH + 2 * sc));
}
__global__ void __launch_bounds__(256) k_akz_hessian(const float* __restrict__ lsm, float* __restrict__ Lx, float* __restrict__ Ly,
                                                     float* __restrict__ Ldet, int w, int h, int st, long long ist, int sc,
                                                     float norm, float wc) {
    extern __shared__ __align__(16) float hsm[];
    const int SW = HS_TW + 4 * sc, SH = HS_TH + 4 * sc;          // staged smoothed level
    const int DW = HS_TW + 2 * sc, DH = HS_TH + 2 * sc;          // first-derivative region
    float* S = hsm; float* DX = hsm + SW * SH; float* DY = DX + DW * DH;
    const int tid = threadIdx.x, f = blockIdx.z;
    const int tx0 = blockIdx.x * HS_TW, ty0 = blockIdx.y * HS_TH;
    const float* src = lsm + f * ist;
    for (int i = tid; i < SW * SH; i += 256) {
        const int r = i / SW, c = i - r * SW;
        S[i] = src[(long long)akz_refl(ty0 - 2 * sc + r, h) * st + akz_refl(tx0 - 2 * sc + c, w)];
    }
    __syncthreads();
    for (int i = tid; i < DW * DH; i += 256) {
        const int r = i / DW, c = i - r * DW;
        const int gy = ty0 - sc + r, gx = tx0 - sc + c;
        if (gy > h - 1 + sc || gx > w - 1 + sc) continue;          // never read by an in-image output
        const int sr = akz_refl(gy, h) - (ty0 - 2 * sc), sx = akz_refl(gx, w) - (tx0 - 2 * sc);
        const float* r0 = S + (sr - sc) * SW; const float* r1 = S + sr * SW; const float* r2 = S + (sr + sc) * SW;
        const int xm = sx - sc, xp = sx + sc;
        DX[i] = norm * ((r0[xp] - r0[xm]) + (r2[xp] - r2[xm])) + wc * (r1[xp] - r1[xm]);
        DY[i] = (norm * (r2[xm] + r2[xp]) + wc * r2[sx]) - (norm * (r0[xm] + r0[xp]) + wc * r0[sx]);
    }
    __syncthreads();
    const float s1 = (float)sc, s2 = (float)(sc * sc);
    for (int i = tid; i < HS_TW * HS_TH; i += 256) {
        const int r = i / HS_TW, c = i - r * HS_TW;
        const int gy = ty0 + r, gx = tx0 + c;
        if (gy >= h || gx >= w) continue;
        const int dr = r + sc, dc = c + sc;
        const float* x0 = DX + (dr - sc) * DW; const float* x1 = DX + dr * DW; const float* x2 = DX + (dr + sc) * DW;
        const float* y0 = DY + (dr - sc) * DW; const float* y2 = DY + (dr + sc) * DW;
        const int cm = dc - sc, cp = dc + sc;
        const float lxx = norm * ((x0[cp] - x0[cm]) + (x2[cp] - x2[cm])) + wc * (x1[cp] - x1[cm]);
        const float lyy = (norm * (y2[cm] + y2[cp]) + wc * y2[dc]) - (norm * (y0[cm] + y0[cp]) + wc * y0[dc]);
        const float lxy = (norm * (x2[cm] + x2[cp]) + wc * x2[dc]) - (norm * (x0[cm] + x0[cp]) + wc * x0[dc]);
        const float a = lxx * s2, b = lyy * s2, cc = lxy * s2;
        const long long p = f * ist + (long long)gy * st + gx;
        Ldet[p] = a * b - cc * cc;
        Lx[p] = x1[dc] * s1; Ly[p] = DY[dr * DW + dc] * s1;
    }
}

__global__ void __launch_bounds__(256) k_akz_extrema(const __grid_constant__ AkzParams P, int i) {
    const AkzLevelG& L = P.lv[i];
    AKZ_PIX();
    const int w = L.w, h = L.h, st = L.stride;
    if (x < 1 || x >= w - 1 || y < 1 || y >= h - 1) return;
    const float* D = L.Ldet + f * L.istride + (long long)y * st + x;
    const float v = D[0];
    if (!(v > P.dth && v >= 0.00001f && v > D[-1] && v > D[1] && v > D[-st - 1] && v > D[-st] && v > D[-st + 1] &&
          v > D[st - 1] && v > D[st] && v > D[st + 1])) return;
    const float smax = 10.0f * 0x1.6a09e6p+0f;        // 10 sqrt(2): libAKAZE's descriptor border for SURF / MLDB (12 sqrt 2 is M-SURF); cv2 agrees
    const float ratio = (float)(1 << L.octave);
    const int sigma_size = (int)(L.psize / ratio + 0.5f);
    const int left = akz_fround((float)x - smax * (float)sigma_size) - 1, right = akz_fround((float)x + smax * (float)sigma_size) + 1;
    const int up = akz_fround((float)y - smax * (float)sigma_size) - 1, down = akz_fround((float)y + smax * (float)sigma_size) + 1;
    if (left < 0 || right >= w || up < 0 || down >= h) return;
    const int slot = atomicAdd(&P.cnt[f * 16 + i], 1);
    if (slot >= L.cand_cap) { atomicOr(&P.status[f], AKZ_ST_CAND_OVERFLOW); return; }
    L.cand[(long long)f * L.cand_cap + slot] = make_uint2(((uint32_t)y << 16) | (uint32_t)x, __float_as_uint(v));
}

__global__ void __launch_bounds__(512) k_akz_sort(const __grid_constant__ AkzParams P) {
    extern __shared__ __align__(16) unsigned long long sk[];
    const int i = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const AkzLevelG& L = P.lv[i];
    const int n = min(P.cnt[f * 16 + i], L.cand_cap);
    if (n == 0) return;
    const uint2* cand = L.cand + (long long)f * L.cand_cap;
    uint2* srt = L.srt + (long long)f * L.cand_cap;
    int np = 1;
    while (np < n) np <<= 1;
    for (int k = tid; k < np; k += 512) sk[k] = k < n ? (((unsigned long long)cand[k].x << 32) | (unsigned)k) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= np; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < np; t += 512) {
                const int txj = t ^ j;
                if (txj > t) {
                    const unsigned long long a = sk[t], b = sk[txj];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) { sk[t] = b; sk[txj] = a; }
                }
            }
            __syncthreads();
        }
    for (int k = tid; k < n; k += 512) srt[k] = cand[(unsigned)(sk[k] & 0xffffffffu)];
}

// ------------------------------------------------------------------------------------------------------
// libAKAZE Find_Scale_Space_Extrema (sequential over raster-ordered candidates, level by level) + Do_Subpixel_Refinement,
// one CTA per frame.  The list lives in shared memory; every candidate is checked against the FIRST close entry (index
// order) of its own or the lower level: the scan is spread over the CTA, the decision is replayed identically by all
// threads.  Then the upper-level filter (parallel: an "exists" query), refinement and an ordered compaction that also
// partitions the survivors per evolution level (= the reference's octave, src/Feature_akaze61.cpp:63-65).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_akz_select(const __grid_constant__ AkzParams P) {
    extern __shared__ __align__(16) unsigned char sraw[];
    const int LC = P.list_cap;
    float* ax = reinterpret_cast<float*>(sraw);
    float* ay = ax + LC;
    float* ar = ay + LC;
    unsigned char* ac = reinterpret_cast<unsigned char*>(ar + LC);
    unsigned char* aflag = ac + LC;
    __shared__ int s_first[2];
    __shared__ int lvcnt[16], lvoff[16], lvrun[16], wcnt[8][16];
    __shared__ int s_total;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    if (tid < 2) s_first[tid] = 0x7fffffff;
    if (tid < 16) { lvcnt[tid] = 0; lvrun[tid] = 0; }
    if (tid == 0) s_total = 0;
    __syncthreads();
    int n = 0, it = 0;
    bool overflow = false;
    for (int i = 0; i < P.nl; ++i) {
        const AkzLevelG& L = P.lv[i];
        const int m = min(P.cnt[f * 16 + i], L.cand_cap);
        const uint2* cand = L.srt + (long long)f * L.cand_cap;
        const float ratio = (float)(1 << L.octave), ps2 = L.psize * L.psize;
        for (int c = 0; c < m; ++c, ++it) {
            const uint2 cd = cand[c];
            const float px = (float)(cd.x & 0xffff) * ratio, py = (float)(cd.x >> 16) * ratio, v = __uint_as_float(cd.y);
            int first = 0x7fffffff;
            for (int k = tid; k < n; k += 256) {
                const int cl = ac[k];
                if (cl == i - 1 || cl == i) {
                    const float dx = px - ax[k], dy = py - ay[k];
                    if (dx * dx + dy * dy <= ps2) { first = k; break; }
                }
            }
            if (first != 0x7fffffff) atomicMin(&s_first[it & 1], first);
            __syncthreads();
            first = s_first[it & 1];
            bool replace = false, append = false;
            if (first != 0x7fffffff) replace = v > ar[first]; else append = true;
            __syncthreads();                               // every thread has read ar[first] / s_first before they change
            if (tid == 0) {
                s_first[it & 1] = 0x7fffffff;
                if (replace) { ax[first] = px; ay[first] = py; ar[first] = v; ac[first] = (unsigned char)i; }
                else if (append && n < LC) { ax[n] = px; ay[n] = py; ar[n] = v; ac[n] = (unsigned char)i; }
            }
            if (append) { if (n < LC) ++n; else overflow = true; }
            // the next iteration uses the other s_first slot; the barrier after its scan orders these writes
            __syncthreads();
        }
    }
    if (overflow && tid == 0) atomicOr(&P.status[f], AKZ_ST_CAND_OVERFLOW);
    // ---- upper-level filter + sub-pixel refinement (results overwrite ax / ay; aflag = survives)
    float rx_[AKZ_LIST_CAP / 256], ry_[AKZ_LIST_CAP / 256];
    for (int i = tid, q = 0; i < n; i += 256, ++q) {
        const int cl = ac[i];
        const float xi = ax[i], yi = ay[i], ri = ar[i];
        const float sz = P.lv[cl].psize, sz2 = sz * sz;
        bool rep = false;
        for (int j = i + 1; j < n && !rep; ++j)
            if (ac[j] == cl + 1) {
                const float dx = xi - ax[j], dy = yi - ay[j];
                if (dx * dx + dy * dy <= sz2 && ri < ar[j]) rep = true;
            }
        bool ok = !rep;
        float nx = 0.f, ny = 0.f;
        if (ok) {
            const AkzLevelG& L = P.lv[cl];
            const float ratio = (float)(1 << L.octave);
            const int x = akz_fround(xi / ratio), y = akz_fround(yi / ratio), st = L.stride;
            const float* D = L.Ldet + f * L.istride + (long long)y * st + x;
            const float Dx = 0.5f * (D[1] - D[-1]), Dy = 0.5f * (D[st] - D[-st]);
            const float Dxx = (D[1] + D[-1]) - 2.0f * D[0], Dyy = (D[st] + D[-st]) - 2.0f * D[0];
            const float Dxy = 0.25f * (D[st + 1] + D[-st - 1]) - 0.25f * (D[-st + 1] + D[st - 1]);
            const float det = Dxx * Dyy - Dxy * Dxy;
            if (det == 0.f) ok = false;
            else {
                const float ox = (Dxy * Dy - Dyy * Dx) / det, oy = (Dxy * Dx - Dxx * Dy) / det;
                if (!(fabsf(ox) <= 1.0f && fabsf(oy) <= 1.0f)) ok = false;
                else {
                    nx = ((float)x + ox) * ratio + 0.5f * (ratio - 1.0f);
                    ny = ((float)y + oy) * ratio + 0.5f * (ratio - 1.0f);
                }
            }
        }
        aflag[i] = ok ? 1 : 0;
        rx_[q] = nx; ry_[q] = ny;
        if (ok) atomicAdd(&lvcnt[cl], 1);
    }
    __syncthreads();                                      // all upper-level scans are done: positions may be overwritten
    for (int i = tid, q = 0; i < n; i += 256, ++q) if (aflag[i]) { ax[i] = rx_[q]; ay[i] = ry_[q]; }
    if (tid == 0) {
        int run = 0;
        for (int k = 0; k < 16; ++k) { lvoff[k] = run; run += lvcnt[k]; }
        if (run > P.key_cap) atomicOr(&P.status[f], AKZ_ST_CAND_OVERFLOW);
    }
    __syncthreads();
    // ---- ordered compaction: list position (kpt / kcls) and per-level segments for the octree
    float4* kpt = P.kpt + (long long)f * P.key_cap; int* kcls = P.kcls + (long long)f * P.key_cap;
    float* okx = P.okx + (long long)f * P.key_cap; float* oky = P.oky + (long long)f * P.key_cap;
    uint32_t* oresp = P.oresp + (long long)f * P.key_cap; int* oidx = P.oidx + (long long)f * P.key_cap;
    for (int base = 0; base < n; base += 256) {
        const int i = base + tid;
        const bool valid = i < n && aflag[i];
        const int cl = valid ? ac[i] : 0;
        int my_rank = 0;
        const unsigned vb = __ballot_sync(0xffffffffu, valid);
        for (int k = 0; k < P.nl; ++k) {
            const unsigned mk = __ballot_sync(0xffffffffu, valid && cl == k);
            if (lane == 0) wcnt[wp][k] = __popc(mk);
            if (valid && cl == k) my_rank = __popc(mk & ((1u << lane) - 1));
        }
        if (lane == 0) wcnt[wp][15] = __popc(vb);
        const int my_list_rank = __popc(vb & ((1u << lane) - 1));
        __syncthreads();
        if (valid) {
            int before = 0, lbefore = 0;
            for (int q = 0; q < wp; ++q) { before += wcnt[q][cl]; lbefore += wcnt[q][15]; }
            const int lpos = s_total + lbefore + my_list_rank;
            const int pos = lvoff[cl] + lvrun[cl] + before + my_rank;
            if (pos < P.key_cap && lpos < P.key_cap) {
                kpt[lpos] = make_float4(ax[i], ay[i], P.lv[cl].psize * 2.0f, ar[i]);
                kcls[lpos] = cl;
                okx[pos] = ax[i]; oky[pos] = ay[i];
                const uint32_t b = __float_as_uint(ar[i]);
                oresp[pos] = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
                oidx[pos] = lpos;
            }
        }
        __syncthreads();
        if (tid < 16) {
            int t = 0;
            for (int q = 0; q < 8; ++q) t += wcnt[q][tid];
            if (tid == 15) s_total += t; else if (tid < P.nl) lvrun[tid] += t;
        }
        __syncthreads();
    }
    if (tid < 16) { P.selinfo[f * 32 + tid] = lvcnt[tid]; P.selinfo[f * 32 + 16 + tid] = lvoff[tid]; }
}

// tap: Feature_Detection list {x, y, size, response, class_id}
__global__ void k_akz_tap_list(const __grid_constant__ AkzParams P, int f, float* out5, int cap, int* n_total) {
    int total = 0;
    for (int k = 0; k < 16; ++k) total += P.selinfo[f * 32 + k];
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_total = total;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total && g < cap && g < P.key_cap; g += gridDim.x * blockDim.x) {
        const float4 k = P.kpt[(long long)f * P.key_cap + g];
        out5[5 * g] = k.x; out5[5 * g + 1] = k.y; out5[5 * g + 2] = k.z; out5[5 * g + 3] = k.w; out5[5 * g + 4] = (float)P.kcls[(long long)f * P.key_cap + g];
    }
}

__global__ void __launch_bounds__(256) k_akz_octree(const __grid_constant__ AkzParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OctWork W;
    oct_carve(smem_raw, P.oct_ncap, W);
    const int lv = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const int M = min(P.selinfo[f * 32 + lv], P.key_cap), off = P.selinfo[f * 32 + 16 + lv];
    if (M == 0 || off + M > P.key_cap) { if (tid == 0) P.keepcnt[f * 16 + lv] = 0; return; }
    const long long base = (long long)f * P.key_cap + off;
    bool overflow = false;
    int size = oct_distribute(W, P.okx + base, P.oky + base, P.knode + base, P.kquad + base, M, P.q_ext[lv], P.n_ini, P.hX,
                              P.H, P.oct_ncap, tid, overflow);
    if (overflow && tid == 0) atomicOr(&P.status[f], AKZ_ST_OCTREE_OVERFLOW);
    // node representative: max response, the first in list order among equals (src/ORBextractor.cc:444-455)
    unsigned long long* best = W.best;
    for (int p = tid; p < size; p += 256) best[p] = 0ull;
    __syncthreads();
    const unsigned short* knode = P.knode + base; const uint32_t* oresp = P.oresp + base;
    for (int k = tid; k < M; k += 256) atomicMax(&best[knode[k]], ((unsigned long long)oresp[k] << 32) | (unsigned long long)(0xffffffffu - (unsigned)k));
    __syncthreads();
    if (size > P.keep_cap) { if (tid == 0) atomicOr(&P.status[f], AKZ_ST_OCTREE_OVERFLOW); size = P.keep_cap; }
    int* keep = P.keep + ((long long)f * P.nlevels + lv) * P.keep_cap;
    for (int p = tid; p < size; p += 256) keep[p] = (int)(0xffffffffu - (unsigned)(best[p] & 0xffffffffu));
    if (tid == 0) P.keepcnt[f * 16 + lv] = size;
}

// ------------------------------------------------------------------------------------------------------
// Compute_Main_Orientation + Get_MLDB_Full_Descriptor, warp per kept keypoint; merged output.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_akz_describe(const __grid_constant__ AkzParams P, afv_keypoint* __restrict__ kps,
                                                      uint8_t* __restrict__ desc, float* __restrict__ kpsize, int* __restrict__ n_out) {
    __shared__ long long s_rx[8][112], s_ry[8][112];
    __shared__ float s_ang[8][112];
    __shared__ float s_val[8][29][3];
    const int f = blockIdx.y, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int j = blockIdx.x * 8 + wp;
    int total = 0, lv = -1, p = 0;
    for (int k = 0; k < P.nlevels; ++k) {
        const int c = P.keepcnt[f * 16 + k];
        if (lv < 0 && j < total + c) { lv = k; p = j - total; }
        total += c;
    }
    if (j == 0 && lane == 0) {
        if (total > P.out_cap) atomicOr(&P.status[f], AKZ_ST_OUT_OVERFLOW);
        n_out[f] = min(total, P.out_cap);
    }
    if (lv < 0 || j >= P.out_cap) return;
    const int key = P.keep[((long long)f * P.nlevels + lv) * P.keep_cap + p];
    const int li = P.oidx[(long long)f * P.key_cap + P.selinfo[f * 32 + 16 + lv] + key];
    const float4 kp = P.kpt[(long long)f * P.key_cap + li];
    const AkzLevelG& L = P.lv[lv];
    const int w = L.w, h = L.h, st = L.stride;
    const float* Lt = L.Lt + f * L.istride; const float* Lx = L.Lx + f * L.istride; const float* Ly = L.Ly + f * L.istride;
    const float ratio = (float)(1 << L.octave);
    const int s = akz_fround(0.5f * kp.z / ratio);
    const float xf = kp.x / ratio, yf = kp.y / ratio;
    // ---- main orientation
    for (int idx = lane; idx < 109; idx += 32) {
        const int i = c_ori_ij[idx][0], jj = c_ori_ij[idx][1];
        const int iy = min(max(akz_fround(yf + (float)(jj * s)), 0), h - 1), ix = min(max(akz_fround(xf + (float)(i * s)), 0), w - 1);
        const float g = akz_gauss25(abs(i), abs(jj));
        const float axv = g * Lx[(long long)iy * st + ix], ayv = g * Ly[(long long)iy * st + ix];
        s_rx[wp][idx] = (long long)rintf(axv * 4294967296.0f); s_ry[wp][idx] = (long long)rintf(ayv * 4294967296.0f);
        s_ang[wp][idx] = akz_atan2(ayv, axv);
    }
    __syncwarp();
    float bm = 0.f; int bt = 1 << 20; long long bsx = 0, bsy = 0;
    for (int t = lane; t < 42; t += 32) {
        const float a1 = (float)t * 0.15f;
        const float a2 = (a1 + AKZ_PI / 3.0f > AKZ_2PI) ? a1 - 5.0f * AKZ_PI / 3.0f : a1 + AKZ_PI / 3.0f;
        long long sx = 0, sy = 0;
        for (int k = 0; k < 109; ++k) {
            const float a = s_ang[wp][k];
            if ((a1 < a2 && a1 < a && a < a2) || (a2 < a1 && ((a > 0.f && a < a2) || (a > a1 && a < AKZ_2PI)))) { sx += s_rx[wp][k]; sy += s_ry[wp][k]; }
        }
        const float fx = (float)sx, fy = (float)sy;
        const float m = fx * fx + fy * fy;
        if (m > bm) { bm = m; bt = t; bsx = sx; bsy = sy; }          // t ascending per lane: first maximum wins
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, bm, off);
        const int ot = __shfl_xor_sync(0xffffffffu, bt, off);
        const long long osx = __shfl_xor_sync(0xffffffffu, bsx, off), osy = __shfl_xor_sync(0xffffffffu, bsy, off);
        if (om > bm || (om == bm && ot < bt)) { bm = om; bt = ot; bsx = osx; bsy = osy; }
    }
    const float angle = bm > 0.f ? akz_atan2((float)bsy, (float)bsx) : 0.f;
    // ---- MLDB: 29 cells (2x2, 3x3, 4x4), lane = cell, samples summed in libAKAZE's loop order
    float si, co;
    akz_sincos(angle, &si, &co);
    const float cs = co * (float)s, ss = si * (float)s;
    if (lane < 29) {
        int g, ci;
        if (lane < 4) { g = 0; ci = lane; } else if (lane < 13) { g = 1; ci = lane - 4; } else { g = 2; ci = lane - 13; }
        const int step = g == 0 ? 10 : g == 1 ? 7 : 5, gn = g + 2;
        const int i0 = -10 + (ci / gn) * step, j0 = -10 + (ci % gn) * step;
        float di = 0.f, dx = 0.f, dy = 0.f;
        int ns = 0;
        for (int k = i0; k < i0 + step; ++k)
            for (int l = j0; l < j0 + step; ++l) {
                const float sy = yf + ((float)l * cs + (float)k * ss);
                const float sx = xf + ((float)k * cs - (float)l * ss);
                const int y1 = min(max(akz_fround(sy), 0), h - 1), x1 = min(max(akz_fround(sx), 0), w - 1);
                const long long q = (long long)y1 * st + x1;
                const float rx = Lx[q], ry = Ly[q];
                di = di + Lt[q];
                dx = dx + (ry * co - rx * si);
                dy = dy + (rx * co + ry * si);
                ++ns;
            }
        s_val[wp][lane][0] = di / (float)ns; s_val[wp][lane][1] = dx / (float)ns; s_val[wp][lane][2] = dy / (float)ns;
    }
    __syncwarp();
    uint8_t* drow = desc + ((long long)f * P.out_cap + j) * 61;
    for (int byte = lane; byte < 61; byte += 32) {
        unsigned v = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int bit = byte * 8 + b;
            if (bit < 486) {
                const uchar4 e = c_mldb_bits[bit];
                if (s_val[wp][e.x][e.z] > s_val[wp][e.y][e.z]) v |= 1u << b;
            }
        }
        drow[byte] = (uint8_t)v;
    }
    if (lane == 0) {
        afv_keypoint o;
        o.x = kp.x; o.y = kp.y; o.size = kp.z; o.angle = angle; o.response = kp.w; o.octave = L.octave; o.class_id = lv;
        kps[(long long)f * P.out_cap + j] = o;
        if (kpsize) kpsize[(long long)f * P.out_cap + j] = P.size_norm[lv];
    }
}

// =====================================================================================================
// host side
// =====================================================================================================
struct AfvAkaze {
    int nfeatures, nlevels, max_batch, max_w, max_h, omax, nsub;
    float scale_factor, detect_th;
    std::vector<void*> allocs;
    AkzParams P;
    AfvBlurTaps taps16, taps10;
    float* lvbuf[AKZ_MAX_LV][5];
    uint2* cand[AKZ_MAX_LV]; uint2* srt[AKZ_MAX_LV]; int cand_cap[AKZ_MAX_LV];
    float* scr[4];                      // full-resolution scratch images [B]
    uint8_t* gray_stage; int* h_status;
    int use_tma;                        // AFV_BLUR_NO_TMA=1: clamped-load staging on every blur tile (A/B check)
    int unfused_fed;                    // AFV_AKAZE_UNFUSED_FED=1: one k_akz_nld launch per FED step and the two-kernel derivative
                                        // path (A/B check of k_akz_fed / k_akz_hessian)
    size_t img_floats;                  // floats per full-resolution frame image (max geometry)
};

template <typename T>
static int akz_alloc(AfvAkaze* s, T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 256);
    if (e != cudaSuccess) { afv_set_error("cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e)); return AFV_ERR_CUDA; }
    s->allocs.push_back(q);
    *p = (T*)q;
    return AFV_OK;
}

static int akz_gauss_taps(float sigma, float* taps) {
    int ks = (int)ceilf(2.0f * (1.0f + (sigma - 0.8f) / 0.3f));
    if ((ks % 2) == 0) ks += 1;
    const int r = ks / 2;
    double wd[32], sum = 0.0;
    for (int j = 0; j <= r; ++j) { wd[j] = exp(-(double)(j * j) / (2.0 * (double)sigma * (double)sigma)); sum += j ? 2.0 * wd[j] : wd[j]; }
    for (int j = 0; j <= r; ++j) taps[j] = (float)(wd[j] / sum);
    return r;
}

static int akz_is_prime(int n) { if (n < 2) return 0; for (int i = 2; i * i <= n; ++i) if (n % i == 0) return 0; return 1; }
// FED cycle (libAKAZE fed.cpp: fed_tau_by_process_time(T, 1, 0.25, reordering))
static int akz_fed_tau(float T, float tau_max, float* tau, int cap) {
    const int n = (int)(ceilf(sqrtf(3.0f * T / tau_max + 0.25f) - 0.5f - 1.0e-8f) + 0.5f);
    if (n <= 0) return 0;
    if (n > cap) return -1;
    const float scale = 3.0f * T / (tau_max * (float)(n * (n + 1)));
    const float c = 1.0f / (4.0f * (float)n + 2.0f), d = scale * tau_max / 2.0f;
    float tauh[64];
    for (int k = 0; k < n; ++k) { const float hc = (float)cos((double)(AKZ_PI * (2.0f * (float)k + 1.0f) * c)); tauh[k] = d / (hc * hc); }
    const int kappa = n / 2;
    int prime = n + 1;
    while (!akz_is_prime(prime)) ++prime;
    for (int k = 0, l = 0; l < n; ++k, ++l) {
        int index;
        while ((index = ((k + 1) * kappa) % prime - 1) >= n) ++k;
        tau[l] = tauh[index];
    }
    return n;
}

void afv_akaze_destroy(AfvAkaze* s) {
    if (!s) return;
    for (void* p : s->allocs) cudaFree(p);
    if (s->h_status) cudaFreeHost(s->h_status);
    delete s;
}

static int pow2ceil_i(int v) { int p = 1; while (p < v) p <<= 1; return p; }

int afv_akaze_create(AfvAkaze** out, int nfeatures, int nlevels, float scale_factor, float detect_th, int max_batch, int max_w, int max_h) {
    *out = nullptr;
    const int omax = nlevels / 4, nsub = nlevels / 2;                 // src/Feature_akaze61.cpp:10-11
    if (omax < 1 || omax * nsub > AKZ_MAX_LV) { afv_set_error("akaze61: numOctaves %d gives %d x %d evolution levels (1..%d supported)", nlevels, omax, nsub, AKZ_MAX_LV); return AFV_ERR_INVALID; }
    if (max_w > 4095 || max_h > 4095) { afv_set_error("akaze61: frame dimension > 4095 not supported"); return AFV_ERR_INVALID; }
    AfvAkaze* s = new AfvAkaze();
    memset(&s->P, 0, sizeof(s->P));
    s->nfeatures = nfeatures; s->nlevels = nlevels; s->scale_factor = scale_factor; s->detect_th = detect_th;
    s->max_batch = max_batch; s->max_w = max_w; s->max_h = max_h; s->omax = omax; s->nsub = nsub;
    s->gray_stage = nullptr; s->h_status = nullptr;
    { const char* e = getenv("AFV_AKAZE_UNFUSED_FED"); s->unfused_fed = e && e[0] == '1'; }
    { const char* e = getenv("AFV_BLUR_NO_TMA"); s->use_tma = !(e && e[0] == '1'); }
    memset(&s->taps16, 0, sizeof(AfvBlurTaps)); memset(&s->taps10, 0, sizeof(AfvBlurTaps));
    if (akz_gauss_taps(1.6f, s->taps16.t) != 4 || akz_gauss_taps(1.0f, s->taps10.t) != 2) { afv_set_error("akaze61: internal: unexpected blur radius"); delete s; return AFV_ERR_INVALID; }
    {
        cudaError_t e = afv_blur_cfg<4, 2>();
        if (e == cudaSuccess) e = afv_blur_cfg<2, 2>();
        if (e == cudaSuccess) e = afv_blur_cfg<2, 0>();
        if (e != cudaSuccess) { afv_set_error("cudaFuncSetAttribute(blur) failed: %s", cudaGetErrorString(e)); delete s; return AFV_ERR_CUDA; }
    }
    // constant tables
    {
        uchar4 bits[488]; memset(bits, 0, sizeof(bits));
        int dpos = 0, cell0 = 0;
        for (int g = 0; g < 3; ++g) {
            const int cnt = (g + 2) * (g + 2);
            for (int ch = 0; ch < 3; ++ch)
                for (int a = 0; a < cnt; ++a)
                    for (int b = a + 1; b < cnt; ++b) { bits[dpos].x = (unsigned char)(cell0 + a); bits[dpos].y = (unsigned char)(cell0 + b); bits[dpos].z = (unsigned char)ch; ++dpos; }
            cell0 += cnt;
        }
        signed char ij[109][2]; int idx = 0;
        for (int i = -6; i <= 6; ++i) for (int j = -6; j <= 6; ++j) if (i * i + j * j < 36) { ij[idx][0] = (signed char)i; ij[idx][1] = (signed char)j; ++idx; }
        cudaError_t e = dpos == 486 && idx == 109 ? cudaSuccess : cudaErrorUnknown;
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_mldb_bits, bits, sizeof(bits));
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_ori_ij, ij, sizeof(ij));
        if (e != cudaSuccess) { afv_set_error("akaze61: constant table upload failed: %s", cudaGetErrorString(e)); delete s; return AFV_ERR_CUDA; }
    }
    int rc = AFV_OK;
    const size_t B = (size_t)max_batch;
    s->img_floats = (size_t)((max_w + 31) & ~31) * max_h;
    int key_cap = 0;
    for (int i = 0; i < omax * nsub && rc == AFV_OK; ++i) {
        const int o = i / nsub;
        const int lw = max_w >> o, lh = max_h >> o;
        const size_t img = (size_t)((lw + 31) & ~31) * lh;
        for (int k = 0; k < 5 && rc == AFV_OK; ++k) {
            if (i == 0 && k == 1) { s->lvbuf[0][1] = s->lvbuf[0][0]; continue; }          // Lsmooth[0] = Lt[0]
            rc = akz_alloc(s, &s->lvbuf[i][k], img * B);
        }
        int cap = pow2ceil_i(lw * lh / 64);
        if (cap < 256) cap = 256;
        if (cap > 8192) cap = 8192;
        s->cand_cap[i] = cap; key_cap += cap;
        if (rc == AFV_OK) rc = akz_alloc(s, &s->cand[i], (size_t)cap * B);
        if (rc == AFV_OK) rc = akz_alloc(s, &s->srt[i], (size_t)cap * B);
    }
    if (key_cap > AKZ_LIST_CAP) key_cap = AKZ_LIST_CAP;
    AkzParams& P = s->P;
    P.key_cap = key_cap;
    for (int k = 0; k < 4 && rc == AFV_OK; ++k) rc = akz_alloc(s, &s->scr[k], s->img_floats * B);
    int maxq = 0;
    if (rc == AFV_OK) {
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
        if (oct_work_bytes(P.oct_ncap) > 227 * 1024) { afv_set_error("akaze61: nfeatures too large for the octree workspace"); rc = AFV_ERR_INVALID; }
    }
    if (rc == AFV_OK) rc = akz_alloc(s, &P.hmax, B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.hist, 304 * B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.kcontrast, B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.cnt, 16 * B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.status, B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.kpt, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.kcls, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.okx, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.oky, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.oresp, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.oidx, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.knode, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.kquad, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.selinfo, 32 * B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.keep, (size_t)P.keep_cap * nlevels * B);
    if (rc == AFV_OK) rc = akz_alloc(s, &P.keepcnt, 16 * B);
    if (rc == AFV_OK) rc = akz_alloc(s, &s->gray_stage, (size_t)max_w * max_h * B);
    if (rc == AFV_OK) {
        cudaError_t e = cudaMallocHost((void**)&s->h_status, sizeof(int) * B);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_akz_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_akz_hessian, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)akz_hess_smem(8));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_akz_fed, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * FED_SH * FED_SW * (int)sizeof(float));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_akz_select, cudaFuncAttributeMaxDynamicSharedMemorySize, 14 * AKZ_LIST_CAP);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_akz_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oct_work_bytes(P.oct_ncap));
        if (e != cudaSuccess) { afv_set_error("akaze61: setup failed: %s", cudaGetErrorString(e)); rc = AFV_ERR_CUDA; }
    }
    if (rc != AFV_OK) { afv_akaze_destroy(s); return rc; }
    *out = s;
    return AFV_OK;
}

uint8_t* afv_akaze_stage(AfvAkaze* s) { return s->gray_stage; }

static int akz_configure(AfvAkaze* s, int w, int h, int B) {
    if (w > s->max_w || h > s->max_h || w < 64 || h < 64) {
        afv_set_error("frame %dx%d outside the extractor's configured range (64..%d x 64..%d)", w, h, s->max_w, s->max_h);
        return AFV_ERR_INVALID;
    }
    if ((w & 1) || (h & 1)) { afv_set_error("akaze61: odd frame sizes are not supported (exact 2x2 area half-sampling only)"); return AFV_ERR_INVALID; }
    AkzParams& P = s->P;
    P.B = B; P.W = w; P.H = h; P.nfeatures = s->nfeatures; P.nlevels = s->nlevels; P.dth = s->detect_th;
    P.n_ini = (int)round((double)((float)w / (float)h));
    if (P.n_ini < 1) { afv_set_error("portrait frames with w/h < 0.5 are not supported (reference divides by zero)"); return AFV_ERR_INVALID; }
    P.hX = (float)w / (float)P.n_ini;
    int n = 0;
    for (int i = 0; i < s->omax; ++i) {
        const float rf = 1.0f / (float)(1 << i);
        const int lw = (int)((float)w * rf), lh = (int)((float)h * rf);
        if ((lw < 80 || lh < 40) && i != 0) break;
        for (int j = 0; j < s->nsub; ++j, ++n) {
            AkzLevelG& L = P.lv[n];
            L.w = lw; L.h = lh; L.stride = (lw + 31) & ~31; L.istride = (long long)L.stride * lh;
            L.octave = i;
            L.esigma = 1.6f * (float)pow(2.0, (double)((float)j / (float)s->nsub + (float)i));
            L.sigma_size = (int)(L.esigma * 1.5f / (float)(1 << i) + 0.5f);
            L.psize = L.esigma * 1.5f;
            L.nsteps = 0;
            L.Lt = s->lvbuf[n][0]; L.Lsm = s->lvbuf[n][1]; L.Lx = s->lvbuf[n][2]; L.Ly = s->lvbuf[n][3]; L.Ldet = s->lvbuf[n][4];
            L.cand = s->cand[n]; L.srt = s->srt[n]; L.cand_cap = s->cand_cap[n];
        }
    }
    P.nl = n;
    // list capacity: one entry per ~64 pixels of the input is an order of magnitude above what textured frames produce
    // (640x480: ~2000 entries); smaller lists let several k_akz_select CTAs share an SM
    { int lc = 1024; while (lc < (w * h) / 96 && lc < AKZ_LIST_CAP) lc <<= 1; P.list_cap = lc < P.key_cap ? lc : P.key_cap; }
    for (int i = 1; i < n; ++i) {
        const float e1 = 0.5f * (P.lv[i].esigma * P.lv[i].esigma), e0 = 0.5f * (P.lv[i - 1].esigma * P.lv[i - 1].esigma);
        float tau[64];
        const int ns = akz_fed_tau(e1 - e0, 0.25f, tau, 16);
        if (ns < 0) { afv_set_error("akaze61: FED cycle longer than 16 steps"); return AFV_ERR_INVALID; }
        P.lv[i].nsteps = ns;
        for (int k = 0; k < ns; ++k) P.lv[i].tau[k] = tau[k];
    }
    return AFV_OK;
}

#define AKZ_GRID(L, B) dim3(((L).w + 63) / 64, ((L).h + 3) / 4, (B))

int afv_akaze_run(AfvAkaze* s, const uint8_t* d_gray, int B, int w, int h, int stride, long frame_stride, afv_keypoint* d_kps,
                  uint8_t* d_desc, float* d_kpsize, int cap, int* d_n_out, cudaStream_t st) {
    if (B < 1 || B > s->max_batch) { afv_set_error("batch %d outside 1..%d", B, s->max_batch); return AFV_ERR_INVALID; }
    int rc = akz_configure(s, w, h, B);
    if (rc) return rc;
    AkzParams P = s->P;
    P.out_cap = cap;
    AFV_CUDA_CHECK(cudaMemsetAsync(P.cnt, 0, sizeof(int) * 16 * B, st));
    AFV_CUDA_CHECK(cudaMemsetAsync(P.status, 0, sizeof(int) * B, st));
    AFV_CUDA_CHECK(cudaMemsetAsync(P.hmax, 0, sizeof(unsigned) * B, st));
    AFV_CUDA_CHECK(cudaMemsetAsync(P.hist, 0, sizeof(int) * 304 * B, st));
    const AkzLevelG& L0 = P.lv[0];
    { AfvProfScope ps("k_akz_base", st);
      afv_blur_launch<4, 2>(d_gray, stride, frame_stride, L0.Lt, nullptr, L0.w, L0.h, L0.stride, L0.istride, s->taps16, B, st); ++g_afv_launches;
      afv_blur_launch<2, 2>(d_gray, stride, frame_stride, s->scr[0], nullptr, L0.w, L0.h, L0.stride, L0.istride, s->taps10, B, st); ++g_afv_launches;
      k_akz_mag<<<AKZ_GRID(L0, B), 256, 0, st>>>(s->scr[0], s->scr[1], L0.w, L0.h, L0.stride, L0.istride, P.hmax); ++g_afv_launches;
      k_akz_hist<<<AKZ_GRID(L0, B), 256, 0, st>>>(s->scr[1], L0.w, L0.h, L0.stride, L0.istride, P.hmax, P.hist); ++g_afv_launches;
      k_akz_kcontrast<<<(B + 127) / 128, 128, 0, st>>>(P.hmax, P.hist, P.kcontrast, B); ++g_afv_launches; }
    { AfvProfScope ps("k_akz_diffusion", st);
    for (int i = 1; i < P.nl; ++i) {
        const AkzLevelG& L = P.lv[i];
        const AkzLevelG& Q = P.lv[i - 1];
        const float* cur = Q.Lt;
        if (L.octave > Q.octave) {
            k_akz_half<<<AKZ_GRID(L, B), 256, 0, st>>>(Q.Lt, Q.stride, Q.istride, s->scr[0], L.w, L.h, L.stride, L.istride); ++g_afv_launches;
            cur = s->scr[0];
        }
        { CUtensorMap tm; const bool ok = s->use_tma && afv_blur_tmap(&tm, cur, L.w, L.h, L.stride, L.istride, B, 2);
          afv_blur_launch<2, 0>(cur, L.stride, L.istride, L.Lsm, nullptr, L.w, L.h, L.stride, L.istride, s->taps10, B, st, ok ? &tm : nullptr, 0); ++g_afv_launches; }
        k_akz_flow<<<AKZ_GRID(L, B), 256, 0, st>>>(L.Lsm, s->scr[2], L.w, L.h, L.stride, L.istride, P.kcontrast, L.octave); ++g_afv_launches;
        if (L.nsteps >= 1 && L.nsteps <= 12 && !s->unfused_fed) {
            AkzTaus taus; memset(&taus, 0, sizeof(taus));
            for (int j = 0; j < L.nsteps; ++j) taus.t[j] = L.tau[j];
            const int ow = FED_SW - 2 * L.nsteps, oh = FED_SH - 2 * L.nsteps;
            k_akz_fed<<<dim3((L.w + ow - 1) / ow, (L.h + oh - 1) / oh, B), 256, 3 * FED_SH * FED_SW * sizeof(float), st>>>(
                cur, s->scr[2], L.Lt, L.w, L.h, L.stride, L.istride, L.nsteps, taus); ++g_afv_launches;
        } else
        for (int j = 0; j < L.nsteps; ++j) {
            float* dst = j == L.nsteps - 1 ? L.Lt : (cur == s->scr[0] ? s->scr[1] : s->scr[0]);
            k_akz_nld<<<AKZ_GRID(L, B), 256, 0, st>>>(cur, s->scr[2], dst, L.w, L.h, L.stride, L.istride, L.tau[j]); ++g_afv_launches;
            cur = dst;
        }
    } }
    { AfvProfScope ps("k_akz_hessian", st);
    for (int i = 0; i < P.nl; ++i) {
        const AkzLevelG& L = P.lv[i];
        const int sc = L.sigma_size;
        const float wgt = 10.0f / 3.0f;
        const float norm = sc == 1 ? 3.0f / 16.0f * 0.5f : 1.0f / (2.0f * (float)sc * (wgt + 2.0f));
        const float wc = sc == 1 ? 10.0f / 16.0f * 0.5f : wgt * norm;
        if (sc <= 8 && !s->unfused_fed) {
            k_akz_hessian<<<dim3((L.w + HS_TW - 1) / HS_TW, (L.h + HS_TH - 1) / HS_TH, B), 256, akz_hess_smem(sc), st>>>(
                L.Lsm, L.Lx, L.Ly, L.Ldet, L.w, L.h, L.stride, L.istride, sc, norm, wc); ++g_afv_launches;
        } else {
        k_akz_deriv1<<<AKZ_GRID(L, B), 256, 0, st>>>(L.Lsm, s->scr[0], s->scr[1], L.w, L.h, L.stride, L.istride, sc, norm, wc); ++g_afv_launches;
        k_akz_deriv2<<<AKZ_GRID(L, B), 256, 0, st>>>(s->scr[0], s->scr[1], L.Lx, L.Ly, L.Ldet, L.w, L.h, L.stride, L.istride, sc, norm, wc); ++g_afv_launches;
        }
        k_akz_extrema<<<AKZ_GRID(L, B), 256, 0, st>>>(P, i); ++g_afv_launches;
    } }
    { AfvProfScope ps("k_akz_sort", st); k_akz_sort<<<dim3(P.nl, B), 512, 8 * 8192, st>>>(P); ++g_afv_launches; }
    { AfvProfScope ps("k_akz_select", st); k_akz_select<<<B, 256, 14 * (size_t)P.list_cap, st>>>(P); ++g_afv_launches; }
    { AfvProfScope ps("k_akz_octree", st); k_akz_octree<<<dim3(P.nlevels, B), 256, oct_work_bytes(P.oct_ncap), st>>>(P); ++g_afv_launches; }
    { AfvProfScope ps("k_akz_describe", st);
      k_akz_describe<<<dim3((cap + 7) / 8, B), 256, 0, st>>>(P, d_kps, d_desc, d_kpsize, d_n_out); ++g_afv_launches; }
    AFV_CUDA_CHECK(cudaGetLastError());
    s->P.B = B;
    return AFV_OK;
}

int afv_akaze_status(AfvAkaze* s, int B, cudaStream_t st) {
    AFV_CUDA_CHECK(cudaMemcpyAsync(s->h_status, s->P.status, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    AFV_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b)
        if (s->h_status[b]) {
            afv_set_error("akaze61: capacity exceeded in frame %d (flags 0x%x: 1 candidate lists, 4 caller cap, 8 octree)", b, s->h_status[b]);
            return AFV_ERR_CAPACITY;
        }
    return AFV_OK;
}

// taps: what = 20..24 Lt / Lsmooth / Lx / Ly / Ldet of evolution level `level`; 25 Feature_Detection list (5 floats each);
// 26 contrast factor of the frame
int afv_akaze_debug_read(AfvAkaze* s, int what, int frame, int level, void* out, long cap_bytes, long* n_bytes) {
    const AkzParams& P = s->P;
    if (what >= 20 && what <= 24) {
        if (level < 0 || level >= P.nl) { afv_set_error("afv_debug_read: bad akaze level"); return AFV_ERR_INVALID; }
        const AkzLevelG& L = P.lv[level];
        const long need = (long)L.w * L.h * 4;
        if (cap_bytes < need) { afv_set_error("buffer too small"); return AFV_ERR_INVALID; }
        const float* src = (what == 20 ? L.Lt : what == 21 ? L.Lsm : what == 22 ? L.Lx : what == 23 ? L.Ly : L.Ldet) + (long long)frame * L.istride;
        AFV_CUDA_CHECK(cudaMemcpy2D(out, (size_t)L.w * 4, src, (size_t)L.stride * 4, (size_t)L.w * 4, L.h, cudaMemcpyDeviceToHost));
        *n_bytes = need;
        return AFV_OK;
    }
    if (what == 25) {
        const int cap = (int)(cap_bytes / 20);
        float* d_out = nullptr; int* d_n = nullptr;
        AFV_CUDA_CHECK(cudaMalloc((void**)&d_out, (size_t)(cap > 0 ? cap : 1) * 20));
        AFV_CUDA_CHECK(cudaMalloc((void**)&d_n, 4));
        k_akz_tap_list<<<32, 256>>>(P, frame, d_out, cap, d_n);
        int n = 0;
        cudaError_t e = cudaMemcpy(&n, d_n, 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && n <= cap) e = cudaMemcpy(out, d_out, (size_t)n * 20, cudaMemcpyDeviceToHost);
        cudaFree(d_out); cudaFree(d_n);
        if (e != cudaSuccess) { afv_set_error("tap copy failed: %s", cudaGetErrorString(e)); return AFV_ERR_CUDA; }
        if (n > cap) { afv_set_error("buffer too small (%d keypoints)", n); return AFV_ERR_INVALID; }
        *n_bytes = (long)n * 20;
        return AFV_OK;
    }
    if (what == 26) {
        if (cap_bytes < 4) { afv_set_error("buffer too small"); return AFV_ERR_INVALID; }
        AFV_CUDA_CHECK(cudaMemcpy(out, P.kcontrast + frame, 4, cudaMemcpyDeviceToHost));
        *n_bytes = 4;
        return AFV_OK;
    }
    afv_set_error("unknown akaze tap %d", what);
    return AFV_ERR_INVALID;
}

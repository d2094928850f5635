// afv_sift.cu -- sift128 extraction (FeatureExtractor_sift128, reference src/Feature_sift128.cpp:9-134) for sm_100a.
//
// The reference delegates the arithmetic to SiftGPU (GLSL); this file implements the published algorithm in the
// parameterisation the reference selects (-fo 0 -d 3 -no 8 -e 10 -tc2 n -da -loweo, default threshold 0.02/3, 2
// orientations) plus everything the reference does around it (octave = int(log2(s/1.6454)), DistributeOctTree per octave
// with quota mnFeaturesPerLevel, descriptor row gather, merge, computeSize).  Arithmetic contract = oracle/afv_oracle_sift.c:
// IEEE float32 without contraction (this TU is compiled --fmad=false), polynomial exp / atan2 / sincos, integer
// histogram accumulation, so results are bit-identical with the oracle.  PARITY vs SiftGPU itself is UNPINNED.
//
// Kernels (batch of B frames, arenas [octave][level][frame][row][stride]):
//   k_sift_blur<R,U8>  separable Gaussian step, 128x32 tile staged in shared memory, register-tiled row / column passes,
//                      writes G[i] and DoG[i-1] = G[i] - G[i-1] in the same pass
//   k_sift_down        next octave's level -1 = every 2nd pixel of level 2
//   k_sift_detect      26-neighbour extremum + darkness-adaptive threshold + edge test + one Newton step -> candidate list
//   k_sift_sort        raster order (bitonic sort of (y, x) keys in shared memory), one CTA per (level, frame)
//   k_sift_orient      warp per candidate: 36-bin integer histogram, smoothing, <= 2 peaks
//   k_sift_select      CTA per frame: -tc2 soft limit, class_id, reference octave, stable partition per octave
//   k_sift_octree      DistributeOctTree per (reference octave, frame) (afv_octree.cuh)
//   k_sift_describe    warp per kept keypoint: 4x4x8 integer histogram, normalise / clamp / renormalise, merged output
#include "afv_common.cuh"
#include "afv_octree.cuh"
#include "afv_blur.cuh"
#include "afv_sift.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define SIFT_S 3
#define SIFT_NL 6
#define SIFT_MAX_OCT 8
#define SIFT_BORDER 5
#define SIFT_QSCALE 1048576.0f
#define SIFT_PI 3.14159265358979323846f
#define SIFT_2PI 6.28318530717958647692f
#define SIFT_MAX_R 13
#define SIFT_NLIST (SIFT_MAX_OCT * SIFT_S)

struct SiftCand { uint32_t key; float xo, yo, ls; };          // key = y << 16 | x (octave grid)

struct SiftOctG {
    int w, h, stride;            // stride in floats (multiple of 32)
    long long istride;           // floats per (level, frame) image
    float* g;                    // [6][B] images
    float* d;                    // [5][B] images
    int cand_cap;                // candidates per (frame, level) list (power of two)
    SiftCand* cand;              // [B][3][cand_cap] unordered
    SiftCand* srt;               // [B][3][cand_cap] raster order
    float4* feat;                // [B][3][2*cand_cap] {xo, yo, ls, ori (<0: empty)}
};

struct SiftParams {
    int B, no, W, H, nfeatures, nlevels, out_cap;
    int n_ini; float hX;
    int q_ext[AFV_MAX_LEVELS]; float size_norm[AFV_MAX_LEVELS];
    SiftOctG oc[SIFT_MAX_OCT];
    int* cnt;                    // [B][SIFT_NLIST] candidate counts
    int* status;                 // [B]
    int key_cap;                 // per-frame capacity of the partitioned key arrays
    float* okx; float* oky; uint32_t* oref; int* ocid;          // [B][key_cap]
    unsigned short* knode; unsigned char* kquad;                // [B][key_cap] octree scratch
    int* selinfo;                // [B][32]: 0..15 keys per reference octave, 16..31 segment offsets
    int* keep; int keep_cap;     // [B][nlevels][keep_cap] key index inside the octave's segment
    int* keepcnt;                // [B][16]
    int oct_ncap;
};

#define SIFT_ST_CAND_OVERFLOW 1
#define SIFT_ST_OCTREE_OVERFLOW 8
#define SIFT_ST_OUT_OVERFLOW 4


// ---- deterministic elementary functions (same polynomials and operation order as the oracle) -----------------------
__device__ __forceinline__ float sift_exp2(float t) {
    if (t < -126.f) t = -126.f;
    if (t > 126.f) t = 126.f;
    const float n = rintf(t);
    const float f = t - n;
    float p = 0x1.430912p-13f;
    p = p * f + 0x1.5d87fep-10f;
    p = p * f + 0x1.3b2ab6p-7f;
    p = p * f + 0x1.c6b08ep-5f;
    p = p * f + 0x1.ebfbep-3f;
    p = p * f + 0x1.62e43p-1f;
    p = p * f + 1.0f;
    return p * __uint_as_float((unsigned)((int)n + 127) << 23);
}
__device__ __forceinline__ float sift_exp(float x) { return sift_exp2(x * 0x1.715476p+0f); }

__device__ __forceinline__ float sift_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
    if (mx == 0.f) return 0.f;
    const float a = mn / mx, z = a * a;
    float p = -0x1.394942p-8f;
    p = p * z + 0x1.9256c4p-6f;
    p = p * z + -0x1.eabc6cp-5f;
    p = p * z + 0x1.974118p-4f;
    p = p * z + -0x1.1f5284p-3f;
    p = p * z + 0x1.990384p-3f;
    p = p * z + -0x1.555216p-2f;
    p = p * z + 0x1.fffffep-1f;
    float r = p * a;
    if (ay > ax) r = 0x1.921fb6p+0f - r;
    if (x < 0.f) r = SIFT_PI - r;
    if (y < 0.f) r = SIFT_2PI - r;
    if (r >= SIFT_2PI) r = r - SIFT_2PI;
    if (r < 0.f) r = 0.f;
    return r;
}

__device__ __forceinline__ void sift_sincos(float a, float* sn, float* cs) {
    const float q = rintf(a * 0x1.45f306p-1f);
    const float r = a - q * 0x1.921fb6p+0f;
    const float z = r * r;
    float s = 0x1.71de3ap-19f;
    s = s * z + -0x1.a01a02p-13f;
    s = s * z + 0x1.111112p-7f;
    s = s * z + -0x1.555556p-3f;
    s = s * z + 1.0f;
    s = s * r;
    float c = 0x1.a01a02p-16f;
    c = c * z + -0x1.6c16c2p-10f;
    c = c * z + 0x1.555556p-5f;
    c = c * z + -0.5f;
    c = c * z + 1.0f;
    const int k = ((int)q) & 3;
    if (k == 0) { *sn = s; *cs = c; }
    else if (k == 1) { *sn = c; *cs = -s; }
    else if (k == 2) { *sn = -s; *cs = -c; }
    else { *sn = -c; *cs = s; }
}

__device__ __forceinline__ float sift_sigma(float ls) { return (1.6f * 0x1.428a3p+0f) * sift_exp2(ls * (1.0f / 3.0f)); }
__device__ __forceinline__ int sift_glevel(float ls) { return min(max((int)rintf(ls) + 1, 1), 3); }

__global__ void __launch_bounds__(256) k_sift_down(const float* __restrict__ src, int sstride, long long sistride,
                                                   float* __restrict__ dst, int w, int h, int stride, long long istride) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6), f = blockIdx.z;
    if (x < w && y < h) dst[(long long)f * istride + (long long)y * stride + x] = src[(long long)f * sistride + (long long)(2 * y) * sstride + 2 * x];
}

// ------------------------------------------------------------------------------------------------------
// Detection: thread per pixel per DoG level (blockIdx.z = frame * 3 + level).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sift_detect(const __grid_constant__ SiftParams P, int o) {
    const SiftOctG& O = P.oc[o];
    const int f = blockIdx.z / SIFT_S, l = blockIdx.z - f * SIFT_S, d = l + 1;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int w = O.w, h = O.h, st = O.stride;
    if (x < SIFT_BORDER || x >= w - SIFT_BORDER || y < SIFT_BORDER || y >= h - SIFT_BORDER) return;
    const long long B = P.B;
    const float* D0 = O.d + ((d - 1) * B + f) * O.istride;
    const float* D1 = O.d + (d * B + f) * O.istride;
    const float* D2 = O.d + ((d + 1) * B + f) * O.istride;
    const int c = y * st + x;
    const float v = D1[c];
    const float g = (O.g + (d * B + f) * O.istride)[c];
    float da = 2.0f * g + 0.1f;
    if (da > 1.0f) da = 1.0f;
    const float T = (0.02f / 3.0f) * da;
    if (!(fabsf(v) > 0.8f * T)) return;
    bool mx = true, mn = true;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int p = c + dy * st + dx;
            if (dx || dy) { const float a = D1[p]; mx = mx && (v > a); mn = mn && (v < a); }
        }
    if (!mx && !mn) return;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int p = c + dy * st + dx;
            const float a = D0[p], b = D2[p];
            mx = mx && (v > a) && (v > b);
            mn = mn && (v < a) && (v < b);
        }
    if (!mx && !mn) return;
    const float dxx = (D1[c + 1] + D1[c - 1]) - 2.0f * v;
    const float dyy = (D1[c + st] + D1[c - st]) - 2.0f * v;
    const float dxy = 0.25f * ((D1[c + st + 1] - D1[c + st - 1]) - (D1[c - st + 1] - D1[c - st - 1]));
    const float det2 = dxx * dyy - dxy * dxy, tr = dxx + dyy;
    if (!(det2 > 0.f)) return;
    if (!(tr * tr * 10.0f < 121.0f * det2)) return;
    const float gx = 0.5f * (D1[c + 1] - D1[c - 1]);
    const float gy = 0.5f * (D1[c + st] - D1[c - st]);
    const float gs = 0.5f * (D2[c] - D0[c]);
    const float dss = (D2[c] + D0[c]) - 2.0f * v;
    const float dxs = 0.25f * ((D2[c + 1] - D2[c - 1]) - (D0[c + 1] - D0[c - 1]));
    const float dys = 0.25f * ((D2[c + st] - D2[c - st]) - (D0[c + st] - D0[c - st]));
    const float a00 = dyy * dss - dys * dys;
    const float a01 = dxs * dys - dxy * dss;
    const float a02 = dxy * dys - dxs * dyy;
    const float a11 = dxx * dss - dxs * dxs;
    const float a12 = dxy * dxs - dxx * dys;
    const float a22 = dxx * dyy - dxy * dxy;
    const float det3 = (dxx * a00 + dxy * a01) + dxs * a02;
    if (det3 == 0.f) return;
    const float ox = -(((a00 * gx + a01 * gy) + a02 * gs) / det3);
    const float oy = -(((a01 * gx + a11 * gy) + a12 * gs) / det3);
    const float os = -(((a02 * gx + a12 * gy) + a22 * gs) / det3);
    if (!(fabsf(ox) < 1.0f && fabsf(oy) < 1.0f && fabsf(os) < 1.0f)) return;
    const float vr = v + 0.5f * ((gx * ox + gy * oy) + gs * os);
    if (!(fabsf(vr) > T)) return;
    const int slot = atomicAdd(&P.cnt[f * SIFT_NLIST + o * SIFT_S + l], 1);
    if (slot >= O.cand_cap) { atomicOr(&P.status[f], SIFT_ST_CAND_OVERFLOW); return; }
    SiftCand r; r.key = ((uint32_t)y << 16) | (uint32_t)x; r.xo = (float)x + ox; r.yo = (float)y + oy; r.ls = (float)(d - 1) + os;
    O.cand[((long long)f * SIFT_S + l) * O.cand_cap + slot] = r;
}

// ------------------------------------------------------------------------------------------------------
// Raster order: bitonic sort of (key << 32 | slot) in shared memory; one CTA (512 threads) per (level, frame).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_sift_sort(const __grid_constant__ SiftParams P, int o) {
    extern __shared__ __align__(16) unsigned long long sk[];
    const SiftOctG& O = P.oc[o];
    const int l = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const int n = min(P.cnt[f * SIFT_NLIST + o * SIFT_S + l], O.cand_cap);
    if (n == 0) return;
    const SiftCand* cand = O.cand + ((long long)f * SIFT_S + l) * O.cand_cap;
    SiftCand* srt = O.srt + ((long long)f * SIFT_S + l) * O.cand_cap;
    int np = 1;
    while (np < n) np <<= 1;
    for (int i = tid; i < np; i += 512) sk[i] = i < n ? (((unsigned long long)cand[i].key << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= np; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np; i += 512) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = sk[i], b = sk[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { sk[i] = b; sk[ixj] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < n; i += 512) srt[i] = cand[(unsigned)(sk[i] & 0xffffffffu)];
}

// ------------------------------------------------------------------------------------------------------
// Orientation: warp per candidate.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sift_orient(const __grid_constant__ SiftParams P, int o) {
    __shared__ unsigned hist[8][36];
    __shared__ float hsm[8][36];
    const SiftOctG& O = P.oc[o];
    const int l = blockIdx.y, f = blockIdx.z, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int n = min(P.cnt[f * SIFT_NLIST + o * SIFT_S + l], O.cand_cap);
    const SiftCand* srt = O.srt + ((long long)f * SIFT_S + l) * O.cand_cap;
    float4* feat = O.feat + ((long long)f * SIFT_S + l) * 2 * O.cand_cap;
    const int w = O.w, h = O.h, st = O.stride;
    for (int r = blockIdx.x * 8 + wp; r < n; r += gridDim.x * 8) {
        const SiftCand cd = srt[r];
        const float xo = cd.xo, yo = cd.yo, sigma = sift_sigma(cd.ls);
        const float* G = O.g + ((long long)sift_glevel(cd.ls) * P.B + f) * O.istride;
        const float sw = 1.5f * sigma;
        const int R = (int)(2.0f * sw + 0.5f);
        const float inv2s2 = -1.0f / (2.0f * sw * sw);
        const int xi = (int)rintf(xo), yi = (int)rintf(yo);
        const float r2max = (float)(R * R) + 0.5f;
        for (int b = lane; b < 36; b += 32) hist[wp][b] = 0;
        __syncwarp();
        for (int j = -R; j <= R; ++j) {
            const int py = yi + j;
            if (py < 1 || py > h - 2) continue;
            for (int i = -R + lane; i <= R; i += 32) {
                const int px = xi + i;
                if (px < 1 || px > w - 2) continue;
                const float dx = (float)px - xo, dy = (float)py - yo;
                const float r2 = dx * dx + dy * dy;
                if (r2 > r2max) continue;
                const int c = py * st + px;
                const float gx = G[c + 1] - G[c - 1], gy = G[c + st] - G[c - st];
                const float mag = sqrtf(gx * gx + gy * gy);
                const float ang = sift_atan2(gy, gx);
                const float wgt = sift_exp(r2 * inv2s2);
                int b = (int)(ang * (36.0f / SIFT_2PI));
                if (b > 35) b = 35;
                atomicAdd(&hist[wp][b], (unsigned)rintf(mag * wgt * SIFT_QSCALE));
            }
        }
        __syncwarp();
        for (int b = lane; b < 36; b += 32) {
            const float hm2 = (float)hist[wp][(b + 34) % 36], hp2 = (float)hist[wp][(b + 2) % 36];
            const float hm1 = (float)hist[wp][(b + 35) % 36], hp1 = (float)hist[wp][(b + 1) % 36];
            hsm[wp][b] = ((hm2 + hp2) * 0.0625f + (hm1 + hp1) * 0.25f) + (float)hist[wp][b] * 0.375f;
        }
        __syncwarp();
        if (lane == 0) {
            const float* hs = hsm[wp];
            float maxv = 0.f;
            for (int b = 0; b < 36; ++b) if (hs[b] > maxv) maxv = hs[b];
            float bv0 = 0.f, bv1 = 0.f, bo0 = 0.f, bo1 = 0.f;
            int nn = 0;
            if (maxv > 0.f) {
                const float th = 0.8f * maxv;
                for (int b = 0; b < 36; ++b) {
                    const float lf = hs[(b + 35) % 36], rt = hs[(b + 1) % 36], c = hs[b];
                    if (c > lf && c > rt && c >= th) {
                        float bin = ((float)b + 0.5f) + 0.5f * (lf - rt) / ((lf - 2.0f * c) + rt);      // peak referred to the bin centre
                        if (bin < 0.f) bin = bin + 36.0f;
                        if (bin >= 36.0f) bin = bin - 36.0f;
                        const float oo = bin * (SIFT_2PI / 36.0f);
                        if (nn < 2) {
                            if (nn == 1 && c > bv0) { bv1 = bv0; bo1 = bo0; bv0 = c; bo0 = oo; }
                            else if (nn == 0) { bv0 = c; bo0 = oo; }
                            else { bv1 = c; bo1 = oo; }
                            ++nn;
                        } else if (c > bv0) { bv1 = bv0; bo1 = bo0; bv0 = c; bo0 = oo; }
                        else if (c > bv1) { bv1 = c; bo1 = oo; }
                    }
                }
            }
            feat[2 * r] = make_float4(xo, yo, cd.ls, nn > 0 ? bo0 : -1.0f);
            feat[2 * r + 1] = make_float4(xo, yo, cd.ls, nn > 1 ? bo1 : -1.0f);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------
// Selection: one CTA (1024 threads) per frame.  SiftGPU list order = (octave, level, raster, orientation rank) = slot
// order; -tc2 soft limit from the coarsest level; class_id = row in that list (src/Feature_sift128.cpp:84-96);
// reference octave = int(log2(s / 1.6454)) (:92); keys are partitioned per reference octave with list order kept.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sift_ref_octave(float s, int nlevels) {
    const double r = (double)s / 1.6454;
    int o = 0;
    double p = 2.0;
    while (r >= p && o < 30) { ++o; p *= 2.0; }
    return min(o, nlevels - 1);
}

__global__ void __launch_bounds__(1024) k_sift_select(const __grid_constant__ SiftParams P) {
    __shared__ int cntv[SIFT_NLIST];
    __shared__ int octcnt[16], octoff[16], octrun[16];
    __shared__ int wcnt[32][16];
    __shared__ int s_keep_from, s_cid_base;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int nlist = P.no * SIFT_S;
    if (tid < SIFT_NLIST) cntv[tid] = 0;
    if (tid < 16) { octcnt[tid] = 0; octrun[tid] = 0; }
    __syncthreads();
    // valid features per list
    for (int L = 0; L < nlist; ++L) {
        const SiftOctG& O = P.oc[L / SIFT_S];
        const int n2 = 2 * min(P.cnt[f * SIFT_NLIST + L], O.cand_cap);
        const float4* feat = O.feat + ((long long)f * SIFT_S + (L % SIFT_S)) * 2 * O.cand_cap;
        int c = 0;
        for (int i = tid; i < n2; i += 1024) c += feat[i].w >= 0.f;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0 && c) atomicAdd(&cntv[L], c);
    }
    __syncthreads();
    if (tid == 0) {
        int keep_from = 0, run = 0;
        for (int i = nlist - 1; i >= 0; --i) {
            if (run > P.nfeatures) { keep_from = i + 1; break; }
            run += cntv[i];
        }
        s_keep_from = keep_from; s_cid_base = 0;
    }
    __syncthreads();
    const int keep_from = s_keep_from;
    // keys per reference octave
    for (int L = keep_from; L < nlist; ++L) {
        const int o = L / SIFT_S;
        const SiftOctG& O = P.oc[o];
        const int n2 = 2 * min(P.cnt[f * SIFT_NLIST + L], O.cand_cap);
        const float4* feat = O.feat + ((long long)f * SIFT_S + (L % SIFT_S)) * 2 * O.cand_cap;
        const float sc = (float)(1 << o);
        for (int i = tid; i < n2; i += 1024) {
            const float4 ft = feat[i];
            if (ft.w >= 0.f) atomicAdd(&octcnt[sift_ref_octave(sift_sigma(ft.z) * sc, P.nlevels)], 1);
        }
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int k = 0; k < 16; ++k) { octoff[k] = run; run += octcnt[k]; }
        if (run > P.key_cap) atomicOr(&P.status[f], SIFT_ST_CAND_OVERFLOW);
    }
    __syncthreads();
    // ordered write
    float* okx = P.okx + (long long)f * P.key_cap; float* oky = P.oky + (long long)f * P.key_cap;
    uint32_t* oref = P.oref + (long long)f * P.key_cap; int* ocid = P.ocid + (long long)f * P.key_cap;
    for (int L = keep_from; L < nlist; ++L) {
        const int o = L / SIFT_S;
        const SiftOctG& O = P.oc[o];
        const int n2 = 2 * min(P.cnt[f * SIFT_NLIST + L], O.cand_cap);
        const float4* feat = O.feat + ((long long)f * SIFT_S + (L % SIFT_S)) * 2 * O.cand_cap;
        const float sc = (float)(1 << o);
        for (int base = 0; base < n2; base += 1024) {
            const int i = base + tid;
            bool valid = false; int oc = 0; float4 ft = make_float4(0, 0, 0, -1);
            if (i < n2) { ft = feat[i]; valid = ft.w >= 0.f; }
            if (valid) oc = sift_ref_octave(sift_sigma(ft.z) * sc, P.nlevels);
            // per-warp ballots for every octave, then prefix over warps
            int my_rank = 0, my_cid_rank = 0;
            const unsigned vb = __ballot_sync(0xffffffffu, valid);
            for (int k = 0; k < P.nlevels; ++k) {
                const unsigned m = __ballot_sync(0xffffffffu, valid && oc == k);
                if (lane == 0) wcnt[wp][k] = __popc(m);
                if (valid && oc == k) my_rank = __popc(m & ((1u << lane) - 1));
            }
            if (lane == 0) wcnt[wp][15] = __popc(vb);
            my_cid_rank = __popc(vb & ((1u << lane) - 1));
            __syncthreads();
            if (valid) {
                int before = 0, cbefore = 0;
                for (int q = 0; q < wp; ++q) { before += wcnt[q][oc]; cbefore += wcnt[q][15]; }
                const int pos = octoff[oc] + octrun[oc] + before + my_rank;
                if (pos < P.key_cap) {
                    okx[pos] = ft.x * sc; oky[pos] = ft.y * sc;
                    oref[pos] = ((uint32_t)L << 20) | (uint32_t)i;
                    ocid[pos] = s_cid_base + cbefore + my_cid_rank;
                }
            }
            __syncthreads();
            if (tid < 16) {
                int t = 0;
                for (int q = 0; q < 32; ++q) t += wcnt[q][tid];
                if (tid == 15) s_cid_base += t; else octrun[tid] += t;
            }
            __syncthreads();
        }
    }
    if (tid < 16) { P.selinfo[f * 32 + tid] = octcnt[tid]; P.selinfo[f * 32 + 16 + tid] = octoff[tid]; }
}

// ------------------------------------------------------------------------------------------------------
// DistributeOctTree per (reference octave, frame): every response is 1, so the node representative is the first key
// in list order (src/ORBextractor.cc:444-455 keeps the first maximum).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sift_octree(const __grid_constant__ SiftParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OctWork W;
    oct_carve(smem_raw, P.oct_ncap, W);
    const int oc = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const int M = min(P.selinfo[f * 32 + oc], P.key_cap), off = P.selinfo[f * 32 + 16 + oc];
    if (M == 0 || off + M > P.key_cap) { if (tid == 0) P.keepcnt[f * 16 + oc] = 0; return; }
    const long long base = (long long)f * P.key_cap + off;
    bool overflow = false;
    int size = oct_distribute(W, P.okx + base, P.oky + base, P.knode + base, P.kquad + base, M, P.q_ext[oc], P.n_ini, P.hX,
                              P.H, P.oct_ncap, tid, overflow);
    if (overflow && tid == 0) atomicOr(&P.status[f], SIFT_ST_OCTREE_OVERFLOW);
    unsigned* first = reinterpret_cast<unsigned*>(W.best);
    for (int p = tid; p < size; p += 256) first[p] = 0xffffffffu;
    __syncthreads();
    const unsigned short* knode = P.knode + base;
    for (int k = tid; k < M; k += 256) atomicMin(&first[knode[k]], (unsigned)k);
    __syncthreads();
    if (size > P.keep_cap) { if (tid == 0) atomicOr(&P.status[f], SIFT_ST_OCTREE_OVERFLOW); size = P.keep_cap; }
    int* keep = P.keep + ((long long)f * P.nlevels + oc) * P.keep_cap;
    for (int p = tid; p < size; p += 256) keep[p] = (int)first[p];
    if (tid == 0) P.keepcnt[f * 16 + oc] = size;
}

// ------------------------------------------------------------------------------------------------------
// Descriptor + merged output: warp per kept keypoint (levels ascending, octree list order inside a level).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sift_describe(const __grid_constant__ SiftParams P, afv_keypoint* __restrict__ kps,
                                                       float* __restrict__ desc, float* __restrict__ kpsize, int* __restrict__ n_out) {
    __shared__ unsigned acc[8][128];
    __shared__ unsigned squeue[8][64];
    const int f = blockIdx.y, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int j = blockIdx.x * 8 + wp;
    int total = 0, oc = -1, p = 0;
    for (int k = 0; k < P.nlevels; ++k) {
        const int c = P.keepcnt[f * 16 + k];
        if (oc < 0 && j < total + c) { oc = k; p = j - total; }
        total += c;
    }
    if (j == 0 && lane == 0) {
        if (total > P.out_cap) atomicOr(&P.status[f], SIFT_ST_OUT_OVERFLOW);
        n_out[f] = min(total, P.out_cap);
    }
    if (oc < 0 || j >= P.out_cap) return;
    const int key = P.keep[((long long)f * P.nlevels + oc) * P.keep_cap + p];
    const long long gk = (long long)f * P.key_cap + P.selinfo[f * 32 + 16 + oc] + key;
    const uint32_t ref = P.oref[gk];
    const int L = ref >> 20, slot = ref & 0xfffff, o = L / SIFT_S;
    const SiftOctG& O = P.oc[o];
    const float4 ft = (O.feat + ((long long)f * SIFT_S + (L % SIFT_S)) * 2 * O.cand_cap)[slot];
    const float xo = ft.x, yo = ft.y, ori = ft.w, sigma = sift_sigma(ft.z);
    const float* G = O.g + ((long long)sift_glevel(ft.z) * P.B + f) * O.istride;
    const int w = O.w, h = O.h, st = O.stride;
    for (int k = lane; k < 128; k += 32) acc[wp][k] = 0;
    __syncwarp();
    float sn, cs;
    sift_sincos(ori, &sn, &cs);
    const float hw = 3.0f * sigma;
    const int R = (int)rintf(hw * 0x1.6a09e6p+0f * 2.5f);
    const float cw = cs / hw, sw_ = sn / hw;
    const int xi = (int)rintf(xo), yi = (int)rintf(yo);
    // Two phases per 32 window positions (ncu: with the test and the accumulation in one loop only 56 % of the lanes did the
    // ~150-instruction accumulation, the rest had been rejected by the rotated-window test).  Phase 1: cheap test, survivors
    // are pushed (ballot + prefix) into a per-warp queue; phase 2 runs whenever 32 survivors are queued, so the accumulation
    // always executes with full warps.  Integer atomics: the changed order is exact.
    auto accumulate = [&](unsigned e) {
        const int i = (int)(e & 0xffu) - 128, jj = (int)(e >> 8) - 128;
        const int px = xi + i, py = yi + jj;
        const float dx = (float)px - xo, dy = (float)py - yo;
        const float cr = dx * cw + dy * sw_;
        const float rr = dy * cw - dx * sw_;
        const float cb = cr + 1.5f, rb = rr + 1.5f;
        const int c = py * st + px;
        const float gx = G[c + 1] - G[c - 1], gy = G[c + st] - G[c - st];
        const float mag = sqrtf(gx * gx + gy * gy);
        float ang = sift_atan2(gy, gx) - ori;
        if (ang < 0.f) ang = ang + SIFT_2PI;
        const float ob = ang * (8.0f / SIFT_2PI);
        const float wgt = sift_exp((cr * cr + rr * rr) * -0.125f);
        const float m = mag * wgt;
        const float c0f = floorf(cb), r0f = floorf(rb), o0f = floorf(ob);
        const float fc = cb - c0f, fr = rb - r0f, fo = ob - o0f;
        const int c0 = (int)c0f, r0 = (int)r0f, o0 = (int)o0f;
#pragma unroll
        for (int dr = 0; dr < 2; ++dr) {
            const int r_ = r0 + dr;
            if (r_ < 0 || r_ > 3) continue;
            const float wr = m * (dr ? fr : 1.0f - fr);
#pragma unroll
            for (int dc = 0; dc < 2; ++dc) {
                const int c_ = c0 + dc;
                if (c_ < 0 || c_ > 3) continue;
                const float wc = wr * (dc ? fc : 1.0f - fc);
#pragma unroll
                for (int dq = 0; dq < 2; ++dq) {
                    const int o_ = (o0 + dq) & 7;
                    const float wo = wc * (dq ? fo : 1.0f - fo);
                    atomicAdd(&acc[wp][(r_ * 4 + c_) * 8 + o_], (unsigned)rintf(wo * SIFT_QSCALE));
                }
            }
        }
    };
    int qn = 0;                                            // queued survivors (warp-uniform)
    for (int jj = -R; jj <= R; ++jj) {
        const int py = yi + jj;
        if (py < 1 || py > h - 2) continue;
        for (int i0 = -R; i0 <= R; i0 += 32) {
            const int i = i0 + lane, px = xi + i;
            bool ok = i <= R && px >= 1 && px <= w - 2;
            if (ok) {
                const float dx = (float)px - xo, dy = (float)py - yo;
                const float cb = (dx * cw + dy * sw_) + 1.5f, rb = (dy * cw - dx * sw_) + 1.5f;
                ok = cb > -1.0f && cb < 4.0f && rb > -1.0f && rb < 4.0f;
            }
            const unsigned msk = __ballot_sync(0xffffffffu, ok);
            if (ok) squeue[wp][qn + __popc(msk & ((1u << lane) - 1))] = (unsigned)(i + 128) | ((unsigned)(jj + 128) << 8);
            qn += __popc(msk);
            __syncwarp();
            if (qn >= 32) {
                accumulate(squeue[wp][lane]);
                const unsigned rest = lane < qn - 32 ? squeue[wp][32 + lane] : 0u;
                __syncwarp();
                if (lane < qn - 32) squeue[wp][lane] = rest;
                qn -= 32;
                __syncwarp();
            }
        }
    }
    if (lane < qn) accumulate(squeue[wp][lane]);
    __syncwarp();
    unsigned u[4];
    unsigned long long s1 = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) { u[t] = acc[wp][lane + 32 * t] >> 6; s1 += (unsigned long long)u[t] * u[t]; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    float* drow = desc + ((long long)f * P.out_cap + j) * 128;
    if (s1 == 0) {
#pragma unroll
        for (int t = 0; t < 4; ++t) drow[lane + 32 * t] = 0.f;
    } else {
        const float n1 = sqrtf((float)s1);
        unsigned q[4];
        unsigned long long s2 = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float v = (float)u[t] / n1;
            if (v > 0.2f) v = 0.2f;
            q[t] = (unsigned)rintf(v * SIFT_QSCALE);
            s2 += (unsigned long long)q[t] * q[t];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        const float n2 = sqrtf((float)s2);
#pragma unroll
        for (int t = 0; t < 4; ++t) drow[lane + 32 * t] = (float)q[t] / n2;
    }
    if (lane == 0) {
        const float sc = (float)(1 << o);
        afv_keypoint kp;
        kp.x = xo * sc; kp.y = yo * sc; kp.size = sigma * sc; kp.angle = ori; kp.response = 1.0f;
        kp.octave = oc; kp.class_id = P.ocid[gk];
        kps[(long long)f * P.out_cap + j] = kp;
        if (kpsize) kpsize[(long long)f * P.out_cap + j] = P.size_norm[oc];
    }
}

// tap: SiftGPU-order list after the -tc2 limit, gathered by class_id: out[cid] = (x, y, s, o)
__global__ void k_sift_tap_list(const __grid_constant__ SiftParams P, int f, float4* out, int cap, int* n_total) {
    int total = 0;
    for (int k = 0; k < 16; ++k) total += P.selinfo[f * 32 + k];
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_total = total;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total && g < P.key_cap; g += gridDim.x * blockDim.x) {
        const long long gk = (long long)f * P.key_cap + g;
        const uint32_t ref = P.oref[gk];
        const int L = ref >> 20, slot = ref & 0xfffff, o = L / SIFT_S;
        const SiftOctG& O = P.oc[o];
        const float4 ft = (O.feat + ((long long)f * SIFT_S + (L % SIFT_S)) * 2 * O.cand_cap)[slot];
        const float sc = (float)(1 << o);
        const int cid = P.ocid[gk];
        if (cid < cap) out[cid] = make_float4(ft.x * sc, ft.y * sc, sift_sigma(ft.z) * sc, ft.w);
    }
}

// =====================================================================================================
// host side
// =====================================================================================================
struct AfvSift {
    int nfeatures, nlevels, max_batch, max_w, max_h, cur_w, cur_h;
    float scale_factor;
    int rad[6];
    AfvBlurTaps taps[6];
    int use_tma;                 // AFV_BLUR_NO_TMA=1 forces the clamped-load staging on every tile (A/B check)
    std::vector<void*> allocs;
    SiftParams P;
    float* g[SIFT_MAX_OCT]; float* d[SIFT_MAX_OCT];
    SiftCand* cand[SIFT_MAX_OCT]; SiftCand* srt[SIFT_MAX_OCT]; float4* feat[SIFT_MAX_OCT];
    int cand_cap[SIFT_MAX_OCT];
    uint8_t* gray_stage; int* h_status;
    int max_no;
};

static int sift_num_octaves(int w, int h) {
    int m = w < h ? w : h, lg = 0;
    while ((1 << (lg + 1)) <= m) ++lg;
    int o = lg - 3;
    if (o > SIFT_MAX_OCT) o = SIFT_MAX_OCT;
    if (o < 1) o = 1;
    return o;
}

static int sift_gauss_kernel(double sigma, float* taps, int max_r) {
    int r = (int)ceil(4.0 * sigma);
    if (r < 1) r = 1;
    if (r > max_r) return -1;
    double sum = 0.0, wd[64];
    for (int j = 0; j <= r; ++j) { wd[j] = exp(-(double)(j * j) / (2.0 * sigma * sigma)); sum += j ? 2.0 * wd[j] : wd[j]; }
    for (int j = 0; j <= r; ++j) taps[j] = (float)(wd[j] / sum);
    return r;
}

template <typename T>
static int sift_alloc(AfvSift* s, T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 256);
    if (e != cudaSuccess) { afv_set_error("cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e)); return AFV_ERR_CUDA; }
    s->allocs.push_back(q);
    *p = (T*)q;
    return AFV_OK;
}

// the five incremental steps have fixed radii (5, 7, 8, 10, 13) and the base step 7 for the SiftGPU sigmas
static const int k_expected_rad[6] = {7, 5, 7, 8, 10, 13};

void afv_sift_destroy(AfvSift* s) {
    if (!s) return;
    for (void* p : s->allocs) cudaFree(p);
    if (s->h_status) cudaFreeHost(s->h_status);
    delete s;
}

static int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

int afv_sift_create(AfvSift** out, int nfeatures, int nlevels, float scale_factor, int max_batch, int max_w, int max_h) {
    *out = nullptr;
    if (max_w > 4095 || max_h > 4095) { afv_set_error("sift128: frame dimension > 4095 not supported"); return AFV_ERR_INVALID; }
    AfvSift* s = new AfvSift();
    memset(&s->P, 0, sizeof(s->P));
    s->nfeatures = nfeatures; s->nlevels = nlevels; s->scale_factor = scale_factor;
    s->max_batch = max_batch; s->max_w = max_w; s->max_h = max_h; s->cur_w = s->cur_h = 0;
    s->gray_stage = nullptr; s->h_status = nullptr;
    { const char* e = getenv("AFV_BLUR_NO_TMA"); s->use_tma = !(e && e[0] == '1'); }
    // blur taps (same double arithmetic as the oracle)
    const double k = pow(2.0, 1.0 / SIFT_S), sigma0 = 1.6 * k, sigman = 0.5;
    double dsig[6];
    const double s_m1 = sigma0 / k;
    dsig[0] = sqrt(s_m1 * s_m1 - sigman * sigman);
    const double dsigma0 = sigma0 * sqrt(1.0 - 1.0 / (k * k));
    for (int l = 0; l <= 4; ++l) dsig[l + 1] = dsigma0 * pow(k, (double)l);
    float taps[6][16];
    memset(taps, 0, sizeof(taps));
    for (int i = 0; i < 6; ++i) {
        s->rad[i] = sift_gauss_kernel(dsig[i], taps[i], SIFT_MAX_R);
        if (s->rad[i] != k_expected_rad[i]) { afv_set_error("sift128: internal: blur radius %d of step %d unexpected", s->rad[i], i); delete s; return AFV_ERR_INVALID; }
    }
    int rc = AFV_OK;
    for (int i = 0; i < 6; ++i) memcpy(s->taps[i].t, taps[i], sizeof(float) * 16);
    {
        cudaError_t e = afv_blur_cfg<7, 1>();
        if (e == cudaSuccess) e = afv_blur_cfg<5, 0>();
        if (e == cudaSuccess) e = afv_blur_cfg<7, 0>();
        if (e == cudaSuccess) e = afv_blur_cfg<8, 0>();
        if (e == cudaSuccess) e = afv_blur_cfg<10, 0>();
        if (e == cudaSuccess) e = afv_blur_cfg<13, 0>();
        if (e != cudaSuccess) { afv_set_error("cudaFuncSetAttribute(blur) failed: %s", cudaGetErrorString(e)); delete s; return AFV_ERR_CUDA; }
    }
    s->max_no = sift_num_octaves(max_w, max_h);
    const size_t B = (size_t)max_batch;
    int ow = max_w, oh = max_h, key_cap = 0;
    for (int o = 0; o < s->max_no && rc == AFV_OK; ++o) {
        const size_t img = (size_t)((ow + 31) & ~31) * oh;
        int cap = pow2ceil(ow * oh / 64);
        if (cap < 256) cap = 256;
        if (cap > 16384) cap = 16384;
        s->cand_cap[o] = cap;
        key_cap += SIFT_S * 2 * cap;
        if ((rc = sift_alloc(s, &s->g[o], img * SIFT_NL * B))) break;
        if ((rc = sift_alloc(s, &s->d[o], img * (SIFT_NL - 1) * B))) break;
        if ((rc = sift_alloc(s, &s->cand[o], (size_t)cap * SIFT_S * B))) break;
        if ((rc = sift_alloc(s, &s->srt[o], (size_t)cap * SIFT_S * B))) break;
        if ((rc = sift_alloc(s, &s->feat[o], (size_t)cap * 2 * SIFT_S * B))) break;
        ow /= 2; oh /= 2;
    }
    SiftParams& P = s->P;
    P.key_cap = key_cap;
    int maxq = 0;
    if (rc == AFV_OK) {
        // mnFeaturesPerLevel (reference src/FeatureExtractor.cpp:97-108) and computeSize (:132-142, GetKeypointSize =
        // powf(scaleFactor0, octave), src/Feature_sift128.cpp:124-126)
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
        if (oct_work_bytes(P.oct_ncap) > 227 * 1024) { afv_set_error("sift128: nfeatures too large for the octree workspace"); rc = AFV_ERR_INVALID; }
    }
    if (rc == AFV_OK) rc = sift_alloc(s, &P.cnt, SIFT_NLIST * B);
    if (rc == AFV_OK) rc = sift_alloc(s, &P.status, B);
    if (rc == AFV_OK) rc = sift_alloc(s, &P.okx, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = sift_alloc(s, &P.oky, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = sift_alloc(s, &P.oref, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = sift_alloc(s, &P.ocid, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = sift_alloc(s, &P.knode, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = sift_alloc(s, &P.kquad, (size_t)key_cap * B);
    if (rc == AFV_OK) rc = sift_alloc(s, &P.selinfo, 32 * B);
    if (rc == AFV_OK) rc = sift_alloc(s, &P.keep, (size_t)P.keep_cap * nlevels * B);
    if (rc == AFV_OK) rc = sift_alloc(s, &P.keepcnt, 16 * B);
    if (rc == AFV_OK) rc = sift_alloc(s, &s->gray_stage, (size_t)max_w * max_h * B);
    if (rc == AFV_OK) {
        cudaError_t e = cudaMallocHost((void**)&s->h_status, sizeof(int) * B);
        if (e != cudaSuccess) { afv_set_error("cudaMallocHost failed: %s", cudaGetErrorString(e)); rc = AFV_ERR_CUDA; }
    }
    if (rc == AFV_OK) {
        cudaError_t e = cudaFuncSetAttribute(k_sift_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sift_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oct_work_bytes(P.oct_ncap));
        if (e != cudaSuccess) { afv_set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); rc = AFV_ERR_CUDA; }
    }
    if (rc != AFV_OK) { afv_sift_destroy(s); return rc; }
    *out = s;
    return AFV_OK;
}

uint8_t* afv_sift_stage(AfvSift* s) { return s->gray_stage; }

static int sift_configure(AfvSift* s, int w, int h, int B) {
    if (w > s->max_w || h > s->max_h || w < 64 || h < 64) {
        afv_set_error("frame %dx%d outside the extractor's configured range (64..%d x 64..%d)", w, h, s->max_w, s->max_h);
        return AFV_ERR_INVALID;
    }
    SiftParams& P = s->P;
    P.B = B; P.W = w; P.H = h; P.nfeatures = s->nfeatures; P.nlevels = s->nlevels;
    P.no = sift_num_octaves(w, h);
    P.n_ini = (int)round((double)((float)w / (float)h));
    if (P.n_ini < 1) { afv_set_error("portrait frames with w/h < 0.5 are not supported (reference divides by zero)"); return AFV_ERR_INVALID; }
    P.hX = (float)w / (float)P.n_ini;
    int ow = w, oh = h;
    for (int o = 0; o < P.no; ++o) {
        SiftOctG& O = P.oc[o];
        O.w = ow; O.h = oh; O.stride = (ow + 31) & ~31; O.istride = (long long)O.stride * oh;
        O.g = s->g[o]; O.d = s->d[o]; O.cand_cap = s->cand_cap[o]; O.cand = s->cand[o]; O.srt = s->srt[o]; O.feat = s->feat[o];
        ow /= 2; oh /= 2;
    }
    s->cur_w = w; s->cur_h = h;
    return AFV_OK;
}

int afv_sift_run(AfvSift* s, const uint8_t* d_gray, int B, int w, int h, int stride, long frame_stride, afv_keypoint* d_kps,
                 float* d_desc, float* d_kpsize, int cap, int* d_n_out, cudaStream_t st) {
    if (B < 1 || B > s->max_batch) { afv_set_error("batch %d outside 1..%d", B, s->max_batch); return AFV_ERR_INVALID; }
    int rc = sift_configure(s, w, h, B);
    if (rc) return rc;
    SiftParams P = s->P;
    P.out_cap = cap;
    AFV_CUDA_CHECK(cudaMemsetAsync(P.cnt, 0, sizeof(int) * SIFT_NLIST * B, st));
    AFV_CUDA_CHECK(cudaMemsetAsync(P.status, 0, sizeof(int) * B, st));
    // scale space
    for (int o = 0; o < P.no; ++o) {
        const SiftOctG& O = P.oc[o];
        auto G = [&](int i) { return O.g + (long long)i * B * O.istride; };
        auto D = [&](int i) { return O.d + (long long)i * B * O.istride; };
        { AfvProfScope ps("k_sift_blur", st);
        if (o == 0) { afv_blur_launch<7, 1>(d_gray, stride, frame_stride, G(0), nullptr, O.w, O.h, O.stride, O.istride, s->taps[0], B, st); ++g_afv_launches; }
        else {
            const SiftOctG& Q = P.oc[o - 1];
            dim3 g((O.w + 63) / 64, (O.h + 3) / 4, B);
            k_sift_down<<<g, 256, 0, st>>>(Q.g + (long long)SIFT_S * B * Q.istride, Q.stride, Q.istride, G(0), O.w, O.h, O.stride, O.istride);
            ++g_afv_launches;
        }
        { CUtensorMap tm; const bool ok = s->use_tma && afv_blur_tmap(&tm, O.g, O.w, O.h, O.stride, O.istride, (long long)SIFT_NL * B, 5);
          afv_blur_launch<5, 0>(G(0), O.stride, O.istride, G(1), D(0), O.w, O.h, O.stride, O.istride, s->taps[1], B, st, ok ? &tm : nullptr, 0 * B); ++g_afv_launches; }
        { CUtensorMap tm; const bool ok = s->use_tma && afv_blur_tmap(&tm, O.g, O.w, O.h, O.stride, O.istride, (long long)SIFT_NL * B, 7);
          afv_blur_launch<7, 0>(G(1), O.stride, O.istride, G(2), D(1), O.w, O.h, O.stride, O.istride, s->taps[2], B, st, ok ? &tm : nullptr, 1 * B); ++g_afv_launches; }
        { CUtensorMap tm; const bool ok = s->use_tma && afv_blur_tmap(&tm, O.g, O.w, O.h, O.stride, O.istride, (long long)SIFT_NL * B, 8);
          afv_blur_launch<8, 0>(G(2), O.stride, O.istride, G(3), D(2), O.w, O.h, O.stride, O.istride, s->taps[3], B, st, ok ? &tm : nullptr, 2 * B); ++g_afv_launches; }
        { CUtensorMap tm; const bool ok = s->use_tma && afv_blur_tmap(&tm, O.g, O.w, O.h, O.stride, O.istride, (long long)SIFT_NL * B, 10);
          afv_blur_launch<10, 0>(G(3), O.stride, O.istride, G(4), D(3), O.w, O.h, O.stride, O.istride, s->taps[4], B, st, ok ? &tm : nullptr, 3 * B); ++g_afv_launches; }
        { CUtensorMap tm; const bool ok = s->use_tma && afv_blur_tmap(&tm, O.g, O.w, O.h, O.stride, O.istride, (long long)SIFT_NL * B, 13);
          afv_blur_launch<13, 0>(G(4), O.stride, O.istride, G(5), D(4), O.w, O.h, O.stride, O.istride, s->taps[5], B, st, ok ? &tm : nullptr, 4 * B); ++g_afv_launches; }
        }
    }
    { AfvProfScope ps("k_sift_detect", st);
    for (int o = 0; o < P.no; ++o) {
        const SiftOctG& O = P.oc[o];
        if (O.w <= 2 * SIFT_BORDER || O.h <= 2 * SIFT_BORDER) continue;
        dim3 g((O.w + 31) / 32, (O.h + 7) / 8, B * SIFT_S);
        k_sift_detect<<<g, 256, 0, st>>>(P, o); ++g_afv_launches;
    } }
    { AfvProfScope ps("k_sift_sort_orient", st);
    for (int o = 0; o < P.no; ++o) {
        const SiftOctG& O = P.oc[o];
        k_sift_sort<<<dim3(SIFT_S, B), 512, 8 * (size_t)O.cand_cap, st>>>(P, o); ++g_afv_launches;
        int gx = O.cand_cap / 8; if (gx > 48) gx = 48;
        k_sift_orient<<<dim3(gx, SIFT_S, B), 256, 0, st>>>(P, o); ++g_afv_launches;
    } }
    { AfvProfScope ps("k_sift_select", st); k_sift_select<<<B, 1024, 0, st>>>(P); ++g_afv_launches; }
    { AfvProfScope ps("k_sift_octree", st);
      k_sift_octree<<<dim3(P.nlevels, B), 256, oct_work_bytes(P.oct_ncap), st>>>(P); ++g_afv_launches; }
    { AfvProfScope ps("k_sift_describe", st);
      k_sift_describe<<<dim3((cap + 7) / 8, B), 256, 0, st>>>(P, d_kps, d_desc, d_kpsize, d_n_out); ++g_afv_launches; }
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

int afv_sift_status(AfvSift* s, int B, cudaStream_t st) {
    AFV_CUDA_CHECK(cudaMemcpyAsync(s->h_status, s->P.status, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    AFV_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b)
        if (s->h_status[b]) {
            afv_set_error("sift128: capacity exceeded in frame %d (flags 0x%x: 1 candidate lists, 4 caller cap, 8 octree)", b, s->h_status[b]);
            return AFV_ERR_CAPACITY;
        }
    return AFV_OK;
}

// taps: what = 10 Gaussian image, 11 DoG image (level = octave * 8 + index), 12 detected list (x, y, s, o) in SiftGPU order
int afv_sift_debug_read(AfvSift* s, int what, int frame, int level, void* out, long cap_bytes, long* n_bytes) {
    const SiftParams& P = s->P;
    if (what == 10 || what == 11) {
        const int o = level / 8, i = level % 8;
        if (o < 0 || o >= P.no || i < 0 || i >= (what == 10 ? SIFT_NL : SIFT_NL - 1)) { afv_set_error("afv_debug_read: bad sift level"); return AFV_ERR_INVALID; }
        const SiftOctG& O = P.oc[o];
        const long need = (long)O.w * O.h * 4;
        if (cap_bytes < need) { afv_set_error("buffer too small"); return AFV_ERR_INVALID; }
        const float* src = (what == 10 ? O.g : O.d) + ((long long)i * P.B + frame) * O.istride;
        AFV_CUDA_CHECK(cudaMemcpy2D(out, (size_t)O.w * 4, src, (size_t)O.stride * 4, (size_t)O.w * 4, O.h, cudaMemcpyDeviceToHost));
        *n_bytes = need;
        return AFV_OK;
    }
    if (what == 12) {
        const int cap = (int)(cap_bytes / 16);
        float4* d_out = nullptr; int* d_n = nullptr;
        AFV_CUDA_CHECK(cudaMalloc((void**)&d_out, (size_t)(cap > 0 ? cap : 1) * 16));
        AFV_CUDA_CHECK(cudaMalloc((void**)&d_n, 4));
        k_sift_tap_list<<<64, 256>>>(P, frame, d_out, cap, d_n);
        int n = 0;
        cudaError_t e = cudaMemcpy(&n, d_n, 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && n <= cap) e = cudaMemcpy(out, d_out, (size_t)n * 16, cudaMemcpyDeviceToHost);
        cudaFree(d_out); cudaFree(d_n);
        if (e != cudaSuccess) { afv_set_error("tap copy failed: %s", cudaGetErrorString(e)); return AFV_ERR_CUDA; }
        if (n > cap) { afv_set_error("buffer too small (%d features)", n); return AFV_ERR_INVALID; }
        *n_bytes = (long)n * 16;
        return AFV_OK;
    }
    afv_set_error("unknown sift tap %d", what);
    return AFV_ERR_INVALID;
}

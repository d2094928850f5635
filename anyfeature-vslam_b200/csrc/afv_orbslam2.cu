// afv_orbslam2.cu -- hand-written sm_100a kernels of the reference's VANILLA ORB-SLAM2 extractor:
// FeatureExtractor::operator()(..., vanillaOrbslam) of src/ORBextractor.cc:568-645 as built with VANILLA_ORB_SLAM2
// (include/Definitions.h:8), behind afv_extractor_create(AFV_FEAT_ORB32_VANILLA).
//   k_os2_resize    ComputePyramid (:647-674): cv::resize INTER_LINEAR, 11-bit fixed point, level l from level l-1
//   k_os2_score     dense FAST-9/16 corner score at minThFAST over the detection rectangle of a level
//   k_os2_cells     ComputeKeyPointsOctTree (:464-531): W = 30 cells, FAST(iniThFAST) + 3x3 NMS confined to the cell, FAST(minThFAST)
//                   when the cell is empty; warp per cell, count pass then ordered emit pass (cell-major, raster inside the cell)
//   k_os2_octree    DistributeOctTree (:239-458) with the border rectangle as bounds (afv_octree.cuh), first key in push order
//                   wins among equal FAST scores
//   k_os2_blur      GaussianBlur(7x7, sigma 2) of the clone()d level (:603-604): OpenCV's 8.8 / 16.16 fixed-point path
//   k_os2_describe  IC_Angle (:138-166) on the un-blurred level + computeOrbDescriptor (include/FeatureExtractor.h:178-217) on the
//                   blurred one, merge (:613-626) and the size override (:629-639)
// Arithmetic is integer except the angle / rotation, which use the same explicit float sequence as the orb32 path.  Results are
// bit-exact with oracle/afv_oracle_orbslam2.c, which is itself checked against the reference's compiled code running on cv2.
#include "afv_common.cuh"
#include "afv_octree.cuh"
#include "afv_orbslam2.h"
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <vector>

#define OS2_MINB 16                   // EDGE_THRESHOLD - 3 (src/ORBextractor.cc:71, :468)
#define OS2_ST_DET_OVERFLOW 2
#define OS2_ST_OUT_OVERFLOW 4
#define OS2_ST_OCTREE_OVERFLOW 8

struct Os2Level {
    int w, h, stride; long long fstride;                 // arenas: [frame][row][stride]
    const uint8_t* img; int img_stride; long long img_fstride;   // un-blurred level (level 0 = the caller's frames)
    uint8_t* blur; uint8_t* score;
    int ncols, nrows, wcell, hcell, cell_base, ncells;   // cell grid of ComputeKeyPointsOctTree
    int q, n_ini, octH; float hX;                        // octree quota and roots
    int det_cap; uint32_t* det;                          // [frame][det_cap] (x - 16) | (y - 16) << 12 | score << 24, push order
    float* kx; float* ky; unsigned short* knode; unsigned char* kquad;   // octree key scratch [frame][det_cap]
    int keep_cap; uint2* keep;                           // [frame][keep_cap] {x | y << 12 (level coords), score as float bits}
    const uint2* xtab; const uint2* ytab;                // resize: {source index, c0 | c1 << 16}
    float sf, size_norm;                                 // mvScaleFactor[l], computeSize value
};
struct Os2Params {
    int nlevels, B, ini_th, min_th, out_cap, ncells_total, oct_ncap;
    int* cellcnt;                                        // [frame][ncells_total] count | 1 << 30 when the iniThFAST set is used
    uint32_t* cellmask;                                  // [frame][ncells_total][32] per-column row masks of the chosen maxima (cells <= 32 x 32)
    int* counts; int* status;
    Os2Level lv[AFV_MAX_LEVELS];
};

static __constant__ __align__(16) int8_t c_os2_pattern[1024] = {
#include "orb_pattern.inc"
};

// ---------------------------------------------------------------------------------------------------------------------------
// cv::resize INTER_LINEAR (8UC1): D = S[sx] * a0 + S[sx + 1] * a1; out = (((b0 * (D0 >> 4)) >> 16) + ((b1 * (D1 >> 4)) >> 16) + 2) >> 2
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_os2_resize(const __grid_constant__ Os2Params P, int l) {
    const Os2Level& D = P.lv[l];
    const Os2Level& S = P.lv[l - 1];
    const int f = blockIdx.z, lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 128 + 4 * lane, y = blockIdx.y * 8 + wrp;
    if (x0 >= D.w || y >= D.h) return;
    const uint2 yt = D.ytab[y];
    const int b0 = (int)(yt.y & 0xffff), b1 = (int)(yt.y >> 16);
    const int s0 = min(max((int)yt.x, 0), S.h - 1), s1 = min(max((int)yt.x + 1, 0), S.h - 1);
    const uint8_t* src = S.img + (long long)f * S.img_fstride;
    const uint8_t* r0 = src + (long long)s0 * S.img_stride;
    const uint8_t* r1 = src + (long long)s1 * S.img_stride;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = min(x0 + k, D.w - 1);
        const uint2 xt = D.xtab[x];
        const int sx = (int)xt.x, sx1 = min(sx + 1, S.w - 1), a0 = (int)(xt.y & 0xffff), a1 = (int)(xt.y >> 16);
        const int D0 = r0[sx] * a0 + r0[sx1] * a1;
        const int D1 = r1[sx] * a0 + r1[sx1] * a1;
        const int v = (((b0 * (D0 >> 4)) >> 16) + ((b1 * (D1 >> 4)) >> 16) + 2) >> 2;
        out |= (uint32_t)min(max(v, 0), 255) << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(D.img) + (long long)f * D.img_fstride + (long long)y * D.img_stride + x0) = out;   // rows padded to 128 B
}

// ---------------------------------------------------------------------------------------------------------------------------
// Dense FAST-9/16 corner score (OpenCV cornerScore<16>: max over the 16 arcs of 9 of min|v - p| of one sign, minus 1) for every
// pixel of the detection rectangle [19, w - 19) x [19, h - 19) that is a corner at minThFAST, else 0.  A pixel is a corner at
// threshold t iff score >= t, so ONE map serves both FAST(iniThFAST) and the FAST(minThFAST) fallback of a cell.
// ---------------------------------------------------------------------------------------------------------------------------
#define SC_W 64
#define SC_H 16
#define SC_P 72
__global__ void __launch_bounds__(256) k_os2_score(const __grid_constant__ Os2Params P, int l) {
    __shared__ __align__(4) uint8_t tile[SC_H + 6][SC_P];
    __shared__ uint16_t slist[SC_W * SC_H];              // pixels that pass the compass pre-test: row << 8 | column
    __shared__ int nlist;
    const Os2Level& L = P.lv[l];
    const int f = blockIdx.z, tid = threadIdx.x, lane = tid & 31;
    const int tx0 = 19 + blockIdx.x * SC_W, ty0 = 19 + blockIdx.y * SC_H;
    const uint8_t* img = L.img + (long long)f * L.img_fstride;
    if (tid == 0) nlist = 0;
    for (int i = tid; i < (SC_H + 6) * (SC_W + 6); i += 256) {
        const int r = i / (SC_W + 6), c = i - r * (SC_W + 6);
        const int gx = tx0 - 3 + c, gy = ty0 - 3 + r;
        tile[r][c] = (gx < L.w && gy < L.h) ? img[(long long)gy * L.img_stride + gx] : 0;      // gx, gy >= 16
    }
    __syncthreads();
    const int t = P.min_th;
    uint8_t* sc = L.score + (long long)f * L.fstride;
    // stage 1: an arc of 9 of the 16 circle pixels holds two ADJACENT compass points, so a corner needs (up or down) and (left or right)
    // beyond the threshold with one sign.  At minThFAST = 7 a third of all pixels still pass, i.e. every warp would run the ~150
    // instruction score for all of its pixels: the survivors are compacted first and scored with full warps (as in k_fast).
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
        const int i = tid + 256 * rep;
        const int r = i >> 6, c = i & 63;
        const int gx = tx0 + c, gy = ty0 + r;
        const bool in = gx < L.w - 19 && gy < L.h - 19;
        bool pass = false;
        if (in) {
            const uint8_t* p = &tile[r + 3][c + 3];
            const int v = p[0], hi = v + t, lo = v - t;
            const int up = p[-3 * SC_P], dn = p[3 * SC_P], lf = p[-3], rt = p[3];
            pass = (max(up, dn) > hi && max(lf, rt) > hi) || (min(up, dn) < lo && min(lf, rt) < lo);
            if (!pass) sc[(long long)gy * L.stride + gx] = 0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            int base = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) base = atomicAdd(&nlist, __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (pass) slist[base + __popc(m & ((1u << lane) - 1))] = (uint16_t)((r << 8) | c);
        }
    }
    __syncthreads();
    // stage 2: OpenCV cornerScore<16> of every survivor (packed (d, -d) sliding minimum over 9 by doubling, see k_fast in afv_orb.cu)
    const int n = nlist;
    for (int j = tid; j < n; j += 256) {
        const int e0 = slist[j], r = e0 >> 8, c = e0 & 255;
        const uint8_t* p = &tile[r + 3][c + 3];
        const int v = p[0];
        const uint32_t A = (uint32_t)(256 + v) | ((uint32_t)(256 - v) << 16);
        uint32_t e[16];
#define OS2_CIRC(F) F(0, 0, 3) F(1, 1, 3) F(2, 2, 2) F(3, 3, 1) F(4, 3, 0) F(5, 3, -1) F(6, 2, -2) F(7, 1, -3) \
                    F(8, 0, -3) F(9, -1, -3) F(10, -2, -2) F(11, -3, -1) F(12, -3, 0) F(13, -3, 1) F(14, -2, 2) F(15, -1, 3)
#define FDIFF(k, dx, dy) e[k] = (uint32_t)p[(dy) * SC_P + (dx)] * 0xffffu + A;
        OS2_CIRC(FDIFF)
#undef FDIFF
        uint32_t m2[16], m4[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) m2[k] = __vmins2(e[k], e[(k + 1) & 15]);
#pragma unroll
        for (int k = 0; k < 16; ++k) m4[k] = __vmins2(m2[k], m2[(k + 2) & 15]);
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = __vmaxs2(acc, __vmins2(__vmins2(m4[k], m4[(k + 4) & 15]), e[(k + 8) & 15]));
        const int bd = (int)(acc & 0xffffu) - 256, bb = (int)(acc >> 16) - 256;
        const int best = bd > bb ? bd : bb;
        sc[(long long)(ty0 + r) * L.stride + tx0 + c] = (uint8_t)(best > t ? best - 1 : 0);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Cells of ComputeKeyPointsOctTree: warp per cell.  EMIT = false: count the NMS maxima with score >= iniThFAST and all maxima;
// EMIT = true: write the chosen set at its place in the level's push order (prefix over the preceding cells of the level).
// ---------------------------------------------------------------------------------------------------------------------------
#define CL_P 64
#define CL_R 62
template <bool EMIT>
__global__ void __launch_bounds__(256) k_os2_cells(const __grid_constant__ Os2Params P) {
    __shared__ __align__(4) uint8_t tiles[8][CL_R][CL_P];
    const int f = blockIdx.y, lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int cell = blockIdx.x * 8 + wrp;
    if (cell >= P.ncells_total) return;
    int l = 0;
    while (l + 1 < P.nlevels && cell >= P.lv[l + 1].cell_base) ++l;
    const Os2Level& L = P.lv[l];
    const int ci = cell - L.cell_base, i = ci / L.ncols, j = ci - i * L.ncols;
    const int maxBX = L.w - OS2_MINB, maxBY = L.h - OS2_MINB;
    const int y0 = OS2_MINB + i * L.hcell, x0 = OS2_MINB + j * L.wcell;
    const int y1 = min(y0 + L.hcell + 6, maxBY), x1 = min(x0 + L.wcell + 6, maxBX);
    const bool skip = (y0 >= maxBY - 3) || (x0 >= maxBX - 6);                      // src/ORBextractor.cc:492-493, :502-503
    const int vx0 = x0 + 3, vy0 = y0 + 3;
    const int wv = skip ? 0 : max(x1 - 3 - vx0, 0), hv = skip ? 0 : max(y1 - 3 - vy0, 0);
    int* cc = P.cellcnt + (long long)f * P.ncells_total;
    uint8_t (*T)[CL_P] = tiles[wrp];
    const uint8_t* sc = L.score + (long long)f * L.fstride;
    if (wv <= 32 && hv <= 32) {
        // Common case (cells are ceil(width / floor(width / 30)) wide: <= 32 for every level with >= 15 cell columns): lane = column,
        // the score rows of the cell are read straight from the map (one byte per lane and row, all loads in flight together),
        // horizontal neighbours come by shuffle, and a neighbour outside the cell is 0 exactly as FAST's sub-image NMS sees it.
        // Pass 1 leaves, per lane, a 32-bit ROW MASK of the maxima of the chosen set; pass 2 only walks those masks in raster order.
        uint32_t* cm = P.cellmask + ((long long)f * P.ncells_total + cell) * 32;
        uint32_t* det = L.det + (long long)f * L.det_cap;
        const bool act = lane < wv;
        const uint8_t* col = sc + (long long)vy0 * L.stride + vx0 + lane;
        if (!EMIT) {
            int v[32];
#pragma unroll
            for (int r = 0; r < 32; ++r) v[r] = (act && r < hv) ? (int)col[(long long)r * L.stride] : 0;
            uint32_t m_all = 0, m_ini = 0;
            int h3_prev = 0;
            int lf = __shfl_up_sync(0xffffffffu, v[0], 1), rt = __shfl_down_sync(0xffffffffu, v[0], 1);
            if (lane == 0) lf = 0;
            if (lane == 31) rt = 0;
            int hlr = max(lf, rt), h3 = max(hlr, v[0]);
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                int hlr_n = 0, h3_n = 0;
                if (r + 1 < 32) {
                    int l2 = __shfl_up_sync(0xffffffffu, v[r + 1], 1), r2 = __shfl_down_sync(0xffffffffu, v[r + 1], 1);
                    if (lane == 0) l2 = 0;
                    if (lane == 31) r2 = 0;
                    hlr_n = max(l2, r2); h3_n = max(hlr_n, v[r + 1]);
                }
                const int s = v[r];
                const bool mx = s > 0 && s > hlr && s > h3_prev && s > h3_n;
                m_all |= (uint32_t)mx << r;
                m_ini |= (uint32_t)(mx && s >= P.ini_th) << r;
                h3_prev = h3; hlr = hlr_n; h3 = h3_n;
            }
            const int n_all = __reduce_add_sync(0xffffffffu, __popc(m_all)), n_ini = __reduce_add_sync(0xffffffffu, __popc(m_ini));
            cm[lane] = n_ini ? m_ini : m_all;
            if (lane == 0) cc[cell] = n_ini ? (n_ini | (1 << 30)) : n_all;
        } else {
            int part = 0;
            for (int k = L.cell_base + lane; k < cell; k += 32) part += cc[k] & 0x3fffffff;
            const int base = __reduce_add_sync(0xffffffffu, part);
            const uint32_t m = cm[lane];
            int run = 0;
            uint32_t rows_any = __reduce_or_sync(0xffffffffu, m);
            while (rows_any) {
                const int r = __ffs(rows_any) - 1;
                rows_any &= rows_any - 1;
                const bool bit = (m >> r) & 1u;
                const unsigned rm = __ballot_sync(0xffffffffu, bit);
                if (bit) {
                    const int pos = base + run + __popc(rm & ((1u << lane) - 1));
                    const uint32_t sv = col[(long long)r * L.stride];
                    if (pos < L.det_cap) det[pos] = (uint32_t)(vx0 + lane - OS2_MINB) | ((uint32_t)(vy0 + r - OS2_MINB) << 12) | (sv << 24);
                }
                run += __popc(rm);
            }
            if (lane == 0 && ci == L.ncells - 1) {
                int tot = base + run;
                if (tot > L.det_cap) { atomicOr(&P.status[f], OS2_ST_DET_OVERFLOW); tot = L.det_cap; }
                P.counts[afv_cnt_idx(f, AFV_CNT_DET, l)] = tot;
            }
        }
        return;
    }
    // stage the valid region with a zero ring: FAST's non-max suppression only sees scores of its own sub-image
    for (int r = 0; r < hv + 2; ++r)
        for (int c = lane; c < wv + 2; c += 32) {
            const bool in = r >= 1 && r <= hv && c >= 1 && c <= wv;
            T[r][c] = in ? sc[(long long)(vy0 + r - 1) * L.stride + vx0 + c - 1] : 0;
        }
    __syncwarp();
    int base = 0, use_ini = 0;
    if (EMIT) {
        int part = 0;
        for (int k = L.cell_base + lane; k < cell; k += 32) part += cc[k] & 0x3fffffff;
#pragma unroll
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        base = part;
        use_ini = (cc[cell] >> 30) & 1;
    }
    int n_ini = 0, n_all = 0;
    uint32_t* det = L.det + (long long)f * L.det_cap;
    for (int r = 1; r <= hv; ++r)
        for (int c0 = 1; c0 <= wv; c0 += 32) {
            const int c = c0 + lane;
            int s = 0;
            bool mx = false;
            if (c <= wv) {
                s = T[r][c];
                mx = s > 0 && s > T[r][c - 1] && s > T[r][c + 1] && s > T[r - 1][c - 1] && s > T[r - 1][c] && s > T[r - 1][c + 1] &&
                     s > T[r + 1][c - 1] && s > T[r + 1][c] && s > T[r + 1][c + 1];
            }
            if (!EMIT) {
                n_all += __popc(__ballot_sync(0xffffffffu, mx));
                n_ini += __popc(__ballot_sync(0xffffffffu, mx && s >= P.ini_th));
            } else {
                const bool pass = mx && (!use_ini || s >= P.ini_th);
                const unsigned m = __ballot_sync(0xffffffffu, pass);
                if (pass) {
                    const int pos = base + n_all + __popc(m & ((1u << lane) - 1));
                    if (pos < L.det_cap) det[pos] = (uint32_t)(vx0 + c - 1 - OS2_MINB) | ((uint32_t)(vy0 + r - 1 - OS2_MINB) << 12) | ((uint32_t)s << 24);
                }
                n_all += __popc(m);
            }
        }
    if (lane == 0) {
        if (!EMIT) cc[cell] = n_ini ? (n_ini | (1 << 30)) : n_all;
        else if (ci == L.ncells - 1) {
            int tot = base + n_all;
            if (tot > L.det_cap) { atomicOr(&P.status[f], OS2_ST_DET_OVERFLOW); tot = L.det_cap; }
            P.counts[afv_cnt_idx(f, AFV_CNT_DET, l)] = tot;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// DistributeOctTree over (minBorderX, maxBorderX, minBorderY, maxBorderY) (src/ORBextractor.cc:533-534): keys are relative to the
// border corner; per node the FIRST key in push order with the largest response survives (:444-455).
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t os2_f2ord(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__global__ void __launch_bounds__(256) k_os2_octree(const __grid_constant__ Os2Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OctWork W;
    oct_carve(smem_raw, P.oct_ncap, W);
    const int l = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const Os2Level& L = P.lv[l];
    const int M = min(P.counts[afv_cnt_idx(f, AFV_CNT_DET, l)], L.det_cap);
    const uint32_t* det = L.det + (long long)f * L.det_cap;
    float* kx = L.kx + (long long)f * L.det_cap; float* ky = L.ky + (long long)f * L.det_cap;
    unsigned short* knode = L.knode + (long long)f * L.det_cap; unsigned char* kquad = L.kquad + (long long)f * L.det_cap;
    uint2* keep = L.keep + (long long)f * L.keep_cap;
    for (int k = tid; k < M; k += 256) { const uint32_t d = det[k]; kx[k] = (float)(d & 0xfff); ky[k] = (float)((d >> 12) & 0xfff); }
    __syncthreads();
    bool overflow = false;
    int size = oct_distribute(W, kx, ky, knode, kquad, M, L.q, L.n_ini, L.hX, L.octH, P.oct_ncap, tid, overflow);
    if (overflow && tid == 0) atomicOr(&P.status[f], OS2_ST_OCTREE_OVERFLOW);
    unsigned long long* best = W.best;
    for (int p = tid; p < size; p += 256) best[p] = 0ull;
    __syncthreads();
    for (int k = tid; k < M; k += 256)
        atomicMax(&best[knode[k]], ((unsigned long long)(det[k] >> 24) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)k));
    __syncthreads();
    for (int p = tid; p < size; p += 256) {
        const uint32_t k = 0xffffffffu - (uint32_t)(best[p] & 0xffffffffu);
        const uint32_t d = det[k];
        if (p < L.keep_cap) keep[p] = make_uint2(((d & 0xfff) + OS2_MINB) | ((((d >> 12) & 0xfff) + OS2_MINB) << 12), __float_as_uint((float)(d >> 24)));
    }
    if (tid == 0) {
        if (size > L.keep_cap) { atomicOr(&P.status[f], OS2_ST_OCTREE_OVERFLOW); size = L.keep_cap; }
        P.counts[afv_cnt_idx(f, AFV_CNT_KEEP, l)] = size;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) for CV_8U: fixed-point kernel {18, 34, 48, 56, 48, 34, 18} / 256, rows in 8.8,
// columns in 16.16, (v + 2^15) >> 16.
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int os2_refl101(int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * n - 2 - i; return min(max(i, 0), n - 1); }
#define BL_W 64
#define BL_H 32
#define BL_P 72                        // staged bytes per row: columns tx0 - 4 .. tx0 + 67 (word aligned)
__global__ void __launch_bounds__(256) k_os2_blur(const __grid_constant__ Os2Params P, int l) {
    __shared__ __align__(16) uint8_t in[BL_H + 6][BL_P];
    __shared__ __align__(16) uint16_t mid[BL_H + 6][BL_W];
    const Os2Level& L = P.lv[l];
    const int f = blockIdx.z, tid = threadIdx.x;
    const int tx0 = blockIdx.x * BL_W, ty0 = blockIdx.y * BL_H;
    const uint8_t* img = L.img + (long long)f * L.img_fstride;
    // stage rows ty0 - 3 .. ty0 + BL_H + 2, columns tx0 - 4 .. tx0 + 67: interior tiles of a 4-byte aligned image with aligned 32-bit
    // loads, everything else byte by byte through BORDER_REFLECT_101
    const bool aligned = (((unsigned long long)img | (unsigned long long)L.img_stride) & 3ull) == 0;
    if (aligned && tx0 >= 4 && ty0 >= 3 && tx0 + BL_W + 4 <= L.w && ty0 + BL_H + 3 <= L.h) {
        for (int i = tid; i < (BL_H + 6) * (BL_P / 4); i += 256) {
            const int r = i / (BL_P / 4), q = i - r * (BL_P / 4);
            reinterpret_cast<uint32_t*>(&in[r][0])[q] =
                *reinterpret_cast<const uint32_t*>(img + (long long)(ty0 - 3 + r) * L.img_stride + tx0 - 4 + 4 * q);
        }
    } else {
        for (int i = tid; i < (BL_H + 6) * BL_P; i += 256) {
            const int r = i / BL_P, c = i - r * BL_P;
            in[r][c] = img[(long long)os2_refl101(ty0 - 3 + r, L.h) * L.img_stride + os2_refl101(tx0 - 4 + c, L.w)];
        }
    }
    __syncthreads();
    // rows: 4 outputs per thread from 3 words (staged bytes 4q + 1 .. 4q + 10), symmetric taps, 8.8 fixed point (<= 65280)
    for (int i = tid; i < (BL_H + 6) * (BL_W / 4); i += 256) {
        const int r = i / (BL_W / 4), q = i - r * (BL_W / 4);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&in[r][0]) + q;
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
        int p[10];
        p[0] = (w0 >> 8) & 0xff; p[1] = (w0 >> 16) & 0xff; p[2] = w0 >> 24;
        p[3] = w1 & 0xff; p[4] = (w1 >> 8) & 0xff; p[5] = (w1 >> 16) & 0xff; p[6] = w1 >> 24;
        p[7] = w2 & 0xff; p[8] = (w2 >> 8) & 0xff; p[9] = (w2 >> 16) & 0xff;
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = 18 * (p[k] + p[k + 6]) + 34 * (p[k + 1] + p[k + 5]) + 48 * (p[k + 2] + p[k + 4]) + 56 * p[k + 3];
        *reinterpret_cast<uint2*>(&mid[r][4 * q]) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
    }
    __syncthreads();
    // columns: 4 outputs per thread and row, 16.16 fixed point, (v + 2^15) >> 16
    uint8_t* out = L.blur + (long long)f * L.fstride;
    for (int i = tid; i < BL_H * (BL_W / 4); i += 256) {
        const int r = i / (BL_W / 4), q = i - r * (BL_W / 4);
        const int gx = tx0 + 4 * q, gy = ty0 + r;
        if (gx >= L.w || gy >= L.h) continue;
        uint32_t acc[4] = {0, 0, 0, 0};
        constexpr int KW[7] = {18, 34, 48, 56, 48, 34, 18};
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const uint2 m = *reinterpret_cast<const uint2*>(&mid[r + j][4 * q]);
            acc[0] += KW[j] * (m.x & 0xffffu); acc[1] += KW[j] * (m.x >> 16); acc[2] += KW[j] * (m.y & 0xffffu); acc[3] += KW[j] * (m.y >> 16);
        }
        uint32_t pk = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) pk |= ((acc[k] + 32768u) >> 16) << (8 * k);
        *reinterpret_cast<uint32_t*>(out + (long long)gy * L.stride + gx) = pk;          // rows padded to 128 B
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// warp per kept keypoint: IC_Angle, computeOrbDescriptor, merged output rows (levels ascending, octree list order inside a level)
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float os2_fast_atan2_deg(float y, float x) {       // scalar cv::fastAtan2 (see afv_orb.cu)
    const float p1 = 0x1.ca44dep+5f, p3 = -0x1.2aaddcp+4f, p5 = 0x1.1d3f7ep+3f, p7 = -0x1.4515b2p+1f;
    const float eps = 0x1p-52f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

__global__ void __launch_bounds__(256) k_os2_describe(const __grid_constant__ Os2Params P, afv_keypoint* __restrict__ kps,
                                                      uint8_t* __restrict__ desc, float* __restrict__ kpsize, int* __restrict__ n_out) {
    __shared__ uint32_t patw[8][32];
    __shared__ int lvl_start[AFV_MAX_LEVELS + 1];
    const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    { const int pl = tid & 31, pk = tid >> 5; patw[pk][pl] = *reinterpret_cast<const uint32_t*>(&c_os2_pattern[pl * 32 + pk * 4]); }
    if (tid == 0) {
        int acc = 0;
        for (int l = 0; l < P.nlevels; ++l) { lvl_start[l] = acc; acc += min(P.counts[afv_cnt_idx(f, AFV_CNT_KEEP, l)], P.lv[l].keep_cap); }
        lvl_start[P.nlevels] = acc;
        if (blockIdx.x == 0) {
            if (acc > P.out_cap) atomicOr(&P.status[f], OS2_ST_OUT_OVERFLOW);
            n_out[f] = min(acc, P.out_cap);
        }
    }
    __syncthreads();
    const int total = min(lvl_start[P.nlevels], P.out_cap);
    const int i = blockIdx.x * 8 + warp;
    if (i >= total) return;
    int l = 0;
    while (i >= lvl_start[l + 1]) ++l;
    const Os2Level& L = P.lv[l];
    const uint2 kd = (L.keep + (long long)f * L.keep_cap)[i - lvl_start[l]];
    const int x0 = kd.x & 0xfff, y0 = (kd.x >> 12) & 0xfff;              // >= 19 px inside the level: no border handling anywhere
    const uint8_t* img = L.img + (long long)f * L.img_fstride;
    const uint8_t* blr = L.blur + (long long)f * L.fstride;
    int m10 = 0, m01 = 0;
    {
        constexpr int UM[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
        const int u = lane - 15, au = u < 0 ? -u : u;
        const uint8_t* base = img + (long long)(y0 - 15) * L.img_stride + x0 + u;
        int vals[31];
#pragma unroll
        for (int r = 0; r < 31; ++r) vals[r] = lane < 31 ? base[(long long)r * L.img_stride] : 0;
        int colsum = 0;
#pragma unroll
        for (int r = 0; r < 31; ++r) {
            const int v = r - 15, av = v < 0 ? -v : v;
            const int val = (au <= UM[av]) ? vals[r] : 0;
            colsum += val; m01 += v * val;
        }
        m10 = u * colsum;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
    const float angle = os2_fast_atan2_deg((float)m01, (float)m10);
    const float ang = __fmul_rn(angle, 0x1.1df46ap-6f);                  // factorPI = (float)(CV_PI/180.f)
    const float a = (float)cos((double)ang), b = (float)sin((double)ang);   // correctly rounded; the reference's cosf / sinf agree except
                                                                          // on rare 1-ulp cases that would still have to flip a cvRound
    const uint8_t* center = blr + (long long)y0 * L.stride + x0;
    uint32_t byte = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t pw = patw[k][lane];
        int tv[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float px = (float)(int)(int8_t)(pw >> (16 * j)), py = (float)(int)(int8_t)(pw >> (16 * j + 8));
            const int iy = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
            const int ix = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
            tv[j] = center[(long long)iy * L.stride + ix];
        }
        byte |= (uint32_t)(tv[0] < tv[1]) << k;
    }
    const long long o = (long long)f * P.out_cap + i;
    desc[o * 32 + lane] = (uint8_t)byte;
    if (lane == 0) {
        afv_keypoint kp;
        kp.x = l ? __fmul_rn((float)x0, L.sf) : (float)x0; kp.y = l ? __fmul_rn((float)y0, L.sf) : (float)y0;
        kp.size = L.sf; kp.angle = angle;
        kp.response = __uint_as_float(kd.y); kp.octave = l; kp.class_id = -1;
        kps[o] = kp;
        if (kpsize) kpsize[o] = L.size_norm;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------
struct AfvOs2 {
    int nfeatures, nlevels, max_batch, max_w, max_h, device, ini_th, min_th;
    float scale_factor;
    std::vector<void*> allocs;
    uint8_t* img[AFV_MAX_LEVELS]; uint8_t* blur[AFV_MAX_LEVELS]; uint8_t* score[AFV_MAX_LEVELS];
    uint32_t* det[AFV_MAX_LEVELS]; float* kx[AFV_MAX_LEVELS]; float* ky[AFV_MAX_LEVELS];
    unsigned short* knode[AFV_MAX_LEVELS]; unsigned char* kquad[AFV_MAX_LEVELS]; uint2* keep[AFV_MAX_LEVELS]; uint2* tab[AFV_MAX_LEVELS];
    int det_cap[AFV_MAX_LEVELS], keep_cap[AFV_MAX_LEVELS], max_stride[AFV_MAX_LEVELS], max_lh[AFV_MAX_LEVELS], q[AFV_MAX_LEVELS];
    float sf[AFV_MAX_LEVELS], size_norm[AFV_MAX_LEVELS];
    int* cellcnt; uint32_t* cellmask; int cells_cap;
    uint8_t* gray_stage;
    int* h_status; int* h_counts;
    int cur_w, cur_h, last_B;
    size_t oct_smem;
    Os2Params P;
};

template <typename T>
static int os2_alloc(AfvOs2* s, T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 256);
    if (e != cudaSuccess) { afv_set_error("orbslam2: cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e)); return AFV_ERR_CUDA; }
    s->allocs.push_back(q);
    *p = (T*)q;
    return AFV_OK;
}

// ComputePyramid geometry (src/ORBextractor.cc:84-98, :651-652)
static void os2_geometry(int w, int h, int nlevels, float scale_factor, int* lw, int* lh, float* sf) {
    sf[0] = 1.0f;
    for (int l = 1; l < nlevels; ++l) sf[l] = sf[l - 1] * scale_factor;
    for (int l = 0; l < nlevels; ++l) {
        const float inv = 1.0f / sf[l];
        lw[l] = (int)lrintf((float)w * inv);
        lh[l] = (int)lrintf((float)h * inv);
    }
}
struct Os2Cells { int ncols, nrows, wcell, hcell; };
static bool os2_cells(int lw, int lh, Os2Cells* c) {                      // src/ORBextractor.cc:466-484
    const float width = (float)(lw - 2 * OS2_MINB), height = (float)(lh - 2 * OS2_MINB);
    c->ncols = (int)(width / 30.f); c->nrows = (int)(height / 30.f);
    if (c->ncols < 1 || c->nrows < 1) return false;
    c->wcell = (int)ceilf(width / (float)c->ncols); c->hcell = (int)ceilf(height / (float)c->nrows);
    return c->wcell <= CL_P - 4 && c->hcell <= CL_R - 2;
}
// cv::resize INTER_LINEAR coefficient tables (see oracle/afv_oracle_orbslam2.c)
static void os2_lin_tab(int ssize, int dsize, bool is_x, uint2* t) {
    const double inv_scale = (double)dsize / (double)ssize, scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        float f = (float)(((double)d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (is_x) { if (s < 0) { f = 0.f; s = 0; } if (s >= ssize - 1) { f = 0.f; s = ssize - 1; } }
        const int a0 = (int)lrintf((1.f - f) * 2048.f), a1 = (int)lrintf(f * 2048.f);
        t[d] = make_uint2((uint32_t)s, (uint32_t)(a0 & 0xffff) | ((uint32_t)(a1 & 0xffff) << 16));
    }
}

int afv_os2_create(AfvOs2** out, int nfeatures, int nlevels, float scale_factor, float detect_th, int max_batch, int max_w, int max_h) {
    *out = nullptr;
    const int ini_th = (int)detect_th, min_th = 7;       // FeatureExtractorSettings: iniThFAST (= detectTh) / minThFAST (src/FeatureExtractor.cpp:40-46)
    if (ini_th < min_th || ini_th > 254) { afv_set_error("orbslam2: iniThFAST %d outside %d..254", ini_th, min_th); return AFV_ERR_INVALID; }
    if (max_w > 4095 + 2 * OS2_MINB || max_h > 4095 + 2 * OS2_MINB) { afv_set_error("orbslam2: frame dimension too large for the packed key format"); return AFV_ERR_INVALID; }
    AfvOs2* s = new AfvOs2();
    s->nfeatures = nfeatures; s->nlevels = nlevels; s->max_batch = max_batch; s->max_w = max_w; s->max_h = max_h;
    s->scale_factor = scale_factor; s->ini_th = ini_th; s->min_th = min_th; s->cur_w = s->cur_h = 0; s->last_B = 0;
    s->h_status = nullptr; s->h_counts = nullptr; s->gray_stage = nullptr; s->cellcnt = nullptr;
    memset(&s->P, 0, sizeof(s->P));
    cudaGetDevice(&s->device);
    int lw[AFV_MAX_LEVELS], lh[AFV_MAX_LEVELS];
    os2_geometry(max_w, max_h, nlevels, scale_factor, lw, lh, s->sf);
    {   // mnFeaturesPerLevel (src/ORBextractor.cc:102-113)
        float factor = 1.0f / scale_factor;
        float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nlevels - 1; ++l) { s->q[l] = (int)lrintf(nDesired); sum += s->q[l]; nDesired *= factor; }
        s->q[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    }
    {   // keyPt.size = mvScaleFactor[octave] (:629-636 also raise settings->maxKeyPtSize / lower minKeyPtSize to the sizes seen: the
        // table is the steady state after every level has produced a keypoint), then computeSize (src/FeatureExtractor.cpp:132-142)
        const float maxSize0 = powf(1.2f, (float)(8 - 1.0));
        float maxSize = maxSize0, minSize = 1.0f;
        for (int l = 0; l < nlevels; ++l) { if (s->sf[l] > maxSize) maxSize = s->sf[l]; if (s->sf[l] < minSize) minSize = s->sf[l]; }
        for (int l = 0; l < nlevels; ++l) {
            const float sz = powf(scale_factor, (float)l);
            float sn = maxSize;
            if (maxSize > minSize) sn = 1.0f + (sz - minSize) * (maxSize0 - 1.0f) / (maxSize - minSize);
            s->size_norm[l] = sn;
        }
    }
    int rc = AFV_OK, maxq = 0;
    const size_t B = (size_t)max_batch;
    s->cells_cap = 0;
    for (int l = 0; l < nlevels && rc == AFV_OK; ++l) {
        s->max_stride[l] = ((lw[l] + 127) & ~127) + 128; s->max_lh[l] = lh[l] + 2;
        const size_t bytes = (size_t)s->max_stride[l] * s->max_lh[l];
        int dc = lw[l] * lh[l] / 16; if (dc < 1024) dc = 1024;
        s->det_cap[l] = dc; s->keep_cap[l] = s->q[l] + 8;
        if (s->q[l] > maxq) maxq = s->q[l];
        s->cells_cap += (lw[l] / 30 + 2) * (lh[l] / 30 + 2);
        s->img[l] = nullptr;
        if (l > 0) rc = os2_alloc(s, &s->img[l], bytes * B);
        if (rc == AFV_OK) rc = os2_alloc(s, &s->blur[l], bytes * B);
        if (rc == AFV_OK) rc = os2_alloc(s, &s->score[l], bytes * B);
        if (rc == AFV_OK) rc = os2_alloc(s, &s->det[l], (size_t)dc * B);
        if (rc == AFV_OK) rc = os2_alloc(s, &s->kx[l], (size_t)dc * B);
        if (rc == AFV_OK) rc = os2_alloc(s, &s->ky[l], (size_t)dc * B);
        if (rc == AFV_OK) rc = os2_alloc(s, &s->knode[l], (size_t)dc * B);
        if (rc == AFV_OK) rc = os2_alloc(s, &s->kquad[l], (size_t)dc * B);
        if (rc == AFV_OK) rc = os2_alloc(s, &s->keep[l], (size_t)s->keep_cap[l] * B);
        if (rc == AFV_OK) rc = os2_alloc(s, &s->tab[l], (size_t)(lw[l] + lh[l] + 8));
    }
    if (rc == AFV_OK) rc = os2_alloc(s, &s->cellcnt, (size_t)s->cells_cap * B);
    if (rc == AFV_OK) rc = os2_alloc(s, &s->cellmask, (size_t)s->cells_cap * 32 * B);
    if (rc == AFV_OK) rc = os2_alloc(s, &s->P.counts, 4 * AFV_MAX_LEVELS * B);
    if (rc == AFV_OK) rc = os2_alloc(s, &s->P.status, B);
    if (rc == AFV_OK) rc = os2_alloc(s, &s->gray_stage, (size_t)max_w * max_h * B);
    if (rc == AFV_OK) {
        cudaError_t e = cudaMallocHost((void**)&s->h_status, sizeof(int) * B);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_counts, sizeof(int) * 4 * AFV_MAX_LEVELS * B);
        if (e != cudaSuccess) { afv_set_error("orbslam2: cudaMallocHost failed: %s", cudaGetErrorString(e)); rc = AFV_ERR_CUDA; }
    }
    if (rc == AFV_OK) {
        s->P.oct_ncap = maxq + 16;
        s->oct_smem = oct_work_bytes(s->P.oct_ncap);
        if (s->oct_smem > 227 * 1024) { afv_set_error("orbslam2: nfeatures too large for the octree workspace"); rc = AFV_ERR_INVALID; }
        else if (s->oct_smem > 40 * 1024) {
            // per-device attribute: only ever raised (an earlier, larger extractor keeps working); static shared memory of the kernel
            // comes on top, so the limit asked for is what this extractor needs plus headroom, not the 227 KB maximum
            const cudaError_t e = cudaFuncSetAttribute(k_os2_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(s->oct_smem + 1024 > 200 * 1024 ? s->oct_smem : 200 * 1024));
            if (e != cudaSuccess) { afv_set_error("orbslam2: cudaFuncSetAttribute(%zu) failed: %s", s->oct_smem, cudaGetErrorString(e)); rc = AFV_ERR_CUDA; }
        }
    }
    if (rc != AFV_OK) { afv_os2_destroy(s); return rc; }
    *out = s;
    return AFV_OK;
}

void afv_os2_destroy(AfvOs2* s) {
    if (!s) return;
    for (void* p : s->allocs) cudaFree(p);
    if (s->h_status) cudaFreeHost(s->h_status);
    if (s->h_counts) cudaFreeHost(s->h_counts);
    delete s;
}

uint8_t* afv_os2_stage(AfvOs2* s) { return s->gray_stage; }

static int os2_configure(AfvOs2* s, int w, int h) {
    if (w == s->cur_w && h == s->cur_h) return AFV_OK;
    if (w > s->max_w || h > s->max_h) { afv_set_error("orbslam2: frame %dx%d larger than the extractor's configured maximum %dx%d", w, h, s->max_w, s->max_h); return AFV_ERR_INVALID; }
    int lw[AFV_MAX_LEVELS], lh[AFV_MAX_LEVELS]; float sf[AFV_MAX_LEVELS];
    os2_geometry(w, h, s->nlevels, s->scale_factor, lw, lh, sf);
    Os2Params& P = s->P;
    int* counts = P.counts; int* status = P.status;
    memset(&P, 0, sizeof(P));
    P.counts = counts; P.status = status; P.cellcnt = s->cellcnt; P.cellmask = s->cellmask;
    P.nlevels = s->nlevels; P.ini_th = s->ini_th; P.min_th = s->min_th;
    P.oct_ncap = (int)0;
    int maxq = 0, cells = 0;
    for (int l = 0; l < s->nlevels; ++l) {
        Os2Level& L = P.lv[l];
        Os2Cells c;
        if (!os2_cells(lw[l], lh[l], &c)) {
            afv_set_error("orbslam2: level %d (%dx%d) of a %dx%d frame is too small for the 30-pixel cell grid (the reference divides by zero)", l, lw[l], lh[l], w, h);
            return AFV_ERR_INVALID;
        }
        L.w = lw[l]; L.h = lh[l]; L.stride = (lw[l] + 127) & ~127;
        if (L.stride > s->max_stride[l] || lh[l] > s->max_lh[l]) { afv_set_error("orbslam2: internal: level %d larger than its arena", l); return AFV_ERR_INVALID; }
        L.fstride = (long long)L.stride * lh[l];
        L.img = s->img[l]; L.img_stride = L.stride; L.img_fstride = L.fstride;
        L.blur = s->blur[l]; L.score = s->score[l];
        L.ncols = c.ncols; L.nrows = c.nrows; L.wcell = c.wcell; L.hcell = c.hcell; L.cell_base = cells; L.ncells = c.ncols * c.nrows;
        cells += L.ncells;
        L.q = s->q[l]; if (L.q > maxq) maxq = L.q;
        const int bw = lw[l] - 2 * OS2_MINB, bh = lh[l] - 2 * OS2_MINB;
        L.n_ini = (int)round((double)((float)bw / (float)bh));                        // src/ORBextractor.cc:243
        if (L.n_ini < 1) { afv_set_error("orbslam2: portrait level with w/h < 0.5 is not supported (the reference divides by zero)"); return AFV_ERR_INVALID; }
        L.hX = (float)bw / (float)L.n_ini; L.octH = bh;
        L.det_cap = s->det_cap[l]; L.det = s->det[l]; L.kx = s->kx[l]; L.ky = s->ky[l]; L.knode = s->knode[l]; L.kquad = s->kquad[l];
        L.keep_cap = s->keep_cap[l]; L.keep = s->keep[l];
        L.sf = s->sf[l]; L.size_norm = s->size_norm[l];
        if (l > 0) {
            std::vector<uint2> t((size_t)lw[l] + lh[l]);
            os2_lin_tab(lw[l - 1], lw[l], true, t.data());
            os2_lin_tab(lh[l - 1], lh[l], false, t.data() + lw[l]);
            AFV_CUDA_CHECK(cudaMemcpy(s->tab[l], t.data(), t.size() * sizeof(uint2), cudaMemcpyHostToDevice));
            L.xtab = s->tab[l]; L.ytab = s->tab[l] + lw[l];
        }
    }
    if (cells > s->cells_cap) { afv_set_error("orbslam2: internal: cell table too small"); return AFV_ERR_INVALID; }
    P.ncells_total = cells; P.oct_ncap = maxq + 16;
    s->cur_w = w; s->cur_h = h;
    return AFV_OK;
}

int afv_os2_run(AfvOs2* s, const uint8_t* d_gray, int B, int w, int h, int stride, long frame_stride, afv_keypoint* d_kps,
                uint8_t* d_desc, float* d_kpsize, int cap, int* d_n_out, cudaStream_t st) {
    int rc = os2_configure(s, w, h);
    if (rc) return rc;
    Os2Params P = s->P;
    P.B = B; P.out_cap = cap;
    P.lv[0].img = d_gray; P.lv[0].img_stride = stride; P.lv[0].img_fstride = frame_stride;
    AFV_CUDA_CHECK(cudaMemsetAsync(P.counts, 0, sizeof(int) * 4 * AFV_MAX_LEVELS * B, st));
    AFV_CUDA_CHECK(cudaMemsetAsync(P.status, 0, sizeof(int) * B, st));
    for (int l = 1; l < P.nlevels; ++l) {
        AfvProfScope ps("k_os2_resize", st);
        k_os2_resize<<<dim3((P.lv[l].w + 127) / 128, (P.lv[l].h + 7) / 8, B), 256, 0, st>>>(P, l); ++g_afv_launches;
    }
    for (int l = 0; l < P.nlevels; ++l) {
        const int rw = P.lv[l].w - 38, rh = P.lv[l].h - 38;                       // detection rectangle [19, w - 19) x [19, h - 19)
        AfvProfScope ps("k_os2_score", st);
        k_os2_score<<<dim3((rw + SC_W - 1) / SC_W, (rh + SC_H - 1) / SC_H, B), 256, 0, st>>>(P, l); ++g_afv_launches;
    }
    { AfvProfScope ps("k_os2_cells", st);
      k_os2_cells<false><<<dim3((P.ncells_total + 7) / 8, B), 256, 0, st>>>(P); ++g_afv_launches;
      k_os2_cells<true><<<dim3((P.ncells_total + 7) / 8, B), 256, 0, st>>>(P); ++g_afv_launches; }
    { AfvProfScope ps("k_os2_octree", st);
      k_os2_octree<<<dim3(P.nlevels, B), 256, s->oct_smem, st>>>(P); ++g_afv_launches; }
    for (int l = 0; l < P.nlevels; ++l) {
        AfvProfScope ps("k_os2_blur", st);
        k_os2_blur<<<dim3((P.lv[l].w + BL_W - 1) / BL_W, (P.lv[l].h + BL_H - 1) / BL_H, B), 256, 0, st>>>(P, l); ++g_afv_launches;
    }
    { AfvProfScope ps("k_os2_describe", st);
      k_os2_describe<<<dim3((cap + 7) / 8, B), 256, 0, st>>>(P, d_kps, d_desc, d_kpsize, d_n_out); ++g_afv_launches; }
    AFV_CUDA_CHECK(cudaGetLastError());
    s->last_B = B;
    return AFV_OK;
}

int afv_os2_status(AfvOs2* s, int B, cudaStream_t st) {
    AFV_CUDA_CHECK(cudaMemcpyAsync(s->h_status, s->P.status, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    AFV_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b)
        if (s->h_status[b]) {
            afv_set_error("orbslam2: capacity exceeded in frame %d (flags 0x%x: 2 detect list, 4 caller cap, 8 octree)", b, s->h_status[b]);
            return AFV_ERR_CAPACITY;
        }
    return AFV_OK;
}

// what = 40 level image (level >= 1), 41 blurred level, 42 score map, 43 detect list in push order (uint32 (x-16) | (y-16) << 12 |
// score << 24), 44 octree keep list (8 bytes: x | y << 12 in level coordinates, float response)
int afv_os2_debug_read(AfvOs2* s, int what, int frame, int level, void* out, long cap_bytes, long* n_bytes) {
    if (level < 0 || level >= s->nlevels || frame < 0 || frame >= s->last_B) { afv_set_error("orbslam2: afv_debug_read: bad frame / level"); return AFV_ERR_INVALID; }
    const Os2Level& L = s->P.lv[level];
    if (what == 40 || what == 41 || what == 42) {
        const long need = (long)L.w * L.h;
        if (cap_bytes < need) { afv_set_error("buffer too small"); return AFV_ERR_INVALID; }
        if (what == 40 && level == 0) { afv_set_error("level 0 is the caller's input"); return AFV_ERR_INVALID; }
        const uint8_t* src = (what == 40 ? s->img[level] : what == 41 ? s->blur[level] : s->score[level]) + (long long)frame * L.fstride;
        AFV_CUDA_CHECK(cudaMemcpy2D(out, L.w, src, L.stride, L.w, L.h, cudaMemcpyDeviceToHost));
        *n_bytes = need;
        return AFV_OK;
    }
    if (what != 43 && what != 44) { afv_set_error("orbslam2: unknown tap %d", what); return AFV_ERR_INVALID; }
    AFV_CUDA_CHECK(cudaMemcpy(s->h_counts, s->P.counts, sizeof(int) * 4 * AFV_MAX_LEVELS * s->last_B, cudaMemcpyDeviceToHost));
    int n = s->h_counts[afv_cnt_idx(frame, what == 43 ? AFV_CNT_DET : AFV_CNT_KEEP, level)];
    const int capn = what == 43 ? L.det_cap : L.keep_cap;
    if (n > capn) n = capn;
    const long esz = what == 43 ? 4 : 8;
    if (cap_bytes < n * esz) { afv_set_error("buffer too small"); return AFV_ERR_INVALID; }
    const void* src = what == 43 ? (const void*)(L.det + (long long)frame * L.det_cap) : (const void*)(L.keep + (long long)frame * L.keep_cap);
    AFV_CUDA_CHECK(cudaMemcpy(out, src, n * esz, cudaMemcpyDeviceToHost));
    *n_bytes = n * esz;
    return AFV_OK;
}

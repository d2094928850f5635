// afv_match.cu -- hand-written sm_100a kernels + C ABI of the FeatureMatcher path.
//
// Reference: src/FeatureMatcher.cc (SearchForInitialization :399-557, SearchByBoW :186-283, window search of
// SearchByProjection :73-154, DescriptorDistance :1508-1531, rotation histogram :1579-1668) and the Frame grid
// (src/Frame.cc:225-240, :333-394).  The reference walks Frame/KeyFrame/MapPoint graphs one candidate at a
// time; here the same decisions are taken on flat arrays:
//   * distances are warp/CTA-parallel popcounts (binary descriptors) or float-diff / double-accumulate L2^2;
//   * "first minimum wins" of the sequential loops is an explicit order key carried through the reductions:
//     (distance, enumeration position) compared lexicographically, top-2 of that order = (best, second);
//   * loops whose iterations depend on earlier ones ("already matched", match stealing) stay sequential over
//     queries inside one CTA per frame pair, with all candidate work of a query done in parallel.
#include "afv_common.cuh"
#include <algorithm>
#include <atomic>
#include <float.h>

#define NCELLS (AFV_GRID_COLS * AFV_GRID_ROWS)

static inline cudaStream_t as_stream(void* s) { return (cudaStream_t)s; }

__host__ __device__ inline int desc_bytes(int desc_type) {
    return desc_type == AFV_FEAT_ORB32 ? 32 : desc_type == AFV_FEAT_AKAZE61 ? 61 : desc_type == AFV_FEAT_BRISK48 ? 48
         : desc_type == AFV_FEAT_SIFT128 ? 512 : -1;
}

// ---- distances ----------------------------------------------------------------------------------------
// Binary: popcount of XOR over D bytes (== the reference's bit-hack for orb32, cv::norm(NORM_HAMMING) for
// akaze61/brisk48).  `a`/`b` may be unaligned (61-byte rows): word path only when both are 4-byte aligned.
__device__ __forceinline__ int hamming_bytes(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int D) {
    int d = 0;
    if ((((uintptr_t)a | (uintptr_t)b) & 3) == 0) {
        const uint32_t* x = reinterpret_cast<const uint32_t*>(a);
        const uint32_t* y = reinterpret_cast<const uint32_t*>(b);
        const int nw = D >> 2;
        for (int i = 0; i < nw; ++i) d += __popc(x[i] ^ y[i]);
        for (int i = nw * 4; i < D; ++i) d += __popc((uint32_t)(a[i] ^ b[i]));
    } else {
        for (int i = 0; i < D; ++i) d += __popc((uint32_t)(a[i] ^ b[i]));
    }
    return d;
}
// cv::norm(a,b,NORM_L2SQR) on CV_32F (normL2Sqr<float,double>): float differences, double accumulation in
// groups of four, result narrowed to float (Descriptor_Distance_Type).
__device__ __forceinline__ float l2sqr128(const float* __restrict__ a, const float* __restrict__ b) {
    double s = 0.0;
    for (int i = 0; i < 128; i += 4) {
        const double v0 = (double)(a[i] - b[i]), v1 = (double)(a[i + 1] - b[i + 1]);
        const double v2 = (double)(a[i + 2] - b[i + 2]), v3 = (double)(a[i + 3] - b[i + 3]);
        s += v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3;
    }
    return (float)s;
}
__device__ __forceinline__ float desc_distance(int desc_type, const void* a, const void* b, int D) {
    if (desc_type == AFV_FEAT_SIFT128) return l2sqr128((const float*)a, (const float*)b);
    return (float)hamming_bytes((const uint8_t*)a, (const uint8_t*)b, D);
}

// ---- (distance, order) top-2 --------------------------------------------------------------------------
// key = distance bits (non-negative float => monotone as uint) << 32 | enumeration order.  Smaller = better.
struct Top2 { unsigned long long k1, k2; };
#define KEY_NONE 0xffffffffffffffffull
__device__ __forceinline__ void top2_push(Top2& t, unsigned long long k) {
    if (k < t.k1) { t.k2 = t.k1; t.k1 = k; } else if (k < t.k2) t.k2 = k;
}
__device__ __forceinline__ void top2_warp_reduce(Top2& t) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, t.k1, o);
        const unsigned long long b = __shfl_xor_sync(0xffffffffu, t.k2, o);
        top2_push(t, a); top2_push(t, b);
    }
}
__device__ __forceinline__ unsigned long long make_key(float dist, uint32_t order) {
    return ((unsigned long long)__float_as_uint(dist) << 32) | order;
}
__device__ __forceinline__ float key_dist(unsigned long long k) {
    return k == KEY_NONE ? FLT_MAX : __uint_as_float((uint32_t)(k >> 32));
}

// ---- Frame grid ---------------------------------------------------------------------------------------
// Frame::PosInGrid (src/Frame.cc:384-394): round() of the scaled coordinate, dropped when outside 64x48.
__device__ __forceinline__ int grid_cell(float x, float y, float minX, float minY, float invW, float invH) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(x, minX), invW));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(y, minY), invH));
    if (px < 0 || px >= AFV_GRID_COLS || py < 0 || py >= AFV_GRID_ROWS) return -1;
    return px * AFV_GRID_ROWS + py;
}
// Frame::GetFeaturesInArea cell range (src/Frame.cc:339-353), trimmed.  Returns false when the window misses the grid.
// The reference's floor / ceil range is up to two cells wider per axis than the cells that can hold a point passing its own
// |dx| < r, |dy| < r gate, because PosInGrid rounds to the NEAREST cell.  Every search here applies that gate, so the outer cells
// contribute nothing: dropping them leaves the enumeration order of the surviving candidates, hence every result, unchanged and
// removes 40 % of the candidates walked at r = 15 (10 % at r = 100).  The 0.01-cell margin (0.1 px) is three orders above the
// rounding of the float gate and of the cell computation, both monotone in the coordinate.
__device__ __forceinline__ bool window_cells(float x, float y, float r, float minX, float minY, float invW, float invH,
                                             int& c0, int& c1, int& r0, int& r1) {
    c0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, minX), r), invW)));
    if (c0 >= AFV_GRID_COLS) return false;
    c1 = min(AFV_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, minX), r), invW)));
    if (c1 < 0) return false;
    r0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, minY), r), invH)));
    if (r0 >= AFV_GRID_ROWS) return false;
    r1 = min(AFV_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, minY), r), invH)));
    if (r1 < 0) return false;
    const float gx = __fmul_rn(__fsub_rn(x, minX), invW), gy = __fmul_rn(__fsub_rn(y, minY), invH), rx = r * invW, ry = r * invH;
    c0 = max(c0, (int)floorf(gx - rx + 0.49f)); c1 = min(c1, (int)floorf(gx + rx + 0.51f));
    r0 = max(r0, (int)floorf(gy - ry + 0.49f)); r1 = min(r1, (int)floorf(gy + ry + 0.51f));
    return c0 <= c1 && r0 <= r1;
}

// One CTA per frame: CSR grid with items ascending inside each cell (== push_back order of the reference).
__global__ void __launch_bounds__(256) k_grid_build(const afv_keypoint* __restrict__ kps, const int* __restrict__ n_arr,
                                                    int cap, float minX, float minY, float invW, float invH,
                                                    int* __restrict__ cell_start, int* __restrict__ cell_items) {
    __shared__ int cnt[NCELLS + 1];
    __shared__ int warp_tot[8];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = min(n_arr[f], cap);
    const afv_keypoint* k = kps + (long long)f * cap;
    int* cs = cell_start + (long long)f * (NCELLS + 1);
    int* ci = cell_items + (long long)f * cap;
    for (int i = tid; i <= NCELLS; i += 256) cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int c = grid_cell(k[i].x, k[i].y, minX, minY, invW, invH);
        if (c >= 0) atomicAdd(&cnt[c], 1);
    }
    __syncthreads();
    // exclusive scan of 3072 counters: 12 per thread
    const int items = (NCELLS + 255) / 256, beg = tid * items, end = min(beg + items, NCELLS);
    int sum = 0;
    for (int i = beg; i < end; ++i) sum += cnt[i];
    const int lane = tid & 31, wid = tid >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < wid; ++w) base += warp_tot[w];
    int run = base + incl - sum;
    for (int i = beg; i < end; ++i) { const int v = cnt[i]; cnt[i] = run; cs[i] = run; run += v; }
    if (tid == 255) cs[NCELLS] = run;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int c = grid_cell(k[i].x, k[i].y, minX, minY, invW, invH);
        if (c >= 0) ci[atomicAdd(&cnt[c], 1)] = i;
    }
    __syncthreads();
    // restore ascending index order inside each cell (cells hold a handful of items)
    for (int c = tid; c < NCELLS; c += 256) {
        const int b = cs[c], e = cnt[c];
        for (int i = b + 1; i < e; ++i) {
            const int v = ci[i];
            int j = i - 1;
            while (j >= b && ci[j] > v) { ci[j + 1] = ci[j]; --j; }
            ci[j + 1] = v;
        }
    }
}

extern "C" int afv_grid_build(const afv_keypoint* d_kps, const int* d_n, int B, int cap,
                              float min_x, float min_y, float max_x, float max_y,
                              int* d_cell_start, int* d_cell_items, void* cuda_stream) {
    if (!d_kps || !d_n || !d_cell_start || !d_cell_items || B < 1) { afv_set_error("afv_grid_build: bad argument"); return AFV_ERR_INVALID; }
    const float invW = (float)AFV_GRID_COLS / (max_x - min_x), invH = (float)AFV_GRID_ROWS / (max_y - min_y);
    k_grid_build<<<B, 256, 0, as_stream(cuda_stream)>>>(d_kps, d_n, cap, min_x, min_y, invW, invH, d_cell_start, d_cell_items);
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

// ---- Frame::UndistortKeyPoints (src/Frame.cc:403-433) -> cv::undistortPoints(mat, mat, mK, mDistCoef, Mat(), mK) ---------
// OpenCV's arithmetic, pinned bit for bit to cv2 4.13.0: double precision, x = (u - cx) / fx as a multiply by 1/fx, FIVE
// fixed-point iterations of the inverse distortion model (k1 k2 p1 p2 k3; the default TermCriteria is count-only), then
// u' = x fx + cx narrowed to float.  Thread per keypoint; every other cv::KeyPoint field is copied (Frame.cc:427-431).
__global__ void __launch_bounds__(256) k_undistort(const afv_keypoint* __restrict__ kps, const int* __restrict__ n_arr, int cap,
                                                   double fx, double fy, double cx, double cy, double k1, double k2, double p1,
                                                   double p2, double k3, int identity, afv_keypoint* __restrict__ out) {
    const int f = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    if (i >= min(n_arr[f], cap)) return;
    afv_keypoint kp = kps[(long long)f * cap + i];
    if (!identity) {
        const double ifx = 1.0 / fx, ify = 1.0 / fy;
        double x = ((double)kp.x - cx) * ifx, y = ((double)kp.y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            const double icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2);
            if (icdist < 0) { x = x0; y = y0; break; }
            const double dX = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
            const double dY = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
            x = (x0 - dX) * icdist; y = (y0 - dY) * icdist;
        }
        kp.x = (float)(x * fx + cx); kp.y = (float)(y * fy + cy);
    }
    out[(long long)f * cap + i] = kp;
}

extern "C" int afv_undistort_keypoints(const afv_keypoint* d_kps, const int* d_n, int B, int cap, const float* K4, const float* dist5,
                                       afv_keypoint* d_kps_un, void* cuda_stream) {
    if (!d_kps || !d_n || !K4 || !dist5 || !d_kps_un || B < 1 || cap < 1) { afv_set_error("afv_undistort_keypoints: bad argument"); return AFV_ERR_INVALID; }
    const int identity = dist5[0] == 0.0f;                          // mDistCoef.at<float>(0) == 0.0 -> mvKeysUn = mvKeys (:405-409)
    k_undistort<<<dim3((cap + 255) / 256, B), 256, 0, as_stream(cuda_stream)>>>(d_kps, d_n, cap, (double)K4[0], (double)K4[1], (double)K4[2],
            (double)K4[3], (double)dist5[0], (double)dist5[1], (double)dist5[2], (double)dist5[3], (double)dist5[4], identity, d_kps_un);
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

// ---- Frame::isInFrustum (src/Frame.cc:276-331) + the window prologue of SearchByProjection (src/FeatureMatcher.cc:86-95) ----
// Thread per map point.  IEEE float32, every operation rounded once (explicit intrinsics; the TU is also built --fmad=false),
// 3-term sums left to right: the arithmetic contract of oracle/afv_oracle_match.c::orc_is_in_frustum (the reference evaluates
// the same expressions through Eigen under -march=native, i.e. to ~1e-6 relative of this, not to the bit).
struct AfvPose { float R[9], t[3], c[3], fx, fy, cx, cy, mbf, minX, maxX, minY, maxY; };
__global__ void __launch_bounds__(256) k_in_frustum(const float* __restrict__ Pw, const float* __restrict__ nrm, const float* __restrict__ mind,
        const float* __restrict__ maxd, const float* __restrict__ rsz, const float* __restrict__ rsg, const float* __restrict__ rds, int M,
        const AfvPose S, float cos_limit, float radius_factor, float size_tol, uint8_t* __restrict__ in_view, float* __restrict__ proj3,
        float* __restrict__ track3, float* __restrict__ qr, float* __restrict__ qmin, float* __restrict__ qmax) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= M) return;
    const float P0 = Pw[3 * i], P1 = Pw[3 * i + 1], P2 = Pw[3 * i + 2];
    float u = 0.f, v = 0.f, ur = 0.f, tsz = 0.f, tsg = 0.f, vc = 0.f, r = -1.0f, lo = 0.f, hi = 0.f;
    bool ok = false;
    float Pc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        Pc[k] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(S.R[3 * k], P0), __fmul_rn(S.R[3 * k + 1], P1)), __fmul_rn(S.R[3 * k + 2], P2)), S.t[k]);
    if (!(Pc[2] < 0.0f)) {
        const float invz = __fdiv_rn(1.0f, Pc[2]);
        const float uu = __fadd_rn(__fmul_rn(__fmul_rn(S.fx, Pc[0]), invz), S.cx);
        const float vv = __fadd_rn(__fmul_rn(__fmul_rn(S.fy, Pc[1]), invz), S.cy);
        if (!(uu < S.minX || uu > S.maxX) && !(vv < S.minY || vv > S.maxY)) {
            const float a = __fsub_rn(P0, S.c[0]), b = __fsub_rn(P1, S.c[1]), c = __fsub_rn(P2, S.c[2]);
            const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
            if (!(dist < mind[i] || dist > maxd[i])) {
                const float dot = __fadd_rn(__fadd_rn(__fmul_rn(a, nrm[3 * i]), __fmul_rn(b, nrm[3 * i + 1])), __fmul_rn(c, nrm[3 * i + 2]));
                const float viewCos = __fdiv_rn(dot, dist);
                if (!(viewCos < cos_limit)) {
                    ok = true; u = uu; v = vv; ur = __fsub_rn(uu, __fmul_rn(S.mbf, invz));
                    tsz = __fdiv_rn(__fmul_rn(rsz[i], rds[i]), dist); tsg = __fdiv_rn(__fmul_rn(rsg[i], rds[i]), dist); vc = viewCos;
                    const float rv = (double)viewCos > 0.998 ? 2.5f : 4.0f;
                    r = __fmul_rn(__fmul_rn(radius_factor, rv), tsz); lo = __fdiv_rn(tsz, size_tol); hi = __fmul_rn(tsz, size_tol);
                }
            }
        }
    }
    in_view[i] = ok ? 1 : 0;
    proj3[3 * i] = u; proj3[3 * i + 1] = v; proj3[3 * i + 2] = ur;
    track3[3 * i] = tsz; track3[3 * i + 1] = tsg; track3[3 * i + 2] = vc;
    if (qr) { qr[i] = r; qmin[i] = lo; qmax[i] = hi; }
}

extern "C" int afv_is_in_frustum(const float* d_Pw, const float* d_normal, const float* d_min_dist, const float* d_max_dist,
                                 const float* d_ref_size, const float* d_ref_sigma, const float* d_ref_dist, int M, const float* pose16,
                                 const float* cam5, const float* bounds4, float viewing_cos_limit, float radius_factor, float size_tolerance,
                                 uint8_t* d_in_view, float* d_proj3, float* d_track3, float* d_qr, float* d_qmin, float* d_qmax,
                                 void* cuda_stream) {
    if (M == 0) return AFV_OK;                                           // empty local map: nothing to do (pointers may be NULL)
    if (!d_Pw || !d_normal || !d_min_dist || !d_max_dist || !d_ref_size || !d_ref_sigma || !d_ref_dist || !pose16 || !cam5 || !bounds4 ||
        !d_in_view || !d_proj3 || !d_track3 || M < 0 || (d_qr && (!d_qmin || !d_qmax))) { afv_set_error("afv_is_in_frustum: bad argument"); return AFV_ERR_INVALID; }
    (void)cudaGetLastError();            // a stale error left by another library in this thread must not be reported as this launch's
    AfvPose S;
    for (int i = 0; i < 9; ++i) S.R[i] = pose16[i];
    for (int i = 0; i < 3; ++i) { S.t[i] = pose16[9 + i]; S.c[i] = pose16[12 + i]; }
    S.fx = cam5[0]; S.fy = cam5[1]; S.cx = cam5[2]; S.cy = cam5[3]; S.mbf = cam5[4];
    S.minX = bounds4[0]; S.maxX = bounds4[1]; S.minY = bounds4[2]; S.maxY = bounds4[3];
    k_in_frustum<<<(M + 255) / 256, 256, 0, as_stream(cuda_stream)>>>(d_Pw, d_normal, d_min_dist, d_max_dist, d_ref_size, d_ref_sigma, d_ref_dist, M, S,
            viewing_cos_limit, radius_factor, size_tolerance, d_in_view, d_proj3, d_track3, d_qr, d_qmin, d_qmax);
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

// ---- packed per-rank results for the multi-GPU gather: n[B] | nmatches[B] | matches12[B][cap] | kps[B][cap] |
// desc[B][cap][D] copied into one contiguous message by ONE launch (the five tensors are 4-byte aligned: 32-bit copies) ------------
struct AfvPackSeg { const uint32_t* src; long long off_w, nw; };            // destination offset / length in 32-bit words
struct AfvPackSegs { AfvPackSeg s[5]; long long total_w; };
__global__ void __launch_bounds__(256) k_pack_results(const AfvPackSegs S, uint32_t* __restrict__ dst) {
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < S.total_w; i += stride) {
        int k = 0;
#pragma unroll
        for (int j = 1; j < 5; ++j) k += i >= S.s[j].off_w;
        dst[i] = S.s[k].src[i - S.s[k].off_w];
    }
}
extern "C" size_t afv_pack_results_bytes(int B, int cap, int desc_bytes) {
    return (size_t)B * 8 + (size_t)B * cap * (4 + 28) + (((size_t)B * cap * desc_bytes + 3) & ~(size_t)3);
}
extern "C" int afv_pack_results(const int* d_n, const int* d_nmatches, const int* d_matches12, const afv_keypoint* d_kps, const void* d_desc,
                                int B, int cap, int desc_bytes, void* d_pack, void* cuda_stream) {
    if (!d_n || !d_nmatches || !d_matches12 || !d_kps || !d_desc || !d_pack || B < 1 || cap < 1 || desc_bytes < 1 || ((uintptr_t)d_desc & 3) || ((uintptr_t)d_pack & 3)) {
        afv_set_error("afv_pack_results: bad argument"); return AFV_ERR_INVALID;
    }
    AfvPackSegs S;
    const long long nw[5] = {B, B, (long long)B * cap, (long long)B * cap * 7, ((long long)B * cap * desc_bytes + 3) / 4};
    const void* src[5] = {d_n, d_nmatches, d_matches12, d_kps, d_desc};
    long long off = 0;
    for (int i = 0; i < 5; ++i) { S.s[i].src = (const uint32_t*)src[i]; S.s[i].off_w = off; S.s[i].nw = nw[i]; off += nw[i]; }
    S.total_w = off;
    (void)cudaGetLastError();
    const int blocks = (int)std::min<long long>((off + 255) / 256, 148 * 16);
    k_pack_results<<<blocks, 256, 0, as_stream(cuda_stream)>>>(S, (uint32_t*)d_pack);
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

// ---- stateless window search: one warp per query ---------------------------------------------------------
__global__ void __launch_bounds__(256) k_match_window(int desc_type, int D, const uint8_t* __restrict__ q,
        const float* __restrict__ qxy, const float* __restrict__ qr, const float* __restrict__ qmin,
        const float* __restrict__ qmax, int nq, const afv_keypoint* __restrict__ tk, const uint8_t* __restrict__ td,
        const float* __restrict__ tsize, const int* __restrict__ cs, const int* __restrict__ ci,
        float minX, float minY, float invW, float invH,
        int* __restrict__ best, float* __restrict__ bestd, float* __restrict__ secondd,
        float* __restrict__ best_size, float* __restrict__ second_size) {
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (qi >= nq) return;
    const float x = qxy[2 * qi], y = qxy[2 * qi + 1], r = qr[qi], smin = qmin[qi], smax = qmax[qi];
    Top2 t; t.k1 = t.k2 = KEY_NONE;
    int c0, c1, r0, r1;
    if (window_cells(x, y, r, minX, minY, invW, invH, c0, c1, r0, r1)) {
        const uint8_t* qd = q + (long long)qi * D;
        uint32_t ebase = 0;
        for (int ix = c0; ix <= c1; ++ix) {
            const int sb = cs[ix * AFV_GRID_ROWS + r0], se = cs[ix * AFV_GRID_ROWS + r1 + 1];
            for (int j = sb + lane; j < se; j += 32) {
                const int idx = ci[j];
                const float sz = tsize[idx];
                if (sz < smin || sz > smax) continue;
                const float dx = __fsub_rn(tk[idx].x, x), dy = __fsub_rn(tk[idx].y, y);
                if (fabsf(dx) < r && fabsf(dy) < r) {
                    const float d = desc_distance(desc_type, qd, td + (long long)idx * D, D);
                    // order = enumeration position; idx is recovered from ci[] through it
                    top2_push(t, make_key(d, ebase + (uint32_t)(j - sb)));
                }
            }
            ebase += (uint32_t)(se - sb);
        }
        top2_warp_reduce(t);
    }
    if (lane == 0) {
        // map enumeration positions back to train indices
        int bi = -1, si = -1;
        if (t.k1 != KEY_NONE || t.k2 != KEY_NONE) {
            uint32_t e1 = (uint32_t)t.k1, e2 = (uint32_t)t.k2, ebase = 0;
            for (int ix = c0; ix <= c1; ++ix) {
                const int sb = cs[ix * AFV_GRID_ROWS + r0], se = cs[ix * AFV_GRID_ROWS + r1 + 1];
                const uint32_t len = (uint32_t)(se - sb);
                if (t.k1 != KEY_NONE && e1 >= ebase && e1 < ebase + len) bi = ci[sb + (e1 - ebase)];
                if (t.k2 != KEY_NONE && e2 >= ebase && e2 < ebase + len) si = ci[sb + (e2 - ebase)];
                ebase += len;
            }
        }
        best[qi] = bi; bestd[qi] = key_dist(t.k1); secondd[qi] = key_dist(t.k2);
        if (best_size) best_size[qi] = bi >= 0 ? tsize[bi] : -1.0f;
        if (second_size) second_size[qi] = si >= 0 ? tsize[si] : -1.0f;
    }
}

extern "C" int afv_match_window(int desc_type, const void* d_q, const float* d_qxy, const float* d_qr,
                                const float* d_qmin_size, const float* d_qmax_size, int nq,
                                const afv_keypoint* d_tk, const void* d_td, const float* d_tsize, int nt,
                                const int* d_cell_start, const int* d_cell_items,
                                float min_x, float min_y, float max_x, float max_y,
                                int* d_best, float* d_bestd, float* d_secondd, float* d_best_size, float* d_second_size,
                                void* cuda_stream) {
    const int D = desc_bytes(desc_type);
    if (D < 0 || !d_q || !d_qxy || !d_qr || !d_qmin_size || !d_qmax_size || !d_tk || !d_td || !d_tsize || !d_cell_start ||
        !d_cell_items || !d_best || !d_bestd || !d_secondd || nq < 0 || nt < 0) { afv_set_error("afv_match_window: bad argument"); return AFV_ERR_INVALID; }
    if (nq == 0) return AFV_OK;
    const float invW = (float)AFV_GRID_COLS / (max_x - min_x), invH = (float)AFV_GRID_ROWS / (max_y - min_y);
    k_match_window<<<(nq + 7) / 8, 256, 0, as_stream(cuda_stream)>>>(desc_type, D, (const uint8_t*)d_q, d_qxy, d_qr, d_qmin_size,
        d_qmax_size, nq, d_tk, (const uint8_t*)d_td, d_tsize, d_cell_start, d_cell_items, min_x, min_y, invW, invH,
        d_best, d_bestd, d_secondd, d_best_size, d_second_size);
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

// ---- windowed matcher batched over frame pairs (the throughput form of the SearchByProjection / GetFeaturesInArea core) ----
// One CTA per pair: the train frame (positions, sizes, descriptors) and its prebuilt grid (afv_grid_build = the reference's
// Frame::AssignFeaturesToGrid, done once per frame) are staged in shared memory with coalesced loads -- these loads ARE the
// algorithmic HBM traffic of the pair.  Then the WARP is the unit of work.  A warp takes 32 queries; the cell-column ranges of
// their windows are walked slot-parallel, the candidates that pass the size and radius gates are pushed (query lane, CSR position)
// into a per-warp shared-memory queue by ballot compaction, and the queue is drained one SLOT per lane: Hamming distance between
// the staged train row and the query row (kept word-major in shared memory for the 32 queries of the batch), key = distance << 16 |
// CSR position.  CSR positions ascend in the reference's enumeration order (cell x, cell y, index), so the smallest key is "first
// minimum wins" and the second smallest key carries the multiset-second-smallest distance, whatever order the slots are evaluated in.
// Only __syncwarp() after the staging barrier.
//
// What was measured on B200 before this form (10 240 pairs, r = 15 / 100 px; profiles/r02_summary.md):
//   thread per query walking its cell columns, query in registers ........ 1.03 / 10.1 ms  (issue bound, 8 of 32 lanes: the window
//                                                                                            populations of 32 neighbouring queries differ)
//   + lane-per-query cursor feeding the queue, atomicMin top-2 ............ 0.87 / 10.5 ms  (25 walk iterations per batch at 13 lanes)
//   + scan-based slot expansion of the walk ............................... 0.82 / 13.0 ms  (LSU data pipe 84 %: shared-memory wavefronts)
//   + chunk-major descriptors read with LDS.128 ........................... 0.79 / 11.9 ms  (bank conflicts 80 M -> 68 M)
//   + owner read-back of contiguous runs instead of atomics ............... 0.83 /  9.8 ms  (no pass 2; kept: never behind, no atomics)
//   + slot owners by rank instead of a shuffle binary search (this kernel)  0.81 /  9.3 ms  (the search was 13 % of the stall samples)
// and, in round 1: enumerate -> flat candidate list -> uniform distance pass -> per-query reduce with block barriers (2 - 3.5x slower),
// 4 lanes per query over interleaved cell columns (1.0 - 1.3x slower), pairs grouped by train frame to stage it once per 8 pairs
// (1.20 vs 1.10 ms: wave-quantisation tail; the staging overlaps with other CTAs' query loops anyway).  The kernel sits at 70 - 84 %
// of the LSU data pipe: 15 - 18 k shared-memory wavefronts per pair (random 16-byte rows, queue traffic, shuffles) against 113 KB
// = 0.9 k wavefronts of compulsory staging, which is why it stays near 0.2 of the HBM roofline.
#define MWQ_QCAP 256
#define MWQ_RUNS 8                                                                      // column steps whose runs can wait for one drain
template <int NW, int THREADS>
__global__ void __launch_bounds__(THREADS) k_match_window_pairs(int D, const afv_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc,
        const float* __restrict__ kpsize, const int* __restrict__ n_arr, int cap, const int* __restrict__ cell_start,
        const int* __restrict__ cell_items, const int* __restrict__ pair_a, const int* __restrict__ pair_b,
        const float* __restrict__ qxy, const float* __restrict__ qr, float r_all, const float* __restrict__ qmin, const float* __restrict__ qmax,
        float minX, float minY, float invW, float invH, int* __restrict__ best, float* __restrict__ bestd, float* __restrict__ secondd) {
    extern __shared__ __align__(16) unsigned char sm[];
    constexpr int WARPS = THREADS / 32;
    constexpr int NC = NW / 4;                                                          // 16-byte chunks per descriptor
    constexpr int WSCR = MWQ_QCAP + NW * 32 + MWQ_RUNS * 32 + 32;                       // words of scratch per warp
    // train descriptors chunk-major: sdesc[c][row] is the c-th 16-byte quarter of row `row`.  A drain reads one row per lane with NC
    // LDS.128; rows are random, and a quarter warp of 8 random 16-byte slots (bank group = row mod 8) collides far less often than
    // 32 random 4-byte words did in the word-major layout of the thread-per-query kernel (28 -> 19 wavefronts per 32 descriptors).
    uint4* sdesc = reinterpret_cast<uint4*>(sm);                                        // [NC][cap]
    uint32_t* wscr = reinterpret_cast<uint32_t*>(sdesc + (size_t)NC * cap);             // [WARPS][WSCR]
    float2* sxy = reinterpret_cast<float2*>(wscr + WARPS * WSCR);                       // [cap]
    float* ssz = reinterpret_cast<float*>(sxy + cap);                                   // [cap]
    unsigned short* scs = reinterpret_cast<unsigned short*>(ssz + cap);                 // [NCELLS + 1]
    unsigned short* sci = scs + ((NCELLS + 1 + 7) & ~7);                                // [cap]
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fa = pair_a[p], fb = pair_b[p];
    const int n1 = min(n_arr[fa], cap), n2 = min(n_arr[fb], cap);
    const afv_keypoint* k1p = kps + (long long)fa * cap; const afv_keypoint* k2p = kps + (long long)fb * cap;
    const uint8_t* d1 = desc + (long long)fa * cap * D; const uint8_t* d2 = desc + (long long)fb * cap * D;
    const float* sz2 = kpsize + (long long)fb * cap;
    const int* cs = cell_start + (long long)fb * (NCELLS + 1); const int* ci = cell_items + (long long)fb * cap;
    // ---- stage the train frame
    if ((D & 15) == 0) {
        const uint4* w2 = reinterpret_cast<const uint4*>(d2);
        for (int i = tid; i < n2 * NC; i += THREADS) { const int r = i / NC, c = i - r * NC; sdesc[(size_t)c * cap + r] = w2[i]; }
    } else {
        for (int i = tid; i < n2 * NC; i += THREADS) {
            const int r = i / NC, c = i - r * NC;
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k] = 0;
                if ((D & 3) == 0) { if (c * 16 + k * 4 < D) v[k] = *reinterpret_cast<const uint32_t*>(d2 + (long long)r * D + c * 16 + k * 4); }
                else for (int b = 0; b < 4; ++b) { const int o = c * 16 + k * 4 + b; if (o < D) v[k] |= (uint32_t)d2[(long long)r * D + o] << (8 * b); }
            }
            sdesc[(size_t)c * cap + r] = make_uint4(v[0], v[1], v[2], v[3]);
        }
    }
    for (int i = tid; i < n2; i += THREADS) { sxy[i] = make_float2(k2p[i].x, k2p[i].y); ssz[i] = sz2[i]; sci[i] = (unsigned short)ci[i]; }
    for (int i = tid; i <= NCELLS; i += THREADS) scs[i] = (unsigned short)cs[i];
    __syncthreads();
    // ---- one warp per batch of 32 queries
    uint32_t* wq = wscr + warp * WSCR;                                                  // [MWQ_QCAP] queue: lane << 16 | CSR position, then lane << 27 | key
    uint32_t* wsq = wq + MWQ_QCAP;                                                      // [NW][32] query descriptors of the batch
    uint32_t* wrun = wsq + NW * 32;                                                     // [MWQ_RUNS][32] finished runs: start << 16 | count
    uint32_t* wown = wrun + MWQ_RUNS * 32;                                              // [32] owner table of the current column step
    const unsigned lt = (1u << lane) - 1u;
    for (int qb = warp * 32; qb < n1; qb += WARPS * 32) {
        const int qi = qb + lane;
        const bool act = qi < n1;
        const long long qo = (long long)p * cap + qi;
        float x = 0.0f, y = 0.0f, r = -1.0f, smin = -FLT_MAX, smax = FLT_MAX;
        if (act) {
            x = qxy ? qxy[2 * qo] : k1p[qi].x; y = qxy ? qxy[2 * qo + 1] : k1p[qi].y;
            r = qr ? qr[qo] : r_all;
            if (qmin) smin = qmin[qo];
            if (qmax) smax = qmax[qo];
            if ((D & 15) == 0) {
                const uint4* row = reinterpret_cast<const uint4*>(d1 + (long long)qi * D);
#pragma unroll
                for (int w = 0; w < NW / 4; ++w) { const uint4 v = row[w]; wsq[(4 * w) * 32 + lane] = v.x; wsq[(4 * w + 1) * 32 + lane] = v.y; wsq[(4 * w + 2) * 32 + lane] = v.z; wsq[(4 * w + 3) * 32 + lane] = v.w; }
            } else if ((D & 3) == 0) {
                const uint32_t* row = reinterpret_cast<const uint32_t*>(d1 + (long long)qi * D);
#pragma unroll
                for (int w = 0; w < NW; ++w) wsq[w * 32 + lane] = row[w];
            } else {
#pragma unroll
                for (int w = 0; w < NW; ++w) { uint32_t v = 0; for (int b = 0; b < 4; ++b) { const int o = w * 4 + b; if (o < D) v |= (uint32_t)d1[(long long)qi * D + o] << (8 * b); } wsq[w * 32 + lane] = v; }
            }
        }
        uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;
        int c0 = 0, c1 = -1, r0 = 0, r1 = 0;
        const bool win = act && !(r < 0.0f) && window_cells(x, y, r, minX, minY, invW, invH, c0, c1, r0, r1);
        // walk: column step k takes the k-th cell column of every query of the batch; the 32 CSR ranges are laid end to end (warp scan
        // of their lengths) and handed out one SLOT per lane, so the gate runs on full warps however unevenly the ranges are filled
        // (a lane-per-query cursor ran 25 iterations per batch at 13 of 32 lanes: corners cluster across pyramid levels).  Slots that
        // pass the gate are appended to the queue in slot order, so the pushes of one query form one contiguous RUN of the queue; the
        // owner lane tracks (run_start, run_cnt) from the ballot alone.  drain: distances one queue slot per lane, then every owner
        // reads its run back and folds the keys into its register top-2 -- no atomics, no second pass.
        const int ncol = win ? c1 - c0 + 1 : 0, dr = r1 + 1 - r0;
        const int maxcol = __reduce_max_sync(0xffffffffu, ncol);
        const bool gate_sz = qmin != nullptr || qmax != nullptr;
        int qn = 0, run_start = 0, run_cnt = 0, nrec = 0;
        auto drain = [&]() {
            __syncwarp();
            for (int s = lane; s < qn; s += 32) {
                const uint32_t e = wq[s];
                const int ql = (int)(e >> 16), cj = (int)(e & 0xffffu), idx = sci[cj];
                int d = 0;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const uint4 tv = sdesc[(size_t)c * cap + idx];
                    d += __popc(wsq[(4 * c) * 32 + ql] ^ tv.x) + __popc(wsq[(4 * c + 1) * 32 + ql] ^ tv.y) +
                         __popc(wsq[(4 * c + 2) * 32 + ql] ^ tv.z) + __popc(wsq[(4 * c + 3) * 32 + ql] ^ tv.w);
                }
                wq[s] = ((uint32_t)d << 16) | (uint32_t)cj;
            }
            __syncwarp();
            for (int rec = 0; rec <= nrec; ++rec) {                                     // finished runs of earlier column steps, then the open one
                const uint32_t rv = rec < nrec ? wrun[rec * 32 + lane] : (((uint32_t)run_start << 16) | (uint32_t)run_cnt);
                const int rs = (int)(rv >> 16), rc = (int)(rv & 0xffffu);
                for (int i = 0; i < rc; ++i) {
                    const uint32_t key = wq[rs + i];
                    if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
                }
            }
            run_cnt = 0; qn = 0; nrec = 0;
            __syncwarp();
        };
        __syncwarp();
        for (int k = 0; k < maxcol; ++k) {
            int j0 = 0, len = 0;
            if (k < ncol) { const int cp = (c0 + k) * AFV_GRID_ROWS + r0; j0 = scs[cp]; len = (int)scs[cp + dr] - j0; }
            int incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const int T = __shfl_sync(0xffffffffu, incl, 31);
            const int excl = incl - len, base = j0 - excl;                              // slot t of this lane's range is CSR position base + t
            // owner table of the step: the non-empty ranges in lane order, (lane, base) packed in one word.  A slot finds its owner
            // by RANK -- how many non-empty ranges end at or before it -- from one REDUX.OR mask per pass instead of a five-deep
            // chain of dependent shuffles (the binary search over the inclusive sums was 13 % of the kernel's stall samples).
            const unsigned nonempty = __ballot_sync(0xffffffffu, len > 0);
            __syncwarp();                                                               // readers of the previous step's table are done
            if (len > 0) wown[__popc(nonempty & lt)] = ((uint32_t)lane << 26) | (uint32_t)(base + (1 << 22));
            __syncwarp();
            for (int t0 = 0; t0 < T; t0 += 32) {
                const int t = t0 + lane;
                const int e = incl - t0 - 1;                                            // this lane's range ends at slot e of the pass
                const unsigned ends = __reduce_or_sync(0xffffffffu, (len > 0 && e >= 0 && e < 32) ? (1u << e) : 0u);
                const int before = __popc(__ballot_sync(0xffffffffu, len > 0 && incl <= t0));
                const uint32_t ow = wown[min(before + __popc(ends & lt), 31)];
                const int lo = (int)(ow >> 26);
                const int j = (int)(ow & 0x03ffffffu) - (1 << 22) + t;
                const float ox = __shfl_sync(0xffffffffu, x, lo), oy = __shfl_sync(0xffffffffu, y, lo);
                const float orr = qr ? __shfl_sync(0xffffffffu, r, lo) : r_all;
                float omin = -FLT_MAX, omax = FLT_MAX;
                if (gate_sz) { omin = __shfl_sync(0xffffffffu, smin, lo); omax = __shfl_sync(0xffffffffu, smax, lo); }
                bool push = false;
                if (t < T) {
                    const int idx = sci[j];
                    const float2 tp = sxy[idx];
                    push = fabsf(__fsub_rn(tp.x, ox)) < orr && fabsf(__fsub_rn(tp.y, oy)) < orr;
                    if (gate_sz && push) { const float sz = ssz[idx]; push = !(sz < omin || sz > omax); }
                }
                const unsigned pm = __ballot_sync(0xffffffffu, push);
                if (push) wq[qn + __popc(pm & lt)] = ((uint32_t)lo << 16) | (uint32_t)j;
                {                                                                       // this lane as an OWNER: its slots of the pass are lanes [a, b)
                    const int a = min(max(excl - t0, 0), 32), b = min(max(incl - t0, 0), 32);
                    const unsigned below_a = a >= 32 ? 0xffffffffu : ((1u << a) - 1u), below_b = b >= 32 ? 0xffffffffu : ((1u << b) - 1u);
                    if (run_cnt == 0) run_start = qn + __popc(pm & below_a);
                    run_cnt += __popc(pm & below_b & ~below_a);
                }
                qn += __popc(pm);
                if (qn > MWQ_QCAP - 32) drain();
            }
            // end of the column step: park this lane's run; drain when the queue or the run table is nearly full, or at the end
            // (nrec is per lane: only non-empty runs are parked, so the read-back loops over the steps that DID push for this query)
            if (k == maxcol - 1 || qn > MWQ_QCAP - 64 || __any_sync(0xffffffffu, nrec == MWQ_RUNS - 1)) { if (qn > 0) drain(); }
            else if (run_cnt > 0) { wrun[nrec * 32 + lane] = ((uint32_t)run_start << 16) | (uint32_t)run_cnt; ++nrec; run_cnt = 0; }
        }
        if (act) {
            best[qo] = k1 == 0xffffffffu ? -1 : (int)sci[k1 & 0xffffu];
            bestd[qo] = k1 == 0xffffffffu ? FLT_MAX : (float)(k1 >> 16);
            secondd[qo] = k2 == 0xffffffffu ? FLT_MAX : (float)(k2 >> 16);
        }
        __syncwarp();                                                                   // wsq is rewritten by the next batch
    }
}

extern "C" int afv_match_window_pairs(int desc_type, const afv_keypoint* d_kps, const void* d_desc, const float* d_kpsize, const int* d_n,
        int B, int cap, const int* d_cell_start, const int* d_cell_items, const int* d_pair_a, const int* d_pair_b, int P,
        const float* d_qxy, const float* d_qr, float radius, const float* d_qmin_size, const float* d_qmax_size,
        float min_x, float min_y, float max_x, float max_y, int* d_best, float* d_bestd, float* d_secondd, void* cuda_stream) {
    const int D = desc_bytes(desc_type);
    if (D < 0 || desc_type == AFV_FEAT_SIFT128 || !d_kps || !d_desc || !d_kpsize || !d_n || !d_cell_start || !d_cell_items || !d_pair_a || !d_pair_b ||
        !d_best || !d_bestd || !d_secondd || B < 1 || P < 0 || cap < 1 || cap >= 65535) { afv_set_error("afv_match_window_pairs: bad argument (binary descriptors only)"); return AFV_ERR_INVALID; }
    if (P == 0) return AFV_OK;
    const int NW = (((D + 3) / 4) + 3) & ~3;                                           // descriptor words, whole 16-byte chunks (8 / 12 / 16)
    const size_t smem_frame = (size_t)NW * cap * 4 + (size_t)cap * (8 + 4 + 2) + (size_t)((NCELLS + 1 + 7) & ~7) * 2 + 16;
    const size_t scratch_warp = (size_t)(MWQ_QCAP + NW * 32 + MWQ_RUNS * 32 + 32) * 4;
    const size_t smem_limit = 227 * 1024;                                               // dynamic shared memory per CTA on sm_100
    // 16 warps per CTA when the frame leaves room for their queues, else 8
    const int threads = smem_frame + 16 * scratch_warp <= smem_limit ? 512 : 256;
    const size_t smem = smem_frame + (size_t)(threads / 32) * scratch_warp;
    if (smem > smem_limit) { afv_set_error("afv_match_window_pairs: cap %d too large for the shared-memory staging", cap); return AFV_ERR_INVALID; }
    cudaStream_t st = as_stream(cuda_stream);
    const float invW = (float)AFV_GRID_COLS / (max_x - min_x), invH = (float)AFV_GRID_ROWS / (max_y - min_y);
    AfvProfScope ps("k_match_window_pairs", st);
#define MWP_LAUNCH(W, T) do { AFV_CUDA_CHECK(cudaFuncSetAttribute(k_match_window_pairs<W, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_match_window_pairs<W, T><<<P, T, smem, st>>>(D, d_kps, (const uint8_t*)d_desc, d_kpsize, d_n, cap, d_cell_start, d_cell_items, d_pair_a, d_pair_b, \
            d_qxy, d_qr, radius, d_qmin_size, d_qmax_size, min_x, min_y, invW, invH, d_best, d_bestd, d_secondd); } while (0)
    if (threads == 512) { if (NW == 8) MWP_LAUNCH(8, 512); else if (NW == 12) MWP_LAUNCH(12, 512); else MWP_LAUNCH(16, 512); }
    else                { if (NW == 8) MWP_LAUNCH(8, 256); else if (NW == 12) MWP_LAUNCH(12, 256); else MWP_LAUNCH(16, 256); }
#undef MWP_LAUNCH
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

// ---- SearchForInitialization: one CTA per frame pair ----------------------------------------------------
// Train frame staged in shared memory (x, y, size, grid cell, descriptor when it fits); queries processed in
// index order; per query every thread scans a slice of the train keypoints: a keypoint is a candidate iff its
// grid cell lies in the query's cell range (this reproduces the round()-vs-floor/ceil quirk of the reference
// grid exactly), passes the size gate and |dx|,|dy| < r.  Enumeration order of the reference (cell x, cell y,
// index) is the order key.
__device__ __forceinline__ int rot_bin(float a1, float a2) {       // updateRotationHistogram (:1587-1597)
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    const float rotFactor = __fdiv_rn(1.0f, (float)AFV_HISTO_LENGTH);
    int bin = (int)roundf(__fmul_rn(rot, rotFactor));
    if (bin == AFV_HISTO_LENGTH) bin = 0;
    return bin;
}
__device__ __forceinline__ void three_maxima(const int* cnt, int& ind1, int& ind2, int& ind3) {   // :1631-1668
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < AFV_HISTO_LENGTH; ++i) {
        const int s = cnt[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
}

// v4 design: two kernels.
//  k_sfi_lists   (throughput part, one CTA per frame pair, one WARP per query): the train frame is counting-sorted by
//                grid column into shared memory (positions, cells, descriptors) so a query scans only its cell-column
//                range with converged lanes; every (query, candidate) distance is computed exactly once and appended
//                to the pair's candidate pool in global memory (two passes per query: count -> bump-allocate -> fill).
//                Nothing here depends on the sequential matching state.
//  k_sfi_resolve (sequential part, one WARP per frame pair, state in shared memory): queries in index order; lanes
//                stream the query's candidate list, drop candidates whose recorded match distance is <= this distance
//                (:511-512), reduce (distance, reference enumeration order) top-2, and lane 0 applies the acceptance /
//                stealing / rotation-histogram logic.  A query whose list did not fit the pool is rescanned brute force
//                with the same candidate definition (exact fallback).
#define SFL_THREADS 256
#define SFR_WARPS 2
#define SFR_QSTAGE 512
__host__ __device__ inline size_t sfl_base_bytes(int cap) { return (((size_t)cap * (4 * 2 + 2 * 4)) + 15) & ~(size_t)15; }


struct SfiQuery { float x, y; int c0, c1, r0, r1; bool ok; };
struct SfiQMeta { int i1, off, cnt; float angle; };         // per query, written by k_sfi_lists

// candidate test shared by both kernels: grid cell of the train keypoint inside the query's cell range + window test
__device__ __forceinline__ bool sfi_in_window(const SfiQuery& q, int cx, int cy, float tx, float ty, float window) {
    if (cx < q.c0 || cx > q.c1 || cy < q.r0 || cy > q.r1) return false;
    const float dx = __fsub_rn(tx, q.x), dy = __fsub_rn(ty, q.y);
    return fabsf(dx) < window && fabsf(dy) < window;
}

// Hamming distance of a query held in registers against slot `sl` of the word-interleaved staged train descriptors; NW words
template <int NW>
__device__ __forceinline__ int sfl_ham(const uint32_t* qd, const uint32_t* y32, int cap) {
    int d = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) d += __popc(qd[w] ^ y32[(size_t)w * cap]);
    return d;
}
#define SFL_WBUF 256         // per-warp buffer of in-window slots: distances are computed from it with all 32 lanes busy

template <bool BINARY>
__global__ void __launch_bounds__(SFL_THREADS) k_sfi_lists(int desc_type, int D, int Dpad, int stage_desc,
        const afv_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, const float* __restrict__ kpsize,
        const int* __restrict__ n_arr, int cap, const int* __restrict__ pair_a, const int* __restrict__ pair_b,
        float minX, float minY, float invW, float invH, float max_kpt_size,
        const float* __restrict__ prev_matched, float window,
        void* __restrict__ pool_v, int pool_cap, SfiQMeta* __restrict__ qmeta, int* __restrict__ nq_out,
        int* __restrict__ gpool, int nsplit) {
    extern __shared__ __align__(16) unsigned char sm[];
    // blockIdx.y = slice of the pair's queries: with few pairs the queries of one pair are spread over nsplit CTAs (each
    // re-sorts the train frame; list space comes from one global bump counter per pair)
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int fa = pair_a[p], fb = pair_b[p];
    const int n1 = min(n_arr[fa], cap), n2 = min(n_arr[fb], cap);
    const afv_keypoint* k1 = kps + (long long)fa * cap;
    const afv_keypoint* k2 = kps + (long long)fb * cap;
    const uint8_t* d1 = desc + (long long)fa * cap * D;
    const uint8_t* d2 = desc + (long long)fb * cap * D;
    const float* size2 = kpsize + (long long)fb * cap;
    const float* pm = prev_matched ? prev_matched + (long long)p * cap * 2 : nullptr;
    SfiQMeta* qm = qmeta + (long long)p * cap;

    float* sx = reinterpret_cast<float*>(sm); float* sy = sx + cap;
    unsigned short* scell = reinterpret_cast<unsigned short*>(sy + cap);
    unsigned short* sorig = scell + cap; unsigned short* qlist = sorig + cap; unsigned short* tmpc = qlist + cap;
    uint8_t* sdesc = sm + sfl_base_bytes(cap);
    __shared__ int colstart[AFV_GRID_COLS + 1], colfill[AFV_GRID_COLS];
    __shared__ int s_nq, s_pool;
    const int nw = Dpad / 4;
    const unsigned short NONE16 = 0xffff;

    if (tid < AFV_GRID_COLS) { colstart[tid] = 0; colfill[tid] = 0; }
    if (tid == 0) { s_nq = 0; s_pool = 0; colstart[AFV_GRID_COLS] = 0; }
    __syncthreads();
    for (int i = tid; i < n2; i += SFL_THREADS) {
        const float sz = size2[i];
        const int c = grid_cell(k2[i].x, k2[i].y, minX, minY, invW, invH);
        // GetFeaturesInArea(..., minSize 0, maxSize F1.maxKeyPtSize) size gate folded in (:495-496, Frame.cc:365-368)
        const bool ok = c >= 0 && !(sz < 0.0f) && !(sz > max_kpt_size);
        tmpc[i] = ok ? (unsigned short)c : NONE16;
        if (ok) atomicAdd(&colstart[c / AFV_GRID_ROWS + 1], 1);
    }
    __syncthreads();
    if (wid == 0) {                                                 // inclusive scan of 64 column counts
        int a0 = colstart[1 + lane], a1 = colstart[33 + lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t0 = __shfl_up_sync(0xffffffffu, a0, o), t1 = __shfl_up_sync(0xffffffffu, a1, o); if (lane >= o) { a0 += t0; a1 += t1; } }
        const int tot0 = __shfl_sync(0xffffffffu, a0, 31);
        colstart[1 + lane] = a0; colstart[33 + lane] = tot0 + a1;
    }
    if (wid == 1) {                                                 // ordered compaction of the octave-0 queries (:491-493)
        int base = 0;
        for (int i0 = 0; i0 < n1; i0 += 32) {
            const int i = i0 + lane;
            const bool q = i < n1 && k1[i].octave <= 0;
            const unsigned m = __ballot_sync(0xffffffffu, q);
            if (q) qlist[base + __popc(m & ((1u << lane) - 1))] = (unsigned short)i;
            base += __popc(m);
        }
        if (lane == 0) { s_nq = base; if (blockIdx.y == 0) nq_out[p] = base; }
    }
    __syncthreads();
    for (int i = tid; i < n2; i += SFL_THREADS) {                   // scatter into column order
        const int c = tmpc[i];
        if (c == NONE16) continue;
        const int cx = c / AFV_GRID_ROWS;
        const int slot = colstart[cx] + atomicAdd(&colfill[cx], 1);
        sx[slot] = k2[i].x; sy[slot] = k2[i].y;
        scell[slot] = (unsigned short)((cx << 8) | (c % AFV_GRID_ROWS));
        sorig[slot] = (unsigned short)i;
    }
    __syncthreads();
    const int n2s = colstart[AFV_GRID_COLS];
    const int nq = s_nq;
    if (stage_desc) {
        for (int i = tid; i < n2s * nw; i += SFL_THREADS) {
            const int r = i / nw, w = i % nw;
            const uint8_t* row = d2 + (long long)sorig[r] * D;
            uint32_t v;
            if ((D & 3) == 0) v = reinterpret_cast<const uint32_t*>(row)[w];
            else { v = 0; for (int b = 0; b < 4; ++b) { const int o = w * 4 + b; if (o < D) v |= (uint32_t)row[o] << (8 * b); } }
            reinterpret_cast<uint32_t*>(sdesc)[(size_t)w * cap + r] = v;      // word-interleaved: lanes on consecutive slots hit consecutive banks
        }
    }
    __syncthreads();

    for (int qi = wid + (SFL_THREADS / 32) * blockIdx.y; qi < nq; qi += (SFL_THREADS / 32) * nsplit) {
        const int i1 = qlist[qi];
        SfiQuery q;
        q.x = pm ? pm[2 * i1] : k1[i1].x; q.y = pm ? pm[2 * i1 + 1] : k1[i1].y;
        q.ok = window_cells(q.x, q.y, window, minX, minY, invW, invH, q.c0, q.c1, q.r0, q.r1);
        int cnt = 0, off = 0;
        if (q.ok) {
            const int s0 = colstart[q.c0], s1 = colstart[q.c1 + 1];
            // pass 1: count
            for (int sb = s0; sb < s1; sb += 32) {
                const int sl = sb + lane;
                bool pass = false;
                if (sl < s1) { const unsigned short cc = scell[sl]; pass = sfi_in_window(q, cc >> 8, cc & 0xff, sx[sl], sy[sl], window); }
                cnt += __popc(__ballot_sync(0xffffffffu, pass));
            }
            if (cnt > 0) {
                if (lane == 0) off = atomicAdd(&gpool[p], cnt);
                off = __shfl_sync(0xffffffffu, off, 0);
                if (off + cnt > pool_cap) off = -1;                 // pool exhausted: the resolver rescans this query
            }
            if (cnt > 0 && off >= 0) {
                uint32_t qd[16];
                if (BINARY) {
                    const uint8_t* qrow = d1 + (long long)i1 * D;
                    if ((D & 3) == 0) {
#pragma unroll
                        for (int w = 0; w < 16; ++w) if (w < nw) qd[w] = reinterpret_cast<const uint32_t*>(qrow)[w];
                    } else {
#pragma unroll
                        for (int w = 0; w < 16; ++w) if (w < nw) { uint32_t v = 0; for (int b = 0; b < 4; ++b) { const int o = w * 4 + b; if (o < D) v |= (uint32_t)qrow[o] << (8 * b); } qd[w] = v; }
                    }
                }
                int run = off;
                if (BINARY && stage_desc) {
                    // The window test passes for about a third of the scanned slots, so computing the distance inside this loop ran at
                    // 11 of 32 lanes (38 % of the kernel's instructions).  The in-window slots are collected in a per-warp buffer (in
                    // enumeration order) and the distances are computed from the buffer with every lane busy, NW unrolled.
                    __shared__ unsigned short wbuf[SFL_THREADS / 32][SFL_WBUF];
                    unsigned short* wb = wbuf[wid];
                    uint32_t* pool = reinterpret_cast<uint32_t*>(pool_v) + (long long)p * pool_cap;
                    int nb = 0;
                    auto flush = [&]() {
                        __syncwarp();
                        for (int j = lane; j < nb; j += 32) {
                            const int sl = wb[j];
                            const uint32_t* y32 = reinterpret_cast<const uint32_t*>(sdesc) + sl;
                            int d;
                            switch (nw) {
                                case 8: d = sfl_ham<8>(qd, y32, cap); break;
                                case 12: d = sfl_ham<12>(qd, y32, cap); break;
                                default: d = sfl_ham<16>(qd, y32, cap); break;          // 61-byte rows: 16 words, zero padded
                            }
                            pool[run + j] = ((uint32_t)d << 20) | (uint32_t)sorig[sl];
                        }
                        run += nb; nb = 0;
                        __syncwarp();
                    };
                    for (int sb = s0; sb < s1; sb += 32) {
                        const int sl = sb + lane;
                        bool pass = false;
                        if (sl < s1) { const unsigned short cc = scell[sl]; pass = sfi_in_window(q, cc >> 8, cc & 0xff, sx[sl], sy[sl], window); }
                        const unsigned m = __ballot_sync(0xffffffffu, pass);
                        if (pass) wb[nb + __popc(m & ((1u << lane) - 1))] = (unsigned short)sl;
                        nb += __popc(m);
                        if (nb > SFL_WBUF - 32) flush();
                    }
                    if (nb) flush();
                } else
                for (int sb = s0; sb < s1; sb += 32) {
                    const int sl = sb + lane;
                    bool pass = false;
                    if (sl < s1) { const unsigned short cc = scell[sl]; pass = sfi_in_window(q, cc >> 8, cc & 0xff, sx[sl], sy[sl], window); }
                    const unsigned m = __ballot_sync(0xffffffffu, pass);
                    if (pass) {
                        const int i2 = sorig[sl];
                        const int pos = run + __popc(m & ((1u << lane) - 1));
                        if (BINARY) {
                            int d = 0;
                            if (stage_desc) {
                                const uint32_t* y32 = reinterpret_cast<const uint32_t*>(sdesc) + sl;
#pragma unroll
                                for (int w = 0; w < 16; ++w) if (w < nw) d += __popc(qd[w] ^ y32[(size_t)w * cap]);
                            } else d = hamming_bytes(d1 + (long long)i1 * D, d2 + (long long)i2 * D, D);
                            reinterpret_cast<uint32_t*>(pool_v)[(long long)p * pool_cap + pos] = ((uint32_t)d << 20) | (uint32_t)i2;
                        } else {
                            const float dist = l2sqr128((const float*)(d1 + (long long)i1 * D), (const float*)(d2 + (long long)i2 * D));
                            reinterpret_cast<unsigned long long*>(pool_v)[(long long)p * pool_cap + pos] = make_key(dist, (uint32_t)i2);
                        }
                    }
                    run += __popc(m);
                }
            }
        }
        if (lane == 0) { SfiQMeta m; m.i1 = i1; m.off = off; m.cnt = cnt; m.angle = k1[i1].angle; qm[qi] = m; }
    }
}

// k_sfi_resolve, v5: one CTA (8 warps) per frame pair.  What the sequential semantics really need per query is the first
// two candidates, in (distance, reference enumeration order), that are not masked by the matching state (:511-512).  The
// state-INDEPENDENT part of that -- the sorted order itself -- is produced ahead of the sequential loop: for a block of
// SFR_QB queries all 8 warps extract the exact K smallest keys of each query's candidate list (K rounds of warp-min over
// the list held in registers).  Warp 0 then walks the block in query order: K lanes test their entry against the state,
// a ballot picks the first two survivors, lane 0 applies acceptance / stealing / histogram.  If fewer than two of the K
// survive and the list is longer than K, that query falls back to the full scan (exact).  The sequential chain per query
// drops from ~350 dependent instructions (list scan + 64-bit warp reduction) to ~50.
#define SFR_THREADS 256
#define SFR_QB 64            // queries per block (phase 1 parallel, phase 2 sequential)
#define SFR_K 8              // sorted prefix length per query
#define SFR_LMAX 256         // lists longer than this skip the prefix (full scan in phase 2)
__host__ __device__ inline size_t sfr_cta_bytes(int cap) {
    return ((((size_t)cap * (4 * 2 + 2 * 3 + 1)) + 15) & ~(size_t)15) + (size_t)SFR_QSTAGE * 16 + (size_t)SFR_QB * SFR_K * 8 + (size_t)SFR_QB * 4;
}

template <bool BINARY>
__global__ void __launch_bounds__(SFR_THREADS) k_sfi_resolve(int desc_type, int D,
        const afv_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, const float* __restrict__ kpsize,
        const int* __restrict__ n_arr, int cap, const int* __restrict__ pair_a, const int* __restrict__ pair_b, int P,
        float minX, float minY, float invW, float invH, float max_kpt_size,
        float* __restrict__ prev_matched, float window, float th_low, float nnratio, int check_ori,
        const void* __restrict__ pool_v, int pool_cap, const SfiQMeta* __restrict__ qmeta, const int* __restrict__ nq_arr,
        int* __restrict__ matches12, int* __restrict__ nmatches) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int p = blockIdx.x;
    float* matched = reinterpret_cast<float*>(sm); float* tang = matched + cap;
    unsigned short* m21 = reinterpret_cast<unsigned short*>(tang + cap);
    unsigned short* m12 = m21 + cap; unsigned short* tcell = m12 + cap;
    signed char* hbin = reinterpret_cast<signed char*>(tcell + cap);
    unsigned char* tail = sm + ((((size_t)cap * (4 * 2 + 2 * 3 + 1)) + 15) & ~(size_t)15);
    SfiQMeta* qms = reinterpret_cast<SfiQMeta*>(tail);
    unsigned long long* topk = reinterpret_cast<unsigned long long*>(tail + (size_t)SFR_QSTAGE * 16);      // [SFR_QB][SFR_K]
    int* tflag = reinterpret_cast<int*>(tail + (size_t)SFR_QSTAGE * 16 + (size_t)SFR_QB * SFR_K * 8);       // [SFR_QB] 1 = prefix valid
    __shared__ int hist[AFV_HISTO_LENGTH + 2];
    __shared__ int s_nm;
    const unsigned short NONE16 = 0xffff;

    const int fa = pair_a[p], fb = pair_b[p];
    const int n1 = min(n_arr[fa], cap), n2 = min(n_arr[fb], cap);
    const afv_keypoint* k1 = kps + (long long)fa * cap;
    const afv_keypoint* k2 = kps + (long long)fb * cap;
    const uint8_t* d1 = desc + (long long)fa * cap * D;
    const uint8_t* d2 = desc + (long long)fb * cap * D;
    const float* size2 = kpsize + (long long)fb * cap;
    float* pm = prev_matched ? prev_matched + (long long)p * cap * 2 : nullptr;
    int* m12g = matches12 + (long long)p * cap;
    const SfiQMeta* qm = qmeta + (long long)p * cap;
    const int nq = nq_arr[p];

    for (int i = tid; i < n2; i += SFR_THREADS) {
        matched[i] = FLT_MAX; m21[i] = NONE16; tang[i] = k2[i].angle;
        const float sz = size2[i];
        const int c = grid_cell(k2[i].x, k2[i].y, minX, minY, invW, invH);
        const bool ok = c >= 0 && !(sz < 0.0f) && !(sz > max_kpt_size);
        tcell[i] = ok ? (unsigned short)(((c / AFV_GRID_ROWS) << 8) | (c % AFV_GRID_ROWS)) : NONE16;
    }
    for (int i = tid; i < n1; i += SFR_THREADS) { m12[i] = NONE16; hbin[i] = -1; }
    if (tid < AFV_HISTO_LENGTH) hist[tid] = 0;
    for (int i = tid; i < min(nq, SFR_QSTAGE); i += SFR_THREADS) qms[i] = qm[i];
    __syncthreads();
    int nm = 0;                                                    // thread 0's copy is authoritative

    auto pool_key = [&](const SfiQMeta& m, int j) -> unsigned long long {          // candidate j of a query as a (distance, order) key
        unsigned long long k;
        if (BINARY) {
            const uint32_t e = reinterpret_cast<const uint32_t*>(pool_v)[(long long)p * pool_cap + m.off + j];
            k = ((unsigned long long)__float_as_uint((float)(e >> 20)) << 32) | (e & 0xfffffu);
        } else k = reinterpret_cast<const unsigned long long*>(pool_v)[(long long)p * pool_cap + m.off + j];
        const int i2 = (int)((uint32_t)k & 0xfffffu);
        const unsigned short cc = tcell[i2];
        return (k & 0xffffffff00000000ull) | ((uint32_t)(cc >> 8) << 26) | ((uint32_t)(cc & 0xff) << 20) | (uint32_t)i2;
    };

    for (int qb = 0; qb < nq; qb += SFR_QB) {
        const int nb = min(SFR_QB, nq - qb);
        // ---- phase 1 (all warps): exact sorted prefix of every query of the block
        for (int s = wid; s < nb; s += SFR_THREADS / 32) {
            const int qi = qb + s;
            const SfiQMeta m = qi < SFR_QSTAGE ? qms[qi] : qm[qi];
            const bool ok = m.off >= 0 && m.cnt > 0 && m.cnt <= SFR_LMAX;
            if (ok && BINARY && cap <= 2048) {
                // compact keys: Hamming distance (<= 488 < 511: 9 bits) | cell x (6) | cell y (6) | train index (11) fit 32 bits, so a
                // round is 7 unsigned minima, ONE redux.sync warp minimum and 8 compare-selects instead of 64-bit compare chains and ten
                // shuffles (the top-K extraction was 60 % of this kernel's instructions and makes it ALU bound at >= 2048 pairs)
                uint32_t key[SFR_LMAX / 32];
                const uint32_t NONE32 = 0xffffffffu;
#pragma unroll
                for (int u = 0; u < SFR_LMAX / 32; ++u) {
                    const int j = u * 32 + lane;
                    uint32_t k = NONE32;
                    if (j < m.cnt) {
                        const uint32_t e = reinterpret_cast<const uint32_t*>(pool_v)[(long long)p * pool_cap + m.off + j];
                        const uint32_t i2 = e & 0xfffffu;
                        const unsigned short cc = tcell[i2];
                        k = ((e >> 20) << 23) | ((uint32_t)(cc >> 8) << 17) | ((uint32_t)(cc & 0xff) << 11) | i2;
                    }
                    key[u] = k;
                }
                for (int r = 0; r < SFR_K; ++r) {
                    uint32_t lmin = key[0];
#pragma unroll
                    for (int u = 1; u < SFR_LMAX / 32; ++u) lmin = min(lmin, key[u]);
                    const uint32_t g = __reduce_min_sync(0xffffffffu, lmin);
                    if (g == NONE32) { if (lane == 0) for (int r2 = r; r2 < SFR_K; ++r2) topk[s * SFR_K + r2] = KEY_NONE; break; }
                    if (lane == 0)      // back to the resolver's key format: float distance bits << 32 | cell x << 26 | cell y << 20 | index
                        topk[s * SFR_K + r] = ((unsigned long long)__float_as_uint((float)(g >> 23)) << 32) | (((g >> 17) & 63u) << 26) |
                                              (((g >> 11) & 63u) << 20) | (g & 2047u);
#pragma unroll
                    for (int u = 0; u < SFR_LMAX / 32; ++u) if (key[u] == g) key[u] = NONE32;       // keys are unique (index bits)
                }
            } else if (ok) {
                unsigned long long key[SFR_LMAX / 32];
#pragma unroll
                for (int u = 0; u < SFR_LMAX / 32; ++u) { const int j = u * 32 + lane; key[u] = j < m.cnt ? pool_key(m, j) : KEY_NONE; }
                for (int r = 0; r < SFR_K; ++r) {
                    unsigned long long lmin = key[0];
#pragma unroll
                    for (int u = 1; u < SFR_LMAX / 32; ++u) lmin = key[u] < lmin ? key[u] : lmin;
                    unsigned long long g = lmin;
#pragma unroll
                    for (int o = 16; o; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, g, o); g = t < g ? t : g; }
                    if (lane == 0) topk[s * SFR_K + r] = g;
                    if (g == KEY_NONE) { if (lane == 0) for (int r2 = r + 1; r2 < SFR_K; ++r2) topk[s * SFR_K + r2] = KEY_NONE; break; }
#pragma unroll
                    for (int u = 0; u < SFR_LMAX / 32; ++u) if (key[u] == g) key[u] = KEY_NONE;     // keys are unique (index bits)
                }
            }
            if (lane == 0) tflag[s] = ok ? 1 : 0;
        }
        __syncthreads();
        // ---- phase 2 (warp 0): the reference's sequential loop over the block
        if (wid == 0) {
            for (int s = 0; s < nb; ++s) {
                const int qi = qb + s;
                const SfiQMeta q = qi < SFR_QSTAGE ? qms[qi] : qm[qi];
                if (q.cnt == 0) continue;
                unsigned long long b1 = KEY_NONE, b2 = KEY_NONE;
                bool done = false;
                if (tflag[s]) {
                    unsigned long long key = lane < SFR_K ? topk[s * SFR_K + lane] : KEY_NONE;
                    bool unf = false;
                    if (key != KEY_NONE) { const int i2 = (int)((uint32_t)key & 0xfffffu); unf = !(matched[i2] <= key_dist(key)); }   // :511-512
                    unsigned msk = __ballot_sync(0xffffffffu, unf);
                    if (__popc(msk) >= 2 || q.cnt <= SFR_K) {
                        if (msk) { const int l1 = __ffs(msk) - 1; b1 = __shfl_sync(0xffffffffu, key, l1); msk &= msk - 1; }
                        if (msk) { const int l2 = __ffs(msk) - 1; b2 = __shfl_sync(0xffffffffu, key, l2); }
                        done = true;
                    }
                }
                if (!done) {
                    Top2 t; t.k1 = t.k2 = KEY_NONE;
                    if (q.off >= 0) {
                        for (int j = lane; j < q.cnt; j += 32) {
                            const unsigned long long key = pool_key(q, j);
                            if (matched[(int)((uint32_t)key & 0xfffffu)] <= key_dist(key)) continue;                // :511-512
                            top2_push(t, key);
                        }
                    } else {
                        // exact fallback (candidate pool exhausted): rescan every train keypoint with the same candidate definition
                        SfiQuery w;
                        w.x = pm ? pm[2 * q.i1] : k1[q.i1].x; w.y = pm ? pm[2 * q.i1 + 1] : k1[q.i1].y;
                        w.ok = window_cells(w.x, w.y, window, minX, minY, invW, invH, w.c0, w.c1, w.r0, w.r1);
                        for (int i2 = lane; w.ok && i2 < n2; i2 += 32) {
                            const unsigned short cc = tcell[i2];
                            if (cc == NONE16 || !sfi_in_window(w, cc >> 8, cc & 0xff, k2[i2].x, k2[i2].y, window)) continue;
                            const float dist = desc_distance(desc_type, d1 + (long long)q.i1 * D, d2 + (long long)i2 * D, D);
                            if (matched[i2] <= dist) continue;
                            top2_push(t, make_key(dist, ((uint32_t)(cc >> 8) << 26) | ((uint32_t)(cc & 0xff) << 20) | (uint32_t)i2));
                        }
                    }
                    top2_warp_reduce(t);
                    b1 = t.k1; b2 = t.k2;
                }
                if (lane == 0 && b1 != KEY_NONE) {
                    const float bestDist = key_dist(b1), bestDist2 = key_dist(b2);
                    if (bestDist <= th_low && bestDist < __fmul_rn(bestDist2, nnratio)) {           // :526-528
                        const int bestIdx2 = (int)((uint32_t)b1 & 0xfffffu);
                        if (m21[bestIdx2] != NONE16) { m12[m21[bestIdx2]] = NONE16; nm--; }         // :530-534
                        m12[q.i1] = (unsigned short)bestIdx2; m21[bestIdx2] = (unsigned short)q.i1; matched[bestIdx2] = bestDist; nm++;
                        if (check_ori) { const int bin = rot_bin(q.angle, tang[bestIdx2]); hbin[q.i1] = (signed char)bin; hist[bin]++; }
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }
    if (tid == 0) s_nm = nm;
    __syncthreads();
    if (wid == 0) {
        nm = s_nm;
        if (check_ori) {
            int kb0 = 0, kb1 = 0, kb2 = 0;
            if (lane == 0) three_maxima(hist, kb0, kb1, kb2);
            kb0 = __shfl_sync(0xffffffffu, kb0, 0); kb1 = __shfl_sync(0xffffffffu, kb1, 0); kb2 = __shfl_sync(0xffffffffu, kb2, 0);
            int removed = 0;
            for (int i = lane; i < n1; i += 32) {
                const int b = hbin[i];
                if (b < 0 || b == kb0 || b == kb1 || b == kb2) continue;
                if (m12[i] != NONE16) { m12[i] = NONE16; ++removed; }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
            nm -= removed;
        }
        if (lane == 0) nmatches[p] = nm;
    }
    __syncthreads();
    for (int i = tid; i < n1; i += SFR_THREADS) {
        const int m = m12[i] == NONE16 ? -1 : (int)m12[i];
        m12g[i] = m;
        if (pm && m >= 0) { pm[2 * i] = k2[m].x; pm[2 * i + 1] = k2[m].y; }                 // :552-554
    }
}

#define SFI_POOL_CAP (64 * 1024)          // candidate-pool entries per frame pair
extern "C" size_t afv_search_for_initialization_workspace_bytes(int desc_type, int P, int cap) {
    const size_t esz = desc_type != AFV_FEAT_SIFT128 ? 4 : 8;
    return (size_t)P * SFI_POOL_CAP * esz + (size_t)P * cap * sizeof(SfiQMeta) + 2 * (size_t)P * sizeof(int) + 256;
}
extern "C" int afv_search_for_initialization_ws(int desc_type, const afv_keypoint* d_kps, const void* d_desc,
        const float* d_kpsize, const int* d_n, int B, int cap, const int* d_pair_a, const int* d_pair_b, int P,
        float min_x, float min_y, float max_x, float max_y, float max_kpt_size,
        float* d_prev_matched, int window, float th_low, float nnratio, int check_orientation,
        int* d_matches12, int* d_nmatches, void* d_workspace, size_t workspace_bytes, void* cuda_stream);
extern "C" int afv_search_for_initialization(int desc_type, const afv_keypoint* d_kps, const void* d_desc,
        const float* d_kpsize, const int* d_n, int B, int cap, const int* d_pair_a, const int* d_pair_b, int P,
        float min_x, float min_y, float max_x, float max_y, float max_kpt_size,
        float* d_prev_matched, int window, float th_low, float nnratio, int check_orientation,
        int* d_matches12, int* d_nmatches, void* cuda_stream) {
    return afv_search_for_initialization_ws(desc_type, d_kps, d_desc, d_kpsize, d_n, B, cap, d_pair_a, d_pair_b, P, min_x, min_y, max_x, max_y,
                                            max_kpt_size, d_prev_matched, window, th_low, nnratio, check_orientation, d_matches12, d_nmatches,
                                            nullptr, 0, cuda_stream);
}
extern "C" int afv_search_for_initialization_ws(int desc_type, const afv_keypoint* d_kps, const void* d_desc,
        const float* d_kpsize, const int* d_n, int B, int cap, const int* d_pair_a, const int* d_pair_b, int P,
        float min_x, float min_y, float max_x, float max_y, float max_kpt_size,
        float* d_prev_matched, int window, float th_low, float nnratio, int check_orientation,
        int* d_matches12, int* d_nmatches, void* d_workspace, size_t workspace_bytes, void* cuda_stream) {
    const int D = desc_bytes(desc_type);
    if (D < 0 || !d_kps || !d_desc || !d_kpsize || !d_n || !d_pair_a || !d_pair_b || !d_matches12 ||
        !d_nmatches || B < 1 || P < 0 || cap < 1) { afv_set_error("afv_search_for_initialization: bad argument"); return AFV_ERR_INVALID; }
    if (P == 0) return AFV_OK;
    if (cap >= 65535) { afv_set_error("cap too large for the 16-bit indices"); return AFV_ERR_INVALID; }
    cudaStream_t st = as_stream(cuda_stream);
    const bool binary = desc_type != AFV_FEAT_SIFT128;
    const int Dpad = (D + 3) & ~3;
    const size_t baseA = sfl_base_bytes(cap);
    const int stage = binary && Dpad <= 64 && baseA + (size_t)cap * Dpad <= 200 * 1024;
    const size_t smemA = baseA + (stage ? (size_t)cap * Dpad : 0);
    const size_t smemB = sfr_cta_bytes(cap);
    if (smemA > 220 * 1024 || smemB > 220 * 1024) { afv_set_error("cap %d too large for the shared-memory staging", cap); return AFV_ERR_INVALID; }
    // the opt-in shared-memory size is a per-device function attribute: set on every call (cheap, idempotent, thread-safe)
    AFV_CUDA_CHECK(binary ? cudaFuncSetAttribute(k_sfi_lists<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA)
                          : cudaFuncSetAttribute(k_sfi_lists<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
    AFV_CUDA_CHECK(binary ? cudaFuncSetAttribute(k_sfi_resolve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB)
                          : cudaFuncSetAttribute(k_sfi_resolve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB));
    {       // keep freed scratch cached in the stream-ordered pool of THIS device instead of returning it to the OS (once per device)
        static std::atomic<unsigned long long> pool_cfg_mask{0};
        int dev = 0; cudaMemPool_t mp;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !(pool_cfg_mask.load() & (1ull << dev)) &&
            cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess) {
            unsigned long long thr = ~0ull; cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
            pool_cfg_mask.fetch_or(1ull << dev);
        }
    }
    // candidate pool: entries per pair (u32 for Hamming: dist<<20 | index; u64 for L2: float bits<<32 | index)
    const int pool_cap = SFI_POOL_CAP;
    const size_t esz = binary ? 4 : 8;
    unsigned char* scratch = (unsigned char*)d_workspace;
    const size_t pool_bytes = (size_t)P * pool_cap * esz, meta_bytes = (size_t)P * cap * sizeof(SfiQMeta), nq_bytes = (size_t)P * sizeof(int);
    const size_t need = afv_search_for_initialization_workspace_bytes(desc_type, P, cap);
    if (scratch && (workspace_bytes < need || ((uintptr_t)scratch & 15))) {
        afv_set_error("afv_search_for_initialization: workspace %zu B < required %zu B (or not 16-byte aligned)", workspace_bytes, need); return AFV_ERR_INVALID;
    }
    if (!scratch) AFV_CUDA_CHECK(cudaMallocAsync((void**)&scratch, need, st));      // no caller workspace: stream-ordered allocation (cached pool)
    void* pool = scratch;
    SfiQMeta* qmeta = reinterpret_cast<SfiQMeta*>(scratch + pool_bytes);
    int* nq = reinterpret_cast<int*>(scratch + pool_bytes + meta_bytes);
    int* gpool = nq + P;
    AFV_CUDA_CHECK(cudaMemsetAsync(gpool, 0, nq_bytes, st));
    int nsplit = 1;                                   // ~2 list CTAs per SM keep the chip busy when there are few pairs
    while (nsplit < 8 && P * nsplit * 2 <= 296) nsplit *= 2;
    const float invW = (float)AFV_GRID_COLS / (max_x - min_x), invH = (float)AFV_GRID_ROWS / (max_y - min_y);
    {
        AfvProfScope ps("k_sfi_lists", st);
        if (binary) k_sfi_lists<true><<<dim3(P, nsplit), SFL_THREADS, smemA, st>>>(desc_type, D, Dpad, stage, d_kps, (const uint8_t*)d_desc, d_kpsize, d_n, cap,
                d_pair_a, d_pair_b, min_x, min_y, invW, invH, max_kpt_size, d_prev_matched, (float)window, pool, pool_cap, qmeta, nq, gpool, nsplit);
        else k_sfi_lists<false><<<dim3(P, nsplit), SFL_THREADS, smemA, st>>>(desc_type, D, Dpad, stage, d_kps, (const uint8_t*)d_desc, d_kpsize, d_n, cap,
                d_pair_a, d_pair_b, min_x, min_y, invW, invH, max_kpt_size, d_prev_matched, (float)window, pool, pool_cap, qmeta, nq, gpool, nsplit);
        ++g_afv_launches;
    }
    {
        AfvProfScope ps("k_sfi_resolve", st);
        const int grid = P;
        if (binary) k_sfi_resolve<true><<<grid, SFR_THREADS, smemB, st>>>(desc_type, D, d_kps, (const uint8_t*)d_desc, d_kpsize, d_n, cap, d_pair_a, d_pair_b, P,
                min_x, min_y, invW, invH, max_kpt_size, d_prev_matched, (float)window, th_low, nnratio, check_orientation, pool, pool_cap, qmeta, nq, d_matches12, d_nmatches);
        else k_sfi_resolve<false><<<grid, SFR_THREADS, smemB, st>>>(desc_type, D, d_kps, (const uint8_t*)d_desc, d_kpsize, d_n, cap, d_pair_a, d_pair_b, P,
                min_x, min_y, invW, invH, max_kpt_size, d_prev_matched, (float)window, th_low, nnratio, check_orientation, pool, pool_cap, qmeta, nq, d_matches12, d_nmatches);
        ++g_afv_launches;
    }
    AFV_CUDA_CHECK(cudaGetLastError());
    if (!d_workspace) AFV_CUDA_CHECK(cudaFreeAsync(scratch, st));
    return AFV_OK;
}

// ---- SearchByProjection family: same two-kernel split as SearchForInitialization ------------------------------
// k_proj_lists: CTA per problem, warp per query, train frame counting-sorted by grid column in shared memory; candidate =
// grid cell in the query's cell range, size inside [min,max] (Frame::GetFeaturesInArea size gate, src/Frame.cc:365-368),
// |dx|,|dy| < r.  k_proj_resolve: warp per problem, "occupied" flags in shared memory, queries in order.
__host__ __device__ inline size_t prl_base_bytes(int cap) { return (((size_t)cap * (4 * 3 + 2 * 3)) + 15) & ~(size_t)15; }
__host__ __device__ inline size_t prr_warp_bytes(int cap) { return (((size_t)cap * (4 + 2 + 1)) + 15) & ~(size_t)15; }
struct ProjQMeta { int off, cnt; };

template <bool BINARY>
__global__ void __launch_bounds__(SFL_THREADS) k_proj_lists(int desc_type, int D, int Dpad, int stage_desc,
        const uint8_t* __restrict__ qdesc, const float* __restrict__ qxy, const float* __restrict__ qr,
        const float* __restrict__ qmin, const float* __restrict__ qmax, const int* __restrict__ q_start,
        const afv_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, const float* __restrict__ kpsize,
        const int* __restrict__ n_arr, int cap, const int* __restrict__ frame, const float* __restrict__ tinf1d,
        float minX, float minY, float invW, float invH, void* __restrict__ pool_v, int pool_cap, ProjQMeta* __restrict__ qmeta) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int fb = frame[p];
    const float* inf2 = tinf1d ? tinf1d + (long long)fb * cap : nullptr;
    const int n2 = min(n_arr[fb], cap);
    const afv_keypoint* k2 = kps + (long long)fb * cap;
    const uint8_t* d2 = desc + (long long)fb * cap * D;
    const float* size2 = kpsize + (long long)fb * cap;
    float* sx = reinterpret_cast<float*>(sm); float* sy = sx + cap; float* ssz = sy + cap;
    unsigned short* scell = reinterpret_cast<unsigned short*>(ssz + cap);
    unsigned short* sorig = scell + cap; unsigned short* tmpc = sorig + cap;
    uint8_t* sdesc = sm + prl_base_bytes(cap);
    __shared__ int colstart[AFV_GRID_COLS + 1], colfill[AFV_GRID_COLS];
    __shared__ int s_pool;
    const int nw = Dpad / 4;
    const unsigned short NONE16 = 0xffff;
    if (tid < AFV_GRID_COLS) { colstart[tid] = 0; colfill[tid] = 0; }
    if (tid == 0) { s_pool = 0; colstart[AFV_GRID_COLS] = 0; }
    __syncthreads();
    for (int i = tid; i < n2; i += SFL_THREADS) {
        const int c = grid_cell(k2[i].x, k2[i].y, minX, minY, invW, invH);
        tmpc[i] = c >= 0 ? (unsigned short)c : NONE16;
        if (c >= 0) atomicAdd(&colstart[c / AFV_GRID_ROWS + 1], 1);
    }
    __syncthreads();
    if (wid == 0) {
        int a0 = colstart[1 + lane], a1 = colstart[33 + lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t0 = __shfl_up_sync(0xffffffffu, a0, o), t1 = __shfl_up_sync(0xffffffffu, a1, o); if (lane >= o) { a0 += t0; a1 += t1; } }
        const int tot0 = __shfl_sync(0xffffffffu, a0, 31);
        colstart[1 + lane] = a0; colstart[33 + lane] = tot0 + a1;
    }
    __syncthreads();
    for (int i = tid; i < n2; i += SFL_THREADS) {
        const int c = tmpc[i];
        if (c == NONE16) continue;
        const int cx = c / AFV_GRID_ROWS;
        const int slot = colstart[cx] + atomicAdd(&colfill[cx], 1);
        sx[slot] = k2[i].x; sy[slot] = k2[i].y; ssz[slot] = size2[i];
        scell[slot] = (unsigned short)((cx << 8) | (c % AFV_GRID_ROWS));
        sorig[slot] = (unsigned short)i;
    }
    __syncthreads();
    const int n2s = colstart[AFV_GRID_COLS];
    if (stage_desc) {
        for (int i = tid; i < n2s * nw; i += SFL_THREADS) {
            const int r = i / nw, w = i % nw;
            const uint8_t* row = d2 + (long long)sorig[r] * D;
            uint32_t v;
            if ((D & 3) == 0) v = reinterpret_cast<const uint32_t*>(row)[w];
            else { v = 0; for (int b = 0; b < 4; ++b) { const int o = w * 4 + b; if (o < D) v |= (uint32_t)row[o] << (8 * b); } }
            reinterpret_cast<uint32_t*>(sdesc)[(size_t)w * cap + r] = v;
        }
    }
    __syncthreads();
    const int q0 = q_start[p], q1 = q_start[p + 1];
    for (int qi = q0 + wid; qi < q1; qi += SFL_THREADS / 32) {
        SfiQuery q;
        q.x = qxy[2 * qi]; q.y = qxy[2 * qi + 1];
        const float r = qr[qi], smin = qmin[qi], smax = qmax[qi];
        q.ok = !(r < 0.0f) && window_cells(q.x, q.y, r, minX, minY, invW, invH, q.c0, q.c1, q.r0, q.r1);   // r < 0: query skipped by the caller's prologue
        int cnt = 0, off = 0;
        if (q.ok) {
            const int s0 = colstart[q.c0], s1 = colstart[q.c1 + 1];
            auto is_cand = [&](int sl) {
                const unsigned short cc = scell[sl];
                const float sz = ssz[sl];
                if (!(!(sz < smin) && !(sz > smax) && sfi_in_window(q, cc >> 8, cc & 0xff, sx[sl], sy[sl], r))) return false;
                if (inf2) {                                      // Fuse, monocular reprojection gate (src/FeatureMatcher.cc:905-915)
                    const float ex = __fsub_rn(q.x, sx[sl]), ey = __fsub_rn(q.y, sy[sl]);
                    const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                    if ((double)__fmul_rn(e2, inf2[sorig[sl]]) > 5.99) return false;
                }
                return true;
            };
            for (int sb = s0; sb < s1; sb += 32) {
                const int sl = sb + lane;
                cnt += __popc(__ballot_sync(0xffffffffu, sl < s1 && is_cand(sl)));
            }
            if (cnt > 0) {
                if (lane == 0) off = atomicAdd(&s_pool, cnt);
                off = __shfl_sync(0xffffffffu, off, 0);
                if (off + cnt > pool_cap) off = -1;
            }
            if (cnt > 0 && off >= 0) {
                uint32_t qd[16];
                const uint8_t* qrow = qdesc + (long long)qi * D;
                if (BINARY) {
                    if ((D & 3) == 0 && ((uintptr_t)qrow & 3) == 0) {
#pragma unroll
                        for (int w = 0; w < 16; ++w) if (w < nw) qd[w] = reinterpret_cast<const uint32_t*>(qrow)[w];
                    } else {
#pragma unroll
                        for (int w = 0; w < 16; ++w) if (w < nw) { uint32_t v = 0; for (int b = 0; b < 4; ++b) { const int o = w * 4 + b; if (o < D) v |= (uint32_t)qrow[o] << (8 * b); } qd[w] = v; }
                    }
                }
                int run = off;
                for (int sb = s0; sb < s1; sb += 32) {
                    const int sl = sb + lane;
                    const bool pass = sl < s1 && is_cand(sl);
                    const unsigned m = __ballot_sync(0xffffffffu, pass);
                    if (pass) {
                        const int i2 = sorig[sl];
                        const int pos = run + __popc(m & ((1u << lane) - 1));
                        if (BINARY) {
                            int d = 0;
                            if (stage_desc) {
                                const uint32_t* y32 = reinterpret_cast<const uint32_t*>(sdesc) + sl;
#pragma unroll
                                for (int w = 0; w < 16; ++w) if (w < nw) d += __popc(qd[w] ^ y32[(size_t)w * cap]);
                            } else d = hamming_bytes(qrow, d2 + (long long)i2 * D, D);
                            reinterpret_cast<uint32_t*>(pool_v)[(long long)p * pool_cap + pos] = ((uint32_t)d << 20) | (uint32_t)i2;
                        } else {
                            const float dist = l2sqr128((const float*)qrow, (const float*)(d2 + (long long)i2 * D));
                            reinterpret_cast<unsigned long long*>(pool_v)[(long long)p * pool_cap + pos] = make_key(dist, (uint32_t)i2);
                        }
                    }
                    run += __popc(m);
                }
            }
        }
        if (lane == 0) { ProjQMeta m; m.off = off; m.cnt = cnt; qmeta[qi] = m; }
    }
}

template <bool BINARY>
__global__ void __launch_bounds__(SFR_WARPS * 32) k_proj_resolve(int desc_type, int D,
        const uint8_t* __restrict__ qdesc, const float* __restrict__ qxy, const float* __restrict__ qr,
        const float* __restrict__ qmin, const float* __restrict__ qmax, const int* __restrict__ q_start, int P,
        const afv_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, const float* __restrict__ kpsize,
        const int* __restrict__ n_arr, int cap, const int* __restrict__ frame, const uint8_t* __restrict__ occupied_in,
        const float* __restrict__ tinf1d, const float* __restrict__ qangle, int claim,
        float minX, float minY, float invW, float invH, float th, float nnratio, int ratio_same_scale, float tol,
        const void* __restrict__ pool_v, int pool_cap, ProjQMeta* __restrict__ qmeta,
        int* __restrict__ match_q, int* __restrict__ nmatches) {
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ int s_hist[SFR_WARPS][32];
    __shared__ int s_keep[SFR_WARPS][3];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int p = blockIdx.x * SFR_WARPS + wid;
    if (p >= P) return;
    unsigned char* base = sm + (size_t)wid * prr_warp_bytes(cap);
    float* tsize = reinterpret_cast<float*>(base);
    unsigned short* tcell = reinterpret_cast<unsigned short*>(tsize + cap);
    unsigned char* occ = reinterpret_cast<unsigned char*>(tcell + cap);
    const int fb = frame[p];
    const int n2 = min(n_arr[fb], cap);
    const afv_keypoint* k2 = kps + (long long)fb * cap;
    const uint8_t* d2 = desc + (long long)fb * cap * D;
    const float* size2 = kpsize + (long long)fb * cap;
    const unsigned short NONE16 = 0xffff;
    for (int i = lane; i < n2; i += 32) {
        tsize[i] = size2[i];
        const int c = grid_cell(k2[i].x, k2[i].y, minX, minY, invW, invH);
        tcell[i] = c >= 0 ? (unsigned short)(((c / AFV_GRID_ROWS) << 8) | (c % AFV_GRID_ROWS)) : NONE16;
        occ[i] = occupied_in ? occupied_in[(long long)fb * cap + i] : 0;
    }
    s_hist[wid][lane] = 0;
    __syncwarp();
    const float invtol = __fdiv_rn(1.0f, tol);
    const float* inf2 = tinf1d ? tinf1d + (long long)fb * cap : nullptr;
    int nm = 0;
    const int q0 = q_start[p], q1 = q_start[p + 1];
    for (int qi = q0; qi < q1; ++qi) {
        const ProjQMeta q = qmeta[qi];
        int bestIdx = -1;
        if (q.cnt > 0) {
            Top2 t; t.k1 = t.k2 = KEY_NONE;
            if (q.off >= 0) {
                for (int j0 = 0; j0 < q.cnt; j0 += 32) {
                    const int j = j0 + lane;
                    if (j >= q.cnt) continue;
                    unsigned long long key;
                    if (BINARY) {
                        const uint32_t e = reinterpret_cast<const uint32_t*>(pool_v)[(long long)p * pool_cap + q.off + j];
                        key = ((unsigned long long)__float_as_uint((float)(e >> 20)) << 32) | (e & 0xfffffu);
                    } else key = reinterpret_cast<const unsigned long long*>(pool_v)[(long long)p * pool_cap + q.off + j];
                    const int i2 = (int)((uint32_t)key & 0xfffffu);
                    if (occ[i2]) continue;                                  // F.pts[idx] already holds a map point
                    const unsigned short cc = tcell[i2];
                    top2_push(t, make_key(key_dist(key), ((uint32_t)(cc >> 8) << 26) | ((uint32_t)(cc & 0xff) << 20) | (uint32_t)i2));
                }
            } else {
                // exact fallback (pool exhausted): rescan the train frame with the same candidate definition
                SfiQuery w;
                w.x = qxy[2 * qi]; w.y = qxy[2 * qi + 1];
                const float r = qr[qi], smin = qmin[qi], smax = qmax[qi];
                w.ok = !(r < 0.0f) && window_cells(w.x, w.y, r, minX, minY, invW, invH, w.c0, w.c1, w.r0, w.r1);
                for (int i2 = lane; w.ok && i2 < n2; i2 += 32) {
                    const unsigned short cc = tcell[i2];
                    if (cc == NONE16 || occ[i2] || tsize[i2] < smin || tsize[i2] > smax) continue;
                    if (!sfi_in_window(w, cc >> 8, cc & 0xff, k2[i2].x, k2[i2].y, r)) continue;
                    if (inf2) {
                        const float ex = __fsub_rn(w.x, k2[i2].x), ey = __fsub_rn(w.y, k2[i2].y);
                        if ((double)__fmul_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), inf2[i2]) > 5.99) continue;
                    }
                    const float dist = desc_distance(desc_type, qdesc + (long long)qi * D, d2 + (long long)i2 * D, D);
                    top2_push(t, make_key(dist, ((uint32_t)(cc >> 8) << 26) | ((uint32_t)(cc & 0xff) << 20) | (uint32_t)i2));
                }
            }
            top2_warp_reduce(t);
            if (t.k1 != KEY_NONE) {
                const float bestDist = key_dist(t.k1), bestDist2 = key_dist(t.k2);
                const int bi = (int)((uint32_t)t.k1 & 0xfffffu);
                if (bestDist <= th) {
                    bool accept = true;
                    if (ratio_same_scale) {                                 // :139-146
                        const float bestSize = tsize[bi];
                        const float bestSize2 = t.k2 != KEY_NONE ? tsize[(uint32_t)t.k2 & 0xfffffu] : -1.0f;
                        const float ratio = __fdiv_rn(bestSize, bestSize2);
                        if (ratio < tol && ratio > invtol && bestSize2 > 0.0f)
                            if (bestDist > __fmul_rn(nnratio, bestDist2)) accept = false;
                    }
                    if (accept) bestIdx = bi;
                }
            }
        }
        if (bestIdx >= 0) {
            ++nm;
            if (lane == 0) {
                if (claim) occ[bestIdx] = 1;
                if (qangle) {                                    // updateRotationHistogram(rotHist, bestIdx, query keypoint, train keypoint)
                    const int bin = rot_bin(qangle[qi], k2[bestIdx].angle);
                    s_hist[wid][bin]++; qmeta[qi].off = bin;     // the list offset of this query is not needed any more
                }
            }
        }
        if (lane == 0) match_q[qi] = bestIdx;
        __syncwarp();
    }
    if (qangle) {                                                // filterMatchesWithOrientation (:1601-1612)
        if (lane == 0) three_maxima(s_hist[wid], s_keep[wid][0], s_keep[wid][1], s_keep[wid][2]);
        __syncwarp();
        int removed = 0;
        for (int qi = q0 + lane; qi < q1; qi += 32) {
            if (match_q[qi] < 0) continue;
            const int bn = qmeta[qi].off;
            if (bn == s_keep[wid][0] || bn == s_keep[wid][1] || bn == s_keep[wid][2]) continue;
            match_q[qi] = -1; ++removed;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nm -= removed;
    }
    if (lane == 0) nmatches[p] = nm;
}

static const int PROJ_POOL_CAP = 64 * 1024;
extern "C" size_t afv_search_by_projection_workspace_bytes(int desc_type, int P, int nq_total) {
    const size_t esz = desc_type != AFV_FEAT_SIFT128 ? 4 : 8;
    return (size_t)(P > 0 ? P : 0) * PROJ_POOL_CAP * esz + (size_t)(nq_total > 0 ? nq_total : 0) * sizeof(ProjQMeta) + 256;
}

extern "C" int afv_search_by_projection_ex(int desc_type, const void* d_qdesc, const float* d_qxy, const float* d_qr,
        const float* d_qmin_size, const float* d_qmax_size, const float* d_qangle, const int* d_q_start, int P, int nq_total,
        const afv_keypoint* d_kps, const void* d_desc, const float* d_kpsize, const float* d_inf1d, const int* d_n, int B, int cap,
        const int* d_frame, const uint8_t* d_occupied, int claim, float min_x, float min_y, float max_x, float max_y, float th,
        float nnratio, int ratio_same_scale_only, float size_tolerance, int* d_match_q, int* d_nmatches,
        void* d_workspace, size_t workspace_bytes, void* cuda_stream) {
    const int D = desc_bytes(desc_type);
    if (D < 0 || !d_qdesc || !d_qxy || !d_qr || !d_qmin_size || !d_qmax_size || !d_q_start || !d_kps || !d_desc || !d_kpsize || !d_n ||
        !d_frame || !d_match_q || !d_nmatches || B < 1 || P < 0 || cap < 1 || !(size_tolerance > 0.0f)) {
        afv_set_error("afv_search_by_projection: bad argument"); return AFV_ERR_INVALID;
    }
    if (P == 0) return AFV_OK;
    if (cap >= 65535) { afv_set_error("cap too large for the 16-bit indices"); return AFV_ERR_INVALID; }
    cudaStream_t st = as_stream(cuda_stream);
    const bool binary = desc_type != AFV_FEAT_SIFT128;
    const int Dpad = (D + 3) & ~3;
    const size_t baseA = prl_base_bytes(cap);
    const int stage = binary && Dpad <= 64 && baseA + (size_t)cap * Dpad <= 200 * 1024;
    const size_t smemA = baseA + (stage ? (size_t)cap * Dpad : 0);
    const size_t smemB = prr_warp_bytes(cap) * SFR_WARPS;
    if (smemA > 220 * 1024 || smemB > 220 * 1024) { afv_set_error("cap %d too large for the shared-memory staging", cap); return AFV_ERR_INVALID; }
    // the opt-in shared-memory size is a per-device function attribute: set it on every call (cheap, idempotent, thread-safe)
    AFV_CUDA_CHECK(binary ? cudaFuncSetAttribute(k_proj_lists<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA)
                          : cudaFuncSetAttribute(k_proj_lists<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
    AFV_CUDA_CHECK(binary ? cudaFuncSetAttribute(k_proj_resolve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB)
                          : cudaFuncSetAttribute(k_proj_resolve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB));
    if (nq_total < 0) {                               // query count unknown to the host: read it back (the only synchronising path)
        AFV_CUDA_CHECK(cudaMemcpyAsync(&nq_total, d_q_start + P, sizeof(int), cudaMemcpyDeviceToHost, st));
        AFV_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    if (nq_total <= 0) { AFV_CUDA_CHECK(cudaMemsetAsync(d_nmatches, 0, sizeof(int) * P, st)); return AFV_OK; }
    const int pool_cap = PROJ_POOL_CAP;
    const size_t esz = binary ? 4 : 8;
    const size_t pool_bytes = (size_t)P * pool_cap * esz, need = afv_search_by_projection_workspace_bytes(desc_type, P, nq_total);
    unsigned char* scratch = (unsigned char*)d_workspace;
    if (scratch && workspace_bytes < need) { afv_set_error("afv_search_by_projection: workspace %zu B < required %zu B", workspace_bytes, need); return AFV_ERR_INVALID; }
    if (!scratch) AFV_CUDA_CHECK(cudaMallocAsync((void**)&scratch, need, st));
    void* pool = scratch;
    ProjQMeta* qmeta = reinterpret_cast<ProjQMeta*>(scratch + pool_bytes);
    const float invW = (float)AFV_GRID_COLS / (max_x - min_x), invH = (float)AFV_GRID_ROWS / (max_y - min_y);
    {
        AfvProfScope ps("k_proj_lists", st);
        if (binary) k_proj_lists<true><<<P, SFL_THREADS, smemA, st>>>(desc_type, D, Dpad, stage, (const uint8_t*)d_qdesc, d_qxy, d_qr, d_qmin_size, d_qmax_size,
                d_q_start, d_kps, (const uint8_t*)d_desc, d_kpsize, d_n, cap, d_frame, d_inf1d, min_x, min_y, invW, invH, pool, pool_cap, qmeta);
        else k_proj_lists<false><<<P, SFL_THREADS, smemA, st>>>(desc_type, D, Dpad, stage, (const uint8_t*)d_qdesc, d_qxy, d_qr, d_qmin_size, d_qmax_size,
                d_q_start, d_kps, (const uint8_t*)d_desc, d_kpsize, d_n, cap, d_frame, d_inf1d, min_x, min_y, invW, invH, pool, pool_cap, qmeta);
        ++g_afv_launches;
    }
    {
        AfvProfScope ps("k_proj_resolve", st);
        const int grid = (P + SFR_WARPS - 1) / SFR_WARPS;
        if (binary) k_proj_resolve<true><<<grid, SFR_WARPS * 32, smemB, st>>>(desc_type, D, (const uint8_t*)d_qdesc, d_qxy, d_qr, d_qmin_size, d_qmax_size, d_q_start, P,
                d_kps, (const uint8_t*)d_desc, d_kpsize, d_n, cap, d_frame, d_occupied, d_inf1d, d_qangle, claim, min_x, min_y, invW, invH, th, nnratio,
                ratio_same_scale_only, size_tolerance, pool, pool_cap, qmeta, d_match_q, d_nmatches);
        else k_proj_resolve<false><<<grid, SFR_WARPS * 32, smemB, st>>>(desc_type, D, (const uint8_t*)d_qdesc, d_qxy, d_qr, d_qmin_size, d_qmax_size, d_q_start, P,
                d_kps, (const uint8_t*)d_desc, d_kpsize, d_n, cap, d_frame, d_occupied, d_inf1d, d_qangle, claim, min_x, min_y, invW, invH, th, nnratio,
                ratio_same_scale_only, size_tolerance, pool, pool_cap, qmeta, d_match_q, d_nmatches);
        ++g_afv_launches;
    }
    AFV_CUDA_CHECK(cudaGetLastError());
    if (!d_workspace) AFV_CUDA_CHECK(cudaFreeAsync(scratch, st));
    return AFV_OK;
}

extern "C" int afv_search_by_projection(int desc_type, const void* d_qdesc, const float* d_qxy, const float* d_qr,
        const float* d_qmin_size, const float* d_qmax_size, const int* d_q_start, int P,
        const afv_keypoint* d_kps, const void* d_desc, const float* d_kpsize, const int* d_n, int B, int cap, const int* d_frame,
        const uint8_t* d_occupied, float min_x, float min_y, float max_x, float max_y, float th, float nnratio,
        int ratio_same_scale_only, float size_tolerance, int* d_match_q, int* d_nmatches, void* cuda_stream) {
    return afv_search_by_projection_ex(desc_type, d_qdesc, d_qxy, d_qr, d_qmin_size, d_qmax_size, nullptr, d_q_start, P, -1, d_kps, d_desc, d_kpsize,
                                       nullptr, d_n, B, cap, d_frame, d_occupied, 1, min_x, min_y, max_x, max_y, th, nnratio, ratio_same_scale_only,
                                       size_tolerance, d_match_q, d_nmatches, nullptr, 0, cuda_stream);
}

// ---- SearchBySim3 (src/FeatureMatcher.cc:1066-1287): two stateless directed searches + agreement -------------------------
__global__ void k_sim3_agree(const int* __restrict__ m1, const int* __restrict__ q_start1, const int* __restrict__ m2,
                             const int* __restrict__ q_start2, int P, int* __restrict__ match12, int* __restrict__ nfound) {
    const int p = blockIdx.x;
    if (p >= P) return;
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int a0 = q_start1[p], a1 = q_start1[p + 1], b0 = q_start2[p], b1 = q_start2[p + 1];
    int local = 0;
    for (int i = a0 + threadIdx.x; i < a1; i += blockDim.x) {
        const int idx2 = m1[i];
        int out = -1;
        if (idx2 >= 0 && b0 + idx2 < b1 && m2[b0 + idx2] == i - a0) { out = idx2; ++local; }
        match12[i] = out;
    }
    atomicAdd(&s_cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) nfound[p] = s_cnt;
}

extern "C" int afv_search_by_sim3(int desc_type,
        const void* d_q1desc, const float* d_q1xy, const float* d_q1r, const float* d_q1min, const float* d_q1max, const int* d_q_start1, int nq1_total,
        const void* d_q2desc, const float* d_q2xy, const float* d_q2r, const float* d_q2min, const float* d_q2max, const int* d_q_start2, int nq2_total,
        int P, const afv_keypoint* d_kps, const void* d_desc, const float* d_kpsize, const int* d_n, int B, int cap,
        const int* d_frame1, const int* d_frame2, float min_x, float min_y, float max_x, float max_y, float th_high,
        int* d_match12, int* d_nfound, void* cuda_stream) {
    if (P < 0 || nq1_total < 0 || nq2_total < 0 || !d_match12 || !d_nfound || !d_q_start1 || !d_q_start2) { afv_set_error("afv_search_by_sim3: bad argument"); return AFV_ERR_INVALID; }
    if (P == 0) return AFV_OK;
    cudaStream_t st = as_stream(cuda_stream);
    int* tmp = nullptr;
    AFV_CUDA_CHECK(cudaMallocAsync((void**)&tmp, sizeof(int) * ((size_t)nq1_total + nq2_total + 2 * (size_t)P + 4), st));
    int* m1 = tmp; int* m2 = tmp + nq1_total; int* c1 = m2 + nq2_total; int* c2 = c1 + P;
    // direction 1 -> 2: the map points of frame1's keypoints searched in frame2, and vice versa; no occupied flags, no claims
    int rc = afv_search_by_projection_ex(desc_type, d_q1desc, d_q1xy, d_q1r, d_q1min, d_q1max, nullptr, d_q_start1, P, nq1_total, d_kps, d_desc, d_kpsize,
                                         nullptr, d_n, B, cap, d_frame2, nullptr, 0, min_x, min_y, max_x, max_y, th_high, 1.0f, 0, 1.0f, m1, c1, nullptr, 0, cuda_stream);
    if (rc == AFV_OK)
        rc = afv_search_by_projection_ex(desc_type, d_q2desc, d_q2xy, d_q2r, d_q2min, d_q2max, nullptr, d_q_start2, P, nq2_total, d_kps, d_desc, d_kpsize,
                                         nullptr, d_n, B, cap, d_frame1, nullptr, 0, min_x, min_y, max_x, max_y, th_high, 1.0f, 0, 1.0f, m2, c2, nullptr, 0, cuda_stream);
    if (rc == AFV_OK) {
        if (nq1_total == 0) AFV_CUDA_CHECK(cudaMemsetAsync(d_nfound, 0, sizeof(int) * P, st));
        else { k_sim3_agree<<<P, 256, 0, st>>>(m1, d_q_start1, m2, d_q_start2, P, d_match12, d_nfound); ++g_afv_launches; AFV_CUDA_CHECK(cudaGetLastError()); }
    }
    AFV_CUDA_CHECK(cudaFreeAsync(tmp, st));
    return rc;
}

// ---- BoW merge-join searches on per-feature node ids, batched over frame pairs --------------------------------------------
// mode 0 SearchByBoW(KF, F) (:186-283), 1 SearchByBoW(KF, KF) (:561-660), 2 SearchForTriangulation (:662-790, monocular).
// One CTA per pair: both frames' (node, feature) keys are bitonic-sorted in shared memory (= DBoW2's FeatureVector order: nodes
// ascending, features of a node in increasing index); the sequential "already matched" state of the reference only couples the
// features of ONE node, so the shared nodes are independent and go to the 8 warps round-robin; inside a node the frame-1
// features are taken in order and the lanes share the node's frame-2 features.
#define BOW_THREADS 256
template <int MODE>
__global__ void __launch_bounds__(BOW_THREADS) k_bow_match(int desc_type, int D, const afv_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc,
        const int* __restrict__ n_arr, int cap, int capp, const int* __restrict__ node_id, const uint8_t* __restrict__ valid,
        const int* __restrict__ pair_a, const int* __restrict__ pair_b, float th_low, float nnratio, int check_ori,
        const float* __restrict__ F12s, const float* __restrict__ epis, const float* __restrict__ sigma2,
        int* __restrict__ match, int* __restrict__ nmatches) {
    extern __shared__ __align__(16) unsigned char sm[];
    unsigned long long* key1 = reinterpret_cast<unsigned long long*>(sm);         // [capp]
    unsigned long long* key2 = key1 + capp;                                       // [capp]
    int* seg = reinterpret_cast<int*>(key2 + capp);                               // [capp + 1] segment starts of list 1
    unsigned char* matched2 = reinterpret_cast<unsigned char*>(seg + capp + 1);   // [cap]
    unsigned char* bin_of = matched2 + cap;                                       // [cap]
    __shared__ int hist[32], s_m1, s_m2, s_nseg, s_nm, keepbin[3];
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int fa = pair_a[p], fb = pair_b[p];
    const int n1 = min(n_arr[fa], cap), n2 = min(n_arr[fb], cap);
    const afv_keypoint* k1 = kps + (long long)fa * cap; const afv_keypoint* k2 = kps + (long long)fb * cap;
    const uint8_t* d1 = desc + (long long)fa * cap * D; const uint8_t* d2 = desc + (long long)fb * cap * D;
    const int* nd1 = node_id + (long long)fa * cap; const int* nd2 = node_id + (long long)fb * cap;
    const uint8_t* v1 = valid ? valid + (long long)fa * cap : nullptr; const uint8_t* v2 = valid ? valid + (long long)fb * cap : nullptr;
    int* out = match + (long long)p * cap;
    if (tid < 32) hist[tid] = 0;
    if (tid == 0) { s_m1 = 0; s_m2 = 0; s_nseg = 0; s_nm = 0; }
    __syncthreads();
    int c1 = 0, c2 = 0;
    for (int i = tid; i < capp; i += BOW_THREADS) {
        unsigned long long a = ~0ull, b = ~0ull;
        if (i < n1 && nd1[i] >= 0) { a = ((unsigned long long)(unsigned)nd1[i] << 32) | (unsigned)i; ++c1; }
        if (i < n2 && nd2[i] >= 0) { b = ((unsigned long long)(unsigned)nd2[i] << 32) | (unsigned)i; ++c2; }
        key1[i] = a; key2[i] = b;
        if (i < cap) { matched2[i] = 0; bin_of[i] = 0xff; out[i] = -1; }
    }
    atomicAdd(&s_m1, c1); atomicAdd(&s_m2, c2);
    __syncthreads();
    for (int k = 2; k <= capp; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < capp; t += BOW_THREADS) {
                const int txj = t ^ j;
                if (txj > t) {
                    const bool up = (t & k) == 0;
                    unsigned long long a = key1[t], b = key1[txj];
                    if ((a > b) == up) { key1[t] = b; key1[txj] = a; }
                    a = key2[t]; b = key2[txj];
                    if ((a > b) == up) { key2[t] = b; key2[txj] = a; }
                }
            }
            __syncthreads();
        }
    const int m1 = s_m1, m2 = s_m2;
    // segment heads of list 1 (ordered compaction by one warp: m1 <= 4096 entries)
    if (wid == 0) {
        int base = 0;
        for (int t0 = 0; t0 < m1; t0 += 32) {
            const int t = t0 + lane;
            const bool head = t < m1 && (t == 0 || (key1[t] >> 32) != (key1[t - 1] >> 32));
            const unsigned m = __ballot_sync(0xffffffffu, head);
            if (head) seg[base + __popc(m & ((1u << lane) - 1))] = t;
            base += __popc(m);
        }
        if (lane == 0) { seg[base] = m1; s_nseg = base; }
    }
    __syncthreads();
    const int nseg = s_nseg;
    float F[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ex = 0.f, ey = 0.f;
    const float* sg2 = nullptr;
    if (MODE == 2) {
        for (int i = 0; i < 9; ++i) F[i] = F12s[(long long)p * 9 + i];
        ex = epis[2 * p]; ey = epis[2 * p + 1];
        sg2 = sigma2 + (long long)fb * cap;
    }
    int nm = 0;
    for (int s = wid; s < nseg; s += BOW_THREADS / 32) {
        const int a0 = seg[s], a1 = seg[s + 1];
        const unsigned node = (unsigned)(key1[a0] >> 32);
        // lower bound of (node, 0) in list 2
        int lo = 0, hi = m2;
        const unsigned long long target = (unsigned long long)node << 32;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (key2[mid] < target) lo = mid + 1; else hi = mid; }
        if (lo >= m2 || (unsigned)(key2[lo] >> 32) != node) continue;
        const int b0 = lo;
        int b1 = b0;
        { int l2 = b0, h2 = m2; const unsigned long long t2 = ((unsigned long long)node + 1ull) << 32;
          while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (key2[mid] < t2) l2 = mid + 1; else h2 = mid; }
          b1 = l2; }
        for (int ia = a0; ia < a1; ++ia) {
            const int idx1 = (int)(unsigned)(key1[ia] & 0xffffffffu);
            if (MODE == 2) { if (v1 && v1[idx1]) continue; }
            else if (v1 && !v1[idx1]) continue;
            const uint8_t* ref = d1 + (long long)idx1 * D;
            Top2 t; t.k1 = t.k2 = KEY_NONE;
            for (int ib = b0 + lane; ib < b1; ib += 32) {
                const int idx2 = (int)(unsigned)(key2[ib] & 0xffffffffu);
                if (MODE == 2) {
                    if (v2 && v2[idx2]) continue;
                    const float dist = desc_distance(desc_type, ref, d2 + (long long)idx2 * D, D);
                    if (dist > th_low) continue;
                    const float distex = __fsub_rn(ex, k2[idx2].x), distey = __fsub_rn(ey, k2[idx2].y);
                    if (__fadd_rn(__fmul_rn(distex, distex), __fmul_rn(distey, distey)) < __fmul_rn(100.0f, __fsqrt_rn(sg2[idx2]))) continue;
                    const float x1 = k1[idx1].x, y1 = k1[idx1].y, x2 = k2[idx2].x, y2 = k2[idx2].y;
                    const float la = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[0]), __fmul_rn(y1, F[3])), F[6]);
                    const float lb = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[1]), __fmul_rn(y1, F[4])), F[7]);
                    const float lc = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[2]), __fmul_rn(y1, F[5])), F[8]);
                    const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, x2), __fmul_rn(lb, y2)), lc);
                    const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
                    if (den == 0.0f) continue;
                    const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
                    if (!(dsqr < __fmul_rn(3.84f, sg2[idx2]))) continue;
                    // smallest distance wins, the LAST of equal distances (candidates with dist > best are skipped, :733)
                    top2_push(t, make_key(dist, (uint32_t)(b1 - 1 - ib)));
                } else {
                    if (matched2[idx2]) continue;
                    if (MODE == 1 && v2 && !v2[idx2]) continue;
                    top2_push(t, make_key(desc_distance(desc_type, ref, d2 + (long long)idx2 * D, D), (uint32_t)(ib - b0)));
                }
            }
            top2_warp_reduce(t);
            if (lane == 0 && t.k1 != KEY_NONE) {
                const float bd1 = key_dist(t.k1), bd2 = key_dist(t.k2);
                if (MODE == 2) {
                    const int ib = b1 - 1 - (int)(uint32_t)t.k1;
                    out[idx1] = (int)(unsigned)(key2[ib] & 0xffffffffu); ++nm;
                } else {
                    const bool pass_th = MODE == 0 ? (bd1 <= th_low) : (bd1 < th_low);
                    if (pass_th && bd1 < __fmul_rn(nnratio, bd2)) {
                        const int best2 = (int)(unsigned)(key2[b0 + (int)(uint32_t)t.k1] & 0xffffffffu);
                        matched2[best2] = 1; ++nm;
                        const int oi = MODE == 0 ? best2 : idx1;
                        out[oi] = MODE == 0 ? idx1 : best2;
                        if (check_ori) { const int bin = rot_bin(k1[idx1].angle, k2[best2].angle); bin_of[oi] = (unsigned char)bin; atomicAdd(&hist[bin], 1); }
                    }
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0 && nm) atomicAdd(&s_nm, nm);
    __syncthreads();
    if (check_ori && MODE != 2) {
        if (tid == 0) three_maxima(hist, keepbin[0], keepbin[1], keepbin[2]);
        __syncthreads();
        int removed = 0;
        const int nout = MODE == 0 ? n2 : n1;
        for (int i = tid; i < nout; i += BOW_THREADS) {
            const int bn = bin_of[i];
            if (bn == 0xff || bn == keepbin[0] || bn == keepbin[1] || bn == keepbin[2]) continue;
            out[i] = -1; ++removed;
        }
        if (removed) atomicSub(&s_nm, removed);
        __syncthreads();
    }
    if (tid == 0) nmatches[p] = s_nm;
}

extern "C" int afv_bow_match(int mode, int desc_type, const afv_keypoint* d_kps, const void* d_desc, const int* d_n, int B, int cap,
        const int* d_node_id, const uint8_t* d_valid, const int* d_pair_a, const int* d_pair_b, int P, float th_low, float nnratio,
        int check_orientation, const float* d_F12, const float* d_epipole, const float* d_sigma2, int* d_match, int* d_nmatches, void* cuda_stream) {
    const int D = desc_bytes(desc_type);
    if (D < 0 || mode < 0 || mode > 2 || !d_kps || !d_desc || !d_n || !d_node_id || !d_pair_a || !d_pair_b || !d_match || !d_nmatches || B < 1 || P < 0 || cap < 1 ||
        (mode == 2 && (!d_F12 || !d_epipole || !d_sigma2))) { afv_set_error("afv_bow_match: bad argument"); return AFV_ERR_INVALID; }
    if (P == 0) return AFV_OK;
    int capp = 1;
    while (capp < cap) capp <<= 1;
    const size_t smem = (size_t)capp * 16 + (size_t)(capp + 1) * 4 + (size_t)cap * 2 + 16;
    if (smem > 200 * 1024) { afv_set_error("afv_bow_match: cap %d too large for the shared-memory sort", cap); return AFV_ERR_INVALID; }
    cudaStream_t st = as_stream(cuda_stream);
    AfvProfScope ps("k_bow_match", st);
#define BOW_LAUNCH(M) do { AFV_CUDA_CHECK(cudaFuncSetAttribute(k_bow_match<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_bow_match<M><<<P, BOW_THREADS, smem, st>>>(desc_type, D, d_kps, (const uint8_t*)d_desc, d_n, cap, capp, d_node_id, d_valid, d_pair_a, d_pair_b, \
            th_low, nnratio, check_orientation, d_F12, d_epipole, d_sigma2, d_match, d_nmatches); } while (0)
    if (mode == 0) BOW_LAUNCH(0); else if (mode == 1) BOW_LAUNCH(1); else BOW_LAUNCH(2);
#undef BOW_LAUNCH
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

// ---- brute force N x M: one warp per query, lanes stride over train descriptors ----------------------------
__global__ void __launch_bounds__(256) k_match_bf(int desc_type, int D, const uint8_t* __restrict__ q, int nq,
        const uint8_t* __restrict__ t, int nt, int* __restrict__ best, float* __restrict__ bestd, float* __restrict__ secondd) {
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (qi >= nq) return;
    Top2 tt; tt.k1 = tt.k2 = KEY_NONE;
    const uint8_t* qd = q + (long long)qi * D;
    if (desc_type == AFV_FEAT_ORB32 && (((uintptr_t)q | (uintptr_t)t) & 15) == 0) {
        const uint4 a0 = reinterpret_cast<const uint4*>(qd)[0], a1 = reinterpret_cast<const uint4*>(qd)[1];
        for (int j = lane; j < nt; j += 32) {
            const uint4 b0 = reinterpret_cast<const uint4*>(t + (long long)j * 32)[0], b1 = reinterpret_cast<const uint4*>(t + (long long)j * 32)[1];
            const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                          __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            top2_push(tt, make_key((float)d, (uint32_t)j));
        }
    } else {
        for (int j = lane; j < nt; j += 32)
            top2_push(tt, make_key(desc_distance(desc_type, qd, t + (long long)j * D, D), (uint32_t)j));
    }
    top2_warp_reduce(tt);
    if (lane == 0) {
        best[qi] = tt.k1 == KEY_NONE ? -1 : (int)(uint32_t)tt.k1;
        bestd[qi] = key_dist(tt.k1); secondd[qi] = key_dist(tt.k2);
    }
}

// Tiled Hamming brute force: a thread owns one query (descriptor words in registers); the train descriptors pass through shared
// memory in tiles of BF_TILE rows, every lane reads the SAME row (128-bit broadcast loads), train index ascending so the first
// minimum wins without tie handling.  Runs at 96 % of the chip's POPC rate (148 SMs x 16 lanes x clock; bench.py --workload m1).
#define BF_TILE 256
template <int NW>
__global__ void __launch_bounds__(256) k_match_bf_tiled(int D, const uint8_t* __restrict__ qbase, int nq_one, const uint8_t* __restrict__ tbase, int nt_one,
        const int* __restrict__ n_arr, int cap, const int* __restrict__ pair_a, const int* __restrict__ pair_b,
        int* __restrict__ best, float* __restrict__ bestd, float* __restrict__ secondd) {
    __shared__ __align__(16) uint32_t tile[BF_TILE][NW];
    const int p = blockIdx.y, tid = threadIdx.x;
    const uint8_t* qd = qbase; const uint8_t* td = tbase;
    int nq = nq_one, nt = nt_one;
    long long oo = 0;
    if (pair_a) {
        const int fa = pair_a[p], fb = pair_b[p];
        qd = qbase + (long long)fa * cap * D; td = qbase + (long long)fb * cap * D;
        nq = min(n_arr[fa], cap); nt = min(n_arr[fb], cap); oo = (long long)p * cap;
    }
    const int qi = blockIdx.x * 256 + tid;
    if (blockIdx.x * 256 >= nq) return;
    uint32_t q[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) q[w] = 0;
    if (qi < nq) {
        if ((D & 3) == 0 && (((uintptr_t)qd) & 3) == 0) {
            const uint32_t* row = reinterpret_cast<const uint32_t*>(qd + (long long)qi * D);
#pragma unroll
            for (int w = 0; w < NW; ++w) q[w] = row[w];
        } else {
#pragma unroll
            for (int w = 0; w < NW; ++w) { uint32_t v = 0; for (int b = 0; b < 4; ++b) { const int o = w * 4 + b; if (o < D) v |= (uint32_t)qd[(long long)qi * D + o] << (8 * b); } q[w] = v; }
        }
    }
    int bi = -1, b1 = 0x7fffffff, b2 = 0x7fffffff;
    for (int t0 = 0; t0 < nt; t0 += BF_TILE) {
        const int nrow = min(BF_TILE, nt - t0);
        __syncthreads();
        if ((D & 3) == 0 && (((uintptr_t)td) & 3) == 0) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(td + (long long)t0 * D);
            for (int i = tid; i < nrow * NW; i += 256) (&tile[0][0])[i] = src[i];
        } else {
            for (int i = tid; i < nrow * NW; i += 256) {
                const int r = i / NW, w = i - r * NW;
                uint32_t v = 0;
                for (int b = 0; b < 4; ++b) { const int o = w * 4 + b; if (o < D) v |= (uint32_t)td[(long long)(t0 + r) * D + o] << (8 * b); }
                tile[r][w] = v;
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < nrow; ++j) {
            int d = 0;
#pragma unroll
            for (int w4 = 0; w4 < NW / 4; ++w4) {
                const uint4 v = *reinterpret_cast<const uint4*>(&tile[j][4 * w4]);
                d += __popc(q[4 * w4] ^ v.x) + __popc(q[4 * w4 + 1] ^ v.y) + __popc(q[4 * w4 + 2] ^ v.z) + __popc(q[4 * w4 + 3] ^ v.w);
            }
            if (d < b1) { b2 = b1; b1 = d; bi = t0 + j; } else if (d < b2) b2 = d;
        }
    }
    if (qi < nq) {
        best[oo + qi] = bi;
        bestd[oo + qi] = b1 == 0x7fffffff ? FLT_MAX : (float)b1;
        secondd[oo + qi] = b2 == 0x7fffffff ? FLT_MAX : (float)b2;
    }
}

static int launch_bf_tiled(int D, const uint8_t* q, int nq, const uint8_t* t, int nt, const int* d_n, int cap, const int* pa,
                           const int* pb, int P, int* best, float* bestd, float* secondd, cudaStream_t st) {
    const int NW = (D + 3) / 4;
    const dim3 grid(((pa ? cap : nq) + 255) / 256, P);
    AfvProfScope ps("k_match_bf", st);
    if (NW == 8) k_match_bf_tiled<8><<<grid, 256, 0, st>>>(D, q, nq, t, nt, d_n, cap, pa, pb, best, bestd, secondd);
    else if (NW == 12) k_match_bf_tiled<12><<<grid, 256, 0, st>>>(D, q, nq, t, nt, d_n, cap, pa, pb, best, bestd, secondd);
    else k_match_bf_tiled<16><<<grid, 256, 0, st>>>(D, q, nq, t, nt, d_n, cap, pa, pb, best, bestd, secondd);
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

extern "C" int afv_match_bruteforce(int desc_type, const void* d_q, int nq, const void* d_t, int nt,
                                    int* d_best, float* d_bestd, float* d_secondd, void* cuda_stream) {
    const int D = desc_bytes(desc_type);
    if (D < 0 || nq < 0 || nt < 0) { afv_set_error("afv_match_bruteforce: bad argument"); return AFV_ERR_INVALID; }
    if (nq == 0) return AFV_OK;
    if (!d_q || (!d_t && nt > 0) || !d_best || !d_bestd || !d_secondd) { afv_set_error("afv_match_bruteforce: NULL argument"); return AFV_ERR_INVALID; }
    if (desc_type != AFV_FEAT_SIFT128)
        return launch_bf_tiled(D, (const uint8_t*)d_q, nq, (const uint8_t*)d_t, nt, nullptr, 0, nullptr, nullptr, 1, d_best, d_bestd, d_secondd, as_stream(cuda_stream));
    k_match_bf<<<(nq + 7) / 8, 256, 0, as_stream(cuda_stream)>>>(desc_type, D, (const uint8_t*)d_q, nq, (const uint8_t*)d_t, nt, d_best, d_bestd, d_secondd);
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

// brute force for P frame pairs of a B x cap extraction result (binary descriptors): outputs are P x cap
extern "C" int afv_match_bruteforce_pairs(int desc_type, const void* d_desc, const int* d_n, int B, int cap, const int* d_pair_a,
                                          const int* d_pair_b, int P, int* d_best, float* d_bestd, float* d_secondd, void* cuda_stream) {
    const int D = desc_bytes(desc_type);
    if (D < 0 || desc_type == AFV_FEAT_SIFT128 || !d_desc || !d_n || !d_pair_a || !d_pair_b || !d_best || !d_bestd || !d_secondd || B < 1 || P < 0 || cap < 1) {
        afv_set_error("afv_match_bruteforce_pairs: bad argument (binary descriptors only)"); return AFV_ERR_INVALID;
    }
    if (P == 0) return AFV_OK;
    return launch_bf_tiled(D, (const uint8_t*)d_desc, 0, nullptr, 0, d_n, cap, d_pair_a, d_pair_b, P, d_best, d_bestd, d_secondd, as_stream(cuda_stream));
}

// ---- SearchByBoW(KF, F): one warp walks the merge-join; lanes share a bucket's frame features ---------------
__global__ void __launch_bounds__(32) k_search_bow(int desc_type, int D, const uint8_t* __restrict__ dkf,
        const afv_keypoint* __restrict__ kkf, const int* __restrict__ kf_node, const int* __restrict__ kf_start,
        const int* __restrict__ kf_idx, int kf_nodes, const uint8_t* __restrict__ df, const afv_keypoint* __restrict__ kf_f,
        int nf, const int* __restrict__ f_node, const int* __restrict__ f_start, const int* __restrict__ f_idx, int f_nodes,
        float th_low, float nnratio, int check_ori, int* __restrict__ match_f, int* __restrict__ nmatches, signed char* __restrict__ bin_of) {
    __shared__ int hist[AFV_HISTO_LENGTH];
    const int lane = threadIdx.x;
    for (int i = lane; i < nf; i += 32) { match_f[i] = -1; bin_of[i] = -1; }
    if (lane < AFV_HISTO_LENGTH) hist[lane] = 0;
    __syncwarp();
    int nm = 0, a = 0, b = 0;
    while (a < kf_nodes && b < f_nodes) {
        const int na = kf_node[a], nb = f_node[b];
        if (na == nb) {
            const int fb0 = f_start[b], fb1 = f_start[b + 1];
            for (int iKF = kf_start[a]; iKF < kf_start[a + 1]; ++iKF) {
                const int realKF = kf_idx[iKF];
                const uint8_t* ref = dkf + (long long)realKF * D;
                Top2 t; t.k1 = t.k2 = KEY_NONE;
                for (int iF = fb0 + lane; iF < fb1; iF += 32) {
                    const int realF = f_idx[iF];
                    if (match_f[realF] >= 0) continue;                       // :232-233
                    top2_push(t, make_key(desc_distance(desc_type, ref, df + (long long)realF * D, D), (uint32_t)(iF - fb0)));
                }
                top2_warp_reduce(t);
                if (lane == 0 && t.k1 != KEY_NONE) {
                    const float bd1 = key_dist(t.k1), bd2 = key_dist(t.k2);
                    if (bd1 <= th_low && bd1 < __fmul_rn(nnratio, bd2)) {
                        const int bestF = f_idx[fb0 + (int)(uint32_t)t.k1];
                        match_f[bestF] = realKF; ++nm;
                        if (check_ori) { const int bin = rot_bin(kkf[realKF].angle, kf_f[bestF].angle); bin_of[bestF] = (signed char)bin; hist[bin]++; }
                    }
                }
                __syncwarp();
                __threadfence_block();
            }
            ++a; ++b;
        } else if (na < nb) ++a; else ++b;
    }
    nm = __shfl_sync(0xffffffffu, nm, 0);
    if (check_ori) {
        __shared__ int keepbin[3];
        if (lane == 0) three_maxima(hist, keepbin[0], keepbin[1], keepbin[2]);
        __syncwarp();
        int removed = 0;
        for (int i = lane; i < nf; i += 32) {
            const int bn = bin_of[i];
            if (bn < 0 || bn == keepbin[0] || bn == keepbin[1] || bn == keepbin[2]) continue;
            match_f[i] = -1; ++removed;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nm -= removed;
    }
    if (lane == 0) *nmatches = nm;
}

extern "C" int afv_search_by_bow(int desc_type, const void* d_dkf, const afv_keypoint* d_kkf,
        const int* d_kf_node, const int* d_kf_start, const int* d_kf_idx, int kf_nodes,
        const void* d_df, const afv_keypoint* d_kf_f, int nf,
        const int* d_f_node, const int* d_f_start, const int* d_f_idx, int f_nodes,
        float th_low, float nnratio, int check_orientation, int* d_match_f, int* d_nmatches, void* cuda_stream) {
    const int D = desc_bytes(desc_type);
    if (D < 0 || !d_dkf || !d_kkf || !d_kf_node || !d_kf_start || !d_kf_idx || !d_df || !d_kf_f || !d_f_node || !d_f_start ||
        !d_f_idx || !d_match_f || !d_nmatches || nf < 0) { afv_set_error("afv_search_by_bow: bad argument"); return AFV_ERR_INVALID; }
    signed char* bin_of = nullptr;
    AFV_CUDA_CHECK(cudaMallocAsync((void**)&bin_of, (size_t)(nf > 0 ? nf : 1), as_stream(cuda_stream)));
    k_search_bow<<<1, 32, 0, as_stream(cuda_stream)>>>(desc_type, D, (const uint8_t*)d_dkf, d_kkf, d_kf_node, d_kf_start, d_kf_idx, kf_nodes,
        (const uint8_t*)d_df, d_kf_f, nf, d_f_node, d_f_start, d_f_idx, f_nodes, th_low, nnratio, check_orientation, d_match_f, d_nmatches, bin_of);
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    AFV_CUDA_CHECK(cudaFreeAsync(bin_of, as_stream(cuda_stream)));
    return AFV_OK;
}

// ---- DescriptorDistance for n pairs ----------------------------------------------------------------------
__global__ void k_desc_distance(int desc_type, int D, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = desc_distance(desc_type, a + (long long)i * D, b + (long long)i * D, D);
}
extern "C" int afv_descriptor_distance(int desc_type, const void* d_a, const void* d_b, int n, float* d_out, void* cuda_stream) {
    const int D = desc_bytes(desc_type);
    if (D < 0 || !d_a || !d_b || !d_out || n < 0) { afv_set_error("afv_descriptor_distance: bad argument"); return AFV_ERR_INVALID; }
    if (n == 0) return AFV_OK;
    k_desc_distance<<<(n + 127) / 128, 128, 0, as_stream(cuda_stream)>>>(desc_type, D, (const uint8_t*)d_a, (const uint8_t*)d_b, n, d_out);
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

// ---- Image::GetGrayImage (K1): OpenCV 8-bit cvtColor, 15-bit fixed point ---------------------------------------
__global__ void k_gray(const uint8_t* __restrict__ src, int ch, int c0, int c1, int c2, int w, int h, int sstride, long long sfs,
                       uint8_t* __restrict__ dst, int dstride, long long dfs) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= w || y >= h) return;
    const uint8_t* p = src + f * sfs + (long long)y * sstride + (long long)x * ch;
    dst[f * dfs + (long long)y * dstride + x] = (uint8_t)((p[0] * c0 + p[1] * c1 + p[2] * c2 + (1 << 14)) >> 15);
}
extern "C" int afv_gray_from_color(const uint8_t* d_src, int channels, int rgb, int B, int w, int h, int src_stride,
                                   long src_frame_stride, uint8_t* d_gray, int gray_stride, long gray_frame_stride, void* cuda_stream) {
    if (!d_src || !d_gray || (channels != 3 && channels != 4) || B < 1 || w < 1 || h < 1 || src_stride < w * channels || gray_stride < w) {
        afv_set_error("afv_gray_from_color: bad argument"); return AFV_ERR_INVALID;
    }
    const int RY = 9798, GY = 19235, BY = 3735;
    k_gray<<<dim3((w + 255) / 256, h, B), 256, 0, as_stream(cuda_stream)>>>(d_src, channels, rgb ? RY : BY, GY, rgb ? BY : RY, w, h, src_stride,
                                                                           src_frame_stride, d_gray, gray_stride, gray_frame_stride);
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

// ---- DBoW2 tree descent (Vocabulary::transform): one warp per feature, lanes over the children of the current node ----
__device__ __forceinline__ double dbow_distance(int desc_type, const uint8_t* a, const uint8_t* b) {
    if (desc_type == AFV_FEAT_SIFT128) {                 // FSift128::distance: float products accumulated in double
        const float* x = (const float*)a; const float* y = (const float*)b;
        double sqd = 0.;
        for (int i = 0; i < 128; ++i) { const float d = x[i] - y[i]; sqd += (double)(d * d); }
        return sqd;
    }
    const int nb = desc_type == AFV_FEAT_ORB32 ? 32 : desc_type == AFV_FEAT_AKAZE61 ? 56 : 48;   // FAkaze61 compares 7 x 8 bytes
    int d = 0;
    for (int i = 0; i < nb; ++i) d += __popc((uint32_t)(a[i] ^ b[i]));
    return (double)d;
}
__global__ void __launch_bounds__(256) k_bow_transform(int desc_type, int D, const uint8_t* __restrict__ desc, int n,
        const int* __restrict__ child_off, const int* __restrict__ child_ids, const uint8_t* __restrict__ node_desc,
        const int* __restrict__ node_word, const double* __restrict__ node_weight, int depth_L, int levelsup,
        int* __restrict__ word_id, double* __restrict__ weight, int* __restrict__ node_id) {
    const int fi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (fi >= n) return;
    const uint8_t* f = desc + (long long)fi * D;
    const int nid_level = depth_L - levelsup;
    int nid = 0, final_id = 0, level = 0;
    for (;;) {
        const int c0 = child_off[final_id], c1 = child_off[final_id + 1];
        if (c1 <= c0) break;                                            // leaf
        ++level;
        // (distance, child position) minimum: first minimum wins
        double bd = 1e300; int bpos = 0x7fffffff;
        for (int c = c0 + lane; c < c1; c += 32) {
            const double d = dbow_distance(desc_type, f, node_desc + (long long)child_ids[c] * D);
            if (d < bd) { bd = d; bpos = c; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o); const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
            if (od < bd || (od == bd && op < bpos)) { bd = od; bpos = op; }
        }
        final_id = child_ids[bpos];
        if (level == nid_level) nid = final_id;
    }
    if (lane == 0) { word_id[fi] = node_word[final_id]; weight[fi] = node_weight[final_id]; node_id[fi] = nid_level <= 0 ? 0 : nid; }
}
extern "C" int afv_bow_transform(int desc_type, const void* d_desc, int n, const int* d_child_off, const int* d_child_ids,
                                 const void* d_node_desc, const int* d_node_word, const double* d_node_weight, int n_nodes, int depth_L,
                                 int levelsup, int* d_word_id, double* d_weight, int* d_node_id, void* cuda_stream) {
    const int D = desc_bytes(desc_type);
    if (D < 0 || n < 0 || n_nodes < 1 || !d_child_off || !d_child_ids || !d_node_desc || !d_node_word || !d_node_weight) {
        afv_set_error("afv_bow_transform: bad argument"); return AFV_ERR_INVALID;
    }
    if (n == 0) return AFV_OK;
    if (!d_desc || !d_word_id || !d_weight || !d_node_id) { afv_set_error("afv_bow_transform: NULL argument"); return AFV_ERR_INVALID; }
    k_bow_transform<<<(n + 7) / 8, 256, 0, as_stream(cuda_stream)>>>(desc_type, D, (const uint8_t*)d_desc, n, d_child_off, d_child_ids,
        (const uint8_t*)d_node_desc, d_node_word, d_node_weight, depth_L, levelsup, d_word_id, d_weight, d_node_id);
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

// ---- MapPoint::ComputeDistinctiveDescriptors: one warp per map point ---------------------------------------------------
#define DD_MAXN 256
__global__ void __launch_bounds__(128) k_distinctive(int desc_type, int D, const uint8_t* __restrict__ desc, const int* __restrict__ obs,
                                                     const int* __restrict__ seg_start, int M, int* __restrict__ best) {
    __shared__ float rowbuf[4][DD_MAXN];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int m = blockIdx.x * 4 + wid;
    if (m >= M) return;
    const int s0 = seg_start[m], N = min(seg_start[m + 1] - s0, DD_MAXN);
    float* row = rowbuf[wid];
    float bestMedian = FLT_MAX; int bestIdx = N > 0 ? 0 : -1;
    const int kth = (int)(0.5 * (double)(N - 1));                       // vDists[0.5*(N-1)]
    for (int i = 0; i < N; ++i) {
        const uint8_t* di = desc + (long long)obs[s0 + i] * D;
        for (int j = lane; j < N; j += 32)
            row[j] = j == i ? 0.0f : desc_distance(desc_type, j > i ? di : desc + (long long)obs[s0 + j] * D,
                                                   j > i ? desc + (long long)obs[s0 + j] * D : di, D);   // Distances[i][j] = dist(min,max)
        __syncwarp();
        float median = FLT_MAX;
        for (int j = lane; j < N; j += 32) {
            const float v = row[j];
            int rank = 0;
            for (int l = 0; l < N; ++l) { const float w = row[l]; rank += (w < v) || (w == v && l < j); }
            if (rank == kth) median = v;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) median = fminf(median, __shfl_xor_sync(0xffffffffu, median, o));
        if (median < bestMedian) { bestMedian = median; bestIdx = i; }
        __syncwarp();
    }
    if (lane == 0) best[m] = bestIdx;
}
extern "C" int afv_distinctive_descriptors(int desc_type, const void* d_desc, const int* d_obs, const int* d_seg_start, int M,
                                           int max_seg, int* d_best, void* cuda_stream) {
    const int D = desc_bytes(desc_type);
    if (D < 0 || M < 0 || (M > 0 && (!d_desc || !d_obs || !d_seg_start || !d_best)) || max_seg > DD_MAXN) {
        afv_set_error("afv_distinctive_descriptors: bad argument (segments are limited to %d observations)", DD_MAXN); return AFV_ERR_INVALID;
    }
    if (M == 0) return AFV_OK;
    k_distinctive<<<(M + 3) / 4, 128, 0, as_stream(cuda_stream)>>>(desc_type, D, (const uint8_t*)d_desc, d_obs, d_seg_start, M, d_best);
    ++g_afv_launches;
    AFV_CUDA_CHECK(cudaGetLastError());
    return AFV_OK;
}

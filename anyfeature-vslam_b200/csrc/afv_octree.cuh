// afv_octree.cuh -- DistributeOctTree (reference src/ORBextractor.cc:181-458) as a CTA-wide device routine shared by the
// orb32 (k_octree) and sift128 (k_sift_octree) extractors.
// The reference grows a std::list of nodes: every pass divides nodes into 4 children that are pushed to
// the FRONT of the list; when the next full pass could overshoot N it divides the biggest nodes first
// (sort by (nKeys, node*)) and stops at N.  Here the list is an array rebuilt per pass: position of the
// t-th child of the r-th divided node = K-1-(P_r+t) (K = children created this pass, P_r = exclusive prefix
// in division order), undivided nodes follow in their old order.  Keys only carry their node's list
// position.  Ties: equal-nKeys nodes are divided later-created-first (= smaller list position first).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct OctNode { short ulx, uly, urx, bry; };

__device__ __forceinline__ int block_scan_excl(int* data, int n, int* warp_tot, int tid) {
    // in-place exclusive scan of data[0..n) by 256 threads; returns the total. n <= 256*items.
    const int items = (n + 255) / 256;
    const int beg = tid * items, end = min(beg + items, n);
    int sum = 0;
    for (int i = beg; i < end; ++i) sum += data[i];
    const int lane = tid & 31, wid = tid >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const int v = warp_tot[w]; if (w < wid) wbase += v; total += v; }
    int run = wbase + incl - sum;
    for (int i = beg; i < end; ++i) { const int v = data[i]; data[i] = run; run += v; }
    __syncthreads();
    return total;
}

// node-side workspace (shared memory), ncap = node capacity
struct OctWork {
    OctNode* nd[2]; int* ncnt[2];
    int *qcnt, *scanA, *scanB, *drank, *order, *newpos;
    unsigned long long* best;
};
__host__ __device__ inline size_t oct_work_bytes(int ncap) {
    size_t off = sizeof(OctNode) * (size_t)ncap * 2 + 4 * (size_t)ncap * 2 + 16 * (size_t)ncap + 4 * (size_t)ncap * 5;
    off = (off + 7) & ~(size_t)7;
    return off + 8 * (size_t)ncap;
}
// base must be 16-byte aligned
__device__ __forceinline__ void oct_carve(unsigned char* base, int ncap, OctWork& W) {
    size_t off = 0;
    W.nd[0] = reinterpret_cast<OctNode*>(base + off); off += sizeof(OctNode) * ncap;
    W.nd[1] = reinterpret_cast<OctNode*>(base + off); off += sizeof(OctNode) * ncap;
    W.ncnt[0] = reinterpret_cast<int*>(base + off); off += 4 * ncap;
    W.ncnt[1] = reinterpret_cast<int*>(base + off); off += 4 * ncap;
    W.qcnt = reinterpret_cast<int*>(base + off); off += 16 * ncap;         // [ncap][4]
    W.scanA = reinterpret_cast<int*>(base + off); off += 4 * ncap;         // children prefix (division order)
    W.scanB = reinterpret_cast<int*>(base + off); off += 4 * ncap;         // survivor prefix (list order)
    W.drank = reinterpret_cast<int*>(base + off); off += 4 * ncap;         // division rank of node (or -1)
    W.order = reinterpret_cast<int*>(base + off); off += 4 * ncap;         // node at division rank r
    W.newpos = reinterpret_cast<int*>(base + off); off += 4 * ncap;        // survivor's new position / child base
    W.best = reinterpret_cast<unsigned long long*>(base + ((off + 7) & ~(size_t)7));
}

// Distributes M keys (kx, ky in full-image coordinates; any address space) over the quadtree until >= N nodes.
// On return knode[k] = final list position of key k's node; returns the node count (list order = reference list
// order).  Called by all 256 threads of the CTA.  overflow is set when the node list would exceed ncap.
// Key positions come through an accessor (x(k), y(k)): plain float arrays (OctXYArrays, any address space) or, for orb32, level
// coordinates packed in one word and scaled on the fly (k_octree: 11 instead of 19 bytes of shared memory per key).
struct OctXYArrays {
    const float* kx; const float* ky;
    __device__ __forceinline__ float x(int k) const { return kx[k]; }
    __device__ __forceinline__ float y(int k) const { return ky[k]; }
};
template <class XY>
__device__ __forceinline__ int oct_distribute_xy(const OctWork& W, const XY xy, unsigned short* knode,
                                                 unsigned char* kquad, int M, int N, int nIni, float hX, int H, int ncap,
                                                 int tid, bool& overflow);
__device__ __forceinline__ int oct_distribute(const OctWork& W, const float* kx, const float* ky, unsigned short* knode,
                                              unsigned char* kquad, int M, int N, int nIni, float hX, int H, int ncap,
                                              int tid, bool& overflow) {
    OctXYArrays a; a.kx = kx; a.ky = ky;
    return oct_distribute_xy(W, a, knode, kquad, M, N, nIni, hX, H, ncap, tid, overflow);
}
template <class XY>
__device__ __forceinline__ int oct_distribute_xy(const OctWork& W, const XY xy, unsigned short* knode,
                                                 unsigned char* kquad, int M, int N, int nIni, float hX, int H, int ncap,
                                                 int tid, bool& overflow) {
    __shared__ int warp_tot[8];
    __shared__ int s_J, s_nexp;
    OctNode* const* nd = W.nd; int* const* ncnt = W.ncnt;
    int* qcnt = W.qcnt; int* scanA = W.scanA; int* scanB = W.scanB; int* drank = W.drank; int* order = W.order; int* newpos = W.newpos;
    // ---- roots (reference :243-283): nIni = round(w/h) boxes of width hX; empty roots erased
    for (int i = tid; i < nIni; i += 256) {
        OctNode n; n.ulx = (short)(int)__fmul_rn(hX, (float)i); n.uly = 0;
        n.urx = (short)(int)__fmul_rn(hX, (float)(i + 1)); n.bry = (short)H;
        nd[0][i] = n; ncnt[0][i] = 0;
    }
    __syncthreads();
    for (int k = tid; k < M; k += 256) {
        const int b = min((int)__fdiv_rn(xy.x(k), hX), nIni - 1);
        knode[k] = (unsigned short)b;
        atomicAdd(&ncnt[0][b], 1);
    }
    __syncthreads();
    // compact away empty roots (order preserved)
    for (int i = tid; i < nIni; i += 256) scanB[i] = ncnt[0][i] > 0 ? 1 : 0;
    __syncthreads();
    int size = block_scan_excl(scanB, nIni, warp_tot, tid);
    for (int i = tid; i < nIni; i += 256)
        if (ncnt[0][i] > 0) { nd[1][scanB[i]] = nd[0][i]; ncnt[1][scanB[i]] = ncnt[0][i]; }
    for (int k = tid; k < M; k += 256) knode[k] = (unsigned short)scanB[knode[k]];
    __syncthreads();
    int cur = 1;

    bool finish = false, careful = false;
    while (!finish) {
        const int prevSize = size;
        OctNode* A = nd[cur]; int* Ac = ncnt[cur];
        OctNode* Bn = nd[cur ^ 1]; int* Bc = ncnt[cur ^ 1];
        // ---- tentative division of every expandable node: quadrant populations
        for (int i = tid; i < size * 4; i += 256) qcnt[i] = 0;
        __syncthreads();
        for (int k = tid; k < M; k += 256) {
            const int p = knode[k];
            if (Ac[p] > 1) {
                const OctNode n = A[p];
                const int halfX = (int)ceilf(__fdiv_rn((float)(n.urx - n.ulx), 2.f));
                const int halfY = (int)ceilf(__fdiv_rn((float)(n.bry - n.uly), 2.f));
                const float sx = (float)(n.ulx + halfX), sy = (float)(n.uly + halfY);
                const float px = xy.x(k), py = xy.y(k);
                const int q = (px < sx) ? ((py < sy) ? 0 : 2) : ((py < sy) ? 1 : 3);
                kquad[k] = (unsigned char)q;
                atomicAdd(&qcnt[p * 4 + q], 1);
            }
        }
        __syncthreads();
        // ---- division order: list order (full pass) or (nKeys desc, position asc) (careful pass)
        if (!careful) {
            for (int p = tid; p < size; p += 256) scanA[p] = Ac[p] > 1 ? 1 : 0;
            __syncthreads();
            const int ne = block_scan_excl(scanA, size, warp_tot, tid);
            for (int p = tid; p < size; p += 256) {
                if (Ac[p] > 1) { drank[p] = scanA[p]; order[scanA[p]] = p; } else drank[p] = -1;
            }
            if (tid == 0) s_nexp = ne;
            __syncthreads();
        } else {
            int ne_local = 0;
            for (int p = tid; p < size; p += 256) {
                int r = -1;
                if (Ac[p] > 1) {
                    r = 0;
                    const int c = Ac[p];
                    for (int o = 0; o < size; ++o) {
                        const int co = Ac[o];
                        r += (co > 1) && (co > c || (co == c && o < p));
                    }
                    order[r] = p;
                    ++ne_local;
                }
                drank[p] = r;
            }
            if (tid == 0) s_nexp = 0;
            __syncthreads();
            if (ne_local) atomicAdd(&s_nexp, ne_local);
            __syncthreads();
        }
        const int nexp = s_nexp;
        // children per division rank, prefix in division order
        for (int r = tid; r < nexp; r += 256) {
            const int p = order[r];
            scanA[r] = (qcnt[p * 4] > 0) + (qcnt[p * 4 + 1] > 0) + (qcnt[p * 4 + 2] > 0) + (qcnt[p * 4 + 3] > 0);
        }
        __syncthreads();
        // how many divisions happen (J): all in a full pass; in a careful pass stop once size >= N
        if (tid == 0) s_J = nexp;
        __syncthreads();
        if (careful) {
            // running size after r+1 divisions = size + sum_{i<=r}(c_i - 1); find the first r reaching N
            for (int r = tid; r < nexp; r += 256) newpos[r] = scanA[r] - 1;
            __syncthreads();
            block_scan_excl(newpos, nexp, warp_tot, tid);
            for (int r = tid; r < nexp; r += 256)
                if (size + newpos[r] + scanA[r] - 1 >= N) atomicMin(&s_J, r + 1);
            __syncthreads();
        }
        const int J = s_J;
        const int K = block_scan_excl(scanA, J, warp_tot, tid);            // scanA[r] = P_r for r < J
        // survivors (every node that is not divided this pass), prefix in list order
        for (int p = tid; p < size; p += 256) scanB[p] = (drank[p] >= 0 && drank[p] < J) ? 0 : 1;
        __syncthreads();
        const int nsurv = block_scan_excl(scanB, size, warp_tot, tid);
        const int newSize = K + nsurv;
        if (newSize > ncap) {                 // cannot happen with ncap >= N + 4*... ; flag and stop
            overflow = true;
            break;
        }
        if (tid == 0) s_nexp = 0;
        __syncthreads();
        // ---- build the new list
        int nexp_local = 0;
        for (int p = tid; p < size; p += 256) {
            const int r = drank[p];
            if (r >= 0 && r < J) {
                const OctNode n = A[p];
                const int halfX = (int)ceilf(__fdiv_rn((float)(n.urx - n.ulx), 2.f));
                const int halfY = (int)ceilf(__fdiv_rn((float)(n.bry - n.uly), 2.f));
                int tnum = 0;
                const int base = K - 1 - scanA[r];
                newpos[p] = base;                                  // children: base - t
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = qcnt[p * 4 + q];
                    if (c > 0) {
                        OctNode ch;
                        ch.ulx = (q & 1) ? (short)(n.ulx + halfX) : n.ulx;
                        ch.urx = (q & 1) ? n.urx : (short)(n.ulx + halfX);
                        ch.uly = (q & 2) ? (short)(n.uly + halfY) : n.uly;
                        ch.bry = (q & 2) ? n.bry : (short)(n.uly + halfY);
                        Bn[base - tnum] = ch; Bc[base - tnum] = c;
                        qcnt[p * 4 + q] = -(base - tnum) - 1;       // remember the child's position (negative-coded)
                        nexp_local += (c > 1);
                        ++tnum;
                    }
                }
            } else {
                const int np = K + scanB[p];
                newpos[p] = np;
                Bn[np] = A[p]; Bc[np] = Ac[p];
            }
        }
        if (nexp_local) atomicAdd(&s_nexp, nexp_local);
        __syncthreads();
        for (int k = tid; k < M; k += 256) {
            const int p = knode[k];
            const int r = drank[p];
            knode[k] = (unsigned short)((r >= 0 && r < J) ? (-qcnt[p * 4 + kquad[k]] - 1) : newpos[p]);
        }
        __syncthreads();
        const int nToExpand = s_nexp;
        size = newSize;
        cur ^= 1;
        // ---- termination (reference :357-362 / :437-438)
        if (size >= N || size == prevSize) finish = true;
        else if (!careful && size + nToExpand * 3 > N) careful = true;
        __syncthreads();
    }
    return size;
}

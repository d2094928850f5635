// afv_orb.cu -- hand-written sm_100a kernels of the orb32 extraction path.
//
// Replaces, behind the C ABI of include/afv.h, what the reference's FeatureExtractor_orb32 runs on the CPU
// per frame (reference src/Feature_orb32.cpp:11-65 -> cv::ORB::detect / DistributeOctTree / cv::ORB::compute):
//   K2 k_resize          cv::resize INTER_LINEAR_EXACT pyramid (level l from l-1)
//   K3 k_fast            FAST-9/16 score + 3x3 NMS, unordered candidate append
//   K4/K5 k_harris_select retainBest(2q, FAST score) -> Harris 7x7 -> retainBest(q, Harris) (radix select)
//   K9 k_octree          DistributeOctTree (src/ORBextractor.cc:181-458), one CTA per (frame, level)
//   K7 k_blur            7x7 sigma-2 float separable blur of every level
//   K6/K8/K10 k_describe IC angle + 256-bit steered BRIEF + merge into the caller's cv::KeyPoint layout
// Every arithmetic step is written to round exactly like the pinned CPU path (see oracle/afv_oracle.c):
// integer fixed point, explicit __f*_rn intrinsics (no contraction) and explicit __fmaf_rn where the pinned
// binary fuses.  Results are bit-exact with the oracle; tests/test_extract_gpu.py checks that.
#include "afv_common.cuh"
#include "afv_octree.cuh"
#include <cuda.h>
#include <stdio.h>
#include <string.h>

// ---- TMA (cp.async.bulk.tensor) + mbarrier primitives, inline PTX (sm_90+/sm_100a) -----------------------------
struct AfvTmaps { CUtensorMap m[AFV_MAX_LEVELS]; };        // one 3-D map (x bytes, y rows, frame) per pyramid level

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok = 0;
    for (int spin = 0; !ok; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
        if (spin > (1 << 24)) __trap();                  // a lost TMA completion must not hang the box
    }
}

// Global, not __constant__: the 256 threads of a CTA read 256 different words of it, which the constant cache serialises
// (7 % of k_describe's stall samples sat on that conversion); as a coalesced LDG.32 per thread it is one L2-resident KB.
static __device__ __align__(16) int8_t g_pattern[1024] = {
#include "orb_pattern.inc"
};

__device__ __forceinline__ int refl101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

// ------------------------------------------------------------------------------------------------------
// K2: pyramid level from the previous one.  thread = 4 consecutive output pixels (one 32-bit store).
// ------------------------------------------------------------------------------------------------------
// One CTA = 128x32 output pixels.  The source footprint (<= 160 x 41 bytes at scale 1.2) is staged in shared memory
// with aligned 32-bit loads; taps are byte LDS with 32-bit addressing (the global-pointer version spent most of its
// instructions on 64-bit address arithmetic).  Thread = 4 columns x 2 rows; coefficients come as one 128-bit load of
// 4 packed table entries (offset << 16 | c1).  (128 x 16 tiles spent a quarter of their instructions on the per-CTA prologue.)
#define RS_ROWS 44                    // source rows of 32 output rows at scale 1.2 (38.4) + the second tap + slack
#define RS_OUT_H 32
#define RS_PITCH 176
__global__ void __launch_bounds__(256) k_resize(const __grid_constant__ AfvParams P, const __grid_constant__ CUtensorMap tm_src, int l) {
    __shared__ __align__(128) uint8_t sp[RS_ROWS][RS_PITCH];
    __shared__ __align__(8) uint64_t tma_bar;
    const AfvLevel& D = P.lv[l];
    const AfvLevel& S = P.lv[l - 1];
    const int f = blockIdx.z, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int tx0 = blockIdx.x * 128, ty0 = blockIdx.y * RS_OUT_H;
    const int sxb = (int)(D.xtab[tx0] >> 16) & ~15;                // first staged source column (16-B aligned for TMA)
    const int syb = (int)(D.ytab[ty0] >> 16);                      // first staged source row
    // source footprint (<= 172 x 22 bytes incl. alignment slack) by one TMA box; rows / columns past the source are
    // zero-filled and only ever multiplied by a zero coefficient (the tables clamp the last tap)
    if (tid == 0) mbar_init(&tma_bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&tma_bar, RS_ROWS * RS_PITCH);
        tma_load_3d(&sp[0][0], &tm_src, &tma_bar, sxb, syb, f);
    }
    mbar_wait(&tma_bar, 0);
    __syncthreads();
    const int x0 = tx0 + 4 * lane;
    if (x0 >= D.w) return;
    const uint4 xt = *reinterpret_cast<const uint4*>(D.xtab + x0);       // table padded to a multiple of 4 entries
    const uint32_t xe[4] = {xt.x, xt.y, xt.z, xt.w};
    uint8_t* dstf = const_cast<uint8_t*>(D.img) + (long long)f * D.img_fstride;
#pragma unroll
    for (int rr = 0; rr < RS_OUT_H / 8; ++rr) {
        const int y = ty0 + (RS_OUT_H / 8) * wrp + rr;
        if (y >= D.h) break;
        const uint32_t yt = D.ytab[y];
        const int yo = yt >> 16, c1y = yt & 0xffff, c0y = 256 - c1y;
        const uint8_t* r0 = &sp[yo - syb][0];
        const uint8_t* r1 = &sp[min(yo + 1, S.h - 1) - syb][0];
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xr = (int)(xe[k] >> 16) - sxb, c1 = xe[k] & 0xffff, c0 = 256 - c1;
            const uint32_t h0 = c0 * r0[xr] + c1 * r0[xr + 1];
            const uint32_t h1 = c0 * r1[xr] + c1 * r1[xr + 1];
            const uint32_t v = (uint32_t)c0y * h0 + (uint32_t)c1y * h1;
            out |= ((v + (1u << 15)) >> 16) << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(dstf + (long long)y * D.img_stride + x0) = out;   // rows padded to 128 B
    }
}

// ------------------------------------------------------------------------------------------------------
// K3: FAST-9/16 + NMS.  Tile 128x16 (+4 halo), 256 threads.
// ------------------------------------------------------------------------------------------------------
#define FT_W AFV_TILE_W
#define FT_H AFV_TILE_H
#define FT_PW (FT_W + 8)
#define FT_PH (FT_H + 8)
#define FT_RW (FT_W + 2)
#define FT_RH (FT_H + 2)
#define FT_SW 132

__device__ __forceinline__ bool has_arc9(uint32_t m) {
    m |= m << 16;
    uint32_t a = m & (m >> 1);
    a &= a >> 2;
    a &= a >> 4;
    a &= m >> 8;
    return (a & 0xffffu) != 0;
}

#define CIRC16(F) F(0, 0, 3) F(1, 1, 3) F(2, 2, 2) F(3, 3, 1) F(4, 3, 0) F(5, 3, -1) F(6, 2, -2) F(7, 1, -3) \
                  F(8, 0, -3) F(9, -1, -3) F(10, -2, -2) F(11, -3, -1) F(12, -3, 0) F(13, -3, 1) F(14, -2, 2) F(15, -1, 3)

// Two-stage segment test with compaction (keeps warps converged): stage 1 runs the two cheap antipodal-pair
// rejections on every pixel of the score region and compacts the survivors (warp ballot + one shared atomic per
// warp); stage 2 builds the two 16-bit brighter/darker masks only for survivors and compacts the corners.
#define FT_WP 40                      // words per staged row: 160 B TMA box starting at x0-16 (16-B aligned)
#define FT_XOFF 12                    // byte offset of pixel x0-4 inside a staged row
__device__ __forceinline__ void warp_push(bool pass, uint16_t val, uint16_t* list, int* counter, int lane) {
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m == 0) return;
    int base = 0;
    const int leader = __ffs(m) - 1;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pass) list[base + __popc(m & ((1u << lane) - 1))] = val;
}

__global__ void __launch_bounds__(256) k_fast(const __grid_constant__ AfvParams P, const __grid_constant__ AfvTmaps TM) {
    __shared__ __align__(128) uint32_t pixw[FT_PH][FT_WP];
    __shared__ __align__(8) uint64_t tma_bar;
    __shared__ __align__(16) uint8_t score[FT_RH][FT_SW + 4];          // 34 x 136 = 289 x 16 bytes: zeroed with 128-bit stores
    __shared__ uint16_t slist[FT_RW * FT_RH];        // stage-1 survivors
    __shared__ uint16_t clist[FT_RW * FT_RH];        // corners
    __shared__ uint32_t surv[(FT_W / 2) * (FT_H / 2) + 64];
    __shared__ int nstage1, ncorner, nsurv, gbase;

    const AfvTile ti = P.tiles[blockIdx.x];
    const int l = ti.level;
    const AfvLevel& L = P.lv[l];
    const int f = blockIdx.y;
    const int x0 = ti.x0, y0 = ti.y0;
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int t = P.fast_th;

    // stage the pixel tile with ONE TMA box load: FT_WP*4 = 160 bytes x FT_PH rows of frame f, starting at
    // (x0 - 16, y0 - 4): the inner coordinate of a u8 box must be 16-byte aligned (an unaligned start raises an
    // illegal-instruction fault, tools/dbg/tma_dbg.cu), so pixel x0-4 sits at byte FT_XOFF of every staged row;
    // everything outside the level (halo of border tiles) is zero-filled by the TMA unit.
    if (tid == 0) {
        nstage1 = 0; ncorner = 0; nsurv = 0;
        mbar_init(&tma_bar, 1);
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&tma_bar, FT_PH * FT_WP * 4);
        tma_load_3d(&pixw[0][0], &TM.m[l], &tma_bar, x0 - 16, y0 - 4, f);
    }
    static_assert((FT_RH * (FT_SW + 4)) % 16 == 0, "score tile is a whole number of uint4");
    for (int i = tid; i < FT_RH * (FT_SW + 4) / 16; i += 256) reinterpret_cast<uint4*>(&score[0][0])[i] = make_uint4(0, 0, 0, 0);
    mbar_wait(&tma_bar, 0);
    __syncthreads();
    (void)wrp;

    const uint8_t* pb = reinterpret_cast<const uint8_t*>(&pixw[0][0]);
    constexpr int PITCH = FT_WP * 4;
    // stage 1: sign-consistent compass test on FOUR pixels per thread and row, two pixels per 32-bit register.
    // An arc of 9 of the 16 circle pixels always contains two ADJACENT compass points (0, 4, 8, 12), so a corner needs two
    // adjacent compass points that are both brighter than v+t or both darker than v-t: (dn|up) & (rt|lf) per sign.
    // A staged word w = (p0,p1,p2,p3) is split into E = (p0,p2) and O = (p1,p3), one pixel per 16-bit half.  With
    // C = (0x8000 + t) in both halves, bit 15 of a half of (E + C) - x is CLEAR iff x > v + t and bit 15 of x + (C - E) is
    // CLEAR iff x < v - t (all halves stay in [0x7f01, 0x81fe]: no carry between halves), so one IADD does a compare for two
    // pixels and the pass condition is five LOP3s.  A thread owns one word column and FS1_ROWS rows and keeps the E / O words of
    // its column in a rolling register window (up / down neighbours); left / right neighbours come from the adjacent words.
    constexpr int FS1_ROWS = 5, FS1_WC0 = 3, FS1_NWC = 34, FS1_CH = 7;       // words 3..36 = staged bytes 12..147 cover the region
    static_assert(FS1_ROWS * FS1_CH >= FT_RH && FS1_NWC * FS1_CH <= 256, "stage-1 work split");
    static_assert(FT_XOFF + 3 == 15 && FS1_WC0 * 4 <= 15 && (FS1_WC0 + FS1_NWC) * 4 > 15 + FT_RW - 1, "word columns cover the region");
    {
        uint32_t smask = 0;                                        // bit 4*j + k: row j of the chunk, nibble bit k (see below)
        int wc = 0, ch = 0;
        if (tid < FS1_NWC * FS1_CH) {
            ch = tid / FS1_NWC; wc = FS1_WC0 + tid - ch * FS1_NWC;
            const uint32_t C2 = (0x8000u + (uint32_t)t) * 0x10001u, G = 0x80008000u;
            uint32_t E[FS1_ROWS + 6], O[FS1_ROWS + 6];
#pragma unroll
            for (int k = 0; k < FS1_ROWS + 6; ++k) {               // staged rows ch*FS1_ROWS + k (region row + 3 = staged row)
                const uint32_t w = pixw[min(ch * FS1_ROWS + k, FT_PH - 1)][wc];
                E[k] = w & 0x00ff00ffu; O[k] = __byte_perm(w, 0, 0x4341);
            }
#pragma unroll
            for (int j = 0; j < FS1_ROWS; ++j) {
                const int rs = min(ch * FS1_ROWS + j + 3, FT_PH - 1);
                const uint32_t wm = pixw[rs][wc - 1], wp = pixw[rs][wc + 1];
                const uint32_t Em = wm & 0x00ff00ffu, Om = __byte_perm(wm, 0, 0x4341);
                const uint32_t Ep = wp & 0x00ff00ffu, Op = __byte_perm(wp, 0, 0x4341);
                // pixels (p0,p2): left = (p-3,p-1) = Om, right = (p3,p5) = (O.hi, Op.lo); pixels (p1,p3): left = (p-2,p0) = (Em.hi, E.lo), right = (p4,p6) = Ep
                const uint32_t lfA = Om, rtA = __byte_perm(O[j + 3], Op, 0x5432);
                const uint32_t lfB = __byte_perm(Em, E[j + 3], 0x5432), rtB = Ep;
                // bright pair exists iff min(max(dn,up), max(rt,lf)) > v + t; dark pair iff max(min(dn,up), min(rt,lf)) < v - t
                const uint32_t K1a = E[j + 3] + C2, K2a = C2 - E[j + 3];
                const uint32_t bA = __vmins2(__vmaxs2(E[j + 6], E[j]), __vmaxs2(rtA, lfA));
                const uint32_t dA = __vmaxs2(__vmins2(E[j + 6], E[j]), __vmins2(rtA, lfA));
                const uint32_t K1b = O[j + 3] + C2, K2b = C2 - O[j + 3];
                const uint32_t bB = __vmins2(__vmaxs2(O[j + 6], O[j]), __vmaxs2(rtB, lfB));
                const uint32_t dB = __vmaxs2(__vmins2(O[j + 6], O[j]), __vmins2(rtB, lfB));
                const uint32_t pa = ~((K1a - bA) & (dA + K2a)) & G, pbm = ~((K1b - bB) & (dB + K2b)) & G;   // pass flags at bits 15 / 31
                // nibble: bit 0 = p0, bit 1 = p1, bit 2 = p2, bit 3 = p3
                const uint32_t nib = ((pa >> 15) & 1u) | ((pbm >> 14) & 2u) | ((pa >> 29) & 4u) | ((pbm >> 28) & 8u);
                smask |= nib << (4 * j);
            }
            // validity: region columns c = 4*wc + k - 15 in [0, FT_RW) with 3 <= gx < w-3; region rows r = ch*FS1_ROWS + j < FT_RH with 3 <= gy < h-3
            // region columns 4*wc - 15 + k: only the first and the last word column hold columns outside [0, FT_RW)
            uint32_t cm = wc == FS1_WC0 ? 0x8u : (wc == FS1_WC0 + FS1_NWC - 1 ? 0x1u : 0xfu);
            static_assert(4 * FS1_WC0 - 15 + 3 == 0 && 4 * (FS1_WC0 + FS1_NWC - 1) - 15 == FT_RW - 1, "edge word columns");
            uint32_t vm;
            if (x0 >= 4 && y0 >= 4 && x0 + FT_W + 4 <= L.w && y0 + FT_H + 4 <= L.h) {      // interior tile (uniform): no image-bound tests
                vm = cm * 0x11111u;
                if (ch == FS1_CH - 1) vm &= (1u << (4 * (FT_RH - (FS1_CH - 1) * FS1_ROWS))) - 1u;      // rows >= FT_RH of the last chunk
            } else {
                cm = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = 4 * wc + k - 15, gx = x0 - 1 + c;
                    if (c >= 0 && c < FT_RW && gx >= 3 && gx < L.w - 3) cm |= 1u << k;
                }
                vm = 0;
#pragma unroll
                for (int j = 0; j < FS1_ROWS; ++j) {
                    const int r = ch * FS1_ROWS + j, gy = y0 - 1 + r;
                    if (r < FT_RH && gy >= 3 && gy < L.h - 3) vm |= cm << (4 * j);
                }
            }
            smask &= vm;
        }
        const int cnt = __popc(smask);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t_ = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t_; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int base = 0;
        if (total) {
            if (lane == 31) base = atomicAdd(&nstage1, total);
            base = __shfl_sync(0xffffffffu, base, 31);
        }
        int o = base + incl - cnt;
        const int i0 = ((ch * FS1_ROWS) << 8) + 4 * wc - 15;            // list entries are (region row << 8 | region column): no div / mod later
        // predicated, fully unrolled writer: 16 % of the kernel's warp instructions but mostly predicated off and off the ALU pipe
        // that bounds this kernel; a per-lane `while (smask)` loop issues fewer instructions (12 %) and is SLOWER (1.17 vs 1.13 ms per
        // 512 frames: ~5 of 32 lanes active, ffs / clear / address arithmetic on the ALU pipe).  entry = row << 8 | column
#pragma unroll
        for (int b = 0; b < 4 * FS1_ROWS; ++b)
            if (smask & (1u << b)) slist[o++] = (uint16_t)(i0 + ((b >> 2) << 8) + (b & 3));
    }
    __syncthreads();
    // stage 2: corner score of every stage-1 survivor = max over the 16 arcs of 9 of min|v - p| (same sign), minus 1 (OpenCV
    // cornerScore<16>); the pixel is a FAST-9 corner at threshold t iff that maximum exceeds t, so no separate segment test is
    // needed.  d[k] = v - p[k] in [-255,255] packed as (d, -d) in the two signed 16-bit halves: one __vmins2 chain gives both
    // min(d) (darker arcs) and min(-d) (brighter arcs); sliding minimum over 9 by doubling.
    // NB: the scalar form max(min9(d), -max9(d)) is MISCOMPILED by nvcc 12.9 / ptxas for sm_100a (3-input VIMNMX3 with a
    // negated operand; repro in tools/dbg/minmax_dbg.cu) - keep this formulation.
    const int n1 = nstage1;
    for (int j0 = 0; j0 < n1; j0 += 256) {
        const int j = j0 + tid;
        bool corner = false;
        int i = 0;
        if (j < n1) {
            i = slist[j];
            const int r = i >> 8, c = i & 255;
            const uint8_t* p = pb + (r + 3) * PITCH + FT_XOFF + (c + 3);
            const int v = p[0];
            uint32_t e[16];
            // e[k] = (256 + v - p_k) | (256 - v + p_k) << 16 as ONE multiply-add per circle pixel: p * 0xffff = (p << 16) - p, added to
            // A = (256 + v) | (256 - v) << 16; both halves stay in [1, 511], so there is no borrow between them.  The bias of 256 is common
            // to every half and drops out of the min / max chain (removed from the result below).  The kernel is ALU-pipe bound
            // (ncu: ALU 76 %): IMAD issues on the FMA pipe, the sub / and / neg / shift-or form cost 4 ALU instructions per pixel.
            const uint32_t A = (uint32_t)(256 + v) | ((uint32_t)(256 - v) << 16);
#define FDIFF(k, dx, dy) e[k] = (uint32_t)p[(dy) * PITCH + (dx)] * 0xffffu + A;
            CIRC16(FDIFF)
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
            corner = best > t;
            if (corner) score[r][c] = (uint8_t)(best - 1);
        }
        warp_push(corner, (uint16_t)i, clist, &ncorner, lane);
    }
    __syncthreads();
    const int nc = ncorner;

    // 3x3 non-max suppression (strictly greater than all 8 neighbours), interior of the tile only
    for (int j0 = 0; j0 < nc; j0 += 256) {
        const int j = j0 + tid;
        bool keep = false;
        uint32_t val = 0;
        if (j < nc) {
            const int i = clist[j];
            const int r = i >> 8, c = i & 255;
            if (r >= 1 && r <= FT_H && c >= 1 && c <= FT_W) {
                const int s = score[r][c];
                keep = s > score[r][c - 1] && s > score[r][c + 1] && s > score[r - 1][c - 1] && s > score[r - 1][c] &&
                       s > score[r - 1][c + 1] && s > score[r + 1][c - 1] && s > score[r + 1][c] && s > score[r + 1][c + 1];
                val = (uint32_t)(x0 - 1 + c) | ((uint32_t)(y0 - 1 + r) << 12) | ((uint32_t)s << 24);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);               // one shared atomic per warp
        if (m) {
            int base = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) base = atomicAdd(&nsurv, __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (keep) surv[base + __popc(m & ((1u << lane) - 1))] = val;
        }
    }
    __syncthreads();
    const int ns = nsurv;
    if (ns == 0) return;
    if (tid == 0) {
        gbase = atomicAdd(&P.counts[afv_cnt_idx(f, AFV_CNT_CAND, l)], ns);
        if (gbase + ns > L.cand_cap) atomicOr(&P.status[f], AFV_ST_CAND_OVERFLOW);
    }
    __syncthreads();
    uint32_t* out = L.cand + (long long)f * L.cand_cap;
    for (int j = tid; j < ns; j += 256)
        if (gbase + j < L.cand_cap) out[gbase + j] = surv[j];
}

// ------------------------------------------------------------------------------------------------------
// K4/K5: per (frame, level): retainBest(2*q_orb) on the FAST score, Harris for the survivors,
// retainBest(q_orb) on the Harris response.  "retainBest(n)" keeps the n best plus everything tying with
// the n-th (OpenCV KeyPointsFilter::retainBest) -> exact 256-bin / radix selection, no sort needed.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2ord(float f) {       // order-preserving float -> uint
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__device__ float harris7(const uint8_t* img, int w, int h, int stride, int x0, int y0) {
    // 9x9 support walked with three rolling rows (27 live pixels instead of 81: the full patch cost 128 registers)
    const bool inner = x0 >= 4 && y0 >= 4 && x0 + 4 < w && y0 + 4 < h;
    // rows are 4-byte aligned (arena strides are multiples of 128, an aliased input has stride % 16 == 0): the 9 bytes of a row come
    // from three aligned 32-bit loads + funnel shifts instead of nine byte loads (the kernel waits on these loads: 44 % of its stall
    // samples); the 12-byte window must end inside the row
    const bool wide = inner && x0 + 8 <= w;
    const int xa = (x0 - 4) & ~3, sh = ((x0 - 4) & 3) * 8;
    int ra[9], rb[9], rc[9];
    auto load_row = [&](int rr, int* dst) {
        if (wide) {
            const uint32_t* q = reinterpret_cast<const uint32_t*>(img + (long long)(y0 - 4 + rr) * stride + xa);
            const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
            const uint32_t u0 = __funnelshift_r(w0, w1, sh), u1 = __funnelshift_r(w1, w2, sh), u2 = w2 >> sh;
#pragma unroll
            for (int c = 0; c < 4; ++c) { dst[c] = (u0 >> (8 * c)) & 0xff; dst[4 + c] = (u1 >> (8 * c)) & 0xff; }
            dst[8] = u2 & 0xff;
        } else if (inner) {
            const uint8_t* p = img + (long long)(y0 - 4 + rr) * stride + x0 - 4;
#pragma unroll
            for (int c = 0; c < 9; ++c) dst[c] = p[c];
        } else {
            const uint8_t* p = img + (long long)refl101(y0 - 4 + rr, h) * stride;
#pragma unroll
            for (int c = 0; c < 9; ++c) dst[c] = p[refl101(x0 - 4 + c, w)];
        }
    };
    load_row(0, ra); load_row(1, rb);
    int a = 0, b = 0, c2 = 0;
#pragma unroll
    for (int r = 1; r < 8; ++r) {
        load_row(r + 1, rc);
#pragma unroll
        for (int c = 1; c < 8; ++c) {
            const int Ix = (rb[c + 1] - rb[c - 1]) * 2 + (ra[c + 1] - ra[c - 1]) + (rc[c + 1] - rc[c - 1]);
            const int Iy = (rc[c] - ra[c]) * 2 + (rc[c - 1] - ra[c - 1]) + (rc[c + 1] - ra[c + 1]);
            a += Ix * Ix; b += Iy * Iy; c2 += Ix * Iy;
        }
#pragma unroll
        for (int c = 0; c < 9; ++c) { ra[c] = rb[c]; rb[c] = rc[c]; }
    }
    // ((float)a*b - (float)c*c - 0.04f*((float)a+b)*((float)a+b)) * scale^4, float32, one rounding per op
    const float scale = __fdiv_rn(1.f, 7140.f);            // 1/((1<<2)*7*255.f)
    const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    const float fa = (float)a, fb = (float)b, fc = (float)c2;
    const float s = __fadd_rn(fa, fb);
    const float t2 = __fmul_rn(__fmul_rn(0.04f, s), s);
    const float r = __fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), t2);
    return __fmul_rn(r, s4);
}

// Suffix sums of a 256-bin histogram by the 256 threads of the CTA: S[b] = sum of hist[b..255] (the serial walk from bin 255 down by
// one thread cost ~10 us per pass: the kernel is latency bound and runs 5 such selections per (frame, level)).
__device__ __forceinline__ int suffix_sum_256(const int* hist, int* wsum, int tid) {
    const int lane = tid & 31, wid = tid >> 5;
    int v = hist[tid];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_down_sync(0xffffffffu, v, o); if (lane + o < 32) v += t; }
    if (lane == 0) wsum[wid] = v;
    __syncthreads();
#pragma unroll
    for (int w = 1; w < 8; ++w) if (wid + w < 8) v += wsum[wid + w];
    __syncthreads();
    return v;
}

__global__ void __launch_bounds__(256) k_harris_select(const __grid_constant__ AfvParams P) {
    __shared__ int hist[256];
    __shared__ int wsum[8];
    __shared__ int s_thr, s_np, s_k, s_m;
    __shared__ uint32_t s_prefix, s_mask;
    const int l = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const AfvLevel& L = P.lv[l];
    const int cnt = min(P.counts[afv_cnt_idx(f, AFV_CNT_CAND, l)], L.cand_cap);
    const uint32_t* cand = L.cand + (long long)f * L.cand_cap;
    float* resp = L.cand_resp + (long long)f * L.cand_cap;
    uint2* det = L.det + (long long)f * L.det_cap;
    const uint8_t* img = L.img + (long long)f * L.img_fstride;

    hist[tid] = 0;
    if (tid == 0) s_m = 0;
    __syncthreads();
    for (int i = tid; i < cnt; i += 256) atomicAdd(&hist[cand[i] >> 24], 1);
    __syncthreads();
    {
        // retainBest(2q) on the FAST score: threshold = the largest score s with #(score >= s) >= 2q (ties of the 2q-th all stay)
        const int n2 = 2 * L.q_orb;
        const int S = suffix_sum_256(hist, wsum, tid);                   // S(tid) = #(score >= tid), non-increasing in tid
        if (tid == 0) { s_thr = 0; s_np = cnt; }
        __syncthreads();
        if (cnt > n2) {
            if (n2 == 0) { if (tid == 0) { s_thr = 256; s_np = 0; } }
            else {
                const int S1 = S - hist[tid];                             // S(tid + 1)
                if (S >= n2 && S1 < n2) { s_thr = tid; s_np = S; }       // exactly one thread
            }
        }
        __syncthreads();
    }
    const int thr = s_thr, np = s_np;
    // (45 % of the kernel's stall samples are the wait for the patch rows below; prefetching the rows of the thread's next candidate
    // into L1 one iteration ahead was measured SLOWER, 0.375 vs 0.338 ms per 512 frames: 18 more instructions per candidate and a
    // 122-register body that has to be capped back to 80 for 3 CTAs per SM.)
    for (int i = tid; i < cnt; i += 256) {
        const uint32_t c = cand[i];
        if ((int)(c >> 24) >= thr) resp[i] = harris7(img, L.w, L.h, L.img_stride, c & 0xfff, (c >> 12) & 0xfff);
    }
    __syncthreads();

    uint32_t kth = 0;                        // keep ord(resp) >= kth
    bool none = false;
    if (np > L.q_orb) {
        if (L.q_orb == 0) none = true;
        else {
            if (tid == 0) { s_prefix = 0; s_mask = 0; s_k = L.q_orb; }
            for (int pass = 3; pass >= 0; --pass) {
                hist[tid] = 0;
                __syncthreads();
                const uint32_t prefix = s_prefix, mask = s_mask;
                for (int i = tid; i < cnt; i += 256) {
                    if ((int)(cand[i] >> 24) < thr) continue;
                    const uint32_t key = f2ord(resp[i]);
                    if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255], 1);
                }
                __syncthreads();
                {
                    // digit = the largest d with #(digit >= d) >= k; the k-th largest then has rank k - #(digit > d) inside that digit
                    const int k = s_k;
                    const int S = suffix_sum_256(hist, wsum, tid);
                    const int S1 = S - hist[tid];
                    if (S >= k && S1 < k) {                               // exactly one thread (S is non-increasing, S(0) >= k)
                        s_k = k - S1;
                        s_prefix = prefix | ((uint32_t)tid << (8 * pass));
                        s_mask = mask | (0xffu << (8 * pass));
                    }
                }
                __syncthreads();
            }
            kth = s_prefix;
        }
    }
    if (!none) {
        for (int i = tid; i < cnt; i += 256) {
            const uint32_t c = cand[i];
            if ((int)(c >> 24) < thr) continue;
            const float r = resp[i];
            if (f2ord(r) >= kth) {
                const int slot = atomicAdd(&s_m, 1);
                if (slot < L.det_cap) det[slot] = make_uint2(c & 0xffffffu, __float_as_uint(r));
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        int m = s_m;
        if (m > L.det_cap) { atomicOr(&P.status[f], AFV_ST_DET_OVERFLOW); m = L.det_cap; }
        P.counts[afv_cnt_idx(f, AFV_CNT_DET, l)] = m;
    }
}

// ------------------------------------------------------------------------------------------------------
// K9: DistributeOctTree, one CTA per (frame, level); the list algorithm is in afv_octree.cuh.
// The per-node max-response pick prefers the smaller raster index (the oracle's canonical order).
// ------------------------------------------------------------------------------------------------------
// keys of one (frame, level) in shared memory: packed level coordinates (scaled to level-0 on the fly exactly as the float arrays
// were: x * scale, one rounding) + ordered response + node position + quadrant = 11 bytes per key
struct OctXYPacked {
    const uint32_t* pk; float scale;
    __device__ __forceinline__ float x(int k) const { return __fmul_rn((float)(pk[k] & 0xfffu), scale); }
    __device__ __forceinline__ float y(int k) const { return __fmul_rn((float)((pk[k] >> 12) & 0xfffu), scale); }
};
__global__ void __launch_bounds__(256) k_octree(const __grid_constant__ AfvParams P, int mcap, int ncap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* kpk = reinterpret_cast<uint32_t*>(smem_raw);           // [mcap] x | y << 12 (level coordinates)
    uint32_t* kresp = kpk + mcap;                                    // [mcap] ordered response
    unsigned short* knode = reinterpret_cast<unsigned short*>(kresp + mcap);   // [mcap] list position of the key's node
    unsigned char* kquad = reinterpret_cast<unsigned char*>(knode + mcap);    // [mcap]
    const size_t off = ((size_t)mcap * (4 + 4 + 2 + 1) + 15) & ~(size_t)15;
    OctWork W;
    oct_carve(smem_raw + off, ncap, W);
    unsigned long long* best = W.best;

    const int l = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const AfvLevel& L = P.lv[l];
    const int N = L.q_ext;
    int M = min(P.counts[afv_cnt_idx(f, AFV_CNT_DET, l)], L.det_cap);
    if (M > mcap) { if (tid == 0) atomicOr(&P.status[f], AFV_ST_OCTREE_OVERFLOW); M = mcap; }
    const uint2* det = L.det + (long long)f * L.det_cap;
    uint2* keep = L.keep + (long long)f * L.keep_cap;

    for (int k = tid; k < M; k += 256) {
        const uint2 d = det[k];
        kpk[k] = d.x & 0xffffffu;
        kresp[k] = f2ord(__uint_as_float(d.y));
    }
    __syncthreads();
    bool overflow = false;
    OctXYPacked xy; xy.pk = kpk; xy.scale = L.scale;
    int size = oct_distribute_xy(W, xy, knode, kquad, M, N, P.n_ini, P.hX, P.H, ncap, tid, overflow);
    if (overflow && tid == 0) atomicOr(&P.status[f], AFV_ST_OCTREE_OVERFLOW);

    // ---- best key per node: max response, ties -> smaller raster index
    for (int p = tid; p < size; p += 256) best[p] = 0ull;
    __syncthreads();
    for (int k = tid; k < M; k += 256) {
        const uint32_t pk = kpk[k];
        const uint32_t ras = ((pk >> 12) & 0xfffu) * (uint32_t)L.w + (pk & 0xfffu);
        atomicMax(&best[knode[k]], ((unsigned long long)kresp[k] << 32) | (unsigned long long)(0xffffffffu - ras));
    }
    __syncthreads();
    for (int p = tid; p < size; p += 256) {
        const unsigned long long b = best[p];
        const uint32_t ras = 0xffffffffu - (uint32_t)(b & 0xffffffffu);
        const int x = ras % L.w, y = ras / L.w;
        if (p < L.keep_cap) keep[p] = make_uint2((uint32_t)x | ((uint32_t)y << 12), __float_as_uint(ord2f((uint32_t)(b >> 32))));
    }
    if (tid == 0) {
        if (size > L.keep_cap) { atomicOr(&P.status[f], AFV_ST_OCTREE_OVERFLOW); size = L.keep_cap; }
        P.counts[afv_cnt_idx(f, AFV_CNT_KEEP, l)] = size;
    }
}

size_t afv_octree_smem_bytes(int mcap, int ncap) {
    const size_t off = ((size_t)mcap * (4 + 4 + 2 + 1) + 15) & ~(size_t)15;
    return off + oct_work_bytes(ncap);
}

// ------------------------------------------------------------------------------------------------------
// K7: 7x7 sigma=2 blur, float separable, arithmetic pinned to OpenCV's FMA-contracted filter engine:
//   row:    s = k0*p0; s = fma(kj, pj, s)                     column: s = k3*r0; s = fma(k3+j, r+j + r-j, s)
// Tile 128x16 outputs, 256 threads; each thread finishes 8 output pixels (2 x uchar4 stores).
// ------------------------------------------------------------------------------------------------------
#define BT_W AFV_TILE_W
#define BT_H AFV_TILE_H
__constant__ float c_g7[7] = {0x1.1f5f62p-4f, 0x1.0c70fcp-3f, 0x1.869472p-3f, 0x1.ba95cp-3f,
                              0x1.869472p-3f, 0x1.0c70fcp-3f, 0x1.1f5f62p-4f};

// Staged input: rows y0-3 .. y0+BT_H+2, byte columns x0-16 .. x0+BT_W+27 (160 B, 16-B aligned start) by ONE TMA box
// load; the TMA unit zero-fills outside the level, so border tiles then patch the few REFLECT_101 halo bytes they
// need (<= 3 rows above/below, <= 3 columns left/right).  Row pass: one thread = 4 outputs from 3 words; column
// pass: one thread = 4 columns x 4 rows from 10 float4 rows.
#define BT_WP 40                      // words per staged row
#define BT_XOFF 16                    // byte offset of pixel x0 inside a staged row
__device__ __forceinline__ float u8f(uint32_t w, int k) {   // (float) of byte k of w: 0x4B0000bb is 2^23 + bb
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 | k)), 8388608.0f);
}
__global__ void __launch_bounds__(256) k_blur(const __grid_constant__ AfvParams P, const __grid_constant__ AfvTmaps TM) {
    __shared__ __align__(128) uint32_t in[BT_H + 6][BT_WP];
    __shared__ __align__(16) float mid[BT_H + 6][BT_W];
    __shared__ __align__(8) uint64_t tma_bar;
    const AfvTile ti = P.tiles[blockIdx.x];
    const AfvLevel& L = P.lv[ti.level];
    const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int x0 = ti.x0, y0 = ti.y0;
    if (tid == 0) mbar_init(&tma_bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&tma_bar, (BT_H + 6) * BT_WP * 4);
        tma_load_3d(&in[0][0], &TM.m[ti.level], &tma_bar, x0 - 16, y0 - 3, f);
    }
    mbar_wait(&tma_bar, 0);
    // REFLECT_101 patch of the halo that fell outside the level.  More than half of all tiles touch a level border (128 x 32 tiles on
    // levels 640 .. 179 pixels wide), so the patch is kept off the common path of each case: the mirrored pixels are already in the
    // staged tile (reflection distance <= 3), rows first (top / bottom tiles only), then 3 columns per side with one thread per byte.
    const int xend = min(x0 + BT_W, L.w);                         // outputs exist for gx < xend, taps reach xend + 2
    {
        uint8_t* inb = reinterpret_cast<uint8_t*>(&in[0][0]);
        constexpr int PB = BT_WP * 4;
        const bool top = y0 == 0, bot = y0 + BT_H + 3 > L.h;
        if (top || bot) {                                         // uniform
            __syncthreads();
            // staged row r holds image row y0 - 3 + r; rows -3..-1 mirror rows 3..1, rows h..h+2 mirror rows h-2..h-4
            const int k = tid / 40, wq = tid - k * 40;            // 3 rows x 40 words per side
            if (k < 3) {
                if (top) in[2 - k][wq] = in[4 + k][wq];
                if (bot) { const int rd = L.h + k - (y0 - 3), rs = L.h - 2 - k - (y0 - 3); if (rd < BT_H + 6) in[rd][wq] = in[rs][wq]; }
            }
        }
        const bool lft = x0 == 0, rgt = xend + 3 > L.w;
        if (lft || rgt) {                                         // uniform
            __syncthreads();
            const int r = tid / 6, k = tid - r * 6;               // 38 rows x (3 left + 3 right) bytes
            if (r < BT_H + 6) {
                uint8_t* dst = inb + r * PB + BT_XOFF - x0;       // dst[gx]
                if (k < 3) { if (lft) dst[-1 - k] = dst[1 + k]; }
                else if (rgt && L.w + (k - 3) < xend + 3) dst[L.w + (k - 3)] = dst[L.w - 2 - (k - 3)];
            }
        }
    }
    __syncthreads();
    // row pass: outputs 4q..4q+3 need staged bytes 4q+1 .. 4q+10; warp w takes rows w, w+8, ..., lane = quad
    for (int r = wrp; r < BT_H + 6; r += 8) {
        const int q = lane;
        const uint32_t w0 = in[r][q + 3], w1 = in[r][q + 4], w2 = in[r][q + 5];      // staged bytes 4q+13 .. 4q+22
        // byte -> float without the conversion unit: one PRMT drops the byte into the mantissa of 2^23 and one FADD removes the 2^23
        // (exact).  The I2F / F2I conversions were this kernel's busiest pipe (XU 55 % in ncu r02A, 16 lanes per clock per SM).
        float pf[10];
        pf[0] = u8f(w0, 1); pf[1] = u8f(w0, 2); pf[2] = u8f(w0, 3);
        pf[3] = u8f(w1, 0); pf[4] = u8f(w1, 1); pf[5] = u8f(w1, 2); pf[6] = u8f(w1, 3);
        pf[7] = u8f(w2, 0); pf[8] = u8f(w2, 1); pf[9] = u8f(w2, 2);
        float4 o;
        float* op = &o.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float sacc = __fmul_rn(c_g7[0], pf[k]);
#pragma unroll
            for (int j = 1; j < 7; ++j) sacc = __fmaf_rn(c_g7[j], pf[k + j], sacc);
            op[k] = sacc;
        }
        *reinterpret_cast<float4*>(&mid[r][4 * q]) = o;
    }
    __syncthreads();
    uint8_t* out = L.blur + (long long)f * L.fstride;
    {
        const int q = lane, rg = wrp;                             // 32 column quads x 8 groups of 4 rows
        const int gx = x0 + 4 * q;
        if (gx < L.w) {
            float4 m[10];
#pragma unroll
            for (int j = 0; j < 10; ++j) m[j] = *reinterpret_cast<const float4*>(&mid[4 * rg + j][4 * q]);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int gy = y0 + 4 * rg + rr;
                if (gy >= L.h) break;
                uint32_t ub[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float sacc = __fmul_rn(c_g7[3], (&m[rr + 3].x)[k]);
#pragma unroll
                    for (int j = 1; j <= 3; ++j)
                        sacc = __fmaf_rn(c_g7[3 + j], __fadd_rn((&m[rr + 3 + j].x)[k], (&m[rr + 3 - j].x)[k]), sacc);
                    // cv::saturate_cast<uchar>(float) = round half to even, clamp.  0 <= sacc <= 254.99997 (positive taps whose float sum
                    // is below 1, u8 inputs), so adding 1.5 * 2^23 leaves rint(sacc) in the low mantissa byte and the clamp never acts
                    ub[k] = __float_as_uint(__fadd_rn(sacc, 12582912.0f));
                }
                const uint32_t pk = __byte_perm(__byte_perm(ub[0], ub[1], 0x0040), __byte_perm(ub[2], ub[3], 0x0040), 0x5410);
                *reinterpret_cast<uint32_t*>(out + (long long)gy * L.stride + gx) = pk;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K6/K8/K10: one warp per kept keypoint: IC angle (31 rows of the radius-15 disc, lane = column), steered
// BRIEF (lane = descriptor byte), and the merged cv::KeyPoint / descriptor / size rows (levels ascending).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float p1 = 0x1.ca44dep+5f, p3 = -0x1.2aaddcp+4f, p5 = 0x1.1d3f7ep+3f, p7 = -0x1.4515b2p+1f;
    const float eps = 0x1p-52f;                                   // (float)DBL_EPSILON
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

// 4 CTAs (32 warps) per SM at 64 registers: measured 0.635 ms / 512 frames against 0.652 at 5 CTAs / 48 registers and 0.678 at 6 / 40
// (spills); the kernel is bound by instruction issue, not by latency (prefetching the next keypoint's rows made it slower).
__global__ void __launch_bounds__(256, 4) k_describe(const __grid_constant__ AfvParams P, afv_keypoint* __restrict__ kps,
                                                  uint8_t* __restrict__ desc, float* __restrict__ kpsize,
                                                  int* __restrict__ n_out) {
    __shared__ float2 patf[8][2][32];        // patf[k][j][lane] = point j (x, y) of test k of descriptor byte `lane`, as floats
    __shared__ uint32_t spatch[8][372];      // per warp: 37 rows x 40 bytes of the blurred level around the keypoint
    __shared__ int lvl_start[AFV_MAX_LEVELS + 1];
    const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {   // the int8 -> float conversion of the 512 pattern points is done once per CTA, not once per tap (6 % of the kernel):
        // thread t converts test t = bit (t & 7) of descriptor byte (t >> 3)
        const int pl = tid >> 3, pk = tid & 7;
        const uint32_t pw = reinterpret_cast<const uint32_t*>(g_pattern)[tid];
        patf[pk][0][pl] = make_float2((float)(int)(int8_t)pw, (float)(int)(int8_t)(pw >> 8));
        patf[pk][1][pl] = make_float2((float)(int)(int8_t)(pw >> 16), (float)(int)(int8_t)(pw >> 24));
    }
    if (warp == 0) {                         // level offsets of the merged output: one count per lane, warp scan (was a serial loop of loads)
        const int c = lane < P.nlevels ? min(P.counts[afv_cnt_idx(f, AFV_CNT_KEEP, lane)], P.lv[lane].keep_cap) : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane <= P.nlevels) lvl_start[lane] = incl - c;       // lane == nlevels holds the total (its own count is 0)
        if (blockIdx.x == 0 && lane == 31) {
            if (incl > P.out_cap) atomicOr(&P.status[f], AFV_ST_OUT_OVERFLOW);
            n_out[f] = min(incl, P.out_cap);
        }
    }
    __syncthreads();
    const int total = min(lvl_start[P.nlevels], P.out_cap);
    // 32 keypoints per CTA, 4 per warp: the prologue above and its barrier were 20 % of the stall samples at 8 keypoints per CTA
    for (int it = 0; it < 4; ++it) {
    const int i = blockIdx.x * 32 + it * 8 + warp;
    if (i >= total) break;
    int l = 0;
    while (i >= lvl_start[l + 1]) ++l;
    const AfvLevel& L = P.lv[l];
    const uint2 kd = (L.keep + (long long)f * L.keep_cap)[i - lvl_start[l]];
    const int x0 = kd.x & 0xfff, y0 = (kd.x >> 12) & 0xfff;
    const uint8_t* img = L.img + (long long)f * L.img_fstride;
    const uint8_t* blr = L.blur + (long long)f * L.fstride;

    // intensity-centroid moments over the radius-15 disc (cv::ORB ICAngles); lane = u + 15.  Fully unrolled so the
    // 31 independent row loads are in flight together (the loop was latency-bound on one load per iteration).
    // (Measured and rejected: lane = ROW with nine aligned words per row, a chord byte mask and 8 + 8 dp4a -- a third of the
    // instructions, but every warp-wide load then touches 31 rows = 31 sectors instead of one or two: 0.875 vs 0.633 ms per 512 frames.)
    int m10 = 0, m01 = 0;
    const int u = lane - 15;
    const int au = u < 0 ? -u : u;
    const bool inner = (x0 >= 15 && y0 >= 15 && x0 + 16 < L.w && y0 + 15 < L.h);
    {
        constexpr int UM[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
        int vals[31];
        if (inner) {
            const uint8_t* base = img + (long long)(y0 - 15) * L.img_stride + x0 + u;
#pragma unroll
            for (int r = 0; r < 31; ++r) vals[r] = base[(long long)r * L.img_stride];
        } else {
            const int xx = refl101(x0 + u, L.w);
#pragma unroll
            for (int r = 0; r < 31; ++r) vals[r] = img[(long long)refl101(y0 - 15 + r, L.h) * L.img_stride + xx];
        }
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
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    const float fx = __fmul_rn((float)x0, L.scale), fy = __fmul_rn((float)y0, L.scale);
    // descriptor centre as cv::ORB::compute recovers it from the scaled keypoint
    const int cx = __float2int_rn(__fmul_rn(fx, L.inv_scale)), cy = __float2int_rn(__fmul_rn(fy, L.inv_scale));
    const float ang = __fmul_rn(angle, 0x1.1df46ap-6f);            // (float)(CV_PI/180.f)
    double sd, cd;
    sincos((double)ang, &sd, &cd);                                // one argument reduction for both (was 10 % of the kernel as cos + sin)
    const float a = (float)cd, b = (float)sd;
    const bool dinner = (cx >= 19 && cy >= 19 && cx + 19 < L.w && cy + 19 < L.h);
    uint32_t byte = 0;
    if (dinner) {                            // warp-uniform: every tap lies inside the blurred level (rotated pattern radius < 19)
        // The 512 taps of a keypoint are scattered over a 37 x 37 patch: as global byte loads each of the 16 warp-wide gathers touches
        // ~20 sectors and the kernel sat at 73 % of the L1 data pipe (ncu r02A).  The patch is copied once -- 37 rows x 10 aligned
        // words, 12 coalesced LDG.32 -- into this warp's shared-memory tile and the gathers become LDS.U8.
        uint32_t* sp = spatch[warp];
        const uint8_t* org = blr + (long long)(cy - 18) * L.stride + (cx - 18);
        const int shift = (int)((uintptr_t)org & 3);             // rows share it: the level stride is a multiple of 128
        const uint8_t* al = org - shift;
        const int stride = L.stride;
#pragma unroll
        for (int half = 0; half < 2; ++half) {                   // 6 loads in flight at a time
            uint32_t wv[6];
#pragma unroll
            for (int it2 = 0; it2 < 6; ++it2) {
                const int item = (half * 6 + it2) * 32 + lane, row = item / 10, col = item - row * 10;
                wv[it2] = item < 370 ? *reinterpret_cast<const uint32_t*>(al + row * stride + col * 4) : 0u;
            }
#pragma unroll
            for (int it2 = 0; it2 < 6; ++it2) { const int item = (half * 6 + it2) * 32 + lane; if (item < 370) sp[item] = wv[it2]; }
        }
        __syncwarp();
        const uint8_t* ctr = reinterpret_cast<const uint8_t*>(sp) + 18 * 40 + 18 + shift;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int tv[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float2 pt = patf[k][j][lane];
                const int dx = __float2int_rn(__fsub_rn(__fmul_rn(pt.x, a), __fmul_rn(pt.y, b)));
                const int dy = __float2int_rn(__fadd_rn(__fmul_rn(pt.x, b), __fmul_rn(pt.y, a)));
                tv[j] = ctr[dy * 40 + dx];
            }
            byte |= (uint32_t)(tv[0] < tv[1]) << k;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int tv[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float2 pt = patf[k][j][lane];
                const int ix = cx + __float2int_rn(__fsub_rn(__fmul_rn(pt.x, a), __fmul_rn(pt.y, b)));
                const int iy = cy + __float2int_rn(__fadd_rn(__fmul_rn(pt.x, b), __fmul_rn(pt.y, a)));
                if (ix >= 0 && ix < L.w && iy >= 0 && iy < L.h) tv[j] = blr[(long long)iy * L.stride + ix];
                else tv[j] = img[(long long)refl101(iy, L.h) * L.img_stride + refl101(ix, L.w)];
            }
            byte |= (uint32_t)(tv[0] < tv[1]) << k;
        }
    }
    const long long o = (long long)f * P.out_cap + i;
    desc[o * 32 + lane] = (uint8_t)byte;
    if (lane == 0) {
        afv_keypoint kp;
        kp.x = fx; kp.y = fy; kp.size = L.kp_size; kp.angle = angle;
        kp.response = __uint_as_float(kd.y); kp.octave = l; kp.class_id = -1;
        kps[o] = kp;
        if (kpsize) kpsize[o] = L.size_norm;
    }
    __syncwarp();                            // spatch is rewritten by the next keypoint
    }
}

// ------------------------------------------------------------------------------------------------------
// host-side launch sequence for one batch (all on one stream, no host synchronisation inside)
// ------------------------------------------------------------------------------------------------------
// Octree capacities belong to the extractor (AfvParams::oct_mcap / oct_ncap): the reference builds two extractors with
// different nfeatures (src/Tracking.cc:78,84).  The dynamic shared-memory attribute of k_octree is per device and only
// ever raised to the running maximum, so an earlier, larger extractor keeps working after a smaller one is created.
#include <mutex>
static std::mutex g_oct_mu;
static size_t g_oct_smem_set[64] = {0};                  // per device

int afv_orb_configure(int max_det_cap, int max_keep_cap, int* mcap_out, int* ncap_out) {
    const int mcap = max_det_cap, ncap = max_keep_cap + 8;
    size_t smem = afv_octree_smem_bytes(mcap, ncap);
    if (smem > 227 * 1024) { afv_set_error("octree shared memory %zu B exceeds 227 KB (nfeatures too large)", smem); return AFV_ERR_INVALID; }
    int dev = 0;
    AFV_CUDA_CHECK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(g_oct_mu);
        if (dev < 0 || dev >= 64 || smem > g_oct_smem_set[dev]) {
            AFV_CUDA_CHECK(cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) g_oct_smem_set[dev] = smem;
        }
    }
    *mcap_out = mcap; *ncap_out = ncap;
    return AFV_OK;
}

typedef CUresult (*afv_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static afv_encode_tiled_fn g_encode = nullptr;

// 3-D u8 tensor map of one level: (x, y, frame) with the k_fast staging box.
static int make_level_tmap(CUtensorMap* m, const AfvLevel& L, int B, int box_w, int box_h) {
    if (!g_encode) {
        void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) {
            afv_set_error("cuTensorMapEncodeTiled not available from the driver"); return AFV_ERR_CUDA;
        }
        g_encode = (afv_encode_tiled_fn)fn;
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)L.w, (cuuint64_t)L.h, (cuuint64_t)B};
    const cuuint64_t gstr[2] = {(cuuint64_t)L.img_stride, (cuuint64_t)L.img_fstride};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(L.img), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { afv_set_error("cuTensorMapEncodeTiled failed (%d) for a %dx%d level, stride %d", (int)r, L.w, L.h, L.img_stride); return AFV_ERR_CUDA; }
    return AFV_OK;
}

int afv_launch_extract(const AfvParams& P, afv_keypoint* d_kps, uint8_t* d_desc, float* d_kpsize,
                       int* d_n_out, cudaStream_t st, const AfvAux& aux) {
    AfvTmaps TM, TMB;                         // k_fast / k_blur staging boxes
    memset(&TM, 0, sizeof(TM)); memset(&TMB, 0, sizeof(TMB));
    for (int l = 0; l < P.nlevels; ++l) {
        int rc = make_level_tmap(&TM.m[l], P.lv[l], P.B, FT_WP * 4, FT_PH);
        if (!rc) rc = make_level_tmap(&TMB.m[l], P.lv[l], P.B, BT_WP * 4, BT_H + 6);
        if (rc) return rc;
    }
    const int acc = P.ntiles;
    cudaMemsetAsync(P.counts, 0, sizeof(int) * 4 * AFV_MAX_LEVELS * P.B, st);
    cudaMemsetAsync(P.status, 0, sizeof(int) * P.B, st);
    for (int l = 1; l < P.nlevels; ++l) {
        dim3 g((P.lv[l].w + 127) / 128, (P.lv[l].h + RS_OUT_H - 1) / RS_OUT_H, P.B);
        CUtensorMap tms;
        { const int rc = make_level_tmap(&tms, P.lv[l - 1], P.B, RS_PITCH, RS_ROWS); if (rc) return rc; }
        AfvProfScope ps("k_resize", st);
        k_resize<<<g, 256, 0, st>>>(P, tms, l);
        ++g_afv_launches;
    }
    // After k_fast the work forks: the blur (ALU bound, 275 k tiles per 512 frames) stays on the caller's stream, the selection kernels
    // (latency bound: one CTA per (frame, level), occupancy 37 %) go to the extractor's HIGH-PRIORITY side stream so that the block
    // scheduler places their CTAs first and the blur fills the remaining slots; k_describe joins both.  (A low-priority blur on the
    // side stream did not overlap: its tiles were queued first and the selection CTAs waited behind them.)  With per-kernel profiling
    // on, everything runs on the caller's stream so that the event times are additive.
    { AfvProfScope ps("k_fast", st); k_fast<<<dim3(acc, P.B), 256, 0, st>>>(P, TM); ++g_afv_launches; }
    const cudaStream_t sst = afv_prof_is_on() ? st : aux.stream;
    cudaEventRecord(aux.ev_pyr, st);
    cudaStreamWaitEvent(sst, aux.ev_pyr, 0);
    { AfvProfScope ps("k_harris_select", sst); k_harris_select<<<dim3(P.nlevels, P.B), 256, 0, sst>>>(P); ++g_afv_launches; }
    { AfvProfScope ps("k_octree", sst);
      k_octree<<<dim3(P.nlevels, P.B), 256, afv_octree_smem_bytes(P.oct_mcap, P.oct_ncap), sst>>>(P, P.oct_mcap, P.oct_ncap);
      ++g_afv_launches; }
    cudaEventRecord(aux.ev_blur, sst);
    { AfvProfScope ps("k_blur", st); k_blur<<<dim3(acc, P.B), 256, 0, st>>>(P, TMB); ++g_afv_launches; }
    cudaStreamWaitEvent(st, aux.ev_blur, 0);
    { AfvProfScope ps("k_describe", st);
      k_describe<<<dim3((P.out_cap + 31) / 32, P.B), 256, 0, st>>>(P, d_kps, d_desc, d_kpsize, d_n_out); ++g_afv_launches; }
    return AFV_OK;
}

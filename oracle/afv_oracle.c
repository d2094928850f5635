/*
 * afv_oracle.c -- CPU ORACLE for the orb32 extraction path (test infrastructure, NOT product code).
 *
 * Restates, in plain C, what `FeatureExtractor_orb32` does per frame (reference:
 * src/Feature_orb32.cpp:11-65, src/FeatureExtractor.cpp:97-172,276-308, src/ORBextractor.cc:181-458).
 * The numerics behind `cv::ORB::detect/compute` live in OpenCV (un-vendored, unpinned by the reference);
 * they are restated here from OpenCV's algorithm and pinned bit-for-bit to cv2 4.13.0
 * (tests/test_oracle_golden.py, tests/golden/).  Build with -ffp-contract=off: every float expression
 * below is meant to round after each operation (no FMA), which is what the pinned binary does.
 */
#include "afv_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

static const int8_t kPattern[256 * 4] = {
#include "orb_pattern.inc"
};

static inline int cv_round_f(float v) { return (int)lrintf(v); }      /* cvRound: round-half-even */
static inline int refl101(int i, int n) {
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}
static inline int px_r(const uint8_t* img, int w, int h, int stride, int x, int y) {
    return img[(long)refl101(y, h) * stride + refl101(x, w)];
}

/* ---------------------------------------------------------------- geometry / quotas -------------- */
/* cv::ORB layer geometry: scale_l = (float)pow((double)scaleFactor, l); size = cvRound(dim * (1.f/scale)). */
int orc_orb_level_geometry(int w, int h, int nlevels, float sf, int* lw, int* lh, float* lscale) {
    for (int l = 0; l < nlevels; ++l) {
        float scale = (float)pow((double)sf, (double)l);
        float inv = 1.0f / scale;
        lscale[l] = scale;
        lw[l] = cv_round_f((float)w * inv);
        lh[l] = cv_round_f((float)h * inv);
    }
    return 0;
}

/* reference src/FeatureExtractor.cpp:97-108 (and the identical formula inside cv::ORB). */
void orc_features_per_level(int nfeatures, int nlevels, float scale_factor, int* quota) {
    float factor = 1.0f / scale_factor;
    float nDesired = (float)nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        quota[l] = cv_round_f(nDesired);
        sum += quota[l];
        nDesired *= factor;
    }
    int rest = nfeatures - sum;
    quota[nlevels - 1] = rest > 0 ? rest : 0;
}

/* ---------------------------------------------------------------- INTER_LINEAR_EXACT ------------- */
/* OpenCV resize_bitExact<uchar, interpolationLinear>: 8.8 fixed-point coefficients from
 * f = (1/inv_scale)*(d+0.5)-0.5 (double, two roundings), c1 = rint(frac*256), c0 = 256-c1;
 * horizontal pass keeps 8.8, vertical pass rounds (v + 2^15) >> 16. Out-of-range taps replicate the edge. */
static void lin_exact_coeffs(int ssize, int dsize, int* ofs, int* c0, int* c1) {
    double inv_scale = (double)dsize / (double)ssize;
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        double f = scale * ((double)d + 0.5) - 0.5;
        int i = (int)floor(f);
        if (i >= 0 && ssize > 1) {
            if (i < ssize - 1) {
                int k1 = (int)lrint((f - (double)i) * 256.0);
                ofs[d] = i; c1[d] = k1; c0[d] = 256 - k1;
            } else { ofs[d] = ssize - 1; c0[d] = 256; c1[d] = 0; }
        } else { ofs[d] = 0; c0[d] = 256; c1[d] = 0; }
    }
}

void orc_resize_linear_exact_u8(const uint8_t* src, int sw, int sh, int sstride,
                                uint8_t* dst, int dw, int dh, int dstride) {
    int* xo = (int*)malloc(sizeof(int) * 3 * dw);
    int* yo = (int*)malloc(sizeof(int) * 3 * dh);
    int *xc0 = xo + dw, *xc1 = xo + 2 * dw, *yc0 = yo + dh, *yc1 = yo + 2 * dh;
    lin_exact_coeffs(sw, dw, xo, xc0, xc1);
    lin_exact_coeffs(sh, dh, yo, yc0, yc1);
    for (int y = 0; y < dh; ++y) {
        const uint8_t* r0 = src + (long)yo[y] * sstride;
        const uint8_t* r1 = src + (long)(yo[y] + 1 < sh ? yo[y] + 1 : sh - 1) * sstride;
        for (int x = 0; x < dw; ++x) {
            int i0 = xo[x], i1 = i0 + 1 < sw ? i0 + 1 : sw - 1;
            uint32_t h0 = (uint32_t)(xc0[x] * r0[i0] + xc1[x] * r0[i1]);
            uint32_t h1 = (uint32_t)(xc0[x] * r1[i0] + xc1[x] * r1[i1]);
            uint32_t v = (uint32_t)yc0[y] * h0 + (uint32_t)yc1[y] * h1;
            dst[(long)y * dstride + x] = (uint8_t)((v + (1u << 15)) >> 16);
        }
    }
    free(xo); free(yo);
}

long orc_orb_pyramid(const uint8_t* gray, int w, int h, int stride, int nlevels, float sf,
                     uint8_t* out, long* offs, int* lw, int* lh, float* lscale) {
    orc_orb_level_geometry(w, h, nlevels, sf, lw, lh, lscale);
    long total = 0;
    for (int l = 0; l < nlevels; ++l) { offs[l] = total; total += (long)lw[l] * lh[l]; }
    if (!out) return total;
    for (int y = 0; y < h; ++y) memcpy(out + (long)y * w, gray + (long)y * stride, (size_t)w);
    for (int l = 1; l < nlevels; ++l)
        orc_resize_linear_exact_u8(out + offs[l - 1], lw[l - 1], lh[l - 1], lw[l - 1],
                                   out + offs[l], lw[l], lh[l], lw[l]);
    return total;
}

/* ---------------------------------------------------------------- FAST-9/16 + NMS ---------------- */
/* Bresenham circle of radius 3, in order around the circle (OpenCV makeOffsets, patternSize 16). */
static const int kCirc[16][2] = {
    {0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
    {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

/* OpenCV's cornerScore<16>: the largest threshold t' for which the pixel still passes the 9-contiguous
 * test, i.e. max over the 16 arcs of 9 of min(v-p) (darker) / min(p-v) (brighter), minus 1.
 * Returns 0 when the pixel is not a corner at `threshold`. */
static inline int fast_score(const uint8_t* p, int stride, int threshold) {
    int v = p[0], d[25];
    for (int k = 0; k < 16; ++k) d[k] = v - p[kCirc[k][1] * stride + kCirc[k][0]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = -256;
    for (int s = 0; s < 16; ++s) {
        int mn = d[s], mx = d[s];
        for (int k = 1; k < 9; ++k) { int t = d[s + k]; if (t < mn) mn = t; if (t > mx) mx = t; }
        if (mn > best) best = mn;          /* all darker than v by at least mn */
        if (-mx > best) best = -mx;        /* all brighter than v by at least -mx */
    }
    return best > threshold ? best - 1 : 0;
}

int orc_fast9_16_nms(const uint8_t* img, int w, int h, int stride, int threshold,
                     int* xs, int* ys, int* scores, int cap) {
    if (threshold < 0) threshold = 0;
    if (threshold > 255) threshold = 255;
    int n = 0;
    if (w < 7 || h < 7) return 0;
    uint8_t* sc = (uint8_t*)calloc((size_t)w * h, 1);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            /* cheap necessary condition first (any arc of 9 holds one pixel of each antipodal pair), as
             * OpenCV's scalar loop does; fast_score decides. */
            const uint8_t* p = img + (long)y * stride + x;
            int v = p[0], hi = v + threshold, lo = v - threshold;
            int a = p[3 * stride], b = p[-3 * stride];
            if (!(a > hi || a < lo || b > hi || b < lo)) continue;
            a = p[3]; b = p[-3];
            if (!(a > hi || a < lo || b > hi || b < lo)) continue;
            a = p[2 * stride + 2]; b = p[-2 * stride - 2];
            if (!(a > hi || a < lo || b > hi || b < lo)) continue;
            a = p[-2 * stride + 2]; b = p[2 * stride - 2];
            if (!(a > hi || a < lo || b > hi || b < lo)) continue;
            sc[(long)y * w + x] = (uint8_t)fast_score(p, stride, threshold);
        }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = sc[(long)y * w + x];
            if (!s) continue;
            const uint8_t* q = sc + (long)y * w + x;
            if (s > q[-1] && s > q[1] && s > q[-w - 1] && s > q[-w] && s > q[-w + 1] &&
                s > q[w - 1] && s > q[w] && s > q[w + 1]) {
                if (n < cap) { xs[n] = x; ys[n] = y; scores[n] = s; }
                ++n;
            }
        }
    free(sc);
    return n;
}

/* ---------------------------------------------------------------- Harris, IC angle --------------- */
/* cv::ORB HarrisResponses(blockSize=7, k=0.04f): int sums of Sobel-like gradients, float32 score. */
float orc_harris7(const uint8_t* img, int w, int h, int stride, int x0, int y0) {
    int a = 0, b = 0, c = 0;
    for (int dy = -3; dy <= 3; ++dy)
        for (int dx = -3; dx <= 3; ++dx) {
            int x = x0 + dx, y = y0 + dy;
            int p00 = px_r(img, w, h, stride, x - 1, y - 1), p01 = px_r(img, w, h, stride, x, y - 1),
                p02 = px_r(img, w, h, stride, x + 1, y - 1), p10 = px_r(img, w, h, stride, x - 1, y),
                p12 = px_r(img, w, h, stride, x + 1, y), p20 = px_r(img, w, h, stride, x - 1, y + 1),
                p21 = px_r(img, w, h, stride, x, y + 1), p22 = px_r(img, w, h, stride, x + 1, y + 1);
            int Ix = (p12 - p10) * 2 + (p02 - p00) + (p22 - p20);
            int Iy = (p21 - p01) * 2 + (p20 - p00) + (p22 - p02);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
    float scale = 1.f / ((1 << 2) * 7 * 255.f);
    float scale_sq_sq = scale * scale * scale * scale;
    float fa = (float)a, fb = (float)b, fc = (float)c;
    float t0 = fa * fb;
    float t1 = fc * fc;
    float s = fa + fb;
    float t2 = 0.04f * s;
    t2 = t2 * s;
    float r = t0 - t1;
    r = r - t2;
    return r * scale_sq_sq;
}

/* scalar cv::fastAtan2 (degrees, [0,360)). */
float orc_fast_atan2(float y, float x) {
    const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

static const int kUmax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

/* cv::ORB ICAngles, half patch 15. */
float orc_ic_angle(const uint8_t* img, int w, int h, int stride, int x0, int y0) {
    int m01 = 0, m10 = 0;
    for (int u = -15; u <= 15; ++u) m10 += u * px_r(img, w, h, stride, x0 + u, y0);
    for (int v = 1; v <= 15; ++v) {
        int vsum = 0, d = kUmax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = px_r(img, w, h, stride, x0 + u, y0 + v), vm = px_r(img, w, h, stride, x0 + u, y0 - v);
            vsum += vp - vm;
            m10 += u * (vp + vm);
        }
        m01 += v * vsum;
    }
    return orc_fast_atan2((float)m01, (float)m10);
}

/* ---------------------------------------------------------------- blur + rBRIEF ------------------ */
/* GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) as cv::ORB::compute runs it on a level ROI of the padded
 * pyramid buffer: the non-isolated-submatrix case takes OpenCV's float separable filter engine
 * (RowFilter<uchar,float> then SymmColumnFilter<Cast<float,uchar>>), kernel = getGaussianKernel(7,2,CV_32F).
 * Pinned arithmetic (cv2 4.13.0 on an AVX2+FMA3 x86-64 host, i.e. the FMA-contracted dispatch of
 * filter.simd.hpp; probe: 0 differing pixels on every level of the toy + synthetic frames, while the
 * un-fused or non-symmetric orders differ in 1-6 px per level):
 *   row:    s = k0*p0;  s = fma(k_j, p_j, s)  j=1..6            (left to right)
 *   column: s = k3*r0;  s = fma(k_{3+j}, r_{+j} + r_{-j}, s) j=1..3   (symmetric form)
 *   out = saturate_u8(rint(s)) (round-half-even). */
static const float kG7[7] = {0x1.1f5f62p-4f, 0x1.0c70fcp-3f, 0x1.869472p-3f, 0x1.ba95cp-3f,
                             0x1.869472p-3f, 0x1.0c70fcp-3f, 0x1.1f5f62p-4f};

void orc_blur7_level(const uint8_t* img, int w, int h, int stride, uint8_t* out, int ostride) {
    float* rows = (float*)malloc(sizeof(float) * (size_t)w * (h + 6));
    for (int yy = -3; yy < h + 3; ++yy) {
        const uint8_t* r = img + (long)refl101(yy, h) * stride;
        float* o = rows + (long)(yy + 3) * w;
        for (int x = 0; x < w; ++x) {
            float s = kG7[0] * (float)r[refl101(x - 3, w)];
            for (int k = 1; k < 7; ++k) s = fmaf(kG7[k], (float)r[refl101(x - 3 + k, w)], s);
            o[x] = s;
        }
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float* c = rows + (long)(y + 3) * w + x;
            float s = kG7[3] * c[0];
            for (int k = 1; k <= 3; ++k) s = fmaf(kG7[3 + k], c[(long)k * w] + c[-(long)k * w], s);
            int v = (int)lrintf(s);
            out[(long)y * ostride + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    free(rows);
}

/* cv::ORB computeOrbDescriptors (WTA_K 2): taps inside the level come from the blurred level, taps in the
 * 23-px REFLECT_101 padding come from the un-blurred level (the padding is never blurred). */
void orc_rbrief32(const uint8_t* img, const uint8_t* blur, int w, int h, int stride,
                  int cx, int cy, float angle_deg, uint8_t* desc) {
    float angle = angle_deg * (float)(M_PI / 180.f);
    float a = (float)cos(angle), b = (float)sin(angle);
    const int8_t* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            int t[2];
            for (int j = 0; j < 2; ++j) {
                float fx = (float)pat[4 * k + 2 * j], fy = (float)pat[4 * k + 2 * j + 1];
                float rx = fx * a - fy * b;
                float ry = fx * b + fy * a;
                int x = cx + cv_round_f(rx), y = cy + cv_round_f(ry);
                if (x >= 0 && x < w && y >= 0 && y < h) t[j] = blur[(long)y * stride + x];
                else t[j] = px_r(img, w, h, stride, x, y);
            }
            val |= (t[0] < t[1]) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ---------------------------------------------------------------- cv::ORB::detect, one level ----- */
static int cmp_int_desc(const void* a, const void* b) { return *(const int*)b - *(const int*)a; }
static int cmp_float_desc(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x < y) - (x > y);
}

/* KeyPointsFilter::retainBest semantics: keep the n best AND everything tying with the n-th. Order kept. */
int orc_orb_detect_level(const uint8_t* img, int w, int h, int stride, int fast_th, int quota,
                         int* xs, int* ys, float* harris, int* fastscore, int cap) {
    int n = orc_fast9_16_nms(img, w, h, stride, fast_th, xs, ys, fastscore, cap);
    if (n > cap) return -n;                       /* caller's buffers too small */
    int n2 = 2 * quota;
    if (n > n2) {
        if (n2 == 0) return 0;
        int* tmp = (int*)malloc(sizeof(int) * n);
        memcpy(tmp, fastscore, sizeof(int) * n);
        qsort(tmp, n, sizeof(int), cmp_int_desc);
        int thr = tmp[n2 - 1], m = 0;
        free(tmp);
        for (int i = 0; i < n; ++i)
            if (fastscore[i] >= thr) { xs[m] = xs[i]; ys[m] = ys[i]; fastscore[m] = fastscore[i]; ++m; }
        n = m;
    }
    for (int i = 0; i < n; ++i) harris[i] = orc_harris7(img, w, h, stride, xs[i], ys[i]);
    if (n > quota) {
        if (quota == 0) return 0;
        float* tmp = (float*)malloc(sizeof(float) * n);
        memcpy(tmp, harris, sizeof(float) * n);
        qsort(tmp, n, sizeof(float), cmp_float_desc);
        float thr = tmp[quota - 1];
        int m = 0;
        free(tmp);
        for (int i = 0; i < n; ++i)
            if (harris[i] >= thr) {
                xs[m] = xs[i]; ys[m] = ys[i]; fastscore[m] = fastscore[i]; harris[m] = harris[i]; ++m;
            }
        n = m;
    }
    return n;
}

/* ---------------------------------------------------------------- DistributeOctTree -------------- */
/* Restatement of src/ORBextractor.cc:181-237 (DivideNode) and :239-458 (DistributeOctTree).
 * The node list is a doubly-linked list over a node pool (push_front / erase as in the reference).
 * Tie rules the reference leaves to the allocator / to std::nth_element are fixed here:
 *  - sort of (nKeys, node*) pairs (:381): equal nKeys ordered by creation sequence (later = larger);
 *  - max-response pick (:444-455): first key in node order wins; node order = input order (stable). */
typedef struct {
    int ulx, uly, urx, brx, bry, bly;               /* UL, UR.x, BR, BL.y (axis-aligned box) */
    int* keys; int nkeys;
    int nomore, prev, next, seq;
} onode;

typedef struct { onode* pool; int used, cap, head, tail, size, seq; } olist;

static int ol_new(olist* L) {
    if (L->used == L->cap) { L->cap *= 2; L->pool = (onode*)realloc(L->pool, sizeof(onode) * L->cap); }
    onode* n = &L->pool[L->used];
    memset(n, 0, sizeof(*n));
    n->prev = n->next = -1; n->seq = L->seq++;
    return L->used++;
}
static void ol_push_front(olist* L, int id) {
    onode* n = &L->pool[id];
    n->prev = -1; n->next = L->head;
    if (L->head >= 0) L->pool[L->head].prev = id; else L->tail = id;
    L->head = id; L->size++;
}
static void ol_push_back(olist* L, int id) {
    onode* n = &L->pool[id];
    n->next = -1; n->prev = L->tail;
    if (L->tail >= 0) L->pool[L->tail].next = id; else L->head = id;
    L->tail = id; L->size++;
}
static int ol_erase(olist* L, int id) {          /* returns the following node */
    onode* n = &L->pool[id];
    int nx = n->next;
    if (n->prev >= 0) L->pool[n->prev].next = n->next; else L->head = n->next;
    if (n->next >= 0) L->pool[n->next].prev = n->prev; else L->tail = n->prev;
    L->size--;
    free(n->keys); n->keys = NULL;
    return nx;
}

/* DivideNode: four children, keys routed by float compares against the (int) split lines. */
static void divide_node(olist* L, int id, const float* px, const float* py, int kid[4]) {
    int ulx = L->pool[id].ulx, uly = L->pool[id].uly, urx = L->pool[id].urx;
    int brx = L->pool[id].brx, bry = L->pool[id].bry, bly = L->pool[id].bly;
    int nk = L->pool[id].nkeys;
    int halfX = (int)ceilf((float)(urx - ulx) / 2);
    int halfY = (int)ceilf((float)(bry - uly) / 2);
    for (int q = 0; q < 4; ++q) { kid[q] = ol_new(L); L->pool[kid[q]].keys = (int*)malloc(sizeof(int) * (nk ? nk : 1)); }
    onode *p = &L->pool[id], *n1 = &L->pool[kid[0]], *n2 = &L->pool[kid[1]], *n3 = &L->pool[kid[2]], *n4 = &L->pool[kid[3]];
    n1->ulx = ulx; n1->uly = uly; n1->urx = ulx + halfX; n1->brx = ulx + halfX; n1->bry = uly + halfY; n1->bly = uly + halfY;
    n2->ulx = n1->urx; n2->uly = uly; n2->urx = urx; n2->brx = urx; n2->bry = uly + halfY; n2->bly = n1->bry;
    n3->ulx = ulx; n3->uly = n1->bly; n3->urx = n1->brx; n3->brx = n1->brx; n3->bry = bly; n3->bly = bly;
    n4->ulx = n3->urx; n4->uly = n2->bry; n4->urx = n2->brx; n4->brx = brx; n4->bry = bry; n4->bly = n3->bry;
    for (int i = 0; i < nk; ++i) {
        int k = p->keys[i];
        onode* dst;
        if (px[k] < (float)n1->urx) dst = (py[k] < (float)n1->bry) ? n1 : n3;
        else dst = (py[k] < (float)n1->bry) ? n2 : n4;
        dst->keys[dst->nkeys++] = k;
    }
    for (int q = 0; q < 4; ++q) if (L->pool[kid[q]].nkeys == 1) L->pool[kid[q]].nomore = 1;
}

typedef struct { int nkeys, seq, id; } sz_ptr;
static int cmp_sz_ptr(const void* a, const void* b) {
    const sz_ptr *x = (const sz_ptr*)a, *y = (const sz_ptr*)b;
    if (x->nkeys != y->nkeys) return x->nkeys < y->nkeys ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}

int orc_distribute_octree(const float* px, const float* py, const float* resp, const int* order, int n,
                          int minX, int maxX, int minY, int maxY, int N, int* keep, int cap) {
    (void)order;
    const int nIni = (int)round((double)((float)(maxX - minX) / (float)(maxY - minY)));
    if (nIni < 1) return -1;
    const float hX = (float)(maxX - minX) / (float)nIni;
    olist L; L.cap = 64; L.pool = (onode*)malloc(sizeof(onode) * L.cap);
    L.used = 0; L.head = L.tail = -1; L.size = 0; L.seq = 0;
    int* ini = (int*)malloc(sizeof(int) * nIni);
    for (int i = 0; i < nIni; ++i) {
        int id = ol_new(&L);
        onode* ni = &L.pool[id];
        ni->ulx = (int)(hX * (float)i); ni->uly = 0;
        ni->urx = (int)(hX * (float)(i + 1));
        ni->bly = maxY - minY; ni->brx = ni->urx; ni->bry = maxY - minY;
        ni->keys = (int*)malloc(sizeof(int) * (n ? n : 1));
        ol_push_back(&L, id);
        ini[i] = id;
    }
    for (int i = 0; i < n; ++i) {
        int b = (int)(px[i] / hX);
        onode* nd = &L.pool[ini[b]];
        nd->keys[nd->nkeys++] = i;
    }
    free(ini);
    for (int it = L.head; it >= 0;) {
        onode* nd = &L.pool[it];
        if (nd->nkeys == 1) { nd->nomore = 1; it = nd->next; }
        else if (nd->nkeys == 0) it = ol_erase(&L, it);
        else it = nd->next;
    }
    int finish = 0;
    sz_ptr* vsz = (sz_ptr*)malloc(sizeof(sz_ptr) * (size_t)(4 * (n + 8)));
    sz_ptr* vprev = (sz_ptr*)malloc(sizeof(sz_ptr) * (size_t)(4 * (n + 8)));
    int nsz = 0;
    while (!finish) {
        int prevSize = L.size, nToExpand = 0;
        nsz = 0;
        for (int it = L.head; it >= 0;) {
            if (L.pool[it].nomore) { it = L.pool[it].next; continue; }
            int kid[4];
            divide_node(&L, it, px, py, kid);
            for (int q = 0; q < 4; ++q) {
                if (L.pool[kid[q]].nkeys > 0) {
                    ol_push_front(&L, kid[q]);
                    if (L.pool[kid[q]].nkeys > 1) {
                        ++nToExpand;
                        vsz[nsz].nkeys = L.pool[kid[q]].nkeys; vsz[nsz].seq = L.pool[kid[q]].seq; vsz[nsz].id = kid[q]; ++nsz;
                    }
                } else { free(L.pool[kid[q]].keys); L.pool[kid[q]].keys = NULL; }
            }
            it = ol_erase(&L, it);
        }
        if (L.size >= N || L.size == prevSize) finish = 1;
        else if (L.size + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = L.size;
                int np = nsz;
                memcpy(vprev, vsz, sizeof(sz_ptr) * np);
                nsz = 0;
                qsort(vprev, np, sizeof(sz_ptr), cmp_sz_ptr);
                for (int j = np - 1; j >= 0; --j) {
                    int kid[4];
                    divide_node(&L, vprev[j].id, px, py, kid);
                    for (int q = 0; q < 4; ++q) {
                        if (L.pool[kid[q]].nkeys > 0) {
                            ol_push_front(&L, kid[q]);
                            if (L.pool[kid[q]].nkeys > 1) {
                                vsz[nsz].nkeys = L.pool[kid[q]].nkeys; vsz[nsz].seq = L.pool[kid[q]].seq; vsz[nsz].id = kid[q]; ++nsz;
                            }
                        } else { free(L.pool[kid[q]].keys); L.pool[kid[q]].keys = NULL; }
                    }
                    ol_erase(&L, vprev[j].id);
                    if (L.size >= N) break;
                }
                if (L.size >= N || L.size == prevSize) finish = 1;
            }
        }
    }
    int m = 0;
    for (int it = L.head; it >= 0; it = L.pool[it].next) {
        onode* nd = &L.pool[it];
        int bk = nd->keys[0];
        float mr = resp[bk];
        for (int k = 1; k < nd->nkeys; ++k)
            if (resp[nd->keys[k]] > mr) { bk = nd->keys[k]; mr = resp[bk]; }
        if (m < cap) keep[m] = bk;
        ++m;
    }
    for (int i = 0; i < L.used; ++i) free(L.pool[i].keys);
    free(L.pool); free(vsz); free(vprev);
    return m;
}

/* ---------------------------------------------------------------- full orb32 operator() ---------- */
int orc_orb32_extract(const uint8_t* gray, int w, int h, int stride,
                      int nfeatures, int nlevels, float scale_factor, float detect_th,
                      orc_keypoint* kps, uint8_t* desc, float* kpsize, int cap, int* n_out,
                      int* n_candidates) {
    if (nlevels < 1 || nlevels > ORC_MAX_LEVELS) return -1;
    const float orb_sf = 1.2f;     /* cv::ORB::create() default; the reference never calls setScaleFactor */
    int lw[ORC_MAX_LEVELS], lh[ORC_MAX_LEVELS], q_orb[ORC_MAX_LEVELS], q_ext[ORC_MAX_LEVELS];
    long offs[ORC_MAX_LEVELS];
    float lscale[ORC_MAX_LEVELS];
    long total = orc_orb_pyramid(gray, w, h, stride, nlevels, orb_sf, NULL, offs, lw, lh, lscale);
    uint8_t* pyr = (uint8_t*)malloc((size_t)total);
    uint8_t* blr = (uint8_t*)malloc((size_t)total);
    orc_orb_pyramid(gray, w, h, stride, nlevels, orb_sf, pyr, offs, lw, lh, lscale);
    orc_features_per_level(nfeatures * 10, nlevels, orb_sf, q_orb);       /* src/Feature_orb32.cpp:22 */
    orc_features_per_level(nfeatures, nlevels, scale_factor, q_ext);      /* src/FeatureExtractor.cpp:97-108 */
    const int fast_th = (int)detect_th;                                   /* src/Feature_orb32.cpp:30 */

    /* settings: maxKeyPtSize0 = maxKeyPtSize = pow(1.2f, 7.f), minKeyPtSize = 1 (src/FeatureExtractor.cpp:52-55) */
    const float maxSize0 = powf(1.2f, (float)(8 - 1.0));
    const float maxSize = maxSize0, minSize = 1.0f;

    int n = 0, ncand = 0, rc = 0;
    for (int l = 0; l < nlevels && rc == 0; ++l) {
        const uint8_t* img = pyr + offs[l];
        int W = lw[l], H = lh[l];
        int ccap = W * H / 4 + 16;
        int* xs = (int*)malloc(sizeof(int) * ccap); int* ys = (int*)malloc(sizeof(int) * ccap);
        int* fs = (int*)malloc(sizeof(int) * ccap); float* hr = (float*)malloc(sizeof(float) * ccap);
        int m = orc_orb_detect_level(img, W, H, W, fast_th, q_orb[l], xs, ys, hr, fs, ccap);
        if (m < 0) m = 0;
        ncand += m;
        float* fx = (float*)malloc(sizeof(float) * (m + 1)); float* fy = (float*)malloc(sizeof(float) * (m + 1));
        for (int i = 0; i < m; ++i) { fx[i] = (float)xs[i] * lscale[l]; fy[i] = (float)ys[i] * lscale[l]; }
        int* keep = (int*)malloc(sizeof(int) * (m + 1));
        int nk = m ? orc_distribute_octree(fx, fy, hr, NULL, m, 0, w, 0, h, q_ext[l], keep, m) : 0;
        if (nk > 0) orc_blur7_level(img, W, H, W, blr + offs[l], W);
        for (int j = 0; j < nk; ++j) {
            int i = keep[j];
            if (n >= cap) { rc = -2; break; }
            orc_keypoint* kp = &kps[n];
            kp->x = fx[i]; kp->y = fy[i];
            kp->size = 31 * lscale[l];
            kp->angle = orc_ic_angle(img, W, H, W, xs[i], ys[i]);
            kp->response = hr[i]; kp->octave = l; kp->class_id = -1;
            float inv = 1.f / lscale[l];
            int cx = cv_round_f(kp->x * inv), cy = cv_round_f(kp->y * inv);
            orc_rbrief32(img, blr + offs[l], W, H, W, cx, cy, kp->angle, desc + (long)n * 32);
            if (kpsize) {
                float s = powf(scale_factor, (float)l);              /* src/Feature_orb32.cpp:59-61 */
                float sn = maxSize;
                if (maxSize > minSize) sn = 1.0f + (s - minSize) * (maxSize0 - 1.0f) / (maxSize - minSize);
                kpsize[n] = sn;
            }
            ++n;
        }
        free(xs); free(ys); free(fs); free(hr); free(fx); free(fy); free(keep);
    }
    free(pyr); free(blr);
    if (n_out) *n_out = n;
    if (n_candidates) *n_candidates = ncand;
    return rc;
}

/* ---------------------------------------------------------------- Image::GetGrayImage ------------------------- */
/* reference src/Image.cpp:30-53 -> cv::cvtColor(CV_RGB2GRAY / CV_BGR2GRAY / ...A2GRAY), 8-bit: OpenCV's 15-bit fixed
 * point (R,G,B = 9798,19235,3735; +2^14; >>15), pinned to cv2 4.13.0 in tests/test_oracle_golden.py. */
void orc_gray_from_color(const uint8_t* src, int channels, int rgb, int w, int h, int sstride, uint8_t* dst, int dstride) {
    const int c0 = rgb ? 9798 : 3735, c1 = 19235, c2 = rgb ? 3735 : 9798;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint8_t* p = src + (long)y * sstride + (long)x * channels;
            dst[(long)y * dstride + x] = (uint8_t)((p[0] * c0 + p[1] * c1 + p[2] * c2 + (1 << 14)) >> 15);
        }
}

/*
 * afv_oracle_brisk.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) of the brisk48 extraction path.
 *
 * PARITY UNPINNED vs ETH brisk v2.  The reference's brisk48 arithmetic lives in ETH "brisk" (fontan::brisk, un-pinned; upstream
 * gwli/brisk = ethz-asl BRISK v2; reference call sites src/Feature_brisk48.cpp:24-26 (BriskFeatureDetector(int(detectTh),
 * nOctaves/2, suppressScaleNonmaxima = true)), :42-45 (BriskDescriptorExtractor(rotationInvariant = true, scaleInvariant = true,
 * Version::briskV2) -> 48-byte descriptors)), which is NOT vendored under the reference tree.  This file restates the PUBLISHED
 * algorithm (Leutenegger, Chli, Siegwart: "BRISK: Binary Robust Invariant Scalable Keypoints", ICCV 2011): AGAST/FAST 9-16 score
 * pyramid of `octaves` octaves + intra-octaves, 3x3 non-maxima suppression per layer, maxima search in the layers above / below,
 * 2-D quadratic sub-pixel fit per layer and 1-D parabola over scale, then the 60-point sampling pattern (radii 0.85 * {0, 2.9, 4.9,
 * 7.4, 10.8}, 1 + 10 + 14 + 15 + 20 points), box-smoothed intensities from an integral image, long-pair gradient orientation
 * (d > 8.2) and short-pair brightness comparisons.  The structure followed is the authors' own reference implementation as it
 * ships in OpenCV (cv::BRISK, features2d/src/brisk.cpp, contributed by the BRISK authors), which is SEMI-PINNED: with the 64-byte
 * short-pair table of the paper (d < 5.85, 512 pairs) this oracle reproduces cv2 4.13.0's BRISK_create(34, 4) keypoints,
 * orientations and 512-bit descriptors (tests/test_oracle_brisk.py, fixture from tools/make_golden_brisk.py).  What cannot be
 * pinned: ETH brisk v2's 48-byte pair table (an embedded pattern blob in that library).  The 48-byte layout here keeps the same 60
 * sample points and takes the 384 SHORTEST of the paper's 512 short pairs (ties by enumeration order), packed in enumeration
 * order -- documented in DESIGN.md as the one place where the product's brisk48 bits are a stand-in.
 *
 * Around it, everything the reference does: octave := keypoint.octave = layer index 0..7 (src/Feature_brisk48.cpp:29-30, :50-52),
 * DistributeOctTree per layer with quota mnFeaturesPerLevel (:58-60 -> src/FeatureExtractor.cpp:276-284), all layers merged into
 * level 0 before compute (:38-41), border keypoints removed by the descriptor extractor, computeSize with powf(scaleFactor0,
 * octave) (:54-56).
 *
 * Score cache: the authors' implementation computes AGAST scores lazily and caches them in a per-layer byte image that the 3x3
 * maximum test reads RAW, so the result of an exact-tie decision can depend on which sub-threshold scores earlier keypoints
 * happened to touch.  mode = ORC_BRISK_SEQUENTIAL follows that literally (this is the cv2-pinned mode); mode = ORC_BRISK_DENSE is
 * the order-independent contract the CUDA path implements: every score is the true score (as if all had been touched with
 * threshold 1), which changes only exact-tie decisions between equal neighbouring maxima (measured in tests/test_oracle_brisk.py).
 *
 * Arithmetic contract: integer where the original is integer; IEEE float32 without contraction, operation order as written;
 * libm only at table-construction time (pattern cos/sin/pow, log for the size -> scale-index thresholds) and for the orientation
 * angle, which is atan2 evaluated in DOUBLE by the shared polynomial below and rounded to float (libm atan2f in cv2-pinned mode).
 */
#include "afv_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define BRK_MAX_LAYERS 16
#define BRK_POINTS 60
#define BRK_SCALES 64
#define BRK_NROT 1024
#define BRK_BASIC_SIZE 12.0f

/* ---------------------------------------------------------------------------------------------------------------------
 * cv::resize(..., INTER_AREA) for 8UC1 as BriskLayer::halfsample / twothirdsample call it: exact 2x2 mean when both scale factors
 * are exactly 2, else the general area path (float weights, sequential float accumulation, round-half-even saturate).
 * Pinned to cv2 4.13.0 (tests/test_oracle_brisk.py).
 * ------------------------------------------------------------------------------------------------------------------- */
typedef struct { int di, si; float alpha; } brk_tab;

static int brk_area_tab(int ssize, int dsize, brk_tab* tab) {
    const double scale = 1.0 / ((double)dsize / (double)ssize);
    int k = 0;
    for (int dx = 0; dx < dsize; ++dx) {
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        const double cell = scale < ssize - fsx1 ? scale : ssize - fsx1;
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        if (sx2 > ssize - 1) sx2 = ssize - 1;
        if (sx1 > sx2) sx1 = sx2;
        if (sx1 - fsx1 > 1e-3) { tab[k].di = dx; tab[k].si = sx1 - 1; tab[k++].alpha = (float)((sx1 - fsx1) / cell); }
        for (int sx = sx1; sx < sx2; ++sx) { tab[k].di = dx; tab[k].si = sx; tab[k++].alpha = (float)(1.0 / cell); }
        if (fsx2 - sx2 > 1e-3) {
            double a = fsx2 - sx2; if (a > 1.0) a = 1.0; if (a > cell) a = cell;
            tab[k].di = dx; tab[k].si = sx2; tab[k++].alpha = (float)(a / cell);
        }
    }
    return k;
}

void orc_resize_area_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
    if (sw == 2 * dw && sh == 2 * dh) {
        for (int y = 0; y < dh; ++y)
            for (int x = 0; x < dw; ++x) {
                const uint8_t* s = src + (size_t)(2 * y) * sstride + 2 * x;
                dst[(size_t)y * dstride + x] = (uint8_t)((s[0] + s[1] + s[sstride] + s[sstride + 1] + 2) >> 2);
            }
        return;
    }
    brk_tab* xt = (brk_tab*)malloc(sizeof(brk_tab) * (size_t)(2 * sw + 2));
    brk_tab* yt = (brk_tab*)malloc(sizeof(brk_tab) * (size_t)(2 * sh + 2));
    const int nx = brk_area_tab(sw, dw, xt), ny = brk_area_tab(sh, dh, yt);
    float* buf = (float*)malloc(sizeof(float) * (size_t)dw);
    float* sum = (float*)calloc((size_t)dw, sizeof(float));
    int prev = ny ? yt[0].di : 0;
    for (int j = 0; j < ny; ++j) {
        const float beta = yt[j].alpha; const int dy = yt[j].di;
        const uint8_t* S = src + (size_t)yt[j].si * sstride;
        for (int x = 0; x < dw; ++x) buf[x] = 0.f;
        for (int k = 0; k < nx; ++k) buf[xt[k].di] += (float)S[xt[k].si] * xt[k].alpha;
        if (dy != prev) {
            for (int x = 0; x < dw; ++x) { dst[(size_t)prev * dstride + x] = (uint8_t)lrintf(sum[x] < 0.f ? 0.f : (sum[x] > 255.f ? 255.f : sum[x])); sum[x] = beta * buf[x]; }
            prev = dy;
        } else {
            for (int x = 0; x < dw; ++x) sum[x] += beta * buf[x];
        }
    }
    for (int x = 0; x < dw; ++x) dst[(size_t)prev * dstride + x] = (uint8_t)lrintf(sum[x] < 0.f ? 0.f : (sum[x] > 255.f ? 255.f : sum[x]));
    free(xt); free(yt); free(buf); free(sum);
}

/* ---------------------------------------------------------------------------------------------------------------------
 * AGAST corner scores.  agast_cornerScore<OAST_9_16>(p, T-1) bisects for the largest b in [T-1, 254] at which the pixel still is a
 * 9-of-16 segment corner (strict >), i.e. max(best - 1, T - 1) with best = max over the 16 arcs of 9 of the smallest |difference|
 * of one sign; BriskLayer::getAgastScore then zeroes anything < T.  AGAST_5_8 is the same on the 8-neighbour ring with arcs of 5.
 * ------------------------------------------------------------------------------------------------------------------- */
static const int brk_c16x[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int brk_c16y[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
static const int brk_c8x[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
static const int brk_c8y[8] = {0, 1, 1, 1, 0, -1, -1, -1};

static int brk_arc_best(const int* d, int n, int arc) {      /* max over arcs of min(d) and of min(-d) */
    int best = -256;
    for (int k = 0; k < n; ++k) {
        int mn = 256, mx = -256;
        for (int j = 0; j < arc; ++j) { const int v = d[(k + j) % n]; if (v < mn) mn = v; if (v > mx) mx = v; }
        if (mn > best) best = mn;
        if (-mx > best) best = -mx;
    }
    return best;
}
/* true score: largest b >= 0 with a corner at threshold b (strict), -1 if none */
static int brk_score_9_16(const uint8_t* img, int stride, int x, int y) {
    const uint8_t* p = img + (size_t)y * stride + x;
    int d[16];
    for (int k = 0; k < 16; ++k) d[k] = (int)p[brk_c16y[k] * stride + brk_c16x[k]] - (int)p[0];
    return brk_arc_best(d, 16, 9) - 1;
}
static int brk_score_5_8(const uint8_t* img, int stride, int x, int y) {
    const uint8_t* p = img + (size_t)y * stride + x;
    int d[8];
    for (int k = 0; k < 8; ++k) d[k] = (int)p[brk_c8y[k] * stride + brk_c8x[k]] - (int)p[0];
    return brk_arc_best(d, 8, 5) - 1;
}

typedef struct {
    int w, h;
    uint8_t* img;            /* tight, stride w */
    uint8_t* scores;         /* BriskLayer::scores_ */
    float scale, offset;
    int dense;               /* ORC_BRISK_DENSE: scores holds every true score >= 1 up front */
} brk_layer;

/* BriskLayer::getAgastScore(int x, int y, int threshold) incl. its cache rule (a cached value > 2 is returned as is) */
static int brk_get_score(brk_layer* L, int x, int y, int threshold) {
    if (x < 3 || y < 3) return 0;
    if (x >= L->w - 3 || y >= L->h - 3) return 0;
    uint8_t* sc = L->scores + (size_t)y * L->w + x;
    if (L->dense) return *sc >= threshold ? *sc : 0;
    if (*sc > 2) return *sc;
    int s = brk_score_9_16(L->img, L->w, x, y);
    if (s < threshold - 1) s = threshold - 1;
    if (s > 255) s = 255;
    if (s < threshold) s = 0;
    *sc = (uint8_t)s;
    return s;
}
static int brk_get_score_5_8(const brk_layer* L, int x, int y, int threshold) {
    if (x < 2 || y < 2) return 0;
    if (x >= L->w - 2 || y >= L->h - 2) return 0;
    int s = brk_score_5_8(L->img, L->w, x, y);
    if (s < threshold - 1) s = threshold - 1;
    if (s < threshold) s = 0;
    return s;
}
/* BriskLayer::getAgastScore(float xf, float yf, int threshold, scale_in = 1): bilinear inside the layer, truncated to uchar */
static int brk_get_score_f(brk_layer* L, float xf, float yf, int threshold) {
    const int x = (int)xf; const float rx1 = xf - (float)x, rx = 1.0f - rx1;
    const int y = (int)yf; const float ry1 = yf - (float)y, ry = 1.0f - ry1;
    const float v = rx * ry * (float)brk_get_score(L, x, y, threshold) + rx1 * ry * (float)brk_get_score(L, x + 1, y, threshold) +
                    rx * ry1 * (float)brk_get_score(L, x, y + 1, threshold) + rx1 * ry1 * (float)brk_get_score(L, x + 1, y + 1, threshold);
    return (int)(uint8_t)(int)v;
}

/* BriskScaleSpace::isMax2D: >= all 8 neighbours on the RAW score image; equal neighbours are compared by their 3x3 binomial sums */
static int brk_is_max2d(const brk_layer* L, int x, int y) {
    const int W = L->w;
    const uint8_t* d = L->scores + (size_t)y * W + x;
    const int c = d[0];
    const int s_10 = d[-1], s10 = d[1], s0_1 = d[-W], s01 = d[W], s_1_1 = d[-W - 1], s1_1 = d[-W + 1], s_11 = d[W - 1], s11 = d[W + 1];
    if (c < s_10 || c < s10 || c < s0_1 || c < s01 || c < s_1_1 || c < s1_1 || c < s_11 || c < s11) return 0;
    int delta[16], nd = 0;
    if (c == s_1_1) { delta[nd++] = -1; delta[nd++] = -1; }
    if (c == s0_1) { delta[nd++] = 0; delta[nd++] = -1; }
    if (c == s1_1) { delta[nd++] = 1; delta[nd++] = -1; }
    if (c == s_10) { delta[nd++] = -1; delta[nd++] = 0; }
    if (c == s10) { delta[nd++] = 1; delta[nd++] = 0; }
    if (c == s_11) { delta[nd++] = -1; delta[nd++] = 1; }
    if (c == s01) { delta[nd++] = 0; delta[nd++] = 1; }
    if (c == s11) { delta[nd++] = 1; delta[nd++] = 1; }
    if (nd) {
        const int smoothed = 4 * c + 2 * (s_10 + s10 + s0_1 + s01) + s_1_1 + s1_1 + s_11 + s11;
        for (int i = 0; i < nd; i += 2) {
            const uint8_t* q = L->scores + (size_t)(y - 1 + delta[i + 1]) * W + x + delta[i] - 1;
            int o = q[0] + 2 * q[1] + q[2];
            q += W; o += 2 * q[0] + 4 * q[1] + 2 * q[2];
            q += W; o += q[0] + 2 * q[1] + q[2];
            if (o > smoothed) return 0;
        }
    }
    return 1;
}

/* BriskScaleSpace::subpixel2D: least-squares quadratic through a 3x3 score patch */
static float brk_subpixel2d(int s_0_0, int s_0_1, int s_0_2, int s_1_0, int s_1_1, int s_1_2, int s_2_0, int s_2_1, int s_2_2,
                            float* delta_x, float* delta_y) {
    const int tmp1 = s_0_0 + s_0_2 - 2 * s_1_1 + s_2_0 + s_2_2;
    const int coeff1 = 3 * (tmp1 + s_0_1 - ((s_1_0 + s_1_2) << 1) + s_2_1);
    const int coeff2 = 3 * (tmp1 - ((s_0_1 + s_2_1) << 1) + s_1_0 + s_1_2);
    const int tmp2 = s_0_2 - s_2_0;
    const int tmp3 = (s_0_0 + tmp2 - s_2_2);
    const int tmp4 = tmp3 - 2 * tmp2;
    const int coeff3 = -3 * (tmp3 + s_0_1 - s_2_1);
    const int coeff4 = -3 * (tmp4 + s_1_0 - s_1_2);
    const int coeff5 = (s_0_0 - s_0_2 - s_2_0 + s_2_2) << 2;
    const int coeff6 = -(s_0_0 + s_0_2 - ((s_1_0 + s_0_1 + s_1_2 + s_2_1) << 1) - 5 * s_1_1 + s_2_0 + s_2_2) << 1;
    const int H_det = 4 * coeff1 * coeff2 - coeff5 * coeff5;
    if (H_det == 0) { *delta_x = 0.0f; *delta_y = 0.0f; return (float)coeff6 / 18.0f; }
    if (!(H_det > 0 && coeff1 < 0)) {
        int tmp_max = coeff3 + coeff4 + coeff5;
        *delta_x = 1.0f; *delta_y = 1.0f;
        int tmp = -coeff3 + coeff4 - coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; *delta_x = -1.0f; *delta_y = 1.0f; }
        tmp = coeff3 - coeff4 - coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; *delta_x = 1.0f; *delta_y = -1.0f; }
        tmp = -coeff3 - coeff4 + coeff5;
        if (tmp > tmp_max) { tmp_max = tmp; *delta_x = -1.0f; *delta_y = -1.0f; }
        return (float)(tmp_max + coeff1 + coeff2 + coeff6) / 18.0f;
    }
    *delta_x = (float)(2 * coeff2 * coeff3 - coeff4 * coeff5) / (float)(-H_det);
    *delta_y = (float)(2 * coeff1 * coeff4 - coeff3 * coeff5) / (float)(-H_det);
    int tx = 0, tx_ = 0, ty = 0, ty_ = 0;
    if (*delta_x > 1.0f) tx = 1; else if (*delta_x < -1.0f) tx_ = 1;
    if (*delta_y > 1.0f) ty = 1;
    if (*delta_y < -1.0f) ty_ = 1;
    const float c1 = (float)coeff1, c2 = (float)coeff2, c3 = (float)coeff3, c4 = (float)coeff4, c5 = (float)coeff5, c6 = (float)coeff6;
    if (tx || tx_ || ty || ty_) {
        float dx1 = 0.0f, dx2 = 0.0f, dy1 = 0.0f, dy2 = 0.0f;
        if (tx) {
            dx1 = 1.0f; dy1 = -(float)(coeff4 + coeff5) / (float)(2 * coeff2);
            if (dy1 > 1.0f) dy1 = 1.0f; else if (dy1 < -1.0f) dy1 = -1.0f;
        } else if (tx_) {
            dx1 = -1.0f; dy1 = -(float)(coeff4 - coeff5) / (float)(2 * coeff2);
            if (dy1 > 1.0f) dy1 = 1.0f; else if (dy1 < -1.0f) dy1 = -1.0f;
        }
        if (ty) {
            dy2 = 1.0f; dx2 = -(float)(coeff3 + coeff5) / (float)(2 * coeff1);
            if (dx2 > 1.0f) dx2 = 1.0f; else if (dx2 < -1.0f) dx2 = -1.0f;
        } else if (ty_) {
            dy2 = -1.0f; dx2 = -(float)(coeff3 - coeff5) / (float)(2 * coeff1);
            if (dx2 > 1.0f) dx2 = 1.0f; else if (dx2 < -1.0f) dx2 = -1.0f;
        }
        const float max1 = (c1 * dx1 * dx1 + c2 * dy1 * dy1 + c3 * dx1 + c4 * dy1 + c5 * dx1 * dy1 + c6) / 18.0f;
        const float max2 = (c1 * dx2 * dx2 + c2 * dy2 * dy2 + c3 * dx2 + c4 * dy2 + c5 * dx2 * dy2 + c6) / 18.0f;
        if (max1 > max2) { *delta_x = dx1; *delta_y = dy1; return max1; }
        *delta_x = dx2; *delta_y = dy2; return max2;
    }
    const float dx = *delta_x, dy = *delta_y;
    return (c1 * dx * dx + c2 * dy * dy + c3 * dx + c4 * dy + c5 * dx * dy + c6) / 18.0f;
}

/* refine1D (octave, layer > 0), refine1D_1 (intra-octave), refine1D_2 (layer 0): parabola through the three scale samples */
static float brk_refine1d(int kind, float s_05, float s0, float s05, float* max) {
    const int i_05 = (int)(1024.0 * s_05 + 0.5), i0 = (int)(1024.0 * s0 + 0.5), i05 = (int)(1024.0 * s05 + 0.5);
    int a, b, c; float lo, hi, div;
    if (kind == 0) { a = 16 * i_05 - 24 * i0 + 8 * i05; b = -40 * i_05 + 54 * i0 - 14 * i05; c = 24 * i_05 - 27 * i0 + 6 * i05; lo = 0.75f; hi = 1.5f; div = 3072.0f; }
    else if (kind == 1) { a = 9 * i_05 - 18 * i0 + 9 * i05; b = -21 * i_05 + 36 * i0 - 15 * i05; c = 12 * i_05 - 16 * i0 + 6 * i05; lo = 0.6666666666666666666666666667f; hi = 1.3333333333333333333333333333f; div = 2048.0f; }
    else { a = 2 * i_05 - 4 * i0 + 2 * i05; b = -5 * i_05 + 8 * i0 - 3 * i05; c = 3 * i_05 - 3 * i0 + 1 * i05; lo = 0.7f; hi = 1.5f; div = 1024.0f; }
    if (a >= 0) {
        if (s0 >= s_05 && s0 >= s05) { *max = s0; return 1.0f; }
        if (s_05 >= s0 && s_05 >= s05) { *max = s_05; return lo; }
        if (s05 >= s0 && s05 >= s_05) { *max = s05; return hi; }
    }
    float r = -(float)b / (float)(2 * a);
    if (r < lo) r = lo; else if (r > hi) r = hi;
    *max = (float)c + (float)a * r * r + (float)b * r;
    *max /= div;
    return r;
}

typedef struct { brk_layer L[BRK_MAX_LAYERS]; int layers; } brk_space;

static void brk_patch9(brk_layer* L, int x, int y, int* s) {      /* s_0_0 s_1_0 s_2_0 s_2_1 s_1_1 s_0_1 s_0_2 s_1_2 s_2_2 evaluation order */
    s[0] = brk_get_score(L, x - 1, y - 1, 1); s[3] = brk_get_score(L, x, y - 1, 1); s[6] = brk_get_score(L, x + 1, y - 1, 1);
    s[7] = brk_get_score(L, x + 1, y, 1); s[4] = brk_get_score(L, x, y, 1); s[1] = brk_get_score(L, x - 1, y, 1);
    s[2] = brk_get_score(L, x - 1, y + 1, 1); s[5] = brk_get_score(L, x, y + 1, 1); s[8] = brk_get_score(L, x + 1, y + 1, 1);
}   /* s[3*i + j] = s_i_j (i = column offset, j = row offset) */

/* BriskScaleSpace::getScoreMaxAbove */
static float brk_max_above(brk_space* S, int layer, int x_layer, int y_layer, int threshold, int* ismax, float* dx, float* dy) {
    *ismax = 0;
    float x_1, x1, y_1, y1;
    brk_layer* A = &S->L[layer + 1];
    if (layer % 2 == 0) {
        x_1 = (float)(4 * x_layer - 1 - 2) / 6.0f; x1 = (float)(4 * x_layer - 1 + 2) / 6.0f;
        y_1 = (float)(4 * y_layer - 1 - 2) / 6.0f; y1 = (float)(4 * y_layer - 1 + 2) / 6.0f;
    } else {
        x_1 = (float)(6 * x_layer - 1 - 3) / 8.0f; x1 = (float)(6 * x_layer - 1 + 3) / 8.0f;
        y_1 = (float)(6 * y_layer - 1 - 3) / 8.0f; y1 = (float)(6 * y_layer - 1 + 3) / 8.0f;
    }
    int max_x = (int)x_1 + 1, max_y = (int)y_1 + 1;
    float tmp_max;
    float maxval = (float)brk_get_score_f(A, x_1, y_1, 1);
    if (maxval > threshold) return 0;
    for (int x = (int)x_1 + 1; x <= (int)x1; ++x) {
        tmp_max = (float)brk_get_score_f(A, (float)x, y_1, 1);
        if (tmp_max > threshold) return 0;
        if (tmp_max > maxval) { maxval = tmp_max; max_x = x; }
    }
    tmp_max = (float)brk_get_score_f(A, x1, y_1, 1);
    if (tmp_max > threshold) return 0;
    if (tmp_max > maxval) { maxval = tmp_max; max_x = (int)x1; }
    for (int y = (int)y_1 + 1; y <= (int)y1; ++y) {
        tmp_max = (float)brk_get_score_f(A, x_1, (float)y, 1);
        if (tmp_max > threshold) return 0;
        if (tmp_max > maxval) { maxval = tmp_max; max_x = (int)(x_1 + 1); max_y = y; }
        for (int x = (int)x_1 + 1; x <= (int)x1; ++x) {
            tmp_max = (float)brk_get_score(A, x, y, 1);
            if (tmp_max > threshold) return 0;
            if (tmp_max > maxval) { maxval = tmp_max; max_x = x; max_y = y; }
        }
        tmp_max = (float)brk_get_score_f(A, x1, (float)y, 1);
        if (tmp_max > threshold) return 0;
        if (tmp_max > maxval) { maxval = tmp_max; max_x = (int)x1; max_y = y; }
    }
    tmp_max = (float)brk_get_score_f(A, x_1, y1, 1);
    if (tmp_max > maxval) { maxval = tmp_max; max_x = (int)(x_1 + 1); max_y = (int)y1; }
    for (int x = (int)x_1 + 1; x <= (int)x1; ++x) {
        tmp_max = (float)brk_get_score_f(A, (float)x, y1, 1);
        if (tmp_max > maxval) { maxval = tmp_max; max_x = x; max_y = (int)y1; }
    }
    tmp_max = (float)brk_get_score_f(A, x1, y1, 1);
    if (tmp_max > maxval) { maxval = tmp_max; max_x = (int)x1; max_y = (int)y1; }

    int s[9];
    s[0] = brk_get_score(A, max_x - 1, max_y - 1, 1); s[3] = brk_get_score(A, max_x, max_y - 1, 1); s[6] = brk_get_score(A, max_x + 1, max_y - 1, 1);
    s[7] = brk_get_score(A, max_x + 1, max_y, 1); s[4] = brk_get_score(A, max_x, max_y, 1); s[1] = brk_get_score(A, max_x - 1, max_y, 1);
    s[2] = brk_get_score(A, max_x - 1, max_y + 1, 1); s[5] = brk_get_score(A, max_x, max_y + 1, 1); s[8] = brk_get_score(A, max_x + 1, max_y + 1, 1);
    float dx_1, dy_1;
    const float refined_max = brk_subpixel2d(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], &dx_1, &dy_1);
    const float real_x = (float)max_x + dx_1, real_y = (float)max_y + dy_1;
    int returnrefined = 1;
    if (layer % 2 == 0) {
        *dx = (real_x * 6.0f + 1.0f) / 4.0f - (float)x_layer;
        *dy = (real_y * 6.0f + 1.0f) / 4.0f - (float)y_layer;
    } else {
        *dx = (real_x * 8.0f + 1.0f) / 6.0f - (float)x_layer;
        *dy = (real_y * 8.0f + 1.0f) / 6.0f - (float)y_layer;
    }
    if (*dx > 1.0f) { *dx = 1.0f; returnrefined = 0; }
    if (*dx < -1.0f) { *dx = -1.0f; returnrefined = 0; }
    if (*dy > 1.0f) { *dy = 1.0f; returnrefined = 0; }
    if (*dy < -1.0f) { *dy = -1.0f; returnrefined = 0; }
    *ismax = 1;
    if (returnrefined) return refined_max > maxval ? refined_max : maxval;
    return maxval;
}

static int brk_ring_sum(brk_layer* B, int x, int y) {
    return 2 * (brk_get_score(B, x - 1, y, 1) + brk_get_score(B, x + 1, y, 1) + brk_get_score(B, x, y + 1, 1) + brk_get_score(B, x, y - 1, 1)) +
           (brk_get_score(B, x + 1, y + 1, 1) + brk_get_score(B, x - 1, y + 1, 1) + brk_get_score(B, x + 1, y - 1, 1) + brk_get_score(B, x - 1, y - 1, 1));
}

/* BriskScaleSpace::getScoreMaxBelow */
static float brk_max_below(brk_space* S, int layer, int x_layer, int y_layer, int threshold, int* ismax, float* dx, float* dy) {
    *ismax = 0;
    float x_1, x1, y_1, y1;
    if (layer % 2 == 0) {
        x_1 = (float)(8 * x_layer + 1 - 4) / 6.0f; x1 = (float)(8 * x_layer + 1 + 4) / 6.0f;
        y_1 = (float)(8 * y_layer + 1 - 4) / 6.0f; y1 = (float)(8 * y_layer + 1 + 4) / 6.0f;
    } else {
        x_1 = (float)(6 * x_layer + 1 - 3) / 4.0f; x1 = (float)(6 * x_layer + 1 + 3) / 4.0f;
        y_1 = (float)(6 * y_layer + 1 - 3) / 4.0f; y1 = (float)(6 * y_layer + 1 + 3) / 4.0f;
    }
    brk_layer* B = &S->L[layer - 1];
    int max_x = (int)x_1 + 1, max_y = (int)y_1 + 1;
    float tmp_max;
    float max = (float)brk_get_score_f(B, x_1, y_1, 1);
    if (max > threshold) return 0;
    for (int x = (int)x_1 + 1; x <= (int)x1; ++x) {
        tmp_max = (float)brk_get_score_f(B, (float)x, y_1, 1);
        if (tmp_max > threshold) return 0;
        if (tmp_max > max) { max = tmp_max; max_x = x; }
    }
    tmp_max = (float)brk_get_score_f(B, x1, y_1, 1);
    if (tmp_max > threshold) return 0;
    if (tmp_max > max) { max = tmp_max; max_x = (int)x1; }
    for (int y = (int)y_1 + 1; y <= (int)y1; ++y) {
        tmp_max = (float)brk_get_score_f(B, x_1, (float)y, 1);
        if (tmp_max > threshold) return 0;
        if (tmp_max > max) { max = tmp_max; max_x = (int)(x_1 + 1); max_y = y; }
        for (int x = (int)x_1 + 1; x <= (int)x1; ++x) {
            tmp_max = (float)brk_get_score(B, x, y, 1);
            if (tmp_max > threshold) return 0;
            if (tmp_max == max) {
                const int t1 = brk_ring_sum(B, x, y);
                const int t2 = brk_ring_sum(B, max_x, max_y);
                if (t1 > t2) { max_x = x; max_y = y; }
            }
            if (tmp_max > max) { max = tmp_max; max_x = x; max_y = y; }
        }
        tmp_max = (float)brk_get_score_f(B, x1, (float)y, 1);
        if (tmp_max > threshold) return 0;
        if (tmp_max > max) { max = tmp_max; max_x = (int)x1; max_y = y; }
    }
    tmp_max = (float)brk_get_score_f(B, x_1, y1, 1);
    if (tmp_max > max) { max = tmp_max; max_x = (int)(x_1 + 1); max_y = (int)y1; }
    for (int x = (int)x_1 + 1; x <= (int)x1; ++x) {
        tmp_max = (float)brk_get_score_f(B, (float)x, y1, 1);
        if (tmp_max > max) { max = tmp_max; max_x = x; max_y = (int)y1; }
    }
    tmp_max = (float)brk_get_score_f(B, x1, y1, 1);
    if (tmp_max > max) { max = tmp_max; max_x = (int)x1; max_y = (int)y1; }

    int s[9];
    s[0] = brk_get_score(B, max_x - 1, max_y - 1, 1); s[3] = brk_get_score(B, max_x, max_y - 1, 1); s[6] = brk_get_score(B, max_x + 1, max_y - 1, 1);
    s[7] = brk_get_score(B, max_x + 1, max_y, 1); s[4] = brk_get_score(B, max_x, max_y, 1); s[1] = brk_get_score(B, max_x - 1, max_y, 1);
    s[2] = brk_get_score(B, max_x - 1, max_y + 1, 1); s[5] = brk_get_score(B, max_x, max_y + 1, 1); s[8] = brk_get_score(B, max_x + 1, max_y + 1, 1);
    float dx_1, dy_1;
    const float refined_max = brk_subpixel2d(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], &dx_1, &dy_1);
    const float real_x = (float)max_x + dx_1, real_y = (float)max_y + dy_1;
    int returnrefined = 1;
    if (layer % 2 == 0) {
        *dx = (float)((real_x * 6.0 + 1.0) / 8.0) - (float)x_layer;
        *dy = (float)((real_y * 6.0 + 1.0) / 8.0) - (float)y_layer;
    } else {
        *dx = (float)((real_x * 4.0 - 1.0) / 6.0) - (float)x_layer;
        *dy = (float)((real_y * 4.0 - 1.0) / 6.0) - (float)y_layer;
    }
    if (*dx > 1.0f) { *dx = 1.0f; returnrefined = 0; }
    if (*dx < -1.0f) { *dx = -1.0f; returnrefined = 0; }
    if (*dy > 1.0f) { *dy = 1.0f; returnrefined = 0; }
    if (*dy < -1.0f) { *dy = -1.0f; returnrefined = 0; }
    *ismax = 1;
    if (returnrefined) return refined_max > max ? refined_max : max;
    return max;
}

/* BriskScaleSpace::refine3D */
static float brk_refine3d(brk_space* S, int layer, int x_layer, int y_layer, float* x, float* y, float* scale, int* ismax) {
    *ismax = 1;
    brk_layer* T = &S->L[layer];
    const int center = brk_get_score(T, x_layer, y_layer, 1);
    float delta_x_above = 0, delta_y_above = 0;
    const float max_above = brk_max_above(S, layer, x_layer, y_layer, center, ismax, &delta_x_above, &delta_y_above);
    if (!*ismax) return 0.0f;
    float max;
    int s[9];
    if (layer % 2 == 0) {
        float delta_x_below = 0, delta_y_below = 0, max_below_float;
        if (layer == 0) {
            int q[9], max_below;
            q[0] = brk_get_score_5_8(T, x_layer - 1, y_layer - 1, 1); max_below = q[0];
            q[3] = brk_get_score_5_8(T, x_layer, y_layer - 1, 1); if (q[3] > max_below) max_below = q[3];
            q[6] = brk_get_score_5_8(T, x_layer + 1, y_layer - 1, 1); if (q[6] > max_below) max_below = q[6];
            q[7] = brk_get_score_5_8(T, x_layer + 1, y_layer, 1); if (q[7] > max_below) max_below = q[7];
            q[4] = brk_get_score_5_8(T, x_layer, y_layer, 1); if (q[4] > max_below) max_below = q[4];
            q[1] = brk_get_score_5_8(T, x_layer - 1, y_layer, 1); if (q[1] > max_below) max_below = q[1];
            q[2] = brk_get_score_5_8(T, x_layer - 1, y_layer + 1, 1); if (q[2] > max_below) max_below = q[2];
            q[5] = brk_get_score_5_8(T, x_layer, y_layer + 1, 1); if (q[5] > max_below) max_below = q[5];
            q[8] = brk_get_score_5_8(T, x_layer + 1, y_layer + 1, 1); if (q[8] > max_below) max_below = q[8];
            (void)brk_subpixel2d(q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], &delta_x_below, &delta_y_below);
            max_below_float = (float)max_below;
        } else {
            max_below_float = brk_max_below(S, layer, x_layer, y_layer, center, ismax, &delta_x_below, &delta_y_below);
            if (!*ismax) return 0;
        }
        brk_patch9(T, x_layer, y_layer, s);
        float delta_x_layer, delta_y_layer;
        const float max_layer = brk_subpixel2d(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], &delta_x_layer, &delta_y_layer);
        const float mid = (float)center > max_layer ? (float)center : max_layer;
        if (layer == 0) *scale = brk_refine1d(2, max_below_float, mid, max_above, &max);
        else *scale = brk_refine1d(0, max_below_float, mid, max_above, &max);
        if (*scale > 1.0f) {
            const float r0 = (1.5f - *scale) / .5f, r1 = 1.0f - r0;
            *x = (r0 * delta_x_layer + r1 * delta_x_above + (float)x_layer) * T->scale + T->offset;
            *y = (r0 * delta_y_layer + r1 * delta_y_above + (float)y_layer) * T->scale + T->offset;
        } else if (layer == 0) {
            const float r0 = (*scale - 0.5f) / 0.5f, r_1 = 1.0f - r0;
            *x = r0 * delta_x_layer + r_1 * delta_x_below + (float)x_layer;
            *y = r0 * delta_y_layer + r_1 * delta_y_below + (float)y_layer;
        } else {
            const float r0 = (*scale - 0.75f) / 0.25f, r_1 = 1.0f - r0;
            *x = (r0 * delta_x_layer + r_1 * delta_x_below + (float)x_layer) * T->scale + T->offset;
            *y = (r0 * delta_y_layer + r_1 * delta_y_below + (float)y_layer) * T->scale + T->offset;
        }
    } else {
        float delta_x_below, delta_y_below;
        const float max_below = brk_max_below(S, layer, x_layer, y_layer, center, ismax, &delta_x_below, &delta_y_below);
        if (!*ismax) return 0.0f;
        brk_patch9(T, x_layer, y_layer, s);
        float delta_x_layer, delta_y_layer;
        const float max_layer = brk_subpixel2d(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], &delta_x_layer, &delta_y_layer);
        const float mid = (float)center > max_layer ? (float)center : max_layer;
        *scale = brk_refine1d(1, max_below, mid, max_above, &max);
        if (*scale > 1.0f) {
            const float r0 = 4.0f - *scale * 3.0f, r1 = 1.0f - r0;
            *x = (r0 * delta_x_layer + r1 * delta_x_above + (float)x_layer) * T->scale + T->offset;
            *y = (r0 * delta_y_layer + r1 * delta_y_above + (float)y_layer) * T->scale + T->offset;
        } else {
            const float r0 = *scale * 3.0f - 2.0f, r_1 = 1.0f - r0;
            *x = (r0 * delta_x_layer + r_1 * delta_x_below + (float)x_layer) * T->scale + T->offset;
            *y = (r0 * delta_y_layer + r_1 * delta_y_below + (float)y_layer) * T->scale + T->offset;
        }
    }
    *scale *= T->scale;
    return max;
}

static void brk_space_free(brk_space* S) {
    for (int i = 0; i < S->layers; ++i) { free(S->L[i].img); free(S->L[i].scores); }
}

/* BriskScaleSpace::constructPyramid: layer 0 = image, layer 1 = 2/3, then layer i = half of layer i-2 */
static int brk_space_build(brk_space* S, const uint8_t* gray, int w, int h, int stride, int octaves) {
    memset(S, 0, sizeof(*S));
    S->layers = octaves == 0 ? 1 : 2 * octaves;
    if (S->layers > BRK_MAX_LAYERS) return -1;
    for (int i = 0; i < S->layers; ++i) {
        brk_layer* L = &S->L[i];
        if (i == 0) { L->w = w; L->h = h; L->scale = 1.0f; L->offset = 0.0f; }
        else if (i == 1) { L->w = 2 * (w / 3); L->h = 2 * (h / 3); L->scale = 1.5f; L->offset = 0.5f * L->scale - 0.5f; }
        else { const brk_layer* P = &S->L[i - 2]; L->w = P->w / 2; L->h = P->h / 2; L->scale = P->scale * 2.0f; L->offset = 0.5f * L->scale - 0.5f; }
        if (L->w < 8 || L->h < 8) { S->layers = i; break; }
        L->img = (uint8_t*)malloc((size_t)L->w * L->h);
        L->scores = (uint8_t*)calloc((size_t)L->w * L->h, 1);
        if (i == 0) for (int y = 0; y < h; ++y) memcpy(L->img + (size_t)y * w, gray + (size_t)y * stride, (size_t)w);
        else if (i == 1) orc_resize_area_u8(S->L[0].img, w, h, w, L->img, L->w, L->h, L->w);
        else orc_resize_area_u8(S->L[i - 2].img, S->L[i - 2].w, S->L[i - 2].h, S->L[i - 2].w, L->img, L->w, L->h, L->w);
    }
    return 0;
}

typedef struct { float x, y, size, response; int layer; } brk_kp;

/* BriskScaleSpace::getKeypoints.  mode: ORC_BRISK_SEQUENTIAL (0) lazy score cache as in the original; ORC_BRISK_DENSE (1). */
static int brk_detect(brk_space* S, int threshold, int mode, brk_kp** out) {
    int cap = 4096, n = 0;
    brk_kp* K = (brk_kp*)malloc(sizeof(brk_kp) * (size_t)cap);
    int* ax[BRK_MAX_LAYERS]; int an[BRK_MAX_LAYERS];
    /* getAgastPoints: OAST 9-16 detections at `threshold` (no NMS), raster order; their scores are written to scores_ */
    for (int i = 0; i < S->layers; ++i) {
        brk_layer* L = &S->L[i];
        L->dense = 0;
        int acap = 1024; an[i] = 0; ax[i] = (int*)malloc(sizeof(int) * 2 * (size_t)acap);
        for (int y = 3; y < L->h - 3; ++y)
            for (int x = 3; x < L->w - 3; ++x) {
                const int s = brk_score_9_16(L->img, L->w, x, y);
                if (mode >= 1 && s >= 1) L->scores[(size_t)y * L->w + x] = (uint8_t)(s > 255 ? 255 : s);
                if (s >= threshold) {
                    if (an[i] == acap) { acap *= 2; ax[i] = (int*)realloc(ax[i], sizeof(int) * 2 * (size_t)acap); }
                    ax[i][2 * an[i]] = x; ax[i][2 * an[i] + 1] = y; ++an[i];
                    L->scores[(size_t)y * L->w + x] = (uint8_t)(s > 255 ? 255 : s);
                }
            }
    }
    /* dense contract: isMax2D sees the thresholded score image (sub-threshold entries read as 0), every getAgastScore the true score */
    uint8_t* thr[BRK_MAX_LAYERS];
    for (int i = 0; i < S->layers; ++i) {
        thr[i] = NULL;
        if (mode == 1) {
            brk_layer* L = &S->L[i];
            L->dense = 1;
            thr[i] = (uint8_t*)malloc((size_t)L->w * L->h);
            for (size_t p = 0; p < (size_t)L->w * L->h; ++p) thr[i][p] = L->scores[p] >= threshold ? L->scores[p] : 0;
        }
    }
    for (int i = 0; i < S->layers; ++i) {
        brk_layer* L = &S->L[i];
        brk_layer Lmax = *L;                       /* view used by isMax2D */
        if (mode == 1) Lmax.scores = thr[i];
        if (mode == 2) { L->dense = 1; }
        for (int k = 0; k < an[i]; ++k) {
            const int px = ax[i][2 * k], py = ax[i][2 * k + 1];
            if (!brk_is_max2d(&Lmax, px, py)) continue;
            brk_kp kp; int ismax = 0;
            if (S->layers == 1) {
                int s[9]; brk_patch9(L, px, py, s);
                float dx, dy;
                const float mx = brk_subpixel2d(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], &dx, &dy);
                kp.x = (float)px + dx; kp.y = (float)py + dy; kp.size = BRK_BASIC_SIZE; kp.response = mx; kp.layer = 0;
            } else if (i == S->layers - 1) {
                float dx, dy;
                /* l.getAgastScore(point.x, point.y, safeThreshold_): float overload -> four integer look-ups with the detection threshold */
                const int center = brk_get_score_f(L, (float)px, (float)py, threshold);
                (void)brk_max_below(S, i, px, py, center, &ismax, &dx, &dy);
                if (!ismax) continue;
                /* the nine patch scores go through the float overload too (each touches a 2x2 block with threshold 1) */
                int s[9];
                s[0] = brk_get_score_f(L, (float)px - 1, (float)py - 1, 1); s[3] = brk_get_score_f(L, (float)px, (float)py - 1, 1);
                s[6] = brk_get_score_f(L, (float)px + 1, (float)py - 1, 1); s[7] = brk_get_score_f(L, (float)px + 1, (float)py, 1);
                s[4] = brk_get_score_f(L, (float)px, (float)py, 1); s[1] = brk_get_score_f(L, (float)px - 1, (float)py, 1);
                s[2] = brk_get_score_f(L, (float)px - 1, (float)py + 1, 1); s[5] = brk_get_score_f(L, (float)px, (float)py + 1, 1);
                s[8] = brk_get_score_f(L, (float)px + 1, (float)py + 1, 1);
                float delta_x, delta_y;
                const float mx = brk_subpixel2d(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], &delta_x, &delta_y);
                kp.x = ((float)px + delta_x) * L->scale + L->offset; kp.y = ((float)py + delta_y) * L->scale + L->offset;
                kp.size = BRK_BASIC_SIZE * L->scale; kp.response = mx; kp.layer = i;
            } else {
                float x, y, scale;
                const float score = brk_refine3d(S, i, px, py, &x, &y, &scale, &ismax);
                if (!ismax) continue;
                if (!(score > (float)threshold)) continue;
                kp.x = x; kp.y = y; kp.size = BRK_BASIC_SIZE * scale; kp.response = score; kp.layer = i;
            }
            if (n == cap) { cap *= 2; K = (brk_kp*)realloc(K, sizeof(brk_kp) * (size_t)cap); }
            K[n++] = kp;
        }
    }
    for (int i = 0; i < S->layers; ++i) { free(ax[i]); free(thr[i]); }
    *out = K;
    return n;
}

/* ---------------------------------------------------------------------------------------------------------------------
 * Sampling pattern + pairs (BRISK_Impl::generateKernel), built once.
 * ------------------------------------------------------------------------------------------------------------------- */
typedef struct { float x, y, sigma; } brk_pp;
typedef struct { unsigned i, j; int wdx, wdy; } brk_long;
typedef struct { unsigned i, j; } brk_short;
static brk_pp* g_pat = NULL;                    /* [scale][rot][point] */
static float g_scale_list[BRK_SCALES];
static unsigned g_size_list[BRK_SCALES];
static brk_long g_long[BRK_POINTS * (BRK_POINTS - 1) / 2];
static brk_short g_short[BRK_POINTS * (BRK_POINTS - 1) / 2];      /* the paper's 512 short pairs, enumeration order */
static brk_short g_short48[384];                                  /* 384 shortest of them, enumeration order */
static int g_nlong = 0, g_nshort = 0;

static int cmp_pairdist(const void* a, const void* b) {
    const float* x = (const float*)a; const float* y = (const float*)b;
    if (x[0] < y[0]) return -1;
    if (x[0] > y[0]) return 1;
    return x[1] < y[1] ? -1 : (x[1] > y[1] ? 1 : 0);
}

static void brk_build_pattern_once(void);
static pthread_once_t g_pat_once = PTHREAD_ONCE_INIT;
static void brk_build_pattern(void) { pthread_once(&g_pat_once, brk_build_pattern_once); }      /* thread-safe (batch CPU arm) */
static void brk_build_pattern_once(void) {
    const float f = 0.85f * 1.0f;
    const float rList[5] = {(float)(f * 0.), (float)(f * 2.9), (float)(f * 4.9), (float)(f * 7.4), (float)(f * 10.8)};
    const int nList[5] = {1, 10, 14, 15, 20};
    const float dMax = 5.85f, dMin = 8.2f;
    brk_pp* pat = (brk_pp*)malloc(sizeof(brk_pp) * (size_t)BRK_POINTS * BRK_SCALES * BRK_NROT);
    const float lb_scale = (float)((double)logf(30.f) / log(2.0));          /* std::log(scalerange_) on a float is logf */
    const float lb_scale_step = lb_scale / (float)BRK_SCALES;
    const float sigma_scale = 1.3f;
    brk_pp* it = pat;
    for (unsigned scale = 0; scale < BRK_SCALES; ++scale) {
        g_scale_list[scale] = (float)pow(2.0, (double)((float)scale * lb_scale_step));
        g_size_list[scale] = 0;
        for (unsigned rot = 0; rot < BRK_NROT; ++rot) {
            const double theta = (double)rot * 2 * M_PI / (double)BRK_NROT;
            for (int ring = 0; ring < 5; ++ring)
                for (int num = 0; num < nList[ring]; ++num) {
                    const double alpha = ((double)num) * 2 * M_PI / (double)nList[ring];
                    it->x = (float)((double)(g_scale_list[scale] * rList[ring]) * cos(alpha + theta));
                    it->y = (float)((double)(g_scale_list[scale] * rList[ring]) * sin(alpha + theta));
                    if (ring == 0) it->sigma = sigma_scale * g_scale_list[scale] * 0.5f;
                    else it->sigma = (float)((double)(sigma_scale * g_scale_list[scale]) * ((double)rList[ring]) * sin(M_PI / nList[ring]));
                    const unsigned size = (unsigned)((int)ceil((double)((g_scale_list[scale] * rList[ring]) + it->sigma)) + 1);
                    if (g_size_list[scale] < size) g_size_list[scale] = size;
                    ++it;
                }
        }
    }
    const float dMin_sq = dMin * dMin, dMax_sq = dMax * dMax;
    g_nlong = 0; g_nshort = 0;
    float sd[BRK_POINTS * (BRK_POINTS - 1) / 2][2];
    for (unsigned i = 1; i < BRK_POINTS; ++i)
        for (unsigned j = 0; j < i; ++j) {
            const float dx = pat[j].x - pat[i].x, dy = pat[j].y - pat[i].y;
            const float norm_sq = (dx * dx + dy * dy);
            if (norm_sq > dMin_sq) {
                brk_long* lp = &g_long[g_nlong++];
                lp->wdx = (int)((double)(dx / (norm_sq)) * 2048.0 + 0.5);
                lp->wdy = (int)((double)(dy / (norm_sq)) * 2048.0 + 0.5);
                lp->i = i; lp->j = j;
            } else if (norm_sq < dMax_sq) {
                sd[g_nshort][0] = norm_sq; sd[g_nshort][1] = (float)g_nshort;
                g_short[g_nshort].i = i; g_short[g_nshort].j = j; ++g_nshort;
            }
        }
    /* 48-byte stand-in table: the 384 shortest short pairs (ties by enumeration order), kept in enumeration order */
    {
        qsort(sd, (size_t)g_nshort, sizeof(sd[0]), cmp_pairdist);
        char* take = (char*)calloc((size_t)g_nshort, 1);
        for (int k = 0; k < 384 && k < g_nshort; ++k) take[(int)sd[k][1]] = 1;
        int m = 0;
        for (int k = 0; k < g_nshort; ++k) if (take[k] && m < 384) g_short48[m++] = g_short[k];
        free(take);
    }
    g_pat = pat;
}

/* BRISK_Impl::smoothedIntensity */
static int brk_smoothed(const uint8_t* image, int imagecols, const int* integral, float key_x, float key_y, unsigned scale, unsigned rot,
                        unsigned point) {
    const brk_pp* bp = &g_pat[((size_t)scale * BRK_NROT + rot) * BRK_POINTS + point];
    const float xf = bp->x + key_x, yf = bp->y + key_y;
    const int x = (int)xf, y = (int)yf;
    const float sigma_half = bp->sigma;
    const float area = 4.0f * sigma_half * sigma_half;
    int ret_val;
    if (sigma_half < 0.5) {
        const int r_x = (int)((xf - (float)x) * 1024), r_y = (int)((yf - (float)y) * 1024);
        const int r_x_1 = (1024 - r_x), r_y_1 = (1024 - r_y);
        const uint8_t* ptr = image + x + y * imagecols;
        ret_val = (r_x_1 * r_y_1 * (int)(*ptr)); ptr++;
        ret_val += (r_x * r_y_1 * (int)(*ptr)); ptr += imagecols;
        ret_val += (r_x * r_y * (int)(*ptr)); ptr--;
        ret_val += (r_x_1 * r_y * (int)(*ptr));
        return (ret_val + 512) / 1024;
    }
    const int scaling = (int)(4194304.0 / area);
    const int scaling2 = (int)((float)scaling * area / 1024.0);
    const int integralcols = imagecols + 1;
    const float x_1 = xf - sigma_half, x1 = xf + sigma_half, y_1 = yf - sigma_half, y1 = yf + sigma_half;
    const int x_left = (int)(x_1 + 0.5), y_top = (int)(y_1 + 0.5), x_right = (int)(x1 + 0.5), y_bottom = (int)(y1 + 0.5);
    const float r_x_1 = (float)x_left - x_1 + 0.5f, r_y_1 = (float)y_top - y_1 + 0.5f;
    const float r_x1 = x1 - (float)x_right + 0.5f, r_y1 = y1 - (float)y_bottom + 0.5f;
    const int dx = x_right - x_left - 1, dy = y_bottom - y_top - 1;
    const int A = (int)((r_x_1 * r_y_1) * (float)scaling), B = (int)((r_x1 * r_y_1) * (float)scaling);
    const int C = (int)((r_x1 * r_y1) * (float)scaling), D = (int)((r_x_1 * r_y1) * (float)scaling);
    const int r_x_1_i = (int)(r_x_1 * (float)scaling), r_y_1_i = (int)(r_y_1 * (float)scaling);
    const int r_x1_i = (int)(r_x1 * (float)scaling), r_y1_i = (int)(r_y1 * (float)scaling);
    if (dx + dy > 2) {
        const uint8_t* ptr = image + x_left + imagecols * y_top;
        ret_val = A * (int)(*ptr); ptr += dx + 1;
        ret_val += B * (int)(*ptr); ptr += (dy + 1) * imagecols;
        ret_val += C * (int)(*ptr); ptr -= dx + 1;
        ret_val += D * (int)(*ptr);
        const int* pi = integral + x_left + integralcols * y_top + 1;
        const int tmp1 = (*pi); pi += dx;
        const int tmp2 = (*pi); pi += integralcols;
        const int tmp3 = (*pi); pi++;
        const int tmp4 = (*pi); pi += dy * integralcols;
        const int tmp5 = (*pi); pi--;
        const int tmp6 = (*pi); pi += integralcols;
        const int tmp7 = (*pi); pi -= dx;
        const int tmp8 = (*pi); pi -= integralcols;
        const int tmp9 = (*pi); pi--;
        const int tmp10 = (*pi); pi -= dy * integralcols;
        const int tmp11 = (*pi); pi++;
        const int tmp12 = (*pi);
        const int upper = (tmp3 - tmp2 + tmp1 - tmp12) * r_y_1_i;
        const int middle = (tmp6 - tmp3 + tmp12 - tmp9) * scaling;
        const int left = (tmp9 - tmp12 + tmp11 - tmp10) * r_x_1_i;
        const int right = (tmp5 - tmp4 + tmp3 - tmp6) * r_x1_i;
        const int bottom = (tmp7 - tmp6 + tmp9 - tmp8) * r_y1_i;
        return (ret_val + upper + middle + left + right + bottom + scaling2 / 2) / scaling2;
    }
    const uint8_t* ptr = image + x_left + imagecols * y_top;
    ret_val = A * (int)(*ptr); ptr++;
    const uint8_t* end1 = ptr + dx;
    for (; ptr < end1; ptr++) ret_val += r_y_1_i * (int)(*ptr);
    ret_val += B * (int)(*ptr);
    ptr += imagecols - dx - 1;
    const uint8_t* end_j = ptr + dy * imagecols;
    for (; ptr < end_j; ptr += imagecols - dx - 1) {
        ret_val += r_x_1_i * (int)(*ptr); ptr++;
        const uint8_t* end2 = ptr + dx;
        for (; ptr < end2; ptr++) ret_val += (int)(*ptr) * scaling;
        ret_val += r_x1_i * (int)(*ptr);
    }
    ret_val += D * (int)(*ptr); ptr++;
    const uint8_t* end3 = ptr + dx;
    for (; ptr < end3; ptr++) ret_val += r_y1_i * (int)(*ptr);
    ret_val += C * (int)(*ptr);
    return (ret_val + scaling2 / 2) / scaling2;
}

/* atan2 in double from +,-,*,/ only (identical results on the CPU and in the CUDA kernel): argument reduction to [0, tan(pi/12)]
 * + odd polynomial; absolute error < 1e-15, so the float rounding is the correctly rounded atan2f except in ~1e-8 of the cases. */
double orc_brisk_atan2(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    if (ax == 0.0 && ay == 0.0) return 0.0;
    const int swap = ay > ax;
    double t = swap ? ax / ay : ay / ax;                    /* in [0, 1] */
    /* reduce with atan(t) = pi/6 + atan((t*sqrt3 - 1) / (sqrt3 + t)) for t > tan(pi/12) */
    const double SQ3 = 1.7320508075688772, T12 = 0.2679491924311227, PI6 = 0.5235987755982989, PI2 = 1.5707963267948966, PI = 3.141592653589793;
    int red = 0;
    if (t > T12) { t = (t * SQ3 - 1.0) / (SQ3 + t); red = 1; }
    const double z = t * t;
    /* atan(t) = t * sum_{k>=0} (-1)^k z^k / (2k+1), |t| <= 0.268: 14 terms give < 1e-17 */
    double p = 1.0 / 29.0;
    p = 1.0 / 27.0 - z * p; p = 1.0 / 25.0 - z * p; p = 1.0 / 23.0 - z * p; p = 1.0 / 21.0 - z * p; p = 1.0 / 19.0 - z * p;
    p = 1.0 / 17.0 - z * p; p = 1.0 / 15.0 - z * p; p = 1.0 / 13.0 - z * p; p = 1.0 / 11.0 - z * p; p = 1.0 / 9.0 - z * p;
    p = 1.0 / 7.0 - z * p; p = 1.0 / 5.0 - z * p; p = 1.0 / 3.0 - z * p; p = 1.0 - z * p;
    double a = t * p;
    if (red) a += PI6;
    if (swap) a = PI2 - a;
    if (x < 0.0) a = PI - a;
    if (y < 0.0) a = -a;
    return a;
}

/* size -> pattern scale index exactly as BRISK_Impl::computeDescriptorsAndOrOrientation */
static unsigned brk_scale_index(float size) {
    static const float log2c = 0.693147180559945f;
    const float lb_scalerange = (float)(logf(30.f) / (log2c));
    const float basicSize06 = BRK_BASIC_SIZE * 0.6f;
    int scale = (int)((float)BRK_SCALES / lb_scalerange * (logf(size / (basicSize06)) / log2c) + 0.5);
    if (scale < 0) scale = 0;
    if (scale >= BRK_SCALES) scale = BRK_SCALES - 1;
    return (unsigned)scale;
}
int orc_brisk_scale_index(float size) { brk_build_pattern(); return (int)brk_scale_index(size); }
int orc_brisk_size_list(unsigned* out64) { brk_build_pattern(); memcpy(out64, g_size_list, sizeof(g_size_list)); return BRK_SCALES; }

static void brk_integral(const uint8_t* img, int w, int h, int stride, int* integral) {     /* cv::integral CV_32S, (h+1) x (w+1) */
    const int iw = w + 1;
    for (int x = 0; x < iw; ++x) integral[x] = 0;
    for (int y = 0; y < h; ++y) {
        int rs = 0;
        integral[(size_t)(y + 1) * iw] = 0;
        for (int x = 0; x < w; ++x) { rs += img[(size_t)y * stride + x]; integral[(size_t)(y + 1) * iw + x + 1] = integral[(size_t)y * iw + x + 1] + rs; }
    }
}

/* BRISK_Impl::computeDescriptorsAndOrOrientation on a keypoint list (in place): removes border keypoints, writes angle and the
 * descriptor rows.  nbytes = 64: the paper's 512 short pairs (cv2-comparable); 48: the 384-pair stand-in table.
 * libm_angle != 0: angle through libm atan2f like cv2; 0: shared double polynomial (product contract).  Returns the new count. */
int orc_brisk_describe(const uint8_t* gray, int w, int h, int stride, orc_keypoint* kps, int n, int nbytes, int libm_angle, uint8_t* desc,
                       int* kept_index) {
    brk_build_pattern();
    uint8_t* image = (uint8_t*)malloc((size_t)w * h);
    for (int y = 0; y < h; ++y) memcpy(image + (size_t)y * w, gray + (size_t)y * stride, (size_t)w);
    int* integral = (int*)malloc(sizeof(int) * (size_t)(w + 1) * (h + 1));
    brk_integral(image, w, h, w, integral);
    const brk_short* sp = nbytes == 48 ? g_short48 : g_short;
    const int nsp = nbytes == 48 ? 384 : g_nshort;
    int m = 0;
    for (int k = 0; k < n; ++k) {
        orc_keypoint kp = kps[k];
        const unsigned scale = brk_scale_index(kp.size);
        const int border = (int)g_size_list[scale];
        const float minX = (float)border, minY = (float)border, maxX = (float)(w - border), maxY = (float)(h - border);
        if ((kp.x < minX) || (kp.x >= maxX) || (kp.y < minY) || (kp.y >= maxY)) continue;
        int values[BRK_POINTS];
        for (unsigned i = 0; i < BRK_POINTS; ++i) values[i] = brk_smoothed(image, w, integral, kp.x, kp.y, scale, 0, i);
        int direction0 = 0, direction1 = 0;
        for (int p = 0; p < g_nlong; ++p) {
            const int delta_t = values[g_long[p].i] - values[g_long[p].j];
            direction0 += delta_t * g_long[p].wdx / 1024;
            direction1 += delta_t * g_long[p].wdy / 1024;
        }
        if (libm_angle == 2) kp.angle = (float)(atan2((double)(float)direction1, (double)(float)direction0) / M_PI * 180.0);
        else if (libm_angle) kp.angle = (float)(atan2f((float)direction1, (float)direction0) / M_PI * 180.0);
        else kp.angle = (float)(orc_brisk_atan2((double)(float)direction1, (double)(float)direction0) / M_PI * 180.0);
        int theta;
        if (kp.angle == -1) theta = 0;
        else {
            theta = (int)(BRK_NROT * (kp.angle / (360.0)) + 0.5);
            if (theta < 0) theta += BRK_NROT;
            if (theta >= (int)BRK_NROT) theta -= BRK_NROT;
        }
        if (kp.angle < 0) kp.angle += 360.f;
        for (unsigned i = 0; i < BRK_POINTS; ++i) values[i] = brk_smoothed(image, w, integral, kp.x, kp.y, scale, (unsigned)theta, i);
        uint8_t* row = desc + (size_t)nbytes * m;
        memset(row, 0, (size_t)nbytes);
        for (int p = 0; p < nsp; ++p)
            if (values[sp[p].i] > values[sp[p].j]) row[p >> 3] |= (uint8_t)(1u << (p & 7));
        if (kept_index) kept_index[m] = k;
        kps[m++] = kp;
    }
    free(image); free(integral);
    return m;
}

/* raw detector tap: BriskFeatureDetector(threshold, octaves).detect -> (x, y, size, response, layer) per keypoint, detection order */
int orc_brisk_detect(const uint8_t* gray, int w, int h, int stride, int threshold, int octaves, int mode, float* out5, int cap) {
    brk_space S;
    if (brk_space_build(&S, gray, w, h, stride, octaves)) return -1;
    brk_kp* K = NULL;
    const int n = brk_detect(&S, threshold, mode, &K);
    if (n <= cap) for (int i = 0; i < n; ++i) { out5[5 * i] = K[i].x; out5[5 * i + 1] = K[i].y; out5[5 * i + 2] = K[i].size; out5[5 * i + 3] = K[i].response; out5[5 * i + 4] = (float)K[i].layer; }
    free(K); brk_space_free(&S);
    return n <= cap ? n : -n;
}

/* pyramid / score taps for the stage tests: what = 0 layer image, 1 true 9-16 score image (0 below 1) */
long orc_brisk_layer(const uint8_t* gray, int w, int h, int stride, int octaves, int what, int layer, uint8_t* out, int* ow, int* oh) {
    brk_space S;
    if (brk_space_build(&S, gray, w, h, stride, octaves)) return -1;
    if (layer < 0 || layer >= S.layers) { brk_space_free(&S); return -1; }
    const brk_layer* L = &S.L[layer];
    *ow = L->w; *oh = L->h;
    if (out) {
        if (what == 0) memcpy(out, L->img, (size_t)L->w * L->h);
        else {
            memset(out, 0, (size_t)L->w * L->h);
            for (int y = 3; y < L->h - 3; ++y)
                for (int x = 3; x < L->w - 3; ++x) { const int s = brk_score_9_16(L->img, L->w, x, y); out[(size_t)y * L->w + x] = (uint8_t)(s >= 1 ? (s > 255 ? 255 : s) : 0); }
        }
    }
    const long nb = (long)L->w * L->h;
    brk_space_free(&S);
    return nb;
}

/* full FeatureExtractor_brisk48::operator() (src/Feature_brisk48.cpp:11-60 + src/FeatureExtractor.cpp:111-142) */
int orc_brisk48_extract(const uint8_t* gray, int w, int h, int stride, int nfeatures, int nlevels, float scale_factor, float detect_th,
                        int mode, orc_keypoint* kps, uint8_t* desc, float* kpsize, int cap, int* n_out, int* n_detected) {
    if (nlevels < 1 || nlevels > ORC_MAX_LEVELS) return -1;
    brk_space S;
    if (brk_space_build(&S, gray, w, h, stride, nlevels / 2)) return -1;
    brk_kp* K = NULL;
    const int n = brk_detect(&S, (int)detect_th, mode, &K);
    brk_space_free(&S);
    if (n_detected) *n_detected = n;
    int q_ext[ORC_MAX_LEVELS];
    orc_features_per_level(nfeatures, nlevels, scale_factor, q_ext);
    float* lx = (float*)malloc(sizeof(float) * (size_t)(n + 1)); float* ly = (float*)malloc(sizeof(float) * (size_t)(n + 1));
    float* lr = (float*)malloc(sizeof(float) * (size_t)(n + 1)); int* idx = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    int* keep = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    orc_keypoint* merged = (orc_keypoint*)malloc(sizeof(orc_keypoint) * (size_t)(n + 1));
    int m = 0, rc = 0;
    for (int l = 0; l < nlevels; ++l) {          /* keypoints_level[octave] in std::map order; levels >= nlevels cannot occur (layers = 2*(nlevels/2)) */
        int nl = 0;
        for (int i = 0; i < n; ++i) if (K[i].layer == l) { lx[nl] = K[i].x; ly[nl] = K[i].y; lr[nl] = K[i].response; idx[nl] = i; ++nl; }
        if (!nl) continue;
        const int nk = orc_distribute_octree(lx, ly, lr, NULL, nl, 0, w, 0, h, q_ext[l], keep, nl);
        for (int j = 0; j < nk; ++j) {
            const brk_kp* p = &K[idx[keep[j]]];
            orc_keypoint* kp = &merged[m++];
            kp->x = p->x; kp->y = p->y; kp->size = p->size; kp->angle = -1.0f; kp->response = p->response; kp->octave = p->layer; kp->class_id = -1;
        }
    }
    uint8_t* d = (uint8_t*)malloc((size_t)48 * (size_t)(m + 1));
    const int mk = orc_brisk_describe(gray, w, h, stride, merged, m, 48, 0, d, NULL);
    if (mk > cap) rc = -2;
    else {
        const float maxSize0 = powf(1.2f, (float)(8 - 1.0)), maxSize = maxSize0, minSize = 1.0f;
        for (int i = 0; i < mk; ++i) {
            kps[i] = merged[i];
            memcpy(desc + (size_t)48 * i, d + (size_t)48 * i, 48);
            if (kpsize) {
                const float s = powf(scale_factor, (float)merged[i].octave);
                float sn = maxSize;
                if (maxSize > minSize) sn = 1.0f + (s - minSize) * (maxSize0 - 1.0f) / (maxSize - minSize);
                kpsize[i] = sn;
            }
        }
        if (n_out) *n_out = mk;
    }
    free(lx); free(ly); free(lr); free(idx); free(keep); free(merged); free(d); free(K);
    return rc;
}

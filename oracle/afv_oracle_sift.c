/*
 * afv_oracle_sift.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) of the sift128 extraction path.
 *
 * PARITY UNPINNED.  The reference's sift128 arithmetic lives in SiftGPU (fontan::siftgpu, un-pinned, GLSL path;
 * reference call sites src/Feature_sift128.cpp:9-62 (arguments), :76-98 (keypoints), :100-118 (descriptors)), which is
 * NOT vendored under the reference tree and cannot run here (needs an OpenGL context).  This file restates the
 * PUBLISHED algorithm (Lowe 2004 DoG SIFT in the sift++ / SiftGPU parameterisation the reference selects with its
 * argument list) and everything the reference itself does around it:
 *   -fo 0 (no up-sampling), -d 3 (3 DoG levels: 6 Gaussian levels -1..4 per octave), sigma0 = 1.6*2^(1/3),
 *   nominal input blur 0.5, -no 8 (octaves = min(8, floor(log2(min(w,h))) - 3)), DoG threshold 0.02/3 (default -t),
 *   -e 10, sub-pixel localisation on (one Newton step; 0.8*t pre-test), at most 2 orientations per keypoint (default
 *   -m 2), -loweo (pixel centres), -tc2 nfeatures (soft limit, coarsest levels first), -da (darkness adaptivity:
 *   threshold scaled by min(2*g + 0.1, 1) with g the smoothed intensity), unit-L2 128-float descriptors
 *   (4x4x8, clamp 0.2, renormalise);
 *   then src/Feature_sift128.cpp:84-97 (octave = int(log2(s/1.6454)), response 1, class_id = list row),
 *   filterKeypoints_notScaled (src/FeatureExtractor.cpp:276-284 -> DistributeOctTree, quota mnFeaturesPerLevel),
 *   computeDescriptors (row gather), mergeKeypointLevels (:296-308), computeSize (:132-142, GetKeypointSize =
 *   powf(scaleFactor0, octave), src/Feature_sift128.cpp:124-126).
 *
 * Cross-check (tests/test_oracle_sift.py::test_pinned_to_cv2_precise_per_octave): against cv2 4.13.0 SIFT with
 * enable_precise_upscale, 86 / 92 % of cv2's octave 0 / 1 keypoints have a keypoint of this file at the same scale within
 * 0.25 px (no systematic offset, 0.08 px std per axis) and their orientations agree to 2.2 degrees of spread with no
 * offset; the 128-float descriptors equal cv2's up to the direction in which the 8 orientation bins are counted (median cosine
 * 0.997).  (Until that check existed sift_orientations() mapped a histogram peak to the LOWER EDGE of its 10-degree bin:
 * a constant -5.0 degrees against cv2, sift++ and Lowe, all of which refer the peak to the bin centre; corrected.)
 *
 * Arithmetic contract (what the CUDA path must reproduce bit for bit): IEEE float32, round-to-nearest, NO fused
 * multiply-add except the explicit fmaf() of the Gaussian taps, operation order exactly as written here; exp / atan2 / sin / cos are the polynomial forms below (not
 * libm); every histogram is accumulated in INTEGERS (contributions quantised with rintf(v * 2^20)), so the sums do not
 * depend on the order in which a parallel implementation adds them.
 */
#include "afv_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SIFT_S 3                 /* DoG levels per octave (-d 3) */
#define SIFT_NL 6                /* Gaussian levels per octave: level -1 .. 4 */
#define SIFT_MAX_OCT 8           /* -no 8 */
#define SIFT_BORDER 5
#define SIFT_QSCALE 1048576.0f   /* 2^20 */
#define SIFT_PI 3.14159265358979323846f
#define SIFT_2PI 6.28318530717958647692f

/* ------------------------------------------------------------------ deterministic elementary functions ---------- */
static inline float sift_exp2(float t) {
    if (t < -126.f) t = -126.f;
    if (t > 126.f) t = 126.f;
    const float n = rintf(t);
    const float f = t - n;                           /* [-0.5, 0.5] */
    float p = 0x1.430912p-13f;                       /* (ln 2)^6 / 720 */
    p = p * f + 0x1.5d87fep-10f;                     /* (ln 2)^5 / 120 */
    p = p * f + 0x1.3b2ab6p-7f;                      /* (ln 2)^4 / 24  */
    p = p * f + 0x1.c6b08ep-5f;                      /* (ln 2)^3 / 6   */
    p = p * f + 0x1.ebfbep-3f;                       /* (ln 2)^2 / 2   */
    p = p * f + 0x1.62e43p-1f;                       /* ln 2           */
    p = p * f + 1.0f;
    union { unsigned u; float f; } s; s.u = (unsigned)((int)n + 127) << 23;
    return p * s.f;
}
static inline float sift_exp(float x) { return sift_exp2(x * 0x1.715476p+0f); }

/* angle of (x, y) in radians, [0, 2 pi) */
static inline float sift_atan2(float y, float x) {
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

static inline void sift_sincos(float a, float* sn, float* cs) {
    const float q = rintf(a * 0x1.45f306p-1f);       /* a * 2/pi */
    const float r = a - q * 0x1.921fb6p+0f;          /* [-pi/4, pi/4] (+ rounding) */
    const float z = r * r;
    float s = 0x1.71de3ap-19f;                       /* 1/9! */
    s = s * z + -0x1.a01a02p-13f;                    /* -1/7! */
    s = s * z + 0x1.111112p-7f;                      /* 1/5! */
    s = s * z + -0x1.555556p-3f;                     /* -1/3! */
    s = s * z + 1.0f;
    s = s * r;
    float c = 0x1.a01a02p-16f;                       /* 1/8! */
    c = c * z + -0x1.6c16c2p-10f;                    /* -1/6! */
    c = c * z + 0x1.555556p-5f;                      /* 1/4! */
    c = c * z + -0.5f;
    c = c * z + 1.0f;
    const int k = ((int)q) & 3;
    if (k == 0) { *sn = s; *cs = c; }
    else if (k == 1) { *sn = c; *cs = -s; }
    else if (k == 2) { *sn = -s; *cs = -c; }
    else { *sn = -c; *cs = s; }
}

float orc_sift_exp(float x) { return sift_exp(x); }
float orc_sift_exp2(float x) { return sift_exp2(x); }
float orc_sift_atan2(float y, float x) { return sift_atan2(y, x); }
void  orc_sift_sincos(float a, float* s, float* c) { sift_sincos(a, s, c); }

/* ------------------------------------------------------------------ scale space ------------------------------- */
/* Gaussian taps for one blur step: radius ceil(4 sigma), weights in double, normalised, narrowed to float. */
int orc_sift_gauss_kernel(double sigma, float* taps /* [0..r] centre outward */, int max_r) {
    int r = (int)ceil(4.0 * sigma);
    if (r < 1) r = 1;
    if (r > max_r) return -1;
    double sum = 0.0, wd[64];
    for (int j = 0; j <= r; ++j) { wd[j] = exp(-(double)(j * j) / (2.0 * sigma * sigma)); sum += j ? 2.0 * wd[j] : wd[j]; }
    for (int j = 0; j <= r; ++j) taps[j] = (float)(wd[j] / sum);
    return r;
}

/* the per-level incremental blur sigmas (sift++ / SiftGPU ParamSIFT): level -1 of octave 0 from the input,
 * level l (0..4) from level l-1 */
void orc_sift_sigmas(double* dsig /* [6] */) {
    const double k = pow(2.0, 1.0 / SIFT_S), sigma0 = 1.6 * k, sigman = 0.5;
    const double s_m1 = sigma0 / k;                                  /* level -1: 1.6 */
    dsig[0] = sqrt(s_m1 * s_m1 - sigman * sigman);
    const double dsigma0 = sigma0 * sqrt(1.0 - 1.0 / (k * k));
    for (int l = 0; l <= 4; ++l) dsig[l + 1] = dsigma0 * pow(k, (double)l);
}

int orc_sift_num_octaves(int w, int h) {
    int m = w < h ? w : h, lg = 0;
    while ((1 << (lg + 1)) <= m) ++lg;                               /* floor(log2(min)) */
    int o = lg - 3;
    if (o > SIFT_MAX_OCT) o = SIFT_MAX_OCT;
    if (o < 1) o = 1;
    return o;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* separable blur, rows then columns, clamp-to-edge; acc = t0*c; acc = fma(tj, l + r, acc) for j = 1..r (the one place
 * where the contract FUSES: one rounding per tap, and one FFMA instead of FMUL + FADD on the GPU) */
static void blur_sep(const float* src, float* dst, float* tmp, int w, int h, const float* taps, int r) {
    for (int y = 0; y < h; ++y) {
        const float* s = src + (size_t)y * w;
        float* t = tmp + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            float acc = taps[0] * s[x];
            for (int j = 1; j <= r; ++j) acc = fmaf(taps[j], s[clampi(x - j, 0, w - 1)] + s[clampi(x + j, 0, w - 1)], acc);
            t[x] = acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        float* d = dst + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            float acc = taps[0] * tmp[(size_t)y * w + x];
            for (int j = 1; j <= r; ++j)
                acc = fmaf(taps[j], tmp[(size_t)clampi(y - j, 0, h - 1) * w + x] + tmp[(size_t)clampi(y + j, 0, h - 1) * w + x], acc);
            d[x] = acc;
        }
    }
}

typedef struct { int w, h; float* g[SIFT_NL]; float* d[SIFT_NL - 1]; } sift_oct;
typedef struct { int no; sift_oct o[SIFT_MAX_OCT]; } sift_pyr;

static void pyr_free(sift_pyr* P) {
    for (int o = 0; o < P->no; ++o) {
        for (int i = 0; i < SIFT_NL; ++i) free(P->o[o].g[i]);
        for (int i = 0; i < SIFT_NL - 1; ++i) free(P->o[o].d[i]);
    }
}

static int pyr_build(sift_pyr* P, const uint8_t* gray, int w, int h, int stride) {
    double dsig[6];
    orc_sift_sigmas(dsig);
    float taps[6][64]; int rad[6];
    for (int i = 0; i < 6; ++i) { rad[i] = orc_sift_gauss_kernel(dsig[i], taps[i], 63); if (rad[i] < 0) return -1; }
    P->no = orc_sift_num_octaves(w, h);
    float* tmp = (float*)malloc(sizeof(float) * (size_t)w * h);
    float* base = (float*)malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) base[(size_t)y * w + x] = (float)gray[(size_t)y * stride + x] / 255.0f;
    int ow = w, oh = h;
    for (int o = 0; o < P->no; ++o) {
        sift_oct* O = &P->o[o];
        O->w = ow; O->h = oh;
        const size_t npx = (size_t)ow * oh;
        for (int i = 0; i < SIFT_NL; ++i) O->g[i] = (float*)malloc(sizeof(float) * npx);
        for (int i = 0; i < SIFT_NL - 1; ++i) O->d[i] = (float*)malloc(sizeof(float) * npx);
        if (o == 0) blur_sep(base, O->g[0], tmp, ow, oh, taps[0], rad[0]);
        else {
            const sift_oct* Q = &P->o[o - 1];                        /* level 2 (index 3) of the previous octave, every 2nd pixel */
            for (int y = 0; y < oh; ++y) for (int x = 0; x < ow; ++x) O->g[0][(size_t)y * ow + x] = Q->g[SIFT_S][(size_t)(2 * y) * Q->w + 2 * x];
        }
        for (int i = 1; i < SIFT_NL; ++i) {
            blur_sep(O->g[i - 1], O->g[i], tmp, ow, oh, taps[i], rad[i]);
            for (size_t p = 0; p < npx; ++p) O->d[i - 1][p] = O->g[i][p] - O->g[i - 1][p];
        }
        ow /= 2; oh /= 2;
    }
    free(tmp); free(base);
    return 0;
}

/* tap for tests: Gaussian (what = 0) or DoG (what = 1) image `idx` of octave `oct`; returns w*h or <0 */
long orc_sift_scale_space(const uint8_t* gray, int w, int h, int stride, int what, int oct, int idx, float* out, int* ow, int* oh) {
    sift_pyr P;
    if (pyr_build(&P, gray, w, h, stride)) return -1;
    long n = -1;
    if (oct >= 0 && oct < P.no && idx >= 0 && idx < (what ? SIFT_NL - 1 : SIFT_NL)) {
        const sift_oct* O = &P.o[oct];
        n = (long)O->w * O->h;
        if (out) memcpy(out, what ? O->d[idx] : O->g[idx], sizeof(float) * (size_t)n);
        if (ow) *ow = O->w;
        if (oh) *oh = O->h;
    }
    pyr_free(&P);
    return n;
}

/* ------------------------------------------------------------------ detection ---------------------------------- */
typedef struct { float x, y, ls; int oct, lvl; float sigma; /* octave-relative */ float ori; } sift_feat;

/* extremum + threshold + edge + one Newton step; returns 1 and fills (xo, yo, ls) when the point survives */
static int sift_test_point(const sift_oct* O, int d, int x, int y, float* xo, float* yo, float* ls) {
    const int w = O->w;
    const float* D0 = O->d[d - 1]; const float* D1 = O->d[d]; const float* D2 = O->d[d + 1];
    const size_t c = (size_t)y * w + x;
    const float v = D1[c];
    const float g = O->g[d][c];                                       /* smoothed intensity (darkness adaptivity) */
    float da = 2.0f * g + 0.1f;
    if (da > 1.0f) da = 1.0f;
    const float T = (0.02f / 3.0f) * da;
    if (!(fabsf(v) > 0.8f * T)) return 0;
    int mx = 1, mn = 1;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const size_t p = (size_t)(y + dy) * w + (x + dx);
            if (dx || dy) { if (!(v > D1[p])) mx = 0; if (!(v < D1[p])) mn = 0; }
            if (!(v > D0[p])) mx = 0;
            if (!(v < D0[p])) mn = 0;
            if (!(v > D2[p])) mx = 0;
            if (!(v < D2[p])) mn = 0;
        }
    if (!mx && !mn) return 0;
    const float dxx = (D1[c + 1] + D1[c - 1]) - 2.0f * v;
    const float dyy = (D1[c + w] + D1[c - w]) - 2.0f * v;
    const float dxy = 0.25f * ((D1[c + w + 1] - D1[c + w - 1]) - (D1[c - w + 1] - D1[c - w - 1]));
    const float det2 = dxx * dyy - dxy * dxy, tr = dxx + dyy;
    if (!(det2 > 0.f)) return 0;
    if (!(tr * tr * 10.0f < 121.0f * det2)) return 0;                /* (r+1)^2 / r with r = 10 */
    const float gx = 0.5f * (D1[c + 1] - D1[c - 1]);
    const float gy = 0.5f * (D1[c + w] - D1[c - w]);
    const float gs = 0.5f * (D2[c] - D0[c]);
    const float dss = (D2[c] + D0[c]) - 2.0f * v;
    const float dxs = 0.25f * ((D2[c + 1] - D2[c - 1]) - (D0[c + 1] - D0[c - 1]));
    const float dys = 0.25f * ((D2[c + w] - D2[c - w]) - (D0[c + w] - D0[c - w]));
    /* H^-1 by the adjugate (H symmetric) */
    const float a00 = dyy * dss - dys * dys;
    const float a01 = dxs * dys - dxy * dss;
    const float a02 = dxy * dys - dxs * dyy;
    const float a11 = dxx * dss - dxs * dxs;
    const float a12 = dxy * dxs - dxx * dys;
    const float a22 = dxx * dyy - dxy * dxy;
    const float det3 = (dxx * a00 + dxy * a01) + dxs * a02;
    if (det3 == 0.f) return 0;
    const float ox = -(((a00 * gx + a01 * gy) + a02 * gs) / det3);
    const float oy = -(((a01 * gx + a11 * gy) + a12 * gs) / det3);
    const float os = -(((a02 * gx + a12 * gy) + a22 * gs) / det3);
    if (!(fabsf(ox) < 1.0f && fabsf(oy) < 1.0f && fabsf(os) < 1.0f)) return 0;
    const float vr = v + 0.5f * ((gx * ox + gy * oy) + gs * os);
    if (!(fabsf(vr) > T)) return 0;
    *xo = (float)x + ox; *yo = (float)y + oy; *ls = (float)(d - 1) + os;
    return 1;
}

/* 36-bin orientation histogram around (xo, yo) on Gaussian image G; returns up to 2 orientations, strongest first */
static int sift_orientations(const float* G, int w, int h, float xo, float yo, float sigma, float* ori) {
    const float sw = 1.5f * sigma;
    const int R = (int)(2.0f * sw + 0.5f);
    const float inv2s2 = -1.0f / (2.0f * sw * sw);
    const int xi = (int)rintf(xo), yi = (int)rintf(yo);
    unsigned hist[36];
    memset(hist, 0, sizeof(hist));
    const float r2max = (float)(R * R) + 0.5f;
    for (int j = -R; j <= R; ++j) {
        const int py = yi + j;
        if (py < 1 || py > h - 2) continue;
        for (int i = -R; i <= R; ++i) {
            const int px = xi + i;
            if (px < 1 || px > w - 2) continue;
            const float dx = (float)px - xo, dy = (float)py - yo;
            const float r2 = dx * dx + dy * dy;
            if (r2 > r2max) continue;
            const size_t c = (size_t)py * w + px;
            const float gx = G[c + 1] - G[c - 1], gy = G[c + w] - G[c - w];
            const float mag = sqrtf(gx * gx + gy * gy);
            const float ang = sift_atan2(gy, gx);
            const float wgt = sift_exp(r2 * inv2s2);
            int b = (int)(ang * (36.0f / SIFT_2PI));
            if (b > 35) b = 35;
            hist[b] += (unsigned)rintf(mag * wgt * SIFT_QSCALE);
        }
    }
    float hf[36], hs[36];
    for (int b = 0; b < 36; ++b) hf[b] = (float)hist[b];
    float maxv = 0.f;
    for (int b = 0; b < 36; ++b) {
        hs[b] = ((hf[(b + 34) % 36] + hf[(b + 2) % 36]) * 0.0625f + (hf[(b + 35) % 36] + hf[(b + 1) % 36]) * 0.25f) + hf[b] * 0.375f;
        if (hs[b] > maxv) maxv = hs[b];
    }
    if (!(maxv > 0.f)) return 0;
    const float th = 0.8f * maxv;
    float bv[2] = {0.f, 0.f}, bo[2] = {0.f, 0.f};
    int n = 0;
    for (int b = 0; b < 36; ++b) {
        const float l = hs[(b + 35) % 36], r = hs[(b + 1) % 36], c = hs[b];
        if (c > l && c > r && c >= th) {
            /* bin b covers [10 b, 10 b + 10) degrees: the peak is referred to the bin CENTRE (sift++: th = 2 pi (i + di + 0.5) / nbins) */
            float bin = ((float)b + 0.5f) + 0.5f * (l - r) / ((l - 2.0f * c) + r);
            if (bin < 0.f) bin = bin + 36.0f;
            if (bin >= 36.0f) bin = bin - 36.0f;
            const float o = bin * (SIFT_2PI / 36.0f);
            /* keep the two largest peaks; ties keep the lower bin first */
            if (n < 2) {
                if (n == 1 && c > bv[0]) { bv[1] = bv[0]; bo[1] = bo[0]; bv[0] = c; bo[0] = o; }
                else { bv[n] = c; bo[n] = o; }
                ++n;
            } else if (c > bv[0]) { bv[1] = bv[0]; bo[1] = bo[0]; bv[0] = c; bo[0] = o; }
            else if (c > bv[1]) { bv[1] = c; bo[1] = o; }
        }
    }
    for (int i = 0; i < n; ++i) ori[i] = bo[i];
    return n;
}

/* 4x4x8 descriptor on Gaussian image G (octave grid) */
static void sift_descriptor(const float* G, int w, int h, float xo, float yo, float sigma, float ori, float* out) {
    unsigned acc[128];
    memset(acc, 0, sizeof(acc));
    float sn, cs;
    sift_sincos(ori, &sn, &cs);
    const float hw = 3.0f * sigma;                                    /* spatial bin width in pixels */
    const int R = (int)rintf(hw * 0x1.6a09e6p+0f * 2.5f);            /* hw * sqrt(2) * (d + 1) / 2 */
    const float cw = cs / hw, sw_ = sn / hw;
    const int xi = (int)rintf(xo), yi = (int)rintf(yo);
    for (int j = -R; j <= R; ++j) {
        const int py = yi + j;
        if (py < 1 || py > h - 2) continue;
        for (int i = -R; i <= R; ++i) {
            const int px = xi + i;
            if (px < 1 || px > w - 2) continue;
            const float dx = (float)px - xo, dy = (float)py - yo;
            const float cr = dx * cw + dy * sw_;                      /* rotated, in bin units */
            const float rr = dy * cw - dx * sw_;
            const float cb = cr + 1.5f, rb = rr + 1.5f;               /* + d/2 - 0.5 */
            if (!(cb > -1.0f && cb < 4.0f && rb > -1.0f && rb < 4.0f)) continue;
            const size_t c = (size_t)py * w + px;
            const float gx = G[c + 1] - G[c - 1], gy = G[c + w] - G[c - w];
            const float mag = sqrtf(gx * gx + gy * gy);
            float ang = sift_atan2(gy, gx) - ori;
            if (ang < 0.f) ang = ang + SIFT_2PI;
            const float ob = ang * (8.0f / SIFT_2PI);
            const float wgt = sift_exp((cr * cr + rr * rr) * -0.125f);   /* sigma = d/2 bins */
            const float m = mag * wgt;
            const float c0f = floorf(cb), r0f = floorf(rb), o0f = floorf(ob);
            const float fc = cb - c0f, fr = rb - r0f, fo = ob - o0f;
            const int c0 = (int)c0f, r0 = (int)r0f, o0 = (int)o0f;
            for (int dr = 0; dr < 2; ++dr) {
                const int r_ = r0 + dr;
                if (r_ < 0 || r_ > 3) continue;
                const float wr = m * (dr ? fr : 1.0f - fr);
                for (int dc = 0; dc < 2; ++dc) {
                    const int c_ = c0 + dc;
                    if (c_ < 0 || c_ > 3) continue;
                    const float wc = wr * (dc ? fc : 1.0f - fc);
                    for (int dq = 0; dq < 2; ++dq) {
                        const int o_ = (o0 + dq) & 7;
                        const float wo = wc * (dq ? fo : 1.0f - fo);
                        acc[(r_ * 4 + c_) * 8 + o_] += (unsigned)rintf(wo * SIFT_QSCALE);
                    }
                }
            }
        }
    }
    /* normalise -> clamp 0.2 -> renormalise, all reductions in integers */
    unsigned long long s1 = 0;
    for (int k = 0; k < 128; ++k) { const unsigned long long u = acc[k] >> 6; s1 += u * u; }
    if (s1 == 0) { for (int k = 0; k < 128; ++k) out[k] = 0.f; return; }
    const float n1 = sqrtf((float)s1);
    unsigned q[128];
    unsigned long long s2 = 0;
    for (int k = 0; k < 128; ++k) {
        float v = (float)(acc[k] >> 6) / n1;
        if (v > 0.2f) v = 0.2f;
        q[k] = (unsigned)rintf(v * SIFT_QSCALE);
        s2 += (unsigned long long)q[k] * q[k];
    }
    const float n2 = sqrtf((float)s2);
    for (int k = 0; k < 128; ++k) out[k] = (float)q[k] / n2;
}

/* SiftGPU-equivalent feature list (x, y, s, o), order = octave asc, level asc, raster, orientation rank; after the
 * -tc2 soft limit.  feats sized by the caller; returns the count (or -needed). */
static int sift_detect(const sift_pyr* P, int nfeatures, sift_feat** out_feats) {
    int cap = 4096, n = 0;
    sift_feat* F = (sift_feat*)malloc(sizeof(sift_feat) * cap);
    int lvl_begin[SIFT_MAX_OCT * SIFT_S + 1];
    for (int o = 0; o < P->no; ++o) {
        const sift_oct* O = &P->o[o];
        for (int l = 0; l < SIFT_S; ++l) {
            lvl_begin[o * SIFT_S + l] = n;
            const int d = l + 1;
            for (int y = SIFT_BORDER; y < O->h - SIFT_BORDER; ++y)
                for (int x = SIFT_BORDER; x < O->w - SIFT_BORDER; ++x) {
                    float xo, yo, ls;
                    if (!sift_test_point(O, d, x, y, &xo, &yo, &ls)) continue;
                    const float sigma = (1.6f * 0x1.428a3p+0f) * sift_exp2(ls * (1.0f / 3.0f));   /* sigma0 * 2^(ls/3) */
                    int gi = (int)rintf(ls) + 1;                       /* nearest Gaussian level */
                    gi = clampi(gi, 1, 3);
                    float ori[2];
                    const int no = sift_orientations(O->g[gi], O->w, O->h, xo, yo, sigma, ori);
                    for (int k = 0; k < no; ++k) {
                        if (n == cap) { cap *= 2; F = (sift_feat*)realloc(F, sizeof(sift_feat) * cap); }
                        sift_feat f; f.x = xo; f.y = yo; f.ls = ls; f.oct = o; f.lvl = l; f.sigma = sigma; f.ori = ori[k];
                        F[n++] = f;
                    }
                }
        }
    }
    const int nl = P->no * SIFT_S;
    lvl_begin[nl] = n;
    /* -tc2: visit levels coarsest first; once the running count exceeds the limit every finer level is dropped */
    int keep_from = 0, run = 0;
    for (int i = nl - 1; i >= 0; --i) {
        if (run > nfeatures) { keep_from = i + 1; break; }
        run += lvl_begin[i + 1] - lvl_begin[i];
    }
    const int first = lvl_begin[keep_from];
    if (first > 0) memmove(F, F + first, sizeof(sift_feat) * (size_t)(n - first));
    *out_feats = F;
    return n - first;
}

/* raw detector tap: xyso[4*i] = (x, y, s, o) in image coordinates, desc[128*i]; returns n (or -needed if cap is small) */
int orc_sift_detect(const uint8_t* gray, int w, int h, int stride, int nfeatures, float* xyso, float* desc, int cap) {
    sift_pyr P;
    if (pyr_build(&P, gray, w, h, stride)) return -1;
    sift_feat* F = NULL;
    const int n = sift_detect(&P, nfeatures, &F);
    if (n > cap) { free(F); pyr_free(&P); return -n; }
    for (int i = 0; i < n; ++i) {
        const float sc = (float)(1 << F[i].oct);
        xyso[4 * i] = F[i].x * sc; xyso[4 * i + 1] = F[i].y * sc; xyso[4 * i + 2] = F[i].sigma * sc; xyso[4 * i + 3] = F[i].ori;
        if (desc) {
            const sift_oct* O = &P.o[F[i].oct];
            const int gi = clampi((int)rintf(F[i].ls) + 1, 1, 3);
            sift_descriptor(O->g[gi], O->w, O->h, F[i].x, F[i].y, F[i].sigma, F[i].ori, desc + (size_t)128 * i);
        }
    }
    free(F); pyr_free(&P);
    return n;
}

/* reference octave of a SiftGPU scale: int(log2(s / 1.6454)) (src/Feature_sift128.cpp:92), via exact thresholds */
int orc_sift_ref_octave(float s) {
    const double r = (double)s / 1.6454;
    int o = 0;
    double p = 2.0;
    while (r >= p && o < 30) { ++o; p *= 2.0; }
    return o;
}

/* full FeatureExtractor_sift128::operator() (src/Feature_sift128.cpp:64-118 + base class post-processing) */
int orc_sift128_extract(const uint8_t* gray, int w, int h, int stride, int nfeatures, int nlevels, float scale_factor,
                        orc_keypoint* kps, float* desc, float* kpsize, int cap, int* n_out, int* n_detected) {
    if (nlevels < 1 || nlevels > ORC_MAX_LEVELS) return -1;
    sift_pyr P;
    if (pyr_build(&P, gray, w, h, stride)) return -1;
    sift_feat* F = NULL;
    const int n = sift_detect(&P, nfeatures, &F);
    if (n_detected) *n_detected = n;
    int q_ext[ORC_MAX_LEVELS];
    orc_features_per_level(nfeatures, nlevels, scale_factor, q_ext);
    const float maxSize0 = powf(1.2f, (float)(8 - 1.0)), maxSize = maxSize0, minSize = 1.0f;
    float* fx = (float*)malloc(sizeof(float) * (n + 1)); float* fy = (float*)malloc(sizeof(float) * (n + 1));
    float* fs = (float*)malloc(sizeof(float) * (n + 1)); float* rs = (float*)malloc(sizeof(float) * (n + 1));
    int* oc = (int*)malloc(sizeof(int) * (n + 1)); int* idx = (int*)malloc(sizeof(int) * (n + 1));
    int* keep = (int*)malloc(sizeof(int) * (n + 1));
    float* lx = (float*)malloc(sizeof(float) * (n + 1)); float* ly = (float*)malloc(sizeof(float) * (n + 1));
    for (int i = 0; i < n; ++i) {
        const float sc = (float)(1 << F[i].oct);
        fx[i] = F[i].x * sc; fy[i] = F[i].y * sc; fs[i] = F[i].sigma * sc; rs[i] = 1.0f;
        oc[i] = orc_sift_ref_octave(fs[i]);
        if (oc[i] > nlevels - 1) oc[i] = nlevels - 1;                /* the reference would index past mnFeaturesPerLevel */
    }
    int m = 0, rc = 0;
    for (int l = 0; l < nlevels && rc == 0; ++l) {                   /* std::map order = ascending octave */
        int nl = 0;
        for (int i = 0; i < n; ++i) if (oc[i] == l) { lx[nl] = fx[i]; ly[nl] = fy[i]; idx[nl] = i; ++nl; }
        if (!nl) continue;
        const int nk = orc_distribute_octree(lx, ly, rs, NULL, nl, 0, w, 0, h, q_ext[l], keep, nl);
        for (int j = 0; j < nk; ++j) {
            const int i = idx[keep[j]];
            if (m >= cap) { rc = -2; break; }
            orc_keypoint* kp = &kps[m];
            kp->x = fx[i]; kp->y = fy[i]; kp->size = fs[i]; kp->angle = F[i].ori; kp->response = 1.0f;
            kp->octave = l; kp->class_id = i;
            const sift_oct* O = &P.o[F[i].oct];
            const int gi = clampi((int)rintf(F[i].ls) + 1, 1, 3);
            sift_descriptor(O->g[gi], O->w, O->h, F[i].x, F[i].y, F[i].sigma, F[i].ori, desc + (size_t)128 * m);
            if (kpsize) {
                const float s = powf(scale_factor, (float)l);
                float sn = maxSize;
                if (maxSize > minSize) sn = 1.0f + (s - minSize) * (maxSize0 - 1.0f) / (maxSize - minSize);
                kpsize[m] = sn;
            }
            ++m;
        }
    }
    free(fx); free(fy); free(fs); free(rs); free(oc); free(idx); free(keep); free(lx); free(ly); free(F);
    pyr_free(&P);
    if (n_out) *n_out = m;
    return rc;
}

/* ---------------------------------------------------------------- batch driver for the CPU arm (bench.py) ------ */
/* One bench step of the sift128 workload (BASELINE configs[2]) on host threads: extraction of B frames, then
 * SearchForInitialization with the L2 distance (DescriptorDistance_sift128) for P frame pairs. */
#include <pthread.h>
#include <malloc.h>
typedef struct {
    const uint8_t* frames; int B, w, h, nfeatures, nlevels; float scale_factor;
    const int *pair_a, *pair_b; int P, window; float th_low, nnratio; int check_ori;
    int cap; orc_keypoint* kps; float* desc; float* ksz; int* n;
    int next_frame, next_pair, err; long total; int phase;
} sift_batch_job;

static void* sift_batch_worker(void* arg) {
    sift_batch_job* J = (sift_batch_job*)arg;
    const float max_size = powf(1.2f, 7.0f);
    if (J->phase == 0) {
        for (;;) {
            const int b = __atomic_fetch_add(&J->next_frame, 1, __ATOMIC_RELAXED);
            if (b >= J->B) break;
            int rc = orc_sift128_extract(J->frames + (size_t)b * J->w * J->h, J->w, J->h, J->w, J->nfeatures, J->nlevels, J->scale_factor,
                                         J->kps + (size_t)b * J->cap, J->desc + (size_t)b * J->cap * 128, J->ksz + (size_t)b * J->cap,
                                         J->cap, &J->n[b], NULL);
            if (rc) __atomic_store_n(&J->err, 1, __ATOMIC_RELAXED);
        }
    } else {
        long local = 0;
        for (;;) {
            const int p = __atomic_fetch_add(&J->next_pair, 1, __ATOMIC_RELAXED);
            if (p >= J->P) break;
            const int a = J->pair_a[p], b = J->pair_b[p];
            const int na = J->n[a];
            float* prev = (float*)malloc(sizeof(float) * 2 * (size_t)(na + 1));
            int* m12 = (int*)malloc(sizeof(int) * (size_t)(na + 1));
            for (int i = 0; i < na; ++i) { prev[2 * i] = J->kps[(size_t)a * J->cap + i].x; prev[2 * i + 1] = J->kps[(size_t)a * J->cap + i].y; }
            local += orc_search_for_initialization(5, J->kps + (size_t)a * J->cap, J->desc + (size_t)a * J->cap * 128, na,
                                                   J->kps + (size_t)b * J->cap, J->desc + (size_t)b * J->cap * 128, J->ksz + (size_t)b * J->cap, J->n[b],
                                                   0.0f, 0.0f, (float)J->w, (float)J->h, max_size, prev, J->window, J->th_low, J->nnratio, J->check_ori, m12);
            free(prev); free(m12);
        }
        __atomic_fetch_add(&J->total, local, __ATOMIC_RELAXED);
    }
    return NULL;
}

long orc_sift128_extract_match_batch(const uint8_t* frames, int B, int w, int h, int nfeatures, int nlevels, float scale_factor,
                                     const int* pair_a, const int* pair_b, int P, int window, float th_low, float nnratio,
                                     int check_ori, int nthreads) {
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    sift_batch_job J;
    memset(&J, 0, sizeof(J));
    J.frames = frames; J.B = B; J.w = w; J.h = h; J.nfeatures = nfeatures; J.nlevels = nlevels; J.scale_factor = scale_factor;
    J.pair_a = pair_a; J.pair_b = pair_b; J.P = P; J.window = window; J.th_low = th_low; J.nnratio = nnratio; J.check_ori = check_ori;
    J.cap = nfeatures + 3 * nlevels + 64;
    J.kps = (orc_keypoint*)malloc(sizeof(orc_keypoint) * (size_t)B * J.cap);
    J.desc = (float*)malloc(sizeof(float) * (size_t)B * J.cap * 128);
    J.ksz = (float*)malloc(sizeof(float) * (size_t)B * J.cap);
    J.n = (int*)calloc(B > 0 ? B : 1, sizeof(int));
    if (nthreads < 1) nthreads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    for (int phase = 0; phase < 2; ++phase) {
        J.phase = phase;
        if (nthreads == 1) { sift_batch_worker(&J); continue; }
        int started = 0;
        for (int t = 0; t < nthreads; ++t) if (pthread_create(&th[t], NULL, sift_batch_worker, &J) == 0) ++started; else break;
        if (started == 0) sift_batch_worker(&J);
        for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    }
    const long total = J.err ? -1 : J.total;
    free(th); free(J.kps); free(J.desc); free(J.ksz); free(J.n);
    return total;
}

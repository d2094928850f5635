/*
 * afv_oracle_orbslam2.c -- CPU ORACLE (test infrastructure, NOT product code): the reference's vanilla ORB-SLAM2 extractor
 * (SURVEY 8f-4), i.e. FeatureExtractor::operator()(..., vanillaOrbslam) of src/ORBextractor.cc:460-676 as built with
 * VANILLA_ORB_SLAM2 (include/Definitions.h:8): ComputePyramid (:647-674), ComputeKeyPointsOctTree (:460-556), IC_Angle /
 * computeOrientation (:138-177), GaussianBlur + computeOrbDescriptor (:603-611, include/FeatureExtractor.h:178-217), merge (:613-626),
 * size override (:629-639).
 *
 * The control flow is the reference's own (in-repo) code and is checked against that code compiled from /root/reference
 * (oracle/_ref, tests/test_oracle_orbslam2.py); the OpenCV calls inside it (cv::resize INTER_LINEAR, cv::FAST, cv::GaussianBlur,
 * cv::fastAtan2, copyMakeBorder) are restated from OpenCV's algorithms and PINNED bit for bit to the cv2 4.13.0 binary of the
 * build container (same tests: the reference code calls the real cv2 functions through callbacks).
 */
#include "afv_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const int8_t kPattern[256 * 4] = {
#include "orb_pattern.inc"
};

static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int refl101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

/* cv::resize(src, dst, sz, 0, 0, INTER_LINEAR) for CV_8UC1 (src/ORBextractor.cc:660): OpenCV's 11-bit fixed-point path
 * (INTER_RESIZE_COEF_BITS = 11).  Coefficients: f = (float)((d + 0.5) * scale - 0.5) with scale = 1 / ((double)dsize / ssize),
 * s = floor(f), f -= s; x only: s < 0 -> (0, f = 0), s >= ssize - 1 -> (ssize - 1, f = 0); c0 = cvRound((1.f - f) * 2048),
 * c1 = cvRound(f * 2048) (float products).  Rows: index clamped to [0, ssize - 1], fraction kept.
 * HResizeLinear: D = S[s] * c0 + S[s + 1] * c1 (int); VResizeLinear<uchar>: ((b0 * (D0 >> 4)) >> 16) + ((b1 * (D1 >> 4)) >> 16) + 2) >> 2. */
static void lin_coeffs(int ssize, int dsize, int is_x, int* ofs, int* c0, int* c1) {
    const double inv_scale = (double)dsize / (double)ssize;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        float f = (float)(((double)d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (is_x) {
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
        }
        ofs[d] = s;
        int a0 = cv_round_f((1.f - f) * 2048.f), a1 = cv_round_f(f * 2048.f);
        c0[d] = a0 > 32767 ? 32767 : a0; c1[d] = a1 > 32767 ? 32767 : a1;
    }
}

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
    int* xo = (int*)malloc(sizeof(int) * 3 * (size_t)dw); int* xc0 = xo + dw; int* xc1 = xc0 + dw;
    int* yo = (int*)malloc(sizeof(int) * 3 * (size_t)dh); int* yc0 = yo + dh; int* yc1 = yc0 + dh;
    lin_coeffs(sw, dw, 1, xo, xc0, xc1);
    lin_coeffs(sh, dh, 0, yo, yc0, yc1);
    for (int y = 0; y < dh; ++y) {
        int s0 = yo[y], s1 = yo[y] + 1;
        s0 = s0 < 0 ? 0 : (s0 > sh - 1 ? sh - 1 : s0);
        s1 = s1 < 0 ? 0 : (s1 > sh - 1 ? sh - 1 : s1);
        const uint8_t* r0 = src + (long)s0 * sstride;
        const uint8_t* r1 = src + (long)s1 * sstride;
        for (int x = 0; x < dw; ++x) {
            const int sx = xo[x], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;       /* c1 = 0 where sx + 1 would leave the row */
            const int D0 = r0[sx] * xc0[x] + r0[sx1] * xc1[x];
            const int D1 = r1[sx] * xc0[x] + r1[sx1] * xc1[x];
            const int v = (((yc0[y] * (D0 >> 4)) >> 16) + ((yc1[y] * (D1 >> 4)) >> 16) + 2) >> 2;
            dst[(long)y * dstride + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(xo); free(yo);
}

/* cv::GaussianBlur(u8, Size(7,7), 2, 2, BORDER_REFLECT_101) on a standalone image (src/ORBextractor.cc:603-604, the level is
 * clone()d first): OpenCV's fixed-point path.  Kernel = getGaussianKernelFixedPoint_ED (8 fractional bits, error diffusion from
 * the ends, centre = 256 - rest): {18, 34, 48, 56, 48, 34, 18} / 256; horizontal pass in 8.8 (ufixedpoint16), vertical in 16.16
 * (ufixedpoint32), result (v + 2^15) >> 16.  Exact integer arithmetic, no saturation can occur (sum of the kernel = 256). */
static const int kG7fix[7] = {18, 34, 48, 56, 48, 34, 18};
void orc_gaussblur7_fixed_u8(const uint8_t* img, int w, int h, int stride, uint8_t* out, int ostride) {
    int* rows = (int*)malloc(sizeof(int) * (size_t)w * (h + 6));
    for (int yy = -3; yy < h + 3; ++yy) {
        const uint8_t* r = img + (long)refl101(yy, h) * stride;
        int* o = rows + (long)(yy + 3) * w;
        for (int x = 0; x < w; ++x) {
            int s = 0;
            for (int k = 0; k < 7; ++k) s += kG7fix[k] * r[refl101(x - 3 + k, w)];
            o[x] = s;
        }
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int* c = rows + (long)(y + 3) * w + x;
            int s = 0;
            for (int k = -3; k <= 3; ++k) s += kG7fix[k + 3] * c[(long)k * w];
            out[(long)y * ostride + x] = (uint8_t)((s + 32768) >> 16);
        }
    free(rows);
}

/* Per-level geometry of ComputePyramid (src/ORBextractor.cc:647-674): mvScaleFactor by repeated float multiplication
 * (:84-90), mvInvScaleFactor = 1.0f / it (:93-98), size = cvRound((float)cols * inv). */
int orc_orbslam2_geometry(int w, int h, int nlevels, float scale_factor, int* lw, int* lh, float* sf, float* isf) {
    if (nlevels < 1 || nlevels > ORC_MAX_LEVELS) return -1;
    sf[0] = 1.0f;
    for (int l = 1; l < nlevels; ++l) sf[l] = sf[l - 1] * scale_factor;
    for (int l = 0; l < nlevels; ++l) {
        isf[l] = 1.0f / sf[l];
        lw[l] = cv_round_f((float)w * isf[l]);
        lh[l] = cv_round_f((float)h * isf[l]);
    }
    return 0;
}

/* ComputeKeyPointsOctTree, detection half for one level (src/ORBextractor.cc:464-531): W = 30 cells with a 3-pixel FAST margin,
 * FAST(iniThFAST) per cell sub-image (its non-max suppression sees only the cell), FAST(minThFAST) when the cell came back empty.
 * Output in the reference's push order (cell rows, cell columns, raster inside the cell), coordinates relative to
 * (minBorderX, minBorderY) = (16, 16) as the reference hands them to DistributeOctTree.  Returns the count (may exceed cap). */
#define OS2_EDGE 19
int orc_orbslam2_detect_level(const uint8_t* img, int cols, int rows, int stride, int ini_th, int min_th,
                              int* xs, int* ys, int* scores, int cap) {
    const float W = 30;
    const int minBorderX = OS2_EDGE - 3, minBorderY = minBorderX;
    const int maxBorderX = cols - OS2_EDGE + 3, maxBorderY = rows - OS2_EDGE + 3;
    const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols < 1 || nRows < 1) return -1;
    const int wCell = (int)ceilf(width / (float)nCols), hCell = (int)ceilf(height / (float)nRows);
    const int ccap = (wCell + 6) * (hCell + 6);
    int* cx = (int*)malloc(sizeof(int) * 3 * (size_t)ccap); int* cy = cx + ccap; int* cs = cy + ccap;
    int n = 0;
    for (int i = 0; i < nRows; ++i) {
        const float iniY = (float)(minBorderY + i * hCell);
        float maxY = iniY + (float)hCell + 6;
        if (iniY >= (float)(maxBorderY - 3)) continue;
        if (maxY > (float)maxBorderY) maxY = (float)maxBorderY;
        for (int j = 0; j < nCols; ++j) {
            const float iniX = (float)(minBorderX + j * wCell);
            float maxX = iniX + (float)wCell + 6;
            if (iniX >= (float)(maxBorderX - 6)) continue;
            if (maxX > (float)maxBorderX) maxX = (float)maxBorderX;
            const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
            const uint8_t* sub = img + (long)y0 * stride + x0;
            int m = orc_fast9_16_nms(sub, x1 - x0, y1 - y0, stride, ini_th, cx, cy, cs, ccap);
            if (m == 0) m = orc_fast9_16_nms(sub, x1 - x0, y1 - y0, stride, min_th, cx, cy, cs, ccap);
            for (int k = 0; k < m; ++k) {
                if (n < cap) { xs[n] = cx[k] + j * wCell; ys[n] = cy[k] + i * hCell; scores[n] = cs[k]; }
                ++n;
            }
        }
    }
    free(cx);
    return n;
}

/* computeOrbDescriptor (include/FeatureExtractor.h:178-217) on the blurred level: float angle = kpt.angle * factorPI;
 * a = cos(angle), b = sin(angle) resolve to the float overloads (the header sits behind `using namespace std`), sample
 * (cvRound(x*b + y*a), cvRound(x*a - y*b)); no border handling (keypoints are >= 19 px inside). */
void orc_orbslam2_descriptor(const uint8_t* blur, int stride, int cx, int cy, float angle_deg, uint8_t* desc) {
    const float factorPI = (float)(M_PI / 180.f);
    const float angle = angle_deg * factorPI;
    const float a = cosf(angle), b = sinf(angle);
    const uint8_t* center = blur + (long)cy * stride + cx;
    const int8_t* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            int t[2];
            for (int j = 0; j < 2; ++j) {
                const float px = (float)pat[4 * k + 2 * j], py = (float)pat[4 * k + 2 * j + 1];
                t[j] = center[(long)cv_round_f(px * b + py * a) * stride + cv_round_f(px * a - py * b)];
            }
            val |= (t[0] < t[1]) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* FeatureExtractor::operator()(img, keypoints, descriptors, sigma2, inf, size, vanillaOrbslam) (src/ORBextractor.cc:568-645).
 * kpsize = computeSize (src/FeatureExtractor.cpp:132-142) with GetKeypointSize = powf(scaleFactor0, octave) and the settings'
 * maxKeyPtSize / minKeyPtSize as :629-636 leave them once every level has produced a keypoint (steady state of the reference's
 * mutated settings object).  what_tap / tap: optional stage output for the parity tests (0 none). */
int orc_orbslam2_extract(const uint8_t* gray, int w, int h, int stride, int nfeatures, int nlevels, float scale_factor,
                         int ini_th, int min_th, orc_keypoint* kps, uint8_t* desc, float* kpsize, int cap, int* n_out) {
    int lw[ORC_MAX_LEVELS], lh[ORC_MAX_LEVELS], quota[ORC_MAX_LEVELS];
    float sf[ORC_MAX_LEVELS], isf[ORC_MAX_LEVELS];
    if (orc_orbslam2_geometry(w, h, nlevels, scale_factor, lw, lh, sf, isf)) return -1;
    orc_features_per_level(nfeatures, nlevels, scale_factor, quota);                 /* src/ORBextractor.cc:102-113 */
    const float maxSize0 = powf(1.2f, (float)(8 - 1.0));                             /* src/FeatureExtractor.cpp:52-55 */
    float maxSize = maxSize0, minSize = 1.0f;
    for (int l = 0; l < nlevels; ++l) { if (sf[l] > maxSize) maxSize = sf[l]; if (sf[l] < minSize) minSize = sf[l]; }
    int n = 0, rc = 0;
    uint8_t* prev = NULL; int pw = 0, ph = 0;
    for (int l = 0; l < nlevels; ++l) {
        const int W = lw[l], H = lh[l];
        uint8_t* img = (uint8_t*)malloc((size_t)W * H);
        if (l == 0) for (int y = 0; y < H; ++y) memcpy(img + (long)y * W, gray + (long)y * stride, (size_t)W);
        else orc_resize_linear_u8(prev, pw, ph, pw, img, W, H, W);
        if (rc == 0) {
            const int ccap = W * H / 4 + 16;
            int* xs = (int*)malloc(sizeof(int) * 3 * (size_t)ccap); int* ys = xs + ccap; int* sc = ys + ccap;
            int m = orc_orbslam2_detect_level(img, W, H, W, ini_th, min_th, xs, ys, sc, ccap);
            if (m < 0) { rc = -1; m = 0; }
            if (m > ccap) m = ccap;
            float* fx = (float*)malloc(sizeof(float) * 3 * (size_t)(m + 1)); float* fy = fx + m + 1; float* fr = fy + m + 1;
            for (int i = 0; i < m; ++i) { fx[i] = (float)xs[i]; fy[i] = (float)ys[i]; fr[i] = (float)sc[i]; }
            int* keep = (int*)malloc(sizeof(int) * (size_t)(m + 1));
            const int minB = OS2_EDGE - 3;
            const int nk = m ? orc_distribute_octree(fx, fy, fr, NULL, m, minB, W - minB, minB, H - minB, quota[l], keep, m) : 0;
            uint8_t* blr = NULL;
            if (nk > 0) { blr = (uint8_t*)malloc((size_t)W * H); orc_gaussblur7_fixed_u8(img, W, H, W, blr, W); }
            for (int j = 0; j < nk; ++j) {
                const int i = keep[j];
                if (n >= cap) { rc = -2; break; }
                const int x = xs[i] + minB, y = ys[i] + minB;
                orc_keypoint* kp = &kps[n];
                kp->angle = orc_ic_angle(img, W, H, W, x, y);
                orc_orbslam2_descriptor(blr, W, x, y, kp->angle, desc + (long)n * 32);
                kp->x = (float)x; kp->y = (float)y;
                if (l != 0) { kp->x *= sf[l]; kp->y *= sf[l]; }
                kp->size = sf[l]; kp->response = fr[i]; kp->octave = l; kp->class_id = -1;
                if (kpsize) {
                    const float s = powf(scale_factor, (float)l);
                    float sn = maxSize;
                    if (maxSize > minSize) sn = 1.0f + (s - minSize) * (maxSize0 - 1.0f) / (maxSize - minSize);
                    kpsize[n] = sn;
                }
                ++n;
            }
            free(xs); free(fx); free(keep); free(blr);
        }
        free(prev);
        prev = img; pw = W; ph = H;
    }
    free(prev);
    if (n_out) *n_out = n;
    return rc;
}

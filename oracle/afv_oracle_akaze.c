/*
 * afv_oracle_akaze.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) of the akaze61 extraction path.
 *
 * PARITY UNPINNED vs libAKAZE, PINNED to cv2 4.13.0 up to two documented differences.  The reference's akaze61 arithmetic lives in libAKAZE (fontan::akaze, un-pinned; upstream
 * pablofdezalc/akaze; reference call sites src/Feature_akaze61.cpp:7-13 (options), :24-34 (scale space), :36-45
 * (detection), :47-61 (descriptors)), which is NOT vendored under the reference tree.  This file restates the PUBLISHED
 * algorithm (Alcantarilla, Nuevo, Bartoli: "Fast Explicit Diffusion for Accelerated Features in Nonlinear Scale Spaces",
 * BMVC 2013) in libAKAZE's structure with the options the reference sets: omax = numOctaves/4 = 2, nsublevels =
 * numOctaves/2 = 4, dthreshold = detectionTh = 5e-4, everything else at the AKAZEOptions defaults (soffset 1.6,
 * derivative_factor 1.5, sderivatives 1.0, PM_G2 diffusivity, kcontrast percentile 0.7 over 300 bins, MLDB full 486-bit
 * descriptor, 3 channels, pattern size 10), and everything the reference does around it: octave := class_id (evolution
 * level, :63-65), DistributeOctTree per level with quota mnFeaturesPerLevel (src/FeatureExtractor.cpp:276-284), all
 * levels merged before Compute_Descriptors (:51-60), computeSize with powf(scaleFactor0, class_id) (:67-69).
 * cv2.AKAZE (OpenCV's port of libAKAZE by the same author) PINS everything except the duplicate filter and the orientation search:
 * with orc_akaze_set_cv2_filter(1) (OpenCV's restructured cross-level filter, akz_detect_cv below) the oracle reproduces cv2 4.13.0's
 * AKAZE_create(threshold 5e-4, 2 octaves, 4 layers) keypoint list EXACTLY -- same count, same order, positions to 1e-4 px, sizes
 * identical, responses to 3e-7 -- and, where the two orientation searches agree (< 0.001 deg, 43 % of the keypoints), 99 % of the
 * 486-bit descriptors bit for bit (all within 2 bits).  OpenCV quantises the 109 sample angles into 42 slices before sliding the
 * pi/3 window, libAKAZE (restated here) slides it over the exact angles: 11 % of the keypoints get an orientation > 1 deg apart.
 * With OpenCV's orientation search selected as well (orc_akaze_set_cv2_filter(3), akz_orientation_cv below) ALL 1751 orientations
 * agree with cv2 (99.9 % within 0.01 deg, none above 1 deg: the derivative arrays differ in the last bits) and 95.7 % of the
 * descriptors are identical, 99.8 % within 2 bits -- so the sampling and MLDB code is pinned on every keypoint, and the two
 * places where the default mode follows libAKAZE instead of OpenCV are exactly the two selectable functions.
 * The default mode (libAKAZE's sequential Find_Scale_Space_Extrema) finds 99.8 % of cv2's keypoints plus the ones OpenCV's filter
 * drops.  tests/test_oracle_akaze.py holds both checks.
 *
 * Arithmetic contract (what the CUDA path reproduces bit for bit): IEEE float32, round to nearest, NO fused multiply-add
 * except the explicit fmaf() of the Gaussian taps, operation order exactly as written; exp / atan2 / sin / cos are the polynomial forms shared with the sift128 oracle;
 * the orientation window sums are accumulated in 64-bit INTEGERS (responses quantised with rintf(v * 2^32)).
 */
#include "afv_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define AKZ_MAX_LEVELS 16
#define AKZ_PI 3.14159265358979323846f
#define AKZ_2PI 6.28318530717958647692f

static inline int akz_round(float v) { return (int)(v + 0.5f); }            /* libAKAZE fRound (non-negative inputs; floor for +) */
static inline int akz_fround(float v) { return (int)floorf(v + 0.5f); }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int refl101(int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * n - 2 - i; return clampi(i, 0, n - 1); }

/* ---- evolution geometry (libAKAZE Allocate_Memory_Evolution) + FED schedule (fed.cpp) --------------------------- */
typedef struct { int w, h, octave, sublevel, sigma_size; float esigma, etime; int nsteps; float tau[64]; } akz_level;

static int fed_is_prime(int n) {
    if (n < 2) return 0;
    for (int i = 2; i * i <= n; ++i) if (n % i == 0) return 0;
    return 1;
}

static int fed_tau(float T, float tau_max, float* tau, int cap) {
    const int n = (int)(ceilf(sqrtf(3.0f * T / tau_max + 0.25f) - 0.5f - 1.0e-8f) + 0.5f);
    if (n <= 0) return 0;
    if (n > cap) return -1;
    const float scale = 3.0f * T / (tau_max * (float)(n * (n + 1)));
    const float c = 1.0f / (4.0f * (float)n + 2.0f), d = scale * tau_max / 2.0f;
    float tauh[64];
    for (int k = 0; k < n; ++k) {
        const float hc = (float)cos((double)(AKZ_PI * (2.0f * (float)k + 1.0f) * c));
        tauh[k] = d / (hc * hc);
    }
    const int kappa = n / 2;
    int prime = n + 1;
    while (!fed_is_prime(prime)) ++prime;
    for (int k = 0, l = 0; l < n; ++k, ++l) {
        int index;
        while ((index = ((k + 1) * kappa) % prime - 1) >= n) ++k;
        tau[l] = tauh[index];
    }
    return n;
}

int orc_akaze_levels(int w, int h, int omax, int nsub, akz_level* L) {
    int n = 0;
    for (int i = 0; i < omax; ++i) {
        const float rf = 1.0f / (float)(1 << i);
        const int lw = (int)((float)w * rf), lh = (int)((float)h * rf);
        if ((lw < 80 || lh < 40) && i != 0) break;
        for (int j = 0; j < nsub; ++j) {
            akz_level* e = &L[n++];
            e->w = lw; e->h = lh; e->octave = i; e->sublevel = j;
            e->esigma = 1.6f * (float)pow(2.0, (double)((float)j / (float)nsub + (float)i));
            e->sigma_size = akz_round(e->esigma * 1.5f / (float)(1 << i));
            e->etime = 0.5f * (e->esigma * e->esigma);
            e->nsteps = 0;
        }
    }
    for (int i = 1; i < n; ++i) {
        L[i].nsteps = fed_tau(L[i].etime - L[i - 1].etime, 0.25f, L[i].tau, 64);
        if (L[i].nsteps < 0) return -1;
    }
    return n;
}

/* tap for the CUDA side / tests: flat description of the schedule */
int orc_akaze_schedule(int w, int h, int omax, int nsub, int* lw, int* lh, int* octave, int* sigma_size, float* esigma,
                       int* nsteps, float* tau /* [levels][64] */) {
    akz_level L[AKZ_MAX_LEVELS];
    if (omax * nsub > AKZ_MAX_LEVELS) return -1;
    const int n = orc_akaze_levels(w, h, omax, nsub, L);
    for (int i = 0; i < n; ++i) {
        lw[i] = L[i].w; lh[i] = L[i].h; octave[i] = L[i].octave; sigma_size[i] = L[i].sigma_size; esigma[i] = L[i].esigma;
        nsteps[i] = L[i].nsteps;
        memcpy(tau + 64 * i, L[i].tau, sizeof(float) * 64);
    }
    return n;
}

/* ---- image primitives ------------------------------------------------------------------------------------------ */
/* Gaussian taps like cv::getGaussianKernel(ksize, sigma): ksize = ceil(2 (1 + (sigma - 0.8) / 0.3)) made odd */
int orc_akaze_gauss_taps(float sigma, float* taps /* centre outward */) {
    int ks = (int)ceilf(2.0f * (1.0f + (sigma - 0.8f) / 0.3f));
    if ((ks % 2) == 0) ks += 1;
    const int r = ks / 2;
    double wd[32], sum = 0.0;
    for (int j = 0; j <= r; ++j) { wd[j] = exp(-(double)(j * j) / (2.0 * (double)sigma * (double)sigma)); sum += j ? 2.0 * wd[j] : wd[j]; }
    for (int j = 0; j <= r; ++j) taps[j] = (float)(wd[j] / sum);
    return r;
}

/* separable blur, rows then columns, BORDER_REPLICATE; acc = t0*c; acc = fma(tj, l + r, acc) (fused, like the sift128 blur) */
static void gauss_blur(const float* src, float* dst, float* tmp, int w, int h, float sigma) {
    float t[32];
    const int r = orc_akaze_gauss_taps(sigma, t);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float* s = src + (size_t)y * w;
            float acc = t[0] * s[x];
            for (int j = 1; j <= r; ++j) acc = fmaf(t[j], s[clampi(x - j, 0, w - 1)] + s[clampi(x + j, 0, w - 1)], acc);
            tmp[(size_t)y * w + x] = acc;
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float acc = t[0] * tmp[(size_t)y * w + x];
            for (int j = 1; j <= r; ++j)
                acc = fmaf(t[j], tmp[(size_t)clampi(y - j, 0, h - 1) * w + x] + tmp[(size_t)clampi(y + j, 0, h - 1) * w + x], acc);
            dst[(size_t)y * w + x] = acc;
        }
}

/* cv::Scharr(scale 1, BORDER_REFLECT_101): derivative [-1 0 1] along one axis, [3 10 3] across */
static void scharr3(const float* src, float* lx, float* ly, int w, int h) {
    for (int y = 0; y < h; ++y) {
        const float* r0 = src + (size_t)refl101(y - 1, h) * w; const float* r1 = src + (size_t)y * w; const float* r2 = src + (size_t)refl101(y + 1, h) * w;
        for (int x = 0; x < w; ++x) {
            const int xm = refl101(x - 1, w), xp = refl101(x + 1, w);
            lx[(size_t)y * w + x] = 3.0f * ((r0[xp] - r0[xm]) + (r2[xp] - r2[xm])) + 10.0f * (r1[xp] - r1[xm]);
            ly[(size_t)y * w + x] = 3.0f * ((r2[xm] - r0[xm]) + (r2[xp] - r0[xp])) + 10.0f * (r2[x] - r0[x]);
        }
    }
}

/* libAKAZE compute_k_percentile(img, 0.7, gscale 1.0, 300 bins) */
static float k_percentile(const float* img, int w, int h, float* g, float* tmp, float* lx, float* ly) {
    gauss_blur(img, g, tmp, w, h, 1.0f);
    scharr3(g, lx, ly, w, h);
    float hmax = 0.f;
    for (int y = 1; y < h - 1; ++y)
        for (int x = 1; x < w - 1; ++x) {
            const float a = lx[(size_t)y * w + x], b = ly[(size_t)y * w + x];
            const float m = sqrtf(a * a + b * b);
            if (m > hmax) hmax = m;
        }
    int hist[300];
    memset(hist, 0, sizeof(hist));
    int npoints = 0;
    for (int y = 1; y < h - 1; ++y)
        for (int x = 1; x < w - 1; ++x) {
            const float a = lx[(size_t)y * w + x], b = ly[(size_t)y * w + x];
            const float m = sqrtf(a * a + b * b);
            if (m != 0.0f) {
                int nbin = (int)floorf(300.0f * (m / hmax));
                if (nbin == 300) nbin = 299;
                hist[nbin]++; npoints++;
            }
        }
    const int nthreshold = (int)((float)npoints * 0.7f);
    int k = 0, nelements = 0;
    for (k = 0; nelements < nthreshold && k < 300; ++k) nelements += hist[k];
    if (nelements < nthreshold) return 0.03f;
    return hmax * ((float)k / 300.0f);
}

/* Perona-Malik g2 conductivity on the un-normalised Scharr gradient of the smoothed level */
static void pm_g2(const float* lx, const float* ly, float* flow, int n, float k) {
    const float inv_k = 1.0f / (k * k);
    for (int i = 0; i < n; ++i) flow[i] = 1.0f / (1.0f + inv_k * (lx[i] * lx[i] + ly[i] * ly[i]));
}

/* one explicit diffusion step (libAKAZE nld_step_scalar): Ld += 0.5 * tau * div(c grad Ld); one-sided at the borders */
static void nld_step(float* Ld, const float* c, float* step, int w, int h, float tau) {
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const size_t p = (size_t)y * w + x;
            float xpos = 0.f, xneg = 0.f, ypos = 0.f, yneg = 0.f;
            if (x + 1 < w) xpos = (c[p] + c[p + 1]) * (Ld[p + 1] - Ld[p]);
            if (x > 0) xneg = (c[p - 1] + c[p]) * (Ld[p] - Ld[p - 1]);
            if (y + 1 < h) ypos = (c[p] + c[p + w]) * (Ld[p + w] - Ld[p]);
            if (y > 0) yneg = (c[p - w] + c[p]) * (Ld[p] - Ld[p - w]);
            step[p] = (0.5f * tau) * ((xpos - xneg) + (ypos - yneg));
        }
    const size_t n = (size_t)w * h;
    for (size_t p = 0; p < n; ++p) Ld[p] = Ld[p] + step[p];
}

/* libAKAZE compute_scharr_derivatives(src, dst, xorder, yorder, scale): sepFilter2D with a [-1 .. 0 .. 1] derivative
 * kernel and a [norm .. w*norm .. norm] smoothing kernel of size 3 + 2 (scale - 1) (scale == 1: Scharr /32),
 * BORDER_REFLECT_101.  dir = 0: d/dx, 1: d/dy. */
static void scharr_scaled(const float* src, float* dst, float* tmp, int w, int h, int dir, int scale) {
    const float wgt = 10.0f / 3.0f;
    const float norm = scale == 1 ? 3.0f / 16.0f * 0.5f : 1.0f / (2.0f * (float)scale * (wgt + 2.0f));
    const float wc = scale == 1 ? 10.0f / 16.0f * 0.5f : wgt * norm;
    const int s = scale;
    /* rows (x direction) then columns (y direction), like sepFilter2D */
    for (int y = 0; y < h; ++y) {
        const float* r = src + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            const float a = r[refl101(x - s, w)], b = r[refl101(x + s, w)];
            tmp[(size_t)y * w + x] = dir == 0 ? (b - a) : (norm * (a + b) + wc * r[x]);
        }
    }
    for (int y = 0; y < h; ++y) {
        const float* r0 = tmp + (size_t)refl101(y - s, h) * w; const float* r1 = tmp + (size_t)y * w; const float* r2 = tmp + (size_t)refl101(y + s, h) * w;
        for (int x = 0; x < w; ++x) dst[(size_t)y * w + x] = dir == 0 ? (norm * (r0[x] + r2[x]) + wc * r1[x]) : (r2[x] - r0[x]);
    }
}

/* ---- scale space ------------------------------------------------------------------------------------------------- */
typedef struct {
    int nl; akz_level L[AKZ_MAX_LEVELS];
    float* Lt[AKZ_MAX_LEVELS]; float* Lsm[AKZ_MAX_LEVELS];
    float* Lx[AKZ_MAX_LEVELS]; float* Ly[AKZ_MAX_LEVELS]; float* Ldet[AKZ_MAX_LEVELS];
    float kcontrast0;
} akz_space;

static void space_free(akz_space* S) {
    for (int i = 0; i < S->nl; ++i) { free(S->Lt[i]); free(S->Lsm[i]); free(S->Lx[i]); free(S->Ly[i]); free(S->Ldet[i]); }
}

static int space_build(akz_space* S, const uint8_t* gray, int w, int h, int stride, int omax, int nsub) {
    if ((w & 1) || (h & 1)) return -3;                        /* exact 2x2 area half-sampling only */
    memset(S, 0, sizeof(*S));
    S->nl = orc_akaze_levels(w, h, omax, nsub, S->L);
    if (S->nl < 1) return -1;
    const size_t n0 = (size_t)w * h;
    float* img = (float*)malloc(sizeof(float) * n0);
    float* tmp = (float*)malloc(sizeof(float) * n0);
    float* lx = (float*)malloc(sizeof(float) * n0); float* ly = (float*)malloc(sizeof(float) * n0);
    float* flow = (float*)malloc(sizeof(float) * n0); float* aux = (float*)malloc(sizeof(float) * n0);
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) img[(size_t)y * w + x] = (float)gray[(size_t)y * stride + x] * (1.0f / 255.0f);
    for (int i = 0; i < S->nl; ++i) {
        const size_t n = (size_t)S->L[i].w * S->L[i].h;
        S->Lt[i] = (float*)malloc(sizeof(float) * n); S->Lsm[i] = (float*)malloc(sizeof(float) * n);
        S->Lx[i] = (float*)malloc(sizeof(float) * n); S->Ly[i] = (float*)malloc(sizeof(float) * n); S->Ldet[i] = (float*)malloc(sizeof(float) * n);
    }
    gauss_blur(img, S->Lt[0], tmp, w, h, 1.6f);
    memcpy(S->Lsm[0], S->Lt[0], sizeof(float) * n0);
    float kcontrast = k_percentile(img, w, h, aux, tmp, lx, ly);
    S->kcontrast0 = kcontrast;
    for (int i = 1; i < S->nl; ++i) {
        const int lw = S->L[i].w, lh = S->L[i].h;
        const size_t n = (size_t)lw * lh;
        if (S->L[i].octave > S->L[i - 1].octave) {
            const int pw = S->L[i - 1].w;
            const float* P = S->Lt[i - 1];
            for (int y = 0; y < lh; ++y) for (int x = 0; x < lw; ++x) {
                const size_t q = (size_t)(2 * y) * pw + 2 * x;
                S->Lt[i][(size_t)y * lw + x] = ((P[q] + P[q + 1]) + (P[q + pw] + P[q + pw + 1])) * 0.25f;
            }
            kcontrast = kcontrast * 0.75f;
        } else memcpy(S->Lt[i], S->Lt[i - 1], sizeof(float) * n);
        gauss_blur(S->Lt[i], S->Lsm[i], tmp, lw, lh, 1.0f);
        scharr3(S->Lsm[i], lx, ly, lw, lh);
        pm_g2(lx, ly, flow, (int)n, kcontrast);
        for (int j = 0; j < S->L[i].nsteps; ++j) nld_step(S->Lt[i], flow, aux, lw, lh, S->L[i].tau[j]);
    }
    /* multiscale derivatives + determinant of the Hessian (Compute_Multiscale_Derivatives / _Determinant_Hessian_Response) */
    for (int i = 0; i < S->nl; ++i) {
        const int lw = S->L[i].w, lh = S->L[i].h, ss = S->L[i].sigma_size;
        const size_t n = (size_t)lw * lh;
        float* lxx = lx; float* lyy = ly; float* lxy = flow;
        scharr_scaled(S->Lsm[i], S->Lx[i], tmp, lw, lh, 0, ss);
        scharr_scaled(S->Lsm[i], S->Ly[i], tmp, lw, lh, 1, ss);
        scharr_scaled(S->Lx[i], lxx, tmp, lw, lh, 0, ss);
        scharr_scaled(S->Ly[i], lyy, tmp, lw, lh, 1, ss);
        scharr_scaled(S->Lx[i], lxy, tmp, lw, lh, 1, ss);
        const float s1 = (float)ss, s2 = (float)(ss * ss);
        for (size_t p = 0; p < n; ++p) {
            const float a = lxx[p] * s2, b = lyy[p] * s2, c = lxy[p] * s2;
            S->Ldet[i][p] = a * b - c * c;
            S->Lx[i][p] = S->Lx[i][p] * s1; S->Ly[i][p] = S->Ly[i][p] * s1;
        }
    }
    free(img); free(tmp); free(lx); free(ly); free(flow); free(aux);
    return 0;
}

long orc_akaze_scale_space(const uint8_t* gray, int w, int h, int stride, int omax, int nsub, int what, int level, float* out,
                           int* ow, int* oh, float* kcontrast) {
    akz_space S;
    if (space_build(&S, gray, w, h, stride, omax, nsub)) return -1;
    long n = -1;
    if (level >= 0 && level < S.nl) {
        n = (long)S.L[level].w * S.L[level].h;
        const float* src = what == 0 ? S.Lt[level] : what == 1 ? S.Lsm[level] : what == 2 ? S.Lx[level] : what == 3 ? S.Ly[level] : S.Ldet[level];
        if (out) memcpy(out, src, sizeof(float) * (size_t)n);
        if (ow) *ow = S.L[level].w;
        if (oh) *oh = S.L[level].h;
    }
    if (kcontrast) *kcontrast = S.kcontrast0;
    space_free(&S);
    return n;
}

/* ---- detection (Find_Scale_Space_Extrema + Do_Subpixel_Refinement) --------------------------------------------- */
typedef struct { float x, y, size, response; int octave, class_id; } akz_pt;

static int akz_detect(const akz_space* S, float dthreshold, akz_pt** out) {
    int cap = 4096, n = 0;
    akz_pt* A = (akz_pt*)malloc(sizeof(akz_pt) * cap);
    const float smax = 10.0f * 0x1.6a09e6p+0f;                       /* MLDB: 10 sqrt(2) (libAKAZE: SURF and MLDB descriptors 10 sqrt 2, M-SURF 12 sqrt 2; cv2 keeps keypoints from row 29 on at sigma_size 2) */
    for (int i = 0; i < S->nl; ++i) {
        const int w = S->L[i].w, h = S->L[i].h;
        const float* D = S->Ldet[i];
        const float ratio = (float)(1 << S->L[i].octave);
        const float psize = S->L[i].esigma * 1.5f;
        const int sigma_size = akz_round(psize / ratio);
        for (int y = 1; y < h - 1; ++y)
            for (int x = 1; x < w - 1; ++x) {
                const size_t c = (size_t)y * w + x;
                const float v = D[c];
                if (!(v > dthreshold && v >= 0.00001f && v > D[c - 1] && v > D[c + 1] &&
                      v > D[c - w - 1] && v > D[c - w] && v > D[c - w + 1] && v > D[c + w - 1] && v > D[c + w] && v > D[c + w + 1])) continue;
                int is_ext = 1, is_rep = 0, id_rep = 0;
                const float px = (float)x * ratio, py = (float)y * ratio;
                for (int k = 0; k < n; ++k) {
                    if (A[k].class_id == i - 1 || A[k].class_id == i) {
                        const float dx = px - A[k].x, dy = py - A[k].y;
                        if (dx * dx + dy * dy <= psize * psize) {
                            if (v > A[k].response) { id_rep = k; is_rep = 1; } else is_ext = 0;
                            break;
                        }
                    }
                }
                if (!is_ext) continue;
                const int left = akz_fround((float)x - smax * (float)sigma_size) - 1, right = akz_fround((float)x + smax * (float)sigma_size) + 1;
                const int up = akz_fround((float)y - smax * (float)sigma_size) - 1, down = akz_fround((float)y + smax * (float)sigma_size) + 1;
                if (left < 0 || right >= w || up < 0 || down >= h) continue;
                akz_pt p; p.x = px; p.y = py; p.size = psize; p.response = v; p.octave = S->L[i].octave; p.class_id = i;
                if (is_rep) A[id_rep] = p;
                else { if (n == cap) { cap *= 2; A = (akz_pt*)realloc(A, sizeof(akz_pt) * cap); } A[n++] = p; }
            }
    }
    /* filter against the upper level, then sub-pixel refinement */
    akz_pt* K = (akz_pt*)malloc(sizeof(akz_pt) * (n + 1));
    int m = 0;
    for (int i = 0; i < n; ++i) {
        int rep = 0;
        for (int j = i + 1; j < n; ++j)
            if (A[j].class_id == A[i].class_id + 1) {
                const float dx = A[i].x - A[j].x, dy = A[i].y - A[j].y;
                if (dx * dx + dy * dy <= A[i].size * A[i].size) { if (A[i].response < A[j].response) { rep = 1; break; } }
            }
        if (rep) continue;
        akz_pt p = A[i];
        const int lv = p.class_id, w = S->L[lv].w;
        const float ratio = (float)(1 << p.octave);
        const int x = akz_fround(p.x / ratio), y = akz_fround(p.y / ratio);
        const float* D = S->Ldet[lv];
        const size_t c = (size_t)y * w + x;
        const float Dx = 0.5f * (D[c + 1] - D[c - 1]), Dy = 0.5f * (D[c + w] - D[c - w]);
        const float Dxx = (D[c + 1] + D[c - 1]) - 2.0f * D[c], Dyy = (D[c + w] + D[c - w]) - 2.0f * D[c];
        const float Dxy = 0.25f * (D[c + w + 1] + D[c - w - 1]) - 0.25f * (D[c - w + 1] + D[c + w - 1]);
        const float det = Dxx * Dyy - Dxy * Dxy;
        if (det == 0.f) continue;
        const float ox = (Dxy * Dy - Dyy * Dx) / det, oy = (Dxy * Dx - Dxx * Dy) / det;     /* solve [Dxx Dxy; Dxy Dyy] o = -[Dx Dy] */
        if (!(fabsf(ox) <= 1.0f && fabsf(oy) <= 1.0f)) continue;
        p.x = ((float)x + ox) * ratio + 0.5f * (ratio - 1.0f);
        p.y = ((float)y + oy) * ratio + 0.5f * (ratio - 1.0f);
        p.size = p.size * 2.0f;
        K[m++] = p;
    }
    free(A);
    *out = K;
    return m;
}

/* ---- detection, OpenCV's variant (cv::AKAZEFeatures::Find_Scale_Space_Extrema + Do_Subpixel_Refinement of OpenCV 4.x) ---------
 * OpenCV restructured libAKAZE's sequential duplicate filter for parallel execution: (1) per level, a byte mask of the strict 3x3
 * maxima above the threshold inside the descriptor border, where a new maximum first looks for an already marked one within
 * sigma_size (first hit of a square raster scan that passes the radius test) and the weaker of the two goes; (2) for every level i >= 1 in raster order, the FIRST marked point of level
 * i - 1 inside a search window around the projected position (square scan, radius test) is cleared when the level-i response is
 * larger; (3) the same from the top level down against level i + 1; (4) sub-pixel refinement of what is left, level by level in raster
 * order.  This mode exists ONLY to pin the rest of the pipeline (scale space, Hessian responses,
 * sub-pixel fit, orientation, MLDB) to the cv2 4.13.0 binary keypoint by keypoint; the reference links libAKAZE, whose filter is the
 * default mode above. */
static int akz_find_neighbor(const unsigned char* mask, int w, int h, int x, int y, int r, int* idx) {
    for (int i = y - r; i < y + r; ++i) {
        if (i < 0 || i >= h) continue;
        for (int j = x - r; j < x + r; ++j) {
            if (j < 0 || j >= w || !mask[(size_t)i * w + j]) continue;
            if ((i - y) * (i - y) + (j - x) * (j - x) <= r * r) { *idx = i * w + j; return 1; }
        }
    }
    return 0;
}
static int akz_detect_cv(const akz_space* S, float dthreshold, akz_pt** out) {
    const float smax = 10.0f * 0x1.6a09e6p+0f;
    unsigned char* mask[ORC_MAX_LEVELS];
    int total = 0;
    for (int i = 0; i < S->nl; ++i) {
        const int w = S->L[i].w, h = S->L[i].h;
        const float* D = S->Ldet[i];
        mask[i] = (unsigned char*)calloc((size_t)w * h, 1);
        const int border = (int)lrintf(smax * (float)S->L[i].sigma_size) + 1;
        for (int y = border; y < h - border; ++y)
            for (int x = border; x < w - border; ++x) {
                const size_t c = (size_t)y * w + x;
                const float v = D[c];
                if (!(v > dthreshold && v > D[c - 1] && v > D[c + 1] && v > D[c - w - 1] && v > D[c - w] && v > D[c - w + 1] &&
                      v > D[c + w - 1] && v > D[c + w] && v > D[c + w + 1])) continue;
                int idx = 0;                                        /* same scale: the better of two close maxima stays (raster order) */
                if (akz_find_neighbor(mask[i], w, h, x, y, S->L[i].sigma_size, &idx)) {
                    if (v > D[idx]) mask[i][idx] = 0; else continue;
                }
                mask[i][c] = 1;
            }
    }
    for (int i = 1; i < S->nl; ++i) {                               /* against the lower level */
        const int w = S->L[i].w, h = S->L[i].h, wp = S->L[i - 1].w, hp = S->L[i - 1].h;
        const int diff = (1 << S->L[i].octave) / (1 << S->L[i - 1].octave);
        const int r = S->L[i].sigma_size * diff;
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                if (!mask[i][(size_t)y * w + x]) continue;
                int idx = 0;
                if (akz_find_neighbor(mask[i - 1], wp, hp, x * diff, y * diff, r, &idx) && S->Ldet[i][(size_t)y * w + x] > S->Ldet[i - 1][idx])
                    mask[i - 1][idx] = 0;
            }
    }
    for (int i = S->nl - 2; i >= 0; --i) {                          /* against the upper level */
        const int w = S->L[i].w, h = S->L[i].h, wn = S->L[i + 1].w, hn = S->L[i + 1].h;
        const int diff = (1 << S->L[i + 1].octave) / (1 << S->L[i].octave);
        const int r = S->L[i + 1].sigma_size;
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                if (!mask[i][(size_t)y * w + x]) continue;
                int idx = 0;
                if (akz_find_neighbor(mask[i + 1], wn, hn, x / diff, y / diff, r, &idx) && S->Ldet[i][(size_t)y * w + x] > S->Ldet[i + 1][idx])
                    mask[i + 1][idx] = 0;
            }
    }
    for (int i = 0; i < S->nl; ++i) for (size_t c = 0; c < (size_t)S->L[i].w * S->L[i].h; ++c) total += mask[i][c];
    akz_pt* K = (akz_pt*)malloc(sizeof(akz_pt) * (size_t)(total + 1));
    int m = 0;
    for (int i = 0; i < S->nl; ++i) {
        const int w = S->L[i].w, h = S->L[i].h;
        const float* D = S->Ldet[i];
        const float ratio = (float)(1 << S->L[i].octave);
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const size_t c = (size_t)y * w + x;
                if (!mask[i][c]) continue;
                const float Dx = 0.5f * (D[c + 1] - D[c - 1]), Dy = 0.5f * (D[c + w] - D[c - w]);
                const float Dxx = (D[c + 1] + D[c - 1]) - 2.0f * D[c], Dyy = (D[c + w] + D[c - w]) - 2.0f * D[c];
                const float Dxy = 0.25f * (D[c + w + 1] + D[c - w - 1]) - 0.25f * (D[c - w + 1] + D[c + w - 1]);
                const float det = Dxx * Dyy - Dxy * Dxy;
                if (det == 0.f) continue;
                const float ox = (Dxy * Dy - Dyy * Dx) / det, oy = (Dxy * Dx - Dxx * Dy) / det;
                if (!(fabsf(ox) <= 1.0f && fabsf(oy) <= 1.0f)) continue;
                akz_pt p;
                p.x = ((float)x + ox) * ratio + 0.5f * (ratio - 1.0f);
                p.y = ((float)y + oy) * ratio + 0.5f * (ratio - 1.0f);
                p.size = S->L[i].esigma * 1.5f * 2.0f; p.response = D[c]; p.octave = S->L[i].octave; p.class_id = i;
                K[m++] = p;
            }
        free(mask[i]);
    }
    *out = K;
    return m;
}
static int g_akz_cv_mode = 0;
void orc_akaze_set_cv2_filter(int on) { g_akz_cv_mode = on; }      /* test hook, bit 0: OpenCV's duplicate filter, bit 1: OpenCV's orientation search */
static int akz_detect_dispatch(const akz_space* S, float dthreshold, akz_pt** out) {
    return (g_akz_cv_mode & 1) ? akz_detect_cv(S, dthreshold, out) : akz_detect(S, dthreshold, out);
}

/* ---- descriptors (Compute_Main_Orientation + Get_MLDB_Full_Descriptor) ------------------------------------------- */
static inline float akz_exp2(float t) {
    if (t < -126.f) t = -126.f;
    if (t > 126.f) t = 126.f;
    const float n = rintf(t), f = t - n;
    float p = 0x1.430912p-13f;
    p = p * f + 0x1.5d87fep-10f; p = p * f + 0x1.3b2ab6p-7f; p = p * f + 0x1.c6b08ep-5f;
    p = p * f + 0x1.ebfbep-3f; p = p * f + 0x1.62e43p-1f; p = p * f + 1.0f;
    union { unsigned u; float f; } s; s.u = (unsigned)((int)n + 127) << 23;
    return p * s.f;
}
static inline float akz_gauss25(int i, int j) {                       /* sigma 2.5 table of libAKAZE, recomputed */
    return 0x1.a13714p-6f * akz_exp2((float)(i * i + j * j) * (-0.08f * 0x1.715476p+0f));      /* 1 / (2 pi 6.25) */
}
static inline float akz_atan2(float y, float x) {
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
static inline void akz_sincos(float a, float* sn, float* cs) {
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

static float akz_orientation(const akz_space* S, const akz_pt* kp) {
    const int lv = kp->class_id, w = S->L[lv].w, h = S->L[lv].h;
    const float ratio = (float)(1 << kp->octave);
    const int s = akz_fround(0.5f * kp->size / ratio);
    const float xf = kp->x / ratio, yf = kp->y / ratio;
    long long rx[109], ry[109]; float ang[109];
    int idx = 0;
    for (int i = -6; i <= 6; ++i)
        for (int j = -6; j <= 6; ++j) {
            if (i * i + j * j >= 36) continue;
            const int iy = clampi(akz_fround(yf + (float)(j * s)), 0, h - 1), ix = clampi(akz_fround(xf + (float)(i * s)), 0, w - 1);
            const float g = akz_gauss25(i < 0 ? -i : i, j < 0 ? -j : j);
            const float ax = g * S->Lx[lv][(size_t)iy * w + ix], ay = g * S->Ly[lv][(size_t)iy * w + ix];
            rx[idx] = (long long)rintf(ax * 4294967296.0f); ry[idx] = (long long)rintf(ay * 4294967296.0f);
            ang[idx] = akz_atan2(ay, ax);
            ++idx;
        }
    float best = 0.f, angle = 0.f;
    for (int t = 0; t < 42; ++t) {                                    /* ang1 = 0, 0.15, ... < 2 pi */
        const float a1 = (float)t * 0.15f;
        const float a2 = (a1 + AKZ_PI / 3.0f > AKZ_2PI) ? a1 - 5.0f * AKZ_PI / 3.0f : a1 + AKZ_PI / 3.0f;
        long long sx = 0, sy = 0;
        for (int k = 0; k < 109; ++k) {
            const float a = ang[k];
            if ((a1 < a2 && a1 < a && a < a2) || (a2 < a1 && ((a > 0.f && a < a2) || (a > a1 && a < AKZ_2PI)))) { sx += rx[k]; sy += ry[k]; }
        }
        const float fx = (float)sx, fy = (float)sy;
        const float m = fx * fx + fy * fy;
        if (m > best) { best = m; angle = akz_atan2(fy, fx); }
    }
    return angle;
}

/* OpenCV's Compute_Main_Orientation (modules/features2d/src/kaze/AKAZEFeatures.cpp), selected with orc_akaze_set_cv2_filter(mode & 2):
 * the TEST-ONLY variant that pins the sampling and MLDB code below to the cv2 binary on every keypoint, not only where the two
 * orientation searches happen to agree.  Differences from libAKAZE's search restated above: the centre and the scale are rounded
 * first (cvRound) and the 109 samples sit on that integer lattice; the weights are the 8-decimal gauss25 literals; angles come from
 * hal::fastAtan2 (radians) and are counting-sorted into 42 slices of 2 pi / 42; the pi / 3 window is 7 whole slices, summed in
 * float in sorted order (slices ascending, original index descending inside a slice); the result is getAngle(sumX, sumY). */
static float akz_fast_atan2_rad(float y, float x) {                    /* cv::hal::fastAtan32f, angleInDegrees = false */
    const float p1 = 0.9997878412794807f * 57.29577951308232f, p3 = -0.3258083974640975f * 57.29577951308232f;
    const float p5 = 0.1555786518463281f * 57.29577951308232f, p7 = -0.04432655554792128f * 57.29577951308232f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) { c = ay / (ax + 2.220446049250313e-16f); c2 = c * c; a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    else { c = ax / (ay + 2.220446049250313e-16f); c2 = c * c; a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    if (x < 0.f) a = 180.f - a;
    if (y < 0.f) a = 360.f - a;
    return a * 0.017453292519943295f;
}
static float akz_get_angle_cv(float x, float y) {                      /* kaze/utils.h getAngle */
    if (x >= 0 && y >= 0) return atanf(y / x);
    if (x < 0 && y >= 0) return AKZ_PI - atanf(-y / x);
    if (x < 0 && y < 0) return AKZ_PI + atanf(y / x);
    if (x >= 0 && y < 0) return AKZ_2PI - atanf(-y / x);
    return 0.f;
}
static float akz_orientation_cv(const akz_space* S, const akz_pt* kp) {
    const int lv = kp->class_id, w = S->L[lv].w, h = S->L[lv].h;
    const float ratio = (float)(1 << kp->octave);
    const int scale = (int)lrintf(0.5f * kp->size / ratio);
    const int x0 = (int)lrintf(kp->x / ratio), y0 = (int)lrintf(kp->y / ratio);
    static float g25[7][7];
    static int g25_ok = 0;
    if (!g25_ok) {                                                     /* the table's literals: the closed form rounded to 8 decimals */
        for (int i = 0; i < 7; ++i)
            for (int j = 0; j < 7; ++j) {
                const double v = exp(-(double)(i * i + j * j) / 12.5) / (2.0 * 3.14159265358979323846 * 6.25);
                g25[i][j] = (float)(floor(v * 1e8 + 0.5) / 1e8);
            }
        g25_ok = 1;
    }
    float rx[109], ry[109], ang[109];
    int n = 0;
    for (int i = -6; i <= 6; ++i)
        for (int j = -6; j <= 6; ++j) {
            if (i * i + j * j >= 36) continue;
            const int y = clampi(y0 + j * scale, 0, h - 1), x = clampi(x0 + i * scale, 0, w - 1);
            const float g = g25[i < 0 ? -i : i][j < 0 ? -j : j];
            rx[n] = g * S->Lx[lv][(size_t)y * w + x]; ry[n] = g * S->Ly[lv][(size_t)y * w + x];
            ang[n] = akz_fast_atan2_rad(ry[n], rx[n]);
            ++n;
        }
    enum { SL = 42, WIN = 7 };
    const float step = (float)(2.0 * 3.14159265358979323846 / SL);
    int cum[SL + 1], sorted[109];
    memset(cum, 0, sizeof(cum));
    for (int i = 0; i < n; ++i) { int b = (int)(ang[i] / step); if (b < 0 || b >= SL) b = 0; cum[b]++; }
    for (int i = 1; i <= SL; ++i) cum[i] += cum[i - 1];
    for (int i = 0; i < n; ++i) { int b = (int)(ang[i] / step); if (b < 0 || b >= SL) b = 0; sorted[--cum[b]] = i; }
    float maxX = 0.f, maxY = 0.f;
    for (int i = cum[0]; i < cum[WIN]; ++i) { maxX += rx[sorted[i]]; maxY += ry[sorted[i]]; }
    float maxNorm = maxX * maxX + maxY * maxY;
    for (int sn = 1; sn <= SL - WIN; ++sn) {
        if (cum[sn] == cum[sn - 1] && cum[sn + WIN] == cum[sn + WIN - 1]) continue;
        float sx = 0.f, sy = 0.f;
        for (int i = cum[sn]; i < cum[sn + WIN]; ++i) { sx += rx[sorted[i]]; sy += ry[sorted[i]]; }
        const float nm = sx * sx + sy * sy;
        if (nm > maxNorm) { maxNorm = nm; maxX = sx; maxY = sy; }
    }
    for (int sn = SL - WIN + 1; sn < SL; ++sn) {
        const int remain = sn + WIN - SL;
        if (cum[sn] == cum[sn - 1] && cum[remain] == cum[remain - 1]) continue;
        float sx = 0.f, sy = 0.f;
        for (int i = cum[sn]; i < cum[SL]; ++i) { sx += rx[sorted[i]]; sy += ry[sorted[i]]; }
        for (int i = cum[0]; i < cum[remain]; ++i) { sx += rx[sorted[i]]; sy += ry[sorted[i]]; }
        const float nm = sx * sx + sy * sy;
        if (nm > maxNorm) { maxNorm = nm; maxX = sx; maxY = sy; }
    }
    return akz_get_angle_cv(maxX, maxY);
}

static void akz_mldb(const akz_space* S, const akz_pt* kp, float angle, uint8_t* desc) {
    memset(desc, 0, 61);
    const int lv = kp->class_id, w = S->L[lv].w, h = S->L[lv].h;
    const float ratio = (float)(1 << kp->octave);
    const int scale = akz_fround(0.5f * kp->size / ratio);
    const float xf = kp->x / ratio, yf = kp->y / ratio;
    float si, co;
    akz_sincos(angle, &si, &co);
    const float cs = co * (float)scale, ss = si * (float)scale;
    static const int steps[3] = {10, 7, 5};
    int dpos = 0;
    for (int g = 0; g < 3; ++g) {
        const int step = steps[g], cnt = (g + 2) * (g + 2);
        float val[16][3];
        int vp = 0;
        for (int i = -10; i < 10; i += step)
            for (int j = -10; j < 10; j += step) {
                float di = 0.f, dx = 0.f, dy = 0.f;
                int ns = 0;
                for (int k = i; k < i + step; ++k)
                    for (int l = j; l < j + step; ++l) {
                        const float sy = yf + ((float)l * cs + (float)k * ss);
                        const float sx = xf + ((float)k * cs - (float)l * ss);
                        const int y1 = clampi(akz_fround(sy), 0, h - 1), x1 = clampi(akz_fround(sx), 0, w - 1);
                        const size_t p = (size_t)y1 * w + x1;
                        const float rx = S->Lx[lv][p], ry = S->Ly[lv][p];
                        di = di + S->Lt[lv][p];
                        dx = dx + (ry * co - rx * si);
                        dy = dy + (rx * co + ry * si);
                        ++ns;
                    }
                val[vp][0] = di / (float)ns; val[vp][1] = dx / (float)ns; val[vp][2] = dy / (float)ns;
                ++vp;
            }
        for (int ch = 0; ch < 3; ++ch)
            for (int a = 0; a < cnt; ++a)
                for (int b = a + 1; b < cnt; ++b) {
                    if (val[a][ch] > val[b][ch]) desc[dpos >> 3] |= (uint8_t)(1u << (dpos & 7));
                    ++dpos;
                }
    }
}

/* raw detector tap (Feature_Detection): x, y, size, response, class_id(as float) per keypoint, list order */
int orc_akaze_detect(const uint8_t* gray, int w, int h, int stride, int omax, int nsub, float dth, float* out5, int cap) {
    akz_space S;
    int rc = space_build(&S, gray, w, h, stride, omax, nsub);
    if (rc) return rc;
    akz_pt* K = NULL;
    const int n = akz_detect_dispatch(&S, dth, &K);
    if (n <= cap) for (int i = 0; i < n; ++i) { out5[5 * i] = K[i].x; out5[5 * i + 1] = K[i].y; out5[5 * i + 2] = K[i].size; out5[5 * i + 3] = K[i].response; out5[5 * i + 4] = (float)K[i].class_id; }
    free(K); space_free(&S);
    return n <= cap ? n : -n;
}

/* full FeatureExtractor_akaze61::operator() */
int orc_akaze61_extract(const uint8_t* gray, int w, int h, int stride, int nfeatures, int nlevels, float scale_factor,
                        float detect_th, orc_keypoint* kps, uint8_t* desc, float* kpsize, int cap, int* n_out, int* n_detected) {
    if (nlevels < 1 || nlevels > ORC_MAX_LEVELS) return -1;
    akz_space S;
    int rc = space_build(&S, gray, w, h, stride, nlevels / 4, nlevels / 2);
    if (rc) return rc;
    akz_pt* K = NULL;
    const int n = akz_detect_dispatch(&S, detect_th, &K);
    if (n_detected) *n_detected = n;
    int q_ext[ORC_MAX_LEVELS];
    orc_features_per_level(nfeatures, nlevels, scale_factor, q_ext);
    const float maxSize0 = powf(1.2f, (float)(8 - 1.0)), maxSize = maxSize0, minSize = 1.0f;
    float* lx = (float*)malloc(sizeof(float) * (n + 1)); float* ly = (float*)malloc(sizeof(float) * (n + 1));
    float* lr = (float*)malloc(sizeof(float) * (n + 1)); int* idx = (int*)malloc(sizeof(int) * (n + 1)); int* keep = (int*)malloc(sizeof(int) * (n + 1));
    int m = 0;
    for (int l = 0; l < nlevels && rc == 0; ++l) {
        int nl = 0;
        for (int i = 0; i < n; ++i) if (K[i].class_id == l) { lx[nl] = K[i].x; ly[nl] = K[i].y; lr[nl] = K[i].response; idx[nl] = i; ++nl; }
        if (!nl) continue;
        const int nk = orc_distribute_octree(lx, ly, lr, NULL, nl, 0, w, 0, h, q_ext[l], keep, nl);
        for (int j = 0; j < nk; ++j) {
            const akz_pt* p = &K[idx[keep[j]]];
            if (m >= cap) { rc = -2; break; }
            const float ang = (g_akz_cv_mode & 2) ? akz_orientation_cv(&S, p) : akz_orientation(&S, p);
            orc_keypoint* kp = &kps[m];
            kp->x = p->x; kp->y = p->y; kp->size = p->size; kp->angle = ang; kp->response = p->response; kp->octave = p->octave; kp->class_id = p->class_id;
            akz_mldb(&S, p, ang, desc + (size_t)61 * m);
            if (kpsize) {
                const float s = powf(scale_factor, (float)l);
                float sn = maxSize;
                if (maxSize > minSize) sn = 1.0f + (s - minSize) * (maxSize0 - 1.0f) / (maxSize - minSize);
                kpsize[m] = sn;
            }
            ++m;
        }
    }
    free(lx); free(ly); free(lr); free(idx); free(keep); free(K);
    space_free(&S);
    if (n_out) *n_out = m;
    return rc;
}

/*
 * afv_oracle_batch.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY): batch driver of the CPU arm for the akaze61 + brisk48 workload
 * (BASELINE configs[3]) on host threads: akaze61 AND brisk48 extraction of B frames, then
 * FeatureMatcher::SearchForInitialization for P frame pairs on each feature's own keypoints and descriptors (61-byte and
 * 48-byte Hamming distances).
 */
#include "afv_oracle.h"
#include <malloc.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const uint8_t* frames; int B, w, h, nfeatures, nlevels; float scale_factor, detect_th;
    const int *pair_a, *pair_b; int P, window; float th_akaze, th_brisk, nnratio; int check_ori;
    int cap; orc_keypoint* kps; uint8_t* desc; float* ksz; int* n;
    orc_keypoint* kps48; uint8_t* desc48; float* ksz48; int* n48;
    int next_frame, next_pair, err; long total; int phase;
} akz_batch_job;

static void* akz_batch_worker(void* arg) {
    akz_batch_job* J = (akz_batch_job*)arg;
    const float max_size = powf(1.2f, 7.0f);
    if (J->phase == 0) {
        for (;;) {
            const int b = __atomic_fetch_add(&J->next_frame, 1, __ATOMIC_RELAXED);
            if (b >= J->B) break;
            uint8_t* d = J->desc + (size_t)b * J->cap * 61;
            int rc = orc_akaze61_extract(J->frames + (size_t)b * J->w * J->h, J->w, J->h, J->w, J->nfeatures, J->nlevels, J->scale_factor,
                                         J->detect_th, J->kps + (size_t)b * J->cap, d, J->ksz + (size_t)b * J->cap, J->cap, &J->n[b], NULL);
            if (rc) __atomic_store_n(&J->err, 1, __ATOMIC_RELAXED);
            /* FeatureExtractor_brisk48 with its own settings (settings/brisk48_settings.yaml: 8 octaves, 1.5, 34) */
            rc = orc_brisk48_extract(J->frames + (size_t)b * J->w * J->h, J->w, J->h, J->w, J->nfeatures, 8, 1.5f, 34.0f, ORC_BRISK_DENSE,
                                     J->kps48 + (size_t)b * J->cap, J->desc48 + (size_t)b * J->cap * 48, J->ksz48 + (size_t)b * J->cap, J->cap,
                                     &J->n48[b], NULL);
            if (rc) __atomic_store_n(&J->err, 1, __ATOMIC_RELAXED);
        }
    } else {
        long local = 0;
        for (;;) {
            const int p = __atomic_fetch_add(&J->next_pair, 1, __ATOMIC_RELAXED);
            if (p >= J->P) break;
            const int a = J->pair_a[p], b = J->pair_b[p];
            for (int pass = 0; pass < 2; ++pass) {
                const orc_keypoint* K = pass == 0 ? J->kps : J->kps48;
                const uint8_t* D = pass == 0 ? J->desc : J->desc48;
                const float* SZ = pass == 0 ? J->ksz : J->ksz48;
                const int* N = pass == 0 ? J->n : J->n48;
                const size_t db = pass == 0 ? 61 : 48;
                const int na = N[a];
                float* prev = (float*)malloc(sizeof(float) * 2 * (size_t)(na + 1));
                int* m12 = (int*)malloc(sizeof(int) * (size_t)(na + 1));
                for (int i = 0; i < na; ++i) { prev[2 * i] = K[(size_t)a * J->cap + i].x; prev[2 * i + 1] = K[(size_t)a * J->cap + i].y; }
                local += orc_search_for_initialization(pass == 0 ? 1 : 2, K + (size_t)a * J->cap, D + (size_t)a * J->cap * db, na,
                                                       K + (size_t)b * J->cap, D + (size_t)b * J->cap * db, SZ + (size_t)b * J->cap, N[b],
                                                       0.0f, 0.0f, (float)J->w, (float)J->h, max_size, prev, J->window,
                                                       pass == 0 ? J->th_akaze : J->th_brisk, J->nnratio, J->check_ori, m12);
                free(prev); free(m12);
            }
        }
        __atomic_fetch_add(&J->total, local, __ATOMIC_RELAXED);
    }
    return NULL;
}

long orc_akaze61_extract_match_batch(const uint8_t* frames, int B, int w, int h, int nfeatures, int nlevels, float scale_factor,
                                     float detect_th, const int* pair_a, const int* pair_b, int P, int window, float th_akaze,
                                     float th_brisk, float nnratio, int check_ori, int nthreads) {
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    akz_batch_job J;
    memset(&J, 0, sizeof(J));
    J.frames = frames; J.B = B; J.w = w; J.h = h; J.nfeatures = nfeatures; J.nlevels = nlevels; J.scale_factor = scale_factor;
    J.detect_th = detect_th; J.pair_a = pair_a; J.pair_b = pair_b; J.P = P; J.window = window; J.th_akaze = th_akaze;
    J.th_brisk = th_brisk; J.nnratio = nnratio; J.check_ori = check_ori;
    J.cap = nfeatures + 3 * nlevels + 64;
    J.kps = (orc_keypoint*)malloc(sizeof(orc_keypoint) * (size_t)B * J.cap);
    J.desc = (uint8_t*)malloc((size_t)B * J.cap * 61);
    J.desc48 = (uint8_t*)malloc((size_t)B * J.cap * 48);
    J.kps48 = (orc_keypoint*)malloc(sizeof(orc_keypoint) * (size_t)B * J.cap);
    J.ksz48 = (float*)malloc(sizeof(float) * (size_t)B * J.cap);
    J.n48 = (int*)calloc(B > 0 ? B : 1, sizeof(int));
    J.ksz = (float*)malloc(sizeof(float) * (size_t)B * J.cap);
    J.n = (int*)calloc(B > 0 ? B : 1, sizeof(int));
    if (nthreads < 1) nthreads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    for (int phase = 0; phase < 2; ++phase) {
        J.phase = phase;
        if (nthreads == 1) { akz_batch_worker(&J); continue; }
        int started = 0;
        for (int t = 0; t < nthreads; ++t) if (pthread_create(&th[t], NULL, akz_batch_worker, &J) == 0) ++started; else break;
        if (started == 0) akz_batch_worker(&J);
        for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    }
    const long total = J.err ? -1 : J.total;
    free(th); free(J.kps); free(J.desc); free(J.desc48); free(J.ksz); free(J.n); free(J.kps48); free(J.ksz48); free(J.n48);
    return total;
}

/* ---- vanilla ORB-SLAM2 extractor (afv_oracle_orbslam2.c) + SearchForInitialization, the CPU arm of bench.py --workload c2v ------- */
typedef struct {
    const uint8_t* frames; int B, w, h, nfeatures, nlevels; float scale_factor;
    const int *pair_a, *pair_b; int P, window; float th_low, nnratio; int check_ori;
    int cap; orc_keypoint* kps; uint8_t* desc; float* ksz; int* n;
    int next_frame, next_pair, err; long total; int phase;
} os2_batch_job;

static void* os2_batch_worker(void* arg) {
    os2_batch_job* J = (os2_batch_job*)arg;
    const float max_size = powf(1.2f, 7.0f);
    if (J->phase == 0) {
        for (;;) {
            const int b = __atomic_fetch_add(&J->next_frame, 1, __ATOMIC_RELAXED);
            if (b >= J->B) break;
            if (orc_orbslam2_extract(J->frames + (size_t)b * J->w * J->h, J->w, J->h, J->w, J->nfeatures, J->nlevels, J->scale_factor, 20, 7,
                                     J->kps + (size_t)b * J->cap, J->desc + (size_t)b * J->cap * 32, J->ksz + (size_t)b * J->cap, J->cap, &J->n[b]))
                __atomic_store_n(&J->err, 1, __ATOMIC_RELAXED);
        }
    } else {
        long local = 0;
        for (;;) {
            const int p = __atomic_fetch_add(&J->next_pair, 1, __ATOMIC_RELAXED);
            if (p >= J->P) break;
            const int a = J->pair_a[p], b = J->pair_b[p], na = J->n[a];
            float* prev = (float*)malloc(sizeof(float) * 2 * (size_t)(na + 1));
            int* m12 = (int*)malloc(sizeof(int) * (size_t)(na + 1));
            for (int i = 0; i < na; ++i) { prev[2 * i] = J->kps[(size_t)a * J->cap + i].x; prev[2 * i + 1] = J->kps[(size_t)a * J->cap + i].y; }
            local += orc_search_for_initialization(0, J->kps + (size_t)a * J->cap, J->desc + (size_t)a * J->cap * 32, na, J->kps + (size_t)b * J->cap,
                                                   J->desc + (size_t)b * J->cap * 32, J->ksz + (size_t)b * J->cap, J->n[b], 0.0f, 0.0f, (float)J->w,
                                                   (float)J->h, max_size, prev, J->window, J->th_low, J->nnratio, J->check_ori, m12);
            free(prev); free(m12);
        }
        __atomic_fetch_add(&J->total, local, __ATOMIC_RELAXED);
    }
    return NULL;
}

long orc_orbslam2_extract_match_batch(const uint8_t* frames, int B, int w, int h, int nfeatures, int nlevels, float scale_factor,
                                      const int* pair_a, const int* pair_b, int P, int window, float th_low, float nnratio, int check_ori,
                                      int nthreads) {
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    os2_batch_job J;
    memset(&J, 0, sizeof(J));
    J.frames = frames; J.B = B; J.w = w; J.h = h; J.nfeatures = nfeatures; J.nlevels = nlevels; J.scale_factor = scale_factor;
    J.pair_a = pair_a; J.pair_b = pair_b; J.P = P; J.window = window; J.th_low = th_low; J.nnratio = nnratio; J.check_ori = check_ori;
    J.cap = nfeatures + 3 * nlevels + 64;
    J.kps = (orc_keypoint*)malloc(sizeof(orc_keypoint) * (size_t)B * J.cap);
    J.desc = (uint8_t*)malloc((size_t)B * J.cap * 32);
    J.ksz = (float*)malloc(sizeof(float) * (size_t)B * J.cap);
    J.n = (int*)calloc(B > 0 ? B : 1, sizeof(int));
    if (nthreads < 1) nthreads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    for (int phase = 0; phase < 2; ++phase) {
        J.phase = phase;
        if (nthreads == 1) { os2_batch_worker(&J); continue; }
        int started = 0;
        for (int t = 0; t < nthreads; ++t) if (pthread_create(&th[t], NULL, os2_batch_worker, &J) == 0) ++started; else break;
        if (started == 0) os2_batch_worker(&J);
        for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    }
    const long total = J.err ? -1 : J.total;
    free(th); free(J.kps); free(J.desc); free(J.ksz); free(J.n);
    return total;
}

/*
 * afv_oracle_match.c -- CPU ORACLE for the FeatureMatcher path (test infrastructure, NOT product code).
 *
 * Restates the reference's matcher loops on plain arrays (no Frame/KeyFrame/MapPoint graph):
 * src/FeatureMatcher.cc:399-557 (SearchForInitialization), :186-283 (SearchByBoW), :73-154 (window search
 * core of SearchByProjection), :1508-1531 (DescriptorDistance), :1579-1668 (rotation histogram), and
 * src/Frame.cc:225-240, :333-394 (grid).  All of this arithmetic is in-repo C++ in the reference (integer
 * popcounts, float compares), so the restatement is definitional; no third-party numerics involved except
 * cv::norm(NORM_L2SQR) for sift128 (float diff, double accumulation; 1e-5 tolerance applies).
 */
#include "afv_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define GRID_COLS 64   /* include/Frame.h:41 */
#define GRID_ROWS 48   /* include/Frame.h:40 */
#define HISTO_LENGTH 30 /* src/FeatureMatcher.cc:64 */

int orc_descriptor_bytes(int desc_type) {
    switch (desc_type) {
        case 0: return 32;      /* DESC_ORB */
        case 1: return 61;      /* DESC_AKAZE61 */
        case 2: return 48;      /* DESC_BRISK */
        case 5: return 512;     /* DESC_SIFT128: 128 float */
        default: return -1;
    }
}

static inline int popc8(unsigned v) { return __builtin_popcount(v & 0xffu); }

float orc_descriptor_distance(int desc_type, const void* a, const void* b) {
    if (desc_type == 5) {                       /* cv::norm(a,b,NORM_L2SQR) on CV_32F: normL2Sqr<float,double> */
        const float *x = (const float*)a, *y = (const float*)b;
        double s = 0;
        for (int i = 0; i < 128; i += 4) {
            double v0 = (double)(x[i] - y[i]), v1 = (double)(x[i + 1] - y[i + 1]);
            double v2 = (double)(x[i + 2] - y[i + 2]), v3 = (double)(x[i + 3] - y[i + 3]);
            s += v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3;
        }
        return (float)s;
    }
    int nb = orc_descriptor_bytes(desc_type), d = 0;
    const uint8_t *x = (const uint8_t*)a, *y = (const uint8_t*)b;
    for (int i = 0; i < nb; ++i) d += popc8((unsigned)(x[i] ^ y[i]));
    return (float)d;
}

/* ---------------------------------------------------------------- grid ---------------------------- */
/* Frame::PosInGrid + AssignFeaturesToGrid; cells stored [ix][iy] (ix major) as CSR, items ascending. */
static inline int pos_in_grid(const orc_keypoint* kp, float minX, float minY, float invW, float invH,
                              int* cx, int* cy) {
    int px = (int)roundf((kp->x - minX) * invW);
    int py = (int)roundf((kp->y - minY) * invH);
    if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) return 0;
    *cx = px; *cy = py;
    return 1;
}

void orc_grid_build(const orc_keypoint* kps, int n, float minX, float minY, float invW, float invH,
                    int* cell_start, int* cell_items) {
    const int nc = GRID_COLS * GRID_ROWS;
    memset(cell_start, 0, sizeof(int) * (nc + 1));
    for (int i = 0; i < n; ++i) {
        int cx, cy;
        if (pos_in_grid(&kps[i], minX, minY, invW, invH, &cx, &cy)) cell_start[cx * GRID_ROWS + cy + 1]++;
    }
    for (int c = 0; c < nc; ++c) cell_start[c + 1] += cell_start[c];
    int* cur = (int*)malloc(sizeof(int) * nc);
    memcpy(cur, cell_start, sizeof(int) * nc);
    for (int i = 0; i < n; ++i) {
        int cx, cy;
        if (pos_in_grid(&kps[i], minX, minY, invW, invH, &cx, &cy)) cell_items[cur[cx * GRID_ROWS + cy]++] = i;
    }
    free(cur);
}

int orc_features_in_area(const orc_keypoint* kps, const float* kpsize, const int* cell_start,
                         const int* cell_items, float minX, float minY, float invW, float invH,
                         float x, float y, float r, float minSize, float maxSize, int* out, int cap) {
    int n = 0;
    int c0 = (int)floorf((x - minX - r) * invW); if (c0 < 0) c0 = 0;
    if (c0 >= GRID_COLS) return 0;
    int c1 = (int)ceilf((x - minX + r) * invW); if (c1 > GRID_COLS - 1) c1 = GRID_COLS - 1;
    if (c1 < 0) return 0;
    int r0 = (int)floorf((y - minY - r) * invH); if (r0 < 0) r0 = 0;
    if (r0 >= GRID_ROWS) return 0;
    int r1 = (int)ceilf((y - minY + r) * invH); if (r1 > GRID_ROWS - 1) r1 = GRID_ROWS - 1;
    if (r1 < 0) return 0;
    for (int ix = c0; ix <= c1; ++ix)
        for (int iy = r0; iy <= r1; ++iy) {
            int c = ix * GRID_ROWS + iy;
            for (int j = cell_start[c]; j < cell_start[c + 1]; ++j) {
                int idx = cell_items[j];
                if (kpsize[idx] < minSize) continue;
                if (kpsize[idx] > maxSize) continue;
                float dx = kps[idx].x - x, dy = kps[idx].y - y;
                if (fabsf(dx) < r && fabsf(dy) < r) { if (n < cap) out[n] = idx; ++n; }
            }
        }
    return n;
}

/* ---------------------------------------------------------------- rotation histogram -------------- */
/* updateRotationHistogram (src/FeatureMatcher.cc:1587-1597): rotFactor = 1/30 so only bins 0..12 fill. */
int orc_rot_bin(float angle1, float angle2) {
    float rot = angle1 - angle2;
    if (rot < 0.0) rot += 360.0f;
    const float rotFactor = 1.0f / (float)HISTO_LENGTH;
    int bin = (int)roundf(rot * rotFactor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

void orc_three_maxima(const int* cnt, int len, int* ind1, int* ind2, int* ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    *ind1 = *ind2 = *ind3 = -1;
    for (int i = 0; i < len; ++i) {
        const int s = cnt[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; *ind3 = *ind2; *ind2 = *ind1; *ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; *ind3 = *ind2; *ind2 = i; }
        else if (s > max3) { max3 = s; *ind3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { *ind2 = -1; *ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) { *ind3 = -1; }
}

/* ---------------------------------------------------------------- SearchForInitialization --------- */
int orc_search_for_initialization(int desc_type,
        const orc_keypoint* k1, const void* d1, int n1,
        const orc_keypoint* k2, const void* d2, const float* size2, int n2,
        float minX, float minY, float maxX, float maxY, float max_kpt_size,
        float* prev_matched, int window, float th_low, float nnratio, int check_ori, int* matches12) {
    const int D = orc_descriptor_bytes(desc_type);
    const float invW = (float)GRID_COLS / (maxX - minX), invH = (float)GRID_ROWS / (maxY - minY);
    int* cs = (int*)malloc(sizeof(int) * (GRID_COLS * GRID_ROWS + 1));
    int* ci = (int*)malloc(sizeof(int) * (n2 + 1));
    orc_grid_build(k2, n2, minX, minY, invW, invH, cs, ci);
    int* cand = (int*)malloc(sizeof(int) * (n2 + 1));
    float* matchedDist = (float*)malloc(sizeof(float) * (n2 + 1));
    int* matches21 = (int*)malloc(sizeof(int) * (n2 + 1));
    int* hist_items = (int*)malloc(sizeof(int) * (n1 + 1));   /* (bin of i1) or -1 */
    int hist_cnt[HISTO_LENGTH] = {0};
    for (int i = 0; i < n2; ++i) { matchedDist[i] = FLT_MAX; matches21[i] = -1; }
    for (int i = 0; i < n1; ++i) { matches12[i] = -1; hist_items[i] = -1; }
    int nMatches = 0;
    const float r = (float)window;
    for (int i1 = 0; i1 < n1; ++i1) {
        if (k1[i1].octave > 0) continue;
        int nc = orc_features_in_area(k2, size2, cs, ci, minX, minY, invW, invH,
                                      prev_matched[2 * i1], prev_matched[2 * i1 + 1], r, 0.0f, max_kpt_size, cand, n2);
        if (nc == 0) continue;
        const uint8_t* ref = (const uint8_t*)d1 + (long)i1 * D;
        float bestDist = FLT_MAX, bestDist2 = FLT_MAX;
        int bestIdx2 = -1;
        for (int c = 0; c < nc; ++c) {
            int i2 = cand[c];
            float dist = orc_descriptor_distance(desc_type, ref, (const uint8_t*)d2 + (long)i2 * D);
            if (matchedDist[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= th_low) {
            if (bestDist < bestDist2 * nnratio) {
                if (matches21[bestIdx2] >= 0) { matches12[matches21[bestIdx2]] = -1; nMatches--; }
                matches12[i1] = bestIdx2;
                matches21[bestIdx2] = i1;
                matchedDist[bestIdx2] = bestDist;
                nMatches++;
                if (check_ori) {
                    int bin = orc_rot_bin(k1[i1].angle, k2[bestIdx2].angle);
                    hist_items[i1] = bin; hist_cnt[bin]++;
                }
            }
        }
    }
    if (check_ori) {
        int i1m, i2m, i3m;
        orc_three_maxima(hist_cnt, HISTO_LENGTH, &i1m, &i2m, &i3m);
        for (int i1 = 0; i1 < n1; ++i1) {
            int b = hist_items[i1];
            if (b < 0 || b == i1m || b == i2m || b == i3m) continue;
            if (matches12[i1] >= 0) { matches12[i1] = -1; nMatches--; }
        }
    }
    for (int i1 = 0; i1 < n1; ++i1)
        if (matches12[i1] >= 0) { prev_matched[2 * i1] = k2[matches12[i1]].x; prev_matched[2 * i1 + 1] = k2[matches12[i1]].y; }
    free(cs); free(ci); free(cand); free(matchedDist); free(matches21); free(hist_items);
    return nMatches;
}

/* ---------------------------------------------------------------- stateless window search --------- */
void orc_match_window(int desc_type, const void* q, const float* qxy, const float* qr,
        const float* qmin_size, const float* qmax_size, int nq,
        const orc_keypoint* tk, const void* td, const float* tsize, int nt,
        float minX, float minY, float maxX, float maxY,
        int* best, float* bestd, float* secondd, float* best_size, float* second_size) {
    const int D = orc_descriptor_bytes(desc_type);
    const float invW = (float)GRID_COLS / (maxX - minX), invH = (float)GRID_ROWS / (maxY - minY);
    int* cs = (int*)malloc(sizeof(int) * (GRID_COLS * GRID_ROWS + 1));
    int* ci = (int*)malloc(sizeof(int) * (nt + 1));
    int* cand = (int*)malloc(sizeof(int) * (nt + 1));
    orc_grid_build(tk, nt, minX, minY, invW, invH, cs, ci);
    for (int i = 0; i < nq; ++i) {
        int nc = orc_features_in_area(tk, tsize, cs, ci, minX, minY, invW, invH, qxy[2 * i], qxy[2 * i + 1],
                                      qr[i], qmin_size[i], qmax_size[i], cand, nt);
        float bd = FLT_MAX, bd2 = FLT_MAX, bs = -1.0f, bs2 = -1.0f;
        int bi = -1;
        const uint8_t* ref = (const uint8_t*)q + (long)i * D;
        for (int c = 0; c < nc; ++c) {
            int idx = cand[c];
            float dist = orc_descriptor_distance(desc_type, ref, (const uint8_t*)td + (long)idx * D);
            if (dist < bd) { bd2 = bd; bd = dist; bi = idx; bs2 = bs; bs = tsize[idx]; }
            else if (dist < bd2) { bd2 = dist; bs2 = tsize[idx]; }
        }
        best[i] = bi; bestd[i] = bd; secondd[i] = bd2;
        if (best_size) best_size[i] = bs;
        if (second_size) second_size[i] = bs2;
    }
    free(cs); free(ci); free(cand);
}

void orc_match_bruteforce(int desc_type, const void* q, int nq, const void* t, int nt,
                          int* best, float* bestd, float* secondd) {
    const int D = orc_descriptor_bytes(desc_type);
    for (int i = 0; i < nq; ++i) {
        float bd = FLT_MAX, bd2 = FLT_MAX;
        int bi = -1;
        const uint8_t* ref = (const uint8_t*)q + (long)i * D;
        for (int j = 0; j < nt; ++j) {
            float dist = orc_descriptor_distance(desc_type, ref, (const uint8_t*)t + (long)j * D);
            if (dist < bd) { bd2 = bd; bd = dist; bi = j; }
            else if (dist < bd2) bd2 = dist;
        }
        best[i] = bi; bestd[i] = bd; secondd[i] = bd2;
    }
}

/* ---------------------------------------------------------------- SearchByBoW(KF, F) --------------- */
int orc_search_by_bow(int desc_type,
        const void* dkf, const int* kf_node, const int* kf_start, const int* kf_idx, int kf_nodes,
        const orc_keypoint* kkf,
        const void* df, const int* f_node, const int* f_start, const int* f_idx, int f_nodes,
        const orc_keypoint* kf_f, int nf,
        float th_low, float nnratio, int check_ori, int* match_f) {
    const int D = orc_descriptor_bytes(desc_type);
    int* bin_of = (int*)malloc(sizeof(int) * (nf + 1));
    int hist_cnt[HISTO_LENGTH] = {0};
    for (int i = 0; i < nf; ++i) { match_f[i] = -1; bin_of[i] = -1; }
    int nMatches = 0, a = 0, b = 0;
    while (a < kf_nodes && b < f_nodes) {
        if (kf_node[a] == f_node[b]) {
            for (int iKF = kf_start[a]; iKF < kf_start[a + 1]; ++iKF) {
                const int realKF = kf_idx[iKF];
                const uint8_t* ref = (const uint8_t*)dkf + (long)realKF * D;
                float bd1 = FLT_MAX, bd2 = FLT_MAX;
                int bestF = -1;
                for (int iF = f_start[b]; iF < f_start[b + 1]; ++iF) {
                    const int realF = f_idx[iF];
                    if (match_f[realF] >= 0) continue;
                    float dist = orc_descriptor_distance(desc_type, ref, (const uint8_t*)df + (long)realF * D);
                    if (dist < bd1) { bd2 = bd1; bd1 = dist; bestF = realF; }
                    else if (dist < bd2) bd2 = dist;
                }
                if (bd1 <= th_low) {
                    if (bd1 < nnratio * bd2) {
                        match_f[bestF] = realKF;
                        nMatches++;
                        if (check_ori) {
                            int bin = orc_rot_bin(kkf[realKF].angle, kf_f[bestF].angle);
                            bin_of[bestF] = bin; hist_cnt[bin]++;
                        }
                    }
                }
            }
            ++a; ++b;
        } else if (kf_node[a] < f_node[b]) {
            while (a < kf_nodes && kf_node[a] < f_node[b]) ++a;     /* lower_bound */
        } else {
            while (b < f_nodes && f_node[b] < kf_node[a]) ++b;
        }
    }
    if (check_ori) {
        int i1m, i2m, i3m;
        orc_three_maxima(hist_cnt, HISTO_LENGTH, &i1m, &i2m, &i3m);
        for (int i = 0; i < nf; ++i) {
            int bn = bin_of[i];
            if (bn < 0 || bn == i1m || bn == i2m || bn == i3m) continue;
            match_f[i] = -1; nMatches--;
        }
    }
    free(bin_of);
    return nMatches;
}

/* ---------------------------------------------------------------- SearchByProjection family --------------- */
/* FeatureMatcher::SearchByProjection(F, vpMapPoints, th) (src/FeatureMatcher.cc:73-154) with ratio_same_scale = 1, and
 * the best-only variant (:287-397) with ratio_same_scale = 0, on plain arrays: per query the projected position, the
 * search radius and the accepted size range; a train keypoint that already holds a map point (occupied) is skipped and
 * an accepted match occupies its keypoint for the later queries. Returns the number of matches. */
int orc_search_by_projection(int desc_type, const void* qdesc, const float* qxy, const float* qr, const float* qmin,
        const float* qmax, int nq, const orc_keypoint* tk, const void* td, const float* tsize, int nt,
        const uint8_t* occupied_in, float minX, float minY, float maxX, float maxY,
        float th, float nnratio, int ratio_same_scale, float tol, int* match_q) {
    const int D = orc_descriptor_bytes(desc_type);
    const float invW = (float)GRID_COLS / (maxX - minX), invH = (float)GRID_ROWS / (maxY - minY);
    const float invtol = 1.0f / tol;
    int* cs = (int*)malloc(sizeof(int) * (GRID_COLS * GRID_ROWS + 1));
    int* ci = (int*)malloc(sizeof(int) * (nt + 1));
    int* cand = (int*)malloc(sizeof(int) * (nt + 1));
    uint8_t* occ = (uint8_t*)calloc((size_t)nt + 1, 1);
    if (occupied_in) memcpy(occ, occupied_in, (size_t)nt);
    orc_grid_build(tk, nt, minX, minY, invW, invH, cs, ci);
    int nmatches = 0;
    for (int i = 0; i < nq; ++i) {
        match_q[i] = -1;
        int nc = orc_features_in_area(tk, tsize, cs, ci, minX, minY, invW, invH, qxy[2 * i], qxy[2 * i + 1], qr[i], qmin[i], qmax[i], cand, nt);
        if (nc == 0) continue;
        const uint8_t* ref = (const uint8_t*)qdesc + (long)i * D;
        float bestDist = FLT_MAX, bestDist2 = FLT_MAX, bestSize = -1.0f, bestSize2 = -1.0f;
        int bestIdx = -1;
        for (int c = 0; c < nc; ++c) {
            const int idx = cand[c];
            if (occ[idx]) continue;
            const float d = orc_descriptor_distance(desc_type, ref, (const uint8_t*)td + (long)idx * D);
            if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestIdx = idx; bestSize2 = bestSize; bestSize = tsize[idx]; }
            else if (d < bestDist2) { bestDist2 = d; bestSize2 = tsize[idx]; }
        }
        if (bestDist <= th) {
            if (ratio_same_scale) {
                if ((bestSize / bestSize2 < tol) && (bestSize / bestSize2 > invtol) && (bestSize2 > 0.0f)) {
                    if (bestDist > nnratio * bestDist2) continue;
                }
            }
            match_q[i] = bestIdx; occ[bestIdx] = 1; nmatches++;
        }
    }
    free(cs); free(ci); free(cand); free(occ);
    return nmatches;
}

/* ---------------------------------------------------------------- ComputeDistinctiveDescriptors ----------- */
/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:279-348): descriptor with the least median distance. */
static int cmp_float_asc(const void* a, const void* b) { float x = *(const float*)a, y = *(const float*)b; return (x > y) - (x < y); }
int orc_distinctive_descriptor(int desc_type, const void* desc, const int* obs, int N) {
    if (N <= 0) return -1;
    const int D = orc_descriptor_bytes(desc_type);
    float* dist = (float*)malloc(sizeof(float) * (size_t)N * N);
    for (int i = 0; i < N; ++i) {
        dist[i * N + i] = 0.0f;
        for (int j = i + 1; j < N; ++j) {
            float d = orc_descriptor_distance(desc_type, (const uint8_t*)desc + (long)obs[i] * D, (const uint8_t*)desc + (long)obs[j] * D);
            dist[i * N + j] = d; dist[j * N + i] = d;
        }
    }
    float bestMedian = FLT_MAX; int bestIdx = 0;
    float* v = (float*)malloc(sizeof(float) * N);
    for (int i = 0; i < N; ++i) {
        memcpy(v, dist + (size_t)i * N, sizeof(float) * N);
        qsort(v, N, sizeof(float), cmp_float_asc);
        float median = v[(int)(0.5 * (N - 1))];
        if (median < bestMedian) { bestMedian = median; bestIdx = i; }
    }
    free(dist); free(v);
    return bestIdx;
}

/* ---------------------------------------------------------------- DBoW2 tree descent ---------------------- */
/* TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)
 * (Thirdparty/DBoW2/include/DBoW2/TemplatedVocabulary.h:1346-1387) with the per-feature distances of
 * Thirdparty/DBoW2/src/FOrb.cpp:73-92, FAkaze61.cpp:85-100 (56 of 61 bytes), FBrisk.cpp:78-100, FSift128.cpp:48-59. */
static double orc_dbow_distance(int desc_type, const uint8_t* a, const uint8_t* b) {
    if (desc_type == 5) {
        const float *x = (const float*)a, *y = (const float*)b;
        double sqd = 0.;
        for (int i = 0; i < 128; ++i) sqd += (x[i] - y[i]) * (x[i] - y[i]);
        return sqd;
    }
    const int nb = desc_type == 0 ? 32 : desc_type == 1 ? 56 : 48;
    int d = 0;
    for (int i = 0; i < nb; ++i) d += popc8((unsigned)(a[i] ^ b[i]));
    return (double)d;
}
void orc_bow_transform(int desc_type, const void* desc, int n, const int* child_off, const int* child_ids, const void* node_desc,
                       const int* node_word, const double* node_weight, int depth_L, int levelsup,
                       int* word_id, double* weight, int* node_id) {
    const int D = orc_descriptor_bytes(desc_type);
    const int nid_level = depth_L - levelsup;
    for (int i = 0; i < n; ++i) {
        const uint8_t* f = (const uint8_t*)desc + (long)i * D;
        int final_id = 0, level = 0, nid = 0;
        while (child_off[final_id + 1] > child_off[final_id]) {
            ++level;
            const int c0 = child_off[final_id], c1 = child_off[final_id + 1];
            int best = child_ids[c0];
            double best_d = orc_dbow_distance(desc_type, f, (const uint8_t*)node_desc + (long)best * D);
            for (int c = c0 + 1; c < c1; ++c) {
                const int id = child_ids[c];
                const double d = orc_dbow_distance(desc_type, f, (const uint8_t*)node_desc + (long)id * D);
                if (d < best_d) { best_d = d; best = id; }
            }
            final_id = best;
            if (level == nid_level) nid = final_id;
        }
        word_id[i] = node_word[final_id]; weight[i] = node_weight[final_id]; node_id[i] = nid_level <= 0 ? 0 : nid;
    }
}

/* ---------------------------------------------------------------- batch driver for the CPU arm ------ */
/* Same work as one bench.py step, all in C so host threads scale (pthreads pulling frames, then pairs, from an
 * atomic counter): orb32 extraction of B frames + SearchForInitialization of frame pair_a[p] against pair_b[p].
 * Returns the total number of matches (a checksum so the work cannot be optimised away), -1 on error. */
#include <pthread.h>
#include <malloc.h>
typedef struct {
    const uint8_t* frames; int B, w, h, nfeatures, nlevels; float scale_factor, detect_th;
    const int *pair_a, *pair_b; int P, window; float th_low, nnratio; int check_ori;
    int cap; orc_keypoint* kps; uint8_t* desc; float* ksz; int* n;
    int next_frame, next_pair, err; long total; int phase;
} orc_batch_job;

static void* orc_batch_worker(void* arg) {
    orc_batch_job* J = (orc_batch_job*)arg;
    const float max_size = powf(1.2f, 7.0f);
    if (J->phase == 0) {
        for (;;) {
            const int b = __atomic_fetch_add(&J->next_frame, 1, __ATOMIC_RELAXED);
            if (b >= J->B) break;
            int rc = orc_orb32_extract(J->frames + (size_t)b * J->w * J->h, J->w, J->h, J->w, J->nfeatures, J->nlevels, J->scale_factor,
                                       J->detect_th, J->kps + (size_t)b * J->cap, J->desc + (size_t)b * J->cap * 32,
                                       J->ksz + (size_t)b * J->cap, J->cap, &J->n[b], NULL);
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
            local += orc_search_for_initialization(0, J->kps + (size_t)a * J->cap, J->desc + (size_t)a * J->cap * 32, na,
                                                   J->kps + (size_t)b * J->cap, J->desc + (size_t)b * J->cap * 32, J->ksz + (size_t)b * J->cap, J->n[b],
                                                   0.0f, 0.0f, (float)J->w, (float)J->h, max_size, prev, J->window, J->th_low, J->nnratio, J->check_ori, m12);
            free(prev); free(m12);
        }
        __atomic_fetch_add(&J->total, local, __ATOMIC_RELAXED);
    }
    return NULL;
}

long orc_orb32_extract_match_batch(const uint8_t* frames, int B, int w, int h, int nfeatures, int nlevels,
                                   float scale_factor, float detect_th, const int* pair_a, const int* pair_b, int P,
                                   int window, float th_low, float nnratio, int check_ori, int nthreads) {
    /* keep the per-frame scratch (pyramids, score maps: ~MBs) inside the per-thread heaps: with the default mmap /
     * trim thresholds every frame would mmap+munmap them and 128 threads serialise on the kernel's mm lock */
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    orc_batch_job J;
    memset(&J, 0, sizeof(J));
    J.frames = frames; J.B = B; J.w = w; J.h = h; J.nfeatures = nfeatures; J.nlevels = nlevels; J.scale_factor = scale_factor;
    J.detect_th = detect_th; J.pair_a = pair_a; J.pair_b = pair_b; J.P = P; J.window = window; J.th_low = th_low; J.nnratio = nnratio;
    J.check_ori = check_ori; J.cap = nfeatures + 3 * nlevels + 64;
    J.kps = (orc_keypoint*)malloc(sizeof(orc_keypoint) * (size_t)B * J.cap);
    J.desc = (uint8_t*)malloc((size_t)B * J.cap * 32);
    J.ksz = (float*)malloc(sizeof(float) * (size_t)B * J.cap);
    J.n = (int*)calloc(B > 0 ? B : 1, sizeof(int));
    if (nthreads < 1) nthreads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    for (int phase = 0; phase < 2; ++phase) {
        J.phase = phase;
        if (nthreads == 1) { orc_batch_worker(&J); continue; }
        int started = 0;
        for (int t = 0; t < nthreads; ++t) if (pthread_create(&th[t], NULL, orc_batch_worker, &J) == 0) ++started; else break;
        if (started == 0) orc_batch_worker(&J);
        for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    }
    const long total = J.err ? -1 : J.total;
    free(th); free(J.kps); free(J.desc); free(J.ksz); free(J.n);
    return total;
}


/* ---------------------------------------------------------------- Frame::UndistortKeyPoints -------------------- */
/* reference src/Frame.cc:403-433 -> cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK): OpenCV computes in double,
 * normalises with 1/fx, runs 5 fixed iterations of the inverse distortion model (default TermCriteria is count-only) and
 * re-projects with the same K.  Pinned to cv2 4.13.0 (tests/golden/undistort_cv2.npz, tools/make_golden_undistort.py). */
void orc_undistort_keypoints(const orc_keypoint* kps, int n, const float* K4, const float* dist5, orc_keypoint* out) {
    const int identity = dist5[0] == 0.0f;
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    const double k1 = dist5[0], k2 = dist5[1], p1 = dist5[2], p2 = dist5[3], k3 = dist5[4];
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    for (int i = 0; i < n; ++i) {
        orc_keypoint kp = kps[i];
        if (!identity) {
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
        out[i] = kp;
    }
}

/* ---------------------------------------------------------------- SearchByProjection family, all variants -- */
/* One restatement with the options that distinguish the reference's projection-type searches (everything after the
 * projection prologue, which stays with the caller: queries arrive as projected position, radius and accepted size range):
 *   :73-154   TrackLocalMap        occupied + claim, ratio rule when best / second have the same scale (ratio_same_scale = 1)
 *   :287-397  Sim3 (loop closing)  occupied (vpMatched) + claim, best only, th = TH_LOW
 *   :1291-1402 TrackWithMotionModel occupied + claim, best only, th = TH_HIGH, orientation histogram (qangle)
 *   :1406-1506 relocalisation      occupied (CurrentFrame.pts) + claim, best only, th = reloc threshold, orientation histogram
 *   :794-942  Fuse (local mapping) NO occupied test, NO claim, best only, th = TH_LOW, reprojection gate
 *                                  e2 * GetKeyPt1DInf(idx) > 5.99 (monocular branch, :905-915) -> tinf1d
 *   :944-1064 Fuse (loop closing)  NO occupied test, NO claim, best only
 *   :1066-1287 SearchBySim3        two such stateless best-only searches (th = TH_HIGH) + agreement (orc_search_by_sim3)
 * A query with qr < 0 is skipped (the reference's `continue`s of the projection prologue).  match_q[i] = train index or -1.
 * With an orientation histogram the entries are the TRAIN indices (points[j] = NULL, :1601-1612): the match of that train
 * keypoint is removed.  Returns the number of matches left. */
int orc_search_by_projection_ex(int desc_type, const void* qdesc, const float* qxy, const float* qr, const float* qmin,
        const float* qmax, const float* qangle, int nq, const orc_keypoint* tk, const void* td, const float* tsize,
        const float* tinf1d, int nt, const uint8_t* occupied_in, int claim, float minX, float minY, float maxX, float maxY,
        float th, float nnratio, int ratio_same_scale, float tol, int* match_q) {
    const int D = orc_descriptor_bytes(desc_type);
    const float invW = (float)GRID_COLS / (maxX - minX), invH = (float)GRID_ROWS / (maxY - minY);
    const float invtol = 1.0f / tol;
    int* cs = (int*)malloc(sizeof(int) * (GRID_COLS * GRID_ROWS + 1));
    int* ci = (int*)malloc(sizeof(int) * (nt + 1));
    int* cand = (int*)malloc(sizeof(int) * (nt + 1));
    int* bin_of_q = (int*)malloc(sizeof(int) * (nq + 1));
    uint8_t* occ = (uint8_t*)calloc((size_t)nt + 1, 1);
    int hist_cnt[HISTO_LENGTH] = {0};
    if (occupied_in) memcpy(occ, occupied_in, (size_t)nt);
    orc_grid_build(tk, nt, minX, minY, invW, invH, cs, ci);
    int nmatches = 0;
    for (int i = 0; i < nq; ++i) {
        match_q[i] = -1; bin_of_q[i] = -1;
        if (qr[i] < 0.0f) continue;
        int nc = orc_features_in_area(tk, tsize, cs, ci, minX, minY, invW, invH, qxy[2 * i], qxy[2 * i + 1], qr[i], qmin[i], qmax[i], cand, nt);
        if (nc == 0) continue;
        const uint8_t* ref = (const uint8_t*)qdesc + (long)i * D;
        float bestDist = FLT_MAX, bestDist2 = FLT_MAX, bestSize = -1.0f, bestSize2 = -1.0f;
        int bestIdx = -1;
        for (int c = 0; c < nc; ++c) {
            const int idx = cand[c];
            if (occ[idx]) continue;
            if (tinf1d) {                                        /* Fuse, monocular branch (:905-915) */
                const float ex = qxy[2 * i] - tk[idx].x, ey = qxy[2 * i + 1] - tk[idx].y;
                const float e2 = ex * ex + ey * ey;
                if (e2 * tinf1d[idx] > 5.99) continue;
            }
            const float d = orc_descriptor_distance(desc_type, ref, (const uint8_t*)td + (long)idx * D);
            if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestIdx = idx; bestSize2 = bestSize; bestSize = tsize[idx]; }
            else if (d < bestDist2) { bestDist2 = d; bestSize2 = tsize[idx]; }
        }
        if (bestDist <= th) {
            if (ratio_same_scale) {
                if ((bestSize / bestSize2 < tol) && (bestSize / bestSize2 > invtol) && (bestSize2 > 0.0f)) {
                    if (bestDist > nnratio * bestDist2) continue;
                }
            }
            match_q[i] = bestIdx; nmatches++;
            if (claim) occ[bestIdx] = 1;
            if (qangle) { const int bin = orc_rot_bin(qangle[i], tk[bestIdx].angle); bin_of_q[i] = bin; hist_cnt[bin]++; }
        }
    }
    if (qangle) {
        int i1m, i2m, i3m;
        orc_three_maxima(hist_cnt, HISTO_LENGTH, &i1m, &i2m, &i3m);
        for (int i = 0; i < nq; ++i) {
            const int b = bin_of_q[i];
            if (b < 0 || b == i1m || b == i2m || b == i3m) continue;
            match_q[i] = -1; nmatches--;
        }
    }
    free(cs); free(ci); free(cand); free(occ); free(bin_of_q);
    return nmatches;
}

/* FeatureMatcher::SearchBySim3 (src/FeatureMatcher.cc:1066-1287) after the projection prologue: direction 1 -> 2 projects the
 * map points of KF1's keypoints into KF2 (queries indexed by KF1 keypoint, qr < 0 = skipped), direction 2 -> 1 likewise; both are
 * stateless best-only searches with th = TH_HIGH; a pair is accepted when both directions agree (:1270-1284).
 * match12[i1] = i2 or -1.  Returns nFound. */
int orc_search_by_sim3(int desc_type,
        const void* q1desc, const float* q1xy, const float* q1r, const float* q1min, const float* q1max, int n1,
        const void* q2desc, const float* q2xy, const float* q2r, const float* q2min, const float* q2max, int n2,
        const orc_keypoint* k1, const void* d1, const float* size1, const orc_keypoint* k2, const void* d2, const float* size2,
        float minX, float minY, float maxX, float maxY, float th_high, int* match12) {
    int* m1 = (int*)malloc(sizeof(int) * (n1 + 1)); int* m2 = (int*)malloc(sizeof(int) * (n2 + 1));
    orc_search_by_projection_ex(desc_type, q1desc, q1xy, q1r, q1min, q1max, NULL, n1, k2, d2, size2, NULL, n2, NULL, 0, minX, minY, maxX, maxY,
                                th_high, 1.0f, 0, 1.0f, m1);
    orc_search_by_projection_ex(desc_type, q2desc, q2xy, q2r, q2min, q2max, NULL, n2, k1, d1, size1, NULL, n1, NULL, 0, minX, minY, maxX, maxY,
                                th_high, 1.0f, 0, 1.0f, m2);
    int nFound = 0;
    for (int i1 = 0; i1 < n1; ++i1) {
        match12[i1] = -1;
        const int idx2 = m1[i1];
        if (idx2 >= 0 && m2[idx2] == i1) { match12[i1] = idx2; nFound++; }
    }
    free(m1); free(m2);
    return nFound;
}

/* ---------------------------------------------------------------- BoW merge-join searches on per-feature node ids ----
 * The FeatureVector of a frame is given as the node id of every feature (what Vocabulary::transform returns per feature,
 * src/Vocabulary.cpp:200-204; < 0 = feature not in the vector): DBoW2's map<node, vector<feature>> lists the features of a node
 * in increasing feature index, nodes ascending, which is the (node, index) order used here.
 *   mode 0  SearchByBoW(KF, F)   src/FeatureMatcher.cc:186-283: frame-1 features need valid1 (map point present and good, :216-222),
 *           frame-2 features already matched are skipped (:232-233), best <= th_low and best < nnratio * second;
 *           out[i2] = i1 (vpMapPointMatches is indexed by the FRAME feature), orientation entries are frame-2 indices
 *   mode 1  SearchByBoW(KF, KF)  :561-660: valid1 / valid2 gates, vbMatched2, best < th_low (strict, :630) and the ratio;
 *           out[i1] = i2, orientation entries are frame-1 indices
 *   mode 2  SearchForTriangulation :662-790 (monocular): only features WITHOUT a map point on both sides (valid = "has map point"),
 *           candidates with dist > th_low or dist > best are skipped, epipole distance gate (:744-751), epipolar line gate
 *           (CheckDistEpipolarLine :165-183); vbMatched2 is never set by the reference, so frame-1 features are independent;
 *           out[i1] = i2.  F12 row-major, (ex, ey) the epipole in image 2, sigma2_2 = GetKeyPt1DSigma2 of frame 2.
 * Returns the number of matches. */
typedef struct { int node, idx; } bow_ent;
static int cmp_bow_ent(const void* a, const void* b) {
    const bow_ent* x = (const bow_ent*)a; const bow_ent* y = (const bow_ent*)b;
    if (x->node != y->node) return x->node < y->node ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}
int orc_bow_match(int mode, int desc_type,
        const orc_keypoint* k1, const void* d1, const int* node1, const uint8_t* valid1, int n1,
        const orc_keypoint* k2, const void* d2, const int* node2, const uint8_t* valid2, int n2,
        float th_low, float nnratio, int check_ori, const float* F12, float ex, float ey, const float* sigma2_2, int* out) {
    const int D = orc_descriptor_bytes(desc_type);
    bow_ent* e1 = (bow_ent*)malloc(sizeof(bow_ent) * (n1 + 1)); bow_ent* e2 = (bow_ent*)malloc(sizeof(bow_ent) * (n2 + 1));
    int m1 = 0, m2 = 0;
    for (int i = 0; i < n1; ++i) if (node1[i] >= 0) { e1[m1].node = node1[i]; e1[m1].idx = i; ++m1; }
    for (int i = 0; i < n2; ++i) if (node2[i] >= 0) { e2[m2].node = node2[i]; e2[m2].idx = i; ++m2; }
    qsort(e1, (size_t)m1, sizeof(bow_ent), cmp_bow_ent); qsort(e2, (size_t)m2, sizeof(bow_ent), cmp_bow_ent);
    const int nout = mode == 0 ? n2 : n1;
    int* bin_of = (int*)malloc(sizeof(int) * (nout + 1));
    uint8_t* matched2 = (uint8_t*)calloc((size_t)n2 + 1, 1);
    int hist_cnt[HISTO_LENGTH] = {0};
    for (int i = 0; i < nout; ++i) { out[i] = -1; bin_of[i] = -1; }
    int nMatches = 0, a = 0, b = 0;
    while (a < m1 && b < m2) {
        if (e1[a].node == e2[b].node) {
            int a1 = a, b1 = b;
            while (a1 < m1 && e1[a1].node == e1[a].node) ++a1;
            while (b1 < m2 && e2[b1].node == e2[b].node) ++b1;
            for (int ia = a; ia < a1; ++ia) {
                const int idx1 = e1[ia].idx;
                const uint8_t* ref = (const uint8_t*)d1 + (long)idx1 * D;
                if (mode == 2) {
                    if (valid1 && valid1[idx1]) continue;                       /* already a map point (:701-704) */
                    float bestDist = th_low; int bestIdx2 = -1;
                    for (int ib = b; ib < b1; ++ib) {
                        const int idx2 = e2[ib].idx;
                        if (valid2 && valid2[idx2]) continue;                   /* :722-724 (vbMatched2 is never set) */
                        const float dist = orc_descriptor_distance(desc_type, ref, (const uint8_t*)d2 + (long)idx2 * D);
                        if (dist > th_low || dist > bestDist) continue;
                        const float distex = ex - k2[idx2].x, distey = ey - k2[idx2].y;
                        if (distex * distex + distey * distey < 100.0f * sqrtf(sigma2_2[idx2])) continue;
                        /* CheckDistEpipolarLine */
                        const float la = k1[idx1].x * F12[0] + k1[idx1].y * F12[3] + F12[6];
                        const float lb = k1[idx1].x * F12[1] + k1[idx1].y * F12[4] + F12[7];
                        const float lc = k1[idx1].x * F12[2] + k1[idx1].y * F12[5] + F12[8];
                        const float num = la * k2[idx2].x + lb * k2[idx2].y + lc;
                        const float den = la * la + lb * lb;
                        if (den == 0) continue;
                        const float dsqr = num * num / den;
                        if (dsqr < 3.84f * sigma2_2[idx2]) { bestIdx2 = idx2; bestDist = dist; }
                    }
                    if (bestIdx2 >= 0) { out[idx1] = bestIdx2; nMatches++; }
                    continue;
                }
                if (valid1 && !valid1[idx1]) continue;
                float bd1 = FLT_MAX, bd2 = FLT_MAX; int best2 = -1;
                for (int ib = b; ib < b1; ++ib) {
                    const int idx2 = e2[ib].idx;
                    if (matched2[idx2]) continue;
                    if (mode == 1 && valid2 && !valid2[idx2]) continue;
                    const float dist = orc_descriptor_distance(desc_type, ref, (const uint8_t*)d2 + (long)idx2 * D);
                    if (dist < bd1) { bd2 = bd1; bd1 = dist; best2 = idx2; }
                    else if (dist < bd2) bd2 = dist;
                }
                const int pass_th = mode == 0 ? (bd1 <= th_low) : (bd1 < th_low);
                if (pass_th && bd1 < nnratio * bd2) {
                    matched2[best2] = 1; nMatches++;
                    const int oi = mode == 0 ? best2 : idx1;
                    out[oi] = mode == 0 ? idx1 : best2;
                    if (check_ori) { const int bin = orc_rot_bin(k1[idx1].angle, k2[best2].angle); bin_of[oi] = bin; hist_cnt[bin]++; }
                }
            }
            a = a1; b = b1;
        } else if (e1[a].node < e2[b].node) {
            while (a < m1 && e1[a].node < e2[b].node) ++a;
        } else {
            while (b < m2 && e2[b].node < e1[a].node) ++b;
        }
    }
    if (check_ori && mode != 2) {
        int i1m, i2m, i3m;
        orc_three_maxima(hist_cnt, HISTO_LENGTH, &i1m, &i2m, &i3m);
        for (int i = 0; i < nout; ++i) {
            const int bn = bin_of[i];
            if (bn < 0 || bn == i1m || bn == i2m || bn == i3m) continue;
            out[i] = -1; nMatches--;
        }
    }
    free(e1); free(e2); free(bin_of); free(matched2);
    return nMatches;
}

/* ---------------------------------------------------------------- Frame::isInFrustum ---------------------------------------- */
/* Frame::isInFrustum (src/Frame.cc:276-331) for M map points against one frame pose, followed by the window prologue of
 * SearchByProjection(F, vpMapPoints, radiusTh) (src/FeatureMatcher.cc:86-93).  IEEE float32 without contraction; 3-term sums are
 * evaluated left to right ((a0*b0 + a1*b1) + a2*b2) -- the reference uses Eigen (not vendored) compiled with -march=native, whose
 * evaluation order / FMA use is not defined by the reference sources, so agreement with a reference BUILD is to ~1e-6 relative, not
 * to the bit (tests use 1e-5).  pose16 = {Rcw row-major (9), tcw (3), twc (3), pad}; cam = {fx, fy, cx, cy, mbf};
 * bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY}.  Outputs per point: in_view, proj = (mTrackProjX, mTrackProjY, mTrackProjXR),
 * track = (trackSize, trackSigma, trackViewCos); optional query arrays for the projection search: qr = radius_factor *
 * RadiusByViewingCos(viewCos) * trackSize (-1 when not in view), qmin = trackSize / tol, qmax = trackSize * tol. */
void orc_is_in_frustum(const float* Pw, const float* normal, const float* min_dist, const float* max_dist, const float* ref_size,
                       const float* ref_sigma, const float* ref_dist, int M, const float* pose16, const float* cam5, const float* bounds4,
                       float viewing_cos_limit, float radius_factor, float size_tol, uint8_t* in_view, float* proj3, float* track3,
                       float* qr, float* qmin, float* qmax) {
    const float* R = pose16; const float* t = pose16 + 9; const float* c = pose16 + 12;
    for (int i = 0; i < M; ++i) {
        const float* P = Pw + 3 * i;
        in_view[i] = 0;
        if (qr) { qr[i] = -1.0f; qmin[i] = 0.0f; qmax[i] = 0.0f; }
        for (int k = 0; k < 3; ++k) { proj3[3 * i + k] = 0.0f; track3[3 * i + k] = 0.0f; }
        float Pc[3];
        for (int r = 0; r < 3; ++r) Pc[r] = ((R[3 * r] * P[0] + R[3 * r + 1] * P[1]) + R[3 * r + 2] * P[2]) + t[r];
        if (Pc[2] < 0.0f) continue;
        const float invz = 1.0f / Pc[2];
        const float u = cam5[0] * Pc[0] * invz + cam5[2];
        const float v = cam5[1] * Pc[1] * invz + cam5[3];
        if (u < bounds4[0] || u > bounds4[1]) continue;
        if (v < bounds4[2] || v > bounds4[3]) continue;
        const float PO[3] = {P[0] - c[0], P[1] - c[1], P[2] - c[2]};
        const float dist = sqrtf((PO[0] * PO[0] + PO[1] * PO[1]) + PO[2] * PO[2]);
        if (dist < min_dist[i] || dist > max_dist[i]) continue;
        const float* Pn = normal + 3 * i;
        const float viewCos = ((PO[0] * Pn[0] + PO[1] * Pn[1]) + PO[2] * Pn[2]) / dist;
        if (viewCos < viewing_cos_limit) continue;
        in_view[i] = 1;
        proj3[3 * i] = u; proj3[3 * i + 1] = v; proj3[3 * i + 2] = u - cam5[4] * invz;
        const float tsz = ref_size[i] * ref_dist[i] / dist;
        track3[3 * i] = tsz; track3[3 * i + 1] = ref_sigma[i] * ref_dist[i] / dist; track3[3 * i + 2] = viewCos;
        if (qr) {
            const float rv = (double)viewCos > 0.998 ? 2.5f : 4.0f;                /* RadiusByViewingCos (src/FeatureMatcher.cc:156-162) */
            qr[i] = radius_factor * rv * tsz;                                      /* radiusScale * radiusTh * ... * predictedSize (:91) */
            qmin[i] = tsz / size_tol; qmax[i] = tsz * size_tol;                    /* :95 */
        }
    }
}

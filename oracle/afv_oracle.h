/*
 * afv_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's per-frame feature front end for orb32 and of the
 * FeatureMatcher inner loops.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path (anyfeature-vslam_b200/csrc) never does.
 *
 * Where the reference's arithmetic lives in un-vendored OpenCV (cv::ORB), the algorithm is restated from
 * OpenCV's published behaviour and PINNED against cv2 4.13.0 run in the build container
 * (tests/test_oracle_golden.py, golden fixtures in tests/golden/ made by tools/make_golden.py).
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 */
#ifndef AFV_ORACLE_H
#define AFV_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LEVELS 16

/* == cv::KeyPoint field order (28 bytes) */
typedef struct { float x, y, size, angle, response; int octave, class_id; } orc_keypoint;

/* Geometry of cv::ORB's pyramid (scaleFactor 1.2f default kept by src/Feature_orb32.cpp:20-24). */
int  orc_orb_level_geometry(int w, int h, int nlevels, float orb_scale_factor,
                            int* lw, int* lh, float* lscale);
/* cv::ORB's per-level quota from maxFeatures (= nfeatures*10, src/Feature_orb32.cpp:22) and the
 * extractor's own per-level quota (src/FeatureExtractor.cpp:97-108): same formula. */
void orc_features_per_level(int nfeatures, int nlevels, float scale_factor, int* quota);

/* cv::resize(..., INTER_LINEAR_EXACT) 8UC1 (used by cv::ORB to build level l from level l-1). */
void orc_resize_linear_exact_u8(const uint8_t* src, int sw, int sh, int sstride,
                                uint8_t* dst, int dw, int dh, int dstride);
/* Whole pyramid into one tight buffer; level l starts at offs[l], stride lw[l]. Returns total bytes. */
long orc_orb_pyramid(const uint8_t* gray, int w, int h, int stride, int nlevels, float orb_scale_factor,
                     uint8_t* out, long* offs, int* lw, int* lh, float* lscale);

/* cv::FAST(threshold, nonmax=true, TYPE_9_16): raster-ordered (x,y,score). Returns count (may exceed cap;
 * only the first cap are written). */
int  orc_fast9_16_nms(const uint8_t* img, int w, int h, int stride, int threshold,
                      int* xs, int* ys, int* scores, int cap);
/* HarrisResponses(blockSize 7, k 0.04) of cv::ORB at integer level coords, reflect-101 outside. */
float orc_harris7(const uint8_t* img, int w, int h, int stride, int x, int y);
/* ICAngles (radius 15) + scalar cv::fastAtan2. */
float orc_ic_angle(const uint8_t* img, int w, int h, int stride, int x, int y);
float orc_fast_atan2(float y, float x);
/* The 7x7 sigma=2 blur cv::ORB::compute applies to each level ROI (float separable path). */
void orc_blur7_level(const uint8_t* img, int w, int h, int stride, uint8_t* out, int ostride);
/* 256-bit steered BRIEF at level coords (x,y); img = un-blurred level (read outside), blur = blurred. */
void orc_rbrief32(const uint8_t* img, const uint8_t* blur, int w, int h, int stride,
                  int x, int y, float angle_deg, uint8_t* desc32);

/* cv::ORB::detect restated for one level: FAST -> retainBest(2q) -> Harris -> retainBest(q), canonical
 * (raster) order. Writes level coords + harris response + fast score. Returns count. */
int  orc_orb_detect_level(const uint8_t* img, int w, int h, int stride, int fast_th, int quota,
                          int* xs, int* ys, float* harris, int* fastscore, int cap);

/* DistributeOctTree (src/ORBextractor.cc:239-458) on keypoints in level-0 coords; bounds (0,w,0,h).
 * `order[i]` is the canonical rank used to break exact max-response ties inside one node (smaller wins);
 * pointer-order ties in the (size,node*) sort are broken by node creation order (later-created = larger).
 * Writes the indices (into the input arrays) of the kept keypoints in list order. Returns count. */
int  orc_distribute_octree(const float* px, const float* py, const float* resp, const int* order, int n,
                           int minX, int maxX, int minY, int maxY, int N, int* keep, int cap);

/* Full orb32 FeatureExtractor::operator() (src/FeatureExtractor.cpp:111-121 with
 * src/Feature_orb32.cpp:11-65): returns merged keypoints (levels ascending) + n x 32 descriptors +
 * per-keypoint size (computeSize, src/FeatureExtractor.cpp:132-142). n_out may be NULL-checked by caller. */
int  orc_orb32_extract(const uint8_t* gray, int w, int h, int stride,
                       int nfeatures, int nlevels, float scale_factor, float detect_th,
                       orc_keypoint* kps, uint8_t* desc, float* kpsize, int cap, int* n_out,
                       int* n_candidates /* optional: total cv::ORB::detect keypoints */);

/* Image::GetGrayImage (src/Image.cpp:30-53): cvtColor 8-bit fixed point; rgb=1 -> channel 0 weighted as R. */
void orc_gray_from_color(const uint8_t* src, int channels, int rgb, int w, int h, int sstride, uint8_t* dst, int dstride);

/* ------------------------------------------------------------------ matcher restatement ---------- */
/* DescriptorDistance_* (src/Feature_orb32.cpp:67-83, Feature_akaze61.cpp:75-77, Feature_brisk48.cpp:62-64,
 * Feature_sift128.cpp:132-134). desc_type uses include/Types.h:24-34 ids (0 orb,1 akaze61,2 brisk,5 sift). */
float orc_descriptor_distance(int desc_type, const void* a, const void* b);
int   orc_descriptor_bytes(int desc_type);

/* Frame grid (src/Frame.cc:225-240, :384-394; 64x48 cells include/Frame.h:40-41).
 * cell_start has 64*48+1 entries (column-major [ix][iy] like mGrid), cell_items n entries. */
void orc_grid_build(const orc_keypoint* kps, int n, float minX, float minY, float invW, float invH,
                    int* cell_start, int* cell_items);
/* Frame::GetFeaturesInArea (src/Frame.cc:333-382). Returns count, indices in reference order. */
int  orc_features_in_area(const orc_keypoint* kps, const float* kpsize, const int* cell_start,
                          const int* cell_items, float minX, float minY, float invW, float invH,
                          float x, float y, float r, float minSize, float maxSize, int* out, int cap);

/* FeatureMatcher::SearchForInitialization (src/FeatureMatcher.cc:399-557) on plain arrays.
 * prev_matched (n1 x 2 floats) is updated in place like vbPrevMatched; matches12 gets n1 ints. */
int  orc_search_for_initialization(int desc_type,
        const orc_keypoint* k1, const void* d1, int n1,
        const orc_keypoint* k2, const void* d2, const float* size2, int n2,
        float minX, float minY, float maxX, float maxY, float max_kpt_size,
        float* prev_matched, int window, float th_low, float nnratio, int check_ori, int* matches12);

/* Stateless windowed best/second search (the data-parallel core of SearchByProjection,
 * src/FeatureMatcher.cc:73-154: candidates from GetFeaturesInArea, best/second distance + sizes). */
void orc_match_window(int desc_type, const void* q, const float* qxy, const float* qr,
        const float* qmin_size, const float* qmax_size, int nq,
        const orc_keypoint* tk, const void* td, const float* tsize, int nt,
        float minX, float minY, float maxX, float maxY,
        int* best, float* bestd, float* secondd, float* best_size, float* second_size);

/* Brute force N x M best / second (first minimum wins). */
void orc_match_bruteforce(int desc_type, const void* q, int nq, const void* t, int nt,
                          int* best, float* bestd, float* secondd);

/* SearchByBoW(KF,F) (src/FeatureMatcher.cc:186-283) with the FeatureVectors given as sorted
 * (node id, index list) segments; every KF keypoint is assumed to carry a valid map point. */
int  orc_search_by_bow(int desc_type,
        const void* dkf, const int* kf_node, const int* kf_start, const int* kf_idx, int kf_nodes,
        const orc_keypoint* kkf,
        const void* df, const int* f_node, const int* f_start, const int* f_idx, int f_nodes,
        const orc_keypoint* kf_f, int nf,
        float th_low, float nnratio, int check_ori, int* match_f /* nf: KF index or -1 */);

/* One bench.py step on the CPU (OpenMP over frames / pairs): extraction of B frames + SearchForInitialization of the
 * given pairs. Returns the total number of matches. Used only by bench.py's cpu_baseline / --impl reference legs. */
long orc_orb32_extract_match_batch(const uint8_t* frames, int B, int w, int h, int nfeatures, int nlevels,
                                   float scale_factor, float detect_th, const int* pair_a, const int* pair_b, int P,
                                   int window, float th_low, float nnratio, int check_ori, int nthreads);

/* SearchByProjection family on arrays (src/FeatureMatcher.cc:73-154 with ratio_same_scale=1, :287-397 with 0). */
int orc_search_by_projection(int desc_type, const void* qdesc, const float* qxy, const float* qr, const float* qmin,
        const float* qmax, int nq, const orc_keypoint* tk, const void* td, const float* tsize, int nt,
        const uint8_t* occupied_in, float minX, float minY, float maxX, float maxY,
        float th, float nnratio, int ratio_same_scale, float tol, int* match_q);

/* All projection-type searches of the reference after their projection prologue (options select the variant: see the table
 * above the definition in afv_oracle_match.c); qr < 0 skips a query; qangle != NULL adds the orientation histogram;
 * tinf1d != NULL adds Fuse's monocular reprojection gate; claim = accepted matches occupy their train keypoint. */
int orc_search_by_projection_ex(int desc_type, const void* qdesc, const float* qxy, const float* qr, const float* qmin,
        const float* qmax, const float* qangle, int nq, const orc_keypoint* tk, const void* td, const float* tsize,
        const float* tinf1d, int nt, const uint8_t* occupied_in, int claim, float minX, float minY, float maxX, float maxY,
        float th, float nnratio, int ratio_same_scale, float tol, int* match_q);
/* SearchBySim3 (src/FeatureMatcher.cc:1066-1287): two stateless directed searches + agreement; match12[i1] = i2 or -1. */
int orc_search_by_sim3(int desc_type,
        const void* q1desc, const float* q1xy, const float* q1r, const float* q1min, const float* q1max, int n1,
        const void* q2desc, const float* q2xy, const float* q2r, const float* q2min, const float* q2max, int n2,
        const orc_keypoint* k1, const void* d1, const float* size1, const orc_keypoint* k2, const void* d2, const float* size2,
        float minX, float minY, float maxX, float maxY, float th_high, int* match12);
/* BoW merge-join searches on per-feature node ids: mode 0 SearchByBoW(KF,F) (:186-283), 1 SearchByBoW(KF,KF) (:561-660),
 * 2 SearchForTriangulation (:662-790, monocular). */
int orc_bow_match(int mode, int desc_type,
        const orc_keypoint* k1, const void* d1, const int* node1, const uint8_t* valid1, int n1,
        const orc_keypoint* k2, const void* d2, const int* node2, const uint8_t* valid2, int n2,
        float th_low, float nnratio, int check_ori, const float* F12, float ex, float ey, const float* sigma2_2, int* out);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:279-348) for one map point: index into obs[0..N). */
int orc_distinctive_descriptor(int desc_type, const void* desc, const int* obs, int N);

/* DBoW2 tree descent per feature (Vocabulary::transform, src/Vocabulary.cpp:156-207). */
void orc_bow_transform(int desc_type, const void* desc, int n, const int* child_off, const int* child_ids, const void* node_desc,
                       const int* node_word, const double* node_weight, int depth_L, int levelsup,
                       int* word_id, double* weight, int* node_id);

/* Frame::UndistortKeyPoints (src/Frame.cc:403-433), pinned to cv2 4.13.0 undistortPoints. */
void orc_undistort_keypoints(const orc_keypoint* kps, int n, const float* K4, const float* dist5, orc_keypoint* out);

/* rotation-consistency helpers (src/FeatureMatcher.cc:1579-1668) */
int  orc_rot_bin(float angle1, float angle2);
void orc_three_maxima(const int* hist_counts, int len, int* ind1, int* ind2, int* ind3);


/* ---- sift128 (afv_oracle_sift.c; PARITY UNPINNED: SiftGPU is not vendored by the reference) ------------------- */
float orc_sift_exp(float x);
float orc_sift_exp2(float x);
float orc_sift_atan2(float y, float x);
void  orc_sift_sincos(float a, float* s, float* c);
int   orc_sift_gauss_kernel(double sigma, float* taps, int max_r);
void  orc_sift_sigmas(double* dsig);
int   orc_sift_num_octaves(int w, int h);
long  orc_sift_scale_space(const uint8_t* gray, int w, int h, int stride, int what, int oct, int idx, float* out, int* ow, int* oh);
int   orc_sift_detect(const uint8_t* gray, int w, int h, int stride, int nfeatures, float* xyso, float* desc, int cap);
int   orc_sift_ref_octave(float s);
int   orc_sift128_extract(const uint8_t* gray, int w, int h, int stride, int nfeatures, int nlevels, float scale_factor,
                          orc_keypoint* kps, float* desc, float* kpsize, int cap, int* n_out, int* n_detected);
long  orc_sift128_extract_match_batch(const uint8_t* frames, int B, int w, int h, int nfeatures, int nlevels, float scale_factor,
                                      const int* pair_a, const int* pair_b, int P, int window, float th_low, float nnratio,
                                      int check_ori, int nthreads);


/* ---- akaze61 (afv_oracle_akaze.c; PARITY UNPINNED: libAKAZE is not vendored by the reference) -------------------- */
int   orc_akaze_schedule(int w, int h, int omax, int nsub, int* lw, int* lh, int* octave, int* sigma_size, float* esigma,
                         int* nsteps, float* tau);
int   orc_akaze_gauss_taps(float sigma, float* taps);
long  orc_akaze_scale_space(const uint8_t* gray, int w, int h, int stride, int omax, int nsub, int what, int level, float* out,
                            int* ow, int* oh, float* kcontrast);
int   orc_akaze_detect(const uint8_t* gray, int w, int h, int stride, int omax, int nsub, float dth, float* out5, int cap);
int   orc_akaze61_extract(const uint8_t* gray, int w, int h, int stride, int nfeatures, int nlevels, float scale_factor,
                          float detect_th, orc_keypoint* kps, uint8_t* desc, float* kpsize, int cap, int* n_out, int* n_detected);
void  orc_akaze_set_cv2_filter(int on);   /* test hook, bit 0: OpenCV's cross-level duplicate filter, bit 1: OpenCV's orientation search, instead of libAKAZE's (pins the rest to cv2) */
long  orc_akaze61_extract_match_batch(const uint8_t* frames, int B, int w, int h, int nfeatures, int nlevels, float scale_factor,
                                      float detect_th, const int* pair_a, const int* pair_b, int P, int window, float th_akaze,
                                      float th_brisk, float nnratio, int check_ori, int nthreads);

/* ---- brisk48 (afv_oracle_brisk.c; PARITY UNPINNED vs ETH brisk v2, detector / orientation / 512-bit core semi-pinned to
 * cv2 4.13.0's cv::BRISK) ------------------------------------------------------------------------------------------ */
#define ORC_BRISK_SEQUENTIAL 0   /* lazy score cache exactly as the authors' implementation (cv2-comparable) */
#define ORC_BRISK_DENSE      1   /* order-independent contract implemented by the CUDA path */
void  orc_resize_area_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
long  orc_brisk_layer(const uint8_t* gray, int w, int h, int stride, int octaves, int what, int layer, uint8_t* out, int* ow, int* oh);
int   orc_brisk_detect(const uint8_t* gray, int w, int h, int stride, int threshold, int octaves, int mode, float* out5, int cap);
int   orc_brisk_describe(const uint8_t* gray, int w, int h, int stride, orc_keypoint* kps, int n, int nbytes, int libm_angle,
                         uint8_t* desc, int* kept_index);
int   orc_brisk_scale_index(float size);
int   orc_brisk_size_list(unsigned* out64);
double orc_brisk_atan2(double y, double x);
int   orc_brisk48_extract(const uint8_t* gray, int w, int h, int stride, int nfeatures, int nlevels, float scale_factor,
                          float detect_th, int mode, orc_keypoint* kps, uint8_t* desc, float* kpsize, int cap, int* n_out,
                          int* n_detected);

/* ---- vanilla ORB-SLAM2 extractor (afv_oracle_orbslam2.c; reference src/ORBextractor.cc:460-676 built with VANILLA_ORB_SLAM2):
 * control flow checked against the reference's compiled code, OpenCV stages pinned to cv2 4.13.0 ---------------------------- */
void  orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void  orc_gaussblur7_fixed_u8(const uint8_t* img, int w, int h, int stride, uint8_t* out, int ostride);
int   orc_orbslam2_geometry(int w, int h, int nlevels, float scale_factor, int* lw, int* lh, float* sf, float* isf);
int   orc_orbslam2_detect_level(const uint8_t* img, int cols, int rows, int stride, int ini_th, int min_th,
                                int* xs, int* ys, int* scores, int cap);
void  orc_orbslam2_descriptor(const uint8_t* blur, int stride, int cx, int cy, float angle_deg, uint8_t* desc32);
int   orc_orbslam2_extract(const uint8_t* gray, int w, int h, int stride, int nfeatures, int nlevels, float scale_factor,
                           int ini_th, int min_th, orc_keypoint* kps, uint8_t* desc, float* kpsize, int cap, int* n_out);

long  orc_orbslam2_extract_match_batch(const uint8_t* frames, int B, int w, int h, int nfeatures, int nlevels, float scale_factor,
                                       const int* pair_a, const int* pair_b, int P, int window, float th_low, float nnratio, int check_ori,
                                       int nthreads);
/* Frame::isInFrustum (src/Frame.cc:276-331) + the window prologue of SearchByProjection (src/FeatureMatcher.cc:86-95), M map points. */
void  orc_is_in_frustum(const float* Pw, const float* normal, const float* min_dist, const float* max_dist, const float* ref_size,
                        const float* ref_sigma, const float* ref_dist, int M, const float* pose16, const float* cam5, const float* bounds4,
                        float viewing_cos_limit, float radius_factor, float size_tol, uint8_t* in_view, float* proj3, float* track3,
                        float* qr, float* qmin, float* qmax);

#ifdef __cplusplus
}
#endif
#endif

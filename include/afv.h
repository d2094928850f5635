/*
 * afv.h -- C ABI of the B200-native feature front end (drop-in boundary for AnyFeature-VSLAM's
 * FeatureExtractor / FeatureMatcher hot path).
 *
 * The reference has no FFI: its boundary is two C++ types.  Each entry point below names the reference
 * interface it replaces (file:line relative to the reference tree); the C++ mirror classes that keep the
 * reference's virtual signatures and call these functions are in anyfeature-vslam_b200/host/, and the
 * binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes, no exceptions across the ABI, 0 = success / negative = error
 * (text via afv_last_error()), caller-owned output buffers with explicit capacity.  All compute runs in
 * hand-written sm_100a CUDA kernels; there is NO CPU fallback: without a usable CUDA device every entry
 * point fails with AFV_ERR_NO_DEVICE.
 */
#ifndef AFV_H
#define AFV_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define AFV_OK               0
#define AFV_ERR_INVALID     -1   /* bad argument */
#define AFV_ERR_NO_DEVICE   -2   /* CUDA device / driver unavailable: the product path has no CPU fallback */
#define AFV_ERR_CUDA        -3   /* CUDA runtime error (see afv_last_error) */
#define AFV_ERR_CAPACITY    -4   /* an internal or caller buffer capacity was exceeded (frame reported in text) */
#define AFV_ERR_UNSUPPORTED -5   /* feature id without an extractor (surf64, kaze64, r2d2_128, anyFeat*) */

/* feature / descriptor ids == reference include/Types.h:11-45 and get_feature_id (:102-124) */
#define AFV_FEAT_ORB32      0
#define AFV_FEAT_AKAZE61    1
#define AFV_FEAT_BRISK48    2
#define AFV_FEAT_SIFT128    5
/* extractor-only id (descriptor type stays AFV_FEAT_ORB32 for the matcher): the reference's FeatureExtractor built with
 * VANILLA_ORB_SLAM2 (include/Definitions.h:8), i.e. operator()(..., vanillaOrbslam) of src/ORBextractor.cc:568-645 */
#define AFV_FEAT_ORB32_VANILLA 100

/* == cv::KeyPoint POD (28 bytes, same field order): what the reference's operator() fills
 * (include/FeatureExtractor.h:84-93). */
typedef struct { float x, y, size, angle, response; int octave, class_id; } afv_keypoint;

typedef struct afv_extractor afv_extractor;

/* ---- FeatureExtractor (reference include/FeatureExtractor.h:68-161, src/FeatureExtractor.cpp:74-129,
 *      src/Feature_orb32.cpp:11-65; factory src/Tracking.cc:1505-1553) ------------------------------------ */

/* Replaces the FeatureExtractor_<feat> constructor + FeatureExtractorSettings (nfeatures, numOctaves,
 * scaleFactor, detectionTh from settings/<feat>_settings.yaml).  max_batch/max_w/max_h size the device
 * arenas once (no allocation on the per-frame path).  Extractors built:
 *   AFV_FEAT_ORB32   (src/Feature_orb32.cpp)   descriptors CV_8U  N x 32, pinned bit-exact to cv2 4.13.0's cv::ORB
 *   AFV_FEAT_SIFT128 (src/Feature_sift128.cpp) descriptors CV_32F N x 128 (512-byte rows), angle in radians, class_id =
 *                    row in the SiftGPU list; n_octaves / scale_factor only drive mnFeaturesPerLevel and computeSize
 *                    (SiftGPU's own arguments are fixed by the reference, :13-52)
 *   AFV_FEAT_AKAZE61 (src/Feature_akaze61.cpp) descriptors CV_8U  N x 61 (MLDB-486), octave = libAKAZE octave, class_id =
 *                    evolution level (the reference's "octave", :63-65); omax = n_octaves/4, nsublevels = n_octaves/2,
 *                    dthreshold = detect_th (:10-12); frame width and height must be even
 *   AFV_FEAT_BRISK48 (src/Feature_brisk48.cpp) descriptors CV_8U  N x 48, octave = BRISK layer 0 .. 2*(n_octaves/2)-1 (:29-30),
 *                    angle in degrees, class_id -1; detector = BriskFeatureDetector(int(detect_th), n_octaves / 2, true) (:24-26),
 *                    keypoints whose sampling pattern leaves the image are removed by the descriptor stage like brisk's compute()
   AFV_FEAT_ORB32_VANILLA (src/ORBextractor.cc:79-177, :460-676; Frame::ExtractFeatures src/Frame.cc:245-246) descriptors CV_8U N x 32:
 *                    INTER_LINEAR pyramid, FAST(iniThFAST = int(detect_th)) per 30-px cell with the minThFAST = 7 fallback, octree over
 *                    the 16-px border rectangle, IC angle, fixed-point GaussianBlur, steered BRIEF; size = mvScaleFactor[octave],
 *                    response = FAST score; every pyramid level must be at least 62 x 62.  OpenCV stages pinned to cv2 4.13.0
 * sift128 / akaze61 / brisk48 implement the published algorithms in the parameterisation the reference selects; SiftGPU /
 * libAKAZE / ETH brisk are not vendored by the reference, so parity with them is UNPINNED (oracle/afv_oracle_{sift,akaze,brisk}.c
 * headers; the brisk48 detector, orientation and 512-bit descriptor core are pinned to cv2.BRISK, the 48-byte pair table is a
 * documented stand-in). */
int  afv_extractor_create(afv_extractor** out, int feature_id, int nfeatures, int n_octaves,
                          float scale_factor, float detect_th, int device,
                          int max_batch, int max_w, int max_h);
void afv_extractor_destroy(afv_extractor* ex);

/* Smallest `cap` the per-frame output buffers must have: nfeatures + 3*n_octaves (DistributeOctTree can
 * overshoot each level's quota by up to 3, src/ORBextractor.cc:357-360). */
int  afv_extractor_output_cap(const afv_extractor* ex);
/* GetScaleFactors / mnFeaturesPerLevel (include/FeatureExtractor.h:95-99, src/FeatureExtractor.cpp:97-108). */
int  afv_extractor_levels(const afv_extractor* ex, float* scale_factors, int* features_per_level);

/* FeatureExtractor::operator()(Image, keypoints, descriptors, ..., size) for ONE frame, host buffers
 * (src/FeatureExtractor.cpp:111-129).  gray: 8-bit single channel, `stride` bytes per row.
 * kps[cap], desc[cap*D] (D = 32 bytes orb32, 61 bytes akaze61, 48 bytes brisk48, 512 bytes = 128 floats sift128), kpsize[cap] (computeSize,
 * :132-142; may be NULL). */
int  afv_extract(afv_extractor* ex, const uint8_t* gray, int w, int h, int stride,
                 afv_keypoint* kps, void* desc, float* kpsize, int cap, int* n_out);

/* Batched form (additive extension: the reference handles one frame per call).  B frames, frame b starts
 * at gray + b*frame_stride.  Outputs are B x cap.  Host buffers; H2D/D2H copies are inside the call. */
int  afv_extract_batch(afv_extractor* ex, const uint8_t* gray, int B, int w, int h, int stride,
                       long frame_stride, afv_keypoint* kps, void* desc, float* kpsize, int cap, int* n_out);

/* Same with DEVICE buffers, asynchronous on `cuda_stream` (a cudaStream_t).  NULL = the extractor's own non-blocking
 * stream, ordered after everything already queued on the legacy default stream and synchronised before return; to run
 * ON the legacy default stream pass cudaStreamLegacy ((void*)0x1).  Capacity overflows are reported by
 * afv_extractor_status(). */
int  afv_extract_batch_device(afv_extractor* ex, const uint8_t* d_gray, int B, int w, int h, int stride,
                              long frame_stride, afv_keypoint* d_kps, void* d_desc, float* d_kpsize,
                              int cap, int* d_n_out, void* cuda_stream);
/* Synchronises the last device batch and returns AFV_OK or AFV_ERR_CAPACITY. */
int  afv_extractor_status(afv_extractor* ex);

/* Stage taps for parity tests (device -> host copy of an intermediate of the LAST batch).
 * sift128: what = 10 Gaussian image / 11 DoG image (level = octave * 8 + index, float), 12 SiftGPU-order list after the -tc2
 *          limit (x, y, s, o floats).  akaze61: what = 20..24 Lt / Lsmooth / Lx / Ly / Ldet of evolution level `level`,
 *          25 Feature_Detection list (x, y, size, response, class_id floats), 26 contrast factor.  brisk48: what = 30 layer image,
 *          31 AGAST 9-16 score image of layer `level` (u8), 32 detect list (x, y, size, response, layer floats).  vanilla ORB-SLAM2: what = 40 level image,
 *          41 blurred level, 42 FAST score map, 43 detect list in push order (uint32 (x-16) | (y-16) << 12 | score << 24), 44 octree
 *          keep list (8 bytes as in 4, level coordinates, FAST score as float).  orb32:
 *   what = 0: pyramid level image (w_l*h_l bytes, tight)      1: blurred level image
 *          2: FAST+NMS candidates (uint32 x | y<<12 | score<<24, unordered)
 *          3: cv::ORB::detect-equivalent list after both retainBest culls (uint32 packed xy, float response
 *             pairs: 8 bytes each, unordered)
 *          4: octree-kept list in node-list order (8 bytes each as in 3)                                   */
int  afv_debug_read(afv_extractor* ex, int what, int frame, int level, void* out, long cap_bytes, long* n_bytes);

/* Image::GetGrayImage (src/Image.cpp:30-53): interleaved 3- or 4-channel 8-bit image -> gray with OpenCV's 8-bit
 * cvtColor arithmetic (R,G,B weights 9798/19235/3735, +2^14, >>15).  `rgb` != 0 selects CV_RGB(A)2GRAY (channel 0
 * weighted as R), 0 selects CV_BGR(A)2GRAY.  NB the reference always passes rgb = true on imread's BGR data
 * (Tracking::mbRGB is never assigned, src/Tracking.cc:1447 vs include/Tracking.h:230): pass rgb = 1 to reproduce it.
 * Device buffers, B frames, asynchronous on `cuda_stream`. */
int  afv_gray_from_color(const uint8_t* d_src, int channels, int rgb, int B, int w, int h, int src_stride,
                         long src_frame_stride, uint8_t* d_gray, int gray_stride, long gray_frame_stride,
                         void* cuda_stream);

/* ---- FeatureMatcher (reference include/FeatureMatcher.h:36-118, src/FeatureMatcher.cc) ------------------
 * The reference methods take Frame/KeyFrame/MapPoint graphs; the ABI sits one step inside, on arrays:
 * (query descriptors, window / bucket / all candidates, train descriptors) -> best index, best and second
 * distance.  desc_type uses the ids above; distances are returned as float like Descriptor_Distance_Type
 * (include/Types.h:127).  All pointers are DEVICE pointers; calls are asynchronous on `cuda_stream`.       */

/* Train-frame side of Frame::AssignFeaturesToGrid (src/Frame.cc:225-240, 64x48 cells): builds the CSR grid
 * cell_start[64*48+1], cell_items[nt] for each of B frames (frame b at offset b*cap). n[b] keypoints each. */
int  afv_grid_build(const afv_keypoint* d_kps, const int* d_n, int B, int cap,
                    float min_x, float min_y, float max_x, float max_y,
                    int* d_cell_start, int* d_cell_items, void* cuda_stream);

/* Frame::UndistortKeyPoints (src/Frame.cc:403-433): cv::undistortPoints(pts, pts, mK, mDistCoef, Mat(), mK) on the keypoint
 * positions of B frames (device, B x cap rows, n[b] valid), every other cv::KeyPoint field copied.  K4 = {fx, fy, cx, cy} and
 * dist5 = {k1, k2, p1, p2, k3} are HOST float arrays (the reference keeps mK / mDistCoef as CV_32F).  dist5[0] == 0 copies the
 * keypoints unchanged like the reference (:405-409).  Arithmetic pinned bit for bit to cv2 4.13.0 (double precision, 5 fixed
 * iterations).  Chain afv_grid_build on the result for Frame::AssignFeaturesToGrid (:225-240). */
int  afv_undistort_keypoints(const afv_keypoint* d_kps, const int* d_n, int B, int cap, const float* K4, const float* dist5,
                             afv_keypoint* d_kps_un, void* cuda_stream);

/* Frame::isInFrustum (src/Frame.cc:276-331) for M map points against one frame, fused with the window prologue of
 * SearchByProjection(F, vpMapPoints, radiusTh) (src/FeatureMatcher.cc:86-95) so the result feeds afv_search_by_projection_ex directly.
 * Device arrays per map point: world position (M x 3), mean viewing normal (M x 3), GetMin/MaxDistanceInvariance, and the
 * PredictSize / PredictSigma inputs refSize, refSigma, refDistance (src/MapPoint.cc:432-442).  HOST arrays: pose16 = {Rcw row-major
 * (9), tcw (3), twc = camera centre (3), unused}, cam5 = {fx, fy, cx, cy, mbf}, bounds4 = {mnMinX, mnMaxX, mnMinY, mnMaxY}.
 * Outputs: d_in_view[M] (mbTrackInView), d_proj3 = (mTrackProjX, mTrackProjY, mTrackProjXR), d_track3 = (trackSize, trackSigma,
 * trackViewCos); optional (all three or none) d_qr = radius_factor * RadiusByViewingCos(viewCos) * trackSize with radius_factor =
 * radiusScale * radiusTh (-1 for a point out of view = skipped query), d_qmin / d_qmax = trackSize / and * size_tolerance.
 * Floating point: IEEE float32, one rounding per operation, 3-term sums left to right; the reference evaluates the same
 * expressions through Eigen built -march=native, so a reference build agrees to ~1e-6 relative (tests: 1e-5), not to the bit. */
int  afv_is_in_frustum(const float* d_Pw, const float* d_normal, const float* d_min_dist, const float* d_max_dist,
                       const float* d_ref_size, const float* d_ref_sigma, const float* d_ref_dist, int M, const float* pose16,
                       const float* cam5, const float* bounds4, float viewing_cos_limit, float radius_factor, float size_tolerance,
                       uint8_t* d_in_view, float* d_proj3, float* d_track3, float* d_qr, float* d_qmin, float* d_qmax,
                       void* cuda_stream);

/* Data-parallel core of SearchByProjection (src/FeatureMatcher.cc:73-154) / GetFeaturesInArea
 * (src/Frame.cc:333-382): for each query (descriptor, window centre qxy, radius, size gate) the best and
 * second-best train keypoint inside the window, first-minimum-wins in (cell x, cell y, index) order.      */
int  afv_match_window(int desc_type, const void* d_q, const float* d_qxy, const float* d_qr,
                      const float* d_qmin_size, const float* d_qmax_size, int nq,
                      const afv_keypoint* d_tk, const void* d_td, const float* d_tsize, int nt,
                      const int* d_cell_start, const int* d_cell_items,
                      float min_x, float min_y, float max_x, float max_y,
                      int* d_best, float* d_bestd, float* d_secondd, float* d_best_size, float* d_second_size,
                      void* cuda_stream);

/* FeatureMatcher::SearchForInitialization (src/FeatureMatcher.cc:399-557), batched over P independent
 * frame pairs: pair p matches frame a[p] (queries, octave-0 only) against frame b[p] (train) of a B x cap
 * extraction result.  Reproduces the sequential "already matched with a smaller distance" rule (:511-512),
 * match stealing (:530-537), the rotation histogram (:1579-1668) and the vbPrevMatched update (:552-554).
 * d_prev_matched: P x cap x 2 floats in/out, or NULL = first call of MonocularInitialization: vbPrevMatched is
 * F1's own keypoint positions (src/Tracking.cc:440-445) and nothing is written back;
 * d_matches12: P x cap ints out; d_nmatches: P ints out.                                                   */
int  afv_search_for_initialization(int desc_type, const afv_keypoint* d_kps, const void* d_desc,
                                   const float* d_kpsize, const int* d_n, int B, int cap,
                                   const int* d_pair_a, const int* d_pair_b, int P,
                                   float min_x, float min_y, float max_x, float max_y, float max_kpt_size,
                                   float* d_prev_matched, int window, float th_low, float nnratio,
                                   int check_orientation, int* d_matches12, int* d_nmatches,
                                   void* cuda_stream);

/* The same with a caller-owned workspace (afv_search_for_initialization_workspace_bytes, 16-byte aligned): no allocation of any kind
 * on the call path.  The plain form above takes its scratch from the device's stream-ordered pool (cudaMallocAsync, cached). */
size_t afv_search_for_initialization_workspace_bytes(int desc_type, int P, int cap);
int  afv_search_for_initialization_ws(int desc_type, const afv_keypoint* d_kps, const void* d_desc,
                                      const float* d_kpsize, const int* d_n, int B, int cap,
                                      const int* d_pair_a, const int* d_pair_b, int P,
                                      float min_x, float min_y, float max_x, float max_y, float max_kpt_size,
                                      float* d_prev_matched, int window, float th_low, float nnratio,
                                      int check_orientation, int* d_matches12, int* d_nmatches,
                                      void* d_workspace, size_t workspace_bytes, void* cuda_stream);

/* SearchByProjection, the two classic variants (src/FeatureMatcher.cc:73-154 TrackLocalMap, :287-397 Sim3); kept as the short form
 * of afv_search_by_projection_ex below (it reads the query count back from the device, i.e. it synchronises the stream once):
 * P independent problems, problem p projects queries q_start[p]..q_start[p+1]
 * (descriptor, projected position, search radius, accepted size range = predicted size / and * sizeTolerance) into train
 * frame d_frame[p] of a B x cap extraction result.  Sequential semantics reproduced exactly: queries in order; a train
 * keypoint that already holds a map point (d_occupied, or claimed by an earlier query of this call) is skipped;
 * best / second distance with their keypoint sizes; accept if best <= th and, when ratio_same_scale_only != 0, reject
 * when best and second are within the size tolerance of each other and best > nnratio * second (:139-146);
 * ratio_same_scale_only == 0 gives the best-only rule of :381-386.  d_match_q[q] = train index or -1. */
int  afv_search_by_projection(int desc_type, const void* d_qdesc, const float* d_qxy, const float* d_qr,
                              const float* d_qmin_size, const float* d_qmax_size, const int* d_q_start, int P,
                              const afv_keypoint* d_kps, const void* d_desc, const float* d_kpsize, const int* d_n,
                              int B, int cap, const int* d_frame, const uint8_t* d_occupied,
                              float min_x, float min_y, float max_x, float max_y,
                              float th, float nnratio, int ratio_same_scale_only, float size_tolerance,
                              int* d_match_q, int* d_nmatches, void* cuda_stream);

/* Every projection-type search of the reference AFTER its projection prologue (the caller keeps the few lines that project a
 * map point and read its predicted size; queries arrive as projected position, radius, accepted size range, descriptor):
 *   reference method (src/FeatureMatcher.cc)            occupied  claim  ratio_same_scale  d_qangle  d_inf1d   th
 *   SearchByProjection(F, vpMapPoints)        :73-154   F.pts     1      1                 NULL      NULL      TH_HIGH
 *   SearchByProjection(pKF, Scw, ..) Sim3     :287-397  vpMatched 1      0                 NULL      NULL      TH_LOW
 *   SearchByProjection(Cur, Last) motion      :1291-1402 Cur.pts  1      0                 Last angles NULL    TH_HIGH
 *   SearchByProjection(Cur, pKF, ..) reloc    :1406-1506 Cur.pts  1      0                 KF angles NULL      reloc th
 *   Fuse(pKF, vpMapPoints)                    :794-942  NULL      0      0                 NULL      GetKeyPt1DInf  TH_LOW
 *   Fuse(pKF, Scw, ..)                        :944-1064 NULL      0      0                 NULL      NULL      TH_LOW
 * (Fuse returns the keypoint every map point lands on; the add / replace decision on the map graph stays with the caller.)
 * d_qr[q] < 0 marks a query the prologue skipped.  d_qangle != NULL adds the orientation histogram (:1579-1668) over the accepted
 * matches and removes the matches outside the three dominant bins; d_inf1d ([B][cap]) adds Fuse's monocular reprojection gate
 * e2 * inf > 5.99 (:905-915).  nq_total = d_q_start[P] when the host knows it (no synchronisation), -1 to read it back.
 * d_workspace (afv_search_by_projection_workspace_bytes) may be NULL: stream-ordered allocation inside the call. */
size_t afv_search_by_projection_workspace_bytes(int desc_type, int P, int nq_total);
int  afv_search_by_projection_ex(int desc_type, const void* d_qdesc, const float* d_qxy, const float* d_qr,
                                 const float* d_qmin_size, const float* d_qmax_size, const float* d_qangle, const int* d_q_start, int P,
                                 int nq_total, const afv_keypoint* d_kps, const void* d_desc, const float* d_kpsize, const float* d_inf1d,
                                 const int* d_n, int B, int cap, const int* d_frame, const uint8_t* d_occupied, int claim,
                                 float min_x, float min_y, float max_x, float max_y, float th, float nnratio,
                                 int ratio_same_scale_only, float size_tolerance, int* d_match_q, int* d_nmatches,
                                 void* d_workspace, size_t workspace_bytes, void* cuda_stream);

/* FeatureMatcher::SearchBySim3 (src/FeatureMatcher.cc:1066-1287) after its projection prologue, P keyframe pairs: direction 1
 * projects the map points of frame1's keypoints into frame2 (problem p = queries q_start1[p]..q_start1[p+1], query j = keypoint j
 * of frame d_frame1[p]; d_qr < 0 = no map point / already matched / outside), direction 2 likewise; both are stateless best-only
 * searches with th_high; match12[q_start1[p] + i1] = i2 where both directions agree (:1270-1284), else -1. */
int  afv_search_by_sim3(int desc_type,
                        const void* d_q1desc, const float* d_q1xy, const float* d_q1r, const float* d_q1min, const float* d_q1max,
                        const int* d_q_start1, int nq1_total,
                        const void* d_q2desc, const float* d_q2xy, const float* d_q2r, const float* d_q2min, const float* d_q2max,
                        const int* d_q_start2, int nq2_total, int P,
                        const afv_keypoint* d_kps, const void* d_desc, const float* d_kpsize, const int* d_n, int B, int cap,
                        const int* d_frame1, const int* d_frame2, float min_x, float min_y, float max_x, float max_y, float th_high,
                        int* d_match12, int* d_nfound, void* cuda_stream);

/* BoW merge-join searches on per-feature node ids (what afv_bow_transform returns per feature; < 0 = not in the FeatureVector),
 * batched over P frame pairs (frame1 = d_pair_a[p], frame2 = d_pair_b[p]) of a B x cap extraction result:
 *   mode 0  SearchByBoW(KF, F)  (:186-283): d_valid[frame1] = keypoint holds a good map point (:216-222); d_match[p][i2] = i1
 *   mode 1  SearchByBoW(KF, KF) (:561-660): d_valid on both frames, strict best < th_low (:630);       d_match[p][i1] = i2
 *   mode 2  SearchForTriangulation (:662-790, monocular): d_valid = keypoint already HAS a map point (skipped on both sides),
 *           epipole gate (:744-751) and CheckDistEpipolarLine (:165-183) with d_F12[p] (row-major), d_epipole[p] = (ex, ey) in
 *           image 2 and d_sigma2 ([B][cap]) = GetKeyPt1DSigma2;                                         d_match[p][i1] = i2
 * d_valid may be NULL (mode 0 / 1: every keypoint valid; mode 2: no keypoint has a map point).  d_match is P x cap. */
int  afv_bow_match(int mode, int desc_type, const afv_keypoint* d_kps, const void* d_desc, const int* d_n, int B, int cap,
                   const int* d_node_id, const uint8_t* d_valid, const int* d_pair_a, const int* d_pair_b, int P,
                   float th_low, float nnratio, int check_orientation, const float* d_F12, const float* d_epipole,
                   const float* d_sigma2, int* d_match, int* d_nmatches, void* cuda_stream);

/* Brute-force N x M best / second (upper bound of every matcher; also MapPoint::ComputeDistinctiveDescriptors'
 * distance matrix, src/MapPoint.cc:312-324). */
int  afv_match_bruteforce(int desc_type, const void* d_q, int nq, const void* d_t, int nt,
                          int* d_best, float* d_bestd, float* d_secondd, void* cuda_stream);

/* The same for P frame pairs (queries = frame d_pair_a[p], train = frame d_pair_b[p]) of a B x cap extraction result, binary
 * descriptors; outputs are P x cap. */
int  afv_match_bruteforce_pairs(int desc_type, const void* d_desc, const int* d_n, int B, int cap, const int* d_pair_a,
                                const int* d_pair_b, int P, int* d_best, float* d_bestd, float* d_secondd, void* cuda_stream);

/* Windowed best / second search batched over P frame pairs (throughput form of afv_match_window): queries = the keypoints of
 * frame d_pair_a[p] (their descriptors), searched in frame d_pair_b[p] through its grid from afv_grid_build.  Window centre =
 * d_qxy[p][i] (P x cap x 2) or, when NULL, the query keypoint's own position; radius = d_qr[p][i] (P x cap; < 0 skips the query) or,
 * when NULL, `radius`; accepted size range d_qmin_size / d_qmax_size (P x cap) or NULL = no size gate.  Same candidate definition
 * and enumeration order as Frame::GetFeaturesInArea (src/Frame.cc:333-382) -- train keypoints that Frame::PosInGrid drops
 * (they round to column 64 / row 48) are not candidates, a frame matched against itself included; binary descriptors.
 * Outputs are P x cap; only the first d_n[d_pair_a[p]] entries of a row are written.  The train frame is staged in shared memory:
 * AFV_ERR_INVALID when cap exceeds what fits (about 4300 / 3100 / 2400 keypoints for 32 / 48 / 61-byte descriptors). */
int  afv_match_window_pairs(int desc_type, const afv_keypoint* d_kps, const void* d_desc, const float* d_kpsize, const int* d_n,
                            int B, int cap, const int* d_cell_start, const int* d_cell_items, const int* d_pair_a,
                            const int* d_pair_b, int P, const float* d_qxy, const float* d_qr, float radius,
                            const float* d_qmin_size, const float* d_qmax_size, float min_x, float min_y, float max_x, float max_y,
                            int* d_best, float* d_bestd, float* d_secondd, void* cuda_stream);

/* SearchByBoW(KF,F) (src/FeatureMatcher.cc:186-283) on FeatureVector segments (sorted node ids + CSR) for ONE pair.  kf_idx
 * must list only keyframe features that hold a good map point (the reference skips the others, :216-222); afv_bow_match is the
 * batched form with explicit validity masks. */
int  afv_search_by_bow(int desc_type,
                       const void* d_dkf, const afv_keypoint* d_kkf,
                       const int* d_kf_node, const int* d_kf_start, const int* d_kf_idx, int kf_nodes,
                       const void* d_df, const afv_keypoint* d_kf_f, int nf,
                       const int* d_f_node, const int* d_f_start, const int* d_f_idx, int f_nodes,
                       float th_low, float nnratio, int check_orientation,
                       int* d_match_f, int* d_nmatches, void* cuda_stream);

/* Vocabulary::transform (src/Vocabulary.cpp:156-207) -> DBoW2 TemplatedVocabulary::transform per feature
 * (Thirdparty/DBoW2/include/DBoW2/TemplatedVocabulary.h:1346-1387): descend the k-ary tree from the root choosing the
 * child with the smallest F<feat>::distance (first minimum wins; Thirdparty/DBoW2/src/FOrb.cpp:73-92 32 bytes,
 * FAkaze61.cpp:85-100 only the first 56 of 61 bytes, FBrisk.cpp:78-100 48 bytes, FSift128.cpp:48-59 L2^2) until a
 * leaf.  Tree as flat arrays: children of node i are child_ids[child_off[i] .. child_off[i+1]) (empty = leaf),
 * node_desc[i] its descriptor, node_word[i] / node_weight[i] the word id / weight of leaves.  Outputs per feature:
 * word id, weight, and the ancestor at level (depth L - levelsup) that keys the FeatureVector used by SearchByBoW
 * (levelsup = 4 in the reference).  All device pointers. */
int  afv_bow_transform(int desc_type, const void* d_desc, int n,
                       const int* d_child_off, const int* d_child_ids, const void* d_node_desc,
                       const int* d_node_word, const double* d_node_weight, int n_nodes, int depth_L, int levelsup,
                       int* d_word_id, double* d_weight, int* d_node_id, void* cuda_stream);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:279-348), batched over M map points: map point m observes the
 * descriptors d_desc[d_obs[s]] for s in [d_seg_start[m], d_seg_start[m+1]); all pairwise distances, per row the median
 * sorted[0.5*(N-1)], and the first row with the smallest median wins.  d_best[m] = position inside the segment (-1 for an
 * empty segment).  Segments longer than 256 observations are rejected (AFV_ERR_INVALID). */
int  afv_distinctive_descriptors(int desc_type, const void* d_desc, const int* d_obs, const int* d_seg_start, int M,
                                 int max_seg, int* d_best, void* cuda_stream);

/* FeatureMatcher::DescriptorDistance (src/FeatureMatcher.cc:1508-1531) for n pairs (a[i], b[i]). */
int  afv_descriptor_distance(int desc_type, const void* d_a, const void* d_b, int n, float* d_out,
                             void* cuda_stream);

/* Multi-GPU result exchange (BASELINE configs[4]): one rank's fixed-capacity results as ONE contiguous message
 *   n[B] i32 | nmatches[B] i32 | matches12[B][cap] i32 | kps[B][cap] (28 B) | desc[B][cap][desc_bytes] (padded to 4 bytes)
 * written by a single launch; the host side gathers the messages with NCCL (anyfeature-vslam_b200/sharding.py).  The last byte of an
 * unaligned descriptor block (B * cap * desc_bytes not a multiple of 4) is read up to 3 bytes past the array: allocate accordingly. */
size_t afv_pack_results_bytes(int B, int cap, int desc_bytes);
int    afv_pack_results(const int* d_n, const int* d_nmatches, const int* d_matches12, const afv_keypoint* d_kps, const void* d_desc,
                        int B, int cap, int desc_bytes, void* d_pack, void* cuda_stream);

/* ---- misc ------------------------------------------------------------------------------------------------ */
const char* afv_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long   afv_kernel_launches(void);
const char* afv_version(void);
/* Optional per-kernel CUDA-event timing used by bench.py's roofline leg (events on the launching stream).
 * afv_profile_read synchronises, sums elapsed ms and launch counts per kernel name (32-char slots), clears. */
int         afv_profile_enable(int on);
int         afv_profile_read(char* names, float* ms, int* calls, int max_n);

#ifdef __cplusplus
}
#endif
#endif /* AFV_H */

// ref_shim.hpp -- TEST INFRASTRUCTURE.  The smallest set of declarations that lets a handful of function bodies of the
// REFERENCE (extracted at build time from /root/reference by oracle/build_ref.py, never copied into this repo) compile
// without OpenCV / Eigen: cv::KeyPoint / Point / Mat look-alikes with the members those bodies touch, and skeleton
// declarations of the reference classes they belong to (member names as in include/Frame.h, include/FeatureMatcher.h,
// include/FeatureExtractor.h:164-175 of the reference).  cv::norm is this repo's restatement of OpenCV (third party).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

static inline int cvRound(double v) { return (int)lrint(v); }       // OpenCV: round half to even
typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_32F 5
#define GL_LUMINANCE 0x1909
#define GL_UNSIGNED_BYTE 0x1401
namespace cv {
template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T x_, T y_) : x(x_), y(y_) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; } };
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x = 0, y = 0, width = 0, height = 0; Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };
enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16, INTER_LINEAR = 1 };
#define CV_8UC1 0
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point2i Point;
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };
enum { NORM_L2SQR = 5, NORM_HAMMING = 6 };
struct Mat {                                   // continuous row-major view; type 0 = CV_8U, 5 = CV_32F
    int rows = 0, cols = 0, type_ = 0;
    unsigned char* data = nullptr;
    size_t step = 0;
    std::shared_ptr<std::vector<unsigned char>> own;      // set when the matrix owns its storage
    Mat() {}
    Mat(int r, int c, int t, void* d) : rows(r), cols(c), type_(t), data((unsigned char*)d), step((size_t)c * (t == 5 ? 4 : 1)) {}
    Mat(int r, int c, int t) { create(r, c, t); }
    void create(int r, int c, int t) {
        rows = r; cols = c; type_ = t; step = (size_t)c * (t == 5 ? 4 : 1);
        own = std::make_shared<std::vector<unsigned char>>((size_t)r * step + 8, 0); data = own->data();
    }
    void release() { rows = cols = 0; data = nullptr; own.reset(); }
    Mat row(int i) const { Mat m(1, cols, type_, data + (size_t)i * step); return m; }
    // the pieces of cv::Mat the vanilla ORB-SLAM2 extractor bodies (src/ORBextractor.cc:460-676) use: ROI views share the storage
    Mat(Size sz, int t) { create(sz.height, sz.width, t); }
    Mat view(int r0, int r1, int c0, int c1) const { Mat m(r1 - r0, c1 - c0, type_, data + (size_t)r0 * step + (size_t)c0 * (type_ == 5 ? 4 : 1)); m.step = step; m.own = own; return m; }
    Mat rowRange(int a, int b) const { return view(a, b, 0, cols); }
    Mat colRange(int a, int b) const { return view(0, rows, a, b); }
    Mat operator()(const Rect& r) const { return view(r.y, r.y + r.height, r.x, r.x + r.width); }
    Mat clone() const { Mat m(rows, cols, type_); for (int i = 0; i < rows; ++i) memcpy(m.data + (size_t)i * m.step, data + (size_t)i * step, m.step); return m; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t step1() const { return step / (type_ == 5 ? 4 : 1); }
    Mat getMat() const { return *this; }
    // Mat::zeros returns an expression in OpenCV: assigned to a header of the same size / type it clears the EXISTING storage
    // (computeDescriptorsORB, src/ORBextractor.cc:561, relies on that: its argument is a row range of the output matrix)
    struct Zeros { int r, c, t; };
    static Zeros zeros(int r, int c, int t) { return Zeros{r, c, t}; }
    Mat& operator=(const Zeros& z) {
        if (data && rows == z.r && cols == z.c && type_ == z.t) { for (int i = 0; i < rows; ++i) memset(data + (size_t)i * step, 0, (size_t)cols * (type_ == 5 ? 4 : 1)); }
        else create(z.r, z.c, z.t);
        return *this;
    }
    unsigned char* ptr(int r = 0) { return data + (size_t)r * step; }
    const unsigned char* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T& at(int r, int c) { return reinterpret_cast<T*>(data + (size_t)r * step)[c]; }
    template <typename T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(data + (size_t)r * step)[c]; }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
};
inline void vconcat(const std::vector<Mat>& src, Mat& dst) {      // row-wise concatenation of equally wide matrices
    int rows = 0, cols = 0, type = 0;
    for (const Mat& m : src) { rows += m.rows; if (m.rows) { cols = m.cols; type = m.type_; } }
    dst.create(rows, cols, type);
    size_t off = 0;
    for (const Mat& m : src) { if (m.rows) memcpy(dst.data + off, m.data, (size_t)m.rows * m.step); off += (size_t)m.rows * m.step; }
}
// OpenCV: NORM_HAMMING = popcount of the xor over all bytes; NORM_L2SQR on CV_32F accumulates the squared differences in double
inline double norm(const Mat& a, const Mat& b, int normType) {
    if (normType == NORM_HAMMING) {
        int d = 0;
        for (int i = 0; i < a.cols; ++i) d += __builtin_popcount((unsigned)(a.data[i] ^ b.data[i]));
        return (double)d;
    }
    const float* pa = a.ptr<float>(); const float* pb = b.ptr<float>();
    double s = 0;
    for (int i = 0; i < a.cols; ++i) { const double v = (double)pa[i] - (double)pb[i]; s += v * v; }
    return s;
}
typedef Mat& OutputArray;
// OpenCV functions the vanilla extractor calls: the tests plug the REAL cv2 4.13.0 functions in through these callbacks
struct VanillaCallbacks {
    int (*fast)(const unsigned char* data, int rows, int cols, int step, int threshold, float* xys, int cap);   // cv::FAST(img, kps, th, true)
    void (*resize)(const unsigned char* src, int srows, int scols, int sstep, unsigned char* dst, int drows, int dcols, int dstep);   // INTER_LINEAR
    void (*blur)(unsigned char* data, int rows, int cols, int step);                                        // GaussianBlur 7x7, sigma 2, REFLECT_101
    float (*atan2)(float y, float x);                                                                       // cv::fastAtan2
};
extern VanillaCallbacks g_vanilla;
inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nonmax) {
    (void)nonmax;
    std::vector<float> buf((size_t)img.rows * img.cols * 3 + 3);
    const int n = g_vanilla.fast(img.data, img.rows, img.cols, (int)img.step, threshold, buf.data(), img.rows * img.cols);
    kps.resize(n);
    for (int i = 0; i < n; ++i) { kps[i] = KeyPoint(); kps[i].pt.x = buf[3 * i]; kps[i].pt.y = buf[3 * i + 1]; kps[i].response = buf[3 * i + 2]; kps[i].size = 7.f; }
}
inline void resize(const Mat& src, Mat& dst, Size sz, double, double, int) {
    if (dst.rows != sz.height || dst.cols != sz.width) dst.create(sz.height, sz.width, src.type_);     // cv::Mat::create keeps a fitting ROI
    g_vanilla.resize(src.data, src.rows, src.cols, (int)src.step, dst.data, dst.rows, dst.cols, (int)dst.step);
}
inline void GaussianBlur(const Mat& src, Mat& dst, Size, double, double, int) { (void)src; g_vanilla.blur(dst.data, dst.rows, dst.cols, (int)dst.step); }   // in place in the reference
inline float fastAtan2(float y, float x) { return g_vanilla.atan2(y, x); }
inline int borderInterpolate101(int p, int len) { if (len == 1) return 0; while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; } return p; }
// copyMakeBorder(BORDER_REFLECT_101 [+ ISOLATED]) into a destination of the final size; src may already be the centre ROI of dst
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int) {
    if (dst.rows != src.rows + top + bottom || dst.cols != src.cols + left + right) dst.create(src.rows + top + bottom, src.cols + left + right, src.type_);
    Mat s = src.clone();
    for (int y = 0; y < dst.rows; ++y) for (int x = 0; x < dst.cols; ++x)
        dst.data[(size_t)y * dst.step + x] = s.data[(size_t)borderInterpolate101(y - top, s.rows) * s.step + borderInterpolate101(x - left, s.cols)];
}
}  // namespace cv
static inline int cvFloor(double v) { return (int)floor(v); }
static inline int cvCeil(double v) { return (int)ceil(v); }

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace ANYFEATURE_VSLAM {
using namespace std;
using namespace cv;
typedef float Descriptor_Distance_Type;
using KeypointIndex = int;
enum DescriptorType { DESC_ANYFEATNONBIN = 8, DESC_ANYFEATBIN = 7, DESC_R2D2 = 6, DESC_SIFT128 = 5, DESC_KAZE64 = 4, DESC_SURF64 = 3,
                      DESC_BRISK = 2, DESC_AKAZE61 = 1, DESC_ORB = 0 };
// the few Eigen fixed-size operations the projection prologue of SearchByProjection (src/FeatureMatcher.cc:1299-1308) uses
struct vec3f {
    float v[3] = {0, 0, 0}; float operator()(int i) const { return v[i]; } float& operator()(int i) { return v[i]; }
    float dot(const vec3f& o) const { return (v[0] * o.v[0] + v[1] * o.v[1]) + v[2] * o.v[2]; }
    float norm() const { return std::sqrt(dot(*this)); }
};
inline vec3f operator+(const vec3f& a, const vec3f& b) { vec3f r; for (int i = 0; i < 3; ++i) r.v[i] = a.v[i] + b.v[i]; return r; }
inline vec3f operator-(const vec3f& a, const vec3f& b) { vec3f r; for (int i = 0; i < 3; ++i) r.v[i] = a.v[i] - b.v[i]; return r; }
struct mat3f {
    float m[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    mat3f transpose() const { mat3f r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = m[j][i]; return r; }
    mat3f operator-() const { mat3f r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = -m[i][j]; return r; }
    vec3f operator*(const vec3f& x) const { vec3f r; for (int i = 0; i < 3; ++i) r.v[i] = (m[i][0] * x.v[0] + m[i][1] * x.v[1]) + m[i][2] * x.v[2]; return r; }
    vec3f row(int i) const { vec3f r; for (int j = 0; j < 3; ++j) r.v[j] = m[i][j]; return r; }
    float operator()(int i, int j) const { return m[i][j]; }
    mat3f operator/(float s) const { mat3f r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = m[i][j] / s; return r; }
};
inline mat3f operator*(float s, const mat3f& a) { mat3f r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = s * a.m[i][j]; return r; }
inline mat3f operator*(double s, const mat3f& a) { mat3f r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = (float)(s * a.m[i][j]); return r; }
template <int R, int C> struct blk_t { typedef mat3f type; };
template <> struct blk_t<3, 1> { typedef vec3f type; };
struct mat4f {
    float m[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    template <int R, int C> typename blk_t<R, C>::type block(int r, int c) const;
};
template <> inline mat3f mat4f::block<3, 3>(int r, int c) const { mat3f o; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o.m[i][j] = m[r + i][c + j]; return o; }
template <> inline vec3f mat4f::block<3, 1>(int r, int c) const { vec3f o; for (int i = 0; i < 3; ++i) o.v[i] = m[r + i][c]; return o; }

class KeyFrame;
struct MapPoint {                                // include/MapPoint.h: the members the extracted FeatureMatcher bodies touch
    bool mbTrackInView = true, bad = false;
    // Fuse / Sim3 / relocalisation searches: invariance range, normal, predicted size, bookkeeping the bodies call
    float minDist = 0.0f, maxDist = 3.0e38f; vec3f normal; bool inKF = false; int idxInKF2 = -1;
    MapPoint* replacedBy = nullptr; int addedObsIdx = -1;
    float GetMaxDistanceInvariance() const { return maxDist; }
    float GetMinDistanceInvariance() const { return minDist; }
    vec3f GetNormal() const { return normal; }
    // PredictSize / PredictSigma: the searches' test set-up presets trackSize; with realPredict the reference's own bodies
    // (src/MapPoint.cc:432-442, cut as PredictSizeRef / PredictSigmaRef) run on refSize / refSigma / refDistance
    bool realPredict = false; float refSize = 1, refSigma = 1, refDistance = 1; std::mutex mMutexPos;
    float PredictSizeRef(const float& currentDist); float PredictSigmaRef(const float& currentDist);
    float PredictSize(const float& d) { return realPredict ? PredictSizeRef(d) : trackSize; }
    float PredictSigma(const float& d) { return realPredict ? PredictSigmaRef(d) : trackSigma; }
    bool IsInKeyFrame(const std::shared_ptr<KeyFrame>&) const { return inKF; }
    int GetIndexInKeyFrame(const std::shared_ptr<KeyFrame>&) const { return idxInKF2; }
    void Replace(const std::shared_ptr<MapPoint>& p) { replacedBy = p.get(); bad = true; }
    void AddObservation(const std::shared_ptr<KeyFrame>&, size_t idx) { addedObsIdx = (int)idx; idxInKF2 = (int)idx; nobs++; }
    float trackSize = 1, trackViewCos = 1, mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, trackSigma = 1;
    cv::Mat desc; int nobs = 1;
    vec3f worldPos;
    vec3f GetWorldPos() const { return worldPos; }
    bool isBad() const { return bad; }
    cv::Mat GetDescriptor() const { return desc; }
    int NumberOfObservations() const { return nobs; }
    DescriptorType descriptorType = DESC_ORB;
};
typedef std::shared_ptr<MapPoint> Pt;

// per-feature distances: orb32 / akaze61 / brisk48 / sift128 bodies are extracted from the reference; the rest is unused here
float DescriptorDistance_orb32(const cv::Mat& a, const cv::Mat& b);
float DescriptorDistance_akaze61(const cv::Mat& a, const cv::Mat& b);
float DescriptorDistance_brisk48(const cv::Mat& a, const cv::Mat& b);
float DescriptorDistance_sift128(const cv::Mat& a, const cv::Mat& b);
inline float DescriptorDistance_anyFeatureNonBin(const cv::Mat&, const cv::Mat&) { return 0; }
inline float DescriptorDistance_anyFeatureBin(const cv::Mat&, const cv::Mat&) { return 0; }
inline float DescriptorDistance_r2d2_128(const cv::Mat&, const cv::Mat&) { return 0; }
inline float DescriptorDistance_kaze64(const cv::Mat&, const cv::Mat&) { return 0; }
inline float DescriptorDistance_surf64(const cv::Mat&, const cv::Mat&) { return 0; }

namespace DBoW2_shim { typedef std::map<unsigned int, std::vector<unsigned int>> FeatureVector; }   // DBoW2::FeatureVector
}  // namespace ANYFEATURE_VSLAM
namespace DBoW2 { using ANYFEATURE_VSLAM::DBoW2_shim::FeatureVector; }
namespace ANYFEATURE_VSLAM {

class Frame {                                   // include/Frame.h: the members the extracted bodies use
public:
    int N = 0;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors;
    vector<float> keyPtsSize{};
    float maxKeyPtSize{};
    static float mfGridElementWidthInv, mfGridElementHeightInv, mnMinX, mnMaxX, mnMinY, mnMaxY;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<cv::KeyPoint> mvKeys;
    mat4f Tcw; float fx = 1, fy = 1, cx = 0, cy = 0, mbf = 0, mb = 0;
    mat3f Rcw; vec3f tcw, twc;                   // SetPose / UpdatePoseMatrices members isInFrustum reads (src/Frame.cc:262-274)
    bool isInFrustum(Pt pMP, float viewingCosLimit);
    std::vector<bool> mvbOutlier;
    DBoW2::FeatureVector mFeatVec;
    std::vector<Pt> pts;                         // map point held by each keypoint
    std::vector<float> mvuRight;                 // < 0 in the monocular case
    float sizeTolerance{}, invSizeTolerance{};
    float GetKeyPtSize(const KeypointIndex& idx) const { return keyPtsSize[idx]; }
    void AssignFeaturesToGrid();
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const float& minSize, const float& maxSize) const;
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
};

class KeyFrame {                                // include/KeyFrame.h: what the extracted FeatureMatcher bodies read / call
public:
    std::vector<Pt> mappoints;
    std::vector<Pt> GetMapPointMatches() { return mappoints; }
    Pt GetMapPoint(const size_t& idx) { return mappoints[idx]; }
    std::set<Pt> GetMapPoints() { std::set<Pt> s; for (auto& p : mappoints) if (p && !p->isBad()) s.insert(p); return s; }
    void AddMapPoint(Pt pMP, const size_t& idx) { mappoints[idx] = pMP; }
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<cv::KeyPoint> mvKeysUn;
    int N = 0;
    float fx = 1, fy = 1, cx = 0, cy = 0, mbf = 0;
    std::vector<float> mvuRight, keyPtsSize, sigma2_1d, inf_1d;
    float sizeTolerance{};
    mat3f Rcw; vec3f tcw, Ow;
    mat3f GetRotation() { return Rcw; }
    vec3f GetTranslation() { return tcw; }
    vec3f GetCameraCenter() { return Ow; }
    float GetKeyPtSize(const KeypointIndex& i) const { return keyPtsSize[i]; }
    float GetKeyPt1DSigma2(const KeypointIndex& i) const { return sigma2_1d[i]; }
    float GetKeyPt1DInf(const KeypointIndex& i) const { return inf_1d[i]; }
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0, mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    std::vector<std::vector<std::vector<size_t>>> mGrid;
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;
    bool IsInImage(const float& x, const float& y) const;
};
typedef std::shared_ptr<KeyFrame> Keyframe;

class FeatureMatcher {                          // include/FeatureMatcher.h:36-118
public:
    FeatureMatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
    int SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12, const int& windowSize,
                                const DescriptorType& descriptorType);
    static Descriptor_Distance_Type DescriptorDistance(const cv::Mat& a, const cv::Mat& b, const DescriptorType& descriptorType_);
    int SearchByProjection(Frame& F, const vector<Pt>& vpMapPoints, const float& radiusTh);
    int SearchByBoW(Keyframe pKF, Frame& F, vector<Pt>& vpMapPointMatches);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float& radiusTh, const bool bMono);
    int SearchByProjection(Frame& CurrentFrame, Keyframe pKF, const std::set<Pt>& sAlreadyFound, const float& radiusTh, const bool& useHighMatchingThreshold);
    int SearchByProjection(Keyframe pKF, const mat4f& Scw, const std::vector<Pt>& vpPoints, std::vector<Pt>& vpMatched, const float& radiusTh);
    int SearchByBoW(Keyframe pKF1, Keyframe pKF2, std::vector<Pt>& vpMatches12);
    int SearchForTriangulation(Keyframe pKF1, Keyframe pKF2, const mat3f& F12, std::vector<pair<size_t, size_t>>& vMatchedPairs, const bool bOnlyStereo,
                               const DescriptorType& descriptorType);
    int SearchBySim3(Keyframe pKF1, Keyframe pKF2, std::vector<Pt>& vpMatches12, const float& s12, const mat3f& R12, const vec3f& t12, const float& radiusTh);
    int Fuse(Keyframe pKF, const vector<Pt>& vpMapPoints, const float& radiusTh);
    int Fuse(Keyframe pKF, const mat4f& Scw, const std::vector<Pt>& vpPoints, const float& radiusTh, vector<Pt>& vpReplacePoint);
    bool CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const mat3f& F12, const Keyframe pKF, const float& sigma2_kp2);
    float RadiusByViewingCos(const float& viewCos);
    static float radiusScale;
    static Descriptor_Distance_Type TH_LOW, TH_HIGH, descDistTh_high_reloc, descDistTh_low_reloc;
    static const int HISTO_LENGTH;
    vector<vector<int>> initRotationHistogram(float& rotFactor, const int& histLength);
    void updateRotationHistogram(vector<vector<int>>& rotHist, const KeypointIndex& idx, const cv::KeyPoint& keyPt, const cv::KeyPoint& refKeyPt,
                                 const float& rotFactor, const int& histLength);
    void filterMatchesWithOrientation(vector<vector<int>>& rotHist, vector<Pt>& points, int& nMatches);
    void filterMatchesWithOrientation(vector<vector<int>>& rotHist, vector<int>& matches, int& nMatches);
    void computeThreeMaxima(vector<vector<int>>& rotHist, int& ind1, int& ind2, int& ind3);
    float mfNNratio; bool mbCheckOrientation;
    const Descriptor_Distance_Type highestPossibleDistance{std::numeric_limits<Descriptor_Distance_Type>::max()};
};

class ExtractorNode {                           // include/FeatureExtractor.h:164-175
public:
    ExtractorNode() : bNoMore(false) {}
    void DivideNode(ExtractorNode& n1, ExtractorNode& n2, ExtractorNode& n3, ExtractorNode& n4);
    std::vector<cv::KeyPoint> vKeys;
    cv::Point2i UL, UR, BL, BR;
    std::list<ExtractorNode>::iterator lit;
    bool bNoMore;
};
class Image { public: cv::Mat img{}, grayImg{}, mask{}; };           // include/Image.h:12-27 (data members)

// cv::ORB look-alike: records the parameters the reference sets; detect() hands back a prepared list (cv2's real output in the
// tests), compute() calls back into the test harness (which runs cv2's real ORB.compute on exactly those keypoints)
}  // namespace ANYFEATURE_VSLAM
namespace cv {
template <typename T> using Ptr = std::shared_ptr<T>;
struct ORB {
    int maxFeatures = 500, edgeThreshold = 31, fastThreshold = 20, nLevels = 8;
    static std::vector<KeyPoint>* g_detect;                                        // prepared detect() output
    static void (*g_compute)(const KeyPoint* kps, int n, unsigned char* desc32);   // harness callback
    static int g_last[4];
    static Ptr<ORB> create() { return std::make_shared<ORB>(); }
    void setMaxFeatures(int v) { maxFeatures = v; } void setEdgeThreshold(int v) { edgeThreshold = v; }
    void setFastThreshold(int v) { fastThreshold = v; } void setNLevels(int v) { nLevels = v; }
    void detect(const Mat&, std::vector<KeyPoint>& k) { g_last[0] = maxFeatures; g_last[1] = edgeThreshold; g_last[2] = fastThreshold; g_last[3] = nLevels; k = *g_detect; }
    void compute(const Mat&, std::vector<KeyPoint>& k, Mat& desc) { desc.create((int)k.size(), 32, 0); if (!k.empty()) g_compute(k.data(), (int)k.size(), desc.data); }
};
}
namespace ANYFEATURE_VSLAM {

// SiftGPU look-alike (third party, not vendored): hands back a prepared feature list
struct SiftGPU {
    struct SiftKeypoint { float x, y, s, o; };
    std::vector<SiftKeypoint> keys; std::vector<float> desc;
    int RunSIFT(int, int, const void*, unsigned, unsigned) { return 1; }
    int GetFeatureNum() { return (int)keys.size(); }
    void GetFeatureVector(SiftKeypoint* k, float* d) {
        if (k) std::copy(keys.begin(), keys.end(), k);
        if (d) std::copy(desc.begin(), desc.end(), d);
    }
};
// libAKAZE look-alike (third party, not vendored): prepared detection list; descriptors / orientation by keypoint identity
struct AKAZEOptions { int omax = 4, nsublevels = 4, img_width = 0, img_height = 0; float dthreshold = 0.001f; };
namespace libAKAZE {
struct AKAZE {
    std::vector<cv::KeyPoint> detected;                              // Feature_Detection output
    std::vector<cv::KeyPoint> table_kp; std::vector<float> table_angle; std::vector<unsigned char> table_desc;   // per detected keypoint
    void Feature_Detection(std::vector<cv::KeyPoint>& kpts) { kpts = detected; }
    void Compute_Descriptors(std::vector<cv::KeyPoint>& kpts, cv::Mat& desc) {
        desc.create((int)kpts.size(), 61, CV_8U);
        for (size_t i = 0; i < kpts.size(); ++i)
            for (size_t j = 0; j < table_kp.size(); ++j)
                if (table_kp[j].pt.x == kpts[i].pt.x && table_kp[j].pt.y == kpts[i].pt.y && table_kp[j].class_id == kpts[i].class_id) {
                    kpts[i].angle = table_angle[j]; memcpy(desc.data + i * 61, table_desc.data() + j * 61, 61); break;
                }
    }
};
}

struct FeatureExtractorSettings {               // include/FeatureExtractor.h:24-66 (the fields the constructor / computeSize read)
    float detectTh = 0;
    int iniThFAST = 20, minThFAST = 7;
    float scaleFactor = 1.2f; int nOctaves = 8;
    float maxKeyPtSize = 0, minKeyPtSize = 1.0f, maxKeyPtSize0 = 0;
    static float scaleFactor0;
    static float GetDetectorNominalScaleFactor() { return scaleFactor0; }
};
// include/Types.h: mat2f (Eigen::Matrix2f) and CovarianceMethod as computeSigma / the vanilla operator() use them
struct mat2f {
    float m[2][2] = {{0, 0}, {0, 0}};
    float& operator()(int i, int j) { return m[i][j]; } float operator()(int i, int j) const { return m[i][j]; }
    static mat2f Identity() { mat2f r; r.m[0][0] = r.m[1][1] = 1.0f; return r; }
};
inline mat2f operator*(float s, const mat2f& a) { mat2f r; for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) r.m[i][j] = s * a.m[i][j]; return r; }
enum CovarianceMethod { NONE = 0, SIZE = 1 };
// src/ORBextractor.cc:68-72 with VANILLA_ORB_SLAM2 defined
const int PATCH_SIZE = 31;
const int HALF_PATCH_SIZE = 15;
const int EDGE_THRESHOLD = 19;
class FeatureExtractor {
public:
    FeatureExtractor() {}
    // vanilla ORB-SLAM2 members (include/FeatureExtractor.h:74-82, :140-158); initVanilla = the VANILLA constructor body (:79-136)
    void initVanilla(const int& nfeatures_, std::shared_ptr<FeatureExtractorSettings>& settings_);
    void operator()(const Image& img, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors, std::vector<mat2f>& keyPtsSigma2,
                    std::vector<mat2f>& keyPtsInf, std::vector<float>& keyPtsSize, const bool& vanillaOrbslam);
    void ComputePyramid(cv::Mat image);
    void ComputeKeyPointsOctTree(std::vector<std::vector<cv::KeyPoint>>& allKeypoints, const cv::Mat& mask);
    void computeSigma(std::vector<mat2f>& keyPtsSigma2, std::vector<mat2f>& keyPtsInf, const std::vector<float>& keyPtsSize,
                      const std::vector<cv::KeyPoint>& keypoints, const Image& img, const CovarianceMethod& method);
    std::vector<cv::Point> pattern;
    std::vector<int> umax;
    FeatureExtractor(const int& nfeatures_, std::shared_ptr<FeatureExtractorSettings>& settings_);
    virtual ~FeatureExtractor() {}
    std::shared_ptr<FeatureExtractorSettings> settings{};
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<cv::Mat> mvImagePyramid;
    std::vector<int> mnFeaturesPerLevel;
    void computeSize(std::vector<float>& keyPtsSize, const std::vector<cv::KeyPoint>& keypoints);
    // hooks of include/FeatureExtractor.h:114-134 with the reference's signatures
    virtual void detectAndCompute(const Image& img, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors) {}
    virtual void detectKeypoints(std::map<int, std::vector<cv::KeyPoint>>& keypoints_level, const Image& img, const float& detectTh, const int& nOctaves) const {}
    virtual void computeDescriptors(std::map<int, cv::Mat>& descriptors_level, std::map<int, std::vector<cv::KeyPoint>>& keypoints_level, const Image& img) const {}
    virtual void mergeKeypointLevels(std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors, std::map<int, cv::Mat>& descriptors_level,
                                     std::map<int, std::vector<cv::KeyPoint>>& keypoints_level) const;
    virtual void filterKeypoints(std::map<int, std::vector<cv::KeyPoint>>& keypoints_level, const cv::Mat& image, const cv::Mat& mask) const {}
    void filterKeypoints_notScaled(std::map<int, std::vector<cv::KeyPoint>>& keypoints_level, const cv::Mat& image, const cv::Mat& mask) const;
    virtual int GetKeypointOctave(const cv::KeyPoint& keypoint) const { return keypoint.octave; }
    virtual float GetKeypointSize(const cv::KeyPoint& keypoint) const {      // src/Feature_orb32.cpp:59-61 (same in every Feature_*)
        return powf(settings->GetDetectorNominalScaleFactor(), float(GetKeypointOctave(keypoint))); }
    int nfeatures = 1000;
    std::vector<cv::KeyPoint> DistributeOctTree(const std::vector<cv::KeyPoint>& vToDistributeKeys, const int& minX, const int& maxX,
                                                const int& minY, const int& maxY, const int& N, const int& level) const;
};
class FeatureExtractor_orb32 : public FeatureExtractor {            // include/Feature_orb32.h
public:
    cv::Ptr<cv::ORB> orb32_extractor;
    FeatureExtractor_orb32(const int& nfeatures_, std::shared_ptr<FeatureExtractorSettings>& settings_);
    void detectAndCompute(const Image& img, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors) override;
    void initializeExtractor(const Image& img);
    void detectKeypoints(std::map<int, std::vector<cv::KeyPoint>>& keypoints_level, const Image& img, const float& detectTh, const int& nOctaves) const override;
    void computeDescriptors(std::map<int, cv::Mat>& descriptors_level, std::map<int, std::vector<cv::KeyPoint>>& keypoints_level, const Image& img) const override;
    void filterKeypoints(std::map<int, std::vector<cv::KeyPoint>>& keypoints_level, const cv::Mat& image, const cv::Mat& mask) const override;
    int GetKeypointOctave(const cv::KeyPoint& keypoint) const override;
    float GetKeypointSize(const cv::KeyPoint& keypoint) const override;
};
class FeatureExtractor_sift128 : public FeatureExtractor {          // include/Feature_sift128.h (constructor replaced: no GL context)
public:
    std::shared_ptr<SiftGPU> sift;
    FeatureExtractor_sift128(const int& nfeatures_, std::shared_ptr<FeatureExtractorSettings>& settings_) : FeatureExtractor(nfeatures_, settings_), sift(std::make_shared<SiftGPU>()) {}
    void detectAndCompute(const Image& img, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors) override;
    void detectKeypoints(std::map<int, std::vector<cv::KeyPoint>>& keypoints_level, const Image& img, const float& detectTh, const int& nOctaves) const override;
    void computeDescriptors(std::map<int, cv::Mat>& descriptors_level, std::map<int, std::vector<cv::KeyPoint>>& keypoints_level, const Image& img) const override;
    void filterKeypoints(std::map<int, std::vector<cv::KeyPoint>>& keypoints_level, const cv::Mat& image, const cv::Mat& mask) const override;
    int GetKeypointOctave(const cv::KeyPoint& keypoint) const override;
    float GetKeypointSize(const cv::KeyPoint& keypoint) const override;
};
class FeatureExtractor_akaze61 : public FeatureExtractor {          // include/Feature_akaze61.h (constructor replaced)
public:
    std::shared_ptr<libAKAZE::AKAZE> evolution{};
    AKAZEOptions akazeOptions{};
    FeatureExtractor_akaze61(const int& nfeatures_, std::shared_ptr<FeatureExtractorSettings>& settings_) : FeatureExtractor(nfeatures_, settings_), evolution(std::make_shared<libAKAZE::AKAZE>()) {}
    void detectAndCompute(const Image& img, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors) override;
    void detectKeypoints(std::map<int, std::vector<cv::KeyPoint>>& keypoints_level, const Image& img, const float& detectTh, const int& nOctaves) const override;
    void computeDescriptors(std::map<int, cv::Mat>& descriptors_level, std::map<int, std::vector<cv::KeyPoint>>& keypoints_level, const Image& img) const override;
    void filterKeypoints(std::map<int, std::vector<cv::KeyPoint>>& keypoints_level, const cv::Mat& image, const cv::Mat& mask) const override;
    int GetKeypointOctave(const cv::KeyPoint& keypoint) const override;
    float GetKeypointSize(const cv::KeyPoint& keypoint) const override;
};
}  // namespace ANYFEATURE_VSLAM

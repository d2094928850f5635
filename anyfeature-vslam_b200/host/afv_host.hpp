// afv_host.hpp -- C++ host mirror of the reference's feature front end API on top of the C ABI (include/afv.h).
//
// Same class names, method names, argument meaning and error behaviour as the reference:
//   FeatureExtractorSettings  include/FeatureExtractor.h:24-66, src/FeatureExtractor.cpp:21-56
//   FeatureExtractor          include/FeatureExtractor.h:68-161 (operator() 6-arg / 3-arg, getters, protected virtuals)
//   FeatureExtractor_orb32    include/Feature_orb32.h, src/Feature_orb32.cpp
//   FeatureExtractor_sift128  include/Feature_sift128.h, src/Feature_sift128.cpp (CV_32F N x 128, angle in radians)
//   FeatureExtractor_akaze61  include/Feature_akaze61.h, src/Feature_akaze61.cpp (CV_8U N x 61, octave := class_id)
//   FeatureExtractor_brisk48  include/Feature_brisk48.h, src/Feature_brisk48.cpp (CV_8U N x 48, octave = BRISK layer 0..7)
//   getFeatureExtractor       src/Tracking.cc:1505-1553 (factory, nfeatures clamp :1515-1520)
//   FeatureMatcher            include/FeatureMatcher.h:36-118: all 11 search methods (SearchForInitialization, SearchByProjection x4,
//                             SearchByBoW x2, SearchForTriangulation, SearchBySim3, Fuse x2), static DescriptorDistance,
//                             setDescriptorDistanceThresholds; TH_LOW/TH_HIGH statics
//   Frame helpers             UndistortKeyPoints / AssignFeaturesToGrid / isInFrustum (src/Frame.cc:403-433, :225-240, :276-331)
//   vanilla ORB-SLAM2         FeatureExtractor::operator()(..., vanillaOrbslam) + mvImagePyramid (include/FeatureExtractor.h:76-82, :142)
// Differences, all additive: the per-frame work runs in libafv_b200.so (CUDA, no CPU fallback: a missing device
// terminates like the reference's fatal paths, src/Feature_sift128.cpp:61); `Image` carries only the gray image;
// the matcher works on `FrameView`s (the arrays the reference's Frame owns: mvKeysUn, mDescriptors, keyPtsSize,
// image bounds, maxKeyPtSize) instead of the full Frame graph.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include "cv_compat.h"
#include "../../include/afv.h"

namespace ANYFEATURE_VSLAM_B200 {
using afv_host::mat2f;
using afvcv::KeyPoint;
using afvcv::Mat;

enum KeypointType { KEYP_ORB = 0, KEYP_AKAZE = 1, KEYP_BRISK = 2, KEYP_SIFT = 5 };         // include/Types.h:11-21
enum DescriptorType { DESC_ORB = 0, DESC_AKAZE61 = 1, DESC_BRISK = 2, DESC_SIFT128 = 5 };  // include/Types.h:23-33
enum FeatureType { FEAT_ORB = 0, FEAT_AKAZE61 = 1, FEAT_BRISK = 2, FEAT_SIFT128 = 5 };     // include/Types.h:35-45
int get_feature_id(const std::string& str);                                                // include/Types.h:102-124
typedef float Descriptor_Distance_Type;                                                    // include/Types.h:127

struct Image {                     // include/Image.h:12-27 (gray image only; imread / cvtColor stay with the caller)
    Mat grayImg, mask;
};

class FeatureExtractorSettings {
public:
    static int numOctaves0; static float scaleFactor0; static float th0;
    // "none" keeps the static nominal values of the previous constructor (reference quirk, :26-38)
    FeatureExtractorSettings(const KeypointType& keypointType_, const DescriptorType& descriptorType_, const std::string& settingsYamlFile);
    static float GetDetectorNominalScaleFactor() { return scaleFactor0; }
    static int GetDetectorNominalNumOctaves() { return numOctaves0; }
    static float GetDetectorNominalThreshold() { return th0; }
    bool ON_automaticTuning; float scaleFactor; int nOctaves; int iniThFAST, minThFAST; float detectTh;
    KeypointType keypointType; DescriptorType descriptorType;
    float maxKeyPtSize{0.0f}, minKeyPtSize{1.0f}, maxKeyPtSize0{}, maxKeyPtSigma0{};
    float scaleFactorOrb{1.2f}; int nOctavesOrb{8};
};

class FeatureExtractor {
public:
    FeatureExtractor(const int& nfeatures_, std::shared_ptr<FeatureExtractorSettings>& settings_);
    virtual ~FeatureExtractor();
    std::shared_ptr<FeatureExtractorSettings> settings{};
    void operator()(const Image& img, std::vector<KeyPoint>& keypoints, Mat& descriptors, std::vector<mat2f>& keyPtsSigma2,
                    std::vector<mat2f>& keyPtsInf, std::vector<float>& keyPtsSize);
    void operator()(const Image& img, std::vector<KeyPoint>& keypoints, Mat& descriptors);
    // VANILLA ORB-SLAM2 form (include/FeatureExtractor.h:76-82, src/ORBextractor.cc:568-645; Frame::ExtractFeatures src/Frame.cc:245-246):
    // runs the AFV_FEAT_ORB32_VANILLA extractor with this object's nfeatures / settings and fills mvImagePyramid
    void operator()(const Image& img, std::vector<KeyPoint>& keypoints, Mat& descriptors, std::vector<mat2f>& keyPtsSigma2,
                    std::vector<mat2f>& keyPtsInf, std::vector<float>& keyPtsSize, const bool& vanillaOrbslam);
    // include/FeatureExtractor.h:142: sized nOctaves by the constructor; only the vanilla path fills it (ComputePyramid), exactly
    // like the reference (the default build never writes it)
    std::vector<Mat> mvImagePyramid;
    // batched extension (not in the reference): B frames of equal size in one call
    void extractBatch(const std::vector<const Image*>& imgs, std::vector<std::vector<KeyPoint>>& keypoints, std::vector<Mat>& descriptors,
                      std::vector<std::vector<float>>& sizes);
    int GetLevels() { return settings->nOctaves; }
    float GetScaleFactor() { return settings->scaleFactor; }
    std::vector<float> GetScaleFactors() { return mvScaleFactor; }
    float GetMaxKeyPtSize() const { return settings->maxKeyPtSize0; }
    float GetMaxKeyPtSigma() const { return settings->maxKeyPtSigma0; }
    const std::vector<int>& GetFeaturesPerLevel() const { return mnFeaturesPerLevel; }
protected:
    int nfeatures;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor;
    void computeSigma(std::vector<mat2f>& keyPtsSigma2, std::vector<mat2f>& keyPtsInf, const std::vector<float>& keyPtsSize);
    void automaticTuning(const Image& img);
    // Functions to override (same hooks as the reference, include/FeatureExtractor.h:114-130).  detectAndCompute is ONE C-ABI call that
    // performs detectKeypoints + filterKeypoints + computeDescriptors + mergeKeypointLevels on the device; the per-stage hooks keep the
    // reference's names and signatures so a subclass written against the reference still compiles and can override them: the base
    // versions expose the device results of the LAST detectAndCompute per level where a stage tap exists (orb32: cv::ORB::detect list
    // and octree-kept list, without the angle cv::ORB fills) and are otherwise no-ops the ABI call has already satisfied.
    virtual void initializeExtractor(const Image&) {}
    virtual void detectKeypoints(std::map<int, std::vector<KeyPoint>>& keypoints_level, const Image& img, const float& detectTh, const int& nOctaves) const;
    virtual void filterKeypoints(std::map<int, std::vector<KeyPoint>>& keypoints_level, const Mat& image, const Mat& mask) const;
    virtual void computeDescriptors(std::map<int, Mat>&, std::map<int, std::vector<KeyPoint>>&, const Image&) const {}
    virtual void mergeKeypointLevels(std::vector<KeyPoint>&, Mat&, std::map<int, Mat>&, std::map<int, std::vector<KeyPoint>>&) const {}
    virtual void scaleKeypoints(std::map<int, std::vector<KeyPoint>>&) const {}
    void levelListFromTap(int what, std::map<int, std::vector<KeyPoint>>& keypoints_level) const;
    afv_extractor* vanilla_handle_ = nullptr; int vanilla_w_ = 0, vanilla_h_ = 0;
    virtual int GetKeypointOctave(const KeyPoint& keypoint) const = 0;
    virtual float GetKeypointSize(const KeyPoint& keypoint) const = 0;
    virtual void detectAndCompute(const Image& img, std::vector<KeyPoint>& keypoints, Mat& descriptors, std::vector<float>& sizes) = 0;
    // descriptor layout of the subclass: C-ABI feature id, row width, element type (CV_8U / CV_32F)
    virtual int featureId() const = 0;
    virtual int descCols() const = 0;
    virtual int descType() const { return afvcv::CV_8U; }
    void detectAndComputeABI(const Image& img, std::vector<KeyPoint>& keypoints, Mat& descriptors, std::vector<float>& sizes);
    afv_extractor* handle_ = nullptr;
    int handle_w_ = 0, handle_h_ = 0, handle_batch_ = 0;
    void ensureHandle(int feature_id, int w, int h, int batch);
};

class FeatureExtractor_orb32 : public FeatureExtractor {
public:
    FeatureExtractor_orb32(const int& nfeatures_, std::shared_ptr<FeatureExtractorSettings>& settings_) : FeatureExtractor(nfeatures_, settings_) {}
protected:
    int GetKeypointOctave(const KeyPoint& keypoint) const override { return keypoint.octave; }
    float GetKeypointSize(const KeyPoint& keypoint) const override;
    void detectAndCompute(const Image& img, std::vector<KeyPoint>& keypoints, Mat& descriptors, std::vector<float>& sizes) override { detectAndComputeABI(img, keypoints, descriptors, sizes); }
    int featureId() const override { return AFV_FEAT_ORB32; }
    int descCols() const override { return 32; }
};

class FeatureExtractor_sift128 : public FeatureExtractor {           // src/Feature_sift128.cpp:9-134
public:
    FeatureExtractor_sift128(const int& nfeatures_, std::shared_ptr<FeatureExtractorSettings>& settings_) : FeatureExtractor(nfeatures_, settings_) {}
protected:
    int GetKeypointOctave(const KeyPoint& keypoint) const override { return keypoint.octave; }                       // :120-122
    float GetKeypointSize(const KeyPoint& keypoint) const override;                                                  // :124-126
    void detectAndCompute(const Image& img, std::vector<KeyPoint>& keypoints, Mat& descriptors, std::vector<float>& sizes) override { detectAndComputeABI(img, keypoints, descriptors, sizes); }
    int featureId() const override { return AFV_FEAT_SIFT128; }
    int descCols() const override { return 128; }
    int descType() const override { return afvcv::CV_32F; }
};

class FeatureExtractor_akaze61 : public FeatureExtractor {           // src/Feature_akaze61.cpp:7-77
public:
    FeatureExtractor_akaze61(const int& nfeatures_, std::shared_ptr<FeatureExtractorSettings>& settings_) : FeatureExtractor(nfeatures_, settings_) {}
protected:
    int GetKeypointOctave(const KeyPoint& keypoint) const override { return keypoint.class_id; }                     // :63-65
    float GetKeypointSize(const KeyPoint& keypoint) const override;                                                  // :67-69
    void detectAndCompute(const Image& img, std::vector<KeyPoint>& keypoints, Mat& descriptors, std::vector<float>& sizes) override { detectAndComputeABI(img, keypoints, descriptors, sizes); }
    int featureId() const override { return AFV_FEAT_AKAZE61; }
    int descCols() const override { return 61; }
};

class FeatureExtractor_brisk48 : public FeatureExtractor {           // src/Feature_brisk48.cpp:7-64 (CV_8U N x 48, octave = BRISK layer)
public:
    FeatureExtractor_brisk48(const int& nfeatures_, std::shared_ptr<FeatureExtractorSettings>& settings_) : FeatureExtractor(nfeatures_, settings_) {}
protected:
    int GetKeypointOctave(const KeyPoint& keypoint) const override { return keypoint.octave; }                       // :50-52
    float GetKeypointSize(const KeyPoint& keypoint) const override;                                                  // :54-56
    void detectAndCompute(const Image& img, std::vector<KeyPoint>& keypoints, Mat& descriptors, std::vector<float>& sizes) override { detectAndComputeABI(img, keypoints, descriptors, sizes); }
    int featureId() const override { return AFV_FEAT_BRISK48; }
    int descCols() const override { return 48; }
};

// Tracking::getFeatureExtractor (src/Tracking.cc:1505-1553): nfeatures scaled with resolution, clamped to [1000,2000]
std::shared_ptr<FeatureExtractor> getFeatureExtractor(const int& scaleNumFeaturesMonocular, const std::string& feature_settings_yaml_file,
                                                      const std::string& feature, int imWidth, int imHeight);

// the arrays a reference Frame / KeyFrame hands to the matcher
struct FrameView {
    std::vector<KeyPoint> mvKeysUn; Mat mDescriptors; std::vector<float> keyPtsSize;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0, maxKeyPtSize = 0;
    float sizeTolerance = 1.2f;                       // Frame::sizeTolerance (src/Frame.cc:182-183)
    std::vector<uint8_t> hasMapPoint;                 // pts[i] != nullptr (with observations): keypoint i already holds a map point
    std::vector<int> featNode;                        // mFeatVec as the node id of every feature (Vocabulary::transform; < 0 = none)
    std::vector<float> inf_1d, sigma2_1d;             // GetKeyPt1DInf / GetKeyPt1DSigma2
    std::vector<int> gridCellStart, gridCellItems;    // mGrid as CSR, filled by AssignFeaturesToGrid
};
struct KeyFrameView : FrameView {};                   // lets the (KF, F) / (KF, KF) overloads of the reference keep their shape
// What the projection prologue of a search leaves per map point (the few lines that project the point and predict its size stay with
// the caller or come from isInFrustum): descriptor, projected position, search radius (< 0 = skipped by the prologue), accepted
// keypoint-size range, and the angle of the source keypoint where the search checks orientation.
struct ProjectedPoints {
    Mat descriptors; std::vector<afvcv::Point2f> uv; std::vector<float> radius, minSize, maxSize, angle;
};
// map points of the local map as isInFrustum reads them (src/Frame.cc:276-331, src/MapPoint.cc:432-442)
struct MapPointsView {
    std::vector<float> worldPos, normal;              // M x 3
    std::vector<float> minDistance, maxDistance, refSize, refSigma, refDistance;
};
struct PoseView { float Rcw[9]; float tcw[3]; float twc[3]; float fx, fy, cx, cy, mbf; };

// Frame::UndistortKeyPoints (src/Frame.cc:403-433): K = {fx, fy, cx, cy}, distCoef = {k1, k2, p1, p2, k3}
void UndistortKeyPoints(const std::vector<KeyPoint>& mvKeys, const float K[4], const float distCoef[5], std::vector<KeyPoint>& mvKeysUn);
// Frame::AssignFeaturesToGrid (src/Frame.cc:225-240): fills F.gridCellStart (64*48+1) / F.gridCellItems
void AssignFeaturesToGrid(FrameView& F);
// Frame::isInFrustum (src/Frame.cc:276-331) for the whole local map + the window prologue of SearchByProjection(F, points, th)
// (src/FeatureMatcher.cc:86-95, radius = radiusScale * radiusTh * RadiusByViewingCos * predicted size); descriptors are copied through
void isInFrustum(const FrameView& F, const PoseView& pose, const MapPointsView& points, const Mat& pointDescriptors, float viewingCosLimit,
                 float radiusFactor, std::vector<uint8_t>& mbTrackInView, ProjectedPoints& out, std::vector<float>& trackViewCos);

class FeatureMatcher {
public:
    FeatureMatcher(float nnratio = 0.6f, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
    static Descriptor_Distance_Type DescriptorDistance(const Mat& a, const Mat& b, const DescriptorType& descriptorType_);
    int SearchForInitialization(FrameView& F1, FrameView& F2, std::vector<afvcv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                const int& windowSize, const DescriptorType& descriptorType);
    // ---- the other ten searches (include/FeatureMatcher.h:47-66) after their projection prologue.  vnMatch[i] = matched keypoint of
    // query i or -1; searches that claim (every SearchByProjection) also mark F.hasMapPoint like the reference sets F.pts[idx].
    enum ProjectionVariant { TRACK_LOCAL_MAP,         // SearchByProjection(Frame&, vector<Pt>&, th)               src/FeatureMatcher.cc:73-154
                             SIM3,                    // SearchByProjection(pKF, Scw, points, vpMatched, th)        :287-397
                             MOTION_MODEL,            // SearchByProjection(CurrentFrame, LastFrame, th, bMono)     :1291-1402
                             RELOCALISATION };        // SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, high) :1406-1506
    int SearchByProjection(ProjectionVariant variant, FrameView& F, const ProjectedPoints& vpMapPoints, std::vector<int>& vnMatch,
                           const DescriptorType& descriptorType);
    int Fuse(KeyFrameView& pKF, const ProjectedPoints& vpMapPoints, bool monoReprojectionGate, std::vector<int>& vnMatch,
             const DescriptorType& descriptorType);   // :794-942 (gate on) / :944-1064 (gate off); add / replace stays with the caller
    int SearchBySim3(KeyFrameView& pKF1, KeyFrameView& pKF2, const ProjectedPoints& points1in2, const ProjectedPoints& points2in1,
                     std::vector<int>& vnMatches12, const DescriptorType& descriptorType);                       // :1066-1287
    int SearchByBoW(KeyFrameView& pKF, FrameView& F, std::vector<int>& vnMatchF, const DescriptorType& descriptorType);          // :186-283
    int SearchByBoW(KeyFrameView& pKF1, KeyFrameView& pKF2, std::vector<int>& vnMatches12, const DescriptorType& descriptorType); // :561-660
    int SearchForTriangulation(KeyFrameView& pKF1, KeyFrameView& pKF2, const float F12[9], const float epipole2[2],
                               std::vector<std::pair<size_t, size_t>>& vMatchedPairs, const DescriptorType& descriptorType);     // :662-790
    static void setDescriptorDistanceThresholds(const std::string& feature_settings_yaml_file);
    static Descriptor_Distance_Type TH_LOW, TH_HIGH, descDistTh_high_reloc, descDistTh_low_reloc;
    static const int HISTO_LENGTH;
protected:
    float mfNNratio; bool mbCheckOrientation;
    int projectionCore(FrameView& F, const ProjectedPoints& P, bool useOccupied, bool claim, bool ratioSameScale, bool useAngle, bool useInf,
                       float th, std::vector<int>& vnMatch, int descType);
    int bowCore(int mode, FrameView& A, FrameView& B, const float* F12, const float* epipole, std::vector<int>& match, int descType);
};
}  // namespace ANYFEATURE_VSLAM_B200

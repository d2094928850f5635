// afv_host.cpp -- implementation of the C++ host mirror (see afv_host.hpp). Plain C++17 + CUDA runtime API for the
// device buffers the matcher ABI needs; all compute is behind include/afv.h.
#include "afv_host.hpp"
#include <cuda_runtime_api.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <fstream>
#include <sstream>

namespace ANYFEATURE_VSLAM_B200 {

[[noreturn]] static void fatal(const char* what) {        // the reference terminates on unrecoverable set-up errors
    std::fprintf(stderr, "anyfeature-vslam_b200: %s: %s\n", what, afv_last_error());
    std::terminate();
}
#define AFV_OK_OR_DIE(x) do { if ((x) != AFV_OK) fatal(#x); } while (0)
#define CU_OK_OR_DIE(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); std::terminate(); } } while (0)

int get_feature_id(const std::string& str) {
    static const char* names[] = {"orb32", "akaze61", "brisk48", "surf64", "kaze64", "sift128", "r2d2_128", "anyfeatbin", "anyfeatnonbin"};
    for (int i = 0; i < 9; ++i)
        if (str == names[i]) return i;
    return 0;
}

// the 4-key settings files are flat "key: value" YAML (settings/<feat>_settings.yaml)
static bool yaml_value(const std::string& file, const std::string& key, float& out) {
    std::ifstream f(file);
    std::string line;
    while (std::getline(f, line)) {
        const size_t p = line.find(key + ":");
        if (p == std::string::npos || line.find('#') < p) continue;
        out = std::strtof(line.c_str() + p + key.size() + 1, nullptr);
        return true;
    }
    return false;
}

int FeatureExtractorSettings::numOctaves0 = 8;
float FeatureExtractorSettings::scaleFactor0 = 1.2f;
float FeatureExtractorSettings::th0 = 20.0f;

FeatureExtractorSettings::FeatureExtractorSettings(const KeypointType& keypointType_, const DescriptorType& descriptorType_, const std::string& settingsYamlFile)
    : keypointType(keypointType_), descriptorType(descriptorType_) {
    if (settingsYamlFile != "none") {
        float v;
        if (yaml_value(settingsYamlFile, "FeatureExtractor.numOctaves", v)) numOctaves0 = (int)v;
        if (yaml_value(settingsYamlFile, "FeatureExtractor.scaleFactor", v)) scaleFactor0 = v;
        if (yaml_value(settingsYamlFile, "FeatureExtractor.detectionTh", v)) th0 = v;
    }
    scaleFactor = GetDetectorNominalScaleFactor(); nOctaves = GetDetectorNominalNumOctaves(); detectTh = GetDetectorNominalThreshold();
    ON_automaticTuning = true; iniThFAST = 20; minThFAST = 7;
    maxKeyPtSize0 = std::pow(scaleFactorOrb, float(nOctavesOrb - 1.0)); maxKeyPtSigma0 = maxKeyPtSize0;
    maxKeyPtSize = maxKeyPtSize0; minKeyPtSize = 1.0f;
}

FeatureExtractor::FeatureExtractor(const int& nfeatures_, std::shared_ptr<FeatureExtractorSettings>& settings_) : settings(settings_), nfeatures(nfeatures_) {
    mvScaleFactor.resize(settings->nOctaves);
    mvScaleFactor[0] = 1.0f;
    for (int i = 1; i < settings->nOctaves; i++) mvScaleFactor[i] = mvScaleFactor[i - 1] * settings->scaleFactor;
    mnFeaturesPerLevel.assign(settings->nOctaves, 0);
    mvImagePyramid.resize(settings->nOctaves);                     // src/FeatureExtractor.cpp:93 (filled by the vanilla path only)
}
FeatureExtractor::~FeatureExtractor() { if (handle_) afv_extractor_destroy(handle_); if (vanilla_handle_) afv_extractor_destroy(vanilla_handle_); }

// per-level lists of the LAST extraction from the stage taps of the orb32 device pipeline (what = 3 cv::ORB::detect-equivalent
// list, 4 octree-kept list): level coordinates scaled to the full image, size 31 * scale, response = Harris, octave = level
void FeatureExtractor::levelListFromTap(int what, std::map<int, std::vector<KeyPoint>>& keypoints_level) const {
    keypoints_level.clear();
    if (!handle_ || featureId() != AFV_FEAT_ORB32) return;
    std::vector<uint32_t> buf(1 << 16);
    for (int l = 0; l < settings->nOctaves; ++l) {
        long nb = 0;
        AFV_OK_OR_DIE(afv_debug_read(handle_, what, 0, l, buf.data(), (long)(buf.size() * 4), &nb));
        const float scale = (float)std::pow((double)1.2f, (double)l);                  // cv::ORB's own scale factor (never overridden)
        std::vector<KeyPoint>& v = keypoints_level[l];
        for (long i = 0; i < nb / 8; ++i) {
            KeyPoint k; const uint32_t xy = buf[2 * i]; float r; std::memcpy(&r, &buf[2 * i + 1], 4);
            k.pt.x = (float)(xy & 0xfff) * scale; k.pt.y = (float)((xy >> 12) & 0xfff) * scale; k.size = 31 * scale; k.response = r; k.octave = l;
            v.push_back(k);
        }
    }
}
void FeatureExtractor::detectKeypoints(std::map<int, std::vector<KeyPoint>>& keypoints_level, const Image&, const float&, const int&) const {
    levelListFromTap(3, keypoints_level);
}
void FeatureExtractor::filterKeypoints(std::map<int, std::vector<KeyPoint>>& keypoints_level, const Mat&, const Mat&) const {
    levelListFromTap(4, keypoints_level);
}

// src/ORBextractor.cc:568-645 (built with VANILLA_ORB_SLAM2): the settings constructor pins scaleFactor / nOctaves / thresholds there
// (src/FeatureExtractor.cpp:40-46); here the object's own settings are used and detectTh plays iniThFAST
void FeatureExtractor::operator()(const Image& img, std::vector<KeyPoint>& keypoints, Mat& descriptors, std::vector<mat2f>& keyPtsSigma2,
                                  std::vector<mat2f>& keyPtsInf, std::vector<float>& keyPtsSize, const bool& vanillaOrbslam) {
    if (!vanillaOrbslam) { (*this)(img, keypoints, descriptors, keyPtsSigma2, keyPtsInf, keyPtsSize); return; }
    const Mat& g = img.grayImg;
    if (g.empty()) return;                                                            // :571-572
    if (!vanilla_handle_ || g.cols > vanilla_w_ || g.rows > vanilla_h_) {
        if (vanilla_handle_) afv_extractor_destroy(vanilla_handle_);
        vanilla_handle_ = nullptr;
        AFV_OK_OR_DIE(afv_extractor_create(&vanilla_handle_, AFV_FEAT_ORB32_VANILLA, nfeatures, settings->nOctaves, settings->scaleFactor,
                                           (float)settings->iniThFAST, 0, 1, g.cols, g.rows));
        vanilla_w_ = g.cols; vanilla_h_ = g.rows;
    }
    const int cap = afv_extractor_output_cap(vanilla_handle_);
    keypoints.assign(cap, KeyPoint());
    Mat d(cap, 32, afvcv::CV_8U);
    keyPtsSize.assign(cap, 0.f);
    int n = 0;
    AFV_OK_OR_DIE(afv_extract(vanilla_handle_, g.data(), g.cols, g.rows, (int)g.step(), reinterpret_cast<afv_keypoint*>(keypoints.data()), d.data(),
                              keyPtsSize.data(), cap, &n));
    keypoints.resize(n); keyPtsSize.resize(n);
    if (n == 0) descriptors.release();                                               // :587-588
    else { descriptors.create(n, 32, afvcv::CV_8U); std::memcpy(descriptors.data(), d.data(), (size_t)n * 32); }
    computeSigma(keyPtsSigma2, keyPtsInf, keyPtsSize);
    // mvImagePyramid (ComputePyramid :647-674): level 0 = the input, levels >= 1 from the device pyramid
    mvImagePyramid.resize(settings->nOctaves);
    for (int l = 0; l < settings->nOctaves; ++l) {
        const float inv = 1.0f / mvScaleFactor[l];
        const int lw = (int)std::lrintf((float)g.cols * inv), lh = (int)std::lrintf((float)g.rows * inv);
        mvImagePyramid[l].create(lh, lw, afvcv::CV_8U);
        if (l == 0) { std::memcpy(mvImagePyramid[0].data(), g.data(), (size_t)lw * lh); continue; }
        long nb = 0;
        AFV_OK_OR_DIE(afv_debug_read(vanilla_handle_, 40, 0, l, mvImagePyramid[l].data(), (long)lw * lh, &nb));
    }
}

void FeatureExtractor::ensureHandle(int feature_id, int w, int h, int batch) {
    if (handle_ && w <= handle_w_ && h <= handle_h_ && batch <= handle_batch_) return;
    if (handle_) afv_extractor_destroy(handle_);
    handle_ = nullptr;
    AFV_OK_OR_DIE(afv_extractor_create(&handle_, feature_id, nfeatures, settings->nOctaves, settings->scaleFactor, settings->detectTh, 0, batch, w, h));
    handle_w_ = w; handle_h_ = h; handle_batch_ = batch;
    std::vector<float> sf(settings->nOctaves);
    afv_extractor_levels(handle_, sf.data(), mnFeaturesPerLevel.data());
}

void FeatureExtractor::automaticTuning(const Image&) {         // src/FeatureExtractor.cpp:195-274: effectively detectTh = th0, flag off
    settings->detectTh = settings->GetDetectorNominalThreshold();
    settings->ON_automaticTuning = false;
}

void FeatureExtractor::computeSigma(std::vector<mat2f>& s2, std::vector<mat2f>& inf, const std::vector<float>& size) {   // :144-172, SIZE
    s2.clear(); inf.clear(); s2.reserve(size.size()); inf.reserve(size.size());
    for (float s : size) { const float v = s * s; s2.push_back(mat2f::scaledIdentity(v)); inf.push_back(mat2f::scaledIdentity(1.0f / v)); }
}

void FeatureExtractor::operator()(const Image& img, std::vector<KeyPoint>& keypoints, Mat& descriptors, std::vector<mat2f>& keyPtsSigma2,
                                  std::vector<mat2f>& keyPtsInf, std::vector<float>& keyPtsSize) {
    initializeExtractor(img);
    if (settings->ON_automaticTuning) automaticTuning(img);
    detectAndCompute(img, keypoints, descriptors, keyPtsSize);
    computeSigma(keyPtsSigma2, keyPtsInf, keyPtsSize);
}
void FeatureExtractor::operator()(const Image& img, std::vector<KeyPoint>& keypoints, Mat& descriptors) {
    std::vector<float> sizes;
    initializeExtractor(img);
    if (settings->ON_automaticTuning) automaticTuning(img);
    detectAndCompute(img, keypoints, descriptors, sizes);
}

float FeatureExtractor_orb32::GetKeypointSize(const KeyPoint& keypoint) const {
    return powf(settings->GetDetectorNominalScaleFactor(), float(GetKeypointOctave(keypoint)));
}

float FeatureExtractor_sift128::GetKeypointSize(const KeyPoint& keypoint) const {
    return powf(settings->GetDetectorNominalScaleFactor(), float(GetKeypointOctave(keypoint)));
}
float FeatureExtractor_akaze61::GetKeypointSize(const KeyPoint& keypoint) const {
    return powf(settings->GetDetectorNominalScaleFactor(), float(GetKeypointOctave(keypoint)));
}
float FeatureExtractor_brisk48::GetKeypointSize(const KeyPoint& keypoint) const {
    return powf(settings->GetDetectorNominalScaleFactor(), float(GetKeypointOctave(keypoint)));
}

// detectKeypoints + filterKeypoints + computeDescriptors + mergeKeypointLevels + computeSize of the subclass as ONE C-ABI call
void FeatureExtractor::detectAndComputeABI(const Image& img, std::vector<KeyPoint>& keypoints, Mat& descriptors, std::vector<float>& sizes) {
    const Mat& g = img.grayImg;
    ensureHandle(featureId(), g.cols, g.rows, 1);
    const int cap = afv_extractor_output_cap(handle_);
    const int dc = descCols(), dt = descType();
    keypoints.assign(cap, KeyPoint());
    Mat d(cap, dc, dt);
    sizes.assign(cap, 0.f);
    int n = 0;
    AFV_OK_OR_DIE(afv_extract(handle_, g.data(), g.cols, g.rows, (int)g.step(), reinterpret_cast<afv_keypoint*>(keypoints.data()), d.data(), sizes.data(), cap, &n));
    keypoints.resize(n); sizes.resize(n);
    descriptors.create(n, dc, dt);
    std::memcpy(descriptors.data(), d.data(), (size_t)n * d.step());
}

void FeatureExtractor::extractBatch(const std::vector<const Image*>& imgs, std::vector<std::vector<KeyPoint>>& keypoints, std::vector<Mat>& descriptors,
                                    std::vector<std::vector<float>>& sizes) {
    const int B = (int)imgs.size();
    if (!B) return;
    const int w = imgs[0]->grayImg.cols, h = imgs[0]->grayImg.rows;
    ensureHandle(featureId(), w, h, B);
    const int cap = afv_extractor_output_cap(handle_);
    const size_t db = (size_t)descCols() * (descType() == afvcv::CV_32F ? 4 : 1);
    std::vector<uint8_t> gray((size_t)B * w * h);
    for (int b = 0; b < B; ++b) std::memcpy(gray.data() + (size_t)b * w * h, imgs[b]->grayImg.data(), (size_t)w * h);
    std::vector<KeyPoint> k((size_t)B * cap); std::vector<uint8_t> d((size_t)B * cap * db); std::vector<float> s((size_t)B * cap); std::vector<int> n(B);
    AFV_OK_OR_DIE(afv_extract_batch(handle_, gray.data(), B, w, h, w, (long)w * h, reinterpret_cast<afv_keypoint*>(k.data()), d.data(), s.data(), cap, n.data()));
    keypoints.resize(B); descriptors.resize(B); sizes.resize(B);
    for (int b = 0; b < B; ++b) {
        keypoints[b].assign(k.begin() + (size_t)b * cap, k.begin() + (size_t)b * cap + n[b]);
        sizes[b].assign(s.begin() + (size_t)b * cap, s.begin() + (size_t)b * cap + n[b]);
        descriptors[b].create(n[b], descCols(), descType());
        std::memcpy(descriptors[b].data(), d.data() + (size_t)b * cap * db, (size_t)n[b] * db);
    }
}

std::shared_ptr<FeatureExtractor> getFeatureExtractor(const int& scaleNumFeaturesMonocular, const std::string& yaml, const std::string& feature, int imWidth, int imHeight) {
    // src/Tracking.cc:1514-1520: linear interpolation between 1000 features at 640x480 and 2000 at 1241x376 (double
    // arithmetic on a float pixel count, truncated to int), clamped to [1000, 2000], then scaled
    const int numFeatures0 = 1000;
    const int w = imWidth, h = imHeight;
    int nFeatures = ((2000.0 - 1000.0) / (1241.0 * 376.0 - 640.0 * 480.0)) * (float(w * h) - 640.0 * 480.0) + numFeatures0;
    if (nFeatures > 2000) nFeatures = 2000;
    else if (nFeatures < 1000) nFeatures = 1000;
    nFeatures *= scaleNumFeaturesMonocular;
    const int id = get_feature_id(feature);
    auto settings = std::make_shared<FeatureExtractorSettings>((KeypointType)id, (DescriptorType)id, yaml);
    switch (id) {
        case FEAT_ORB: return std::make_shared<FeatureExtractor_orb32>(nFeatures, settings);
        case FEAT_SIFT128: return std::make_shared<FeatureExtractor_sift128>(nFeatures, settings);
        case FEAT_AKAZE61: return std::make_shared<FeatureExtractor_akaze61>(nFeatures, settings);
        case FEAT_BRISK: return std::make_shared<FeatureExtractor_brisk48>(nFeatures, settings);
        default:
            std::fprintf(stderr, "getFeatureExtractor: feature '%s' has no B200 extractor (orb32, akaze61, brisk48, sift128 are built)\n", feature.c_str());
            std::terminate();                                                         // include/Types.h:67-70 behaviour
    }
}

// ---------------------------------------------------------------- matcher ---------------------------------
Descriptor_Distance_Type FeatureMatcher::TH_HIGH = 0.0f, FeatureMatcher::TH_LOW = 0.0f;
Descriptor_Distance_Type FeatureMatcher::descDistTh_high_reloc = 0.0f, FeatureMatcher::descDistTh_low_reloc = 0.0f;
const int FeatureMatcher::HISTO_LENGTH = 30;

void FeatureMatcher::setDescriptorDistanceThresholds(const std::string& yaml) {        // src/FeatureMatcher.cc:1533-1545
    float th = 0;
    if (!yaml_value(yaml, "FeatureMatcher.matchingTh", th)) { std::fprintf(stderr, "missing FeatureMatcher.matchingTh in %s\n", yaml.c_str()); std::terminate(); }
    TH_LOW = th; TH_HIGH = TH_LOW; descDistTh_low_reloc = TH_LOW; descDistTh_high_reloc = TH_LOW;
}

template <typename T> struct DevBuf {
    T* p = nullptr; size_t n = 0;
    explicit DevBuf(size_t n_) : n(n_) { CU_OK_OR_DIE(cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T))); }
    ~DevBuf() { cudaFree(p); }
    void up(const void* h, size_t cnt) { CU_OK_OR_DIE(cudaMemcpy(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice)); }
    void down(void* h, size_t cnt) { CU_OK_OR_DIE(cudaMemcpy(h, p, cnt * sizeof(T), cudaMemcpyDeviceToHost)); }
};

Descriptor_Distance_Type FeatureMatcher::DescriptorDistance(const Mat& a, const Mat& b, const DescriptorType& t) {
    const size_t bytes = (size_t)a.cols * a.elemSize();
    DevBuf<uint8_t> da(bytes), db(bytes); DevBuf<float> out(1);
    da.up(a.data(), bytes); db.up(b.data(), bytes);
    AFV_OK_OR_DIE(afv_descriptor_distance((int)t, da.p, db.p, 1, out.p, nullptr));
    float r = 0; out.down(&r, 1);
    return r;
}

int FeatureMatcher::SearchForInitialization(FrameView& F1, FrameView& F2, std::vector<afvcv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                            const int& windowSize, const DescriptorType& descriptorType) {
    const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
    const int cap = std::max(std::max(n1, n2), 1);
    const size_t D = (size_t)F1.mDescriptors.cols * F1.mDescriptors.elemSize();
    DevBuf<afv_keypoint> dk((size_t)2 * cap); DevBuf<uint8_t> dd((size_t)2 * cap * D); DevBuf<float> ds((size_t)2 * cap), dpm((size_t)cap * 2);
    DevBuf<int> dn(2), dpa(1), dpb(1), dm((size_t)cap), dnm(1);
    std::vector<float> s1(n1, 1.0f);
    dk.up(F1.mvKeysUn.data(), n1); CU_OK_OR_DIE(cudaMemcpy(dk.p + cap, F2.mvKeysUn.data(), (size_t)n2 * sizeof(afv_keypoint), cudaMemcpyHostToDevice));
    dd.up(F1.mDescriptors.data(), (size_t)n1 * D); CU_OK_OR_DIE(cudaMemcpy(dd.p + (size_t)cap * D, F2.mDescriptors.data(), (size_t)n2 * D, cudaMemcpyHostToDevice));
    ds.up(F1.keyPtsSize.empty() ? s1.data() : F1.keyPtsSize.data(), n1);
    CU_OK_OR_DIE(cudaMemcpy(ds.p + cap, F2.keyPtsSize.data(), (size_t)n2 * sizeof(float), cudaMemcpyHostToDevice));
    const int nn[2] = {n1, n2}, a = 0, b = 1;
    dn.up(nn, 2); dpa.up(&a, 1); dpb.up(&b, 1);
    dpm.up(vbPrevMatched.data(), (size_t)n1 * 2);
    AFV_OK_OR_DIE(afv_search_for_initialization((int)descriptorType, dk.p, dd.p, ds.p, dn.p, 2, cap, dpa.p, dpb.p, 1, F2.mnMinX, F2.mnMinY, F2.mnMaxX, F2.mnMaxY,
                                                F1.maxKeyPtSize, dpm.p, windowSize, TH_LOW, mfNNratio, mbCheckOrientation ? 1 : 0, dm.p, dnm.p, nullptr));
    vnMatches12.assign(n1, -1);
    dm.down(vnMatches12.data(), n1);
    dpm.down(vbPrevMatched.data(), (size_t)n1 * 2);
    int nmatches = 0; dnm.down(&nmatches, 1);
    return nmatches;
}

// ---------------------------------------------------------------- Frame helpers -------------------------------------------------
void UndistortKeyPoints(const std::vector<KeyPoint>& mvKeys, const float K[4], const float distCoef[5], std::vector<KeyPoint>& mvKeysUn) {
    const int n = (int)mvKeys.size();
    mvKeysUn.assign(n, KeyPoint());
    if (!n) return;
    DevBuf<afv_keypoint> a(n), b(n); DevBuf<int> dn(1);
    a.up(mvKeys.data(), n); dn.up(&n, 1);
    AFV_OK_OR_DIE(afv_undistort_keypoints(a.p, dn.p, 1, n, K, distCoef, b.p, nullptr));
    b.down(mvKeysUn.data(), n);
}

void AssignFeaturesToGrid(FrameView& F) {
    const int n = (int)F.mvKeysUn.size(), cap = std::max(n, 1);
    DevBuf<afv_keypoint> k(cap); DevBuf<int> dn(1), cs(64 * 48 + 1), ci(cap);
    k.up(F.mvKeysUn.data(), n); dn.up(&n, 1);
    AFV_OK_OR_DIE(afv_grid_build(k.p, dn.p, 1, cap, F.mnMinX, F.mnMinY, F.mnMaxX, F.mnMaxY, cs.p, ci.p, nullptr));
    F.gridCellStart.assign(64 * 48 + 1, 0); F.gridCellItems.assign(cap, 0);
    cs.down(F.gridCellStart.data(), 64 * 48 + 1); ci.down(F.gridCellItems.data(), cap);
    F.gridCellItems.resize(F.gridCellStart.back());
}

void isInFrustum(const FrameView& F, const PoseView& pose, const MapPointsView& pts, const Mat& pointDescriptors, float viewingCosLimit,
                 float radiusFactor, std::vector<uint8_t>& mbTrackInView, ProjectedPoints& out, std::vector<float>& trackViewCos) {
    const int M = (int)pts.minDistance.size();
    mbTrackInView.assign(M, 0); trackViewCos.assign(M, 0.f);
    out.descriptors = pointDescriptors; out.uv.assign(M, afvcv::Point2f()); out.radius.assign(M, -1.f); out.minSize.assign(M, 0.f); out.maxSize.assign(M, 0.f);
    out.angle.clear();
    if (!M) return;
    DevBuf<float> dP((size_t)3 * M), dN((size_t)3 * M), dmin(M), dmax(M), drs(M), drg(M), drd(M), dproj((size_t)3 * M), dtrack((size_t)3 * M), dqr(M), dqmin(M), dqmax(M);
    DevBuf<uint8_t> div(M);
    dP.up(pts.worldPos.data(), (size_t)3 * M); dN.up(pts.normal.data(), (size_t)3 * M); dmin.up(pts.minDistance.data(), M); dmax.up(pts.maxDistance.data(), M);
    drs.up(pts.refSize.data(), M); drg.up(pts.refSigma.data(), M); drd.up(pts.refDistance.data(), M);
    float pose16[16] = {0}, cam5[5] = {pose.fx, pose.fy, pose.cx, pose.cy, pose.mbf}, bounds4[4] = {F.mnMinX, F.mnMaxX, F.mnMinY, F.mnMaxY};
    std::memcpy(pose16, pose.Rcw, 36); std::memcpy(pose16 + 9, pose.tcw, 12); std::memcpy(pose16 + 12, pose.twc, 12);
    AFV_OK_OR_DIE(afv_is_in_frustum(dP.p, dN.p, dmin.p, dmax.p, drs.p, drg.p, drd.p, M, pose16, cam5, bounds4, viewingCosLimit, radiusFactor,
                                    F.sizeTolerance, div.p, dproj.p, dtrack.p, dqr.p, dqmin.p, dqmax.p, nullptr));
    std::vector<float> proj((size_t)3 * M), track((size_t)3 * M);
    div.down(mbTrackInView.data(), M); dproj.down(proj.data(), (size_t)3 * M); dtrack.down(track.data(), (size_t)3 * M);
    dqr.down(out.radius.data(), M); dqmin.down(out.minSize.data(), M); dqmax.down(out.maxSize.data(), M);
    for (int i = 0; i < M; ++i) { out.uv[i] = afvcv::Point2f(proj[3 * i], proj[3 * i + 1]); trackViewCos[i] = track[3 * i + 2]; }
}

// ---------------------------------------------------------------- the remaining FeatureMatcher searches -------------------------
// one frame (B = 1) on the device: keypoints, descriptors, sizes, count
struct DevFrame {
    DevBuf<afv_keypoint> k; DevBuf<uint8_t> d; DevBuf<float> s; DevBuf<int> n; int cnt, cap;
    DevFrame(const FrameView& F, size_t D) : k(std::max<size_t>(F.mvKeysUn.size(), 1)), d(std::max<size_t>(F.mvKeysUn.size(), 1) * D),
                                             s(std::max<size_t>(F.mvKeysUn.size(), 1)), n(1), cnt((int)F.mvKeysUn.size()), cap(std::max((int)F.mvKeysUn.size(), 1)) {
        std::vector<float> one(cnt, 1.0f);
        k.up(F.mvKeysUn.data(), cnt); d.up(F.mDescriptors.data(), (size_t)cnt * D);
        s.up(F.keyPtsSize.size() == (size_t)cnt ? F.keyPtsSize.data() : one.data(), cnt); n.up(&cnt, 1);
    }
};

int FeatureMatcher::projectionCore(FrameView& F, const ProjectedPoints& P, bool useOccupied, bool claim, bool ratioSameScale, bool useAngle,
                                   bool useInf, float th, std::vector<int>& vnMatch, int descType) {
    const int nq = (int)P.uv.size();
    vnMatch.assign(nq, -1);
    if (!nq || F.mvKeysUn.empty()) return 0;
    const size_t D = (size_t)F.mDescriptors.cols * F.mDescriptors.elemSize();
    DevFrame T(F, D);
    DevBuf<uint8_t> qd((size_t)nq * D), occ(T.cap); DevBuf<float> qxy((size_t)2 * nq), qr(nq), qmin(nq), qmax(nq), qang(nq), inf(T.cap);
    DevBuf<int> qs(2), fr(1), mq(nq), nm(1);
    qd.up(P.descriptors.data(), (size_t)nq * D); qxy.up(P.uv.data(), (size_t)2 * nq); qr.up(P.radius.data(), nq); qmin.up(P.minSize.data(), nq); qmax.up(P.maxSize.data(), nq);
    if (useAngle) qang.up(P.angle.data(), nq);
    if (useInf) inf.up(F.inf_1d.data(), T.cnt);
    std::vector<uint8_t> o(T.cap, 0);
    if (useOccupied && F.hasMapPoint.size() == (size_t)T.cnt) std::copy(F.hasMapPoint.begin(), F.hasMapPoint.end(), o.begin());
    occ.up(o.data(), T.cap);
    const int starts[2] = {0, nq}, zero = 0;
    qs.up(starts, 2); fr.up(&zero, 1);
    AFV_OK_OR_DIE(afv_search_by_projection_ex(descType, qd.p, qxy.p, qr.p, qmin.p, qmax.p, useAngle ? qang.p : nullptr, qs.p, 1, nq, T.k.p, T.d.p, T.s.p,
                                              useInf ? inf.p : nullptr, T.n.p, 1, T.cap, fr.p, useOccupied ? occ.p : nullptr, claim ? 1 : 0, F.mnMinX, F.mnMinY,
                                              F.mnMaxX, F.mnMaxY, th, mfNNratio, ratioSameScale ? 1 : 0, F.sizeTolerance, mq.p, nm.p, nullptr, 0, nullptr));
    mq.down(vnMatch.data(), nq);
    int n = 0; nm.down(&n, 1);
    if (claim) { F.hasMapPoint.resize(T.cnt, 0); for (int v : vnMatch) if (v >= 0) F.hasMapPoint[v] = 1; }   // F.pts[bestIdx] = pMP
    return n;
}

int FeatureMatcher::SearchByProjection(ProjectionVariant variant, FrameView& F, const ProjectedPoints& P, std::vector<int>& vnMatch, const DescriptorType& dt) {
    switch (variant) {                                  // option table of include/afv.h (afv_search_by_projection_ex)
        case TRACK_LOCAL_MAP: return projectionCore(F, P, true, true, true, false, false, TH_HIGH, vnMatch, (int)dt);
        case SIM3: return projectionCore(F, P, true, true, false, false, false, TH_LOW, vnMatch, (int)dt);
        case MOTION_MODEL: return projectionCore(F, P, true, true, false, mbCheckOrientation, false, TH_HIGH, vnMatch, (int)dt);
        default: return projectionCore(F, P, true, true, false, mbCheckOrientation, false, descDistTh_high_reloc, vnMatch, (int)dt);
    }
}
int FeatureMatcher::Fuse(KeyFrameView& pKF, const ProjectedPoints& P, bool monoReprojectionGate, std::vector<int>& vnMatch, const DescriptorType& dt) {
    return projectionCore(pKF, P, false, false, false, false, monoReprojectionGate, TH_LOW, vnMatch, (int)dt);
}

int FeatureMatcher::SearchBySim3(KeyFrameView& K1, KeyFrameView& K2, const ProjectedPoints& P1, const ProjectedPoints& P2, std::vector<int>& m12,
                                 const DescriptorType& dt) {
    const int n1 = (int)K1.mvKeysUn.size(), n2 = (int)K2.mvKeysUn.size(), cap = std::max(std::max(n1, n2), 1);
    m12.assign(n1, -1);
    if (!n1 || !n2) return 0;
    const size_t D = (size_t)K1.mDescriptors.cols * K1.mDescriptors.elemSize();
    DevBuf<afv_keypoint> dk((size_t)2 * cap); DevBuf<uint8_t> dd((size_t)2 * cap * D); DevBuf<float> ds((size_t)2 * cap); DevBuf<int> dn(2);
    CU_OK_OR_DIE(cudaMemcpy(dk.p, K1.mvKeysUn.data(), (size_t)n1 * 28, cudaMemcpyHostToDevice)); CU_OK_OR_DIE(cudaMemcpy(dk.p + cap, K2.mvKeysUn.data(), (size_t)n2 * 28, cudaMemcpyHostToDevice));
    CU_OK_OR_DIE(cudaMemcpy(dd.p, K1.mDescriptors.data(), (size_t)n1 * D, cudaMemcpyHostToDevice)); CU_OK_OR_DIE(cudaMemcpy(dd.p + (size_t)cap * D, K2.mDescriptors.data(), (size_t)n2 * D, cudaMemcpyHostToDevice));
    CU_OK_OR_DIE(cudaMemcpy(ds.p, K1.keyPtsSize.data(), (size_t)n1 * 4, cudaMemcpyHostToDevice)); CU_OK_OR_DIE(cudaMemcpy(ds.p + cap, K2.keyPtsSize.data(), (size_t)n2 * 4, cudaMemcpyHostToDevice));
    const int nn[2] = {n1, n2}; dn.up(nn, 2);
    auto upq = [&](const ProjectedPoints& P, int n, DevBuf<uint8_t>& qd, DevBuf<float>& xy, DevBuf<float>& r, DevBuf<float>& mn, DevBuf<float>& mx, DevBuf<int>& st) {
        qd.up(P.descriptors.data(), (size_t)n * D); xy.up(P.uv.data(), (size_t)2 * n); r.up(P.radius.data(), n); mn.up(P.minSize.data(), n); mx.up(P.maxSize.data(), n);
        const int s2[2] = {0, n}; st.up(s2, 2);
    };
    DevBuf<uint8_t> q1d((size_t)n1 * D), q2d((size_t)n2 * D); DevBuf<float> x1((size_t)2 * n1), r1(n1), a1(n1), b1(n1), x2((size_t)2 * n2), r2(n2), a2(n2), b2(n2);
    DevBuf<int> s1(2), s2(2), f1(1), f2(1), dm(n1), nf(1);
    upq(P1, n1, q1d, x1, r1, a1, b1, s1); upq(P2, n2, q2d, x2, r2, a2, b2, s2);
    const int i0 = 0, i1 = 1; f1.up(&i0, 1); f2.up(&i1, 1);
    AFV_OK_OR_DIE(afv_search_by_sim3((int)dt, q1d.p, x1.p, r1.p, a1.p, b1.p, s1.p, n1, q2d.p, x2.p, r2.p, a2.p, b2.p, s2.p, n2, 1, dk.p, dd.p, ds.p, dn.p, 2, cap,
                                     f1.p, f2.p, K2.mnMinX, K2.mnMinY, K2.mnMaxX, K2.mnMaxY, TH_HIGH, dm.p, nf.p, nullptr));
    dm.down(m12.data(), n1);
    int n = 0; nf.down(&n, 1);
    return n;
}

// mode 0 SearchByBoW(KF, F) (match indexed by F's keypoints), 1 SearchByBoW(KF, KF), 2 SearchForTriangulation (match indexed by A's keypoints)
int FeatureMatcher::bowCore(int mode, FrameView& A, FrameView& B, const float* F12, const float* epipole, std::vector<int>& match, int descType) {
    const int n1 = (int)A.mvKeysUn.size(), n2 = (int)B.mvKeysUn.size(), cap = std::max(std::max(n1, n2), 1);
    match.assign(mode == 0 ? n2 : n1, -1);
    if (!n1 || !n2) return 0;
    const size_t D = (size_t)A.mDescriptors.cols * A.mDescriptors.elemSize();
    DevBuf<afv_keypoint> dk((size_t)2 * cap); DevBuf<uint8_t> dd((size_t)2 * cap * D), dv((size_t)2 * cap); DevBuf<int> dn(2), node((size_t)2 * cap), pa(1), pb(1), dm(cap), nm(1);
    DevBuf<float> sg((size_t)2 * cap), dF(9), de(2);
    CU_OK_OR_DIE(cudaMemcpy(dk.p, A.mvKeysUn.data(), (size_t)n1 * 28, cudaMemcpyHostToDevice)); CU_OK_OR_DIE(cudaMemcpy(dk.p + cap, B.mvKeysUn.data(), (size_t)n2 * 28, cudaMemcpyHostToDevice));
    CU_OK_OR_DIE(cudaMemcpy(dd.p, A.mDescriptors.data(), (size_t)n1 * D, cudaMemcpyHostToDevice)); CU_OK_OR_DIE(cudaMemcpy(dd.p + (size_t)cap * D, B.mDescriptors.data(), (size_t)n2 * D, cudaMemcpyHostToDevice));
    std::vector<int> nd((size_t)2 * cap, -1); std::vector<uint8_t> v((size_t)2 * cap, mode == 2 ? 0 : 1); std::vector<float> s2((size_t)2 * cap, 1.0f);
    for (int i = 0; i < n1 && i < (int)A.featNode.size(); ++i) nd[i] = A.featNode[i];
    for (int i = 0; i < n2 && i < (int)B.featNode.size(); ++i) nd[cap + i] = B.featNode[i];
    if (A.hasMapPoint.size() == (size_t)n1) std::copy(A.hasMapPoint.begin(), A.hasMapPoint.end(), v.begin());
    if (mode != 0 && B.hasMapPoint.size() == (size_t)n2) std::copy(B.hasMapPoint.begin(), B.hasMapPoint.end(), v.begin() + cap);   // (KF, F): every frame keypoint takes part
    if (B.sigma2_1d.size() == (size_t)n2) std::copy(B.sigma2_1d.begin(), B.sigma2_1d.end(), s2.begin() + cap);
    node.up(nd.data(), nd.size()); dv.up(v.data(), v.size()); sg.up(s2.data(), s2.size());
    const int nn[2] = {n1, n2}, i0 = 0, i1 = 1; dn.up(nn, 2); pa.up(&i0, 1); pb.up(&i1, 1);
    if (F12) { dF.up(F12, 9); de.up(epipole, 2); }
    AFV_OK_OR_DIE(afv_bow_match(mode, descType, dk.p, dd.p, dn.p, 2, cap, node.p, dv.p, pa.p, pb.p, 1, TH_LOW, mfNNratio, mbCheckOrientation ? 1 : 0,
                                F12 ? dF.p : nullptr, F12 ? de.p : nullptr, mode == 2 ? sg.p : nullptr, dm.p, nm.p, nullptr));
    dm.down(match.data(), match.size());
    int n = 0; nm.down(&n, 1);
    return n;
}
int FeatureMatcher::SearchByBoW(KeyFrameView& pKF, FrameView& F, std::vector<int>& vnMatchF, const DescriptorType& dt) { return bowCore(0, pKF, F, nullptr, nullptr, vnMatchF, (int)dt); }
int FeatureMatcher::SearchByBoW(KeyFrameView& pKF1, KeyFrameView& pKF2, std::vector<int>& vnMatches12, const DescriptorType& dt) { return bowCore(1, pKF1, pKF2, nullptr, nullptr, vnMatches12, (int)dt); }
int FeatureMatcher::SearchForTriangulation(KeyFrameView& pKF1, KeyFrameView& pKF2, const float F12[9], const float epipole2[2],
                                           std::vector<std::pair<size_t, size_t>>& vMatchedPairs, const DescriptorType& dt) {
    std::vector<int> m;
    const int n = bowCore(2, pKF1, pKF2, F12, epipole2, m, (int)dt);
    vMatchedPairs.clear();
    for (size_t i = 0; i < m.size(); ++i) if (m[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m[i]));       // :781-787
    return n;
}
}  // namespace ANYFEATURE_VSLAM_B200

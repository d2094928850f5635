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
}
FeatureExtractor::~FeatureExtractor() { if (handle_) afv_extractor_destroy(handle_); }

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
}  // namespace ANYFEATURE_VSLAM_B200

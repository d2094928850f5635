// Drives the C++ host mirror like the reference's Frame / Tracking would: factory -> operator() on two PGM frames ->
// FeatureMatcher::SearchForInitialization; prints results for the pytest harness to compare with the Python path.
#include "../../anyfeature-vslam_b200/host/afv_host.hpp"
#include <cstdio>
#include <fstream>
using namespace ANYFEATURE_VSLAM_B200;
static bool read_pgm(const char* path, Image& im) {
    std::ifstream f(path, std::ios::binary);
    std::string magic; int w, h, mx;
    f >> magic >> w >> h >> mx; f.get();
    if (magic != "P5" || mx != 255) return false;
    im.grayImg.create(h, w, afvcv::CV_8U);
    f.read(reinterpret_cast<char*>(im.grayImg.data()), (size_t)w * h);
    return (bool)f;
}
int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: %s settings.yaml a.pgm b.pgm [orb32|sift128|akaze61|brisk48]\n", argv[0]); return 2; }
    const std::string feature = argc > 4 ? argv[4] : "orb32";
    Image A, B;
    if (!read_pgm(argv[2], A) || !read_pgm(argv[3], B)) return 3;
    auto ext = getFeatureExtractor(1, argv[1], feature, A.grayImg.cols, A.grayImg.rows);
    FeatureMatcher::setDescriptorDistanceThresholds(argv[1]);
    FrameView F[2];
    const Image* ims[2] = {&A, &B};
    for (int i = 0; i < 2; ++i) {
        std::vector<mat2f> s2, inf;
        (*ext)(*ims[i], F[i].mvKeysUn, F[i].mDescriptors, s2, inf, F[i].keyPtsSize);
        F[i].mnMaxX = (float)A.grayImg.cols; F[i].mnMaxY = (float)A.grayImg.rows; F[i].maxKeyPtSize = ext->GetMaxKeyPtSize();
    }
    std::vector<afvcv::Point2f> prev(F[0].mvKeysUn.size());
    for (size_t i = 0; i < prev.size(); ++i) prev[i] = F[0].mvKeysUn[i].pt;
    std::vector<int> m12;
    FeatureMatcher matcher(0.9f, true);
    const int nm = matcher.SearchForInitialization(F[0], F[1], prev, m12, 100, (DescriptorType)get_feature_id(feature));
    unsigned long long h = 1469598103934665603ull;
    for (int i = 0; i < 2; ++i) for (int r = 0; r < F[i].mDescriptors.rows; ++r) for (size_t c = 0; c < F[i].mDescriptors.step(); ++c) { h ^= F[i].mDescriptors.ptr<uint8_t>(r)[c]; h *= 1099511628211ull; }
    for (int v : m12) { h ^= (unsigned)(v + 1); h *= 1099511628211ull; }
    std::printf("n0=%zu n1=%zu matches=%d levels=%d q0=%d hash=%llu\n", F[0].mvKeysUn.size(), F[1].mvKeysUn.size(), nm, ext->GetLevels(),
                ext->GetFeaturesPerLevel()[0], h);
    return 0;
}

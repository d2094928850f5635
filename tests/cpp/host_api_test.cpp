// Drives the C++ host mirror like the reference's Frame / Tracking / LocalMapping / LoopClosing would: factory -> operator() on two
// PGM frames -> every FeatureMatcher search, the Frame helpers, the vanilla ORB-SLAM2 operator() and the protected stage hooks; prints
// one "key=value" token per result for the pytest harness (tests/test_matcher_gpu.py), which repeats the sequence through the Python path.
#include "../../anyfeature-vslam_b200/host/afv_host.hpp"
#include <cstdio>
#include <cstring>
#include <fstream>
using namespace ANYFEATURE_VSLAM_B200;
typedef unsigned long long u64;
static const u64 FNV0 = 1469598103934665603ull, FNVP = 1099511628211ull;
static bool read_pgm(const char* path, Image& im) {
    std::ifstream f(path, std::ios::binary);
    std::string magic; int w, h, mx;
    f >> magic >> w >> h >> mx; f.get();
    if (magic != "P5" || mx != 255) return false;
    im.grayImg.create(h, w, afvcv::CV_8U);
    f.read(reinterpret_cast<char*>(im.grayImg.data()), (size_t)w * h);
    return (bool)f;
}
static u64 hash_ints(const std::vector<int>& v) { u64 h = FNV0; for (int x : v) { h ^= (unsigned)(x + 1); h *= FNVP; } return h; }
static u64 hash_bytes(const void* p, size_t n, u64 h = FNV0) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= FNVP; } return h; }

struct Probe : FeatureExtractor_orb32 {          // a subclass written against the reference's protected hooks
    using FeatureExtractor_orb32::FeatureExtractor_orb32;
    using FeatureExtractor::detectKeypoints; using FeatureExtractor::filterKeypoints;
};

// projection prologue stand-in: query i = keypoint i of `Q` shifted by (dx, dy), radius 6 x size (every 10th skipped), size range size / 1.2 .. size * 1.2
static ProjectedPoints queries(const FrameView& Q, float dx, float dy) {
    ProjectedPoints P;
    const size_t n = Q.mvKeysUn.size();
    P.descriptors = Q.mDescriptors; P.uv.resize(n); P.radius.resize(n); P.minSize.resize(n); P.maxSize.resize(n); P.angle.resize(n);
    for (size_t i = 0; i < n; ++i) {
        P.uv[i] = afvcv::Point2f(Q.mvKeysUn[i].pt.x + dx, Q.mvKeysUn[i].pt.y + dy);
        P.radius[i] = i % 10 == 9 ? -1.0f : 6.0f * Q.keyPtsSize[i];
        P.minSize[i] = Q.keyPtsSize[i] / 1.2f; P.maxSize[i] = Q.keyPtsSize[i] * 1.2f; P.angle[i] = Q.mvKeysUn[i].angle;
    }
    return P;
}

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: %s settings.yaml a.pgm b.pgm [orb32|sift128|akaze61|brisk48]\n", argv[0]); return 2; }
    const std::string feature = argc > 4 ? argv[4] : "orb32";
    Image A, B;
    if (!read_pgm(argv[2], A) || !read_pgm(argv[3], B)) return 3;
    auto ext = getFeatureExtractor(1, argv[1], feature, A.grayImg.cols, A.grayImg.rows);
    FeatureMatcher::setDescriptorDistanceThresholds(argv[1]);
    const DescriptorType dt = (DescriptorType)get_feature_id(feature);
    KeyFrameView F[2];
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
    const int nm = matcher.SearchForInitialization(F[0], F[1], prev, m12, 100, dt);
    u64 h = FNV0;
    for (int i = 0; i < 2; ++i) for (int r = 0; r < F[i].mDescriptors.rows; ++r) for (size_t c = 0; c < F[i].mDescriptors.step(); ++c) { h ^= F[i].mDescriptors.ptr<uint8_t>(r)[c]; h *= FNVP; }
    for (int v : m12) { h ^= (unsigned)(v + 1); h *= FNVP; }
    std::printf("n0=%zu n1=%zu matches=%d levels=%d q0=%d hash=%llu\n", F[0].mvKeysUn.size(), F[1].mvKeysUn.size(), nm, ext->GetLevels(),
                ext->GetFeaturesPerLevel()[0], h);
    std::printf("pyr_default=%zu\n", ext->mvImagePyramid.size() * 100 + (ext->mvImagePyramid[0].empty() ? 0 : 1));

    // ---- the ten other searches; F[1] is the train frame, queries come from F[0]
    const size_t n0 = F[0].mvKeysUn.size(), n1 = F[1].mvKeysUn.size();
    const ProjectedPoints P = queries(F[0], 1.5f, -2.0f);
    std::vector<uint8_t> occ(n1); for (size_t i = 0; i < n1; ++i) occ[i] = i % 7 == 0;
    const char* pv[4] = {"proj_local", "proj_sim3", "proj_motion", "proj_reloc"};
    FeatureMatcher fm(0.8f, true);
    for (int v = 0; v < 4; ++v) {
        KeyFrameView T = F[1]; T.hasMapPoint = occ;
        std::vector<int> m;
        const int n = fm.SearchByProjection((FeatureMatcher::ProjectionVariant)v, T, P, m, dt);
        size_t claimed = 0; for (uint8_t b : T.hasMapPoint) claimed += b;
        std::printf("%s=%d %s_h=%llu %s_occ=%zu\n", pv[v], n, pv[v], hash_ints(m), pv[v], claimed);
    }
    for (int gate = 1; gate >= 0; --gate) {
        KeyFrameView T = F[1];
        T.inf_1d.resize(n1); for (size_t i = 0; i < n1; ++i) T.inf_1d[i] = 1.0f / (T.keyPtsSize[i] * T.keyPtsSize[i]);
        std::vector<int> m;
        const int n = fm.Fuse(T, P, gate != 0, m, dt);
        std::printf("fuse%d=%d fuse%d_h=%llu\n", gate, n, gate, hash_ints(m));
    }
    {
        const ProjectedPoints P2 = queries(F[1], -1.5f, 2.0f);
        std::vector<int> m;
        const int n = fm.SearchBySim3(F[0], F[1], P, P2, m, dt);
        std::printf("sim3=%d sim3_h=%llu\n", n, hash_ints(m));
    }
    if (dt != DESC_SIFT128) {                                  // FeatureVector stand-in: node id = first descriptor byte mod 24
        KeyFrameView K1 = F[0], K2 = F[1];
        K1.featNode.resize(n0); K2.featNode.resize(n1); K1.hasMapPoint.resize(n0); K2.hasMapPoint.resize(n1); K2.sigma2_1d.resize(n1);
        for (size_t i = 0; i < n0; ++i) { K1.featNode[i] = K1.mDescriptors.ptr<uint8_t>((int)i)[0] % 24; K1.hasMapPoint[i] = i % 5 != 0; }
        for (size_t i = 0; i < n1; ++i) { K2.featNode[i] = K2.mDescriptors.ptr<uint8_t>((int)i)[0] % 24; K2.hasMapPoint[i] = i % 4 != 0; K2.sigma2_1d[i] = K2.keyPtsSize[i] * K2.keyPtsSize[i]; }
        FeatureMatcher fb(0.7f, true);
        std::vector<int> m;
        FrameView Fr = F[1]; Fr.featNode = K2.featNode;
        int n = fb.SearchByBoW(K1, Fr, m, dt);
        std::printf("bow_kf_f=%d bow_kf_f_h=%llu\n", n, hash_ints(m));
        n = fb.SearchByBoW(K1, K2, m, dt);
        std::printf("bow_kf_kf=%d bow_kf_kf_h=%llu\n", n, hash_ints(m));
        for (size_t i = 0; i < n0; ++i) K1.hasMapPoint[i] = i % 3 == 0;          // triangulation: keypoints that already HAVE a point are skipped
        for (size_t i = 0; i < n1; ++i) K2.hasMapPoint[i] = i % 3 == 1;
        const float F12[9] = {0, 0, 0, 0, 0, -1, 0, 1, 0}, epi[2] = {1.0e6f, 240.0f};        // pure x translation: horizontal epipolar lines
        std::vector<std::pair<size_t, size_t>> pairs;
        FeatureMatcher ft(0.6f, false);
        n = ft.SearchForTriangulation(K1, K2, F12, epi, pairs, dt);
        u64 hp = FNV0; for (auto& pr : pairs) { hp ^= (unsigned)pr.first; hp *= FNVP; hp ^= (unsigned)pr.second; hp *= FNVP; }
        std::printf("triang=%d triang_h=%llu\n", n, hp);
    }
    // ---- Frame helpers
    {
        const float K[4] = {520.f, 520.f, 320.f, 240.f}, dist[5] = {-0.28f, 0.07f, 0.0002f, 0.00002f, 0.f};
        std::vector<KeyPoint> un;
        UndistortKeyPoints(F[0].mvKeysUn, K, dist, un);
        std::printf("undist_h=%llu\n", hash_bytes(un.data(), un.size() * sizeof(KeyPoint)));
        FrameView G = F[1];
        AssignFeaturesToGrid(G);
        std::printf("grid_items=%zu grid_h=%llu\n", G.gridCellItems.size(), hash_bytes(G.gridCellItems.data(), G.gridCellItems.size() * 4, hash_bytes(G.gridCellStart.data(), G.gridCellStart.size() * 4)));
        MapPointsView mp; PoseView pose; std::memset(&pose, 0, sizeof(pose));
        pose.Rcw[0] = pose.Rcw[4] = pose.Rcw[8] = 1.f; pose.fx = pose.fy = 520.f; pose.cx = 320.f; pose.cy = 240.f;
        for (size_t i = 0; i < n0; ++i) {
            float depth = 2.0f + (float)(i % 5);
            if (i % 9 == 0) depth = -depth;
            const KeyPoint& k = F[0].mvKeysUn[i];
            mp.worldPos.push_back((k.pt.x - 320.f) / 520.f * depth); mp.worldPos.push_back((k.pt.y - 240.f) / 520.f * depth); mp.worldPos.push_back(depth);
            mp.normal.push_back(0.f); mp.normal.push_back(0.f); mp.normal.push_back(1.f);
            mp.minDistance.push_back(0.1f); mp.maxDistance.push_back(i % 11 == 0 ? 1.0f : 100.f);
            mp.refSize.push_back(F[0].keyPtsSize[i]); mp.refSigma.push_back(1.f); mp.refDistance.push_back(3.f);
        }
        std::vector<uint8_t> inview; std::vector<float> vcos; ProjectedPoints Q;
        isInFrustum(F[1], pose, mp, F[0].mDescriptors, 0.5f, 3.0f, inview, Q, vcos);
        size_t nin = 0; for (uint8_t b : inview) nin += b;
        std::printf("frustum_in=%zu frustum_h=%llu\n", nin, hash_bytes(Q.radius.data(), Q.radius.size() * 4, hash_bytes(Q.uv.data(), Q.uv.size() * 8, hash_bytes(inview.data(), inview.size()))));
        KeyFrameView T = F[1];
        std::vector<int> m;
        FeatureMatcher fl(0.8f, true);
        const int n = fl.SearchByProjection(FeatureMatcher::TRACK_LOCAL_MAP, T, Q, m, dt);      // Tracking::SearchLocalPoints
        std::printf("local_points=%d local_points_h=%llu\n", n, hash_ints(m));
    }
    // ---- vanilla ORB-SLAM2 operator() + mvImagePyramid, and the protected per-stage hooks (orb32)
    if (dt == DESC_ORB) {
        std::vector<KeyPoint> k; Mat d; std::vector<mat2f> s2, inf; std::vector<float> sz;
        (*ext)(A, k, d, s2, inf, sz, true);
        u64 hv = hash_bytes(d.data(), (size_t)d.rows * 32, hash_bytes(k.data(), k.size() * sizeof(KeyPoint)));
        const Mat& l3 = ext->mvImagePyramid[3];
        std::printf("vanilla_n=%zu vanilla_h=%llu pyr3=%dx%d pyr3_h=%llu\n", k.size(), hv, l3.cols, l3.rows, hash_bytes(l3.data(), (size_t)l3.cols * l3.rows));
        auto st = std::make_shared<FeatureExtractorSettings>(KEYP_ORB, DESC_ORB, "none");
        Probe pr(1000, st);
        std::vector<KeyPoint> k2; Mat d2;
        pr(A, k2, d2);
        std::map<int, std::vector<KeyPoint>> det, kept;
        pr.detectKeypoints(det, A, st->detectTh, st->nOctaves);
        pr.filterKeypoints(kept, A.grayImg, A.mask);
        size_t nd = 0, nk = 0; for (auto& kv : det) nd += kv.second.size(); for (auto& kv : kept) nk += kv.second.size();
        std::printf("hook_detect=%zu hook_kept=%zu hook_n=%zu\n", nd, nk, k2.size());
    }
    return 0;
}

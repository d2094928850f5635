"""Builds oracle/_ref/libafv_ref.so: function bodies of the REFERENCE, compiled from the sources where they lie under
/root/reference (never copied into the repo; the generated translation unit lives only in the git-ignored oracle/_ref/).

The reference's hot path as a whole cannot be built here (every TU needs OpenCV C++ headers, Eigen, ...), but the in-repo
parts of the path are plain C++ over cv::KeyPoint / cv::Mat / STL.  This script cuts those function definitions out of the
reference sources by signature, puts them behind oracle/ref_shim.hpp and adds extern "C" drivers:
  src/ORBextractor.cc   ExtractorNode::DivideNode, FeatureExtractor::DistributeOctTree            (:181-458)
  src/Frame.cc          Frame::AssignFeaturesToGrid, GetFeaturesInArea, PosInGrid                 (:225-240, :333-394)
  src/FeatureMatcher.cc SearchForInitialization, DescriptorDistance, rotation-histogram helpers   (:399-557, :1508-1531, :1579-1668)
  src/Feature_{orb32,akaze61,brisk48,sift128}.cpp   DescriptorDistance_<feat>
tests/test_oracle_vs_ref.py checks the oracle restatement against this library.  TEST INFRASTRUCTURE ONLY.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AFV_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "libafv_ref.so")


def cut(path, signature_regex, all_matches=False):
    """Text of the function definition(s) whose first line matches the regex, up to the matching closing brace."""
    src = open(os.path.join(REF, path), errors="ignore").read()
    out = []
    for m in re.finditer(signature_regex, src, flags=re.M):
        i = src.index("{", m.end() - 1) if src[m.end() - 1] != "{" else m.end() - 1
        depth, j = 0, i
        while True:
            c = src[j]
            if c == "{":
                depth += 1
            elif c == "}":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        out.append(src[m.start():j + 1])
        if not all_matches:
            break
    if not out:
        raise SystemExit("build_ref: no match for %r in %s" % (signature_regex, path))
    return "\n\n".join(out)


def cut_between(path, start_marker, end_marker):
    """Text of a reference source between two literal markers (start inclusive, end exclusive)."""
    src = open(os.path.join(REF, path), errors="ignore").read()
    a = src.index(start_marker)
    b = src.index(end_marker, a)
    return src[a:b]


DRIVERS = r'''
// ---- deterministic heap for the reference code -------------------------------------------------------------------------
// DistributeOctTree sorts (nKeys, ExtractorNode*) pairs (src/ORBextractor.cc:381): among nodes with equal key counts the
// division order -- and with it the kept set (by a few keypoints) and the output order -- depends on HEAP ADDRESSES.  With
// the stock allocator the compiled reference does not even agree with itself between two calls on the same input
// (tests/test_oracle_vs_ref.py::test_reference_octree_depends_on_heap_addresses).  For the comparison with the oracle this
// library therefore runs the unmodified reference code on a monotonic bump allocator (-Wl,-Bsymbolic binds the library's
// operator new / delete to these definitions): a later allocation has a larger address, which is the canonical tie order the
// oracle documents.  ref_set_bump(0) restores malloc.
#include <sys/mman.h>
#include <cstdlib>
#include <functional>
#include <new>
static char* g_arena = nullptr; static size_t g_arena_off = 0; static const size_t ARENA = (size_t)4 << 30; static int g_bump = 1;
static void arena_reset() {
    if (!g_arena) g_arena = (char*)mmap(nullptr, ARENA, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    else madvise(g_arena, g_arena_off, MADV_DONTNEED);
    g_arena_off = 0;
}
static inline bool in_arena(void* p) { return g_arena && (char*)p >= g_arena && (char*)p < g_arena + ARENA; }
void* operator new(size_t n) {
    if (g_bump && g_arena) { const size_t a = (g_arena_off + 15) & ~(size_t)15; if (a + n <= ARENA) { g_arena_off = a + n; return g_arena + a; } }
    void* p = malloc(n ? n : 1); if (!p) throw std::bad_alloc(); return p;
}
void operator delete(void* p) noexcept { if (p && !in_arena(p)) free(p); }
void operator delete(void* p, size_t) noexcept { if (p && !in_arena(p)) free(p); }
extern "C" void ref_set_bump(int on) { g_bump = on; }

namespace ANYFEATURE_VSLAM {
float Frame::mfGridElementWidthInv = 0, Frame::mfGridElementHeightInv = 0, Frame::mnMinX = 0, Frame::mnMaxX = 0, Frame::mnMinY = 0, Frame::mnMaxY = 0;
Descriptor_Distance_Type FeatureMatcher::TH_HIGH = 0, FeatureMatcher::TH_LOW = 0, FeatureMatcher::descDistTh_high_reloc = 0, FeatureMatcher::descDistTh_low_reloc = 0;
float FeatureMatcher::radiusScale = 1.0f;
float FeatureExtractorSettings::scaleFactor0 = 1.2f;
}
cv::VanillaCallbacks cv::g_vanilla = {nullptr, nullptr, nullptr, nullptr};
std::vector<cv::KeyPoint>* cv::ORB::g_detect = nullptr;
void (*cv::ORB::g_compute)(const cv::KeyPoint*, int, unsigned char*) = nullptr;
int cv::ORB::g_last[4] = {0, 0, 0, 0};
namespace ANYFEATURE_VSLAM {
const int FeatureMatcher::HISTO_LENGTH = 30;                       // src/FeatureMatcher.cc:64
}
using namespace ANYFEATURE_VSLAM;
struct kp7 { float x, y, size, angle, response; int octave, class_id; };
static void fill_frame(Frame& F, const kp7* k, int n, void* desc, int dcols, int dtype, const float* ksize, float max_kpt_size) {
    F.N = n; F.mvKeysUn.resize(n); F.keyPtsSize.assign(ksize, ksize + n); F.maxKeyPtSize = max_kpt_size;
    for (int i = 0; i < n; ++i) { cv::KeyPoint p; p.pt.x = k[i].x; p.pt.y = k[i].y; p.size = k[i].size; p.angle = k[i].angle; p.response = k[i].response; p.octave = k[i].octave; p.class_id = k[i].class_id; F.mvKeysUn[i] = p; }
    F.mDescriptors = cv::Mat(n, dcols, dtype, desc);
    F.AssignFeaturesToGrid();
}
extern "C" {
// FeatureExtractor::DistributeOctTree on keypoints (x, y, response); returns the kept keypoints' (x, y, response) in list order
int ref_distribute_octree(const float* px, const float* py, const float* resp, int n, int minX, int maxX, int minY, int maxY, int N,
                          float* ox, float* oy, float* oresp, int cap) {
    if (g_bump) arena_reset();
    int m = 0;
    {
        std::vector<cv::KeyPoint> v(n);
        for (int i = 0; i < n; ++i) { v[i].pt.x = px[i]; v[i].pt.y = py[i]; v[i].response = resp[i]; v[i].class_id = i; }
        FeatureExtractor fe;
        std::vector<cv::KeyPoint> r = fe.DistributeOctTree(v, minX, maxX, minY, maxY, N, 0);
        for (size_t i = 0; i < r.size() && (int)i < cap; ++i) { ox[i] = r[i].pt.x; oy[i] = r[i].pt.y; oresp[i] = (float)r[i].class_id; }
        m = (int)r.size();
    }
    return m;
}
// FeatureMatcher::SearchForInitialization between two frames given as arrays; prev (n1 x 2) is updated like vbPrevMatched
int ref_search_for_initialization(int desc_type, int dcols, int dtype, const kp7* k1, void* d1, const float* s1, int n1,
                                  const kp7* k2, void* d2, const float* s2, int n2, float minX, float minY, float maxX, float maxY,
                                  float max_kpt_size, float* prev, int window, float th_low, float nnratio, int check_ori, int* matches12) {
    Frame::mnMinX = minX; Frame::mnMinY = minY; Frame::mnMaxX = maxX; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (maxX - minX);      // src/Frame.cc:198-199
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (maxY - minY);
    FeatureMatcher::TH_LOW = th_low; FeatureMatcher::TH_HIGH = th_low;
    Frame F1, F2;
    fill_frame(F1, k1, n1, d1, dcols, dtype, s1, max_kpt_size);
    fill_frame(F2, k2, n2, d2, dcols, dtype, s2, max_kpt_size);
    std::vector<cv::Point2f> pm(n1);
    for (int i = 0; i < n1; ++i) { pm[i].x = prev[2 * i]; pm[i].y = prev[2 * i + 1]; }
    std::vector<int> m12;
    FeatureMatcher fm(nnratio, check_ori != 0);
    const int nm = fm.SearchForInitialization(F1, F2, pm, m12, window, (DescriptorType)desc_type);
    for (int i = 0; i < n1; ++i) { matches12[i] = m12[i]; prev[2 * i] = pm[i].x; prev[2 * i + 1] = pm[i].y; }
    return nm;
}
// Frame::GetFeaturesInArea on one frame
int ref_features_in_area(const kp7* k, const float* ksize, int n, float minX, float minY, float maxX, float maxY, float x, float y, float r,
                         float min_size, float max_size, int* out, int cap) {
    Frame::mnMinX = minX; Frame::mnMinY = minY; Frame::mnMaxX = maxX; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (maxX - minX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (maxY - minY);
    Frame F;
    unsigned char dummy = 0;
    fill_frame(F, k, n, &dummy, 1, 0, ksize, 0.f);
    std::vector<size_t> v = F.GetFeaturesInArea(x, y, r, min_size, max_size);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int)v[i];
    return (int)v.size();
}
// FeatureMatcher::SearchByProjection(Frame&, vector<Pt>&, radiusTh) (TrackLocalMap): query i = a map point with descriptor,
// projected position, predicted size and viewing cosine; occupied[idx] != 0 = keypoint idx already holds an observed map point.
int ref_search_by_projection(int desc_type, int dcols, int dtype, void* qdesc, const float* qxy, const float* qsize, const float* qcos, int nq,
                             const kp7* k, void* d, const float* ksize, int n, const unsigned char* occupied, float minX, float minY,
                             float maxX, float maxY, float radius_th, float radius_scale, float size_tol, float th_high, float nnratio,
                             int* match_q) {
    Frame::mnMinX = minX; Frame::mnMinY = minY; Frame::mnMaxX = maxX; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (maxX - minX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (maxY - minY);
    FeatureMatcher::TH_HIGH = th_high; FeatureMatcher::TH_LOW = th_high; FeatureMatcher::radiusScale = radius_scale;
    Frame F;
    fill_frame(F, k, n, d, dcols, dtype, ksize, 0.f);
    F.sizeTolerance = size_tol; F.invSizeTolerance = 1.0f / size_tol;       // src/Frame.cc:182-183
    F.pts.assign(n, Pt()); F.mvuRight.assign(n, -1.0f);
    Pt held = std::make_shared<MapPoint>();
    for (int i = 0; i < n; ++i) if (occupied && occupied[i]) F.pts[i] = held;
    std::vector<Pt> q(nq);
    cv::Mat Q(nq, dcols, dtype, qdesc);
    for (int i = 0; i < nq; ++i) {
        q[i] = std::make_shared<MapPoint>();
        q[i]->desc = Q.row(i); q[i]->descriptorType = (DescriptorType)desc_type; q[i]->trackSize = qsize[i]; q[i]->trackViewCos = qcos[i];
        q[i]->mTrackProjX = qxy[2 * i]; q[i]->mTrackProjY = qxy[2 * i + 1];
    }
    FeatureMatcher fm(nnratio, true);
    const int nm = fm.SearchByProjection(F, q, radius_th);
    for (int i = 0; i < nq; ++i) match_q[i] = -1;
    for (int idx = 0; idx < n; ++idx)
        if (F.pts[idx] && F.pts[idx] != held)
            for (int i = 0; i < nq; ++i) if (F.pts[idx] == q[i]) { match_q[i] = idx; break; }
    return nm;
}
// FeatureMatcher::SearchByBoW(KF, F): FeatureVectors as (node ids, CSR starts, feature indices); every keyframe keypoint
// holds a good map point.  match_f[f] = keyframe keypoint index matched to frame keypoint f, or -1.
int ref_search_by_bow(int desc_type, int dcols, int dtype, const kp7* kkf, void* dkf, int nkf, const int* kf_node, const int* kf_start,
                      const int* kf_idx, int kf_nodes, const kp7* kf_, void* df, int nf, const int* f_node, const int* f_start,
                      const int* f_idx, int f_nodes, float th_low, float nnratio, int check_ori, int* match_f) {
    FeatureMatcher::TH_LOW = th_low; FeatureMatcher::TH_HIGH = th_low;
    Keyframe KF = std::make_shared<KeyFrame>();
    KF->mvKeysUn.resize(nkf); KF->mappoints.resize(nkf);
    for (int i = 0; i < nkf; ++i) {
        cv::KeyPoint p; p.pt.x = kkf[i].x; p.pt.y = kkf[i].y; p.angle = kkf[i].angle; p.octave = kkf[i].octave; KF->mvKeysUn[i] = p;
        KF->mappoints[i] = std::make_shared<MapPoint>(); KF->mappoints[i]->descriptorType = (DescriptorType)desc_type;
    }
    KF->mDescriptors = cv::Mat(nkf, dcols, dtype, dkf);
    for (int s = 0; s < kf_nodes; ++s) for (int j = kf_start[s]; j < kf_start[s + 1]; ++j) KF->mFeatVec[(unsigned)kf_node[s]].push_back((unsigned)kf_idx[j]);
    Frame F;
    F.N = nf; F.mvKeys.resize(nf);
    for (int i = 0; i < nf; ++i) { cv::KeyPoint p; p.pt.x = kf_[i].x; p.pt.y = kf_[i].y; p.angle = kf_[i].angle; p.octave = kf_[i].octave; F.mvKeys[i] = p; }
    F.mDescriptors = cv::Mat(nf, dcols, dtype, df);
    for (int s = 0; s < f_nodes; ++s) for (int j = f_start[s]; j < f_start[s + 1]; ++j) F.mFeatVec[(unsigned)f_node[s]].push_back((unsigned)f_idx[j]);
    std::vector<Pt> matches;
    FeatureMatcher fm(nnratio, check_ori != 0);
    const int nm = fm.SearchByBoW(KF, F, matches);
    for (int f = 0; f < nf; ++f) {
        match_f[f] = -1;
        if (matches[f]) for (int i = 0; i < nkf; ++i) if (KF->mappoints[i] == matches[f]) { match_f[f] = i; break; }
    }
    return nm;
}
// Vocabulary::transform (src/Vocabulary.cpp:156-207 -> TemplatedVocabulary::transform per feature, levelsup 4 in the reference):
// the tree comes as CSR children lists + per-node descriptor / word id / weight (node 0 = root)
}   // extern "C"
template <class F, class MakeDesc>
static void bow_run(int n, const int* child_off, const int* child_ids, int nnodes, const int* node_word, const double* node_weight, int L,
                    int levelsup, MakeDesc mk_node, MakeDesc mk_feat, int* word_id, double* weight, int* node_id) {
    DBoW2::TemplatedVocabulary<typename F::TDescriptor, F> V;
    V.m_L = L; V.m_nodes.resize(nnodes);
    for (int i = 0; i < nnodes; ++i) {
        V.m_nodes[i].id = i; V.m_nodes[i].weight = node_weight[i]; V.m_nodes[i].word_id = node_word[i] < 0 ? 0 : node_word[i];
        V.m_nodes[i].descriptor = mk_node(i);
        for (int j = child_off[i]; j < child_off[i + 1]; ++j) V.m_nodes[i].children.push_back((DBoW2::NodeId)child_ids[j]);
    }
    for (int f = 0; f < n; ++f) {
        DBoW2::WordId w = 0; DBoW2::WordValue wt = 0; DBoW2::NodeId nid = 0;
        V.transform(mk_feat(f), w, wt, &nid, levelsup);
        word_id[f] = (int)w; weight[f] = wt; node_id[f] = (int)nid;
    }
}
extern "C" {
int ref_bow_transform(int desc_type, void* desc, int n, const int* child_off, const int* child_ids, int nnodes, void* node_desc,
                      const int* node_word, const double* node_weight, int L, int levelsup, int* word_id, double* weight, int* node_id) {
    if (desc_type == 5) {
        const float* nd = (const float*)node_desc; const float* fd = (const float*)desc;
        auto mkn = [&](int i) { return std::vector<float>(nd + (size_t)i * 128, nd + (size_t)(i + 1) * 128); };
        auto mkf = [&](int i) { return std::vector<float>(fd + (size_t)i * 128, fd + (size_t)(i + 1) * 128); };
        bow_run<DBoW2::FSift128, std::function<std::vector<float>(int)>>(n, child_off, child_ids, nnodes, node_word, node_weight, L, levelsup, mkn, mkf, word_id, weight, node_id);
        return 0;
    }
    const int D = desc_type == 0 ? 32 : desc_type == 1 ? 61 : 48;
    // rows are copied into 8-byte aligned, zero-padded storage: the DBoW2 distances read uint64_t / int32_t words
    const int Dp = (D + 7) & ~7;
    std::vector<uint64_t> nbuf((size_t)nnodes * Dp / 8, 0), fbuf((size_t)(n > 0 ? n : 1) * Dp / 8, 0);
    for (int i = 0; i < nnodes; ++i) memcpy((char*)nbuf.data() + (size_t)i * Dp, (char*)node_desc + (size_t)i * D, D);
    for (int i = 0; i < n; ++i) memcpy((char*)fbuf.data() + (size_t)i * Dp, (char*)desc + (size_t)i * D, D);
    auto mkn = [&](int i) { return cv::Mat(1, D, 0, (char*)nbuf.data() + (size_t)i * Dp); };
    auto mkf = [&](int i) { return cv::Mat(1, D, 0, (char*)fbuf.data() + (size_t)i * Dp); };
    typedef std::function<cv::Mat(int)> MK;
    if (desc_type == 0) bow_run<DBoW2::FOrb, MK>(n, child_off, child_ids, nnodes, node_word, node_weight, L, levelsup, mkn, mkf, word_id, weight, node_id);
    else if (desc_type == 1) bow_run<DBoW2::FAkaze61, MK>(n, child_off, child_ids, nnodes, node_word, node_weight, L, levelsup, mkn, mkf, word_id, weight, node_id);
    else bow_run<DBoW2::FBrisk, MK>(n, child_off, child_ids, nnodes, node_word, node_weight, L, levelsup, mkn, mkf, word_id, weight, node_id);
    return 0;
}
// FeatureExtractor constructor (src/FeatureExtractor.cpp:74-109: mvScaleFactor, mnFeaturesPerLevel) + computeSize (:132-142)
int ref_extractor_tables(int nfeatures, int nlevels, float scale_factor, float* scale_factors, int* quota, float* size_norm) {
    std::shared_ptr<FeatureExtractorSettings> st = std::make_shared<FeatureExtractorSettings>();
    st->scaleFactor = scale_factor; st->nOctaves = nlevels; FeatureExtractorSettings::scaleFactor0 = scale_factor;
    st->maxKeyPtSize0 = pow(1.2f, float(8 - 1.0)); st->maxKeyPtSize = st->maxKeyPtSize0; st->minKeyPtSize = 1.0f;   // :52-55
    FeatureExtractor fe(nfeatures, st);
    std::vector<cv::KeyPoint> k(nlevels);
    for (int l = 0; l < nlevels; ++l) k[l].octave = l;
    std::vector<float> sz;
    fe.computeSize(sz, k);
    for (int l = 0; l < nlevels; ++l) { scale_factors[l] = fe.mvScaleFactor[l]; quota[l] = fe.mnFeaturesPerLevel[l]; size_norm[l] = sz[l]; }
    return nlevels;
}
// MapPoint::ComputeDistinctiveDescriptors on rows obs[0..n) of a descriptor matrix: returns the position (into obs) of the pick
int ref_distinctive_descriptor(int desc_type, int dcols, int dtype, void* desc, const int* obs, int n) {
    cv::Mat M(1 << 30, dcols, dtype, desc);
    std::vector<cv::Mat> d;
    for (int i = 0; i < n; ++i) d.push_back(M.row(obs[i]));
    return ref_distinctive_core(d, (DescriptorType)desc_type);
}
// computeOrbDescriptor of the reference header on a (blurred) level image
void ref_orb_descriptor(unsigned char* img, int w, int h, int stride, float x, float y, float angle_deg, unsigned char* desc32) {
    cv::Mat M(h, w, 0, img); M.step = (size_t)stride;
    cv::KeyPoint kp; kp.pt.x = x; kp.pt.y = y; kp.angle = angle_deg;
    computeOrbDescriptor(kp, M, (const cv::Point*)bit_pattern_31_, desc32);
}
const int* ref_orb_pattern() { return bit_pattern_31_; }
// FeatureMatcher::SearchByProjection(CurrentFrame, LastFrame, radiusTh, bMono) (TrackWithMotionModel, :1291-1402): identity
// poses, fx = fy = 1, cx = cy = 0 and world points (u, v, 1) make the projection of last-frame map point i land exactly on
// qxy[i]; LastFrame.GetKeyPtSize(i) = qsize[i].  match_q[i] = current-frame keypoint matched to query i, or -1.
int ref_search_by_projection_frames(int desc_type, int dcols, int dtype, void* qdesc, const float* qxy, const float* qsize, const float* qangle,
                                    int nq, const kp7* k, void* d, const float* ksize, int n, const unsigned char* occupied, float minX,
                                    float minY, float maxX, float maxY, float radius_th, float radius_scale, float size_tol, float th_high,
                                    int check_ori, int* match_q) {
    Frame::mnMinX = minX; Frame::mnMinY = minY; Frame::mnMaxX = maxX; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (maxX - minX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (maxY - minY);
    FeatureMatcher::TH_HIGH = th_high; FeatureMatcher::TH_LOW = th_high; FeatureMatcher::radiusScale = radius_scale;
    Frame Cur;
    fill_frame(Cur, k, n, d, dcols, dtype, ksize, 0.f);
    Cur.sizeTolerance = size_tol; Cur.invSizeTolerance = 1.0f / size_tol;
    Cur.pts.assign(n, Pt()); Cur.mvuRight.assign(n, -1.0f);
    Pt held = std::make_shared<MapPoint>();
    for (int i = 0; i < n; ++i) if (occupied && occupied[i]) Cur.pts[i] = held;
    Frame Last;
    Last.N = nq; Last.mvKeysUn.resize(nq); Last.keyPtsSize.assign(qsize, qsize + nq); Last.pts.resize(nq); Last.mvbOutlier.assign(nq, false);
    cv::Mat Q(nq, dcols, dtype, qdesc);
    for (int i = 0; i < nq; ++i) {
        Last.mvKeysUn[i].angle = qangle[i];
        Last.pts[i] = std::make_shared<MapPoint>();
        Last.pts[i]->desc = Q.row(i); Last.pts[i]->descriptorType = (DescriptorType)desc_type;
        Last.pts[i]->worldPos.v[0] = qxy[2 * i]; Last.pts[i]->worldPos.v[1] = qxy[2 * i + 1]; Last.pts[i]->worldPos.v[2] = 1.0f;
    }
    FeatureMatcher fm(0.9f, check_ori != 0);
    const int nm = fm.SearchByProjection(Cur, Last, radius_th, true);
    for (int i = 0; i < nq; ++i) match_q[i] = -1;
    for (int idx = 0; idx < n; ++idx)
        if (Cur.pts[idx] && Cur.pts[idx] != held)
            for (int i = 0; i < nq; ++i) if (Cur.pts[idx] == Last.pts[i]) { match_q[i] = idx; break; }
    return nm;
}
// FeatureExtractor_sift128::detectAndCompute + computeSize of the reference on a prepared SiftGPU feature list (x, y, s, o, desc)
int ref_sift128_glue(const float* xyso, const float* desc, int n, int w, int h, int nfeatures, int nlevels, float scale_factor,
                     kp7* okps, float* odesc, float* osize, int cap) {
    if (g_bump) arena_reset();                               // DistributeOctTree runs inside: monotonic heap (see above)
    std::shared_ptr<FeatureExtractorSettings> st = std::make_shared<FeatureExtractorSettings>();
    st->scaleFactor = scale_factor; st->nOctaves = nlevels; FeatureExtractorSettings::scaleFactor0 = scale_factor; st->detectTh = 10.0f;
    st->maxKeyPtSize0 = pow(1.2f, float(8 - 1.0)); st->maxKeyPtSize = st->maxKeyPtSize0; st->minKeyPtSize = 1.0f;
    FeatureExtractor_sift128 fe(nfeatures, st);
    fe.sift->keys.resize(n); fe.sift->desc.assign(desc, desc + (size_t)n * 128);
    for (int i = 0; i < n; ++i) { fe.sift->keys[i].x = xyso[4 * i]; fe.sift->keys[i].y = xyso[4 * i + 1]; fe.sift->keys[i].s = xyso[4 * i + 2]; fe.sift->keys[i].o = xyso[4 * i + 3]; }
    Image img; unsigned char dummy = 0; img.grayImg = cv::Mat(h, w, CV_8U, &dummy);
    std::vector<cv::KeyPoint> k; cv::Mat d; std::vector<float> sz;
    fe.detectAndCompute(img, k, d);
    fe.computeSize(sz, k);
    const int m = (int)k.size();
    for (int i = 0; i < m && i < cap; ++i) {
        okps[i].x = k[i].pt.x; okps[i].y = k[i].pt.y; okps[i].size = k[i].size; okps[i].angle = k[i].angle; okps[i].response = k[i].response;
        okps[i].octave = k[i].octave; okps[i].class_id = k[i].class_id; osize[i] = sz[i];
        memcpy(odesc + (size_t)i * 128, d.ptr<float>(i), 512);
    }
    return m;
}
// FeatureExtractor_akaze61::detectAndCompute + computeSize on a prepared Feature_Detection list; descriptors / angles of every
// detected keypoint come as a table (what libAKAZE::Compute_Descriptors would produce for it)
int ref_akaze61_glue(const kp7* det, const float* det_angle, const unsigned char* det_desc, int n, int w, int h, int nfeatures, int nlevels,
                     float scale_factor, float detect_th, kp7* okps, unsigned char* odesc, float* osize, int cap) {
    if (g_bump) arena_reset();
    std::shared_ptr<FeatureExtractorSettings> st = std::make_shared<FeatureExtractorSettings>();
    st->scaleFactor = scale_factor; st->nOctaves = nlevels; FeatureExtractorSettings::scaleFactor0 = scale_factor; st->detectTh = detect_th;
    st->maxKeyPtSize0 = pow(1.2f, float(8 - 1.0)); st->maxKeyPtSize = st->maxKeyPtSize0; st->minKeyPtSize = 1.0f;
    FeatureExtractor_akaze61 fe(nfeatures, st);
    fe.evolution->detected.resize(n); fe.evolution->table_angle.assign(det_angle, det_angle + n);
    fe.evolution->table_desc.assign(det_desc, det_desc + (size_t)n * 61);
    for (int i = 0; i < n; ++i) {
        cv::KeyPoint p; p.pt.x = det[i].x; p.pt.y = det[i].y; p.size = det[i].size; p.angle = 0; p.response = det[i].response; p.octave = det[i].octave; p.class_id = det[i].class_id;
        fe.evolution->detected[i] = p;
    }
    fe.evolution->table_kp = fe.evolution->detected;
    Image img; unsigned char dummy = 0; img.grayImg = cv::Mat(h, w, CV_8U, &dummy);
    std::vector<cv::KeyPoint> k; cv::Mat d; std::vector<float> sz;
    fe.detectAndCompute(img, k, d);
    fe.computeSize(sz, k);
    const int m = (int)k.size();
    for (int i = 0; i < m && i < cap; ++i) {
        okps[i].x = k[i].pt.x; okps[i].y = k[i].pt.y; okps[i].size = k[i].size; okps[i].angle = k[i].angle; okps[i].response = k[i].response;
        okps[i].octave = k[i].octave; okps[i].class_id = k[i].class_id; osize[i] = sz[i];
        memcpy(odesc + (size_t)i * 61, d.data + (size_t)i * 61, 61);
    }
    return m;
}
// FeatureExtractor_orb32: initializeExtractor + detectAndCompute + computeSize of the reference.  det = cv::ORB::detect output
// (the tests pass cv2's real one), compute_cb = ORB.compute on the keypoints the glue selects (the tests call cv2's real one).
// orb_params receives the (maxFeatures, edgeThreshold, fastThreshold, nLevels) the reference configured cv::ORB with.
int ref_orb32_glue(const kp7* det, int n, int w, int h, int nfeatures, int nlevels, float scale_factor, float detect_th,
                   void (*compute_cb)(const cv::KeyPoint*, int, unsigned char*), kp7* okps, unsigned char* odesc, float* osize, int cap,
                   int* orb_params) {
    if (g_bump) arena_reset();
    std::shared_ptr<FeatureExtractorSettings> st = std::make_shared<FeatureExtractorSettings>();
    st->scaleFactor = scale_factor; st->nOctaves = nlevels; FeatureExtractorSettings::scaleFactor0 = scale_factor; st->detectTh = detect_th;
    st->maxKeyPtSize0 = pow(1.2f, float(8 - 1.0)); st->maxKeyPtSize = st->maxKeyPtSize0; st->minKeyPtSize = 1.0f;
    std::vector<cv::KeyPoint> dv(n);
    for (int i = 0; i < n; ++i) { dv[i].pt.x = det[i].x; dv[i].pt.y = det[i].y; dv[i].size = det[i].size; dv[i].angle = det[i].angle; dv[i].response = det[i].response; dv[i].octave = det[i].octave; dv[i].class_id = det[i].class_id; }
    cv::ORB::g_detect = &dv; cv::ORB::g_compute = compute_cb;
    FeatureExtractor_orb32 fe(nfeatures, st);
    Image img; unsigned char dummy = 0; img.grayImg = cv::Mat(h, w, CV_8U, &dummy);
    fe.initializeExtractor(img);
    std::vector<cv::KeyPoint> k; cv::Mat d; std::vector<float> sz;
    fe.detectAndCompute(img, k, d);
    fe.computeSize(sz, k);
    for (int i = 0; i < 4; ++i) orb_params[i] = cv::ORB::g_last[i];
    const int m = (int)k.size();
    for (int i = 0; i < m && i < cap; ++i) {
        okps[i].x = k[i].pt.x; okps[i].y = k[i].pt.y; okps[i].size = k[i].size; okps[i].angle = k[i].angle; okps[i].response = k[i].response;
        okps[i].octave = k[i].octave; okps[i].class_id = k[i].class_id; osize[i] = sz[i];
        memcpy(odesc + (size_t)i * 32, d.data + (size_t)i * 32, 32);
    }
    return m;
}

// ---- the remaining FeatureMatcher searches (rows a19 / a20): look-alike KeyFrame / MapPoint objects are set up so that the
// reference's own projection prologue lands exactly on the given (u, v): identity pose, fx = fy = 1, cx = cy = 0 and world
// point (u, v, 1); PredictSize returns the per-point predicted size; distance-invariance range wide open; normal = viewing ray.
static Keyframe make_kf(const kp7* k, int n, void* desc, int dcols, int dtype, const float* ksize, float minX, float minY, float maxX, float maxY,
                        float size_tol) {
    Frame::mnMinX = minX; Frame::mnMinY = minY; Frame::mnMaxX = maxX; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (maxX - minX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (maxY - minY);
    Frame F;
    fill_frame(F, k, n, desc, dcols, dtype, ksize, 0.f);
    Keyframe KF = std::make_shared<KeyFrame>();
    KF->N = n; KF->mvKeysUn = F.mvKeysUn; KF->mDescriptors = F.mDescriptors; KF->keyPtsSize = F.keyPtsSize;
    KF->mappoints.assign(n, Pt()); KF->mvuRight.assign(n, -1.0f); KF->sizeTolerance = size_tol;
    KF->sigma2_1d.assign(n, 1.0f); KF->inf_1d.assign(n, 1.0f);
    KF->mnMinX = minX; KF->mnMinY = minY; KF->mnMaxX = maxX; KF->mnMaxY = maxY;
    KF->mfGridElementWidthInv = Frame::mfGridElementWidthInv; KF->mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    KF->mGrid.resize(FRAME_GRID_COLS);                                   // KeyFrame copies the Frame's grid (src/KeyFrame.cc:57-63)
    for (int i = 0; i < FRAME_GRID_COLS; ++i) { KF->mGrid[i].resize(FRAME_GRID_ROWS); for (int j = 0; j < FRAME_GRID_ROWS; ++j) KF->mGrid[i][j] = F.mGrid[i][j]; }
    return KF;
}
static Pt make_point(int desc_type, const cv::Mat& Q, int i, float u, float v, float pred_size, bool skip) {
    Pt p = std::make_shared<MapPoint>();
    p->desc = Q.row(i); p->descriptorType = (DescriptorType)desc_type; p->trackSize = pred_size; p->bad = skip;
    p->worldPos(0) = u; p->worldPos(1) = v; p->worldPos(2) = 1.0f;
    const float nn = std::sqrt(u * u + v * v + 1.0f);
    p->normal(0) = u / nn; p->normal(1) = v / nn; p->normal(2) = 1.0f / nn;
    return p;
}
static void feat_vec(DBoW2::FeatureVector& fv, const int* node, int n) {
    for (int i = 0; i < n; ++i) if (node[i] >= 0) fv[(unsigned)node[i]].push_back((unsigned)i);
}
// SearchByProjection(pKF, Scw, vpPoints, vpMatched, radiusTh) (:287-397, loop closing): occupied = vpMatched already set
int ref_search_by_projection_sim3(int desc_type, int dcols, int dtype, void* qdesc, const float* qxy, const float* qsize, const unsigned char* qskip,
                                  int nq, const kp7* k, void* d, const float* ksize, int n, const unsigned char* occupied, float minX, float minY,
                                  float maxX, float maxY, float radius_th, float size_tol, float th_low, int* match_q) {
    FeatureMatcher::TH_LOW = th_low; FeatureMatcher::TH_HIGH = th_low; FeatureMatcher::radiusScale = 1.0f;
    Keyframe KF = make_kf(k, n, d, dcols, dtype, ksize, minX, minY, maxX, maxY, size_tol);
    cv::Mat Q(nq, dcols, dtype, qdesc);
    std::vector<Pt> q(nq), matched(n);
    Pt held = std::make_shared<MapPoint>();
    for (int i = 0; i < n; ++i) if (occupied && occupied[i]) matched[i] = held;
    for (int i = 0; i < nq; ++i) q[i] = make_point(desc_type, Q, i, qxy[2 * i], qxy[2 * i + 1], qsize[i], qskip && qskip[i]);
    FeatureMatcher fm(0.6f, true);
    mat4f Scw;
    const int nm = fm.SearchByProjection(KF, Scw, q, matched, radius_th);
    for (int i = 0; i < nq; ++i) match_q[i] = -1;
    for (int idx = 0; idx < n; ++idx) if (matched[idx] && matched[idx] != held) for (int i = 0; i < nq; ++i) if (matched[idx] == q[i]) { match_q[i] = idx; break; }
    return nm;
}
// SearchByProjection(CurrentFrame, pKF, sAlreadyFound, radiusTh, useHigh) (:1406-1506, relocalisation): query i = map point of
// keyframe keypoint i (angle qangle[i]); occupied = CurrentFrame.pts already set
int ref_search_by_projection_reloc(int desc_type, int dcols, int dtype, void* qdesc, const float* qxy, const float* qsize, const float* qangle,
                                   const unsigned char* qskip, int nq, const kp7* k, void* d, const float* ksize, int n,
                                   const unsigned char* occupied, float minX, float minY, float maxX, float maxY, float radius_th, float size_tol,
                                   float th, int check_ori, int* match_q) {
    Frame::mnMinX = minX; Frame::mnMinY = minY; Frame::mnMaxX = maxX; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (maxX - minX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (maxY - minY);
    FeatureMatcher::descDistTh_low_reloc = th; FeatureMatcher::descDistTh_high_reloc = th; FeatureMatcher::radiusScale = 1.0f;
    Frame F;
    fill_frame(F, k, n, d, dcols, dtype, ksize, 0.f);
    F.sizeTolerance = size_tol; F.invSizeTolerance = 1.0f / size_tol;
    F.pts.assign(n, Pt()); F.mvuRight.assign(n, -1.0f);
    Pt held = std::make_shared<MapPoint>();
    for (int i = 0; i < n; ++i) if (occupied && occupied[i]) F.pts[i] = held;
    Keyframe KF = std::make_shared<KeyFrame>();
    cv::Mat Q(nq, dcols, dtype, qdesc);
    KF->mappoints.resize(nq); KF->mvKeysUn.resize(nq);
    for (int i = 0; i < nq; ++i) { KF->mappoints[i] = make_point(desc_type, Q, i, qxy[2 * i], qxy[2 * i + 1], qsize[i], qskip && qskip[i]); KF->mvKeysUn[i].angle = qangle[i]; }
    std::set<Pt> found;
    FeatureMatcher fm(0.6f, check_ori != 0);
    const int nm = fm.SearchByProjection(F, KF, found, radius_th, false);
    for (int i = 0; i < nq; ++i) match_q[i] = -1;
    for (int idx = 0; idx < n; ++idx) if (F.pts[idx] && F.pts[idx] != held) for (int i = 0; i < nq; ++i) if (F.pts[idx] == KF->mappoints[i]) { match_q[i] = idx; break; }
    return nm;
}
// Fuse(pKF, vpMapPoints, radiusTh) (variant 1, :794-942, with the monocular reprojection gate on inf1d) and
// Fuse(pKF, Scw, vpPoints, radiusTh, vpReplacePoint) (variant 2, :944-1064).  has_mp[idx] != 0: the keyframe keypoint already holds
// a map point (with more observations than any query, so variant 1 always takes pMP->Replace(pMPinKF)).  match_q[i] = keypoint.
int ref_fuse(int variant, int desc_type, int dcols, int dtype, void* qdesc, const float* qxy, const float* qsize, const unsigned char* qskip, int nq,
             const kp7* k, void* d, const float* ksize, const float* inf1d, int n, const unsigned char* has_mp, float minX, float minY, float maxX,
             float maxY, float radius_th, float size_tol, float th_low, int* match_q) {
    FeatureMatcher::TH_LOW = th_low; FeatureMatcher::TH_HIGH = th_low; FeatureMatcher::radiusScale = 1.0f;
    Keyframe KF = make_kf(k, n, d, dcols, dtype, ksize, minX, minY, maxX, maxY, size_tol);
    if (inf1d) KF->inf_1d.assign(inf1d, inf1d + n);
    for (int i = 0; i < n; ++i) if (has_mp && has_mp[i]) { KF->mappoints[i] = std::make_shared<MapPoint>(); KF->mappoints[i]->nobs = 1000; KF->mappoints[i]->idxInKF2 = i; }
    cv::Mat Q(nq, dcols, dtype, qdesc);
    std::vector<Pt> q(nq), repl(nq);
    for (int i = 0; i < nq; ++i) q[i] = make_point(desc_type, Q, i, qxy[2 * i], qxy[2 * i + 1], qsize[i], qskip && qskip[i]);
    FeatureMatcher fm(0.6f, true);
    int nf;
    if (variant == 1) nf = fm.Fuse(KF, q, radius_th);
    else { mat4f Scw; nf = fm.Fuse(KF, Scw, q, radius_th, repl); }
    for (int i = 0; i < nq; ++i) {
        match_q[i] = -1;
        if (q[i]->addedObsIdx >= 0) match_q[i] = q[i]->addedObsIdx;
        else if (variant == 1 && q[i]->replacedBy) match_q[i] = q[i]->replacedBy->idxInKF2;
        else if (variant == 2 && repl[i]) match_q[i] = repl[i]->idxInKF2;
    }
    return nf;
}
// SearchBySim3 (:1066-1287): identity relative pose, scale 1; KF1 keypoint i1 carries a map point that projects to q1xy[i1] in KF2
// (skip = no map point), and vice versa
int ref_search_by_sim3(int desc_type, int dcols, int dtype, const kp7* k1, void* d1, const float* size1, const float* q1xy, const float* q1size,
                       const unsigned char* q1skip, int n1, const kp7* k2, void* d2, const float* size2, const float* q2xy, const float* q2size,
                       const unsigned char* q2skip, int n2, float minX, float minY, float maxX, float maxY, float radius_th, float size_tol,
                       float th_high, int* match12) {
    FeatureMatcher::TH_LOW = th_high; FeatureMatcher::TH_HIGH = th_high; FeatureMatcher::radiusScale = 1.0f;
    Keyframe KF1 = make_kf(k1, n1, d1, dcols, dtype, size1, minX, minY, maxX, maxY, size_tol);
    Keyframe KF2 = make_kf(k2, n2, d2, dcols, dtype, size2, minX, minY, maxX, maxY, size_tol);
    cv::Mat Q1(n1, dcols, dtype, d1), Q2(n2, dcols, dtype, d2);
    // the map point's own descriptor is the keypoint's descriptor here (GetDescriptor of a single-observation point)
    for (int i = 0; i < n1; ++i) if (!(q1skip && q1skip[i])) KF1->mappoints[i] = make_point(desc_type, Q1, i, q1xy[2 * i], q1xy[2 * i + 1], q1size[i], false);
    for (int i = 0; i < n2; ++i) if (!(q2skip && q2skip[i])) KF2->mappoints[i] = make_point(desc_type, Q2, i, q2xy[2 * i], q2xy[2 * i + 1], q2size[i], false);
    std::vector<Pt> m12(n1);
    FeatureMatcher fm(0.6f, true);
    mat3f R12; vec3f t12;
    const int nf = fm.SearchBySim3(KF1, KF2, m12, 1.0f, R12, t12, radius_th);
    for (int i = 0; i < n1; ++i) {
        match12[i] = -1;
        if (m12[i]) for (int j = 0; j < n2; ++j) if (KF2->mappoints[j] == m12[i]) { match12[i] = j; break; }
    }
    return nf;
}
// SearchByBoW(pKF1, pKF2, vpMatches12) (:561-660): FeatureVectors from per-feature node ids; valid = keypoint holds a good map point
int ref_search_by_bow_kfkf(int desc_type, int dcols, int dtype, const kp7* k1, void* d1, const int* node1, const unsigned char* valid1, int n1,
                           const kp7* k2, void* d2, const int* node2, const unsigned char* valid2, int n2, float th_low, float nnratio, int check_ori,
                           int* match12) {
    FeatureMatcher::TH_LOW = th_low; FeatureMatcher::TH_HIGH = th_low;
    std::vector<float> s1(n1, 1.0f), s2(n2, 1.0f);
    Keyframe KF1 = make_kf(k1, n1, d1, dcols, dtype, s1.data(), 0, 0, 640, 480, 1.5f);
    Keyframe KF2 = make_kf(k2, n2, d2, dcols, dtype, s2.data(), 0, 0, 640, 480, 1.5f);
    for (int i = 0; i < n1; ++i) if (!valid1 || valid1[i]) { KF1->mappoints[i] = std::make_shared<MapPoint>(); KF1->mappoints[i]->descriptorType = (DescriptorType)desc_type; }
    for (int i = 0; i < n2; ++i) if (!valid2 || valid2[i]) { KF2->mappoints[i] = std::make_shared<MapPoint>(); KF2->mappoints[i]->descriptorType = (DescriptorType)desc_type; }
    feat_vec(KF1->mFeatVec, node1, n1); feat_vec(KF2->mFeatVec, node2, n2);
    std::vector<Pt> m12;
    FeatureMatcher fm(nnratio, check_ori != 0);
    const int nm = fm.SearchByBoW(KF1, KF2, m12);
    for (int i = 0; i < n1; ++i) {
        match12[i] = -1;
        if (m12[i]) for (int j = 0; j < n2; ++j) if (KF2->mappoints[j] == m12[i]) { match12[i] = j; break; }
    }
    return nm;
}
// SearchForTriangulation (:662-790, monocular): has_mp = keypoint already holds a map point; (ex, ey) = epipole in image 2 (the
// reference computes it from KF1's camera centre: identity pose of KF2 and centre (ex, ey, 1)); sigma2 = GetKeyPt1DSigma2 of KF2
int ref_search_for_triangulation(int desc_type, int dcols, int dtype, const kp7* k1, void* d1, const int* node1, const unsigned char* has_mp1, int n1,
                                 const kp7* k2, void* d2, const int* node2, const unsigned char* has_mp2, const float* sigma2, int n2,
                                 const float* F12, float ex, float ey, float th_low, int* match12) {
    FeatureMatcher::TH_LOW = th_low; FeatureMatcher::TH_HIGH = th_low;
    std::vector<float> s1(n1, 1.0f), s2(n2, 1.0f);
    Keyframe KF1 = make_kf(k1, n1, d1, dcols, dtype, s1.data(), 0, 0, 640, 480, 1.5f);
    Keyframe KF2 = make_kf(k2, n2, d2, dcols, dtype, s2.data(), 0, 0, 640, 480, 1.5f);
    for (int i = 0; i < n1; ++i) if (has_mp1 && has_mp1[i]) KF1->mappoints[i] = std::make_shared<MapPoint>();
    for (int i = 0; i < n2; ++i) if (has_mp2 && has_mp2[i]) KF2->mappoints[i] = std::make_shared<MapPoint>();
    KF2->sigma2_1d.assign(sigma2, sigma2 + n2);
    KF1->Ow(0) = ex; KF1->Ow(1) = ey; KF1->Ow(2) = 1.0f;
    feat_vec(KF1->mFeatVec, node1, n1); feat_vec(KF2->mFeatVec, node2, n2);
    mat3f F;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) F.m[i][j] = F12[3 * i + j];
    std::vector<std::pair<size_t, size_t>> pairs;
    FeatureMatcher fm(0.6f, false);
    const int nm = fm.SearchForTriangulation(KF1, KF2, F, pairs, false, (DescriptorType)desc_type);
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    for (auto& pr : pairs) match12[pr.first] = (int)pr.second;
    return nm;
}
// vanilla ORB-SLAM2 FeatureExtractor::operator()(..., vanillaOrbslam) (src/ORBextractor.cc:568-645) on one gray frame; the OpenCV
// functions it calls (FAST, resize INTER_LINEAR, GaussianBlur, fastAtan2) are the harness callbacks (the tests pass cv2's real ones)
int ref_orbslam2_extract(const unsigned char* gray, int w, int h, int nfeatures, int nlevels, float scale_factor, int ini_th, int min_th,
                         int (*fast_cb)(const unsigned char*, int, int, int, int, float*, int),
                         void (*resize_cb)(const unsigned char*, int, int, int, unsigned char*, int, int, int),
                         void (*blur_cb)(unsigned char*, int, int, int), float (*atan2_cb)(float, float),
                         kp7* okps, unsigned char* odesc, float* osize, int cap) {
    if (g_bump) arena_reset();
    cv::g_vanilla.fast = fast_cb; cv::g_vanilla.resize = resize_cb; cv::g_vanilla.blur = blur_cb; cv::g_vanilla.atan2 = atan2_cb;
    std::shared_ptr<FeatureExtractorSettings> st = std::make_shared<FeatureExtractorSettings>();
    st->scaleFactor = scale_factor; st->nOctaves = nlevels; FeatureExtractorSettings::scaleFactor0 = scale_factor;
    st->iniThFAST = ini_th; st->minThFAST = min_th; st->detectTh = (float)ini_th;
    st->maxKeyPtSize0 = pow(1.2f, float(8 - 1.0)); st->maxKeyPtSize = st->maxKeyPtSize0; st->minKeyPtSize = 1.0f;      // src/FeatureExtractor.cpp:52-55
    int m = 0;
    {
        FeatureExtractor fe;
        fe.initVanilla(nfeatures, st);
        Image img; img.grayImg = cv::Mat(h, w, CV_8U, (void*)gray);
        std::vector<cv::KeyPoint> k; cv::Mat d; std::vector<float> sz; std::vector<mat2f> s2, inf;
        fe(img, k, d, s2, inf, sz, true);
        // second call: computeSize then sees the settings' maxKeyPtSize / minKeyPtSize as the first frame left them (steady state)
        fe(img, k, d, s2, inf, sz, true);
        m = (int)k.size();
        for (int i = 0; i < m && i < cap; ++i) {
            okps[i].x = k[i].pt.x; okps[i].y = k[i].pt.y; okps[i].size = k[i].size; okps[i].angle = k[i].angle; okps[i].response = k[i].response;
            okps[i].octave = k[i].octave; okps[i].class_id = k[i].class_id; osize[i] = sz[i];
            memcpy(odesc + (size_t)i * 32, d.data + (size_t)i * d.step, 32);
        }
    }
    return m;
}
// Frame::isInFrustum (src/Frame.cc:276-331) for M map points; Eigen is replaced by the shim's left-to-right float look-alikes
void ref_is_in_frustum(const float* Pw, const float* normal, const float* min_dist, const float* max_dist, const float* ref_size,
                       const float* ref_sigma, const float* ref_dist, int M, const float* pose16, const float* cam5, const float* bounds4,
                       float viewing_cos_limit, unsigned char* in_view, float* proj3, float* track3) {
    Frame::mnMinX = bounds4[0]; Frame::mnMaxX = bounds4[1]; Frame::mnMinY = bounds4[2]; Frame::mnMaxY = bounds4[3];
    Frame F;
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) F.Rcw.m[i][j] = pose16[3 * i + j]; F.tcw.v[i] = pose16[9 + i]; F.twc.v[i] = pose16[12 + i]; }
    F.fx = cam5[0]; F.fy = cam5[1]; F.cx = cam5[2]; F.cy = cam5[3]; F.mbf = cam5[4];
    for (int i = 0; i < M; ++i) {
        Pt p = std::make_shared<MapPoint>();
        for (int k = 0; k < 3; ++k) { p->worldPos.v[k] = Pw[3 * i + k]; p->normal.v[k] = normal[3 * i + k]; }
        p->minDist = min_dist[i]; p->maxDist = max_dist[i]; p->realPredict = true;
        p->refSize = ref_size[i]; p->refSigma = ref_sigma[i]; p->refDistance = ref_dist[i];
        p->mTrackProjX = p->mTrackProjY = p->mTrackProjXR = 0; p->trackSize = p->trackSigma = p->trackViewCos = 0;
        in_view[i] = F.isInFrustum(p, viewing_cos_limit) ? 1 : 0;
        proj3[3 * i] = p->mTrackProjX; proj3[3 * i + 1] = p->mTrackProjY; proj3[3 * i + 2] = p->mTrackProjXR;
        track3[3 * i] = p->trackSize; track3[3 * i + 1] = p->trackSigma; track3[3 * i + 2] = p->trackViewCos;
    }
}
float ref_descriptor_distance(int desc_type, int dcols, int dtype, void* a, void* b) {
    return FeatureMatcher::DescriptorDistance(cv::Mat(1, dcols, dtype, a), cv::Mat(1, dcols, dtype, b), (DescriptorType)desc_type);
}
}
'''


def build(force=False):
    if not os.path.isdir(REF):
        return None                                        # e.g. on the GPU box: only the prebuilt library is used
    os.makedirs(OUT_DIR, exist_ok=True)
    gen = os.path.join(OUT_DIR, "ref_extract.cpp")
    if not force and os.path.exists(OUT_SO) and os.path.getmtime(OUT_SO) > max(os.path.getmtime(__file__), os.path.getmtime(os.path.join(HERE, "ref_shim.hpp"))):
        return OUT_SO
    parts = ['#include "../ref_shim.hpp"', "namespace ANYFEATURE_VSLAM {"]
    # steered BRIEF of the reference header (include/FeatureExtractor.h:177-217) and its 256 x 4 pattern table (:219-477)
    parts.append("const float factorPI = (float)(CV_PI/180.f);")
    parts.append(cut("include/FeatureExtractor.h", r"^\s*static void computeOrbDescriptor\("))
    parts.append(cut("include/FeatureExtractor.h", r"^\s*static int bit_pattern_31_\[256\*4\] =") + ";")
    parts.append(cut("src/ORBextractor.cc", r"^void ExtractorNode::DivideNode\("))
    parts.append(cut("src/ORBextractor.cc", r"^vector<cv::KeyPoint> FeatureExtractor::DistributeOctTree\("))
    parts.append(cut("src/FeatureExtractor.cpp", r"^ANYFEATURE_VSLAM::FeatureExtractor::FeatureExtractor\(const int& nfeatures_"))
    parts.append(cut("src/FeatureExtractor.cpp", r"^void ANYFEATURE_VSLAM::FeatureExtractor::computeSize\("))
    parts.append(cut("src/Frame.cc", r"^void Frame::AssignFeaturesToGrid\(\)"))
    parts.append(cut("src/Frame.cc", r"^vector<size_t> Frame::GetFeaturesInArea\("))
    parts.append(cut("src/Frame.cc", r"^bool Frame::PosInGrid\("))
    parts.append(cut("src/Frame.cc", r"^bool Frame::isInFrustum\("))
    parts.append(cut("src/MapPoint.cc", r"^float MapPoint::PredictSize\(").replace("MapPoint::PredictSize(", "MapPoint::PredictSizeRef("))
    parts.append(cut("src/MapPoint.cc", r"^float MapPoint::PredictSigma\(").replace("MapPoint::PredictSigma(", "MapPoint::PredictSigmaRef("))
    parts.append(cut("src/FeatureMatcher.cc", r"^int FeatureMatcher::SearchForInitialization\("))
    parts.append(cut("src/FeatureMatcher.cc", r"^int FeatureMatcher::SearchByProjection\(Frame &F, const vector<Pt> &vpMapPoints, const float& radiusTh\)"))
    parts.append(cut("src/FeatureMatcher.cc", r"^float FeatureMatcher::RadiusByViewingCos\("))
    parts.append(cut("src/FeatureMatcher.cc", r"^int FeatureMatcher::SearchByProjection\(Frame &CurrentFrame, const Frame &LastFrame, const float& radiusTh, const bool bMono\)"))
    parts.append(cut("src/FeatureMatcher.cc", r"^int FeatureMatcher::SearchByBoW\(Keyframe pKF, Frame &F, vector<Pt> &vpMapPointMatches\)"))
    # rows a19 / a20: the remaining searches + the KeyFrame helpers they call
    parts.append(cut("src/FeatureMatcher.cc", r"^int FeatureMatcher::SearchByProjection\(Keyframe pKF, const mat4f& Scw, const vector<Pt> &vpPoints, vector<Pt> &vpMatched, const float& radiusTh\)"))
    parts.append(cut("src/FeatureMatcher.cc", r"^int FeatureMatcher::SearchByProjection\(Frame &CurrentFrame, Keyframe pKF, const set<Pt> &sAlreadyFound, const float& radiusTh, const bool& useHighMatchingThreshold\)"))
    parts.append(cut("src/FeatureMatcher.cc", r"^int FeatureMatcher::SearchByBoW\(Keyframe pKF1, Keyframe pKF2, vector<Pt > &vpMatches12\)"))
    parts.append(cut("src/FeatureMatcher.cc", r"^bool FeatureMatcher::CheckDistEpipolarLine\("))
    parts.append(cut("src/FeatureMatcher.cc", r"^int FeatureMatcher::SearchForTriangulation\("))
    parts.append(cut("src/FeatureMatcher.cc", r"^int FeatureMatcher::Fuse\(Keyframe pKF, const vector<Pt> &vpMapPoints, const float& radiusTh\)"))
    parts.append(cut("src/FeatureMatcher.cc", r"^int FeatureMatcher::Fuse\(Keyframe pKF, const mat4f& Scw, const vector<Pt> &vpPoints, const float& radiusTh, vector<Pt> &vpReplacePoint\)"))
    parts.append(cut("src/FeatureMatcher.cc", r"^int FeatureMatcher::SearchBySim3\("))
    parts.append(cut("src/KeyFrame.cc", r"^vector<size_t> KeyFrame::GetFeaturesInArea\("))
    parts.append(cut("src/KeyFrame.cc", r"^bool KeyFrame::IsInImage\("))
    parts.append(cut("src/FeatureMatcher.cc", r"^Descriptor_Distance_Type FeatureMatcher::DescriptorDistance\("))
    parts.append(cut("src/FeatureMatcher.cc", r"^\s*vector<vector<int>> FeatureMatcher::initRotationHistogram\("))
    parts.append(cut("src/FeatureMatcher.cc", r"^\s*void FeatureMatcher::updateRotationHistogram\("))
    parts.append(cut("src/FeatureMatcher.cc", r"^\s*void FeatureMatcher::filterMatchesWithOrientation\(", all_matches=True))
    parts.append(cut("src/FeatureMatcher.cc", r"^\s*void FeatureMatcher::computeThreeMaxima\("))
    # ---- vanilla ORB-SLAM2 extractor: src/ORBextractor.cc:79-177 and :460-676 as built with VANILLA_ORB_SLAM2.  The
    # constructor is cut and turned into a member function (the shim's class already has the default build's constructor).
    ctor = cut("src/ORBextractor.cc", r"^\s*FeatureExtractor::FeatureExtractor\(const int& nfeatures_, shared_ptr<FeatureExtractorSettings>& settings_\):\s*\n\s*nfeatures\(nfeatures_\), settings\(settings_\)")
    body = ctor[ctor.index("{") + 1:]
    parts.append("void FeatureExtractor::initVanilla(const int& nfeatures_, shared_ptr<FeatureExtractorSettings>& settings_)\n{\n    nfeatures = nfeatures_; settings = settings_;" + body)
    parts.append(cut("src/ORBextractor.cc", r"^static float IC_Angle\("))
    parts.append(cut("src/ORBextractor.cc", r"^static void computeOrientation\("))
    parts.append(cut("src/ORBextractor.cc", r"^void FeatureExtractor::ComputeKeyPointsOctTree\("))
    parts.append(cut("src/ORBextractor.cc", r"^static void computeDescriptorsORB\("))
    parts.append(cut("src/ORBextractor.cc", r"^void FeatureExtractor::operator\(\)\(const Image & img, vector<KeyPoint>& _keypoints, OutputArray _descriptors,"))
    parts.append(cut("src/ORBextractor.cc", r"^void FeatureExtractor::ComputePyramid\("))
    parts.append("}  // namespace ANYFEATURE_VSLAM")
    parts.append(cut("src/FeatureExtractor.cpp", r"^void ANYFEATURE_VSLAM::FeatureExtractor::computeSigma\("))
    # ---- the reference-side glue of the sift128 / akaze61 extractors (everything around the un-vendored third-party library)
    parts.append(cut("src/FeatureExtractor.cpp", r"^void ANYFEATURE_VSLAM::FeatureExtractor::filterKeypoints_notScaled\("))
    parts.append(cut("src/FeatureExtractor.cpp", r"^void ANYFEATURE_VSLAM::FeatureExtractor::mergeKeypointLevels\("))
    parts.append(cut("src/Feature_orb32.cpp", r"^ANYFEATURE_VSLAM::FeatureExtractor_orb32::FeatureExtractor_orb32\("))
    parts.append(cut("src/Feature_orb32.cpp", r"^void ANYFEATURE_VSLAM::FeatureExtractor_orb32::initializeExtractor\("))
    for feat in ("orb32", "sift128", "akaze61"):
        for fn in ("detectAndCompute", "detectKeypoints", "computeDescriptors", "filterKeypoints", "GetKeypointOctave", "GetKeypointSize"):
            rt = {"GetKeypointOctave": "int", "GetKeypointSize": "float"}.get(fn, "void")
            parts.append(cut("src/Feature_%s.cpp" % feat, r"^%s ANYFEATURE_VSLAM::FeatureExtractor_%s::%s\(" % (rt, feat, fn)))
    for feat in ("orb32", "akaze61", "brisk48", "sift128"):
        parts.append(cut("src/Feature_%s.cpp" % feat, r"^float ANYFEATURE_VSLAM::DescriptorDistance_%s\(" % feat))
    # ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:279-348): the distance matrix + least-median pick, i.e. the
    # function body between the collection of the observed descriptors and the locked write-back
    parts.append("namespace ANYFEATURE_VSLAM {\nint ref_distinctive_core(const std::vector<cv::Mat>& descriptors, DescriptorType descriptorType) {")
    parts.append(cut_between("src/MapPoint.cc", "    // Compute distances between them", "    {\n        unique_lock<mutex> lock(mMutexFeatures);\n        mDescriptor = descriptors[BestIdx].clone();"))
    parts.append("    return BestIdx;\n}\n}  // namespace ANYFEATURE_VSLAM")
    # ---- DBoW2 (vendored by the reference under Thirdparty/DBoW2): tree descent + the per-feature distances it uses
    parts.append("""
#include <climits>
namespace DBoW2 {
typedef unsigned int WordId; typedef double WordValue; typedef unsigned int NodeId;
template <class TDescriptor, class F> class TemplatedVocabulary {     // TemplatedVocabulary.h: the members transform() reads
public:
    struct Node { NodeId id; WordValue weight; std::vector<NodeId> children; NodeId parent; TDescriptor descriptor; WordId word_id;
                  Node() : id(0), weight(0), parent(0), word_id(0) {} inline bool isLeaf() const { return children.empty(); } };
    int m_L; std::vector<Node> m_nodes;
    void transform(const TDescriptor& feature, WordId& word_id, WordValue& weight, NodeId* nid, int levelsup) const;
};
struct FOrb { typedef cv::Mat TDescriptor; static const int L = 32; static double distance(const TDescriptor& a, const TDescriptor& b); };
struct FAkaze61 { typedef cv::Mat TDescriptor; static const int L = 61; static double distance(const TDescriptor& a, const TDescriptor& b); };
struct FBrisk { typedef cv::Mat TDescriptor; static const int L = 48; static double distance(const TDescriptor& a, const TDescriptor& b); };
struct FSift128 { typedef std::vector<float> TDescriptor; static const int L = 128; static double distance(const TDescriptor& a, const TDescriptor& b); };
""")
    parts.append(cut("Thirdparty/DBoW2/include/DBoW2/TemplatedVocabulary.h",
                     r"^template<class TDescriptor, class F>\nvoid TemplatedVocabulary<TDescriptor,F>::transform\(const TDescriptor &feature, \n\s*WordId &word_id, WordValue &weight, NodeId \*nid, int levelsup\) const"))
    parts.append(cut("Thirdparty/DBoW2/src/FOrb.cpp", r"^\s*double FOrb::distance\("))
    parts.append(cut("Thirdparty/DBoW2/src/FAkaze61.cpp", r"^\s*double FAkaze61::distance\("))
    parts.append(cut("Thirdparty/DBoW2/src/FBrisk.cpp", r"^\s*double FBrisk::distance\("))
    parts.append(cut("Thirdparty/DBoW2/src/FSift128.cpp", r"^\s*double FSift128::distance\("))
    parts.append("}  // namespace DBoW2")
    parts.append(DRIVERS)
    open(gen, "w").write("\n\n".join(parts) + "\n")
    # same optimisation level family as the reference's CMakeLists.txt:16 (-O3 -march=native); -ffp-contract=off: with GCC's default
    # contraction the float gates of the matcher (epipolar distance, reprojection chi2) would round differently per host CPU, i.e. the
    # reference would not be a reproducible checker; IEEE evaluation without fusing is the contract the oracle restates
    cmd = ["g++", "-std=c++17", "-O3", "-march=native", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-Wl,-Bsymbolic", "-o", OUT_SO, gen]
    subprocess.check_call(cmd)
    return OUT_SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

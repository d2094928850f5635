// afv_capi.cu -- C ABI (include/afv.h) of the extractor: handle, device arenas, batch orchestration.
//
// Host-side mirror of what the reference's FeatureExtractor constructor prepares once
// (reference src/FeatureExtractor.cpp:74-109: scale factors, per-level quotas) plus cv::ORB's layer geometry
// and the INTER_LINEAR_EXACT coefficient tables; everything per frame runs in afv_orb.cu kernels.
#include "afv_common.cuh"
#include "afv_sift.h"
#include "afv_akaze.h"
#include "afv_brisk.h"
#include "afv_orbslam2.h"
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

std::atomic<long long> g_afv_launches{0};
static thread_local char g_err[512] = "";

void afv_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* afv_last_error(void) { return g_err; }
extern "C" long long afv_kernel_launches(void) { return g_afv_launches.load(); }
extern "C" const char* afv_version(void) { return "afv-b200 0.1 (sm_100a)"; }

// ---- per-kernel event timing --------------------------------------------------------------------------
#include <map>
#include <string>
#include <mutex>
static std::atomic<bool> g_prof_on{false};
struct ProfRec { std::string name; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;                         // records and event pool (matcher calls arrive from several host threads)
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
int afv_prof_begin(const char* name, cudaStream_t st) {
    if (!g_prof_on.load()) return -1;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r; r.name = name; r.e0 = prof_event(); r.e1 = prof_event();
    cudaEventRecord(r.e0, st);
    g_prof_recs.push_back(r);
    return (int)g_prof_recs.size() - 1;
}
void afv_prof_end(int id, cudaStream_t st) {
    if (id < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (id < (int)g_prof_recs.size()) cudaEventRecord(g_prof_recs[id].e1, st);
}
bool afv_prof_is_on() { return g_prof_on.load(); }
extern "C" int afv_profile_enable(int on) { g_prof_on.store(on != 0); return AFV_OK; }
// Synchronises, sums the recorded launches per kernel name and clears the records.
// names: max_n x 32 chars; ms / calls: max_n.  Returns the number of distinct kernels.
extern "C" int afv_profile_read(char* names, float* ms, int* calls, int max_n) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<double, int>> acc;
    std::vector<std::string> order;
    for (auto& r : g_prof_recs) {
        cudaEventSynchronize(r.e1);
        float t = 0; cudaEventElapsedTime(&t, r.e0, r.e1);
        if (!acc.count(r.name)) order.push_back(r.name);
        acc[r.name].first += t; acc[r.name].second += 1;
        g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1);
    }
    g_prof_recs.clear();
    int n = 0;
    for (auto& k : order) {
        if (n >= max_n) break;
        strncpy(names + 32 * n, k.c_str(), 31); names[32 * n + 31] = 0;
        ms[n] = (float)acc[k].first; calls[n] = acc[k].second; ++n;
    }
    return n;
}

static inline int round_half_even_f(float v) { return (int)lrintf(v); }     // cvRound

struct afv_extractor {
    int feature_id, nfeatures, nlevels, device;
    float scale_factor, detect_th;
    int max_batch, max_w, max_h;
    int cur_w, cur_h;                 // geometry the tables / params are currently built for
    cudaStream_t stream;
    cudaEvent_t ev_order;             // orders the private stream after the legacy default stream (NULL-stream calls)
    AfvAux aux;
    std::vector<void*> allocs;        // everything to cudaFree
    uint8_t* pyr[AFV_MAX_LEVELS];     // un-blurred arena per level (level 0 = staging copy of the input)
    uint8_t* blur[AFV_MAX_LEVELS];
    uint32_t* cand[AFV_MAX_LEVELS]; float* cand_resp[AFV_MAX_LEVELS];
    uint2* det[AFV_MAX_LEVELS]; uint2* keep[AFV_MAX_LEVELS];
    uint32_t* tab[AFV_MAX_LEVELS];    // xtab | ytab (offset << 16 | c1)
    int* counts; int* status;
    AfvTile* tiles; int tiles_cap;
    afv_keypoint* o_kps; uint8_t* o_desc; float* o_size; int* o_n;   // device outputs for the host-buffer API
    int* h_status; int* h_counts;     // pinned
    int max_stride[AFV_MAX_LEVELS], max_lh[AFV_MAX_LEVELS], cand_cap[AFV_MAX_LEVELS], det_cap[AFV_MAX_LEVELS], keep_cap[AFV_MAX_LEVELS];
    int q_orb[AFV_MAX_LEVELS], q_ext[AFV_MAX_LEVELS];
    float scale[AFV_MAX_LEVELS], size_norm[AFV_MAX_LEVELS], ext_scale[AFV_MAX_LEVELS];
    AfvParams P;
    int oct_mcap, oct_ncap;           // this extractor's k_octree capacities
    int last_B;
    cudaStream_t last_stream;
    AfvSift* sift;                    // sift128 state (feature_id == AFV_FEAT_SIFT128), else NULL
    AfvAkaze* akaze;                  // akaze61 state (feature_id == AFV_FEAT_AKAZE61), else NULL
    AfvBrisk* brisk;                  // brisk48 state (feature_id == AFV_FEAT_BRISK48), else NULL
    AfvOs2* os2;                      // vanilla ORB-SLAM2 state (feature_id == AFV_FEAT_ORB32_VANILLA), else NULL
    int desc_bytes;                   // bytes per descriptor row: 32 (orb32) / 61 (akaze61) / 512 (sift128: 128 floats)
};

// reference src/FeatureExtractor.cpp:97-108 (same formula inside cv::ORB for its maxFeatures quota)
static void features_per_level(int nfeatures, int nlevels, float scale_factor, int* quota) {
    float factor = 1.0f / scale_factor;
    float nDesired = (float)nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        quota[l] = round_half_even_f(nDesired);
        sum += quota[l];
        nDesired *= factor;
    }
    quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
}

// OpenCV resize_bitExact / interpolationLinear coefficients (8.8 fixed point), see oracle for the derivation.
static void lin_exact_tables(int ssize, int dsize, uint16_t* ofs, uint16_t* c1) {
    double inv_scale = (double)dsize / (double)ssize;
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        double f = scale * ((double)d + 0.5) - 0.5;
        int i = (int)floor(f);
        if (i >= 0 && ssize > 1) {
            if (i < ssize - 1) { ofs[d] = (uint16_t)i; c1[d] = (uint16_t)lrint((f - (double)i) * 256.0); }
            else { ofs[d] = (uint16_t)(ssize - 1); c1[d] = 0; }
        } else { ofs[d] = 0; c1[d] = 0; }
    }
}

static void level_geometry(int w, int h, int nlevels, int* lw, int* lh, float* lscale) {
    const float orb_sf = 1.2f;      // cv::ORB::create() default, never overridden (reference src/Feature_orb32.cpp:20-24)
    for (int l = 0; l < nlevels; ++l) {
        float s = (float)pow((double)orb_sf, (double)l);
        float inv = 1.0f / s;
        lscale[l] = s;
        lw[l] = round_half_even_f((float)w * inv);
        lh[l] = round_half_even_f((float)h * inv);
    }
}

template <typename T>
static int dev_alloc(afv_extractor* ex, T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 256);
    if (e != cudaSuccess) { afv_set_error("cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e)); return AFV_ERR_CUDA; }
    ex->allocs.push_back(q);
    *p = (T*)q;
    return AFV_OK;
}

// (re)build geometry-dependent tables and the kernel parameter block for a w x h input
static int configure_geometry(afv_extractor* ex, int w, int h) {
    if (w == ex->cur_w && h == ex->cur_h) return AFV_OK;
    if (w > ex->max_w || h > ex->max_h || w < 64 || h < 64) {
        afv_set_error("frame %dx%d outside the extractor's configured range (64..%d x 64..%d)", w, h, ex->max_w, ex->max_h);
        return AFV_ERR_INVALID;
    }
    if (w > 4095 || h > 4095) { afv_set_error("frame dimension > 4095 not supported by the packed candidate format"); return AFV_ERR_INVALID; }
    int lw[AFV_MAX_LEVELS], lh[AFV_MAX_LEVELS];
    level_geometry(w, h, ex->nlevels, lw, lh, ex->scale);
    AfvParams& P = ex->P;
    memset(&P, 0, sizeof(P));
    P.nlevels = ex->nlevels; P.W = w; P.H = h;
    P.fast_th = (int)ex->detect_th;                       // reference src/Feature_orb32.cpp:30 int(detectTh)
    if (P.fast_th < 0) P.fast_th = 0;
    if (P.fast_th > 255) P.fast_th = 255;
    P.n_ini = (int)round((double)((float)w / (float)h));  // reference src/ORBextractor.cc:243
    if (P.n_ini < 1) { afv_set_error("portrait frames with w/h < 0.5 are not supported (reference divides by zero)"); return AFV_ERR_INVALID; }
    P.hX = (float)w / (float)P.n_ini;
    P.counts = ex->counts; P.status = ex->status;
    P.oct_mcap = ex->oct_mcap; P.oct_ncap = ex->oct_ncap;
    for (int l = 0; l < ex->nlevels; ++l) {
        AfvLevel& L = P.lv[l];
        L.w = lw[l]; L.h = lh[l];
        L.stride = (lw[l] + 127) & ~127;
        if (L.stride > ex->max_stride[l] || lh[l] > ex->max_lh[l]) { afv_set_error("internal: level %d larger than arena", l); return AFV_ERR_INVALID; }
        L.fstride = (long long)L.stride * lh[l];
        L.img = ex->pyr[l]; L.img_stride = L.stride; L.img_fstride = L.fstride;
        L.blur = ex->blur[l];
        L.scale = ex->scale[l]; L.inv_scale = 1.f / ex->scale[l];
        L.kp_size = 31 * ex->scale[l];
        L.size_norm = ex->size_norm[l];
        L.q_orb = ex->q_orb[l]; L.q_ext = ex->q_ext[l];
        L.cand_cap = ex->cand_cap[l]; L.cand = ex->cand[l]; L.cand_resp = ex->cand_resp[l];
        L.det_cap = ex->det_cap[l]; L.det = ex->det[l];
        L.keep_cap = ex->keep_cap[l]; L.keep = ex->keep[l];
        if (l > 0) {
            const int wpad = (lw[l] + 3) & ~3;
            std::vector<uint16_t> xo(lw[l]), xc(lw[l]), yo(lh[l]), yc(lh[l]);
            lin_exact_tables(lw[l - 1], lw[l], xo.data(), xc.data());
            lin_exact_tables(lh[l - 1], lh[l], yo.data(), yc.data());
            std::vector<uint32_t> t((size_t)wpad + lh[l]);
            for (int x = 0; x < wpad; ++x) { const int xx = x < lw[l] ? x : lw[l] - 1; t[x] = ((uint32_t)xo[xx] << 16) | xc[xx]; }
            for (int y = 0; y < lh[l]; ++y) t[wpad + y] = ((uint32_t)yo[y] << 16) | yc[y];
            AFV_CUDA_CHECK(cudaMemcpy(ex->tab[l], t.data(), t.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            L.xtab = ex->tab[l]; L.ytab = ex->tab[l] + wpad;
        }
    }
    {
        std::vector<AfvTile> tl;
        for (int l = 0; l < ex->nlevels; ++l)
            for (int y = 0; y < lh[l]; y += AFV_TILE_H)
                for (int x = 0; x < lw[l]; x += AFV_TILE_W) { AfvTile t; t.level = (short)l; t.pad = 0; t.x0 = (short)x; t.y0 = (short)y; tl.push_back(t); }
        if ((int)tl.size() > ex->tiles_cap) { afv_set_error("internal: tile table too small"); return AFV_ERR_INVALID; }
        AFV_CUDA_CHECK(cudaMemcpy(ex->tiles, tl.data(), tl.size() * sizeof(AfvTile), cudaMemcpyHostToDevice));
        P.tiles = ex->tiles; P.ntiles = (int)tl.size();
    }
    ex->cur_w = w; ex->cur_h = h;
    return AFV_OK;
}

extern "C" int afv_extractor_create(afv_extractor** out, int feature_id, int nfeatures, int n_octaves,
                                    float scale_factor, float detect_th, int device,
                                    int max_batch, int max_w, int max_h) {
    if (!out) { afv_set_error("out is NULL"); return AFV_ERR_INVALID; }
    *out = nullptr;
    if (feature_id != AFV_FEAT_ORB32 && feature_id != AFV_FEAT_SIFT128 && feature_id != AFV_FEAT_AKAZE61 && feature_id != AFV_FEAT_BRISK48 &&
        feature_id != AFV_FEAT_ORB32_VANILLA) {
        afv_set_error("feature id %d: no extractor for this feature (orb32, akaze61, brisk48, sift128 are built)", feature_id);
        return AFV_ERR_UNSUPPORTED;
    }
    if (feature_id == AFV_FEAT_SIFT128 && n_octaves > 15) { afv_set_error("sift128: at most 15 levels"); return AFV_ERR_INVALID; }
    if (nfeatures < 1 || n_octaves < 1 || n_octaves > AFV_MAX_LEVELS || max_batch < 1 || max_w < 64 || max_h < 64 || !(scale_factor > 1.0f)) {
        afv_set_error("invalid extractor arguments"); return AFV_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        afv_set_error("no usable CUDA device %d (%s): this library has no CPU fallback", device, e == cudaSuccess ? "not present" : cudaGetErrorString(e));
        return AFV_ERR_NO_DEVICE;
    }
    AFV_CUDA_CHECK(cudaSetDevice(device));
    afv_extractor* ex = new afv_extractor();
    ex->feature_id = feature_id; ex->nfeatures = nfeatures; ex->nlevels = n_octaves; ex->device = device;
    ex->scale_factor = scale_factor; ex->detect_th = detect_th;
    ex->max_batch = max_batch; ex->max_w = max_w; ex->max_h = max_h; ex->cur_w = ex->cur_h = 0;
    ex->last_B = 0; ex->last_stream = nullptr;
    ex->sift = nullptr; ex->akaze = nullptr; ex->brisk = nullptr; ex->os2 = nullptr; ex->ev_order = nullptr; ex->stream = nullptr;
    ex->desc_bytes = feature_id == AFV_FEAT_SIFT128 ? 512 : feature_id == AFV_FEAT_AKAZE61 ? 61 : feature_id == AFV_FEAT_BRISK48 ? 48 : 32;
    ex->h_status = nullptr; ex->h_counts = nullptr; ex->aux.stream = nullptr; ex->aux.ev_pyr = nullptr; ex->aux.ev_blur = nullptr;
    AFV_CUDA_CHECK(cudaStreamCreateWithFlags(&ex->stream, cudaStreamNonBlocking));
    AFV_CUDA_CHECK(cudaEventCreateWithFlags(&ex->ev_order, cudaEventDisableTiming));
    if (feature_id == AFV_FEAT_SIFT128 || feature_id == AFV_FEAT_AKAZE61 || feature_id == AFV_FEAT_BRISK48 || feature_id == AFV_FEAT_ORB32_VANILLA) {
        // vanilla ORB-SLAM2 (reference src/ORBextractor.cc:79-136 constructor, :568-645 operator()): iniThFAST = int(detect_th), minThFAST = 7
        // FeatureExtractor_brisk48 (reference src/Feature_brisk48.cpp:20-48): BriskFeatureDetector(int(detect_th), n_octaves / 2, true).
        // FeatureExtractor_akaze61 (reference src/Feature_akaze61.cpp:7-13): omax = n_octaves / 4, nsublevels = n_octaves / 2,
        // dthreshold = detect_th.
        // FeatureExtractor_sift128 (reference src/Feature_sift128.cpp:9-62): SiftGPU arguments are fixed by the reference;
        // n_octaves / scale_factor only drive mnFeaturesPerLevel and computeSize (settings/sift128_settings.yaml: 8, 2.0)
        for (int l = 0; l < n_octaves; ++l) ex->ext_scale[l] = l == 0 ? 1.0f : ex->ext_scale[l - 1] * scale_factor;
        features_per_level(nfeatures, n_octaves, scale_factor, ex->q_ext);
        int rc = feature_id == AFV_FEAT_SIFT128 ? afv_sift_create(&ex->sift, nfeatures, n_octaves, scale_factor, max_batch, max_w, max_h)
               : feature_id == AFV_FEAT_BRISK48 ? afv_brisk_create(&ex->brisk, nfeatures, n_octaves, scale_factor, detect_th, max_batch, max_w, max_h)
               : feature_id == AFV_FEAT_ORB32_VANILLA ? afv_os2_create(&ex->os2, nfeatures, n_octaves, scale_factor, detect_th, max_batch, max_w, max_h)
                                                : afv_akaze_create(&ex->akaze, nfeatures, n_octaves, scale_factor, detect_th, max_batch, max_w, max_h);
        const int ocap = nfeatures + 3 * n_octaves;
        if (rc == AFV_OK) rc = dev_alloc(ex, &ex->o_kps, (size_t)ocap * max_batch);
        if (rc == AFV_OK) rc = dev_alloc(ex, &ex->o_desc, (size_t)ocap * ex->desc_bytes * max_batch);
        if (rc == AFV_OK) rc = dev_alloc(ex, &ex->o_size, (size_t)ocap * max_batch);
        if (rc == AFV_OK) rc = dev_alloc(ex, &ex->o_n, (size_t)max_batch);
        if (rc != AFV_OK) { afv_extractor_destroy(ex); return rc; }
        *out = ex;
        return AFV_OK;
    }
    {   // the side stream carries the latency-bound selection kernels (k_harris_select, k_octree): highest priority, so their few
        // CTAs are placed ahead of the hundreds of thousands of blur tiles queued on the caller's stream
        int prio_least = 0, prio_greatest = 0;
        AFV_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        AFV_CUDA_CHECK(cudaStreamCreateWithPriority(&ex->aux.stream, cudaStreamNonBlocking, prio_greatest));
    }
    AFV_CUDA_CHECK(cudaEventCreateWithFlags(&ex->aux.ev_pyr, cudaEventDisableTiming));
    AFV_CUDA_CHECK(cudaEventCreateWithFlags(&ex->aux.ev_blur, cudaEventDisableTiming));

    features_per_level(nfeatures * 10, n_octaves, 1.2f, ex->q_orb);          // cv::ORB, maxFeatures = 10*nfeatures
    features_per_level(nfeatures, n_octaves, scale_factor, ex->q_ext);       // mnFeaturesPerLevel
    // computeSize (reference src/FeatureExtractor.cpp:132-142) with GetKeypointSize = powf(scaleFactor0, octave)
    const float maxSize0 = powf(1.2f, (float)(8 - 1.0)), maxSize = maxSize0, minSize = 1.0f;
    for (int l = 0; l < n_octaves; ++l) {
        float s = powf(scale_factor, (float)l);
        float sn = maxSize;
        if (maxSize > minSize) sn = 1.0f + (s - minSize) * (maxSize0 - 1.0f) / (maxSize - minSize);
        ex->size_norm[l] = sn;
        ex->ext_scale[l] = l == 0 ? 1.0f : ex->ext_scale[l - 1] * scale_factor;   // mvScaleFactor (:81)
    }
    int lw[AFV_MAX_LEVELS], lh[AFV_MAX_LEVELS]; float ls[AFV_MAX_LEVELS];
    level_geometry(max_w, max_h, n_octaves, lw, lh, ls);
    int rc = AFV_OK;
    const size_t B = (size_t)max_batch;
    for (int l = 0; l < n_octaves && rc == AFV_OK; ++l) {
        ex->max_stride[l] = ((lw[l] + 127) & ~127) + 128;
        ex->max_lh[l] = lh[l] + 2;
        const size_t lvl_bytes = (size_t)ex->max_stride[l] * ex->max_lh[l];
        ex->cand_cap[l] = lw[l] * lh[l] / 8 > 1024 ? lw[l] * lh[l] / 8 : 1024;
        ex->det_cap[l] = ex->q_orb[l] + 64;
        ex->keep_cap[l] = ex->q_ext[l] + 8;
        if ((rc = dev_alloc(ex, &ex->pyr[l], lvl_bytes * B))) break;
        if ((rc = dev_alloc(ex, &ex->blur[l], lvl_bytes * B))) break;
        if ((rc = dev_alloc(ex, &ex->cand[l], (size_t)ex->cand_cap[l] * B))) break;
        if ((rc = dev_alloc(ex, &ex->cand_resp[l], (size_t)ex->cand_cap[l] * B))) break;
        if ((rc = dev_alloc(ex, &ex->det[l], (size_t)ex->det_cap[l] * B))) break;
        if ((rc = dev_alloc(ex, &ex->keep[l], (size_t)ex->keep_cap[l] * B))) break;
        if ((rc = dev_alloc(ex, &ex->tab[l], (size_t)(lw[l] + lh[l]) + 16))) break;
    }
    const int ocap = nfeatures + 3 * n_octaves;
    if (rc == AFV_OK) {
        ex->tiles_cap = 0;
        for (int l = 0; l < n_octaves; ++l) ex->tiles_cap += ((lw[l] + AFV_TILE_W - 1) / AFV_TILE_W + 1) * ((lh[l] + AFV_TILE_H - 1) / AFV_TILE_H + 1);
        rc = dev_alloc(ex, &ex->tiles, (size_t)ex->tiles_cap);
    }
    if (rc == AFV_OK) rc = dev_alloc(ex, &ex->counts, 4 * AFV_MAX_LEVELS * B);
    if (rc == AFV_OK) rc = dev_alloc(ex, &ex->status, B);
    if (rc == AFV_OK) rc = dev_alloc(ex, &ex->o_kps, (size_t)ocap * B);
    if (rc == AFV_OK) rc = dev_alloc(ex, &ex->o_desc, (size_t)ocap * 32 * B);
    if (rc == AFV_OK) rc = dev_alloc(ex, &ex->o_size, (size_t)ocap * B);
    if (rc == AFV_OK) rc = dev_alloc(ex, &ex->o_n, B);
    if (rc == AFV_OK) {
        cudaError_t e2 = cudaMallocHost((void**)&ex->h_status, sizeof(int) * B);
        if (e2 == cudaSuccess) e2 = cudaMallocHost((void**)&ex->h_counts, sizeof(int) * 4 * AFV_MAX_LEVELS * B);
        if (e2 != cudaSuccess) { afv_set_error("cudaMallocHost failed: %s", cudaGetErrorString(e2)); rc = AFV_ERR_CUDA; }
    }
    if (rc == AFV_OK) {
        int mdet = 0, mkeep = 0;
        for (int l = 0; l < n_octaves; ++l) { if (ex->det_cap[l] > mdet) mdet = ex->det_cap[l]; if (ex->keep_cap[l] > mkeep) mkeep = ex->keep_cap[l]; }
        rc = afv_orb_configure(mdet, mkeep, &ex->oct_mcap, &ex->oct_ncap);
    }
    if (rc != AFV_OK) { afv_extractor_destroy(ex); return rc; }
    *out = ex;
    return AFV_OK;
}

extern "C" void afv_extractor_destroy(afv_extractor* ex) {
    if (!ex) return;
    cudaSetDevice(ex->device);
    if (ex->stream) { cudaStreamSynchronize(ex->stream); cudaStreamDestroy(ex->stream); }
    if (ex->aux.stream) { cudaStreamSynchronize(ex->aux.stream); cudaStreamDestroy(ex->aux.stream); }
    if (ex->ev_order) cudaEventDestroy(ex->ev_order);
    if (ex->aux.ev_pyr) cudaEventDestroy(ex->aux.ev_pyr);
    if (ex->aux.ev_blur) cudaEventDestroy(ex->aux.ev_blur);
    for (void* p : ex->allocs) cudaFree(p);
    if (ex->sift) afv_sift_destroy(ex->sift);
    if (ex->akaze) afv_akaze_destroy(ex->akaze);
    if (ex->brisk) afv_brisk_destroy(ex->brisk);
    if (ex->os2) afv_os2_destroy(ex->os2);
    if (ex->h_status) cudaFreeHost(ex->h_status);
    if (ex->h_counts) cudaFreeHost(ex->h_counts);
    delete ex;
}

extern "C" int afv_extractor_output_cap(const afv_extractor* ex) { return ex ? ex->nfeatures + 3 * ex->nlevels : AFV_ERR_INVALID; }

extern "C" int afv_extractor_levels(const afv_extractor* ex, float* scale_factors, int* features_per_level_out) {
    if (!ex) return AFV_ERR_INVALID;
    for (int l = 0; l < ex->nlevels; ++l) {
        if (scale_factors) scale_factors[l] = ex->ext_scale[l];
        if (features_per_level_out) features_per_level_out[l] = ex->q_ext[l];
    }
    return ex->nlevels;
}

static int run_device(afv_extractor* ex, const uint8_t* d_gray, int B, int w, int h, int stride, long frame_stride,
                      afv_keypoint* d_kps, void* d_desc, float* d_kpsize, int cap, int* d_n_out, cudaStream_t st,
                      bool gray_is_staged) {
    if (B < 1 || B > ex->max_batch) { afv_set_error("batch %d outside 1..%d", B, ex->max_batch); return AFV_ERR_INVALID; }
    if (cap < afv_extractor_output_cap(ex)) { afv_set_error("cap %d < required %d", cap, afv_extractor_output_cap(ex)); return AFV_ERR_INVALID; }
    if (ex->sift || ex->akaze || ex->brisk || ex->os2) {
        const int src = ex->os2 ? afv_os2_run(ex->os2, d_gray, B, w, h, stride, frame_stride, d_kps, (uint8_t*)d_desc, d_kpsize, cap, d_n_out, st)
                      : ex->sift ? afv_sift_run(ex->sift, d_gray, B, w, h, stride, frame_stride, d_kps, (float*)d_desc, d_kpsize, cap, d_n_out, st)
                      : ex->brisk ? afv_brisk_run(ex->brisk, d_gray, B, w, h, stride, frame_stride, d_kps, (uint8_t*)d_desc, d_kpsize, cap, d_n_out, st)
                                 : afv_akaze_run(ex->akaze, d_gray, B, w, h, stride, frame_stride, d_kps, (uint8_t*)d_desc, d_kpsize, cap, d_n_out, st);
        if (src) return src;
        ex->last_B = B; ex->last_stream = st;
        return AFV_OK;
    }
    int rc = configure_geometry(ex, w, h);
    if (rc) return rc;
    AfvParams P = ex->P;
    P.B = B; P.out_cap = cap;
    if (!gray_is_staged) {
        const bool alias_ok = (((uintptr_t)d_gray & 15) == 0) && (stride % 16 == 0) && (frame_stride % 16 == 0) && stride >= w;
        if (alias_ok) { P.lv[0].img = d_gray; P.lv[0].img_stride = stride; P.lv[0].img_fstride = frame_stride; }
        else {
            for (int b = 0; b < B; ++b)
                AFV_CUDA_CHECK(cudaMemcpy2DAsync(ex->pyr[0] + (long long)b * P.lv[0].fstride, P.lv[0].stride,
                                                 d_gray + (long long)b * frame_stride, stride, w, h, cudaMemcpyDeviceToDevice, st));
        }
    }
    { const int lrc = afv_launch_extract(P, d_kps, (uint8_t*)d_desc, d_kpsize, d_n_out, st, ex->aux); if (lrc) return lrc; }
    AFV_CUDA_CHECK(cudaGetLastError());
    ex->last_B = B; ex->last_stream = st;
    return AFV_OK;
}

static int check_status(afv_extractor* ex, int B, cudaStream_t st) {
    if (ex->sift) return afv_sift_status(ex->sift, B, st);
    if (ex->akaze) return afv_akaze_status(ex->akaze, B, st);
    if (ex->brisk) return afv_brisk_status(ex->brisk, B, st);
    if (ex->os2) return afv_os2_status(ex->os2, B, st);
    AFV_CUDA_CHECK(cudaMemcpyAsync(ex->h_status, ex->status, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    AFV_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b)
        if (ex->h_status[b]) {
            afv_set_error("capacity exceeded in frame %d (flags 0x%x: 1 FAST candidates, 2 detect list, 4 caller cap, 8 octree)", b, ex->h_status[b]);
            return AFV_ERR_CAPACITY;
        }
    return AFV_OK;
}

extern "C" int afv_extract_batch_device(afv_extractor* ex, const uint8_t* d_gray, int B, int w, int h, int stride,
                                        long frame_stride, afv_keypoint* d_kps, void* d_desc, float* d_kpsize,
                                        int cap, int* d_n_out, void* cuda_stream) {
    if (!ex || !d_gray || !d_kps || !d_desc || !d_n_out) { afv_set_error("NULL argument"); return AFV_ERR_INVALID; }
    AFV_CUDA_CHECK(cudaSetDevice(ex->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ex->stream;
    if (!cuda_stream) {
        // the private stream is cudaStreamNonBlocking: order it after whatever the legacy default stream has queued (the
        // usual producer of d_gray and of freshly zeroed output buffers) so NULL never races with the caller's work
        AFV_CUDA_CHECK(cudaEventRecord(ex->ev_order, cudaStreamLegacy));
        AFV_CUDA_CHECK(cudaStreamWaitEvent(ex->stream, ex->ev_order, 0));
    }
    int rc = run_device(ex, d_gray, B, w, h, stride, frame_stride, d_kps, d_desc, d_kpsize, cap, d_n_out, st, false);
    if (rc) return rc;
    if (!cuda_stream) return check_status(ex, B, st);
    return AFV_OK;
}

extern "C" int afv_extractor_status(afv_extractor* ex) {
    if (!ex) return AFV_ERR_INVALID;
    if (!ex->last_B) return AFV_OK;
    AFV_CUDA_CHECK(cudaSetDevice(ex->device));
    return check_status(ex, ex->last_B, ex->last_stream);
}

extern "C" int afv_extract_batch(afv_extractor* ex, const uint8_t* gray, int B, int w, int h, int stride,
                                 long frame_stride, afv_keypoint* kps, void* desc, float* kpsize, int cap, int* n_out) {
    if (!ex || !gray || !kps || !desc || !n_out) { afv_set_error("NULL argument"); return AFV_ERR_INVALID; }
    AFV_CUDA_CHECK(cudaSetDevice(ex->device));
    const int ocap = afv_extractor_output_cap(ex);
    if (cap < ocap) { afv_set_error("cap %d < required %d", cap, ocap); return AFV_ERR_INVALID; }
    if (stride < w) { afv_set_error("stride < w"); return AFV_ERR_INVALID; }
    cudaStream_t st = ex->stream;
    if (ex->sift || ex->akaze || ex->brisk || ex->os2) {
        if (w > ex->max_w || h > ex->max_h) { afv_set_error("frame %dx%d larger than the extractor's configured maximum", w, h); return AFV_ERR_INVALID; }
        uint8_t* stage = ex->os2 ? afv_os2_stage(ex->os2) : ex->sift ? afv_sift_stage(ex->sift) : ex->brisk ? afv_brisk_stage(ex->brisk) : afv_akaze_stage(ex->akaze);
        const size_t DB = (size_t)ex->desc_bytes;
        for (int b0 = 0; b0 < B; b0 += ex->max_batch) {
            const int nb = B - b0 < ex->max_batch ? B - b0 : ex->max_batch;
            if (frame_stride == (long)stride * h)
                AFV_CUDA_CHECK(cudaMemcpy2DAsync(stage, w, gray + (long long)b0 * frame_stride, stride, w, (size_t)h * nb, cudaMemcpyHostToDevice, st));
            else
                for (int b = 0; b < nb; ++b)
                    AFV_CUDA_CHECK(cudaMemcpy2DAsync(stage + (size_t)b * w * h, w, gray + (long long)(b0 + b) * frame_stride, stride, w, h, cudaMemcpyHostToDevice, st));
            int rc = run_device(ex, stage, nb, w, h, w, (long)w * h, ex->o_kps, ex->o_desc, ex->o_size, ocap, ex->o_n, st, true);
            if (rc) return rc;
            AFV_CUDA_CHECK(cudaMemcpy2DAsync(kps + (size_t)b0 * cap, (size_t)cap * sizeof(afv_keypoint), ex->o_kps, (size_t)ocap * sizeof(afv_keypoint),
                                             (size_t)ocap * sizeof(afv_keypoint), nb, cudaMemcpyDeviceToHost, st));
            AFV_CUDA_CHECK(cudaMemcpy2DAsync((uint8_t*)desc + (size_t)b0 * cap * DB, (size_t)cap * DB, ex->o_desc, (size_t)ocap * DB,
                                             (size_t)ocap * DB, nb, cudaMemcpyDeviceToHost, st));
            if (kpsize)
                AFV_CUDA_CHECK(cudaMemcpy2DAsync(kpsize + (size_t)b0 * cap, (size_t)cap * sizeof(float), ex->o_size, (size_t)ocap * sizeof(float),
                                                 (size_t)ocap * sizeof(float), nb, cudaMemcpyDeviceToHost, st));
            AFV_CUDA_CHECK(cudaMemcpyAsync(n_out + b0, ex->o_n, sizeof(int) * nb, cudaMemcpyDeviceToHost, st));
            rc = check_status(ex, nb, st);
            if (rc) return rc;
        }
        return AFV_OK;
    }
    for (int b0 = 0; b0 < B; b0 += ex->max_batch) {
        const int nb = B - b0 < ex->max_batch ? B - b0 : ex->max_batch;
        int rc = configure_geometry(ex, w, h);
        if (rc) return rc;
        const AfvLevel& L0 = ex->P.lv[0];
        if (stride == w && frame_stride == (long)w * h && L0.stride == w) {
            AFV_CUDA_CHECK(cudaMemcpyAsync(ex->pyr[0], gray + (long long)b0 * frame_stride, (size_t)nb * w * h, cudaMemcpyHostToDevice, st));
        } else {
            for (int b = 0; b < nb; ++b)
                AFV_CUDA_CHECK(cudaMemcpy2DAsync(ex->pyr[0] + (long long)b * L0.fstride, L0.stride,
                                                 gray + (long long)(b0 + b) * frame_stride, stride, w, h, cudaMemcpyHostToDevice, st));
        }
        rc = run_device(ex, ex->pyr[0], nb, w, h, L0.stride, (long)L0.fstride, ex->o_kps, ex->o_desc, ex->o_size, ocap, ex->o_n, st, true);
        if (rc) return rc;
        // D2H: rows are ocap wide on the device, cap wide at the caller
        AFV_CUDA_CHECK(cudaMemcpy2DAsync(kps + (size_t)b0 * cap, (size_t)cap * sizeof(afv_keypoint), ex->o_kps, (size_t)ocap * sizeof(afv_keypoint),
                                         (size_t)ocap * sizeof(afv_keypoint), nb, cudaMemcpyDeviceToHost, st));
        AFV_CUDA_CHECK(cudaMemcpy2DAsync((uint8_t*)desc + (size_t)b0 * cap * 32, (size_t)cap * 32, ex->o_desc, (size_t)ocap * 32,
                                         (size_t)ocap * 32, nb, cudaMemcpyDeviceToHost, st));
        if (kpsize)
            AFV_CUDA_CHECK(cudaMemcpy2DAsync(kpsize + (size_t)b0 * cap, (size_t)cap * sizeof(float), ex->o_size, (size_t)ocap * sizeof(float),
                                             (size_t)ocap * sizeof(float), nb, cudaMemcpyDeviceToHost, st));
        AFV_CUDA_CHECK(cudaMemcpyAsync(n_out + b0, ex->o_n, sizeof(int) * nb, cudaMemcpyDeviceToHost, st));
        rc = check_status(ex, nb, st);
        if (rc) return rc;
    }
    return AFV_OK;
}

extern "C" int afv_extract(afv_extractor* ex, const uint8_t* gray, int w, int h, int stride,
                           afv_keypoint* kps, void* desc, float* kpsize, int cap, int* n_out) {
    return afv_extract_batch(ex, gray, 1, w, h, stride, (long)stride * h, kps, desc, kpsize, cap, n_out);
}

extern "C" int afv_debug_read(afv_extractor* ex, int what, int frame, int level, void* out, long cap_bytes, long* n_bytes) {
    if (ex && (ex->sift || ex->akaze || ex->brisk || ex->os2)) {
        if (!out || !n_bytes || frame < 0 || frame >= ex->last_B) { afv_set_error("afv_debug_read: bad argument"); return AFV_ERR_INVALID; }
        AFV_CUDA_CHECK(cudaSetDevice(ex->device));
        AFV_CUDA_CHECK(cudaStreamSynchronize(ex->last_stream));
        return ex->os2 ? afv_os2_debug_read(ex->os2, what, frame, level, out, cap_bytes, n_bytes)
             : ex->sift ? afv_sift_debug_read(ex->sift, what, frame, level, out, cap_bytes, n_bytes)
             : ex->brisk ? afv_brisk_debug_read(ex->brisk, what, frame, level, out, cap_bytes, n_bytes)
                        : afv_akaze_debug_read(ex->akaze, what, frame, level, out, cap_bytes, n_bytes);
    }
    if (!ex || !out || !n_bytes || level < 0 || level >= ex->nlevels || frame < 0 || frame >= ex->last_B) {
        afv_set_error("afv_debug_read: bad argument"); return AFV_ERR_INVALID;
    }
    AFV_CUDA_CHECK(cudaSetDevice(ex->device));
    AFV_CUDA_CHECK(cudaStreamSynchronize(ex->last_stream));
    const AfvLevel& L = ex->P.lv[level];
    AFV_CUDA_CHECK(cudaMemcpy(ex->h_counts, ex->counts, sizeof(int) * 4 * AFV_MAX_LEVELS * ex->last_B, cudaMemcpyDeviceToHost));
    if (what == 0 || what == 1) {
        long need = (long)L.w * L.h;
        if (cap_bytes < need) { afv_set_error("buffer too small"); return AFV_ERR_INVALID; }
        if (what == 0 && level == 0) { afv_set_error("level 0 of the un-blurred pyramid is the caller's input"); return AFV_ERR_INVALID; }
        const uint8_t* src = (what == 0 ? ex->pyr[level] : ex->blur[level]) + (long long)frame * L.fstride;
        AFV_CUDA_CHECK(cudaMemcpy2D(out, L.w, src, L.stride, L.w, L.h, cudaMemcpyDeviceToHost));
        *n_bytes = need;
        return AFV_OK;
    }
    int which = what == 2 ? AFV_CNT_CAND : what == 3 ? AFV_CNT_DET : what == 4 ? AFV_CNT_KEEP : -1;
    if (which < 0) { afv_set_error("unknown tap %d", what); return AFV_ERR_INVALID; }
    int n = ex->h_counts[afv_cnt_idx(frame, which, level)];
    const int capn = what == 2 ? L.cand_cap : what == 3 ? L.det_cap : L.keep_cap;
    if (n > capn) n = capn;
    const long esz = what == 2 ? 4 : 8;
    if (cap_bytes < n * esz) { afv_set_error("buffer too small"); return AFV_ERR_INVALID; }
    const void* src = what == 2 ? (const void*)(L.cand + (long long)frame * L.cand_cap)
                    : what == 3 ? (const void*)(L.det + (long long)frame * L.det_cap)
                                : (const void*)(L.keep + (long long)frame * L.keep_cap);
    AFV_CUDA_CHECK(cudaMemcpy(out, src, n * esz, cudaMemcpyDeviceToHost));
    *n_bytes = n * esz;
    return AFV_OK;
}

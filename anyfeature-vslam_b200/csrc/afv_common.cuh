// afv_common.cuh -- shared declarations of the B200 feature front end (device layout + launch plumbing).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/afv.h"

#define AFV_MAX_LEVELS 16
#define AFV_GRID_COLS 64      // reference include/Frame.h:41
#define AFV_GRID_ROWS 48      // reference include/Frame.h:40
#define AFV_HISTO_LENGTH 30   // reference src/FeatureMatcher.cc:64

// Per-level geometry + arena pointers.  Arenas are level-major: [level][frame][row][stride] so a kernel over
// (tile, frame) of one level is a plain 3-D problem and rows are 128-byte aligned for vector / TMA access.
struct AfvLevel {
    int w, h;                 // level size (cv::ORB geometry: cvRound(dim / 1.2^l))
    int stride;               // bytes per row in pyr / blur arenas (multiple of 128)
    long long fstride;        // bytes per frame in pyr / blur arenas
    const uint8_t* img;       // un-blurred level, frame 0 (level 0 may alias the caller's input)
    int img_stride; long long img_fstride;
    uint8_t* blur;            // blurred level, frame 0
    float scale;              // (float)pow(1.2f as double, l)
    float inv_scale;          // 1.f / scale
    float kp_size;            // 31 * scale  (cv::KeyPoint::size)
    float size_norm;          // computeSize() value for this octave (reference src/FeatureExtractor.cpp:132-142)
    int q_orb;                // cv::ORB's own per-level quota from maxFeatures = 10*nfeatures
    int q_ext;                // extractor quota mnFeaturesPerLevel (octree N)
    int cand_cap;             // FAST+NMS candidate capacity per frame
    uint32_t* cand;           // [frame][cand_cap] packed x | y<<12 | score<<24
    float* cand_resp;         // [frame][cand_cap] Harris response scratch
    int det_cap;              // capacity of the detect list (q_orb + slack for exact ties)
    uint2* det;               // [frame][det_cap] {packed xy, response bits}
    int keep_cap;             // q_ext + 3 (+ slack)
    uint2* keep;              // [frame][keep_cap] octree output in node-list order
    // resize tables for building this level from level l-1 (INTER_LINEAR_EXACT 8.8 fixed point)
    const uint32_t* xtab;     // [w rounded up to 4] source offset << 16 | c1 (8.8 fixed point, c0 = 256 - c1)
    const uint32_t* ytab;     // [h]
};

#define AFV_TILE_W 128
#define AFV_TILE_H 32
struct AfvTile { short level, pad; short x0, y0; };      // one stencil tile of one pyramid level

struct AfvParams {
    int nlevels, B;
    const AfvTile* tiles; int ntiles;   // AFV_TILE_W x AFV_TILE_H tiles of all levels (k_fast / k_blur grids)
    int W, H;                 // level-0 size
    int fast_th;
    int n_ini; float hX;      // octree roots (reference src/ORBextractor.cc:243-245)
    int out_cap;              // per-frame output capacity (caller's cap)
    int oct_mcap, oct_ncap;   // k_octree capacities of THIS extractor (keys / nodes held in shared memory)
    int* counts;              // [frame][4][AFV_MAX_LEVELS]: 0 = #candidates, 1 = #det, 2 = #kept, 3 = spare
    int* status;              // [frame] bit flags (AFV_ST_*)
    AfvLevel lv[AFV_MAX_LEVELS];
};

#define AFV_ST_CAND_OVERFLOW 1
#define AFV_ST_DET_OVERFLOW  2
#define AFV_ST_OUT_OVERFLOW  4
#define AFV_ST_OCTREE_OVERFLOW 8

#define AFV_CNT_CAND 0
#define AFV_CNT_DET  1
#define AFV_CNT_KEEP 2
__host__ __device__ inline int afv_cnt_idx(int frame, int which, int level) {
    return (frame * 4 + which) * AFV_MAX_LEVELS + level;
}

// launch wrappers (afv_orb.cu)
struct AfvAux { cudaStream_t stream; cudaEvent_t ev_pyr, ev_blur; };   // side stream for kernels that only need the pyramid
int afv_launch_extract(const AfvParams& P, afv_keypoint* d_kps, uint8_t* d_desc, float* d_kpsize,
                       int* d_n_out, cudaStream_t st, const AfvAux& aux);
size_t afv_octree_smem_bytes(int mcap, int ncap);
int afv_orb_configure(int max_det_cap, int max_keep_cap, int* mcap_out, int* ncap_out);   // raises the smem attribute; returns 0 / error

// matcher launches (afv_match.cu) are called directly from afv_capi.cu through the C ABI.

extern std::atomic<long long> g_afv_launches;     // kernels launched by this library (host threads of a SLAM system call in concurrently)
// optional per-kernel CUDA-event timing (bench.py's roofline leg): AFV_PROF_BEGIN/END bracket one launch
int  afv_prof_begin(const char* name, cudaStream_t st);       // returns the record id (-1 when profiling is off)
void afv_prof_end(int id, cudaStream_t st);
bool afv_prof_is_on();
struct AfvProfScope { cudaStream_t st; int id; AfvProfScope(const char* n, cudaStream_t s) : st(s), id(afv_prof_begin(n, s)) {} ~AfvProfScope() { afv_prof_end(id, st); } };
void afv_set_error(const char* fmt, ...);
#define AFV_CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    afv_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); return AFV_ERR_CUDA; } } while (0)

// standalone debug harness: k_fast clone with smem dumps vs host emulation
#include "../../anyfeature-vslam_b200/csrc/afv_orb.cu"
#include <vector>
#include <cstring>
long long g_afv_launches = 0;
void afv_set_error(const char*, ...) {}
__global__ void __launch_bounds__(256) k_fast_dbg(const uint8_t* img, int w, int h, int stride, int t, int x0, int y0,
        uint8_t* o_pix, uint8_t* o_score, uint16_t* o_clist, int* o_nc, int* o_d) {
    __shared__ __align__(16) uint8_t pix[FT_PH][FT_PW];
    __shared__ uint8_t score[FT_RH][FT_SW];
    __shared__ uint16_t clist[FT_RW * FT_RH];
    __shared__ int ncorner;
    const int tid = threadIdx.x;
    if (tid == 0) ncorner = 0;
    for (int i = tid; i < FT_PH * (FT_PW / 4); i += 256) {
        const int r = i / (FT_PW / 4), c4 = i % (FT_PW / 4);
        const int gy = y0 - 4 + r, gx = x0 - 4 + c4 * 4;
        uint32_t v = 0;
        if (gy >= 0 && gy < h && gx >= 0 && gx < stride) v = *reinterpret_cast<const uint32_t*>(img + (long long)gy * stride + gx);
        *reinterpret_cast<uint32_t*>(&pix[r][c4 * 4]) = v;
    }
    for (int i = tid; i < FT_RH * FT_SW / 4; i += 256) reinterpret_cast<uint32_t*>(&score[0][0])[i] = 0;
    __syncthreads();
    for (int i = tid; i < FT_RW * FT_RH; i += 256) {
        const int r = i / FT_RW, c = i % FT_RW;
        const int gx = x0 - 1 + c, gy = y0 - 1 + r;
        if (gx < 3 || gy < 3 || gx >= w - 3 || gy >= h - 3) continue;
        const uint8_t* p = &pix[r + 3][c + 3];
        const int v = p[0], hi = v + t, lo = v - t;
        const int p0 = p[3 * FT_PW], p8 = p[-3 * FT_PW];
        if (!((p0 > hi) | (p0 < lo) | (p8 > hi) | (p8 < lo))) continue;
        const int p4 = p[3], p12 = p[-3];
        if (!((p4 > hi) | (p4 < lo) | (p12 > hi) | (p12 < lo))) continue;
        uint32_t br = 0, dk = 0;
#define FMASK(k, dx, dy) { const int q = p[(dy) * FT_PW + (dx)]; br |= (uint32_t)(q > hi) << k; dk |= (uint32_t)(q < lo) << k; }
        CIRC16(FMASK)
#undef FMASK
        if (has_arc9(br) || has_arc9(dk)) clist[atomicAdd(&ncorner, 1)] = (uint16_t)i;
    }
    __syncthreads();
    const int nc = ncorner;
    for (int j = tid; j < nc; j += 256) {
        const int i = clist[j];
        const int r = i / FT_RW, c = i % FT_RW;
        const uint8_t* p = &pix[r + 3][c + 3];
        const int v = p[0];
        int d[16];
#define FDIFF(k, dx, dy) d[k] = v - (int)p[(dy) * FT_PW + (dx)];
        CIRC16(FDIFF)
#undef FDIFF
        for (int k = 0; k < 16; ++k) o_d[j * 16 + k] = d[k];
        int mn2[16], mx2[16], mn4[16], mx4[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { mn2[k] = min(d[k], d[(k + 1) & 15]); mx2[k] = max(d[k], d[(k + 1) & 15]); }
#pragma unroll
        for (int k = 0; k < 16; ++k) { mn4[k] = min(mn2[k], mn2[(k + 2) & 15]); mx4[k] = max(mx2[k], mx2[(k + 2) & 15]); }
        int best = -256;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int mn9 = min(min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
            const int mx9 = max(max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
            best = max(best, max(mn9, -mx9));
        }
        score[r][c] = (uint8_t)(best - 1);
    }
    __syncthreads();
    for (int i = tid; i < FT_PH * FT_PW; i += 256) o_pix[i] = (&pix[0][0])[i];
    for (int i = tid; i < FT_RH * FT_SW; i += 256) o_score[i] = (&score[0][0])[i];
    for (int i = tid; i < nc; i += 256) o_clist[i] = clist[i];
    if (tid == 0) *o_nc = nc;
}
int main() {
    int w = 256, h = 64; std::vector<uint8_t> img(w * h); srand(3);
    for (auto& v : img) v = rand() % 256;
    for (int it = 0; it < 2; ++it) { auto t = img; for (int y = 1; y < h - 1; ++y) for (int x = 1; x < w - 1; ++x) { int s = 0; for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) s += t[(y + dy) * w + x + dx]; img[y * w + x] = s / 9; } }
    for (int i = 0; i < 40; ++i) { int x0 = rand() % (w - 40), y0 = rand() % (h - 40), g = rand() % 256; for (int y = y0; y < y0 + 30; ++y) for (int x = x0; x < x0 + 30; ++x) img[y * w + x] = (img[y * w + x] + g) / 2; }
    uint8_t *d_img, *d_pix, *d_score; uint16_t* d_clist; int *d_nc, *d_d;
    cudaMalloc(&d_img, w * h); cudaMemcpy(d_img, img.data(), w * h, cudaMemcpyHostToDevice);
    cudaMalloc(&d_pix, FT_PH * FT_PW); cudaMalloc(&d_score, FT_RH * FT_SW); cudaMalloc(&d_clist, 2 * FT_RW * FT_RH); cudaMalloc(&d_nc, 4); cudaMalloc(&d_d, 4 * 16 * FT_RW * FT_RH);
    int x0 = 128, y0 = 16, t = 20;
    k_fast_dbg<<<1, 256>>>(d_img, w, h, w, t, x0, y0, d_pix, d_score, d_clist, d_nc, d_d);
    printf("launch: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    std::vector<uint8_t> g_pix(FT_PH * FT_PW), g_score(FT_RH * FT_SW); std::vector<uint16_t> g_clist(FT_RW * FT_RH); int g_nc; std::vector<int> g_d(16 * FT_RW * FT_RH);
    cudaMemcpy(g_pix.data(), d_pix, g_pix.size(), cudaMemcpyDeviceToHost); cudaMemcpy(g_score.data(), d_score, g_score.size(), cudaMemcpyDeviceToHost);
    cudaMemcpy(&g_nc, d_nc, 4, cudaMemcpyDeviceToHost); cudaMemcpy(g_clist.data(), d_clist, 2 * g_nc, cudaMemcpyDeviceToHost); cudaMemcpy(g_d.data(), d_d, 4 * 16 * g_nc, cudaMemcpyDeviceToHost);
    // host emulation
    static uint8_t pix[FT_PH][FT_PW]; static uint8_t score[FT_RH][FT_SW]; memset(score, 0, sizeof(score));
    for (int r = 0; r < FT_PH; ++r) for (int c = 0; c < FT_PW; ++c) { int gy = y0 - 4 + r, gx = x0 - 4 + c; pix[r][c] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? img[gy * w + gx] : 0; }
    int pixbad = 0; for (int i = 0; i < FT_PH * FT_PW; ++i) pixbad += g_pix[i] != (&pix[0][0])[i];
    printf("pix mismatches %d, gpu corners %d\n", pixbad, g_nc);
    int shown = 0, sbad = 0;
    for (int j = 0; j < g_nc; ++j) {
        int i = g_clist[j], r = i / FT_RW, c = i % FT_RW; const uint8_t* p = &pix[r + 3][c + 3]; int v = p[0]; int d[16];
#define FDIFF(k, dx, dy) d[k] = v - (int)p[(dy) * FT_PW + (dx)];
        CIRC16(FDIFF)
        int best = -256; for (int s = 0; s < 16; ++s) { int mn = d[s], mx = d[s]; for (int k = 1; k < 9; ++k) { int tt = d[(s + k) & 15]; mn = min(mn, tt); mx = max(mx, tt); } best = max(best, max(mn, -mx)); }
        int hs = best - 1; int gs = g_score[r * FT_SW + c];
        bool dbad = false; for (int k = 0; k < 16; ++k) dbad |= d[k] != g_d[j * 16 + k];
        if (hs != gs || dbad) { ++sbad; if (shown++ < 6) { printf("corner i=%d r=%d c=%d host score %d gpu %d dbad %d\n  host d:", i, r, c, hs, gs, (int)dbad); for (int k = 0; k < 16; ++k) printf(" %d", d[k]); printf("\n  gpu  d:"); for (int k = 0; k < 16; ++k) printf(" %d", g_d[j * 16 + k]); printf("\n"); } }
    }
    printf("score mismatches %d of %d\n", sbad, g_nc);
    return 0;
}

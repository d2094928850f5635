#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
__device__ __forceinline__ int v1(const int* d) {
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
    return best - 1;
}
__device__ __forceinline__ int v2(const int* d) {   // min-only on d and -d
    int A = -256;
#pragma unroll
    for (int sgn = 0; sgn < 2; ++sgn) {
        int e[16], m2[16], m4[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = sgn ? -d[k] : d[k];
#pragma unroll
        for (int k = 0; k < 16; ++k) m2[k] = min(e[k], e[(k + 1) & 15]);
#pragma unroll
        for (int k = 0; k < 16; ++k) m4[k] = min(m2[k], m2[(k + 2) & 15]);
#pragma unroll
        for (int k = 0; k < 16; ++k) A = max(A, min(min(m4[k], m4[(k + 4) & 15]), e[(k + 8) & 15]));
    }
    return A - 1;
}
__device__ __forceinline__ int pmin(int a, int b) { int r; asm("min.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ int pmax(int a, int b) { int r; asm("max.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ int v3(const int* d) {   // same as v1 with explicit PTX min/max
    int mn2[16], mx2[16], mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn2[k] = pmin(d[k], d[(k + 1) & 15]); mx2[k] = pmax(d[k], d[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn4[k] = pmin(mn2[k], mn2[(k + 2) & 15]); mx4[k] = pmax(mx2[k], mx2[(k + 2) & 15]); }
    int best = -256;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int mn9 = pmin(pmin(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
        const int mx9 = pmax(pmax(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
        best = pmax(best, pmax(mn9, -mx9));
    }
    return best - 1;
}
__device__ __forceinline__ int v4(const int* d) {   // packed 16-bit SIMD: lanes (d, -d)
    unsigned e[16], m2[16], m4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) e[k] = ((unsigned)(d[k] & 0xffff)) | ((unsigned)((-d[k]) & 0xffff) << 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) m2[k] = __vmins2(e[k], e[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m4[k] = __vmins2(m2[k], m2[(k + 2) & 15]);
    unsigned A = 0x80008000u;
#pragma unroll
    for (int k = 0; k < 16; ++k) A = __vmaxs2(A, __vmins2(__vmins2(m4[k], m4[(k + 4) & 15]), e[(k + 8) & 15]));
    int lo = (short)(A & 0xffff), hi = (short)(A >> 16);
    return max(lo, hi) - 1;
}
__device__ __forceinline__ int v5(const int* d) {   // straightforward double loop
    int best = -256;
    for (int s = 0; s < 16; ++s) { int mn = d[s], mx = d[s]; for (int k = 1; k < 9; ++k) { int t = d[(s + k) & 15]; mn = min(mn, t); mx = max(mx, t); } best = max(best, max(mn, -mx)); }
    return best - 1;
}
__global__ void k(const int* d, int n, int* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    int dd[16];
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) dd[k2] = d[i * 16 + k2];
    out[i * 5 + 0] = v1(dd); out[i * 5 + 1] = v2(dd); out[i * 5 + 2] = v3(dd); out[i * 5 + 3] = v4(dd); out[i * 5 + 4] = v5(dd);
}
int main() {
    int n = 1 << 16; std::vector<int> d(n * 16); srand(1);
    for (int i = 0; i < n; ++i) { int mode = rand() % 3; for (int k2 = 0; k2 < 16; ++k2) d[i * 16 + k2] = mode == 0 ? rand() % 511 - 255 : (rand() % 4 ? -(rand() % 80) : rand() % 3); }
    int *dd, *dout; cudaMalloc(&dd, n * 64); cudaMalloc(&dout, n * 20); cudaMemcpy(dd, d.data(), n * 64, cudaMemcpyHostToDevice);
    k<<<n / 128, 128>>>(dd, n, dout); printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    std::vector<int> o(n * 5); cudaMemcpy(o.data(), dout, n * 20, cudaMemcpyDeviceToHost);
    int bad[5] = {0};
    for (int i = 0; i < n; ++i) { int* x = &d[i * 16]; int best = -256; for (int s = 0; s < 16; ++s) { int mn = x[s], mx = x[s]; for (int k2 = 1; k2 < 9; ++k2) { int t = x[(s + k2) & 15]; mn = std::min(mn, t); mx = std::max(mx, t); } best = std::max(best, std::max(mn, -mx)); }
        for (int v = 0; v < 5; ++v) bad[v] += o[i * 5 + v] != best - 1; }
    printf("bad: v1 %d v2 %d v3(ptx) %d v4(simd2) %d v5(loop) %d of %d\n", bad[0], bad[1], bad[2], bad[3], bad[4], n);
}

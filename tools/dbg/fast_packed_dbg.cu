// fast_packed_dbg.cu -- standalone A/B harness for the k_fast stage-1 idea of DESIGN.md section 10 (NOT part of the library):
// the sign-consistent compass test on TWO pixels per 32-bit register with VIMNMX.S16x2, bytes widened to 0x0100 | p by PRMT
// and per-half predicates read from bit 15 (identities checked exhaustively on the CPU by tools/check_packed_fast.py).
// Build + run on a B200:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fast_packed_dbg fast_packed_dbg.cu && ./fast_packed_dbg
// Prints whether the packed masks equal the scalar ones on a random image and the time of both kernels.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define W 2048
#define H 2048
#define ROWS 16                 // rows walked per thread

// scalar reference: one thread = one column, ROWS rows; bit j of out = pass(row y0 + j)
__global__ void k_scalar(const uint8_t* __restrict__ img, uint32_t* __restrict__ out, int t) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x + 3;
    const int y0 = blockIdx.y * ROWS + 3;
    if (x >= W - 3) return;
    uint32_t m = 0;
    for (int j = 0; j < ROWS; ++j) {
        const int y = y0 + j;
        if (y >= H - 3) break;
        const int v = img[y * W + x], hi = v + t, lo = v - t;
        const int dn = img[(y + 3) * W + x], up = img[(y - 3) * W + x], rt = img[y * W + x + 3], lf = img[y * W + x - 3];
        const bool pass = (((dn > hi) | (up > hi)) & ((rt > hi) | (lf > hi))) | (((dn < lo) | (up < lo)) & ((rt < lo) | (lf < lo)));
        m |= (uint32_t)pass << j;
    }
    out[blockIdx.y * W + x] = m;
}

// packed: one thread = columns x and x + 1 (x even + 3 offset handled by the caller's indexing), ROWS rows
__device__ __forceinline__ uint32_t pack2(uint32_t a, uint32_t b) { return (a | 0x100u) | ((b | 0x100u) << 16); }
__global__ void k_packed(const uint8_t* __restrict__ img, uint32_t* __restrict__ out, int t) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 2 + 3;
    const int y0 = blockIdx.y * ROWS + 3;
    if (x + 1 >= W - 3) return;
    const uint32_t tt = (uint32_t)(t + 1) * 0x10001u;
    const uint32_t C = 0x80008000u - tt;
    uint32_t ma = 0, mb = 0;
    for (int j = 0; j < ROWS; ++j) {
        const int y = y0 + j;
        if (y >= H - 3) break;
        const uint8_t* r = img + y * W + x;
        const uint32_t v2 = pack2(r[0], r[1]);
        const uint32_t dn = pack2(r[3 * W], r[3 * W + 1]), up = pack2(r[-3 * W], r[-3 * W + 1]);
        const uint32_t rt = pack2(r[3], r[4]), lf = pack2(r[-3], r[-2]);
        const uint32_t mx = __vmins2(__vmaxs2(dn, up), __vmaxs2(rt, lf));       // both compass pairs have a pixel > hi  <=>  mx > hi
        const uint32_t mn = __vmaxs2(__vmins2(dn, up), __vmins2(rt, lf));       // both compass pairs have a pixel < lo  <=>  mn < lo
        const uint32_t e = mx + (C - v2);                                        // bit 15 / 31: mx > v + t
        const uint32_t f = (v2 + C) - mn;                                        // bit 15 / 31: mn < v - t
        const uint32_t p = (e | f) & 0x80008000u;
        ma |= ((p >> 15) & 1u) << j;
        mb |= (p >> 31) << j;
    }
    out[blockIdx.y * W + x] = ma;
    out[blockIdx.y * W + x + 1] = mb;
}

int main() {
    uint8_t* h = (uint8_t*)malloc((size_t)W * H);
    srand(1);
    for (size_t i = 0; i < (size_t)W * H; ++i) h[i] = (uint8_t)((rand() % 64) + ((i / 7) % 3) * 60 + (rand() % 100 < 3 ? 120 : 0));
    uint8_t* d; uint32_t *oa, *ob;
    const int nby = (H - 6 + ROWS - 1) / ROWS;
    cudaMalloc(&d, (size_t)W * H); cudaMalloc(&oa, (size_t)nby * W * 4); cudaMalloc(&ob, (size_t)nby * W * 4);
    cudaMemcpy(d, h, (size_t)W * H, cudaMemcpyHostToDevice);
    cudaMemset(oa, 0, (size_t)nby * W * 4); cudaMemset(ob, 0, (size_t)nby * W * 4);
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    const int t = 20;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_scalar<<<dim3((W - 6 + 255) / 256, nby), 256>>>(d, oa, t);
        cudaEventRecord(e1);
        k_packed<<<dim3(((W - 6) / 2 + 255) / 256, nby), 256>>>(d, ob, t);
        cudaEventRecord(e2);
        cudaEventSynchronize(e2);
    }
    float ms_a, ms_b; cudaEventElapsedTime(&ms_a, e0, e1); cudaEventElapsedTime(&ms_b, e1, e2);
    uint32_t* ha = (uint32_t*)malloc((size_t)nby * W * 4); uint32_t* hb = (uint32_t*)malloc((size_t)nby * W * 4);
    cudaMemcpy(ha, oa, (size_t)nby * W * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb, ob, (size_t)nby * W * 4, cudaMemcpyDeviceToHost);
    long bad = 0, pass = 0;
    for (int by = 0; by < nby; ++by) for (int x = 3; x < W - 4; ++x) { bad += ha[by * W + x] != hb[by * W + x]; pass += __builtin_popcount(ha[by * W + x]); }
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    printf("stage-1 pass rate %.3f, mismatching words %ld, scalar %.3f ms, packed %.3f ms\n", (double)pass / ((double)(W - 7) * (H - 6)), bad, ms_a, ms_b);
    return bad != 0;
}

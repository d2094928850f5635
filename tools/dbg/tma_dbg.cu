// standalone check of the TMA 3-D u8 box load used by k_fast (coordinates: negative / unaligned / OOB)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#define BW 144
#define BH 40
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, uint8_t* out, int* status) {
    __shared__ __align__(128) uint8_t buf[BH][BW];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(BW * BH) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(smem_u32(&buf[0][0])), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    uint32_t ok = 0; int spin = 0;
    while (!ok && spin < 100000) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        ++spin;
    }
    if (threadIdx.x == 0) { status[0] = ok; status[1] = spin; }
    __syncthreads();
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = (&buf[0][0])[i];
}
int main() {
    const int w = 533, h = 400, B = 3, stride = 640; const long fstride = (long)stride * h;
    std::vector<uint8_t> img((size_t)fstride * B);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t* d; cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap tm; cuuint64_t gd[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B}; cuuint64_t gs[2] = {(cuuint64_t)stride, (cuuint64_t)fstride};
    cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
    CUresult r = ((enc_t)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d fn=%p\n", (int)r, fn);
    uint8_t* dout; int* dst; cudaMalloc(&dout, BW * BH); cudaMalloc(&dst, 8);
    const int tests[][3] = {{0, 0, 0}, {128, 32, 1}, {120, 28, 1}, {-8, -4, 0}, {504, 380, 2}, {3, 5, 1}};
    for (auto& t : tests) {
        cudaMemset(dout, 0xEE, BW * BH); cudaMemset(dst, 0, 8);
        k<<<1, 128>>>(tm, t[0], t[1], t[2], dout, dst);
        cudaError_t e = cudaDeviceSynchronize();
        int st[2]; std::vector<uint8_t> o(BW * BH); cudaMemcpy(st, dst, 8, cudaMemcpyDeviceToHost); cudaMemcpy(o.data(), dout, BW * BH, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int r2 = 0; r2 < BH; ++r2) for (int c = 0; c < BW; ++c) {
            const int x = t[0] + c, y = t[1] + r2; uint8_t exp = 0;
            if (x >= 0 && x < w && y >= 0 && y < h) exp = img[(size_t)t[2] * fstride + (size_t)y * stride + x];
            bad += o[r2 * BW + c] != exp;
        }
        printf("coord (%d,%d,%d): %s done=%d spins=%d mismatches=%ld\n", t[0], t[1], t[2], cudaGetErrorString(e), st[0], st[1], bad);
        if (e != cudaSuccess) break;
    }
    return 0;
}

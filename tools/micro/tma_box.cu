// Micro test: 4-D TMA box {16 x, 4 y, 128 c, 1 b} of an fp16 NCHW tensor with SWIZZLE_128B -- does it load, and how
// does it land in shared memory?   nvcc -gencode arch=compute_100a,code=sm_100a -I ../../squeezedet-pytorch_b200/csrc -o tma_box tma_box.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "tc_ptx.cuh"
using namespace sqd_tc;

__global__ void k(const __grid_constant__ CUtensorMap map, int x0, int y0, int c0, int b0, uint16_t *out, int bytes) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ volatile int abort_flag;
    if (threadIdx.x == 0) {
        abort_flag = 0;
        mbar_init(&bar, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, bytes);
        tma_load_4d(&map, &bar, smem, x0, y0, c0, b0);
    }
    bool ok = mbar_wait(&bar, 0, &abort_flag);
    __syncthreads();
    for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t *>(smem)[i];
    if (threadIdx.x == 0 && !ok) out[0] = 0xDEAD;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 80, H = 24, C = 256, B = 2;
    const int swz = argc > 2 ? atoi(argv[2]) : 3;
    const int bx = argc > 3 ? atoi(argv[3]) : 16, by = argc > 4 ? atoi(argv[4]) : 4, bc = argc > 5 ? atoi(argv[5]) : 128;
    const int xs = argc > 6 ? atoi(argv[6]) : 15;
    std::vector<uint16_t> h((size_t)B * C * H * W);
    for (int b = 0; b < B; ++b) for (int c = 0; c < C; ++c) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x)
        h[(((size_t)b * C + c) * H + y) * W + x] = (uint16_t)((c << 8) | (y << 4) | (x & 15));   // tag: channel, row, col%16
    uint16_t *d, *out;
    cudaMalloc(&d, h.size() * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    const int bytes = bx * by * bc * 2;
    cudaMalloc(&out, bytes);
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    alignas(64) CUtensorMap map;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)C * H * W * 2};
    const cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bc, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     (CUtensorMapSwizzle)swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode W=%d swizzle=%d box {%d,%d,%d} x0=%d -> %d\n", W, swz, bx, by, bc, xs, (int)r);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 2048);
    k<<<1, 128, bytes + 2048>>>(map, xs, 3, 64, 1, out, bytes);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint16_t> o(bytes / 2);
    cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
    // expected (address-based 128B swizzle): element (c, y, x) of the box at row c (128 B), logical byte (y*16 + x)*2,
    // 16-byte chunk index XOR (c & 7)
    int bad = 0;
    for (int c = 0; c < bc; ++c) for (int y = 0; y < by; ++y) for (int x = 0; x < bx; ++x) {
        const int lb = (y * bx + x) * 2;
        const int chunk = swz == 3 ? ((lb >> 4) ^ (c & 7)) : (lb >> 4);
        const uint16_t got = o[(c * 128 + chunk * 16 + (lb & 15)) / 2];
        const int gx = xs + x, gy = 3 + y;
        const uint16_t want = gx < W ? (uint16_t)(((64 + c) << 8) | (gy << 4) | (gx & 15)) : 0;
        if (got != want && bad++ < 5) printf("mismatch c=%d y=%d x=%d got %04x want %04x\n", c, y, x, got, want);
    }
    printf("mismatches: %d of %d\n", bad, bc * by * bx);
    return 0;
}

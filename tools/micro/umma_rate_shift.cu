// Microbenchmark: does a SWIZZLE_128B K-major A operand that starts at a row which is NOT a multiple of 8 (the flat-tile
// tap shift of convdet_fused.cu) cost tensor-pipe time?  Same harness as umma_rate_2cta.cu (M = 256 across a CTA pair,
// K = 16, the production N1 = 144 / N2 = 80 alternation), A descriptors start `shift` rows (x 128 B) into the buffer;
// B either SWIZZLE_128B (+32 B per K step) or SWIZZLE_32B tiles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I squeezedet-pytorch_b200/csrc -o umma_rate_shift umma_rate_shift.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace sqd_tc;

__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ uint64_t desc_sw32(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}

// mode 0: the staged kernel's pattern (4 K steps of one tap: A +32 B, B SW128 +32 B); mode 1: the slice-major pattern
// (9 taps of one K step: A row shifts dy*pw + dx, B SW32 tap tiles)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate_kernel(int mode, int shift, int pw, int iters, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int abort_flag;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        abort_flag = 0;
        mbar_init(&bar, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc_2cta(&slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tm = slot;
    long long t0 = 0, t1 = 0;
    if (warp == 0 && rank == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 100 * 1024);
        const uint32_t id1 = idesc_f16(256, 144), id2 = idesc_f16(256, 80);
        const uint64_t ad0 = umma_desc_sw128(a0) + (uint64_t)(shift * 128 >> 4), ad1 = ad0 + (uint64_t)(36864 >> 4);
        const uint64_t bd128 = umma_desc_sw128(b0), bd32 = desc_sw32(b0);
        t0 = clock64();
        if (elect_one_sync()) {
            if (mode == 0) {
                for (int i = 0; i < iters; i += 8) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * 2);
                        umma_f16_ss_2cta(tm, ad0 + adv, bd128 + adv, id1, 1u);
                        umma_f16_ss_2cta(tm + 256, ad1 + adv, bd128 + adv, id2, 1u);
                    }
                }
            } else {
                for (int i = 0; i < iters; i += 18) {
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint64_t a1 = ad0 + (uint64_t)((dy * pw + dx) * 128 >> 4);
                            const uint64_t b = (mode == 1 ? bd32 + (uint64_t)((dy * 3 + dx) * 2304 >> 4) : bd128 + (uint64_t)((dy * 3 + dx) * 9216 >> 4));
                            umma_f16_ss_2cta(tm, a1, b, id1, 1u);
                            umma_f16_ss_2cta(tm + 256, a1 + (uint64_t)(36864 >> 4), b, id2, 1u);
                        }
                }
            }
            umma_commit_2cta(&bar, 1);
        }
        __syncwarp();
        mbar_wait(&bar, 0, &abort_flag);
        t1 = clock64();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc_2cta(tm, 512);
    }
    if (threadIdx.x == 0 && rank == 0) out[blockIdx.x / 2] = t1 - t0;
}

int main() {
    long long *d;
    cudaMalloc(&d, 148 * sizeof(long long));
    const int iters = 4608;   // multiple of 8 and 18
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024);
    const int cases[][3] = {{0, 0, 0}, {0, 1, 0}, {0, 4, 0}, {0, 7, 0}, {0, 8, 0}, {0, 16, 0},
                            {1, 0, 79}, {1, 0, 80}, {1, 0, 0}, {1, 3, 80}, {2, 0, 79}, {2, 0, 80}, {2, 0, 0}};
    for (auto &c : cases) {
        rate_kernel<<<148, 128, 204 * 1024>>>(c[0], c[1], c[2], iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("mode %d shift %d: %s\n", c[0], c[1], cudaGetErrorString(e));
            return 1;
        }
        long long h[74];
        cudaMemcpy(h, d, 74 * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 74; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("mode %d (%s) A row shift %2d pw %2d: %7.1f cycles per (144 + 80) K step\n", c[0],
               c[0] == 0 ? "4 K steps per tap, B SW128" : c[0] == 1 ? "9 taps per K step, B SW32" : "9 taps per K step, B SW128 tiles", c[1], c[2],
               2.0 * mx / iters);
    }
    return 0;
}

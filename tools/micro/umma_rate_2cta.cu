// Microbenchmark: execution rate of tcgen05.mma.cta_group::2.kind::f16 (M = 256 across a CTA pair, K = 16, SS operands,
// SWIZZLE_128B K-major) as a function of N -- is the cost linear in N or quantised?  Operands are zeros; the leader CTA's
// elected thread issues `iters` MMAs back to back (the production kernel's pattern: 4 K steps per descriptor pair) and
// the time to the commit's completion is measured with clock64.  Also the alternating pattern N1 / N2 of the f16x3
// scheme (A1 x [w1|w2], A2 x w1).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I squeezedet-pytorch_b200/csrc -o umma_rate_2cta umma_rate_2cta.cu
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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate_kernel(int n1, int n2, int iters, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int abort_flag;
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
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
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
        const uint32_t id1 = idesc_f16(256, n1), id2 = idesc_f16(256, n2);
        const uint64_t ad0 = umma_desc_sw128(a0), ad1 = umma_desc_sw128(a0 + 20480), bd = umma_desc_sw128(b0);
        t0 = clock64();
        if (elect_one_sync()) {
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);
                    umma_f16_ss_2cta(tm, ad0 + adv, bd + adv, id1, 1u);
                    umma_f16_ss_2cta(tm + 256, ad1 + adv, bd + adv, id2, 1u);
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
    const int iters = 4800;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
    const int cases[][2] = {{32, 32}, {64, 64}, {80, 80}, {96, 96}, {112, 112}, {128, 128}, {144, 144}, {160, 160}, {192, 192},
                            {224, 224}, {256, 256}, {160, 80}, {144, 80}, {144, 64}, {160, 96}, {128, 96}, {224, 32}};
    for (auto &c : cases)
        for (int grid : {2, 148}) {
            rate_kernel<<<grid, 128, 140 * 1024>>>(c[0], c[1], iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("N1 %d N2 %d: %s\n", c[0], c[1], cudaGetErrorString(e));
                return 1;
            }
            long long h[74];
            cudaMemcpy(h, d, (grid / 2) * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < grid / 2; ++i) mx = h[i] > mx ? h[i] : mx;
            const double per_pair = 2.0 * mx / iters;   // one N1 MMA + one N2 MMA
            printf("N1=%3d N2=%3d pairs=%2d  %7.1f cycles per (N1 + N2) K step   nominal (N1 + N2) / 2 = %5.1f   ratio %.2f\n", c[0],
                   c[1], grid / 2, per_pair, (c[0] + c[1]) / 2.0, per_pair / ((c[0] + c[1]) / 2.0));
        }
    return 0;
}

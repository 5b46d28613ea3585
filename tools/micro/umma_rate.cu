// Microbenchmark: execution rate of tcgen05.mma (cta_group::1, M=128, kind::tf32) as a function of N,
// with A from shared memory (SS) or from tensor memory (TS).  One CTA per SM; operands are zeros; MMAs are
// issued back to back from an unrolled block with precomputed descriptors (like the production kernel).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I squeezedet-pytorch_b200/csrc -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

using namespace sqd_tc;

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc),
                 "r"(idesc), "r"(acc)
                 : "memory");
}

// MODE 0: SS, same A/B tiles.  1: SS, 3 A tiles x 2 B tiles rotating (production pattern).  2: TS (A in TMEM).
// 3: SS pattern of the "concatenated B" scheme: alternate N and N/2 MMAs is emulated by the caller via n.
template <int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int iters, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int abort_flag;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        abort_flag = 0;
        mbar_init(&bar, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
        const uint32_t idesc = umma_idesc_tf32(128, n);
        uint64_t ad[3], bd[2];
        for (int k = 0; k < 3; ++k) ad[k] = umma_desc_sw128(a0 + k * 20480);
        for (int k = 0; k < 2; ++k) bd[k] = umma_desc_sw128(b0 + k * 32768);
        t0 = clock64();
        if (elect_one_sync()) {
            for (int i = 0; i < iters; i += 12) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);
                    if (MODE == 0) {
                        umma_tf32(tm, ad[0] + adv, bd[0] + adv, idesc, 1u);
                        umma_tf32(tm, ad[0] + adv, bd[0] + adv, idesc, 1u);
                        umma_tf32(tm, ad[0] + adv, bd[0] + adv, idesc, 1u);
                    } else if (MODE == 1) {
                        umma_tf32(tm, ad[1] + adv, bd[0] + adv, idesc, 1u);
                        umma_tf32(tm, ad[0] + adv, bd[1] + adv, idesc, 1u);
                        umma_tf32(tm, ad[0] + adv, bd[0] + adv, idesc, 1u);
                    } else {
                        umma_tf32_ts(tm, tm + 320 + ks * 8, bd[0] + adv, idesc, 1u);
                        umma_tf32_ts(tm, tm + 352 + ks * 8, bd[1] + adv, idesc, 1u);
                        umma_tf32_ts(tm, tm + 320 + ks * 8, bd[0] + adv, idesc, 1u);
                    }
                }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0, &abort_flag);
        t1 = clock64();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tm, 512);
    }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, long long *d) {
    const int iters = 4800;
    cudaFuncSetAttribute(rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    for (int n : {16, 64, 80, 96, 112, 128, 160, 256})
        for (int grid : {1, 148}) {
            rate_kernel<MODE><<<grid, 128, 170 * 1024>>>(n, iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("%s n %d: %s\n", name, n, cudaGetErrorString(e));
                exit(1);
            }
            long long h[148];
            cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("%-28s N=%3d grid=%3d  %7.1f cycles/MMA   (N/2 = %5.1f)\n", name, n, grid, (double)mx / iters, n / 2.0);
        }
}

int main(int argc, char **argv) {
    long long *d;
    cudaMalloc(&d, 148 * sizeof(long long));
    run<0>("tf32 SS same tiles", d);
    run<1>("tf32 SS rotating A/B tiles", d);
    if (argc > 1) run<2>("tf32 TS (A in TMEM)", d);
    return 0;
}

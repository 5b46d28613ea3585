// Microbenchmark: cost of tcgen05.commit inside a stream of TS-mode MMAs (production pattern: N=160 then N=80).
// A commit to an mbarrier is inserted after every `every` MMAs (0 = only one at the end).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace sqd_tc;

__device__ __forceinline__ void umma_tf32_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

template <int EVERY, int NBAR, bool SS>
__global__ void __launch_bounds__(128, 1) k(int iters, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, dummy[4];
    __shared__ uint32_t slot;
    __shared__ volatile int abort_flag;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        abort_flag = 0; mbar_init(&bar, 1);
        for (int q = 0; q < 4; ++q) mbar_init(&dummy[q], 1);
        fence_barrier_init(); fence_proxy_async();
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
        const uint32_t id_cat = umma_idesc_tf32(128, 160), id_one = umma_idesc_tf32(128, 80);
        const uint64_t bd = umma_desc_sw128(b0), ad = umma_desc_sw128(a0);
        t0 = clock64();
        if (elect_one_sync()) {
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (SS) {
                        umma_tf32(tm, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), id_one, 1u);
                        umma_tf32(tm, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), id_one, 1u);
                    } else {
                        umma_tf32_ts(tm, tm + 320 + ks * 8, bd + (uint64_t)(ks * 2), id_cat, 1u);
                        umma_tf32_ts(tm, tm + 352 + ks * 8, bd + (uint64_t)(ks * 2), id_one, 1u);
                    }
                    if (EVERY > 0 && ((ks * 2 + 2) % EVERY) == 0) {
#pragma unroll
                        for (int q = 0; q < NBAR; ++q) umma_commit(&dummy[q]);
                    }
                }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0, &abort_flag);
        t1 = clock64();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tm, 512); }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int EVERY, int NBAR, bool SS>
void run(long long *d) {
    const int iters = 4800;
    cudaFuncSetAttribute(k<EVERY, NBAR, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    k<EVERY, NBAR, SS><<<148, 128, 170 * 1024>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); exit(1); }
    long long h[148];
    cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%s  commit x%d after every %d MMAs: %6.1f cycles per MMA\n", SS ? "SS N=80        " : "TS N=160/80 mix", NBAR, EVERY, (double)mx / iters);
}

int main() {
    long long *d; cudaMalloc(&d, 148 * sizeof(long long));
    run<0, 1, false>(d); run<8, 1, false>(d); run<8, 2, false>(d); run<4, 1, false>(d); run<2, 1, false>(d);
    run<0, 1, true>(d);  run<8, 1, true>(d);  run<8, 2, true>(d);  run<4, 1, true>(d);  run<2, 1, true>(d);
    return 0;
}

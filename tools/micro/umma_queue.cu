// Microbenchmark: how far can the issuing thread run ahead of the tensor pipe?  Time to ISSUE k TS-mode MMAs
// (N=160/80 mix, 60 cycles each when executing) vs time until they COMPLETE; plus the cost of an
// mbarrier.try_wait on an already-completed phase, measured in the same warp.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace sqd_tc;

__device__ __forceinline__ void umma_tf32_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

template <int K>
__global__ void __launch_bounds__(128, 1) k(long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, done_bar;
    __shared__ uint32_t slot;
    __shared__ volatile int abort_flag;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        abort_flag = 0; mbar_init(&bar, 1); mbar_init(&done_bar, 1);
        fence_barrier_init(); fence_proxy_async();
        mbar_arrive(&done_bar);  // phase 0 of done_bar is complete from the start
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    if (warp == 0) {
        const uint32_t b0 = smem_u32(smem + 96 * 1024);
        const uint32_t id_cat = umma_idesc_tf32(128, 160), id_one = umma_idesc_tf32(128, 80);
        const uint64_t bd = umma_desc_sw128(b0);
        t0 = clock64();
        if (elect_one_sync()) {
#pragma unroll
            for (int i = 0; i < K; i += 2) {
                umma_tf32_ts(tm, tm + 320 + (i & 6) * 4, bd + (uint64_t)(i & 6), id_cat, 1u);
                umma_tf32_ts(tm, tm + 352 + (i & 6) * 4, bd + (uint64_t)(i & 6), id_one, 1u);
            }
            umma_commit(&bar);
        }
        __syncwarp();
        t1 = clock64();
        mbar_wait(&bar, 0, &abort_flag);
        t2 = clock64();
        // cost of polling an already-complete barrier, warp-uniform flavour and single flavour
        bool r = mbar_wait_warp(&done_bar, 0, &abort_flag);
        t3 = clock64();
        r &= mbar_wait(&done_bar, 0, &abort_flag);
        t4 = clock64();
        if (!r) out[100] = 1;
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tm, 512); }
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = t3 - t2; out[3] = t4 - t3; }
}

template <int K>
void run(long long *d) {
    cudaFuncSetAttribute(k<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    k<K><<<1, 128, 170 * 1024>>>(d);
    k<K><<<1, 128, 170 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); exit(1); }
    long long h[4];
    cudaMemcpy(h, d, 4 * sizeof(long long), cudaMemcpyDeviceToHost);
    printf("k=%3d MMAs: issue returns after %6lld cycles, complete after %6lld (ideal exec %5d) | try_wait on complete barrier: warp-uniform %lld, plain %lld cycles\n",
           K, h[0], h[1], K * 60, h[2], h[3]);
}

int main() {
    long long *d; cudaMalloc(&d, 128 * sizeof(long long));
    run<2>(d); run<4>(d); run<8>(d); run<16>(d); run<24>(d); run<32>(d); run<64>(d); run<128>(d);
    return 0;
}

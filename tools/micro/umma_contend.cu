// Microbenchmark: does other traffic slow TS-mode tcgen05.mma (A in TMEM) down?  Warp 0 issues the production
// pattern (N=160 then N=80 per K step); warps 1..4 generate background traffic:
//   0 none | 1 tcgen05.st into other TMEM columns | 2 tcgen05.ld from the accumulator | 3 LDS.128 smem reads
//   4 STS.128 smem writes | 5 tcgen05.st + wait::st per 64 columns (converter pattern)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace sqd_tc;

__device__ __forceinline__ void umma_tf32_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

__global__ void __launch_bounds__(160, 1) k(int bg, int iters, long long *out, float *sink) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int abort_flag, done;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) { abort_flag = 0; done = 0; mbar_init(&bar, 1); fence_barrier_init(); fence_proxy_async(); }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        const uint32_t b0 = smem_u32(smem + 96 * 1024);
        const uint32_t id_cat = umma_idesc_tf32(128, 160), id_one = umma_idesc_tf32(128, 80);
        uint64_t bd[2];
        for (int q = 0; q < 2; ++q) bd[q] = umma_desc_sw128(b0 + q * 20480);
        t0 = clock64();
        if (elect_one_sync()) {
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    umma_tf32_ts(tm, tm + 320 + ks * 8, bd[0] + (uint64_t)(ks * 2), id_cat, 1u);
                    umma_tf32_ts(tm, tm + 352 + ks * 8, bd[0] + (uint64_t)(ks * 2), id_one, 1u);
                }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0, &abort_flag);
        t1 = clock64();
        done = 1;
    } else {
        const int q = warp & 3;
        uint32_t r[16];
        for (int i = 0; i < 16; ++i) r[i] = i + lane;
        float acc = 0.f;
        const uint32_t lanebase = tm + ((uint32_t)(q * 32) << 16);
        int it = 0;
        while (!done) {
            if (bg == 1) {
                tmem_st_x16(lanebase + 384 + (it & 7) * 16, r);
            } else if (bg == 2) {
                uint32_t v[16];
                tmem_ld_x16(lanebase + 160 + (it & 7) * 16, v);
                tmem_ld_wait();
                acc += __uint_as_float(v[3]);
            } else if (bg == 3) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                             : "r"(smem_u32(smem + ((it * 512 + threadIdx.x * 16) & 0xFFFF))));
                acc += v.x;
            } else if (bg == 4) {
                asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(smem_u32(smem + ((it * 512 + threadIdx.x * 16) & 0xFFFF))), "f"(acc) : "memory");
            } else if (bg == 5) {
                tmem_st_x16(lanebase + 384 + 0, r); tmem_st_x16(lanebase + 384 + 16, r);
                tmem_st_x16(lanebase + 384 + 32, r); tmem_st_x16(lanebase + 384 + 48, r);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            } else {
                __nanosleep(100);
            }
            ++it;
        }
        if (bg == 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (acc == 12345.f) sink[0] = acc;
        if (lane == 0 && blockIdx.x == 0) out[200 + warp] = it;
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tm, 512); }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
    long long *d; float *sink;
    cudaMalloc(&d, 256 * sizeof(long long)); cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    const char *names[] = {"none", "tcgen05.st stream", "tcgen05.ld+wait", "LDS.128", "STS.128", "4x tcgen05.st + wait::st"};
    const int iters = 4800;
    for (int bg = 0; bg < 6; ++bg) {
        cudaMemset(d, 0, 256 * sizeof(long long));
        k<<<148, 160, 170 * 1024>>>(bg, iters, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("bg %d: %s\n", bg, cudaGetErrorString(e)); return 1; }
        long long h[256];
        cudaMemcpy(h, d, 256 * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("background %-26s: %6.1f cycles per MMA (avg of N=160 and N=80; ideal 60)   bg iterations/warp %lld over %lld cycles\n",
               names[bg], (double)mx / iters, h[201], mx);
    }
    return 0;
}

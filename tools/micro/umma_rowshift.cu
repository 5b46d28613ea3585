// Micro test: can a K-major SWIZZLE_128B UMMA operand start at a ROW offset that is not a multiple of the 8-row swizzle
// atom (start address = 1024-aligned base + s * 128 B)?  This is what a one-pixel tap shift of an implicit-GEMM A patch
// needs (the ConvDet kernel re-fetches the patch per dx because it assumed the answer is no).
//   D_s[m][n] = sum_k A[m + s][k] * B[n][k],  m < 128, n < 16, k < 64,  for s = 0..8
// A (144 rows) and B (16 rows) are loaded with TMA (SWIZZLE_128B, 128-byte rows), the MMA is tcgen05.mma kind::f16
// M = 128, N = 16, four K = 16 steps.  mode 0: descriptor base_offset field = 0; mode 1: base_offset = (addr >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I ../../squeezedet-pytorch_b200/csrc -o umma_rowshift umma_rowshift.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "tc_ptx.cuh"
using namespace sqd_tc;

constexpr int kRows = 144, kN = 16, kK = 64, kShifts = 9;

__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                         int mode, float *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sa = smem, *sb = smem + 20 * 1024;   // A: 144 rows x 128 B = 18 KB
    __shared__ __align__(8) uint64_t bar_ld, bar_mma;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int abort_flag;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        abort_flag = 0;
        mbar_init(&bar_ld, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar_ld, (kRows + kN) * 128);
        tma_load_2d(&map_a, &bar_ld, sa, 0, 0);
        tma_load_2d(&map_b, &bar_ld, sb, 0, 0);
    }
    mbar_wait(&bar_ld, 0, &abort_flag);
    tc_fence_after();
    for (int s = 0; s < kShifts; ++s) {
        if (threadIdx.x == 0) {
            const uint32_t a_addr = smem_u32(sa) + s * 128;
            uint64_t ad = umma_desc_sw128(a_addr);
            if (mode == 1) ad |= (uint64_t)((a_addr >> 7) & 7u) << 49;   // base_offset
            const uint64_t bd = umma_desc_sw128(smem_u32(sb));
            for (int ks = 0; ks < kK / 16; ++ks) {
                const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                umma_f16_ss(tmem, ad + adv, bd + adv, idesc_f16(128, kN), ks ? 1u : 0u);
            }
            umma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, s & 1, &abort_flag);
        tc_fence_after();
        uint32_t r[16];
        tmem_ld_x16(tmem + ((uint32_t)(warp * 32) << 16), r);
        tmem_ld_wait();
        for (int n = 0; n < kN; ++n) out[((size_t)s * 128 + threadIdx.x) * kN + n] = __uint_as_float(r[n]);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 0) tmem_dealloc(tmem, 32);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    std::vector<__half> ha((size_t)kRows * kK), hb((size_t)kN * kK);
    std::vector<float> fa(ha.size()), fb(hb.size());
    srand(7);
    for (size_t i = 0; i < ha.size(); ++i) { fa[i] = (float)(rand() % 17 - 8); ha[i] = __float2half(fa[i]); }
    for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (float)(rand() % 9 - 4); hb[i] = __float2half(fb[i]); }
    __half *da, *db;
    float *dout;
    cudaMalloc(&da, ha.size() * 2);
    cudaMalloc(&db, hb.size() * 2);
    cudaMalloc(&dout, (size_t)kShifts * 128 * kN * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    alignas(64) CUtensorMap ma, mb;
    const cuuint32_t es[2] = {1, 1};
    {
        const cuuint64_t dims[2] = {kK, kRows};
        const cuuint64_t strides[1] = {kK * 2};
        const cuuint32_t box[2] = {kK, kRows};
        CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, da, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode A failed %d\n", (int)r); return 1; }
    }
    {
        const cuuint64_t dims[2] = {kK, kN};
        const cuuint64_t strides[1] = {kK * 2};
        const cuuint32_t box[2] = {kK, kN};
        CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode B failed %d\n", (int)r); return 1; }
    }
    const int smem = 24 * 1024 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(dout, 0, (size_t)kShifts * 128 * kN * 4);
        k<<<1, 128, smem>>>(ma, mb, mode, dout);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mode %d (base_offset %s): kernel %s\n", mode, mode ? "= (addr>>7)&7" : "= 0", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<float> o((size_t)kShifts * 128 * kN);
        cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
        for (int s = 0; s < kShifts; ++s) {
            int bad = 0, first_m = -1;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < kN; ++n) {
                    float want = 0.f;
                    for (int kk = 0; kk < kK; ++kk) want += fa[(size_t)(m + s) * kK + kk] * fb[(size_t)n * kK + kk];
                    if (o[((size_t)s * 128 + m) * kN + n] != want) {
                        if (first_m < 0) first_m = m;
                        ++bad;
                    }
                }
            printf("  row shift %d: %s (%d of %d wrong%s)\n", s, bad ? "MISMATCH" : "exact", bad, 128 * kN,
                   bad ? (first_m == 0 ? ", from row 0" : ", later rows") : "");
        }
    }
    return 0;
}

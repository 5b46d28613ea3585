// a1 (validation yardstick): ConvDet 3x3 head as a CUDA-core fp32 FMA implicit GEMM.
// Reference: SqueezeDetBase.convdet + permute/view, src/model/squeezedet.py:73-75,83-87.
//
// This is NOT the production head (that is convdet_f16.cu / convdet_fused.cu: tcgen05 / TMEM / TMA, fp16x3).  It exists
// so that the tensor-core kernel can be checked on the GPU at full size against an independent
// fp32 evaluation of the same contraction (plain fmaf accumulation, k = tap-major, channel-minor),
// and it accepts NCHW features directly.  GEMM view: M = B*gh*gw cells, N = Cout, K = 9*Cin.
#include "common.cuh"

namespace {

constexpr int TM = 64;   // cells per CTA
constexpr int TK = 16;   // k-slab (one tap, 16 channels)
constexpr int TNMAX = 128;
constexpr int kThreads = 256;

// (Cout, Cin, 3, 3) -> Wt[(tap*Cin + c)][n], n padded to a multiple of 16 with zeros
__global__ void simt_pack_kernel(const float *w, int cout, int cin, int npad, float *wt) {
    const long long total = (long long)9 * cin * npad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i % npad);
        const long long k = i / npad;
        const int c = (int)(k % cin), tap = (int)(k / cin);
        wt[i] = n < cout ? w[((size_t)n * cin + c) * 9 + tap] : 0.f;
    }
}

template <int NJ>  // NJ = npad / 16 output columns per thread
__global__ void __launch_bounds__(kThreads) convdet_simt_kernel(const float *feat, int layout, const float *wt,
                                                                const float *bias, int B, int cin, int gh, int gw,
                                                                int cout, float *pred) {
    constexpr int TN = NJ * 16;
    __shared__ __align__(16) float As[TK][TM];
    __shared__ __align__(16) float Bs[TK][TN];
    const int cells = gh * gw;
    const long long M = (long long)B * cells;
    const long long m0 = (long long)blockIdx.x * TM;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // tx -> n, ty -> m

    float acc[4][NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;

    const int chunks_per_tap = cin / TK;
    for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        for (int ch = 0; ch < chunks_per_tap; ++ch) {
            const int c0 = ch * TK;
            // A slab: As[kk][cell]
#pragma unroll
            for (int i = 0; i < (TM * TK) / kThreads; ++i) {
                int cell, kk;
                if (layout == SQD_LAYOUT_NCHW) {
                    cell = threadIdx.x % TM;
                    kk = threadIdx.x / TM + (kThreads / TM) * i;
                } else {
                    kk = threadIdx.x % TK;
                    cell = threadIdx.x / TK + (kThreads / TK) * i;
                }
                const long long m = m0 + cell;
                float v = 0.f;
                if (m < M) {
                    const int b = (int)(m / cells), r = (int)(m % cells);
                    const int y = r / gw + dy, x = r % gw + dx;
                    if (y >= 0 && y < gh && x >= 0 && x < gw) {
                        const int c = c0 + kk;
                        v = layout == SQD_LAYOUT_NCHW ? __ldg(feat + (((size_t)b * cin + c) * gh + y) * gw + x)
                                                      : __ldg(feat + (((size_t)b * gh + y) * gw + x) * cin + c);
                    }
                }
                As[kk][cell] = v;
            }
            // B slab: Bs[kk][n]
            const float *wrow = wt + ((size_t)tap * cin + c0) * TN;
            for (int i = threadIdx.x; i < TK * TN; i += kThreads) Bs[i / TN][i % TN] = __ldg(wrow + i);
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < TK; ++kk) {
                const float4 a4 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                float bv[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int n = tx + 16 * j;
            if (n < cout) pred[(size_t)m * cout + n] = acc[i][j] + __ldg(bias + n);
        }
    }
}

}  // namespace

size_t sqd_simt_workspace_bytes(int cin, int cout) {
    const int npad = (cout + 15) / 16 * 16;
    return (size_t)9 * cin * npad * sizeof(float);
}

int sqd_convdet_simt(const float *d_feat, int layout, const float *d_weight, const float *d_bias, int batch, int cin,
                     int gh, int gw, int cout, float *d_pred, void *d_workspace, cudaStream_t st) {
    const int npad = (cout + 15) / 16 * 16;
    SQD_REQUIRE(npad <= TNMAX, SQD_E_SHAPE, "convdet (simt): Cout %d > %d", cout, TNMAX);
    SQD_REQUIRE(cin % TK == 0, SQD_E_SHAPE, "convdet (simt): Cin %d must be a multiple of %d", cin, TK);
    float *wt = static_cast<float *>(d_workspace);
    simt_pack_kernel<<<2 * SQD_SM_COUNT, 256, 0, st>>>(d_weight, cout, cin, npad, wt);
    SQD_LAUNCH_CHECK("simt_pack_kernel");
    const long long M = (long long)batch * gh * gw;
    const unsigned grid = (unsigned)((M + TM - 1) / TM);
#define SQD_SIMT_CASE(NJ)                                                                                           \
    case NJ:                                                                                                        \
        convdet_simt_kernel<NJ><<<grid, kThreads, 0, st>>>(d_feat, layout, wt, d_bias, batch, cin, gh, gw, cout, d_pred); \
        break;
    switch (npad / 16) {
        SQD_SIMT_CASE(1) SQD_SIMT_CASE(2) SQD_SIMT_CASE(3) SQD_SIMT_CASE(4)
        SQD_SIMT_CASE(5) SQD_SIMT_CASE(6) SQD_SIMT_CASE(7) SQD_SIMT_CASE(8)
        default:
            SQD_REQUIRE(false, SQD_E_SHAPE, "convdet (simt): unsupported Cout %d", cout);
    }
#undef SQD_SIMT_CASE
    SQD_LAUNCH_CHECK("convdet_simt_kernel");
    return SQD_OK;
}

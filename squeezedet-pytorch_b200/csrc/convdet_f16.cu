// a1: ConvDet 3x3 head as a persistent tcgen05 / TMEM implicit GEMM fed by TMA (sm_100a), "f16x3":
// fp32-level products from THREE half-precision tensor-core passes.
// Reference: SqueezeDetBase.convdet + permute(0,2,3,1) + view, src/model/squeezedet.py:73-75,83-87
// (cuDNN conv with N=72 plus an NCHW->NHWC copy kernel there).
//
// GEMM view per image: M = gh*gw cells, N = Cout = K_anchors*(C+5) (72 KITTI, padded to 80), K = 9*Cin = 6912.
//
// Why fp16 and not tf32: both carry an 11-bit significand, but kind::f16 runs at twice the kind::tf32 rate and
// its operands are half as wide, which matters because this small-N GEMM is SHARED-MEMORY bound (an SS-mode
// tcgen05.mma streams A (128 rows) and B (N rows) through the 128 B/clk port; tools/micro/umma_rate.cu).  What
// fp16 lacks is exponent range, so both operands are scaled by exact powers of two first:
//     x*s = x1 + x2/2^11 (+ <= 2^-22 relative),  x1 = fp16(x*s),  x2 = fp16((x*s - x1) * 2^11)
// with s = 2^e chosen per (IMAGE, 64-CHANNEL BLOCK) for the features and per tensor for the weights so that
// max|x*s| lies in [2^13, 2^14): no overflow, and elements down to 2^-27 of the block maximum keep full precision
// (smaller ones degrade gracefully to an absolute error of 2^-49 of the maximum).  Then
//     a*w*(s_a*s_w) = a1*w1 + (a1*w2 + a2*w1)/2^11   (the dropped a2*w2 term is 2^-22 relative, like 3xTF32)
// A TMEM accumulation chunk never crosses a channel block, and the accumulate warps multiply it by the block's 1/s_a
// (exact) while folding it into their fp32 registers.
//
//  * pre-pass (the only extra HBM traffic): features fp32 NCHW -> two fp16 NHWC planes (x1, x2) in ONE pass: a
//    thread-block cluster holds an (image, channel block) slab in shared memory, finds its max and writes the split,
//    transposed planes (4 B read + 4 B written per element; the K-major A operand needs the transpose anyway).
//    channels_last input and oversized grids take a max pass + a split pass.
//  * M tile = 8 x 16 cells = 128 rows = one UMMA_M.  A "unit" of work is (tile, 64-channel block, dx): ONE 4-D TMA
//    box {64 ch, 16 x, 10 y, 1 img} per plane at (c0, x0+dx-1, y0-1, b); conv padding = TMA out-of-bounds zero fill.
//    The box lands as 160 rows x 128 B, 128B-swizzled; the three dy taps are the SAME patch read through UMMA
//    descriptors offset by 16 rows (2048 B, swizzle-atom aligned) -> A is fetched 3x per channel block, not 9x.
//  * B = packed weights, one fp16 matrix [2*Npad][9*Cin] K-major (k = tap*Cin + c): rows [0,Npad) hold w2, rows
//    [Npad,2Npad) hold w1, so per K step (16 channels) only TWO MMAs are issued:
//        D[:, 0:2N]  (+)= A1 * [w2 | w1]^T     (N = 2*Npad: cross term a1*w2 | main term a1*w1)
//        D[:, 0:N]    +=  A2 * w1^T            (cross term a2*w1, same 2^11 scale as a1*w2)
//  * Chunked accumulation: the tensor core truncates when adding into the fp32 TMEM accumulator (measured:
//    profiles/r01_tc_accuracy_vs_chunk.txt), so every unit (24 MMAs) accumulates from zero into one of two TMEM
//    accumulators and four accumulate warps fold finished units into fp32 registers (main + cross/2^11, round to
//    nearest) while the next unit's MMAs run into the other accumulator.
//  * Persistent, balanced schedule: grid = min(#SMs, #tiles); the unit range is cut evenly, so a CTA owns
//    [tail of a tile][whole tiles][head of a tile].  A split tile is finished deterministically: the head
//    holder publishes its partial sums, the tail holder (higher CTA index, its tail segment is processed
//    LAST) adds them in a fixed order.  Waiters only ever wait for lower-indexed CTAs.
//  * Warp roles (224 threads): 0 A-TMA, 1 TMEM alloc + MMA issue (converged warp, one elected lane), 2 B-TMA,
//    3..6 accumulate + epilogue (x 1/(s_a*s_w), + bias -> pred in the reference's (B, A, C+5) layout).
//  * Every wait is bounded: on timeout the CTA raises a status word and drains instead of hanging.
// Algorithmic FLOPs per image: 2*M*Cout*K (the three passes and the N padding are NOT counted).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <cooperative_groups.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace sqd_tc;

constexpr int kTileX = 16, kTileY = 8, kPatchY = kTileY + 2;
constexpr int kBlockK = 64;   // channels per unit (128 B of fp16 = one swizzle row)
constexpr int kUmmaK = 16;    // f16 MMA K
constexpr int kPlaneBytes = kPatchY * kTileX * kBlockK * 2;  // 20480: one plane of one A stage
constexpr int kAStageBytes = 2 * kPlaneBytes;                // x1 patch then x2 patch
constexpr int kDyBytes = kTileX * kBlockK * 2;               // 2048: one y row of the patch = descriptor step per dy
constexpr int kThreads = 224;
constexpr int kWarpATma = 0, kWarpMma = 1, kWarpBTma = 2, kWarpAcc0 = 3;
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;    // 2^11: the second term is stored scaled up
constexpr int kHeaderBytes = 256;                            // packed weights: [header][fp16 matrix]

struct PackedHeader {
    unsigned amax_bits;  // max |w| as fp32 bits
    float scale;         // s_w = 2^e
    float inv_scale;     // 1 / s_w
    int npad, cin;
    int hr;              // output channels per CTA in the CTA-pair layout (pair_hr_of)
};

// power-of-two scale s with amax*s in [2^13, 2^14); exponent clamped so that s and 1/s are normal floats
__device__ __forceinline__ float pow2_scale_for(float amax) {
    if (!(amax > 0.f) || amax > 3.0e38f) return 1.f;
    int ex;
    frexpf(amax, &ex);  // amax = m * 2^ex, m in [0.5, 1)
    int e = 14 - ex;
    e = e < -126 ? -126 : (e > 126 ? 126 : e);
    return ldexpf(1.f, e);
}

__device__ __forceinline__ void split_f16(float xs, __half &h1, __half &h2) {
    h1 = __float2half_rn(xs);
    h2 = __float2half_rn((xs - __half2float(h1)) * kLoScale);
}

// ---------------------------------------------------------------------------------------------------
// pre-passes: fp32 features -> two fp16 NHWC planes, scaled per (image, 64-channel block)
// ---------------------------------------------------------------------------------------------------
// Scale granularity.  One tcgen05.mma needs a single scale for all rows of its A slab, and a slab is (pixels of one
// tile, 16 channels of ONE 64-channel block), so the finest granularity that keeps the GEMM exact is (image, channel
// block): the accumulate warps multiply each drained chunk (which never crosses a channel block) by the block's
// 2^-e while folding it into fp32 registers (a power of two: exact).  An (image, block) slab of an NCHW tensor is one
// contiguous run of 64*P floats, small enough for a thread-block CLUSTER to hold in shared memory -- so max|x| and
// the split need ONE pass over HBM (4 B read + 4 B written per element) instead of a max pass plus a split pass.

// max |x| of contiguous runs of n4 float4 (one run per blockIdx.y).  Non-negative floats order like their bit
// patterns, so the reduction is an integer atomicMax.  NaN inputs are ignored by fmaxf.
__global__ void __launch_bounds__(256) absmax_kernel(const float4 *__restrict__ in, size_t n4, unsigned *__restrict__ amax_bits) {
    const float4 *src = in + (size_t)blockIdx.y * n4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    sqd_pdl_trigger();   // a GEMM launched behind this kernel as a programmatic dependent may set up while it runs
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {  // four independent 16-byte loads in flight per thread
        const float4 a = ld_stream_f4(src + i), b = ld_stream_f4(src + i + stride);
        const float4 c = ld_stream_f4(src + i + 2 * stride), d = ld_stream_f4(src + i + 3 * stride);
        m0 = fmaxf(fmaxf(m0, fmaxf(fabsf(a.x), fabsf(a.y))), fmaxf(fabsf(a.z), fabsf(a.w)));
        m1 = fmaxf(fmaxf(m1, fmaxf(fabsf(b.x), fabsf(b.y))), fmaxf(fabsf(b.z), fabsf(b.w)));
        m2 = fmaxf(fmaxf(m2, fmaxf(fabsf(c.x), fabsf(c.y))), fmaxf(fabsf(c.z), fabsf(c.w)));
        m3 = fmaxf(fmaxf(m3, fmaxf(fabsf(d.x), fabsf(d.y))), fmaxf(fabsf(d.z), fabsf(d.w)));
    }
    for (; i < n4; i += stride) {
        const float4 a = ld_stream_f4(src + i);
        m0 = fmaxf(fmaxf(m0, fmaxf(fabsf(a.x), fabsf(a.y))), fmaxf(fabsf(a.z), fabsf(a.w)));
    }
    float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float s[32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) atomicMax(amax_bits + blockIdx.y, __float_as_uint(m));
    }
}

// NHWC (channels_last): per (image, channel block) max.  blockDim.x = cin/4: thread t owns channel quad t (so its
// channel block is fixed), the block walks cells blockIdx.x, +gridDim.x, ...; 16 consecutive lanes share a block.
__global__ void absmax_nhwc_kernel(const float4 *__restrict__ in, int P, int cin4, unsigned *__restrict__ amax_bits) {
    const float4 *src = in + (size_t)blockIdx.y * P * cin4 + threadIdx.x;
    float m0 = 0.f, m1 = 0.f;
    int q = blockIdx.x;
    for (; q + (int)gridDim.x < P; q += 2 * gridDim.x) {
        const float4 a = ld_stream_f4(src + (size_t)q * cin4), b = ld_stream_f4(src + (size_t)(q + gridDim.x) * cin4);
        m0 = fmaxf(fmaxf(m0, fmaxf(fabsf(a.x), fabsf(a.y))), fmaxf(fabsf(a.z), fabsf(a.w)));
        m1 = fmaxf(fmaxf(m1, fmaxf(fabsf(b.x), fabsf(b.y))), fmaxf(fabsf(b.z), fabsf(b.w)));
    }
    if (q < P) {
        const float4 a = ld_stream_f4(src + (size_t)q * cin4);
        m0 = fmaxf(fmaxf(m0, fmaxf(fabsf(a.x), fabsf(a.y))), fmaxf(fabsf(a.z), fabsf(a.w)));
    }
    float m = fmaxf(m0, m1);
#pragma unroll
    for (int o = 8; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));  // within the 16 lanes of a block
    if ((threadIdx.x & 15) == 0) atomicMax(amax_bits + (size_t)blockIdx.y * (cin4 >> 4) + (threadIdx.x >> 4), __float_as_uint(m));
}

// NHWC fp32 -> x1 / x2 fp16 planes (same layout); blockDim.x = cin/4 like absmax_nhwc_kernel; grid.y = image
__global__ void split_nhwc_f16_kernel(const float4 *__restrict__ in, uint2 *__restrict__ p1, uint2 *__restrict__ p2,
                                      int P, int cin4, const unsigned *__restrict__ amax_bits) {
    const float s = pow2_scale_for(__uint_as_float(amax_bits[(size_t)blockIdx.y * (cin4 >> 4) + (threadIdx.x >> 4)]));
    const size_t base = (size_t)blockIdx.y * P * cin4 + threadIdx.x;
    for (int q = blockIdx.x; q < P; q += gridDim.x) {
        const float4 v = ld_stream_f4(in + base + (size_t)q * cin4);
        __half a1[4], a2[4];
        split_f16(v.x * s, a1[0], a2[0]);
        split_f16(v.y * s, a1[1], a2[1]);
        split_f16(v.z * s, a1[2], a2[2]);
        split_f16(v.w * s, a1[3], a2[3]);
        p1[base + (size_t)q * cin4] = *reinterpret_cast<const uint2 *>(a1);
        p2[base + (size_t)q * cin4] = *reinterpret_cast<const uint2 *>(a2);
    }
}

// NCHW (B,Cin,P) fp32 -> NHWC (B,P,Cin) fp16 planes through a 64 ch x 64 cell shared tile, scales already known
// (two-pass fallback for shapes the one-pass kernel does not take).  P = gh*gw.  256 threads.
__global__ void __launch_bounds__(256) split_nchw_f16_kernel(const float *__restrict__ in, uint2 *__restrict__ p1,
                                                             uint2 *__restrict__ p2, int cin, int P,
                                                             const unsigned *__restrict__ amax_bits) {
    __shared__ float tile[64][65];
    const int b = blockIdx.z, c0 = blockIdx.y * 64, q0 = blockIdx.x * 64;
    const float s = pow2_scale_for(__uint_as_float(amax_bits[(size_t)b * (cin >> 6) + blockIdx.y]));
    const float *src = in + (size_t)b * cin * P;
    const int col = threadIdx.x & 63, r0 = threadIdx.x >> 6;
    const int q = q0 + col;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = q < P ? __ldg(src + (size_t)(c0 + r0 + 4 * j) * P + q) : 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) tile[r0 + 4 * j][col] = v[j];
    __syncthreads();
    const int l16 = threadIdx.x & 15, g = threadIdx.x >> 4;  // 16 lanes x 4 channels per cell, 16 cells per pass
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int j = g + 16 * jj;
        const int qq = q0 + j;
        if (qq < P) {
            __half a1[4], a2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_f16(tile[4 * l16 + e][j] * s, a1[e], a2[e]);
            const size_t o = (((size_t)b * P + qq) * cin + c0) / 4 + l16;
            p1[o] = *reinterpret_cast<const uint2 *>(a1);
            p2[o] = *reinterpret_cast<const uint2 *>(a2);
        }
    }
}

// ONE-PASS NCHW split.  A cluster of `cs` CTAs owns one (image, 64-channel block) slab = 64 rows of P floats; CTA r
// takes the float4 columns [r*n4/cs, (r+1)*n4/cs) of every row (n4 = P/4).
//   1. stream the sub-slab into shared memory (16-byte loads; 16-byte chunks XOR-swizzled by the channel so that the
//      transposed reads of step 3 are at most 2-way bank conflicted), tracking max|x|;
//   2. block max -> pushed into every CTA of the cluster through distributed shared memory -> cluster.sync();
//   3. scale by 2^e, split into (x1, x2) and write the NHWC planes: 16 lanes x 4 channels = one 128-byte row of a cell.
constexpr int kSplitMaxCluster = 16;
template <int kSplitThreads, bool kRowLoads>
__global__ void __launch_bounds__(kSplitThreads) split_nchw_cluster_kernel(const float *__restrict__ in, uint2 *__restrict__ p1,
                                                                           uint2 *__restrict__ p2, int cin, int P,
                                                                           unsigned *__restrict__ amax_bits, int cs, int nq4p) {
    extern __shared__ __align__(16) float4 tile4[];  // [64][nq4p]
    __shared__ float s_red[kSplitThreads / 32];
    __shared__ unsigned s_cmax[kSplitMaxCluster];
    namespace cgx = cooperative_groups;
    cgx::cluster_group cluster = cgx::this_cluster();
    const int rank = (int)cluster.block_rank();
    // index arithmetic without integer divisions (they were a quarter of this kernel's instructions): cs is a power of two
    const int csh = 31 - __clz(cs);
    const int slab = blockIdx.x >> csh, ncb = cin >> 6;
    const int b = slab / ncb, cb = slab - b * ncb, c0 = cb << 6;
    const int n4 = P >> 2;
    const int q4b = (rank * n4) >> csh, q4e = ((rank + 1) * n4) >> csh, nq4 = q4e - q4b;
    const float4 *src = reinterpret_cast<const float4 *>(in + ((size_t)b * cin + c0) * P) + q4b;
    sqd_pdl_trigger();   // the GEMM behind this kernel may be scheduled as SMs drain; it waits for this grid to complete

    // 1. stream the sub-slab in: 4 independent 16-byte loads in flight per thread
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kWarps = kSplitThreads / 32;
    float m = 0.f;
    if (kRowLoads) {
        // a warp covers a row segment; rows warp, warp + kWarps, ... (no index division)
        for (int c0r = warp; c0r < 64; c0r += 4 * kWarps) {
#pragma unroll 2
            for (int jb = 0; jb < nq4; jb += 32) {
                const int j = jb + lane;
                float4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    v[i] = (j < nq4 && c0r + kWarps * i < 64) ? ld_stream_f4(src + (size_t)(c0r + kWarps * i) * n4 + j)
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = c0r + kWarps * i;
                    if (j < nq4 && c < 64) {
                        m = fmaxf(fmaxf(m, fmaxf(fabsf(v[i].x), fabsf(v[i].y))), fmaxf(fabsf(v[i].z), fabsf(v[i].w)));
                        tile4[c * nq4p + (j ^ ((c >> 2) & 7))] = v[i];
                    }
                }
            }
        }
    } else {
        const int total = 64 * nq4;
        // idx / nq4 as a multiply-high: exact for idx < 2^16 and nq4 < 256 (idx < 64 * nq4 here)
        const bool use_magic = nq4 < 256;
        const unsigned magic = use_magic ? (unsigned)(0xFFFFFFFFu / (unsigned)nq4) + 1u : 0u;
        for (int base = 0; base < total; base += 4 * kSplitThreads) {
            float4 v[4];
            int cc[4], jj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * kSplitThreads + threadIdx.x;
                cc[u] = use_magic ? (int)__umulhi((unsigned)idx, magic) : idx / nq4;
                jj[u] = idx - cc[u] * nq4;
                v[u] = idx < total ? ld_stream_f4(src + (size_t)cc[u] * n4 + jj[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (base + u * kSplitThreads + (int)threadIdx.x < total) {
                    m = fmaxf(fmaxf(m, fmaxf(fabsf(v[u].x), fabsf(v[u].y))), fmaxf(fabsf(v[u].z), fabsf(v[u].w)));
                    tile4[cc[u] * nq4p + (jj[u] ^ ((cc[u] >> 2) & 7))] = v[u];
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < kSplitThreads / 32 ? s_red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((int)threadIdx.x < cs) cluster.map_shared_rank(s_cmax, threadIdx.x)[rank] = __float_as_uint(m);
    }
    cluster.sync();  // every CTA's maximum has landed in every CTA's s_cmax; also the block barrier for tile4
    unsigned mb = 0u;
    for (int r = 0; r < cs; ++r) mb = max(mb, s_cmax[r]);
    if (rank == 0 && threadIdx.x == 0) amax_bits[slab] = mb;
    const float s = pow2_scale_for(__uint_as_float(mb));

    // 3. 16 lanes x 4 channels = the 128-byte row of one cell and plane; 32 cells per pass
    const int l16 = threadIdx.x & 15, slot = threadIdx.x >> 4;
    const float *tile = reinterpret_cast<const float *>(tile4);
    const int sw = l16 & 7;  // == ((4*l16 + k) >> 2) & 7 for k = 0..3
    const float *trow = tile + (size_t)(4 * l16) * nq4p * 4;
    const int kstride = nq4p * 4;
    const size_t cin4 = (size_t)(cin >> 2);
    size_t o = ((size_t)b * P + 4 * q4b + slot) * cin4 + (c0 >> 2) + l16;
    for (int q = slot; q < 4 * nq4; q += kSplitThreads / 16, o += (kSplitThreads / 16) * cin4) {
        const float *src_q = trow + ((q >> 2) ^ sw) * 4 + (q & 3);
        const float x0 = src_q[0] * s, x1 = src_q[kstride] * s, x2 = src_q[2 * kstride] * s, x3 = src_q[3 * kstride] * s;
        // two-term split, two elements per instruction where the ISA has a packed form
        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn((x0 - f01.x) * kLoScale, (x1 - f01.y) * kLoScale);
        const __half2 l23 = __floats2half2_rn((x2 - f23.x) * kLoScale, (x3 - f23.y) * kLoScale);
        uint2 w1, w2;
        w1.x = *reinterpret_cast<const unsigned *>(&h01);
        w1.y = *reinterpret_cast<const unsigned *>(&h23);
        w2.x = *reinterpret_cast<const unsigned *>(&l01);
        w2.y = *reinterpret_cast<const unsigned *>(&l23);
        p1[o] = w1;
        p2[o] = w2;
    }
}

// ONE-PASS NCHW split, register resident (opt-in SQD_SPLIT_REGS=1, for slabs of at most 8 x 64 float4 columns: P <= 2048;
// bit-identical planes, MEASURED SLOWER than the shared-memory kernel below: step 156.0 vs 147.2 us at KITTI B = 20 -- 128
// registers x 256 threads leave two CTAs = 16 warps per SM to hide the shuffle / convert chain; with the rows stored
// straight from registers, 32 lines per store instruction, it was 184.6 us.  profiles/r02_prepass_register_resident.txt).
// Same decomposition -- a cluster of 8 CTAs per (image, 64-channel block) slab, CTA r the float4 columns
// [r*n4/8, (r+1)*n4/8) -- but nothing goes through shared memory: a thread issues ALL 16 of its 16-byte loads up front
// (channels 4s + cl, s = 0..15, of one column: 64 KB in flight per CTA), the block maximum goes round the cluster, and the
// NCHW -> NHWC transposition is a 4 x 4 register transpose per load step among the four lanes of a column (two rounds of
// two shuffles).  A lane ends up with all 64 channels of ONE cell = the 128-byte row of that cell in each plane, which it
// writes with 16-byte stores.  No staging tile, no block barrier besides the cluster's; planes bit-identical to the
// shared-memory kernel's (same scale, same split arithmetic).
constexpr int kRsThreads = 256, kRsCluster = 8;
__global__ void __cluster_dims__(kRsCluster, 1, 1) __launch_bounds__(kRsThreads, 2)
split_nchw_regs_cluster_kernel(const float *__restrict__ in, uint4 *__restrict__ p1, uint4 *__restrict__ p2, int cin, int P,
                               unsigned *__restrict__ amax_bits) {
    __shared__ float s_red[kRsThreads / 32];
    __shared__ unsigned s_cmax[kRsCluster];
    namespace cgx = cooperative_groups;
    cgx::cluster_group cluster = cgx::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int slab = blockIdx.x / kRsCluster, ncb = cin >> 6;
    const int b = slab / ncb, cb = slab - b * ncb, c0 = cb << 6;
    const int n4 = P >> 2;
    const int q4b = (rank * n4) / kRsCluster, q4e = ((rank + 1) * n4) / kRsCluster;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci = lane >> 2, cl = lane & 3;               // column within the warp's group of 8, channel within a quad
    const int chunk = q4b + warp * 8 + ci;
    const bool valid = chunk < q4e;
    const float4 *src = reinterpret_cast<const float4 *>(in + ((size_t)b * cin + c0) * P) + chunk;
    sqd_pdl_trigger();   // the GEMM behind this kernel may be scheduled as SMs drain; it waits for this grid to complete

    float4 v[16];
#pragma unroll
    for (int s = 0; s < 16; ++s)
        v[s] = valid ? ld_stream_f4(src + (size_t)(4 * s + cl) * n4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float m = 0.f;
#pragma unroll
    for (int s = 0; s < 16; ++s)
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v[s].x), fabsf(v[s].y))), fmaxf(fabsf(v[s].z), fabsf(v[s].w)));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < kRsThreads / 32 ? s_red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((int)threadIdx.x < kRsCluster) cluster.map_shared_rank(s_cmax, threadIdx.x)[rank] = __float_as_uint(m);
    }
    cluster.sync();  // every CTA's maximum has landed in every CTA's s_cmax
    unsigned mb = 0u;
#pragma unroll
    for (int r = 0; r < kRsCluster; ++r) mb = max(mb, s_cmax[r]);
    if (rank == 0 && threadIdx.x == 0) amax_bits[slab] = mb;
    const float sc = pow2_scale_for(__uint_as_float(mb));

    // transpose + split: after step s this lane holds channels 4s..4s+3 of cell 4*chunk + cl
    const bool hi2 = (cl & 2) != 0, hi1 = (cl & 1) != 0;
    uint2 w1[16], w2[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const float4 t = v[s];
        // round 1 (partner cl ^ 2): swap the off-diagonal 2 x 2 blocks
        const float r0 = __shfl_xor_sync(0xffffffffu, hi2 ? t.x : t.z, 2), r1 = __shfl_xor_sync(0xffffffffu, hi2 ? t.y : t.w, 2);
        const float ax = hi2 ? r0 : t.x, ay = hi2 ? r1 : t.y, az = hi2 ? t.z : r0, aw = hi2 ? t.w : r1;
        // round 2 (partner cl ^ 1): transpose inside the 2 x 2 blocks
        const float q0 = __shfl_xor_sync(0xffffffffu, hi1 ? ax : ay, 1), q1 = __shfl_xor_sync(0xffffffffu, hi1 ? az : aw, 1);
        const float x0 = (hi1 ? q0 : ax) * sc, x1 = (hi1 ? ay : q0) * sc, x2 = (hi1 ? q1 : az) * sc, x3 = (hi1 ? aw : q1) * sc;
        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn((x0 - f01.x) * kLoScale, (x1 - f01.y) * kLoScale);
        const __half2 l23 = __floats2half2_rn((x2 - f23.x) * kLoScale, (x3 - f23.y) * kLoScale);
        w1[s].x = *reinterpret_cast<const unsigned *>(&h01);
        w1[s].y = *reinterpret_cast<const unsigned *>(&h23);
        w2[s].x = *reinterpret_cast<const unsigned *>(&l01);
        w2[s].y = *reinterpret_cast<const unsigned *>(&l23);
    }
    // Lane L now holds cell L of the warp's 32 cells (4*ci + cl == lane): a 128-byte row per plane.  Written directly that
    // is 32 different lines per store instruction (measured: the whole kernel 88 us); instead the rows pass through a
    // warp-private staging tile (pitch 144 B: conflict-free both ways, __syncwarp only) and leave as 4 cells x 128
    // contiguous bytes per instruction.
    __shared__ uint4 s_stage[kRsThreads / 32][32 * 9];
    uint4 *stg = s_stage[warp];
    const int pr0 = lane >> 3, piece = lane & 7;
    const size_t cell0 = (size_t)b * P + 4 * (size_t)(q4b + warp * 8);
#pragma unroll
    for (int plane = 0; plane < 2; ++plane) {
        uint4 *dst = plane == 0 ? p1 : p2;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            stg[lane * 9 + j] = plane == 0 ? make_uint4(w1[2 * j].x, w1[2 * j].y, w1[2 * j + 1].x, w1[2 * j + 1].y)
                                           : make_uint4(w2[2 * j].x, w2[2 * j].y, w2[2 * j + 1].x, w2[2 * j + 1].y);
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int pr = 4 * t + pr0;                                   // cell of the warp: column pr >> 2, pixel pr & 3
            if (q4b + warp * 8 + (pr >> 2) < q4e) dst[(((cell0 + pr) * cin + c0) >> 3) + piece] = stg[pr * 9 + piece];
        }
    }
}

__global__ void weight_absmax_kernel(const float *__restrict__ w, size_t n, PackedHeader *hdr) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(&hdr->amax_bits, __float_as_uint(m));
}

// (Cout,Cin,3,3) fp32 -> two fp16 matrices [2*npad][9*cin]:
//   mat  (1-CTA kernel): rows [0,npad) = w2 (scaled residual), rows [npad,2npad) = w1
//   mat2 (CTA-pair kernel), hr = output channels per CTA of the pair: rows [r*2hr, r*2hr+hr) = w1 of channels
//        [r*hr,(r+1)*hr), the next hr rows = w2 of the same channels, for pair rank r = 0,1 -- each CTA of a pair loads
//        its own contiguous 2*hr rows.  hr = npad/2, or 36 for Cout <= 72 (KITTI): 144 instead of 160 B rows per MMA.
__global__ void pack_weights_f16_kernel(const float *__restrict__ w, int cout, int cin, int npad, int hr, PackedHeader *hdr,
                                        __half *__restrict__ mat, __half *__restrict__ mat2) {
    const float s = pow2_scale_for(__uint_as_float(hdr->amax_bits));
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        hdr->scale = s;
        hdr->inv_scale = 1.f / s;
        hdr->npad = npad;
        hdr->cin = cin;
        hdr->hr = hr;
    }
    const size_t ktot = (size_t)9 * cin, total = (size_t)npad * ktot;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / ktot);
        const size_t k = i % ktot;
        const int tap = (int)(k / cin), c = (int)(k % cin);
        float v = 0.f;
        if (n < cout) v = w[((size_t)n * cin + c) * 9 + tap];
        __half h1, h2;
        split_f16(v * s, h1, h2);
        mat[i] = h2;
        mat[total + i] = h1;
        if (n < 2 * hr) {   // channels >= 2*hr are padding beyond Cout and have no row in the pair layout
            const int r = n / hr, j = n - r * hr;
            mat2[((size_t)(r * 2 * hr + j)) * ktot + k] = h1;
            mat2[((size_t)(r * 2 * hr + hr + j)) * ktot + k] = h2;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------
// cute::UMMA::InstrDescriptor: c=F32 (1 @bit 4), a=b=F16 (0 @bits 7,10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, M=128, kind::f16
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }  // the 4 accumulate warps

struct F16Params {
    int cin, gh, gw, cout;
    int tiles_x, tiles_per_img, total_tiles;
    int upt;            // units per tile = (cin/64) * 3
    int units_per_cta;  // even cut of total_tiles*upt over the grid (>= upt)
    int a_stages, b_stages;
    int dbg;            // debug (SQD_F16_DBG): 1 = skip MMA issue, 2 = skip A loads, 4 = skip B loads
    const float *bias;
    const unsigned *amax_bits;   // (B, Cin/64) max|x| per image and channel block, fp32 bits
    const PackedHeader *whdr;
    float *pred;
    float *partial;  // (grid, 128, NPAD) partial sums of split tiles
    int *flags;      // (grid) 1 = partial[cta] published
    int *status;     // 0 ok; else the role whose bounded wait timed out
    long long *trace;  // debug (SQD_F16_TRACE = device address): per-unit clock64 stamps of CTA trace_cta, 32 slots per unit
    int trace_cta;
};

struct Sched {  // the permuted unit sequence of one CTA: [whole tiles + head segment][deferred tail segment]
    long long u0;
    int n, main_len, upt;
    int blk = 3;   // units per scale block: 3 (one 64-channel block); upt when every block of an image shares one scale
    __device__ __forceinline__ long long unit(int i) const {
        const int len_tail = n - main_len;
        return i < main_len ? u0 + len_tail + i : u0 + (i - main_len);
    }
    // unit i closes its accumulation chunk: the chunk is full, or the unit ends a segment of the sequence
    // (end of a tile, end of the main part, end of the CTA's range).  in_chunk = units already in the chunk.
    // A chunk never crosses a 64-channel block (r = cb*3 + dx): the block's feature scale is applied per chunk.
    __device__ __forceinline__ bool chunk_ends(int i, int r, int in_chunk, int chunk_units) const {
        // blk is 3 or upt (r < upt): keep the modulo a compile-time constant
        return in_chunk + 1 >= chunk_units || (blk == 3 ? r % 3 == 2 : r == blk - 1) || i == n - 1 || i == main_len - 1 ||
               r == upt - 1;
    }
};

// Walks the unit sequence of a CTA without per-unit divisions: (tile, r) are decoded from the linear unit index only
// at the two points where the sequence jumps (i == 0 and i == main_len) and stepped incrementally otherwise.
struct UnitIter {
    int tile, r;          // tile index, unit index inside the tile (r = cb*3 + dxi)
    int cb, dxi;          // 64-channel block, dx tap
    int img, tx, ty;      // image, tile column / row inside the image
    __device__ __forceinline__ void seek(long long u, const F16Params &p);
    __device__ __forceinline__ void next(const F16Params &p);
};

#ifdef SQD_ENABLE_TRACE
#define SQD_TRACE(slot, i) \
    do { if (p.trace && cta == p.trace_cta && lane == 0 && (i) < 512) p.trace[(i) * 32 + (slot)] = clock64(); } while (0)
#else
#define SQD_TRACE(slot, i) do { } while (0)
#endif

struct Ring {  // stage index + phase bit of a circular buffer of n stages
    int s;
    uint32_t ph;
    __device__ __forceinline__ void advance(int n) {
        if (++s == n) {
            s = 0;
            ph ^= 1u;
        }
    }
};

__device__ __forceinline__ void UnitIter::seek(long long u, const F16Params &p) {
    tile = (int)(u / p.upt);
    r = (int)(u - (long long)tile * p.upt);
    cb = r / 3;
    dxi = r - cb * 3;
    img = tile / p.tiles_per_img;
    const int t = tile - img * p.tiles_per_img;
    ty = t / p.tiles_x;
    tx = t - ty * p.tiles_x;
}
__device__ __forceinline__ void UnitIter::next(const F16Params &p) {
    ++r;
    if (++dxi == 3) {
        dxi = 0;
        ++cb;
    }
    if (r == p.upt) {
        r = 0;
        cb = 0;
        ++tile;
        if (++tx == p.tiles_x) {
            tx = 0;
            if (++ty * p.tiles_x == p.tiles_per_img) {
                ty = 0;
                ++img;
            }
        }
    }
}

template <int NPAD>
__global__ void __launch_bounds__(kThreads, 1)
convdet_f16_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_a2,
                   const __grid_constant__ CUtensorMap map_b, const F16Params p) {
    constexpr int kBStageBytes = 2 * NPAD * kBlockK * 2;  // [w2 rows | w1 rows] of one tap: one K-major tile of 2*NPAD rows
    constexpr int kW1Offset = NPAD * kBlockK * 2;         // byte offset of the w1 rows inside a B stage (multiple of 2048)
    constexpr int kAccCols = 2 * NPAD;                    // [cross | main]
    constexpr uint32_t kTmemCols = 512;
    constexpr uint32_t kIdescCat = umma_idesc_f16(128, 2 * NPAD);
    constexpr uint32_t kIdescOne = umma_idesc_f16(128, NPAD);
    static_assert(2 * kAccCols <= 512, "TMEM budget");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int AS = p.a_stages, BS = p.b_stages;
    uint8_t *a_ring = smem;
    uint8_t *b_ring = smem + (size_t)AS * kAStageBytes;
    uint8_t *ctrl = b_ring + (size_t)BS * kBStageBytes;
    uint64_t *a_full = reinterpret_cast<uint64_t *>(ctrl);    // [4]
    uint64_t *a_empty = a_full + 4;                           // [4]
    uint64_t *b_full = a_empty + 4;                           // [8]
    uint64_t *b_empty = b_full + 8;                           // [8]
    uint64_t *tmem_full = b_empty + 8;                        // [2]
    uint64_t *tmem_empty = tmem_full + 2;                     // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    volatile int *abort_flag = reinterpret_cast<volatile int *>(tmem_slot + 1);
    float *s_bias = reinterpret_cast<float *>(tmem_slot + 2);  // NPAD floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;

    // ---- this CTA's slice of the unit space --------------------------------------------------------------
    const long long total_units = (long long)p.total_tiles * p.upt;
    Sched sc;
    sc.upt = p.upt;
    sc.u0 = (long long)cta * p.units_per_cta;
    {
        long long u1 = sc.u0 + p.units_per_cta;
        if (u1 > total_units) u1 = total_units;
        sc.n = u1 > sc.u0 ? (int)(u1 - sc.u0) : 0;
        const int r0 = (int)(sc.u0 % p.upt);
        int len_tail = r0 ? p.upt - r0 : 0;   // the range starts inside a tile: that tail segment is done last
        if (len_tail > sc.n) len_tail = sc.n;
        sc.main_len = sc.n - len_tail;
    }
    const int n_units = sc.n;

    if (threadIdx.x == 0) {
        *abort_flag = 0;
        for (int s = 0; s < AS; ++s) {
            mbar_init(a_full + s, 1);
            mbar_init(a_empty + s, 1);
        }
        for (int s = 0; s < BS; ++s) {
            mbar_init(b_full + s, 1);
            mbar_init(b_empty + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tmem_full + b, 1);
            mbar_init(tmem_empty + b, 4);  // one arrival per accumulate warp
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if ((warp == kWarpATma || warp == kWarpBTma) && lane == 0) {
        tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_a2);
        tma_prefetch_desc(&map_b);
    }
    for (int i = threadIdx.x; i < NPAD; i += kThreads) s_bias[i] = i < p.cout ? __ldg(p.bias + i) : 0.f;
    if (warp == kWarpMma) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kWarpATma) {
        // ===== A producer: the x1 and x2 patches of one unit (warp stays converged, one elected lane issues) =====
        UnitIter it;
        Ring ra{0, 0};
        for (int i = 0; i < n_units; ++i) {
            if (i == 0 || i == sc.main_len) it.seek(sc.unit(i), p); else it.next(p);
            if (!mbar_wait_warp(a_empty + ra.s, ra.ph ^ 1u, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 1);
                break;
            }
            SQD_TRACE(0, i);
            if (elect_one_sync()) {
                uint8_t *st = a_ring + (size_t)ra.s * kAStageBytes;
                const int x = it.tx * kTileX + it.dxi - 1, y = it.ty * kTileY - 1;
                if (p.dbg & 2) {
                    mbar_arrive(a_full + ra.s);
                } else {
                    mbar_arrive_expect_tx(a_full + ra.s, kAStageBytes);
                    tma_load_4d(&map_a1, a_full + ra.s, st, it.cb * kBlockK, x, y, it.img);
                    tma_load_4d(&map_a2, a_full + ra.s, st + kPlaneBytes, it.cb * kBlockK, x, y, it.img);
                }
            }
            __syncwarp();
            ra.advance(AS);
        }
    } else if (warp == kWarpBTma) {
        // ===== B producer: the [w2 | w1] tile of one tap per step =====
        UnitIter it;
        Ring rb{0, 0};
        bool ok = true;
        for (int i = 0; i < n_units && ok; ++i) {
            if (i == 0 || i == sc.main_len) it.seek(sc.unit(i), p); else it.next(p);
            for (int dyi = 0; dyi < 3; ++dyi) {
                if (!mbar_wait_warp(b_empty + rb.s, rb.ph ^ 1u, abort_flag)) {
                    if (lane == 0) atomicCAS(p.status, 0, 5);
                    ok = false;
                    break;
                }
                if (dyi == 0) SQD_TRACE(1, i);
                const int tap = dyi * 3 + it.dxi;
                if (elect_one_sync()) {
                    if (p.dbg & 4) {
                        mbar_arrive(b_full + rb.s);
                    } else {
                        mbar_arrive_expect_tx(b_full + rb.s, kBStageBytes);
                        tma_load_2d(&map_b, b_full + rb.s, b_ring + (size_t)rb.s * kBStageBytes, tap * p.cin + it.cb * kBlockK, 0);
                    }
                }
                __syncwarp();
                rb.advance(BS);
            }
        }
    } else if (warp == kWarpMma) {
        // ===== MMA issuer: 24 SS-mode MMAs per unit into a fresh TMEM accumulator.  The warp stays converged and
        // one elected lane issues, so descriptors live in uniform registers. =====
        Ring ra{0, 0}, rb{0, 0};
        bool ok = true;
        for (int i = 0; i < n_units && ok; ++i) {
            const int buf = i & 1;
            const uint32_t acc_ph = (uint32_t)(i >> 1) & 1u;
            if (!mbar_wait_warp(tmem_empty + buf, acc_ph ^ 1u, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 4);
                break;
            }
            SQD_TRACE(2, i);
            if (!mbar_wait_warp(a_full + ra.s, ra.ph, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 2);
                break;
            }
            SQD_TRACE(3, i);
            const uint32_t d_tmem = tmem_base + (uint32_t)buf * kAccCols;
            const uint32_t a_addr = smem_u32(a_ring + (size_t)ra.s * kAStageBytes);
            for (int dyi = 0; dyi < 3; ++dyi) {
                if (!mbar_wait_warp(b_full + rb.s, rb.ph, abort_flag)) {
                    if (lane == 0) atomicCAS(p.status, 0, 6);
                    ok = false;
                    break;
                }
                SQD_TRACE(4 + 2 * dyi, i);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(b_ring + (size_t)rb.s * kBStageBytes);
                const uint64_t b_cat = umma_desc_sw128(b_addr);              // 2*NPAD rows: w2 then w1
                const uint64_t b_w1 = umma_desc_sw128(b_addr + kW1Offset);   // NPAD rows of w1
                const uint64_t a1 = umma_desc_sw128(a_addr + dyi * kDyBytes);
                const uint64_t a2 = umma_desc_sw128(a_addr + kPlaneBytes + dyi * kDyBytes);
                if (elect_one_sync()) {
                    if (!(p.dbg & 1)) {
#pragma unroll
                        for (int ks = 0; ks < kBlockK / kUmmaK; ++ks) {
                            const uint64_t adv = (uint64_t)((ks * kUmmaK * 2) >> 4);  // +32 B per K step, in 16 B units
                            umma_f16_ss(d_tmem, a1 + adv, b_cat + adv, kIdescCat, (dyi | ks) ? 1u : 0u);
                            umma_f16_ss(d_tmem, a2 + adv, b_w1 + adv, kIdescOne, 1u);
                        }
                    }
                    umma_commit(b_empty + rb.s);  // weight slot reusable once these MMAs have read it
                }
                __syncwarp();
                SQD_TRACE(5 + 2 * dyi, i);
                rb.advance(BS);
            }
            if (elect_one_sync()) {
                umma_commit(a_empty + ra.s);    // patch slot reusable
                umma_commit(tmem_full + buf);   // unit complete (also fires after an aborted tap loop)
            }
            __syncwarp();
            SQD_TRACE(10, i);
            ra.advance(AS);
        }
    } else {
        // ===== accumulate + epilogue warps: TMEM unit -> fp32 registers (RN) ... -> scale, +bias -> pred =====
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // accumulator row == cell inside the 8x16 tile
        const int et = threadIdx.x - kWarpAcc0 * 32;  // 0..127
        const float inv_sw = p.whdr->inv_scale;
        float acc[NPAD];
        int seg_r0 = 0;
        UnitIter it;
        for (int i = 0; i < n_units; ++i) {
            if (i == 0 || i == sc.main_len) it.seek(sc.unit(i), p); else it.next(p);
            const int r = it.r;
            if (i == 0 || r == 0 || i == sc.main_len) {
                seg_r0 = r;
#pragma unroll
                for (int n = 0; n < NPAD; ++n) acc[n] = 0.f;
            }
            const int buf = i & 1;
            const uint32_t acc_ph = (uint32_t)(i >> 1) & 1u;
            if (!mbar_wait(tmem_full + buf, acc_ph, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 3);
                break;
            }
            tc_fence_after();
            __syncwarp();
            if (warp == kWarpAcc0) SQD_TRACE(11, i);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * kAccCols;
            // 1 / (feature scale of this image and channel block): a power of two, so the fma below rounds once,
            // exactly like an fp32 add of the unscaled unit
            const float inv_a = 1.f / pow2_scale_for(__uint_as_float(__ldg(p.amax_bits + (size_t)it.img * (p.cin / kBlockK) + it.cb)));
#pragma unroll
            for (int n0 = 0; n0 < NPAD; n0 += 16) {
                uint32_t v[16], w[16];
                tmem_ld_x16(taddr + n0, v);          // a1*w2 + a2*w1   (x 2^11)
                tmem_ld_x16(taddr + NPAD + n0, w);   // a1*w1
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    acc[n0 + k] = fmaf(fmaf(__uint_as_float(v[k]), kLoInv, __uint_as_float(w[k])), inv_a, acc[n0 + k]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + buf);  // this warp is done reading the accumulator
            if (warp == kWarpAcc0) SQD_TRACE(12, i);

            const bool seg_end = (i == n_units - 1) || (r == p.upt - 1) || (i == sc.main_len - 1);
            if (!seg_end) continue;
            const bool from_start = seg_r0 == 0, to_end = r == p.upt - 1;
            if (from_start && !to_end) {
                // head of a split tile: publish the partial sums for the next CTA (which holds the tail)
                float4 *dst = reinterpret_cast<float4 *>(p.partial + ((size_t)cta * 128 + row) * NPAD);
#pragma unroll
                for (int n = 0; n < NPAD; n += 4) dst[n >> 2] = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
                __threadfence();
                epi_bar();
                if (et == 0) st_release(p.flags + cta, 1);
                continue;
            }
            if (!from_start && to_end) {
                // tail of a split tile (processed last): add the head published by the previous CTA, fixed order
                if (et == 0) {
                    unsigned spin = 0;
                    while (ld_acquire(p.flags + cta - 1) == 0) {
                        if (++spin > kSpinLimit || *abort_flag) {
                            *abort_flag = 1;
                            atomicCAS(p.status, 0, 8);
                            break;
                        }
                    }
                }
                epi_bar();  // (on abort keep going: every later wait fails for all four warps at the same unit)
                const float4 *src = reinterpret_cast<const float4 *>(p.partial + ((size_t)(cta - 1) * 128 + row) * NPAD);
#pragma unroll
                for (int n = 0; n < NPAD; n += 4) {
                    const float4 h = __ldcg(src + (n >> 2));
                    acc[n] = fadd(h.x, acc[n]); acc[n + 1] = fadd(h.y, acc[n + 1]);
                    acc[n + 2] = fadd(h.z, acc[n + 2]); acc[n + 3] = fadd(h.w, acc[n + 3]);
                }
            } else if (!(from_start && to_end)) {
                if (lane == 0) atomicCAS(p.status, 0, 9);  // a segment strictly inside a tile: scheduler invariant broken
                continue;
            }
            // whole tile in registers: x 1/(s_a*s_w), + bias -> pred
            const int x = it.tx * kTileX + row % kTileX, y = it.ty * kTileY + row / kTileX;
            const float inv = inv_sw;   // the feature scales were divided out chunk by chunk
            if (y < p.gh && x < p.gw) {
                float *out = p.pred + (((size_t)it.img * p.gh + y) * p.gw + x) * p.cout;
                if ((p.cout & 3) == 0) {
                    float4 *o4 = reinterpret_cast<float4 *>(out);
#pragma unroll
                    for (int n = 0; n < NPAD; n += 4)
                        if (n < p.cout)
                            o4[n >> 2] = make_float4(fadd(fmul(acc[n], inv), s_bias[n]), fadd(fmul(acc[n + 1], inv), s_bias[n + 1]),
                                                     fadd(fmul(acc[n + 2], inv), s_bias[n + 2]),
                                                     fadd(fmul(acc[n + 3], inv), s_bias[n + 3]));
                } else {
#pragma unroll
                    for (int n = 0; n < NPAD; ++n)
                        if (n < p.cout) out[n] = fadd(fmul(acc[n], inv), s_bias[n]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kWarpMma) {
        __syncwarp();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair kernel (cta_group::2): two SMs of a TPC work on two M tiles with ONE issuing thread and ONE copy of B.
// ---------------------------------------------------------------------------------------------------
// Why: the 1-CTA kernel is bound by (1) L2->SM traffic -- every CTA streams the whole 2.2 MB weight matrix per tile --
// and (2) the single MMA-issuing thread (barrier round trips + ~46 cycles per tcgen05.mma).  With cta_group::2 an MMA has
// M = 256 (128 rows per CTA), each CTA holds only HALF of the B rows (the tensor core reads the other half from the
// peer's shared memory), and one thread issues for both SMs: B traffic per tile and issue cost per FLOP both halve.
//
//  * pair-tile j = tiles {j, j + PT} (PT = ceil(tiles/2)); rank r of the pair owns tile j + r*PT.  A ghost tile
//    (odd tile count) loads zeros (TMA out-of-bounds image index) and is never stored.
//  * B tile of one tap in CTA r (HR = output channels per CTA): rows [0,HR) = w1 of channels [r*HR,(r+1)*HR), rows
//    [HR,2HR) = w2 (the scaled residual) of the same channels.
//        MMA1: D[:, 0:4HR)      (+)= A1 * B^T           -> columns [main(r=0) | cross(r=0) | main(r=1) | cross(r=1)]
//        MMA2: D[:, 4HR:+2*N2H) (+)= A2 * B[0:N2H)^T    -> columns [cross2(r=0) | cross2(r=1)], N2H = round8(HR); its
//              columns HR..N2H-1 of each half are A2 * w2 products nobody reads
//    HR = Npad/2 in general; HR = 36 for Cout <= 72 (KITTI's 72 channels): N = 144 + 80 instead of 160 + 80 per K step.
//  * unit-granular stages: one stage = the unit's two A patches + its three B tiles (67 KB at Cout = 72, 3 stages).
//    Per unit the issuing thread does two barrier waits, 24 MMAs and two multicast commits:
//        full[s]   leader only: both CTAs' TMA bytes of stage s landed (cta_group::2 TMA credits the leader's barrier)
//        sfree[s]  both CTAs  : the MMAs reading stage s completed            (tcgen05.commit multicast)
//        tfull[b]  both CTAs  : TMEM accumulator b holds a finished unit      (tcgen05.commit multicast)
//        tempty[b] leader only: all 8 accumulate warps of the pair drained accumulator b (remote mbarrier arrives)
//  * 192 threads: warp 0 TMA producer, warp 1 TMEM alloc (+ MMA issue in the leader), warps 2..5 accumulate/epilogue.
constexpr size_t kSmemLimit = 227 * 1024;
constexpr size_t kCtrlBytes = 1024;
constexpr int kThreads2 = 192;
constexpr int kEpiPitch = 68;                                  // floats per staged pixel row: 64 channels + 4 (bank spread)
constexpr size_t kEpiStageBytes = (size_t)4 * 32 * kEpiPitch * 4;   // PK kernel only: one staging tile per accumulate warp
constexpr int kWarpTma2 = 0, kWarpMma2 = 1, kWarpAcc2 = 2;

__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

struct PairParams {
    int cin, gh, gw, cout, batch;
    int tiles_x, tiles_per_img, total_tiles, pair_tiles;
    int upt;             // units per tile = (cin/64) * 3
    int units_per_pair;  // even cut of pair_tiles*upt over the pairs (>= upt)
    int stages;
    int chunk_units;     // units accumulated inside TMEM before the sum moves to registers
    int chunk_blk;       // units that share one feature scale: 3 (a 64-channel block) or upt (one scale per image)
    int dbg;
    const float *bias;
    const unsigned *amax_bits;
    const PackedHeader *whdr;
    float *pred;
    float *partial;  // (grid, 128, NPAD)
    int *flags;      // (grid)
    int *status;
    long long *trace;
    int trace_cta;
    int out_stride;      // floats between consecutive cells of the output (== cout for pred; Cin for the dgrad slabs)
    int pdl;             // host only: launch as a programmatic dependent of the pre-pass kernel
    int a_bufs;          // AO kernels: operand buffers (whole (tile, block) patches) in front of the weight stages
    int ksteps_last;     // PK kernels: valid 16-channel K steps of the last channel block (1..4)
    // PK kernels (the dgrad GEMM): `slabs` weight matrices share the same input planes; the tile space is slab-major,
    // tiles_per_slab (even) tiles per slab; slab s uses the weights of map_b's 3rd coordinate s (header `slab_stride`
    // bytes further) and writes output columns [s*NPAD, +NPAD).
    int slabs, tiles_per_slab;
    size_t slab_stride;
    // fused score epilogue (template CS > 0 only): candidates above the score threshold go to per-image lists
    SqdCand cand;        // cand.count == nullptr: no emission
    float score_thr;
    int anchors_per_cell;
};

// like UnitIter, for the tile of one rank; img == batch marks a ghost tile (loads zeros, stores nothing).
//  PK = false: pair-tile j = tiles {j, j + pair_tiles}; `sel` = the rank's tile offset.
//  PK = true (several weight slabs over the same images): the two ranks of a pair MUST work on the same slab (they share
//  the B operand), so pair-tile j = the adjacent tiles {2j, 2j+1} of a slab-major tile space whose slabs are padded to
//  an even tile count (tiles_per_slab); `sel` = the rank.
template <bool PK>
struct PairIterT {
    int r, cb, dxi, img, tx, ty, slab, pt, sel;
    __device__ __forceinline__ void locate(const PairParams &p) {
        if (PK) {
            const int tile = 2 * pt + sel;
            slab = tile / p.tiles_per_slab;
            const int local = tile - slab * p.tiles_per_slab;
            if (slab >= p.slabs) {
                slab = p.slabs - 1;
                img = p.batch;
                tx = ty = 0;
                return;
            }
            img = local / p.tiles_per_img;       // == batch for the slab's padding tile
            const int t = local - img * p.tiles_per_img;
            ty = t / p.tiles_x;
            tx = t - ty * p.tiles_x;
        } else {
            slab = 0;
            const int tile = pt + sel;
            img = tile / p.tiles_per_img;
            const int t = tile - img * p.tiles_per_img;
            ty = t / p.tiles_x;
            tx = t - ty * p.tiles_x;
        }
    }
    __device__ __forceinline__ void seek(long long u, int sel_, const PairParams &p) {
        sel = sel_;
        pt = (int)(u / p.upt);
        r = (int)(u - (long long)pt * p.upt);
        cb = r / 3;
        dxi = r - cb * 3;
        locate(p);
    }
    __device__ __forceinline__ void next(const PairParams &p) {
        ++r;
        if (++dxi == 3) {
            dxi = 0;
            ++cb;
        }
        if (r == p.upt) {
            r = 0;
            cb = 0;
            ++pt;
            if (PK) {
                locate(p);
            } else if (++tx == p.tiles_x) {
                tx = 0;
                if (++ty * p.tiles_x == p.tiles_per_img) {
                    ty = 0;
                    ++img;
                }
            }
        }
    }
};

#ifdef SQD_ENABLE_TRACE
#define SQD_TRACE2(slot, i) \
    do { if (p.trace && cta == p.trace_cta && lane == 0 && (i) < 512) p.trace[(i) * 32 + (slot)] = clock64(); } while (0)
// kernel phases (thread 0 of the traced CTA): row 511 of the trace buffer
#define SQD_TRACE_PH(slot) \
    do { if (p.trace && blockIdx.x == (unsigned)p.trace_cta && threadIdx.x == 0) p.trace[511 * 32 + (slot)] = clock64(); } while (0)
#else
#define SQD_TRACE2(slot, i) do { } while (0)
#define SQD_TRACE_PH(slot) do { } while (0)
#endif

__device__ __forceinline__ bool pair_wait_warp(uint64_t *bar, uint32_t parity, volatile int *abort_flag, bool spin) {
    return spin ? mbar_spin_warp(bar, parity, abort_flag) : mbar_wait_warp(bar, parity, abort_flag);
}

// CS > 0: the epilogue also scores the cell's anchors (CS classes, NF = CS+5 fields per anchor) from the finished
// fp32 logits -- class softmax x confidence sigmoid, first-max argmax, the same sqd_score_anchor every filter kernel
// uses -- and appends the anchors above the score threshold to per-image candidate lists (SqdCand), so the filter
// that follows never scans pred.
// PK: the last 64-channel block is only partly filled (p.ksteps_last valid 16-channel K steps; the rest is zero padding,
// e.g. the 72 -> 128 padded gradient channels of the dgrad GEMM): its all-zero K steps are not issued.
// AO ("A once"): the operand patch of a (tile, channel block) is fetched ONCE instead of once per dx tap.  The TMA box is
// {64 ch, 10 y, 18 x} on a tensor map whose dimension order is (channel, y, x, image), so the patch lands x-major / y-minor:
// row = px*10 + py.  An 8-row swizzle atom is then one tile COLUMN (8 consecutive y of one x), atoms follow each other at a
// uniform 10 rows (SBO = 1280 B), and tap (dy, dx) is the same patch read from row dx*10 + dy on -- legal because the tensor
// core derives the 128-byte swizzle phase from absolute address bits (profiles/r01_umma_row_offset_microtest.txt), at the
// aligned rate (tools/micro/umma_rate_shift.cu).  M row m of a tile = cell (y0 + m % 8, x0 + m / 8).  46 KB per (tile, block)
// instead of 3 x 41 KB: the kernel is bound by the shared-memory port (profiles/r02_one_kernel_convdet.txt), and this takes
// 77 KB of TMA writes per block off it.  A ring of whole-block operand buffers + the usual ring of weight stages.
constexpr int kAoY = kTileY + 2, kAoX = kTileX + 2;
constexpr int kAoBoxBytes = kAoX * kAoY * kBlockK * 2;                     // 23040: what one TMA box delivers
constexpr int kAoPlaneBytes = (kAoBoxBytes + 1023) / 1024 * 1024;          // 23552: swizzle-atom aligned
constexpr int kAoBufBytes = 2 * kAoPlaneBytes;
__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

constexpr int kScorerWarp0 = kThreads2 / 32;                  // CS > 0: four scorer warps follow the six pipeline warps
// AO kernels: one more warp, the operand-patch producer (its waits for a free operand buffer must not hold up the
// weight-stage refills of the other producer warp)
__host__ __device__ constexpr int pair_threads(int cs, bool ao = false) { return kThreads2 + (cs > 0 ? 128 : 0) + (ao ? 32 : 0); }

template <int NPAD, int CS, bool PK = false, int HR = NPAD / 2, bool K3 = (NPAD >= 96), bool AO = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(pair_threads(CS, AO), 1)
convdet_f16_pair_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_a2,
                        const __grid_constant__ CUtensorMap map_b, const PairParams p) {
    // This CTA's half of one tap is [w1: HR rows | w2: HR rows] for its HR output channels.  MMA1 = A1 x all 2*HR rows
    // (N = 4*HR across the pair), MMA2 = A2 x the first N2H = round8(HR) rows (w1, plus up to 4 w2 rows whose columns
    // nobody reads): both B descriptors start at row 0, so HR only has to be a multiple of 4.
    constexpr int N1H = 2 * HR, N2H = (HR + 7) / 8 * 8;
    static_assert(HR % 4 == 0 && 2 * HR <= NPAD && N2H <= N1H, "pair layout");
    constexpr int kBTapBytes = N1H * kBlockK * 2;           // multiple of 1024: N1H is a multiple of 8
    static_assert(!AO || (CS == 0 && !PK), "the A-once layout is built for the plain forward kernel");
    constexpr int kBOff = AO ? 0 : kAStageBytes;            // the three weight tiles inside a stage
    constexpr int kStageBytes = kBOff + 3 * kBTapBytes;
    // K3 (Npad >= 96: the stress shape's 117 channels, the dgrad GEMM's 128): THREE MMAs of N = 2*HR per K step instead --
    // A1 x w1 -> main columns, A1 x w2 and A2 x w1 -> the SAME cross columns (both carry the 2^-11 scale).  Same tensor
    // time in the linear regime of the MMA rate (N >= 96: 128 + 64 = 3 x 64 cycles), but 4*HR accumulator columns instead
    // of 6*HR: two TMEM buffers fit and the drain of a chunk overlaps the MMAs of the next.  (For KITTI's N = 72 the three
    // MMAs would sit on the ~45-cycle instruction floor: 135 vs 116 cycles per K step; 2 x 224 columns fit anyway.)
    constexpr bool k3 = K3 && HR % 8 == 0;
    constexpr int kAccCols = k3 ? 4 * HR : 2 * N1H + 2 * N2H;   // [MMA1: 4*HR | MMA2: 2*N2H]  or  [main 2*HR | cross 2*HR]
    constexpr int kAccBufs = (2 * kAccCols <= 512) ? 2 : 1;
    constexpr uint32_t kTmemCols = 512;
    constexpr uint32_t kIdesc1 = umma_idesc_f16(256, 2 * N1H);
    constexpr uint32_t kIdesc2 = umma_idesc_f16(256, 2 * N2H);
    constexpr uint32_t kIdesc3 = umma_idesc_f16(256, 2 * HR);
    static_assert(kAccCols <= 512 && (2 * N1H) % 16 == 0 && (2 * N2H) % 16 == 0, "TMEM budget / UMMA N");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int S = p.stages;
    uint8_t *a_ring = smem;                                 // AO: p.a_bufs whole-block operand buffers in front of the stages
    if (AO) smem += (size_t)p.a_bufs * kAoBufBytes;
    uint8_t *ctrl = smem + (size_t)S * kStageBytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(ctrl);   // [4]
    uint64_t *sfree = full + 4;                            // [4]
    uint64_t *tfull = sfree + 4;                           // [2]
    uint64_t *tempty = tfull + 2;                          // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    volatile int *abort_flag = reinterpret_cast<volatile int *>(tmem_slot + 1);
    float *s_bias = reinterpret_cast<float *>(tmem_slot + 2);  // NPAD floats
    // CS > 0: accumulate warp q hands every finished tile to scorer warp q through a two-entry ring
    uint64_t *afull = reinterpret_cast<uint64_t *>(ctrl + 768);     // AO (never with CS > 0): [4] leader: both CTAs' patch landed
    uint64_t *afree = afull + 4;                                    //                         [4] both: its MMAs completed
    uint64_t *sc_full = reinterpret_cast<uint64_t *>(ctrl + 768);   // [4][2] tile record written, pred rows stored
    uint64_t *sc_empty = sc_full + 8;                               // [4][2] record consumed
    int *sc_rec = reinterpret_cast<int *>(sc_empty + 8);            // [4][2][4] {image (-1: no more tiles), tile x, tile y, -}

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const uint32_t rank = cluster_ctarank();
    const int pair = cta >> 1, npairs = gridDim.x >> 1;
    const int tile_offset = PK ? (int)rank : (rank ? p.pair_tiles : 0);
    using PairIter = PairIterT<PK>;
    const bool spin = (p.dbg & 8) != 0;   // debug: 8 = poll mbarrier.test_wait instead of parking in try_wait (slower)
    SQD_TRACE_PH(0);

    // ---- this pair's slice of the (pair-tile, unit) space ----------------------------------------------------
    const long long total_units = (long long)p.pair_tiles * p.upt;
    Sched sc;
    sc.upt = p.upt;
    sc.blk = p.chunk_blk;
    sc.u0 = (long long)pair * p.units_per_pair;
    {
        long long u1 = sc.u0 + p.units_per_pair;
        if (u1 > total_units) u1 = total_units;
        sc.n = u1 > sc.u0 ? (int)(u1 - sc.u0) : 0;
        const int r0 = (int)(sc.u0 % p.upt);
        int len_tail = r0 ? p.upt - r0 : 0;
        if (len_tail > sc.n) len_tail = sc.n;
        sc.main_len = sc.n - len_tail;
    }
    const int n_units = sc.n;
    (void)npairs;

    if (threadIdx.x == 0) {
        *abort_flag = 0;
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, 1);
            mbar_init(sfree + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull + b, 1);
            mbar_init(tempty + b, 8);  // 4 accumulate warps x 2 CTAs
        }
        if (CS > 0)
            for (int b = 0; b < 8; ++b) {
                mbar_init(sc_full + b, 1);
                mbar_init(sc_empty + b, 1);
            }
        if (AO)
            for (int b = 0; b < 4; ++b) {
                mbar_init(afull + b, 1);
                mbar_init(afree + b, 1);
            }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == kWarpTma2 && lane == 0) {
        tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_a2);
        tma_prefetch_desc(&map_b);
    }
    for (int i = threadIdx.x; i < NPAD; i += pair_threads(CS, AO)) s_bias[i] = (i < p.cout && p.bias) ? __ldg(p.bias + i) : 0.f;
    SQD_TRACE_PH(1);
    if (warp == kWarpMma2) tmem_alloc_2cta(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    SQD_TRACE_PH(2);
    cluster_sync_all();  // both CTAs: barriers initialised, TMEM allocated, before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    SQD_TRACE_PH(3);
    // Programmatic dependent launch: everything above ran while the pre-pass kernel was still finishing; from here on
    // its outputs (fp16 planes, block maxima) are needed.
    sqd_pdl_wait();
    SQD_TRACE_PH(4);

    if (warp == kWarpTma2) {
        // ===== producer: this CTA's A patches (own tile) and its half of the three B tiles of one unit =====
        PairIter it;
        Ring rs{0, 0};
        for (int i = 0; i < n_units; ++i) {
            if (i == 0 || i == sc.main_len) it.seek(sc.unit(i), tile_offset, p); else it.next(p);
            if (!pair_wait_warp(sfree + rs.s, rs.ph ^ 1u, abort_flag, spin)) {
                if (lane == 0) atomicCAS(p.status, 0, 1);
                break;
            }
            SQD_TRACE2(0, i);
            if (elect_one_sync()) {
                uint8_t *st = smem + (size_t)rs.s * kStageBytes;
                const int x = it.tx * kTileX + it.dxi - 1, y = it.ty * kTileY - 1;
                const int bytes = ((p.dbg & 2) || AO ? 0 : kAStageBytes) + ((p.dbg & 4) ? 0 : 3 * kBTapBytes);
                if (rank == 0) mbar_arrive_expect_tx(full + rs.s, 2 * bytes);  // both CTAs' bytes
                if (!AO && !(p.dbg & 2)) {   // ghost tile: image index == batch -> out of bounds -> zeros
                    tma_load_4d_2cta(&map_a1, full + rs.s, st, it.cb * kBlockK, x, y, it.img);
                    tma_load_4d_2cta(&map_a2, full + rs.s, st + kPlaneBytes, it.cb * kBlockK, x, y, it.img);
                }
                if (!(p.dbg & 4)) {
                    const int slab = it.slab;
#pragma unroll
                    for (int dyi = 0; dyi < 3; ++dyi) {
                        if (PK)
                            tma_load_3d_2cta(&map_b, full + rs.s, st + kBOff + dyi * kBTapBytes,
                                             (dyi * 3 + it.dxi) * p.cin + it.cb * kBlockK, (int)rank * N1H, slab);
                        else
                            tma_load_2d_2cta(&map_b, full + rs.s, st + kBOff + dyi * kBTapBytes,
                                             (dyi * 3 + it.dxi) * p.cin + it.cb * kBlockK, (int)rank * N1H);
                    }
                }
            }
            __syncwarp();
            rs.advance(S);
        }
    } else if (AO && warp == kScorerWarp0) {
        // ===== AO operand producer: the patch of this CTA's tile, fetched by the first unit of every (tile, block) group
        // of the pair's sequence (a group is cut where the sequence is: at its start, its end and at the main / tail seam) =====
        PairIter it;
        Ring ra{0, 0};
        for (int i = 0; i < n_units; ++i) {
            if (i == 0 || i == sc.main_len) it.seek(sc.unit(i), tile_offset, p); else it.next(p);
            if (!(it.dxi == 0 || i == 0 || i == sc.main_len)) continue;
            if (!pair_wait_warp(afree + ra.s, ra.ph ^ 1u, abort_flag, spin)) {
                if (lane == 0) atomicCAS(p.status, 0, 12);
                break;
            }
            if (elect_one_sync()) {
                uint8_t *ab = a_ring + (size_t)ra.s * kAoBufBytes;
                if (rank == 0) mbar_arrive_expect_tx(afull + ra.s, (p.dbg & 2) ? 0 : 4 * kAoBoxBytes);   // 2 planes x 2 CTAs
                if (!(p.dbg & 2)) {   // map dimension order (channel, y, x, image); out of bounds (padding, ghost tile) -> zeros
                    const int x = it.tx * kTileX - 1, y = it.ty * kTileY - 1;
                    tma_load_4d_2cta(&map_a1, afull + ra.s, ab, it.cb * kBlockK, y, x, it.img);
                    tma_load_4d_2cta(&map_a2, afull + ra.s, ab + kAoPlaneBytes, it.cb * kBlockK, y, x, it.img);
                }
            }
            __syncwarp();
            ra.advance(p.a_bufs);
        }
    } else if (warp == kWarpMma2) {
        if (rank == 0) {
            // ===== MMA issuer (leader CTA): 24 M=256 MMAs per unit; converged warp, one elected lane issues =====
            Ring rs{0, 0}, ra{0, 0};
            int chunk = 0, in_chunk = 0;  // a chunk = up to p.chunk_units consecutive units of one segment in one accumulator
            PairIter it;
            for (int i = 0; i < n_units; ++i) {
                if (i == 0 || i == sc.main_len) it.seek(sc.unit(i), tile_offset, p); else it.next(p);
                const bool a_first = AO && (it.dxi == 0 || i == 0 || i == sc.main_len);
                const bool a_last = AO && (it.dxi == 2 || i == n_units - 1 || i == sc.main_len - 1);
                const int buf = kAccBufs == 2 ? (chunk & 1) : 0;
                const uint32_t acc_ph = kAccBufs == 2 ? ((uint32_t)(chunk >> 1) & 1u) : ((uint32_t)chunk & 1u);
                if (in_chunk == 0 && !(p.dbg & 128)) {   // dbg 128 / 64: timing experiments (skip the accumulator / operand waits)
                    if (!pair_wait_warp(tempty + buf, acc_ph ^ 1u, abort_flag, spin)) {
                        if (lane == 0) atomicCAS(p.status, 0, 4);
                        break;
                    }
                }
                SQD_TRACE2(2, i);
                if (a_first && !(p.dbg & 64) && !pair_wait_warp(afull + ra.s, ra.ph, abort_flag, spin)) {
                    if (lane == 0) atomicCAS(p.status, 0, 13);
                    break;
                }
                if (!(p.dbg & 64) && !pair_wait_warp(full + rs.s, rs.ph, abort_flag, spin)) {
                    if (lane == 0) atomicCAS(p.status, 0, 2);
                    break;
                }
                SQD_TRACE2(3, i);
                tc_fence_after();
                const uint32_t d1 = tmem_base + (uint32_t)buf * kAccCols, d2 = d1 + 2 * N1H;
                const uint32_t st = smem_u32(smem + (size_t)rs.s * kStageBytes);
                const uint32_t ab = smem_u32(a_ring + (size_t)ra.s * kAoBufBytes) + (uint32_t)(it.dxi * kAoY) * 128u;   // AO: tap column
                const bool chunk_end = sc.chunk_ends(i, it.r, in_chunk, p.chunk_units);
                const int nks = (PK && it.cb == p.cin / kBlockK - 1) ? p.ksteps_last : kBlockK / kUmmaK;
                if (elect_one_sync()) {
                    if (!(p.dbg & 1)) {
#pragma unroll
                        for (int dyi = 0; dyi < 3; ++dyi) {
                            // dbg 32 (timing experiment, wrong results): atom-aligned taps and SBO = 1024 on the AO buffers
                            const uint32_t ab_t = (p.dbg & 32) ? (ab & ~1023u) + dyi * 1024 : ab + dyi * 128;
                            const uint32_t sbo_t = (p.dbg & 32) ? 1024 : kAoY * 128;
                            const uint64_t a1 = AO ? umma_desc_sw128_sbo(ab_t, sbo_t) : umma_desc_sw128(st + dyi * kDyBytes);
                            const uint64_t a2 = AO ? umma_desc_sw128_sbo(ab_t + kAoPlaneBytes, sbo_t)
                                                   : umma_desc_sw128(st + kPlaneBytes + dyi * kDyBytes);
                            const uint64_t b = umma_desc_sw128(st + kBOff + dyi * kBTapBytes);
                            const uint64_t bw1 = b;   // the w1 rows come first
                            const uint64_t bw2 = umma_desc_sw128(st + kBOff + dyi * kBTapBytes + HR * kBlockK * 2);
#pragma unroll
                            for (int ks = 0; ks < kBlockK / kUmmaK; ++ks) {
                                if (PK && ks >= nks) continue;
                                const uint64_t adv = (uint64_t)((ks * kUmmaK * 2) >> 4);
                                const uint32_t accum = (in_chunk | dyi | ks) ? 1u : 0u;
                                if (k3) {
                                    umma_f16_ss_2cta(d1, a1 + adv, bw1 + adv, kIdesc3, accum);            // main  = a1*w1
                                    umma_f16_ss_2cta(d1 + 2 * HR, a1 + adv, bw2 + adv, kIdesc3, accum);   // cross = a1*w2
                                    umma_f16_ss_2cta(d1 + 2 * HR, a2 + adv, bw1 + adv, kIdesc3, 1u);      //       + a2*w1
                                } else {
                                    umma_f16_ss_2cta(d1, a1 + adv, b + adv, kIdesc1, accum);
                                    umma_f16_ss_2cta(d2, a2 + adv, bw1 + adv, kIdesc2, accum);
                                }
                            }
                        }
                    }
                    umma_commit_2cta(sfree + rs.s, 3);                  // stage reusable in both CTAs
                    if (a_last) umma_commit_2cta(afree + ra.s, 3);      // operand buffer reusable in both CTAs
                    if (chunk_end) umma_commit_2cta(tfull + buf, 3);    // chunk complete (both CTAs' accumulate warps)
                }
                __syncwarp();
                SQD_TRACE2(10, i);
                rs.advance(S);
                if (a_last) ra.advance(p.a_bufs);
                if (chunk_end) {
                    ++chunk;
                    in_chunk = 0;
                } else {
                    ++in_chunk;
                }
            }
        }
        // All MMAs of this pair are issued: let the kernel behind us (the scan) be scheduled while the last chunks drain
        // and the epilogue runs.  (Triggering at kernel start parks its CTAs next to ours for the whole GEMM: +5 %.)
        sqd_pdl_trigger();
    } else if (warp < kScorerWarp0) {
        // ===== accumulate + epilogue warps (both CTAs, each on its own 128 TMEM lanes) =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - kWarpAcc2 * 32;  // 0..127
        const float inv_sw = p.whdr->inv_scale;
        float acc[NPAD];
        int seg_r0 = 0, chunk = 0, in_chunk = 0;
        PairIter it;
        // ---- fused score epilogue (CS > 0 and a candidate sink only): every finished tile is handed to scorer warp q, which
        // scores the cells this warp just stored while this warp goes on draining TMEM (see the scorer branch below)
        const bool emit = CS > 0 && p.cand.count != nullptr;
        int n_handed = 0;
        auto hand_over = [&](int img, int tx, int ty) {
            const int slot = n_handed & 1;
            const uint32_t ph = (uint32_t)(n_handed >> 1) & 1u;
            if (!mbar_wait_warp(sc_empty + q * 2 + slot, ph ^ 1u, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 10);
                return;
            }
            __threadfence();       // this lane's pred stores are visible before the record is
            __syncwarp();
            if (lane == 0) {
                int *rec = sc_rec + (q * 2 + slot) * 4;
                rec[0] = img; rec[1] = tx; rec[2] = ty;
                mbar_arrive(sc_full + q * 2 + slot);
            }
            ++n_handed;
        };
        for (int i = 0; i < n_units; ++i) {
            if (i == 0 || i == sc.main_len) it.seek(sc.unit(i), tile_offset, p); else it.next(p);
            const int r = it.r;
            if (i == 0 || r == 0 || i == sc.main_len) {
                seg_r0 = r;
#pragma unroll
                for (int n = 0; n < NPAD; ++n) acc[n] = 0.f;
            }
            const bool chunk_end = sc.chunk_ends(i, r, in_chunk, p.chunk_units);
            if (!chunk_end) {   // the tensor core keeps accumulating this chunk in TMEM
                ++in_chunk;
                continue;
            }
            in_chunk = 0;
            const int buf = kAccBufs == 2 ? (chunk & 1) : 0;
            const uint32_t acc_ph = kAccBufs == 2 ? ((uint32_t)(chunk >> 1) & 1u) : ((uint32_t)chunk & 1u);
            ++chunk;
            if (!pair_wait_warp(tfull + buf, acc_ph, abort_flag, spin)) {
                if (lane == 0) atomicCAS(p.status, 0, 3);
                break;
            }
            tc_fence_after();
            __syncwarp();
            if (warp == kWarpAcc2) SQD_TRACE2(11, i);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * kAccCols;
            // 1 / (feature scale of this image and channel block); all units of a chunk share the block.  A power of
            // two, so the fma below rounds once, exactly like an fp32 add of the unscaled chunk.  Ghost tile: any value.
            const float inv_a = it.img < p.batch
                ? 1.f / pow2_scale_for(__uint_as_float(__ldg(p.amax_bits + (size_t)it.img * (p.cin / kBlockK) + it.cb)))
                : 1.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int j0 = 0; j0 < HR; j0 += 8) {
                    if (k3) {
                        uint32_t mn[8], cr[8];
                        tmem_ld_x8(taddr + h * HR + j0, mn);                 // a1*w1
                        tmem_ld_x8(taddr + 2 * HR + h * HR + j0, cr);        // a1*w2 + a2*w1
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            acc[h * HR + j0 + k] = fmaf(fmaf(__uint_as_float(cr[k]), kLoInv, __uint_as_float(mn[k])), inv_a,
                                                        acc[h * HR + j0 + k]);
                        continue;
                    }
                    uint32_t c1[8], mn[8], c2[8];
                    tmem_ld_x8(taddr + h * N1H + j0, mn);              // a1*w1
                    tmem_ld_x8(taddr + h * N1H + HR + j0, c1);         // a1*w2
                    tmem_ld_x8(taddr + 2 * N1H + h * N2H + j0, c2);    // a2*w1
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (j0 + k < HR) {   // HR % 8 == 4: the last group's upper half belongs to the neighbouring block
                            const float cross = fadd(__uint_as_float(c1[k]), __uint_as_float(c2[k]));
                            acc[h * HR + j0 + k] = fmaf(fmaf(cross, kLoInv, __uint_as_float(mn[k])), inv_a, acc[h * HR + j0 + k]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(tempty + buf);
                else if (p.dbg & 16) mbar_arrive_cluster(tempty + buf, 0);
                else mbar_arrive_remote(tempty + buf, 0);   // TMEM reads are ordered by wait::ld + the tcgen05 fence above
            }
            if (warp == kWarpAcc2) SQD_TRACE2(12, i);

            const bool seg_end = (i == n_units - 1) || (r == p.upt - 1) || (i == sc.main_len - 1);
            if (!seg_end) continue;
            const bool from_start = seg_r0 == 0, to_end = r == p.upt - 1;
            if (from_start && !to_end) {
                // head of a split pair-tile: publish for the same rank of the next pair
                // quad-major layout [n/4][row]: the 32 lanes (rows) of a warp write 512 contiguous bytes per instruction
                float4 *dst = reinterpret_cast<float4 *>(p.partial + (size_t)cta * 128 * NPAD) + row;
#pragma unroll
                for (int n = 0; n < NPAD; n += 4) dst[(n >> 2) * 128] = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
                __threadfence();
                epi_bar();
                if (et == 0) st_release(p.flags + cta, 1);
                continue;
            }
            if (!from_start && to_end) {
                if (et == 0) {
                    unsigned spin = 0;
                    while (ld_acquire(p.flags + cta - 2) == 0) {
                        if (++spin > kSpinLimit || *abort_flag) {
                            *abort_flag = 1;
                            atomicCAS(p.status, 0, 8);
                            break;
                        }
                    }
                }
                epi_bar();
                const float4 *src = reinterpret_cast<const float4 *>(p.partial + (size_t)(cta - 2) * 128 * NPAD) + row;
#pragma unroll
                for (int n = 0; n < NPAD; n += 4) {
                    const float4 hd = __ldcg(src + (n >> 2) * 128);
                    acc[n] = fadd(hd.x, acc[n]); acc[n + 1] = fadd(hd.y, acc[n + 1]);
                    acc[n + 2] = fadd(hd.z, acc[n + 2]); acc[n + 3] = fadd(hd.w, acc[n + 3]);
                }
            } else if (!(from_start && to_end)) {
                if (lane == 0) atomicCAS(p.status, 0, 9);
                continue;
            }
            // whole tile in registers: x 1/(s_a*s_w), + bias -> pred   (ghost tile: img == batch, nothing stored)
            const int x = it.tx * kTileX + (AO ? row / kTileY : row % kTileX), y = it.ty * kTileY + (AO ? row % kTileY : row / kTileX);
            const bool inb = it.img < p.batch && y < p.gh && x < p.gw;
            const int slab = it.slab;
            // the feature scales were divided out chunk by chunk; the weight scale belongs to the slab
            const float inv = PK ? reinterpret_cast<const PackedHeader *>(reinterpret_cast<const char *>(p.whdr) +
                                                                          (size_t)slab * p.slab_stride)->inv_scale
                                 : inv_sw;
#pragma unroll
            for (int n = 0; n < NPAD; ++n) acc[n] = fadd(fmul(acc[n], inv), s_bias[n]);
            if (PK && NPAD == 128) {
                // The dgrad GEMM writes 512 B per pixel.  One thread per pixel row (the TMEM lane layout) makes every store
                // instruction touch 32 different lines -- ncu: 16 % of all stall samples are LSU throttling on these stores.
                // Transpose through a warp-private staging tile instead (64 channels per round): a store instruction then
                // covers two pixels x 256 contiguous bytes.
                float *stg = reinterpret_cast<float *>(ctrl + kCtrlBytes) + (size_t)q * (32 * kEpiPitch);
                const float *my_out = inb ? p.pred + (((size_t)it.img * p.gh + y) * p.gw + x) * p.out_stride + slab * NPAD : nullptr;
                const unsigned long long my_bits = reinterpret_cast<unsigned long long>(my_out);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    __syncwarp();
#pragma unroll
                    for (int n4 = 0; n4 < 16; ++n4)
                        *reinterpret_cast<float4 *>(stg + lane * kEpiPitch + 4 * n4) =
                            make_float4(acc[half * 64 + 4 * n4], acc[half * 64 + 4 * n4 + 1], acc[half * 64 + 4 * n4 + 2],
                                        acc[half * 64 + 4 * n4 + 3]);
                    __syncwarp();
#pragma unroll
                    for (int itp = 0; itp < 16; ++itp) {
                        const int pp = 2 * itp + (lane >> 4), q4 = lane & 15;
                        const float4 v = *reinterpret_cast<const float4 *>(stg + pp * kEpiPitch + 4 * q4);
                        const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)my_bits, pp);
                        const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(my_bits >> 32), pp);
                        float *dst = reinterpret_cast<float *>(((unsigned long long)hi << 32) | lo);
                        if (dst != nullptr) *reinterpret_cast<float4 *>(dst + half * 64 + 4 * q4) = v;
                    }
                }
            } else if (inb) {
                float *out = p.pred + (((size_t)it.img * p.gh + y) * p.gw + x) * p.out_stride + (PK ? slab * NPAD : 0);
                if (((p.cout | p.out_stride) & 3) == 0) {
                    float4 *o4 = reinterpret_cast<float4 *>(out);
#pragma unroll
                    for (int n = 0; n < NPAD; n += 4)
                        if (n < p.cout) o4[n >> 2] = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
                } else {
#pragma unroll
                    for (int n = 0; n < NPAD; ++n)
                        if (n < p.cout) out[n] = acc[n];
                }
            }
            if (emit && it.img < p.batch) hand_over(it.img, it.tx, it.ty);   // ghost tiles have nothing to score (warp-uniform)
        }
        if (emit) hand_over(-1, 0, 0);   // no more tiles: the scorer warp leaves
    } else if (CS > 0 && warp >= kScorerWarp0) {
        // ===== scorer warps (CS > 0): class softmax x confidence sigmoid, first-max argmax (the same sqd_score_candidate every
        // filter kernel uses) on the fp32 logits of a finished tile, read back from the pred rows accumulate warp q stored
        // (L2 hits); anchors above the score threshold go to the image's candidate list with ONE atomic per warp and tile.
        // These warps are off the TMEM drain path: a tile period is ~58 k cycles, scoring its 9 x 32 anchors takes a few k.
        constexpr int kNF = CS + 5, KMAX = NPAD / kNF;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const sqd_u64 floor_key = sqd_score_floor_key(p.score_thr);
        if (p.cand.count != nullptr) {
            for (int n = 0;; ++n) {
                const int slot = n & 1;
                const uint32_t ph = (uint32_t)(n >> 1) & 1u;
                if (!mbar_wait_warp(sc_full + q * 2 + slot, ph, abort_flag)) {
                    if (lane == 0) atomicCAS(p.status, 0, 11);
                    break;
                }
                const int *rec = sc_rec + (q * 2 + slot) * 4;
                const int img = rec[0], tx = rec[1], ty = rec[2];
                if (img < 0) break;
                const int x = tx * kTileX + row % kTileX, y = ty * kTileY + row / kTileX;
                const bool inb = y < p.gh && x < p.gw;
                const float *cell = p.pred + (((size_t)img * p.gh + y) * p.gw + x) * p.out_stride;
                const int a0 = (y * p.gw + x) * p.anchors_per_cell;
                bool pass[KMAX];
                sqd_u64 key[KMAX];
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    const bool want = inb && k < p.anchors_per_cell;
                    float f[CS + 1];
#pragma unroll
                    for (int j = 0; j <= CS; ++j) f[j] = 0.f;
                    if (want) {
                        const float *src = cell + k * kNF;
                        if (CS == 3) {
                            const float4 v = __ldcg(reinterpret_cast<const float4 *>(src));
                            f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
                        } else {
#pragma unroll
                            for (int j = 0; j <= CS; ++j) f[j] = __ldcg(src + j);
                        }
                    }
                    float scv = 0.f;
                    int cl = 0;
                    const bool cand_ok = sqd_score_candidate<CS>(f, CS, p.score_thr, scv, cl);
                    key[k] = sqd_make_key(scv, a0 + k, cl);
                    pass[k] = want && cand_ok && key[k] > floor_key;
                }
                sqd_cand_append_warp<KMAX>(p.cand, img, pass, key);
                __syncwarp();
                if (lane == 0) mbar_arrive(sc_empty + q * 2 + slot);
            }
        }
    }
    SQD_TRACE_PH(5);     // the producer warp is through (thread 0 is its lane 0)
    tc_fence_before();
    __syncthreads();
    SQD_TRACE_PH(6);
    cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still touch its shared memory / barriers
    tc_fence_after();
    SQD_TRACE_PH(7);
    if (warp == kWarpMma2) {
        __syncwarp();
        tmem_dealloc_2cta(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;  // benign race: every thread resolves the same pointer
    if (fn) return fn;
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
    return fn;
}

int npad_of(int cout) { return (cout + 15) / 16 * 16; }
// output channels per CTA of a pair (see pack_weights_f16_kernel / convdet_f16_pair_kernel)
int pair_hr_of(int cout) {
    const int npad = npad_of(cout);
    if (npad == 80 && cout <= 72) return 36;      // KITTI: 9 anchors x (3 + 5)
    return npad / 2;
}


int a_stages_for(int npad) {
    int s = sqd_opt(SQD_OPT_F16_A_STAGES);
    (void)npad;
    return s < 1 ? 1 : (s > 4 ? 4 : s);
}

int b_stages_for(int npad) {
    const size_t stage = (size_t)2 * npad * kBlockK * 2;
    size_t s = (kSmemLimit - 1024 /*align*/ - kCtrlBytes - (size_t)a_stages_for(npad) * kAStageBytes) / stage;
    if (s > 8) s = 8;
    const int cap = sqd_opt(SQD_OPT_F16_B_STAGES);
    if ((int)s > cap && cap >= 1) s = cap;
    return (int)s;
}

size_t smem_bytes_for(int npad, int a_stages, int b_stages) {
    return 1024 + (size_t)a_stages * kAStageBytes + (size_t)b_stages * 2 * npad * kBlockK * 2 + kCtrlBytes;
}

int grid_for(int total_tiles) { return total_tiles < SQD_SM_COUNT ? total_tiles : SQD_SM_COUNT; }

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// planes buffer (also the SQD_LAYOUT_SPLIT_NHWC input):
//   [max|x| bits per (image, 64-channel block): B*(Cin/64) x u32, padded to 256 B][x1 plane][x2 plane]
struct PlaneLayout {
    size_t p1_off, p2_off, total;
};
PlaneLayout plane_layout(int batch, int cin, int gh, int gw) {
    PlaneLayout l;
    const size_t plane = align256((size_t)batch * gh * gw * cin * sizeof(__half));
    l.p1_off = align256((size_t)batch * (cin / kBlockK) * sizeof(unsigned));
    l.p2_off = l.p1_off + plane;
    l.total = l.p2_off + plane;
    return l;
}

// workspace layout: [status (256 B)][flags: 256 ints][partials: #SM*128*npad floats][planes unless pre-split input]
struct WsLayout {
    size_t flags_off, partial_off, planes_off, total;
};
WsLayout ws_layout(int batch, int cin, int gh, int gw, int cout, int layout) {
    WsLayout w;
    w.flags_off = 256;
    w.partial_off = w.flags_off + 256 * sizeof(int);
    w.planes_off = align256(w.partial_off + (size_t)SQD_SM_COUNT * 128 * npad_of(cout) * sizeof(float));
    w.total = w.planes_off + (layout == SQD_LAYOUT_SPLIT_NHWC ? 0 : plane_layout(batch, cin, gh, gw).total);
    return w;
}

template <int NPAD>
int launch_f16(const CUtensorMap *maps, const F16Params &p, int grid, cudaStream_t st) {
    const size_t smem = smem_bytes_for(NPAD, p.a_stages, p.b_stages);
    SQD_CUDA(cudaFuncSetAttribute(convdet_f16_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    convdet_f16_kernel<NPAD><<<grid, kThreads, smem, st>>>(maps[0], maps[1], maps[2], p);
    SQD_LAUNCH_CHECK("convdet_f16_kernel");
    return SQD_OK;
}

}  // namespace

size_t sqd_f16_packed_bytes(int cout, int cin) {
    return kHeaderBytes + (size_t)2 * 2 * npad_of(cout) * 9 * cin * sizeof(__half);  // 1-CTA and CTA-pair layouts
}

int sqd_f16_pack_weights(const float *d_weight, int cout, int cin, void *d_packed, cudaStream_t st) {
    SQD_REQUIRE(cin % kBlockK == 0, SQD_E_SHAPE, "convdet (tcgen05): Cin %d must be a multiple of %d", cin, kBlockK);
    SQD_REQUIRE(cout >= 1 && cout <= 128, SQD_E_SHAPE, "convdet (tcgen05): Cout %d outside [1,128]", cout);
    PackedHeader *hdr = static_cast<PackedHeader *>(d_packed);
    __half *mat = reinterpret_cast<__half *>(static_cast<char *>(d_packed) + kHeaderBytes);
    SQD_CUDA(cudaMemsetAsync(d_packed, 0, kHeaderBytes, st));
    const size_t n = (size_t)cout * cin * 9;
    weight_absmax_kernel<<<SQD_SM_COUNT, 256, 0, st>>>(d_weight, n, hdr);
    SQD_LAUNCH_CHECK("weight_absmax_kernel");
    pack_weights_f16_kernel<<<SQD_SM_COUNT * 4, 256, 0, st>>>(d_weight, cout, cin, npad_of(cout), pair_hr_of(cout), hdr, mat,
                                                              mat + (size_t)2 * npad_of(cout) * 9 * cin);
    SQD_LAUNCH_CHECK("pack_weights_f16_kernel");
    return SQD_OK;
}

// ---- the dgrad GEMM's weight slabs, all of them in TWO launches ---------------------------------------------------------
// Slab s is the packed form (header + both fp16 layouts, exactly what sqd_f16_pack_weights writes) of the flipped /
// transposed matrix W2_s (slab_rows, kp, 3, 3), W2_s[j][n][t] = W[n][s*slab_rows + j][8 - t], zero outside W.  Built
// slab by slab that is a transpose kernel, a memset, a max kernel and a pack kernel per slab -- 24 launches of a few
// microseconds each for the KITTI head, 70 us of a training step that re-packs after every optimizer step; here one block
// per slab reduces max|W| of its channels straight from W and one grid packs every slab from W.  Same bytes.
namespace {
constexpr int kSlabMaxBlocks = 24;   // blocks per slab of the max kernel; their partial maxima go through `partial`
__global__ void __launch_bounds__(256) dgrad_slab_absmax_kernel(const float *__restrict__ w, int cout, int cin, int slab_rows,
                                                                float *__restrict__ partial) {
    const int s = blockIdx.y;
    const int c0 = s * slab_rows, nc = min(slab_rows, cin - c0), run = nc * 9;   // per output channel: one contiguous run
    float m = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cout * run; i += gridDim.x * blockDim.x) {
        const int n = i / run, r = i - n * run;
        m = fmaxf(m, fabsf(w[((size_t)n * cin + c0) * 9 + r]));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float s_m[8];
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) m = fmaxf(m, s_m[i]);
        partial[s * kSlabMaxBlocks + blockIdx.x] = m;
    }
}

__global__ void pack_dgrad_slabs_kernel(const float *__restrict__ w, int cout, int cin, int slab_rows, int kp, int npad, int hr,
                                        char *base, size_t slab_bytes, const float *__restrict__ partial) {
    const int s = blockIdx.y;
    PackedHeader *hdr = reinterpret_cast<PackedHeader *>(base + (size_t)s * slab_bytes);
    const size_t ktot = (size_t)9 * kp, total = (size_t)npad * ktot;
    __half *mat = reinterpret_cast<__half *>(base + (size_t)s * slab_bytes + kHeaderBytes), *mat2 = mat + 2 * total;
    float amax = 0.f;
    for (int i = 0; i < kSlabMaxBlocks; ++i) amax = fmaxf(amax, __ldg(partial + s * kSlabMaxBlocks + i));
    const float sc = pow2_scale_for(amax);
    if (blockIdx.x == 0) {   // the header: every byte, like the memset + field writes of sqd_f16_pack_weights
        if (threadIdx.x < kHeaderBytes / 4) reinterpret_cast<unsigned *>(hdr)[threadIdx.x] = 0u;
        __syncthreads();
        if (threadIdx.x == 0) {
            hdr->amax_bits = __float_as_uint(amax);
            hdr->scale = sc;
            hdr->inv_scale = 1.f / sc;
            hdr->npad = npad;
            hdr->cin = kp;
            hdr->hr = hr;
        }
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / ktot);                       // row of the slab: feature channel s*slab_rows + n
        const size_t k = i % ktot;
        const int tap = (int)(k / kp), c = (int)(k % kp);    // c: gradient (output) channel, padded to kp
        const int fc = s * slab_rows + n;
        float v = 0.f;
        if (n < slab_rows && c < cout && fc < cin) v = w[((size_t)c * cin + fc) * 9 + (8 - tap)];
        __half h1, h2;
        split_f16(v * sc, h1, h2);
        mat[i] = h2;
        mat[total + i] = h1;
        if (n < 2 * hr) {
            const int r = n / hr, j = n - r * hr;
            mat2[((size_t)(r * 2 * hr + j)) * ktot + k] = h1;
            mat2[((size_t)(r * 2 * hr + hr + j)) * ktot + k] = h2;
        }
    }
}
}  // namespace

// d_scratch: at least nslabs * 24 floats (the caller passes the scratch matrix behind the slabs)
int sqd_f16_pack_dgrad_slabs(const float *d_weight, int cout, int cin, int slab_rows, int kp, int nslabs, void *d_packed,
                             size_t slab_bytes, float *d_scratch, cudaStream_t st) {
    SQD_REQUIRE(kp % kBlockK == 0 && slab_rows >= 1 && slab_rows <= 128, SQD_E_SHAPE, "dgrad weight slabs: bad shape");
    char *base = static_cast<char *>(d_packed);
    dgrad_slab_absmax_kernel<<<dim3(kSlabMaxBlocks, nslabs), 256, 0, st>>>(d_weight, cout, cin, slab_rows, d_scratch);
    SQD_LAUNCH_CHECK("dgrad_slab_absmax_kernel");
    pack_dgrad_slabs_kernel<<<dim3(SQD_SM_COUNT, nslabs), 256, 0, st>>>(d_weight, cout, cin, slab_rows, kp, npad_of(slab_rows),
                                                                         pair_hr_of(slab_rows), base, slab_bytes, d_scratch);
    SQD_LAUNCH_CHECK("pack_dgrad_slabs_kernel");
    return SQD_OK;
}

namespace {
int pair_stages_for(int n1h) {
    const size_t stage = (size_t)kAStageBytes + (size_t)3 * n1h * kBlockK * 2;
    size_t s = (kSmemLimit - 1024 - kCtrlBytes) / stage;
    if (s > 4) s = 4;
    const int cap = sqd_opt(SQD_OPT_F16_PAIR_STAGES);
    if ((int)s > cap && cap >= 1) s = cap;
    return (int)s;
}

// A-once layout (AO kernels): the largest operand-buffer count (<= 3) that leaves room for three weight stages
int ao_bufs_for(int n1h) {
    const size_t avail = kSmemLimit - 1024 - kCtrlBytes - (size_t)3 * 3 * n1h * kBlockK * 2;
    size_t n = avail / kAoBufBytes;
    const int cap = sqd_opt(SQD_OPT_F16_AO_BUFS);   // 2: a buffer is refilled a whole block (three units) ahead of its use
    if ((int)n > cap && cap >= 2) n = cap;
    return n > 3 ? 3 : (int)n;
}

template <int NPAD, int CS = 0, bool PK = false, int HR = NPAD / 2, bool K3 = (NPAD >= 96), bool AO = false>
int launch_pair(const CUtensorMap *maps, const PairParams &p, int grid, cudaStream_t st) {
    const size_t smem = AO ? 1024 + (size_t)p.a_bufs * kAoBufBytes + (size_t)p.stages * (3 * 2 * HR * kBlockK * 2) + kCtrlBytes
                           : 1024 + (size_t)p.stages * (kAStageBytes + 3 * 2 * HR * kBlockK * 2) + kCtrlBytes + (PK ? kEpiStageBytes : 0);
    SQD_REQUIRE(smem <= kSmemLimit, SQD_E_SHAPE, "convdet (tcgen05): shared memory budget exceeded (%zu bytes)", smem);
    SQD_CUDA(cudaFuncSetAttribute(convdet_f16_pair_kernel<NPAD, CS, PK, HR, K3, AO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // dependent launch when the pre-pass kernel directly precedes it on the stream (p.pdl)
    cudaError_t e = sqd_launch_dependent(convdet_f16_pair_kernel<NPAD, CS, PK, HR, K3, AO>, dim3(grid), dim3(pair_threads(CS, AO)), smem, st, p.pdl != 0,
                                         maps[0], maps[1], maps[2], p);
    if (e != cudaSuccess) {
        sqd_set_error("launch of convdet_f16_pair_kernel failed: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return SQD_OK;
}
}  // namespace

int sqd_f16_split_features(const float *d_feat, int layout, int batch, int cin, int gh, int gw, void *d_planes,
                           cudaStream_t st);

// emit (optional): ask the epilogue to score the anchors and fill per-image candidate lists (emit->cand, counts already
// zeroed on `st`).  *emit->done is set to 1 when the epilogue will do it; the caller otherwise runs the stand-alone
// scan over pred.  MEASURED (profiles/r01_fused_score_epilogue.txt): scoring inside the GEMM kernel costs more than the
// separate 8 us scan kernel it replaces (+23 us at B = 20, +15 % at B = 256, immediate or software-pipelined), so it is
// an opt-in experiment (SQD_FUSED_SCORE=1), not the default.
int sqd_convdet_f16_pair(const float *d_feat, int layout, const void *d_packed, const float *d_bias, int batch, int cin,
                         int gh, int gw, int cout, float *d_pred, void *d_workspace, cudaStream_t st,
                         const SqdCandEmit *emit, int out_stride, int ksteps_last, int slabs, size_t slab_stride,
                         int image_scales) {
    SQD_REQUIRE(cin % kBlockK == 0, SQD_E_SHAPE, "convdet (tcgen05): Cin %d must be a multiple of %d", cin, kBlockK);
    SQD_REQUIRE(cout >= 1 && cout <= 128, SQD_E_SHAPE, "convdet (tcgen05): Cout %d outside [1,128]", cout);
    EncodeTiledFn encode = get_encode_fn();
    SQD_REQUIRE(encode != nullptr, SQD_E_DRIVER, "cuTensorMapEncodeTiled not available from the CUDA driver");
    const int npad = npad_of(cout), n1h = 2 * pair_hr_of(cout);
    const WsLayout w = ws_layout(batch, cin, gh, gw, cout, layout);
    char *ws = static_cast<char *>(d_workspace);
    SQD_CUDA(cudaMemsetAsync(ws, 0, w.partial_off, st));  // status + flags

    const char *planes = reinterpret_cast<const char *>(d_feat);
    bool after_prepass = false;
    if (layout != SQD_LAYOUT_SPLIT_NHWC) {
        int rc = sqd_f16_split_features(d_feat, layout, batch, cin, gh, gw, ws + w.planes_off, st);
        if (rc) return rc;
        planes = ws + w.planes_off;
        after_prepass = true;   // a kernel directly precedes the GEMM on the stream: programmatic dependent launch
    }
    const PlaneLayout pl = plane_layout(batch, cin, gh, gw);

    if (slabs < 1) slabs = 1;
    const bool multi = (slabs > 1 || (ksteps_last >= 1 && ksteps_last < 4)) && npad == 128;
    // A-once operand layout (one patch fetch per (tile, block), see kAoY): the plain forward kernel of the two-MMA scheme
    const bool fused_score = emit && sqd_opt(SQD_OPT_FUSED_SCORE);
    const bool ao = !multi && !fused_score && npad <= 80 && ao_bufs_for(n1h) >= 2 && sqd_opt(SQD_OPT_F16_A_ONCE);
    alignas(64) CUtensorMap maps[3];
    for (int i = 0; i < 2; ++i) {
        // AO: dimension order (channel, y, x, image), so that the box lands x-major / y-minor in shared memory
        const cuuint64_t dims_xy[4] = {(cuuint64_t)cin, (cuuint64_t)gw, (cuuint64_t)gh, (cuuint64_t)batch};
        const cuuint64_t strides_xy[3] = {(cuuint64_t)cin * 2, (cuuint64_t)gw * cin * 2, (cuuint64_t)gh * gw * cin * 2};
        const cuuint32_t box_xy[4] = {kBlockK, kTileX, kPatchY, 1};
        const cuuint64_t dims_yx[4] = {(cuuint64_t)cin, (cuuint64_t)gh, (cuuint64_t)gw, (cuuint64_t)batch};
        const cuuint64_t strides_yx[3] = {(cuuint64_t)gw * cin * 2, (cuuint64_t)cin * 2, (cuuint64_t)gh * gw * cin * 2};
        const cuuint32_t box_yx[4] = {kBlockK, kAoY, kAoX, 1};
        const cuuint64_t *dims = ao ? dims_yx : dims_xy, *strides = ao ? strides_yx : strides_xy;
        const cuuint32_t *box = ao ? box_yx : box_xy;
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        void *base = const_cast<char *>(planes + (i == 0 ? pl.p1_off : pl.p2_off));
        CUresult r = encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SQD_REQUIRE(r == CUDA_SUCCESS, SQD_E_DRIVER, "cuTensorMapEncodeTiled(features) failed: CUresult %d", (int)r);
    }
    // the dgrad GEMM (PK kernel, `multi`): several weight matrices over the same planes in one launch and / or a partly
    // filled last channel block; it addresses the weights through a 3-D map {k, row, slab}
    SQD_REQUIRE(slabs == 1 || multi, SQD_E_UNSUPPORTED, "convdet (tcgen05): multi-slab launches need Cout == 128 per slab");
    if (slabs == 1) slab_stride = 0;
    {
        const size_t ktot = (size_t)9 * cin;
        void *mat2 = const_cast<char *>(static_cast<const char *>(d_packed) + kHeaderBytes) + (size_t)2 * npad * ktot * sizeof(__half);
        CUresult r;
        if (multi) {
            SQD_REQUIRE(slab_stride % 16 == 0, SQD_E_SHAPE, "convdet (tcgen05): bad slab layout");
            const cuuint64_t dims[3] = {(cuuint64_t)ktot, (cuuint64_t)(2 * n1h), (cuuint64_t)slabs};
            const cuuint64_t strides[2] = {(cuuint64_t)ktot * 2, (cuuint64_t)(slabs > 1 ? slab_stride : (size_t)2 * npad * ktot * 2)};
            const cuuint32_t box[3] = {kBlockK, (cuuint32_t)n1h, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            r = encode(&maps[2], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, mat2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)(2 * n1h)};
            const cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
            const cuuint32_t box[2] = {kBlockK, (cuuint32_t)n1h};    // one CTA's half: 2*hr rows
            const cuuint32_t estr[2] = {1, 1};
            r = encode(&maps[2], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, mat2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        SQD_REQUIRE(r == CUDA_SUCCESS, SQD_E_DRIVER, "cuTensorMapEncodeTiled(weights) failed: CUresult %d", (int)r);
    }

    PairParams p;
    p.cin = cin; p.gh = gh; p.gw = gw; p.cout = cout; p.batch = batch;
    p.slabs = slabs; p.slab_stride = slab_stride;
    p.out_stride = out_stride > 0 ? out_stride : cout;
    p.pdl = after_prepass ? 1 : 0;
    p.ksteps_last = (ksteps_last >= 1 && ksteps_last < 4) ? ksteps_last : 4;
    p.tiles_x = (gw + kTileX - 1) / kTileX;
    p.tiles_per_img = p.tiles_x * ((gh + kTileY - 1) / kTileY);
    p.tiles_per_slab = (p.tiles_per_img * batch + 1) & ~1;   // PK: slabs padded to an even tile count
    const long long total_tiles = multi ? (long long)p.tiles_per_slab * slabs : (long long)p.tiles_per_img * batch;
    SQD_REQUIRE(total_tiles < (1ll << 30), SQD_E_SHAPE, "convdet (tcgen05): too many tiles");
    p.total_tiles = (int)total_tiles;
    p.pair_tiles = (int)((total_tiles + 1) / 2);
    p.upt = cin / kBlockK * 3;
    const int max_pairs = SQD_SM_COUNT / 2;
    int npairs = p.pair_tiles < max_pairs ? p.pair_tiles : max_pairs;
    const long long total_units = (long long)p.pair_tiles * p.upt;
    long long upp = (total_units + npairs - 1) / npairs;
    if (upp < p.upt) upp = p.upt;
    // Opt-in (SQD_F16_HALF_TILES=1) for small batches (fewer pair-tiles than half the CTA pairs): give every pair-tile to
    // TWO pairs, exactly half of its channel blocks each -- the head / tail hand-off handles two holders per tile --
    // which cuts the latency of a batch-1 call from 60 to 51 us.  Not the default: as long as every pair owns whole
    // tiles (up to 74 tiles = 4 KITTI images per launch) the fp32 summation order of a cell does not depend on how a
    // batch is chunked, and the host-buffer entry point relies on that to return bit-identical results for any chunk size.
    if (2 * p.pair_tiles <= max_pairs && (cin / kBlockK) % 2 == 0 && sqd_opt(SQD_OPT_F16_HALF_TILES)) {
        npairs = 2 * p.pair_tiles;
        upp = p.upt / 2;
    }
    p.units_per_pair = (int)upp;
    p.stages = pair_stages_for(n1h);
    p.a_bufs = 0;
    if (ao) {
        p.a_bufs = ao_bufs_for(n1h);
        const size_t left = kSmemLimit - 1024 - kCtrlBytes - (size_t)p.a_bufs * kAoBufBytes;
        int sb = (int)(left / ((size_t)3 * n1h * kBlockK * 2));
        const int cap = sqd_opt(SQD_OPT_F16_PAIR_STAGES);
        p.stages = sb > 4 ? 4 : sb;
        if (p.stages > cap && cap >= 2) p.stages = cap;
    }
    SQD_REQUIRE(p.stages >= 2, SQD_E_SHAPE, "convdet (tcgen05): shared memory too small for two stages");
    p.chunk_units = sqd_opt(SQD_OPT_F16_CHUNK);   // one 64-channel block (3 dx units) per TMEM chunk
    if (p.chunk_units < 1) p.chunk_units = 1;
    p.chunk_blk = 3;
    if (image_scales && p.upt <= 6) {              // every block of an image shares one scale: a whole tile per chunk
        p.chunk_blk = p.upt;
        p.chunk_units = p.upt;
    }
    p.dbg = sqd_opt(SQD_OPT_F16_DBG);
    p.bias = d_bias;
    p.amax_bits = reinterpret_cast<const unsigned *>(planes);
    p.whdr = static_cast<const PackedHeader *>(d_packed);
    p.pred = d_pred;
    p.partial = reinterpret_cast<float *>(ws + w.partial_off);
    p.flags = reinterpret_cast<int *>(ws + w.flags_off);
    p.status = reinterpret_cast<int *>(ws);
    p.trace = nullptr;
    p.trace_cta = sqd_opt(SQD_OPT_F16_TRACE_CTA);
#ifdef SQD_ENABLE_TRACE   // profiling builds only (SQD_BUILD_TRACE=1 python csrc/build.py): device address of a clock64 stamp buffer
    if (const char *e = getenv("SQD_F16_TRACE")) p.trace = reinterpret_cast<long long *>(strtoull(e, nullptr, 0));
#endif
    const int grid = 2 * npairs;
    p.cand.count = nullptr;
    p.cand.keys = nullptr;
    p.cand.stride = 0;
    p.score_thr = 0.f;
    p.anchors_per_cell = 0;
    if (emit) {
        *emit->done = 0;
        const int nf = emit->num_classes + 5;
        const bool shape_ok = cout % nf == 0 && ((emit->num_classes == 3 && npad == 80) || (emit->num_classes == 8 && npad == 128));
        if (shape_ok && sqd_opt(SQD_OPT_FUSED_SCORE)) {
            p.cand = emit->cand;
            p.score_thr = emit->score_thr;
            p.anchors_per_cell = cout / nf;
            *emit->done = 1;
            if (npad == 80) return n1h == 72 ? launch_pair<80, 3, false, 36>(maps, p, grid, st) : launch_pair<80, 3>(maps, p, grid, st);
            return launch_pair<128, 8>(maps, p, grid, st);
        }
    }
    if (multi) return launch_pair<128, 0, true>(maps, p, grid, st);   // the dgrad GEMM
    p.ksteps_last = 4;
    if (ao) {
        switch (npad / 16) {
            case 1: return launch_pair<16, 0, false, 8, false, true>(maps, p, grid, st);
            case 2: return launch_pair<32, 0, false, 16, false, true>(maps, p, grid, st);
            case 3: return launch_pair<48, 0, false, 24, false, true>(maps, p, grid, st);
            case 4: return launch_pair<64, 0, false, 32, false, true>(maps, p, grid, st);
            case 5: return n1h == 72 ? launch_pair<80, 0, false, 36, false, true>(maps, p, grid, st)
                                     : launch_pair<80, 0, false, 40, false, true>(maps, p, grid, st);
        }
    }
    switch (npad / 16) {
        case 1: return launch_pair<16>(maps, p, grid, st);
        case 2: return launch_pair<32>(maps, p, grid, st);
        case 3: return launch_pair<48>(maps, p, grid, st);
        case 4: return launch_pair<64>(maps, p, grid, st);
        case 5: return n1h == 72 ? launch_pair<80, 0, false, 36>(maps, p, grid, st) : launch_pair<80>(maps, p, grid, st);
        case 6: return launch_pair<96>(maps, p, grid, st);
        case 7: return launch_pair<112>(maps, p, grid, st);
        case 8: return launch_pair<128>(maps, p, grid, st);
    }
    SQD_REQUIRE(false, SQD_E_SHAPE, "convdet (tcgen05): unsupported Cout %d", cout);
}

size_t sqd_f16_split_bytes(int batch, int cin, int gh, int gw) { return plane_layout(batch, cin, gh, gw).total; }

namespace {
// cluster size for the one-pass NCHW split: the smallest that lets >= 3 CTAs share an SM (<= 72 KB each); 0 = shape not
// eligible (two-pass fallback).  A slab that needs more than 72 KB per CTA even at the largest cluster (the stress grid:
// 64 x 7488 floats = 120 KB per CTA at 16) would run ONE CTA per SM, whose load / barrier / store phases then have nothing
// to overlap with: measured 1.010 ms for the one-pass kernel against 0.655 ms for max pass + split pass at 64 stress
// images (tools/shape_bench.py), so such shapes take the two passes.  SQD_SPLIT_CS forces a size (up to 200 KB per CTA).
int split_cluster_size(int P, size_t *smem_out, int *nq4p_out) {
    if (P % 4 != 0 || sqd_opt(SQD_OPT_SPLIT_TWO_PASS)) return 0;
    const int n4 = P / 4;
    int pick = 0;
    const int force = sqd_opt(SQD_OPT_SPLIT_CS);
    for (int pass = 0; pass < (force ? 2 : 1) && !pick; ++pass)
        for (int cs = 1; cs <= kSplitMaxCluster; cs <<= 1) {
            if (cs > n4) break;
            if (force && cs != force) continue;
            const int nq4 = (n4 + cs - 1) / cs;              // largest share of any rank
            const int nq4p = (nq4 + 7) & ~7;
            const size_t smem = (size_t)64 * nq4p * sizeof(float4);
            if (smem <= (pass == 0 ? (size_t)72 * 1024 : (size_t)200 * 1024)) {
                pick = cs;
                *smem_out = smem;
                *nq4p_out = nq4p;
                break;
            }
        }
    return pick;
}
}  // namespace

int sqd_f16_split_features(const float *d_feat, int layout, int batch, int cin, int gh, int gw, void *d_planes,
                           cudaStream_t st) {
    SQD_REQUIRE(cin % kBlockK == 0, SQD_E_SHAPE, "convdet (tcgen05): Cin %d must be a multiple of %d", cin, kBlockK);
    SQD_REQUIRE(cin <= 4096, SQD_E_SHAPE, "convdet (tcgen05): Cin %d > 4096", cin);
    const PlaneLayout pl = plane_layout(batch, cin, gh, gw);
    char *base = static_cast<char *>(d_planes);
    unsigned *amax = reinterpret_cast<unsigned *>(base);
    uint2 *p1 = reinterpret_cast<uint2 *>(base + pl.p1_off), *p2 = reinterpret_cast<uint2 *>(base + pl.p2_off);
    const int P = gh * gw, ncb = cin / kBlockK;
    if (layout == SQD_LAYOUT_NHWC) {
        // channels_last: a (image, block) slab is strided, so max and split stay two passes over the image
        SQD_CUDA(cudaMemsetAsync(amax, 0, pl.p1_off, st));
        int bx = P < 64 ? P : 64;
        if ((long long)bx * batch < 2 * SQD_SM_COUNT) bx = P < 256 ? P : 256;
        absmax_nhwc_kernel<<<dim3(bx, batch), cin / 4, 0, st>>>(reinterpret_cast<const float4 *>(d_feat), P, cin / 4, amax);
        SQD_LAUNCH_CHECK("absmax_nhwc_kernel");
        int sx = P < 128 ? P : 128;
        if ((long long)sx * batch < 4 * SQD_SM_COUNT) sx = P < 512 ? P : 512;
        split_nhwc_f16_kernel<<<dim3(sx, batch), cin / 4, 0, st>>>(reinterpret_cast<const float4 *>(d_feat), p1, p2, P,
                                                                  cin / 4, amax);
        SQD_LAUNCH_CHECK("split_nhwc_f16_kernel");
        return SQD_OK;
    }
    size_t smem = 0;
    int nq4p = 0;
    const int cs = split_cluster_size(P, &smem, &nq4p);
    if (cs > 0 && sqd_opt(SQD_OPT_SPLIT_REGS) && P % 4 == 0 && (P / 4 + kRsCluster - 1) / kRsCluster <= 8 * (kRsThreads / 32) &&
        !sqd_opt(SQD_OPT_SPLIT_CS)) {
        // register-resident one-pass kernel: a CTA covers at most 64 float4 columns (8 warps x 8)
        split_nchw_regs_cluster_kernel<<<(unsigned)((size_t)batch * ncb * kRsCluster), kRsThreads, 0, st>>>(
            d_feat, reinterpret_cast<uint4 *>(p1), reinterpret_cast<uint4 *>(p2), cin, P, amax);
        SQD_LAUNCH_CHECK("split_nchw_regs_cluster_kernel");
        return SQD_OK;
    }
    if (cs > 0) {
        // one pass: a cluster per (image, channel block) slab
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((size_t)batch * ncb * cs));
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const int threads = sqd_opt(SQD_OPT_SPLIT_THREADS), rows = sqd_opt(SQD_OPT_SPLIT_ROWS);
        cudaError_t e = cudaErrorInvalidValue;
#define SQD_SPLIT_LAUNCH(T, R)                                                                                              \
    do {                                                                                                                    \
        SQD_CUDA(cudaFuncSetAttribute(split_nchw_cluster_kernel<T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        if (cs > 8)                                                                                                         \
            SQD_CUDA(cudaFuncSetAttribute(split_nchw_cluster_kernel<T, R>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)); \
        cfg.blockDim = dim3(T);                                                                                             \
        e = cudaLaunchKernelEx(&cfg, split_nchw_cluster_kernel<T, R>, d_feat, p1, p2, cin, P, amax, cs, nq4p);              \
    } while (0)
        if (threads == 256 && rows) SQD_SPLIT_LAUNCH(256, true);
        else if (threads == 256) SQD_SPLIT_LAUNCH(256, false);
        else if (threads == 1024 && rows) SQD_SPLIT_LAUNCH(1024, true);
        else if (threads == 1024) SQD_SPLIT_LAUNCH(1024, false);
        else if (rows) SQD_SPLIT_LAUNCH(512, true);
        else SQD_SPLIT_LAUNCH(512, false);
#undef SQD_SPLIT_LAUNCH
        if (e != cudaSuccess) {
            sqd_set_error("launch of split_nchw_cluster_kernel failed: %s", cudaGetErrorString(e));
            return (int)e;
        }
        return SQD_OK;
    }
    // two passes: per-slab max (a slab is one contiguous run of 64*P floats), then the tiled transpose/split
    SQD_CUDA(cudaMemsetAsync(amax, 0, pl.p1_off, st));
    {
        const size_t n4 = (size_t)kBlockK * P / 4;
        int bx = (int)((n4 + 256 * 8 - 1) / (256 * 8));
        if (bx > 8) bx = 8;
        if (bx < 1) bx = 1;
        absmax_kernel<<<dim3(bx, batch * ncb), 256, 0, st>>>(reinterpret_cast<const float4 *>(d_feat), n4, amax);
        SQD_LAUNCH_CHECK("absmax_kernel");
    }
    dim3 grid((P + 63) / 64, ncb, batch);
    split_nchw_f16_kernel<<<grid, 256, 0, st>>>(d_feat, p1, p2, cin, P, amax);
    SQD_LAUNCH_CHECK("split_nchw_f16_kernel");
    return SQD_OK;
}

// max |x| of `nruns` contiguous runs of run_floats fp32 values (d_amax zeroed by the caller); used by the wgrad pre-pass
int sqd_f16_absmax_runs(const float *d_in, size_t run_floats, int nruns, unsigned *d_amax, cudaStream_t st) {
    SQD_REQUIRE(run_floats % 4 == 0 && nruns >= 1 && nruns <= 65535, SQD_E_SHAPE, "absmax: bad run shape");
    const size_t n4 = run_floats / 4;
    int bx = (int)((n4 + 256 * 8 - 1) / (256 * 8));
    if (bx > 8) bx = 8;
    if (bx < 1) bx = 1;
    absmax_kernel<<<dim3(bx, nruns), 256, 0, st>>>(reinterpret_cast<const float4 *>(d_in), n4, d_amax);
    SQD_LAUNCH_CHECK("absmax_kernel");
    return SQD_OK;
}

size_t sqd_f16_workspace_bytes(int batch, int cin, int gh, int gw, int cout, int layout) {
    return ws_layout(batch, cin, gh, gw, cout, layout).total;
}

int sqd_convdet_f16(const float *d_feat, int layout, const void *d_packed, const float *d_bias, int batch, int cin,
                    int gh, int gw, int cout, float *d_pred, void *d_workspace, cudaStream_t st) {
    SQD_REQUIRE(cin % kBlockK == 0, SQD_E_SHAPE, "convdet (tcgen05): Cin %d must be a multiple of %d", cin, kBlockK);
    SQD_REQUIRE(cout >= 1 && cout <= 128, SQD_E_SHAPE, "convdet (tcgen05): Cout %d outside [1,128]", cout);
    EncodeTiledFn encode = get_encode_fn();
    SQD_REQUIRE(encode != nullptr, SQD_E_DRIVER, "cuTensorMapEncodeTiled not available from the CUDA driver");
    const int npad = npad_of(cout);
    const WsLayout w = ws_layout(batch, cin, gh, gw, cout, layout);
    char *ws = static_cast<char *>(d_workspace);
    int *status = reinterpret_cast<int *>(ws);
    int *flags = reinterpret_cast<int *>(ws + w.flags_off);
    float *partial = reinterpret_cast<float *>(ws + w.partial_off);
    SQD_CUDA(cudaMemsetAsync(ws, 0, w.partial_off, st));  // status + flags

    // 1. fp16 planes (x1, x2) + per-image max: from the workspace, or handed in pre-split
    const char *planes = reinterpret_cast<const char *>(d_feat);
    if (layout != SQD_LAYOUT_SPLIT_NHWC) {
        int rc = sqd_f16_split_features(d_feat, layout, batch, cin, gh, gw, ws + w.planes_off, st);
        if (rc) return rc;
        planes = ws + w.planes_off;
    }
    const PlaneLayout pl = plane_layout(batch, cin, gh, gw);

    // 2. tensor maps
    alignas(64) CUtensorMap maps[3];
    for (int i = 0; i < 2; ++i) {
        const cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)gw, (cuuint64_t)gh, (cuuint64_t)batch};
        const cuuint64_t strides[3] = {(cuuint64_t)cin * 2, (cuuint64_t)gw * cin * 2, (cuuint64_t)gh * gw * cin * 2};
        const cuuint32_t box[4] = {kBlockK, kTileX, kPatchY, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        void *base = const_cast<char *>(planes + (i == 0 ? pl.p1_off : pl.p2_off));
        CUresult r = encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SQD_REQUIRE(r == CUDA_SUCCESS, SQD_E_DRIVER, "cuTensorMapEncodeTiled(features) failed: CUresult %d", (int)r);
    }
    {
        const size_t ktot = (size_t)9 * cin;
        void *mat = const_cast<char *>(static_cast<const char *>(d_packed) + kHeaderBytes);
        const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)(2 * npad)};
        const cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
        const cuuint32_t box[2] = {kBlockK, (cuuint32_t)(2 * npad)};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&maps[2], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, mat, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SQD_REQUIRE(r == CUDA_SUCCESS, SQD_E_DRIVER, "cuTensorMapEncodeTiled(weights) failed: CUresult %d", (int)r);
    }

    // 3. the persistent GEMM
    F16Params p;
    p.cin = cin; p.gh = gh; p.gw = gw; p.cout = cout;
    p.tiles_x = (gw + kTileX - 1) / kTileX;
    p.tiles_per_img = p.tiles_x * ((gh + kTileY - 1) / kTileY);
    const long long total_tiles = (long long)p.tiles_per_img * batch;
    SQD_REQUIRE(total_tiles < (1ll << 30), SQD_E_SHAPE, "convdet (tcgen05): too many tiles");
    p.total_tiles = (int)total_tiles;
    p.upt = cin / kBlockK * 3;
    const int grid = grid_for(p.total_tiles);
    const long long total_units = total_tiles * p.upt;
    long long upc = (total_units + grid - 1) / grid;
    if (upc < p.upt) upc = p.upt;  // grid == #tiles: whole tiles only
    p.units_per_cta = (int)upc;
    p.a_stages = a_stages_for(npad);
    p.b_stages = b_stages_for(npad);
    p.dbg = sqd_opt(SQD_OPT_F16_DBG);
    p.bias = d_bias;
    p.amax_bits = reinterpret_cast<const unsigned *>(planes);
    p.whdr = static_cast<const PackedHeader *>(d_packed);
    p.pred = d_pred;
    p.partial = partial;
    p.flags = flags;
    p.status = status;
    p.trace = nullptr;
    p.trace_cta = sqd_opt(SQD_OPT_F16_TRACE_CTA);
#ifdef SQD_ENABLE_TRACE   // profiling builds only (SQD_BUILD_TRACE=1 python csrc/build.py): device address of a clock64 stamp buffer
    if (const char *e = getenv("SQD_F16_TRACE")) p.trace = reinterpret_cast<long long *>(strtoull(e, nullptr, 0));
#endif
    switch (npad / 16) {
        case 1: return launch_f16<16>(maps, p, grid, st);
        case 2: return launch_f16<32>(maps, p, grid, st);
        case 3: return launch_f16<48>(maps, p, grid, st);
        case 4: return launch_f16<64>(maps, p, grid, st);
        case 5: return launch_f16<80>(maps, p, grid, st);
        case 6: return launch_f16<96>(maps, p, grid, st);
        case 7: return launch_f16<112>(maps, p, grid, st);
        case 8: return launch_f16<128>(maps, p, grid, st);
    }
    SQD_REQUIRE(false, SQD_E_SHAPE, "convdet (tcgen05): unsupported Cout %d", cout);
}

// a1: ConvDet 3x3 head as a persistent tcgen05 / TMEM implicit GEMM fed by TMA (sm_100a), 3xTF32.
// Reference: SqueezeDetBase.convdet + permute(0,2,3,1) + view, src/model/squeezedet.py:73-75,83-87
// (cuDNN conv with N=72 plus an NCHW->NHWC copy kernel there).
//
// GEMM view per image: M = gh*gw cells, N = Cout = K_anchors*(C+5) (72 KITTI, padded to 80), K = 9*Cin = 6912.
// The first version of this kernel (one TMA box per tap, pre-split hi/lo planes) was L2-bandwidth bound:
// 11.2 MB of TMA traffic per 128-cell tile, 10.3 TB/s aggregate, tensor pipe ~1/3 busy (profiles/r01_*).
// This version cuts the traffic and balances the machine:
//
//  * M tile = 8 x 16 cells = 128 rows = one UMMA_M.  A "unit" of work is (tile, 32-channel block, dx):
//    ONE 4-D TMA box {32 ch, 16 x, 10 y, 1 img} at (c0, x0+dx, y0-1, b) of the RAW fp32 NHWC feature map.
//    Conv padding = TMA out-of-bounds zero fill.  The box lands as 160 rows x 128 B, 128B-swizzled; the three
//    dy taps are the SAME patch read through UMMA descriptors offset by 16 rows (2048 B, swizzle-atom
//    aligned), so A is fetched 3x per channel block instead of 9x (and once, not hi+lo twice).
//  * 3xTF32 split in the kernel: four converter warps turn the raw patch into tf32-exact hi (in place) and
//    lo planes in shared memory (hi = rna_tf32(x), lo = rna_tf32(x - hi)); D += A_lo*B_hi + A_hi*B_lo +
//    A_hi*B_hi.  channels_last features are consumed zero-copy; there is no hi/lo pre-pass any more.
//  * B = packed weights [Npad][9*Cin] K-major hi/lo planes (k = tap*Cin + c), its own TMA ring of
//    {32, Npad} boxes, decoupled from the A ring (A stages live for 3 taps, B stages for one).
//  * Chunked accumulation: the tensor core truncates when adding into the fp32 TMEM accumulator (measured:
//    -2e-5 relative bias over 2592 MMAs, profiles/r01_tc_accuracy_vs_chunk.txt), so every unit (36 MMAs)
//    accumulates from zero into one of two TMEM accumulators and four accumulate warps add finished units
//    into fp32 registers with round-to-nearest while the next unit's MMAs run.
//  * Persistent, balanced schedule: grid = min(#SMs, #tiles); the unit range is cut evenly, so a CTA owns
//    [tail of a tile][whole tiles][head of a tile].  A split tile is finished deterministically: the head
//    holder publishes its partial sums, the tail holder (higher CTA index, its tail segment is processed
//    LAST) adds them in a fixed order.  Waiters only ever wait for lower-indexed CTAs.
//  * Warp roles (352 threads): 0 A-TMA, 1 TMEM alloc + MMA issue (one lane), 2 B-TMA, 3..6 accumulate +
//    epilogue (+bias -> pred in the reference's (B, A, C+5) layout), 7..10 hi/lo converters.
//  * Every wait is bounded: on timeout the CTA raises a status word and drains instead of hanging.
// Algorithmic FLOPs per image: 2*M*Cout*K (the 3 passes and the N padding are NOT counted).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace sqd_tc;

constexpr int kTileX = 16, kTileY = 8, kPatchY = kTileY + 2;
constexpr int kBlockK = 32;   // channels per unit (128 B of fp32 = one swizzle row)
constexpr int kUmmaK = 8;     // tf32 MMA K
constexpr int kPatchBytes = kPatchY * kTileX * kBlockK * 4;  // 20480: one plane of one A stage
constexpr int kDyBytes = kTileX * kBlockK * 4;               // 2048: one y row of the patch = descriptor step per dy
constexpr int kAStages = 2;
constexpr int kThreads = 352;
constexpr int kWarpATma = 0, kWarpMma = 1, kWarpBTma = 2, kWarpAcc0 = 3, kWarpCvt0 = 7;

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
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

// NCHW (B,Cin,P) -> NHWC (B,P,Cin) raw fp32 through a 32x33 shared tile; P = gh*gw
__global__ void nchw_to_nhwc_kernel(const float *__restrict__ in, float *__restrict__ out, int cin, int P) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const float *src = in + (size_t)b * cin * P;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, p = p0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < cin && p < P) ? __ldg(src + (size_t)c * P + p) : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int p = p0 + j, c = c0 + threadIdx.x;
        if (p < P && c < cin) out[((size_t)b * P + p) * cin + c] = tile[threadIdx.x][j];
    }
}

struct Tc2Params {
    int cin, gh, gw, cout;
    int tiles_x, tiles_per_img, total_tiles;
    int upt;            // units per tile = (cin/32) * 3
    int units_per_cta;  // even cut of total_tiles*upt over the grid (>= upt)
    int b_stages;
    const float *bias;
    float *pred;
    float *partial;  // (grid, 128, NPAD) partial sums of split tiles
    int *flags;      // (grid) 1 = partial[cta] published
    int *status;     // 0 ok; else the role whose bounded wait timed out
    long long *trace;  // debug: per-unit clock64 timestamps of CTA 0 (8 slots per unit), or NULL
};

#define SQD_TRACE(slot, i) \
    do { if (p.trace && cta == 0 && lane == 0 && (i) < 512) p.trace[(i) * 8 + (slot)] = clock64(); } while (0)

struct Sched {  // the permuted unit sequence of one CTA: [whole tiles + head segment][deferred tail segment]
    long long u0;
    int n, main_len, upt;
    __device__ __forceinline__ long long unit(int i) const {
        const int len_tail = n - main_len;
        return i < main_len ? u0 + len_tail + i : u0 + (i - main_len);
    }
};

template <int NPAD>
__global__ void __launch_bounds__(kThreads, 1)
convdet_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b_hi,
                   const __grid_constant__ CUtensorMap map_b_lo, const Tc2Params p) {
    constexpr int kBBytes = NPAD * kBlockK * 4;       // one weight plane of one tap
    constexpr int kBStageBytes = 2 * kBBytes;         // hi + lo
    constexpr int kAStageBytes = 2 * kPatchBytes;     // raw->hi, lo
    constexpr uint32_t kAccStride = NPAD <= 16 ? 16 : (NPAD <= 32 ? 32 : (NPAD <= 64 ? 64 : 128));
    constexpr uint32_t kTmemCols = 2 * kAccStride;
    constexpr uint32_t kIdesc = umma_idesc_tf32(128, NPAD);

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int BS = p.b_stages;
    uint8_t *a_ring = smem;
    uint8_t *b_ring = smem + kAStages * kAStageBytes;
    uint8_t *ctrl = b_ring + (size_t)BS * kBStageBytes;
    uint64_t *a_full = reinterpret_cast<uint64_t *>(ctrl);  // [2]  TMA landed the raw patch
    uint64_t *a_conv = a_full + 2;                          // [2]  hi/lo planes written (128 arrivals)
    uint64_t *a_empty = a_conv + 2;                         // [2]  MMAs done reading the stage
    uint64_t *b_full = a_empty + 2;                         // [8]
    uint64_t *b_empty = b_full + 8;                         // [8]
    uint64_t *tmem_full = b_empty + 8;                      // [2]
    uint64_t *tmem_empty = tmem_full + 2;                   // [2]
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
        for (int s = 0; s < kAStages; ++s) {
            mbar_init(a_full + s, 1);
            mbar_init(a_conv + s, 128);
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
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
    }
    for (int i = threadIdx.x; i < NPAD; i += kThreads) s_bias[i] = i < p.cout ? __ldg(p.bias + i) : 0.f;
    if (warp == kWarpMma) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kWarpATma) {
        // ===== A producer: one raw patch per unit (warp stays converged, one elected lane issues) =====
        for (int i = 0; i < n_units; ++i) {
            const int s = i % kAStages;
            const uint32_t ph = (uint32_t)(i / kAStages) & 1u;
            if (!mbar_wait_warp(a_empty + s, ph ^ 1u, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 1);
                break;
            }
            const long long u = sc.unit(i);
            const int tile = (int)(u / p.upt), r = (int)(u % p.upt);
            const int cb = r / 3, dxi = r - cb * 3;
            const int img = tile / p.tiles_per_img, t = tile - img * p.tiles_per_img;
            const int x0 = (t % p.tiles_x) * kTileX, y0 = (t / p.tiles_x) * kTileY;
            SQD_TRACE(0, i);
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(a_full + s, kPatchBytes);
                tma_load_4d(&map_a, a_full + s, a_ring + (size_t)s * kAStageBytes, cb * kBlockK, x0 + dxi - 1, y0 - 1, img);
            }
            __syncwarp();
        }
    } else if (warp == kWarpBTma) {
        // ===== B producer: hi + lo weight tiles of one tap per step =====
        bool ok = true;
        for (int i = 0; i < n_units && ok; ++i) {
            const long long u = sc.unit(i);
            const int r = (int)(u % p.upt);
            const int cb = r / 3, dxi = r - cb * 3;
            for (int dyi = 0; dyi < 3; ++dyi) {
                const int j = i * 3 + dyi;
                const int s = j % BS;
                const uint32_t ph = (uint32_t)(j / BS) & 1u;
                if (!mbar_wait_warp(b_empty + s, ph ^ 1u, abort_flag)) {
                    if (lane == 0) atomicCAS(p.status, 0, 5);
                    ok = false;
                    break;
                }
                const int tap = dyi * 3 + dxi;
                uint8_t *st = b_ring + (size_t)s * kBStageBytes;
                if (dyi == 0) SQD_TRACE(7, i);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(b_full + s, kBStageBytes);
                    tma_load_2d(&map_b_hi, b_full + s, st, tap * p.cin + cb * kBlockK, 0);
                    tma_load_2d(&map_b_lo, b_full + s, st + kBBytes, tap * p.cin + cb * kBlockK, 0);
                }
                __syncwarp();
            }
        }
    } else if (warp == kWarpMma) {
        // ===== MMA issuer: 36 MMAs per unit into a fresh TMEM accumulator.  The warp stays converged and one
        // elected lane issues, so descriptors live in uniform registers (no per-MMA R2UR retry loop). =====
        bool ok = true;
        for (int i = 0; i < n_units && ok; ++i) {
            const int as = i % kAStages;
            const uint32_t a_ph = (uint32_t)(i / kAStages) & 1u;
            const int buf = i & 1;
            const uint32_t acc_ph = (uint32_t)(i >> 1) & 1u;
            if (!mbar_wait_warp(tmem_empty + buf, acc_ph ^ 1u, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 4);
                break;
            }
            if (!mbar_wait_warp(a_conv + as, a_ph, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 2);
                break;
            }
            tc_fence_after();
            SQD_TRACE(3, i);
            const uint32_t d_tmem = tmem_base + (uint32_t)buf * kAccStride;
            const uint32_t a_addr = smem_u32(a_ring + (size_t)as * kAStageBytes);
            for (int dyi = 0; dyi < 3; ++dyi) {
                const int j = i * 3 + dyi;
                const int bs = j % BS;
                const uint32_t b_ph = (uint32_t)(j / BS) & 1u;
                if (!mbar_wait_warp(b_full + bs, b_ph, abort_flag)) {
                    if (lane == 0) atomicCAS(p.status, 0, 6);
                    ok = false;
                    break;
                }
                tc_fence_after();
                const uint32_t b_addr = smem_u32(b_ring + (size_t)bs * kBStageBytes);
                const uint64_t a_hi = umma_desc_sw128(a_addr + dyi * kDyBytes);
                const uint64_t a_lo = umma_desc_sw128(a_addr + kPatchBytes + dyi * kDyBytes);
                const uint64_t b_hi = umma_desc_sw128(b_addr), b_lo = umma_desc_sw128(b_addr + kBBytes);
                if (elect_one_sync()) {
#pragma unroll
                    for (int ks = 0; ks < kBlockK / kUmmaK; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * kUmmaK * 4) >> 4);  // +32 B per K step, in 16 B units
                        // small cross terms first, then the dominant hi*hi product; each unit starts from zero
                        umma_tf32(d_tmem, a_lo + adv, b_hi + adv, kIdesc, (dyi | ks) ? 1u : 0u);
                        umma_tf32(d_tmem, a_hi + adv, b_lo + adv, kIdesc, 1u);
                        umma_tf32(d_tmem, a_hi + adv, b_hi + adv, kIdesc, 1u);
                    }
                    umma_commit(b_empty + bs);  // weight slot reusable once these MMAs have read it
                }
                __syncwarp();
            }
            SQD_TRACE(4, i);
            if (elect_one_sync()) {
                umma_commit(a_empty + as);      // patch slot reusable
                umma_commit(tmem_full + buf);   // unit complete (also fires after an aborted tap loop)
            }
            __syncwarp();
        }
    } else if (warp >= kWarpCvt0) {
        // ===== converters: raw fp32 patch -> tf32 hi (in place) + lo, same swizzled positions =====
        const int t = threadIdx.x - kWarpCvt0 * 32;  // 0..127
        for (int i = 0; i < n_units; ++i) {
            const int s = i % kAStages;
            const uint32_t ph = (uint32_t)(i / kAStages) & 1u;
            if (!mbar_wait(a_full + s, ph, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 7);
                break;
            }
            if (warp == kWarpCvt0) SQD_TRACE(1, i);
            float4 *hi = reinterpret_cast<float4 *>(a_ring + (size_t)s * kAStageBytes);
            float4 *lo = reinterpret_cast<float4 *>(a_ring + (size_t)s * kAStageBytes + kPatchBytes);
#pragma unroll
            for (int k = 0; k < kPatchBytes / 16 / 128; ++k) {
                const int idx = k * 128 + t;
                const float4 v = hi[idx];
                float4 h, l;
                h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
                l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
                hi[idx] = h;
                lo[idx] = l;
            }
            fence_proxy_async();        // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(a_conv + s);
            if (warp == kWarpCvt0) SQD_TRACE(2, i);
        }
    } else {
        // ===== accumulate + epilogue warps: TMEM unit -> fp32 registers (RN) ... -> (+bias) -> pred =====
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // accumulator row == cell inside the 8x16 tile
        const int et = threadIdx.x - kWarpAcc0 * 32;  // 0..127
        float acc[NPAD];
        bool ok = true;
        int seg_r0 = 0;
        for (int i = 0; i < n_units; ++i) {
            const long long u = sc.unit(i);
            const int tile = (int)(u / p.upt), r = (int)(u % p.upt);
            if (i == 0 || r == 0 || i == sc.main_len) {
                seg_r0 = r;
#pragma unroll
                for (int n = 0; n < NPAD; ++n) acc[n] = 0.f;
            }
            const int buf = i & 1;
            const uint32_t acc_ph = (uint32_t)(i >> 1) & 1u;
            if (!mbar_wait(tmem_full + buf, acc_ph, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 3);
                ok = false;
                break;
            }
            tc_fence_after();
            __syncwarp();
            if (warp == kWarpAcc0) SQD_TRACE(5, i);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * kAccStride;
#pragma unroll
            for (int n0 = 0; n0 < NPAD; n0 += 16) {
                uint32_t v[16];
                tmem_ld_x16(taddr + n0, v);
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 16; ++k) acc[n0 + k] += __uint_as_float(v[k]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + buf);  // this warp is done reading the accumulator
            if (warp == kWarpAcc0) SQD_TRACE(6, i);

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
                    acc[n] = h.x + acc[n]; acc[n + 1] = h.y + acc[n + 1]; acc[n + 2] = h.z + acc[n + 2]; acc[n + 3] = h.w + acc[n + 3];
                }
            } else if (!(from_start && to_end)) {
                if (lane == 0) atomicCAS(p.status, 0, 9);  // a segment strictly inside a tile: scheduler invariant broken
                continue;
            }
            // whole tile in registers: + bias -> pred
            const int img = tile / p.tiles_per_img, t = tile - img * p.tiles_per_img;
            const int x = (t % p.tiles_x) * kTileX + row % kTileX, y = (t / p.tiles_x) * kTileY + row / kTileX;
            if (y < p.gh && x < p.gw) {
                float *out = p.pred + (((size_t)img * p.gh + y) * p.gw + x) * p.cout;
                if ((p.cout & 3) == 0) {
                    float4 *o4 = reinterpret_cast<float4 *>(out);
#pragma unroll
                    for (int n = 0; n < NPAD; n += 4)
                        if (n < p.cout)
                            o4[n >> 2] = make_float4(acc[n] + s_bias[n], acc[n + 1] + s_bias[n + 1],
                                                     acc[n + 2] + s_bias[n + 2], acc[n + 3] + s_bias[n + 3]);
                } else {
#pragma unroll
                    for (int n = 0; n < NPAD; ++n)
                        if (n < p.cout) out[n] = acc[n] + s_bias[n];
                }
            }
        }
        (void)ok;
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

constexpr size_t kSmemLimit = 227 * 1024;
constexpr size_t kCtrlBytes = 1024;

int b_stages_for(int npad) {
    const size_t stage = (size_t)2 * npad * kBlockK * 4;
    size_t s = (kSmemLimit - 1024 /*align*/ - kCtrlBytes - (size_t)kAStages * 2 * kPatchBytes) / stage;
    if (s > 6) s = 6;
    return (int)s;
}

size_t smem_bytes_for(int npad, int b_stages) {
    return 1024 + (size_t)kAStages * 2 * kPatchBytes + (size_t)b_stages * 2 * npad * kBlockK * 4 + kCtrlBytes;
}

int grid_for(int total_tiles) { return total_tiles < SQD_SM_COUNT ? total_tiles : SQD_SM_COUNT; }

// workspace layout: [status (256 B)][flags: 256 ints][partials: grid*128*npad floats][NHWC copy when input is NCHW]
struct WsLayout {
    size_t flags_off, partial_off, nhwc_off, total;
};
WsLayout ws_layout(int batch, int cin, int gh, int gw, int cout, int layout) {
    WsLayout w;
    w.flags_off = 256;
    w.partial_off = w.flags_off + 256 * sizeof(int);
    w.nhwc_off = w.partial_off + (size_t)SQD_SM_COUNT * 128 * npad_of(cout) * sizeof(float);
    w.total = w.nhwc_off + (layout == SQD_LAYOUT_NCHW ? (size_t)batch * gh * gw * cin * sizeof(float) : 0);
    return w;
}

template <int NPAD>
int launch_tc2(const CUtensorMap *maps, const Tc2Params &p, int grid, cudaStream_t st) {
    const size_t smem = smem_bytes_for(NPAD, p.b_stages);
    SQD_CUDA(cudaFuncSetAttribute(convdet_tc2_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    convdet_tc2_kernel<NPAD><<<grid, kThreads, smem, st>>>(maps[0], maps[1], maps[2], p);
    SQD_LAUNCH_CHECK("convdet_tc2_kernel");
    return SQD_OK;
}

}  // namespace

size_t sqd_tc2_workspace_bytes(int batch, int cin, int gh, int gw, int cout, int layout) {
    return ws_layout(batch, cin, gh, gw, cout, layout).total;
}

int sqd_convdet_tc2(const float *d_feat, int layout, const void *d_packed, const float *d_bias, int batch, int cin,
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

    // 1. NCHW input: one transposing copy to NHWC (channels_last input is consumed in place)
    const float *nhwc = d_feat;
    if (layout == SQD_LAYOUT_NCHW) {
        float *copy = reinterpret_cast<float *>(ws + w.nhwc_off);
        const int P = gh * gw;
        dim3 grid((P + 31) / 32, (cin + 31) / 32, batch);
        nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, st>>>(d_feat, copy, cin, P);
        SQD_LAUNCH_CHECK("nchw_to_nhwc_kernel");
        nhwc = copy;
    }

    // 2. tensor maps
    alignas(64) CUtensorMap maps[3];
    {
        const cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)gw, (cuuint64_t)gh, (cuuint64_t)batch};
        const cuuint64_t strides[3] = {(cuuint64_t)cin * 4, (cuuint64_t)gw * cin * 4, (cuuint64_t)gh * gw * cin * 4};
        const cuuint32_t box[4] = {kBlockK, kTileX, kPatchY, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&maps[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(nhwc), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SQD_REQUIRE(r == CUDA_SUCCESS, SQD_E_DRIVER, "cuTensorMapEncodeTiled(features) failed: CUresult %d", (int)r);
    }
    {
        const size_t ktot = (size_t)9 * cin;
        const float *b_hi = static_cast<const float *>(d_packed);
        const float *b_lo = b_hi + (size_t)npad * ktot;
        const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)npad};
        const cuuint64_t strides[1] = {(cuuint64_t)ktot * 4};
        const cuuint32_t box[2] = {kBlockK, (cuuint32_t)npad};
        const cuuint32_t estr[2] = {1, 1};
        for (int i = 0; i < 2; ++i) {
            CUresult r = encode(&maps[1 + i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, i == 0 ? (void *)b_hi : (void *)b_lo,
                                dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            SQD_REQUIRE(r == CUDA_SUCCESS, SQD_E_DRIVER, "cuTensorMapEncodeTiled(weights) failed: CUresult %d", (int)r);
        }
    }

    // 3. the persistent GEMM
    Tc2Params p;
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
    p.b_stages = b_stages_for(npad);
    p.bias = d_bias;
    p.pred = d_pred;
    p.partial = partial;
    p.flags = flags;
    p.status = status;
    p.trace = nullptr;
    if (const char *e = getenv("SQD_TC_TRACE")) p.trace = reinterpret_cast<long long *>(strtoull(e, nullptr, 0));  // debug
    switch (npad / 16) {
        case 1: return launch_tc2<16>(maps, p, grid, st);
        case 2: return launch_tc2<32>(maps, p, grid, st);
        case 3: return launch_tc2<48>(maps, p, grid, st);
        case 4: return launch_tc2<64>(maps, p, grid, st);
        case 5: return launch_tc2<80>(maps, p, grid, st);
        case 6: return launch_tc2<96>(maps, p, grid, st);
        case 7: return launch_tc2<112>(maps, p, grid, st);
        case 8: return launch_tc2<128>(maps, p, grid, st);
    }
    SQD_REQUIRE(false, SQD_E_SHAPE, "convdet (tcgen05): unsupported Cout %d", cout);
}

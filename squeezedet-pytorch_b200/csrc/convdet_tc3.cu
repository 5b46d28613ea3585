// a1: ConvDet 3x3 head as a persistent tcgen05 / TMEM implicit GEMM fed by TMA (sm_100a), 3xTF32.
// Reference: SqueezeDetBase.convdet + permute(0,2,3,1) + view, src/model/squeezedet.py:73-75,83-87
// (cuDNN conv with N=72 plus an NCHW->NHWC copy kernel there).
//
// GEMM view per image: M = gh*gw cells, N = Cout = K_anchors*(C+5) (72 KITTI, padded to 80), K = 9*Cin = 6912.
//
// What bounds this kernel (all measured on B200, profiles/r01_*):
//   v1 (one TMA box per tap, pre-split hi/lo planes, 3 SS-mode MMAs per K step) was L2 bound: 11.2 MB of TMA
//   traffic per 128-cell tile.  v2 (patch reuse + in-kernel split) was SHARED-MEMORY bound: an SS-mode
//   tcgen05.mma reads A (4 KB) and B (N*32 B) through the same 128 B/clk port as TMA writes and the hi/lo
//   converters (micro-benchmark tools/micro/umma_rate.cu: SS = 32 + N/4 cycles, TS = N/2 cycles).
// So this version keeps A OUT of shared memory on the MMA side:
//
//  * M tile = 8 x 16 cells = 128 rows = one UMMA_M.  A "unit" of work is (tile, 32-channel block, dx):
//    ONE 4-D TMA box {32 ch, 16 x, 10 y, 1 img} at (c0, x0+dx, y0-1, b) of the RAW fp32 NHWC feature map
//    (conv padding = TMA out-of-bounds zero fill; channels_last features are consumed zero-copy).  The three
//    dy taps reuse the patch, so A is fetched 3x per channel block instead of 9x.
//  * 3xTF32 split in the kernel, straight into TENSOR MEMORY: converter thread m reads patch row m+16*dy
//    (un-swizzling the 128 B row), forms tf32-exact hi = rna(x) and lo = rna(x - hi) and writes them with
//    tcgen05.st into a 64-column A slot (one slot per dy).  The MMAs take A from TMEM (TS mode, N/2 cycles).
//  * Two MMAs per K step instead of three: B_hi and B_lo tiles are adjacent in shared memory, so
//    D[:, 0:2N] (+)= A_hi * [B_hi | B_lo]^T is ONE N=2*Npad MMA, then D[:, 0:N] += A_lo * B_hi^T.
//    The epilogue adds the two halves.  A_hi is read once, B_hi/B_lo traffic is unchanged.
//  * B = packed weights [Npad][9*Cin] K-major hi/lo planes (k = tap*Cin + c), its own TMA ring.
//  * Chunked accumulation: the tensor core truncates when adding into the fp32 TMEM accumulator (measured:
//    -2e-5 relative bias over 2592 MMAs, profiles/r01_tc_accuracy_vs_chunk.txt), so every unit (24 MMAs)
//    accumulates from zero into a TMEM accumulator and four accumulate warps add finished units into fp32
//    registers with round-to-nearest while the next unit's MMAs run into the other accumulator.
//  * Persistent, balanced schedule: grid = min(#SMs, #tiles); the unit range is cut evenly, so a CTA owns
//    [tail of a tile][whole tiles][head of a tile].  A split tile is finished deterministically: the head
//    holder publishes its partial sums, the tail holder (higher CTA index, its tail segment is processed
//    LAST) adds them in a fixed order.  Waiters only ever wait for lower-indexed CTAs.
//  * Warp roles (352 threads): 0 A-TMA, 1 TMEM alloc + MMA issue (converged warp, one elected lane), 2 B-TMA,
//    3..6 accumulate + epilogue (+bias -> pred in the reference's (B, A, C+5) layout), 7..10 converters.
//  * Every wait is bounded: on timeout the CTA raises a status word and drains instead of hanging.
// Algorithmic FLOPs per image: 2*M*Cout*K (the 3xTF32 passes and the N padding are NOT counted).
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
constexpr int kASlotCols = 64;  // one TMEM A slot: 32 columns of hi + 32 of lo (K = 32)
constexpr int kThreads = 352;
constexpr int kWarpATma = 0, kWarpMma = 1, kWarpBTma = 2, kWarpAcc0 = 3, kWarpCvt0 = 7;

// round-to-nearest (ties away) to tf32 with integer ops (== cvt.rna.tf32.f32 for finite inputs): full-rate ALU
__device__ __forceinline__ uint32_t tf32_rna_bits(uint32_t b) { return (b + 0x1000u) & 0xffffe000u; }

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T, M=128, kind::tf32 (A: lane = row, column = k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
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

struct Tc3Params {
    int cin, gh, gw, cout;
    int tiles_x, tiles_per_img, total_tiles;
    int upt;            // units per tile = (cin/32) * 3
    int units_per_cta;  // even cut of total_tiles*upt over the grid (>= upt)
    int b_stages, raw_stages;
    const float *bias;
    float *pred;
    float *partial;  // (grid, 128, NPAD) partial sums of split tiles
    int *flags;      // (grid) 1 = partial[cta] published
    int *status;     // 0 ok; else the role whose bounded wait timed out
    long long *trace;  // debug: per-unit clock64 timestamps of CTA 0 (8 slots per unit), or NULL
};

#define SQD_TRACE(slot, i) \
    do { if (p.trace && cta == 0 && lane == 0 && (i) < 512) p.trace[(i) * 32 + (slot)] = clock64(); } while (0)

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
convdet_tc3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b_hi,
                   const __grid_constant__ CUtensorMap map_b_lo, const Tc3Params p) {
    constexpr int kBBytes = NPAD * kBlockK * 4;       // one weight plane of one tap
    constexpr int kBStageBytes = 2 * kBBytes;         // hi rows then lo rows: also ONE K-major tile of 2*NPAD rows
    constexpr int kAccCols = 2 * NPAD;                // [A_hi*B_hi + A_lo*B_hi | A_hi*B_lo]
    constexpr int kAccBufs = (2 * kAccCols + 3 * kASlotCols <= 512) ? 2 : 1;
    constexpr uint32_t kASlotBase = kAccBufs * kAccCols;
    constexpr uint32_t kTmemCols = 512;
    constexpr uint32_t kIdescCat = umma_idesc_tf32(128, 2 * NPAD);
    constexpr uint32_t kIdescOne = umma_idesc_tf32(128, NPAD);
    static_assert(kASlotBase + 3 * kASlotCols <= 512, "TMEM budget");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int BS = p.b_stages, RS = p.raw_stages;
    uint8_t *a_ring = smem;                                   // RS raw patches
    uint8_t *b_ring = smem + (size_t)RS * kPatchBytes;
    uint8_t *ctrl = b_ring + (size_t)BS * kBStageBytes;
    uint64_t *raw_full = reinterpret_cast<uint64_t *>(ctrl);  // [4]  TMA landed the raw patch
    uint64_t *raw_empty = raw_full + 4;                       // [4]  converters done reading it (128 arrivals)
    uint64_t *aslot_full = raw_empty + 4;                     // [3]  hi/lo of one dy written to TMEM (128 arrivals)
    uint64_t *aslot_empty = aslot_full + 3;                   // [3]  MMAs done reading the slot
    uint64_t *b_full = aslot_empty + 3;                       // [8]
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
        for (int s = 0; s < RS; ++s) {
            mbar_init(raw_full + s, 1);
            mbar_init(raw_empty + s, 128);
        }
        for (int s = 0; s < 3; ++s) {
            mbar_init(aslot_full + s, 128);
            mbar_init(aslot_empty + s, 1);
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
            const int s = i % RS;
            const uint32_t ph = (uint32_t)(i / RS) & 1u;
            if (!mbar_wait_warp(raw_empty + s, ph ^ 1u, abort_flag)) {
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
                mbar_arrive_expect_tx(raw_full + s, kPatchBytes);
                tma_load_4d(&map_a, raw_full + s, a_ring + (size_t)s * kPatchBytes, cb * kBlockK, x0 + dxi - 1, y0 - 1, img);
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
        // ===== MMA issuer: 24 TS-mode MMAs per unit into a fresh TMEM accumulator.  The warp stays converged and
        // one elected lane issues, so descriptors live in uniform registers (no per-MMA R2UR retry loop). =====
        bool ok = true;
        for (int i = 0; i < n_units && ok; ++i) {
            const int buf = kAccBufs == 2 ? (i & 1) : 0;
            const uint32_t acc_ph = kAccBufs == 2 ? ((uint32_t)(i >> 1) & 1u) : ((uint32_t)i & 1u);
            if (!mbar_wait_warp(tmem_empty + buf, acc_ph ^ 1u, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 4);
                break;
            }
            SQD_TRACE(3, i);
            const uint32_t d_tmem = tmem_base + (uint32_t)buf * kAccCols;
            for (int dyi = 0; dyi < 3; ++dyi) {
                const int j = i * 3 + dyi;
                const int bs = j % BS;
                const uint32_t b_ph = (uint32_t)(j / BS) & 1u;
                if (!mbar_wait_warp(aslot_full + dyi, (uint32_t)i & 1u, abort_flag)) {
                    if (lane == 0) atomicCAS(p.status, 0, 2);
                    ok = false;
                    break;
                }
                SQD_TRACE(16 + dyi * 3 + 0, i);   // A slot ready
                if (!mbar_wait_warp(b_full + bs, b_ph, abort_flag)) {
                    if (lane == 0) atomicCAS(p.status, 0, 6);
                    ok = false;
                    break;
                }
                SQD_TRACE(16 + dyi * 3 + 1, i);   // B stage ready
                tc_fence_after();
                const uint32_t b_addr = smem_u32(b_ring + (size_t)bs * kBStageBytes);
                const uint64_t b_cat = umma_desc_sw128(b_addr);  // 2*NPAD rows: B_hi then B_lo
                const uint32_t a_hi = tmem_base + kASlotBase + (uint32_t)dyi * kASlotCols, a_lo = a_hi + 32;
                if (elect_one_sync()) {
#pragma unroll
                    for (int ks = 0; ks < kBlockK / kUmmaK; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * kUmmaK * 4) >> 4);  // +32 B per K step, in 16 B units
                        umma_tf32_ts(d_tmem, a_hi + ks * kUmmaK, b_cat + adv, kIdescCat, (dyi | ks) ? 1u : 0u);
                        umma_tf32_ts(d_tmem, a_lo + ks * kUmmaK, b_cat + adv, kIdescOne, 1u);
                    }
                    umma_commit(b_empty + bs);        // weight slot reusable once these MMAs have read it
                    umma_commit(aslot_empty + dyi);   // and the TMEM A slot
                }
                __syncwarp();
                SQD_TRACE(16 + dyi * 3 + 2, i);   // step issued
            }
            SQD_TRACE(4, i);
            if (elect_one_sync()) umma_commit(tmem_full + buf);  // unit complete (also fires after an aborted tap loop)
            __syncwarp();
        }
    } else if (warp >= kWarpCvt0) {
        // ===== converters: raw fp32 patch row (swizzled smem) -> tf32 hi / lo -> TMEM A slot of each dy =====
        const int q = warp & 3;          // TMEM lane quarter of this warp
        const int m = q * 32 + lane;     // tile row == TMEM lane
        bool ok = true;
        for (int i = 0; i < n_units && ok; ++i) {
            const int s = i % RS;
            const uint32_t ph = (uint32_t)(i / RS) & 1u;
            if (!mbar_wait(raw_full + s, ph, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 7);
                break;
            }
            if (warp == kWarpCvt0) SQD_TRACE(1, i);
            const uint8_t *raw = a_ring + (size_t)s * kPatchBytes;
            for (int dyi = 0; dyi < 3; ++dyi) {
                const int pr = m + 16 * dyi;  // patch row holding the input of tile row m for this dy
                const uint8_t *rowp = raw + pr * 128;
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {  // 16-byte chunk c of the row sits at chunk position c ^ (row & 7)
                    const float4 v = *reinterpret_cast<const float4 *>(rowp + ((c ^ (pr & 7)) << 4));
                    const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t h = tf32_rna_bits(__float_as_uint(x[e]));
                        hi[c * 4 + e] = h;
                        lo[c * 4 + e] = tf32_rna_bits(__float_as_uint(x[e] - __uint_as_float(h)));
                    }
                }
                if (warp == kWarpCvt0) SQD_TRACE(8 + dyi * 3 + 0, i);
                if (!mbar_wait_warp(aslot_empty + dyi, ((uint32_t)i & 1u) ^ 1u, abort_flag)) {  // MMAs of unit i-1 done
                    if (lane == 0) atomicCAS(p.status, 0, 10);
                    ok = false;
                    break;
                }
                tc_fence_after();
                if (warp == kWarpCvt0 && dyi < 2) SQD_TRACE(8 + dyi * 3 + 1, i);
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kASlotBase + (uint32_t)dyi * kASlotCols;
                tmem_st_x16(taddr, hi);
                tmem_st_x16(taddr + 16, hi + 16);
                tmem_st_x16(taddr + 32, lo);
                tmem_st_x16(taddr + 48, lo + 16);
                tmem_st_wait();
                if (warp == kWarpCvt0 && dyi < 2) SQD_TRACE(8 + dyi * 3 + 2, i);
                tc_fence_before();
                mbar_arrive(aslot_full + dyi);
            }
            mbar_arrive(raw_empty + s);  // all three rows of this thread have been read
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
            const int buf = kAccBufs == 2 ? (i & 1) : 0;
            const uint32_t acc_ph = kAccBufs == 2 ? ((uint32_t)(i >> 1) & 1u) : ((uint32_t)i & 1u);
            if (!mbar_wait(tmem_full + buf, acc_ph, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 3);
                ok = false;
                break;
            }
            tc_fence_after();
            __syncwarp();
            if (warp == kWarpAcc0) SQD_TRACE(5, i);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * kAccCols;
#pragma unroll
            for (int n0 = 0; n0 < NPAD; n0 += 16) {
                uint32_t v[16], w[16];
                tmem_ld_x16(taddr + n0, v);          // A_hi*B_hi + A_lo*B_hi
                tmem_ld_x16(taddr + NPAD + n0, w);   // A_hi*B_lo
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 16; ++k) acc[n0 + k] += __uint_as_float(v[k]) + __uint_as_float(w[k]);
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

int raw_stages_for(int npad) { return npad <= 80 ? 4 : 3; }

int b_stages_for(int npad) {
    const size_t stage = (size_t)2 * npad * kBlockK * 4;
    size_t s = (kSmemLimit - 1024 /*align*/ - kCtrlBytes - (size_t)raw_stages_for(npad) * kPatchBytes) / stage;
    if (s > 6) s = 6;
    return (int)s;
}

size_t smem_bytes_for(int npad, int b_stages) {
    return 1024 + (size_t)raw_stages_for(npad) * kPatchBytes + (size_t)b_stages * 2 * npad * kBlockK * 4 + kCtrlBytes;
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
int launch_tc3(const CUtensorMap *maps, const Tc3Params &p, int grid, cudaStream_t st) {
    const size_t smem = smem_bytes_for(NPAD, p.b_stages);
    SQD_CUDA(cudaFuncSetAttribute(convdet_tc3_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    convdet_tc3_kernel<NPAD><<<grid, kThreads, smem, st>>>(maps[0], maps[1], maps[2], p);
    SQD_LAUNCH_CHECK("convdet_tc3_kernel");
    return SQD_OK;
}

}  // namespace

size_t sqd_tc3_workspace_bytes(int batch, int cin, int gh, int gw, int cout, int layout) {
    return ws_layout(batch, cin, gh, gw, cout, layout).total;
}

int sqd_convdet_tc3(const float *d_feat, int layout, const void *d_packed, const float *d_bias, int batch, int cin,
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
    Tc3Params p;
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
    p.raw_stages = raw_stages_for(npad);
    p.bias = d_bias;
    p.pred = d_pred;
    p.partial = partial;
    p.flags = flags;
    p.status = status;
    p.trace = nullptr;
    if (const char *e = getenv("SQD_TC_TRACE")) p.trace = reinterpret_cast<long long *>(strtoull(e, nullptr, 0));  // debug
    switch (npad / 16) {
        case 1: return launch_tc3<16>(maps, p, grid, st);
        case 2: return launch_tc3<32>(maps, p, grid, st);
        case 3: return launch_tc3<48>(maps, p, grid, st);
        case 4: return launch_tc3<64>(maps, p, grid, st);
        case 5: return launch_tc3<80>(maps, p, grid, st);
        case 6: return launch_tc3<96>(maps, p, grid, st);
        case 7: return launch_tc3<112>(maps, p, grid, st);
        case 8: return launch_tc3<128>(maps, p, grid, st);
    }
    SQD_REQUIRE(false, SQD_E_SHAPE, "convdet (tcgen05): unsupported Cout %d", cout);
}

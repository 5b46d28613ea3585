// a1: ConvDet 3x3 head, ONE kernel from the fp32 NCHW Fire11 features to pred -- no pre-pass, no fp16 planes in HBM.
// Reference: SqueezeDetBase.convdet + permute(0,2,3,1) + view, src/model/squeezedet.py:73-75,83-87.
//
// Same arithmetic idea as convdet_f16.cu ("f16x3": x*s = x1 + x2/2^11 with a power-of-two scale s, three half-precision
// tcgen05 passes, fp32 accumulation chunked per 64-channel block), but the operand pipeline is different:
//
//  * The features are read ONCE, as fp32, straight from the caller's NCHW tensor by CONVERTER warps (coalesced 4-byte
//    16-byte loads).  A thread keeps its share of one (tile, channel block) patch in registers (72 floats) and writes the
//    scaled two-term fp16 split directly as the K-major, 128B-swizzled A operand (two planes) that tcgen05.mma reads;
//    the registers of a finished round are refilled with the NEXT block at once, so loads stream continuously.
//    The scale per (image, 64-channel block) comes from a read-only max pass (absmax_kernel) in front of the GEMM --
//    the accumulate warps fold each TMEM chunk (= one channel block of one tile) into fp32 registers with the exact
//    factor 1/s.  No fp16 planes ever exist in HBM: DRAM traffic per image = the features (twice at most; the second
//    read hits L2 when the batch fits) + pred, against features + 2 x planes written + planes read for the staged path.
//  * M tile = 128 CONSECUTIVE cells of the image in a row-padded flat order (row pitch pw = gw + 1: one zero column
//    between image rows, which is the right pad of row y and the left pad of row y + 1; rows above / below the image are
//    zero too).  In that order a 3x3 tap is a constant offset (dy*pw + dx rows), so ALL NINE taps of a tile read the same
//    patch of R = 128 + 2*(pw + 1) rows through UMMA descriptors that start at different 128-byte rows (legal with
//    SWIZZLE_128B: the tensor core derives the swizzle phase from the absolute address, profiles/
//    r01_umma_row_offset_microtest.txt).  The patch is produced once per (tile, block) instead of once per dx tap.
//  * B = the packed weights of convdet_f16.cu (CTA-pair layout, [w1 | w2] rows of this CTA's HR channels), streamed by
//    TMA tap by tap through its own ring; per 16-channel K step two MMAs: D1 (+)= A1 x [w1|w2]^T, D2 (+)= A2 x w1^T.
//  * CTA pair (cta_group::2, M = 256): each CTA converts the patch of its own tile, the leader issues the MMAs.
//  * 512 threads, register budget rebalanced with setmaxnreg: warps 0-7 converters (128 registers), warps 8-11 accumulate +
//    epilogue (200), warp 12 B-TMA producer, warp 13 TMEM alloc + MMA issue (warps 14-15 idle; 56).
//  * Persistent, balanced schedule with deterministic split tiles, bounded waits and the status word: as convdet_f16.cu.
// Limits: NCHW fp32 input, Cin % 64 == 0, Cout <= 80, gh*gw % 4 == 0 (16-byte loads), gw <= 78 (R <= 288 rows: two operand buffers + the B ring fill the
// 227 KB of shared memory), gh*pw < 65536.  Other shapes take the staged path (pre-pass + convdet_f16_pair_kernel).
// Algorithmic FLOPs per image: 2*M*Cout*K; algorithmic bytes: Cin*P*4 read + P*Cout*4 written.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace sqd_tc;

constexpr int kBlockK = 64;            // channels per block = one 128-byte swizzle row of fp16
constexpr int kUmmaK = 16;
constexpr int kTileM = 128;
constexpr int kThreadsF = 512;
// warp ids: the scheduler favours high warp ids, so the latency-critical roles sit at the top
constexpr int kWarpCvt = 0, kWarpAcc = 8, kWarpB = 12, kWarpMma = 13, kWarpRelay = 14;
constexpr int kCvtThreads = 256, kCvtWarps = 8;
constexpr int kMaxRows = 288;          // patch rows; the converters cover 288 source pixels from an aligned start:
constexpr int kMainPix = 256;          //   256 as 64 quads x 8 channel octets (two rounds of 16-byte loads per thread)
                                       //   + 32 as single pixels x 8 octets (one round of 4-byte loads)
constexpr int kNB = 3;                 // B ring stages; a stage = the three dx taps of one dy row of one channel block
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;
constexpr int kHeaderBytes = 256;
constexpr size_t kSmemLimit = 227 * 1024;
constexpr size_t kCtrlBytes = 1024;
constexpr int kRegsLight = 56, kRegsAcc = 200;   // 128*56 + 128*200 + 256*128 (converters keep the launch value) = 65536

struct PackedHeader {   // must match convdet_f16.cu
    unsigned amax_bits;
    float scale;
    float inv_scale;
    int npad, cin;
    int hr;
};

// power-of-two scale s with amax*s in [2^13, 2^14) -- must match convdet_f16.cu
__device__ __forceinline__ float pow2_scale_for(float amax) {
    if (!(amax > 0.f) || amax > 3.0e38f) return 1.f;
    int ex;
    frexpf(amax, &ex);
    int e = 14 - ex;
    e = e < -126 ? -126 : (e > 126 ? 126 : e);
    return ldexpf(1.f, e);
}

__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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
__device__ __forceinline__ void acc_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 accumulate warps
__device__ __forceinline__ float ldg_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N)); }

// Profiling builds (SQD_BUILD_TRACE=1): cycle accumulators per role, dumped to p.trace[(cta*4 + role)*8 + k]
#ifdef SQD_ENABLE_TRACE
#define PROF_DECL long long pf_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pf_last = clock64()
#define PROF(k) do { const long long n_ = clock64(); pf_t[k] += n_ - pf_last; pf_last = n_; } while (0)
#define PROF_DUMP(role) do { if (p.trace && lane == 0) { for (int k_ = 0; k_ < 8; ++k_) p.trace[((size_t)cta * 4 + (role)) * 8 + k_] = pf_t[k_]; } } while (0)
#else
#define PROF_DECL do { } while (0)
#define PROF(k) do { } while (0)
#define PROF_DUMP(role) do { } while (0)
#endif

struct Ring {
    int s;
    uint32_t ph;
    __device__ __forceinline__ void advance(int n) {
        if (++s == n) {
            s = 0;
            ph ^= 1u;
        }
    }
};

struct FusedParams {
    int cin, gh, gw, pw, cout, batch;
    int P;               // gh*gw
    int rows;            // R = 128 + 2*(pw+1) patch rows actually read
    int plane_bytes;     // round8(R) * 128
    int tiles_per_img, total_tiles, pair_tiles;
    int upt;             // units (= 64-channel blocks) per tile
    int units_per_pair;
    unsigned magic_gw, magic_pw;   // x / d == __umulhi(x, magic) for x < 65536
    int out_stride;
    int dbg;             // developer ablations (SQD_F16_DBG): 1 skip MMA issue, 2 skip feature loads, 4 skip convert + stores, 8 skip B loads
    const float *feat;
    const unsigned *amax_bits;   // (B, Cin/64) max|x| per image and channel block, fp32 bits (absmax pre-kernel)
    const float *bias;
    const PackedHeader *whdr;
    float *pred;
    float *partial;      // (grid, 128, NPAD)
    int *flags;          // (grid)
    int *status;
    long long *trace;    // profiling builds only
};

// the unit sequence of one pair: [whole tiles + head segment][deferred tail segment], as in convdet_f16.cu
struct Sched {
    long long u0;
    int n, main_len, upt;
    __device__ __forceinline__ long long unit(int i) const {
        const int len_tail = n - main_len;
        return i < main_len ? u0 + len_tail + i : u0 + (i - main_len);
    }
};

struct BlockIter {       // (pair-tile, channel block) of unit i, and this rank's tile
    int pt, cb, img, t;  // img == batch: ghost tile
    __device__ __forceinline__ void locate(const FusedParams &p, int tile_offset) {
        const int tile = pt + tile_offset;
        if (tile >= p.total_tiles) {
            img = p.batch;
            t = 0;
        } else {
            img = tile / p.tiles_per_img;
            t = tile - img * p.tiles_per_img;
        }
    }
    __device__ __forceinline__ void seek(long long u, const FusedParams &p, int tile_offset) {
        pt = (int)(u / p.upt);
        cb = (int)(u - (long long)pt * p.upt);
        locate(p, tile_offset);
    }
    __device__ __forceinline__ void next(const FusedParams &p, int tile_offset) {
        if (++cb == p.upt) {
            cb = 0;
            ++pt;
            locate(p, tile_offset);
        }
    }
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {   // K-major SWIZZLE_128B, SBO = 1024 (tc_ptx.cuh)
    return umma_desc_sw128(saddr);
}

// ---------------------------------------------------------------------------------------------------------------------
template <int NPAD, int HR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsF, 1)
convdet_fused_pair_kernel(const __grid_constant__ CUtensorMap map_b, const FusedParams p) {
    constexpr int N1H = 2 * HR, N2H = (HR + 7) / 8 * 8;
    static_assert(HR % 4 == 0 && 2 * HR <= NPAD && N2H <= N1H, "pair layout");
    constexpr int kBTapBytes = N1H * kBlockK * 2;
    constexpr int kBStageBytes = 3 * kBTapBytes;          // the three dx taps of one dy row
    constexpr int kAccCols = 2 * N1H + 2 * N2H;           // [MMA1: 4*HR | MMA2: 2*N2H]
    static_assert(2 * kAccCols <= 512, "two TMEM accumulators must fit");
    constexpr uint32_t kTmemCols = 512;
    constexpr uint32_t kIdesc1 = umma_idesc_f16(256, 2 * N1H);
    constexpr uint32_t kIdesc2 = umma_idesc_f16(256, 2 * N2H);
    static_assert((2 * N1H) % 16 == 0 && (2 * N2H) % 16 == 0 && kBTapBytes % 1024 == 0, "UMMA N / swizzle atom");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *a_buf = smem;                                            // [2 buffers][2 planes][plane_bytes]
    uint8_t *b_ring = smem + (size_t)4 * p.plane_bytes;               // [kNB][3 taps][kBTapBytes]
    uint8_t *ctrl = b_ring + (size_t)kNB * kBStageBytes;
    uint64_t *bfull = reinterpret_cast<uint64_t *>(ctrl);   // [kNB] leader: the stage's bytes of both CTAs landed
    uint64_t *bfree = bfull + kNB;                          // [kNB] both: the MMAs reading the stage completed
    uint64_t *afull = bfree + kNB;                          // [2]   leader: both CTAs' converters filled operand buffer b
    uint64_t *afree = afull + 2;                            // [2]   both: the MMAs reading operand buffer b completed
    uint64_t *aconv = afree + 2;                            // [2]   own CTA: all 8 converter warps stored their share of buffer b
    uint64_t *tfull = aconv + 2;                            // [2]   both: chunk complete in TMEM accumulator b
    uint64_t *tempty = tfull + 2;                           // [2]   leader: all 8 accumulate warps of the pair drained it
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    volatile int *abort_flag = reinterpret_cast<volatile int *>(tmem_slot + 1);
    float *s_bias = reinterpret_cast<float *>(tmem_slot + 2);        // [NPAD]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const uint32_t rank = cluster_ctarank();
    const int pair = cta >> 1;
    const int tile_offset = rank ? p.pair_tiles : 0;

    const long long total_units = (long long)p.pair_tiles * p.upt;
    Sched sc;
    sc.upt = p.upt;
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

    if (threadIdx.x == 0) {
        *abort_flag = 0;
        for (int s = 0; s < kNB; ++s) {
            mbar_init(bfull + s, 1);
            mbar_init(bfree + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(afull + b, 2);    // one arrive per CTA of the pair
            mbar_init(afree + b, 1);
            mbar_init(aconv + b, kCvtWarps);
            mbar_init(tfull + b, 1);
            mbar_init(tempty + b, 8);   // 4 accumulate warps x 2 CTAs
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == kWarpB && lane == 0) tma_prefetch_desc(&map_b);
    for (int i = threadIdx.x; i < NPAD; i += kThreadsF) s_bias[i] = (i < p.cout && p.bias) ? __ldg(p.bias + i) : 0.f;
    if (warp == kWarpMma) tmem_alloc_2cta(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    sqd_pdl_wait();   // no-op unless launched as a programmatic dependent

    if (warp >= kWarpB) {
        reg_dec<kRegsLight>();
        if (warp == kWarpB) {
            // ===== B producer: this CTA's [w1 | w2] rows of the three dx taps of one dy row of one channel block per stage =====
            BlockIter it;
            Ring rb{0, 0};
            bool ok = true;
            PROF_DECL;
            for (int i = 0; i < n_units && ok; ++i) {
                if (i == 0 || i == sc.main_len) it.seek(sc.unit(i), p, tile_offset); else it.next(p, tile_offset);
                for (int dy = 0; dy < 3; ++dy) {
                    PROF(1);
                    if (!mbar_wait_warp(bfree + rb.s, rb.ph ^ 1u, abort_flag)) {
                        if (lane == 0) atomicCAS(p.status, 0, 1);
                        ok = false;
                        break;
                    }
                    PROF(0);
                    if (elect_one_sync()) {
                        if (rank == 0) mbar_arrive_expect_tx(bfull + rb.s, (p.dbg & 8) ? 0 : 2 * kBStageBytes);   // both CTAs' bytes
                        if (!(p.dbg & 8)) {
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx)
                                tma_load_2d_2cta(&map_b, bfull + rb.s, b_ring + (size_t)rb.s * kBStageBytes + dx * kBTapBytes,
                                                 (dy * 3 + dx) * p.cin + it.cb * kBlockK, (int)rank * N1H);
                        }
                    }
                    __syncwarp();
                    rb.advance(kNB);
                }
            }
            PROF_DUMP(3);
        } else if (warp == kWarpMma) {
            if (rank == 0) {
                // ===== MMA issuer (leader CTA): 72 M=256 MMAs per block; converged warp, one elected lane issues =====
                Ring rb{0, 0};
                bool ok = true;
                const uint32_t a0 = smem_u32(a_buf), b0 = smem_u32(b_ring);
                const uint64_t plane_adv = (uint64_t)(p.plane_bytes >> 4);      // descriptor start-address units (16 B)
                const uint64_t dy_adv = (uint64_t)(p.pw * 128 >> 4);
                PROF_DECL;
                for (int i = 0; i < n_units && ok; ++i) {
                    const int buf = i & 1;
                    const uint32_t ph = (uint32_t)(i >> 1) & 1u;
                    PROF(3);
                    if (!mbar_wait_warp(tempty + buf, ph ^ 1u, abort_flag)) {
                        if (lane == 0) atomicCAS(p.status, 0, 4);
                        break;
                    }
                    PROF(0);
                    if (!mbar_wait_warp(afull + buf, ph, abort_flag)) {
                        if (lane == 0) atomicCAS(p.status, 0, 2);
                        break;
                    }
                    PROF(1);
                    tc_fence_after();
                    const uint32_t d1 = tmem_base + (uint32_t)buf * kAccCols, d2 = d1 + 2 * N1H;
                    const uint64_t a_blk = desc_sw128(a0 + (uint32_t)buf * 2u * (uint32_t)p.plane_bytes);
#pragma unroll 1
                    for (int dy = 0; dy < 3; ++dy) {
                        PROF(3);
                        if (!mbar_wait_warp(bfull + rb.s, rb.ph, abort_flag)) {
                            if (lane == 0) atomicCAS(p.status, 0, 5);
                            ok = false;
                            break;
                        }
                        PROF(2);
                        tc_fence_after();
                        if (elect_one_sync()) {
                            const uint64_t a_row = a_blk + (uint64_t)dy * dy_adv;
                            const uint64_t b_stage = desc_sw128(b0 + (uint32_t)rb.s * kBStageBytes);
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                const uint64_t a1 = a_row + (uint64_t)(dx * 128 >> 4);       // one pixel = one 128-byte row
                                const uint64_t a2 = a1 + plane_adv;
                                const uint64_t b = b_stage + (uint64_t)(dx * kBTapBytes >> 4);
#pragma unroll
                                for (int ks = 0; ks < kBlockK / kUmmaK; ++ks) {
                                    if (p.dbg & 1) continue;
                                    const uint64_t adv = (uint64_t)((ks * kUmmaK * 2) >> 4);
                                    const uint32_t accum = (dy | dx | ks) ? 1u : 0u;
                                    umma_f16_ss_2cta(d1, a1 + adv, b + adv, kIdesc1, accum);    // a1 x [w1 | w2]
                                    umma_f16_ss_2cta(d2, a2 + adv, b + adv, kIdesc2, accum);    // a2 x w1 (first rows)
                                }
                            }
                            umma_commit_2cta(bfree + rb.s, 3);
                            if (dy == 2) {
                                umma_commit_2cta(afree + buf, 3);   // operand buffer reusable in both CTAs
                                umma_commit_2cta(tfull + buf, 3);   // chunk complete (both CTAs' accumulate warps)
                            }
                        }
                        __syncwarp();
                        rb.advance(kNB);
                    }
                }
                PROF(3);
                PROF_DUMP(0);
            }
            sqd_pdl_trigger();   // all MMAs issued: the kernel behind us may be scheduled while the last chunks drain
        } else if (warp == kWarpRelay) {
            // ===== relay (both CTAs): converters' generic-proxy stores -> async proxy -> the leader's "operand full" =====
            // The converters keep the NEXT block's global loads in flight while they store, so they must not execute the
            // proxy fence themselves (it is a MEMBAR: it would wait for those loads and expose their whole latency).  They
            // release-arrive on a CTA-local barrier instead; this warp acquires it, fences the proxies in the CTA that
            // holds the data, and only then tells the MMA issuer.
            for (int i = 0; i < n_units; ++i) {
                const int buf = i & 1;
                const uint32_t ph = (uint32_t)(i >> 1) & 1u;
                if (!mbar_wait_warp(aconv + buf, ph, abort_flag)) {
                    if (lane == 0) atomicCAS(p.status, 0, 7);
                    break;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (rank == 0) mbar_arrive(afull + buf);
                    else mbar_arrive_cluster(afull + buf, 0);
                }
            }
        }
    } else if (warp >= kWarpAcc) {
        reg_inc<kRegsAcc>();
        // ===== accumulate + epilogue warps (both CTAs, each on its own 128 TMEM lanes) =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - kWarpAcc * 32;  // 0..127
        const float inv_sw = p.whdr->inv_scale;
        float acc[NPAD];
        int seg_c0 = 0;
        BlockIter it;
        PROF_DECL;
        for (int i = 0; i < n_units; ++i) {
            PROF(2);
            if (i == 0 || i == sc.main_len) it.seek(sc.unit(i), p, tile_offset); else it.next(p, tile_offset);
            const int cb = it.cb;
            if (i == 0 || cb == 0 || i == sc.main_len) {
                seg_c0 = cb;
#pragma unroll
                for (int n = 0; n < NPAD; ++n) acc[n] = 0.f;
            }
            const int buf = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            if (!mbar_wait_warp(tfull + buf, ph, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 3);
                break;
            }
            PROF(0);
            tc_fence_after();
            __syncwarp();
            // 1 / (feature scale of this image and channel block): an exact power of two, so the fma below rounds once,
            // like an fp32 add of the unscaled chunk.  Ghost tile: any value.
            const float inv_a = it.img < p.batch
                ? 1.f / pow2_scale_for(__uint_as_float(__ldg(p.amax_bits + (size_t)it.img * p.upt + cb)))
                : 1.f;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * kAccCols;
            // groups of 8 output channels: g = h*NG + j, columns main h*N1H + 8j, cross1 + HR, cross2 2*N1H + h*N2H + 8j;
            // the loads of group g+1 are in flight while group g is folded
            constexpr int NG = (HR + 7) / 8;
            uint32_t ld[2][24];
            auto issue = [&](int g, uint32_t *dst) {
                const int h = g / NG, j0 = (g - h * NG) * 8;
                tmem_ld_x8(taddr + h * N1H + j0, dst);
                tmem_ld_x8(taddr + h * N1H + HR + j0, dst + 8);
                tmem_ld_x8(taddr + 2 * N1H + h * N2H + j0, dst + 16);
            };
            issue(0, ld[0]);
#pragma unroll
            for (int g = 0; g < 2 * NG; ++g) {
                tmem_ld_wait();
                if (g + 1 < 2 * NG) issue(g + 1, ld[(g + 1) & 1]);
                const int h = g / NG, j0 = (g - h * NG) * 8;
                const uint32_t *v = ld[g & 1];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (j0 + k < HR) {
                        const float cross = fadd(__uint_as_float(v[8 + k]), __uint_as_float(v[16 + k]));
                        acc[h * HR + j0 + k] = fmaf(fmaf(cross, kLoInv, __uint_as_float(v[k])), inv_a, acc[h * HR + j0 + k]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(tempty + buf);
                else mbar_arrive_cluster(tempty + buf, 0);
            }
            PROF(1);

            const bool seg_end = (i == n_units - 1) || (cb == p.upt - 1) || (i == sc.main_len - 1);
            if (!seg_end) continue;
            const bool from_start = seg_c0 == 0, to_end = cb == p.upt - 1;
            if (from_start && !to_end) {
                // head of a split pair-tile: publish for the same rank of the next pair (quad-major: 512 B per warp store)
                float4 *dst = reinterpret_cast<float4 *>(p.partial + (size_t)cta * 128 * NPAD) + row;
#pragma unroll
                for (int n = 0; n < NPAD; n += 4) dst[(n >> 2) * 128] = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
                __threadfence();
                acc_bar();
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
                acc_bar();
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
            // whole tile in registers: x 1/s_w, + bias -> pred   (ghost tile: nothing stored)
            const int f = it.t * kTileM + row;
            const int y = (int)__umulhi((unsigned)f, p.magic_pw), x = f - y * p.pw;
            const bool inb = it.img < p.batch && y < p.gh && x < p.gw;
#pragma unroll
            for (int n = 0; n < NPAD; ++n) acc[n] = fadd(fmul(acc[n], inv_sw), s_bias[n]);
            if (inb) {
                float *out = p.pred + (((size_t)it.img * p.gh + y) * p.gw + x) * p.out_stride;
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
        }
        PROF(2);
        if (warp == kWarpAcc) PROF_DUMP(2);
    } else {
        // ===== converter warps: fp32 NCHW patch -> scale -> two-term fp16 split -> swizzled K-major A operand =====
        // A thread's share of a patch (72 floats in registers), over the 288 source pixels [s_al, s_al + 288) of the 64
        // channels, s_al = the patch's first source pixel rounded down to a multiple of 4:
        //   rounds 0, 1  ("quads"): pixels 4q .. 4q+3 of the 8 channels of octet o -- eight 16-byte loads.  A warp covers
        //                8 quads x 4 octets; lane = o' + 4*(q & 1) + 8*(q >> 1), so that a load instruction reads 4 channels
        //                x 128 contiguous bytes and a quarter warp's 16-byte stores (2 consecutive... rows 4 apart x 4 octets)
        //                hit all 32 banks once;
        //   round 2      ("tail"): pixel 256 + lane of the 8 channels of octet = warp -- eight 4-byte loads.
        // The values of block i+1 are re-loaded round by round right after block i's round has been converted and stored,
        // so the loads are in flight during the rest of the conversion, the publish step and the wait for the next buffer.
        const int cw = warp - kWarpCvt;             // 0..7
        const int ct = threadIdx.x - kWarpCvt * 32; // 0..255
        float v[72];
        BlockIter it, nx;
        struct Geo {
            int s_lo, s_hi, s_al, f0;
            const float *src;
        } cur, nxt;
        auto geometry = [&](const BlockIter &b, Geo &g) {
            g.f0 = b.t * kTileM - p.pw - 1;
            if (b.img >= p.batch) {   // ghost tile: nothing to load, every row is zero
                g.s_lo = g.s_hi = g.s_al = 0;
                g.src = p.feat;
                return;
            }
            const int f_lo = g.f0 < 0 ? 0 : g.f0;
            int f_hi = g.f0 + p.rows;
            if (f_hi > p.gh * p.pw) f_hi = p.gh * p.pw;
            f_hi -= 1;                                               // last padded-flat position of the patch
            const int y_lo = (int)__umulhi((unsigned)f_lo, p.magic_pw), x_lo = f_lo - y_lo * p.pw;
            const int y_hi = (int)__umulhi((unsigned)f_hi, p.magic_pw), x_hi = f_hi - y_hi * p.pw;
            g.s_lo = x_lo == p.gw ? (y_lo + 1) * p.gw : y_lo * p.gw + x_lo;
            g.s_hi = x_hi == p.gw ? (y_hi + 1) * p.gw : y_hi * p.gw + x_hi + 1;
            g.s_al = g.s_lo & ~3;
            g.src = p.feat + ((size_t)b.img * p.cin + (size_t)b.cb * kBlockK) * p.P;
        };
        // quad rounds: this thread's (quad, octet) of round j
        const int lq = ((lane >> 2) & 1) + 2 * (lane >> 3), lo = lane & 3;
        auto quad_of = [&](int j, int &q, int &o) {
            const int T = j * kCvtWarps + cw;         // warp tile 0..15: quad group T & 7, octet group T >> 3
            q = (T & 7) * 8 + lq;
            o = (T >> 3) * 4 + lo;
        };
        auto load_quads = [&](int j, const Geo &g) {
            int q, o;
            quad_of(j, q, o);
            const int s = g.s_al + 4 * q;
            const bool okq = s < g.s_hi && !(p.dbg & 2);      // the quad overlaps [s_lo, s_hi) (s + 3 >= s_lo always holds)
            const float4 *ptr = reinterpret_cast<const float4 *>(g.src + (size_t)(o * 8) * p.P + s);
            const size_t cstride = (size_t)(p.P >> 2);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 t = okq ? ld_stream_f4(ptr + c * cstride) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[j * 32 + c * 4 + 0] = t.x; v[j * 32 + c * 4 + 1] = t.y; v[j * 32 + c * 4 + 2] = t.z; v[j * 32 + c * 4 + 3] = t.w;
            }
        };
        auto load_tail = [&](const Geo &g) {
            const int s = g.s_al + kMainPix + lane;
            const bool okp = s < g.s_hi && !(p.dbg & 2);
            const float *ptr = g.src + (size_t)(cw * 8) * p.P + s;
#pragma unroll
            for (int c = 0; c < 8; ++c) v[64 + c] = okp ? ldg_stream(ptr + (size_t)c * p.P) : 0.f;
        };
        // one pixel (8 channels of octet o, values x[0..7]) -> row of both planes (shared-window addresses pl1, pl1 + plane)
        const unsigned magic_gw = p.magic_gw, magic_pw = p.magic_pw;
        const uint32_t plane_b = (uint32_t)p.plane_bytes;
        const bool no_cvt = (p.dbg & 4) != 0;
        auto put_pixel = [&](const Geo &g, uint32_t pl1, int s, int o, float sa, const float *x) {
            if (s < g.s_lo || s >= g.s_hi || no_cvt) return;
            const int y = (int)__umulhi((unsigned)s, magic_gw);
            const int r = s + y - g.f0;                 // padded-flat position (pw = gw + 1) relative to the patch
            uint32_t h1[4], h2[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float x0 = x[2 * c] * sa, x1 = x[2 * c + 1] * sa;
                const __half2 hi = __floats2half2_rn(x0, x1);
                const float2 fhi = __half22float2(hi);
                const __half2 lo2 = __floats2half2_rn((x0 - fhi.x) * kLoScale, (x1 - fhi.y) * kLoScale);
                h1[c] = *reinterpret_cast<const uint32_t *>(&hi);
                h2[c] = *reinterpret_cast<const uint32_t *>(&lo2);
            }
            const uint32_t off = pl1 + (uint32_t)r * 128u + (uint32_t)((o ^ (r & 7)) << 4);
            sts128(off, h1[0], h1[1], h1[2], h1[3]);
            sts128(off + plane_b, h2[0], h2[1], h2[2], h2[3]);
        };
        const uint32_t a_base = smem_u32(a_buf);
        PROF_DECL;
        int zt0 = -1, zt1 = -1;   // pair-tile whose pad rows operand buffer 0 / 1 currently holds as zeros
        if (n_units > 0) {
            it.seek(sc.unit(0), p, tile_offset);
            geometry(it, cur);
            load_quads(0, cur);
            load_quads(1, cur);
            load_tail(cur);
        }
        for (int i = 0; i < n_units; ++i) {
            const int buf = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            const bool more = i + 1 < n_units;
            if (more) {
                nx = it;
                if (i + 1 == sc.main_len) nx.seek(sc.unit(i + 1), p, tile_offset); else nx.next(p, tile_offset);
                geometry(nx, nxt);
            }
            const float sa = it.img < p.batch
                ? pow2_scale_for(__uint_as_float(__ldg(p.amax_bits + (size_t)it.img * p.upt + it.cb)))
                : 1.f;
            // 1. operand buffer free?
            PROF(5);
            if (!mbar_wait_warp(afree + buf, ph ^ 1u, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 6);
                break;
            }
            PROF(0);
            const uint32_t pl1 = a_base + (uint32_t)buf * 2u * plane_b;
            // 2. pad rows (between image rows, above / below the image): zero, once per (tile, buffer)
            if ((buf ? zt1 : zt0) != it.pt) {
                if (buf) zt1 = it.pt; else zt0 = it.pt;
                const int lim = p.gh * p.pw;
                for (int r = ct; r < p.rows; r += kCvtThreads) {
                    const int f = cur.f0 + r;
                    bool pad = it.img >= p.batch || f < 0 || f >= lim;
                    if (!pad) {
                        const int y = (int)__umulhi((unsigned)f, p.magic_pw);
                        pad = f - y * p.pw == p.gw;
                    }
                    if (pad) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            sts128(pl1 + (uint32_t)r * 128u + c * 16, 0u, 0u, 0u, 0u);
                            sts128(pl1 + plane_b + (uint32_t)r * 128u + c * 16, 0u, 0u, 0u, 0u);
                        }
                    }
                }
            }
            PROF(2);
            // 3. split and store
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int q, o;
                quad_of(j, q, o);
                const int s = cur.s_al + 4 * q;
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                    float x[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) x[c] = v[j * 32 + c * 4 + px];
                    put_pixel(cur, pl1, s + px, o, sa, x);
                }
            }
            {
                float x[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) x[c] = v[64 + c];
                put_pixel(cur, pl1, cur.s_al + kMainPix + lane, cw, sa, x);
            }
            PROF(3);
            // 4. publish to the relay warp: the warp's stores, then one release-arrive (no fence here, see the relay)
            __syncwarp();
            if (lane == 0) mbar_arrive(aconv + buf);
            // 5. the next block's patch into the registers just freed: in flight during the wait for the next buffer.
            //    (ptxas tracks every load of this loop on ONE scoreboard slot, so refilling round by round would make
            //    each round's conversion wait for the loads issued just before it.)
            if (more) {
                load_quads(0, nxt);
                load_quads(1, nxt);
                load_tail(nxt);
            }
            PROF(4);
            if (more) {
                it = nx;
                cur = nxt;
            }
        }
        PROF(5);
        if (cw == 0) PROF_DUMP(1);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still touch its shared memory / barriers
    tc_fence_after();
    if (warp == kWarpMma) {
        __syncwarp();
        tmem_dealloc_2cta(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
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
int pair_hr_of(int cout) {   // must match convdet_f16.cu (the packed weight layout)
    const int npad = npad_of(cout);
    if (npad == 80 && cout <= 72) return 36;
    return npad / 2;
}
unsigned magic_for(int d) { return (unsigned)(0xFFFFFFFFu / (unsigned)d) + 1u; }   // x / d for x < 65536, d < 65536

template <int NPAD, int HR>
int launch_fused(const CUtensorMap &map_b, const FusedParams &p, int grid, size_t smem, cudaStream_t st) {
    SQD_CUDA(cudaFuncSetAttribute(convdet_fused_pair_kernel<NPAD, HR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaError_t e = sqd_launch_dependent(convdet_fused_pair_kernel<NPAD, HR>, dim3(grid), dim3(kThreadsF), smem, st, true, map_b, p);
    if (e != cudaSuccess) {
        sqd_set_error("launch of convdet_fused_pair_kernel failed: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return SQD_OK;
}

size_t fused_smem_bytes(int gw, int n1h) {
    const int pw = gw + 1, rows = kTileM + 2 * (pw + 1), rpad = (rows + 7) & ~7;
    return 1024 + (size_t)4 * rpad * 128 + (size_t)kNB * 3 * n1h * kBlockK * 2 + kCtrlBytes;
}

}  // namespace

// implemented in convdet_f16.cu: max |x| of contiguous runs (one per blockIdx.y), atomicMax into d_amax (zeroed by the caller)
int sqd_f16_absmax_runs(const float *d_in, size_t run_floats, int nruns, unsigned *d_amax, cudaStream_t st);

// Shapes the one-kernel path takes (everything else goes through the pre-pass + convdet_f16_pair_kernel).
bool sqd_convdet_fused_eligible(int layout, int batch, int cin, int gh, int gw, int cout) {
    if (layout != SQD_LAYOUT_NCHW || batch < 1 || (long long)batch * (cin / kBlockK) > 65535 || cin < kBlockK || cin % kBlockK != 0 || cout < 1 || cout > 80) return false;
    const int pw = gw + 1, rows = kTileM + 2 * (pw + 1);
    if (gw < 1 || gh < 1 || rows > kMaxRows || (gh * gw) % 4 != 0 || (long long)gh * pw + 256 >= 65536) return false;
    return fused_smem_bytes(gw, 2 * pair_hr_of(cout)) <= kSmemLimit;
}

// workspace: the [status (256 B)][flags: 256 ints][partials: #SM*128*npad floats] prefix of convdet_f16.cu's layout
int sqd_convdet_fused(const float *d_feat, const void *d_packed, const float *d_bias, int batch, int cin, int gh, int gw,
                      int cout, float *d_pred, void *d_workspace, cudaStream_t st) {
    SQD_REQUIRE(sqd_convdet_fused_eligible(SQD_LAYOUT_NCHW, batch, cin, gh, gw, cout), SQD_E_SHAPE,
                "convdet (fused tcgen05): shape outside the kernel's limits");
    EncodeTiledFn encode = get_encode_fn();
    SQD_REQUIRE(encode != nullptr, SQD_E_DRIVER, "cuTensorMapEncodeTiled not available from the CUDA driver");
    const int npad = npad_of(cout), hr = pair_hr_of(cout), n1h = 2 * hr;
    char *ws = static_cast<char *>(d_workspace);
    const size_t flags_off = 256, partial_off = flags_off + 256 * sizeof(int);
    const size_t amax_off = (partial_off + (size_t)SQD_SM_COUNT * 128 * npad * sizeof(float) + 255) & ~(size_t)255;
    const int ncb = cin / kBlockK;
    SQD_CUDA(cudaMemsetAsync(ws, 0, partial_off, st));   // status + flags
    SQD_CUDA(cudaMemsetAsync(ws + amax_off, 0, (size_t)batch * ncb * sizeof(unsigned), st));
    // read-only max pass: max|x| per (image, 64-channel block) -> the power-of-two feature scales (an (image, block) slab
    // of an NCHW tensor is one contiguous run of 64*P floats).  It also leaves the features in L2 for the GEMM.
    {
        int rc = sqd_f16_absmax_runs(d_feat, (size_t)kBlockK * gh * gw, batch * ncb, reinterpret_cast<unsigned *>(ws + amax_off), st);
        if (rc) return rc;
    }

    alignas(64) CUtensorMap map_b;
    {
        const size_t ktot = (size_t)9 * cin;
        void *mat2 = const_cast<char *>(static_cast<const char *>(d_packed) + kHeaderBytes) + (size_t)2 * npad * ktot * sizeof(__half);
        const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)(2 * n1h)};
        const cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
        const cuuint32_t box[2] = {kBlockK, (cuuint32_t)n1h};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, mat2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SQD_REQUIRE(r == CUDA_SUCCESS, SQD_E_DRIVER, "cuTensorMapEncodeTiled(weights) failed: CUresult %d", (int)r);
    }

    FusedParams p;
    p.cin = cin; p.gh = gh; p.gw = gw; p.pw = gw + 1; p.cout = cout; p.batch = batch;
    p.P = gh * gw;
    p.rows = kTileM + 2 * (p.pw + 1);
    p.plane_bytes = ((p.rows + 7) & ~7) * 128;
    p.tiles_per_img = (gh * p.pw + kTileM - 1) / kTileM;
    const long long total_tiles = (long long)p.tiles_per_img * batch;
    SQD_REQUIRE(total_tiles < (1ll << 30), SQD_E_SHAPE, "convdet (fused tcgen05): too many tiles");
    p.total_tiles = (int)total_tiles;
    p.pair_tiles = (int)((total_tiles + 1) / 2);
    p.upt = cin / kBlockK;
    const int max_pairs = SQD_SM_COUNT / 2;
    const int npairs = p.pair_tiles < max_pairs ? p.pair_tiles : max_pairs;
    const long long total_units = (long long)p.pair_tiles * p.upt;
    long long upp = (total_units + npairs - 1) / npairs;
    if (upp < p.upt) upp = p.upt;
    p.units_per_pair = (int)upp;
    p.magic_gw = magic_for(gw);
    p.magic_pw = magic_for(p.pw);
    p.out_stride = cout;
    p.dbg = sqd_opt(SQD_OPT_F16_DBG);
    p.feat = d_feat;
    p.amax_bits = reinterpret_cast<const unsigned *>(ws + amax_off);
    p.bias = d_bias;
    p.whdr = static_cast<const PackedHeader *>(d_packed);
    p.pred = d_pred;
    p.partial = reinterpret_cast<float *>(ws + partial_off);
    p.flags = reinterpret_cast<int *>(ws + flags_off);
    p.status = reinterpret_cast<int *>(ws);
    p.trace = nullptr;
#ifdef SQD_ENABLE_TRACE
    if (const char *e = getenv("SQD_F16_TRACE")) p.trace = reinterpret_cast<long long *>(strtoull(e, nullptr, 0));
#endif
    const int grid = 2 * npairs;
    const size_t smem = fused_smem_bytes(gw, n1h);
    switch (npad / 16) {
        case 1: return launch_fused<16, 8>(map_b, p, grid, smem, st);
        case 2: return launch_fused<32, 16>(map_b, p, grid, smem, st);
        case 3: return launch_fused<48, 24>(map_b, p, grid, smem, st);
        case 4: return launch_fused<64, 32>(map_b, p, grid, smem, st);
        case 5: return hr == 36 ? launch_fused<80, 36>(map_b, p, grid, smem, st) : launch_fused<80, 40>(map_b, p, grid, smem, st);
    }
    SQD_REQUIRE(false, SQD_E_SHAPE, "convdet (fused tcgen05): unsupported Cout %d", cout);
}

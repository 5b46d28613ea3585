// SURVEY 8(f) rank 2, second half: ConvDet WEIGHT gradient on tcgen05 / TMEM (sm_100a), f16x3 numerics.
// Reference: autograd through nn.Conv2d(768, 72, 3, padding=1), src/model/squeezedet.py:73-75 (cuDNN wgrad there).
//
//   dW[n,c,dy,dx] = sum_{b,y,x} G[b,y,x,n] * X[b,c,y+dy-1,x+dx-1]
// Per tap (dy,dx) this is a GEMM  D[c, n] = sum_pixels Xshift[c, pix] * G[pix, n]  whose contraction runs over PIXELS, so
// both operands must be pixel-contiguous ("K-major" in UMMA terms), and TMA dictates how (measured, tools/micro/tma_box.cu:
// with SWIZZLE_128B the innermost box extent must be the full 128 B, and the innermost start coordinate must be 16-byte
// aligned -- a one-pixel tap shift cannot be a box origin):
//   * every image plane is stored FLAT with rows padded by at least one zero column to gwp (a multiple of 8): pixel
//     (y,x) -> j = y*gwp + x.  A K tile is 64 consecutive flat pixels.  Thanks to the zero pad columns a flat shift by
//     (dy-1)*gwp + (dx-1) equals the 2-D shift with zero padding; shifts past either end are TMA out-of-bounds zeros.
//   * A (M = 128 channels x K = 64 pixels): NCHW fp16 planes of the features, 3-D box {64 px, 128 ch, 1 img} at the
//     aligned origin 64*t: 128 rows of 128 B = one SWIZZLE_128B K-major tile.  Never shifted.
//   * B (N = 80 outputs x K = 64 pixels): G transposed to (B, 80, P_pad) planes; the tap moves G instead of X:
//     G^T[n, q - off].  The row part of the shift, (dy-1)*gwp, is a multiple of 8 and goes into the box origin; the
//     column part (dx-1) needs THREE pre-shifted copies of the (small) G^T planes.  [g2 | g1] are two boxes stacked.
// Numerics as in the forward (convdet_f16.cu): x*s = x1 + x2/2^11, g*t = g1 + g2/2^11 with power-of-two scales per
// (image, 64-channel block) and per (image, 64-output block); per K step of 16 pixels two MMAs
//     D[:, 0:160] (+)= X1 * [g2 | g1]^T        D[:, 0:80] += X2 * g1^T
// A TMEM accumulation chunk = the pixel tiles of ONE image (<= 30 tiles x 4 K steps): the scales are uniform inside it,
// and the accumulate warps fold it into fp32 registers with the row's 1/s (channel block) and the column's 1/t.
// Work item = (128-channel block, tap, slice of the batch's pixel tiles); slices write fp32 partials that a second
// kernel adds in a fixed order (deterministic).  Both operands stream (no operand is reused across K), so the kernel
// is L2 -> SM bound, not tensor bound: 52 KB per 64-pixel step per CTA.
// Warp roles (192 threads): 0 TMA producer, 1 TMEM alloc + MMA issue, 2..5 accumulate.  Every wait is bounded.
#include <cuda.h>
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace sqd_tc;

constexpr int kMTile = 128;            // channels per work item (UMMA_M)
constexpr int kNPad = 80;              // outputs, padded (Cout <= 80)
constexpr int kPixTile = 64;            // K tile = 64 flat pixels = one 128-byte swizzle row of fp16
constexpr int kABytes = kMTile * 128;  // one plane of one stage (16 KB)
constexpr int kBBytes = kNPad * 128;   // one plane of G^T (10 KB)
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;   // x1 | x2 | g2 | g1 = 52 KB
constexpr int kStages = 4;
constexpr int kThreads = 192;
constexpr int kAccCols = 2 * kNPad;    // [cross | main]
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;

__device__ __forceinline__ float pow2_scale_for(float amax) {   // must match convdet_f16.cu
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
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- operand preparation ---------------------------------------------------------------------------------------------
// X (B,Cin,gh,gw) fp32 -> x1 / x2 fp16 planes (B,Cin,gh,gwp), same NCHW order, rows zero-padded to gwp; scale per
// (image, 64-channel block) from amax (computed by absmax over the contiguous slab).
__global__ void __launch_bounds__(256) split_nchw_rows_kernel(const float *__restrict__ x, int cin, int gh, int gw, int gwp,
                                                              const unsigned *__restrict__ amax_bits, __half2 *__restrict__ p1,
                                                              __half2 *__restrict__ p2) {
    // one block per (image, channel) plane; a thread per pair of columns (gw, gwp even: 8-byte loads, 4-byte stores)
    const size_t plane = blockIdx.x;                     // b*cin + c
    const int b = (int)(plane / cin), c = (int)(plane - (size_t)b * cin);
    const float s = pow2_scale_for(__uint_as_float(__ldg(amax_bits + (size_t)b * (cin >> 6) + (c >> 6))));
    const int hp = gwp >> 1, hw = gw >> 1;
    const float2 *src = reinterpret_cast<const float2 *>(x + plane * gh * gw);
    __half2 *d1 = p1 + plane * gh * hp, *d2 = p2 + plane * gh * hp;
    for (int i = threadIdx.x; i < gh * hp; i += blockDim.x) {
        const int y = i / hp, cp = i - y * hp;
        float2 v = make_float2(0.f, 0.f);
        if (cp < hw) v = __ldg(src + y * hw + cp);
        const float x0 = v.x * s, x1 = v.y * s;
        const __half2 h = __floats2half2_rn(x0, x1);
        const float2 f = __half22float2(h);
        d1[i] = h;
        d2[i] = __floats2half2_rn((x0 - f.x) * kLoScale, (x1 - f.y) * kLoScale);
    }
}

// The same planes AND the block maxima in ONE pass over X: a cluster of 8 CTAs owns one (image, 64-channel block) slab,
// CTA r its channels [8r, 8r+8) = one contiguous run of 8*gh*gw floats held in REGISTERS, the block maximum goes round
// the cluster through distributed shared memory, then every thread scales, splits and stores what it holds.  A thread
// item is one 16-byte chunk of an output row (8 columns: four 8-byte loads, the last chunk of a row ends in the zero
// padding) -> one 16-byte store per plane.  4 B read + 4 B written per element instead of 8 + 4.  Up to 4 items per
// thread (KITTI: 8 channels x 24 rows x 10 chunks = 1920 items per CTA); larger grids take the two kernels above.
constexpr int kXsThreads = 512, kXsCluster = 8, kXsItems = 4, kXsChan = 64 / kXsCluster;
__global__ void __cluster_dims__(kXsCluster, 1, 1) __launch_bounds__(kXsThreads)
    split_nchw_rows_cluster_kernel(const float *__restrict__ x, int cin, int gh, int gw, int gwp, unsigned *__restrict__ amax_bits,
                                   uint4 *__restrict__ p1, uint4 *__restrict__ p2) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float s_red[kXsThreads / 32];
    __shared__ unsigned s_cmax[kXsCluster];
    const int rank = (int)cluster.block_rank(), slab = blockIdx.x / kXsCluster, ncb = cin >> 6;
    const int b = slab / ncb, cb = slab - b * ncb;
    const size_t plane0 = (size_t)b * cin + cb * 64 + rank * kXsChan;
    const int cpr = gwp >> 3, rows = kXsChan * gh, total = rows * cpr;       // chunks per row, rows and chunks of this CTA
    const float2 *src = reinterpret_cast<const float2 *>(x + plane0 * gh * gw);
    const int hw = gw >> 1;
    float2 v[kXsItems][4];
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < kXsItems; ++i) {
        const int q = threadIdx.x + i * kXsThreads;
        const int r = q / cpr, k = q - r * cpr;                              // r = channel * gh + y: rows are contiguous in X
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int pc = 4 * k + e;                                        // column pair inside the row
            v[i][e] = (q < total && pc < hw) ? __ldg(src + (size_t)r * hw + pc) : make_float2(0.f, 0.f);
            m = fmaxf(m, fmaxf(fabsf(v[i][e].x), fabsf(v[i][e].y)));
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < kXsThreads / 32 ? s_red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((int)threadIdx.x < kXsCluster) cluster.map_shared_rank(s_cmax, threadIdx.x)[rank] = __float_as_uint(m);
    }
    cluster.sync();
    unsigned mb = 0u;
#pragma unroll
    for (int r = 0; r < kXsCluster; ++r) mb = max(mb, s_cmax[r]);
    if (rank == 0 && threadIdx.x == 0) amax_bits[slab] = mb;
    const float s = pow2_scale_for(__uint_as_float(mb));
    uint4 *d1 = p1 + plane0 * gh * cpr, *d2 = p2 + plane0 * gh * cpr;
#pragma unroll
    for (int i = 0; i < kXsItems; ++i) {
        const int q = threadIdx.x + i * kXsThreads;
        if (q < total) {
            unsigned h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float x0 = v[i][e].x * s, x1 = v[i][e].y * s;
                const __half2 hh = __floats2half2_rn(x0, x1);
                const float2 f = __half22float2(hh);
                const __half2 ll = __floats2half2_rn((x0 - f.x) * kLoScale, (x1 - f.y) * kLoScale);
                h[e] = *reinterpret_cast<const unsigned *>(&hh);
                l[e] = *reinterpret_cast<const unsigned *>(&ll);
            }
            d1[q] = make_uint4(h[0], h[1], h[2], h[3]);                      // chunk q of the CTA's rows == chunk q of its planes
            d2[q] = make_uint4(l[0], l[1], l[2], l[3]);
        }
    }
}

// G (B, gh*gw, cout) fp32 -> three column-shifted copies e = 0,1,2 (shift e-1) of the transposed fp16 planes
// (3, B, 80, gh*gwp):  copy_e[b][n][y*gwp + col] = G[b][y][col - (e-1)][n]  (zero outside the row, for n >= cout).
// One block per image row: the row's gw x cout floats are read coalesced into shared memory and written back
// transposed, pixel-contiguous.  (A flat shift would pull column 0 of the next row into the last pad column of the
// -1 copy; that position only ever meets a zero pad column of X, so writing zero there is equivalent.)
__global__ void __launch_bounds__(256) g_transpose_split_kernel(const float *__restrict__ g, int gh, int gw, int gwp, int cout,
                                                                int batch, const unsigned *__restrict__ amax_bits,
                                                                __half *__restrict__ p1, __half *__restrict__ p2) {
    extern __shared__ float srow[];   // [gw][cout + 1]
    const int b = blockIdx.y, y = blockIdx.x;
    const int ld = cout + 1;          // odd for even cout: the strided reads below are bank-conflict free
    const float *src = g + ((size_t)b * gh + y) * gw * cout;
    for (int i = threadIdx.x; i < gw * cout; i += blockDim.x) {
        const int col = i / cout, n = i - col * cout;
        srow[col * ld + n] = __ldg(src + i);
    }
    __syncthreads();
    const float t0 = pow2_scale_for(__uint_as_float(__ldg(amax_bits + (size_t)b * 2)));
    const float t1 = pow2_scale_for(__uint_as_float(__ldg(amax_bits + (size_t)b * 2 + 1)));
    const size_t ppad = (size_t)gh * gwp;
    // one item = 8 consecutive columns of one output channel of one shifted copy: one 16-byte store per plane
    const int cpr = gwp >> 3, per_copy = kNPad * cpr, total = 3 * per_copy;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int e = i / per_copy, r = i - e * per_copy, n = r / cpr, k = r - n * cpr;
        const float t = n < 64 ? t0 : t1;
        unsigned h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float xs[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int col = 8 * k + 2 * q + u - (e - 1);
                xs[u] = (n < cout && col >= 0 && col < gw) ? srow[col * ld + n] * t : 0.f;
            }
            const __half2 hh = __floats2half2_rn(xs[0], xs[1]);
            const float2 f = __half22float2(hh);
            const __half2 ll = __floats2half2_rn((xs[0] - f.x) * kLoScale, (xs[1] - f.y) * kLoScale);
            h[q] = *reinterpret_cast<const unsigned *>(&hh);
            l[q] = *reinterpret_cast<const unsigned *>(&ll);
        }
        const size_t o = (((size_t)e * batch + b) * kNPad + n) * ppad + (size_t)y * gwp + 8 * k;   // multiple of 8 halfs
        *reinterpret_cast<uint4 *>(p1 + o) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(p2 + o) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// ---- main kernel -----------------------------------------------------------------------------------------------------
struct WgParams {
    int cin, gh, gw, cout, batch;
    int gwp, tiles_per_img;          // padded row length; 64-pixel tiles per image = ceil(gh*gwp / 64)
    int total_tiles;                 // batch * tiles_per_img
    int nslice;
    const unsigned *amax_x;          // (B, Cin/64)
    const unsigned *amax_g;          // (B, 2)
    float *partial;                  // (nslice, 9, Cin, 80)
    int *status;
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                const __grid_constant__ CUtensorMap map_g1a, const __grid_constant__ CUtensorMap map_g1b,
                const __grid_constant__ CUtensorMap map_g1c, const __grid_constant__ CUtensorMap map_g2a,
                const __grid_constant__ CUtensorMap map_g2b, const __grid_constant__ CUtensorMap map_g2c, const WgParams p) {
    constexpr uint32_t kIdescCat = umma_idesc_f16(128, 2 * kNPad);
    constexpr uint32_t kIdescOne = umma_idesc_f16(128, kNPad);
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *ctrl = smem + (size_t)kStages * kStageBytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(ctrl);      // [4]
    uint64_t *empty = full + 4;                               // [4]
    uint64_t *tmem_full = empty + 4;                          // [2]
    uint64_t *tmem_empty = tmem_full + 2;                     // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    volatile int *abort_flag = reinterpret_cast<volatile int *>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x, tap = blockIdx.y, slice = blockIdx.z;
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int k0 = (int)((long long)p.total_tiles * slice / p.nslice), k1 = (int)((long long)p.total_tiles * (slice + 1) / p.nslice);
    const int ntiles = k1 - k0;

    if (threadIdx.x == 0) {
        *abort_flag = 0;
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tmem_full + b, 1);
            mbar_init(tmem_empty + b, 4);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    // the column part of the tap shift selects one of the three pre-shifted copies of G^T
    const CUtensorMap *map_g1 = dx < 0 ? &map_g1a : (dx == 0 ? &map_g1b : &map_g1c);
    const CUtensorMap *map_g2 = dx < 0 ? &map_g2a : (dx == 0 ? &map_g2b : &map_g2c);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x1);
        tma_prefetch_desc(&map_x2);
        tma_prefetch_desc(map_g1);
        tma_prefetch_desc(map_g2);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: x1 | x2 (shifted by the tap) | g2 | g1 of one 64-pixel tile per stage =====
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < ntiles; ++i) {
            const int kt = k0 + i;
            const int img = kt / p.tiles_per_img, t = kt - img * p.tiles_per_img;
            if (!mbar_wait_warp(empty + s, ph ^ 1u, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 1);
                break;
            }
            if (elect_one_sync()) {
                uint8_t *st = smem + (size_t)s * kStageBytes;
                mbar_arrive_expect_tx(full + s, kStageBytes);
                const int q0 = t * kPixTile;                 // flat pixel origin of the X tile (16-byte aligned)
                const int g0 = q0 - dy * p.gwp;              // G^T origin: the row part of the tap shift (multiple of 8)
                tma_load_3d(&map_x1, full + s, st, q0, mt * kMTile, img);
                tma_load_3d(&map_x2, full + s, st + kABytes, q0, mt * kMTile, img);
                tma_load_3d(map_g2, full + s, st + 2 * kABytes, g0, 0, img);
                tma_load_3d(map_g1, full + s, st + 2 * kABytes + kBBytes, g0, 0, img);
            }
            __syncwarp();
            if (++s == kStages) {
                s = 0;
                ph ^= 1u;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: 8 MMAs per tile; a chunk (= the tiles of one image) accumulates in one TMEM buffer =====
        int s = 0, chunk = 0;
        uint32_t ph = 0;
        bool fresh = true;
        for (int i = 0; i < ntiles; ++i) {
            const int kt = k0 + i;
            const int img = kt / p.tiles_per_img;
            const bool chunk_end = (i == ntiles - 1) || ((kt + 1) / p.tiles_per_img != img);
            const int buf = chunk & 1;
            if (fresh) {
                if (!mbar_wait_warp(tmem_empty + buf, (((uint32_t)(chunk >> 1)) & 1u) ^ 1u, abort_flag)) {
                    if (lane == 0) atomicCAS(p.status, 0, 4);
                    break;
                }
            }
            if (!mbar_wait_warp(full + s, ph, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 2);
                break;
            }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)buf * kAccCols;
            const uint32_t st = smem_u32(smem + (size_t)s * kStageBytes);
            const uint64_t a1 = umma_desc_sw128(st), a2 = umma_desc_sw128(st + kABytes);
            const uint64_t b_cat = umma_desc_sw128(st + 2 * kABytes), b_g1 = umma_desc_sw128(st + 2 * kABytes + kBBytes);
            if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t adv = (uint64_t)((ks * 32) >> 4);   // +32 B per K step of 16 pixels
                    umma_f16_ss(d_tmem, a1 + adv, b_cat + adv, kIdescCat, (fresh && ks == 0) ? 0u : 1u);
                    umma_f16_ss(d_tmem, a2 + adv, b_g1 + adv, kIdescOne, 1u);
                }
                umma_commit(empty + s);
                if (chunk_end) umma_commit(tmem_full + buf);
            }
            __syncwarp();
            fresh = chunk_end;
            if (chunk_end) ++chunk;
            if (++s == kStages) {
                s = 0;
                ph ^= 1u;
            }
        }
    } else {
        // ===== accumulate warps: TMEM chunk -> fp32 registers with the image's scales divided out =====
        const int q = warp & 3;
        const int row = q * 32 + lane;                 // channel inside the 128-channel block
        const int c = mt * kMTile + row;
        float acc[kNPad];
#pragma unroll
        for (int n = 0; n < kNPad; ++n) acc[n] = 0.f;
        int chunk = 0;
        for (int i = 0; i < ntiles; ++i) {
            const int kt = k0 + i;
            const int img = kt / p.tiles_per_img;
            const bool chunk_end = (i == ntiles - 1) || ((kt + 1) / p.tiles_per_img != img);
            if (!chunk_end) continue;
            const int buf = chunk & 1;
            const uint32_t ph = ((uint32_t)(chunk >> 1)) & 1u;
            ++chunk;
            if (!mbar_wait_warp(tmem_full + buf, ph, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 3);
                break;
            }
            tc_fence_after();
            __syncwarp();
            const int ncb = p.cin >> 6;
            const float inv_x = c < p.cin ? 1.f / pow2_scale_for(__uint_as_float(__ldg(p.amax_x + (size_t)img * ncb + (c >> 6)))) : 0.f;
            const float inv_g0 = 1.f / pow2_scale_for(__uint_as_float(__ldg(p.amax_g + (size_t)img * 2)));
            const float inv_g1 = 1.f / pow2_scale_for(__uint_as_float(__ldg(p.amax_g + (size_t)img * 2 + 1)));
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * kAccCols;
#pragma unroll
            for (int n0 = 0; n0 < kNPad; n0 += 16) {
                uint32_t v[16], w[16];
                tmem_ld_x16(taddr + n0, v);            // x1*g2 + x2*g1   (x 2^11)
                tmem_ld_x16(taddr + kNPad + n0, w);    // x1*g1
                tmem_ld_wait();
                const float inv_g = n0 < 64 ? inv_g0 : inv_g1;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const float t = fmaf(__uint_as_float(v[k]), kLoInv, __uint_as_float(w[k])) * inv_x;   // exact scaling
                    acc[n0 + k] = fmaf(t, inv_g, acc[n0 + k]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + buf);
        }
        if (c < p.cin) {
            float4 *dst = reinterpret_cast<float4 *>(p.partial + (((size_t)slice * 9 + tap) * p.cin + c) * kNPad);
#pragma unroll
            for (int n = 0; n < kNPad; n += 4) dst[n >> 2] = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---- three taps per CTA ---------------------------------------------------------------------------------------------
// The feature tile is the same for all nine taps (only G moves), and the single-tap kernel above is L2 -> SM bound
// (tensor pipe 39 % active).  Here a CTA owns (128-channel block, column shift dx, slice) and runs the THREE row shifts
// dy against one load of the feature tile: 32 KB of X + 3 x 20 KB of G per 64-pixel step instead of 3 x 52 KB.  Three
// accumulators of 160 TMEM columns (single buffered: the issuer waits for the twelve accumulate warps once per image),
// 14 warps: TMA producer, MMA issuer, 3 groups of 4 accumulate warps (group t owns the accumulator of dy = t - 1).
constexpr int kTaps3 = 3;
constexpr int kStage3Bytes = 2 * kABytes + kTaps3 * 2 * kBBytes;   // 92 KB
constexpr int kStages3 = 2;
constexpr int kThreads3 = 64 + kTaps3 * 128;                       // 448

__global__ void __launch_bounds__(kThreads3, 1)
wgrad_tc3_kernel(const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                 const __grid_constant__ CUtensorMap map_g1a, const __grid_constant__ CUtensorMap map_g1b,
                 const __grid_constant__ CUtensorMap map_g1c, const __grid_constant__ CUtensorMap map_g2a,
                 const __grid_constant__ CUtensorMap map_g2b, const __grid_constant__ CUtensorMap map_g2c, const WgParams p) {
    constexpr uint32_t kIdescCat = umma_idesc_f16(128, 2 * kNPad);
    constexpr uint32_t kIdescOne = umma_idesc_f16(128, kNPad);
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *ctrl = smem + (size_t)kStages3 * kStage3Bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(ctrl);      // [2]
    uint64_t *empty = full + 2;                               // [2]
    uint64_t *tmem_full = empty + 2;                          // [1]
    uint64_t *tmem_empty = tmem_full + 1;                     // [1]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 1);
    volatile int *abort_flag = reinterpret_cast<volatile int *>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x, dxi = blockIdx.y, slice = blockIdx.z;   // dxi = dx + 1
    const int k0 = (int)((long long)p.total_tiles * slice / p.nslice), k1 = (int)((long long)p.total_tiles * (slice + 1) / p.nslice);
    const int ntiles = k1 - k0;

    if (threadIdx.x == 0) {
        *abort_flag = 0;
        for (int s = 0; s < kStages3; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 4 * kTaps3);
        fence_barrier_init();
        fence_proxy_async();
    }
    const CUtensorMap *map_g1 = dxi == 0 ? &map_g1a : (dxi == 1 ? &map_g1b : &map_g1c);
    const CUtensorMap *map_g2 = dxi == 0 ? &map_g2a : (dxi == 1 ? &map_g2b : &map_g2c);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x1);
        tma_prefetch_desc(&map_x2);
        tma_prefetch_desc(map_g1);
        tma_prefetch_desc(map_g2);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < ntiles; ++i) {
            const int kt = k0 + i;
            const int img = kt / p.tiles_per_img, t = kt - img * p.tiles_per_img;
            if (!mbar_wait_warp(empty + s, ph ^ 1u, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 1);
                break;
            }
            if (elect_one_sync()) {
                uint8_t *st = smem + (size_t)s * kStage3Bytes;
                mbar_arrive_expect_tx(full + s, kStage3Bytes);
                const int q0 = t * kPixTile;
                tma_load_3d(&map_x1, full + s, st, q0, mt * kMTile, img);
                tma_load_3d(&map_x2, full + s, st + kABytes, q0, mt * kMTile, img);
#pragma unroll
                for (int ti = 0; ti < kTaps3; ++ti) {   // dy = ti - 1: G^T origin q0 - dy*gwp
                    uint8_t *bt = st + 2 * kABytes + ti * 2 * kBBytes;
                    tma_load_3d(map_g2, full + s, bt, q0 - (ti - 1) * p.gwp, 0, img);
                    tma_load_3d(map_g1, full + s, bt + kBBytes, q0 - (ti - 1) * p.gwp, 0, img);
                }
            }
            __syncwarp();
            if (++s == kStages3) {
                s = 0;
                ph ^= 1u;
            }
        }
    } else if (warp == 1) {
        int s = 0, chunk = 0;
        uint32_t ph = 0;
        bool fresh = true;
        int rem = k0 % p.tiles_per_img;   // tile inside the image, stepped without divisions
        for (int i = 0; i < ntiles; ++i) {
            const bool last_of_img = rem == p.tiles_per_img - 1;
            rem = last_of_img ? 0 : rem + 1;
            const bool chunk_end = (i == ntiles - 1) || last_of_img;
            if (fresh) {   // single accumulator set: wait until the previous image has been drained
                if (!mbar_wait_warp(tmem_empty, ((uint32_t)chunk & 1u) ^ 1u, abort_flag)) {
                    if (lane == 0) atomicCAS(p.status, 0, 4);
                    break;
                }
            }
            if (!mbar_wait_warp(full + s, ph, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 2);
                break;
            }
            tc_fence_after();
            const uint32_t st = smem_u32(smem + (size_t)s * kStage3Bytes);
            const uint64_t a1 = umma_desc_sw128(st), a2 = umma_desc_sw128(st + kABytes);
            if (elect_one_sync()) {
#pragma unroll
                for (int ti = 0; ti < kTaps3; ++ti) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)ti * kAccCols;
                    const uint64_t b_cat = umma_desc_sw128(st + 2 * kABytes + ti * 2 * kBBytes);
                    const uint64_t b_g1 = umma_desc_sw128(st + 2 * kABytes + ti * 2 * kBBytes + kBBytes);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                        umma_f16_ss(d_tmem, a1 + adv, b_cat + adv, kIdescCat, (fresh && ks == 0) ? 0u : 1u);
                        umma_f16_ss(d_tmem, a2 + adv, b_g1 + adv, kIdescOne, 1u);
                    }
                }
                umma_commit(empty + s);
                if (chunk_end) umma_commit(tmem_full);
            }
            __syncwarp();
            fresh = chunk_end;
            if (chunk_end) ++chunk;
            if (++s == kStages3) {
                s = 0;
                ph ^= 1u;
            }
        }
    } else {
        const int ti = (warp - 2) >> 2;                // accumulator / dy of this group
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int c = mt * kMTile + row;
        float acc[kNPad];
#pragma unroll
        for (int n = 0; n < kNPad; ++n) acc[n] = 0.f;
        int chunk = 0;
        // (image, tile inside the image) stepped incrementally: two integer divisions per K tile and warp were a third of
        // this kernel's instructions
        int img_it = k0 / p.tiles_per_img, rem = k0 - img_it * p.tiles_per_img;
        for (int i = 0; i < ntiles; ++i) {
            const int img = img_it;
            const bool last_of_img = rem == p.tiles_per_img - 1;
            if (last_of_img) {
                rem = 0;
                ++img_it;
            } else {
                ++rem;
            }
            const bool chunk_end = (i == ntiles - 1) || last_of_img;
            if (!chunk_end) continue;
            const uint32_t ph = (uint32_t)chunk & 1u;
            ++chunk;
            if (!mbar_wait_warp(tmem_full, ph, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 3);
                break;
            }
            tc_fence_after();
            __syncwarp();
            const int ncb = p.cin >> 6;
            const float inv_x = c < p.cin ? 1.f / pow2_scale_for(__uint_as_float(__ldg(p.amax_x + (size_t)img * ncb + (c >> 6)))) : 0.f;
            const float inv_g0 = 1.f / pow2_scale_for(__uint_as_float(__ldg(p.amax_g + (size_t)img * 2)));
            const float inv_g1 = 1.f / pow2_scale_for(__uint_as_float(__ldg(p.amax_g + (size_t)img * 2 + 1)));
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ti * kAccCols;
#pragma unroll
            for (int n0 = 0; n0 < kNPad; n0 += 16) {
                uint32_t v[16], w[16];
                tmem_ld_x16(taddr + n0, v);
                tmem_ld_x16(taddr + kNPad + n0, w);
                tmem_ld_wait();
                const float inv_g = n0 < 64 ? inv_g0 : inv_g1;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const float t = fmaf(__uint_as_float(v[k]), kLoInv, __uint_as_float(w[k])) * inv_x;
                    acc[n0 + k] = fmaf(t, inv_g, acc[n0 + k]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
        }
        {
            // The warp's 32 rows x 80 floats are one contiguous 10 KB block of the partial buffer.  One thread per row
            // would touch 32 lines per store instruction (ncu: 7.5 % of the kernel's stall samples were LSU throttling
            // here); the stage ring is idle by now -- the last chunk's tmem_full means every MMA, hence every TMA load, has
            // completed -- so the rows are transposed through it and written as 512 contiguous bytes per instruction.
            constexpr int kPitch = kNPad + 4;
            float *stg = reinterpret_cast<float *>(smem) + (size_t)(warp - 2) * (32 * kPitch);
#pragma unroll
            for (int n = 0; n < kNPad; n += 4)
                *reinterpret_cast<float4 *>(stg + lane * kPitch + n) = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
            __syncwarp();
            const int tap = ti * 3 + dxi;              // dy-major tap index of the weight tensor
            const int c_warp = mt * kMTile + q * 32;   // first channel row of this warp
            float4 *dst = reinterpret_cast<float4 *>(p.partial + (((size_t)slice * 9 + tap) * p.cin + c_warp) * kNPad);
            constexpr int kQuadsPerRow = kNPad / 4;
#pragma unroll
            for (int it = 0; it < kQuadsPerRow; ++it) {
                const int f = it * 32 + lane, r = f / kQuadsPerRow, q4 = f - r * kQuadsPerRow;
                if (c_warp + r < p.cin) dst[f] = *reinterpret_cast<const float4 *>(stg + r * kPitch + 4 * q4);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 512);
    }
}

// dW[n][c][tap] = sum_slices partial[slice][tap][c][n]   (fixed order; reads coalesced along n, 4-byte scattered writes)
__global__ void wgrad_tc_reduce_kernel(const float *__restrict__ partial, int nslice, int cin, int cout, float *__restrict__ gw_out) {
    const size_t n_in = (size_t)9 * cin * kNPad;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % kNPad);
        const size_t r = i / kNPad;
        const int c = (int)(r % cin), tap = (int)(r / cin);
        if (n >= cout) continue;
        float s = 0.f;
        for (int k = 0; k < nslice; ++k) s += partial[(size_t)k * n_in + i];
        gw_out[((size_t)n * cin + c) * 9 + tap] = s;
    }
}

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

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
int gwp_of(int gw) { return (gw + 1 + 7) & ~7; }   // at least one zero pad column, rows a multiple of 16 bytes

// debug aid (SQD_WG_SYNC=1): synchronise after every stage so that a faulting kernel is named
int stage_check(const char *what, cudaStream_t st) {
    if (!sqd_opt(SQD_OPT_WG_SYNC)) return SQD_OK;
    cudaError_t err = cudaStreamSynchronize(st);
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) {
        sqd_set_error("wgrad_tc stage '%s' failed: %s", what, cudaGetErrorString(err));
        return (int)err;
    }
    return SQD_OK;
}

struct WgWs {
    size_t status_off, amax_x_off, amax_g_off, x1_off, x2_off, g1_off, g2_off, partial_off, total;
    int nslice;
};
WgWs wg_ws(int batch, int cin, int gh, int gw) {
    WgWs w;
    const int gwp = gwp_of(gw);
    const int tiles = batch * ((gh * gwp + kPixTile - 1) / kPixTile);
    // slices: three waves of work items (channel block x taps), at least ~8 pixel tiles each
    const int items = ((cin + kMTile - 1) / kMTile) * (sqd_opt(SQD_OPT_WG_SINGLE_TAP) ? 9 : 3);
    int ns = (3 * SQD_SM_COUNT) / items;   // <= 3 full waves of one CTA per SM (a 4th, nearly empty wave costs a full wave time)
    if (ns > tiles / 8) ns = tiles / 8;
    if (ns < 1) ns = 1;
    w.nslice = ns;
    size_t off = 0;
    w.status_off = off;  off += 256;
    w.amax_x_off = off;  off += align256((size_t)batch * (cin / 64) * sizeof(unsigned));
    w.amax_g_off = off;  off += align256((size_t)batch * 2 * sizeof(unsigned));
    const size_t xplane = align256((size_t)batch * cin * gh * gwp * sizeof(__half));
    const size_t gplane = align256((size_t)3 * batch * kNPad * gh * gwp * sizeof(__half));   // three column shifts
    w.x1_off = off;  off += xplane;
    w.x2_off = off;  off += xplane;
    w.g1_off = off;  off += gplane;
    w.g2_off = off;  off += gplane;
    w.partial_off = off;  off += align256((size_t)ns * 9 * cin * kNPad * sizeof(float));
    w.total = off;
    return w;
}

}  // namespace

// implemented in convdet_f16.cu: max |x| of contiguous runs (one per blockIdx.y)
int sqd_f16_absmax_runs(const float *d_in, size_t run_floats, int nruns, unsigned *d_amax, cudaStream_t st);
int sqd_gpred_absmax(const float *d_gpred, int batch, int P, int cout, int ncb, unsigned *amax_bits, cudaStream_t st,
                     int per_image);

extern "C" size_t sqd_convdet_wgrad_tc_workspace_bytes(int batch, int cin, int gh, int gw, int cout) {
    if (batch <= 0 || cin <= 0 || gh <= 0 || gw <= 0 || cout <= 0) return 256;
    return wg_ws(batch, cin, gh, gw).total;
}

extern "C" int sqd_convdet_wgrad_tc(const float *d_feat_nchw, const float *d_gpred, int batch, int cin, int gh, int gw, int cout,
                                    float *d_gweight, void *d_workspace, size_t workspace_bytes, void *stream) {
    SQD_REQUIRE(d_feat_nchw && d_gpred && d_gweight && d_workspace, SQD_E_NULL, "sqd_convdet_wgrad_tc: NULL pointer");
    SQD_REQUIRE(batch >= 1 && cin >= 64 && cin % 64 == 0 && gh > 0 && gw > 0 && cout >= 1 && cout <= kNPad, SQD_E_SHAPE,
                "sqd_convdet_wgrad_tc: bad shape (Cin a multiple of 64, Cout <= %d)", kNPad);
    SQD_REQUIRE(sqd_aligned16(d_workspace), SQD_E_ALIGN, "sqd_convdet_wgrad_tc: workspace must be 16-byte aligned");
    const WgWs w = wg_ws(batch, cin, gh, gw);
    SQD_REQUIRE(workspace_bytes >= w.total, SQD_E_WORKSPACE, "sqd_convdet_wgrad_tc: workspace too small (%zu < %zu bytes)",
                workspace_bytes, w.total);
    EncodeTiledFn encode = get_encode_fn();
    SQD_REQUIRE(encode != nullptr, SQD_E_DRIVER, "cuTensorMapEncodeTiled not available from the CUDA driver");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char *ws = static_cast<char *>(d_workspace);
    const int gwp = gwp_of(gw), P = gh * gw, ncb = cin / 64;
    unsigned *amax_x = reinterpret_cast<unsigned *>(ws + w.amax_x_off), *amax_g = reinterpret_cast<unsigned *>(ws + w.amax_g_off);
    __half *x1 = reinterpret_cast<__half *>(ws + w.x1_off), *x2 = reinterpret_cast<__half *>(ws + w.x2_off);
    __half *g1 = reinterpret_cast<__half *>(ws + w.g1_off), *g2 = reinterpret_cast<__half *>(ws + w.g2_off);
    SQD_CUDA(cudaMemsetAsync(ws, 0, w.x1_off, st));   // status + both amax arrays

    // 1. operands: pixel-major fp16 planes
    int rc = SQD_OK;
    SQD_REQUIRE(gw % 2 == 0, SQD_E_SHAPE, "sqd_convdet_wgrad_tc: grid width %d must be even", gw);
    if (kXsChan * gh * (gwp / 8) <= kXsItems * kXsThreads && !sqd_opt(SQD_OPT_BWD_OLD_PREPASS)) {
        split_nchw_rows_cluster_kernel<<<(unsigned)(batch * ncb * kXsCluster), kXsThreads, 0, st>>>(
            d_feat_nchw, cin, gh, gw, gwp, amax_x, reinterpret_cast<uint4 *>(x1), reinterpret_cast<uint4 *>(x2));
        SQD_LAUNCH_CHECK("split_nchw_rows_cluster_kernel");
    } else {
        rc = sqd_f16_absmax_runs(d_feat_nchw, (size_t)64 * P, batch * ncb, amax_x, st);
        if (rc) return rc;
        if ((rc = stage_check("absmax x", st))) return rc;
        split_nchw_rows_kernel<<<(unsigned)((size_t)batch * cin), 256, 0, st>>>(d_feat_nchw, cin, gh, gw, gwp, amax_x,
                                                                               reinterpret_cast<__half2 *>(x1),
                                                                               reinterpret_cast<__half2 *>(x2));
        SQD_LAUNCH_CHECK("split_nchw_rows_kernel");
    }
    if ((rc = stage_check("split x", st))) return rc;
    if ((rc = sqd_gpred_absmax(d_gpred, batch, P, cout, 2, amax_g, st, 0))) return rc;
    if ((rc = stage_check("absmax g", st))) return rc;
    {
        const size_t smem_g = (size_t)gw * (cout + 1) * sizeof(float);
        if (smem_g > 48 * 1024)
            SQD_CUDA(cudaFuncSetAttribute(g_transpose_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
        g_transpose_split_kernel<<<dim3(gh, batch), 256, smem_g, st>>>(d_gpred, gh, gw, gwp, cout, batch, amax_g, g1, g2);
        SQD_LAUNCH_CHECK("g_transpose_split_kernel");
        if ((rc = stage_check("transpose g", st))) return rc;
    }

    // 2. tensor maps over the flat planes: {padded pixel, channel / output, image}; maps 2..4 = g1 shifted -1,0,+1, 5..7 = g2
    alignas(64) CUtensorMap maps[8];
    const size_t ppad = (size_t)gh * gwp;
    for (int i = 0; i < 8; ++i) {
        const int chan = i < 2 ? cin : kNPad;
        void *base = i == 0 ? (void *)x1 : i == 1 ? (void *)x2
                   : i < 5 ? (void *)(g1 + (size_t)(i - 2) * batch * kNPad * ppad) : (void *)(g2 + (size_t)(i - 5) * batch * kNPad * ppad);
        const cuuint64_t dims[3] = {(cuuint64_t)ppad, (cuuint64_t)chan, (cuuint64_t)batch};
        const cuuint64_t strides[2] = {(cuuint64_t)ppad * 2, (cuuint64_t)chan * ppad * 2};
        const cuuint32_t box[3] = {kPixTile, (cuuint32_t)(i < 2 ? kMTile : kNPad), 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SQD_REQUIRE(r == CUDA_SUCCESS, SQD_E_DRIVER, "cuTensorMapEncodeTiled(wgrad operand %d) failed: CUresult %d", i, (int)r);
    }

    // 3. the GEMM and the fixed-order reduction
    WgParams p;
    p.cin = cin; p.gh = gh; p.gw = gw; p.cout = cout; p.batch = batch;
    p.gwp = gwp;
    p.tiles_per_img = (gh * gwp + kPixTile - 1) / kPixTile;
    p.total_tiles = batch * p.tiles_per_img;
    p.nslice = w.nslice;
    p.amax_x = amax_x;
    p.amax_g = amax_g;
    p.partial = reinterpret_cast<float *>(ws + w.partial_off);
    p.status = reinterpret_cast<int *>(ws + w.status_off);
    if (!sqd_opt(SQD_OPT_WG_SINGLE_TAP)) {
        // three row shifts per CTA: a third of the feature-tile traffic
        const size_t smem3 = 1024 + (size_t)kStages3 * kStage3Bytes + 1024;
        SQD_CUDA(cudaFuncSetAttribute(wgrad_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        const dim3 grid3((cin + kMTile - 1) / kMTile, 3, w.nslice);
        wgrad_tc3_kernel<<<grid3, kThreads3, smem3, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], p);
        SQD_LAUNCH_CHECK("wgrad_tc3_kernel");
    } else {
        const size_t smem = 1024 + (size_t)kStages * kStageBytes + 1024;
        SQD_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const dim3 grid((cin + kMTile - 1) / kMTile, 9, w.nslice);
        wgrad_tc_kernel<<<grid, kThreads, smem, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], p);
        SQD_LAUNCH_CHECK("wgrad_tc_kernel");
    }
    if ((rc = stage_check("gemm", st))) return rc;
    const size_t n_in = (size_t)9 * cin * kNPad;
    wgrad_tc_reduce_kernel<<<(int)((n_in + 255) / 256), 256, 0, st>>>(p.partial, w.nslice, cin, cout, d_gweight);
    SQD_LAUNCH_CHECK("wgrad_tc_reduce_kernel");
    if ((rc = stage_check("reduce", st))) return rc;
    return SQD_OK;
}

// 0 = the last sqd_convdet_wgrad_tc on this workspace drained cleanly (synchronises the stream)
extern "C" int sqd_convdet_wgrad_tc_status(const void *d_workspace, void *stream) {
    SQD_REQUIRE(d_workspace, SQD_E_NULL, "sqd_convdet_wgrad_tc_status: NULL workspace");
    int h = -1;
    SQD_CUDA(cudaMemcpyAsync(&h, d_workspace, sizeof(int), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    SQD_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    if (h != 0) sqd_set_error("tcgen05 wgrad pipeline timed out (role %d)", h);
    return h;
}

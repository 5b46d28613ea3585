// SURVEY 8(f) rank 2, first half: ConvDet backward with respect to the Fire11 features (dgrad) on the tcgen05
// forward kernel, plus the bias gradient.  Reference: autograd through nn.Conv2d(768, 72, 3, padding=1)
// (src/model/squeezedet.py:73-75,83-87; trainer.py:43-46 calls loss.backward()) -> cuDNN dgrad there.
//
// dX[b,y,x,c] = sum_{dy,dx,n} G[b, y+1-dy, x+1-dx, n] * W[n,c,dy,dx]
//             = a 3x3 pad-1 convolution of G (the gradient of pred, NHWC with Cout = K*(C+5) channels) with the
//               weights W2[c][n][dy'][dx'] = W[n][c][2-dy'][2-dx']   (taps flipped, channel roles swapped)
// so it is the forward implicit GEMM with roles swapped: "input channels" = Cout padded to a multiple of 64 (72 -> 128),
// "output channels" = Cin = 768 in slabs of 128 columns (six launches of convdet_f16_pair_kernel<128,0>, all reading the
// same small fp16 planes of G, which stay in L2).  Same f16x3 numerics as the forward (fp32-grade products).
// The result is written NHWC (B, gh, gw, Cin): the channels_last memory format of the logical (B, Cin, gh, gw) tensor.
// The weight gradient (wgrad: a contraction over B*gh*gw pixels) is NOT native yet -- see DESIGN.md.
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

size_t sqd_f16_packed_bytes(int cout, int cin);
int sqd_f16_pack_weights(const float *d_weight, int cout, int cin, void *d_packed, cudaStream_t st);
int sqd_f16_pack_dgrad_slabs(const float *d_weight, int cout, int cin, int slab_rows, int kp, int nslabs, void *d_packed,
                             size_t slab_bytes, float *d_scratch, cudaStream_t st);
size_t sqd_f16_split_bytes(int batch, int cin, int gh, int gw);
size_t sqd_f16_workspace_bytes(int batch, int cin, int gh, int gw, int cout, int layout);
int sqd_convdet_f16_pair(const float *d_feat, int layout, const void *d_packed, const float *d_bias, int batch, int cin,
                         int gh, int gw, int cout, float *d_pred, void *d_workspace, cudaStream_t st,
                         const SqdCandEmit *emit, int out_stride, int ksteps_last, int slabs, size_t slab_stride,
                         int image_scales);

namespace {

constexpr int kSlab = 128;  // output columns (feature channels) per GEMM launch

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
int kpad_of(int cout) { return (cout + 63) / 64 * 64; }
int nslab_of(int cin) { return (cin + kSlab - 1) / kSlab; }

// W (cout, cin, 3, 3) -> W2 slab (kSlab, kp, 3, 3): W2[j][n][t] = W[n][s*kSlab + j][8 - t], zero outside
__global__ void flip_transpose_weights_kernel(const float *__restrict__ w, int cout, int cin, int kp, int slab,
                                              float *__restrict__ w2) {
    const int total = kSlab * kp * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int t = i % 9, n = (i / 9) % kp, j = i / (9 * kp);
        const int c = slab * kSlab + j;
        w2[i] = (n < cout && c < cin) ? w[((size_t)n * cin + c) * 9 + (8 - t)] : 0.f;
    }
}

// power-of-two scale s with amax*s in [2^13, 2^14) -- must match convdet_f16.cu (the GEMM divides it out again)
__device__ __forceinline__ float pow2_scale_for(float amax) {
    if (!(amax > 0.f) || amax > 3.0e38f) return 1.f;
    int ex;
    frexpf(amax, &ex);
    int e = 14 - ex;
    e = e < -126 ? -126 : (e > 126 ? 126 : e);
    return ldexpf(1.f, e);
}

// max |g| per (image, 64-channel block of the padded channel axis): one thread per cell walks the cell's channels
// block by block (register maximum), a warp reduction and one shared atomic per warp and block follow
__global__ void __launch_bounds__(256) gpred_absmax_kernel(const float *__restrict__ g, int P, int cout, int ncb,
                                                           unsigned *__restrict__ amax_bits) {
    extern __shared__ unsigned s_max[];   // ncb
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < ncb; i += blockDim.x) s_max[i] = 0u;
    __syncthreads();
    const float *src = g + (size_t)b * P * cout;
    for (int cell0 = blockIdx.x * blockDim.x; cell0 < P; cell0 += gridDim.x * blockDim.x) {   // warp-uniform trip count
        const int cell = cell0 + threadIdx.x;
        for (int cb = 0; cb < ncb; ++cb) {
            float m = 0.f;
            if (cell < P) {
                const int c1 = min(cout, (cb + 1) * 64);
                for (int ch = cb * 64; ch < c1; ++ch) m = fmaxf(m, fabsf(__ldg(src + (size_t)cell * cout + ch)));
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if ((threadIdx.x & 31) == 0) atomicMax(&s_max[cb], __float_as_uint(m));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ncb; i += blockDim.x) atomicMax(amax_bits + (size_t)b * ncb + i, s_max[i]);
}

// The same maxima for cout % 4 == 0 (KITTI: 72): an image's (P, cout) block read as a flat run of 16-byte quads, fully
// coalesced; a quad never straddles a pixel or a 64-channel block.  Grid (chunks, B); at most 2 blocks (cout <= 128).
__global__ void __launch_bounds__(256) gpred_absmax_flat_kernel(const float4 *__restrict__ g, int quads_per_image, int qpp,
                                                                int ncb, unsigned *__restrict__ amax_bits, int per_image) {
    __shared__ unsigned s_max[2];
    if (threadIdx.x < 2) s_max[threadIdx.x] = 0u;
    __syncthreads();
    const int b = blockIdx.y;
    const float4 *src = g + (size_t)b * quads_per_image;
    float m0 = 0.f, m1 = 0.f;
    const int stride = gridDim.x * blockDim.x;
#pragma unroll 4
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < quads_per_image; q += stride) {
        const float4 v = __ldg(src + q);
        const float m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
        if ((q % qpp) < 16) m0 = fmaxf(m0, m); else m1 = fmaxf(m1, m);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&s_max[0], __float_as_uint(m0));
        atomicMax(&s_max[1], __float_as_uint(m1));
    }
    __syncthreads();
    // per_image: every block of the image gets the image maximum (one scale per image: the GEMM may then accumulate a
    // whole tile in TMEM before folding it)
    if ((int)threadIdx.x < ncb)
        atomicMax(amax_bits + (size_t)b * ncb + threadIdx.x, per_image ? max(s_max[0], s_max[1]) : s_max[threadIdx.x]);
}

// G (B, P, cout) fp32 -> x1 / x2 fp16 planes (B, P, kp), channels >= cout zero; one thread per (cell, 4 channels)
__global__ void __launch_bounds__(256) gpred_split_pad_kernel(const float *__restrict__ g, int P, int cout, int kp,
                                                              const unsigned *__restrict__ amax_bits, uint2 *__restrict__ p1,
                                                              uint2 *__restrict__ p2, size_t total) {
    const int kq = kp >> 2, ncb = kp >> 6;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % kq);
        const size_t cell = i / kq;           // b*P + pixel
        const int b = (int)(cell / P);
        const float s = pow2_scale_for(__uint_as_float(amax_bits[(size_t)b * ncb + (q >> 4)]));
        unsigned short h1[4], h2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int ch = 4 * q + e;
            const float xs = ch < cout ? g[cell * cout + ch] * s : 0.f;
            const __half a = __float2half_rn(xs);
            const __half r = __float2half_rn((xs - __half2float(a)) * 2048.f);
            h1[e] = __half_as_ushort(a);
            h2[e] = __half_as_ushort(r);
        }
        p1[i] = make_uint2((unsigned)h1[0] | ((unsigned)h1[1] << 16), (unsigned)h1[2] | ((unsigned)h1[3] << 16));
        p2[i] = make_uint2((unsigned)h2[0] | ((unsigned)h2[1] << 16), (unsigned)h2[2] | ((unsigned)h2[3] << 16));
    }
}

// db[n] = sum over (b, pixel) of G[b, pixel, n]: fixed-order two-level sum (deterministic), double accumulation
__global__ void __launch_bounds__(256) bias_grad_kernel(const float *__restrict__ g, size_t rows, int cout,
                                                        float *__restrict__ db) {
    const int n = blockIdx.x;
    double acc = 0.0;
    for (size_t r = threadIdx.x; r < rows; r += blockDim.x) acc += (double)g[r * cout + n];
    __shared__ double s[256];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) db[n] = (float)s[0];
}

// The same sums for cout % 4 == 0: coalesced 16-byte loads instead of one strided column per CTA.  A cluster of 8 CTAs
// owns a group of 6 quads (24 columns, 96 B = three whole sectors of every 4*cout-byte row) and splits the rows; a
// thread sums one quad over every 170th row in double, the CTA adds its 170 row lanes in lane order, CTA 0 of the
// cluster adds the 8 CTA sums in rank order through distributed shared memory: a fixed order, deterministic, no scratch.
constexpr int kBgThreads = 1024, kBgQuads = 6, kBgCluster = 8, kBgLanes = kBgThreads / kBgQuads;   // 170 row lanes
__global__ void __cluster_dims__(kBgCluster, 1, 1) __launch_bounds__(kBgThreads)
    bias_grad_cluster_kernel(const float4 *__restrict__ g, size_t rows, int qpp, float *__restrict__ db) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double s_lane[kBgLanes][kBgQuads * 4];
    __shared__ double s_cta[kBgCluster][kBgQuads * 4];
    const int rank = (int)cluster.block_rank(), group = blockIdx.x / kBgCluster;
    const int q0 = group * kBgQuads, nq = min(kBgQuads, qpp - q0);
    const int lq = threadIdx.x % kBgQuads, lr = threadIdx.x / kBgQuads;
    const size_t per = (rows + kBgCluster - 1) / kBgCluster;
    const size_t r0 = (size_t)rank * per, r1 = r0 + per < rows ? r0 + per : rows;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (lr < kBgLanes && lq < nq) {
#pragma unroll 8
        for (size_t r = r0 + lr; r < r1; r += kBgLanes) {
            const float4 v = __ldg(g + r * qpp + q0 + lq);
            a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
        }
    }
    if (lr < kBgLanes) {
        s_lane[lr][lq * 4 + 0] = a0; s_lane[lr][lq * 4 + 1] = a1; s_lane[lr][lq * 4 + 2] = a2; s_lane[lr][lq * 4 + 3] = a3;
    }
    __syncthreads();
    if ((int)threadIdx.x < kBgQuads * 4) {
        double t = 0.0;
        for (int l = 0; l < kBgLanes; ++l) t += s_lane[l][threadIdx.x];
        cluster.map_shared_rank(&s_cta[0][0], 0)[rank * kBgQuads * 4 + threadIdx.x] = t;
    }
    cluster.sync();
    if (rank == 0 && (int)threadIdx.x < nq * 4) {
        double t = 0.0;
        for (int r = 0; r < kBgCluster; ++r) t += s_cta[r][threadIdx.x];
        db[q0 * 4 + threadIdx.x] = (float)t;
    }
}

// ---- weight gradient (fp32 CUDA-core implicit GEMM, split over the pixel axis) ---------------------------------------
// dW[n,c,dy,dx] = sum_{b,y,x} G[b,y,x,n] * X[b,c,y+dy-1,x+dx-1]: M = Cout (72), N = 9*Cin (6912), K = B*gh*gw (37,440 at
// KITTI B = 20).  A CTA owns (one tap, 128 input channels, all Cout <= 80 output channels) for a slice of the image rows
// and walks it in 32-pixel row segments: X segment (128 ch x 32 px, NCHW rows are pixel-contiguous, shifted by the tap,
// zero outside the image = the conv padding) and G segment (32 px x Cout, NHWC rows) staged in shared memory, each of
// the 320 threads accumulating an 8 (channels) x 4 (outputs) register tile from three 16-byte shared loads per pixel.  Slices write fp32 partials that a second
// kernel adds in a fixed order (deterministic; no atomics).  Not a tensor-core kernel: the tcgen05 version (pixel-major
// operands, chunked TMEM accumulation per image) is round-2 work; this one removes the cuDNN dependency of the training
// mirror and is the yardstick it will be checked against.
constexpr int kWgC = 128, kWgN = 80, kWgPix = 32, kWgThreads = 320;  // 16 channel octets x 20 output quads
constexpr int kWgXs = kWgC + 4;                                        // padded row: 16-byte aligned, conflict-free LDS.128
constexpr int kWgXLoads = (kWgC * kWgPix + kWgThreads - 1) / kWgThreads;   // 13 scalar X loads per thread and segment
constexpr int kWgGLoads = (kWgPix * (kWgN / 4) + kWgThreads - 1) / kWgThreads;  // 2 float4 G loads per thread and segment

// Global -> register fetch of one 32-pixel segment (software pipelined: issued before the previous segment is
// multiplied).  X: warp w covers channels w, w+10, ... of the CTA's 128, lane = pixel of the segment, shifted by the
// tap and zero outside the image (= the conv padding).  G: 32 pixels x Cout floats as float4 (Cout % 4 == 0).
struct WgFetch {
    float x[kWgXLoads];
    float4 g[kWgGLoads];
};

// n0: first output channel of this launch's chunk (<= kWgN channels per launch); rows of G that are not 16-byte
// aligned for the chunk (Cout or n0 not a multiple of 4, e.g. the stress head's 117 channels) are read with scalar loads.
__device__ __forceinline__ void wg_fetch(WgFetch &f, const float *__restrict__ x, const float *__restrict__ g, int b, int y,
                                         int x0, int dy, int dx, int c0, int cin, int gh, int gw, int cout, int n0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ys = y + dy, xsrc = x0 + lane + dx;
    const bool ok = ys >= 0 && ys < gh && xsrc >= 0 && xsrc < gw && x0 + lane < gw;
    const float *xp = x + (((size_t)b * cin + c0 + warp) * gh + (ok ? ys : 0)) * gw + (ok ? xsrc : 0);
    const size_t cstep = (size_t)(kWgThreads / 32) * gh * gw;
#pragma unroll
    for (int j = 0; j < kWgXLoads; ++j) {
        const int c = warp + (kWgThreads / 32) * j;
        f.x[j] = (ok && c < kWgC && c0 + c < cin) ? __ldg(xp + j * cstep) : 0.f;
    }
    const float *gbase = g + (((size_t)b * gh + y) * gw + x0) * cout + n0;
    const int npix = min(kWgPix, gw - x0), nn = min(kWgN, cout - n0);
    if (((cout | n0) & 3) == 0) {
        const float4 *grow = reinterpret_cast<const float4 *>(gbase);
        const int q4 = cout >> 2, qn = nn >> 2;
#pragma unroll
        for (int j = 0; j < kWgGLoads; ++j) {
            const int i = threadIdx.x + j * kWgThreads;
            const int p = i / (kWgN / 4), q = i - p * (kWgN / 4);
            f.g[j] = (p < npix && q < qn) ? __ldg(grow + (size_t)p * q4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
#pragma unroll
        for (int j = 0; j < kWgGLoads; ++j) {
            const int i = threadIdx.x + j * kWgThreads;
            const int p = i / (kWgN / 4), q = i - p * (kWgN / 4);
            const float *src = gbase + (size_t)p * cout + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < npix) {
                if (4 * q < nn) v.x = __ldg(src);
                if (4 * q + 1 < nn) v.y = __ldg(src + 1);
                if (4 * q + 2 < nn) v.z = __ldg(src + 2);
                if (4 * q + 3 < nn) v.w = __ldg(src + 3);
            }
            f.g[j] = v;
        }
    }
}

__global__ void __launch_bounds__(kWgThreads) wgrad_partial_kernel(const float *__restrict__ x, const float *__restrict__ g,
                                                                   int batch, int cin, int gh, int gw, int cout, int nslice,
                                                                   float *__restrict__ partial, int n0) {
    __shared__ __align__(16) float xs[kWgPix][kWgXs];  // [pixel][channel]
    __shared__ __align__(16) float gs[kWgPix][kWgN];   // [pixel][output channel]
    const int tap = blockIdx.x, cb = blockIdx.y, slice = blockIdx.z;
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int c0 = cb * kWgC;
    const int tc = threadIdx.x & 15, tn = threadIdx.x >> 4;   // channels 4*tc.. and 64+4*tc.., outputs 4*tn..
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int rows = batch * gh;                               // image rows in the whole batch
    const int r_begin = (int)((long long)rows * slice / nslice), r_end = (int)((long long)rows * (slice + 1) / nslice);
    const int segs_per_row = (gw + kWgPix - 1) / kWgPix;
    const int nseg = (r_end - r_begin) * segs_per_row;
    WgFetch f;
    if (nseg > 0) wg_fetch(f, x, g, r_begin / gh, r_begin % gh, 0, dy, dx, c0, cin, gh, gw, cout, n0);
    for (int sgi = 0; sgi < nseg; ++sgi) {
        const int r = r_begin + sgi / segs_per_row, x0 = (sgi % segs_per_row) * kWgPix;
        const int npix = min(kWgPix, gw - x0);
        // registers -> shared memory
#pragma unroll
        for (int j = 0; j < kWgXLoads; ++j) {
            const int c = warp + (kWgThreads / 32) * j;
            if (c < kWgC) xs[lane][c] = f.x[j];
        }
#pragma unroll
        for (int j = 0; j < kWgGLoads; ++j) {
            const int i = threadIdx.x + j * kWgThreads;
            const int p = i / (kWgN / 4), q = i - p * (kWgN / 4);
            if (p < kWgPix) *reinterpret_cast<float4 *>(&gs[p][4 * q]) = f.g[j];
        }
        __syncthreads();
        if (sgi + 1 < nseg) {   // next segment's loads fly while this one is multiplied
            const int rn = r_begin + (sgi + 1) / segs_per_row;
            wg_fetch(f, x, g, rn / gh, rn % gh, ((sgi + 1) % segs_per_row) * kWgPix, dy, dx, c0, cin, gh, gw, cout, n0);
        }
#pragma unroll 2
        for (int p = 0; p < npix; ++p) {
            const float4 a = *reinterpret_cast<const float4 *>(&xs[p][4 * tc]);
            const float4 a2 = *reinterpret_cast<const float4 *>(&xs[p][64 + 4 * tc]);
            const float4 bq = *reinterpret_cast<const float4 *>(&gs[p][4 * tn]);
            const float av[8] = {a.x, a.y, a.z, a.w, a2.x, a2.y, a2.z, a2.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
        (void)r;
    }
    // partial[slice][n][c][tap]  (the layout of the weight tensor, so the reduction is a plain strided sum)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + (i < 4 ? 4 * tc + i : 64 + 4 * tc + i - 4), n = n0 + 4 * tn + j;
            if (c < cin && n < cout) partial[(((size_t)slice * cout + n) * cin + c) * 9 + tap] = acc[i][j];
        }
}

__global__ void wgrad_reduce_kernel(const float *__restrict__ partial, size_t n, int nslice, float *__restrict__ gw_out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < nslice; ++k) s += partial[(size_t)k * n + i];   // fixed order
        gw_out[i] = s;
    }
}

int wgrad_slices(int batch, int gh) {
    int s = (batch * gh + 29) / 30;     // ~30 image rows per slice
    return s < 1 ? 1 : (s > 32 ? 32 : s);
}

struct DgradWs {
    size_t planes_off, gemm_off, total;
};
DgradWs dgrad_ws(int batch, int cin, int gh, int gw, int cout) {
    DgradWs w;
    const int kp = kpad_of(cout);
    w.planes_off = 0;
    w.gemm_off = align256(sqd_f16_split_bytes(batch, kp, gh, gw));
    w.total = w.gemm_off + align256(sqd_f16_workspace_bytes(batch, kp, gh, gw, kSlab, SQD_LAYOUT_SPLIT_NHWC));
    (void)cin;
    return w;
}

}  // namespace

extern "C" size_t sqd_convdet_dgrad_packed_bytes(int cout, int cin) {
    if (cout <= 0 || cin <= 0) return 0;
    const int kp = kpad_of(cout);
    return (size_t)nslab_of(cin) * align256(sqd_f16_packed_bytes(kSlab, kp)) + align256((size_t)kSlab * kp * 9 * sizeof(float));
}

extern "C" int sqd_convdet_dgrad_pack_weights(const float *d_weight, int cout, int cin, void *d_packed, void *stream) {
    SQD_REQUIRE(d_weight && d_packed, SQD_E_NULL, "sqd_convdet_dgrad_pack_weights: NULL pointer");
    SQD_REQUIRE(cout >= 1 && cout <= 1024 && cin >= 1, SQD_E_SHAPE, "sqd_convdet_dgrad_pack_weights: bad shape (%d, %d)", cout, cin);
    SQD_REQUIRE(sqd_aligned16(d_packed), SQD_E_ALIGN, "sqd_convdet_dgrad_pack_weights: buffer must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int kp = kpad_of(cout), ns = nslab_of(cin);
    const size_t slab_bytes = align256(sqd_f16_packed_bytes(kSlab, kp));
    char *base = static_cast<char *>(d_packed);
    if (!sqd_opt(SQD_OPT_DGRAD_PACK_LOOP))   // default: every slab in two launches, straight from W (same bytes)
        return sqd_f16_pack_dgrad_slabs(d_weight, cout, cin, kSlab, kp, ns, d_packed, slab_bytes,
                                        reinterpret_cast<float *>(base + (size_t)ns * slab_bytes), st);
    float *tmp = reinterpret_cast<float *>(base + (size_t)ns * slab_bytes);
    for (int s = 0; s < ns; ++s) {
        flip_transpose_weights_kernel<<<SQD_SM_COUNT, 256, 0, st>>>(d_weight, cout, cin, kp, s, tmp);
        SQD_LAUNCH_CHECK("flip_transpose_weights_kernel");
        int rc = sqd_f16_pack_weights(tmp, kSlab, kp, base + (size_t)s * slab_bytes, st);
        if (rc) return rc;
    }
    return SQD_OK;
}

// max |G| per (image, 64-channel block of the channel axis padded to a multiple of 64) into amax_bits (zeroed by the
// caller; image stride ncb >= ceil(cout / 64)): the scale granularity of both the dgrad planes and the wgrad G^T copies.
int sqd_gpred_absmax(const float *d_gpred, int batch, int P, int cout, int ncb, unsigned *amax_bits, cudaStream_t st,
                     int per_image) {
    if (cout % 4 == 0 && cout <= 128 && ncb <= 2 && (per_image || !sqd_opt(SQD_OPT_BWD_OLD_PREPASS))) {
        gpred_absmax_flat_kernel<<<dim3(16, batch), 256, 0, st>>>(reinterpret_cast<const float4 *>(d_gpred), P * (cout / 4),
                                                                 cout / 4, ncb, amax_bits, per_image);
        SQD_LAUNCH_CHECK("gpred_absmax_flat_kernel");
    } else {
        gpred_absmax_kernel<<<dim3(16, batch), 256, ncb * sizeof(unsigned), st>>>(d_gpred, P, cout, ncb, amax_bits);
        SQD_LAUNCH_CHECK("gpred_absmax_kernel");
    }
    return SQD_OK;
}

extern "C" size_t sqd_convdet_dgrad_workspace_bytes(int batch, int cin, int gh, int gw, int cout) {
    if (batch <= 0 || cin <= 0 || gh <= 0 || gw <= 0 || cout <= 0) return 256;
    return dgrad_ws(batch, cin, gh, gw, cout).total;
}

extern "C" int sqd_convdet_dgrad(const float *d_gpred, const void *d_dgrad_packed, int batch, int cin, int gh, int gw,
                                 int cout, float *d_gfeat_nhwc, void *d_workspace, size_t workspace_bytes, void *stream) {
    if (batch == 0) return SQD_OK;
    SQD_REQUIRE(d_gpred && d_dgrad_packed && d_gfeat_nhwc && d_workspace, SQD_E_NULL, "sqd_convdet_dgrad: NULL pointer");
    SQD_REQUIRE(batch > 0 && batch <= 65535 && cin >= kSlab && cin % kSlab == 0 && gh > 0 && gw > 0 && cout >= 1 && cout <= 1024,
                SQD_E_SHAPE, "sqd_convdet_dgrad: bad shape (Cin must be a multiple of %d)", kSlab);
    SQD_REQUIRE(sqd_aligned16(d_gpred) && sqd_aligned16(d_gfeat_nhwc) && sqd_aligned16(d_workspace), SQD_E_ALIGN,
                "sqd_convdet_dgrad: pointers must be 16-byte aligned");
    const DgradWs w = dgrad_ws(batch, cin, gh, gw, cout);
    SQD_REQUIRE(workspace_bytes >= w.total, SQD_E_WORKSPACE, "sqd_convdet_dgrad: workspace too small (%zu < %zu bytes)",
                workspace_bytes, w.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int kp = kpad_of(cout), ncb = kp / 64, P = gh * gw, ns = nslab_of(cin);
    // ONE scale per image for G when the flat max kernel applies (72 channels = 2 blocks): a tile of the dgrad GEMM is
    // then a single TMEM chunk (one drain of the 384-column accumulator per tile instead of two).  fp16 keeps 11
    // significant bits down to 2^-27 of the image maximum, so a per-block scale buys nothing for a gradient tensor.
    const int image_scales = (cout % 4 == 0 && cout <= 128 && ncb == 2 && !sqd_opt(SQD_OPT_DGRAD_BLOCK_SCALES)) ? 1 : 0;
    char *ws = static_cast<char *>(d_workspace);
    // 1. G -> zero-padded fp16 planes with per-(image, block) scales (G is 1/10 of the features: two small passes)
    char *planes = ws + w.planes_off;
    const size_t amax_bytes = align256((size_t)batch * ncb * sizeof(unsigned));
    const size_t plane_bytes = align256((size_t)batch * P * kp * sizeof(unsigned short));
    unsigned *amax = reinterpret_cast<unsigned *>(planes);
    SQD_CUDA(cudaMemsetAsync(amax, 0, amax_bytes, st));
    {
        int rc = sqd_gpred_absmax(d_gpred, batch, P, cout, ncb, amax, st, image_scales);
        if (rc) return rc;
    }
    const size_t total = (size_t)batch * P * (kp / 4);
    int gx = (int)((total + 255) / 256);
    if (gx > SQD_SM_COUNT * 16) gx = SQD_SM_COUNT * 16;
    gpred_split_pad_kernel<<<gx, 256, 0, st>>>(d_gpred, P, cout, kp, amax, reinterpret_cast<uint2 *>(planes + amax_bytes),
                                              reinterpret_cast<uint2 *>(planes + amax_bytes + plane_bytes), total);
    SQD_LAUNCH_CHECK("gpred_split_pad_kernel");
    // 2. one implicit GEMM per slab of 128 feature channels; the zero-padded K steps of the last gradient-channel block
    //    (72 channels: block 1 holds 8 real ones = one 16-channel step of four) are not issued
    const int rem = cout % 64;
    const int ksteps_last = rem == 0 ? 4 : (rem + 15) / 16;
    // ONE launch over all slabs: virtual image v = slab * B + b reads the planes of image b, the weights of `slab` and
    // writes feature channels [slab*128, +128).  (Six launches of one slab each spend most of their time in prologue,
    // pipeline fill and the last tile's epilogue: 2 tiles per CTA pair.)
    const size_t slab_bytes = align256(sqd_f16_packed_bytes(kSlab, kp));
    if (sqd_opt(SQD_OPT_DGRAD_PER_SLAB)) {
        for (int s = 0; s < ns; ++s) {
            int rc = sqd_convdet_f16_pair(reinterpret_cast<const float *>(planes), SQD_LAYOUT_SPLIT_NHWC,
                                          static_cast<const char *>(d_dgrad_packed) + (size_t)s * slab_bytes, nullptr, batch, kp, gh,
                                          gw, kSlab, d_gfeat_nhwc + (size_t)s * kSlab, ws + w.gemm_off, st, nullptr, cin, ksteps_last,
                                          1, 0, image_scales);
            if (rc) return rc;
        }
        return SQD_OK;
    }
    int rc = sqd_convdet_f16_pair(reinterpret_cast<const float *>(planes), SQD_LAYOUT_SPLIT_NHWC, d_dgrad_packed, nullptr, batch, kp,
                                  gh, gw, kSlab, d_gfeat_nhwc, ws + w.gemm_off, st, nullptr, cin, ksteps_last, ns, slab_bytes, image_scales);
    if (rc) return rc;
    return SQD_OK;
}

extern "C" int sqd_convdet_bias_grad(const float *d_gpred, int batch, int gh, int gw, int cout, float *d_gbias, void *stream) {
    SQD_REQUIRE(d_gpred && d_gbias, SQD_E_NULL, "sqd_convdet_bias_grad: NULL pointer");
    SQD_REQUIRE(batch >= 0 && gh > 0 && gw > 0 && cout >= 1, SQD_E_SHAPE, "sqd_convdet_bias_grad: bad shape");
    const size_t rows = (size_t)batch * gh * gw;
    if (cout % 4 == 0 && sqd_aligned16(d_gpred) && !sqd_opt(SQD_OPT_BWD_OLD_PREPASS)) {
        const int qpp = cout / 4, groups = (qpp + kBgQuads - 1) / kBgQuads;
        bias_grad_cluster_kernel<<<groups * kBgCluster, kBgThreads, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float4 *>(d_gpred), rows, qpp, d_gbias);
        SQD_LAUNCH_CHECK("bias_grad_cluster_kernel");
        return SQD_OK;
    }
    bias_grad_kernel<<<cout, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_gpred, rows, cout, d_gbias);
    SQD_LAUNCH_CHECK("bias_grad_kernel");
    return SQD_OK;
}

extern "C" size_t sqd_convdet_wgrad_workspace_bytes(int batch, int cin, int gh, int gw, int cout) {
    if (batch <= 0 || cin <= 0 || gh <= 0 || gw <= 0 || cout <= 0) return 256;
    return align256((size_t)wgrad_slices(batch, gh) * cout * cin * 9 * sizeof(float));
}

extern "C" int sqd_convdet_wgrad(const float *d_feat_nchw, const float *d_gpred, int batch, int cin, int gh, int gw, int cout,
                                 float *d_gweight, void *d_workspace, size_t workspace_bytes, void *stream) {
    SQD_REQUIRE(d_feat_nchw && d_gpred && d_gweight && d_workspace, SQD_E_NULL, "sqd_convdet_wgrad: NULL pointer");
    SQD_REQUIRE(batch >= 1 && cin >= 1 && gh > 0 && gw > 0 && cout >= 1 && cout <= 4096, SQD_E_SHAPE,
                "sqd_convdet_wgrad: bad shape (Cout in [1, 4096])");
    SQD_REQUIRE(sqd_aligned16(d_gpred), SQD_E_ALIGN, "sqd_convdet_wgrad: gpred must be 16-byte aligned");
    SQD_REQUIRE(workspace_bytes >= sqd_convdet_wgrad_workspace_bytes(batch, cin, gh, gw, cout), SQD_E_WORKSPACE,
                "sqd_convdet_wgrad: workspace too small (%zu bytes)", workspace_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ns = wgrad_slices(batch, gh);
    float *partial = static_cast<float *>(d_workspace);
    const dim3 grid(9, (cin + kWgC - 1) / kWgC, ns);
    for (int n0 = 0; n0 < cout; n0 += kWgN) {   // <= 80 output channels per launch (KITTI's 72: one launch)
        wgrad_partial_kernel<<<grid, kWgThreads, 0, st>>>(d_feat_nchw, d_gpred, batch, cin, gh, gw, cout, ns, partial, n0);
        SQD_LAUNCH_CHECK("wgrad_partial_kernel");
    }
    const size_t n = (size_t)cout * cin * 9;
    wgrad_reduce_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(partial, n, ns, d_gweight);
    SQD_LAUNCH_CHECK("wgrad_reduce_kernel");
    return SQD_OK;
}

// a1: ConvDet 3x3 head as a tcgen05 / TMEM implicit GEMM fed by TMA (sm_100a), 3xTF32.
// Reference: SqueezeDetBase.convdet + permute(0,2,3,1) + view, src/model/squeezedet.py:73-75,83-87
// (cuDNN conv with N=72 plus an NCHW->NHWC copy kernel there).
//
// GEMM view per image: M = gh*gw cells, N = Cout = K_anchors*(C+5) (72 KITTI, padded to 80), K = 9*Cin = 6912.
//
//  * M tile = 8 x 16 spatial block of cells = 128 rows = one UMMA_M.  For filter tap (dy,dx) and channel
//    block c0 the A operand is ONE 4-D TMA box {32 ch, 16 x, 8 y, 1 img} at (c0, x0+dx, y0+dy, b): the
//    conv padding is TMA's out-of-bounds zero fill, so there is no im2col buffer and no halo logic.
//    The box lands as 128 rows x 128 B, 128B-swizzled = the canonical K-major UMMA layout.
//  * B operand = packed weights [Npad][9*Cin] K-major (k = tap*Cin + c), a 2-D TMA box {32, Npad}.
//  * 3xTF32: features and weights are pre-split into tf32-exact hi and lo planes
//    (hi = rna_tf32(x), lo = rna_tf32(x - hi)); D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi.  One tf32 pass flips
//    top-k order; three passes recover fp32-level products.
//    Algorithmic FLOPs per image: 2*M*Cout*K (the 3 passes and N padding are NOT counted).
//  * Chunked accumulation: the tensor core TRUNCATES when it adds into the fp32 TMEM accumulator (measured on
//    B200: accumulating all 2592 MMAs in TMEM biases every output towards zero by ~2e-5 relative, 20x the
//    fp32 rounding noise).  So K is cut into chunks of `chunk_stages` pipeline stages; each chunk accumulates
//    from zero into one of two TMEM accumulators, and the epilogue warps add finished chunks into fp32
//    REGISTERS with round-to-nearest while the MMAs of the next chunk run into the other accumulator.
//  * Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected
//    lane), warps 2..5 = accumulate/epilogue (tcgen05.ld -> register sum -> +bias -> pred in the reference's
//    (B, A, C+5) layout).
//    smem ring of kStages x {A_hi, A_lo, B_hi, B_lo} with full/empty mbarriers; tcgen05.commit frees slots.
//  * Every mbarrier wait is bounded: on timeout the CTA sets a status word and drains instead of hanging.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace sqd_tc;

constexpr int kTileX = 16, kTileY = 8;        // cells per M tile (8 x 16 = 128 = UMMA_M)
constexpr int kBlockK = 32;                   // channels per pipeline stage (128 B of fp32 = one swizzle row)
constexpr int kUmmaK = 8;                     // tf32 MMA K
constexpr int kABytes = 128 * kBlockK * 4;    // 16 KiB per A plane per stage
constexpr int kThreadsTC = 192;

// ---------------------------------------------------------------------------------------------------
// pre-passes: hi/lo split of features (-> NHWC planes) and of the weights (-> [Npad][9*Cin] planes)
// ---------------------------------------------------------------------------------------------------
__global__ void split_nhwc_kernel(const float4 *__restrict__ in, float4 *__restrict__ hi, float4 *__restrict__ lo,
                                  size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = ld_stream_f4(in + i);
        float4 h, l;
        h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
        l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
        hi[i] = h;
        lo[i] = l;
    }
}

// NCHW (B,Cin,P) -> NHWC (B,P,Cin) planes through a 32x33 shared tile; P = gh*gw
__global__ void split_nchw_kernel(const float *__restrict__ in, float *__restrict__ hi, float *__restrict__ lo, int cin,
                                  int P) {
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
        if (p < P && c < cin) {
            const float v = tile[threadIdx.x][j];
            const float h = tf32_rna(v);
            const size_t o = ((size_t)b * P + p) * cin + c;
            hi[o] = h;
            lo[o] = tf32_rna(v - h);
        }
    }
}

__global__ void pack_weights_tc_kernel(const float *__restrict__ w, int cout, int cin, int npad, float *__restrict__ hi,
                                       float *__restrict__ lo) {
    const size_t ktot = (size_t)9 * cin, total = (size_t)npad * ktot;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / ktot);
        const size_t k = i % ktot;
        const int tap = (int)(k / cin), c = (int)(k % cin);
        float v = 0.f;
        if (n < cout) v = w[((size_t)n * cin + c) * 9 + tap];
        const float h = tf32_rna(v);
        hi[i] = h;
        lo[i] = tf32_rna(v - h);
    }
}

// ---------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------
struct TcParams {
    int cin, gh, gw, cout;
    int tiles_x;
    int num_stages;
    int chunk_stages;  // pipeline stages accumulated inside TMEM before the sum moves to registers
    const float *bias;
    float *pred;
    int *status;  // 0 ok; else the role whose bounded wait timed out (1 TMA, 2 MMA/full, 3 accumulate, 4 MMA/tmem)
};

template <int NPAD>
__global__ void __launch_bounds__(kThreadsTC, 1)
convdet_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                  const TcParams p) {
    constexpr int kBBytes = NPAD * kBlockK * 4;
    constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
    constexpr uint32_t kAccStride = NPAD <= 16 ? 16 : (NPAD <= 32 ? 32 : (NPAD <= 64 ? 64 : 128));  // columns per accumulator
    constexpr uint32_t kTmemCols = 2 * kAccStride;                                                   // two accumulators
    constexpr uint32_t kIdesc = umma_idesc_tf32(128, NPAD);

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int S = p.num_stages;
    uint8_t *ctrl = smem + (size_t)S * kStageBytes;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(ctrl);
    uint64_t *empty_bar = full_bar + 8;
    uint64_t *tmem_full_bar = empty_bar + 8;   // [2]
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;  // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);
    volatile int *abort_flag = reinterpret_cast<volatile int *>(tmem_slot + 1);
    float *s_bias = reinterpret_cast<float *>(tmem_slot + 2);  // NPAD floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, img = blockIdx.y;
    const int x0 = (tile % p.tiles_x) * kTileX, y0 = (tile / p.tiles_x) * kTileY;
    const int cblocks = p.cin / kBlockK;
    const int iters = cblocks * 9;
    const int chunk = p.chunk_stages;
    const int chunks = (iters + chunk - 1) / chunk;

    if (threadIdx.x == 0) {
        *abort_flag = 0;
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tmem_full_bar + b, 1);
            mbar_init(tmem_empty_bar + b, 4);  // one arrival per accumulate warp
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
    }
    for (int i = threadIdx.x; i < NPAD; i += kThreadsTC) s_bias[i] = i < p.cout ? __ldg(p.bias + i) : 0.f;
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                if (!mbar_wait(empty_bar + s, ph ^ 1u, abort_flag)) {
                    atomicCAS(p.status, 0, 1);
                    break;
                }
                const int cb = it / 9, tap = it - cb * 9;   // channel block outer, tap inner: the 9 shifted
                const int dy = tap / 3 - 1, dx = tap % 3 - 1;  // windows of one channel block hit L2 back to back
                uint8_t *st = smem + (size_t)s * kStageBytes;
                mbar_arrive_expect_tx(full_bar + s, kStageBytes);
                tma_load_4d(&map_a_hi, full_bar + s, st, cb * kBlockK, x0 + dx, y0 + dy, img);
                tma_load_4d(&map_a_lo, full_bar + s, st + kABytes, cb * kBlockK, x0 + dx, y0 + dy, img);
                tma_load_2d(&map_b_hi, full_bar + s, st + 2 * kABytes, tap * p.cin + cb * kBlockK, 0);
                tma_load_2d(&map_b_lo, full_bar + s, st + 2 * kABytes + kBBytes, tap * p.cin + cb * kBlockK, 0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single thread) =====
        if (lane == 0) {
            bool ok = true;
            int it = 0;
            for (int c = 0; c < chunks && ok; ++c) {
                const int buf = c & 1;
                const uint32_t acc_ph = (uint32_t)(c >> 1) & 1u;
                if (!mbar_wait(tmem_empty_bar + buf, acc_ph ^ 1u, abort_flag)) {  // accumulator drained by the warps
                    atomicCAS(p.status, 0, 4);
                    ok = false;
                    break;
                }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * kAccStride;
                for (int j = 0; j < chunk && it < iters; ++j, ++it) {
                    const int s = it % S;
                    const uint32_t ph = (uint32_t)(it / S) & 1u;
                    if (!mbar_wait(full_bar + s, ph, abort_flag)) {
                        atomicCAS(p.status, 0, 2);
                        ok = false;
                        break;
                    }
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)s * kStageBytes);
                    const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + kABytes);
                    const uint64_t b_hi = umma_desc_sw128(sa + 2 * kABytes),
                                   b_lo = umma_desc_sw128(sa + 2 * kABytes + kBBytes);
#pragma unroll
                    for (int ks = 0; ks < kBlockK / kUmmaK; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * kUmmaK * 4) >> 4);  // +32 B per K step, in 16 B units
                        // small cross terms first, then the dominant hi*hi product; each chunk starts from zero
                        umma_tf32(d_tmem, a_lo + adv, b_hi + adv, kIdesc, (j | ks) ? 1u : 0u);
                        umma_tf32(d_tmem, a_hi + adv, b_lo + adv, kIdesc, 1u);
                        umma_tf32(d_tmem, a_hi + adv, b_hi + adv, kIdesc, 1u);
                    }
                    umma_commit(empty_bar + s);  // slot reusable once these MMAs have read it
                }
                umma_commit(tmem_full_bar + buf);  // chunk complete (also fires after an aborted loop)
            }
        }
    } else {
        // ===== accumulate + epilogue warps: TMEM chunk -> fp32 registers (RN) ... -> (+bias) -> pred =====
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // accumulator row == cell inside the 8x16 tile
        const int y = y0 + row / kTileX, x = x0 + row % kTileX;
        float acc[NPAD];
#pragma unroll
        for (int n = 0; n < NPAD; ++n) acc[n] = 0.f;
        bool ok = true;
        for (int c = 0; c < chunks; ++c) {
            const int buf = c & 1;
            const uint32_t acc_ph = (uint32_t)(c >> 1) & 1u;
            if (!mbar_wait(tmem_full_bar + buf, acc_ph, abort_flag)) {
                if (lane == 0) atomicCAS(p.status, 0, 3);
                ok = false;
                break;
            }
            tc_fence_after();
            __syncwarp();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * kAccStride;
#pragma unroll
            for (int n0 = 0; n0 < NPAD; n0 += 16) {
                uint32_t v[16];
                tmem_ld_x16(taddr + n0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[n0 + i] += __uint_as_float(v[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar + buf);  // this warp is done reading the accumulator
        }
        __syncwarp();
        if (ok && y < p.gh && x < p.gw) {
            float *out = p.pred + (((size_t)img * p.gh + y) * p.gw + x) * p.cout;
            if ((p.cout & 3) == 0) {
                float4 *o4 = reinterpret_cast<float4 *>(out);
#pragma unroll
                for (int n = 0; n < NPAD; n += 4)
                    if (n < p.cout)
                        o4[n >> 2] = make_float4(acc[n] + s_bias[n], acc[n + 1] + s_bias[n + 1], acc[n + 2] + s_bias[n + 2],
                                                 acc[n + 3] + s_bias[n + 3]);
            } else {
#pragma unroll
                for (int n = 0; n < NPAD; ++n)
                    if (n < p.cout) out[n] = acc[n] + s_bias[n];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
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

int stages_for(int npad) {
    const int stage = 2 * kABytes + 2 * npad * kBlockK * 4;
    int s = (227 * 1024 - 2048 - 1024) / stage;
    if (s > 4) s = 4;
    return s;
}

size_t smem_bytes_for(int npad, int stages) {
    return (size_t)stages * (2 * kABytes + 2 * npad * kBlockK * 4) + 1024 /*align*/ + 256 /*barriers*/ + npad * 4;
}

}  // namespace

size_t sqd_tc_packed_bytes(int cout, int cin) { return (size_t)2 * npad_of(cout) * 9 * cin * sizeof(float); }

// workspace = [status word, padded to 256 B | A_hi plane | A_lo plane]; pre-split input needs only the status word
size_t sqd_tc_workspace_bytes(int batch, int cin, int gh, int gw, int layout) {
    if (layout == SQD_LAYOUT_SPLIT_NHWC) return 256;
    return (size_t)2 * batch * gh * gw * cin * sizeof(float) + 256;
}

// stand-alone hi/lo split (the first half of sqd_convdet_tc), for callers that keep the planes
int sqd_tc_split_features(const float *d_feat, int layout, int batch, int cin, int gh, int gw, float *d_planes,
                          cudaStream_t st) {
    const size_t plane = (size_t)batch * gh * gw * cin;
    float *a_hi = d_planes, *a_lo = d_planes + plane;
    if (layout == SQD_LAYOUT_NHWC) {
        split_nhwc_kernel<<<4 * SQD_SM_COUNT, 256, 0, st>>>(reinterpret_cast<const float4 *>(d_feat),
                                                            reinterpret_cast<float4 *>(a_hi),
                                                            reinterpret_cast<float4 *>(a_lo), plane / 4);
    } else {
        const int P = gh * gw;
        dim3 grid((P + 31) / 32, (cin + 31) / 32, batch);
        split_nchw_kernel<<<grid, dim3(32, 8), 0, st>>>(d_feat, a_hi, a_lo, cin, P);
    }
    SQD_LAUNCH_CHECK("split kernel");
    return SQD_OK;
}

int sqd_tc_pack_weights(const float *d_weight, int cout, int cin, void *d_packed, cudaStream_t st) {
    const int npad = npad_of(cout);
    float *hi = static_cast<float *>(d_packed);
    float *lo = hi + (size_t)npad * 9 * cin;
    pack_weights_tc_kernel<<<2 * SQD_SM_COUNT, 256, 0, st>>>(d_weight, cout, cin, npad, hi, lo);
    SQD_LAUNCH_CHECK("pack_weights_tc_kernel");
    return SQD_OK;
}

template <int NPAD>
static int launch_tc(const CUtensorMap *maps, const TcParams &p, int tiles, int batch, cudaStream_t st) {
    const size_t smem = smem_bytes_for(NPAD, p.num_stages);
    SQD_CUDA(cudaFuncSetAttribute(convdet_tc_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    convdet_tc_kernel<NPAD><<<dim3(tiles, batch), kThreadsTC, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
    SQD_LAUNCH_CHECK("convdet_tc_kernel");
    return SQD_OK;
}

// d_workspace: [status word (256 B) | A_hi plane | A_lo plane]
int sqd_convdet_tc(const float *d_feat, int layout, const void *d_packed, const float *d_bias, int batch, int cin,
                   int gh, int gw, int cout, float *d_pred, void *d_workspace, cudaStream_t st) {
    SQD_REQUIRE(cin % kBlockK == 0, SQD_E_SHAPE, "convdet (tcgen05): Cin %d must be a multiple of %d", cin, kBlockK);
    SQD_REQUIRE(cout >= 1 && cout <= 128, SQD_E_SHAPE, "convdet (tcgen05): Cout %d outside [1,128]", cout);
    SQD_REQUIRE(batch <= 65535, SQD_E_SHAPE, "convdet (tcgen05): batch %d > 65535 (split the call)", batch);
    EncodeTiledFn encode = get_encode_fn();
    SQD_REQUIRE(encode != nullptr, SQD_E_DRIVER, "cuTensorMapEncodeTiled not available from the CUDA driver");
    const int npad = npad_of(cout);
    const size_t plane = (size_t)batch * gh * gw * cin;
    int *status = static_cast<int *>(d_workspace);
    SQD_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
    float *a_hi, *a_lo;
    if (layout == SQD_LAYOUT_SPLIT_NHWC) {
        // caller already holds the tf32 hi/lo NHWC planes (sqd_convdet_split_features)
        a_hi = const_cast<float *>(d_feat);
        a_lo = a_hi + plane;
    } else {
        // 1. hi/lo split into NHWC planes
        a_hi = reinterpret_cast<float *>(static_cast<char *>(d_workspace) + 256);
        a_lo = a_hi + plane;
        int rc = sqd_tc_split_features(d_feat, layout, batch, cin, gh, gw, a_hi, st);
        if (rc) return rc;
    }

    // 2. tensor maps
    alignas(64) CUtensorMap maps[4];
    {
        const cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)gw, (cuuint64_t)gh, (cuuint64_t)batch};
        const cuuint64_t strides[3] = {(cuuint64_t)cin * 4, (cuuint64_t)gw * cin * 4, (cuuint64_t)gh * gw * cin * 4};
        const cuuint32_t box[4] = {kBlockK, kTileX, kTileY, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        for (int i = 0; i < 2; ++i) {
            CUresult r = encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, i == 0 ? (void *)a_hi : (void *)a_lo, dims,
                                strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            SQD_REQUIRE(r == CUDA_SUCCESS, SQD_E_DRIVER, "cuTensorMapEncodeTiled(features) failed: CUresult %d", (int)r);
        }
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
            CUresult r = encode(&maps[2 + i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, i == 0 ? (void *)b_hi : (void *)b_lo,
                                dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            SQD_REQUIRE(r == CUDA_SUCCESS, SQD_E_DRIVER, "cuTensorMapEncodeTiled(weights) failed: CUresult %d", (int)r);
        }
    }

    // 3. the GEMM
    TcParams p;
    p.cin = cin; p.gh = gh; p.gw = gw; p.cout = cout;
    p.tiles_x = (gw + kTileX - 1) / kTileX;
    p.num_stages = stages_for(npad);
    p.chunk_stages = 4;  // 48 MMAs per TMEM accumulation epoch: accuracy at fp32 level, TMEM reads fully overlapped
    if (const char *e = getenv("SQD_TC_CHUNK")) {  // tuning / experiment knob
        const int v = atoi(e);
        if (v >= 1) p.chunk_stages = v;
    }
    p.bias = d_bias;
    p.pred = d_pred;
    p.status = status;
    const int tiles = p.tiles_x * ((gh + kTileY - 1) / kTileY);
    switch (npad / 16) {
        case 1: return launch_tc<16>(maps, p, tiles, batch, st);
        case 2: return launch_tc<32>(maps, p, tiles, batch, st);
        case 3: return launch_tc<48>(maps, p, tiles, batch, st);
        case 4: return launch_tc<64>(maps, p, tiles, batch, st);
        case 5: return launch_tc<80>(maps, p, tiles, batch, st);
        case 6: return launch_tc<96>(maps, p, tiles, batch, st);
        case 7: return launch_tc<112>(maps, p, tiles, batch, st);
        case 8: return launch_tc<128>(maps, p, tiles, batch, st);
    }
    SQD_REQUIRE(false, SQD_E_SHAPE, "convdet (tcgen05): unsupported Cout %d", cout);
}

// status word is the first word of the workspace (see sqd_convdet_tc)
const int *sqd_tc_status_ptr(const void *d_workspace) { return static_cast<const int *>(d_workspace); }

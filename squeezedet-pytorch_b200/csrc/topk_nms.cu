// a8-a9: Detector.filter for a whole batch -- exact top-k, per-class greedy NMS, score threshold --
// and its fusion with the decode (sqd_detect_from_pred).
// Reference: src/engine/detector.py:87-122 (full torch.argsort of A scores, 3 torchvision.ops.nms
// calls and >= 3C+3 host syncs PER IMAGE) and torchvision's CPU nms kernel for the IoU arithmetic.
//
// One thread-block CLUSTER of 1/2/4/8 CTAs per image (small batches would otherwise leave most SMs idle: one CTA
// per image is latency bound at batch 20).  Each CTA scans its slice of the anchors and keeps its local top-k;
// rank 0 then pulls the other ranks' survivors through distributed shared memory, selects the top-k of the <= 2048
// candidates, sorts them and runs phases 2-3.  The union of local top-k lists contains the global top-k, and the 64-bit keys
// are totally ordered, so the result does not depend on the cluster size.
//  1. Scan: every thread scores its anchors; a candidate is a 64-bit key
//        [ order-preserving score bits : 32 | 0xFFFFFF - anchor : 24 | class : 8 ]
//     so "larger key" == (score desc, anchor index asc) -- the declared tie policy.
//     Candidates above the running k-th-best threshold are appended to a 2048-entry shared buffer;
//     when a round could overflow it, the exact top-k of the buffer is SELECTED (MSB-first radix select on the
//     keys, no sort) and the threshold raised.  After the first cut almost nothing passes (expected
//     k*ln(A/2048) more candidates).  Only the final k survivors are sorted (by ranking).
//  2. The k survivors (sorted) get their boxes; a k x k same-class IoU bitmask is built with one
//     ballot per 32 pairs; one warp runs the sequential greedy sweep over the mask rows.
//  3. Kept rows with score > thresh are emitted class-ascending / score-descending.
// The running threshold starts at the score threshold (exact, see score_floor_key), so on real inputs only a few
// hundred anchors per image ever enter the candidate buffer and mid-scan compactions do not happen.
// Algorithmic HBM bytes per image: A*(C+5)*4 (fused) or A*4 (+ a few KB of gathers) for the dense form.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 512;
constexpr int kUnroll = 2;
constexpr int kRound = kThreads * kUnroll;  // anchors consumed per round
constexpr int kCap = 2048;                  // candidate buffer entries (>= SQD_MAX_TOPK + kRound)
static_assert(kCap >= SQD_MAX_TOPK + kRound, "candidate buffer too small");
static_assert(kCap / 2 >= SQD_MAX_TOPK, "rank_sort uses the upper half of the buffer as its destination");

typedef unsigned long long u64;

__device__ __forceinline__ unsigned order_bits(float s) {
    const unsigned b = __float_as_uint(s);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unorder_bits(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ u64 make_key(float score, int anchor, int cls) {
    return ((u64)order_bits(score) << 32) | ((u64)(0xFFFFFFu - (unsigned)anchor) << 8) | (u64)(cls & 0xFF);
}
__device__ __forceinline__ int key_anchor(u64 k) { return (int)(0xFFFFFFu - (unsigned)((k >> 8) & 0xFFFFFFu)); }
__device__ __forceinline__ int key_class(u64 k) { return (int)(k & 0xFFu); }
__device__ __forceinline__ float key_score(u64 k) { return unorder_bits((unsigned)(k >> 32)); }

// Exact pre-filter: an anchor with score <= score_thresh can never be emitted (final strict filter) and can never
// suppress an emitted box (greedy NMS only lets HIGHER scores suppress), so it only ever occupies a top-k slot that
// no surviving anchor needs.  Starting the running threshold at "largest key with score == score_thresh" keeps
// such anchors out of the candidate buffer altogether; the result is identical to the reference's
// top-k -> NMS -> score filter order (detector.py:88-114).
__device__ __forceinline__ u64 score_floor_key(float score_thr) {
    return ((u64)order_bits(score_thr) << 32) | 0xFFFFFFFFull;
}

// ---- candidate sources ---------------------------------------------------------------------------
template <int CS>
struct FromPred {  // fused: score and box straight from the ConvDet output
    const float *pred;  // this image, (A, C+5)
    const float4 *anchors;
    int C;
    float wmax, hmax;
    __device__ __forceinline__ u64 key(int a) const {
        const int Cn = CS > 0 ? CS : C;
        const int NF = Cn + 5;
        float f[SQD_CMAX(CS) + 1];
        const float *row = pred + (size_t)a * NF;
        if (CS == 3) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(row));
            f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < SQD_CMAX(CS) + 1; ++j)
                if (j <= Cn) f[j] = __ldg(row + j);
        }
        float s;
        int c;
        sqd_score_anchor<CS>(f, Cn, s, c);
        return make_key(s, a, c);
    }
    __device__ __forceinline__ float4 box(int a) const {
        const int Cn = CS > 0 ? CS : C;
        const float *row = pred + (size_t)a * (Cn + 5) + Cn + 1;
        return sqd_decode_box(__ldg(anchors + a), __ldg(row), __ldg(row + 1), __ldg(row + 2), __ldg(row + 3), wmax,
                              hmax);
    }
};

struct FromDense {  // Detector.filter's own contract: dense ids / scores / boxes
    const long long *class_ids;
    const float *scores;
    const float4 *boxes;
    __device__ __forceinline__ float score(int a) const { return __ldg(scores + a); }
    __device__ __forceinline__ u64 key_of(float s, int a) const { return make_key(s, a, (int)__ldg(class_ids + a)); }
    __device__ __forceinline__ float4 box(int a) const { return __ldg(boxes + a); }
};

struct Shared {
    u64 buf[kCap];
    int hist[256];
    int count;
    int n_valid;
    int sel_bin, sel_need, sel_done;
    u64 thresh;
};

// ---- exact top-k SELECTION (no sort) ---------------------------------------------------------------
// MSB-first radix select over the 64-bit keys with 8-bit digits: every pass histograms the digit of the keys that
// still share the prefix of the k-th largest key and narrows the prefix by one digit; it stops as soon as the
// selected bin is needed in full.  Keys are unique (the anchor index is part of the key), so exactly k keys
// satisfy key >= sel_lo.  Keys live in registers during the passes (kCap / kThreads per thread); the survivors
// are written back unsorted, and the running threshold becomes sel_lo - 1.  All threads must call.
__device__ void select_topk(Shared &sh, int k) {
    constexpr int kPer = kCap / kThreads;
    __syncthreads();
    const int cnt = min(sh.count, kCap);
    if (cnt <= k) {           // nothing to cut (uniform)
        if (threadIdx.x == 0) sh.count = cnt;
        __syncthreads();
        return;
    }
    u64 my[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const int idx = threadIdx.x + i * kThreads;
        my[i] = idx < cnt ? sh.buf[idx] : 0ull;  // 0 is below every real key (score bits of a finite float are never 0)
    }
    const int lane = threadIdx.x & 31;
    u64 prefix = 0ull;
    int need = k, shift = 56;
    for (; shift >= 0; shift -= 8) {
        if (threadIdx.x < 256) sh.hist[threadIdx.x] = 0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const bool in = my[i] != 0ull && (shift == 56 || (my[i] >> (shift + 8)) == (prefix >> (shift + 8)));
            const unsigned digit = (unsigned)(my[i] >> shift) & 255u;
            // warp-aggregated histogram update: one atomic per distinct digit per warp
            const unsigned act = __ballot_sync(0xffffffffu, in);
            if (in) {
                const unsigned peers = __match_any_sync(act, digit);
                if (lane == __ffs(peers) - 1) atomicAdd(&sh.hist[digit], __popc(peers));
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // lane l owns bins [8l, 8l+8); find the bin b with  #(digit > b) < need <= #(digit >= b)
            int c[8], own = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                c[j] = sh.hist[lane * 8 + j];
                own += c[j];
            }
            int above = own;  // inclusive suffix sum over lanes >= this one
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_down_sync(0xffffffffu, above, o);
                if (lane + o < 32) above += v;
            }
            const int higher = above - own;  // keys in bins of higher lanes
            if (higher < need && need <= above) {
                int acc = higher;
                for (int j = 7; j >= 0; --j) {
                    if (acc + c[j] >= need) {
                        sh.sel_bin = lane * 8 + j;
                        sh.sel_need = need - acc;
                        sh.sel_done = (c[j] == need - acc) ? 1 : 0;
                        break;
                    }
                    acc += c[j];
                }
            }
        }
        __syncthreads();
        prefix |= (u64)(unsigned)sh.sel_bin << shift;
        need = sh.sel_need;
        if (sh.sel_done) break;  // the whole bin is needed: no further digits to resolve (uniform)
    }
    if (shift < 0) shift = 0;
    const u64 sel_lo = prefix;  // lower (unresolved) digits are zero: selected <=> key >= sel_lo
    __syncthreads();            // everyone has read sel_*; buf may be overwritten now
    if (threadIdx.x == 0) {
        sh.count = 0;
        sh.thresh = sel_lo - 1ull;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPer; ++i)
        if (my[i] >= sel_lo && my[i] != 0ull) sh.buf[atomicAdd(&sh.count, 1)] = my[i];
    __syncthreads();
}

// Sort the (<= k <= kCap/2) survivors in descending order by ranking: rank(i) = #{j : key_j > key_i}.
// k*k/kThreads comparisons per thread (8 for k = 64); keys are unique so ranks are a permutation.
__device__ void rank_sort(Shared &sh) {
    const int m = sh.count;
    u64 *out = sh.buf + kCap / 2;
    for (int i = threadIdx.x; i < m; i += kThreads) {
        const u64 key = sh.buf[i];
        int rank = 0;
        for (int j = 0; j < m; ++j) rank += sh.buf[j] > key ? 1 : 0;
        out[rank] = key;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += kThreads) sh.buf[i] = out[i];
    __syncthreads();
}

// Cluster merge: every rank holds its (unsorted) local top-k in sh.buf[0..sh.count).  Rank 0 appends the other
// ranks' lists (read through DSMEM) to its own and cuts the union to k again.  All threads of all CTAs must call.
__device__ void cluster_merge(Shared &sh, int k, int cs) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();  // local lists complete and visible cluster-wide
    if (cluster.block_rank() == 0) {
        int total = sh.count;
        __syncthreads();
        for (int r = 1; r < cs; ++r) {
            const Shared *rs = cluster.map_shared_rank(&sh, r);
            const int cnt = min(rs->count, k);
            for (int i = threadIdx.x; i < cnt; i += kThreads) sh.buf[total + i] = rs->buf[i];
            total += cnt;
        }
        __syncthreads();
        if (threadIdx.x == 0) sh.count = total;
        select_topk(sh, k);
    }
    cluster.sync();  // remote lists no longer needed: the other ranks may exit
}

struct FilterOut {
    int *count;
    int *anchor;
    int *cls;
    float *score;
    float4 *box;
};

// Phase 2+3, shared by both sources.  sh.buf[0..m) holds the sorted survivors.
template <class Src>
__device__ void nms_and_emit(const Src &src, Shared &sh, unsigned char *dyn, int k, int num_classes, float nms_thr_f,
                             float score_thr_f, const FilterOut &o, int img) {
    const int m = sh.count;
    const int wpr = (k + 31) >> 5;  // mask words per row
    float4 *sbox = reinterpret_cast<float4 *>(dyn);
    float *sarea = reinterpret_cast<float *>(sbox + k);
    unsigned *mask = reinterpret_cast<unsigned *>(sarea + k);
    unsigned char *valid = reinterpret_cast<unsigned char *>(mask + (size_t)k * wpr);

    for (int i = threadIdx.x; i < m; i += kThreads) {
        const float4 b = src.box(key_anchor(sh.buf[i]));
        sbox[i] = b;
        sarea[i] = fmul(fsub(b.z, b.x), fsub(b.w, b.y));  // torchvision: (x2-x1)*(y2-y1), no +1
        valid[i] = 0;
    }
    __syncthreads();

    // same-class suppression mask: bit j of row i set iff j>i, class equal and IoU(i,j) > thresh
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = kThreads >> 5;
    for (int item = warp; item < m * wpr; item += nwarp) {
        const int i = item / wpr, w = item - i * wpr;
        const int j = (w << 5) + lane;
        bool sup = false;
        if (j > i && j < m && key_class(sh.buf[j]) == key_class(sh.buf[i])) {
            const float4 a = sbox[i], b = sbox[j];
            const float iw = fmaxf(0.f, fsub(fminf(a.z, b.z), fmaxf(a.x, b.x)));
            const float ih = fmaxf(0.f, fsub(fminf(a.w, b.w), fmaxf(a.y, b.y)));
            const float inter = fmul(iw, ih);
            const float iou = fdiv(inter, fsub(fadd(sarea[i], sarea[j]), inter));
            sup = iou > nms_thr_f;  // false for NaN (0/0 of zero-area boxes), like the reference
        }
        const unsigned bits = __ballot_sync(0xffffffffu, sup);
        if (lane == 0) mask[item] = bits;
    }
    __syncthreads();

    // greedy sweep in descending-score order (one warp; lane l owns word l of the removed set)
    if (warp == 0) {
        unsigned removed = 0u;  // word `lane`
        for (int i = 0; i < m; ++i) {
            const unsigned word = __shfl_sync(0xffffffffu, removed, i >> 5);
            if (!((word >> (i & 31)) & 1u)) {
                if (lane < wpr) removed |= mask[i * wpr + lane];
                if (lane == 0) valid[i] = key_score(sh.buf[i]) > score_thr_f ? 1 : 0;
            }
        }
    }
    __syncthreads();

    // emit: class ascending, then descending score (== position) inside a class
    int n_valid_local = 0;
    for (int t = threadIdx.x; t < m; t += kThreads) {
        if (!valid[t]) continue;
        const int ct = key_class(sh.buf[t]);
        int pos = 0;
        for (int u = 0; u < m; ++u) {
            if (!valid[u]) continue;
            const int cu = key_class(sh.buf[u]);
            pos += (cu < ct || (cu == ct && u < t)) ? 1 : 0;
        }
        const size_t r = (size_t)img * k + pos;
        o.anchor[r] = key_anchor(sh.buf[t]);
        o.cls[r] = ct;
        o.score[r] = key_score(sh.buf[t]);
        o.box[r] = sbox[t];
        ++n_valid_local;
    }
    if (threadIdx.x == 0) sh.n_valid = 0;
    __syncthreads();
    if (n_valid_local) atomicAdd(&sh.n_valid, n_valid_local);
    __syncthreads();
    const int nv = sh.n_valid;
    if (threadIdx.x == 0) o.count[img] = nv;
    for (int t = nv + threadIdx.x; t < k; t += kThreads) {  // deterministic padding rows
        const size_t r = (size_t)img * k + t;
        o.anchor[r] = -1;
        o.cls[r] = -1;
        o.score[r] = 0.f;
        o.box[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    (void)num_classes;
}

template <int CS>
__global__ void __launch_bounds__(kThreads) detect_from_pred_kernel(const float *pred, const float4 *anchors, int A,
                                                                    int C, float wmax, float hmax, int k,
                                                                    float nms_thr_f, float score_thr_f, FilterOut o,
                                                                    int cs) {
    __shared__ Shared sh;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int img = blockIdx.x / cs, rank = blockIdx.x - img * cs;
    const int chunk = (A + cs - 1) / cs;
    const int a_begin = rank * chunk, a_end = min(A, a_begin + chunk);
    FromPred<CS> src;
    src.pred = pred + (size_t)img * A * ((CS > 0 ? CS : C) + 5);
    src.anchors = anchors;
    src.C = C;
    src.wmax = wmax;
    src.hmax = hmax;
    if (threadIdx.x == 0) {
        sh.count = 0;
        sh.thresh = score_floor_key(score_thr_f);
    }
    __syncthreads();
    for (int base = a_begin; base < a_end; base += kRound) {
        const u64 thr = sh.thresh;
        u64 keys[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int a = base + u * kThreads + threadIdx.x;
            keys[u] = a < a_end ? src.key(a) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (keys[u] > thr) {
                const int pos = atomicAdd(&sh.count, 1);
                if (pos < kCap) sh.buf[pos] = keys[u];
            }
        // Barrier + vote in one: the thread that performs the round's last append observes the final
        // count, so the OR is true for everyone iff the buffer could overflow next round (uniform branch).
        if (__syncthreads_or(*(volatile int *)&sh.count > kCap - kRound)) select_topk(sh, k);
    }
    select_topk(sh, k);
    if (cs > 1) {
        cluster_merge(sh, k, cs);
        if (rank != 0) return;
    }
    rank_sort(sh);
    nms_and_emit(src, sh, dyn, k, C, nms_thr_f, score_thr_f, o, img);
}

__global__ void __launch_bounds__(kThreads) filter_dense_kernel(const long long *class_ids, const float *scores,
                                                               const float4 *boxes, int A, int C, int k,
                                                               float nms_thr_f, float score_thr_f, FilterOut o, int cs) {
    __shared__ Shared sh;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int img = blockIdx.x / cs, rank = blockIdx.x - img * cs;
    const int chunk = (A + cs - 1) / cs;
    const int a_begin = rank * chunk, a_end = min(A, a_begin + chunk);
    FromDense src;
    src.class_ids = class_ids + (size_t)img * A;
    src.scores = scores + (size_t)img * A;
    src.boxes = boxes + (size_t)img * A;
    if (threadIdx.x == 0) {
        sh.count = 0;
        sh.thresh = score_floor_key(score_thr_f);
    }
    __syncthreads();
    for (int base = a_begin; base < a_end; base += kRound) {
        const unsigned thr_hi = (unsigned)(sh.thresh >> 32);
        float s[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int a = base + u * kThreads + threadIdx.x;
            s[u] = a < a_end ? src.score(a) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int a = base + u * kThreads + threadIdx.x;
            if (a < a_end && order_bits(s[u]) >= thr_hi) {  // cheap pre-test on the score word only
                const u64 key = src.key_of(s[u], a);
                if (key > sh.thresh) {
                    const int pos = atomicAdd(&sh.count, 1);
                    if (pos < kCap) sh.buf[pos] = key;
                }
            }
        }
        if (__syncthreads_or(*(volatile int *)&sh.count > kCap - kRound)) select_topk(sh, k);
    }
    select_topk(sh, k);
    if (cs > 1) {
        cluster_merge(sh, k, cs);
        if (rank != 0) return;
    }
    rank_sort(sh);
    nms_and_emit(src, sh, dyn, k, C, nms_thr_f, score_thr_f, o, img);
}

size_t dyn_smem_bytes(int k) {
    const size_t wpr = (k + 31) / 32;
    return (size_t)k * 16 + (size_t)k * 4 + (size_t)k * wpr * 4 + (size_t)k + 16;
}

float float_at_or_below(double t) {  // largest float <= t: (double)iou > t  <=>  iou > this
    float f = (float)t;
    if ((double)f > t) f = nextafterf(f, -INFINITY);
    return f;
}

int check_common(const char *fn, int batch, int A, int C, int k, const void *count, const void *anchor,
                 const void *cls, const void *score, const void *box) {
    SQD_REQUIRE(count && anchor && cls && score && box, SQD_E_NULL, "%s: an output pointer is NULL", fn);
    SQD_REQUIRE(batch >= 0 && A > 0 && A <= 0xFFFFFF, SQD_E_SHAPE, "%s: num_anchors %d outside (0, 2^24)", fn, A);
    SQD_REQUIRE(C >= 1 && C <= SQD_MAX_CLASSES, SQD_E_SHAPE, "%s: num_classes %d outside [1,%d]", fn, C,
                SQD_MAX_CLASSES);
    SQD_REQUIRE(k >= 1 && k <= SQD_MAX_TOPK, SQD_E_SHAPE, "%s: top_k %d outside [1,%d]", fn, k, SQD_MAX_TOPK);
    SQD_REQUIRE(sqd_aligned16(box), SQD_E_ALIGN, "%s: out_box must be 16-byte aligned", fn);
    return SQD_OK;
}

template <class K>
int opt_in_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) SQD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return SQD_OK;
}

// CTAs per image: as many as keep the merged candidate list inside the shared buffer (k*cs <= kCap) and are
// useful for filling the GPU at this batch size (at most ~4 CTAs per SM in flight); at least 2 rounds per CTA.
int cluster_size_for(int batch, int A, int k) {
    int cs = 8;
    while (cs > 1 && ((long long)k * cs > kCap || (long long)batch * cs > 4ll * SQD_SM_COUNT || A / cs < 2 * kRound)) cs >>= 1;
    return cs;
}

template <class K, class... Args>
int launch_clustered(const char *name, K kernel, int batch, int cs, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(batch * cs));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    if (e != cudaSuccess) {
        sqd_set_error("launch of %s failed: %s", name, cudaGetErrorString(e));
        return (int)e;
    }
    return SQD_OK;
}

}  // namespace

extern "C" int sqd_topk_nms(const int64_t *d_class_ids, const float *d_scores, const float *d_boxes, int batch,
                            int num_anchors, int num_classes, int top_k, double nms_thresh, double score_thresh,
                            int32_t *d_count, int32_t *d_out_anchor, int32_t *d_out_class, float *d_out_score,
                            float *d_out_box, void *stream) {
    if (batch == 0) return SQD_OK;  // empty batch: nothing to enqueue, pointers may be NULL
    SQD_REQUIRE(d_class_ids && d_scores && d_boxes, SQD_E_NULL, "sqd_topk_nms: an input pointer is NULL");
    int rc = check_common("sqd_topk_nms", batch, num_anchors, num_classes, top_k, d_count, d_out_anchor, d_out_class,
                          d_out_score, d_out_box);
    if (rc) return rc;
    SQD_REQUIRE(sqd_aligned16(d_boxes), SQD_E_ALIGN, "sqd_topk_nms: boxes must be 16-byte aligned");
    FilterOut o{d_count, d_out_anchor, d_out_class, d_out_score, reinterpret_cast<float4 *>(d_out_box)};
    const size_t smem = dyn_smem_bytes(top_k);
    rc = opt_in_smem(filter_dense_kernel, smem);
    if (rc) return rc;
    return launch_clustered("filter_dense_kernel", filter_dense_kernel, batch, cluster_size_for(batch, num_anchors, top_k),
                            smem, static_cast<cudaStream_t>(stream), reinterpret_cast<const long long *>(d_class_ids),
                            d_scores, reinterpret_cast<const float4 *>(d_boxes), num_anchors, num_classes, top_k,
                            float_at_or_below(nms_thresh), (float)score_thresh, o, cluster_size_for(batch, num_anchors, top_k));
}

extern "C" int sqd_detect_from_pred(const float *d_pred, const float *d_anchors, int batch, int num_anchors,
                                    int num_classes, int input_h, int input_w, int top_k, double nms_thresh,
                                    double score_thresh, int32_t *d_count, int32_t *d_out_anchor,
                                    int32_t *d_out_class, float *d_out_score, float *d_out_box, void *stream) {
    if (batch == 0) return SQD_OK;  // empty batch: nothing to enqueue, pointers may be NULL
    SQD_REQUIRE(d_pred && d_anchors, SQD_E_NULL, "sqd_detect_from_pred: pred/anchors is NULL");
    int rc = check_common("sqd_detect_from_pred", batch, num_anchors, num_classes, top_k, d_count, d_out_anchor,
                          d_out_class, d_out_score, d_out_box);
    if (rc) return rc;
    SQD_REQUIRE(sqd_aligned16(d_pred) && sqd_aligned16(d_anchors), SQD_E_ALIGN,
                "sqd_detect_from_pred: pred/anchors must be 16-byte aligned");
    FilterOut o{d_count, d_out_anchor, d_out_class, d_out_score, reinterpret_cast<float4 *>(d_out_box)};
    const size_t smem = dyn_smem_bytes(top_k);
    const float4 *anc = reinterpret_cast<const float4 *>(d_anchors);
    const float wmax = (float)(input_w - 1), hmax = (float)(input_h - 1);
    const float nthr = float_at_or_below(nms_thresh), sthr = (float)score_thresh;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int cs = cluster_size_for(batch, num_anchors, top_k);
    if (num_classes == 3) {
        rc = opt_in_smem(detect_from_pred_kernel<3>, smem);
        if (rc) return rc;
        return launch_clustered("detect_from_pred_kernel", detect_from_pred_kernel<3>, batch, cs, smem, st, d_pred, anc,
                                num_anchors, num_classes, wmax, hmax, top_k, nthr, sthr, o, cs);
    } else if (num_classes == 8) {
        rc = opt_in_smem(detect_from_pred_kernel<8>, smem);
        if (rc) return rc;
        return launch_clustered("detect_from_pred_kernel", detect_from_pred_kernel<8>, batch, cs, smem, st, d_pred, anc,
                                num_anchors, num_classes, wmax, hmax, top_k, nthr, sthr, o, cs);
    }
    rc = opt_in_smem(detect_from_pred_kernel<0>, smem);
    if (rc) return rc;
    return launch_clustered("detect_from_pred_kernel", detect_from_pred_kernel<0>, batch, cs, smem, st, d_pred, anc,
                            num_anchors, num_classes, wmax, hmax, top_k, nthr, sthr, o, cs);
}

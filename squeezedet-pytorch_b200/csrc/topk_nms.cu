// a8-a9: Detector.filter for a whole batch -- exact top-k, per-class greedy NMS, score threshold --
// and its fusion with the decode (sqd_detect_from_pred).
// Reference: src/engine/detector.py:87-122 (full torch.argsort of A scores, 3 torchvision.ops.nms
// calls and >= 3C+3 host syncs PER IMAGE) and torchvision's CPU nms kernel for the IoU arithmetic.
//
// One CTA per image.
//  1. Scan: every thread scores its anchors; a candidate is a 64-bit key
//        [ order-preserving score bits : 32 | 0xFFFFFF - anchor : 24 | class : 8 ]
//     so "larger key" == (score desc, anchor index asc) -- the declared tie policy.
//     Candidates above the running k-th-best threshold are appended to a 2048-entry shared buffer;
//     when a round could overflow it, the buffer is bitonic-sorted, cut to k and the threshold
//     raised.  After the first cut almost nothing passes (expected k*ln(A/2048) more candidates).
//  2. The k survivors (sorted) get their boxes; a k x k same-class IoU bitmask is built with one
//     ballot per 32 pairs; one warp runs the sequential greedy sweep over the mask rows.
//  3. Kept rows with score > thresh are emitted class-ascending / score-descending.
// Algorithmic HBM bytes per image: A*(C+5)*4 (fused) or A*4 (+ a few KB of gathers) for the dense form.
#include "common.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kUnroll = 2;
constexpr int kRound = kThreads * kUnroll;  // anchors consumed per round
constexpr int kCap = 2048;                  // candidate buffer entries (>= SQD_MAX_TOPK + kRound)
static_assert(kCap >= SQD_MAX_TOPK + kRound, "candidate buffer too small");

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

// ---- block-wide bitonic sort (descending) of buf[0..n2), n2 a power of two ------------------------
__device__ void bitonic_desc(u64 *buf, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n2 >> 1); t += kThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair
                const int l = i | j;
                const bool desc = ((i & k) == 0);
                const u64 x = buf[i], y = buf[l];
                if ((x < y) == desc) {
                    buf[i] = y;
                    buf[l] = x;
                }
            }
            __syncthreads();
        }
    }
}

struct Shared {
    u64 buf[kCap];
    int count;
    int n_valid;
    u64 thresh;
};

// Sort the candidates, keep the best k, raise the threshold.  All threads must call.
__device__ void compact(Shared &sh, int k) {
    const int cnt = min(sh.count, kCap);
    int n2 = 2;
    while (n2 < cnt) n2 <<= 1;
    for (int i = cnt + threadIdx.x; i < n2; i += kThreads) sh.buf[i] = 0ull;
    __syncthreads();
    bitonic_desc(sh.buf, n2);
    if (threadIdx.x == 0) {
        if (cnt >= k) {
            sh.count = k;
            sh.thresh = sh.buf[k - 1];
        } else {
            sh.count = cnt;
        }
    }
    __syncthreads();
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
                                                                    float nms_thr_f, float score_thr_f, FilterOut o) {
    __shared__ Shared sh;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int img = blockIdx.x;
    FromPred<CS> src;
    src.pred = pred + (size_t)img * A * ((CS > 0 ? CS : C) + 5);
    src.anchors = anchors;
    src.C = C;
    src.wmax = wmax;
    src.hmax = hmax;
    if (threadIdx.x == 0) {
        sh.count = 0;
        sh.thresh = 0ull;
    }
    __syncthreads();
    for (int base = 0; base < A; base += kRound) {
        const u64 thr = sh.thresh;
        u64 keys[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int a = base + u * kThreads + threadIdx.x;
            keys[u] = a < A ? src.key(a) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (keys[u] > thr) {
                const int pos = atomicAdd(&sh.count, 1);
                if (pos < kCap) sh.buf[pos] = keys[u];
            }
        // Barrier + vote in one: the thread that performs the round's last append observes the final
        // count, so the OR is true for everyone iff the buffer could overflow next round (uniform branch).
        if (__syncthreads_or(*(volatile int *)&sh.count > kCap - kRound)) compact(sh, k);
    }
    __syncthreads();
    compact(sh, k);
    nms_and_emit(src, sh, dyn, k, C, nms_thr_f, score_thr_f, o, img);
}

__global__ void __launch_bounds__(kThreads) filter_dense_kernel(const long long *class_ids, const float *scores,
                                                               const float4 *boxes, int A, int C, int k,
                                                               float nms_thr_f, float score_thr_f, FilterOut o) {
    __shared__ Shared sh;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int img = blockIdx.x;
    FromDense src;
    src.class_ids = class_ids + (size_t)img * A;
    src.scores = scores + (size_t)img * A;
    src.boxes = boxes + (size_t)img * A;
    if (threadIdx.x == 0) {
        sh.count = 0;
        sh.thresh = 0ull;
    }
    __syncthreads();
    for (int base = 0; base < A; base += kRound) {
        const unsigned thr_hi = (unsigned)(sh.thresh >> 32);
        float s[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int a = base + u * kThreads + threadIdx.x;
            s[u] = a < A ? src.score(a) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int a = base + u * kThreads + threadIdx.x;
            if (a < A && order_bits(s[u]) >= thr_hi) {  // cheap pre-test on the score word only
                const u64 key = src.key_of(s[u], a);
                if (key > sh.thresh) {
                    const int pos = atomicAdd(&sh.count, 1);
                    if (pos < kCap) sh.buf[pos] = key;
                }
            }
        }
        if (__syncthreads_or(*(volatile int *)&sh.count > kCap - kRound)) compact(sh, k);
    }
    __syncthreads();
    compact(sh, k);
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
    filter_dense_kernel<<<batch, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long *>(d_class_ids), d_scores, reinterpret_cast<const float4 *>(d_boxes),
        num_anchors, num_classes, top_k, float_at_or_below(nms_thresh), (float)score_thresh, o);
    SQD_LAUNCH_CHECK("filter_dense_kernel");
    return SQD_OK;
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
    if (num_classes == 3) {
        rc = opt_in_smem(detect_from_pred_kernel<3>, smem);
        if (rc) return rc;
        detect_from_pred_kernel<3><<<batch, kThreads, smem, st>>>(d_pred, anc, num_anchors, num_classes, wmax, hmax,
                                                                  top_k, nthr, sthr, o);
    } else if (num_classes == 8) {
        rc = opt_in_smem(detect_from_pred_kernel<8>, smem);
        if (rc) return rc;
        detect_from_pred_kernel<8><<<batch, kThreads, smem, st>>>(d_pred, anc, num_anchors, num_classes, wmax, hmax,
                                                                  top_k, nthr, sthr, o);
    } else {
        rc = opt_in_smem(detect_from_pred_kernel<0>, smem);
        if (rc) return rc;
        detect_from_pred_kernel<0><<<batch, kThreads, smem, st>>>(d_pred, anc, num_anchors, num_classes, wmax, hmax,
                                                                  top_k, nthr, sthr, o);
    }
    SQD_LAUNCH_CHECK("detect_from_pred_kernel");
    return SQD_OK;
}

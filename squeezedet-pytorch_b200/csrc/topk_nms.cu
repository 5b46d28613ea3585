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
//     ballot per 32 pairs (the reference's `float(inter/union) > thresh` decided exactly WITHOUT the division, see
//     nms_and_emit); greedy NMS is then solved block by block (32 candidates) as a fixed point of
//     word' = ballot(pre & (own & word) == 0): no k-step serial chain, typically 2-4 rounds per block.
//  3. Kept rows with score > thresh are emitted class-ascending / score-descending.
// The running threshold starts at the score threshold (exact, see score_floor_key), so on real inputs only a few
// hundred anchors per image ever enter the candidate buffer and mid-scan compactions do not happen.
// Algorithmic HBM bytes per image: A*(C+5)*4 (fused) or A*4 (+ a few KB of gathers) for the dense form.
#include <stdlib.h>

#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 512;
constexpr int kUnroll = 2;
constexpr int kRound = kThreads * kUnroll;  // anchors consumed per round
constexpr int kCap = 2048;                  // candidate buffer entries (>= SQD_MAX_TOPK + kRound)
static_assert(kCap >= SQD_MAX_TOPK + kRound, "candidate buffer too small");
static_assert(SQD_MAX_TOPK <= kCap / 2, "the emit phase handles ceil(SQD_MAX_TOPK / T) candidates per thread, T = 128 or 512");
static_assert(kCap / 2 >= SQD_MAX_TOPK, "rank_sort uses the upper half of the buffer as its destination");

typedef sqd_u64 u64;

// Profiling builds (SQD_BUILD_TRACE=1): clock64 stamps of the tail kernel's phases, image SQD_TAIL_TRACE_IMG, 16 slots
#ifdef SQD_ENABLE_TRACE
__device__ long long *g_tail_trace = nullptr;
__device__ int g_tail_trace_img = 0;
#define TAIL_STAMP(slot) do { if (g_tail_trace && (int)blockIdx.x == g_tail_trace_img && threadIdx.x == 0) g_tail_trace[slot] = clock64(); } while (0)
#else
#define TAIL_STAMP(slot) do { } while (0)
#endif

__device__ __forceinline__ unsigned order_bits(float s) { return sqd_order_bits(s); }
__device__ __forceinline__ u64 make_key(float score, int anchor, int cls) { return sqd_make_key(score, anchor, cls); }
__device__ __forceinline__ int key_anchor(u64 k) { return (int)(0xFFFFFFu - (unsigned)((k >> 8) & 0xFFFFFFu)); }
__device__ __forceinline__ int key_class(u64 k) { return (int)(k & 0xFFu); }
__device__ __forceinline__ float key_score(u64 k) { return sqd_unorder_bits((unsigned)(k >> 32)); }
// The running threshold starts at the score threshold: exact, see sqd_score_floor_key (common.cuh).
__device__ __forceinline__ u64 score_floor_key(float score_thr) { return sqd_score_floor_key(score_thr); }

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
template <int T = kThreads>
__device__ void select_topk(Shared &sh, int k) {
    constexpr int kPer = kCap / T;
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
        const int idx = threadIdx.x + i * T;
        my[i] = idx < cnt ? sh.buf[idx] : 0ull;  // 0 is below every real key (score bits of a finite float are never 0)
    }
    const int lane = threadIdx.x & 31;
    u64 prefix = 0ull;
    int need = k, shift = 56;
    for (; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += T) sh.hist[i] = 0;
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
template <int T = kThreads>
__device__ void rank_sort(Shared &sh) {
    const int m = sh.count;
    u64 *out = sh.buf + kCap / 2;
    for (int i = threadIdx.x; i < m; i += T) {
        const u64 key = sh.buf[i];
        int rank = 0;
        for (int j = 0; j < m; ++j) rank += sh.buf[j] > key ? 1 : 0;
        out[rank] = key;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += T) sh.buf[i] = out[i];
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

// Phase 2+3, shared by all sources.  sh.buf[0..m) holds the sorted survivors (m <= k <= 2*kThreads).
//  * suppressor bitsets: word w of candidate i has bit b set iff j = 32w+b < i, same class and IoU(i,j) > thresh
//  * greedy sweep without a 64-step serial chain: keep(i) = !exists j<i: keep(j) & sup(j,i) is solved block by block
//    (32 candidates per block, final keep words of earlier blocks applied first); inside a block the warp iterates
//    word' = ballot(pre & (own & word) == 0) to its fixed point -- after r rounds the first r lanes are final, and a
//    fixed point satisfies the defining recurrence, whose solution is unique.  Typically 2-4 rounds per block.
//  * output position = rank of (class, index) among the kept rows: class ascending / score descending.
template <class Src, int T = kThreads>
__device__ void nms_and_emit(const Src &src, Shared &sh, unsigned char *dyn, int k, int num_classes, float nms_thr_f,
                             float score_thr_f, const FilterOut &o, int img) {
    const int m = sh.count;
    const int wpr = (k + 31) >> 5;  // mask words per candidate
    float4 *sbox = reinterpret_cast<float4 *>(dyn);
    float *sarea = reinterpret_cast<float *>(sbox + k);
    unsigned *col = reinterpret_cast<unsigned *>(sarea + k);
    unsigned *okey = col + (size_t)k * wpr;
    unsigned *keepw = okey + k;

    for (int i = threadIdx.x; i < m; i += T) {
        const float4 b = src.box(key_anchor(sh.buf[i]));
        sbox[i] = b;
        sarea[i] = fmul(fsub(b.z, b.x), fsub(b.w, b.y));  // torchvision: (x2-x1)*(y2-y1), no +1
    }
    __syncthreads();
    TAIL_STAMP(4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = T >> 5;
    // One (candidate i, 32-candidate word w) item per warp step, lanes = the j of the word.  Four items are evaluated
    // together (independent dependency chains), and only lanes whose boxes intersect at all are tested (inter > 0:
    // otherwise the quotient is 0 or NaN and cannot exceed a threshold >= 0).
    // The reference's test is RN32(inter / uni) > thr (torchvision: float division, strict >).  For 0 <= thr and a finite
    // positive divisor that is decided WITHOUT the ~40-instruction IEEE division: with thr+ the next float above thr,
    //   RN32(q) > thr  <=>  RN32(q) >= thr+  <=>  q > mid, or q == mid and the tie rounds to thr+ (its mantissa is even),
    // mid = (thr + thr+)/2, and q vs mid is inter vs mid*uni -- exact in double (25 x 24 significant bits).  With
    // inter > 0 the only divisor the product form gets wrong is a negative one (quotient < 0 <= thr: never suppresses;
    // uni is never -0 here): uni == 0 -> +inf > thr, uni == +inf -> 0, NaN -> false all come out right, and inter == inf
    // makes uni -inf or NaN.  Negative (or absurdly large) thresholds take the division.
    const float thr_up = __uint_as_float(__float_as_uint(nms_thr_f) + 1u);
    const bool fast_ok = nms_thr_f >= 0.f && nms_thr_f < 1e30f;
    const double thr_mid = 0.5 * ((double)nms_thr_f + (double)thr_up);
    const bool tie_up = (__float_as_uint(thr_up) & 1u) == 0u;
    // Word by word: a lane keeps candidate j = 32w + lane of the word (box, area, class) in registers and the warp walks
    // the candidates i >= 32w it owns (every nwarp-th), four at a time.  Branch free -- invalid lanes are masked at the
    // end -- so the four items' shared-memory loads and arithmetic overlap; the division path is taken by a warp only if
    // one of its lanes needs it.  Item (i, w) with 32w == i has no j < i and writes 0 (the sweep reads it as `own`).
    for (int w = 0; (w << 5) < m; ++w) {
        const int j = (w << 5) + lane, jc = min(j, m - 1);
        const float4 bj = sbox[jc];
        const float aj = sarea[jc];
        const int cj = key_class(sh.buf[jc]);
        for (int i0 = (w << 5) + warp; i0 < m; i0 += 4 * nwarp) {
            unsigned bits[4];
            if (fast_ok) {   // uniform.  Bitwise (not short-circuit) logic: no branches, the four chains interleave
                float4 a[4];
                float ai[4];
                int ci[4], ii[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ii[u] = min(i0 + u * nwarp, m - 1);
                    a[u] = sbox[ii[u]];
                    ai[u] = sarea[ii[u]];
                    ci[u] = key_class(sh.buf[ii[u]]);
                }
                bool sup[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float iw = fmaxf(0.f, fsub(fminf(a[u].z, bj.z), fmaxf(a[u].x, bj.x)));
                    const float ih = fmaxf(0.f, fsub(fminf(a[u].w, bj.w), fmaxf(a[u].y, bj.y)));
                    const float inter = fmul(iw, ih);
                    const float uni = fsub(fadd(ai[u], aj), inter);
                    const double prod = thr_mid * (double)uni, di = (double)inter;
                    const bool cmp = tie_up ? di >= prod : di > prod;
                    sup[u] = (j < ii[u]) & (cj == ci[u]) & (inter > 0.f) & !(uni < 0.f) & cmp;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) bits[u] = __ballot_sync(0xffffffffu, sup[u]);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = min(i0 + u * nwarp, m - 1);
                    const float4 a = sbox[i];
                    const float iw = fmaxf(0.f, fsub(fminf(a.z, bj.z), fmaxf(a.x, bj.x)));
                    const float ih = fmaxf(0.f, fsub(fminf(a.w, bj.w), fmaxf(a.y, bj.y)));
                    const float inter = fmul(iw, ih);
                    const float uni = fsub(fadd(sarea[i], aj), inter);
                    const bool same = j < i && cj == key_class(sh.buf[i]);
                    const bool sup = same && (inter > 0.f || nms_thr_f < 0.f) && fdiv(inter, uni) > nms_thr_f;  // false for NaN
                    bits[u] = __ballot_sync(0xffffffffu, sup);
                }
            }
            if (lane < 4) {
                const int i = i0 + lane * nwarp;
                const unsigned mine = lane == 0 ? bits[0] : lane == 1 ? bits[1] : lane == 2 ? bits[2] : bits[3];
                if (i < m) col[i * wpr + w] = mine;
            }
        }
    }
    __syncthreads();
    TAIL_STAMP(5);

    if (warp == 0) {
        for (int t = 0; (t << 5) < m; ++t) {
            const int i = (t << 5) + lane;
            bool pre = i < m;
            for (int w = 0; w < t; ++w) pre = pre && (col[i * wpr + w] & keepw[w]) == 0u;
            const unsigned own = i < m ? col[i * wpr + t] : 0u;
            unsigned word = __ballot_sync(0xffffffffu, pre);
            for (;;) {
                const unsigned nw = __ballot_sync(0xffffffffu, pre && (own & word) == 0u);
                if (nw == word) break;
                word = nw;
            }
            if (lane == 0) keepw[t] = word;
            __syncwarp();
        }
    }
    __syncthreads();
    TAIL_STAMP(6);

    // emit: class ascending, then descending score (== position) inside a class
    constexpr int kEmitPer = (SQD_MAX_TOPK + T - 1) / T;   // candidates per thread (m <= k <= SQD_MAX_TOPK)
    unsigned mine[kEmitPer];
    int n_valid = 0;
#pragma unroll
    for (int h = 0; h < kEmitPer; ++h) {
        const int t = threadIdx.x + h * T;
        bool v = false;
        if (t < m) v = ((keepw[t >> 5] >> (t & 31)) & 1u) && key_score(sh.buf[t]) > score_thr_f;
        mine[h] = v ? ((unsigned)key_class(sh.buf[t]) << 16) | (unsigned)t : 0xFFFFFFFFu;
        if (t < m) okey[t] = mine[h];
        n_valid += __syncthreads_count(v);
    }
#pragma unroll
    for (int h = 0; h < kEmitPer; ++h) {
        const int t = threadIdx.x + h * T;
        if (mine[h] == 0xFFFFFFFFu) continue;
        int pos = 0;
        for (int u = 0; u < m; ++u) pos += okey[u] < mine[h] ? 1 : 0;
        const size_t r = (size_t)img * k + pos;
        o.anchor[r] = key_anchor(sh.buf[t]);
        o.cls[r] = key_class(sh.buf[t]);
        o.score[r] = key_score(sh.buf[t]);
        o.box[r] = sbox[t];
    }
    if (threadIdx.x == 0) o.count[img] = n_valid;
    for (int t = n_valid + threadIdx.x; t < k; t += T) {  // deterministic padding rows
        const size_t r = (size_t)img * k + t;
        o.anchor[r] = -1;
        o.cls[r] = -1;
        o.score[r] = 0.f;
        o.box[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    TAIL_STAMP(7);
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

// ---- two-phase form: (1) streaming scan -> per-image candidate lists, (2) per-image select / NMS / emit --------------
// Phase 1 is a pure HBM stream over pred (the whole grid works on the whole batch, no per-image serial phases);
// only anchors above the score threshold (exact pre-filter) are appended, typically 1-3 % of them.  Phase 2 reads
// those few keys and runs the same select -> sort -> NMS -> emit tail as the one-kernel form.  The tcgen05 ConvDet
// epilogue can produce the candidate lists itself (convdet_f16.cu), in which case phase 1 never runs and pred is
// only touched for the <= k surviving boxes of an image.
constexpr int kScanThreads = 256;
constexpr int kScanUnroll = 4;

// grid = (chunks of an image, images): every row of a block belongs to image blockIdx.y, so there is no index
// division and a warp appends all its candidates with ONE atomic.
template <int CS>
__global__ void __launch_bounds__(kScanThreads) score_candidates_kernel(const float *__restrict__ pred, int A, int C,
                                                                        float score_thr_f, SqdCand cand) {
    const u64 floor_key = score_floor_key(score_thr_f);
    const int img = blockIdx.y;
    sqd_pdl_trigger();   // the tail kernel may be scheduled; it waits for this grid to complete
    sqd_pdl_wait();      // pred (written by the preceding kernel of the chain) is complete and visible
    if (CS == 3) {
        // 32-byte rows: the four scoring fields are the first 16 bytes of a row
        const float4 *p4 = reinterpret_cast<const float4 *>(pred) + 2 * (size_t)img * A;
        for (int base = blockIdx.x * kScanThreads * kScanUnroll; base < A; base += gridDim.x * kScanThreads * kScanUnroll) {
            float4 v[kScanUnroll];
#pragma unroll
            for (int u = 0; u < kScanUnroll; ++u) {
                const int a = base + u * kScanThreads + threadIdx.x;
                v[u] = a < A ? ld_stream_f4(p4 + 2 * (size_t)a) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            bool pass[kScanUnroll];
            u64 key[kScanUnroll];
#pragma unroll
            for (int u = 0; u < kScanUnroll; ++u) {
                const int a = base + u * kScanThreads + threadIdx.x;
                const float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                float s = 0.f;
                int c = 0;
                const bool cand_ok = sqd_score_candidate<3>(f, 3, score_thr_f, s, c);
                key[u] = make_key(s, a, c);
                pass[u] = a < A && cand_ok && key[u] > floor_key;
            }
            sqd_cand_append_warp<kScanUnroll>(cand, img, pass, key);
        }
    } else {
        // any field count: stage 256 contiguous rows through shared memory with 16-byte streaming loads
        extern __shared__ __align__(16) float slab[];
        const int Cn = CS > 0 ? CS : C;
        const int NF = Cn + 5;
        const float *ipred = pred + (size_t)img * A * NF;
        for (int row0 = blockIdx.x * kScanThreads; row0 < A; row0 += gridDim.x * kScanThreads) {
            const int n = min(kScanThreads, A - row0);
            const float *src = ipred + (size_t)row0 * NF;
            // 16-byte loads when this image's rows start 16-byte aligned (256*NF*4 is a multiple of 16, so then every
            // slab of the image is); otherwise (A*NF odd multiples of 4 bytes) plain 4-byte loads
            const int count = n * NF;
            const int nvec = (reinterpret_cast<uintptr_t>(src) & 15u) == 0 ? count >> 2 : 0;
            const float4 *src4 = reinterpret_cast<const float4 *>(src);
            float4 *dst4 = reinterpret_cast<float4 *>(slab);
            for (int i = threadIdx.x; i < nvec; i += kScanThreads) dst4[i] = ld_stream_f4(src4 + i);
            for (int i = (nvec << 2) + threadIdx.x; i < count; i += kScanThreads) slab[i] = src[i];
            __syncthreads();
            const bool in = (int)threadIdx.x < n;
            float f[SQD_CMAX(CS) + 1];
#pragma unroll
            for (int j = 0; j < SQD_CMAX(CS) + 1; ++j)
                if (j <= Cn) f[j] = in ? slab[threadIdx.x * NF + j] : 0.f;
            float s = 0.f;
            int c = 0;
            const bool cand_ok = sqd_score_candidate<CS>(f, Cn, score_thr_f, s, c);
            bool pass[1];
            u64 key[1];
            key[0] = make_key(s, row0 + (int)threadIdx.x, c);
            pass[0] = in && cand_ok && key[0] > floor_key;
            sqd_cand_append_warp<1>(cand, img, pass, key);
            __syncthreads();
        }
    }
}

// Exact top-k of a long candidate list without sorting or compaction rounds: MSB-first radix select on the 32 score
// bits with 11-bit digits (a 2048-bin shared histogram per level, keys re-read from L2), stopping as soon as the keys
// at or above the threshold bin fit the sort buffer; those are then collected (unordered) into sh.buf.  Typical lists
// (thousands of distinct scores) need ONE level: histogram pass + collect pass.  Returns false (nothing collected) if
// even single-score bins overflow the buffer (massive exact ties): the caller then uses the running-threshold loop.
template <bool kInRegs, int T = kThreads>
__device__ bool hist_select_collect(Shared &sh, const u64 *__restrict__ keys, int n, int k, float score_thr_f) {
    constexpr int kBins = 2048, kPerThread = kBins / T;
    constexpr int kCollectCap = kCap / 2;
    constexpr int kHold = 16;  // keys per thread kept in registers when the list has <= kHold*kThreads entries
    int *hist = reinterpret_cast<int *>(sh.buf + kCap / 2);  // upper half of the key buffer: free until the sort
    // bin b lives at hist[b ^ ((b >> 5) & 31)]: thread t reads its kPerThread CONSECUTIVE bins one per step, which without
    // the swizzle is a 16-way (128 threads) / 4-way (512 threads) bank conflict per step; with it every step is conflict free
    auto hpos = [](unsigned b) { return b ^ ((b >> 5) & 31u); };
    __shared__ int s_wtot[T / 32];
    __shared__ int s_tb, s_above_add, s_tb_cnt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 my[kInRegs ? kHold : 1];
    if (kInRegs) {  // ONE pass over the list, all loads in flight together; 0 = no key (never a real key)
#pragma unroll
        for (int u = 0; u < kHold; ++u) {
            const int i = threadIdx.x + u * T;
            my[u] = i < n ? __ldcg(keys + i) : 0ull;
        }
    }
    // visits every key of the list: from registers, or re-read from L2 eight at a time
    auto for_each_key = [&](auto &&fn) {
        if (kInRegs) {
#pragma unroll
            for (int u = 0; u < kHold; ++u)
                if (my[u] != 0ull) fn(my[u]);
        } else {
            for (int base = 0; base < n; base += 8 * T) {
                u64 t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = base + u * T + threadIdx.x;
                    t[u] = i < n ? __ldcg(keys + i) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (t[u] != 0ull) fn(t[u]);
            }
        }
    };
    TAIL_STAMP(10);
    unsigned lo = order_bits(score_thr_f) + 1u, hi = 0xFFFFFFFFu;  // undecided score-bit range (inclusive)
    int above = 0;                                                 // keys with score bits > hi: selected for sure
    const unsigned top = order_bits(1.0f);                         // scores are probabilities: <= 1
    const unsigned span = (top > lo ? top : lo) - lo;
    int shift = 0;
    while ((span >> shift) >= (unsigned)kBins) ++shift;
    for (int level = 0; level < 4; ++level) {
        for (int i = threadIdx.x; i < kBins; i += T) hist[i] = 0;
        __syncthreads();
        if (level == 0) TAIL_STAMP(11);
        for_each_key([&](u64 key) {
            const unsigned sb = (unsigned)(key >> 32);
            if (sb >= lo && sb <= hi) atomicAdd(&hist[hpos(min((sb - lo) >> shift, (unsigned)(kBins - 1)))], 1);
        });
        __syncthreads();
        if (level == 0) TAIL_STAMP(12);
        int c[kPerThread], own = 0;
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
            c[j] = hist[hpos(threadIdx.x * kPerThread + j)];
            own += c[j];
        }
        int suf = own;  // inclusive suffix sum over the lanes >= this one
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_down_sync(0xffffffffu, suf, o);
            if (lane + o < 32) suf += v;
        }
        if (lane == 0) s_wtot[warp] = suf;
        __syncthreads();
        int higher = 0;
        for (int w = warp + 1; w < T / 32; ++w) higher += s_wtot[w];
        const int incl = suf + higher, excl = incl - own;  // keys in bins of threads >= / > this one
        const int need = k - above;
        if (excl < need && need <= incl) {  // exactly one thread: the range holds at least `need` keys
            int acc = excl;
#pragma unroll
            for (int j = kPerThread - 1; j >= 0; --j) {
                if (acc + c[j] >= need) {
                    s_tb = threadIdx.x * kPerThread + j;
                    s_above_add = acc;
                    s_tb_cnt = c[j];
                    break;
                }
                acc += c[j];
            }
        }
        __syncthreads();
        if (level == 0) TAIL_STAMP(13);
        const int tb = s_tb;
        const unsigned lo_new = lo + ((unsigned)tb << shift);
        const int above_new = above + s_above_add;
        if (above_new + s_tb_cnt <= kCollectCap) {
            // collect every key at or above the threshold bin (>= k of them, <= kCollectCap)
            __syncthreads();  // all reads of hist / s_* done before buf and the counter are written
            for_each_key([&](u64 key) {
                if ((unsigned)(key >> 32) >= lo_new) sh.buf[atomicAdd(&sh.count, 1)] = key;
            });
            __syncthreads();
            return true;
        }
        if (shift == 0) break;  // one exact score value with too many anchors: give up (uniform)
        if (tb != kBins - 1) hi = lo_new + ((1u << shift) - 1u);
        lo = lo_new;
        above = above_new;
        shift = shift > 11 ? shift - 11 : 0;
        __syncthreads();
    }
    return false;
}

// Phase 2: one CTA per image over its candidate list (a few hundred keys on real inputs; at most A).
// T threads per CTA: 512 for small batches (the shortest latency chain per image), 128 for large ones -- the kernel is a
// chain of short barrier-separated phases, so what a full GPU needs is MANY resident CTAs (8 per SM at 128 threads
// against 2 at 512: 1024 images in one wave instead of 3.5).
template <int T>
__global__ void __launch_bounds__(T, T == 128 ? 8 : 1) detect_from_candidates_kernel(SqdCand cand, const float *pred,
                                                                   const float4 *anchors, int A, int C, float wmax,
                                                                   float hmax, int k, float nms_thr_f,
                                                                   float score_thr_f, FilterOut o) {
    constexpr int kRoundT = T * kUnroll;
    __shared__ Shared sh;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int img = blockIdx.x;
    TAIL_STAMP(0);
    sqd_pdl_wait();      // candidate lists and pred are complete and visible
    TAIL_STAMP(1);
    const int n = min(cand.count[img], cand.stride);
    const u64 *keys = cand.keys + (size_t)img * cand.stride;
    FromPred<0> src;
    src.pred = pred + (size_t)img * A * (C + 5);
    src.anchors = anchors;
    src.C = C;
    src.wmax = wmax;
    src.hmax = hmax;
    if (threadIdx.x == 0) {
        sh.count = 0;
        sh.thresh = score_floor_key(score_thr_f);
    }
    __syncthreads();
    bool have = false;
    if (n <= max(k + 192, 256) && n <= kCap / 2) {
        // short list: sort it directly
        for (int i = threadIdx.x; i < n; i += T) sh.buf[i] = __ldcg(keys + i);
        if (threadIdx.x == 0) sh.count = n;
        __syncthreads();
        have = true;
    } else {
        have = n <= 16 * T ? hist_select_collect<true, T>(sh, keys, n, k, score_thr_f)
                           : hist_select_collect<false, T>(sh, keys, n, k, score_thr_f);
    }
    TAIL_STAMP(2);
    if (have) {
        rank_sort<T>(sh);  // descending: the first k entries are the top-k
        if (threadIdx.x == 0 && sh.count > k) sh.count = k;
        __syncthreads();
    } else {
        // fallback (massive exact score ties): running-threshold scan with exact radix selects
        if (threadIdx.x == 0) sh.count = 0;
        __syncthreads();
        for (int base = 0; base < n; base += kRoundT) {
            const u64 thr = sh.thresh;
            u64 kk[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int i = base + u * T + threadIdx.x;
                kk[u] = i < n ? __ldcg(keys + i) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                if (kk[u] > thr) {
                    const int pos = atomicAdd(&sh.count, 1);
                    if (pos < kCap) sh.buf[pos] = kk[u];
                }
            if (__syncthreads_or(*(volatile int *)&sh.count > kCap - kRoundT)) select_topk<T>(sh, k);
        }
        select_topk<T>(sh, k);
        rank_sort<T>(sh);
    }
    TAIL_STAMP(3);
#ifdef SQD_ENABLE_TRACE
    if (g_tail_trace && (int)blockIdx.x == g_tail_trace_img && threadIdx.x == 0) {
        g_tail_trace[8] = n;
        g_tail_trace[9] = sh.count;
    }
#endif
    nms_and_emit<FromPred<0>, T>(src, sh, dyn, k, C, nms_thr_f, score_thr_f, o, img);
}

size_t dyn_smem_bytes(int k) {
    const size_t wpr = (k + 31) / 32;
    return (size_t)k * 16 + (size_t)k * 4 + (size_t)k * wpr * 4 + (size_t)k * 4 + wpr * 4 + 16;
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

// ---- two-phase host side (also used by api.cu for the fused head) -------------------------------------------
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

size_t sqd_cand_bytes(int batch, int num_anchors) {
    return align256((size_t)batch * sizeof(int)) + align256((size_t)batch * num_anchors * sizeof(u64));
}

SqdCand sqd_cand_layout(void *ws, int batch, int num_anchors) {
    SqdCand c;
    c.count = static_cast<int *>(ws);
    c.keys = reinterpret_cast<u64 *>(static_cast<char *>(ws) + align256((size_t)batch * sizeof(int)));
    c.stride = num_anchors;
    return c;
}

// phase 1: pred -> candidate lists (cand.count must have been zeroed on the stream)
// pdl: the launch directly follows the kernel that produced d_pred on `st` (programmatic dependent launch)
int sqd_score_candidates(const float *d_pred, int batch, int num_anchors, int num_classes, double score_thresh,
                         SqdCand cand, cudaStream_t st, bool pdl) {
    SQD_REQUIRE(batch <= 65535, SQD_E_SHAPE, "detect: batch %d > 65535 (split the call)", batch);
    const float sthr = (float)score_thresh;
    const int per_block = num_classes == 3 ? kScanThreads * kScanUnroll : kScanThreads;
    const int gx = (num_anchors + per_block - 1) / per_block;   // one pass per block (the in-kernel loop is for safety)
    const dim3 grid((unsigned)gx, (unsigned)batch);
    cudaError_t e;
    if (num_classes == 3) {
        e = sqd_launch_dependent(score_candidates_kernel<3>, grid, dim3(kScanThreads), 0, st, pdl, d_pred, num_anchors, 3, sthr, cand);
    } else {
        const size_t smem = (size_t)kScanThreads * (num_classes + 5) * sizeof(float);
        if (num_classes == 8)
            e = sqd_launch_dependent(score_candidates_kernel<8>, grid, dim3(kScanThreads), smem, st, pdl, d_pred, num_anchors, 8, sthr, cand);
        else
            e = sqd_launch_dependent(score_candidates_kernel<0>, grid, dim3(kScanThreads), smem, st, pdl, d_pred, num_anchors,
                                     num_classes, sthr, cand);
    }
    if (e != cudaSuccess) {
        sqd_set_error("launch of score_candidates_kernel failed: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return SQD_OK;
}

// phase 2: candidate lists -> final detections
int sqd_detect_from_candidates(SqdCand cand, const float *d_pred, const float *d_anchors, int batch, int num_anchors,
                               int num_classes, int input_h, int input_w, int top_k, double nms_thresh,
                               double score_thresh, int32_t *d_count, int32_t *d_out_anchor, int32_t *d_out_class,
                               float *d_out_score, float *d_out_box, cudaStream_t st) {
    FilterOut o{d_count, d_out_anchor, d_out_class, d_out_score, reinterpret_cast<float4 *>(d_out_box)};
    const size_t smem = dyn_smem_bytes(top_k);
    // large batches: 128-thread CTAs, 8 per SM, so that every image's tail is resident at once (SQD_TAIL_THREADS forces)
    const int tail_opt = sqd_opt(SQD_OPT_TAIL_THREADS);
    // (measured, KITTI clustered pred: 1024 images 115.4 us with 128 threads vs 127.5 with 512; 256 images 53.5 vs 41.8 --
    // up to two waves of 512-thread CTAs are the shorter chain)
    const bool small_cta = tail_opt ? tail_opt == 128 : batch > 4 * SQD_SM_COUNT;
    int rc = small_cta ? opt_in_smem(detect_from_candidates_kernel<128>, smem) : opt_in_smem(detect_from_candidates_kernel<kThreads>, smem);
    if (rc) return rc;
#ifdef SQD_ENABLE_TRACE
    {
        long long *tp = nullptr;
        if (const char *e = getenv("SQD_TAIL_TRACE")) tp = reinterpret_cast<long long *>(strtoull(e, nullptr, 0));
        const int ti = getenv("SQD_TAIL_TRACE_IMG") ? atoi(getenv("SQD_TAIL_TRACE_IMG")) : 0;
        cudaMemcpyToSymbolAsync(g_tail_trace, &tp, sizeof(tp), 0, cudaMemcpyHostToDevice, st);
        cudaMemcpyToSymbolAsync(g_tail_trace_img, &ti, sizeof(ti), 0, cudaMemcpyHostToDevice, st);
    }
#endif
    // always a dependent launch: its predecessor on the stream is the scan (or the GEMM whose epilogue scored)
    const float4 *anc = reinterpret_cast<const float4 *>(d_anchors);
    const float wmax = (float)(input_w - 1), hmax = (float)(input_h - 1), nthr = float_at_or_below(nms_thresh);
    cudaError_t e = small_cta
        ? sqd_launch_dependent(detect_from_candidates_kernel<128>, dim3(batch), dim3(128), smem, st, true, cand, d_pred, anc,
                               num_anchors, num_classes, wmax, hmax, top_k, nthr, (float)score_thresh, o)
        : sqd_launch_dependent(detect_from_candidates_kernel<kThreads>, dim3(batch), dim3(kThreads), smem, st, true, cand, d_pred,
                               anc, num_anchors, num_classes, wmax, hmax, top_k, nthr, (float)score_thresh, o);
    if (e != cudaSuccess) {
        sqd_set_error("launch of detect_from_candidates_kernel failed: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return SQD_OK;
}

int sqd_detect_check_args(const char *fn, const void *d_pred, const void *d_anchors, int batch, int num_anchors,
                          int num_classes, int top_k, const void *count, const void *anchor, const void *cls,
                          const void *score, const void *box) {
    SQD_REQUIRE(d_pred && d_anchors, SQD_E_NULL, "%s: pred/anchors is NULL", fn);
    int rc = check_common(fn, batch, num_anchors, num_classes, top_k, count, anchor, cls, score, box);
    if (rc) return rc;
    SQD_REQUIRE(sqd_aligned16(d_pred) && sqd_aligned16(d_anchors), SQD_E_ALIGN, "%s: pred/anchors must be 16-byte aligned", fn);
    return SQD_OK;
}

extern "C" size_t sqd_detect_workspace_bytes(int batch, int num_anchors) {
    if (batch <= 0 || num_anchors <= 0) return 256;
    return sqd_cand_bytes(batch, num_anchors);
}

extern "C" int sqd_detect_from_pred(const float *d_pred, const float *d_anchors, int batch, int num_anchors,
                                    int num_classes, int input_h, int input_w, int top_k, double nms_thresh,
                                    double score_thresh, int32_t *d_count, int32_t *d_out_anchor,
                                    int32_t *d_out_class, float *d_out_score, float *d_out_box, void *d_workspace,
                                    size_t workspace_bytes, void *stream) {
    if (batch == 0) return SQD_OK;  // empty batch: nothing to enqueue, pointers may be NULL
    int rc = sqd_detect_check_args("sqd_detect_from_pred", d_pred, d_anchors, batch, num_anchors, num_classes, top_k,
                                   d_count, d_out_anchor, d_out_class, d_out_score, d_out_box);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d_workspace) {
        // two-phase form: streaming scan -> candidate lists -> per-image tail
        SQD_REQUIRE(workspace_bytes >= sqd_cand_bytes(batch, num_anchors), SQD_E_WORKSPACE,
                    "sqd_detect_from_pred: workspace too small (%zu bytes)", workspace_bytes);
        SQD_REQUIRE(sqd_aligned16(d_workspace), SQD_E_ALIGN, "sqd_detect_from_pred: workspace must be 16-byte aligned");
        const SqdCand cand = sqd_cand_layout(d_workspace, batch, num_anchors);
        SQD_CUDA(cudaMemsetAsync(cand.count, 0, (size_t)batch * sizeof(int), st));
        rc = sqd_score_candidates(d_pred, batch, num_anchors, num_classes, score_thresh, cand, st, false);  // follows a memset
        if (rc) return rc;
        return sqd_detect_from_candidates(cand, d_pred, d_anchors, batch, num_anchors, num_classes, input_h, input_w, top_k,
                                          nms_thresh, score_thresh, d_count, d_out_anchor, d_out_class, d_out_score,
                                          d_out_box, st);
    }
    // workspace-free form: one launch, a cluster of CTAs per image scans, selects and filters
    FilterOut o{d_count, d_out_anchor, d_out_class, d_out_score, reinterpret_cast<float4 *>(d_out_box)};
    const size_t smem = dyn_smem_bytes(top_k);
    const float4 *anc = reinterpret_cast<const float4 *>(d_anchors);
    const float wmax = (float)(input_w - 1), hmax = (float)(input_h - 1);
    const float nthr = float_at_or_below(nms_thresh), sthr = (float)score_thresh;
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

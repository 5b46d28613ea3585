// a11-a13: training-side anchor <-> ground-truth matching and the dense target tensor.
// Reference: numpy compute_overlaps / compute_deltas, src/utils/boxes.py:70-135 (one full
// np.argsort of A IoUs per GT box, in DataLoader workers) and BaseDataset.prepare_annotations,
// src/datasets/base.py:61-76.
//
// One thread-block CLUSTER per image (8 CTAs: the anchors are split between them, so a GT box costs 4 anchors per
// thread instead of 66); GT boxes are processed sequentially (the greedy assignment is order dependent by
// definition), each one as a masked arg-max over the A anchors in float64 with the reference's exact operation order
// (no FMA contraction), so equal IoUs are bit-equal here exactly when they are in numpy: block arg-max per CTA, the
// CTAs' candidates meet in rank 0 through distributed shared memory, and the winner is broadcast back.  Tie policy:
// lowest anchor index (a stable argsort in the reference) -- the key (value, index) is totally ordered, so the result
// does not depend on how the anchors are partitioned.  "Taken" anchors live in a shared-memory bitmask (each CTA keeps
// the bits of its own anchors).
// Batched rounds: the greedy order only matters when two boxes want the same anchor, which is rare.  A round computes the
// best untaken anchor of up to 32 pending boxes in ONE pass (anchor geometry in registers, all candidates exchanged with
// one pair of cluster barriers) and accepts them in annotation order up to the first box whose candidate was just
// taken by an earlier box of the round (or that has no overlapping anchor: distance fallback, one box the old way);
// the rest is recomputed in the next round.  Every round settles at least one box, so the result is the sequential
// one by construction; typical images need one or two rounds instead of G.  Bytes per image: G * A * 32 (float64 anchor table, L2-resident across the batch).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 512;
constexpr int kMaxCluster = 8;
constexpr int kRoundBoxes = 32;   // boxes settled per batched round

struct Best {
    double v;
    int idx;
};

// larger v wins; ties -> lower index.  idx == INT_MAX marks "none".
__device__ __forceinline__ Best better_max(Best a, Best b) {
    if (b.idx == 0x7fffffff) return a;
    if (a.idx == 0x7fffffff) return b;
    if (b.v > a.v || (b.v == a.v && b.idx < a.idx)) return b;
    return a;
}
__device__ __forceinline__ Best better_min(Best a, Best b) {
    if (b.idx == 0x7fffffff) return a;
    if (a.idx == 0x7fffffff) return b;
    if (b.v < a.v || (b.v == a.v && b.idx < a.idx)) return b;
    return a;
}

template <bool kMax>
__device__ Best block_reduce(Best x, Best *scratch) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Best y;
        y.v = __shfl_down_sync(0xffffffffu, x.v, off);
        y.idx = __shfl_down_sync(0xffffffffu, x.idx, off);
        x = kMax ? better_max(x, y) : better_min(x, y);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();  // scratch reuse across calls
    if (lane == 0) scratch[warp] = x;
    __syncthreads();
    if (warp == 0) {
        x = lane < (kThreads >> 5) ? scratch[lane] : Best{0.0, 0x7fffffff};
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            Best y;
            y.v = __shfl_down_sync(0xffffffffu, x.v, off);
            y.idx = __shfl_down_sync(0xffffffffu, x.idx, off);
            x = kMax ? better_max(x, y) : better_min(x, y);
        }
        if (lane == 0) scratch[0] = x;
    }
    __syncthreads();
    return scratch[0];
}

// Cluster-wide arg-max / arg-min: every CTA contributes its block result, all CTAs get the winner.  `xchg` counts
// the exchanges of this CTA (identical in all CTAs of the cluster): slots alternate so that two cluster barriers per
// exchange are enough.  All threads of all CTAs of the cluster must call.
template <bool kMax>
__device__ Best cluster_reduce(Best mine, int cs, int rank, int &xchg, Best (*s_cand)[kMaxCluster], Best *s_win) {
    if (cs == 1) return mine;
    cg::cluster_group cluster = cg::this_cluster();
    const int slot = xchg & 1;
    ++xchg;
    if (threadIdx.x == 0) cluster.map_shared_rank(&s_cand[slot][0], 0)[rank] = mine;
    cluster.sync();
    if (rank == 0 && threadIdx.x == 0) {
        Best w = s_cand[slot][0];
        for (int r = 1; r < cs; ++r) w = kMax ? better_max(w, s_cand[slot][r]) : better_min(w, s_cand[slot][r]);
        for (int r = 0; r < cs; ++r) *cluster.map_shared_rank(&s_win[slot], r) = w;
    }
    cluster.sync();
    return s_win[slot];
}

__global__ void __launch_bounds__(kThreads) match_kernel(const float4 *gt_boxes, const int *gt_count, int gmax,
                                                         const double *anchors, int A, int *out_idx,
                                                         float4 *out_deltas, int cs, int no_batch) {
    extern __shared__ unsigned taken[];  // bits of this CTA's anchors [a_begin, a_end): ceil(chunk/32) words
    __shared__ Best scratch[kThreads / 32];
    __shared__ Best s_cand[2][kMaxCluster];
    __shared__ Best s_win[2];
    const int img = blockIdx.x / cs, rank = blockIdx.x - img * cs;
    const int chunk = ((A + cs - 1) / cs + 31) & ~31;           // multiple of 32: bitmask words never straddle CTAs
    const int a_begin = min(A, rank * chunk), a_end = min(A, a_begin + chunk);
    const int G = min(max(gt_count[img], 0), gmax);
    int xchg = 0;
    for (int i = threadIdx.x; i < (chunk + 31) / 32; i += kThreads) taken[i] = 0u;
    if (rank == 0)
        for (int g = G + threadIdx.x; g < gmax; g += kThreads) {
            out_idx[(size_t)img * gmax + g] = -1;
            out_deltas[(size_t)img * gmax + g] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    __syncthreads();

    __shared__ Best s_wbest[kRoundBoxes][kThreads / 32];
    __shared__ Best s_cand2[2][kRoundBoxes][kMaxCluster];
    __shared__ Best s_final[2][kRoundBoxes];
    __shared__ int s_nacc[2];
    const bool batched = !no_batch;
    int g_next = 0;      // boxes [0, g_next) are settled
    int force_seq = 0;   // the next box must take the sequential path (distance fallback)
    while (g_next < G) {
        if (batched && !force_seq) {
            // ---- one batched round over boxes [g_next, g_next + nb) ----
            const int nb = min(kRoundBoxes, G - g_next);
            const int slot = xchg & 1;
            ++xchg;
            constexpr int kMaxPer = 5;                      // anchors per thread kept in registers per pass (KITTI: 2112 per CTA)
            for (int a0 = a_begin; a0 < a_end || a0 == a_begin; a0 += kThreads * kMaxPer) {
                double ax1[kMaxPer], ay1[kMaxPer], ax2[kMaxPer], ay2[kMaxPer], aar[kMaxPer];
                bool live[kMaxPer];
#pragma unroll
                for (int k = 0; k < kMaxPer; ++k) {
                    const int a = a0 + k * kThreads + threadIdx.x;
                    live[k] = a < a_end && !((taken[(a - a_begin) >> 5] >> ((a - a_begin) & 31)) & 1u);
                    if (live[k]) {
                        const double2 xy = *reinterpret_cast<const double2 *>(anchors + (size_t)a * 4);
                        const double2 wh = *reinterpret_cast<const double2 *>(anchors + (size_t)a * 4 + 2);
                        const double hw = d_mul(0.5, d_sub(wh.x, 1.0)), hh = d_mul(0.5, d_sub(wh.y, 1.0));
                        ax1[k] = d_sub(xy.x, hw); ay1[k] = d_sub(xy.y, hh); ax2[k] = d_add(xy.x, hw); ay2[k] = d_add(xy.y, hh);
                        aar[k] = d_mul(d_sub(ax2[k], ax1[k]), d_sub(ay2[k], ay1[k]));
                    }
                }
                for (int j = 0; j < nb; ++j) {
                    const float4 b = gt_boxes[(size_t)img * gmax + g_next + j];
                    const double bx1 = b.x, by1 = b.y, bx2 = b.z, by2 = b.w;
                    const double area_g = (double)fmul(fsub(b.z, b.x), fsub(b.w, b.y));
                    Best best{0.0, 0x7fffffff};
#pragma unroll
                    for (int k = 0; k < kMaxPer; ++k) {
                        if (!live[k]) continue;
                        const double lr = fmax(d_sub(fmin(ax2[k], bx2), fmax(ax1[k], bx1)), 0.0);
                        const double tb = fmax(d_sub(fmin(ay2[k], by2), fmax(ay1[k], by1)), 0.0);
                        const double inter = d_mul(lr, tb);
                        const double uni = d_sub(d_add(aar[k], area_g), inter);
                        const double iou = d_div(inter, d_add(uni, 1e-10));
                        if (iou > 0.0 && (best.idx == 0x7fffffff || iou > best.v)) {   // a ascends with k: first max kept
                            best.v = iou;
                            best.idx = a0 + k * kThreads + threadIdx.x;
                        }
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        Best y;
                        y.v = __shfl_down_sync(0xffffffffu, best.v, off);
                        y.idx = __shfl_down_sync(0xffffffffu, best.idx, off);
                        best = better_max(best, y);
                    }
                    if ((threadIdx.x & 31) == 0) {
                        Best &dst = s_wbest[j][threadIdx.x >> 5];
                        dst = a0 == a_begin ? best : better_max(dst, best);   // several passes when a CTA owns > 4096 anchors
                    }
                }
                if (a0 + kThreads * kMaxPer >= a_end) break;
            }
            __syncthreads();
            if ((int)threadIdx.x < nb) {   // block result of box j -> rank 0
                Best w = s_wbest[threadIdx.x][0];
                for (int k = 1; k < kThreads / 32; ++k) w = better_max(w, s_wbest[threadIdx.x][k]);
                if (cs > 1) cg::this_cluster().map_shared_rank(&s_cand2[slot][threadIdx.x][0], 0)[rank] = w;
                else s_cand2[slot][threadIdx.x][0] = w;
            }
            if (cs > 1) cg::this_cluster().sync(); else __syncthreads();
            if (rank == 0) {
                if ((int)threadIdx.x < nb) {
                    Best w = s_cand2[slot][threadIdx.x][0];
                    for (int r = 1; r < cs; ++r) w = better_max(w, s_cand2[slot][threadIdx.x][r]);
                    s_final[slot][threadIdx.x] = w;
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    // accept in annotation order up to the first box that has no candidate or collides with an accepted one
                    int n_acc = 0;
                    for (; n_acc < nb; ++n_acc) {
                        const int c = s_final[slot][n_acc].idx;
                        bool clash = c == 0x7fffffff;
                        for (int q = 0; q < n_acc && !clash; ++q) clash = s_final[slot][q].idx == c;
                        if (clash) break;
                    }
                    for (int r = 0; r < cs; ++r) {
                        if (cs > 1) {
                            Best *rf = cg::this_cluster().map_shared_rank(&s_final[slot][0], r);
                            if (r) for (int q = 0; q < n_acc; ++q) rf[q] = s_final[slot][q];
                            *cg::this_cluster().map_shared_rank(&s_nacc[slot], r) = n_acc;
                        } else {
                            s_nacc[slot] = n_acc;
                        }
                    }
                }
            }
            if (cs > 1) cg::this_cluster().sync(); else __syncthreads();
            const int n_acc = s_nacc[slot];
            for (int q = threadIdx.x; q < n_acc; q += kThreads) {
                const int j = s_final[slot][q].idx;
                if (j >= a_begin && j < a_end) atomicOr(&taken[(j - a_begin) >> 5], 1u << ((j - a_begin) & 31));
                if (rank == 0) {
                    const float4 b = gt_boxes[(size_t)img * gmax + g_next + q];
                    const float gx = fdiv(fadd(b.x, b.z), 2.0f), gy = fdiv(fadd(b.y, b.w), 2.0f);
                    const float gw = fadd(fsub(b.z, b.x), 1.0f), gh = fadd(fsub(b.w, b.y), 1.0f);
                    const double ax = anchors[(size_t)j * 4], ay = anchors[(size_t)j * 4 + 1];
                    const double aw = anchors[(size_t)j * 4 + 2], ah = anchors[(size_t)j * 4 + 3];
                    float4 d;
                    d.x = (float)d_div(d_sub((double)gx, ax), aw);   // boxes.py:125-128, float64 then cast
                    d.y = (float)d_div(d_sub((double)gy, ay), ah);
                    d.z = (float)log(d_div((double)gw, aw));
                    d.w = (float)log(d_div((double)gh, ah));
                    out_idx[(size_t)img * gmax + g_next + q] = j;
                    out_deltas[(size_t)img * gmax + g_next + q] = d;
                }
            }
            __syncthreads();
            g_next += n_acc;
            // the box that stopped the round: no overlapping untaken anchor -> distance fallback (sequential path);
            // a clash -> simply recomputed by the next round
            if (n_acc < nb && s_final[slot][n_acc].idx == 0x7fffffff && rank == 0) force_seq = 1;
            if (cs > 1) {   // make the verdict cluster-uniform (only rank 0 holds the un-accepted candidates)
                __shared__ int s_force[2];
                if (rank == 0 && threadIdx.x == 0)
                    for (int r = 0; r < cs; ++r) *cg::this_cluster().map_shared_rank(&s_force[slot], r) = force_seq;
                cg::this_cluster().sync();
                force_seq = s_force[slot];
            }
            continue;
        }
        force_seq = 0;
        const int g = g_next++;
        const float4 b = gt_boxes[(size_t)img * gmax + g];
        // float32 scalar arithmetic of boxes.py:17-22 (xyxy_to_xywh on the float32 GT array)
        const float gx = fdiv(fadd(b.x, b.z), 2.0f);
        const float gy = fdiv(fadd(b.y, b.w), 2.0f);
        const float gw = fadd(fsub(b.z, b.x), 1.0f);
        const float gh = fadd(fsub(b.w, b.y), 1.0f);
        const double bx1 = b.x, by1 = b.y, bx2 = b.z, by2 = b.w;
        const double area_g = (double)fmul(fsub(b.z, b.x), fsub(b.w, b.y));  // float32 product, then promoted

        // pass 1: best IoU > 0 among untaken anchors (boxes.py:70-81,104-111)
        Best best{0.0, 0x7fffffff};
        for (int a = a_begin + threadIdx.x; a < a_end; a += kThreads) {
            const int la = a - a_begin;
            if ((taken[la >> 5] >> (la & 31)) & 1u) continue;
            const double2 xy = *reinterpret_cast<const double2 *>(anchors + (size_t)a * 4);
            const double2 wh = *reinterpret_cast<const double2 *>(anchors + (size_t)a * 4 + 2);
            const double hw = d_mul(0.5, d_sub(wh.x, 1.0)), hh = d_mul(0.5, d_sub(wh.y, 1.0));
            const double x1 = d_sub(xy.x, hw), y1 = d_sub(xy.y, hh), x2 = d_add(xy.x, hw), y2 = d_add(xy.y, hh);
            const double lr = fmax(d_sub(fmin(x2, bx2), fmax(x1, bx1)), 0.0);
            const double tb = fmax(d_sub(fmin(y2, by2), fmax(y1, by1)), 0.0);
            const double inter = d_mul(lr, tb);
            const double area_a = d_mul(d_sub(x2, x1), d_sub(y2, y1));
            const double uni = d_sub(d_add(area_a, area_g), inter);
            const double iou = d_div(inter, d_add(uni, 1e-10));
            if (iou > 0.0 && (best.idx == 0x7fffffff || iou > best.v)) {  // a ascends per thread: first max kept
                best.v = iou;
                best.idx = a;
            }
        }
        best = block_reduce<true>(best, scratch);
        best = cluster_reduce<true>(best, cs, rank, xchg, s_cand, s_win);

        if (best.idx == 0x7fffffff) {   // cluster-uniform
            // pass 2: nearest untaken anchor in squared xywh distance (boxes.py:115-121)
            Best nb{0.0, 0x7fffffff};
            for (int a = a_begin + threadIdx.x; a < a_end; a += kThreads) {
                const int la = a - a_begin;
                if ((taken[la >> 5] >> (la & 31)) & 1u) continue;
                const double2 xy = *reinterpret_cast<const double2 *>(anchors + (size_t)a * 4);
                const double2 wh = *reinterpret_cast<const double2 *>(anchors + (size_t)a * 4 + 2);
                const double d0 = d_sub((double)gx, xy.x), d1 = d_sub((double)gy, xy.y);
                const double d2 = d_sub((double)gw, wh.x), d3 = d_sub((double)gh, wh.y);
                const double dist = d_add(d_add(d_add(d_mul(d0, d0), d_mul(d1, d1)), d_mul(d2, d2)), d_mul(d3, d3));
                if (nb.idx == 0x7fffffff || dist < nb.v) {
                    nb.v = dist;
                    nb.idx = a;
                }
            }
            nb = block_reduce<false>(nb, scratch);
            best = cluster_reduce<false>(nb, cs, rank, xchg, s_cand, s_win);
        }

        if (threadIdx.x == 0) {
            int j = best.idx;
            if (j != 0x7fffffff && j >= a_begin && j < a_end) taken[(j - a_begin) >> 5] |= 1u << ((j - a_begin) & 31);
            if (rank == 0) {
                float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j != 0x7fffffff) {
                    const double ax = anchors[(size_t)j * 4], ay = anchors[(size_t)j * 4 + 1];
                    const double aw = anchors[(size_t)j * 4 + 2], ah = anchors[(size_t)j * 4 + 3];
                    d.x = (float)d_div(d_sub((double)gx, ax), aw);   // boxes.py:125-128, float64 then cast
                    d.y = (float)d_div(d_sub((double)gy, ay), ah);
                    d.z = (float)log(d_div((double)gw, aw));
                    d.w = (float)log(d_div((double)gh, ah));
                } else {
                    j = A;  // more GT boxes than anchors: the reference leaves anchor_idx == num_anchors
                }
                out_idx[(size_t)img * gmax + g] = j;
                out_deltas[(size_t)img * gmax + g] = d;
            }
        }
        __syncthreads();  // taken[] update visible before the next GT box
    }
    if (cs > 1) cg::this_cluster().sync();   // nobody exits while a peer may still write its exchange slots
}

// scatter the matched rows into the (already zeroed) dense target, base.py:69-74
__global__ void scatter_targets_kernel(const float4 *gt_boxes, const int *gt_classes, const int *gt_count,
                                       const int *anchor_idx, const float4 *deltas, int gmax, int A, int C,
                                       float *gt_dense) {
    const int img = blockIdx.x;
    const int G = min(max(gt_count[img], 0), gmax);
    const int W = C + 9;
    // later rows win where the reference's fancy-index assignment would (duplicates cannot occur for idx < A)
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const int j = anchor_idx[(size_t)img * gmax + g];
        if (j < 0 || j >= A) continue;
        float *row = gt_dense + ((size_t)img * A + j) * W;
        const float4 b = gt_boxes[(size_t)img * gmax + g];
        const float4 d = deltas[(size_t)img * gmax + g];
        row[0] = 1.f;
        row[1] = b.x; row[2] = b.y; row[3] = b.z; row[4] = b.w;
        row[5] = d.x; row[6] = d.y; row[7] = d.z; row[8] = d.w;
        const int c = gt_classes[(size_t)img * gmax + g];
        if (c >= 0 && c < C) row[9 + c] = 1.f;
    }
}

}  // namespace

extern "C" int sqd_match_anchors(const float *d_gt_boxes, const int32_t *d_gt_count, int batch, int gmax,
                                 const double *d_anchors64, int num_anchors, int32_t *d_anchor_idx,
                                 float *d_deltas, void *stream) {
    if (batch == 0) return SQD_OK;  // empty batch: nothing to enqueue, pointers may be NULL
    SQD_REQUIRE(d_gt_boxes && d_gt_count && d_anchors64 && d_anchor_idx && d_deltas, SQD_E_NULL,
                "sqd_match_anchors: NULL pointer");
    SQD_REQUIRE(batch >= 0 && gmax >= 1 && gmax <= SQD_MAX_GT, SQD_E_SHAPE, "sqd_match_anchors: gmax %d outside [1,%d]",
                gmax, SQD_MAX_GT);
    SQD_REQUIRE(num_anchors > 0 && num_anchors <= (1 << 20), SQD_E_SHAPE, "sqd_match_anchors: bad num_anchors %d",
                num_anchors);
    SQD_REQUIRE(sqd_aligned16(d_gt_boxes) && sqd_aligned16(d_anchors64) && sqd_aligned16(d_deltas), SQD_E_ALIGN,
                "sqd_match_anchors: gt_boxes/anchors/deltas must be 16-byte aligned");
    // CTAs per image: enough to bring a GT box down to a few anchors per thread, not more than fills the GPU
    int cs = kMaxCluster;
    while (cs > 1 && (num_anchors / cs < 2 * kThreads || (long long)batch * cs > 4ll * SQD_SM_COUNT * 2)) cs >>= 1;
    const int chunk = ((num_anchors + cs - 1) / cs + 31) & ~31;
    const size_t smem = (size_t)((chunk + 31) / 32) * sizeof(unsigned);
    if (smem > 48 * 1024)
        SQD_CUDA(cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(batch * cs));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, match_kernel, reinterpret_cast<const float4 *>(d_gt_boxes), d_gt_count, gmax,
                                       d_anchors64, num_anchors, d_anchor_idx, reinterpret_cast<float4 *>(d_deltas), cs,
                                       sqd_opt(SQD_OPT_MATCH_SEQUENTIAL) ? 1 : 0);
    if (e != cudaSuccess) {
        sqd_set_error("launch of match_kernel failed: %s", cudaGetErrorString(e));
        return (int)e;
    }
    SQD_LAUNCH_CHECK("match_kernel");
    return SQD_OK;
}

extern "C" int sqd_build_targets(const float *d_gt_boxes, const int32_t *d_gt_classes, const int32_t *d_gt_count,
                                 const int32_t *d_anchor_idx, const float *d_deltas, int batch, int gmax,
                                 int num_anchors, int num_classes, float *d_gt_dense, void *stream) {
    if (batch == 0) return SQD_OK;  // empty batch: nothing to enqueue, pointers may be NULL
    SQD_REQUIRE(d_gt_boxes && d_gt_classes && d_gt_count && d_anchor_idx && d_deltas && d_gt_dense, SQD_E_NULL,
                "sqd_build_targets: NULL pointer");
    SQD_REQUIRE(batch >= 0 && gmax >= 1 && gmax <= SQD_MAX_GT && num_anchors > 0, SQD_E_SHAPE,
                "sqd_build_targets: bad shape");
    SQD_REQUIRE(num_classes >= 1 && num_classes <= SQD_MAX_CLASSES, SQD_E_SHAPE, "sqd_build_targets: bad num_classes");
    SQD_REQUIRE(sqd_aligned16(d_gt_boxes) && sqd_aligned16(d_deltas), SQD_E_ALIGN,
                "sqd_build_targets: gt_boxes/deltas must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SQD_CUDA(cudaMemsetAsync(d_gt_dense, 0, (size_t)batch * num_anchors * (num_classes + 9) * sizeof(float), st));
    scatter_targets_kernel<<<batch, 64, 0, st>>>(reinterpret_cast<const float4 *>(d_gt_boxes), d_gt_classes,
                                                 d_gt_count, d_anchor_idx, reinterpret_cast<const float4 *>(d_deltas),
                                                 gmax, num_anchors, num_classes, d_gt_dense);
    SQD_LAUNCH_CHECK("scatter_targets_kernel");
    return SQD_OK;
}

// a14-a16: SqueezeDet loss forward + analytic backward in three small launches.
// Reference: Loss.forward, src/model/squeezedet.py:133-174 (PredictionResolver with log_softmax,
// torch compute_overlaps modules.py:48-63, four masked sums; ~40 ATen kernels + the autograd graph).
//
//   count   : per (image, slice) partial sums of the anchor mask            -> n = num_objects
//   main    : per anchor, all four loss terms and d(loss)/d(pred); per-slice partial sums
//   finalize: fixed-order sum of the slice partials, the "/ n" and "/ (A - n)" normalisations
// Every reduction has a fixed order, so results are run-to-run deterministic.
//
// Backward (oracle/oracle.py: loss_backward states the same formulas):
//   d/dz_k   = w_cls * m/n * (sum_c(y_c) * p_k - y_k)
//   d/ds     = (w_pos*m/n + w_neg*(1-m)/(A-n)) * 2*(iou - sig) * (-sig*(1-sig))
//   d/ddelta = w_box*m/n*2*(delta - t)  +  [IoU target is NOT detached in the reference:]
//              dL/dIoU * dIoU/dbox * dbox/ddelta, with torch's clamp pass-through mask
//              (gradient only where the raw coordinate lies inside [0, W-1] / [0, H-1], modules.py:42-43)
//              and torch's min/max tie rule (gradient halved on exact ties).
// A zero-object image gives 0/0 = NaN losses and gradients, as in the reference (not "fixed").
// Bytes per image: count reads the mask column (A*(C+9)*4 at sector granularity), main reads
// pred + gt once and writes dpred: A*((C+5)*2 + (C+9))*4.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxSlices = 64;

__device__ __forceinline__ float block_sum(float v, float *scratch) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = 0.f;
    if (warp == 0) {
        r = lane < (kThreads >> 5) ? scratch[lane] : 0.f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) r += __shfl_down_sync(0xffffffffu, r, off);
    }
    return r;  // valid in thread 0
}

__global__ void __launch_bounds__(kThreads) loss_count_kernel(const float *gt, int A, int W, int per_slice,
                                                              float *partial_n) {
    __shared__ float scratch[kThreads / 32];
    const int img = blockIdx.y, s = blockIdx.x, S = gridDim.x;
    const int a0 = s * per_slice, a1 = min(A, a0 + per_slice);
    const float *g = gt + (size_t)img * A * W;
    float acc = 0.f;
    for (int a = a0 + threadIdx.x; a < a1; a += kThreads) acc += __ldg(g + (size_t)a * W);
    const float tot = block_sum(acc, scratch);
    if (threadIdx.x == 0) partial_n[img * S + s] = tot;
}

struct LossArgs {
    const float *pred;
    const float *gt;
    const float4 *anchors;
    int A, C;
    float wmax, hmax;
    float w_cls, w_pos, w_neg, w_box;
    const float *grad_loss;  // (B,4) or null
    float *dpred;            // or null
    const float *partial_n;
    float *partial_loss;  // (B, S, 4)
    int per_slice;
};

// gradient share of `a` in min(a,b): 1, 1/2 on an exact tie, 0 (torch's minimum/maximum backward)
__device__ __forceinline__ float share_min(float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); }

template <int CS>
__global__ void __launch_bounds__(kThreads) loss_main_kernel(LossArgs p) {
    __shared__ float scratch[kThreads / 32];
    __shared__ float s_n;
    const int C = CS > 0 ? CS : p.C;
    const int NF = C + 5, W = C + 9;
    const int img = blockIdx.y, s = blockIdx.x, S = gridDim.x;
    if (threadIdx.x == 0) {
        float n = 0.f;
        for (int i = 0; i < S; ++i) n += p.partial_n[img * S + i];  // fixed order
        s_n = n;
    }
    __syncthreads();
    const float n = s_n;
    const float fA = (float)p.A;
    const bool want_grad = p.dpred != nullptr;
    // upstream gradient of the four per-image terms {class, positive score, negative score, bbox}
    float go_cls = 1.f, go_pos = 1.f, go_neg = 1.f, go_box = 1.f;
    if (want_grad && p.grad_loss) {
        go_cls = __ldg(p.grad_loss + img * 4 + 0);
        go_pos = __ldg(p.grad_loss + img * 4 + 1);
        go_neg = __ldg(p.grad_loss + img * 4 + 2);
        go_box = __ldg(p.grad_loss + img * 4 + 3);
    }

    const int a0 = s * p.per_slice, a1 = min(p.A, a0 + p.per_slice);
    float acc_cls = 0.f, acc_pos = 0.f, acc_neg = 0.f, acc_box = 0.f;

    for (int a = a0 + threadIdx.x; a < a1; a += kThreads) {
        const size_t row = (size_t)img * p.A + a;
        float f[SQD_CMAX(CS) + 5], g[SQD_CMAX(CS) + 9];
        const float *pr = p.pred + row * NF;
        const float *gr = p.gt + row * W;
        if (CS == 3) {
            const float4 *p4 = reinterpret_cast<const float4 *>(pr);
            const float4 *g4 = reinterpret_cast<const float4 *>(gr);
            const float4 v0 = ld_stream_f4(p4), v1 = ld_stream_f4(p4 + 1);
            const float4 u0 = ld_stream_f4(g4), u1 = ld_stream_f4(g4 + 1), u2 = ld_stream_f4(g4 + 2);
            f[0] = v0.x; f[1] = v0.y; f[2] = v0.z; f[3] = v0.w; f[4] = v1.x; f[5] = v1.y; f[6] = v1.z; f[7] = v1.w;
            g[0] = u0.x; g[1] = u0.y; g[2] = u0.z; g[3] = u0.w; g[4] = u1.x; g[5] = u1.y; g[6] = u1.z; g[7] = u1.w;
            g[8] = u2.x; g[9] = u2.y; g[10] = u2.z; g[11] = u2.w;
        } else {
#pragma unroll
            for (int j = 0; j < SQD_CMAX(CS) + 5; ++j)
                if (j < NF) f[j] = __ldg(pr + j);
#pragma unroll
            for (int j = 0; j < SQD_CMAX(CS) + 9; ++j)
                if (j < W) g[j] = __ldg(gr + j);
        }
        const float m = g[0];
        float prob[SQD_CMAX(CS)];
        float zmax, sum;
        const float sig = sqd_softmax_conf<CS>(f, C, prob, &zmax, &sum);
        const float lse = logf(sum);

        // class term: sum_c w*m*y_c*(-logp_c)
        float cls = 0.f, ysum = 0.f;
#pragma unroll
        for (int c = 0; c < SQD_CMAX(CS); ++c)
            if (c < C) {
                const float y = g[9 + c];
                const float logp = (f[c] - zmax) - lse;
                cls += p.w_cls * m * y * (-logp);
                ysum += y;
            }
        acc_cls += cls;

        // decoded box (clamped) and the raw coordinates for the clamp pass-through mask
        const float4 anc = __ldg(p.anchors + a);
        const float dx = f[C + 1], dy = f[C + 2], dw = f[C + 3], dh = f[C + 4];
        const float cx = fadd(anc.x, fmul(anc.z, dx)), cy = fadd(anc.y, fmul(anc.w, dy));
        const float bw = fmul(anc.z, expf(dw)), bh = fmul(anc.w, expf(dh));
        const float hw = fmul(0.5f, fsub(bw, 1.f)), hh = fmul(0.5f, fsub(bh, 1.f));
        const float r0 = fsub(cx, hw), r1 = fsub(cy, hh), r2 = fadd(cx, hw), r3 = fadd(cy, hh);
        const float p0 = sqd_clamp(r0, p.wmax), p1 = sqd_clamp(r1, p.hmax);
        const float p2 = sqd_clamp(r2, p.wmax), p3 = sqd_clamp(r3, p.hmax);

        // IoU(gt box, predicted box), modules.py:48-63
        const float lr_raw = fsub(fminf(g[3], p2), fmaxf(g[1], p0));
        const float tb_raw = fsub(fminf(g[4], p3), fmaxf(g[2], p1));
        const float lr = fmaxf(lr_raw, 0.f), tb = fmaxf(tb_raw, 0.f);
        const float inter = fmul(lr, tb);
        const float pw = fsub(p2, p0), ph = fsub(p3, p1);
        const float uni = fsub(fadd(fmul(fsub(g[3], g[1]), fsub(g[4], g[2])), fmul(pw, ph)), inter);
        const float den = fadd(uni, 1e-10f);
        const float iou = fmul(fdiv(inter, den), m);
        const float resid = iou - sig;
        const float r2sq = resid * resid;
        acc_pos += p.w_pos * m * r2sq;
        acc_neg += p.w_neg * (1.f - m) * r2sq;

        float box = 0.f;
        const float e0 = dx - g[5], e1 = dy - g[6], e2 = dw - g[7], e3 = dh - g[8];
        box = p.w_box * m * (e0 * e0) + p.w_box * m * (e1 * e1) + p.w_box * m * (e2 * e2) + p.w_box * m * (e3 * e3);
        acc_box += box;

        if (want_grad) {
            const float k_obj = m / n;                       // NaN when n == 0, like autograd's 0/0
            const float k_bg = (1.f - m) / (fA - n);
            float d[SQD_CMAX(CS) + 5];
#pragma unroll
            for (int c = 0; c < SQD_CMAX(CS); ++c)
                if (c < C) d[c] = go_cls * p.w_cls * k_obj * (ysum * prob[c] - g[9 + c]);
            const float k_sc = go_pos * p.w_pos * k_obj + go_neg * p.w_neg * k_bg;
            d[C] = k_sc * 2.f * resid * (-sig * (1.f - sig));

            const float dL_diou = k_sc * 2.f * resid * m;
            const float inv_den2 = 1.f / (den * den);
            const float d_inter = (den + inter) * inv_den2;
            const float d_area = -inter * inv_den2;
            const float on_lr = lr_raw >= 0.f ? 1.f : 0.f, on_tb = tb_raw >= 0.f ? 1.f : 0.f;
            // x1/y1 enter through max(g, p): share of p is that of -p in min(-p, -g)
            const float dI0 = -tb * on_lr * share_min(-p0, -g[1]);
            const float dI1 = -lr * on_tb * share_min(-p1, -g[2]);
            const float dI2 = tb * on_lr * share_min(p2, g[3]);
            const float dI3 = lr * on_tb * share_min(p3, g[4]);
            const float pass0 = (r0 >= 0.f && r0 <= p.wmax) ? 1.f : 0.f, pass1 = (r1 >= 0.f && r1 <= p.hmax) ? 1.f : 0.f;
            const float pass2 = (r2 >= 0.f && r2 <= p.wmax) ? 1.f : 0.f, pass3 = (r3 >= 0.f && r3 <= p.hmax) ? 1.f : 0.f;
            const float G0 = dL_diou * (d_inter * dI0 + d_area * (-ph)) * pass0;
            const float G1 = dL_diou * (d_inter * dI1 + d_area * (-pw)) * pass1;
            const float G2 = dL_diou * (d_inter * dI2 + d_area * ph) * pass2;
            const float G3 = dL_diou * (d_inter * dI3 + d_area * pw) * pass3;
            const float kb = go_box * p.w_box * k_obj * 2.f;
            d[C + 1] = kb * e0 + anc.z * (G0 + G2);
            d[C + 2] = kb * e1 + anc.w * (G1 + G3);
            d[C + 3] = kb * e2 + 0.5f * bw * (G2 - G0);
            d[C + 4] = kb * e3 + 0.5f * bh * (G3 - G1);
            float *out = p.dpred + row * NF;
            if (CS == 3) {
                float4 *o4 = reinterpret_cast<float4 *>(out);
                o4[0] = make_float4(d[0], d[1], d[2], d[3]);
                o4[1] = make_float4(d[4], d[5], d[6], d[7]);
            } else {
#pragma unroll
                for (int j = 0; j < SQD_CMAX(CS) + 5; ++j)
                    if (j < NF) out[j] = d[j];
            }
        }
    }
    float *pl = p.partial_loss + ((size_t)img * S + s) * 4;
    float t;
    t = block_sum(acc_cls, scratch); if (threadIdx.x == 0) pl[0] = t;
    t = block_sum(acc_pos, scratch); if (threadIdx.x == 0) pl[1] = t;
    t = block_sum(acc_neg, scratch); if (threadIdx.x == 0) pl[2] = t;
    t = block_sum(acc_box, scratch); if (threadIdx.x == 0) pl[3] = t;
}

__global__ void loss_finalize_kernel(const float *partial_n, const float *partial_loss, int S, int A, float *losses) {
    const int img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= gridDim.x * blockDim.x) return;
    float n = 0.f, c = 0.f, ps = 0.f, ng = 0.f, bx = 0.f;
    for (int s = 0; s < S; ++s) {
        n += partial_n[img * S + s];
        const float *pl = partial_loss + ((size_t)img * S + s) * 4;
        c += pl[0]; ps += pl[1]; ng += pl[2]; bx += pl[3];
    }
    losses[img * 4 + 0] = c / n;
    losses[img * 4 + 1] = ps / n;
    losses[img * 4 + 2] = ng / ((float)A - n);
    losses[img * 4 + 3] = bx / n;
}

int pick_slices(int batch) {
    int s = (2 * SQD_SM_COUNT + batch - 1) / batch;
    if (s < 1) s = 1;
    if (s > kMaxSlices) s = kMaxSlices;
    return s;
}

}  // namespace

extern "C" size_t sqd_loss_workspace_bytes(int batch, int num_anchors) {
    (void)num_anchors;
    if (batch <= 0) return 256;
    return (size_t)batch * kMaxSlices * 5 * sizeof(float) + 256;
}

extern "C" int sqd_loss_fwd_bwd(const float *d_pred, const float *d_gt, const float *d_anchors, int batch,
                                int num_anchors, int num_classes, int input_h, int input_w, const float *weights,
                                const float *d_grad_loss, float *d_losses, float *d_dpred, void *d_workspace,
                                size_t workspace_bytes, void *stream) {
    if (batch == 0) return SQD_OK;  // empty batch: nothing to enqueue, pointers may be NULL
    SQD_REQUIRE(d_pred && d_gt && d_anchors && weights && d_losses && d_workspace, SQD_E_NULL,
                "sqd_loss_fwd_bwd: NULL pointer");
    SQD_REQUIRE(batch >= 0 && num_anchors > 0, SQD_E_SHAPE, "sqd_loss_fwd_bwd: bad shape");
    SQD_REQUIRE(num_classes >= 1 && num_classes <= SQD_MAX_CLASSES, SQD_E_SHAPE, "sqd_loss_fwd_bwd: bad num_classes");
    SQD_REQUIRE(workspace_bytes >= sqd_loss_workspace_bytes(batch, num_anchors), SQD_E_WORKSPACE,
                "sqd_loss_fwd_bwd: workspace too small");
    SQD_REQUIRE(sqd_aligned16(d_pred) && sqd_aligned16(d_gt) && sqd_aligned16(d_anchors) && sqd_aligned16(d_dpred) &&
                    sqd_aligned16(d_workspace),
                SQD_E_ALIGN, "sqd_loss_fwd_bwd: pointers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int S = pick_slices(batch);
    int per_slice = (num_anchors + S - 1) / S;
    per_slice = ((per_slice + kThreads - 1) / kThreads) * kThreads;
    float *partial_n = static_cast<float *>(d_workspace);
    float *partial_loss = partial_n + (size_t)batch * kMaxSlices;
    dim3 grid(S, batch);
    loss_count_kernel<<<grid, kThreads, 0, st>>>(d_gt, num_anchors, num_classes + 9, per_slice, partial_n);
    SQD_LAUNCH_CHECK("loss_count_kernel");
    LossArgs p;
    p.pred = d_pred;
    p.gt = d_gt;
    p.anchors = reinterpret_cast<const float4 *>(d_anchors);
    p.A = num_anchors;
    p.C = num_classes;
    p.wmax = (float)(input_w - 1);
    p.hmax = (float)(input_h - 1);
    p.w_cls = weights[0];
    p.w_pos = weights[1];
    p.w_neg = weights[2];
    p.w_box = weights[3];
    p.grad_loss = d_grad_loss;
    p.dpred = d_dpred;
    p.partial_n = partial_n;
    p.partial_loss = partial_loss;
    p.per_slice = per_slice;
    if (num_classes == 3)
        loss_main_kernel<3><<<grid, kThreads, 0, st>>>(p);
    else if (num_classes == 8)
        loss_main_kernel<8><<<grid, kThreads, 0, st>>>(p);
    else
        loss_main_kernel<0><<<grid, kThreads, 0, st>>>(p);
    SQD_LAUNCH_CHECK("loss_main_kernel");
    loss_finalize_kernel<<<batch, 1, 0, st>>>(partial_n, partial_loss, S, num_anchors, d_losses);
    SQD_LAUNCH_CHECK("loss_finalize_kernel");
    return SQD_OK;
}

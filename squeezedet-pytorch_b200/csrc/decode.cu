// a2-a7: PredictionResolver.forward + SqueezeDet.forward scoring as one HBM-bound pass over pred.
// Reference: src/model/squeezedet.py:109-120,200-205 and src/model/modules.py:17-45,66-68
// (~25 ATen kernels, a host sync and an H2D anchor copy per call there).
//
// Layout: pred is (B*A, NF=C+5) fp32 with fields interleaved per anchor.  One thread owns one
// anchor; a 256-thread block stages its contiguous 256*NF-float slab through shared memory with
// 16-byte streaming loads (fully coalesced for every NF), then each thread reads its NF fields
// (stride NF words: conflict-free for odd NF, LDS.128 for NF==8).  Algorithmic bytes per anchor:
// NF*4 read + 28 written (int64 id, score, box) for the SqueezeDet.forward contract.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct DecodeArgs {
    const float *pred;
    const float4 *anchors;
    long long total;  // B*A
    int num_anchors;
    int num_classes;
    float wmax, hmax;
    long long *class_ids;
    float *scores;
    float4 *boxes;
    float *probs;
    float *logp;
    float *conf;
    float4 *deltas;
};

template <int CS>
__global__ void __launch_bounds__(kThreads) decode_kernel(DecodeArgs a) {
    extern __shared__ __align__(16) float slab[];
    const int C = CS > 0 ? CS : a.num_classes;
    const int NF = C + 5;
    const long long g0 = (long long)blockIdx.x * kThreads;
    const int n = (int)min((long long)kThreads, a.total - g0);
    const float *src = a.pred + g0 * NF;
    float f[SQD_CMAX(CS) + 5];
    if (CS == 3) {
        // 32-byte rows: a thread reads its own row with two 16-byte loads (a warp covers 1 KB contiguous; the second
        // load of a 32-byte sector hits L1) -- no shared-memory round trip, no block barrier
        if ((int)threadIdx.x >= n) return;
        const float4 *row = reinterpret_cast<const float4 *>(src) + 2 * threadIdx.x;
        const float4 lo = __ldg(row), hi = __ldg(row + 1);
        f[0] = lo.x; f[1] = lo.y; f[2] = lo.z; f[3] = lo.w;
        f[4] = hi.x; f[5] = hi.y; f[6] = hi.z; f[7] = hi.w;
    } else {
        const int count = n * NF;
        {
            const int nvec = count >> 2;  // slab start is 16-byte aligned: kThreads*NF*4 is a multiple of 16
            const float4 *src4 = reinterpret_cast<const float4 *>(src);
            float4 *dst4 = reinterpret_cast<float4 *>(slab);
            for (int i = threadIdx.x; i < nvec; i += kThreads) dst4[i] = ld_stream_f4(src4 + i);
            for (int i = (nvec << 2) + threadIdx.x; i < count; i += kThreads) slab[i] = src[i];
        }
        __syncthreads();
        if ((int)threadIdx.x >= n) return;
#pragma unroll
        for (int j = 0; j < SQD_CMAX(CS) + 5; ++j)
            if (j < NF) f[j] = slab[threadIdx.x * NF + j];
    }

    const long long g = g0 + threadIdx.x;
    const int anchor = (int)(g % a.num_anchors);

    float p[SQD_CMAX(CS)];
    float zmax, sum;
    const float conf = sqd_softmax_conf<CS>(f, C, p, &zmax, &sum);

    if (a.class_ids != nullptr || a.scores != nullptr) {
        float sbest = fmul(p[0], conf);
        int best = 0;
#pragma unroll
        for (int c = 1; c < SQD_CMAX(CS); ++c)
            if (c < C) {
                const float s = fmul(p[c], conf);
                if (s > sbest) {
                    sbest = s;
                    best = c;
                }
            }
        if (a.class_ids) a.class_ids[g] = best;
        if (a.scores) a.scores[g] = sbest;
    }
    if (a.boxes) {
        const float4 anc = __ldg(a.anchors + anchor);
        a.boxes[g] = sqd_decode_box(anc, f[C + 1], f[C + 2], f[C + 3], f[C + 4], a.wmax, a.hmax);
    }
    if (a.probs) {
#pragma unroll
        for (int c = 0; c < SQD_CMAX(CS); ++c)
            if (c < C) a.probs[g * C + c] = p[c];
    }
    if (a.logp) {
        const float lse = logf(sum);
#pragma unroll
        for (int c = 0; c < SQD_CMAX(CS); ++c)
            if (c < C) a.logp[g * C + c] = fsub(fsub(f[c], zmax), lse);
    }
    if (a.conf) a.conf[g] = conf;
    if (a.deltas) a.deltas[g] = make_float4(f[C + 1], f[C + 2], f[C + 3], f[C + 4]);
}

}  // namespace

extern "C" int sqd_decode_scores(const float *d_pred, const float *d_anchors, int batch, int num_anchors,
                                 int num_classes, int input_h, int input_w, int64_t *d_class_ids, float *d_scores,
                                 float *d_boxes, float *d_probs, float *d_logp, float *d_conf, float *d_deltas,
                                 void *stream) {
    if (batch == 0) return SQD_OK;  // empty batch: nothing to enqueue, pointers may be NULL
    SQD_REQUIRE(d_pred && d_anchors, SQD_E_NULL, "sqd_decode_scores: pred/anchors is NULL");
    SQD_REQUIRE(batch >= 0 && num_anchors > 0, SQD_E_SHAPE, "sqd_decode_scores: bad batch/num_anchors");
    SQD_REQUIRE(num_classes >= 1 && num_classes <= SQD_MAX_CLASSES, SQD_E_SHAPE,
                "sqd_decode_scores: num_classes %d outside [1,%d]", num_classes, SQD_MAX_CLASSES);
    SQD_REQUIRE(sqd_aligned16(d_pred) && sqd_aligned16(d_anchors) && sqd_aligned16(d_boxes) && sqd_aligned16(d_deltas),
                SQD_E_ALIGN, "sqd_decode_scores: pred/anchors/boxes/deltas must be 16-byte aligned");
    DecodeArgs a;
    a.pred = d_pred;
    a.anchors = reinterpret_cast<const float4 *>(d_anchors);
    a.total = (long long)batch * num_anchors;
    a.num_anchors = num_anchors;
    a.num_classes = num_classes;
    a.wmax = (float)(input_w - 1);
    a.hmax = (float)(input_h - 1);
    a.class_ids = reinterpret_cast<long long *>(d_class_ids);
    a.scores = d_scores;
    a.boxes = reinterpret_cast<float4 *>(d_boxes);
    a.probs = d_probs;
    a.logp = d_logp;
    a.conf = d_conf;
    a.deltas = reinterpret_cast<float4 *>(d_deltas);
    const long long blocks = (a.total + kThreads - 1) / kThreads;
    SQD_REQUIRE(blocks < 0x7fffffffLL, SQD_E_SHAPE, "sqd_decode_scores: too many anchors");
    const size_t smem = (size_t)kThreads * (num_classes + 5) * sizeof(float);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (num_classes == 3)
        decode_kernel<3><<<(unsigned)blocks, kThreads, 0, st>>>(a);
    else if (num_classes == 8)
        decode_kernel<8><<<(unsigned)blocks, kThreads, smem, st>>>(a);
    else
        decode_kernel<0><<<(unsigned)blocks, kThreads, smem, st>>>(a);
    SQD_LAUNCH_CHECK("decode_kernel");
    return SQD_OK;
}

// Shared device/host helpers for libsqdet_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "sqdet_b200.h"

#define SQD_SM_COUNT 148

// ---- error plumbing (thread-local message, never throws across the ABI) ----------------------
void sqd_set_error(const char *fmt, ...);

#define SQD_REQUIRE(cond, code, ...)   \
    do {                               \
        if (!(cond)) {                 \
            sqd_set_error(__VA_ARGS__); \
            return (code);             \
        }                              \
    } while (0)

#define SQD_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            sqd_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                              \
        }                                                                                \
    } while (0)

#define SQD_LAUNCH_CHECK(name)                                                           \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            sqd_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));      \
            return (int)_e;                                                              \
        }                                                                                \
    } while (0)

// ---- developer options --------------------------------------------------------------------------
// Alternative routes (cross-checked bit for bit by the tests) and tuning knobs.  One table, filled ONCE from the
// environment on first use and changed afterwards only through sqd_set_option(); no entry point calls getenv().
enum SqdOptId {
    SQD_OPT_NO_PDL, SQD_OPT_FUSED_SCORE, SQD_OPT_SPLIT_TWO_PASS, SQD_OPT_SPLIT_CS, SQD_OPT_SPLIT_THREADS, SQD_OPT_SPLIT_ROWS,
    SQD_OPT_DGRAD_PER_SLAB, SQD_OPT_DGRAD_BLOCK_SCALES, SQD_OPT_WG_SINGLE_TAP, SQD_OPT_WG_SYNC, SQD_OPT_BWD_OLD_PREPASS,
    SQD_OPT_MATCH_SEQUENTIAL, SQD_OPT_F16_HALF_TILES, SQD_OPT_F16_CHUNK, SQD_OPT_F16_DBG, SQD_OPT_F16_PAIR_STAGES,
    SQD_OPT_F16_A_STAGES, SQD_OPT_F16_B_STAGES, SQD_OPT_F16_TRACE_CTA, SQD_OPT_HEAD_ONE_KERNEL, SQD_OPT_F16_A_ONCE, SQD_OPT_F16_AO_BUFS, SQD_OPT_TAIL_THREADS, SQD_OPT_DGRAD_PACK_LOOP, SQD_OPT_SPLIT_REGS, SQD_OPT_COUNT
};
int sqd_opt(SqdOptId id);

static inline bool sqd_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- arithmetic that must round exactly like the reference's separate torch / numpy ops -------
// (the _rn intrinsics are never contracted into FMAs by nvcc)
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double d_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double d_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double d_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double d_div(double a, double b) { return __ddiv_rn(a, b); }

// streaming 16-byte load that does not pollute L1 (inputs are read exactly once)
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

// ---- per-anchor decode shared by every kernel that scores anchors ------------------------------
#define SQD_CMAX(CS) ((CS) > 0 ? (CS) : SQD_MAX_CLASSES)

// Class softmax (modules.py:66-68: e=exp(z-max), e/sum with a left-to-right sum) and confidence sigmoid
// (squeezedet.py:114).  Fills e[] with the softmax PROBABILITIES, returns conf; *lse = log(sum) for logp.
template <int CS>
__device__ __forceinline__ float sqd_softmax_conf(const float *f, int C_rt, float *p, float *zmax_out, float *sum_out) {
    const int C = CS > 0 ? CS : C_rt;
    float zmax = f[0];
#pragma unroll
    for (int c = 1; c < SQD_CMAX(CS); ++c)
        if (c < C) zmax = fmaxf(zmax, f[c]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < SQD_CMAX(CS); ++c)
        if (c < C) {
            p[c] = expf(fsub(f[c], zmax));
            sum = (c == 0) ? p[c] : fadd(sum, p[c]);
        }
#pragma unroll
    for (int c = 0; c < SQD_CMAX(CS); ++c)
        if (c < C) p[c] = fdiv(p[c], sum);
    *zmax_out = zmax;
    *sum_out = sum;
    return fdiv(1.0f, fadd(1.0f, expf(-f[C])));
}

// probs *= conf ; argmax (first maximum wins) ; max.   squeezedet.py:200-202
template <int CS>
__device__ __forceinline__ void sqd_score_anchor(const float *f, int C_rt, float &score, int &cls) {
    const int C = CS > 0 ? CS : C_rt;
    float p[SQD_CMAX(CS)];
    float zmax, sum;
    const float conf = sqd_softmax_conf<CS>(f, C_rt, p, &zmax, &sum);
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
    score = sbest;
    cls = best;
}

// Candidate test + score with the work the RESULT needs, bit-identical to sqd_score_anchor wherever it returns true:
//  * score = max_c p_c*conf <= conf (p_c <= 1), so conf <= thr decides "not a candidate" after one exp and one division;
//  * the first-maximum logit has e = exp(0) = 1 exactly and the largest p; unless another class comes within 1e-5 of
//    it (then rounding could tie the products and the reference's first-index rule matters: full path), the score is
//    (1/sum)*conf and the other C-1 divisions and products are never needed.  The sum keeps the reference's order.
template <int CS>
__device__ __forceinline__ bool sqd_score_candidate(const float *f, int C_rt, float score_thr, float &score, int &cls) {
    const int C = CS > 0 ? CS : C_rt;
    const float conf = fdiv(1.0f, fadd(1.0f, expf(-f[C])));
    if (!(conf > score_thr)) {
        if (conf == conf) return false;
        sqd_score_anchor<CS>(f, C_rt, score, cls);   // NaN confidence: rank it exactly like the full path does
        return true;
    }
    float zmax = f[0];
    int imax = 0;
#pragma unroll
    for (int c = 1; c < SQD_CMAX(CS); ++c)
        if (c < C && f[c] > zmax) {
            zmax = f[c];
            imax = c;
        }
    float sum = 0.f;
    bool near = conf < 1e-30f;  // products could underflow into ties
#pragma unroll
    for (int c = 0; c < SQD_CMAX(CS); ++c)
        if (c < C) {
            const float e = c == imax ? 1.0f : expf(fsub(f[c], zmax));
            near = near || (c != imax && !(e <= 0.99999f));   // also catches NaN
            sum = (c == 0) ? e : fadd(sum, e);
        }
    // a non-finite maximum (NaN, or +-inf: the reference's exp(z - max) is then exp(inf - inf) = NaN for the maximum
    // itself) takes the full path, which reproduces torch's NaN propagation
    if (near || !(fabsf(zmax) <= 3.4028235e38f)) {
        sqd_score_anchor<CS>(f, C_rt, score, cls);
        return true;
    }
    score = fmul(fdiv(1.0f, sum), conf);
    cls = imax;
    return true;
}

__device__ __forceinline__ float sqd_clamp(float v, float hi) {  // torch.clamp: NaN passes through
    return v < 0.f ? 0.f : (v > hi ? hi : v);
}

// Anchor delta decoding.  modules.py:27-45 with xywh_to_xyxy of modules.py:17-24 and the clamps.
__device__ __forceinline__ float4 sqd_decode_box(float4 anc, float dx, float dy, float dw, float dh, float wmax,
                                                 float hmax) {
    const float cx = fadd(anc.x, fmul(anc.z, dx));
    const float cy = fadd(anc.y, fmul(anc.w, dy));
    const float w = fmul(anc.z, expf(dw));
    const float h = fmul(anc.w, expf(dh));
    const float hw = fmul(0.5f, fsub(w, 1.0f));
    const float hh = fmul(0.5f, fsub(h, 1.0f));
    float4 b;
    b.x = sqd_clamp(fsub(cx, hw), wmax);
    b.y = sqd_clamp(fsub(cy, hh), hmax);
    b.z = sqd_clamp(fadd(cx, hw), wmax);
    b.w = sqd_clamp(fadd(cy, hh), hmax);
    return b;
}

// ---- ranking keys and per-image candidate lists (shared by the ConvDet epilogue and the filter kernels) ----
// A candidate is a 64-bit key  [ order-preserving score bits : 32 | 0xFFFFFF - anchor : 24 | class : 8 ]
// so "larger key" == (score desc, anchor index asc) -- the declared tie policy.
typedef unsigned long long sqd_u64;

// NaN (either sign) maps to the largest key: torch's descending sort ranks NaN before +inf (detector.py:88)
__device__ __forceinline__ unsigned sqd_order_bits(float s) {
    const unsigned b = __float_as_uint(s);
    if (s != s) return 0xFFFFFFFFu;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float sqd_unorder_bits(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ sqd_u64 sqd_make_key(float score, int anchor, int cls) {
    return ((sqd_u64)sqd_order_bits(score) << 32) | ((sqd_u64)(0xFFFFFFu - (unsigned)anchor) << 8) | (sqd_u64)(cls & 0xFF);
}
// Exact pre-filter: an anchor with score <= score_thresh can never be emitted (final strict filter) and can never
// suppress an emitted box (greedy NMS only lets HIGHER scores suppress), so it only ever occupies a top-k slot that
// no surviving anchor needs.  Keys above "largest key with score == score_thresh" are the only ones that matter;
// the result is identical to the reference's top-k -> NMS -> score filter order (detector.py:88-114).
__device__ __forceinline__ sqd_u64 sqd_score_floor_key(float score_thr) {
    return ((sqd_u64)sqd_order_bits(score_thr) << 32) | 0xFFFFFFFFull;
}

// Per-image candidate lists in global memory: keys[img*stride + i], i < count[img] (unordered; the keys are unique
// and totally ordered, so every consumer is order independent).  stride >= num_anchors, so a list never overflows.
struct SqdCand {
    int *count;       // (B), zeroed by the host wrapper before the producer kernel
    sqd_u64 *keys;    // (B, stride)
    int stride;
};

// Warp-aggregated append (one atomic per image present in the warp).  Must be called by all 32 lanes, converged.
__device__ __forceinline__ void sqd_cand_append(const SqdCand &c, bool pass, int img, sqd_u64 key) {
    const unsigned act = __ballot_sync(0xffffffffu, pass);
    if (!pass) return;
    const unsigned peers = __match_any_sync(act, img);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(c.count + img, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    const int pos = base + __popc(peers & ((1u << lane) - 1u));
    if (pos < c.stride) c.keys[(size_t)img * c.stride + pos] = key;
}

// N candidates per lane, ALL rows of the warp in the same image (img warp-uniform): one atomic per call.
// Must be called by all 32 lanes, converged.
template <int N>
__device__ __forceinline__ void sqd_cand_append_warp(const SqdCand &c, int img, const bool (&pass)[N],
                                                     const sqd_u64 (&key)[N]) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int off[N];
    int total = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const unsigned b = __ballot_sync(0xffffffffu, pass[i]);
        off[i] = total + __popc(b & lt);
        total += __popc(b);
    }
    if (total == 0) return;  // warp-uniform
    int base = 0;
    if (lane == 0) base = atomicAdd(c.count + img, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    sqd_u64 *dst = c.keys + (size_t)img * c.stride + base;
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (pass[i] && base + off[i] < c.stride) dst[off[i]] = key[i];
}

// request to the tcgen05 ConvDet launcher: fill `cand` from the epilogue if the shape allows it
struct SqdCandEmit {
    SqdCand cand;
    int num_classes;
    float score_thr;
    int *done;   // host: set to 1 if the epilogue will emit, 0 if the caller must scan pred itself
};

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// The path is a chain of kernels on one stream (pre-pass -> GEMM -> scan -> tail).  A kernel launched with the
// programmatic-stream-serialization attribute may start while its predecessor is still running: its CTAs do their
// set-up (barrier init, TMEM allocation, descriptor prefetch ...) and then block in sqd_pdl_wait() until the predecessor
// grid has completed and its writes are visible.  The predecessor calls sqd_pdl_trigger() early to allow that.  Both
// are no-ops for kernels launched the ordinary way.  Disabled with SQD_NO_PDL=1 (then every launch is fully serialised).
__device__ __forceinline__ void sqd_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void sqd_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool sqd_pdl_enabled();   // api.cu (reads SQD_NO_PDL once per call)

template <class K, class... Args>
cudaError_t sqd_launch_dependent(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && sqd_pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

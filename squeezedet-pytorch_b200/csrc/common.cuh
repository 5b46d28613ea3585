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

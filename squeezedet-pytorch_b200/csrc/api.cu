// C-ABI glue: error plumbing, ConvDet algorithm dispatch, the fused head->detections entry point and
// the boxes_postprocess epilogue.  See include/sqdet_b200.h for the contract of every symbol.
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

// implemented in convdet_simt.cu / convdet_f16.cu
size_t sqd_simt_workspace_bytes(int cin, int cout);
int sqd_convdet_simt(const float *d_feat, int layout, const float *d_weight, const float *d_bias, int batch, int cin,
                     int gh, int gw, int cout, float *d_pred, void *d_workspace, cudaStream_t st);
size_t sqd_f16_packed_bytes(int cout, int cin);
int sqd_f16_pack_weights(const float *d_weight, int cout, int cin, void *d_packed, cudaStream_t st);
size_t sqd_f16_split_bytes(int batch, int cin, int gh, int gw);
int sqd_f16_split_features(const float *d_feat, int layout, int batch, int cin, int gh, int gw, void *d_planes,
                           cudaStream_t st);
size_t sqd_f16_workspace_bytes(int batch, int cin, int gh, int gw, int cout, int layout);
int sqd_convdet_f16_pair(const float *d_feat, int layout, const void *d_packed, const float *d_bias, int batch, int cin,
                         int gh, int gw, int cout, float *d_pred, void *d_workspace, cudaStream_t st,
                         const SqdCandEmit *emit, int out_stride, int ksteps_last, int slabs, size_t slab_stride,
                         int image_scales);
// implemented in topk_nms.cu
size_t sqd_cand_bytes(int batch, int num_anchors);
SqdCand sqd_cand_layout(void *ws, int batch, int num_anchors);
int sqd_score_candidates(const float *d_pred, int batch, int num_anchors, int num_classes, double score_thresh,
                         SqdCand cand, cudaStream_t st, bool pdl);
int sqd_detect_from_candidates(SqdCand cand, const float *d_pred, const float *d_anchors, int batch, int num_anchors,
                               int num_classes, int input_h, int input_w, int top_k, double nms_thresh,
                               double score_thresh, int32_t *d_count, int32_t *d_out_anchor, int32_t *d_out_class,
                               float *d_out_score, float *d_out_box, cudaStream_t st);
int sqd_detect_check_args(const char *fn, const void *d_pred, const void *d_anchors, int batch, int num_anchors,
                          int num_classes, int top_k, const void *count, const void *anchor, const void *cls,
                          const void *score, const void *box);
int sqd_convdet_f16(const float *d_feat, int layout, const void *d_packed, const float *d_bias, int batch, int cin,
                    int gh, int gw, int cout, float *d_pred, void *d_workspace, cudaStream_t st);
// implemented in convdet_fused.cu: the one-kernel path (fp32 NCHW features -> pred, operands converted inside the GEMM)
bool sqd_convdet_fused_eligible(int layout, int batch, int cin, int gh, int gw, int cout);
int sqd_convdet_fused(const float *d_feat, const void *d_packed, const float *d_bias, int batch, int cin, int gh, int gw,
                      int cout, float *d_pred, void *d_workspace, cudaStream_t st);

static thread_local char g_err[512] = "";

void sqd_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- developer options: one table, read from the environment once ---------------------------------------------
namespace {
struct OptDef {
    const char *name;
    int dflt;
};
const OptDef kOptDefs[SQD_OPT_COUNT] = {
    {"SQD_NO_PDL", 0}, {"SQD_FUSED_SCORE", 0}, {"SQD_SPLIT_TWO_PASS", 0}, {"SQD_SPLIT_CS", 0}, {"SQD_SPLIT_THREADS", 512},
    {"SQD_SPLIT_ROWS", 0}, {"SQD_DGRAD_PER_SLAB", 0}, {"SQD_DGRAD_BLOCK_SCALES", 0}, {"SQD_WG_SINGLE_TAP", 0}, {"SQD_WG_SYNC", 0},
    {"SQD_BWD_OLD_PREPASS", 0}, {"SQD_MATCH_SEQUENTIAL", 0}, {"SQD_F16_HALF_TILES", 0}, {"SQD_F16_CHUNK", 3}, {"SQD_F16_DBG", 0},
    {"SQD_F16_PAIR_STAGES", 4}, {"SQD_F16_A_STAGES", 2}, {"SQD_F16_B_STAGES", 8}, {"SQD_F16_TRACE_CTA", 0}, {"SQD_HEAD_ONE_KERNEL", 0},
    {"SQD_F16_A_ONCE", 0}, {"SQD_F16_AO_BUFS", 2}, {"SQD_TAIL_THREADS", 0}, {"SQD_DGRAD_PACK_LOOP", 0}, {"SQD_SPLIT_REGS", 0},
};
std::atomic<int> g_opt[SQD_OPT_COUNT];
std::once_flag g_opt_once;
void opt_init() {
    for (int i = 0; i < SQD_OPT_COUNT; ++i) {
        const char *e = getenv(kOptDefs[i].name);
        g_opt[i].store(e ? (*e ? atoi(e) : 1) : kOptDefs[i].dflt, std::memory_order_relaxed);
    }
}
int opt_index(const char *name) {
    if (!name) return -1;
    for (int i = 0; i < SQD_OPT_COUNT; ++i)
        if (strcmp(name, kOptDefs[i].name) == 0) return i;
    return -1;
}
}  // namespace

int sqd_opt(SqdOptId id) {
    std::call_once(g_opt_once, opt_init);
    return g_opt[id].load(std::memory_order_relaxed);
}

extern "C" int sqd_set_option(const char *name, int value) {
    SQD_REQUIRE(name, SQD_E_NULL, "sqd_set_option: NULL name");
    const int i = opt_index(name);
    SQD_REQUIRE(i >= 0, SQD_E_UNSUPPORTED, "sqd_set_option: unknown option %s", name ? name : "(null)");
    std::call_once(g_opt_once, opt_init);
    g_opt[i].store(value, std::memory_order_relaxed);
    return SQD_OK;
}

extern "C" int sqd_get_option(const char *name, int *value) {
    SQD_REQUIRE(name && value, SQD_E_NULL, "sqd_get_option: NULL pointer");
    const int i = opt_index(name);
    SQD_REQUIRE(i >= 0, SQD_E_UNSUPPORTED, "sqd_get_option: unknown option %s", name);
    *value = sqd_opt(static_cast<SqdOptId>(i));
    return SQD_OK;
}

bool sqd_pdl_enabled() { return sqd_opt(SQD_OPT_NO_PDL) == 0; }

extern "C" int sqd_abi_version(void) { return SQD_ABI_VERSION; }
extern "C" const char *sqd_last_error(void) { return g_err; }

// ---- a1 -------------------------------------------------------------------------------------------
extern "C" size_t sqd_convdet_packed_weight_bytes(int cout, int cin) {
    if (cout <= 0 || cin <= 0) return 0;
    return sqd_f16_packed_bytes(cout, cin);
}

extern "C" int sqd_convdet_pack_weights(const float *d_weight, int cout, int cin, void *d_packed, void *stream) {
    SQD_REQUIRE(d_weight && d_packed, SQD_E_NULL, "sqd_convdet_pack_weights: NULL pointer");
    SQD_REQUIRE(cout >= 1 && cout <= 128 && cin >= 64 && cin % 64 == 0, SQD_E_SHAPE,
                "sqd_convdet_pack_weights: need 1<=Cout<=128 and Cin a multiple of 64 (got %d, %d)", cout, cin);
    SQD_REQUIRE(sqd_aligned16(d_packed), SQD_E_ALIGN, "sqd_convdet_pack_weights: packed buffer must be 16-byte aligned");
    return sqd_f16_pack_weights(d_weight, cout, cin, d_packed, static_cast<cudaStream_t>(stream));
}

extern "C" size_t sqd_convdet_workspace_bytes(int batch, int cin, int gh, int gw, int cout, int layout, int algo) {
    if (batch <= 0 || cin <= 0 || gh <= 0 || gw <= 0 || cout <= 0) return 256;
    if (algo == SQD_CONV_SIMT_FP32) return align_up(sqd_simt_workspace_bytes(cin, cout), 256);
    return align_up(sqd_f16_workspace_bytes(batch, cin, gh, gw, cout, layout), 256);
}

static int convdet_forward_impl(const float *d_feat, int layout, const void *d_packed, const float *d_weight,
                                const float *d_bias, int batch, int cin, int gh, int gw, int cout, float *d_pred,
                                void *d_workspace, size_t workspace_bytes, int algo, void *stream, const SqdCandEmit *emit) {
    if (emit) *emit->done = 0;
    if (batch == 0) return SQD_OK;  // empty batch: nothing to enqueue, pointers may be NULL
    SQD_REQUIRE(d_feat && d_bias && d_pred && d_workspace, SQD_E_NULL, "sqd_convdet_forward: NULL pointer");
    SQD_REQUIRE(layout == SQD_LAYOUT_NCHW || layout == SQD_LAYOUT_NHWC || layout == SQD_LAYOUT_SPLIT_NHWC, SQD_E_SHAPE,
                "sqd_convdet_forward: bad layout %d", layout);
    SQD_REQUIRE(!(layout == SQD_LAYOUT_SPLIT_NHWC && algo == SQD_CONV_SIMT_FP32), SQD_E_UNSUPPORTED,
                "sqd_convdet_forward: pre-split planes are only consumed by the tcgen05 algorithm");
    SQD_REQUIRE(batch >= 0 && cin > 0 && gh > 0 && gw > 0 && cout > 0, SQD_E_SHAPE, "sqd_convdet_forward: bad shape");
    SQD_REQUIRE(sqd_aligned16(d_feat) && sqd_aligned16(d_pred) && sqd_aligned16(d_workspace), SQD_E_ALIGN,
                "sqd_convdet_forward: feat/pred/workspace must be 16-byte aligned");
    SQD_REQUIRE(workspace_bytes >= sqd_convdet_workspace_bytes(batch, cin, gh, gw, cout, layout, algo), SQD_E_WORKSPACE,
                "sqd_convdet_forward: workspace too small (%zu bytes)", workspace_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (algo == SQD_CONV_SIMT_FP32) {
        SQD_REQUIRE(d_weight, SQD_E_NULL, "sqd_convdet_forward: SIMT algorithm needs the raw weight tensor");
        return sqd_convdet_simt(d_feat, layout, d_weight, d_bias, batch, cin, gh, gw, cout, d_pred, d_workspace, st);
    }
    SQD_REQUIRE(algo == SQD_CONV_TCGEN05_F16X3 || algo == SQD_CONV_TCGEN05_F16X3_1CTA, SQD_E_UNSUPPORTED,
                "sqd_convdet_forward: unknown algo %d", algo);
    SQD_REQUIRE(d_packed, SQD_E_NULL, "sqd_convdet_forward: tcgen05 algorithm needs packed weights");
    // Opt-in route (SQD_HEAD_ONE_KERNEL=1): max pass + ONE GEMM kernel that converts the fp32 features itself (no fp16
    // planes in HBM, A patch produced once per (tile, block)).  Parity-checked against the default route by the tests;
    // MEASURED slower than pre-pass + GEMM at KITTI B = 20 (136 vs 132 us, profiles/r02_one_kernel_convdet.txt: register
    // loads through the LSU cannot stream 74 KB per block and SM), so it is not the default.
    if (algo == SQD_CONV_TCGEN05_F16X3 && sqd_opt(SQD_OPT_HEAD_ONE_KERNEL) && !(emit && sqd_opt(SQD_OPT_FUSED_SCORE)) &&
        sqd_convdet_fused_eligible(layout, batch, cin, gh, gw, cout))
        return sqd_convdet_fused(d_feat, d_packed, d_bias, batch, cin, gh, gw, cout, d_pred, d_workspace, st);
    if (algo == SQD_CONV_TCGEN05_F16X3)
        return sqd_convdet_f16_pair(d_feat, layout, d_packed, d_bias, batch, cin, gh, gw, cout, d_pred, d_workspace, st, emit, 0, 0, 1, 0, 0);
    return sqd_convdet_f16(d_feat, layout, d_packed, d_bias, batch, cin, gh, gw, cout, d_pred, d_workspace, st);
}

extern "C" int sqd_convdet_forward(const float *d_feat, int layout, const void *d_packed, const float *d_weight,
                                   const float *d_bias, int batch, int cin, int gh, int gw, int cout, float *d_pred,
                                   void *d_workspace, size_t workspace_bytes, int algo, void *stream) {
    return convdet_forward_impl(d_feat, layout, d_packed, d_weight, d_bias, batch, cin, gh, gw, cout, d_pred, d_workspace,
                                workspace_bytes, algo, stream, nullptr);
}

extern "C" size_t sqd_convdet_split_bytes(int batch, int cin, int gh, int gw) {
    if (batch <= 0 || cin <= 0 || gh <= 0 || gw <= 0) return 0;
    return sqd_f16_split_bytes(batch, cin, gh, gw);
}

extern "C" int sqd_convdet_split_features(const float *d_feat, int layout, int batch, int cin, int gh, int gw,
                                          void *d_planes, void *stream) {
    if (batch == 0) return SQD_OK;  // empty batch: nothing to enqueue, pointers may be NULL
    SQD_REQUIRE(d_feat && d_planes, SQD_E_NULL, "sqd_convdet_split_features: NULL pointer");
    SQD_REQUIRE(layout == SQD_LAYOUT_NCHW || layout == SQD_LAYOUT_NHWC, SQD_E_SHAPE,
                "sqd_convdet_split_features: bad layout %d", layout);
    SQD_REQUIRE(batch >= 0 && cin >= 64 && cin % 64 == 0 && gh > 0 && gw > 0, SQD_E_SHAPE,
                "sqd_convdet_split_features: bad shape");
    SQD_REQUIRE(sqd_aligned16(d_feat) && sqd_aligned16(d_planes), SQD_E_ALIGN,
                "sqd_convdet_split_features: pointers must be 16-byte aligned");
    return sqd_f16_split_features(d_feat, layout, batch, cin, gh, gw, d_planes, static_cast<cudaStream_t>(stream));
}

// Synchronises `stream` and reports whether the last tcgen05 launch that used this workspace drained cleanly
// (0) or hit a bounded-wait timeout (>0: 1 producer, 2 MMA issuer, 3 epilogue).  Debug / test aid.
extern "C" int sqd_convdet_status(const void *d_workspace, void *stream) {
    SQD_REQUIRE(d_workspace, SQD_E_NULL, "sqd_convdet_status: NULL workspace");
    int h = -1;
    SQD_CUDA(cudaMemcpyAsync(&h, static_cast<const int *>(d_workspace), sizeof(int), cudaMemcpyDeviceToHost,
                             static_cast<cudaStream_t>(stream)));
    SQD_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    if (h != 0) sqd_set_error("tcgen05 ConvDet pipeline timed out (role %d)", h);
    return h;
}

// ---- fused a1-a9 -----------------------------------------------------------------------------------
// workspace: [pred (B,A,C+5)][candidate lists: counts + (B,A) keys][ConvDet workspace]
extern "C" size_t sqd_head_detect_workspace_bytes(int batch, int cin, int gh, int gw, int cout, int layout, int algo) {
    if (batch <= 0 || cin <= 0 || gh <= 0 || gw <= 0 || cout <= 0) return 256;
    const size_t pred = align_up((size_t)batch * gh * gw * cout * sizeof(float), 256);
    // anchors per image = gh*gw*K <= gh*gw*cout/6 (an anchor has at least 6 fields)
    const size_t cand = sqd_cand_bytes(batch, gh * gw * (cout / 6 + 1));
    return pred + cand + sqd_convdet_workspace_bytes(batch, cin, gh, gw, cout, layout, algo);
}

// Byte offset, inside a sqd_head_detect_fused workspace, of the tcgen05 pipeline status word (int32; 0 = the last call
// drained cleanly, else the role whose bounded wait timed out).  Callers that synchronise anyway read it with their
// results instead of paying sqd_convdet_status' extra synchronisation.
extern "C" size_t sqd_head_detect_status_offset(int batch, int gh, int gw, int cout) {
    if (batch <= 0 || gh <= 0 || gw <= 0 || cout <= 0) return 0;
    return align_up((size_t)batch * gh * gw * cout * sizeof(float), 256) + sqd_cand_bytes(batch, gh * gw * (cout / 6 + 1));
}

namespace {
// Shared body of sqd_head_detect_fused / sqd_head_detect_profile.  ev (optional, 3 events) are recorded after the
// split pre-pass (profile form only, where the caller ran it), after the GEMM and after the filter.
int head_detect_impl(const float *d_feat, int layout, const void *d_packed, const float *d_weight, const float *d_bias,
                     const float *d_anchors, int batch, int cin, int gh, int gw, int anchors_per_grid, int num_classes,
                     int input_h, int input_w, int top_k, double nms_thresh, double score_thresh, int32_t *d_count,
                     int32_t *d_out_anchor, int32_t *d_out_class, float *d_out_score, float *d_out_box,
                     void *d_workspace, size_t workspace_bytes, int algo, void *stream, cudaEvent_t *ev) {
    const int cout = anchors_per_grid * (num_classes + 5);
    const int A = gh * gw * anchors_per_grid;
    const size_t pred_bytes = align_up((size_t)batch * gh * gw * cout * sizeof(float), 256);
    const size_t cand_bytes = sqd_cand_bytes(batch, gh * gw * (cout / 6 + 1));
    float *pred = static_cast<float *>(d_workspace);
    void *cand_ws = static_cast<char *>(d_workspace) + pred_bytes;
    void *conv_ws = static_cast<char *>(d_workspace) + pred_bytes + cand_bytes;
    int rc = sqd_detect_check_args("sqd_head_detect_fused", pred, d_anchors, batch, A, num_classes, top_k, d_count,
                                   d_out_anchor, d_out_class, d_out_score, d_out_box);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const SqdCand cand = sqd_cand_layout(cand_ws, batch, A);
    SQD_CUDA(cudaMemsetAsync(cand.count, 0, (size_t)batch * sizeof(int), st));
    int emitted = 0;
    SqdCandEmit emit{cand, num_classes, (float)score_thresh, &emitted};
    // a1 (+ a2-a7 scoring in the GEMM epilogue where the shape has a fused instantiation)
    rc = convdet_forward_impl(d_feat, layout, d_packed, d_weight, d_bias, batch, cin, gh, gw, cout, pred, conv_ws,
                              workspace_bytes - pred_bytes - cand_bytes, algo, stream, &emit);
    if (rc) return rc;
    if (ev) SQD_CUDA(cudaEventRecord(ev[0], st));
    if (!emitted) {
        rc = sqd_score_candidates(pred, batch, A, num_classes, score_thresh, cand, st, /*pdl=*/algo != SQD_CONV_SIMT_FP32);
        if (rc) return rc;
    }
    // a8-a9 on the candidates
    rc = sqd_detect_from_candidates(cand, pred, d_anchors, batch, A, num_classes, input_h, input_w, top_k, nms_thresh,
                                    score_thresh, d_count, d_out_anchor, d_out_class, d_out_score, d_out_box, st);
    if (rc) return rc;
    if (ev) SQD_CUDA(cudaEventRecord(ev[1], st));
    return SQD_OK;
}
}  // namespace

extern "C" int sqd_head_detect_fused(const float *d_feat, int layout, const void *d_packed, const float *d_weight,
                                     const float *d_bias, const float *d_anchors, int batch, int cin, int gh, int gw,
                                     int anchors_per_grid, int num_classes, int input_h, int input_w, int top_k,
                                     double nms_thresh, double score_thresh, int32_t *d_count, int32_t *d_out_anchor,
                                     int32_t *d_out_class, float *d_out_score, float *d_out_box, void *d_workspace,
                                     size_t workspace_bytes, int algo, void *stream) {
    if (batch == 0) return SQD_OK;  // empty batch: nothing to enqueue, pointers may be NULL
    SQD_REQUIRE(anchors_per_grid >= 1 && num_classes >= 1, SQD_E_SHAPE, "sqd_head_detect_fused: bad anchor/class count");
    const int cout = anchors_per_grid * (num_classes + 5);
    SQD_REQUIRE(d_workspace, SQD_E_NULL, "sqd_head_detect_fused: NULL workspace");
    SQD_REQUIRE(workspace_bytes >= sqd_head_detect_workspace_bytes(batch, cin, gh, gw, cout, layout, algo),
                SQD_E_WORKSPACE, "sqd_head_detect_fused: workspace too small (%zu bytes)", workspace_bytes);
    return head_detect_impl(d_feat, layout, d_packed, d_weight, d_bias, d_anchors, batch, cin, gh, gw, anchors_per_grid,
                            num_classes, input_h, input_w, top_k, nms_thresh, score_thresh, d_count, d_out_anchor,
                            d_out_class, d_out_score, d_out_box, d_workspace, workspace_bytes, algo, stream, nullptr);
}

// Diagnostic twin of sqd_head_detect_fused: the same kernel sequence with CUDA events recorded on `stream` between
// the stages; synchronises the stream and returns the stage durations in milliseconds:
// h_stage_ms[0] split pre-pass (max|x| + fp16 planes), [1] ConvDet GEMM (+ score epilogue), [2] filter (select/NMS/emit,
// preceded by the pred scan when the epilogue did not score).  tcgen05 algorithm, NCHW / NHWC input only.
extern "C" size_t sqd_head_detect_profile_workspace_bytes(int batch, int cin, int gh, int gw, int cout) {
    if (batch <= 0 || cin <= 0 || gh <= 0 || gw <= 0 || cout <= 0) return 256;
    return sqd_head_detect_workspace_bytes(batch, cin, gh, gw, cout, SQD_LAYOUT_SPLIT_NHWC, SQD_CONV_TCGEN05_F16X3) +
           align_up(sqd_f16_split_bytes(batch, cin, gh, gw), 256);
}

extern "C" int sqd_head_detect_profile(const float *d_feat, int layout, const void *d_packed, const float *d_bias,
                                       const float *d_anchors, int batch, int cin, int gh, int gw, int anchors_per_grid,
                                       int num_classes, int input_h, int input_w, int top_k, double nms_thresh,
                                       double score_thresh, int32_t *d_count, int32_t *d_out_anchor,
                                       int32_t *d_out_class, float *d_out_score, float *d_out_box, void *d_workspace,
                                       size_t workspace_bytes, void *stream, float *h_stage_ms) {
    SQD_REQUIRE(batch > 0 && anchors_per_grid >= 1 && num_classes >= 1, SQD_E_SHAPE, "sqd_head_detect_profile: bad shape");
    SQD_REQUIRE(layout == SQD_LAYOUT_NCHW || layout == SQD_LAYOUT_NHWC, SQD_E_SHAPE, "sqd_head_detect_profile: bad layout");
    SQD_REQUIRE(d_feat && d_workspace && h_stage_ms, SQD_E_NULL, "sqd_head_detect_profile: NULL pointer");
    const int cout = anchors_per_grid * (num_classes + 5);
    const size_t fused_bytes =
        sqd_head_detect_workspace_bytes(batch, cin, gh, gw, cout, SQD_LAYOUT_SPLIT_NHWC, SQD_CONV_TCGEN05_F16X3);
    SQD_REQUIRE(workspace_bytes >= sqd_head_detect_profile_workspace_bytes(batch, cin, gh, gw, cout), SQD_E_WORKSPACE,
                "sqd_head_detect_profile: workspace too small (%zu bytes)", workspace_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void *planes = static_cast<char *>(d_workspace) + fused_bytes;
    cudaEvent_t ev[4];
    for (int i = 0; i < 4; ++i) SQD_CUDA(cudaEventCreate(&ev[i]));
    int rc = SQD_OK;
    SQD_CUDA(cudaEventRecord(ev[0], st));
    if (sqd_opt(SQD_OPT_HEAD_ONE_KERNEL) && sqd_convdet_fused_eligible(layout, batch, cin, gh, gw, cout)) {
        // one-kernel route: there is no pre-pass (stage 0 = 0 ms), the GEMM reads the fp32 features itself
        SQD_CUDA(cudaEventRecord(ev[1], st));
        rc = head_detect_impl(d_feat, layout, d_packed, nullptr, d_bias, d_anchors, batch, cin, gh, gw, anchors_per_grid,
                              num_classes, input_h, input_w, top_k, nms_thresh, score_thresh, d_count, d_out_anchor,
                              d_out_class, d_out_score, d_out_box, d_workspace, workspace_bytes, SQD_CONV_TCGEN05_F16X3, stream, ev + 2);
    } else {
        rc = sqd_convdet_split_features(d_feat, layout, batch, cin, gh, gw, planes, stream);
        if (rc == SQD_OK) {
            SQD_CUDA(cudaEventRecord(ev[1], st));
            rc = head_detect_impl(static_cast<const float *>(planes), SQD_LAYOUT_SPLIT_NHWC, d_packed, nullptr, d_bias, d_anchors,
                                  batch, cin, gh, gw, anchors_per_grid, num_classes, input_h, input_w, top_k, nms_thresh,
                                  score_thresh, d_count, d_out_anchor, d_out_class, d_out_score, d_out_box, d_workspace,
                                  fused_bytes, SQD_CONV_TCGEN05_F16X3, stream, ev + 2);
        }
    }
    if (rc == SQD_OK) {
        SQD_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < 3; ++i) SQD_CUDA(cudaEventElapsedTime(h_stage_ms + i, ev[i], ev[i + 1]));
    }
    for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]);
    return rc;
}

// ---- fused a1-a9 with HOST buffers: chunked copy/compute pipeline ---------------------------------------------
namespace {
struct HostWs {
    size_t feat_off, count_off, anchor_off, cls_off, score_off, box_off, fused_off, total;
};
HostWs host_ws_layout(int batch, int cin, int gh, int gw, int cout, int top_k, int layout, int algo, int chunk) {
    HostWs w;
    size_t off = 0;
    w.feat_off = off;   off += align_up((size_t)batch * cin * gh * gw * sizeof(float), 256);
    w.count_off = off;  off += align_up((size_t)batch * sizeof(int32_t), 256);
    w.anchor_off = off; off += align_up((size_t)batch * top_k * sizeof(int32_t), 256);
    w.cls_off = off;    off += align_up((size_t)batch * top_k * sizeof(int32_t), 256);
    w.score_off = off;  off += align_up((size_t)batch * top_k * sizeof(float), 256);
    w.box_off = off;    off += align_up((size_t)batch * top_k * 4 * sizeof(float), 256);
    w.fused_off = off;  off += sqd_head_detect_workspace_bytes(chunk, cin, gh, gw, cout, layout, algo);
    w.total = off;
    return w;
}
int clamp_chunk(int batch, int chunk) { return chunk <= 0 || chunk > batch ? batch : chunk; }

// Cross-stream ordering events of sqd_head_detect_host: a few per (host thread, device), created once and re-recorded.
// Safe: cudaStreamWaitEvent binds to the record that is current when it is CALLED, a later re-record does not move it.
struct EventRing {
    static constexpr int kN = 4;
    cudaEvent_t ev[kN] = {nullptr, nullptr, nullptr, nullptr};
    int device = -1, next = 0;
};
int host_event(cudaEvent_t *out) {
    static thread_local EventRing rings[16];
    int dev = 0;
    SQD_CUDA(cudaGetDevice(&dev));
    EventRing &r = rings[dev & 15];
    if (r.device != dev) {           // first use on this device by this thread (or a colliding device id: rebuild)
        for (int i = 0; i < EventRing::kN; ++i) {
            if (r.ev[i]) cudaEventDestroy(r.ev[i]);
            SQD_CUDA(cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming));
        }
        r.device = dev;
        r.next = 0;
    }
    *out = r.ev[r.next];
    r.next = (r.next + 1) % EventRing::kN;
    return SQD_OK;
}
}  // namespace

extern "C" size_t sqd_head_detect_host_workspace_bytes(int batch, int cin, int gh, int gw, int cout, int top_k, int layout,
                                                       int algo, int chunk_images) {
    if (batch <= 0 || cin <= 0 || gh <= 0 || gw <= 0 || cout <= 0 || top_k <= 0) return 256;
    return host_ws_layout(batch, cin, gh, gw, cout, top_k, layout, algo, clamp_chunk(batch, chunk_images)).total;
}

// Byte offsets of the five result arrays (count, anchor, class, score, box) inside ONE host block of *total bytes that
// mirrors the device-side result block: a caller that passes views of such a block to sqd_head_detect_host gets its
// detections with a single device-to-host copy instead of five.
extern "C" int sqd_head_detect_host_result_layout(int batch, int top_k, size_t *offsets5, size_t *total) {
    SQD_REQUIRE(offsets5 && total, SQD_E_NULL, "sqd_head_detect_host_result_layout: NULL pointer");
    SQD_REQUIRE(batch >= 1 && top_k >= 1, SQD_E_SHAPE, "sqd_head_detect_host_result_layout: bad shape");
    const HostWs w = host_ws_layout(batch, 64, 1, 1, 6, top_k, SQD_LAYOUT_NCHW, SQD_CONV_TCGEN05_F16X3, 1);
    offsets5[0] = 0;
    offsets5[1] = w.anchor_off - w.count_off;
    offsets5[2] = w.cls_off - w.count_off;
    offsets5[3] = w.score_off - w.count_off;
    offsets5[4] = w.box_off - w.count_off;
    *total = w.fused_off - w.count_off;
    return SQD_OK;
}

extern "C" int sqd_head_detect_host(const float *h_feat, int layout, const void *d_packed, const float *d_weight,
                                    const float *d_bias, const float *d_anchors, int batch, int cin, int gh, int gw,
                                    int anchors_per_grid, int num_classes, int input_h, int input_w, int top_k,
                                    double nms_thresh, double score_thresh, int32_t *h_count, int32_t *h_out_anchor,
                                    int32_t *h_out_class, float *h_out_score, float *h_out_box, void *d_workspace,
                                    size_t workspace_bytes, int algo, int chunk_images, void *stream, void *copy_stream,
                                    int flags) {
    if (batch == 0) return SQD_OK;
    SQD_REQUIRE((flags & ~SQD_HOST_NO_STAGING_FENCE) == 0, SQD_E_UNSUPPORTED, "sqd_head_detect_host: unknown flags 0x%x",
                flags);
    SQD_REQUIRE(h_feat && h_count && h_out_anchor && h_out_class && h_out_score && h_out_box && d_workspace, SQD_E_NULL,
                "sqd_head_detect_host: NULL pointer");
    SQD_REQUIRE(layout == SQD_LAYOUT_NCHW || layout == SQD_LAYOUT_NHWC, SQD_E_SHAPE, "sqd_head_detect_host: bad layout %d",
                layout);
    SQD_REQUIRE(anchors_per_grid >= 1 && num_classes >= 1 && batch > 0 && top_k >= 1, SQD_E_SHAPE,
                "sqd_head_detect_host: bad shape");
    const int cout = anchors_per_grid * (num_classes + 5);
    const int chunk = clamp_chunk(batch, chunk_images);
    const HostWs w = host_ws_layout(batch, cin, gh, gw, cout, top_k, layout, algo, chunk);
    SQD_REQUIRE(workspace_bytes >= w.total, SQD_E_WORKSPACE, "sqd_head_detect_host: workspace too small (%zu < %zu bytes)",
                workspace_bytes, w.total);
    SQD_REQUIRE(sqd_aligned16(d_workspace), SQD_E_ALIGN, "sqd_head_detect_host: workspace must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaStream_t cst = copy_stream ? static_cast<cudaStream_t>(copy_stream) : st;
    char *ws = static_cast<char *>(d_workspace);
    float *d_feat = reinterpret_cast<float *>(ws + w.feat_off);
    int32_t *d_count = reinterpret_cast<int32_t *>(ws + w.count_off);
    int32_t *d_anchor = reinterpret_cast<int32_t *>(ws + w.anchor_off);
    int32_t *d_cls = reinterpret_cast<int32_t *>(ws + w.cls_off);
    float *d_score = reinterpret_cast<float *>(ws + w.score_off);
    float *d_box = reinterpret_cast<float *>(ws + w.box_off);
    const size_t img_elems = (size_t)cin * gh * gw;
    const int nchunks = (batch + chunk - 1) / chunk;
    cudaEvent_t ev = nullptr;
    if (cst != st && !(flags & SQD_HOST_NO_STAGING_FENCE)) {
        // the staging buffer may still be read by kernels of the previous call on `stream`
        if (int rc_ev = host_event(&ev)) return rc_ev;
        SQD_CUDA(cudaEventRecord(ev, st));
        SQD_CUDA(cudaStreamWaitEvent(cst, ev, 0));
    }
    int rc = SQD_OK;
    for (int c = 0; c < nchunks && rc == SQD_OK; ++c) {
        const int b0 = c * chunk, nb = (b0 + chunk <= batch) ? chunk : batch - b0;
        SQD_CUDA(cudaMemcpyAsync(d_feat + b0 * img_elems, h_feat + b0 * img_elems, nb * img_elems * sizeof(float),
                                 cudaMemcpyHostToDevice, cst));
        if (cst != st) {
            if (int rc_ev = host_event(&ev)) return rc_ev;
            SQD_CUDA(cudaEventRecord(ev, cst));
            SQD_CUDA(cudaStreamWaitEvent(st, ev, 0));
        }
        rc = sqd_head_detect_fused(d_feat + b0 * img_elems, layout, d_packed, d_weight, d_bias, d_anchors, nb, cin, gh, gw,
                                   anchors_per_grid, num_classes, input_h, input_w, top_k, nms_thresh, score_thresh,
                                   d_count + b0, d_anchor + (size_t)b0 * top_k, d_cls + (size_t)b0 * top_k,
                                   d_score + (size_t)b0 * top_k, d_box + (size_t)b0 * top_k * 4, ws + w.fused_off,
                                   workspace_bytes - w.fused_off, algo, stream);
    }
    if (rc) return rc;
    {   // one D2H when the caller's five buffers are views of ONE block laid out like the device block
        // (sqd_head_detect_host_result_layout): the usual case for the Python wrapper's HostDetections
        const char *h0 = reinterpret_cast<const char *>(h_count);
        const bool mirrored = reinterpret_cast<const char *>(h_out_anchor) - h0 == (ptrdiff_t)(w.anchor_off - w.count_off) &&
                              reinterpret_cast<const char *>(h_out_class) - h0 == (ptrdiff_t)(w.cls_off - w.count_off) &&
                              reinterpret_cast<const char *>(h_out_score) - h0 == (ptrdiff_t)(w.score_off - w.count_off) &&
                              reinterpret_cast<const char *>(h_out_box) - h0 == (ptrdiff_t)(w.box_off - w.count_off);
        if (mirrored) {
            SQD_CUDA(cudaMemcpyAsync(h_count, d_count, w.fused_off - w.count_off, cudaMemcpyDeviceToHost, st));
            return SQD_OK;
        }
    }
    SQD_CUDA(cudaMemcpyAsync(h_count, d_count, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SQD_CUDA(cudaMemcpyAsync(h_out_anchor, d_anchor, (size_t)batch * top_k * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SQD_CUDA(cudaMemcpyAsync(h_out_class, d_cls, (size_t)batch * top_k * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SQD_CUDA(cudaMemcpyAsync(h_out_score, d_score, (size_t)batch * top_k * sizeof(float), cudaMemcpyDeviceToHost, st));
    SQD_CUDA(cudaMemcpyAsync(h_out_box, d_box, (size_t)batch * top_k * 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
    return SQD_OK;
}

// ---- 8(f) rank 1: boxes_postprocess -----------------------------------------------------------------
// Reference order (src/utils/boxes.py:145-166): /scale, -padding, +crops, flip, +drifts.  One 10-float record
// per image: [scale_y, scale_x, pad_top, pad_left, crop_top, crop_left, flip_width (<=0: not flipped),
// drift_y, drift_x, 0]; absent keys are passed as their identity (scale 1, offsets 0).
namespace {
__global__ void postprocess_kernel(float4 *boxes, const int *count, const float *meta, int k) {
    const int img = blockIdx.x;
    const int n = count[img];
    const float *m = meta + (size_t)img * 10;
    const float sy = m[0], sx = m[1], pt = m[2], pl = m[3], ct = m[4], cl = m[5], fw = m[6], dy = m[7], dx = m[8];
    for (int i = threadIdx.x; i < n && i < k; i += blockDim.x) {
        float4 b = boxes[(size_t)img * k + i];
        b.x = fdiv(b.x, sx); b.z = fdiv(b.z, sx); b.y = fdiv(b.y, sy); b.w = fdiv(b.w, sy);
        b.x = fsub(b.x, pl); b.z = fsub(b.z, pl); b.y = fsub(b.y, pt); b.w = fsub(b.w, pt);
        b.x = fadd(b.x, cl); b.z = fadd(b.z, cl); b.y = fadd(b.y, ct); b.w = fadd(b.w, ct);
        if (fw > 0.f) {
            const float w = fadd(fsub(b.z, b.x), 1.f);
            b.x = fsub(fsub(fw, 1.f), b.z);
            b.z = fsub(fadd(b.x, w), 1.f);
        }
        b.x = fadd(b.x, dx); b.z = fadd(b.z, dx); b.y = fadd(b.y, dy); b.w = fadd(b.w, dy);
        boxes[(size_t)img * k + i] = b;
    }
}
}  // namespace

extern "C" int sqd_boxes_postprocess(float *d_boxes, const int32_t *d_count, const float *d_meta, int batch, int top_k,
                                     void *stream) {
    if (batch == 0) return SQD_OK;  // empty batch: nothing to enqueue, pointers may be NULL
    SQD_REQUIRE(d_boxes && d_count && d_meta, SQD_E_NULL, "sqd_boxes_postprocess: NULL pointer");
    SQD_REQUIRE(batch >= 0 && top_k >= 1, SQD_E_SHAPE, "sqd_boxes_postprocess: bad shape");
    SQD_REQUIRE(sqd_aligned16(d_boxes), SQD_E_ALIGN, "sqd_boxes_postprocess: boxes must be 16-byte aligned");
    postprocess_kernel<<<batch, 64, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<float4 *>(d_boxes), d_count,
                                                                           d_meta, top_k);
    SQD_LAUNCH_CHECK("postprocess_kernel");
    return SQD_OK;
}

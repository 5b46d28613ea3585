// SURVEY 8(f) "next" rows: the steps immediately before and after the detection path.
//   8f.1  boxes_postprocess + result packing: ONE (B, k, 6) array + counts, one D2H copy for a batch
//         (src/utils/boxes.py:138-168, src/engine/detector.py:37-40)
//   8f.3  KITTI result writer: the text lines of KITTI.save_results (src/datasets/kitti.py:78-97), formatted on the
//         host from the packed array
//   8f.4  input pre-processing: whiten + bilinear resize + HWC->CHW (src/utils/image.py:9-19,77-88 with cv2.resize's
//         float INTER_LINEAR arithmetic; src/datasets/base.py:33,49-59) as one HBM-bound kernel
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace {

// boxes_postprocess in the reference's order (/scale, -padding, +crops, flip, +drifts), meta record as documented
// for sqd_boxes_postprocess in the header.
__device__ __forceinline__ float4 postprocess_box(float4 b, const float *m) {
    const float sy = m[0], sx = m[1], pt = m[2], pl = m[3], ct = m[4], cl = m[5], fw = m[6], dy = m[7], dx = m[8];
    b.x = fdiv(b.x, sx); b.z = fdiv(b.z, sx); b.y = fdiv(b.y, sy); b.w = fdiv(b.w, sy);
    b.x = fsub(b.x, pl); b.z = fsub(b.z, pl); b.y = fsub(b.y, pt); b.w = fsub(b.w, pt);
    b.x = fadd(b.x, cl); b.z = fadd(b.z, cl); b.y = fadd(b.y, ct); b.w = fadd(b.w, ct);
    if (fw > 0.f) {
        const float w = fadd(fsub(b.z, b.x), 1.f);
        b.x = fsub(fsub(fw, 1.f), b.z);
        b.z = fsub(fadd(b.x, w), 1.f);
    }
    b.x = fadd(b.x, dx); b.z = fadd(b.z, dx); b.y = fadd(b.y, dy); b.w = fadd(b.w, dy);
    return b;
}

__global__ void pack_results_kernel(const int *__restrict__ count, const int *__restrict__ cls, const float *__restrict__ score,
                                    const float4 *__restrict__ box, const float *__restrict__ meta, int k,
                                    float *__restrict__ packed) {
    const int img = blockIdx.x;
    const int n = min(count[img], k);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const size_t r = (size_t)img * k + i;
        float *o = packed + r * 6;
        if (i < n) {
            float4 b = box[r];
            if (meta) b = postprocess_box(b, meta + (size_t)img * 10);
            o[0] = (float)cls[r];
            o[1] = score[r];
            o[2] = b.x; o[3] = b.y; o[4] = b.z; o[5] = b.w;
        } else {
            o[0] = -1.f;
            o[1] = o[2] = o[3] = o[4] = o[5] = 0.f;
        }
    }
}

// cv2.resize(float32, INTER_LINEAR) source coordinate of a destination index: fx = (d + 0.5) * scale - 0.5 in DOUBLE
// rounded to float, s = floor(fx), weight fx - s, clamped at both borders the way cv::resize builds its xofs / alpha
// tables (imgproc/src/resize.cpp).
__device__ __forceinline__ void src_coord(int d, double scale, int src_n, int &s0, int &s1, float &w1) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) {
        f = 0.f;
        s = 0;
    }
    if (s >= src_n - 1) {
        f = 0.f;
        s = src_n - 1;
    }
    s0 = s;
    s1 = min(s + 1, src_n - 1);
    w1 = f;
}

template <typename T>
__global__ void __launch_bounds__(256) preprocess_kernel(const T *__restrict__ img, int H0, int W0, float3 mean, float3 stdv,
                                                         int H, int W, double scale_y, double scale_x,
                                                         float *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    int x0, x1, y0, y1;
    float fx, fy;
    src_coord(x, scale_x, W0, x0, x1, fx);
    src_coord(y, scale_y, H0, y0, y1, fy);
    const T *src = img + (size_t)b * H0 * W0 * 3;
    const float m[3] = {mean.x, mean.y, mean.z}, s[3] = {stdv.x, stdv.y, stdv.z};
    const float a0 = 1.f - fx, b0 = 1.f - fy;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // whiten first (image.py:16), then interpolate: horizontal pass per source row, then vertical (cv::resize order)
        const float v00 = fdiv(fsub((float)src[((size_t)y0 * W0 + x0) * 3 + c], m[c]), s[c]);
        const float v01 = fdiv(fsub((float)src[((size_t)y0 * W0 + x1) * 3 + c], m[c]), s[c]);
        const float v10 = fdiv(fsub((float)src[((size_t)y1 * W0 + x0) * 3 + c], m[c]), s[c]);
        const float v11 = fdiv(fsub((float)src[((size_t)y1 * W0 + x1) * 3 + c], m[c]), s[c]);
        const float r0 = fadd(fmul(v00, a0), fmul(v01, fx));
        const float r1 = fadd(fmul(v10, a0), fmul(v11, fx));
        out[(((size_t)b * 3 + c) * H + y) * W + x] = fadd(fmul(r0, b0), fmul(r1, fy));
    }
}

}  // namespace

extern "C" int sqd_pack_results(const int32_t *d_count, const int32_t *d_class, const float *d_score, const float *d_box,
                                const float *d_meta, int batch, int top_k, float *d_packed, void *stream) {
    if (batch == 0) return SQD_OK;
    SQD_REQUIRE(d_count && d_class && d_score && d_box && d_packed, SQD_E_NULL, "sqd_pack_results: NULL pointer");
    SQD_REQUIRE(batch > 0 && top_k >= 1, SQD_E_SHAPE, "sqd_pack_results: bad shape");
    SQD_REQUIRE(sqd_aligned16(d_box), SQD_E_ALIGN, "sqd_pack_results: boxes must be 16-byte aligned");
    pack_results_kernel<<<batch, 64, 0, static_cast<cudaStream_t>(stream)>>>(d_count, d_class, d_score,
                                                                           reinterpret_cast<const float4 *>(d_box), d_meta,
                                                                           top_k, d_packed);
    SQD_LAUNCH_CHECK("pack_results_kernel");
    return SQD_OK;
}

// Host-side formatter.  One text block per image, the lines of KITTI.save_results (kitti.py:91-96):
//   "{class_name.lower()} -1 -1 0 {x1:.2f} {y1:.2f} {x2:.2f} {y2:.2f} 0 0 0 0 0 0 0 {score:.3f}\n"
// h_offsets[b] .. h_offsets[b+1] is image b's block inside h_out.  Returns the number of bytes needed (so a call with
// cap == 0 sizes the buffer), or < 0 on a bad argument.  Python's '{:.2f}'.format(np.float32) and printf("%.2f",
// (double)f) both print the correctly rounded decimal of the same binary value.
extern "C" long long sqd_format_kitti(const float *h_packed, const int32_t *h_count, int batch, int top_k,
                                      const char *const *class_names_lower, int num_classes, char *h_out, size_t cap,
                                      long long *h_offsets) {
    if (!h_packed || !h_count || !class_names_lower || batch < 0 || top_k < 1 || num_classes < 1) {
        sqd_set_error("sqd_format_kitti: bad argument");
        return SQD_E_NULL;
    }
    size_t pos = 0;
    char line[512];
    for (int b = 0; b < batch; ++b) {
        if (h_offsets) h_offsets[b] = (long long)pos;
        const int n = h_count[b] < top_k ? h_count[b] : top_k;
        for (int i = 0; i < n; ++i) {
            const float *r = h_packed + ((size_t)b * top_k + i) * 6;
            const int c = (int)r[0];
            if (c < 0 || c >= num_classes) {
                sqd_set_error("sqd_format_kitti: class id %d outside [0,%d) at image %d row %d", c, num_classes, b, i);
                return SQD_E_SHAPE;
            }
            const int len = snprintf(line, sizeof(line), "%s -1 -1 0 %.2f %.2f %.2f %.2f 0 0 0 0 0 0 0 %.3f\n",
                                     class_names_lower[c], (double)r[2], (double)r[3], (double)r[4], (double)r[5],
                                     (double)r[1]);
            if (len < 0 || len >= (int)sizeof(line)) {
                sqd_set_error("sqd_format_kitti: line too long");
                return SQD_E_SHAPE;
            }
            if (h_out && pos + (size_t)len <= cap) memcpy(h_out + pos, line, (size_t)len);
            pos += (size_t)len;
        }
    }
    if (h_offsets) h_offsets[batch] = (long long)pos;
    return (long long)pos;
}

extern "C" int sqd_preprocess(const void *d_images, int dtype, int batch, int src_h, int src_w, const float *mean3,
                              const float *std3, int dst_h, int dst_w, float *d_out, void *stream) {
    if (batch == 0) return SQD_OK;
    SQD_REQUIRE(d_images && d_out && mean3 && std3, SQD_E_NULL, "sqd_preprocess: NULL pointer");
    SQD_REQUIRE(batch > 0 && batch <= 65535 && src_h >= 1 && src_w >= 1 && dst_h >= 1 && dst_h <= 65535 && dst_w >= 1,
                SQD_E_SHAPE, "sqd_preprocess: bad shape");
    SQD_REQUIRE(dtype == 0 || dtype == 1, SQD_E_UNSUPPORTED, "sqd_preprocess: dtype must be 0 (uint8) or 1 (float32)");
    const float3 mean = make_float3(mean3[0], mean3[1], mean3[2]), stdv = make_float3(std3[0], std3[1], std3[2]);
    // cv::resize: inv_scale = (double)dst / src, scale = 1. / inv_scale
    const double sy = 1.0 / ((double)dst_h / src_h), sx = 1.0 / ((double)dst_w / src_w);
    const dim3 grid((dst_w + 255) / 256, dst_h, batch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == 0)
        preprocess_kernel<unsigned char><<<grid, 256, 0, st>>>(static_cast<const unsigned char *>(d_images), src_h, src_w, mean,
                                                              stdv, dst_h, dst_w, sy, sx, d_out);
    else
        preprocess_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float *>(d_images), src_h, src_w, mean, stdv, dst_h,
                                                      dst_w, sy, sx, d_out);
    SQD_LAUNCH_CHECK("preprocess_kernel");
    return SQD_OK;
}

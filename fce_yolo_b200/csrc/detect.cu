// Detect decode: DFL softmax-integral + dist2bbox + stride scaling + class sigmoid, fused
// (reference head.py:149-167, block.py:76-79, tal.py:352-376).  One thread per anchor; the output
// [B, 4+nc, A] is written channel-major so stores coalesce across anchors.
// HBM-bound: algorithmic bytes per image = (4*reg_max+nc)*A*4 read + (4+nc)*A*4 written.
#include "common.cuh"

namespace fce {
namespace {

constexpr int NT = 128;
constexpr int MAX_REG = 16;

__global__ void __launch_bounds__(NT) decode_kernel(const fce_decode_desc d, const float* __restrict__ r0,
                                                    const float* __restrict__ r1, const float* __restrict__ r2,
                                                    const float* __restrict__ r3, float* __restrict__ y, int A) {
    const int b = blockIdx.y;
    const int a = blockIdx.x * NT + threadIdx.x;
    if (a >= A) return;
    // locate the level (levels are concatenated P3, P4, P5: head.py:157)
    int lvl = 0, rem = a;
    const float* raws[4] = {r0, r1, r2, r3};
    while (lvl < d.nl - 1 && rem >= d.H[lvl] * d.W[lvl]) {
        rem -= d.H[lvl] * d.W[lvl];
        ++lvl;
    }
    const int hw = d.H[lvl] * d.W[lvl];
    const int ph = rem / d.W[lvl], pw = rem - ph * d.W[lvl];
    const float* p = raws[lvl] + ((size_t)b * hw + rem) * d.raw_pitch[lvl];
    const float stride = d.stride[lvl];
    const int R = d.reg_max;

    float dist[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        float v[MAX_REG];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < MAX_REG; i += 4) {
            if (i < R) {
                const float4 t = *reinterpret_cast<const float4*>(p + s * R + i);
                v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
            }
        }
#pragma unroll
        for (int i = 0; i < MAX_REG; ++i)
            if (i < R) mx = fmaxf(mx, v[i]);
        float sum = 0.f, wsum = 0.f;
#pragma unroll
        for (int i = 0; i < MAX_REG; ++i)
            if (i < R) {
                const float e = expf(v[i] - mx);
                sum += e;
                wsum = fmaf(e, (float)i, wsum);
            }
        dist[s] = wsum / sum;
    }
    const float ax = (float)pw + 0.5f, ay = (float)ph + 0.5f;
    const float x1 = ax - dist[0], y1 = ay - dist[1], x2 = ax + dist[2], y2 = ay + dist[3];
    const size_t yb = (size_t)b * (4 + d.nc) * A + a;
    y[yb] = (x1 + x2) * 0.5f * stride;
    y[yb + (size_t)A] = (y1 + y2) * 0.5f * stride;
    y[yb + (size_t)2 * A] = (x2 - x1) * stride;
    y[yb + (size_t)3 * A] = (y2 - y1) * stride;
    const float* pc = p + 4 * R;
    for (int c = 0; c < d.nc; c += 4) {
        const float4 t = *reinterpret_cast<const float4*>(pc + c);
        const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (c + j < d.nc) y[yb + (size_t)(4 + c + j) * A] = sigmoid_acc(tv[j]);
    }
}


// ------------------------------------------------------------------------------------------------ scale_boxes
// ops.scale_boxes + clip_boxes (ultralytics/utils/ops.py:102-134, 152-177) as DetectionPredictor.construct_result
// applies them (models/yolo/detect/predict.py:109-122): detections of the letterboxed image -> original image
// coordinates, in place on the padded [B, max_det, 6] buffer.  meta[b] = {gain, pad_x, pad_y, w0, h0} fp32.  Same fp32
// operations in the same order as the reference's CPU path (subtract pad, IEEE divide by gain, clamp).
__global__ void __launch_bounds__(NT) scale_boxes_kernel(float* __restrict__ det, const int32_t* __restrict__ count,
                                                         const float* __restrict__ meta, int max_det) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * NT + threadIdx.x;
    if (q >= max_det || q >= count[b]) return;
    const float gain = meta[b * 5], px = meta[b * 5 + 1], py = meta[b * 5 + 2], w0 = meta[b * 5 + 3], h0 = meta[b * 5 + 4];
    float4* p = reinterpret_cast<float4*>(det + ((size_t)b * max_det + q) * 6);  // 24-byte rows: 8-byte aligned only
    float* f = reinterpret_cast<float*>(p);
    const float x1 = __fdiv_rn(__fsub_rn(f[0], px), gain), y1 = __fdiv_rn(__fsub_rn(f[1], py), gain);
    const float x2 = __fdiv_rn(__fsub_rn(f[2], px), gain), y2 = __fdiv_rn(__fsub_rn(f[3], py), gain);
    f[0] = fminf(fmaxf(x1, 0.f), w0);
    f[1] = fminf(fmaxf(y1, 0.f), h0);
    f[2] = fminf(fmaxf(x2, 0.f), w0);
    f[3] = fminf(fmaxf(y2, 0.f), h0);
}

}  // namespace
}  // namespace fce

using namespace fce;

extern "C" int fce_detect_decode(const fce_decode_desc* d, const float* raw0, const float* raw1, const float* raw2,
                                 const float* raw3, float* y, void* stream) {
    if (!d || !raw0 || !y || d->B <= 0 || d->nl < 1 || d->nl > 4) return FCE_ERR_BAD_ARG;
    if (d->reg_max > MAX_REG || (d->reg_max % 4) || (d->nc % 4)) return FCE_ERR_UNSUPPORTED;
    const float* r[4] = {raw0, raw1, raw2, raw3};
    int A = 0;
    for (int i = 0; i < d->nl; ++i) {
        if (!r[i] || d->H[i] <= 0 || d->W[i] <= 0) return FCE_ERR_BAD_ARG;
        if ((d->raw_pitch[i] % 4) || (((uintptr_t)r[i]) & 15)) return FCE_ERR_ALIGNMENT;
        A += d->H[i] * d->W[i];
    }
    dim3 grid((A + NT - 1) / NT, d->B);
    decode_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(*d, raw0, raw1, raw2, raw3, y, A);
    return check_launch();
}

extern "C" int fce_scale_boxes(float* det, const int32_t* count, const float* meta, int32_t B, int32_t max_det,
                               void* stream) {
    if (!det || !count || !meta || B <= 0 || max_det <= 0) return FCE_ERR_BAD_ARG;
    if (B > 65535) return FCE_ERR_UNSUPPORTED;
    dim3 grid((max_det + NT - 1) / NT, B);
    scale_boxes_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(det, count, meta, max_det);
    return check_launch();
}

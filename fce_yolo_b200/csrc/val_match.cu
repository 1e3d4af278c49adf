// Validation matching on the GPU: box_iou (ultralytics/utils/metrics.py:57-77) + the non-scipy branch of
// BaseValidator.match_predictions (ultralytics/engine/validator.py:266-306) as DetectionValidator._process_batch
// (ultralytics/models/yolo/detect/val.py:274-288) applies them per image - in the reference a host-side numpy loop
// over ten IoU thresholds after a device->host copy of the IoU matrix.  Here one CTA per image works on the padded NMS
// output in place:
//   1. every detection finds its best same-class label (max fp32 IoU in the reference's operation order, ties to
//      the higher label index);
//   2. per threshold, every label takes the LOWEST-index detection among those whose best label it is and whose
//      IoU passes (what the reference's sort -> unique(detections) -> unique(labels) sequence computes);
//   3. tp[d][t] = 1 for those detections.
// Integer / compare work on a few KB per image: latency-bound, no tensor cores.
#include "common.cuh"

namespace fce {
namespace {

constexpr int MT = 256;

__global__ void __launch_bounds__(MT) match_kernel(const float* __restrict__ det, const int32_t* __restrict__ count,
                                                   const float* __restrict__ gt_boxes, const float* __restrict__ gt_cls,
                                                   const int32_t* __restrict__ gt_off, const float* __restrict__ iouv,
                                                   int max_det, int n_iou, int max_g, uint8_t* __restrict__ tp) {
    extern __shared__ __align__(16) unsigned char msm[];
    int* best_l = reinterpret_cast<int*>(msm);                 // [max_det]
    float* best_i = reinterpret_cast<float*>(best_l + max_det);  // [max_det]
    int* best_d = reinterpret_cast<int*>(best_i + max_det);      // [n_iou][max_g]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = min(count[b], max_det);
    const int g0 = gt_off[b], g = gt_off[b + 1] - g0;
    const float* db = det + (size_t)b * max_det * 6;
    uint8_t* tb = tp + (size_t)b * max_det * n_iou;
    for (int i = tid; i < n_iou * g; i += MT) best_d[(i / g) * max_g + (i % g)] = 0x7fffffff;
    for (int d = tid; d < n; d += MT) {
        const float px1 = db[d * 6], py1 = db[d * 6 + 1], px2 = db[d * 6 + 2], py2 = db[d * 6 + 3], pc = db[d * 6 + 5];
        const float area_p = __fmul_rn(__fsub_rn(px2, px1), __fsub_rn(py2, py1));
        float bi = -1.f;
        int bl = 0;
        for (int l = 0; l < g; ++l) {
            const float4 gb = *reinterpret_cast<const float4*>(gt_boxes + (size_t)(g0 + l) * 4);
            float v = 0.f;
            if (gt_cls[g0 + l] == pc) {  // iou * correct_class (validator.py:287-288)
                const float w = fmaxf(__fsub_rn(fminf(gb.z, px2), fmaxf(gb.x, px1)), 0.f);
                const float h = fmaxf(__fsub_rn(fminf(gb.w, py2), fmaxf(gb.y, py1)), 0.f);
                const float inter = __fmul_rn(w, h);
                const float area_g = __fmul_rn(__fsub_rn(gb.z, gb.x), __fsub_rn(gb.w, gb.y));
                const float uni = __fadd_rn(__fsub_rn(__fadd_rn(area_g, area_p), inter), 1e-7f);
                v = __fdiv_rn(inter, uni);
            }
            if (v >= bi) {  // ascending l with >= : ties go to the higher label index
                bi = v;
                bl = l;
            }
        }
        best_l[d] = bl;
        best_i[d] = bi;
    }
    __syncthreads();
    if (g > 0) {
        for (int d = tid; d < n; d += MT) {
            const float bi = best_i[d];
            const int bl = best_l[d];
            for (int t = 0; t < n_iou; ++t)
                if (bi >= iouv[t]) atomicMin(&best_d[t * max_g + bl], d);
        }
    }
    __syncthreads();
    for (int i = tid; i < max_det * n_iou; i += MT) {
        const int d = i / n_iou, t = i - d * n_iou;
        uint8_t ok = 0;
        if (d < n && g > 0) ok = (best_i[d] >= iouv[t] && best_d[t * max_g + best_l[d]] == d) ? 1 : 0;
        tb[i] = ok;
    }
}

}  // namespace
}  // namespace fce

using namespace fce;

extern "C" int fce_match_predictions(const float* det, const int32_t* count, const float* gt_boxes, const float* gt_cls,
                                     const int32_t* gt_offsets, const float* iouv, int32_t B, int32_t max_det,
                                     int32_t n_iou, int32_t max_gt_per_image, uint8_t* tp, void* stream) {
    if (!det || !count || !gt_offsets || !iouv || !tp || B <= 0 || max_det <= 0 || n_iou <= 0 || max_gt_per_image < 0)
        return FCE_ERR_BAD_ARG;
    if (max_gt_per_image > 0 && (!gt_boxes || !gt_cls)) return FCE_ERR_BAD_ARG;
    if (((uintptr_t)gt_boxes) & 15) return FCE_ERR_ALIGNMENT;
    const int mg = max_gt_per_image > 0 ? max_gt_per_image : 1;
    const size_t smem = (size_t)max_det * 8 + (size_t)n_iou * mg * 4;
    if (smem > 200 * 1024) return FCE_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_cuda_error(e);
            return FCE_ERR_CUDA;
        }
    }
    match_kernel<<<B, MT, smem, (cudaStream_t)stream>>>(det, count, gt_boxes, gt_cls, gt_offsets, iouv, max_det, n_iou, mg, tp);
    return check_launch();
}

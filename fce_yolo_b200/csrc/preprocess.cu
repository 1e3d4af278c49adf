// Predict-side preprocessing on the GPU: LetterBox (ultralytics/data/augment.py:1589-1631) + the
// BGR->RGB flip of BasePredictor.preprocess (ultralytics/engine/predictor.py:163-165) for a whole batch of
// differently sized uint8 HWC BGR images in ONE launch, written straight into the network's uint8 NHWC RGB input
// buffer (the 1/255 scale lives in the stem weights).  Bit-exact against cv2.resize(INTER_LINEAR) for 8-bit images:
// 11-bit fixed-point taps, vertical pass (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2 (OpenCV
// imgproc/resize.cpp, VResizeLinear for uchar).  The per-axis tap tables {i0, i1, w0, w1} are computed on the host
// (double -> float32 coordinate arithmetic identical to OpenCV's) and shipped with the batch.
// Integer / byte work, HBM-bound: source bytes read once (+ L1-served neighbours), 3 bytes written per pixel.
#include "common.cuh"

namespace fce {
namespace {

constexpr int LB_THREADS = 128, LB_PX = 4;  // pixels per thread: 12 output bytes = three aligned 32-bit stores

__global__ void __launch_bounds__(LB_THREADS) letterbox_kernel(const fce_letterbox_item* __restrict__ items,
                                                               const int4* __restrict__ xtab,
                                                               const int4* __restrict__ ytab, int out_h, int out_w,
                                                               uint8_t* __restrict__ out, int pad_value) {
    const int b = blockIdx.z, y = blockIdx.y;
    const int x0 = (blockIdx.x * LB_THREADS + threadIdx.x) * LB_PX;
    if (x0 >= out_w) return;
    const fce_letterbox_item it = items[b];
    const uint8_t* src = reinterpret_cast<const uint8_t*>(it.src);
    const int yy = y - it.top;
    const bool row_in = yy >= 0 && yy < it.new_h;
    int4 yt = make_int4(0, 0, 0, 0);
    if (row_in) yt = __ldg(ytab + (size_t)b * out_h + yy);
    const uint8_t* r0 = src + (size_t)yt.x * it.src_pitch;
    const uint8_t* r1 = src + (size_t)yt.y * it.src_pitch;
    uint8_t px[LB_PX * 3];
#pragma unroll
    for (int i = 0; i < LB_PX; ++i) {
        const int xx = x0 + i - it.left;
        if (row_in && xx >= 0 && xx < it.new_w) {
            const int4 xt = __ldg(xtab + (size_t)b * out_w + xx);
            const uint8_t* p00 = r0 + xt.x * 3;
            const uint8_t* p01 = r0 + xt.y * 3;
            const uint8_t* p10 = r1 + xt.x * 3;
            const uint8_t* p11 = r1 + xt.y * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int s0 = (int)__ldg(p00 + c) * xt.z + (int)__ldg(p01 + c) * xt.w;  // HResizeLinear, scale 2^11
                const int s1 = (int)__ldg(p10 + c) * xt.z + (int)__ldg(p11 + c) * xt.w;
                int v = (((yt.z * (s0 >> 4)) >> 16) + ((yt.w * (s1 >> 4)) >> 16) + 2) >> 2;
                v = v < 0 ? 0 : (v > 255 ? 255 : v);
                px[i * 3 + (2 - c)] = (uint8_t)v;  // BGR -> RGB
            }
        } else {
            px[i * 3] = px[i * 3 + 1] = px[i * 3 + 2] = (uint8_t)pad_value;
        }
    }
    uint8_t* o = out + (((size_t)b * out_h + y) * out_w + x0) * 3;
    if (x0 + LB_PX <= out_w && (out_w & 3) == 0) {
        uint32_t wds[3];
#pragma unroll
        for (int q = 0; q < 3; ++q)
            wds[q] = (uint32_t)px[4 * q] | ((uint32_t)px[4 * q + 1] << 8) | ((uint32_t)px[4 * q + 2] << 16) |
                     ((uint32_t)px[4 * q + 3] << 24);
        uint32_t* ow = reinterpret_cast<uint32_t*>(o);
        ow[0] = wds[0];
        ow[1] = wds[1];
        ow[2] = wds[2];
    } else {
        for (int i = 0; i < LB_PX && x0 + i < out_w; ++i) {
            o[i * 3] = px[i * 3];
            o[i * 3 + 1] = px[i * 3 + 1];
            o[i * 3 + 2] = px[i * 3 + 2];
        }
    }
}

}  // namespace
}  // namespace fce

using namespace fce;

extern "C" int fce_letterbox(const fce_letterbox_item* items, const int32_t* xtab, const int32_t* ytab, int32_t B,
                             int32_t out_h, int32_t out_w, int32_t pad_value, uint8_t* out, void* stream) {
    if (!items || !xtab || !ytab || !out || B <= 0 || out_h <= 0 || out_w <= 0) return FCE_ERR_BAD_ARG;
    if (B > 65535 || out_h > 65535) return FCE_ERR_UNSUPPORTED;
    if ((((uintptr_t)xtab | (uintptr_t)ytab) & 15) || ((uintptr_t)items & 7) || ((uintptr_t)out & 3)) return FCE_ERR_ALIGNMENT;
    dim3 grid((out_w + LB_THREADS * LB_PX - 1) / (LB_THREADS * LB_PX), out_h, B);
    letterbox_kernel<<<grid, LB_THREADS, 0, (cudaStream_t)stream>>>(items, reinterpret_cast<const int4*>(xtab),
                                                                    reinterpret_cast<const int4*>(ytab), out_h, out_w,
                                                                    out, pad_value);
    return check_launch();
}

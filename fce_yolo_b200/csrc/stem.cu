// Stem patch packing: the first Conv of the graph (3x3, stride 2, Cin = 3; reference conv.py:80-89 on the
// network input, yolo11-fce.yaml:20) has K = 27 - far too thin for an implicit-GEMM main loop and not
// addressable by TMA (3 channels = 3 or 12 bytes per pixel).  This kernel gathers each output pixel's 3x3x3
// patch into one 64-byte row [27 values | 5 zeros] of a bf16 matrix A[M = B*Ho*Wo, 32]; the stem then runs as a
// 1x1 convolution (K = 32) on the tcgen05 kernel with the reference's OHWI weights laid out the same way.
// uint8 images are stored unscaled (0..255 are exact in bf16) - the 1/255 is folded into the weights.
// HBM-bound: algorithmic bytes per output pixel = 27 * e_in / 4 (input, each byte used ~2.25 times) + 64 written.
#include "common.cuh"

namespace fce {
namespace {

constexpr int NT = 256;

template <typename TI, int LAYOUT>
__global__ void __launch_bounds__(NT) stem_pack_kernel(const fce_pack_desc d, const TI* __restrict__ x,
                                                       uint4* __restrict__ a, int Ho, int Wo, int M) {
    const int m = blockIdx.x * NT + threadIdx.x;
    if (m >= M) return;
    const int hw = Ho * Wo;
    const int b = m / hw;
    const int rem = m - b * hw;
    const int ho = rem / Wo, wo = rem - ho * Wo;
    const int h0 = ho * 2 - 1, w0 = wo * 2 - 1;
    float v[27];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int hi = h0 + kh;
        const bool h_ok = hi >= 0 && hi < d.H;
        const int hc = h_ok ? hi : 0;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int wi = w0 + kw;
            const bool ok = h_ok && wi >= 0 && wi < d.W;
            const int wc = ok ? wi : 0;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                size_t idx;
                if (LAYOUT == FCE_NHWC)
                    idx = ((size_t)(b * d.H + hc) * d.W + wc) * 3 + ci;
                else
                    idx = ((size_t)(b * 3 + ci) * d.H + hc) * d.W + wc;
                const float t = Elem<TI>::to_f(__ldg(x + idx));
                v[(kh * 3 + kw) * 3 + ci] = ok ? t : 0.f;
            }
        }
    }
    uint32_t o[16];
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        o[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[26], 0.f);
        o[13] = *reinterpret_cast<uint32_t*>(&h);
    }
    o[14] = o[15] = 0u;
    uint4* dst = a + (size_t)m * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
}

}  // namespace
}  // namespace fce

using namespace fce;

extern "C" int fce_stem_pack(const fce_pack_desc* d, const void* x, void* a, void* stream) {
    if (!d || !x || !a || d->B <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    if (d->Cin != 3 || d->k != 3 || d->stride != 2 || d->Kpad != 32) return FCE_ERR_UNSUPPORTED;
    if (((uintptr_t)a) & 15) return FCE_ERR_ALIGNMENT;
    const int Ho = (d->H + 2 - 3) / 2 + 1, Wo = (d->W + 2 - 3) / 2 + 1;
    const long long M = (long long)d->B * Ho * Wo;
    if (M > 0x7fffffffLL) return FCE_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)((M + NT - 1) / NT);
    if (d->in_dtype == FCE_U8 && d->in_layout == FCE_NHWC)
        stem_pack_kernel<uint8_t, FCE_NHWC><<<grid, NT, 0, st>>>(*d, (const uint8_t*)x, (uint4*)a, Ho, Wo, (int)M);
    else if (d->in_dtype == FCE_F32 && d->in_layout == FCE_NCHW)
        stem_pack_kernel<float, FCE_NCHW><<<grid, NT, 0, st>>>(*d, (const float*)x, (uint4*)a, Ho, Wo, (int)M);
    else if (d->in_dtype == FCE_F32 && d->in_layout == FCE_NHWC)
        stem_pack_kernel<float, FCE_NHWC><<<grid, NT, 0, st>>>(*d, (const float*)x, (uint4*)a, Ho, Wo, (int)M);
    else
        return FCE_ERR_UNSUPPORTED;
    return check_launch();
}

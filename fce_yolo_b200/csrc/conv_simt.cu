// Dense convolution, fp32-accurate SIMT implicit GEMM (CUDA cores, fp32 FMA).
//
// Role: (1) the fp32-mode path (north_star: 1e-4 relative per layer, which rules out tensor-core
// tf32/bf16 inputs), (2) shapes the tcgen05 kernel does not take (Cin = 3 stem, NCHW/u8 image input,
// channel counts that are not multiples of 8).  The bf16 hot path is conv_tc.cu.
//
// GEMM view: M = B*Ho*Wo output pixels, N = Cout, K = k*k*Cin with K index = (kh*k + kw)*Cin + ci,
// matching the OHWI weight layout.  64x64 CTA tile, BK = 16, 256 threads, 4x4 register tile.
#include "common.cuh"

namespace fce {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename TI, typename TW, typename TO, int LAYOUT>
__global__ void __launch_bounds__(NT) conv_simt_kernel(const fce_conv_desc d, const TI* __restrict__ x,
                                                       const TW* __restrict__ w, const float* __restrict__ bias,
                                                       const TO* res, TO* y, int Ho, int Wo) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int M = d.B * Ho * Wo;
    const int K = d.k * d.k * d.Cin;
    const int pad = d.k >> 1;
    const int m_base = blockIdx.x * BM, n_base = blockIdx.y * BN;

    // loader role: row (pixel for A / cout for B) lr, 4 consecutive k starting at lk
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    const int gm = m_base + lr;
    const bool m_ok = gm < M;
    int pb = 0, ph = 0, pw = 0;
    if (m_ok) {
        pb = gm / (Ho * Wo);
        int r = gm - pb * Ho * Wo;
        ph = r / Wo;
        pw = r - ph * Wo;
    }
    const int h0 = ph * d.stride - pad, w0 = pw * d.stride - pad;
    const int gn = n_base + lr;
    const bool n_ok = gn < d.Cout;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        // ---- A tile: im2col gather ----
        {
            int kidx = k0 + lk;
            int tap = kidx / d.Cin;
            int c = kidx - tap * d.Cin;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = 0.f;
                if (m_ok && kidx + j < K) {
                    int kh = tap / d.k, kw = tap - kh * d.k;
                    int hi = h0 + kh, wi = w0 + kw;
                    if (hi >= 0 && hi < d.H && wi >= 0 && wi < d.W) {
                        size_t idx;
                        if (LAYOUT == FCE_NHWC)
                            idx = ((size_t)(pb * d.H + hi) * d.W + wi) * d.in_pitch + d.in_off + c;
                        else
                            idx = ((size_t)(pb * d.Cin + c) * d.H + hi) * d.W + wi;
                        v = Elem<TI>::to_f(x[idx]) * d.in_scale;
                    }
                }
                As[lk + j][lr] = v;
                if (++c == d.Cin) { c = 0; ++tap; }
            }
        }
        // ---- B tile: weights [Cout][K] ----
        {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int kidx = k0 + lk + j;
                float v = 0.f;
                if (n_ok && kidx < K) v = Elem<TW>::to_f(w[(size_t)gn * K + kidx]);
                Bs[lk + j][lr] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue: bias, activation, residual, store ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m_base + ty * 4 + i;
        if (m >= M) continue;
        const size_t obase = (size_t)m * d.out_pitch + d.out_off;
        const size_t rbase = (size_t)m * d.res_pitch + d.res_off;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n_base + tx * 4 + j;
            if (n >= d.Cout) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (d.act == FCE_ACT_SILU) v = silu_acc(v);
            else if (d.act == FCE_ACT_SIGMOID) v = sigmoid_acc(v);
            if (res) v += Elem<TO>::to_f(res[rbase + n]);
            y[obase + n] = Elem<TO>::from_f(v);
        }
    }
}

template <typename TI, typename TW, typename TO>
int launch_simt(const fce_conv_desc* d, const void* x, const void* w, const float* bias, const void* res, void* y,
                cudaStream_t st) {
    const int pad = d->k / 2;
    const int Ho = (d->H + 2 * pad - d->k) / d->stride + 1;
    const int Wo = (d->W + 2 * pad - d->k) / d->stride + 1;
    const long long M = (long long)d->B * Ho * Wo;
    dim3 grid(ceil_div(M, BM), ceil_div(d->Cout, BN));
    if (d->in_layout == FCE_NHWC)
        conv_simt_kernel<TI, TW, TO, FCE_NHWC><<<grid, NT, 0, st>>>(*d, (const TI*)x, (const TW*)w, bias,
                                                                    (const TO*)res, (TO*)y, Ho, Wo);
    else
        conv_simt_kernel<TI, TW, TO, FCE_NCHW><<<grid, NT, 0, st>>>(*d, (const TI*)x, (const TW*)w, bias,
                                                                    (const TO*)res, (TO*)y, Ho, Wo);
    return check_launch();
}

}  // namespace

int conv2d_simt(const fce_conv_desc* d, const void* x, const void* w, const float* bias, const void* res, void* y,
                cudaStream_t st) {
    using bf = __nv_bfloat16;
    const int key = d->in_dtype * 100 + d->w_dtype * 10 + d->out_dtype;
    switch (key) {
        case FCE_F32 * 100 + FCE_F32 * 10 + FCE_F32: return launch_simt<float, float, float>(d, x, w, bias, res, y, st);
        case FCE_BF16 * 100 + FCE_BF16 * 10 + FCE_BF16: return launch_simt<bf, bf, bf>(d, x, w, bias, res, y, st);
        case FCE_BF16 * 100 + FCE_BF16 * 10 + FCE_F32: return launch_simt<bf, bf, float>(d, x, w, bias, res, y, st);
        case FCE_F32 * 100 + FCE_BF16 * 10 + FCE_BF16: return launch_simt<float, bf, bf>(d, x, w, bias, res, y, st);
        case FCE_U8 * 100 + FCE_BF16 * 10 + FCE_BF16: return launch_simt<uint8_t, bf, bf>(d, x, w, bias, res, y, st);
        case FCE_U8 * 100 + FCE_F32 * 10 + FCE_F32: return launch_simt<uint8_t, float, float>(d, x, w, bias, res, y, st);
        default: return FCE_ERR_UNSUPPORTED;
    }
}

}  // namespace fce

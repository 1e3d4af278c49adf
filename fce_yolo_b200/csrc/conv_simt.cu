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


// ------------------------------------------------------------------------------------------------ strip GEMM
// 1x1 convolutions on the pooled strips of the coordinate-attention family (fce_block.py:105,112-113,165-168,178,
// 246-249,258,264-268,276): y[M, N] = act(x[M, K] w[N, K]^T + b) with M = B*(H+W) or B*H rows, K and N between 8
// and a few hundred, everything fp32 (strips, gates and their weights are fp32 in both modes).  The generic
// kernel above spends its time in im2col index arithmetic; this one is a plain tiled SGEMM with 16-byte loads:
// 32 x 64 CTA tile, BK = 32, 256 threads x (2 x 4) outputs.  HBM-bound on the strip itself (a few MB).
constexpr int SG_BM = 32, SG_BN = 64, SG_BK = 32;

template <bool VEC>
__global__ void __launch_bounds__(NT) strip_gemm_kernel(const fce_conv_desc d, const float* __restrict__ x,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        const float* res, float* y, int M) {
    __shared__ float As[SG_BK][SG_BM + 1];
    __shared__ float Bs[SG_BK][SG_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // outputs: rows ty*2 .. +1, cols tx*4 .. +3
    const int K = d.Cin, N = d.Cout;
    const int m_base = blockIdx.x * SG_BM, n_base = blockIdx.y * SG_BN;
    const int ar = tid >> 3, ak = (tid & 7) * 4;  // A loader: row ar, 4 consecutive k
    const float* arow = x + (size_t)(m_base + ar) * d.in_pitch + d.in_off;
    const bool a_ok = m_base + ar < M;
    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // register double buffering: the global loads of K chunk k0 + BK are in flight while chunk k0 is multiplied
    float va[4], vb[2][4];
    auto fetch = [&](int k0) {
        const int k = k0 + ak;
#pragma unroll
        for (int j = 0; j < 4; ++j) va[j] = 0.f;
        if (a_ok) {
            if (VEC && k + 3 < K) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(arow + k));
                va[0] = t.x; va[1] = t.y; va[2] = t.z; va[3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (k + j < K) va[j] = __ldg(arow + k + j);
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // B loader: rows (tid>>3) + 32*h, 4 consecutive k
            const int n = n_base + (tid >> 3) + 32 * h;
#pragma unroll
            for (int j = 0; j < 4; ++j) vb[h][j] = 0.f;
            if (n < N) {
                const float* wr = w + (size_t)n * K + k;
                if (VEC && k + 3 < K) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(wr));
                    vb[h][0] = t.x; vb[h][1] = t.y; vb[h][2] = t.z; vb[h][3] = t.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (k + j < K) vb[h][j] = __ldg(wr + j);
                }
            }
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[ak + j][ar] = va[j];
            Bs[ak + j][tid >> 3] = vb[0][j];
            Bs[ak + j][(tid >> 3) + 32] = vb[1][j];
        }
        __syncthreads();
        if (k0 + SG_BK < K) fetch(k0 + SG_BK);
#pragma unroll
        for (int kk = 0; kk < SG_BK; ++kk) {
            const float a0 = As[kk][ty * 2], a1 = As[kk][ty * 2 + 1];
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
            acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
            acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
            acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m_base + ty * 2 + i;
        if (m >= M) continue;
        float* yo = y + (size_t)m * d.out_pitch + d.out_off;
        const float* ro = res ? res + (size_t)m * d.res_pitch + d.res_off : nullptr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n_base + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            v = apply_act(v, d.act);
            if (ro) v += ro[n];
            yo[n] = v;
        }
    }
}

int launch_strip_gemm(const fce_conv_desc* d, const void* x, const void* w, const float* bias, const void* res, void* y,
                      cudaStream_t st) {
    const long long M = (long long)d->B * d->H * d->W;
    if (M > 0x7fffffffLL) return FCE_ERR_UNSUPPORTED;
    dim3 grid(ceil_div(M, SG_BM), ceil_div(d->Cout, SG_BN));
    const bool vec = d->Cin % 4 == 0 && d->in_pitch % 4 == 0 && d->in_off % 4 == 0 && (((uintptr_t)x | (uintptr_t)w) & 15) == 0;
    if (vec)
        strip_gemm_kernel<true><<<grid, NT, 0, st>>>(*d, (const float*)x, (const float*)w, bias, (const float*)res,
                                                     (float*)y, (int)M);
    else
        strip_gemm_kernel<false><<<grid, NT, 0, st>>>(*d, (const float*)x, (const float*)w, bias, (const float*)res,
                                                      (float*)y, (int)M);
    return check_launch();
}

}  // namespace

int conv2d_simt(const fce_conv_desc* d, const void* x, const void* w, const float* bias, const void* res, void* y,
                cudaStream_t st) {
    using bf = __nv_bfloat16;
    const int key = d->in_dtype * 100 + d->w_dtype * 10 + d->out_dtype;
    if (key == FCE_F32 * 100 + FCE_F32 * 10 + FCE_F32 && d->k == 1 && d->stride == 1 && d->in_layout == FCE_NHWC &&
        d->in_scale == 1.0f)
        return launch_strip_gemm(d, x, w, bias, res, y, st);  // strips, and every 1x1 of fp32 mode
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

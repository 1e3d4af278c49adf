// C2PSA self-attention core (reference block.py:1293-1302), flash-style: a CTA owns 64 queries of one
// (image, head), streams 64-key tiles through shared memory with an online softmax, and never
// materialises the N x N score matrix.  key_dim = 32, head_dim = 64 always hold on this path
// (heads = c/64, attn_ratio = 0.5).  fp32 CUDA-core math (this is also the fp32-mode kernel).
#include <atomic>

#include "common.cuh"

namespace fce {
namespace {

constexpr int KD = 32, HD = 64, BR = 64, BC = 64, NT = 256;

template <typename T>
__global__ void __launch_bounds__(NT) psa_kernel(const fce_psa_desc d, const T* __restrict__ qkv, T* __restrict__ out) {
    extern __shared__ __align__(16) float psa_smem[];
    float (*Qs)[BR + 4] = reinterpret_cast<float (*)[BR + 4]>(psa_smem);                 // [k][query]
    float (*Ks)[BC + 4] = reinterpret_cast<float (*)[BC + 4]>(psa_smem + KD * (BR + 4)); // [k][key]
    float (*Vs)[HD + 4] = reinterpret_cast<float (*)[HD + 4]>(psa_smem + KD * (BR + 4) + KD * (BC + 4));  // [key][d]
    float (*Ps)[BC + 4] = reinterpret_cast<float (*)[BC + 4]>(psa_smem + KD * (BR + 4) + KD * (BC + 4) + BC * (HD + 4));
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * BR;
    const T* base = qkv + (size_t)b * d.N * d.qkv_pitch;
    const int qo = d.q_off + head * KD, ko = d.k_off + head * KD, vo = d.v_off + head * HD;

    for (int i = tid; i < BR * KD; i += NT) {
        const int r = i / KD, c = i % KD;
        const int n = q0 + r;
        Qs[c][r] = n < d.N ? Elem<T>::to_f(base[(size_t)n * d.qkv_pitch + qo + c]) * d.scale : 0.f;
    }
    float m_i[4], l_i[4], o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_i[i] = -INFINITY;
        l_i[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    }

    for (int k0 = 0; k0 < d.N; k0 += BC) {
        __syncthreads();  // previous tile fully consumed (also orders the Qs fill on the first pass)
        for (int i = tid; i < BC * KD; i += NT) {
            const int r = i / KD, c = i % KD;
            const int n = k0 + r;
            Ks[c][r] = n < d.N ? Elem<T>::to_f(base[(size_t)n * d.qkv_pitch + ko + c]) : 0.f;
        }
        for (int i = tid; i < BC * HD; i += NT) {
            const int r = i / HD, c = i % HD;
            const int n = k0 + r;
            Vs[r][c] = n < d.N ? Elem<T>::to_f(base[(size_t)n * d.qkv_pitch + vo + c]) : 0.f;
        }
        __syncthreads();
        // S = (scale*Q) K^T : this thread owns queries ty*4.., keys tx*4..
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
        for (int kk = 0; kk < KD; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&Qs[kk][ty * 4]);
            const float4 c = *reinterpret_cast<const float4*>(&Ks[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], cv[j], s[i][j]);
        }
        // online softmax per query row; the 16 threads sharing ty sit in one half-warp
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k0 + tx * 4 + j >= d.N) s[i][j] = -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float m_new = fmaxf(m_i[i], mx);
            const float alpha = expf(m_i[i] - m_new);  // first tile: exp(-inf) = 0
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = expf(s[i][j] - m_new);
                Ps[ty * 4 + i][tx * 4 + j] = p;
                rs += p;
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
            l_i[i] = l_i[i] * alpha + rs;
            m_i[i] = m_new;
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] *= alpha;
        }
        __syncthreads();
        // O += P V : queries ty*4.., output dims tx*4..
#pragma unroll 8
        for (int kk = 0; kk < BC; ++kk) {
            const float4 vv = *reinterpret_cast<const float4*>(&Vs[kk][tx * 4]);
            const float vvv[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float p = Ps[ty * 4 + i][kk];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[i][j] = fmaf(p, vvv[j], o[i][j]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = q0 + ty * 4 + i;
        if (n >= d.N) continue;
        const float inv = 1.f / l_i[i];
        T* op = out + ((size_t)b * d.N + n) * d.out_pitch + d.out_off + head * HD + tx * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) op[j] = Elem<T>::from_f(o[i][j] * inv);
    }
}

}  // namespace
}  // namespace fce

using namespace fce;

extern "C" int fce_psa_attention(const fce_psa_desc* d, const void* qkv, void* out, void* stream) {
    if (!d || !qkv || !out || d->B <= 0 || d->N <= 0 || d->heads <= 0) return FCE_ERR_BAD_ARG;
    if (d->kd != KD || d->hd != HD) return FCE_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((d->N + BR - 1) / BR, d->heads, d->B);
    constexpr size_t smem = sizeof(float) * (KD * (BR + 4) + KD * (BC + 4) + BC * (HD + 4) + BR * (BC + 4));
    static_assert(smem <= 100 * 1024, "psa smem");
    static std::atomic<bool> attr_done{false};  // set once per process (not during graph capture replays)
    if (!attr_done.load(std::memory_order_acquire)) {
        cudaError_t e1 = cudaFuncSetAttribute(psa_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaError_t e2 = cudaFuncSetAttribute(psa_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { set_cuda_error(e1 != cudaSuccess ? e1 : e2); return FCE_ERR_CUDA; }
        attr_done.store(true, std::memory_order_release);
    }
    if (d->dtype == FCE_BF16)
        psa_kernel<__nv_bfloat16><<<grid, NT, smem, st>>>(*d, (const __nv_bfloat16*)qkv, (__nv_bfloat16*)out);
    else if (d->dtype == FCE_F32)
        psa_kernel<float><<<grid, NT, smem, st>>>(*d, (const float*)qkv, (float*)out);
    else
        return FCE_ERR_UNSUPPORTED;
    return check_launch();
}

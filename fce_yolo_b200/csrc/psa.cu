// C2PSA self-attention core (reference block.py:1293-1302), flash-style: a CTA owns 64 queries of one
// (image, head), streams 64-key tiles through shared memory with an online softmax, and never
// materialises the N x N score matrix.  key_dim = 32, head_dim = 64 always hold on this path
// (heads = c/64, attn_ratio = 0.5).
//   psa_kernel     : fp32 CUDA-core math - the fp32-mode (1e-4 parity) kernel.
//   psa_mma_kernel : bf16 mode.  Warp-level tensor-core MMA (mma.sync m16n8k16, fp32 accumulate): per (image,
//                    head) the problem is 400 x 400 x (32 + 64) at 640^2 - far too small and too softmax-heavy for
//                    a tcgen05/TMEM pipeline to pay; a warp owns 16 queries, S = Q K^T stays in registers, the
//                    probabilities are re-packed in registers as the A operand of P V (no shared-memory round
//                    trip), K/V tiles are double-buffered with cp.async.
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


// ------------------------------------------------------------------------------------------------ bf16 / mma.sync
constexpr int MQ = 64;            // queries per CTA (4 warps x 16)
constexpr int MK = 80;            // keys per tile: 400 = 5 x 80, 1600 = 20 x 80 (no ragged tile at 640^2 / 1280^2)
constexpr int MNT = 128;
constexpr int QK_PITCH = KD + 8;  // bf16 elements; 80-byte rows -> the 8 rows of an ldmatrix hit 8 distinct 16-B slots
constexpr int V_PITCH = HD + 8;   // 144-byte rows, same property

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;  // src-size 0 -> 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(MNT) psa_mma_kernel(const fce_psa_desc d, const __nv_bfloat16* __restrict__ qkv,
                                                      __nv_bfloat16* __restrict__ out) {
    pdl_trigger();
    __shared__ __align__(16) __nv_bfloat16 Qs[MQ * QK_PITCH];
    __shared__ __align__(16) __nv_bfloat16 Ks[2][MK * QK_PITCH];
    __shared__ __align__(16) __nv_bfloat16 Vs[2][MK * V_PITCH];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * MQ;
    const __nv_bfloat16* base = qkv + (size_t)b * d.N * d.qkv_pitch;
    const int qo = d.q_off + head * KD, ko = d.k_off + head * KD, vo = d.v_off + head * HD;
    const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(Qs);
    const uint32_t sK = (uint32_t)__cvta_generic_to_shared(Ks), sV = (uint32_t)__cvta_generic_to_shared(Vs);

    auto load_kv = [&](int tile, int buf) {
        const int k0 = tile * MK;
        for (int i = tid; i < MK * (KD / 8); i += MNT) {  // 4 x 16-B chunks per key row
            const int r = i >> 2, c = i & 3, n = k0 + r;
            const bool ok = n < d.N;
            cp_async16(sK + (uint32_t)((buf * MK + r) * QK_PITCH + c * 8) * 2,
                       base + (size_t)(ok ? n : 0) * d.qkv_pitch + ko + c * 8, ok);
        }
        for (int i = tid; i < MK * (HD / 8); i += MNT) {  // 8 chunks per value row
            const int r = i >> 3, c = i & 7, n = k0 + r;
            const bool ok = n < d.N;
            cp_async16(sV + (uint32_t)((buf * MK + r) * V_PITCH + c * 8) * 2,
                       base + (size_t)(ok ? n : 0) * d.qkv_pitch + vo + c * 8, ok);
        }
    };
    for (int i = tid; i < MQ * (KD / 8); i += MNT) {
        const int r = i >> 2, c = i & 3, n = q0 + r;
        const bool ok = n < d.N;
        cp_async16(sQ + (uint32_t)(r * QK_PITCH + c * 8) * 2, base + (size_t)(ok ? n : 0) * d.qkv_pitch + qo + c * 8, ok);
    }
    load_kv(0, 0);
    asm volatile("cp.async.commit_group;");

    const int tiles = (d.N + MK - 1) / MK;
    const float sc = d.scale * 1.4426950408889634f;  // softmax in base 2
    uint32_t qa[2][4];
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // rows g and g+8 of this warp's 16 queries

    for (int tile = 0; tile < tiles; ++tile) {
        const int buf = tile & 1;
        if (tile + 1 < tiles) {
            load_kv(tile + 1, buf ^ 1);
            asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 1;");
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncthreads();
        if (tile == 0) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
                ldsm_x4(sQ + (uint32_t)((warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * QK_PITCH + ks * 16 + 8 * (lane >> 4)) * 2,
                        qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
        }
        // S = Q K^T for 16 queries x 80 keys
        float s[MK / 8][4];
#pragma unroll
        for (int nt = 0; nt < MK / 8; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
            uint32_t b0, b1, b2, b3;
            ldsm_x4(sK + (uint32_t)((buf * MK + nt * 8 + (lane & 7)) * QK_PITCH + (lane >> 3) * 8) * 2, b0, b1, b2, b3);
            mma_bf16(s[nt], qa[0][0], qa[0][1], qa[0][2], qa[0][3], b0, b1);
            mma_bf16(s[nt], qa[1][0], qa[1][1], qa[1][2], qa[1][3], b2, b3);
        }
        const int k0 = tile * MK;
        const bool ragged = k0 + MK > d.N;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < MK / 8; ++nt) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = s[nt][j] * sc;
                if (ragged && k0 + nt * 8 + 2 * t + (j & 1) >= d.N) v = -INFINITY;
                s[nt][j] = v;
            }
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);  // finite: every tile holds at least one valid key
        const float a0 = exp2f(m0 - n0), a1 = exp2f(m1 - n1);
        m0 = n0;
        m1 = n1;
        float r0 = 0.f, r1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < MK / 8; ++nt) {
            s[nt][0] = exp2f(s[nt][0] - n0);
            s[nt][1] = exp2f(s[nt][1] - n0);
            s[nt][2] = exp2f(s[nt][2] - n1);
            s[nt][3] = exp2f(s[nt][3] - n1);
            r0 += s[nt][0] + s[nt][1];
            r1 += s[nt][2] + s[nt][3];
        }
        l0 = l0 * a0 + r0;  // per-thread partial sums; reduced over the quad once at the end
        l1 = l1 * a1 + r1;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
            o[i][0] *= a0;
            o[i][1] *= a0;
            o[i][2] *= a1;
            o[i][3] *= a1;
        }
        // O += P V : the S accumulators of key tiles 2j, 2j+1 are exactly the A fragment of k-step j
#pragma unroll
        for (int j = 0; j < MK / 16; ++j) {
            const uint32_t p0 = pack_bf16(s[2 * j][0], s[2 * j][1]), p1 = pack_bf16(s[2 * j][2], s[2 * j][3]);
            const uint32_t p2 = pack_bf16(s[2 * j + 1][0], s[2 * j + 1][1]), p3 = pack_bf16(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
            for (int i = 0; i < HD / 8; i += 2) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(sV + (uint32_t)((buf * MK + j * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * V_PITCH + 8 * (i + (lane >> 4))) * 2,
                          b0, b1, b2, b3);
                mma_bf16(o[i], p0, p1, p2, p3, b0, b1);
                mma_bf16(o[i + 1], p0, p1, p2, p3, b2, b3);
            }
        }
        __syncthreads();  // tile fully consumed before its buffer is refilled
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    __nv_bfloat16* op = out + (size_t)b * d.N * d.out_pitch + d.out_off + head * HD + 2 * t;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
        if (r0 < d.N) *reinterpret_cast<uint32_t*>(op + (size_t)r0 * d.out_pitch + i * 8) = pack_bf16(o[i][0] * i0, o[i][1] * i0);
        if (r1 < d.N) *reinterpret_cast<uint32_t*>(op + (size_t)r1 * d.out_pitch + i * 8) = pack_bf16(o[i][2] * i1, o[i][3] * i1);
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
    static DeviceOnce attr_once;  // once per DEVICE (the attribute is per device; never during graph capture replays)
    int dev_ = 0;
    if (attr_once.pending(&dev_)) {
        cudaError_t e1 = cudaFuncSetAttribute(psa_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaError_t e2 = cudaFuncSetAttribute(psa_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { set_cuda_error(e1 != cudaSuccess ? e1 : e2); return FCE_ERR_CUDA; }
        attr_once.done(dev_);
    }
    const bool vec_ok = d->qkv_pitch % 8 == 0 && d->q_off % 8 == 0 && d->k_off % 8 == 0 && d->v_off % 8 == 0 &&
                        d->out_pitch % 2 == 0 && d->out_off % 2 == 0 && (uintptr_t)qkv % 16 == 0 && (uintptr_t)out % 4 == 0;
    if (d->dtype == FCE_BF16 && vec_ok) {
        dim3 mgrid((d->N + MQ - 1) / MQ, d->heads, d->B);
        psa_mma_kernel<<<mgrid, MNT, 0, st>>>(*d, (const __nv_bfloat16*)qkv, (__nv_bfloat16*)out);
    } else if (d->dtype == FCE_BF16)
        psa_kernel<__nv_bfloat16><<<grid, NT, smem, st>>>(*d, (const __nv_bfloat16*)qkv, (__nv_bfloat16*)out);
    else if (d->dtype == FCE_F32)
        psa_kernel<float><<<grid, NT, smem, st>>>(*d, (const float*)qkv, (float*)out);
    else
        return FCE_ERR_UNSUPPORTED;
    return check_launch();
}

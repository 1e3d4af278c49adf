// Dense convolution as an NHWC implicit GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Replaces the cuDNN call behind Conv.forward_fuse (ultralytics/nn/modules/conv.py:80-89) and every bare
// nn.Conv2d on the bf16 path, with bias + SiLU/sigmoid + residual + concat-offset epilogues fused.
//
// GEMM view:  D[M, N] = A[M, K] * B[N, K]^T,  M = B*Ho*Wo output pixels, N = Cout, K = k*k*Cin with
// K index = (kh*k + kw)*Cin + ci (OHWI weights), both operands K-major in shared memory.
//
//   * A (activations) is never materialised: for 1x1 convs one 2-D tiled TMA box {kc channels, 128 pixels}
//     of the [M, pitch] view; for 3x3 convs one IM2COL-mode TMA per (tap, channel chunk): the TMA unit walks
//     128 consecutive output pixels (across rows and images), applies the conv stride and the (kw, kh) tap
//     offset and zero-fills the padding halo.  The input map is read from HBM once; the 9 taps hit L2.
//   * B (weights) is a 2-D tiled box {kc, bn} of the [Cout, K] matrix.
//   * One CTA per SM, persistent over (m, n) tiles, warp-specialised:
//       warp 0    : TMA producer (one elected lane), `stages`-deep full/empty mbarrier ring
//       warp 1    : TMEM allocator + single-thread tcgen05.mma issuer; the accumulator (128 x bn fp32) lives in
//                   TMEM and is double buffered so the epilogue of tile i overlaps the main loop of tile i+1
//       warps 2-9 : epilogue: tcgen05.ld -> +bias -> SiLU -> (+residual) -> bf16/fp32 -> NHWC store at the
//                   channel offset of the destination view (concat fusion)
//   * Shared-memory tiles use the hardware swizzle that matches the K chunk (128B / 64B / 32B for
//     kc = 64 / 32 / 16 channels), identical in the TMA descriptor and the UMMA shared-memory descriptor.
#include <cuda.h>  // CUtensorMap types only - the two driver entry points are resolved at run time via cudart

#include "common.cuh"

namespace fce {
namespace {

constexpr int BM = 128;           // UMMA M (cta_group::1)
constexpr int NUM_EPI_WARPS = 8;  // two warps per TMEM lane quarter, interleaved over 16-column chunks
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 200 * 1024;

struct TcParams {
    int M, Cout, Cin;
    int taps, ksz, stride, pad;
    int Ho, Wo;
    int kc, chunks;  // channels per K step, K steps per tap
    int bn, n_tiles, m_tiles;
    int stages;
    uint32_t a_bytes, b_bytes;    // TMA bytes per stage
    uint32_t a_stride, b_stride;  // stage strides in shared memory (1024-aligned)
    uint32_t tmem_cols;
    int out_pitch, res_pitch, act, out_f32;
    uint32_t desc_hi;  // upper 32 bits of the UMMA shared-memory descriptor (SBO, version, swizzle)
    uint32_t idesc;    // UMMA instruction descriptor
};

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must fault the launch (reported through the C ABI), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w, int h,
                                                int n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulation, issued by ONE thread for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on the mbarrier once every MMA issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct Ring {
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance(int n) {
        if (++stage == n) {
            stage = 0;
            phase ^= 1;
        }
    }
};

__device__ __forceinline__ float act_fast(float v, int act) {
    if (act == FCE_ACT_SILU) return __fdividef(v, 1.f + __expf(-v));
    if (act == FCE_ACT_SIGMOID) return __fdividef(1.f, 1.f + __expf(-v));
    return v;
}

// ------------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p,
               const float* __restrict__ bias, const __nv_bfloat16* __restrict__ res, void* __restrict__ y) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base;
    const uint32_t sB = sA + p.stages * p.a_stride;
    const uint32_t bars = sB + p.stages * p.b_stride;
    const uint32_t full0 = bars, empty0 = bars + 8 * MAX_STAGES;
    const uint32_t tfull0 = bars + 16 * MAX_STAGES, tempty0 = tfull0 + 16;
    const uint32_t tmem_slot = tempty0 + 16;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles;
    const int k_iters = p.taps * p.chunks;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, NUM_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            Ring r;
            const int hw = p.Ho * p.Wo;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m0 = (tile / p.n_tiles) * BM, n0 = (tile % p.n_tiles) * p.bn;
                int img = 0, w0 = 0, h0 = 0;
                if (p.taps > 1) {
                    img = m0 / hw;
                    const int rem = m0 - img * hw;
                    const int ph = rem / p.Wo;
                    h0 = ph * p.stride - p.pad;
                    w0 = (rem - ph * p.Wo) * p.stride - p.pad;
                }
                for (int tap = 0; tap < p.taps; ++tap) {
                    const int kh = tap / p.ksz, kw = tap - kh * p.ksz;
                    for (int ch = 0; ch < p.chunks; ++ch) {
                        const uint32_t fb = full0 + 8 * r.stage;
                        mbar_wait(empty0 + 8 * r.stage, r.phase ^ 1);
                        mbar_expect_tx(fb, p.a_bytes + p.b_bytes);
                        if (p.taps == 1)
                            tma_load_2d(sA + r.stage * p.a_stride, &tmA, fb, ch * p.kc, m0);
                        else
                            tma_load_im2col(sA + r.stage * p.a_stride, &tmA, fb, ch * p.kc, w0, h0, img, (uint16_t)kw,
                                            (uint16_t)kh);
                        tma_load_2d(sB + r.stage * p.b_stride, &tmB, fb, tap * p.Cin + ch * p.kc, n0);
                        r.advance(p.stages);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            Ring r;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint64_t hi = (uint64_t)p.desc_hi << 32;
            const int kk = p.kc >> 4;  // UMMA K = 16 bf16 = 32 bytes
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * p.bn;
                for (int it = 0; it < k_iters; ++it) {
                    mbar_wait(full0 + 8 * r.stage, r.phase);
                    tc_fence_after();
                    const uint32_t a_addr = sA + r.stage * p.a_stride, b_addr = sB + r.stage * p.b_stride;
                    for (int k = 0; k < kk; ++k) {
                        const uint64_t ad = hi | (uint64_t)(((a_addr + 32 * k) >> 4) & 0x3FFF) | (1ull << 16);
                        const uint64_t bd = hi | (uint64_t)(((b_addr + 32 * k) >> 4) & 0x3FFF) | (1ull << 16);
                        umma_bf16(d_tmem, ad, bd, p.idesc, (it | k) != 0);
                    }
                    umma_commit(empty0 + 8 * r.stage);  // frees the smem slot when these MMAs retire
                    r.advance(p.stages);
                }
                umma_commit(tfull0 + 8 * acc);  // accumulator complete -> epilogue
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue
        const int quarter = warp & 3;         // TMEM lanes [32*quarter, +32) are the only ones this warp may read
        const int half = (warp - 2) >> 2;     // which interleaved set of 16-column chunks
        const int row = quarter * 32 + lane;  // accumulator row = output pixel within the tile
        const int n_chunks = p.bn >> 4;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m0 = (tile / p.n_tiles) * BM, n0 = (tile % p.n_tiles) * p.bn;
            const int m = m0 + row;
            const bool m_ok = m < p.M;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * p.bn;
            for (int j = half; j < n_chunks; j += 2) {
                const int n = n0 + j * 16;
                if (n >= p.Cout) break;
                uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
                if (res != nullptr && m_ok) {
                    const uint4* rp = reinterpret_cast<const uint4*>(res + (size_t)m * p.res_pitch + n);
                    r0 = __ldg(rp);
                    r1 = __ldg(rp + 1);
                }
                uint32_t v[16];
                tmem_ld16(t_row + j * 16, v);
                const float4* bp = reinterpret_cast<const float4*>(bias + n);
                float4 b4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) b4[q] = __ldg(bp + q);
                tmem_ld_wait();
                float f[16];
                const float* bf = reinterpret_cast<const float*>(b4);
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = act_fast(__uint_as_float(v[i]) + bf[i], p.act);
                if (res != nullptr) {
                    const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        f[2 * i] += __uint_as_float(rr[i] << 16);
                        f[2 * i + 1] += __uint_as_float(rr[i] & 0xffff0000u);
                    }
                }
                if (m_ok) {
                    if (p.out_f32) {
                        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + (size_t)m * p.out_pitch + n);
#pragma unroll
                        for (int q = 0; q < 4; ++q) op[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                    } else {
                        uint32_t o[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                            o[i] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        uint4* op =
                            reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + (size_t)m * p.out_pitch + n);
                        op[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        op[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);  // this warp has drained its part of the accumulator
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

struct DriverApi {
    EncodeTiledFn tiled = nullptr;
    EncodeIm2colFn im2col = nullptr;
    int driver_version = 0;
    bool ok = false;
    DriverApi() {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return;
        tiled = (EncodeTiledFn)f;
        f = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return;
        im2col = (EncodeIm2colFn)f;
        cudaDriverGetVersion(&driver_version);
        ok = true;
    }
};
const DriverApi& driver() {
    static DriverApi api;
    return api;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int pick_kc(int cin) { return cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : 16); }

}  // namespace

bool conv2d_tc_supported(const fce_conv_desc* d, const void* x, const void* w, const void* res, const void* y) {
    if (d->in_dtype != FCE_BF16 || d->w_dtype != FCE_BF16 || d->in_layout != FCE_NHWC) return false;
    if (d->out_dtype != FCE_BF16 && d->out_dtype != FCE_F32) return false;
    if (d->in_scale != 1.0f) return false;
    if (d->Cin % 16 || d->Cout % 16) return false;
    if (d->in_pitch % 8 || d->in_off % 8 || !aligned16(x) || !aligned16(w) || !aligned16(y)) return false;
    if (d->out_dtype == FCE_BF16 ? (d->out_pitch % 8 || d->out_off % 8) : (d->out_pitch % 4 || d->out_off % 4)) return false;
    if (res && (d->out_dtype != FCE_BF16 || d->res_pitch % 8 || d->res_off % 8 || !aligned16(res))) return false;
    if (d->k == 1 && d->stride != 1) return false;
    if (d->H > 32000 || d->W > 32000) return false;
    return driver().ok;
}

int conv2d_tc(const fce_conv_desc* d, const void* x, const void* w, const float* bias, const void* res, void* y,
              cudaStream_t st) {
    if (!bias) return FCE_ERR_BAD_ARG;
    const DriverApi& api = driver();
    if (!api.ok) return FCE_ERR_CUDA;
    const int pad = d->k / 2;
    TcParams p{};
    p.Ho = (d->H + 2 * pad - d->k) / d->stride + 1;
    p.Wo = (d->W + 2 * pad - d->k) / d->stride + 1;
    const long long M = (long long)d->B * p.Ho * p.Wo;
    if (M > 0x7fffff00LL) return FCE_ERR_UNSUPPORTED;
    p.M = (int)M;
    p.Cout = d->Cout;
    p.Cin = d->Cin;
    p.ksz = d->k;
    p.taps = d->k * d->k;
    p.stride = d->stride;
    p.pad = pad;
    p.kc = pick_kc(d->Cin);
    p.chunks = d->Cin / p.kc;
    p.m_tiles = ceil_div(M, BM);
    p.n_tiles = ceil_div(d->Cout, 256);
    p.bn = ceil_div(ceil_div(d->Cout, p.n_tiles), 16) * 16;
    // small problems: narrower N tiles give the persistent grid more tiles to balance over 148 SMs
    while ((long long)p.m_tiles * p.n_tiles < 2 * kNumSMs && p.bn >= 128 && (p.bn / 2) % 16 == 0) {
        p.bn /= 2;
        p.n_tiles = ceil_div(d->Cout, p.bn);
    }
    p.a_bytes = BM * p.kc * 2;
    p.b_bytes = p.bn * p.kc * 2;
    p.a_stride = (p.a_bytes + 1023u) & ~1023u;
    p.b_stride = (p.b_bytes + 1023u) & ~1023u;
    p.stages = SMEM_BUDGET / (int)(p.a_stride + p.b_stride);
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    if (p.stages < 2) return FCE_ERR_UNSUPPORTED;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2u * p.bn) p.tmem_cols <<= 1;
    p.out_pitch = d->out_pitch;
    p.res_pitch = d->res_pitch;
    p.act = d->act;
    p.out_f32 = d->out_dtype == FCE_F32;
    const uint32_t row_bytes = p.kc * 2;                                        // swizzle span = K-chunk row
    const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);  // UMMA LayoutType
    const uint32_t sbo = 8 * row_bytes;                                         // 8-row core-matrix group pitch
    p.desc_hi = (sbo >> 4) | (1u << 14) | (layout << 29);
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                     : CU_TENSOR_MAP_SWIZZLE_32B;
    alignas(64) CUtensorMap tmA, tmB;
    const __nv_bfloat16* xin = reinterpret_cast<const __nv_bfloat16*>(x) + d->in_off;
    CUresult cr;
    if (d->k == 1) {
        const cuuint64_t gdim[2] = {(cuuint64_t)d->Cin, (cuuint64_t)M};
        const cuuint64_t gstr[1] = {(cuuint64_t)d->in_pitch * 2};
        const cuuint32_t box[2] = {(cuuint32_t)p.kc, (cuuint32_t)BM};
        const cuuint32_t est[2] = {1, 1};
        cr = api.tiled(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)xin, gdim, gstr, box, est,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t gdim[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
        const cuuint64_t gstr[3] = {(cuuint64_t)d->in_pitch * 2, (cuuint64_t)d->W * d->in_pitch * 2,
                                    (cuuint64_t)d->H * d->W * d->in_pitch * 2};
        // base pixels span [-pad, dim - 1 + pad - (k - 1)] so that base + tap offset covers the padded map
        const int lower[2] = {-pad, -pad};
        const int upper[2] = {pad - (d->k - 1), pad - (d->k - 1)};
        const cuuint32_t est[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
        cr = api.im2col(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)xin, gdim, gstr, lower, upper,
                        (cuuint32_t)p.kc, (cuuint32_t)BM, est, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        // Same descriptor fix-up CUTLASS applies for drivers <= 13.1 on tensors smaller than 128 KiB.
        const unsigned long long bytes = (unsigned long long)d->B * d->H * d->W * d->in_pitch * 2;
        if (cr == CUDA_SUCCESS && api.driver_version <= 13010 && bytes < 131072ull)
            reinterpret_cast<uint64_t*>(&tmA)[1] &= ~(1ull << 21);
    }
    if (cr != CUDA_SUCCESS) return FCE_ERR_UNSUPPORTED;
    {
        const cuuint64_t K = (cuuint64_t)p.taps * d->Cin;
        const cuuint64_t gdim[2] = {K, (cuuint64_t)d->Cout};
        const cuuint64_t gstr[1] = {K * 2};
        const cuuint32_t box[2] = {(cuuint32_t)p.kc, (cuuint32_t)p.bn};
        const cuuint32_t est[2] = {1, 1};
        cr = api.tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, est,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return FCE_ERR_UNSUPPORTED;
    }

    const size_t smem = (size_t)p.stages * (p.a_stride + p.b_stride) + 1024 + 256;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            set_cuda_error(e);
            return FCE_ERR_CUDA;
        }
        attr_set = true;
    }
    const int total = p.m_tiles * p.n_tiles;
    const int grid = total < kNumSMs ? total : kNumSMs;
    const __nv_bfloat16* rp = res ? reinterpret_cast<const __nv_bfloat16*>(res) + d->res_off : nullptr;
    void* yp = d->out_dtype == FCE_F32 ? (void*)(reinterpret_cast<float*>(y) + d->out_off)
                                       : (void*)(reinterpret_cast<__nv_bfloat16*>(y) + d->out_off);
    conv_tc_kernel<<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, p, bias, rp, yp);
    return check_launch();
}

}  // namespace fce

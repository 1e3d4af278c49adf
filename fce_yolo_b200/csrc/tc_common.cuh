// Shared device wrappers (mbarrier, TMA, tcgen05/TMEM, epilogue math) and the run-time resolved driver entry
// points of the tensor-core convolution kernels (conv_tc.cu: TMA-im2col implicit GEMM; conv_halo.cu: 3x3 stride-1
// with the input strip resident in shared memory).  sm_100a only.
#pragma once
#include <cstdlib>
#include <cuda.h>  // CUtensorMap types only - the two driver entry points are resolved at run time via cudart

#include "common.cuh"

namespace fce {
namespace tc {

// Programmatic dependent launch (PDL).  A kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still draining: everything it does before pdl_wait() - barrier init,
// TMEM allocation, tensor-map prefetch, loads of CONSTANT data (weights, bias) - overlaps the predecessor's tail.
// pdl_wait() returns once the predecessor grid has completed and its writes are visible; every access to
// activations (reads AND writes - arena buffers are recycled) must come after it.  pdl_launch_dependents() lets
// the successor's CTAs be scheduled as soon as all CTAs of this grid have issued it (or exited).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }


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
// The clock is read once per 256 polls: with clock64() and a 64-bit compare in every iteration the idle roles' spin loops were
// 25 - 40 % of all issued instructions of the fused producer kernels (ncu source view of conv_dwpw / conv_stem2), taking
// issue slots from the warps that do the work.  (A try_wait with a suspend-time hint compiled to NANOSLEEP.SYNCS + the same
// clock arithmetic and polled just as often.)
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 256; ++i) {
            if (mbar_try_wait(bar, parity)) return;  // (a __nanosleep(40) back-off here measured 6 % SLOWER on conv_stem2)
        }
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
// ---- CTA pairs (cta_group::2): two CTAs of a {2,1,1} cluster on the two SMs of one TPC run ONE tcgen05.mma of M = 256.
// Each CTA stages its own 128 rows of A and HALF of the B tile; the leader (cluster rank 0) issues the MMAs and its
// commits arrive on the barriers of both CTAs.  TMA loads of either CTA signal the LEADER's full barrier.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// data lands in THIS CTA's shared memory; the completion bytes go to `bar`, a shared::cluster address that may be the
// peer CTA's barrier (what .cta_group::2 permits)
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                                int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col_cg2(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w, int h,
                                                    int n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
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

// The pair forms: D[256 x N] over the TMEM of both CTAs (128 lanes each), A = 128 rows per CTA, B = N/2 rows per CTA.
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far retire.
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// One MUFU per element: silu(v) = h + h*tanh(h), sigmoid(v) = 0.5 + 0.5*tanh(h), h = v/2.  tanh.approx has
// 2^-11 relative error - below the bf16 rounding applied to the result.
__device__ __forceinline__ float act_fast(float v, int act) {
    if (act == FCE_ACT_SILU) {
        const float h = 0.5f * v;
        return fmaf(h, tanh_fast(h), h);
    }
    if (act == FCE_ACT_SIGMOID) return fmaf(0.5f, tanh_fast(0.5f * v), 0.5f);
    return v;
}

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

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
// 32 bytes from L2 in one request (sm_100: LDG.256); p must be 32-byte aligned
__device__ __forceinline__ void ldcg256(const uint4* p, uint4& a, uint4& b) {
    asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tm), "r"(src),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src),
                 "r"(c0), "r"(c1)
                 : "memory");
}

// bias + activation (+ residual) of 16 consecutive output channels of one pixel, on PACKED fp32 pairs (FFMA2 /
// FADD2: sm_100's two-wide fp32 pipe halves the issue slots of the epilogue, which - not HBM and not the tensor
// pipe - sets the pace of the short-K layers).  `sbias` holds the bias PRE-SCALED by epi_bias_scale(act):
// SiLU as h = 0.5*acc + 0.5*bias (one FFMA2 per pair), t = tanh.approx(h) (the one MUFU op per element),
// out = h + h*t (one FFMA2 per pair).
// (Measured alternatives: ex2+rcp = two MUFU ops; f16x2 tanh halves the MUFU work but costs more issue slots in
// conversions and was 5% slower - the epilogue is issue-bound, not MUFU-bound.)
__host__ __device__ __forceinline__ float epi_bias_scale(int act) { return act == FCE_ACT_SILU ? 0.5f : 1.f; }
// osc / rsc: weighted-sum epilogue y = osc * act(..) + rsc * res (BiFPN fusion); both 1 for the plain conv.
// pre: the residual joins the accumulator BEFORE bias and activation, y = act(acc + res + bias) - the second half of a
// 1x1 conv over a weighted sum of two maps (BiFPN node folded into its consumer).
__device__ __forceinline__ void epi_math16(const uint32_t* vin, const float* sbias, int act, bool has_res, uint4 r0,
                                           uint4 r1, float* f, float osc = 1.f, float rsc = 1.f, bool pre = false) {
    float2 o[8];
    uint32_t v[16];
    if (pre && has_res) {
        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 a = __fadd2_rn(make_float2(__uint_as_float(vin[2 * i]), __uint_as_float(vin[2 * i + 1])),
                                        make_float2(__uint_as_float(rr[i] << 16), __uint_as_float(rr[i] & 0xffff0000u)));
            v[2 * i] = __float_as_uint(a.x);
            v[2 * i + 1] = __float_as_uint(a.y);
        }
        has_res = false;
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = vin[i];
    }
    if (act == FCE_ACT_SILU) {
        const float2 half2 = make_float2(0.5f, 0.5f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 b = reinterpret_cast<const float4*>(sbias)[q];
            const float2 h0 = __ffma2_rn(make_float2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])), half2,
                                         make_float2(b.x, b.y));
            const float2 h1 = __ffma2_rn(make_float2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])), half2,
                                         make_float2(b.z, b.w));
            o[2 * q] = __ffma2_rn(h0, make_float2(tanh_fast(h0.x), tanh_fast(h0.y)), h0);
            o[2 * q + 1] = __ffma2_rn(h1, make_float2(tanh_fast(h1.x), tanh_fast(h1.y)), h1);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 b = reinterpret_cast<const float4*>(sbias)[q];
            o[2 * q] = __fadd2_rn(make_float2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])),
                                  make_float2(b.x, b.y));
            o[2 * q + 1] = __fadd2_rn(make_float2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])),
                                      make_float2(b.z, b.w));
        }
        if (act != FCE_ACT_NONE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = make_float2(act_fast(o[i].x, act), act_fast(o[i].y, act));
        }
    }
    const bool scaled = osc != 1.f || rsc != 1.f;  // warp-uniform
    if (has_res) {
        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        if (!scaled) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                o[i] = __fadd2_rn(o[i], make_float2(__uint_as_float(rr[i] << 16), __uint_as_float(rr[i] & 0xffff0000u)));
        } else {
            const float2 o2 = make_float2(osc, osc), r2 = make_float2(rsc, rsc);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                o[i] = __ffma2_rn(o[i], o2, __fmul2_rn(make_float2(__uint_as_float(rr[i] << 16),
                                                                   __uint_as_float(rr[i] & 0xffff0000u)), r2));
        }
    } else if (scaled) {
        const float2 o2 = make_float2(osc, osc);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __fmul2_rn(o[i], o2);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        f[2 * i] = o[i].x;
        f[2 * i + 1] = o[i].y;
    }
}
__device__ __forceinline__ void pack16(const float* f, uint32_t* o) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        o[i] = *reinterpret_cast<uint32_t*>(&h2);
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
inline const DriverApi& driver() {
    static DriverApi api;
    return api;
}


inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }


// Host: launch with the programmatic-stream-serialization attribute (see pdl_wait above).  Captured into CUDA
// graphs as a programmatic dependency edge.  FCE_NO_PDL=1 in the environment falls back to plain launches (A/B).
// `cluster` > 1: thread-block clusters of that many CTAs along x (CTA pairs: 2).
template <typename... KArgs, typename... Args>
inline int launch_pdl_cluster(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, int cluster,
                              Args... args) {
    static const bool no_pdl = [] {
        const char* e = getenv("FCE_NO_PDL");
        return e && e[0] == '1';
    }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cluster > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = (unsigned)cluster;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (!no_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    if (e != cudaSuccess) {
        set_cuda_error(e);
        return FCE_ERR_CUDA;
    }
    return check_launch();
}

template <typename... KArgs, typename... Args>
inline int launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    static const bool no_pdl = [] {
        const char* e = getenv("FCE_NO_PDL");
        return e && e[0] == '1';
    }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    if (e != cudaSuccess) {
        set_cuda_error(e);
        return FCE_ERR_CUDA;
    }
    return check_launch();
}

}  // namespace tc
}  // namespace fce

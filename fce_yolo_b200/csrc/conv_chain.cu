// Two chained 1x1 convolutions in ONE pass: the tail of a C3k2 block whose last inner block is a C3k,
//     t = act1(W1 * x1 + b1)                   C3k.cv3   (ultralytics/nn/modules/block.py:338-340, Conv.forward_fuse conv.py:80-89)
//     y = act2(W2 * [x2 ; t] + b2)             C3k2.cv2 over the concat [y0, y1, ..., t]   (block.py:303-307)
// As two launches t (c channels per pixel) is written to HBM by cv3 and read back by cv2 - both layers sit left of the
// ridge, so that round trip is pure cost (m scale, batch 256, 160x160: 0.84 GB out + 0.84 GB in, and cv3 itself is a 0.42 ms
// launch at 4 TB/s).  Here t never leaves the SM: GEMM 1 accumulates in its own TMEM columns, four epilogue warps apply
// bias + SiLU, round to bf16 (exactly what cv3 stores) and write the tile K-major / 128B-swizzled into shared memory -
// the layout a TMA load of the stored map would have produced - where it is the LAST K chunks of GEMM 2's A operand.
// Both GEMMs use the instruction shapes and the K order of the two-launch route: results are bit-identical to it.
//
//   work unit   one 128-pixel M tile; a CTA (one per SM, persistent) walks tiles b, b + grid, ... and, inside a tile, the
//               N tiles (128 columns) of GEMM 2 - t is computed once per M tile and reused by every N tile
//   A ring      TMA: x1 chunks [128 px, 64 ch] for GEMM 1, x2 chunks for GEMM 2 (re-read per N tile: L2 hits)
//   B ring      TMA: W1 chunks [c, 64] for GEMM 1, W2 chunks [128, 64] for GEMM 2 (all L2 hits)
//   MMA warp    per M tile i:  G2(i, n = 0) x2 part | wait t(i) | G1(i + 1) | G2(i, 0) t part | G2(i, n >= 1) ...
//               GEMM 1 of the NEXT tile is issued one tile ahead, so its epilogue (TMEM -> SiLU -> smem) runs under GEMM 2
//   warps 16-19 GEMM-1 epilogue: tcgen05.ld -> bias + act -> bf16 -> swizzled st.shared into t buffer i % TB
//   warps 0-11  GEMM-2 epilogue, three groups of four over three TMEM stages (conv_tc.cu's): -> staging slab -> TMA store
//
// Algorithmic HBM bytes per pixel: 2 * (c1 + c2) read + 2 * Cout written (the two launches: + 4 * c).
#include "tc_common.cuh"

namespace fce {
using namespace tc;
namespace {

constexpr int BM = 128, BN = 128, KC = 64;
constexpr int NUM_EPI2_WARPS = 12, NUM_GROUPS = 3;
constexpr int WARP_PROD_A = 12, WARP_PROD_B = 13, WARP_MMA = 14, WARP_ALLOC = 15, WARP_EPI1 = 16;
constexpr int NUM_THREADS = 20 * 32;
constexpr int MAX_STAGES = 8;
constexpr uint32_t STAGE = BM * KC * 2;   // 16 KB: one [128 rows x 128 bytes] K chunk (A, W2 tile, t chunk)
constexpr int STG_BYTES = 32 * 64;        // epilogue staging slab: 32 rows x 64 bytes (64B swizzle)
constexpr int SMEM_LIMIT = 227 * 1024;

struct ChainParams {
    int M, m_tiles, n_tiles;
    int c1, cm, c2, Cout;   // x1 channels (K of GEMM 1), t channels (N of GEMM 1), x2 channels, output channels
    int k1, kt, ky;         // K chunks: c1 / 64, cm / 64, c2 / 64
    int stages, tbufs;
    int act1, act2;
    uint32_t w1_bytes;      // one W1 chunk: cm x 128 bytes
    uint32_t bias_bytes;
    uint32_t desc_hi, idesc1, idesc2;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_chain_kernel(const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
                  const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                  const __grid_constant__ CUtensorMap tmC, const ChainParams p, const float* __restrict__ b1,
                  const float* __restrict__ b2) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int S = p.stages, TB = p.tbufs;
    const uint32_t sA = base;
    const uint32_t sB = sA + S * STAGE;
    const uint32_t sT = sB + S * STAGE;                        // TB x kt chunks
    const uint32_t sC = sT + TB * p.kt * STAGE;                // NUM_EPI2_WARPS x 2 slabs
    const uint32_t sBias = sC + NUM_EPI2_WARPS * 2 * STG_BYTES;  // [cm] GEMM-1 bias, then [Cout] GEMM-2 bias (pre-scaled)
    const uint32_t bars = sBias + p.bias_bytes;
    const uint32_t a_full0 = bars, a_empty0 = a_full0 + 8 * MAX_STAGES;
    const uint32_t b_full0 = a_empty0 + 8 * MAX_STAGES, b_empty0 = b_full0 + 8 * MAX_STAGES;
    const uint32_t tfull0 = b_empty0 + 8 * MAX_STAGES, tempty0 = tfull0 + 8 * NUM_GROUPS;  // GEMM-2 accumulator stages
    const uint32_t g1_full = tempty0 + 8 * NUM_GROUPS, g1_empty = g1_full + 8;             // GEMM-1 accumulator
    const uint32_t t_full0 = g1_empty + 8, t_empty0 = t_full0 + 16;                        // t buffers (TB <= 2)
    const uint32_t tmem_slot = t_empty0 + 16;
    float* bias_s = reinterpret_cast<float*>(smem_raw + (sBias - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = p.n_tiles, k1 = p.k1, kt = p.kt, ky = p.ky;
    // this CTA's M tiles: blockIdx.x, + gridDim.x, ...
    const int my_tiles = ((int)blockIdx.x < p.m_tiles) ? (p.m_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    pdl_launch_dependents();
    if (warp == WARP_PROD_A && lane == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(a_full0 + 8 * i, 1);
            mbar_init(a_empty0 + 8 * i, 1);
            mbar_init(b_full0 + 8 * i, 1);
            mbar_init(b_empty0 + 8 * i, 1);
        }
        for (int a = 0; a < NUM_GROUPS; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, 4);
        }
        mbar_init(g1_full, 1);
        mbar_init(g1_empty, 4);
        for (int t = 0; t < 2; ++t) {
            mbar_init(t_full0 + 8 * t, 4);
            mbar_init(t_empty0 + 8 * t, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmX1);
        tma_prefetch_desc(&tmX2);
        tma_prefetch_desc(&tmW1);
        tma_prefetch_desc(&tmW2);
        tma_prefetch_desc(&tmC);
    }
    if (warp == WARP_ALLOC) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // biases pre-scaled for epi_math16 (x 1/2 under SiLU)
        const float s1 = epi_bias_scale(p.act1), s2 = epi_bias_scale(p.act2);
        for (int i = threadIdx.x; i < p.cm; i += NUM_THREADS) bias_s[i] = b1[i] * s1;
        for (int i = threadIdx.x; i < p.Cout; i += NUM_THREADS) bias_s[p.cm + i] = b2[i] * s2;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const uint32_t tmem_g1 = tmem_base + NUM_GROUPS * BN;  // GEMM-1 accumulator: columns [384, 384 + cm)

    if (warp == WARP_PROD_A) {
        // ------------------------------------------------------------------ A producer: x1 chunks (GEMM 1), x2 chunks (GEMM 2)
        Ring r;
        auto load = [&](const CUtensorMap* tm, int c0, int m0) {
            mbar_wait(a_empty0 + 8 * r.stage, r.phase ^ 1);
            if (elect_one()) {
                const uint32_t fb = a_full0 + 8 * r.stage;
                mbar_expect_tx(fb, STAGE);
                tma_load_2d(sA + r.stage * STAGE, tm, fb, c0, m0);
            }
            __syncwarp();
            r.advance(S);
        };
        pdl_wait();  // activations come from the previous kernel
        if (my_tiles > 0)
            for (int g = 0; g < k1; ++g) load(&tmX1, g * KC, (int)blockIdx.x * BM);
        for (int i = 0; i < my_tiles; ++i) {
            const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * BM;
            for (int n = 0; n < n_tiles; ++n) {
                for (int g = 0; g < ky; ++g) load(&tmX2, g * KC, m0);
                if (n == 0 && i + 1 < my_tiles)
                    for (int g = 0; g < k1; ++g) load(&tmX1, g * KC, m0 + (int)gridDim.x * BM);
            }
        }
    } else if (warp == WARP_PROD_B) {
        // ------------------------------------------------------------------ B producer: W1 chunks, W2 tiles (constants: no pdl_wait)
        Ring r;
        auto load = [&](const CUtensorMap* tm, uint32_t bytes, int c0, int n0) {
            mbar_wait(b_empty0 + 8 * r.stage, r.phase ^ 1);
            if (elect_one()) {
                const uint32_t fb = b_full0 + 8 * r.stage;
                mbar_expect_tx(fb, bytes);
                tma_load_2d(sB + r.stage * STAGE, tm, fb, c0, n0);
            }
            __syncwarp();
            r.advance(S);
        };
        if (my_tiles > 0)
            for (int g = 0; g < k1; ++g) load(&tmW1, p.w1_bytes, g * KC, 0);
        for (int i = 0; i < my_tiles; ++i)
            for (int n = 0; n < n_tiles; ++n) {
                for (int g = 0; g < ky; ++g) load(&tmW2, STAGE, g * KC, n * BN);
                if (n == 0 && i + 1 < my_tiles)
                    for (int g = 0; g < k1; ++g) load(&tmW1, p.w1_bytes, g * KC, 0);
                for (int g = 0; g < kt; ++g) load(&tmW2, STAGE, (ky + g) * KC, n * BN);
            }
    } else if (warp == WARP_MMA) {
        // ------------------------------------------------------------------ MMA issuer
        Ring ra, rb;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t dhi = p.desc_hi, idesc1 = p.idesc1, idesc2 = p.idesc2;
        auto lo = [](uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); };
        // GEMM 1 of this CTA's tile `idx`: accumulator drained by the epilogue of tile idx - 1
        auto gemm1 = [&](int idx) {
            mbar_wait(g1_empty, (uint32_t)((idx & 1) ^ 1));
            tc_fence_after();
#pragma unroll 1
            for (int g = 0; g < k1; ++g) {
                mbar_wait(a_full0 + 8 * ra.stage, ra.phase);
                mbar_wait(b_full0 + 8 * rb.stage, rb.phase);
                tc_fence_after();
                const uint32_t a_lo = lo(sA + ra.stage * STAGE), b_lo = lo(sB + rb.stage * STAGE);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < KC / 16; ++k)
                        umma_bf16(tmem_g1, make_desc(dhi, a_lo + 2 * k), make_desc(dhi, b_lo + 2 * k), idesc1, (g | k) != 0);
                    umma_commit(a_empty0 + 8 * ra.stage);
                    umma_commit(b_empty0 + 8 * rb.stage);
                    if (g == k1 - 1) umma_commit(g1_full);
                }
                __syncwarp();
                ra.advance(S);
                rb.advance(S);
            }
        };
        if (my_tiles > 0) gemm1(0);
        for (int i = 0; i < my_tiles; ++i) {
            const int tb = i % TB;
            const uint32_t t_phase = (uint32_t)((i / TB) & 1);
            for (int n = 0; n < n_tiles; ++n) {
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
#pragma unroll 1
                for (int g = 0; g < ky; ++g) {   // x2 part of K
                    mbar_wait(a_full0 + 8 * ra.stage, ra.phase);
                    mbar_wait(b_full0 + 8 * rb.stage, rb.phase);
                    tc_fence_after();
                    const uint32_t a_lo = lo(sA + ra.stage * STAGE), b_lo = lo(sB + rb.stage * STAGE);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k)
                            umma_bf16(d_tmem, make_desc(dhi, a_lo + 2 * k), make_desc(dhi, b_lo + 2 * k), idesc2, (g | k) != 0);
                        umma_commit(a_empty0 + 8 * ra.stage);
                        umma_commit(b_empty0 + 8 * rb.stage);
                    }
                    __syncwarp();
                    ra.advance(S);
                    rb.advance(S);
                }
                if (n == 0) {
                    mbar_wait(t_full0 + 8 * tb, t_phase);  // t of this tile is in shared memory (and GEMM 1's accumulator is free)
                    tc_fence_after();
                    if (i + 1 < my_tiles) gemm1(i + 1);
                }
#pragma unroll 1
                for (int g = 0; g < kt; ++g) {   // t part of K: A operand = the t buffer
                    mbar_wait(b_full0 + 8 * rb.stage, rb.phase);
                    tc_fence_after();
                    const uint32_t a_lo = lo(sT + (tb * kt + g) * STAGE), b_lo = lo(sB + rb.stage * STAGE);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k)
                            umma_bf16(d_tmem, make_desc(dhi, a_lo + 2 * k), make_desc(dhi, b_lo + 2 * k), idesc2, 1u);
                        umma_commit(b_empty0 + 8 * rb.stage);
                        if (g == kt - 1) {
                            if (n == n_tiles - 1) umma_commit(t_empty0 + 8 * tb);   // last reader of this t buffer
                            umma_commit(tfull0 + 8 * acc);                          // accumulator complete -> epilogue 2
                        }
                    }
                    __syncwarp();
                    rb.advance(S);
                }
                if (++acc == NUM_GROUPS) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= WARP_EPI1) {
        // ------------------------------------------------------------------ GEMM-1 epilogue: TMEM -> bias + act -> bf16 -> t buffer
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t t_row = tmem_g1 + ((uint32_t)(quarter * 32) << 16);
        const uint32_t row_off = (uint32_t)row * 128u, rx = (uint32_t)(row & 7);
        const uint4 z = make_uint4(0, 0, 0, 0);
        const int act1 = p.act1, cm = p.cm;
        for (int i = 0; i < my_tiles; ++i) {
            const int tb = i % TB;
            mbar_wait(g1_full, (uint32_t)(i & 1));
            tc_fence_after();
            mbar_wait(t_empty0 + 8 * tb, (uint32_t)(((i / TB) & 1) ^ 1));  // GEMM 2 of the tile that used this buffer has retired
            const uint32_t tbuf = sT + tb * kt * STAGE;
#pragma unroll 1
            for (int c0 = 0; c0 < cm; c0 += 32) {
                uint32_t v0[16], v1[16];
                tmem_ld16(t_row + c0, v0);
                tmem_ld16(t_row + c0 + 16, v1);
                tmem_ld_wait();
                const uint32_t chunk = tbuf + (uint32_t)(c0 >> 6) * STAGE + row_off;
                const uint32_t u0 = (uint32_t)((c0 & 63) >> 3);  // first 16-byte unit (8 channels) of this column group
                float f[16];
                uint32_t o[8];
                epi_math16(v0, bias_s + c0, act1, false, z, z, f);
                pack16(f, o);
                st_shared_v4(chunk + (((u0 + 0) ^ rx) << 4), o[0], o[1], o[2], o[3]);
                st_shared_v4(chunk + (((u0 + 1) ^ rx) << 4), o[4], o[5], o[6], o[7]);
                epi_math16(v1, bias_s + c0 + 16, act1, false, z, z, f);
                pack16(f, o);
                st_shared_v4(chunk + (((u0 + 2) ^ rx) << 4), o[0], o[1], o[2], o[3]);
                st_shared_v4(chunk + (((u0 + 3) ^ rx) << 4), o[4], o[5], o[6], o[7]);
            }
            tc_fence_before();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(g1_empty);
                mbar_arrive(t_full0 + 8 * tb);
            }
        }
    } else if (warp < NUM_EPI2_WARPS) {
        // ------------------------------------------------------------------ GEMM-2 epilogue (conv_tc.cu's, bf16 out, no residual)
        const int quarter = warp & 3, group = warp >> 2;
        const int act2 = p.act2, Cout = p.Cout;
        const float* bias2 = bias_s + p.cm;
        const uint32_t stg0 = sC + warp * 2 * STG_BYTES;
        const uint32_t swz = (uint32_t)((lane >> 1) & 3);  // 64B swizzle: 16-byte unit u of row r lives at u ^ ((r >> 1) & 3)
        const uint32_t my_row = stg0 + lane * 64;
        const uint4 z = make_uint4(0, 0, 0, 0);
        int buf = 0, acc = 0;
        uint32_t acc_phase = 0;
        pdl_wait();  // the output buffer may still be in use by the previous kernel (arena buffers are recycled)
        for (int i = 0; i < my_tiles; ++i) {
            const int m_warp = ((int)blockIdx.x + i * (int)gridDim.x) * BM + quarter * 32;
            for (int nt = 0; nt < n_tiles; ++nt) {
                const bool mine = acc == group;
                const int my_acc = acc;
                const uint32_t my_phase = acc_phase;
                if (++acc == NUM_GROUPS) {
                    acc = 0;
                    acc_phase ^= 1;
                }
                if (!mine) continue;
                const int n0 = nt * BN;
                mbar_wait(tfull0 + 8 * my_acc, my_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + my_acc * BN;
#pragma unroll 1
                for (int sl = 0; sl < BN / 32; ++sl) {
                    const int c0 = sl * 32, n = n0 + c0;
                    if (n >= Cout) break;
                    uint32_t v0[16], v1[16];
                    tmem_ld16(t_row + c0, v0);
                    tmem_ld16(t_row + c0 + 16, v1);
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // slab of two stores ago drained
                    __syncwarp();
                    tmem_ld_wait();
                    const uint32_t rowp = my_row + buf * STG_BYTES;
                    float f[16];
                    uint32_t o[8];
                    epi_math16(v0, bias2 + n, act2, false, z, z, f);
                    pack16(f, o);
                    st_shared_v4(rowp + ((0 ^ swz) << 4), o[0], o[1], o[2], o[3]);
                    st_shared_v4(rowp + ((1 ^ swz) << 4), o[4], o[5], o[6], o[7]);
                    epi_math16(v1, bias2 + n + 16, act2, false, z, z, f);
                    pack16(f, o);
                    st_shared_v4(rowp + ((2 ^ swz) << 4), o[0], o[1], o[2], o[3]);
                    st_shared_v4(rowp + ((3 ^ swz) << 4), o[4], o[5], o[6], o[7]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && m_warp < p.M) {
                        tma_store_2d(&tmC, stg0 + buf * STG_BYTES, n, m_warp);  // rows >= M are clipped
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    buf ^= 1;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8 * my_acc);
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == WARP_ALLOC)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

size_t chain_smem(const ChainParams& p) {
    return (size_t)(2 * p.stages + p.tbufs * p.kt) * STAGE + NUM_EPI2_WARPS * 2 * STG_BYTES + p.bias_bytes + 512 + 1024;
}

// Shape rules and the shared-memory split; pure arithmetic (no CUDA calls): also behind fce_conv1x1_chain_route.
bool chain_plan(const fce_chain_desc* d, ChainParams& p) {
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return false;
    if (d->c1 <= 0 || d->c1 % KC || d->c2 <= 0 || d->c2 % KC) return false;
    if (d->cm != 64 && d->cm != 128) return false;          // GEMM-1 accumulator: the 128 TMEM columns next to 3 x 128
    if (d->Cout <= 0 || d->Cout % BN) return false;
    if (d->x1_pitch % 8 || d->x1_off % 8 || d->x2_pitch % 8 || d->x2_off % 8 || d->out_pitch % 8 || d->out_off % 8) return false;
    if (d->x1_pitch < d->c1 || d->x2_pitch < d->c2 || d->out_pitch < d->Cout) return false;
    for (int a : {d->act1, d->act2})
        if (a != FCE_ACT_SILU && a != FCE_ACT_NONE && a != FCE_ACT_SIGMOID) return false;
    const long long M = (long long)d->B * d->H * d->W;
    if (M > 0x7fffff00LL) return false;
    p.M = (int)M;
    p.m_tiles = ceil_div(M, BM);
    p.n_tiles = d->Cout / BN;
    p.c1 = d->c1; p.cm = d->cm; p.c2 = d->c2; p.Cout = d->Cout;
    p.k1 = d->c1 / KC; p.kt = d->cm / KC; p.ky = d->c2 / KC;
    p.act1 = d->act1; p.act2 = d->act2;
    p.w1_bytes = (uint32_t)d->cm * 128u;
    p.bias_bytes = ((uint32_t)(d->cm + d->Cout) * 4u + 1023u) & ~1023u;
    // two t buffers (the GEMM-1 epilogue of tile i + 1 never waits for GEMM 2 of tile i) where four ring stages still fit
    auto stages_with = [&](int tb) {
        const long long fixed = (long long)tb * p.kt * STAGE + NUM_EPI2_WARPS * 2 * STG_BYTES + p.bias_bytes + 512 + 1024;
        const int st = (int)((SMEM_LIMIT - fixed) / (2 * (long long)STAGE));
        return st > MAX_STAGES ? MAX_STAGES : st;
    };
    static const int tb_env = [] { const char* e = getenv("FCE_CHAIN_TB"); return e && *e ? atoi(e) : 0; }();  // A/B timing
    p.tbufs = (tb_env == 1 || tb_env == 2) ? tb_env : (stages_with(2) >= 4 ? 2 : 1);
    p.stages = stages_with(p.tbufs);
    if (p.stages < 3) return false;
    p.desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 8 rows x 128 bytes, descriptor version 1, 128B swizzle
    p.idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(d->cm >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    p.idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    return true;
}

}  // namespace
}  // namespace fce

using namespace fce;

extern "C" int fce_conv1x1_chain_route(const fce_chain_desc* d) {
    if (!d) return FCE_ERR_BAD_ARG;
    ChainParams p{};
    return chain_plan(d, p) ? 1 : 0;
}

extern "C" int fce_conv1x1_chain(const fce_chain_desc* d, const void* x1, const void* w1, const float* b1, const void* x2,
                                 const void* w2, const float* b2, void* y, void* stream) {
    if (!d || !x1 || !w1 || !b1 || !x2 || !w2 || !b2 || !y) return FCE_ERR_BAD_ARG;
    if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->c1 <= 0 || d->cm <= 0 || d->c2 <= 0 || d->Cout <= 0) return FCE_ERR_BAD_ARG;
    ChainParams p{};
    if (!chain_plan(d, p)) return FCE_ERR_UNSUPPORTED;
    if (!aligned16(x1) || !aligned16(x2) || !aligned16(w1) || !aligned16(w2) || !aligned16(y)) return FCE_ERR_ALIGNMENT;
    const DriverApi& api = driver();
    if (!api.ok) return FCE_ERR_CUDA;
    const cuuint32_t est[2] = {1, 1};
    auto act_map = [&](CUtensorMap* tm, const void* ptr, int off, int ch, int pitch) {
        const cuuint64_t gdim[2] = {(cuuint64_t)ch, (cuuint64_t)p.M};
        const cuuint64_t gstr[1] = {(cuuint64_t)pitch * 2};
        const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)BM};
        return api.tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)(reinterpret_cast<const __nv_bfloat16*>(ptr) + off), gdim,
                         gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    auto w_map = [&](CUtensorMap* tm, const void* ptr, int K, int N, int rows) {
        const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
        const cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
        const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)rows};
        return api.tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, est,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    alignas(64) CUtensorMap tmX1, tmX2, tmW1, tmW2, tmC;
    if (act_map(&tmX1, x1, d->x1_off, d->c1, d->x1_pitch) != CUDA_SUCCESS) return FCE_ERR_UNSUPPORTED;
    if (act_map(&tmX2, x2, d->x2_off, d->c2, d->x2_pitch) != CUDA_SUCCESS) return FCE_ERR_UNSUPPORTED;
    if (w_map(&tmW1, w1, d->c1, d->cm, d->cm) != CUDA_SUCCESS) return FCE_ERR_UNSUPPORTED;
    if (w_map(&tmW2, w2, d->c2 + d->cm, d->Cout, BN) != CUDA_SUCCESS) return FCE_ERR_UNSUPPORTED;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)d->Cout, (cuuint64_t)p.M};
        const cuuint64_t gstr[1] = {(cuuint64_t)d->out_pitch * 2};
        const cuuint32_t box[2] = {32, 32};
        if (api.tiled(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)(reinterpret_cast<__nv_bfloat16*>(y) + d->out_off), gdim,
                      gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCE_ERR_UNSUPPORTED;
    }
    static DeviceOnce attr_once;  // the shared-memory opt-in is a per-device attribute
    int dev = 0;
    if (attr_once.pending(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) {
            set_cuda_error(e);
            return FCE_ERR_CUDA;
        }
        attr_once.done(dev);
    }
    const int grid = p.m_tiles < kNumSMs ? p.m_tiles : kNumSMs;
    return launch_pdl(conv_chain_kernel, grid, NUM_THREADS, chain_smem(p), (cudaStream_t)stream, tmX1, tmX2, tmW1, tmW2, tmC, p,
                      b1, b2);
}

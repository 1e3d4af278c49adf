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
//   * One CTA per SM, persistent over (m, n) tiles, warp-specialised (16 warps):
//       warps 0-11 : epilogue, three groups of four (one warp per TMEM lane quarter), tile i -> group i % 3:
//                    tcgen05.ld -> +bias -> SiLU -> (+residual) -> bf16/fp32 -> swizzled staging slab -> TMA store
//                    at the channel offset of the destination view (concat fusion)
//       warp 12/13 : TMA producers for A and B (one elected lane), `stages`-deep full/empty mbarrier ring
//       warp 14    : single-thread tcgen05.mma issuer; the accumulator (128 x bn fp32) lives in TMEM with up to three
//                    stages so the epilogues of tiles i, i+1 overlap the main loop of tile i+2
//       warp 15    : TMEM allocator
//   * Shared-memory tiles use the hardware swizzle that matches the K chunk (128B / 64B / 32B for
//     kc = 64 / 32 / 16 channels), identical in the TMA descriptor and the UMMA shared-memory descriptor.
//   * CTA PAIRS (template PAIR, cta_group::2): two CTAs of a {2,1,1} cluster - the two SMs of a TPC - own a 256-pixel
//     M tile.  Each CTA stages ITS 128 pixels of A and HALF of the [bn, kc] weight tile (bn/2 rows); the leader CTA
//     (cluster rank 0) issues ONE tcgen05.mma.cta_group::2 of M = 256 per K step, which reads both halves of B across
//     the pair and accumulates 128 rows in each CTA's TMEM.  Per K step an SM reads 4 KB of A + bn*16 bytes of B from
//     shared memory instead of 4 KB + bn*32 (the SS-mode operand bandwidth that capped the N <= 128 layers), and the
//     L2 -> SM weight traffic per output pixel halves (what capped the long-K 3x3 layers: a 128 x 256 tile moves
//     (128 + 256) * kc * 2 bytes per K step, a pair (256 + 256) for twice the work).  Barriers: TMA loads of both
//     CTAs complete on the LEADER's full barrier (.cta_group::2 loads, the leader's producers post the expected
//     bytes of both CTAs); the leader's commits are multicast to the empty / accumulator-full barriers of both CTAs;
//     the epilogue warps of both CTAs arrive on the leader's accumulator-empty barrier (remote arrive).  The epilogue
//     itself is unchanged: every CTA drains its own 128 TMEM lanes and stores its own rows.
#include <atomic>

#include "tc_common.cuh"

namespace fce {
using namespace tc;
namespace {

constexpr int BM = 128;  // UMMA M (cta_group::1)
// Warp roles: 12 epilogue warps (three groups of four, one warp per TMEM lane quarter; tiles rotate over the
// groups), then the single-lane roles.
constexpr int WARP_EPI0 = 0;
constexpr int NUM_EPI_WARPS = 12;  // three groups x four lane quarters
constexpr int NUM_GROUPS = NUM_EPI_WARPS / 4;
constexpr int MAX_ACC = 3;         // TMEM accumulator stages (3 x bn <= 512 columns, else 2)
constexpr int WARP_PROD_A = 12, WARP_PROD_B = 13, WARP_MMA = 14, WARP_ALLOC = 15;
constexpr int NUM_THREADS = 16 * 32;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 172 * 1024;  // A/B tiles + bias; staging slabs and barriers come on top
constexpr int B_RESIDENT_MAX = 96 * 1024;
constexpr int STG_BYTES = 32 * 64;  // one epilogue staging slab: 32 rows x 64 bytes (64B swizzle)

struct TcParams {
    int M, Cout, Cin;
    int taps, ksz, stride, pad;
    int Ho, Wo;
    int kc, chunks;   // channels per K step, K steps per tap
    int S, groups;    // K steps per pipeline stage, stages per tile (= taps*chunks / S)
    int bn, n_tiles, m_tiles;
    int stages;
    int b_resident;   // whole [Cout, K] weight matrix parked in shared memory for the life of the CTA
    uint32_t a_sub, b_sub;      // bytes of one K-step sub-tile: 128 x kc and bn x kc bf16
    uint32_t a_stage, b_stage;  // S sub-tiles
    uint32_t b_total;           // shared-memory bytes of the B region
    uint32_t bias_bytes;
    uint32_t tmem_cols;
    int acc_stages;    // TMEM accumulator stages: tile `it` of a CTA accumulates in stage it % acc_stages
    int m_units;       // M work units: m_tiles for single CTAs, ceil(m_tiles / 2) for CTA pairs
    int out_pitch, res_pitch, act, out_f32;
    uint32_t desc_hi;  // upper 32 bits of the UMMA shared-memory descriptor (SBO, version, swizzle)
    uint32_t idesc;    // UMMA instruction descriptor
    int dbg;           // debug: bit 0 = producers skip the TMA loads (times the MMA + epilogue pipeline alone)
    // Detect-head epilogues (fce_conv2d_detect): the accumulator never becomes a logit tensor in HBM
    int epi_mode;      // 0 = plain; 1 = class scores -> y[b, 4 + n, a] = sigmoid; 2 = DFL boxes -> y[b, 0..3, a]
    int epi_A, epi_abase, epi_rows;  // anchors per image, first anchor of this level, rows of y per image (4 + nc)
    float epi_stride;
    // weighted-sum epilogue (BiFPN fused into its realign conv): y = out_scale * act(..) + res_scale * res
    float out_scale, res_scale;
    int res_up;        // res is a half-resolution map read through a nearest 2x upsample
    int res_wide;      // residual rows are 32-byte aligned: 256-bit loads
    int res_pre;       // res joins the accumulator before bias + activation (fce_conv_desc.weighted == 2)
};

// Per-role cycle accounting and the load / store / math switch-off bits exist in -DFCE_DEBUG builds only (debug entry
// points fce_conv_tc_set_profile / fce_conv_tc_profile, declared in api.cu, NOT in the public header):
// [cta][0..1] A producer wait/total, [4..6] MMA wait-full / wait-tmem-empty / total,
// [7..8] epilogue warp 0 wait-tmem-full / total.
constexpr int PROF_SLOTS = 16;
#ifdef FCE_DEBUG
constexpr bool PROF = true, DBG = true;
__device__ long long g_prof[kNumSMs * PROF_SLOTS];
#else
constexpr bool PROF = false, DBG = false;
__device__ long long g_prof[1];
#endif

#define PROF_T0() long long _t0 = 0; if (PROF) _t0 = clock64()
#define PROF_ACC(var) if (PROF) (var) += clock64() - _t0

// ------------------------------------------------------------------------------------------------ kernel
// The single-thread roles (TMA producers, MMA issuer) run their loops WARP-UNIFORMLY - every lane executes the
// loop control and the barrier waits, and only the TMA / tcgen05 instruction itself is predicated on one
// elected lane.  This keeps addresses, phases and descriptors in the uniform datapath (the async-proxy
// instructions take uniform registers) instead of a per-thread dependent chain with R2UR moves.
// KK = K chunk / 16 (UMMA instructions per K step), S = K steps per pipeline stage: compile-time so that the
// issue loops are straight-line code (descriptor = base + constant).
template <int KK, int S, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const TcParams p, const float* __restrict__ bias,
               const __nv_bfloat16* __restrict__ res, float* __restrict__ ydet) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base;
    const uint32_t sB = sA + p.stages * p.a_stage;
    const uint32_t sC = sB + p.b_total;  // epilogue staging: NUM_EPI_WARPS x 2 x (32 rows x 64 bytes)
    const uint32_t sBias = sC + NUM_EPI_WARPS * 2 * STG_BYTES;
    const uint32_t bars = sBias + p.bias_bytes;
    const uint32_t full0 = bars, empty0 = bars + 8 * MAX_STAGES;
    const uint32_t tfull0 = bars + 16 * MAX_STAGES, tempty0 = tfull0 + 32;
    const uint32_t bfull = tempty0 + 32;
    const uint32_t tmem_slot = bfull + 8;
    float* bias_s = reinterpret_cast<float*>(smem_raw + (sBias - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work units: (M unit, N tile).  Single CTA: M unit = one 128-row tile, CTA b takes units b, b + grid, ...  Pair: M unit
    // = two consecutive 128-row tiles (rank r of the pair owns tile 2 * unit + r), pair q = blockIdx.x / 2 takes units
    // q, q + grid / 2, ... - BOTH CTAs of a pair walk the same unit sequence (same ring / accumulator phases).
    constexpr int NCTA = PAIR ? 2 : 1;
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int total_tiles = p.m_units * p.n_tiles;
    // barriers the PEER must reach live in the leader's shared memory
    const uint32_t full0_sig = PAIR ? mapa_shared(full0, 0) : full0;      // TMA completion target
    const uint32_t bfull_sig = PAIR ? mapa_shared(bfull, 0) : bfull;
    const uint32_t tempty0_sig = PAIR ? mapa_shared(tempty0, 0) : tempty0;  // epilogue -> MMA issuer

    pdl_launch_dependents();  // the next kernel's prologue may overlap this kernel's tail
    if (warp == WARP_PROD_A && lane == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(full0 + 8 * i, p.b_resident ? 1 : 2);  // arrivals: this CTA's producers (pair: the leader's only)
            mbar_init(empty0 + 8 * i, 1);
        }
        for (int a = 0; a < MAX_ACC; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, 4 * NCTA);  // one group of four epilogue warps (per CTA) drains an accumulator
        }
        mbar_init(bfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
    }
    if (warp == WARP_ALLOC) {
        if (PAIR) {  // one warp of EACH CTA of the pair: the allocation is made in both tensor memories at once
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    {   // bias pre-scaled for the epilogue: x 1/2 for SiLU (epi_math16), x -log2(e) for the fused class sigmoid
        const float bsc = p.epi_mode == 1 ? -1.4426950408889634f : epi_bias_scale(p.act);
        for (int i = threadIdx.x; i < (int)(p.bias_bytes >> 2); i += NUM_THREADS) bias_s[i] = i < p.Cout ? bias[i] * bsc : 0.f;
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();  // barrier inits and the TMEM allocation of BOTH CTAs are visible to both
    else __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == WARP_PROD_A) {
        // ------------------------------------------------------------------ A producer (activations)
        Ring r;
        constexpr int kc = KK * 16;
        constexpr uint32_t a_sub = BM * kc * 2, a_stage = S * a_sub;
        const int hw = p.Ho * p.Wo, groups = p.groups, chunks = p.chunks, ksz = p.ksz;
        const bool is_1x1 = p.taps == 1, skip = DBG && (p.dbg & 1) != 0;
        const int n_tiles = p.n_tiles, nstages = p.stages;
        long long pw = 0, pt0 = PROF ? clock64() : 0;
        pdl_wait();  // activations are produced by the previous kernel (weights / bias above are constants)
        for (int tile = unit0; tile < total_tiles; tile += unit_step) {
            int mt = (n_tiles == 1 ? tile : tile / n_tiles) * NCTA + (int)cta_rank;
            if (PAIR && mt >= p.m_tiles) mt = p.m_tiles - 1;  // odd tile count: the pair's last half re-reads valid rows (never stored)
            const int m0 = mt * BM;
            int img = 0, w0 = 0, h0 = 0;
            if (!is_1x1) {
                img = m0 / hw;
                const int rem = m0 - img * hw;
                const int ph = rem / p.Wo;
                h0 = ph * p.stride - p.pad;
                w0 = (rem - ph * p.Wo) * p.stride - p.pad;
            }
            int ch = 0, kw = 0, kh = 0;
#pragma unroll 1
            for (int g = 0; g < groups; ++g) {
                const uint32_t fb = full0 + 8 * r.stage;
                {
                    PROF_T0();
                    mbar_wait(empty0 + 8 * r.stage, r.phase ^ 1);
                    PROF_ACC(pw);
                }
                const uint32_t dst = sA + r.stage * a_stage;
                const uint32_t fsig = full0_sig + 8 * r.stage;
                if (elect_one()) {
                    if (skip) {
                        if (leader) mbar_arrive(fb);
                    } else {
                        if (leader) mbar_expect_tx(fb, a_stage * NCTA);  // the bytes of both CTAs' A tiles
                        int c_ = ch, kw_ = kw, kh_ = kh;
#pragma unroll
                        for (int q = 0; q < S; ++q) {
                            if (PAIR) {
                                if (is_1x1)
                                    tma_load_2d_cg2(dst + q * a_sub, &tmA, fsig, c_ * kc, m0);
                                else
                                    tma_load_im2col_cg2(dst + q * a_sub, &tmA, fsig, c_ * kc, w0, h0, img, (uint16_t)kw_,
                                                        (uint16_t)kh_);
                            } else if (is_1x1)
                                tma_load_2d(dst + q * a_sub, &tmA, fb, c_ * kc, m0);
                            else
                                tma_load_im2col(dst + q * a_sub, &tmA, fb, c_ * kc, w0, h0, img, (uint16_t)kw_,
                                                (uint16_t)kh_);
                            if (++c_ == chunks) {
                                c_ = 0;
                                if (++kw_ == ksz) {
                                    kw_ = 0;
                                    ++kh_;
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < S; ++q) {
                    if (++ch == chunks) {
                        ch = 0;
                        if (++kw == ksz) {
                            kw = 0;
                            ++kh;
                        }
                    }
                }
                r.advance(nstages);
            }
        }
        if (PROF && lane == 0) {
            g_prof[blockIdx.x * PROF_SLOTS + 0] = pw;
            g_prof[blockIdx.x * PROF_SLOTS + 1] = clock64() - pt0;
        }
    } else if (warp == WARP_PROD_B) {
        // ------------------------------------------------------------------ B producer (weights)
        constexpr int kc = KK * 16;
        const int k_steps = p.groups * S;
        const uint32_t b_sub = p.b_sub;
        const bool skip = DBG && (p.dbg & 1) != 0;
        const int nrow0 = (int)cta_rank * (p.bn / NCTA);  // pair: this CTA stages rows [rank * bn/2, +bn/2) of the weight tile
        if (p.b_resident) {
            // weight-stationary: one load of the whole [Cout, K] matrix (pair: this CTA's half of the rows) for all tiles
            if (elect_one()) {
                if (leader) mbar_expect_tx(bfull, (uint32_t)k_steps * b_sub * NCTA);
                for (int it = 0; it < k_steps; ++it) {
                    if (PAIR) tma_load_2d_cg2(sB + it * b_sub, &tmB, bfull_sig, it * kc, nrow0);
                    else tma_load_2d(sB + it * b_sub, &tmB, bfull, it * kc, 0);
                }
            }
        } else {
            Ring r;
            const int groups = p.groups, n_tiles = p.n_tiles, nstages = p.stages, bn = p.bn;
            const uint32_t b_stage = p.b_stage;
            for (int tile = unit0; tile < total_tiles; tile += unit_step) {
                const int n0 = (n_tiles == 1 ? 0 : tile % n_tiles) * bn + nrow0;
                int kcol = 0;
#pragma unroll 1
                for (int g = 0; g < groups; ++g) {
                    const uint32_t fb = full0 + 8 * r.stage;
                    mbar_wait(empty0 + 8 * r.stage, r.phase ^ 1);
                    const uint32_t dst = sB + r.stage * b_stage;
                    const uint32_t fsig = full0_sig + 8 * r.stage;
                    if (elect_one()) {
                        if (skip) {
                            if (leader) mbar_arrive(fb);
                        } else {
                            if (leader) mbar_expect_tx(fb, b_stage * NCTA);
#pragma unroll
                            for (int q = 0; q < S; ++q) {
                                if (PAIR) tma_load_2d_cg2(dst + q * b_sub, &tmB, fsig, kcol + q * kc, n0);
                                else tma_load_2d(dst + q * b_sub, &tmB, fb, kcol + q * kc, n0);
                            }
                        }
                    }
                    kcol += S * kc;
                    r.advance(nstages);
                }
            }
        }
    } else if (warp == WARP_MMA && leader) {
        // ------------------------------------------------------------------ MMA issuer (pair: the leader CTA's only)
        Ring r;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t dhi = p.desc_hi;
        constexpr int kc = KK * 16;
        constexpr uint32_t a_sub = BM * kc * 2, a_stage = S * a_sub;
        const int groups = p.groups, nstages = p.stages, bn = p.bn;
        const bool resident = p.b_resident != 0;
        const uint32_t b_sub16 = p.b_sub >> 4, b_stage = p.b_stage, idesc = p.idesc;
        long long wf = 0, we = 0, mt0 = PROF ? clock64() : 0;
        if (resident) {
            mbar_wait(bfull, 0);
            tc_fence_after();
        }
        for (int tile = unit0; tile < total_tiles; tile += unit_step) {
            {
                PROF_T0();
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                PROF_ACC(we);
            }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * bn;
#pragma unroll 1
            for (int g = 0; g < groups; ++g) {
                {
                    PROF_T0();
                    mbar_wait(full0 + 8 * r.stage, r.phase);
                    PROF_ACC(wf);
                }
                tc_fence_after();
                // descriptor low words: (addr >> 4) | LBO(=1) << 16 ; a K step of 16 bf16 advances the address by 32 B
                const uint32_t a_lo = (((sA + r.stage * a_stage) >> 4) & 0x3FFF) | (1u << 16);
                const uint32_t b_lo = (((resident ? sB + g * b_stage : sB + r.stage * b_stage) >> 4) & 0x3FFF) | (1u << 16);
                if (elect_one()) {
#pragma unroll
                    for (int q = 0; q < S; ++q) {
#pragma unroll
                        for (int k = 0; k < KK; ++k) {
                            const uint64_t ad = make_desc(dhi, a_lo + q * (a_sub >> 4) + 2 * k);
                            const uint64_t bd = make_desc(dhi, b_lo + q * b_sub16 + 2 * k);
                            if (PAIR) umma_bf16_cg2(d_tmem, ad, bd, idesc, (g | q | k) != 0);
                            else umma_bf16(d_tmem, ad, bd, idesc, (g | q | k) != 0);
                        }
                    }
                    if (PAIR) {  // multicast: the same barrier offsets in both CTAs of the pair
                        umma_commit_cg2(empty0 + 8 * r.stage);
                        if (g == groups - 1) umma_commit_cg2(tfull0 + 8 * acc);
                    } else {
                        umma_commit(empty0 + 8 * r.stage);  // frees the smem slot when these MMAs retire
                        if (g == groups - 1) umma_commit(tfull0 + 8 * acc);  // accumulator complete -> epilogue
                    }
                }
                __syncwarp();
                r.advance(nstages);
            }
            if (++acc == p.acc_stages) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (PROF && lane == 0) {
            g_prof[blockIdx.x * PROF_SLOTS + 4] = wf;
            g_prof[blockIdx.x * PROF_SLOTS + 5] = we;
            g_prof[blockIdx.x * PROF_SLOTS + 6] = clock64() - mt0;
        }
    } else if (warp < WARP_EPI0 + NUM_EPI_WARPS) {
        // ------------------------------------------------------------------ epilogue
        // TMEM -> registers -> bias/activation/residual -> swizzled staging slab in smem -> TMA store.
        // Each warp owns 32 accumulator rows and two private 32-row x 64-byte slabs, so the only
        // synchronisation is __syncwarp + the bulk-group wait that recycles a slab.
        // The twelve warps form THREE groups of four (one warp per TMEM lane quarter); tile `it` of this CTA goes to
        // accumulator stage it % acc_stages and to the group of the same index.  Measured on the 1x1 layers: with loads, stores and
        // math switched off in turn, neither HBM nor the TMA stores set the pace - the epilogue warps did (8 warps =
        // 2 per scheduler, each a dependent chain barrier -> tcgen05.ld -> FFMA/MUFU -> st.shared -> fence -> TMA
        // store): more warps in flight is what raises the rate, and the tensor pipe has slack to run ahead into
        // the extra accumulator stage.
        const int e = warp - WARP_EPI0;
        const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the only ones this warp may read
        const int group = e >> 2;      // which accumulator stage
        const int n_tiles = p.n_tiles, bn = p.bn, Cout = p.Cout, act = p.act;
        const bool has_res = res != nullptr, out_f32 = p.out_f32 != 0, res_pre = p.res_pre != 0;
        const bool dbg_nostore = DBG && (p.dbg & 2) != 0, dbg_nomath = DBG && (p.dbg & 4) != 0;  // debug timing only
        const int slab_cols = out_f32 ? 16 : 32;  // 64 bytes of output per row
        const int n_slabs = (bn + slab_cols - 1) / slab_cols;
        const uint32_t stg0 = sC + e * 2 * STG_BYTES;
        const uint32_t swz = (uint32_t)((lane >> 1) & 3);  // 64B swizzle: 16-byte unit u of row r lives at u ^ ((r>>1)&3)
        const uint32_t my_row = stg0 + lane * 64;
        int buf = 0;
        int acc = 0;             // accumulator stage and phase of tile `it` (advanced for every tile, mine or not)
        uint32_t acc_phase = 0;
        int it = 0;
        long long ew = 0, et0 = PROF ? clock64() : 0;
        pdl_wait();  // residual reads and output stores touch buffers the previous kernel may still be using
        for (int tile = unit0; tile < total_tiles; tile += unit_step, ++it) {
            // group g owns accumulator stage g: every barrier has ONE waiting group that sees its phases in order (a
            // group waiting two phases ahead of a barrier would be fooled by the parity wrap-around).  With two
            // stages (bn > 170) the third group idles.
            const bool mine = acc == group;
            const int my_acc = acc;
            const uint32_t my_phase = acc_phase;
            if (++acc == p.acc_stages) {
                acc = 0;
                acc_phase ^= 1;
            }
            if (!mine) continue;
            const int mu = n_tiles == 1 ? tile : tile / n_tiles;
            const int n0 = n_tiles == 1 ? 0 : (tile - mu * n_tiles) * bn;
            const int mt = mu * NCTA + (int)cta_rank;  // this CTA's 128-row tile (pair, odd tile count: may be past M)
            const int m_warp = mt * BM + quarter * 32;
            const int m = m_warp + lane;
            const bool m_ok = m < p.M;
            size_t ridx = (size_t)m;
            if (p.res_up && has_res && m_ok) {  // nearest 2x: output pixel (h, w) reads low-resolution pixel (h/2, w/2)
                const int hw = p.Ho * p.Wo;
                const int bimg = m / hw, rem = m - bimg * hw;
                const int ph = rem / p.Wo, pw = rem - ph * p.Wo;
                ridx = ((size_t)bimg * (p.Ho >> 1) + (ph >> 1)) * (p.Wo >> 1) + (pw >> 1);
            }
            const __nv_bfloat16* rrow = res + ridx * p.res_pitch;
            // Residual of slab `s`, fetched ONE SLAB AHEAD (slab 0 before the accumulator wait): the four 16-byte loads of
            // a lane touch a row of their own (32 sectors per request) and come from L2 - issued at the top of their own
            // slab, their latency was exposed once per slab (1x1 512 -> 256 at 80x80, residual at half resolution:
            // 725 us against 498 us without the residual).
            uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0, q2 = q0, q3 = q0;
            auto res_fetch = [&](int s) {
                const int c0 = s * slab_cols, n = n0 + c0;
                if (s >= n_slabs || n >= Cout || !m_ok) return;
                const uint4* rp = reinterpret_cast<const uint4*>(rrow + n);
                const bool four = !out_f32 && c0 + 16 < bn && n + 16 < Cout;
                if (p.res_wide) {  // 32-byte aligned rows: 256-bit loads, half the requests (L2 only, like __ldcg below)
                    ldcg256(rp, q0, q1);
                    if (four) ldcg256(rp + 2, q2, q3);
                    return;
                }
                q0 = __ldcg(rp);  // L2 only: with PDL this SM's L1 may still hold the previous grid's lines
                q1 = __ldcg(rp + 1);
                if (four) {
                    q2 = __ldcg(rp + 2);
                    q3 = __ldcg(rp + 3);
                }
            };
            if (has_res && p.epi_mode == 0) res_fetch(0);
            {
                PROF_T0();
                mbar_wait(tfull0 + 8 * my_acc, my_phase);
                PROF_ACC(ew);
            }
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + my_acc * bn;
            if (p.epi_mode != 0) {
                // Detect head, last conv of a branch (head.py:94,103 -> :149-167): this thread's accumulator row IS the
                // logit vector of anchor `rem` of image `bimg`; decode it here and store straight into the prediction
                // tensor y[B, 4 + nc, A] (anchor-major rows: a warp's 32 pixels are 32 consecutive floats per channel).
                const int hw = p.Ho * p.Wo;
                const int mm = m_ok ? m : 0;
                const int bimg = mm / hw, rem = mm - bimg * hw;
                float* yb = ydet + (size_t)bimg * p.epi_rows * p.epi_A + p.epi_abase + rem;
                const size_t A = (size_t)p.epi_A;
                if (p.epi_mode == 1) {
#pragma unroll 1
                    for (int c0 = 0; c0 < bn; c0 += 16) {
                        const int n = n0 + c0;
                        if (n >= Cout) break;
                        uint32_t v[16];
                        tmem_ld16(t_row + c0, v);
                        tmem_ld_wait();
                        // sigmoid(acc + b) = 1 / (1 + 2^(-(acc + b) log2 e)) on packed pairs; the bias sits in shared
                        // memory pre-multiplied by -log2 e (see the prologue): FFMA2, 2 x ex2, FADD2, 2 x rcp per pair
                        float2 o[8];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 bq = reinterpret_cast<const float4*>(bias_s + n)[q];
                            const float2 nl2e = make_float2(-1.4426950408889634f, -1.4426950408889634f);
                            const float2 t0 = __ffma2_rn(make_float2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])),
                                                         nl2e, make_float2(bq.x, bq.y));
                            const float2 t1 = __ffma2_rn(make_float2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])),
                                                         nl2e, make_float2(bq.z, bq.w));
                            const float2 one = make_float2(1.f, 1.f);
                            const float2 d0 = __fadd2_rn(make_float2(ex2_fast(t0.x), ex2_fast(t0.y)), one);
                            const float2 d1 = __fadd2_rn(make_float2(ex2_fast(t1.x), ex2_fast(t1.y)), one);
                            o[2 * q] = make_float2(rcp_fast(d0.x), rcp_fast(d0.y));
                            o[2 * q + 1] = make_float2(rcp_fast(d1.x), rcp_fast(d1.y));
                        }
                        if (m_ok) {  // Cout is a multiple of 16: every column of the group is a real class
                            float* yc = yb + (size_t)(4 + n) * A;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                yc[(size_t)(2 * j) * A] = o[j].x;
                                yc[(size_t)(2 * j + 1) * A] = o[j].y;
                            }
                        }
                    }
                } else {
                    // DFL (block.py:76-79): softmax over the 16 bins of each side, expectation, then dist2bbox
                    // (tal.py:367-376) around the anchor centre (tal.py:352-364) and the level's stride
                    float dist[4];
#pragma unroll
                    for (int sd = 0; sd < 4; ++sd) {
                        uint32_t v[16];
                        tmem_ld16(t_row + 16 * sd, v);
                        tmem_ld_wait();
                        float f[16], mx = -INFINITY;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            f[j] = __uint_as_float(v[j]) + bias_s[16 * sd + j];
                            mx = fmaxf(mx, f[j]);
                        }
                        float sum = 0.f, wsum = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float e = __expf(f[j] - mx);
                            sum += e;
                            wsum = fmaf(e, (float)j, wsum);
                        }
                        dist[sd] = __fdividef(wsum, sum);
                    }
                    const int ph = rem / p.Wo, pw = rem - ph * p.Wo;
                    const float ax = (float)pw + 0.5f, ay = (float)ph + 0.5f;
                    const float x1 = ax - dist[0], y1 = ay - dist[1], x2 = ax + dist[2], y2 = ay + dist[3];
                    if (m_ok) {
                        yb[0] = (x1 + x2) * 0.5f * p.epi_stride;
                        yb[A] = (y1 + y2) * 0.5f * p.epi_stride;
                        yb[2 * A] = (x2 - x1) * p.epi_stride;
                        yb[3 * A] = (y2 - y1) * p.epi_stride;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_cluster(tempty0_sig + 8 * my_acc);
                    else mbar_arrive(tempty0 + 8 * my_acc);
                }
                continue;
            }
#pragma unroll 1
            for (int sl = 0; sl < n_slabs; ++sl) {
                const int c0 = sl * slab_cols;  // first column of the slab within the tile
                const int n = n0 + c0;
                if (n >= Cout) break;
                const bool two = !out_f32 && (c0 + 16 < bn);
                const uint4 r0 = q0, r1 = q1, r2 = q2, r3 = q3;
                if (has_res) res_fetch(sl + 1);
                uint32_t v0[16], v1[16];
                tmem_ld16(t_row + c0, v0);
                if (two) tmem_ld16(t_row + c0 + 16, v1);
                // the bulk store that read this slab two slabs ago must have drained it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                tmem_ld_wait();
                const uint32_t rowp = my_row + buf * STG_BYTES;
                if (dbg_nomath) {
                    st_shared_v4(rowp, v0[0], v0[1], v0[2], v0[3]);
                } else if (out_f32) {
                    float f[16];
                    epi_math16(v0, bias_s + n, act, false, r0, r1, f);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        st_shared_v4(rowp + ((q ^ swz) << 4), __float_as_uint(f[4 * q]), __float_as_uint(f[4 * q + 1]),
                                     __float_as_uint(f[4 * q + 2]), __float_as_uint(f[4 * q + 3]));
                } else {
                    float f[16];
                    uint32_t o[8];
                    epi_math16(v0, bias_s + n, act, has_res, r0, r1, f, p.out_scale, p.res_scale, res_pre);
                    pack16(f, o);
                    st_shared_v4(rowp + ((0 ^ swz) << 4), o[0], o[1], o[2], o[3]);
                    st_shared_v4(rowp + ((1 ^ swz) << 4), o[4], o[5], o[6], o[7]);
                    if (two) {
                        epi_math16(v1, bias_s + n + 16, act, has_res, r2, r3, f, p.out_scale, p.res_scale, res_pre);
                        pack16(f, o);
                        st_shared_v4(rowp + ((2 ^ swz) << 4), o[0], o[1], o[2], o[3]);
                        st_shared_v4(rowp + ((3 ^ swz) << 4), o[4], o[5], o[6], o[7]);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0 && !dbg_nostore && m_warp < p.M) {
                    tma_store_2d(&tmC, stg0 + buf * STG_BYTES, n, m_warp);  // rows >= M / cols >= Cout are clipped
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                buf ^= 1;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {  // this warp has drained its part of the accumulator
                if (PAIR) mbar_arrive_cluster(tempty0_sig + 8 * my_acc);
                else mbar_arrive(tempty0 + 8 * my_acc);
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
        if (PROF && warp == WARP_EPI0 && lane == 0) {
            g_prof[blockIdx.x * PROF_SLOTS + 7] = ew;
            g_prof[blockIdx.x * PROF_SLOTS + 8] = clock64() - et0;
        }
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all();  // the peer's shared memory and barriers are in use until the last MMA / arrive
    else __syncthreads();
    tc_fence_after();
    if (warp == WARP_ALLOC) {
        if (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host
#ifdef FCE_DEBUG
int g_debug_flags = 0;
#endif

int pick_kc(int cin) { return cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : 16); }

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

}  // namespace

// Launch counters of the convolution front door (fce_conv_stats): how many launches took which route since the last
// reset.  A bf16 conv that silently left the tensor cores (SIMT route) must be visible to the plan compiler / tests.
std::atomic<long long> g_conv_stats[4];  // 0 tcgen05 single CTA, 1 tcgen05 CTA pair, 2 tcgen05 3x3 strip kernel, 3 SIMT

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

int conv2d_halo_choice(const fce_conv_desc* d, bool has_res, int mode);  // 0 no strip kernel, 1 single CTA, 2 CTA pair
bool conv_halo_ran_last();
void conv_halo_clear_last();
#ifdef FCE_DEBUG
void conv_halo_set_profile(bool on);
int conv_halo_profile(long long* out, int n);
#endif
int conv2d_halo(const fce_conv_desc*, const void*, const void*, const float*, const void*, void*, cudaStream_t, int mode);

int conv2d_tc(const fce_conv_desc* d, const void* x, const void* w, const float* bias, const void* res, void* y,
              cudaStream_t st, const fce_detect_epi_desc* epi) {
    if (!bias) return FCE_ERR_BAD_ARG;
    // thin 3x3 stride-1 convs: input strip resident in shared memory (conv_halo.cu)
    // impl: 0 / 2 automatic kernel choice, 3 = CTA-pair kernel required (error when the shape has no legal pair tiling),
    // 4 = single-CTA implicit-GEMM kernel required (3 and 4 bypass the strip kernel), 5 / 6 = single-CTA / CTA-pair 3x3 strip
    // kernel required (tests, A/B timing)
    const bool force_pair = d->impl == 3, force_single = d->impl == 4;
    if (d->impl == 5 || d->impl == 6) {
        if (epi || !conv2d_halo_choice(d, res != nullptr, d->impl - 4)) return FCE_ERR_UNSUPPORTED;
        g_conv_stats[d->impl == 6 ? 1 : 2].fetch_add(1, std::memory_order_relaxed);
        return conv2d_halo(d, x, w, bias, res, y, st, d->impl - 4);
    }
    if (!epi && !force_pair && !force_single) {
        const int hc = conv2d_halo_choice(d, res != nullptr, 0);
        if (hc) {
            g_conv_stats[hc == 2 ? 1 : 2].fetch_add(1, std::memory_order_relaxed);
            return conv2d_halo(d, x, w, bias, res, y, st, 0);
        }
    }
    conv_halo_clear_last();
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
    // Kernel-selection knobs (they change WHICH kernel computes the same result, never the result): FCE_BN_MAX=128 caps
    // the N tile; FCE_1X1_CAP_K: 1x1 layers with at most this many input channels use 128-column tiles;
    // FCE_PAIR = 0 never / 1 whenever legal / unset: automatic.
    static const int bn_cap = [] { const int v = env_int("FCE_BN_MAX", 0); return v >= 32 ? v : 256; }();
    static const int cap1x1_k = env_int("FCE_1X1_CAP_K", 384);  // (384->512 at 80x80, batch 256: 955 -> 904 us with 128-column tiles)
    static const int pair_mode = env_int("FCE_PAIR", -1);
    // CTA pairs: legal when the pair's 256 x bn MMA exists (bn % 16, each CTA stages bn / 2 rows = whole 8-row swizzle
    // atoms) and there are at least two M tiles.
    const int k_steps = p.taps * p.chunks;
    auto plan_tiles = [&](bool pair) {
        // short-K 1x1 layers are HBM / epilogue bound: 128-column tiles keep three TMEM stages (and all twelve epilogue
        // warps) busy; re-reading the small [128, Cin] A tile for the second N tile is an L2 hit.  Long-K 1x1s and 3x3
        // layers keep 256 columns (their A tiles are large, and each extra N tile re-reads them).
        p.n_tiles = ceil_div(d->Cout, d->k == 1 && d->Cin <= cap1x1_k && bn_cap > 128 ? 128 : bn_cap);
        // N tiles that do not end the channel range must end on a 32-column staging-slab boundary
        p.bn = p.n_tiles == 1 ? d->Cout : ceil_div(ceil_div(d->Cout, p.n_tiles), 32) * 32;
        p.m_units = pair ? (p.m_tiles + 1) / 2 : p.m_tiles;
        const int workers = pair ? kNumSMs / 2 : kNumSMs;
        // small problems: narrower N tiles give the persistent grid more units to balance over the SMs
        while ((long long)p.m_units * p.n_tiles < 2 * workers && p.bn >= 128 && (p.bn / 2) % 32 == 0) {
            p.bn /= 2;
            p.n_tiles = ceil_div(d->Cout, p.bn);
        }
    };
    plan_tiles(false);
    bool pair = false;
    if ((pair_mode != 0 || force_pair) && !force_single && p.m_tiles >= 2) {
        plan_tiles(true);
        pair = p.bn % 16 == 0 && p.bn >= 32;
        if (pair && pair_mode < 0 && !force_pair) {
            // automatic (measured per layer on B200 with tools/plan_conv_ab.py, profiles/r02_conv_ab_*.csv): pairs win
            // wherever the operand feed - shared-memory reads of A + B per MMA, L2 -> SM weight traffic per tile - is the
            // limit: every 3x3 with Cin >= 64 (m scale, batch 256: 3x3/s2 512->512 1748 -> 1230 us, 256->256 1684 -> 1324 us,
            // 3x3/s1 128->128 146 -> 122 us) and 1x1 layers with K >= 512 (768->512: 341 -> 296 us).  Short-K 1x1 layers
            // are HBM / epilogue bound and LOSE 10-20 % to the coupling of the two CTAs' epilogues (256->256: 386 -> 460 us),
            // the 32-channel 3x3/s2 is bound by the TMA-im2col request rate either way.
            pair = (d->k == 3 ? d->Cin >= 64 : d->Cin >= env_int("FCE_PAIR_MIN_K1", 512)) &&
                   p.m_units * p.n_tiles >= kNumSMs / 4;
        }
        if (!pair) plan_tiles(false);
    }
    if (force_pair && !pair) return FCE_ERR_UNSUPPORTED;
    const int ncta = pair ? 2 : 1;
    // narrow K chunks: put three K steps behind one barrier round trip (one kernel row of a 3x3 / a 96-wide 1x1)
    p.S = (p.kc < 64 && k_steps % 3 == 0) ? 3 : 1;
    p.groups = k_steps / p.S;
    p.a_sub = BM * p.kc * 2;
    p.b_sub = (p.bn / ncta) * p.kc * 2;  // per CTA: a pair stages half of the weight tile's rows in each CTA
    p.a_stage = p.S * p.a_sub;
    p.b_stage = p.S * p.b_sub;
    const uint32_t w_bytes = (uint32_t)k_steps * p.b_sub;
    // Weights parked in shared memory for the life of the CTA when there is ONE N tile.  (Tried and dropped: with several
    // N tiles, a grid that is a multiple of n_tiles so that every CTA keeps one tile column and parks its [bn, K] slice.
    // Isolated, L2 flushed: 1x1 192 -> 256 at 160x160 1149 -> 1086 us, 256 -> 256 at 80x80 357 -> 338 us, but 384 -> 512
    // at 80x80 842 -> 916 us with only four A stages left; inside the captured step the m-scale forward came out
    // 0.5 % SLOWER, 10.29k vs 10.35k images/s on the same box - the streamed weight tiles are L2 hits that overlap the A
    // stream, and the parked slice costs pipeline depth.)
    p.b_resident = (p.n_tiles == 1 && w_bytes <= (uint32_t)B_RESIDENT_MAX) ? 1 : 0;
    p.bias_bytes = (uint32_t)(p.n_tiles * p.bn * 4 + 1023) & ~1023u;
    int room = SMEM_BUDGET - (int)p.bias_bytes;
    if (p.b_resident) {
        p.b_total = (w_bytes + 1023u) & ~1023u;
        p.stages = (room - (int)p.b_total) / (int)p.a_stage;
    } else {
        p.stages = room / (int)(p.a_stage + p.b_stage);
        p.b_total = 0;  // set below once the stage count is known
    }
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    if (p.stages < 2) return FCE_ERR_UNSUPPORTED;
    if (!p.b_resident) p.b_total = p.stages * p.b_stage;
    p.acc_stages = 3 * p.bn <= 512 ? 3 : 2;
    p.tmem_cols = 32;
    while (p.tmem_cols < (uint32_t)(p.acc_stages * p.bn)) p.tmem_cols <<= 1;
    p.out_pitch = d->out_pitch;
    p.res_pitch = d->res_pitch;
    p.res_wide = res && d->res_pitch % 16 == 0 && d->res_off % 16 == 0 && (reinterpret_cast<uintptr_t>(res) & 31) == 0;
    p.act = d->act;
    p.out_f32 = d->out_dtype == FCE_F32;
#ifdef FCE_DEBUG
    p.dbg = g_debug_flags;
#endif
    p.out_scale = d->weighted == 1 ? d->out_scale : 1.f;
    p.res_scale = d->weighted == 1 ? d->res_scale : 1.f;
    p.res_up = d->res_up;
    p.res_pre = d->weighted == 2 ? 1 : 0;
    if (epi) {
        if (p.n_tiles != 1 && epi->mode == 2) return FCE_ERR_UNSUPPORTED;
        p.epi_mode = epi->mode;
        p.epi_A = epi->A;
        p.epi_abase = epi->a_base;
        p.epi_rows = epi->rows;
        p.epi_stride = epi->stride;
    }
    const uint32_t row_bytes = p.kc * 2;                                        // swizzle span = K-chunk row
    const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);  // UMMA LayoutType
    const uint32_t sbo = 8 * row_bytes;                                         // 8-row core-matrix group pitch
    p.desc_hi = (sbo >> 4) | (1u << 14) | (layout << 29);
    // instruction descriptor: fp32 accumulate, bf16 A / B, K-major both, N >> 3 at bit 17, M >> 4 at bit 24 (M = 256 for a pair)
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)((BM * ncta) >> 4) << 24);

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
        static const int a_promo = env_int("FCE_A_L2PROMO", 128);
        cr = api.tiled(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)xin, gdim, gstr, box, est,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                       a_promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                      : (a_promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B),
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
        const cuuint32_t box[2] = {(cuuint32_t)p.kc, (cuuint32_t)(p.bn / ncta)};
        const cuuint32_t est[2] = {1, 1};
        cr = api.tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, est,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return FCE_ERR_UNSUPPORTED;
    }

    alignas(64) CUtensorMap tmC;
    if (epi) {
        tmC = tmA;  // never used for a store in the Detect epilogues; only prefetched
    } else {
        const bool f32 = d->out_dtype == FCE_F32;
        const cuuint64_t gdim[2] = {(cuuint64_t)d->Cout, (cuuint64_t)M};
        const cuuint64_t gstr[1] = {(cuuint64_t)d->out_pitch * (f32 ? 4 : 2)};
        const cuuint32_t box[2] = {(cuuint32_t)(f32 ? 16 : 32), 32};
        const cuuint32_t est[2] = {1, 1};
        void* yp = f32 ? (void*)(reinterpret_cast<float*>(y) + d->out_off)
                       : (void*)(reinterpret_cast<__nv_bfloat16*>(y) + d->out_off);
        cr = api.tiled(&tmC, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, yp, gdim, gstr,
                       box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return FCE_ERR_UNSUPPORTED;
    }
    const size_t smem = (size_t)p.stages * p.a_stage + p.b_total + NUM_EPI_WARPS * 2 * STG_BYTES + p.bias_bytes + 1024 + 256;
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams, const float*,
                             const __nv_bfloat16*, float*);
    // [single CTA | CTA pair] x (KK, S) variants: S = 3 only with narrow K chunks
    static const KernelFn table[2][5] = {
        {conv_tc_kernel<1, 1, false>, conv_tc_kernel<2, 1, false>, conv_tc_kernel<4, 1, false>,
         conv_tc_kernel<1, 3, false>, conv_tc_kernel<2, 3, false>},
        {conv_tc_kernel<1, 1, true>, conv_tc_kernel<2, 1, true>, conv_tc_kernel<4, 1, true>,
         conv_tc_kernel<1, 3, true>, conv_tc_kernel<2, 3, true>}};
    static DeviceOnce attr_once;  // the shared-memory opt-in is a per-device attribute
    int dev = 0;
    if (attr_once.pending(&dev)) {
        for (int a = 0; a < 2; ++a)
            for (int v = 0; v < 5; ++v) {
                cudaError_t e = cudaFuncSetAttribute(table[a][v], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                if (e != cudaSuccess) {
                    set_cuda_error(e);
                    return FCE_ERR_CUDA;
                }
            }
        attr_once.done(dev);
    }
    const int kk = p.kc >> 4;
    const int variant = p.S == 1 ? (kk == 1 ? 0 : (kk == 2 ? 1 : 2)) : (kk == 1 ? 3 : 4);
    const int total = p.m_units * p.n_tiles;
    const __nv_bfloat16* rp = res ? reinterpret_cast<const __nv_bfloat16*>(res) + d->res_off : nullptr;
    float* ydet = epi ? reinterpret_cast<float*>(y) : nullptr;
    g_conv_stats[pair ? 1 : 0].fetch_add(1, std::memory_order_relaxed);
    if (pair) {
        const int pairs = total < kNumSMs / 2 ? total : kNumSMs / 2;
        return launch_pdl_cluster(table[1][variant], 2 * pairs, NUM_THREADS, smem, st, 2, tmA, tmB, tmC, p, bias, rp, ydet);
    }
    const int grid = total < kNumSMs ? total : kNumSMs;
    return launch_pdl(table[0][variant], grid, NUM_THREADS, smem, st, tmA, tmB, tmC, p, bias, rp, ydet);
}

#ifdef FCE_DEBUG
void conv_tc_set_profile(int on) {
    conv_halo_set_profile((on & 1) != 0);
    g_debug_flags = on >> 1;  // bit 1 of `on`: skip TMA loads (debug timing only - results are garbage)
}

int conv_tc_profile(long long* out, int n) {
    if (conv_halo_ran_last()) return conv_halo_profile(out, n);
    if (n > kNumSMs * PROF_SLOTS) n = kNumSMs * PROF_SLOTS;
    cudaError_t e = cudaMemcpyFromSymbol(out, g_prof, (size_t)n * sizeof(long long));
    if (e != cudaSuccess) {
        set_cuda_error(e);
        return FCE_ERR_CUDA;
    }
    return n;
}
#endif

}  // namespace fce

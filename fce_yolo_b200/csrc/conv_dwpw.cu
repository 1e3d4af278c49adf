// Depthwise 3x3 + pointwise 1x1 in ONE pass: Detect's class branch cv3[i][j] = Sequential(DWConv(c, c, 3), Conv(c, c3, 1))
// (ultralytics/nn/modules/head.py:101-102; DWConv conv.py:185-199, Conv.forward_fuse conv.py:80-89).
//
// As two launches the pair moves the depthwise result through HBM twice (written by fce_dwconv3x3, read back by the 1x1:
// at m scale, batch 256, 80x80x256 that is 2 x 0.84 GB and two kernels of 0.48 + 0.34 ms, the first instruction-bound at
// half of the copy bandwidth).  Here the depthwise result never leaves the SM: CUDA-core warps compute it tile by tile
// straight into the shared-memory A operand of the 1x1, which runs on the tensor cores underneath them.
//
//   unit   = TH x 16 output pixels of one image (TH = 8, or 7 when the parked weights leave less room): M = 128 rows of
//            one tcgen05.mma (rows past TH * 16 are never stored)
//   K loop = C / 64 chunks of 64 channels (128-byte rows):
//     TMA warp      one tiled 4-D box {64 channels, 18 columns, TH + 2 rows} of the NHWC input per (unit, chunk); the TMA
//                   unit's out-of-bounds zero fill is the conv padding.  Ring of three to eight stages (what fits next to the
//                   parked weights).
//     DW teams      two teams of eight warps; team t takes chunks t, t + 2, ... of the (unit, chunk) sequence.  A warp owns two
//                   adjacent columns, a lane two channels (every shared-memory access of a warp is one 128-byte pixel row:
//                   no bank conflicts); it walks down the box rows with three rolling accumulator rows on
//                   packed fp32 pairs (the arithmetic of dwconv_tma.cu, same operation order: bit-identical values), then
//                   bias + SiLU -> bf16 -> its team's A stage [128 rows x 128 bytes], 128B-swizzled K-major - exactly the
//                   layout a TMA load of the stored map would have produced.
//     MMA warp      4 x tcgen05.mma (M = 128, N = Cout, K = 16) per chunk from the A stage and the PARKED 1x1 weights
//                   ([Cout, C] bf16, loaded once per CTA; where they do not fit - 512 -> 256: 256 KB - chunk c of the weights
//                   is streamed with input chunk c through a three-stage ring), accumulator in TMEM, two accumulator stages.
//   epilogue      eight warps (two per TMEM lane quarter, half of the channels each): tcgen05.ld -> bias + SiLU -> bf16 ->
//                   32-byte global stores (one pixel row per thread; no staging buffer: shared memory goes to the parked
//                   weights).
//
//   Measured and dropped (git history, commit 7b7fe32): a CTA-pair variant (cta_group::2, half of the weight rows parked
//   per CTA so that five input stages fit at m scale) - the coupling of the two CTAs' producer teams through the joint
//   M = 256 MMA cost more than the deeper ring bought (m scale P3 block, batch 256: 690 us against 520 us for single CTAs
//   with three stages; switching the input loads off entirely changes nothing: the kernel is not paced by them).
//
// Algorithmic HBM bytes per output pixel: 2 * C read (+ 1/8 .. 1/4 halo re-read, an L2 hit) + 2 * Cout written.
// Bound (m scale, C = Cout = 256, batch 256): HBM 0.26 ms at the copy peak.  Measured 0.52 ms (two launches: 0.74 ms): the
// SM side has no single saturated pipe (ncu: issue slots 58 %, XU 40 %, FMA 36 %, L1 / shared 68 %, tensor 21 %; the
// depthwise warps are busy 93 % of the time at ~11 cycles per issued instruction, 1.4 eligible warps per scheduler) -
// ~21 k warp instructions per unit, of which ~13.5 k are the arithmetic itself.
#include "tc_common.cuh"

namespace fce {
using namespace tc;
namespace {

constexpr int TW = 16;                 // output columns of a unit
constexpr int BC = TW + 2;             // input box columns
constexpr int KC = 64;                 // channels per K chunk (128-byte rows, 128B swizzle)
constexpr int TEAMS = 2, TEAM_WARPS = TW / 2;  // one warp per pair of output columns
constexpr int WARP_EPI0 = TEAMS * TEAM_WARPS;  // 16 .. 23: warp & 3 = TMEM lane quarter, (warp - 16) >> 2 = channel half
constexpr int NUM_EPI_WARPS = 8;
constexpr int WARP_TMA = WARP_EPI0 + NUM_EPI_WARPS, WARP_MMA = WARP_TMA + 1;
constexpr int NUM_THREADS = (WARP_MMA + 1) * 32;  // 832
constexpr int MAX_IN_STAGES = 8;
constexpr int W_STAGES = 3;            // streamed weights: ring of [Cout, 64] chunks
constexpr uint32_t A_STAGE = 128 * 128;  // one A stage: 128 rows x 128 bytes
constexpr int SMEM_LIMIT = 227 * 1024;

struct DwpwParams {
    int B, H, W, C, Cout;
    int tiles_h, tiles_w, units, chunks;
    int in_stages;      // depth of the input ring
    int out_pitch;
    int dw_act, pw_act;
    uint32_t in_stage;  // bytes of one input stage = one TMA box
    uint32_t w_chunk;   // bytes of one weight chunk: Cout x 128
    int w_stream;       // the [Cout, C] weights do not fit next to the pipeline: chunk c travels with input chunk c (W_STAGES ring)
    uint32_t bias_bytes, tmem_cols;
    uint32_t desc_hi, idesc;
    int wide_store;     // output rows 32-byte aligned: 256-bit stores
    int dbg;            // -DFCE_DEBUG builds (FCE_DWPW_DBG): 1 no input loads, 2 no depthwise math, 4 no epilogue - GARBAGE results, timing only
};
#ifdef FCE_DEBUG
constexpr bool DBG = true;
// per-CTA cycle accounting: [0] DW warp 0 wait-input [1] wait-A-empty [2] total | [4] MMA wait-A-full [5] wait-acc-empty [6] total
// | [8] epilogue warp 0 wait-acc-full [9] total
__device__ long long g_dprof[kNumSMs * 16];
#else
constexpr bool DBG = false;
__device__ long long g_dprof[1];
#endif
#define DP_T0() long long _t0 = 0; if (DBG) _t0 = clock64()
#define DP_ACC(var) if (DBG) (var) += clock64() - _t0

__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* o) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                 "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                 : "memory");
}

template <int TH>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_dwpw_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const DwpwParams p,
                 const float* __restrict__ w_dw, const float* __restrict__ b_dw, const float* __restrict__ b_pw,
                 __nv_bfloat16* __restrict__ y) {
    constexpr int BR = TH + 2;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw0 = smem_u32(smem_raw);
    const uint32_t base = (raw0 + 1023u) & ~1023u;
    const uint32_t sA = base;                              // TEAMS x A_STAGE
    const uint32_t sW = sA + TEAMS * A_STAGE;              // chunks (parked) or W_STAGES (streamed) x [Cout][64] bf16, 128B-swizzled
    const uint32_t sIn = sW + (p.w_stream ? W_STAGES : p.chunks) * p.w_chunk;  // in_stages x [BR][BC][64] bf16
    const uint32_t sBias = sIn + p.in_stages * p.in_stage;
    const uint32_t bars = sBias + p.bias_bytes;
    const uint32_t in_full0 = bars, in_empty0 = bars + 8 * MAX_IN_STAGES;
    const uint32_t a_full0 = in_empty0 + 8 * MAX_IN_STAGES, a_empty0 = a_full0 + 8 * TEAMS;
    const uint32_t tfull0 = a_empty0 + 8 * TEAMS, tempty0 = tfull0 + 16;
    const uint32_t wfull = tempty0 + 16, tmem_slot = wfull + 8;
    const uint32_t ws_full0 = tmem_slot + 8, ws_empty0 = ws_full0 + 8 * W_STAGES;  // streamed-weight ring
    float* bias_s = reinterpret_cast<float*>(smem_raw + (sBias - raw0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item0 = (int)blockIdx.x, item_step = (int)gridDim.x, n_items = p.units;  // persistent: units b, b + grid, ...
    const int nst = p.in_stages;

    pdl_launch_dependents();
    if (warp == WARP_TMA && lane == 0) {
        for (int i = 0; i < nst; ++i) {
            mbar_init(in_full0 + 8 * i, 1);
            mbar_init(in_empty0 + 8 * i, TEAM_WARPS);  // the four warps of the team that consumed the stage
        }
        for (int t = 0; t < TEAMS; ++t) {
            mbar_init(a_full0 + 8 * t, TEAM_WARPS);
            mbar_init(a_empty0 + 8 * t, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, NUM_EPI_WARPS);
        }
        mbar_init(wfull, 1);
        for (int i = 0; i < W_STAGES; ++i) {
            mbar_init(ws_full0 + 8 * i, 1);
            mbar_init(ws_empty0 + 8 * i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        const float bsc = epi_bias_scale(p.pw_act);  // pre-scaled for epi_math16
        for (int i = threadIdx.x; i < (int)(p.bias_bytes >> 2); i += NUM_THREADS) bias_s[i] = i < p.Cout ? b_pw[i] * bsc : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int chunks = p.chunks;
    auto unit_coords = [&](int u, int& b, int& h0, int& w0) {
        const int tw = u % p.tiles_w;
        const int r = u / p.tiles_w;
        const int th = r % p.tiles_h;
        b = r / p.tiles_h;
        h0 = th * TH;
        w0 = tw * TW;
    };

    if (warp == WARP_TMA) {
        // ------------------------------------------------------------------ loads: parked 1x1 weights, then input boxes
        const bool w_stream = p.w_stream != 0;
        if (!w_stream && elect_one()) {
            mbar_expect_tx(wfull, (uint32_t)chunks * p.w_chunk);
            for (int c = 0; c < chunks; ++c) tma_load_2d(sW + c * p.w_chunk, &tmW, wfull, c * KC, 0);
        }
        __syncwarp();
        pdl_wait();  // activations come from the previous kernel (the weights above are constants)
        Ring r, rw;
        for (int it = item0; it < n_items; it += item_step) {
            int b, h0, w0;
            unit_coords(it, b, h0, w0);
            for (int c = 0; c < chunks; ++c) {
                if (w_stream) {  // this chunk's weights (an L2 hit after the first unit): consumed by the MMA warp
                    mbar_wait(ws_empty0 + 8 * rw.stage, rw.phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(ws_full0 + 8 * rw.stage, p.w_chunk);
                        tma_load_2d(sW + rw.stage * p.w_chunk, &tmW, ws_full0 + 8 * rw.stage, c * KC, 0);
                    }
                    __syncwarp();
                    rw.advance(W_STAGES);
                }
                mbar_wait(in_empty0 + 8 * r.stage, r.phase ^ 1);
                if (elect_one()) {
                    const uint32_t fb = in_full0 + 8 * r.stage;
                    if (DBG && (p.dbg & 1)) {
                        mbar_arrive(fb);
                    } else {
                        mbar_expect_tx(fb, p.in_stage);
                        tma_load_4d(sIn + r.stage * p.in_stage, &tmX, fb, c * KC, w0 - 1, h0 - 1, b);
                    }
                }
                __syncwarp();
                r.advance(nst);
            }
        }
    } else if (warp == WARP_MMA) {
        // ------------------------------------------------------------------ MMA issuer
        const bool w_stream = p.w_stream != 0;
        if (!w_stream) {
            mbar_wait(wfull, 0);
            tc_fence_after();
        }
        Ring rw;
        int acc = 0, turn = 0;
        uint32_t acc_phase = 0, a_phase = 0;  // bit t of a_phase: parity of team t's A-full barrier
        const uint32_t dhi = p.desc_hi, idesc = p.idesc;
        long long m_wa = 0, m_we = 0, m_t0 = DBG ? clock64() : 0;
        for (int it = item0; it < n_items; it += item_step) {
            {
                DP_T0();
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                DP_ACC(m_we);
            }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * p.Cout;
#pragma unroll 1
            for (int c = 0; c < chunks; ++c) {
                {
                    DP_T0();
                    mbar_wait(a_full0 + 8 * turn, (a_phase >> turn) & 1u);
                    DP_ACC(m_wa);
                }
                if (w_stream) mbar_wait(ws_full0 + 8 * rw.stage, rw.phase);
                tc_fence_after();
                const uint32_t a_lo = (((sA + turn * A_STAGE) >> 4) & 0x3FFF) | (1u << 16);
                const uint32_t b_lo = (((sW + (w_stream ? rw.stage : c) * p.w_chunk) >> 4) & 0x3FFF) | (1u << 16);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < KC / 16; ++k)
                        umma_bf16(d_tmem, make_desc(dhi, a_lo + 2 * k), make_desc(dhi, b_lo + 2 * k), idesc, (c | k) != 0);
                    umma_commit(a_empty0 + 8 * turn);                   // the team may overwrite its A stage
                    if (w_stream) umma_commit(ws_empty0 + 8 * rw.stage);  // the weight stage may be refilled
                    if (c == chunks - 1) umma_commit(tfull0 + 8 * acc);  // accumulator complete -> epilogue
                }
                __syncwarp();
                a_phase ^= 1u << turn;
                if (++turn == TEAMS) turn = 0;
                if (w_stream) rw.advance(W_STAGES);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (DBG && lane == 0) {
            g_dprof[blockIdx.x * 16 + 4] = m_wa;
            g_dprof[blockIdx.x * 16 + 5] = m_we;
            g_dprof[blockIdx.x * 16 + 6] = clock64() - m_t0;
        }
    } else if (warp < WARP_EPI0) {
        // ------------------------------------------------------------------ depthwise teams: input box -> A stage
        const int team = warp / TEAM_WARPS;
        const int col = 2 * (warp - team * TEAM_WARPS);  // this WARP owns columns col, col + 1; lane = channel pair 2 * lane
        const uint8_t* in_gen = smem_raw + (sIn - raw0) + lane * 4;
        // A-stage byte offset of (row m, this lane's 4 bytes): 16-byte chunk (lane >> 2) ^ (m & 7), m & 7 = (col + q) & 7.
        // A warp store covers one whole 128-byte row: conflict-free, like the 128-byte pixel reads.
        const uint32_t a_col[2] = {
            sA + team * A_STAGE + (uint32_t)col * 128u + (uint32_t)((((lane >> 2) ^ (col & 7)) << 4) | ((lane & 3) << 2)),
            sA + team * A_STAGE + (uint32_t)(col + 1) * 128u +
                (uint32_t)((((lane >> 2) ^ ((col + 1) & 7)) << 4) | ((lane & 3) << 2))};
        const bool silu = p.dw_act == FCE_ACT_SILU;
        int turn = 0;
        Ring rin;
        uint32_t a_phase = 0;
        long long d_wi = 0, d_wa = 0, d_t0 = DBG ? clock64() : 0;
        for (int it = item0; it < n_items; it += item_step) {
            int b, h0, w0;
            unit_coords(it, b, h0, w0);
            const bool active = w0 + col < p.W && !(DBG && (p.dbg & 2));  // ragged last column tile (warp-uniform): rows stay stale
            for (int c = 0; c < chunks; ++c) {
                if (turn == team) {
                    // depthwise taps / bias of this lane's two channels: constants (L1 hits), in flight across the wait below
                    // With SiLU the taps and the bias are halved (exact), so the accumulator IS h = a / 2 of silu(a) = h + h * tanh(h):
                    // the same values as dwconv_tma.cu's a * 0.5 afterwards, one multiply per pair less.
                    float2 wt[9], bs;
                    {
                        const int ch = c * KC + lane * 2;
                        const float2 hs = silu ? make_float2(0.5f, 0.5f) : make_float2(1.f, 1.f);
#pragma unroll
                        for (int t = 0; t < 9; ++t) wt[t] = __fmul2_rn(__ldg(reinterpret_cast<const float2*>(w_dw + t * p.C + ch)), hs);
                        bs = __fmul2_rn(__ldg(reinterpret_cast<const float2*>(b_dw + ch)), hs);
                    }
                    {
                        DP_T0();
                        mbar_wait(in_full0 + 8 * rin.stage, rin.phase);
                        DP_ACC(d_wi);
                    }
                    if (active) {
                        const uint8_t* tile = in_gen + rin.stage * p.in_stage;
                        float2 acc[3][2];
#pragma unroll
                        for (int k = 0; k < 3; ++k) acc[k][0] = acc[k][1] = bs;
#pragma unroll
                        for (int br = 0; br < BR; ++br) {
                            float2 v[4];
                            const uint8_t* rowp = tile + (br * BC + col) * 128;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t r = *reinterpret_cast<const uint32_t*>(rowp + k * 128);
                                v[k] = make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
                            }
                            // box row br is tap row kh of output row br - kh (same operation order as dwconv_tma.cu)
#pragma unroll
                            for (int kh = 0; kh < 3; ++kh) {
                                const int o = br - kh;
                                if (o < 0 || o >= TH) continue;
#pragma unroll
                                for (int q = 0; q < 2; ++q)
#pragma unroll
                                    for (int kw = 0; kw < 3; ++kw)
                                        acc[o % 3][q] = __ffma2_rn(v[q + kw], wt[3 * kh + kw], acc[o % 3][q]);
                            }
                            if (br == 2) {  // first store of this chunk: the MMAs that read this A stage (two chunks ago) have
                                DP_T0();    // retired - they ran under rows 0..2 above
                                mbar_wait(a_empty0 + 8 * team, a_phase ^ 1);
                                DP_ACC(d_wa);
                            }
                            if (br >= 2) {
                                const int o = br - 2;  // finished output row
#pragma unroll
                                for (int q = 0; q < 2; ++q) {
                                    float2 out = acc[o % 3][q];
                                    if (silu)  // h + h * tanh(h), the accumulator is h: one MUFU per element
                                        out = __ffma2_rn(out, make_float2(tanh_fast(out.x), tanh_fast(out.y)), out);
                                    acc[o % 3][q] = bs;
                                    __nv_bfloat162 ob = __floats2bfloat162_rn(out.x, out.y);
                                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_col[q] + (uint32_t)o * (TW * 128u)),
                                                 "r"(*reinterpret_cast<uint32_t*>(&ob))
                                                 : "memory");
                                }
                            }
                        }
                    }
                    else mbar_wait(a_empty0 + 8 * team, a_phase ^ 1);  // (keeps the barrier phases in step)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(in_empty0 + 8 * rin.stage);
                        mbar_arrive(a_full0 + 8 * team);
                    }
                    a_phase ^= 1;
                }
                rin.advance(nst);
                if (++turn == TEAMS) turn = 0;
            }
        }
        if (DBG && warp == 0 && lane == 0) {
            g_dprof[blockIdx.x * 16 + 0] = d_wi;
            g_dprof[blockIdx.x * 16 + 1] = d_wa;
            g_dprof[blockIdx.x * 16 + 2] = clock64() - d_t0;
        }
    } else {
        // ------------------------------------------------------------------ epilogue: TMEM -> bias + act -> bf16 -> global
        // eight warps: warp & 3 = TMEM lane quarter (32 pixels, one per lane), two warps per quarter split the channel range
        const int quarter = warp & 3, half = (warp - WARP_EPI0) >> 2;
        const int m = quarter * 32 + lane, r = m >> 4, cc = m & 15;
        const int act = p.pw_act, Cout = p.Cout;
        const int split = ((Cout / 16 + 1) / 2) * 16;
        const int n_lo = half ? split : 0, n_hi = half ? Cout : split;
        const bool wide = p.wide_store != 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        pdl_wait();  // the output buffer may still be in use by the previous kernel (arena buffers are recycled)
        long long e_w = 0, e_t0 = DBG ? clock64() : 0;
        for (int it = item0; it < n_items; it += item_step) {
            int b, h0, w0;
            unit_coords(it, b, h0, w0);
            const bool ok = r < TH && h0 + r < p.H && w0 + cc < p.W;
            __nv_bfloat16* yrow = y + (((size_t)b * p.H + (h0 + r)) * p.W + (w0 + cc)) * p.out_pitch;
            {
                DP_T0();
                mbar_wait(tfull0 + 8 * acc, acc_phase);
                DP_ACC(e_w);
            }
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * Cout;
            const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll 1
            for (int n = n_lo; n < ((DBG && (p.dbg & 4)) ? n_lo : n_hi); n += 32) {
                const bool two = n + 16 < n_hi;
                uint32_t v0[16], v1[16];
                tmem_ld16(t_row + n, v0);
                if (two) tmem_ld16(t_row + n + 16, v1);
                tmem_ld_wait();
                float f[16];
                uint32_t o[8];
                epi_math16(v0, bias_s + n, act, false, z, z, f);
                pack16(f, o);
                if (ok) {
                    if (wide) st_global_v8(yrow + n, o);
                    else {
                        *reinterpret_cast<uint4*>(yrow + n) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4*>(yrow + n + 8) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
                if (two) {
                    epi_math16(v1, bias_s + n + 16, act, false, z, z, f);
                    pack16(f, o);
                    if (ok) {
                        if (wide) st_global_v8(yrow + n + 16, o);
                        else {
                            *reinterpret_cast<uint4*>(yrow + n + 16) = make_uint4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<uint4*>(yrow + n + 24) = make_uint4(o[4], o[5], o[6], o[7]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (DBG && warp == WARP_EPI0 && lane == 0) {
            g_dprof[blockIdx.x * 16 + 8] = e_w;
            g_dprof[blockIdx.x * 16 + 9] = clock64() - e_t0;
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == WARP_MMA)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
}

size_t dwpw_smem(const DwpwParams& p) {
    return (size_t)TEAMS * A_STAGE + (size_t)(p.w_stream ? W_STAGES : p.chunks) * p.w_chunk + (size_t)p.in_stages * p.in_stage +
           p.bias_bytes + 512 + 1024;
}

// Ring depth a tile height leaves room for next to the weights (parked, or the streamed ring); 0 = does not fit.
int dwpw_stages(const fce_dwpw_desc* d, int th, bool stream) {
    const size_t fixed = (size_t)TEAMS * A_STAGE + (size_t)(stream ? W_STAGES : d->C / KC) * (size_t)d->Cout * 128u +
                         (((size_t)d->Cout * 4u + 255u) & ~(size_t)255u) + 512 + 1024;
    if (fixed >= (size_t)SMEM_LIMIT) return 0;
    const int st = (int)(((size_t)SMEM_LIMIT - fixed) / ((size_t)(th + 2) * BC * 128u));
    return st > MAX_IN_STAGES ? MAX_IN_STAGES : st;
}

// Shape rules, tile height and ring depth; pure arithmetic (no CUDA calls): also behind fce_dwpw_route.
// Returns the tile height (0 = not a shape of this kernel).
int dwpw_plan(const fce_dwpw_desc* d, DwpwParams& p) {
    if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0 || d->Cout <= 0) return 0;
    if (d->C % KC || d->Cout % 16 || d->Cout > 256) return 0;
    if (d->in_pitch % 8 || d->in_off % 8 || d->out_pitch % 8 || d->out_off % 8) return 0;
    if (d->dw_act != FCE_ACT_SILU && d->dw_act != FCE_ACT_NONE) return 0;
    if (d->pw_act != FCE_ACT_SILU && d->pw_act != FCE_ACT_NONE && d->pw_act != FCE_ACT_SIGMOID) return 0;
    if (d->H > 32000 || d->W > 32000) return 0;
    // Weights parked for the life of the CTA where they fit next to at least three input stages (two being consumed, one in
    // flight); otherwise chunk c of the weights is streamed with input chunk c (m scale: the 512 -> 256 blocks, 256 KB of
    // weights - 32 KB per chunk and unit from L2, 13 B / clk / SM next to a depthwise stage that takes ~2 k cycles).
    int TH = 0, stages = 0;
    bool stream = false;
    for (int pass = 0; pass < 2 && !TH; ++pass)
        for (int th = 8; th >= 7 && !TH; --th) {
            const int st = dwpw_stages(d, th, pass == 1);
            if (st >= 3) {
                TH = th;
                stages = st;
                stream = pass == 1;
            }
        }
    if (!TH) return 0;
    p.w_stream = stream ? 1 : 0;
    p.B = d->B; p.H = d->H; p.W = d->W; p.C = d->C; p.Cout = d->Cout;
    p.chunks = d->C / KC;
    p.w_chunk = (uint32_t)d->Cout * 128u;
    p.bias_bytes = ((uint32_t)d->Cout * 4u + 255u) & ~255u;
    p.out_pitch = d->out_pitch;
    p.dw_act = d->dw_act;
    p.pw_act = d->pw_act;
    p.in_stage = (uint32_t)(TH + 2) * BC * 128u;
    p.in_stages = stages;
    p.tiles_h = (d->H + TH - 1) / TH;
    p.tiles_w = (d->W + TW - 1) / TW;
    const long long units = (long long)d->B * p.tiles_h * p.tiles_w;
    if (units > 0x7fffffffLL) return 0;
    p.units = (int)units;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2u * (uint32_t)d->Cout) p.tmem_cols <<= 1;
    p.desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 8 rows x 128 bytes, descriptor version 1, 128B swizzle
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(d->Cout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    return TH;
}

}  // namespace
}  // namespace fce

using namespace fce;

#ifdef FCE_DEBUG
extern "C" int fce_dwpw_profile(long long* out, int n) {  // debug builds only, not in the public header
    if (n > kNumSMs * 16) n = kNumSMs * 16;
    return cudaMemcpyFromSymbol(out, g_dprof, (size_t)n * sizeof(long long)) == cudaSuccess ? n : FCE_ERR_CUDA;
}
#endif

extern "C" int fce_dwpw_route(const fce_dwpw_desc* d) {
    if (!d) return FCE_ERR_BAD_ARG;
    DwpwParams p{};
    return dwpw_plan(d, p) ? 1 : 0;
}

extern "C" int fce_dwpw_conv(const fce_dwpw_desc* d, const void* x, const float* w_dw, const float* b_dw, const void* w_pw,
                             const float* b_pw, void* y, void* stream) {
    if (!d || !x || !w_dw || !b_dw || !w_pw || !b_pw || !y) return FCE_ERR_BAD_ARG;
    if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0 || d->Cout <= 0) return FCE_ERR_BAD_ARG;
    DwpwParams p{};
    const int TH = dwpw_plan(d, p);
    if (!TH) return FCE_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(w_pw) || !aligned16(y) || !aligned16(w_dw) || !aligned16(b_dw)) return FCE_ERR_ALIGNMENT;
    const DriverApi& api = driver();
    if (!api.ok) return FCE_ERR_CUDA;
    const __nv_bfloat16* xin = reinterpret_cast<const __nv_bfloat16*>(x) + d->in_off;
    __nv_bfloat16* yout = reinterpret_cast<__nv_bfloat16*>(y) + d->out_off;
#ifdef FCE_DEBUG
    {
        const char* e = getenv("FCE_DWPW_DBG");
        p.dbg = e && *e ? atoi(e) : 0;
    }
#endif
    p.wide_store = (d->out_pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(yout) & 31) == 0) ? 1 : 0;
    alignas(64) CUtensorMap tmX, tmW;
    {
        const cuuint64_t gdim[4] = {(cuuint64_t)d->C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
        const cuuint64_t gstr[3] = {(cuuint64_t)d->in_pitch * 2, (cuuint64_t)d->W * d->in_pitch * 2,
                                    (cuuint64_t)d->H * d->W * d->in_pitch * 2};
        const cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)BC, (cuuint32_t)(TH + 2), 1};
        const cuuint32_t est[4] = {1, 1, 1, 1};
        if (api.tiled(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)xin, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCE_ERR_UNSUPPORTED;
    }
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)d->C, (cuuint64_t)d->Cout};
        const cuuint64_t gstr[1] = {(cuuint64_t)d->C * 2};
        const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)d->Cout};
        const cuuint32_t est[2] = {1, 1};
        if (api.tiled(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_pw), gdim, gstr, box, est,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCE_ERR_UNSUPPORTED;
    }
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const DwpwParams, const float*, const float*, const float*,
                             __nv_bfloat16*);
    static const KernelFn table[2] = {conv_dwpw_kernel<7>, conv_dwpw_kernel<8>};
    static DeviceOnce attr_once;  // the shared-memory opt-in is a per-device attribute
    int dev = 0;
    if (attr_once.pending(&dev)) {
        for (int v = 0; v < 2; ++v) {
            cudaError_t e = cudaFuncSetAttribute(table[v], cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
            if (e != cudaSuccess) {
                set_cuda_error(e);
                return FCE_ERR_CUDA;
            }
        }
        attr_once.done(dev);
    }
    const int grid = p.units < kNumSMs ? p.units : kNumSMs;
    return launch_pdl(table[TH - 7], grid, NUM_THREADS, dwpw_smem(p), (cudaStream_t)stream, tmX, tmW, p, w_dw, b_dw, b_pw, yout);
}

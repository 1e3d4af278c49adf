// 3x3 stride-1 convolution with the INPUT STRIP RESIDENT IN SHARED MEMORY ("halo" kernel), tcgen05 + TMEM + TMA.
//
// The TMA-im2col kernel (conv_tc.cu) fetches every input pixel nine times (once per tap) through L2 and pays the
// TMA unit's per-pixel request cost nine times; for the thin 3x3 convs of the C3k2 / C3k bottlenecks and the Detect
// box branch (conv.py:80-89 inside block.py:474-476, head.py:92-94) that cost, not HBM or the tensor pipe, set
// the pace.  Here a CTA owns a band of R output rows of one image:
//
//   * ONE tiled-mode TMA box {kc channels, W+2 columns, R+2 rows} per K chunk lands the zero-padded input strip in
//     shared memory as a "padded-flat" pixel list  j = row * (W+2) + col  (the TMA unit's out-of-bounds zero fill
//     produces the padding columns and rows for free);
//   * output pixel o = ro * (W+2) + co reads tap (kh, kw) at strip row  o + kh*(W+2) + kw, so the A operand of
//     every tap is the SAME strip seen through a shifted shared-memory descriptor: 9 * kc/16 tcgen05.mma per
//     128-row block and K chunk, no data movement between taps;
//   * the two padding columns of every row produce garbage accumulator rows that are simply never stored;
//   * the whole [Cout, 9*Cin] weight matrix is parked in shared memory for the life of the CTA - or, when it does not
//     fit (Detect box branch: 3x3 with Cin = 128..512 -> 64), the nine tap tiles of K chunk c are streamed together
//     with strip chunk c into the same ring stage: weights cross L2->SM once per band instead of once per 128-pixel
//     tile, and the input (R+2)/R times instead of nine (the TMA-im2col kernel ran these layers at 37 % of their
//     roofline, bound by the TMA request rate).
//
// Each input element crosses L2 -> SM (R+2)/R times instead of 9, and the producer issues one TMA per
// (band, chunk) instead of nine per 128 pixels.  Warp roles as in conv_tc.cu.
//
// CTA PAIRS (template PAIR, cta_group::2): the two CTAs of a {2,1,1} cluster each own a band (units 2q and 2q+1) with the
// SAME geometry, stage their own strip and HALF of the rows of every weight tile; the leader issues tcgen05.mma of
// M = 256 whose shifted-strip descriptors address both CTAs' shared memory at the same offsets.  Per K = 16 step an SM
// reads 4 KB of A + Cout * 16 bytes of B instead of Cout * 32 (Cout = 64: 160 instead of 192 B/clk against the 128 B/clk
// shared-memory port), parked weights take half the shared memory, and streamed weights cross L2 -> SM once per TWO bands:
// 128 -> 128 channels at 40x40 moves 0.97 KB per output pixel instead of the 3.4 KB of the TMA-im2col pair kernel.
// Barrier protocol as in conv_tc.cu: loads of both CTAs complete on the leader's full barrier, commits are multicast,
// epilogue warps of both CTAs arrive on the leader's accumulator-empty barrier.
#include <cstdio>

#include "tc_common.cuh"

namespace fce {
using namespace tc;
namespace {

constexpr int NUM_EPI_WARPS = 12;  // three groups x four TMEM lane quarters (the epilogue, not HBM, paces these layers)
constexpr int NUM_GROUPS = NUM_EPI_WARPS / 4;
constexpr int WARP_PROD_A = 12, WARP_PROD_B = 13, WARP_MMA = 14, WARP_ALLOC = 15;
constexpr int NUM_THREADS = 16 * 32;
constexpr int A_STAGES = 2;
constexpr int SMEM_LIMIT = 222 * 1024;

struct HaloParams {
    int B, H, W, Wp;  // Wp = Wt + 2: padded width of one strip
    int Wt, wtiles;   // column tiles of Wt output columns (maps wider than 254 columns: the TMA box holds at most 256)
    int Cin, Cout;
    int R, bands, nb;  // rows per band, bands per image, 128-row MMA blocks per band
    int chunks;        // K chunks (kc channels each)
    int units;         // B * bands
    int acc_sets;      // 1 or 2 accumulator sets of nb * Cout TMEM columns
    uint32_t strip_bytes, strip_tx;  // shared-memory bytes of one strip stage / bytes one TMA box delivers
    uint32_t b_sub;                  // bytes of one (tap, chunk) weight tile PER CTA: Cout (pair: Cout / 2) x kc bf16
    int pair_units;                  // ceil(units / 2): work items of a CTA pair
    int issuers;                     // 1 or 2 MMA-issuing warps (2: 128-row blocks alternate between them)
    uint32_t b_total, bias_bytes;
    int kc;        // channels per K chunk
    int b_stream;  // weights too large to park: the 9 tap tiles of chunk c travel with strip chunk c (same ring stage)
    uint32_t slab_bytes;  // one staged output slab: R*W rows x slab_cols*2 bytes (1024-aligned)
    int n_slabs, slab_cols, n_stg, has_res;  // slab_cols: 64 (128-byte rows, 128B swizzle) or 32 (64-byte rows, 64B swizzle)
    uint32_t tmem_cols;
    int act;
    uint32_t desc_hi, idesc;
};

// debug cycle accounting, same slot layout as conv_tc.cu: [4] MMA wait-full [5] MMA wait-tmem [6] MMA total
// [7] epilogue wait-tmem-full [8] epilogue total [9] epilogue wait-staging [10] DMA wait-written [11] DMA total
constexpr int PROF_SLOTS = 16;
#ifdef FCE_DEBUG
constexpr bool PROF = true;
__device__ long long g_hprof[kNumSMs * PROF_SLOTS];
#else
constexpr bool PROF = false;
__device__ long long g_hprof[1];
#endif
#define HP_T0() long long _t0 = 0; if (PROF) _t0 = clock64()
#define HP_ACC(var) if (PROF) (var) += clock64() - _t0

template <int KK, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const HaloParams p,
                 const float* __restrict__ bias) {
    constexpr int kc = KK * 16;
    constexpr uint32_t row_b = kc * 2;  // bytes of one strip row (= swizzle span)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base;
    const uint32_t sB = sA + A_STAGES * p.strip_bytes;
    const uint32_t sStg = sB + p.b_total;  // output staging: n_stg x n_slabs x [R*W rows][slab row], swizzled
    const uint32_t stg_bytes = p.n_slabs * p.slab_bytes;
    const uint32_t sBias = sStg + p.n_stg * stg_bytes;
    const uint32_t bars = sBias + p.bias_bytes;
    const uint32_t full0 = bars, empty0 = bars + 16, tfull0 = bars + 32, tempty0 = bars + 48, bfull = bars + 64;
    const uint32_t stg_ready0 = bars + 72, stg_written0 = bars + 88;
    const uint32_t tmem_slot = bars + 104;
    float* bias_s = reinterpret_cast<float*>(smem_raw + (sBias - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Work items: single CTA b walks units b, b + grid, ...; a pair q = blockIdx.x / 2 walks pair-units q, q + grid / 2, ...
    // and its CTA of rank r owns unit 2 * item + r (an odd unit count leaves the last pair's rank 1 with a dummy: it
    // re-reads the last unit's strip for the joint MMA and stores nothing).
    constexpr int NCTA = PAIR ? 2 : 1;
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int item_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int n_items = PAIR ? p.pair_units : p.units;
    const uint32_t full0_sig = PAIR ? mapa_shared(full0, 0) : full0;      // barriers of the LEADER that the peer signals
    const uint32_t bfull_sig = PAIR ? mapa_shared(bfull, 0) : bfull;
    const uint32_t tempty0_sig = PAIR ? mapa_shared(tempty0, 0) : tempty0;

    pdl_launch_dependents();  // the next kernel's prologue may overlap this kernel's tail
    if (warp == WARP_PROD_A && lane == 0) {
        for (int i = 0; i < A_STAGES; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, p.issuers);  // every issuing warp commits once per stage
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8 * a, p.issuers);
            mbar_init(tempty0 + 8 * a, NUM_EPI_WARPS * NCTA);
        }
        mbar_init(bfull, 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(stg_ready0 + 8 * a, 1);
            mbar_init(stg_written0 + 8 * a, NUM_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        if (p.has_res) tma_prefetch_desc(&tmR);
    }
    if (warp == WARP_ALLOC) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    for (int i = threadIdx.x; i < (int)(p.bias_bytes >> 2); i += NUM_THREADS) bias_s[i] = i < p.Cout ? bias[i] * epi_bias_scale(p.act) : 0.f;  // pre-scaled for epi_math16
    tc_fence_before();
    if (PAIR) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int units = p.units, chunks = p.chunks, nb = p.nb, bn = p.Cout;

    if (warp == WARP_PROD_A) {
        // ------------------------------------------------------------------ strip producer
        Ring r;
        const int nrow0 = (int)cta_rank * (p.Cout / NCTA);  // pair: this CTA stages rows [rank * Cout/2, +Cout/2) of each tile
        pdl_wait();  // activations come from the previous kernel (the weights loaded by the B producer do not)
        for (int it = item0; it < n_items; it += item_step) {
            int u = it * NCTA + (int)cta_rank;
            if (u >= units) u = units - 1;  // the dummy half of the last pair
            const int ub = p.wtiles == 1 ? u : u / p.wtiles, wt = u - ub * p.wtiles;  // one tile: no division
            const int b = ub / p.bands, band = ub - b * p.bands;
            const int h0 = band * p.R, w0 = wt * p.Wt;
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(empty0 + 8 * r.stage, r.phase ^ 1);
                if (elect_one()) {
                    const uint32_t fb = full0 + 8 * r.stage, fsig = full0_sig + 8 * r.stage;
                    if (leader) mbar_expect_tx(fb, (p.b_stream ? p.strip_tx + 9u * p.b_sub : p.strip_tx) * NCTA);
                    if (PAIR) tma_load_4d_cg2(sA + r.stage * p.strip_bytes, &tmA, fsig, c * kc, w0 - 1, h0 - 1, b);
                    else tma_load_4d(sA + r.stage * p.strip_bytes, &tmA, fb, c * kc, w0 - 1, h0 - 1, b);
                    if (p.b_stream) {  // this chunk's nine weight tiles ride in the same stage
                        const uint32_t bdst = sB + r.stage * 9u * p.b_sub;
                        for (int t = 0; t < 9; ++t) {
                            if (PAIR) tma_load_2d_cg2(bdst + t * p.b_sub, &tmB, fsig, t * p.Cin + c * kc, nrow0);
                            else tma_load_2d(bdst + t * p.b_sub, &tmB, fb, t * p.Cin + c * kc, 0);
                        }
                    }
                }
                r.advance(A_STAGES);
            }
        }
    } else if (warp == WARP_PROD_B || warp == WARP_MMA) {
        if (warp == WARP_PROD_B) {
            // -------------------------------------------------------------- weights: parked once
            if (!p.b_stream && elect_one()) {
                const int tiles = 9 * chunks;
                const int nrow0 = (int)cta_rank * (p.Cout / NCTA);
                if (leader) mbar_expect_tx(bfull, (uint32_t)tiles * p.b_sub * NCTA);
                // tile index = chunk * 9 + tap ; K column of the OHWI matrix = tap * Cin + chunk * kc
                for (int c = 0; c < chunks; ++c)
                    for (int t = 0; t < 9; ++t) {
                        if (PAIR) tma_load_2d_cg2(sB + (c * 9 + t) * p.b_sub, &tmB, bfull_sig, t * p.Cin + c * kc, nrow0);
                        else tma_load_2d(sB + (c * 9 + t) * p.b_sub, &tmB, bfull, t * p.Cin + c * kc, 0);
                    }
            }
            __syncwarp();
        }
        // ------------------------------------------------------------------ MMA issuers (pair: the leader CTA's only)
        // p.issuers == 2 (experiment, off by default - see the host code): the weight-producer warp, idle after its one
        // load, issues the odd 128-row blocks and this warp the even ones; each block's accumulation stays in one thread's
        // program order, each issuer commits its own MMAs (the stage-empty / accumulator-full barriers count both).
        const int me = warp == WARP_MMA ? 0 : 1;
        if (leader && me < p.issuers) {
        Ring r;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t dhi = p.desc_hi, idesc = p.idesc, b_sub16 = p.b_sub >> 4;
        const int blk_step = p.issuers;
        // strip-row offset of every tap, in 16-byte units
        uint32_t tap16[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) tap16[t] = (uint32_t)((t / 3) * p.Wp + (t % 3)) * (row_b >> 4);
        if (!p.b_stream) {
            mbar_wait(bfull, 0);
            tc_fence_after();
        }
        long long wf = 0, we = 0, mt0 = PROF ? clock64() : 0;
        for (int it = item0; it < n_items; it += item_step) {
            {
                HP_T0();
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                HP_ACC(we);
            }
            tc_fence_after();
            const uint32_t d0 = tmem_base + acc * nb * bn;
            for (int c = 0; c < chunks; ++c) {
                {
                    HP_T0();
                    mbar_wait(full0 + 8 * r.stage, r.phase);
                    HP_ACC(wf);
                }
                tc_fence_after();
                const uint32_t a16 = (sA + r.stage * p.strip_bytes) >> 4;
                const uint32_t b16 = (sB + (p.b_stream ? r.stage : c) * 9 * p.b_sub) >> 4;
                if (elect_one()) {
#pragma unroll 1
                    for (int blk = me; blk < nb; blk += blk_step) {
                        const uint32_t d_tmem = d0 + blk * bn;
                        const uint32_t blk16 = a16 + (uint32_t)blk * (128u * (row_b >> 4));
#pragma unroll
                        for (int t = 0; t < 9; ++t) {
                            // Start address >> 4 of this tap's A operand: the strip shifted by whole pixel rows.  The
                            // operand fetch applies the swizzle XOR to the final shared-memory address (as the TMA
                            // write did), so a start that is only row-aligned needs no descriptor base offset
                            // (verified on B200: setting the base-offset field from address bits 7..9 breaks it).
                            const uint32_t as = blk16 + tap16[t];
                            const uint32_t hi = dhi;
                            const uint32_t bs = b16 + t * b_sub16;
#pragma unroll
                            for (int k = 0; k < KK; ++k) {
                                const uint64_t ad = make_desc(hi, ((as + 2 * k) & 0x3FFF) | (1u << 16));
                                const uint64_t bd = make_desc(dhi, ((bs + 2 * k) & 0x3FFF) | (1u << 16));
                                if (PAIR) umma_bf16_cg2(d_tmem, ad, bd, idesc, (c | t | k) != 0);
                                else umma_bf16(d_tmem, ad, bd, idesc, (c | t | k) != 0);
                            }
                        }
                    }
                    if (PAIR) {
                        umma_commit_cg2(empty0 + 8 * r.stage);
                        if (c == chunks - 1) umma_commit_cg2(tfull0 + 8 * acc);
                    } else {
                        umma_commit(empty0 + 8 * r.stage);
                        if (c == chunks - 1) umma_commit(tfull0 + 8 * acc);
                    }
                }
                __syncwarp();
                r.advance(A_STAGES);
            }
            if (p.acc_sets == 2) {
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            } else {
                acc_phase ^= 1;
            }
        }
        if (PROF && lane == 0 && me == 0) {
            g_hprof[blockIdx.x * PROF_SLOTS + 4] = wf;
            g_hprof[blockIdx.x * PROF_SLOTS + 5] = we;
            g_hprof[blockIdx.x * PROF_SLOTS + 6] = clock64() - mt0;
        }
        }
    } else if (warp == WARP_ALLOC) {
        // ------------------------------------------------------------------ staging DMA: residual in, result out
        // Per band i (staging buffer i % n_stg): the residual tile arrives by TMA (or the buffer is simply declared
        // free) -> stg_ready; the eight epilogue warps fill it -> stg_written; one bulk store per slab.  With two
        // staging buffers the residual of band i+1 is fetched while band i is still in the epilogue.
        const int n_stg = p.n_stg, n_slabs = p.n_slabs, slab_cols = p.slab_cols;
        const uint32_t res_tx = (uint32_t)n_slabs * (uint32_t)(slab_cols * 2) * (uint32_t)p.Wt * (uint32_t)p.R;
        auto make_ready = [&](int u, int buf) {  // executed by one elected lane
            if (p.has_res) {
                const int ub = p.wtiles == 1 ? u : u / p.wtiles, wt = u - ub * p.wtiles;  // one tile: no division
                const int b = ub / p.bands, band = ub - b * p.bands;
                mbar_expect_tx(stg_ready0 + 8 * buf, res_tx);
                for (int sl = 0; sl < n_slabs; ++sl)
                    tma_load_4d(sStg + buf * stg_bytes + sl * p.slab_bytes, &tmR, stg_ready0 + 8 * buf, sl * slab_cols,
                                wt * p.Wt, band * p.R, b);
            } else {
                mbar_arrive(stg_ready0 + 8 * buf);
            }
        };
        uint32_t ph[2] = {0, 0};
        int i = 0;
        long long dw = 0, dt0 = PROF ? clock64() : 0;
        // this CTA's unit of work item `it`; -1 = the dummy half of the last pair (no residual, no store)
        auto unit_of = [&](int it) { const int u = it * NCTA + (int)cta_rank; return u < units ? u : -1; };
        auto ready_or_free = [&](int it, int buf) {  // executed by one elected lane
            const int u = unit_of(it);
            if (u >= 0) make_ready(u, buf);
            else mbar_arrive(stg_ready0 + 8 * buf);
        };
        pdl_wait();  // residual reads and output stores touch buffers the previous kernel may still be using
        if (item0 < n_items && elect_one()) ready_or_free(item0, 0);
        for (int it = item0; it < n_items; it += item_step, ++i) {
            const int u = unit_of(it);
            const int buf = n_stg == 2 ? (i & 1) : 0;
            if (n_stg == 2) {
                // the other buffer's last store (band i-1) must have drained it before band i+1 may load into it
                const int itn = it + item_step;
                if (elect_one()) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    if (itn < n_items) ready_or_free(itn, buf ^ 1);
                }
            }
            {
                HP_T0();
                mbar_wait(stg_written0 + 8 * buf, ph[buf]);
                HP_ACC(dw);
            }
            ph[buf] ^= 1;
            if (elect_one()) {
                if (u >= 0) {
                    const int ub = p.wtiles == 1 ? u : u / p.wtiles, wt = u - ub * p.wtiles;  // one tile: no division
                    const int b = ub / p.bands, band = ub - b * p.bands;
                    for (int sl = 0; sl < n_slabs; ++sl)  // columns past W (ragged last tile) are clipped by the TMA unit
                        tma_store_4d(&tmC, sStg + buf * stg_bytes + sl * p.slab_bytes, sl * slab_cols, wt * p.Wt, band * p.R, b);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (n_stg == 1) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    const int itn = it + item_step;
                    if (itn < n_items) ready_or_free(itn, 0);
                }
            }
            __syncwarp();
        }
        if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
        if (PROF && lane == 0) {
            g_hprof[blockIdx.x * PROF_SLOTS + 10] = dw;
            g_hprof[blockIdx.x * PROF_SLOTS + 11] = clock64() - dt0;
        }
    } else if (warp < NUM_EPI_WARPS) {
        // ------------------------------------------------------------------ epilogue
        // TMEM -> registers -> bias / activation (+ residual, read from the staged tile) -> staging, compacted from the
        // padded-flat accumulator rows to [R][W] so that ONE 4-D bulk store per 32-column slab writes the band.
        const int quarter = warp & 3, grp = warp >> 2;
        const int n_chunks = bn >> 4, act = p.act, Wp = p.Wp, W = p.W, Wt = p.Wt, R = p.R, H = p.H;
        const int pairs = (n_chunks + 1) >> 1;  // work item = (128-row block, pair of 16-column chunks), dealt round-robin
        const int items = nb * pairs;           // to the three warp groups
        const bool has_res = p.has_res != 0;
        const bool wide = p.slab_cols == 64;    // 128-byte staging rows (128B swizzle) vs 64-byte rows (64B swizzle)
        const uint32_t row_bytes = wide ? 128u : 64u, swz_mask = wide ? 7u : 3u;
        int acc = 0, i = 0;
        uint32_t acc_phase = 0, sph[2] = {0, 0};
        long long ew = 0, es = 0, et0 = PROF ? clock64() : 0;
        for (int it = item0; it < n_items; it += item_step, ++i) {
            const int u_raw = it * NCTA + (int)cta_rank;
            const bool live = u_raw < units;  // false: the dummy half of the last pair
            const int u = live ? u_raw : units - 1;
            const int ub = p.wtiles == 1 ? u : u / p.wtiles;  // one column tile (maps up to 254 wide): no division
            const int band = ub % p.bands;
            const int h0 = band * R, w0 = (u - ub * p.wtiles) * Wt;
            const int wlim = min(Wt, W - w0), rlim = min(R, H - h0);  // valid columns / rows of this unit
            const int buf = p.n_stg == 2 ? (i & 1) : 0;
            const uint32_t stg = sStg + buf * stg_bytes;
            {
                HP_T0();
                mbar_wait(tfull0 + 8 * acc, acc_phase);
                HP_ACC(ew);
            }
            {
                HP_T0();
                mbar_wait(stg_ready0 + 8 * buf, sph[buf]);
                HP_ACC(es);
            }
            sph[buf] ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int item = grp; item < (live ? items : 0); item += NUM_GROUPS) {
                const int blk = item / pairs;
                const int j = (item - blk * pairs) * 2;
                const int o = blk * 128 + quarter * 32 + lane;  // padded-flat output index inside the band
                const int ro = o / Wp, co = o - ro * Wp;
                const bool ok = co < wlim && ro < rlim;
                const uint32_t lin = (uint32_t)(ro * Wt + co) * row_bytes;  // byte offset of this pixel's row in a slab
                const uint32_t swz = ((lin >> 7) & swz_mask) << 4;           // swizzle: address bits 7.. -> bits 4..
                const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (acc * nb + blk) * bn;
                {
                    const int n = j * 16;
                    const bool two = j + 1 < n_chunks;
                    // 16-column chunk j lives in slab j / (slab_cols / 16), at byte (j % ...) * 32 of the row
                    const uint32_t rowp = stg + (uint32_t)(wide ? (j >> 2) : (j >> 1)) * p.slab_bytes + lin;
                    const uint32_t cb = wide ? (uint32_t)(j & 2) * 32u : 0u;  // j is even: chunk pair offset 0 or 64
                    uint32_t v0[16], v1[16];
                    tmem_ld16(t_row + n, v0);
                    if (two) tmem_ld16(t_row + n + 16, v1);
                    uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, r2 = r0, r3 = r0;
                    if (has_res && ok) {
                        r0 = ld_shared_v4(rowp + ((cb + 0u) ^ swz));
                        r1 = ld_shared_v4(rowp + ((cb + 16u) ^ swz));
                        if (two) {
                            r2 = ld_shared_v4(rowp + ((cb + 32u) ^ swz));
                            r3 = ld_shared_v4(rowp + ((cb + 48u) ^ swz));
                        }
                    }
                    tmem_ld_wait();
                    float f[16];
                    uint32_t ov[8];
                    epi_math16(v0, bias_s + n, act, has_res, r0, r1, f);
                    pack16(f, ov);
                    if (ok) {
                        st_shared_v4(rowp + ((cb + 0u) ^ swz), ov[0], ov[1], ov[2], ov[3]);
                        st_shared_v4(rowp + ((cb + 16u) ^ swz), ov[4], ov[5], ov[6], ov[7]);
                    }
                    if (two) {
                        epi_math16(v1, bias_s + n + 16, act, has_res, r2, r3, f);
                        pack16(f, ov);
                        if (ok) {
                            st_shared_v4(rowp + ((cb + 32u) ^ swz), ov[0], ov[1], ov[2], ov[3]);
                            st_shared_v4(rowp + ((cb + 48u) ^ swz), ov[4], ov[5], ov[6], ov[7]);
                        }
                    }
                }
            }
            tc_fence_before();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(tempty0_sig + 8 * acc);
                else mbar_arrive(tempty0 + 8 * acc);
                mbar_arrive(stg_written0 + 8 * buf);
            }
            if (p.acc_sets == 2) {
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            } else {
                acc_phase ^= 1;
            }
        }
        if (PROF && warp == 0 && lane == 0) {
            g_hprof[blockIdx.x * PROF_SLOTS + 7] = ew;
            g_hprof[blockIdx.x * PROF_SLOTS + 8] = clock64() - et0;
            g_hprof[blockIdx.x * PROF_SLOTS + 9] = es;
        }
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    if (warp == WARP_ALLOC) {
        if (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

bool g_halo_prof = false, g_halo_last = false;
// Kernel-selection knob (same results either way): FCE_HALO_MODE = 0 strip kernel off (every 3x3 through the TMA-im2col
// kernel), 1 on (default), 2 on with resident weights only.  Read once.
const int g_halo_mode = [] {
    const char* e = getenv("FCE_HALO_MODE");
    return e && *e ? atoi(e) : 1;
}();

}  // namespace

bool conv_halo_ran_last() { return g_halo_last; }
void conv_halo_clear_last() { g_halo_last = false; }
#ifdef FCE_DEBUG
void conv_halo_set_profile(bool on) { g_halo_prof = on; }
int conv_halo_profile(long long* out, int n) {
    if (n > kNumSMs * PROF_SLOTS) n = kNumSMs * PROF_SLOTS;
    cudaError_t e = cudaMemcpyFromSymbol(out, g_hprof, (size_t)n * sizeof(long long));
    if (e != cudaSuccess) {
        set_cuda_error(e);
        return FCE_ERR_CUDA;
    }
    return n;
}
#endif

static const int g_halo_wtiles_max = [] {  // extra column tiles tried when the plain plan does not fit (FCE_HALO_WTILES=1: off)
    const char* e = getenv("FCE_HALO_WTILES");
    return e && *e ? atoi(e) : 4;
}();

// Band height for a given K chunk / weight mode / column-tile count; returns the efficiency estimate (0 = does not fit).
static double halo_plan_tiles(const fce_conv_desc* d, bool has_res, int kc, bool stream, int ncta, int wtiles, HaloParams& p) {
    const int chunks = d->Cin / kc;
    const uint32_t row_b = kc * 2;
    const uint32_t b_sub = (uint32_t)(d->Cout / ncta) * row_b;  // a pair stages half of every weight tile's rows per CTA
    const uint32_t b_total = ((stream ? (uint32_t)A_STAGES * 9u : 9u * chunks) * b_sub + 1023u) & ~1023u;
    const uint32_t bias_bytes = ((uint32_t)d->Cout * 4 + 1023u) & ~1023u;
    const int Wt = (d->W + wtiles - 1) / wtiles;
    const int Wp = Wt + 2;
    int best_R = 0;
    double best_eff = 0.0;
    for (int R = 1; R <= d->H && R + 2 <= 256; ++R) {
        const int nb = (R * Wp + 127) / 128;
        if (2 * nb * d->Cout > 512) break;  // two accumulator sets: the epilogue of a band overlaps the next band's MMAs
        const uint32_t rows = 128u * nb + 2u * Wp + 2u;
        const uint32_t strip = (rows * row_b + 1023u) & ~1023u;
        const uint32_t slab_cols = d->Cout % 64 == 0 ? 64u : 32u;
        const uint32_t slab = ((uint32_t)R * Wt * slab_cols * 2u + 1023u) & ~1023u;
        const uint32_t n_slabs = ((uint32_t)d->Cout + slab_cols - 1) / slab_cols;
        // with a residual the staging is double buffered (the next band's residual streams in during the epilogue)
        const uint32_t n_stg = has_res ? 2u : 1u;
        if ((size_t)A_STAGES * strip + b_total + n_stg * n_slabs * slab + bias_bytes + 2048 > (size_t)SMEM_LIMIT) break;
        // useful fraction of the issued MMA rows, discounted by the halo re-read
        const int bands = (d->H + R - 1) / R;
        const double eff = (double)d->H * d->W / ((double)bands * wtiles * nb * 128) * (0.75 + 0.25 * R / (R + 2.0)) *
                           (wtiles == 1 ? 1.0 : 0.75 + 0.25 * Wt / (Wt + 2.0));
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best_R = R;
        }
    }
    if (best_R == 0) return 0.0;
    p.B = d->B; p.H = d->H; p.W = d->W; p.Wp = Wp;
    p.Wt = Wt; p.wtiles = wtiles;
    p.Cin = d->Cin; p.Cout = d->Cout;
    p.kc = kc;
    p.b_stream = stream ? 1 : 0;
    p.R = best_R;
    p.bands = (d->H + best_R - 1) / best_R;
    p.nb = (best_R * Wp + 127) / 128;
    p.chunks = chunks;
    p.units = d->B * p.bands * wtiles;
    p.pair_units = (p.units + 1) / 2;
    p.acc_sets = 2 * p.nb * d->Cout <= 512 ? 2 : 1;
    p.strip_bytes = ((128u * p.nb + 2u * Wp + 2u) * row_b + 1023u) & ~1023u;
    p.strip_tx = row_b * (uint32_t)Wp * (uint32_t)(best_R + 2);
    p.b_sub = b_sub;
    p.b_total = b_total;
    p.bias_bytes = bias_bytes;
    p.slab_cols = d->Cout % 64 == 0 ? 64 : 32;
    p.slab_bytes = ((uint32_t)best_R * Wt * (uint32_t)p.slab_cols * 2u + 1023u) & ~1023u;
    p.n_slabs = (d->Cout + p.slab_cols - 1) / p.slab_cols;
    // a second staging buffer (residual prefetch / store overlap) when shared memory allows
    p.n_stg = ((size_t)A_STAGES * p.strip_bytes + b_total + 2ull * p.n_slabs * p.slab_bytes + bias_bytes + 2048 <=
               (size_t)SMEM_LIMIT) ? 2 : 1;
    p.tmem_cols = 32;
    while (p.tmem_cols < (uint32_t)(p.acc_sets * p.nb * d->Cout)) p.tmem_cols <<= 1;
    return best_eff;
}

// Column tiles: a TMA box dimension holds at most 256 elements, so maps wider than 254 columns (320 x 320 maps of 1280^2
// inputs) are cut into the fewest equal tiles; every tile is its own unit with its own halo columns.  More tiles than
// that only where the plain plan does not fit the accumulator: 96 outputs leave room for two 128-row blocks per
// accumulator set - one 160-wide row wastes 38 % of them, three rows of an 80-wide tile 6 %.
static double halo_plan_one(const fce_conv_desc* d, bool has_res, int kc, bool stream, int ncta, HaloParams& p) {
    const int wmin = (d->W + 253) / 254;
    double best = halo_plan_tiles(d, has_res, kc, stream, ncta, wmin, p);
    if (best >= 0.55 || g_halo_wtiles_max <= 1) return best;
    for (int wt = wmin + 1; wt <= wmin + g_halo_wtiles_max - 1 && d->W / wt >= 32; ++wt) {
        HaloParams q{};
        const double e = halo_plan_tiles(d, has_res, kc, stream, ncta, wt, q);
        if (e > best + 1e-9) {
            best = e;
            p = q;
        }
    }
    return best;
}

static const int g_stream_max_cout = [] {  // widest output the streamed-weight mode takes (FCE_STREAM_MAX_COUT to vary)
    const char* e = getenv("FCE_STREAM_MAX_COUT");
    return e ? atoi(e) : 128;
}();

// Automatic pair rule, from per-layer A/B timings on B200 (tools/plan_conv_ab.py, profiles/r02_conv_ab_*_v3.csv; m scale,
// batch 256): pairs pay where the operand feed limits the single CTA - 64 -> 64 at 80x80: 193 -> 173 us; Detect box branch
// 256 -> 64 at 80x80 (streamed weights): 546 -> 405 us, 512 -> 64 at 40x40: 316 -> 235 us; 128 -> 128 at 40x40: 136 us against
// 140 us of the im2col pair kernel (155 vs 169 us with the residual).  Thin layers (Cin or Cout of 32 and less) are HBM /
// epilogue bound and LOSE to the coupling of the two CTAs' epilogues (32 -> 32 at 160x160 with residual: 267 -> 351 us).
static bool halo_pair_pays(const fce_conv_desc* d, const HaloParams& q) {
    (void)q;
    return d->Cout >= 64 && d->Cin >= 64;
}

// Picks the weight mode, the K chunk and the band height; returns false when the shape does not fit this kernel.
// ncta = 2 plans the CTA-pair variant (half of each weight tile per CTA: more shapes park their weights).
static bool halo_plan(const fce_conv_desc* d, bool has_res, int ncta, HaloParams& p) {
    if (g_halo_mode == 0) return false;
    if (d->k != 3 || d->stride != 1) return false;
    if (d->Cout > 256 || d->W > 2032 || d->out_dtype != FCE_BF16) return false;
    if (ncta == 2 && (d->Cout % 16 || d->Cout < 32)) return false;  // M = 256 MMAs: N % 16, whole swizzle atoms per CTA
    const int kc0 = d->Cin % 64 == 0 ? 64 : (d->Cin % 32 == 0 ? 32 : 16);
    if (halo_plan_one(d, has_res, kc0, false, ncta, p) >= 0.55) return true;  // weights parked in shared memory
    if (g_halo_mode == 2) return false;                                      // debug: streaming mode off
    // weights streamed with the strip.  Single CTA: worth it for narrow outputs (measured: Cout = 64: 117 -> 82 us, Cout =
    // 128 on a 20x20 map: 26.6 -> 20.5 us), where the TMA-im2col kernel's nine reads of the input per output pixel
    // dominate; 65..128 outputs only on small maps of small batches - from ~600 128-row tiles on, the CTA-pair im2col
    // kernel is faster (m scale, batch 256, 128->128 at 40x40: 146 us streamed strips vs 122 us pairs).  A PAIR streams
    // half of the weights per CTA and band: up to 128 outputs at any size.
    if (d->Cout > g_stream_max_cout) return false;
    if (ncta == 1 && d->Cout > 64 && (d->H * d->W > 1600 || (long long)d->B * d->H * d->W >= 75000)) return false;
    HaloParams best{};
    double best_eff = 0.0;
    for (int kc = 64; kc >= 32; kc >>= 1) {
        if (d->Cin % kc) continue;
        HaloParams q{};
        const double e = halo_plan_one(d, has_res, kc, true, ncta, q);
        if (e > best_eff + 0.05) {  // prefer the wider chunk unless the narrower one buys a clearly better band
            best_eff = e;
            best = q;
        }
    }
    if (best_eff < 0.55) return false;
    p = best;
    return true;
}

// 0 = not a strip-kernel shape, 1 = single CTA, 2 = CTA pair.  Automatic choice (mode 0) from per-layer A/B timings
// (tools/plan_conv_ab.py); mode 1 / 2 force single / pair.
static int halo_choice(const fce_conv_desc* d, bool has_res, int mode, HaloParams& p) {
    static const int pair_env = [] { const char* e = getenv("FCE_HALO_PAIR"); return e && *e ? atoi(e) : -1; }();
    if (mode == 1) return halo_plan(d, has_res, 1, p) ? 1 : 0;
    if (mode == 2) return (halo_plan(d, has_res, 2, p) && p.units >= 2) ? 2 : 0;
    // 128 outputs and more WITH a residual on a large problem: the im2col CTA-pair kernel (conv_tc.cu), whose epilogue now
    // prefetches the residual with 32-byte loads, is ahead of the streamed-weight pair strips (m scale, batch 256,
    // 128 -> 128 at 40x40 + res: 153 us against 163 us; equal at 20x20)
    if (has_res && d->Cout >= 128 && (long long)d->B * d->H * d->W >= 200000) return 0;
    if (pair_env != 0) {
        HaloParams q{};
        if (halo_plan(d, has_res, 2, q) && q.units >= kNumSMs / 2 && (pair_env == 1 || halo_pair_pays(d, q))) {
            p = q;
            return 2;
        }
    }
    return halo_plan(d, has_res, 1, p) ? 1 : 0;
}

int conv2d_halo_choice(const fce_conv_desc* d, bool has_res, int mode) {
    HaloParams p{};
    return halo_choice(d, has_res, mode, p);
}

int conv2d_halo(const fce_conv_desc* d, const void* x, const void* w, const float* bias, const void* res, void* y,
                cudaStream_t st, int mode) {
    const DriverApi& api = driver();
    HaloParams p{};
    const int choice = api.ok ? halo_choice(d, res != nullptr, mode, p) : 0;
    if (!choice) return FCE_ERR_UNSUPPORTED;
    const int ncta = choice;
    const int kc = p.kc;
    const uint32_t row_b = kc * 2;
    p.act = d->act;
    p.has_res = res != nullptr;
    // One issuing warp.  FCE_HALO_ISSUERS=2 lets the (idle) weight-producer warp issue every other 128-row block - an
    // experiment that did NOT pay (m scale, batch 256, same box: 64->64 at 80x80 157 -> 177 us, 128->128 at 40x40 138 -> 146 us,
    // with residual 159 -> 142 us; step total 20.61 -> 20.86 ms): the short-MMA layers are paced by the tensor core's own
    // dispatch rate, not by the issuing thread's instruction stream (4-5 uniform-datapath instructions per UTCHMMA).
    static const int issuers_env = [] { const char* e = getenv("FCE_HALO_ISSUERS"); return e && *e ? atoi(e) : 1; }();
    p.issuers = issuers_env >= 2 && p.nb >= 2 ? 2 : 1;
    const uint32_t layout = row_b == 128 ? 2u : (row_b == 64 ? 4u : 6u);
    const uint32_t sbo = 8 * row_b;
    p.desc_hi = (sbo >> 4) | (1u << 14) | (layout << 29);
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(d->Cout >> 3) << 17) | ((uint32_t)((128 * ncta) >> 4) << 24);
    const CUtensorMapSwizzle swz = row_b == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : row_b == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                 : CU_TENSOR_MAP_SWIZZLE_32B;
    alignas(64) CUtensorMap tmA, tmB, tmC, tmR;
    const __nv_bfloat16* xin = reinterpret_cast<const __nv_bfloat16*>(x) + d->in_off;
    // output / residual bands: box {32 channels, W, R rows, 1 image} of the NHWC view, 64B-swizzled staging
    for (int which = 0; which < 2; ++which) {
        const bool is_res = which == 1;
        if (is_res && !res) {
            tmR = tmC;
            break;
        }
        const int pitch = is_res ? d->res_pitch : d->out_pitch;
        void* ptr = is_res ? (void*)(reinterpret_cast<const __nv_bfloat16*>(res) + d->res_off)
                           : (void*)(reinterpret_cast<__nv_bfloat16*>(y) + d->out_off);
        const cuuint64_t gdim[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
        const cuuint64_t gstr[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)d->W * pitch * 2, (cuuint64_t)d->H * d->W * pitch * 2};
        const cuuint32_t box[4] = {(cuuint32_t)p.slab_cols, (cuuint32_t)p.Wt, (cuuint32_t)p.R, 1};
        const cuuint32_t est[4] = {1, 1, 1, 1};
        if (api.tiled(is_res ? &tmR : &tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, gdim, gstr, box, est,
                      CU_TENSOR_MAP_INTERLEAVE_NONE,
                      p.slab_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCE_ERR_UNSUPPORTED;
    }
    {
        const cuuint64_t gdim[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
        const cuuint64_t gstr[3] = {(cuuint64_t)d->in_pitch * 2, (cuuint64_t)d->W * d->in_pitch * 2,
                                    (cuuint64_t)d->H * d->W * d->in_pitch * 2};
        const cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)p.Wp, (cuuint32_t)(p.R + 2), 1};
        const cuuint32_t est[4] = {1, 1, 1, 1};
        if (api.tiled(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)xin, gdim, gstr, box, est,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCE_ERR_UNSUPPORTED;
    }
    {
        const cuuint64_t K = 9ull * d->Cin;
        const cuuint64_t gdim[2] = {K, (cuuint64_t)d->Cout};
        const cuuint64_t gstr[1] = {K * 2};
        const cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)(d->Cout / ncta)};
        const cuuint32_t est[2] = {1, 1};
        if (api.tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, est,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCE_ERR_UNSUPPORTED;
    }
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const HaloParams,
                             const float*);
    static const KernelFn table[6] = {conv_halo_kernel<1, false>, conv_halo_kernel<2, false>, conv_halo_kernel<4, false>,
                                      conv_halo_kernel<1, true>,  conv_halo_kernel<2, true>,  conv_halo_kernel<4, true>};
    static DeviceOnce attr_once;  // the shared-memory opt-in is a per-device attribute
    int dev = 0;
    if (attr_once.pending(&dev)) {
        for (int v = 0; v < 6; ++v) {
            cudaError_t e = cudaFuncSetAttribute(table[v], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) {
                set_cuda_error(e);
                return FCE_ERR_CUDA;
            }
        }
        attr_once.done(dev);
    }
    const size_t smem = (size_t)A_STAGES * p.strip_bytes + p.b_total + (size_t)p.n_stg * p.n_slabs * p.slab_bytes +
                        p.bias_bytes + 1024 + 256;
    g_halo_last = true;
    if (g_halo_prof)
        fprintf(stderr, "[halo] ncta=%d R=%d bands=%d nb=%d units=%d acc_sets=%d n_stg=%d slab_cols=%d stream=%d smem=%zu\n", ncta,
                p.R, p.bands, p.nb, p.units, p.acc_sets, p.n_stg, p.slab_cols, p.b_stream, smem);
    const int kv = kc == 16 ? 0 : (kc == 32 ? 1 : 2);
    if (ncta == 2) {
        const int pairs = p.pair_units < kNumSMs / 2 ? p.pair_units : kNumSMs / 2;
        return launch_pdl_cluster(table[3 + kv], 2 * pairs, NUM_THREADS, smem, st, 2, tmA, tmB, tmC, tmR, p, bias);
    }
    const int grid = p.units < kNumSMs ? p.units : kNumSMs;
    return launch_pdl(table[kv], grid, NUM_THREADS, smem, st, tmA, tmB, tmC, tmR, p, bias);
}

}  // namespace fce

// The first TWO convolutions of the graph in one pass: Conv(3, C0, 3, 2) on the uint8 image followed by Conv(C0, C1, 3, 2)
// (yolo11-fce.yaml:20-21; Conv.forward_fuse conv.py:80-89), C0 = 64.
//
// As two launches the stem writes its [B, H/2, W/2, 64] map to HBM and the second conv reads it back: at m scale, batch
// 256, 640^2 that is 2 x 3.36 GB of the step's ~73 GB and 1.04 + 1.06 ms (the stem at 0.84 of its write-aware roofline,
// the second conv HBM-bound at 0.72 of the copy peak).  Here the stem map never exists in HBM.
//
//   unit = 7 x 16 output pixels of the second conv (M = 112 of the 128 rows of one tcgen05.mma).  It needs a 15 x 33
//   window of stem pixels, which needs 31 x 67 image pixels.
//   * the stem window is itself a small GEMM on the tensor cores: four blocks of 128 stem pixels, [128 x 32] x [32 x 64].
//     BUILDER warps (2 teams of 4, one stem pixel per thread) read the pixel's 3 x 10 image bytes (L1-cached words, funnel-shifted to
//     the byte origin), convert them to fp16 (byte / 256, the magic-number trick of stem.cu) and write one 64-byte A row,
//     K order k' = 10 * kh + j; the stem weights sit in shared memory as the fp16 B operand (x 256).  A first version
//     computed the stem with warp-level mma.sync like stem.cu: 19 warp instructions per stem pixel made the fused kernel
//     stem-bound at 2.3 ms; the operand build costs ~3.
//   * STEM EPILOGUE warps (2 groups of 4, blocks alternate; the bias rides in the GEMM as K column 30): TMEM -> SiLU -> bf16 -> the STEM TILE in shared memory, laid out as FOUR PARITY PLANES
//     (stem row parity x column parity), each a padded-flat pixel list of pitch 17 with 128-byte rows swizzled by their
//     absolute shared-memory address.  A stride-2 3x3 tap (kh, kw) of output pixel o = r * 17 + c reads plane
//     (kh & 1, kw & 1) at flat index o + (kh >> 1) * 17 + (kw >> 1): every tap's A operand is ONE shifted shared-memory
//     descriptor over a plane (conv_halo.cu's trick, per parity) - no im2col traffic at all.  The 17th column of every row
//     produces accumulator rows that are never stored.  Stem pixels outside the stem map are written as zeros: they are
//     the zero padding of the second conv.
//   * MMA warp: the planes are produced and consumed in the order P11, P10, P01, P00 (1, 2, 2, 4 taps; the four stem blocks
//     walk the slot list in the same order, so plane q is complete when block q is); each plane has its own full / free
//     barrier, so the stem epilogue refills plane q for the next unit while the tensor core still works on planes q+1.. of
//     this one - one stem tile in shared memory behaves like a four-stage ring.  36 tcgen05.mma (M = 128, N = C1, K = 16)
//     per unit against the PARKED weights (all nine [C1, 64] tap tiles, loaded once per CTA); the stem GEMMs (four TMEM
//     stages) and the plane steps are two independent instruction streams of this warp, each issued when its barriers allow.
//   * epilogue warps (4): TMEM -> bias + SiLU -> bf16 -> 32-byte global stores, two accumulator stages.
//
// HBM bytes per output pixel: 48 image bytes (+ halo, L2) read, 2 * C1 written.  m scale, batch 256: 0.3 + 1.68 GB instead
// of 0.3 + 3.36 + 3.36 + 1.68 GB.
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace fce {
using namespace tc;
namespace {

// C0 (template): stem channels = K chunk of the second conv: 64 (m / l scale: 128-byte rows, 128B swizzle) or 32 (s scale:
// 64-byte rows, 64B swizzle)
constexpr int TH = 7, TW = 16, PW = TW + 1; // unit of the second conv; plane pitch
constexpr int SC = 2 * TW + 1;                      // stem window: 15 rows x 33 columns
// warp roles (21 warps: 80 registers per thread; 25 warps with eight epilogue warps measured slower: the SM is paced by
// its aggregate instruction rate, more warps only dilute it)
constexpr int TEAM_WARPS = 4;                       //     // 0 .. 7: two builder teams (blocks alternate), one stem pixel (A row) per thread
constexpr int WARP_SEPI0 = 8;                       // 8 .. 15 stem epilogue: warp & 3 = TMEM lane quarter, (warp - 8) >> 2 = group
constexpr int WARP_EPI0 = 16, NEPI_WARPS = 4;       // 16 .. 19 epilogue of the second conv
constexpr int WARP_MMA = 20;
constexpr int NUM_THREADS = 21 * 32;
// planes in memory order P00, P01, P10, P11 (index = (row parity) * 2 + column parity)
// (136, 136, 119, 119 pixel slots)
__host__ __device__ constexpr int pl_base(int pl) { return pl == 0 ? 0 : pl == 1 ? 8 * PW : pl == 2 ? 16 * PW : 23 * PW; }  // 0 136 272 391
constexpr int ST_ROWS_ALLOC = 520;  // the last descriptor reads rows [391 + 1, +128)
// Production / consumption order q = 0..3 -> plane P11, P10, P01, P00.  The stem window is computed as FOUR blocks of 128
// slots of the list [P11 | P10 | P01 | P00] (119 + 119 + 136 + 136 = 510 slots): plane q is complete when block q is.
__host__ __device__ constexpr int q_plane(int q) { return 3 - q; }
constexpr int N_SLOTS = 510;
constexpr int STEM_STAGES = 4;           // stem accumulator stages in TMEM (64 columns each): block blk of a unit -> stage blk
constexpr uint32_t A0_BYTES = 128 * 64;  // stem GEMM A operand: 128 rows x K = 32 fp16 (64-byte rows, 64B swizzle)
// stem weights (B operand of the stem GEMM): C0 rows x 32 fp16 = C0 x 64 bytes
constexpr int SMEM_LIMIT = 227 * 1024;

struct Stem2Params {
    int B, H, W;        // image
    int H0, W0;         // stem map
    int H1, W1, C1;     // output
    int tiles_h, tiles_w, units;
    int out_pitch, act0, act1;
    uint32_t w_tile;    // bytes of one tap's weight tile: C1 x 128
    uint32_t bias_bytes, tmem_cols;
    uint32_t desc_hi, idesc;     // second conv: 128-byte rows, bf16
    uint32_t desc_hi0, idesc0;   // stem GEMM: 64-byte rows, fp16
    int wide_store;
};

#ifdef FCE_DEBUG
constexpr bool DBG = true;
// per-CTA cycle accounting: [0] builder wait-A-empty [1] builder total | [2] stem epilogue wait-acc [3] wait-plane-free [4] total
// | [5] MMA wait-A-full [6] wait-stem-acc-empty [7] wait-plane-full [8] wait-acc-empty [9] total | [10] epilogue wait [11] total
__device__ long long g_sprof[kNumSMs * 16];
#else
constexpr bool DBG = false;
__device__ long long g_sprof[1];
#endif
#define SP_T0() long long _t0 = 0; if (DBG) _t0 = clock64()
#define SP_ACC(var) if (DBG) (var) += clock64() - _t0

// non-blocking phase test (try_wait may suspend the thread for a hardware-defined time before it reports failure)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void st2_global_v8(void* p, const uint32_t* o) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                 "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                 : "memory");
}
// global slot of the unit's list -> (plane, slot inside the plane); slots >= N_SLOTS give an index past the plane
__device__ __forceinline__ void slot_to_plane(int gs, int& pl, int& i) {
    if (gs < 119) { pl = 3; i = gs; }
    else if (gs < 238) { pl = 2; i = gs - 119; }
    else if (gs < 374) { pl = 1; i = gs - 238; }
    else { pl = 0; i = gs - 374; }
}

template <int C0>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_stem2_kernel(const __grid_constant__ CUtensorMap tmW, const Stem2Params p, const uint8_t* __restrict__ x,
                  const __nv_bfloat16* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ b1,
                  __nv_bfloat16* __restrict__ y) {
    constexpr uint32_t RB = C0 * 2;              // bytes of one stem-tile / weight row = swizzle span
    constexpr uint32_t SWZ = RB == 128 ? 7u : 3u;  // 16-byte chunk c of a row lives at c ^ ((address >> 7) & SWZ)
    constexpr uint32_t STILE_BYTES = (ST_ROWS_ALLOC * RB + 1023u) & ~1023u, B0_BYTES = C0 * 64;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw0 = smem_u32(smem_raw);
    const uint32_t base = (raw0 + 1023u) & ~1023u;
    const uint32_t sW = base;                           // 9 x [C1][C0] bf16, swizzled
    const uint32_t sT = sW + 9 * p.w_tile;              // stem tile: four parity planes, 128-byte rows
    const uint32_t sA = sT + STILE_BYTES;               // stem GEMM A operand (one stage)
    const uint32_t sB0 = sA + A0_BYTES;                 // stem GEMM B operand
    const uint32_t sBias = sB0 + B0_BYTES;
    const uint32_t bars = sBias + p.bias_bytes;
    const uint32_t a_full = bars, a_empty0 = bars + 8, s_tfull0 = bars + 24, s_tempty0 = bars + 56;
    const uint32_t pfull0 = bars + 88, pfree0 = bars + 120, tfull0 = bars + 152, tempty0 = bars + 168, wfull = bars + 184;
    const uint32_t tmem_slot = bars + 192;
    float* bias_s = reinterpret_cast<float*>(smem_raw + (sBias - raw0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = p.units;

    pdl_launch_dependents();
    if (warp == WARP_MMA && lane == 0) {
        mbar_init(a_full, TEAM_WARPS);
        for (int a = 0; a < 2; ++a) {
            mbar_init(a_empty0 + 8 * a, 1);
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, NEPI_WARPS);
        }
        for (int a = 0; a < STEM_STAGES; ++a) {
            mbar_init(s_tfull0 + 8 * a, 1);
            mbar_init(s_tempty0 + 8 * a, 4);  // the stem-epilogue group that drains the stage
        }
        for (int q = 0; q < 4; ++q) {
            mbar_init(pfull0 + 8 * q, q == 0 ? 4 : 8);  // the group of block q + (q > 0) the group of block q - 1 (plane head)
            mbar_init(pfree0 + 8 * q, 1);
        }
        mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmW);
    }
    if (warp == WARP_EPI0) {  // TMEM allocator (a warp that is idle until the first accumulator is complete)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        // Stem weights as the B operand of the stem GEMM: fp16 [64][32], K order k' = 10 * kh + j (j = 0..9: the nine values of
        // patch row kh and their zero-weight neighbour; k' 30 = the bias, 31 zero), x 256 (x 128 with SiLU) because the A operand holds byte / 256
        // (bf16 -> fp16 and the power of two are exact; the bias keeps 11 bits - three more than the bf16 result); 64-byte
        // rows, 64B swizzle (16-byte chunk c of row n at c ^ ((n >> 1) & 3))
        const unsigned short* wus = reinterpret_cast<const unsigned short*>(w0);
        const float wsc = p.act0 == FCE_ACT_SILU ? 128.f : 256.f;  // SiLU as h + h * tanh(h): the GEMM produces h = v / 2 directly
        for (int idx = threadIdx.x; idx < C0 * 32; idx += NUM_THREADS) {
            const int n = idx >> 5, kp = idx & 31;
            const int kh = kp / 10, j = kp - kh * 10;
            float v = 0.f;
            if (kp < 30 && j < 9) v = __uint_as_float((uint32_t)__ldg(wus + n * 32 + kh * 9 + j) << 16) * wsc;
            if (kp == 30) v = __ldg(b0 + n) * wsc;  // the bias rides in the GEMM: A column 30 is the constant 1 / 256
            const __half hv = __float2half_rn(v);
            const uint32_t addr = sB0 + (uint32_t)n * 64u + ((((uint32_t)kp >> 3) ^ (((uint32_t)n >> 1) & 3u)) << 4) + ((uint32_t)kp & 7u) * 2u;
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const unsigned short*>(&hv)) : "memory");
        }
        const float bsc1 = epi_bias_scale(p.act1);  // pre-scaled for epi_math16
        for (int i = threadIdx.x; i < (int)(p.bias_bytes >> 2); i += NUM_THREADS) bias_s[i] = i < p.C1 ? b1[i] * bsc1 : 0.f;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the B operand is read by the tensor core
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    auto unit_coords = [&](int u, int& b, int& oh0, int& ow0) {
        const int tw = u % p.tiles_w;
        const int r = u / p.tiles_w;
        const int th = r % p.tiles_h;
        b = r / p.tiles_h;
        oh0 = th * TH;
        ow0 = tw * TW;
    };
    // this thread's stem pixel of block blk: plane, slot in the plane, window coordinates, liveness
    auto stem_slot = [&](int blk, int row, int& pl, int& i, int& r, int& c) {
        const int gs = blk * 128 + row;
        slot_to_plane(gs, pl, i);
        const int pi = i / PW, pj = i - pi * PW;
        r = 2 * pi + (pl >> 1);
        c = 2 * pj + (pl & 1);
        return gs < N_SLOTS && c < SC;
    };
    const int my_units = (int)blockIdx.x < units ? (units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == WARP_MMA) {
        // ------------------------------------------------------------------ weights (once), then the MMA issue loop
        if (elect_one()) {
            mbar_expect_tx(wfull, 9u * p.w_tile);
            for (int t = 0; t < 9; ++t) tma_load_2d(sW + t * p.w_tile, &tmW, wfull, t * C0, 0);
        }
        __syncwarp();
        mbar_wait(wfull, 0);
        tc_fence_after();
        const uint32_t dhi = p.desc_hi, idesc = p.idesc, dhi0 = p.desc_hi0, idesc0 = p.idesc0;
        const int S = 4 * my_units;  // stem blocks = plane steps of this CTA
        long long m_af = 0, m_se = 0, m_pf = 0, m_te = 0, m_t0 = DBG ? clock64() : 0;  // (the waits of this warp are polls now)
        // Two instruction streams share this warp: the stem GEMMs (block sb: [128 slots x 32] x [32 x 64] -> stem accumulator
        // stage sb & 3) and the plane steps of the second conv (step lp: the taps of plane order q = lp & 3 of unit lp >> 2).
        // Each is issued as soon as ITS barriers allow (non-blocking tests): a blocking wait for one stream would hold back the
        // other - the first version waited in program order and the stem epilogue sat idle 38 % of the time.
        auto ready = [&](uint32_t bar, uint32_t parity) { return __all_sync(0xffffffffu, mbar_test(bar, parity)) != 0; };
        int sb = 0, lp = 0;
        (void)m_af; (void)m_se; (void)m_pf; (void)m_te;
        long long t_prog = clock64();  // bounded like every other wait: a protocol bug faults the launch instead of hanging
        int idle = 0;
        while (lp < S) {
            if (++idle == 4096) {
                idle = 0;
                if (clock64() - t_prog > 4000000000LL) __trap();
            }
            if (sb < S && sb < lp + STEM_STAGES) {
                const int stage = sb & 3;
                // the stem epilogue has drained the stage (block sb - 4), the builders have filled the A operand
                if (ready(s_tempty0 + 8 * stage, (((uint32_t)sb >> 2) & 1u) ^ 1u) && ready(a_full, (uint32_t)sb & 1u)) {
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a16 = sA >> 4, b16 = sB0 >> 4;
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            umma_bf16(tmem_base + stage * C0, make_desc(dhi0, ((a16 + 2 * k) & 0x3FFF) | (1u << 16)),
                                      make_desc(dhi0, ((b16 + 2 * k) & 0x3FFF) | (1u << 16)), idesc0, (uint32_t)k);
                        umma_commit(a_empty0 + 8 * (sb & 1));  // the OTHER builder team may fill the stage with block sb + 1
                        umma_commit(s_tfull0 + 8 * stage);     // -> stem epilogue
                    }
                    __syncwarp();
                    ++sb;
                    idle = 0;
                    t_prog = clock64();
                }
            }
            {
                const int q = lp & 3, itu = lp >> 2, acc = itu & 1;
                // the plane is complete; for the first plane of a unit the epilogue has drained the accumulator stage
                if ((q != 0 || ready(tempty0 + 8 * acc, (((uint32_t)itu >> 1) & 1u) ^ 1u)) && ready(pfull0 + 8 * q, (uint32_t)itu & 1u)) {
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t d_tmem = tmem_base + STEM_STAGES * C0 + acc * p.C1;
                        const int pl = q_plane(q), pr = pl >> 1, pc = pl & 1;
                        bool first = q == 0;
                        for (int kh = pr; kh < 3; kh += 2)
                            for (int kw = pc; kw < 3; kw += 2) {
                                const uint32_t a16 = (sT + (uint32_t)(pl_base(pl) + (kh >> 1) * PW + (kw >> 1)) * RB) >> 4;
                                const uint32_t b16 = (sW + (uint32_t)(kh * 3 + kw) * p.w_tile) >> 4;
#pragma unroll
                                for (int k = 0; k < C0 / 16; ++k) {
                                    umma_bf16(d_tmem, make_desc(dhi, ((a16 + 2 * k) & 0x3FFF) | (1u << 16)),
                                              make_desc(dhi, ((b16 + 2 * k) & 0x3FFF) | (1u << 16)), idesc, first ? 0u : 1u);
                                    first = false;
                                }
                            }
                        umma_commit(pfree0 + 8 * q);                // the stem epilogue may refill this plane
                        if (q == 3) umma_commit(tfull0 + 8 * acc);  // accumulator complete -> epilogue
                    }
                    __syncwarp();
                    ++lp;
                    idle = 0;
                }
            }
        }
        if (DBG && lane == 0) {
            long long* g = g_sprof + blockIdx.x * 16;
            g[5] = m_af; g[6] = m_se; g[7] = m_pf; g[8] = m_te; g[9] = clock64() - m_t0;
        }
    } else if (warp < WARP_SEPI0) {
        // ------------------------------------------------------------------ builders: image patches -> A rows of the stem GEMM
        // Two teams of four warps; team t builds the blocks sb = t, t + 2, ... of the CTA's block sequence: the loads and
        // conversions of block sb + 1 run while block sb waits for the (single) A stage.
        const int team = warp / TEAM_WARPS;
        const int row = (int)threadIdx.x - team * (TEAM_WARPS * 32);  // 0 .. 127
        const uint32_t* xw = reinterpret_cast<const uint32_t*>(x);
        const int row_words = p.W * 3 / 4;
        const __half2 four = __float2half2_rn(4.f);
        auto cvt2 = [&](uint32_t v, uint32_t sel) {  // two bytes -> two fp16 (byte / 256): 0x4400 | byte = 4 + byte / 256
            const uint32_t m = __byte_perm(v, 0x44444444u, sel);
            const __half2 h = __hsub2(*reinterpret_cast<const __half2*>(&m), four);
            return *reinterpret_cast<const uint32_t*>(&h);
        };
        const uint32_t my_row = sA + (uint32_t)row * 64u, sw = ((uint32_t)row >> 1) & 3u;
        // this thread's two stem pixels (blocks team and team + 2) do not depend on the unit: window coordinates r | c << 8,
        // -1 = no pixel
        int slot_rc[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            int pl, i, r, c;
            slot_rc[j] = stem_slot(2 * j + team, row, pl, i, r, c) ? (r | (c << 8)) : -1;
        }
        // next unit's image window -> L1 / L2 while this unit is computed: 31 rows x (up to) three 128-byte lines
        auto prefetch_window = [&](int u) {
            int b, oh0, ow0;
            unit_coords(u, b, oh0, ow0);
            const int r = row / 3, k = row - 3 * r;  // 93 of the 128 threads: row r, line k
            const int ih = 4 * oh0 - 3 + r;
            const long long bcol = 3LL * (4 * ow0 - 3) + 128 * k;
            if (r < 31 && ih >= 0 && ih < p.H && bcol >= -127 && bcol < 3LL * p.W && (k < 2 || bcol < 3LL * (4 * ow0 + 64))) {
                const uint8_t* a = x + ((size_t)(b * p.H + ih) * p.W) * 3 + (bcol < 0 ? 0 : bcol);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
            }
        };
        long long b_w = 0, b_t0 = DBG ? clock64() : 0;
        // Software pipeline over this team's block sequence (block n: unit blockIdx.x + (n >> 1) * grid, slot j = n & 1): the
        // twelve image words of block n + 1 are requested before block n is converted and stored, so that their latency (L2
        // or HBM: ~2.7 k cycles per block were spent here) hides behind a whole block of work.
        uint32_t wd[3][4];
        uint32_t sh = 0;
        bool inside = false;
        auto request = [&](int n) {  // loads of block n -> wd / sh / inside
            const int u = (int)blockIdx.x + (n >> 1) * (int)gridDim.x;
            inside = false;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int k = 0; k < 4; ++k) wd[kh][k] = 0u;
            if (u >= units) return;
            int b, oh0, ow0;
            unit_coords(u, b, oh0, ow0);
            const int rc = slot_rc[n & 1];
            const int sr = 2 * oh0 - 1 + (rc & 255), scc = 2 * ow0 - 1 + (rc >> 8);  // stem coordinates of this thread's pixel
            inside = rc >= 0 && sr >= 0 && sr < p.H0 && scc >= 0 && scc < p.W0;
            if (!inside) return;
            // patch row kh: image row 2 sr - 1 + kh, bytes [6 sc - 3, +9] (three pixels x 3 channels + one neighbour)
            const int bs = 6 * scc - 3, w_lo = bs >> 2;
            sh = 8u * (uint32_t)(bs & 3);
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int ih = 2 * sr - 1 + kh;
                if (ih < 0 || ih >= p.H) continue;  // zero padding of the stem conv
                const uint32_t* rowp = xw + (size_t)(b * p.H + ih) * row_words;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (w_lo + k >= 0 && w_lo + k < row_words) wd[kh][k] = __ldg(rowp + w_lo + k);
            }
        };
        const int n_blocks = 2 * my_units;
        pdl_wait();  // the image comes from the previous work in the stream
        if (n_blocks > 0) request(0);
#pragma unroll 1
        for (int nb = 0; nb < n_blocks; ++nb) {  // block sb = 2 * nb + team of the CTA's sequence
            if (team == 0 && (nb & 1) == 0) {
                const int un = (int)blockIdx.x + ((nb >> 1) + 1) * (int)gridDim.x;
                if (un < units) prefetch_window(un);
            }
            uint32_t hw[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) hw[k] = 0u;
            hw[15] = 0x00001C00u;  // k' = 30: fp16 1 / 256, the multiplier of the bias row of B
            if (inside) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const uint32_t v0 = __funnelshift_r(wd[kh][0], wd[kh][1], sh), v1 = __funnelshift_r(wd[kh][1], wd[kh][2], sh),
                                   v2 = __funnelshift_r(wd[kh][2], wd[kh][3], sh);
                    hw[5 * kh + 0] = cvt2(v0, 0x4140);
                    hw[5 * kh + 1] = cvt2(v0, 0x4342);
                    hw[5 * kh + 2] = cvt2(v1, 0x4140);
                    hw[5 * kh + 3] = cvt2(v1, 0x4342);
                    hw[5 * kh + 4] = cvt2(v2, 0x4140);
                }
            }
            request(nb + 1);  // in flight across the wait, the stores and the hand-over below
            {
                // the stem GEMM of the previous block (sb - 1, built by the other team) has read the stage: its commits go
                // to a_empty[(sb - 1) & 1] - a barrier whose every phase this team observes (no parity aliasing)
                SP_T0();
                mbar_wait(a_empty0 + 8 * (team ^ 1), team ? ((uint32_t)nb & 1u) : (((uint32_t)nb & 1u) ^ 1u));
                SP_ACC(b_w);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                st_shared_v4(my_row + ((((uint32_t)k) ^ sw) << 4), hw[4 * k], hw[4 * k + 1], hw[4 * k + 2], hw[4 * k + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
        if (DBG && warp == 0 && lane == 0) {
            g_sprof[blockIdx.x * 16 + 0] = b_w;
            g_sprof[blockIdx.x * 16 + 1] = clock64() - b_t0;
        }
    } else if (warp < WARP_EPI0) {
        // ------------------------------------------------------------------ stem epilogue: TMEM -> SiLU -> bf16 -> planes
        // Two groups of four warps (one warp per TMEM lane quarter); group g drains the blocks g and g + 2 of every unit from
        // stem accumulator stage g, all 64 channels of its 32 pixels per warp: two blocks are in flight at a time.
        const int quarter = warp & 3, grp = (warp - WARP_SEPI0) >> 2;
        const int row = quarter * 32 + lane;
        const bool silu = p.act0 == FCE_ACT_SILU;
        // this thread's two stem pixels do not depend on the unit: window coordinates and the row address in the stem tile
        int slot_rc[2];
        uint32_t slot_ra[2];
#pragma unroll
        for (int j2 = 0; j2 < 2; ++j2) {
            int pl, i, r, c;
            const bool live = stem_slot(2 * j2 + grp, row, pl, i, r, c);
            slot_rc[j2] = r | (c << 8);
            slot_ra[j2] = live ? sT + (uint32_t)(pl_base(pl) + i) * RB : 0u;
        }
        int itu = 0;
        long long e_wa = 0, e_wp = 0, e_t0 = DBG ? clock64() : 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x, ++itu) {
            int b, oh0, ow0;
            unit_coords(u, b, oh0, ow0);
            const int sr0 = 2 * oh0 - 1, sc0 = 2 * ow0 - 1;
#pragma unroll
            for (int j2 = 0; j2 < 2; ++j2) {
                const int blk = 2 * j2 + grp;  // block sb = 4 * itu + blk of the CTA's sequence, stem accumulator stage blk
                const uint32_t ra = slot_ra[j2];
                const int sr = sr0 + (slot_rc[j2] & 255), scc = sc0 + (slot_rc[j2] >> 8);
                const bool inside = ra != 0u && sr >= 0 && sr < p.H0 && scc >= 0 && scc < p.W0;
                {
                    SP_T0();
                    mbar_wait(s_tfull0 + 8 * blk, (uint32_t)itu & 1u);  // one completion of stage blk per unit
                    SP_ACC(e_wa);
                }
                {
                    // the MMAs of the previous unit are done with the planes this block writes: plane step blk, head of blk + 1
                    SP_T0();
                    mbar_wait(pfree0 + 8 * blk, ((uint32_t)itu & 1u) ^ 1u);
                    if (blk < 3) mbar_wait(pfree0 + 8 * (blk + 1), ((uint32_t)itu & 1u) ^ 1u);
                    SP_ACC(e_wp);
                }
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + blk * C0;
                const uint32_t xr = (ra >> 7) & SWZ;
                // 16 channels (two 16-byte chunks of the pixel's row) at a time.  (Keeping the TMEM load of group cg + 1 in flight
                // while group cg is computed measured SLOWER: 1.32 -> 1.52 ms.)
#pragma unroll
                for (int cg = 0; cg < C0 / 16; ++cg) {
                    uint32_t v[16], o[8];
                    tmem_ld16(t_row + 16 * cg, v);
                    tmem_ld_wait();
                    if (cg == C0 / 16 - 1) {  // the accumulator is in registers: the stage is free for block sb + 4
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(s_tempty0 + 8 * blk);
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float2 a = make_float2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
                        if (silu) a = __ffma2_rn(a, make_float2(tanh_fast(a.x), tanh_fast(a.y)), a);  // the accumulator is h = v / 2
                        __nv_bfloat162 ob = __floats2bfloat162_rn(a.x, a.y);
                        o[k] = *reinterpret_cast<uint32_t*>(&ob);
                    }
                    if (!inside) {  // outside the stem map: the second conv's zero padding
#pragma unroll
                        for (int k = 0; k < 8; ++k) o[k] = 0u;
                    }
                    if (ra != 0u) {
                        st_shared_v4(ra + (((uint32_t)(2 * cg) ^ xr) << 4), o[0], o[1], o[2], o[3]);
                        st_shared_v4(ra + (((uint32_t)(2 * cg + 1) ^ xr) << 4), o[4], o[5], o[6], o[7]);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
                __syncwarp();
                if (lane == 0) {  // plane step blk is complete once blocks blk - 1 (its head) and blk are
                    mbar_arrive(pfull0 + 8 * blk);
                    if (blk < 3) mbar_arrive(pfull0 + 8 * (blk + 1));
                }
            }
        }
        if (DBG && warp == WARP_SEPI0 && lane == 0) {
            g_sprof[blockIdx.x * 16 + 2] = e_wa;
            g_sprof[blockIdx.x * 16 + 3] = e_wp;
            g_sprof[blockIdx.x * 16 + 4] = clock64() - e_t0;
        }
    } else {
        // ------------------------------------------------------------------ epilogue: TMEM -> bias + act -> bf16 -> global
        const int quarter = warp & 3;
        const int o = quarter * 32 + lane, r = o / PW, cc = o - r * PW;  // padded-flat output index -> (row, column) of the unit
        const int act = p.act1, C1n = p.C1;
        const int n_lo = 0, n_hi = C1n;
        const bool wide = p.wide_store != 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        pdl_wait();  // the output buffer may still be in use by the previous kernel (arena buffers are recycled)
        long long l_w = 0, l_t0 = DBG ? clock64() : 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            int b, oh0, ow0;
            unit_coords(u, b, oh0, ow0);
            const bool ok = r < TH && cc < TW && oh0 + r < p.H1 && ow0 + cc < p.W1;
            __nv_bfloat16* yrow = y + (((size_t)b * p.H1 + (oh0 + r)) * p.W1 + (ow0 + cc)) * p.out_pitch;
            {
                SP_T0();
                mbar_wait(tfull0 + 8 * acc, acc_phase);
                SP_ACC(l_w);
            }
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + STEM_STAGES * C0 + acc * C1n;
            const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll 1
            for (int n = n_lo; n < n_hi; n += 32) {
                const bool two = n + 16 < n_hi;
                uint32_t v0[16], v1[16];
                tmem_ld16(t_row + n, v0);
                if (two) tmem_ld16(t_row + n + 16, v1);
                tmem_ld_wait();
                float f[16];
                uint32_t ov[8];
                epi_math16(v0, bias_s + n, act, false, z, z, f);
                pack16(f, ov);
                if (ok) {
                    if (wide) st2_global_v8(yrow + n, ov);
                    else {
                        *reinterpret_cast<uint4*>(yrow + n) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
                        *reinterpret_cast<uint4*>(yrow + n + 8) = make_uint4(ov[4], ov[5], ov[6], ov[7]);
                    }
                }
                if (two) {
                    epi_math16(v1, bias_s + n + 16, act, false, z, z, f);
                    pack16(f, ov);
                    if (ok) {
                        if (wide) st2_global_v8(yrow + n + 16, ov);
                        else {
                            *reinterpret_cast<uint4*>(yrow + n + 16) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
                            *reinterpret_cast<uint4*>(yrow + n + 24) = make_uint4(ov[4], ov[5], ov[6], ov[7]);
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
            g_sprof[blockIdx.x * 16 + 10] = l_w;
            g_sprof[blockIdx.x * 16 + 11] = clock64() - l_t0;
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == WARP_EPI0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
}

size_t stem2_smem(const Stem2Params& p, int c0) {
    return 9ull * p.w_tile + (((size_t)ST_ROWS_ALLOC * c0 * 2 + 1023) & ~(size_t)1023) + A0_BYTES + (size_t)c0 * 64 + p.bias_bytes + 256 + 1024;
}

// Shape rules; pure arithmetic (also behind fce_stem2_route).
bool stem2_plan(const fce_stem2_desc* d, Stem2Params& p) {
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return false;
    if ((d->C0 != 64 && d->C0 != 32) || d->C1 % 16 || d->C1 < 16 || d->C1 > 256) return false;
    const int C0 = d->C0;
    if (d->H % 4 || d->W % 4 || d->H > 16000 || d->W > 16000) return false;  // whole words per image row, both strides exact
    if (d->out_pitch % 8 || d->out_off % 8) return false;
    if (d->act0 != FCE_ACT_SILU && d->act0 != FCE_ACT_NONE) return false;
    if (d->act1 != FCE_ACT_SILU && d->act1 != FCE_ACT_NONE) return false;
    p.B = d->B; p.H = d->H; p.W = d->W;
    p.H0 = d->H / 2; p.W0 = d->W / 2;
    p.H1 = d->H / 4; p.W1 = d->W / 4;
    p.C1 = d->C1;
    p.tiles_h = (p.H1 + TH - 1) / TH;
    p.tiles_w = (p.W1 + TW - 1) / TW;
    const long long units = (long long)d->B * p.tiles_h * p.tiles_w;
    if (units > 0x7fffffffLL) return false;
    p.units = (int)units;
    p.out_pitch = d->out_pitch;
    p.act0 = d->act0;
    p.act1 = d->act1;
    p.w_tile = (uint32_t)d->C1 * (uint32_t)C0 * 2u;
    p.bias_bytes = ((uint32_t)d->C1 * 4u + 255u) & ~255u;
    if (stem2_smem(p, C0) > (size_t)SMEM_LIMIT) return false;
    if (STEM_STAGES * C0 + 2 * d->C1 > 512) return false;  // tensor memory: four stem stages + two accumulator stages
    p.tmem_cols = 32;
    while (p.tmem_cols < (uint32_t)STEM_STAGES * C0 + 2u * (uint32_t)d->C1) p.tmem_cols <<= 1;  // stem stages + two accumulator stages
    // SBO = 8 rows of the swizzle span, descriptor version 1, 128B (layout 2) / 64B (layout 4) swizzle
    p.desc_hi = C0 == 64 ? ((1024u >> 4) | (1u << 14) | (2u << 29)) : ((512u >> 4) | (1u << 14) | (4u << 29));
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(d->C1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    p.desc_hi0 = (512u >> 4) | (1u << 14) | (4u << 29);  // stem GEMM: 8 rows x 64 bytes, 64B swizzle
    p.idesc0 = (1u << 4) | ((uint32_t)(C0 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // fp16 A / B, fp32 accumulate, N = C0
    return true;
}

}  // namespace
}  // namespace fce

using namespace fce;

#ifdef FCE_DEBUG
extern "C" int fce_stem2_profile(long long* out, int n) {  // debug builds only, not in the public header
    if (n > kNumSMs * 16) n = kNumSMs * 16;
    return cudaMemcpyFromSymbol(out, g_sprof, (size_t)n * sizeof(long long)) == cudaSuccess ? n : FCE_ERR_CUDA;
}
#endif

extern "C" int fce_stem2_route(const fce_stem2_desc* d) {
    if (!d) return FCE_ERR_BAD_ARG;
    Stem2Params p{};
    return stem2_plan(d, p) ? 1 : 0;
}

extern "C" int fce_stem2_conv(const fce_stem2_desc* d, const void* x, const void* w0, const float* b0, const void* w1,
                              const float* b1, void* y, void* stream) {
    if (!d || !x || !w0 || !b0 || !w1 || !b1 || !y) return FCE_ERR_BAD_ARG;
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    Stem2Params p{};
    if (!stem2_plan(d, p)) return FCE_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) & 3) || (reinterpret_cast<uintptr_t>(w0) & 3) || !aligned16(w1) || !aligned16(y))
        return FCE_ERR_ALIGNMENT;
    const DriverApi& api = driver();
    if (!api.ok) return FCE_ERR_CUDA;
    __nv_bfloat16* yout = reinterpret_cast<__nv_bfloat16*>(y) + d->out_off;
    p.wide_store = (d->out_pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(yout) & 31) == 0) ? 1 : 0;
    alignas(64) CUtensorMap tmW;
    {
        const int C0 = d->C0;
        const cuuint64_t K = 9ull * C0;
        const cuuint64_t gdim[2] = {K, (cuuint64_t)d->C1};
        const cuuint64_t gstr[1] = {K * 2};
        const cuuint32_t box[2] = {(cuuint32_t)C0, (cuuint32_t)d->C1};
        const cuuint32_t est[2] = {1, 1};
        if (api.tiled(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w1), gdim, gstr, box, est,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, C0 == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCE_ERR_UNSUPPORTED;
    }
    typedef void (*KernelFn)(const CUtensorMap, const Stem2Params, const uint8_t*, const __nv_bfloat16*, const float*,
                             const float*, __nv_bfloat16*);
    static const KernelFn table[2] = {conv_stem2_kernel<32>, conv_stem2_kernel<64>};
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
    return launch_pdl(table[d->C0 == 64 ? 1 : 0], grid, NUM_THREADS, stem2_smem(p, d->C0), (cudaStream_t)stream, tmW, p,
                      reinterpret_cast<const uint8_t*>(x), reinterpret_cast<const __nv_bfloat16*>(w0), b0, b1, yout);
}

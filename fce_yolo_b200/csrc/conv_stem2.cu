// The first TWO convolutions of the graph in one pass: Conv(3, C0, 3, 2) on the uint8 image followed by Conv(C0, C1, 3, 2)
// (yolo11-fce.yaml:20-21; Conv.forward_fuse conv.py:80-89), C0 = 64.
//
// As two launches the stem writes its [B, H/2, W/2, 64] map to HBM and the second conv reads it back: at m scale, batch
// 256, 640^2 that is 2 x 3.36 GB of the step's ~73 GB and 1.04 + 1.06 ms (the stem at 0.84 of its write-aware roofline,
// the second conv HBM-bound at 0.72 of the copy peak).  Here the stem map never exists in HBM.
//
//   unit = 7 x 16 output pixels of the second conv (M = 112 of the 128 rows of one tcgen05.mma).  It needs a 15 x 33
//   window of stem pixels, which needs 31 x 67 image pixels.
//   * stem warps (11): stage the image window as fp16 (byte / 256) in shared memory, then compute the stem window with
//     warp-level mma.sync m16n8k16 (the arithmetic of stem.cu: same K order, same operand scaling, bias + SiLU on packed
//     fp32 pairs) in runs of 16 stem pixels, and write bf16 rows of 64 channels (128 bytes) into the STEM TILE.  Stem
//     pixels outside the stem map are written as zeros: they are the zero padding of the second conv.
//   * the stem tile is laid out as FOUR PARITY PLANES (stem row parity x column parity), each a padded-flat pixel list
//     of pitch 17, rows 128B-swizzled by their absolute shared-memory address.  A stride-2 3x3 tap (kh, kw) of output
//     pixel o = r * 17 + c reads plane (kh & 1, kw & 1) at flat index o + (kh >> 1) * 17 + (kw >> 1): every tap's A operand
//     is ONE shifted shared-memory descriptor over a plane (conv_halo.cu's trick, per parity) - no im2col traffic at all.
//     The 17th column of every row produces accumulator rows that are never stored.
//   * MMA warp: the planes are produced and consumed in the order P11, P10, P01, P00 (1, 2, 2, 4 taps); each plane has its
//     own full / free barrier, so the stem warps refill plane q for the next unit while the tensor core still works on
//     planes q+1.. of this one - one stem tile in shared memory behaves like a four-stage ring.  36 tcgen05.mma (M = 128,
//     N = C1, K = 16) per unit against the PARKED weights (all nine [C1, 64] tap tiles, loaded once per CTA).
//   * epilogue warps (4): TMEM -> bias + SiLU -> bf16 -> 32-byte global stores, two accumulator stages.
//
// HBM bytes per output pixel: 48 image bytes (+ halo, L2) read, 2 * C1 written.  m scale, batch 256: 0.3 + 1.68 GB instead
// of 0.3 + 3.36 + 3.36 + 1.68 GB.
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace fce {
using namespace tc;
namespace {

constexpr int C0 = 64;                      // stem channels = K chunk of the second conv (128-byte rows)
constexpr int NT = C0 / 8;                  // mma.sync n-tiles of the stem
constexpr int TH = 7, TW = 16, PW = TW + 1; // unit of the second conv; plane pitch
constexpr int SR = 2 * TH + 1, SC = 2 * TW + 1;     // stem window 15 x 33
constexpr int IR = 2 * SR + 1;                      // 31 image rows
constexpr int ISHIFT = 1;                           // one extra image column on the left: the byte origin is 4-byte aligned
constexpr int HP = 208;                             // fp16 elements per staged row: element 1 + b holds byte b of the row
constexpr int NSTEM_WARPS = 11, STEM_THREADS = NSTEM_WARPS * 32;
constexpr int WARP_MMA = NSTEM_WARPS, WARP_EPI0 = WARP_MMA + 1;  // 12 .. 15: warp & 3 = TMEM lane quarter
constexpr int NUM_THREADS = (WARP_EPI0 + 4) * 32;   // 512: four warps per scheduler, 128 registers per thread
// planes in memory order P00, P01, P10, P11 (index = (row parity) * 2 + column parity)
__host__ __device__ constexpr int pl_rows(int pl) { return pl < 2 ? 8 * PW : 7 * PW; }  // 136 136 119 119 pixel slots
__host__ __device__ constexpr int pl_base(int pl) { return pl == 0 ? 0 : pl == 1 ? 8 * PW : pl == 2 ? 16 * PW : 23 * PW; }  // 0 136 272 391
constexpr int ST_ROWS_ALLOC = 520;  // the last descriptor reads rows [391 + 1, +128)
// production / consumption order q = 0..3 -> plane P11, P10, P01, P00, and the global list of runs (16 slots each): 8, 8, 9, 9
__host__ __device__ constexpr int q_plane(int q) { return 3 - q; }
__host__ __device__ constexpr int q_run0(int q) { return q == 0 ? 0 : q == 1 ? 8 : q == 2 ? 16 : q == 3 ? 25 : 34; }
constexpr uint32_t STILE_BYTES = ST_ROWS_ALLOC * 128;
constexpr uint32_t ITILE_BYTES = IR * HP * 2;
constexpr int SMEM_LIMIT = 227 * 1024;

struct Stem2Params {
    int B, H, W;        // image
    int H0, W0;         // stem map
    int H1, W1, C1;     // output
    int tiles_h, tiles_w, units;
    int out_pitch, act0, act1;
    uint32_t w_tile;    // bytes of one tap's weight tile: C1 x 128
    uint32_t bias_bytes, tmem_cols;
    uint32_t desc_hi, idesc;
    int wide_store;
};

__device__ __forceinline__ void st2_global_v8(void* p, const uint32_t* o) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                 "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                 : "memory");
}
__device__ __forceinline__ void stem_bar() { asm volatile("bar.sync 1, %0;" ::"n"(STEM_THREADS) : "memory"); }

// K axis of the stem GEMM, as in stem.cu: k' 0..23 = the first eight values of patch rows 0 / 1 / 2, k' 24..29 = (value 8,
// the zero-weight neighbour 9) of rows 0 / 1 / 2, k' 30..31 zero.
__host__ __device__ constexpr int s2_kh(int kp) { return kp < 24 ? kp / 8 : (kp - 24) / 2; }
__host__ __device__ constexpr int s2_j(int kp) { return kp < 24 ? kp % 8 : 8 + (kp - 24) % 2; }

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_stem2_kernel(const __grid_constant__ CUtensorMap tmW, const Stem2Params p, const uint8_t* __restrict__ x,
                  const __nv_bfloat16* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ b1,
                  __nv_bfloat16* __restrict__ y) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw0 = smem_u32(smem_raw);
    const uint32_t base = (raw0 + 1023u) & ~1023u;
    const uint32_t sW = base;                           // 9 x [C1][64] bf16, 128B-swizzled
    const uint32_t sT = sW + 9 * p.w_tile;              // stem tile: four parity planes, 128-byte rows
    const uint32_t sI = sT + STILE_BYTES;               // staged image window, fp16 [IR][HP]
    const uint32_t sBias = sI + ITILE_BYTES;
    const uint32_t bars = sBias + p.bias_bytes;
    const uint32_t pfull0 = bars, pfree0 = bars + 32, tfull0 = bars + 64, tempty0 = bars + 80, wfull = bars + 96;
    const uint32_t tmem_slot = bars + 104;
    float* bias_s = reinterpret_cast<float*>(smem_raw + (sBias - raw0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = p.units;

    pdl_launch_dependents();
    if (warp == WARP_MMA && lane == 0) {
        for (int q = 0; q < 4; ++q) {
            mbar_init(pfull0 + 8 * q, NSTEM_WARPS);
            mbar_init(pfree0 + 8 * q, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, 4);
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
        const float bsc = epi_bias_scale(p.act1);  // pre-scaled for epi_math16
        for (int i = threadIdx.x; i < (int)(p.bias_bytes >> 2); i += NUM_THREADS) bias_s[i] = i < p.C1 ? b1[i] * bsc : 0.f;
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

    if (warp == WARP_MMA) {
        // ------------------------------------------------------------------ weights (once), then the MMA issue loop
        if (elect_one()) {
            mbar_expect_tx(wfull, 9u * p.w_tile);
            for (int t = 0; t < 9; ++t) tma_load_2d(sW + t * p.w_tile, &tmW, wfull, t * C0, 0);
        }
        __syncwarp();
        mbar_wait(wfull, 0);
        tc_fence_after();
        const uint32_t dhi = p.desc_hi, idesc = p.idesc;
        int acc = 0, it = 0;
        uint32_t acc_phase = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
            mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * p.C1;
            bool first = true;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                mbar_wait(pfull0 + 8 * q, (uint32_t)it & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const int pl = q_plane(q), pr = pl >> 1, pc = pl & 1;
#pragma unroll
                    for (int kh = pr; kh < 3; kh += 2)
#pragma unroll
                        for (int kw = pc; kw < 3; kw += 2) {
                            const uint32_t a16 = (sT + (uint32_t)(pl_base(pl) + (kh >> 1) * PW + (kw >> 1)) * 128u) >> 4;
                            const uint32_t b16 = (sW + (uint32_t)(kh * 3 + kw) * p.w_tile) >> 4;
#pragma unroll
                            for (int k = 0; k < C0 / 16; ++k) {
                                const uint64_t ad = make_desc(dhi, ((a16 + 2 * k) & 0x3FFF) | (1u << 16));
                                const uint64_t bd = make_desc(dhi, ((b16 + 2 * k) & 0x3FFF) | (1u << 16));
                                umma_bf16(d_tmem, ad, bd, idesc, first ? 0u : 1u);
                                first = false;
                            }
                        }
                    umma_commit(pfree0 + 8 * q);             // the stem warps may refill this plane
                    if (q == 3) umma_commit(tfull0 + 8 * acc);  // accumulator complete -> epilogue
                }
                __syncwarp();
                first = false;
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else if (warp < NSTEM_WARPS) {
        // ------------------------------------------------------------------ stem warps
        const int tid = threadIdx.x, g = lane >> 2, t = lane & 3;
        // B fragments (stem weights, fp16) in the re-ordered K axis; bf16 -> fp16 is exact, x 256 because the staged image is
        // byte / 256, x 1/2 for the SiLU form h + h * tanh(h) (stem.cu)
        uint32_t bf[2][NT][2];
        float bs[NT][2];
        {
            const unsigned short* wus = reinterpret_cast<const unsigned short*>(w0);
            const float sc = 256.f * (p.act0 == FCE_ACT_SILU ? 0.5f : 1.f);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const unsigned short* wr = wus + (nt * 8 + g) * 32;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t v = 0u;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int kp = ks * 16 + hf * 8 + 2 * t + e;
                            const int kh = s2_kh(kp), j = s2_j(kp);
                            if (kp < 30 && j < 9) v |= (uint32_t)__ldg(wr + kh * 9 + j) << (16 * e);
                        }
                        const __half2 h = __floats2half2_rn(__uint_as_float(v << 16) * sc, __uint_as_float(v & 0xffff0000u) * sc);
                        bf[ks][nt][hf] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                const float bscale = p.act0 == FCE_ACT_SILU ? 0.5f : 1.f;
                bs[nt][0] = __ldg(b0 + nt * 8 + 2 * t) * bscale;
                bs[nt][1] = __ldg(b0 + nt * 8 + 2 * t + 1) * bscale;
            }
        }
        // this thread's 4 A-fragment word offsets relative to a patch origin (32-bit words of the fp16 window)
        int aoff[2][2];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int kp = ks * 16 + hf * 8 + 2 * t;
                aoff[ks][hf] = kp < 30 ? (s2_kh(kp) * HP + s2_j(kp)) / 2 : -1;
            }
        const uint32_t* xw = reinterpret_cast<const uint32_t*>(x);
        const int row_words = p.W * 3 / 4;
        constexpr int GROUPS = IR * (HP / 8);                          // 806 groups of eight staged values
        constexpr int TRIPS = (GROUPS + STEM_THREADS - 1) / STEM_THREADS;  // 3
        uint32_t wa[TRIPS], wb[TRIPS], wc[TRIPS];
        auto fetch = [&](int u) {  // this thread's words of unit u's image window -> registers
            int b, oh0, ow0;
            unit_coords(u, b, oh0, ow0);
            const int ih0 = 4 * oh0 - 3;
            const int w_first = (4 * ow0 - 3 - ISHIFT) * 3 / 4;  // exact: (4 ow0 - 4) * 3 is a multiple of 4 (may be negative)
#pragma unroll
            for (int k = 0; k < TRIPS; ++k) {
                const int i = tid + k * STEM_THREADS;
                const int r = i / (HP / 8), qq = 2 * (i - r * (HP / 8));
                const int hi = ih0 + r, wq = w_first + qq;
                wa[k] = wb[k] = wc[k] = 0u;
                if (i < GROUPS && hi >= 0 && hi < p.H) {
                    const uint32_t* rowp = xw + (size_t)(b * p.H + hi) * row_words;
                    if (wq - 1 >= 0 && wq - 1 < row_words) wa[k] = __ldg(rowp + wq - 1);
                    if (wq >= 0 && wq < row_words) wb[k] = __ldg(rowp + wq);
                    if (wq + 1 >= 0 && wq + 1 < row_words) wc[k] = __ldg(rowp + wq + 1);
                }
            }
        };
        uint4* iw = reinterpret_cast<uint4*>(smem_raw + (sI - raw0));
        const uint32_t* il = reinterpret_cast<const uint32_t*>(smem_raw + (sI - raw0));
        const __half2 four = __float2half2_rn(4.f);
        auto cvt2 = [&](uint32_t v, uint32_t sel) {  // two bytes -> two fp16 (byte / 256): 0x4400 | byte = 4 + byte / 256
            const uint32_t m = __byte_perm(v, 0x44444444u, sel);
            const __half2 h = __hsub2(*reinterpret_cast<const __half2*>(&m), four);
            return *reinterpret_cast<const uint32_t*>(&h);
        };
        const bool silu = p.act0 == FCE_ACT_SILU;
        pdl_wait();  // the image comes from the previous work in the stream
        if ((int)blockIdx.x < units) fetch(blockIdx.x);
        int it = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
            int b, oh0, ow0;
            unit_coords(u, b, oh0, ow0);
            const int sr0 = 2 * oh0 - 1, sc0 = 2 * ow0 - 1;  // stem coordinates of the window origin
            // ---- stage this unit's window (already in registers), then prefetch the next unit's
#pragma unroll
            for (int k = 0; k < TRIPS; ++k) {
                const int i = tid + k * STEM_THREADS;
                if (i >= GROUPS) continue;
                const uint32_t v0 = __funnelshift_l(wa[k], wb[k], 8), v1 = __funnelshift_l(wb[k], wc[k], 8);
                iw[i] = make_uint4(cvt2(v0, 0x4140), cvt2(v0, 0x4342), cvt2(v1, 0x4140), cvt2(v1, 0x4342));
            }
            stem_bar();
            if (u + (int)gridDim.x < units) fetch(u + gridDim.x);
            // ---- the stem window, plane by plane (order P11, P10, P01, P00), runs of 16 slots dealt round-robin to the warps
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                const int pl = q_plane(q), pr = pl >> 1, pc = pl & 1;
                mbar_wait(pfree0 + 8 * q, ((uint32_t)it & 1u) ^ 1u);  // the MMAs of the previous unit are done with this plane
                const int n_slots = pl_rows(pl);
                const uint32_t pbase = sT + (uint32_t)pl_base(pl) * 128u;
                const int run_lo = q_run0(q), run_hi = q_run0(q + 1);
#pragma unroll 1
                for (int run = run_lo + ((warp - run_lo % NSTEM_WARPS + NSTEM_WARPS) % NSTEM_WARPS); run < run_hi;
                     run += NSTEM_WARPS) {
                    const int s0 = (run - run_lo) * 16;
                    // slots of fragment rows g and g + 8 -> stem window coordinates -> patch origin
                    int slot[2], po[2];
                    bool inside[2], live[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int s = s0 + g + 8 * h;
                        slot[h] = s;
                        const int pi = s / PW, pj = s - pi * PW;
                        const int r = 2 * pi + pr, c = 2 * pj + pc;  // stem window coordinates
                        live[h] = s < n_slots;
                        const int sr = sr0 + r, scc = sc0 + c;
                        inside[h] = live[h] && c < SC && sr >= 0 && sr < p.H0 && scc >= 0 && scc < p.W0;
                        // patch origin: image window row 2 r, element 1 + 3 * (2 c + ISHIFT) (even) -> 32-bit words
                        const int rr = live[h] && c < SC ? r : 0, cc = live[h] && c < SC ? c : 0;
                        po[h] = (2 * rr) * (HP / 2) + 3 * cc + (1 + 3 * ISHIFT) / 2;
                    }
                    uint32_t a[2][4];
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const int o = aoff[ks][hf];
                            a[ks][2 * hf] = o >= 0 ? il[po[0] + o] : 0u;
                            a[ks][2 * hf + 1] = o >= 0 ? il[po[1] + o] : 0u;
                        }
                    // swizzled row addresses of the two pixels (absolute address bits 7..9 select the XOR)
                    const uint32_t row0 = pbase + (uint32_t)slot[0] * 128u, row1 = pbase + (uint32_t)slot[1] * 128u;
                    const uint32_t x0 = (row0 >> 7) & 7u, x1 = (row1 >> 7) & 7u;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        float c[4] = {bs[nt][0], bs[nt][1], bs[nt][0], bs[nt][1]};
                        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                                     : "r"(a[0][0]), "r"(a[0][1]), "r"(a[0][2]), "r"(a[0][3]), "r"(bf[0][nt][0]), "r"(bf[0][nt][1]));
                        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                                     : "r"(a[1][0]), "r"(a[1][1]), "r"(a[1][2]), "r"(a[1][3]), "r"(bf[1][nt][0]), "r"(bf[1][nt][1]));
                        if (silu) {  // c holds h = v / 2: silu(v) = h + h * tanh(h), on packed pairs
#pragma unroll
                            for (int j = 0; j < 4; j += 2) {
                                const float2 h = make_float2(c[j], c[j + 1]);
                                const float2 o = __ffma2_rn(h, make_float2(tanh_fast(h.x), tanh_fast(h.y)), h);
                                c[j] = o.x;
                                c[j + 1] = o.y;
                            }
                        }
                        __nv_bfloat162 lo = __floats2bfloat162_rn(c[0], c[1]), hi = __floats2bfloat162_rn(c[2], c[3]);
                        const uint32_t vlo = inside[0] ? *reinterpret_cast<uint32_t*>(&lo) : 0u;  // outside the stem map: the
                        const uint32_t vhi = inside[1] ? *reinterpret_cast<uint32_t*>(&hi) : 0u;  // second conv's zero padding
                        if (live[0])
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(row0 + ((((uint32_t)nt ^ x0) << 4) | (uint32_t)(t << 2))), "r"(vlo)
                                         : "memory");
                        if (live[1])
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(row1 + ((((uint32_t)nt ^ x1) << 4) | (uint32_t)(t << 2))), "r"(vhi)
                                         : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
                __syncwarp();
                if (lane == 0) mbar_arrive(pfull0 + 8 * q);
            }
            stem_bar();  // every warp is done reading the staged window
        }
    } else {
        // ------------------------------------------------------------------ epilogue: TMEM -> bias + act -> bf16 -> global
        const int quarter = warp & 3;
        const int o = quarter * 32 + lane, r = o / PW, cc = o - r * PW;  // padded-flat output index -> (row, column) of the unit
        const int act = p.act1, C1n = p.C1;
        const bool wide = p.wide_store != 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        pdl_wait();  // the output buffer may still be in use by the previous kernel (arena buffers are recycled)
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            int b, oh0, ow0;
            unit_coords(u, b, oh0, ow0);
            const bool ok = r < TH && cc < TW && oh0 + r < p.H1 && ow0 + cc < p.W1;
            __nv_bfloat16* yrow = y + (((size_t)b * p.H1 + (oh0 + r)) * p.W1 + (ow0 + cc)) * p.out_pitch;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * C1n;
            const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll 1
            for (int n = 0; n < C1n; n += 32) {
                const bool two = n + 16 < C1n;
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
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == WARP_EPI0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
}

size_t stem2_smem(const Stem2Params& p) {
    return 9ull * p.w_tile + STILE_BYTES + ITILE_BYTES + p.bias_bytes + 256 + 1024;
}

// Shape rules; pure arithmetic (also behind fce_stem2_route).
bool stem2_plan(const fce_stem2_desc* d, Stem2Params& p) {
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return false;
    if (d->C0 != C0 || d->C1 % 16 || d->C1 < 16 || d->C1 > 256) return false;
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
    p.w_tile = (uint32_t)d->C1 * 128u;
    p.bias_bytes = ((uint32_t)d->C1 * 4u + 255u) & ~255u;
    if (stem2_smem(p) > (size_t)SMEM_LIMIT) return false;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2u * (uint32_t)d->C1) p.tmem_cols <<= 1;
    p.desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 8 rows x 128 bytes, descriptor version 1, 128B swizzle
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(d->C1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    return true;
}

}  // namespace
}  // namespace fce

using namespace fce;

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
        const cuuint64_t K = 9ull * C0;
        const cuuint64_t gdim[2] = {K, (cuuint64_t)d->C1};
        const cuuint64_t gstr[1] = {K * 2};
        const cuuint32_t box[2] = {(cuuint32_t)C0, (cuuint32_t)d->C1};
        const cuuint32_t est[2] = {1, 1};
        if (api.tiled(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w1), gdim, gstr, box, est,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCE_ERR_UNSUPPORTED;
    }
    static DeviceOnce attr_once;  // the shared-memory opt-in is a per-device attribute
    int dev = 0;
    if (attr_once.pending(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(conv_stem2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) {
            set_cuda_error(e);
            return FCE_ERR_CUDA;
        }
        attr_once.done(dev);
    }
    const int grid = p.units < kNumSMs ? p.units : kNumSMs;
    return launch_pdl(conv_stem2_kernel, grid, NUM_THREADS, stem2_smem(p), (cudaStream_t)stream, tmW, p,
                      reinterpret_cast<const uint8_t*>(x), reinterpret_cast<const __nv_bfloat16*>(w0), b0, b1, yout);
}

// Depthwise 3x3 (stride 1, pad 1) fed by TMA: DWConv of the Detect class branch (ultralytics/nn/modules/conv.py:185-199
// in head.py:101-102) and Attention.pe with its add (block.py:1282,1302).  HBM-bound: 2 (or 3) x C*H*W*e bytes.
//
// A persistent CTA walks (image, 10-row x 20-column tile, 128-byte channel group) units - 20 x 10 tiles cover the
// 20 / 40 / 80 / 160-pixel maps of 640^2 and 1280^2 inputs without ragged edges.  ONE tiled-mode TMA box {128 bytes of
// channels, 22 columns, 12 rows} per unit lands the zero-padded input tile in shared memory (the TMA unit's
// out-of-bounds zero fill IS the conv padding, at the image border and for ragged tiles alike) through a full/empty
// ring: two stages of 34 KB per CTA and three CTAs (15 warps) per SM by default - measured against three stages / two
// CTAs: 69.6 vs 81.9 us at 80x80x128, batch 64 (FCE_DW_STAGES=3 selects the deeper ring).  The kernel is paced by HBM and
// by the latency of its unit loop (barrier + TMA wait), not by per-thread load latency.
// The first TMA version (one 16-byte channel vector of ONE column per thread, scalar FMAs) was issue-bound: 62 % issue
// utilisation at 2.6 TB/s, ~170 instructions per output vector.  Now a thread owns 8 bytes of channels (4 bf16 / 2
// fp32) of TWO adjacent columns and walks down the 12 box rows: four 8-byte shared-memory reads per row feed both
// columns (2 loads + 2 unpacks per output instead of 3), and all arithmetic runs on packed fp32 pairs (FFMA2, sm_100's
// two-wide fp32 pipe): 9 FFMA2 per channel pair per output.  Three rolling accumulator rows (outputs r-2, r-1, r);
// one 8-byte global store per output pixel, + the optional add operand (which may alias the destination: the same
// thread reads it right before it writes).
// No tensor cores: there is no contraction over channels.
#include "tc_common.cuh"

namespace fce {
using namespace tc;
namespace {

constexpr int TH = 10, TW = 20;           // output tile
constexpr int BR = TH + 2, BC = TW + 2;   // input box rows / columns
constexpr int LANES = 16;                 // 8-byte channel lanes per pixel and channel group (128 bytes)
constexpr int THREADS = (TW / 2) * LANES; // 160
constexpr uint32_t STAGE_BYTES = BR * BC * 128;  // 33792

struct DwParams {
    int B, H, W, C;
    int tiles_h, tiles_w, cgroups;  // unit = ((cg * B + b) * tiles_h + th) * tiles_w + tw : channel group slowest
    int units;
    int out_pitch, add_pitch;       // elements
    int act;
};

// 8 bytes of channels as packed fp32 pairs: 2 pairs for bf16, 1 for fp32
template <typename T>
struct Lane8 {
    static constexpr int NP = 4 / sizeof(T);  // float2 pairs
    static constexpr int NC = 2 * NP;         // channels
    static __device__ __forceinline__ void unpack(uint2 r, float2* f) {
        if constexpr (sizeof(T) == 2) {
            f[0] = make_float2(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u));
            f[1] = make_float2(__uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
        } else {
            f[0] = make_float2(__uint_as_float(r.x), __uint_as_float(r.y));
        }
    }
    static __device__ __forceinline__ uint2 pack(const float2* f) {
        if constexpr (sizeof(T) == 2) {
            __nv_bfloat162 a = __floats2bfloat162_rn(f[0].x, f[0].y), b = __floats2bfloat162_rn(f[1].x, f[1].y);
            return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
        } else {
            return make_uint2(__float_as_uint(f[0].x), __float_as_uint(f[0].y));
        }
    }
};

template <typename T, int STAGES>
__global__ void __launch_bounds__(THREADS) dwconv_tma_kernel(const __grid_constant__ CUtensorMap tmX, const DwParams p,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             const T* add, T* __restrict__ y) {
    pdl_trigger();
    constexpr int NP = Lane8<T>::NP, NC = Lane8<T>::NC;
    constexpr int CB = LANES * NC;  // channels per group
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint8_t* sgen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    const int tid = threadIdx.x, lane8 = tid % LANES, col = 2 * (tid / LANES);  // first of this thread's two columns

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmX);
    }
    __syncthreads();

    auto unit_coords = [&](int u, int& cg, int& b, int& h0, int& w0) {
        const int tw = u % p.tiles_w;
        int r = u / p.tiles_w;
        const int th = r % p.tiles_h;
        r /= p.tiles_h;
        b = r % p.B;
        cg = r / p.B;
        h0 = th * TH;
        w0 = tw * TW;
    };
    auto issue = [&](int u, int stage) {  // one elected thread
        int cg, b, h0, w0;
        unit_coords(u, cg, b, h0, w0);
        const uint32_t bar = bars + 8 * stage;
        mbar_expect_tx(bar, STAGE_BYTES);
        tma_load_4d(base + stage * STAGE_BYTES, &tmX, bar, cg * CB, w0 - 1, h0 - 1, b);
    };
    const int first = blockIdx.x, step = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s)
            if (first + s * step < p.units) issue(first + s * step, s);
    }

    float2 wt[9][NP], bs[NP];
    int cur_cg = -1;
    int it = 0;
    for (int u = first; u < p.units; u += step, ++it) {
        const int stage = it % STAGES;
        const uint32_t phase = (uint32_t)(it / STAGES) & 1u;
        int cg, b, h0, w0;
        unit_coords(u, cg, b, h0, w0);
        const int c = cg * CB + lane8 * NC;
        const bool c_ok = c < p.C;
        if (cg != cur_cg) {  // rare: the channel group is the slowest unit coordinate
            cur_cg = cg;
            const int cc = c_ok ? c : 0;
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int j = 0; j < NP; ++j)
                    wt[t][j] = make_float2(__ldg(w + t * p.C + cc + 2 * j), __ldg(w + t * p.C + cc + 2 * j + 1));
#pragma unroll
            for (int j = 0; j < NP; ++j) bs[j] = make_float2(__ldg(bias + cc + 2 * j), __ldg(bias + cc + 2 * j + 1));
        }
        mbar_wait(bars + 8 * stage, phase);
        const uint8_t* tile = sgen + stage * STAGE_BYTES + lane8 * 8;
        const int wc = w0 + col;
        const bool ok0 = c_ok && wc < p.W, ok1 = c_ok && wc + 1 < p.W;
        // acc[o % 3][q]: output row o of column q - the slot index is a compile-time constant of the unrolled row
        // loop, so the three rolling rows never move between registers
        float2 acc[3][2][NP];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int j = 0; j < NP; ++j) acc[k][q][j] = bs[j];
        // pointers of output row h0 (advanced by one image row per finished output row)
        const size_t pix0 = ((size_t)b * p.H + h0) * p.W + wc;
        T* yp = y + pix0 * p.out_pitch + c;
        const T* ap = add ? add + pix0 * p.add_pitch + c : nullptr;
        const size_t y_row = (size_t)p.W * p.out_pitch, a_row = (size_t)p.W * p.add_pitch;
        const int rows_ok = p.H - h0;  // output rows of this tile inside the image
        const bool silu = p.act == FCE_ACT_SILU, plain = p.act == FCE_ACT_NONE;
#pragma unroll
        for (int br = 0; br < BR; ++br) {
            float2 v[4][NP];
            const uint8_t* rowp = tile + (br * BC + col) * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) Lane8<T>::unpack(*reinterpret_cast<const uint2*>(rowp + k * 128), v[k]);
            // the add operand of the row this iteration finishes: issued now, consumed after the FMAs
            uint2 araw[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
            if (ap && br >= 2 && br - 2 < rows_ok) {
                if (ok0) araw[0] = *reinterpret_cast<const uint2*>(ap);
                if (ok1) araw[1] = *reinterpret_cast<const uint2*>(ap + p.add_pitch);
            }
            // box row br is tap row kh of output row br - kh
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int o = br - kh;  // compile-time after unrolling
                if (o < 0 || o >= TH) continue;
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                        for (int j = 0; j < NP; ++j)
                            acc[o % 3][q][j] = __ffma2_rn(v[q + kw][j], wt[3 * kh + kw][j], acc[o % 3][q][j]);
            }
            if (br >= 2) {
                const int o = br - 2;  // finished output row
                if (o < rows_ok) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (q == 0 ? ok0 : ok1) {
                            float2 out[NP];
#pragma unroll
                            for (int j = 0; j < NP; ++j) {
                                const float2 a = acc[o % 3][q][j];
                                if (plain) {
                                    out[j] = a;
                                } else if (sizeof(T) == 2 && silu) {  // h + h*tanh(h), h = a/2: one MUFU per element
                                    const float2 h = __fmul2_rn(a, make_float2(0.5f, 0.5f));
                                    out[j] = __ffma2_rn(h, make_float2(tanh_fast(h.x), tanh_fast(h.y)), h);
                                } else if (sizeof(T) == 2) {
                                    out[j] = make_float2(act_fast(a.x, p.act), act_fast(a.y, p.act));
                                } else {
                                    out[j] = make_float2(apply_act(a.x, p.act), apply_act(a.y, p.act));
                                }
                            }
                            if (ap) {
                                float2 af[NP];
                                Lane8<T>::unpack(araw[q], af);
#pragma unroll
                                for (int j = 0; j < NP; ++j) out[j] = __fadd2_rn(out[j], af[j]);
                            }
                            *reinterpret_cast<uint2*>(yp + (size_t)q * p.out_pitch) = Lane8<T>::pack(out);
                        }
                    }
                }
                yp += y_row;
                if (ap) ap += a_row;
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int j = 0; j < NP; ++j) acc[o % 3][q][j] = bs[j];
            }
        }
        __syncthreads();  // every thread is done reading this stage
        if (tid == 0 && u + STAGES * step < p.units) issue(u + STAGES * step, stage);
    }
}

}  // namespace

// Returns FCE_ERR_UNSUPPORTED when the views cannot be described to TMA (the caller falls back to the register
// window kernel); x / y / add already point at the first channel of their views.
int dwconv3x3_tma(const fce_dwconv_desc* d, const void* x, const float* w, const float* bias, const void* add, void* y,
                  cudaStream_t st) {
    const DriverApi& api = driver();
    if (!api.ok) return FCE_ERR_UNSUPPORTED;
    const int esz = d->dtype == FCE_BF16 ? 2 : 4;
    const int n = 16 / esz;
    if (d->dtype != FCE_BF16 && d->dtype != FCE_F32) return FCE_ERR_UNSUPPORTED;
    if (d->C % n || d->in_pitch % n || d->out_pitch % n || (add && d->add_pitch % n)) return FCE_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(y) || (add && !aligned16(add))) return FCE_ERR_UNSUPPORTED;
    const int cb = 128 / esz;
    DwParams p;
    p.B = d->B; p.H = d->H; p.W = d->W; p.C = d->C;
    p.tiles_h = (d->H + TH - 1) / TH;
    p.tiles_w = (d->W + TW - 1) / TW;
    p.cgroups = (d->C + cb - 1) / cb;
    const long long units = (long long)p.cgroups * d->B * p.tiles_h * p.tiles_w;
    if (units > 0x7fffffffLL) return FCE_ERR_UNSUPPORTED;
    p.units = (int)units;
    p.out_pitch = d->out_pitch;
    p.add_pitch = d->add_pitch;
    p.act = d->act;
    CUtensorMap tmX;
    {
        const cuuint64_t gdim[4] = {(cuuint64_t)d->C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
        const cuuint64_t gstr[3] = {(cuuint64_t)d->in_pitch * esz, (cuuint64_t)d->W * d->in_pitch * esz,
                                    (cuuint64_t)d->H * d->W * d->in_pitch * esz};
        const cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)BC, (cuuint32_t)BR, 1};
        const cuuint32_t est[4] = {1, 1, 1, 1};
        if (api.tiled(&tmX, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                      const_cast<void*>(x), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCE_ERR_UNSUPPORTED;
    }
    // ring depth: 2 stages -> 68 KB, three CTAs (15 warps) per SM; 3 stages -> 101 KB, two CTAs (FCE_DW_STAGES to vary)
    static const int stages = [] {
        const char* e = getenv("FCE_DW_STAGES");
        const int v = e ? atoi(e) : 2;
        return v == 3 ? 3 : 2;
    }();
    const size_t smem = (size_t)stages * STAGE_BYTES + 8 * stages + 128;
    typedef void (*KernelFn)(const CUtensorMap, const DwParams, const float*, const float*, const void*, void*);
    static const KernelFn table[2][2] = {
        {(KernelFn)(void*)dwconv_tma_kernel<__nv_bfloat16, 2>, (KernelFn)(void*)dwconv_tma_kernel<__nv_bfloat16, 3>},
        {(KernelFn)(void*)dwconv_tma_kernel<float, 2>, (KernelFn)(void*)dwconv_tma_kernel<float, 3>}};
    static DeviceOnce attr_once;  // per-device attribute
    int dev_ = 0;
    if (attr_once.pending(&dev_)) {
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
                cudaError_t e = cudaFuncSetAttribute((const void*)table[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     3 * (int)STAGE_BYTES + 256);
                if (e != cudaSuccess) {
                    set_cuda_error(e);
                    return FCE_ERR_CUDA;
                }
            }
        attr_once.done(dev_);
    }
    int grid = (stages == 2 ? 3 : 2) * kNumSMs;
    if (grid > p.units) grid = p.units;
    void* args[] = {(void*)&tmX, (void*)&p, (void*)&w, (void*)&bias, (void*)&add, (void*)&y};
    cudaError_t le = cudaLaunchKernel((const void*)table[esz == 2 ? 0 : 1][stages - 2], dim3(grid), dim3(THREADS), args, smem, st);
    if (le != cudaSuccess) {
        set_cuda_error(le);
        return FCE_ERR_CUDA;
    }
    return check_launch();
}

}  // namespace fce

// Depthwise 3x3 (stride 1, pad 1) fed by TMA: DWConv of the Detect class branch (ultralytics/nn/modules/conv.py:185-199
// in head.py:101-102) and Attention.pe with its add (block.py:1282,1302).  HBM-bound: 2 (or 3) x C*H*W*e bytes.
//
// A persistent CTA walks (image, 8-row x 16-column tile, 128-byte channel group) units.  ONE tiled-mode TMA box
// {128 bytes of channels, 18 columns, 10 rows} per unit lands the zero-padded input tile in shared memory (the TMA
// unit's out-of-bounds zero fill IS the conv padding, at the image border and for ragged tiles alike) through a
// 3-stage full/empty ring, so ~69 KB of loads are in flight per CTA and three CTAs share an SM: the kernel is paced
// by HBM, not by per-thread load latency (the register-window kernels it replaces ran at 25 % of HBM peak with
// 60 % issue utilisation).  A thread owns one 16-byte channel vector of one tile column and walks down the 10 box
// rows: three conflict-free 16-byte shared-memory reads per row scatter into three rolling accumulators (output
// rows r-2, r-1, r), 9 FMAs per channel per output; one 16-byte global store per output, + the optional add
// operand (which may alias the destination: the same thread reads it right before it writes).
// No tensor cores: there is no contraction over channels.
#include "tc_common.cuh"

namespace fce {
using namespace tc;
namespace {

constexpr int TH = 8, TW = 16;            // output tile
constexpr int BR = TH + 2, BC = TW + 2;   // input box rows / columns
constexpr int VECS = 8;                   // 16-byte vectors per pixel and channel group (128 bytes)
constexpr int THREADS = TW * VECS;        // 128
constexpr int STAGES = 3;
constexpr uint32_t STAGE_BYTES = BR * BC * 128;  // 23040

struct DwParams {
    int B, H, W, C;
    int tiles_h, tiles_w, cgroups;  // unit = ((cg * B + b) * tiles_h + th) * tiles_w + tw : channel group slowest
    int units;
    int out_pitch, add_pitch;       // elements
    int act;
};

template <typename T>
__global__ void __launch_bounds__(THREADS) dwconv_tma_kernel(const __grid_constant__ CUtensorMap tmX, const DwParams p,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             const T* add, T* __restrict__ y) {
    pdl_trigger();
    constexpr int N = Vec16<T>::N;  // channels per thread
    constexpr int CB = VECS * N;    // channels per group
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint8_t* sgen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    const int tid = threadIdx.x, vec = tid % VECS, col = tid / VECS;

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

    float wt[9][N], bs[N];
    int cur_cg = -1;
    int it = 0;
    for (int u = first; u < p.units; u += step, ++it) {
        const int stage = it % STAGES;
        const uint32_t phase = (uint32_t)(it / STAGES) & 1u;
        int cg, b, h0, w0;
        unit_coords(u, cg, b, h0, w0);
        const int c = cg * CB + vec * N;
        const bool c_ok = c < p.C;
        if (cg != cur_cg) {  // rare: the channel group is the slowest unit coordinate
            cur_cg = cg;
            const int cc = c_ok ? c : 0;
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int j = 0; j < N; ++j) wt[t][j] = __ldg(w + t * p.C + cc + j);
#pragma unroll
            for (int j = 0; j < N; ++j) bs[j] = __ldg(bias + cc + j);
        }
        mbar_wait(bars + 8 * stage, phase);
        const uint8_t* tile = sgen + stage * STAGE_BYTES + vec * 16;
        const int wc = w0 + col;
        const bool st_ok = c_ok && wc < p.W;
        float a0[N], a1[N], a2[N];
#pragma unroll
        for (int j = 0; j < N; ++j) a0[j] = a1[j] = a2[j] = bs[j];
#pragma unroll
        for (int br = 0; br < BR; ++br) {
            Vec16<T> vl, vm, vr;
            const uint8_t* rowp = tile + (br * BC + col) * 128;
            vl.raw = *reinterpret_cast<const uint4*>(rowp);
            vm.raw = *reinterpret_cast<const uint4*>(rowp + 128);
            vr.raw = *reinterpret_cast<const uint4*>(rowp + 256);
            float fl[N], fm[N], fr[N];
            vl.unpack(fl);
            vm.unpack(fm);
            vr.unpack(fr);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                // box row br is tap row kh = 2 of output br-2, kh = 1 of output br-1, kh = 0 of output br
                a0[j] = fmaf(fl[j], wt[6][j], fmaf(fm[j], wt[7][j], fmaf(fr[j], wt[8][j], a0[j])));
                a1[j] = fmaf(fl[j], wt[3][j], fmaf(fm[j], wt[4][j], fmaf(fr[j], wt[5][j], a1[j])));
                a2[j] = fmaf(fl[j], wt[0][j], fmaf(fm[j], wt[1][j], fmaf(fr[j], wt[2][j], a2[j])));
            }
            if (br >= 2) {
                const int ho = h0 + br - 2;
                if (st_ok && ho < p.H) {
                    float out[N];
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        if (sizeof(T) == 2) out[j] = act_fast(a0[j], p.act);
                        else out[j] = apply_act(a0[j], p.act);
                    }
                    const size_t pix = ((size_t)b * p.H + ho) * p.W + wc;
                    if (add) {
                        Vec16<T> av;
                        float af[N];
                        av.load(add + pix * p.add_pitch + c);
                        av.unpack(af);
#pragma unroll
                        for (int j = 0; j < N; ++j) out[j] += af[j];
                    }
                    Vec16<T> ov;
                    ov.pack(out);
                    ov.store(y + pix * p.out_pitch + c);
                }
            }
#pragma unroll
            for (int j = 0; j < N; ++j) {
                a0[j] = a1[j];
                a1[j] = a2[j];
                a2[j] = bs[j];
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
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 8 * STAGES + 128;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e1 = cudaFuncSetAttribute(dwconv_tma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaError_t e2 = cudaFuncSetAttribute(dwconv_tma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            set_cuda_error(e1 != cudaSuccess ? e1 : e2);
            return FCE_ERR_CUDA;
        }
        attr_set = true;
    }
    int grid = 3 * kNumSMs;
    if (grid > p.units) grid = p.units;
    if (esz == 2)
        dwconv_tma_kernel<__nv_bfloat16><<<grid, THREADS, smem, st>>>(tmX, p, w, bias, (const __nv_bfloat16*)add,
                                                                      (__nv_bfloat16*)y);
    else
        dwconv_tma_kernel<float><<<grid, THREADS, smem, st>>>(tmX, p, w, bias, (const float*)add, (float*)y);
    return check_launch();
}

}  // namespace fce

// Coordinate pooling fed by TMA (bf16, large batches): the H-pool / W-pool of CoordAtt, CoordCrossAtt and BiCoordCrossAtt
// (ultralytics/nn/modules/fce_block.py:81-82,101-102,140-141,159-160,212-213,239-240) in ONE pass over x.
// HBM-bound: algorithmic bytes = C*H*W*2 read once (+ (H+W)*C*4 written).
//
// The register-window kernel in coord.cu alternates load phases with reduction phases (a warp that is reducing has no
// load in flight): 47 % of the HBM roofline at batch 64.  Here the loads are asynchronous: a CTA owns (image, 64-channel
// chunk) and walks the image top to bottom in bands of RB rows through a three-stage ring of tiled-mode TMA boxes
// {128 bytes of channels, W columns, RB rows} - three CTAs per SM, each with up to two bands (<= 20 KB each) in flight while it reduces
// the third.  Per band: a thread owns an 8-byte channel lane of every 16th column, adds its values to the column sums it
// keeps in registers for the whole image and to RB row partials; the row partials are combined by one shuffle + a
// shared-memory transpose (8 warps), the finished row means go to the strip.  No workspace, no finish kernel, no atomics:
// deterministic.  Out-of-bounds rows / channels arrive as zeros (TMA fill) and add nothing.
#include "tc_common.cuh"

namespace fce {
using namespace tc;
namespace {

constexpr int PT = 256, STAGES = 3, CCH = 64;  // threads, ring depth, channels per CTA

template <int RB, int NCOL>
__global__ void __launch_bounds__(PT, 3) coord_pool_tma_kernel(const __grid_constant__ CUtensorMap tm, const fce_pool_desc d,
                                                              float* __restrict__ strip, int chunks, int units) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    __shared__ float red[PT / 32][RB][CCH];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint8_t* sgen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t stage_bytes = (uint32_t)RB * d.W * 128u;
    const uint32_t bars = base + STAGES * stage_bytes;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lane8 = tid & 15, wslot = tid >> 4;  // 8-byte channel lane (4 channels), column slot (every 16th column)
    const int nbands = (d.H + RB - 1) / RB;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tm);
    }
    __syncthreads();
    // Persistent: this CTA's units (image, channel chunk) are u = blockIdx.x, + gridDim.x, ...; their bands form ONE
    // sequence of items that the ring walks without draining between units (item -> unit index item / nbands).
    const int my_units = (units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_items = my_units * nbands;
    auto issue = [&](int item) {  // one thread
        const int ui = item / nbands, band = item - ui * nbands;
        const int u = (int)blockIdx.x + ui * (int)gridDim.x;
        const int b = u / chunks, chunk = u - b * chunks;
        const uint32_t bar = bars + 8 * (item % STAGES);
        mbar_expect_tx(bar, stage_bytes);
        tma_load_4d(base + (item % STAGES) * stage_bytes, &tm, bar, chunk * CCH, 0, band * RB, b);
    };
    if (tid == 0)
        for (int s = 0; s < STAGES && s < n_items; ++s) issue(s);

    const float inv_w = 1.f / (float)d.W, inv_h = 1.f / (float)d.H;
    int item = 0;
    for (int ui = 0; ui < my_units; ++ui) {
        const int u = (int)blockIdx.x + ui * (int)gridDim.x;
        const int b = u / chunks, chunk = u - b * chunks;
        float col[NCOL][4];
#pragma unroll
        for (int i = 0; i < NCOL; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) col[i][j] = 0.f;
        float* xh = strip + (size_t)b * d.H * d.C;
        for (int band = 0; band < nbands; ++band, ++item) {
            const int stage = item % STAGES;
            mbar_wait(bars + 8 * stage, (uint32_t)(item / STAGES) & 1u);
            const uint8_t* tile = sgen + stage * stage_bytes + lane8 * 8;
            float row[RB][4];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
#pragma unroll
                for (int j = 0; j < 4; ++j) row[r][j] = 0.f;
#pragma unroll
                for (int i = 0; i < NCOL; ++i) {
                    const int w = wslot + 16 * i;
                    if (w < d.W) {  // a warp reads 2 columns x 128 contiguous bytes: conflict-free
                        const uint2 v = *reinterpret_cast<const uint2*>(tile + (size_t)(r * d.W + w) * 128);
                        const float f0 = __uint_as_float(v.x << 16), f1 = __uint_as_float(v.x & 0xffff0000u);
                        const float f2 = __uint_as_float(v.y << 16), f3 = __uint_as_float(v.y & 0xffff0000u);
                        row[r][0] += f0; row[r][1] += f1; row[r][2] += f2; row[r][3] += f3;
                        col[i][0] += f0; col[i][1] += f1; col[i][2] += f2; col[i][3] += f3;
                    }
                }
            }
            // row partials: the two column slots of a warp by one shuffle, the eight warps through shared memory
#pragma unroll
            for (int r = 0; r < RB; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) row[r][j] += __shfl_xor_sync(0xffffffffu, row[r][j], 16);
            __syncthreads();  // every thread is done with this stage AND with the previous band's red[]
            if (tid == 0 && item + STAGES < n_items) issue(item + STAGES);
            if (lane < 16) {
#pragma unroll
                for (int r = 0; r < RB; ++r)
                    *reinterpret_cast<float4*>(&red[warp][r][lane8 * 4]) =
                        make_float4(row[r][0], row[r][1], row[r][2], row[r][3]);
            }
            __syncthreads();
            for (int o = tid; o < RB * CCH; o += PT) {
                const int r = o / CCH, c = o - r * CCH;
                const int h = band * RB + r, cg = chunk * CCH + c;
                if (h < d.H && cg < d.C) {
                    float s = 0.f;
#pragma unroll
                    for (int wq = 0; wq < PT / 32; ++wq) s += red[wq][r][c];
                    xh[(size_t)h * d.C + cg] = s * inv_w;
                }
            }
        }
        // column means: every (column, channel lane) is owned by exactly one thread
        const int c0 = chunk * CCH + lane8 * 4;
        if (c0 < d.C) {  // C is a multiple of 8: the four channels of a lane are inside or outside together
            float* colp = strip + ((size_t)d.B * d.H + (size_t)b * d.W) * d.C + c0;
#pragma unroll
            for (int i = 0; i < NCOL; ++i) {
                const int w = wslot + 16 * i;
                if (w < d.W)
                    *reinterpret_cast<float4*>(colp + (size_t)w * d.C) =
                        make_float4(col[i][0] * inv_h, col[i][1] * inv_h, col[i][2] * inv_h, col[i][3] * inv_h);
            }
        }
    }
}

}  // namespace

// FCE_ERR_UNSUPPORTED = not a shape for this kernel (the caller keeps the register-window kernel).
int coord_pool_tma(const fce_pool_desc* d, const void* x, float* strip, cudaStream_t st) {
    const DriverApi& api = driver();
    if (!api.ok || d->dtype != FCE_BF16) return FCE_ERR_UNSUPPORTED;
    // measured (batch 64, bf16): 80x80x256 59.4 -> 51.2 us; 40x40x512 36.9 -> 38.9 us - narrow maps keep the register-window
    // kernel (its 32-lane channel groups read 512 contiguous bytes per pixel; here a pixel contributes 128)
    if (d->W <= 48 || d->W > 160 || d->C % 8 || d->pitch % 8 || d->off % 8 || !aligned16(x) || !aligned16(strip)) return FCE_ERR_UNSUPPORTED;
    const int chunks = (d->C + CCH - 1) / CCH;
    static const int min_ctas = [] { const char* e = getenv("FCE_POOL_TMA_MIN"); return e && *e ? atoi(e) : 96; }();
    if ((long long)d->B * chunks < min_ctas) return FCE_ERR_UNSUPPORTED;  // small batches: bands supply the parallelism
    const int ncol = (d->W + 15) / 16;
    int rb, variant;
    // stages of at most 20 KB: three CTAs (3 x 3 stages) per SM - a single CTA per SM with 40 KB stages reached only half
    // of an SM's share of the HBM bandwidth (latency-bound ring), and 256 CTAs on 148 SMs quantise badly
    if (ncol <= 5) { rb = 2; variant = 0; }  // 49..80 columns
    else { rb = 1; variant = 1; }            // 81..160 columns
    CUtensorMap tm;
    {
        const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x) + d->off;
        const cuuint64_t gdim[4] = {(cuuint64_t)d->C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
        const cuuint64_t gstr[3] = {(cuuint64_t)d->pitch * 2, (cuuint64_t)d->W * d->pitch * 2,
                                    (cuuint64_t)d->H * d->W * d->pitch * 2};
        const cuuint32_t box[4] = {(cuuint32_t)CCH, (cuuint32_t)d->W, (cuuint32_t)rb, 1};
        const cuuint32_t est[4] = {1, 1, 1, 1};
        if (api.tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)xp, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCE_ERR_UNSUPPORTED;
    }
    typedef void (*KernelFn)(const CUtensorMap, const fce_pool_desc, float*, int, int);
    static const KernelFn table[2] = {coord_pool_tma_kernel<2, 5>, coord_pool_tma_kernel<1, 10>};
    static DeviceOnce attr_once;
    int dev = 0;
    if (attr_once.pending(&dev)) {
        for (int v = 0; v < 2; ++v) {
            cudaError_t e = cudaFuncSetAttribute(table[v], cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
            if (e != cudaSuccess) {
                set_cuda_error(e);
                return FCE_ERR_CUDA;
            }
        }
        attr_once.done(dev);
    }
    const size_t smem = (size_t)STAGES * rb * d->W * 128 + 8 * STAGES + 256;
    if (smem > 72 * 1024) return FCE_ERR_UNSUPPORTED;
    const int units = d->B * chunks;
    const int grid = units < 3 * kNumSMs ? units : 3 * kNumSMs;  // persistent: three CTAs per SM
    table[variant]<<<grid, PT, smem, st>>>(tm, *d, strip, chunks, units);
    return check_launch();
}

}  // namespace fce

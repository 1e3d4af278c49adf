// Coordinate pooling and the strip cross-attention core of CoordAtt / CoordCrossAtt / BiCoordCrossAtt.
//
// fce_coord_pool: ONE pass over x producing both strips (mean over W per row, mean over H per
// column), fp32 accumulation, deterministic (no atomics).  A CTA owns (image, 8-row band, 64-channel
// chunk) [32 channels in fp32 mode]; a warp reads 4 pixels x 128 contiguous bytes per request.
// HBM-bound: algorithmic bytes = C*H*W*e read (+ (H+W)*C*4 written; + the band partials, 1/8 of x in fp32... 
// i.e. 2*W*C*4 bytes per 8 rows, which is why the band is not made smaller).
//
// fce_strip_attn: softmax(q k^T * scale) v over strips; the whole per-(image, head) K/V fits in
// shared memory (L <= 160, dh <= 32 on this path), fp32 CUDA-core math - far too small for tensor cores.
#include <atomic>

#include "common.cuh"

namespace fce {
namespace {

constexpr int PT = 256;       // threads
constexpr int CVT = 8;        // channel-vector lanes per CTA
constexpr int SLOTS = 32;     // pixel slots along W
constexpr int RB = 4;         // rows per reduction group (register budget: 2 CTAs per SM)
constexpr int WS_ROWS = 8;    // the workspace contract: at most one band per 8 rows (fce_coord_pool_workspace)
constexpr int BAND = 3;       // column groups held in registers -> 96 columns per sweep

// A CTA owns (image, band of rows, 64-channel chunk [32 in fp32 mode]) and walks its band RB rows at a time.  Row
// means are complete inside the CTA; the band's column sums stay in registers across the walk and go to the
// workspace ([B][bands][W][C] fp32), where a second, tiny kernel adds the bands in a fixed order (deterministic,
// no atomics) - or straight to the strip when one band covers the image.  The host picks the band count so that
// the grid is about two CTAs per SM: at batch 64 that is ONE band (no workspace traffic at all), at batch 1 the
// bands supply the parallelism.  6 independent 16-byte loads are in flight per thread, two CTAs per SM.
template <typename T>
__global__ void __launch_bounds__(PT, 2) coord_pool_kernel(const fce_pool_desc d, const T* __restrict__ x,
                                                        float* __restrict__ strip, float* __restrict__ ws, int bands,
                                                        int band_rows) {
    constexpr int N = Vec16<T>::N;
    constexpr int CC = CVT * N;  // channels per CTA
    __shared__ float red[PT / 32][RB][CC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cvt = tid & (CVT - 1);
    const int slot = tid / CVT;  // 0..31 ; a warp holds 4 consecutive slots
    const int chunks = (d.C + CC - 1) / CC;
    const int chunk = blockIdx.x % chunks;
    const int band = (blockIdx.x / chunks) % bands;
    const int b = blockIdx.x / (chunks * bands);
    const int c0 = chunk * CC + cvt * N;
    const bool c_ok = c0 < d.C;  // C is a multiple of N (checked on host)
    const int h_begin = band * band_rows;
    const int h_end = min(d.H, h_begin + band_rows);
    const T* xb_safe = x + (size_t)b * d.H * d.W * d.pitch + d.off + (c_ok ? c0 : 0);
    float* xh = strip + ((size_t)b * d.H) * d.C;  // rows [b*H, b*H+H)
    // column partial sums of this band: straight into the strip when there is a single band
    float* colp = bands == 1 ? strip + ((size_t)d.B * d.H + (size_t)b * d.W) * d.C
                             : ws + ((size_t)(b * bands + band) * d.W) * d.C;
    const float col_scale = bands == 1 ? 1.f / (float)d.H : 1.f;
    const float inv_w = 1.f / (float)d.W;

    for (int w_base = 0; w_base < d.W; w_base += SLOTS * BAND) {
        float col[BAND][N];
#pragma unroll
        for (int g = 0; g < BAND; ++g)
#pragma unroll
            for (int j = 0; j < N; ++j) col[g][j] = 0.f;
        for (int h_base = h_begin; h_base < h_end; h_base += RB) {
            float row[RB][N];
#pragma unroll
            for (int r = 0; r < RB; ++r)
#pragma unroll
                for (int j = 0; j < N; ++j) row[r][j] = 0.f;
            // all loads of two rows (2 x BAND 16-byte vectors) are issued before any is consumed; out-of-range
            // pixels issue no load at all (W = 80 / 40 maps fill only part of the 96-column sweep)
#pragma unroll
            for (int r = 0; r < RB; r += 2) {
                Vec16<T> v[2][BAND];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int h = h_base + r + rr;
#pragma unroll
                    for (int g = 0; g < BAND; ++g) {
                        const int w = w_base + g * SLOTS + slot;
                        v[rr][g].raw = make_uint4(0u, 0u, 0u, 0u);
                        if (c_ok && h < h_end && w < d.W) v[rr][g].load_nc(xb_safe + ((size_t)h * d.W + w) * d.pitch);
                    }
                }
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                    for (int g = 0; g < BAND; ++g) {
                        float f[N];
                        v[rr][g].unpack(f);
#pragma unroll
                        for (int j = 0; j < N; ++j) {  // skipped loads contribute exact zeros
                            row[r + rr][j] += f[j];
                            col[g][j] += f[j];
                        }
                    }
            }
            // reduce row sums over the 32 slots: 4 slots inside the warp (lanes differ by 8, 16), 8 warps via smem
#pragma unroll
            for (int r = 0; r < RB; ++r)
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    float v = row[r][j];
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    row[r][j] = v;
                }
            __syncthreads();  // previous group's readers are done with red[]
            if (lane < CVT) {
#pragma unroll
                for (int r = 0; r < RB; ++r)
#pragma unroll
                    for (int j = 0; j < N; ++j) red[warp][r][cvt * N + j] = row[r][j];
            }
            __syncthreads();
            for (int o = tid; o < RB * CC; o += PT) {
                const int r = o / CC, c = o % CC;
                const int h = h_base + r;
                const int cg = chunk * CC + c;
                if (h < h_end && cg < d.C) {
                    float s = 0.f;
#pragma unroll
                    for (int wq = 0; wq < PT / 32; ++wq) s += red[wq][r][c];
                    float* dst = xh + (size_t)h * d.C + cg;
                    if (w_base > 0) s += *dst;  // later sweeps of very wide maps accumulate (same CTA, ordered)
                    if (w_base + SLOTS * BAND >= d.W) s *= inv_w;
                    *dst = s;
                }
            }
        }
        // column sums: every (slot, group) column is owned by exactly one thread
        if (c_ok) {
#pragma unroll
            for (int g = 0; g < BAND; ++g) {
                const int w = w_base + g * SLOTS + slot;
                if (w < d.W) {
                    float* dst = colp + (size_t)w * d.C + c0;
#pragma unroll
                    for (int j = 0; j < N; ++j) dst[j] = col[g][j] * col_scale;
                }
            }
        }
    }
}

// strip[B*H + b*W + w][c] = (sum over bands of ws[b][band][w][c]) / H
__global__ void __launch_bounds__(PT) coord_pool_finish(const fce_pool_desc d, const float* __restrict__ ws,
                                                        float* __restrict__ strip, int bands) {
    const size_t n = (size_t)d.B * d.W * d.C;
    const float inv_h = 1.f / (float)d.H;
    const size_t wc = (size_t)d.W * d.C;
    for (size_t i = blockIdx.x * (size_t)PT + threadIdx.x; i < n; i += (size_t)gridDim.x * PT) {
        const size_t b = i / wc, r = i - b * wc;
        const float* p = ws + (b * bands) * wc + r;
        float s = 0.f;
        for (int k = 0; k < bands; ++k) s += p[(size_t)k * wc];
        strip[(size_t)d.B * d.H * d.C + i] = s * inv_h;
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int AT = 256;

__global__ void __launch_bounds__(AT) strip_attn_kernel(const fce_strip_attn_desc d, const float* __restrict__ q,
                                                        const float* __restrict__ k, const float* __restrict__ v,
                                                        float* __restrict__ out) {
    extern __shared__ float sm[];
    const int dh = d.dh, Lk = d.Lk;
    float* Ks = sm;                    // [Lk][dh]
    float* Vs = Ks + Lk * dh;          // [Lk][dh]
    float* Ss = Vs + Lk * dh;          // [warps][Lk]
    float* Qs = Ss + (AT / 32) * Lk;   // [warps][dh]
    const int b = blockIdx.x / d.heads, head = blockIdx.x % d.heads;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* kb = k + b * d.k_bstride + head * dh;
    const float* vb = v + b * d.v_bstride + head * dh;
    for (int i = tid; i < Lk * dh; i += AT) {
        const int m = i / dh, j = i % dh;
        Ks[i] = kb[m * d.k_rstride + j];
        Vs[i] = vb[m * d.v_rstride + j];
    }
    __syncthreads();
    float* sw = Ss + warp * Lk;
    float* qw = Qs + warp * dh;
    for (int l = blockIdx.y * (AT / 32) + warp; l < d.Lq; l += gridDim.y * (AT / 32)) {
        const float* qp = q + b * d.q_bstride + l * d.q_rstride + head * dh;
        for (int j = lane; j < dh; j += 32) qw[j] = qp[j];
        __syncwarp();
        float mx = -INFINITY;
        for (int m = lane; m < Lk; m += 32) {
            float s = 0.f;
            for (int j = 0; j < dh; ++j) s = fmaf(qw[j], Ks[m * dh + j], s);
            s *= d.scale;
            sw[m] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int m = lane; m < Lk; m += 32) {
            const float e = expf(sw[m] - mx);
            sw[m] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        __syncwarp();
        const float inv = 1.f / sum;
        float* op = out + b * d.o_bstride + l * d.o_rstride + head * dh;
        for (int j = lane; j < dh; j += 32) {
            float acc = 0.f;
            for (int m = 0; m < Lk; ++m) acc = fmaf(sw[m], Vs[m * dh + j], acc);
            op[j] = acc * inv;
        }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// CoordAtt gate MLP on the pooled strips (fce_block.py:104-113): per strip row s[C],
//   y = SiLU(W1 s + b1) [mip],  a = sigmoid(W2 y + b2) [oup],  W2/b2 = (w_h, b_h) for the B*H rows pooled over W and
//   (w_w, b_w) for the B*W rows pooled over H.
// One launch instead of three strip GEMMs (cv1, cv_h, cv_w) whose fixed costs dominated their ~2 us of work.  A CTA
// owns MR consecutive rows of one of the two row ranges: the rows sit in shared memory, (row, m) pairs are dealt to
// threads for layer 1, then thread o keeps MR accumulators for output channel o in layer 2.  Weights arrive
// transposed ([C][mip] and [mip][oup]) so that a warp's loads are consecutive; they stay L1/L2 resident.
constexpr int MT = 256;  // threads
constexpr int MR = 16;   // strip rows per CTA

__global__ void __launch_bounds__(MT) coordatt_mlp_kernel(const fce_coordatt_mlp_desc d, const float* __restrict__ strip,
                                                          const float* __restrict__ w1t, const float* __restrict__ b1,
                                                          const float* __restrict__ wht, const float* __restrict__ bh,
                                                          const float* __restrict__ wwt, const float* __restrict__ bw,
                                                          float* __restrict__ out) {
    extern __shared__ float msm[];
    const int C = d.C, mip = d.mip, oup = d.oup;
    const int sp = C + 4;          // padded row pitch (keeps 16-byte alignment, staggers the banks of adjacent rows)
    float* ss = msm;               // [MR][sp]
    float* ys = msm + MR * sp;     // [MR][mip]
    const int nh = (d.rows_h + MR - 1) / MR;
    const bool is_w = (int)blockIdx.x >= nh;
    const int r0 = is_w ? d.rows_h + ((int)blockIdx.x - nh) * MR : (int)blockIdx.x * MR;
    const int r_end = is_w ? d.rows_h + d.rows_w : d.rows_h;
    const int nr = min(MR, r_end - r0);
    const int tid = threadIdx.x;
    pdl_trigger();
    for (int i = tid; i < MR * (C >> 2); i += MT) {
        const int r = i / (C >> 2), c4 = i - r * (C >> 2);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nr) v = *reinterpret_cast<const float4*>(strip + (size_t)(r0 + r) * d.s_pitch + 4 * c4);
        *reinterpret_cast<float4*>(ss + r * sp + 4 * c4) = v;
    }
    __syncthreads();
    for (int i = tid; i < MR * mip; i += MT) {
        const int r = i / mip, m = i - r * mip;
        const float* sr = ss + r * sp;
        const float* wp = w1t + m;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int c = 0; c < C; c += 4) {
            const float4 s4 = *reinterpret_cast<const float4*>(sr + c);
            a0 = fmaf(s4.x, __ldg(wp + (size_t)c * mip), a0);
            a1 = fmaf(s4.y, __ldg(wp + (size_t)(c + 1) * mip), a1);
            a2 = fmaf(s4.z, __ldg(wp + (size_t)(c + 2) * mip), a2);
            a3 = fmaf(s4.w, __ldg(wp + (size_t)(c + 3) * mip), a3);
        }
        ys[i] = apply_act((a0 + a1) + (a2 + a3) + __ldg(b1 + m), d.act1);
    }
    __syncthreads();
    const float* w2t = is_w ? wwt : wht;
    const float* b2 = is_w ? bw : bh;
    for (int o = tid; o < oup; o += MT) {
        float acc[MR];
        const float bo = __ldg(b2 + o);
#pragma unroll
        for (int r = 0; r < MR; ++r) acc[r] = bo;
        for (int m = 0; m < mip; ++m) {
            const float wv = __ldg(w2t + (size_t)m * oup + o);
#pragma unroll
            for (int r = 0; r < MR; ++r) acc[r] = fmaf(wv, ys[r * mip + m], acc[r]);
        }
#pragma unroll
        for (int r = 0; r < MR; ++r)
            if (r < nr) out[(size_t)(r0 + r) * d.out_pitch + o] = apply_act(acc[r], d.act2);
    }
}

}  // namespace
}  // namespace fce

using namespace fce;

extern "C" size_t fce_coord_pool_workspace(const fce_pool_desc* d) {
    if (!d || d->H <= WS_ROWS) return 0;
    const size_t bands = (size_t)(d->H + WS_ROWS - 1) / WS_ROWS;
    return (size_t)d->B * bands * d->W * d->C * sizeof(float);
}

extern "C" int fce_coord_pool(const fce_pool_desc* d, const void* x, float* strip, void* ws, size_t ws_bytes,
                              void* stream) {
    if (!d || !x || !strip || d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = d->dtype == FCE_BF16 ? 8 : 4;
    if (d->dtype != FCE_BF16 && d->dtype != FCE_F32) return FCE_ERR_UNSUPPORTED;
    if ((d->C % n) || (d->pitch % n) || (d->off % n) || (((uintptr_t)x) & 15)) return FCE_ERR_ALIGNMENT;
    const int cc = CVT * n;
    const int chunks = (d->C + cc - 1) / cc;
    // bands: enough CTAs for ~4 per SM, never more than one band per WS_ROWS rows
    const int max_bands = (d->H + WS_ROWS - 1) / WS_ROWS;
    int bands = (4 * kNumSMs + d->B * chunks - 1) / (d->B * chunks);
    bands = bands < 1 ? 1 : (bands > max_bands ? max_bands : bands);
    int band_rows = ((d->H + bands - 1) / bands + WS_ROWS - 1) / WS_ROWS * WS_ROWS;
    bands = (d->H + band_rows - 1) / band_rows;
    if (bands > 1 && (!ws || ws_bytes < (size_t)d->B * bands * d->W * d->C * sizeof(float))) return FCE_ERR_WORKSPACE;
    const int grid = d->B * bands * chunks;
    if (d->dtype == FCE_BF16)
        coord_pool_kernel<__nv_bfloat16><<<grid, PT, 0, st>>>(*d, (const __nv_bfloat16*)x, strip, (float*)ws, bands, band_rows);
    else
        coord_pool_kernel<float><<<grid, PT, 0, st>>>(*d, (const float*)x, strip, (float*)ws, bands, band_rows);
    int rc = check_launch();
    if (rc != FCE_OK || bands == 1) return rc;
    const size_t nitems = (size_t)d->B * d->W * d->C;
    int fgrid = (int)((nitems + PT - 1) / PT);
    if (fgrid > kNumSMs * 8) fgrid = kNumSMs * 8;
    coord_pool_finish<<<fgrid, PT, 0, st>>>(*d, (const float*)ws, strip, bands);
    return check_launch();
}

extern "C" int fce_strip_attn(const fce_strip_attn_desc* d, const float* q, const float* k, const float* v,
                              float* out, void* stream) {
    if (!d || !q || !k || !v || !out || d->B <= 0 || d->heads <= 0 || d->dh <= 0 || d->Lq <= 0 || d->Lk <= 0)
        return FCE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * ((size_t)2 * d->Lk * d->dh + (AT / 32) * (size_t)d->Lk + (AT / 32) * (size_t)d->dh);
    if (smem > 200 * 1024) return FCE_ERR_UNSUPPORTED;
    static std::atomic<bool> attr_done{false};
    if (!attr_done.load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(strip_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) { set_cuda_error(e); return FCE_ERR_CUDA; }
        attr_done.store(true, std::memory_order_release);
    }
    int qblocks = (d->Lq + (AT / 32) - 1) / (AT / 32);
    // enough CTAs to cover the SMs, without re-loading K/V more often than needed
    int want = (2 * kNumSMs + d->B * d->heads - 1) / (d->B * d->heads);
    if (qblocks > want) qblocks = want;
    if (qblocks < 1) qblocks = 1;
    dim3 grid(d->B * d->heads, qblocks);
    strip_attn_kernel<<<grid, AT, smem, st>>>(*d, q, k, v, out);
    return check_launch();
}

extern "C" int fce_coordatt_mlp(const fce_coordatt_mlp_desc* d, const float* strip, const float* w1t, const float* b1,
                                const float* wht, const float* bh, const float* wwt, const float* bw, float* out,
                                void* stream) {
    if (!d || !strip || !w1t || !b1 || !wht || !bh || !wwt || !bw || !out) return FCE_ERR_BAD_ARG;
    if (d->rows_h < 0 || d->rows_w < 0 || d->rows_h + d->rows_w <= 0 || d->C <= 0 || d->mip <= 0 || d->oup <= 0)
        return FCE_ERR_BAD_ARG;
    if ((d->C & 3) || (d->s_pitch & 3) || (((uintptr_t)strip) & 15)) return FCE_ERR_ALIGNMENT;
    const size_t smem = sizeof(float) * ((size_t)MR * (d->C + 4) + (size_t)MR * d->mip);
    if (smem > 160 * 1024) return FCE_ERR_UNSUPPORTED;
    static std::atomic<bool> attr_done{false};
    if (!attr_done.load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(coordatt_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) { set_cuda_error(e); return FCE_ERR_CUDA; }
        attr_done.store(true, std::memory_order_release);
    }
    const int grid = (d->rows_h + MR - 1) / MR + (d->rows_w + MR - 1) / MR;
    coordatt_mlp_kernel<<<grid, MT, smem, (cudaStream_t)stream>>>(*d, strip, w1t, b1, wht, bh, wwt, bw, out);
    return check_launch();
}

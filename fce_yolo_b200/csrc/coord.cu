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
// CVT = channel-vector lanes per CTA (template parameter: 8 / 16 / 32), SLOTS = PT / CVT pixel slots along W: the
// host picks the widest CVT whose 3 * SLOTS columns per sweep still cover W, so 80 / 40 / 20-pixel maps all keep
// 83 % of the load slots busy (with CVT fixed at 8 a 40-pixel map used 42 % of them: 1.6 TB/s)
constexpr int RB = 4;         // rows per reduction group (register budget: 2 CTAs per SM)
constexpr int WS_ROWS = 8;    // the workspace contract: at most one band per 8 rows (fce_coord_pool_workspace)
constexpr int BAND = 3;       // column groups held in registers -> 96 columns per sweep

// A CTA owns (image, band of rows, 64-channel chunk [32 in fp32 mode]) and walks its band RB rows at a time.  Row
// means are complete inside the CTA; the band's column sums stay in registers across the walk and go to the
// workspace ([B][bands][W][C] fp32), where a second, tiny kernel adds the bands in a fixed order (deterministic,
// no atomics) - or straight to the strip when one band covers the image.  The host picks the band count so that
// the grid is about two CTAs per SM: at batch 64 that is ONE band (no workspace traffic at all), at batch 1 the
// bands supply the parallelism.  6 independent 16-byte loads are in flight per thread, two CTAs per SM.
template <typename T, int CVT>
__global__ void __launch_bounds__(PT, 2) coord_pool_kernel(const fce_pool_desc d, const T* __restrict__ x,
                                                        float* __restrict__ strip, float* __restrict__ ws, int bands,
                                                        int band_rows) {
    constexpr int N = Vec16<T>::N;
    constexpr int CC = CVT * N;  // channels per CTA
    constexpr int SLOTS = PT / CVT;
    __shared__ float red[PT / 32][RB][CC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cvt = tid & (CVT - 1);
    const int slot = tid / CVT;  // a warp holds 32 / CVT consecutive slots
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
            // reduce row sums over the slots: the 32 / CVT slots inside the warp by shuffles, the 8 warps via smem
#pragma unroll
            for (int r = 0; r < RB; ++r)
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    float v = row[r][j];
                    if (CVT <= 8) v += __shfl_xor_sync(0xffffffffu, v, 8);
                    if (CVT <= 16) v += __shfl_xor_sync(0xffffffffu, v, 16);
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
// One launch instead of three strip GEMMs (cv1, cv_h, cv_w) whose fixed costs dominated their ~2 us of work.
// 84 M MACs on 10 K rows: too thin for the tensor cores in fp32, so the kernel is built to issue few instructions
// per MAC (two earlier versions - weights through L1, then a warp reduction per hidden unit - both took 35 us):
//   * persistent CTAs, blockIdx.y picks the row range (and with it W2); both weight matrices are parked in shared
//     memory once per CTA; W1 arrives regrouped as [C/4][mip][4] so that lanes read consecutive hidden units;
//   * a warp owns a PAIR of rows; layer 1: lane (half, m) computes the whole dot product of row `half` with hidden
//     unit m - no cross-lane reduction; the row is a shared-memory broadcast, the math packed fp32 pairs (FFMA2);
//   * layer 2: lane owns float4 groups of output channels for both rows, so every W2^T read feeds eight MACs;
//     conflict-free shared-memory reads, 16-byte coalesced stores.
constexpr int MT = 256;    // threads
constexpr int MIP_MAX = 64;

__device__ __forceinline__ float mlp_act(float v, int act) {  // fast forms: ~1e-6 relative, far inside fp32-mode parity
    if (act == FCE_ACT_SILU) return silu_f(v);
    if (act == FCE_ACT_SIGMOID) return sigmoid_f(v);
    return v;
}

__global__ void __launch_bounds__(MT) coordatt_mlp_kernel(const fce_coordatt_mlp_desc d, const float* __restrict__ strip,
                                                          const float* __restrict__ w1q_g, const float* __restrict__ b1,
                                                          const float* __restrict__ wht, const float* __restrict__ bh,
                                                          const float* __restrict__ wwt, const float* __restrict__ bw,
                                                          float* __restrict__ out) {
    extern __shared__ float4 msm4[];
    const int C4 = d.C >> 2, O4 = d.oup >> 2, mip = d.mip;
    float4* w1q = msm4;                      // [C4][mip]: w1q[c4 * mip + m] = W1[m][4 c4 .. 4 c4 + 3]
    float4* w2s = w1q + mip * C4;            // [mip][O4]  (W2 transposed)
    float4* b2s = w2s + mip * O4;            // [O4]
    float4* srow = b2s + O4;                 // [warps][2][C4]
    float* b1s = reinterpret_cast<float*>(srow + (MT / 32) * 2 * C4);  // [MIP_MAX]
    float2* ysm = reinterpret_cast<float2*>(b1s + MIP_MAX);            // [warps][MIP_MAX] (row A, row B)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_w = blockIdx.y == 1;
    const int seg_rows = is_w ? d.rows_w : d.rows_h;
    const int row0 = is_w ? d.rows_h : 0;
    pdl_trigger();
    if ((int)blockIdx.x * (MT / 16) >= seg_rows) return;  // this CTA has no rows (uniform)
    {
        const float4* g1 = reinterpret_cast<const float4*>(w1q_g);
        const float4* g2 = reinterpret_cast<const float4*>(is_w ? wwt : wht);
        const float4* gb = reinterpret_cast<const float4*>(is_w ? bw : bh);
        for (int i = tid; i < mip * C4; i += MT) w1q[i] = __ldg(g1 + i);  // already regrouped by the caller
        for (int i = tid; i < mip * O4; i += MT) w2s[i] = __ldg(g2 + i);
        for (int i = tid; i < O4; i += MT) b2s[i] = __ldg(gb + i);
        for (int i = tid; i < mip; i += MT) b1s[i] = __ldg(b1 + i);
    }
    __syncthreads();
    const int half = lane >> 4, ml = lane & 15;
    float4* sw = srow + (warp * 2 + half) * C4;
    float2* yw = ysm + warp * MIP_MAX;
    const int pstep = gridDim.x * (MT / 32);
    const int pairs = (seg_rows + 1) >> 1;
    for (int pr = blockIdx.x * (MT / 32) + warp; pr < pairs; pr += pstep) {
        const int rA = 2 * pr, rB = min(2 * pr + 1, seg_rows - 1);  // an odd tail row is computed twice, stored once
        const size_t my_row = (size_t)(row0 + (half ? rB : rA));
        {
            const float4* sp = reinterpret_cast<const float4*>(strip + my_row * d.s_pitch);
            for (int i = ml; i < C4; i += 16) sw[i] = sp[i];
        }
        __syncwarp();
        for (int m = ml; m < mip; m += 16) {
            const float4* wq = w1q + m;
            float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
            int c4 = 0;
            for (; c4 + 2 <= C4; c4 += 2) {
                const float4 s0 = sw[c4], s1 = sw[c4 + 1];
                const float4 u0 = wq[c4 * mip], u1 = wq[(c4 + 1) * mip];
                a0 = __ffma2_rn(make_float2(s0.x, s0.y), make_float2(u0.x, u0.y), a0);
                a1 = __ffma2_rn(make_float2(s0.z, s0.w), make_float2(u0.z, u0.w), a1);
                a2 = __ffma2_rn(make_float2(s1.x, s1.y), make_float2(u1.x, u1.y), a2);
                a3 = __ffma2_rn(make_float2(s1.z, s1.w), make_float2(u1.z, u1.w), a3);
            }
            if (c4 < C4) {
                const float4 s0 = sw[c4];
                const float4 u0 = wq[c4 * mip];
                a0 = __ffma2_rn(make_float2(s0.x, s0.y), make_float2(u0.x, u0.y), a0);
                a1 = __ffma2_rn(make_float2(s0.z, s0.w), make_float2(u0.z, u0.w), a1);
            }
            const float2 t = __fadd2_rn(__fadd2_rn(a0, a1), __fadd2_rn(a2, a3));
            const float yv = mlp_act(t.x + t.y + b1s[m], d.act1);
            if (half) yw[m].y = yv;
            else yw[m].x = yv;
        }
        __syncwarp();
        float4* opA = reinterpret_cast<float4*>(out + (size_t)(row0 + rA) * d.out_pitch);
        float4* opB = reinterpret_cast<float4*>(out + (size_t)(row0 + rB) * d.out_pitch);
        for (int o4 = lane; o4 < O4; o4 += 32) {
            const float4 bb = b2s[o4];
            float2 pA0 = make_float2(bb.x, bb.y), pA1 = make_float2(bb.z, bb.w), pB0 = pA0, pB1 = pA1;
#pragma unroll 4
            for (int m = 0; m < mip; ++m) {
                const float2 yy = yw[m];
                const float4 w = w2s[m * O4 + o4];
                const float2 ya = make_float2(yy.x, yy.x), yb = make_float2(yy.y, yy.y);
                pA0 = __ffma2_rn(ya, make_float2(w.x, w.y), pA0);
                pA1 = __ffma2_rn(ya, make_float2(w.z, w.w), pA1);
                pB0 = __ffma2_rn(yb, make_float2(w.x, w.y), pB0);
                pB1 = __ffma2_rn(yb, make_float2(w.z, w.w), pB1);
            }
            opA[o4] = make_float4(mlp_act(pA0.x, d.act2), mlp_act(pA0.y, d.act2), mlp_act(pA1.x, d.act2),
                                  mlp_act(pA1.y, d.act2));
            if (rB != rA)
                opB[o4] = make_float4(mlp_act(pB0.x, d.act2), mlp_act(pB0.y, d.act2), mlp_act(pB1.x, d.act2),
                                      mlp_act(pB1.y, d.act2));
        }
        __syncwarp();  // sw / yw are rewritten by the next pair
    }
}

}  // namespace
}  // namespace fce

namespace fce {
int coord_pool_tma(const fce_pool_desc* d, const void* x, float* strip, cudaStream_t st);  // coord_pool_tma.cu
}

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
    {   // bf16, enough (image, 64-channel chunk) pairs to fill the GPU: the TMA-fed kernel (asynchronous loads)
        const int rc = coord_pool_tma(d, x, strip, st);
        if (rc != FCE_ERR_UNSUPPORTED) return rc;
    }
    // widest channel group whose sweep (3 column groups of PT / cvt slots) still covers W
    const int cvt = (d->W <= 24 && d->C >= 32 * n) ? 32 : ((d->W <= 48 && d->C >= 16 * n) ? 16 : 8);
    const int cc = cvt * n;
    const int chunks = (d->C + cc - 1) / cc;
    // bands: with (image, chunk) pairs for at least three quarters of the SMs a single band is best - no workspace traffic, no finish
    // kernel, and splitting the pairs further only re-slices the same two-CTAs-per-SM rounds; small batches get their
    // parallelism from the bands (never more than one per WS_ROWS rows)
    const int max_bands = (d->H + WS_ROWS - 1) / WS_ROWS;
    int bands = d->B * chunks >= (kNumSMs * 3) / 4 ? 1 : (4 * kNumSMs + d->B * chunks - 1) / (d->B * chunks);
    bands = bands < 1 ? 1 : (bands > max_bands ? max_bands : bands);
    int band_rows = ((d->H + bands - 1) / bands + WS_ROWS - 1) / WS_ROWS * WS_ROWS;
    bands = (d->H + band_rows - 1) / band_rows;
    if (bands > 1 && (!ws || ws_bytes < (size_t)d->B * bands * d->W * d->C * sizeof(float))) return FCE_ERR_WORKSPACE;
    const int grid = d->B * bands * chunks;
#define FCE_POOL_LAUNCH(T, CV) \
    coord_pool_kernel<T, CV><<<grid, PT, 0, st>>>(*d, (const T*)x, strip, (float*)ws, bands, band_rows)
    if (d->dtype == FCE_BF16) {
        if (cvt == 32) FCE_POOL_LAUNCH(__nv_bfloat16, 32);
        else if (cvt == 16) FCE_POOL_LAUNCH(__nv_bfloat16, 16);
        else FCE_POOL_LAUNCH(__nv_bfloat16, 8);
    } else {
        if (cvt == 32) FCE_POOL_LAUNCH(float, 32);
        else if (cvt == 16) FCE_POOL_LAUNCH(float, 16);
        else FCE_POOL_LAUNCH(float, 8);
    }
#undef FCE_POOL_LAUNCH
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
    static DeviceOnce attr_once;  // per-device attribute
    int dev_ = 0;
    if (attr_once.pending(&dev_)) {
        cudaError_t e = cudaFuncSetAttribute(strip_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) { set_cuda_error(e); return FCE_ERR_CUDA; }
        attr_once.done(dev_);
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

extern "C" int fce_coordatt_mlp(const fce_coordatt_mlp_desc* d, const float* strip, const float* w1, const float* b1,
                                const float* wht, const float* bh, const float* wwt, const float* bw, float* out,
                                void* stream) {
    if (!d || !strip || !w1 || !b1 || !wht || !bh || !wwt || !bw || !out) return FCE_ERR_BAD_ARG;
    if (d->rows_h < 0 || d->rows_w < 0 || d->rows_h + d->rows_w <= 0 || d->C <= 0 || d->mip <= 0 || d->oup <= 0)
        return FCE_ERR_BAD_ARG;
    if ((d->C & 3) || (d->oup & 3) || (d->s_pitch & 3) || (d->out_pitch & 3)) return FCE_ERR_ALIGNMENT;
    for (const void* p : {(const void*)strip, (const void*)w1, (const void*)wht, (const void*)bh, (const void*)wwt,
                          (const void*)bw, (const void*)out})
        if (((uintptr_t)p) & 15) return FCE_ERR_ALIGNMENT;
    if (d->mip > MIP_MAX) return FCE_ERR_UNSUPPORTED;
    const size_t smem = sizeof(float) * ((size_t)d->mip * (d->C + d->oup) + d->oup + (MT / 32) * 2 * (size_t)d->C +
                                         MIP_MAX + (MT / 32) * 2 * MIP_MAX);
    if (smem > 200 * 1024) return FCE_ERR_UNSUPPORTED;
    static DeviceOnce attr_once;  // per-device attribute
    int dev_ = 0;
    if (attr_once.pending(&dev_)) {
        cudaError_t e = cudaFuncSetAttribute(coordatt_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) { set_cuda_error(e); return FCE_ERR_CUDA; }
        attr_once.done(dev_);
    }
    // one CTA per SM and row range: the weights are read once per CTA, row pairs are dealt to warps round-robin
    const int seg = d->rows_h > d->rows_w ? d->rows_h : d->rows_w;
    int gx = (seg + MT / 16 - 1) / (MT / 16);
    if (gx > kNumSMs) gx = kNumSMs;
    coordatt_mlp_kernel<<<dim3(gx, 2), MT, smem, (cudaStream_t)stream>>>(*d, strip, w1, b1, wht, bh, wwt, bw, out);
    return check_launch();
}

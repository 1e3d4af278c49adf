// Bandwidth-bound NHWC kernels: depthwise 3x3, SPPF max-pool pyramid, nearest 2x upsample,
// BiFPN weighted fusion, gate application, view copy.  No tensor cores: every thread owns one
// 16-byte channel vector (8 bf16 / 4 fp32) of one pixel, so a warp touches whole 128-byte lines;
// one CTA per image row, no integer division on the device.
#include <cstdlib>

#include "common.cuh"

namespace fce {
namespace {

constexpr int NT = 256;

// Thread-block shape shared by the pointwise kernels: blockDim.x walks the 16-byte channel vectors of one
// pixel (coalesced 128-byte lines), blockDim.y walks pixels of one image row; blockIdx.x is the image row
// (b*H + h).  No integer division anywhere on the device.
struct RowGrid {
    dim3 grid, block;
};
inline RowGrid row_grid(int rows, int cv) {
    int cx = 1;
    while (cx < cv && cx < 32) cx <<= 1;
    RowGrid g;
    g.block = dim3(cx, NT / cx);
    g.grid = dim3(rows);
    return g;
}

template <typename T, bool VEC>
struct CV {  // channel-vector accessor: VEC -> 16-byte vectors, else single elements
    static constexpr int N = VEC ? Vec16<T>::N : 1;
    static __device__ __forceinline__ void load(const T* p, float* f) {
        if constexpr (VEC) {
            Vec16<T> v;
            v.load(p);
            v.unpack(f);
        } else {
            f[0] = Elem<T>::to_f(*p);
        }
    }
    static __device__ __forceinline__ void store(T* p, const float* f) {
        if constexpr (VEC) {
            Vec16<T> v;
            v.pack(f);
            v.store(p);
        } else {
            *p = Elem<T>::from_f(f[0]);
        }
    }
};

template <typename T>
__device__ __forceinline__ float silu_for(float v) {
    if constexpr (sizeof(T) == 2) return silu_f(v);  // bf16 storage: fast exp is far below the output rounding
    else return silu_acc(v);
}

// ------------------------------------------------------------------------------------------------
// Depthwise 3x3, sliding window.  A warp owns one image row and 32 lanes x 4 bytes of channels (64 bf16 / 32
// fp32): every load and store is one fully coalesced 128-byte line.  The warp walks the row left to right in
// chunks of 4 pixels keeping the 3 x 6 input window in registers, so each input element is loaded ONCE per
// output row (3 loads per output instead of 9) and unpacked once; the 12 loads of the next chunk are issued
// before the current chunk is computed.  Padding: out-of-range rows get zero weights, out-of-range columns
// zero values; addresses are clamped so that every load is in bounds.
template <typename T>
struct Lane4;  // 4 bytes of channels per lane
template <>
struct Lane4<__nv_bfloat16> {
    static constexpr int E = 2;
    static __device__ __forceinline__ void unpack(uint32_t r, float* f) {
        f[0] = __uint_as_float(r << 16);
        f[1] = __uint_as_float(r & 0xffff0000u);
    }
    static __device__ __forceinline__ uint32_t pack(const float* f) {
        __nv_bfloat162 h = __floats2bfloat162_rn(f[0], f[1]);
        return *reinterpret_cast<uint32_t*>(&h);
    }
};
template <>
struct Lane4<float> {
    static constexpr int E = 1;
    static __device__ __forceinline__ void unpack(uint32_t r, float* f) { f[0] = __uint_as_float(r); }
    static __device__ __forceinline__ uint32_t pack(const float* f) { return __float_as_uint(f[0]); }
};

// ------------------------------------------------------------------------------------------------
// Depthwise 3x3, version 2 (used whenever the views are 16-byte aligned): a thread owns ONE 16-byte channel
// vector (8 bf16 / 4 fp32) of one image column and walks DOWN a band of rows.  Input row r is loaded once per
// thread as three vectors (columns w-1, w, w+1: the two neighbours are L1 hits - the adjacent threads of the
// same CTA load them as their own centre) and scattered into three rolling accumulators (output rows r-1, r,
// r+1), so each output costs one HBM read of its input element (x (TH+2)/TH for the band halo), 9 FMAs per
// channel and one 16-byte store.  The next row's vectors are in flight while the current row is computed.
// CTA = 128 threads = 16 columns x 8 vectors (64 bf16 / 32 fp32 channels: 128 contiguous bytes per pixel).
constexpr int DW2_COLS = 16, DW2_VECS = 8, DW2_THREADS = DW2_COLS * DW2_VECS;

template <typename T>
__global__ void __launch_bounds__(DW2_THREADS) dwconv3x3_v2_kernel(const fce_dwconv_desc d, const T* __restrict__ x,
                                                                   const float* __restrict__ w,
                                                                   const float* __restrict__ bias,
                                                                   const T* __restrict__ add, T* __restrict__ y,
                                                                   int band_rows, int bands, int cgroups) {
    pdl_trigger();
    constexpr int N = Vec16<T>::N;
    const int vec = threadIdx.x % DW2_VECS, col = threadIdx.x / DW2_VECS;
    const int cg = blockIdx.x % cgroups;
    const int wt_ = blockIdx.x / cgroups;
    const int band = blockIdx.y % bands, b = blockIdx.y / bands;
    const int c = (cg * DW2_VECS + vec) * N;
    const int wc = wt_ * DW2_COLS + col;
    if (c >= d.C || wc >= d.W) return;
    const int h0 = band * band_rows, h1 = min(d.H, h0 + band_rows);
    float wt[9][N], bs[N];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < N; ++j) wt[t][j] = __ldg(w + t * d.C + c + j);
#pragma unroll
    for (int j = 0; j < N; ++j) bs[j] = __ldg(bias + c + j);
    const bool l_ok = wc > 0, r_ok = wc + 1 < d.W;
    const T* xp = x + ((size_t)b * d.H * d.W + wc) * d.in_pitch + d.in_off + c;  // row 0 of this column
    const size_t row_stride = (size_t)d.W * d.in_pitch;

    Vec16<T> nl, nm, nr;  // next input row, raw
    auto load_row = [&](int r) {
        nl.raw = nm.raw = nr.raw = make_uint4(0u, 0u, 0u, 0u);
        if (r >= 0 && r < d.H) {
            const T* p = xp + (size_t)r * row_stride;
            nm.load(p);
            if (l_ok) nl.load(p - d.in_pitch);
            if (r_ok) nr.load(p + d.in_pitch);
        }
    };
    float a0[N], a1[N], a2[N];  // accumulators of output rows r-1, r, r+1 while input row r is processed
#pragma unroll
    for (int j = 0; j < N; ++j) a0[j] = a1[j] = a2[j] = bs[j];
    load_row(h0 - 1);
    for (int r = h0 - 1; r <= h1; ++r) {
        float fl[N], fm[N], fr[N];
        nl.unpack(fl);
        nm.unpack(fm);
        nr.unpack(fr);
        if (r < h1) load_row(r + 1);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            // input row r is tap row kh = 2 of output r-1, kh = 1 of output r, kh = 0 of output r+1
            a0[j] = fmaf(fl[j], wt[6][j], fmaf(fm[j], wt[7][j], fmaf(fr[j], wt[8][j], a0[j])));
            a1[j] = fmaf(fl[j], wt[3][j], fmaf(fm[j], wt[4][j], fmaf(fr[j], wt[5][j], a1[j])));
            a2[j] = fmaf(fl[j], wt[0][j], fmaf(fm[j], wt[1][j], fmaf(fr[j], wt[2][j], a2[j])));
        }
        const int o = r - 1;  // complete now
        if (o >= h0) {
            float out[N];
#pragma unroll
            for (int j = 0; j < N; ++j) out[j] = d.act == FCE_ACT_SILU ? silu_for<T>(a0[j]) : a0[j];
            const size_t pix = ((size_t)b * d.H + o) * d.W + wc;
            if (add) {
                Vec16<T> av;
                float af[N];
                av.load(add + pix * d.add_pitch + d.add_off + c);
                av.unpack(af);
#pragma unroll
                for (int j = 0; j < N; ++j) out[j] += af[j];
            }
            Vec16<T> ov;
            ov.pack(out);
            ov.store(y + pix * d.out_pitch + d.out_off + c);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            a0[j] = a1[j];
            a1[j] = a2[j];
            a2[j] = bs[j];
        }
    }
}

constexpr int DW_WARPS = 4;  // warps per CTA, each an independent (row, channel group)
constexpr int DW_CHUNK = 4;  // pixels per register chunk

template <typename T>
__global__ void __launch_bounds__(DW_WARPS * 32) dwconv3x3_kernel(const fce_dwconv_desc d, const T* __restrict__ x,
                                                                  const float* __restrict__ w,
                                                                  const float* __restrict__ bias,
                                                                  const T* __restrict__ add, T* __restrict__ y,
                                                                  int groups, int total_units) {
    constexpr int E = Lane4<T>::E;
    const int lane = threadIdx.x & 31;
    const int unit = blockIdx.x * DW_WARPS + (threadIdx.x >> 5);  // (row, channel group)
    if (unit >= total_units) return;
    const int row = unit / groups, cg = unit - row * groups;  // row = b*H + h
    const int ph = row % d.H;
    int c = (cg * 32 + lane) * E;
    const bool c_ok = c < d.C;
    if (!c_ok) c = 0;  // idle lanes shadow channel 0 (loads stay in bounds), stores are predicated
    const bool up_ok = ph > 0, dn_ok = ph < d.H - 1;
    const uint32_t* r1 = reinterpret_cast<const uint32_t*>(x + (size_t)row * d.W * d.in_pitch + d.in_off + c);
    const size_t pitch4 = (size_t)d.in_pitch * sizeof(T) / 4;  // pixel pitch in 4-byte units
    const uint32_t* r0 = up_ok ? r1 - (size_t)d.W * pitch4 : r1;
    const uint32_t* r2 = dn_ok ? r1 + (size_t)d.W * pitch4 : r1;
    float wt[9][E], bs[E];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const bool rok = t < 3 ? up_ok : (t < 6 ? true : dn_ok);
#pragma unroll
        for (int j = 0; j < E; ++j) wt[t][j] = rok ? w[t * d.C + c + j] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < E; ++j) bs[j] = bias[c + j];
    const uint32_t* ap = add ? reinterpret_cast<const uint32_t*>(add + (size_t)row * d.W * d.add_pitch + d.add_off + c) : nullptr;
    const size_t apitch4 = (size_t)d.add_pitch * sizeof(T) / 4;
    uint32_t* yp = reinterpret_cast<uint32_t*>(y + (size_t)row * d.W * d.out_pitch + d.out_off + c);
    const size_t opitch4 = (size_t)d.out_pitch * sizeof(T) / 4;
    const int W = d.W;

    // window columns: win[k][col] for col = chunk_start-1 .. chunk_start+CHUNK  (CHUNK + 2 columns)
    float win[3][DW_CHUNK + 2][E];
    uint32_t nxt[3][DW_CHUNK], nadd[DW_CHUNK];
    auto load_cols = [&](int w0) {  // raw loads of columns w0 .. w0+CHUNK-1 (clamped), + the add operand of the
#pragma unroll                     // chunk starting at w0 - 1
        for (int q = 0; q < DW_CHUNK; ++q) {
            const int wc = w0 + q < W ? w0 + q : W - 1;
            nxt[0][q] = __ldg(r0 + (size_t)wc * pitch4);
            nxt[1][q] = __ldg(r1 + (size_t)wc * pitch4);
            nxt[2][q] = __ldg(r2 + (size_t)wc * pitch4);
        }
    };
    // prologue: window = [pad, col0 .. col CHUNK]: load columns 0..CHUNK-1 into win[.][1..CHUNK], then column
    // CHUNK .. 2*CHUNK-1 into nxt (its first column completes the window)
    load_cols(0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int j = 0; j < E; ++j) win[k][0][j] = 0.f;
#pragma unroll
        for (int q = 0; q < DW_CHUNK; ++q) {
            Lane4<T>::unpack(nxt[k][q], win[k][q + 1]);
            if (q >= W) {
#pragma unroll
                for (int j = 0; j < E; ++j) win[k][q + 1][j] = 0.f;
            }
        }
    }
    for (int w0 = 0; w0 < W; w0 += DW_CHUNK) {
        load_cols(w0 + DW_CHUNK);  // columns of the NEXT chunk; in flight while this chunk is computed
        if (ap) {
#pragma unroll
            for (int q = 0; q < DW_CHUNK; ++q) {
                const int wc = w0 + q < W ? w0 + q : W - 1;
                nadd[q] = __ldg(ap + (size_t)wc * apitch4);
            }
        }
        // the window's last column is the first column of the next chunk
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            Lane4<T>::unpack(nxt[k][0], win[k][DW_CHUNK + 1]);
            if (w0 + DW_CHUNK >= W) {
#pragma unroll
                for (int j = 0; j < E; ++j) win[k][DW_CHUNK + 1][j] = 0.f;
            }
        }
#pragma unroll
        for (int q = 0; q < DW_CHUNK; ++q) {
            if (w0 + q >= W) break;
            float acc[E];
#pragma unroll
            for (int j = 0; j < E; ++j) {
                float a = bs[j];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    a = fmaf(win[k][q][j], wt[k * 3 + 0][j], a);
                    a = fmaf(win[k][q + 1][j], wt[k * 3 + 1][j], a);
                    a = fmaf(win[k][q + 2][j], wt[k * 3 + 2][j], a);
                }
                acc[j] = a;
            }
            if (d.act == FCE_ACT_SILU) {
#pragma unroll
                for (int j = 0; j < E; ++j) acc[j] = silu_for<T>(acc[j]);
            }
            if (ap) {
                float av[E];
                Lane4<T>::unpack(nadd[q], av);
#pragma unroll
                for (int j = 0; j < E; ++j) acc[j] += av[j];
            }
            if (c_ok) yp[(size_t)(w0 + q) * opitch4] = Lane4<T>::pack(acc);
        }
        // slide: the last two columns of this window open the next one, the rest comes from nxt
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int j = 0; j < E; ++j) {
                win[k][0][j] = win[k][DW_CHUNK][j];
                win[k][1][j] = win[k][DW_CHUNK + 1][j];
            }
#pragma unroll
            for (int q = 1; q < DW_CHUNK; ++q) {
                Lane4<T>::unpack(nxt[k][q], win[k][q + 1]);
                if (w0 + DW_CHUNK + q >= W) {
#pragma unroll
                    for (int j = 0; j < E; ++j) win[k][q + 1][j] = 0.f;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// SPPF pyramid: three chained 5x5/s1/p2 max-pools (block.py:228-232).  A CTA owns one image and 32 channels
// (lane = channel): the H x W plane is staged in shared memory once and each pool runs as a separable 5-tap row
// pass + 5-tap column pass (10 shared-memory reads per output instead of the 169 global reads of the naive
// window); every pool's result is the next pool's input and is written to its concat slice.  Warps walk the
// pixels incrementally (no integer division).  max() is exact in any precision.
constexpr int SPPF_CH = 32;
template <typename T>
__global__ void __launch_bounds__(NT) sppf_kernel(const fce_sppf_desc d, T* buf) {
    extern __shared__ float sp[];  // [2][H*W][32]
    const int HW = d.H * d.W, W = d.W, H = d.H;
    float* A = sp;
    float* Bm = sp + (size_t)HW * SPPF_CH;
    const int chunks = (d.C + SPPF_CH - 1) / SPPF_CH;
    const int b = blockIdx.x / chunks, c0 = (blockIdx.x % chunks) * SPPF_CH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = NT / 32;
    const bool c_ok = c0 + lane < d.C;
    T* base = buf + (size_t)b * HW * d.pitch + d.off + c0 + lane;
    for (int p = warp; p < HW; p += NW) A[p * SPPF_CH + lane] = c_ok ? Elem<T>::to_f(base[(size_t)p * d.pitch]) : 0.f;
    __syncthreads();
    const int h_step = NW / W, w_step = NW % W;  // pixel index advances by NW per iteration
    for (int level = 1; level <= 3; ++level) {
        int h = warp / W, w0 = warp % W;
        for (int p = warp; p < HW; p += NW) {  // row pass A -> Bm
            float m = A[p * SPPF_CH + lane];
            if (w0 >= 1) m = fmaxf(m, A[(p - 1) * SPPF_CH + lane]);
            if (w0 >= 2) m = fmaxf(m, A[(p - 2) * SPPF_CH + lane]);
            if (w0 + 1 < W) m = fmaxf(m, A[(p + 1) * SPPF_CH + lane]);
            if (w0 + 2 < W) m = fmaxf(m, A[(p + 2) * SPPF_CH + lane]);
            Bm[p * SPPF_CH + lane] = m;
            w0 += w_step;
            h += h_step;
            if (w0 >= W) {
                w0 -= W;
                ++h;
            }
        }
        __syncthreads();
        h = warp / W;
        w0 = warp % W;
        for (int p = warp; p < HW; p += NW) {  // column pass Bm -> A (+ global slice `level`)
            float m = Bm[p * SPPF_CH + lane];
            if (h >= 1) m = fmaxf(m, Bm[(p - W) * SPPF_CH + lane]);
            if (h >= 2) m = fmaxf(m, Bm[(p - 2 * W) * SPPF_CH + lane]);
            if (h + 1 < H) m = fmaxf(m, Bm[(p + W) * SPPF_CH + lane]);
            if (h + 2 < H) m = fmaxf(m, Bm[(p + 2 * W) * SPPF_CH + lane]);
            A[p * SPPF_CH + lane] = m;
            if (c_ok) base[(size_t)p * d.pitch + level * d.C] = Elem<T>::from_f(m);
            w0 += w_step;
            h += h_step;
            if (w0 >= W) {
                w0 -= W;
                ++h;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// SPPF pyramid, version 2 (16-byte aligned views): same chained separable passes, but an item is a 16-byte channel
// vector (8 bf16 / 4 fp32) kept in its STORAGE type in shared memory - max() is exact in any precision, so bf16
// planes need half the space - and every global access is a 16-byte vector.  A CTA owns (image, 16 bf16 / 8 fp32
// channels): 2 x HW x 32 bytes of shared memory (25.6 KB for the 20 x 20 map of a 640^2 image, 102 KB for the
// 40 x 40 map at 1280^2), so several CTAs share an SM and hide each other's barriers.
template <typename T>
__device__ __forceinline__ uint4 vmax16(uint4 a, uint4 b);
template <>
__device__ __forceinline__ uint4 vmax16<__nv_bfloat16>(uint4 a, uint4 b) {
    uint4 r;
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
    __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
    return r;
}
template <>
__device__ __forceinline__ uint4 vmax16<float>(uint4 a, uint4 b) {
    return make_uint4(__float_as_uint(fmaxf(__uint_as_float(a.x), __uint_as_float(b.x))),
                      __float_as_uint(fmaxf(__uint_as_float(a.y), __uint_as_float(b.y))),
                      __float_as_uint(fmaxf(__uint_as_float(a.z), __uint_as_float(b.z))),
                      __float_as_uint(fmaxf(__uint_as_float(a.w), __uint_as_float(b.w))));
}

template <typename T>
__global__ void __launch_bounds__(NT) sppf_v2_kernel(const fce_sppf_desc d, T* buf, int vpc) {
    pdl_trigger();
    extern __shared__ uint4 spv[];  // [2][H*W][vpc]
    constexpr int N = Vec16<T>::N;
    const int HW = d.H * d.W, W = d.W, H = d.H;
    const int items = HW * vpc;
    uint4* A = spv;
    uint4* Bm = spv + items;
    const int chunks = (d.C + vpc * N - 1) / (vpc * N);
    const int b = blockIdx.x / chunks, c0 = (blockIdx.x % chunks) * vpc * N;
    T* base = buf + (size_t)b * HW * d.pitch + d.off + c0;
    // per-thread item list is fixed across passes: item i -> pixel p = i / vpc (h, w), vector v = i % vpc
    for (int i = threadIdx.x; i < items; i += NT) {
        const int p = i / vpc, v = i - p * vpc;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (c0 + v * N < d.C) val = *reinterpret_cast<const uint4*>(base + (size_t)p * d.pitch + v * N);
        A[i] = val;
    }
    __syncthreads();
    for (int level = 1; level <= 3; ++level) {
        for (int i = threadIdx.x; i < items; i += NT) {  // row pass A -> Bm
            const int p = i / vpc;
            const int w0 = p % W;
            uint4 m = A[i];
            if (w0 >= 1) m = vmax16<T>(m, A[i - vpc]);
            if (w0 >= 2) m = vmax16<T>(m, A[i - 2 * vpc]);
            if (w0 + 1 < W) m = vmax16<T>(m, A[i + vpc]);
            if (w0 + 2 < W) m = vmax16<T>(m, A[i + 2 * vpc]);
            Bm[i] = m;
        }
        __syncthreads();
        const int rs = W * vpc;
        for (int i = threadIdx.x; i < items; i += NT) {  // column pass Bm -> A (+ global slice `level`)
            const int p = i / vpc, v = i - p * vpc;
            const int h = p / W;
            uint4 m = Bm[i];
            if (h >= 1) m = vmax16<T>(m, Bm[i - rs]);
            if (h >= 2) m = vmax16<T>(m, Bm[i - 2 * rs]);
            if (h + 1 < H) m = vmax16<T>(m, Bm[i + rs]);
            if (h + 2 < H) m = vmax16<T>(m, Bm[i + 2 * rs]);
            A[i] = m;
            if (c0 + v * N < d.C)
                *reinterpret_cast<uint4*>(base + (size_t)p * d.pitch + level * d.C + v * N) = m;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(NT) upsample_kernel(const fce_upsample_desc d, const T* __restrict__ x, T* y) {
    pdl_trigger();
    constexpr int N = CV<T, VEC>::N;
    const int cv = d.C / N;
    const int Wo = d.W * 2;
    const int row = blockIdx.x;  // b*Ho + ho
    const int Ho = d.H * 2;
    const int b = row / Ho, ho = row - b * Ho;
    const T* xr = x + ((size_t)(b * d.H + (ho >> 1)) * d.W) * d.in_pitch + d.in_off;
    T* yr = y + (size_t)row * Wo * d.out_pitch + d.out_off;
    for (int cb = threadIdx.x; cb < cv; cb += blockDim.x)
        for (int pw = threadIdx.y; pw < Wo; pw += blockDim.y) {
            float v[N];
            CV<T, VEC>::load(xr + (size_t)(pw >> 1) * d.in_pitch + cb * N, v);
            CV<T, VEC>::store(yr + (size_t)pw * d.out_pitch + cb * N, v);
        }
}

// ------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(NT) bifpn_kernel(const fce_bifpn_desc d, const T* __restrict__ x0,
                                                   const T* __restrict__ x1, const T* __restrict__ x2, T* y) {
    pdl_trigger();
    constexpr int N = CV<T, VEC>::N;
    const int cv = d.C / N;
    const int row = blockIdx.x;  // b*H + h
    const int b = row / d.H, h = row - b * d.H;
    const T* xs[3] = {x0, x1, x2};
    const T* xr[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        if (s < d.n)
            xr[s] = d.up[s] ? xs[s] + ((size_t)(b * (d.H >> 1) + (h >> 1)) * (d.W >> 1)) * d.pitch[s] + d.off[s]
                            : xs[s] + (size_t)row * d.W * d.pitch[s] + d.off[s];
        else
            xr[s] = nullptr;
    }
    T* yr = y + (size_t)row * d.W * d.out_pitch + d.out_off;
    for (int cb = threadIdx.x; cb < cv; cb += blockDim.x)
        for (int pw = threadIdx.y; pw < d.W; pw += blockDim.y) {
            float acc[N];
#pragma unroll
            for (int j = 0; j < N; ++j) acc[j] = 0.f;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                if (s >= d.n) break;
                float v[N];
                CV<T, VEC>::load(xr[s] + (size_t)(d.up[s] ? (pw >> 1) : pw) * d.pitch[s] + cb * N, v);
#pragma unroll
                for (int j = 0; j < N; ++j) acc[j] = fmaf(d.wn[s], v[j], acc[j]);
            }
            CV<T, VEC>::store(yr + (size_t)pw * d.out_pitch + cb * N, acc);
        }
}

// ------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(NT) copy_kernel(const fce_copy_desc d, const T* __restrict__ x, T* y) {
    pdl_trigger();
    constexpr int N = CV<T, VEC>::N;
    const int cv = d.C / N;
    const size_t row = blockIdx.x;
    const T* xr = x + row * d.W * d.in_pitch + d.in_off;
    T* yr = y + row * d.W * d.out_pitch + d.out_off;
    for (int cb = threadIdx.x; cb < cv; cb += blockDim.x)
        for (int pw = threadIdx.y; pw < d.W; pw += blockDim.y) {
            float v[N];
            CV<T, VEC>::load(xr + (size_t)pw * d.in_pitch + cb * N, v);
            CV<T, VEC>::store(yr + (size_t)pw * d.out_pitch + cb * N, v);
        }
}

// ------------------------------------------------------------------------------------------------
// N consecutive fp32 gate values: 16-byte loads on the vector path (the host checks the alignment)
template <int N, bool VEC>
__device__ __forceinline__ void ldf(const float* __restrict__ p, float* f) {
    if constexpr (VEC && N % 4 == 0) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p) + q);
            f[4 * q] = v.x;
            f[4 * q + 1] = v.y;
            f[4 * q + 2] = v.z;
            f[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) f[j] = __ldg(p + j);
    }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(NT) gate_kernel(const fce_gate_desc d, const T* __restrict__ x,
                                                  const float* __restrict__ gh, const float* __restrict__ gw, T* y) {
    pdl_trigger();
    constexpr int N = CV<T, VEC>::N;
    const int cv = d.C / N;
    const int row = blockIdx.x;  // b*H + h
    const int b = row / d.H, h = row - b * d.H;
    const T* xr = x + (size_t)row * d.W * d.in_pitch + d.in_off;
    T* yr = y + (size_t)row * d.W * d.out_pitch + d.out_off;
    const float* ar = gh + b * d.gh_bstride + h * d.gh_rstride;
    const float* br = d.mode == 1 ? nullptr : gw + b * d.gw_bstride;
    for (int cb = threadIdx.x; cb < cv; cb += blockDim.x) {
        const int c = cb * N;
        float a[N];
        ldf<N, VEC>(ar + c, a);
        for (int pw0 = threadIdx.y; pw0 < d.W; pw0 += 2 * blockDim.y) {
            const int pw1 = pw0 + blockDim.y;
            const bool two = pw1 < d.W;
            float v0[N], v1[N], b0[N], b1[N];
            CV<T, VEC>::load(xr + (size_t)pw0 * d.in_pitch + c, v0);
            if (two) CV<T, VEC>::load(xr + (size_t)pw1 * d.in_pitch + c, v1);
            if (d.mode != 1) {
                ldf<N, VEC>(br + pw0 * d.gw_rstride + c, b0);
                if (two) ldf<N, VEC>(br + pw1 * d.gw_rstride + c, b1);
            }
#pragma unroll
            for (int j = 0; j < N; ++j) {
                if (d.mode == 1) {
                    v0[j] *= a[j];
                    v1[j] *= a[j];
                } else if (d.mode == 0) {
                    v0[j] = v0[j] * a[j] * b0[j];
                    v1[j] = v1[j] * a[j] * b1[j];
                } else {
                    v0[j] *= sigmoid_f(a[j] + b0[j]);
                    v1[j] *= sigmoid_f(a[j] + b1[j]);
                }
            }
            CV<T, VEC>::store(yr + (size_t)pw0 * d.out_pitch + c, v0);
            if (two) CV<T, VEC>::store(yr + (size_t)pw1 * d.out_pitch + c, v1);
        }
    }
}

// Column-walking variant (vector path): a thread owns one 16-byte channel vector of one pixel COLUMN and walks
// down a band of image rows, four rows per step (four independent 16-byte loads in flight).  Its W-gate values stay
// in registers for the whole band, the H-gate row (1 KB per 256 channels) is shared by the CTA's eight pixel columns
// through L1.  The row-per-CTA kernel above re-read the whole [W, C] fp32 W-gate of the image in every CTA - twice
// the bytes of x itself through L2 - and kept only two loads in flight per thread (4.2 TB/s).
constexpr int GC_PX = 8;     // pixel columns per CTA (blockDim.y)
constexpr int GC_ROWS = 4;   // rows per step
template <typename T>
__global__ void __launch_bounds__(32 * GC_PX) gate_cols_kernel(const fce_gate_desc d, const T* __restrict__ x,
                                                               const float* __restrict__ gh, const float* __restrict__ gw,
                                                               T* y, int col_groups, int band_rows) {
    pdl_trigger();
    constexpr int N = Vec16<T>::N;
    int u = blockIdx.x;  // (b, band, column group): column group fastest
    const int cg = u % col_groups;
    u /= col_groups;
    const int bands = (d.H + band_rows - 1) / band_rows;
    const int band = u % bands, b = u / bands;
    const int c = (blockIdx.y * 32 + threadIdx.x) * N;
    const int pw = cg * GC_PX + threadIdx.y;
    if (c >= d.C || pw >= d.W) return;
    const int h0 = band * band_rows, h1 = min(d.H, h0 + band_rows);
    float bwv[N];
    if (d.mode != 1) ldf<N, true>(gw + b * d.gw_bstride + pw * d.gw_rstride + c, bwv);
    const size_t in_row = (size_t)d.W * d.in_pitch, out_row = (size_t)d.W * d.out_pitch;
    const T* xp = x + ((size_t)(b * d.H + h0) * d.W + pw) * d.in_pitch + d.in_off + c;
    T* yp = y + ((size_t)(b * d.H + h0) * d.W + pw) * d.out_pitch + d.out_off + c;
    const float* ap = gh + b * d.gh_bstride + h0 * d.gh_rstride + c;
    for (int h = h0; h < h1; h += GC_ROWS) {
        Vec16<T> v[GC_ROWS];
        float a[GC_ROWS][N];
#pragma unroll
        for (int r = 0; r < GC_ROWS; ++r)
            if (h + r < h1) {
                v[r].load_nc(xp + r * in_row);
                ldf<N, true>(ap + r * d.gh_rstride, a[r]);
            }
#pragma unroll
        for (int r = 0; r < GC_ROWS; ++r)
            if (h + r < h1) {
                float f[N];
                v[r].unpack(f);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    if (d.mode == 1) f[j] *= a[r][j];
                    else if (d.mode == 0) f[j] = f[j] * a[r][j] * bwv[j];
                    else f[j] *= sigmoid_f(a[r][j] + bwv[j]);
                }
                Vec16<T> o;
                o.pack(f);
                o.store(yp + r * out_row);
            }
        xp += GC_ROWS * in_row;
        yp += GC_ROWS * out_row;
        ap += GC_ROWS * d.gh_rstride;
    }
}

inline bool multiple_of(int n, std::initializer_list<long long> vals) {
    for (long long v : vals)
        if (v % n) return false;
    return true;
}
inline bool ptr16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// Calls f(T{}, vec) with T = storage type; vec = whether 16-byte channel vectors are legal.
template <typename F>
int by_dtype(int dtype, F&& f) {
    if (dtype == FCE_BF16) return f(__nv_bfloat16{});
    if (dtype == FCE_F32) return f(float{});
    return FCE_ERR_UNSUPPORTED;
}

}  // namespace
}  // namespace fce

namespace fce {
int dwconv3x3_tma(const fce_dwconv_desc* d, const void* x, const float* w, const float* bias, const void* add, void* y,
                  cudaStream_t st);
// FCE_DW_IMPL=1: column-walk kernel, 2: 4-byte-lane row kernel (A/B timing of the depthwise implementations)
static const int g_dw_impl = [] {
    const char* e = getenv("FCE_DW_IMPL");
    return e ? atoi(e) : 0;
}();
}  // namespace fce

using namespace fce;

extern "C" int fce_dwconv3x3(const fce_dwconv_desc* d, const void* x, const float* w, const float* bias,
                             const void* add, void* y, void* stream) {
    if (!d || !x || !w || !bias || !y || d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if ((d->dtype == FCE_BF16 || d->dtype == FCE_F32) && g_dw_impl == 0) {  // TMA-pipelined kernel when TMA can describe the views
        const int esz = d->dtype == FCE_BF16 ? 2 : 4;
        const int rc = dwconv3x3_tma(d, (const char*)x + (size_t)d->in_off * esz, w, bias,
                                     add ? (const char*)add + (size_t)d->add_off * esz : nullptr,
                                     (char*)y + (size_t)d->out_off * esz, st);
        if (rc != FCE_ERR_UNSUPPORTED) return rc;
    }
    return by_dtype(d->dtype, [&](auto tag) -> int {
        using T = decltype(tag);
        constexpr int E = 4 / (int)sizeof(T);  // channels per 4-byte lane
        if (!multiple_of(E, {d->C, d->in_pitch, d->in_off, d->out_pitch, d->out_off, add ? d->add_pitch : 0,
                             add ? d->add_off : 0}) ||
            (((uintptr_t)x | (uintptr_t)y | (uintptr_t)add) & 3))
            return FCE_ERR_ALIGNMENT;
        constexpr int N = 16 / (int)sizeof(T);
        if (g_dw_impl <= 1 && multiple_of(N, {d->C, d->in_pitch, d->in_off, d->out_pitch, d->out_off, add ? d->add_pitch : 0,
                            add ? d->add_off : 0}) && ptr16(x) && ptr16(y) && ptr16(add)) {
            const int cgroups = (d->C + DW2_VECS * N - 1) / (DW2_VECS * N);
            const int wtiles = (d->W + DW2_COLS - 1) / DW2_COLS;
            // bands of ~20 rows (halo overhead 10 %), but enough of them to give every SM a few CTAs
            int bands = (d->H + 19) / 20;
            while ((long long)d->B * bands * wtiles * cgroups < 4 * kNumSMs && bands * 4 <= d->H) bands *= 2;
            const int band_rows = (d->H + bands - 1) / bands;
            bands = (d->H + band_rows - 1) / band_rows;
            if ((long long)d->B * bands > 65535) return FCE_ERR_UNSUPPORTED;
            dim3 grid(wtiles * cgroups, d->B * bands);
            dwconv3x3_v2_kernel<T><<<grid, DW2_THREADS, 0, st>>>(*d, (const T*)x, w, bias, (const T*)add, (T*)y,
                                                                 band_rows, bands, cgroups);
            return check_launch();
        }
        const int groups = (d->C + 32 * E - 1) / (32 * E);
        const long long units = (long long)d->B * d->H * groups;
        if (units > 0x7fffffffLL) return FCE_ERR_UNSUPPORTED;
        const int grid = (int)((units + DW_WARPS - 1) / DW_WARPS);
        dwconv3x3_kernel<T><<<grid, DW_WARPS * 32, 0, st>>>(*d, (const T*)x, w, bias, (const T*)add, (T*)y, groups,
                                                            (int)units);
        return check_launch();
    });
}

extern "C" int fce_sppf_pool(const fce_sppf_desc* d, void* buf, void* stream) {
    if (!d || !buf || d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    if (d->off + 4 * d->C > d->pitch) return FCE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return by_dtype(d->dtype, [&](auto tag) -> int {
        using T = decltype(tag);
        constexpr int N = 16 / (int)sizeof(T);
        if (multiple_of(N, {d->C, d->pitch, d->off}) && ptr16(buf)) {
            int vpc = 2;
            size_t sm2 = (size_t)2 * d->H * d->W * vpc * 16;
            if (sm2 > 110 * 1024) { vpc = 1; sm2 /= 2; }  // very large maps: one vector per CTA
            if (sm2 <= 220 * 1024) {
                static DeviceOnce attr2[2];  // per-device attribute, one flag set per element type
                constexpr int ti2 = sizeof(T) == 2 ? 0 : 1;
                int dev2 = 0;
                if (attr2[ti2].pending(&dev2)) {
                    cudaError_t e = cudaFuncSetAttribute(sppf_v2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
                    if (e != cudaSuccess) {
                        set_cuda_error(e);
                        return FCE_ERR_CUDA;
                    }
                    attr2[ti2].done(dev2);
                }
                const int chunks2 = (d->C + vpc * N - 1) / (vpc * N);
                sppf_v2_kernel<T><<<d->B * chunks2, NT, sm2, st>>>(*d, (T*)buf, vpc);
                return check_launch();
            }
        }
        // two fp32 copies of a 32-channel plane in shared memory: up to 29 x 29 ... 880 x 880 pixels of input per
        // 100 KB; the path's P5 maps are 20 x 20 (640) and 40 x 40 (1280)
        const int HW = d->H * d->W;
        const size_t smem = (size_t)2 * HW * SPPF_CH * sizeof(float);
        if (smem > 200 * 1024) return FCE_ERR_UNSUPPORTED;
        static DeviceOnce attr_set[2];
        constexpr int ti = sizeof(T) == 2 ? 0 : 1;
        int dev_ = 0;
        if (attr_set[ti].pending(&dev_)) {
            cudaError_t e = cudaFuncSetAttribute(sppf_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) {
                set_cuda_error(e);
                return FCE_ERR_CUDA;
            }
            attr_set[ti].done(dev_);
        }
        const int chunks = (d->C + SPPF_CH - 1) / SPPF_CH;
        sppf_kernel<T><<<d->B * chunks, NT, smem, st>>>(*d, (T*)buf);
        return check_launch();
    });
}

extern "C" int fce_upsample2x(const fce_upsample_desc* d, const void* x, void* y, void* stream) {
    if (!d || !x || !y || d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return by_dtype(d->dtype, [&](auto tag) -> int {
        using T = decltype(tag);
        constexpr int N = 16 / (int)sizeof(T);
        const bool vec = multiple_of(N, {d->C, d->in_pitch, d->in_off, d->out_pitch, d->out_off}) && ptr16(x) && ptr16(y);
        const RowGrid g = row_grid(d->B * d->H * 2, d->C / (vec ? N : 1));
        if (vec) upsample_kernel<T, true><<<g.grid, g.block, 0, st>>>(*d, (const T*)x, (T*)y);
        else upsample_kernel<T, false><<<g.grid, g.block, 0, st>>>(*d, (const T*)x, (T*)y);
        return check_launch();
    });
}

extern "C" int fce_bifpn_fuse(const fce_bifpn_desc* d, const void* x0, const void* x1, const void* x2, void* y,
                              void* stream) {
    if (!d || !x0 || !x1 || !y || d->n < 2 || d->n > 3 || (d->n == 3 && !x2)) return FCE_ERR_BAD_ARG;
    if (d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    for (int i = 0; i < d->n; ++i)
        if (d->up[i] && ((d->H | d->W) & 1)) return FCE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return by_dtype(d->dtype, [&](auto tag) -> int {
        using T = decltype(tag);
        constexpr int N = 16 / (int)sizeof(T);
        bool vec = multiple_of(N, {d->C, d->out_pitch, d->out_off}) && ptr16(x0) && ptr16(x1) && ptr16(x2) && ptr16(y);
        for (int i = 0; i < d->n; ++i) vec = vec && multiple_of(N, {d->pitch[i], d->off[i]});
        const RowGrid g = row_grid(d->B * d->H, d->C / (vec ? N : 1));
        if (vec) bifpn_kernel<T, true><<<g.grid, g.block, 0, st>>>(*d, (const T*)x0, (const T*)x1, (const T*)x2, (T*)y);
        else bifpn_kernel<T, false><<<g.grid, g.block, 0, st>>>(*d, (const T*)x0, (const T*)x1, (const T*)x2, (T*)y);
        return check_launch();
    });
}

extern "C" int fce_copy_view(const fce_copy_desc* d, const void* x, void* y, void* stream) {
    if (!d || !x || !y || d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return by_dtype(d->dtype, [&](auto tag) -> int {
        using T = decltype(tag);
        constexpr int N = 16 / (int)sizeof(T);
        const bool vec = multiple_of(N, {d->C, d->in_pitch, d->in_off, d->out_pitch, d->out_off}) && ptr16(x) && ptr16(y);
        const RowGrid g = row_grid(d->B * d->H, d->C / (vec ? N : 1));
        if (vec) copy_kernel<T, true><<<g.grid, g.block, 0, st>>>(*d, (const T*)x, (T*)y);
        else copy_kernel<T, false><<<g.grid, g.block, 0, st>>>(*d, (const T*)x, (T*)y);
        return check_launch();
    });
}

extern "C" int fce_gate_apply(const fce_gate_desc* d, const void* x, const float* gh, const float* gw, void* y,
                              void* stream) {
    if (!d || !x || !gh || !y || d->mode < 0 || d->mode > 2 || (d->mode != 1 && !gw)) return FCE_ERR_BAD_ARG;
    if (d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return by_dtype(d->dtype, [&](auto tag) -> int {
        using T = decltype(tag);
        constexpr int N = 16 / (int)sizeof(T);
        const bool vec = multiple_of(N, {d->C, d->in_pitch, d->in_off, d->out_pitch, d->out_off}) && ptr16(x) && ptr16(y) &&
                         multiple_of(4, {d->gh_bstride, d->gh_rstride, d->gw_bstride, d->gw_rstride}) && ptr16(gh) && ptr16(gw);
        static const bool rows_only = [] {  // FCE_GATE_ROWS=1: the row-per-CTA kernel everywhere (A/B timing)
            const char* e = getenv("FCE_GATE_ROWS");
            return e && e[0] == '1';
        }();
        if (vec && !rows_only && x != y && d->C / N >= 16) {
            // column walk: (image, row band, 8-column group) CTAs; bands sized for ~8 CTAs per SM
            const int col_groups = (d->W + GC_PX - 1) / GC_PX;
            const int cchunks = (d->C / N + 31) / 32;
            int bands = (8 * kNumSMs + d->B * col_groups * cchunks - 1) / (d->B * col_groups * cchunks);
            const int max_bands = (d->H + 2 * GC_ROWS - 1) / (2 * GC_ROWS);
            bands = bands < 1 ? 1 : (bands > max_bands ? max_bands : bands);
            const int band_rows = ((d->H + bands - 1) / bands + GC_ROWS - 1) / GC_ROWS * GC_ROWS;
            bands = (d->H + band_rows - 1) / band_rows;
            const long long ctas = (long long)d->B * bands * col_groups;
            if (ctas <= 0x7fffffffLL && cchunks <= 65535) {
                gate_cols_kernel<T><<<dim3((unsigned)ctas, cchunks), dim3(32, GC_PX), 0, st>>>(*d, (const T*)x, gh, gw, (T*)y,
                                                                                           col_groups, band_rows);
                return check_launch();
            }
        }
        const RowGrid g = row_grid(d->B * d->H, d->C / (vec ? N : 1));
        if (vec) gate_kernel<T, true><<<g.grid, g.block, 0, st>>>(*d, (const T*)x, gh, gw, (T*)y);
        else gate_kernel<T, false><<<g.grid, g.block, 0, st>>>(*d, (const T*)x, gh, gw, (T*)y);
        return check_launch();
    });
}

// Bandwidth-bound NHWC kernels: depthwise 3x3, SPPF max-pool pyramid, nearest 2x upsample,
// BiFPN weighted fusion, gate application, view copy.  No tensor cores: every thread owns one
// 16-byte channel vector (8 bf16 / 4 fp32) of one pixel, so a warp touches whole 128-byte lines;
// grids are sized in multiples of the SM count and walk the tensor with a grid-stride loop.
#include "common.cuh"

namespace fce {
namespace {

constexpr int NT = 256;

inline int grid_for(long long items) {
    long long blocks = (items + NT - 1) / NT;
    long long cap = (long long)kNumSMs * 16;  // up to 16 resident 256-thread CTAs' worth of work per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
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

// ------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(NT) dwconv3x3_kernel(const fce_dwconv_desc d, const T* __restrict__ x,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       const T* add, T* y) {
    constexpr int N = CV<T, VEC>::N;
    const int cv = d.C / N;
    const long long total = (long long)d.B * d.H * d.W * cv;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % cv) * N;
        long long p = i / cv;
        const int pw = (int)(p % d.W);
        const int ph = (int)((p / d.W) % d.H);
        const int pb = (int)(p / ((long long)d.W * d.H));
        float acc[N];
#pragma unroll
        for (int j = 0; j < N; ++j) acc[j] = bias[c + j];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int hi = ph + kh - 1;
            if (hi < 0 || hi >= d.H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int wi = pw + kw - 1;
                if (wi < 0 || wi >= d.W) continue;
                float v[N];
                CV<T, VEC>::load(x + ((size_t)(pb * d.H + hi) * d.W + wi) * d.in_pitch + d.in_off + c, v);
                const float* wp = w + (kh * 3 + kw) * d.C + c;
#pragma unroll
                for (int j = 0; j < N; ++j) acc[j] = fmaf(v[j], wp[j], acc[j]);
            }
        }
        if (d.act == FCE_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < N; ++j) acc[j] = silu_acc(acc[j]);
        }
        if (add) {
            float v[N];
            CV<T, VEC>::load(add + (size_t)p * d.add_pitch + d.add_off + c, v);
#pragma unroll
            for (int j = 0; j < N; ++j) acc[j] += v[j];
        }
        CV<T, VEC>::store(y + (size_t)p * d.out_pitch + d.out_off + c, acc);
    }
}

// ------------------------------------------------------------------------------------------------
// SPPF: windows 5/9/13 around each pixel, computed from one sweep over the 13x13 neighbourhood
// (nested row maxima), identical to three chained 5x5/s1/p2 pools with -inf padding.
template <typename T, bool VEC>
__global__ void __launch_bounds__(NT) sppf_kernel(const fce_sppf_desc d, T* buf) {
    constexpr int N = CV<T, VEC>::N;
    const int cv = d.C / N;
    const long long total = (long long)d.B * d.H * d.W * cv;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % cv) * N;
        long long p = i / cv;
        const int pw = (int)(p % d.W);
        const int ph = (int)((p / d.W) % d.H);
        const int pb = (int)(p / ((long long)d.W * d.H));
        float m5[N], m9[N], m13[N];
#pragma unroll
        for (int j = 0; j < N; ++j) m5[j] = m9[j] = m13[j] = -INFINITY;
        for (int dy = -6; dy <= 6; ++dy) {
            const int hi = ph + dy;
            if (hi < 0 || hi >= d.H) continue;
            float r5[N], r9[N], r13[N];
#pragma unroll
            for (int j = 0; j < N; ++j) r5[j] = r9[j] = r13[j] = -INFINITY;
            for (int dx = -6; dx <= 6; ++dx) {
                const int wi = pw + dx;
                if (wi < 0 || wi >= d.W) continue;
                float v[N];
                CV<T, VEC>::load(buf + ((size_t)(pb * d.H + hi) * d.W + wi) * d.pitch + d.off + c, v);
                const int ax = dx < 0 ? -dx : dx;
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    r13[j] = fmaxf(r13[j], v[j]);
                    if (ax <= 4) r9[j] = fmaxf(r9[j], v[j]);
                    if (ax <= 2) r5[j] = fmaxf(r5[j], v[j]);
                }
            }
            const int ay = dy < 0 ? -dy : dy;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                m13[j] = fmaxf(m13[j], r13[j]);
                if (ay <= 4) m9[j] = fmaxf(m9[j], r9[j]);
                if (ay <= 2) m5[j] = fmaxf(m5[j], r5[j]);
            }
        }
        T* o = buf + (size_t)p * d.pitch + d.off + c;
        CV<T, VEC>::store(o + d.C, m5);
        CV<T, VEC>::store(o + 2 * d.C, m9);
        CV<T, VEC>::store(o + 3 * d.C, m13);
    }
}

// ------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(NT) upsample_kernel(const fce_upsample_desc d, const T* __restrict__ x, T* y) {
    constexpr int N = CV<T, VEC>::N;
    const int cv = d.C / N;
    const int Ho = d.H * 2, Wo = d.W * 2;
    const long long total = (long long)d.B * Ho * Wo * cv;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % cv) * N;
        long long p = i / cv;
        const int pw = (int)(p % Wo);
        const int ph = (int)((p / Wo) % Ho);
        const int pb = (int)(p / ((long long)Wo * Ho));
        float v[N];
        CV<T, VEC>::load(x + ((size_t)(pb * d.H + (ph >> 1)) * d.W + (pw >> 1)) * d.in_pitch + d.in_off + c, v);
        CV<T, VEC>::store(y + (size_t)p * d.out_pitch + d.out_off + c, v);
    }
}

// ------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(NT) bifpn_kernel(const fce_bifpn_desc d, const T* __restrict__ x0,
                                                   const T* __restrict__ x1, const T* __restrict__ x2, T* y) {
    constexpr int N = CV<T, VEC>::N;
    const int cv = d.C / N;
    const long long total = (long long)d.B * d.H * d.W * cv;
    const T* xs[3] = {x0, x1, x2};
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % cv) * N;
        long long p = i / cv;
        const int pw = (int)(p % d.W);
        const int ph = (int)((p / d.W) % d.H);
        const int pb = (int)(p / ((long long)d.W * d.H));
        float acc[N];
#pragma unroll
        for (int j = 0; j < N; ++j) acc[j] = 0.f;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            if (s >= d.n) break;
            size_t pix = d.up[s] ? ((size_t)(pb * (d.H >> 1) + (ph >> 1)) * (d.W >> 1) + (pw >> 1)) : (size_t)p;
            float v[N];
            CV<T, VEC>::load(xs[s] + pix * d.pitch[s] + d.off[s] + c, v);
#pragma unroll
            for (int j = 0; j < N; ++j) acc[j] = fmaf(d.wn[s], v[j], acc[j]);
        }
        CV<T, VEC>::store(y + (size_t)p * d.out_pitch + d.out_off + c, acc);
    }
}

// ------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(NT) copy_kernel(const fce_copy_desc d, const T* __restrict__ x, T* y) {
    constexpr int N = CV<T, VEC>::N;
    const int cv = d.C / N;
    const long long total = (long long)d.B * d.H * d.W * cv;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % cv) * N;
        const long long p = i / cv;
        float v[N];
        CV<T, VEC>::load(x + (size_t)p * d.in_pitch + d.in_off + c, v);
        CV<T, VEC>::store(y + (size_t)p * d.out_pitch + d.out_off + c, v);
    }
}

// ------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(NT) gate_kernel(const fce_gate_desc d, const T* __restrict__ x,
                                                  const float* __restrict__ gh, const float* __restrict__ gw, T* y) {
    constexpr int N = CV<T, VEC>::N;
    const int cv = d.C / N;
    const long long total = (long long)d.B * d.H * d.W * cv;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % cv) * N;
        long long p = i / cv;
        const int pw = (int)(p % d.W);
        const int ph = (int)((p / d.W) % d.H);
        const int pb = (int)(p / ((long long)d.W * d.H));
        float v[N];
        CV<T, VEC>::load(x + (size_t)p * d.in_pitch + d.in_off + c, v);
        const float* a = gh + pb * d.gh_bstride + ph * d.gh_rstride + c;
        if (d.mode == 1) {
#pragma unroll
            for (int j = 0; j < N; ++j) v[j] *= a[j];
        } else {
            const float* b = gw + pb * d.gw_bstride + pw * d.gw_rstride + c;
            if (d.mode == 0) {
#pragma unroll
                for (int j = 0; j < N; ++j) v[j] = v[j] * a[j] * b[j];
            } else {
#pragma unroll
                for (int j = 0; j < N; ++j) v[j] *= sigmoid_f(a[j] + b[j]);
            }
        }
        CV<T, VEC>::store(y + (size_t)p * d.out_pitch + d.out_off + c, v);
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

using namespace fce;

extern "C" int fce_dwconv3x3(const fce_dwconv_desc* d, const void* x, const float* w, const float* bias,
                             const void* add, void* y, void* stream) {
    if (!d || !x || !w || !bias || !y || d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return by_dtype(d->dtype, [&](auto tag) {
        using T = decltype(tag);
        constexpr int N = 16 / (int)sizeof(T);
        const bool vec = multiple_of(N, {d->C, d->in_pitch, d->in_off, d->out_pitch, d->out_off,
                                         add ? d->add_pitch : 0, add ? d->add_off : 0}) &&
                         ptr16(x) && ptr16(y) && ptr16(add) && ptr16(w) && ptr16(bias);
        const long long items = (long long)d->B * d->H * d->W * (d->C / (vec ? N : 1));
        if (vec)
            dwconv3x3_kernel<T, true><<<grid_for(items), NT, 0, st>>>(*d, (const T*)x, w, bias, (const T*)add, (T*)y);
        else
            dwconv3x3_kernel<T, false><<<grid_for(items), NT, 0, st>>>(*d, (const T*)x, w, bias, (const T*)add, (T*)y);
        return check_launch();
    });
}

extern "C" int fce_sppf_pool(const fce_sppf_desc* d, void* buf, void* stream) {
    if (!d || !buf || d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    if (d->off + 4 * d->C > d->pitch) return FCE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return by_dtype(d->dtype, [&](auto tag) {
        using T = decltype(tag);
        constexpr int N = 16 / (int)sizeof(T);
        const bool vec = multiple_of(N, {d->C, d->pitch, d->off}) && ptr16(buf);
        const long long items = (long long)d->B * d->H * d->W * (d->C / (vec ? N : 1));
        if (vec) sppf_kernel<T, true><<<grid_for(items), NT, 0, st>>>(*d, (T*)buf);
        else sppf_kernel<T, false><<<grid_for(items), NT, 0, st>>>(*d, (T*)buf);
        return check_launch();
    });
}

extern "C" int fce_upsample2x(const fce_upsample_desc* d, const void* x, void* y, void* stream) {
    if (!d || !x || !y || d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return by_dtype(d->dtype, [&](auto tag) {
        using T = decltype(tag);
        constexpr int N = 16 / (int)sizeof(T);
        const bool vec = multiple_of(N, {d->C, d->in_pitch, d->in_off, d->out_pitch, d->out_off}) && ptr16(x) && ptr16(y);
        const long long items = (long long)d->B * d->H * d->W * 4 * (d->C / (vec ? N : 1));
        if (vec) upsample_kernel<T, true><<<grid_for(items), NT, 0, st>>>(*d, (const T*)x, (T*)y);
        else upsample_kernel<T, false><<<grid_for(items), NT, 0, st>>>(*d, (const T*)x, (T*)y);
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
    return by_dtype(d->dtype, [&](auto tag) {
        using T = decltype(tag);
        constexpr int N = 16 / (int)sizeof(T);
        bool vec = multiple_of(N, {d->C, d->out_pitch, d->out_off}) && ptr16(x0) && ptr16(x1) && ptr16(x2) && ptr16(y);
        for (int i = 0; i < d->n; ++i) vec = vec && multiple_of(N, {d->pitch[i], d->off[i]});
        const long long items = (long long)d->B * d->H * d->W * (d->C / (vec ? N : 1));
        if (vec) bifpn_kernel<T, true><<<grid_for(items), NT, 0, st>>>(*d, (const T*)x0, (const T*)x1, (const T*)x2, (T*)y);
        else bifpn_kernel<T, false><<<grid_for(items), NT, 0, st>>>(*d, (const T*)x0, (const T*)x1, (const T*)x2, (T*)y);
        return check_launch();
    });
}

extern "C" int fce_copy_view(const fce_copy_desc* d, const void* x, void* y, void* stream) {
    if (!d || !x || !y || d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return by_dtype(d->dtype, [&](auto tag) {
        using T = decltype(tag);
        constexpr int N = 16 / (int)sizeof(T);
        const bool vec = multiple_of(N, {d->C, d->in_pitch, d->in_off, d->out_pitch, d->out_off}) && ptr16(x) && ptr16(y);
        const long long items = (long long)d->B * d->H * d->W * (d->C / (vec ? N : 1));
        if (vec) copy_kernel<T, true><<<grid_for(items), NT, 0, st>>>(*d, (const T*)x, (T*)y);
        else copy_kernel<T, false><<<grid_for(items), NT, 0, st>>>(*d, (const T*)x, (T*)y);
        return check_launch();
    });
}

extern "C" int fce_gate_apply(const fce_gate_desc* d, const void* x, const float* gh, const float* gw, void* y,
                              void* stream) {
    if (!d || !x || !gh || !y || d->mode < 0 || d->mode > 2 || (d->mode != 1 && !gw)) return FCE_ERR_BAD_ARG;
    if (d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return by_dtype(d->dtype, [&](auto tag) {
        using T = decltype(tag);
        constexpr int N = 16 / (int)sizeof(T);
        const bool vec = multiple_of(N, {d->C, d->in_pitch, d->in_off, d->out_pitch, d->out_off}) && ptr16(x) && ptr16(y);
        const long long items = (long long)d->B * d->H * d->W * (d->C / (vec ? N : 1));
        if (vec) gate_kernel<T, true><<<grid_for(items), NT, 0, st>>>(*d, (const T*)x, gh, gw, (T*)y);
        else gate_kernel<T, false><<<grid_for(items), NT, 0, st>>>(*d, (const T*)x, gh, gw, (T*)y);
        return check_launch();
    });
}

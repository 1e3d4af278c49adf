// Shared device helpers for the fce_yolo_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/fce_yolo_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "fce_yolo_b200 kernels target sm_100a (B200) only"
#endif

namespace fce {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_cuda_error(cudaError_t e);

// cudaFuncSetAttribute (the > 48 KB dynamic shared-memory opt-in) applies to the CURRENT device only: one process driving
// several GPUs must repeat it per device.  Bit d of the mask = done on device d; setting twice is harmless, so two
// threads racing through pending() at worst both set the attributes.
struct DeviceOnce {
    std::atomic<unsigned long long> mask{0};
    bool pending(int* dev) {
        *dev = 0;
        if (cudaGetDevice(dev) != cudaSuccess || *dev < 0 || *dev > 63) return true;
        return ((mask.load(std::memory_order_acquire) >> *dev) & 1ull) == 0;
    }
    void done(int dev) {
        if (dev >= 0 && dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release);
    }
};

inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_cuda_error(e);
        return FCE_ERR_CUDA;
    }
    return FCE_OK;
}

template <typename T>
struct Elem;
template <>
struct Elem<float> {
    static __device__ __forceinline__ float to_f(float v) { return v; }
    static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <>
struct Elem<__nv_bfloat16> {
    static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct Elem<uint8_t> {
    static __device__ __forceinline__ float to_f(uint8_t v) { return (float)v; }
};

// 16-byte vector of T: 8 bf16 or 4 fp32.
template <typename T>
struct Vec16 {
    static constexpr int N = 16 / sizeof(T);
    uint4 raw;
    __device__ __forceinline__ void load(const T* p) { raw = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void load_nc(const T* p) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
                     : "l"(p));
    }
    __device__ __forceinline__ void store(T* p) const { *reinterpret_cast<uint4*>(p) = raw; }
    __device__ __forceinline__ void unpack(float* f) const;
    __device__ __forceinline__ void pack(const float* f);
};
template <>
__device__ __forceinline__ void Vec16<float>::unpack(float* f) const {
    f[0] = __uint_as_float(raw.x); f[1] = __uint_as_float(raw.y);
    f[2] = __uint_as_float(raw.z); f[3] = __uint_as_float(raw.w);
}
template <>
__device__ __forceinline__ void Vec16<float>::pack(const float* f) {
    raw.x = __float_as_uint(f[0]); raw.y = __float_as_uint(f[1]);
    raw.z = __float_as_uint(f[2]); raw.w = __float_as_uint(f[3]);
}
template <>
__device__ __forceinline__ void Vec16<__nv_bfloat16>::unpack(float* f) const {
    const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(r[i] << 16);
        f[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u);
    }
}
template <>
__device__ __forceinline__ void Vec16<__nv_bfloat16>::pack(const float* f) {
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        r[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    raw = make_uint4(r[0], r[1], r[2], r[3]);
}

__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.f + __expf(-v)); }
__device__ __forceinline__ float sigmoid_f(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }
// accurate versions for fp32 mode (expf is ~1 ulp; __expf has ~2^-21 relative error near 0)
__device__ __forceinline__ float silu_acc(float v) { return v / (1.f + expf(-v)); }
__device__ __forceinline__ float sigmoid_acc(float v) { return 1.f / (1.f + expf(-v)); }

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == FCE_ACT_SILU) return silu_acc(v);
    if (act == FCE_ACT_SIGMOID) return sigmoid_acc(v);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Programmatic dependent launch, primary side: lets the CTAs of the NEXT kernel in the stream be scheduled (and run
// their prologue up to their own griddepcontrol.wait) once every CTA of this grid has issued it.  The tcgen05 conv
// kernels are launched with the programmatic-serialization attribute and wait before touching activations
// (tc_common.cuh); for a successor launched normally this is a no-op.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace fce

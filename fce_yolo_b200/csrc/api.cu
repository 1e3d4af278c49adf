// C-ABI glue: version / error reporting and the convolution front door (kernel selection).
#include <atomic>

#include "common.cuh"

namespace fce {

static thread_local cudaError_t g_last_err = cudaSuccess;
void set_cuda_error(cudaError_t e) { g_last_err = e; }

int conv2d_simt(const fce_conv_desc*, const void*, const void*, const float*, const void*, void*, cudaStream_t);
int conv2d_tc(const fce_conv_desc*, const void*, const void*, const float*, const void*, void*, cudaStream_t,
              const fce_detect_epi_desc* epi = nullptr);
bool conv2d_tc_supported(const fce_conv_desc*, const void*, const void*, const void*, const void*);
extern std::atomic<long long> g_conv_stats[4];
#ifdef FCE_DEBUG
void conv_tc_set_profile(int on);
int conv_tc_profile(long long* out, int n);
#endif

}  // namespace fce

using namespace fce;

extern "C" int fce_abi_version(void) { return 1; }

extern "C" const char* fce_last_cuda_error(void) { return cudaGetErrorString(g_last_err); }

extern "C" int fce_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
    return p.major == 10 ? 1 : 0;
}

extern "C" int fce_conv2d(const fce_conv_desc* d, const void* x, const void* w, const float* bias, const void* res,
                          void* y, void* stream) {
    if (!d || !x || !w || !y) return FCE_ERR_BAD_ARG;
    if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0) return FCE_ERR_BAD_ARG;
    if ((d->k != 1 && d->k != 3) || (d->stride != 1 && d->stride != 2)) return FCE_ERR_UNSUPPORTED;
    if (d->in_layout == FCE_NCHW && d->in_dtype == FCE_BF16) return FCE_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const bool tc_ok = conv2d_tc_supported(d, x, w, res, y);
    const bool weighted = d->weighted != 0 || d->res_up != 0;
    if (weighted) {  // BiFPN-fused epilogue: 1x1 convs on the tcgen05 kernel only
        if (d->res_up && (!res || (d->H & 1) || (d->W & 1))) return FCE_ERR_BAD_ARG;
        if (d->weighted < 0 || d->weighted > 2 || (d->weighted == 2 && !res)) return FCE_ERR_BAD_ARG;
        if (!tc_ok || d->k != 1 || d->impl == 1 || d->out_dtype != FCE_BF16) return FCE_ERR_UNSUPPORTED;
        return conv2d_tc(d, x, w, bias, res, y, st);
    }
    if (d->impl >= 2) return tc_ok ? conv2d_tc(d, x, w, bias, res, y, st) : FCE_ERR_UNSUPPORTED;
    if (d->impl == 0 && tc_ok) return conv2d_tc(d, x, w, bias, res, y, st);
    g_conv_stats[3].fetch_add(1, std::memory_order_relaxed);
    return conv2d_simt(d, x, w, bias, res, y, st);
}

// Would fce_conv2d run this descriptor on the tensor cores?  1 = tcgen05 kernel, 0 = the CUDA-core route (fp32 mode,
// strips, channel counts that are not multiples of 16, ...).  Pure function of the descriptor and pointer alignment: the
// plan compiler calls it to count (or, in strict mode, reject) bf16 convs that would silently leave the tensor pipe.
extern "C" int fce_conv2d_route(const fce_conv_desc* d, const void* x, const void* w, const void* res, const void* y) {
    if (!d) return FCE_ERR_BAD_ARG;
    if (d->impl == 1) return 0;
    return conv2d_tc_supported(d, x, w, res, y) ? 1 : 0;
}

// Launch counters since the last call with reset != 0: out[0] tcgen05 single-CTA launches, [1] tcgen05 CTA-pair launches,
// [2] tcgen05 3x3 strip-kernel launches, [3] CUDA-core (SIMT) launches.
extern "C" int fce_conv_stats(long long* out, int reset) {
    if (!out) return FCE_ERR_BAD_ARG;
    for (int i = 0; i < 4; ++i) out[i] = reset ? g_conv_stats[i].exchange(0) : g_conv_stats[i].load();
    return FCE_OK;
}

// Last conv of a Detect branch with the decode fused into the tcgen05 epilogue (conv_tc.cu): y is the prediction
// tensor [B, 4 + nc, A] fp32, not a logit map.  tcgen05 path only - there is no SIMT twin (the plan compiler keeps
// the unfused fce_conv2d + fce_detect_decode route for every shape this entry point rejects).
extern "C" int fce_conv2d_detect(const fce_conv_desc* d, const fce_detect_epi_desc* e, const void* x, const void* w,
                                 const float* bias, float* y, void* stream) {
    if (!d || !e || !x || !w || !bias || !y) return FCE_ERR_BAD_ARG;
    if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0) return FCE_ERR_BAD_ARG;
    if (e->mode != 1 && e->mode != 2) return FCE_ERR_BAD_ARG;
    if (e->A <= 0 || e->a_base < 0 || e->a_base + d->H * d->W > e->A || e->rows < 4) return FCE_ERR_BAD_ARG;
    if (d->k != 1 || d->stride != 1 || d->act != FCE_ACT_NONE) return FCE_ERR_UNSUPPORTED;
    if (d->in_dtype != FCE_BF16 || d->w_dtype != FCE_BF16 || d->in_layout != FCE_NHWC || d->in_scale != 1.0f)
        return FCE_ERR_UNSUPPORTED;
    if (d->Cin % 16 || d->Cout % 16) return FCE_ERR_UNSUPPORTED;
    if (e->mode == 1 && e->rows < 4 + d->Cout) return FCE_ERR_BAD_ARG;
    if (e->mode == 2 && (e->reg_max != 16 || d->Cout != 64)) return FCE_ERR_UNSUPPORTED;
    if (d->in_pitch % 8 || d->in_off % 8 || (((uintptr_t)x) & 15) || (((uintptr_t)w) & 15) || (((uintptr_t)y) & 3))
        return FCE_ERR_ALIGNMENT;
    fce_conv_desc dd = *d;
    dd.out_dtype = FCE_F32;  // the epilogue writes fp32 straight from the accumulator
    return conv2d_tc(&dd, x, w, bias, nullptr, y, (cudaStream_t)stream, e);
}

#ifdef FCE_DEBUG
// Debug builds only (python fce_yolo_b200/build.py with FCE_DEBUG=1; not declared in the public header): per-role cycle
// accounting and switches that make the tcgen05 kernels skip loads / stores / epilogue math - the results are GARBAGE,
// they exist to time one pipeline stage at a time (tools/conv_bench.py --prof / --dbg).
extern "C" void fce_conv_tc_set_profile(int on) {
    conv_tc_set_profile(on & 15);  // bit 0 profile, bit 1 skip TMA loads, bit 2 skip TMA stores, bit 3 skip epilogue math
}
extern "C" int fce_conv_tc_profile(long long* out, int n) { return out ? conv_tc_profile(out, n) : FCE_ERR_BAD_ARG; }
#endif

// Detect decode: DFL softmax-integral + dist2bbox + stride scaling + class sigmoid, fused
// (reference head.py:149-167, block.py:76-79, tal.py:352-376).  One thread per anchor; the output
// [B, 4+nc, A] is written channel-major so stores coalesce across anchors.
// HBM-bound: algorithmic bytes per image = (4*reg_max+nc)*A*4 read + (4+nc)*A*4 written.
//
// A warp owns 32 consecutive anchors of one level.  Their raw rows (32 x 576 bytes at nc = 80) are first copied into
// the warp's shared-memory slab with fully coalesced 16-byte cp.async copies - 36 in flight per lane - and
// each lane then reads ITS anchor's row from the slab (row pitch padded by 16 bytes: conflict-free).  The first
// version let every lane walk its own 576-byte row in global memory: each warp load touched 32 different lines, and
// the L1 tag stage (71 % busy in the ncu capture), not HBM, set the pace at 3.9 TB/s.
// Math: 64 exponentials + 80 sigmoids per anchor would make the kernel issue-bound with libm-grade expf and IEEE
// division, so it uses ex2.approx / rcp.approx forms (relative error ~1e-6 on the probabilities, < 1e-2 pixel on the
// boxes - two orders inside the fp32-mode parity bound of 1e-4 relative).
#include "common.cuh"

namespace fce {
namespace {

constexpr int NT = 128;
constexpr int MAX_REG = 16;

__global__ void __launch_bounds__(NT) decode_kernel(const fce_decode_desc d, const float* __restrict__ r0,
                                                    const float* __restrict__ r1, const float* __restrict__ r2,
                                                    const float* __restrict__ r3, float* __restrict__ y, int A,
                                                    int blk1, int blk2, int blk3) {
    extern __shared__ float4 dsm[];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // level of this block (levels are concatenated P3, P4, P5: head.py:157) and its first anchor inside the level
    const int bx = blockIdx.x;
    const int lvl = bx < blk1 ? 0 : (bx < blk2 ? 1 : (bx < blk3 ? 2 : 3));
    const int first_blk = lvl == 0 ? 0 : (lvl == 1 ? blk1 : (lvl == 2 ? blk2 : blk3));
    const float* raws[4] = {r0, r1, r2, r3};
    int a_base = 0;
    for (int l = 0; l < lvl; ++l) a_base += d.H[l] * d.W[l];
    const int hw = d.H[lvl] * d.W[lvl];
    const int rem0 = (bx - first_blk) * NT + warp * 32;  // first anchor of this warp inside the level
    if (rem0 >= hw) return;
    const int n_here = min(32, hw - rem0);
    const int R = d.reg_max;
    const int n4 = (4 * R + d.nc) >> 2;   // float4 per raw row
    const int sp4 = n4 + 1;               // padded slab pitch
    const int pitch4 = d.raw_pitch[lvl] >> 2;
    float4* slab = dsm + warp * 32 * sp4;
    {
        const float4* src = reinterpret_cast<const float4*>(raws[lvl]) + ((size_t)b * hw + rem0) * pitch4;
        // asynchronous 16-byte copies: all of a lane's ~36 requests are in flight at once (a register-staged loop
        // serialises on each load's latency - in-order issue stalls at the first dependent shared-memory store).
        // Two commit groups: the DFL columns first, the class columns second, so the box math below runs while the
        // class logits are still landing.
        const uint32_t slab_s = (uint32_t)__cvta_generic_to_shared(slab);
        const int nbox4 = R;  // 4 * R floats of box logits = R float4
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
            const int q0 = part ? nbox4 : 0, qn = part ? n4 - nbox4 : nbox4;  // column range [q0, q0 + qn)
            int an = 0, q = lane;
            while (q >= qn && an < n_here) { q -= qn; ++an; }
            while (an < n_here) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slab_s + (uint32_t)(an * sp4 + q0 + q) * 16u),
                             "l"(src + (size_t)an * pitch4 + q0 + q)
                             : "memory");
                q += 32;
                while (q >= qn && an < n_here) { q -= qn; ++an; }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.wait_group 1;" ::: "memory");
    }
    __syncwarp();
    const bool active = lane < n_here;  // idle lanes of a ragged tail stay for the waits below (their copies feed others)
    const int rem = rem0 + (active ? lane : 0);
    const int a = a_base + rem;
    const int ph = rem / d.W[lvl], pw = rem - ph * d.W[lvl];
    const float4* p4 = slab + (active ? lane : 0) * sp4;
    const float stride = d.stride[lvl];

    float dist[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        float v[MAX_REG];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < MAX_REG; i += 4) {
            if (i < R) {
                const float4 t = p4[(s * R + i) >> 2];
                v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
            }
        }
#pragma unroll
        for (int i = 0; i < MAX_REG; ++i)
            if (i < R) mx = fmaxf(mx, v[i]);
        float sum = 0.f, wsum = 0.f;
#pragma unroll
        for (int i = 0; i < MAX_REG; ++i)
            if (i < R) {
                const float e = __expf(v[i] - mx);
                sum += e;
                wsum = fmaf(e, (float)i, wsum);
            }
        dist[s] = __fdividef(wsum, sum);
    }
    const float ax = (float)pw + 0.5f, ay = (float)ph + 0.5f;
    const float x1 = ax - dist[0], y1 = ay - dist[1], x2 = ax + dist[2], y2 = ay + dist[3];
    const size_t yb = (size_t)b * (4 + d.nc) * A + a;
    if (active) {
        y[yb] = (x1 + x2) * 0.5f * stride;
        y[yb + (size_t)A] = (y1 + y2) * 0.5f * stride;
        y[yb + (size_t)2 * A] = (x2 - x1) * stride;
        y[yb + (size_t)3 * A] = (y2 - y1) * stride;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");  // class columns (every lane waits on its own copies ...
    __syncwarp();                                          // ... and then on everybody else's)
    if (!active) return;
    const float4* pc = p4 + R;
    for (int c = 0; c < d.nc; c += 4) {
        const float4 t = pc[c >> 2];
        const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (c + j < d.nc) y[yb + (size_t)(4 + c + j) * A] = sigmoid_f(tv[j]);
    }
}


// ------------------------------------------------------------------------------------------------ scale_boxes
// ops.scale_boxes + clip_boxes (ultralytics/utils/ops.py:102-134, 152-177) as DetectionPredictor.construct_result
// applies them (models/yolo/detect/predict.py:109-122): detections of the letterboxed image -> original image
// coordinates, in place on the padded [B, max_det, 6] buffer.  meta[b] = {gain, pad_x, pad_y, w0, h0} fp32.  Same fp32
// operations in the same order as the reference's CPU path (subtract pad, IEEE divide by gain, clamp).
__global__ void __launch_bounds__(NT) scale_boxes_kernel(float* __restrict__ det, const int32_t* __restrict__ count,
                                                         const float* __restrict__ meta, int max_det) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * NT + threadIdx.x;
    if (q >= max_det || q >= count[b]) return;
    const float gain = meta[b * 5], px = meta[b * 5 + 1], py = meta[b * 5 + 2], w0 = meta[b * 5 + 3], h0 = meta[b * 5 + 4];
    float4* p = reinterpret_cast<float4*>(det + ((size_t)b * max_det + q) * 6);  // 24-byte rows: 8-byte aligned only
    float* f = reinterpret_cast<float*>(p);
    const float x1 = __fdiv_rn(__fsub_rn(f[0], px), gain), y1 = __fdiv_rn(__fsub_rn(f[1], py), gain);
    const float x2 = __fdiv_rn(__fsub_rn(f[2], px), gain), y2 = __fdiv_rn(__fsub_rn(f[3], py), gain);
    f[0] = fminf(fmaxf(x1, 0.f), w0);
    f[1] = fminf(fmaxf(y1, 0.f), h0);
    f[2] = fminf(fmaxf(x2, 0.f), w0);
    f[3] = fminf(fmaxf(y2, 0.f), h0);
}

}  // namespace
}  // namespace fce

using namespace fce;

extern "C" int fce_detect_decode(const fce_decode_desc* d, const float* raw0, const float* raw1, const float* raw2,
                                 const float* raw3, float* y, void* stream) {
    if (!d || !raw0 || !y || d->B <= 0 || d->nl < 1 || d->nl > 4) return FCE_ERR_BAD_ARG;
    if (d->reg_max > MAX_REG || (d->reg_max % 4) || (d->nc % 4)) return FCE_ERR_UNSUPPORTED;
    const float* r[4] = {raw0, raw1, raw2, raw3};
    int A = 0;
    for (int i = 0; i < d->nl; ++i) {
        if (!r[i] || d->H[i] <= 0 || d->W[i] <= 0) return FCE_ERR_BAD_ARG;
        if ((d->raw_pitch[i] % 4) || (((uintptr_t)r[i]) & 15)) return FCE_ERR_ALIGNMENT;
        A += d->H[i] * d->W[i];
    }
    int blk[5] = {0, 0, 0, 0, 0};  // first block of each level
    for (int i = 0; i < d->nl; ++i) blk[i + 1] = blk[i] + (d->H[i] * d->W[i] + NT - 1) / NT;
    for (int i = d->nl; i < 4; ++i) blk[i + 1] = blk[d->nl];
    if (d->B > 65535) return FCE_ERR_UNSUPPORTED;
    const size_t smem = (size_t)(NT / 32) * 32 * ((4 * d->reg_max + d->nc) / 4 + 1) * sizeof(float4);
    if (smem > 200 * 1024) return FCE_ERR_UNSUPPORTED;
    static DeviceOnce attr_once;  // per-device attribute
    int dev_ = 0;
    if (smem > 48 * 1024 && attr_once.pending(&dev_)) {
        cudaError_t e = cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            set_cuda_error(e);
            return FCE_ERR_CUDA;
        }
        attr_once.done(dev_);
    }
    dim3 grid(blk[d->nl], d->B);
    decode_kernel<<<grid, NT, smem, (cudaStream_t)stream>>>(*d, raw0, raw1, raw2, raw3, y, A, blk[1], blk[2], blk[3]);
    return check_launch();
}

extern "C" int fce_scale_boxes(float* det, const int32_t* count, const float* meta, int32_t B, int32_t max_det,
                               void* stream) {
    if (!det || !count || !meta || B <= 0 || max_det <= 0) return FCE_ERR_BAD_ARG;
    if (B > 65535) return FCE_ERR_UNSUPPORTED;
    dim3 grid((max_det + NT - 1) / NT, B);
    scale_boxes_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(det, count, meta, max_det);
    return check_launch();
}

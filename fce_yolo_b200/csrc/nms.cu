// Batched GPU NMS with the exact semantics of the reference's non_max_suppression
// (ultralytics/utils/nms.py:13-166) + torchvision.ops.nms (:151-154): one CTA per image runs
//   1. candidate generation in anchor order (best class, or multi-label (anchor, class) pairs in
//      torch.where's row-major order), optional class filter;
//   2. if more than max_nms candidates: exact radix-select of the max_nms best (ties -> lowest index);
//   3. stable LSD radix sort by descending score (ties -> ascending index, = torchvision's stable sort);
//   4. greedy suppression against the kept list, 256 sorted candidates at a time, stopping at max_det
//      (the reference truncates after NMS, nms.py:157 - identical result, bounded work);
//   5. emission of [max_det, 6] rows, int64 anchor indices and the count.
// All box arithmetic uses explicit round-to-nearest fp32 intrinsics so nothing is contracted into FMAs:
// keep indices and class ids are bit-exact against the CPU reference, including the fp32 class offset
// (cls * max_wh added to the coordinates, nms.py:143,149).
// Integer/bandwidth work: no tensor cores.
#include <atomic>

#include "common.cuh"

namespace fce {
namespace {

constexpr int NT = 512;
constexpr int NW = NT / 32;
constexpr int CH = 256;  // candidates resolved per suppression round

__device__ __forceinline__ uint32_t desc_key(float s) {
    // ascending order of the returned key == descending order of the score
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ordered ascending
    return ~u;
}

// exclusive block scan of one int per thread; returns the thread's offset, total through *total
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // warp_sums reuse
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < NW ? warp_sums[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < NW) warp_sums[lane] = winc - w;
        if (lane == NW - 1) warp_sums[NW] = winc;
    }
    __syncthreads();
    *total = warp_sums[NW];
    return warp_sums[warp] + inc - v;
}

struct Box {
    float x1, y1, x2, y2;
};

__device__ __forceinline__ Box load_box(const float* pb, int A, int a) {
    // xywh -> xyxy exactly as ops.py:234-239 (wh / 2, xy -/+ half)
    const float cx = pb[a], cy = pb[(size_t)A + a], w = pb[(size_t)2 * A + a], h = pb[(size_t)3 * A + a];
    const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
    Box r;
    r.x1 = __fsub_rn(cx, hw);
    r.y1 = __fsub_rn(cy, hh);
    r.x2 = __fadd_rn(cx, hw);
    r.y2 = __fadd_rn(cy, hh);
    return r;
}

__device__ __forceinline__ bool suppresses(const Box& a, float area_a, const Box& b, float area_b, float thr, int ge) {
    const float xx1 = fmaxf(a.x1, b.x1), yy1 = fmaxf(a.y1, b.y1);
    const float xx2 = fminf(a.x2, b.x2), yy2 = fminf(a.y2, b.y2);
    const float w = fmaxf(__fsub_rn(xx2, xx1), 0.f), h = fmaxf(__fsub_rn(yy2, yy1), 0.f);
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
    return ge ? (ovr >= thr) : (ovr > thr);
}

__global__ void __launch_bounds__(NT) nms_kernel(const fce_nms_desc d, const float* __restrict__ pred,
                                                 const int32_t* __restrict__ classes, float* __restrict__ det,
                                                 int64_t* __restrict__ keep, int32_t* __restrict__ count,
                                                 uint32_t* __restrict__ ws, int cap, float thr_f, int thr_ge) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int warp_sums[NW + 1];
    __shared__ int hist[256];
    __shared__ int digit_base[256];
    __shared__ int s_n, s_kept, s_prefix_digit, s_need;
    __shared__ uint32_t alive[CH / 32];
    __shared__ uint32_t sup[CH][CH / 32];
    __shared__ Box cbox[CH];
    __shared__ float carea[CH];
    __shared__ int warp_hist[NW][256];

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = d.A, nc = d.nc;
    const float* pb = pred + (size_t)b * (4 + nc) * A;
    const float* ps = pb + (size_t)4 * A;
    uint32_t* keyA = ws + (size_t)b * cap * 4;
    uint32_t* idxA = keyA + cap;
    uint32_t* keyB = idxA + cap;
    uint32_t* idxB = keyB + cap;
    const float conf = d.conf_thres;

    // kept boxes live in dynamic smem: [max_det] Box + area + idx
    Box* kbox = reinterpret_cast<Box*>(smem_raw);
    float* karea = reinterpret_cast<float*>(kbox + d.max_det);
    uint32_t* kidx = reinterpret_cast<uint32_t*>(karea + d.max_det);
    float* kscore = reinterpret_cast<float*>(kidx + d.max_det);

    // ---------------- 1. candidates, in (anchor, class) order ----------------
    int n = 0;
    for (int a0 = 0; a0 < A; a0 += NT) {
        const int a = a0 + tid;
        int cnt = 0;
        float best = -INFINITY;
        int bestc = 0;
        if (a < A) {
            if (d.multi_label) {
                for (int j = 0; j < nc; ++j) {
                    const float s = ps[(size_t)j * A + a];
                    if (s > conf) {
                        bool ok = true;
                        if (d.n_classes > 0) {
                            ok = false;
                            for (int q = 0; q < d.n_classes; ++q) ok |= (classes[q] == j);
                        }
                        cnt += ok;
                    }
                }
            } else {
                for (int j = 0; j < nc; ++j) {
                    const float s = ps[(size_t)j * A + a];
                    if (s > best) { best = s; bestc = j; }
                }
                bool ok = best > conf;
                if (ok && d.n_classes > 0) {
                    ok = false;
                    for (int q = 0; q < d.n_classes; ++q) ok |= (classes[q] == bestc);
                }
                cnt = ok;
            }
        }
        int total;
        int off = n + block_excl_scan(cnt, warp_sums, &total);
        if (cnt) {
            if (d.multi_label) {
                for (int j = 0; j < nc; ++j) {
                    const float s = ps[(size_t)j * A + a];
                    if (s > conf) {
                        bool ok = true;
                        if (d.n_classes > 0) {
                            ok = false;
                            for (int q = 0; q < d.n_classes; ++q) ok |= (classes[q] == j);
                        }
                        if (ok) {
                            keyA[off] = desc_key(s);
                            idxA[off] = (uint32_t)a * nc + j;
                            ++off;
                        }
                    }
                }
            } else {
                keyA[off] = desc_key(best);
                idxA[off] = (uint32_t)a * nc + bestc;
            }
        }
        n += total;
    }
    __syncthreads();
    uint32_t *kin = keyA, *iin = idxA, *kout = keyB, *iout = idxB;

    // ---------------- 2. exact top-max_nms selection (nms.py:136-140) ----------------
    if (n > d.max_nms) {
        uint32_t prefix = 0, mask = 0;
        int need = d.max_nms;  // rank (1-based) of the last key to keep, among keys matching the prefix
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += NT) hist[i] = 0;
            __syncthreads();
            for (int i = tid; i < n; i += NT) {
                const uint32_t k = kin[i];
                if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255], 1);
            }
            __syncthreads();
            if (tid == 0) {
                int acc = 0, dg = 0;
                for (; dg < 256; ++dg) {
                    if (acc + hist[dg] >= need) break;
                    acc += hist[dg];
                }
                s_prefix_digit = dg;
                s_need = need - acc;
            }
            __syncthreads();
            prefix |= ((uint32_t)s_prefix_digit) << shift;
            mask |= 255u << shift;
            need = s_need;
            __syncthreads();
        }
        // prefix == threshold key T; keep all keys < T and the first `need` keys == T (index order)
        const uint32_t T = prefix;
        int out_n = 0, eq_seen = 0;
        for (int i0 = 0; i0 < n; i0 += NT) {
            const int i = i0 + tid;
            uint32_t k = 0;
            int is_lt = 0, is_eq = 0;
            if (i < n) {
                k = kin[i];
                is_lt = k < T;
                is_eq = k == T;
            }
            int tot_eq;
            const int eq_rank = eq_seen + block_excl_scan(is_eq, warp_sums, &tot_eq);
            const int take = is_lt | (is_eq && eq_rank < need);
            int tot;
            const int off = out_n + block_excl_scan(take, warp_sums, &tot);
            if (take) {
                kout[off] = k;
                iout[off] = iin[i];
            }
            out_n += tot;
            eq_seen += tot_eq;
        }
        __syncthreads();
        n = out_n;
        uint32_t* t;
        t = kin; kin = kout; kout = t;
        t = iin; iin = iout; iout = t;
    }

    // ---------------- 3. stable LSD radix sort, 4 x 8 bits ----------------
    if (n > 1) {
        for (int shift = 0; shift < 32; shift += 8) {
            for (int i = tid; i < 256; i += NT) hist[i] = 0;
            __syncthreads();
            for (int i = tid; i < n; i += NT) atomicAdd(&hist[(kin[i] >> shift) & 255], 1);
            __syncthreads();
            if (warp == 0) {  // exclusive scan of 256 bins by one warp (8 bins per lane)
                int loc[8], s = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { loc[j] = hist[lane * 8 + j]; s += loc[j]; }
                int inc = s;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                int base = inc - s;
#pragma unroll
                for (int j = 0; j < 8; ++j) { digit_base[lane * 8 + j] = base; base += loc[j]; }
            }
            __syncthreads();
            const bool skip = false;
            (void)skip;
            for (int i0 = 0; i0 < n; i0 += NT) {
                const int i = i0 + tid;
                const bool valid = i < n;
                uint32_t k = 0, id = 0;
                int dg = 256 + 0;  // invalid lanes use a private pseudo digit
                if (valid) { k = kin[i]; id = iin[i]; dg = (k >> shift) & 255; }
                for (int j = lane; j < 256; j += 32) warp_hist[warp][j] = 0;
                __syncwarp();
                const uint32_t peers = __match_any_sync(0xffffffffu, dg);
                const int rank_in_warp = __popc(peers & ((1u << lane) - 1));
                if (valid && rank_in_warp == 0) warp_hist[warp][dg] = __popc(peers);
                __syncthreads();
                // per digit: exclusive prefix over warps, then advance the running base
                if (tid < 256) {
                    int run = digit_base[tid];
#pragma unroll
                    for (int wq = 0; wq < NW; ++wq) {
                        const int c = warp_hist[wq][tid];
                        warp_hist[wq][tid] = run;
                        run += c;
                    }
                    digit_base[tid] = run;
                }
                __syncthreads();
                if (valid) {
                    const int pos = warp_hist[warp][dg] + rank_in_warp;
                    kout[pos] = k;
                    iout[pos] = id;
                }
                __syncthreads();
            }
            uint32_t* t;
            t = kin; kin = kout; kout = t;
            t = iin; iin = iout; iout = t;
        }
    }
    __syncthreads();

    // ---------------- 4. greedy suppression against the kept list ----------------
    if (tid == 0) s_kept = 0;
    __syncthreads();
    const float offs_scale = d.agnostic ? 0.f : d.max_wh;
    for (int c0 = 0; c0 < n; c0 += CH) {
        const int kept0 = s_kept;
        if (kept0 >= d.max_det) break;
        const int m = min(CH, n - c0);
        bool dead = true;
        Box mine = {0, 0, 0, 0};
        float marea = 0.f;
        if (tid < m) {
            const uint32_t id = iin[c0 + tid];
            const int a = id / nc, cls = id - a * nc;
            Box r = load_box(pb, A, a);
            const float off = __fmul_rn((float)cls, offs_scale);  // nms.py:143
            mine.x1 = __fadd_rn(r.x1, off); mine.y1 = __fadd_rn(r.y1, off);  // nms.py:149
            mine.x2 = __fadd_rn(r.x2, off); mine.y2 = __fadd_rn(r.y2, off);
            marea = __fmul_rn(__fsub_rn(mine.x2, mine.x1), __fsub_rn(mine.y2, mine.y1));
            cbox[tid] = mine;
            carea[tid] = marea;
            dead = false;
            for (int q = 0; q < kept0; ++q)
                if (suppresses(kbox[q], karea[q], mine, marea, thr_f, thr_ge)) { dead = true; break; }
        }
        if (tid < CH / 32) alive[tid] = 0;
        __syncthreads();
        if (tid < m && !dead) atomicOr(&alive[tid >> 5], 1u << (tid & 31));
        // sup[i] = bitmask of later chunk members j > i that candidate i would suppress
        for (int e = tid; e < CH * (CH / 32); e += NT) {
            const int i = e / (CH / 32), wq = e % (CH / 32);
            uint32_t bits = 0;
            if (i < m) {
                const Box bi = cbox[i];
                const float ai = carea[i];
                for (int t = 0; t < 32; ++t) {
                    const int j = wq * 32 + t;
                    if (j > i && j < m && suppresses(bi, ai, cbox[j], carea[j], thr_f, thr_ge)) bits |= 1u << t;
                }
            }
            sup[i][wq] = bits;
        }
        __syncthreads();
        if (warp == 0) {  // sequential resolve, 8 words handled by lanes 0..7
            uint32_t al = lane < CH / 32 ? alive[lane] : 0;
            int kept = kept0;
            for (int i = 0; i < m && kept < d.max_det; ++i) {
                const uint32_t wbits = __shfl_sync(0xffffffffu, al, i >> 5);
                if ((wbits >> (i & 31)) & 1u) {
                    if (lane == 0) {
                        kbox[kept] = cbox[i];
                        karea[kept] = carea[i];
                        kidx[kept] = iin[c0 + i];
                        kscore[kept] = 0.f;
                    }
                    ++kept;
                    if (lane < CH / 32) al &= ~sup[i][lane];
                }
            }
            if (lane == 0) s_kept = kept;
        }
        __syncthreads();
    }
    __syncthreads();

    // ---------------- 5. emit ----------------
    const int kept = s_kept;
    if (tid == 0) count[b] = kept;
    for (int q = tid; q < d.max_det; q += NT) {
        float* o = det + ((size_t)b * d.max_det + q) * 6;
        if (q < kept) {
            const uint32_t id = kidx[q];
            const int a = id / nc, cls = id - a * nc;
            const Box r = load_box(pb, A, a);
            o[0] = r.x1; o[1] = r.y1; o[2] = r.x2; o[3] = r.y2;
            o[4] = ps[(size_t)cls * A + a];
            o[5] = (float)cls;
            keep[(size_t)b * d.max_det + q] = a;
        } else {
            o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.f;
            keep[(size_t)b * d.max_det + q] = -1;
        }
    }
}

inline size_t kept_smem(int max_det) { return (size_t)max_det * (sizeof(Box) + 3 * sizeof(float)); }

}  // namespace
}  // namespace fce

using namespace fce;

extern "C" size_t fce_nms_workspace(const fce_nms_desc* d) {
    if (!d) return 0;
    const size_t cap = d->multi_label ? (size_t)d->A * d->nc : (size_t)d->A;
    return (size_t)d->B * cap * 4 * sizeof(uint32_t);
}

extern "C" int fce_nms(const fce_nms_desc* d, const float* pred, const int32_t* classes, float* det, int64_t* keep,
                       int32_t* count, void* ws, size_t ws_bytes, void* stream) {
    if (!d || !pred || !det || !keep || !count || !ws) return FCE_ERR_BAD_ARG;
    if (d->B <= 0 || d->A <= 0 || d->nc <= 0 || d->max_det <= 0 || d->max_nms <= 0) return FCE_ERR_BAD_ARG;
    if (d->n_classes > 0 && !classes) return FCE_ERR_BAD_ARG;
    if (!(d->conf_thres >= 0.f && d->conf_thres <= 1.f) || !(d->iou_thres >= 0.0 && d->iou_thres <= 1.0))
        return FCE_ERR_BAD_ARG;  // nms.py:59-60 asserts
    if ((long long)d->A * d->nc >= (1ll << 31)) return FCE_ERR_UNSUPPORTED;
    if (ws_bytes < fce_nms_workspace(d)) return FCE_ERR_WORKSPACE;
    const size_t smem = kept_smem(d->max_det);
    if (smem > 96 * 1024) return FCE_ERR_UNSUPPORTED;
    static std::atomic<bool> attr_done{false};
    if (!attr_done.load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess) { set_cuda_error(e); return FCE_ERR_CUDA; }
        attr_done.store(true, std::memory_order_release);
    }
    // torchvision compares the fp32 IoU with a C double threshold: reproduce with an fp32 compare
    const float tf = (float)d->iou_thres;
    const int ge = ((double)tf > d->iou_thres) ? 1 : 0;
    const int cap = d->multi_label ? d->A * d->nc : d->A;
    nms_kernel<<<d->B, NT, smem, (cudaStream_t)stream>>>(*d, pred, classes, det, keep, count, (uint32_t*)ws, cap, tf, ge);
    return check_launch();
}

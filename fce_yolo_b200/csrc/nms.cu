// Batched GPU NMS with the exact semantics of the reference's non_max_suppression
// (ultralytics/utils/nms.py:13-166) + torchvision.ops.nms (:151-154).
//
// Kernel 1 (fully parallel, HBM-bound: one coalesced pass over pred): every candidate gets a 32-bit sort key
// (descending score); non-candidates (score <= conf, class filtered out) get the sentinel 0xFFFFFFFF.
//   predict mode : one key per anchor (best class, nms.py:120-122) + the class id;
//   multi-label  : one key per (class, anchor) score (nms.py:115-116), same [nc, A] layout as pred.
// Kernel 2 (one CTA per image) works on 64-bit composites  key << 32 | (anchor * nc + class), which are UNIQUE
// and whose ascending order is exactly the reference's order: descending score, ties by ascending
// (anchor, class) = torch.where's row-major order / torchvision's stable sort.  In rounds of 1024 candidates:
//   a. exact radix select of the 1024 smallest unprocessed composites (8-bit digits, shared-memory histograms);
//   b. unordered compaction into shared memory + bitonic sort (uniqueness makes stability moot);
//   c. greedy suppression against the kept list, 256 sorted candidates at a time, stopping at max_det
//      (the reference truncates after NMS, nms.py:157 - identical result, bounded work) or after max_nms
//      candidates (nms.py:136-140 keeps only the max_nms best).
// All box arithmetic uses explicit round-to-nearest fp32 intrinsics so nothing is contracted into FMAs:
// keep indices and class ids are bit-exact against the CPU reference, including the fp32 class offset
// (cls * max_wh added to the coordinates, nms.py:143,149).
// Integer/bandwidth work: no tensor cores.
#include <atomic>

#include "common.cuh"

namespace fce {
namespace {

constexpr int NT = 1024;
constexpr int TSEL = 1024;  // candidates selected + sorted per round
constexpr int CH = 256;     // candidates resolved per suppression step
constexpr uint32_t INVALID_KEY = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t desc_key(float s) {
    // ascending order of the returned key == descending order of the score
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ordered ascending
    return ~u;
}

__device__ __forceinline__ bool class_allowed(const int32_t* classes, int n_classes, int c) {
    if (n_classes <= 0) return true;
    bool ok = false;
    for (int q = 0; q < n_classes; ++q) ok |= (classes[q] == c);
    return ok;
}

// ---------------------------------------------------------------- kernel 1: keys
__global__ void __launch_bounds__(256) nms_keys_single(const fce_nms_desc d, const float* __restrict__ pred,
                                                       const int32_t* __restrict__ classes,
                                                       uint32_t* __restrict__ keys, uint32_t* __restrict__ cls) {
    const int b = blockIdx.y;
    const int a = blockIdx.x * 256 + threadIdx.x;
    if (a >= d.A) return;
    const float* ps = pred + ((size_t)b * (4 + d.nc) + 4) * d.A + a;
    float best = -INFINITY;
    int bestc = 0;
#pragma unroll 8
    for (int j = 0; j < d.nc; ++j) {
        const float s = __ldg(ps + (size_t)j * d.A);
        if (s > best) {  // first maximum wins, like torch.max (nms.py:120)
            best = s;
            bestc = j;
        }
    }
    const bool ok = best > d.conf_thres && class_allowed(classes, d.n_classes, bestc);
    keys[(size_t)b * d.A + a] = ok ? desc_key(best) : INVALID_KEY;
    cls[(size_t)b * d.A + a] = (uint32_t)bestc;
}

__global__ void __launch_bounds__(256) nms_keys_multi(const fce_nms_desc d, const float* __restrict__ pred,
                                                      const int32_t* __restrict__ classes,
                                                      uint32_t* __restrict__ keys) {
    const int b = blockIdx.z, j = blockIdx.y;
    const int a = blockIdx.x * 256 + threadIdx.x;
    if (a >= d.A) return;
    const float s = __ldg(pred + ((size_t)b * (4 + d.nc) + 4 + j) * d.A + a);
    const bool ok = s > d.conf_thres && class_allowed(classes, d.n_classes, j);
    keys[((size_t)b * d.nc + j) * d.A + a] = ok ? desc_key(s) : INVALID_KEY;
}

// ---------------------------------------------------------------- kernel 2: select / sort / suppress
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
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    // Same decision as the reference's  inter / uni > thr  in IEEE fp32, without paying for the IEEE division on
    // every pair: disjoint boxes (the bulk) give exactly 0, and a 2-ulp approximate quotient decides every pair
    // that is not within 1e-4 (relative) of the threshold; only those take the exact division.
    if (inter == 0.f && uni > 0.f) return ge ? (0.f >= thr) : (0.f > thr);
    const float q = __fdividef(inter, uni);
    if (fabsf(q - thr) > 1e-4f * fmaxf(thr, 1e-3f)) return q > thr;  // false for NaN: falls through to the exact path
    const float ovr = __fdiv_rn(inter, uni);
    return ge ? (ovr >= thr) : (ovr > thr);
}

// Visits every composite slot of image b: f(valid, composite) runs on ALL 32 lanes of every warp the same number of times (loop bounds
// rounded up to a multiple of 32), so f may use full-mask warp collectives - the histogram and compaction steps
// aggregate their shared-memory atomics per warp (scores of one image share their leading digits: un-aggregated,
// thousands of atomics serialise on one or two bins).
template <typename F>
__device__ __forceinline__ void for_each_slot(const fce_nms_desc& d, const uint32_t* __restrict__ keys,
                                              const uint32_t* __restrict__ cls, F&& f) {
    const int A = d.A, nc = d.nc;
    const int A_up = (A + 31) & ~31;
    if (d.multi_label) {
        for (int j = 0; j < nc; ++j) {
            const uint32_t* kj = keys + (size_t)j * A;
            for (int a = threadIdx.x; a < A_up; a += NT) {
                const uint32_t k = a < A ? kj[a] : INVALID_KEY;
                f(k != INVALID_KEY, ((unsigned long long)k << 32) | (uint32_t)(a * nc + j));
            }
        }
    } else {
        for (int a = threadIdx.x; a < A_up; a += NT) {
            const uint32_t k = a < A ? keys[a] : INVALID_KEY;
            const uint32_t c = a < A ? cls[a] : 0u;
            f(k != INVALID_KEY, ((unsigned long long)k << 32) | (uint32_t)(a * nc + (int)c));
        }
    }
}

__global__ void __launch_bounds__(NT) nms_kernel(const fce_nms_desc d, const float* __restrict__ pred,
                                                 const uint32_t* __restrict__ keys_all,
                                                 const uint32_t* __restrict__ cls_all, float* __restrict__ det,
                                                 int64_t* __restrict__ keep, int32_t* __restrict__ count, float thr_f,
                                                 int thr_ge) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long scomp[TSEL];
    __shared__ int hist[256];
    __shared__ int s_digit, s_need, s_cnt, s_kept, s_valid, s_full;
    __shared__ uint32_t alive[CH / 32];
    __shared__ uint32_t sup[CH][CH / 32];
    __shared__ Box cbox[CH];
    __shared__ float carea[CH];
    __shared__ unsigned short korder[CH];

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = d.A, nc = d.nc;
    const float* pb = pred + (size_t)b * (4 + nc) * A;
    const float* ps = pb + (size_t)4 * A;
    const size_t n_keys = d.multi_label ? (size_t)A * nc : (size_t)A;
    const uint32_t* keys = keys_all + (size_t)b * n_keys;
    const uint32_t* cls = d.multi_label ? nullptr : cls_all + (size_t)b * A;

    // kept boxes live in dynamic smem: [max_det] Box + area + idx
    Box* kbox = reinterpret_cast<Box*>(smem_raw);
    float* karea = reinterpret_cast<float*>(kbox + d.max_det);
    uint32_t* kidx = reinterpret_cast<uint32_t*>(karea + d.max_det);

    // number of candidates
    if (tid == 0) { s_valid = 0; s_kept = 0; }
    __syncthreads();
    {
        int local = 0;  // per warp (every lane holds the same count)
        for_each_slot(d, keys, cls, [&](bool ok, unsigned long long) { local += __popc(__ballot_sync(0xffffffffu, ok)); });
        if (lane == 0 && local) atomicAdd(&s_valid, local);
    }
    __syncthreads();
    const int limit = min(s_valid, d.max_nms);  // nms.py:136-140: only the max_nms best take part
    const float offs_scale = d.agnostic ? 0.f : d.max_wh;

    unsigned long long bnd = 0;  // composites <= bnd have been processed (valid once processed > 0)
    int processed = 0;
    while (processed < limit && s_kept < d.max_det) {
        const int target = min(TSEL, limit - processed);
        const bool have_bnd = processed > 0;
        // ---- a. exact radix select: the composite of rank `target` among the unprocessed ones
        unsigned long long prefix = 0, mask = 0;
        int need = target;
        // fast path (the usual predict case: a few hundred candidates): everything still unprocessed fits one
        // round, so the rank-`target` composite need not be found - take them all
        // (<= target, not <= TSEL: when max_nms cuts the list, limit - processed < remaining candidates and the BEST
        // `target` of them must be selected - nms.py:136-140 keeps the top max_nms by score)
        const bool take_all = s_valid - processed <= target;
        if (take_all) prefix = ~0ull;
        for (int shift = 56; shift >= 0 && !take_all; shift -= 8) {
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            for_each_slot(d, keys, cls, [&](bool ok, unsigned long long c) {
                ok = ok && (!have_bnd || c > bnd) && (c & mask) == prefix;
                // one atomic per distinct digit per warp: lanes with the same digit elect their lowest lane
                const int bin = ok ? (int)((c >> shift) & 255) : 256 + lane;
                const uint32_t peers = __match_any_sync(0xffffffffu, bin);
                if (ok && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
            });
            __syncthreads();
            if (warp == 0) {  // first digit whose cumulative count reaches `need` (8 bins per lane)
                int loc[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { loc[j] = hist[lane * 8 + j]; sum += loc[j]; }
                int inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                int before = inc - sum;  // candidates in lower digits
                const bool mine = before < need && inc >= need;
                if (mine) {
                    int dg = 0;
                    for (; dg < 8; ++dg) {
                        if (before + loc[dg] >= need) break;
                        before += loc[dg];
                    }
                    s_digit = lane * 8 + dg;
                    s_need = need - before;
                    s_full = loc[dg] == need - before;  // the whole bin is wanted: no need to look at lower digits
                }
            }
            __syncthreads();
            prefix |= (unsigned long long)s_digit << shift;
            mask |= 255ull << shift;
            need = s_need;
            const bool full = s_full != 0;
            __syncthreads();
            if (full) {  // distinct scores: reached after the four score digits, the four index digits are skipped
                prefix |= shift ? ((1ull << shift) - 1ull) : 0ull;
                break;
            }
        }
        const unsigned long long thr = prefix;
        // ---- b. compaction (any order) + bitonic sort
        if (tid == 0) s_cnt = 0;
        for (int i = tid; i < TSEL; i += NT) scomp[i] = ~0ull;
        __syncthreads();
        for_each_slot(d, keys, cls, [&](bool ok, unsigned long long c) {
            ok = ok && (!have_bnd || c > bnd) && c <= thr;
            const uint32_t sel = __ballot_sync(0xffffffffu, ok);
            if (sel) {  // warp-uniform: one atomic per warp reserves the slots
                int base_pos = 0;
                if (lane == __ffs(sel) - 1) base_pos = atomicAdd(&s_cnt, __popc(sel));
                base_pos = __shfl_sync(0xffffffffu, base_pos, __ffs(sel) - 1);
                const int pos = base_pos + __popc(sel & ((1u << lane) - 1u));
                if (ok && pos < TSEL) scomp[pos] = c;
            }
        });
        __syncthreads();
        int n_sort = 2;  // bitonic network over the next power of two >= target (the tail is ~0 padding)
        while (n_sort < target) n_sort <<= 1;
        for (int k = 2; k <= n_sort; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < n_sort; i += NT) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long x = scomp[i], y = scomp[ixj];
                        const bool asc = (i & k) == 0;
                        if ((x > y) == asc) {
                            scomp[i] = y;
                            scomp[ixj] = x;
                        }
                    }
                }
                __syncthreads();
            }
        }
        // ---- c. greedy suppression against the kept list
        for (int c0 = 0, m = 0; c0 < target; c0 += m) {
            const int kept0 = s_kept;
            if (kept0 >= d.max_det) break;
            // chunk length: no longer than needed to fill the remaining max_det slots with ~2x head-room (rounded to
            // whole 32-candidate words) - the all-pairs work of a chunk grows with its square, and after the first
            // chunk only a few dozen boxes are usually missing
            m = min(min(CH, target - c0), (2 * (d.max_det - kept0) + 63) & ~31);
            bool dead = true;
            if (tid < m) {
                const uint32_t id = (uint32_t)scomp[c0 + tid];
                const int a = id / nc, cl = id - a * nc;
                const Box r = load_box(pb, A, a);
                const float off = __fmul_rn((float)cl, offs_scale);  // nms.py:143
                Box mine;
                mine.x1 = __fadd_rn(r.x1, off); mine.y1 = __fadd_rn(r.y1, off);  // nms.py:149
                mine.x2 = __fadd_rn(r.x2, off); mine.y2 = __fadd_rn(r.y2, off);
                const float marea = __fmul_rn(__fsub_rn(mine.x2, mine.x1), __fsub_rn(mine.y2, mine.y1));
                cbox[tid] = mine;
                carea[tid] = marea;
            }
            if (tid < CH / 32) alive[tid] = 0;
            __syncthreads();
            // candidate i (0..m) against the kept list: 4 threads per candidate split the list
            {
                const int i = tid >> 2, part = tid & 3;
                bool hit = false;
                if (i < m) {
                    const Box mine = cbox[i];
                    const float marea = carea[i];
                    for (int q = part; q < kept0; q += 4)
                        if (suppresses(kbox[q], karea[q], mine, marea, thr_f, thr_ge)) { hit = true; break; }
                }
                hit |= __shfl_xor_sync(0xffffffffu, hit, 1);
                hit |= __shfl_xor_sync(0xffffffffu, hit, 2);
                dead = hit;
                if (i < m && part == 0 && !dead) atomicOr(&alive[i >> 5], 1u << (i & 31));
            }
            // sup[i] = bitmask of later chunk members j > i that candidate i would suppress
            for (int e = tid; e < CH * (CH / 32); e += NT) {
                const int i = e / (CH / 32), wq = e % (CH / 32);
                uint32_t bits = 0;
                if (i < m && wq * 32 + 31 > i) {
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
            if (warp == 0) {
                // Greedy resolve of the chunk, one 32-candidate word at a time.  Lane j stands for candidate 32w + j
                // and holds its suppression row (8 words).  Inside a word the scan is inherently serial, but the
                // dependent chain per candidate is only test-bit / mask / and-not (the rows are fetched by 32
                // independent shuffles up front); across words the rows of the kept candidates are or-reduced over
                // the warp.  ~700 cycles per word instead of ~450 per kept box for a per-box loop.
                uint32_t al = lane < CH / 32 ? alive[lane] : 0;  // lane w holds alive word w
                int kept = kept0;
                const int n_words = (m + 31) >> 5;
                for (int w = 0; w < n_words && kept < d.max_det; ++w) {
                    uint32_t row[CH / 32];
#pragma unroll
                    for (int q = 0; q < CH / 32; ++q) row[q] = sup[w * 32 + lane][q];
                    uint32_t K = __shfl_sync(0xffffffffu, al, w);  // alive candidates of this word (uniform)
                    uint32_t intra = 0u;
#pragma unroll
                    for (int q = 0; q < CH / 32; ++q)
                        if (q == w) intra = row[q];
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        const uint32_t rt = __shfl_sync(0xffffffffu, intra, t);  // independent of K
                        const uint32_t on = 0u - ((K >> t) & 1u);
                        K &= ~(rt & on);
                    }
                    const int room = d.max_det - kept;
                    if (__popc(K) > room) K &= (1u << __fns(K, 0, room + 1)) - 1u;  // keep the first `room` only
                    const bool mine = (K >> lane) & 1u;
                    if (mine) korder[kept - kept0 + __popc(K & ((1u << lane) - 1u))] = (unsigned short)(w * 32 + lane);
                    kept += __popc(K);
#pragma unroll
                    for (int q = 0; q < CH / 32; ++q) {
                        const uint32_t rem = __reduce_or_sync(0xffffffffu, mine ? row[q] : 0u);
                        if (lane == q && q > w) al &= ~rem;
                    }
                }
                if (lane == 0) s_kept = kept;
            }
            __syncthreads();
            for (int q = tid; q < s_kept - kept0; q += NT) {  // kept boxes of this chunk -> kept list
                const int i = korder[q];
                kbox[kept0 + q] = cbox[i];
                karea[kept0 + q] = carea[i];
                kidx[kept0 + q] = (uint32_t)scomp[c0 + i];
            }
            __syncthreads();
        }
        processed += target;
        bnd = thr;
        __syncthreads();
    }
    __syncthreads();

    // ---------------- emit ----------------
    const int kept = s_kept;
    if (tid == 0) count[b] = kept;
    for (int q = tid; q < d.max_det; q += NT) {
        float* o = det + ((size_t)b * d.max_det + q) * 6;
        if (q < kept) {
            const uint32_t id = kidx[q];
            const int a = id / nc, cl = id - a * nc;
            const Box r = load_box(pb, A, a);
            o[0] = r.x1; o[1] = r.y1; o[2] = r.x2; o[3] = r.y2;
            o[4] = ps[(size_t)cl * A + a];
            o[5] = (float)cl;
            keep[(size_t)b * d.max_det + q] = a;
        } else {
            o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.f;
            keep[(size_t)b * d.max_det + q] = -1;
        }
    }
}

inline size_t kept_smem(int max_det) { return (size_t)max_det * (sizeof(Box) + 2 * sizeof(float)); }

}  // namespace
}  // namespace fce

using namespace fce;

extern "C" size_t fce_nms_workspace(const fce_nms_desc* d) {
    if (!d) return 0;
    // keys: one uint32 per anchor (predict) or per (class, anchor) (multi-label); + class ids in predict mode
    const size_t n_keys = d->multi_label ? (size_t)d->A * d->nc : (size_t)d->A;
    return (size_t)d->B * (n_keys + (d->multi_label ? 0 : (size_t)d->A)) * sizeof(uint32_t);
}

extern "C" int fce_nms(const fce_nms_desc* d, const float* pred, const int32_t* classes, float* det, int64_t* keep,
                       int32_t* count, void* ws, size_t ws_bytes, void* stream) {
    if (!d || !pred || !det || !keep || !count || !ws) return FCE_ERR_BAD_ARG;
    if (d->B <= 0 || d->A <= 0 || d->nc <= 0 || d->max_det <= 0 || d->max_nms <= 0) return FCE_ERR_BAD_ARG;
    if (d->n_classes > 0 && !classes) return FCE_ERR_BAD_ARG;
    if (!(d->conf_thres >= 0.f && d->conf_thres <= 1.f) || !(d->iou_thres >= 0.0 && d->iou_thres <= 1.0))
        return FCE_ERR_BAD_ARG;  // nms.py:59-60 asserts
    if ((long long)d->A * d->nc >= (1ll << 31)) return FCE_ERR_UNSUPPORTED;
    if (d->nc > 65535 || d->B > 65535) return FCE_ERR_UNSUPPORTED;
    if (ws_bytes < fce_nms_workspace(d)) return FCE_ERR_WORKSPACE;
    const size_t smem = kept_smem(d->max_det);
    if (smem > 96 * 1024) return FCE_ERR_UNSUPPORTED;
    static DeviceOnce attr_once;  // per-device attribute
    int dev_ = 0;
    if (attr_once.pending(&dev_)) {
        cudaError_t e = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess) { set_cuda_error(e); return FCE_ERR_CUDA; }
        attr_once.done(dev_);
    }
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* keys = (uint32_t*)ws;
    const size_t n_keys = d->multi_label ? (size_t)d->A * d->nc : (size_t)d->A;
    uint32_t* cls = keys + (size_t)d->B * n_keys;
    const int ablocks = (d->A + 255) / 256;
    if (d->multi_label)
        nms_keys_multi<<<dim3(ablocks, d->nc, d->B), 256, 0, st>>>(*d, pred, classes, keys);
    else
        nms_keys_single<<<dim3(ablocks, d->B), 256, 0, st>>>(*d, pred, classes, keys, cls);
    int rc = check_launch();
    if (rc != FCE_OK) return rc;
    // torchvision compares the fp32 IoU with a C double threshold: reproduce with an fp32 compare
    const float tf = (float)d->iou_thres;
    const int ge = ((double)tf > d->iou_thres) ? 1 : 0;
    nms_kernel<<<d->B, NT, smem, st>>>(*d, pred, keys, cls, det, keep, count, tf, ge);
    return check_launch();
}

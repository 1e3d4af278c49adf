// Stem patch packing: the first Conv of the graph (3x3, stride 2, Cin = 3; reference conv.py:80-89 on the
// network input, yolo11-fce.yaml:20) has K = 27 - far too thin for an implicit-GEMM main loop and not
// addressable by TMA (3 channels = 3 or 12 bytes per pixel).  This kernel gathers each output pixel's 3x3x3
// patch into one 64-byte row [27 values | 5 zeros] of a bf16 matrix A[M = B*Ho*Wo, 32]; the stem then runs as a
// 1x1 convolution (K = 32) on the tcgen05 kernel with the reference's OHWI weights laid out the same way.
// uint8 images are stored unscaled (0..255 are exact in bf16) - the 1/255 is folded into the weights.
// HBM-bound: algorithmic bytes per output pixel = 27 * e_in / 4 (input, each byte used ~2.25 times) + 64 written.
#include <cuda_fp16.h>

#include "common.cuh"

namespace fce {
namespace {

constexpr int NT = 256;

template <typename TI, int LAYOUT>
__global__ void __launch_bounds__(NT) stem_pack_kernel(const fce_pack_desc d, const TI* __restrict__ x,
                                                       uint4* __restrict__ a, int Ho, int Wo, int M) {
    const int m = blockIdx.x * NT + threadIdx.x;
    if (m >= M) return;
    const int hw = Ho * Wo;
    const int b = m / hw;
    const int rem = m - b * hw;
    const int ho = rem / Wo, wo = rem - ho * Wo;
    const int h0 = ho * 2 - 1, w0 = wo * 2 - 1;
    float v[27];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int hi = h0 + kh;
        const bool h_ok = hi >= 0 && hi < d.H;
        const int hc = h_ok ? hi : 0;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int wi = w0 + kw;
            const bool ok = h_ok && wi >= 0 && wi < d.W;
            const int wc = ok ? wi : 0;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                size_t idx;
                if (LAYOUT == FCE_NHWC)
                    idx = ((size_t)(b * d.H + hc) * d.W + wc) * 3 + ci;
                else
                    idx = ((size_t)(b * 3 + ci) * d.H + hc) * d.W + wc;
                const float t = Elem<TI>::to_f(__ldg(x + idx));
                v[(kh * 3 + kw) * 3 + ci] = ok ? t : 0.f;
            }
        }
    }
    uint32_t o[16];
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        o[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[26], 0.f);
        o[13] = *reinterpret_cast<uint32_t*>(&h);
    }
    o[14] = o[15] = 0u;
    uint4* dst = a + (size_t)m * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
}


// ------------------------------------------------------------------------------------------------ fused stem
// fce_stem_conv: the whole first layer in ONE pass - image tile -> shared memory (fp16) -> per-warp
// mma.sync m16n8k16 (M = 16 output pixels, K = 27 padded to 32, N = Cout) -> bias + SiLU -> bf16 NHWC, staged
// through shared memory so that every global store is a full 16-byte chunk of a contiguous pixel run.
// Algorithmic HBM bytes per output pixel: 12 * e_in read (each input byte lands in ~2.25 patches but is fetched
// once per CTA tile) + 2 * Cout written; the [M, 32] patch matrix of the two-kernel route (64 B written + 64 B
// read per pixel) never exists.  A tcgen05 pipeline has nothing to offer at K = 32, N <= 96: the layer is
// bound by the output write and by instruction issue, so the kernel is written for few instructions per pixel:
//   * the tile keeps (w, ci) interleaved exactly like the image, so the 9 values of one patch row kh are 9
//     CONTIGUOUS fp16 starting at a 4-byte aligned address.  The GEMM K axis is therefore re-ordered to
//     (kh, j) = stem_kh / stem_j(k') below (j = 0..8 = kw*3+ci, j = 9 and k' >= 30 carry zero weights): every A-fragment register
//     (two consecutive k') is ONE aligned 32-bit shared-memory load, no byte gathering or packing;
//   * uint8 images are staged eight values per item from three aligned words (the tile origin is shifted left to a
//     4-byte boundary, one leading pad element per row keeps the patch origins even) and converted two at a time to
//     fp16 (magic-number trick in the half domain); the operands are fp16, not bf16: byte/256 is exact in fp16 and the
//     bf16 weights convert exactly;
//   * SiLU is h + h * tanh(h), h = v/2: one MUFU per value; the weights and the bias are halved when the B fragments
//     are built (exact in bf16), so the accumulator IS h, and the h + h*t runs on packed fp32 pairs (FFMA2).
constexpr int ST_ROWS = 8, ST_COLS = 64;               // output tile of one CTA
constexpr int ST_IR = 2 * ST_ROWS + 1;                 // input rows incl. the top halo
constexpr int ST_SHIFT = 3;                            // extra left columns so that the u8 tile starts 4-byte aligned
constexpr int ST_IC = 2 * ST_COLS + 2 + ST_SHIFT;      // 133 input columns (399 values) per tile row
constexpr int ST_PITCH = 400;                          // fp16 elements per tile row: element 1 + 3*c + ci holds (c, ci);
                                                       // the leading pad element makes every patch origin EVEN
constexpr int ST_THREADS = 128;
// Re-ordered K axis: k' 0..23 = the first eight values of patch rows 0 / 1 / 2, k' 24..29 = (value 8, the zero-weight
// neighbour 9) of rows 0 / 1 / 2, k' 30..31 zero.  Three of the four A-fragment loads of a warp then stay inside ONE
// patch row and the fourth reads the same word of three rows: no shared-memory bank conflicts.  (The first ordering,
// k' = 10 * row + j, mixed two rows in one load: 47 % of the shared-load wavefronts were conflict replays and the LSU
// data pipe sat at 89 % in the ncu capture.)
__host__ __device__ constexpr int stem_kh(int kp) { return kp < 24 ? kp / 8 : (kp - 24) / 2; }
__host__ __device__ constexpr int stem_j(int kp) { return kp < 24 ? kp % 8 : 8 + (kp - 24) % 2; }

template <typename TI, int LAYOUT, int NTILES>
__global__ void __launch_bounds__(ST_THREADS) stem_fused_kernel(const fce_stem_desc d, const TI* __restrict__ x,
                                                                const __nv_bfloat16* __restrict__ w,
                                                                const float* __restrict__ bias,
                                                                __nv_bfloat16* __restrict__ y, int Ho, int Wo) {
    pdl_trigger();
    constexpr int COUT = NTILES * 8;
    constexpr int OPITCH = COUT + 8;  // staged pixel pitch (bf16): +16 B keeps the quad-strided writes conflict-free
    __shared__ __align__(16) __half tile[ST_IR * ST_PITCH];  // fp16: u8 / 256 and [0, 1] floats are (near-)exact
    __shared__ __align__(16) __nv_bfloat16 stage[4][16 * OPITCH];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int b = blockIdx.z, ho0 = blockIdx.y * ST_ROWS, wo0 = blockIdx.x * ST_COLS;
    const int hi0 = 2 * ho0 - 1, wi0 = 2 * wo0 - 1 - ST_SHIFT;  // tile origin in the image

    // ---- stage the input tile (zero outside the image = the conv padding).  Tile column c is image column
    // wi0 + c; value (c, ci) lives at element 1 + 3*c + ci of its row.
    if (LAYOUT == FCE_NHWC && sizeof(TI) == 1 && (d.W & 3) == 0) {
        // u8: the tile row starts at byte 3 * wi0 = 6 * wo0 - 12 of an image row of 3 * W bytes - both multiples of
        // 4.  Elements 4q .. 4q+3 are image bytes 4q-1 .. 4q+2 of the row segment: the top byte of word q-1 and the
        // low three bytes of word q (one funnel shift).
        const uint32_t* xw = reinterpret_cast<const uint32_t*>(x);
        uint4* tw = reinterpret_cast<uint4*>(tile);
        const int row_words = d.W * 3 / 4;
        const int w_first = wi0 * 3 / 4;  // exact (may be negative)
        // A thread converts EIGHT values per item from three consecutive words (the first version loaded two words per
        // four values and spent half of the kernel's issue slots on this staging); fully unrolled so that all global
        // loads are in flight before the first conversion.
        constexpr int ITEMS = ST_IR * (ST_PITCH / 8);                      // 850 groups of eight values
        constexpr int TRIPS = (ITEMS + ST_THREADS - 1) / ST_THREADS;       // 7
        uint32_t wa[TRIPS], wb[TRIPS], wc[TRIPS];
#pragma unroll
        for (int t = 0; t < TRIPS; ++t) {
            const int i = tid + t * ST_THREADS;
            const int r = i / (ST_PITCH / 8), q = 2 * (i - r * (ST_PITCH / 8));
            const int hi = hi0 + r, wq = w_first + q;
            wa[t] = wb[t] = wc[t] = 0u;
            if (i < ITEMS && hi >= 0 && hi < d.H) {
                const uint32_t* rowp = xw + (size_t)(b * d.H + hi) * row_words;
                if (wq - 1 >= 0 && wq - 1 < row_words) wa[t] = __ldg(rowp + wq - 1);
                if (wq >= 0 && wq < row_words) wb[t] = __ldg(rowp + wq);
                if (wq + 1 >= 0 && wq + 1 < row_words) wc[t] = __ldg(rowp + wq + 1);
            }
        }
        // two bytes -> two fp16 at once: 0x4400 | byte is the half 4 + byte/256 (ulp 2^-8 in [4, 8)); subtracting 4
        // leaves byte/256 exactly.  The 256 is folded into the B fragments below.
        const __half2 four = __float2half2_rn(4.f);
        auto cvt2 = [&](uint32_t v, uint32_t sel) {
            const uint32_t m = __byte_perm(v, 0x44444444u, sel);
            const __half2 h = __hsub2(*reinterpret_cast<const __half2*>(&m), four);
            return *reinterpret_cast<const uint32_t*>(&h);
        };
#pragma unroll
        for (int t = 0; t < TRIPS; ++t) {
            const int i = tid + t * ST_THREADS;
            if (i >= ITEMS) continue;
            const uint32_t v0 = __funnelshift_l(wa[t], wb[t], 8), v1 = __funnelshift_l(wb[t], wc[t], 8);
            tw[i] = make_uint4(cvt2(v0, 0x4140), cvt2(v0, 0x4342), cvt2(v1, 0x4140), cvt2(v1, 0x4342));
        }
    } else if (LAYOUT == FCE_NHWC) {
        for (int i = tid; i < ST_IR * ST_PITCH; i += ST_THREADS) {
            const int r = i / ST_PITCH, e = i - r * ST_PITCH - 1;  // e = 3*c + ci
            const int c = e / 3;
            const int hi = hi0 + r, wi = wi0 + c;
            float v = 0.f;
            if (e >= 0 && hi >= 0 && hi < d.H && wi >= 0 && wi < d.W)
                v = Elem<TI>::to_f(__ldg(x + ((long long)(b * d.H + hi) * d.W + wi) * 3 + (e - c * 3)));
            tile[i] = __float2half_rn(sizeof(TI) == 1 ? v * (1.f / 256.f) : v);  // u8 tiles hold byte/256 on every path
        }
    } else {
        for (int i = tid; i < 3 * ST_IR * ST_IC; i += ST_THREADS) {
            const int ci = i / (ST_IR * ST_IC), rem = i - ci * (ST_IR * ST_IC);
            const int r = rem / ST_IC, c = rem - r * ST_IC;
            const int hi = hi0 + r, wi = wi0 + c;
            float v = 0.f;
            if (hi >= 0 && hi < d.H && wi >= 0 && wi < d.W)
                v = Elem<TI>::to_f(__ldg(x + ((size_t)(b * 3 + ci) * d.H + hi) * d.W + wi));
            tile[r * ST_PITCH + 1 + c * 3 + ci] = __float2half_rn(v);
        }
    }
    // ---- B fragments in the re-ordered K axis: k' -> (kh, j)  <->  k = kh*9 + j of the [COUT][32] weight rows
    uint32_t bf[2][NTILES][2];
    float bs[NTILES][2];
    const unsigned short* wus = reinterpret_cast<const unsigned short*>(w);
#pragma unroll
    for (int nt = 0; nt < NTILES; ++nt) {
        const unsigned short* wr = wus + (nt * 8 + g) * 32;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t v = 0u;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int kp = ks * 16 + hf * 8 + 2 * t + e;
                    const int kh = stem_kh(kp), j = stem_j(kp);
                    if (kp < 30 && j < 9) v |= (uint32_t)__ldg(wr + kh * 9 + j) << (16 * e);
                }
                // bf16 weights -> fp16 B fragments, scaled by 256 for u8 tiles (which hold byte/256) and by 1/2 for the
                // SiLU epilogue (the MMA then produces h = v/2 directly, see below); powers of two: exact
                {
                    const float sc = (sizeof(TI) == 1 ? 256.f : 1.f) *
                                     (d.act == FCE_ACT_SILU ? 0.5f : 1.f);
                    const __half2 h = __floats2half2_rn(__uint_as_float(v << 16) * sc, __uint_as_float(v & 0xffff0000u) * sc);
                    v = *reinterpret_cast<const uint32_t*>(&h);
                }
                bf[ks][nt][hf] = v;
            }
        const float bscale = d.act == FCE_ACT_SILU ? 0.5f : 1.f;
        bs[nt][0] = __ldg(bias + nt * 8 + 2 * t) * bscale;
        bs[nt][1] = __ldg(bias + nt * 8 + 2 * t + 1) * bscale;
    }
    // ---- this thread's 4 A-fragment word offsets (in 32-bit words, relative to the patch origin)
    int aoff[2][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int kp = ks * 16 + hf * 8 + 2 * t;
            const int kh = stem_kh(kp), j = stem_j(kp);
            aoff[ks][hf] = kp < 30 ? (kh * ST_PITCH + j) / 2 : -1;
        }
    __syncthreads();

    const uint32_t* tl = reinterpret_cast<const uint32_t*>(tile);
    __nv_bfloat16* stg = stage[warp];
    // warp w owns output rows 2w, 2w+1 of the tile: 8 runs of 16 consecutive pixels
#pragma unroll 1
    for (int run = 0; run < 8; ++run) {
        const int ro = warp * 2 + (run >> 2), co = (run & 3) * 16;
        const int ho = ho0 + ro;
        if (ho >= Ho || wo0 + co >= Wo) continue;  // warp-uniform
        // patch origin of pixel g: row 2*ro, column 2*(co+g) + ST_SHIFT -> element 1 + 3*(2*(co+g) + 3) = 6*(co+g) + 10
        // (even), in 32-bit words; pixel g+8 is 48 elements = 24 words further
        const int p0 = ro * ST_PITCH + 3 * (co + g) + 5;
        uint32_t a[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int o = aoff[ks][hf];
                a[ks][2 * hf] = o >= 0 ? tl[p0 + o] : 0u;           // row g
                a[ks][2 * hf + 1] = o >= 0 ? tl[p0 + 24 + o] : 0u;  // row g+8
            }
        __syncwarp();  // the previous run's staged pixels have been read
#pragma unroll
        for (int nt = 0; nt < NTILES; ++nt) {
            float c[4] = {bs[nt][0], bs[nt][1], bs[nt][0], bs[nt][1]};
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                         : "r"(a[0][0]), "r"(a[0][1]), "r"(a[0][2]), "r"(a[0][3]), "r"(bf[0][nt][0]), "r"(bf[0][nt][1]));
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                         : "r"(a[1][0]), "r"(a[1][1]), "r"(a[1][2]), "r"(a[1][3]), "r"(bf[1][nt][0]), "r"(bf[1][nt][1]));
            if (d.act == FCE_ACT_SILU) {  // c holds h = v/2: silu(v) = h + h*tanh(h), on packed pairs
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    const float2 h = make_float2(c[j], c[j + 1]);
                    float t0, t1;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h.x));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h.y));
                    const float2 o = __ffma2_rn(h, make_float2(t0, t1), h);
                    c[j] = o.x;
                    c[j + 1] = o.y;
                }
            }
            __nv_bfloat162 lo = __floats2bfloat162_rn(c[0], c[1]), hi = __floats2bfloat162_rn(c[2], c[3]);
            *reinterpret_cast<__nv_bfloat162*>(stg + g * OPITCH + nt * 8 + 2 * t) = lo;
            *reinterpret_cast<__nv_bfloat162*>(stg + (g + 8) * OPITCH + nt * 8 + 2 * t) = hi;
        }
        __syncwarp();
        // 16 pixels x NTILES 16-byte chunks, contiguous in global memory when out_pitch == COUT
        const size_t pix0 = ((size_t)(b * Ho + ho) * Wo + wo0 + co);
#pragma unroll
        // chunk-fastest: a lane quad writes one pixel's 64 contiguous bytes.  (Pixel-fastest reads would be free of the
        // one bank conflict per staged read - pixel 1's last chunk sits 128 bytes after pixel 0's first - but every warp
        // store then covers 16 half-written 64-byte pixel rows: measured 135 -> 161 us.)
        for (int i = lane; i < 16 * NTILES; i += 32) {
            const int px = i / NTILES, ch = i - px * NTILES;
            if (wo0 + co + px < Wo) {
                const uint4 v = *reinterpret_cast<const uint4*>(stg + px * OPITCH + ch * 8);
                *reinterpret_cast<uint4*>(y + (pix0 + px) * d.out_pitch + d.out_off + ch * 8) = v;
            }
        }
    }
}

template <typename TI, int LAYOUT>
int launch_stem(const fce_stem_desc* d, const void* x, const void* w, const float* bias, void* y, cudaStream_t st) {
    const int Ho = (d->H - 1) / 2 + 1, Wo = (d->W - 1) / 2 + 1;
    dim3 grid((Wo + ST_COLS - 1) / ST_COLS, (Ho + ST_ROWS - 1) / ST_ROWS, d->B);
    const TI* xi = (const TI*)x;
    const __nv_bfloat16* wi = (const __nv_bfloat16*)w;
    __nv_bfloat16* yo = (__nv_bfloat16*)y;
    switch (d->Cout / 8) {
        case 2: stem_fused_kernel<TI, LAYOUT, 2><<<grid, ST_THREADS, 0, st>>>(*d, xi, wi, bias, yo, Ho, Wo); break;
        case 4: stem_fused_kernel<TI, LAYOUT, 4><<<grid, ST_THREADS, 0, st>>>(*d, xi, wi, bias, yo, Ho, Wo); break;
        case 6: stem_fused_kernel<TI, LAYOUT, 6><<<grid, ST_THREADS, 0, st>>>(*d, xi, wi, bias, yo, Ho, Wo); break;
        case 8: stem_fused_kernel<TI, LAYOUT, 8><<<grid, ST_THREADS, 0, st>>>(*d, xi, wi, bias, yo, Ho, Wo); break;
        case 12: stem_fused_kernel<TI, LAYOUT, 12><<<grid, ST_THREADS, 0, st>>>(*d, xi, wi, bias, yo, Ho, Wo); break;
        default: return FCE_ERR_UNSUPPORTED;
    }
    return check_launch();
}

}  // namespace
}  // namespace fce

using namespace fce;

extern "C" int fce_stem_pack(const fce_pack_desc* d, const void* x, void* a, void* stream) {
    if (!d || !x || !a || d->B <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    if (d->Cin != 3 || d->k != 3 || d->stride != 2 || d->Kpad != 32) return FCE_ERR_UNSUPPORTED;
    if (((uintptr_t)a) & 15) return FCE_ERR_ALIGNMENT;
    const int Ho = (d->H + 2 - 3) / 2 + 1, Wo = (d->W + 2 - 3) / 2 + 1;
    const long long M = (long long)d->B * Ho * Wo;
    if (M > 0x7fffffffLL) return FCE_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)((M + NT - 1) / NT);
    if (d->in_dtype == FCE_U8 && d->in_layout == FCE_NHWC)
        stem_pack_kernel<uint8_t, FCE_NHWC><<<grid, NT, 0, st>>>(*d, (const uint8_t*)x, (uint4*)a, Ho, Wo, (int)M);
    else if (d->in_dtype == FCE_F32 && d->in_layout == FCE_NCHW)
        stem_pack_kernel<float, FCE_NCHW><<<grid, NT, 0, st>>>(*d, (const float*)x, (uint4*)a, Ho, Wo, (int)M);
    else if (d->in_dtype == FCE_F32 && d->in_layout == FCE_NHWC)
        stem_pack_kernel<float, FCE_NHWC><<<grid, NT, 0, st>>>(*d, (const float*)x, (uint4*)a, Ho, Wo, (int)M);
    else
        return FCE_ERR_UNSUPPORTED;
    return check_launch();
}

extern "C" int fce_stem_conv(const fce_stem_desc* d, const void* x, const void* w, const float* bias, void* y,
                             void* stream) {
    if (!d || !x || !w || !bias || !y || d->B <= 0 || d->H <= 0 || d->W <= 0) return FCE_ERR_BAD_ARG;
    if (d->Cout % 8 || d->Cout <= 0 || (d->act != FCE_ACT_SILU && d->act != FCE_ACT_NONE)) return FCE_ERR_UNSUPPORTED;
    if (d->B > 65535) return FCE_ERR_UNSUPPORTED;
    if (d->out_pitch % 8 || d->out_off % 8 || ((uintptr_t)y & 15) || ((uintptr_t)w & 3) || ((uintptr_t)x & 3))
        return FCE_ERR_ALIGNMENT;
    cudaStream_t st = (cudaStream_t)stream;
    if (d->in_dtype == FCE_U8 && d->in_layout == FCE_NHWC) return launch_stem<uint8_t, FCE_NHWC>(d, x, w, bias, y, st);
    if (d->in_dtype == FCE_F32 && d->in_layout == FCE_NCHW) return launch_stem<float, FCE_NCHW>(d, x, w, bias, y, st);
    if (d->in_dtype == FCE_F32 && d->in_layout == FCE_NHWC) return launch_stem<float, FCE_NHWC>(d, x, w, bias, y, st);
    return FCE_ERR_UNSUPPORTED;
}

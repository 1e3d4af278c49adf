/* fce_yolo_b200 - C ABI of the B200 (sm_100a) FCE-YOLOv11 detection forward path.
 *
 * The reference (ShioMisaka/fce-yolo, a fork of Ultralytics 8.3.242) has no native code: every
 * operator is a torch.nn.Module calling ATen.  Each entry point below replaces the ATen work of
 * one reference operator (file:line cited per function, relative to the reference root).  The
 * Python host (fce_yolo_b200/_lib.py) binds them with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *  - Plain pointers and sizes only.  All pointers are DEVICE pointers unless stated; the caller
 *    (PyTorch caching allocator) owns every buffer, the library allocates nothing.
 *  - Activations are NHWC ("channels_last"): element (b,h,w,c) of a view lives at
 *    base + ((b*H + h)*W + w)*pitch + off + c, where `pitch` is the channel count of the
 *    underlying buffer and `off` the first channel of the view.  Channel-concats and chunk()s of
 *    the reference (block.py:232,303-307,340,1464; head.py:120) are therefore free: producers
 *    write, and consumers read, channel slices of one buffer.
 *  - dtype codes: FCE_BF16 activations/weights with fp32 accumulation ("bf16 mode") or FCE_F32
 *    end to end ("fp32 mode").  Strips (pooled [B,L,C] vectors), gates, biases, logits and
 *    detections are always fp32.
 *  - `stream` is a cudaStream_t passed as void*.  Calls are asynchronous, re-entrant and
 *    stream-ordered; there is no global mutable state.
 *  - Return value: 0 on success, negative fce_status otherwise (the Python shim raises).
 *    There is no CPU fallback anywhere.
 */
#ifndef FCE_YOLO_B200_H
#define FCE_YOLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    FCE_OK = 0,
    FCE_ERR_BAD_ARG = -1,      /* null pointer, non-positive size */
    FCE_ERR_UNSUPPORTED = -2,  /* shape / dtype combination outside the path */
    FCE_ERR_ALIGNMENT = -3,    /* view not aligned for the vector width the kernel needs */
    FCE_ERR_WORKSPACE = -4,    /* workspace too small */
    FCE_ERR_CUDA = -5          /* launch / driver error (see fce_last_cuda_error) */
} fce_status;

typedef enum { FCE_BF16 = 0, FCE_F32 = 1, FCE_U8 = 2 } fce_dtype;
typedef enum { FCE_ACT_NONE = 0, FCE_ACT_SILU = 1, FCE_ACT_SIGMOID = 2 } fce_act;
typedef enum { FCE_NHWC = 0, FCE_NCHW = 1 } fce_layout;

int fce_abi_version(void);
/* last cudaError_t seen by a failing call on this thread, as text */
const char* fce_last_cuda_error(void);
/* 1 if the current device is sm_100 (B200) */
int fce_device_ok(void);

/* ---------------------------------------------------------------------------------------------
 * Dense convolution k in {1,3}, stride in {1,2}, pad k/2, groups 1, + bias, + SiLU, + residual.
 * Replaces Conv.forward_fuse (ultralytics/nn/modules/conv.py:80-89) and every bare nn.Conv2d on
 * the path (fce_block.py:95,233 identity; head.py:94,103), with the Bottleneck / PSABlock
 * residual adds (block.py:474-476, 1352-1353) and torch.cat/chunk fused as view offsets.
 * Weights are OHWI: w[co][kh][kw][ci], dense.  y = act(conv(x*in_scale) + bias) (+ res), optionally as a weighted sum
 * (out_scale / res_scale / res_up below).
 * impl: 0 = auto, 1 = force the fp32-accurate SIMT kernel, 2 = force a tcgen05 kernel (strip / single-CTA / CTA-pair chosen
 *       automatically), 3 = force the CTA-pair (cta_group::2) implicit-GEMM kernel, 4 = force the single-CTA
 *       implicit-GEMM kernel, 5 / 6 = force the 3x3 stride-1 strip kernel as single CTAs / CTA pairs.  2-6 return
 *       FCE_ERR_UNSUPPORTED when the shape has no such kernel.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t B, H, W;          /* input spatial size */
    int32_t Cin, Cout;
    int32_t in_pitch, in_off;
    int32_t out_pitch, out_off;
    int32_t res_pitch, res_off;
    int32_t k, stride;
    int32_t act;              /* fce_act */
    int32_t in_dtype, w_dtype, out_dtype;
    int32_t in_layout;        /* fce_layout; FCE_NCHW only for the network input image */
    float in_scale;           /* 1.0, or 1/255 for u8 images */
    int32_t impl;
    /* Weighted-sum epilogue (BiFPN_Concat.forward, fce_block.py:55-63, fused into its realign conv), when
     * weighted == 1:   y = out_scale * act(conv(x) + bias) + res_scale * res     (a scale of 0 is legitimate: relu(w) = 0)
     * weighted == 2:   y = act(conv(x) + res + bias): res is the partial sum W' x' of the OTHER input of a two-input BiFPN
     *                  node whose consumer is this 1x1 conv (conv(w0 a + w1 b) = (w0 W) a + (w1 W) b; the node's scales are
     *                  folded into the two weight matrices by the caller, the fused map is never stored)
     * res_up = 1: res is a [B, Ho/2, Wo/2, Cout] map read through a nearest 2x upsample (the Upsample layer in front of
     * the BiFPN node, yolo11-fce.yaml:40,44).  weighted / res_up: 1x1 convs on the tcgen05 path only
     * (FCE_ERR_UNSUPPORTED otherwise). */
    int32_t weighted;
    float out_scale, res_scale;
    int32_t res_up;
} fce_conv_desc;

int fce_conv2d(const fce_conv_desc* d, const void* x, const void* w, const float* bias, const void* res,
               void* y, void* stream);

/* Stem patch packing for the first Conv of the graph (3x3, stride 2, pad 1, Cin = 3; conv.py:80-89 applied to the
 * network input, yolo11-fce.yaml:20): gathers every output pixel's 3x3x3 patch into one row of a bf16 matrix
 * a[B*Ho*Wo][Kpad = 32] (K index (kh*3+kw)*3+ci as in the OHWI weights, columns 27..31 zero), so that the stem
 * runs as a K = 32 1x1 fce_conv2d on the tensor cores.  x: uint8 NHWC (values stored unscaled - fold 1/255 into
 * the weights), fp32 NCHW (the reference's tensor input, loaders.py:562-632) or fp32 NHWC. */
typedef struct {
    int32_t B, H, W, Cin;
    int32_t k, stride, Kpad;
    int32_t in_dtype, in_layout;
} fce_pack_desc;
int fce_stem_pack(const fce_pack_desc* d, const void* x, void* a, void* stream);

/* Fused stem: the same first Conv (3x3, stride 2, pad 1, Cin = 3, + bias + SiLU; conv.py:80-89 on the network
 * input) in one pass, bf16 NHWC output - no patch matrix in HBM.  w: bf16 [Cout][32], K index (kh*3+kw)*3+ci as in
 * the OHWI weights, columns 27..31 zero, any input scale (1/255 for u8) folded in.  Cout in {16,32,48,64,96}
 * (the n/s/-/m,l/x stems).  x: uint8 NHWC, fp32 NCHW or fp32 NHWC as for fce_stem_pack. */
typedef struct {
    int32_t B, H, W;          /* input size */
    int32_t Cout;
    int32_t out_pitch, out_off;
    int32_t act;
    int32_t in_dtype, in_layout;
} fce_stem_desc;
int fce_stem_conv(const fce_stem_desc* d, const void* x, const void* w, const float* bias, void* y, void* stream);

/* The first TWO convolutions of the graph in one pass (yolo11-fce.yaml:20-21; conv.py:80-89): Conv(3, C0, 3, 2) on the
 * uint8 NHWC image followed by Conv(C0, C1, 3, 2), both + bias + act; the stem map [B, H/2, W/2, C0] stays in shared
 * memory and never exists in HBM.  w0 / b0 as for fce_stem_conv (bf16 [C0][32] with the input scale folded in, fp32 [C0]);
 * w1 bf16 OHWI [C1][3][3][C0], b1 fp32 [C1]; y bf16 NHWC view [B, H/4, W/4, C1].  Same values as fce_stem_conv +
 * fce_conv2d up to rounding noise (same operands, same bf16 rounding of the intermediate; tests/test_gpu_stem2.py).
 * Shapes: C0 in {32, 64}, C1 % 16 == 0, C1 <= 128 for C0 = 64 / <= 192 for C0 = 32 (the nine [C1, C0] tap tiles are parked
 * in shared memory, 4 * C0 + 2 * C1 <= 512 tensor-memory columns), H and W multiples of 4; anything else returns
 * FCE_ERR_UNSUPPORTED (fce_stem2_route: 1 = taken, 0 = not, without launching) and the caller issues the two launches. */
typedef struct {
    int32_t B, H, W;          /* image size */
    int32_t C0, C1;
    int32_t out_pitch, out_off;
    int32_t act0, act1;
} fce_stem2_desc;
int fce_stem2_conv(const fce_stem2_desc* d, const void* x, const void* w0, const float* b0, const void* w1,
                   const float* b1, void* y, void* stream);
int fce_stem2_route(const fce_stem2_desc* d);

/* Predict-side preprocessing, the step right before the path (LetterBox, ultralytics/data/augment.py:1589-1631,
 * + the BGR->RGB flip of BasePredictor.preprocess, ultralytics/engine/predictor.py:163-165): a batch of differently
 * sized uint8 HWC BGR images -> uint8 NHWC RGB [B, out_h, out_w, 3], resized with cv2.resize(INTER_LINEAR)'s 8-bit
 * fixed-point arithmetic (bit-exact) and padded with pad_value (114).  items: DEVICE array of B records.
 * xtab [B][out_w][4], ytab [B][out_h][4] (DEVICE, int32): per resized column / row {tap 0 index, tap 1 index,
 * weight 0, weight 1} with 11-bit weights, entries [0, new_w) / [0, new_h) used; the host computes them with
 * OpenCV's coordinate arithmetic (fce_yolo_b200/preprocess.py). */
typedef struct {
    uint64_t src;             /* device pointer of the source image (uint8, HWC, BGR) */
    int32_t src_pitch;        /* bytes per source row */
    int32_t H, W;             /* source size */
    int32_t new_w, new_h;     /* resized (un-padded) size */
    int32_t top, left;        /* where the resized image sits in the output */
    int32_t reserved;
} fce_letterbox_item;
int fce_letterbox(const fce_letterbox_item* items, const int32_t* xtab, const int32_t* ytab, int32_t B, int32_t out_h,
                  int32_t out_w, int32_t pad_value, uint8_t* out, void* stream);

/* Which route fce_conv2d takes for a descriptor: 1 = tcgen05 tensor-core kernel, 0 = CUDA-core kernel (fp32 mode,
 * pooled strips, channel counts that are not multiples of 16 ...), < 0 = fce_status.  Pure function of the descriptor and
 * the pointers' alignment (nothing is launched): lets the caller SEE a bf16 convolution that would leave the tensor
 * pipe instead of finding out from a profile. */
int fce_conv2d_route(const fce_conv_desc* d, const void* x, const void* w, const void* res, const void* y);

/* Launch counters of fce_conv2d / fce_conv2d_detect since the last call with reset != 0:
 * out[0] tcgen05 single-CTA launches, out[1] tcgen05 CTA-pair (cta_group::2) launches, out[2] tcgen05 3x3 strip-kernel
 * launches, out[3] CUDA-core launches.  (Monotonic counters; the only process-wide state of the library besides its
 * immutable kernel / tensor-map setup.) */
int fce_conv_stats(long long* out, int reset);

/* Depthwise 3x3 stride 1 (+bias, +SiLU, + optional add of a second map).  Replaces DWConv
 * (conv.py:185-199) in Detect.cv3 (head.py:101-102) and Attention.pe with its add
 * (block.py:1282,1302).  w is fp32 [9][C] (tap-major), bias fp32 [C]. */
typedef struct {
    int32_t B, H, W, C;
    int32_t in_pitch, in_off, out_pitch, out_off, add_pitch, add_off;
    int32_t act, dtype;
} fce_dwconv_desc;
int fce_dwconv3x3(const fce_dwconv_desc* d, const void* x, const float* w, const float* bias, const void* add,
                  void* y, void* stream);

/* Depthwise 3x3 stride 1 + bias + act FOLLOWED BY a 1x1 conv + bias + act, in one pass: one block of Detect's class
 * branch, cv3[i][j] = Sequential(DWConv(c, c, 3), Conv(c, c3, 1)) (head.py:101-102; conv.py:185-199, :80-89).  The
 * depthwise result (rounded to bf16, as the two-launch route stores it) goes straight into the shared-memory A operand
 * of the 1x1 on the tensor cores and never exists in HBM; results are bit-identical to fce_dwconv3x3 + fce_conv2d.
 * bf16 NHWC views in and out; w_dw fp32 [9][C] (tap-major), b_dw fp32 [C], w_pw bf16 [Cout][C], b_pw fp32 [Cout].
 * Shapes: C % 64 == 0, Cout % 16 == 0, Cout <= 256 (the [Cout, C] weights are parked in shared memory where they fit next
 * to the pipeline - C * Cout <= 64 Ki elements - and streamed chunk by chunk otherwise); anything else returns
 * FCE_ERR_UNSUPPORTED and the caller issues the two launches.  fce_dwpw_route answers that question without launching (1 = this entry point takes the shape, 0 = not;
 * pure function of the descriptor). */
typedef struct {
    int32_t B, H, W;
    int32_t C, Cout;
    int32_t in_pitch, in_off, out_pitch, out_off;
    int32_t dw_act, pw_act;      /* fce_act of the depthwise conv / of the 1x1 */
} fce_dwpw_desc;
int fce_dwpw_conv(const fce_dwpw_desc* d, const void* x, const float* w_dw, const float* b_dw, const void* w_pw,
                  const float* b_pw, void* y, void* stream);
int fce_dwpw_route(const fce_dwpw_desc* d);

/* Two chained 1x1 convs in one pass - the tail of a C3k2 block whose last inner block is a C3k:
 *     t = act1(W1 * x1 + b1)          C3k.cv3   (block.py:338-340; Conv.forward_fuse conv.py:80-89)
 *     y = act2(W2 * [x2 ; t] + b2)    C3k2.cv2 over the concat [y0, y1, ..., t]   (block.py:303-307)
 * t (rounded to bf16, as the two-launch route stores it) is written by the first GEMM's epilogue into the shared-memory
 * A operand of the second GEMM and never exists in HBM; results are bit-identical to two fce_conv2d launches.
 * bf16 NHWC views: x1 [M, c1], x2 [M, c2], y [M, Cout] (M = B*H*W); w1 bf16 [cm][c1], w2 bf16 [Cout][c2 + cm] (the t
 * columns last, as the concat orders them), b1 fp32 [cm], b2 fp32 [Cout].
 * Shapes: c1 % 64 == 0, c2 % 64 == 0, cm in {64, 128}, Cout % 128 == 0; anything else returns FCE_ERR_UNSUPPORTED and the
 * caller issues the two launches.  fce_conv1x1_chain_route answers that without launching (1 / 0, pure function of the
 * descriptor). */
typedef struct {
    int32_t B, H, W;
    int32_t c1, cm, c2, Cout;
    int32_t x1_pitch, x1_off, x2_pitch, x2_off, out_pitch, out_off;
    int32_t act1, act2;          /* fce_act of the first / second conv */
} fce_chain_desc;
int fce_conv1x1_chain(const fce_chain_desc* d, const void* x1, const void* w1, const float* b1, const void* x2,
                      const void* w2, const float* b2, void* y, void* stream);
int fce_conv1x1_chain_route(const fce_chain_desc* d);

/* SPPF pyramid: three chained 5x5/s1/p2 max-pools (= 5x5, 9x9, 13x13 windows) of slice 0 of the
 * concat buffer written to slices 1..3 (block.py:228-232).  buf has 4*C channels. */
typedef struct {
    int32_t B, H, W, C;
    int32_t pitch, off;       /* slice j lives at channel off + j*C */
    int32_t dtype;
} fce_sppf_desc;
int fce_sppf_pool(const fce_sppf_desc* d, void* buf, void* stream);

/* Nearest 2x upsample (nn.Upsample(None, 2, "nearest"), yolo11-fce.yaml:40,44). H,W = input size. */
typedef struct {
    int32_t B, H, W, C;
    int32_t in_pitch, in_off, out_pitch, out_off, dtype;
} fce_upsample_desc;
int fce_upsample2x(const fce_upsample_desc* d, const void* x, void* y, void* stream);

/* BiFPN_Concat weighted fusion (fce_block.py:55-61): y = sum_i wn[i] * x_i, n in {2,3}, where wn is
 * the already normalised relu(w)/(sum relu(w)+1e-4) (host computes it once per weight load) and
 * input i may be a half-resolution map read through the nearest-2x pattern (up[i] = 1), which
 * fuses the preceding nn.Upsample.  H,W = output size. */
typedef struct {
    int32_t B, H, W, C, n;
    int32_t pitch[3], off[3], up[3];
    float wn[3];
    int32_t out_pitch, out_off, dtype;
} fce_bifpn_desc;
int fce_bifpn_fuse(const fce_bifpn_desc* d, const void* x0, const void* x1, const void* x2, void* y, void* stream);

/* Channel concat for the stock Concat module (conv.py:616-641) when it cannot be fused away:
 * copies a view into a channel slice of another buffer. */
typedef struct {
    int32_t B, H, W, C;
    int32_t in_pitch, in_off, out_pitch, out_off, dtype;
} fce_copy_desc;
int fce_copy_view(const fce_copy_desc* d, const void* x, void* y, void* stream);

/* Coordinate pooling (fce_block.py:81-82,101-102,140-141,159-160,212-213,239-240): one pass over x
 * producing strip[b, 0:H, c] = mean_w x and strip[b, H:H+W, c] = mean_h x, fp32 [B, H+W, C]. */
typedef struct {
    int32_t B, H, W, C;
    int32_t pitch, off, dtype;
} fce_pool_desc;
int fce_coord_pool(const fce_pool_desc* d, const void* x, float* strip, void* ws, size_t ws_bytes, void* stream);
size_t fce_coord_pool_workspace(const fce_pool_desc* d);

/* 1x1 convs on strips - cv1/cv_h/cv_w (fce_block.py:105,112-113), q/k/v/proj (:165-168,178) and the
 * eight BiCoordCrossAtt projections (:246-249,258,264-268,276) - are fce_conv2d calls on the fp32 strip
 * seen as a [1, rows, 1, C] map (k = 1, all dtypes FCE_F32, act NONE / SILU / SIGMOID). */

/* Last conv of a Detect branch with the decode fused into the epilogue (head.py:94,103 + :149-167, DFL block.py:76-79,
 * anchors / dist2bbox tal.py:352-376): the [B, H, W, Cin] feature map goes through the 1x1 conv on the tensor cores and
 * the accumulator row of each pixel is decoded in registers and stored into the prediction tensor y [B, rows, A] fp32 -
 * the logit maps never exist in HBM.
 *   mode 1 (cv3[i][2], Cout = nc):      y[b, 4 + n, a_base + p] = sigmoid(conv + bias)
 *   mode 2 (cv2[i][2], Cout = 4 * 16):  y[b, 0..3, a_base + p]  = (cx, cy, w, h) * stride from the softmax-integral
 * p = h * W + w.  bf16 NHWC input, bf16 OHWI weights, k = 1 only; anything else returns FCE_ERR_UNSUPPORTED and the
 * caller uses fce_conv2d (fp32 logits) + fce_detect_decode. */
typedef struct {
    int32_t mode, A, a_base, rows, reg_max;
    float stride;
} fce_detect_epi_desc;
int fce_conv2d_detect(const fce_conv_desc* d, const fce_detect_epi_desc* e, const void* x, const void* w,
                      const float* bias, float* y, void* stream);

/* CoordAtt gate MLP on the pooled strips, fused (fce_block.py:104-113 - cv1 with its folded BN + SiLU, then cv_h /
 * cv_w + sigmoid): out[r, :] = act2(W2 act1(W1 strip[r, :] + b1) + b2) with (W2, b2) = (w_h, b_h) for rows
 * [0, rows_h) and (w_w, b_w) for rows [rows_h, rows_h + rows_w).  All operands fp32, 16-byte aligned, C / oup /
 * pitches multiples of 4, mip <= 64, weights + 16 staged rows within 200 KB of shared memory; w1q is W1 regrouped as [C/4][mip][4] (w1q[c4][m][j] = W1[m][4 c4 + j]), the
 * second layer TRANSPOSED: wht / wwt [mip][oup].  Replaces three fce_conv2d calls on the strips. */
typedef struct {
    int32_t rows_h, rows_w, C, mip, oup;
    int32_t s_pitch, out_pitch; /* row pitches of strip / out, elements */
    int32_t act1, act2;
} fce_coordatt_mlp_desc;
int fce_coordatt_mlp(const fce_coordatt_mlp_desc* d, const float* strip, const float* w1q, const float* b1,
                     const float* wht, const float* bh, const float* wwt, const float* bw, float* out, void* stream);

/* Strip cross-attention core (fce_block.py:171-175, 252-256, 271-275):
 * out[b,l,h*dh+j] = sum_m softmax_m(scale * <q[b,l,h,:], k[b,m,h,:]>) * v[b,m,h*dh+j].
 * q rows: Lq, k/v rows: Lk; channel c = head*dh + j (the reference's view(n,heads,dh,L)). */
typedef struct {
    int32_t B, heads, dh, Lq, Lk;
    float scale;
    int64_t q_bstride, q_rstride, k_bstride, k_rstride, v_bstride, v_rstride, o_bstride, o_rstride;
} fce_strip_attn_desc;
int fce_strip_attn(const fce_strip_attn_desc* d, const float* q, const float* k, const float* v, float* out,
                   void* stream);

/* Gate application (fce_block.py:116,180,283-284), one read + one write of x:
 *  mode 0 (CoordAtt)        y = x * gh[b,h,c] * gw[b,w,c]        (gh, gw already sigmoid-ed)
 *  mode 1 (CoordCrossAtt)   y = x * gh[b,h,c]
 *  mode 2 (BiCoordCrossAtt) y = x * sigmoid(gh[b,h,c] + gw[b,w,c])
 * gh/gw fp32 with row stride g_rstride and batch stride g_bstride (elements). */
typedef struct {
    int32_t B, H, W, C, mode;
    int32_t in_pitch, in_off, out_pitch, out_off, dtype;
    int64_t gh_bstride, gh_rstride, gw_bstride, gw_rstride;
} fce_gate_desc;
int fce_gate_apply(const fce_gate_desc* d, const void* x, const float* gh, const float* gw, void* y, void* stream);

/* C2PSA self-attention core (block.py:1293-1302): per image and head,
 * out[:, n] = sum_m softmax_m(scale * <q[:,n], k[:,m]>) v[:, m].  qkv is one NHWC buffer whose
 * channels were re-ordered at weight-pack time to [q(all heads) | k(all heads) | v(all heads)]. */
typedef struct {
    int32_t B, N, heads, kd, hd;
    int32_t qkv_pitch, q_off, k_off, v_off, out_pitch, out_off, dtype;
    float scale;
} fce_psa_desc;
int fce_psa_attention(const fce_psa_desc* d, const void* qkv, void* out, void* stream);

/* Detect decode (head.py:149-167 + block.py:76-79 DFL + tal.py:352-376): raw per-level logits
 * fp32 NHWC [B,Hi,Wi,4*reg_max+nc] -> y fp32 [B, 4+nc, A] = (cx,cy,w,h in input pixels, sigmoid cls). */
typedef struct {
    int32_t B, nl, nc, reg_max;
    int32_t H[4], W[4];
    float stride[4];
    int32_t raw_pitch[4];
} fce_decode_desc;
int fce_detect_decode(const fce_decode_desc* d, const float* raw0, const float* raw1, const float* raw2,
                      const float* raw3, float* y, void* stream);

/* Batched NMS (ultralytics/utils/nms.py:13-166 with torchvision.ops.nms semantics, :151-154).
 * pred fp32 [B, 4+nc, A].  Outputs: det fp32 [B, max_det, 6] (x1,y1,x2,y2,conf,cls), keep int64
 * [B, max_det] anchor indices (-1 padded), count int32 [B].  classes: optional device int32 list of
 * allowed class ids.  Keep indices and class ids are bit-exact w.r.t. the reference. */
typedef struct {
    int32_t B, A, nc;
    float conf_thres;
    double iou_thres;
    int32_t max_det, max_nms, multi_label, agnostic;
    float max_wh;
    int32_t n_classes;
} fce_nms_desc;
size_t fce_nms_workspace(const fce_nms_desc* d);
int fce_nms(const fce_nms_desc* d, const float* pred, const int32_t* classes, float* det, int64_t* keep,
            int32_t* count, void* ws, size_t ws_bytes, void* stream);

/* Back to original-image coordinates, the step right after NMS in predict (ops.scale_boxes + clip_boxes,
 * ultralytics/utils/ops.py:102-134,152-177, applied by models/yolo/detect/predict.py:109-122): rows [0, count[b]) of
 * det [B, max_det, 6] are rewritten in place as clamp((xyxy - pad) / gain, 0, (w0, h0)); meta[b] = {gain, pad_x, pad_y,
 * w0, h0} fp32 (host: fce_yolo_b200.predict.scale_meta).  Same fp32 operations as the reference's CPU path. */
int fce_scale_boxes(float* det, const int32_t* count, const float* meta, int32_t B, int32_t max_det, void* stream);

/* Validation matching, the step right after NMS in val (box_iou, ultralytics/utils/metrics.py:57-77, + the non-scipy
 * BaseValidator.match_predictions, ultralytics/engine/validator.py:266-306, as called per image by
 * DetectionValidator._process_batch, models/yolo/detect/val.py:274-288).  det [B, max_det, 6] / count [B] are the NMS
 * outputs (boxes in the same coordinate space as the labels); labels of image b are rows gt_offsets[b] ..
 * gt_offsets[b+1] of gt_boxes [G, 4] (xyxy) / gt_cls [G]; iouv [n_iou] the IoU thresholds (0.5 .. 0.95).
 * tp [B, max_det, n_iou] uint8: 1 where detection d is a true positive at threshold t, 0 elsewhere (rows >= count
 * too).  Bit-exact w.r.t. the reference except for exact fp32 IoU ties between two labels of one detection, which the
 * reference orders with an unstable sort (here: the higher label index). */
int fce_match_predictions(const float* det, const int32_t* count, const float* gt_boxes, const float* gt_cls,
                          const int32_t* gt_offsets, const float* iouv, int32_t B, int32_t max_det, int32_t n_iou,
                          int32_t max_gt_per_image, uint8_t* tp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FCE_YOLO_B200_H */

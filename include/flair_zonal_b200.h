/*
 * flair_zonal_b200.h -- C ABI of libfz_b200.so, the sm_100a implementation of
 * kezakool/flair-for-aigle's zonal segmentation hot path.
 *
 * The reference has no FFI of its own (it is pure Python calling PyTorch / smp / timm / numpy);
 * each entry point below names the reference call site it replaces.  INTEGRATION.md shows
 * the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a raw DEVICE pointer unless the name ends in _host; the caller owns
 *     all memory; the library never allocates device memory.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it and
 *     capturable in a CUDA graph.
 *   - return 0 on success, <0 on error; fz_last_error() returns a thread-local message.
 *   - activations are NHWC.  16-bit tensors of the INFERENCE entry points are in the library's operand format FZ_OP16
 *     (IEEE fp16 in this build, see fz_operand_format() below) even where a parameter is still spelled "_bf16"; the
 *     TRAINING entry points (backward / optimizer section) take __nv_bfloat16.
 */
#ifndef FLAIR_ZONAL_B200_H
#define FLAIR_ZONAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FZ_ABI_VERSION 3

const char* fz_last_error(void);
int fz_abi_version(void);
int fz_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);

/* ---------------------------------------------------------------- tile feeder
 * Replaces flair_zonal_detection/dataset.py:89-124,174-209 (_load_patch boundless windowed
 * read with fill 0, then norm.py:37-44 (x-mean)/std) and inference.py:280-286 (H2D).
 * raster: uint8 [C][H][W] band-sequential (what rasterio's read() returns), resident in HBM.
 * origins: int32 [n_tiles][2] = (row0, col0) of each tile's full window (may be negative /
 * past the edge: zero fill BEFORE normalisation, like the reference).
 * out: float32 [n_tiles][C][P][P] normalised, NCHW -- exactly the tensor the reference's
 * model receives as batch[<MOD>].
 */
int fz_gather_tiles_f32(const uint8_t* raster, int C, int H, int W, const int32_t* origins, int n_tiles, int P,
                        const float* mean, const float* std, float* out, void* stream);
/* The same window gather for a float32 raster (a DEM / elevation modality next to the uint8 imagery, dataset.py:89-124):
 * raster float32 [C][H][W], zero fill outside, (x - mean) / std in float64 rounded once to float32. */
int fz_gather_tiles_f32_from_f32(const float* raster, int C, int H, int W, const int32_t* origins, int n_tiles, int P,
                                 const float* mean, const float* stdv, float* out, void* stream);
/* The RESAMPLED window read of dataset.py:97-115 for a modality whose pixel size differs from the reference modality's:
 * windows double [n_tiles][4] = (row_off, col_off, height, width) of every tile in THIS raster's pixels (fractional),
 * resampled to ps x ps like rasterio's read(out_shape=..., resampling=bilinear, boundless=True, fill_value=0) -- GDAL's
 * RasterIO convolution restated in oracle/resample.py -- then (x - mean) / std.  raster: uint8 (src_f32 = 0, resampled
 * values rounded half up like GDAL) or float32 (src_f32 = 1) [C][H][W]; max_ratio = max over tiles of window size / ps
 * (host-checked against the kernel's tap budget).  out float32 [n_tiles][C][ps][ps]. */
int fz_gather_tiles_resampled(const void* raster, int src_f32, int C, int H, int W, const double* windows, int n_tiles, int ps,
                              double max_ratio, const float* mean, const float* stdv, float* out, void* stream);
/* Same window rule, raw bytes, NHWC uint8 [n_tiles][P][P][4] (C<=4, missing channels = 0):
 * the operand the fused stem kernel consumes (20 B/px feeder of SURVEY 8d becomes 4+4 B/px). */
int fz_gather_tiles_u8(const uint8_t* raster, int C, int H, int W, const int32_t* origins, int n_tiles, int P,
                       uint8_t* out, void* stream);

/* ---------------------------------------------------------------- crop / softmax / blend / argmax
 * Replaces inference.py:295-352 (D2H of fp32 logits, per-tile numpy crop + convert + windowed
 * write) and postprocess.py:9-30 (convert).
 * logits: [n_tiles][n_cls][P][P] (layout FZ_NCHW) or [n_tiles][P][P][cstride] (FZ_NHWC, class
 * fastest, first n_cls of cstride valid); dtype FZ_F32, FZ_BF16 or FZ_F16.
 * plan: int32 [n_tiles][6] = (row0, col0, top_px, left_px, height_px, width_px): the write
 * window of the margin-cropped prediction (inference.py:318-343); height_px<=0 = skipped.
 * own:  int32 [n_tiles][4] = (r0, r1, c0, c1) sub-window (absolute raster px) of the write
 * window that survives later tiles' overwrites ("last writer wins", inference.py:343-352);
 * NULL = write the whole window (only valid when tiles in one call do not overlap).
 */
#define FZ_F32 0
#define FZ_BF16 1
#define FZ_F16 2
/* FZ_OP16: the 16-bit OPERAND format of the inference kernels -- what they store between layers and feed to the tensor
 * cores.  fp16 by default (11 significand bits; class-map agreement with the fp32 reference 99.0 % -> see DESIGN.md
 * section 2), bf16 when the library is built with -DFZ_OPERANDS_BF16 (round-1 behaviour, A/B measurements).  Wherever an
 * inference entry point below says "bf16" / "_bf16" for an activation or weight tensor, read "FZ_OP16"; fz_operand_format()
 * returns which one this build uses.  The training entry points (backward / optimizer section) are always bf16, and
 * fz_gemm_bf16 takes the format per call (FZ_EPI_OPERANDS_F16). */
int fz_operand_format(void); /* FZ_F16 or FZ_BF16 */
#define FZ_NCHW 0
#define FZ_NHWC 1
#define FZ_NHWC_UP4 2 /* float logits at quarter resolution [n][P/4][P/4][cstride]; the x4 bilinear (align_corners=True)
                       * of smp's UPerNet SegmentationHead is evaluated inside the crop kernels */
int fz_crop_argmax_write(const void* logits, int dtype, int layout, int cstride, int n_tiles, int n_cls, int P,
                         int margin, const int32_t* plan, const int32_t* own, uint8_t* out_raster, int H, int W,
                         void* stream);
/* output_type == "class_prob": out_raster is uint8 [n_cls][H][W] = round(softmax*255). */
int fz_crop_softmax_write(const void* logits, int dtype, int layout, int cstride, int n_tiles, int n_cls, int P,
                          int margin, const int32_t* plan, const int32_t* own, uint8_t* out_raster, int H, int W,
                          void* stream);
/* output_px_meters != reference resolution (inference.py:212-226,299-312): the cropped prediction of a tile is zoomed
 * (nearest neighbour, scipy.ndimage.zoom order 0) before it is written.  zmap int32 [zoomed]: crop pixel read by output
 * offset o, or -1 where scipy writes its constant fill 0 (host: slicing.zoom_map); plan / own are in OUTPUT pixels with
 * window sizes <= zoomed.
 * mode 0 = argmax -> uint8 [H][W], 1 = class_prob -> uint8 [n_cls][H][W]. */
int fz_crop_zoom_write(int mode, const void* logits, int dtype, int layout, int cstride, int n_tiles, int n_cls, int P,
                       int margin, const int32_t* plan, const int32_t* own, const int32_t* zmap, int zoomed,
                       uint8_t* out_raster, int H, int W, void* stream);
/* Intended semantics of inference.py:468-564: canvas[n_cls][H][W] (float32, planar like the
 * reference's raster_logits) += weight(y,x) * softmax(cropped logits).
 * weight: float32 [P-2m][P-2m] or NULL (=1). */
int fz_crop_softmax_accumulate(const void* logits, int dtype, int layout, int cstride, int n_tiles, int n_cls, int P,
                               int margin, const int32_t* plan, const int32_t* plan_host, const float* weight, float* canvas,
                               int H, int W, void* stream);
/* plan_host: the same plan in HOST memory, or NULL.  Overlapping write windows have to be accumulated one after the other (no
 * atomics: the canvas is bit-reproducible).  Without the host copy every tile is its own launch; with it (n_tiles <= 96) the
 * tiles are levelled on the host -- a tile's level is one more than the highest level of an earlier tile it overlaps -- and
 * each level is ONE launch, so every pixel still receives its contributions in tile order: same bits, a few launches. */
/* The same accumulation on the rescaled output grid (inference.py:515-523 then :525-562): the cropped logits are zoomed
 * with zmap (see fz_crop_zoom_write; constant-fill pixels carry zero logits, i.e. a uniform softmax) before the softmax;
 * plan is in OUTPUT pixels. */
int fz_crop_zoom_accumulate(const void* logits, int dtype, int layout, int cstride, int n_tiles, int n_cls, int P,
                            int margin, const int32_t* plan, const int32_t* plan_host, const int32_t* zmap, int zoomed,
                            float* canvas, int H, int W, void* stream);
/* inference.py:566-572: labels = argmax_c canvas (uint8), confidence = max_c canvas (float32,
 * may be NULL).  canvas is [n_cls][n_px]. */
int fz_canvas_argmax(const float* canvas, int n_cls, int64_t n_px, uint8_t* labels, float* confidence, void* stream);
/* postprocess.py:9-30 on one (C,h,w) fp32 NCHW array: mode 0 argmax -> uint8 [h][w];
 * mode 1 class_prob -> uint8 [C][h][w]. */
int fz_convert(const float* img, int C, int h, int w, int mode, uint8_t* out, void* stream);

/* ---------------------------------------------------------------- bf16 GEMM (tcgen05 / TMEM / TMA)
 * C[m,n] = sum_k A[m,k] * B[b][n,k] (+ epilogue).  A: bf16 [M][K]; B: bf16 [b_batch][N][K];
 * b = (m / rows_per_sample) when b_batch > 1 (per-sample GRN-scaled fc2 weights) else 0.
 * Replaces the nn.Linear / 1x1 / 2x2-stride-2 conv calls under flair_model.py:376,539-541.
 */
#define FZ_EPI_BF16 0       /* out bf16 = acc + bias                                              */
#define FZ_EPI_GELU_SUMSQ 1 /* out bf16 = gelu(acc+bias); sumsq[m/128][n] = sum over the 128-row tile of out^2 (of the
                             * bf16-rounded values, i.e. exactly what the next GEMM reads) */
#define FZ_EPI_RESID_F32 2  /* out f32  = acc + bias + resid                                      */
#define FZ_EPI_F32 3        /* out f32  = acc + bias                                              */
#define FZ_EPI_RELU_BF16 4  /* out bf16 = relu(acc + bias)                                        */
#define FZ_EPI_REVERSE_TILES 0x100 /* OR into mode: walk the tile list backwards, so that an A operand the previous
                                    * kernel has just streamed out (larger than L2) is consumed newest-first */
#define FZ_EPI_GELU_BF16 5  /* out bf16 = gelu(acc + bias)   (timm Mlp.fc1 + nn.GELU of a Swin block)       */
#define FZ_EPI_OPERANDS_F16 0x200 /* OR into mode: A, B and a 16-bit output are IEEE fp16 instead of bf16 (same MMA rate,
                                   * 3 more significand bits; outputs saturate at +-65504).  The inference engines use
                                   * fp16 operands; the training step fp16 for its forward GEMMs and bf16 for its
                                   * gradient GEMMs (see the format note at the training entry points). */
int fz_gemm_bf16(const void* A, const void* B, void* out, const float* bias, const float* resid, float* sumsq, int M,
                 int N, int K, int b_batch, int rows_per_sample, int mode, void* stream);
/* Split-K variant for a small output with a long reduction (the training step's weight gradients dW = dY^T X): out float
 * [M][N] = A [M][K] * B [N][K]^T, no bias, K cut into `splits` pieces that run as independent work items of the same
 * persistent kernel (fp32 partial tiles in workspace [splits][M][N]) and are added in index order by a second kernel --
 * deterministic.  flags: FZ_EPI_OPERANDS_F16 or 0 (bf16).  fz_gemm_splitk_max_splits: the split count that gives about two
 * work items per SM while every piece keeps >= 4 k-blocks (1 = not worth splitting). */
int fz_gemm_bf16_splitk(const void* A, const void* B, float* out, float* workspace, int M, int N, int K, int splits, int flags,
                        void* stream);
int fz_gemm_splitk_max_splits(int M, int N, int K);
/* The same with both operands TRANSPOSED in memory: out [M][N] = At^T Bt with At [K][M] and Bt [K][N] row-major (MN-major
 * tcgen05 operands; M % 8 == 0).  dW = dY^T X reads dY [rows][N_l] and X [rows][K_l] as they are -- no transposed copies. */
int fz_gemm_bf16_splitk_tn(const void* At, const void* Bt, float* out, float* workspace, int M, int N, int K, int splits,
                           int flags, void* stream);
/* Diagnostics: when set to a device buffer of 64*8 uint64, CTA 0 of every following fz_gemm_bf16 launch
 * records clock64 stamps per tile (producer start, MMA arrive, accumulator free, first operands landed,
 * MMAs issued, epilogue wait, accumulator complete, epilogue done).  NULL switches it off. */
int fz_gemm_set_trace(void* device_buffer_64x8_u64);
/* Same contract on CUDA cores (exact erff); test/bring-up cross-check only. */
int fz_gemm_bf16_simt(const void* A, const void* B, void* out, const float* bias, const float* resid, float* sumsq,
                      int M, int N, int K, int b_batch, int rows_per_sample, int mode, void* stream);

/* ---------------------------------------------------------------- ConvNeXt-V2 encoder pieces
 * (timm ConvNeXtV2 as wrapped by smp TimmUniversalEncoder; call site flair_model.py:376)
 *
 * fz_stem_ln: conv4x4 stride 4 (+bias) + LayerNorm2d on raw uint8 tiles [B][P][P][4]; the
 *   (x-mean)/std normalisation of norm.py:37-44 is folded into w/bias by the host:
 *   w float [64][C0] with row k = ky*16+kx*4+c = weight[n][c][ky][kx]/std[c];
 *   bias[n] = conv_bias[n] - sum_k w[k][n]*mean[c(k)].  out float [B][P/4][P/4][C0].
 * fz_dwconv7_ln: depthwise 7x7 (pad 3, bias) + LayerNorm over C; x float NHWC, wdw float
 *   [49][C] (tap-major), out bf16 NHWC (the fc1 GEMM's A operand).
 * fz_ln2d_s2d: LayerNorm2d + space-to-depth: out bf16 [B][H/2][W/2][4C] with
 *   k = (y&1)*2C + (x&1)*C + c, the A operand of the 2x2/s2 downsample conv as a GEMM.
 * fz_grn_scale: scale[b][k] = 1 + gamma[k]*Gx/(mean_k Gx + eps), Gx = sqrt(sum_t partial[b*tps+t][k])
 *   (timm GlobalResponseNorm) from the fc1 epilogue's per-128-row-tile partial sums; fixed
 *   summation order, so the result is independent of batching.
 * fz_scale_weights: out[b][n][k] = bf16(w[n][k]*scale[b][k]) (GRN folded into fc2's weights);
 * fz_scale_rows:    h[m][k] *= scale[m/rows_per_sample][k] (GRN applied to the activations).
 */
int fz_stem_ln(const uint8_t* tiles_u8, const float* w, const float* bias, const float* ln_w, const float* ln_b,
               float* out, int B, int P, int C0, float eps, void* stream);
/* same on already-normalised float32 NCHW input [B][Cin<=4][P][P] (the tensor the reference's
 * model receives); w rows for c >= Cin must be zero. */
int fz_stem_ln_f32(const float* x_nchw, int Cin, const float* w, const float* bias, const float* ln_w,
                   const float* ln_b, float* out, int B, int P, int C0, float eps, void* stream);
int fz_dwconv7_ln(const float* x, const float* wdw, const float* bdw, const float* ln_w, const float* ln_b,
                  void* out_bf16, int B, int H, int W, int C, float eps, void* stream);
int fz_ln2d_s2d(const float* x, const float* ln_w, const float* ln_b, void* out_bf16, int B, int H, int W, int C,
                float eps, void* stream);
/* same, and the un-normalised input as bf16 rows [B*H*W][C] into copy_bf16 (may be NULL): the U-Net decoder's skip
 * operand comes for free while the stage output is in registers.  C = 128, 256, 512 when copy_bf16 is given. */
int fz_ln2d_s2d_copy(const float* x, const float* ln_w, const float* ln_b, void* out_bf16, void* copy_bf16, int B, int H,
                     int W, int C, float eps, void* stream);
int fz_grn_scale(const float* sumsq_partial, int tiles_per_sample, const float* gamma, float* scale, float* scratch,
                 int B, int K, float eps, void* stream); /* scratch: B*K/64 floats */
int fz_scale_weights(const void* w_bf16, const float* scale, void* out_bf16, int B, int N, int K, void* stream);
int fz_scale_rows(void* h_bf16, const float* scale, int64_t M, int K, int rows_per_sample, void* stream);

/* ---------------------------------------------------------------- U-Net decoder pieces
 * (smp 0.4.0 UnetDecoder + SegmentationHead; call site flair_model.py:417-419)
 *
 * fz_upsample2_concat: out[b][y][x][0:C1] = a[b][y/2][x/2][:] (F.interpolate nearest x2),
 *   out[..][C1:C1+C2] = s[b][y][x][:] (skip; C2 may be 0).  a/s dtype FZ_F32 or FZ_BF16,
 *   out bf16 [B][H][W][C1+C2].
 * fz_conv3x3_bf16: 3x3 pad 1 conv as implicit GEMM on tcgen05.  in bf16 [B][H][W][Cin];
 *   w bf16 [w_rows][3][3][Cin] (rows zero-padded to the N tile); scale/bias float [w_rows]:
 *   result = acc*scale + bias (eval BatchNorm folded into scale/bias; scale NULL = 1).  mode FZ_CONV_RELU_BF16: out bf16 [B][H][W][Cout] = relu(conv+bias);
 *   FZ_CONV_LOGITS_F32: out float [B][H][W][cstride] (first Cout valid) = conv+bias;
 *   FZ_CONV_ARGMAX_RASTER: no logits are stored -- argmax over the Cout classes is written
 *   straight into raster[RH][RW] at the tile's margin-cropped window (plan/own as in
 *   fz_crop_argmax_write; tile index = batch index).
 */
#define FZ_CONV_RELU_BF16 0
#define FZ_CONV_LOGITS_F32 1
#define FZ_CONV_ARGMAX_RASTER 2
#define FZ_CONV_LOGITS_F32_NCHW 3 /* out float [B][Cout][H][W]: the reference's logits layout */
#define FZ_CONV_ADD_RELU_BF16 4   /* out bf16 = relu(conv*scale + bias + resid): torchvision BasicBlock tail */
#define FZ_CONV_BF16 5            /* out bf16 = conv*scale + bias (no activation): BasicBlock downsample branch */
int fz_upsample2_concat(const void* a, int a_dtype, const void* s, int s_dtype, void* out16, int out_dtype, int B, int H, int W,
                        int C1, int C2, void* stream);
int fz_conv3x3_bf16(const void* in, const void* w, const float* scale, const float* bias, void* out, int B, int H,
                    int W, int Cin, int Cout, int w_rows, int mode, int cstride, const int32_t* plan,
                    const int32_t* own, uint8_t* raster, int RH, int RW, int margin, void* stream);

/* fz_upconv3x3_bn_relu: relu(bn(conv3x3(F.interpolate(a, scale_factor=2, mode='nearest')))) -- the first convolution of
 * the smp U-Net decoder blocks without a skip connection -- computed from `a` directly (sub-pixel decomposition; the
 * upsampled tensor is never built).  in bf16 [B][H][W][Cin] (SOURCE size, W % 128 == 0, H % 16 == 0, Cin 32|64);
 * w16 bf16 [w_rows][16][Cin]: merged taps, tile ((py*2+px)*2+ra)*2+ca = sum of the 3x3 taps (ky, kx) with
 * ky in G(py,ra), kx in G(px,ca), G(0,0)={0}, G(0,1)={1,2}, G(1,0)={0,1}, G(1,1)={2}; out bf16 [B][2H][2W][Cout],
 * Cout 16|32; scale/bias as in fz_conv3x3_bf16. */
int fz_upconv3x3_bn_relu(const void* in, const void* w16, const float* scale, const float* bias, void* out, int B, int H,
                         int W, int Cin, int Cout, int w_rows, void* stream);

/* fz_catconv3x3_bn_relu: relu(bn(conv3x3(cat(nearest_up2(a), skip)))) -- the first convolution of the smp U-Net decoder
 * blocks WITH a skip connection -- as one implicit GEMM; the concatenated tensor is never built.  a bf16
 * [B][Hs][Ws][C1]; skip bf16 [B][2Hs][2Ws][C2]; w16a bf16 [w_rows][16][C1] = merged sub-pixel taps of the first C1 input
 * channels (layout as in fz_upconv3x3_bn_relu); w bf16 [w_rows][3][3][C1+C2] = the convolution's own weights (their
 * skip-channel part is used); out bf16 [B][2Hs][2Ws][Cout].  C1, C2, Cout multiples of 64. */
int fz_catconv3x3_bn_relu(const void* a, const void* skip, const void* w16a, const void* w, const float* scale,
                          const float* bias, void* out, int B, int Hs, int Ws, int C1, int C2, int Cout, int w_rows,
                          void* stream);

/* ---------------------------------------------------------------- ResNet-34 encoder front end
 * (smp native ResNetEncoder = torchvision ResNet without fc; `resnet34-unet`, BASELINE.json configs[0])
 * fz_conv7x7s2_bn_relu: conv 7x7 stride 2 pad 3 + eval BatchNorm (scale/bias) + ReLU.  in: uint8 [B][P][P][4]
 *   (in_is_f32 = 0; normalisation folded into w/bias by the host) or float [B][Cin][P][P] (in_is_f32 = 1);
 *   w float [196][64] with row k = (ky*7+kx)*4 + c; out bf16 [B][P/2][P/2][64].
 * fz_maxpool3x3s2: max-pool 3x3 stride 2 pad 1 on bf16 NHWC. */
int fz_conv7x7s2_bn_relu(const void* in, int in_is_f32, int Cin, const float* w, const float* scale, const float* bias,
                         void* out_bf16, int B, int P, void* stream);
int fz_maxpool3x3s2(const void* in_bf16, void* out_bf16, int B, int H, int W, int C, void* stream);

/* fz_conv3x3_ex: fz_conv3x3_bf16 plus stride (1 | 2; H, W are the OUTPUT size, the input is H*stride x W*stride,
 * padding 1) and a residual operand (bf16 [B][H][W][Cout]) for FZ_CONV_ADD_RELU_BF16.  Covers the smp/torchvision
 * ResNet BasicBlock (conv3x3 s1|s2 + BN + ReLU, conv3x3 + BN, + identity, ReLU) called at flair_model.py:376 for
 * `resnet34-unet`; the 1x1/s2 downsample conv is expressed as a 3x3 with only the centre tap non-zero. */
int fz_conv3x3_ex(const void* in, const void* w, const float* scale, const float* bias, void* out, const void* resid,
                  int B, int H, int W, int Cin, int Cout, int w_rows, int stride, int mode, int cstride,
                  const int32_t* plan, const int32_t* own, uint8_t* raster, int RH, int RW, int margin, void* stream);

/* ---------------------------------------------------------------- Swin encoder pieces
 * (timm SwinTransformer behind smp TimmUniversalEncoder, `swin_base_patch4_window12_384-upernet`, BASELINE.json
 * configs[2]; call site flair_model.py:376).  The linears run through fz_gemm_bf16; patch embedding through
 * fz_stem_ln / fz_stem_ln_f32 with eps = 1e-5.
 * fz_layernorm_rows: nn.LayerNorm(C) of float rows [rows][C] -> bf16 (C in 128, 256, 512, 1024, 2048).
 * fz_merge_ln: timm PatchMerging up to its Linear: x float [B][H][W][C] -> bf16 [B][H/2][W/2][4C], channel blocks
 *   (h0w0, h1w0, h0w1, h1w1), LayerNorm over 4C.
 * fz_swin_window_attn: timm SwinTransformerBlock._attn between the qkv and proj linears: cyclic shift (roll by
 *   -shift), zero-pad bottom/right to a multiple of `window`, window partition, softmax(q k^T * scale +
 *   relative-position bias + shift mask (-100, regions on the padded grid)) v, window reverse, crop, roll back.
 *   qkv bf16 [B][H][W][3C] (q | k | v, channel = head*32 + d) in natural token order; qkv_bias bf16 [3C] = q/k/v of
 *   a padded (zero) token; table float [heads][(2*window-1)^2]; out bf16 [B][H][W][C].  window <= 12, head dim 32.
 * fz_cast_f32_bf16: stage outputs (float) -> decoder operands (bf16). */
int fz_layernorm_rows(const float* x, const float* w, const float* b, void* out_bf16, int64_t rows, int C, float eps,
                      void* stream);
int fz_merge_ln(const float* x, const float* w, const float* b, void* out_bf16, int B, int H, int W, int C, float eps,
                void* stream);
int fz_swin_window_attn(const void* qkv_bf16, const void* qkv_bias_bf16, const float* table, void* out_bf16, int B,
                        int H, int W, int C, int heads, int window, int shift, float scale, void* stream);
int fz_cast_f32_bf16(const float* in, void* out_bf16, int64_t n, void* stream);
/* fp32 -> FZ_BF16 or FZ_F16 (round to nearest even; fp16 saturates at +-65504) */
int fz_cast_f32_16(const float* in, void* out16, int out_dtype, int64_t n, void* stream);

/* smp UPerNetDecoder fuse input in one pass (the five strided slice writes it replaces are the calls above):
 * out bf16 [B][H][H][5C] = [bilinear(p0: s0^2 -> H^2) | bilinear(p1) | bilinear(p2) | p3 | down2(up2(p3))], all maps bf16
 * NHWC with C channels, C / 8 a divisor of 256. */
int fz_pyramid_concat(const void* p0, int s0, const void* p1, int s1, const void* p2, int s2, const void* p3,
                      void* out_bf16, int B, int H, int C, void* stream);
/* ---------------------------------------------------------------- UPerNet decoder pieces
 * (smp 0.4.0 UPerNetDecoder + SegmentationHead(kernel_size=1, upsampling=4); call site flair_model.py:417-419)
 * fz_adaptive_avgpool: nn.AdaptiveAvgPool2d(S) on bf16 NHWC -> bf16 [B][S][S][C].
 * fz_bilinear_slice: out[b][y][x][c0:c0+C] = bilinear(in, size=(H,W), align_corners=False)[b][y][x][:] (+ add[b][y][x][:]
 *   when add != NULL); in bf16 [B][h][w][C], add bf16 [B][H][W][C], out bf16 [B][H][W][Ctot].  h == H is a copy.
 * fz_updown_slice: same destination convention for down2(up2(in)) at in's own size (the 0-channel FPN stage followed
 *   by the resize back to H/4): separable [1/8, 3/4, 1/8] with replicated borders.
 * fz_head_upsample4: nn.UpsamplingBilinear2d(scale_factor=4) (align_corners=True): logits float [B][h][w][cstride]
 *   (first n_cls valid) -> float [B][n_cls][4h][4w]. */
int fz_adaptive_avgpool(const void* in_bf16, void* out_bf16, int B, int H, int W, int C, int S, void* stream);
int fz_bilinear_slice(const void* in_bf16, const void* add_bf16, void* out_bf16, int B, int h, int w, int H, int W,
                      int C, int Ctot, int c0, void* stream);
int fz_updown_slice(const void* in_bf16, void* out_bf16, int B, int H, int W, int C, int Ctot, int c0, void* stream);
int fz_head_upsample4(const float* logits, float* out_nchw, int B, int h, int w, int cstride, int n_cls, void* stream);

/* ---- polygonisation of the class raster (inference.py:356-407, the consumer of inference_and_write's output in
 * scripts/run_fast_aigle_segmentation.py:119) ----
 * Connected components, 4-connectivity (rasterio.features.shapes' default), all classes at once: labels int32 [H][W],
 * label = smallest linear pixel index of the component (order independent).  raster/labels on the device. */
int fz_ccl_label(const uint8_t* raster, int32_t* labels, int H, int W, void* stream);
/* area[root] += pixel count (area int32 [H*W], zeroed by the caller); *n_roots (zeroed) = number of components. */
int fz_ccl_areas(const int32_t* labels, int32_t* area_zeroed, int32_t* n_roots_zeroed, int H, int W, void* stream);
/* Compacted table of the components with area >= min_area_px and class != ignore_class (-1: none), arbitrary order:
 * records int32 [capacity][3] = (root, area, class); *counter (zeroed) ends as their number -- when it exceeds capacity
 * only the first `capacity` were stored: call again with a larger buffer. */
int fz_ccl_table(const uint8_t* raster, const int32_t* labels, const int32_t* area, int32_t* counter_zeroed,
                 int32_t* records, int capacity, int min_area_px, int ignore_class, int H, int W, void* stream);
/* HOST side of the same stage: boundary rings of the components listed in keep_roots (sorted ascending) on a label image
 * in host memory, in pixel-corner coordinates (x right, y down; vertex (x, y) = top-left corner of pixel (x, y)),
 * closed (first point repeated), exterior rings and hole rings (flagged), Douglas-Peucker simplified with tolerance
 * simplify_px (0 = corners only).  Results are held per thread until fz_trace_rings_fetch copies them out:
 * ring_root int32 [n_rings], ring_is_hole uint8 [n_rings], ring_offset int64 [n_rings + 1], xy float64 [n_points][2]. */
int fz_trace_rings(const int32_t* labels_host, int H, int W, const int32_t* keep_roots, int n_keep, double simplify_px,
                   int64_t* n_rings, int64_t* n_points);
int fz_trace_rings_fetch(int32_t* ring_root, uint8_t* ring_is_hole, int64_t* ring_offset, double* xy);

/* ---- training step, loss and optimizer side (the model's backward is not built yet) ----
 * tasks_module.py:153-154: targets int32 [B][H][W] = argmax over C of the one-hot float labels [B][C][H][W]. */
int fz_onehot_argmax(const float* onehot, int32_t* targets, int B, int C, int H, int W, void* stream);
/* nn.CrossEntropyLoss(weight=class_weight) of module_setup.py:155, 'mean' reduction = sum(w[t] nll) / sum(w[t]), times
 * task_weight (tasks_module.py:162-163), on fp32 NCHW logits.  lse float [B][H][W] (kept for the backward), preds int32
 * [B][H][W] = argmax(softmax(logits)) (tasks_module.py:159; may be NULL), workspace = fz_ce_workspace_doubles(B,H,W)
 * doubles, loss_out float[2] = {loss, sum of w[t]}.  Deterministic (fixed-order reductions in double). */
int64_t fz_ce_workspace_doubles(int B, int H, int W);
int fz_ce_loss_forward(const float* logits, const int32_t* targets, const float* class_weight, float task_weight, float* lse,
                       int32_t* preds, double* workspace, float* loss_out, int B, int C, int H, int W, void* stream);
/* dlogits = grad_scale * d loss / d logits = grad_scale * task_weight * w[t] / sum(w) * (softmax - onehot(t)). */
int fz_ce_loss_backward(const float* logits, const int32_t* targets, const float* class_weight, float task_weight,
                        const float* lse, const float* loss_out, float grad_scale, float* dlogits, int B, int C, int H,
                        int W, void* stream);
/* Confusion matrix of the training / validation metrics (tasks_module.py:74-90,212,274-275: torchmetrics
 * MulticlassJaccardIndex; prediction_writer.py:64: sklearn confusion_matrix(labels=range(C))): cm int64 [C][C] (device),
 * cm[t][p] += number of pixels with label t and prediction p; pixels with t or p outside 0..C-1 are skipped.  ACCUMULATES
 * (zero cm first); exact integer counts, independent of scheduling.  C <= 96. */
int fz_confusion_matrix(const int32_t* target, const int32_t* pred, int64_t n, int C, int64_t* cm, void* stream);
/* torch.optim.AdamW (tasks_module.py:385-389: lr, weight_decay, betas from the config; eps 1e-8) on one flat fp32
 * buffer, in place; step counts from 1. */
int fz_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, double weight_decay, int step, void* stream);
/* The same update with the step counter ON THE DEVICE: step_dev (int64, device) is incremented and the bias corrections are
 * computed by a one-thread kernel into hyper_dev (float[2], device), so the call has no step-dependent host argument and a
 * captured training step can be replayed as a CUDA graph. */
int fz_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                      double beta2, double eps, double weight_decay, int64_t* step_dev, float* hyper_dev, void* stream);

/* First building blocks of the model backward (a Linear layer's gradients through the K-major tcgen05 GEMM):
 * out bf16 [C][R] = in bf16 [R][C]^T (C <= 2 097 120); out float [N] = column sums of in bf16 [M][N]
 * (bias gradient; partial = float [chunks][N] workspace, fixed summation order). */
int fz_transpose_bf16(const void* in, void* out, int R, int C, void* stream);
int fz_colsum_bf16(const void* in, float* partial, float* out, int64_t M, int N, int chunks, void* stream);

/* ---- backward of one ConvNeXt-V2 block (flair_model.py:376 -> timm ConvNeXtBlock: dwconv7x7 -> LN -> fc1 -> GELU -> GRN ->
 * fc2 -> + x), first slice of the model backward; correctness-first kernels (csrc/backward_ops.cu), the Linear layers go
 * through fz_gemm_bf16 / the transposes above.  All activations NHWC = rows [B*H*W][C]. ----
 * depthwise 7x7 in fp32: out = bias + sum_k in(shifted) * w[k][c]; flip = 1 uses w[48-k] (= the data gradient). */
int fz_dwconv7_f32(const float* in, const float* w, const float* bias, float* out, int B, int H, int W, int C, int flip,
                   void* stream);
/* the same with `add` (float, out's layout, may be NULL) summed into the result: the block backward's dx = dy + dwconv^T(du)
 * in one pass.  H % 4 == 0 and W % 8 == 0 take the 4 x 8 register-tiled kernel. */
int fz_dwconv7_f32_add(const float* in, const float* w, const float* bias, const float* add, float* out, int B, int H, int W,
                       int C, int flip, void* stream);
/* dw float [49][C], db float [C]: gradients of the depthwise weights / bias from x and the output gradient du.  The one entry
 * point that owns device memory: a per-device scratch buffer for its fixed-order partial sums, grown on demand. */
int fz_dwconv7_wgrad(const float* x, const float* du, float* dw, float* db, int B, int H, int W, int C, void* stream);
/* LayerNorm over C with saved row statistics (training forward), and its backward: dx float [M][C]; dgamma_dbeta float [2][C]
 * (partial = float [blocks][2][C] workspace; blocks fixes the reduction order). */
int fz_layernorm_fwd_stats(const float* x, const float* g, const float* b, void* out_bf16, void* out2_bf16, float* mean,
                           float* rstd, int64_t M, int C, float eps, int out_f16, void* stream);
int fz_layernorm_bwd(const void* dy_bf16, const float* x, const float* mean, const float* rstd, const float* g, float* dx,
                     float* partial, float* dgamma_dbeta, int64_t M, int C, int blocks, void* stream);
/* exact (erf) GELU on bf16. */
int fz_gelu_fwd(const void* h_bf16, void* g_bf16, int64_t n, void* stream);
/* out float [B][C] = sum over a sample's HW rows of a*a (mode 0), a*b (1) or a (2); a, b bf16 [B][HW][C]. */
int fz_sample_colreduce(const void* a_bf16, const void* b_bf16, float* out, int B, int HW, int C, int mode, void* stream);
/* s1 = sum_hw a*b and s0 = sum_hw a in ONE pass over both tensors (C % 8 == 0): the GRN backward's two reductions. */
int fz_sample_colreduce2(const void* a_bf16, const void* b_16, float* s1, float* s0, int B, int HW, int C, int b_f16,
                         void* stream);
/* g = GELU(h) (bf16, stored), optionally GELU'(h) (bf16, stored; NULL = not wanted), and sumsq float [B][C] = sum_hw g^2 of
 * the stored values, one pass (C % 8 == 0).  erf through Abramowitz-Stegun 7.1.26, |Phi error| <= 3e-7. */
int fz_gelu_fwd_sumsq(const void* h_bf16, void* g_bf16, void* dgelu_bf16, float* sumsq, int B, int HW, int C, int act_f16,
                      void* stream);
/* GRN, training forward: gx = sqrt(sumsq), mu = mean_c gx, nx = gx / (mu + eps), y = g (1 + gamma nx) + beta. */
int fz_grn_train_forward(const void* g_bf16, const float* sumsq, const float* gamma, const float* beta, float* gx, float* nx,
                         float* mu, void* y_bf16, void* y2_bf16, int B, int HW, int C, float eps, int act_f16, void* stream);
/* GRN + GELU backward: dy = gradient at the GRN output, g = GELU(h), s1 = sum_hw dy g, s0 = sum_hw dy (fz_sample_colreduce);
 * dh bf16 = gradient at the pre-GELU activations; dgamma / dbeta float [C]; coef_a / coef_b float [B][C] workspaces. */
int fz_grn_gelu_backward(const void* dy_bf16, const void* g_bf16, const void* h_bf16, const float* s1, const float* s0,
                         const float* gx, const float* nx, const float* mu, const float* gamma, float* coef_a, float* coef_b,
                         float* dgamma, float* dbeta, void* dh_bf16, int B, int HW, int C, float eps, void* stream);
/* the same, and dbias float [C] (may be NULL) = column sums of the stored dh over all B*HW rows: the bias gradient of the
 * Linear that produced h, without another pass over dh. */
int fz_grn_gelu_backward_db(const void* dy_bf16, const void* g_bf16, const void* h_bf16, const float* s1, const float* s0,
                            const float* gx, const float* nx, const float* mu, const float* gamma, float* coef_a,
                            float* coef_b, float* dgamma, float* dbeta, void* dh_bf16, float* dbias, int B, int HW, int C,
                            float eps, void* stream);
/* the same with GELU'(h) as saved by fz_gelu_fwd_sumsq in place of h: no transcendental in the backward pass. */
int fz_grn_gelu_backward_saved(const void* dy_bf16, const void* g_bf16, const void* dgelu_bf16, const float* s1,
                               const float* s0, const float* gx, const float* nx, const float* mu, const float* gamma,
                               float* coef_a, float* coef_b, float* dgamma, float* dbeta, void* dh_bf16, float* dbias, int B,
                               int HW, int C, float eps, int act_f16, void* stream);
/* ---- U-Net decoder convolutions of the training step without im2col (csrc/conv3x3_small.cu): 3x3 / pad 1 on bf16 NHWC maps
 * with H % 8 == 0, W % 32 == 0 and 16 / 32 / 48 / 64 channels on either side (fz_conv3x3_small_supported: 1 / 0).
 * forward: out[px][co] = bias[co] + sum_tap sum_ci in[px + tap][ci] * w[tap][co][ci], w bf16 [9][Cout][Cin], tap = ky*3 + kx;
 * out fp32 (out_bf16 = 0) or bf16 [pixels][ldo], columns < n_store written.  The data gradient is the same call on the output
 * gradient with w'[tap][ci][co] = w[8 - tap][co][ci].
 * wgrad: dw float [9][Cout][Cin] = sum_px dconv[px][co] * x[px + tap][ci]; dconv bf16 [pixels][ldd]; Cout <= 32; partial sums
 * live in a per-device scratch buffer grown on demand and are added in a fixed order. */
int fz_conv3x3_small_supported(int H, int W, int Cin, int Cout);
int fz_conv3x3_small_forward(const void* in_bf16, const void* w_bf16, const float* bias, void* out, int out_bf16, int B, int H,
                             int W, int Cin, int Cout, int n_store, int ldo, int in_f16, void* stream);
int fz_conv3x3_small_wgrad(const void* x_bf16, const void* dconv_bf16, int ldd, float* dw, int B, int H, int W, int Cin,
                           int Cout, int x_f16, void* stream);
/* IEEE fp16 -> bf16 (round to nearest even), n values, 16-byte aligned buffers: the forward activations on their way into a
 * weight-gradient GEMM whose other operand is a bf16 gradient. */
int fz_cast_f16_bf16(const void* in_f16, void* out_bf16, int64_t n, void* stream);
int fz_add_f32(const float* a, const float* b, float* out, int64_t n, void* stream);
/* out float [N] = sum over s of partial float [S][N], in the order s = 0 .. S-1 (split reductions stay reproducible). */
int fz_reduce_rows_f32(const float* partial, float* out, int N, int S, void* stream);
/* FORMATS of the training step's 16-bit tensors.  Gradients are bf16 (range).  FORWARD activations -- LayerNorm outputs, the
 * hidden tensors h / GELU(h) / GELU'(h) / GRN output, decoder activations, stem patches, im2col rows -- are bf16 or, where an
 * entry point takes an `act_f16` / `out_f16` / `in_f16` / `x_f16` / `b_f16` flag set to 1, IEEE fp16 (the trainer's choice:
 * tests/diag/grad_precision_budget.py shows the bf16 rounding of the forward tensors alone costs the gradients 0.8 % of
 * cosine against fp32 autograd, fp16 0.1 %).  The pointer parameters keep their *_bf16 names.  fz_bn_relu_backward reads its
 * `y_bf16` for the sign only and accepts either format.  out2_bf16 / y2_bf16 (may be NULL): a second copy of the output,
 * always bf16 -- the operand of the weight-gradient GEMM, whose other operand is a bf16 gradient (one MMA, one format).
 * Encoder plumbing of the same slice: LayerNorm with a bf16 and / or an fp32 output (LayerNorm2d of the stem feeds the fp32
 * residual stream, the one in front of a downsample conv feeds a GEMM); space-to-depth for the 2x2/s2 convolutions as GEMMs,
 * k = (ky*s + kx)*C + c, and its inverse (inverse = 1: `in` is the [B][H/s][W/s][s*s*C] side); the stem's 4x4/s4 patches of a
 * normalised fp32 NCHW tile as bf16 rows [B*(P/4)^2][Kpad], k = (c*4 + ky)*4 + kx, zero padded to Kpad. */
int fz_layernorm_fwd_stats2(const float* x, const float* g, const float* b, void* out_bf16, float* out_f32, float* mean,
                            float* rstd, int64_t M, int C, float eps, int out_f16, void* stream);
int fz_s2d_bf16(const void* in, void* out, int B, int H, int W, int C, int s, int inverse, void* stream);
int fz_patchify4_nchw(const float* in, void* out_bf16, int B, int Cin, int P, int Kpad, int out_f16, void* stream);
/* U-Net decoder, training mode (smp UnetDecoder block: nearest x2 -> concat skip -> [conv3x3 -> BatchNorm -> ReLU] x 2):
 * the 3x3 convolutions run as GEMMs over an explicit im2col (correctness-first; col bf16 [B*H*W][Kpad], k = (ky*3+kx)*C + c)
 * with col2im as the data gradient; BatchNorm uses the batch statistics (biased variance) and keeps mean / rstd; x is the
 * convolution output, fp32 with row stride ldx (the GEMM's padded N).  workspace: (chunks + 1) * 2 * C floats.
 * fz_bn_relu_backward: dx bf16 with row stride ldd (padding columns zeroed), dbeta_dgamma float [2][C]. */
int fz_im2col3x3_bf16(const void* in, void* col, int B, int H, int W, int C, int Kpad, void* stream);
int fz_col2im3x3(const void* dcol_bf16, float* dx, int B, int H, int W, int C, int Kpad, void* stream);
int fz_bn_relu_train_forward(const float* x, int ldx, const float* gamma, const float* beta, void* y_bf16, float* mean,
                             float* rstd, float* workspace, int64_t M, int C, int chunks, float eps, int out_f16,
                             void* stream);
int fz_bn_relu_backward(const float* x, int ldx, const void* dy_bf16, const void* y_bf16, const float* mean, const float* rstd,
                        const float* gamma, void* dx_bf16, int ldd, float* dbeta_dgamma, float* workspace, int64_t M, int C,
                        int chunks, void* stream);
/* BatchNorm2d's running statistics after a training forward (momentum 0.1 in smp's decoder): running_mean / running_var are
 * updated in place from the batch mean and rstd of fz_bn_relu_train_forward (unbiased variance, like PyTorch). */
int fz_bn_update_running(const float* mean, const float* rstd, float* running_mean, float* running_var, int C, int64_t M,
                         float eps, float momentum, void* stream);
/* gradient of fz_upsample2_concat: da float [B][H][W][C1] (2x2 sums), dskip float [B][2H][2W][C2] from dcat float
 * [B][2H][2W][C1+C2]. */
int fz_upsample2_concat_backward(const float* dcat, float* da, float* dskip, int B, int H, int W, int C1, int C2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLAIR_ZONAL_B200_H */

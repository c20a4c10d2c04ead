// CUDA-core kernels around the tensor-core GEMMs of the ConvNeXt-V2 encoder and U-Net decoder:
//   stem (conv4x4 s4 on raw uint8 with the normalisation folded in, + LayerNorm2d),
//   depthwise 7x7 + LayerNorm fused, LayerNorm2d + space-to-depth for the 2x2/s2 downsample,
//   GRN statistics -> per-sample scales, weight / activation scaling, nearest-up x2 + concat.
// Activations are NHWC; the residual stream is fp32, GEMM operands are bf16.
//
// Replaces the timm ConvNeXtBlock / ConvNeXtStage arithmetic invoked from
// flair_hub/models/flair_model.py:376 and smp UnetDecoder's interpolate+cat invoked from :418
// (SURVEY.md K1/K4).
#include "common.h"
#include "../../include/flair_zonal_b200.h"

#include "ptx.cuh"
#include "operand.cuh"

namespace fz {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------ stem
// in  : uint8 [B][P][P][4]           (fz_gather_tiles_u8)
// w   : float [64][C0], row k = ky*16 + kx*4 + c, already divided by std[c]
// bias: float [C0] = conv bias - sum_k w[k][n] * mean[c(k)]
// out : float [B][P/4][P/4][C0] = LayerNorm2d(conv)
constexpr int STEM_PX = 32;        // output pixels per segment (8 per warp)
constexpr int STEM_IN_STRIDE = 80; // floats per pixel row in smem (64 used; 80 keeps stores conflict-free)

template <int CPL, bool F32IN>
__global__ void __launch_bounds__(128) stem_ln_kernel(const void* __restrict__ in_raw, int Cin,
                                                      const float* __restrict__ w, const float* __restrict__ bias,
                                                      const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                      float* __restrict__ out, int P, float eps) {
  constexpr int C0 = 32 * CPL;
  extern __shared__ float smem_f[];
  float* sW = smem_f;                 // [64][C0]
  float* sIn = smem_f + 64 * C0;      // [STEM_PX][STEM_IN_STRIDE]
  const int OW = P / 4;
  const int segs = OW / STEM_PX;
  const int oy = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 64 * C0; i += 128) sW[i] = w[i];
  float bz[CPL], gw[CPL], gb[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    bz[j] = bias[lane * CPL + j];
    gw[j] = ln_w[lane * CPL + j];
    gb[j] = ln_b[lane * CPL + j];
  }

  for (int seg = 0; seg < segs; ++seg) {
    const int ox0 = seg * STEM_PX;
    __syncthreads();
    // 4 input rows x 128 input px: thread t loads px (4*ox0 + t) of each row
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      if (F32IN) {
        // float32 NCHW [B][Cin][P][P]: one coalesced load per channel
        const float* xin = reinterpret_cast<const float*>(in_raw);
        const size_t o = (static_cast<size_t>(b) * Cin * P + (4 * oy + r)) * P + 4 * ox0 + tid;
        const size_t plane = static_cast<size_t>(P) * P;
        f.x = xin[o];
        if (Cin > 1) f.y = xin[o + plane];
        if (Cin > 2) f.z = xin[o + 2 * plane];
        if (Cin > 3) f.w = xin[o + 3 * plane];
      } else {
        const uchar4 u =
            reinterpret_cast<const uchar4*>(in_raw)[(static_cast<size_t>(b) * P + (4 * oy + r)) * P + 4 * ox0 + tid];
        f = make_float4(u.x, u.y, u.z, u.w);
      }
      *reinterpret_cast<float4*>(&sIn[(tid >> 2) * STEM_IN_STRIDE + r * 16 + (tid & 3) * 4]) = f;
    }
    __syncthreads();

    float acc[8][CPL];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[i][j] = bz[j];

#pragma unroll 4
    for (int k = 0; k < 64; k += 4) {
      float wv[4][CPL];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (CPL % 4 == 0) {
#pragma unroll
          for (int j = 0; j < CPL; j += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(&sW[(k + kk) * C0 + lane * CPL + j]);
            wv[kk][j] = t4.x; wv[kk][j + 1] = t4.y; wv[kk][j + 2] = t4.z; wv[kk][j + 3] = t4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < CPL; ++j) wv[kk][j] = sW[(k + kk) * C0 + lane * CPL + j];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 x4 = *reinterpret_cast<const float4*>(&sIn[(warp * 8 + i) * STEM_IN_STRIDE + k]);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          acc[i][j] = fmaf(x4.x, wv[0][j], acc[i][j]);
          acc[i][j] = fmaf(x4.y, wv[1][j], acc[i][j]);
          acc[i][j] = fmaf(x4.z, wv[2][j], acc[i][j]);
          acc[i][j] = fmaf(x4.w, wv[3][j], acc[i][j]);
        }
      }
    }
    // LayerNorm2d over the C0 channels of each pixel (two-pass, fp32)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) s += acc[i][j];
      const float mean = warp_sum(s) * (1.0f / C0);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const float d = acc[i][j] - mean;
        q = fmaf(d, d, q);
      }
      const float rstd = rsqrtf(warp_sum(q) * (1.0f / C0) + eps);
      float* o = out + ((static_cast<size_t>(b) * OW + oy) * OW + ox0 + warp * 8 + i) * C0 + lane * CPL;
#pragma unroll
      for (int j = 0; j < CPL; ++j) o[j] = (acc[i][j] - mean) * rstd * gw[j] + gb[j];
    }
  }
}

// ------------------------------------------------------------------------------------ stem on tensor cores (uint8 input)
// ncu-level arithmetic: the stem is 64 x C0 MACs per output pixel = 9.9 GFLOP per batch of 37 tiles, which the fp32 FMA
// kernel above executes at 33 TFLOP/s (295 us, 2.2 % of the step) while its HBM traffic (39 MB in, 310 MB out) needs 70 us.
// Here the same sums run on mma.sync.m16n8k8 TF32 with the 3xTF32 split of the WEIGHTS (w = w_hi + w_lo, both TF32; the
// uint8 inputs are exact in TF32), fp32 accumulation: every product is exact, the result differs from the fp32 kernel by
// summation order only.  Block = one output row of one sample, 4 warps x 16 pixels per iteration (P % 256 == 0).
constexpr int STEM_MMA_PX = 64;       // output pixels per block iteration
constexpr int STEM_MMA_LDI = 68;      // floats per pixel row of the input tile (64 used; 68 = 4 mod 32: conflict-free A loads)

__device__ __forceinline__ void mma_tf32_1688(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NB>   // C0 = 8 * NB output channels
__global__ void __launch_bounds__(128, 2) stem_ln_mma_kernel(const uchar4* __restrict__ in, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const float* __restrict__ ln_w,
                                                             const float* __restrict__ ln_b, float* __restrict__ out, int P,
                                                             int n_rows, float eps) {
  constexpr int C0 = 8 * NB, LDW = C0 + 8;                       // LDW = 8 mod 32: conflict-free B loads
  extern __shared__ float smem_f[];
  float* sHi = smem_f;                                           // [64][LDW]
  float* sLo = sHi + 64 * LDW;                                   // [64][LDW]
  float* sIn = sLo + 64 * LDW;                                   // [STEM_MMA_PX][STEM_MMA_LDI]
  float* sPar = sIn + STEM_MMA_PX * STEM_MMA_LDI;                // bias | ln_w | ln_b, C0 each
  const int OW = P / 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  // persistent blocks (one wave): the 64 x C0 weights are split and staged ONCE per block, not once per output row -- with
  // a block per row the staging was most of the kernel (227 us for 4736 rows; ncu launch list of the bench)
  for (int i = tid; i < 64 * C0; i += 128) {
    const int k = i / C0, n = i - k * C0;
    const float v = w[i];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);     // the TF32 part of the weight
    sHi[k * LDW + n] = hi;
    sLo[k * LDW + n] = v - hi;                                               // exact in fp32; its TF32 part is what the MMA reads
  }
  for (int i = tid; i < C0; i += 128) {
    sPar[i] = bias[i];
    sPar[C0 + i] = ln_w[i];
    sPar[2 * C0 + i] = ln_b[i];
  }
  for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
  const int b = row / OW, oy = row - b * OW;                     // output maps are square: OH = OW
  for (int ox0 = 0; ox0 < OW; ox0 += STEM_MMA_PX) {
    __syncthreads();                                             // weights staged / previous iteration's reads done
    // 4 input rows x 256 input pixels of 4 bytes: thread -> input pixels tid and tid + 128 of each row
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ip = tid + h * 128;                            // input pixel within the 256-pixel span
        const uchar4 u = in[(static_cast<size_t>(b) * P + (4 * oy + r)) * P + 4 * ox0 + ip];
        *reinterpret_cast<float4*>(&sIn[(ip >> 2) * STEM_MMA_LDI + r * 16 + (ip & 3) * 4]) = make_float4(u.x, u.y, u.z, u.w);
      }
    __syncthreads();
    float acc[NB][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      acc[nb][0] = acc[nb][2] = sPar[nb * 8 + 2 * t];
      acc[nb][1] = acc[nb][3] = sPar[nb * 8 + 2 * t + 1];
    }
    const float* arow = sIn + (warp * 16 + g) * STEM_MMA_LDI + t;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t a[4];
      a[0] = __float_as_uint(arow[ks * 8]);
      a[1] = __float_as_uint(arow[8 * STEM_MMA_LDI + ks * 8]);
      a[2] = __float_as_uint(arow[ks * 8 + 4]);
      a[3] = __float_as_uint(arow[8 * STEM_MMA_LDI + ks * 8 + 4]);
      const float* bh = sHi + (ks * 8 + t) * LDW + g;
      const float* bl = sLo + (ks * 8 + t) * LDW + g;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        mma_tf32_1688(acc[nb], a, __float_as_uint(bh[nb * 8]), __float_as_uint(bh[4 * LDW + nb * 8]));
        mma_tf32_1688(acc[nb], a, __float_as_uint(bl[nb * 8]), __float_as_uint(bl[4 * LDW + nb * 8]));
      }
    }
    // LayerNorm2d over the C0 channels of each pixel: a pixel's channels sit in the 4 lanes of one g (rows g and g + 8)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float s = 0.f;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) s += acc[nb][half * 2] + acc[nb][half * 2 + 1];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      const float mean = s * (1.0f / C0);
      float q = 0.f;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const float d0 = acc[nb][half * 2] - mean, d1 = acc[nb][half * 2 + 1] - mean;
        q = fmaf(d0, d0, q);
        q = fmaf(d1, d1, q);
      }
      q += __shfl_xor_sync(0xffffffffu, q, 1);
      q += __shfl_xor_sync(0xffffffffu, q, 2);
      const float rstd = rsqrtf(q * (1.0f / C0) + eps);
      float* o = out + ((static_cast<size_t>(b) * OW + oy) * OW + ox0 + warp * 16 + g + half * 8) * C0 + 2 * t;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int c = nb * 8 + 2 * t;
        *reinterpret_cast<float2*>(o + nb * 8) =
            make_float2((acc[nb][half * 2] - mean) * rstd * sPar[C0 + c] + sPar[2 * C0 + c],
                        (acc[nb][half * 2 + 1] - mean) * rstd * sPar[C0 + c + 1] + sPar[2 * C0 + c + 1]);
      }
    }
  }
  }
}

// ------------------------------------------------------------------------------------ dwconv7x7 + LN
// x  : float [B][H][W][C] residual stream;  wdw: float [49][C];  bdw: float [C]
// out: bf16  [B][H][W][C] = LayerNorm_C(dwconv7x7(x) + bdw) * ln_w + ln_b
//
// One CTA = one 4 x SW pixel tile, ALL channels: warp w owns channels [32w, 32w+32) (and, for
// CPT = 2, the same range in the upper half of C), lane = channel, so every global access is a
// coalesced 128 B row and no shared-memory staging of the input is needed.  Each thread keeps
// the 4 x SW outputs of its channel(s) in registers and walks the (4+6) x (SW+6) input window
// once (14 loads feed up to 224 FMAs).  LayerNorm statistics: a 31-shuffle transpose-reduce
// gives per-pixel sums over the warp's channels, a tiny shared array combines the warps
// (fixed order -> deterministic), two passes (mean, then centred variance) in fp32.
template <int OFF>
__device__ __forceinline__ void colsum32_step(float (&s)[32], int lane) {
  const bool upper = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < OFF; ++i) {
    const float a = s[i], b = s[i + OFF];
    s[i] = (upper ? b : a) + __shfl_xor_sync(0xffffffffu, upper ? a : b, OFF);
  }
}
// on return s[0] of lane l = sum over lanes of the input s[l]
__device__ __forceinline__ void warp_colsum32(float (&s)[32], int lane) {
  colsum32_step<16>(s, lane);
  colsum32_step<8>(s, lane);
  colsum32_step<4>(s, lane);
  colsum32_step<2>(s, lane);
  colsum32_step<1>(s, lane);
}

template <int NW, int CPT, int C, bool PERSIST>
__global__ void __launch_bounds__(NW * 32, (512 / (NW * 32)) > 0 ? 512 / (NW * 32) : 1) dwconv_ln_kernel(const float* __restrict__ x, const float* __restrict__ wdw,
                                                            const float* __restrict__ bdw,
                                                            const float* __restrict__ ln_w,
                                                            const float* __restrict__ ln_b,
                                                            op_t* __restrict__ out, int H, int W, float eps,
                                                            int tiles_x, int tiles_y, int n_tiles) {
  static_assert(NW * 32 * CPT == C, "one lane per channel (two for CPT = 2)");
  constexpr int SH = 4;
  constexpr int SW = 8 / CPT;      // 32 accumulators per thread either way
  constexpr int NPX = SH * SW;     // pixels per tile (32 or 16)
  __shared__ float red[NW][32];
  __shared__ float s_stat[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // PERSIST (one channel per lane only): the CTA walks a list of tiles and loads the 49 taps + bias ONCE -- they were 26 %
  // of all loads and a dependent prologue in front of each tile's first FMA (stage 2, C = 512: 86 -> 78 us; C = 128 with
  // four small CTAs per SM measured slower, 271 -> 283 us, and stays one tile per CTA)
  static_assert(!PERSIST || CPT == 1, "the persistent walk keeps one channel's taps in registers");
  float wreg[49];
  float bias = 0.f;
  if (PERSIST) {
    const int c = warp * 32 + lane;
#pragma unroll
    for (int t = 0; t < 49; ++t) wreg[t] = wdw[t * C + c];
    bias = bdw[c];
  }
  int tile = blockIdx.x;
  do {
  const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
  const int x0 = tx * SW, y0 = ty * SH;
  const float* xb = x + static_cast<size_t>(b) * H * W * C;
  // CTA-uniform edge flags: out-of-image taps are zero (conv padding); no per-load predicate math
  const bool left_edge = x0 < 3, right_edge = x0 + SW + 3 > W;

  float acc[CPT][NPX];
#pragma unroll
  for (int cc = 0; cc < CPT; ++cc) {
    const int c = cc * (C / CPT) + warp * 32 + lane;
    if (!PERSIST) {
#pragma unroll
      for (int t = 0; t < 49; ++t) wreg[t] = wdw[t * C + c];
      bias = bdw[c];
    }
#pragma unroll
    for (int p = 0; p < NPX; ++p) acc[cc][p] = bias;
    // software-pipelined rows: row iy+1 is in flight while row iy feeds the FMAs
    float nxt[SW + 6];
    // running row pointer: one 64-bit add per input row, every tap is [row + immediate] (C is constexpr)
    const ptrdiff_t row_stride = static_cast<ptrdiff_t>(W) * C;
    const float* row0 = xb + (static_cast<ptrdiff_t>(y0 - 3) * W + (x0 - 3)) * C + c;
    auto load_row = [&](int iy, float (&dst)[SW + 6]) {
      const int gy = y0 - 3 + iy;
      if (gy < 0 || gy >= H) {                               // CTA-uniform: zero padding row
#pragma unroll
        for (int ix = 0; ix < SW + 6; ++ix) dst[ix] = 0.0f;
        return;
      }
      const float* row = row0 + iy * row_stride;
#pragma unroll
      for (int ix = 0; ix < SW + 6; ++ix) {
        if (ix < 3) dst[ix] = (left_edge && x0 - 3 + ix < 0) ? 0.0f : row[ix * C];
        else if (ix >= SW + 3) dst[ix] = (right_edge && x0 - 3 + ix >= W) ? 0.0f : row[ix * C];
        else dst[ix] = row[ix * C];
      }
    };
    load_row(0, nxt);
#pragma unroll
    for (int iy = 0; iy < SH + 6; ++iy) {
      float in[SW + 6];
#pragma unroll
      for (int ix = 0; ix < SW + 6; ++ix) in[ix] = nxt[ix];
      if (iy + 1 < SH + 6) load_row(iy + 1, nxt);
#pragma unroll
      for (int oy = 0; oy < SH; ++oy) {
        const int ky = iy - oy;
        if (ky >= 0 && ky < 7) {
#pragma unroll
          for (int kx = 0; kx < 7; ++kx)
#pragma unroll
            for (int ox = 0; ox < SW; ++ox) acc[cc][oy * SW + ox] = fmaf(in[ox + kx], wreg[ky * 7 + kx], acc[cc][oy * SW + ox]);
        }
      }
    }
  }

  // ---- LayerNorm over C per pixel.  s[] layout: CPT == 1: s[p]; CPT == 2: s[cc*16 + p]
  float s[32];
#pragma unroll
  for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
    for (int p = 0; p < NPX; ++p) s[cc * NPX + p] = acc[cc][p];
  warp_colsum32(s, lane);
  float part = s[0];
  if (CPT == 2) part += __shfl_xor_sync(0xffffffffu, part, 16);   // lane l and l^16 hold the two channel halves
  red[warp][lane] = part;
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < NW; ++w2) t += red[w2][lane];
    s_stat[lane] = t / C;                                           // mean of pixel (lane % NPX)
  }
  __syncthreads();
#pragma unroll
  for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
    for (int p = 0; p < NPX; ++p) {
      acc[cc][p] -= s_stat[p];
      s[cc * NPX + p] = acc[cc][p] * acc[cc][p];
    }
  warp_colsum32(s, lane);
  part = s[0];
  if (CPT == 2) part += __shfl_xor_sync(0xffffffffu, part, 16);
  __syncthreads();                                                  // everyone has read the means
  red[warp][lane] = part;
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < NW; ++w2) t += red[w2][lane];
    s_stat[lane] = rsqrtf(t / C + eps);
  }
  __syncthreads();
#pragma unroll
  for (int cc = 0; cc < CPT; ++cc) {
    const int c = cc * (C / CPT) + warp * 32 + lane;
    const float g = ln_w[c], be = ln_b[c];
    op_t* o0 = out + ((static_cast<size_t>(b) * H + y0) * W + x0) * C + c;
    const ptrdiff_t orow = static_cast<ptrdiff_t>(W) * C;
#pragma unroll
    for (int oy = 0; oy < SH; ++oy) {
      op_t* orp = o0 + oy * orow;
#pragma unroll
      for (int ox = 0; ox < SW; ++ox)
        orp[ox * C] = f2op(acc[cc][oy * SW + ox] * s_stat[oy * SW + ox] * g + be);
    }
  }
    tile += gridDim.x;
  } while (PERSIST && tile < n_tiles);   // the next tile's first __syncthreads orders its shared-memory writes after these reads
}

// ------------------------------------------------------------------------------------ dwconv7x7 + LN, channel pairs
// Round 2.  ncu on the kernel above (profiles/r2_ncu_dwconv_full_summary.txt): DRAM traffic = algorithmic, FMA pipe 43.6 %,
// issue slots 63 %, and only 58 % of the 2690 instructions a thread executes for its 1568 FMAs ARE FMAs -- the rest is
// per-channel overhead: 140 scalar loads, 32 two-byte stores, the LayerNorm shuffles.  Here a lane owns the channel PAIR
// (2l, 2l+1) of its warp's 64 channels: every load is an 8-byte float2 (still a coalesced 256 B row per warp), every FMA is
// Blackwell's packed fma.rn.f32x2 (two IEEE fp32 FMAs in one instruction: same bits as the scalar kernel, same summation
// order), every store one packed 16-bit pair, and the LayerNorm statistics of the two channels share their shuffles.
// Tile 2 x 8 pixels x 2 channels = 16 float2 accumulators (the 4 x 8 tile of the scalar kernel would need 150 registers and
// halve the occupancy); the 49 x C taps live in shared memory (a lane reads its float2 per tap: 7 LDS.64 per input row and
// output row), loaded once by the persistent CTA.
template <int NW, int C, int SH>
__global__ void __launch_bounds__(NW * 32, SH == 2 ? 512 / (NW * 32) : (NW >= 8 ? 1 : 256 / (NW * 32))) dwconv_ln_pair_kernel(const float* __restrict__ x, const float* __restrict__ wdw,
                                                                 const float* __restrict__ bdw,
                                                                 const float* __restrict__ ln_w,
                                                                 const float* __restrict__ ln_b, op_t* __restrict__ out,
                                                                 int H, int W, float eps, int tiles_x, int tiles_y,
                                                                 int n_tiles) {
  static_assert(NW * 64 == C, "one lane per channel pair");
  static_assert(SH == 2 || SH == 4, "2 x 8 or 4 x 8 pixel tiles");
  constexpr int SW = 8, NPX = SH * SW;
  extern __shared__ float s_taps[];                    // [49][C]
  __shared__ float red[NW][32];
  __shared__ float s_stat[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x * 4; i < 49 * C; i += NW * 32 * 4)
    *reinterpret_cast<float4*>(s_taps + i) = *reinterpret_cast<const float4*>(wdw + i);
  const int c = warp * 64 + lane * 2;
  const float2 bias = *reinterpret_cast<const float2*>(bdw + c);
  const float2 g2 = *reinterpret_cast<const float2*>(ln_w + c), be2 = *reinterpret_cast<const float2*>(ln_b + c);
  __syncthreads();
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int x0 = tx * SW, y0 = ty * SH;
    const float* xb = x + static_cast<size_t>(b) * H * W * C;
    const bool left_edge = x0 < 3, right_edge = x0 + SW + 3 > W;
    float2 acc[NPX];
#pragma unroll
    for (int p = 0; p < NPX; ++p) acc[p] = bias;
    float2 nxt[SW + 6];
    const ptrdiff_t row_stride = static_cast<ptrdiff_t>(W) * C;
    const float* row0 = xb + (static_cast<ptrdiff_t>(y0 - 3) * W + (x0 - 3)) * C + c;
    auto load_row = [&](int iy, float2 (&dst)[SW + 6]) {
      const int gy = y0 - 3 + iy;
      if (gy < 0 || gy >= H) {                               // CTA-uniform: zero padding row
#pragma unroll
        for (int ix = 0; ix < SW + 6; ++ix) dst[ix] = make_float2(0.f, 0.f);
        return;
      }
      const float* row = row0 + iy * row_stride;
#pragma unroll
      for (int ix = 0; ix < SW + 6; ++ix) {
        const bool pad = (ix < 3 && left_edge && x0 - 3 + ix < 0) || (ix >= SW + 3 && right_edge && x0 - 3 + ix >= W);
        dst[ix] = pad ? make_float2(0.f, 0.f) : *reinterpret_cast<const float2*>(row + ix * C);
      }
    };
    load_row(0, nxt);
#pragma unroll
    for (int iy = 0; iy < SH + 6; ++iy) {
      float2 in[SW + 6];
#pragma unroll
      for (int ix = 0; ix < SW + 6; ++ix) in[ix] = nxt[ix];
      if (iy + 1 < SH + 6) load_row(iy + 1, nxt);
#pragma unroll
      for (int oy = 0; oy < SH; ++oy) {
        const int ky = iy - oy;
        if (ky >= 0 && ky < 7) {
          float2 wk[7];
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) wk[kx] = *reinterpret_cast<const float2*>(s_taps + (ky * 7 + kx) * C + c);
#pragma unroll
          for (int kx = 0; kx < 7; ++kx)
#pragma unroll
            for (int ox = 0; ox < SW; ++ox) acc[oy * SW + ox] = __ffma2_rn(in[ox + kx], wk[kx], acc[oy * SW + ox]);
        }
      }
    }
    // ---- LayerNorm over C per pixel.  SH = 2: s[cc * 16 + p] = channel cc of the pair (like the CPT = 2 path above);
    //      SH = 4: the pair is summed first, s[p] over 32 pixels
    float s[32];
    float part;
    if (SH == 2) {
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        s[p] = acc[p].x;
        s[16 + p] = acc[p].y;
      }
      warp_colsum32(s, lane);
      part = s[0] + __shfl_xor_sync(0xffffffffu, s[0], 16);
    } else {
#pragma unroll
      for (int p = 0; p < 32; ++p) s[p] = acc[p % NPX].x + acc[p % NPX].y;
      warp_colsum32(s, lane);
      part = s[0];
    }
    red[warp][lane] = part;
    __syncthreads();
    if (warp == 0) {
      float t = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < NW; ++w2) t += red[w2][lane];
      s_stat[lane] = t / C;                                           // mean of pixel (lane % NPX)
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < NPX; ++p) {
      acc[p].x -= s_stat[p];
      acc[p].y -= s_stat[p];
    }
    if (SH == 2) {
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        s[p] = acc[p].x * acc[p].x;
        s[16 + p] = acc[p].y * acc[p].y;
      }
      warp_colsum32(s, lane);
      part = s[0] + __shfl_xor_sync(0xffffffffu, s[0], 16);
    } else {
#pragma unroll
      for (int p = 0; p < 32; ++p) s[p] = acc[p % NPX].x * acc[p % NPX].x + acc[p % NPX].y * acc[p % NPX].y;
      warp_colsum32(s, lane);
      part = s[0];
    }
    __syncthreads();                                                  // everyone has read the means
    red[warp][lane] = part;
    __syncthreads();
    if (warp == 0) {
      float t = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < NW; ++w2) t += red[w2][lane];
      s_stat[lane] = rsqrtf(t / C + eps);
    }
    __syncthreads();
    op_t* o0 = out + ((static_cast<size_t>(b) * H + y0) * W + x0) * C + c;
    const ptrdiff_t orow = static_cast<ptrdiff_t>(W) * C;
#pragma unroll
    for (int oy = 0; oy < SH; ++oy) {
      op_t* orp = o0 + oy * orow;
#pragma unroll
      for (int ox = 0; ox < SW; ++ox) {
        const float r = s_stat[oy * SW + ox];
        *reinterpret_cast<op2_t*>(orp + ox * C) = ff2op2(acc[oy * SW + ox].x * r * g2.x + be2.x, acc[oy * SW + ox].y * r * g2.y + be2.y);
      }
    }
    // the next tile's first __syncthreads orders its shared-memory writes after these reads of s_stat
  }
}

// ------------------------------------------------------------------------------------ LN2d + space-to-depth
// x: float [B][H][W][C] -> out: bf16 [B][H/2][W/2][4C], k = (y&1)*2C + (x&1)*C + c  (the K order of
// the 2x2/s2 conv weights repacked as [N][ky][kx][c]).  One warp per pixel.
__global__ void __launch_bounds__(256) ln2d_s2d_kernel(const float* __restrict__ x, const float* __restrict__ ln_w,
                                                       const float* __restrict__ ln_b, op_t* __restrict__ out,
                                                       int n_px, int H, int W, int C, float eps) {
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= n_px) return;
  const int xw = p % W, yh = (p / W) % H, b = p / (W * H);
  const float* row = x + static_cast<size_t>(p) * C;
  op_t* o = out + ((static_cast<size_t>(b) * (H / 2) + yh / 2) * (W / 2) + xw / 2) * 4 * C +
                     ((yh & 1) * 2 + (xw & 1)) * C;
  float s = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    s += v.x + v.y + v.z + v.w;
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    const float a = v.x - mean, bb = v.y - mean, cc = v.z - mean, d = v.w - mean;
    q += a * a + bb * bb + cc * cc + d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    const float4 g = *reinterpret_cast<const float4*>(ln_w + c);
    const float4 be = *reinterpret_cast<const float4*>(ln_b + c);
    op2_t lo = ff2op2((v.x - mean) * rstd * g.x + be.x, (v.y - mean) * rstd * g.y + be.y);
    op2_t hi = ff2op2((v.z - mean) * rstd * g.z + be.z, (v.w - mean) * rstd * g.w + be.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(o + c) = pk;
  }
}

// Same for C = NV*128 with the pixel's channels held in registers: one global read instead of three dependent passes
// (the kernel above measured 29 % of its HBM roofline at the stage-0 shape).
template <int NV>
__global__ void __launch_bounds__(256) ln2d_s2d_reg_kernel(const float* __restrict__ x, const float* __restrict__ ln_w,
                                                           const float* __restrict__ ln_b,
                                                           op_t* __restrict__ out,
                                                           op_t* __restrict__ copy, int n_px, int H, int W,
                                                           float eps) {
  constexpr int C = NV * 128;
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= n_px) return;
  const int xw = p % W, yh = (p / W) % H, b = p / (W * H);
  const float4* row = reinterpret_cast<const float4*>(x + static_cast<size_t>(p) * C);
  op_t* o = out + ((static_cast<size_t>(b) * (H / 2) + yh / 2) * (W / 2) + xw / 2) * 4 * C +
                     ((yh & 1) * 2 + (xw & 1)) * C;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = __ldg(row + i * 32 + lane);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  if (copy != nullptr) {      // the un-normalised stage output as bf16: the decoder's skip operand, no extra read
    uint2* cp = reinterpret_cast<uint2*>(copy + static_cast<size_t>(p) * C);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      op2_t lo = ff2op2(v[i].x, v[i].y), hi = ff2op2(v[i].z, v[i].w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      cp[i * 32 + lane] = pk;
    }
  }
  const float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, bb = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + bb * bb) + (cc * cc + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = i * 32 + lane;
    const float4 g = __ldg(reinterpret_cast<const float4*>(ln_w) + idx);
    const float4 be = __ldg(reinterpret_cast<const float4*>(ln_b) + idx);
    op2_t lo = ff2op2((v[i].x - mean) * rstd * g.x + be.x, (v[i].y - mean) * rstd * g.y + be.y);
    op2_t hi = ff2op2((v[i].z - mean) * rstd * g.z + be.z, (v[i].w - mean) * rstd * g.w + be.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(o)[idx] = pk;
  }
}

// ------------------------------------------------------------------------------------ GRN
// scale[b][k] = 1 + gamma[k] * Gx / (mean_k Gx + eps), Gx = sqrt(sum_t partial[b*tps + t][k]) where the
// partials are the fc1 epilogue's per-row-tile sums of squares.  Two small grid-parallel kernels, every
// sum in a fixed order (the result does not depend on how tiles are batched, no atomics).  timm
// GlobalResponseNorm; the beta term is folded into fc2's bias on the host.
//   pass 1: grid (K/64, B), 256 threads = 64 channels x 4 row-tile groups -> Gx into scale[], block sum into psum
//   pass 2: grid (ceil(K/256), B): mean from the K/64 block sums, scale in place
__global__ void __launch_bounds__(256) grn_gx_kernel(const float* __restrict__ partial, int tps,
                                                     float* __restrict__ gx_out, float* __restrict__ psum, int K) {
  __shared__ float sm[4][64];
  const int b = blockIdx.y, kk = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int k = blockIdx.x * 64 + kk;
  const float* base = partial + static_cast<size_t>(b) * tps * K + k;
  float acc = 0.f;
  for (int t = g; t < tps; t += 4) acc += base[static_cast<size_t>(t) * K];
  sm[g][kk] = acc;
  __syncthreads();
  if (g == 0) {
    const float gx = sqrtf((sm[0][kk] + sm[1][kk]) + (sm[2][kk] + sm[3][kk]));
    gx_out[static_cast<size_t>(b) * K + k] = gx;
    float s = warp_sum(gx);                       // shuffle tree: fixed order
    if ((kk & 31) == 0) sm[0][kk >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) psum[b * gridDim.x + blockIdx.x] = sm[0][0] + sm[0][1];
}

__global__ void __launch_bounds__(256) grn_apply_kernel(float* __restrict__ scale, const float* __restrict__ psum,
                                                        int nblk, const float* __restrict__ gamma, int K, float eps) {
  const int b = blockIdx.y;
  float tot = 0.f;
  for (int i = 0; i < nblk; ++i) tot += psum[b * nblk + i];     // <= 64 values, same order in every thread
  const float inv = 1.0f / (tot / K + eps);
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k < K) scale[static_cast<size_t>(b) * K + k] = 1.0f + gamma[k] * scale[static_cast<size_t>(b) * K + k] * inv;
}

// out[b][n][k] = bf16(w[n][k] * scale[b][k])   (8 elements per thread)
__global__ void __launch_bounds__(256) scale_weights_kernel(const op_t* __restrict__ w,
                                                            const float* __restrict__ scale,
                                                            op_t* __restrict__ out, int N, int K) {
  const int b = blockIdx.y;
  const size_t i8 = (static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x) * 8;
  if (i8 >= static_cast<size_t>(N) * K) return;
  const int k = static_cast<int>(i8 % K);
  const uint4 raw = *reinterpret_cast<const uint4*>(w + i8);
  const op2_t* wp = reinterpret_cast<const op2_t*>(&raw);
  const float4 s0 = *reinterpret_cast<const float4*>(scale + static_cast<size_t>(b) * K + k);
  const float4 s1 = *reinterpret_cast<const float4*>(scale + static_cast<size_t>(b) * K + k + 4);
  uint4 o;
  op2_t t;
  float2 f;
  f = op22ff(wp[0]); t = ff2op2(f.x * s0.x, f.y * s0.y); o.x = *reinterpret_cast<uint32_t*>(&t);
  f = op22ff(wp[1]); t = ff2op2(f.x * s0.z, f.y * s0.w); o.y = *reinterpret_cast<uint32_t*>(&t);
  f = op22ff(wp[2]); t = ff2op2(f.x * s1.x, f.y * s1.y); o.z = *reinterpret_cast<uint32_t*>(&t);
  f = op22ff(wp[3]); t = ff2op2(f.x * s1.z, f.y * s1.w); o.w = *reinterpret_cast<uint32_t*>(&t);
  *reinterpret_cast<uint4*>(out + static_cast<size_t>(b) * N * K + i8) = o;
}

// Round 2 tried twice to shorten the GRN side path (grn_gx + grn_apply + scale_weights = 7.6 % of the step) and measured
// both variants as no faster, so the three small kernels stay:
//  (1) ONE launch that also recomputed Gx per CTA from the fc1 partials: bit-identical, but 12.6 % of the step -- the
//      per-sample statistics are a serial chain of dependent loads, and a few large CTAs had too little memory parallelism
//      for the 77 MB of scaled weights;
//  (2) grn_apply folded into scale_weights (scale = 1 + gamma * Gx * inv computed per element): bit-identical, 8.1 % -- the
//      kernel is bound by its 77 MB of writes (3.3 TB/s), and the extra Gx / gamma reads cost what the saved launch gave.

// h[m][k] = bf16(h[m][k] * scale[m / rows_per_sample][k]) in place
__global__ void __launch_bounds__(256) scale_rows_kernel(op_t* __restrict__ h, const float* __restrict__ scale,
                                                         size_t total, int K, int rows_per_sample) {
  const size_t i8 = (static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x) * 8;
  if (i8 >= total) return;
  const size_t m = i8 / K;
  const int k = static_cast<int>(i8 % K);
  const size_t b = m / rows_per_sample;
  uint4 raw = *reinterpret_cast<const uint4*>(h + i8);
  op2_t* hp = reinterpret_cast<op2_t*>(&raw);
  const float4 s0 = *reinterpret_cast<const float4*>(scale + b * K + k);
  const float4 s1 = *reinterpret_cast<const float4*>(scale + b * K + k + 4);
  float2 f;
  f = op22ff(hp[0]); hp[0] = ff2op2(f.x * s0.x, f.y * s0.y);
  f = op22ff(hp[1]); hp[1] = ff2op2(f.x * s0.z, f.y * s0.w);
  f = op22ff(hp[2]); hp[2] = ff2op2(f.x * s1.x, f.y * s1.y);
  f = op22ff(hp[3]); hp[3] = ff2op2(f.x * s1.z, f.y * s1.w);
  *reinterpret_cast<uint4*>(h + i8) = raw;
}

// ------------------------------------------------------------------------------------ up x2 + concat
// out[b][y][x][0:C1] = a[b][y/2][x/2][:] (nearest, smp DecoderBlock) ; out[..][C1:C1+C2] = s[b][y][x][:]
// a / s are 16-bit (copied bit for bit: they must already be in the output's format) or fp32 (encoder features, converted
// to the output's 16-bit format: fp16 saturating, or bf16 for the training step).  8 channels per thread.
struct Raw16 {};   // "some 16-bit format", moved without interpretation
template <typename T, bool F16>
__device__ __forceinline__ uint4 load8_16(const void* p, size_t idx);
template <>
__device__ __forceinline__ uint4 load8_16<Raw16, true>(const void* p, size_t idx) {
  return *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p) + idx);
}
template <>
__device__ __forceinline__ uint4 load8_16<Raw16, false>(const void* p, size_t idx) {
  return *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p) + idx);
}
template <bool F16>
__device__ __forceinline__ uint4 cvt8_16(const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  return make_uint4(pack16<F16>(a.x, a.y), pack16<F16>(a.z, a.w), pack16<F16>(b.x, b.y), pack16<F16>(b.z, b.w));
}
template <>
__device__ __forceinline__ uint4 load8_16<float, true>(const void* p, size_t idx) {
  return cvt8_16<true>(reinterpret_cast<const float*>(p) + idx);
}
template <>
__device__ __forceinline__ uint4 load8_16<float, false>(const void* p, size_t idx) {
  return cvt8_16<false>(reinterpret_cast<const float*>(p) + idx);
}

template <typename TA, typename TS, bool F16>
__global__ void __launch_bounds__(256) upcat_kernel(const void* __restrict__ a, const void* __restrict__ s,
                                                    uint16_t* __restrict__ out, size_t n_vec, int H, int W,
                                                    int C1, int C2) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n_vec) return;
  const int CT = C1 + C2;
  const int vec_per_px = CT / 8;
  const size_t px = i / vec_per_px;
  const int c = static_cast<int>(i % vec_per_px) * 8;
  const int xw = static_cast<int>(px % W), yh = static_cast<int>((px / W) % H);
  const size_t b = px / (static_cast<size_t>(W) * H);
  uint4 v;
  if (c < C1) {
    v = load8_16<TA, F16>(a, ((b * (H / 2) + yh / 2) * (W / 2) + xw / 2) * C1 + c);
  } else {
    v = load8_16<TS, F16>(s, px * C2 + (c - C1));
  }
  *reinterpret_cast<uint4*>(out + px * CT + c) = v;
}

template <bool F16>
static void launch_upcat(bool a16, bool s16, const void* a, const void* s, uint16_t* o, size_t n_vec, int H, int W, int C1,
                         int C2, unsigned grid, cudaStream_t st) {
  if (a16 && s16) upcat_kernel<Raw16, Raw16, F16><<<grid, 256, 0, st>>>(a, s, o, n_vec, H, W, C1, C2);
  else if (a16) upcat_kernel<Raw16, float, F16><<<grid, 256, 0, st>>>(a, s, o, n_vec, H, W, C1, C2);
  else if (s16) upcat_kernel<float, Raw16, F16><<<grid, 256, 0, st>>>(a, s, o, n_vec, H, W, C1, C2);
  else upcat_kernel<float, float, F16><<<grid, 256, 0, st>>>(a, s, o, n_vec, H, W, C1, C2);
}

}  // namespace fz

// ============================================================================ C ABI
namespace fz {
template <bool F32IN>
static int launch_stem(const void* in, int Cin, const float* w, const float* bias, const float* ln_w,
                       const float* ln_b, float* out, int B, int P, int C0, float eps, cudaStream_t st) {
  FZ_REQUIRE(P % 128 == 0, "fz_stem_ln: P=%d must be a multiple of 128", P);
  FZ_REQUIRE(C0 % 32 == 0 && C0 >= 32 && C0 <= 384, "fz_stem_ln: C0=%d unsupported", C0);
  FZ_REQUIRE(Cin >= 1 && Cin <= 4, "fz_stem_ln: Cin=%d must be 1..4", Cin);
  if (B <= 0) return 0;
  // uint8 tiles, the common widths, P a multiple of 256: the tensor-core kernel (FZ_STEM_MMA=0 keeps the fp32 FMA kernel)
  static int use_mma = -1;
  if (use_mma < 0) {
    const char* e = getenv("FZ_STEM_MMA");
    use_mma = (e && e[0] == '0') ? 0 : 1;
  }
  if (!F32IN && use_mma && Cin == 4 && P % 256 == 0 && (C0 == 96 || C0 == 128 || C0 == 192)) {
    const size_t sm = (2 * 64 * static_cast<size_t>(C0 + 8) + STEM_MMA_PX * STEM_MMA_LDI + 3 * C0) * sizeof(float);
    const int n_rows = B * (P / 4);
    const int g2 = n_rows < 2 * 148 ? n_rows : 2 * 148;          // two 89 KB blocks per SM
    auto in4 = reinterpret_cast<const uchar4*>(in);
#define FZ_STEM_MMA(NBV)                                                                                 \
  case NBV * 8: {                                                                                        \
    FZ_ENSURE_SMEM((stem_ln_mma_kernel<NBV>), static_cast<int>(sm));                                     \
    stem_ln_mma_kernel<NBV><<<g2, 128, sm, st>>>(in4, w, bias, ln_w, ln_b, out, P, n_rows, eps);         \
    break;                                                                                               \
  }
    switch (C0) { FZ_STEM_MMA(12) FZ_STEM_MMA(16) FZ_STEM_MMA(24) }
#undef FZ_STEM_MMA
    FZ_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const size_t smem = (64 * static_cast<size_t>(C0) + STEM_PX * STEM_IN_STRIDE) * sizeof(float);
  dim3 grid(P / 4, B);
#define FZ_STEM(CPL)                                                                                           \
  case CPL: {                                                                                                  \
    FZ_ENSURE_SMEM((stem_ln_kernel<CPL, F32IN>), static_cast<int>(smem));                                      \
    stem_ln_kernel<CPL, F32IN><<<grid, 128, smem, st>>>(in, Cin, w, bias, ln_w, ln_b, out, P, eps);            \
    break;                                                                                                     \
  }
  switch (C0 / 32) {
    FZ_STEM(1) FZ_STEM(2) FZ_STEM(3) FZ_STEM(4) FZ_STEM(6) FZ_STEM(8) FZ_STEM(11) FZ_STEM(12)
    default: set_error("fz_stem_ln: C0=%d not instantiated", C0); return -1;
  }
#undef FZ_STEM
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace fz

extern "C" int fz_stem_ln(const uint8_t* tiles_u8, const float* w, const float* bias, const float* ln_w,
                          const float* ln_b, float* out, int B, int P, int C0, float eps, void* stream) {
  return fz::launch_stem<false>(tiles_u8, 4, w, bias, ln_w, ln_b, out, B, P, C0, eps,
                                reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int fz_stem_ln_f32(const float* x_nchw, int Cin, const float* w, const float* bias, const float* ln_w,
                              const float* ln_b, float* out, int B, int P, int C0, float eps, void* stream) {
  return fz::launch_stem<true>(x_nchw, Cin, w, bias, ln_w, ln_b, out, B, P, C0, eps,
                               reinterpret_cast<cudaStream_t>(stream));
}

namespace fz {
template <int NW, int CPT, int C>
static int launch_dwconv(const float* x, const float* wdw, const float* bdw, const float* ln_w, const float* ln_b,
                         op_t* out, int B, int H, int W, float eps, cudaStream_t st) {
  constexpr int SW = 8 / CPT;
  FZ_REQUIRE(H % 4 == 0 && W % SW == 0, "fz_dwconv7_ln: H=%d W=%d must be multiples of 4 x %d", H, W, SW);
  const int tiles_x = W / SW, tiles_y = H / 4, n_tiles = tiles_x * tiles_y * B;
  // persistent walk for the 16-warp one-channel-per-lane shape: one CTA per SM, x-fastest tile order (neighbouring
  // tiles run concurrently on different SMs and share their halos in L2)
  static int persist = -1;
  if (persist < 0) {
    const char* e = getenv("FZ_DWCONV_PERSIST");
    persist = (e && e[0] == '0') ? 0 : 1;
  }
  if constexpr (NW == 16 && CPT == 1) {
   if (persist) {
    const int sm_count = device_sm_count();
    if (sm_count <= 0) return -2;
    const int grid = n_tiles < sm_count ? n_tiles : sm_count;
    dwconv_ln_kernel<NW, 1, C, true><<<grid, NW * 32, 0, st>>>(x, wdw, bdw, ln_w, ln_b, out, H, W, eps, tiles_x, tiles_y, n_tiles);
    FZ_CHECK_CUDA(cudaGetLastError());
    return 0;
   }
  }
  dwconv_ln_kernel<NW, CPT, C, false><<<n_tiles, NW * 32, 0, st>>>(x, wdw, bdw, ln_w, ln_b, out, H, W, eps, tiles_x, tiles_y,
                                                                   n_tiles);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace fz

namespace fz {
template <int NW, int C, int SH>
static int launch_dwconv_pair(const float* x, const float* wdw, const float* bdw, const float* ln_w, const float* ln_b,
                              op_t* out, int B, int H, int W, float eps, cudaStream_t st) {
  FZ_REQUIRE(H % SH == 0 && W % 8 == 0, "fz_dwconv7_ln: H=%d W=%d must be multiples of %d x 8", H, W, SH);
  const int tiles_x = W / 8, tiles_y = H / SH, n_tiles = tiles_x * tiles_y * B;
  constexpr int SMEM = 49 * C * 4;
  auto kern = dwconv_ln_pair_kernel<NW, C, SH>;
  FZ_ENSURE_SMEM(kern, SMEM);
  const int sm_count = device_sm_count();
  if (sm_count <= 0) return -2;
  // persistent CTAs: as many as fit per SM by shared memory (taps) and threads, x-fastest tile order
  int per_sm = (220 * 1024) / (SMEM + 2048);
  if (per_sm > 2048 / (NW * 32)) per_sm = 2048 / (NW * 32);
  const int thread_cap = SH == 2 ? 512 : 256;                       // ~126 (2 x 8 tile) / ~170 (4 x 8) registers per thread
  if (per_sm > thread_cap / (NW * 32)) per_sm = thread_cap / (NW * 32);
  if (per_sm < 1) per_sm = 1;
  const long long want = static_cast<long long>(sm_count) * per_sm;
  const int grid = static_cast<int>(n_tiles < want ? n_tiles : want);
  kern<<<grid, NW * 32, SMEM, st>>>(x, wdw, bdw, ln_w, ln_b, out, H, W, eps, tiles_x, tiles_y, n_tiles);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace fz

extern "C" int fz_dwconv7_ln(const float* x, const float* wdw, const float* bdw, const float* ln_w, const float* ln_b,
                             void* out_bf16, int B, int H, int W, int C, float eps, void* stream) {
  using namespace fz;
  if (B <= 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  op_t* o = reinterpret_cast<op_t*>(out_bf16);
  // Kernel choice, measured at B = 37 (profiles/r2_dwconv_pair_ab.txt): the one-channel-per-lane kernel, the channel-pair
  // kernel with 2 x 8 tiles (40 % fewer instructions, 16 warps / SM) and with 4 x 8 tiles (half the L2 reads of the former,
  // 8 warps / SM) all take the SAME time at C = 128 and 512 (264 / 267 / 274 us and 78.2 / 78.6 / 79.3 us): the kernel is
  // bound by neither issue slots nor L2 traffic nor occupancy but by the fp32 FMA work itself under the power cap.  The pair
  // kernels win where the scalar one had the wrong shape: C = 256 (4 x 8 tiles: 142 vs 149 us) and C = 1024 (2 x 8: 33.5 vs
  // 43.7 us).  FZ_DWCONV_PAIR = 0 / 2 / 4 forces the scalar / 2 x 8 / 4 x 8 kernel everywhere (A/B measurements).
  static int pair = -1;
  if (pair < 0) {
    const char* e = getenv("FZ_DWCONV_PAIR");
    pair = e ? atoi(e) : 1;
  }
  const bool ok2 = H % 2 == 0 && W % 8 == 0, ok4 = H % 4 == 0 && W % 8 == 0;
  if (ok4 && (pair == 4 || (pair == 1 && C == 256))) {
    switch (C) {
      case 128: return launch_dwconv_pair<2, 128, 4>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
      case 256: return launch_dwconv_pair<4, 256, 4>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
      case 512: return launch_dwconv_pair<8, 512, 4>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
      default: break;
    }
  }
  if (ok2 && (pair == 2 || pair == 4 || (pair == 1 && C == 1024))) {
    switch (C) {
      case 128: return launch_dwconv_pair<2, 128, 2>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
      case 256: return launch_dwconv_pair<4, 256, 2>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
      case 512: return launch_dwconv_pair<8, 512, 2>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
      case 1024: return launch_dwconv_pair<16, 1024, 2>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
      default: break;
    }
  }
  switch (C) {   // one warp per 32 channels; above 512 channels each thread carries two
    case 64: return launch_dwconv<2, 1, 64>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
    case 96: return launch_dwconv<3, 1, 96>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
    case 128: return launch_dwconv<4, 1, 128>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
    case 192: return launch_dwconv<6, 1, 192>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
    case 256: return launch_dwconv<8, 1, 256>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
    case 384: return launch_dwconv<12, 1, 384>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
    case 512: return launch_dwconv<16, 1, 512>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
    case 768: return launch_dwconv<12, 2, 768>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
    case 1024: return launch_dwconv<16, 2, 1024>(x, wdw, bdw, ln_w, ln_b, o, B, H, W, eps, st);
  }
  set_error("fz_dwconv7_ln: C=%d has no instantiation (64,96,128,192,256,384,512,768,1024)", C);
  return -1;
}

extern "C" int fz_ln2d_s2d(const float* x, const float* ln_w, const float* ln_b, void* out_bf16, int B, int H, int W,
                           int C, float eps, void* stream) {
  return fz_ln2d_s2d_copy(x, ln_w, ln_b, out_bf16, nullptr, B, H, W, C, eps, stream);
}

extern "C" int fz_ln2d_s2d_copy(const float* x, const float* ln_w, const float* ln_b, void* out_bf16, void* copy_bf16,
                                int B, int H, int W, int C, float eps, void* stream) {
  using namespace fz;
  FZ_REQUIRE(copy_bf16 == nullptr || C == 128 || C == 256 || C == 512,
             "fz_ln2d_s2d_copy: the bf16 copy is built for C = 128, 256, 512 (got %d)", C);
  op_t* cpy = reinterpret_cast<op_t*>(copy_bf16);
  FZ_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "fz_ln2d_s2d: bad shape H=%d W=%d C=%d", H, W, C);
  const int n_px = B * H * W;
  if (n_px <= 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  op_t* o = reinterpret_cast<op_t*>(out_bf16);
  const unsigned grid = (n_px + 7) / 8;
  switch (C) {
    case 128: ln2d_s2d_reg_kernel<1><<<grid, 256, 0, st>>>(x, ln_w, ln_b, o, cpy, n_px, H, W, eps); break;
    case 256: ln2d_s2d_reg_kernel<2><<<grid, 256, 0, st>>>(x, ln_w, ln_b, o, cpy, n_px, H, W, eps); break;
    case 512: ln2d_s2d_reg_kernel<4><<<grid, 256, 0, st>>>(x, ln_w, ln_b, o, cpy, n_px, H, W, eps); break;
    default: ln2d_s2d_kernel<<<grid, 256, 0, st>>>(x, ln_w, ln_b, o, n_px, H, W, C, eps);
  }
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_grn_scale(const float* sumsq_partial, int tiles_per_sample, const float* gamma, float* scale,
                            float* scratch, int B, int K, float eps, void* stream) {
  using namespace fz;
  FZ_REQUIRE(tiles_per_sample >= 1 && K >= 64 && K % 64 == 0, "fz_grn_scale: bad arguments tps=%d K=%d (K %% 64)",
             tiles_per_sample, K);
  FZ_REQUIRE(scratch != nullptr, "fz_grn_scale: scratch (B*K/64 floats) required");
  if (B <= 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  grn_gx_kernel<<<dim3(K / 64, B), 256, 0, st>>>(sumsq_partial, tiles_per_sample, scale, scratch, K);
  grn_apply_kernel<<<dim3((K + 255) / 256, B), 256, 0, st>>>(scale, scratch, K / 64, gamma, K, eps);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_scale_weights(const void* w_bf16, const float* scale, void* out_bf16, int B, int N, int K,
                                void* stream) {
  using namespace fz;
  FZ_REQUIRE(K % 8 == 0, "fz_scale_weights: K=%d must be a multiple of 8", K);
  if (B <= 0) return 0;
  const size_t n8 = static_cast<size_t>(N) * K / 8;
  dim3 grid(static_cast<unsigned>((n8 + 255) / 256), B);
  scale_weights_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const op_t*>(w_bf16), scale, reinterpret_cast<op_t*>(out_bf16), N, K);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_scale_rows(void* h_bf16, const float* scale, int64_t M, int K, int rows_per_sample, void* stream) {
  using namespace fz;
  FZ_REQUIRE(K % 8 == 0 && rows_per_sample > 0, "fz_scale_rows: bad arguments");
  const size_t total = static_cast<size_t>(M) * K;
  if (total == 0) return 0;
  scale_rows_kernel<<<static_cast<unsigned>((total / 8 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<op_t*>(h_bf16), scale, total, K, rows_per_sample);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_upsample2_concat(const void* a, int a_dtype, const void* s, int s_dtype, void* out16, int out_dtype,
                                   int B, int H, int W, int C1, int C2, void* stream) {
  using namespace fz;
  FZ_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0 && C1 > 0 && C2 >= 0, "fz_upsample2_concat: C1=%d C2=%d must be multiples of 8",
             C1, C2);
  FZ_REQUIRE(H % 2 == 0 && W % 2 == 0, "fz_upsample2_concat: output H, W must be even");
  FZ_REQUIRE(out_dtype == FZ_BF16 || out_dtype == FZ_F16, "fz_upsample2_concat: the output is FZ_BF16 or FZ_F16");
  FZ_REQUIRE((a_dtype == FZ_F32 || a_dtype == out_dtype) && (C2 == 0 || s_dtype == FZ_F32 || s_dtype == out_dtype),
             "fz_upsample2_concat: a 16-bit input must have the output's format (a %d, s %d, out %d)", a_dtype, s_dtype,
             out_dtype);
  const size_t n_vec = static_cast<size_t>(B) * H * W * (C1 + C2) / 8;
  if (n_vec == 0) return 0;
  const unsigned grid = static_cast<unsigned>((n_vec + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint16_t* o = reinterpret_cast<uint16_t*>(out16);
  if (out_dtype == FZ_F16) launch_upcat<true>(a_dtype != FZ_F32, s_dtype != FZ_F32, a, s, o, n_vec, H, W, C1, C2, grid, st);
  else launch_upcat<false>(a_dtype != FZ_F32, s_dtype != FZ_F32, a, s, o, n_vec, H, W, C1, C2, grid, st);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

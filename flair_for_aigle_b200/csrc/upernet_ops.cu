// CUDA-core kernels of the smp 0.4.0 UPerNet decoder + SegmentationHead(kernel 1, upsampling 4) -- decoder of
// `swin_base_patch4_window12_384-upernet` (BASELINE.json configs[2]; reference call site
// flair_hub/models/flair_model.py:417-419 through monotemp_model.py:7-31):
//   fz_adaptive_avgpool   nn.AdaptiveAvgPool2d(s) of the PSP module
//   fz_bilinear_slice     F.interpolate(mode='bilinear', align_corners=False) of a bf16 NHWC map, written into a
//                         channel slice of a wider NHWC tensor (the torch.cat operands of PSP and of the FPN fuse),
//                         optionally + a lateral map (FPNBlock: upsample(x) + skip_conv(skip))
//   fz_updown_slice       down2(up2(x)): what the 0-channel FPN stage followed by the resize to H/4 computes,
//                         = separable [1/8, 3/4, 1/8] filter with replicated borders; the 2H x 2W map is never stored
//   fz_head_upsample4     nn.UpsamplingBilinear2d(scale_factor=4) (align_corners=True) of the class logits -> NCHW
// The 1x1 convolutions run as GEMMs and the 3x3 fuse convolution as an implicit GEMM on tcgen05.
#include "common.h"
#include "ptx.cuh"
#include "../../include/flair_zonal_b200.h"

#include "ptx.cuh"
#include "operand.cuh"

namespace fz {

__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
  const op2_t* h = reinterpret_cast<const op2_t*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = op22ff(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_op(f[0], f[1]), pack_op(f[2], f[3]), pack_op(f[4], f[5]), pack_op(f[6], f[7]));
}

// torch adaptive pooling bins: [floor(i*H/S), ceil((i+1)*H/S))
__global__ void __launch_bounds__(128) adaptive_avgpool_kernel(const op_t* __restrict__ in,
                                                               op_t* __restrict__ out, int H, int W, int C,
                                                               int S) {
  const int cell = blockIdx.x % (S * S);
  const int b = blockIdx.x / (S * S);
  const int oy = cell / S, ox = cell % S;
  const int y0 = (oy * H) / S, y1 = ((oy + 1) * H + S - 1) / S;
  const int x0 = (ox * W) / S, x1 = ((ox + 1) * W + S - 1) / S;
  const float inv = 1.0f / static_cast<float>((y1 - y0) * (x1 - x0));
  for (int c = threadIdx.x * 8; c < C; c += 128 * 8) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(in + ((static_cast<size_t>(b) * H + y) * W + x) * C + c), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    *reinterpret_cast<uint4*>(out + (static_cast<size_t>(blockIdx.x)) * C + c) = pack8(acc);
  }
}

// PyTorch's bilinear source index (align_corners=False): src = max((dst + 0.5) * scale - 0.5, 0), scale = in/out
__device__ __forceinline__ void bilin_coord(int d, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = (static_cast<float>(d) + 0.5f) * scale - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = static_cast<int>(s);
  i0 = i0 > in_size - 1 ? in_size - 1 : i0;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - static_cast<float>(i0);
}

constexpr int PYR_ROWS = 8;

// x-interpolated source row: hx * f[yy][x0] + lx * f[yy][x1]
__device__ __forceinline__ void pyr_xlerp(const op_t* src, int yy, int w, int C, int x0, int x1, float hx, float lx,
                                          float (&dst)[8]) {
  float f0[8], f1[8];
  unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(yy) * w + x0) * C), f0);
  unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(yy) * w + x1) * C), f1);
#pragma unroll
  for (int j = 0; j < 8; ++j) dst[j] = hx * f0[j] + lx * f1[j];
}
// horizontally filtered source row of the down2(up2(.)) stencil: 0.125 f[x-1] + 0.75 f[x] + 0.125 f[x+1] (clamped)
__device__ __forceinline__ void pyr_xtap3(const op_t* src, int yy, int w, int C, int xm, int x, int xp,
                                          float (&dst)[8]) {
  float f0[8], f1[8], f2[8];
  unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(yy) * w + xm) * C), f0);
  unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(yy) * w + x) * C), f1);
  unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(yy) * w + xp) * C), f2);
#pragma unroll
  for (int j = 0; j < 8; ++j) dst[j] = 0.125f * f0[j] + 0.75f * f1[j] + 0.125f * f2[j];
}

// MODE 0: bilinear resize (h,w)->(H,W) (+ add); MODE 1: down2(up2(.)) at the same size (h == H, w == W)
// One CTA owns a (256 / vectors-per-pixel) x PYR_ROWS output patch of one sample and walks its rows: the row and sample are
// CTA-uniform, index math is 32-bit, and both resamplings run separably with the x-filtered source rows kept in
// registers across the output rows that share them (the first version decoded a flat 64-bit index per 16-byte vector
// and re-fetched 4 / 9 source vectors per output: 218 us for a 310 MB slice).
template <int MODE>
__global__ void __launch_bounds__(256) resize_slice_kernel(const op_t* __restrict__ in,
                                                           const op_t* __restrict__ add,
                                                           op_t* __restrict__ out, int h, int w, int H, int W,
                                                           int C, int Ctot, int c0) {
  const int vpp = C / 8;
  const int vper = vpp < 256 ? vpp : 256;                   // vectors of one pixel handled side by side
  const int xi = threadIdx.x / vper, v = threadIdx.x - xi * vper;
  const int x = blockIdx.x * (256 / vper) + xi;
  if (x >= W || xi >= 256 / vper) return;
  const size_t b = blockIdx.z;
  const int ya = blockIdx.y * PYR_ROWS, yb = min(ya + PYR_ROWS, H);
  const size_t orow = static_cast<size_t>(W) * Ctot, arow = static_cast<size_t>(W) * C;
  for (int c = v * 8; c < C; c += 256 * 8) {
    const op_t* src = in + b * h * w * C + c;
    op_t* o = out + ((b * H + ya) * W + x) * static_cast<size_t>(Ctot) + c0 + c;
    const op_t* ad = add ? add + ((b * H + ya) * W + x) * static_cast<size_t>(C) + c : nullptr;
    auto emit = [&](float (&acc)[8]) {
      if (ad != nullptr) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(ad), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
        ad += arow;
      }
      *reinterpret_cast<uint4*>(o) = pack8(acc);
      o += orow;
    };
    if (MODE == 0 && h == H && w == W) {
      for (int y = ya; y < yb; ++y) {
        float acc[8];
        unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(y) * w + x) * C), acc);
        emit(acc);
      }
    } else if (MODE == 0) {
      int x0, x1, cy0 = -1, cy1 = -1;
      float lx;
      bilin_coord(x, static_cast<float>(w) / W, w, x0, x1, lx);
      const float hx = 1.f - lx;
      float top[8], bot[8];
      for (int y = ya; y < yb; ++y) {                        // y0, y1 are CTA-uniform: no divergence
        int y0, y1;
        float ly;
        bilin_coord(y, static_cast<float>(h) / H, h, y0, y1, ly);
        if (y0 != cy0) {
          if (y0 == cy1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) top[j] = bot[j];
          } else {
            pyr_xlerp(src, y0, w, C, x0, x1, hx, lx, top);
          }
          cy0 = y0;
        }
        if (y1 != cy1) {
          if (y1 == y0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bot[j] = top[j];
          } else {
            pyr_xlerp(src, y1, w, C, x0, x1, hx, lx, bot);
          }
          cy1 = y1;
        }
        const float hy = 1.f - ly;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = hy * top[j] + ly * bot[j];
        emit(acc);
      }
    } else {
      const int xm = x > 0 ? x - 1 : 0, xp = x < W - 1 ? x + 1 : W - 1;
      float rm[8], r0[8], rp[8];
      pyr_xtap3(src, ya > 0 ? ya - 1 : 0, W, C, xm, x, xp, rm);
      pyr_xtap3(src, ya, W, C, xm, x, xp, r0);
      for (int y = ya; y < yb; ++y) {
        pyr_xtap3(src, y < H - 1 ? y + 1 : H - 1, W, C, xm, x, xp, rp);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] = 0.125f * rm[j] + 0.75f * r0[j] + 0.125f * rp[j];
          rm[j] = r0[j];
          r0[j] = rp[j];
        }
        emit(acc);
      }
    }
  }
}

// The UPerNet fuse input in one launch: out [B][H][H][5C] = [bilinear(p0) | bilinear(p1) | bilinear(p2) | p3 | down2(up2(p3))]
// (smp UPerNetDecoder: the four FPN maps resized to the stride-4 grid plus the unused fifth block's map).  The slice-by-
// slice version ran at 1.4-1.7 TB/s of output because every output vector pulled its 4 (or 9) source vectors through
// L2 -> SM again (~6 TB/s of gather traffic).  Here one CTA owns ONE slice of a PX x 4 pixel patch and walks its 4 rows
// in turn: source vectors shared by neighbouring output pixels are L1 hits, and both resamplings run separably (x pass
// per source row, kept in registers across the output rows that share it), which halves the arithmetic per output.
__global__ void __launch_bounds__(256) pyramid_concat_kernel(const op_t* __restrict__ p0, int s0,
                                                             const op_t* __restrict__ p1, int s1,
                                                             const op_t* __restrict__ p2, int s2,
                                                             const op_t* __restrict__ p3,
                                                             op_t* __restrict__ out, int H, int C) {
  const int vpp = C / 8;                                   // 16-byte vectors per pixel and slice; 256 % vpp == 0
  const int xi = threadIdx.x / vpp, c = (threadIdx.x - xi * vpp) * 8;
  const int x = blockIdx.x * (256 / vpp) + xi;
  const int k = blockIdx.z % 5;                            // CTA-uniform slice
  const size_t b = blockIdx.z / 5;
  if (x >= H) return;
  const op_t* in = k == 0 ? p0 : (k == 1 ? p1 : (k == 2 ? p2 : p3));
  const int h = k == 0 ? s0 : (k == 1 ? s1 : (k == 2 ? s2 : H));
  const op_t* src = in + b * h * h * C + c;
  const int ya = blockIdx.y * PYR_ROWS, yb = min(ya + PYR_ROWS, H);
  op_t* o = out + ((b * H + ya) * H + x) * (5 * static_cast<size_t>(C)) + k * C + c;
  const size_t orow = static_cast<size_t>(H) * 5 * C;
  if (k < 4 && h == H) {                                   // same size: copy
    for (int y = ya; y < yb; ++y, o += orow)
      *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(src + (static_cast<size_t>(y) * h + x) * C);
  } else if (k < 4) {
    // bilinear, separable: the x-interpolated source rows y0 / y1 are kept while consecutive output rows share them
    int x0, x1, cy0 = -1, cy1 = -1;
    float lx;
    bilin_coord(x, static_cast<float>(h) / H, h, x0, x1, lx);
    const float hx = 1.f - lx;
    float top[8], bot[8];
    for (int y = ya; y < yb; ++y, o += orow) {             // y0, y1 are CTA-uniform: no divergence
      int y0, y1;
      float ly;
      bilin_coord(y, static_cast<float>(h) / H, h, y0, y1, ly);
      if (y0 != cy0) {
        if (y0 == cy1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) top[j] = bot[j];
        } else {
          pyr_xlerp(src, y0, h, C, x0, x1, hx, lx, top);
        }
        cy0 = y0;
      }
      if (y1 != cy1) {
        if (y1 == y0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) bot[j] = top[j];
        } else {
          pyr_xlerp(src, y1, h, C, x0, x1, hx, lx, bot);
        }
        cy1 = y1;
      }
      const float hy = 1.f - ly;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = hy * top[j] + ly * bot[j];
      *reinterpret_cast<uint4*>(o) = pack8(acc);
    }
  } else {
    // down2(up2(.)) = the separable 3-tap stencil [1/8 3/4 1/8] with clamped borders: rolling window of filtered rows
    const int xm = x > 0 ? x - 1 : 0, xp = x < H - 1 ? x + 1 : H - 1;
    float rm[8], r0[8], rp[8];
    pyr_xtap3(src, ya > 0 ? ya - 1 : 0, H, C, xm, x, xp, rm);
    pyr_xtap3(src, ya, H, C, xm, x, xp, r0);
    for (int y = ya; y < yb; ++y, o += orow) {
      pyr_xtap3(src, y < H - 1 ? y + 1 : H - 1, H, C, xm, x, xp, rp);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] = 0.125f * rm[j] + 0.75f * r0[j] + 0.125f * rp[j];
        rm[j] = r0[j];
        r0[j] = rp[j];
      }
      *reinterpret_cast<uint4*>(o) = pack8(acc);
    }
  }
}

// logits float [B][h][w][cstride] (first n_cls valid) -> float [B][n_cls][4h][4w], bilinear align_corners=True
__global__ void __launch_bounds__(256) head_upsample4_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                             size_t n_px, int h, int w, int cstride, int n_cls) {
  const size_t px = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (px >= n_px) return;
  const int H = 4 * h, W = 4 * w;
  const int x = static_cast<int>(px % W), y = static_cast<int>((px / W) % H);
  const size_t b = px / (static_cast<size_t>(W) * H);
  const float sy = static_cast<float>(h - 1) / static_cast<float>(H - 1);
  const float sx = static_cast<float>(w - 1) / static_cast<float>(W - 1);
  const float fy = y * sy, fx = x * sx;
  int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  y0 = y0 > h - 1 ? h - 1 : y0;
  x0 = x0 > w - 1 ? w - 1 : x0;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
  const float* p00 = in + ((b * h + y0) * w + x0) * cstride;
  const float* p01 = in + ((b * h + y0) * w + x1) * cstride;
  const float* p10 = in + ((b * h + y1) * w + x0) * cstride;
  const float* p11 = in + ((b * h + y1) * w + x1) * cstride;
  float* o = out + (b * n_cls * H + y) * W + x;
  const size_t plane = static_cast<size_t>(H) * W;
  for (int c = 0; c < n_cls; ++c)
    o[c * plane] = hy * (hx * __ldg(p00 + c) + lx * __ldg(p01 + c)) + ly * (hx * __ldg(p10 + c) + lx * __ldg(p11 + c));
}

}  // namespace fz

extern "C" int fz_adaptive_avgpool(const void* in_bf16, void* out_bf16, int B, int H, int W, int C, int S,
                                   void* stream) {
  using namespace fz;
  FZ_REQUIRE(B >= 0 && H > 0 && W > 0 && S > 0 && S <= H && S <= W && C % 8 == 0, "fz_adaptive_avgpool: bad shape");
  if (B == 0) return 0;
  adaptive_avgpool_kernel<<<B * S * S, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const op_t*>(in_bf16), reinterpret_cast<op_t*>(out_bf16), H, W, C, S);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int resize_common(int mode, const void* in, const void* add, void* out, int B, int h, int w, int H, int W, int C,
                         int Ctot, int c0, void* stream) {
  using namespace fz;
  FZ_REQUIRE(B >= 0 && h > 0 && w > 0 && H > 0 && W > 0, "fz_bilinear_slice: bad shape");
  FZ_REQUIRE(C % 8 == 0 && Ctot % 8 == 0 && c0 % 8 == 0 && c0 + C <= Ctot, "fz_bilinear_slice: bad channel slice");
  if (B == 0) return 0;
  FZ_REQUIRE(B <= 65535, "fz_bilinear_slice: B=%d exceeds the grid limit", B);
  const int vpp = C / 8, vper = vpp < 256 ? vpp : 256, px = 256 / vper;
  const dim3 grid((W + px - 1) / px, (H + PYR_ROWS - 1) / PYR_ROWS, B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const op_t* i = reinterpret_cast<const op_t*>(in);
  const op_t* a = reinterpret_cast<const op_t*>(add);
  op_t* o = reinterpret_cast<op_t*>(out);
  if (mode == 0) resize_slice_kernel<0><<<grid, 256, 0, st>>>(i, a, o, h, w, H, W, C, Ctot, c0);
  else resize_slice_kernel<1><<<grid, 256, 0, st>>>(i, a, o, h, w, H, W, C, Ctot, c0);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_bilinear_slice(const void* in_bf16, const void* add_bf16, void* out_bf16, int B, int h, int w, int H,
                                 int W, int C, int Ctot, int c0, void* stream) {
  return resize_common(0, in_bf16, add_bf16, out_bf16, B, h, w, H, W, C, Ctot, c0, stream);
}

extern "C" int fz_updown_slice(const void* in_bf16, void* out_bf16, int B, int H, int W, int C, int Ctot, int c0,
                               void* stream) {
  return resize_common(1, in_bf16, nullptr, out_bf16, B, H, W, H, W, C, Ctot, c0, stream);
}

extern "C" int fz_pyramid_concat(const void* p0, int s0, const void* p1, int s1, const void* p2, int s2, const void* p3,
                                 void* out_bf16, int B, int H, int C, void* stream) {
  using namespace fz;
  FZ_REQUIRE(B >= 0 && H > 0 && s0 > 0 && s1 > 0 && s2 > 0 && s0 <= H && s1 <= H && s2 <= H, "fz_pyramid_concat: bad sizes");
  FZ_REQUIRE(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0, "fz_pyramid_concat: C=%d: C/8 must divide 256", C);
  FZ_REQUIRE(H <= 65535 * 4 && B * 5 <= 65535, "fz_pyramid_concat: H=%d B=%d exceed the grid limits", H, B);
  if (B == 0) return 0;
  const int px = 256 / (C / 8);
  const dim3 grid((H + px - 1) / px, (H + PYR_ROWS - 1) / PYR_ROWS, B * 5);
  pyramid_concat_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const op_t*>(p0), s0, reinterpret_cast<const op_t*>(p1), s1,
      reinterpret_cast<const op_t*>(p2), s2, reinterpret_cast<const op_t*>(p3),
      reinterpret_cast<op_t*>(out_bf16), H, C);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_head_upsample4(const float* logits, float* out_nchw, int B, int h, int w, int cstride, int n_cls,
                                 void* stream) {
  using namespace fz;
  FZ_REQUIRE(B >= 0 && h > 1 && w > 1 && n_cls >= 1 && n_cls <= cstride, "fz_head_upsample4: bad shape");
  if (B == 0) return 0;
  const size_t n_px = static_cast<size_t>(B) * 16 * h * w;
  head_upsample4_kernel<<<static_cast<unsigned>((n_px + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      logits, out_nchw, n_px, h, w, cstride, n_cls);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

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

#include <cuda_bf16.h>

namespace fz {

__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// torch adaptive pooling bins: [floor(i*H/S), ceil((i+1)*H/S))
__global__ void __launch_bounds__(128) adaptive_avgpool_kernel(const __nv_bfloat16* __restrict__ in,
                                                               __nv_bfloat16* __restrict__ out, int H, int W, int C,
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

// MODE 0: bilinear resize (h,w)->(H,W); MODE 1: down2(up2(.)) at the same size (h == H, w == W)
template <int MODE>
__global__ void __launch_bounds__(256) resize_slice_kernel(const __nv_bfloat16* __restrict__ in,
                                                           const __nv_bfloat16* __restrict__ add,
                                                           __nv_bfloat16* __restrict__ out, size_t n_vec, int h, int w,
                                                           int H, int W, int C, int Ctot, int c0) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n_vec) return;
  const int vpp = C / 8;
  const int c = static_cast<int>(i % vpp) * 8;
  const size_t px = i / vpp;
  const int x = static_cast<int>(px % W), y = static_cast<int>((px / W) % H);
  const size_t b = px / (static_cast<size_t>(W) * H);
  const __nv_bfloat16* src = in + b * h * w * C + c;
  float acc[8];
  if (MODE == 0) {
    if (h == H && w == W) {
      unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(y) * w + x) * C), acc);
    } else {
      int y0, y1, x0, x1;
      float ly, lx;
      bilin_coord(y, static_cast<float>(h) / H, h, y0, y1, ly);
      bilin_coord(x, static_cast<float>(w) / W, w, x0, x1, lx);
      float f00[8], f01[8], f10[8], f11[8];
      unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(y0) * w + x0) * C), f00);
      unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(y0) * w + x1) * C), f01);
      unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(y1) * w + x0) * C), f10);
      unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(y1) * w + x1) * C), f11);
      const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = hy * (hx * f00[j] + lx * f01[j]) + ly * (hx * f10[j] + lx * f11[j]);
    }
  } else {
    const int ym = y > 0 ? y - 1 : 0, yp = y < H - 1 ? y + 1 : H - 1;
    const int xm = x > 0 ? x - 1 : 0, xp = x < W - 1 ? x + 1 : W - 1;
    const int ys[3] = {ym, y, yp}, xs[3] = {xm, x, xp};
    const float k[3] = {0.125f, 0.75f, 0.125f};
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(ys[a]) * w + xs[e]) * C), f);
        const float wt = k[a] * k[e];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += wt * f[j];
      }
  }
  if (add != nullptr) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(add + px * C + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
  }
  *reinterpret_cast<uint4*>(out + px * Ctot + c0 + c) = pack8(acc);
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
      reinterpret_cast<const __nv_bfloat16*>(in_bf16), reinterpret_cast<__nv_bfloat16*>(out_bf16), H, W, C, S);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int resize_common(int mode, const void* in, const void* add, void* out, int B, int h, int w, int H, int W, int C,
                         int Ctot, int c0, void* stream) {
  using namespace fz;
  FZ_REQUIRE(B >= 0 && h > 0 && w > 0 && H > 0 && W > 0, "fz_bilinear_slice: bad shape");
  FZ_REQUIRE(C % 8 == 0 && Ctot % 8 == 0 && c0 % 8 == 0 && c0 + C <= Ctot, "fz_bilinear_slice: bad channel slice");
  if (B == 0) return 0;
  const size_t n_vec = static_cast<size_t>(B) * H * W * (C / 8);
  const unsigned grid = static_cast<unsigned>((n_vec + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* i = reinterpret_cast<const __nv_bfloat16*>(in);
  const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(add);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (mode == 0) resize_slice_kernel<0><<<grid, 256, 0, st>>>(i, a, o, n_vec, h, w, H, W, C, Ctot, c0);
  else resize_slice_kernel<1><<<grid, 256, 0, st>>>(i, a, o, n_vec, h, w, H, W, C, Ctot, c0);
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

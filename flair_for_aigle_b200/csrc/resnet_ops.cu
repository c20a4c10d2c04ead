// CUDA-core kernels of the ResNet-34 encoder front end (smp native ResNetEncoder = torchvision ResNet
// minus fc; call site flair_hub/models/flair_model.py:376 for `resnet34-unet`, BASELINE.json configs[0]):
//   conv 7x7 / stride 2 / pad 3 (C_in <= 4) + eval BatchNorm + ReLU  -> bf16 NHWC
//   max-pool 3x3 / stride 2 / pad 1                                  -> bf16 NHWC
// The 3x3 convolutions of the BasicBlocks run on tcgen05 (conv3x3_tcgen05.cu, stride 1|2, residual epilogue).
#include "common.h"
#include "../../include/flair_zonal_b200.h"

#include "ptx.cuh"
#include "operand.cuh"

namespace fz {

// in : uint8 [B][P][P][4] (normalisation folded into w / bias by the host) or float [B][Cin][P][P]
// w  : float [196][64], row k = (ky*7 + kx)*4 + c
// out: bf16 [B][P/2][P/2][64] = relu((conv) * scale + bias)
constexpr int C7_SEG = 32;                    // output pixels per segment
constexpr int C7_INPX = 2 * C7_SEG + 5;       // 69 input pixels per row of a segment

template <bool F32IN>
__global__ void __launch_bounds__(128) conv7x7s2_kernel(const void* __restrict__ in_raw, int Cin,
                                                        const float* __restrict__ w, const float* __restrict__ scale,
                                                        const float* __restrict__ bias,
                                                        op_t* __restrict__ out, int P) {
  extern __shared__ float smem_f[];
  float* sW = smem_f;                        // [196][64]
  float* sIn = smem_f + 196 * 64;            // [7][C7_INPX + 3][4]
  constexpr int ROWF = (C7_INPX + 3) * 4;
  const int OW = P / 2;
  const int oy = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x;
  const int cg = tid & 15, pg = tid >> 4;    // 16 channel groups of 4, 8 pixel groups of 4
  for (int i = tid; i < 196 * 64; i += 128) sW[i] = w[i];
  float sc[4], bi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = scale[cg * 4 + j];
    bi[j] = bias[cg * 4 + j];
  }
  for (int seg = 0; seg < OW / C7_SEG; ++seg) {
    const int ox0 = seg * C7_SEG;
    __syncthreads();
    for (int i = tid; i < 7 * C7_INPX; i += 128) {
      const int r = i / C7_INPX, px = i % C7_INPX;
      const int gy = 2 * oy - 3 + r, gx = 2 * ox0 - 3 + px;
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < P && gx >= 0 && gx < P) {
        if (F32IN) {
          const float* xin = reinterpret_cast<const float*>(in_raw);
          const size_t o = (static_cast<size_t>(b) * Cin * P + gy) * P + gx;
          const size_t plane = static_cast<size_t>(P) * P;
          f.x = xin[o];
          if (Cin > 1) f.y = xin[o + plane];
          if (Cin > 2) f.z = xin[o + 2 * plane];
          if (Cin > 3) f.w = xin[o + 3 * plane];
        } else {
          const uchar4 u = reinterpret_cast<const uchar4*>(in_raw)[(static_cast<size_t>(b) * P + gy) * P + gx];
          f = make_float4(u.x, u.y, u.z, u.w);
        }
      }
      *reinterpret_cast<float4*>(&sIn[r * ROWF + px * 4]) = f;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[a][j] = 0.f;
    for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        float4 wv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) wv[c] = *reinterpret_cast<const float4*>(&sW[((ky * 7 + kx) * 4 + c) * 64 + cg * 4]);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float4 x4 = *reinterpret_cast<const float4*>(&sIn[ky * ROWF + ((pg * 4 + a) * 2 + kx) * 4]);
          acc[a][0] += x4.x * wv[0].x + x4.y * wv[1].x + x4.z * wv[2].x + x4.w * wv[3].x;
          acc[a][1] += x4.x * wv[0].y + x4.y * wv[1].y + x4.z * wv[2].y + x4.w * wv[3].y;
          acc[a][2] += x4.x * wv[0].z + x4.y * wv[1].z + x4.z * wv[2].z + x4.w * wv[3].z;
          acc[a][3] += x4.x * wv[0].w + x4.y * wv[1].w + x4.z * wv[2].w + x4.w * wv[3].w;
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int ox = ox0 + pg * 4 + a;
      op2_t lo = ff2op2(fmaxf(acc[a][0] * sc[0] + bi[0], 0.f), fmaxf(acc[a][1] * sc[1] + bi[1], 0.f));
      op2_t hi = ff2op2(fmaxf(acc[a][2] * sc[2] + bi[2], 0.f), fmaxf(acc[a][3] * sc[3] + bi[3], 0.f));
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(out + ((static_cast<size_t>(b) * OW + oy) * OW + ox) * 64 + cg * 4) = pk;
    }
  }
}

// in bf16 [B][H][W][C] -> out bf16 [B][H/2][W/2][C]; 8 channels per thread; padding never wins (-inf)
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const op_t* __restrict__ in,
                                                           op_t* __restrict__ out, size_t n_vec, int H, int W,
                                                           int C) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n_vec) return;
  const int vpp = C / 8;
  const int c = static_cast<int>(i % vpp) * 8;
  const size_t px = i / vpp;
  const int OW = W / 2, OH = H / 2;
  const int ox = static_cast<int>(px % OW), oy = static_cast<int>((px / OW) % OH);
  const size_t b = px / (static_cast<size_t>(OW) * OH);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int y = 2 * oy + dy;
    if (y < 0 || y >= H) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int x = 2 * ox + dx;
      if (x < 0 || x >= W) continue;
      const uint4 raw = *reinterpret_cast<const uint4*>(in + ((b * H + y) * W + x) * C + c);
      const op2_t* h = reinterpret_cast<const op2_t*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = op22ff(h[j]);
        m[2 * j] = fmaxf(m[2 * j], f.x);
        m[2 * j + 1] = fmaxf(m[2 * j + 1], f.y);
      }
    }
  }
  uint4 o;
  op2_t t;
  t = ff2op2(m[0], m[1]); o.x = *reinterpret_cast<uint32_t*>(&t);
  t = ff2op2(m[2], m[3]); o.y = *reinterpret_cast<uint32_t*>(&t);
  t = ff2op2(m[4], m[5]); o.z = *reinterpret_cast<uint32_t*>(&t);
  t = ff2op2(m[6], m[7]); o.w = *reinterpret_cast<uint32_t*>(&t);
  *reinterpret_cast<uint4*>(out + px * C + c) = o;
}

}  // namespace fz

extern "C" int fz_conv7x7s2_bn_relu(const void* in, int in_is_f32, int Cin, const float* w, const float* scale,
                                    const float* bias, void* out_bf16, int B, int P, void* stream) {
  using namespace fz;
  FZ_REQUIRE(P % 64 == 0 && Cin >= 1 && Cin <= 4, "fz_conv7x7s2_bn_relu: bad shape P=%d Cin=%d", P, Cin);
  if (B <= 0) return 0;
  const size_t smem = (196 * 64 + 7 * (C7_INPX + 3) * 4) * sizeof(float);
  FZ_ENSURE_SMEM(conv7x7s2_kernel<true>, static_cast<int>(smem));
  FZ_ENSURE_SMEM(conv7x7s2_kernel<false>, static_cast<int>(smem));
  dim3 grid(P / 2, B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  op_t* o = reinterpret_cast<op_t*>(out_bf16);
  if (in_is_f32) conv7x7s2_kernel<true><<<grid, 128, smem, st>>>(in, Cin, w, scale, bias, o, P);
  else conv7x7s2_kernel<false><<<grid, 128, smem, st>>>(in, 4, w, scale, bias, o, P);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_maxpool3x3s2(const void* in_bf16, void* out_bf16, int B, int H, int W, int C, void* stream) {
  using namespace fz;
  FZ_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "fz_maxpool3x3s2: bad shape H=%d W=%d C=%d", H, W, C);
  const size_t n_vec = static_cast<size_t>(B) * (H / 2) * (W / 2) * (C / 8);
  if (n_vec == 0) return 0;
  maxpool3x3s2_kernel<<<static_cast<unsigned>((n_vec + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const op_t*>(in_bf16), reinterpret_cast<op_t*>(out_bf16), n_vec, H, W, C);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Training-mode forward pieces and the backward of the ConvNeXt-V2 block, the encoder plumbing and the U-Net decoder:
//   x -> dwconv7x7 (+b) -> LayerNorm -> fc1 -> GELU -> GRN -> fc2 -> + x;  stem / downsample LayerNorms, space-to-depth, patches;
//   decoder: im2col / col2im for the deep layers, BatchNorm + ReLU on batch statistics, upsample-concat gradient.
// The Linear layers run on the tcgen05 GEMM (fz_gemm_bf16, native.linear_backward), the wide decoder convolutions on
// conv3x3_small.cu; the kernels here are the memory- and issue-bound rest.  Conventions: fixed reduction orders and no
// floating-point atomics (gradients are bit-reproducible, graph replays equal eager steps); one thread per 8-channel group
// with 16-byte accesses, blocks walking row chunks (reduce_vec.cuh); grids sized to one wave of resident blocks; forward
// activations in bf16 or IEEE fp16 (flag per entry point), gradients in bf16.  Saved tensors follow PyTorch's autograd of the
// same modules (oracle/models.py ConvNeXtBlock): the LayerNorm input with its row statistics, GELU(h) and GELU'(h), the GRN
// norms.  The first, one-thread-per-element versions of several kernels are kept as the fallback for shapes the vector forms
// do not take (channel counts that are not multiples of 8 / 4, ragged maps).
#include <cuda_bf16.h>

#include "common.h"
#include "reduce_vec.cuh"
#include "../../include/flair_zonal_b200.h"

namespace fz {

// ------------------------------------------------------------------------------------------------ depthwise 7x7, fp32
// out[b,y,x,c] = bias[c] + sum_k in[b,y+ky-3,x+kx-3,c] * w[k][c]   (flip: w[48-k], i.e. the data gradient)
__global__ void __launch_bounds__(256) dwconv7_f32_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ out, int H,
                                                          int W, int C, int flip) {
  // grid (ceil(W*C / 256), H, B): the row and the sample come from the grid, the rest is 32-bit index math
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= W * C) return;
  const int x = t / C, c = t - x * C, y = blockIdx.y;
  const size_t img = static_cast<size_t>(blockIdx.z) * H * W * C;
  float acc = bias ? bias[c] : 0.f;
#pragma unroll
  for (int ky = 0; ky < 7; ++ky) {
    const int iy = y + ky - 3;
    if (iy < 0 || iy >= H) continue;
    const float* row = in + img + static_cast<size_t>(iy) * W * C + c;
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
      const int ix = x + kx - 3;
      if (ix < 0 || ix >= W) continue;
      const int k = ky * 7 + kx;
      acc = fmaf(row[ix * C], w[(flip ? 48 - k : k) * C + c], acc);
    }
  }
  out[img + (static_cast<size_t>(y) * W + x) * C + c] = acc;
}

// Register-tiled form (W % 8 == 0): lane = channel, a warp computes 1 x 8 pixel tiles with the 49 taps in registers, four
// tiles per warp so that the tap loads are amortised: 7 x 14 input loads per 392 FMAs instead of one load per FMA.
// grid (C/32, ceil(n_tiles / 32)).
__global__ void __launch_bounds__(256) dwconv7_f32_tiled_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ out, int B,
                                                                int H, int W, int C, int flip) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  if (c >= C) return;
  const int tiles_x = W / 8, n_tiles = B * H * tiles_x;
  float wr[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) wr[k] = w[(flip ? 48 - k : k) * C + c];
  const float b0 = bias ? bias[c] : 0.f;
  for (int i = 0; i < 4; ++i) {
    const int t = (blockIdx.y * 8 + warp) * 4 + i;
    if (t >= n_tiles) return;
    const int row = t / tiles_x, tx = t - row * tiles_x, y = row % H, b = row / H;
    const int x0 = tx * 8;
    const size_t img = static_cast<size_t>(b) * H * W;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = b0;
#pragma unroll
    for (int ky = 0; ky < 7; ++ky) {
      const int iy = y + ky - 3;
      if (iy < 0 || iy >= H) continue;
      const float* xr = in + (img + static_cast<size_t>(iy) * W) * C + c;
      float xs[14];
#pragma unroll
      for (int q = 0; q < 14; ++q) {
        const int ix = x0 - 3 + q;
        xs[q] = (ix >= 0 && ix < W) ? xr[static_cast<size_t>(ix) * C] : 0.f;
      }
#pragma unroll
      for (int kx = 0; kx < 7; ++kx)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(xs[j + kx], wr[ky * 7 + kx], acc[j]);
    }
    float* op = out + (img + static_cast<size_t>(y) * W + x0) * C + c;
#pragma unroll
    for (int j = 0; j < 8; ++j) op[static_cast<size_t>(j) * C] = acc[j];
  }
}

// 4 x 8 pixel tiles (H % 4 == 0, W % 8 == 0): an input row loaded once feeds up to four output rows, 140 loads per 1568 FMAs
// (the 1 x 8 form above: 98 per 392), which is what bounds these kernels -- the loads hit L1/L2, the LSU issue rate is the
// limit.  `add` (optional, same layout as out) is summed into the result: the block backward's dx = dy + dwconv^T(du) without
// a separate pass.  grid (C/32, ceil(n_tiles / 16)), two tiles per warp.
// CT: the channel count as a compile-time constant (0 = runtime): every tap address becomes [row pointer + immediate]
// instead of a 64-bit multiply-add per load -- the instruction mix, not DRAM, bounds these kernels.
template <int CT>
__global__ void __launch_bounds__(256, 2) dwconv7_f32_tile48_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, const float* __restrict__ add,
                                                                 float* __restrict__ out, int B, int H, int W, int C_rt, int flip) {
  const int C = CT ? CT : C_rt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  if (c >= C) return;
  const int tiles_x = W / 8, tiles_y = H / 4, n_tiles = B * tiles_y * tiles_x;
  float wr[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) wr[k] = w[(flip ? 48 - k : k) * C + c];
  const float b0 = bias ? bias[c] : 0.f;
  for (int i = 0; i < 2; ++i) {
    const int t = (blockIdx.y * 8 + warp) * 2 + i;
    if (t >= n_tiles) return;
    const int trow = t / tiles_x, tx = t - trow * tiles_x, ty = trow % tiles_y, b = trow / tiles_y;
    const int x0 = tx * 8, y0 = ty * 4;
    const size_t img = static_cast<size_t>(b) * H * W;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = b0;
    // software-pipelined rows: row iy + 1 is in flight while row iy feeds the FMAs (16 resident warps per SM cannot hide the
    // load latency on their own).  Rows outside the image are skipped, loads and FMAs both (warp-uniform).
    const float* base = in + (img + static_cast<size_t>(y0 - 3) * W + x0) * C + c;      // pixel (y0 - 3, x0); only dereferenced inside
    const ptrdiff_t row_stride = static_cast<ptrdiff_t>(W) * C;
    auto load_row = [&](int iy, float (&dst)[14]) {
      const int gy = y0 - 3 + iy;
      if (gy < 0 || gy >= H) return;
      const float* xr = base + iy * row_stride;
#pragma unroll
      for (int q = 0; q < 14; ++q) {
        // only the three halo pixels on either side can fall outside the row (W % 8 == 0)
        const bool inside = q < 3 ? x0 - 3 + q >= 0 : (q >= 11 ? x0 - 3 + q < W : true);
        dst[q] = inside ? xr[(q - 3) * C] : 0.f;
      }
    };
    float nxt[14];
    load_row(0, nxt);
#pragma unroll
    for (int iy = 0; iy < 10; ++iy) {
      float xs[14];
#pragma unroll
      for (int q = 0; q < 14; ++q) xs[q] = nxt[q];
      if (iy + 1 < 10) load_row(iy + 1, nxt);
      const int gy = y0 - 3 + iy;
      if (gy < 0 || gy >= H) continue;                       // warp-uniform
#pragma unroll
      for (int oy = 0; oy < 4; ++oy) {
        const int ky = iy - oy;
        if (ky >= 0 && ky < 7) {
#pragma unroll
          for (int kx = 0; kx < 7; ++kx)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[oy * 8 + j] = fmaf(xs[j + kx], wr[ky * 7 + kx], acc[oy * 8 + j]);
        }
      }
    }
#pragma unroll
    for (int oy = 0; oy < 4; ++oy) {
      const size_t o = (img + static_cast<size_t>(y0 + oy) * W + x0) * C + c;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = acc[oy * 8 + j];
        if (add) v += add[o + j * C];
        out[o + j * C] = v;
      }
    }
  }
}

// partial[z][k][c] = sum over pixel chunk z of du[p][c] * x[p + tap k][c]  (k < 49);  k == 49: sum du.   grid (50, C/32, S)
__global__ void __launch_bounds__(256) dwconv7_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ du,
                                                            float* __restrict__ partial, int B, int H, int W, int C,
                                                            int px_per_chunk) {
  __shared__ float red[8][32];
  const int k = blockIdx.x, c = blockIdx.y * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  const int ky = k / 7, kx = k % 7;
  const int npx = B * H * W;                                   // < 2^31 (host check): 32-bit pixel arithmetic
  const int p0 = blockIdx.z * px_per_chunk;
  const int p1 = p0 + px_per_chunk < npx ? p0 + px_per_chunk : npx;
  const int shift = (ky - 3) * W + (kx - 3);
  float acc = 0.f;
  if (c < C)
    for (int p = p0 + ty; p < p1; p += 8) {
      const float g = du[static_cast<size_t>(p) * C + c];
      if (k == 49) {
        acc += g;
      } else {
        const int row = p / W, xw = p - row * W, y = row % H;
        const int iy = y + ky - 3, ix = xw + kx - 3;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) acc = fmaf(g, x[static_cast<size_t>(p + shift) * C + c], acc);
      }
    }
  red[ty][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    partial[(static_cast<size_t>(blockIdx.z) * 50 + k) * C + c] = t;
  }
}

// Register-tiled form of the same reduction (W % 8 == 0): lane = channel, a warp walks 1 x 8 pixel tiles and keeps all 49 tap
// sums + the bias sum in registers; per tile it loads 8 output gradients and 7 x 14 inputs for 392 MACs (the per-tap kernel
// above re-reads both operands for every tap: 13 GB through L2 per launch at stage 0, 1.5 ms).  grid (C/32, chunks).
__global__ void __launch_bounds__(256) dwconv7_wgrad_tiled_kernel(const float* __restrict__ x, const float* __restrict__ du,
                                                                  float* __restrict__ partial, int B, int H, int W, int C,
                                                                  int tiles_per_chunk) {
  __shared__ float red[50][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int tiles_x = W / 8, n_tiles = B * H * tiles_x;
  const int t0 = blockIdx.y * tiles_per_chunk;
  const int t1 = t0 + tiles_per_chunk < n_tiles ? t0 + tiles_per_chunk : n_tiles;
  float acc[50];
#pragma unroll
  for (int k = 0; k < 50; ++k) acc[k] = 0.f;
  if (c < C)
    for (int t = t0 + warp; t < t1; t += 8) {
      const int row = t / tiles_x, tx = t - row * tiles_x, y = row % H, b = row / H;
      const int x0 = tx * 8;
      const size_t img = static_cast<size_t>(b) * H * W;
      const float* gp = du + ((img + static_cast<size_t>(y) * W + x0) * C + c);
      float g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g[j] = gp[static_cast<size_t>(j) * C];
        acc[49] += g[j];
      }
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        const int iy = y + ky - 3;
        if (iy < 0 || iy >= H) continue;
        const float* xr = x + (img + static_cast<size_t>(iy) * W) * C + c;
        float xs[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) {
          const int ix = x0 - 3 + i;
          xs[i] = (ix >= 0 && ix < W) ? xr[static_cast<size_t>(ix) * C] : 0.f;
        }
#pragma unroll
        for (int kx = 0; kx < 7; ++kx)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[ky * 7 + kx] = fmaf(g[j], xs[j + kx], acc[ky * 7 + kx]);
      }
    }
  // the 8 warps' sums, added in warp order (fixed: reproducible)
  for (int w2 = 0; w2 < 8; ++w2) {
    if (warp == w2) {
#pragma unroll
      for (int k = 0; k < 50; ++k) red[k][lane] = (w2 == 0 ? 0.f : red[k][lane]) + acc[k];
    }
    __syncthreads();
  }
  if (c < C)
    for (int k = warp; k < 50; k += 8) partial[(static_cast<size_t>(blockIdx.y) * 50 + k) * C + c] = red[k][lane];
}

// The same reduction over 4 x 8 pixel tiles (H % 4 == 0): 32 output gradients and 10 x 14 inputs per 1568 MACs.
// grid (C/32, chunks); a chunk is a run of tiles_per_chunk tiles, its 8 warps take them round robin.
template <int CT>
__global__ void __launch_bounds__(256) dwconv7_wgrad_tile48_kernel(const float* __restrict__ x, const float* __restrict__ du,
                                                                   float* __restrict__ partial, int B, int H, int W, int C_rt,
                                                                   int tiles_per_chunk) {
  const int C = CT ? CT : C_rt;
  __shared__ float red[50][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int tiles_x = W / 8, tiles_y = H / 4, n_tiles = B * tiles_y * tiles_x;
  const int t0 = blockIdx.y * tiles_per_chunk;
  const int t1 = t0 + tiles_per_chunk < n_tiles ? t0 + tiles_per_chunk : n_tiles;
  float acc[50];
#pragma unroll
  for (int k = 0; k < 50; ++k) acc[k] = 0.f;
  if (c < C)
    for (int t = t0 + warp; t < t1; t += 8) {
      const int trow = t / tiles_x, tx = t - trow * tiles_x, ty = trow % tiles_y, b = trow / tiles_y;
      const int x0 = tx * 8, y0 = ty * 4;
      const size_t img = static_cast<size_t>(b) * H * W;
      float g[32];
#pragma unroll
      for (int oy = 0; oy < 4; ++oy) {
        const float* gp = du + ((img + static_cast<size_t>(y0 + oy) * W + x0) * C + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          g[oy * 8 + j] = gp[j * C];
          acc[49] += g[oy * 8 + j];
        }
      }
#pragma unroll
      for (int iy = 0; iy < 10; ++iy) {
        const int gy = y0 - 3 + iy;
        if (gy < 0 || gy >= H) continue;
        const float* xr = x + (img + static_cast<size_t>(gy) * W + x0) * C + c;
        float xs[14];
#pragma unroll
        for (int q = 0; q < 14; ++q) {
          const bool inside = q < 3 ? x0 - 3 + q >= 0 : (q >= 11 ? x0 - 3 + q < W : true);
          xs[q] = inside ? xr[(q - 3) * C] : 0.f;
        }
#pragma unroll
        for (int oy = 0; oy < 4; ++oy) {
          const int ky = iy - oy;
          if (ky >= 0 && ky < 7) {
#pragma unroll
            for (int kx = 0; kx < 7; ++kx)
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[ky * 7 + kx] = fmaf(g[oy * 8 + j], xs[j + kx], acc[ky * 7 + kx]);
          }
        }
      }
    }
  for (int w2 = 0; w2 < 8; ++w2) {                           // the 8 warps' sums, added in warp order (reproducible)
    if (warp == w2) {
#pragma unroll
      for (int k = 0; k < 50; ++k) red[k][lane] = (w2 == 0 ? 0.f : red[k][lane]) + acc[k];
    }
    __syncthreads();
  }
  if (c < C)
    for (int k = warp; k < 50; k += 8) partial[(static_cast<size_t>(blockIdx.y) * 50 + k) * C + c] = red[k][lane];
}

// ------------------------------------------------------------------------------------------------ LayerNorm rows
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per row: out = (x - mean) * rstd * g + b (bf16), mean / rstd saved
__global__ void __launch_bounds__(256) ln_fwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                           const float* __restrict__ b, __nv_bfloat16* __restrict__ out,
                                                           float* __restrict__ out_f32, float* __restrict__ mean,
                                                           float* __restrict__ rstd, int64_t M, int C, float eps, int out_f16,
                                                           __nv_bfloat16* __restrict__ out2) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= M) return;
  const float* row = x + r * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += row[c];
  const float mu = warp_sum(s) / C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = row[c] - mu;
    v += d * d;
  }
  const float rs = rsqrtf(warp_sum(v) / C + eps);
  for (int c = lane; c < C; c += 32) {
    const float v2 = (row[c] - mu) * rs * g[c] + b[c];
    if (out) {
      if (out_f16) reinterpret_cast<__half*>(out)[r * C + c] = __float2half_rn(fminf(fmaxf(v2, -65504.f), 65504.f));
      else out[r * C + c] = __float2bfloat16_rn(v2);
    }
    if (out_f32) out_f32[r * C + c] = v2;
    if (out2) out2[r * C + c] = __float2bfloat16_rn(v2);
  }
  if (lane == 0) {
    mean[r] = mu;
    rstd[r] = rs;
  }
}

// dx = rstd * (dy*g - mean_c(dy*g) - xhat * mean_c(dy*g*xhat));  partial[blk][0][c] = sum dy*xhat, [1][c] = sum dy over the
// block's rows (block b's warp w walks rows (b*8 + w) + k * gridDim.x*8: fixed assignment, fixed order)
__global__ void __launch_bounds__(256) ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ g, float* __restrict__ dx,
                                                     float* __restrict__ partial, int64_t M, int C) {
  extern __shared__ float sh[];                      // [8 warps][2][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* mine = sh + static_cast<size_t>(warp) * 2 * C;
  for (int c = lane; c < 2 * C; c += 32) mine[c] = 0.f;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + warp; r < M; r += static_cast<int64_t>(gridDim.x) * 8) {
    const float mu = mean[r], rs = rstd[r];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xh = (x[r * C + c] - mu) * rs, d = __bfloat162float(dy[r * C + c]), gg = d * g[c];
      s1 += gg;
      s2 += gg * xh;
      mine[c] += d * xh;
      mine[C + c] += d;
    }
    const float m1 = warp_sum(s1) / C, m2 = warp_sum(s2) / C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (x[r * C + c] - mu) * rs, gg = __bfloat162float(dy[r * C + c]) * g[c];
      dx[r * C + c] = rs * (gg - m1 - xh * m2);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) t += sh[static_cast<size_t>(w2) * 2 * C + c];
    partial[static_cast<size_t>(blockIdx.x) * 2 * C + c] = t;
  }
}

__device__ __forceinline__ void load4_bf16(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 q = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(q.x << 16);
  v[1] = __uint_as_float(q.x & 0xffff0000u);
  v[2] = __uint_as_float(q.y << 16);
  v[3] = __uint_as_float(q.y & 0xffff0000u);
}
// "value > 0" on the bit pattern of a 16-bit float: holds for bf16 and IEEE fp16 alike (sign bit clear, magnitude non-zero;
// neither format's ReLU outputs are NaN), so the ReLU mask of the backward needs no format flag
__device__ __forceinline__ bool positive16(__nv_bfloat16 v) {
  const unsigned short b = *reinterpret_cast<const unsigned short*>(&v);
  return b != 0 && (b & 0x8000u) == 0;
}
__device__ __forceinline__ void positive16x4(const __nv_bfloat16* p, bool (&pos)[4]) {
  const uint2 q = *reinterpret_cast<const uint2*>(p);
  const unsigned w[4] = {q.x & 0xffffu, q.x >> 16, q.y & 0xffffu, q.y >> 16};
#pragma unroll
  for (int j = 0; j < 4; ++j) pos[j] = w[j] != 0 && (w[j] & 0x8000u) == 0;
}
__device__ __forceinline__ void store4_16(__nv_bfloat16* p, const float (&v)[4], bool f16) {   // bf16 or (f16) IEEE half
  uint2 q;
  q.x = rv_pack2(v[0], v[1], f16);
  q.y = rv_pack2(v[2], v[3], f16);
  *reinterpret_cast<uint2*>(p) = q;
}
__device__ __forceinline__ void load4_16(const __nv_bfloat16* p, float (&v)[4], bool f16) {
  const uint2 q = *reinterpret_cast<const uint2*>(p);
  if (f16) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&q.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
    v[0] = __uint_as_float(q.x << 16);
    v[1] = __uint_as_float(q.x & 0xffff0000u);
    v[2] = __uint_as_float(q.y << 16);
    v[3] = __uint_as_float(q.y & 0xffff0000u);
  }
}
__device__ __forceinline__ void store4_bf16(__nv_bfloat16* p, const float (&v)[4]) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 q;
  q.x = *reinterpret_cast<const uint32_t*>(&a);
  q.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = q;
}
// ---- the same two kernels for C = 128 * NV: a lane owns NV float4 quads of the row (quad k*32 + lane) and keeps them in
// registers, so the row is read from global memory once (the scalar forms re-read it for every pass, 4 bytes per load) and the
// per-channel sums of the backward live in registers instead of a shared-memory read-modify-write per element.
template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                         const float* __restrict__ b, __nv_bfloat16* __restrict__ out,
                                                         float* __restrict__ out_f32, float* __restrict__ mean,
                                                         float* __restrict__ rstd, int64_t M, float eps, int out_f16,
                                                         __nv_bfloat16* __restrict__ out2) {
  constexpr int C = 128 * NV;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= M) return;
  const float4* row = reinterpret_cast<const float4*>(x + r * C);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    v[k] = row[k * 32 + lane];
    s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  const float mu = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    v[k].x -= mu; v[k].y -= mu; v[k].z -= mu; v[k].w -= mu;
    q += (v[k].x * v[k].x + v[k].y * v[k].y) + (v[k].z * v[k].z + v[k].w * v[k].w);
  }
  const float rs = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * 4;
    const float4 gg = *reinterpret_cast<const float4*>(g + c), bb = *reinterpret_cast<const float4*>(b + c);
    const float o[4] = {v[k].x * rs * gg.x + bb.x, v[k].y * rs * gg.y + bb.y, v[k].z * rs * gg.z + bb.z,
                        v[k].w * rs * gg.w + bb.w};
    if (out) store4_16(out + r * C + c, o, out_f16);
    if (out2) store4_16(out2 + r * C + c, o, false);
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + r * C + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
  if (lane == 0) {
    mean[r] = mu;
    rstd[r] = rs;
  }
}

template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_vec_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         const float* __restrict__ g, float* __restrict__ dx,
                                                         float* __restrict__ partial, int64_t M) {
  constexpr int C = 128 * NV;
  extern __shared__ float sh[];                      // [8 warps][2][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 gam[NV];
  float a_dg[NV][4], a_db[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    gam[k] = *reinterpret_cast<const float4*>(g + (k * 32 + lane) * 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) a_dg[k][j] = a_db[k][j] = 0.f;
  }
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + warp; r < M; r += static_cast<int64_t>(gridDim.x) * 8) {
    const float mu = mean[r], rs = rstd[r];
    float xh[NV][4], gg[NV][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 4;
      const float4 xv = *reinterpret_cast<const float4*>(x + r * C + c);
      float d[4];
      load4_bf16(dy + r * C + c, d);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      const float gs[4] = {gam[k].x, gam[k].y, gam[k].z, gam[k].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xh[k][j] = (xs[j] - mu) * rs;
        gg[k][j] = d[j] * gs[j];
        s1 += gg[k][j];
        s2 += gg[k][j] * xh[k][j];
        a_dg[k][j] += d[j] * xh[k][j];
        a_db[k][j] += d[j];
      }
    }
    const float m1 = warp_sum(s1) / C, m2 = warp_sum(s2) / C;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 4;
      *reinterpret_cast<float4*>(dx + r * C + c) =
          make_float4(rs * (gg[k][0] - m1 - xh[k][0] * m2), rs * (gg[k][1] - m1 - xh[k][1] * m2),
                      rs * (gg[k][2] - m1 - xh[k][2] * m2), rs * (gg[k][3] - m1 - xh[k][3] * m2));
    }
  }
  float* mine = sh + static_cast<size_t>(warp) * 2 * C;
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mine[(k * 32 + lane) * 4 + j] = a_dg[k][j];
      mine[C + (k * 32 + lane) * 4 + j] = a_db[k][j];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) t += sh[static_cast<size_t>(w2) * 2 * C + c];
    partial[static_cast<size_t>(blockIdx.x) * 2 * C + c] = t;
  }
}

// out[n] = sum_s partial[s][n]: 32 columns x 8 row lanes per block, lane k adds rows k, k+8, ... in order and the eight lane
// sums are added in lane order (a fixed order; the one-thread-per-column form chained up to 592 dependent loads).
// grid ceil(N / 32)
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ partial, float* __restrict__ out, int N,
                                                          int S) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + (threadIdx.x & 31), k = threadIdx.x >> 5;
  float t = 0.f;
  if (n < N)
    for (int s = k; s < S; s += 8) t += partial[static_cast<size_t>(s) * N + n];
  red[k][threadIdx.x & 31] = t;
  __syncthreads();
  if (k == 0 && n < N) {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v += red[j][threadIdx.x];
    out[n] = v;
  }
}

// ------------------------------------------------------------------------------------------------ GELU (erf)
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = __uint_as_float(w[j] << 16);
    v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 q;
  uint32_t* w = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = q;
}
// n is a multiple of 8 (C4 is a multiple of 8)
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ g,
                                                       int64_t n) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) * 8;
  if (i >= n) return;
  float v[8];
  load8(h + i, v);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = gelu_exact(v[j]);
  store8(g + i, v);
}

// ------------------------------------------------------------------------------------------------ GRN
// out[b][c] = sum_hw f(a, bb): mode 0 a*a, 1 a*bb, 2 a.   grid (C/32, B), 32 channels x 8 row lanes per block
__global__ void __launch_bounds__(256) sample_colreduce_kernel(const __nv_bfloat16* __restrict__ a,
                                                               const __nv_bfloat16* __restrict__ bb, float* __restrict__ out,
                                                               int HW, int C, int mode) {
  __shared__ float red[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  const int64_t base = static_cast<int64_t>(blockIdx.y) * HW * C;
  float acc = 0.f;
  if (c < C)
    for (int r = ty; r < HW; r += 8) {
      const float v = __bfloat162float(a[base + static_cast<int64_t>(r) * C + c]);
      acc += mode == 0 ? v * v : (mode == 1 ? v * __bfloat162float(bb[base + static_cast<int64_t>(r) * C + c]) : v);
    }
  red[ty][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    out[static_cast<size_t>(blockIdx.y) * C + c] = t;
  }
}

// one block per sample: Gx = sqrt(sumsq), mu = mean_c Gx, Nx = Gx / (mu + eps)
__global__ void __launch_bounds__(256) grn_norms_kernel(const float* __restrict__ sumsq, float* __restrict__ gx,
                                                        float* __restrict__ nx, float* __restrict__ mu_out, int C, float eps) {
  __shared__ float red[256];
  const int b = blockIdx.x;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float v = sqrtf(sumsq[static_cast<size_t>(b) * C + c]);
    gx[static_cast<size_t>(b) * C + c] = v;
    s += v;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const float mu = red[0] / C;
  if (threadIdx.x == 0) mu_out[b] = mu;
  for (int c = threadIdx.x; c < C; c += 256) nx[static_cast<size_t>(b) * C + c] = gx[static_cast<size_t>(b) * C + c] / (mu + eps);
}

// y = g * (1 + gamma * Nx[b]) + beta.   grid (ceil(per_sample / 2048), B), 8 channels per thread (C % 8 == 0)
__global__ void __launch_bounds__(256) grn_apply_train_kernel(const __nv_bfloat16* __restrict__ g, const float* __restrict__ nx,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              __nv_bfloat16* __restrict__ y, int per_sample, int C) {
  const int e = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (e >= per_sample) return;
  const int c = e % C;
  const size_t b = blockIdx.y, i = b * per_sample + e;
  float v[8];
  load8(g + i, v);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = v[j] * (1.0f + gamma[c + j] * nx[b * C + c + j]) + beta[c + j];
  store8(y + i, v);
}

// one block per sample: coefA = 1 + gamma*Nx, coefB = dGx / Gx with
// dGx = gamma*S1/(mu+eps) - (sum_c gamma*S1*Gx) / (C (mu+eps)^2)
__global__ void __launch_bounds__(256) grn_bwd_coef_kernel(const float* __restrict__ s1, const float* __restrict__ gx,
                                                           const float* __restrict__ nx, const float* __restrict__ mu,
                                                           const float* __restrict__ gamma, float* __restrict__ coef_a,
                                                           float* __restrict__ coef_b, int C, float eps) {
  __shared__ float red[256];
  const int b = blockIdx.x;
  const size_t o = static_cast<size_t>(b) * C;
  float t = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) t += gamma[c] * s1[o + c] * gx[o + c];
  red[threadIdx.x] = t;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  const float d = mu[b] + eps, tb = red[0] / (static_cast<float>(C) * d * d);
  for (int c = threadIdx.x; c < C; c += 256) {
    coef_a[o + c] = 1.0f + gamma[c] * nx[o + c];
    const float dgx = gamma[c] * s1[o + c] / d - tb;
    coef_b[o + c] = gx[o + c] > 0.f ? dgx / gx[o + c] : 0.f;
  }
}

// dgamma[c] = sum_b Nx*S1, dbeta[c] = sum_b S0
__global__ void __launch_bounds__(256) grn_param_grad_kernel(const float* __restrict__ s1, const float* __restrict__ s0,
                                                             const float* __restrict__ nx, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, int B, int C) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, bsum = 0.f;
  for (int b = 0; b < B; ++b) {
    a += nx[static_cast<size_t>(b) * C + c] * s1[static_cast<size_t>(b) * C + c];
    bsum += s0[static_cast<size_t>(b) * C + c];
  }
  dgamma[c] = a;
  dbeta[c] = bsum;
}

// dh = (dy * coefA[b] + g * coefB[b]) * GELU'(h).   grid (ceil(per_sample / 2048), B), 8 channels per thread
__global__ void __launch_bounds__(256) grn_gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ g,
                                                           const __nv_bfloat16* __restrict__ h, const float* __restrict__ coef_a,
                                                           const float* __restrict__ coef_b, __nv_bfloat16* __restrict__ dh,
                                                           int per_sample, int C) {
  const int e = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (e >= per_sample) return;
  const int c = e % C;
  const size_t b = blockIdx.y, i = b * per_sample + e;
  float d[8], gg[8], hh[8];
  load8(dy + i, d);
  load8(g + i, gg);
  load8(h + i, hh);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    d[j] = (d[j] * coef_a[b * C + c + j] + gg[j] * coef_b[b * C + c + j]) * gelu_grad(hh[j]);
  store8(dh + i, d);
}

// Row-walking forms of the two kernels above (reduce_vec.cuh's geometry: a thread owns one 8-channel group, the block walks
// its row chunk): the per-(sample, channel) coefficients are loaded ONCE into registers instead of 16-24 scalar loads per
// 16-byte payload load.  grid (slabs, chunks, B).
__global__ void __launch_bounds__(256) grn_apply_rows_kernel(const __nv_bfloat16* __restrict__ g, const float* __restrict__ nx,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ y2,
                                                             int rows, int C, int cgs, int rows_per_chunk, int act_f16) {
  const int rp = 256 / cgs;
  const int cl = threadIdx.x % cgs, rr = threadIdx.x / cgs;
  const int cg = blockIdx.x * cgs + cl;
  if (rr >= rp || cg * 8 >= C) return;
  const size_t b = blockIdx.z;
  float ca[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ca[j] = 1.0f + gamma[cg * 8 + j] * nx[b * C + cg * 8 + j];
    be[j] = beta[cg * 8 + j];
  }
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  const size_t base = b * rows * C + static_cast<size_t>(cg) * 8;
#pragma unroll 4
  for (int r = r0 + rr; r < r1; r += rp) {
    const size_t o = base + static_cast<size_t>(r) * C;
    float v[8];
    rv_load8(g + o, v, act_f16);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = v[j] * ca[j] + be[j];
    float v2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v2[j] = v[j];
    rv_store8_round(y + o, v, act_f16);
    if (y2) rv_store8_round(y2 + o, v2, false);        // a bf16 copy of the same values: the weight gradient's operand
  }
}

// dh = (dy * coefA[b] + g * coefB[b]) * GELU'(h), and the column sums of the stored (rounded) dh per (sample, chunk): fc1's
// bias gradient without another pass over dh.   db_partial[(b * chunks + chunk) * C + c]
__global__ void __launch_bounds__(256) grn_gelu_bwd_rows_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                const __nv_bfloat16* __restrict__ g,
                                                                const __nv_bfloat16* __restrict__ h,
                                                                const float* __restrict__ coef_a, const float* __restrict__ coef_b,
                                                                __nv_bfloat16* __restrict__ dh, float* __restrict__ db_partial,
                                                                int rows, int C, int cgs, int rows_per_chunk, int h_is_dgelu,
                                                                int act_f16) {
  __shared__ float red[256][9];
  const int rp = 256 / cgs;
  const int cl = threadIdx.x % cgs, rr = threadIdx.x / cgs;
  const int cg = blockIdx.x * cgs + cl;
  const bool active = rr < rp && cg * 8 < C;
  const size_t b = blockIdx.z;
  float ca[8], cb[8], acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ca[j] = active ? coef_a[b * C + cg * 8 + j] : 0.f;
    cb[j] = active ? coef_b[b * C + cg * 8 + j] : 0.f;
    acc[j] = 0.f;
  }
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  const size_t base = b * rows * C + static_cast<size_t>(cg) * 8;
  if (active) {
#pragma unroll 2
    for (int r = r0 + rr; r < r1; r += rp) {
      const size_t o = base + static_cast<size_t>(r) * C;
      float d[8], gg[8], hh[8];
      load8(dy + o, d);
      rv_load8(g + o, gg, act_f16);
      rv_load8(h + o, hh, act_f16);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = (d[j] * ca[j] + gg[j] * cb[j]) * (h_is_dgelu ? hh[j] : rv_gelu_grad(hh[j]));
      rv_store8_round(dh + o, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += d[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (rr == 0 && cg * 8 < C) {
    float* dst = db_partial + (b * gridDim.y + blockIdx.y) * C + static_cast<size_t>(cg) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = 0.f;
      for (int q = 0; q < rp; ++q) t += red[q * cgs + cl][j];
      dst[j] = t;
    }
  }
}

__global__ void __launch_bounds__(256) add_f32x4_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                        float4* __restrict__ out, int64_t n4) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 u = a[i], v = b[i];
  out[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
}

// IEEE fp16 -> bf16, 8 values per thread (the tail one by one): the weight-gradient GEMMs want both operands in one format
// (a tcgen05 kind::f16 MMA with a_format = BF16 and b_format = F16 is an illegal instruction on sm_100a -- tried), so the
// fp16 forward activations are re-rounded to bf16 on their way into dW = dY^T X
__global__ void __launch_bounds__(256) cast_f16_bf16_kernel(const __half* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                            int64_t n) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float v[8];
    rv_load8(in + i, v, true);
    rv_store8_round(out + i, v, false);
  } else {
    for (int64_t j = i; j < n; ++j) out[j] = __float2bfloat16_rn(__half2float(in[j]));
  }
}

__global__ void __launch_bounds__(256) add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                      float* __restrict__ out, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// space-to-depth for a k x k / stride k convolution as a GEMM: out[b][y/s][x/s][(ky*s + kx)*C + c] = in[b][y][x][c]
// (inverse: the same index map read the other way -- the data gradient of the gather)
__global__ void __launch_bounds__(256) s2d_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                       int H, int W, int C, int s, int inverse, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;      // index into the [B][H][W][C] side
  if (i >= n) return;
  const int c = static_cast<int>(i % C);
  const int64_t px = i / C;
  const int x = static_cast<int>(px % W), y = static_cast<int>((px / W) % H);
  const int64_t b = px / (static_cast<int64_t>(W) * H);
  const int64_t j = (((b * (H / s) + y / s) * (W / s) + x / s) * (s * s) + (y % s) * s + (x % s)) * C + c;
  if (inverse) out[i] = in[j];
  else out[j] = in[i];
}

// the same map moving 8 channels (16 bytes) per thread, C % 8 == 0
__global__ void __launch_bounds__(256) s2d_bf16_vec_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H, int W,
                                                           int C8, int s, int inverse, int64_t n8) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;      // index into the [B][H][W][C/8] side
  if (i >= n8) return;
  const int c = static_cast<int>(i % C8);
  const int64_t px = i / C8;
  const int x = static_cast<int>(px % W), y = static_cast<int>((px / W) % H);
  const int64_t b = px / (static_cast<int64_t>(W) * H);
  const int64_t j = (((b * (H / s) + y / s) * (W / s) + x / s) * (s * s) + (y % s) * s + (x % s)) * C8 + c;
  if (inverse) out[i] = in[j];
  else out[j] = in[i];
}

// stem patches: out bf16 [B*(P/4)^2][Kpad], k = (c*4 + ky)*4 + kx (the flattening of a [Cout][Cin][4][4] weight), zero padded
__global__ void __launch_bounds__(256) patchify4_nchw_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                             int Cin, int P, int Kpad, int64_t n, int out_f16) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;      // over rows * Kpad
  if (i >= n) return;
  const int k = static_cast<int>(i % Kpad);
  const int64_t row = i / Kpad;
  const int q = P / 4;
  const int px = static_cast<int>(row % q), py = static_cast<int>((row / q) % q);
  const int64_t b = row / (static_cast<int64_t>(q) * q);
  float v = 0.f;
  if (k < Cin * 16) {
    const int c = k / 16, ky = (k / 4) % 4, kx = k % 4;
    v = in[((b * Cin + c) * P + py * 4 + ky) * P + px * 4 + kx];
  }
  if (out_f16) reinterpret_cast<__half*>(out)[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  else out[i] = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------------ U-Net decoder
// col[p][(ky*3+kx)*C + c] = in[p + (ky-1, kx-1)][c] (zero outside the image and in the K padding)
__global__ void __launch_bounds__(256) im2col3x3_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ col,
                                                        int H, int W, int C, int Kpad, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const int k = static_cast<int>(i % Kpad);
  const int64_t p = i / Kpad;
  __nv_bfloat16 v = __float2bfloat16_rn(0.f);
  if (k < 9 * C) {
    const int tap = k / C, c = k - tap * C, ky = tap / 3, kx = tap - ky * 3;
    const int x = static_cast<int>(p % W), y = static_cast<int>((p / W) % H);
    const int iy = y + ky - 1, ix = x + kx - 1;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = in[(p + static_cast<int64_t>(ky - 1) * W + (kx - 1)) * C + c];
  }
  col[i] = v;
}

// dx[q][c] = sum_taps dcol[q - (ky-1, kx-1)][(ky*3+kx)*C + c]
__global__ void __launch_bounds__(256) col2im3x3_kernel(const __nv_bfloat16* __restrict__ dcol, float* __restrict__ dx, int H,
                                                        int W, int C, int Kpad, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const int c = static_cast<int>(i % C);
  const int64_t q = i / C;
  const int x = static_cast<int>(q % W), y = static_cast<int>((q / W) % H);
  float acc = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int py = y - (ky - 1), px = x - (kx - 1);
      if (py >= 0 && py < H && px >= 0 && px < W)
        acc += __bfloat162float(dcol[(q - static_cast<int64_t>(ky - 1) * W - (kx - 1)) * Kpad + (ky * 3 + kx) * C + c]);
    }
  dx[i] = acc;
}

// 8 channels (16 bytes) per thread, C % 8 == 0: one thread per (pixel, 8-wide slice of the Kpad row); 32-bit tap arithmetic
__global__ void __launch_bounds__(256) im2col3x3_vec_kernel(const uint4* __restrict__ in, uint4* __restrict__ col, int H, int W,
                                                            int C8, int K8, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const int k8 = static_cast<int>(i % K8);
  const int64_t p = i / K8;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (k8 < 9 * C8) {
    const int tap = k8 / C8, c8 = k8 - tap * C8, ky = tap / 3, kx = tap - ky * 3;
    const int x = static_cast<int>(p % W), y = static_cast<int>((p / W) % H);
    const int iy = y + ky - 1, ix = x + kx - 1;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = in[(p + static_cast<int64_t>(ky - 1) * W + (kx - 1)) * C8 + c8];
  }
  col[i] = v;
}

// one thread per (pixel, 8 channels): nine 16-byte reads of dcol, two float4 writes
__global__ void __launch_bounds__(256) col2im3x3_vec_kernel(const __nv_bfloat16* __restrict__ dcol, float* __restrict__ dx, int H,
                                                            int W, int C, int Kpad, int64_t n8) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n8) return;
  const int C8 = C / 8;
  const int c = static_cast<int>(i % C8) * 8;
  const int64_t q = i / C8;
  const int x = static_cast<int>(q % W), y = static_cast<int>((q / W) % H);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int py = y - (ky - 1), px = x - (kx - 1);
      if (py >= 0 && py < H && px >= 0 && px < W) {
        float v[8];
        load8(dcol + (q - static_cast<int64_t>(ky - 1) * W - (kx - 1)) * Kpad + (ky * 3 + kx) * C + c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    }
  float4* o = reinterpret_cast<float4*>(dx + q * C + c);
  o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// partial[s][0][c] = sum_rows f0, partial[s][1][c] = sum_rows f1 over row chunk s.   grid (C/32, chunks)
//   mode 0 (BN forward, pass 1): f0 = x                      mode 1 (pass 2): f0 = (x - mean)^2
//   mode 2 (BN+ReLU backward):   f0 = g, f1 = g * xhat with g = dy * (y > 0), xhat = (x - mean) * rstd
__global__ void __launch_bounds__(256) bn_partial_kernel(const float* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ dy,
                                                         const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                                         const float* __restrict__ rstd, float* __restrict__ partial, int64_t M,
                                                         int C, int rows_per_chunk, int mode) {
  __shared__ float r0[8][32], r1[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * rows_per_chunk;
  const int64_t m1 = m0 + rows_per_chunk < M ? m0 + rows_per_chunk : M;
  float a = 0.f, b = 0.f;
  if (c < C) {
    const float mu = mode ? mean[c] : 0.f, rs = mode == 2 ? rstd[c] : 0.f;
    for (int64_t m = m0 + ty; m < m1; m += 8) {
      const float xv = x[m * ldx + c];
      if (mode == 0) a += xv;
      else if (mode == 1) a += (xv - mu) * (xv - mu);
      else {
        const float g = positive16(y[m * C + c]) ? __bfloat162float(dy[m * C + c]) : 0.f;
        a += g;
        b += g * (xv - mu) * rs;
      }
    }
  }
  r0[ty][threadIdx.x & 31] = a;
  r1[ty][threadIdx.x & 31] = b;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      t0 += r0[j][threadIdx.x];
      t1 += r1[j][threadIdx.x];
    }
    partial[(static_cast<size_t>(blockIdx.y) * 2) * C + c] = t0;
    partial[(static_cast<size_t>(blockIdx.y) * 2 + 1) * C + c] = t1;
  }
}
// mode 0: mean = sum / M; mode 1: rstd = rsqrt(sum / M + eps) (biased variance, what BatchNorm normalises with)
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ sums, float* __restrict__ out, int C, float M,
                                                          float eps, int mode) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < C) out[c] = mode == 0 ? sums[c] / M : rsqrtf(sums[c] / M + eps);
}
// y = relu((x - mean) * rstd * gamma + beta)
__global__ void __launch_bounds__(256) bn_relu_apply_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                            int C, int64_t n, int out_f16) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const int c = static_cast<int>(i % C);
  const int64_t m = i / C;
  const float v = (x[m * ldx + c] - mean[c]) * rstd[c] * gamma[c] + beta[c];
  if (out_f16) reinterpret_cast<__half*>(y)[i] = __float2half_rn(v > 0.f ? fminf(v, 65504.f) : 0.f);
  else y[i] = __float2bfloat16_rn(v > 0.f ? v : 0.f);
}
// dx[m][c] = gamma * rstd * (g - dbeta / M - xhat * dgamma / M), written bf16 with row stride ldd (padding columns zero)
__global__ void __launch_bounds__(256) bn_relu_bwd_apply_kernel(const float* __restrict__ x, int ldx,
                                                                const __nv_bfloat16* __restrict__ dy,
                                                                const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                const float* __restrict__ dgb, __nv_bfloat16* __restrict__ dx,
                                                                int ldd, int C, float invM, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;      // over M * ldd
  if (i >= n) return;
  const int c = static_cast<int>(i % ldd);
  const int64_t m = i / ldd;
  float v = 0.f;
  if (c < C) {
    const float g = positive16(y[m * C + c]) ? __bfloat162float(dy[m * C + c]) : 0.f;
    const float xh = (x[m * ldx + c] - mean[c]) * rstd[c];
    v = gamma[c] * rstd[c] * (g - dgb[c] * invM - xh * dgb[C + c] * invM);      // dgb = [dbeta | dgamma]
  }
  dx[i] = __float2bfloat16_rn(v);
}

// ---- the same three kernels, four channels (one float4 / one 8-byte bf16x4) per thread; C, ldx, ldd multiples of 4.
// The scalar forms above move 4 bytes per thread per row and leave the 16-channel, 4-million-row layers of the decoder with
// 256 blocks of work; here a thread owns one channel quad and the block walks its row chunk 256 / (C/4) rows at a time.
// grid (slabs, chunks); partial[chunk][2][C] as above
template <int MODE>
__global__ void __launch_bounds__(256) bn_partial_vec_kernel(const float* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ dy,
                                                             const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, float* __restrict__ partial,
                                                             int64_t M, int C, int cgs, int rows_per_chunk) {
  __shared__ float red[2][256][5];
  const int rp = 256 / cgs;
  const int cl = threadIdx.x % cgs, rr = threadIdx.x / cgs;
  const int c = (blockIdx.x * cgs + cl) * 4;
  const bool active = rr < rp && c < C;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * rows_per_chunk;
  const int64_t m1 = m0 + rows_per_chunk < M ? m0 + rows_per_chunk : M;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    float mu[4] = {0.f, 0.f, 0.f, 0.f}, rs[4] = {0.f, 0.f, 0.f, 0.f};
    if (MODE >= 1) {
      const float4 t = *reinterpret_cast<const float4*>(mean + c);
      mu[0] = t.x; mu[1] = t.y; mu[2] = t.z; mu[3] = t.w;
    }
    if (MODE == 2) {
      const float4 t = *reinterpret_cast<const float4*>(rstd + c);
      rs[0] = t.x; rs[1] = t.y; rs[2] = t.z; rs[3] = t.w;
    }
#pragma unroll 4
    for (int64_t m = m0 + rr; m < m1; m += rp) {
      const float4 t = *reinterpret_cast<const float4*>(x + m * ldx + c);
      const float xv[4] = {t.x, t.y, t.z, t.w};
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] += xv[j];
      } else if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] += (xv[j] - mu[j]) * (xv[j] - mu[j]);
      } else {
        float dv[4];
        bool pos[4];
        positive16x4(y + m * C + c, pos);
        load4_bf16(dy + m * C + c, dv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float g = pos[j] ? dv[j] : 0.f;
          a[j] += g;
          b[j] += g * (xv[j] - mu[j]) * rs[j];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[0][threadIdx.x][j] = a[j];
    red[1][threadIdx.x][j] = b[j];
  }
  __syncthreads();
  if (rr == 0 && c < C) {
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = 0.f;
        for (int q = 0; q < rp; ++q) t += red[o][q * cgs + cl][j];
        partial[(static_cast<size_t>(blockIdx.y) * 2 + o) * C + c + j] = t;
      }
  }
}
__global__ void __launch_bounds__(256) bn_relu_apply_vec_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                                int C4, int64_t n4, int out_f16) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n4) return;
  const int c = static_cast<int>(i % C4) * 4;
  const int64_t m = i / C4;
  const float4 xv = *reinterpret_cast<const float4*>(x + m * ldx + c);
  const float4 mu = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
  const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
  float v[4] = {(xv.x - mu.x) * rs.x * ga.x + be.x, (xv.y - mu.y) * rs.y * ga.y + be.y, (xv.z - mu.z) * rs.z * ga.z + be.z,
                (xv.w - mu.w) * rs.w * ga.w + be.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = v[j] > 0.f ? v[j] : 0.f;
  store4_16(y + m * (C4 * 4) + c, v, out_f16);
}
__global__ void __launch_bounds__(256) bn_relu_bwd_apply_vec_kernel(const float* __restrict__ x, int ldx,
                                                                    const __nv_bfloat16* __restrict__ dy,
                                                                    const __nv_bfloat16* __restrict__ y,
                                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                    const float* __restrict__ gamma, const float* __restrict__ dgb,
                                                                    __nv_bfloat16* __restrict__ dx, int ldd, int C, float invM,
                                                                    int64_t n4) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;      // over M * ldd / 4
  if (i >= n4) return;
  const int L4 = ldd / 4;
  const int c = static_cast<int>(i % L4) * 4;
  const int64_t m = i / L4;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    float dv[4];
    bool pos[4];
    positive16x4(y + m * C + c, pos);
    load4_bf16(dy + m * C + c, dv);
    const float4 t = *reinterpret_cast<const float4*>(x + m * ldx + c);
    const float xv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float g = pos[j] ? dv[j] : 0.f;
      const float xh = (xv[j] - mean[c + j]) * rstd[c + j];
      v[j] = gamma[c + j] * rstd[c + j] * (g - dgb[c + j] * invM - xh * dgb[C + c + j] * invM);
    }
  }
  store4_bf16(dx + m * ldd + c, v);
}

// nearest x2 upsample + concat, backward: da[b][y][x][c] = sum_{2x2} dcat[b][2y+i][2x+j][c] (c < C1);
// dskip[b][Y][X][c] = dcat[b][Y][X][C1 + c]
__global__ void __launch_bounds__(256) upcat_bwd_kernel(const float* __restrict__ dcat, float* __restrict__ da,
                                                        float* __restrict__ dskip, int H, int W, int C1, int C2, int64_t n_a,
                                                        int64_t n_s) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const int CT = C1 + C2;
  if (i < n_a) {
    const int c = static_cast<int>(i % C1);
    const int64_t p = i / C1;
    const int x = static_cast<int>(p % W), y = static_cast<int>((p / W) % H);
    const int64_t b = p / (static_cast<int64_t>(W) * H);
    const int64_t base = ((b * 2 * H + 2 * y) * 2 * W + 2 * x) * CT + c;
    da[i] = dcat[base] + dcat[base + CT] + dcat[base + static_cast<int64_t>(2 * W) * CT] +
            dcat[base + static_cast<int64_t>(2 * W) * CT + CT];
  } else if (i < n_a + n_s) {
    const int64_t j = i - n_a;
    const int c = static_cast<int>(j % C2);
    dskip[j] = dcat[(j / C2) * CT + C1 + c];
  }
}

// nn.BatchNorm2d's side effect in training mode: running = (1 - momentum) * running + momentum * batch statistic, with the
// UNBIASED batch variance (var * M / (M - 1)); the biased variance is recovered from rstd = 1 / sqrt(var + eps)
__global__ void __launch_bounds__(256) bn_update_running_kernel(const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                float* __restrict__ running_mean, float* __restrict__ running_var,
                                                                int C, float M, float eps, float momentum) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const float var = 1.0f / (rstd[c] * rstd[c]) - eps;
  running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean[c];
  running_var[c] = (1.0f - momentum) * running_var[c] + momentum * var * (M / (M - 1.0f));
}

static inline unsigned blocks_for(int64_t n) { return static_cast<unsigned>((n + 255) / 256); }

}  // namespace fz

using namespace fz;
#define ST(stream) reinterpret_cast<cudaStream_t>(stream)
typedef const __nv_bfloat16* cbf;
typedef __nv_bfloat16* bf;

// C = 128 * NV rows: the register-resident LayerNorm forward; false = shape not covered (caller takes the scalar kernel)
static bool launch_ln_fwd_vec(const float* x, const float* g, const float* b, bf out, float* out_f32, float* mean, float* rstd,
                              int64_t M, int C, float eps, int out_f16, bf out2, cudaStream_t st) {
  const unsigned grid = static_cast<unsigned>((M + 7) / 8);
  switch (C % 128 == 0 ? C / 128 : 0) {
    case 1: ln_fwd_vec_kernel<1><<<grid, 256, 0, st>>>(x, g, b, out, out_f32, mean, rstd, M, eps, out_f16, out2); return true;
    case 2: ln_fwd_vec_kernel<2><<<grid, 256, 0, st>>>(x, g, b, out, out_f32, mean, rstd, M, eps, out_f16, out2); return true;
    case 3: ln_fwd_vec_kernel<3><<<grid, 256, 0, st>>>(x, g, b, out, out_f32, mean, rstd, M, eps, out_f16, out2); return true;
    case 4: ln_fwd_vec_kernel<4><<<grid, 256, 0, st>>>(x, g, b, out, out_f32, mean, rstd, M, eps, out_f16, out2); return true;
    case 6: ln_fwd_vec_kernel<6><<<grid, 256, 0, st>>>(x, g, b, out, out_f32, mean, rstd, M, eps, out_f16, out2); return true;
    case 8: ln_fwd_vec_kernel<8><<<grid, 256, 0, st>>>(x, g, b, out, out_f32, mean, rstd, M, eps, out_f16, out2); return true;
    default: return false;
  }
}

extern "C" int fz_dwconv7_f32_add(const float* in, const float* w, const float* bias, const float* add, float* out, int B,
                                  int H, int W, int C, int flip, void* stream) {
  FZ_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && in && w && out, "fz_dwconv7_f32: bad arguments");
  FZ_REQUIRE(H <= 65535 && B <= 65535, "fz_dwconv7_f32: H=%d B=%d exceed the grid limits", H, B);
  if (W % 8 == 0 && H % 4 == 0) {
    const int64_t n_tiles = static_cast<int64_t>(B) * (H / 4) * (W / 8);
    FZ_REQUIRE(n_tiles < (1LL << 31) && (n_tiles + 15) / 16 <= 65535, "fz_dwconv7_f32: too many tiles");
    const dim3 grid((C + 31) / 32, static_cast<unsigned>((n_tiles + 15) / 16));
#define FZ_DW_CASE(ct) \
  case ct: dwconv7_f32_tile48_kernel<ct><<<grid, 256, 0, ST(stream)>>>(in, w, bias, add, out, B, H, W, C, flip); break;
    switch (C) {                       // the ConvNeXt-V2 widths (tiny / base / large); anything else takes the runtime form
      FZ_DW_CASE(96) FZ_DW_CASE(128) FZ_DW_CASE(192) FZ_DW_CASE(256) FZ_DW_CASE(384) FZ_DW_CASE(512) FZ_DW_CASE(768)
      FZ_DW_CASE(1024) FZ_DW_CASE(1536)
      default: dwconv7_f32_tile48_kernel<0><<<grid, 256, 0, ST(stream)>>>(in, w, bias, add, out, B, H, W, C, flip);
    }
#undef FZ_DW_CASE
    FZ_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (W % 8 == 0) {
    const int64_t n_tiles = static_cast<int64_t>(B) * H * (W / 8);
    FZ_REQUIRE(n_tiles < (1LL << 31) && (n_tiles + 31) / 32 <= 65535, "fz_dwconv7_f32: too many tiles");
    dwconv7_f32_tiled_kernel<<<dim3((C + 31) / 32, static_cast<unsigned>((n_tiles + 31) / 32)), 256, 0, ST(stream)>>>(
        in, w, bias, out, B, H, W, C, flip);
  } else {
    dwconv7_f32_kernel<<<dim3((W * C + 255) / 256, H, B), 256, 0, ST(stream)>>>(in, w, bias, out, H, W, C, flip);
  }
  FZ_CHECK_CUDA(cudaGetLastError());
  if (add) {                                                   // ragged shapes: the sum as its own pass (out += add)
    const int64_t n = static_cast<int64_t>(B) * H * W * C;
    add_f32_kernel<<<blocks_for(n), 256, 0, ST(stream)>>>(out, add, out, n);
    FZ_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int fz_dwconv7_f32(const float* in, const float* w, const float* bias, float* out, int B, int H, int W, int C,
                              int flip, void* stream) {
  return fz_dwconv7_f32_add(in, w, bias, nullptr, out, B, H, W, C, flip, stream);
}

extern "C" int fz_dwconv7_wgrad(const float* x, const float* du, float* dw, float* db, int B, int H, int W, int C,
                                void* stream) {
  FZ_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && x && du && dw && db, "fz_dwconv7_wgrad: bad arguments");
  // the pixel walk is split into chunks (grid.z) so that every SM has work and no thread chains thousands of dependent
  // global loads; partial[chunk][50][C] lives in a per-device scratch buffer and is summed in a fixed order
  const int64_t npx = static_cast<int64_t>(B) * H * W;
  FZ_REQUIRE(npx < (1LL << 31), "fz_dwconv7_wgrad: B*H*W must fit int32");
  const bool tiled = W % 8 == 0, tile48 = tiled && H % 4 == 0;
  int chunks = tile48 ? static_cast<int>(npx / 32 / 8)                                    // >= 8 tiles (one per warp) per chunk
                      : (tiled ? static_cast<int>(npx / 8 / 8) : static_cast<int>(npx / 2048));
  // the tile kernels hold 2 blocks of 128 registers per SM: size the grid to ONE wave of them (ncu: 1024 blocks with one
  // tile per warp each spent their time in the 50-value block reduction, profiles/r2_ncu_train_bwd_summary.txt)
  const int one_wave = 2 * 148 / ((C + 31) / 32);
  if (tiled && chunks > one_wave) chunks = one_wave;
  chunks = chunks < 1 ? 1 : (chunks > 128 ? 128 : chunks);
  const int ppc = static_cast<int>((npx + chunks - 1) / chunks);
  static float* scratch[64] = {nullptr};
  static size_t scratch_floats[64] = {0};
  int dev = 0;
  FZ_CHECK_CUDA(cudaGetDevice(&dev));
  FZ_REQUIRE(dev >= 0 && dev < 64, "fz_dwconv7_wgrad: device index %d", dev);
  const size_t need = static_cast<size_t>(chunks + 1) * 50 * C;
  if (scratch_floats[dev] < need) {
    if (scratch[dev]) cudaFree(scratch[dev]);
    FZ_CHECK_CUDA(cudaMalloc(&scratch[dev], need * sizeof(float)));
    scratch_floats[dev] = need;
  }
  float* partial = scratch[dev];
  float* sums = partial + static_cast<size_t>(chunks) * 50 * C;
  if (tile48) {
    const int n_tiles = static_cast<int>(npx / 32);
    const dim3 grid((C + 31) / 32, chunks);
    const int tpc = (n_tiles + chunks - 1) / chunks;
#define FZ_DW_CASE(ct) \
  case ct: dwconv7_wgrad_tile48_kernel<ct><<<grid, 256, 0, ST(stream)>>>(x, du, partial, B, H, W, C, tpc); break;
    switch (C) {
      FZ_DW_CASE(96) FZ_DW_CASE(128) FZ_DW_CASE(192) FZ_DW_CASE(256) FZ_DW_CASE(384) FZ_DW_CASE(512) FZ_DW_CASE(768)
      FZ_DW_CASE(1024) FZ_DW_CASE(1536)
      default: dwconv7_wgrad_tile48_kernel<0><<<grid, 256, 0, ST(stream)>>>(x, du, partial, B, H, W, C, tpc);
    }
#undef FZ_DW_CASE
  } else if (tiled) {
    const int n_tiles = static_cast<int>(npx / 8);
    dwconv7_wgrad_tiled_kernel<<<dim3((C + 31) / 32, chunks), 256, 0, ST(stream)>>>(x, du, partial, B, H, W, C,
                                                                                   (n_tiles + chunks - 1) / chunks);
  } else {
    dwconv7_wgrad_kernel<<<dim3(50, (C + 31) / 32, chunks), 256, 0, ST(stream)>>>(x, du, partial, B, H, W, C, ppc);
  }
  reduce_rows_kernel<<<(50 * C + 31) / 32, 256, 0, ST(stream)>>>(partial, sums, 50 * C, chunks);
  FZ_CHECK_CUDA(cudaMemcpyAsync(dw, sums, static_cast<size_t>(49) * C * sizeof(float), cudaMemcpyDeviceToDevice, ST(stream)));
  FZ_CHECK_CUDA(cudaMemcpyAsync(db, sums + static_cast<size_t>(49) * C, C * sizeof(float), cudaMemcpyDeviceToDevice, ST(stream)));
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_layernorm_fwd_stats(const float* x, const float* g, const float* b, void* out_bf16, void* out2_bf16, float* mean,
                                      float* rstd, int64_t M, int C, float eps, int out_f16, void* stream) {
  FZ_REQUIRE(M > 0 && C > 0 && x && g && b && out_bf16 && mean && rstd, "fz_layernorm_fwd_stats: bad arguments");
  if (!launch_ln_fwd_vec(x, g, b, reinterpret_cast<bf>(out_bf16), nullptr, mean, rstd, M, C, eps, out_f16,
                         reinterpret_cast<bf>(out2_bf16), ST(stream)))
    ln_fwd_stats_kernel<<<static_cast<unsigned>((M + 7) / 8), 256, 0, ST(stream)>>>(
        x, g, b, reinterpret_cast<bf>(out_bf16), nullptr, mean, rstd, M, C, eps, out_f16, reinterpret_cast<bf>(out2_bf16));
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_layernorm_bwd(const void* dy_bf16, const float* x, const float* mean, const float* rstd, const float* g,
                                float* dx, float* partial, float* dgamma_dbeta, int64_t M, int C, int blocks, void* stream) {
  FZ_REQUIRE(M > 0 && C > 0 && blocks >= 1 && C <= 2048, "fz_layernorm_bwd: M=%lld C=%d (C <= 2048)", (long long)M, C);
  const int smem = 8 * 2 * C * 4;
#define FZ_LN_CASE(nv)                                                                                          \
  case nv: {                                                                                                    \
    auto kv = ln_bwd_vec_kernel<nv>;                                                                            \
    FZ_ENSURE_SMEM(kv, 8 * 2 * 128 * nv * 4);                                                                   \
    kv<<<blocks, 256, smem, ST(stream)>>>(reinterpret_cast<cbf>(dy_bf16), x, mean, rstd, g, dx, partial, M);    \
  } break;
  switch (C % 128 == 0 ? C / 128 : 0) {
    FZ_LN_CASE(1) FZ_LN_CASE(2) FZ_LN_CASE(3) FZ_LN_CASE(4) FZ_LN_CASE(6) FZ_LN_CASE(8)
    default: {
      auto kern = ln_bwd_kernel;
      FZ_ENSURE_SMEM(kern, 8 * 2 * 2048 * 4);        // opt in once for the largest C (the attribute is set once per kernel)
      kern<<<blocks, 256, smem, ST(stream)>>>(reinterpret_cast<cbf>(dy_bf16), x, mean, rstd, g, dx, partial, M, C);
    }
  }
#undef FZ_LN_CASE
  reduce_rows_kernel<<<(2 * C + 31) / 32, 256, 0, ST(stream)>>>(partial, dgamma_dbeta, 2 * C, blocks);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_gelu_fwd(const void* h_bf16, void* g_bf16, int64_t n, void* stream) {
  FZ_REQUIRE(n > 0 && h_bf16 && g_bf16, "fz_gelu_fwd: bad arguments");
  FZ_REQUIRE(n % 8 == 0, "fz_gelu_fwd: n=%lld must be a multiple of 8", (long long)n);
  gelu_fwd_kernel<<<blocks_for(n / 8), 256, 0, ST(stream)>>>(reinterpret_cast<cbf>(h_bf16), reinterpret_cast<bf>(g_bf16), n);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_sample_colreduce(const void* a_bf16, const void* b_bf16, float* out, int B, int HW, int C, int mode,
                                   void* stream) {
  FZ_REQUIRE(B > 0 && HW > 0 && C > 0 && mode >= 0 && mode <= 2 && a_bf16 && out && (mode != 1 || b_bf16) && B <= 65535,
             "fz_sample_colreduce: bad arguments");
  if (C % 8 == 0 && mode != 1) {
    FZ_CHECK_CUDA(mode == 0 ? rv_colreduce<0>(a_bf16, nullptr, nullptr, out, nullptr, B, HW, C, ST(stream))
                            : rv_colreduce<2>(a_bf16, nullptr, nullptr, out, nullptr, B, HW, C, ST(stream)));
    return 0;
  }
  sample_colreduce_kernel<<<dim3((C + 31) / 32, B), 256, 0, ST(stream)>>>(reinterpret_cast<cbf>(a_bf16),
                                                                          reinterpret_cast<cbf>(b_bf16), out, HW, C, mode);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// s1[b][c] = sum_hw a*b, s0[b][c] = sum_hw a in ONE pass over both tensors (the GRN backward's two reductions)
extern "C" int fz_sample_colreduce2(const void* a_bf16, const void* b_16, float* s1, float* s0, int B, int HW, int C,
                                    int b_f16, void* stream) {
  const void* b_bf16 = b_16;
  FZ_REQUIRE(B > 0 && HW > 0 && C > 0 && C % 8 == 0 && a_bf16 && b_bf16 && s1 && s0 && B <= 65535,
             "fz_sample_colreduce2: bad arguments (C %% 8 == 0)");
  FZ_CHECK_CUDA(rv_colreduce<3>(a_bf16, b_bf16, nullptr, s1, s0, B, HW, C, ST(stream), nullptr, b_f16 ? 2 : 0));
  return 0;
}

// g = GELU(h) (bf16) and sumsq[b][c] = sum_hw g^2 of the stored values, one pass
extern "C" int fz_gelu_fwd_sumsq(const void* h_bf16, void* g_bf16, void* dgelu_bf16, float* sumsq, int B, int HW, int C,
                                 int act_f16, void* stream) {
  FZ_REQUIRE(B > 0 && HW > 0 && C > 0 && C % 8 == 0 && h_bf16 && g_bf16 && sumsq && B <= 65535,
             "fz_gelu_fwd_sumsq: bad arguments (C %% 8 == 0)");
  FZ_CHECK_CUDA(rv_colreduce<1>(h_bf16, nullptr, g_bf16, sumsq, nullptr, B, HW, C, ST(stream), dgelu_bf16, act_f16 ? 5 : 0));
  return 0;
}

extern "C" int fz_grn_train_forward(const void* g_bf16, const float* sumsq, const float* gamma, const float* beta, float* gx,
                                    float* nx, float* mu, void* y_bf16, void* y2_bf16, int B, int HW, int C, float eps,
                                    int act_f16, void* stream) {
  FZ_REQUIRE(B > 0 && HW > 0 && C > 0 && g_bf16 && sumsq && gamma && beta && gx && nx && mu && y_bf16,
             "fz_grn_train_forward: bad arguments");
  grn_norms_kernel<<<B, 256, 0, ST(stream)>>>(sumsq, gx, nx, mu, C, eps);
  const int64_t per = static_cast<int64_t>(HW) * C;
  FZ_REQUIRE(C % 8 == 0 && per < (1LL << 31) && B <= 65535, "fz_grn_train_forward: C %% 8, HW*C < 2^31, B <= 65535");
  const RvGeom geo = rv_geometry(B, HW, C, rv_resident(grn_apply_rows_kernel));
  grn_apply_rows_kernel<<<dim3(geo.slabs, geo.chunks, B), 256, 0, ST(stream)>>>(
      reinterpret_cast<cbf>(g_bf16), nx, gamma, beta, reinterpret_cast<bf>(y_bf16), reinterpret_cast<bf>(y2_bf16), HW, C,
      geo.cgs, geo.rows_per_chunk, act_f16);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int grn_gelu_backward_impl(const void* dy_bf16, const void* g_bf16, const void* h_bf16, const float* s1,
                                  const float* s0, const float* gx, const float* nx, const float* mu, const float* gamma,
                                  float* coef_a, float* coef_b, float* dgamma, float* dbeta, void* dh_bf16, float* dbias,
                                  int B, int HW, int C, float eps, void* stream, int h_is_dgelu, int act_f16) {
  FZ_REQUIRE(B > 0 && HW > 0 && C > 0 && dy_bf16 && g_bf16 && h_bf16 && s1 && s0 && gx && nx && mu && gamma && coef_a &&
                 coef_b && dgamma && dbeta && dh_bf16,
             "fz_grn_gelu_backward: bad arguments");
  grn_bwd_coef_kernel<<<B, 256, 0, ST(stream)>>>(s1, gx, nx, mu, gamma, coef_a, coef_b, C, eps);
  grn_param_grad_kernel<<<(C + 255) / 256, 256, 0, ST(stream)>>>(s1, s0, nx, dgamma, dbeta, B, C);
  const int64_t per = static_cast<int64_t>(HW) * C;
  FZ_REQUIRE(C % 8 == 0 && per < (1LL << 31) && B <= 65535, "fz_grn_gelu_backward: C %% 8, HW*C < 2^31, B <= 65535");
  // dh, and (dbias != NULL) the column sums of dh over all B*HW rows = the bias gradient of the Linear that produced h
  const RvGeom geo = rv_geometry(B, HW, C, rv_resident(grn_gelu_bwd_rows_kernel));
  float* partial = rv_scratch(static_cast<size_t>(B) * geo.chunks * C);
  FZ_REQUIRE(partial != nullptr, "fz_grn_gelu_backward: no scratch memory");
  grn_gelu_bwd_rows_kernel<<<dim3(geo.slabs, geo.chunks, B), 256, 0, ST(stream)>>>(
      reinterpret_cast<cbf>(dy_bf16), reinterpret_cast<cbf>(g_bf16), reinterpret_cast<cbf>(h_bf16), coef_a, coef_b,
      reinterpret_cast<bf>(dh_bf16), partial, HW, C, geo.cgs, geo.rows_per_chunk, h_is_dgelu, act_f16);
  if (dbias) colreduce_final_kernel<<<dim3((C + 31) / 32, 1), 256, 0, ST(stream)>>>(partial, dbias, nullptr, B * geo.chunks, C, 1);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_grn_gelu_backward_db(const void* dy_bf16, const void* g_bf16, const void* h_bf16, const float* s1,
                                       const float* s0, const float* gx, const float* nx, const float* mu, const float* gamma,
                                       float* coef_a, float* coef_b, float* dgamma, float* dbeta, void* dh_bf16, float* dbias,
                                       int B, int HW, int C, float eps, void* stream) {
  return grn_gelu_backward_impl(dy_bf16, g_bf16, h_bf16, s1, s0, gx, nx, mu, gamma, coef_a, coef_b, dgamma, dbeta, dh_bf16,
                                dbias, B, HW, C, eps, stream, 0, 0);
}

extern "C" int fz_grn_gelu_backward_saved(const void* dy_bf16, const void* g_bf16, const void* dgelu_bf16, const float* s1,
                                          const float* s0, const float* gx, const float* nx, const float* mu,
                                          const float* gamma, float* coef_a, float* coef_b, float* dgamma, float* dbeta,
                                          void* dh_bf16, float* dbias, int B, int HW, int C, float eps, int act_f16,
                                          void* stream) {
  return grn_gelu_backward_impl(dy_bf16, g_bf16, dgelu_bf16, s1, s0, gx, nx, mu, gamma, coef_a, coef_b, dgamma, dbeta, dh_bf16,
                                dbias, B, HW, C, eps, stream, 1, act_f16);
}

extern "C" int fz_grn_gelu_backward(const void* dy_bf16, const void* g_bf16, const void* h_bf16, const float* s1,
                                    const float* s0, const float* gx, const float* nx, const float* mu, const float* gamma,
                                    float* coef_a, float* coef_b, float* dgamma, float* dbeta, void* dh_bf16, int B, int HW,
                                    int C, float eps, void* stream) {
  return fz_grn_gelu_backward_db(dy_bf16, g_bf16, h_bf16, s1, s0, gx, nx, mu, gamma, coef_a, coef_b, dgamma, dbeta, dh_bf16,
                                 nullptr, B, HW, C, eps, stream);
}

extern "C" int fz_cast_f16_bf16(const void* in_f16, void* out_bf16, int64_t n, void* stream) {
  FZ_REQUIRE(n >= 0 && (n == 0 || (in_f16 && out_bf16)), "fz_cast_f16_bf16: bad arguments");
  FZ_REQUIRE(((reinterpret_cast<uintptr_t>(in_f16) | reinterpret_cast<uintptr_t>(out_bf16)) & 15) == 0,
             "fz_cast_f16_bf16: 16-byte aligned buffers required");
  if (n == 0) return 0;
  cast_f16_bf16_kernel<<<blocks_for((n + 7) / 8), 256, 0, ST(stream)>>>(reinterpret_cast<const __half*>(in_f16),
                                                                         reinterpret_cast<bf>(out_bf16), n);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_add_f32(const float* a, const float* b, float* out, int64_t n, void* stream) {
  FZ_REQUIRE(n > 0 && a && b && out, "fz_add_f32: bad arguments");
  const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (n % 4 == 0 && aligned)
    add_f32x4_kernel<<<blocks_for(n / 4), 256, 0, ST(stream)>>>(reinterpret_cast<const float4*>(a),
                                                                reinterpret_cast<const float4*>(b),
                                                                reinterpret_cast<float4*>(out), n / 4);
  else
    add_f32_kernel<<<blocks_for(n), 256, 0, ST(stream)>>>(a, b, out, n);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_layernorm_fwd_stats2(const float* x, const float* g, const float* b, void* out_bf16, float* out_f32,
                                       float* mean, float* rstd, int64_t M, int C, float eps, int out_f16, void* stream) {
  FZ_REQUIRE(M > 0 && C > 0 && x && g && b && (out_bf16 || out_f32) && mean && rstd, "fz_layernorm_fwd_stats2: bad arguments");
  if (!launch_ln_fwd_vec(x, g, b, reinterpret_cast<bf>(out_bf16), out_f32, mean, rstd, M, C, eps, out_f16, nullptr, ST(stream)))
    ln_fwd_stats_kernel<<<static_cast<unsigned>((M + 7) / 8), 256, 0, ST(stream)>>>(x, g, b, reinterpret_cast<bf>(out_bf16),
                                                                                   out_f32, mean, rstd, M, C, eps, out_f16, nullptr);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_s2d_bf16(const void* in, void* out, int B, int H, int W, int C, int s, int inverse, void* stream) {
  FZ_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && s >= 1 && H % s == 0 && W % s == 0 && in && out, "fz_s2d_bf16: bad arguments");
  const int64_t n = static_cast<int64_t>(B) * H * W * C;
  if (C % 8 == 0)
    s2d_bf16_vec_kernel<<<blocks_for(n / 8), 256, 0, ST(stream)>>>(reinterpret_cast<const uint4*>(in),
                                                                   reinterpret_cast<uint4*>(out), H, W, C / 8, s, inverse, n / 8);
  else
    s2d_bf16_kernel<<<blocks_for(n), 256, 0, ST(stream)>>>(reinterpret_cast<cbf>(in), reinterpret_cast<bf>(out), H, W, C, s,
                                                           inverse, n);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_patchify4_nchw(const float* in, void* out_bf16, int B, int Cin, int P, int Kpad, int out_f16, void* stream) {
  FZ_REQUIRE(B > 0 && Cin > 0 && P > 0 && P % 4 == 0 && Kpad >= Cin * 16 && in && out_bf16, "fz_patchify4_nchw: bad arguments");
  const int64_t n = static_cast<int64_t>(B) * (P / 4) * (P / 4) * Kpad;
  patchify4_nchw_kernel<<<blocks_for(n), 256, 0, ST(stream)>>>(in, reinterpret_cast<bf>(out_bf16), Cin, P, Kpad, n, out_f16);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_im2col3x3_bf16(const void* in, void* col, int B, int H, int W, int C, int Kpad, void* stream) {
  FZ_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && Kpad >= 9 * C && in && col, "fz_im2col3x3_bf16: bad arguments");
  const int64_t n = static_cast<int64_t>(B) * H * W * Kpad;
  if (C % 8 == 0 && Kpad % 8 == 0)
    im2col3x3_vec_kernel<<<blocks_for(n / 8), 256, 0, ST(stream)>>>(reinterpret_cast<const uint4*>(in),
                                                                    reinterpret_cast<uint4*>(col), H, W, C / 8, Kpad / 8, n / 8);
  else
    im2col3x3_kernel<<<blocks_for(n), 256, 0, ST(stream)>>>(reinterpret_cast<cbf>(in), reinterpret_cast<bf>(col), H, W, C, Kpad, n);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_col2im3x3(const void* dcol_bf16, float* dx, int B, int H, int W, int C, int Kpad, void* stream) {
  FZ_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && Kpad >= 9 * C && dcol_bf16 && dx, "fz_col2im3x3: bad arguments");
  const int64_t n = static_cast<int64_t>(B) * H * W * C;
  if (C % 8 == 0 && Kpad % 8 == 0)
    col2im3x3_vec_kernel<<<blocks_for(n / 8), 256, 0, ST(stream)>>>(reinterpret_cast<cbf>(dcol_bf16), dx, H, W, C, Kpad, n / 8);
  else
    col2im3x3_kernel<<<blocks_for(n), 256, 0, ST(stream)>>>(reinterpret_cast<cbf>(dcol_bf16), dx, H, W, C, Kpad, n);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_bn_relu_train_forward(const float* x, int ldx, const float* gamma, const float* beta, void* y_bf16,
                                        float* mean, float* rstd, float* workspace, int64_t M, int C, int chunks, float eps,
                                        int out_f16, void* stream) {
  FZ_REQUIRE(M > 0 && C > 0 && ldx >= C && chunks >= 1 && chunks <= 65535 && x && gamma && beta && y_bf16 && mean && rstd &&
                 workspace,
             "fz_bn_relu_train_forward: bad arguments");
  cudaStream_t st = ST(stream);
  const int rpc = static_cast<int>((M + chunks - 1) / chunks);
  float* partial = workspace;                                  // [chunks][2][C]
  float* sums = workspace + static_cast<size_t>(chunks) * 2 * C;   // [2][C]
  const bool vec = C % 4 == 0 && ldx % 4 == 0;
  const int groups = C / 4, cgs = vec ? (groups < 256 ? groups : 256) : 1;
  const dim3 grid((C + 31) / 32, chunks), vgrid(vec ? (groups + cgs - 1) / cgs : 1, chunks);
  if (vec) bn_partial_vec_kernel<0><<<vgrid, 256, 0, st>>>(x, ldx, nullptr, nullptr, nullptr, nullptr, partial, M, C, cgs, rpc);
  else bn_partial_kernel<<<grid, 256, 0, st>>>(x, ldx, nullptr, nullptr, nullptr, nullptr, partial, M, C, rpc, 0);
  reduce_rows_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(partial, sums, 2 * C, chunks);
  bn_finalize_kernel<<<(C + 255) / 256, 256, 0, st>>>(sums, mean, C, static_cast<float>(M), eps, 0);
  if (vec) bn_partial_vec_kernel<1><<<vgrid, 256, 0, st>>>(x, ldx, nullptr, nullptr, mean, nullptr, partial, M, C, cgs, rpc);
  else bn_partial_kernel<<<grid, 256, 0, st>>>(x, ldx, nullptr, nullptr, mean, nullptr, partial, M, C, rpc, 1);
  reduce_rows_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(partial, sums, 2 * C, chunks);
  bn_finalize_kernel<<<(C + 255) / 256, 256, 0, st>>>(sums, rstd, C, static_cast<float>(M), eps, 1);
  const int64_t n = M * C;
  if (vec)
    bn_relu_apply_vec_kernel<<<blocks_for(n / 4), 256, 0, st>>>(x, ldx, mean, rstd, gamma, beta, reinterpret_cast<bf>(y_bf16),
                                                                C / 4, n / 4, out_f16);
  else
    bn_relu_apply_kernel<<<blocks_for(n), 256, 0, st>>>(x, ldx, mean, rstd, gamma, beta, reinterpret_cast<bf>(y_bf16), C, n, out_f16);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_bn_relu_backward(const float* x, int ldx, const void* dy_bf16, const void* y_bf16, const float* mean,
                                   const float* rstd, const float* gamma, void* dx_bf16, int ldd, float* dbeta_dgamma,
                                   float* workspace, int64_t M, int C, int chunks, void* stream) {
  FZ_REQUIRE(M > 0 && C > 0 && ldx >= C && ldd >= C && chunks >= 1 && chunks <= 65535 && x && dy_bf16 && y_bf16 && mean && rstd &&
                 gamma && dx_bf16 && dbeta_dgamma && workspace,
             "fz_bn_relu_backward: bad arguments");
  cudaStream_t st = ST(stream);
  const int rpc = static_cast<int>((M + chunks - 1) / chunks);
  const bool vec = C % 4 == 0 && ldx % 4 == 0 && ldd % 4 == 0;
  const int groups = C / 4, cgs = vec ? (groups < 256 ? groups : 256) : 1;
  if (vec)
    bn_partial_vec_kernel<2><<<dim3((groups + cgs - 1) / cgs, chunks), 256, 0, st>>>(
        x, ldx, reinterpret_cast<cbf>(dy_bf16), reinterpret_cast<cbf>(y_bf16), mean, rstd, workspace, M, C, cgs, rpc);
  else
    bn_partial_kernel<<<dim3((C + 31) / 32, chunks), 256, 0, st>>>(x, ldx, reinterpret_cast<cbf>(dy_bf16),
                                                                   reinterpret_cast<cbf>(y_bf16), mean, rstd, workspace, M, C, rpc, 2);
  reduce_rows_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(workspace, dbeta_dgamma, 2 * C, chunks);
  const int64_t n = M * ldd;
  if (vec)
    bn_relu_bwd_apply_vec_kernel<<<blocks_for(n / 4), 256, 0, st>>>(
        x, ldx, reinterpret_cast<cbf>(dy_bf16), reinterpret_cast<cbf>(y_bf16), mean, rstd, gamma, dbeta_dgamma,
        reinterpret_cast<bf>(dx_bf16), ldd, C, 1.0f / static_cast<float>(M), n / 4);
  else
    bn_relu_bwd_apply_kernel<<<blocks_for(n), 256, 0, st>>>(x, ldx, reinterpret_cast<cbf>(dy_bf16), reinterpret_cast<cbf>(y_bf16),
                                                            mean, rstd, gamma, dbeta_dgamma, reinterpret_cast<bf>(dx_bf16), ldd, C,
                                                            1.0f / static_cast<float>(M), n);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_upsample2_concat_backward(const float* dcat, float* da, float* dskip, int B, int H, int W, int C1, int C2,
                                            void* stream) {
  FZ_REQUIRE(B > 0 && H > 0 && W > 0 && C1 > 0 && C2 >= 0 && dcat && da && (C2 == 0 || dskip),
             "fz_upsample2_concat_backward: bad arguments");
  const int64_t n_a = static_cast<int64_t>(B) * H * W * C1, n_s = static_cast<int64_t>(B) * 4 * H * W * C2;
  upcat_bwd_kernel<<<blocks_for(n_a + n_s), 256, 0, ST(stream)>>>(dcat, da, dskip, H, W, C1, C2, n_a, n_s);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_reduce_rows_f32(const float* partial, float* out, int N, int S, void* stream) {
  FZ_REQUIRE(N > 0 && S >= 1 && partial && out, "fz_reduce_rows_f32: bad arguments");
  reduce_rows_kernel<<<(N + 31) / 32, 256, 0, ST(stream)>>>(partial, out, N, S);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_bn_update_running(const float* mean, const float* rstd, float* running_mean, float* running_var, int C,
                                    int64_t M, float eps, float momentum, void* stream) {
  FZ_REQUIRE(C > 0 && M > 1 && mean && rstd && running_mean && running_var, "fz_bn_update_running: bad arguments");
  bn_update_running_kernel<<<(C + 255) / 256, 256, 0, ST(stream)>>>(mean, rstd, running_mean, running_var, C,
                                                                    static_cast<float>(M), eps, momentum);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

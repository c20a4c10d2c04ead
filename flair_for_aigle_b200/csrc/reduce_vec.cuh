// Vectorised column reductions over bf16 [B][rows][C] activations for the training step (C % 8 == 0).
//
// The round-1 reductions gave every thread one channel and 2-byte loads (sample_colreduce_kernel: 0.6 TB/s, 20 ms per step).
// Here a thread owns one 8-channel group (one 16-byte load per row) and the block walks its row chunk `rp` rows at a time,
// so a warp reads 512 contiguous bytes per row and the per-channel coefficients / accumulators stay in registers.  Sums are
// written per (sample, chunk) and added in chunk order by colreduce_final_kernel: the result does not depend on scheduling.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace fz {

// Activation tensors of the training step come in two 16-bit formats: bf16 (gradients: range) and IEEE fp16 (forward
// activations: three more significand bits -- tests/diag/grad_precision_budget.py: the bf16 rounding of the FORWARD tensors
// alone costs the gradients 0.8 % of cosine, fp16 0.1 %).  `f16` selects the format of a tensor; it is warp-uniform.
__device__ __forceinline__ void rv_load8(const void* p, float (&v)[8], bool f16 = false) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (f16) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      v[2 * j] = f.x;
      v[2 * j + 1] = f.y;
    } else {
      v[2 * j] = __uint_as_float(w[j] << 16);
      v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
  }
}
__device__ __forceinline__ uint32_t rv_pack2(float lo, float hi, bool f16) {
  uint32_t r;
  if (f16) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    r = *reinterpret_cast<const uint32_t*>(&h);
  }
  return r;
}
// rounds to the 16-bit format, stores, and leaves the ROUNDED values in v (what a later pass over the stored tensor would read)
__device__ __forceinline__ void rv_store8_round(void* p, float (&v)[8], bool f16 = false) {
  uint4 q;
  uint32_t* w = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    w[j] = rv_pack2(v[2 * j], v[2 * j + 1], f16);
    if (f16) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      v[2 * j] = f.x;
      v[2 * j + 1] = f.y;
    } else {
      v[2 * j] = __uint_as_float(w[j] << 16);
      v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
  }
  *reinterpret_cast<uint4*>(p) = q;
}
// Phi(x) = (1 + erf(x / sqrt 2)) / 2 and phi(x) = exp(-x^2 / 2) / sqrt(2 pi) from ONE exponential and one reciprocal
// (Abramowitz & Stegun 7.1.26: erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z), |error| <= 1.5e-7 -- below
// fp32 resolution of erf near 1 and four orders below the bf16 rounding of the stored result).  erff() + expf() made the
// GELU passes of the training step instruction-bound (2.8 TB/s for a read+write kernel); this form is ~15 instructions.
// The negative tail is formed as poly * e / 2 directly, not as 1 - (1 - poly * e), so it keeps its relative accuracy.
__device__ __forceinline__ void rv_gauss_cdf_pdf(float x, float& cdf, float& pdf) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  const float e = __expf(-z * z);
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float tail = 0.5f * poly * t * e;                      // = (1 - erf(z)) / 2
  cdf = x < 0.f ? tail : 1.0f - tail;
  pdf = 0.3989422804014327f * e;
}
// GELU(x) and GELU'(x) of a PAIR of values with the packed fp32 instructions (mul / fma.rn.f32x2, two IEEE operations per issue
// slot): the same formula as above, about 12 issue slots per element instead of 22 -- the forward GELU pass of the training
// step is instruction-bound (ncu: issue slots 72 %, DRAM 32 %, profiles/r2_ncu_train_fwd_summary.txt).
__device__ __forceinline__ void rv_gelu_pair(float2 x, float2& g, float2& dg) {
  const float2 e_arg = __fmul2_rn(__fmul2_rn(x, x), make_float2(-0.72134752044448170f, -0.72134752044448170f));   // -x^2/2 log2(e)
  float2 e, t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(e_arg.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(e_arg.y));
  t.x = __frcp_rn(fmaf(0.23164189f, fabsf(x.x), 1.0f));            // 0.3275911 / sqrt(2)
  t.y = __frcp_rn(fmaf(0.23164189f, fabsf(x.y), 1.0f));
  float2 poly = __ffma2_rn(t, make_float2(1.061405429f, 1.061405429f), make_float2(-1.453152027f, -1.453152027f));
  poly = __ffma2_rn(t, poly, make_float2(1.421413741f, 1.421413741f));
  poly = __ffma2_rn(t, poly, make_float2(-0.284496736f, -0.284496736f));
  poly = __ffma2_rn(t, poly, make_float2(0.254829592f, 0.254829592f));
  const float2 tail = __fmul2_rn(__fmul2_rn(poly, t), __fmul2_rn(e, make_float2(0.5f, 0.5f)));     // (1 - erf(|x| / sqrt 2)) / 2
  const float2 cdf = make_float2(x.x < 0.f ? tail.x : 1.0f - tail.x, x.y < 0.f ? tail.y : 1.0f - tail.y);
  const float2 pdf = __fmul2_rn(e, make_float2(0.3989422804014327f, 0.3989422804014327f));
  dg = __ffma2_rn(x, pdf, cdf);
  g = __fmul2_rn(x, cdf);
}
__device__ __forceinline__ float rv_gelu(float x) {
  float cdf, pdf;
  rv_gauss_cdf_pdf(x, cdf, pdf);
  return x * cdf;
}
__device__ __forceinline__ float rv_gelu_grad(float x) {
  float cdf, pdf;
  rv_gauss_cdf_pdf(x, cdf, pdf);
  return fmaf(x, pdf, cdf);
}

// Geometry shared by the kernels below and their launchers: `cgs` column groups (of 8 channels) per block, rp = 256 / cgs
// row lanes, grid (slabs, chunks, B).
struct RvGeom {
  int cgs, rp, slabs, chunks, rows_per_chunk;
};
// ctas_per_sm: how many 256-thread blocks of the kernel are resident on one SM (rv_resident).  The grid is sized to ONE wave of
// resident blocks: ncu on the first version (profiles/r2_ncu_train_bwd_summary.txt) showed 592 blocks of a 76-register kernel,
// of which only 3 x 148 were resident -- a second, one-third-full wave.
static inline RvGeom rv_geometry(int B, int rows, int C, int ctas_per_sm = 4, int max_chunks = 1 << 16) {
  RvGeom g;
  const int groups = C / 8;
  g.cgs = groups < 256 ? groups : 256;
  g.rp = 256 / g.cgs;
  g.slabs = (groups + g.cgs - 1) / g.cgs;
  int want = (148 * ctas_per_sm) / (B * g.slabs);                     // one wave of resident blocks over the whole grid
  const int most = rows / (g.rp * 8) > 1 ? rows / (g.rp * 8) : 1;     // at least 8 rows per thread
  want = want < 1 ? 1 : (want > most ? most : want);
  g.chunks = want > max_chunks ? max_chunks : want;
  g.rows_per_chunk = (rows + g.chunks - 1) / g.chunks;
  g.chunks = (rows + g.rows_per_chunk - 1) / g.rows_per_chunk;
  return g;
}

// MODE 0: sum a^2      1: g = GELU(a) stored to gout, GELU'(a) stored to dout (when given), sum g^2 (of the stored,
// rounded g)      2: sum a      3: out0 = sum a*b, out1 = sum a
// partial[((b * chunks + chunk) * NOUT + o) * C + c]
// fmt: bit 0 = `a` is fp16, bit 1 = `bb` is fp16, bit 2 = the stored outputs (gout, dout) are fp16; 0 = all bf16
template <int MODE>
__global__ void __launch_bounds__(256) colreduce_vec_kernel(const __nv_bfloat16* __restrict__ a,
                                                            const __nv_bfloat16* __restrict__ bb,
                                                            __nv_bfloat16* __restrict__ gout, __nv_bfloat16* __restrict__ dout,
                                                            float* __restrict__ partial, int rows, int C, int cgs,
                                                            int rows_per_chunk, int fmt) {
  const bool a16 = fmt & 1, b16 = fmt & 2, o16 = fmt & 4;
  constexpr int NOUT = MODE == 3 ? 2 : 1;
  __shared__ float red[NOUT][256][9];                          // 9: the 8-float rows would collide 4-way on the banks
  const int rp = 256 / cgs;
  const int cl = threadIdx.x % cgs, rr = threadIdx.x / cgs;
  const int cg = blockIdx.x * cgs + cl;
  const bool active = rr < rp && cg * 8 < C;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  const size_t base = static_cast<size_t>(blockIdx.z) * rows * C + static_cast<size_t>(cg) * 8;
  float acc0[8], acc1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc0[j] = acc1[j] = 0.f;
  if (active) {
#pragma unroll 4
    for (int r = r0 + rr; r < r1; r += rp) {
      const size_t o = base + static_cast<size_t>(r) * C;
      float v[8];
      rv_load8(a + o, v, a16);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc0[j] = fmaf(v[j], v[j], acc0[j]);
      } else if (MODE == 1) {
        float d[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {                            // GELU and GELU' (the backward reads it instead of recomputing)
          float2 g2, d2;
          rv_gelu_pair(make_float2(v[2 * j], v[2 * j + 1]), g2, d2);
          v[2 * j] = g2.x;
          v[2 * j + 1] = g2.y;
          d[2 * j] = d2.x;
          d[2 * j + 1] = d2.y;
        }
        if (dout) rv_store8_round(dout + o, d, o16);
        rv_store8_round(gout + o, v, o16);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc0[j] = fmaf(v[j], v[j], acc0[j]);
      } else if (MODE == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc0[j] += v[j];
      } else {
        float w[8];
        rv_load8(bb + o, w, b16);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc0[j] = fmaf(v[j], w[j], acc0[j]);
          acc1[j] += v[j];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][threadIdx.x][j] = acc0[j];
    if (NOUT == 2) red[NOUT - 1][threadIdx.x][j] = acc1[j];
  }
  __syncthreads();
  if (rr == 0 && cg * 8 < C) {
    float* dst = partial + (static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * NOUT * C + static_cast<size_t>(cg) * 8;
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = 0.f;
        for (int q = 0; q < rp; ++q) t += red[o][q * cgs + cl][j];        // row lanes in a fixed order
        dst[static_cast<size_t>(o) * C + j] = t;
      }
  }
}

// out_o[b][c] = sum_chunk partial[b][chunk][o][c].  32 channels x 8 chunk lanes per block: lane k adds chunks k, k+8, ... in
// order, then the eight lane sums are added in lane order (fixed: reproducible).   grid (ceil(C / 32), B)
static __global__ void __launch_bounds__(256) colreduce_final_kernel(const float* __restrict__ partial, float* __restrict__ out0,
                                                                     float* __restrict__ out1, int chunks, int C, int nout) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), k = threadIdx.x >> 5;
  const size_t b = blockIdx.y;
  for (int o = 0; o < nout; ++o) {
    float t = 0.f;
    if (c < C)
      for (int s = k; s < chunks; s += 8) t += partial[((b * chunks + s) * nout + o) * C + c];
    red[k][threadIdx.x & 31] = t;
    __syncthreads();
    if (k == 0 && c < C) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) v += red[j][threadIdx.x];
      (o ? out1 : out0)[b * C + c] = v;
    }
    __syncthreads();
  }
}

// resident 256-thread blocks per SM of a kernel (register / shared-memory limited), asked once per kernel
template <typename K>
static inline int rv_resident(K kernel) {
  static int n = 0;
  if (!n) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, 256, 0) != cudaSuccess || n < 1) n = 1;
    if (n > 8) n = 8;
  }
  return n;
}

// Per-device scratch for the partial sums (the training step runs its kernels on one stream).
static inline float* rv_scratch(size_t floats) {
  static float* buf[64] = {nullptr};
  static size_t have[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (have[dev] < floats) {
    if (buf[dev]) cudaFree(buf[dev]);
    buf[dev] = nullptr;
    have[dev] = 0;
    if (cudaMalloc(&buf[dev], floats * sizeof(float)) != cudaSuccess) return nullptr;
    have[dev] = floats;
  }
  return buf[dev];
}

// Launches MODE over a [B][rows][C] tensor; returns a cudaError_t.
template <int MODE>
static inline cudaError_t rv_colreduce(const void* a, const void* bb, void* gout, float* out0, float* out1, int B, int rows,
                                       int C, cudaStream_t st, void* dout = nullptr, int fmt = 0) {
  constexpr int NOUT = MODE == 3 ? 2 : 1;
  const RvGeom g = rv_geometry(B, rows, C, rv_resident(colreduce_vec_kernel<MODE>));
  float* partial = rv_scratch(static_cast<size_t>(B) * g.chunks * NOUT * C);
  if (!partial) return cudaErrorMemoryAllocation;
  colreduce_vec_kernel<MODE><<<dim3(g.slabs, g.chunks, B), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(bb),
      reinterpret_cast<__nv_bfloat16*>(gout), reinterpret_cast<__nv_bfloat16*>(dout), partial, rows, C, g.cgs, g.rows_per_chunk, fmt);
  colreduce_final_kernel<<<dim3((C + 31) / 32, B), 256, 0, st>>>(partial, out0, out1, g.chunks, C, NOUT);
  return cudaGetLastError();
}

}  // namespace fz

// Epilogue shared by the one-SM (gemm_tcgen05.cu) and CTA-pair (gemm_tcgen05_2sm.cu) bf16 GEMM kernels:
// one warp moves a 32-row x 128-byte chunk of the accumulator TMEM -> registers -> (+bias, activation,
// residual) -> XOR-swizzled per-warp staging tile -> global memory in full 128 B lines.
#pragma once
#include "ptx.cuh"
#include "../../include/flair_zonal_b200.h"

namespace fz {

struct GemmParams {
  int M, N, K;
  int rows_per_sample;  // rows of A per sample (H*W); selects the B batch
  int b_batched;        // 1: B is [num_samples][N][K] and the tile uses batch m0 / rows_per_sample
  const float* bias;    // [N], never null
  void* out;            // [M,N] bf16 or f32
  const float* resid;   // [M,N] f32 (FZ_EPI_RESID_F32), may alias out
  float* sumsq;         // [ceil(M/128), N] f32 per-128-row partial sums of out^2 (FZ_EPI_GELU_SUMSQ)
  int splits;           // split-K: the K range is cut into `splits` pieces of kb_per_split k-blocks; piece s writes its fp32
  int kb_per_split;     //   partial tile to out + s * M * N (fz_gemm_bf16_splitk sums the pieces in order); 1 = off
  int nobias;           // 1: the epilogue adds no bias (split-K partials)
  int tma_store;        // 1: the staged 32 x 128 B chunk leaves through a TMA store (CTA-pair kernel; tmO is valid)
  int f16;              // 1: A, B and a 16-bit output are fp16 (FZ_EPI_OPERANDS_F16), 0: bf16
  int reverse;          // 1: walk the tile list backwards (consume a just-written operand newest-first, while it is in L2)
  unsigned long long* trace;  // optional: CTA 0 writes clock64 stamps [tile][8] (diagnostics, see fz_gemm_set_trace)
};

template <int MODE>
struct EpiShape {
  static constexpr bool F32OUT = (MODE == FZ_EPI_RESID_F32 || MODE == FZ_EPI_F32);
  static constexpr int CH_COLS = F32OUT ? 32 : 64;   // 128 B of output per row
  static constexpr int ESZ = F32OUT ? 4 : 2;
};

// taddr : TMEM address of the chunk's first column at the warp's lane base (lane quarter already applied)
// row0  : global output row of the warp's first row (its 32 rows are row0 .. row0+31)
// col0  : global output column of the chunk
// stg   : this warp's 4 KB staging tile; 16 B segments XOR-swizzled with row&7 so that both the row-per-lane and
//         the row-contiguous access patterns are bank-conflict free.  A row-per-thread STG/LDG would touch 32
//         lines per instruction, which was the first epilogue's bottleneck.
// sq_dst: (GELU_SUMSQ) 64 floats: 32-row column sums of out^2 for the chunk's columns (8-byte aligned)
template <int MODE, bool F16>
__device__ __forceinline__ void epi_chunk(const GemmParams& p, uint32_t taddr, int row0, int col0, char* stg, int lane,
                                          float* sq_dst, void* out_base, const CUtensorMap* tmO = nullptr) {
  constexpr bool F32OUT = EpiShape<MODE>::F32OUT;
  constexpr int CH_COLS = EpiShape<MODE>::CH_COLS;
  constexpr int ESZ = EpiShape<MODE>::ESZ;
  const int rsub = lane >> 3, seg = lane & 7;              // row-contiguous mapping: 4 rows x 8 segments
  const size_t row_bytes = static_cast<size_t>(p.N) * ESZ;
  char* gout = reinterpret_cast<char*>(out_base) + static_cast<size_t>(row0) * row_bytes + static_cast<size_t>(col0) * ESZ;
  if (MODE == FZ_EPI_RESID_F32) {
    if (tmO != nullptr) {          // the previous chunk's TMA store has read this staging tile
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
    }
    // coalesced residual tile -> staging
    const char* gres = reinterpret_cast<const char*>(p.resid) + static_cast<size_t>(row0) * row_bytes +
                       static_cast<size_t>(col0) * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = i * 4 + rsub;
      uint4 x = make_uint4(0, 0, 0, 0);
      if (row0 + rr < p.M) x = *reinterpret_cast<const uint4*>(gres + static_cast<size_t>(rr) * row_bytes + seg * 16);
      *reinterpret_cast<uint4*>(stg + rr * 128 + ((seg ^ (rr & 7)) << 4)) = x;
    }
    __syncwarp();
  }
#pragma unroll
  for (int h = 0; h < CH_COLS / 32; ++h) {
    uint32_t r[32];
    tmem_ld32(taddr + h * 32, r);
    const float4* bp = reinterpret_cast<const float4*>(p.bias + col0 + h * 32);
    float4 b4[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b4[j] = p.nobias ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(bp + j);
    float v[32];
    tmem_ld_wait();
    // packed fp32 adds / FMAs (add / fma.rn.f32x2: two IEEE operations per issue slot, same bits as the scalar forms) wherever
    // the epilogue works on column pairs: the epilogue, not the MMA, bounds the K = 512 GEMMs
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 lo = __fadd2_rn(make_float2(__uint_as_float(r[4 * j + 0]), __uint_as_float(r[4 * j + 1])),
                                   make_float2(b4[j].x, b4[j].y));
      const float2 hi = __fadd2_rn(make_float2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])),
                                   make_float2(b4[j].z, b4[j].w));
      v[4 * j + 0] = lo.x;
      v[4 * j + 1] = lo.y;
      v[4 * j + 2] = hi.x;
      v[4 * j + 3] = hi.y;
    }
    if (MODE == FZ_EPI_GELU_SUMSQ || MODE == FZ_EPI_GELU_BF16) {
#ifdef FZ_GELU_SCALAR                                      // A/B build: one column per instruction (round 2 before the packed form)
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
#else
#pragma unroll
      for (int j = 0; j < 16; ++j) {                       // two columns per packed fp32 instruction (same bits as scalar)
        const float2 g2 = gelu_erf_fast2(make_float2(v[2 * j], v[2 * j + 1]));
        v[2 * j] = g2.x;
        v[2 * j + 1] = g2.y;
      }
#endif
    } else if (MODE == FZ_EPI_RELU_BF16) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
    } else if (MODE == FZ_EPI_RESID_F32) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 x = *reinterpret_cast<const float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4));
        const float2 lo = __fadd2_rn(make_float2(v[4 * j + 0], v[4 * j + 1]), make_float2(x.x, x.y));
        const float2 hi = __fadd2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(x.z, x.w));
        v[4 * j + 0] = lo.x;
        v[4 * j + 1] = lo.y;
        v[4 * j + 2] = hi.x;
        v[4 * j + 3] = hi.y;
      }
      __syncwarp();   // everyone has read its residual row before the tile is overwritten
    }
    // own row -> staging
    if (tmO != nullptr && MODE != FZ_EPI_RESID_F32 && h == 0) {   // the previous chunk's TMA store has read this staging tile
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
    }
    if (F32OUT) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(stg + lane * 128 + (((h * 4 + j) ^ (lane & 7)) << 4)) =
            make_uint4(pack16<F16>(v[8 * j], v[8 * j + 1]), pack16<F16>(v[8 * j + 2], v[8 * j + 3]),
                       pack16<F16>(v[8 * j + 4], v[8 * j + 5]), pack16<F16>(v[8 * j + 6], v[8 * j + 7]));
    }
  }
  __syncwarp();
  if (MODE == FZ_EPI_GELU_SUMSQ) {
    // GRN statistics: column sums of out^2 over this warp's 32 rows, read back from the staged 16-bit tile (the values
    // fc2 will actually consume): lane l owns columns 2l, 2l+1 = one 32-bit word per row, conflict-free under the
    // XOR swizzle.  2.5 instructions per element instead of 4.9 for the register transpose-reduce it replaces.
    // M is a multiple of 128 in this mode (host check): no row mask.
    float2 acc = make_float2(0.f, 0.f);
    const int sidx = lane >> 2, woff = (lane & 3) * 4;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(stg + r * 128 + ((sidx ^ (r & 7)) << 4) + woff);
      const float2 f = unpack16<F16>(w);
      acc = __ffma2_rn(f, f, acc);                          // both columns' sums in one packed FMA
    }
    *reinterpret_cast<float2*>(sq_dst + 2 * lane) = acc;
  }
  if (tmO != nullptr) {
    // staging -> global as ONE TMA store of the 32-row x 128-byte box: the staging tile's XOR pattern (16-byte segment ^
    // (row & 7), 4 KB-aligned) IS the 128-byte TMA swizzle; rows beyond M are clipped by the hardware.  Replaces 8 LDS.128 +
    // 8 STG.128 per thread and chunk in the issue-bound epilogue.
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmO, stg, col0, row0);
      tma_store_commit();
    }
    return;       // the staging tile is released by the wait at its next use
  }
  // staging -> global, 4 full 128 B lines per instruction
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + rsub;
    const uint4 x = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((seg ^ (rr & 7)) << 4));
    if (row0 + rr < p.M) *reinterpret_cast<uint4*>(gout + static_cast<size_t>(rr) * row_bytes + seg * 16) = x;
  }
  __syncwarp();       // staging is reused by the next chunk
}

// launchers of the two kernels (defined in their .cu files)
int gemm_pair_launch(const void* A, const void* B, const GemmParams& p, int b_batch, int mode, cudaStream_t stream);

}  // namespace fz

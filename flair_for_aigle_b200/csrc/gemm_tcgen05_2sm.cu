// CTA-pair (tcgen05 cta_group::2) variant of the bf16 GEMM in gemm_tcgen05.cu: same contract, same epilogues.
//
// Why: the one-SM kernel's 256x128 tile pulls 48 KB per k-block through the L2->SM fabric and measured exactly
// that fabric's cap (~6.3 KB/clk over 148 SMs = 1.0 PFLOP/s with epilogues off).  A pair of CTAs on neighbouring
// SMs computes one 256x256 tile: each CTA loads 128 rows of A and 128 of the 256 rows of B (32 KB per k-block),
// the leader's UMMA (M=256, N=256) reads both CTAs' shared memory, and each CTA ends up with its 128 rows x 256
// columns of the accumulator in its own TMEM (2 x 256 columns: the epilogue still overlaps the next tile).
// Operand traffic per flop drops by 1.5x against the 256x128 tile.
//
// Protocol (per pair; "leader" = cluster rank 0):
//   full[s]   (leader)      1 arrival: leader's arrive.expect_tx(64 KB); both CTAs' TMA loads complete_tx on it
//   empty[s]  (each CTA)    1 arrival: leader's tcgen05.commit multicast to both CTAs
//   tfull[a]  (each CTA)    1 arrival: leader's tcgen05.commit multicast at the end of a tile
//   tempty[a] (leader)      32 arrivals: 16 epilogue warps of each CTA (the peer's arrive remotely)
#include "common.h"
#include "ptx.cuh"
#include "gemm_epilogue.cuh"

namespace fz {

namespace pair {
constexpr int BM = 256, BN = 256, BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = 128 * BK * 2;           // this CTA's 128 rows of A
constexpr int B_BYTES = 128 * BK * 2;           // this CTA's 128 rows of B
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int OFF_SQ = STAGES * STAGE_BYTES;    // [tile parity][lane quarter][256] f32
constexpr int OFF_STG = OFF_SQ + 2 * 4 * BN * 4;
constexpr int OFF_BAR = OFF_STG + 16 * 4096;
constexpr int OFF_TSLOT = OFF_BAR + (2 * STAGES + 4) * 8;
constexpr int SMEM_BYTES = OFF_TSLOT + 16 + 1024;
constexpr int THREADS = 640;
constexpr int EPI_WARP0 = 4;
}  // namespace pair

template <int MODE, bool F16>
__global__ void __launch_bounds__(pair::THREADS, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmO, GemmParams p) {
  using namespace pair;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sSq = reinterpret_cast<float*>(smem + OFF_SQ);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;     // [2]
  uint64_t* tempty = tfull + 2;         // [2]
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + OFF_TSLOT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int n_tiles = p.N / BN;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int total_tiles = n_tiles * m_tiles;
  const int num_kb = p.K / BK;
  const int t_first = blockIdx.x >> 1, t_step = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 32);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair(tslot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers and TMEM exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = t_first; t < total_tiles; t += t_step) {
        const int te = p.reverse ? total_tiles - 1 - t : t;
        const int n0 = (te % n_tiles) * BN;
        const int m0 = (te / n_tiles) * BM;
        const int bcoord = p.b_batched ? (m0 / p.rows_per_sample) : 0;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * STAGE_BYTES);
          const uint32_t lead_full = mapa_u32(&full[s], 0);
          uint8_t* st = smem + s * STAGE_BYTES;
          tma_load_2d_pair(&tmA, lead_full, st, kb * BK, m0 + rank * 128);
          tma_load_3d_pair(&tmB, lead_full, st + A_BYTES, kb * BK, n0 + rank * 128, bcoord);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc16(BM, BN, F16);
      uint32_t it = 0, lt = 0;
      for (int t = t_first; t < total_tiles; t += t_step, ++lt) {
        const uint32_t as = lt & 1;
        const uint32_t aph = (lt >> 1) & 1;
        mbar_wait(&tempty[as], aph ^ 1);           // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t acc = tmem + as * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t ad = umma_smem_desc(smem_u32(smem + s * STAGE_BYTES), 128);
          const uint64_t bd = umma_smem_desc(smem_u32(smem + s * STAGE_BYTES + A_BYTES), 128);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_pair(acc, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair(&empty[s], 0x3);
        }
        umma_commit_pair(&tfull[as], 0x3);
      }
    }
    __syncwarp();
  } else if (warp >= EPI_WARP0) {
    const int ew = warp - EPI_WARP0;            // 0..15
    const int q = warp & 3;                     // TMEM lane quarter this warp may read
    const int colgrp = ew >> 2;                 // chunks colgrp, colgrp+4, ...
    constexpr int CH_COLS = EpiShape<MODE>::CH_COLS;
    char* stg = reinterpret_cast<char*>(smem + OFF_STG) + ew * 4096;
    uint32_t lt = 0;
    for (int t = t_first; t < total_tiles; t += t_step, ++lt) {
      const int te = p.reverse ? total_tiles - 1 - t : t;
      const int n0 = (te % n_tiles) * BN;
      const int m0 = (te / n_tiles) * BM + rank * 128;
      const uint32_t as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      float* sq_buf = sSq + (lt & 1) * 4 * BN;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const uint32_t tbase = tmem + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = colgrp; c < BN / CH_COLS; c += 4)
        epi_chunk<MODE, F16>(p, tbase + c * CH_COLS, m0 + q * 32, n0 + c * CH_COLS, stg, lane,
                        sq_buf + q * BN + c * CH_COLS, p.out, p.tma_store ? &tmO : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty[as], 0));
      if (MODE == FZ_EPI_GELU_SUMSQ) {
        // deterministic: fixed-order sum of the four lane-quarter warps (same order as the one-SM kernel).  Only the four
        // warps that share this column group meet at the barrier (round 1 synchronised all 16 epilogue warps per tile:
        // `barrier` was the largest stall reason of the fc1 launch, 1.4 warps per issue-active cycle).
        asm volatile("bar.sync %0, 128;" ::"r"(1 + colgrp) : "memory");
        if (m0 < p.M && q < CH_COLS / 32)
          for (int c = colgrp; c < BN / CH_COLS; c += 4) {
            const int i = c * CH_COLS + q * 32 + lane;
            p.sumsq[static_cast<size_t>(m0 / 128) * p.N + n0 + i] =
                (sq_buf[i] + sq_buf[BN + i]) + (sq_buf[2 * BN + i] + sq_buf[3 * BN + i]);
          }
      }
    }
  }
  if (p.tma_store && warp >= EPI_WARP0 && lane == 0) tma_store_wait_all();   // this warp's last TMA stores have landed
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's smem / barriers / TMEM stay alive until both CTAs are done
  if (warp == 1) tmem_dealloc_pair(tmem, 512);
}

template <int MODE, bool F16>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const GemmParams& p,
                       cudaStream_t stream) {
  auto kern = gemm_bf16_pair_kernel<MODE, F16>;
  FZ_ENSURE_SMEM(kern, pair::SMEM_BYTES);
  const int sm_count = device_sm_count();
  if (sm_count <= 0) return -2;
  const int tiles = ((p.M + pair::BM - 1) / pair::BM) * (p.N / pair::BN);
  int grid = 2 * tiles < sm_count ? 2 * tiles : (sm_count & ~1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(pair::THREADS);
  cfg.dynamicSmemBytes = pair::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FZ_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmO, p));
  return 0;
}

// Caller (fz_gemm_bf16) has validated the contract; here N % 256 == 0 and K % 64 == 0.
int gemm_pair_launch(const void* A, const void* B, const GemmParams& p, int b_batch, int mode, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.M};
    const uint64_t strides[1] = {(uint64_t)p.K * 2};
    const uint32_t box[2] = {pair::BK, 128};
    int rc = make_tmap16(&tmA, A, 2, dims, strides, box, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)p.K, (uint64_t)p.N, (uint64_t)b_batch};
    const uint64_t strides[2] = {(uint64_t)p.K * 2, (uint64_t)p.K * 2 * (uint64_t)p.N};
    const uint32_t box[3] = {pair::BK, 128, 1};
    int rc = make_tmap16(&tmB, B, 3, dims, strides, box, 128);
    if (rc) return rc;
  }
  // output map for the epilogue's TMA stores: box = 32 rows x 128 bytes (one staged chunk), 128-byte swizzle.
  // FZ_GEMM_TMA_STORE=0 keeps the round-1 register -> smem -> st.global path (A/B measurements).
  GemmParams q = p;
  static int tma_store = -1;
  if (tma_store < 0) {
    const char* e = getenv("FZ_GEMM_TMA_STORE");
    tma_store = (e && e[0] == '0') ? 0 : 1;
  }
  const bool f32out = mode == FZ_EPI_RESID_F32 || mode == FZ_EPI_F32;
  CUtensorMap tmO = tmA;     // placeholder when unused
  q.tma_store = 0;
  if (tma_store && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
    const uint64_t dims[2] = {(uint64_t)p.N, (uint64_t)p.M};
    const uint64_t strides[1] = {(uint64_t)p.N * (f32out ? 4 : 2)};
    const uint32_t box[2] = {f32out ? 32u : 64u, 32u};
    int rc = f32out ? make_tmap32(&tmO, p.out, 2, dims, strides, box, 128) : make_tmap16(&tmO, p.out, 2, dims, strides, box, 128);
    if (rc) return rc;
    q.tma_store = 1;
  }
  switch (mode) {
    case FZ_EPI_BF16: return q.f16 ? launch_pair<FZ_EPI_BF16, true>(tmA, tmB, tmO, q, stream) : launch_pair<FZ_EPI_BF16, false>(tmA, tmB, tmO, q, stream);
    case FZ_EPI_GELU_SUMSQ: return q.f16 ? launch_pair<FZ_EPI_GELU_SUMSQ, true>(tmA, tmB, tmO, q, stream) : launch_pair<FZ_EPI_GELU_SUMSQ, false>(tmA, tmB, tmO, q, stream);
    case FZ_EPI_RESID_F32: return q.f16 ? launch_pair<FZ_EPI_RESID_F32, true>(tmA, tmB, tmO, q, stream) : launch_pair<FZ_EPI_RESID_F32, false>(tmA, tmB, tmO, q, stream);
    case FZ_EPI_F32: return q.f16 ? launch_pair<FZ_EPI_F32, true>(tmA, tmB, tmO, q, stream) : launch_pair<FZ_EPI_F32, false>(tmA, tmB, tmO, q, stream);
    case FZ_EPI_RELU_BF16: return q.f16 ? launch_pair<FZ_EPI_RELU_BF16, true>(tmA, tmB, tmO, q, stream) : launch_pair<FZ_EPI_RELU_BF16, false>(tmA, tmB, tmO, q, stream);
    case FZ_EPI_GELU_BF16: return q.f16 ? launch_pair<FZ_EPI_GELU_BF16, true>(tmA, tmB, tmO, q, stream) : launch_pair<FZ_EPI_GELU_BF16, false>(tmA, tmB, tmO, q, stream);
  }
  set_error("fz_gemm_bf16: unknown epilogue mode %d", mode);
  return -1;
}

}  // namespace fz

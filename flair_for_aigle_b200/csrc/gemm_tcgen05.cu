// bf16 GEMM on tcgen05 / TMEM fed by TMA, with the fused epilogues the ConvNeXt-V2 / U-Net
// forward needs.  C[m,n] = sum_k A[m,k] * B[b(m)][n,k]   (both operands K-major, i.e. the
// activations as NHWC rows and nn.Linear / 1x1-conv weights as stored).
//
// Replaces the cuBLAS/cuDNN fp32 calls under flair_hub/models/flair_model.py:376 (timm
// ConvNeXtBlock mlp.fc1 / mlp.fc2, stage downsample conv2x2) and :539-541 (FusionHandler 1x1).
//
// Persistent kernel, one CTA per SM, static tile schedule (tile = blockIdx.x + i*gridDim.x,
// n fastest so CTAs running together share the A rows in L2).  A CTA tile is 256 x BN: two
// M=128 UMMA accumulators that share one B tile, which raises the operand intensity to
// 256*BN*64*2 / ((256+BN)*128) flop per smem byte (87 at BN=128) -- the 128x128 tile of the first version measured
// exactly its L2->SM bandwidth bound (64 flop/B * 8.6 TB/s).  Outputs with N % 256 == 0 go to the CTA-pair kernel
// (gemm_tcgen05_2sm.cu), which gets 131 flop/B out of a 256x256 tile split over two SMs.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM owner), warps 4..19 epilogue:
// 16 warps = 2 row halves x 4 TMEM lane quarters (= warp % 4) x 2 interleaved column groups, i.e.
// four epilogue warps per SM sub-partition to hide the TMEM-load / SFU / global latencies.
// smem ring of STAGES k-blocks runs continuously across tiles; with BN=128 the accumulators
// are double buffered in TMEM (2 x 256 columns) so tile i's epilogue overlaps tile i+1's MMAs.
#include "common.h"
#include "ptx.cuh"
#include "gemm_epilogue.cuh"
#include "../../include/flair_zonal_b200.h"

namespace fz {

constexpr int BM = 256;              // rows per CTA tile (two UMMA M=128 halves)
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int GEMM_THREADS = 640;   // 4 control warps + 16 epilogue warps
constexpr int EPI_WARP0 = 4;

static unsigned long long* g_trace = nullptr;
#define FZ_TRACE(slot)                                                          \
  do {                                                                          \
    if (p.trace && blockIdx.x == 0 && lt < 64) p.trace[lt * 8 + (slot)] = clock64(); \
  } while (0)

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int OFF_B = STAGES * A_STAGE_BYTES;
  static constexpr int OFF_SQ = OFF_B + STAGES * B_STAGE_BYTES;
  static constexpr int OFF_STG = OFF_SQ + 2 * 8 * BN * 4;       // [tile parity][half*4+quarter][BN] sumsq rows
  static constexpr int OFF_BAR = OFF_STG + 16 * 4096;           // 4 KB store-staging tile per epilogue warp
  static constexpr int OFF_TSLOT = OFF_BAR + (2 * STAGES + 4) * 8;
  static constexpr int BYTES = OFF_TSLOT + 16 + 1024;            // + worst-case alignment pad
};

// MNMAJOR: both operands are given transposed -- A as [K][M], B as [K][N], row-major (the weight-gradient GEMM dW = dY^T X reads
// dY [rows][N_l] and X [rows][K_l] as they are, no transposed copies).  Each 64-wide block of M / N is its own TMA box.
template <int BN, int STAGES, int ACC_STAGES, int MODE, bool F16, bool MNMAJOR = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, GemmParams p) {
  using L = GemmSmem<BN, STAGES>;
  static_assert(ACC_STAGES * 2 * BN <= 512, "TMEM has 512 columns");
  constexpr int TCOLS = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + L::OFF_B;
  float* sSq = reinterpret_cast<float*>(smem + L::OFF_SQ);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;     // [2]
  uint64_t* tempty = tfull + 2;         // [2]
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + L::OFF_TSLOT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = p.N / BN;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int total_tiles = n_tiles * m_tiles;
  const int total_work = total_tiles * p.splits;      // split-K (p.splits > 1): work item = (tile, K piece)
  const int num_kb = p.K / BK;
  // persistent schedule: tile t -> (m = t / n_tiles, n = t % n_tiles)
  const int t_first = blockIdx.x, t_step = gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 16);       // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tslot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, lt = 0;
      for (int t = t_first; t < total_work; t += t_step, ++lt) {
        const int tile = t / p.splits, sp = t - tile * p.splits;
        const int te = p.reverse ? total_tiles - 1 - tile : tile;
        const int n0 = (te % n_tiles) * BN;
        const int m0 = (te / n_tiles) * BM;
        const int bcoord = p.b_batched ? (m0 / p.rows_per_sample) : 0;
        const int kb0 = sp * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        FZ_TRACE(0);   // producer starts issuing this tile's loads
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], A_STAGE_BYTES + L::B_STAGE_BYTES);
          if (MNMAJOR) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(&tmA, &full[s], sA + s * A_STAGE_BYTES + j * (BK * 128), m0 + j * 64, kb * BK);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(&tmB, &full[s], sB + s * L::B_STAGE_BYTES + j * (BK * 128), n0 + j * 64, kb * BK);
          } else {
            tma_load_2d(&tmA, &full[s], sA + s * A_STAGE_BYTES, kb * BK, m0);
            tma_load_3d(&tmB, &full[s], sB + s * L::B_STAGE_BYTES, kb * BK, n0, bcoord);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // (a_format = BF16 with b_format = F16 in one descriptor was tried for the training weight gradients: the MMA is an
      //  illegal instruction on sm_100a, kind::f16 wants one format for both operands)
      constexpr uint32_t idesc = MNMAJOR ? umma_idesc16_mn(128, BN, F16) : umma_idesc16(128, BN, F16);
      uint32_t it = 0, lt = 0;
      for (int t = t_first; t < total_work; t += t_step, ++lt) {
        const int sp = t % p.splits;
        const int kb0 = sp * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        const uint32_t as = lt % ACC_STAGES;
        const uint32_t aph = (lt / ACC_STAGES) & 1;
        FZ_TRACE(1);   // MMA warp reaches the tile
        mbar_wait(&tempty[as], aph ^ 1);           // epilogue has drained this accumulator stage
        tc_fence_after();
        FZ_TRACE(2);   // accumulator stage free
        const uint32_t acc0 = tmem + as * (2 * BN);
        const uint32_t acc1 = acc0 + BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (kb == kb0) FZ_TRACE(3);   // first k-block landed
          const uint32_t a_addr = smem_u32(sA + s * A_STAGE_BYTES), b_addr = smem_u32(sB + s * L::B_STAGE_BYTES);
          // K-major: rows of 128 B, +32 bytes (16 elements) along K inside the swizzle atom = +2 in the >>4 address field.
          // MN-major: 64-element MN blocks BK*128 B apart (LBO), 8-k-row groups 1024 B apart (SBO); the second M = 128 half
          // starts two MN blocks further; 16 k-rows = +2048 B = +128 in the address field.
          const uint64_t ad0 = MNMAJOR ? umma_smem_desc_mn(a_addr, BK * 128, 1024) : umma_smem_desc(a_addr, 128);
          const uint64_t ad1 = MNMAJOR ? umma_smem_desc_mn(a_addr + 2 * BK * 128, BK * 128, 1024)
                                       : umma_smem_desc(a_addr + 128 * BK * 2, 128);
          const uint64_t bd = MNMAJOR ? umma_smem_desc_mn(b_addr, BK * 128, 1024) : umma_smem_desc(b_addr, 128);
          constexpr uint32_t KSTEP = MNMAJOR ? 128u : 2u;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint32_t accum = ((kb - kb0) | k) != 0 ? 1u : 0u;
            umma_bf16(acc0, ad0 + KSTEP * k, bd + KSTEP * k, idesc, accum);
            umma_bf16(acc1, ad1 + KSTEP * k, bd + KSTEP * k, idesc, accum);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[as]);
        FZ_TRACE(4);   // all MMAs of the tile issued
      }
    }
    __syncwarp();
  } else if (warp >= EPI_WARP0) {
    const int ew = warp - EPI_WARP0;            // 0..15
    const int q = warp & 3;                     // TMEM lane quarter this warp may read
    const int half = (ew >> 2) & 1;             // which M=128 accumulator
    const int colgrp = ew >> 3;                 // chunks colgrp, colgrp+2, ...
    const int bar_id = 1 + half * 2 + colgrp;   // named barriers 1..4: one per (row half, column group) of 4 warps
    uint32_t lt = 0;
    for (int t = t_first; t < total_work; t += t_step, ++lt) {
      const int tile = t / p.splits, sp = t - tile * p.splits;
      const int te = p.reverse ? total_tiles - 1 - tile : tile;
      const int n0 = (te % n_tiles) * BN;
      const int m0 = (te / n_tiles) * BM;
      // split-K piece sp writes its fp32 partial tile to its own [M][N] plane of the workspace
      void* out_base = reinterpret_cast<char*>(p.out) + static_cast<size_t>(sp) * p.M * p.N * sizeof(float);
      const uint32_t as = lt % ACC_STAGES;
      const uint32_t aph = (lt / ACC_STAGES) & 1;
      float* sq_buf = sSq + (lt & 1) * 8 * BN;
      char* stg = reinterpret_cast<char*>(smem + L::OFF_STG) + ew * 4096;
      if (ew == 0 && lane == 0) FZ_TRACE(5);   // epilogue warp 0 waits for the accumulator
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      if (ew == 0 && lane == 0) FZ_TRACE(6);   // accumulator complete
      const uint32_t tbase = tmem + (static_cast<uint32_t>(q * 32) << 16) + as * (2 * BN) + half * BN;
      constexpr int CH_COLS = EpiShape<MODE>::CH_COLS;
#pragma unroll 1
      for (int c = colgrp; c < BN / CH_COLS; c += 2)
        epi_chunk<MODE, F16>(p, tbase + c * CH_COLS, m0 + half * 128 + q * 32, n0 + c * CH_COLS, stg, lane,
                        sq_buf + (half * 4 + q) * BN + c * CH_COLS, out_base, p.tma_store ? &tmO : nullptr);
      // all TMEM reads of this stage are complete (tmem_ld_wait above): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (ew == 0 && lane == 0) FZ_TRACE(7);   // epilogue warp 0 done with the tile
      if (MODE == FZ_EPI_GELU_SUMSQ) {
        // deterministic: fixed-order sum of the four lane-quarter warps, one plain store per column.
        // sq_buf alternates with the tile parity, so one barrier per tile is enough.
        // only the four warps sharing this (row half, column group) meet at the barrier
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        const float* sq = sq_buf + half * 4 * BN;
        if (m0 + half * 128 < p.M && q < CH_COLS / 32)
          for (int c = colgrp; c < BN / CH_COLS; c += 2) {
            const int i = c * CH_COLS + q * 32 + lane;
            p.sumsq[static_cast<size_t>(m0 / 128 + half) * p.N + n0 + i] =
                (sq[i] + sq[BN + i]) + (sq[2 * BN + i] + sq[3 * BN + i]);
          }
      }
    }
  }
  if (p.tma_store && warp >= EPI_WARP0 && lane == 0) tma_store_wait_all();   // this warp's last TMA stores have landed
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TCOLS);
}

template <int BN, int STAGES, int ACC_STAGES, int MODE, bool F16, bool MNMAJOR = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p0, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES>;
  // output map for the epilogue's TMA stores (box = one staged chunk: 32 rows x 128 bytes); not with split-K, whose pieces
  // write row-offset planes of a workspace.  FZ_GEMM_TMA_STORE=0: the round-1 st.global path.
  GemmParams p = p0;
  CUtensorMap tmO = tmA;
  p.tma_store = 0;
  static int tma_store = -1;
  if (tma_store < 0) {
    const char* e = getenv("FZ_GEMM_TMA_STORE");
    tma_store = (e && e[0] == '0') ? 0 : 1;
  }
  if (tma_store && p.splits == 1 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
    constexpr bool F32OUT = EpiShape<MODE>::F32OUT;
    const uint64_t dims[2] = {(uint64_t)p.N, (uint64_t)p.M};
    const uint64_t strides[1] = {(uint64_t)p.N * (F32OUT ? 4 : 2)};
    const uint32_t box[2] = {F32OUT ? 32u : 64u, 32u};
    if (int rc = F32OUT ? make_tmap32(&tmO, p.out, 2, dims, strides, box, 128) : make_tmap16(&tmO, p.out, 2, dims, strides, box, 128))
      return rc;
    p.tma_store = 1;
  }
  auto kern = gemm_bf16_kernel<BN, STAGES, ACC_STAGES, MODE, F16, MNMAJOR>;
  FZ_ENSURE_SMEM(kern, L::BYTES);
  const int sm_count = device_sm_count();
  if (sm_count <= 0) return -2;
  const int tiles = ((p.M + BM - 1) / BM) * (p.N / BN) * p.splits;
  const int grid = tiles < sm_count ? tiles : sm_count;
  kern<<<grid, GEMM_THREADS, L::BYTES, stream>>>(tmA, tmB, tmO, p);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// BN = 128: 3 smem stages (48 KB each) + 64 KB store staging, accumulators double buffered; BN = 64 for narrow outputs.
// (Wide outputs, N % 256 == 0, go to the CTA-pair kernel in gemm_tcgen05_2sm.cu.)
template <int BN, int STAGES, int ACC_STAGES, bool F16>
static int dispatch_mode_fmt(int mode, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t st) {
  switch (mode) {
    case FZ_EPI_BF16: return launch_gemm<BN, STAGES, ACC_STAGES, FZ_EPI_BF16, F16>(a, b, p, st);
    case FZ_EPI_GELU_SUMSQ: return launch_gemm<BN, STAGES, ACC_STAGES, FZ_EPI_GELU_SUMSQ, F16>(a, b, p, st);
    case FZ_EPI_RESID_F32: return launch_gemm<BN, STAGES, ACC_STAGES, FZ_EPI_RESID_F32, F16>(a, b, p, st);
    case FZ_EPI_F32: return launch_gemm<BN, STAGES, ACC_STAGES, FZ_EPI_F32, F16>(a, b, p, st);
    case FZ_EPI_RELU_BF16: return launch_gemm<BN, STAGES, ACC_STAGES, FZ_EPI_RELU_BF16, F16>(a, b, p, st);
    case FZ_EPI_GELU_BF16: return launch_gemm<BN, STAGES, ACC_STAGES, FZ_EPI_GELU_BF16, F16>(a, b, p, st);
  }
  set_error("fz_gemm_bf16: unknown epilogue mode %d", mode);
  return -1;
}
template <int BN, int STAGES, int ACC_STAGES>
static int dispatch_mode(int mode, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t st) {
  return p.f16 ? dispatch_mode_fmt<BN, STAGES, ACC_STAGES, true>(mode, a, b, p, st)
               : dispatch_mode_fmt<BN, STAGES, ACC_STAGES, false>(mode, a, b, p, st);
}

}  // namespace fz

extern "C" int fz_gemm_set_trace(void* device_buffer_64x8_u64) {
  fz::g_trace = reinterpret_cast<unsigned long long*>(device_buffer_64x8_u64);
  return 0;
}

extern "C" int fz_gemm_bf16(const void* A, const void* B, void* out, const float* bias, const float* resid,
                            float* sumsq, int M, int N, int K, int b_batch, int rows_per_sample, int mode,
                            void* stream) {
  using namespace fz;
  const int reverse = (mode & FZ_EPI_REVERSE_TILES) ? 1 : 0;
  const int f16 = (mode & FZ_EPI_OPERANDS_F16) ? 1 : 0;
  mode &= ~(FZ_EPI_REVERSE_TILES | FZ_EPI_OPERANDS_F16);
  FZ_REQUIRE(M > 0 && N > 0 && K > 0, "fz_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
  FZ_REQUIRE(K % BK == 0, "fz_gemm_bf16: K=%d must be a multiple of %d", K, BK);
  FZ_REQUIRE(N % 64 == 0, "fz_gemm_bf16: N=%d must be a multiple of 64", N);
  FZ_REQUIRE(b_batch >= 1, "fz_gemm_bf16: b_batch must be >= 1");
  if (b_batch > 1)
    FZ_REQUIRE(rows_per_sample > 0 && rows_per_sample % BM == 0,
               "fz_gemm_bf16: rows_per_sample=%d must be a positive multiple of %d", rows_per_sample, BM);
  FZ_REQUIRE(mode != FZ_EPI_GELU_SUMSQ || sumsq != nullptr, "fz_gemm_bf16: sumsq buffer required");
  FZ_REQUIRE(mode != FZ_EPI_GELU_SUMSQ || M % 128 == 0, "fz_gemm_bf16: M=%d must be a multiple of 128 with GELU_SUMSQ", M);
  FZ_REQUIRE(bias != nullptr, "fz_gemm_bf16: bias is required (pass zeros)");
  FZ_REQUIRE(mode != FZ_EPI_RESID_F32 || resid != nullptr, "fz_gemm_bf16: residual buffer required");
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : M;
  p.b_batched = b_batch > 1 ? 1 : 0;
  p.bias = bias; p.out = out; p.resid = resid; p.sumsq = sumsq; p.trace = g_trace;
  p.reverse = reverse;
  p.f16 = f16;
  p.splits = 1; p.kb_per_split = K / BK; p.nobias = 0; p.tma_store = 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // CTA-pair kernel (256x256 tile over two SMs, gemm_tcgen05_2sm.cu): wide outputs with enough tiles for 74 pairs.
  // FZ_GEMM_PAIR=0 disables, =2 forces it whenever N % 256 == 0.
  const char* pair_e = getenv("FZ_GEMM_PAIR");
  const int pair_env = pair_e ? atoi(pair_e) : 1;
  if (pair_env != 0 && N % 256 == 0) {
    const long long pair_tiles = static_cast<long long>((M + BM - 1) / BM) * (N / 256);
    if (pair_env == 2 || pair_tiles >= 74) return gemm_pair_launch(A, B, p, b_batch, mode, st);
  }
  // one-SM kernel: 256 x 128 tiles (double-buffered accumulators), 256 x 64 for narrow outputs
  const int BN = (N % 128 == 0) ? 128 : 64;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {BK, BM};
    int rc = make_tmap16(&tmA, A, 2, dims, strides, box, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)b_batch};
    const uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)K * 2 * (uint64_t)N};
    const uint32_t box[3] = {BK, (uint32_t)BN, 1};
    int rc = make_tmap16(&tmB, B, 3, dims, strides, box, 128);
    if (rc) return rc;
  }
  if (BN == 128) return dispatch_mode<128, 3, 2>(mode, tmA, tmB, p, st);
  return dispatch_mode<64, 3, 2>(mode, tmA, tmB, p, st);
}

// ----------------------------------------------------------------------------------------
// Split-K: C[M,N] (fp32) = A[M,K] B[N,K]^T for a SMALL output and a LONG reduction -- the weight gradients of the training
// step (dW = dY^T X: a few output tiles, K = every pixel of the batch).  A plain launch would keep 2..32 of the 148 SMs
// busy; here the K range is cut into `splits` pieces, every (tile, piece) is a work item of the same persistent kernel and
// writes an fp32 partial tile to workspace[piece], and a second kernel adds the pieces in index order (deterministic, no
// atomics).  No bias.
// ----------------------------------------------------------------------------------------
namespace fz {
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float4* __restrict__ ws, float4* __restrict__ out,
                                                            size_t n4, int splits) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n4) return;
  float4 a = ws[i];
  for (int s = 1; s < splits; ++s) {
    const float4 b = ws[static_cast<size_t>(s) * n4 + i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  out[i] = a;
}
}  // namespace fz

extern "C" int fz_gemm_splitk_max_splits(int M, int N, int K) {
  using namespace fz;
  if (M <= 0 || N <= 0 || K <= 0 || K % BK || N % 64) return 1;
  const int sms = device_sm_count();
  const int BN = (N % 128 == 0) ? 128 : 64;
  const long long tiles = static_cast<long long>((M + BM - 1) / BM) * (N / BN);
  const int num_kb = K / BK;
  long long s = sms > 0 ? (2LL * sms) / tiles : 1;      // about two work items per SM
  if (s > num_kb / 4) s = num_kb / 4;                   // every piece keeps >= 4 k-blocks
  return static_cast<int>(s < 1 ? 1 : (s > 64 ? 64 : s));
}

extern "C" int fz_gemm_bf16_splitk(const void* A, const void* B, float* out, float* workspace, int M, int N, int K, int splits,
                                   int flags, void* stream) {
  using namespace fz;
  FZ_REQUIRE(M > 0 && N > 0 && K > 0, "fz_gemm_bf16_splitk: bad shape M=%d N=%d K=%d", M, N, K);
  FZ_REQUIRE(K % BK == 0 && N % 64 == 0, "fz_gemm_bf16_splitk: K=%d must be a multiple of %d and N=%d of 64", K, BK, N);
  FZ_REQUIRE((static_cast<long long>(M) * N) % 4 == 0, "fz_gemm_bf16_splitk: M*N must be a multiple of 4");
  const int num_kb = K / BK;
  FZ_REQUIRE(splits >= 1 && splits <= num_kb, "fz_gemm_bf16_splitk: splits=%d out of range (1..%d)", splits, num_kb);
  FZ_REQUIRE(splits == 1 || workspace != nullptr, "fz_gemm_bf16_splitk: workspace (splits*M*N floats) required");
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.rows_per_sample = M; p.b_batched = 0;
  p.bias = nullptr; p.resid = nullptr; p.sumsq = nullptr; p.trace = nullptr; p.reverse = 0;
  p.f16 = (flags & FZ_EPI_OPERANDS_F16) ? 1 : 0;
  p.nobias = 1; p.tma_store = 0;
  p.kb_per_split = (num_kb + splits - 1) / splits;
  p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;          // no empty piece
  p.out = p.splits > 1 ? static_cast<void*>(workspace) : static_cast<void*>(out);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int BN = (N % 128 == 0) ? 128 : 64;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {BK, BM};
    if (int rc = make_tmap16(&tmA, A, 2, dims, strides, box, 128)) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, 1};
    const uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)K * 2 * (uint64_t)N};
    const uint32_t box[3] = {BK, (uint32_t)BN, 1};
    if (int rc = make_tmap16(&tmB, B, 3, dims, strides, box, 128)) return rc;
  }
  int rc = (BN == 128) ? dispatch_mode<128, 3, 2>(FZ_EPI_F32, tmA, tmB, p, st) : dispatch_mode<64, 3, 2>(FZ_EPI_F32, tmA, tmB, p, st);
  if (rc) return rc;
  if (p.splits > 1) {
    const size_t n4 = static_cast<size_t>(M) * N / 4;
    splitk_reduce_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(workspace), reinterpret_cast<float4*>(out), n4, p.splits);
    FZ_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// Split-K with BOTH operands transposed in memory: out [M][N] = At^T Bt, At [K][M], Bt [K][N] row-major (MN-major UMMA
// operands).  This is dW = dY^T X straight from dY [rows][N_l] and X [rows][K_l]: round 1 made transposed copies of both for
// every layer and step (499 launches, 19 ms of the training step).
extern "C" int fz_gemm_bf16_splitk_tn(const void* At, const void* Bt, float* out, float* workspace, int M, int N, int K,
                                      int splits, int flags, void* stream) {
  using namespace fz;
  FZ_REQUIRE(M > 0 && N > 0 && K > 0, "fz_gemm_bf16_splitk_tn: bad shape M=%d N=%d K=%d", M, N, K);
  FZ_REQUIRE(K % BK == 0 && N % 64 == 0 && M % 8 == 0,
             "fz_gemm_bf16_splitk_tn: K=%d must be a multiple of %d, N=%d of 64, M=%d of 8", K, BK, N, M);
  FZ_REQUIRE((static_cast<long long>(M) * N) % 4 == 0, "fz_gemm_bf16_splitk_tn: M*N must be a multiple of 4");
  const int num_kb = K / BK;
  FZ_REQUIRE(splits >= 1 && splits <= num_kb, "fz_gemm_bf16_splitk_tn: splits=%d out of range (1..%d)", splits, num_kb);
  FZ_REQUIRE(splits == 1 || workspace != nullptr, "fz_gemm_bf16_splitk_tn: workspace (splits*M*N floats) required");
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.rows_per_sample = M; p.b_batched = 0;
  p.bias = nullptr; p.resid = nullptr; p.sumsq = nullptr; p.trace = nullptr; p.reverse = 0;
  p.f16 = (flags & FZ_EPI_OPERANDS_F16) ? 1 : 0;
  p.nobias = 1; p.tma_store = 0;
  p.kb_per_split = (num_kb + splits - 1) / splits;
  p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.out = p.splits > 1 ? static_cast<void*>(workspace) : static_cast<void*>(out);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int BN = (N % 128 == 0) ? 128 : 64;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)M, (uint64_t)K};          // innermost = the GEMM's M index
    const uint64_t strides[1] = {(uint64_t)M * 2};
    const uint32_t box[2] = {64, BK};
    if (int rc = make_tmap16(&tmA, At, 2, dims, strides, box, 128)) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
    const uint64_t strides[1] = {(uint64_t)N * 2};
    const uint32_t box[2] = {64, BK};
    if (int rc = make_tmap16(&tmB, Bt, 2, dims, strides, box, 128)) return rc;
  }
  int rc;
  if (BN == 128)
    rc = p.f16 ? launch_gemm<128, 3, 2, FZ_EPI_F32, true, true>(tmA, tmB, p, st) : launch_gemm<128, 3, 2, FZ_EPI_F32, false, true>(tmA, tmB, p, st);
  else
    rc = p.f16 ? launch_gemm<64, 3, 2, FZ_EPI_F32, true, true>(tmA, tmB, p, st) : launch_gemm<64, 3, 2, FZ_EPI_F32, false, true>(tmA, tmB, p, st);
  if (rc) return rc;
  if (p.splits > 1) {
    const size_t n4 = static_cast<size_t>(M) * N / 4;
    splitk_reduce_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(workspace), reinterpret_cast<float4*>(out), n4, p.splits);
    FZ_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// ----------------------------------------------------------------------------------------
// Plain SIMT GEMM with the same contract (exact erff GELU).  Used by the GPU tests to check
// the tcgen05 kernel independently of any library, and selectable with FZ_GEMM_IMPL=simt for
// bring-up.  Not a fallback: the engine never selects it on its own.
// ----------------------------------------------------------------------------------------
namespace fz {
template <bool F16> struct Op16 { typedef __nv_bfloat16 T; };
template <> struct Op16<true> { typedef __half T; };
__device__ __forceinline__ float ld16(const __nv_bfloat16& x) { return __bfloat162float(x); }
__device__ __forceinline__ float ld16(const __half& x) { return __half2float(x); }
__device__ __forceinline__ void st16(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void st16(__half* p, float v) { *p = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f)); }

template <int MODE, bool F16>
__global__ void gemm_simt_kernel(const void* __restrict__ Av, const void* __restrict__ Bv, GemmParams p) {
  typedef typename Op16<F16>::T T16;
  const T16* A = reinterpret_cast<const T16*>(Av);
  const T16* B = reinterpret_cast<const T16*>(Bv);
  __shared__ float sa[16][17];
  __shared__ float sb[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int n = blockIdx.x * 16 + tx;
  const int m = blockIdx.y * 16 + ty;
  const int m_tile0 = blockIdx.y * 16;
  const int bidx = p.b_batched ? (m_tile0 / p.rows_per_sample) : 0;
  const T16* Bb = B + static_cast<size_t>(bidx) * p.N * p.K;
  float acc = 0.0f;
  for (int k0 = 0; k0 < p.K; k0 += 16) {
    const int am = blockIdx.y * 16 + ty;
    sa[ty][tx] = (am < p.M) ? ld16(A[static_cast<size_t>(am) * p.K + k0 + tx]) : 0.0f;
    const int bn = blockIdx.x * 16 + ty;
    sb[ty][tx] = (bn < p.N) ? ld16(Bb[static_cast<size_t>(bn) * p.K + k0 + tx]) : 0.0f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(sa[ty][k], sb[tx][k], acc);
    __syncthreads();
  }
  if (m >= p.M || n >= p.N) return;
  float v = acc + (p.bias ? p.bias[n] : 0.0f);
  const size_t off = static_cast<size_t>(m) * p.N + n;
  if (MODE == FZ_EPI_GELU_SUMSQ) {
    v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    T16 tmp;
    st16(&tmp, v);
    const float vr = ld16(tmp);     // statistics of the stored (16-bit) value
    atomicAdd(&p.sumsq[static_cast<size_t>(m / 128) * p.N + n], vr * vr);
  } else if (MODE == FZ_EPI_GELU_BF16) {
    v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
  } else if (MODE == FZ_EPI_RELU_BF16) {
    v = fmaxf(v, 0.0f);
  } else if (MODE == FZ_EPI_RESID_F32) {
    v += p.resid[off];
  }
  if (MODE == FZ_EPI_RESID_F32 || MODE == FZ_EPI_F32)
    reinterpret_cast<float*>(p.out)[off] = v;
  else
    st16(reinterpret_cast<T16*>(p.out) + off, v);
}
}  // namespace fz

extern "C" int fz_gemm_bf16_simt(const void* A, const void* B, void* out, const float* bias, const float* resid,
                                 float* sumsq, int M, int N, int K, int b_batch, int rows_per_sample, int mode,
                                 void* stream) {
  using namespace fz;
  FZ_REQUIRE(M > 0 && N > 0 && K > 0 && K % 16 == 0, "fz_gemm_bf16_simt: bad shape M=%d N=%d K=%d", M, N, K);
  if (b_batch > 1) FZ_REQUIRE(rows_per_sample % 16 == 0, "fz_gemm_bf16_simt: rows_per_sample %% 16 != 0");
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : M;
  p.b_batched = b_batch > 1 ? 1 : 0;
  p.bias = bias; p.out = out; p.resid = resid; p.sumsq = sumsq; p.trace = nullptr;
  p.reverse = 0;
  p.splits = 1; p.kb_per_split = K / 16; p.nobias = 0; p.tma_store = 0;
  p.f16 = (mode & FZ_EPI_OPERANDS_F16) ? 1 : 0;
  mode &= ~(FZ_EPI_REVERSE_TILES | FZ_EPI_OPERANDS_F16);
  dim3 grid((N + 15) / 16, (M + 15) / 16), block(16, 16);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (mode == FZ_EPI_GELU_SUMSQ)
    FZ_CHECK_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(float) * static_cast<size_t>((M + 127) / 128) * N, st));
#define FZ_SIMT_CASE(M)                                                     \
  case M:                                                                  \
    if (p.f16) gemm_simt_kernel<M, true><<<grid, block, 0, st>>>(A, B, p); \
    else gemm_simt_kernel<M, false><<<grid, block, 0, st>>>(A, B, p);      \
    break;
  switch (mode) {
    FZ_SIMT_CASE(FZ_EPI_BF16)
    FZ_SIMT_CASE(FZ_EPI_GELU_SUMSQ)
    FZ_SIMT_CASE(FZ_EPI_RESID_F32)
    FZ_SIMT_CASE(FZ_EPI_F32)
    FZ_SIMT_CASE(FZ_EPI_RELU_BF16)
    FZ_SIMT_CASE(FZ_EPI_GELU_BF16)
    default: set_error("fz_gemm_bf16_simt: unknown mode %d", mode); return -1;
  }
#undef FZ_SIMT_CASE
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

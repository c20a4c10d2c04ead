// Row-streaming 3x3 convolution for the HBM-bound tail of the U-Net decoder (256^2 / 512^2 feature
// maps, 16..64 input channels, <= 32 output channels) and the segmentation head.
//
// Same arithmetic and epilogues as conv3x3_tcgen05.cu (smp Conv2dReLU / SegmentationHead,
// flair_hub/models/flair_model.py:418; argmax mode = inference.py:295-352), different data movement:
// the tile-per-CTA kernel re-fetches every input row nine times (one TMA box per tap) and pays a
// full pipeline start-up for K = 9*C_in <= 576, which left these layers 4-15x off the HBM roofline.
// Here a persistent CTA walks DOWN a 128-pixel-wide column strip: every input row (130 px with the
// halo) is fetched from HBM ONCE into a ring of shared-memory row buffers, and the nine taps of an
// output row are nine UMMA descriptors pointing INTO those buffers (row ky, pixel offset kx) --
// possible because the buffers use the un-swizzled K-major core-matrix layout
// [8-channel group][pixel][16 B], in which a one-pixel shift is a +16 B start address.
// Weights (all 9 taps) stay resident in shared memory; accumulators rotate through 4 TMEM stages so
// the MMA warp runs ahead of the 4 epilogue warps.
#include "common.h"
#include "ptx.cuh"
#include "operand.cuh"
#include "../../include/flair_zonal_b200.h"

namespace fz {

struct RowConvParams {
  int B, H, W, Cout;
  int R;                 // output rows per work item
  const float* bias;
  const float* scale;
  void* out;
  int cstride;
  const int32_t* plan;
  const int32_t* own;
  uint8_t* raster;
  int RH, RW, margin;
};

constexpr int ROW_PX = 130;                 // 128 output pixels + 1 halo pixel each side
constexpr int ROW_NR = 8;                   // row-buffer ring slots
constexpr int ROW_AS = 4;                   // TMEM accumulator stages

template <int CIN>
struct RowSmem {
  static constexpr int GROUPS = CIN / 8;
  static constexpr int GROUP_BYTES = ((ROW_PX * 16 + 127) / 128) * 128;   // 2080 -> 2176: TMA destinations are 128 B aligned
  static constexpr int SLOT_BYTES = GROUPS * GROUP_BYTES;
  static constexpr int ROW_TX_BYTES = GROUPS * ROW_PX * 16;              // bytes the TMA actually writes per row
};

// no-swizzle K-major descriptor: 8x16B core matrices, SBO = 128 B between 8-pixel groups,
// LBO = distance between the two 8-channel halves of a K=16 step
__device__ __forceinline__ uint64_t umma_desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;   // layout_type = 0 (SWIZZLE_NONE)
}

template <int CIN, int BN, int MODE>
__global__ void __launch_bounds__(192)
conv3x3_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, RowConvParams p) {
  using L = RowSmem<CIN>;
  constexpr int SWZ = CIN * 2;                       // weight tile swizzle (32/64/128 B rows)
  constexpr int WTAP_BYTES = BN * CIN * 2;
  constexpr int OFF_ROWS = ((9 * WTAP_BYTES + 1023) / 1024) * 1024;
  constexpr int OFF_BAR = OFF_ROWS + ROW_NR * L::SLOT_BYTES;
  constexpr int TCOLS = (ROW_AS * BN) < 32 ? 32 : (ROW_AS * BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sRow = smem + OFF_ROWS;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty = full + ROW_NR;
  uint64_t* tfull = empty + ROW_NR;
  uint64_t* tempty = tfull + ROW_AS;
  uint64_t* wfull = tempty + ROW_AS;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int segs = p.W / 128;
  const int runs = (p.H + p.R - 1) / p.R;   // the last run of an image may be shorter
  const int items = p.B * segs * runs;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < ROW_NR; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < ROW_AS; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 4);
    }
    mbar_init(wfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tslot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp == 0) {
    if (lane == 0) {
      // resident weights: 9 tap tiles [BN][CIN], hardware-swizzled
      mbar_arrive_expect_tx(wfull, 9 * WTAP_BYTES);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(&tmB, wfull, sW + tap * WTAP_BYTES, tap * CIN, 0);
      uint32_t g = 0;   // global input-row counter
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int seg = it % segs, run = (it / segs) % runs, b = it / (runs * segs);   // segments of a row run side by side
        const int y0 = run * p.R, x0 = seg * 128;
        const int Rn = min(p.R, p.H - y0);
        for (int r = 0; r < Rn + 2; ++r, ++g) {
          const int s = g % ROW_NR;
          const uint32_t ph = (g / ROW_NR) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], L::ROW_TX_BYTES);
#pragma unroll
          for (int grp = 0; grp < L::GROUPS; ++grp)
            tma_load_4d(&tmA, &full[s], sRow + s * L::SLOT_BYTES + grp * L::GROUP_BYTES, grp * 8, x0 - 1, y0 - 1 + r, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc16(128, BN, OP_F16);
      mbar_wait(wfull, 0);
      uint32_t g = 0, ro = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int Rn = min(p.R, p.H - ((it / segs) % runs) * p.R);
        for (int j = 0; j < Rn; ++j, ++ro) {
          // input rows g+j .. g+j+2 must have landed (rows land in order)
          const int first = (j == 0) ? 0 : 2;
          for (int d = first; d < 3; ++d) {
            const uint32_t gi = g + j + d;
            mbar_wait(&full[gi % ROW_NR], (gi / ROW_NR) & 1);
          }
          const uint32_t as = ro % ROW_AS;
          mbar_wait(&tempty[as], ((ro / ROW_AS) & 1) ^ 1);
          tc_fence_after();
          const uint32_t acc = tmem + as * BN;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint32_t slot = (g + j + ky) % ROW_NR;
            const uint32_t rbase = smem_u32(sRow + slot * L::SLOT_BYTES);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint64_t bd = umma_smem_desc(smem_u32(sW + (ky * 3 + kx) * WTAP_BYTES), SWZ);
#pragma unroll
              for (int k = 0; k < CIN / 16; ++k) {
                const uint64_t ad = umma_desc_noswz(rbase + (2 * k) * L::GROUP_BYTES + kx * 16, L::GROUP_BYTES, 128);
                umma_bf16(acc, ad, bd + 2 * k, idesc, (ky | kx | k) != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(&tfull[as]);
          umma_commit(&empty[(g + j) % ROW_NR]);          // oldest row of the window is done
        }
        // the last two rows of the item were only read, release them too
        umma_commit(&empty[(g + Rn) % ROW_NR]);
        umma_commit(&empty[(g + Rn + 1) % ROW_NR]);
        g += Rn + 2;
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    float sc[BN], bi[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) {
      sc[j] = p.scale ? p.scale[j] : 1.0f;
      bi[j] = p.bias ? p.bias[j] : 0.0f;
    }
    uint32_t ro = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int seg = it % segs, run = (it / segs) % runs, b = it / (runs * segs);   // segments of a row run side by side
      const int x = seg * 128 + m;
      const int Rn = min(p.R, p.H - run * p.R);
      for (int j = 0; j < Rn; ++j, ++ro) {
        const int y = run * p.R + j;
        const uint32_t as = ro % ROW_AS;
        mbar_wait(&tfull[as], (ro / ROW_AS) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
        float v[BN];
#pragma unroll
        for (int c = 0; c < BN / 16; ++c) {
          uint32_t r[16];
          tmem_ld16(taddr + c * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) v[c * 16 + jj] = fmaf(__uint_as_float(r[jj]), sc[c * 16 + jj], bi[c * 16 + jj]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
        const size_t pix = (static_cast<size_t>(b) * p.H + y) * p.W + x;
        if (MODE == FZ_CONV_RELU_BF16) {
          uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<op_t*>(p.out) + pix * p.Cout);
#pragma unroll
          for (int c = 0; c < BN / 8; ++c)
            if (c * 8 < p.Cout)
              op[c] = make_uint4(pack_op(fmaxf(v[8 * c], 0.f), fmaxf(v[8 * c + 1], 0.f)),
                                 pack_op(fmaxf(v[8 * c + 2], 0.f), fmaxf(v[8 * c + 3], 0.f)),
                                 pack_op(fmaxf(v[8 * c + 4], 0.f), fmaxf(v[8 * c + 5], 0.f)),
                                 pack_op(fmaxf(v[8 * c + 6], 0.f), fmaxf(v[8 * c + 7], 0.f)));
        } else if (MODE == FZ_CONV_LOGITS_F32) {
          float* op = reinterpret_cast<float*>(p.out) + pix * p.cstride;
#pragma unroll
          for (int c = 0; c < BN; c += 4)
            if (c < p.cstride)
              *reinterpret_cast<float4*>(op + c) = make_float4(v[c], c + 1 < p.Cout ? v[c + 1] : 0.f,
                                                               c + 2 < p.Cout ? v[c + 2] : 0.f,
                                                               c + 3 < p.Cout ? v[c + 3] : 0.f);
        } else if (MODE == FZ_CONV_LOGITS_F32_NCHW) {
          float* op = reinterpret_cast<float*>(p.out) + (static_cast<size_t>(b) * p.Cout * p.H + y) * p.W + x;
          const size_t plane = static_cast<size_t>(p.H) * p.W;
#pragma unroll
          for (int c = 0; c < BN; ++c)
            if (c < p.Cout) op[c * plane] = v[c];
        } else {  // FZ_CONV_ARGMAX_RASTER: first maximal class wins (np.argmax)
          int best = 0;
          float bv = v[0];
#pragma unroll
          for (int c = 1; c < BN; ++c)
            if (c < p.Cout && v[c] > bv) {
              bv = v[c];
              best = c;
            }
          const int32_t* pl = p.plan + 6 * b;
          const int top = pl[2], left = pl[3];
          int r0 = top, r1 = top + pl[4], c0 = left, c1 = left + pl[5];
          if (p.own) {
            const int32_t* o = p.own + 4 * b;
            r0 = max(r0, o[0]); r1 = min(r1, o[1]); c0 = max(c0, o[2]); c1 = min(c1, o[3]);
          }
          const int rr = top + (y - p.margin), cc = left + (x - p.margin);
          if (rr >= r0 && rr < r1 && cc >= c0 && cc < c1) p.raster[static_cast<size_t>(rr) * p.RW + cc] = static_cast<uint8_t>(best);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TCOLS);
}

template <int CIN, int BN, int MODE>
static int launch_rows(const CUtensorMap& a, const CUtensorMap& b, const RowConvParams& p, cudaStream_t st) {
  using L = RowSmem<CIN>;
  constexpr int WTAP_BYTES = BN * CIN * 2;
  constexpr int OFF_ROWS = ((9 * WTAP_BYTES + 1023) / 1024) * 1024;
  constexpr int BYTES = OFF_ROWS + ROW_NR * L::SLOT_BYTES + (2 * ROW_NR + 2 * ROW_AS + 1) * 8 + 16 + 1024;
  auto kern = conv3x3_rows_kernel<CIN, BN, MODE>;
  FZ_ENSURE_SMEM(kern, BYTES);
  const int sm_count = device_sm_count();
  if (sm_count <= 0) return -2;
  // co-resident CTAs hide the per-row latency chain; bounded by shared memory and by TMEM (512 columns)
  int per_sm = (227 * 1024) / (BYTES + 1024);
  constexpr int tcols = (ROW_AS * BN) < 32 ? 32 : (ROW_AS * BN);
  if (per_sm > 512 / tcols) per_sm = 512 / tcols;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 6) per_sm = 6;
  // rows per work item: every item re-reads a 2-row halo and refills the row pipeline, so long runs win as long as
  // every SM still holds >= 2 CTAs (measured at B = 37, 512^2, 16 -> 16: R = 16/32/64/128 -> 208/201/172/169 us; run
  // lengths that are not a power of two are slower, R = 26 -> 250 us, so they are not considered)
  RowConvParams q = p;
  int grid = 0;
  for (int R = 128; R >= 16; R >>= 1) {
    if (R > p.H && R > 16) continue;
    const int it = p.B * (p.W / 128) * ((p.H + R - 1) / R);
    q.R = R;
    grid = it < sm_count * per_sm ? it : sm_count * per_sm;
    if (it >= 2 * sm_count) break;
  }
  if (const char* e = getenv("FZ_ROWS_R")) {   // experiment override
    q.R = atoi(e);
    const int it = p.B * (p.W / 128) * ((p.H + q.R - 1) / q.R);
    grid = it < sm_count * per_sm ? it : sm_count * per_sm;
  }
  kern<<<grid, 192, BYTES, st>>>(a, b, q);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int MODE>
static int dispatch_rows(int CIN, int BN, const CUtensorMap& a, const CUtensorMap& b, const RowConvParams& p,
                         cudaStream_t st) {
#define FZ_CASE(cin, bn) \
  if (CIN == cin && BN == bn) return launch_rows<cin, bn, MODE>(a, b, p, st);
  if (MODE == FZ_CONV_RELU_BF16) {
    FZ_CASE(16, 16) FZ_CASE(32, 16) FZ_CASE(32, 32) FZ_CASE(64, 32) FZ_CASE(64, 16) FZ_CASE(16, 32)
  } else {
    FZ_CASE(16, 32) FZ_CASE(32, 32) FZ_CASE(64, 32)
  }
#undef FZ_CASE
  set_error("conv3x3 rows: no kernel for Cin=%d BN=%d mode=%d", CIN, BN, MODE);
  return -1;
}

bool conv_rows_applicable(int H, int W, int Cin, int Cout, int mode) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("FZ_CONV_ROWS");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled) return false;
  if (W % 128 != 0 || H % 16 != 0) return false;
  if (!(Cin == 16 || Cin == 32 || Cin == 64)) return false;
  if (mode == FZ_CONV_RELU_BF16) return Cout == 16 || Cout == 32;
  return Cout <= 32;
}

int conv_rows_launch(const void* in, const void* w, const float* scale, const float* bias, void* out, int B, int H,
                     int W, int Cin, int Cout, int w_rows, int mode, int cstride, const int32_t* plan,
                     const int32_t* own, uint8_t* raster, int RH, int RW, int margin, cudaStream_t st) {
  const int BN = (mode == FZ_CONV_RELU_BF16) ? Cout : 32;
  FZ_REQUIRE(w_rows >= BN, "conv3x3 rows: weight rows %d < %d", w_rows, BN);
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    const uint32_t box[4] = {8, ROW_PX, 1, 1};
    int rc = make_tmap16(&tmA, in, 4, dims, strides, box, 0);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)9 * Cin, (uint64_t)w_rows};
    const uint64_t strides[1] = {(uint64_t)9 * Cin * 2};
    const uint32_t box[2] = {(uint32_t)Cin, (uint32_t)BN};
    int rc = make_tmap16(&tmB, w, 2, dims, strides, box, Cin * 2);
    if (rc) return rc;
  }
  RowConvParams p;
  p.B = B; p.H = H; p.W = W; p.Cout = Cout;
  p.R = 32;   // launch_rows picks the real value
  p.bias = bias; p.scale = scale; p.out = out; p.cstride = cstride; p.plan = plan; p.own = own; p.raster = raster;
  p.RH = RH; p.RW = RW; p.margin = margin;
  switch (mode) {
    case FZ_CONV_RELU_BF16: return dispatch_rows<FZ_CONV_RELU_BF16>(Cin, BN, tmA, tmB, p, st);
    case FZ_CONV_LOGITS_F32: return dispatch_rows<FZ_CONV_LOGITS_F32>(Cin, BN, tmA, tmB, p, st);
    case FZ_CONV_LOGITS_F32_NCHW: return dispatch_rows<FZ_CONV_LOGITS_F32_NCHW>(Cin, BN, tmA, tmB, p, st);
    case FZ_CONV_ARGMAX_RASTER: return dispatch_rows<FZ_CONV_ARGMAX_RASTER>(Cin, BN, tmA, tmB, p, st);
  }
  set_error("conv3x3 rows: unknown mode %d", mode);
  return -1;
}

}  // namespace fz

// conv3x3(nearest_upsample_x2(a)) + BN + ReLU without ever building the upsampled tensor: the first convolution of the
// U-Net decoder blocks that have no skip connection (smp UnetDecoder blocks 3 / 4 for the timm encoders, block 4 for
// ResNet; flair_hub/models/flair_model.py:417-419 -> smp DecoderBlock: F.interpolate(scale_factor=2, 'nearest') ->
// Conv2dReLU).  These layers sit on the HBM roofline; the generic path wrote the 4x larger upsampled tensor
// (fz_upsample2_concat) and read it back through the 3x3 convolution.
//
// Sub-pixel decomposition: output pixel (2y+py, 2x+px) only ever sees the 2x2 source pixels
//   rows {y-1, y} (py = 0) or {y, y+1} (py = 1),  columns {x-1, x} (px = 0) or {x, x+1} (px = 1),
// because the 3 taps along an axis fall on 2 source pixels.  The taps that hit the same source pixel are summed on the
// host (fp32 sum, one bf16 rounding): 4 phases x 4 merged taps = 16 weight tiles instead of 9 taps on 4x the pixels,
// i.e. 2.25x fewer MMAs and 4x fewer input bytes.  Zero padding of the upsampled image coincides with TMA's
// out-of-bounds zero fill on the source (row -1 / H, column -1 / W).
//
// Data movement is conv3x3_rows.cu's: a persistent CTA walks down a strip of 128 SOURCE pixels (256 output pixels),
// each source row (130 px with halo) is fetched once into a ring of un-swizzled K-major row buffers, taps are shifted
// UMMA descriptors into those buffers.  One source row yields four [128 x BN] accumulators (py, px); an epilogue thread
// owns one source pixel and writes, per output row, the two adjacent output pixels (2*Cout bf16, contiguous).
#include "common.h"
#include "ptx.cuh"
#include "operand.cuh"
#include "../../include/flair_zonal_b200.h"

namespace fz {

struct UpConvParams {
  int B, H, W, Cout;     // SOURCE height / width
  int R;                 // source rows per work item
  const float* bias;
  const float* scale;
  op_t* out;    // [B][2H][2W][Cout]
};

constexpr int UP_PX = 130;
constexpr int UP_NR = 8;                    // row-buffer ring slots

template <int CIN>
struct UpSmem {
  static constexpr int GROUPS = CIN / 8;
  static constexpr int GROUP_BYTES = ((UP_PX * 16 + 127) / 128) * 128;
  static constexpr int SLOT_BYTES = GROUPS * GROUP_BYTES;
  static constexpr int ROW_TX_BYTES = GROUPS * UP_PX * 16;
};

__device__ __forceinline__ uint64_t up_desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

template <int CIN, int BN>
__global__ void __launch_bounds__(192)
upconv3x3_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, UpConvParams p) {
  using L = UpSmem<CIN>;
  constexpr int SWZ = CIN * 2;
  constexpr int WTAP_BYTES = BN * CIN * 2;
  constexpr int OFF_ROWS = ((16 * WTAP_BYTES + 1023) / 1024) * 1024;
  constexpr int OFF_BAR = OFF_ROWS + UP_NR * L::SLOT_BYTES;
  constexpr int AS = (CIN == 32 && BN == 32) ? 2 : 4;                  // accumulator stages of 4 x BN columns
  constexpr int TCOLS = AS * 4 * BN < 32 ? 32 : AS * 4 * BN;
  static_assert(AS >= 2, "two accumulator stages at least");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sRow = smem + OFF_ROWS;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty = full + UP_NR;
  uint64_t* tfull = empty + UP_NR;
  uint64_t* tempty = tfull + AS;
  uint64_t* wfull = tempty + AS;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int segs = p.W / 128;
  const int runs = (p.H + p.R - 1) / p.R;   // the last run of an image may be shorter
  const int items = p.B * segs * runs;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < UP_NR; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < AS; ++s) {
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
      // resident weights: 16 merged-tap tiles [BN][CIN], tile index = ((py*2+px)*2+ra)*2+ca
      mbar_arrive_expect_tx(wfull, 16 * WTAP_BYTES);
      for (int tap = 0; tap < 16; ++tap) tma_load_2d(&tmB, wfull, sW + tap * WTAP_BYTES, tap * CIN, 0);
      uint32_t g = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int seg = it % segs, run = (it / segs) % runs, b = it / (runs * segs);   // segments of a row run side by side
        const int y0 = run * p.R, x0 = seg * 128;
        const int Rn = min(p.R, p.H - y0);
        for (int r = 0; r < Rn + 2; ++r, ++g) {
          const int s = g % UP_NR;
          const uint32_t ph = (g / UP_NR) & 1;
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
          const int first = (j == 0) ? 0 : 2;       // source rows g+j .. g+j+2 (rows land in order)
          for (int d = first; d < 3; ++d) {
            const uint32_t gi = g + j + d;
            mbar_wait(&full[gi % UP_NR], (gi / UP_NR) & 1);
          }
          const uint32_t as = ro % AS;
          mbar_wait(&tempty[as], ((ro / AS) & 1) ^ 1);
          tc_fence_after();
#pragma unroll
          for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int px = 0; px < 2; ++px) {
              const uint32_t acc = tmem + as * (4 * BN) + (py * 2 + px) * BN;
#pragma unroll
              for (int ra = 0; ra < 2; ++ra) {
                const uint32_t slot = (g + j + py + ra) % UP_NR;          // source row y-1+py+ra
                const uint32_t rbase = smem_u32(sRow + slot * L::SLOT_BYTES);
#pragma unroll
                for (int ca = 0; ca < 2; ++ca) {
                  const int tile = ((py * 2 + px) * 2 + ra) * 2 + ca;
                  const uint64_t bd = umma_smem_desc(smem_u32(sW + tile * WTAP_BYTES), SWZ);
#pragma unroll
                  for (int k = 0; k < CIN / 16; ++k) {
                    // buffer pixel 0 is source column x0-1: source column x0+i-1+px+ca sits at i + px + ca
                    const uint64_t ad = up_desc_noswz(rbase + (2 * k) * L::GROUP_BYTES + (px + ca) * 16,
                                                      L::GROUP_BYTES, 128);
                    umma_bf16(acc, ad, bd + 2 * k, idesc, (ra | ca | k) != 0 ? 1u : 0u);
                  }
                }
              }
            }
          umma_commit(&tfull[as]);
          umma_commit(&empty[(g + j) % UP_NR]);
        }
        umma_commit(&empty[(g + Rn) % UP_NR]);
        umma_commit(&empty[(g + Rn + 1) % UP_NR]);
        g += Rn + 2;
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;                       // source pixel of this thread inside the 128-px segment
    float sc[BN], bi[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) {
      sc[j] = p.scale ? p.scale[j] : 1.0f;
      bi[j] = p.bias ? p.bias[j] : 0.0f;
    }
    const int OW = 2 * p.W;
    uint32_t ro = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int seg = it % segs, run = (it / segs) % runs, b = it / (runs * segs);   // segments of a row run side by side
      const int xs = seg * 128 + m;
      const int Rn = min(p.R, p.H - run * p.R);
      for (int j = 0; j < Rn; ++j, ++ro) {
        const int ys = run * p.R + j;
        const uint32_t as = ro % AS;
        mbar_wait(&tfull[as], (ro / AS) & 1);
        tc_fence_after();
        const uint32_t tbase = tmem + (static_cast<uint32_t>(q * 32) << 16) + as * (4 * BN);
#pragma unroll
        for (int py = 0; py < 2; ++py) {
          // output pixels (2ys+py, 2xs) and (2ys+py, 2xs+1): 2*Cout bf16 contiguous
          op_t* op = p.out + ((static_cast<size_t>(b) * 2 * p.H + 2 * ys + py) * OW + 2 * xs) * p.Cout;
#pragma unroll
          for (int px = 0; px < 2; ++px) {
            float v[BN];
#pragma unroll
            for (int c = 0; c < BN / 16; ++c) {
              uint32_t r[16];
              tmem_ld16(tbase + (py * 2 + px) * BN + c * 16, r);
              tmem_ld_wait();
#pragma unroll
              for (int jj = 0; jj < 16; ++jj)
                v[c * 16 + jj] = fmaxf(fmaf(__uint_as_float(r[jj]), sc[c * 16 + jj], bi[c * 16 + jj]), 0.f);
            }
            uint4* o4 = reinterpret_cast<uint4*>(op + px * p.Cout);
#pragma unroll
            for (int c = 0; c < BN / 8; ++c)
              o4[c] = make_uint4(pack_op(v[8 * c], v[8 * c + 1]), pack_op(v[8 * c + 2], v[8 * c + 3]),
                                 pack_op(v[8 * c + 4], v[8 * c + 5]), pack_op(v[8 * c + 6], v[8 * c + 7]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TCOLS);
}

template <int CIN, int BN>
static int launch_upconv(const CUtensorMap& a, const CUtensorMap& b, const UpConvParams& p, cudaStream_t st) {
  using L = UpSmem<CIN>;
  constexpr int WTAP_BYTES = BN * CIN * 2;
  constexpr int OFF_ROWS = ((16 * WTAP_BYTES + 1023) / 1024) * 1024;
  constexpr int AS = (CIN == 32 && BN == 32) ? 2 : 4;
  constexpr int BYTES = OFF_ROWS + UP_NR * L::SLOT_BYTES + (2 * UP_NR + 2 * AS + 1) * 8 + 16 + 1024;
  auto kern = upconv3x3_rows_kernel<CIN, BN>;
  FZ_ENSURE_SMEM(kern, BYTES);
  const int sm_count = device_sm_count();
  if (sm_count <= 0) return -2;
  int per_sm = (227 * 1024) / (BYTES + 1024);
  constexpr int tcols = AS * 4 * BN < 32 ? 32 : AS * 4 * BN;
  if (per_sm > 512 / tcols) per_sm = 512 / tcols;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 6) per_sm = 6;
  // rows per work item: every item re-reads a 2-row halo and refills the row pipeline, so long runs win as long as
  // every SM still holds >= 2 CTAs (measured at B = 37, 512^2, 16 -> 16: R = 16/32/64/128 -> 208/201/172/169 us; run
  // lengths that are not a power of two are slower, R = 26 -> 250 us, so they are not considered)
  UpConvParams q = p;
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

}  // namespace fz

extern "C" int fz_upconv3x3_bn_relu(const void* in, const void* w16, const float* scale, const float* bias, void* out,
                                    int B, int H, int W, int Cin, int Cout, int w_rows, void* stream) {
  using namespace fz;
  FZ_REQUIRE(B >= 0 && H > 0 && W > 0, "fz_upconv3x3_bn_relu: bad shape");
  FZ_REQUIRE(W % 128 == 0 && H % 16 == 0, "fz_upconv3x3_bn_relu: source H=%d W=%d must be multiples of 16 x 128", H, W);
  FZ_REQUIRE((Cin == 32 || Cin == 64) && (Cout == 16 || Cout == 32),
             "fz_upconv3x3_bn_relu: Cin=%d Cout=%d not instantiated (32|64 -> 16|32)", Cin, Cout);
  FZ_REQUIRE(w_rows >= Cout, "fz_upconv3x3_bn_relu: weight rows %d < %d", w_rows, Cout);
  if (B == 0) return 0;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    const uint32_t box[4] = {8, UP_PX, 1, 1};
    int rc = make_tmap16(&tmA, in, 4, dims, strides, box, 0);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)16 * Cin, (uint64_t)w_rows};
    const uint64_t strides[1] = {(uint64_t)16 * Cin * 2};
    const uint32_t box[2] = {(uint32_t)Cin, (uint32_t)Cout};
    int rc = make_tmap16(&tmB, w16, 2, dims, strides, box, Cin * 2);
    if (rc) return rc;
  }
  UpConvParams p;
  p.B = B; p.H = H; p.W = W; p.Cout = Cout;
  p.R = 32;   // launch_rows picks the real value
  p.bias = bias; p.scale = scale;
  p.out = reinterpret_cast<op_t*>(out);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (Cin == 64 && Cout == 32) return launch_upconv<64, 32>(tmA, tmB, p, st);
  if (Cin == 64 && Cout == 16) return launch_upconv<64, 16>(tmA, tmB, p, st);
  if (Cin == 32 && Cout == 32) return launch_upconv<32, 32>(tmA, tmB, p, st);
  return launch_upconv<32, 16>(tmA, tmB, p, st);
}

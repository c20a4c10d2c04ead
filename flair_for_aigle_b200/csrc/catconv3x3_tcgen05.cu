// conv3x3(cat(nearest_upsample_x2(a), skip)) + BN + ReLU as ONE implicit GEMM, without building the concatenated tensor:
// the first convolution of the smp U-Net decoder blocks that do have a skip connection
// (flair_hub/models/flair_model.py:417-419 -> smp DecoderBlock: F.interpolate(x, 2, 'nearest'); torch.cat([x, skip], 1);
// Conv2dReLU).  fz_upsample2_concat + fz_conv3x3_bf16 wrote and re-read a tensor 4x the size of `a`.
//
// K is split by source tensor:
//   * the `a` channels use the sub-pixel decomposition of upconv3x3_rows.cu: an output pixel of phase (py, px) sees 2x2
//     source pixels, the 9 taps collapse into 4 merged taps (host: fp32 sum, one bf16 rounding) -> 4/9 of the MACs;
//   * the `skip` channels keep their 9 taps at output resolution.
// A CTA tile is 128 output pixels of ONE phase of a (2 TH) x (2 TW) output region, so that both parts are plain TMA
// boxes: a TH x TW box of `a` at (ys0 + py + ra - 1, xs0 + px + ca - 1), and a TH x TW box of `skip` traversed with
// element stride 2 at (2 ys0 + py + ky - 1, 2 xs0 + px + kx - 1).  Out-of-image coordinates are zero-filled by the TMA
// unit = the convolution's padding of the upsampled / skip image.  Everything accumulates into one TMEM tile.
#include "common.h"
#include "ptx.cuh"
#include "operand.cuh"
#include "../../include/flair_zonal_b200.h"

namespace fz {

struct CatConvParams {
  int B, Hs, Ws;            // SOURCE (a) height / width; output is 2Hs x 2Ws
  int C1, C2, Cout;
  int TW, TH;               // source tile (TW*TH = 128)
  const float* bias;
  const float* scale;
  op_t* out;       // [B][2Hs][2Ws][Cout]
};

template <int BN, int KC, int STAGES>
struct CatSmem {
  static constexpr int A_BYTES = 128 * KC * 2;
  static constexpr int B_BYTES = BN * KC * 2;
  static constexpr int OFF_B = STAGES * A_BYTES;
  static constexpr int OFF_BIAS = OFF_B + STAGES * B_BYTES;
  static constexpr int OFF_SCALE = OFF_BIAS + BN * 4;
  static constexpr int OFF_BAR = (OFF_SCALE + BN * 4 + 7) & ~7;
  static constexpr int OFF_TSLOT = OFF_BAR + (2 * STAGES + 1) * 8;
  static constexpr int BYTES = OFF_TSLOT + 16 + 1024;
};

template <int BN, int KC>
__global__ void __launch_bounds__(192, 2)
catconv3x3_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, CatConvParams p) {
  constexpr int STAGES = 4;
  constexpr int SWZ = KC * 2;
  constexpr int TCOLS = BN < 32 ? 32 : BN;
  using L = CatSmem<BN, KC, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + L::OFF_B;
  float* sBias = reinterpret_cast<float*>(smem + L::OFF_BIAS);
  float* sScale = reinterpret_cast<float*>(smem + L::OFF_SCALE);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + L::OFF_TSLOT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile decomposition: n fastest, then phase, x, y, b (the four phases of a region run together: same `a` boxes in L2)
  const int n_tiles_n = (p.Cout + BN - 1) / BN;
  const int tiles_x = p.Ws / p.TW, tiles_y = p.Hs / p.TH;
  int t = blockIdx.x;
  const int n0 = (t % n_tiles_n) * BN;  t /= n_tiles_n;
  const int phase = t % 4;              t /= 4;
  const int xs0 = (t % tiles_x) * p.TW; t /= tiles_x;
  const int ys0 = (t % tiles_y) * p.TH; t /= tiles_y;
  const int b = t;
  const int py = phase >> 1, px = phase & 1;
  const int chunks1 = p.C1 / KC, chunks2 = p.C2 / KC;
  const int kb1 = 4 * chunks1;
  const int num_kb = kb1 + 9 * chunks2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmB2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tslot, TCOLS);
  if (warp >= 2)
    for (int i = threadIdx.x - 64; i < BN; i += 128) {
      sBias[i] = p.bias ? p.bias[n0 + i] : 0.0f;
      sScale[i] = p.scale ? p.scale[n0 + i] : 1.0f;
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], L::A_BYTES + L::B_BYTES);
        if (kb < kb1) {
          const int t4 = kb / chunks1, c0 = (kb % chunks1) * KC;       // t4 = ra*2 + ca
          const int ra = t4 >> 1, ca = t4 & 1;
          tma_load_4d(&tmA1, &full[s], sA + s * L::A_BYTES, c0, xs0 + px + ca - 1, ys0 + py + ra - 1, b);
          tma_load_2d(&tmB1, &full[s], sB + s * L::B_BYTES, (phase * 4 + t4) * p.C1 + c0, n0);
        } else {
          const int k2 = kb - kb1;
          const int tap = k2 / chunks2, c0 = (k2 % chunks2) * KC;
          const int ky = tap / 3, kx = tap % 3;
          tma_load_4d(&tmA2, &full[s], sA + s * L::A_BYTES, c0, 2 * xs0 + px + kx - 1, 2 * ys0 + py + ky - 1, b);
          tma_load_2d(&tmB2, &full[s], sB + s * L::B_BYTES, tap * (p.C1 + p.C2) + p.C1 + c0, n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc16(128, BN, OP_F16);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t ad = umma_smem_desc(smem_u32(sA + s * L::A_BYTES), SWZ);
        const uint64_t bd = umma_smem_desc(smem_u32(sB + s * L::B_BYTES), SWZ);
#pragma unroll
        for (int k = 0; k < KC / 16; ++k) umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(tfull);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int y = 2 * (ys0 + m / p.TW) + py, x = 2 * (xs0 + m % p.TW) + px;
    mbar_wait(tfull, 0);
    tc_fence_after();
    const uint32_t trow = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const size_t pix = (static_cast<size_t>(b) * 2 * p.Hs + y) * (2 * p.Ws) + x;
#pragma unroll 1
    for (int c = 0; c < BN / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(trow + c * 16, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        v[j] = fmaxf(fmaf(__uint_as_float(r[j]), sScale[c * 16 + j], sBias[c * 16 + j]), 0.0f);
      if (n0 + c * 16 < p.Cout) {
        uint4* op = reinterpret_cast<uint4*>(p.out + pix * p.Cout + n0 + c * 16);
        op[0] = make_uint4(pack_op(v[0], v[1]), pack_op(v[2], v[3]), pack_op(v[4], v[5]), pack_op(v[6], v[7]));
        op[1] = make_uint4(pack_op(v[8], v[9]), pack_op(v[10], v[11]), pack_op(v[12], v[13]),
                           pack_op(v[14], v[15]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TCOLS);
}

template <int BN, int KC>
static int launch_catconv(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b1, const CUtensorMap& b2,
                          const CatConvParams& p, cudaStream_t st) {
  using L = CatSmem<BN, KC, 4>;
  auto kern = catconv3x3_kernel<BN, KC>;
  FZ_ENSURE_SMEM(kern, L::BYTES);
  const int n_tiles_n = (p.Cout + BN - 1) / BN;
  const long long grid = 4LL * p.B * (p.Hs / p.TH) * (p.Ws / p.TW) * n_tiles_n;
  kern<<<static_cast<unsigned>(grid), 192, L::BYTES, st>>>(a1, a2, b1, b2, p);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fz

extern "C" int fz_catconv3x3_bn_relu(const void* a, const void* skip, const void* w16a, const void* w,
                                     const float* scale, const float* bias, void* out, int B, int Hs, int Ws, int C1,
                                     int C2, int Cout, int w_rows, void* stream) {
  using namespace fz;
  FZ_REQUIRE(B >= 0 && Hs > 0 && Ws > 0, "fz_catconv3x3_bn_relu: bad shape");
  FZ_REQUIRE(C1 % 64 == 0 && C2 % 64 == 0 && C1 > 0 && C2 > 0, "fz_catconv3x3_bn_relu: C1=%d C2=%d must be multiples of 64",
             C1, C2);
  FZ_REQUIRE(Cout % 64 == 0, "fz_catconv3x3_bn_relu: Cout=%d must be a multiple of 64", Cout);
  if (B == 0) return 0;
  constexpr int KC = 64;
  const int BN = (Cout % 128 == 0) ? 128 : 64;
  const int n_tiles_n = (Cout + BN - 1) / BN;
  FZ_REQUIRE(w_rows >= n_tiles_n * BN, "fz_catconv3x3_bn_relu: weight rows %d < %d", w_rows, n_tiles_n * BN);
  const int TW = Ws < 128 ? Ws : 128;
  FZ_REQUIRE(128 % TW == 0, "fz_catconv3x3_bn_relu: Ws=%d must divide 128 or be a multiple of 128", Ws);
  const int TH = 128 / TW;
  FZ_REQUIRE(Ws % TW == 0 && Hs % TH == 0, "fz_catconv3x3_bn_relu: Hs=%d Ws=%d not tileable by %dx%d", Hs, Ws, TH, TW);
  FZ_REQUIRE(2 * TW <= 256 && 2 * TH <= 256, "fz_catconv3x3_bn_relu: tile too wide for a strided TMA box");
  CUtensorMap tmA1, tmA2, tmB1, tmB2;
  {
    const uint64_t dims[4] = {(uint64_t)C1, (uint64_t)Ws, (uint64_t)Hs, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C1 * 2, (uint64_t)Ws * C1 * 2, (uint64_t)Hs * Ws * C1 * 2};
    const uint32_t box[4] = {KC, (uint32_t)TW, (uint32_t)TH, 1};
    int rc = make_tmap16(&tmA1, a, 4, dims, strides, box, KC * 2);
    if (rc) return rc;
  }
  {
    const uint64_t W2 = 2ull * Ws, H2 = 2ull * Hs;
    const uint64_t dims[4] = {(uint64_t)C2, W2, H2, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C2 * 2, W2 * C2 * 2, H2 * W2 * C2 * 2};
    const uint32_t box[4] = {KC, (uint32_t)(2 * TW), (uint32_t)(2 * TH), 1};     // traversed elements; stride 2 -> TW x TH land
    const uint32_t estr[4] = {1, 2, 2, 1};
    int rc = make_tmap16(&tmA2, skip, 4, dims, strides, box, KC * 2, estr);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)16 * C1, (uint64_t)w_rows};
    const uint64_t strides[1] = {(uint64_t)16 * C1 * 2};
    const uint32_t box[2] = {KC, (uint32_t)BN};
    int rc = make_tmap16(&tmB1, w16a, 2, dims, strides, box, KC * 2);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)9 * (C1 + C2), (uint64_t)w_rows};
    const uint64_t strides[1] = {(uint64_t)9 * (C1 + C2) * 2};
    const uint32_t box[2] = {KC, (uint32_t)BN};
    int rc = make_tmap16(&tmB2, w, 2, dims, strides, box, KC * 2);
    if (rc) return rc;
  }
  CatConvParams p;
  p.B = B; p.Hs = Hs; p.Ws = Ws; p.C1 = C1; p.C2 = C2; p.Cout = Cout; p.TW = TW; p.TH = TH;
  p.bias = bias; p.scale = scale;
  p.out = reinterpret_cast<op_t*>(out);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (BN == 128) return launch_catconv<128, KC>(tmA1, tmA2, tmB1, tmB2, p, st);
  return launch_catconv<64, KC>(tmA1, tmA2, tmB1, tmB2, p, st);
}

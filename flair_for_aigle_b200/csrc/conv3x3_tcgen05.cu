// 3x3 / pad 1 / stride 1 convolution on NHWC bf16 as an implicit GEMM on tcgen05 + TMEM.
//
// Replaces smp UnetDecoder's Conv2dReLU (conv3x3 + BatchNorm + ReLU, BN folded on the host) and
// SegmentationHead conv3x3 invoked from flair_hub/models/flair_model.py:418
// (monotemp_model.py:22-31), SURVEY.md K4, and -- in its argmax mode -- the logits D2H +
// numpy crop/argmax/write of flair_zonal_detection/inference.py:295-352 (K8).
//
// M = 128 output pixels of one image (a TH x TW box, TH*TW = 128), N = output channels,
// K = 9 taps x C_in.  The A operand is never materialised: for tap (ky,kx) and channel chunk c0
// the producer issues ONE 4-D TMA box {KC, TW, TH, 1} at (c0, x0+kx-1, y0+ky-1, b); the TMA
// unit zero-fills out-of-image coordinates, which is exactly the conv's zero padding, and
// writes the box as 128 K-major rows with the hardware swizzle the UMMA descriptor expects.
#include "common.h"
#include "ptx.cuh"
#include "operand.cuh"
#include "../../include/flair_zonal_b200.h"

namespace fz {

struct ConvParams {
  int B, H, W, Cin, Cout;   // H, W = OUTPUT size; Cout = channels actually stored / compared
  int stride;               // 1 or 2 (input is H*stride x W*stride)
  const void* resid;        // ADD_RELU: bf16 [B,H,W,Cout] added before the ReLU (ResNet BasicBlock identity)
  int TW, TH;               // output tile (TW*TH = 128)
  const float* bias;        // [n_tiles_n * BN] (zero padded)
  const float* scale;       // [n_tiles_n * BN] per-channel multiplier (folded BatchNorm) or nullptr
  void* out;                // RELU_BF16: bf16 [B,H,W,Cout]; LOGITS_F32: f32 [B,H,W,cstride]
  int cstride;
  // ARGMAX_RASTER
  const int32_t* plan;      // [B][6]
  const int32_t* own;       // [B][4] or nullptr
  uint8_t* raster;          // [RH][RW]
  int RH, RW, margin;
};

template <int BN, int KC, int STAGES>
struct ConvSmem {
  static constexpr int A_BYTES = 128 * KC * 2;
  static constexpr int B_BYTES = BN * KC * 2;
  static constexpr int OFF_B = STAGES * A_BYTES;
  static constexpr int OFF_BIAS = OFF_B + STAGES * B_BYTES;
  static constexpr int OFF_SCALE = OFF_BIAS + BN * 4;
  static constexpr int OFF_BAR = (OFF_SCALE + BN * 4 + 7) & ~7;
  static constexpr int OFF_TSLOT = OFF_BAR + (2 * STAGES + 1) * 8;
  static constexpr int BYTES = OFF_TSLOT + 16 + 1024;
};

template <int BN, int KC, int MODE>
__global__ void __launch_bounds__(192, 2)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvParams p) {
  constexpr int STAGES = 4;
  constexpr int SWZ = KC * 2;
  constexpr int TCOLS = BN < 32 ? 32 : BN;
  using L = ConvSmem<BN, KC, STAGES>;
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
  // tile decomposition: n fastest, then x, y, b
  const int n_tiles_n = (p.Cout + BN - 1) / BN;
  const int tiles_x = p.W / p.TW, tiles_y = p.H / p.TH;
  int t = blockIdx.x;
  const int n0 = (t % n_tiles_n) * BN;  t /= n_tiles_n;
  const int x0 = (t % tiles_x) * p.TW;  t /= tiles_x;
  const int y0 = (t % tiles_y) * p.TH;  t /= tiles_y;
  const int b = t;
  const int chunks = p.Cin / KC;
  const int num_kb = 9 * chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
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
        const int tap = kb / chunks, c0 = (kb % chunks) * KC;
        const int ky = tap / 3, kx = tap % 3;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], L::A_BYTES + L::B_BYTES);
        tma_load_4d(&tmA, &full[s], sA + s * L::A_BYTES, c0, x0 * p.stride + kx - 1, y0 * p.stride + ky - 1, b);
        tma_load_2d(&tmB, &full[s], sB + s * L::B_BYTES, tap * p.Cin + c0, n0);
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
    const int y = y0 + m / p.TW, x = x0 + m % p.TW;
    mbar_wait(tfull, 0);
    tc_fence_after();
    const uint32_t trow = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const size_t pix = (static_cast<size_t>(b) * p.H + y) * p.W + x;
    if (MODE == FZ_CONV_RELU_BF16 || MODE == FZ_CONV_ADD_RELU_BF16 || MODE == FZ_CONV_BF16) {
#pragma unroll 1
      for (int c = 0; c < BN / 16; ++c) {
        uint32_t r[16];
        tmem_ld16(trow + c * 16, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(__uint_as_float(r[j]), sScale[c * 16 + j], sBias[c * 16 + j]);
        if (MODE == FZ_CONV_ADD_RELU_BF16 && n0 + c * 16 < p.Cout) {
          const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const op_t*>(p.resid) +
                                                           pix * p.Cout + n0 + c * 16);
          const uint4 r0 = rp[0], r1 = rp[1];
          const op2_t* h0 = reinterpret_cast<const op2_t*>(&r0);
          const op2_t* h1 = reinterpret_cast<const op2_t*>(&r1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 a = op22ff(h0[j]), bq = op22ff(h1[j]);
            v[2 * j] += a.x; v[2 * j + 1] += a.y; v[8 + 2 * j] += bq.x; v[8 + 2 * j + 1] += bq.y;
          }
        }
        if (MODE != FZ_CONV_BF16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        if (n0 + c * 16 < p.Cout) {
          uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<op_t*>(p.out) + pix * p.Cout + n0 + c * 16);
          op[0] = make_uint4(pack_op(v[0], v[1]), pack_op(v[2], v[3]), pack_op(v[4], v[5]), pack_op(v[6], v[7]));
          op[1] = make_uint4(pack_op(v[8], v[9]), pack_op(v[10], v[11]), pack_op(v[12], v[13]),
                             pack_op(v[14], v[15]));
        }
      }
    } else {
      // head: BN == 32 covers all classes (Cout <= 32) in one N tile
      float v[BN];
#pragma unroll
      for (int c = 0; c < BN / 16; ++c) {
        uint32_t r[16];
        tmem_ld16(trow + c * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[c * 16 + j] = fmaf(__uint_as_float(r[j]), sScale[c * 16 + j], sBias[c * 16 + j]);
      }
      if (MODE == FZ_CONV_LOGITS_F32) {
        float* op = reinterpret_cast<float*>(p.out) + pix * p.cstride;
#pragma unroll
        for (int j = 0; j < BN; j += 4)
          if (j < p.cstride)
            *reinterpret_cast<float4*>(op + j) = make_float4(v[j], j + 1 < p.Cout ? v[j + 1] : 0.f,
                                                             j + 2 < p.Cout ? v[j + 2] : 0.f,
                                                             j + 3 < p.Cout ? v[j + 3] : 0.f);
      } else if (MODE == FZ_CONV_LOGITS_F32_NCHW) {
        // [B][Cout][H][W]: consecutive lanes = consecutive x, one coalesced store per class
        float* op = reinterpret_cast<float*>(p.out) + (static_cast<size_t>(b) * p.Cout * p.H + y) * p.W + x;
        const size_t plane = static_cast<size_t>(p.H) * p.W;
#pragma unroll
        for (int j = 0; j < BN; ++j)
          if (j < p.Cout) op[j * plane] = v[j];
      } else {  // FZ_CONV_ARGMAX_RASTER: first maximal class wins (np.argmax)
        int best = 0;
        float bv = v[0];
#pragma unroll
        for (int j = 1; j < BN; ++j)
          if (j < p.Cout && v[j] > bv) {
            bv = v[j];
            best = j;
          }
        const int32_t* pl = p.plan + 6 * b;
        const int top = pl[2], left = pl[3];
        int r0 = top, r1 = top + pl[4], c0 = left, c1 = left + pl[5];
        if (p.own) {
          const int32_t* o = p.own + 4 * b;
          r0 = max(r0, o[0]); r1 = min(r1, o[1]); c0 = max(c0, o[2]); c1 = min(c1, o[3]);
        }
        const int rr = top + (y - p.margin), cc = left + (x - p.margin);
        if (y >= p.margin && x >= p.margin && rr >= r0 && rr < r1 && cc >= c0 && cc < c1)
          p.raster[static_cast<size_t>(rr) * p.RW + cc] = static_cast<uint8_t>(best);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TCOLS);
}

template <int BN, int KC, int MODE>
static int launch_conv(const CUtensorMap& a, const CUtensorMap& b, const ConvParams& p, cudaStream_t st) {
  using L = ConvSmem<BN, KC, 4>;
  auto kern = conv3x3_kernel<BN, KC, MODE>;
  FZ_ENSURE_SMEM(kern, L::BYTES);
  const int n_tiles_n = (p.Cout + BN - 1) / BN;
  const long long grid = 1LL * p.B * (p.H / p.TH) * (p.W / p.TW) * n_tiles_n;
  kern<<<static_cast<unsigned>(grid), 192, L::BYTES, st>>>(a, b, p);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int MODE>
static int dispatch_conv(int BN, int KC, const CUtensorMap& a, const CUtensorMap& b, const ConvParams& p,
                         cudaStream_t st) {
#define FZ_CASE(bn, kc) \
  if (BN == bn && KC == kc) return launch_conv<bn, kc, MODE>(a, b, p, st);
  if (MODE == FZ_CONV_RELU_BF16) {
    FZ_CASE(128, 64) FZ_CASE(64, 64) FZ_CASE(32, 64) FZ_CASE(32, 32) FZ_CASE(16, 32) FZ_CASE(16, 16)
    FZ_CASE(128, 32) FZ_CASE(64, 32) FZ_CASE(128, 16) FZ_CASE(64, 16) FZ_CASE(32, 16) FZ_CASE(16, 64)
  } else if (MODE == FZ_CONV_ADD_RELU_BF16 || MODE == FZ_CONV_BF16) {
    FZ_CASE(128, 64) FZ_CASE(64, 64)
  } else {
    FZ_CASE(32, 16) FZ_CASE(32, 32) FZ_CASE(32, 64)
  }
#undef FZ_CASE
  set_error("fz_conv3x3_bf16: no kernel for BN=%d KC=%d mode=%d", BN, KC, MODE);
  return -1;
}

bool conv_rows_applicable(int H, int W, int Cin, int Cout, int mode);
int conv_rows_launch(const void* in, const void* w, const float* scale, const float* bias, void* out, int B, int H,
                     int W, int Cin, int Cout, int w_rows, int mode, int cstride, const int32_t* plan,
                     const int32_t* own, uint8_t* raster, int RH, int RW, int margin, cudaStream_t st);

}  // namespace fz

extern "C" int fz_conv3x3_ex(const void* in, const void* w, const float* scale, const float* bias, void* out,
                             const void* resid, int B, int H, int W, int Cin, int Cout, int w_rows, int stride, int mode,
                             int cstride, const int32_t* plan, const int32_t* own, uint8_t* raster, int RH, int RW,
                             int margin, void* stream);

extern "C" int fz_conv3x3_bf16(const void* in, const void* w, const float* scale, const float* bias, void* out, int B,
                               int H, int W, int Cin, int Cout, int w_rows, int mode, int cstride, const int32_t* plan,
                               const int32_t* own, uint8_t* raster, int RH, int RW, int margin, void* stream) {
  return fz_conv3x3_ex(in, w, scale, bias, out, nullptr, B, H, W, Cin, Cout, w_rows, 1, mode, cstride, plan, own, raster,
                       RH, RW, margin, stream);
}

extern "C" int fz_conv3x3_ex(const void* in, const void* w, const float* scale, const float* bias, void* out,
                             const void* resid, int B, int H, int W, int Cin, int Cout, int w_rows, int stride, int mode,
                             int cstride, const int32_t* plan, const int32_t* own, uint8_t* raster, int RH, int RW,
                             int margin, void* stream) {
  using namespace fz;
  FZ_REQUIRE(B > 0 && H > 0 && W > 0, "fz_conv3x3_bf16: bad shape");
  FZ_REQUIRE(stride == 1 || stride == 2, "fz_conv3x3_ex: stride %d unsupported", stride);
  FZ_REQUIRE(mode != FZ_CONV_ADD_RELU_BF16 || resid != nullptr, "fz_conv3x3_ex: residual required");
  FZ_REQUIRE(Cin % 16 == 0, "fz_conv3x3_bf16: Cin=%d must be a multiple of 16", Cin);
  if (mode == FZ_CONV_ARGMAX_RASTER) FZ_REQUIRE(plan && raster, "fz_conv3x3_bf16: argmax mode needs plan and raster");
  if (mode == FZ_CONV_LOGITS_F32)
    FZ_REQUIRE(cstride % 4 == 0 && cstride >= Cout && cstride <= 32, "fz_conv3x3_bf16: bad cstride %d", cstride);
  // HBM-bound tail layers (wide maps, few channels): row-streaming kernel, every input row read once
  if (stride == 1 && conv_rows_applicable(H, W, Cin, Cout, mode))
    return conv_rows_launch(in, w, scale, bias, out, B, H, W, Cin, Cout, w_rows, mode, cstride, plan, own, raster, RH,
                            RW, margin, reinterpret_cast<cudaStream_t>(stream));
  const int KC = (Cin % 64 == 0) ? 64 : (Cin % 32 == 0 ? 32 : 16);
  int BN;
  if (mode == FZ_CONV_RELU_BF16 || mode == FZ_CONV_ADD_RELU_BF16 || mode == FZ_CONV_BF16) {
    FZ_REQUIRE(Cout % 16 == 0, "fz_conv3x3_bf16: Cout=%d must be a multiple of 16", Cout);
    BN = (Cout % 128 == 0) ? 128 : (Cout % 64 == 0 ? 64 : (Cout % 32 == 0 ? 32 : 16));
  } else {
    FZ_REQUIRE(Cout >= 1 && Cout <= 32, "fz_conv3x3_bf16: head supports <= 32 classes, got %d", Cout);
    BN = 32;
    if (mode == FZ_CONV_LOGITS_F32)
      FZ_REQUIRE(cstride % 4 == 0 && cstride >= Cout && cstride <= 32, "fz_conv3x3_bf16: bad cstride %d", cstride);
    else if (mode == FZ_CONV_ARGMAX_RASTER)
      FZ_REQUIRE(plan && raster, "fz_conv3x3_bf16: argmax mode needs plan and raster");
  }
  const int n_tiles_n = (Cout + BN - 1) / BN;
  FZ_REQUIRE(w_rows >= n_tiles_n * BN, "fz_conv3x3_bf16: weight rows %d < %d (zero-pad on the host)", w_rows,
             n_tiles_n * BN);
  int TW = W < 128 ? W : 128;
  FZ_REQUIRE(128 % TW == 0, "fz_conv3x3_bf16: W=%d must divide 128 or be a multiple of 128", W);
  int TH = 128 / TW;
  FZ_REQUIRE(W % TW == 0 && H % TH == 0, "fz_conv3x3_bf16: H=%d W=%d not tileable by %dx%d", H, W, TH, TW);

  CUtensorMap tmA, tmB;
  {
    const uint64_t Wi = (uint64_t)W * stride, Hi = (uint64_t)H * stride;     // input extent
    const uint64_t dims[4] = {(uint64_t)Cin, Wi, Hi, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, Wi * Cin * 2, Hi * Wi * Cin * 2};
    // with element strides the box is given in traversed input elements: TW*stride -> TW pixels land in smem
    const uint32_t box[4] = {(uint32_t)KC, (uint32_t)(TW * stride), (uint32_t)(TH * stride), 1};
    const uint32_t estr[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    FZ_REQUIRE(TW * stride <= 256, "fz_conv3x3_ex: tile too wide for a strided TMA box");
    int rc = make_tmap16(&tmA, in, 4, dims, strides, box, KC * 2, estr);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)9 * Cin, (uint64_t)w_rows};
    const uint64_t strides[1] = {(uint64_t)9 * Cin * 2};
    const uint32_t box[2] = {(uint32_t)KC, (uint32_t)BN};
    int rc = make_tmap16(&tmB, w, 2, dims, strides, box, KC * 2);
    if (rc) return rc;
  }
  ConvParams p;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.TW = TW; p.TH = TH;
  p.stride = stride; p.resid = resid;
  p.bias = bias; p.scale = scale; p.out = out; p.cstride = cstride; p.plan = plan; p.own = own; p.raster = raster;
  p.RH = RH; p.RW = RW; p.margin = margin;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (mode) {
    case FZ_CONV_RELU_BF16: return dispatch_conv<FZ_CONV_RELU_BF16>(BN, KC, tmA, tmB, p, st);
    case FZ_CONV_ADD_RELU_BF16: return dispatch_conv<FZ_CONV_ADD_RELU_BF16>(BN, KC, tmA, tmB, p, st);
    case FZ_CONV_BF16: return dispatch_conv<FZ_CONV_BF16>(BN, KC, tmA, tmB, p, st);
    case FZ_CONV_LOGITS_F32: return dispatch_conv<FZ_CONV_LOGITS_F32>(BN, KC, tmA, tmB, p, st);
    case FZ_CONV_LOGITS_F32_NCHW: return dispatch_conv<FZ_CONV_LOGITS_F32_NCHW>(BN, KC, tmA, tmB, p, st);
    case FZ_CONV_ARGMAX_RASTER: return dispatch_conv<FZ_CONV_ARGMAX_RASTER>(BN, KC, tmA, tmB, p, st);
  }
  set_error("fz_conv3x3_bf16: unknown mode %d", mode);
  return -1;
}

// 3x3 / pad 1 convolutions of the U-Net decoder's wide, few-channel layers in the TRAINING step (smp Conv2dReLU and
// SegmentationHead under tasks_module.py:133-167): forward, data gradient and weight gradient without an im2col matrix.
//
// At 256^2 / 512^2 resolution with 16-64 channels these layers are HBM-bound: the explicit im2col of round 1 wrote a
// [pixels][9 C_in] matrix (2.7 GB for one layer), the GEMM read it, the weight gradient read it again and the data gradient
// wrote and re-read a second one -- 13.9 ms of a 81 ms step.  Here a CTA stages a (8+2) x (32+2) pixel tile WITH its halo in
// shared memory once and forms all nine taps from it:
//   forward / data gradient:  out[px][co] = sum_tap sum_ci x[px + tap][ci] * w[tap][co][ci]        (M = pixels, N = co, K = ci)
//   weight gradient:          dw[tap][co][ci] = sum_px dconv[px][co] * x[px + tap][ci]             (M = co, N = ci, K = pixels)
// on warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate): with K = 16-64 per tap and N = 16-32 the arithmetic is a few
// percent of the tensor peak either way; what matters is that every input byte is read from HBM once.  The data gradient is
// the forward kernel run on dconv with the flipped, transposed weights (the host prepares them with the 16-bit copies).
// Shapes outside the template list keep the im2col + tcgen05 GEMM path (engine/convnext_train.py).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.h"
#include "../../include/flair_zonal_b200.h"

namespace fz {

namespace cs {
constexpr int TH = 8, TW = 32, HP = TH + 2, WP = TW + 2;      // tile and halo'd tile, pixels

__device__ __forceinline__ uint32_t sa(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(sa(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(sa(p)));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Stages the halo'd tile of a [B][H][W][C] bf16 map at (b, y0 - 1, x0 - 1): sX[(r * WP + c) * LD + ch], zero outside the image.
// f16_to_bf16: the map holds IEEE fp16 and the tile is wanted in bf16 (the weight gradient multiplies bf16 output gradients
// with fp16 forward activations and mma.sync takes one format): converted on the way in, through registers instead of cp.async
template <int C, int NT>
__device__ __forceinline__ void load_halo_tile(__nv_bfloat16* sX, const __nv_bfloat16* in, int b, int y0, int x0, int H, int W,
                                               bool f16_to_bf16 = false) {
  constexpr int LD = C + 8, C8 = C / 8;
  for (int i = threadIdx.x; i < HP * WP * C8; i += NT) {
    const int c8 = i % C8, pix = i / C8;
    const int r = pix / WP, c = pix - r * WP;
    const int gy = y0 - 1 + r, gx = x0 - 1 + c;
    __nv_bfloat16* dst = sX + pix * LD + c8 * 8;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const __nv_bfloat16* src = in + ((static_cast<size_t>(b) * H + gy) * W + gx) * C + c8 * 8;
      if (f16_to_bf16) {
        const uint4 q = *reinterpret_cast<const uint4*>(src);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint4 o;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
          const __nv_bfloat162 h = __floats2bfloat162_rn(f.x, f.y);
          ow[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(dst) = o;
      } else {
        cp_async16(dst, src);
      }
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}
}  // namespace cs

// ------------------------------------------------------------------------------------------ forward / data gradient
// 8 warps; warp w computes row w of the tile (32 pixels = two 16-pixel M blocks) for all COUT channels.
// w: bf16 [9][COUT][CIN]; out: fp32 or bf16 [pixels][ldo], columns < n_store written (+ bias[col] when given).
// The kernel is latency-bound between its load, compute and store phases (ncu, profiles/r2_ncu_train_conv_summary.txt: two
// resident blocks per SM, issue slots 28 % busy, DRAM 29 %): the 16-output-channel variants (76 registers) are compiled for 3
// resident blocks so that one block's loads overlap the others' MMAs and stores; grids are one wave of resident blocks.
template <int CIN, int COUT, bool OUT16, bool F16>
__global__ void __launch_bounds__(256, COUT == 16 ? 3 : 2) conv3x3_small_fwd_kernel(const __nv_bfloat16* __restrict__ in,
                                                                   const __nv_bfloat16* __restrict__ w,
                                                                   const float* __restrict__ bias, void* __restrict__ out, int B,
                                                                   int H, int W, int n_store, int ldo) {
  using namespace cs;
  constexpr int LDX = CIN + 8, LDW = CIN + 8;
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* sW = reinterpret_cast<__nv_bfloat16*>(smem);                          // [9 * COUT][LDW]
  __nv_bfloat16* sX = sW + 9 * COUT * LDW;                                             // [HP * WP][LDX]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 9 * COUT * (CIN / 8); i += 256) {
    const int c8 = i % (CIN / 8), row = i / (CIN / 8);
    cp_async16(sW + row * LDW + c8 * 8, w + static_cast<size_t>(row) * CIN + c8 * 8);
  }
  const int tiles_x = W / TW, tiles_y = H / TH, n_tiles = B * tiles_y * tiles_x;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
    const int x0 = tx * TW, y0 = ty * TH;
    __syncthreads();                                   // the previous tile's reads of sX are done
    load_halo_tile<CIN, 256>(sX, in, b, y0, x0, H, W);
    cp_async_wait_all();
    __syncthreads();
    float acc[2][COUT / 8][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < COUT / 8; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
#pragma unroll
      for (int ks = 0; ks < CIN / 16; ++ks) {
        uint32_t a[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)                    // A: pixels (rows) x channels, K contiguous
          ldsm_x4(a[m], sX + ((warp + ky) * WP + m * 16 + kx + (lane & 7) + ((lane >> 3) & 1) * 8) * LDX + ks * 16 +
                            (lane >> 4) * 8);
#pragma unroll
        for (int np = 0; np < COUT / 16; ++np) {       // B: two 8-channel N blocks per ldmatrix.x4
          uint32_t bb[4];
          ldsm_x4(bb, sW + (tap * COUT + np * 16 + (lane >> 4) * 8 + (lane & 7)) * LDW + ks * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            if (F16) {
              mma_f16(acc[m][np * 2], a[m], bb[0], bb[1]);
              mma_f16(acc[m][np * 2 + 1], a[m], bb[2], bb[3]);
            } else {
              mma_bf16(acc[m][np * 2], a[m], bb[0], bb[1]);
              mma_bf16(acc[m][np * 2 + 1], a[m], bb[2], bb[3]);
            }
          }
        }
      }
    }
    // D fragment: rows lane / 4 (+8), columns (lane % 4) * 2 (+1) of each 16 x 8 block
    const size_t row_px = (static_cast<size_t>(b) * H + y0 + warp) * W + x0;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const size_t px = row_px + m * 16 + (lane >> 2) + half * 8;
#pragma unroll
        for (int n = 0; n < COUT / 8; ++n) {
          const int col = n * 8 + (lane & 3) * 2;
          if (col < n_store) {
            float v0 = acc[m][n][half * 2], v1 = acc[m][n][half * 2 + 1];
            if (bias) {
              v0 += bias[col];
              v1 += bias[col + 1];
            }
            if (OUT16) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
              *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(out) + px * ldo + col) = h;
            } else {
              *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + px * ldo + col) = make_float2(v0, v1);
            }
          }
        }
      }
  }
  cp_async_wait_all();                                 // a CTA without tiles still has the weight copies in flight
}

// ------------------------------------------------------------------------------------------ weight gradient
// 9 warps, warp = tap.  Per tile the warp walks the 256 pixels in 16 chunks of 16 (K) and accumulates its [COUT][CIN] block in
// registers over ALL the tiles of this CTA; partial[cta][tap][co][ci] is summed over CTAs in order by a second kernel.
template <int CIN, int COUT>
__global__ void __launch_bounds__(288, CIN * COUT <= 512 ? 3 : 2) conv3x3_small_wgrad_kernel(const __nv_bfloat16* __restrict__ x,
                                                                     const __nv_bfloat16* __restrict__ dconv, int ldd,
                                                                     float* __restrict__ partial, int B, int H, int W, int x_f16) {
  using namespace cs;
  constexpr int LDX = CIN + 8, LDD = COUT + 8;
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* sX = reinterpret_cast<__nv_bfloat16*>(smem);                          // [HP * WP][LDX]
  __nv_bfloat16* sD = sX + HP * WP * LDX;                                              // [TH * TW][LDD]
  const int tap = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ky = tap / 3, kx = tap % 3;
  float acc[COUT / 16][CIN / 8][4];
#pragma unroll
  for (int m = 0; m < COUT / 16; ++m)
#pragma unroll
    for (int n = 0; n < CIN / 8; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
  const int tiles_x = W / TW, tiles_y = H / TH, n_tiles = B * tiles_y * tiles_x;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
    const int x0 = tx * TW, y0 = ty * TH;
    __syncthreads();
    load_halo_tile<CIN, 288>(sX, x, b, y0, x0, H, W, x_f16 != 0);
    for (int i = threadIdx.x; i < TH * TW * (COUT / 8); i += 288) {
      const int c8 = i % (COUT / 8), pix = i / (COUT / 8);
      const int r = pix / TW, c = pix - r * TW;
      cp_async16(sD + pix * LDD + c8 * 8, dconv + ((static_cast<size_t>(b) * H + y0 + r) * W + x0 + c) * ldd + c8 * 8);
    }
    cp_async_wait_all();
    __syncthreads();
#pragma unroll 2
    for (int ch = 0; ch < TH * TW / 16; ++ch) {
      const int r = ch >> 1, xo = (ch & 1) * 16;
      uint32_t a[COUT / 16][4];
#pragma unroll
      for (int m = 0; m < COUT / 16; ++m)               // A = dconv^T: stored [pixel][co], transposed on load
        ldsm_x4_t(a[m], sD + (r * TW + xo + (lane & 7) + (lane >> 4) * 8) * LDD + m * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
      for (int np = 0; np < CIN / 16; ++np) {           // B = shifted x: stored [pixel][ci], transposed on load
        uint32_t bb[4];
        ldsm_x4_t(bb, sX + ((r + ky) * WP + xo + kx + (lane & 7) + ((lane >> 3) & 1) * 8) * LDX + (np * 2 + (lane >> 4)) * 8);
#pragma unroll
        for (int m = 0; m < COUT / 16; ++m) {
          mma_bf16(acc[m][np * 2], a[m], bb[0], bb[1]);
          mma_bf16(acc[m][np * 2 + 1], a[m], bb[2], bb[3]);
        }
      }
    }
  }
  float* dst = partial + (static_cast<size_t>(blockIdx.x) * 9 + tap) * COUT * CIN;
#pragma unroll
  for (int m = 0; m < COUT / 16; ++m)
#pragma unroll
    for (int n = 0; n < CIN / 8; ++n)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int co = m * 16 + (lane >> 2) + half * 8, ci = n * 8 + (lane & 3) * 2;
        *reinterpret_cast<float2*>(dst + co * CIN + ci) = make_float2(acc[m][n][half * 2], acc[m][n][half * 2 + 1]);
      }
}

// out[i] = sum_cta partial[cta][i] in CTA order: 32 elements x 8 lanes per block like reduce_rows_kernel
static __global__ void __launch_bounds__(256) conv_small_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                                       int n, int ctas) {
  __shared__ float red[8][33];
  const int i = blockIdx.x * 32 + (threadIdx.x & 31), k = threadIdx.x >> 5;
  float t = 0.f;
  if (i < n)
    for (int s = k; s < ctas; s += 8) t += partial[static_cast<size_t>(s) * n + i];
  red[k][threadIdx.x & 31] = t;
  __syncthreads();
  if (k == 0 && i < n) {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v += red[j][threadIdx.x];
    out[i] = v;
  }
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// resident blocks per SM of a kernel at its launch configuration (asked once per kernel): persistent grids are one wave
template <typename K>
static int resident_blocks(K kernel, int threads, int smem) {
  static int n = 0;
  if (!n) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 1;
  }
  return n;
}

template <int CIN, int COUT>
static int launch_fwd(const void* in, const void* w, const float* bias, void* out, int out16, int in_f16, int B, int H, int W,
                      int n_store, int ldo, cudaStream_t st) {
  constexpr int SMEM = (9 * COUT * (CIN + 8) + cs::HP * cs::WP * (CIN + 8)) * 2;
  const int n_tiles = B * (H / cs::TH) * (W / cs::TW);
  auto a = reinterpret_cast<const __nv_bfloat16*>(in);
  auto b = reinterpret_cast<const __nv_bfloat16*>(w);
#define FZ_CS_LAUNCH(O16, F16)                                                                   \
  {                                                                                              \
    auto k = conv3x3_small_fwd_kernel<CIN, COUT, O16, F16>;                                      \
    FZ_ENSURE_SMEM(k, SMEM);                                                                     \
    const int wave = resident_blocks(k, 256, SMEM) * sm_count();                                 \
    k<<<n_tiles < wave ? n_tiles : wave, 256, SMEM, st>>>(a, b, bias, out, B, H, W, n_store, ldo); \
  }
  // the data gradient (bf16 operands) may write bf16; the forward (fp16 or bf16 operands) writes fp32
  if (out16) FZ_CS_LAUNCH(true, false)
  else if (in_f16) FZ_CS_LAUNCH(false, true)
  else FZ_CS_LAUNCH(false, false)
#undef FZ_CS_LAUNCH
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int CIN, int COUT>
static int launch_wgrad(const void* x, const void* dconv, int ldd, float* dw, float* partial, int max_ctas, int B, int H, int W,
                        int x_f16, cudaStream_t st) {
  constexpr int SMEM = (cs::HP * cs::WP * (CIN + 8) + cs::TH * cs::TW * (COUT + 8)) * 2;
  auto k = conv3x3_small_wgrad_kernel<CIN, COUT>;
  FZ_ENSURE_SMEM(k, SMEM);
  const int wave = resident_blocks(k, 288, SMEM) * sm_count();
  const int ctas = max_ctas < wave ? max_ctas : wave;                 // the scratch buffer is sized for max_ctas
  k<<<ctas, 288, SMEM, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dconv), ldd,
                             partial, B, H, W, x_f16);
  const int n = 9 * COUT * CIN;
  conv_small_reduce_kernel<<<(n + 31) / 32, 256, 0, st>>>(partial, dw, n, ctas);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static bool small_channels(int c) { return c == 16 || c == 32 || c == 48 || c == 64; }

}  // namespace fz

using namespace fz;

extern "C" int fz_conv3x3_small_supported(int H, int W, int Cin, int Cout) {
  return H > 0 && W > 0 && H % cs::TH == 0 && W % cs::TW == 0 && small_channels(Cin) && small_channels(Cout) ? 1 : 0;
}

extern "C" int fz_conv3x3_small_forward(const void* in_bf16, const void* w_bf16, const float* bias, void* out, int out_bf16,
                                        int B, int H, int W, int Cin, int Cout, int n_store, int ldo, int in_f16, void* stream) {
  FZ_REQUIRE(!(in_f16 && out_bf16), "fz_conv3x3_small_forward: fp16 operands write fp32 only");
  FZ_REQUIRE(B > 0 && in_bf16 && w_bf16 && out && fz_conv3x3_small_supported(H, W, Cin, Cout),
             "fz_conv3x3_small_forward: B=%d H=%d W=%d Cin=%d Cout=%d not covered (H %% 8, W %% 32, channels 16/32/48/64)", B, H,
             W, Cin, Cout);
  FZ_REQUIRE(n_store >= 1 && n_store <= Cout && n_store % 2 == 0 && ldo >= n_store && ldo % 2 == 0,
             "fz_conv3x3_small_forward: n_store=%d ldo=%d (even, n_store <= Cout, ldo >= n_store)", n_store, ldo);
  FZ_REQUIRE(static_cast<int64_t>(B) * (H / cs::TH) * (W / cs::TW) < (1LL << 31), "fz_conv3x3_small_forward: too many tiles");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define FZ_CS(ci, co) \
  if (Cin == ci && Cout == co) return launch_fwd<ci, co>(in_bf16, w_bf16, bias, out, out_bf16, in_f16, B, H, W, n_store, ldo, st);
  FZ_CS(16, 16) FZ_CS(16, 32) FZ_CS(16, 48) FZ_CS(16, 64) FZ_CS(32, 16) FZ_CS(32, 32) FZ_CS(32, 48) FZ_CS(32, 64)
  FZ_CS(48, 16) FZ_CS(48, 32) FZ_CS(48, 48) FZ_CS(48, 64) FZ_CS(64, 16) FZ_CS(64, 32) FZ_CS(64, 48) FZ_CS(64, 64)
#undef FZ_CS
  set_error("fz_conv3x3_small_forward: no kernel for Cin=%d Cout=%d", Cin, Cout);
  return -1;
}

extern "C" int fz_conv3x3_small_wgrad(const void* x_bf16, const void* dconv_bf16, int ldd, float* dw, int B, int H, int W,
                                      int Cin, int Cout, int x_f16, void* stream) {
  FZ_REQUIRE(B > 0 && x_bf16 && dconv_bf16 && dw && fz_conv3x3_small_supported(H, W, Cin, Cout) && Cout <= 32,
             "fz_conv3x3_small_wgrad: B=%d H=%d W=%d Cin=%d Cout=%d not covered (H %% 8, W %% 32, Cin 16/32/48/64, Cout 16/32)", B,
             H, W, Cin, Cout);
  FZ_REQUIRE(ldd >= Cout && ldd % 8 == 0, "fz_conv3x3_small_wgrad: ldd=%d (>= Cout, multiple of 8)", ldd);
  const int64_t n_tiles = static_cast<int64_t>(B) * (H / cs::TH) * (W / cs::TW);
  FZ_REQUIRE(n_tiles < (1LL << 31), "fz_conv3x3_small_wgrad: too many tiles");
  const int ctas = n_tiles < 4 * sm_count() ? static_cast<int>(n_tiles) : 4 * sm_count();      // upper bound; launch_wgrad trims to one wave
  static float* scratch[64] = {nullptr};              // partial sums [ctas][9][Cout][Cin]: per-device, grown on demand
  static size_t have[64] = {0};
  int dev = 0;
  FZ_CHECK_CUDA(cudaGetDevice(&dev));
  FZ_REQUIRE(dev >= 0 && dev < 64, "fz_conv3x3_small_wgrad: device index %d", dev);
  const size_t need = static_cast<size_t>(ctas) * 9 * Cout * Cin;
  if (have[dev] < need) {
    if (scratch[dev]) cudaFree(scratch[dev]);
    scratch[dev] = nullptr;
    have[dev] = 0;
    FZ_CHECK_CUDA(cudaMalloc(&scratch[dev], need * sizeof(float)));
    have[dev] = need;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define FZ_CS(ci, co) \
  if (Cin == ci && Cout == co) return launch_wgrad<ci, co>(x_bf16, dconv_bf16, ldd, dw, scratch[dev], ctas, B, H, W, x_f16, st);
  FZ_CS(16, 16) FZ_CS(32, 16) FZ_CS(48, 16) FZ_CS(64, 16) FZ_CS(16, 32) FZ_CS(32, 32) FZ_CS(48, 32) FZ_CS(64, 32)
#undef FZ_CS
  set_error("fz_conv3x3_small_wgrad: no kernel for Cin=%d Cout=%d", Cin, Cout);
  return -1;
}

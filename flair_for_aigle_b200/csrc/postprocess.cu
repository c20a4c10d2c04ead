// Memory-bound post-processing kernels: margin crop + argmax / softmax + overlap accumulate,
// written straight into the zone raster in HBM.
//
// Replaces flair_zonal_detection/inference.py:295-352 (20 MB/tile fp32 D2H, numpy crop,
// postprocess.convert, rasterio windowed write) and postprocess.py:9-30; the accumulating
// variant implements the intended semantics of inference.py:468-572.
//
// Access pattern: one thread per output pixel (4 px per thread on the aligned NCHW path);
// NCHW logits are read plane by plane so every warp load is a contiguous 128/512 B run;
// NHWC logits (the engine's own layout, class fastest, 16 B-aligned records) are read with
// 16 B vector loads and reduced in registers.
#include "common.h"
#include "../../include/flair_zonal_b200.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace fz {

constexpr int MAX_CLS = 32;

struct TileWin {
  int y0, x0;      // first tile-local pixel (inside the P x P tile)
  int r0, c0;      // raster pixel it maps to
  int h, w;        // extent
};

// Resolve tile t's window: write window from the plan, optionally trimmed to the owned sub-window.
__device__ __forceinline__ TileWin tile_window(const int32_t* plan, const int32_t* own, int t, int margin) {
  const int32_t* p = plan + 6 * t;
  const int top = p[2], left = p[3];
  int r0 = top, r1 = top + p[4], c0 = left, c1 = left + p[5];
  if (own) {
    const int32_t* o = own + 4 * t;
    r0 = max(r0, o[0]);
    r1 = min(r1, o[1]);
    c0 = max(c0, o[2]);
    c1 = min(c1, o[3]);
  }
  TileWin w;
  w.r0 = r0;
  w.c0 = c0;
  w.h = r1 - r0;
  w.w = c1 - c0;
  w.y0 = margin + (r0 - top);
  w.x0 = margin + (c0 - left);
  return w;
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// Load the n_cls logits of one pixel into registers.
template <typename T, int LAYOUT>
__device__ __forceinline__ void load_pixel(const T* __restrict__ logits, int t, int n_cls, int cstride, int P, int y,
                                           int x, float (&v)[MAX_CLS]) {
  if (LAYOUT == FZ_NCHW) {
    const T* base = logits + (static_cast<size_t>(t) * n_cls * P + y) * P + x;
    const size_t plane = static_cast<size_t>(P) * P;
#pragma unroll
    for (int c = 0; c < MAX_CLS; ++c)
      if (c < n_cls) v[c] = to_f32<T>(base[c * plane]);
  } else if (LAYOUT == FZ_NHWC_UP4) {
    // logits live at quarter resolution [P/4][P/4][cstride]; nn.UpsamplingBilinear2d(scale_factor=4)
    // (align_corners=True) of smp's UPerNet head is evaluated here, with fz_head_upsample4's exact arithmetic
    const int h = P / 4;
    const float sc = static_cast<float>(h - 1) / static_cast<float>(P - 1);
    const float fy = y * sc, fx = x * sc;
    int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    y0 = y0 > h - 1 ? h - 1 : y0;
    x0 = x0 > h - 1 ? h - 1 : x0;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < h - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const T* tb = logits + static_cast<size_t>(t) * h * h * cstride;
    const T* p00 = tb + (static_cast<size_t>(y0) * h + x0) * cstride;
    const T* p01 = tb + (static_cast<size_t>(y0) * h + x1) * cstride;
    const T* p10 = tb + (static_cast<size_t>(y1) * h + x0) * cstride;
    const T* p11 = tb + (static_cast<size_t>(y1) * h + x1) * cstride;
#pragma unroll
    for (int c = 0; c < MAX_CLS; ++c)
      if (c < n_cls)
        v[c] = hy * (hx * to_f32<T>(p00[c]) + lx * to_f32<T>(p01[c])) + ly * (hx * to_f32<T>(p10[c]) + lx * to_f32<T>(p11[c]));
  } else {
    const T* base = logits + ((static_cast<size_t>(t) * P + y) * P + x) * cstride;
    if ((static_cast<size_t>(cstride) * sizeof(T)) % 16 == 0) {
      constexpr int PER = 16 / sizeof(T);
#pragma unroll
      for (int c = 0; c < MAX_CLS; c += PER) {
        if (c < n_cls) {
          const uint4 raw = *reinterpret_cast<const uint4*>(base + c);
          const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
          for (int j = 0; j < PER; ++j) v[c + j] = to_f32<T>(e[j]);
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < MAX_CLS; ++c)
        if (c < n_cls) v[c] = to_f32<T>(base[c]);
    }
  }
}

__device__ __forceinline__ int argmax_first(const float (&v)[MAX_CLS], int n_cls) {
  // np.argmax: first maximal index wins; NaN handling is not needed (finite logits).
  int best = 0;
  float bv = v[0];
#pragma unroll
  for (int c = 1; c < MAX_CLS; ++c)
    if (c < n_cls && v[c] > bv) {
      bv = v[c];
      best = c;
    }
  return best;
}

// softmax over the classes of one pixel.  ex2.approx (2^-22 relative) and ONE reciprocal instead of 19 expf() + 19 IEEE
// divisions: the class_prob / accumulate kernels were instruction-bound, not memory-bound (round 2: 576 GB/s with expf and
// '/'), and the result feeds a uint8 (1/255 steps) or an fp32 canvas compared at 2e-6 absolute.
__device__ __forceinline__ void softmax_inplace(float (&v)[MAX_CLS], int n_cls) {
  float mx = v[0];
#pragma unroll
  for (int c = 1; c < MAX_CLS; ++c)
    if (c < n_cls) mx = fmaxf(mx, v[c]);
  float sum = 0.0f;
  constexpr float LOG2E = 1.4426950408889634f;
  const float mxl = mx * LOG2E;
#pragma unroll
  for (int c = 0; c < MAX_CLS; ++c)
    if (c < n_cls) {
      float e;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v[c], LOG2E, -mxl)));
      v[c] = e;
      sum += e;
    }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int c = 0; c < MAX_CLS; ++c)
    if (c < n_cls) v[c] = v[c] * inv;
}

constexpr int PP_THREADS = 128;
constexpr int PP_ROWS = 4;  // rows of the window per CTA

// Which tiles of the call a launch covers: n == 0 -> tile = blockIdx.x; else tile = idx[blockIdx.x] (a conflict-free group of
// the accumulating variant, passed by value: no device-side index buffer, capturable in a CUDA graph).
constexpr int TL_MAX = 96;
struct TileList {
  int n;
  uint8_t idx[TL_MAX];
};

// MODE 0: argmax -> uint8 [H][W]; 1: round(softmax*255) -> uint8 [n_cls][H][W];
// 2: canvas[n_cls][H][W] += w * softmax
template <typename T, int LAYOUT, int MODE>
__global__ void __launch_bounds__(PP_THREADS) crop_kernel(const T* __restrict__ logits, int n_cls, int cstride, int P,
                                                          int margin, const int32_t* __restrict__ plan,
                                                          const int32_t* __restrict__ own,
                                                          const float* __restrict__ weight, uint8_t* __restrict__ out8,
                                                          float* __restrict__ canvas, int H, int W,
                                                          const int32_t* __restrict__ zmap, const TileList tl) {
  const int t = tl.n ? tl.idx[blockIdx.x] : blockIdx.x;
  const TileWin win = tile_window(plan, own, t, margin);
  if (win.h <= 0 || win.w <= 0) return;
  const int row_begin = blockIdx.y * PP_ROWS;
  const size_t plane = static_cast<size_t>(H) * W;
  const int S = P - 2 * margin;
  // zmap (output_px_meters != reference resolution): nearest-neighbour zoom of the cropped prediction,
  // inference.py:212-226,303-312 -- output offset o inside the tile's window reads crop pixel zmap[o]
  for (int ry = row_begin; ry < min(row_begin + PP_ROWS, win.h); ++ry) {
    const int zy = zmap ? zmap[win.y0 - margin + ry] : 0;     // < 0: scipy's constant fill (zero logits / label 0)
    const int y = zmap ? margin + zy : win.y0 + ry;
    const size_t orow = static_cast<size_t>(win.r0 + ry) * W + win.c0;
    for (int rx = threadIdx.x; rx < win.w; rx += PP_THREADS) {
      const int zx = zmap ? zmap[win.x0 - margin + rx] : 0;
      const int xs = zmap ? margin + zx : win.x0 + rx;
      float v[MAX_CLS];
      if (zy < 0 || zx < 0) {
#pragma unroll
        for (int c = 0; c < MAX_CLS; ++c) v[c] = 0.0f;
      } else {
        load_pixel<T, LAYOUT>(logits, t, n_cls, cstride, P, y, xs, v);
      }
      if (MODE == 0) {
        out8[orow + rx] = static_cast<uint8_t>(argmax_first(v, n_cls));
      } else {
        softmax_inplace(v, n_cls);
        if (MODE == 1) {
#pragma unroll
          for (int c = 0; c < MAX_CLS; ++c)
            if (c < n_cls) out8[c * plane + orow + rx] = static_cast<uint8_t>(rintf(v[c] * 255.0f));
        } else {
          const float wgt = weight ? weight[(y - margin) * S + (xs - margin)] : 1.0f;
#pragma unroll
          for (int c = 0; c < MAX_CLS; ++c)
            if (c < n_cls) canvas[c * plane + orow + rx] += wgt * v[c];
        }
      }
    }
  }
}

// fp32 logits, no zoom: 4 consecutive pixels per thread -- float4 loads of every class plane (NCHW) and 4-byte / 16-byte
// stores per output plane instead of 19 single-byte (MODE 1) or 4-byte (MODE 2) accesses per pixel.  Round 2 measured the
// one-pixel-per-thread kernel at 576 GB/s (class_prob planes) and 183 GB/s (accumulate, one launch per tile).
// MODE 1: round(softmax * 255) -> uint8 [n_cls][H][W];  MODE 2: canvas[n_cls][H][W] += w * softmax.
constexpr int PV_THREADS = 128;
template <int LAYOUT, int MODE>
__global__ void __launch_bounds__(PV_THREADS) crop_vec4_kernel(const float* __restrict__ logits, int n_cls, int cstride, int P,
                                                               int margin, const int32_t* __restrict__ plan,
                                                               const int32_t* __restrict__ own,
                                                               const float* __restrict__ weight, uint8_t* __restrict__ out8,
                                                               float* __restrict__ canvas, int H, int W, const TileList tl) {
  const int t = tl.n ? tl.idx[blockIdx.x] : blockIdx.x;
  const TileWin win = tile_window(plan, own, t, margin);
  if (win.h <= 0 || win.w <= 0) return;
  const int gw = (win.w + 3) >> 2;                      // groups of 4 pixels per row
  const int g = blockIdx.y * PV_THREADS + threadIdx.x;
  if (g >= gw * win.h) return;
  const int ry = g / gw, rx = (g - ry * gw) * 4;
  const int npx = min(4, win.w - rx);
  const int y = win.y0 + ry, x = win.x0 + rx;
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t opix = static_cast<size_t>(win.r0 + ry) * W + win.c0 + rx;
  const int S = P - 2 * margin;
  float v[4][MAX_CLS];
  const size_t iplane = static_cast<size_t>(P) * P;
  const float* ibase = logits + (static_cast<size_t>(t) * n_cls * P + y) * P + x;      // NCHW
  const bool vec_in = LAYOUT == FZ_NCHW && npx == 4 && ((reinterpret_cast<uintptr_t>(ibase) & 15) == 0) && (iplane % 4 == 0);
  if (vec_in) {
#pragma unroll
    for (int c = 0; c < MAX_CLS; ++c)
      if (c < n_cls) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(ibase + c * iplane));
        v[0][c] = q.x; v[1][c] = q.y; v[2][c] = q.z; v[3][c] = q.w;
      }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < npx) load_pixel<float, LAYOUT>(logits, t, n_cls, cstride, P, y, x + j, v[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < npx) softmax_inplace(v[j], n_cls);
  if (MODE == 1) {
    const bool vec_out = npx == 4 && ((opix & 3) == 0) && (plane % 4 == 0);
#pragma unroll
    for (int c = 0; c < MAX_CLS; ++c)
      if (c < n_cls) {
        uint8_t* o = out8 + c * plane + opix;
        if (vec_out) {
          *reinterpret_cast<uchar4*>(o) = make_uchar4(static_cast<uint8_t>(rintf(v[0][c] * 255.0f)), static_cast<uint8_t>(rintf(v[1][c] * 255.0f)),
                                                      static_cast<uint8_t>(rintf(v[2][c] * 255.0f)), static_cast<uint8_t>(rintf(v[3][c] * 255.0f)));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < npx) o[j] = static_cast<uint8_t>(rintf(v[j][c] * 255.0f));
        }
      }
  } else {
    float wg[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) wg[j] = (weight && j < npx) ? weight[(y - margin) * S + (x + j - margin)] : 1.0f;
    const bool vec_out = npx == 4 && ((opix & 3) == 0) && (plane % 4 == 0);
#pragma unroll
    for (int c = 0; c < MAX_CLS; ++c)
      if (c < n_cls) {
        float* o = canvas + c * plane + opix;
        if (vec_out) {
          float4 a = *reinterpret_cast<float4*>(o);
          a.x += wg[0] * v[0][c]; a.y += wg[1] * v[1][c]; a.z += wg[2] * v[2][c]; a.w += wg[3] * v[3][c];
          *reinterpret_cast<float4*>(o) = a;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < npx) o[j] += wg[j] * v[j][c];
        }
      }
  }
}

template <int MODE>
static int launch_crop(const void* logits, int dtype, int layout, int cstride, int n_tiles, int n_cls, int P,
                       int margin, const int32_t* plan, const int32_t* own, const float* weight, uint8_t* out8,
                       float* canvas, int H, int W, cudaStream_t st, const int32_t* zmap = nullptr, int zoomed = 0,
                       const TileList* tiles = nullptr) {
  FZ_REQUIRE(n_cls >= 1 && n_cls <= MAX_CLS, "crop kernels support 1..%d classes, got %d", MAX_CLS, n_cls);
  FZ_REQUIRE(P > 2 * margin && margin >= 0, "bad patch/margin %d/%d", P, margin);
  FZ_REQUIRE(dtype == FZ_F32 || dtype == FZ_BF16 || dtype == FZ_F16, "bad dtype %d", dtype);
  FZ_REQUIRE(layout == FZ_NCHW || layout == FZ_NHWC || layout == FZ_NHWC_UP4, "bad layout %d", layout);
  if (layout == FZ_NHWC_UP4)
    FZ_REQUIRE(dtype == FZ_F32 && P % 4 == 0 && P >= 8 && cstride >= n_cls, "quarter-resolution logits: fp32, P %% 4 == 0");
  if (layout == FZ_NHWC) FZ_REQUIRE(cstride >= n_cls, "cstride %d < n_cls %d", cstride, n_cls);
  if (layout == FZ_NHWC && (static_cast<size_t>(cstride) * (dtype == FZ_F32 ? 4 : 2)) % 16 == 0)
    FZ_REQUIRE(cstride >= ((n_cls + (dtype == FZ_F32 ? 3 : 7)) / (dtype == FZ_F32 ? 4 : 8)) * (dtype == FZ_F32 ? 4 : 8),
               "vector path needs cstride >= n_cls rounded up to 16 B");
  if (n_tiles <= 0) return 0;
  const int S = zmap ? zoomed : P - 2 * margin;      // rows a tile's write window can have
  FZ_REQUIRE(S >= 1, "bad zoomed window size %d", S);
  TileList tl;
  tl.n = 0;
  if (tiles) tl = *tiles;
  const int n_launch = tl.n ? tl.n : n_tiles;
  if ((MODE == 1 || MODE == 2) && dtype == FZ_F32 && zmap == nullptr && (layout == FZ_NCHW || layout == FZ_NHWC)) {
    // fp32 logits without zoom: the 4-pixels-per-thread kernel
    dim3 vgrid(n_launch, (((S + 3) / 4) * S + PV_THREADS - 1) / PV_THREADS);
    if (layout == FZ_NCHW)
      crop_vec4_kernel<FZ_NCHW, MODE><<<vgrid, PV_THREADS, 0, st>>>(static_cast<const float*>(logits), n_cls, cstride, P, margin,
                                                                    plan, own, weight, out8, canvas, H, W, tl);
    else
      crop_vec4_kernel<FZ_NHWC, MODE><<<vgrid, PV_THREADS, 0, st>>>(static_cast<const float*>(logits), n_cls, cstride, P, margin,
                                                                    plan, own, weight, out8, canvas, H, W, tl);
    FZ_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  dim3 grid(n_launch, (S + PP_ROWS - 1) / PP_ROWS), block(PP_THREADS);
#define FZ_LAUNCH(T, LAY)                                                                                       \
  crop_kernel<T, LAY, MODE><<<grid, block, 0, st>>>(reinterpret_cast<const T*>(logits), n_cls, cstride, P, margin, \
                                                    plan, own, weight, out8, canvas, H, W, zmap, tl)
  if (layout == FZ_NHWC_UP4) FZ_LAUNCH(float, FZ_NHWC_UP4);
  else if (dtype == FZ_F32 && layout == FZ_NCHW) FZ_LAUNCH(float, FZ_NCHW);
  else if (dtype == FZ_F32) FZ_LAUNCH(float, FZ_NHWC);
  else if (dtype == FZ_BF16 && layout == FZ_NCHW) FZ_LAUNCH(__nv_bfloat16, FZ_NCHW);
  else if (dtype == FZ_BF16) FZ_LAUNCH(__nv_bfloat16, FZ_NHWC);
  else if (layout == FZ_NCHW) FZ_LAUNCH(__half, FZ_NCHW);
  else FZ_LAUNCH(__half, FZ_NHWC);
#undef FZ_LAUNCH
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void canvas_argmax_kernel(const float* __restrict__ canvas, int n_cls, int64_t n_px,
                                     uint8_t* __restrict__ labels, float* __restrict__ conf) {
  // 4 pixels per thread, float4 plane loads (n_px is padded by the tail loop below).
  const int64_t i4 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n_px) return;
  if (i4 + 4 <= n_px && (n_px % 4 == 0)) {
    float4 best = *reinterpret_cast<const float4*>(canvas + i4);
    uchar4 idx = make_uchar4(0, 0, 0, 0);
    for (int c = 1; c < n_cls; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(canvas + c * n_px + i4);
      if (v.x > best.x) { best.x = v.x; idx.x = c; }
      if (v.y > best.y) { best.y = v.y; idx.y = c; }
      if (v.z > best.z) { best.z = v.z; idx.z = c; }
      if (v.w > best.w) { best.w = v.w; idx.w = c; }
    }
    *reinterpret_cast<uchar4*>(labels + i4) = idx;
    if (conf) *reinterpret_cast<float4*>(conf + i4) = best;
  } else {
    for (int64_t i = i4; i < min(i4 + 4, n_px); ++i) {
      float best = canvas[i];
      int idx = 0;
      for (int c = 1; c < n_cls; ++c) {
        const float v = canvas[c * n_px + i];
        if (v > best) { best = v; idx = c; }
      }
      labels[i] = static_cast<uint8_t>(idx);
      if (conf) conf[i] = best;
    }
  }
}

__global__ void convert_kernel(const float* __restrict__ img, int C, int64_t n_px, int mode,
                               uint8_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  float v[MAX_CLS];
#pragma unroll
  for (int c = 0; c < MAX_CLS; ++c)
    if (c < C) v[c] = img[c * n_px + i];
  if (mode == 0) {
    out[i] = static_cast<uint8_t>(argmax_first(v, C));
  } else {
    softmax_inplace(v, C);
#pragma unroll
    for (int c = 0; c < MAX_CLS; ++c)
      if (c < C) out[c * n_px + i] = static_cast<uint8_t>(rintf(v[c] * 255.0f));
  }
}

}  // namespace fz

extern "C" int fz_crop_argmax_write(const void* logits, int dtype, int layout, int cstride, int n_tiles, int n_cls,
                                    int P, int margin, const int32_t* plan, const int32_t* own, uint8_t* out_raster,
                                    int H, int W, void* stream) {
  return fz::launch_crop<0>(logits, dtype, layout, cstride, n_tiles, n_cls, P, margin, plan, own, nullptr, out_raster,
                            nullptr, H, W, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int fz_crop_softmax_write(const void* logits, int dtype, int layout, int cstride, int n_tiles, int n_cls,
                                     int P, int margin, const int32_t* plan, const int32_t* own, uint8_t* out_raster,
                                     int H, int W, void* stream) {
  return fz::launch_crop<1>(logits, dtype, layout, cstride, n_tiles, n_cls, P, margin, plan, own, nullptr, out_raster,
                            nullptr, H, W, reinterpret_cast<cudaStream_t>(stream));
}

namespace fz {
// Tiles of one call may overlap in the canvas (clamped edge rows/columns): serialise them.
// The grid is a product grid, so tiles that overlap are never in the same call when the
// caller batches by grid column; to stay safe for any batch we launch tile by tile.
static int accumulate_tiles(const void* logits, int dtype, int layout, int cstride, int n_tiles, int n_cls, int P,
                            int margin, const int32_t* plan, const float* weight, float* canvas, int H, int W,
                            cudaStream_t st, const int32_t* zmap, int zoomed, const int32_t* plan_host) {
  if (plan_host != nullptr && n_tiles > 1 && n_tiles <= TL_MAX) {
    // Overlapping write windows must be accumulated one after the other (no atomics: the canvas is bit-reproducible), but
    // most windows of a tile grid are disjoint.  Level the tiles: level(t) = 1 + max level of an EARLIER tile whose window
    // intersects t's; tiles of one level never overlap and go into ONE launch, levels run in order -- so every pixel still
    // receives its contributions in tile order, i.e. the result equals the one-launch-per-tile sequence bit for bit.
    const int S = zmap ? zoomed : P - 2 * margin;
    int level[TL_MAX], n_levels = 0;
    for (int t = 0; t < n_tiles; ++t) {
      const int32_t* pt = plan_host + 6 * t;
      const int th = zmap ? pt[4] : (pt[4] < S ? pt[4] : S), tw = zmap ? pt[5] : (pt[5] < S ? pt[5] : S);
      int lv = 0;
      for (int u = 0; u < t; ++u) {
        const int32_t* pu = plan_host + 6 * u;
        const int uh = zmap ? pu[4] : (pu[4] < S ? pu[4] : S), uw = zmap ? pu[5] : (pu[5] < S ? pu[5] : S);
        const bool apart = pt[2] + th <= pu[2] || pu[2] + uh <= pt[2] || pt[3] + tw <= pu[3] || pu[3] + uw <= pt[3];
        if (!apart && th > 0 && tw > 0 && uh > 0 && uw > 0 && level[u] + 1 > lv) lv = level[u] + 1;
      }
      level[t] = lv;
      if (lv + 1 > n_levels) n_levels = lv + 1;
    }
    for (int lv = 0; lv < n_levels; ++lv) {
      TileList tl;
      tl.n = 0;
      for (int t = 0; t < n_tiles; ++t)
        if (level[t] == lv) tl.idx[tl.n++] = static_cast<uint8_t>(t);
      int rc = launch_crop<2>(logits, dtype, layout, cstride, n_tiles, n_cls, P, margin, plan, nullptr, weight, nullptr, canvas,
                              H, W, st, zmap, zoomed, &tl);
      if (rc) return rc;
    }
    return 0;
  }
  for (int t = 0; t < n_tiles; ++t) {
    int rc = launch_crop<2>(logits, dtype, layout, cstride, 1, n_cls, P, margin, plan, nullptr, weight, nullptr, canvas, H,
                            W, st, zmap, zoomed);
    if (rc) return rc;
    // advance to the next tile: logits and plan are indexed by blockIdx.x inside the kernel
    const size_t esz = dtype == FZ_F32 ? 4 : 2;
    const size_t per_tile = layout == FZ_NCHW ? static_cast<size_t>(n_cls) * P * P
                            : layout == FZ_NHWC_UP4 ? static_cast<size_t>(P / 4) * (P / 4) * cstride
                                                    : static_cast<size_t>(P) * P * cstride;
    logits = static_cast<const char*>(logits) + per_tile * esz;
    plan += 6;
  }
  return 0;
}
}  // namespace fz

extern "C" int fz_crop_softmax_accumulate(const void* logits, int dtype, int layout, int cstride, int n_tiles,
                                          int n_cls, int P, int margin, const int32_t* plan, const int32_t* plan_host,
                                          const float* weight, float* canvas, int H, int W, void* stream) {
  return fz::accumulate_tiles(logits, dtype, layout, cstride, n_tiles, n_cls, P, margin, plan, weight, canvas, H, W,
                              reinterpret_cast<cudaStream_t>(stream), nullptr, 0, plan_host);
}

extern "C" int fz_crop_zoom_accumulate(const void* logits, int dtype, int layout, int cstride, int n_tiles, int n_cls,
                                       int P, int margin, const int32_t* plan, const int32_t* plan_host, const int32_t* zmap,
                                       int zoomed, float* canvas, int H, int W, void* stream) {
  FZ_REQUIRE(zmap != nullptr && zoomed >= 1, "fz_crop_zoom_accumulate: zoom map required");
  return fz::accumulate_tiles(logits, dtype, layout, cstride, n_tiles, n_cls, P, margin, plan, nullptr, canvas, H, W,
                              reinterpret_cast<cudaStream_t>(stream), zmap, zoomed, plan_host);
}

extern "C" int fz_canvas_argmax(const float* canvas, int n_cls, int64_t n_px, uint8_t* labels, float* confidence,
                                void* stream) {
  FZ_REQUIRE(n_cls >= 1 && n_px >= 0, "fz_canvas_argmax: bad arguments");
  if (n_px == 0) return 0;
  const int threads = 256;
  const int64_t blocks = (n_px + threads * 4 - 1) / (threads * 4);
  fz::canvas_argmax_kernel<<<static_cast<unsigned>(blocks), threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      canvas, n_cls, n_px, labels, confidence);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_convert(const float* img, int C, int h, int w, int mode, uint8_t* out, void* stream) {
  FZ_REQUIRE(C >= 1 && C <= fz::MAX_CLS, "fz_convert: supports 1..%d classes, got %d", fz::MAX_CLS, C);
  FZ_REQUIRE(mode == 0 || mode == 1, "fz_convert: Unknown output type: %d", mode);
  const int64_t n_px = static_cast<int64_t>(h) * w;
  if (n_px == 0) return 0;
  const int threads = 256;
  fz::convert_kernel<<<static_cast<unsigned>((n_px + threads - 1) / threads), threads, 0,
                       reinterpret_cast<cudaStream_t>(stream)>>>(img, C, n_px, mode, out);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_crop_zoom_write(int mode, const void* logits, int dtype, int layout, int cstride, int n_tiles, int n_cls,
                                  int P, int margin, const int32_t* plan, const int32_t* own, const int32_t* zmap,
                                  int zoomed, uint8_t* out_raster, int H, int W, void* stream) {
  using namespace fz;
  FZ_REQUIRE(zmap != nullptr && zoomed >= 1, "fz_crop_zoom_write: zoom map required");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (mode == 0)
    return launch_crop<0>(logits, dtype, layout, cstride, n_tiles, n_cls, P, margin, plan, own, nullptr, out_raster,
                          nullptr, H, W, st, zmap, zoomed);
  if (mode == 1)
    return launch_crop<1>(logits, dtype, layout, cstride, n_tiles, n_cls, P, margin, plan, own, nullptr, out_raster,
                          nullptr, H, W, st, zmap, zoomed);
  set_error("fz_crop_zoom_write: mode %d (0 argmax, 1 class_prob)", mode);
  return -1;
}

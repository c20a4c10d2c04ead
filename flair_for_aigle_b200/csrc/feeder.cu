// Tile feeder kernels: boundless windowed read of the zone raster (zero fill) + normalisation.
//
// Replaces flair_zonal_detection/dataset.py:89-117 (_load_patch: rasterio boundless read,
// fill_value=0), :119-124 + flair_hub/data/utils_data/norm.py:37-44 ((x-mean)/std in
// float64 -> float32) and the per-window Python loop of the DataLoader workers.  The raster
// stays resident in HBM as uint8; a tile is 1 byte/channel/pixel read.
#include "common.h"
#include "../../include/flair_zonal_b200.h"

namespace fz {

// out[t][c][y][x] = (raster[c][row0+y][col0+x] (0 outside) - mean[c]) / std[c]
// One thread = 4 consecutive x of one (t, c, y): 4 byte loads (coalesced across the warp),
// one float4 store.
template <typename T>
__global__ void gather_f32_kernel(const T* __restrict__ raster, int C, int H, int W,
                                  const int32_t* __restrict__ origins, int P, const float* __restrict__ mean,
                                  const float* __restrict__ stdv, float* __restrict__ out) {
  const int t = blockIdx.z;
  const int c = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over P*P/4
  const int per_row = P / 4;
  if (idx >= P * per_row) return;
  const int y = idx / per_row;
  const int x = (idx % per_row) * 4;
  const int r = origins[2 * t] + y;
  const int c0 = origins[2 * t + 1] + x;
  // float64 like norm.py:40-43, rounded once to float32 (torch.tensor(..., float32))
  const double m = static_cast<double>(mean[c]);
  const double s = static_cast<double>(stdv[c]);
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cc = c0 + j;
    const double px = (r >= 0 && r < H && cc >= 0 && cc < W)
                          ? static_cast<double>(raster[(static_cast<size_t>(c) * H + r) * W + cc])
                          : 0.0;
    v[j] = static_cast<float>((px - m) / s);
  }
  float4* dst = reinterpret_cast<float4*>(out + ((static_cast<size_t>(t) * C + c) * P + y) * P + x);
  *dst = make_float4(v[0], v[1], v[2], v[3]);
}

// out[t][y][x][0..3] = raster[c][row0+y][col0+x] (0 outside, 0 for c >= C)
__global__ void gather_u8_kernel(const uint8_t* __restrict__ raster, int C, int H, int W,
                                 const int32_t* __restrict__ origins, int P, uint8_t* __restrict__ out) {
  const int t = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over P*P
  if (idx >= P * P) return;
  const int y = idx / P, x = idx % P;
  const int r = origins[2 * t] + y;
  const int cc = origins[2 * t + 1] + x;
  uchar4 px = make_uchar4(0, 0, 0, 0);
  if (r >= 0 && r < H && cc >= 0 && cc < W) {
    const size_t o = static_cast<size_t>(r) * W + cc;
    const size_t plane = static_cast<size_t>(H) * W;
    px.x = raster[o];
    if (C > 1) px.y = raster[plane + o];
    if (C > 2) px.z = raster[2 * plane + o];
    if (C > 3) px.w = raster[3 * plane + o];
  }
  reinterpret_cast<uchar4*>(out)[static_cast<size_t>(t) * P * P + idx] = px;
}

// Resampled window read (dataset.py:97-115 when a modality's pixel size differs from the reference modality's: the tile's
// window is a FRACTIONAL number of the modality's pixels and rasterio resamples it to patch_sizes[mod] with
// Resampling.bilinear, boundless, fill 0).  Restates GDAL's RasterIO convolution (see oracle/resample.py, which this follows
// operation for operation, in double): destination pixel i sits at source coordinate off + (i + 0.5) * (win / ps); triangle
// kernel on pixel-centre distances, widened by win / ps when the read downsamples, weights normalised; fill pixels take part
// like data; integer rasters are rounded half up before the normalisation.
// win: double [n][4] = (row_off, col_off, height, width) in the modality's pixels.  One thread per output pixel.
constexpr int RS_MAXT = 20;   // taps per axis: ceil(2 * max(ratio, 1)) + 1 -> ratios up to 9.5
template <typename T>
__global__ void __launch_bounds__(256) gather_resampled_kernel(const T* __restrict__ raster, int C, int H, int W,
                                                               const double* __restrict__ win, int ps,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ stdv, float* __restrict__ out) {
  const int t = blockIdx.z, c = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= ps * ps) return;
  const int oy = p / ps, ox = p - oy * ps;
  const double row_off = win[4 * t], col_off = win[4 * t + 1], wh = win[4 * t + 2], ww = win[4 * t + 3];
  const double ry = wh / ps, rx = ww / ps;
  const double sy = ry > 1.0 ? ry : 1.0, sx = rx > 1.0 ? rx : 1.0;
  const int ty = static_cast<int>(ceil(2.0 * sy)) + 1, tx = static_cast<int>(ceil(2.0 * sx)) + 1;
  const double cy = row_off + (oy + 0.5) * ry, cx = col_off + (ox + 0.5) * rx;
  const int fy = static_cast<int>(floor(cy - 0.5 - sy)) + 1, fx = static_cast<int>(floor(cx - 0.5 - sx)) + 1;
  double wy[RS_MAXT], wx[RS_MAXT];
  double sumy = 0.0, sumx = 0.0;
  for (int k = 0; k < ty; ++k) {
    const double w = 1.0 - fabs((fy + k + 0.5) - cy) / sy;
    wy[k] = w > 0.0 ? w : 0.0;
    sumy += wy[k];
  }
  for (int k = 0; k < tx; ++k) {
    const double w = 1.0 - fabs((fx + k + 0.5) - cx) / sx;
    wx[k] = w > 0.0 ? w : 0.0;
    sumx += wx[k];
  }
  for (int k = 0; k < ty; ++k) wy[k] = sumy > 0.0 ? wy[k] / sumy : 0.0;
  for (int k = 0; k < tx; ++k) wx[k] = sumx > 0.0 ? wx[k] / sumx : 0.0;
  const T* plane = raster + static_cast<size_t>(c) * H * W;
  double acc = 0.0;
  for (int kx = 0; kx < tx; ++kx) {
    const int xi = fx + kx;
    double r = 0.0;                                   // rows first, like the oracle: r = sum_ky wy * v(yi, xi)
    if (xi >= 0 && xi < W) {
      for (int ky = 0; ky < ty; ++ky) {
        const int yi = fy + ky;
        const double v = (yi >= 0 && yi < H) ? static_cast<double>(plane[static_cast<size_t>(yi) * W + xi]) : 0.0;
        r += wy[ky] * v;
      }
    }
    acc += wx[kx] * r;
  }
  if (sizeof(T) == 1) {                               // uint8 raster: GDAL rounds the resampled value half up and clamps
    acc = floor(acc + 0.5);
    acc = acc < 0.0 ? 0.0 : (acc > 255.0 ? 255.0 : acc);
  }
  out[((static_cast<size_t>(t) * C + c) * ps + oy) * ps + ox] =
      static_cast<float>((acc - static_cast<double>(mean[c])) / static_cast<double>(stdv[c]));
}

}  // namespace fz

extern "C" int fz_gather_tiles_resampled(const void* raster, int src_f32, int C, int H, int W, const double* windows,
                                         int n_tiles, int ps, double max_ratio, const float* mean, const float* stdv, float* out,
                                         void* stream) {
  FZ_REQUIRE(C >= 1 && H > 0 && W > 0 && ps > 0, "fz_gather_tiles_resampled: bad shape C=%d H=%d W=%d ps=%d", C, H, W, ps);
  FZ_REQUIRE(max_ratio > 0.0 && static_cast<int>(ceil(2.0 * (max_ratio > 1.0 ? max_ratio : 1.0))) + 1 <= fz::RS_MAXT,
             "fz_gather_tiles_resampled: window / patch ratio %.3f needs more than %d taps per axis", max_ratio, fz::RS_MAXT);
  if (n_tiles <= 0) return 0;
  dim3 grid((ps * ps + 255) / 256, C, n_tiles);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (src_f32)
    fz::gather_resampled_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(raster), C, H, W, windows, ps, mean,
                                                             stdv, out);
  else
    fz::gather_resampled_kernel<uint8_t><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(raster), C, H, W, windows, ps,
                                                               mean, stdv, out);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int gather_f32_launch(const void* raster, int src_f32, int C, int H, int W, const int32_t* origins, int n_tiles, int P,
                             const float* mean, const float* stdv, float* out, void* stream) {
  FZ_REQUIRE(C >= 1 && H > 0 && W > 0 && P > 0 && P % 4 == 0, "fz_gather_tiles_f32: bad shape C=%d H=%d W=%d P=%d", C, H,
             W, P);
  if (n_tiles <= 0) return 0;
  const int threads = 256;
  dim3 grid((P * (P / 4) + threads - 1) / threads, C, n_tiles);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (src_f32)
    fz::gather_f32_kernel<float><<<grid, threads, 0, st>>>(static_cast<const float*>(raster), C, H, W, origins, P, mean, stdv,
                                                           out);
  else
    fz::gather_f32_kernel<uint8_t><<<grid, threads, 0, st>>>(static_cast<const uint8_t*>(raster), C, H, W, origins, P, mean,
                                                             stdv, out);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_gather_tiles_f32(const uint8_t* raster, int C, int H, int W, const int32_t* origins, int n_tiles,
                                   int P, const float* mean, const float* stdv, float* out, void* stream) {
  return gather_f32_launch(raster, 0, C, H, W, origins, n_tiles, P, mean, stdv, out, stream);
}

extern "C" int fz_gather_tiles_f32_from_f32(const float* raster, int C, int H, int W, const int32_t* origins, int n_tiles,
                                            int P, const float* mean, const float* stdv, float* out, void* stream) {
  return gather_f32_launch(raster, 1, C, H, W, origins, n_tiles, P, mean, stdv, out, stream);
}

extern "C" int fz_gather_tiles_u8(const uint8_t* raster, int C, int H, int W, const int32_t* origins, int n_tiles,
                                  int P, uint8_t* out, void* stream) {
  FZ_REQUIRE(C >= 1 && C <= 4 && H > 0 && W > 0 && P > 0, "fz_gather_tiles_u8: bad shape C=%d H=%d W=%d P=%d", C, H, W,
             P);
  if (n_tiles <= 0) return 0;
  const int threads = 256;
  dim3 grid((P * P + threads - 1) / threads, n_tiles);
  fz::gather_u8_kernel<<<grid, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(raster, C, H, W, origins, P, out);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Loss and optimizer kernels of the training step: FLAIRLosses' nn.CrossEntropyLoss(weight=w)
// (flair_hub/tasks/module_setup.py:119-200) on (B, C, H, W) logits with the targets of tasks_module.py:153-154, its
// gradient, the predictions of tasks_module.py:159, and torch.optim.AdamW's update (tasks_module.py:377-391).
// All HBM-bound: the loss reads C*4 B per pixel once, the gradient reads and writes C*4 B per pixel, AdamW moves
// 28 B per parameter (p, g, m, v read; p, m, v written).  The backward of the model itself is not built yet.
#include <cuda_bf16.h>

#include "common.h"
#include "reduce_vec.cuh"
#include "../../include/flair_zonal_b200.h"

namespace fz {

constexpr int CE_MAX_CLS = 32;
constexpr int CE_THREADS = 256;

// targets[b][y][x] = argmax_c onehot[b][c][y][x] (first maximum), tasks_module.py:154
__global__ void __launch_bounds__(CE_THREADS) onehot_argmax_kernel(const float* __restrict__ onehot, int32_t* __restrict__ tgt,
                                                                   int C, int64_t plane, int64_t n_px) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * CE_THREADS + threadIdx.x;
  if (i >= n_px) return;
  const int64_t b = i / plane, p = i - b * plane;
  const float* src = onehot + b * C * plane + p;
  float bv = src[0];
  int best = 0;
  for (int c = 1; c < C; ++c) {
    const float v = src[c * plane];
    if (v > bv) {
      bv = v;
      best = c;
    }
  }
  tgt[i] = best;
}

// cm[t][p] += 1 for every pixel with 0 <= t, p < C (pixels whose label or prediction is out of range are skipped, like
// sklearn's confusion_matrix(labels=range(C)) at prediction_writer.py:64 and torchmetrics' bincount of t * C + p).
// Integer counting: per-block histogram in shared memory, lanes holding the same (t, p) elect one to add their count
// (__match_any_sync: class rasters are spatially coherent, most warps hit one or two bins), one 64-bit global add per
// non-empty bin per block.  Sums of integers: the result is exact and order independent.
__global__ void __launch_bounds__(256) confusion_kernel(const int32_t* __restrict__ target, const int32_t* __restrict__ pred,
                                                        int64_t n, int C, unsigned long long* __restrict__ cm) {
  extern __shared__ unsigned int hist[];                       // [C * C]
  for (int k = threadIdx.x; k < C * C; k += 256) hist[k] = 0u;
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * 256;
  // whole warps iterate together (the loop bound is rounded up to a multiple of the stride) so the match is convergent
  for (int64_t i0 = static_cast<int64_t>(blockIdx.x) * 256; i0 < n; i0 += stride) {
    const int64_t i = i0 + threadIdx.x;
    int key = -1;
    if (i < n) {
      const int t = target[i], p = pred[i];
      if (t >= 0 && t < C && p >= 0 && p < C) key = t * C + p;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[key], static_cast<unsigned>(__popc(peers)));
  }
  __syncthreads();
  for (int k = threadIdx.x; k < C * C; k += 256)
    if (hist[k]) atomicAdd(&cm[k], static_cast<unsigned long long>(hist[k]));
}

// per pixel: lse, prediction, and the block's partial sums of w[t] * nll and w[t] (fixed-order tree: deterministic).
// grid = (blocks per sample, B): no 64-bit index division; MAXC = class count rounded up to 8 (19 -> 24 unrolled slots).
template <int MAXC>
__global__ void __launch_bounds__(CE_THREADS) ce_forward_kernel(const float* __restrict__ logits,
                                                                const int32_t* __restrict__ tgt,
                                                                const float* __restrict__ weight, float* __restrict__ lse,
                                                                int32_t* __restrict__ preds, double* __restrict__ partials,
                                                                int C, int plane) {
  __shared__ double s_l[CE_THREADS / 32], s_w[CE_THREADS / 32];
  const int p = blockIdx.x * CE_THREADS + threadIdx.x;
  const int64_t b = blockIdx.y;
  double wl = 0.0, ww = 0.0;
  if (p < plane) {
    const int64_t i = b * plane + p;
    const float* src = logits + b * C * plane + p;
    // all class planes are requested before the first use (the first version interleaved load and compare and paid one
    // DRAM round trip per class: 305 us; long_scoreboard 16 per issue in ncu)
    float v[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) v[c] = (c < C) ? __ldg(src + static_cast<int64_t>(c) * plane) : -INFINITY;
    float mx = v[0];
    int best = 0;
#pragma unroll
    for (int c = 1; c < MAXC; ++c)
      if (v[c] > mx) {          // padding classes hold -inf: never selected, exp() = 0
        mx = v[c];
        best = c;
      }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) sum += expf(v[c] - mx);
    const float l = logf(sum) + mx;
    lse[i] = l;
    if (preds) preds[i] = best;                    // argmax(softmax(logits)) = argmax(logits), tasks_module.py:159
    const int t = tgt[i];
    float xt = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c == t) xt = v[c];
    // a target outside [0, C) -- nn.CrossEntropyLoss's ignore_index (-100) or a bad label -- is IGNORED: weight 0, left
    // out of sum(w), zero gradient (PyTorch ignores -100 and raises on other values; it never reads out of bounds)
    const bool valid = t >= 0 && t < C;
    const float w = valid ? (weight ? weight[t] : 1.0f) : 0.0f;
    wl = valid ? static_cast<double>(w) * static_cast<double>(l - xt) : 0.0;
    ww = static_cast<double>(w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wl += __shfl_xor_sync(0xffffffffu, wl, o);
    ww += __shfl_xor_sync(0xffffffffu, ww, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_l[warp] = wl;
    s_w[warp] = ww;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b2 = 0.0;
    for (int k = 0; k < CE_THREADS / 32; ++k) {
      a += s_l[k];
      b2 += s_w[k];
    }
    const int64_t blk = b * gridDim.x + blockIdx.x;
    partials[2 * blk] = a;
    partials[2 * blk + 1] = b2;
  }
}

// out[0] = task_weight * sum(w * nll) / sum(w) (CrossEntropyLoss 'mean' with class weights), out[1] = sum(w)
__global__ void __launch_bounds__(256) ce_reduce_kernel(const double* __restrict__ partials, int n_blocks, float task_weight,
                                                        float* __restrict__ out) {
  __shared__ double s_l[256], s_w[256];
  double a = 0.0, b = 0.0;
  for (int k = threadIdx.x; k < n_blocks; k += 256) {
    a += partials[2 * k];
    b += partials[2 * k + 1];
  }
  s_l[threadIdx.x] = a;
  s_w[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_l[threadIdx.x] += s_l[threadIdx.x + o];
      s_w[threadIdx.x] += s_w[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = static_cast<float>(task_weight * s_l[0] / s_w[0]);
    out[1] = static_cast<float>(s_w[0]);
  }
}

// dlogits[b][c][p] = grad_scale * task_weight / sum(w) * w[t] * (softmax_c - [c == t])
__global__ void __launch_bounds__(CE_THREADS) ce_backward_kernel(const float* __restrict__ logits,
                                                                 const int32_t* __restrict__ tgt,
                                                                 const float* __restrict__ weight,
                                                                 const float* __restrict__ lse,
                                                                 const float* __restrict__ loss_out, float scale,
                                                                 float* __restrict__ dlogits, int C, int64_t plane,
                                                                 int64_t n_px) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * CE_THREADS + threadIdx.x;
  if (i >= n_px) return;
  const int64_t b = i / plane, p = i - b * plane;
  const int t = tgt[i];
  const bool valid = t >= 0 && t < C;                 // ignored targets (see ce_forward_kernel) get a zero gradient
  const float g = valid ? scale / loss_out[1] * (weight ? weight[t] : 1.0f) : 0.0f;
  const float l = lse[i];
  const float* src = logits + b * C * plane + p;
  float* dst = dlogits + b * C * plane + p;
  for (int c0 = 0; c0 < C; c0 += 8) {               // 8 planes in flight per thread
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (c0 + j < C) ? __ldg(src + (c0 + j) * plane) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j < C) dst[(c0 + j) * plane] = g * (expf(x[j] - l) - (c0 + j == t ? 1.0f : 0.0f));
  }
}

// torch.optim.AdamW, single-tensor formulation (decoupled weight decay, bias-corrected step)
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, int64_t n, float decay, float one_minus_b1,
                                                    float b2, float one_minus_b2, float step_size, float bc2_sqrt,
                                                    float eps) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  float pi = p[i] * decay;                              // param.mul_(1 - lr * weight_decay)
  float mi = m[i];
  mi = mi + one_minus_b1 * (gi - mi);                   // exp_avg.lerp_(grad, 1 - beta1)
  float vi = v[i] * b2;
  vi = vi + one_minus_b2 * (gi * gi);                   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pi = pi - step_size * (mi / denom);                   // param.addcdiv_(exp_avg, denom, value=-step_size)
  p[i] = pi;
  m[i] = mi;
  v[i] = vi;
}

// The step-dependent scalars computed ON THE DEVICE from a device-resident step counter, so that a whole training step
// (this kernel included) can sit in a CUDA graph and be replayed: state[0] += 1; hyper = {lr / (1 - b1^t), sqrt(1 - b2^t)}
__global__ void adamw_advance_kernel(long long* __restrict__ step, float* __restrict__ hyper, double lr, double beta1,
                                     double beta2) {
  const long long t = step[0] + 1;
  step[0] = t;
  hyper[0] = static_cast<float>(lr / (1.0 - pow(beta1, static_cast<double>(t))));
  hyper[1] = static_cast<float>(sqrt(1.0 - pow(beta2, static_cast<double>(t))));
}
__global__ void __launch_bounds__(256) adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, float decay, float one_minus_b1,
                                                        float b2, float one_minus_b2, const float* __restrict__ hyper,
                                                        float eps) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const float step_size = hyper[0], bc2_sqrt = hyper[1];
  const float gi = g[i];
  float pi = p[i] * decay;
  float mi = m[i];
  mi = mi + one_minus_b1 * (gi - mi);
  float vi = v[i] * b2;
  vi = vi + one_minus_b2 * (gi * gi);
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pi = pi - step_size * (mi / denom);
  p[i] = pi;
  m[i] = mi;
  v[i] = vi;
}

// out[c][r] = in[r][c] (bf16), 32 x 32 tiles through padded shared memory: the operand transposes that let the K-major
// tcgen05 GEMM compute dX = dY W and dW = dY^T X (first backward building block; an MN-major operand path would save them)
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in,
                                                             __nv_bfloat16* __restrict__ out, int R, int C) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int c0 = blockIdx.y * 32, r0 = blockIdx.x * 32;     // rows on grid.x: activations have millions of rows
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = r0 + ty + 8 * j, c = c0 + tx;
    if (r < R && c < C) tile[ty + 8 * j][tx] = in[static_cast<size_t>(r) * C + c];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + ty + 8 * j, r = r0 + tx;
    if (r < R && c < C) out[static_cast<size_t>(c) * R + r] = tile[tx][ty + 8 * j];
  }
}

// partial[s][n] = sum over the rows of chunk s of in[m][n]; then out[n] = sum_s partial[s][n] in a fixed order
__global__ void __launch_bounds__(256) colsum_partial_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ partial,
                                                             int64_t M, int N, int rows_per_chunk) {
  __shared__ float red[8][32];
  const int n = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * rows_per_chunk;
  const int64_t m1 = m0 + rows_per_chunk < M ? m0 + rows_per_chunk : M;
  float a = 0.f;
  if (n < N)
    for (int64_t m = m0 + ty; m < m1; m += 8) a += __bfloat162float(in[m * N + n]);
  red[ty][threadIdx.x & 31] = a;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    partial[static_cast<size_t>(blockIdx.y) * N + n] = t;
  }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int N,
                                                           int chunks) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  float t = 0.f;
  for (int s = 0; s < chunks; ++s) t += partial[static_cast<size_t>(s) * N + n];
  out[n] = t;
}

}  // namespace fz

extern "C" int fz_onehot_argmax(const float* onehot, int32_t* targets, int B, int C, int H, int W, void* stream) {
  using namespace fz;
  FZ_REQUIRE(B >= 0 && C >= 1 && H > 0 && W > 0, "fz_onehot_argmax: bad shape");
  const int64_t plane = static_cast<int64_t>(H) * W, n_px = plane * B;
  if (n_px == 0) return 0;
  onehot_argmax_kernel<<<static_cast<unsigned>((n_px + CE_THREADS - 1) / CE_THREADS), CE_THREADS, 0,
                         reinterpret_cast<cudaStream_t>(stream)>>>(onehot, targets, C, plane, n_px);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_confusion_matrix(const int32_t* target, const int32_t* pred, int64_t n, int C, int64_t* cm, void* stream) {
  using namespace fz;
  FZ_REQUIRE(n >= 0 && C >= 1 && C <= 96 && cm && (n == 0 || (target && pred)), "fz_confusion_matrix: n=%lld C=%d (1..96 classes)",
             (long long)n, C);
  if (n == 0) return 0;
  int64_t blocks = (n + 256 * 16 - 1) / (256 * 16);            // >= 16 pixels per thread before a block's global adds
  blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
  confusion_kernel<<<static_cast<unsigned>(blocks), 256, C * C * sizeof(unsigned int), reinterpret_cast<cudaStream_t>(stream)>>>(
      target, pred, n, C, reinterpret_cast<unsigned long long*>(cm));
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t fz_ce_workspace_doubles(int B, int H, int W) {
  const int64_t plane = static_cast<int64_t>(H) * W;
  return 2 * B * ((plane + fz::CE_THREADS - 1) / fz::CE_THREADS);
}

extern "C" int fz_ce_loss_forward(const float* logits, const int32_t* targets, const float* class_weight, float task_weight,
                                  float* lse, int32_t* preds, double* workspace, float* loss_out, int B, int C, int H,
                                  int W, void* stream) {
  using namespace fz;
  FZ_REQUIRE(B >= 1 && B <= 65535 && C >= 1 && C <= CE_MAX_CLS && H > 0 && W > 0,
             "fz_ce_loss_forward: B=%d C=%d (1..65535 samples, 1..%d classes)", B, C, CE_MAX_CLS);
  FZ_REQUIRE(logits && targets && lse && workspace && loss_out, "fz_ce_loss_forward: null pointer");
  const int64_t plane = static_cast<int64_t>(H) * W;
  FZ_REQUIRE(plane < (1LL << 31), "fz_ce_loss_forward: H*W must fit int32");
  const int per_sample = static_cast<int>((plane + CE_THREADS - 1) / CE_THREADS);
  const int64_t blocks = static_cast<int64_t>(per_sample) * B;
  FZ_REQUIRE(blocks < (1LL << 31), "fz_ce_loss_forward: too many pixels");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid(per_sample, B);
#define FZ_CE(maxc)                                                                                                  \
  ce_forward_kernel<maxc><<<grid, CE_THREADS, 0, st>>>(logits, targets, class_weight, lse, preds, workspace, C, \
                                                       static_cast<int>(plane))
  if (C <= 8) FZ_CE(8);
  else if (C <= 16) FZ_CE(16);
  else if (C <= 24) FZ_CE(24);
  else FZ_CE(32);
#undef FZ_CE
  ce_reduce_kernel<<<1, 256, 0, st>>>(workspace, static_cast<int>(blocks), task_weight, loss_out);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_ce_loss_backward(const float* logits, const int32_t* targets, const float* class_weight, float task_weight,
                                   const float* lse, const float* loss_out, float grad_scale, float* dlogits, int B, int C,
                                   int H, int W, void* stream) {
  using namespace fz;
  FZ_REQUIRE(B >= 1 && C >= 1 && H > 0 && W > 0 && logits && targets && lse && loss_out && dlogits,
             "fz_ce_loss_backward: bad arguments");
  const int64_t plane = static_cast<int64_t>(H) * W, n_px = plane * B;
  ce_backward_kernel<<<static_cast<unsigned>((n_px + CE_THREADS - 1) / CE_THREADS), CE_THREADS, 0,
                       reinterpret_cast<cudaStream_t>(stream)>>>(logits, targets, class_weight, lse, loss_out,
                                                                 grad_scale * task_weight, dlogits, C, plane, n_px);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                             double beta1, double beta2, double eps, double weight_decay, int step, void* stream) {
  using namespace fz;
  FZ_REQUIRE(n >= 0 && step >= 1 && param && grad && exp_avg && exp_avg_sq, "fz_adamw_step: bad arguments (step counts from 1)");
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
  adamw_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, n, static_cast<float>(1.0 - lr * weight_decay), static_cast<float>(1.0 - beta1),
      static_cast<float>(beta2), static_cast<float>(1.0 - beta2), static_cast<float>(lr / bc1),
      static_cast<float>(sqrt(bc2)), static_cast<float>(eps));
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                                 double beta1, double beta2, double eps, double weight_decay, int64_t* step_dev,
                                 float* hyper_dev, void* stream) {
  using namespace fz;
  FZ_REQUIRE(n >= 0 && param && grad && exp_avg && exp_avg_sq && step_dev && hyper_dev, "fz_adamw_step_dev: bad arguments");
  if (n == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  adamw_advance_kernel<<<1, 1, 0, st>>>(reinterpret_cast<long long*>(step_dev), hyper_dev, lr, beta1, beta2);
  adamw_dev_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      param, grad, exp_avg, exp_avg_sq, n, static_cast<float>(1.0 - lr * weight_decay), static_cast<float>(1.0 - beta1),
      static_cast<float>(beta2), static_cast<float>(1.0 - beta2), hyper_dev, static_cast<float>(eps));
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_transpose_bf16(const void* in, void* out, int R, int C, void* stream) {
  using namespace fz;
  FZ_REQUIRE(R > 0 && C > 0 && in && out, "fz_transpose_bf16: bad arguments");
  const dim3 grid((R + 31) / 32, (C + 31) / 32);
  FZ_REQUIRE(grid.y <= 65535, "fz_transpose_bf16: C=%d exceeds %d columns", C, 65535 * 32);
  transpose_bf16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), R, C);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_colsum_bf16(const void* in, float* partial, float* out, int64_t M, int N, int chunks, void* stream) {
  using namespace fz;
  FZ_REQUIRE(M > 0 && N > 0 && chunks >= 1 && chunks <= 65535 && in && partial && out, "fz_colsum_bf16: bad arguments");
  const int rows_per_chunk = static_cast<int>((M + chunks - 1) / chunks);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (N % 8 == 0 && M < (1LL << 31)) {         // 16-byte loads, row-walking blocks; partial sums in the library's own scratch
    FZ_CHECK_CUDA(fz::rv_colreduce<2>(in, nullptr, nullptr, out, nullptr, 1, static_cast<int>(M), N, st));
    return 0;
  }
  colsum_partial_kernel<<<dim3((N + 31) / 32, chunks), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(in), partial, M, N,
                                                                      rows_per_chunk);
  colsum_final_kernel<<<(N + 255) / 256, 256, 0, st>>>(partial, out, N, chunks);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

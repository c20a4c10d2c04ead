// Swin transformer pieces that are not plain GEMMs (timm SwinTransformerBlock / PatchMerging as run by smp's
// TimmUniversalEncoder for `swin_base_patch4_window12_384-upernet`, BASELINE.json configs[2]; reference call site
// flair_hub/models/flair_model.py:376):
//   fz_layernorm_rows   nn.LayerNorm over C of the fp32 residual stream -> bf16 GEMM operand
//   fz_merge_ln         PatchMerging gather (h0w0,h1w0,h0w1,h1w1) + LayerNorm(4C) -> bf16 GEMM operand
//   fz_swin_window_attn cyclic shift + zero-pad + window partition + softmax(QK^T*scale + rel-pos bias + shift mask) V
//                       + window reverse + crop + un-shift, all as index arithmetic around one CTA per (window, head)
//   fz_cast_f32_bf16    stage outputs -> decoder operands
// The qkv / proj / fc1 / fc2 / reduction linears run on tcgen05 (gemm_tcgen05*.cu).
//
// The attention contraction is 6 % of the model's FLOPs in 144x32x144 pieces; it runs on mma.sync (bf16, fp32
// accumulate) fragments held in registers, FlashAttention-2 style (S never leaves the register file).
#include "common.h"
#include "ptx.cuh"
#include "../../include/flair_zonal_b200.h"

#include "operand.cuh"

namespace fz {

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per output row of NV*128 floats; the row is the concatenation of NSEG equal segments read from
// different source rows (NSEG = 1: plain LayerNorm; NSEG = 4: patch merging).
template <int NV, int NSEG>
__global__ void __launch_bounds__(256) ln_gather_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bvec, op_t* __restrict__ out,
                                                        long long rows, int H, int W, int C, float eps) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  constexpr int CT = NV * 128;               // normalised width (= NSEG * C)
  const float* seg[NSEG];
  if (NSEG == 1) {
    seg[0] = x + row * C;
  } else {
    const int OW = W / 2, OH = H / 2;
    const int ox = static_cast<int>(row % OW), oy = static_cast<int>((row / OW) % OH);
    const long long b = row / (static_cast<long long>(OW) * OH);
    const float* base = x + ((b * H + 2 * oy) * W + 2 * ox) * C;
#pragma unroll
    for (int s = 0; s < NSEG; ++s) seg[s] = base + (static_cast<long long>(s & 1) * W + (s >> 1)) * C;  // dy = s&1, dx = s>>1
  }
  const int v_per_seg = C / 4;               // float4 per segment
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = i * 32 + lane;           // float4 index in the row
    const int s = NSEG == 1 ? 0 : idx / v_per_seg;
    const int o = NSEG == 1 ? idx : idx % v_per_seg;
    const float* sp = seg[0];
#pragma unroll
    for (int k = 1; k < NSEG; ++k) sp = (s == k) ? seg[k] : sp;
    v[i] = __ldg(reinterpret_cast<const float4*>(sp) + o);
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.0f / CT);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b2 = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    var += (a * a + b2 * b2) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var * (1.0f / CT) + eps);
  op_t* orow = out + row * CT;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = i * 32 + lane;
    const float4 g = __ldg(reinterpret_cast<const float4*>(w) + idx);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bvec) + idx);
    uint2 pk;
    pk.x = pack_op((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
    pk.y = pack_op((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
    reinterpret_cast<uint2*>(orow)[idx] = pk;
  }
}

template <int NSEG>
static int launch_ln(const float* x, const float* w, const float* b, void* out, long long rows, int H, int W, int C,
                     float eps, cudaStream_t st) {
  const int ct = NSEG * C;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  op_t* o = reinterpret_cast<op_t*>(out);
  switch (ct / 128) {
    case 1: ln_gather_kernel<1, NSEG><<<grid, 256, 0, st>>>(x, w, b, o, rows, H, W, C, eps); break;
    case 2: ln_gather_kernel<2, NSEG><<<grid, 256, 0, st>>>(x, w, b, o, rows, H, W, C, eps); break;
    case 4: ln_gather_kernel<4, NSEG><<<grid, 256, 0, st>>>(x, w, b, o, rows, H, W, C, eps); break;
    case 8: ln_gather_kernel<8, NSEG><<<grid, 256, 0, st>>>(x, w, b, o, rows, H, W, C, eps); break;
    case 16: ln_gather_kernel<16, NSEG><<<grid, 256, 0, st>>>(x, w, b, o, rows, H, W, C, eps); break;
    default: set_error("layernorm: width %d unsupported (128, 256, 512, 1024 or 2048)", ct); return -1;
  }
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <bool F16>
__global__ void __launch_bounds__(256) cast_f32_16_kernel(const float4* __restrict__ in, uint2* __restrict__ out,
                                                          size_t n4) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldg(in + i);
  out[i] = make_uint2(pack16<F16>(v.x, v.y), pack16<F16>(v.z, v.w));
}

// ------------------------------------------------------------------------------------------------ window attention
constexpr int WA_N = 144;        // token slots per window (window side <= 12)
constexpr int WA_LD = 40;        // bf16 per smem row: 80 B stride keeps ldmatrix conflict-free
constexpr int WA_THREADS = 288;  // 9 warps x 16 query rows
constexpr int WA_D = 32;         // head dim (all timm Swin variants)

struct WinAttnParams {
  const op_t* qkv;       // [B][H][W][3C]: q | k | v, channel = head*32 + d
  const op_t* qkv_bias;  // [3C] bf16: q/k/v of a zero (padded) token
  const float* table;             // [heads][(2ws-1)^2] relative position bias
  op_t* out;             // [B][H][W][C]
  int H, W, C, heads, ws, shift, nwy, nwx, hgroup;
  float scale;
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
// operands in the inference operand format (operand.cuh: fp16 unless built with FZ_OPERANDS_BF16), fp32 accumulate
__device__ __forceinline__ void mma_op_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef FZ_OPERANDS_BF16
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#else
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#endif
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int WA_MAT = WA_N * WA_LD;                      // bf16 elements of one staged q, k or v matrix
constexpr int WA_TAB = 23 * 23 + 3;                       // floats per staged bias table (padded to 16 B)
constexpr int WA_SMEM = 2 * 3 * WA_MAT * 2 + 2 * WA_TAB * 4 + 2 * WA_N * 4;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One CTA = one window x `hgroup` consecutive heads.  The q/k/v rows of head i+1 stream into the second smem buffer
// (cp.async) while head i is computed, so only the first head's load latency is exposed.
__global__ void __launch_bounds__(WA_THREADS, 2) swin_window_attn_kernel(WinAttnParams p) {
  extern __shared__ __align__(16) uint8_t wa_smem[];
  op_t* sBuf = reinterpret_cast<op_t*>(wa_smem);                 // [2][3][WA_MAT]
  float* sTabs = reinterpret_cast<float*>(wa_smem + 2 * 3 * WA_MAT * 2);          // [2][WA_TAB]
  int* sSrc = reinterpret_cast<int*>(sTabs + 2 * WA_TAB);   // source token (y*W+x) | -1 padded token | -2 unused slot
  int* sInfo = sSrc + WA_N;                                 // rel-pos code (ty*(2ws-1)+tx) | region << 16 | unused << 24

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nw = p.nwy * p.nwx;
  const int ngrp = p.heads / p.hgroup;
  const int head0 = (blockIdx.x % ngrp) * p.hgroup;
  const int win = (blockIdx.x / ngrp) % nw;
  const int b = blockIdx.x / (ngrp * nw);
  const int ws = p.ws, n_tok = ws * ws, span = 2 * ws - 1;
  const int wy = win / p.nwx, wx = win % p.nwx;

  if (tid < WA_N) {
    int src = -2, info = 1 << 24;
    if (tid < n_tok) {
      const int ty = tid / ws, tx = tid % ws;
      const int yp = wy * ws + ty, xp = wx * ws + tx;
      int reg = 0;
      if (p.shift > 0) {   // regions of the shift mask, on the PADDED grid
        const int Hp = p.nwy * ws, Wp = p.nwx * ws;
        const int rh = yp < Hp - ws ? 0 : (yp < Hp - p.shift ? 1 : 2);
        const int rw = xp < Wp - ws ? 0 : (xp < Wp - p.shift ? 1 : 2);
        reg = rh * 3 + rw;
      }
      info = (ty * span + tx) | (reg << 16);
      if (yp < p.H && xp < p.W) {
        int ys = yp + p.shift, xs = xp + p.shift;   // roll(-shift): shifted[yp] = x[(yp+shift) % H]
        if (ys >= p.H) ys -= p.H;
        if (xs >= p.W) xs -= p.W;
        src = ys * p.W + xs;
      } else {
        src = -1;
      }
    }
    sSrc[tid] = src;
    sInfo[tid] = info;
  }
  __syncthreads();

  // 144 tokens x (q,k,v) x 4 chunks of 16 B = 6 per thread.  The chunk -> (token, matrix) mapping and the global
  // offsets are the same for every head; unused slots (window side < 12) are zeroed once in both buffers.
  static_assert(WA_N * 12 == 6 * WA_THREADS, "load loop is unrolled for 6 chunks per thread");
  const size_t tok0 = static_cast<size_t>(b) * p.H * p.W;
  const int C3 = 3 * p.C;
  const op_t* gsrc[6];
  int soff[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int i = tid + k * WA_THREADS;
    const int tok = i / 12, rem = i % 12, m = rem >> 2, ch = rem & 3;
    const int src = sSrc[tok];
    const int coff = m * p.C + ch * 8;
    soff[k] = m * WA_MAT + tok * WA_LD + ch * 8;
    gsrc[k] = src >= 0 ? p.qkv + (tok0 + src) * C3 + coff : (src == -1 ? p.qkv_bias + coff : nullptr);
    if (src == -2) {
      *reinterpret_cast<uint4*>(sBuf + soff[k]) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sBuf + 3 * WA_MAT + soff[k]) = make_uint4(0, 0, 0, 0);
    }
  }
#define FZ_WA_ISSUE_HEAD(head_, buf_)                                                                        \
  do {                                                                                                         \
    _Pragma("unroll") for (int k = 0; k < 6; ++k)                                                              \
      if (gsrc[k] != nullptr) cp_async16(sBuf + (buf_) * 3 * WA_MAT + soff[k], gsrc[k] + (head_) * WA_D);      \
    cp_async_commit();                                                                                         \
  } while (0)
  const int tab_n = span * span;
  FZ_WA_ISSUE_HEAD(head0, 0);
  for (int i = tid; i < tab_n; i += WA_THREADS)
    sTabs[i] = p.table[static_cast<size_t>(head0) * tab_n + i] * 1.4426950408889634f;   // log2 domain

  const int r0 = warp * 16;
  const int lr = lane >> 2, lc = (lane & 3) * 2;
  const bool general = (p.shift > 0 && (wy == p.nwy - 1 || wx == p.nwx - 1)) || n_tok < WA_N;

  for (int hh = 0; hh < p.hgroup; ++hh) {
    const int head = head0 + hh;
    const int cur = hh & 1;
    op_t* sQ = sBuf + cur * 3 * WA_MAT;
    op_t* sK = sQ + WA_MAT;
    op_t* sV = sK + WA_MAT;
    const float* sTab = sTabs + cur * WA_TAB;
    cp_async_wait_all();
    __syncthreads();          // head's q/k/v + table visible; everyone is done with the other buffer
    float tnext[2] = {0.f, 0.f};
    const bool more = hh + 1 < p.hgroup;
    if (more) {
      FZ_WA_ISSUE_HEAD(head + 1, cur ^ 1);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (tid + k * WA_THREADS < tab_n) tnext[k] = __ldg(p.table + static_cast<size_t>(head + 1) * tab_n + tid + k * WA_THREADS);
    }

  uint32_t qa[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    ldsm_x4(qa[ks], &sQ[(r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * WA_LD + ks * 16 + (lane >> 4) * 8]);

  float s[18][4];
#pragma unroll
  for (int nt = 0; nt < 18; ++nt) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    uint32_t kb[4];
    ldsm_x4(kb, &sK[(nt * 8 + (lane & 7)) * WA_LD + (lane >> 3) * 8]);
    mma_op_16816(s[nt], qa[0], kb[0], kb[1]);
    mma_op_16816(s[nt], qa[1], kb[2], kb[3]);
  }

  // scale + relative position bias + shift mask, then softmax over the 144 key slots (unused slots excluded).
  // Everything is kept in the log2 domain (scale, table and mask pre-multiplied by log2 e) so that the
  // exponential is one ex2.approx.  The mask / unused-slot handling is skipped (CTA-uniform branch) for the
  // windows that cannot need it: un-shifted blocks, and interior windows of shifted blocks.
  constexpr float LOG2E = 1.4426950408889634f;
  const float sl2 = p.scale * LOG2E;
  const int base = (ws - 1) * span + (ws - 1);
  const int info_lo = sInfo[r0 + lr], info_hi = sInfo[r0 + lr + 8];
  const int code_lo = (info_lo & 0xffff) + base, code_hi = (info_hi & 0xffff) + base;
  float m_lo = -INFINITY, m_hi = -INFINITY;
  if (!general) {
#pragma unroll
    for (int nt = 0; nt < 18; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = sInfo[nt * 8 + lc + e] & 0xffff;
        const float a = fmaf(s[nt][e], sl2, sTab[code_lo - cc]);
        const float c = fmaf(s[nt][2 + e], sl2, sTab[code_hi - cc]);
        s[nt][e] = a;
        s[nt][2 + e] = c;
        m_lo = fmaxf(m_lo, a);
        m_hi = fmaxf(m_hi, c);
      }
    }
  } else {
    const int reg_lo = (info_lo >> 16) & 0xff, reg_hi = (info_hi >> 16) & 0xff;
    constexpr float MASKED = -100.0f * LOG2E;
#pragma unroll
    for (int nt = 0; nt < 18; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ci = sInfo[nt * 8 + lc + e];
        const int cc = ci & 0xffff, cr = (ci >> 16) & 0xff;
        const bool unused = (ci >> 24) != 0;
        float a = fmaf(s[nt][e], sl2, sTab[code_lo - cc]) + (cr != reg_lo ? MASKED : 0.0f);
        float c = fmaf(s[nt][2 + e], sl2, sTab[code_hi - cc]) + (cr != reg_hi ? MASKED : 0.0f);
        a = unused ? -INFINITY : a;
        c = unused ? -INFINITY : c;
        s[nt][e] = a;
        s[nt][2 + e] = c;
        m_lo = fmaxf(m_lo, a);
        m_hi = fmaxf(m_hi, c);
      }
    }
  }
  m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
  m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
  m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
  m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
  float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
  for (int nt = 0; nt < 18; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float a = ex2_approx(s[nt][e] - m_lo);
      const float c = ex2_approx(s[nt][2 + e] - m_hi);
      s[nt][e] = a;
      s[nt][2 + e] = c;
      sum_lo += a;
      sum_hi += c;
    }
  }
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);

  float o[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 9; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_op(s[2 * kk][0], s[2 * kk][1]);
    pa[1] = pack_op(s[2 * kk][2], s[2 * kk][3]);
    pa[2] = pack_op(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[3] = pack_op(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t vb[4];
      ldsm_x4_t(vb, &sV[(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * WA_LD + (np * 2 + (lane >> 4)) * 8]);
      mma_op_16816(o[np * 2], pa, vb[0], vb[1]);
      mma_op_16816(o[np * 2 + 1], pa, vb[2], vb[3]);
    }
  }
  const float inv_lo = 1.0f / sum_lo, inv_hi = 1.0f / sum_hi;
  // stage this warp's 16 output rows in its own (already consumed) Q rows, then 16 B stores per token
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    *reinterpret_cast<uint32_t*>(&sQ[(r0 + lr) * WA_LD + nt * 8 + lc]) = pack_op(o[nt][0] * inv_lo, o[nt][1] * inv_lo);
    *reinterpret_cast<uint32_t*>(&sQ[(r0 + lr + 8) * WA_LD + nt * 8 + lc]) = pack_op(o[nt][2] * inv_hi, o[nt][3] * inv_hi);
  }
  __syncwarp();
#pragma unroll
  for (int i = lane; i < 64; i += 32) {
    const int tok = r0 + (i >> 2), ch = i & 3;
    const int src = sSrc[tok];
    if (src >= 0)
      *reinterpret_cast<uint4*>(p.out + (tok0 + src) * p.C + head * WA_D + ch * 8) =
          *reinterpret_cast<const uint4*>(&sQ[tok * WA_LD + ch * 8]);
  }
    if (more) {   // next head's table (log2 domain); its previous user finished before this iteration's barrier
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (tid + k * WA_THREADS < tab_n) sTabs[(cur ^ 1) * WA_TAB + tid + k * WA_THREADS] = tnext[k] * 1.4426950408889634f;
    }
  }   // heads
#undef FZ_WA_ISSUE_HEAD
}

}  // namespace fz

extern "C" int fz_layernorm_rows(const float* x, const float* w, const float* b, void* out_bf16, int64_t rows, int C,
                                 float eps, void* stream) {
  using namespace fz;
  FZ_REQUIRE(rows >= 0 && C > 0 && C % 128 == 0, "fz_layernorm_rows: C=%d must be a multiple of 128", C);
  if (rows == 0) return 0;
  return launch_ln<1>(x, w, b, out_bf16, rows, 0, 0, C, eps, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int fz_merge_ln(const float* x, const float* w, const float* b, void* out_bf16, int B, int H, int W, int C,
                           float eps, void* stream) {
  using namespace fz;
  FZ_REQUIRE(B >= 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 32 == 0,
             "fz_merge_ln: bad shape H=%d W=%d C=%d (even H, W; C multiple of 32)", H, W, C);
  if (B == 0) return 0;
  return launch_ln<4>(x, w, b, out_bf16, static_cast<long long>(B) * (H / 2) * (W / 2), H, W, C, eps,
                      reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int fz_cast_f32_16(const float* in, void* out16, int out_dtype, int64_t n, void* stream) {
  using namespace fz;
  FZ_REQUIRE(n >= 0 && n % 4 == 0, "fz_cast_f32_16: n must be a multiple of 4");
  FZ_REQUIRE(out_dtype == FZ_BF16 || out_dtype == FZ_F16, "fz_cast_f32_16: the output is FZ_BF16 or FZ_F16");
  if (n == 0) return 0;
  const size_t n4 = static_cast<size_t>(n) / 4;
  const unsigned grid = static_cast<unsigned>((n4 + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (out_dtype == FZ_F16)
    cast_f32_16_kernel<true><<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<uint2*>(out16), n4);
  else
    cast_f32_16_kernel<false><<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<uint2*>(out16), n4);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_cast_f32_bf16(const float* in, void* out_bf16, int64_t n, void* stream) {
  return fz_cast_f32_16(in, out_bf16, FZ_BF16, n, stream);
}

extern "C" int fz_swin_window_attn(const void* qkv_bf16, const void* qkv_bias_bf16, const float* table, void* out_bf16,
                                   int B, int H, int W, int C, int heads, int window, int shift, float scale,
                                   void* stream) {
  using namespace fz;
  FZ_REQUIRE(B >= 0 && H > 0 && W > 0, "fz_swin_window_attn: bad shape");
  FZ_REQUIRE(window >= 1 && window <= 12, "fz_swin_window_attn: window %d unsupported (1..12)", window);
  FZ_REQUIRE(heads > 0 && C == heads * WA_D, "fz_swin_window_attn: C=%d must be heads*32 (heads=%d)", C, heads);
  FZ_REQUIRE(shift >= 0 && shift < window, "fz_swin_window_attn: shift %d out of range", shift);
  if (B == 0) return 0;
  WinAttnParams p;
  p.qkv = reinterpret_cast<const op_t*>(qkv_bf16);
  p.qkv_bias = reinterpret_cast<const op_t*>(qkv_bias_bf16);
  p.table = table;
  p.out = reinterpret_cast<op_t*>(out_bf16);
  p.H = H; p.W = W; p.C = C; p.heads = heads; p.ws = window; p.shift = shift;
  p.nwy = (H + window - 1) / window;
  p.nwx = (W + window - 1) / window;
  p.scale = scale;
  p.hgroup = heads % 4 == 0 ? 4 : (heads % 2 == 0 ? 2 : 1);
  const long long grid = static_cast<long long>(B) * p.nwy * p.nwx * (heads / p.hgroup);
  FZ_REQUIRE(grid < (1ll << 31), "fz_swin_window_attn: grid too large");
  FZ_ENSURE_SMEM(swin_window_attn_kernel, WA_SMEM);
  swin_window_attn_kernel<<<static_cast<unsigned>(grid), WA_THREADS, WA_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// The 16-bit OPERAND format of the inference kernels (what is fed to the tensor cores and stored between layers).
//
// Round 1 used bf16 (8 significand bits): logits were within 0.9 % of the fp32 reference but only 99.0-99.3 % of the
// pixels kept its class.  tcgen05.mma.kind::f16 takes fp16 operands at the same rate, fp16 has 11 significand bits, and
// every 16-bit tensor of the forward is bounded (LayerNorm / BatchNorm+ReLU outputs, GELU of a normalised input,
// GRN-scaled weights), so the inference path stores fp16: 8x less rounding noise at the same bandwidth and MMA rate
// (tests/error_budget.py has the per-family table).  Conversions saturate at +-65504 instead of producing inf.
// The training kernels (backward_ops.cu, training_ops.cu) keep bf16 -- gradients need its exponent range -- and the
// GEMM takes the format per call (FZ_EPI_OPERANDS_F16).
//
// Build with -DFZ_OPERANDS_BF16 for the round-1 behaviour (A/B measurements only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace fz {

#ifdef FZ_OPERANDS_BF16
typedef __nv_bfloat16 op_t;
typedef __nv_bfloat162 op2_t;
constexpr bool OP_F16 = false;
#define FZ_OP_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
__device__ __forceinline__ op_t f2op(float x) { return __float2bfloat16_rn(x); }
__device__ __forceinline__ op2_t ff2op2(float lo, float hi) { return __floats2bfloat162_rn(lo, hi); }
__device__ __forceinline__ float op2f(op_t x) { return __bfloat162float(x); }
__device__ __forceinline__ float2 op22ff(op2_t x) { return __bfloat1622float2(x); }
#else
typedef __half op_t;
typedef __half2 op2_t;
constexpr bool OP_F16 = true;
#define FZ_OP_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
__device__ __forceinline__ op2_t ff2op2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return *reinterpret_cast<op2_t*>(&r);
}
__device__ __forceinline__ op_t f2op(float x) {
  uint16_t r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return *reinterpret_cast<op_t*>(&r);
}
__device__ __forceinline__ float op2f(op_t x) { return __half2float(x); }
__device__ __forceinline__ float2 op22ff(op2_t x) { return __half22float2(x); }
#endif

__device__ __forceinline__ uint32_t pack_op(float lo, float hi) {
  op2_t v = ff2op2(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_op(uint32_t w) { return op22ff(*reinterpret_cast<const op2_t*>(&w)); }

}  // namespace fz

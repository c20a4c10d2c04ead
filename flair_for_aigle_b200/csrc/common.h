// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace fz {

void set_error(const char* fmt, ...);
const char* last_error();

// Encode a tiled tensor map over a 16-bit (bf16 / fp16: same element size, zero OOB fill) tensor (driver entry point resolved lazily through
// cudaGetDriverEntryPoint, so the library does not link libcuda directly).
// dims/strides are innermost-first; strides_bytes has rank-1 entries (dim 1..rank-1).
// swizzle_bytes in {0 (none), 32, 64, 128}.
int make_tmap16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, uint32_t swizzle_bytes,
                   const uint32_t* elem_strides = nullptr);

// The same over a 32-bit (float) tensor: the TMA STORE of fp32 GEMM outputs.
int make_tmap32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, uint32_t swizzle_bytes, const uint32_t* elem_strides = nullptr);

// One-time (per device, thread-safe) opt-in of a kernel to `bytes` of dynamic shared memory; 0 or -2 (error set).
int ensure_dynamic_smem(const void* kernel, int bytes);
// Multiprocessor count of the CURRENT device (cached per device); 0 on error (error set).
int device_sm_count();

#define FZ_ENSURE_SMEM(kernel, bytes)                                                        \
  do {                                                                                       \
    if (int _rc = fz::ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), (bytes))) return _rc; \
  } while (0)

#define FZ_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      fz::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                             \
    }                                                                                        \
  } while (0)

#define FZ_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      fz::set_error(__VA_ARGS__);  \
      return -1;                   \
    }                              \
  } while (0)

}  // namespace fz

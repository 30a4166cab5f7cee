// Host-side helpers shared by the C-ABI translation units: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/b21.h"

namespace b21 {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define B21_CHECK_ARG(cond, ...)   \
  do {                             \
    if (!(cond)) {                 \
      b21::set_error(__VA_ARGS__); \
      return B21_ERR_BAD_ARG;      \
    }                              \
  } while (0)

#define B21_CUDA(expr)                                                    \
  do {                                                                    \
    cudaError_t e__ = (expr);                                             \
    if (e__ != cudaSuccess) return b21::cuda_fail(e__, #expr);            \
  } while (0)

#define B21_LAUNCH_CHECK(name)                                            \
  do {                                                                    \
    cudaError_t e__ = cudaGetLastError();                                 \
    if (e__ != cudaSuccess) return b21::cuda_fail(e__, name);             \
  } while (0)

// Encode a bf16 tiled tensor map (rank <= 5). dims/box innermost-first, strides in BYTES for dims 1..rank-1.
// swizzle: 0 = none, 3 = 128B (CUtensorMapSwizzle values).  OOB elements read as zero.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle);

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace b21

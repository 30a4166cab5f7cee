// Shared pieces of the folded-EvoNorm conv epilogues (see fold.cu for the scheme).
#pragma once
#include "ptx.cuh"

namespace b21 {

struct FoldExtras {
  const float* table;   // bias table [N][ncls][Cout] fp32 (ncls = 27 border classes for k = 3, 1 for k = 1); replaces bias
  float* chan_sum;      // [N][Cout] fp32: per-channel sums of the STORED (bf16-rounded) outputs (SE squeeze), or NULL
  long long wstride;    // bytes between the per-sample packed weights (0 = one shared weight set)
  int act;              // 0 = identity, 1 = x * sigmoid(x) applied before the store (after bias and statistics)
};

__device__ __forceinline__ int border_class(int i, int n) { return i == 0 ? 0 : (i == n - 1 ? 2 : 1); }
// x * sigmoid(x) = x / (1 + 2^(-x log2 e)) on the MUFU pipe (ex2 + rcp, flush-to-zero: no range fix-up branches).
// x -> -inf: ex2 -> +inf, rcp -> 0, result -0;  x -> +inf: ex2 -> 0, result x.
__device__ __forceinline__ float swishf(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return x * r;
}

// Transposed warp reduction: every lane holds 32 values v[0..32); afterwards v[0] of lane l is the sum over all
// lanes of their v[l] (31 shuffles instead of 32 x 5).
__device__ __forceinline__ float warp_transpose_sum32(float* v, int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

}  // namespace b21

// Shared pieces of the folded-EvoNorm conv epilogues (see fold.cu for the scheme).
#pragma once
#include "ptx.cuh"

namespace b21 {

struct FoldExtras {
  const float* table;   // bias table [N][ncls][Cout] fp32 (ncls = 27 border classes for k = 3, 1 for k = 1); replaces bias
  float* chan_sum;      // [N][Cout] fp32: per-channel sums of the STORED (bf16-rounded) outputs (SE squeeze), or NULL
  long long wstride;    // bytes between the per-sample packed weights (0 = one shared weight set)
  int act;              // 0 = identity, 1 = x * sigmoid(x) applied before the store (after bias and statistics),
                        // 2 = the same through ONE MUFU op (tanh.approx): for epilogue-bound layers (Cin = 8)
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

// x * sigmoid(x) = h + h * tanh(h), h = x / 2: one MUFU op instead of two.  tanh.approx.f32 has a maximum relative
// error of 2^-11, i.e. |error| <= 2.5e-4 * |x|: below the bf16 rounding of the stored value for x > 0 and about one
// bf16 ulp on the negative tail.  Used where the MUFU pipe, not the tensor pipe, bounds the kernel (the 8 -> 48 input
// conv: 96 MUFU ops per voxel against 9 MMAs per 128 voxels).
__device__ __forceinline__ float swishf_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
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

// Bias-table rows computed by the extra blocks appended to the per-sample weight packing kernels (one launch per
// consumer conv): T[n][cls][co] = bias[co] + sum_ci ws[cls][ci][co] * B[n][ci].
struct BiasTableArgs {
  const float* ws;     // [ncls][cin][cout] border-class tap sums of the fp32 weight (b21_border_weight_sums)
  const float* bias;   // [cout] or NULL
  const float* b_in;   // B[n][ci], row stride ld
  float* table;        // [n][ncls][cout]
  int ld, cout, cin, ncls;
};
__device__ __forceinline__ void bias_table_block(const BiasTableArgs& t, int cls, int n) {
  // ws is cold in L2 by the time the next forward needs it: issue the loads of a row block up front (independent of
  // the accumulation chain) instead of one dependent load per multiply-add
  const float* bp = t.b_in + size_t(n) * t.ld;
  for (int co = threadIdx.x; co < t.cout; co += blockDim.x) {
    const float* wp = t.ws + size_t(cls) * t.cin * t.cout + co;
    float s = t.bias ? t.bias[co] : 0.f;
    int ci = 0;
    for (; ci + 16 <= t.cin; ci += 16) {
      float wv[16], bv[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        wv[k] = __ldg(wp + size_t(ci + k) * t.cout);
        bv[k] = __ldg(bp + ci + k);
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) s = fmaf(wv[k], bv[k], s);
    }
    for (; ci < t.cin; ++ci) s = fmaf(__ldg(wp + size_t(ci) * t.cout), __ldg(bp + ci), s);
    t.table[(size_t(n) * t.ncls + cls) * t.cout + co] = s;
  }
}

}  // namespace b21

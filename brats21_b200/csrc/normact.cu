// normact.cu — the rest of the reference's norm / activation factory for EquiUnet (networks/factory.py:179-200):
// get_norm_layer {"group", "instance", "batch", "none"} x get_act {"relu", "leakyrelu", "elu"}.
// (GroupNorm(8) + ReLU, the README recipe, keeps its dedicated kernel: elementwise.cu norm_apply.)
//
// HBM-bound, vectorised (8 bf16 channels = 16 B per thread and access), channels-last:
//   channel_stats   per (n, c): sum and sum of squares over the voxels (double), one read pass        2 B / element
//   norm_coeffs     tiny: statistics (+ gamma, beta, running stats) -> per-(n, c) affine (a, b), i.e. the normalisation
//                   folded into one multiply-add; updates BatchNorm's running statistics in training mode
//   affine_act      y = act(a[n][c] * x + b[n][c]), one read + one write pass                          4 B / element
#include "ptx.cuh"
#include "host_common.h"
#include "reduce.cuh"
#include <math.h>

namespace b21 {

namespace {

__device__ __forceinline__ void unpack8n(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8n(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

}  // namespace

// grid = (gx, N), gx * blockDim a multiple of C/8: a thread always serves the same 8 channels
__global__ void __launch_bounds__(256) channel_stats_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                            double* __restrict__ out, long long nvox, int C) {
  extern __shared__ float sm[];  // [2][C]
  const int n = blockIdx.y;
  const int chunks = C >> 3;
  const long long total = nvox * chunks, T = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ck = int(i % chunks);
  const long long vstep = T / chunks;
  long long v = i / chunks;
  const __nv_bfloat16* xn = x + size_t(n) * nvox * ldx + ck * 8;
  float sq[2][8];
  float (&s)[8] = sq[0], (&q)[8] = sq[1];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  // short fp32 runs (<= nvox / grid voxels per thread), combined in double below
  for (; i < total; i += T, v += vstep) {
    float f[8];
    unpack8n(__ldg(reinterpret_cast<const uint4*>(xn + v * ldx)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += f[j];
      q[j] = fmaf(f[j], f[j], q[j]);
    }
  }
  block_chunk_reduce<2>(sq, chunks, C, sm);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(out + (size_t(n) * C + c) * 2, double(sm[c]));
    atomicAdd(out + (size_t(n) * C + c) * 2 + 1, double(sm[C + c]));
  }
}

// kind: 0 group (conv-epilogue statistics double[SLOTS][N][8][2]), 1 instance, 2 batch (training: batch statistics,
// running statistics updated), 3 batch (eval: running statistics), 4 none (a = 1, b = 0)
__global__ void norm_coeffs_kernel(int kind, const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var,
                                   float momentum, float* __restrict__ a_out, float* __restrict__ b_out, int N, int C,
                                   long long nvox, float eps) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    if (kind == 4) {
      for (int n = 0; n < N; ++n) { a_out[size_t(n) * C + c] = 1.f; b_out[size_t(n) * C + c] = 0.f; }
    } else if (kind == 3) {
      const float r = rsqrtf(running_var[c] + eps);
      for (int n = 0; n < N; ++n) { a_out[size_t(n) * C + c] = g * r; b_out[size_t(n) * C + c] = bt - running_mean[c] * g * r; }
    } else if (kind == 2) {
      double s = 0.0, q = 0.0;
      for (int n = 0; n < N; ++n) { s += stats[(size_t(n) * C + c) * 2]; q += stats[(size_t(n) * C + c) * 2 + 1]; }
      const double cnt = double(N) * double(nvox), mean = s / cnt;
      double var = q / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      const float r = float(1.0 / sqrt(var + double(eps)));
      for (int n = 0; n < N; ++n) { a_out[size_t(n) * C + c] = g * r; b_out[size_t(n) * C + c] = bt - float(mean) * g * r; }
      if (running_mean && running_var) {  // nn.BatchNorm: running_var tracks the UNBIASED batch variance
        const double varu = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * float(mean);
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * float(varu);
      }
    } else {
      for (int n = 0; n < N; ++n) {
        double s = 0.0, q = 0.0, cnt;
        if (kind == 1) {
          s = stats[(size_t(n) * C + c) * 2];
          q = stats[(size_t(n) * C + c) * 2 + 1];
          cnt = double(nvox);
        } else {
          const int gsz = C / 8, grp = c / gsz;
          for (int slot = 0; slot < B21_STAT_SLOTS; ++slot) {
            const double* st = stats + ((size_t(slot) * N + n) * 8 + grp) * 2;
            s += st[0];
            q += st[1];
          }
          cnt = double(nvox) * gsz;
        }
        const double mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        const float r = float(1.0 / sqrt(var + double(eps)));
        a_out[size_t(n) * C + c] = g * r;
        b_out[size_t(n) * C + c] = bt - float(mean) * g * r;
      }
    }
  }
}

// act: 0 identity, 1 ReLU, 2 LeakyReLU(slope), 3 ELU(alpha = 1)  (MONAI Act["relu" | "leakyrelu" | "elu"])
template <int ACT>
__global__ void __launch_bounds__(256) affine_act_kernel(const __nv_bfloat16* x, int ldx, __nv_bfloat16* y, int ldy,
                                                         const float* __restrict__ a_in, const float* __restrict__ b_in,
                                                         float slope, long long nvox, int C) {
  const int n = blockIdx.y;
  const int chunks = C >> 3;
  const long long total = nvox * chunks, T = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ck = int(i % chunks);
  const long long vstep = T / chunks;
  long long v = i / chunks;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = a_in[size_t(n) * C + ck * 8 + j];
    b[j] = b_in[size_t(n) * C + ck * 8 + j];
  }
  const __nv_bfloat16* xn = x + size_t(n) * nvox * ldx + ck * 8;
  __nv_bfloat16* yn = y + size_t(n) * nvox * ldy + ck * 8;
  constexpr int U = 4;
  for (; i < total; i += T * U, v += vstep * U) {
    uint4 in[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * T < total) in[u] = *reinterpret_cast<const uint4*>(xn + (v + u * vstep) * ldx);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u * T < total) {
        float f[8];
        unpack8n(in[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = fmaf(f[j], a[j], b[j]);
          if (ACT == 1) f[j] = fmaxf(t, 0.f);
          else if (ACT == 2) f[j] = t > 0.f ? t : t * slope;
          else if (ACT == 3) f[j] = t > 0.f ? t : expm1f(t);
          else f[j] = t;
        }
        *reinterpret_cast<uint4*>(yn + (v + u * vstep) * ldy) = pack8n(f);
      }
    }
  }
}

static int stats_grid(long long nvox, int chunks, int n, int per_thread) {
  long long blocks = (nvox * chunks + 256LL * per_thread - 1) / (256LL * per_thread);
  const long long cap = (long long)num_sms() * 8 / n > 0 ? (long long)num_sms() * 8 / n : 1;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return int((blocks + chunks - 1) / chunks * chunks);  // gx * 256 must be a multiple of C / 8
}

}  // namespace b21

using namespace b21;
typedef __nv_bfloat16 bf16;

extern "C" int b21_channel_stats(const void* x, int ldx, double* out, int n, long long nvox, int c, void* stream) {
  B21_CHECK_ARG(x && out && n > 0 && nvox > 0, "channel_stats: bad arguments");
  B21_CHECK_ARG(c % 8 == 0 && c >= 8 && ldx % 8 == 0 && ldx >= c, "channel_stats: C/ld must be multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  B21_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 2 * size_t(n) * c, st));
  dim3 grid(stats_grid(nvox, c / 8, n, 64), n);
  channel_stats_kernel<<<grid, 256, sizeof(float) * 2 * c, st>>>((const bf16*)x, ldx, out, nvox, c);
  B21_LAUNCH_CHECK("channel_stats_kernel");
  return B21_OK;
}

extern "C" int b21_norm_coeffs(int kind, const double* stats, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, float momentum, float* a_out, float* b_out, int n, int c,
                               long long nvox, float eps, void* stream) {
  B21_CHECK_ARG(kind >= 0 && kind <= 4 && a_out && b_out && n > 0 && c > 0, "norm_coeffs: bad arguments");
  B21_CHECK_ARG(kind >= 3 || stats, "norm_coeffs: statistics missing");
  B21_CHECK_ARG(kind != 3 || (running_mean && running_var), "norm_coeffs: eval-mode batch norm needs running statistics");
  B21_CHECK_ARG(kind != 0 || c % 8 == 0, "norm_coeffs: GroupNorm(8) needs C divisible by 8");
  norm_coeffs_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(kind, stats, gamma, beta, running_mean, running_var,
                                                                        momentum, a_out, b_out, n, c, nvox, eps);
  B21_LAUNCH_CHECK("norm_coeffs_kernel");
  return B21_OK;
}

extern "C" int b21_affine_act(const void* x, int ldx, void* y, int ldy, const float* a, const float* b, int act,
                              float slope, int n, long long nvox, int c, void* stream) {
  B21_CHECK_ARG(x && y && a && b && n > 0 && nvox > 0, "affine_act: bad arguments");
  B21_CHECK_ARG(c % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= c && ldy >= c, "affine_act: bad C/ld");
  B21_CHECK_ARG(act >= 0 && act <= 3, "affine_act: act must be 0 (none), 1 (relu), 2 (leakyrelu), 3 (elu)");
  dim3 grid(stats_grid(nvox, c / 8, n, 4), n);
  cudaStream_t st = (cudaStream_t)stream;
  switch (act) {
    case 0: affine_act_kernel<0><<<grid, 256, 0, st>>>((const bf16*)x, ldx, (bf16*)y, ldy, a, b, slope, nvox, c); break;
    case 1: affine_act_kernel<1><<<grid, 256, 0, st>>>((const bf16*)x, ldx, (bf16*)y, ldy, a, b, slope, nvox, c); break;
    case 2: affine_act_kernel<2><<<grid, 256, 0, st>>>((const bf16*)x, ldx, (bf16*)y, ldy, a, b, slope, nvox, c); break;
    default: affine_act_kernel<3><<<grid, 256, 0, st>>>((const bf16*)x, ldx, (bf16*)y, ldy, a, b, slope, nvox, c); break;
  }
  B21_LAUNCH_CHECK("affine_act_kernel");
  return B21_OK;
}

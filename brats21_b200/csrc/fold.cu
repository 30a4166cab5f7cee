// fold.cu — host-side glue kernels of the "folded EvoNorm" inference path.
//
// EvoNorm-S0 (networks/equiunet2021.py:48-52,95-105) is  y = x*sigmoid(x) * a_c + b_c  with
// a_c = gamma_c / sqrt(var_unbiased[group(c)] + eps), b_c = beta_c: only the per-channel AFFINE depends on the global
// group statistics, the non-linearity does not.  The conv epilogues therefore store S = swish(conv + bias) directly
// (one bf16 write, no separate normalisation pass over HBM) together with the group statistics, and the affine
// (A, B) — including the ResidualSE gate s_c, since (a S + b) s = (a s) S + (b s) — is folded into whatever consumes
// the tensor:
//   * a following convolution multiplies its weights by A[n][ci] (per-sample packed weights) and adds the response
//     to the constant image B, which near the zero-padded border depends on which taps fall inside the volume:
//     a bias TABLE T[n][border class (3x3x3)][co] = bias[co] + sum_ci B[n][ci] * sum_{valid taps} W[co][ci][tap];
//   * the MaxAvgPool (MONAI, equiunet2021.py:261) applies A, B on the fly (max for A >= 0, min for A < 0);
//   * trilinear up-sampling commutes with the affine (interpolation weights sum to 1);
//   * the 1x1 heads take W*A and bias + W*B.
// Kernels here: evo_se_affine (statistics [+ SE MLP] -> A, B), border_weight_sums / bias_table, affine_pool.
#include "ptx.cuh"
#include "host_common.h"

namespace b21 {

__device__ __forceinline__ void f_unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 f_pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------------------------------------ statistics -> (A, B)
// One block per sample.  A[n][c] = a_c * s_c, B[n][c] = beta_c * s_c with s_c = 1 when there is no SE gate, else the
// MONAI ResidualSELayer gate 1 + sigmoid(W2 relu(W1 m + b1) + b2) on m_c = a_c * mean(S_c) + beta_c.
__global__ void __launch_bounds__(1024) evo_se_affine_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, const float* __restrict__ chan_sum,
                                     const float* __restrict__ w1, const float* __restrict__ b1,
                                     const float* __restrict__ w2, const float* __restrict__ b2,
                                     float* __restrict__ A, float* __restrict__ B, int ldab, int N, int C, int Hd,
                                     long long nvox, float eps) {
  extern __shared__ float sm[];  // a[C], m[C], hid[Hd], w1[Hd*C], w2[C*Hd]
  __shared__ double sred[16];    // per (group, {sum, sumsq}) totals over the contention-spreading slots
  float* sa = sm;
  float* m = sm + C;
  float* hid = sm + 2 * C;
  float* sw1 = hid + Hd;
  float* sw2 = sw1 + Hd * C;
  const int n = blockIdx.x;
  const int gsz = C / 8;
  if (chan_sum) {  // the MLP weights are cold in L2 by now: fetch them with all loads in flight, under the stats reduction
    for (int i = threadIdx.x; i < Hd * C; i += blockDim.x) {
      sw1[i] = __ldg(w1 + i);
      sw2[i] = __ldg(w2 + i);
    }
  }
  if (threadIdx.x < 16) sred[threadIdx.x] = 0.0;
  __syncthreads();
  for (int i = threadIdx.x; i < B21_STAT_SLOTS * 16; i += blockDim.x) {  // all slot loads in flight at once
    double v = stats[(size_t(i >> 4) * N + n) * 16 + (i & 15)];
    v += __shfl_xor_sync(0xffffffffu, v, 16);  // lanes l and l + 16 hold the same (group, k)
    if ((threadIdx.x & 31) < 16) atomicAdd(&sred[i & 15], v);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / gsz;
    const double s = sred[g * 2], q = sred[g * 2 + 1];
    const double cnt = double(nvox) * gsz;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const double varu = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;  // torch.var default: unbiased
    const float a = gamma[c] * float(1.0 / sqrt(varu + double(eps)));
    sa[c] = a;
    if (chan_sum) {
      m[c] = a * (chan_sum[size_t(n) * C + c] / float(nvox)) + beta[c];
    } else {
      A[size_t(n) * ldab + c] = a;
      B[size_t(n) * ldab + c] = beta[c];
    }
  }
  if (!chan_sum) return;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < Hd; j += nw) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += sw1[j * C + c] * m[c];
    s = warp_sum(s);
    if (lane == 0) hid[j] = fmaxf(s + b1[j], 0.f);
  }
  __syncthreads();
  for (int c = warp; c < C; c += nw) {
    float s = 0.f;
    for (int j = lane; j < Hd; j += 32) s += sw2[c * Hd + j] * hid[j];
    s = warp_sum(s);
    if (lane == 0) {
      const float gate = 1.f + 1.f / (1.f + expf(-(s + b2[c])));
      A[size_t(n) * ldab + c] = sa[c] * gate;
      B[size_t(n) * ldab + c] = beta[c] * gate;
    }
  }
}

// ------------------------------------------------------------------------------------------ border weight sums
// ws[cls][ci][co] = sum over the taps that stay inside the volume for border class cls = cd*9 + ch*3 + cw
// (c = 0: first voxel of the axis, 1: interior, 2: last voxel) of w[co][ci][tap];  taps == 1: ws[0][ci][co] = w.
__global__ void border_weight_sums_kernel(const float* __restrict__ w, float* __restrict__ ws, int cout, int cin,
                                          int taps) {
  const int ncls = taps == 27 ? 27 : 1;
  const long long total = (long long)ncls * cin * cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = int(i % cout);
    const int ci = int((i / cout) % cin);
    const int cls = int(i / ((long long)cout * cin));
    const float* wp = w + (size_t(co) * cin + ci) * taps;
    float s = 0.f;
    if (taps == 1) {
      s = wp[0];
    } else {
      const int cd = cls / 9, ch = (cls / 3) % 3, cw = cls % 3;
      for (int kd = 0; kd < 3; ++kd) {
        if ((cd == 0 && kd == 0) || (cd == 2 && kd == 2)) continue;
        for (int kh = 0; kh < 3; ++kh) {
          if ((ch == 0 && kh == 0) || (ch == 2 && kh == 2)) continue;
          for (int kw = 0; kw < 3; ++kw) {
            if ((cw == 0 && kw == 0) || (cw == 2 && kw == 2)) continue;
            s += wp[kd * 9 + kh * 3 + kw];
          }
        }
      }
    }
    ws[i] = s;
  }
}

// T[n][cls][co] = bias[co] + sum_ci ws[cls][ci][co] * B[n][ci]      grid (ncls, N), threads over co
__global__ void bias_table_kernel(const float* __restrict__ ws, const float* __restrict__ bias,
                                  const float* __restrict__ B, int ldab, float* __restrict__ T, int cout, int cin,
                                  int ncls) {
  const int cls = blockIdx.x, n = blockIdx.y;
  for (int co = threadIdx.x; co < cout; co += blockDim.x) {
    float s = bias ? bias[co] : 0.f;
    const float* wp = ws + size_t(cls) * cin * cout + co;
    for (int ci = 0; ci < cin; ++ci) s = fmaf(wp[size_t(ci) * cout], __ldg(B + size_t(n) * ldab + ci), s);
    T[(size_t(n) * ncls + cls) * cout + co] = s;
  }
}

// ------------------------------------------------------------------------------------------ affine + pool
// pooled[n][dp][hp][wp][0:C | C:2C] = [max | mean] over the 2x2x2 block of (A[n][c] * S + B[n][c]) (mode 2), or
// max only (mode 1).  max(A s + B) = A * (A >= 0 ? max s : min s) + B.  One thread per (pooled voxel, 8 channels).
__global__ void __launch_bounds__(256) affine_pool_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                          const float* __restrict__ A, const float* __restrict__ B,
                                                          int ldab, __nv_bfloat16* __restrict__ pooled, int ldpool,
                                                          int mode, int N, int D, int H, int W, int C) {
  const int chunks = C >> 3;
  const int Dp = D >> 1, Hp = H >> 1, Wp = W >> 1;
  const long long total = (long long)N * Dp * Hp * Wp * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ck = int(i % chunks);
    long long v = i / chunks;
    const int wp = int(v % Wp); v /= Wp;
    const int hp = int(v % Hp); v /= Hp;
    const int dp = int(v % Dp);
    const int n = int(v / Dp);
    uint4 raw[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const size_t vox = ((size_t(n) * D + (2 * dp + (t >> 2))) * H + (2 * hp + ((t >> 1) & 1))) * W + (2 * wp + (t & 1));
      raw[t] = __ldg(reinterpret_cast<const uint4*>(x + vox * ldx + ck * 8));
    }
    float mx[8], mn[8], sm[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { mx[j] = -INFINITY; mn[j] = INFINITY; sm[j] = 0.f; }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float f[8];
      f_unpack8(raw[t], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { mx[j] = fmaxf(mx[j], f[j]); mn[j] = fminf(mn[j], f[j]); sm[j] += f[j]; }
    }
    float om[8], oa[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = __ldg(A + size_t(n) * ldab + ck * 8 + j), b = __ldg(B + size_t(n) * ldab + ck * 8 + j);
      om[j] = fmaf(a, a >= 0.f ? mx[j] : mn[j], b);
      oa[j] = fmaf(a, sm[j] * 0.125f, b);
    }
    const size_t pv = ((size_t(n) * Dp + dp) * Hp + hp) * Wp + wp;
    *reinterpret_cast<uint4*>(pooled + pv * ldpool + ck * 8) = f_pack8(om);
    if (mode == 2) *reinterpret_cast<uint4*>(pooled + pv * ldpool + C + ck * 8) = f_pack8(oa);
  }
}

}  // namespace b21

using namespace b21;
typedef __nv_bfloat16 bf16;

extern "C" int b21_evo_se_affine(const double* stats, const float* gamma, const float* beta, const float* chan_sum,
                                 const float* w1, const float* b1, const float* w2, const float* b2, float* a_out,
                                 float* b_out, int ldab, int n, int c, int hidden, long long nvox, float eps,
                                 void* stream) {
  B21_CHECK_ARG(stats && gamma && beta && a_out && b_out, "evo_se_affine: null pointer");
  B21_CHECK_ARG(n > 0 && c > 0 && c % 8 == 0 && ldab >= c && nvox > 0, "evo_se_affine: bad sizes");
  if (chan_sum) B21_CHECK_ARG(w1 && b1 && w2 && b2 && hidden > 0, "evo_se_affine: SE gate needs its MLP");
  const size_t smem = sizeof(float) * (2 * c + (chan_sum ? hidden + 2 * size_t(hidden) * c : 0));
  B21_CHECK_ARG(smem <= 200 * 1024, "evo_se_affine: SE MLP of %d channels does not fit in shared memory", c);
  static bool attr_set = false;
  if (!attr_set) {
    B21_CUDA(cudaFuncSetAttribute(evo_se_affine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  // the SE variant stages up to 147 KB of MLP weights and runs two dependent mat-vecs in ONE block per sample: it is
  // latency-bound, so it gets 32 warps
  evo_se_affine_kernel<<<n, chan_sum ? 1024 : 256, smem, (cudaStream_t)stream>>>(
      stats, gamma, beta, chan_sum, w1, b1, w2, b2, a_out, b_out, ldab, n, c, hidden, nvox, eps);
  B21_LAUNCH_CHECK("evo_se_affine_kernel");
  return B21_OK;
}

extern "C" int b21_border_weight_sums(const float* w, float* ws, int cout, int cin, int taps, void* stream) {
  B21_CHECK_ARG(w && ws && cout > 0 && cin > 0 && (taps == 1 || taps == 27), "border_weight_sums: bad args");
  const long long total = (long long)(taps == 27 ? 27 : 1) * cin * cout;
  const int blocks = int((total + 255) / 256) < 1024 ? int((total + 255) / 256) : 1024;
  border_weight_sums_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, ws, cout, cin, taps);
  B21_LAUNCH_CHECK("border_weight_sums_kernel");
  return B21_OK;
}

extern "C" int b21_bias_table(const float* ws, const float* bias, const float* b_in, int ldab, float* table, int n,
                              int cout, int cin, int ncls, void* stream) {
  B21_CHECK_ARG(ws && b_in && table && n > 0 && cout > 0 && cin > 0 && (ncls == 1 || ncls == 27), "bias_table: bad args");
  dim3 grid(ncls, n);
  bias_table_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(ws, bias, b_in, ldab, table, cout, cin, ncls);
  B21_LAUNCH_CHECK("bias_table_kernel");
  return B21_OK;
}

extern "C" int b21_affine_pool(const void* x, int ldx, const float* a_in, const float* b_in, int ldab, void* pooled,
                               int ldpool, int mode, int n, int d, int h, int w, int c, void* stream) {
  B21_CHECK_ARG(x && a_in && b_in && pooled, "affine_pool: null pointer");
  B21_CHECK_ARG(c % 8 == 0 && ldx % 8 == 0 && ldpool % 8 == 0 && (mode == 1 || mode == 2), "affine_pool: bad C/ld/mode");
  B21_CHECK_ARG(d % 2 == 0 && h % 2 == 0 && w % 2 == 0 && ldpool >= (mode == 2 ? 2 * c : c), "affine_pool: bad dims");
  const long long items = (long long)n * (d / 2) * (h / 2) * (w / 2) * (c / 8);
  long long blocks = (items + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  affine_pool_kernel<<<int(blocks), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, a_in, b_in, ldab, (bf16*)pooled,
                                                                    ldpool, mode, n, d, h, w, c);
  B21_LAUNCH_CHECK("affine_pool_kernel");
  return B21_OK;
}

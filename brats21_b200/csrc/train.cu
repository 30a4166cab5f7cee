// train.cu — HBM-bound kernels of the training step (backward of the network body, Dice criterion).
//
//   dice_fwd / dice_finalize / dice_bwd   monai DiceLoss(sigmoid, squared_pred, batch=True[, jaccard]) as built at
//                                         src/definer.py:184-203 and averaged over heads at learning/engine.py:322-330
//   norm_bwd_reduce / coeffs / apply      backward of EvoNorm3D-S0 (+ MONAI ResidualSELayer folded in) and of
//                                         GroupNorm(8)+ReLU (equiunet2021.py:48-105,204-205; factory.py:182)
//   pool_bwd                              backward of MaxPool3d(2,2) / MONAI MaxAvgPool (equiunet2020.py:433, 2021.py:261)
//   upsample2x_bwd, upsample_f32_bwd      adjoint of nn.Upsample(trilinear, align_corners=True)
//   head_conv_bwd                         backward of the 1x1 class heads (outconv / deep heads)
//   add_inplace                           gradient fan-in (dst += src) on channels-last bf16
// Layouts as in elementwise.cu: NDHWC bf16 with a channel stride `ld`, 16-byte vectors of 8 channels, grids sized
// so that a thread keeps the same 8 channels for its whole grid-stride loop (per-channel partial sums in registers).
#include "ptx.cuh"
#include "host_common.h"
#include "reduce.cuh"

namespace b21 {

__device__ __forceinline__ void unpack8t(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8t(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
// 1 / (1 + 2^(-x log2 e)) on the MUFU pipe (ex2 + rcp, flush-to-zero: no range fix-up instructions)
__device__ __forceinline__ float sigmoidf_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}

static inline int grid_cap(long long blocks, int per_sm = 8) {
  const long long cap = (long long)num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return int(blocks);
}

// =========================================================================================== Dice
// sums[k][3] (double) += { sum t*p, sum t*t, sum p*p } over batch and space, p = sigmoid(logit)
__global__ void __launch_bounds__(256) dice_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                       double* __restrict__ sums, int N, int K, long long nvox) {
  const int k = blockIdx.y;
  float a = 0.f, b = 0.f, c = 0.f;
  for (int n = 0; n < N; ++n) {
    const float* x = logits + (size_t(n) * K + k) * nvox;
    const float* t = target + (size_t(n) * K + k) * nvox;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (long long)gridDim.x * blockDim.x) {
      const float p = 1.f / (1.f + expf(-x[i]));
      const float tv = t[i];
      a = fmaf(tv, p, a);
      b = fmaf(tv, tv, b);
      c = fmaf(p, p, c);
    }
  }
  __shared__ float red[3][8];
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = a; red[1][warp] = b; red[2][warp] = c; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += double(red[threadIdx.x][w]);
    atomicAdd(sums + k * 3 + threadIdx.x, s);
  }
}

// loss += weight * mean_k f_k ;  coef[k] = {c_t, c_p, c_pt}: d f_k / d p = c_t * t + c_p * p   (per channel)
__global__ void dice_finalize_kernel(const double* __restrict__ sums, float* __restrict__ loss, float* __restrict__ coef,
                                     int K, int jaccard, float smooth_nr, float smooth_dr, float weight) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double acc = 0.0;
  for (int k = 0; k < K; ++k) {
    const double I = sums[k * 3], G = sums[k * 3 + 1], P = sums[k * 3 + 2];
    double den = G + P;
    if (jaccard) den = 2.0 * (den - I);
    const double num = 2.0 * I + double(smooth_nr);
    const double D = den + double(smooth_dr);
    acc += 1.0 - num / D;
    // f = 1 - num/D ; dnum/dp = 2 t ; dD/dp = 2 p (dice) or 2 (2 p - t) (jaccard)
    double ct, cp;
    if (!jaccard) {
      ct = -2.0 / D;
      cp = 2.0 * num / (D * D);
    } else {
      ct = -2.0 / D - 2.0 * num / (D * D);
      cp = 4.0 * num / (D * D);
    }
    coef[k * 2] = float(ct / K);
    coef[k * 2 + 1] = float(cp / K);
  }
  atomicAdd(loss, float(acc / K) * weight);
}

// Cross-entropy half of DiceCELoss (learning/losses.py:470-595): torch.nn.CrossEntropyLoss(reduction="mean") over the
// K class logits of every voxel against y = argmax_k target (first maximum; `ce()` at losses.py:562-577).
__device__ __forceinline__ int first_argmax(const float* t, int K) {
  int y = 0;
  float best = t[0];
  for (int k = 1; k < K; ++k)
    if (t[k] > best) { best = t[k]; y = k; }
  return y;
}

constexpr int kMaxClasses = 16;

__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                     double* __restrict__ sum, int N, int K, long long nvox) {
  float acc = 0.f;
  const long long total = (long long)N * nvox;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / nvox, v = i - n * nvox;
    const float* x = logits + size_t(n) * K * nvox + v;
    const float* t = target + size_t(n) * K * nvox + v;
    float xs[kMaxClasses], ts[kMaxClasses];
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) {
      xs[k] = x[size_t(k) * nvox];
      ts[k] = t[size_t(k) * nvox];
      m = fmaxf(m, xs[k]);
    }
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(xs[k] - m);
    acc += m + logf(se) - xs[first_argmax(ts, K)];
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += double(red[w]);
    atomicAdd(sum, s);
  }
}

__global__ void ce_finalize_kernel(const double* __restrict__ sum, float* __restrict__ loss, double count, float weight) {
  if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(loss, float(sum[0] / count) * weight);
}

// dlogits = gscale * [ (c_t t + c_p p) * p (1 - p)  +  ce_w * (softmax_k - [k == y]) ]
__global__ void __launch_bounds__(256) dice_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                       const float* __restrict__ coef, const float* __restrict__ gout,
                                                       float gscale, float ce_w, float* __restrict__ dlogits, int N, int K,
                                                       long long nvox) {
  const long long total = (long long)N * K * nvox;
  const float g = gscale * (gout ? __ldg(gout) : 1.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = int((i / nvox) % K);
    const float p = 1.f / (1.f + expf(-logits[i]));
    const float t = target[i];
    float d = (__ldg(coef + 2 * k) * t + __ldg(coef + 2 * k + 1) * p) * p * (1.f - p);
    if (ce_w != 0.f) {
      const long long base = i - (long long)k * nvox;  // channel 0 of this voxel
      float ts[kMaxClasses];
      float m = -INFINITY;
      for (int j = 0; j < K; ++j) {
        ts[j] = target[base + (long long)j * nvox];
        m = fmaxf(m, logits[base + (long long)j * nvox]);
      }
      float se = 0.f;
      for (int j = 0; j < K; ++j) se += expf(logits[base + (long long)j * nvox] - m);
      const float sm = expf(logits[i] - m) / se;
      d += ce_w * (sm - (first_argmax(ts, K) == k ? 1.f : 0.f));
    }
    dlogits[i] = g * d;
  }
}

// =========================================================================================== norm backward
// MODE 0: y = relu(gamma (z - mu) r + beta)       (GroupNorm(8) + ReLU)
// MODE 1: y = z sigmoid(z) r gamma + beta          (EvoNorm3D-S0), optionally followed by out = y * sc[n][c]
// reduce: per (n, c)   R1 = sum dyA,  R2 = sum dyA * u,  R3 = sum u
//         MODE 0: dyA = dy * [y > 0], u = z ;  MODE 1: dyA = dy, u = z sigmoid(z)
constexpr int kRedSlots = 32;  // copies of the reduction table (norm_bwd_reduce), folded by norm_bwd_coeffs
constexpr int kColSlots = 16;  // copies of the bias-gradient row (norm_bwd_apply), folded by its last block
struct NormFwdCoef {  // how the forward pass maps z to y for channel c of sample n: y = u * a + b (MODE 1) / z*a+b (0)
  float a, b;
};

template <int MODE>
__global__ void __launch_bounds__(256, 3) norm_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, int lddy,
                                                              const __nv_bfloat16* __restrict__ z, int ldz,
                                                              const float* __restrict__ fa, const float* __restrict__ fb,
                                                              double* __restrict__ red, long long nvox, int C) {
  extern __shared__ float sm[];  // [3][C] block sums, filled by block_chunk_reduce<3>
  const int n = blockIdx.y;
  const int chunks = C >> 3;
  const long long total = nvox * chunks;
  const long long T = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ck = int(i % chunks);
  float a[8], b[8], rr[3][8];
  float (&r1)[8] = rr[0], (&r2)[8] = rr[1], (&r3)[8] = rr[2];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = MODE == 0 ? fa[size_t(n) * C + ck * 8 + j] : 0.f;
    b[j] = MODE == 0 ? fb[size_t(n) * C + ck * 8 + j] : 0.f;
    r1[j] = r2[j] = r3[j] = 0.f;
  }
  const __nv_bfloat16* dyn = dy + size_t(n) * nvox * lddy;
  const __nv_bfloat16* zn = z + size_t(n) * nvox * ldz;
  constexpr int U = 4;  // independent 16 B loads in flight per thread and operand
  const long long vstep = T / chunks;  // T is a multiple of `chunks`: no 64-bit division in the loop
  long long v0 = i / chunks;
  for (; i < total; i += T * U, v0 += vstep * U) {
    uint4 gv[U], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + u * T;
      if (iu < total) {
        const long long v = v0 + u * vstep;
        gv[u] = *reinterpret_cast<const uint4*>(dyn + v * lddy + ck * 8);
        xv[u] = *reinterpret_cast<const uint4*>(zn + v * ldz + ck * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u * T < total) {
        float g[8], x[8];
        unpack8t(gv[u], g);
        unpack8t(xv[u], x);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (MODE == 0) {
            const float gg = fmaf(x[j], a[j], b[j]) > 0.f ? g[j] : 0.f;
            r1[j] += gg;
            r2[j] = fmaf(gg, x[j], r2[j]);
          } else {
            const float uu = x[j] * sigmoidf_fast(x[j]);
            r1[j] += g[j];
            r2[j] = fmaf(g[j], uu, r2[j]);
            r3[j] += uu;
          }
        }
      }
    }
  }
  block_chunk_reduce<3>(rr, chunks, C, sm);
  // kRedSlots copies of the table: ~1200 blocks adding to the same 3*C addresses serialise in the L2 atomic unit
  // (measured 75 us per launch, whatever the tensor size); with 32 slots a launch sees < 40 adds per address
  double* rs = red + (size_t(blockIdx.x % kRedSlots) * gridDim.y + n) * C * 3;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    for (int q = 0; q < 3; ++q) atomicAdd(rs + size_t(c) * 3 + q, double(sm[q * C + c]));
}

// One block.  From the reductions, the forward statistics and (optionally) the squeeze-excite state, produce the
// per-(n, c) coefficient table of the apply pass   dz = (dy * p0 + p1) * D(z) - p2 * z + p3
// (D = swish' for MODE 1, the ReLU mask for MODE 0) and ACCUMULATE the parameter gradients.
struct NormBwdArgs {
  const double* red;      // [N][C][3] (copy 0 of redw after the fold)
  double* redw;           // [kRedSlots][N][C][3] as written by norm_bwd_reduce
  const double* stats;    // forward statistics, double[SLOTS][N][8][2]
  const float* gamma;     // [C]
  const float* beta;      // [C]
  float* fa;              // out (MODE 0 only needs it earlier; recomputed here for the apply pass) [N][C]
  float* fb;
  float* coef;            // out [N][C][4]
  float* dgamma;          // accumulate [C]
  float* dbeta;           // accumulate [C]
  // squeeze-excite (MODE 1 only; se_w1 == NULL -> none)
  const float* se_scale;  // [N][C] = 1 + sigmoid(.)
  const float* se_mean;   // [N][C] channel means of y (forward)
  const float* se_w1; const float* se_b1; const float* se_w2; const float* se_b2;
  float* d_w1; float* d_b1; float* d_w2; float* d_b2;  // accumulate
  int N, C, hidden;
  long long nvox;
  float eps;
};

// One block:  y[n] = sum_m M[m][n] x[m]   and   dM[m][n] += x[m] * z[n]    (M, dM row-major [rows][cols], cols % 4 == 0;
// x, y in shared memory, z anywhere).  Thread t owns four columns and every S-th row: its accumulators live in
// registers, loads are 16 B and independent (a thread per output looping over a whole row / column of a 384 x 192
// matrix took 30 us, shared-memory float atomics -- compare-and-swap loops -- 270 us).  part: blockDim.x * 4 floats.
__device__ __forceinline__ void matvec_t_outer(const float* __restrict__ M, float* dM, const float* x, const float* z,
                                               float* y, float* part, int rows, int cols) {
  const int tid = threadIdx.x, q = cols >> 2;
  int S = int(blockDim.x) / q;
  if (S > rows) S = rows;
  if (S < 1) S = 1;  // cols > 4 * blockDim.x: column groups are walked in passes
  for (int g0 = 0; g0 < q; g0 += int(blockDim.x)) {
    const int g = g0 + (S > 1 ? tid % q : tid), slice = S > 1 ? tid / q : 0;
    const bool on = g < q && slice < S;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) {
      const float4 zv = make_float4(z[4 * g], z[4 * g + 1], z[4 * g + 2], z[4 * g + 3]);
#pragma unroll 4
      for (int m = slice; m < rows; m += S) {
        const size_t off = size_t(m) * cols + 4 * g;
        const float4 w = *reinterpret_cast<const float4*>(M + off);
        float4 d = *reinterpret_cast<const float4*>(dM + off);
        const float xm = x[m];
        acc.x = fmaf(w.x, xm, acc.x); acc.y = fmaf(w.y, xm, acc.y); acc.z = fmaf(w.z, xm, acc.z); acc.w = fmaf(w.w, xm, acc.w);
        d.x = fmaf(xm, zv.x, d.x); d.y = fmaf(xm, zv.y, d.y); d.z = fmaf(xm, zv.z, d.z); d.w = fmaf(xm, zv.w, d.w);
        *reinterpret_cast<float4*>(dM + off) = d;
      }
      if (S == 1) { y[4 * g] = acc.x; y[4 * g + 1] = acc.y; y[4 * g + 2] = acc.z; y[4 * g + 3] = acc.w; }
    }
    if (S > 1) {  // (then q <= blockDim.x and there is a single pass)
      if (on) *reinterpret_cast<float4*>(part + 4 * (slice * q + g)) = acc;
      __syncthreads();
      for (int n = tid; n < cols; n += blockDim.x) {
        float t = 0.f;
        for (int sl = 0; sl < S; ++sl) t += part[4 * (sl * q + (n >> 2)) + (n & 3)];
        y[n] = t;
      }
    }
  }
  __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(512) norm_bwd_coeffs_kernel(NormBwdArgs p) {
  extern __shared__ float sm[];  // mu[8], r[8], A[8], B[8], then per-channel scratch: dsc[C], du[C], hid[H], dh[H], dm[C]
  float* s_mu = sm;
  float* s_r = sm + 8;
  float* s_A = sm + 16;
  float* s_B = sm + 24;
  float* s_du = sm + 32;
  float* s_hid = s_du + p.C;
  float* s_dh = s_hid + p.hidden;
  float* s_dm = s_dh + p.hidden;
  float* s_part = s_dm + p.C;  // blockDim.x * 4 floats (squeeze-excite only)
  const int C = p.C, gsz = C / 8, tid = threadIdx.x, nt = blockDim.x;
  const double cnt = double(p.nvox) * gsz;
  {  // fold the kRedSlots copies of the reduction table into copy 0
    const int tot = p.N * C * 3;
    for (int i = tid; i < tot; i += nt) {
      double s = 0.0;
#pragma unroll 8
      for (int slot = 0; slot < kRedSlots; ++slot) s += p.redw[size_t(slot) * tot + i];
      p.redw[i] = s;
    }
    __syncthreads();
  }
  for (int n = 0; n < p.N; ++n) {
    if (tid < 8) {
      double s = 0.0, q = 0.0;
      for (int slot = 0; slot < B21_STAT_SLOTS; ++slot) {
        const double* st = p.stats + ((size_t(slot) * p.N + n) * 8 + tid) * 2;
        s += st[0];
        q += st[1];
      }
      const double mean = s / cnt;
      double var = q / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      if (MODE == 1) var = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
      s_mu[tid] = float(mean);
      s_r[tid] = float(1.0 / sqrt(var + double(p.eps)));
      s_A[tid] = 0.f;
      s_B[tid] = 0.f;
    }
    for (int c = tid; c < C; c += nt) s_dm[c] = 0.f;
    __syncthreads();
    const bool se = MODE == 1 && p.se_w1 != nullptr;
    if (se) {
      // dsc[c] = sum_v dy * y = (gamma r) R2 + beta R1 ;  du = dsc * s (1 - s), s = sc - 1
      for (int c = tid; c < C; c += nt) {
        const double* R = p.red + (size_t(n) * C + c) * 3;
        const float r = s_r[c / gsz];
        const float dsc = p.gamma[c] * r * float(R[1]) + p.beta[c] * float(R[0]);
        const float s = p.se_scale[size_t(n) * C + c] - 1.f;
        s_du[c] = dsc * s * (1.f - s);
      }
      for (int j = tid >> 5; j < p.hidden; j += nt >> 5) {  // one warp per hidden unit: coalesced rows of W1
        float h = 0.f;
        for (int c = tid & 31; c < C; c += 32) h = fmaf(p.se_w1[size_t(j) * C + c], p.se_mean[size_t(n) * C + c], h);
        h = warp_sum(h);
        if ((tid & 31) == 0) s_hid[j] = fmaxf(h + p.se_b1[j], 0.f);
      }
      __syncthreads();
      // dW2[c][j] += du[c] * hid[j] ; dh_raw[j] = sum_c W2[c][j] du[c]
      matvec_t_outer(p.se_w2, p.d_w2, s_du, s_hid, s_dh, s_part, C, p.hidden);
      for (int c = tid; c < C; c += nt) p.d_b2[c] += s_du[c];
      for (int j = tid; j < p.hidden; j += nt) {
        const float d = s_hid[j] > 0.f ? s_dh[j] : 0.f;
        s_dh[j] = d;
        p.d_b1[j] += d;
      }
      __syncthreads();
      // dW1[j][c] += dh[j] * mean[c] ; dm[c] = sum_j W1[j][c] dh[j]
      matvec_t_outer(p.se_w1, p.d_w1, s_dh, p.se_mean + size_t(n) * C, s_dm, s_part, p.hidden, C);
      for (int c = tid; c < C; c += nt) s_dm[c] /= float(p.nvox);  // gradient of every voxel of channel c through the mean
      __syncthreads();
    }
    // per-channel sums of the gradient w.r.t. the norm output y: dyE = dy * sc + dmv
    for (int c = tid; c < C; c += nt) {
      const double* R = p.red + (size_t(n) * C + c) * 3;
      const int g = c / gsz;
      const float sc = se ? p.se_scale[size_t(n) * C + c] : 1.f;
      const float dmv = s_dm[c];
      const float gam = p.gamma[c], r = s_r[g], mu = s_mu[g];
      if (MODE == 1) {
        const float S = sc * float(R[1]) + dmv * float(R[2]);          // sum dyE * swish(z)
        const float Sb = sc * float(R[0]) + dmv * float(p.nvox);       // sum dyE
        p.dgamma[c] += r * S;
        p.dbeta[c] += Sb;
        atomicAdd(&s_A[g], gam * S);
      } else {
        const float Q1 = float(R[0]), Q2 = float(R[1]);                // sum dyM, sum dyM * z
        const float xh = (Q2 - mu * Q1) * r;                           // sum dyM * xhat
        p.dgamma[c] += xh;
        p.dbeta[c] += Q1;
        atomicAdd(&s_A[g], gam * Q1);
        atomicAdd(&s_B[g], gam * xh);
      }
    }
    __syncthreads();
    for (int c = tid; c < C; c += nt) {
      const int g = c / gsz;
      const float gam = p.gamma[c], r = s_r[g], mu = s_mu[g];
      float* co = p.coef + (size_t(n) * C + c) * 4;
      if (MODE == 1) {
        const float sc = se ? p.se_scale[size_t(n) * C + c] : 1.f;
        const float Kg = s_A[g] * r * r * r / float(cnt - 1.0);
        co[0] = sc * gam * r;
        co[1] = s_dm[c] * gam * r;
        co[2] = Kg;
        co[3] = Kg * mu;
      } else {
        const float inv = 1.f / float(cnt);
        co[0] = gam * r;
        co[1] = 0.f;
        co[2] = r * r * s_B[g] * inv;
        co[3] = -r * s_A[g] * inv + mu * r * r * s_B[g] * inv;
        p.fa[size_t(n) * C + c] = gam * r;
        p.fb[size_t(n) * C + c] = p.beta[c] - mu * gam * r;
      }
    }
    __syncthreads();
  }
}

// forward affine (a, b) of GroupNorm for the ReLU mask of the reduce pass (MODE 0)
__global__ void gn_fwd_affine_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ fa, float* __restrict__ fb, int N,
                                     int C, long long nvox, float eps) {
  const int n = blockIdx.x, gsz = C / 8;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / gsz;
    double s = 0.0, q = 0.0;
    for (int slot = 0; slot < B21_STAT_SLOTS; ++slot) {
      const double* st = stats + ((size_t(slot) * N + n) * 8 + g) * 2;
      s += st[0];
      q += st[1];
    }
    const double cnt = double(nvox) * gsz, mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float r = float(1.0 / sqrt(var + double(eps)));
    fa[size_t(n) * C + c] = gamma[c] * r;
    fb[size_t(n) * C + c] = beta[c] - float(mean) * gamma[c] * r;
  }
}

// dz = (dy * p0 + p1) * D(z) - p2 * z + p3 ; optional column sums of dz (bias gradient of the producing conv)
template <int MODE>
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(const __nv_bfloat16* dy, int lddy,
                                                             const __nv_bfloat16* __restrict__ z, int ldz,
                                                             __nv_bfloat16* dz, int lddz, const float* __restrict__ coef,
                                                             const float* __restrict__ fa, const float* __restrict__ fb,
                                                             float* colsum, float* cs_ws, unsigned int* cs_count,
                                                             long long nvox, int C) {
  extern __shared__ float sm[];  // [C] column sums
  const int n = blockIdx.y;
  const int chunks = C >> 3;
  const long long total = nvox * chunks;
  const long long T = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ck = int(i % chunks);
  float p0[8], p1[8], p2[8], p3[8], a[8], b[8], accv[1][8];
  float (&acc)[8] = accv[0];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float* co = coef + (size_t(n) * C + ck * 8 + j) * 4;
    p0[j] = co[0]; p1[j] = co[1]; p2[j] = co[2]; p3[j] = co[3];
    a[j] = MODE == 0 ? fa[size_t(n) * C + ck * 8 + j] : 0.f;
    b[j] = MODE == 0 ? fb[size_t(n) * C + ck * 8 + j] : 0.f;
    acc[j] = 0.f;
  }
  const __nv_bfloat16* dyn = dy + size_t(n) * nvox * lddy;
  const __nv_bfloat16* zn = z + size_t(n) * nvox * ldz;
  __nv_bfloat16* dzn = dz + size_t(n) * nvox * lddz;
  constexpr int U = 4;  // independent 16 B loads in flight per thread and operand (dz may alias dy: own elements only)
  const long long vstep = T / chunks;  // T is a multiple of `chunks`: no 64-bit division in the loop
  long long v0 = i / chunks;
  for (; i < total; i += T * U, v0 += vstep * U) {
    uint4 gv[U], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + u * T;
      if (iu < total) {
        const long long v = v0 + u * vstep;
        gv[u] = *reinterpret_cast<const uint4*>(dyn + v * lddy + ck * 8);
        xv[u] = *reinterpret_cast<const uint4*>(zn + v * ldz + ck * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + u * T;
      if (iu < total) {
        const long long v = v0 + u * vstep;
        float g[8], x[8], o[8];
        unpack8t(gv[u], g);
        unpack8t(xv[u], x);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float d;
          if (MODE == 0) {
            d = fmaf(x[j], a[j], b[j]) > 0.f ? 1.f : 0.f;
          } else {
            const float s = sigmoidf_fast(x[j]);
            d = s * fmaf(x[j], 1.f - s, 1.f);
          }
          o[j] = fmaf(fmaf(g[j], p0[j], p1[j]), d, fmaf(-p2[j], x[j], p3[j]));
        }
        const uint4 pk = pack8t(o);
        if (colsum) {
          float q[8];
          unpack8t(pk, q);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += q[j];
        }
        *reinterpret_cast<uint4*>(dzn + v * lddz + ck * 8) = pk;
      }
    }
  }
  if (colsum) {
    block_chunk_reduce<1>(accv, chunks, C, sm);
    // kColSlots copies of the row (see norm_bwd_reduce); the last block to finish folds them into colsum
    float* cs = cs_ws + size_t(blockIdx.x % kColSlots) * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(cs + c, sm[c]);
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(cs_count, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (last) {
      __threadfence();
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
#pragma unroll
        for (int slot = 0; slot < kColSlots; ++slot) t += __ldcg(cs_ws + size_t(slot) * C + c);
        colsum[c] += t;
      }
    }
  }
}

// =========================================================================================== pooling backward
// y: pooled input (full resolution, as the forward saw it); dpool: [N, D/2, H/2, W/2, C or 2C]; add: optional
// extra gradient of y (same shape as y); dy = add + max-route(dpool[:C]) + dpool[C:] / 8
__global__ void __launch_bounds__(256) pool_bwd_kernel(const __nv_bfloat16* __restrict__ y, int ldy,
                                                       const __nv_bfloat16* __restrict__ dpool, int ldp,
                                                       const __nv_bfloat16* add, int ldadd, __nv_bfloat16* dy, int lddy,
                                                       int mode, int N, int D, int H, int W, int C) {
  const int chunks = C >> 3;
  const int Dp = D >> 1, Hp = H >> 1, Wp = W >> 1;
  const long long total = (long long)N * Dp * Hp * Wp * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ck = int(i % chunks);
    long long v = i / chunks;
    const int wp = int(v % Wp); v /= Wp;
    const int hp = int(v % Hp); v /= Hp;
    const int dp = int(v % Dp);
    const int n = int(v / Dp);
    const size_t pv = ((size_t(n) * Dp + dp) * Hp + hp) * Wp + wp;
    float gm[8], ga[8], best[8];
    int arg[8];
    unpack8t(*reinterpret_cast<const uint4*>(dpool + pv * ldp + ck * 8), gm);
    if (mode == 2) {
      unpack8t(*reinterpret_cast<const uint4*>(dpool + pv * ldp + C + ck * 8), ga);
#pragma unroll
      for (int j = 0; j < 8; ++j) ga[j] *= 0.125f;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) ga[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = 0; }
    size_t vox[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      vox[t] = ((size_t(n) * D + (2 * dp + (t >> 2))) * H + (2 * hp + ((t >> 1) & 1))) * W + (2 * wp + (t & 1));
      float f[8];
      unpack8t(*reinterpret_cast<const uint4*>(y + vox[t] * ldy + ck * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (f[j] > best[j]) { best[j] = f[j]; arg[j] = t; }  // first maximum in scan order, as torch max_pool3d
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float o[8];
      if (add) unpack8t(*reinterpret_cast<const uint4*>(add + vox[t] * ldadd + ck * 8), o);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += ga[j] + (arg[j] == t ? gm[j] : 0.f);
      *reinterpret_cast<uint4*>(dy + vox[t] * lddy + ck * 8) = pack8t(o);
    }
  }
}

// =========================================================================================== upsample backward
__device__ __forceinline__ void lerp_setup_t(int o, int I, int O, int& i0, int& i1, float& l1) {
  const float sc = (O > 1) ? float(I - 1) / float(O - 1) : 0.f;
  const float src = sc * float(o);
  i0 = int(src);
  if (i0 > I - 1) i0 = I - 1;
  i1 = i0 + (i0 < I - 1 ? 1 : 0);
  l1 = src - float(i0);
}
// taps of the adjoint along one axis: output indices o (and weights) that read input index i in the forward pass
template <int MAXT>
__device__ __forceinline__ int adjoint_taps(int i, int I, int O, int S, int* to, float* tw) {
  int cnt = 0;
  // forward: src(o) = o (I-1)/(O-1); input i is read by the outputs with src in (i-1, i+1)
  int lo = 0, hi = O - 1;
  if (I > 1) {
    const float rr = float(O - 1) / float(I - 1);
    lo = int(floorf(float(i - 1) * rr)) - 1;
    hi = int(ceilf(float(i + 1) * rr)) + 1;
    lo = lo < 0 ? 0 : lo;
    hi = hi > O - 1 ? O - 1 : hi;
  }
  (void)S;
  for (int o = lo; o <= hi; ++o) {
    int i0, i1;
    float l;
    lerp_setup_t(o, I, O, i0, i1, l);
    float wgt = 0.f;
    if (i0 == i) wgt += 1.f - l;
    if (i1 == i) wgt += l;
    if (wgt != 0.f && cnt < MAXT) { to[cnt] = o; tw[cnt] = wgt; ++cnt; }
  }
  return cnt;
}

// dx[n,d,h,w,c] = sum over the output voxels that interpolate from (d,h,w); dy is [n,2d,2h,2w,c]
__global__ void __launch_bounds__(128) upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int lddy,
                                                             __nv_bfloat16* __restrict__ dx, int lddx, int N, int D,
                                                             int H, int W, int C) {
  const int chunks = C >> 3;
  const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)N * D * H * W * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ck = int(i % chunks);
    long long v = i / chunks;
    const int w = int(v % W); v /= W;
    const int h = int(v % H); v /= H;
    const int d = int(v % D);
    const int n = int(v / D);
    int od[6], oh[6], ow[6];
    float wd[6], wh[6], ww[6];
    const int nd = adjoint_taps<6>(d, D, Do, 2, od, wd);
    const int nh = adjoint_taps<6>(h, H, Ho, 2, oh, wh);
    const int nw = adjoint_taps<6>(w, W, Wo, 2, ow, ww);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int a = 0; a < nd; ++a)
      for (int b = 0; b < nh; ++b) {
        const float wab = wd[a] * wh[b];
        const size_t rowv = ((size_t(n) * Do + od[a]) * Ho + oh[b]) * Wo;
        for (int c = 0; c < nw; ++c) {
          float f[8];
          unpack8t(__ldg(reinterpret_cast<const uint4*>(dy + (rowv + ow[c]) * lddy + ck * 8)), f);
          const float wt = wab * ww[c];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt, f[j], acc[j]);
        }
      }
    const size_t iv = ((size_t(n) * D + d) * H + h) * W + w;
    *reinterpret_cast<uint4*>(dx + iv * lddx + ck * 8) = pack8t(acc);
  }
}

// same adjoint on NCDHW fp32 planes with an integer factor S (deep-supervision heads)
__global__ void __launch_bounds__(128) upsample_f32_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                               int planes, int D, int H, int W, int S) {
  const int Do = S * D, Ho = S * H, Wo = S * W;
  const long long total = (long long)planes * D * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long v = i;
    const int w = int(v % W); v /= W;
    const int h = int(v % H); v /= H;
    const int d = int(v % D);
    const int pl = int(v / D);
    int od[24], oh[24], ow[24];
    float wd[24], wh[24], ww[24];
    const int nd = adjoint_taps<24>(d, D, Do, S, od, wd);
    const int nh = adjoint_taps<24>(h, H, Ho, S, oh, wh);
    const int nw = adjoint_taps<24>(w, W, Wo, S, ow, ww);
    const float* yp = dy + size_t(pl) * Do * Ho * Wo;
    float acc = 0.f;
    for (int a = 0; a < nd; ++a)
      for (int b = 0; b < nh; ++b) {
        const float wab = wd[a] * wh[b];
        const float* row = yp + (size_t(od[a]) * Ho + oh[b]) * Wo;
        for (int c = 0; c < nw; ++c) acc = fmaf(wab * ww[c], __ldg(row + ow[c]), acc);
      }
    dx[i] = acc;
  }
}

// =========================================================================================== head backward
// logits[n][k][v] = b[k] + sum_c w[k][c] * s[n][c] * x[n][v][c]  (s = optional SE scale folded into the head)
//   dx[n][v][c]   = s[n][c] * sum_k w[k][c] dl[n][k][v]                        (bf16, gradient w.r.t. x)
//   dws[n][k][c] += sum_v dl[n][k][v] * x[n][v][c]        (fp32; dW = s * dws, ds = sum_k w dws: tiny, done by caller)
//   db[k]        += sum_v dl[n][k][v]
template <int K>
__global__ void __launch_bounds__(256) head_conv_bwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                            const float* __restrict__ scale, const float* __restrict__ w,
                                                            const float* __restrict__ dl, __nv_bfloat16* __restrict__ dx,
                                                            int lddx, int accumulate, float* __restrict__ dws,
                                                            float* __restrict__ db, int slots, long long nvox, int C) {
  extern __shared__ float sm[];  // [K][C] partial dws + [K] db
  const int n = blockIdx.y;
  const int chunks = C >> 3;
  const long long total = nvox * chunks;
  const long long T = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ck = int(i % chunks);
  float ws[K][8], acc[K][8], accb[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    accb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ck * 8 + j;
      ws[k][j] = w[k * C + c] * (scale ? scale[size_t(n) * C + c] : 1.f);
      acc[k][j] = 0.f;
    }
  }
  const __nv_bfloat16* xn = x + size_t(n) * nvox * ldx;
  __nv_bfloat16* dxn = dx + size_t(n) * nvox * lddx;
  const float* dln = dl + size_t(n) * K * nvox;
  for (; i < total; i += T) {
    const long long v = i / chunks;
    float f[8], o[8], g[K];
    unpack8t(*reinterpret_cast<const uint4*>(xn + v * ldx + ck * 8), f);
#pragma unroll
    for (int k = 0; k < K; ++k) g[k] = __ldg(dln + size_t(k) * nvox + v);
    if (accumulate) unpack8t(*reinterpret_cast<const uint4*>(dxn + v * lddx + ck * 8), o);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (ck == 0) accb[k] += g[k];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = fmaf(ws[k][j], g[k], o[j]);
        acc[k][j] = fmaf(g[k], f[j], acc[k][j]);
      }
    }
    *reinterpret_cast<uint4*>(dxn + v * lddx + ck * 8) = pack8t(o);
  }
  block_chunk_reduce<K>(acc, chunks, C, sm);
  // bias gradient: warp sums first (only the ck == 0 threads hold non-zero values), then <= 8 adds per address
  if (threadIdx.x < K) sm[K * C + threadIdx.x] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float t = warp_sum(accb[k]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sm[K * C + k], t);
  }
  __syncthreads();
  // `slots` copies of the tables (the caller sums them): ~1200 blocks adding to K * C addresses serialise in L2
  const int slot = blockIdx.x % slots;
  float* dst = dws + (size_t(slot) * gridDim.y + n) * K * C;
  for (int c = threadIdx.x; c < K * C; c += blockDim.x) atomicAdd(dst + c, sm[c]);
  if (threadIdx.x < K) atomicAdd(db + slot * K + threadIdx.x, sm[K * C + threadIdx.x]);
}

// dst += src  (channels-last bf16 with channel strides)
__global__ void __launch_bounds__(256) add_inplace_kernel(__nv_bfloat16* dst, int ldd, const __nv_bfloat16* __restrict__ src,
                                                          int lds, long long nvox_total, int C) {
  const int chunks = C >> 3;
  const long long total = nvox_total * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ck = int(i % chunks);
    const long long v = i / chunks;
    float a[8], b[8];
    unpack8t(*reinterpret_cast<const uint4*>(dst + v * ldd + ck * 8), a);
    unpack8t(*reinterpret_cast<const uint4*>(src + v * lds + ck * 8), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    *reinterpret_cast<uint4*>(dst + v * ldd + ck * 8) = pack8t(a);
  }
}

static inline int chunk_grid(long long nvox, int chunks, int n, int per_thread = 4) {
  long long blocks = (nvox * chunks + 256 * per_thread - 1) / (256 * per_thread);
  long long cap = (long long)num_sms() * 8 / (n > 0 ? n : 1);
  if (cap < 1) cap = 1;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return int((blocks + chunks - 1) / chunks * chunks);  // gx * 256 is a multiple of `chunks`
}

}  // namespace b21

using namespace b21;
typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------ Dice
extern "C" int b21_dice_fwd(const float* logits, const float* target, double* sums, float* loss, float* coef, int n,
                            int k, long long nvox, int jaccard, float smooth_nr, float smooth_dr, float weight,
                            void* stream) {
  B21_CHECK_ARG(logits && target && sums && loss && coef, "dice_fwd: null pointer");
  B21_CHECK_ARG(n > 0 && k > 0 && k <= 16 && nvox > 0, "dice_fwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  B21_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 3 * k, st));
  dim3 grid(grid_cap((nvox + 1023) / 1024, 4), k);
  dice_fwd_kernel<<<grid, 256, 0, st>>>(logits, target, sums, n, k, nvox);
  dice_finalize_kernel<<<1, 32, 0, st>>>(sums, loss, coef, k, jaccard, smooth_nr, smooth_dr, weight);
  B21_LAUNCH_CHECK("dice_fwd_kernel");
  return B21_OK;
}

extern "C" int b21_ce_fwd(const float* logits, const float* target, double* scratch, float* loss, int n, int k,
                          long long nvox, float weight, void* stream) {
  B21_CHECK_ARG(logits && target && scratch && loss, "ce_fwd: null pointer");
  B21_CHECK_ARG(n > 0 && k > 1 && k <= kMaxClasses && nvox > 0, "ce_fwd: bad sizes (2..%d classes)", kMaxClasses);
  cudaStream_t st = (cudaStream_t)stream;
  B21_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), st));
  ce_fwd_kernel<<<grid_cap(((long long)n * nvox + 1023) / 1024, 4), 256, 0, st>>>(logits, target, scratch, n, k, nvox);
  ce_finalize_kernel<<<1, 32, 0, st>>>(scratch, loss, double(n) * double(nvox), weight);
  B21_LAUNCH_CHECK("ce_fwd_kernel");
  return B21_OK;
}

extern "C" int b21_dice_bwd(const float* logits, const float* target, const float* coef, const float* gout, float gscale,
                            float ce_weight, float* dlogits, int n, int k, long long nvox, void* stream) {
  B21_CHECK_ARG(logits && target && coef && dlogits, "dice_bwd: null pointer");
  B21_CHECK_ARG(ce_weight == 0.f || (k > 1 && k <= kMaxClasses), "dice_bwd: the cross-entropy term needs 2..%d classes", kMaxClasses);
  const long long total = (long long)n * k * nvox;
  // ce_weight is d(total loss)/d(ce loss) = lambda_ce; the mean over n * nvox voxels is applied here
  const float ce_w = ce_weight / float(double(n) * double(nvox));
  dice_bwd_kernel<<<grid_cap((total + 1023) / 1024), 256, 0, (cudaStream_t)stream>>>(logits, target, coef, gout, gscale,
                                                                                      ce_w, dlogits, n, k, nvox);
  B21_LAUNCH_CHECK("dice_bwd_kernel");
  return B21_OK;
}

// ------------------------------------------------------------------------------------------------ norm backward
extern "C" long long b21_norm_bwd_workspace_bytes(int n, int c) {
  if (n <= 0 || c <= 0) return 0;
  return (long long)(size_t(kRedSlots) * n * c * 3 * sizeof(double) + size_t(kColSlots) * c * sizeof(float) + 16 +
                     size_t(n) * c * 6 * sizeof(float));
}

extern "C" int b21_norm_bwd(const void* dy, int lddy, const void* z, int ldz, void* dz, int lddz, const double* stats,
                            const float* gamma, const float* beta, float* dgamma, float* dbeta, float* colsum,
                            const float* se_scale, const float* se_mean, const float* se_w1, const float* se_b1,
                            const float* se_w2, const float* se_b2, float* d_w1, float* d_b1, float* d_w2, float* d_b2,
                            int hidden, void* workspace, long long workspace_bytes, int mode, int n, long long nvox, int c,
                            float eps, void* stream) {
  B21_CHECK_ARG(dy && z && dz && stats && gamma && beta && dgamma && dbeta && workspace, "norm_bwd: null pointer");
  B21_CHECK_ARG(mode == 0 || mode == 1, "norm_bwd: mode must be 0 (GN+ReLU) or 1 (EvoNorm-S0)");
  B21_CHECK_ARG(c % 8 == 0 && lddy % 8 == 0 && ldz % 8 == 0 && lddz % 8 == 0, "norm_bwd: C/ld must be multiples of 8");
  B21_CHECK_ARG(!(mode == 0 && se_w1), "norm_bwd: squeeze-excite only follows EvoNorm");
  B21_CHECK_ARG(!se_w1 || (se_scale && se_mean && se_b1 && se_w2 && se_b2 && d_w1 && d_b1 && d_w2 && d_b2 && hidden > 0),
                "norm_bwd: incomplete squeeze-excite arguments");
  // workspace: red double[kRedSlots][n][c][3] | cs float[kColSlots][c] | counter (16 B) | coef float[n][c][4] |
  //            fa float[n][c] | fb float[n][c]        (the first three are zeroed by ONE memset per call)
  const size_t need = (size_t)b21_norm_bwd_workspace_bytes(n, c);
  B21_CHECK_ARG((size_t)workspace_bytes >= need, "norm_bwd: workspace too small (%lld < %zu)", workspace_bytes, need);
  B21_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "norm_bwd: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  double* red = reinterpret_cast<double*>(workspace);
  float* cs_ws = reinterpret_cast<float*>(red + size_t(kRedSlots) * n * c * 3);
  unsigned int* cs_count = reinterpret_cast<unsigned int*>(cs_ws + size_t(kColSlots) * c);
  float* coef = reinterpret_cast<float*>(cs_count + 4);
  float* fa = coef + size_t(n) * c * 4;
  float* fb = fa + size_t(n) * c;
  B21_CUDA(cudaMemsetAsync(red, 0, sizeof(double) * kRedSlots * n * c * 3 + sizeof(float) * kColSlots * c + 16, st));
  const int chunks = c / 8;
  dim3 grid(chunk_grid(nvox, chunks, n), n);
  if (mode == 0) {
    gn_fwd_affine_kernel<<<n, 256, 0, st>>>(stats, gamma, beta, fa, fb, n, c, nvox, eps);
    norm_bwd_reduce_kernel<0><<<grid, 256, sizeof(float) * 3 * c, st>>>((const bf16*)dy, lddy, (const bf16*)z, ldz, fa, fb, red, nvox, c);
  } else {
    norm_bwd_reduce_kernel<1><<<grid, 256, sizeof(float) * 3 * c, st>>>((const bf16*)dy, lddy, (const bf16*)z, ldz, fa, fb, red, nvox, c);
  }
  NormBwdArgs a;
  a.red = red; a.redw = red; a.stats = stats; a.gamma = gamma; a.beta = beta; a.fa = fa; a.fb = fb; a.coef = coef;
  a.dgamma = dgamma; a.dbeta = dbeta;
  a.se_scale = se_scale; a.se_mean = se_mean; a.se_w1 = se_w1; a.se_b1 = se_b1; a.se_w2 = se_w2; a.se_b2 = se_b2;
  a.d_w1 = d_w1; a.d_b1 = d_b1; a.d_w2 = d_w2; a.d_b2 = d_b2;
  a.N = n; a.C = c; a.hidden = se_w1 ? hidden : 0; a.nvox = nvox; a.eps = eps;
  const size_t smc = sizeof(float) * (32 + 2 * c + 2 * a.hidden + (a.hidden ? 512 * 4 : 0));
  if (mode == 0) norm_bwd_coeffs_kernel<0><<<1, 512, smc, st>>>(a);
  else norm_bwd_coeffs_kernel<1><<<1, 512, smc, st>>>(a);
  if (mode == 0)
    norm_bwd_apply_kernel<0><<<grid, 256, sizeof(float) * c, st>>>((const bf16*)dy, lddy, (const bf16*)z, ldz, (bf16*)dz, lddz, coef, fa, fb, colsum, cs_ws, cs_count, nvox, c);
  else
    norm_bwd_apply_kernel<1><<<grid, 256, sizeof(float) * c, st>>>((const bf16*)dy, lddy, (const bf16*)z, ldz, (bf16*)dz, lddz, coef, fa, fb, colsum, cs_ws, cs_count, nvox, c);
  B21_LAUNCH_CHECK("norm_bwd kernels");
  return B21_OK;
}

extern "C" int b21_pool_bwd(const void* y, int ldy, const void* dpool, int ldp, const void* add, int ldadd, void* dy,
                            int lddy, int mode, int n, int d, int h, int w, int c, void* stream) {
  B21_CHECK_ARG(y && dpool && dy, "pool_bwd: null pointer");
  B21_CHECK_ARG(mode == 1 || mode == 2, "pool_bwd: mode 1 (max) or 2 (max|avg)");
  B21_CHECK_ARG(c % 8 == 0 && d % 2 == 0 && h % 2 == 0 && w % 2 == 0, "pool_bwd: C multiple of 8, even dims");
  const long long items = (long long)n * (d / 2) * (h / 2) * (w / 2) * (c / 8);
  pool_bwd_kernel<<<grid_cap((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)y, ldy, (const bf16*)dpool, ldp, (const bf16*)add, ldadd, (bf16*)dy, lddy, mode, n, d, h, w, c);
  B21_LAUNCH_CHECK("pool_bwd_kernel");
  return B21_OK;
}

extern "C" int b21_upsample2x_bwd(const void* dy, int lddy, void* dx, int lddx, int n, int d, int h, int w, int c,
                                  void* stream) {
  B21_CHECK_ARG(dy && dx && c % 8 == 0, "upsample2x_bwd: bad args");
  const long long items = (long long)n * d * h * w * (c / 8);
  upsample2x_bwd_kernel<<<grid_cap((items + 127) / 128, 16), 128, 0, (cudaStream_t)stream>>>((const bf16*)dy, lddy, (bf16*)dx, lddx, n, d, h, w, c);
  B21_LAUNCH_CHECK("upsample2x_bwd_kernel");
  return B21_OK;
}

extern "C" int b21_upsample_f32_bwd(const float* dy, float* dx, int planes, int d, int h, int w, int s, void* stream) {
  B21_CHECK_ARG(dy && dx && planes > 0 && s >= 1 && s <= 8, "upsample_f32_bwd: bad args (factor 1..8)");
  const long long items = (long long)planes * d * h * w;
  upsample_f32_bwd_kernel<<<grid_cap((items + 127) / 128, 16), 128, 0, (cudaStream_t)stream>>>(dy, dx, planes, d, h, w, s);
  B21_LAUNCH_CHECK("upsample_f32_bwd_kernel");
  return B21_OK;
}

extern "C" int b21_head_conv_bwd(const void* x, int ldx, const float* scale, const float* w, const float* dl, void* dx,
                                 int lddx, int accumulate, float* dws, float* db, int slots, int n, long long nvox, int c,
                                 int k, void* stream) {
  B21_CHECK_ARG(x && w && dl && dx && dws && db, "head_conv_bwd: null pointer");
  B21_CHECK_ARG(k >= 1 && k <= 4 && c % 8 == 0 && c <= 1024, "head_conv_bwd: K 1..4, C multiple of 8, C <= 1024");
  B21_CHECK_ARG(slots >= 1, "head_conv_bwd: slots must be >= 1");
  const int chunks = c / 8;
  dim3 grid(chunk_grid(nvox, chunks, n, 2), n);
  const size_t smem = sizeof(float) * (k * c + k);
  cudaStream_t st = (cudaStream_t)stream;
#define B21_HEAD_BWD(KK) head_conv_bwd_kernel<KK><<<grid, 256, smem, st>>>((const bf16*)x, ldx, scale, w, dl, (bf16*)dx, lddx, accumulate, dws, db, slots, nvox, c)
  switch (k) {
    case 1: B21_HEAD_BWD(1); break;
    case 2: B21_HEAD_BWD(2); break;
    case 3: B21_HEAD_BWD(3); break;
    default: B21_HEAD_BWD(4); break;
  }
#undef B21_HEAD_BWD
  B21_LAUNCH_CHECK("head_conv_bwd_kernel");
  return B21_OK;
}

extern "C" int b21_add_inplace(void* dst, int ldd, const void* src, int lds, long long nvox_total, int c, void* stream) {
  B21_CHECK_ARG(dst && src && c % 8 == 0 && ldd % 8 == 0 && lds % 8 == 0, "add_inplace: bad args");
  const long long items = nvox_total * (c / 8);
  add_inplace_kernel<<<grid_cap((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>((bf16*)dst, ldd, (const bf16*)src, lds, nvox_total, c);
  B21_LAUNCH_CHECK("add_inplace_kernel");
  return B21_OK;
}

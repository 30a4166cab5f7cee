// elementwise.cu — HBM-bound kernels of the network body on channels-last (NDHWC) bf16 activations.
// Every kernel moves 16-byte vectors (8 channels), keeps consecutive lanes on consecutive addresses, holds
// per-channel parameters in registers/shared memory and sizes its grid as a multiple of the SM count.
//
//   norm_apply   GroupNorm(8,C)+ReLU (networks/factory.py:182, equiunet2020.py:60-61) or EvoNorm3D-S0
//                (networks/equiunet2021.py:48-52,95-105) from the (sum, sumsq) group statistics the conv epilogue
//                produced; optional per-channel output sums for the squeeze-excite mean.
//   se_gate      MONAI ResidualSELayer MLP -> per (n, c) scale 1 + sigmoid(...)   (equiunet2021.py:204-205)
//   scale_pool   x * scale written full-res and/or 2x2x2 max / [max, avg] pooled  (MaxPool3d equiunet2020.py:433,
//                MONAI MaxAvgPool equiunet2021.py:261)
//   upsample2x   trilinear x2, align_corners=True, into a channel slice           (equiunet2020.py:439)
//   head_conv    1x1 conv C -> K<=4 logits, NCDHW fp32 out                         (equiunet2020.py:441, 2021.py:271)
//   upsample_f32 trilinear xS (align_corners) on NCDHW fp32 (deep-supervision heads, equiunet2021.py:274-280)
#include "ptx.cuh"
#include "fold.cuh"
#include "host_common.h"
#include "reduce.cuh"
#include <stdlib.h>

namespace b21 {

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------------------------------------ norm apply
// grid = (gx, N) with gx * blockDim a multiple of C/8, so that each thread always serves the same 8 channels.
template <int MODE>  // 0 = GroupNorm+ReLU, 1 = EvoNorm-S0
__global__ void __launch_bounds__(256) norm_apply_kernel(const __nv_bfloat16* x, int ldx,  // x may alias y
                                                         __nv_bfloat16* y, int ldy,
                                                         const double* __restrict__ stats,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* chan_sum,
                                                         int chan_slots, int N, long long nvox, int C, float eps) {
  extern __shared__ float sm[];  // a[C], b[C], csum[C]
  float* sa = sm;
  float* sb = sm + C;
  float* ssum = sm + 2 * C;
  const int n = blockIdx.y;
  const int gsz = C / 8;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / gsz;
    double s = 0.0, q = 0.0;
    for (int slot = 0; slot < B21_STAT_SLOTS; ++slot) {
      const double* p = stats + ((size_t(slot) * N + n) * 8 + g) * 2;
      s += p[0];
      q += p[1];
    }
    const double cnt = double(nvox) * gsz;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    if (MODE == 0) {
      const float rstd = float(1.0 / sqrt(var + double(eps)));
      sa[c] = gamma[c] * rstd;
      sb[c] = beta[c] - float(mean) * gamma[c] * rstd;
    } else {
      const double varu = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;  // torch.var default: unbiased
      sa[c] = gamma[c] * float(1.0 / sqrt(varu + double(eps)));
      sb[c] = beta[c];
    }
    ssum[c] = 0.f;
  }
  __syncthreads();
  const int chunks = C >> 3;
  const long long total = nvox * chunks;
  const long long T = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ck = int(i % chunks);  // invariant: T % chunks == 0
  float a[8], b[8], accv[1][8];
  float (&acc)[8] = accv[0];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = sa[ck * 8 + j];
    b[j] = sb[ck * 8 + j];
    acc[j] = 0.f;
  }
  const __nv_bfloat16* xn = x + size_t(n) * nvox * ldx;
  __nv_bfloat16* yn = y + size_t(n) * nvox * ldy;
  constexpr int U = 4;
  // T is a multiple of `chunks`: the voxel index advances by T / chunks per step (no 64-bit division in the loop)
  const long long vstep = T / chunks;
  long long v0 = i / chunks;
  for (; i < total; i += T * U, v0 += vstep * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + u * T;
      if (iu < total) v[u] = *reinterpret_cast<const uint4*>(xn + (v0 + u * vstep) * ldx + ck * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + u * T;
      if (iu < total) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float r;
          if (MODE == 0) {
            r = fmaxf(fmaf(f[j], a[j], b[j]), 0.f);
          } else {
            const float sw = swishf(f[j]);  // x * sigmoid(x)
            r = fmaf(sw, a[j], b[j]);
          }
          f[j] = r;
        }
        const uint4 o = pack8(f);
        if (chan_sum) {  // sums of the ROUNDED outputs: what the next layer actually sees
          float g[8];
          unpack8(o, g);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += g[j];
        }
        *reinterpret_cast<uint4*>(yn + (v0 + u * vstep) * ldy + ck * 8) = o;
      }
    }
  }
  if (chan_sum) {
    block_chunk_reduce<1>(accv, chunks, C, ssum);
    // chan_slots copies of the table: ~1200 blocks adding to the same C addresses serialise in the L2 atomic unit
    float* cs = chan_sum + (size_t(blockIdx.x % chan_slots) * N + n) * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(cs + c, ssum[c]);
  }
}

// ------------------------------------------------------------------------------------------ SE gate
// one block per n: scale[n][c] = 1 + sigmoid(W2 relu(W1 m + b1) + b2), m = chan_sum * inv_count
__global__ void __launch_bounds__(1024) se_gate_kernel(const float* __restrict__ chan_sum, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, float* __restrict__ scale, int C,
                                                       int Hd, float inv_count) {
  extern __shared__ float sm[];  // m[C], hid[Hd]
  float* m = sm;
  float* hid = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) m[c] = chan_sum[size_t(n) * C + c] * inv_count;
  __syncthreads();
  // latency-bound (4 CTAs, two dependent mat-vecs): 32 warps, every lane keeps up to 12 independent loads in flight
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < Hd; j += nw) {
    const float* wr = w1 + size_t(j) * C;
    float s = 0.f;
    int c = lane;
    for (; c + 352 < C; c += 384) {
      float t[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) t[k] = __ldg(wr + c + 32 * k);
#pragma unroll
      for (int k = 0; k < 12; ++k) s = fmaf(t[k], m[c + 32 * k], s);
    }
    for (; c < C; c += 32) s = fmaf(__ldg(wr + c), m[c], s);
    s = warp_sum(s);
    if (lane == 0) hid[j] = fmaxf(s + b1[j], 0.f);
  }
  __syncthreads();
  for (int c = warp; c < C; c += nw) {
    const float* wr = w2 + size_t(c) * Hd;
    float s = 0.f;
    int j = lane;
    for (; j + 160 < Hd; j += 192) {
      float t[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) t[k] = __ldg(wr + j + 32 * k);
#pragma unroll
      for (int k = 0; k < 6; ++k) s = fmaf(t[k], hid[j + 32 * k], s);
    }
    for (; j < Hd; j += 32) s = fmaf(__ldg(wr + j), hid[j], s);
    s = warp_sum(s);
    if (lane == 0) scale[size_t(n) * C + c] = 1.f + 1.f / (1.f + expf(-(s + b2[c])));
  }
}

// ------------------------------------------------------------------------------------------ scale + pool
// One thread per (pooled voxel, 8-channel chunk): reads the 2x2x2 block, optionally writes it back scaled
// (full-res, may alias x) and writes max (mode 1) or [max | avg] (mode 2) at half resolution.
__global__ void __launch_bounds__(256) scale_pool_kernel(const __nv_bfloat16* x, int ldx,
                                                         const float* __restrict__ scale, __nv_bfloat16* full,
                                                         int ldfull, __nv_bfloat16* __restrict__ pooled, int ldpool,
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
    float sc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sc[j] = scale ? __ldg(scale + size_t(n) * C + ck * 8 + j) : 1.f;
    float mx[8], sm[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { mx[j] = -INFINITY; sm[j] = 0.f; }
    uint4 raw[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const size_t vox = ((size_t(n) * D + (2 * dp + (t >> 2))) * H + (2 * hp + ((t >> 1) & 1))) * W + (2 * wp + (t & 1));
      raw[t] = *reinterpret_cast<const uint4*>(x + vox * ldx + ck * 8);
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float f[8];
      unpack8(raw[t], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= sc[j];
      uint4 o = pack8(f);
      if (full) {
        const size_t vox = ((size_t(n) * D + (2 * dp + (t >> 2))) * H + (2 * hp + ((t >> 1) & 1))) * W + (2 * wp + (t & 1));
        *reinterpret_cast<uint4*>(full + vox * ldfull + ck * 8) = o;
      }
      unpack8(o, f);  // pool what downstream layers see (bf16-rounded)
#pragma unroll
      for (int j = 0; j < 8; ++j) { mx[j] = fmaxf(mx[j], f[j]); sm[j] += f[j]; }
    }
    if (pooled) {
      const size_t pv = ((size_t(n) * Dp + dp) * Hp + hp) * Wp + wp;
      *reinterpret_cast<uint4*>(pooled + pv * ldpool + ck * 8) = pack8(mx);
      if (mode == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[j] *= 0.125f;
        *reinterpret_cast<uint4*>(pooled + pv * ldpool + C + ck * 8) = pack8(sm);
      }
    }
  }
}

// scale only (no pooling): y = x * scale[n][c]
__global__ void __launch_bounds__(256) scale_kernel(const __nv_bfloat16* x, int ldx, const float* __restrict__ scale,
                                                    __nv_bfloat16* y, int ldy, int N, long long nvox, int C) {
  const int chunks = C >> 3;
  const long long total = (long long)N * nvox * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ck = int(i % chunks);
    const long long v = i / chunks;
    const int n = int(v / nvox);
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(x + v * ldx + ck * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= __ldg(scale + size_t(n) * C + ck * 8 + j);
    *reinterpret_cast<uint4*>(y + v * ldy + ck * 8) = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------ trilinear x2
__device__ __forceinline__ void lerp_setup(int o, int I, int O, int& i0, int& i1, float& l1) {
  // torch upsample, align_corners=True: src = o * (I-1)/(O-1)
  const float sc = (O > 1) ? float(I - 1) / float(O - 1) : 0.f;
  const float src = sc * float(o);
  i0 = int(src);
  if (i0 > I - 1) i0 = I - 1;
  i1 = i0 + (i0 < I - 1 ? 1 : 0);
  l1 = src - float(i0);
}

__global__ void __launch_bounds__(256) upsample2x_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                         __nv_bfloat16* __restrict__ y, int ldy, int N, int D, int H,
                                                         int W, int C) {
  const int chunks = C >> 3;
  const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)N * Do * Ho * Wo * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ck = int(i % chunks);
    long long v = i / chunks;
    const int wo = int(v % Wo); v /= Wo;
    const int ho = int(v % Ho); v /= Ho;
    const int d_o = int(v % Do);
    const int n = int(v / Do);
    int d0, d1, h0, h1, w0, w1;
    float ld, lh, lw;
    lerp_setup(d_o, D, Do, d0, d1, ld);
    lerp_setup(ho, H, Ho, h0, h1, lh);
    lerp_setup(wo, W, Wo, w0, w1, lw);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int dd = (t & 4) ? d1 : d0, hh = (t & 2) ? h1 : h0, ww = (t & 1) ? w1 : w0;
      const float wt = ((t & 4) ? ld : 1.f - ld) * ((t & 2) ? lh : 1.f - lh) * ((t & 1) ? lw : 1.f - lw);
      const size_t vox = ((size_t(n) * D + dd) * H + hh) * W + ww;
      float f[8];
      unpack8(ldg16(x + vox * ldx + ck * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt, f[j], acc[j]);
    }
    const size_t ov = ((size_t(n) * Do + d_o) * Ho + ho) * Wo + wo;
    *reinterpret_cast<uint4*>(y + ov * ldy + ck * 8) = pack8(acc);
  }
}

// Faster x2 variant: one thread per (INPUT voxel, 8-channel chunk) produces the 2x2x2 output block.  The two outputs
// of an axis interpolate between the inputs {i-1, i, i+1} only (align_corners=True: src = o (I-1)/(2I-1)), so the
// block needs 2 x 9 loads per output plane instead of 8 per output voxel, and the interpolation is done separably
// (d, then h, then w).  Weights come from the same lerp_setup as the reference formulation.
__device__ __forceinline__ void axis_slots(int i, int I, float w0[3], float w1[3]) {
  // weights of outputs 2i (w0) and 2i+1 (w1) over the input slots (i-1, i, i+1)
#pragma unroll
  for (int k = 0; k < 3; ++k) w0[k] = w1[k] = 0.f;
#pragma unroll
  for (int par = 0; par < 2; ++par) {
    int i0, i1;
    float l;
    lerp_setup(2 * i + par, I, 2 * I, i0, i1, l);
    float* wv = par ? w1 : w0;
    const int s0 = i0 - i + 1, s1 = i1 - i + 1;
#pragma unroll
    for (int k = 0; k < 3; ++k) wv[k] += (s0 == k ? 1.f - l : 0.f) + (s1 == k ? l : 0.f);
  }
}

__global__ void __launch_bounds__(128, 4) upsample2x_block_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                               __nv_bfloat16* __restrict__ y, int ldy, int N, int D,
                                                               int H, int W, int C) {
  const int chunks = C >> 3;
  const int Ho = 2 * H, Wo = 2 * W;
  // grid: x over (iw, chunk), y = ih, z = (n, id, od): no 64-bit divisions in the index math
  {
    const int tx = blockIdx.x * blockDim.x + threadIdx.x;
    if (tx >= W * chunks) return;
    const int ck = tx % chunks;
    const int iw = tx / chunks;
    const int ih = blockIdx.y;
    const int od = blockIdx.z & 1;  // one thread per output d plane of the block (register pressure)
    const int id = (blockIdx.z >> 1) % D;
    const int n = (blockIdx.z >> 1) / D;
    float wd[2][3], wh[2][3], ww[2][3];
    axis_slots(id, D, wd[0], wd[1]);
    axis_slots(ih, H, wh[0], wh[1]);
    axis_slots(iw, W, ww[0], ww[1]);
    int hs[3], wsl[3], ds[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ds[k] = min(max(id + k - 1, 0), D - 1);
      hs[k] = min(max(ih + k - 1, 0), H - 1);
      wsl[k] = min(max(iw + k - 1, 0), W - 1);
    }
    const __nv_bfloat16* xn = x + size_t(n) * D * H * W * ldx + ck * 8;
    {
      // the two d planes this output plane interpolates: slots (0,1) for od = 0, (1,2) for od = 1; a weight that
      // rounding put on the third slot (at most 1 ulp of the coordinate) is added to its neighbour
      float wdo[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) wdo[k] = od ? wd[1][k] : wd[0][k];  // selects, not dynamic register indexing
      const int dsa = od == 0 ? ds[0] : ds[1], dsb = od == 0 ? ds[1] : ds[2];
      const float wa = od == 0 ? wdo[0] : wdo[1] + wdo[0], wb = od == 0 ? wdo[1] + wdo[2] : wdo[2];
      float t[3][3][8];
#pragma unroll
      for (int b = 0; b < 3; ++b) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const size_t off = (size_t(hs[b]) * W + wsl[c]) * ldx;
          float fa[8], fb[8];
          unpack8(ldg16(xn + size_t(dsa) * H * W * ldx + off), fa);
          unpack8(ldg16(xn + size_t(dsb) * H * W * ldx + off), fb);
#pragma unroll
          for (int j = 0; j < 8; ++j) t[b][c][j] = fmaf(wb, fb[j], wa * fa[j]);
        }
      }
#pragma unroll
      for (int oh = 0; oh < 2; ++oh) {
        float u[3][8];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            u[c][j] = fmaf(wh[oh][2], t[2][c][j], fmaf(wh[oh][1], t[1][c][j], wh[oh][0] * t[0][c][j]));
#pragma unroll
        for (int ow = 0; ow < 2; ++ow) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(ww[ow][2], u[2][j], fmaf(ww[ow][1], u[1][j], ww[ow][0] * u[0][j]));
          const size_t ov = ((size_t(n) * 2 * D + (2 * id + od)) * Ho + (2 * ih + oh)) * Wo + (2 * iw + ow);
          *reinterpret_cast<uint4*>(y + ov * ldy + ck * 8) = pack8(o);
        }
      }
    }
  }
}

// Row-marching x2 variant (the one the networks run): a block owns one output plane (n, od) and a segment of its rows;
// thread = (input column iw, 8-channel chunk).  Separable interpolation with every intermediate computed ONCE:
//   d stage  T[ih]      = lerp_d(x[d0][ih], x[d1][ih])      registers, refreshed as the march along h advances (2 loads)
//   h stage  U[oh]      = lerp_h(T[h0], T[h1])              one lerp per output row, exchanged through shared memory
//   w stage  y[2iw + p] = lerp_w(U[w0], U[w1])              two outputs per thread
// = 1.75 lerps per output vector (the 2x2x2-block kernel above: 4.75) and ~4 instructions per output element.
__global__ void __launch_bounds__(1024) upsample2x_march_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                                __nv_bfloat16* __restrict__ y, int ldy, int D, int H,
                                                                int W, int C, int rows_per_block) {
  extern __shared__ float4 urow[];  // [2 buffers][W * chunks][2 float4]
  const int chunks = C >> 3;
  const int cols = W * chunks;
  const int t = threadIdx.x;
  const bool live = t < cols;
  const int iw = live ? t / chunks : 0;
  const int ck = live ? t - iw * chunks : 0;
  const int od = blockIdx.y, n = blockIdx.z;
  const int Ho = 2 * H, Wo = 2 * W;
  int d0, d1;
  float ld;
  lerp_setup(od, D, 2 * D, d0, d1, ld);
  // the two outputs of this thread along w
  int wa0, wa1, wb0, wb1;
  float la, lb;
  lerp_setup(2 * iw, W, Wo, wa0, wa1, la);
  lerp_setup(2 * iw + 1, W, Wo, wb0, wb1, lb);
  const __nv_bfloat16* x0 = x + ((size_t(n) * D + d0) * H * W + iw) * ldx + ck * 8;
  const __nv_bfloat16* x1 = x + ((size_t(n) * D + d1) * H * W + iw) * ldx + ck * 8;
  const size_t xrow = size_t(W) * ldx;
  const int oh_begin = blockIdx.x * rows_per_block;
  const int oh_end = min(oh_begin + rows_per_block, Ho);
  float tp[8], tc[8];  // T rows h0 and h1 of the current output row
  int have0 = -1, have1 = -1;
  auto load_t = [&](int ih, float* dst) {
    float a[8], b[8];
    unpack8(ldg16(x0 + size_t(ih) * xrow), a);
    unpack8(ldg16(x1 + size_t(ih) * xrow), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = fmaf(ld, b[j] - a[j], a[j]);
  };
  for (int oh = oh_begin; oh < oh_end; ++oh) {
    int h0, h1;
    float lh;
    lerp_setup(oh, H, Ho, h0, h1, lh);
    float4* buf = urow + size_t(oh & 1) * cols * 2;
    if (live) {
      if (have0 != h0) {
        if (have1 == h0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) tp[j] = tc[j];
        } else {
          load_t(h0, tp);
        }
        have0 = h0;
      }
      if (have1 != h1) {
        if (h1 == h0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) tc[j] = tp[j];
        } else {
          load_t(h1, tc);
        }
        have1 = h1;
      }
      float u[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = fmaf(lh, tc[j] - tp[j], tp[j]);
      buf[t * 2] = make_float4(u[0], u[1], u[2], u[3]);
      buf[t * 2 + 1] = make_float4(u[4], u[5], u[6], u[7]);
    }
    __syncthreads();  // (double-buffered rows: one barrier per output row)
    if (live) {
      __nv_bfloat16* yo = y + (((size_t(n) * 2 * D + od) * Ho + oh) * Wo + 2 * iw) * ldy + ck * 8;
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int w0 = par ? wb0 : wa0, w1 = par ? wb1 : wa1;
        const float lw = par ? lb : la;
        const float4 p0 = buf[(w0 * chunks + ck) * 2], p1 = buf[(w0 * chunks + ck) * 2 + 1];
        const float4 q0 = buf[(w1 * chunks + ck) * 2], q1 = buf[(w1 * chunks + ck) * 2 + 1];
        float o[8];
        o[0] = fmaf(lw, q0.x - p0.x, p0.x); o[1] = fmaf(lw, q0.y - p0.y, p0.y);
        o[2] = fmaf(lw, q0.z - p0.z, p0.z); o[3] = fmaf(lw, q0.w - p0.w, p0.w);
        o[4] = fmaf(lw, q1.x - p1.x, p1.x); o[5] = fmaf(lw, q1.y - p1.y, p1.y);
        o[6] = fmaf(lw, q1.z - p1.z, p1.z); o[7] = fmaf(lw, q1.w - p1.w, p1.w);
        *reinterpret_cast<uint4*>(yo + size_t(par) * ldy) = pack8(o);
      }
    }
  }
}

// trilinear xS on NCDHW fp32 planes (deep-supervision heads): one thread per output element
__global__ void __launch_bounds__(256) upsample_f32_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           int planes, int D, int H, int W, int S) {
  const int Do = S * D, Ho = S * H, Wo = S * W;
  const long long total = (long long)planes * Do * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long v = i;
    const int wo = int(v % Wo); v /= Wo;
    const int ho = int(v % Ho); v /= Ho;
    const int d_o = int(v % Do);
    const int pl = int(v / Do);
    int d0, d1, h0, h1, w0, w1;
    float ld, lh, lw;
    lerp_setup(d_o, D, Do, d0, d1, ld);
    lerp_setup(ho, H, Ho, h0, h1, lh);
    lerp_setup(wo, W, Wo, w0, w1, lw);
    const float* xp = x + size_t(pl) * D * H * W;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int dd = (t & 4) ? d1 : d0, hh = (t & 2) ? h1 : h0, ww = (t & 1) ? w1 : w0;
      const float wt = ((t & 4) ? ld : 1.f - ld) * ((t & 2) ? lh : 1.f - lh) * ((t & 1) ? lw : 1.f - lw);
      acc = fmaf(wt, __ldg(xp + (size_t(dd) * H + hh) * W + ww), acc);
    }
    y[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------ 1x1 head conv
// logits[n][k][vox] = b[k] + sum_c w[k][c] * (scale[n][c] * x[n][vox][c] + offset[n][c]);  K <= 4, C <= 1024
// (scale / offset = the folded affine of the input: SE gate, or EvoNorm (A, B) of the folded inference path)
template <int K>
__global__ void __launch_bounds__(256) head_conv_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                        const float* __restrict__ scale,
                                                        const float* __restrict__ offset, int ldso,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        float* __restrict__ out, int N, long long nvox, int C) {
  extern __shared__ float sw[];  // [K][C] (pre-scaled per n), then [K] biases
  float* sb = sw + K * C;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < K * C; i += blockDim.x)
    sw[i] = w[i] * (scale ? scale[size_t(n) * ldso + (i % C)] : 1.f);
  if (threadIdx.x < K) {
    float s = b ? b[threadIdx.x] : 0.f;
    if (offset)
      for (int c = 0; c < C; ++c) s = fmaf(w[threadIdx.x * C + c], offset[size_t(n) * ldso + c], s);
    sb[threadIdx.x] = s;
  }
  __syncthreads();
  const __nv_bfloat16* xn = x + size_t(n) * nvox * ldx;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvox;
       v += (long long)gridDim.x * blockDim.x) {
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = sb[k];
    const __nv_bfloat16* xv = xn + v * ldx;
    for (int c = 0; c < C; c += 8) {
      float f[8];
      unpack8(ldg16(xv + c), f);
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k] = fmaf(f[j], sw[k * C + c + j], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) out[(size_t(n) * K + k) * nvox + v] = acc[k];
  }
}

static inline int grid_for(long long work_items, int threads, int multiple_of = 1) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  blocks = (blocks + multiple_of - 1) / multiple_of * multiple_of;
  return int(blocks);
}

}  // namespace b21

using namespace b21;
typedef __nv_bfloat16 bf16;

extern "C" int b21_norm_apply(const void* x, int ldx, void* y, int ldy, const double* stats, const float* gamma,
                              const float* beta, float* chan_sum, int chan_slots, int mode, int n, long long nvox, int c,
                              float eps, void* stream) {
  B21_CHECK_ARG(x && y && stats && gamma && beta, "norm_apply: null pointer");
  B21_CHECK_ARG(!chan_sum || chan_slots >= 1, "norm_apply: chan_slots must be >= 1");
  B21_CHECK_ARG(mode == 0 || mode == 1, "norm_apply: mode must be 0 (GN+ReLU) or 1 (EvoNorm-S0)");
  B21_CHECK_ARG(c % 8 == 0 && (c / 8) >= 1 && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= c && ldy >= c, "norm_apply: bad C/ld");
  B21_CHECK_ARG(n > 0 && nvox > 0, "norm_apply: empty tensor");
  const int chunks = c / 8;
  long long blocks = (nvox * chunks + 256 * 4 - 1) / (256 * 4);
  const long long cap = (long long)num_sms() * 8 / n > 0 ? (long long)num_sms() * 8 / n : 1;
  if (blocks > cap) blocks = cap;
  const int gx = int((blocks + chunks - 1) / chunks * chunks);  // gx*256 must be a multiple of C/8
  const size_t smem = sizeof(float) * 3 * c;
  dim3 grid(gx, n);
  if (mode == 0)
    norm_apply_kernel<0><<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)x, ldx, (bf16*)y, ldy, stats, gamma, beta, chan_sum, chan_slots, n, nvox, c, eps);
  else
    norm_apply_kernel<1><<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)x, ldx, (bf16*)y, ldy, stats, gamma, beta, chan_sum, chan_slots, n, nvox, c, eps);
  B21_LAUNCH_CHECK("norm_apply_kernel");
  return B21_OK;
}

extern "C" int b21_se_gate(const float* chan_sum, const float* w1, const float* b1, const float* w2, const float* b2,
                           float* scale, int n, int c, int hidden, float inv_count, void* stream) {
  B21_CHECK_ARG(chan_sum && w1 && b1 && w2 && b2 && scale, "se_gate: null pointer");
  B21_CHECK_ARG(n > 0 && c > 0 && hidden > 0, "se_gate: bad sizes");
  se_gate_kernel<<<n, 1024, sizeof(float) * (c + hidden), (cudaStream_t)stream>>>(chan_sum, w1, b1, w2, b2, scale, c, hidden, inv_count);
  B21_LAUNCH_CHECK("se_gate_kernel");
  return B21_OK;
}

extern "C" int b21_scale_pool(const void* x, int ldx, const float* scale, void* full, int ldfull, void* pooled,
                              int ldpool, int mode, int n, int d, int h, int w, int c, void* stream) {
  B21_CHECK_ARG(x, "scale_pool: null input");
  B21_CHECK_ARG(c % 8 == 0 && ldx % 8 == 0, "scale_pool: C/ld must be multiples of 8");
  B21_CHECK_ARG(mode >= 0 && mode <= 2, "scale_pool: mode 0 (scale only), 1 (max), 2 (max|avg)");
  if (mode == 0) {
    B21_CHECK_ARG(scale && full, "scale_pool: mode 0 needs scale and full");
    const long long nvox = (long long)d * h * w;
    scale_kernel<<<grid_for((long long)n * nvox * (c / 8), 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, scale, (bf16*)full, ldfull, n, nvox, c);
    B21_LAUNCH_CHECK("scale_kernel");
    return B21_OK;
  }
  B21_CHECK_ARG(pooled && d % 2 == 0 && h % 2 == 0 && w % 2 == 0, "scale_pool: pooling needs even dims and an output");
  B21_CHECK_ARG(ldpool % 8 == 0 && ldpool >= (mode == 2 ? 2 * c : c), "scale_pool: pooled ld too small");
  const long long items = (long long)n * (d / 2) * (h / 2) * (w / 2) * (c / 8);
  scale_pool_kernel<<<grid_for(items, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, scale, (bf16*)full, ldfull, (bf16*)pooled, ldpool, mode, n, d, h, w, c);
  B21_LAUNCH_CHECK("scale_pool_kernel");
  return B21_OK;
}

extern "C" int b21_upsample2x(const void* x, int ldx, void* y, int ldy, int n, int d, int h, int w, int c,
                              void* stream) {
  B21_CHECK_ARG(x && y && c % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "upsample2x: bad args");
  const int cols = w * (c / 8);
  static int use_march = -1;
  if (use_march < 0) {
    const char* e = getenv("B21_UPSAMPLE_MARCH");
    use_march = e ? atoi(e) : 1;
  }
  if (use_march && d >= 2 && h >= 2 && w >= 2 && cols <= 1024 && 2 * d <= 65535 && n <= 65535) {
    const int threads = (cols + 31) / 32 * 32;
    // rows per block: enough blocks to fill the GPU a few times over, long enough marches to amortise the d stage
    int segs = 1;
    while ((long long)n * 2 * d * segs < 6LL * num_sms() && (2 * h) / (segs * 2) >= 8) segs *= 2;
    const int rows = (2 * h + segs - 1) / segs;
    dim3 grid((2 * h + rows - 1) / rows, 2 * d, n);
    const size_t smem = size_t(2) * cols * 2 * sizeof(float4);
    static bool attr_set = false;
    if (!attr_set) {
      B21_CUDA(cudaFuncSetAttribute(upsample2x_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      attr_set = true;
    }
    upsample2x_march_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>((const bf16*)x, ldx, (bf16*)y, ldy, d, h, w, c, rows);
    B21_LAUNCH_CHECK("upsample2x_march_kernel");
    return B21_OK;
  }
  if (d >= 2 && h >= 2 && w >= 2) {
    B21_CHECK_ARG(h <= 65535 && (long long)n * d * 2 <= 65535, "upsample2x: volume too large for the launch grid");
    dim3 grid((w * (c / 8) + 127) / 128, h, n * d * 2);
    upsample2x_block_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (bf16*)y, ldy, n, d, h, w, c);
    B21_LAUNCH_CHECK("upsample2x_block_kernel");
    return B21_OK;
  }
  const long long items = (long long)n * d * h * w * 8 * (c / 8);
  upsample2x_kernel<<<grid_for(items, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (bf16*)y, ldy, n, d, h, w, c);
  B21_LAUNCH_CHECK("upsample2x_kernel");
  return B21_OK;
}

extern "C" int b21_upsample_f32(const float* x, float* y, int planes, int d, int h, int w, int s, void* stream) {
  B21_CHECK_ARG(x && y && planes > 0 && s >= 1, "upsample_f32: bad args");
  const long long items = (long long)planes * d * h * w * s * s * s;
  upsample_f32_kernel<<<grid_for(items, 256), 256, 0, (cudaStream_t)stream>>>(x, y, planes, d, h, w, s);
  B21_LAUNCH_CHECK("upsample_f32_kernel");
  return B21_OK;
}

extern "C" int b21_head_conv(const void* x, int ldx, const float* scale, const float* offset, int ldso, const float* w,
                             const float* b, float* out,
                             int n, long long nvox, int c, int k, void* stream) {
  B21_CHECK_ARG(x && w && out, "head_conv: null pointer");
  B21_CHECK_ARG(k >= 1 && k <= 4 && c % 8 == 0 && c <= 1024, "head_conv: K must be 1..4 and C a multiple of 8");
  dim3 grid(grid_for(nvox, 256), n);
  const size_t smem = sizeof(float) * (k * c + k);
  cudaStream_t st = (cudaStream_t)stream;
  switch (k) {
    case 1: head_conv_kernel<1><<<grid, 256, smem, st>>>((const bf16*)x, ldx, scale, offset, ldso, w, b, out, n, nvox, c); break;
    case 2: head_conv_kernel<2><<<grid, 256, smem, st>>>((const bf16*)x, ldx, scale, offset, ldso, w, b, out, n, nvox, c); break;
    case 3: head_conv_kernel<3><<<grid, 256, smem, st>>>((const bf16*)x, ldx, scale, offset, ldso, w, b, out, n, nvox, c); break;
    default: head_conv_kernel<4><<<grid, 256, smem, st>>>((const bf16*)x, ldx, scale, offset, ldso, w, b, out, n, nvox, c); break;
  }
  B21_LAUNCH_CHECK("head_conv_kernel");
  return B21_OK;
}

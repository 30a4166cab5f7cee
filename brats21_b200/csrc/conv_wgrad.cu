// conv_wgrad.cu — weight gradient of conv3d (k = 1 or 3, any dilation, stride 1, "same" padding) on sm_100a.
//
//   dW[co][ci][tap] = sum over voxels v of  dz[v][co] * x[v + off(tap)][ci]
//
// GEMM view per tap: M = output channels (128-row tile), N = input channels (tile <= 256), K = voxels.  Both
// operands are read straight out of the NDHWC activations: a TMA box {64 channels, tw, th, td} lands in shared
// memory as [128 voxels][128 B] with the 128B swizzle, which IS the canonical "MN-major" UMMA operand layout
// (channels contiguous, K = voxel rows 128 B apart, 8-row groups 1024 B apart, next 64 channels LBO bytes away), so
// no transposed copy of the activations is ever made.  The zero padding of the convolution is the TMA out-of-bounds
// fill of the shifted x box.  Each CTA owns (M tile, N tile, a group of taps that fits TMEM: taps x N <= 512 columns)
// and a contiguous range of voxel tiles (split-K across the grid); accumulators stay in TMEM for the whole range and
// are flushed once with fp32 atomics into the torch-layout gradient [cout][cin][taps].
// Warp roles: warp 0 TMA producer (dz tile once per voxel tile, x tile once per tap), warp 1 MMA issuer,
// warps 2..5 epilogue.  Backward of networks/equiunet2020.py:19-41, networks/equiunet2021.py:165-172,192-222.
#include "ptx.cuh"
#include "host_common.h"
#include <stdlib.h>

namespace b21 {

constexpr int kWgThreads = 192;
constexpr int kWgChunk = 128 * 128;  // one TMA box: 128 voxels x 64 channels bf16

struct WgradParams {
  float* dw;
  int N, D, H, W, Cin, Cout, taps, dil;
  int tilesD, tilesH, tilesW, ntiles_total;
  int ltw, lth;
  int mtiles, ntiles, BN, nchunks;  // N tile (multiple of 16) and its 64-channel chunks
  int G, groups;                    // taps per CTA, tap groups
  int tiles_per_split;
  int xstages;
};

__device__ __forceinline__ uint32_t umma_idesc_bf16_mn(int M, int N) {
  // as umma_idesc_bf16, with A and B both MN-major (bits 15, 16)
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmZ, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t zfull[2], zempty[2], xfull[4], xempty[4], acc_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t z_addr = smem_u32(smem);                       // 2 stages x 2 chunks
  const uint32_t x_addr = z_addr + 2u * 2u * kWgChunk;          // xstages x nchunks chunks
  const uint32_t xstage_bytes = uint32_t(p.nchunks) * kWgChunk;

  int u = blockIdx.x;
  const int grp = u % p.groups; u /= p.groups;
  const int nt = u % p.ntiles;
  const int mt = u / p.ntiles;
  const int tap0 = grp * p.G;
  const int ntap = p.taps - tap0 < p.G ? p.taps - tap0 : p.G;
  const int t_begin = blockIdx.y * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split;
  t_end = t_end > p.ntiles_total ? p.ntiles_total : t_end;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&zfull[s], 1); mbar_init(&zempty[s], 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xempty[s], 1); }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmZ);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int tw = 1 << p.ltw, th = 1 << p.lth, td = 128 >> (p.ltw + p.lth);
  const uint32_t zf0 = smem_u32(&zfull[0]), ze0 = smem_u32(&zempty[0]), xf0 = smem_u32(&xfull[0]), xe0 = smem_u32(&xempty[0]);

  if (warp == 0) {
    if (elect_one()) {
      int zs = 0, xs = 0;
      uint32_t zph = 0, xph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        int q = t;
        const int wt = q % p.tilesW; q /= p.tilesW;
        const int ht = q % p.tilesH; q /= p.tilesH;
        const int dt = q % p.tilesD;
        const int n = q / p.tilesD;
        const int w0 = wt * tw, h0 = ht * th, d0 = dt * td;
        mbar_wait_sleep_a(ze0 + 8u * zs, zph ^ 1);
        mbar_expect_tx_a(zf0 + 8u * zs, 2u * kWgChunk);
        for (int c = 0; c < 2; ++c)
          tma_load_5d_a(z_addr + uint32_t(zs * 2 + c) * kWgChunk, &tmZ, zf0 + 8u * zs, mt * 128 + c * 64, w0, h0, d0, n);
        if (++zs == 2) { zs = 0; zph ^= 1; }
        for (int j = 0; j < ntap; ++j) {
          const int tap = tap0 + j;
          int kd = 0, kh = 0, kw = 0;
          if (p.taps == 27) { kd = tap / 9 - 1; kh = (tap / 3) % 3 - 1; kw = tap % 3 - 1; }
          mbar_wait_sleep_a(xe0 + 8u * xs, xph ^ 1);
          mbar_expect_tx_a(xf0 + 8u * xs, xstage_bytes);
          for (int c = 0; c < p.nchunks; ++c)
            tma_load_5d_a(x_addr + uint32_t(xs) * xstage_bytes + uint32_t(c) * kWgChunk, &tmX, xf0 + 8u * xs,
                          nt * p.BN + c * 64, w0 + kw * p.dil, h0 + kh * p.dil, d0 + kd * p.dil, n);
          if (++xs == p.xstages) { xs = 0; xph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // MN-major, 128B swizzle: SBO = 1024 (next 8 voxel rows), LBO = one chunk (next 64 channels)
      const uint64_t dsc = umma_smem_desc(0, kWgChunk, 1024, kLayoutSw128);
      const uint32_t idesc = umma_idesc_bf16_mn(128, p.BN);
      int zs = 0, xs = 0;
      uint32_t zph = 0, xph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait_a(zf0 + 8u * zs, zph);
        const uint64_t ad0 = dsc + uint64_t((z_addr + uint32_t(zs * 2) * kWgChunk) >> 4);
        for (int j = 0; j < ntap; ++j) {
          mbar_wait_a(xf0 + 8u * xs, xph);
          tc_fence_after();
          const uint64_t bd0 = dsc + uint64_t((x_addr + uint32_t(xs) * xstage_bytes) >> 4);
          const uint32_t dcol = tmem_base + uint32_t(j * p.BN);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // K = 16 voxels = two 8-row groups = 2048 B
            umma_bf16(dcol, ad0 + uint64_t(k * 128), bd0 + uint64_t(k * 128), idesc, (t > t_begin || k > 0) ? 1u : 0u);
          umma_commit_a(xe0 + 8u * xs);
          if (++xs == p.xstages) { xs = 0; xph ^= 1; }
        }
        umma_commit_a(ze0 + 8u * zs);
        if (++zs == 2) { zs = 0; zph ^= 1; }
      }
      umma_commit(&acc_bar);
    }
  } else {
    if (t_end > t_begin) {
      const int quad = warp & 3;
      const int co = mt * 128 + quad * 32 + lane;
      mbar_wait_sleep(&acc_bar, 0);
      tc_fence_after();
      for (int j = 0; j < ntap; ++j) {
        const int tap = tap0 + j;
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
          float v[16];
          tmem_ld16(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(j * p.BN + c0), v);
          tmem_ld_wait();
          if (co < p.Cout) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int ci = nt * p.BN + c0 + i;
              if (ci < p.Cin) atomicAdd(p.dw + (size_t(co) * p.Cin + ci) * p.taps + tap, v[i]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static inline int wg_floor_log2(int v) {
  int l = 0;
  while ((2 << l) <= v) ++l;
  return l;
}

}  // namespace b21

using namespace b21;

extern "C" int b21_conv3d_wgrad(const void* x, int ldx, const void* dz, int lddz, float* dw, int n, int d, int h, int w,
                                int cin, int cout, int taps, int dil, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B21_CHECK_ARG(x && dz && dw, "conv3d_wgrad: null pointer");
  B21_CHECK_ARG(taps == 1 || taps == 27, "conv3d_wgrad: taps must be 1 or 27 (got %d)", taps);
  B21_CHECK_ARG(n > 0 && d > 0 && h > 0 && w > 0, "conv3d_wgrad: bad shape");
  B21_CHECK_ARG(cin > 0 && cin % 8 == 0 && ldx >= cin && ldx % 8 == 0, "conv3d_wgrad: cin %d / ldx %d must be multiples of 8", cin, ldx);
  B21_CHECK_ARG(cout > 0 && cout % 8 == 0 && lddz >= cout && lddz % 8 == 0, "conv3d_wgrad: cout %d / lddz %d must be multiples of 8", cout, lddz);
  B21_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0,
                "conv3d_wgrad: pointers must be 16-byte aligned");
  WgradParams p;
  p.dw = dw;
  p.N = n; p.D = d; p.H = h; p.W = w; p.Cin = cin; p.Cout = cout; p.taps = taps; p.dil = dil;
  p.ltw = wg_floor_log2(w) < 3 ? wg_floor_log2(w) : 3;
  p.lth = wg_floor_log2(h) < 3 ? wg_floor_log2(h) : 3;
  const int tw = 1 << p.ltw, th = 1 << p.lth, td = 128 >> (p.ltw + p.lth);
  p.tilesW = (w + tw - 1) / tw;
  p.tilesH = (h + th - 1) / th;
  p.tilesD = (d + td - 1) / td;
  p.ntiles_total = n * p.tilesD * p.tilesH * p.tilesW;
  p.mtiles = (cout + 127) / 128;
  p.ntiles = (cin + 255) / 256;
  p.BN = ((cin + p.ntiles - 1) / p.ntiles + 15) / 16 * 16;
  p.nchunks = (p.BN + 63) / 64;
  p.G = 512 / p.BN;
  if (p.G > taps) p.G = taps;
  if (p.G > 9) p.G = 9;
  p.groups = (taps + p.G - 1) / p.G;
  const int units = p.mtiles * p.ntiles * p.groups;
  // split-K over voxel tiles.  Every CTA ends with taps x BN x 128 fp32 atomics, so the split factor aims at ONE CTA per
  // SM (measured at 16^3: 384x384 0.197 -> 0.105 ms, 96x384 0.170 -> 0.057 ms against four waves) and at two only when
  // a CTA would otherwise march over more than 64 voxel tiles.  B21_WG_WAVES overrides.
  static int waves_env = -1;
  if (waves_env < 0) {
    const char* e = getenv("B21_WG_WAVES");
    waves_env = e ? atoi(e) : 0;
  }
  int splits = (num_sms() + units - 1) / units;
  if (waves_env > 0) splits = (waves_env * num_sms() + units - 1) / units;
  else if (p.ntiles_total / splits > 64) splits *= 2;
  if (splits > p.ntiles_total) splits = p.ntiles_total;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.ntiles_total + splits - 1) / splits;
  splits = (p.ntiles_total + p.tiles_per_split - 1) / p.tiles_per_split;
  const size_t xstage = size_t(p.nchunks) * kWgChunk;
  int xs = int((size_t(200 * 1024) - 4 * kWgChunk) / xstage);
  p.xstages = xs > 4 ? 4 : xs;
  B21_CHECK_ARG(p.xstages >= 2, "conv3d_wgrad: N tile %d does not fit", p.BN);

  CUtensorMap tmX, tmZ;
  {
    const uint64_t dims[5] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)d, (uint64_t)n};
    const uint64_t str[4] = {uint64_t(ldx) * 2, uint64_t(w) * ldx * 2, uint64_t(h) * w * ldx * 2, uint64_t(d) * h * w * ldx * 2};
    const uint32_t box[5] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)td, 1};
    int r = encode_tmap_bf16(&tmX, x, 5, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  {
    const uint64_t dims[5] = {(uint64_t)cout, (uint64_t)w, (uint64_t)h, (uint64_t)d, (uint64_t)n};
    const uint64_t str[4] = {uint64_t(lddz) * 2, uint64_t(w) * lddz * 2, uint64_t(h) * w * lddz * 2, uint64_t(d) * h * w * lddz * 2};
    const uint32_t box[5] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)td, 1};
    int r = encode_tmap_bf16(&tmZ, dz, 5, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  static bool attr_set = false;
  if (!attr_set) {
    B21_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    attr_set = true;
  }
  const size_t smem_bytes = 4 * size_t(kWgChunk) + size_t(p.xstages) * xstage + 1024;
  dim3 grid(units, splits);
  conv_wgrad_kernel<<<grid, kWgThreads, smem_bytes, stream>>>(tmX, tmZ, p);
  B21_LAUNCH_CHECK("conv_wgrad_kernel");
  return B21_OK;
}

// conv_input.cu — the FIRST convolution of the networks (k = 3, dilation 1, at most 4 real input channels, e.g. the
// four MRI modalities -> 48) for sm_100a.
//
// conv_march.cu runs this layer as 9 tcgen05.mma per 128-voxel plane with K = 16 of which 4 are real channels, and —
// with so little tensor work per plane — it is bound by the per-plane hand-shakes of its pipeline (two mbarrier waits,
// two or three tcgen05.commit, one scout message per plane, profiles/r01i_march_variants.md): 0.8-0.9 ms per batch of
// 9 x 128^3 windows against 0.33 ms of HBM time.  With only 8 bytes per voxel there is nothing to gain from keeping
// the halo implicit, so this kernel does the opposite:
//
//   * K is the im2col of ONE input plane: k = (kh * 3 + kw) * 4 + c, 36 real elements padded to 48 = 3 k-steps.  The raw
//     18 x 10 halo planes arrive by TMA (one box of 18 rows x 160 B per plane, zero fill = the convolution's padding, a
//     ring of eight planes hides the HBM latency); two builder warps turn each into a "plane block"
//     ([6 chunks of 8 k][128 voxel rows][16 B], the canonical SWIZZLE_NONE K-major layout) shared -> shared.
//   * The three kd taps are folded into N as in conv_march.cu (B = [W(kd=2) | W(kd=1) | W(kd=0)]), but the output planes
//     are grouped in QUADS that own one TMEM buffer (4 x Cout fp32 columns, double-buffered): block i of a quad feeds
//     the up-to-three adjacent output planes i-2 .. i that lie inside the quad: 18 MMAs (N = Cout, 2 Cout or 3 Cout)
//     per four planes, no accumulator ring shared between hand-shake units.
//   * The bias rides on the MMA: K elements 36 / 37 of every A row are 1.0 and the kd = 1 weight rows 36 / 37 hold the
//     bias split into two bf16 (hi + lo, relative error 2^-16).
//   * Every hand-shake is per block PAIR (builders <-> issuer) or per output QUAD (issuer <-> epilogue); the four
//     epilogue groups drain the four planes of a quad side by side (tcgen05.ld -> re-zero -> GroupNorm/EvoNorm group
//     statistics -> [swish] -> bf16) with packed fp32x2 arithmetic, 16-column passes; each warp stages its 32 voxels x
//     Cout tile in shared memory and ONE TMA store (cp.async.bulk.tensor, box {Cout, 8 w, 4 h}) writes it: a thread-per-voxel
//     st.global.v4 touches 24 lines per instruction and cost 576 LSU cycles per plane (0.3 ms of a 0.85 ms launch).
//   * Warp roles: warp 0 TMA producer, warps 1-2 builders, warp 3 TMEM owner + MMA issuer, warps 4-19 four epilogue groups.
//
// Replaces torch.nn.Conv3d(inplanes, features[0], 3, padding=1) at networks/equiunet2020.py:19-25 (encoder1.ConvBnRelu1)
// and networks/equiunet2021.py:198 (encoder1.conv_evo_1) on both the inference and the training forward pass.
#include "ptx.cuh"
#include "fold.cuh"
#include "host_common.h"
#include "pack.cuh"
#include <stdlib.h>

namespace b21 {

constexpr int kITH = 16, kITW = 8;                   // in-plane output tile: M = 128 rows = 16 h x 8 w
constexpr int kIBuildWarps = 3, kIEpiGroups = 4;     // warp 0 = TMA producer of the raw halo planes, warps 1-2 = builders
constexpr int kIThreads = (kIBuildWarps + 1 + 4 * kIEpiGroups) * 32;  // 640: 96 registers per thread
constexpr int kIHH = kITH + 2, kIHW = kITW + 2;      // raw halo plane: 18 rows x 10 voxels x 16 B (8 channels, 4 real)
constexpr int kIRawData = kIHH * kIHW * 16;          // 2880 B of TMA payload
constexpr int kIRawBytes = (kIRawData + 127) / 128 * 128;
constexpr int kIRawSlots = 8;
constexpr int kIChunk = 2048;                        // 128 rows x 16 B: one chunk of 8 K elements
constexpr int kIBlockChunks = 6;                     // 48 K elements per plane block
constexpr int kIBlockBytes = kIBlockChunks * kIChunk;
constexpr int kIStageBytes = 2 * kIBlockBytes;       // a stage = a PAIR of plane blocks
constexpr int kIMaxStages = 8;
constexpr int kISmemBudget = 225 * 1024;

struct ConvInputParams {
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  const uint8_t* wpk;
  const float* bias;
  double* stats;
  int N, D, H, W, ldx, ldy;
  int tilesH, tilesW, segs, L, items;
  int stages, act;
  uint32_t wbytes;
  int variant;  // debug (B21_INPUT_VARIANT): 1 no global stores, 2 epilogue hand-shake only, 4 no MMAs, 8 no TMA loads of the raw planes
};

struct InputItem {
  int n, h0, w0, d0, Lc;
};
__device__ __forceinline__ InputItem input_decode(const ConvInputParams& p, int item) {
  InputItem it;
  int t = item;
  const int wt = t % p.tilesW; t /= p.tilesW;
  const int ht = t % p.tilesH; t /= p.tilesH;
  const int sg = t % p.segs;
  it.n = t / p.segs;
  it.h0 = ht * kITH;
  it.w0 = wt * kITW;
  it.d0 = sg * p.L;
  it.Lc = p.D - it.d0 < p.L ? p.D - it.d0 : p.L;
  return it;
}

// packed fp32 pairs (FADD2 / FMUL2 / FFMA2): half the issue slots of the epilogue arithmetic
__device__ __forceinline__ uint64_t f2_bits(float2 v) { return *reinterpret_cast<uint64_t*>(&v); }
__device__ __forceinline__ float2 bits_f2(uint64_t v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  uint64_t d;
  asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  uint64_t d;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(d);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
  return bits_f2(d);
}

template <int COUT>
__global__ void __launch_bounds__(kIThreads, 1) conv_input_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const ConvInputParams p) {
  constexpr int NG3 = 3 * COUT / 8;                  // 8-column groups of the B operand [W(kd=2) | W(kd=1) | W(kd=0)]
  constexpr uint32_t kBChunk = NG3 * 128;            // one chunk of 8 K elements of the weight image
  constexpr int GS = COUT / 8;                       // channels per norm group (even: fp32x2 pairs never straddle groups)
  constexpr uint32_t kBufCols = 4 * COUT;            // one TMEM buffer = the four output planes of a quad
  static_assert(GS % 2 == 0 && 2 * kBufCols <= 512, "bad Cout");
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kIMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kIMaxStages];
  __shared__ __align__(8) uint64_t accf_bar[2];
  __shared__ __align__(8) uint64_t acce_bar[2];
  __shared__ __align__(8) uint64_t rfull_bar[kIRawSlots];
  __shared__ __align__(8) uint64_t rempty_bar[kIRawSlots];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_stat[kIEpiGroups][2][16];  // [epilogue group][double buffer][8 groups x (sum, sumsq)]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const uint32_t w_addr = smem_u32(smem);
  const uint32_t a_off = (p.wbytes + 127u) & ~127u;
  const uint32_t a_addr = w_addr + a_off;
  // per epilogue warp: one 32-voxel x Cout bf16 tile (4 h rows x 8 w, dense = the TMA store box), 128-byte aligned
  constexpr uint32_t kStageTile = 32 * COUT * 2;
  const uint32_t y_addr = a_addr + uint32_t(p.stages) * uint32_t(kIStageBytes);
  const uint32_t raw_addr = y_addr + uint32_t(4 * kIEpiGroups) * kStageTile;
  const uint32_t rfull0 = smem_u32(&rfull_bar[0]), rempty0 = smem_u32(&rempty_bar[0]);
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t accf0 = smem_u32(&accf_bar[0]), acce0 = smem_u32(&acce_bar[0]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 64);  // every builder thread arrives after its own stores and proxy fence
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&accf_bar[b], 1);
      mbar_init(&acce_bar[b], 4 * kIEpiGroups);  // every epilogue warp, once per quad
    }
    for (int s = 0; s < kIRawSlots; ++s) {
      mbar_init(&rfull_bar[s], 1);
      mbar_init(&rempty_bar[s], 1);  // the builder warp that owns the plane's block
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (threadIdx.x < kIEpiGroups * 32) s_stat[threadIdx.x >> 5][(threadIdx.x >> 4) & 1][threadIdx.x & 15] = 0.f;
  {
    // weights: [6 chunks][3 Cout / 8][8 n][8 k] bf16, n = (2 - kd) * Cout + co, resident for the whole kernel
    const uint4* src = reinterpret_cast<const uint4*>(p.wpk);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = threadIdx.x; i < p.wbytes / 16; i += kIThreads) dst[i] = __ldg(src + i);
    // K elements 36..47 of every plane block are padding, written once: 36 and 37 are 1.0 (they pick up the bias rows
    // of the weight image), the rest is zero; the builders never touch them
    const uint32_t blocks = uint32_t(p.stages) * 2;
    for (uint32_t i = threadIdx.x; i < blocks * 128; i += kIThreads) {
      uint8_t* blk = smem + a_off + size_t(i >> 7) * kIBlockBytes + (i & 127) * 16;
      *reinterpret_cast<uint2*>(blk + 4 * kIChunk + 8) = make_uint2(0x3F803F80u, 0u);
      *reinterpret_cast<uint4*>(blk + 5 * kIChunk) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  __syncthreads();
  if (threadIdx.x < COUT) {
    // bias rows: k = 36 (hi) and 37 (lo) of the kd = 1 columns (n = Cout + co)
    const float bv = p.bias ? p.bias[threadIdx.x] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(bv);
    const __nv_bfloat16 lo = __float2bfloat16_rn(bv - __bfloat162float(hi));
    const int n = COUT + threadIdx.x;
    __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(smem + 4 * kBChunk + size_t(n >> 3) * 128 + (n & 7) * 16);
    row[4] = hi;
    row[5] = lo;
  }
  fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
  if (warp == kIBuildWarps) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: raw halo planes
    // one box of 18 rows x 160 B (10 voxels x 8 channels, contiguous in HBM) per input plane, zero fill = padding
    if (elect_one()) {
      uint32_t slot = 0, ph = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const InputItem it = input_decode(p, item);
        for (int i = 0; i <= it.Lc + 1; ++i) {
          const int dz = it.d0 - 1 + i;
          if (dz < 0 || dz >= p.D) continue;
          mbar_wait_sleep_a(rempty0 + 8u * slot, ph ^ 1u);
          if (p.variant & 8) {
            mbar_arrive_a(rfull0 + 8u * slot);
          } else {
            mbar_expect_tx_a(rfull0 + 8u * slot, uint32_t(kIRawData));
            tma_load_5d_a(raw_addr + slot * uint32_t(kIRawBytes), &tmX, rfull0 + 8u * slot, (it.w0 - 1) * 8, it.h0 - 1, dz, it.n, 0);
          }
          if (++slot == uint32_t(kIRawSlots)) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp < kIBuildWarps) {
    // ------------------------------------------------------------------ builders: im2col of one plane per block
    // shared -> shared: for its two voxel rows a thread reads the nine taps (8 bytes = 4 channels each) of the raw halo
    // plane and writes the five 16-byte K chunks of the A block.  (First version: 8-byte cp.async straight from global
    // memory — every pair paid one memory latency, because the proxy fence after the copies compiles to MEMBAR.ALL.CTA and
    // also waits for the thread's younger copies; three pairs in flight gave 0.38 ms per launch for the builders alone.)
    // builder warp bw owns block bw of every pair (input planes 2s + bw): one raw-plane wait per pair and warp
    const int bw = warp - 1;
    uint32_t g = 0;  // block pairs of the CTA so far
    uint32_t rcount = 0;  // raw planes consumed by BOTH builder warps so far (valid blocks, in order)
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const InputItem it = input_decode(p, item);
      const int nbp = (it.Lc + 3) >> 1;  // blocks i = 0 .. Lc + 1 (input planes d0 - 1 + i), in pairs
      for (int s = 0; s < nbp; ++s, ++g) {
        const uint32_t stage = g % uint32_t(p.stages), ph = (g / uint32_t(p.stages)) & 1u;
        // valid blocks of this pair: the raw ring holds them in order (block 0 first)
        const int i0 = 2 * s, i1 = 2 * s + 1;
        const bool v0 = i0 <= it.Lc + 1 && it.d0 - 1 + i0 >= 0 && it.d0 - 1 + i0 < p.D;
        const bool v1 = i1 <= it.Lc + 1 && it.d0 - 1 + i1 >= 0 && it.d0 - 1 + i1 < p.D;
        const bool mine = bw == 0 ? v0 : v1;
        const uint32_t rc = rcount + (bw == 1 && v0 ? 1u : 0u);  // ring position of this warp's block
        rcount += (v0 ? 1u : 0u) + (v1 ? 1u : 0u);
        mbar_wait_spin_a(empty0 + 8u * stage, ph ^ 1u);
        if (mine) {
          const uint32_t rslot = rc % uint32_t(kIRawSlots), rph = (rc / uint32_t(kIRawSlots)) & 1u;
          mbar_wait_spin_a(rfull0 + 8u * rslot, rph);
          const uint8_t* raw = smem + (raw_addr - w_addr) + size_t(rslot) * kIRawBytes;
          uint8_t* blk = smem + a_off + size_t(stage) * kIStageBytes + size_t(bw) * kIBlockBytes;
          uint2 t[4][9];
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const int r = rr * 32 + lane;  // voxel row of the tile = TMEM lane of its outputs
            const uint8_t* src = raw + (r >> 3) * (kIHW * 16) + (r & 7) * 16;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw)
                t[rr][kh * 3 + kw] = *reinterpret_cast<const uint2*>(src + kh * (kIHW * 16) + kw * 16);
          }
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            uint8_t* dst = blk + (rr * 32 + lane) * 16;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(dst + c * kIChunk) =
                  make_uint4(t[rr][2 * c].x, t[rr][2 * c].y, t[rr][2 * c + 1].x, t[rr][2 * c + 1].y);
            *reinterpret_cast<uint2*>(dst + 4 * kIChunk) = t[rr][8];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_a(rempty0 + 8u * rslot);  // the raw plane has been read
          fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        }
        mbar_arrive_a(full0 + 8u * stage);
      }
    }
  } else if (warp == kIBuildWarps) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint64_t dA = umma_smem_desc(0, kIChunk, 128, kLayoutNone);   // LBO = next 8 k, SBO = next 8 rows
      const uint64_t dB = umma_smem_desc(0, kBChunk, 128, kLayoutNone);
      const uint32_t id1 = umma_idesc_bf16(128, COUT), id2 = umma_idesc_bf16(128, 2 * COUT), id3 = umma_idesc_bf16(128, 3 * COUT);
      const uint32_t stages = uint32_t(p.stages);
      uint32_t bc_base = 0;  // block pairs of the previous items
      uint32_t waited = 0;   // block pairs whose full barrier has been consumed
      uint32_t rel = 0;      // block pairs released to the builders
      uint32_t qc = 0;       // output quads issued
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const InputItem it = input_decode(p, item);
        const int nbp = (it.Lc + 3) >> 1, nq = (it.Lc + 3) >> 2;
        for (int t = 0; t < nq; ++t, ++qc) {
          // outputs 4t .. 4t + 3 read blocks 4t .. 4t + 5 = block pairs 2t .. 2t + 2
          const int last_pair = 2 * t + 2 < nbp - 1 ? 2 * t + 2 : nbp - 1;
          while (waited < bc_base + uint32_t(last_pair) + 1u) {
            mbar_wait_a(full0 + 8u * (waited % stages), (waited / stages) & 1u);
            ++waited;
          }
          const uint32_t buf = qc & 1u;
          mbar_wait_a(acce0 + 8u * buf, (qc >> 1) & 1u);  // drained AND re-zeroed: every MMA accumulates
          tc_fence_after();
          const int o_base = 4 * t;
          const int o_max = o_base + 3 < it.Lc - 1 ? o_base + 3 : it.Lc - 1;
          if (!(p.variant & 4)) {
#pragma unroll
            for (int ii = 0; ii < 6; ++ii) {
              const int i = o_base + ii;
              const int dz = it.d0 - 1 + i;
              const int o_lo = ii <= 2 ? o_base : i - 2;
              const int o_hi = i < o_max ? i : o_max;
              if (i <= it.Lc + 1 && dz >= 0 && dz < p.D && o_hi >= o_lo) {
                const int n = o_hi - o_lo + 1;
                const uint32_t jlo = uint32_t(2 - (i - o_lo));  // first column group of B: kd = 2 - j
                const uint32_t dcol = tmem_base + buf * kBufCols + uint32_t(o_lo - o_base) * COUT;
                const uint32_t idesc = n == 3 ? id3 : (n == 2 ? id2 : id1);
                const uint32_t st = (bc_base + uint32_t(i >> 1)) % stages;
                const uint64_t ad = dA + uint64_t((a_addr + st * uint32_t(kIStageBytes) + uint32_t(i & 1) * kIBlockBytes) >> 4);
                const uint64_t bd = dB + uint64_t((w_addr + jlo * uint32_t(COUT * 16)) >> 4);
#pragma unroll
                for (int ks = 0; ks < 3; ++ks)
                  umma_bf16(dcol, ad + uint64_t(ks) * (2u * kIChunk >> 4), bd + uint64_t(ks) * (2u * kBChunk >> 4), idesc, 1u);
              }
            }
          }
          // block pairs 2t and 2t + 1 have served their last outputs (the last quad releases the rest of the item)
          const int done_pair = (t == nq - 1 || 2 * t + 1 > nbp - 1) ? nbp - 1 : 2 * t + 1;
          while (rel < bc_base + uint32_t(done_pair) + 1u) {
            umma_commit_a(empty0 + 8u * (rel % stages));
            ++rel;
          }
          umma_commit_a(accf0 + 8u * buf);
        }
        bc_base += uint32_t(nbp);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (TMEM lane quadrant = warp % 4)
    // group g drains output plane 4t + g of every quad
    const int quad = warp & 3;
    const int grp = (warp - kIBuildWarps - 1) >> 2;
    const int row = quad * 32 + lane;
    const int hh = row >> 3, ww = row & 7;
    const uint32_t tlane = tmem_base + (uint32_t(quad * 32) << 16);
    // accumulators start at zero and are re-zeroed after every drain: the MMAs always accumulate
    for (uint32_t b = 0; b < 2; ++b) {
      for (uint32_t c0 = 0; c0 < uint32_t(COUT); c0 += 16) tmem_st16_zero(tlane + b * kBufCols + uint32_t(grp * COUT) + c0);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive_a(acce0);
      mbar_arrive_a(acce0 + 8u);
    }
    uint32_t qc = 0;
    int sbuf = 0;
    const uint32_t y_tile = y_addr + uint32_t(warp - kIBuildWarps - 1) * kStageTile;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const InputItem it = input_decode(p, item);
      const int h = it.h0 + hh, w = it.w0 + ww;
      const bool valid = (h < p.H) && (w < p.W);
      float2 gs[8], gq[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) gs[g] = gq[g] = make_float2(0.f, 0.f);
      const int nq = (it.Lc + 3) >> 2;
      for (int t = 0; t < nq; ++t, ++qc) {
        const uint32_t buf = qc & 1u;
        mbar_wait_sleep_a(accf0 + 8u * buf, (qc >> 1) & 1u);
        tc_fence_after();
        const int o = 4 * t + grp;
        if (o >= it.Lc || (p.variant & 2)) {  // the plane does not exist (plane count not a multiple of four)
          __syncwarp();
          if (lane == 0) mbar_arrive_a(acce0 + 8u * buf);
          continue;
        }
        const uint32_t tcol = tlane + buf * kBufCols + uint32_t(grp * COUT);
        // the previous TMA store of this warp has finished reading the staging tile
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 16) {
          float v[16];
          tmem_ld16(tcol + uint32_t(c0), v);
          tmem_ld_wait();
          tmem_st16_zero(tcol + uint32_t(c0));
          if (c0 + 16 >= COUT) {  // last pass: the accumulator is in registers and zeroed again
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(acce0 + 8u * buf);
          }
          uint32_t pk[8];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const float2 x2 = make_float2(v[c], v[c + 1]);
            if (valid) {
              gs[(c0 + c) / GS] = add2(gs[(c0 + c) / GS], x2);
              gq[(c0 + c) / GS] = fma2(x2, x2, gq[(c0 + c) / GS]);
            }
            float2 o2 = x2;
            if (p.act == 2) {  // x * sigmoid(x) = h + h * tanh(h), h = x / 2 (fold.cuh: swishf_tanh)
              const float2 h2 = mul2(x2, make_float2(0.5f, 0.5f));
              float2 t2;
              asm("tanh.approx.f32 %0, %1;" : "=f"(t2.x) : "f"(h2.x));
              asm("tanh.approx.f32 %0, %1;" : "=f"(t2.y) : "f"(h2.y));
              o2 = fma2(h2, t2, h2);
            } else if (p.act == 1) {
              o2 = make_float2(swishf(x2.x), swishf(x2.y));
            }
            pk[c >> 1] = pack_bf16x2(o2.x, o2.y);
          }
          // staging tile [32 voxels][Cout] bf16: this thread's voxel record
          uint8_t* srow = smem + (y_tile - w_addr) + size_t(lane) * (COUT * 2) + c0 * 2;
          *reinterpret_cast<uint4*>(srow) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(srow + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        fence_proxy_async();  // this thread's staging writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0 && !(p.variant & 1)) {
          // box {Cout, 8 w, 4 h, 1, 1}: rows and columns beyond H / W are clipped by the TMA unit
          tma_store_5d(&tmY, y_tile, 0, it.w0, it.h0 + quad * 4, it.d0 + o, it.n);
          bulk_commit();
        }
      }
      if (p.stats) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float a = warp_sum(gs[g].x + gs[g].y), b = warp_sum(gq[g].x + gq[g].y);
          if (lane == 0) {
            atomicAdd(&s_stat[grp][sbuf][g * 2], a);
            atomicAdd(&s_stat[grp][sbuf][g * 2 + 1], b);
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");  // the four warps of this epilogue group only
        if (quad == 2 && lane < 16) {
          const float sv = s_stat[grp][sbuf][lane];
          s_stat[grp][sbuf][lane] = 0.f;
          if (sv != 0.f) {
            const int slot = item % B21_STAT_SLOTS;
            atomicAdd(p.stats + ((size_t(slot) * p.N + it.n) * 8) * 2 + lane, double(sv));
          }
        }
        sbuf ^= 1;
      }
    }
    if (lane == 0) bulk_wait0();  // the last TMA stores have left shared memory (and completed) before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kIBuildWarps) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ weight repack
// out = shared-memory image [6 chunks][3 cout / 8][8 n][8 k] bf16, n = (2 - kd) * cout + co, k = (kh * 3 + kw) * 4 + c
// (36 real, 48 padded; rows 36 / 37 of the kd = 1 columns receive the bias inside the kernel)
__global__ void pack_input_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin_o) {
  const size_t total = size_t(3) * kIBlockChunks * cout * 8;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16_rn(pack_input_value(w, i, cout, cin_o));
}

static inline size_t input_wbytes(int cout) { return size_t(3) * kIBlockChunks * (cout / 8) * 128; }
static inline size_t input_staging_bytes(int cout) { return size_t(4 * kIEpiGroups) * 32 * cout * 2; }
static inline int input_stages(int cout) {
  const size_t wb = (input_wbytes(cout) + 127) & ~size_t(127);
  size_t st = (size_t(kISmemBudget) - wb - 128 - input_staging_bytes(cout) - size_t(kIRawSlots) * kIRawBytes) / kIStageBytes;
  return int(st > size_t(kIMaxStages) ? size_t(kIMaxStages) : st);
}

template <int COUT>
static int launch_input(const CUtensorMap& tmX, const CUtensorMap& tmY, const ConvInputParams& p, size_t smem_bytes, int grid,
                        cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    B21_CUDA(cudaFuncSetAttribute(conv_input_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kISmemBudget));
    attr_set = true;
  }
  conv_input_kernel<COUT><<<grid, kIThreads, smem_bytes, stream>>>(tmX, tmY, p);
  B21_LAUNCH_CHECK("conv_input_kernel");
  return B21_OK;
}

}  // namespace b21

using namespace b21;

extern "C" int b21_conv_input_supported(int cin_true, int cout) {
  if (cin_true < 1 || cin_true > 4) return 0;
  return (cout == 16 || cout == 32 || cout == 48 || cout == 64) ? 1 : 0;
}

extern "C" long long b21_conv_input_weight_bytes(int cout) { return (long long)input_wbytes(cout); }

extern "C" int b21_pack_conv_weight_input(const float* w, void* packed, int cout, int cin, void* stream) {
  B21_CHECK_ARG(w && packed, "pack_conv_weight_input: null pointer");
  B21_CHECK_ARG(b21_conv_input_supported(cin, cout), "pack_conv_weight_input: (cin %d, cout %d) unsupported", cin, cout);
  const size_t total = input_wbytes(cout) / 2;
  pack_input_weight_kernel<<<int((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w, reinterpret_cast<__nv_bfloat16*>(packed), cout, cin);
  B21_LAUNCH_CHECK("pack_input_weight_kernel");
  return B21_OK;
}

extern "C" int b21_pack_job_input(const float* w, void* packed, int cout, int cin, b21_pack_job* job) {
  B21_CHECK_ARG(w && packed && job, "pack_job_input: null pointer");
  B21_CHECK_ARG(b21_conv_input_supported(cin, cout), "pack_job_input: (cin %d, cout %d) unsupported", cin, cout);
  job->w = w; job->out = packed; job->total = (long long)(input_wbytes(cout) / 2);
  job->kind = kPackInput; job->cout = cout; job->cin = cin; job->tf = 0;
  job->p0 = cout; job->p1 = 0; job->p2 = 0; job->p3 = 0; job->blk0 = 0; job->nblk = 0;
  return B21_OK;
}

// y = conv3d(x, W) + bias [-> swish] with the group statistics of the pre-activation values; x holds the (at most 4)
// real channels in the first 8 bytes of every dense 16-byte voxel record (ldx = 8).  act: 0 none, 1 swish (ex2 + rcp),
// 2 swish through tanh.approx (fold.cuh).
extern "C" int b21_conv3d_input_fwd(const void* x, int ldx, const void* w_input, const float* bias, void* y, int ldy,
                                    double* stats, int act, int n, int d, int h, int w, int cout, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B21_CHECK_ARG(x && w_input && y, "conv3d_input_fwd: null pointer");
  B21_CHECK_ARG(n > 0 && d > 0 && h > 0 && w > 0, "conv3d_input_fwd: bad shape %d %d %d %d", n, d, h, w);
  B21_CHECK_ARG(b21_conv_input_supported(4, cout), "conv3d_input_fwd: cout %d unsupported", cout);
  B21_CHECK_ARG(act >= 0 && act <= 2, "conv3d_input_fwd: act must be 0, 1 or 2 (got %d)", act);
  B21_CHECK_ARG(ldx == 8 && ldy >= cout && ldy % 8 == 0, "conv3d_input_fwd: ldx must be 8 (dense 8-channel records, got %d), ldy %d", ldx, ldy);
  B21_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(w_input) & 15) == 0,
                "conv3d_input_fwd: pointers must be 16-byte aligned");
  ConvInputParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.wpk = reinterpret_cast<const uint8_t*>(w_input);
  p.bias = bias;
  p.stats = stats;
  p.N = n; p.D = d; p.H = h; p.W = w; p.ldx = ldx; p.ldy = ldy;
  p.tilesH = (h + kITH - 1) / kITH;
  p.tilesW = (w + kITW - 1) / kITW;
  p.stages = input_stages(cout);
  p.act = act;
  p.wbytes = (uint32_t)input_wbytes(cout);
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("B21_INPUT_VARIANT");
    variant = e ? atoi(e) : 0;
  }
  p.variant = variant;
  // march length: the longest segment that still balances the persistent grid (each segment rebuilds 2 halo planes)
  const int sms = num_sms();
  const long long tiles = (long long)n * p.tilesH * p.tilesW;
  int bestL = d;
  double best = -1.0;
  for (int segs = 1; segs <= d; ++segs) {
    const int L = (d + segs - 1) / segs;
    if (L < 8 && segs > 1) break;
    if ((d + L - 1) / L != segs) continue;
    const long long items = tiles * segs;
    const long long rounds = (items + sms - 1) / sms;
    const double eff = double(items) / double(rounds * sms) * double(L) / double(L + 2);
    if (eff > best + 1e-9) {
      best = eff;
      bestL = L;
    }
  }
  p.L = bestL;
  p.segs = (d + p.L - 1) / p.L;
  p.items = int(tiles * p.segs);
  const int grid = p.items < sms ? p.items : sms;
  if (stats) B21_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * B21_STAT_SLOTS * n * 16, stream));
  const size_t smem_bytes = ((size_t(p.wbytes) + 127) & ~size_t(127)) + size_t(p.stages) * kIStageBytes +
                            input_staging_bytes(cout) + size_t(kIRawSlots) * kIRawBytes + 128;
  CUtensorMap tmX, tmY;
  {
    // (w, c) merged into one dimension: a halo row of 10 voxels x 8 channels is one contiguous 160-byte run
    const uint64_t dims[5] = {(uint64_t)w * 8, (uint64_t)h, (uint64_t)d, (uint64_t)n, 1};
    const uint64_t str[4] = {uint64_t(w) * 16, uint64_t(h) * w * 16, uint64_t(d) * h * w * 16, uint64_t(n) * d * h * w * 16};
    const uint32_t box[5] = {8 * (uint32_t)kIHW, (uint32_t)kIHH, 1, 1, 1};
    int r = encode_tmap_bf16(&tmX, x, 5, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r) return r;
  }
  {
    const uint64_t dims[5] = {(uint64_t)cout, (uint64_t)w, (uint64_t)h, (uint64_t)d, (uint64_t)n};
    const uint64_t str[4] = {uint64_t(ldy) * 2, uint64_t(w) * ldy * 2, uint64_t(h) * w * ldy * 2, uint64_t(d) * h * w * ldy * 2};
    const uint32_t box[5] = {(uint32_t)cout, (uint32_t)kITW, 4, 1, 1};
    int r = encode_tmap_bf16(&tmY, y, 5, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r) return r;
  }
  switch (cout) {
    case 16: return launch_input<16>(tmX, tmY, p, smem_bytes, grid, stream);
    case 32: return launch_input<32>(tmX, tmY, p, smem_bytes, grid, stream);
    case 48: return launch_input<48>(tmX, tmY, p, smem_bytes, grid, stream);
    default: return launch_input<64>(tmX, tmY, p, smem_bytes, grid, stream);
  }
}

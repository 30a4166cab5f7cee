// conv_point.cu — persistent 1x1x1 convolution (a [voxels x Cin] * [Cin x Cout] GEMM) for sm_100a.
//
// The 1x1 layers of the hot path (ConvEvo bridges / up-convs: networks/equiunet2021.py:214-222,262-269) carry
// almost no FLOPs: they are HBM-bound (read Cin, write Cout bf16 per voxel).  Launching one CTA per 128-voxel
// tile (conv_tap.cu) makes them CTA-launch-latency bound instead (bridge1, 48 -> 24 at 128^3: 10x off the HBM
// roofline), so this kernel is persistent:
//
//   * one CTA per SM walks a contiguous range of 128-voxel tiles; the weights [Cout_pad x Cin] are loaded ONCE into
//     shared memory (TMA, 128B-swizzled K-major) and stay resident;
//   * the activation tiles stream through an 8-stage TMA ring (one {64 ch x 128 voxel} box per stage; channel
//     counts that are not multiples of 64 are zero-filled by TMA and only ceil(valid/16) MMAs are issued);
//   * tcgen05.mma (M = 128, N = Cout_pad, K = 16) accumulates into a DOUBLE-BUFFERED TMEM accumulator so that the
//     epilogue of tile i (tcgen05.ld -> +bias -> group statistics -> bf16 store) overlaps the loads/MMAs of tile i+1;
//   * GroupNorm/EvoNorm group statistics are accumulated in registers across all tiles of a sample and flushed
//     with one round of double atomics per CTA.
//
// Tile GROUPS (round 2): one 128-voxel tile per hand-shake made the kernel protocol-bound (two mbarrier waits and two
// tcgen05.commit per 3-6 MMAs: 1520 cycles per tile against 755 of HBM time at 48 -> 24, 9 x 128^3).  A pipeline stage now
// holds G = 4 (2 for Cout = 96) consecutive tiles of one sample, a TMEM buffer their G accumulators, and every barrier /
// commit is per group.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 and 6..9 = two epilogue groups:
// group g drains accumulator buffer g (tiles with it % 2 == g), so that one group's tcgen05.ld / MUFU / store
// latency chain overlaps the other's (a single group left the SM idle for half of every tile: 1570 cycles per tile
// against 810 cycles of HBM time at 48 -> 24).
#include "ptx.cuh"
#include "fold.cuh"
#include "host_common.h"

namespace b21 {

constexpr int kPtThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..5 / 6..9: two epilogue groups (alternate tiles)
constexpr int kPtABytes = 128 * 128;  // one stage: 128 voxels x 64 bf16
constexpr int kPtMaxStages = 8;
constexpr int kPtSmemBudget = 225 * 1024;
__host__ __device__ constexpr int point_group(int bn) { return bn <= 64 ? 4 : 2; }  // tiles per pipeline stage

struct ConvPointParams {
  __nv_bfloat16* y;
  const float* bias;
  double* stats;
  int N, Cin, ldy;
  long long nvox;      // voxels per sample
  int tiles_per_n, tiles, chunks, stages;
  FoldExtras ex;  // table = per-sample bias [N][Cout]; wstride != 0: per-sample weights (third coordinate of tmB)
};

template <int COUT>
__global__ void __launch_bounds__(kPtThreads, 1)
conv_point_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const ConvPointParams p) {
  constexpr int BN = (COUT + 15) / 16 * 16;
  constexpr int GS = COUT / 8;
  constexpr int G = point_group(BN);
  constexpr uint32_t kBufCols = G * BN;  // one TMEM buffer = the accumulators of a tile group
  constexpr uint32_t TCOLS = 2 * kBufCols <= 32 ? 32 : (2 * kBufCols <= 64 ? 64 : (2 * kBufCols <= 128 ? 128 : (2 * kBufCols <= 256 ? 256 : 512)));
  static_assert(2 * kBufCols <= 512, "accumulator buffers exceed TMEM");
  constexpr uint32_t kStageBytes = G * kPtABytes;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kPtMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kPtMaxStages];
  __shared__ __align__(8) uint64_t accf_bar[2];
  __shared__ __align__(8) uint64_t acce_bar[2];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                   // [chunks][BN rows][128 B]
  uint8_t* sA = smem + size_t(p.chunks) * BN * 128;     // [stages][G tiles][128 rows][128 B]

  // contiguous tile range of this CTA
  const int per = p.tiles / gridDim.x, rem = p.tiles % gridDim.x;
  const int t_begin = blockIdx.x * per + (int(blockIdx.x) < rem ? blockIdx.x : rem);
  const int t_end = t_begin + per + (int(blockIdx.x) < rem ? 1 : 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&accf_bar[b], 1);
      mbar_init(&acce_bar[b], 4);
    }
    mbar_init(&w_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  for (int c = threadIdx.x; c < BN; c += kPtThreads) s_bias[c] = (p.bias && c < COUT) ? p.bias[c] : 0.f;
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, TCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      int w_n = -1;
      for (int tile = t_begin; tile < t_end;) {
        const int n = tile / p.tiles_per_n;
        int g = (n + 1) * p.tiles_per_n - tile;  // group: up to G consecutive tiles of one sample
        g = g > G ? G : g;
        g = g > t_end - tile ? t_end - tile : g;
        const int r0 = (tile - n * p.tiles_per_n) * 128;
        if (w_n < 0 || (p.ex.wstride != 0 && n != w_n)) {
          if (w_n >= 0) {  // every activation stage released <=> every MMA that read the old weights has completed
            for (int k = 0; k < p.stages; ++k) {
              const int s2 = s + k < p.stages ? s + k : s + k - p.stages;
              mbar_wait_sleep(&empty_bar[s2], (s + k < p.stages ? ph : ph ^ 1) ^ 1);
            }
          }
          w_n = n;
          mbar_expect_tx(&w_bar, uint32_t(p.chunks) * BN * 128);
          for (int ck = 0; ck < p.chunks; ++ck)
            tma_load_3d(sB + size_t(ck) * BN * 128, &tmB, &w_bar, ck * 64, 0, p.ex.wstride != 0 ? n : 0);
        }
        for (int ck = 0; ck < p.chunks; ++ck) {
          mbar_wait_sleep(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], uint32_t(g) * kPtABytes);
          for (int j = 0; j < g; ++j)
            tma_load_3d(sA + size_t(s) * kStageBytes + size_t(j) * kPtABytes, &tmA, &full_bar[s], ck * 64, r0 + j * 128, n);
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
        tile += g;
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, BN);
      const uint32_t b_addr = smem_u32(sB), a_addr = smem_u32(sA);
      int s = 0;
      uint32_t ph = 0;
      uint32_t it = 0;
      int w_n = -1;
      uint32_t w_par = 0;
      for (int tile = t_begin; tile < t_end; ++it) {
        const int n = tile / p.tiles_per_n;
        int g = (n + 1) * p.tiles_per_n - tile;
        g = g > G ? G : g;
        g = g > t_end - tile ? t_end - tile : g;
        if (w_n < 0 || (p.ex.wstride != 0 && n != w_n)) {
          w_n = n;
          mbar_wait(&w_bar, w_par);
          w_par ^= 1u;
          tc_fence_after();
        }
        const uint32_t buf = it & 1u;
        mbar_wait(&acce_bar[buf], ((it >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t dcol = tmem_base + buf * kBufCols;
        for (int ck = 0; ck < p.chunks; ++ck) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          int nk = (p.Cin - ck * 64 + 15) >> 4;
          nk = nk > 4 ? 4 : nk;
          const uint32_t b0 = b_addr + uint32_t(ck) * BN * 128;
          for (int j = 0; j < g; ++j) {
            const uint32_t a0 = a_addr + uint32_t(s) * kStageBytes + uint32_t(j) * kPtABytes;
            for (int k = 0; k < nk; ++k) {
              const uint64_t ad = umma_smem_desc(a0 + k * 32, 16, 1024, kLayoutSw128);
              const uint64_t bd = umma_smem_desc(b0 + k * 32, 16, 1024, kLayoutSw128);
              umma_bf16(dcol + uint32_t(j) * BN, ad, bd, idesc, (ck | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&accf_bar[buf]);
        tile += g;
      }
    }
  } else {
    const int quad = warp & 3;
    const uint32_t grp = uint32_t(warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t tlane = tmem_base + (uint32_t(quad * 32) << 16);
    float gs[8], gq[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) gs[g] = gq[g] = 0.f;
    int cur_n = -1;
    auto flush = [&](int n) {
      if (!p.stats || n < 0) return;
      const int slot = blockIdx.x % B21_STAT_SLOTS;
      double* dst = p.stats + ((size_t(slot) * p.N + n) * 8) * 2;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float a = warp_sum(gs[g]), b = warp_sum(gq[g]);
        if (lane == 0) {
          atomicAdd(dst + g * 2, double(a));
          atomicAdd(dst + g * 2 + 1, double(b));
        }
        gs[g] = gq[g] = 0.f;
      }
    };
    uint32_t it = 0;
    for (int tile = t_begin; tile < t_end; ++it) {
      const int n = tile / p.tiles_per_n;
      int g = (n + 1) * p.tiles_per_n - tile;
      g = g > G ? G : g;
      g = g > t_end - tile ? t_end - tile : g;
      const int tile0 = tile;
      tile += g;
      if ((it & 1u) != grp) continue;  // the other group's tiles (and accumulator buffer)
      if (n != cur_n) {
        flush(cur_n);
        cur_n = n;
      }
      const uint32_t buf = it & 1u;
      mbar_wait_sleep(&accf_bar[buf], (it >> 1) & 1u);
      tc_fence_after();
      const float* tb = p.ex.table ? p.ex.table + size_t(n) * COUT : nullptr;
      for (int j = 0; j < g; ++j) {
        const long long r = (long long)(tile0 + j - n * p.tiles_per_n) * 128 + row;
        const bool valid = r < p.nvox;
        float v[BN];
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 16) tmem_ld16(tlane + buf * kBufCols + uint32_t(j * BN + c0), v + c0);
        tmem_ld_wait();
        if (j == g - 1) {  // the whole buffer is in registers: the issuer may start the next group here
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acce_bar[buf]);
        }
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          const float val = v[c] + (tb ? __ldg(tb + c) : s_bias[c]);
          const float sv = valid ? val : 0.f;
          gs[c / GS] += sv;
          gq[c / GS] = fmaf(sv, sv, gq[c / GS]);
          v[c] = val;
        }
        if (p.ex.act) {  // one branch around the whole unrolled loop: the ex2/rcp chains interleave
#pragma unroll
          for (int c = 0; c < COUT; ++c) v[c] = swishf(v[c]);
        }
        if (valid) {
          __nv_bfloat16* yrow = p.y + (size_t(n) * p.nvox + r) * size_t(p.ldy);
#pragma unroll
          for (int c0 = 0; c0 < COUT; c0 += 8) {
            uint4 o;
            o.x = pack_bf16x2(v[c0 + 0], v[c0 + 1]);
            o.y = pack_bf16x2(v[c0 + 2], v[c0 + 3]);
            o.z = pack_bf16x2(v[c0 + 4], v[c0 + 5]);
            o.w = pack_bf16x2(v[c0 + 6], v[c0 + 7]);
            *reinterpret_cast<uint4*>(yrow + c0) = o;
          }
        }
      }
    }
    flush(cur_n);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TCOLS);
}

static inline int point_bn(int cout) { return (cout + 15) / 16 * 16; }
static inline int point_stages(int cin, int cout) {
  const size_t wb = size_t((cin + 63) / 64) * point_bn(cout) * 128;
  const size_t stage = size_t(point_group(point_bn(cout))) * kPtABytes;
  if (wb + 1024 + 2 * stage > size_t(kPtSmemBudget)) return 0;
  size_t st = (size_t(kPtSmemBudget) - 1024 - wb) / stage;
  return int(st > kPtMaxStages ? kPtMaxStages : st);
}

template <int COUT>
static int launch_point(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvPointParams& p, size_t smem_bytes,
                        int grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    B21_CUDA(cudaFuncSetAttribute(conv_point_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kPtSmemBudget));
    attr_set = true;
  }
  conv_point_kernel<COUT><<<grid, kPtThreads, smem_bytes, stream>>>(tmA, tmB, p);
  B21_LAUNCH_CHECK("conv_point_kernel");
  return B21_OK;
}

}  // namespace b21

using namespace b21;

extern "C" int b21_conv_point_supported(int cin, int cout) {
  if (!(cout == 8 || cout == 16 || cout == 24 || cout == 32 || cout == 48 || cout == 64 || cout == 96)) return 0;
  if (cin <= 0 || cin % 8) return 0;
  return point_stages(cin, cout) >= 2 ? 1 : 0;
}

static int point_fwd_impl(const void* x, int ldx, const void* w_packed, const float* bias, void* y, int ldy,
                          double* stats, int n, long long nvox, int cin, int cout, const FoldExtras& ex, void* stream_);

extern "C" int b21_conv1x1_fwd(const void* x, int ldx, const void* w_packed, const float* bias, void* y, int ldy,
                               double* stats, int n, long long nvox, int cin, int cout, void* stream_) {
  FoldExtras ex = {nullptr, nullptr, 0, 0};
  return point_fwd_impl(x, ldx, w_packed, bias, y, ldy, stats, n, nvox, cin, cout, ex, stream_);
}

// Folded-EvoNorm variant (see fold.cu): `per_sample` != 0 -> w_packed holds n weight sets (b21_pack_conv_weight_fold),
// bias_n (or NULL -> `bias`) is a per-sample bias [n][cout]; act != 0 stores x*sigmoid(x).
extern "C" int b21_conv1x1_fwd_fold(const void* x, int ldx, const void* w_packed, int per_sample, const float* bias,
                                    const float* bias_n, void* y, int ldy, double* stats, int act, int n,
                                    long long nvox, int cin, int cout, void* stream_) {
  FoldExtras ex = {bias_n, nullptr, per_sample ? 1 : 0, act};
  return point_fwd_impl(x, ldx, w_packed, bias, y, ldy, stats, n, nvox, cin, cout, ex, stream_);
}

static int point_fwd_impl(const void* x, int ldx, const void* w_packed, const float* bias, void* y, int ldy,
                          double* stats, int n, long long nvox, int cin, int cout, const FoldExtras& ex, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B21_CHECK_ARG(x && w_packed && y, "conv1x1_fwd: null pointer");
  B21_CHECK_ARG(n > 0 && nvox > 0 && nvox < (1ll << 31), "conv1x1_fwd: bad shape n %d nvox %lld", n, nvox);
  B21_CHECK_ARG(b21_conv_point_supported(cin, cout), "conv1x1_fwd: (cin %d, cout %d) unsupported", cin, cout);
  B21_CHECK_ARG(ldx >= cin && ldx % 8 == 0 && ldy >= cout && ldy % 8 == 0, "conv1x1_fwd: bad ldx %d / ldy %d", ldx, ldy);
  B21_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
                "conv1x1_fwd: pointers must be 16-byte aligned");
  ConvPointParams p;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.bias = bias;
  p.stats = stats;
  p.N = n; p.Cin = cin; p.ldy = ldy; p.nvox = nvox;
  p.tiles_per_n = int((nvox + 127) / 128);
  p.tiles = p.tiles_per_n * n;
  p.chunks = (cin + 63) / 64;
  p.stages = point_stages(cin, cout);
  p.ex = ex;
  const int bn = point_bn(cout);

  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[3] = {(uint64_t)cin, (uint64_t)nvox, (uint64_t)n};
    const uint64_t str[2] = {uint64_t(ldx) * 2, uint64_t(nvox) * ldx * 2};
    const uint32_t box[3] = {64, 128, 1};
    int r = encode_tmap_bf16(&tmA, x, 3, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  {
    // packed weight of b21_pack_conv_weight(k = 1): [1][cout_padded = bn][cin]
    const uint64_t dims[3] = {(uint64_t)cin, (uint64_t)bn, (uint64_t)(ex.wstride != 0 ? n : 1)};
    const uint64_t str[2] = {uint64_t(cin) * 2, uint64_t(bn) * cin * 2};
    const uint32_t box[3] = {64, (uint32_t)bn, 1};
    int r = encode_tmap_bf16(&tmB, w_packed, 3, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  if (stats) B21_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * B21_STAT_SLOTS * n * 16, stream));
  const int sms = num_sms();
  const int grid = p.tiles < sms ? p.tiles : sms;
  const size_t smem_bytes = size_t(p.chunks) * bn * 128 + size_t(p.stages) * point_group(bn) * kPtABytes + 1024;
  switch (cout) {
    case 8: return launch_point<8>(tmA, tmB, p, smem_bytes, grid, stream);
    case 16: return launch_point<16>(tmA, tmB, p, smem_bytes, grid, stream);
    case 24: return launch_point<24>(tmA, tmB, p, smem_bytes, grid, stream);
    case 32: return launch_point<32>(tmA, tmB, p, smem_bytes, grid, stream);
    case 48: return launch_point<48>(tmA, tmB, p, smem_bytes, grid, stream);
    case 64: return launch_point<64>(tmA, tmB, p, smem_bytes, grid, stream);
    default: return launch_point<96>(tmA, tmB, p, smem_bytes, grid, stream);
  }
}

// conv_march.cu — "plane-marching" implicit-GEMM conv3d (k = 3, dilation 1, stride 1) for sm_100a.
//
// The tap-streaming kernel (conv_tap.cu) re-fetches its activation tile from L2 once per tap (27x), which makes the
// small-channel layers that hold most of the FLOPs (48 -> 48 at full resolution) L2-bandwidth bound, and with
// N = Cout = 48 every tcgen05.mma spends more cycles reading its 4 KB A tile from shared memory than multiplying.
// This kernel removes both limits:
//
//   * One persistent CTA per SM owns an (h, w) tile of 16 x 8 output voxels (M = 128) and MARCHES along d.  Each
//     input plane (18 x 10 halo, all channels) is brought into shared memory exactly once by TMA (out-of-bounds
//     zero fill = the convolution's zero padding) and serves all 27 taps: a tap is just a different start address
//     of the UMMA shared-memory descriptor.  The plane is stored channel-chunk-major ([Cin/8][18][10][8 ch],
//     SWIZZLE_NONE "interleaved" K-major core matrices: 8 consecutive w voxels x 16 B = one 128 B core matrix), so
//     arbitrary (kh, kw) shifts keep the canonical layout: SBO = 160 B (next h row), LBO = 2944 B (next 8 channels).
//   * All 27 x Cin x Cout weights stay resident in shared memory (loaded once per CTA with cp.async.bulk).
//   * The three kd taps are folded into the N dimension: for input plane d' the MMA computes, per (kh, kw),
//     D[128, 3*Cout] += A[128, Cin] * [W(kd=2) | W(kd=1) | W(kd=0)], whose three column groups belong to output
//     planes d'-1, d', d'+1.  TMEM holds a ring of per-output-plane accumulators (Cout fp32 columns each); the three
//     groups are adjacent ring slots.  N = 144 instead of 48 amortises the A-tile read over 3x the math.
//   * Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2 .. 2+4*NG-1 = NG epilogue groups of
//     four warps (tcgen05.ld -> + bias / bias table -> GroupNorm/EvoNorm group statistics -> [swish] -> bf16 NDHWC
//     store [-> SE channel sums]) that take the output planes round-robin and overlap with the MMAs of later planes,
//     last warp = SCOUT.  tcgen05.mma issue is nearly synchronous (the pipe accepts one or two MMAs ahead), so the
//     issuing thread does nothing but issue and commit: the scout walks the same plane sequence, does the ring / border
//     arithmetic, waits on the mbarriers (plane data, freshly claimed accumulator slots, weights) and hands a 48-byte
//     ready-made descriptor per plane to the issuer through a shared-memory ring.
//   * Template parameters: COUT; NG = epilogue groups (2; 4 with 16-column passes for the Cin = 8 input conv, whose 9 MMAs
//     per plane leave the epilogue on the critical path); KS = compile-time k-steps per tap (two uniform adds per MMA
//     instead of a branchy remainder loop; 0 = runtime loop).
//   * Input variants: `merged` (Cin = 8, dense): the plane is ONE TMA box of 18 rows x 160 B over a merged (w, c)
//     dimension instead of 180 rows x 16 B, and the zero second k-chunk is written once; `split` (b21_..._fold2): the
//     8-channel chunks come from TWO tensors, i.e. a channel concat is read in place from its two dense producers.
//
// Replaces torch.nn.Conv3d(k=3, padding=1) at networks/equiunet2020.py:19-25 and networks/equiunet2021.py:198,201
// for the layers whose weights fit in shared memory (54 * ceil16(Cin) * Cout bytes); other shapes use conv_tap.cu.
#include "ptx.cuh"
#include "fold.cuh"
#include "host_common.h"
#include "pack.cuh"
#include <stdlib.h>

namespace b21 {

// warp 0 TMA, warp 1 MMA, then NG epilogue groups of four warps (planes are dealt round-robin to the groups).
// NG = 2 (320 threads) everywhere except the Cin = 8 input conv: with 9 MMAs per plane its epilogue (about 800
// instructions per warp and plane, issue-latency bound at 2 warps per scheduler: 2000 cycles per plane against 650 of
// MMAs) is the bottleneck, so it runs NG = 4 groups (576 threads, <= 112 registers: 16-column passes).
constexpr int kMThreadsBase = 64;
constexpr int kMTH = 16, kMTW = 8;                // in-plane output tile: M = 128 rows = 16 h x 8 w
constexpr int kMHH = kMTH + 2, kMHW = kMTW + 2;   // halo plane 18 x 10
constexpr int kMChunkData = kMHH * kMHW * 16;     // one 8-channel chunk of a halo plane (2880 B of TMA payload)
constexpr int kMChunkBytes = (kMChunkData + 127) / 128 * 128;  // chunk stride: TMA destinations are 128 B aligned
constexpr int kMMaxStages = 6;
constexpr int kMMaxRing = 16;
constexpr int kMSmemBudget = 225 * 1024;

struct ConvMarchParams {
  __nv_bfloat16* y;
  const uint8_t* wpk;
  const float* bias;
  double* stats;
  int N, D, H, W, ldy;
  int kc;  // 8-channel chunks per plane (Cin rounded up to 16, divided by 8)
  int tilesH, tilesW, segs, L, items;
  int stages, ring;
  uint32_t wbytes;
  FoldExtras ex;
  int split;    // 8-channel chunks [split, kc) come from a SECOND tensor (tmX2): channel concat without a concat buffer
  int merged;   // Cin == 8 dense input (ld = 8): (w, c) merged into one tensor-map dimension, chunk 1 is zero in smem
  int variant;  // debug (B21_MARCH_VARIANT, see profiles/r01i_march_variants.md): 4 no TMA loads, 8 no stores, 16 one kh row of
                // taps only, 32 epilogue handshake only, 64 no TMEM ld/st, 128 two epilogue groups for the input conv,
                // 256 runtime k-step loop, 4096 no MMAs
};

__device__ __forceinline__ uint64_t nosw_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return umma_smem_desc(saddr, lbo, sbo, kLayoutNone);
}

struct MarchItem {
  int n, h0, w0, d0, Lc;
};
__device__ __forceinline__ MarchItem decode_item(const ConvMarchParams& p, int item) {
  MarchItem it;
  int t = item;
  const int wt = t % p.tilesW; t /= p.tilesW;
  const int ht = t % p.tilesH; t /= p.tilesH;
  const int sg = t % p.segs;
  it.n = t / p.segs;
  it.h0 = ht * kMTH;
  it.w0 = wt * kMTW;
  it.d0 = sg * p.L;
  it.Lc = p.D - it.d0 < p.L ? p.D - it.d0 : p.L;
  return it;
}

// scout -> issuer message: everything the issuing thread needs for one input plane (48 bytes, three 16-byte words)
struct PlaneMsg {
  uint32_t a_lo, a_hi, b_lo, b_hi;   // A / B shared-memory descriptors of the plane's first tap / k-step
  uint32_t col0, id0, id1, b1;       // accumulator column, instruction descriptor (id1 / b1: unused since the ring no longer wraps)
  uint32_t stage, accf_lo, accf_hi, pad;  // plane stage to release, ring slots to commit (0xffffffff = none / end)
};
constexpr int kMsgRing = 8;

// KS = compile-time number of 16-channel k-steps (0: runtime loop).  With a runtime count the compiler emits a
// branchy remainder loop around every tcgen05.mma (about 80 cycles of issue per MMA against 72 of execution at
// N = 144, so the tensor pipe drains during the per-plane bookkeeping); fully unrolled, an MMA costs two uniform
// 64-bit adds and the issue thread runs ahead of the pipe.
template <int COUT, int NG, int KS>
__global__ void __launch_bounds__(kMThreadsBase + NG * 128 + 32, 1)
conv_march_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX2,
                  const __grid_constant__ CUtensorMap tmY, const ConvMarchParams p) {
  constexpr int kMThreads = kMThreadsBase + NG * 128 + 32;  // + the scout warp
  constexpr int PC = NG > 2 ? 16 : COUT;  // columns per epilogue pass
  // Accumulator slots in TMEM: RING logical slots + 2 OVERFLOW slots.  The three column groups of a plane are adjacent
  // slots r_lo .. r_lo + 2; a plain ring wraps for 2 of every RING planes and needs two MMAs (N = 96 and 48: 56 + 45
  // cycles instead of 72 for one N = 144 MMA).  With the physical slots RING and RING + 1 standing in for logical
  // slots 0 and 1 whenever a window starts at RING - 2 or RING - 1, every plane is ONE contiguous MMA; the epilogue
  // sums (and re-zeroes) both physical copies of logical slots 0 and 1.  Barriers are per LOGICAL slot.
  constexpr uint32_t PHYS = (512 / COUT) < kMMaxRing ? (512 / COUT) : kMMaxRing;
  constexpr uint32_t RING = PHYS - 2;
  static_assert(RING >= 4, "accumulator ring too short");
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMMaxStages];
  __shared__ __align__(8) uint64_t accf_bar[kMMaxRing];
  __shared__ __align__(8) uint64_t acce_bar[kMMaxRing];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_stat[NG][2][16];  // [epilogue group][double buffer][8 groups x (sum, sumsq)]
  __shared__ float s_bias[COUT];
  __shared__ __align__(16) PlaneMsg msg_ring[kMsgRing];
  __shared__ uint32_t msg_ready;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const uint32_t w_addr = smem_u32(smem);
  const uint32_t p_addr = w_addr + ((p.wbytes + 127u) & ~127u);
  const uint32_t plane_bytes = uint32_t(p.kc) * kMChunkBytes;
  // output staging of the TMA-store epilogue: one dense 32-voxel x COUT bf16 tile (4 h rows x 8 w) per epilogue warp
  constexpr uint32_t kStageTile = 32 * COUT * 2;
  const uint32_t y_addr = p_addr + uint32_t(p.stages) * plane_bytes;
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t accf0 = smem_u32(&accf_bar[0]), acce0 = smem_u32(&acce_bar[0]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (uint32_t r = 0; r < RING; ++r) {
      mbar_init(&accf_bar[r], 1);
      mbar_init(&acce_bar[r], 4);
    }
    mbar_init(&w_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmX2);
    tma_prefetch_desc(&tmY);
  }
  if (threadIdx.x == 0) msg_ready = 0;
  if (threadIdx.x < NG * 32) s_stat[threadIdx.x >> 5][(threadIdx.x >> 4) & 1][threadIdx.x & 15] = 0.f;
  for (int c = threadIdx.x; c < COUT; c += kMThreads) s_bias[c] = p.bias ? p.bias[c] : 0.f;
  if (p.merged) {
    // the input has 8 channels but K = 16 per MMA: the second chunk of every stage is zeroed once and never loaded
    for (int s = 0; s < p.stages; ++s) {
      uint4* z = reinterpret_cast<uint4*>(smem + (p_addr - w_addr) + size_t(s) * plane_bytes + kMChunkBytes);
      for (int i = threadIdx.x; i < kMChunkBytes / 16; i += kMThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: weights once, then planes
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int w_n = -1;  // sample whose weights are resident (per-sample weights: folded EvoNorm affine)
      const uint32_t tx = p.merged ? uint32_t(kMChunkData) : uint32_t(p.kc) * kMChunkData;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const MarchItem it = decode_item(p, item);
        if (w_n < 0 || (p.ex.wstride != 0 && it.n != w_n)) {
          if (w_n >= 0) {
            // every plane stage released <=> every MMA that read the old weights has completed
            for (int k = 0; k < p.stages; ++k) {
              const int s2 = stage + k < p.stages ? stage + k : stage + k - p.stages;
              mbar_wait_sleep_a(empty0 + 8u * s2, (stage + k < p.stages ? phase : phase ^ 1) ^ 1);
            }
          }
          w_n = it.n;
          const uint8_t* wsrc = p.wpk + size_t(p.ex.wstride) * it.n;
          mbar_expect_tx(&w_bar, p.wbytes);
          for (uint32_t off = 0; off < p.wbytes; off += 16384u) {
            const uint32_t nb = p.wbytes - off < 16384u ? p.wbytes - off : 16384u;
            bulk_load_1d(smem + off, wsrc + off, nb, &w_bar);
          }
        }
        for (int i = 0; i <= it.Lc + 1; ++i) {
          const int dz = it.d0 - 1 + i;
          if (dz < 0 || dz >= p.D) continue;
          mbar_wait_sleep_a(empty0 + 8u * stage, phase ^ 1);
          const uint32_t fb = full0 + 8u * stage;
          if (p.variant & 4) {  // debug: no loads
            mbar_arrive_a(fb);
          } else {
            const uint32_t dst = p_addr + uint32_t(stage) * plane_bytes;
            mbar_expect_tx_a(fb, tx);
            if (p.merged) {
              // one box of 18 rows x 160 B (10 voxels x 8 channels, contiguous in HBM) instead of 180 rows x 16 B:
              // the TMA engine is request-rate bound on 16 B rows (2000 cycles per plane against 650 of MMAs)
              tma_load_5d_a(dst, &tmX, fb, (it.w0 - 1) * 8, it.h0 - 1, dz, it.n, 0);
            } else {
              for (int c = 0; c < p.split; ++c)
                tma_load_5d_a(dst + c * kMChunkBytes, &tmX, fb, c * 8, it.w0 - 1, it.h0 - 1, dz, it.n);
              for (int c = p.split; c < p.kc; ++c)
                tma_load_5d_a(dst + c * kMChunkBytes, &tmX2, fb, (c - p.split) * 8, it.w0 - 1, it.h0 - 1, dz, it.n);
            }
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2 + 4 * NG) {
    // ------------------------------------------------------------------ MMA issuer (warp 1) + its scout (last warp)
    // tcgen05.mma issue is close to synchronous (the pipe accepts one or two MMAs ahead), so every cycle the issuing
    // thread spends on anything else is a cycle the tensor pipe idles.  Measured per plane on the 48 -> 48 layer
    // (profiles/r01i_march_variants.md): 27 MMAs 1780 cycles, ring/border arithmetic 270, the two mbarrier waits 600-900
    // (even when both phases completed long ago), commits 300.  The arithmetic and the waits therefore live on a SCOUT
    // thread (one elected lane of the last warp): it walks the same plane sequence, waits for the plane's data
    // (full barrier), for freshly claimed accumulator slots (acce barriers) and for the weights, and publishes a
    // ready-made descriptor through a small shared-memory ring; the issuer only polls a counter, reads 48 bytes and
    // issues.  The scout can never be more than `stages` planes ahead (it waits on the same full barriers), so a ring of
    // 8 entries needs no back-pressure.
    if (elect_one()) {
      constexpr uint32_t lboB = (3 * COUT / 8) * 128, sboB = 128;
      constexpr uint32_t lboA = kMChunkBytes, sboA = kMHW * 16;
      constexpr uint32_t kEnd = 0xffffffffu;
      if (warp != 1) {
        // ---------------------------------------------------------------- scout
        const uint64_t dA = nosw_desc(0, lboA, sboA), dB = nosw_desc(0, lboB, sboB);
        const uint32_t idesc1 = umma_idesc_bf16(128, COUT), idesc2 = umma_idesc_bf16(128, 2 * COUT),
                       idesc3 = umma_idesc_bf16(128, 3 * COUT);
        const uint64_t bd0 = dB + uint64_t(w_addr >> 4);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t sg_base = 0, next_fresh = 0, r_base = 0, claim_par = 0;
        int w_n = -1;
        uint32_t w_par = 0, k = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
          const MarchItem it = decode_item(p, item);
          if (w_n < 0 || (p.ex.wstride != 0 && it.n != w_n)) {
            w_n = it.n;
            mbar_wait_spin(&w_bar, w_par);
            w_par ^= 1u;
          }
          uint32_t r_lo = r_base;
          for (int i = 0; i <= it.Lc + 1; ++i) {
            // column group j (0..2) = tap kd = 2 - j = output plane (local, 1-based) so = i - 1 + j; the lowest valid
            // group of plane i sits in ring slot r_lo, which advances by one per plane once i >= 3.
            if (i >= 3) r_lo = r_lo + 1 == RING ? 0 : r_lo + 1;
            const int dz = it.d0 - 1 + i;
            if (dz < 0 || dz >= p.D) continue;
            const int jlo = i >= 2 ? 0 : 2 - i;
            const int jhi = i + 1 <= it.Lc ? 2 : it.Lc + 1 - i;
            const int ngroups = jhi - jlo + 1;
            const uint32_t sg_lo = sg_base + uint32_t(i + jlo - 2);
#pragma unroll
            for (int g = 0; g < 3; ++g) {
              const uint32_t sg = sg_lo + uint32_t(g);
              if (g < ngroups && sg >= next_fresh) {  // first contribution: wait until the slot is drained and zeroed
                uint32_t r = r_lo + uint32_t(g);
                r = r >= RING ? r - RING : r;
                mbar_wait_spin_a(acce0 + 8u * r, (claim_par >> r) & 1u);
                claim_par ^= 1u << r;
                next_fresh = sg + 1;
              }
            }
            mbar_wait_spin_a(full0 + 8u * stage, phase);
            PlaneMsg m;
            const uint64_t a_row = dA + uint64_t((p_addr + uint32_t(stage) * plane_bytes) >> 4);
            const uint64_t bq = bd0 + uint64_t(jlo) * COUT;  // descriptor address field is in 16 B units
            m.a_lo = uint32_t(a_row); m.a_hi = uint32_t(a_row >> 32);
            m.b_lo = uint32_t(bq); m.b_hi = uint32_t(bq >> 32);
            m.col0 = tmem_base + r_lo * COUT;  // physical slots r_lo .. r_lo + ngroups - 1 <= RING + 1: never wraps
            m.id0 = ngroups == 3 ? idesc3 : (ngroups == 2 ? idesc2 : idesc1);
            m.id1 = 0;
            m.b1 = 0;
            m.stage = uint32_t(stage);
            m.accf_lo = i >= 2 ? r_lo : kEnd;  // output plane so = i - 1 is complete after this plane
            m.accf_hi = kEnd;
            if (i == it.Lc && it.d0 + it.Lc >= p.D) {  // no plane i + 1 exists: so = i is complete as well
              const uint32_t r = r_lo + uint32_t(1 - jlo);
              m.accf_hi = r >= RING ? r - RING : r;
            }
            m.pad = 0;
            uint4* dst = reinterpret_cast<uint4*>(&msg_ring[k & (kMsgRing - 1)]);
            const uint4* src = reinterpret_cast<const uint4*>(&m);
            dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
            ++k;
            __threadfence_block();
            *reinterpret_cast<volatile uint32_t*>(&msg_ready) = k;
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
          sg_base += uint32_t(it.Lc);
          r_base = (r_base + uint32_t(it.Lc)) % RING;
        }
        msg_ring[k & (kMsgRing - 1)].stage = kEnd;  // end marker
        ++k;
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t*>(&msg_ready) = k;
      } else {
        // ---------------------------------------------------------------- issuer
        const uint64_t a_step = 2u * (lboA >> 4), b_step = 2u * (lboB >> 4);
        const int ksteps = p.kc >> 1;
        const int nkh = (p.variant & 4096) ? 0 : ((p.variant & 16) ? 1 : 3);  // debug: bit4 one row of taps, bit12 no MMAs
        for (uint32_t k = 0;; ++k) {
          uint32_t spins = 0;
          while (*reinterpret_cast<volatile uint32_t*>(&msg_ready) <= k) {
            if (++spins > (1u << 28)) {
              printf("b21: march issuer starved block %d plane %u\n", blockIdx.x, k);
              __trap();
            }
          }
          __threadfence_block();
          PlaneMsg m;
          {
            const uint4* src = reinterpret_cast<const uint4*>(&msg_ring[k & (kMsgRing - 1)]);
            uint4* dst = reinterpret_cast<uint4*>(&m);
            dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
          }
          if (m.stage == kEnd) break;
          tc_fence_after();
          const uint64_t a_row = (uint64_t(m.a_hi) << 32) | m.a_lo, bq0 = (uint64_t(m.b_hi) << 32) | m.b_lo;
          if constexpr (KS > 0) {
            constexpr uint64_t kAStep = 2u * (uint64_t(kMChunkBytes) >> 4), kBStep = 2u * (uint64_t((3 * COUT / 8) * 128) >> 4);
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {  // one MMA per (tap, k-step), N = ngroups * COUT
              if (kh < nkh) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                  for (int ks = 0; ks < KS; ++ks)
                    umma_bf16(m.col0, a_row + uint64_t(kh * kMHW + kw) + uint64_t(ks) * kAStep,
                              bq0 + uint64_t((kh * 3 + kw) * KS + ks) * kBStep, m.id0, 1u);
                }
              }
            }
          } else {
            uint64_t a_kh = a_row, bq = bq0;
            for (int kh = 0; kh < nkh; ++kh, a_kh += kMHW) {
              uint64_t a_tap = a_kh;
              for (int kw = 0; kw < 3; ++kw, ++a_tap) {
                uint64_t ad = a_tap;
                for (int ks = 0; ks < ksteps; ++ks, ad += a_step, bq += b_step) {
                  umma_bf16(m.col0, ad, bq, m.id0, 1u);
                }
              }
            }
          }
          umma_commit_a(empty0 + 8u * m.stage);
          if (m.accf_lo != kEnd) umma_commit_a(accf0 + 8u * m.accf_lo);
          if (m.accf_hi != kEnd) umma_commit_a(accf0 + 8u * m.accf_hi);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (TMEM lane quadrant = warp % 4)
    // Two groups of four warps drain alternate output planes (the epilogue, not the MMA, bounds the small-Cin layers).
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int hh = row >> 3, ww = row & 7;
    const uint32_t tlane = tmem_base + (uint32_t(quad * 32) << 16);
    constexpr int GS = COUT / 8;  // channels per norm group
    // accumulators start at zero and are re-zeroed after every drain: the MMAs always accumulate
    if (grp == 0) {
      for (uint32_t c0 = 0; c0 < PHYS * COUT; c0 += 16) tmem_st16_zero(tlane + c0);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0)
        for (uint32_t r = 0; r < RING; ++r) mbar_arrive_a(acce0 + 8u * r);
    }
    uint32_t r = 0, use_par = 0, plane_cnt = 0;
    int buf = 0;
    const uint32_t y_tile = y_addr + uint32_t(warp - 2) * kStageTile;
    uint8_t* y_row = smem + (y_tile - w_addr) + size_t(lane) * (COUT * 2);  // this thread's voxel record of the tile
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const MarchItem it = decode_item(p, item);
      const int h = it.h0 + hh, w = it.w0 + ww;
      const bool valid = (h < p.H) && (w < p.W);
      float gs[8], gq[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) gs[g] = gq[g] = 0.f;
      constexpr int NCS = (COUT + 31) / 32;
      float csum[NCS];  // lane l: running sum of the stored outputs of channels l, 32 + l, ...
#pragma unroll
      for (int b = 0; b < NCS; ++b) csum[b] = 0.f;
      const float* trow = p.ex.table
                              ? p.ex.table + (size_t(it.n) * 27 + border_class(h, p.H) * 3 + border_class(w, p.W)) * COUT
                              : nullptr;
      for (int so = 0; so < it.Lc; ++so) {
        if (int((plane_cnt++) % uint32_t(NG)) != grp) {  // another group's plane: only advance the ring cursor
          use_par ^= 1u << r;
          r = r + 1 == RING ? 0 : r + 1;
          continue;
        }
        mbar_wait_sleep_a(accf0 + 8u * r, (use_par >> r) & 1u);
        use_par ^= 1u << r;
        tc_fence_after();
        if (lane == 0) bulk_wait_read0();  // the previous TMA store of this warp has finished reading the staging tile
        __syncwarp();
        const uint32_t tcol = tlane + r * COUT;
        const bool aliased = r < 2;  // logical slots 0 and 1 have a second physical copy (slots RING, RING + 1)
        const uint32_t tcol2 = tcol + RING * COUT;
        const uint32_t acce_r = acce0 + 8u * r;
        r = r + 1 == RING ? 0 : r + 1;
        const float* tb = (trow && valid) ? trow + size_t(border_class(it.d0 + so, p.D)) * 9 * COUT : nullptr;
        uint4 o[COUT / 8];
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += PC) {
          float v[PC];
          if (!(p.variant & 64)) {
#pragma unroll
            for (int c = 0; c < PC; c += 16) tmem_ld16(tcol + uint32_t(c0 + c), v + c);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < PC; c += 16) tmem_st16_zero(tcol + uint32_t(c0 + c));
            if (aliased) {
#pragma unroll
              for (int c = 0; c < PC; c += 16) {
                float t[16];
                tmem_ld16(tcol2 + uint32_t(c0 + c), t);
                tmem_ld_wait();
                tmem_st16_zero(tcol2 + uint32_t(c0 + c));
#pragma unroll
                for (int q = 0; q < 16; ++q) v[c + q] += t[q];
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < PC; ++c) v[c] = 0.f;
          }
          if (c0 + PC >= COUT) {  // last pass: the slot is drained and zeroed, the MMA warp may reuse it
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(acce_r);
          }
          if (p.variant & 32) continue;  // debug: handshake only
          if (tb) {  // folded input affine: the bias depends on the border class of the output voxel
            const float4* t4p = reinterpret_cast<const float4*>(tb + c0);
#pragma unroll
            for (int c4 = 0; c4 < PC / 4; ++c4) {
              const float4 t4 = __ldg(t4p + c4);
              v[c4 * 4 + 0] += t4.x; v[c4 * 4 + 1] += t4.y; v[c4 * 4 + 2] += t4.z; v[c4 * 4 + 3] += t4.w;
            }
          } else if (!trow) {
#pragma unroll
            for (int c = 0; c < PC; ++c) v[c] += s_bias[c0 + c];
          }
#pragma unroll
          for (int c = 0; c < PC; ++c) {
            const float sv = valid ? v[c] : 0.f;
            gs[(c0 + c) / GS] += sv;
            gq[(c0 + c) / GS] = fmaf(sv, sv, gq[(c0 + c) / GS]);
          }
          if (p.ex.act == 1) {  // one branch around the whole unrolled loop: the ex2/rcp chains interleave
#pragma unroll
            for (int c = 0; c < PC; ++c) v[c] = swishf(v[c]);
          } else if (p.ex.act == 2) {
#pragma unroll
            for (int c = 0; c < PC; ++c) v[c] = swishf_tanh(v[c]);
          }
#pragma unroll
          for (int c = 0; c < PC; c += 8) {
            uint4& q = o[(c0 + c) / 8];
            q.x = pack_bf16x2(v[c + 0], v[c + 1]);
            q.y = pack_bf16x2(v[c + 2], v[c + 3]);
            q.z = pack_bf16x2(v[c + 4], v[c + 5]);
            q.w = pack_bf16x2(v[c + 6], v[c + 7]);
            *reinterpret_cast<uint4*>(y_row + (c0 + c) * 2) = q;
          }
        }
        if (p.variant & 32) continue;
        // One TMA store per warp and plane (box {COUT, 8 w, 4 h}; rows / columns beyond H / W are clipped).  A
        // thread-per-voxel st.global.v4 touches 24 lines per instruction: 576 LSU cycles per 48-channel plane on the
        // L1 / shared-memory data path the tensor core reads its operands through (profiles/r02w_input_conv.md).
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !(p.variant & 8)) {
          tma_store_5d(&tmY, y_tile, 0, it.w0, it.h0 + quad * 4, it.d0 + so, it.n);
          bulk_commit();
        }
        if constexpr (PC == COUT) {
          if (p.ex.chan_sum) {  // SE squeeze: channel sums of what the consumer will read (the rounded values)
#pragma unroll
            for (int b = 0; b < NCS; ++b) {
              float t[32];
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const int c = b * 32 + i;
                float2 f = make_float2(0.f, 0.f);
                if (c < COUT) {
                  const uint4& q = o[c / 8];
                  const uint32_t wd = ((c % 8) / 2) == 0 ? q.x : (((c % 8) / 2) == 1 ? q.y : (((c % 8) / 2) == 2 ? q.z : q.w));
                  f = unpack_bf16x2(wd);
                }
                t[i] = valid ? f.x : 0.f;
                t[i + 1] = valid ? f.y : 0.f;
              }
              csum[b] += warp_transpose_sum32(t, lane);
            }
          }
        }
      }
      if (PC == COUT && p.ex.chan_sum) {
#pragma unroll
        for (int b = 0; b < NCS; ++b)
          if (b * 32 + lane < COUT) atomicAdd(p.ex.chan_sum + size_t(it.n) * COUT + b * 32 + lane, csum[b]);
      }
      if (p.stats) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float a = warp_sum(gs[g]), b = warp_sum(gq[g]);
          if (lane == 0) {
            atomicAdd(&s_stat[grp][buf][g * 2], a);
            atomicAdd(&s_stat[grp][buf][g * 2 + 1], b);
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");  // the four warps of this epilogue group only
        if (quad == 2 && lane < 16) {  // warp 2 / warp 6
          const float sv = s_stat[grp][buf][lane];
          s_stat[grp][buf][lane] = 0.f;
          if (sv != 0.f) {
            const int slot = item % B21_STAT_SLOTS;
            atomicAdd(p.stats + ((size_t(slot) * p.N + it.n) * 8) * 2 + lane, double(sv));
          }
        }
        buf ^= 1;
      }
    }
    if (lane == 0) bulk_wait0();  // the last TMA stores have completed before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ weight repack
// out = shared-memory image: [tap9 = kh*3+kw][kc][ng = 3*cout/8][8 n][8 k] bf16, n = j*cout + co with kd = 2 - j.
// scale != NULL: per-sample copies (blockIdx.y = sample) with the input channels multiplied by scale[sample][ci].
__global__ void pack_march_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout_o,
                                         int cin_o, int rows, int kc, int transpose_flip,
                                         const float* __restrict__ scale = nullptr, int ldscale = 0,
                                         int pack_blocks = 0, BiasTableArgs tab = BiasTableArgs()) {
  if (pack_blocks > 0 && int(blockIdx.x) >= pack_blocks) {  // appended blocks: one bias-table row each
    bias_table_block(tab, blockIdx.x - pack_blocks, blockIdx.y);
    return;
  }
  const int ng = 3 * rows / 8;
  const size_t total = size_t(9) * kc * ng * 64;
  const size_t gstride = size_t(pack_blocks > 0 ? pack_blocks : gridDim.x) * blockDim.x;
  out += size_t(blockIdx.y) * total;
  if (scale) scale += size_t(blockIdx.y) * ldscale;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += gstride)
    out[i] = __float2bfloat16_rn(pack_march_value(w, i, cout_o, cin_o, rows, kc, transpose_flip, scale));
}

static inline int march_kc(int cin) { return (cin + 15) / 16 * 2; }
static inline size_t march_wbytes(int cin, int cout) { return size_t(9) * march_kc(cin) * (3 * cout / 8) * 128; }

// output staging of the epilogue warps (TMA store): 16 warps cover the NG = 4 variant of the Cin = 8 layers
static inline size_t march_staging(int cin, int cout) { return size_t(cin <= 8 ? 16 : 8) * 32 * cout * 2; }
static int march_stages(int cin, int cout) {
  const size_t wb = (march_wbytes(cin, cout) + 127) & ~size_t(127);
  const size_t plane = size_t(march_kc(cin)) * kMChunkBytes;
  if (wb + 128 + march_staging(cin, cout) >= size_t(kMSmemBudget)) return 0;
  size_t st = (size_t(kMSmemBudget) - wb - 128 - march_staging(cin, cout)) / plane;
  return int(st > kMMaxStages ? kMMaxStages : st);
}

}  // namespace b21

using namespace b21;

extern "C" int b21_conv_march_supported(int cin, int cout) {
  if (!(cout == 16 || cout == 32 || cout == 48 || cout == 64)) return 0;
  if (cin <= 0 || cin % 8) return 0;
  return march_stages(cin, cout) >= 3 ? 1 : 0;
}

extern "C" long long b21_conv_march_weight_bytes(int cin, int cout) { return (long long)march_wbytes(cin, cout); }

extern "C" int b21_pack_conv_weight_march(const float* w, void* packed, int cout, int cin, int transpose_flip,
                                          void* stream) {
  B21_CHECK_ARG(w && packed, "pack_conv_weight_march: null pointer");
  const int rows = transpose_flip ? cin : cout, inner = transpose_flip ? cout : cin;
  B21_CHECK_ARG(rows % 8 == 0, "pack_conv_weight_march: output channels %d must be a multiple of 8", rows);
  const int kc = march_kc(inner);
  const size_t total = march_wbytes(inner, rows) / 2;
  const int threads = 256;
  const int blocks = int((total + threads - 1) / threads) < 2048 ? int((total + threads - 1) / threads) : 2048;
  pack_march_weight_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
      w, reinterpret_cast<__nv_bfloat16*>(packed), cout, cin, rows, kc, transpose_flip);
  B21_LAUNCH_CHECK("pack_march_weight_kernel");
  return B21_OK;
}

extern "C" int b21_pack_job_march(const float* w, void* packed, int cout, int cin, int transpose_flip, b21_pack_job* job) {
  B21_CHECK_ARG(w && packed && job, "pack_job_march: null pointer");
  const int rows = transpose_flip ? cin : cout, inner = transpose_flip ? cout : cin;
  B21_CHECK_ARG(rows % 8 == 0, "pack_job_march: output channels %d must be a multiple of 8", rows);
  job->w = w; job->out = packed; job->total = (long long)(march_wbytes(inner, rows) / 2);
  job->kind = kPackMarch; job->cout = cout; job->cin = cin; job->tf = transpose_flip;
  job->p0 = rows; job->p1 = march_kc(inner); job->p2 = 0; job->p3 = 0; job->blk0 = 0; job->nblk = 0;
  return B21_OK;
}

// Per-sample folded packing: packed[s] = pack(w * scale[s][ci]) for s < nsamples, b21_conv_march_weight_bytes apart.
extern "C" int b21_pack_conv_weight_march_fold(const float* w, void* packed, int cout, int cin, const float* scale,
                                               int ldscale, int nsamples, const float* ws, const float* bias,
                                               const float* b_in, float* table, void* stream) {
  B21_CHECK_ARG(w && packed && scale && nsamples > 0 && ldscale >= cin, "pack_conv_weight_march_fold: bad args");
  B21_CHECK_ARG(!table || (ws && b_in), "pack_conv_weight_march_fold: the bias table needs ws and B");
  B21_CHECK_ARG(cout % 8 == 0, "pack_conv_weight_march_fold: output channels %d must be a multiple of 8", cout);
  b21_pack_job job;  // source-tiled packing (pack.cuh); padding is not written: the caller zero-fills the buffer once
  int r = b21_pack_job_march(w, packed, cout, cin, 0, &job);
  if (r) return r;
  BiasTableArgs tab = {ws, bias, b_in, table, ldscale, cout, cin, 27};
  return launch_pack_fold_tile(reinterpret_cast<const PackJob&>(job), scale, ldscale, nsamples, tab, (cudaStream_t)stream);
}

template <int COUT, int NG, int KS>
static int launch_march_ng(const CUtensorMap& tm, const CUtensorMap& tm2, const CUtensorMap& tmY, const ConvMarchParams& p,
                           size_t smem_bytes, int grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    B21_CUDA(cudaFuncSetAttribute(conv_march_kernel<COUT, NG, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kMSmemBudget));
    attr_set = true;
  }
  conv_march_kernel<COUT, NG, KS><<<grid, kMThreadsBase + NG * 128 + 32, smem_bytes, stream>>>(tm, tm2, tmY, p);
  B21_LAUNCH_CHECK("conv_march_kernel");
  return B21_OK;
}

template <int COUT>
static int launch_march(const CUtensorMap& tm, const CUtensorMap& tm2, const CUtensorMap& tmY, const ConvMarchParams& p,
                        size_t smem_bytes, int grid, cudaStream_t stream) {
  const int ks = (p.variant & 256) ? 0 : p.kc >> 1;
  // epilogue-bound input conv (one real 8-channel chunk, no SE channel sums): four epilogue groups
  if (p.merged && !p.ex.chan_sum && !(p.variant & 128)) return launch_march_ng<COUT, 4, 1>(tm, tm2, tmY, p, smem_bytes, grid, stream);
  switch (ks) {
    case 1: return launch_march_ng<COUT, 2, 1>(tm, tm2, tmY, p, smem_bytes, grid, stream);
    case 2: return launch_march_ng<COUT, 2, 2>(tm, tm2, tmY, p, smem_bytes, grid, stream);
    case 3: return launch_march_ng<COUT, 2, 3>(tm, tm2, tmY, p, smem_bytes, grid, stream);
    default: return launch_march_ng<COUT, 2, 0>(tm, tm2, tmY, p, smem_bytes, grid, stream);
  }
}

static int march_fwd_impl(const void* x, int ldx, const void* w_march, const float* bias, void* y, int ldy,
                          double* stats, int n, int d, int h, int w, int cin, int cout, const FoldExtras& ex,
                          void* stream_, const void* x2 = nullptr, int ldx2 = 0, int cin1 = 0);

extern "C" int b21_conv3d_march_fwd(const void* x, int ldx, const void* w_march, const float* bias, void* y, int ldy,
                                    double* stats, int n, int d, int h, int w, int cin, int cout, void* stream_) {
  FoldExtras ex = {nullptr, nullptr, 0, 0};
  return march_fwd_impl(x, ldx, w_march, bias, y, ldy, stats, n, d, h, w, cin, cout, ex, stream_);
}

// Folded-EvoNorm variant (see fold.cu): per-sample weights `wstride_n` bytes apart (0 = shared), bias table
// [n][27][cout] instead of a bias (or NULL -> plain `bias`), optional swish before the store, optional channel sums.
extern "C" int b21_conv3d_march_fwd_fold(const void* x, int ldx, const void* w_march, long long wstride_n,
                                         const float* bias, const float* bias_table, void* y, int ldy, double* stats,
                                         float* chan_sum, int act, int n, int d, int h, int w, int cin, int cout,
                                         void* stream_) {
  B21_CHECK_ARG(!bias_table || (d >= 2 && h >= 2 && w >= 2), "conv3d_march_fwd_fold: border classes need dims >= 2");
  B21_CHECK_ARG(wstride_n >= 0 && wstride_n % 16 == 0, "conv3d_march_fwd_fold: weight stride must be a multiple of 16 B");
  FoldExtras ex = {bias_table, chan_sum, wstride_n, act};
  return march_fwd_impl(x, ldx, w_march, bias, y, ldy, stats, n, d, h, w, cin, cout, ex, stream_);
}

// Same with the input channels split over TWO tensors (x: channels [0, cin1), x2: channels [cin1, cin)): the consumer
// of a channel concat (equiunet2021.py:310,315,319) reads both producers' DENSE outputs, so neither producer writes a
// 48-byte half of a 96-byte record (partial-sector writes cost 1.3x the DRAM traffic, profiles/r01h_point_full.md).
extern "C" int b21_conv3d_march_fwd_fold2(const void* x, int ldx, int cin1, const void* x2, int ldx2,
                                          const void* w_march, long long wstride_n, const float* bias,
                                          const float* bias_table, void* y, int ldy, double* stats, float* chan_sum,
                                          int act, int n, int d, int h, int w, int cin, int cout, void* stream_) {
  B21_CHECK_ARG(x2, "conv3d_march_fwd_fold2: null second input");
  B21_CHECK_ARG(!bias_table || (d >= 2 && h >= 2 && w >= 2), "conv3d_march_fwd_fold2: border classes need dims >= 2");
  B21_CHECK_ARG(wstride_n >= 0 && wstride_n % 16 == 0, "conv3d_march_fwd_fold2: weight stride must be a multiple of 16 B");
  FoldExtras ex = {bias_table, chan_sum, wstride_n, act};
  return march_fwd_impl(x, ldx, w_march, bias, y, ldy, stats, n, d, h, w, cin, cout, ex, stream_, x2, ldx2, cin1);
}

static int march_fwd_impl(const void* x, int ldx, const void* w_march, const float* bias, void* y, int ldy,
                          double* stats, int n, int d, int h, int w, int cin, int cout, const FoldExtras& ex,
                          void* stream_, const void* x2, int ldx2, int cin1) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (x2) {
    B21_CHECK_ARG(cin1 > 0 && cin1 < cin && cin1 % 8 == 0 && (cin - cin1) % 8 == 0 && ldx >= cin1 &&
                      ldx2 >= cin - cin1 && ldx2 % 8 == 0 && (reinterpret_cast<uintptr_t>(x2) & 15) == 0,
                  "conv3d_march_fwd: bad split input (cin %d = %d + %d, ld %d / %d)", cin, cin1, cin - cin1, ldx, ldx2);
  } else {
    cin1 = cin;
  }
  B21_CHECK_ARG(x && w_march && y, "conv3d_march_fwd: null pointer");
  B21_CHECK_ARG(n > 0 && d > 0 && h > 0 && w > 0, "conv3d_march_fwd: bad shape %d %d %d %d", n, d, h, w);
  B21_CHECK_ARG(b21_conv_march_supported(cin, cout), "conv3d_march_fwd: (cin %d, cout %d) unsupported", cin, cout);
  B21_CHECK_ARG(ldx >= cin1 && ldx % 8 == 0 && ldy >= cout && ldy % 8 == 0, "conv3d_march_fwd: bad ldx %d / ldy %d", ldx, ldy);
  B21_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(w_march) & 15) == 0,
                "conv3d_march_fwd: pointers must be 16-byte aligned");

  ConvMarchParams p;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.wpk = reinterpret_cast<const uint8_t*>(w_march);
  p.bias = bias;
  p.stats = stats;
  p.N = n; p.D = d; p.H = h; p.W = w; p.ldy = ldy;
  p.kc = march_kc(cin);
  p.tilesH = (h + kMTH - 1) / kMTH;
  p.tilesW = (w + kMTW - 1) / kMTW;
  p.stages = march_stages(cin, cout);
  p.ring = 512 / cout < kMMaxRing ? 512 / cout : kMMaxRing;
  p.wbytes = (uint32_t)march_wbytes(cin, cout);
  p.ex = ex;
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("B21_MARCH_VARIANT");
    variant = e ? atoi(e) : 0;
  }
  p.variant = variant;

  // march length: the longest segment that still balances the persistent grid (each segment re-reads 2 halo planes)
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

  CUtensorMap tm, tm2;
  p.split = x2 ? cin1 / 8 : p.kc;
  p.merged = (cin == 8 && ldx == 8 && !x2) ? 1 : 0;
  if (p.merged) {
    const uint64_t dims[5] = {(uint64_t)w * 8, (uint64_t)h, (uint64_t)d, (uint64_t)n, 1};
    const uint64_t str[4] = {uint64_t(w) * 16, uint64_t(h) * w * 16, uint64_t(d) * h * w * 16,
                             uint64_t(n) * d * h * w * 16};
    const uint32_t box[5] = {8 * (uint32_t)kMHW, (uint32_t)kMHH, 1, 1, 1};
    int r = encode_tmap_bf16(&tm, x, 5, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r) return r;
  } else {
    const uint64_t dims[5] = {(uint64_t)cin1, (uint64_t)w, (uint64_t)h, (uint64_t)d, (uint64_t)n};
    const uint64_t str[4] = {uint64_t(ldx) * 2, uint64_t(w) * ldx * 2, uint64_t(h) * w * ldx * 2,
                             uint64_t(d) * h * w * ldx * 2};
    const uint32_t box[5] = {8, (uint32_t)kMHW, (uint32_t)kMHH, 1, 1};
    int r = encode_tmap_bf16(&tm, x, 5, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r) return r;
  }
  tm2 = tm;
  if (x2) {
    const uint64_t dims[5] = {(uint64_t)(cin - cin1), (uint64_t)w, (uint64_t)h, (uint64_t)d, (uint64_t)n};
    const uint64_t str[4] = {uint64_t(ldx2) * 2, uint64_t(w) * ldx2 * 2, uint64_t(h) * w * ldx2 * 2,
                             uint64_t(d) * h * w * ldx2 * 2};
    const uint32_t box[5] = {8, (uint32_t)kMHW, (uint32_t)kMHH, 1, 1};
    int r = encode_tmap_bf16(&tm2, x2, 5, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r) return r;
  }
  if (stats) B21_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * B21_STAT_SLOTS * n * 16, stream));
  const size_t smem_bytes = ((size_t(p.wbytes) + 127) & ~size_t(127)) + size_t(p.stages) * p.kc * kMChunkBytes +
                            march_staging(cin, cout) + 128;
  CUtensorMap tmY;
  {
    const uint64_t dims[5] = {(uint64_t)cout, (uint64_t)w, (uint64_t)h, (uint64_t)d, (uint64_t)n};
    const uint64_t str[4] = {uint64_t(ldy) * 2, uint64_t(w) * ldy * 2, uint64_t(h) * w * ldy * 2, uint64_t(d) * h * w * ldy * 2};
    const uint32_t box[5] = {(uint32_t)cout, (uint32_t)kMTW, 4, 1, 1};
    int r = encode_tmap_bf16(&tmY, y, 5, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r) return r;
  }
  switch (cout) {
    case 16: return launch_march<16>(tm, tm2, tmY, p, smem_bytes, grid, stream);
    case 32: return launch_march<32>(tm, tm2, tmY, p, smem_bytes, grid, stream);
    case 48: return launch_march<48>(tm, tm2, tmY, p, smem_bytes, grid, stream);
    default: return launch_march<64>(tm, tm2, tmY, p, smem_bytes, grid, stream);
  }
}

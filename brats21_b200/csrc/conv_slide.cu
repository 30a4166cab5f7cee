// conv_slide.cu — "sliding-window" implicit-GEMM conv3d (k = 3, dilation 1, stride 1) for the mid-channel layers
// (Cin <= 96) whose 27-tap weights do NOT fit in shared memory (conv_march.cu needs them resident).
//
// The tap-streaming kernel (conv_tap.cu) re-fetches a 128-voxel activation tile AND a weight tile from L2 for every
// tap: 1.5 MB of L2 traffic per 128 x 96 x 96 x 27 output tile, which pins the 96 -> 96 layers at 64^3 (31 % of the
// EquiUNet-ASPP-Evo FLOPs) to the L2 bandwidth (~430 TFLOP/s).  This kernel makes the activations shared-memory
// resident and amortises every streamed weight tile over three output planes:
//
//   * A persistent CTA per SM owns a 16 x 8 (h, w) tile and marches along d in groups of R = 3 output planes, whose
//     three accumulators (N = NT fp32 columns each) live in a TMEM ring.
//   * The work is cut into PHASES phi = 3 g + kd.  Phase phi applies the nine (kh, kw) weight tiles of depth tap kd to
//     the three input halo planes phi-1, phi, phi+1 (relative to the segment), accumulating into the three
//     output planes of group g.  Because R equals the kernel depth, the window of input planes slides by exactly one
//     plane per phase: a ring of five 18 x 10 halo planes (three in use, two in flight) is enough, each plane is
//     fetched from L2 once per segment (TMA, out-of-bounds zero fill = the conv padding), and a (kh, kw) tap is only a
//     different start address of the UMMA shared-memory descriptor (SWIZZLE_NONE interleaved core matrices, as in
//     conv_march.cu).
//   * The weight tiles [Cin x NT] stream through their own TMA ring (cp.async.bulk of the pre-packed UMMA image); each
//     tile feeds 3 x Cin/16 tcgen05.mma (M = 128, N = NT, K = 16).  L2 traffic per 128 outputs drops from 1.5 MB to
//     ~0.22 MB (Cin = Cout = 96).
//   * Warp roles: warp 0 = halo-plane TMA producer, warp 1 = TMEM owner + MMA issuer, warp 2 = weight producer,
//     warps 3..6 = epilogue (tcgen05.ld -> +bias -> GroupNorm/EvoNorm group statistics -> bf16 NDHWC store), which
//     drains group g while the MMAs of group g+1 run (TMEM ring of floor(512 / NT) accumulators).
//
// Replaces torch.nn.Conv3d(k=3, padding=1) at networks/equiunet2020.py:19-25 and networks/equiunet2021.py:198,201
// for 16 <= Cin <= 96, Cout a multiple of the N tile (48 or 96), and is reused for the data gradient with the
// transposed + mirrored packing.
#include "ptx.cuh"
#include "fold.cuh"
#include "host_common.h"
#include "pack.cuh"
#include <stdlib.h>

namespace b21 {

constexpr int kSThreads = 224;
constexpr int kSTH = 16, kSTW = 8;
constexpr int kSHH = kSTH + 2, kSHW = kSTW + 2;
constexpr int kSChunkData = kSHH * kSHW * 16;                  // one 8-channel chunk of a halo plane: 2880 B
constexpr int kSChunkBytes = (kSChunkData + 127) / 128 * 128;  // 2944 B (TMA destinations are 128 B aligned)
constexpr int kSMaxPSlots = 8;
constexpr int kSMaxWStages = 8;
constexpr int kSMaxRing = 10;
constexpr int kSSmemBudget = 225 * 1024;

struct ConvSlideParams {
  __nv_bfloat16* y;
  const uint8_t* wpk;
  const float* bias;
  double* stats;
  int N, D, H, W, ldy;
  int kc, ksteps;
  int tilesH, tilesW, segs, L, ntiles, items;
  int pslots, wstages, wsub, ks_sub;
  int nchunks;  // input-channel chunks of 8 * kc channels (> 1: Cin >= 128, each (group, chunk) pass loads its 5 planes)
  uint32_t tap_bytes, wstage_bytes;
  FoldExtras ex;
  int variant;  // debug (B21_SLIDE_VARIANT): bit0 no plane loads, bit1 no weight loads, bit2 no epilogue math/stores, bit3 one MMA per stage
};

struct SlideItem {
  int n, nt, h0, w0, d0, Lc;
};
__device__ __forceinline__ SlideItem slide_decode(const ConvSlideParams& p, int item) {
  SlideItem it;
  int t = item;
  const int wt = t % p.tilesW; t /= p.tilesW;
  const int ht = t % p.tilesH; t /= p.tilesH;
  it.nt = t % p.ntiles; t /= p.ntiles;
  const int sg = t % p.segs;
  it.n = t / p.segs;
  it.h0 = ht * kSTH;
  it.w0 = wt * kSTW;
  it.d0 = sg * p.L;
  it.Lc = p.D - it.d0 < p.L ? p.D - it.d0 : p.L;
  return it;
}

template <int NT, int GS, int KS0, int KS1>  // N tile, channels per norm group, k-steps of the two weight sub-stages
__global__ void __launch_bounds__(kSThreads, 1)
conv_slide_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const ConvSlideParams p) {
  constexpr uint32_t RING = (512 / NT) < kSMaxRing ? (512 / NT) : kSMaxRing;  // accumulators in TMEM
  constexpr int NG = NT / GS;                                                 // norm groups covered by one N tile
  static_assert(NT % 16 == 0 && NT % GS == 0 && NG <= 8 && RING >= 4, "bad tile");
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t pfull_bar[kSMaxPSlots];
  __shared__ __align__(8) uint64_t pempty_bar[kSMaxPSlots];
  __shared__ __align__(8) uint64_t wfull_bar[kSMaxWStages];
  __shared__ __align__(8) uint64_t wempty_bar[kSMaxWStages];
  __shared__ __align__(8) uint64_t accf_bar[kSMaxRing];
  __shared__ __align__(8) uint64_t acce_bar[kSMaxRing];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_stat[2][16];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const uint32_t w_addr = smem_u32(smem);
  const uint32_t p_addr = w_addr + uint32_t(p.wstages) * p.wstage_bytes;
  const uint32_t plane_bytes = uint32_t(p.kc) * kSChunkBytes;
  // output staging of the TMA-store epilogue: per epilogue warp one dense 32-voxel (4 h x 8 w) x SC-channel bf16 tile;
  // a 96-channel tile goes out as two 48-channel halves through the same buffer
  constexpr int SC = NT == 96 ? 48 : NT;
  constexpr uint32_t kStageTile = 32 * SC * 2;
  const uint32_t y_addr = p_addr + uint32_t(p.pslots) * plane_bytes;
  const uint32_t pfull0 = smem_u32(&pfull_bar[0]), pempty0 = smem_u32(&pempty_bar[0]);
  const uint32_t wfull0 = smem_u32(&wfull_bar[0]), wempty0 = smem_u32(&wempty_bar[0]);
  const uint32_t accf0 = smem_u32(&accf_bar[0]), acce0 = smem_u32(&acce_bar[0]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.pslots; ++s) {
      mbar_init(&pfull_bar[s], 1);
      mbar_init(&pempty_bar[s], 1);
    }
    for (int s = 0; s < p.wstages; ++s) {
      mbar_init(&wfull_bar[s], 1);
      mbar_init(&wempty_bar[s], 1);
    }
    for (uint32_t r = 0; r < RING; ++r) {
      mbar_init(&accf_bar[r], 1);
      mbar_init(&acce_bar[r], 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (threadIdx.x < 32) s_stat[threadIdx.x >> 4][threadIdx.x & 15] = 0.f;
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ halo-plane producer
    if (elect_one()) {
      int slot = 0;
      uint32_t phase = 0;
      const uint32_t tx = uint32_t(p.kc) * kSChunkData;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const SlideItem it = slide_decode(p, item);
        const int G = (it.Lc + 2) / 3;
        // one chunk: the window slides across the groups of the item (3G + 2 planes); several chunks: every
        // (group, chunk) pass loads the 5 planes of its three phases (the accumulators stay in TMEM across the chunks)
        const int npass = p.nchunks > 1 ? G * p.nchunks : 1;
        const int NP = p.nchunks > 1 ? 5 : 3 * G + 2;
        for (int ps = 0; ps < npass; ++ps) {
          const int g = p.nchunks > 1 ? ps / p.nchunks : 0;
          const int ch0 = p.nchunks > 1 ? (ps - g * p.nchunks) * p.kc * 8 : 0;
          for (int jj = 0; jj < NP; ++jj) {
            const int j = 3 * g + jj;
            mbar_wait_sleep_a(pempty0 + 8u * slot, phase ^ 1);
            const uint32_t fb = pfull0 + 8u * slot;
            if (j <= it.Lc + 1 && !(p.variant & 1)) {  // planes beyond the segment halo only feed skipped output planes
              const uint32_t dst = p_addr + uint32_t(slot) * plane_bytes;
              mbar_expect_tx_a(fb, tx);
              for (int c = 0; c < p.kc; ++c)
                tma_load_5d_a(dst + c * kSChunkBytes, &tmX, fb, ch0 + c * 8, it.w0 - 1, it.h0 - 1, it.d0 - 1 + j, it.n);
            } else {
              mbar_arrive_a(fb);
            }
            if (++slot == p.pslots) {
              slot = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ weight producer
    if (elect_one()) {
      int ws = 0;
      uint32_t wph = 0;
      constexpr uint32_t sub_bytes0 = uint32_t(KS0) * NT * 32u;
      constexpr uint32_t sub_bytes1 = uint32_t(KS1) * NT * 32u;
      constexpr int kWsub = KS1 > 0 ? 2 : 1;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const SlideItem it = slide_decode(p, item);
        const int G = (it.Lc + 2) / 3;
        const uint8_t* wt0 = p.wpk + size_t(p.ex.wstride) * it.n + size_t(it.nt) * p.nchunks * 27 * p.tap_bytes;
        for (int gc = 0; gc < G * p.nchunks; ++gc) {
          const uint8_t* wt = wt0 + size_t(gc % p.nchunks) * 27 * p.tap_bytes;
          for (int tap = 0; tap < 27; ++tap) {  // tap = kd * 9 + kh * 3 + kw: phase kd walks taps kd*9 .. kd*9+8
            for (int sub = 0; sub < kWsub; ++sub) {
              mbar_wait_sleep_a(wempty0 + 8u * ws, wph ^ 1);
              const uint32_t nb = sub == 0 ? sub_bytes0 : sub_bytes1;
              const uint32_t fb = wfull0 + 8u * ws;
              if (p.variant & 2) {
                mbar_arrive_a(fb);
              } else {
              mbar_expect_tx_a(fb, nb);
              asm volatile(
                  "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                      w_addr + uint32_t(ws) * p.wstage_bytes),
                  "l"(reinterpret_cast<uint64_t>(wt + size_t(tap) * p.tap_bytes + (sub ? sub_bytes0 : 0u))), "r"(nb),
                  "r"(fb)
                  : "memory");
              }
              if (++ws == p.wstages) {
                ws = 0;
                wph ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected thread)
    // The loop is written for a low instruction count per tcgen05.mma: k-steps per weight stage are compile-time
    // (immediate descriptor offsets), the three A descriptors of a phase and the accumulator columns are computed once
    // per phase, ring cursors are running counters.
    if (elect_one()) {
      constexpr uint32_t lboB = (NT / 8) * 128, sboB = 128;
      constexpr uint32_t lboA = kSChunkBytes, sboA = kSHW * 16;
      const uint64_t dA = umma_smem_desc(0, lboA, sboA, kLayoutNone), dB = umma_smem_desc(0, lboB, sboB, kLayoutNone);
      const uint32_t a_hi = uint32_t(dA >> 32), b_hi = uint32_t(dB >> 32);
      const uint32_t a_lo_base = uint32_t(dA) + (p_addr >> 4);  // address field: bits [0,14), smem < 256 KB: no carry
      const uint32_t b_lo_base = uint32_t(dB) + (w_addr >> 4);
      const uint32_t plane16 = plane_bytes >> 4, wstage16 = p.wstage_bytes >> 4;
      const uint32_t idesc = umma_idesc_bf16(128, NT);
      constexpr uint32_t a_step = 2u * (lboA >> 4), b_step = 2u * (lboB >> 4);  // one k-step (16 channels), 16 B units
      const int pslots = p.pslots, wstages = p.wstages, nchunks = p.nchunks;
      int pwait_slot = 0;        // next plane slot to wait for (planes are consumed strictly in ring order)
      uint32_t pwait_ph = 0;
      int win_slot = 0;          // ring slot of plane j = phi (first plane of the current window)
      int ws = 0;
      uint32_t wph = 0;
      uint32_t acc_slot = 0, acc_par = 0;  // TMEM ring cursor of the first output plane of the current group
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const SlideItem it = slide_decode(p, item);
        const int G = (it.Lc + 2) / 3;
        int jwait = 0;
        for (int g = 0; g < G; ++g) {
          const int nvalid = it.Lc - 3 * g < 3 ? it.Lc - 3 * g : 3;
          uint32_t as1 = acc_slot + 1, ap1 = acc_par;
          if (as1 == RING) { as1 = 0; ap1 ^= 1u; }
          uint32_t as2 = as1 + 1, ap2 = ap1;
          if (as2 == RING) { as2 = 0; ap2 ^= 1u; }
          const uint32_t dcol0 = tmem_base + acc_slot * NT, dcol1 = tmem_base + as1 * NT, dcol2 = tmem_base + as2 * NT;
          for (int ch = 0; ch < nchunks; ++ch) {
          if (nchunks > 1) jwait = 0;  // every (group, chunk) pass has its own 5 planes
          for (int kd = 0; kd < 3; ++kd) {
            const int phi = nchunks > 1 ? kd : 3 * g + kd;
            while (jwait <= phi + 2) {
              mbar_wait_a(pfull0 + 8u * pwait_slot, pwait_ph);
              if (++pwait_slot == pslots) {
                pwait_slot = 0;
                pwait_ph ^= 1;
              }
              ++jwait;
            }
            tc_fence_after();
            int s1 = win_slot + 1;
            if (s1 == pslots) s1 = 0;
            int s2 = s1 + 1;
            if (s2 == pslots) s2 = 0;
            const uint32_t a0 = a_lo_base + uint32_t(win_slot) * plane16, a1 = a_lo_base + uint32_t(s1) * plane16,
                           a2 = a_lo_base + uint32_t(s2) * plane16;
#pragma unroll 1
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const uint32_t tap16 = uint32_t(kh * kSHW + kw);
                const uint32_t fresh = (ch | kd | kh | kw) == 0 ? 1u : 0u;  // first contribution to the group's accumulators
                {
                  mbar_wait_a(wfull0 + 8u * ws, wph);
                  tc_fence_after();
                  const uint32_t b_lo = b_lo_base + uint32_t(ws) * wstage16;
#pragma unroll
                  for (int r = 0; r < 3; ++r) {
                    if (r < nvalid) {
                      const uint32_t as = r == 0 ? acc_slot : (r == 1 ? as1 : as2);
                      const uint32_t ap = r == 0 ? acc_par : (r == 1 ? ap1 : ap2);
                      const uint32_t dcol = r == 0 ? dcol0 : (r == 1 ? dcol1 : dcol2);
                      const uint32_t al = (r == 0 ? a0 : (r == 1 ? a1 : a2)) + tap16;
                      if (fresh) {  // the epilogue must have drained this TMEM slot
                        mbar_wait_a(acce0 + 8u * as, ap ^ 1u);
                        tc_fence_after();
                      }
#pragma unroll
                      for (int ks = 0; ks < KS0; ++ks)
                        umma_bf16(dcol, (uint64_t(a_hi) << 32) | (al + ks * a_step),
                                  (uint64_t(b_hi) << 32) | (b_lo + ks * b_step), idesc, (ks == 0 ? fresh ^ 1u : 1u));
                    }
                  }
                  umma_commit_a(wempty0 + 8u * ws);
                  if (++ws == wstages) {
                    ws = 0;
                    wph ^= 1;
                  }
                }
                if (KS1 > 0) {
                  mbar_wait_a(wfull0 + 8u * ws, wph);
                  tc_fence_after();
                  const uint32_t b_lo = b_lo_base + uint32_t(ws) * wstage16;
#pragma unroll
                  for (int r = 0; r < 3; ++r) {
                    if (r < nvalid) {
                      const uint32_t dcol = r == 0 ? dcol0 : (r == 1 ? dcol1 : dcol2);
                      const uint32_t al = (r == 0 ? a0 : (r == 1 ? a1 : a2)) + tap16 + KS0 * a_step;
#pragma unroll
                      for (int ks = 0; ks < KS1; ++ks)
                        umma_bf16(dcol, (uint64_t(a_hi) << 32) | (al + ks * a_step),
                                  (uint64_t(b_hi) << 32) | (b_lo + ks * b_step), idesc, 1u);
                    }
                  }
                  umma_commit_a(wempty0 + 8u * ws);
                  if (++ws == wstages) {
                    ws = 0;
                    wph ^= 1;
                  }
                }
              }
            }
            umma_commit_a(pempty0 + 8u * win_slot);  // plane j = phi has served its last window
            win_slot = s1;
          }
          if (nchunks > 1) {  // the two trailing planes of the pass
            umma_commit_a(pempty0 + 8u * win_slot);
            if (++win_slot == pslots) win_slot = 0;
            umma_commit_a(pempty0 + 8u * win_slot);
            if (++win_slot == pslots) win_slot = 0;
          }
          }
          for (int r = 0; r < nvalid; ++r) {  // the group's accumulators are complete
            umma_commit_a(accf0 + 8u * acc_slot);
            if (++acc_slot == RING) {
              acc_slot = 0;
              acc_par ^= 1u;
            }
          }
        }
        if (nchunks == 1) {  // the two trailing halo planes (j = 3G, 3G+1)
          umma_commit_a(pempty0 + 8u * win_slot);
          if (++win_slot == pslots) win_slot = 0;
          umma_commit_a(pempty0 + 8u * win_slot);
          if (++win_slot == pslots) win_slot = 0;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (TMEM lane quadrant = warp % 4)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int hh = row >> 3, ww = row & 7;
    const uint32_t tlane = tmem_base + (uint32_t(quad * 32) << 16);
    uint32_t slot = 0, par = 0;
    int buf = 0;
    const uint32_t y_tile = y_addr + uint32_t(quad) * kStageTile;
    uint8_t* y_row = smem + (y_tile - w_addr) + size_t(lane) * (SC * 2);  // this thread's voxel record of the tile
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const SlideItem it = slide_decode(p, item);
      const int h = it.h0 + hh, w = it.w0 + ww;
      const bool valid = (h < p.H) && (w < p.W);
      const int cbase = it.nt * NT;
      float gs[NG], gq[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) gs[g] = gq[g] = 0.f;
      const float* bias = p.bias ? p.bias + cbase : nullptr;
      const int cout_all = p.ntiles * NT;
      const float* trow = p.ex.table ? p.ex.table + (size_t(it.n) * 27 + border_class(h, p.H) * 3 + border_class(w, p.W)) *
                                                          cout_all + cbase
                                     : nullptr;
      constexpr int NCS = (NT + 31) / 32;
      float csum[NCS];  // lane l: running sum of the stored outputs of channels l, 32 + l, ...
#pragma unroll
      for (int b = 0; b < NCS; ++b) csum[b] = 0.f;
      for (int so = 0; so < it.Lc; ++so) {
        mbar_wait_sleep_a(accf0 + 8u * slot, par);
        tc_fence_after();
        float v[NT];
#pragma unroll
        for (int c0 = 0; c0 < NT; c0 += 16) tmem_ld16(tlane + slot * NT + uint32_t(c0), v + c0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(acce0 + 8u * slot);  // drained: the MMA warp may start the next group here
        if (++slot == RING) {
          slot = 0;
          par ^= 1u;
        }
        if (p.variant & 4) continue;
        if (trow) {  // folded input affine: the bias depends on the border class of the output voxel
          if (valid) {
            const float4* tb =
                reinterpret_cast<const float4*>(trow + size_t(border_class(it.d0 + so, p.D)) * 9 * cout_all);
#pragma unroll
            for (int c4 = 0; c4 < NT / 4; ++c4) {
              const float4 t4 = __ldg(tb + c4);
              v[c4 * 4 + 0] += t4.x; v[c4 * 4 + 1] += t4.y; v[c4 * 4 + 2] += t4.z; v[c4 * 4 + 3] += t4.w;
            }
          }
        } else if (bias) {
#pragma unroll
          for (int c = 0; c < NT; ++c) v[c] += __ldg(bias + c);
        }
#pragma unroll
        for (int c = 0; c < NT; ++c) {
          const float val = v[c];
          const float sv = valid ? val : 0.f;
          gs[c / GS] += sv;
          gq[c / GS] = fmaf(sv, sv, gq[c / GS]);
        }
        if (p.ex.act) {  // one branch around the whole unrolled loop: the ex2/rcp chains interleave
#pragma unroll
          for (int c = 0; c < NT; ++c) v[c] = swishf(v[c]);
        }
        uint32_t o[NT / 2];
#pragma unroll
        for (int c = 0; c < NT; c += 2) o[c / 2] = pack_bf16x2(v[c], v[c + 1]);
        // TMA store per warp, plane and SC-channel part (box {SC, 8 w, 4 h}; rows / columns beyond H / W are clipped): a
        // thread-per-voxel st.global.v4 touches up to 32 lines per instruction on the L1 / shared-memory data path the
        // tensor core reads its operands through (conv_march.cu, profiles/r02w_input_conv.md)
#pragma unroll
        for (int part = 0; part < NT / SC; ++part) {
          if (lane == 0) bulk_wait_read0();  // the previous store of this warp has finished reading the staging tile
          __syncwarp();
#pragma unroll
          for (int c0 = 0; c0 < SC; c0 += 8) {
            const int q = (part * SC + c0) / 2;
            *reinterpret_cast<uint4*>(y_row + c0 * 2) = make_uint4(o[q], o[q + 1], o[q + 2], o[q + 3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_5d(&tmY, y_tile, cbase + part * SC, it.w0, it.h0 + quad * 4, it.d0 + so, it.n);
            bulk_commit();
          }
        }
        if (p.ex.chan_sum) {  // SE squeeze: channel sums of what the consumer will read (the rounded values)
#pragma unroll
          for (int b = 0; b < NCS; ++b) {
            float t[32];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const int c = b * 32 + i;
              float2 f = make_float2(0.f, 0.f);
              if (c < NT) f = unpack_bf16x2(o[c / 2]);
              t[i] = valid ? f.x : 0.f;
              t[i + 1] = valid ? f.y : 0.f;
            }
            csum[b] += warp_transpose_sum32(t, lane);
          }
        }
      }
      if (p.ex.chan_sum) {
#pragma unroll
        for (int b = 0; b < NCS; ++b)
          if (b * 32 + lane < NT)
            atomicAdd(p.ex.chan_sum + size_t(it.n) * cout_all + cbase + b * 32 + lane, csum[b]);
      }
      if (p.stats) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const float a = warp_sum(gs[g]), b = warp_sum(gq[g]);
          if (lane == 0) {
            atomicAdd(&s_stat[buf][g * 2], a);
            atomicAdd(&s_stat[buf][g * 2 + 1], b);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
        if (warp == 3 && lane < 2 * NG) {
          const float sv = s_stat[buf][lane];
          s_stat[buf][lane] = 0.f;
          if (sv != 0.f) {
            const int slot_s = item % B21_STAT_SLOTS;
            atomicAdd(p.stats + ((size_t(slot_s) * p.N + it.n) * 8 + it.nt * NG) * 2 + lane, double(sv));
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
// out = UMMA shared-memory image per (N tile, tap): [ntile][tap = kd*9+kh*3+kw][kc][NT/8][8 n][8 k] bf16.
// scale != NULL: per-sample copies (blockIdx.y = sample) with the input channels multiplied by scale[sample][ci].
__global__ void pack_slide_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout_o,
                                         int cin_o, int rows, int kc, int nt, int transpose_flip,
                                         const float* __restrict__ scale = nullptr, int ldscale = 0,
                                         int pack_blocks = 0, BiasTableArgs tab = BiasTableArgs(), int nchunks = 1) {
  if (pack_blocks > 0 && int(blockIdx.x) >= pack_blocks) {  // appended blocks: one bias-table row each
    bias_table_block(tab, blockIdx.x - pack_blocks, blockIdx.y);
    return;
  }
  const int ng = nt / 8;
  const int ntiles = rows / nt;
  const size_t total = size_t(ntiles) * nchunks * 27 * kc * ng * 64;  // kc = 8-channel chunks per input-channel chunk
  const size_t gstride = size_t(pack_blocks > 0 ? pack_blocks : gridDim.x) * blockDim.x;
  out += size_t(blockIdx.y) * total;
  if (scale) scale += size_t(blockIdx.y) * ldscale;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += gstride)
    out[i] = __float2bfloat16_rn(pack_slide_value(w, i, cout_o, cin_o, kc, nt, nchunks, transpose_flip, scale));
}

static inline int slide_nchunks(int cin) { return cin > 96 ? cin / 64 : 1; }
static inline int slide_kc(int cin) { return cin > 96 ? 8 : (cin + 15) / 16 * 2; }  // 8-channel chunks per channel chunk
static inline int slide_nt(int cout) { return cout % 96 == 0 ? 96 : (cout % 48 == 0 ? 48 : (cout % 64 == 0 ? 64 : (cout % 32 == 0 ? 32 : 0))); }

struct SlideCfg {
  int kc, ksteps, nt, wsub, ks_sub, pslots, wstages, nchunks;
  uint32_t tap_bytes, wstage_bytes;
  size_t smem_bytes;
};
static bool slide_config(int cin, int cout, SlideCfg* c) {
  if (cin < 16 || cin % 8 || cout % 8) return false;
  if (cin > 96 && cin % 64) return false;  // channel-chunked mode: whole 64-channel chunks
  c->nt = slide_nt(cout);
  if (c->nt == 0) return false;
  if ((cout / 8) == 0 || c->nt % (cout / 8) != 0) return false;  // whole norm groups per N tile
  c->nchunks = slide_nchunks(cin);
  c->kc = slide_kc(cin);
  c->ksteps = c->kc / 2;
  c->tap_bytes = uint32_t(c->kc) * c->nt * 16u;
  c->wsub = c->ksteps >= 4 ? 2 : 1;
  c->ks_sub = c->wsub == 2 ? (c->ksteps + 1) / 2 : c->ksteps;
  c->wstage_bytes = (uint32_t(c->ks_sub) * c->nt * 32u + 127u) & ~127u;
  const size_t plane = size_t(c->kc) * kSChunkBytes;
  // one chunk: 3 planes in use + 2 in flight; several chunks: a pass needs its third plane while the previous pass still
  // holds three, so one more slot keeps the pipeline full
  c->pslots = c->nchunks > 1 ? 6 : 5;
  // output staging of the four epilogue warps (TMA store): 32 voxels x SC channels each, SC = 48 for the 96-wide tile
  const size_t staging = size_t(4) * 32 * (c->nt == 96 ? 48 : c->nt) * 2;
  if (size_t(kSSmemBudget) < 128 + staging + c->pslots * plane + 2 * size_t(c->wstage_bytes)) return false;
  size_t ws = (size_t(kSSmemBudget) - 128 - staging - c->pslots * plane) / c->wstage_bytes;
  c->wstages = int(ws > size_t(kSMaxWStages) ? size_t(kSMaxWStages) : ws);
  if (c->wstages < (c->wsub == 2 ? 3 : 2)) return false;
  // spare room -> extra plane slots (deeper halo prefetch)
  while (c->pslots < kSMaxPSlots &&
         128 + staging + size_t(c->pslots + 1) * plane + size_t(c->wstages) * c->wstage_bytes <= size_t(kSSmemBudget))
    ++c->pslots;
  c->smem_bytes = 128 + staging + size_t(c->pslots) * plane + size_t(c->wstages) * c->wstage_bytes;
  return true;
}

template <int NT, int GS, int KS0, int KS1>
static int launch_slide(const CUtensorMap& tm, const CUtensorMap& tmY, const ConvSlideParams& p, size_t smem_bytes, int grid,
                        cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    B21_CUDA(cudaFuncSetAttribute(conv_slide_kernel<NT, GS, KS0, KS1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kSSmemBudget));
    attr_set = true;
  }
  conv_slide_kernel<NT, GS, KS0, KS1><<<grid, kSThreads, smem_bytes, stream>>>(tm, tmY, p);
  B21_LAUNCH_CHECK("conv_slide_kernel");
  return B21_OK;
}

// Instantiated shapes: (N tile, channels per norm group, k-steps of weight sub-stage 0 / 1).
//   96 -> 96   EquiUNet-ASPP-Evo encoder2/decoder2 (+ data gradients), EquiUNet encoder2.2
//   48 -> 96   EquiUNet encoder2.1;  96 -> 48  EquiUNet decoder1.1 / decoder2.2;  96 -> 192  EquiUNet encoder3.1
//   64 -> 64   level 3 of the width-16 networks the parity tests run
//   192 -> 192, 384 -> 384, 192 -> 96, 384 -> 192, 768 -> 192 ...: channel-chunked mode (64-channel chunks: KS 2 + 2)
#define B21_SLIDE_SHAPES(X) \
  X(96, 12, 3, 3) X(96, 12, 3, 0) X(48, 6, 3, 3) X(96, 24, 3, 3) X(64, 8, 2, 2) X(96, 12, 2, 2) X(96, 24, 2, 2) X(96, 48, 2, 2)

static bool slide_instantiated(const SlideCfg& c, int cout) {
  const int gs = cout / 8, ks1 = c.ksteps - c.ks_sub;
#define X(nt_, g_, k0, k1) if (c.nt == nt_ && gs == g_ && c.ks_sub == k0 && ks1 == k1) return true;
  B21_SLIDE_SHAPES(X)
#undef X
  return false;
}

}  // namespace b21

using namespace b21;

extern "C" int b21_conv_slide_supported(int cin, int cout) {
  SlideCfg c;
  if (!slide_config(cin, cout, &c)) return 0;
  return slide_instantiated(c, cout) ? 1 : 0;
}

extern "C" long long b21_conv_slide_weight_bytes(int cin, int cout) {
  return (long long)27 * slide_nchunks(cin) * slide_kc(cin) * cout * 16;
}

extern "C" int b21_pack_conv_weight_slide(const float* w, void* packed, int cout, int cin, int transpose_flip,
                                          void* stream) {
  B21_CHECK_ARG(w && packed, "pack_conv_weight_slide: null pointer");
  const int rows = transpose_flip ? cin : cout, inner = transpose_flip ? cout : cin;
  SlideCfg c;
  B21_CHECK_ARG(slide_config(inner, rows, &c), "pack_conv_weight_slide: (cin %d, cout %d) unsupported", inner, rows);
  const size_t total = size_t(27) * c.nchunks * c.kc * rows * 8;
  const int threads = 256;
  const int blocks = int((total + threads - 1) / threads) < 2048 ? int((total + threads - 1) / threads) : 2048;
  pack_slide_weight_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
      w, reinterpret_cast<__nv_bfloat16*>(packed), cout, cin, rows, c.kc, c.nt, transpose_flip, nullptr, 0, 0,
      BiasTableArgs(), c.nchunks);
  B21_LAUNCH_CHECK("pack_slide_weight_kernel");
  return B21_OK;
}

extern "C" int b21_pack_job_slide(const float* w, void* packed, int cout, int cin, int transpose_flip, b21_pack_job* job) {
  B21_CHECK_ARG(w && packed && job, "pack_job_slide: null pointer");
  const int rows = transpose_flip ? cin : cout, inner = transpose_flip ? cout : cin;
  SlideCfg c;
  B21_CHECK_ARG(slide_config(inner, rows, &c), "pack_job_slide: (cin %d, cout %d) unsupported", inner, rows);
  job->w = w; job->out = packed; job->total = (long long)27 * c.nchunks * c.kc * rows * 8;
  job->kind = kPackSlide; job->cout = cout; job->cin = cin; job->tf = transpose_flip;
  job->p0 = rows; job->p1 = c.kc; job->p2 = c.nt; job->p3 = c.nchunks; job->blk0 = 0; job->nblk = 0;
  return B21_OK;
}

extern "C" int b21_pack_conv_weight_slide_fold(const float* w, void* packed, int cout, int cin, const float* scale,
                                               int ldscale, int nsamples, const float* ws, const float* bias,
                                               const float* b_in, float* table, void* stream) {
  B21_CHECK_ARG(w && packed && scale && nsamples > 0 && ldscale >= cin, "pack_conv_weight_slide_fold: bad args");
  B21_CHECK_ARG(!table || (ws && b_in), "pack_conv_weight_slide_fold: the bias table needs ws and B");
  b21_pack_job job;  // source-tiled packing (pack.cuh); padding is not written: the caller zero-fills the buffer once
  int r = b21_pack_job_slide(w, packed, cout, cin, 0, &job);
  if (r) return r;
  BiasTableArgs tab = {ws, bias, b_in, table, ldscale, cout, cin, 27};
  return launch_pack_fold_tile(reinterpret_cast<const PackJob&>(job), scale, ldscale, nsamples, tab, (cudaStream_t)stream);
}

static int slide_fwd_impl(const void* x, int ldx, const void* w_slide, const float* bias, void* y, int ldy,
                          double* stats, int n, int d, int h, int w, int cin, int cout, const FoldExtras& ex,
                          void* stream_);

extern "C" int b21_conv3d_slide_fwd(const void* x, int ldx, const void* w_slide, const float* bias, void* y, int ldy,
                                    double* stats, int n, int d, int h, int w, int cin, int cout, void* stream_) {
  FoldExtras ex = {nullptr, nullptr, 0, 0};
  return slide_fwd_impl(x, ldx, w_slide, bias, y, ldy, stats, n, d, h, w, cin, cout, ex, stream_);
}

// Folded-EvoNorm variant (see fold.cu); arguments as b21_conv3d_march_fwd_fold.
extern "C" int b21_conv3d_slide_fwd_fold(const void* x, int ldx, const void* w_slide, long long wstride_n,
                                         const float* bias, const float* bias_table, void* y, int ldy, double* stats,
                                         float* chan_sum, int act, int n, int d, int h, int w, int cin, int cout,
                                         void* stream_) {
  B21_CHECK_ARG(!bias_table || (d >= 2 && h >= 2 && w >= 2), "conv3d_slide_fwd_fold: border classes need dims >= 2");
  B21_CHECK_ARG(wstride_n >= 0 && wstride_n % 16 == 0, "conv3d_slide_fwd_fold: weight stride must be a multiple of 16 B");
  FoldExtras ex = {bias_table, chan_sum, wstride_n, act};
  return slide_fwd_impl(x, ldx, w_slide, bias, y, ldy, stats, n, d, h, w, cin, cout, ex, stream_);
}

static int slide_fwd_impl(const void* x, int ldx, const void* w_slide, const float* bias, void* y, int ldy,
                          double* stats, int n, int d, int h, int w, int cin, int cout, const FoldExtras& ex,
                          void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B21_CHECK_ARG(x && w_slide && y, "conv3d_slide_fwd: null pointer");
  B21_CHECK_ARG(n > 0 && d > 0 && h > 0 && w > 0, "conv3d_slide_fwd: bad shape %d %d %d %d", n, d, h, w);
  B21_CHECK_ARG(b21_conv_slide_supported(cin, cout), "conv3d_slide_fwd: (cin %d, cout %d) unsupported", cin, cout);
  B21_CHECK_ARG(ldx >= cin && ldx % 8 == 0 && ldy >= cout && ldy % 8 == 0, "conv3d_slide_fwd: bad ldx %d / ldy %d", ldx, ldy);
  B21_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(w_slide) & 15) == 0,
                "conv3d_slide_fwd: pointers must be 16-byte aligned");
  SlideCfg c;
  slide_config(cin, cout, &c);
  ConvSlideParams p;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.wpk = reinterpret_cast<const uint8_t*>(w_slide);
  p.bias = bias;
  p.stats = stats;
  p.N = n; p.D = d; p.H = h; p.W = w; p.ldy = ldy;
  p.kc = c.kc; p.ksteps = c.ksteps;
  p.tilesH = (h + kSTH - 1) / kSTH;
  p.tilesW = (w + kSTW - 1) / kSTW;
  p.ntiles = cout / c.nt;
  p.pslots = c.pslots; p.wstages = c.wstages; p.wsub = c.wsub; p.ks_sub = c.ks_sub;
  p.tap_bytes = c.tap_bytes; p.wstage_bytes = c.wstage_bytes;
  p.nchunks = c.nchunks;
  p.ex = ex;
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("B21_SLIDE_VARIANT");
    variant = e ? atoi(e) : 0;
  }
  p.variant = variant;

  // segment length along d: balance the persistent grid (MMA work is proportional to the planes of an item)
  const int sms = num_sms();
  const long long cols = (long long)n * p.tilesH * p.tilesW * p.ntiles;
  int bestL = d;
  double best = -1.0;
  for (int segs = 1; segs <= d; ++segs) {
    const int L = (d + segs - 1) / segs;
    if (L < 3 && segs > 1) break;
    if ((d + L - 1) / L != segs) continue;
    const long long items = cols * segs;
    const long long rounds = (items + sms - 1) / sms;
    const double eff = double(cols) * d / (double(rounds) * sms * (L + 0.5));
    if (eff > best + 1e-9) {
      best = eff;
      bestL = L;
    }
  }
  p.L = bestL;
  p.segs = (d + p.L - 1) / p.L;
  p.items = int(cols * p.segs);
  const int grid = p.items < sms ? p.items : sms;

  CUtensorMap tm;
  {
    const uint64_t dims[5] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)d, (uint64_t)n};
    const uint64_t str[4] = {uint64_t(ldx) * 2, uint64_t(w) * ldx * 2, uint64_t(h) * w * ldx * 2,
                             uint64_t(d) * h * w * ldx * 2};
    const uint32_t box[5] = {8, (uint32_t)kSHW, (uint32_t)kSHH, 1, 1};
    int r = encode_tmap_bf16(&tm, x, 5, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r) return r;
  }
  CUtensorMap tmY;
  {
    const uint64_t dims[5] = {(uint64_t)cout, (uint64_t)w, (uint64_t)h, (uint64_t)d, (uint64_t)n};
    const uint64_t str[4] = {uint64_t(ldy) * 2, uint64_t(w) * ldy * 2, uint64_t(h) * w * ldy * 2, uint64_t(d) * h * w * ldy * 2};
    const uint32_t box[5] = {(uint32_t)(c.nt == 96 ? 48 : c.nt), (uint32_t)kSTW, 4, 1, 1};
    int r = encode_tmap_bf16(&tmY, y, 5, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r) return r;
  }
  if (stats) B21_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * B21_STAT_SLOTS * n * 16, stream));
  const int gs = cout / 8, ks1 = c.ksteps - c.ks_sub;
#define X(nt_, g_, k0, k1) \
  if (c.nt == nt_ && gs == g_ && c.ks_sub == k0 && ks1 == k1) \
    return launch_slide<nt_, g_, k0, k1>(tm, tmY, p, c.smem_bytes, grid, stream);
  B21_SLIDE_SHAPES(X)
#undef X
  set_error("conv3d_slide_fwd: shape not instantiated");
  return B21_ERR_UNSUPPORTED;
}

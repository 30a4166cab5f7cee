// conv_wgrad_march.cu — plane-marching weight gradient of conv3d (k = 3, dilation 1) for the small-channel layers
// (Cin <= 96, Cout <= 128) that hold most of the training FLOPs.
//
//   dW[co][ci][kd,kh,kw] = sum_v dz[v][co] * x[v + (kd-1, kh-1, kw-1)][ci]
//
// The generic kernel (conv_wgrad.cu) re-fetches a shifted 128-voxel x tile from L2 for each of the 27 taps, which makes
// the 48- and 96-channel layers L2-bandwidth bound (27 x 16 KB per 128 voxels) and leaves the tensor pipe at ~13 %.
// Here, as in conv_march.cu, a persistent CTA owns a 16 x 8 (h, w) tile and marches along d:
//   * every x halo plane (18 x 10 voxels, all channels) and every dz plane (16 x 8) is loaded ONCE by TMA into shared
//     memory in the interleaved core-matrix layout [8-channel chunk][h][w][8 ch]; read as an MN-major operand (channels
//     contiguous, K = voxels) this is the canonical SWIZZLE_NONE layout: a core matrix is 8 consecutive w voxels x 16 B,
//     the next 8 voxels of K are the next h row (LBO = row pitch), the next 8 channels the next chunk (SBO);
//   * a tap (kh, kw) is a different start address of the x descriptor; the three kd taps are the three consecutive x
//     planes of the ring and are folded into the MMA N dimension (N = 3 * Cin_pad, up to 256 columns), so one
//     tcgen05.mma (M = 128 = Cout padded, K = 16 voxels) updates the accumulators of three taps;
//   * the accumulators of the CTA's taps ((kh, kw) group x 3 kd x Cin_pad fp32 columns <= 512) stay in TMEM for the whole
//     voxel range of the CTA (split-K over the grid) and are flushed once with fp32 atomics into the torch-layout
//     gradient [cout][cin][27].
//   * tap pairing (Cout <= 64): with M = 128 accumulator rows and Cout = 48 only 37.5 % of every MMA is useful work and the
//     kernel is bound by tensor issue.  The dz plane is therefore staged TWICE, the second copy shifted by one voxel in
//     w (A_1[u] = dz[u - e_w], TMA zero-fills u_w = 0), directly behind the first one, so that rows [Cout, 2 Cout) of the
//     same M = 128 descriptor hold the shifted copy: sum_u dz[u - e_w][co] x[u + t - 1][ci] = dW[co][ci][t + e_w].  One MMA
//     with the x descriptor at tap (kh, 0) updates taps (kh, 0) and (kh, 1); taps (kh, 2) use a single copy.  6 MMA sets
//     instead of 9, 3 per CTA -> two tap groups instead of three.  The sum over u must reach u_w = W (v_w = W - 1 of the
//     shifted copy), hence one extra, almost empty tile column: tilesW = ceil((W + 1) / 8).
//   * optional (B21_WGM_MULTICAST=1, off by default — it measured slower, see the launch code): the tap groups of one
//     voxel range form a thread-block CLUSTER (3 CTAs for the 48- and 96-channel layers), each CTA loads 1/R of the
//     8-channel chunks and TMA-multicasts them to all R CTAs; a ring slot is released to the producers by
//     tcgen05.commit multicast to the `empty` barrier of every CTA.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = TMEM zero-fill at start + final flush.
// Backward of networks/equiunet2020.py:19-25 and networks/equiunet2021.py:198,201 (learning/engine.py:117 backward()).
#include "ptx.cuh"
#include "host_common.h"
#include <stdlib.h>

namespace b21 {

constexpr int kWMThreads = 192;
constexpr int kWMTH = 16, kWMTW = 8, kWMHH = 18, kWMHW = 10;
constexpr int kWMXChunk = 2944;  // (18 * 10 * 16 B = 2880) rounded up to 128
constexpr int kWMXData = 2880;
constexpr int kWMZChunk = 2048;  // 16 * 8 * 16 B
constexpr int kWMMaxX = 8, kWMMaxZ = 4;
constexpr int kWMSmemBudget = 220 * 1024;

struct WgMarchParams {
  float* dw;
  int N, D, H, W, Cin, Cout;
  int kcx, kcz, ncolx;           // x chunks per plane (even), dz chunks, padded input channels (= 8 * kcx)
  int tilesH, tilesW, segs, L, items;
  int G2, groups, fold;          // (kh, kw) taps per CTA, number of tap groups, max planes per MMA (N <= 256)
  int items_per_split;
  int xslots, zstages;
  int pair;     // 1: dz is staged twice (second copy shifted by one voxel in w) and fills M rows [Cout, 2 Cout): one MMA
                // updates the taps (kh, kw) and (kh, kw + 1) -- see "tap pairing" in the header comment
  int csplit;   // input-channel halves (blockIdx.z): each CTA marches over kcx chunks = Cin_pad / csplit channels of x
  int cluster;  // CTAs per thread-block cluster (tap groups that share one voxel range): x / dz planes are TMA-multicast
  int variant;  // debug (B21_WGM_VARIANT): bit0 skip the final atomics, bit1 one K-step per tap
};

struct WgItem {
  int n, h0, w0, d0, Lc;
};
__device__ __forceinline__ WgItem wg_decode(const WgMarchParams& p, int item) {
  WgItem it;
  int t = item;
  const int wt = t % p.tilesW; t /= p.tilesW;
  const int ht = t % p.tilesH; t /= p.tilesH;
  const int sg = t % p.segs;
  it.n = t / p.segs;
  it.h0 = ht * kWMTH;
  it.w0 = wt * kWMTW;
  it.d0 = sg * p.L;
  it.Lc = p.D - it.d0 < p.L ? p.D - it.d0 : p.L;
  return it;
}

// MMA set q of the launch -> x tap offset (kh, kw') and, when pairing, whether rows [Cout, 2 Cout) are a real tap
struct WgSet {
  int kh, kw;
  bool two;
};
__device__ __forceinline__ WgSet wg_set(const WgMarchParams& p, int q) {
  WgSet s;
  if (p.pair) {
    s.kh = q % 3;
    s.kw = q < 3 ? 0 : 2;
    s.two = q < 3;
  } else {
    s.kh = q / 3;
    s.kw = q % 3;
    s.two = false;
  }
  return s;
}

__device__ __forceinline__ uint32_t wgm_idesc(int N) {  // A and B MN-major, bf16 -> fp32, M = 128
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
}

__global__ void __launch_bounds__(kWMThreads, 1)
conv_wgrad_march_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmZ,
                        const WgMarchParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t xfull[kWMMaxX], xempty[kWMMaxX], zfull[kWMMaxZ], zempty[kWMMaxZ], acc_bar, zero_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  // dz stages first: the M = 128 descriptor of a narrower dz tile reads past its end, into the x ring (valid memory;
  // those accumulator rows are never used)
  const uint32_t z_addr = smem_u32(smem);
  const uint32_t zstage_bytes = uint32_t(p.kcz) * kWMZChunk * (p.pair ? 2u : 1u);
  const uint32_t x_addr = z_addr + uint32_t(p.zstages) * zstage_bytes;
  const uint32_t xplane_bytes = uint32_t(p.kcx) * kWMXChunk;
  const uint32_t xf0 = smem_u32(&xfull[0]), xe0 = smem_u32(&xempty[0]), zf0 = smem_u32(&zfull[0]), ze0 = smem_u32(&zempty[0]);

  const int grp = blockIdx.x;
  const int cbase = int(blockIdx.z) * p.ncolx;  // first input channel of this CTA's share
  const int nsets = p.pair ? 6 : 9;
  const int t9_begin = grp * p.G2;  // first MMA set of this CTA
  const int ntap = nsets - t9_begin < p.G2 ? nsets - t9_begin : p.G2;
  const int i_begin = blockIdx.y * p.items_per_split;
  int i_end = i_begin + p.items_per_split;
  i_end = i_end > p.items ? p.items : i_end;

  const int R = p.cluster;
  const uint32_t crank = R > 1 ? cluster_ctarank() : 0;
  const uint16_t cmask = uint16_t((1u << R) - 1u);
  if (threadIdx.x == 0) {
    // `empty` barriers collect one tcgen05.commit arrival from EVERY CTA of the cluster: a slot may be overwritten
    // (by multicast, in all CTAs at once) only when all of them have finished reading it
    for (int s = 0; s < p.xslots; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xempty[s], R); }
    for (int s = 0; s < p.zstages; ++s) { mbar_init(&zfull[s], 1); mbar_init(&zempty[s], R); }
    mbar_init(&acc_bar, 1);
    mbar_init(&zero_bar, 4);
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
  if (R > 1) cluster_sync_all();  // every CTA's barriers exist before the first remote arrive / multicast write
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: x halo planes and dz planes
    if (elect_one()) {
      int xs = 0, zs = 0;
      uint32_t xph = 0, zph = 0;
      const uint32_t xtx = uint32_t(p.kcx) * kWMXData, ztx = zstage_bytes;
      for (int item = i_begin; item < i_end; ++item) {
        const WgItem it = wg_decode(p, item);
        // order of consumption: x planes j = 0, 1, 2 | dz 0 | x 3 | dz 1 | ...
        for (int j = 0; j < it.Lc + 2; ++j) {
          mbar_wait_sleep_a(xe0 + 8u * xs, xph ^ 1);
          mbar_expect_tx_a(xf0 + 8u * xs, xtx);
          const uint32_t dst = x_addr + uint32_t(xs) * xplane_bytes;
          if (R > 1) {
            for (int c = int(crank); c < p.kcx; c += R)
              tma_load_5d_mc(dst + c * kWMXChunk, &tmX, xf0 + 8u * xs, cbase + c * 8, it.w0 - 1, it.h0 - 1, it.d0 - 1 + j, it.n,
                             cmask);
          } else {
            for (int c = 0; c < p.kcx; ++c)
              tma_load_5d_a(dst + c * kWMXChunk, &tmX, xf0 + 8u * xs, cbase + c * 8, it.w0 - 1, it.h0 - 1, it.d0 - 1 + j, it.n);
          }
          if (++xs == p.xslots) { xs = 0; xph ^= 1; }
          if (j >= 2) {
            const int so = j - 2;
            mbar_wait_sleep_a(ze0 + 8u * zs, zph ^ 1);
            mbar_expect_tx_a(zf0 + 8u * zs, ztx);
            const uint32_t zd = z_addr + uint32_t(zs) * zstage_bytes;
            if (R > 1) {
              for (int c = int(crank); c < p.kcz; c += R)
                tma_load_5d_mc(zd + c * kWMZChunk, &tmZ, zf0 + 8u * zs, c * 8, it.w0, it.h0, it.d0 + so, it.n, cmask);
            } else {
              for (int c = 0; c < p.kcz; ++c)
                tma_load_5d_a(zd + c * kWMZChunk, &tmZ, zf0 + 8u * zs, c * 8, it.w0, it.h0, it.d0 + so, it.n);
              if (p.pair)
                for (int c = 0; c < p.kcz; ++c)
                  tma_load_5d_a(zd + (p.kcz + c) * kWMZChunk, &tmZ, zf0 + 8u * zs, c * 8, it.w0 - 1, it.h0, it.d0 + so, it.n);
            }
            if (++zs == p.zstages) { zs = 0; zph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      // MN-major SWIZZLE_NONE: SBO = next 8 channels (chunk stride), LBO = next 8 voxels of K (next h row)
      const uint64_t dZ = umma_smem_desc(0, kWMTW * 16, kWMZChunk, kLayoutNone);
      const uint64_t dX = umma_smem_desc(0, kWMHW * 16, kWMXChunk, kLayoutNone);
      const uint32_t z_hi = uint32_t(dZ >> 32), x_hi = uint32_t(dX >> 32);
      // (the matrix-descriptor start address is the CTA-local offset: in a cluster launch the shared-window address of
      // a CTA carries its rank in the upper bits, which must not spill into the other descriptor fields)
      const uint32_t z_lo0 = uint32_t(dZ) + ((z_addr & 0x3FFFFu) >> 4), x_lo0 = uint32_t(dX) + ((x_addr & 0x3FFFFu) >> 4);
      const uint32_t zstage16 = zstage_bytes >> 4, xplane16 = xplane_bytes >> 4;
      const uint32_t id1 = wgm_idesc(p.ncolx), id2 = wgm_idesc(2 * p.ncolx), id3 = wgm_idesc(3 * p.ncolx);
      int xw = 0;          // next x slot to wait for
      uint32_t xwph = 0;
      int x_lo_slot = 0;   // slot of x plane j = so (lowest plane of the current window)
      int zs = 0;
      uint32_t zph = 0;
      auto release = [&](uint32_t bar) {  // slot free: tell the producer of every CTA in the cluster
        if (R > 1) umma_commit_mc(bar, cmask);
        else umma_commit_a(bar);
      };
      mbar_wait(&zero_bar, 0);  // accumulators zeroed by the epilogue warps: every MMA accumulates
      tc_fence_after();
      for (int item = i_begin; item < i_end; ++item) {
        const WgItem it = wg_decode(p, item);
        int jw = 0;
        for (int so = 0; so < it.Lc; ++so) {
          while (jw <= so + 2) {
            mbar_wait_a(xf0 + 8u * xw, xwph);
            if (++xw == p.xslots) { xw = 0; xwph ^= 1; }
            ++jw;
          }
          mbar_wait_a(zf0 + 8u * zs, zph);
          tc_fence_after();
          // runs of consecutive ring slots (<= fold planes per MMA, split at the ring wrap)
          const uint32_t za = z_lo0 + uint32_t(zs) * zstage16;
          int j = 0;
          while (j < 3) {
            int sl = x_lo_slot + j;
            if (sl >= p.xslots) sl -= p.xslots;
            int len = 3 - j;
            if (len > p.fold) len = p.fold;
            if (sl + len > p.xslots) len = p.xslots - sl;
            const uint32_t idesc = len == 3 ? id3 : (len == 2 ? id2 : id1);
            const uint32_t xs0 = x_lo0 + uint32_t(sl) * xplane16;
            uint32_t dcol = tmem_base + uint32_t(j * p.ncolx);
            for (int t = 0; t < ntap; ++t, dcol += 3u * p.ncolx) {
              const WgSet st = wg_set(p, t9_begin + t);
              const uint32_t xa = xs0 + uint32_t(st.kh * kWMHW + st.kw);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)  // K = 16 voxels = two h rows of the tile
                if (ks == 0 || !(p.variant & 2))
                umma_bf16(dcol, (uint64_t(z_hi) << 32) | (za + ks * 16u), (uint64_t(x_hi) << 32) | (xa + ks * 20u), idesc, 1u);
            }
            j += len;
          }
          release(xe0 + 8u * x_lo_slot);  // x plane j = so has served its last output plane
          release(ze0 + 8u * zs);
          if (++x_lo_slot == p.xslots) x_lo_slot = 0;
          if (++zs == p.zstages) { zs = 0; zph ^= 1; }
        }
        // the two trailing halo planes of the item
        release(xe0 + 8u * x_lo_slot);
        if (++x_lo_slot == p.xslots) x_lo_slot = 0;
        release(xe0 + 8u * x_lo_slot);
        if (++x_lo_slot == p.xslots) x_lo_slot = 0;
      }
      umma_commit(&acc_bar);
    }
  } else {
    // ------------------------------------------------------------------ zero-fill, then final flush
    const int quad = warp & 3;
    const uint32_t tlane = tmem_base + (uint32_t(quad * 32) << 16);
    const int ncols = ntap * 3 * p.ncolx;
    for (int c0 = 0; c0 < ncols; c0 += 16) tmem_st16_zero(tlane + uint32_t(c0));
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&zero_bar);
    if (i_end > i_begin) {
      const int row = quad * 32 + lane;
      const int shift = p.pair && row >= p.Cout ? 1 : 0;  // rows [Cout, 2 Cout): the copy of dz shifted in w
      const int co = row - shift * p.Cout;
      mbar_wait_sleep(&acc_bar, 0);
      tc_fence_after();
      for (int t = 0; t < ntap; ++t) {
        const WgSet st = wg_set(p, t9_begin + t);
        const bool live = co < p.Cout && (shift == 0 || st.two);
        for (int kd = 0; kd < 3; ++kd) {
          const int tap = kd * 9 + st.kh * 3 + st.kw + shift;
          for (int c0 = 0; c0 < p.ncolx; c0 += 16) {
            float v[16];
            tmem_ld16(tlane + uint32_t((t * 3 + kd) * p.ncolx + c0), v);
            tmem_ld_wait();
            if (live && !(p.variant & 1)) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int ci = cbase + c0 + i;
                if (ci < p.Cin) atomicAdd(p.dw + (size_t(co) * p.Cin + ci) * 27 + tap, v[i]);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (R > 1) cluster_sync_all();  // no CTA may exit while a peer can still arrive on its barriers
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

struct WgMarchCfg {
  int kcx, kcz, ncolx, G2, groups, fold, xslots, zstages, csplit, pair;
  size_t smem_bytes;
};
static bool wgm_config(int cin, int cout, WgMarchCfg* c) {
  if (cin <= 0 || cin % 8 || cin > 96 || cout <= 0 || cout % 8 || cout > 128) return false;
  // 80-96 input channels: one tap's accumulators (3 kd x Cin columns) leave room for a single (kh, kw) tap per CTA, so
  // nine tap groups re-read x and dz from L2.  Two input-channel halves x three taps per CTA halve that traffic.
  static int splitting = -1;  // B21_WGM_CSPLIT=0 switches it off (A/B runs)
  if (splitting < 0) {
    const char* e = getenv("B21_WGM_CSPLIT");
    splitting = e ? atoi(e) : 1;
  }
  c->csplit = splitting && cin >= 80 && cin % 32 == 0 ? 2 : 1;
  cin /= c->csplit;
  c->kcx = (cin + 15) / 16 * 2;
  c->ncolx = c->kcx * 8;
  c->kcz = cout / 8;
  static int pairing = -1;  // B21_WGM_PAIR=0 switches tap pairing off (A/B runs)
  if (pairing < 0) {
    const char* e = getenv("B21_WGM_PAIR");
    pairing = e ? atoi(e) : 1;
  }
  c->pair = pairing && 2 * cout <= 128 ? 1 : 0;
  const int nsets = c->pair ? 6 : 9;
  c->G2 = 512 / (3 * c->ncolx);
  if (c->G2 < 1) return false;
  if (c->G2 > nsets) c->G2 = nsets;
  if (c->pair && c->G2 < 6) c->G2 = c->G2 >= 3 ? 3 : c->G2;  // equal groups: 6 = 1 x 6 = 2 x 3 = 3 x 2 = 6 x 1 sets
  c->groups = (nsets + c->G2 - 1) / c->G2;
  c->fold = 256 / c->ncolx > 3 ? 3 : 256 / c->ncolx;
  const size_t xplane = size_t(c->kcx) * kWMXChunk, zstage = size_t(c->kcz) * kWMZChunk * (c->pair ? 2 : 1);
  c->zstages = 2;
  if (size_t(kWMSmemBudget) < 128 + 2 * zstage + 4 * xplane) return false;
  size_t xs = (size_t(kWMSmemBudget) - 128 - 2 * zstage) / xplane;
  c->xslots = int(xs > size_t(kWMMaxX) ? size_t(kWMMaxX) : xs);
  while (c->zstages < kWMMaxZ && 128 + size_t(c->zstages + 1) * zstage + size_t(c->xslots) * xplane <= size_t(kWMSmemBudget))
    ++c->zstages;
  // the M = 128 dz descriptor reads 16 chunks from the start of a stage: keep that inside the allocation
  c->smem_bytes = 128 + size_t(c->zstages) * zstage + size_t(c->xslots) * xplane;
  if (size_t(c->zstages - 1) * zstage + 16 * size_t(kWMZChunk) > c->smem_bytes - 128) return false;
  return true;
}

}  // namespace b21

using namespace b21;

extern "C" int b21_conv_wgrad_march_supported(int cin, int cout) {
  WgMarchCfg c;
  return wgm_config(cin, cout, &c) ? 1 : 0;
}

extern "C" int b21_conv3d_wgrad_march(const void* x, int ldx, const void* dz, int lddz, float* dw, int n, int d, int h,
                                      int w, int cin, int cin_true, int cout, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B21_CHECK_ARG(x && dz && dw, "conv3d_wgrad_march: null pointer");
  B21_CHECK_ARG(n > 0 && d > 0 && h >= 8 && w >= 8, "conv3d_wgrad_march: bad shape %d %d %d %d", n, d, h, w);
  WgMarchCfg c;
  B21_CHECK_ARG(wgm_config(cin, cout, &c), "conv3d_wgrad_march: (cin %d, cout %d) unsupported", cin, cout);
  B21_CHECK_ARG(cin_true > 0 && cin_true <= cin, "conv3d_wgrad_march: bad cin_true %d", cin_true);
  B21_CHECK_ARG(ldx >= cin && ldx % 8 == 0 && lddz >= cout && lddz % 8 == 0, "conv3d_wgrad_march: bad ldx %d / lddz %d", ldx, lddz);
  B21_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0,
                "conv3d_wgrad_march: pointers must be 16-byte aligned");
  WgMarchParams p;
  p.dw = dw;
  p.N = n; p.D = d; p.H = h; p.W = w; p.Cin = cin_true; p.Cout = cout;
  p.kcx = c.kcx; p.kcz = c.kcz; p.ncolx = c.ncolx;
  p.G2 = c.G2; p.groups = c.groups; p.fold = c.fold;
  p.xslots = c.xslots; p.zstages = c.zstages;
  p.csplit = c.csplit;
  p.pair = c.pair;
  p.tilesH = (h + kWMTH - 1) / kWMTH;
  p.tilesW = (w + p.pair + kWMTW - 1) / kWMTW;  // pairing: the shifted copy needs u_w = W (header comment)
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("B21_WGM_VARIANT");
    variant = e ? atoi(e) : 0;
  }
  p.variant = variant;
  // split-K: groups x splits CTAs ~ 1 per SM; d segments so that every split gets several items
  const int sms = num_sms();
  int splits = sms / (p.groups * p.csplit);
  if (splits < 1) splits = 1;
  const long long cols = (long long)n * p.tilesH * p.tilesW;
  int segs = 1;
  while (cols * segs < 4LL * splits && (d + segs) / (segs + 1) >= 8) ++segs;
  p.L = (d + segs - 1) / segs;
  p.segs = (d + p.L - 1) / p.L;
  p.items = int(cols * p.segs);
  if (splits > p.items) splits = p.items;
  p.items_per_split = (p.items + splits - 1) / splits;
  splits = (p.items + p.items_per_split - 1) / p.items_per_split;

  CUtensorMap tmX, tmZ;
  {
    const uint64_t dims[5] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)d, (uint64_t)n};
    const uint64_t str[4] = {uint64_t(ldx) * 2, uint64_t(w) * ldx * 2, uint64_t(h) * w * ldx * 2, uint64_t(d) * h * w * ldx * 2};
    const uint32_t box[5] = {8, (uint32_t)kWMHW, (uint32_t)kWMHH, 1, 1};
    int r = encode_tmap_bf16(&tmX, x, 5, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r) return r;
  }
  {
    const uint64_t dims[5] = {(uint64_t)cout, (uint64_t)w, (uint64_t)h, (uint64_t)d, (uint64_t)n};
    const uint64_t str[4] = {uint64_t(lddz) * 2, uint64_t(w) * lddz * 2, uint64_t(h) * w * lddz * 2, uint64_t(d) * h * w * lddz * 2};
    const uint32_t box[5] = {8, (uint32_t)kWMTW, (uint32_t)kWMTH, 1, 1};
    int r = encode_tmap_bf16(&tmZ, dz, 5, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r) return r;
  }
  static bool attr_set = false;
  if (!attr_set) {
    B21_CUDA(cudaFuncSetAttribute(conv_wgrad_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWMSmemBudget));
    attr_set = true;
  }
  // cluster = the tap groups of one voxel range (all of them, or three at a time when there are nine)
  static int multicast = -1;
  if (multicast < 0) {
    const char* e = getenv("B21_WGM_MULTICAST");
    // measured (profiles/r02h_wgrad_multicast_ab.md): 48x48 @128^3 0.566 -> 0.565 ms, 96x96 @64^3 0.304 -> 0.423 ms: the
    // kernel is bound by tensor issue (M = 128 rows, cout of them useful), not by L2 -> SM traffic, and coupling the ring
    // slots of three CTAs costs more than the saved loads.  Off by default; B21_WGM_MULTICAST=1 enables it.
    multicast = e ? atoi(e) : 0;
  }
  p.cluster = 1;
  if (multicast && p.groups > 1 && !p.pair) p.cluster = p.groups <= 8 ? p.groups : (p.groups % 3 == 0 ? 3 : 1);
  dim3 grid(p.groups, splits, p.csplit);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kWMThreads);
  cfg.dynamicSmemBytes = c.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B21_CUDA(cudaLaunchKernelEx(&cfg, conv_wgrad_march_kernel, tmX, tmZ, p));
  return B21_OK;
}

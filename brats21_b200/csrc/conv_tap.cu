// conv_tap.cu — generic "tap-streaming" implicit-GEMM conv3d for sm_100a.
//
//   GEMM view: M = output voxels (tile of 128 = td x th x tw box), N = output channels (tile BN <= 256),
//              K = taps x Cin, walked as (tap, 64-channel chunk) K-blocks.
//   A operand: one 5-D TMA box {64ch, tw, th, td, 1} of the NDHWC bf16 activation per K-block, shifted by the
//              tap offset; the zero padding of the convolution is the TMA out-of-bounds zero fill, and so is the
//              channel padding to 64 when Cin is not a multiple of 64 (only ceil(valid/16) UMMAs are issued).
//   B operand: one 3-D TMA box {64ch, BN, 1} of the packed weight [taps][cout_padded][cin_padded] bf16.
//   Both land in 128B-swizzled K-major tiles; tcgen05.mma (M=128, N=BN, K=16) accumulates in TMEM (fp32).
//   Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue
//              (tcgen05.ld -> +bias -> group statistics -> bf16 -> global NDHWC store).
//
// Covers every conv of the reference hot path (k=1 and k=3, dilation 1/2/4/6, any Cin/Cout multiple of 8):
// networks/equiunet2020.py:19-41, networks/equiunet2021.py:165-172,192-222.  The plane-marching kernel in
// conv_march.cu is the faster specialisation for the dil=1, small-channel layers that hold most of the FLOPs.
#include "ptx.cuh"
#include "fold.cuh"
#include "host_common.h"
#include "pack.cuh"

namespace b21 {

constexpr int kTapThreads = 192;
constexpr int kTapABytes = 128 * 128;  // 128 rows x 64 bf16
constexpr int kTapMaxStages = 8;

struct ConvTapParams {
  __nv_bfloat16* y;
  const float* bias;
  double* stats;
  int N, D, H, W, Cin, Cout, ldy, taps, dil;
  int tilesD, tilesH, tilesW;
  int ltw, lth;  // log2 of tile width / height; depth = 128 >> (ltw + lth)
  int BN, chunks, gsize, stages, tmem_cols;
};

__global__ void __launch_bounds__(kTapThreads)
conv_tap_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const ConvTapParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kTapMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kTapMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_stat[16];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = kTapABytes + p.BN * 128;

  int t = blockIdx.x;
  const int wt = t % p.tilesW;
  t /= p.tilesW;
  const int ht = t % p.tilesH;
  t /= p.tilesH;
  const int dt = t % p.tilesD;
  const int n = t / p.tilesD;
  const int tw = 1 << p.ltw, th = 1 << p.lth, td = 128 >> (p.ltw + p.lth);
  const int w0 = wt * tw, h0 = ht * th, d0 = dt * td;
  const int ntile = blockIdx.y;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (threadIdx.x < 16) s_stat[threadIdx.x] = 0.f;
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int numK = p.taps * p.chunks;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < numK; ++kb) {
        const int s = kb % p.stages;
        const uint32_t ph = (kb / p.stages) & 1;
        mbar_wait_sleep(&empty_bar[s], ph ^ 1);
        const int tap = kb / p.chunks, ck = kb - tap * p.chunks;
        int kd = 0, kh = 0, kw = 0;
        if (p.taps == 27) {
          kd = tap / 9 - 1;
          kh = (tap / 3) % 3 - 1;
          kw = tap % 3 - 1;
        }
        uint8_t* sA = smem + size_t(s) * stage_bytes;
        uint8_t* sB = sA + kTapABytes;
        mbar_expect_tx(&full_bar[s], stage_bytes);
        tma_load_5d(sA, &tmA, &full_bar[s], ck * 64, w0 + kw * p.dil, h0 + kh * p.dil, d0 + kd * p.dil, n);
        tma_load_3d(sB, &tmB, &full_bar[s], ck * 64, ntile * p.BN, tap);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.BN);
      for (int kb = 0; kb < numK; ++kb) {
        const int s = kb % p.stages;
        const uint32_t ph = (kb / p.stages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const int ck = kb % p.chunks;
        int nk = (p.Cin - ck * 64 + 15) >> 4;
        nk = nk > 4 ? 4 : nk;
        const uint32_t a0 = smem_u32(smem + size_t(s) * stage_bytes);
        const uint32_t b0 = a0 + kTapABytes;
        for (int k = 0; k < nk; ++k) {
          // K-major 128B-swizzled tiles: 8-row groups 1024 B apart; advancing K by 16 elements = +32 B.
          const uint64_t ad = umma_smem_desc(a0 + k * 32, 16, 1024, kLayoutSw128);
          const uint64_t bd = umma_smem_desc(b0 + k * 32, 16, 1024, kLayoutSw128);
          umma_bf16(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage when these MMAs have read it
      }
      umma_commit(&tmem_full_bar);  // accumulator complete
    }
  } else {
    // ---------------- epilogue: TMEM lane quadrant is fixed by (warp index % 4)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int wx = row & (tw - 1), hy = (row >> p.ltw) & (th - 1), dz = row >> (p.ltw + p.lth);
    const int d = d0 + dz, h = h0 + hy, w = w0 + wx;
    const bool valid = (d < p.D) && (h < p.H) && (w < p.W) && (dz < td);
    const size_t vox = ((size_t(n) * p.D + d) * p.H + h) * p.W + w;
    const int cbase = ntile * p.BN;
    __nv_bfloat16* yrow = p.y + vox * size_t(p.ldy) + cbase;

    mbar_wait_sleep(&tmem_full_bar, 0);
    tc_fence_after();

    float gs = 0.f, gq = 0.f;
    int gcnt = 0, g = p.gsize > 0 ? cbase / p.gsize : 0;
    for (int c0 = 0; c0 < p.BN; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(c0), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = cbase + c0 + j;
        float val = v[j];
        if (c < p.Cout) {
          if (p.bias) val += __ldg(p.bias + c);
          if (p.gsize > 0) {
            const float sv = valid ? val : 0.f;
            gs += sv;
            gq += sv * sv;
            if (++gcnt == p.gsize) {  // warp-uniform
              gs = warp_sum(gs);
              gq = warp_sum(gq);
              if (lane == 0) {
                atomicAdd(&s_stat[g * 2], gs);
                atomicAdd(&s_stat[g * 2 + 1], gq);
              }
              gs = gq = 0.f;
              gcnt = 0;
              ++g;
            }
          }
        }
        v[j] = val;
      }
      if (valid) {
#pragma unroll
        for (int hv = 0; hv < 2; ++hv) {
          if (cbase + c0 + hv * 8 < p.Cout) {
            uint4 o;
            o.x = pack_bf16x2(v[hv * 8 + 0], v[hv * 8 + 1]);
            o.y = pack_bf16x2(v[hv * 8 + 2], v[hv * 8 + 3]);
            o.z = pack_bf16x2(v[hv * 8 + 4], v[hv * 8 + 5]);
            o.w = pack_bf16x2(v[hv * 8 + 6], v[hv * 8 + 7]);
            *reinterpret_cast<uint4*>(yrow + c0 + hv * 8) = o;
          }
        }
      }
    }
    if (p.stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
      if (warp == 2 && lane < 16) {
        const float sv = s_stat[lane];
        if (sv != 0.f) {
          const int slot = blockIdx.x % B21_STAT_SLOTS;
          atomicAdd(p.stats + ((size_t(slot) * p.N + n) * 8) * 2 + lane, double(sv));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ---------------------------------------------------------------------------------------- weight repack
// scale != NULL: per-sample copies (blockIdx.y = sample) with the input channels multiplied by scale[sample][ci].
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout,
                                        int cin, int rows_padded, int inner_padded, int T, int transpose_flip,
                                        const float* __restrict__ scale = nullptr, int ldscale = 0,
                                        int pack_blocks = 0, BiasTableArgs tab = BiasTableArgs()) {
  if (pack_blocks > 0 && int(blockIdx.x) >= pack_blocks) {  // appended blocks: one bias-table row each
    bias_table_block(tab, blockIdx.x - pack_blocks, blockIdx.y);
    return;
  }
  const size_t total = size_t(T) * rows_padded * inner_padded;
  const size_t gstride = size_t(pack_blocks > 0 ? pack_blocks : gridDim.x) * blockDim.x;
  out += size_t(blockIdx.y) * total;
  if (scale) scale += size_t(blockIdx.y) * ldscale;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += gstride)
    out[i] = __float2bfloat16_rn(pack_tap_value(w, i, cout, cin, rows_padded, inner_padded, T, transpose_flip, scale));
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline void conv_tile_n(int cout, int* ntiles, int* bn) {
  *ntiles = (cout + 255) / 256;
  *bn = round_up((cout + *ntiles - 1) / *ntiles, 16);
}
static inline int floor_log2(int v) {
  int l = 0;
  while ((2 << l) <= v) ++l;
  return l;
}

}  // namespace b21

using namespace b21;

extern "C" int b21_conv_cout_padded(int cout) {
  int nt, bn;
  conv_tile_n(cout, &nt, &bn);
  return nt * bn;
}

extern "C" int b21_pack_conv_weight(const float* w, void* packed, int cout, int cin, int cin_padded, int k,
                                    int transpose_flip, void* stream) {
  B21_CHECK_ARG(w && packed, "pack_conv_weight: null pointer");
  B21_CHECK_ARG(k == 1 || k == 3, "pack_conv_weight: k must be 1 or 3 (got %d)", k);
  const int rows = transpose_flip ? cin : cout, inner = transpose_flip ? cout : cin;
  B21_CHECK_ARG(cin_padded >= inner && cin_padded % 8 == 0, "pack_conv_weight: bad inner padding %d for %d",
                cin_padded, inner);
  const int T = k * k * k;
  const int rows_padded = b21_conv_cout_padded(rows);
  const size_t total = size_t(T) * rows_padded * cin_padded;
  const int threads = 256;
  const int blocks = int((total + threads - 1) / threads) < 4096 ? int((total + threads - 1) / threads) : 4096;
  pack_conv_weight_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
      w, reinterpret_cast<__nv_bfloat16*>(packed), cout, cin, rows_padded, cin_padded, T, transpose_flip);
  B21_LAUNCH_CHECK("pack_conv_weight_kernel");
  return B21_OK;
}

extern "C" int b21_pack_job_tap(const float* w, void* packed, int cout, int cin, int cin_padded, int k, int transpose_flip,
                                b21_pack_job* job) {
  B21_CHECK_ARG(w && packed && job, "pack_job_tap: null pointer");
  B21_CHECK_ARG(k == 1 || k == 3, "pack_job_tap: k must be 1 or 3 (got %d)", k);
  const int rows = transpose_flip ? cin : cout, inner = transpose_flip ? cout : cin;
  B21_CHECK_ARG(cin_padded >= inner && cin_padded % 8 == 0, "pack_job_tap: bad inner padding %d for %d", cin_padded, inner);
  const int T = k * k * k, rows_padded = b21_conv_cout_padded(rows);
  job->w = w; job->out = packed; job->total = (long long)T * rows_padded * cin_padded;
  job->kind = kPackTap; job->cout = cout; job->cin = cin; job->tf = transpose_flip;
  job->p0 = rows_padded; job->p1 = cin_padded; job->p2 = T; job->p3 = 0; job->blk0 = 0; job->nblk = 0;
  return B21_OK;
}

// Per-sample folded packing (k = 1 or 3): packed[s] = pack(w * scale[s][ci]), taps*cout_padded*cin_padded bf16 apart.
extern "C" int b21_pack_conv_weight_fold(const float* w, void* packed, int cout, int cin, int cin_padded, int k,
                                         const float* scale, int ldscale, int nsamples, const float* ws,
                                         const float* bias, const float* b_in, float* table, void* stream) {
  B21_CHECK_ARG(w && packed && scale && nsamples > 0 && ldscale >= cin, "pack_conv_weight_fold: bad args");
  B21_CHECK_ARG(!table || (ws && b_in), "pack_conv_weight_fold: the bias table needs ws and B");
  B21_CHECK_ARG(k == 1 || k == 3, "pack_conv_weight_fold: k must be 1 or 3 (got %d)", k);
  B21_CHECK_ARG(cin_padded >= cin && cin_padded % 8 == 0, "pack_conv_weight_fold: bad inner padding");
  // source-tiled packing (pack.cuh): the fp32 weight is read in coalesced 16 x 16 x taps tiles, once per sample; padding
  // rows / channels of the images are not written (the caller zero-fills the buffer once)
  const int ncls = k == 3 ? 27 : 1;
  b21_pack_job job;
  int r = b21_pack_job_tap(w, packed, cout, cin, cin_padded, k, 0, &job);
  if (r) return r;
  BiasTableArgs tab = {ws, bias, b_in, table, ldscale, cout, cin, ncls};
  return launch_pack_fold_tile(reinterpret_cast<const PackJob&>(job), scale, ldscale, nsamples, tab, (cudaStream_t)stream);
}

extern "C" int b21_conv3d_fwd(const void* x, int ldx, const void* w_packed, const float* bias, void* y, int ldy,
                              double* stats, int n, int d, int h, int w, int cin, int cout, int taps, int dil,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B21_CHECK_ARG(x && w_packed && y, "conv3d_fwd: null pointer");
  B21_CHECK_ARG(taps == 1 || taps == 27, "conv3d_fwd: taps must be 1 or 27 (got %d)", taps);
  B21_CHECK_ARG(n > 0 && d > 0 && h > 0 && w > 0, "conv3d_fwd: bad shape %d %d %d %d", n, d, h, w);
  B21_CHECK_ARG(cin > 0 && cin % 8 == 0 && ldx >= cin && ldx % 8 == 0, "conv3d_fwd: cin %d / ldx %d must be multiples of 8", cin, ldx);
  B21_CHECK_ARG(cout > 0 && cout % 8 == 0 && ldy >= cout && ldy % 8 == 0, "conv3d_fwd: cout %d / ldy %d must be multiples of 8", cout, ldy);
  B21_CHECK_ARG(dil >= 1 && dil <= 8, "conv3d_fwd: dilation %d unsupported", dil);
  B21_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
                "conv3d_fwd: pointers must be 16-byte aligned");

  ConvTapParams p;
  int ntiles;
  conv_tile_n(cout, &ntiles, &p.BN);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.bias = bias;
  p.stats = stats;
  p.N = n; p.D = d; p.H = h; p.W = w; p.Cin = cin; p.Cout = cout; p.ldy = ldy; p.taps = taps; p.dil = dil;
  p.ltw = floor_log2(w) < 3 ? floor_log2(w) : 3;
  p.lth = floor_log2(h) < 3 ? floor_log2(h) : 3;
  if (p.ltw + p.lth > 7) p.lth = 7 - p.ltw;
  const int tw = 1 << p.ltw, th = 1 << p.lth, td = 128 >> (p.ltw + p.lth);
  p.tilesW = (w + tw - 1) / tw;
  p.tilesH = (h + th - 1) / th;
  p.tilesD = (d + td - 1) / td;
  p.chunks = (cin + 63) / 64;
  p.gsize = stats ? cout / 8 : 0;
  // Small problems (the 16^3 level of a batch-1 training step: 32 voxel tiles): halve the N tile while the grid covers
  // less than half of the SMs -- 384 -> 384 runs 4 x 96 columns on 128 CTAs instead of 2 x 192 on 64, 384 -> 96 runs
  // 2 x 48 on 64 CTAs instead of 32 (the packed weight rows = ntiles * BN do not change).
  const long long mtiles = (long long)p.tilesW * p.tilesH * p.tilesD * n;
  const int sms = num_sms();
  while (mtiles * ntiles * 2 < sms && p.BN % 32 == 0 && p.BN >= 64 && (!stats || (p.BN / 2) % p.gsize == 0)) {
    p.BN /= 2;
    ntiles *= 2;
  }
  if (stats) B21_CHECK_ARG(cout % 8 == 0 && (p.BN % p.gsize == 0 || ntiles == 1), "conv3d_fwd: stats need whole groups per N tile");
  p.tmem_cols = 32;
  while (p.tmem_cols < p.BN) p.tmem_cols <<= 1;
  const int stage_bytes = kTapABytes + p.BN * 128;
  // two co-resident CTAs per SM when the grid has more CTAs than SMs, otherwise one CTA with a deeper ring (the K loop of
  // a small grid is TMA-latency bound: 162 K-blocks through 3 stages measured 107 us for 21 us of MMAs)
  int stages = ((mtiles * ntiles > sms ? 108 : 198) * 1024) / stage_bytes;
  stages = stages < 2 ? 2 : (stages > kTapMaxStages ? kTapMaxStages : stages);
  p.stages = stages;
  const size_t smem_bytes = size_t(stages) * stage_bytes + 1024;

  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[5] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)d, (uint64_t)n};
    const uint64_t str[4] = {uint64_t(ldx) * 2, uint64_t(w) * ldx * 2, uint64_t(h) * w * ldx * 2,
                             uint64_t(d) * h * w * ldx * 2};
    const uint32_t box[5] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)td, 1};
    int r = encode_tmap_bf16(&tmA, x, 5, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  {
    const int cin_padded = cin;  // packed inner dim == cin (already a multiple of 8)
    const int rows = ntiles * p.BN;
    const uint64_t dims[3] = {(uint64_t)cin_padded, (uint64_t)rows, (uint64_t)taps};
    const uint64_t str[2] = {uint64_t(cin_padded) * 2, uint64_t(rows) * cin_padded * 2};
    const uint32_t box[3] = {64, (uint32_t)p.BN, 1};
    int r = encode_tmap_bf16(&tmB, w_packed, 3, dims, str, box, (int)CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  if (stats) B21_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * B21_STAT_SLOTS * n * 16, stream));

  static bool attr_set = false;
  if (!attr_set) {
    B21_CUDA(cudaFuncSetAttribute(conv_tap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid(unsigned(p.tilesW) * p.tilesH * p.tilesD * n, ntiles);
  conv_tap_kernel<<<grid, kTapThreads, smem_bytes, stream>>>(tmA, tmB, p);
  B21_LAUNCH_CHECK("conv_tap_kernel");
  return B21_OK;
}

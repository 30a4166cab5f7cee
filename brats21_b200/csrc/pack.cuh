// pack.cuh — element functions of the three bf16 weight packings (generic tap layout, plane-march image, sliding-window
// image) and the job record of the batched re-pack (b21_pack_batch: ONE launch re-packs every conv of a network after an
// optimizer step instead of ~90 launches of 2-30 us each).  The per-layout kernels in conv_tap.cu / conv_march.cu /
// conv_slide.cu call the same element functions, so both routes produce identical images.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace b21 {

enum { kPackTap = 0, kPackMarch = 1, kPackSlide = 2, kPackInput = 3 };

// mirrors `b21_pack_job` of include/b21.h (64 bytes)
struct PackJob {
  const float* w;       // fp32 weight [cout][cin][k^3]
  __nv_bfloat16* out;   // packed image
  long long total;      // elements of the image
  int kind, cout, cin, tf;
  int p0, p1, p2, p3;   // tap: rows_padded, inner_padded, taps | march: rows, kc | slide: rows, kc, nt, nchunks | input: cout
  int blk0, nblk;       // block range of the job's GROUP (consecutive jobs with the same source weight share one range:
                        // one block per 16 x 16 (cout, cin) source tile)
};
static_assert(sizeof(PackJob) == 64, "PackJob must match b21_pack_job");

// [taps][rows_padded][inner_padded]; transpose_flip: rows = input channels, inner = output channels, taps mirrored
__device__ __forceinline__ float pack_tap_value(const float* __restrict__ w, size_t i, int cout, int cin, int rows_padded,
                                                int inner_padded, int T, int transpose_flip,
                                                const float* __restrict__ scale) {
  const int ki = int(i % inner_padded);
  const int r = int((i / inner_padded) % rows_padded);
  const int tap = int(i / (size_t(inner_padded) * rows_padded));
  float v = 0.f;
  if (!transpose_flip) {
    if (r < cout && ki < cin) v = w[(size_t(r) * cin + ki) * T + tap] * (scale ? scale[ki] : 1.f);
  } else {
    if (r < cin && ki < cout) v = w[(size_t(ki) * cin + r) * T + (T - 1 - tap)];
  }
  return v;
}

// plane-march image [kh,kw][cin/8][3*rows/8][8][8] (conv_march.cu)
__device__ __forceinline__ float pack_march_value(const float* __restrict__ w, size_t i, int cout_o, int cin_o, int rows,
                                                  int kc, int transpose_flip, const float* __restrict__ scale) {
  const int ng = 3 * rows / 8;
  const int k8 = int(i & 7), n8 = int((i >> 3) & 7);
  size_t t = i >> 6;
  const int g = int(t % ng); t /= ng;
  const int c = int(t % kc);
  const int tap9 = int(t / kc);
  const int n = g * 8 + n8, j = n / rows, ro = n % rows, kd = 2 - j, kh = tap9 / 3, kw = tap9 % 3;
  const int ki = c * 8 + k8;
  float v = 0.f;
  if (!transpose_flip) {
    if (ro < cout_o && ki < cin_o) v = w[(size_t(ro) * cin_o + ki) * 27 + (kd * 9 + kh * 3 + kw)] * (scale ? scale[ki] : 1.f);
  } else {
    // rows = original input channels, inner = original output channels, taps mirrored (data gradient)
    if (ro < cin_o && ki < cout_o) v = w[(size_t(ki) * cin_o + ro) * 27 + ((2 - kd) * 9 + (2 - kh) * 3 + (2 - kw))];
  }
  return v;
}

// sliding-window image [N tile][channel chunk][tap][kc][nt/8][8][8] (conv_slide.cu)
__device__ __forceinline__ float pack_slide_value(const float* __restrict__ w, size_t i, int cout_o, int cin_o, int kc,
                                                  int nt, int nchunks, int transpose_flip,
                                                  const float* __restrict__ scale) {
  const int ng = nt / 8;
  const int k8 = int(i & 7), n8 = int((i >> 3) & 7);
  size_t t = i >> 6;
  const int g = int(t % ng); t /= ng;
  const int c = int(t % kc); t /= kc;
  const int tap = int(t % 27); t /= 27;
  const int chunk = int(t % nchunks);
  const int tile = int(t / nchunks);
  const int ro = tile * nt + g * 8 + n8;
  const int ki = (chunk * kc + c) * 8 + k8;
  float v = 0.f;
  if (!transpose_flip) {
    if (ro < cout_o && ki < cin_o) v = w[(size_t(ro) * cin_o + ki) * 27 + tap] * (scale ? scale[ki] : 1.f);
  } else {  // rows = original input channels, inner = original output channels, taps mirrored (data gradient)
    if (ro < cin_o && ki < cout_o) v = w[(size_t(ki) * cin_o + ro) * 27 + (26 - tap)];
  }
  return v;
}

// input-conv image [6 chunks][3 cout / 8][8 n][8 k], n = (2 - kd) * cout + co, k = (kh * 3 + kw) * 4 + c (conv_input.cu)
__device__ __forceinline__ float pack_input_value(const float* __restrict__ w, size_t i, int cout, int cin_o) {
  const int ng = 3 * cout / 8;
  const int k8 = int(i & 7), n8 = int((i >> 3) & 7);
  size_t t = i >> 6;
  const int g = int(t % ng);
  const int chunk = int(t / ng);
  const int k = chunk * 8 + k8, tap9 = k >> 2, c = k & 3, n = g * 8 + n8, kd = 2 - n / cout, co = n % cout;
  return (tap9 < 9 && c < cin_o) ? w[(size_t(co) * cin_o + c) * 27 + kd * 9 + tap9] : 0.f;
}
__device__ __forceinline__ size_t pack_input_index(int co, int ci, int tap, int cout) {
  const int kd = tap / 9, k = (tap % 9) * 4 + ci, n = (2 - kd) * cout + co;
  return ((size_t(k >> 3) * (3 * cout / 8) + (n >> 3)) * 8 + (n & 7)) * 8 + (k & 7);
}

// Inverse maps: destination index of source element (co, ci, tap) in each image (used by the batched re-pack, which
// walks the SOURCE in coalesced tiles).  r / ki are the image's row / inner channel (swapped for transpose_flip).
__device__ __forceinline__ size_t pack_tap_index(int r, int ki, int tap, int rows_padded, int inner_padded) {
  return (size_t(tap) * rows_padded + r) * inner_padded + ki;
}
__device__ __forceinline__ size_t pack_march_index(int ro, int ki, int kd, int kh, int kw, int rows, int kc) {
  const int ng = 3 * rows / 8, n = (2 - kd) * rows + ro;
  return (((size_t(kh * 3 + kw) * kc + (ki >> 3)) * ng + (n >> 3)) * 8 + (n & 7)) * 8 + (ki & 7);
}
__device__ __forceinline__ size_t pack_slide_index(int ro, int ki, int tap, int kc, int nt, int nchunks) {
  const int ng = nt / 8, tile = ro / nt, rem = ro - tile * nt, q = ki >> 3, chunk = q / kc, c = q - chunk * kc;
  return (((((size_t(tile) * nchunks + chunk) * 27 + tap) * kc + c) * ng + (rem >> 3)) * 8 + (rem & 7)) * 8 + (ki & 7);
}

constexpr int kPT = 16;                // source tile: 16 output channels x 16 input channels x all taps
constexpr int kPitch = kPT * 27 + 1;   // odd row pitch of the shared tile: the transposed walk stays (almost) conflict-free

// One block (256 threads) re-packs source tile `t` of the group jobs[first..last] (same source weight): it reads the
// tile's nco x (nci * taps) contiguous floats once into `tile` (kPT * kPitch floats of shared memory) and writes every
// image of the group in runs of 16 consecutive bf16.  scale != NULL multiplies input channel ci by scale[ci]
// (per-sample folded weights; forward images only).  Padding rows / channels of the images are not touched.
__device__ __forceinline__ void pack_tile_block(const PackJob* jobs, int first, int last, int t,
                                                const float* __restrict__ scale, float* tile) {
  const PackJob& j0 = jobs[first];
  const int cout = j0.cout, cin = j0.cin;
  const int T = j0.kind == kPackTap ? j0.p2 : 27;
  const int cin_tiles = (cin + kPT - 1) / kPT;
  const int co0 = (t / cin_tiles) * kPT, ci0 = (t % cin_tiles) * kPT;
  const int nco = cout - co0 < kPT ? cout - co0 : kPT, nci = cin - ci0 < kPT ? cin - ci0 : kPT;
  // source rows: w[co][ci0 .. ci0 + nci)[0 .. T) is one contiguous run of nci * T floats; a warp per row
  const int run = nci * T, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < nco; r += 8) {
    const float* src = j0.w + (size_t(co0 + r) * cin + ci0) * T;
    for (int k = lane; k < run; k += 32) tile[r * kPitch + k] = src[k];
  }
  __syncthreads();
  // thread (a, b) = image (row, inner channel) within the tile, b fastest: 16 consecutive bf16 per run; it walks the
  // taps with a constant destination stride (every image layout is linear in the tap index for a fixed element)
  const int a = threadIdx.x >> 4, b = threadIdx.x & 15;
  for (int q = first; q <= last; ++q) {
    const PackJob& j = jobs[q];
    const int na = j.tf ? nci : nco, nb = j.tf ? nco : nci;
    if (a >= na || b >= nb) continue;
    const int col = j.tf ? b : a, cil = j.tf ? a : b;  // source (co, ci) within the tile
    const float* src = tile + col * kPitch + cil * T;
    const int r = (j.tf ? ci0 : co0) + a, ki = (j.tf ? co0 : ci0) + b;
    const float sc = scale ? scale[ci0 + cil] : 1.f;
    if (j.kind == kPackTap) {
      const size_t step = size_t(j.p0) * j.p1;
      size_t dst = pack_tap_index(r, ki, 0, j.p0, j.p1);
      for (int tap = 0; tap < T; ++tap, dst += step) j.out[dst] = __float2bfloat16_rn(src[j.tf ? T - 1 - tap : tap] * sc);
    } else if (j.kind == kPackMarch) {
      // index(kd, kh, kw) = index(0, 0, 0) + (kh * 3 + kw) * s9 - kd * sd
      const size_t i0 = pack_march_index(r, ki, 0, 0, 0, j.p0, j.p1);
      const size_t s9 = pack_march_index(r, ki, 0, 0, 1, j.p0, j.p1) - i0, sd = i0 - pack_march_index(r, ki, 1, 0, 0, j.p0, j.p1);
#pragma unroll
      for (int kd = 0; kd < 3; ++kd)
#pragma unroll
        for (int t9 = 0; t9 < 9; ++t9) {
          const int tap = kd * 9 + t9;
          j.out[i0 + t9 * s9 - kd * sd] = __float2bfloat16_rn(src[j.tf ? 26 - tap : tap] * sc);
        }
    } else if (j.kind == kPackInput) {
      if (ki < 4) {
#pragma unroll
        for (int tap = 0; tap < 27; ++tap) j.out[pack_input_index(r, ki, tap, j.p0)] = __float2bfloat16_rn(src[tap] * sc);
      }
    } else {
      const size_t i0 = pack_slide_index(r, ki, 0, j.p1, j.p2, j.p3);
      const size_t st = pack_slide_index(r, ki, 1, j.p1, j.p2, j.p3) - i0;
#pragma unroll
      for (int tap = 0; tap < 27; ++tap) j.out[i0 + tap * st] = __float2bfloat16_rn(src[j.tf ? 26 - tap : tap] * sc);
    }
  }
}

struct BiasTableArgs;
// Per-sample folded packing (conv_tap.cu / conv_march.cu / conv_slide.cu): image s = pack(w * scale[s][ci]) at
// out + s * job.total, plus the bias-table rows when tab.table != NULL; one launch (pack_batch.cu).
int launch_pack_fold_tile(const PackJob& job, const float* scale, int ldscale, int nsamples, const BiasTableArgs& tab,
                          cudaStream_t stream);

}  // namespace b21

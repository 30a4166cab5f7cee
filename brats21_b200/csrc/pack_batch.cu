// pack_batch.cu — every bf16 weight image of a network re-packed by ONE launch (after Ranger2020.step changed the fp32
// master weights in place).  The training step spent 0.41-0.56 ms in ~90 back-to-back packing launches of 2-30 us, most
// of it in stride-27 gathers of the fp32 source (one 32-byte sector per 4 useful bytes).  Here a block owns one
// 16 x 16 (cout, cin) tile of ONE source weight: it reads the tile's 16 x (16 * taps) contiguous floats once into
// shared memory and writes every image made from that weight (generic tap layout, plane-march or sliding-window image,
// forward and transposed: consecutive jobs with the same source form a group) in runs of 16 consecutive bf16.
// Padding rows / channels of the images are never touched: they were zeroed by the first, per-layout packing.
#include "ptx.cuh"
#include "host_common.h"
#include "pack.cuh"

namespace b21 {

static_assert(sizeof(b21_pack_job) == sizeof(PackJob), "b21_pack_job and PackJob must have the same layout");
constexpr int kPT = 16;  // source tile: 16 output channels x 16 input channels x all taps

constexpr int kPitch = kPT * 27 + 1;  // odd row pitch of the shared tile: the transposed walk stays (almost) conflict-free

__global__ void __launch_bounds__(256) pack_batch_kernel(const PackJob* __restrict__ jobs, int njobs) {
  __shared__ float tile[kPT * kPitch];
  __shared__ int s_first, s_last;
  if (threadIdx.x == 0) {
    int lo = 0, hi = njobs - 1;  // jobs are in ascending blk0: last job whose first block is <= this block
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].blk0 <= int(blockIdx.x)) lo = mid;
      else hi = mid - 1;
    }
    int first = lo;
    while (first > 0 && jobs[first - 1].blk0 == jobs[lo].blk0) --first;
    s_first = first;
    s_last = lo;
  }
  __syncthreads();
  const int first = s_first, last = s_last;
  const PackJob j0 = jobs[first];
  const int cout = j0.cout, cin = j0.cin;
  const int T = j0.kind == kPackTap ? j0.p2 : 27;
  const int cin_tiles = (cin + kPT - 1) / kPT;
  const int t = int(blockIdx.x) - j0.blk0;
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
    const PackJob j = jobs[q];
    const int na = j.tf ? nci : nco, nb = j.tf ? nco : nci;
    if (a >= na || b >= nb) continue;
    const int col = j.tf ? b : a, cil = j.tf ? a : b;  // source (co, ci) within the tile
    const float* src = tile + col * kPitch + cil * T;
    const int r = (j.tf ? ci0 : co0) + a, ki = (j.tf ? co0 : ci0) + b;
    if (j.kind == kPackTap) {
      const size_t step = size_t(j.p0) * j.p1;
      size_t dst = pack_tap_index(r, ki, 0, j.p0, j.p1);
      for (int tap = 0; tap < T; ++tap, dst += step) j.out[dst] = __float2bfloat16_rn(src[j.tf ? T - 1 - tap : tap]);
    } else if (j.kind == kPackMarch) {
      // index(kd, kh, kw) = index(0, 0, 0) + (kh * 3 + kw) * s9 - kd * sd
      const size_t i0 = pack_march_index(r, ki, 0, 0, 0, j.p0, j.p1);
      const size_t s9 = pack_march_index(r, ki, 0, 0, 1, j.p0, j.p1) - i0, sd = i0 - pack_march_index(r, ki, 1, 0, 0, j.p0, j.p1);
#pragma unroll
      for (int kd = 0; kd < 3; ++kd)
#pragma unroll
        for (int t9 = 0; t9 < 9; ++t9) {
          const int tap = kd * 9 + t9;
          j.out[i0 + t9 * s9 - kd * sd] = __float2bfloat16_rn(src[j.tf ? 26 - tap : tap]);
        }
    } else {
      const size_t i0 = pack_slide_index(r, ki, 0, j.p1, j.p2, j.p3);
      const size_t st = pack_slide_index(r, ki, 1, j.p1, j.p2, j.p3) - i0;
#pragma unroll
      for (int tap = 0; tap < 27; ++tap) j.out[i0 + tap * st] = __float2bfloat16_rn(src[j.tf ? 26 - tap : tap]);
    }
  }
}

}  // namespace b21

using namespace b21;

extern "C" int b21_pack_batch(const b21_pack_job* jobs_dev, int njobs, int total_blocks, void* stream) {
  B21_CHECK_ARG(jobs_dev && njobs > 0 && total_blocks > 0, "pack_batch: bad arguments");
  pack_batch_kernel<<<total_blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const PackJob*>(jobs_dev), njobs);
  B21_LAUNCH_CHECK("pack_batch_kernel");
  return B21_OK;
}

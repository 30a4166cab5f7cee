// pack_batch.cu — every bf16 weight image of a network re-packed by ONE launch (after Ranger2020.step changed the fp32
// master weights in place).  The training step spent 0.41-0.56 ms in ~90 back-to-back packing launches of 2-30 us, most
// of it in stride-27 gathers of the fp32 source (one 32-byte sector per 4 useful bytes).  Here a block owns one
// 16 x 16 (cout, cin) tile of ONE source weight: it reads the tile's 16 x (16 * taps) contiguous floats once into
// shared memory and writes every image made from that weight (generic tap layout, plane-march or sliding-window image,
// forward and transposed: consecutive jobs with the same source form a group) in runs of 16 consecutive bf16.
// Padding rows / channels of the images are never touched: they were zeroed by the first, per-layout packing.
#include "ptx.cuh"
#include "host_common.h"
#include "fold.cuh"
#include "pack.cuh"

namespace b21 {

static_assert(sizeof(b21_pack_job) == sizeof(PackJob), "b21_pack_job and PackJob must have the same layout");
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
  pack_tile_block(jobs, s_first, s_last, int(blockIdx.x) - jobs[s_first].blk0, nullptr, tile);
}

// grid (source tiles + bias-table rows, samples): per-sample folded image, sample s at job.out + s * job.total
__global__ void __launch_bounds__(256) pack_fold_tile_kernel(PackJob job, const float* __restrict__ scale, int ldscale,
                                                             int ntiles, BiasTableArgs tab) {
  __shared__ float tile[kPT * kPitch];
  if (int(blockIdx.x) >= ntiles) {  // appended blocks: one bias-table row each
    bias_table_block(tab, blockIdx.x - ntiles, blockIdx.y);
    return;
  }
  job.out += size_t(blockIdx.y) * size_t(job.total);
  pack_tile_block(&job, 0, 0, blockIdx.x, scale + size_t(blockIdx.y) * ldscale, tile);
}

int launch_pack_fold_tile(const PackJob& job, const float* scale, int ldscale, int nsamples, const BiasTableArgs& tab,
                          cudaStream_t stream) {
  const int ntiles = ((job.cout + kPT - 1) / kPT) * ((job.cin + kPT - 1) / kPT);
  pack_fold_tile_kernel<<<dim3(ntiles + (tab.table ? tab.ncls : 0), nsamples), 256, 0, stream>>>(job, scale, ldscale,
                                                                                             ntiles, tab);
  B21_LAUNCH_CHECK("pack_fold_tile_kernel");
  return B21_OK;
}

}  // namespace b21

using namespace b21;

extern "C" int b21_pack_batch(const b21_pack_job* jobs_dev, int njobs, int total_blocks, void* stream) {
  B21_CHECK_ARG(jobs_dev && njobs > 0 && total_blocks > 0, "pack_batch: bad arguments");
  pack_batch_kernel<<<total_blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const PackJob*>(jobs_dev), njobs);
  B21_LAUNCH_CHECK("pack_batch_kernel");
  return B21_OK;
}

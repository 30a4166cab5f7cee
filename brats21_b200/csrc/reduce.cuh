// reduce.cuh — block-level reduction of per-thread channel-chunk partial sums.
//
// The streaming kernels (norm_apply, norm_bwd_*, head_conv_bwd, channel_stats) give every thread ONE 8-channel chunk
// ck = (global thread index) % chunks and NV running sums per channel.  atomicAdd(float) on shared memory compiles to
// a compare-and-swap spin loop (SASS ATOMS.CAST.SPIN); with 256 / chunks threads of a block adding to the same address
// the spin loops cost 40-80 us per launch (measured, 24-channel tensors worst).  The lanes of a warp that serve the same
// chunk are `chunks` lanes apart, so a shuffle tree over the distances chunks, 2 chunks, 4 chunks, ... leaves the warp's
// total of chunk (c0 + r) % chunks in lane r < chunks; only those lanes touch shared memory (8 warps per address).  No
// scratch memory: these kernels run next to persistent tensor kernels that own ~190 KB of the SM's shared memory.
#pragma once

namespace b21 {

// red: NV * C floats of shared memory (zeroed here); red[q * C + c] is valid for every thread on return.
// Must be reached by every thread of the block.
template <int NV>
__device__ __forceinline__ void block_chunk_reduce(float (&v)[NV][8], int chunks, int C, float* red) {
  for (int i = threadIdx.x; i < NV * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int ck = int(((long long)blockIdx.x * blockDim.x + threadIdx.x) % chunks);
  bool owner = true;
  if (chunks < 32) {
    for (int d = chunks; d < 32; d <<= 1) {
      const bool has = lane + d < 32;
#pragma unroll
      for (int q = 0; q < NV; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float o = __shfl_down_sync(0xffffffffu, v[q][j], d);
          if (has) v[q][j] += o;
        }
    }
    owner = lane < chunks;
  }
  if (owner) {
#pragma unroll
    for (int q = 0; q < NV; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&red[q * C + ck * 8 + j], v[q][j]);
  }
  __syncthreads();
}

}  // namespace b21

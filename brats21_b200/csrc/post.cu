// post.cu — label-map post-processing on the GPU: the step right after the hot path, which the
// reference runs on the CPU through skimage / scipy:
//   * KeepLargestConnectedComponent / get_largest_component (utils/transforms.py:209-230, 579-600):
//     skimage.morphology.label(mask) with full (26-) connectivity, keep the components larger than a threshold
//     (or only the largest one);
//   * ReplaceWithClosestValue / replace_w_closest_value_{2d,3d} (utils/transforms.py:233-268, 603-647): labels with
//     at most `thresh` voxels are replaced, slice by slice along one axis, by the value of the nearest voxel of the
//     same slice that carries a kept label (scipy griddata(method="nearest"));
//   * ConvertToMultiChannelBasedOnBratsClasses (MONAI): label map -> (TC, WT, ET) channels.
// All integer work: bit-exact against the oracle (oracle/prepost.py).  HBM traffic is a few passes over a 9 MB
// label volume + a 36 MB int32 parent array, i.e. launch-latency bound; the kernels are plain coalesced grid-stride
// loops sized in multiples of the SM count.
#include "host_common.h"
#include <stdint.h>

namespace b21 {

static inline int post_grid(long long items, int threads) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  return int(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

// ------------------------------------------------------------------------------------ connected components
// Label-equivalence union-find (Komura / Playne-Hawick): parent[] always points to a smaller linear index, so the
// root of a component is its first voxel in raster order — the same order skimage numbers its labels in.
__device__ __forceinline__ int cc_find(const int* parent, int x) {
  int p = __ldcg(parent + x);
  while (p != x) {
    x = p;
    p = __ldcg(parent + x);
  }
  return x;
}

__device__ __forceinline__ void cc_union(int* parent, int a, int b) {
  bool done;
  do {
    a = cc_find(parent, a);
    b = cc_find(parent, b);
    if (a < b) {
      const int old = atomicMin(parent + b, a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(parent + a, b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

__global__ void __launch_bounds__(256) cc_init_kernel(const uint8_t* __restrict__ label, int* __restrict__ parent,
                                                      int* __restrict__ sizes, long long nvox) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (long long)gridDim.x * blockDim.x) {
    parent[i] = label[i] ? int(i) : -1;
    sizes[i] = 0;
  }
}

// every foreground voxel merges with its 13 raster-order predecessors of the 26-neighbourhood
__global__ void __launch_bounds__(256) cc_merge_kernel(const uint8_t* __restrict__ label, int* __restrict__ parent,
                                                       int D, int H, int W) {
  const long long nvox = (long long)D * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (long long)gridDim.x * blockDim.x) {
    if (!label[i]) continue;
    const int w = int(i % W);
    const int h = int((i / W) % H);
    const int d = int(i / ((long long)W * H));
#pragma unroll
    for (int k = 0; k < 13; ++k) {
      // k 0..8: plane d-1, all (oh, ow); k 9..11: row h-1 of plane d; k 12: the voxel to the left
      const int od = k < 9 ? -1 : 0;
      const int oh = k < 9 ? k / 3 - 1 : (k < 12 ? -1 : 0);
      const int ow = k < 9 ? k % 3 - 1 : (k < 12 ? k - 10 : -1);
      const int d2 = d + od, h2 = h + oh, w2 = w + ow;
      if (d2 < 0 || h2 < 0 || h2 >= H || w2 < 0 || w2 >= W) continue;
      const long long j = ((long long)d2 * H + h2) * W + w2;
      if (label[j]) cc_union(parent, int(i), int(j));
    }
  }
}

// path compression + component sizes (atomicAdd on the root; warp-aggregated when neighbours share the root)
__global__ void __launch_bounds__(256) cc_count_kernel(int* __restrict__ parent, int* __restrict__ sizes, long long nvox) {
  for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < nvox; i0 += (long long)gridDim.x * blockDim.x) {
    const long long i = i0 + threadIdx.x;
    int root = -1;
    if (i < nvox && parent[i] >= 0) {
      root = cc_find(parent, int(i));
      parent[i] = root;
    }
    // aggregate equal roots inside the warp
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, root);
    if (root >= 0 && (__ffs(peers) - 1) == int(threadIdx.x & 31)) atomicAdd(sizes + root, __popc(peers));
  }
}

// best = max over roots of (size << 32 | ~root): the largest component, the first one in raster order on ties
__global__ void __launch_bounds__(256) cc_best_kernel(const int* __restrict__ parent, const int* __restrict__ sizes,
                                                      unsigned long long* __restrict__ best, long long nvox) {
  unsigned long long loc = 0ull;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (long long)gridDim.x * blockDim.x) {
    if (parent[i] == int(i)) {
      const unsigned long long key = ((unsigned long long)(unsigned)sizes[i] << 32) | (unsigned)(~unsigned(i));
      loc = key > loc ? key : loc;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, loc, o);
    loc = other > loc ? other : loc;
  }
  if ((threadIdx.x & 31) == 0 && loc) atomicMax(best, loc);
}

__global__ void __launch_bounds__(256) cc_apply_kernel(uint8_t* __restrict__ label, const int* __restrict__ parent,
                                                       const int* __restrict__ sizes,
                                                       const unsigned long long* __restrict__ best, int threshold,
                                                       long long nvox) {
  int keep_root = -2;
  if (threshold < 0) {
    const unsigned long long b = *best;
    keep_root = b ? int(~unsigned(b & 0xffffffffull)) : -2;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (long long)gridDim.x * blockDim.x) {
    const int root = parent[i];
    if (root < 0) continue;
    const bool keep = threshold < 0 ? (root == keep_root) : (sizes[root] > threshold);
    if (!keep) label[i] = 0;
  }
}

// ------------------------------------------------------------------------------------ rare-label replacement
// work layout (int32): [0..255] histogram, [256] number of masked voxels, [257] "active" flag,
//                      [258 .. 258+cap) masked voxel indices, [258+cap .. 258+2cap) their new values
constexpr int kRareHdr = 258;

__global__ void __launch_bounds__(256) rare_hist_kernel(const uint8_t* __restrict__ label, int* __restrict__ work,
                                                        long long nvox) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  int cur = -1, run = 0;  // run-length per thread: label maps are mostly background
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (long long)gridDim.x * blockDim.x) {
    const int v = label[i];
    if (v != cur) {
      if (run) atomicAdd(&h[cur], run);
      cur = v;
      run = 0;
    }
    ++run;
  }
  if (run) atomicAdd(&h[cur], run);
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(work + threadIdx.x, h[threadIdx.x]);
}

// `values_to_replace.any()` (utils/transforms.py:265): nothing happens unless a NON-ZERO label is rare
__global__ void rare_flag_kernel(int* __restrict__ work, int thresh) {
  const int c = work[threadIdx.x];
  const int rare_nonzero = (threadIdx.x > 0 && c > 0 && c <= thresh) ? 1 : 0;
  const int any = __syncthreads_or(rare_nonzero);
  if (threadIdx.x == 0) work[257] = any;
}

__global__ void __launch_bounds__(256) rare_collect_kernel(const uint8_t* __restrict__ label, int* __restrict__ work,
                                                           int cap, int thresh, long long nvox) {
  if (!work[257]) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (long long)gridDim.x * blockDim.x) {
    if (work[label[i]] <= thresh) {
      const int k = atomicAdd(work + 256, 1);
      if (k < cap) work[kRareHdr + k] = int(i);
    }
  }
}

// one CTA per masked voxel: nearest (squared euclidean distance in the slice, then smallest row-major slice index)
// voxel of the same slice whose label is kept.  dims (n0, n1, n2) with strides (s0, s1, s2): the slice runs over
// axes (a, b) = the two axes other than `axis`, in their original order.
__global__ void __launch_bounds__(256) rare_nearest_kernel(const uint8_t* __restrict__ label, int* __restrict__ work,
                                                           int cap, int thresh, int n0, int n1, int n2, int axis) {
  const int count = min(work[256], cap);
  if (int(blockIdx.x) >= count) return;
  const int idx = work[kRareHdr + blockIdx.x];
  const int c2 = idx % n2, c1 = (idx / n2) % n1, c0 = idx / (n2 * n1);
  const int dims[3] = {n0, n1, n2};
  const long long strides[3] = {(long long)n1 * n2, (long long)n2, 1};
  const int coord[3] = {c0, c1, c2};
  const int a = axis == 0 ? 1 : 0, b = axis == 2 ? 1 : 2;
  const int na = dims[a], nb = dims[b];
  const long long base = (long long)coord[axis] * strides[axis];
  const int pa = coord[a], pb = coord[b];
  unsigned long long bestk = ~0ull;
  for (int t = threadIdx.x; t < na * nb; t += blockDim.x) {
    const int qa = t / nb, qb = t % nb;
    const uint8_t v = label[base + qa * strides[a] + qb * strides[b]];
    if (work[v] <= thresh) continue;  // masked itself
    const long long da = qa - pa, db = qb - pb;
    const unsigned long long key = ((unsigned long long)(da * da + db * db) << 32) | (unsigned)t;
    bestk = key < bestk ? key : bestk;
  }
  __shared__ unsigned long long red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, bestk, o);
    bestk = other < bestk ? other : bestk;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = bestk;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) bestk = red[k] < bestk ? red[k] : bestk;
    int newv = label[idx];  // a slice without any kept voxel keeps its value
    if (bestk != ~0ull) {
      const int t = int(bestk & 0xffffffffull);
      newv = label[base + (t / nb) * strides[a] + (t % nb) * strides[b]];
    }
    work[kRareHdr + cap + blockIdx.x] = newv;
  }
}

__global__ void rare_apply_kernel(uint8_t* __restrict__ label, const int* __restrict__ work, int cap) {
  const int count = min(work[256], cap);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x)
    label[work[kRareHdr + k]] = uint8_t(work[kRareHdr + cap + k]);
}

// ------------------------------------------------------------------------------------ label map -> channels
__global__ void __launch_bounds__(256) labels_to_channels_kernel(const uint8_t* __restrict__ label,
                                                                 uint8_t* __restrict__ onehot, long long nvox) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (long long)gridDim.x * blockDim.x) {
    const uint8_t v = label[i];
    onehot[i] = (v == 1 || v == 4) ? 1 : 0;            // TC
    onehot[nvox + i] = (v == 1 || v == 2 || v == 4) ? 1 : 0;  // WT
    onehot[2 * nvox + i] = (v == 4) ? 1 : 0;            // ET
  }
}

}  // namespace b21

using namespace b21;

extern "C" long long b21_keep_components_workspace_bytes(long long nvox) { return 8 * nvox + 16; }

extern "C" int b21_keep_components(uint8_t* label, void* work, int d, int h, int w, int threshold, void* stream) {
  B21_CHECK_ARG(label && work, "keep_components: null pointer");
  B21_CHECK_ARG(d > 0 && h > 0 && w > 0, "keep_components: bad dims");
  const long long nvox = (long long)d * h * w;
  B21_CHECK_ARG(nvox < (1ll << 31), "keep_components: volume too large for int32 voxel indices");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* best = reinterpret_cast<unsigned long long*>(work);
  int* parent = reinterpret_cast<int*>(best + 2);
  int* sizes = parent + nvox;
  B21_CUDA(cudaMemsetAsync(best, 0, 16, st));
  const int grid = post_grid(nvox, 256);
  cc_init_kernel<<<grid, 256, 0, st>>>(label, parent, sizes, nvox);
  cc_merge_kernel<<<grid, 256, 0, st>>>(label, parent, d, h, w);
  cc_count_kernel<<<grid, 256, 0, st>>>(parent, sizes, nvox);
  if (threshold < 0) cc_best_kernel<<<grid, 256, 0, st>>>(parent, sizes, best, nvox);
  cc_apply_kernel<<<grid, 256, 0, st>>>(label, parent, sizes, best, threshold, nvox);
  B21_LAUNCH_CHECK("keep_components kernels");
  return B21_OK;
}

extern "C" long long b21_replace_rare_workspace_bytes(int thresh) {
  const long long cap = 256ll * (thresh > 0 ? thresh : 1);
  return 4 * (kRareHdr + 2 * cap);
}

extern "C" int b21_replace_rare_labels(uint8_t* label, void* work, int n0, int n1, int n2, int thresh, int axis,
                                       void* stream) {
  B21_CHECK_ARG(label && work, "replace_rare_labels: null pointer");
  B21_CHECK_ARG(n0 > 0 && n1 > 0 && n2 > 0 && axis >= 0 && axis <= 2, "replace_rare_labels: bad dims / axis");
  B21_CHECK_ARG(thresh >= 0 && thresh <= 4096, "replace_rare_labels: thresh must be 0..4096");
  const long long nvox = (long long)n0 * n1 * n2;
  B21_CHECK_ARG(nvox < (1ll << 31), "replace_rare_labels: volume too large for int32 voxel indices");
  cudaStream_t st = (cudaStream_t)stream;
  const int cap = 256 * (thresh > 0 ? thresh : 1);
  int* wk = reinterpret_cast<int*>(work);
  B21_CUDA(cudaMemsetAsync(wk, 0, sizeof(int) * kRareHdr, st));
  const int grid = post_grid(nvox, 256);
  rare_hist_kernel<<<grid, 256, 0, st>>>(label, wk, nvox);
  rare_flag_kernel<<<1, 256, 0, st>>>(wk, thresh);
  rare_collect_kernel<<<grid, 256, 0, st>>>(label, wk, cap, thresh, nvox);
  rare_nearest_kernel<<<cap, 256, 0, st>>>(label, wk, cap, thresh, n0, n1, n2, axis);
  rare_apply_kernel<<<(cap + 255) / 256, 256, 0, st>>>(label, wk, cap);
  B21_LAUNCH_CHECK("replace_rare_labels kernels");
  return B21_OK;
}

extern "C" int b21_labels_to_channels(const uint8_t* label, uint8_t* onehot, long long nvox, void* stream) {
  B21_CHECK_ARG(label && onehot && nvox > 0, "labels_to_channels: null pointer");
  labels_to_channels_kernel<<<post_grid(nvox, 256), 256, 0, (cudaStream_t)stream>>>(label, onehot, nvox);
  B21_LAUNCH_CHECK("labels_to_channels_kernel");
  return B21_OK;
}

// pre.cu — input side of the hot path on the GPU: what the reference's data transforms do on the CPU
// between NIfTI decode and the first convolution (src/definer.py:561-567):
//   * CropForegroundd(source_key="img") (MONAI 0.6.0: bounding box of `img > 0` over ANY channel, margin 0);
//   * NormalizeIntensity(nonzero=True, channel_wise=True[, remove_outliers]) (utils/transforms.py:328-406):
//     per channel, over the voxels != 0: (x - mean) / std (population std; std == 0 -> 1), zeros stay zero,
//     optional clip to +-outliers_value on the non-zero voxels;
//   * shape_to_divisible(k = 8) (utils/transforms.py:483-512): zero padding, ceil(p / 2) before, floor(p / 2) after.
// Three streaming passes over a 143 MB fp32 volume (4 x 240 x 240 x 155): HBM-bound, coalesced float4-free
// grid-stride loops (rows of 155 floats are not 16 B aligned), warp-shuffle + shared-memory reductions, fp64 sums.
#include "host_common.h"
#include <stdint.h>

namespace b21 {

static inline int pre_grid(long long items, int threads) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  return int(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

__global__ void bbox_init_kernel(int* __restrict__ bbox, int D, int H, int W) {
  if (threadIdx.x == 0) {
    bbox[0] = D; bbox[1] = H; bbox[2] = W;
    bbox[3] = 0; bbox[4] = 0; bbox[5] = 0;
  }
}

// bbox = {min d, min h, min w, max d + 1, max h + 1, max w + 1} of the voxels with any channel > 0
__global__ void __launch_bounds__(256) foreground_bbox_kernel(const float* __restrict__ img, int C, int D, int H, int W,
                                                              int* __restrict__ bbox) {
  const long long nvox = (long long)D * H * W;
  int lo[3] = {D, H, W}, hi[3] = {0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (long long)gridDim.x * blockDim.x) {
    bool fg = false;
    for (int c = 0; c < C; ++c) fg |= __ldg(img + c * nvox + i) > 0.f;
    if (fg) {
      const int w = int(i % W), h = int((i / W) % H), d = int(i / ((long long)W * H));
      lo[0] = min(lo[0], d); lo[1] = min(lo[1], h); lo[2] = min(lo[2], w);
      hi[0] = max(hi[0], d + 1); hi[1] = max(hi[1], h + 1); hi[2] = max(hi[2], w + 1);
    }
  }
  __shared__ int s[6];
  if (threadIdx.x < 3) s[threadIdx.x] = 1 << 30;
  else if (threadIdx.x < 6) s[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int a = lo[k], b = hi[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&s[k], a);
      atomicMax(&s[3 + k], b);
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) atomicMin(bbox + threadIdx.x, s[threadIdx.x]);
  else if (threadIdx.x < 6) atomicMax(bbox + threadIdx.x, s[threadIdx.x]);
}

// stats[c] = {count, sum, sum of squares} (fp64) over the voxels != 0 of channel c inside bbox; grid.y = channel
__global__ void __launch_bounds__(256) nonzero_stats_kernel(const float* __restrict__ img, int D, int H, int W,
                                                            const int* __restrict__ bbox, double* __restrict__ stats) {
  const int c = blockIdx.y;
  const int d0 = bbox[0], h0 = bbox[1], w0 = bbox[2];
  const int bd = bbox[3] - d0, bh = bbox[4] - h0, bw = bbox[5] - w0;
  double cnt = 0.0, sum = 0.0, sq = 0.0;
  if (bd > 0 && bh > 0 && bw > 0) {
    const float* src = img + (long long)c * D * H * W;
    const long long rows = (long long)bd * bh;
    // one warp per (d, h) row of the box: coalesced along w
    const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (long long r = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows;
         r += (long long)gridDim.x * warps_per_block) {
      const int d = d0 + int(r / bh), h = h0 + int(r % bh);
      const float* row = src + ((long long)d * H + h) * W + w0;
      float fs = 0.f, fq = 0.f;
      int fc = 0;
      for (int w = lane; w < bw; w += 32) {
        const float v = __ldg(row + w);
        if (v != 0.f) {
          ++fc;
          fs += v;
          fq = fmaf(v, v, fq);
        }
      }
      cnt += fc; sum += fs; sq += fq;  // at most ceil(bw / 32) fp32 terms per partial, then fp64
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  __shared__ double red[8][3];
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5][0] = cnt; red[threadIdx.x >> 5][1] = sum; red[threadIdx.x >> 5][2] = sq;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int k = 0; k < int(blockDim.x >> 5); ++k) t += red[k][threadIdx.x];
    if (t != 0.0) atomicAdd(stats + c * 3 + threadIdx.x, t);
  }
}

// out[c][od][oh][ow] (zero padded) = normalised crop; one thread per output element, coalesced along w
__global__ void __launch_bounds__(256) normalize_crop_pad_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                                 int C, int D, int H, int W, const int* __restrict__ bbox,
                                                                 const double* __restrict__ stats, int OD, int OH, int OW,
                                                                 int pd, int ph, int pw, float clip) {
  const int d0 = bbox[0], h0 = bbox[1], w0 = bbox[2];
  const int bd = bbox[3] - d0, bh = bbox[4] - h0, bw = bbox[5] - w0;
  __shared__ float s_mean[16], s_sd[16];
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    const double n = stats[c * 3];
    const double mean = n > 0.0 ? stats[c * 3 + 1] / n : 0.0;
    double var = n > 0.0 ? stats[c * 3 + 2] / n - mean * mean : 0.0;
    var = var > 0.0 ? var : 0.0;
    float sd = float(sqrt(var));
    s_mean[c] = float(mean);
    s_sd[c] = sd == 0.f ? 1.f : sd;
  }
  __syncthreads();
  const long long per = (long long)OD * OH * OW, total = per * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i / per);
    long long v = i - (long long)c * per;
    const int ow = int(v % OW); v /= OW;
    const int oh = int(v % OH);
    const int od = int(v / OH);
    const int d = od - pd, h = oh - ph, w = ow - pw;
    float r = 0.f;
    if (d >= 0 && d < bd && h >= 0 && h < bh && w >= 0 && w < bw) {
      const float x = __ldg(img + (((long long)c * D + d0 + d) * H + h0 + h) * W + w0 + w);
      if (x != 0.f) {
        r = (x - s_mean[c]) / s_sd[c];
        if (clip > 0.f) r = fminf(fmaxf(r, -clip), clip);
      }
    }
    out[i] = r;
  }
}

}  // namespace b21

using namespace b21;

extern "C" int b21_foreground_bbox(const float* img, int c, int d, int h, int w, int* bbox, void* stream) {
  B21_CHECK_ARG(img && bbox, "foreground_bbox: null pointer");
  B21_CHECK_ARG(c >= 1 && d > 0 && h > 0 && w > 0, "foreground_bbox: bad dims");
  cudaStream_t st = (cudaStream_t)stream;
  bbox_init_kernel<<<1, 32, 0, st>>>(bbox, d, h, w);
  foreground_bbox_kernel<<<pre_grid((long long)d * h * w, 256), 256, 0, st>>>(img, c, d, h, w, bbox);
  B21_LAUNCH_CHECK("foreground_bbox_kernel");
  return B21_OK;
}

extern "C" int b21_nonzero_stats(const float* img, int c, int d, int h, int w, const int* bbox, double* stats,
                                 void* stream) {
  B21_CHECK_ARG(img && bbox && stats, "nonzero_stats: null pointer");
  B21_CHECK_ARG(c >= 1 && d > 0 && h > 0 && w > 0, "nonzero_stats: bad dims");
  cudaStream_t st = (cudaStream_t)stream;
  B21_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 3 * c, st));
  dim3 grid(pre_grid((long long)d * h, 8), c);
  nonzero_stats_kernel<<<grid, 256, 0, st>>>(img, d, h, w, bbox, stats);
  B21_LAUNCH_CHECK("nonzero_stats_kernel");
  return B21_OK;
}

extern "C" int b21_normalize_crop_pad(const float* img, float* out, int c, int d, int h, int w, const int* bbox,
                                      const double* stats, int od, int oh, int ow, int pad_d, int pad_h, int pad_w,
                                      float clip, void* stream) {
  B21_CHECK_ARG(img && out && bbox && stats, "normalize_crop_pad: null pointer");
  B21_CHECK_ARG(c >= 1 && c <= 16 && od > 0 && oh > 0 && ow > 0 && pad_d >= 0 && pad_h >= 0 && pad_w >= 0,
                "normalize_crop_pad: bad dims");
  const long long total = (long long)c * od * oh * ow;
  normalize_crop_pad_kernel<<<pre_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(img, out, c, d, h, w, bbox, stats, od,
                                                                                   oh, ow, pad_d, pad_h, pad_w, clip);
  B21_LAUNCH_CHECK("normalize_crop_pad_kernel");
  return B21_OK;
}

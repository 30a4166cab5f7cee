// ranger.cu — fused multi-tensor Ranger2020 step (RAdam + Lookahead), one launch for all parameter tensors.
// Replaces the per-parameter Python loop of learning/optimizer.py:136-255 (about 20 tiny kernels per tensor).
// fp32 parameters, gradients and state; the scalars that depend only on the step count (rectification term,
// step size, look-ahead phase) are computed on the host exactly as the reference does (optimizer.py:205-217).
#include "ptx.cuh"
#include "host_common.h"

namespace b21 {

constexpr int kRangerChunk = 16384;

// table: int64 [ntensors][6] = {param, grad, exp_avg, exp_avg_sq, slow, numel}; chunks: int32 [nchunks][2] = {tensor, offset}
__global__ void __launch_bounds__(256) ranger_step_kernel(const long long* __restrict__ table, const int* __restrict__ chunks,
                                                          float gscale, float lr_step, float beta1, float beta2,
                                                          float eps, float wd, int rectified, int lookahead, float alpha,
                                                          const float* __restrict__ dyn) {
  if (dyn) {  // step-dependent scalars from device memory: the launch can be replayed from a CUDA graph
    lr_step = dyn[0];
    rectified = dyn[1] != 0.f;
    lookahead = dyn[2] != 0.f;
    gscale = dyn[3];
  }
  const int ti = chunks[blockIdx.x * 2], off = chunks[blockIdx.x * 2 + 1];
  const long long* row = table + size_t(ti) * 6;
  float* p = reinterpret_cast<float*>(row[0]) + off;
  const float* g = reinterpret_cast<const float*>(row[1]) + off;
  float* m = reinterpret_cast<float*>(row[2]) + off;
  float* v = reinterpret_cast<float*>(row[3]) + off;
  float* s = reinterpret_cast<float*>(row[4]) + off;
  long long n = row[5] - off;
  n = n > kRangerChunk ? kRangerChunk : n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float gr = g[i] * gscale;
    const float pv = p[i];
    const float vv = beta2 * v[i] + (1.f - beta2) * gr * gr;   // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
    float mm = beta1 * m[i] + (1.f - beta1) * gr;              // exp_avg.mul_(b1).add_(g, 1-b1)
    float upd;
    if (rectified) {
      upd = mm / (sqrtf(vv) + eps);
      if (wd != 0.f) upd += wd * pv;
    } else {
      // G_grad aliases exp_avg in the reference (optimizer.py:229-233): the decay term lands in the moving average
      if (wd != 0.f) mm += wd * pv;
      upd = mm;
    }
    float pn = pv - lr_step * upd;
    if (lookahead) {  // every k-th step: slow += alpha (fast - slow); fast = slow   (optimizer.py:245-251)
      const float sl = s[i] + alpha * (pn - s[i]);
      s[i] = sl;
      pn = sl;
    }
    v[i] = vv;
    m[i] = mm;
    p[i] = pn;
  }
}

}  // namespace b21

using namespace b21;

extern "C" int b21_ranger_chunk(void) { return kRangerChunk; }

extern "C" int b21_ranger_step(const long long* table, const int* chunks, int nchunks, float gscale, float lr,
                               float step_size, float beta1, float beta2, float eps, float weight_decay,
                               int rectified, int lookahead, float alpha, const float* dyn, void* stream) {
  B21_CHECK_ARG(table && chunks && nchunks > 0, "ranger_step: empty parameter table");
  ranger_step_kernel<<<nchunks, 256, 0, (cudaStream_t)stream>>>(table, chunks, gscale, step_size * lr, beta1, beta2, eps,
                                                                weight_decay, rectified, lookahead, alpha, dyn);
  B21_LAUNCH_CHECK("ranger_step_kernel");
  return B21_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Gradient centralisation (centralized_gradient, learning/optimizer.py:11-20; gc_loc = True: applied to the raw
// gradient before the moment updates, optimizer.py:187-188): every dim-0 slice ("row") of a qualifying gradient
// tensor has its mean subtracted in place.  rows: int64 [nrows][2] = {pointer to the row, row length}; one CTA per row.
namespace b21 {
__global__ void __launch_bounds__(256) grad_centralize_kernel(const long long* __restrict__ rows) {
  float* g = reinterpret_cast<float*>(rows[size_t(blockIdx.x) * 2]);
  const int n = int(rows[size_t(blockIdx.x) * 2 + 1]);
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float red[8];
  __shared__ float mean_s;
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    mean_s = t / float(n);
  }
  __syncthreads();
  const float m = mean_s;
  for (int i = threadIdx.x; i < n; i += blockDim.x) g[i] -= m;
}
}  // namespace b21

extern "C" int b21_grad_centralize(const long long* rows, int nrows, void* stream) {
  B21_CHECK_ARG(rows && nrows > 0, "grad_centralize: empty row table");
  b21::grad_centralize_kernel<<<nrows, 256, 0, (cudaStream_t)stream>>>(rows);
  B21_LAUNCH_CHECK("grad_centralize_kernel");
  return B21_OK;
}

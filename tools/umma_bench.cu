// umma_bench.cu — micro-benchmark: sustained cycles per tcgen05.mma (M=128, K=16, bf16) as a function of the
// shared-memory operand layout and N.  Design input for the conv kernels (DESIGN.md "UMMA operand layouts").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench tools/umma_bench.cu && ./umma_bench
#include "../brats21_b200/csrc/ptx.cuh"
#include <stdio.h>
#include <stdlib.h>
using namespace b21;

struct Cfg {
  int layoutA, layoutB;      // 0 none, 2 sw128, 4 sw64, 6 sw32
  uint32_t lboA, sboA, kadvA, offA;  // bytes; kadv = start-address advance per K step; offA = start offset
  uint32_t lboB, sboB, kadvB;
  int N, iters;
};

__global__ void __launch_bounds__(128, 1) bench(Cfg c, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_s, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_s;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem) + c.offA, b0 = smem_u32(smem + 64 * 1024);
    const uint64_t dA = umma_smem_desc(0, c.lboA, c.sboA, c.layoutA), dB = umma_smem_desc(0, c.lboB, c.sboB, c.layoutB);
    const uint32_t idesc = umma_idesc_bf16(128, c.N);
    uint64_t ad[4], bd[4];
    for (int k = 0; k < 4; ++k) {
      ad[k] = dA | uint64_t(((a0 + k * c.kadvA) & 0x3FFFF) >> 4);
      bd[k] = dB | uint64_t(((b0 + k * c.kadvB) & 0x3FFFF) >> 4);
    }
    // warm-up
    for (int k = 0; k < 4; ++k) umma_bf16(tm, ad[k], bd[k], idesc, k ? 1u : 0u);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t0 = clock64();
    for (int i = 0; i < c.iters; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tm + ((i >> 2) & 1) * 256, ad[k], bd[k], idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    const long long t1 = clock64();
    out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

static double run(const char* name, Cfg c) {
  unsigned long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  bench<<<148, 128, 200 * 1024>>>(c, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-44s N=%3d  ERROR %s\n", name, c.N, cudaGetErrorString(e));
    exit(1);
  }
  unsigned long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  double avg = 0, mx = 0;
  for (int i = 0; i < 148; ++i) {
    avg += double(h[i]);
    mx = h[i] > mx ? double(h[i]) : mx;
  }
  avg /= 148.0 * c.iters;
  mx /= c.iters;
  printf("%-44s N=%3d  cycles/MMA avg %.1f max %.1f   (floor %d)  -> %.0f%% of tensor peak\n", name, c.N, avg, mx,
         c.N / 2, 100.0 * (c.N / 2) / avg);
  return avg;
}

int main() {
  const int iters = 4096;
  const int Ns[] = {48, 96, 144, 192, 256};
  for (int N : Ns) {
    // SW128 K-major both (the tap kernel's layout): rows 128 B, SBO 1024, K advance 32 B
    run("A sw128 / B sw128", Cfg{2, 2, 16, 1024, 32, 0, 16, 1024, 32, N, iters});
    run("A sw128 +128B row shift / B sw128", Cfg{2, 2, 16, 1024, 32, 128, 16, 1024, 32, N, iters});
    run("A sw128 SBO 1280 (halo rows) / B sw128", Cfg{2, 2, 16, 1280, 32, 0, 16, 1024, 32, N, iters});
    run("A sw64 / B sw64", Cfg{4, 4, 16, 512, 32, 0, 16, 512, 32, N, iters});
    run("A sw64 SBO 640 / B sw64", Cfg{4, 4, 16, 640, 32, 0, 16, 512, 32, N, iters});
    run("A sw32 / B sw32", Cfg{6, 6, 16, 256, 0, 0, 16, 256, 0, N, iters});
    run("A sw32 SBO 320 / B sw32", Cfg{6, 6, 16, 320, 0, 0, 16, 256, 0, N, iters});
    // no swizzle ("interleaved"): 8 rows x 16 B core matrices; LBO = K-chunk stride, SBO = 8-row group stride
    run("A none SBO 128 / B none SBO 128", Cfg{0, 0, 4096, 128, 8192, 0, 8192, 128, 16384, N, iters});
    run("A none SBO 160 / B none SBO 128", Cfg{0, 0, 4096, 160, 8192, 0, 8192, 128, 16384, N, iters});
    run("A none SBO 256 / B none SBO 128", Cfg{0, 0, 8192, 256, 16384, 0, 8192, 128, 16384, N, iters});
    run("A sw128 / B none SBO 128", Cfg{2, 0, 16, 1024, 32, 0, 8192, 128, 16384, N, iters});
    run("A none SBO 128 / B sw128", Cfg{0, 2, 4096, 128, 8192, 0, 16, 1024, 32, N, iters});
    printf("\n");
  }
  return 0;
}

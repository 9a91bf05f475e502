// Throughput of the warp-level mma.sync.m16n8k16 (f16 inputs, f32 accumulate) on this GPU: issue-slot and pipe cost of the
// legacy tensor path the step kernel uses for its masked sums.   nvcc -arch=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int ILP, int FMA_PER>
__global__ void __launch_bounds__(64, 14) k(float *out, int iters, uint32_t seed) {
  float c[ILP][4] = {};
  uint32_t a[4] = {seed, seed + 1, seed + 2, seed + 3}, b[2] = {seed + 4, seed + 5};
  float f[8];
  for (int i = 0; i < 8; i++) f[i] = seed * 1e-9f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int q = 0; q < ILP; q++) {
      mma16816(c[q], a, b);
#pragma unroll
      for (int z = 0; z < FMA_PER; z++) f[(q * FMA_PER + z) & 7] = fmaf(f[(q * FMA_PER + z) & 7], 1.0001f, 0.5f);
    }
  }
  float s = 0;
  for (int q = 0; q < ILP; q++) s += c[q][0] + c[q][1] + c[q][2] + c[q][3];
  for (int i = 0; i < 8; i++) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int FMA_PER>
void run(const char *name) {
  int dev = 0, sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 14, iters = 4096;
  float *out;
  cudaMalloc(&out, grid * 64 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP, FMA_PER><<<grid, 64>>>(out, 16, 0);
  cudaEventRecord(e0);
  k<ILP, FMA_PER><<<grid, 64>>>(out, iters, 0);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = (double)grid * 2 * iters * ILP;
  printf("%-28s %8.3f ms  %7.1f G mma/s  %6.1f TFLOP/s  %6.2f SM-cycles per mma (at 1.965 GHz)\n", name, ms, mmas / ms * 1e-6,
         mmas * 4096 * 2 / ms * 1e-9, ms * 1e-3 * 1.965e9 * sms / mmas);
  cudaFree(out);
}
int main() {
  run<4, 0>("mma only, 4 independent");
  run<8, 0>("mma only, 8 independent");
  run<4, 8>("4 mma + 32 ffma per trip");
  run<4, 16>("4 mma + 64 ffma per trip");
  run<1, 32>("1 mma + 32 ffma per trip");
  return 0;
}

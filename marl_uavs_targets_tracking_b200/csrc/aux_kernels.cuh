// aux_kernels.cuh -- reset / begin-episode / random-policy / statistics kernels.
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// reset / begin-episode / random-policy / statistics kernels
// ------------------------------------------------------------------------------------------------
// pre-step observation: lists empty -> -1 blocks, self part x/dc, y/dc, a/Na (src/agent/uav.py:156-190)
__device__ __forceinline__ void write_initial_obs(float *ob, double x, double y, int a, double dc, int na) {
#pragma unroll
  for (int k = 0; k < 9; k++) ob[k] = -1.f;
  ob[9] = (float)(x / dc);
  ob[10] = (float)(y / dc);
  ob[11] = (float)((double)a / (double)na);
}

// Environment.reset (src/environment.py:45-107) with Philox draws instead of Python's `random`.
__global__ void uavsim_reset_kernel(const KParams P, const UavSimBuffers B, uint64_t seed) {
  const int64_t total_u = P.E * P.n, total_t = P.E * P.m;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total_u + total_t; k += stride) {
    if (k < total_u) {
      const int64_t e = k / P.n;
      const int i = (int)(k - e * P.n);
      const uint32_t g = (uint32_t)(P.env_id_offset + e);
      const Philox4 r = philox4x32_10((uint32_t)i, UAVSIM_RNG_UAV_RESET, g, 0u, seed);
      // x_i = i * x_max / (n_uav + 1), i = 1..n ; y = y_max / 2  (environment.py:105-107)
      const double x = (double)(i + 1) * P.x_max / (double)(P.n + 1);
      const double y = P.y_max / 2;
      const double h = -PI_D + (PI_D - (-PI_D)) * philox_u53(r.v[0], r.v[1]);  // random.uniform(-pi, pi)
      const int a = (int)philox_below(r.v[2], (uint32_t)P.na);                   // random.randint(0, na-1)
      B.ux[k] = x; B.uy[k] = y; B.uh[k] = h; B.ua[k] = a;
      write_initial_obs(B.obs + k * 12, x, y, a, P.dc, P.na);
    } else {
      const int64_t kt = k - total_u;
      const int64_t e = kt / P.m;
      const int t = (int)(kt - e * P.m);
      const uint32_t g = (uint32_t)(P.env_id_offset + e);
      const Philox4 r = philox4x32_10((uint32_t)t, UAVSIM_RNG_TGT_POS, g, 0u, seed);
      const Philox4 r2 = philox4x32_10((uint32_t)t, UAVSIM_RNG_TGT_HEAD, g, 0u, seed);
      B.tx[kt] = 0 + (P.x_max - 0) * philox_u53(r.v[0], r.v[1]);   // random.uniform(0, x_max)
      B.ty[kt] = 0 + (P.y_max - 0) * philox_u53(r.v[2], r.v[3]);
      B.th[kt] = -PI_D + (PI_D - (-PI_D)) * philox_u53(r2.v[0], r2.v[1]);
      // r2.v[2..3] is the unused a0 draw (environment.py:81)
    }
  }
}

__global__ void uavsim_initial_obs_kernel(const KParams P, const UavSimBuffers B) {
  const int64_t total_u = P.E * P.n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total_u; k += stride)
    write_initial_obs(B.obs + k * 12, B.ux[k], B.uy[k], B.ua[k], P.dc, P.na);
}

__global__ void uavsim_random_actions_kernel(const KParams P, int32_t *__restrict__ actions, uint64_t seed,
                                             uint32_t step) {
  const int64_t total_u = P.E * P.n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total_u; k += stride) {
    const int64_t e = k / P.n;
    const int i = (int)(k - e * P.n);
    const Philox4 r = philox4x32_10((uint32_t)i, UAVSIM_RNG_ACTION, (uint32_t)(P.env_id_offset + e), step, seed);
    actions[k] = (int32_t)philox_below(r.v[0], (uint32_t)P.na);
  }
}

__global__ void uavsim_stats_clear_kernel(double *stats, int count) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) stats[k] = 0.0;
}

// fixed-order reduction of the per-CTA slots: one CTA of 256 threads, strided partial sums then a shared-memory
// tree (the association order depends only on `slots`, so totals are reproducible)
__global__ void uavsim_stats_reduce_kernel(const double *__restrict__ partial, int slots, double *__restrict__ out) {
  __shared__ double sh[256][STAT_W];
  if (blockIdx.x != 0) return;
  double a[STAT_W] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = threadIdx.x; s < 2 * slots; s += 256) {
    const double *p = partial + (size_t)s * STAT_W;
    for (int k = 0; k < 5; k++) a[k] += p[k];
    a[5] = fmax(a[5], p[5]);
    a[6] += p[6];
  }
  {  // third region: the fast step kernel's reward sums as 64-bit counts of 2^-22 (exact in any order)
    const long long *fx = reinterpret_cast<const long long *>(partial + (size_t)2 * slots * STAT_W);
    long long c[4] = {0, 0, 0, 0};
    for (int s = threadIdx.x; s < slots; s += 256)
      for (int k = 0; k < 4; k++) c[k] += fx[(size_t)s * STAT_W + k];
    __shared__ long long shc[256][4];
    for (int k = 0; k < 4; k++) shc[threadIdx.x][k] = c[k];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
      if ((int)threadIdx.x < w)
        for (int k = 0; k < 4; k++) shc[threadIdx.x][k] += shc[threadIdx.x + w][k];
      __syncthreads();
    }
    if (threadIdx.x == 0)
      for (int k = 0; k < 4; k++) a[k] += (double)shc[0][k] * (1.0 / 4194304.0);
  }
  for (int k = 0; k < STAT_W; k++) sh[threadIdx.x][k] = a[k];
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) {
      for (int k = 0; k < 5; k++) sh[threadIdx.x][k] += sh[threadIdx.x + w][k];
      sh[threadIdx.x][5] = fmax(sh[threadIdx.x][5], sh[threadIdx.x + w][5]);
      sh[threadIdx.x][6] += sh[threadIdx.x + w][6];
    }
    __syncthreads();
  }
  if (threadIdx.x < STAT_W) out[threadIdx.x] = sh[0][threadIdx.x];
}

// policy.cuh -- fused policy step for the batched rollout.
//
// The reference draws one action per agent with a batch-of-one actor forward, `Categorical(probs).sample()` and an
// `.item()` sync (src/models/actor_critic.py:138-148, src/train.py:160-176).  With the environment resident on the GPU
// the policy step of a rollout is `probs = softmax(fc2(relu(fc1(obs))))` over all E*n rows followed by one categorical
// draw per row (src/models/actor_critic.py:85-99): six small library kernels in stock PyTorch, which at 10x10 cost
// twice the environment step.  Here it is ONE launch: a thread owns four rows, packed two by two into the halves of
// Blackwell's f32x2 instructions, so every weight fetched from shared memory (broadcast) serves four rows; the hidden
// activation is never materialised (each hidden unit is consumed by the logit accumulators as soon as it is
// computed); softmax, inverse-CDF draw (Philox4x32-10 keyed (seed; row, counter)) and the optional probability
// output happen in registers.  fp32 throughout, like the torch module.
#pragma once
#include "common.cuh"
#include "philox.cuh"

#define POLICY_NT 128        // threads per CTA
#define POLICY_ROWS 4        // rows per thread
#define POLICY_IN 12         // state_dim (src/environment.py:29)
#define POLICY_AMAX 16       // max actions (reference: na = 12)
#define POLICY_HMAX 512      // max hidden units (reference: 128, src/configs/*.yaml actor_critic.hidden_dim)

struct PolicyDev {
  const float *w1, *b1, *w2, *b2;  // fc1.weight [H,12], fc1.bias [H], fc2.weight [A,H], fc2.bias [A]  (torch layouts)
  int H, A;
};

__device__ __forceinline__ uint64_t pol_pack2(float lo, float hi) {
  return (uint64_t)__float_as_uint(lo) | ((uint64_t)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ uint64_t pol_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float pol_lo(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float pol_hi(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }

// shared memory: per hidden unit j a record of 32 floats {w1[j][0..11], b1[j], 0,0,0, w2[0..A-1][j], 0..} (128-bit loads)
// AP = number of actions rounded up to a multiple of 4; the padding logits are -inf (probability 0)
template <int AP>
__global__ void __launch_bounds__(POLICY_NT)
uavsim_policy_kernel(const float *__restrict__ obs, int64_t rows, const PolicyDev W, uint64_t seed, uint64_t counter,
                     int32_t *__restrict__ actions, float *__restrict__ probs) {
  extern __shared__ __align__(16) float s_w[];
  const int H = W.H, A = W.A;
  for (int k = threadIdx.x; k < H * 32; k += POLICY_NT) {
    const int j = k >> 5, e = k & 31;
    float v = 0.f;
    if (e < POLICY_IN) v = W.w1[j * POLICY_IN + e];
    else if (e == POLICY_IN) v = W.b1[j];
    else if (e >= 16 && e < 16 + A) v = W.w2[(e - 16) * H + j];
    s_w[k] = v;
  }
  __syncthreads();

  const int64_t stride = (int64_t)gridDim.x * POLICY_NT * POLICY_ROWS;
  for (int64_t r0 = ((int64_t)blockIdx.x * POLICY_NT + threadIdx.x) * POLICY_ROWS; r0 < rows; r0 += stride) {
    // rows r0 .. r0+3 (clamped reads; stores are guarded): x01[k] = {row0[k], row1[k]}, x23[k] = {row2[k], row3[k]}
    uint64_t x01[POLICY_IN], x23[POLICY_IN];
    {
      const float4 *p[POLICY_ROWS];
#pragma unroll
      for (int q = 0; q < POLICY_ROWS; q++) p[q] = reinterpret_cast<const float4 *>(obs + min(r0 + q, rows - 1) * POLICY_IN);
#pragma unroll
      for (int v = 0; v < POLICY_IN / 4; v++) {
        const float4 a = p[0][v], b = p[1][v], c = p[2][v], d = p[3][v];
        x01[4 * v + 0] = pol_pack2(a.x, b.x); x01[4 * v + 1] = pol_pack2(a.y, b.y);
        x01[4 * v + 2] = pol_pack2(a.z, b.z); x01[4 * v + 3] = pol_pack2(a.w, b.w);
        x23[4 * v + 0] = pol_pack2(c.x, d.x); x23[4 * v + 1] = pol_pack2(c.y, d.y);
        x23[4 * v + 2] = pol_pack2(c.z, d.z); x23[4 * v + 3] = pol_pack2(c.w, d.w);
      }
    }
    uint64_t l01[AP], l23[AP];
#pragma unroll
    for (int a = 0; a < AP; a++) { const float b = a < A ? W.b2[a] : -INFINITY; l01[a] = pol_pack2(b, b); l23[a] = l01[a]; }

#pragma unroll 2
    for (int j = 0; j < H; j++) {
      const float4 *wj = reinterpret_cast<const float4 *>(s_w + j * 32);
      const float4 wa = wj[0], wb = wj[1], wc = wj[2], wd = wj[3];  // w1[j][0..11], {b1[j], 0, 0, 0}
      const float w1j[POLICY_IN] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y, wc.z, wc.w};
      uint64_t h01 = pol_pack2(wd.x, wd.x), h23 = h01;
#pragma unroll
      for (int k = 0; k < POLICY_IN; k++) {
        const uint64_t w = pol_pack2(w1j[k], w1j[k]);
        h01 = pol_fma2(w, x01[k], h01);
        h23 = pol_fma2(w, x23[k], h23);
      }
      h01 = pol_pack2(fmaxf(pol_lo(h01), 0.f), fmaxf(pol_hi(h01), 0.f));  // ReLU
      h23 = pol_pack2(fmaxf(pol_lo(h23), 0.f), fmaxf(pol_hi(h23), 0.f));
      float w2j[AP];
#pragma unroll
      for (int v = 0; v < AP / 4; v++) {
        const float4 t = wj[4 + v];
        w2j[4 * v] = t.x; w2j[4 * v + 1] = t.y; w2j[4 * v + 2] = t.z; w2j[4 * v + 3] = t.w;
      }
#pragma unroll
      for (int a = 0; a < AP; a++) {
        const uint64_t w = pol_pack2(w2j[a], w2j[a]);
        l01[a] = pol_fma2(w, h01, l01[a]);
        l23[a] = pol_fma2(w, h23, l23[a]);
      }
    }

    // softmax + one categorical draw per row (inverse CDF on the unnormalised weights)
#pragma unroll
    for (int q = 0; q < POLICY_ROWS; q++) {
      const int64_t r = r0 + q;
      if (r >= rows) break;
      float lg[AP];
#pragma unroll
      for (int a = 0; a < AP; a++) {
        const uint64_t v = (q < 2) ? l01[a] : l23[a];
        lg[a] = (q & 1) ? pol_hi(v) : pol_lo(v);
      }
      float mx = lg[0];
#pragma unroll
      for (int a = 1; a < AP; a++) mx = fmaxf(mx, lg[a]);
      float sum = 0.f;
#pragma unroll
      for (int a = 0; a < AP; a++) { lg[a] = expf(lg[a] - mx); sum += lg[a]; }
      const Philox4 rn = philox4x32_10((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)counter, (uint32_t)(counter >> 32), seed);
      const float u = (float)philox_u53(rn.v[0], rn.v[1]) * sum;
      // inverse CDF: the action is the number of prefix sums that do not exceed u, capped at the last action
      float run = 0.f;
      int act = 0;
#pragma unroll
      for (int a = 0; a < AP - 1; a++) { run += lg[a]; act += (run <= u && a < A - 1) ? 1 : 0; }
      actions[r] = act;
      if (probs) {
        const float inv = 1.0f / sum;
#pragma unroll
        for (int a = 0; a < AP; a++) if (a < A) probs[r * A + a] = lg[a] * inv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int AP>
static int policy_launch(const float *obs, int64_t rows, const PolicyDev &W, uint64_t seed, uint64_t counter,
                         int32_t *actions, float *probs, int device, cudaStream_t st) {
  const size_t smem = (size_t)W.H * 32 * sizeof(float);
  int rc = raise_dynamic_smem((const void *)uavsim_policy_kernel<AP>, device, smem);
  if (rc) return rc;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  int64_t blocks = (rows + (int64_t)POLICY_NT * POLICY_ROWS - 1) / ((int64_t)POLICY_NT * POLICY_ROWS);
  const int64_t cap = (int64_t)sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  uavsim_policy_kernel<AP><<<(int)blocks, POLICY_NT, smem, st>>>(obs, rows, W, seed, counter, actions, probs);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int uavsim_policy_sample(const float *obs, int64_t rows, const UavSimPolicyWeights *w, uint64_t seed,
                                    uint64_t counter, int32_t *actions, float *probs, int device, void *stream) {
  if (!w || rows < 0 || (rows > 0 && (!obs || !actions)) || !w->w1 || !w->b1 || !w->w2 || !w->b2) {
    SET_ERR("uavsim_policy_sample: NULL argument");
    return UAVSIM_ERR_ARG;
  }
  if (w->state_dim != POLICY_IN || w->hidden < 1 || w->hidden > POLICY_HMAX || w->n_actions < 2 || w->n_actions > POLICY_AMAX) {
    SET_ERR("uavsim_policy_sample: supports state_dim = %d, hidden <= %d, 2 <= n_actions <= %d (got %d, %d, %d)", POLICY_IN,
            POLICY_HMAX, POLICY_AMAX, w->state_dim, w->hidden, w->n_actions);
    return UAVSIM_ERR_UNSUPPORTED;
  }
  if (rows == 0) return 0;
  CUDA_TRY(cudaSetDevice(device));
  PolicyDev W;
  W.w1 = w->w1; W.b1 = w->b1; W.w2 = w->w2; W.b2 = w->b2; W.H = w->hidden; W.A = w->n_actions;
  cudaStream_t st = (cudaStream_t)stream;
  switch ((w->n_actions + 3) / 4) {
    case 1: return policy_launch<4>(obs, rows, W, seed, counter, actions, probs, device, st);
    case 2: return policy_launch<8>(obs, rows, W, seed, counter, actions, probs, device, st);
    case 3: return policy_launch<12>(obs, rows, W, seed, counter, actions, probs, device, st);
    case 4: return policy_launch<16>(obs, rows, W, seed, counter, actions, probs, device, st);
  }
  return UAVSIM_ERR_UNSUPPORTED;
}

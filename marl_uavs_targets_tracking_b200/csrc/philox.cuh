// Philox4x32-10 counter-based RNG (Salmon et al., SC'11), written out for device and host.
// Counter = (entity, stream, global env id, step); key = 64-bit seed.  The same function in
// numpy lives in tests/philox_ref.py so the reset / random-policy kernels can be checked bit for bit.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define UAVSIM_HD __host__ __device__ __forceinline__
#else
#define UAVSIM_HD inline
#endif

struct Philox4 {
  uint32_t v[4];
};

UAVSIM_HD uint32_t philox_mulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

UAVSIM_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = philox_mulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = philox_mulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

// 53-bit uniform in [0,1), built like CPython's random.random(): (a>>5, b>>6) -> (a*2^26+b)/2^53
UAVSIM_HD double philox_u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// unbiased-enough integer in [0, n): high word of a 32x32 product
UAVSIM_HD uint32_t philox_below(uint32_t r, uint32_t n) { return philox_mulhi(r, n); }

// RNG stream ids (counter word 1)
enum { UAVSIM_RNG_UAV_RESET = 0, UAVSIM_RNG_TGT_POS = 1, UAVSIM_RNG_TGT_HEAD = 2, UAVSIM_RNG_ACTION = 3 };

// step_fast_kernel.cuh -- the fused environment step for swarms of 64 UAVs x 64 targets (Environment.step,
// src/environment.py:120-164), one environment per 64-thread CTA, persistent over environments.
//
// What differs from the generic kernel (step_kernel.cuh), and why:
//   * Data movement.  The eight state / action arrays of an environment (3.5 KB) arrive by cp.async.bulk (TMA engine,
//     mbarrier complete_tx) while the previous environment is computed, and the outputs (new state, observations,
//     four reward planes: 7.4 KB) leave by bulk shared -> global copies; no thread issues a global load or store
//     on the hot path and there is no per-thread 64-bit address arithmetic.
//   * Pair arithmetic in fp32 with a TWO-SIDED guard.  Positions are staged as fp32 relative to the map centre.
//     With g(R) a proven bound on the fp32 error of a squared distance (R = largest |coordinate - centre| of the
//     environment, KParams::GuardK): s_f <= thr^2 - g is certainly inside the radius, s_f > thr^2 + g certainly
//     outside, and only a pair in the band between the two (about one pair in 10^6) is AMBIGUOUS.  A UAV that meets
//     an ambiguous pair re-evaluates its whole row in fp64 with the reference's arithmetic (fast_agent_exact), so
//     every mask, count and coverage bit is decided exactly as before -- but the candidates of an ordinary UAV never
//     touch the fp64 pipe.  The observation sums are fp32 (outputs are fp32, contract 1e-5; measured ~2e-7).
//   * Candidate walks.  A sign-bit prefilter (packed f32x2 FMAs, as in the generic kernel) marks the partners inside
//     each guarded radius; three short walks consume them: targets (observation + tracking reward + coverage),
//     communication partners (new record if the partner moved first, old record otherwise: src/agent/uav.py:124-147
//     in the update order of src/environment.py:133-138), and duplicate-tracking / neighbour partners
//     (src/agent/uav.py:214-229, :305).
//   * fp64 sine / cosine of the headings by fm_sincos_small (fast_math.cuh) instead of the library call.
// Everything that is not a pair test (kinematics, reflection, boundary term, normalisation, cooperative reward,
// statistics) follows the generic kernel line by line.
#pragma once
#include "common.cuh"
#include "fast_math.cuh"
#include <stddef.h>
#include "step_kernel.cuh"

#define FAST_NT 64  // threads per CTA = UAVs = targets of an environment

// per action: dt * heading-rate (src/agent/uav.py:73-81, :96) and the fp32 cosine / sine of that angle
struct ActEntry {
  double dth;
  float cd, sd;
};

template <int N, int M, bool AUX>
struct __align__(128) FastSmem {
  // inputs of the current environment (bulk-loaded)
  double ux[N], uy[N], uh[N];
  double tx[M], ty[M], th[M];
  int32_t ua[N], act[N];
  // outputs (bulk-stored)
  double oux[N], ouy[N], ouh[N];
  double otx[M], oty[M], oth[M];
  int32_t oua[N];
  float rew[4][N];
  float obs[N * 12];
  // working set of the pair phase (fp32, positions relative to the map centre)
  float4 pn2[N / 2];   // new UAV positions, two per entry {x0, x1, y0, y1}: operands of the packed prefilter
  float4 tp2[M / 2];   // target positions, same pairing
  float4 recn[N][2];   // UAV after its move  {x, y, cos h, sin h} {a, -, -, -}
  float4 reco[N][2];   // UAV before its move (same layout; reco - recn is a compile-time offset)
  float4 trec[M];      // target {x, y, cos h * tv/uv, sin h * tv/uv}
  double xo[N], yo[N]; // fp64 positions before the move (exact path only)
  double raw[N];       // weighted raw reward of every UAV (neighbour mean)
  int32_t tcnt[AUX ? M : 1];
  uint32_t cover[2][2];  // per warp: targets with a UAV strictly inside dp
  uint32_t rmax[2];      // per warp: largest |coordinate - centre| as float bits
  unsigned long long mbar;
};

// ---- single-thread async-copy instructions, issued by a CONVERGED warp and predicated on one elected lane inside the
//      asm: under a divergent `if (t == 0)` every operand of these uniform-datapath instructions goes through a
//      per-instruction waterfall loop (see pmi_tc_kernel.cuh) ----
__device__ __forceinline__ void sf_expect_tx(uint32_t lead, uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
               "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes), "r"(lead) : "memory");
}
__device__ __forceinline__ void sf_bulk_g2s(uint32_t lead, uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
               "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "r"(lead) : "memory");
}
__device__ __forceinline__ void sf_bulk_s2g(uint32_t lead, void *dst, uint32_t src, uint32_t bytes) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t"
               "@q cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\t}" ::"l"(dst), "r"(src), "r"(bytes), "r"(lead) : "memory");
}
__device__ __forceinline__ uint32_t sf_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sf_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; spin++) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
    if (spin > (1u << 22)) __trap();  // a copy that never lands traps instead of hanging the GPU
  }
}

__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float f2_lo(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float f2_hi(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ int sf_msb(uint32_t w) {  // index of the highest set bit (w != 0): one FLO
  int r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(w));
  return r;
}
__device__ __forceinline__ float sf_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Candidate sets are kept as an EVEN and an ODD word: bit b of the even word is partner 2b, of the odd word partner
// 2b+1.  Both words span the whole index range, so a lane's candidates split evenly between them and the two walks
// (one per word, each as long as its busiest lane) together take about as many trips as the busiest lane has
// candidates -- with one word per 32 consecutive partners a swarm flying in index order needed up to 1.5 x that.
//
// Sign-bit prefilter over 64 partner positions (32 pairs {x0, x1, y0, y1}) against one or two guarded squared radii:
// the partner's bit is set iff fma(dx, dx, fma(dy, dy, -thr)) < 0.
template <bool TWO>
__device__ __forceinline__ void prefilter64(const float4 *__restrict__ pf, float xf, float yf, float thrA, float thrB,
                                            uint32_t &aE, uint32_t &aO, uint32_t &bE, uint32_t &bO) {
  const uint64_t xf2 = pack2(xf, xf), yf2 = pack2(yf, yf), nA = pack2(-thrA, -thrA), nB = pack2(-thrB, -thrB);
  uint32_t ae = 0, ao = 0, be = 0, bo = 0;
#pragma unroll 1
  for (int k = 0; k < 32; k += 4) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const ulonglong2 p = reinterpret_cast<const ulonglong2 *>(pf)[k + u];
      const uint64_t dx = f2_sub(p.x, xf2), dy = f2_sub(p.y, yf2);
      const uint64_t tA = f2_fma(dx, dx, f2_fma(dy, dy, nA));
      ae = __funnelshift_l((uint32_t)tA, ae, 1);
      ao = __funnelshift_l((uint32_t)(tA >> 32), ao, 1);
      if (TWO) {
        const uint64_t tB = f2_fma(dx, dx, f2_fma(dy, dy, nB));
        be = __funnelshift_l((uint32_t)tB, be, 1);
        bo = __funnelshift_l((uint32_t)(tB >> 32), bo, 1);
      }
    }
  }
  aE = __brev(ae); aO = __brev(ao);  // pair 0 went in first and sits in bit 31
  bE = __brev(be); bO = __brev(bo);
}

// even / odd words -> the natural 64-bit set (bit j = partner j): only the PMI hand-over needs it
__device__ __forceinline__ uint64_t sf_interleave(uint32_t e, uint32_t o) {
  auto spread = [](uint32_t v) {
    uint64_t x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
  };
  return spread(e) | (spread(o) << 1);
}

// ------------------------------------------------------------------------------------------------
// exact path: the reference's arithmetic pair by pair in fp64, including the min(dist, 1) row weights of
// src/agent/uav.py:162-186.  Taken by a UAV that met an ambiguous pair, by UAVs within 2 m of the origin in both
// coordinates (the only place where a row weight differs from 1) and by every UAV of an environment with an
// entity outside the radius the guard is proven for.  Cold: never inlined; its constants and mask pointers travel
// by value (a reference to the kernel's parameter block would force a copy of the whole block into local memory at
// kernel entry) and its results come back through the UAV's own slots of the output staging area (a pointer to
// registers of the caller would push them into local memory on the hot path as well):
//   obs[12 i .. +8] the nine list entries of the local state, obs[12 i + 9] = tt, obs[12 i + 10] = dup,
//   rew[0][i], rew[1][i] = neighbour words (even, odd), rew[2][i], rew[3][i] = coverage words.
// ------------------------------------------------------------------------------------------------
struct ExactK {
  double s_dp_le, s_dp_lt, s_2dp_le, s_dc_le, dp, dc, two_dp;
  int na;
};
struct MaskPtrs {
  uint8_t *obs_mask, *comm_mask, *nbr_mask, *dup_mask, *cover_mask;
};
template <int N, int M, bool AUX>
__device__ __noinline__ void fast_agent_exact(const ExactK P, const MaskPtrs B, FastSmem<N, M, AUX> &S, int i,
                                              int64_t mrow_t, int64_t mrow_u) {
  const double xi = S.oux[i], yi = S.ouy[i];
  const float4 me = S.recn[i][0];
  const double chi = (double)me.z, shi = (double)me.w;
  const int ai = S.oua[i];
  double tt = 0, o0 = 0, o1 = 0, o2 = 0, o3 = 0;
  int nobs = 0;
  uint32_t cov[2] = {0, 0};
  for (int t = 0; t < M; t++) {
    const double dx = S.otx[t] - xi, dy = S.oty[t] - yi;
    const double d2 = dx * dx + dy * dy;
    const bool hit = d2 <= P.s_dp_le, cv = d2 <= P.s_dp_lt;
    if (AUX && B.obs_mask) { B.obs_mask[mrow_t + t] = hit; B.cover_mask[mrow_t + t] = cv; }
    if (cv) cov[t & 1] |= 1u << (t >> 1);
    if (hit) {
      const double d = sqrt(d2);
      tt += 1 + (P.dp - d) / P.dp;  // uav.py:208
      const float4 tr = S.trec[t];
      double rx = dx / P.dp, ry = dy / P.dp, vx = (double)tr.z - chi, vy = (double)tr.w - shi;
      const double wx = rx - xi, wy = ry - yi, w2 = wx * wx + wy * wy;  // uav.py:174-180
      if (w2 < 1.0) { const double w = sqrt(w2); rx /= w; ry /= w; vx /= w; vy /= w; }
      o0 += rx; o1 += ry; o2 += vx; o3 += vy;
      nobs++;
    }
  }
  double dup = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
  int ncomm = 0;
  uint32_t nb[2] = {0, 0};
  for (int j = 0; j < N; j++) {
    if (j == i) {
      if (AUX && B.obs_mask) { B.comm_mask[mrow_u + j] = 0; B.nbr_mask[mrow_u + j] = 0; B.dup_mask[mrow_u + j] = 0; }
      continue;
    }
    const double dxn = S.oux[j] - xi, dyn = S.ouy[j] - yi;
    const double d2n = dxn * dxn + dyn * dyn;
    const bool hit_dup = d2n <= P.s_2dp_le, hit_nbr = d2n <= P.s_dp_le;
    if (hit_dup) { const double d = sqrt(d2n); dup += -0.5 * exp((P.two_dp - d) / P.two_dp); }  // uav.py:226
    if (hit_nbr) nb[j & 1] |= 1u << (j >> 1);
    double dxc, dyc, d2c;
    float4 r0;
    float aj;
    if (j < i) { dxc = dxn; dyc = dyn; d2c = d2n; r0 = S.recn[j][0]; aj = S.recn[j][1].x; }
    else { dxc = S.xo[j] - xi; dyc = S.yo[j] - yi; d2c = dxc * dxc + dyc * dyc; r0 = S.reco[j][0]; aj = S.reco[j][1].x; }
    const bool hit_c = d2c <= P.s_dc_le;
    if (AUX && B.obs_mask) { B.comm_mask[mrow_u + j] = hit_c; B.nbr_mask[mrow_u + j] = hit_nbr; B.dup_mask[mrow_u + j] = hit_dup; }
    if (hit_c) {
      double rx = dxc / P.dc, ry = dyc / P.dc, vx = (double)r0.z - chi, vy = (double)r0.w - shi;
      double da = ((double)aj - (double)ai) / (double)P.na;
      const double wx = rx - xi, wy = ry - yi, w2 = wx * wx + wy * wy;
      if (w2 < 1.0) { const double w = sqrt(w2); rx /= w; ry /= w; vx /= w; vy /= w; da /= w; }
      c0 += rx; c1 += ry; c2 += vx; c3 += vy; c4 += da;
      ncomm++;
    }
  }
  float *ob = S.obs + i * 12;
  if (ncomm) {
    const double k = (double)ncomm;
    ob[0] = (float)(c0 / k); ob[1] = (float)(c1 / k); ob[2] = (float)(c2 / k); ob[3] = (float)(c3 / k); ob[4] = (float)(c4 / k);
  } else {
    ob[0] = ob[1] = ob[2] = ob[3] = ob[4] = -1.f;
  }
  if (nobs) {
    const double k = (double)nobs;
    ob[5] = (float)(o0 / k); ob[6] = (float)(o1 / k); ob[7] = (float)(o2 / k); ob[8] = (float)(o3 / k);
  } else {
    ob[5] = ob[6] = ob[7] = ob[8] = -1.f;
  }
  ob[9] = (float)tt; ob[10] = (float)dup;
  S.rew[0][i] = __uint_as_float(nb[0]); S.rew[1][i] = __uint_as_float(nb[1]);
  S.rew[2][i] = __uint_as_float(cov[0]); S.rew[3][i] = __uint_as_float(cov[1]);
}

// sin / cos of a heading: the wrapped range takes the inline routine, anything else the library one
__device__ __forceinline__ void heading_sincos(double h, double &s, double &c) {
  if (fabs(h) < 4.0) fm_sincos_small(h, &s, &c);
  else sincos_shared(h, &s, &c);
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
#ifndef FAST_CTAS_PER_SM
#define FAST_CTAS_PER_SM 11
#endif
template <int N, int M, bool AUX>
__global__ void __launch_bounds__(FAST_NT, FAST_CTAS_PER_SM)
uavsim_step_fast_kernel(const KParams P, const UavSimBuffers B, const ActEntry *__restrict__ act_tab, int64_t env_begin,
                        int64_t env_count, int mode, double coop, int done_flag, double *__restrict__ stats_partial) {
  static_assert(N == 64 && M == 64 && FAST_NT == 64, "one thread per UAV and per target");
  typedef FastSmem<N, M, AUX> SmemT;
  static_assert(offsetof(SmemT, reco) - offsetof(SmemT, recn) == sizeof(float4) * 2 * N, "record offset");
  extern __shared__ __align__(128) unsigned char fast_smem_raw[];
  SmemT &S = *reinterpret_cast<SmemT *>(fast_smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t bar = sf_smem(&S.mbar);
  constexpr uint32_t IN_BYTES = 3 * N * 8 + 3 * M * 8 + 2 * N * 4;
  const bool pmi_pending = (mode == UAVSIM_MODE_PMI) && (coop != 0.0);

  // warp 0, converged: the eight input arrays of environment e
  auto issue_loads = [&](uint32_t lead, int64_t e) {
    sf_expect_tx(lead, bar, IN_BYTES);
    sf_bulk_g2s(lead, sf_smem(S.ux), B.ux + e * N, N * 8, bar);
    sf_bulk_g2s(lead, sf_smem(S.uy), B.uy + e * N, N * 8, bar);
    sf_bulk_g2s(lead, sf_smem(S.uh), B.uh + e * N, N * 8, bar);
    sf_bulk_g2s(lead, sf_smem(S.tx), B.tx + e * M, M * 8, bar);
    sf_bulk_g2s(lead, sf_smem(S.ty), B.ty + e * M, M * 8, bar);
    sf_bulk_g2s(lead, sf_smem(S.th), B.th + e * M, M * 8, bar);
    sf_bulk_g2s(lead, sf_smem(S.ua), B.ua + e * N, N * 4, bar);
    sf_bulk_g2s(lead, sf_smem(S.act), B.actions + e * N, N * 4, bar);
  };

  if (t == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t lead = 0;
  if (warp == 0) {
    lead = elect_one();
    if ((int64_t)blockIdx.x < env_count) issue_loads(lead, env_begin + blockIdx.x);
  }

  const int64_t plane = P.E * N;
  const float inv_dp_f = (float)P.inv_dp, inv_dc_f = (float)P.inv_dc, inv_na_f = (float)P.inv_na;
  const float k_ex0 = 1.4426950408889634f, k_ex1 = (float)(-1.4426950408889634 / P.two_dp);
  double st_r = 0, st_tt = 0, st_bp = 0, st_dup = 0, st_cov = 0, st_envs = 0;
  int st_cmax = 0;
  uint32_t parity = 0;
  // partners below the own index inside the even / odd word (2b < t, 2b + 1 < t) and the own bit
  const uint32_t ltE = low_bits((t + 1) >> 1), ltO = low_bits(t >> 1);
  const uint32_t selfE = (t & 1) ? 0u : (1u << (t >> 1)), selfO = (t & 1) ? (1u << (t >> 1)) : 0u;

  for (int64_t k = blockIdx.x; k < env_count; k += gridDim.x) {
    const int64_t e = env_begin + k;
    if (lead) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous outputs have left shared memory
    __syncthreads();
    sf_mbar_wait(bar, parity);
    parity ^= 1;

    // ---- phase 0a: target t (src/agent/target.py:27-60) ----
    float rabs;
    {
      double x = S.tx[t], y = S.ty[t], h = S.th[t];
      double sh, ch;
      heading_sincos(h, sh, ch);
      x += P.dtv_t * ch;
      y += P.dtv_t * sh;
      // reflection (target.py:52-58); cos(-h) = cos h, sin(-h) = -sin h, cos(+-pi - h) = -cos h, sin(+-pi - h) = sin h
      if (0 > y || y > P.y_max) { h = -h; sh = -sh; }
      else if (x < 0 || x > P.x_max) { h = (h > 0) ? (PI_D - h) : (-PI_D - h); ch = -ch; }
      S.otx[t] = x; S.oty[t] = y; S.oth[t] = h;
      const float xf = (float)(x - P.cx), yf = (float)(y - P.cy);
      // cos(target.h) * target.v_max / self.v_max  (src/agent/uav.py:115-116)
      S.trec[t] = make_float4(xf, yf, (float)(ch * P.tv_over_uv), (float)(sh * P.tv_over_uv));
      EnvView::put_pair(S.tp2, t, xf, yf);
      rabs = fmaxf(fabsf(xf), fabsf(yf));
      if (AUX) S.tcnt[t] = 0;
    }
    // ---- phase 0b: UAV t (src/agent/uav.py:73-99) ----
    double xi, yi;
    float xf, yf, chf, shf, xl, yl;  // own fp32 position relative to the centre, heading, and what fp32 dropped of the position
    int ai;
    {
      double x = S.ux[t], y = S.uy[t], h = S.uh[t];
      const int a_old = S.ua[t], act = S.act[t];
      double sh, ch;
      heading_sincos(h, sh, ch);
      const float cof = (float)ch, sof = (float)sh;
      const float xof = (float)(x - P.cx), yof = (float)(y - P.cy);
      S.xo[t] = x; S.yo[t] = y;
      S.reco[t][0] = make_float4(xof, yof, cof, sof);
      S.reco[t][1].x = (float)a_old;
      x += P.dtv_u * ch;
      y += P.dtv_u * sh;
      double dh;
      float cd, sd;
      if ((unsigned)act < (unsigned)P.na) {
        const ActEntry en = act_tab[act];
        dh = en.dth; cd = en.cd; sd = en.sd;
      } else {  // the reference's formula accepts any integer (uav.py:73-81)
        dh = P.dt * ((double)(2 * (act + 1) - P.na - 1) * P.uav_h_max / (double)(P.na - 1));
        double sd_, cd_;
        sincos_shared(dh, &sd_, &cd_);
        cd = (float)cd_; sd = (float)sd_;
      }
      h = wrap_heading(h + dh);
      // cos / sin of the new heading by angle addition in fp32: they only feed the observation (the next step
      // evaluates the stored heading again)
      chf = fmaf(cof, cd, -(sof * sd));
      shf = fmaf(sof, cd, cof * sd);
      xf = (float)(x - P.cx); yf = (float)(y - P.cy);
      xl = (float)((x - P.cx) - (double)xf); yl = (float)((y - P.cy) - (double)yf);
      S.oux[t] = x; S.ouy[t] = y; S.ouh[t] = h; S.oua[t] = act;
      S.recn[t][0] = make_float4(xf, yf, chf, shf);
      S.recn[t][1].x = (float)act;
      EnvView::put_pair(S.pn2, t, xf, yf);
      rabs = fmaxf(fmaxf(rabs, fmaxf(fabsf(xf), fabsf(yf))), fmaxf(fabsf(xof), fabsf(yof)));
      // fp32 sums of action indices are exact only for small integers: anything else takes the exact path
      if ((unsigned)act >= 4096u || (unsigned)a_old >= 4096u) rabs = __int_as_float(0x7f800000);
      if (!(rabs == rabs)) rabs = __int_as_float(0x7f800000);
      xi = x; yi = y; ai = act;
    }
    {
      const uint32_t rm = __reduce_max_sync(0xffffffffu, __float_as_uint(rabs));
      if (lane == 0) S.rmax[warp] = rm;
    }
    __syncthreads();
    if (warp == 0 && k + gridDim.x < env_count) issue_loads(lead, e + gridDim.x);  // next environment, behind the pair phase

    // ---- phase 1: pair tests, observation, raw reward ----
    const float R = __uint_as_float(max(S.rmax[0], S.rmax[1]));
    const bool far_env = !(R <= P.r_fast);
    const bool near_origin = fabs(xi) < 2.0 && fabs(yi) < 2.0;  // the only place a row weight differs from 1
    const int64_t mrow_t = (e * N + t) * M, mrow_u = (e * N + t) * N;
    // results of the pair phase: nine list entries of the local state, raw tracking / duplicate terms, the
    // neighbour set d <= dp and the targets strictly inside dp (even / odd words)
    float o0, o1, o2, o3, o4, o5, o6, o7, o8, tt_f, dup_f;
    uint32_t nbE = 0, nbO = 0, cvE = 0, cvO = 0;
    uint32_t cmE = 0, cmO = 0, dpE = 0, dpO = 0;  // AUX: communication / duplicate sets for the mask outputs
    bool exact = far_env || near_origin;
    if (!exact) {
      const float g_dp = fmaf(R, P.g_dp.c1, P.g_dp.c0), g_2dp = fmaf(R, P.g_2dp.c1, P.g_2dp.c0);
      const float g_dc = fmaf(R, P.g_dc.c1, P.g_dc.c0), g_pf = fmaf(R, P.g_pf.c1, P.g_pf.c0);
      const float Tp_hi = __fadd_ru(P.g_dp.t2_up, g_dp), Tp_lo = __fadd_rd(P.g_dp.t2_dn, -g_dp);
      const float T2_hi = __fadd_ru(P.g_2dp.t2_up, g_2dp), T2_lo = __fadd_rd(P.g_2dp.t2_dn, -g_2dp);
      const float Tc_hi = __fadd_ru(P.g_dc.t2_up, g_dc), Tc_lo = __fadd_rd(P.g_dc.t2_dn, -g_dc);
      const float Tf_hi = __fadd_ru(P.g_pf.t2_up, g_pf);
      // ambiguity trackers: the largest squared distance among the pairs each walk ACCEPTS as inside a radius; a
      // walk is unambiguous iff that stays at or below the radius' lower guard
      float smax_t = 0.f, smax_c = 0.f, smax_d = 0.f, smax_n = 0.f;
      const uint64_t me2 = pack2(xf, yf);

      // -- targets: observe_target (uav.py:101-122), tracking reward (uav.py:199-212), coverage (environment.py:246-253)
      {
        uint32_t d0, d1;
        prefilter64<false>(S.tp2, xf, yf, Tp_hi, 0.f, cvE, cvO, d0, d1);
        const int nobs = __popc(cvE) + __popc(cvO);
        uint64_t od = 0, ov = 0;  // packed sums {dx, dy}, {vx, vy}
        float ttacc = 0;
#pragma unroll
        for (int par = 0; par < 2; par++) {
          uint32_t w = par ? cvO : cvE;
          const ulonglong2 *rec = reinterpret_cast<const ulonglong2 *>(S.trec + par);
#pragma unroll 1
          while (w) {
            const int b = sf_msb(w);
            w ^= 1u << b;
            const ulonglong2 tr = rec[2 * b];    // {x, y}, {vx, vy}
            const uint64_t d = f2_sub(tr.x, me2);
            const uint64_t q = f2_mul(d, d);
            const float s = f2_lo(q) + f2_hi(q);
            smax_t = fmaxf(smax_t, s);
            od = f2_add(od, d);
            ov = f2_add(ov, tr.y);
            ttacc = fmaf(fast_sqrtf(s), -inv_dp_f, ttacc);  // sum of (dp - d)/dp - 1
          }
        }
        if (nobs) {
          const float kf = (float)nobs, rk = sf_rcp(kf);
          // the own coordinate's fp32 rounding is common to every row: taken out of the mean exactly
          o5 = fmaf(f2_lo(od), rk, -xl) * inv_dp_f;
          o6 = fmaf(f2_hi(od), rk, -yl) * inv_dp_f;
          o7 = fmaf(f2_lo(ov), rk, -chf);
          o8 = fmaf(f2_hi(ov), rk, -shf);
          tt_f = fmaf(2.0f, kf, ttacc);  // sum of 1 + (dp - d)/dp
        } else {
          o5 = o6 = o7 = o8 = -1.f;
          tt_f = 0.f;
        }
      }

      // -- UAV partners: prefilter on the NEW positions against (i) the communication radius widened by one move
      //    (a partner that moves after this UAV is tested at its old position, at most dt*v from the new one) and
      //    (ii) the duplicate-tracking radius 2 dp
      uint32_t ccE, ccO, cdE, cdO;
      prefilter64<true>(S.pn2, xf, yf, Tf_hi, T2_hi, ccE, ccO, cdE, cdO);
      ccE &= ~selfE; ccO &= ~selfO; cdE &= ~selfE; cdO &= ~selfO;  // never its own partner
      // -- communication partners (uav.py:124-147): partner j < i already moved -> its new record, j > i -> its old one
      {
        uint64_t sd = 0, sh2 = 0;  // packed sums {dx, dy}, {cos, sin}
        float sa = 0;
        int cnt = 0;
        constexpr uint32_t OLD_OFF = (uint32_t)(sizeof(float4) * 2 * N);
#pragma unroll
        for (int par = 0; par < 2; par++) {
          uint32_t w = par ? ccO : ccE;
          const uint32_t lt = par ? ltO : ltE;
          const unsigned char *rec_new = reinterpret_cast<const unsigned char *>(&S.recn[par][0]);
#pragma unroll 1
          while (w) {
            const int b = sf_msb(w);
            const uint32_t bit = 1u << b;
            w ^= bit;
            const unsigned char *rp = ((bit & lt) ? rec_new : rec_new + OLD_OFF) + 64 * b;
            const ulonglong2 r0 = *reinterpret_cast<const ulonglong2 *>(rp);  // {x, y}, {cos h, sin h}
            const uint64_t d = f2_sub(r0.x, me2);
            const uint64_t q = f2_mul(d, d);
            const float s = f2_lo(q) + f2_hi(q);
            if (s <= Tc_hi) {
              const float aj = *reinterpret_cast<const float *>(rp + 16);
              smax_c = fmaxf(smax_c, s);
              sd = f2_add(sd, d);
              sh2 = f2_add(sh2, r0.y);
              sa += aj; cnt++;
              if (AUX) { if (par) cmO |= bit; else cmE |= bit; }
            }
          }
        }
        if (cnt) {
          const float kf = (float)cnt, rk = sf_rcp(kf);
          o0 = fmaf(f2_lo(sd), rk, -xl) * inv_dc_f;
          o1 = fmaf(f2_hi(sd), rk, -yl) * inv_dc_f;
          o2 = fmaf(f2_lo(sh2), rk, -chf);
          o3 = fmaf(f2_hi(sh2), rk, -shf);
          o4 = (sa - kf * (float)ai) * (rk * inv_na_f);
        } else {
          o0 = o1 = o2 = o3 = o4 = -1.f;
        }
      }
      // -- duplicate-tracking punishment (uav.py:214-229) and the neighbour set (uav.py:305), all at NEW positions
      {
        float dup = 0;
        if (AUX) { dpE = cdE; dpO = cdO; }
#pragma unroll
        for (int par = 0; par < 2; par++) {
          uint32_t w = par ? cdO : cdE, nbits = 0;
          const unsigned char *rec_new = reinterpret_cast<const unsigned char *>(&S.recn[par][0]);
#pragma unroll 1
          while (w) {
            const int b = sf_msb(w);
            const uint32_t bit = 1u << b;
            w ^= bit;
            const uint64_t d = f2_sub(*reinterpret_cast<const uint64_t *>(rec_new + 64 * b), me2);
            const uint64_t q = f2_mul(d, d);
            const float s = f2_lo(q) + f2_hi(q);
            smax_d = fmaxf(smax_d, s);
            dup += fast_ex2f(fmaf(fast_sqrtf(s), k_ex1, k_ex0));  // exp((2dp - d)/(2dp))
            if (s <= Tp_hi) {
              nbits |= bit;
              smax_n = fmaxf(smax_n, s);
            }
          }
          if (par) nbO = nbits; else nbE = nbits;
        }
        dup_f = -0.5f * dup;
      }
      exact = (smax_t > Tp_lo) || (smax_c > Tc_lo) || (smax_d > T2_lo) || (smax_n > Tp_lo);
    }
    if (exact) {
      const ExactK XK = {P.s_dp_le, P.s_dp_lt, P.s_2dp_le, P.s_dc_le, P.dp, P.dc, P.two_dp, P.na};
      const MaskPtrs MP = {B.obs_mask, B.comm_mask, B.nbr_mask, B.dup_mask, B.cover_mask};
      fast_agent_exact<N, M, AUX>(XK, MP, S, t, mrow_t, mrow_u);
      const float *ob = S.obs + t * 12;
      o0 = ob[0]; o1 = ob[1]; o2 = ob[2]; o3 = ob[3]; o4 = ob[4]; o5 = ob[5]; o6 = ob[6]; o7 = ob[7]; o8 = ob[8];
      tt_f = ob[9]; dup_f = ob[10];
      nbE = __float_as_uint(S.rew[0][t]); nbO = __float_as_uint(S.rew[1][t]);
      cvE = __float_as_uint(S.rew[2][t]); cvO = __float_as_uint(S.rew[3][t]);
    } else if (AUX && B.obs_mask) {
      for (int j = 0; j < 64; j++) {
        const int b = j >> 1;
        const uint8_t v = (((j & 1) ? cvO : cvE) >> b) & 1u;
        B.obs_mask[mrow_t + j] = v; B.cover_mask[mrow_t + j] = v;
        B.comm_mask[mrow_u + j] = (((j & 1) ? cmO : cmE) >> b) & 1u;
        B.nbr_mask[mrow_u + j] = (((j & 1) ? nbO : nbE) >> b) & 1u;
        B.dup_mask[mrow_u + j] = (((j & 1) ? dpO : dpE) >> b) & 1u;
      }
    }
    if (AUX && B.tracker_cnt) {
#pragma unroll
      for (int par = 0; par < 2; par++) {
        uint32_t w = par ? cvO : cvE;
        while (w) { const int b = sf_msb(w); w ^= 1u << b; atomicAdd(&S.tcnt[2 * b + par], 1); }
      }
    }
    {  // coverage: targets with at least one UAV strictly inside dp (environment.py:246-253), OR over the warp
      const uint32_t c0 = __reduce_or_sync(0xffffffffu, cvE), c1 = __reduce_or_sync(0xffffffffu, cvO);
      if (lane == 0) { S.cover[warp][0] = c0; S.cover[warp][1] = c1; }
    }

    double raw, ttn, bpn, dupn;
    {
      float *ob = S.obs + t * 12;
      reinterpret_cast<float4 *>(ob)[0] = make_float4(o0, o1, o2, o3);
      reinterpret_cast<float4 *>(ob)[1] = make_float4(o4, o5, o6, o7);
      reinterpret_cast<float4 *>(ob)[2] = make_float4(o8, (float)(xi * P.inv_dc), (float)(yi * P.inv_dc), (float)ai * inv_na_f);
      // boundary punishment (uav.py:231-250)
      const double dbdr = fmin(fmin(xi - 0, P.x_max - xi), fmin(yi - 0, P.y_max - yi));
      double bp;
      if (0 <= xi && xi <= P.x_max && 0 <= yi && yi <= P.y_max)
        bp = (dbdr < P.dp) ? (-0.5 * (P.dp - dbdr) * P.inv_dp) : 0.0;
      else
        bp = -0.5;
      // normalise + weights (environment.py:206-220)
      ttn = fmin(fmax((double)tt_f, 0.0), P.tt_hi) * P.inv_tt_hi;
      dupn = (fmin(fmax((double)dup_f, P.dup_lo), 0.0) - P.dup_lo) * P.inv_dup_span - 1.0;
      bpn = (fmin(fmax(bp, -0.5), 0.0) + 0.5) * 2.0 - 1.0;
      raw = P.alpha * ttn + P.beta * bpn + P.gamma * dupn;
      S.raw[t] = raw;
    }
    __syncthreads();

    // ---- phase 2: cooperative reward (environment.py:222-227), coverage count, outputs ----
    {
      double r;
      const int64_t gi = e * N + t;
      if (mode == UAVSIM_MODE_SELF || coop == 0.0) {
        r = raw;  // uav.py:271-272 / :300-301
      } else if (mode == UAVSIM_MODE_MEAN) {
        // uav.py:293-310 -- the conditional expression covers the whole sum: no neighbour -> 0
        double s = 0;
        const int cnt = __popc(nbE) + __popc(nbO);
#pragma unroll
        for (int par = 0; par < 2; par++) {
          uint32_t w = par ? nbO : nbE;
          while (w) { const int b = sf_msb(w); w ^= 1u << b; s += S.raw[2 * b + par]; }
        }
        r = cnt ? ((1 - coop) * raw + coop * s / (double)cnt) : 0.0;
      } else {
        r = 0.0;  // finished by the PMI kernel
        B.raw[gi] = raw;
        B.nbr_bits[gi * 2] = sf_interleave(nbE, nbO);
        B.nbr_bits[gi * 2 + 1] = 0;
      }
      r = fmin(fmax(r, -1.0), 1.0);  // clip_and_normalize(reward, -1, 1) is a plain clip
      S.rew[0][t] = (float)r;
      S.rew[1][t] = (float)ttn;
      S.rew[2][t] = (float)bpn;
      S.rew[3][t] = (float)dupn;
      if (!pmi_pending) st_r += r;
      st_tt += ttn; st_bp += bpn; st_dup += dupn;
      if (AUX && B.tracker_cnt) B.tracker_cnt[e * M + t] = S.tcnt[t];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk copies
    __syncthreads();
    if (warp == 0) {
      sf_bulk_s2g(lead, B.ux + e * N, sf_smem(S.oux), N * 8);
      sf_bulk_s2g(lead, B.uy + e * N, sf_smem(S.ouy), N * 8);
      sf_bulk_s2g(lead, B.uh + e * N, sf_smem(S.ouh), N * 8);
      sf_bulk_s2g(lead, B.ua + e * N, sf_smem(S.oua), N * 4);
      sf_bulk_s2g(lead, B.tx + e * M, sf_smem(S.otx), M * 8);
      sf_bulk_s2g(lead, B.ty + e * M, sf_smem(S.oty), M * 8);
      sf_bulk_s2g(lead, B.th + e * M, sf_smem(S.oth), M * 8);
      sf_bulk_s2g(lead, B.obs + e * N * 12, sf_smem(S.obs), N * 48);
      if (!pmi_pending) sf_bulk_s2g(lead, B.rew4 + e * N, sf_smem(S.rew[0]), N * 4);
      sf_bulk_s2g(lead, B.rew4 + plane + e * N, sf_smem(S.rew[1]), N * 4);
      sf_bulk_s2g(lead, B.rew4 + 2 * plane + e * N, sf_smem(S.rew[2]), N * 4);
      sf_bulk_s2g(lead, B.rew4 + 3 * plane + e * N, sf_smem(S.rew[3]), N * 4);
      if (lead) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (t == 0) {
      const int c = __popc(S.cover[0][0] | S.cover[1][0]) + __popc(S.cover[0][1] | S.cover[1][1]);
      B.covered[e] = c;
      if (B.done) B.done[e] = done_flag;
      st_cov += (double)c;
      st_cmax = max(st_cmax, c);
      st_envs += 1.0;
    }
  }
  if (lead) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // outputs complete before the CTA retires
  __syncthreads();
  block_stats_commit(reinterpret_cast<double *>(S.obs), stats_partial + (size_t)blockIdx.x * STAT_W, st_r, st_tt, st_bp,
                     st_dup, st_cov, st_cmax, st_envs, FAST_NT);
}

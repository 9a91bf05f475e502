// step_kernel.cuh -- the fused environment step kernel (Environment.step, src/environment.py:120-164).
//
// CTA = epb environments, thread q = (env q / n, UAV q % n), epb * n <= NT.  Per step:
//   phase 0  targets move and reflect (src/agent/target.py:27-60), UAVs integrate their heading-rate action
//            (src/agent/uav.py:73-99); old and new UAV records are staged in shared memory as double2 pairs
//            so the pair loops read them with broadcast 128-bit loads.
//   phase 1  one thread per UAV walks all targets and all other UAVs: range tests on exact squared fp64
//            thresholds, branch-free accumulation of the observation sums, tracking / duplicate terms,
//            neighbour bit set, per-target tracker counts; boundary term; normalisation and weights.
//   phase 2  cooperative reward (self / neighbour mean; PMI is finished by uavsim_pmi_kernel), coverage
//            count, coalesced output stores, per-CTA episode statistics.
//
// Precision plan.  Everything that decides an integer output (the five range masks, coverage) is fp64 in
// the reference's evaluation order.  The observation sums are fp64 but LINEAR: mean_j((x_j - x_i)/dc) is
// accumulated as sum(x_j - x_i) and scaled once -- valid because the reference's per-row weight
// min(||(rx,ry) - (x,y)||, 1) (src/agent/uav.py:162-186) is exactly 1 unless |x| < 2 and |y| < 2; UAVs
// inside that 4 m x 4 m corner take the exact per-row path (`EXACTW`).  The transcendental parts of the
// tracking and duplicate terms (sqrt, exp on hits) are evaluated in fp32: they only feed fp32 outputs
// that are normalised by 2m and e/2*n (error <= ~2e-7 against the 1e-5 bar).
#pragma once
#include "common.cuh"

struct StepSmem {
  double2 *tpos, *tvel;          // [epb*m] moved targets: (x, y), (cos h, sin h) * tv / uv
  double2 *npos, *nhd;           // [epb*n] UAV after the move: (x, y), (cos h, sin h)
  double2 *opos, *ohd;           // [epb*n] UAV before the move
  double *raw;                   // [epb*n]
  double *dth;                   // [3*na] per action: dt*rate, cos(dt*rate), sin(dt*rate)
  double *red;                   // [64]
  float *obs;                    // [epb*n*12]
  int *oa, *na_;                 // [epb*n] previous / new action index
  int *tcnt;                     // [epb*m] UAVs strictly within dp of each target
};

static size_t step_smem_bytes(int n, int m, int na, int epb) {
  size_t d = (size_t)epb * m * 4 + (size_t)epb * n * 9 + (size_t)3 * na + 64;
  size_t f = (size_t)epb * n * 12;
  size_t i = (size_t)epb * n * 2 + (size_t)epb * m;
  return d * 8 + 32 + f * 4 + i * 4;
}

__device__ __forceinline__ StepSmem carve(unsigned char *base, int n, int m, int na, int epb) {
  StepSmem s;
  double2 *d2 = reinterpret_cast<double2 *>(base);  // base is 16-byte aligned
  const size_t em = (size_t)epb * m, en = (size_t)epb * n;
  s.tpos = d2; d2 += em; s.tvel = d2; d2 += em;
  s.npos = d2; d2 += en; s.nhd = d2; d2 += en; s.opos = d2; d2 += en; s.ohd = d2; d2 += en;
  double *d = reinterpret_cast<double *>(d2);
  s.raw = d; d += en;
  s.dth = d; d += 3 * na;
  s.red = d; d += 64;
  d += (en + 3 * (size_t)na) & 1;  // keep obs 16-byte aligned with plain pointer arithmetic (stays a shared pointer)
  s.obs = reinterpret_cast<float *>(d);
  int *ip = reinterpret_cast<int *>(s.obs + en * 12);
  s.oa = ip; ip += en; s.na_ = ip; ip += en; s.tcnt = ip;
  return s;
}

// what phase 1 produces for one UAV
struct AgentOut {
  double tt, dup;         // raw tracking reward / duplicate punishment (before normalisation)
  uint32_t nb[4];         // neighbour set d <= dp, bit j of word j/32
};

// approximate fp32 sqrt / exp2 (one MUFU each, ~2 ulp): they only feed fp32 reward terms
__device__ __forceinline__ float fast_sqrtf(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ------------------------------------------------------------------------------------------------
// exact per-row path: the reference's arithmetic hit by hit, including the min(dist,1) row weights.
// Taken by UAVs within 2 m of the origin in both coordinates (never inlined: it is cold).
// ------------------------------------------------------------------------------------------------
template <bool MASKS>
__device__ __noinline__ void agent_exact(const KParams &P, const UavSimBuffers &B, double xi, double yi, double chi,
                                         double shi, int ai, int i, int n, int m, const double2 *tpos,
                                         const double2 *tvel, const double2 *npos, const double2 *nhd,
                                         const double2 *opos, const double2 *ohd, const int *na_, const int *oa,
                                         int *tcnt, float *ob, int64_t mrow_t, int64_t mrow_u, AgentOut &O) {
  double tt = 0, o0 = 0, o1 = 0, o2 = 0, o3 = 0;
  int nobs = 0;
  for (int t = 0; t < m; t++) {
    const double2 tp = tpos[t];
    const double dx = tp.x - xi, dy = tp.y - yi;
    const double d2 = dx * dx + dy * dy;
    const bool hit = d2 <= P.s_dp_le, cov = d2 <= P.s_dp_lt;
    if (MASKS) { B.obs_mask[mrow_t + t] = hit; B.cover_mask[mrow_t + t] = cov; }
    if (hit) {
      const double d = sqrt(d2);
      tt += 1 + (P.dp - d) / P.dp;  // uav.py:208
      const double2 tv = tvel[t];
      double rx = dx / P.dp, ry = dy / P.dp, vx = tv.x - chi, vy = tv.y - shi;
      const double wx = rx - xi, wy = ry - yi, w2 = wx * wx + wy * wy;  // uav.py:174-180
      if (w2 < 1.0) { const double w = sqrt(w2); rx /= w; ry /= w; vx /= w; vy /= w; }
      o0 += rx; o1 += ry; o2 += vx; o3 += vy;
      nobs++;
      if (cov) atomicAdd(&tcnt[t], 1);
    }
  }
  double dup = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
  int ncomm = 0;
  uint32_t nb[4] = {0, 0, 0, 0};
  for (int j = 0; j < n; j++) {
    if (j == i) {
      if (MASKS) { B.comm_mask[mrow_u + j] = 0; B.nbr_mask[mrow_u + j] = 0; B.dup_mask[mrow_u + j] = 0; }
      continue;
    }
    const double2 np = npos[j];
    const double dxn = np.x - xi, dyn = np.y - yi;
    const double d2n = dxn * dxn + dyn * dyn;
    const bool hit_dup = d2n <= P.s_2dp_le, hit_nbr = d2n <= P.s_dp_le;
    if (hit_dup) { const double d = sqrt(d2n); dup += -0.5 * exp((P.two_dp - d) / P.two_dp); }  // uav.py:226
    if (hit_nbr) nb[j >> 5] |= 1u << (j & 31);
    double dxc, dyc, d2c;
    double2 hd;
    int aj;
    if (j < i) { dxc = dxn; dyc = dyn; d2c = d2n; hd = nhd[j]; aj = na_[j]; }
    else { const double2 op = opos[j]; dxc = op.x - xi; dyc = op.y - yi; d2c = dxc * dxc + dyc * dyc; hd = ohd[j]; aj = oa[j]; }
    const bool hit_c = d2c <= P.s_dc_le;
    if (MASKS) { B.comm_mask[mrow_u + j] = hit_c; B.nbr_mask[mrow_u + j] = hit_nbr; B.dup_mask[mrow_u + j] = hit_dup; }
    if (hit_c) {
      double rx = dxc / P.dc, ry = dyc / P.dc, vx = hd.x - chi, vy = hd.y - shi;
      double da = (double)(aj - ai) / (double)P.na;
      const double wx = rx - xi, wy = ry - yi, w2 = wx * wx + wy * wy;
      if (w2 < 1.0) { const double w = sqrt(w2); rx /= w; ry /= w; vx /= w; vy /= w; da /= w; }
      c0 += rx; c1 += ry; c2 += vx; c3 += vy; c4 += da;
      ncomm++;
    }
  }
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
  O.tt = tt; O.dup = dup;
  O.nb[0] = nb[0]; O.nb[1] = nb[1]; O.nb[2] = nb[2]; O.nb[3] = nb[3];
}

// ------------------------------------------------------------------------------------------------
// fast path: same masks, linear sums, fp32 transcendentals.
// CN / CM: compile-time n_uav / m_targets (0 = run-time).  WARP_ENV: every warp lies inside one
// environment (n % 32 == 0), so "j already moved" (j < i) is warp-uniform outside the warp's own 32 UAVs.
// ------------------------------------------------------------------------------------------------
struct CommAcc {
  double sx, sy, sc, ss;
  int sa, cnt;
};

enum { PAIR_MOVED = 0, PAIR_MIXED = 1, PAIR_UNMOVED = 2 };

// One chunk of up to 32 partner UAVs j = jb .. jb+len-1 for UAV i.
//   PAIR_MOVED    every j moved before i (j < i): one distance serves the reward tests and communication
//   PAIR_UNMOVED  every j moves after i: new-new distance for the rewards, new-old for communication
//   PAIR_MIXED    per-lane order (j in the same warp as i, or a run-time sized environment)
template <int KIND, bool MASKS>
__device__ __forceinline__ uint32_t pair_chunk(const KParams &P, const UavSimBuffers &B, int jb, int len, int i,
                                               double xi, double yi, const double2 *__restrict__ npos,
                                               const double2 *__restrict__ nhd, const double2 *__restrict__ opos,
                                               const double2 *__restrict__ ohd, const int *__restrict__ na_,
                                               const int *__restrict__ oa, float k_ex0, float k_ex1, CommAcc &A,
                                               float &dupA, float &dupB, int64_t mrow_u) {
  uint32_t bits = 0, bit = 1;
  auto body = [&](int j, float &dupacc) {
    const double2 np = npos[j];
    const double dxn = np.x - xi, dyn = np.y - yi;
    const double d2n = dxn * dxn + dyn * dyn;
    const bool valid = (KIND != PAIR_MIXED) || (j != i);
    const bool hd = valid && (d2n <= P.s_2dp_le);  // uav.py:225
    const bool hn = valid && (d2n <= P.s_dp_le);   // uav.py:305
    // exp((2dp - d)/(2dp)) = 2^(log2e - d*log2e/(2dp))
    const float v = fast_ex2f(fmaf(fast_sqrtf((float)d2n), k_ex1, k_ex0));
    dupacc += hd ? v : 0.f;
    if (hn) bits |= bit;
    bit += bit;
    bool hc;
    if (KIND == PAIR_MOVED) {
      hc = d2n <= P.s_dc_le;  // uav.py:135, partner already at its new state
      if (hc) { const double2 h = nhd[j]; A.sx += dxn; A.sy += dyn; A.sc += h.x; A.ss += h.y; A.sa += na_[j]; A.cnt++; }
    } else {
      const double2 op = opos[j];
      const double dxo = op.x - xi, dyo = op.y - yi;
      const double d2o = dxo * dxo + dyo * dyo;
      if (KIND == PAIR_UNMOVED) {
        hc = d2o <= P.s_dc_le;  // partner still at its old state
        if (hc) { const double2 h = ohd[j]; A.sx += dxo; A.sy += dyo; A.sc += h.x; A.ss += h.y; A.sa += oa[j]; A.cnt++; }
      } else {
        const bool hc_new = (j < i) && (d2n <= P.s_dc_le);
        const bool hc_old = (j > i) && (d2o <= P.s_dc_le);
        if (hc_new) { const double2 h = nhd[j]; A.sx += dxn; A.sy += dyn; A.sc += h.x; A.ss += h.y; A.sa += na_[j]; A.cnt++; }
        if (hc_old) { const double2 h = ohd[j]; A.sx += dxo; A.sy += dyo; A.sc += h.x; A.ss += h.y; A.sa += oa[j]; A.cnt++; }
        hc = hc_new || hc_old;
      }
    }
    if (MASKS) { B.comm_mask[mrow_u + j] = hc; B.nbr_mask[mrow_u + j] = hn; B.dup_mask[mrow_u + j] = hd; }
  };
  int j = jb;
  const int jend = jb + len;
#pragma unroll 2
  for (; j + 1 < jend; j += 2) { body(j, dupA); body(j + 1, dupB); }
  if (j < jend) body(j, dupA);
  return bits;
}

template <int CN, int CM, bool WARP_ENV, bool MASKS>
__device__ __forceinline__ void agent_fast(const KParams &P, const UavSimBuffers &B, double xi, double yi, double chi,
                                           double shi, int ai, int i, int n_rt, int m_rt,
                                           const double2 *__restrict__ tpos, const double2 *__restrict__ tvel,
                                           const double2 *__restrict__ npos, const double2 *__restrict__ nhd,
                                           const double2 *__restrict__ opos, const double2 *__restrict__ ohd,
                                           const int *__restrict__ na_, const int *__restrict__ oa, int *tcnt,
                                           float *ob, int64_t mrow_t, int64_t mrow_u, AgentOut &O) {
  const int n = CN ? CN : n_rt, m = CM ? CM : m_rt;
  const float inv_dp_f = (float)(1.0 / P.dp);
  const float k_ex0 = 1.4426950408889634f, k_ex1 = (float)(-1.4426950408889634 / P.two_dp);

  // ---- targets: observe_target (uav.py:101-122), tracking reward (uav.py:199-212), coverage (environment.py:246-253)
  double ox = 0, oy = 0, ovx = 0, ovy = 0;
  float ttf = 0.f;
  int nobs = 0;
#pragma unroll 4
  for (int t = 0; t < m; t++) {
    const double2 tp = tpos[t];
    const double dx = tp.x - xi, dy = tp.y - yi;
    const double d2 = dx * dx + dy * dy;
    const bool hit = d2 <= P.s_dp_le;
    if (MASKS) { B.obs_mask[mrow_t + t] = hit; B.cover_mask[mrow_t + t] = (d2 <= P.s_dp_lt); }
    if (hit) {
      const double2 tv = tvel[t];
      ox += dx; oy += dy; ovx += tv.x; ovy += tv.y;
      nobs++;
      ttf += 2.0f - fast_sqrtf((float)d2) * inv_dp_f;  // 1 + (dp - d)/dp
      if (d2 <= P.s_dp_lt) atomicAdd(&tcnt[t], 1);
    }
  }

  // ---- UAVs: observe_uav in the sequential update order (uav.py:124-147, environment.py:133-138),
  //      duplicate-tracking punishment (uav.py:214-229), neighbour set (uav.py:305)
  CommAcc A = {0, 0, 0, 0, 0, 0};
  float dupA = 0.f, dupB = 0.f;  // two partial sums keep the fp32 accumulation error ~1e-7 after normalisation
  uint32_t nb[4] = {0, 0, 0, 0};
  const int w0 = i & ~31;  // first UAV of this warp (WARP_ENV)
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const int jb = 32 * c;
    if (jb < n) {
      const int len = min(32, n - jb);
      if (WARP_ENV && jb < w0)
        nb[c] = pair_chunk<PAIR_MOVED, MASKS>(P, B, jb, len, i, xi, yi, npos, nhd, opos, ohd, na_, oa, k_ex0, k_ex1, A, dupA, dupB, mrow_u);
      else if (WARP_ENV && jb > w0)
        nb[c] = pair_chunk<PAIR_UNMOVED, MASKS>(P, B, jb, len, i, xi, yi, npos, nhd, opos, ohd, na_, oa, k_ex0, k_ex1, A, dupA, dupB, mrow_u);
      else
        nb[c] = pair_chunk<PAIR_MIXED, MASKS>(P, B, jb, len, i, xi, yi, npos, nhd, opos, ohd, na_, oa, k_ex0, k_ex1, A, dupA, dupB, mrow_u);
    }
  }

  // ---- 12-d local state (uav.py:156-190): means of the lists (row weights are all 1 here), -1 blocks when empty
  if (A.cnt) {
    const double k = (double)A.cnt;
    ob[0] = (float)(A.sx / P.dc / k);
    ob[1] = (float)(A.sy / P.dc / k);
    ob[2] = (float)((A.sc - k * chi) / k);
    ob[3] = (float)((A.ss - k * shi) / k);
    ob[4] = (float)((double)(A.sa - A.cnt * ai) / (double)P.na / k);
  } else {
    ob[0] = ob[1] = ob[2] = ob[3] = ob[4] = -1.f;
  }
  if (nobs) {
    const double k = (double)nobs;
    ob[5] = (float)(ox / P.dp / k);
    ob[6] = (float)(oy / P.dp / k);
    ob[7] = (float)((ovx - k * chi) / k);
    ob[8] = (float)((ovy - k * shi) / k);
  } else {
    ob[5] = ob[6] = ob[7] = ob[8] = -1.f;
  }
  O.tt = (double)ttf;
  O.dup = -0.5 * ((double)dupA + (double)dupB);
  O.nb[0] = nb[0]; O.nb[1] = nb[1]; O.nb[2] = nb[2]; O.nb[3] = nb[3];
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int CN, int CM, bool MASKS>
__global__ void __launch_bounds__(NT, 2)
uavsim_step_kernel(const KParams P, const UavSimBuffers B, const double *__restrict__ g_dth, int64_t env_begin,
                   int64_t env_count, int epb, int mode, double coop, int done_flag,
                   double *__restrict__ stats_partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = CN ? CN : P.n, m = CM ? CM : P.m;
  constexpr bool WARP_ENV = (CN > 0) && (CN % 32 == 0);
  const int tid = threadIdx.x;
  const StepSmem S = carve(smem_raw, n, m, P.na, epb);
  const int64_t plane = P.E * n;  // rew4 plane stride

  for (int k = tid; k < 3 * P.na; k += NT) S.dth[k] = g_dth[k];

  const int64_t ngroups = (env_count + epb - 1) / epb;
  // per-thread statistics, reduced once at the end (src/train.py:181-192)
  double st_r = 0, st_tt = 0, st_bp = 0, st_dup = 0, st_cov = 0, st_envs = 0;
  int st_cmax = 0;

  for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int64_t e0 = env_begin + grp * epb;
    const int ne = (int)min((int64_t)epb, env_begin + env_count - e0);
    __syncthreads();  // previous iteration's readers are done; dth table visible

    // ---- phase 0a: targets (src/agent/target.py:27-60) ----
    for (int q = tid; q < ne * m; q += NT) {
      const int64_t gi = e0 * m + q;
      double x = B.tx[gi], y = B.ty[gi], h = B.th[gi];
      double sh, ch;
      sincos(h, &sh, &ch);
      x += P.dtv_t * ch;
      y += P.dtv_t * sh;
      bool refl = false;
      if (0 > y || y > P.y_max) {
        h = -h; refl = true;
      } else if (x < 0 || x > P.x_max) {
        h = (h > 0) ? (PI_D - h) : (-PI_D - h); refl = true;
      }
      if (refl) { sincos(h, &sh, &ch); B.th[gi] = h; }
      B.tx[gi] = x; B.ty[gi] = y;
      S.tpos[q] = make_double2(x, y);
      // cos(target.h) * target.v_max / self.v_max  (src/agent/uav.py:115-116)
      S.tvel[q] = make_double2(ch * P.tv / P.uv, sh * P.tv / P.uv);
      S.tcnt[q] = 0;
    }
    // ---- phase 0b: UAV kinematics (src/agent/uav.py:73-99) ----
    const int q = tid;
    const bool active = q < ne * n;
    const int el = active ? q / n : 0, i = q - el * n;
    const int64_t ge = e0 + el, gi = e0 * n + q;
    if (active) {
      double x = B.ux[gi], y = B.uy[gi], h = B.uh[gi];
      const int a_old = B.ua[gi], act = B.actions[gi];
      double sh, ch;
      sincos(h, &sh, &ch);
      S.opos[q] = make_double2(x, y); S.ohd[q] = make_double2(ch, sh); S.oa[q] = a_old;
      x += P.dtv_u * ch;
      y += P.dtv_u * sh;
      h += S.dth[3 * act];
      h = pymod_pos(h + PI_D, 2 * PI_D) - PI_D;
      // cos/sin of the new heading by angle addition (|error| ~ 3e-16; they only feed the observation).
      // The next step re-evaluates sincos from the stored heading, so the trajectory is unaffected.
      const double cd = S.dth[3 * act + 1], sd = S.dth[3 * act + 2];
      const double chn = ch * cd - sh * sd, shn = sh * cd + ch * sd;
      S.npos[q] = make_double2(x, y); S.nhd[q] = make_double2(chn, shn); S.na_[q] = act;
      B.ux[gi] = x; B.uy[gi] = y; B.uh[gi] = h; B.ua[gi] = act;
    }
    __syncthreads();

    // ---- phase 1: all-pairs tests, observation, raw reward ----
    AgentOut O;
    O.tt = 0; O.dup = 0; O.nb[0] = O.nb[1] = O.nb[2] = O.nb[3] = 0;
    double raw = 0, ttn = 0, bpn = 0, dupn = 0;
    if (active) {
      const double2 me = S.npos[q], mh = S.nhd[q];
      const double xi = me.x, yi = me.y, chi = mh.x, shi = mh.y;
      const int ai = S.na_[q];
      float *ob = S.obs + (size_t)q * 12;
      const int64_t mrow_t = (ge * n + i) * m, mrow_u = (ge * n + i) * n;
      // row weights differ from 1 only if |x| < 2 and |y| < 2 (|rx|,|ry| <= 1 for any row in range)
      const bool near_origin = fabs(xi) < 2.0 && fabs(yi) < 2.0;
      if (near_origin)
        agent_exact<MASKS>(P, B, xi, yi, chi, shi, ai, i, n, m, S.tpos + el * m, S.tvel + el * m, S.npos + el * n,
                           S.nhd + el * n, S.opos + el * n, S.ohd + el * n, S.na_ + el * n, S.oa + el * n,
                           S.tcnt + el * m, ob, mrow_t, mrow_u, O);
      else
        agent_fast<CN, CM, WARP_ENV, MASKS>(P, B, xi, yi, chi, shi, ai, i, n, m, S.tpos + el * m, S.tvel + el * m,
                                            S.npos + el * n, S.nhd + el * n, S.opos + el * n, S.ohd + el * n,
                                            S.na_ + el * n, S.oa + el * n, S.tcnt + el * m, ob, mrow_t, mrow_u, O);
      if (MASKS) { B.comm_mask[mrow_u + i] = 0; B.nbr_mask[mrow_u + i] = 0; B.dup_mask[mrow_u + i] = 0; }
      ob[9] = (float)(xi / P.dc);
      ob[10] = (float)(yi / P.dc);
      ob[11] = (float)((double)ai / (double)P.na);

      // boundary punishment (uav.py:231-250)
      const double dbdr = fmin(fmin(xi - 0, P.x_max - xi), fmin(yi - 0, P.y_max - yi));
      double bp;
      if (0 <= xi && xi <= P.x_max && 0 <= yi && yi <= P.y_max)
        bp = (dbdr < P.dp) ? (-0.5 * (P.dp - dbdr) / P.dp) : 0.0;
      else
        bp = -0.5;
      // normalise + weights (environment.py:206-220)
      ttn = clipnorm_0(O.tt, P.tt_hi);
      dupn = clipnorm_m1(O.dup, P.dup_lo);
      bpn = clipnorm_m1(bp, -0.5);
      raw = P.alpha * ttn + P.beta * bpn + P.gamma * dupn;
      S.raw[q] = raw;
    }
    __syncthreads();

    // ---- phase 2: cooperative reward (environment.py:222-227), coverage, outputs ----
    if (active) {
      double r;
      const bool pmi_pending = (mode == UAVSIM_MODE_PMI) && (coop != 0.0);
      if (mode == UAVSIM_MODE_SELF || coop == 0.0) {
        r = raw;  // uav.py:271-272 / :300-301
      } else if (mode == UAVSIM_MODE_MEAN) {
        // uav.py:293-310 -- the conditional expression covers the whole sum: no neighbour -> 0
        const double *R = S.raw + el * n;
        double s = 0;
        int cnt = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          uint32_t w = O.nb[c];
          while (w) { const int j = __ffs((int)w) - 1; w &= w - 1; s += R[32 * c + j]; cnt++; }
        }
        r = cnt ? ((1 - coop) * raw + coop * s / (double)cnt) : 0.0;
      } else {
        r = 0.0;  // finished by uavsim_pmi_kernel
        B.raw[gi] = raw;
        B.nbr_bits[gi * 2] = (uint64_t)O.nb[0] | ((uint64_t)O.nb[1] << 32);
        B.nbr_bits[gi * 2 + 1] = (uint64_t)O.nb[2] | ((uint64_t)O.nb[3] << 32);
      }
      r = fmin(fmax(r, -1.0), 1.0);  // clip_and_normalize(reward, -1, 1) is a plain clip
      if (!pmi_pending) { B.rew4[gi] = (float)r; st_r += r; }
      B.rew4[plane + gi] = (float)ttn;
      B.rew4[2 * plane + gi] = (float)bpn;
      B.rew4[3 * plane + gi] = (float)dupn;
      st_tt += ttn; st_bp += bpn; st_dup += dupn;
    }
    if (tid < ne) {  // environment.py:246-253: targets with at least one UAV strictly within dp
      int c = 0;
      const int *tc = S.tcnt + tid * m;
      for (int t = 0; t < m; t++) c += (tc[t] > 0);
      B.covered[e0 + tid] = c;
      if (B.done) B.done[e0 + tid] = done_flag;
      st_cov += (double)c;
      st_cmax = max(st_cmax, c);
      st_envs += 1.0;
    }
    if (B.tracker_cnt)
      for (int k = tid; k < ne * m; k += NT) B.tracker_cnt[e0 * m + k] = S.tcnt[k];
    {  // coalesced observation write: ne*n*12 floats = ne*n*3 float4, contiguous in global memory
      const float4 *src = reinterpret_cast<const float4 *>(S.obs);
      float4 *dst = reinterpret_cast<float4 *>(B.obs + e0 * n * 12);
      for (int k = tid; k < ne * n * 3; k += NT) dst[k] = src[k];
    }
  }
  block_stats_commit(S.red, stats_partial + (size_t)blockIdx.x * STAT_W, st_r, st_tt, st_bp, st_dup, st_cov, st_cmax,
                     st_envs, NT);
}

// uavsim.cu -- libuavsim.so: batched UAV/target tracking environment for B200 (sm_100a).
//
// One fused step kernel per block of environments: target motion, UAV kinematics, the all-pairs
// UAV-target / UAV-UAV range tests (positions staged in shared memory), the 12-d weighted-mean
// observation, the three reward terms, the cooperative reward and the coverage count.
// Semantics follow the reference's Environment.step (src/environment.py:120-164) and the UAV/TARGET
// classes (src/agent/uav.py, src/agent/target.py); each section cites the lines it implements.
//
// Numerics: state, distances and every range test are fp64 (the masks must equal the reference's,
// SURVEY.md section 7), compiled with -fmad=false so the parity-critical expressions keep the
// reference's evaluation order.  Range tests compare squared distances against the exact squared
// threshold computed on the host (largest double s with sqrt(s) <= thr), and take sqrt only on hits.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/uavsim.h"
#include "philox.cuh"

#define NT 256       // threads per CTA, step kernel (one thread per UAV, NT/n environments per CTA)
#define PMI_NT 256   // threads per CTA, PMI kernel
#define STAT_W 8     // doubles per statistics slot

static thread_local char g_err[512] = "";
#define SET_ERR(...) snprintf(g_err, sizeof(g_err), __VA_ARGS__)
#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      SET_ERR("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                      \
    }                                                                                      \
  } while (0)

// ------------------------------------------------------------------------------------------------
// kernel-side parameter block
// ------------------------------------------------------------------------------------------------
struct KParams {
  int n, m, na, num_steps;
  int64_t E;                  // environments of this handle (plane stride of rew4 is E*n)
  int64_t env_id_offset;      // global id of env 0 (RNG key only)
  double x_max, y_max;
  double dtv_u, dtv_t;        // dt*v_max of UAVs / targets (Python evaluates dt*v_max first)
  double dc, dp, two_dp;      // two_dp = radio*dp, radio = 2 (src/agent/uav.py:214)
  double tv, uv;              // target / uav v_max
  double s_dp_le, s_dp_lt, s_dc_le, s_2dp_le;  // exact squared thresholds
  double alpha, beta, gamma;
  double tt_hi;               // 2*m_targets            (src/environment.py:207-208)
  double dup_lo;              // -e/2*n_uav             (src/environment.py:209-210)
};

struct PmiDev {
  int H;
  const float *w0, *b0, *w1t, *b1, *w2;  // w1t = fc1 weight transposed to [3H,H]
  float b2;
};

struct uavsim {
  UavSimParams hp;
  KParams kp;
  UavSimBuffers buf;
  bool bound;
  int device, sm_count;
  int64_t E;
  double *d_dth;      // [na] dt * heading-rate per action (src/agent/uav.py:73-81,96)
  double *d_stats;    // [2][slots][STAT_W] per-CTA partial sums (step kernel | pmi kernel)
  double *d_stats8;   // [8] reduced
  double *h_stats8;   // pinned
  int stat_slots;
  int epb, grid_max;
  size_t smem_step;
  // pmi
  bool has_pmi;
  PmiDev pmi;
  float *d_pmi_blob;
  int pmi_g, pmi_pmax, pmi_tm, pmi_grid_max;
  size_t smem_pmi;
  // host-buffer pipeline
  cudaStream_t s_in, s_comp, s_out;
  cudaEvent_t ev_user, ev_in[16], ev_comp[16];
  int64_t t, launches;
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
#define PI_D 3.141592653589793

// Python float `%` with a positive divisor (CPython float_rem): fmod, then shift negatives up.
__device__ __forceinline__ double pymod_pos(double a, double b) {
  double r = fmod(a, b);
  if (r < 0.0) r += b;
  return r;
}

// src/utils/data_util.py:43-56 clip_and_normalize, choice 0 with floor 0
__device__ __forceinline__ double clipnorm_0(double v, double hi) {
  v = fmin(fmax(v, 0.0), hi);
  return (v - 0.0) / (hi - 0.0);
}
// choice -1 with ceil 0: (v-floor)/(0-floor) - 1
__device__ __forceinline__ double clipnorm_m1(double v, double lo) {
  v = fmin(fmax(v, lo), 0.0);
  return (v - lo) / (0.0 - lo) - 1.0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reduction of per-thread statistics into this CTA's slot (accumulating across launches;
// one writer per slot, no atomics, so the totals are reproducible for a fixed launch geometry).
__device__ void block_stats_commit(double *red /*smem [nwarps*6]*/, double *slot, double v0, double v1, double v2,
                                   double v3, double v4, int vmax, double v6, int nthreads) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = nthreads >> 5;
  v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3); v4 = warp_sum(v4);
  v6 = warp_sum(v6);
  vmax = warp_max(vmax);
  __syncthreads();
  if (lane == 0) {
    red[wid * 7 + 0] = v0; red[wid * 7 + 1] = v1; red[wid * 7 + 2] = v2; red[wid * 7 + 3] = v3;
    red[wid * 7 + 4] = v4; red[wid * 7 + 5] = (double)vmax; red[wid * 7 + 6] = v6;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int w = 0; w < nw; w++) {
      for (int k = 0; k < 5; k++) a[k] += red[w * 7 + k];
      a[5] = fmax(a[5], red[w * 7 + 5]);
      a[6] += red[w * 7 + 6];
    }
    for (int k = 0; k < 5; k++) slot[k] += a[k];
    slot[5] = fmax(slot[5], a[5]);
    slot[6] += a[6];
  }
}

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up of the step kernel (per CTA, epb environments)
// ------------------------------------------------------------------------------------------------
struct StepSmem {
  double *tx, *ty, *tvx, *tvy;  // [epb*m] moved targets, cos/sin(h)*tv/uv
  double *ox, *oy, *och, *osh;  // [epb*n] UAV before the move
  double *nx, *ny, *nch, *nsh;  // [epb*n] UAV after the move
  double *raw;                  // [epb*n]
  double *dth;                  // [na]
  double *red;                  // [64]
  float *obs;                   // [epb*n*12] (16-byte aligned)
  int *oa, *na_;                // [epb*n] previous / new action index
  int *tcnt;                    // [epb*m] UAVs strictly within dp of each target
};

static size_t step_smem_bytes(int n, int m, int na, int epb) {
  size_t d = (size_t)epb * m * 4 + (size_t)epb * n * 9 + (size_t)na + 64;
  size_t f = (size_t)epb * n * 12;
  size_t i = (size_t)epb * n * 2 + (size_t)epb * m;
  return d * 8 + 16 + f * 4 + i * 4;
}

__device__ __forceinline__ StepSmem carve(unsigned char *base, int n, int m, int na, int epb) {
  StepSmem s;
  double *d = reinterpret_cast<double *>(base);
  const size_t em = (size_t)epb * m, en = (size_t)epb * n;
  s.tx = d; d += em; s.ty = d; d += em; s.tvx = d; d += em; s.tvy = d; d += em;
  s.ox = d; d += en; s.oy = d; d += en; s.och = d; d += en; s.osh = d; d += en;
  s.nx = d; d += en; s.ny = d; d += en; s.nch = d; d += en; s.nsh = d; d += en;
  s.raw = d; d += en;
  s.dth = d; d += na;
  s.red = d; d += 64;
  uintptr_t p = (reinterpret_cast<uintptr_t>(d) + 15) & ~(uintptr_t)15;
  s.obs = reinterpret_cast<float *>(p);
  int *ip = reinterpret_cast<int *>(s.obs + en * 12);
  s.oa = ip; ip += en; s.na_ = ip; ip += en; s.tcnt = ip;
  return s;
}

// ------------------------------------------------------------------------------------------------
// the fused step kernel: CTA = epb environments, thread q = (env q/n, UAV q%n), epb*n <= NT
// ------------------------------------------------------------------------------------------------
template <bool MASKS>
__global__ void __launch_bounds__(NT)
uavsim_step_kernel(const KParams P, const UavSimBuffers B, const double *__restrict__ g_dth, int64_t env_begin,
                   int64_t env_count, int epb, int mode, double coop, int done_flag,
                   double *__restrict__ stats_partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = P.n, m = P.m;
  const int tid = threadIdx.x;
  const StepSmem S = carve(smem_raw, n, m, P.na, epb);
  const int64_t plane = P.E * n;  // rew4 plane stride

  for (int k = tid; k < P.na; k += NT) S.dth[k] = g_dth[k];

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
      S.tx[q] = x; S.ty[q] = y;
      // cos(target.h) * target.v_max / self.v_max  (src/agent/uav.py:115-116)
      S.tvx[q] = ch * P.tv / P.uv;
      S.tvy[q] = sh * P.tv / P.uv;
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
      S.ox[q] = x; S.oy[q] = y; S.och[q] = ch; S.osh[q] = sh; S.oa[q] = a_old;
      x += P.dtv_u * ch;
      y += P.dtv_u * sh;
      h += S.dth[act];
      h = pymod_pos(h + PI_D, 2 * PI_D) - PI_D;
      sincos(h, &sh, &ch);
      S.nx[q] = x; S.ny[q] = y; S.nch[q] = ch; S.nsh[q] = sh; S.na_[q] = act;
      B.ux[gi] = x; B.uy[gi] = y; B.uh[gi] = h; B.ua[gi] = act;
    }
    __syncthreads();

    // ---- phase 1: all-pairs tests, observation, raw reward ----
    uint64_t nb0 = 0, nb1 = 0;  // neighbour set d <= dp
    double raw = 0, ttn = 0, bpn = 0, dupn = 0;
    if (active) {
      const double xi = S.nx[q], yi = S.ny[q], chi = S.nch[q], shi = S.nsh[q];
      const int ai = S.na_[q];
      const double *Tx = S.tx + el * m, *Ty = S.ty + el * m, *Tvx = S.tvx + el * m, *Tvy = S.tvy + el * m;

      // targets: observe_target (uav.py:101-122), tracking reward (uav.py:199-212), coverage (environment.py:246-253)
      double tt = 0, o0 = 0, o1 = 0, o2 = 0, o3 = 0;
      int nobs = 0;
      for (int t = 0; t < m; t++) {
        const double dx = Tx[t] - xi, dy = Ty[t] - yi;
        const double d2 = dx * dx + dy * dy;
        const bool hit = d2 <= P.s_dp_le;
        const bool cov = d2 <= P.s_dp_lt;
        if (MASKS) {
          B.obs_mask[(ge * n + i) * m + t] = hit;
          B.cover_mask[(ge * n + i) * m + t] = cov;
        }
        if (hit) {
          const double d = sqrt(d2);
          tt += 1 + (P.dp - d) / P.dp;
          double rx = dx / P.dp, ry = dy / P.dp, vx = Tvx[t] - chi, vy = Tvy[t] - shi;
          // weight quirk: min(||(rx,ry) - (x,y)||, 1)  (uav.py:174-180)
          const double wx = rx - xi, wy = ry - yi, w2 = wx * wx + wy * wy;
          if (w2 < 1.0) {
            const double w = sqrt(w2);
            rx /= w; ry /= w; vx /= w; vy /= w;
          }
          o0 += rx; o1 += ry; o2 += vx; o3 += vy;
          nobs++;
          if (cov) atomicAdd(&S.tcnt[el * m + t], 1);
        }
      }

      // UAVs: observe_uav with the sequential update order (uav.py:124-147, environment.py:133-138),
      // duplicate-tracking punishment (uav.py:214-229) and the neighbour set (uav.py:305)
      const double *Nx = S.nx + el * n, *Ny = S.ny + el * n, *Nch = S.nch + el * n, *Nsh = S.nsh + el * n;
      const double *Ox = S.ox + el * n, *Oy = S.oy + el * n, *Och = S.och + el * n, *Osh = S.osh + el * n;
      const int *Na_ = S.na_ + el * n, *Oa = S.oa + el * n;
      double dup = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
      int ncomm = 0;
      for (int j = 0; j < n; j++) {
        if (j == i) {
          if (MASKS) {
            const int64_t o = (ge * n + i) * n + j;
            B.comm_mask[o] = 0; B.nbr_mask[o] = 0; B.dup_mask[o] = 0;
          }
          continue;
        }
        const double dxn = Nx[j] - xi, dyn = Ny[j] - yi;
        const double d2n = dxn * dxn + dyn * dyn;
        const bool hit_dup = d2n <= P.s_2dp_le;
        const bool hit_nbr = d2n <= P.s_dp_le;
        if (hit_dup) {
          const double d = sqrt(d2n);
          dup += -0.5 * exp((P.two_dp - d) / P.two_dp);
        }
        if (hit_nbr) {
          if (j < 64) nb0 |= (1ull << j);
          else nb1 |= (1ull << (j - 64));
        }
        double dxc, dyc, d2c, cj, sj;
        int aj;
        if (j < i) {  // j already moved
          dxc = dxn; dyc = dyn; d2c = d2n; cj = Nch[j]; sj = Nsh[j]; aj = Na_[j];
        } else {      // j not moved yet: old position, heading and action
          dxc = Ox[j] - xi; dyc = Oy[j] - yi; d2c = dxc * dxc + dyc * dyc;
          cj = Och[j]; sj = Osh[j]; aj = Oa[j];
        }
        const bool hit_c = d2c <= P.s_dc_le;
        if (MASKS) {
          const int64_t o = (ge * n + i) * n + j;
          B.comm_mask[o] = hit_c; B.nbr_mask[o] = hit_nbr; B.dup_mask[o] = hit_dup;
        }
        if (hit_c) {
          double rx = dxc / P.dc, ry = dyc / P.dc, vx = cj - chi, vy = sj - shi;
          double da = (double)(aj - ai) / (double)P.na;
          const double wx = rx - xi, wy = ry - yi, w2 = wx * wx + wy * wy;
          if (w2 < 1.0) {
            const double w = sqrt(w2);
            rx /= w; ry /= w; vx /= w; vy /= w; da /= w;
          }
          c0 += rx; c1 += ry; c2 += vx; c3 += vy; c4 += da;
          ncomm++;
        }
      }

      // boundary punishment (uav.py:231-250)
      const double dbdr = fmin(fmin(xi - 0, P.x_max - xi), fmin(yi - 0, P.y_max - yi));
      double bp;
      if (0 <= xi && xi <= P.x_max && 0 <= yi && yi <= P.y_max)
        bp = (dbdr < P.dp) ? (-0.5 * (P.dp - dbdr) / P.dp) : 0.0;
      else
        bp = -0.5;

      // normalise + weights (environment.py:206-220)
      ttn = clipnorm_0(tt, P.tt_hi);
      dupn = clipnorm_m1(dup, P.dup_lo);
      bpn = clipnorm_m1(bp, -0.5);
      raw = P.alpha * ttn + P.beta * bpn + P.gamma * dupn;
      S.raw[q] = raw;

      // 12-d local state (uav.py:156-190): means of the weighted lists, -1 blocks when empty
      float *ob = S.obs + (size_t)q * 12;
      if (ncomm) {
        const double k = (double)ncomm;
        ob[0] = (float)(c0 / k); ob[1] = (float)(c1 / k); ob[2] = (float)(c2 / k);
        ob[3] = (float)(c3 / k); ob[4] = (float)(c4 / k);
      } else {
        ob[0] = ob[1] = ob[2] = ob[3] = ob[4] = -1.f;
      }
      if (nobs) {
        const double k = (double)nobs;
        ob[5] = (float)(o0 / k); ob[6] = (float)(o1 / k); ob[7] = (float)(o2 / k); ob[8] = (float)(o3 / k);
      } else {
        ob[5] = ob[6] = ob[7] = ob[8] = -1.f;
      }
      ob[9] = (float)(xi / P.dc);
      ob[10] = (float)(yi / P.dc);
      ob[11] = (float)((double)ai / (double)P.na);
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
        uint64_t w = nb0;
        while (w) { const int j = __ffsll((long long)w) - 1; w &= w - 1; s += R[j]; cnt++; }
        w = nb1;
        while (w) { const int j = __ffsll((long long)w) - 1; w &= w - 1; s += R[64 + j]; cnt++; }
        r = cnt ? ((1 - coop) * raw + coop * s / (double)cnt) : 0.0;
      } else {
        r = 0.0;  // finished by uavsim_pmi_kernel
        B.raw[gi] = raw;
        B.nbr_bits[gi * 2] = nb0;
        B.nbr_bits[gi * 2 + 1] = nb1;
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

// ------------------------------------------------------------------------------------------------
// PMI reciprocal reward (src/agent/uav.py:262-291 + src/models/PMINet.py:41-72), fp32 CUDA-core GEMM.
// CTA = G environments: enumerate neighbour pairs, run the folded MLP over tiles of TM pair rows
// (layer 0 block-diagonal 12->3H, layer 1 3H->H as a register-tiled SGEMM with fc1^T streamed
// through shared memory, layer 2 H->1 as a shuffle reduction), then the per-UAV softmax and mix.
// ------------------------------------------------------------------------------------------------
#define PMI_KC 32  // k-chunk of fc1 streamed per iteration

struct PmiSmem {
  float *obs;       // [G*n*12]
  double *raw;      // [G*n]
  double *red;      // [64]
  uint32_t *off;    // [G*n+1]
  uint32_t *pair;   // [pmax]  (a << 16) | b, local UAV indices in the group
  float *logit;     // [pmax]
  float *x;         // [TM*12]
  float *h0;        // [TM*(3H+4)]
  float *wchunk;    // [PMI_KC*H]
  float *w0, *b0, *b1, *w2;  // [3H*5] [3H] [H] [H]
};

static size_t pmi_smem_bytes(int n, int H, int G, int pmax, int TM) {
  size_t b = 0;
  b += (size_t)G * n * 8 + 64 * 8;                 // raw, red (doubles first)
  b += (size_t)G * n * 12 * 4;                     // obs
  b += ((size_t)G * n + 1 + 3) / 4 * 4 * 4;        // off (padded)
  b += (size_t)pmax * 8;                           // pair + logit
  b += (size_t)TM * 12 * 4 + (size_t)TM * (3 * H + 4) * 4 + (size_t)PMI_KC * H * 4;
  b += (size_t)(3 * H * 5 + 3 * H + H + H) * 4;
  return b + 32;
}

__device__ __forceinline__ PmiSmem pmi_carve(unsigned char *base, int n, int H, int G, int pmax, int TM) {
  PmiSmem s;
  double *d = reinterpret_cast<double *>(base);
  s.raw = d; d += (size_t)G * n;
  s.red = d; d += 64;
  float *f = reinterpret_cast<float *>(d);
  s.obs = f; f += (size_t)G * n * 12;
  s.off = reinterpret_cast<uint32_t *>(f); f += ((size_t)G * n + 1 + 3) / 4 * 4;
  s.pair = reinterpret_cast<uint32_t *>(f); f += pmax;
  s.logit = f; f += pmax;
  s.x = f; f += (size_t)TM * 12;
  s.h0 = f; f += (size_t)TM * (3 * H + 4);
  s.wchunk = f; f += (size_t)PMI_KC * H;
  s.w0 = f; f += 3 * H * 5;
  s.b0 = f; f += 3 * H;
  s.b1 = f; f += H;
  s.w2 = f; f += H;
  return s;
}

template <int CPT, int TM>  // H = 16*CPT output columns; TM rows per tile (TM/16 rows per thread)
__global__ void __launch_bounds__(PMI_NT)
uavsim_pmi_kernel(const KParams P, const UavSimBuffers B, const PmiDev W, int64_t env_begin, int64_t env_count,
                  int G, int pmax, double coop, double *__restrict__ stats_partial) {
  constexpr int H = 16 * CPT, H3 = 3 * H, LD0 = H3 + 4, RPT = TM / 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = P.n, tid = threadIdx.x;
  const PmiSmem S = pmi_carve(smem_raw, n, H, G, pmax, TM);

  for (int k = tid; k < H3 * 5; k += PMI_NT) S.w0[k] = W.w0[k];
  for (int k = tid; k < H3; k += PMI_NT) S.b0[k] = W.b0[k];
  for (int k = tid; k < H; k += PMI_NT) { S.b1[k] = W.b1[k]; S.w2[k] = W.w2[k]; }

  const int ty = tid >> 4, tx = tid & 15;
  const int64_t ngroups = (env_count + G - 1) / G;
  double st_r = 0;

  for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int64_t e0 = env_begin + grp * G;
    const int ne = (int)min((int64_t)G, env_begin + env_count - e0);
    const int A = ne * n;  // UAVs in this group (<= PMI_NT)
    __syncthreads();
    for (int k = tid; k < A * 12; k += PMI_NT) S.obs[k] = B.obs[e0 * n * 12 + k];
    uint64_t nb0 = 0, nb1 = 0;
    if (tid < A) {
      S.raw[tid] = B.raw[e0 * n + tid];
      nb0 = B.nbr_bits[(e0 * n + tid) * 2];
      nb1 = B.nbr_bits[(e0 * n + tid) * 2 + 1];
      S.off[tid + 1] = __popcll(nb0) + __popcll(nb1);
    }
    __syncthreads();
    if (tid == 0) {  // exclusive scan of the neighbour counts (A <= 256)
      uint32_t acc = 0;
      S.off[0] = 0;
      for (int a = 0; a < A; a++) { acc += S.off[a + 1]; S.off[a + 1] = acc; }
    }
    __syncthreads();
    const int npairs = (int)S.off[A];
    if (tid < A) {  // neighbours in ascending index order, like the reference's loop (uav.py:277-282)
      const int base = (tid / n) * n;
      uint32_t p = S.off[tid];
      uint64_t w = nb0;
      while (w) { const int j = __ffsll((long long)w) - 1; w &= w - 1; S.pair[p++] = ((uint32_t)tid << 16) | (uint32_t)(base + j); }
      w = nb1;
      while (w) { const int j = __ffsll((long long)w) - 1; w &= w - 1; S.pair[p++] = ((uint32_t)tid << 16) | (uint32_t)(base + 64 + j); }
    }
    __syncthreads();

    for (int p0 = 0; p0 < npairs; p0 += TM) {
      const int rows = min(TM, npairs - p0);
      // input rows: la_i * la_j (uav.py:280-281), fp32
      for (int k = tid; k < TM * 12; k += PMI_NT) {
        const int r = k / 12, c = k - r * 12;
        float v = 0.f;
        if (r < rows) {
          const uint32_t pr = S.pair[p0 + r];
          v = S.obs[(pr >> 16) * 12 + c] * S.obs[(pr & 0xffffu) * 12 + c];
        }
        S.x[k] = v;
      }
      __syncthreads();
      // layer 0: three branch Linear+BN(folded)+ReLU, concatenated (PMINet.py:45-58)
      for (int k = tid; k < TM * H3; k += PMI_NT) {
        const int r = k / H3, u = k - r * H3;
        const int b = u / H;
        const int off = (b == 0) ? 0 : (b == 1 ? 5 : 9);
        const int dim = (b == 0) ? 5 : (b == 1 ? 4 : 3);
        const float *xr = S.x + r * 12 + off, *wr = S.w0 + u * 5;
        float acc = S.b0[u];
        for (int c = 0; c < dim; c++) acc = fmaf(wr[c], xr[c], acc);
        S.h0[r * LD0 + u] = fmaxf(acc, 0.f);
      }
      // layer 1: [TM,3H] x [3H,H]
      float acc[RPT][CPT];
#pragma unroll
      for (int a = 0; a < RPT; a++)
#pragma unroll
        for (int c = 0; c < CPT; c++) acc[a][c] = 0.f;
      for (int kc = 0; kc < H3; kc += PMI_KC) {
        __syncthreads();  // h0 complete (first pass) / previous chunk consumed
        for (int k = tid; k < PMI_KC * H; k += PMI_NT) S.wchunk[k] = W.w1t[(size_t)kc * H + k];
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < PMI_KC; k++) {
          float av[RPT], bv[CPT];
#pragma unroll
          for (int a = 0; a < RPT; a++) av[a] = S.h0[(ty + 16 * a) * LD0 + kc + k];
#pragma unroll
          for (int c = 0; c < CPT; c++) bv[c] = S.wchunk[k * H + tx + 16 * c];
#pragma unroll
          for (int a = 0; a < RPT; a++)
#pragma unroll
            for (int c = 0; c < CPT; c++) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
        }
      }
      // bias + ReLU, then layer 2 (PMINet.py:59-62): dot with fc2 across the 16 column threads
#pragma unroll
      for (int a = 0; a < RPT; a++) {
        float part = 0.f;
#pragma unroll
        for (int c = 0; c < CPT; c++) {
          const int col = tx + 16 * c;
          part = fmaf(S.w2[col], fmaxf(acc[a][c] + S.b1[col], 0.f), part);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        const int r = ty + 16 * a;
        if (tx == 0 && r < rows) S.logit[p0 + r] = part + W.b2;
      }
      __syncthreads();
    }
    __syncthreads();

    // softmax over each UAV's neighbours (scipy.special.softmax on float32) and the mix (uav.py:284-290)
    if (tid < A) {
      const uint32_t lo = S.off[tid], hi = S.off[tid + 1];
      const double raw = S.raw[tid];
      double r;
      if (hi > lo) {
        float mx = S.logit[lo];
        for (uint32_t p = lo + 1; p < hi; p++) mx = fmaxf(mx, S.logit[p]);
        float ssum = 0.f;
        for (uint32_t p = lo; p < hi; p++) ssum += expf(S.logit[p] - mx);
        double s = 0;
        for (uint32_t p = lo; p < hi; p++) {
          const float wgt = expf(S.logit[p] - mx) / ssum;
          s += S.raw[S.pair[p] & 0xffffu] * (double)wgt;
        }
        r = (1 - coop) * raw + coop * s;
      } else {
        r = (1 - coop) * raw;
      }
      r = fmin(fmax(r, -1.0), 1.0);
      B.rew4[e0 * n + tid] = (float)r;
      st_r += r;
    }
  }
  block_stats_commit(S.red, stats_partial + (size_t)blockIdx.x * STAT_W, st_r, 0, 0, 0, 0, 0, 0, PMI_NT);
}

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

// fixed-order reduction of the per-CTA slots (single thread: <= a few thousand adds per episode)
__global__ void uavsim_stats_reduce_kernel(const double *__restrict__ partial, int slots, double *__restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double a[STAT_W] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int half = 0; half < 2; half++)
    for (int s = 0; s < slots; s++) {
      const double *p = partial + ((size_t)half * slots + s) * STAT_W;
      for (int k = 0; k < 5; k++) a[k] += p[k];
      a[5] = fmax(a[5], p[5]);
      a[6] += p[6];
    }
  for (int k = 0; k < STAT_W; k++) out[k] = a[k];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// largest double s with sqrt(s) <= thr (strict = false) or sqrt(s) < thr (strict = true); host sqrt and
// device sqrt are both correctly rounded, so `d2 <= s` reproduces `sqrt(d2) <= thr` / `< thr` exactly.
static double exact_sq_threshold(double thr, bool strict) {
  double s = thr * thr;
  auto ok = [&](double v) { double r = sqrt(v); return strict ? (r < thr) : (r <= thr); };
  while (!ok(s)) s = nextafter(s, -INFINITY);
  while (ok(nextafter(s, INFINITY))) s = nextafter(s, INFINITY);
  return s;
}

extern "C" int uavsim_abi_version(void) { return UAVSIM_ABI_VERSION; }
extern "C" const char *uavsim_last_error(void) { return g_err; }
extern "C" int64_t uavsim_launch_count(const uavsim_t *h) { return h ? h->launches : 0; }
extern "C" int64_t uavsim_step_count(const uavsim_t *h) { return h ? h->t : 0; }

extern "C" int uavsim_create(const UavSimParams *p, int64_t n_envs, int64_t env_id_offset, int device,
                             uavsim_t **out) {
  if (!p || !out) { SET_ERR("uavsim_create: NULL argument"); return UAVSIM_ERR_ARG; }
  *out = nullptr;
  if (n_envs <= 0 || p->n_uav <= 0 || p->m_targets <= 0 || p->na < 2 || p->dp <= 0 || p->dc <= 0) {
    SET_ERR("uavsim_create: need n_envs>0, n_uav>0, m_targets>0, na>=2, dp>0, dc>0");
    return UAVSIM_ERR_ARG;
  }
  if (p->n_uav > UAVSIM_MAX_UAV) {
    SET_ERR("uavsim_create: n_uav=%d exceeds UAVSIM_MAX_UAV=%d", p->n_uav, UAVSIM_MAX_UAV);
    return UAVSIM_ERR_UNSUPPORTED;
  }
  if (env_id_offset < 0 || env_id_offset + n_envs > (int64_t)0xffffffffll) {
    SET_ERR("uavsim_create: global env ids must fit 32 bits");
    return UAVSIM_ERR_UNSUPPORTED;
  }
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { SET_ERR("uavsim_create: device %d of %d", device, ndev); return UAVSIM_ERR_ARG; }
  CUDA_TRY(cudaSetDevice(device));
  uavsim *h = (uavsim *)calloc(1, sizeof(uavsim));
  if (!h) { SET_ERR("out of host memory"); return UAVSIM_ERR_ARG; }
  h->hp = *p;
  h->device = device;
  h->E = n_envs;
  CUDA_TRY(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));

  KParams &k = h->kp;
  k.n = p->n_uav; k.m = p->m_targets; k.na = p->na; k.num_steps = p->num_steps;
  k.E = n_envs; k.env_id_offset = env_id_offset;
  k.x_max = p->x_max; k.y_max = p->y_max;
  k.dtv_u = p->dt * p->uav_v_max;
  k.dtv_t = p->dt * p->tgt_v_max;
  k.dc = p->dc; k.dp = p->dp; k.two_dp = 2 * p->dp;
  k.tv = p->tgt_v_max; k.uv = p->uav_v_max;
  k.s_dp_le = exact_sq_threshold(p->dp, false);
  k.s_dp_lt = exact_sq_threshold(p->dp, true);
  k.s_dc_le = exact_sq_threshold(p->dc, false);
  k.s_2dp_le = exact_sq_threshold(2 * p->dp, false);
  k.alpha = p->alpha; k.beta = p->beta; k.gamma = p->gamma;
  k.tt_hi = (double)(2 * p->m_targets);
  k.dup_lo = -2.718281828459045 / 2 * p->n_uav;

  // dt * discrete_action(a) (src/agent/uav.py:73-81, :96), same evaluation order as the reference
  double *dth = (double *)malloc(sizeof(double) * p->na);
  for (int a = 0; a < p->na; a++) {
    const int na1 = a + 1;
    const double rate = (double)(2 * na1 - p->na - 1) * p->uav_h_max / (double)(p->na - 1);
    dth[a] = p->dt * rate;
  }
  CUDA_TRY(cudaMalloc(&h->d_dth, sizeof(double) * p->na));
  CUDA_TRY(cudaMemcpy(h->d_dth, dth, sizeof(double) * p->na, cudaMemcpyHostToDevice));
  free(dth);

  // launch geometry of the step kernel
  int epb = NT / p->n_uav;
  if (epb < 1) epb = 1;
  if ((int64_t)epb > n_envs) epb = (int)n_envs;
  while (epb > 1 && step_smem_bytes(k.n, k.m, k.na, epb) > 100 * 1024) epb--;
  h->epb = epb;
  h->smem_step = step_smem_bytes(k.n, k.m, k.na, epb);
  if (h->smem_step > 227 * 1024) {
    SET_ERR("uavsim_create: n_uav=%d m_targets=%d needs %zu B shared memory per CTA", k.n, k.m, h->smem_step);
    free(h);
    return UAVSIM_ERR_UNSUPPORTED;
  }
  // the attribute is per function, not per handle: only ever raise it (several handles may coexist)
  static size_t s_step_attr[64] = {0};
  if (h->smem_step > s_step_attr[device & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(uavsim_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_step));
    CUDA_TRY(cudaFuncSetAttribute(uavsim_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_step));
    s_step_attr[device & 63] = h->smem_step;
  }
  int occ = 1;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, uavsim_step_kernel<false>, NT, h->smem_step));
  if (occ < 1) occ = 1;
  h->grid_max = h->sm_count * occ;
  h->stat_slots = h->grid_max > 1024 ? h->grid_max : 1024;
  CUDA_TRY(cudaMalloc(&h->d_stats, sizeof(double) * 2 * h->stat_slots * STAT_W));
  CUDA_TRY(cudaMalloc(&h->d_stats8, sizeof(double) * 8));
  CUDA_TRY(cudaMallocHost(&h->h_stats8, sizeof(double) * 8));
  CUDA_TRY(cudaMemset(h->d_stats, 0, sizeof(double) * 2 * h->stat_slots * STAT_W));

  CUDA_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->s_comp, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_user, cudaEventDisableTiming));
  for (int c = 0; c < 16; c++) {
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_in[c], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_comp[c], cudaEventDisableTiming));
  }
  *out = h;
  return 0;
}

extern "C" int uavsim_destroy(uavsim_t *h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaFree(h->d_dth); cudaFree(h->d_stats); cudaFree(h->d_stats8); cudaFreeHost(h->h_stats8);
  cudaFree(h->d_pmi_blob);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_comp) cudaStreamDestroy(h->s_comp);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  if (h->ev_user) cudaEventDestroy(h->ev_user);
  for (int c = 0; c < 16; c++) {
    if (h->ev_in[c]) cudaEventDestroy(h->ev_in[c]);
    if (h->ev_comp[c]) cudaEventDestroy(h->ev_comp[c]);
  }
  free(h);
  return 0;
}

extern "C" int uavsim_bind(uavsim_t *h, const UavSimBuffers *b) {
  if (!h || !b) { SET_ERR("uavsim_bind: NULL argument"); return UAVSIM_ERR_ARG; }
  if (!b->ux || !b->uy || !b->uh || !b->ua || !b->tx || !b->ty || !b->th || !b->actions || !b->obs || !b->rew4 ||
      !b->covered) {
    SET_ERR("uavsim_bind: state, actions, obs, rew4 and covered buffers are required");
    return UAVSIM_ERR_ARG;
  }
  const int nm = (b->obs_mask != 0) + (b->comm_mask != 0) + (b->nbr_mask != 0) + (b->dup_mask != 0) + (b->cover_mask != 0);
  if (nm != 0 && nm != 5) { SET_ERR("uavsim_bind: give all five mask buffers or none"); return UAVSIM_ERR_ARG; }
  if ((reinterpret_cast<uintptr_t>(b->obs) & 15) != 0) { SET_ERR("uavsim_bind: obs must be 16-byte aligned"); return UAVSIM_ERR_ARG; }
  h->buf = *b;
  h->bound = true;
  return 0;
}

extern "C" int uavsim_set_reward_weights(uavsim_t *h, double alpha, double beta, double gamma) {
  if (!h) { SET_ERR("uavsim_set_reward_weights: NULL handle"); return UAVSIM_ERR_ARG; }
  h->hp.alpha = h->kp.alpha = alpha;
  h->hp.beta = h->kp.beta = beta;
  h->hp.gamma = h->kp.gamma = gamma;
  return 0;
}

static int check_bound(uavsim_t *h, const char *who) {
  if (!h) { SET_ERR("%s: NULL handle", who); return UAVSIM_ERR_ARG; }
  if (!h->bound) { SET_ERR("%s: uavsim_bind has not been called", who); return UAVSIM_ERR_UNBOUND; }
  return 0;
}

static int clear_counters(uavsim_t *h, cudaStream_t st) {
  const int count = 2 * h->stat_slots * STAT_W;
  uavsim_stats_clear_kernel<<<(count + 255) / 256, 256, 0, st>>>(h->d_stats, count);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  h->t = 0;
  return 0;
}

static int elementwise_grid(uavsim_t *h, int64_t items) {
  int64_t g = (items + 255) / 256;
  const int64_t cap = (int64_t)h->sm_count * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" int uavsim_reset(uavsim_t *h, uint64_t seed, void *stream) {
  int rc = check_bound(h, "uavsim_reset");
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t items = h->E * (h->kp.n + h->kp.m);
  uavsim_reset_kernel<<<elementwise_grid(h, items), 256, 0, st>>>(h->kp, h->buf, seed);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  return clear_counters(h, st);
}

extern "C" int uavsim_begin_episode(uavsim_t *h, void *stream) {
  int rc = check_bound(h, "uavsim_begin_episode");
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  uavsim_initial_obs_kernel<<<elementwise_grid(h, h->E * h->kp.n), 256, 0, st>>>(h->kp, h->buf);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  return clear_counters(h, st);
}

extern "C" int uavsim_random_actions(uavsim_t *h, uint64_t seed, int64_t step, void *stream) {
  int rc = check_bound(h, "uavsim_random_actions");
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  uavsim_random_actions_kernel<<<elementwise_grid(h, h->E * h->kp.n), 256, 0, st>>>(h->kp, h->buf.actions, seed,
                                                                                   (uint32_t)step);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  return 0;
}

// PMI launch geometry: G environments per CTA iteration, bounded by threads (G*n <= 256) and by the
// pair buffer (G*n*(n-1) <= pmax)
static int pmi_configure(uavsim_t *h) {
  const int n = h->kp.n, H = h->pmi.H;
  const int TM = (H <= 128) ? 64 : 32;
  int G = PMI_NT / n;
  if (G < 1) G = 1;
  if ((int64_t)G > h->E) G = (int)h->E;
  const int per_env = n * (n - 1) > 0 ? n * (n - 1) : 1;
  while (G > 1 && (int64_t)G * per_env > 8192) G--;
  int pmax = G * per_env;
  size_t smem = pmi_smem_bytes(n, H, G, pmax, TM);
  if (smem > 227 * 1024) {
    SET_ERR("PMI mode: n_uav=%d hidden=%d needs %zu B shared memory per CTA", n, H, smem);
    return UAVSIM_ERR_UNSUPPORTED;
  }
  h->pmi_g = G; h->pmi_pmax = pmax; h->pmi_tm = TM; h->smem_pmi = smem;
  return 0;
}

template <int CPT, int TM>
static int pmi_launch_t(uavsim_t *h, int64_t e0, int64_t cnt, double coop, cudaStream_t st, bool configure_only) {
  auto kern = uavsim_pmi_kernel<CPT, TM>;
  if (configure_only) {
    static size_t s_attr[64] = {0};  // per template instance and device: only ever raise it
    if (h->smem_pmi > s_attr[h->device & 63]) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_pmi));
      s_attr[h->device & 63] = h->smem_pmi;
    }
    int occ = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PMI_NT, h->smem_pmi));
    if (occ < 1) occ = 1;
    h->pmi_grid_max = h->sm_count * occ;
    if (h->pmi_grid_max > h->stat_slots) h->pmi_grid_max = h->stat_slots;
    return 0;
  }
  const int64_t ngroups = (cnt + h->pmi_g - 1) / h->pmi_g;
  const int grid = (int)(ngroups < h->pmi_grid_max ? ngroups : h->pmi_grid_max);
  kern<<<grid, PMI_NT, h->smem_pmi, st>>>(h->kp, h->buf, h->pmi, e0, cnt, h->pmi_g, h->pmi_pmax, coop,
                                         h->d_stats + (size_t)h->stat_slots * STAT_W);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  return 0;
}

static int pmi_launch(uavsim_t *h, int64_t e0, int64_t cnt, double coop, cudaStream_t st, bool configure_only) {
  switch (h->pmi.H) {
    case 32: return pmi_launch_t<2, 64>(h, e0, cnt, coop, st, configure_only);
    case 64: return pmi_launch_t<4, 64>(h, e0, cnt, coop, st, configure_only);
    case 128: return pmi_launch_t<8, 64>(h, e0, cnt, coop, st, configure_only);
    case 256: return pmi_launch_t<16, 32>(h, e0, cnt, coop, st, configure_only);
  }
  SET_ERR("PMI hidden size %d not supported (32, 64, 128, 256)", h->pmi.H);
  return UAVSIM_ERR_UNSUPPORTED;
}

extern "C" int uavsim_set_pmi_weights(uavsim_t *h, const UavSimPmiWeights *w, void *stream) {
  if (!h || !w || !w->w0 || !w->b0 || !w->w1 || !w->b1 || !w->w2) { SET_ERR("uavsim_set_pmi_weights: NULL argument"); return UAVSIM_ERR_ARG; }
  const int H = w->hidden;
  if (H != 32 && H != 64 && H != 128 && H != 256) { SET_ERR("PMI hidden size %d not supported (32, 64, 128, 256)", H); return UAVSIM_ERR_UNSUPPORTED; }
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_w0 = (size_t)3 * H * 5, n_b0 = 3 * H, n_w1 = (size_t)H * 3 * H, n_b1 = H, n_w2 = H;
  const size_t total = n_w0 + n_b0 + n_w1 + n_b1 + n_w2;
  float *blob = (float *)malloc(total * sizeof(float));
  float *q = blob;
  memcpy(q, w->w0, n_w0 * 4); q += n_w0;
  memcpy(q, w->b0, n_b0 * 4); q += n_b0;
  for (int k = 0; k < 3 * H; k++)  // transpose fc1 [H,3H] -> [3H,H]
    for (int o = 0; o < H; o++) q[(size_t)k * H + o] = w->w1[(size_t)o * 3 * H + k];
  q += n_w1;
  memcpy(q, w->b1, n_b1 * 4); q += n_b1;
  memcpy(q, w->w2, n_w2 * 4);
  if (h->d_pmi_blob && h->pmi.H != H) { CUDA_TRY(cudaStreamSynchronize(st)); cudaFree(h->d_pmi_blob); h->d_pmi_blob = nullptr; }
  if (!h->d_pmi_blob) CUDA_TRY(cudaMalloc(&h->d_pmi_blob, total * sizeof(float)));
  CUDA_TRY(cudaMemcpyAsync(h->d_pmi_blob, blob, total * sizeof(float), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  free(blob);
  h->pmi.H = H;
  h->pmi.w0 = h->d_pmi_blob;
  h->pmi.b0 = h->pmi.w0 + n_w0;
  h->pmi.w1t = h->pmi.b0 + n_b0;
  h->pmi.b1 = h->pmi.w1t + n_w1;
  h->pmi.w2 = h->pmi.b1 + n_b1;
  h->pmi.b2 = w->b2;
  int rc = pmi_configure(h);
  if (rc) return rc;
  rc = pmi_launch(h, 0, 0, 0.0, st, true);
  if (rc) return rc;
  h->has_pmi = true;
  return 0;
}

// one step over the env range [e0, e0+cnt) on stream st
static int launch_step_range(uavsim_t *h, int mode, double coop, int64_t e0, int64_t cnt, int done_flag, cudaStream_t st) {
  const int64_t ngroups = (cnt + h->epb - 1) / h->epb;
  const int grid = (int)(ngroups < h->grid_max ? ngroups : h->grid_max);
  if (h->buf.obs_mask)
    uavsim_step_kernel<true><<<grid, NT, h->smem_step, st>>>(h->kp, h->buf, h->d_dth, e0, cnt, h->epb, mode, coop, done_flag, h->d_stats);
  else
    uavsim_step_kernel<false><<<grid, NT, h->smem_step, st>>>(h->kp, h->buf, h->d_dth, e0, cnt, h->epb, mode, coop, done_flag, h->d_stats);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  if (mode == UAVSIM_MODE_PMI && coop != 0.0) return pmi_launch(h, e0, cnt, coop, st, false);
  return 0;
}

static int check_step_args(uavsim_t *h, int mode, double coop, const char *who) {
  int rc = check_bound(h, who);
  if (rc) return rc;
  if (mode != UAVSIM_MODE_SELF && mode != UAVSIM_MODE_MEAN && mode != UAVSIM_MODE_PMI) { SET_ERR("%s: bad mode %d", who, mode); return UAVSIM_ERR_ARG; }
  if (mode == UAVSIM_MODE_PMI && coop != 0.0) {
    if (!h->has_pmi) { SET_ERR("%s: mode PMI needs uavsim_set_pmi_weights first", who); return UAVSIM_ERR_NO_PMI; }
    if (!h->buf.raw || !h->buf.nbr_bits) { SET_ERR("%s: mode PMI needs the raw and nbr_bits buffers", who); return UAVSIM_ERR_UNBOUND; }
  }
  return 0;
}

extern "C" int uavsim_step(uavsim_t *h, int mode, double coop, void *stream) {
  int rc = check_step_args(h, mode, coop, "uavsim_step");
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  h->t++;
  const int done = (h->hp.num_steps > 0 && h->t >= h->hp.num_steps) ? 1 : 0;
  return launch_step_range(h, mode, coop, 0, h->E, done, (cudaStream_t)stream);
}

extern "C" int uavsim_step_host(uavsim_t *h, int mode, double coop, const int32_t *h_actions, float *h_obs,
                                float *h_rew4, int32_t *h_covered, int chunks, void *stream) {
  int rc = check_step_args(h, mode, coop, "uavsim_step_host");
  if (rc) return rc;
  if (!h_actions) { SET_ERR("uavsim_step_host: h_actions is NULL"); return UAVSIM_ERR_ARG; }
  CUDA_TRY(cudaSetDevice(h->device));
  if (chunks < 1) chunks = 1;
  if (chunks > 16) chunks = 16;
  if ((int64_t)chunks > h->E) chunks = (int)h->E;
  const int n = h->kp.n;
  const int64_t E = h->E;
  h->t++;
  const int done = (h->hp.num_steps > 0 && h->t >= h->hp.num_steps) ? 1 : 0;
  // order after whatever the caller queued on its stream
  CUDA_TRY(cudaEventRecord(h->ev_user, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamWaitEvent(h->s_in, h->ev_user, 0));
  CUDA_TRY(cudaStreamWaitEvent(h->s_comp, h->ev_user, 0));
  CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->ev_user, 0));
  for (int c = 0; c < chunks; c++) {
    const int64_t e0 = E * c / chunks, e1 = E * (c + 1) / chunks, cnt = e1 - e0;
    CUDA_TRY(cudaMemcpyAsync(h->buf.actions + e0 * n, h_actions + e0 * n, sizeof(int32_t) * cnt * n, cudaMemcpyHostToDevice, h->s_in));
    CUDA_TRY(cudaEventRecord(h->ev_in[c], h->s_in));
    CUDA_TRY(cudaStreamWaitEvent(h->s_comp, h->ev_in[c], 0));
    rc = launch_step_range(h, mode, coop, e0, cnt, done, h->s_comp);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev_comp[c], h->s_comp));
    CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->ev_comp[c], 0));
    if (h_obs)
      CUDA_TRY(cudaMemcpyAsync(h_obs + e0 * n * 12, h->buf.obs + e0 * n * 12, sizeof(float) * cnt * n * 12, cudaMemcpyDeviceToHost, h->s_out));
    if (h_rew4)
      for (int k = 0; k < 4; k++)
        CUDA_TRY(cudaMemcpyAsync(h_rew4 + (k * E + e0) * n, h->buf.rew4 + (k * E + e0) * n, sizeof(float) * cnt * n, cudaMemcpyDeviceToHost, h->s_out));
    if (h_covered)
      CUDA_TRY(cudaMemcpyAsync(h_covered + e0, h->buf.covered + e0, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, h->s_out));
  }
  CUDA_TRY(cudaStreamSynchronize(h->s_out));
  CUDA_TRY(cudaStreamSynchronize(h->s_comp));
  return 0;
}

extern "C" int uavsim_episode_stats(uavsim_t *h, double out[8], void *stream) {
  if (!h || !out) { SET_ERR("uavsim_episode_stats: NULL argument"); return UAVSIM_ERR_ARG; }
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  uavsim_stats_reduce_kernel<<<1, 32, 0, st>>>(h->d_stats, h->stat_slots, h->d_stats8);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  CUDA_TRY(cudaMemcpyAsync(h->h_stats8, h->d_stats8, sizeof(double) * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int k = 0; k < 8; k++) out[k] = h->h_stats8[k];
  return 0;
}

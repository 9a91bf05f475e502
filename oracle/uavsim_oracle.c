/*
 * uavsim_oracle.c -- CPU restatement of the reference environment hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library, and only as
 * the checker or the timed CPU baseline.  The product path (libuavsim.so, CUDA) never
 * links, imports or calls anything in oracle/.
 *
 * Parity pin: the reference ships no tests or golden vectors, so this
 * restatement is pinned against outputs of the reference itself, executed in the build
 * container by tests/golden/make_golden.py and committed as the .npz files under tests/golden/
 * (tests/test_oracle_golden.py).
 *
 * Scalar fp64, one environment at a time, same evaluation order as the reference's
 * Python (left-to-right, no FMA contraction: compile with -ffp-contract=off), same libm
 * entry points CPython's math module calls (cos, sin, sqrt, exp, fmod, pow).
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root).
 */
#include <alloca.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "uavsim_oracle.h"

#define PI 3.141592653589793 /* math.pi */
#define E_ 2.718281828459045 /* math.e  */

/* Python `x ** 2` on floats is libm pow(x, 2.0) (CPython floatobject.c float_pow). */
static inline double sq(double v) { return pow(v, 2.0); }

/* src/agent/uav.py:53-71  UAV.__distance / UAV.distance */
static inline double dist(double x1, double y1, double x2, double y2) {
  return sqrt(sq(x1 - x2) + sq(y1 - y2));
}

/* Python float `%`: fmod, then shift into the divisor's sign (CPython float_rem). */
static inline double pymod(double a, double b) {
  double r = fmod(a, b);
  if (r != 0.0) {
    if ((b < 0.0) != (r < 0.0)) r += b;
  } else {
    r = copysign(0.0, b);
  }
  return r;
}

/* src/utils/data_util.py:43-56  clip_and_normalize */
static inline double clip_norm(double v, double lo, double hi, int choice) {
  if (v < lo) v = lo;
  if (v > hi) v = hi;
  double mid = (lo + hi) / 2;
  if (choice == -1) return (v - lo) / (hi - lo) - 1;
  if (choice == 0) return (v - lo) / (hi - lo);
  return (v - mid) / (mid - lo);
}

/* numpy pairwise float64 sum of a contiguous 1-D array (np.sum in src/agent/uav.py:288);
 * numpy/_core/src/umath/loops_utils.h.src pairwise_sum, block size 128. */
static double np_pairwise_sum(const double *a, int n) {
  if (n < 8) {
    double r = 0.0;  /* numpy starts from -0.0 only for the identity; result identical for our inputs */
    for (int i = 0; i < n; i++) r += a[i];
    return r;
  } else if (n <= 128) {
    double r[8];
    int i;
    for (int k = 0; k < 8; k++) r[k] = a[k];
    for (i = 8; i < n - (n % 8); i += 8)
      for (int k = 0; k < 8; k++) r[k] += a[i + k];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
  } else {
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
  }
}

/* src/models/PMINet.py:41-62  PMINetwork.forward, eval mode (BatchNorm1d uses running stats,
 * eps = 1e-5), batch of one row, fp32 like torch CPU. */
static float pmi_forward(const OraclePmi *p, const float *x) {
  const int H = p->hidden;
  float *cat = (float *)alloca(sizeof(float) * 3 * H);
  const int off[3] = {0, 5, 9}, dim[3] = {5, 4, 3};
  for (int b = 0; b < 3; b++) {
    const float *W = p->w_in[b], *bias = p->b_in[b];
    const float *g = p->bn_in[b][0], *be = p->bn_in[b][1], *mu = p->bn_in[b][2], *var = p->bn_in[b][3];
    for (int h = 0; h < H; h++) {
      float acc = 0.f;
      for (int k = 0; k < dim[b]; k++) acc += W[h * dim[b] + k] * x[off[b] + k];
      acc += bias[h];
      float y = (acc - mu[h]) / sqrtf(var[h] + 1e-5f) * g[h] + be[h];
      cat[b * H + h] = y > 0.f ? y : 0.f;
    }
  }
  float out = 0.f;
  for (int h = 0; h < H; h++) {
    float acc = 0.f;
    const float *w = p->w1 + (size_t)h * 3 * H;
    for (int k = 0; k < 3 * H; k++) acc += w[k] * cat[k];
    acc += p->b1[h];
    float y = (acc - p->bn1[2][h]) / sqrtf(p->bn1[3][h] + 1e-5f) * p->bn1[0][h] + p->bn1[1][h];
    y = y > 0.f ? y : 0.f;
    out += p->w2[h] * y;
  }
  return out + p->b2[0];
}

/* src/agent/uav.py:156-197 (self part) before any step: lists are empty -> -1 blocks */
void oracle_initial_obs(const OracleParams *P, const double *ux, const double *uy, const int32_t *ua, double *obs) {
  for (int i = 0; i < P->n_uav; i++) {
    double *o = obs + 12 * i;
    for (int k = 0; k < 9; k++) o[k] = -1.0;
    o[9] = ux[i] / P->dc;
    o[10] = uy[i] / P->dc;
    o[11] = (double)ua[i] / P->na;
  }
}

/*
 * One Environment.step for one environment (src/environment.py:120-164).
 * All state arrays are updated in place.  Optional outputs may be NULL.
 */
void oracle_step(const OracleParams *P, int mode, double coop, const OraclePmi *pmi,
                 double *ux, double *uy, double *uh, int32_t *ua,
                 double *tx, double *ty, double *th, const int32_t *actions,
                 double *obs, double *rewards, double *tt_n, double *bp_n, double *dup_n, double *raw_out,
                 int32_t *covered, int32_t *tracker_cnt,
                 uint8_t *obs_mask, uint8_t *comm_mask, uint8_t *nbr_mask, uint8_t *dup_mask, uint8_t *cover_mask) {
  const int n = P->n_uav, m = P->m_targets, Na = P->na;
  const double dp = P->dp, dc = P->dc;

  /* src/agent/target.py:27-60 TARGET.update_position (the random draw at :34 is unused) */
  for (int t = 0; t < m; t++) {
    double dx = P->dt * P->tgt_v_max * cos(th[t]);
    double dy = P->dt * P->tgt_v_max * sin(th[t]);
    tx[t] += dx;
    ty[t] += dy;
    if (0 > ty[t] || ty[t] > P->y_max) {
      th[t] = -th[t];
    } else if (tx[t] < 0 || tx[t] > P->x_max) {
      if (th[t] > 0) th[t] = PI - th[t];
      else th[t] = -PI - th[t];
    }
  }

  /* per-UAV observation lists, built in the sequential (Gauss-Seidel) order of
   * src/environment.py:133-138 */
  double *comm = (double *)malloc(sizeof(double) * 5 * (size_t)n * n);
  double *tobs = (double *)malloc(sizeof(double) * 4 * (size_t)n * m);
  int *ncomm = (int *)calloc(n, sizeof(int)), *ntobs = (int *)calloc(n, sizeof(int));
  double *raw = (double *)malloc(sizeof(double) * n);

  for (int i = 0; i < n; i++) {
    /* src/agent/uav.py:73-99 discrete_action + update_position */
    ua[i] = actions[i];
    int na1 = actions[i] + 1;
    double a = (double)(2 * na1 - Na - 1) * P->uav_h_max / (double)(Na - 1);
    double dx = P->dt * P->uav_v_max * cos(uh[i]);
    double dy = P->dt * P->uav_v_max * sin(uh[i]);
    ux[i] += dx;
    uy[i] += dy;
    uh[i] += P->dt * a;
    uh[i] = pymod(uh[i] + PI, 2 * PI) - PI;

    /* src/agent/uav.py:101-122 observe_target (relative=True) */
    for (int t = 0; t < m; t++) {
      double d = dist(ux[i], uy[i], tx[t], ty[t]);
      int hit = d <= dp;
      if (obs_mask) obs_mask[i * m + t] = (uint8_t)hit;
      if (hit) {
        double *o = tobs + 4 * ((size_t)i * m + ntobs[i]++);
        o[0] = (tx[t] - ux[i]) / dp;
        o[1] = (ty[t] - uy[i]) / dp;
        o[2] = cos(th[t]) * P->tgt_v_max / P->uav_v_max - cos(uh[i]);
        o[3] = sin(th[t]) * P->tgt_v_max / P->uav_v_max - sin(uh[i]);
      }
    }
    /* src/agent/uav.py:124-147 observe_uav: j<i already moved, j>i still old (position, heading, action) */
    for (int j = 0; j < n; j++) {
      double d = dist(ux[i], uy[i], ux[j], uy[j]);
      int hit = (d <= dc) && (j != i);
      if (comm_mask) comm_mask[i * n + j] = (uint8_t)hit;
      if (hit) {
        double *o = comm + 5 * ((size_t)i * n + ncomm[i]++);
        o[0] = (ux[j] - ux[i]) / dc;
        o[1] = (uy[j] - uy[i]) / dc;
        o[2] = cos(uh[j]) - cos(uh[i]);
        o[3] = sin(uh[j]) - sin(uh[i]);
        o[4] = (double)(ua[j] - ua[i]) / Na;
      }
    }
  }

  /* src/environment.py:195-220 calculate_rewards, first loop */
  for (int i = 0; i < n; i++) {
    /* src/agent/uav.py:199-212 tracking reward over the TARGET list (call at :257) */
    double tt = 0;
    for (int t = 0; t < m; t++) {
      double d = dist(ux[i], uy[i], tx[t], ty[t]);
      if (d <= dp) tt += 1 + (dp - d) / dp;
    }
    /* src/agent/uav.py:231-250 boundary punishment */
    double x0 = ux[i] - 0, x1 = P->x_max - ux[i], y0 = uy[i] - 0, y1 = P->y_max - uy[i];
    double dbdr = x0;
    if (x1 < dbdr) dbdr = x1;
    if (y0 < dbdr) dbdr = y0;
    if (y1 < dbdr) dbdr = y1;
    double bp;
    if (0 <= ux[i] && ux[i] <= P->x_max && 0 <= uy[i] && uy[i] <= P->y_max) {
      bp = (dbdr < dp) ? -0.5 * (dp - dbdr) / dp : 0.0;
    } else {
      bp = -1.0 / 2;
    }
    /* src/agent/uav.py:214-229 duplicate tracking punishment, radio = 2 */
    double dup = 0;
    for (int j = 0; j < n; j++) {
      if (j == i) continue;
      double d = dist(ux[i], uy[i], ux[j], uy[j]);
      int hit = d <= 2 * dp;
      if (dup_mask) dup_mask[i * n + j] = (uint8_t)hit;
      if (hit) dup += -0.5 * exp((2 * dp - d) / (2 * dp));
    }
    if (dup_mask) dup_mask[i * n + i] = 0;
    double ttn = clip_norm(tt, 0, 2 * m, 0);
    double dupn = clip_norm(dup, -E_ / 2 * n, 0, -1);
    double bpn = clip_norm(bp, -1.0 / 2, 0, -1);
    tt_n[i] = ttn;
    bp_n[i] = bpn;
    dup_n[i] = dupn;
    raw[i] = P->alpha * ttn + P->beta * bpn + P->gamma * dupn;
    if (raw_out) raw_out[i] = raw[i];
  }

  /* src/environment.py:109-118 get_states / src/agent/uav.py:156-190 weighted mean.
   * Computed here because the PMI branch needs every UAV's post-step local state. */
  for (int i = 0; i < n; i++) {
    double *o = obs + 12 * i;
    if (ncomm[i]) {
      double s[5] = {0, 0, 0, 0, 0};
      for (int k = 0; k < ncomm[i]; k++) {
        const double *c = comm + 5 * ((size_t)i * n + k);
        double w = dist(c[0], c[1], ux[i], uy[i]);
        if (!(w < 1)) w = 1;
        for (int q = 0; q < 5; q++) s[q] += c[q] / w;
      }
      for (int q = 0; q < 5; q++) o[q] = s[q] / ncomm[i];
    } else {
      for (int q = 0; q < 5; q++) o[q] = -1.0;
    }
    if (ntobs[i]) {
      double s[4] = {0, 0, 0, 0};
      for (int k = 0; k < ntobs[i]; k++) {
        const double *c = tobs + 4 * ((size_t)i * m + k);
        double w = dist(c[0], c[1], ux[i], uy[i]);
        if (!(w < 1)) w = 1;
        for (int q = 0; q < 4; q++) s[q] += c[q] / w;
      }
      for (int q = 0; q < 4; q++) o[5 + q] = s[q] / ntobs[i];
    } else {
      for (int q = 0; q < 4; q++) o[5 + q] = -1.0;
    }
    o[9] = ux[i] / dc;
    o[10] = uy[i] / dc;
    o[11] = (double)ua[i] / Na;
  }

  /* src/environment.py:222-227 second loop: cooperative reward, then clip to [-1, 1] */
  double *nbr_r = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  float *dep = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    int cnt = 0;
    for (int j = 0; j < n; j++) {
      int hit = (j != i) && dist(ux[i], uy[i], ux[j], uy[j]) <= dp;
      if (nbr_mask) nbr_mask[i * n + j] = (uint8_t)hit;
    }
    double r;
    if (coop == 0) {
      r = raw[i]; /* src/agent/uav.py:271-272 / :300-301 */
    } else if (mode == ORACLE_MODE_PMI) {
      /* src/agent/uav.py:262-291 */
      for (int j = 0; j < n; j++) {
        if (j == i || !(dist(ux[i], uy[i], ux[j], uy[j]) <= dp)) continue;
        float in[12];
        for (int q = 0; q < 12; q++) in[q] = (float)(obs[12 * i + q] * obs[12 * j + q]);
        dep[cnt] = pmi_forward(pmi, in);
        nbr_r[cnt] = raw[j];
        cnt++;
      }
      if (cnt) {
        /* scipy.special.softmax on float32 */
        float mx = dep[0];
        for (int k = 1; k < cnt; k++) if (dep[k] > mx) mx = dep[k];
        float ssum = 0.f;
        for (int k = 0; k < cnt; k++) { dep[k] = expf(dep[k] - mx); }
        /* np.sum over float32 (pairwise; sequential below 8 elements) */
        if (cnt < 8) { for (int k = 0; k < cnt; k++) ssum += dep[k]; }
        else {
          float q[8]; int k;
          for (int u = 0; u < 8; u++) q[u] = dep[u];
          for (k = 8; k < cnt - (cnt % 8); k += 8) for (int u = 0; u < 8; u++) q[u] += dep[k + u];
          ssum = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
          for (; k < cnt; k++) ssum += dep[k];
        }
        for (int k = 0; k < cnt; k++) nbr_r[k] = nbr_r[k] * (double)(dep[k] / ssum);
        r = (1 - coop) * raw[i] + coop * np_pairwise_sum(nbr_r, cnt);
      } else {
        r = (1 - coop) * raw[i];
      }
    } else {
      /* src/agent/uav.py:293-310; note the conditional binds the whole expression: no neighbour -> 0 */
      double s = 0;
      for (int j = 0; j < n; j++) {
        if (j == i || !(dist(ux[i], uy[i], ux[j], uy[j]) <= dp)) continue;
        s += raw[j];
        cnt++;
      }
      r = cnt ? (1 - coop) * raw[i] + coop * s / cnt : 0.0;
    }
    rewards[i] = clip_norm(r, -1, 1, 1);
  }

  /* src/environment.py:246-253 calculate_covered_target (strict <) + per-target tracker counts */
  int cov = 0;
  for (int t = 0; t < m; t++) {
    int c = 0;
    for (int i = 0; i < n; i++) {
      int hit = dist(ux[i], uy[i], tx[t], ty[t]) < dp;
      if (cover_mask) cover_mask[i * m + t] = (uint8_t)hit;
      c += hit;
    }
    if (tracker_cnt) tracker_cnt[t] = c;
    cov += (c > 0);
  }
  *covered = cov;

  free(comm); free(tobs); free(ncomm); free(ntobs); free(raw); free(nbr_r); free(dep);
}

/* ------------------------------------------------------------------------------------------
 * Batched driver used for cross-checking the CUDA path and as the timed CPU baseline:
 * E independent environments, env-major structure-of-arrays ([E,n] / [E,m]), split across
 * pthreads.  Outputs are float64; any of obs/rew4/covered/tracker may be NULL.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const OracleParams *P; int mode; double coop; const OraclePmi *pmi;
  int64_t e0, e1;
  double *ux, *uy, *uh; int32_t *ua; double *tx, *ty, *th; const int32_t *actions;
  double *obs, *rew4; int32_t *covered, *tracker; int64_t E;
} BatchJob;

static void *batch_worker(void *arg) {
  BatchJob *J = (BatchJob *)arg;
  const int n = J->P->n_uav, m = J->P->m_targets;
  double *obs = (double *)malloc(sizeof(double) * 12 * n);
  double *r4 = (double *)malloc(sizeof(double) * 4 * n);
  for (int64_t e = J->e0; e < J->e1; e++) {
    int32_t cov;
    oracle_step(J->P, J->mode, J->coop, J->pmi, J->ux + e * n, J->uy + e * n, J->uh + e * n, J->ua + e * n,
                J->tx + e * m, J->ty + e * m, J->th + e * m, J->actions + e * n,
                J->obs ? J->obs + e * n * 12 : obs, r4, r4 + n, r4 + 2 * n, r4 + 3 * n, NULL,
                &cov, J->tracker ? J->tracker + e * m : NULL, NULL, NULL, NULL, NULL, NULL);
    if (J->rew4)
      for (int k = 0; k < 4; k++) memcpy(J->rew4 + ((size_t)k * J->E + e) * n, r4 + k * n, sizeof(double) * n);
    if (J->covered) J->covered[e] = cov;
  }
  free(obs); free(r4);
  return NULL;
}

void oracle_step_batch(const OracleParams *P, int mode, double coop, const OraclePmi *pmi, int64_t E,
                       double *ux, double *uy, double *uh, int32_t *ua, double *tx, double *ty, double *th,
                       const int32_t *actions, double *obs, double *rew4, int32_t *covered, int32_t *tracker,
                       int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > E) nthreads = (int)(E > 0 ? E : 1);
  pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  BatchJob *jobs = (BatchJob *)malloc(sizeof(BatchJob) * nthreads);
  for (int k = 0; k < nthreads; k++) {
    BatchJob J = {P, mode, coop, pmi, E * k / nthreads, E * (k + 1) / nthreads,
                  ux, uy, uh, ua, tx, ty, th, actions, obs, rew4, covered, tracker, E};
    jobs[k] = J;
    if (nthreads == 1) batch_worker(&jobs[k]);
    else pthread_create(&tid[k], NULL, batch_worker, &jobs[k]);
  }
  if (nthreads > 1)
    for (int k = 0; k < nthreads; k++) pthread_join(tid[k], NULL);
  free(tid); free(jobs);
}

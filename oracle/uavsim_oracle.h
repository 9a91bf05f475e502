/*
 * uavsim_oracle.h -- C interface of the CPU oracle (test infrastructure, see uavsim_oracle.c).
 */
#ifndef UAVSIM_ORACLE_H
#define UAVSIM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_MODE_SELF = 0, ORACLE_MODE_MEAN = 1, ORACLE_MODE_PMI = 2 };

/* scenario constants, already converted the way src/environment.py:97-107 does
 * (h_max fields are pi / yaml value) */
typedef struct OracleParams {
  int32_t n_uav, m_targets, na, _pad;
  double x_max, y_max, dt, uav_v_max, uav_h_max, dc, dp, tgt_v_max, tgt_h_max, alpha, beta, gamma;
} OracleParams;

/* raw (unfolded) PMINetwork parameters, torch layouts ([out,in] row-major), src/models/PMINet.py:29-38.
 * bn arrays: [0]=weight (gamma) [1]=bias (beta) [2]=running_mean [3]=running_var */
typedef struct OraclePmi {
  int32_t hidden, _pad;
  const float *w_in[3], *b_in[3];   /* fc_comm [H,5], fc_obs [H,4], fc_boundary_state [H,3] */
  const float *bn_in[3][4];
  const float *w1, *b1;             /* fc1 [H,3H] */
  const float *bn1[4];
  const float *w2, *b2;             /* fc2 [1,H] */
} OraclePmi;

void oracle_initial_obs(const OracleParams *P, const double *ux, const double *uy, const int32_t *ua, double *obs);

void oracle_step(const OracleParams *P, int mode, double coop, const OraclePmi *pmi,
                 double *ux, double *uy, double *uh, int32_t *ua,
                 double *tx, double *ty, double *th, const int32_t *actions,
                 double *obs, double *rewards, double *tt_n, double *bp_n, double *dup_n, double *raw_out,
                 int32_t *covered, int32_t *tracker_cnt,
                 uint8_t *obs_mask, uint8_t *comm_mask, uint8_t *nbr_mask, uint8_t *dup_mask, uint8_t *cover_mask);

/* E environments, env-major SoA; rew4 is [4][E][n] (rewards, tracking, boundary, duplicate) */
void oracle_step_batch(const OracleParams *P, int mode, double coop, const OraclePmi *pmi, int64_t E,
                       double *ux, double *uy, double *uh, int32_t *ua, double *tx, double *ty, double *th,
                       const int32_t *actions, double *obs, double *rew4, int32_t *covered, int32_t *tracker,
                       int nthreads);

#ifdef __cplusplus
}
#endif
#endif

/* TEST INFRASTRUCTURE ONLY -- see quadsim_oracle.h.  fp64 CPU restatement of the reference hot path.
 *
 * Compile with -ffp-contract=off: the reference is numpy float64 without fused multiply-add, and the
 * expressions below keep the reference's evaluation order so the port stays within a few ulps of it.
 * Parity is PINNED: tests/test_oracle_golden.py checks every function here against tests/golden npz files,
 * which tests/golden/make_golden.py produced by running the unmodified reference.
 */
#include "quadsim_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- constants: dynamics/quadrotor.py:10-63, controller/PIDController.py:10-44 and the envs' __init__ literals, as a
 * parameter block (qo_params, quadsim_oracle.h) so that the parity tests can also run the model with non-default mass /
 * inertia / gains / thresholds (the CUDA library's qs_env_params_t).  Every function takes the block explicitly; the public
 * leaf functions use the reference's literals. ---- */
#define QO_PI 3.141592653589793
static const qo_params QO_DEF = {
    0.02, 9.81, 0.18, 0.086, {0.00025, 0.000232, 0.0003738}, 1.5e-9 / 6.11e-8, {85, 85, 175},
    {-10.0, -10.0, -9.5}, {5.1, 5.1, 4.0}, {-1.0, -1.0, 50.0}, {-1.65, -1.65, 8.0}, {-0.7, -0.7, 1.0},
    {8, -50, 5}, {10, -50, 5}, {10, -50, 5}, {0, 0, 0}, 0.1, 0.1, 0.1, 10, 0, 0.1, 600, 0};
void qo_default_params(int kind, qo_params* out) {
  *out = QO_DEF;
  if (kind == QO_HOVER) { out->chaser_ini_pos[0] = 0; out->chaser_ini_pos[1] = 0; out->chaser_ini_pos[2] = 5; out->range_limit = 100; }
  else if (kind == QO_DOCK_V0) out->range_limit = 3;
  else { out->range_limit = 10; out->target_des_vel[0] = 0.2; }
}

/* utils/transform.py:143-144  deg2rad(angle) = angle * np.pi / 180 */
static double deg2rad(double a) { return a * QO_PI / 180; }
static double sgn(double x) { return (x > 0) - (x < 0); }

/* utils/transform.py:94-120 */
void qo_quat2euler(const double q[4], double e[3]) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  double r10 = 2.0 * (x * y - w * z);
  double r11 = w * w - x * x + y * y - z * z;
  double r12 = 2.0 * (w * x + y * z);
  double r02 = 2.0 * (x * z - w * y);
  double r22 = w * w - x * x - y * y + z * z;
  double phi, theta, psi;
  if (r12 < 1.0) {
    if (r12 < -1.0) {
      phi = -QO_PI / 2.0; psi = atan2(-r10, r11); theta = 0;
    } else {
      phi = asin(r12); psi = atan2(-r10, r11); theta = atan2(-r02, r22);
    }
  } else {
    phi = QO_PI / 2.0; psi = atan2(-r10, r11); theta = 0;
  }
  e[0] = phi; e[1] = theta; e[2] = psi;
}

/* utils/transform.py:123-136 */
void qo_euler2quat(const double e[3], double q[4]) {
  double cy = cos(e[2] * 0.5), sy = sin(e[2] * 0.5);
  double cp = cos(e[1] * 0.5), sp = sin(e[1] * 0.5);
  double cr = cos(e[0] * 0.5), sr = sin(e[0] * 0.5);
  q[0] = cr * cp * cy - sr * sp * sy;
  q[1] = sr * cp * cy - cr * sp * sy;
  q[2] = sr * cp * sy + cr * sp * cy;
  q[3] = cr * cp * sy + sr * sp * cy;
}

/* utils/transform.py:4-20 == dynamics/quadrotor.py:226-245 (elementwise square, raw quat[0]) */
void qo_quat2rot(const double q[4], double R[9]) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double n1 = q[1] / n, n2 = q[2] / n, n3 = q[3] / n;
  double h[9] = {0, -n3, n2, n3, 0, -n1, -n2, n1, 0};
  double w2 = 2 * q[0];
  for (int i = 0; i < 9; i++) R[i] = ((i % 4 == 0 ? 1.0 : 0.0) + (2 * h[i]) * h[i]) + w2 * h[i];
}

/* utils/transform.py:23-46 */
void qo_rot2euler(const double R[9], double e[3]) {
  double phi, theta, psi;
  if (R[5] < 1.0) {
    phi = asin(R[5]); psi = atan2(-R[3], R[4]); theta = atan2(-R[2], R[8]);
    if (R[5] < -1.0) { phi = -QO_PI / 2.0; psi = atan2(-R[3], R[4]); theta = 0; }
  } else {
    phi = QO_PI / 2.0; psi = atan2(-R[3], R[4]); theta = 0;
  }
  e[0] = phi; e[1] = theta; e[2] = psi;
}

/* dynamics/quadrotor.py:171-187 (A = :47-50, B = :52-54; F_max/4 == mass*gravity exactly) */
static void u_limit_p(const qo_params* P, const double u[4], double out[4]) {
  const double QO_L = P->arm_length;
  const double hl = 0.5 / QO_L, fmax = 4 * P->mass * P->gravity / 4, fmin = 0.0;
  double p[4];
  p[0] = 0.25 * u[0] + 0 * u[1] + (-hl) * u[2];
  p[1] = 0.25 * u[0] + hl * u[1] + 0 * u[2];
  p[2] = 0.25 * u[0] + 0 * u[1] + hl * u[2];
  p[3] = 0.25 * u[0] + (-hl) * u[1] + 0 * u[2];
  for (int i = 0; i < 4; i++) {
    if (p[i] > fmax) p[i] = fmax;
    if (p[i] < fmin) p[i] = fmin;
  }
  out[0] = p[0] + p[1] + p[2] + p[3];
  out[1] = 0 * p[0] + QO_L * p[1] + 0 * p[2] + (-QO_L) * p[3];
  out[2] = (-QO_L) * p[0] + 0 * p[1] + QO_L * p[2] + 0 * p[3];
  out[3] = u[3];
}
void qo_u_limit(const double u[4], double out[4]) { u_limit_p(&QO_DEF, u, out); }

/* dynamics/quadrotor.py:80-113 */
static void df_p(const qo_params* P, const double s[13], const double u[4], double ds[13]) {
  const double QO_MASS = P->mass, QO_G = P->gravity;
  const double* QO_I = P->inertia;
  double F = u[0];
  double M[3] = {QO_I[0] * u[1], QO_I[1] * u[2], QO_I[2] * u[3]};
  const double* q = s + 6;
  const double* w = s + 10;
  double R[9];
  qo_quat2rot(q, R);
  /* R_b2w @ [0,0,F] = column 2 of R^T = row 2 of R */
  double g3[3] = {0, 0, QO_MASS * QO_G};
  for (int i = 0; i < 3; i++) ds[3 + i] = 1.0 / QO_MASS * (R[6 + i] * F - g3[i]);
  double e = 1.0 - (q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double sk[4];
  sk[0] = -w[0] * q[1] + -w[1] * q[2] + -w[2] * q[3];
  sk[1] = w[0] * q[0] + -w[1] * q[2] + w[2] * q[3];
  sk[2] = w[1] * q[0] + w[2] * q[1] + -w[0] * q[3];
  sk[3] = w[2] * q[0] + -w[1] * q[1] + w[0] * q[2];
  for (int i = 0; i < 4; i++) ds[6 + i] = -0.5 * sk[i] + 2.0 * e * q[i];
  double Iw[3] = {QO_I[0] * w[0], QO_I[1] * w[1], QO_I[2] * w[2]};
  double cr[3] = {w[1] * Iw[2] - w[2] * Iw[1], w[2] * Iw[0] - w[0] * Iw[2], w[0] * Iw[1] - w[1] * Iw[0]};
  for (int i = 0; i < 3; i++) ds[10 + i] = (1.0 / QO_I[i]) * (M[i] - cr[i]);
  ds[0] = s[3]; ds[1] = s[4]; ds[2] = s[5];
}
void qo_df(const double s[13], const double u[4], double ds[13]) { df_p(&QO_DEF, s, u, ds); }

/* dynamics/quadrotor.py:126-144 with :115-124 -- SURVEY Q1/Q2: SciPy RK45 integrates a constant RHS
 * df(state_at_step_start, u_prev), i.e. one explicit Euler step of length dt; the new command is stored
 * only AFTER the integration (one-step latency).  attitude_limit: quadrotor.py:146-168. */
static void drone_step_p(const qo_params* P, double s[13], double u_prev[4], const double u_cmd[4]) {
  double ds[13];
  df_p(P, s, u_prev, ds);
  for (int i = 0; i < 13; i++) s[i] = s[i] + P->dt * ds[i];
  double e[3], lim[4], tmp[3];
  int flag = 0;
  qo_quat2euler(s + 6, e);
  double r = e[0], p = e[1], y = e[2];
  const double dr = deg2rad(P->att_limit_deg[0]), dp = deg2rad(P->att_limit_deg[1]), dy = deg2rad(P->att_limit_deg[2]);
  if (fabs(r) >= dr) { tmp[0] = sgn(r) * dr; tmp[1] = p; tmp[2] = y; qo_euler2quat(tmp, lim); flag = 1; }
  if (fabs(p) >= dp) { tmp[0] = r; tmp[1] = sgn(p) * dp; tmp[2] = y; qo_euler2quat(tmp, lim); flag = 1; }
  if (fabs(y) >= dy) { tmp[0] = r; tmp[1] = p; tmp[2] = sgn(y) * dy; qo_euler2quat(tmp, lim); flag = 1; }
  if (fabs(r) <= dr && fabs(p) <= dp && fabs(y) <= dy) flag = 0;
  if (flag) {
    for (int i = 0; i < 4; i++) s[6 + i] = lim[i];
    s[10] = s[11] = s[12] = 0.0;
  }
  u_limit_p(P, u_cmd, u_prev);
}
void qo_drone_step(double s[13], double u_prev[4], const double u_cmd[4]) { drone_step_p(&QO_DEF, s, u_prev, u_cmd); }

/* dynamics/quadrotor.py:213-224 with dock_port_inB_pos = [d,0,0]; only 'pos' and 'vel' are ever read */
void qo_dock_port(const double s[13], double d, double pos[3], double vel[3]) {
  double R[9];
  qo_quat2rot(s + 6, R);
  double r[3] = {R[0] * d, R[1] * d, R[2] * d}; /* R_b2w @ [d,0,0] = row 0 of R, times d */
  for (int i = 0; i < 3; i++) pos[i] = s[i] + r[i];
  const double* w = s + 10;
  vel[0] = s[3] + (0 * r[0] + -w[2] * r[1] + w[1] * r[2]);
  vel[1] = s[4] + (w[2] * r[0] + 0 * r[1] + -w[0] * r[2]);
  vel[2] = s[5] + (-w[1] * r[0] + w[0] * r[1] + 0 * r[2]);
}

/* controller/PIDController.py:10-44 gains: qo_params.kp_att / kd_att / kp_pos / kd_pos / kp_vel; the rest are fixed */
#define FF_YAW 0.0
#define KD_VX 0.0
#define KD_VY 0.0
#define KD_VZ 0.1

/* controller/PIDController.py:52-74 */
static void attitude_controller(const qo_params* P, const double des[13], const double now[13], double M[3]) {
  double ad[3], an[3];
  qo_quat2euler(des + 6, ad);
  qo_quat2euler(now + 6, an);
  double ea[3], er[3];
  for (int i = 0; i < 3; i++) { ea[i] = ad[i] - an[i]; er[i] = des[10 + i] - now[10 + i]; }
  M[0] = P->kp_att[0] * ea[0] + P->kd_att[0] * er[0];
  M[1] = P->kp_att[1] * ea[1] + P->kd_att[1] * er[1];
  M[2] = FF_YAW + P->kp_att[2] * ea[2] + P->kd_att[2] * er[2];
}

/* shared tail of hover_controller (PIDController.py:87-104) and vel_controller (:119-134):
 * acc_des -> F, desired attitude; OVERWRITES des[6:10], des[10], des[11] */
static double acc2att(const qo_params* P, const double acc[3], double des[13]) {
  const double QO_MASS = P->mass, QO_G = P->gravity;
  double F = QO_MASS * QO_G + QO_MASS * acc[2];
  double att[3];
  qo_quat2euler(des + 6, att);
  double psi = att[2];
  double phi = (acc[0] * sin(psi) - acc[1] * cos(psi)) / QO_G;
  double theta = (acc[0] * cos(psi) + acc[1] * sin(psi)) / QO_G;
  att[0] = phi; att[1] = theta; att[2] = psi;
  qo_euler2quat(att, des + 6);
  des[10] = 0; des[11] = 0;
  return F;
}

/* controller/PIDController.py:179-185 (hover_controller :76-104 then attitude_controller) */
static void pid_p(const qo_params* P, double des[13], const double now[13], double u[4]) {
  double acc[3];
  acc[0] = P->kp_pos[0] * (des[0] - now[0]) + P->kd_pos[0] * (des[3] - now[3]);
  acc[1] = P->kp_pos[1] * (des[1] - now[1]) + P->kd_pos[1] * (des[4] - now[4]);
  acc[2] = P->kp_pos[2] * (des[2] - now[2]) + P->kd_pos[2] * (des[5] - now[5]);
  u[0] = acc2att(P, acc, des);
  attitude_controller(P, des, now, u + 1);
}
void qo_pid(double des[13], const double now[13], double u[4]) { pid_p(&QO_DEF, des, now, u); }

/* controller/PIDController.py:106-141 */
static void vel_controller_p(const qo_params* P, double des[13], const double now[13], const double last[13], double u[4]) {
  double acc[3];
  acc[0] = P->kp_vel[0] * (des[3] - now[3]) + KD_VX * (now[3] - last[3]);
  acc[1] = P->kp_vel[1] * (des[4] - now[4]) + KD_VY * (now[4] - last[4]);
  acc[2] = P->kp_vel[2] * (des[5] - now[5]) + KD_VZ * (now[5] - last[5]);
  u[0] = acc2att(P, acc, des);
  attitude_controller(P, des, now, u + 1);
}
void qo_vel_controller(double des[13], const double now[13], const double last[13], double u[4]) { vel_controller_p(&QO_DEF, des, now, last, u); }

/* The PID expert of the imitation-learning scripts: run_expert_policy.py:44-69 == run_expert_record.py:40-139.
 *   des_vel = kp * (env.state_target[0:3] + [-0.2,0,0] - env.state_chaser[0:3]) + kd * (-env.state_chaser[3:6])   (:133-134)
 *   if idx_per_epi != 0: state_des[3:6] = des_vel                                                               (:135-136)
 *   u = control.vel_controller(state_des, env.state_chaser, state_last)                                         (:137)
 *   action = (np.linalg.inv(env.chaser.rotor2control) @ u - env.action_mean) / env.action_std                   (:138)
 * mem[8] = state_des[3:6], pad, state_des[6:10].  alias != 0: state_des IS env.chaser_ini_state (:41), which env.reset()
 * hands to Drone.reset as the chaser's state array (quadrotor.py:70,75): on the first step of an episode the chaser is
 * seated on the mutated ini state and des / now / last are ONE array -- reproduced here by really aliasing them.
 * Returns 1 when the chaser state was re-seated (the caller recomputes the reset observation BEFORE the controller ran:
 * use qo_pid_expert_reseat first). */
int qo_pid_expert_reseat(double chaser[13], const double mem[8], int first, int alias) {
  if (!(first && alias)) return 0;
  chaser[3] = mem[0]; chaser[4] = mem[1]; chaser[5] = mem[2];
  memcpy(chaser + 6, mem + 4, 4 * sizeof(double));
  return 1;
}
void qo_pid_expert_action(double chaser[13], const double target[13], double mem[8], int first, int alias, double act[4]) {
  const double kp = 0.35, kd = 0;
  double u[4];
  if (first && alias) {
    qo_vel_controller(chaser, chaser, chaser, u); /* des == now == last: writes the levelled quaternion into the chaser itself */
    mem[0] = chaser[3]; mem[1] = chaser[4]; mem[2] = chaser[5];
    memcpy(mem + 4, chaser + 6, 4 * sizeof(double));
  } else {
    double des[13] = {0};
    const double off[3] = {-0.2, 0, 0};
    if (!first)
      for (int k = 0; k < 3; k++) mem[k] = kp * (target[k] + off[k] - chaser[k]) + kd * (-chaser[3 + k]);
    memcpy(des + 3, mem, 3 * sizeof(double));
    memcpy(des + 6, mem + 4, 4 * sizeof(double));
    qo_vel_controller(des, chaser, chaser, u); /* kd_v* == 0: `last` is never felt (Q5) */
    memcpy(mem + 4, des + 6, 4 * sizeof(double));
  }
  /* inv(rotor2control), quadrotor.py:56-59, in closed form; action_mean = action_std = m g / 2 (docking_env.py:98-99) */
  const double L = 0.086, lam = 1.5e-9 / 6.11e-8, half = 1.0 * 0.18 * 9.81 / 2.0;
  const double f[4] = {u[0] / 4 - u[2] / (2 * L) + u[3] / (4 * lam), u[0] / 4 + u[1] / (2 * L) - u[3] / (4 * lam),
                       u[0] / 4 + u[2] / (2 * L) + u[3] / (4 * lam), u[0] / 4 - u[1] / (2 * L) - u[3] / (4 * lam)};
  for (int k = 0; k < 4; k++) act[k] = (f[k] - half) / half;
}

static void mat3mul(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void mat3vec(const double A[9], const double x[3], double y[3]) {
  for (int i = 0; i < 3; i++) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}

/* gym_docking/envs/docking_env.py:257-295 (== moving_docking_env.py:222-260), with the env's dock ports
 * chaser [+0.1,0,0] (docking_env.py:38) and target [-0.1,0,0] (:51) */
static void state2rel_p(const qo_params* P, const double sc[13], const double st[13], double rel[12]) {
  double RB[9], RA[9], RAt[9], RAB[9], RABA[9];
  double pc[3], vc[3], pt[3], vt[3];
  qo_dock_port(sc, P->dock_port, pc, vc);
  qo_dock_port(st, -P->dock_port, pt, vt);
  qo_quat2rot(st + 6, RB);
  qo_quat2rot(sc + 6, RA);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) RAt[3 * i + j] = RA[3 * j + i];
  for (int i = 0; i < 3; i++) { rel[i] = pt[i] - pc[i]; rel[3 + i] = vt[i] - vc[i]; }
  mat3mul(RB, RAt, RAB);
  qo_rot2euler(RAB, rel + 6);
  double phi = rel[6], theta = rel[7];
  double a[3], b[3];
  mat3vec(RB, st + 10, a);
  mat3mul(RAB, RA, RABA);
  mat3vec(RABA, sc + 10, b);
  double p = a[0] - b[0], q = a[1] - b[1], r = a[2] - b[2];
  rel[9] = p * cos(theta) + r * sin(theta);
  rel[10] = q - tan(phi) * (r * cos(theta) - p * sin(theta));
  rel[11] = (r * cos(theta) - p * sin(theta)) / cos(phi);
}
void qo_state2rel(const double sc[13], const double st[13], double rel[12]) { state2rel_p(&QO_DEF, sc, st, rel); }

/* Philox4x32-10, Salmon et al. SC'11 (Random123).  Not in the reference: the north star's specified
 * extension for randomised auto-reset (SURVEY section 0.1 closing note). */
void qo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void qo_philox_jitter3(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t stream, float scale, float out[3]) {
  uint32_t ctr[4] = {(uint32_t)env_id, (uint32_t)(env_id >> 32), episode, stream};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t x[4];
  qo_philox4x32_10(ctr, key, x);
  for (int i = 0; i < 3; i++) {
    volatile float u = (float)(x[i] >> 8) * 5.9604644775390625e-08f; /* 2^-24, exact */
    volatile float c = u - 0.5f;
    out[i] = c * scale;
  }
}

/* ------------------------------------------------------------------------------------------------ */
struct qo_vec {
  qo_config cfg;
  qo_params P;
  int obs_dim;
  double *chaser, *target, *uc, *ut, *desq, *last_shaping, *ep_return;
  int64_t *t, *episode;
  double stats[5];
};

/* global id of env i: env_id0 + i * id_stride (id_stride > 1: a strided SAMPLE of a larger batch, for the parity tests at
 * sizes the oracle cannot step in full) */
static uint64_t env_gid(const qo_vec* v, int64_t i) { return v->cfg.env_id0 + (uint64_t)i * (uint64_t)(v->cfg.id_stride > 1 ? v->cfg.id_stride : 1); }

static void dock_ini(const qo_vec* v, int64_t i, double sc[13], double st[13]) {
  /* docking_env.py:34-57 / moving_docking_env.py:33-56 ; jitter = the commented-out U(-0.5,0.5)^3 (:34) */
  memset(sc, 0, 13 * sizeof(double));
  memset(st, 0, 13 * sizeof(double));
  float base[3] = {(float)v->P.chaser_ini_pos[0], (float)v->P.chaser_ini_pos[1], (float)v->P.chaser_ini_pos[2]}, j[3] = {0, 0, 0};
  if (v->cfg.randomize != QO_RAND_NONE) {
    uint32_t ep = v->cfg.randomize == QO_RAND_ONCE ? 0u : (uint32_t)v->episode[i];
    qo_philox_jitter3(v->cfg.seed, env_gid(v, i), ep, 0, (float)(2 * v->cfg.half_width), j);
  }
  for (int k = 0; k < 3; k++) { volatile float p = base[k] + j[k]; sc[k] = (double)p; }
  sc[6] = 1.0;
  for (int k = 0; k < 3; k++) st[k] = (double)(float)v->P.target_ini_pos[k];
  st[6] = 1.0;
}

static void hover_ini(const qo_vec* v, int64_t i, double s[13]) {
  /* hovering_env.py:23-29: pos [0,0,5]+U(-1,1)^3, euler U(-0.2,0.2)^3, sampled once per env */
  memset(s, 0, 13 * sizeof(double));
  float base[3] = {(float)v->P.chaser_ini_pos[0], (float)v->P.chaser_ini_pos[1], (float)v->P.chaser_ini_pos[2]}, jp[3] = {0, 0, 0}, je[3] = {0, 0, 0};
  if (v->cfg.randomize != QO_RAND_NONE) {
    uint32_t ep = v->cfg.randomize == QO_RAND_ONCE ? 0u : (uint32_t)v->episode[i];
    qo_philox_jitter3(v->cfg.seed, env_gid(v, i), ep, 0, (float)(2 * v->cfg.half_width), jp);
    qo_philox_jitter3(v->cfg.seed, env_gid(v, i), ep, 1, 0.4f, je);
  }
  for (int k = 0; k < 3; k++) { volatile float p = base[k] + jp[k]; s[k] = (double)p; }
  double e[3] = {(double)je[0], (double)je[1], (double)je[2]};
  qo_euler2quat(e, s + 6);
}

static void env_seat(qo_vec* v, int64_t i, double* obs) {
  /* Drone.reset (quadrotor.py:65-78): state<-ini, u<-0 ; env reset (docking_env.py:233-244): t, shaping <- 0;
   * target_state_des is NOT touched (Q4) */
  if (v->cfg.kind == QO_HOVER) {
    hover_ini(v, i, v->chaser + 13 * i);
    memset(v->uc + 4 * i, 0, 4 * sizeof(double));
    if (obs) memcpy(obs, v->chaser + 13 * i, 13 * sizeof(double));
  } else {
    dock_ini(v, i, v->chaser + 13 * i, v->target + 13 * i);
    memset(v->uc + 4 * i, 0, 4 * sizeof(double));
    memset(v->ut + 4 * i, 0, 4 * sizeof(double));
    if (obs) state2rel_p(&v->P, v->chaser + 13 * i, v->target + 13 * i, obs);
  }
  v->t[i] = 0;
  v->last_shaping[i] = 0.0;
  v->ep_return[i] = 0.0;
}

qo_vec* qo_vec_create(const qo_config* cfg) {
  qo_vec* v = (qo_vec*)calloc(1, sizeof(qo_vec));
  v->cfg = *cfg;
  if (cfg->params) v->P = *cfg->params; else qo_default_params(cfg->kind, &v->P);
  v->cfg.params = NULL;
  if (v->cfg.half_width <= 0) v->cfg.half_width = cfg->kind == QO_HOVER ? 1.0 : 0.5;
  v->obs_dim = cfg->kind == QO_HOVER ? 13 : 12;
  int64_t n = cfg->n_envs;
  v->chaser = (double*)calloc(13 * n, sizeof(double));
  v->target = (double*)calloc(13 * n, sizeof(double));
  v->uc = (double*)calloc(4 * n, sizeof(double));
  v->ut = (double*)calloc(4 * n, sizeof(double));
  v->desq = (double*)calloc(4 * n, sizeof(double));
  v->last_shaping = (double*)calloc(n, sizeof(double));
  v->ep_return = (double*)calloc(n, sizeof(double));
  v->t = (int64_t*)calloc(n, sizeof(int64_t));
  v->episode = (int64_t*)calloc(n, sizeof(int64_t));
  for (int64_t i = 0; i < n; i++) {
    v->desq[4 * i] = 1.0; /* euler2quat(0,0,0): docking_env.py:60-63 */
    env_seat(v, i, NULL);
  }
  return v;
}

void qo_vec_destroy(qo_vec* v) {
  if (!v) return;
  free(v->chaser); free(v->target); free(v->uc); free(v->ut); free(v->desq);
  free(v->last_shaping); free(v->ep_return); free(v->t); free(v->episode);
  free(v);
}

int qo_vec_obs_dim(const qo_vec* v) { return v->obs_dim; }

void* qo_vec_field(qo_vec* v, int f) {
  switch (f) {
    case QO_F_CHASER: return v->chaser;
    case QO_F_TARGET: return v->target;
    case QO_F_UC: return v->uc;
    case QO_F_UT: return v->ut;
    case QO_F_DESQ: return v->desq;
    case QO_F_LAST_SHAPING: return v->last_shaping;
    case QO_F_EP_RETURN: return v->ep_return;
    case QO_F_T: return v->t;
    case QO_F_EPISODE: return v->episode;
  }
  return NULL;
}

/* reset(): every env that has stepped in its current episode starts its next episode; fresh envs stay */
void qo_vec_reset(qo_vec* v, double* obs_out) {
  for (int64_t i = 0; i < v->cfg.n_envs; i++) {
    if (v->t[i] != 0) v->episode[i] += 1;
    env_seat(v, i, obs_out ? obs_out + (int64_t)v->obs_dim * i : NULL);
  }
}

static double norm3(const double* x) { return sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]); }

typedef struct { double ret, len, neps, ndock, nnonfinite; } acc_t;

static void dock_step_one(qo_vec* v, int64_t i, const double* a, double* obs, double* reward, uint8_t* done,
                          uint8_t* info, double* term_obs, acc_t* acc) {
  const int v2 = v->cfg.kind == QO_DOCK_V2;
  const qo_params* P = &v->P;
  const double QO_MASS = P->mass, QO_G = P->gravity, QO_L = P->arm_length;
  double* sc = v->chaser + 13 * i;
  double* st = v->target + 13 * i;
  v->t[i] += 1; /* docking_env.py:108 */
  /* docking_env.py:98-99,115: rotor2control @ (action_std*a + action_mean); rotor2control quadrotor.py:56-59 */
  const double half = 1.0 * QO_MASS * QO_G / 2.0, lam = P->motor_lambda;
  double f[4], u_c[4], u_t[4];
  for (int k = 0; k < 4; k++) f[k] = half * a[k] + half;
  u_c[0] = f[0] + f[1] + f[2] + f[3];
  u_c[1] = 0 * f[0] + QO_L * f[1] + 0 * f[2] + (-QO_L) * f[3];
  u_c[2] = (-QO_L) * f[0] + 0 * f[1] + QO_L * f[2] + 0 * f[3];
  u_c[3] = lam * f[0] + (-lam) * f[1] + lam * f[2] + (-lam) * f[3];
  /* target_state_des: docking_env.py:60-64 / moving_docking_env.py:59-66 ; des[6:10] persists (Q4) */
  double des[13] = {P->target_des_pos[0], P->target_des_pos[1], P->target_des_pos[2], P->target_des_vel[0], P->target_des_vel[1],
                    P->target_des_vel[2], 0, 0, 0, 0, 0, 0, 0};
  memcpy(des + 6, v->desq + 4 * i, 4 * sizeof(double));
  if (v2) vel_controller_p(P, des, st, st, u_t); /* moving_docking_env.py:117,126 -- last IS now (Q5) */
  else pid_p(P, des, st, u_t);                   /* docking_env.py:119 */
  memcpy(v->desq + 4 * i, des + 6, 4 * sizeof(double));
  drone_step_p(P, st, v->ut + 4 * i, u_t); /* :120 */
  drone_step_p(P, sc, v->uc + 4 * i, u_c); /* :121 */
  double rel[12];
  state2rel_p(P, sc, st, rel); /* :124-127 */
  const double d10 = deg2rad(P->dock_angle_deg);
  int docking = norm3(rel) < P->dock_pos_tol && norm3(rel + 3) < P->dock_vel_tol && fabs(rel[6]) < d10 && fabs(rel[7]) < d10 && fabs(rel[8]) < d10;
  const double radius = P->range_limit; /* docking_env.py:141 (3) / moving_docking_env.py:148 (10) */
  int overlimit = norm3(rel) >= radius || sc[2] <= P->min_altitude;
  int overtime = v->t[i] >= P->max_steps;
  int dn = overlimit || overtime;
  double ra = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]);
  double x[3], y[3];
  for (int k = 0; k < 3; k++) { x[k] = rel[k] / radius; y[k] = rel[6 + k] / QO_PI; }
  /* docking_env.py:215-222 / moving_docking_env.py:176-183 */
  double shaping = -10.0 * norm3(x) - 1.0 * norm3(rel + 3) - 10.0 * norm3(y) - 1.0 * norm3(rel + 9) - 0.1 * ra +
                   1.0 * (docking ? 1.0 : 0.0);
  double r = shaping - v->last_shaping[i];
  v->last_shaping[i] = shaping;
  v->ep_return[i] += r;
  *reward = r;
  *done = (uint8_t)dn;
  if (info) *info = (uint8_t)((dn ? QO_INFO_DONE : 0) | (overlimit ? QO_INFO_OVERLIMIT : 0) | (docking ? QO_INFO_DOCKING : 0) |
                              (overtime ? QO_INFO_OVERTIME : 0));
  acc->ndock += docking;
  int finite = isfinite(r);
  for (int k = 0; k < 12; k++) finite = finite && isfinite(rel[k]);
  acc->nnonfinite += !finite;
  if (term_obs) memcpy(term_obs, rel, sizeof rel);
  if (dn && v->cfg.auto_reset) { /* SB VecEnv worker auto-reset */
    acc->ret += v->ep_return[i]; acc->len += (double)v->t[i]; acc->neps += 1;
    v->episode[i] += 1;
    env_seat(v, i, obs);
  } else {
    memcpy(obs, rel, sizeof rel);
  }
}

/* gym_docking/envs/hovering_env.py:47-82 */
static void hover_step_one(qo_vec* v, int64_t i, const double* a, double* obs, double* reward, uint8_t* done,
                           uint8_t* info, double* term_obs, acc_t* acc) {
  double* s = v->chaser + 13 * i;
  const qo_params* P = &v->P;
  const double QO_MASS = P->mass, QO_G = P->gravity, QO_L = P->arm_length;
  v->t[i] += 1;
  const double amax = 1.0 * QO_MASS * QO_G, lam = P->motor_lambda;
  double f[4], u[4];
  for (int k = 0; k < 4; k++) f[k] = amax * a[k];
  u[0] = f[0] + f[1] + f[2] + f[3];
  u[1] = 0 * f[0] + QO_L * f[1] + 0 * f[2] + (-QO_L) * f[3];
  u[2] = (-QO_L) * f[0] + 0 * f[1] + QO_L * f[2] + 0 * f[3];
  u[3] = lam * f[0] + (-lam) * f[1] + lam * f[2] + (-lam) * f[3];
  drone_step_p(P, s, v->uc + 4 * i, u);
  const double des_p[3] = {0, 0, 5};
  double pe[3], ve[3], ae[3], we[3], ed[3], en[3];
  const double qd[4] = {1, 0, 0, 0};
  qo_quat2euler(qd, ed);
  qo_quat2euler(s + 6, en);
  for (int k = 0; k < 3; k++) { pe[k] = des_p[k] - s[k]; ve[k] = 0 - s[3 + k]; ae[k] = ed[k] - en[k]; we[k] = 0 - s[10 + k]; }
  double r_thre = (norm3(pe) < 0.1 && norm3(ve) < 0.1) ? 1.0 : 0.0;
  int dn = norm3(s) > P->range_limit || norm3(s + 3) > P->range_limit;
  double r;
  if (!dn) r = r_thre + 0.1 - 0.01 * norm3(pe) - 0.001 * norm3(ve) - 0.01 * norm3(ae) - 0.001 * norm3(we);
  else r = -0.1;
  v->ep_return[i] += r;
  *reward = r;
  *done = (uint8_t)dn;
  if (info) *info = (uint8_t)((dn ? (QO_INFO_DONE | QO_INFO_OVERLIMIT) : 0) | (r_thre > 0 ? QO_INFO_DOCKING : 0));
  acc->ndock += r_thre > 0;
  int finite = isfinite(r);
  for (int k = 0; k < 13; k++) finite = finite && isfinite(s[k]);
  acc->nnonfinite += !finite;
  if (term_obs) memcpy(term_obs, s, 13 * sizeof(double));
  if (dn && v->cfg.auto_reset) {
    acc->ret += v->ep_return[i]; acc->len += (double)v->t[i]; acc->neps += 1;
    v->episode[i] += 1;
    env_seat(v, i, obs);
  } else {
    memcpy(obs, s, 13 * sizeof(double));
  }
}

void qo_vec_step(qo_vec* v, const double* actions, double* obs, double* reward, uint8_t* done, uint8_t* info,
                 double* term_obs, int nthreads) {
  const int64_t n = v->cfg.n_envs;
  const int od = v->obs_dim;
  double ret = 0, len = 0, neps = 0, ndock = 0, nnf = 0;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static) reduction(+ : ret, len, neps, ndock, nnf)
  for (int64_t i = 0; i < n; i++) {
    acc_t acc = {0, 0, 0, 0, 0};
    if (v->cfg.kind == QO_HOVER)
      hover_step_one(v, i, actions + 4 * i, obs + od * i, reward + i, done + i, info ? info + i : NULL,
                     term_obs ? term_obs + od * i : NULL, &acc);
    else
      dock_step_one(v, i, actions + 4 * i, obs + od * i, reward + i, done + i, info ? info + i : NULL,
                    term_obs ? term_obs + od * i : NULL, &acc);
    ret += acc.ret; len += acc.len; neps += acc.neps; ndock += acc.ndock; nnf += acc.nnonfinite;
  }
  v->stats[0] += ret; v->stats[1] += len; v->stats[2] += neps; v->stats[3] += ndock; v->stats[4] += nnf;
}

void qo_vec_stats(qo_vec* v, double out[5]) {
  memcpy(out, v->stats, sizeof v->stats);
  memset(v->stats, 0, sizeof v->stats);
}

/* rl_baselines/ppo2/ppo2.py:511-521.  numpy's promotion rules make the reference evaluate
 * `mb_rewards[step] + self.gamma * nextvalues * nextnonterminal - mb_values[step]` as: gamma*nextvalues in float32 (python
 * scalar times float32 array), the rest in float64 because nextnonterminal = 1.0 - <bool array> is float64; last_gae_lam
 * stays float64 across iterations; mb_advs[step] is stored float32; mb_returns = mb_advs + mb_values in float32.
 * gamma and lam are python floats (doubles); gamma is rounded to float32 only inside gamma*nextvalues. */
void qo_gae(int64_t T, int64_t n, const float* rew, const float* val, const uint8_t* dones, const float* last_val,
            const uint8_t* last_dones, double gamma, double lam, float* adv, float* ret) {
  const double gl = gamma * lam;
  const float gf = (float)gamma;
  for (int64_t i = 0; i < n; i++) {
    double nextnonterminal = 1.0 - (double)last_dones[i];
    float nextvalues = last_val[i];
    double lastgaelam = 0.0;
    for (int64_t t = T - 1; t >= 0; t--) {
      volatile float gv = gf * nextvalues;
      double delta = ((double)rew[t * n + i] + (double)gv * nextnonterminal) - (double)val[t * n + i];
      lastgaelam = delta + (gl * nextnonterminal) * lastgaelam;
      adv[t * n + i] = (float)lastgaelam;
      volatile float r = adv[t * n + i] + val[t * n + i];
      ret[t * n + i] = r;
      nextnonterminal = 1.0 - (double)dones[t * n + i];
      nextvalues = val[t * n + i];
    }
  }
}

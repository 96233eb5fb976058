// qs_core.cuh -- per-environment arithmetic of the QuadSim hot path, written once for fp32 and fp64.
//
// Every function cites the reference code (sd196821/QuadSim, file:line) whose behaviour it reproduces; the
// semantics are the ones SURVEY.md section 0.1 / Appendix A pin down by running the reference:
//   Q1  Drone.step == one explicit Euler step with the RHS frozen at (state, u_prev)   quadrotor.py:115-134
//   Q2  the new command is stored AFTER integrating (one step of control latency)       quadrotor.py:132-142
//   Q3  "quat2rot" uses an elementwise square: unit diagonal, 6 off-diagonals            quadrotor.py:243
//   Q4  the controller's desired quaternion persists across steps and episodes          PIDController.py:100-102
//   Q5  docking-v2's derivative term is identically zero                                moving_docking_env.py:117,126
//
// Two arithmetic modes, selectable per precision at build time (kFast<T>):
//   "fast" (default for float and double) -- algebraically identical shortcuts that remove transcendentals from the
//                common path (threshold tests on sines/tangents instead of angles, rsqrt normalisation, sin/cos of an
//                atan2 taken from its arguments, quat2euler(euler2quat(.)) elided on its identity range).
//                Parity bars: fp64 1e-9 with exact flags, fp32 1e-4 * max(1, |ref|) -- both tested on the same tapes.
//   "literal" (-DQS_LITERAL64 / -DQS_LITERAL32) -- follows the reference's formulas step by step, for A/B checks.
//   Precision-specific extras (the origin shift, 2-ulp division/sqrt, the in-line sincos) apply to fp32 only.
//
// The header also compiles as plain C++ (tests/host_harness.cpp) so tolerances can be explored without a
// GPU; that build is test-only and never loaded by the package.
#pragma once
#include <stdint.h>
#include <math.h>

#include "../../include/quadsim_b200.h"  // qs_env_params_t

#if defined(__CUDACC__)
#define QS_HD __host__ __device__ __forceinline__
#else
#define QS_HD inline
#endif

namespace qs {

template <typename T> constexpr bool is_f32_v() { return sizeof(T) == 4; }

// ---- physical constants: dynamics/quadrotor.py:10-63 ; gains: controller/PIDController.py:10-44 ----
// K<T>: the reference's literals.  Only pi is read from it directly; everything a caller may override goes through a
// *parameter provider* P (first argument of every model function below):
//   DefaultParams<T, KIND>  every member a static constexpr -> the literals end up as immediates, exactly as before
//                           parameters existed (handles created with use_params = 0);
//   RuntimeParams<T>        the same member names as plain data, filled on the host from qs_env_params_t
//                           (make_runtime_params) and handed to the kernel as a __grid_constant__ parameter, i.e. read
//                           from the constant bank as an instruction operand (handles created with use_params = 1).
// Both spell `p.mass`, so the model code is written once.  Derived members (reciprocals, m*g, thresholds of the
// transcendental-free clamp pre-test ...) are part of the provider so that the division / sin / tan happens once on the host.
template <typename T> struct K {
  static constexpr T dt = T(0.02), g = T(9.81), mass = T(0.18), L = T(0.086);
  static constexpr T I0 = T(0.00025), I1 = T(0.000232), I2 = T(0.0003738);
  static constexpr T lam = T(1.5e-9 / 6.11e-8);
  static constexpr T mg = T(0.18 * 9.81);
  static constexpr T pi = T(3.141592653589793);
  static constexpr T d85 = T(85 * 3.141592653589793 / 180), d175 = T(175 * 3.141592653589793 / 180),
                     d10 = T(10 * 3.141592653589793 / 180);
  static constexpr T kp_roll = T(-10), kp_pitch = T(-10), kp_yaw = T(-9.5);
  static constexpr T kd_roll = T(5.1), kd_pitch = T(5.1), kd_yaw = T(4.0);
  static constexpr T kp_x = T(-1), kp_y = T(-1), kp_z = T(50), kd_x = T(-1.65), kd_y = T(-1.65), kd_z = T(8);
  static constexpr T kp_vx = T(-0.7), kp_vy = T(-0.7), kp_vz = T(1);
};

constexpr int kHover = 0, kDockV0 = 1, kDockV2 = 2;                 // == qs_env_kind

// Docking positions are ~50 m from the world origin (y = -50) while everything the env computes from them is a
// difference of a few metres.  The fp32 path therefore keeps both drones' positions RELATIVE to the fixed point
// O = the target's start (10,-50,5) (== docking-v0 set-point, docking_env.py:47,59): the dynamics are translation invariant,
// so this is the same trajectory with ~10x smaller position rounding.  The fp64 path stays in world coordinates
// (literal).  State export/import (always float64) converts, exactly, at the boundary.
template <typename T> struct Origin { static constexpr T x = T(0), y = T(0), z = T(0); };
#ifndef QS_LITERAL32
template <> struct Origin<float> { static constexpr float x = 10.f, y = -50.f, z = 5.f; };
#endif

// arithmetic type the constants of a provider are folded in (sdouble, the speculative fp64 type below, folds in double)
struct sdouble;
template <typename T> struct Arith { using type = T; };
template <> struct Arith<sdouble> { using type = double; };

template <typename T, int KIND = kDockV2> struct DefaultParams {
  using A = typename Arith<T>::type;
  using C = K<A>;
  using O = Origin<A>;
  static constexpr bool custom = false;
  // plant (quadrotor.py:10-63)
  static constexpr T dt = T(C::dt), mass = T(C::mass), g = T(C::g), L = T(C::L), lam = T(C::lam), mg = T(C::mg);
  static constexpr T inv_m = T(A(1.0) / C::mass), mass_x_g = T(C::mass * C::g), inv_g = T(A(1) / C::g), hl = T(A(0.5 / 0.086));
  static constexpr T I0 = T(C::I0), I1 = T(C::I1), I2 = T(C::I2), inv_I0 = T(A(1) / C::I0), inv_I1 = T(A(1) / C::I1), inv_I2 = T(A(1) / C::I2);
  // attitude_limit (quadrotor.py:146-168): limits in radians; sin / tan forms for the transcendental-free pre-test
  static constexpr T lim_roll = T(C::d85), lim_pitch = T(C::d85), lim_yaw = T(C::d175);
  static constexpr T s_roll = T(A(0.99619469809174555)), t_pitch = T(A(11.430052302761343)), t_yaw = T(A(0.087488663525924007));
  // controller gains (PIDController.py:10-44)
  static constexpr T kp_roll = T(C::kp_roll), kp_pitch = T(C::kp_pitch), kp_yaw = T(C::kp_yaw);
  static constexpr T kd_roll = T(C::kd_roll), kd_pitch = T(C::kd_pitch), kd_yaw = T(C::kd_yaw);
  static constexpr T kp_x = T(C::kp_x), kp_y = T(C::kp_y), kp_z = T(C::kp_z), kd_x = T(C::kd_x), kd_y = T(C::kd_y), kd_z = T(C::kd_z);
  static constexpr T kp_vx = T(C::kp_vx), kp_vy = T(C::kp_vy), kp_vz = T(C::kp_vz);
  // environment (docking_env.py:34-99,130-152 ; moving_docking_env.py:62-66,148 ; hovering_env.py:23,42,68)
  static constexpr T half = T(A(1.0 * 0.18 * 9.81 / 2.0));                // action_mean = action_std = m g / 2
  static constexpr T port = T(A(0.1)), port2 = T(A(-0.2));                // chaser port +d, target port -d; -d - d
  static constexpr T tol_pos = T(A(0.1)), tol_vel = T(A(0.1)), d10 = T(C::d10);
  static constexpr T radius = T(KIND == kHover ? A(100) : (KIND == kDockV0 ? A(3) : A(10)));
  static constexpr T ten_over_radius = T(A(10) / (KIND == kDockV0 ? A(3) : A(10)));
  static constexpr T z_floor = T(A(0.1 - (double)O::z));
  static constexpr int max_steps = 600;
  static constexpr T ox = T(O::x), oy = T(O::y), oz = T(O::z);
  static constexpr float cini_x = KIND == kHover ? 0.f : 8.f, cini_y = KIND == kHover ? 0.f : -50.f, cini_z = 5.f;
  static constexpr T tini_x = T(A(10) - O::x), tini_y = T(A(-50) - O::y), tini_z = T(A(5) - O::z);
  static constexpr T des_px = T(A(10) - O::x), des_py = T(A(-50) - O::y), des_pz = T(A(5) - O::z);
  static constexpr T des_vx = T(KIND == kDockV2 ? A(0.2) : A(0)), des_vy = T(A(0)), des_vz = T(A(0));
};

template <typename T> struct RuntimeParams {
  static constexpr bool custom = true;  // "not the compiled-in immediates"
  T dt, mass, g, L, lam, mg, inv_m, mass_x_g, inv_g, hl, I0, I1, I2, inv_I0, inv_I1, inv_I2;
  T lim_roll, lim_pitch, lim_yaw, s_roll, t_pitch, t_yaw;
  T kp_roll, kp_pitch, kp_yaw, kd_roll, kd_pitch, kd_yaw, kp_x, kp_y, kp_z, kd_x, kd_y, kd_z, kp_vx, kp_vy, kp_vz;
  T half, port, port2, tol_pos, tol_vel, d10, radius, ten_over_radius, z_floor;
  T ox, oy, oz, tini_x, tini_y, tini_z, des_px, des_py, des_pz, des_vx, des_vy, des_vz;
  float cini_x, cini_y, cini_z;
  int max_steps;
};

// Host side: qs_env_params_t -> RuntimeParams<T>.  Every derived member is formed the way the constexpr provider forms it
// (products of doubles rounded once to T where the literal was a double expression; T arithmetic where it was T arithmetic),
// so a RuntimeParams built from qs_default_params() reproduces the compiled-in constants bit for bit -- except the three
// pre-test thresholds, which come from libm here and are pulled in by 2^-20 so that the pre-test can only over-trigger
// the exact (out-of-line) clamp, never miss it.
template <typename T> inline void make_runtime_params(const qs_env_params_t& in, int kind, RuntimeParams<T>& p) {
  const double pi = 3.141592653589793;
  p.dt = T(in.dt); p.mass = T(in.mass); p.g = T(in.gravity); p.L = T(in.arm_length); p.lam = T(in.motor_lambda);
  p.mg = T(in.mass * in.gravity);
  p.inv_m = T(1.0) / p.mass; p.mass_x_g = p.mass * p.g; p.inv_g = T(1) / p.g; p.hl = T(0.5 / in.arm_length);
  p.I0 = T(in.inertia[0]); p.I1 = T(in.inertia[1]); p.I2 = T(in.inertia[2]);
  p.inv_I0 = T(1) / p.I0; p.inv_I1 = T(1) / p.I1; p.inv_I2 = T(1) / p.I2;
  const double lr = in.att_limit_deg[0] * pi / 180, lp = in.att_limit_deg[1] * pi / 180, ly = in.att_limit_deg[2] * pi / 180;
  p.lim_roll = T(lr); p.lim_pitch = T(lp); p.lim_yaw = T(ly);
  const double shrink = 1.0 - 9.5367431640625e-07;
  p.s_roll = T(sin(lr) * shrink); p.t_pitch = T(tan(lp) * shrink); p.t_yaw = T(tan(pi - ly) / shrink);
  p.kp_roll = T(in.kp_att[0]); p.kp_pitch = T(in.kp_att[1]); p.kp_yaw = T(in.kp_att[2]);
  p.kd_roll = T(in.kd_att[0]); p.kd_pitch = T(in.kd_att[1]); p.kd_yaw = T(in.kd_att[2]);
  p.kp_x = T(in.kp_pos[0]); p.kp_y = T(in.kp_pos[1]); p.kp_z = T(in.kp_pos[2]);
  p.kd_x = T(in.kd_pos[0]); p.kd_y = T(in.kd_pos[1]); p.kd_z = T(in.kd_pos[2]);
  p.kp_vx = T(in.kp_vel[0]); p.kp_vy = T(in.kp_vel[1]); p.kp_vz = T(in.kp_vel[2]);
  p.half = T(1.0 * in.mass * in.gravity / 2.0);
  p.port = T(in.dock_port); p.port2 = -p.port - p.port;
  p.tol_pos = T(in.dock_pos_tol); p.tol_vel = T(in.dock_vel_tol); p.d10 = T(in.dock_angle_deg * pi / 180);
  p.radius = T(in.range_limit); p.ten_over_radius = T(10) / p.radius;
  const bool shifted = is_f32_v<T>() && kind != kHover;  // the fp32 docking layout is relative to the target's start position
  const float o[3] = {shifted ? (float)in.target_ini_pos[0] : 0.f, shifted ? (float)in.target_ini_pos[1] : 0.f, shifted ? (float)in.target_ini_pos[2] : 0.f};
  p.ox = T(o[0]); p.oy = T(o[1]); p.oz = T(o[2]);
  p.z_floor = T(in.min_altitude - (double)o[2]);
  p.cini_x = (float)in.chaser_ini_pos[0]; p.cini_y = (float)in.chaser_ini_pos[1]; p.cini_z = (float)in.chaser_ini_pos[2];
  p.tini_x = T((float)in.target_ini_pos[0]) - p.ox; p.tini_y = T((float)in.target_ini_pos[1]) - p.oy; p.tini_z = T((float)in.target_ini_pos[2]) - p.oz;
  p.des_px = T(in.target_des_pos[0]) - p.ox; p.des_py = T(in.target_des_pos[1]) - p.oy; p.des_pz = T(in.target_des_pos[2]) - p.oz;
  p.des_vx = T(in.target_des_vel[0]); p.des_vy = T(in.target_des_vel[1]); p.des_vz = T(in.target_des_vel[2]);
  p.max_steps = in.max_steps;
}

template <typename T> struct is_f32 { static constexpr bool value = false; };
template <> struct is_f32<float> { static constexpr bool value = true; };
// kFast<T>: algebraically identical shortcuts that take transcendentals off the common path (see the functions below).
// On by default in both precisions -- they pass the same parity bars (fp64: 1e-9 and exact flags against the reference tapes,
// the wild single-step states and the oracle; fp32: 1e-4) -- and are worth 1.13x (fp32) / 1.49x (fp64) on the step kernel.
// -DQS_LITERAL32 / -DQS_LITERAL64 build the literal transcription (reference formulas in reference order) for A/B.
#ifdef QS_LITERAL32
constexpr bool kFast32 = false;
#else
constexpr bool kFast32 = true;
#endif
#ifdef QS_LITERAL64
constexpr bool kFast64 = false;
#else
constexpr bool kFast64 = true;
#endif
template <typename T> constexpr bool kFast = is_f32<T>::value ? kFast32 : kFast64;  // sdouble (below): kFast64

// ---- fp64 elementary functions of the fast fp64 path (qs::f64m) ----
// The library's fp64 atan2 / sincos / 1/sqrt cost ~115 / ~68 / ~43 SASS instructions per call inside the step kernel, a
// third of them UMOV/IMAD.MOV pairs that materialise 64-bit polynomial coefficients (profiles/r1_dock_step_f64_fast_*).
// These versions keep the coefficients in constant memory (operands of the DFMA itself), split each polynomial into two
// interleaved Horner chains, and use the MUFU.RCP64H / MUFU.RSQ64H seeds with branch-free Newton steps.  Measured on
// 4e6 random arguments against the host libm (tests/test_core_arith_host.py): <= 2 ulp.  Coefficients:
// tools/gen_f64_math_coefs.py (Chebyshev fits at 60 digits).  Arguments outside the plain range (zero, subnormal, huge,
// non-finite) defer to the library, so the IEEE special cases are the library's.
#define QS_F64M_ATAN {-0x1.5555555555555p-2, 0x1.999999999934ap-3, -0x1.24924924360cbp-3, 0x1.c71c7185314cbp-4, -0x1.745d0b26b83e7p-4, 0x1.3b1262d95579ep-4, -0x1.10fa75382537fp-4, 0x1.dfe61e80903d2p-5, -0x1.a098bb6ba4941p-5, 0x1.41603647c7a7cp-5, -0x1.3a2b7a07caea9p-6}
#define QS_F64M_SIN {-0x1.5555555555555p-3, 0x1.1111111110babp-7, -0x1.a01a019e820aep-13, 0x1.71de37946ed6dp-19, -0x1.ae6008d114485p-26, 0x1.5e0a4fc16be59p-33}
#define QS_F64M_COS {0x1.5555555555555p-5, -0x1.6c16c16c16963p-10, 0x1.a01a019f4ddfdp-16, -0x1.27e4fa16cc2d1p-22, 0x1.1eeb67d026ec6p-29, -0x1.907cd8ff5ede7p-37}
#define QS_F64M_ASIN {0x1.5555555555556p-3, 0x1.3333333332ec4p-4, 0x1.6db6db6e3292ep-5, 0x1.f1c71c1d4aa0dp-6, 0x1.6e8bb1d89492bp-6, 0x1.1c4d343fa57f6p-6, 0x1.c9cf3759fd8ffp-7, 0x1.78246f32d075cp-7, 0x1.524ea8b3aad31p-7, 0x1.653a8f8cfd43cp-8, 0x1.1d661531b222ep-6, -0x1.e7a24f548c99fp-7, 0x1.d78189767524ep-6}
namespace f64m {
#if defined(__CUDACC__)
static __constant__ double c_asin[13] = QS_F64M_ASIN;
static __constant__ double c_atan[11] = QS_F64M_ATAN;
static __constant__ double c_sin[6] = QS_F64M_SIN;
static __constant__ double c_cos[6] = QS_F64M_COS;
#endif
static constexpr double h_asin[13] = QS_F64M_ASIN;
static constexpr double h_atan[11] = QS_F64M_ATAN;
static constexpr double h_sin[6] = QS_F64M_SIN;
static constexpr double h_cos[6] = QS_F64M_COS;
#if defined(__CUDA_ARCH__)
#define QS_F64M_C(name, i) c_##name[i]
#else
#define QS_F64M_C(name, i) h_##name[i]
#endif

QS_HD double qnan() {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(0x7ff8000000000000LL);
#else
  return NAN;
#endif
}
// ok ? v : NaN as an opaque select: written as a plain conditional, the compiler sinks the whole evaluation of v into
// a branch on ok -- exactly the basic-block split the speculative path exists to avoid
QS_HD double keep(bool ok, double v) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("{ .reg .pred p; setp.ne.s32 p, %2, 0; selp.f64 %0, %1, %3, p; }" : "=d"(r) : "d"(v), "r"((int)ok), "d"(qnan()));
  return r;
#else
  return ok ? v : qnan();
#endif
}
// library fallbacks, one out-of-line copy each (cold: zero / subnormal / huge / non-finite arguments only)
#if defined(__CUDACC__)
#define QS_F64M_COLD static __host__ __device__ __noinline__
#else
#define QS_F64M_COLD inline
#endif
struct SC { double s, c; };
QS_F64M_COLD double atan2_cold(double y, double x) { return ::atan2(y, x); }
QS_F64M_COLD double sqrt_cold(double x) { return ::sqrt(x); }
QS_F64M_COLD SC sincos_cold(double x) {
  SC r;
#if defined(__CUDA_ARCH__)
  ::sincos(x, &r.s, &r.c);
#else
  r.s = ::sin(x); r.c = ::cos(x);
#endif
  return r;
}

// 1/d for d in [1e-280, 1e280]: 20-bit seed + two Newton steps (relative error ~2^-52)
QS_HD double rcp(double d) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / d;
#endif
}
// n/d, d in [1e-280, 1e280], |n| <= |d|: one residual correction on top of n * rcp(d) (<= 1 ulp)
QS_HD double div(double n, double d) {
#if defined(__CUDA_ARCH__)
  const double r = rcp(d);
  const double q = n * r;
  return fma(fma(-d, q, n), r, q);
#else
  return n / d;
#endif
}
// 1/sqrt(x), x in [1e-300, 1e300]
QS_HD double rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  // y (1 + e/2 + 3e^2/8), e = 1 - x y^2, twice: 2^-20 -> 2^-57
  double e = fma(-x * y, y, 1.0);
  y = fma(y * e, fma(0.375, e, 0.5), y);
  e = fma(-x * y, y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
#else
  return 1.0 / ::sqrt(x);
#endif
}
QS_HD double sqrt(double x) {
#if defined(__CUDA_ARCH__)
  const double y = rsqrt(x);
  const double g = x * y;
  return fma(fma(-g, g, x), 0.5 * y, g);
#else
  return ::sqrt(x);
#endif
}

// sqrt for x == 0 or x in [1e-300, 1e300], branch-free (the seed is taken of max(x, 1e-300): 0 * finite = 0);
// negative / NaN x -> NaN
QS_HD double sqrt0(double x) {
#if defined(__CUDA_ARCH__)
  const double y = rsqrt(fmax(x, 1e-300));
  const double g = x * y;
  const double r = fma(fma(-g, g, x), 0.5 * y, g);
  return keep(!(x < 0.0), r);
#else
  return ::sqrt(x);
#endif
}
// atan2 for max(|x|,|y|) in [1e-280, 1e280] (anything else: garbage, the callers below guard or poison)
QS_HD double atan2_fast(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const bool sw = ay > ax;
  const double a = sw ? ax : ay, b = sw ? ay : ax;  // a = min, b = max (a NaN lands in b or in the quotient)
  // |a/b| <= 1 ; above tan(pi/8) use atan(a/b) = pi/4 + atan((a-b)/(a+b)): still one division, |u| <= tan(pi/8)
  const bool hi = a > 0x1.a827999fcef32p-2 * b;
  const double u = div(hi ? a - b : a, hi ? a + b : b);
  const double z = u * u, z2 = z * z;
  double pe = fma(z2, QS_F64M_C(atan, 10), QS_F64M_C(atan, 8));
  double po = fma(z2, QS_F64M_C(atan, 9), QS_F64M_C(atan, 7));
  pe = fma(z2, pe, QS_F64M_C(atan, 6)); po = fma(z2, po, QS_F64M_C(atan, 5));
  pe = fma(z2, pe, QS_F64M_C(atan, 4)); po = fma(z2, po, QS_F64M_C(atan, 3));
  pe = fma(z2, pe, QS_F64M_C(atan, 2)); po = fma(z2, po, QS_F64M_C(atan, 1));
  pe = fma(z2, pe, QS_F64M_C(atan, 0));
  double r = fma(u * z, fma(z, po, pe), u);
  if (hi) r = 0x1.921fb54442d18p-1 + (0x1.1a62633145c07p-55 + r);
  if (sw) r = 0x1.921fb54442d18p+0 - (r - 0x1.1a62633145c07p-54);
  if (x < 0.0) r = 0x1.921fb54442d18p+1 - (r - 0x1.1a62633145c07p-53);
  return copysign(r, y);
}
QS_HD bool atan2_in_range(double y, double x) {
  const double b = fmax(fabs(x), fabs(y));  // fmax drops a NaN operand: a NaN argument still reaches the quotient
  return b >= 1e-280 && b <= 1e280;
}
QS_HD double atan2(double y, double x) {
  if (!atan2_in_range(y, x)) return atan2_cold(y, x);
  return atan2_fast(y, x);
}

// asin(x) = u + u z R(z) with (u, z) = (x, x^2) for |x| <= 1/2, else pi/2 - 2 asin(sqrt((1 - |x|) / 2)); |x| > 1 -> NaN
template <bool GUARDED> QS_HD double asin_t(double x) {
  const double ax = fabs(x);
  const bool big = ax > 0.5;
  const double w = fma(-0.5, ax, 0.5);
  const double z = big ? w : x * x;
  double u = ax;
  if (GUARDED) { if (big) u = (w >= 1e-300) ? sqrt(w) : sqrt_cold(w); }  // w = 0 at |x| = 1, negative (NaN) beyond
  else u = big ? sqrt0(w) : ax;                                        // branch-free; |x| > 1 -> NaN as well
  const double z2 = z * z;
  double pe = fma(z2, QS_F64M_C(asin, 12), QS_F64M_C(asin, 10));
  double po = fma(z2, QS_F64M_C(asin, 11), QS_F64M_C(asin, 9));
  pe = fma(z2, pe, QS_F64M_C(asin, 8)); po = fma(z2, po, QS_F64M_C(asin, 7));
  pe = fma(z2, pe, QS_F64M_C(asin, 6)); po = fma(z2, po, QS_F64M_C(asin, 5));
  pe = fma(z2, pe, QS_F64M_C(asin, 4)); po = fma(z2, po, QS_F64M_C(asin, 3));
  pe = fma(z2, pe, QS_F64M_C(asin, 2)); po = fma(z2, po, QS_F64M_C(asin, 1));
  pe = fma(z2, pe, QS_F64M_C(asin, 0));
  double r = fma(u * z, fma(z, po, pe), u);
  if (big) r = 0x1.921fb54442d18p+0 - (2.0 * r - 0x1.1a62633145c07p-54);
  return copysign(r, x);
}
QS_HD double asin(double x) { return asin_t<true>(x); }

// sin and cos for |x| <= 1e5
QS_HD void sincos_fast(double x, double* s, double* c) {
  // k = nearest integer to x * 2/pi via the 1.5 * 2^52 shift (its low word is k mod 2^32)
  const double t = fma(x, 0x1.45f306dc9c883p-1, 6755399441055744.0);
  const double k = t - 6755399441055744.0;
#if defined(__CUDA_ARCH__)
  const int q = __double2loint(t);
#else
  const int q = (int)k;
#endif
  double r = fma(k, -0x1.921fb54400000p+0, x);
  r = fma(k, -0x1.0b4611a600000p-34, r);
  r = fma(k, -0x1.3198a2e037073p-69, r);
  const double z = r * r;
  double ps = fma(z, QS_F64M_C(sin, 5), QS_F64M_C(sin, 4)), pc = fma(z, QS_F64M_C(cos, 5), QS_F64M_C(cos, 4));
  ps = fma(z, ps, QS_F64M_C(sin, 3)); pc = fma(z, pc, QS_F64M_C(cos, 3));
  ps = fma(z, ps, QS_F64M_C(sin, 2)); pc = fma(z, pc, QS_F64M_C(cos, 2));
  ps = fma(z, ps, QS_F64M_C(sin, 1)); pc = fma(z, pc, QS_F64M_C(cos, 1));
  ps = fma(z, ps, QS_F64M_C(sin, 0)); pc = fma(z, pc, QS_F64M_C(cos, 0));
  const double sn = fma(r * z, ps, r);
  const double cs = fma(z * z, pc, fma(-0.5, z, 1.0));
  const double a = (q & 1) ? cs : sn, b = (q & 1) ? sn : cs;
  *s = (q & 2) ? -a : a;
  *c = ((q + 1) & 2) ? -b : b;
}
QS_HD void sincos(double x, double* s, double* c) {
  if (!(fabs(x) <= 1.0e5)) { const SC r = sincos_cold(x); *s = r.s; *c = r.c; return; }
  sincos_fast(x, s, c);
}
}  // namespace f64m

// ---- scalar math wrappers ----
QS_HD float m_sqrt(float x) { return sqrtf(x); }
QS_HD double m_sqrt(double x) {
  if (kFast64) return (x == 0.0 || (x >= 1e-300 && x <= 1e300)) ? f64m::sqrt0(x) : f64m::sqrt_cold(x);  // 0: norms on an episode's first step
  return sqrt(x);
}
QS_HD float m_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
QS_HD double m_rsqrt(double x) {
  if (kFast64) return (x >= 1e-300 && x <= 1e300) ? f64m::rsqrt(x) : 1.0 / f64m::sqrt_cold(x);
  return 1.0 / sqrt(x);
}
QS_HD float m_abs(float x) { return fabsf(x); }
QS_HD double m_abs(double x) { return fabs(x); }
QS_HD float m_asin(float x) { return asinf(x); }
#if defined(__CUDACC__) && defined(QS_F64_OUTLINE)
#define QS_F64_FN __host__ __device__ __noinline__
#else
#define QS_F64_FN QS_HD
#endif
QS_F64_FN double m_asin(double x) { return kFast64 ? f64m::asin(x) : asin(x); }
QS_HD float m_atan2(float y, float x) { return atan2f(y, x); }
QS_F64_FN double m_atan2(double y, double x) { return kFast64 ? f64m::atan2(y, x) : atan2(y, x); }
QS_HD float m_tan(float x) { return tanf(x); }
QS_HD double m_tan(double x) { return tan(x); }
// fp32 sin & cos together.  Every angle the envs feed in is bounded (half-angles of atan2 results, desired tilt = acc/g),
// so the common path is a 3-constant Cody-Waite reduction to [-pi/4, pi/4] plus the Cephes single-precision minimax
// polynomials (~1-2 ulp there): straight-line code, no Payne-Hanek slow path inlined at every call site (that cost ~40 %
// of the kernel's code size and showed up as instruction-cache stalls).  |x| > 1e4 (only reachable from injected wild
// states) defers to the library.
QS_HD void m_sincos_big(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  sincosf(x, s, c);
#else
  *s = sinf(x); *c = cosf(x);
#endif
}
QS_HD void m_sincos(float x, float* s, float* c) {
#ifdef QS_LITERAL32
  m_sincos_big(x, s, c);
#else
  if (!(fabsf(x) <= 1.0e4f)) { m_sincos_big(x, s, c); return; }
  const float k = rintf(x * 0.636619772367581343f);  // 2/pi
  float r = fmaf(k, -1.5707962512969971f, x);        // pi/2 = hi + mid + lo
  r = fmaf(k, -7.5497894158615964e-08f, r);
  r = fmaf(k, -5.3903029534742384e-15f, r);
  const float z = r * r;
  const float sn = fmaf(r * z, fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f), r);
  const float cs = fmaf(z * z, fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f), fmaf(-0.5f, z, 1.0f));
  const int q = (int)k;
  const float a = (q & 1) ? cs : sn, b = (q & 1) ? sn : cs;
  *s = (q & 2) ? -a : a;
  *c = ((q + 1) & 2) ? -b : b;
#endif
}
struct SinCosD { double s, c; };
QS_F64_FN SinCosD m_sincos_d(double x) {
  SinCosD r;
  if (kFast64) { f64m::sincos(x, &r.s, &r.c); return r; }
#if defined(__CUDA_ARCH__)
  sincos(x, &r.s, &r.c);
#else
  r.s = sin(x); r.c = cos(x);
#endif
  return r;
}
QS_HD void m_sincos(double x, double* s, double* c) { SinCosD r = m_sincos_d(x); *s = r.s; *c = r.c; }
QS_HD float m_min(float a, float b) { return fminf(a, b); }
QS_HD double m_min(double a, double b) { return fmin(a, b); }
QS_HD float m_max(float a, float b) { return fmaxf(a, b); }
QS_HD double m_max(double a, double b) { return fmax(a, b); }
template <typename T> QS_HD T m_sign(T x) { return T((x > T(0)) - (x < T(0))); }
template <typename T> QS_HD bool m_finite(T x) { return (x - x) == T(0); }
template <typename T> QS_HD T norm3(T x, T y, T z) { return m_sqrt(x * x + y * y + z * z); }

// ---- sdouble: the SPECULATIVE fp64 arithmetic type ----
// The guarded fp64 functions above each carry a range test with an out-of-line library call behind it.  Those branches
// are never taken on sane states, but they cut the step into ~70 basic blocks that ptxas cannot schedule across; on a
// kernel that runs 3 warps per scheduler that costs 3.5 % (profiles/r1_fp64_f64m_spec_ab.log).  sdouble is a double whose
// elementary functions are the same fast cores WITHOUT the branch: an argument outside a core's range yields NaN instead
// ("poison").  dock_env_step<sdouble> is therefore a handful of long straight-line blocks (only the attitude-clamp tests
// remain); dock_env_step_spec() below runs it, checks that nothing non-finite reached the observation, the reward or the
// parts of the state the observation does not see, and otherwise tells the caller to redo the step from the ORIGINAL
// state with the guarded double arithmetic (a cold, out-of-line replay).  In-range arguments take identical code either
// way, so the results are bit-identical to the guarded path (tests/test_core_arith_host.py checks that on 1.6e6 wild
// env-steps); legitimately non-finite states simply always take the replay.  Rules that keep poison from getting lost:
// NaN-preserving min/max (u_limit), an explicit test before the comparisons of the attitude clamp, and f64m::keep() --
// an opaque select -- wherever a precondition is recorded.
struct sdouble {
  double v;
  sdouble() = default;
  constexpr sdouble(double x) : v(x) {}
  constexpr explicit operator double() const { return v; }
};
template <typename T> struct is_spec { static constexpr bool value = false; };
template <> struct is_spec<sdouble> { static constexpr bool value = true; };
template <typename T> constexpr bool kSpec = is_spec<T>::value;
QS_HD sdouble operator+(sdouble a, sdouble b) { return sdouble(a.v + b.v); }
QS_HD sdouble operator-(sdouble a, sdouble b) { return sdouble(a.v - b.v); }
QS_HD sdouble operator*(sdouble a, sdouble b) { return sdouble(a.v * b.v); }
QS_HD sdouble operator/(sdouble a, sdouble b) { return sdouble(a.v / b.v); }
QS_HD sdouble operator-(sdouble a) { return sdouble(-a.v); }
QS_HD sdouble& operator+=(sdouble& a, sdouble b) { a.v += b.v; return a; }
QS_HD sdouble& operator-=(sdouble& a, sdouble b) { a.v -= b.v; return a; }
QS_HD sdouble& operator*=(sdouble& a, sdouble b) { a.v *= b.v; return a; }
QS_HD bool operator<(sdouble a, sdouble b) { return a.v < b.v; }
QS_HD bool operator>(sdouble a, sdouble b) { return a.v > b.v; }
QS_HD bool operator<=(sdouble a, sdouble b) { return a.v <= b.v; }
QS_HD bool operator>=(sdouble a, sdouble b) { return a.v >= b.v; }
QS_HD bool operator==(sdouble a, sdouble b) { return a.v == b.v; }
QS_HD bool operator!=(sdouble a, sdouble b) { return a.v != b.v; }
// ok ? v : NaN -- how a model-level precondition of a shortcut is recorded in speculative mode
QS_HD sdouble spec_keep(sdouble v, bool ok) { return sdouble(f64m::keep(ok, v.v)); }
QS_HD float spec_keep(float v, bool) { return v; }
QS_HD double spec_keep(double v, bool) { return v; }
QS_HD sdouble m_abs(sdouble x) { return sdouble(fabs(x.v)); }
// NaN-preserving min / max (fmin / fmax would swallow the poison on its way through u_limit)
QS_HD sdouble m_min(sdouble a, sdouble b) { return sdouble((a.v < b.v || a.v != a.v) ? a.v : b.v); }
QS_HD sdouble m_max(sdouble a, sdouble b) { return sdouble((a.v > b.v || a.v != a.v) ? a.v : b.v); }
QS_HD sdouble m_sqrt(sdouble x) {  // 0 is a legitimate argument (norms of zero vectors on the first step of an episode)
  return sdouble(f64m::keep(x.v == 0.0 || (x.v >= 1e-300 && x.v <= 1e300), f64m::sqrt0(x.v)));
}
QS_HD sdouble m_rsqrt(sdouble x) { return sdouble(f64m::keep(x.v >= 1e-300 && x.v <= 1e300, f64m::rsqrt(x.v))); }
QS_HD sdouble m_asin(sdouble x) { return sdouble(f64m::asin_t<false>(x.v)); }
QS_HD sdouble m_atan2(sdouble y, sdouble x) {
  return sdouble(f64m::keep(f64m::atan2_in_range(y.v, x.v), f64m::atan2_fast(y.v, x.v)));
}
QS_HD void m_sincos(sdouble x, sdouble* s, sdouble* c) {
  double sn, cs;
  f64m::sincos_fast(x.v, &sn, &cs);
  const bool ok = fabs(x.v) <= 1.0e5;
  *s = sdouble(f64m::keep(ok, sn)); *c = sdouble(f64m::keep(ok, cs));
}
QS_HD sdouble m_tan(sdouble) { return sdouble(f64m::qnan()); }  // only reached from branches the speculative path replaces

template <typename T> struct Euler { T phi, theta, psi; };
template <typename T> struct Quat { T w, x, y, z; };

// One quadrotor: state[13] (quadrotor.py:26) + the previous, already limited command u (Q2).
template <typename T> struct Drone {
  T px, py, pz, vx, vy, vz, q0, q1, q2, q3, wx, wy, wz;
  T u0, u1, u2, u3;
};

// The six off-diagonal entries of the reference's "rotation matrix" (diagonal is exactly 1, Q3).
template <typename T> struct Rot { T r01, r02, r10, r12, r20, r21; };

// utils/transform.py:94-120  quat2euler (un-normalised quaternion in, Z-X-Y angles out)
template <typename T> QS_HD Euler<T> quat2euler(T w, T x, T y, T z) {
  T r10 = T(2) * (x * y - w * z);
  T r11 = w * w - x * x + y * y - z * z;
  T r12 = T(2) * (w * x + y * z);
  T r02 = T(2) * (x * z - w * y);
  T r22 = w * w - x * x - y * y + z * z;
  Euler<T> e;
  e.psi = m_atan2(-r10, r11);
  if (kSpec<T>) {  // straight line; the +-90 degree branches are left to the guarded replay
    e.phi = spec_keep(m_asin(r12), r12 < T(1) && !(r12 < T(-1)));
    e.theta = m_atan2(-r02, r22);
    return e;
  }
  if (r12 < T(1)) {
    if (r12 < T(-1)) { e.phi = -K<T>::pi / T(2); e.theta = T(0); }
    else { e.phi = m_asin(r12); e.theta = m_atan2(-r02, r22); }
  } else { e.phi = K<T>::pi / T(2); e.theta = T(0); }
  return e;
}

// utils/transform.py:123-136  euler2quat
template <typename T> QS_HD Quat<T> euler2quat(T roll, T pitch, T yaw) {
  T cy, sy, cp, sp, cr, sr;
  m_sincos(yaw * T(0.5), &sy, &cy);
  m_sincos(pitch * T(0.5), &sp, &cp);
  m_sincos(roll * T(0.5), &sr, &cr);
  Quat<T> q;
  q.w = cr * cp * cy - sr * sp * sy;
  q.x = sr * cp * cy - cr * sp * sy;
  q.y = sr * cp * sy + cr * sp * cy;
  q.z = cr * cp * sy + sr * sp * cy;
  return q;
}

// utils/transform.py:4-20 == dynamics/quadrotor.py:226-245 (Q3): R = I + 2*h.*h + 2*q0*h with
// h01=-n3 h02=n2 h12=-n1 h10=n3 h20=-n2 h21=n1, n = q/|q| and q0 RAW.
template <typename T> QS_HD Rot<T> quat2rot(T q0, T q1, T q2, T q3) {
  T n1, n2, n3;
  T ss = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
  if (kFast<T>) { T inv = m_rsqrt(ss); n1 = q1 * inv; n2 = q2 * inv; n3 = q3 * inv; }
  else { T n = m_sqrt(ss); n1 = q1 / n; n2 = q2 / n; n3 = q3 / n; }
  T w2 = T(2) * q0;
  Rot<T> R;
  R.r01 = (T(2) * n3) * n3 - w2 * n3;
  R.r10 = (T(2) * n3) * n3 + w2 * n3;
  R.r02 = (T(2) * n2) * n2 + w2 * n2;
  R.r20 = (T(2) * n2) * n2 - w2 * n2;
  R.r12 = (T(2) * n1) * n1 - w2 * n1;
  R.r21 = (T(2) * n1) * n1 + w2 * n1;
  return R;
}

// utils/transform.py:23-46  rot2euler on (R12, R10, R11, R02, R22)
template <typename T> QS_HD Euler<T> rot2euler(T r12, T r10, T r11, T r02, T r22) {
  Euler<T> e;
  e.psi = m_atan2(-r10, r11);
  if (kSpec<T>) {  // straight line; the +-90 degree branches are left to the guarded replay
    e.phi = spec_keep(m_asin(r12), r12 < T(1) && !(r12 < T(-1)));
    e.theta = m_atan2(-r02, r22);
    return e;
  }
  if (r12 < T(1)) {
    if (r12 < T(-1)) { e.phi = -K<T>::pi / T(2); e.theta = T(0); }
    else { e.phi = m_asin(r12); e.theta = m_atan2(-r02, r22); }
  } else { e.phi = K<T>::pi / T(2); e.theta = T(0); }
  return e;
}

// dynamics/quadrotor.py:171-187  u_limit: per-rotor thrust A@[F,Mx,My] clamped to [0, F_max/4 = m*g], Mz untouched
template <typename T, typename P> QS_HD void u_limit(const P& p, T F, T Mx, T My, T Mz, T& o0, T& o1, T& o2, T& o3) {
  const T hl = p.hl, L = p.L, fmax = p.mg;
  T p0 = T(0.25) * F - hl * My;
  T p1 = T(0.25) * F + hl * Mx;
  T p2 = T(0.25) * F + hl * My;
  T p3 = T(0.25) * F - hl * Mx;
  p0 = m_max(m_min(p0, fmax), T(0));
  p1 = m_max(m_min(p1, fmax), T(0));
  p2 = m_max(m_min(p2, fmax), T(0));
  p3 = m_max(m_min(p3, fmax), T(0));
  o0 = ((p0 + p1) + p2) + p3;
  o1 = L * p1 - L * p3;
  o2 = L * p2 - L * p0;
  o3 = Mz;
}
template <typename T> QS_HD void u_limit(T F, T Mx, T My, T Mz, T& o0, T& o1, T& o2, T& o3) {
  u_limit(DefaultParams<T>{}, F, Mx, My, Mz, o0, o1, o2, o3);
}

// rotor thrusts -> [F, Mx, My, Mz]: rotor2control, dynamics/quadrotor.py:56-59
template <typename T, typename P> QS_HD void rotor2control(const P& p, T f0, T f1, T f2, T f3, T& F, T& Mx, T& My, T& Mz) {
  const T L = p.L, lam = p.lam;
  F = ((f0 + f1) + f2) + f3;
  Mx = L * f1 - L * f3;
  My = L * f2 - L * f0;
  Mz = ((lam * f0 - lam * f1) + lam * f2) - lam * f3;
}

// dynamics/quadrotor.py:146-168  attitude_limit + the overwrite in Drone.step (:135-138).  The reference runs three
// independent `if`s, each rebuilding the quaternion from the RAW other two angles and overriding the previous one, then
// cancels everything if all three are within (<=) their limits: the last limit that fires wins (yaw, else pitch, else
// roll).  Cold path (never taken under the benchmark's action distribution), kept out of line on the device.
// out-of-line euler2quat for cold paths (by value in and out: nothing of the caller's has to live in local memory)
template <typename T>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
Quat<T> euler2quat_cold(T roll, T pitch, T yaw) { return euler2quat(roll, pitch, yaw); }

template <typename T> struct ClampResult { T w, x, y, z; bool fired; };
template <typename T, typename P> QS_HD ClampResult<T> attitude_clamp(const P& p_, T q0, T q1, T q2, T q3) {
  const T d85r = p_.lim_roll, d85p = p_.lim_pitch, d175 = p_.lim_yaw;
  ClampResult<T> out = {q0, q1, q2, q3, false};
  Euler<T> e = quat2euler(q0, q1, q2, q3);
  const T r = e.phi, p = e.theta, y = e.psi;
  if (kSpec<T> && !m_finite((r + p) + y)) {  // poisoned angles must not be swallowed by the comparisons below
    out.w = out.x = out.y = out.z = r + p + y; out.fired = true;
    return out;
  }
  const bool cr = m_abs(r) >= d85r, cp = m_abs(p) >= d85p, cy = m_abs(y) >= d175;
  const bool inside = m_abs(r) <= d85r && m_abs(p) <= d85p && m_abs(y) <= d175;
  if (!(cr || cp || cy) || inside) return out;
  T lr = r, lp = p, ly = y;
  if (cy) ly = m_sign(y) * d175;
  else if (cp) lp = m_sign(p) * d85p;
  else lr = m_sign(r) * d85r;
  Quat<T> lim = euler2quat_cold(lr, lp, ly);
  out.w = lim.w; out.x = lim.x; out.y = lim.y; out.z = lim.z; out.fired = true;
  return out;
}
// out-of-line copy for the fp32 fast path, where the clamp is a cold branch behind a transcendental-free pre-test
// (by-value in/out so the caller's drone state never has to live in local memory)
template <typename T, typename P>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
ClampResult<T> attitude_clamp_cold(const P& p, T q0, T q1, T q2, T q3) { return attitude_clamp(p, q0, q1, q2, q3); }

// dynamics/quadrotor.py:126-144  Drone.step(u_cmd): Euler step with frozen RHS df(state, u_prev) (:80-113),
// attitude_limit (:146-168), then u_prev <- u_limit(u_cmd).  Returns true if the attitude clamp fired.
template <typename T, typename P> QS_HD bool drone_step(const P& p, Drone<T>& d, T c0, T c1, T c2, T c3) {
  const T dt = p.dt;
  // --- df: quadrotor.py:80-113 ---
  T ss = d.q0 * d.q0 + d.q1 * d.q1 + d.q2 * d.q2 + d.q3 * d.q3;
  T n1, n2;
  if (kFast<T>) { T inv = m_rsqrt(ss); n1 = d.q1 * inv; n2 = d.q2 * inv; }
  else { T n = m_sqrt(ss); n1 = d.q1 / n; n2 = d.q2 / n; }
  T w2 = T(2) * d.q0;
  T r20 = (T(2) * n2) * n2 - w2 * n2, r21 = (T(2) * n1) * n1 + w2 * n1;
  const T inv_m = p.inv_m;
  T ax = inv_m * (r20 * d.u0), ay = inv_m * (r21 * d.u0), az = inv_m * (d.u0 - p.mg);
  T e2 = T(2) * (T(1) - ss);  // K_quat * e_quat, quadrotor.py:96-97
  T sk0 = -d.wx * d.q1 - d.wy * d.q2 - d.wz * d.q3;  // q_sk @ q, quadrotor.py:98-103
  T sk1 = d.wx * d.q0 - d.wy * d.q2 + d.wz * d.q3;
  T sk2 = d.wy * d.q0 + d.wz * d.q1 - d.wx * d.q3;
  T sk3 = d.wz * d.q0 - d.wy * d.q1 + d.wx * d.q2;
  T dq0 = T(-0.5) * sk0 + e2 * d.q0, dq1 = T(-0.5) * sk1 + e2 * d.q1;
  T dq2 = T(-0.5) * sk2 + e2 * d.q2, dq3 = T(-0.5) * sk3 + e2 * d.q3;
  // inv(I) @ (I@u[1:] - w x (I w)), quadrotor.py:83,105
  T Iw0 = p.I0 * d.wx, Iw1 = p.I1 * d.wy, Iw2 = p.I2 * d.wz;
  T aw0 = p.inv_I0 * (p.I0 * d.u1 - (d.wy * Iw2 - d.wz * Iw1));
  T aw1 = p.inv_I1 * (p.I1 * d.u2 - (d.wz * Iw0 - d.wx * Iw2));
  T aw2 = p.inv_I2 * (p.I2 * d.u3 - (d.wx * Iw1 - d.wy * Iw0));
  // --- y = s + dt * f (Q1) ---
  d.px += dt * d.vx; d.py += dt * d.vy; d.pz += dt * d.vz;
  d.vx += dt * ax; d.vy += dt * ay; d.vz += dt * az;
  d.q0 += dt * dq0; d.q1 += dt * dq1; d.q2 += dt * dq2; d.q3 += dt * dq3;
  d.wx += dt * aw0; d.wy += dt * aw1; d.wz += dt * aw2;
  // --- attitude_limit: quadrotor.py:146-168 ---
  bool clamp = false;
  if (kFast<T>) {
    // |phi|>=85deg <=> |r12|>=sin85 ; |theta|>=85deg <=> not(r22>0 && |r02|<r22*tan85) ;
    // |psi|>=175deg <=> r11<0 && |r10|<=|r11|*tan5  -- no transcendental unless the clamp fires
    T w = d.q0, x = d.q1, y = d.q2, z = d.q3;
    T r10 = T(2) * (x * y - w * z), r11 = w * w - x * x + y * y - z * z, r12 = T(2) * (w * x + y * z);
    T r02 = T(2) * (x * z - w * y), r22 = w * w - x * x - y * y + z * z;
    const T s85 = p.s_roll, t85 = p.t_pitch, t5 = p.t_yaw;
    bool in_phi = m_abs(r12) < s85;
    bool in_theta = (r22 > T(0)) && (m_abs(r02) < r22 * t85);
    bool in_psi = !((r11 < T(0)) && (m_abs(r10) <= -r11 * t5));
    clamp = !(in_phi && in_theta && in_psi);
    if (!(r12 < T(1)) || r12 < T(-1)) clamp = true;  // phi = +-90deg branch: theta forced to 0
  } else {
    clamp = true;
  }
  if (clamp) {
    ClampResult<T> cr = kFast<T> ? attitude_clamp_cold(p, d.q0, d.q1, d.q2, d.q3) : attitude_clamp(p, d.q0, d.q1, d.q2, d.q3);
    clamp = cr.fired;
    if (cr.fired) {
      d.q0 = cr.w; d.q1 = cr.x; d.q2 = cr.y; d.q3 = cr.z;
      d.wx = T(0); d.wy = T(0); d.wz = T(0);
    }
  }
  // --- u_prev <- u_limit(u_cmd) (Q2) ---
  u_limit(p, c0, c1, c2, c3, d.u0, d.u1, d.u2, d.u3);
  return clamp;
}
template <typename T> QS_HD bool drone_step(Drone<T>& d, T c0, T c1, T c2, T c3) { return drone_step(DefaultParams<T>{}, d, c0, c1, c2, c3); }

// Shared tail of hover_controller (PIDController.py:87-104) and vel_controller (:119-134) followed by
// attitude_controller (:52-74): acc_des -> u = [F, M].  des_q is read (psi_des) and OVERWRITTEN (Q4).
template <typename T, typename P> QS_HD void acc2cmd(const P& p, T ax, T ay, T az, Quat<T>& dq, const Drone<T>& now, T& F, T& M0, T& M1, T& M2) {
  F = p.mass_x_g + p.mass * az;
  T phi_d, theta_d, psi_d, sp, cp;
  {
    // psi_des = quat2euler(des_q)[2]
    T w = dq.w, x = dq.x, y = dq.y, z = dq.z;
    T r10 = T(2) * (x * y - w * z), r11 = w * w - x * x + y * y - z * z;
    psi_d = m_atan2(-r10, r11);
    T h2 = r10 * r10 + r11 * r11;
    if (kSpec<T>) {
      T inv = spec_keep(m_rsqrt(h2), h2 > T(1e-100) && h2 < T(1e100));
      sp = -r10 * inv; cp = r11 * inv;
    } else if (kFast<T> && !is_f32<T>::value && h2 > T(1e-100) && h2 < T(1e100)) {
      T inv = m_rsqrt(h2);  // sin/cos of atan2(-r10, r11) from its arguments
      sp = -r10 * inv; cp = r11 * inv;
    } else {
      m_sincos(psi_d, &sp, &cp);
    }
  }
  if (kFast<T> && !is_f32<T>::value) {
    phi_d = (ax * sp - ay * cp) * p.inv_g;
    theta_d = (ax * cp + ay * sp) * p.inv_g;
  } else {
    phi_d = (ax * sp - ay * cp) / p.g;
    theta_d = (ax * cp + ay * sp) / p.g;
  }
  dq = euler2quat(phi_d, theta_d, psi_d);
  // attitude_controller re-extracts the desired angles from the quaternion it was just given
  Euler<T> ed;
  if (kSpec<T>) {
    ed.phi = spec_keep(phi_d, m_abs(phi_d) < T(1.5) && m_abs(theta_d) < T(3.1)); ed.theta = theta_d; ed.psi = psi_d;
  } else if (kFast<T> && m_abs(phi_d) < T(1.5) && m_abs(theta_d) < T(3.1)) {
    ed.phi = phi_d; ed.theta = theta_d; ed.psi = psi_d;  // quat2euler(euler2quat(.)) == identity on this range
  } else {
    ed = quat2euler(dq.w, dq.x, dq.y, dq.z);
  }
  Euler<T> en = quat2euler(now.q0, now.q1, now.q2, now.q3);
  // des rates: des[10]=des[11]=0 (written), des[12]=0 (never set)
  M0 = p.kp_roll * (ed.phi - en.phi) + p.kd_roll * (T(0) - now.wx);
  M1 = p.kp_pitch * (ed.theta - en.theta) + p.kd_pitch * (T(0) - now.wy);
  M2 = p.kp_yaw * (ed.psi - en.psi) + p.kd_yaw * (T(0) - now.wz);
}

// controller/PIDController.py:179-185  PID(des, now) with des pos/vel given explicitly
template <typename T, typename P> QS_HD void pid(const P& p, T dpx, T dpy, T dpz, T dvx, T dvy, T dvz, Quat<T>& dq, const Drone<T>& now,
                                                 T& F, T& M0, T& M1, T& M2) {
  T ax = p.kp_x * (dpx - now.px) + p.kd_x * (dvx - now.vx);
  T ay = p.kp_y * (dpy - now.py) + p.kd_y * (dvy - now.vy);
  T az = p.kp_z * (dpz - now.pz) + p.kd_z * (dvz - now.vz);
  acc2cmd(p, ax, ay, az, dq, now, F, M0, M1, M2);
}
template <typename T> QS_HD void pid(T dpx, T dpy, T dpz, T dvx, T dvy, T dvz, Quat<T>& dq, const Drone<T>& now, T& F, T& M0, T& M1, T& M2) {
  pid(DefaultParams<T>{}, dpx, dpy, dpz, dvx, dvy, dvz, dq, now, F, M0, M1, M2);
}

// controller/PIDController.py:106-141  vel_controller(des, now, last) with last == now (Q5): the kd_v* terms vanish
template <typename T, typename P> QS_HD void vel_controller(const P& p, T dvx, T dvy, T dvz, Quat<T>& dq, const Drone<T>& now, T& F, T& M0,
                                                            T& M1, T& M2) {
  T ax = p.kp_vx * (dvx - now.vx);
  T ay = p.kp_vy * (dvy - now.vy);
  T az = p.kp_vz * (dvz - now.vz);
  acc2cmd(p, ax, ay, az, dq, now, F, M0, M1, M2);
}
template <typename T> QS_HD void vel_controller(T dvx, T dvy, T dvz, Quat<T>& dq, const Drone<T>& now, T& F, T& M0, T& M1, T& M2) {
  vel_controller(DefaultParams<T>{}, dvx, dvy, dvz, dq, now, F, M0, M1, M2);
}

// gym_docking/envs/docking_env.py:257-295 state2rel, with get_dock_port_state (quadrotor.py:213-224) folded in for
// the ports the envs use: chaser [+0.1,0,0] (docking_env.py:38), target [-0.1,0,0] (:51).
// c = chaser (drone A), t = target (drone B).  rel = [dp(3), dv(3), euler(3), euler rates(3)].
template <typename T, typename P> QS_HD void state2rel(const P& p_, const Drone<T>& c, const Drone<T>& t, T rel[12]) {
  Rot<T> A = quat2rot(c.q0, c.q1, c.q2, c.q3);  // R_I2A
  Rot<T> B = quat2rot(t.q0, t.q1, t.q2, t.q3);  // R_I2B
  // dock ports: r = R^T @ [d,0,0] = d * (1, R01, R02); pos = p + r; vel = v + w x r
  const T dc = p_.port, dt_ = -p_.port;
  T rc0 = dc, rc1 = A.r01 * dc, rc2 = A.r02 * dc;
  T rt0 = dt_, rt1 = B.r01 * dt_, rt2 = B.r02 * dt_;
  if (kFast<T>) {
    // difference the two ~50 m positions first, then add the +-0.1 m lever arms
    rel[0] = (t.px - c.px) + (rt0 - rc0);
    rel[1] = (t.py - c.py) + (rt1 - rc1);
    rel[2] = (t.pz - c.pz) + (rt2 - rc2);
  } else {
    rel[0] = (t.px + rt0) - (c.px + rc0);
    rel[1] = (t.py + rt1) - (c.py + rc1);
    rel[2] = (t.pz + rt2) - (c.pz + rc2);
  }
  T vc0 = c.vx + (-c.wz * rc1 + c.wy * rc2), vc1 = c.vy + (c.wz * rc0 - c.wx * rc2), vc2 = c.vz + (-c.wy * rc0 + c.wx * rc1);
  T vt0 = t.vx + (-t.wz * rt1 + t.wy * rt2), vt1 = t.vy + (t.wz * rt0 - t.wx * rt2), vt2 = t.vz + (-t.wy * rt0 + t.wx * rt1);
  rel[3] = vt0 - vc0; rel[4] = vt1 - vc1; rel[5] = vt2 - vc2;
  // R_A2B = R_I2B @ R_I2A^T  (unit diagonals)
  T ab00 = T(1) + B.r01 * A.r01 + B.r02 * A.r02;
  T ab01 = A.r10 + B.r01 + B.r02 * A.r12;
  T ab02 = A.r20 + B.r01 * A.r21 + B.r02;
  T ab10 = B.r10 + A.r01 + B.r12 * A.r02;
  T ab11 = B.r10 * A.r10 + T(1) + B.r12 * A.r12;
  T ab12 = B.r10 * A.r20 + A.r21 + B.r12;
  T ab20 = B.r20 + B.r21 * A.r01 + A.r02;
  T ab21 = B.r20 * A.r10 + B.r21 + A.r12;
  T ab22 = B.r20 * A.r20 + B.r21 * A.r21 + T(1);
  Euler<T> e = rot2euler(ab12, ab10, ab11, ab02, ab22);
  rel[6] = e.phi; rel[7] = e.theta; rel[8] = e.psi;
  // rel_A2B_inB = R_I2B @ w_B - (R_A2B @ R_I2A) @ w_A ; evaluated as R_A2B @ (R_I2A @ w_A)
  T b0 = t.wx + B.r01 * t.wy + B.r02 * t.wz;
  T b1 = B.r10 * t.wx + t.wy + B.r12 * t.wz;
  T b2 = B.r20 * t.wx + B.r21 * t.wy + t.wz;
  T a0 = c.wx + A.r01 * c.wy + A.r02 * c.wz;
  T a1 = A.r10 * c.wx + c.wy + A.r12 * c.wz;
  T a2 = A.r20 * c.wx + A.r21 * c.wy + c.wz;
  T p = b0 - (ab00 * a0 + ab01 * a1 + ab02 * a2);
  T q = b1 - (ab10 * a0 + ab11 * a1 + ab12 * a2);
  T r = b2 - (ab20 * a0 + ab21 * a1 + ab22 * a2);
  T st, ct, tphi, icphi, cphi = T(1);
  if (kSpec<T>) {
    T h = m_rsqrt(ab02 * ab02 + ab22 * ab22);
    st = -ab02 * h; ct = ab22 * h;
    icphi = spec_keep(m_rsqrt((T(1) - ab12) * (T(1) + ab12)), ab12 < T(1) && ab12 > T(-1));
    tphi = ab12 * icphi;
  } else if (kFast<T> && ab12 < T(1) && ab12 > T(-1)) {
    // theta = atan2(-ab02, ab22), phi = asin(ab12): their sin/cos/tan follow from the arguments
    T h = m_rsqrt(ab02 * ab02 + ab22 * ab22);
    st = -ab02 * h; ct = ab22 * h;
    icphi = m_rsqrt((T(1) - ab12) * (T(1) + ab12));
    tphi = ab12 * icphi;
    if (!(ab02 * ab02 + ab22 * ab22 > T(0))) { st = T(0); ct = T(1); }  // atan2(0,0) = 0
  } else {
    m_sincos(e.theta, &st, &ct);
    tphi = m_tan(e.phi);
    T sphi;
    m_sincos(e.phi, &sphi, &cphi);
    icphi = T(0);
  }
  T rc_ps = r * ct - p * st;
  rel[9] = p * ct + r * st;
  rel[10] = q - tphi * rc_ps;
  if (kSpec<T>) rel[11] = rc_ps * icphi;
  else rel[11] = (icphi != T(0)) ? rc_ps * icphi : rc_ps / cphi;
}
template <typename T> QS_HD void state2rel(const Drone<T>& c, const Drone<T>& t, T rel[12]) { state2rel(DefaultParams<T>{}, c, t, rel); }

// ---- Philox4x32-10 (Salmon et al., SC'11).  Not in the reference: the north star's counter-based per-env RNG
// for the randomised auto-reset (docking_env.py:34 has the distribution commented out). ----
struct U4 { uint32_t x, y, z, w; };
QS_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
    U4 n;
    n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0; n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1; n.w = (uint32_t)p0;
    c = n;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c;
}

// exact fp32 ops (no FMA contraction) so host oracle, fp32 kernel and fp64 kernel all get the same bits
QS_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
QS_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
// jitter_i = ((x_i >> 8) * 2^-24 - 0.5) * scale, all in fp32
QS_HD void philox_jitter3(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t stream, float scale, float j[3]) {
  U4 c = {(uint32_t)env_id, (uint32_t)(env_id >> 32), episode, stream};
  U4 x = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float s = 5.9604644775390625e-08f;
  j[0] = f_mul(f_add(f_mul((float)(x.x >> 8), s), -0.5f), scale);
  j[1] = f_mul(f_add(f_mul((float)(x.y >> 8), s), -0.5f), scale);
  j[2] = f_mul(f_add(f_mul((float)(x.z >> 8), s), -0.5f), scale);
}

// =========================================================================================================
// Environment level: one env's step / (re)seat in registers.  The CUDA kernels (qs_lib.cu) are load -> these -> store.
// =========================================================================================================
constexpr int kRandNone = 0, kRandEpisode = 1, kRandOnce = 2;       // == qs_randomize
constexpr int kInfoDone = 1, kInfoOverlimit = 2, kInfoDocking = 4, kInfoOvertime = 8;  // == QS_INFO_*

struct RngCfg { uint64_t seed, env_id0; int randomize; float jitter_scale; };

template <typename T> struct DockEnv {
  Drone<T> c, t;        // chaser (drone A), target (drone B)
  Quat<T> dq;           // target_state_des[6:10] (Q4)
  T last_shaping;       // also the running episode return: reward_k = shaping_k - shaping_{k-1} and shaping_0 = 0
                        // (docking_env.py:221-222,243), so sum(reward) telescopes to the latest shaping
  int32_t tt, episode;  // env.t ; Philox counter word 2
};
template <typename T> struct HoverEnv {
  Drone<T> d;
  T ep_return;
  int32_t tt, episode;
};
template <typename T> struct StepOut {
  T reward;
  bool done, docking, overlimit, overtime, nonfinite;
  QS_HD int info() const { return (done ? kInfoDone : 0) | (overlimit ? kInfoOverlimit : 0) | (docking ? kInfoDocking : 0) | (overtime ? kInfoOvertime : 0); }
};

template <typename T> QS_HD Drone<T> drone_at(T x, T y, T z) {
  return Drone<T>{x, y, z, T(0), T(0), T(0), T(1), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)};
}

// env.reset(): Drone.reset x2 (quadrotor.py:65-78: state <- ini, u <- 0) and docking_env.py:233-244 (t, shaping <- 0).
// target_state_des (dq) is NOT re-initialised (Q4).  Chaser ini [8,-50,5] + Philox jitter U(-hw,hw)^3 (the distribution
// commented out at docking_env.py:34), target ini [10,-50,5].  obs = state2rel of the seated states, which for identity
// attitudes and zero rates is (dp, 0, ..., 0).
template <typename T, typename P> QS_HD void dock_env_seat(const P& p, DockEnv<T>& e, const RngCfg& rng, int64_t i, T obs[12]) {
  float j[3] = {0.f, 0.f, 0.f};
  if (rng.randomize != kRandNone)
    philox_jitter3(rng.seed, rng.env_id0 + (uint64_t)i, rng.randomize == kRandOnce ? 0u : (uint32_t)e.episode, 0u, rng.jitter_scale, j);
  // the fp32 subtractions of the origin below are exact (the result has fewer significant bits than the minuend)
  e.c = drone_at<T>(T(f_add(p.cini_x, j[0])) - p.ox, T(f_add(p.cini_y, j[1])) - p.oy, T(f_add(p.cini_z, j[2])) - p.oz);
  e.t = drone_at<T>(p.tini_x, p.tini_y, p.tini_z);
  e.tt = 0; e.last_shaping = T(0);
  if (kFast<T>) obs[0] = (e.t.px - e.c.px) + p.port2;
  else obs[0] = (e.t.px + (-p.port)) - (e.c.px + p.port);
  obs[1] = e.t.py - e.c.py; obs[2] = e.t.pz - e.c.pz;
#pragma unroll
  for (int k = 3; k < 12; k++) obs[k] = T(0);
}
template <typename T> QS_HD void dock_env_seat(DockEnv<T>& e, const RngCfg& rng, int64_t i, T obs[12]) { dock_env_seat(DefaultParams<T>{}, e, rng, i, obs); }

// gym_docking/envs/docking_env.py:104-231 (KIND = kDockV0) / moving_docking_env.py:111-192 (KIND = kDockV2).
// rel = the step's observation (terminal observation if done); the caller applies the auto-reset.
template <typename T, int KIND, typename P> QS_HD void dock_env_step(const P& p, DockEnv<T>& e, const T act[4], T rel[12], StepOut<T>& o) {
  using C = K<T>;
  e.tt += 1;  // docking_env.py:108
  // chaser command: rotor2control @ (action_std * a + action_mean), action_std = action_mean = m*g/2 (docking_env.py:98-99,115)
  const T half = p.half;
  T uc0, uc1, uc2, uc3;
  rotor2control(p, half * act[0] + half, half * act[1] + half, half * act[2] + half, half * act[3] + half, uc0, uc1, uc2, uc3);
  // target command from the target state BEFORE stepping (docking_env.py:119 / moving_docking_env.py:126)
  T ut0, ut1, ut2, ut3;
  if (KIND == kDockV0) pid(p, p.des_px, p.des_py, p.des_pz, T(0), T(0), T(0), e.dq, e.t, ut0, ut1, ut2, ut3);  // des: docking_env.py:60-64
  else vel_controller(p, p.des_vx, p.des_vy, p.des_vz, e.dq, e.t, ut0, ut1, ut2, ut3);       // des vel: moving_docking_env.py:62-66
  drone_step(p, e.t, ut0, ut1, ut2, ut3);  // docking_env.py:120
  drone_step(p, e.c, uc0, uc1, uc2, uc3);  // docking_env.py:121
  state2rel(p, e.c, e.t, rel);             // docking_env.py:124-127
  // flags: docking_env.py:130-152 / moving_docking_env.py:137-157
  const T radius = p.radius;
  T np_ = norm3(rel[0], rel[1], rel[2]), nv = norm3(rel[3], rel[4], rel[5]);
  o.docking = (np_ < p.tol_pos) && (nv < p.tol_vel) && (m_abs(rel[6]) < p.d10) && (m_abs(rel[7]) < p.d10) && (m_abs(rel[8]) < p.d10);
  o.overlimit = (np_ >= radius) || (e.c.pz <= p.z_floor);
  o.overtime = e.tt >= p.max_steps;
  o.done = o.overlimit || o.overtime;
  // shaping / reward: docking_env.py:215-222 / moving_docking_env.py:176-183
  T na = m_sqrt(act[0] * act[0] + act[1] * act[1] + act[2] * act[2] + act[3] * act[3]);
  T shaping;
  if (kFast<T>) {
    shaping = -p.ten_over_radius * np_ - nv - T(10.0 / 3.141592653589793) * norm3(rel[6], rel[7], rel[8]) -
              norm3(rel[9], rel[10], rel[11]) - T(0.1) * na + (o.docking ? T(1) : T(0));
  } else {
    shaping = T(-10) * norm3(rel[0] / radius, rel[1] / radius, rel[2] / radius) - T(1) * nv -
              T(10) * norm3(rel[6] / C::pi, rel[7] / C::pi, rel[8] / C::pi) - T(1) * norm3(rel[9], rel[10], rel[11]) -
              T(0.1) * na + T(1) * (o.docking ? T(1) : T(0));
  }
  o.reward = shaping - e.last_shaping;
  e.last_shaping = shaping;
  T fsum = m_abs(o.reward);
#pragma unroll
  for (int k = 0; k < 12; k++) fsum += m_abs(rel[k]);
  o.nonfinite = !m_finite(fsum);
}
template <typename T, int KIND> QS_HD void dock_env_step(DockEnv<T>& e, const T act[4], T rel[12], StepOut<T>& o) {
  dock_env_step<T, KIND>(DefaultParams<T, KIND>{}, e, act, rel, o);
}

// Speculative fp64 step (see sdouble): returns false if the step has to be redone from the ORIGINAL state with the
// guarded arithmetic -- e, rel and o are then meaningless.  Poison can only be born in the elementary functions and the
// shortcut preconditions; from there it reaches either the observation / reward (covered by o.nonfinite) or, through the
// target's controller, the target's stored command and desired quaternion, which this step's observation does not see
// (one step of command latency, Q2) and which are therefore checked explicitly.
template <typename T> QS_HD Drone<sdouble> to_spec(const Drone<T>& d) {
  return Drone<sdouble>{sdouble(d.px), sdouble(d.py), sdouble(d.pz), sdouble(d.vx), sdouble(d.vy), sdouble(d.vz), sdouble(d.q0),
                        sdouble(d.q1), sdouble(d.q2), sdouble(d.q3), sdouble(d.wx), sdouble(d.wy), sdouble(d.wz), sdouble(d.u0),
                        sdouble(d.u1), sdouble(d.u2), sdouble(d.u3)};
}
QS_HD Drone<double> from_spec(const Drone<sdouble>& d) {
  return Drone<double>{d.px.v, d.py.v, d.pz.v, d.vx.v, d.vy.v, d.vz.v, d.q0.v, d.q1.v, d.q2.v, d.q3.v, d.wx.v, d.wy.v, d.wz.v,
                       d.u0.v, d.u1.v, d.u2.v, d.u3.v};
}
// Speculative evaluation reads its constants from the same provider: the compiled-in ones folded for sdouble, or the very
// same runtime block viewed as sdouble members (sdouble is a struct of one double: identical layout).
template <int KIND, typename P> QS_HD bool dock_env_step_spec(const P& p, DockEnv<double>& e, const double act[4], double rel[12], StepOut<double>& o) {
  DockEnv<sdouble> s;
  s.c = to_spec(e.c); s.t = to_spec(e.t);
  s.dq = Quat<sdouble>{sdouble(e.dq.w), sdouble(e.dq.x), sdouble(e.dq.y), sdouble(e.dq.z)};
  s.last_shaping = sdouble(e.last_shaping); s.tt = e.tt; s.episode = e.episode;
  const sdouble a[4] = {sdouble(act[0]), sdouble(act[1]), sdouble(act[2]), sdouble(act[3])};
  sdouble r[12];
  StepOut<sdouble> so;
  if constexpr (P::custom) dock_env_step<sdouble, KIND>(reinterpret_cast<const RuntimeParams<sdouble>&>(p), s, a, r, so);
  else dock_env_step<sdouble, KIND>(DefaultParams<sdouble, KIND>{}, s, a, r, so);
  const double hidden = ((s.t.u0.v + s.t.u1.v) + (s.t.u2.v + s.t.u3.v)) + ((s.dq.w.v + s.dq.x.v) + (s.dq.y.v + s.dq.z.v));
  const bool ok = !so.nonfinite && (hidden - hidden == 0.0);
  e.c = from_spec(s.c); e.t = from_spec(s.t);
  e.dq = Quat<double>{s.dq.w.v, s.dq.x.v, s.dq.y.v, s.dq.z.v};
  e.last_shaping = s.last_shaping.v; e.tt = s.tt;
#pragma unroll
  for (int k = 0; k < 12; k++) rel[k] = r[k].v;
  o.reward = so.reward.v; o.done = so.done; o.docking = so.docking; o.overlimit = so.overlimit; o.overtime = so.overtime;
  o.nonfinite = so.nonfinite;
  return ok;
}
template <int KIND> QS_HD bool dock_env_step_spec(DockEnv<double>& e, const double act[4], double rel[12], StepOut<double>& o) {
  return dock_env_step_spec<KIND>(DefaultParams<double, KIND>{}, e, act, rel, o);
}

// =========================================================================================================
// The PID "expert" that flies the CHASER in the reference's imitation-learning scripts (run_expert_policy.py:44-69,
// run_expert_record.py:40-139):
//     des_vel = kp * (target_pos + [-0.2, 0, 0] - chaser_pos) + kd * (-chaser_vel),  kp = 0.35, kd = 0
//     if not first step of the episode: state_des[3:6] = des_vel          (the first step keeps the stale value)
//     u = controller.vel_controller(state_des, chaser_state, state_last)  (kd_v* = 0: state_last is never felt)
//     action = (inv(rotor2control) @ u - action_mean) / action_std
// ExpertMem = the slice of state_des the controller reads and rewrites: des vel [3:6] and des quaternion [6:10] (Q4).
// `alias`: the scripts bind state_des = env.chaser_ini_state, and env.reset() -> Drone.reset aliases that same array as
// the chaser's state (quadrotor.py:75).  So (i) every episode after the first starts from the MUTATED initial state
// (velocity = last des_vel, attitude = last des quaternion), and (ii) on the first step of an episode des, now and last
// are one array: the velocity error is exactly 0, the controller levels the chaser's own quaternion (keeping its yaw)
// and the attitude errors vanish, so u = [m g, 0, 0, 0].  alias = true reproduces that literally (`reseated` tells the
// caller the chaser state / reset observation changed); alias = false treats state_des as a private copy.
// =========================================================================================================
template <typename T> struct ExpertMem { T dvx, dvy, dvz; Quat<T> dq; };

template <typename T> QS_HD bool pid_expert_reseat(Drone<T>& c, const ExpertMem<T>& m, bool first, bool alias) {
  bool reseated = false;
  if (first && alias) {
    // env.reset() seated the chaser on chaser_ini_state == state_des: position untouched, [3:6] = des vel, [6:10] = des q
    c.vx = m.dvx; c.vy = m.dvy; c.vz = m.dvz;
    c.q0 = m.dq.w; c.q1 = m.dq.x; c.q2 = m.dq.y; c.q3 = m.dq.z;
    reseated = true;
  }
  return reseated;
}

// second half of the aliased first step, after the caller has taken the reset observation from the re-seated chaser
template <typename T, typename P> QS_HD void pid_expert_cmd(const P& p, Drone<T>& c, const Drone<T>& t, ExpertMem<T>& m, bool first, bool alias, T act[4]) {
  T F, M0, M1, M2;
  if (first && alias) {
    // vel_controller(des, now, last) on ONE array: e_vel == 0 -> acc_des = 0 -> att_des = (0, 0, psi_des); the write of
    // state_des[6:10] / [10] / [11] lands in the chaser's own state; attitude_controller then sees des == now.
    T r10 = T(2) * (m.dq.x * m.dq.y - m.dq.w * m.dq.z), r11 = m.dq.w * m.dq.w - m.dq.x * m.dq.x + m.dq.y * m.dq.y - m.dq.z * m.dq.z;
    m.dq = euler2quat(T(0), T(0), m_atan2(-r10, r11));
    c.q0 = m.dq.w; c.q1 = m.dq.x; c.q2 = m.dq.y; c.q3 = m.dq.z;
    c.wx = T(0); c.wy = T(0);
    F = p.mass_x_g + p.mass * T(0); M0 = T(0); M1 = T(0); M2 = T(0);
  } else {
    if (!first) {
      const T kp = T(0.35), kd = T(0);
      m.dvx = kp * ((t.px + T(-0.2)) - c.px) + kd * (-c.vx);
      m.dvy = kp * ((t.py + T(0)) - c.py) + kd * (-c.vy);
      m.dvz = kp * ((t.pz + T(0)) - c.pz) + kd * (-c.vz);
    }
    vel_controller(p, m.dvx, m.dvy, m.dvz, m.dq, c, F, M0, M1, M2);
  }
  // inv(rotor2control) (quadrotor.py:56-59) in closed form, then the env's action normalisation (docking_env.py:98-99)
  const T hl = p.hl, ql = P::custom ? T(0.25) / p.lam : T(0.25 / (1.5e-9 / 6.11e-8)), half = p.half;
  const T f0 = (T(0.25) * F - hl * M1) + ql * M2, f1 = (T(0.25) * F + hl * M0) - ql * M2;
  const T f2 = (T(0.25) * F + hl * M1) + ql * M2, f3 = (T(0.25) * F - hl * M0) - ql * M2;
  act[0] = (f0 - half) / half; act[1] = (f1 - half) / half; act[2] = (f2 - half) / half; act[3] = (f3 - half) / half;
}
template <typename T> QS_HD void pid_expert_cmd(Drone<T>& c, const Drone<T>& t, ExpertMem<T>& m, bool first, bool alias, T act[4]) {
  pid_expert_cmd(DefaultParams<T>{}, c, t, m, first, alias, act);
}

template <typename T> QS_HD void drone_to_obs13(const Drone<T>& d, T o[13]) {
  o[0] = d.px; o[1] = d.py; o[2] = d.pz; o[3] = d.vx; o[4] = d.vy; o[5] = d.vz;
  o[6] = d.q0; o[7] = d.q1; o[8] = d.q2; o[9] = d.q3; o[10] = d.wx; o[11] = d.wy; o[12] = d.wz;
}

// hover (re)seat: hovering_env.py:23-29, :80-82 -- pos [0,0,5] + U(-hw,hw)^3, euler U(-0.2,0.2)^3, zero rates, u <- 0
template <typename T, typename P> QS_HD void hover_env_seat(const P& p, HoverEnv<T>& e, const RngCfg& rng, int64_t i, T obs[13]) {
  float jp[3] = {0.f, 0.f, 0.f}, je[3] = {0.f, 0.f, 0.f};
  if (rng.randomize != kRandNone) {
    uint32_t ep = rng.randomize == kRandOnce ? 0u : (uint32_t)e.episode;
    philox_jitter3(rng.seed, rng.env_id0 + (uint64_t)i, ep, 0u, rng.jitter_scale, jp);
    philox_jitter3(rng.seed, rng.env_id0 + (uint64_t)i, ep, 1u, 0.4f, je);
  }
  Quat<T> q = euler2quat(T(je[0]), T(je[1]), T(je[2]));
  e.d = drone_at<T>(T(f_add(p.cini_x, jp[0])), T(f_add(p.cini_y, jp[1])), T(f_add(p.cini_z, jp[2])));
  e.d.q0 = q.w; e.d.q1 = q.x; e.d.q2 = q.y; e.d.q3 = q.z;
  e.tt = 0; e.ep_return = T(0);
  drone_to_obs13(e.d, obs);
}
template <typename T> QS_HD void hover_env_seat(HoverEnv<T>& e, const RngCfg& rng, int64_t i, T obs[13]) { hover_env_seat(DefaultParams<T, kHover>{}, e, rng, i, obs); }

// hovering-v0 step: gym_docking/envs/hovering_env.py:47-78
template <typename T, typename P> QS_HD void hover_env_step(const P& p, HoverEnv<T>& e, const T act[4], T obs[13], StepOut<T>& o) {
  Drone<T>& d = e.d;
  e.tt += 1;
  const T amax = p.mg;  // action_max = m*g per rotor, hovering_env.py:42
  T u0, u1, u2, u3;
  rotor2control(p, amax * act[0], amax * act[1], amax * act[2], amax * act[3], u0, u1, u2, u3);  // :51
  drone_step(p, d, u0, u1, u2, u3);                                                               // :52
  // errors w.r.t. state_des = [0,0,5, 0.., q=identity, 0..] (hovering_env.py:31-35, :57-60); quat2euler(identity) = 0
  Euler<T> en = quat2euler(d.q0, d.q1, d.q2, d.q3);
  T pe = norm3(T(0) - d.px, T(0) - d.py, T(5) - d.pz), ve = norm3(d.vx, d.vy, d.vz);
  T ae = norm3(en.phi, en.theta, en.psi), we = norm3(d.wx, d.wy, d.wz);
  T r_thre = (pe < T(0.1) && ve < T(0.1)) ? T(1) : T(0);                                      // :62-66
  o.done = (norm3(d.px, d.py, d.pz) > p.radius) || (ve > p.radius);                            // :68
  o.overlimit = o.done; o.overtime = false; o.docking = r_thre > T(0);
  o.reward = o.done ? T(-0.1) : (r_thre + T(0.1) - T(0.01) * pe - T(0.001) * ve - T(0.01) * ae - T(0.001) * we);  // :70-76
  e.ep_return += o.reward;
  drone_to_obs13(d, obs);
  T fsum = m_abs(o.reward);
#pragma unroll
  for (int k = 0; k < 13; k++) fsum += m_abs(obs[k]);
  o.nonfinite = !m_finite(fsum);
}
template <typename T> QS_HD void hover_env_step(HoverEnv<T>& e, const T act[4], T obs[13], StepOut<T>& o) { hover_env_step(DefaultParams<T, kHover>{}, e, act, obs, o); }

}  // namespace qs

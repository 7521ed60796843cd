// wdb_sa_physics.cuh -- one-step physics of the discrete-action single-agent envs as
// out-of-line device functions, shared by the stand-alone step kernels (wdb_small_envs.cu,
// wdb_classic_control.cu) and the whole-rollout kernel (wdb_sa_rollout.cu): ONE compiled body
// per env, so the fused path produces the same bits as the step kernels that are pinned to
// the reference's numba binaries (tests/test_gpu_classic_control.py).
#pragma once
#include <math.h>

#include "wdb_common.cuh"

namespace wdb {

// numba's `_clip(v, lo, hi)` (mountain_car_step_numba.py:5-11): two ordered tests.
static __device__ __forceinline__ double sa_clip_f64(double v, double lo, double hi) {
  if (v < lo) return lo;
  if (v > hi) return hi;
  return v;
}

// ========================================================================= CartPole
// NumbaClassicControlCartPoleEnvStep, cartpole_step_numba.py:42-75.  float32 everywhere
// except where the float64 literal 4.0/3.0 promotes (thetaacc, xacc and the two velocity
// updates).  Fused multiply-adds exactly where the reference binary has them (ptxas on
// numba's PTX, oracle/_ref/numba_cartpole.cubin): FFMA for the force sum, g*sin - cos*temp
// and the two position updates; DFMA for the two velocity updates; plain mul/div elsewhere.
static __device__ __noinline__ float4 cartpole_physics(float4 s, int action, float gravity,
                                                       float masspole, float total_mass,
                                                       float length, float polemass_length,
                                                       float force_mag, float tau) {
  const float x = s.x, x_dot = s.y, theta = s.z, theta_dot = s.w;
  const float force = (action > 0.5f) ? force_mag : -force_mag;
  const float costheta = cosf(theta), sintheta = sinf(theta);
  const float temp = __fdiv_rn(
      __fmaf_rn(__fmul_rn(polemass_length, __fmul_rn(theta_dot, theta_dot)), sintheta, force),
      total_mass);
  const float c2m = __fdiv_rn(__fmul_rn(masspole, __fmul_rn(costheta, costheta)), total_mass);
  const float torque = __fmaf_rn(gravity, sintheta, -__fmul_rn(costheta, temp));
  const double thetaacc =
      __ddiv_rn((double)torque, __dmul_rn((double)length, __dsub_rn(4.0 / 3.0, (double)c2m)));
  const double xacc = __dsub_rn(
      (double)temp,
      __ddiv_rn(__dmul_rn(__dmul_rn((double)polemass_length, thetaacc), (double)costheta),
                (double)total_mass));
  float4 n;
  n.x = __fmaf_rn(tau, x_dot, x);
  n.y = (float)__fma_rn((double)tau, xacc, (double)x_dot);
  n.z = __fmaf_rn(tau, theta_dot, theta);
  n.w = (float)__fma_rn((double)tau, thetaacc, (double)theta_dot);
  return n;
}
static __device__ __forceinline__ bool cartpole_terminated(float4 n, float theta_thr, float x_thr) {
  return (n.x < -x_thr) || (n.x > x_thr) || (n.z < -theta_thr) || (n.z > theta_thr);
}

// ===================================================================== MountainCar
// NumbaClassicControlMountainCarEnvStep, mountain_car_step_numba.py:14-70.
//   velocity += (action - 1) * force + cos(3 * position) * (-gravity)   -- all float64
//   (reference SASS: DMUL g*cos; DFMA force*(action-1) - that; DADD velocity)
//   terminated is decided on the float64 values BEFORE they are rounded into state.
static __device__ __noinline__ float2 mountain_car_physics(float2 s, int action,
                                                           float min_position,
                                                           float max_position, float max_speed,
                                                           float goal_position,
                                                           float goal_velocity, float force,
                                                           float gravity, int *terminated) {
  const double pos0 = (double)s.x;
  const double c = cos(__dmul_rn(pos0, 3.0));
  const double push = __fma_rn((double)force, (double)(long long)(action - 1),
                               -__dmul_rn((double)gravity, c));
  double vel = __dadd_rn(push, (double)s.y);
  vel = sa_clip_f64(vel, (double)(-max_speed), (double)max_speed);
  double pos = __dadd_rn(pos0, vel);
  pos = sa_clip_f64(pos, (double)min_position, (double)max_position);
  if (pos == (double)min_position && vel < 0.0) vel = 0.0;
  *terminated = (pos >= (double)goal_position && vel >= (double)goal_velocity) ? 1 : 0;
  return make_float2((float)pos, (float)vel);
}

// ========================================================================== Acrobot
// NumbaClassicControlAcrobotEnvStep, acrobot_step_numba.py:24-168: RK4 (dt = 0.2) of the
// two-link dynamics `_dsdt` (:70-109), book (not "nips") version.  All link constants are
// 1.0 / 0.5 (:8-14), so the constant sub-expressions below are the values Python's left-to-
// right float64 evaluation gives.  Types per numba: the state and every k / k_update array
// are float32 (each stage is ROUNDED to float32 on store, :116-131); cos/sin of float32
// angles are the float32 routines; everything that touches a constant is float64.
struct Vec4 { float v[4]; };

static __device__ __forceinline__ Vec4 acrobot_dsdt(const Vec4 &s, double torque) {
  constexpr double kPi = 3.141592653589793;
  const float theta1 = s.v[0], theta2 = s.v[1], dtheta1 = s.v[2], dtheta2 = s.v[3];
  const double c2 = (double)cosf(theta2);
  const double s2 = (double)sinf(theta2);
  // d1 = m1*lc1^2 + m2*(l1^2 + lc2^2 + 2*l1*lc2*cos(theta2)) + I1 + I2
  const double d1 = ((0.25 + (1.25 + c2)) + 1.0) + 1.0;
  // d2 = m2*(lc2^2 + l1*lc2*cos(theta2)) + I2
  const double d2 = (0.25 + 0.5 * c2) + 1.0;
  // phi2 = m2*lc2*g*cos(theta1 + theta2 - pi/2); theta1 + theta2 is a float32 add
  const double phi2 = (1.0 * 0.5 * 9.8) * cos((double)__fadd_rn(theta1, theta2) - kPi / 2);
  const double phi1 = ((-0.5 * (double)__fmul_rn(dtheta2, dtheta2)) * s2
                       - ((double)dtheta2 * (double)dtheta1) * s2
                       + ((1.0 * 0.5 + 1.0 * 1.0) * 9.8) * cos((double)theta1 - kPi / 2))
                      + phi2;
  const double ddtheta2 =
      (torque + d2 / d1 * phi1 - (0.5 * (double)__fmul_rn(dtheta1, dtheta1)) * s2 - phi2) /
      ((0.25 + 1.0) - d2 * d2 / d1);
  const double ddtheta1 = -(d2 * ddtheta2 + phi1) / d1;
  Vec4 d;
  d.v[0] = dtheta1;
  d.v[1] = dtheta2;
  d.v[2] = (float)ddtheta1;
  d.v[3] = (float)ddtheta2;
  return d;
}

static __device__ __forceinline__ double acrobot_wrap(double x, double m, double M) {
  const double diff = M - m;
  while (x > M) x = x - diff;
  while (x < m) x = x + diff;
  return x;
}

// one RK4 step + wrap / clip; returns the new state, writes obs[6], reward and the terminal flag
static __device__ __noinline__ float4 acrobot_physics(float4 s4, int action, float *o,
                                                      float *reward, int *terminated_out) {
  constexpr double kPi = 3.141592653589793;
  constexpr double kMaxVel1 = 12.566370614359172, kMaxVel2 = 28.274333882308138;
  constexpr double kDt = 0.2, kDt2 = 0.1;
  const double torque = (double)(action - 1);  // AVAIL_TORQUE = [-1, 0, 1] (:6)
  Vec4 s; s.v[0] = s4.x; s.v[1] = s4.y; s.v[2] = s4.z; s.v[3] = s4.w;
  // rk4 (:112-134)
  const Vec4 k1 = acrobot_dsdt(s, torque);
  Vec4 u;
#pragma unroll
  for (int i = 0; i < 4; i++) u.v[i] = (float)((double)s.v[i] + (double)k1.v[i] * kDt2);
  const Vec4 k2 = acrobot_dsdt(u, torque);
#pragma unroll
  for (int i = 0; i < 4; i++) u.v[i] = (float)((double)s.v[i] + (double)k2.v[i] * kDt2);
  const Vec4 k3 = acrobot_dsdt(u, torque);
#pragma unroll
  for (int i = 0; i < 4; i++) u.v[i] = (float)((double)s.v[i] + (double)k3.v[i] * kDt);
  const Vec4 k4 = acrobot_dsdt(u, torque);
  float ns[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const double sum = (((double)k1.v[i] + 2.0 * (double)k2.v[i]) + 2.0 * (double)k3.v[i]) +
                       (double)k4.v[i];
    ns[i] = (float)((double)s.v[i] + (kDt / 6.0) * sum);
  }
  ns[0] = (float)acrobot_wrap((double)ns[0], -kPi, kPi);
  ns[1] = (float)acrobot_wrap((double)ns[1], -kPi, kPi);
  ns[2] = (float)fmin(fmax((double)ns[2], -kMaxVel1), kMaxVel1);
  ns[3] = (float)fmin(fmax((double)ns[3], -kMaxVel2), kMaxVel2);
  // _terminal (:151-153): float32 throughout
  const float c0 = cosf(ns[0]);
  const bool terminated = __fsub_rn(-c0, cosf(__fadd_rn(ns[1], ns[0]))) > 1.0f;
  *reward = terminated ? 0.0f : -1.0f;
  *terminated_out = terminated ? 1 : 0;
  // _get_ob (:156-168)
  o[0] = c0;
  o[1] = sinf(ns[0]);
  o[2] = cosf(ns[1]);
  o[3] = sinf(ns[1]);
  o[4] = ns[2];
  o[5] = ns[3];
  return make_float4(ns[0], ns[1], ns[2], ns[3]);
}

}  // namespace wdb

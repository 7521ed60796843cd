// wdb_classic_control.cu -- MountainCar, ContinuousMountainCar, Acrobot and Pendulum steps
// (SURVEY section 8 row f2).  The reference has these only as numba kernels
// (example_envs/single_agent/classic_control/*/*_step_numba.py) launched as one 1-thread
// block per env; here one thread per env in 256-thread CTAs, vector state I/O.
//
// Arithmetic types are NOT the ones the Python source suggests at first sight: numba's type
// inference promotes `3 * position` (int64 * float32) and every expression touching a
// Python float literal to float64, keeps float32 where only float32 values meet, and picks
// the float32 or float64 libdevice routine by the argument type.  The types below were read
// off numba's own type annotations of the reference kernels and the fused-multiply-add
// placement off the SASS that ptxas makes of numba's PTX (the test tree compiles both from
// the reference sources; tests/test_gpu_classic_control.py runs those reference binaries
// next to these kernels).  Where a float64 result is rounded to float32 on store, a different FMA
// contraction can only matter in a 1e-8 fraction of the cases, so only the documented
// patterns are forced with intrinsics.
#include <math.h>

#include "wdb_common.cuh"
#include "wdb_sa_physics.cuh"

namespace wdb {
namespace {

constexpr int kCcThreads = 256;


// ===================================================================== MountainCar
// NumbaClassicControlMountainCarEnvStep, mountain_car_step_numba.py:14-70.
//   velocity += (action - 1) * force + cos(3 * position) * (-gravity)   -- all float64
//   (reference SASS: DMUL g*cos; DFMA force*(action-1) - that; DADD velocity)
//   terminated is decided on the float64 values BEFORE they are rounded into state.
//   done = 1 at the episode end, else 2 when the goal is reached (:66-70).
__global__ void __launch_bounds__(kCcThreads)
mountain_car_step_kernel(int n_envs, float2 *__restrict__ state, const int *__restrict__ action,
                         int *__restrict__ done, float *__restrict__ reward,
                         float2 *__restrict__ obs, float min_position, float max_position,
                         float max_speed, float goal_position, float goal_velocity,
                         float force, float gravity, int *__restrict__ env_timestep,
                         int episode_length) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n_envs) return;
  const int t = env_timestep[env] + 1;
  env_timestep[env] = t;
  int terminated = 0;
  const float2 n = mountain_car_physics(state[env], action[env], min_position, max_position,
                                        max_speed, goal_position, goal_velocity, force,
                                        gravity, &terminated);
  state[env] = n;
  obs[env] = n;
  reward[env] = -1.0f;
  if (t == episode_length) done[env] = 1;
  else if (terminated) done[env] = 2;
}

// ============================================================ ContinuousMountainCar
// NumbaClassicControlContinuousMountainCarEnvStep, continuous_mountain_car_step_numba.py:
// 14-71.  force = clip(action) and force * power stay float32; the rest is float64
// (reference SASS: FMUL force*power; DFMA cos * -0.0025 + that; DADD velocity).
// reward = (terminated ? 100 : 0) - pow(action, 2) * 0.1 in float64 (:64-69).
__global__ void __launch_bounds__(kCcThreads)
continuous_mountain_car_step_kernel(int n_envs, float2 *__restrict__ state,
                                    const float *__restrict__ action, int *__restrict__ done,
                                    float *__restrict__ reward, float2 *__restrict__ obs,
                                    float min_action, float max_action, float min_position,
                                    float max_position, float max_speed, float goal_position,
                                    float goal_velocity, float power,
                                    int *__restrict__ env_timestep, int episode_length) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n_envs) return;
  const int t = env_timestep[env] + 1;
  env_timestep[env] = t;
  const float a = action[env];
  const float2 s = state[env];
  float f = a;
  if (a < min_action) f = min_action;
  else if (a > max_action) f = max_action;
  const float fp = __fmul_rn(f, power);
  const double pos0 = (double)s.x;
  const double c = cos(__dmul_rn(pos0, 3.0));
  double vel = __dadd_rn(__fma_rn(c, -0.0025, (double)fp), (double)s.y);
  vel = sa_clip_f64(vel, (double)(-max_speed), (double)max_speed);
  double pos = __dadd_rn(pos0, vel);
  pos = sa_clip_f64(pos, (double)min_position, (double)max_position);
  if (pos == (double)min_position && vel < 0.0) vel = 0.0;
  const float2 n = make_float2((float)pos, (float)vel);
  state[env] = n;
  obs[env] = n;
  const bool terminated = pos >= (double)goal_position && vel >= (double)goal_velocity;
  const double rew = terminated ? 100.0 : 0.0;
  reward[env] = (float)__dsub_rn(rew, __dmul_rn(pow((double)a, 2.0), 0.1));
  if (t == episode_length || terminated) done[env] = 1;
}

// ========================================================================= Pendulum
// NumbaClassicControlPendulumEnvStep, pendulum_step_numba.py:30-72 (g = 9.81, m = l = 1,
// dt = 0.05, max_speed = 8, max_torque = 2 are module constants there, :9-14).
//   u = clip(action, -2, 2) in float64; angle_normalize(th) = ((th + pi) % 2pi) - pi in
//   float64, where NVVM lowers `%` to  a - floor(|a| / b) * b  with the sign of a restored
//   and Python's "result takes the divisor's sign" fix-up; sin(th) is the FLOAT32 routine
//   (th is float32), cos/sin(newth) the float64 ones.  FMA placement from the reference SASS:
//   costs = fma(u*u, 0.001, fma(an, an, 0.1 * (double)(thdot*thdot)))
//   newthdot = fma(fma(u, 3, 14.715 * sin(th)), dt, thdot);  newth = fma(newthdot, dt, th).
__device__ __forceinline__ double python_mod_2pi(double a) {
  const double b = 2.0 * 3.141592653589793;
  const double q = floor(__ddiv_rn(fabs(a), b));
  double r = __fma_rn(-q, b, fabs(a));
  if (!(a >= 0.0)) r = -r;
  if (r < 0.0) r = __dadd_rn(r, b);
  return r;
}

__global__ void __launch_bounds__(kCcThreads)
pendulum_step_kernel(int n_envs, float2 *__restrict__ state, const float *__restrict__ action,
                     int *__restrict__ done, float *__restrict__ reward,
                     float *__restrict__ obs /*[E,1,3]*/, int *__restrict__ env_timestep,
                     int episode_length) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n_envs) return;
  const int t = env_timestep[env] + 1;
  env_timestep[env] = t;
  constexpr double kPi = 3.141592653589793;
  constexpr double kDt = 0.05;
  constexpr double kGain = 3 * 9.81 / (2 * 1.0);  // 3 * g / (2 * l)
  const double u = sa_clip_f64((double)action[env], -2.0, 2.0);
  const float2 s = state[env];
  const double th = (double)s.x, thdot = (double)s.y;
  const double an = __dsub_rn(python_mod_2pi(__dadd_rn(th, kPi)), kPi);
  const double td2 = (double)__fmul_rn(s.y, s.y);
  const double costs = __fma_rn(__dmul_rn(u, u), 0.001, __fma_rn(an, an, __dmul_rn(td2, 0.1)));
  double newthdot = __fma_rn(__fma_rn(u, 3.0, __dmul_rn((double)sinf(s.x), kGain)), kDt, thdot);
  newthdot = sa_clip_f64(newthdot, -8.0, 8.0);
  const double newth = __fma_rn(newthdot, kDt, th);
  state[env] = make_float2((float)newth, (float)newthdot);
  float *o = obs + (size_t)env * 3;
  o[0] = (float)cos(newth);
  o[1] = (float)sin(newth);
  o[2] = (float)newthdot;
  reward[env] = (float)(-costs);
  if (t == episode_length) done[env] = 1;
}

// ========================================================================== Acrobot
// NumbaClassicControlAcrobotEnvStep, acrobot_step_numba.py:24-168: RK4 (dt = 0.2) of the
// two-link dynamics `_dsdt` (:70-109), book (not "nips") version.  All link constants are
// 1.0 / 0.5 (:8-14), so the constant sub-expressions below are the values Python's left-to-
// right float64 evaluation gives.  Types per numba: the state and every k / k_update array
// are float32 (each stage is ROUNDED to float32 on store, :116-131); cos/sin of float32
// angles are the float32 routines; everything that touches a constant is float64.
__global__ void __launch_bounds__(kCcThreads)
acrobot_step_kernel(int n_envs, float4 *__restrict__ state, const int *__restrict__ action,
                    int *__restrict__ done, float *__restrict__ reward,
                    float *__restrict__ obs /*[E,1,6]*/, int *__restrict__ env_timestep,
                    int episode_length) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n_envs) return;
  const int t = env_timestep[env] + 1;
  env_timestep[env] = t;
  float o6[6], rew;
  int terminated = 0;
  state[env] = acrobot_physics(state[env], action[env], o6, &rew, &terminated);
  reward[env] = rew;
  float *o = obs + (size_t)env * 6;
#pragma unroll
  for (int i = 0; i < 6; i++) o[i] = o6[i];
  if (t == episode_length || terminated) done[env] = 1;
}

inline bool misaligned(const void *p, uintptr_t mask) {
  return (reinterpret_cast<uintptr_t>(p) & mask) != 0;
}

}  // namespace
}  // namespace wdb

using namespace wdb;

WDB_API int wdb_mountain_car_step(void *stream, int n_envs, float *state, const int *action,
                                  int *done, float *reward, float *obs, float min_position,
                                  float max_position, float max_speed, float goal_position,
                                  float goal_velocity, float force, float gravity,
                                  int *env_timestep, int episode_length) {
  if (!state || !action || !done || !reward || !obs || !env_timestep || n_envs <= 0)
    return (int)cudaErrorInvalidValue;
  if (misaligned(state, 7) || misaligned(obs, 7)) return (int)cudaErrorMisalignedAddress;
  mountain_car_step_kernel<<<(n_envs + kCcThreads - 1) / kCcThreads, kCcThreads, 0,
                             as_stream(stream)>>>(
      n_envs, reinterpret_cast<float2 *>(state), action, done, reward,
      reinterpret_cast<float2 *>(obs), min_position, max_position, max_speed, goal_position,
      goal_velocity, force, gravity, env_timestep, episode_length);
  return finish_launch();
}

WDB_API int wdb_continuous_mountain_car_step(
    void *stream, int n_envs, float *state, const float *action, int *done, float *reward,
    float *obs, float min_action, float max_action, float min_position, float max_position,
    float max_speed, float goal_position, float goal_velocity, float power,
    int *env_timestep, int episode_length) {
  if (!state || !action || !done || !reward || !obs || !env_timestep || n_envs <= 0)
    return (int)cudaErrorInvalidValue;
  if (misaligned(state, 7) || misaligned(obs, 7)) return (int)cudaErrorMisalignedAddress;
  continuous_mountain_car_step_kernel<<<(n_envs + kCcThreads - 1) / kCcThreads, kCcThreads, 0,
                                        as_stream(stream)>>>(
      n_envs, reinterpret_cast<float2 *>(state), action, done, reward,
      reinterpret_cast<float2 *>(obs), min_action, max_action, min_position, max_position,
      max_speed, goal_position, goal_velocity, power, env_timestep, episode_length);
  return finish_launch();
}

WDB_API int wdb_pendulum_step(void *stream, int n_envs, float *state, const float *action,
                              int *done, float *reward, float *obs, int *env_timestep,
                              int episode_length) {
  if (!state || !action || !done || !reward || !obs || !env_timestep || n_envs <= 0)
    return (int)cudaErrorInvalidValue;
  if (misaligned(state, 7)) return (int)cudaErrorMisalignedAddress;
  pendulum_step_kernel<<<(n_envs + kCcThreads - 1) / kCcThreads, kCcThreads, 0,
                         as_stream(stream)>>>(
      n_envs, reinterpret_cast<float2 *>(state), action, done, reward, obs, env_timestep,
      episode_length);
  return finish_launch();
}

WDB_API int wdb_acrobot_step(void *stream, int n_envs, float *state, const int *action,
                             int *done, float *reward, float *obs, int *env_timestep,
                             int episode_length) {
  if (!state || !action || !done || !reward || !obs || !env_timestep || n_envs <= 0)
    return (int)cudaErrorInvalidValue;
  if (misaligned(state, 15)) return (int)cudaErrorMisalignedAddress;
  acrobot_step_kernel<<<(n_envs + kCcThreads - 1) / kCcThreads, kCcThreads, 0,
                        as_stream(stream)>>>(
      n_envs, reinterpret_cast<float4 *>(state), action, done, reward, obs, env_timestep,
      episode_length);
  return finish_launch();
}

// wdb_small_envs.cu -- TagGridWorld and CartPole step kernels (sm_100a).
//
// Both envs have a handful of threads' worth of work per replica (5 agents / 1 agent),
// so the reference's "one block per env" launch (function_manager.py:65-67) puts 5 or 1
// threads in a CTA.  Here a CTA carries many replicas: gridworld packs
// floor(128 / n_agents) envs per CTA (one thread per agent), CartPole is one thread per
// env with float4 state I/O.  Both are HBM-streaming kernels.
#include "wdb_common.cuh"
#include "wdb_sa_physics.cuh"

using namespace wdb;

// ===================================================================== TagGridWorld
// CudaTagGridWorldStep, example_envs/tag_gridworld/tag_gridworld_step_pycuda.cu:112-251
// (+ CudaTagGridWorldGenerateObservation :9-110).  Integer state: results are bit-exact.
constexpr int kGwThreads = 128;

__global__ void __launch_bounds__(1024)
tag_gridworld_step_kernel(int n_envs, int N, int epb, int *__restrict__ loc_x,
                          int *__restrict__ loc_y, const int *__restrict__ actions,
                          int *__restrict__ done, float *__restrict__ rewards,
                          float *__restrict__ obs, float wall_hit_penalty,
                          float tag_reward_for_tagger, float tag_penalty_for_runner,
                          float step_cost_for_tagger, int use_full_observation, int B,
                          int *__restrict__ env_timestep, int episode_length,
                          const int *__restrict__ index_to_action, int tile_floats,
                          int tile_pad) {
  extern __shared__ __align__(16) int s_mem[];
  int *s_x = s_mem;                 // [epb*N]
  int *s_y = s_x + epb * N;         // [epb*N]
  int *s_tagged = s_y + epb * N;    // [epb]
  int *s_t = s_tagged + epb;        // [epb]
  int *s_closest = s_t + epb;       // [epb]
  int *s_move = s_closest + epb;    // [10]

  const int tid = threadIdx.x;
  const int le = tid / N;           // local env
  const int a = tid - le * N;       // agent
  const int env = blockIdx.x * epb + le;
  const bool active = le < epb && env < n_envs;

  if (tid < 10) s_move[tid] = index_to_action[tid];
  if (active && a == 0) {           // :138-141
    const int t = env_timestep[env] + 1;
    env_timestep[env] = t;
    s_t[le] = t;
    s_tagged[le] = 0;
  }
  __syncthreads();

  float rew = 0.0f;
  int x = 0, y = 0;
  if (active) {                     // :160-191 movement + wall clamp
    const int gi = env * N + a;
    const int ai = actions[gi] * 2;
    x = loc_x[gi] + s_move[ai];
    y = loc_y[gi] + s_move[ai + 1];
    if (x < 0) { x = 0; rew -= wall_hit_penalty; }
    else if (x > B) { x = B; rew -= wall_hit_penalty; }
    if (y < 0) { y = 0; rew -= wall_hit_penalty; }
    else if (y > B) { y = B; rew -= wall_hit_penalty; }
    loc_x[gi] = x;
    loc_y[gi] = y;
    s_x[le * N + a] = x;
    s_y[le * N + a] = y;
  }
  __syncthreads();
  const bool is_tagger = a < N - 1;
  if (active && is_tagger) {        // :200-205
    if (x == s_x[le * N + N - 1] && y == s_y[le * N + N - 1]) atomicAdd(&s_tagged[le], 1);
  }
  __syncthreads();
  if (active) {                     // :214-231
    const int tagged = s_tagged[le];
    if (is_tagger) { if (tagged > 0) rew += tag_reward_for_tagger; else rew -= step_cost_for_tagger; }
    else { if (tagged > 0) rew -= tag_penalty_for_runner; else rew += step_cost_for_tagger; }
    rewards[env * N + a] = rew;
    if (a == 0 && (s_t[le] == episode_length || tagged > 0)) done[env] = 1;  // :245-249
  }

  const float fB = static_cast<float>(B);
  const int envs_here = min(epb, n_envs - blockIdx.x * epb);
  if (use_full_observation) {
    // :25-51.  Every row r of an env is [x_0..x_{N-1}]/B, [y_*]/B, [is_runner_*],
    // onehot(r), t/T.
    const int F = 4 * N + 1;
    float *o = obs + (long long)blockIdx.x * epb * N * F;
    const int total = envs_here * N * F;
    if (tile_floats > 0) {
      // staged: each agent thread assembles ITS row in shared memory from per-env float
      // planes (one int->float division per agent, no index arithmetic per element), then the
      // CTA's tile -- contiguous in `obs` -- leaves as 16-byte vectors
      float *s_fx = reinterpret_cast<float *>(s_move + 10);     // [epb*N] x / B
      float *s_fy = s_fx + epb * N;                              // [epb*N] y / B
      float *tile = s_fy + epb * N + tile_pad;                   // 16-byte aligned
      if (active) {
        s_fx[le * N + a] = x / fB;
        s_fy[le * N + a] = y / fB;
      }
      __syncthreads();
      if (active) {
        float *row = tile + (le * N + a) * F;
        const float *fx = s_fx + le * N, *fy = s_fy + le * N;
        for (int j = 0; j < N; j++) {
          row[j] = fx[j];
          row[N + j] = fy[j];
          row[2 * N + j] = (j == N - 1) ? 1.0f : 0.0f;
          row[3 * N + j] = (j == a) ? 1.0f : 0.0f;
        }
        row[4 * N] = s_t[le] / static_cast<float>(episode_length);
      }
      __syncthreads();
      if ((((uintptr_t)o) & 15) == 0 && (total & 3) == 0) {
        const float4 *src = reinterpret_cast<const float4 *>(tile);
        float4 *dst = reinterpret_cast<float4 *>(o);
        for (int i = tid; i < total / 4; i += blockDim.x) dst[i] = src[i];
      } else {
        for (int i = tid; i < total; i += blockDim.x) o[i] = tile[i];
      }
    } else {
      // large envs (the tile does not fit shared memory): unit-stride element stores
      for (int i = tid; i < total; i += blockDim.x) {
        const int e = i / (N * F);
        const int rem = i - e * (N * F);
        const int row = rem / F;
        const int col = rem - row * F;
        float v;
        if (col < N) v = s_x[e * N + col] / fB;
        else if (col < 2 * N) v = s_y[e * N + col - N] / fB;
        else if (col < 3 * N) v = (col - 2 * N == N - 1) ? 1.0f : 0.0f;
        else if (col < 4 * N) v = (col - 3 * N == row) ? 1.0f : 0.0f;
        else v = s_t[e] / static_cast<float>(episode_length);
        o[i] = v;
      }
    }
  } else {
    // :52-109 partial observation: 6 floats per agent
    if (active && a == N - 1) {     // runner: nearest tagger, first-min, strict <
      int closest = 0, min_d = 2 * B * B;
      for (int g = 0; g < N - 1; g++) {
        const int dx = s_x[le * N + g] - x, dy = s_y[le * N + g] - y;
        const int d = dx * dx + dy * dy;  // exact ints (ref: pow() in f64, truncated)
        if (d < min_d) { min_d = d; closest = g; }
      }
      s_closest[le] = closest;
    }
    __syncthreads();
    if (active) {
      const int other = is_tagger ? N - 1 : s_closest[le];
      float *o = obs + ((long long)env * N + a) * 6;
      o[0] = x / fB;
      o[1] = y / fB;
      o[2] = s_x[le * N + other] / fB;
      o[3] = s_y[le * N + other] / fB;
      o[4] = is_tagger ? 0.0f : 1.0f;
      o[5] = s_t[le] / static_cast<float>(episode_length);
    }
  }
}

WDB_API int wdb_tag_gridworld_step(void *stream, int n_envs, int n_agents, int *loc_x,
                                   int *loc_y, const int *actions, int *done,
                                   float *rewards, float *obs, float wall_hit_penalty,
                                   float tag_reward_for_tagger,
                                   float tag_penalty_for_runner,
                                   float step_cost_for_tagger, int use_full_observation,
                                   int world_boundary, int *env_timestep,
                                   int episode_length, const int *index_to_action) {
  if (!loc_x || !loc_y || !actions || !done || !rewards || !obs || !env_timestep ||
      !index_to_action || n_envs <= 0 || n_agents < 2 || n_agents > 1024)
    return (int)cudaErrorInvalidValue;
  int epb = n_agents >= kGwThreads ? 1 : kGwThreads / n_agents;
  // full observations are staged in a shared-memory tile that leaves as 16-byte vectors:
  // keep the CTA's tile (epb * N * (4N + 1) floats) a multiple of 16 bytes where possible
  const int F = 4 * n_agents + 1;
  if (use_full_observation && epb >= 4 && ((epb * n_agents * F) & 3) != 0) epb &= ~3;
  const int block = max(32, round_up(epb * n_agents, 32));
  const int grid = (n_envs + epb - 1) / epb;
  const int n_int = 2 * epb * n_agents + 3 * epb + 10;
  size_t smem = sizeof(int) * (size_t)n_int;
  int tile_floats = 0, tile_pad = 0;
  if (use_full_observation) {
    const size_t planes = 2ull * epb * n_agents;
    const size_t head = (size_t)n_int + planes;
    tile_pad = (int)((4 - (head & 3)) & 3);
    const size_t tile = (size_t)epb * n_agents * F;
    if ((head + tile_pad + tile) * 4 <= 96 * 1024) {
      tile_floats = (int)tile;
      smem = (head + tile_pad + tile) * 4;
      static size_t configured = 0;
      if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(tag_gridworld_step_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
      }
    } else {
      tile_pad = 0;
    }
  }
  tag_gridworld_step_kernel<<<grid, block, smem, as_stream(stream)>>>(
      n_envs, n_agents, epb, loc_x, loc_y, actions, done, rewards, obs, wall_hit_penalty,
      tag_reward_for_tagger, tag_penalty_for_runner, step_cost_for_tagger,
      use_full_observation, world_boundary, env_timestep, episode_length,
      index_to_action, tile_floats, tile_pad);
  return finish_launch();
}

// ========================================================================= CartPole
// NumbaClassicControlCartPoleEnvStep, example_envs/single_agent/classic_control/
// cartpole/cartpole_step_numba.py:6-83.  One thread per env; the reference launches one
// 1-thread block per env.  Arithmetic follows numba's typing of that source: float32
// everywhere except where the float64 literal 4.0/3.0 promotes (thetaacc, xacc and the
// two velocity updates); checked against numba's type annotations of the reference kernel.
__global__ void __launch_bounds__(256)
cartpole_step_kernel(int n_envs, float4 *__restrict__ state, const int *__restrict__ action,
                     int *__restrict__ done, float *__restrict__ reward,
                     float4 *__restrict__ obs, float gravity, float masspole,
                     float total_mass, float length, float polemass_length,
                     float force_mag, float tau, float theta_thr, float x_thr,
                     int *__restrict__ env_timestep, int episode_length) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n_envs) return;
  const int t = env_timestep[env] + 1;
  env_timestep[env] = t;
  const float4 n = cartpole_physics(state[env], action[env], gravity, masspole, total_mass,
                                    length, polemass_length, force_mag, tau);
  state[env] = n;
  obs[env] = n;
  const bool terminated = (n.x < -x_thr) || (n.x > x_thr) || (n.z < -theta_thr) || (n.z > theta_thr);
  reward[env] = 1.0f;
  if (t == episode_length || terminated) done[env] = 1;
}

WDB_API int wdb_cartpole_step(void *stream, int n_envs, float *state, const int *action,
                              int *done, float *reward, float *obs, float gravity,
                              float masspole, float total_mass, float length,
                              float polemass_length, float force_mag, float tau,
                              float theta_threshold_radians, float x_threshold,
                              int *env_timestep, int episode_length) {
  if (!state || !action || !done || !reward || !obs || !env_timestep || n_envs <= 0)
    return (int)cudaErrorInvalidValue;
  if (((uintptr_t)state | (uintptr_t)obs) & 15) return (int)cudaErrorMisalignedAddress;
  const int block = 256;
  cartpole_step_kernel<<<(n_envs + block - 1) / block, block, 0, as_stream(stream)>>>(
      n_envs, reinterpret_cast<float4 *>(state), action, done, reward,
      reinterpret_cast<float4 *>(obs), gravity, masspole, total_mass, length,
      polemass_length, force_mag, tau, theta_threshold_radians, x_threshold,
      env_timestep, episode_length);
  return finish_launch();
}

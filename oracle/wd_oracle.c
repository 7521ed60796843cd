/*
 * wd_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C restatement of the reference's *CUDA* rollout kernels, used only as
 * the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg.  Nothing under warp_drive_b200/ may import, link or call this file.
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference).  The restatement follows the CUDA-C
 * path op for op: float32 state, float64 distance math, the swap-based partial
 * selection sort (tie order!), strict-< first-min nearest-tagger search.
 *
 * Known, documented deviations from the reference CUDA binary:
 *   - sinf/cosf come from glibc, not CUDA libdevice (<= 2 ulp apart), so
 *     positions can differ in the last bit; parity tests use 1e-5 tolerance
 *     for floats and teacher-force the state where bit-equality is asserted.
 *   - the reference has two data races in the tag phase
 *     (tag_continuous_step_pycuda.cu:324-329: `rewards[tagger] +=` and
 *     `num_runners[env] -= 1` are non-atomic).  The oracle uses the race-free
 *     semantics (every tag counted), which is what the reference's NumPy env
 *     does (tag_continuous.py:660-672).
 *
 * Pinning: see oracle/README.md -- gridworld is pinned by the reference's own
 * golden vectors (tests/golden/gridworld_cuda_golden.npz), tag_continuous by
 * fixtures generated from the reference NumPy env, and on a GPU box by the
 * reference's own kernels compiled into oracle/_ref/ (fatbin files).
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC wd_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define WD_EXPORT __attribute__((visibility("default")))

/* tag_continuous_step_pycuda.cu:7-9 : __constant__ float kPi/kTwoPi/kEpsilon */
static const float kTwoPi = 6.283185308f;
static const float kEpsilon = 1.0e-10f;

WD_EXPORT int wd_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

WD_EXPORT void wd_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* tag_continuous                                                           */
/* ------------------------------------------------------------------------ */

/* ComputeDistance, tag_continuous_step_pycuda.cu:13-26.
 * float subtraction, then pow(double,2)+pow(double,2), sqrt in double,
 * narrowed to float on return. */
static float tc_distance(const float *x, const float *y, int i1, int i2) {
  float dxf = x[i1] - x[i2];
  float dyf = y[i1] - y[i2];
  double dx = (double)dxf, dy = (double)dyf;
  return (float)sqrt(pow(dx, 2.0) + pow(dy, 2.0));
}

typedef struct {
  int n_agents, k_obs, episode_length;
  int use_full_obs, runner_exits;
  float edge_hit_penalty, grid_length, max_speed;
  float margin, tag_reward, tag_penalty, end_reward;
} tc_params;

/* One env of CudaTagContinuousStep (tag_continuous_step_pycuda.cu:351-520).
 * The device barriers (__sync_env_threads) become sequential phases. */
static void tc_step_one_env(const tc_params *P, int env, float *loc_x,
                            float *loc_y, float *speed, float *direction,
                            float *acceleration, const int *agent_types,
                            float *edge_pen, const float *acc_actions,
                            const float *turn_actions, const float *skill,
                            int *alive, float *obs, const int *actions,
                            float *nd /*[N-1] per agent scratch*/,
                            int *nids /*[N-1] per agent scratch*/,
                            int *nearest /*[N,K]*/, float *rewards,
                            const float *step_rewards, int *num_runners,
                            int *done, int *timestep, int scratch_is_global) {
  const int N = P->n_agents;
  const int K = P->k_obs;
  const float L = P->grid_length;
  const int base = env * N;
  float *x = loc_x + base, *y = loc_y + base, *sp = speed + base;
  float *dir = direction + base, *acc = acceleration + base;
  float *ep = edge_pen + base, *rew = rewards + base;
  int *al = alive + base;

  /* :391-393 timestep++ by agent 0 */
  timestep[env] += 1;
  const int t = timestep[env];

  /* :402-465 kinematics */
  for (int a = 0; a < N; a++) {
    const int *act = actions + (size_t)(base + a) * 2;
    float d_acc = acc_actions[act[0]];
    float d_turn = turn_actions[act[1]];
    acc[a] = acc[a] + d_acc;
    /* :413-419 fmod(float,float) -> fmodf; times int flag */
    float nd_ = fmodf(dir[a] + d_turn, kTwoPi) * (float)al[a];
    if (nd_ < 0) nd_ = kTwoPi + nd_;
    dir[a] = nd_;
    /* :421-426 min(float, max(0.0, float)) * int : evaluated in double */
    float cap = P->max_speed * skill[a];
    double s = fmin((double)cap, fmax(0.0, (double)(sp[a] + acc[a]))) *
               (double)al[a];
    sp[a] = (float)s;
    /* :430-434 */
    if ((sp[a] <= 0.0) || (sp[a] >= cap)) acc[a] = 0.0f;
    /* :436-439  nvcc contracts `x += s*cos(d)` into one fma (default
     * -fmad=true); restated with fmaf so the CPU result rounds once too. */
    x[a] = fmaf(sp[a], cosf(dir[a]), x[a]);
    y[a] = fmaf(sp[a], sinf(dir[a]), y[a]);
    /* :442-463 */
    int crossed = (x[a] < 0) | (x[a] > L) | (y[a] < 0) | (y[a] > L);
    if (crossed) {
      if (x[a] < 0) x[a] = 0.0f; else if (x[a] > L) x[a] = L;
      if (y[a] < 0) y[a] = 0.0f; else if (y[a] > L) y[a] = L;
      ep[a] = P->edge_hit_penalty;
    } else {
      ep[a] = 0.0f;
    }
  }

  /* CudaTagContinuousGenerateObservation :29-256 */
  const double diag = sqrt(2.0) * (double)L;      /* :94 double normaliser */
  const float vnorm = P->max_speed + kEpsilon;    /* :101 float add */
  if (P->use_full_obs) {
    const int M = N - 1, F = 7 * M + 1;
    for (int a = 0; a < N; a++) {
      float *o = obs + ((size_t)env * N + a) * F;
      int idx = 0;
      for (int b = 0; b < N; b++) {               /* :63-85 */
        if (b == a) continue;
        o[0 * M + idx] = 0.0f; o[1 * M + idx] = 0.0f; o[2 * M + idx] = 0.0f;
        o[3 * M + idx] = 0.0f; o[4 * M + idx] = 0.0f;
        o[5 * M + idx] = (float)agent_types[b];
        o[6 * M + idx] = (float)al[b];
        idx++;
      }
      o[7 * M] = 0.0f;
      if (al[a]) {                                /* :88-113 */
        idx = 0;
        for (int b = 0; b < N; b++) {
          if (b == a) continue;
          o[0 * M + idx] = (float)((double)(float)(x[b] - x[a]) / diag);
          o[1 * M + idx] = (float)((double)(float)(y[b] - y[a]) / diag);
          o[2 * M + idx] = (float)(sp[b] - sp[a]) / vnorm;
          o[3 * M + idx] = (float)(acc[b] - acc[a]) / vnorm;
          o[4 * M + idx] = (float)(dir[b] - dir[a]) / kTwoPi;
          idx++;
        }
        o[7 * M] = (float)t / P->episode_length;
      }
    }
  } else {
    const int F = 7 * K + 1;
    for (int a = 0; a < N; a++) {
      float *o = obs + ((size_t)env * N + a) * F;
      for (int i = 0; i < F; i++) o[i] = 0.0f;    /* :121-139 */
      if (!al[a]) continue;
      float *d = scratch_is_global ? nd + ((size_t)env * N + a) * (N - 1) : nd;
      int *ids = scratch_is_global ? nids + ((size_t)env * N + a) * (N - 1) : nids;
      int nv = 0;
      for (int b = 0; b < N; b++)                 /* :154-164 */
        if (b != a && al[b]) ids[nv++] = b;
      for (int i = 0; i < nv; i++)                /* :167-176 */
        d[i] = tc_distance(x, y, a, ids[i]);
      const int kk = nv < K ? nv : K;
      for (int i = 0; i < kk; i++) {              /* :179-199 swap selection */
        for (int j = i + 1; j < nv; j++) {
          if (d[j] < d[i]) {
            float td = d[i]; d[i] = d[j]; d[j] = td;
            int ti = ids[i]; ids[i] = ids[j]; ids[j] = ti;
          }
        }
      }
      int *nn = nearest + ((size_t)env * N + a) * K;
      for (int i = 0; i < kk; i++) nn[i] = ids[i]; /* :202-211 */
      for (int i = 0; i < kk; i++) {              /* :214-250 */
        const int b = nn[i];
        o[0 * K + i] = (float)((double)(float)(x[b] - x[a]) / diag);
        o[1 * K + i] = (float)((double)(float)(y[b] - y[a]) / diag);
        o[2 * K + i] = (float)(sp[b] - sp[a]) / vnorm;
        o[3 * K + i] = (float)(acc[b] - acc[a]) / vnorm;
        o[4 * K + i] = (float)(dir[b] - dir[a]) / kTwoPi;
        o[5 * K + i] = (float)agent_types[b];
        o[6 * K + i] = (float)al[b];
      }
      o[7 * K] = (float)t / P->episode_length;    /* :251-253 */
    }
  }

  /* CudaTagContinuousComputeReward :259-349 */
  for (int a = 0; a < N; a++) {                   /* :283-291 */
    float r = 0.0f;
    if (al[a]) { r += ep[a]; r += step_rewards[a]; }
    rew[a] = r;
  }
  for (int a = 0; a < N; a++) {                   /* :296-338 */
    if (agent_types[a] != 0 || !al[a]) continue;
    float min_dist = (float)((double)L * sqrt(2.0));
    int nearest_tagger = -1;
    for (int b = 0; b < N; b++) {
      if (agent_types[b] != 1) continue;
      float dd = tc_distance(x, y, a, b);
      if (dd < min_dist) { min_dist = dd; nearest_tagger = b; }
    }
    if (min_dist < P->margin) {
      rew[a] += P->tag_penalty;
      rew[nearest_tagger] += P->tag_reward;       /* race-free accumulation */
      if (P->runner_exits) { al[a] = 0; num_runners[env] -= 1; }
    }
    if (t == P->episode_length) rew[a] += P->end_reward;  /* :334-337 */
  }
  /* :341-348 */
  if (t == P->episode_length || num_runners[env] == 0) done[env] = 1;
}

/* Argument order == CudaTagContinuousStep signature
 * (tag_continuous_step_pycuda.cu:351-385) with n_envs prepended. */
WD_EXPORT void wd_oracle_tag_continuous_step(
    int n_envs, float *loc_x, float *loc_y, float *speed, float *direction,
    float *acceleration, const int *agent_types, float *edge_hit_reward_penalty,
    float edge_hit_penalty, float grid_length, const float *acceleration_actions,
    const float *turn_actions, float max_speed, int num_other_agents_observed,
    const float *skill_levels, int runner_exits_game_after_tagged,
    int *still_in_the_game, int use_full_observation, float *obs,
    const int *action_indices, float *neighbor_distances,
    int *neighbor_ids_sorted_by_distance, int *nearest_neighbor_ids,
    float *rewards, const float *step_rewards, int *num_runners,
    float distance_margin_for_reward, float tag_reward_for_tagger,
    float tag_penalty_for_runner, float end_of_game_reward_for_runner,
    int *done, int *env_timestep, int n_agents, int episode_length) {
  tc_params P;
  P.n_agents = n_agents; P.k_obs = num_other_agents_observed;
  P.episode_length = episode_length; P.use_full_obs = use_full_observation;
  P.runner_exits = runner_exits_game_after_tagged;
  P.edge_hit_penalty = edge_hit_penalty; P.grid_length = grid_length;
  P.max_speed = max_speed; P.margin = distance_margin_for_reward;
  P.tag_reward = tag_reward_for_tagger; P.tag_penalty = tag_penalty_for_runner;
  P.end_reward = end_of_game_reward_for_runner;
  const int global_scratch =
      (neighbor_distances != NULL && neighbor_ids_sorted_by_distance != NULL);
#pragma omp parallel
  {
    float *nd = NULL; int *nids = NULL;
    if (!global_scratch) {
      nd = (float *)malloc(sizeof(float) * (size_t)(n_agents > 1 ? n_agents : 1));
      nids = (int *)malloc(sizeof(int) * (size_t)(n_agents > 1 ? n_agents : 1));
    }
#pragma omp for schedule(static)
    for (int env = 0; env < n_envs; env++) {
      tc_step_one_env(&P, env, loc_x, loc_y, speed, direction, acceleration,
                      agent_types, edge_hit_reward_penalty, acceleration_actions,
                      turn_actions, skill_levels, still_in_the_game, obs,
                      action_indices,
                      global_scratch ? neighbor_distances : nd,
                      global_scratch ? neighbor_ids_sorted_by_distance : nids,
                      nearest_neighbor_ids, rewards, step_rewards, num_runners,
                      done, env_timestep, global_scratch);
    }
    free(nd); free(nids);
  }
}

/* ------------------------------------------------------------------------ */
/* tag_gridworld : CudaTagGridWorldStep tag_gridworld_step_pycuda.cu:112-251 */
/* ------------------------------------------------------------------------ */
WD_EXPORT void wd_oracle_tag_gridworld_step(
    int n_envs, int n_agents, int *loc_x, int *loc_y, const int *actions,
    int *done, float *rewards, float *obs, float wall_hit_penalty,
    float tag_reward_for_tagger, float tag_penalty_for_runner,
    float step_cost_for_tagger, int use_full_observation, int world_boundary,
    int *env_timestep, int episode_length,
    const int *index_to_action /* kIndexToActionArr[10], :6 */) {
  const int N = n_agents;
  const int B = world_boundary;
#pragma omp parallel for schedule(static)
  for (int env = 0; env < n_envs; env++) {
    int *x = loc_x + env * N, *y = loc_y + env * N;
    float *rew = rewards + env * N;
    env_timestep[env] += 1;                       /* :138-141 */
    const int t = env_timestep[env];
    int tagged = 0;
    float *rr = (float *)malloc(sizeof(float) * (size_t)N);
    for (int a = 0; a < N; a++) {                 /* :160-191 */
      float r = 0.0f;
      int ai = actions[env * N + a] * 2;
      x[a] = x[a] + index_to_action[ai];
      y[a] = y[a] + index_to_action[ai + 1];
      if (x[a] < 0) { x[a] = 0; r -= wall_hit_penalty; }
      else if (x[a] > B) { x[a] = B; r -= wall_hit_penalty; }
      if (y[a] < 0) { y[a] = 0; r -= wall_hit_penalty; }
      else if (y[a] > B) { y[a] = B; r -= wall_hit_penalty; }
      rr[a] = r;
    }
    for (int a = 0; a < N - 1; a++)               /* :200-205 */
      if (x[a] == x[N - 1] && y[a] == y[N - 1]) tagged++;
    for (int a = 0; a < N; a++) {                 /* :214-231 */
      float r = rr[a];
      if (a < N - 1) { if (tagged > 0) r += tag_reward_for_tagger; else r -= step_cost_for_tagger; }
      else { if (tagged > 0) r -= tag_penalty_for_runner; else r += step_cost_for_tagger; }
      rew[a] = r;
    }
    free(rr);
    /* CudaTagGridWorldGenerateObservation :9-110 */
    if (use_full_observation) {
      const int F = 4 * N + 1;
      float *o = obs + (size_t)env * N * F;
      for (int a = 0; a < N; a++) {               /* thread a writes column a */
        for (int row = 0; row < N; row++) {
          float *orow = o + (size_t)row * F;
          orow[a] = (float)x[a] / (float)B;
          orow[N + a] = (float)y[a] / (float)B;
          orow[2 * N + a] = (float)(1.0 * (int)(a == N - 1));
          orow[3 * N + a] = (float)(1.0 * (int)(row == a));
          if (a == N - 1) orow[4 * N] = (float)t / (float)episode_length;
        }
      }
    } else {
      float *o = obs + (size_t)env * N * 6;
      int closest = 0, min_distance = 2 * B * B;  /* :86-93 */
      for (int a = 0; a < N - 1; a++) {
        int dist = (int)(pow((double)(x[a] - x[N - 1]), 2.0) +
                         pow((double)(y[a] - y[N - 1]), 2.0));
        if (dist < min_distance) { min_distance = dist; closest = a; }
      }
      for (int a = 0; a < N; a++) {
        float *oa = o + a * 6;
        oa[0] = (float)x[a] / (float)B;
        oa[1] = (float)y[a] / (float)B;
        int other = (a < N - 1) ? N - 1 : closest;
        oa[2] = (float)x[other] / (float)B;
        oa[3] = (float)y[other] / (float)B;
        oa[4] = (float)(1.0 * (int)(a == N - 1));
        oa[5] = (float)t / (float)episode_length;
      }
    }
    if (t == episode_length || tagged > 0) done[env] = 1;  /* :245-249 */
  }
}

/* ------------------------------------------------------------------------ */
/* CartPole: NumbaClassicControlCartPoleEnvStep                             */
/* example_envs/single_agent/classic_control/cartpole/cartpole_step_numba.py:6-83
 * Numba typing: arrays/scalars are float32; the literal 4.0/3.0 is float64 and
 * promotes thetaacc, xacc, x_dot', theta_dot' to float64 before the float32
 * store (numba BinOp typing; `** 2` on float32 stays float32).              */
/* ------------------------------------------------------------------------ */
WD_EXPORT void wd_oracle_cartpole_step(
    int n_envs, float *state /*[E,1,4]*/, const int *action /*[E,1,1]*/,
    int *done, float *reward /*[E,1]*/, float *obs /*[E,1,4]*/, float gravity,
    float masspole, float total_mass, float length, float polemass_length,
    float force_mag, float tau, float theta_threshold_radians, float x_threshold,
    int *env_timestep, int episode_length) {
#pragma omp parallel for schedule(static)
  for (int env = 0; env < n_envs; env++) {
    env_timestep[env] += 1;
    float *s = state + (size_t)env * 4;
    float x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
    float force = (action[env] > 0.5) ? force_mag : -force_mag;
    float costheta = cosf(theta), sintheta = sinf(theta);
    /* fmaf()/fma() where the reference binary (ptxas on numba's PTX) fuses */
    float temp = fmaf(polemass_length * (theta_dot * theta_dot), sintheta, force) / total_mass;
    float c2m = masspole * (costheta * costheta) / total_mass;
    float torque = fmaf(gravity, sintheta, -(costheta * temp));
    double thetaacc = (double)torque / ((double)length * (4.0 / 3.0 - (double)c2m));
    double xacc = (double)temp - (double)polemass_length * thetaacc * (double)costheta / (double)total_mass;
    float nx = fmaf(tau, x_dot, x);
    float nx_dot = (float)fma((double)tau, xacc, (double)x_dot);
    float ntheta = fmaf(tau, theta_dot, theta);
    float ntheta_dot = (float)fma((double)tau, thetaacc, (double)theta_dot);
    s[0] = nx; s[1] = nx_dot; s[2] = ntheta; s[3] = ntheta_dot;
    float *o = obs + (size_t)env * 4;
    o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[3];
    int terminated = (nx < -x_threshold) || (nx > x_threshold) ||
                     (ntheta < -theta_threshold_radians) ||
                     (ntheta > theta_threshold_radians);
    reward[env] = 1.0f;
    if (env_timestep[env] == episode_length || terminated) done[env] = 1;
  }
}

/* ------------------------------------------------------------------------ */
/* Sampler: sample_actions + search_index, cuda_includes/core/random.cu:33-85 */
/* The uniform draw p is an INPUT (the reference RNG stream is unpinned).    */
/* ------------------------------------------------------------------------ */
static int search_index(const float *distr, float p, int l, int r) {
  const float kEps = 1.0e-8f;                     /* random.cu:9 */
  int left = l, right = r, mid;
  while (left <= right) {
    mid = left + (right - left) / 2;
    if (fabsf(distr[mid] - p) < kEps) return mid - l;
    else if (distr[mid] < p) left = mid + 1;
    else right = mid - 1;
  }
  return left > r ? r - l : left - l;
}

WD_EXPORT void wd_oracle_sample_actions(
    const float *distr /*[n, A]*/, int *action_indices, int action_stride,
    float *cum_distr /*[n, A] or NULL*/, const float *uniforms /*[n]*/,
    long n /* n_envs*n_agents */, int num_actions, int use_argmax) {
  float *tmp = (float *)malloc(sizeof(float) * (size_t)num_actions);
  for (long pos = 0; pos < n; pos++) {
    const float *d = distr + pos * num_actions;
    if (use_argmax) {                             /* random.cu:58-69 */
      float max_p = d[0]; int max_ind = 0;
      for (int i = 1; i < num_actions; i++)
        if (max_p < d[i]) { max_p = d[i]; max_ind = i; }
      action_indices[pos * action_stride] = max_ind;
      continue;
    }
    float *c = cum_distr ? cum_distr + pos * num_actions : tmp;
    c[0] = d[0];                                  /* random.cu:75-80 */
    for (int i = 1; i < num_actions; i++) c[i] = d[i] + c[i - 1];
    action_indices[pos * action_stride] =
        search_index(c, uniforms[pos], 0, num_actions - 1);
  }
  free(tmp);
}

/* ------------------------------------------------------------------------ */
/* Reset: reset_in_*_when_done_{2d,3d} + undo_done_flag_and_reset_timestep   */
/* cuda_includes/core/reset.cu:9-75.  Element size is 4 bytes for both the   */
/* float and int variants, so one byte-copy covers all four kernels.         */
/* ------------------------------------------------------------------------ */
WD_EXPORT void wd_oracle_reset_when_done(void *data, const void *ref,
                                          const int *done, int n_envs,
                                          long elems_per_env, int force_reset) {
  for (int env = 0; env < n_envs; env++) {
    if (force_reset > 0 || done[env] > 0)
      memcpy((char *)data + (size_t)env * elems_per_env * 4,
             (const char *)ref + (size_t)env * elems_per_env * 4,
             (size_t)elems_per_env * 4);
  }
}

WD_EXPORT void wd_oracle_undo_done_and_reset_timestep(int *done, int *timestep,
                                                       int n_envs,
                                                       int force_reset) {
  for (int env = 0; env < n_envs; env++) {
    if (force_reset > 0 || done[env] > 0) { done[env] = 0; timestep[env] = 0; }
  }
}

/* ------------------------------------------------------------------------ */
/* OU process: sample_ou_process, numba_includes/core/random.py:74-105.      */
/* The normal draw is an INPUT (numba's xoroshiro stream is unpinned).       */
/* ------------------------------------------------------------------------ */
WD_EXPORT void wd_oracle_ou_process(const float *distr, float *actions,
                                     float *ou_states, const float *normals,
                                     long n, float damping, float stddev,
                                     float scale) {
  const float kEps = 1.0e-8f;
  for (long i = 0; i < n; i++) {
    if (scale < kEps) { actions[i] = distr[i]; continue; }
    float nv = stddev * normals[i];
    ou_states[i] = (1.0f - damping) * ou_states[i] + nv;
    actions[i] = distr[i] + scale * ou_states[i];
  }
}

/* ------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11; Random123 v1.14 philox.h) -- the     */
/* product's device RNG.  This scalar restatement + the Random123 known-     */
/* answer vectors (tests/test_oracle_cpu.py) pin the stream bit-for-bit.     */
/* ------------------------------------------------------------------------ */
WD_EXPORT void wd_oracle_philox4x32_10(const uint32_t ctr[4],
                                        const uint32_t key[2],
                                        uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ------------------------------------------------------------------------ */
/* Classic control: the reference has these only    */
/* as numba kernels under example_envs/single_agent/classic_control/.        */
/* Arithmetic types follow numba's type inference of that Python source      */
/* (float32 arrays and np.float32 scalar arguments; int * float32 and every  */
/* expression with a Python float literal are float64; math.sin/cos pick the */
/* float32 routine for float32 arguments), read off numba's own type         */
/* annotations; fma() is written where the reference binary (ptxas on        */
/* numba's PTX, oracle/build_ref_numba.py) fuses.  Pinning: on the GPU box    */
/* against those reference binaries (tests/test_gpu_classic_control.py) and  */
/* through the fixtures they produced there (tests/golden/                   */
/* classic_control_numba_golden.npz); on CPU against float64 restatements of */
/* the gym physics the reference's CPU envs delegate to.                     */
/* ------------------------------------------------------------------------ */
static double wd_clip(double v, double lo, double hi) {
  if (v < lo) return lo;                 /* mountain_car_step_numba.py:5-11 */
  if (v > hi) return hi;
  return v;
}

/* NumbaClassicControlMountainCarEnvStep, mountain_car/mountain_car_step_numba.py:14-70 */
WD_EXPORT void wd_oracle_mountain_car_step(
    int n_envs, float *state /*[E,1,2]*/, const int *action /*[E,1,1]*/, int *done,
    float *reward /*[E,1]*/, float *obs /*[E,1,2]*/, float min_position,
    float max_position, float max_speed, float goal_position, float goal_velocity,
    float force, float gravity, int *env_timestep, int episode_length) {
  for (int env = 0; env < n_envs; env++) {
    env_timestep[env] += 1;                                          /* :35 */
    float *s = state + (size_t)env * 2;
    double position = (double)s[0], velocity = (double)s[1];
    /* :46  velocity += (action - 1) * force + math.cos(3 * position) * (-gravity) */
    double c = cos(3.0 * position);
    velocity = fma((double)force, (double)(long)(action[env] - 1), -((double)gravity * c)) +
               velocity;
    velocity = wd_clip(velocity, (double)(-max_speed), (double)max_speed);   /* :47 */
    position = position + velocity;                                          /* :48 */
    position = wd_clip(position, (double)min_position, (double)max_position);
    if (position == (double)min_position && velocity < 0.0) velocity = 0.0;  /* :50-51 */
    s[0] = (float)position; s[1] = (float)velocity;                          /* :53-54 */
    obs[(size_t)env * 2] = s[0]; obs[(size_t)env * 2 + 1] = s[1];
    int terminated = position >= (double)goal_position &&
                     velocity >= (double)goal_velocity;                      /* :59-61 */
    reward[env] = -1.0f;                                                     /* :64 */
    if (env_timestep[env] == episode_length) done[env] = 1;                  /* :66-69 */
    else if (terminated) done[env] = 2;
  }
}

/* NumbaClassicControlContinuousMountainCarEnvStep,
 * continuous_mountain_car/continuous_mountain_car_step_numba.py:14-71 */
WD_EXPORT void wd_oracle_continuous_mountain_car_step(
    int n_envs, float *state, const float *action /*[E,1,1] float32*/, int *done,
    float *reward, float *obs, float min_action, float max_action, float min_position,
    float max_position, float max_speed, float goal_position, float goal_velocity,
    float power, int *env_timestep, int episode_length) {
  for (int env = 0; env < n_envs; env++) {
    env_timestep[env] += 1;
    float *s = state + (size_t)env * 2;
    float a = action[env];
    float f = a;                                      /* :44 _clip on float32 values */
    if (a < min_action) f = min_action; else if (a > max_action) f = max_action;
    float fp = f * power;                             /* float32 product */
    double position = (double)s[0], velocity = (double)s[1];
    double c = cos(3.0 * position);
    velocity = fma(c, -0.0025, (double)fp) + velocity;                       /* :46 */
    velocity = wd_clip(velocity, (double)(-max_speed), (double)max_speed);
    position = position + velocity;
    position = wd_clip(position, (double)min_position, (double)max_position);
    if (position == (double)min_position && velocity < 0.0) velocity = 0.0;
    s[0] = (float)position; s[1] = (float)velocity;
    obs[(size_t)env * 2] = s[0]; obs[(size_t)env * 2 + 1] = s[1];
    int terminated = position >= (double)goal_position &&
                     velocity >= (double)goal_velocity;
    double rew = terminated ? 100.0 : 0.0;                                   /* :64-66 */
    rew -= pow((double)a, 2.0) * 0.1;                                        /* :67 */
    reward[env] = (float)rew;
    if (env_timestep[env] == episode_length || terminated) done[env] = 1;    /* :70-71 */
  }
}

/* `x % (2*pi)` as NVVM lowers Python's float modulo: a - floor(|a|/b)*b (one fma), sign of
 * the dividend restored, then "result takes the divisor's sign". */
static double wd_python_mod_2pi(double a) {
  const double b = 2.0 * 3.141592653589793;
  double q = floor(fabs(a) / b);
  double r = fma(-q, b, fabs(a));
  if (!(a >= 0.0)) r = -r;
  if (r < 0.0) r = r + b;
  return r;
}

/* NumbaClassicControlPendulumEnvStep, pendulum/pendulum_step_numba.py:30-72
 * (module constants :9-14: max_speed 8, max_torque 2, dt 0.05, g 9.81, m = l = 1) */
WD_EXPORT void wd_oracle_pendulum_step(
    int n_envs, float *state /*[E,1,2]*/, const float *action /*[E,1,1]*/, int *done,
    float *reward, float *obs /*[E,1,3]*/, int *env_timestep, int episode_length) {
  const double kPi = 3.141592653589793, dt = 0.05;
  const double gain = 3 * 9.81 / (2 * 1.0);
  for (int env = 0; env < n_envs; env++) {
    env_timestep[env] += 1;
    float *s = state + (size_t)env * 2;
    double u = wd_clip((double)action[env], -2.0, 2.0);                     /* :51 */
    double th = (double)s[0], thdot = (double)s[1];
    double an = wd_python_mod_2pi(th + kPi) - kPi;                           /* :26-27 */
    double td2 = (double)(s[1] * s[1]);               /* float32 ** 2 stays float32 */
    double costs = fma(u * u, 0.001, fma(an, an, td2 * 0.1));                /* :56 */
    double newthdot = fma(fma(u, 3.0, (double)sinf(s[0]) * gain), dt, thdot); /* :58 */
    newthdot = wd_clip(newthdot, -8.0, 8.0);                                 /* :59 */
    double newth = fma(newthdot, dt, th);                                    /* :60 */
    s[0] = (float)newth; s[1] = (float)newthdot;
    float *o = obs + (size_t)env * 3;
    o[0] = (float)cos(newth); o[1] = (float)sin(newth); o[2] = (float)newthdot;
    reward[env] = (float)(-costs);
    if (env_timestep[env] == episode_length) done[env] = 1;
  }
}

/* _dsdt, acrobot/acrobot_step_numba.py:70-109 (all link constants 1.0 / 0.5, g = 9.8) */
static void wd_acrobot_dsdt(const float *st, double torque, float *d) {
  const double kPi = 3.141592653589793;
  float theta1 = st[0], theta2 = st[1], dtheta1 = st[2], dtheta2 = st[3];
  double c2 = (double)cosf(theta2), s2 = (double)sinf(theta2);
  double d1 = ((0.25 + (1.25 + c2)) + 1.0) + 1.0;                           /* :85-90 */
  double d2 = (0.25 + 0.5 * c2) + 1.0;                                       /* :91 */
  double phi2 = (1.0 * 0.5 * 9.8) * cos((double)(float)(theta1 + theta2) - kPi / 2);
  double phi1 = ((-0.5 * (double)(float)(dtheta2 * dtheta2)) * s2
                 - ((double)dtheta2 * (double)dtheta1) * s2
                 + ((1.0 * 0.5 + 1.0 * 1.0) * 9.8) * cos((double)theta1 - kPi / 2))
                + phi2;                                                      /* :93-98 */
  double ddtheta2 = (torque + d2 / d1 * phi1
                     - (0.5 * (double)(float)(dtheta1 * dtheta1)) * s2 - phi2) /
                    ((0.25 + 1.0) - d2 * d2 / d1);                           /* :100-101 */
  double ddtheta1 = -(d2 * ddtheta2 + phi1) / d1;                            /* :102 */
  d[0] = dtheta1; d[1] = dtheta2; d[2] = (float)ddtheta1; d[3] = (float)ddtheta2;
}

static double wd_wrap(double x, double m, double M) {                        /* :137-143 */
  double diff = M - m;
  while (x > M) x = x - diff;
  while (x < m) x = x + diff;
  return x;
}

/* NumbaClassicControlAcrobotEnvStep, acrobot/acrobot_step_numba.py:24-67 + rk4 :112-134 */
WD_EXPORT void wd_oracle_acrobot_step(
    int n_envs, float *state /*[E,1,4]*/, const int *action /*[E,1,1]*/, int *done,
    float *reward, float *obs /*[E,1,6]*/, int *env_timestep, int episode_length) {
  const double kPi = 3.141592653589793;
  const double kMaxVel1 = 12.566370614359172, kMaxVel2 = 28.274333882308138;
  const double dt = 0.2, dt2 = 0.1;
  for (int env = 0; env < n_envs; env++) {
    env_timestep[env] += 1;
    float *s = state + (size_t)env * 4;
    double torque = (double)(action[env] - 1);          /* AVAIL_TORQUE[action], :6 */
    float k1[4], k2[4], k3[4], k4[4], u[4], ns[4];
    wd_acrobot_dsdt(s, torque, k1);
    for (int i = 0; i < 4; i++) u[i] = (float)((double)s[i] + (double)k1[i] * dt2);
    wd_acrobot_dsdt(u, torque, k2);
    for (int i = 0; i < 4; i++) u[i] = (float)((double)s[i] + (double)k2[i] * dt2);
    wd_acrobot_dsdt(u, torque, k3);
    for (int i = 0; i < 4; i++) u[i] = (float)((double)s[i] + (double)k3[i] * dt);
    wd_acrobot_dsdt(u, torque, k4);
    for (int i = 0; i < 4; i++) {
      double sum = (((double)k1[i] + 2.0 * (double)k2[i]) + 2.0 * (double)k3[i]) + (double)k4[i];
      ns[i] = (float)((double)s[i] + (dt / 6.0) * sum);
    }
    ns[0] = (float)wd_wrap((double)ns[0], -kPi, kPi);                        /* :51-54 */
    ns[1] = (float)wd_wrap((double)ns[1], -kPi, kPi);
    ns[2] = (float)fmin(fmax((double)ns[2], -kMaxVel1), kMaxVel1);
    ns[3] = (float)fmin(fmax((double)ns[3], -kMaxVel2), kMaxVel2);
    for (int i = 0; i < 4; i++) s[i] = ns[i];
    float c0 = cosf(ns[0]);
    int terminated = (float)(-c0 - cosf((float)(ns[1] + ns[0]))) > 1.0f;     /* :151-153 */
    reward[env] = terminated ? 0.0f : -1.0f;                                 /* :44,60-61 */
    float *o = obs + (size_t)env * 6;                                        /* :156-168 */
    o[0] = c0; o[1] = sinf(ns[0]); o[2] = cosf(ns[1]); o[3] = sinf(ns[1]);
    o[4] = ns[2]; o[5] = ns[3];
    if (env_timestep[env] == episode_length || terminated) done[env] = 1;    /* :66-67 */
  }
}

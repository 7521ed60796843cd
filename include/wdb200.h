/*
 * wdb200.h -- C ABI of libwdb200.so, the B200-native (sm_100a) rollout kernels that
 * replace WarpDrive's JIT-compiled CUDA module.
 *
 * Boundary being replaced (reference = salesforce/warp-drive v2.7.1): the reference
 * looks kernels up BY NAME in a pycuda module and launches them with positional
 * arguments (warp_drive/managers/pycuda_managers/pycuda_function_manager.py:319-388,
 * `get_function(name)(*args, block=, grid=)`).  Each entry point below takes the same
 * arguments, in the same order, as the reference kernel it replaces, with three
 * changes: (1) a leading `stream` (cudaStream_t as void*; NULL = legacy default
 * stream), (2) the compile-time constants wkNumberEnvs / wkNumberAgents /
 * wkBlocksPerEnv (warp_drive/cuda_includes/template_env_config.h:19-21) become run-time
 * arguments, (3) `bool` kernel parameters become int32 (the reference passes np.int32
 * into `bool` params, SURVEY.md section 8 "Python-side arg passing").
 *
 * Conventions
 *   - every function returns a cudaError_t as int (0 == cudaSuccess) and never throws;
 *     wdb_error_string() turns it into text.  Launches are asynchronous on `stream`.
 *   - all buffers are caller-owned device memory (torch CUDA tensors), row-major,
 *     C-contiguous, 32-bit elements.  State arrays are updated in place
 *     (warp_drive/env_wrapper.py:347-352).
 *   - the library allocates nothing; RNG state is a caller-provided blob of
 *     wdb_rng_state_bytes(n_streams) bytes.
 *   - one host thread per process, one process per GPU.
 */
#ifndef WDB200_H_
#define WDB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WDB_ABI_VERSION 1

int wdb_abi_version(void);
const char *wdb_error_string(int err);
/* number of kernel launches issued through this library since load (bench.py's
 * `gpu_launches` evidence) */
long long wdb_launch_count(void);

/* Tuning / A-B switches (results are identical either way).  "tc_history": use last step's
 * neighbour lists as a distance threshold in the tag_continuous k-nearest search (1, default)
 * or always run the full sorting network (0).  "tc_force_exact": resolve every agent's
 * neighbours with the reference's literal selection (1) instead of only the ambiguous ones
 * (0, default).  "tc_cta_threads": thread budget of one
 * tag_continuous CTA (32..320, default 320); a CTA takes floor(budget / agents) envs.
 * "mlp_max_ctas": persistent CTAs (= SMs) the following wdb_mlp_policy_forward launches may
 * use (0 = all): two policies' forwards can then run concurrently on disjoint SMs. */
int wdb_set_option(const char *name, int value);

/* ------------------------------------------------------------------ RNG ------ */
/* Replaces init_random / free_random (warp_drive/cuda_includes/core/random.cu:14-31):
 * the reference heap-allocates one XORWOW curandState per (env, agent) thread; here the
 * state is a flat caller-owned blob: {u64 seed, u64 n_streams, u64 offset[n_streams]}
 * driving a counter-based Philox4x32-10 (key = seed, counter = {offset, stream}). */
long long wdb_rng_state_bytes(long long n_streams);
int wdb_rng_init(void *stream, void *rng_state, long long n_streams,
                 unsigned long long seed);

/* Test hook: out[4*i .. 4*i+3] = the next Philox block of stream i (advances every
 * stream's offset by one) -- lets tests pin the device RNG against the scalar oracle. */
int wdb_rng_draw_u32x4(void *stream, void *rng_state, unsigned int *out,
                       long long n_streams);

/* Replaces sample_actions (random.cu:51-85; launched from PyCUDASampler.sample,
 * pycuda_function_manager.py:532-572).  probs [n_envs, n_agents, n_actions] f32.
 * actions [n_envs, n_agents, 1] i32.  cum_distr [n_envs, n_agents, n_actions] f32 or
 * NULL (the reference always materialises it; it is scratch).  actions_combined
 * (optional) receives the same index at [pos * combined_stride + combined_offset] --
 * the reference does that with a separate torch copy (trainer_base.py:507-512).
 * uniforms (optional, test hook): if non-NULL the draw p is read from uniforms[pos]
 * instead of the RNG, which makes the index selection parity-testable. */
int wdb_sample_actions(void *stream, void *rng_state, const float *probs, int *actions,
                       float *cum_distr, int n_envs, int n_agents, int n_actions,
                       int use_argmax, int *actions_combined, int combined_stride,
                       int combined_offset, const float *uniforms);

/* Replaces the numba-only sample_ou_process (warp_drive/numba_includes/core/random.py:
 * 74-105): ou = (1-damping)*ou + stddev*N(0,1); action = mean + scale*ou.
 * normals (optional, test hook) replaces the N(0,1) draw. */
int wdb_sample_ou_process(void *stream, void *rng_state, const float *mean,
                          float *actions, float *ou_state, int n_envs, int n_agents,
                          float damping, float stddev, float scale,
                          const float *normals);

/* ------------------------------------------------------------------ reset ---- */
/* Replaces reset_in_{float,int}_when_done_{2d,3d} + undo_done_flag_and_reset_timestep
 * (warp_drive/cuda_includes/core/reset.cu:9-75; one launch per registered array in
 * PyCUDAEnvironmentReset.reset_when_done_deterministic, pycuda_function_manager.py:
 * 668-753) and the numba-only reset_when_done_{1,2,3}d_from_pool
 * (numba_includes/core/pool_reset.py:15-52) with ONE launch over a device-resident
 * descriptor table. */
typedef struct wdb_reset_desc {
  void *dst;                /* data[env]            */
  const void *ref;          /* data_at_reset[env] or pool[row]                      */
  long long bytes_per_env;  /* multiple of 4                                        */
  long long pool_rows;      /* 0: deterministic (ref indexed by env); >0: pool size */
} wdb_reset_desc;

int wdb_reset_when_done(void *stream, const wdb_reset_desc *table_dev, int n_arrays,
                        int *done, int *timestep, int n_envs, int force_reset,
                        int undo_done_and_timestep, void *pool_rng_state);

/* ------------------------------------------------------------------ log ------ */
/* Replace reset_log_mask / update_log_mask / log_one_step_in_{float,int}
 * (warp_drive/cuda_includes/core/log.cu:11-62). */
int wdb_reset_log_mask(void *stream, int *log_mask, int episode_length);
int wdb_update_log_mask(void *stream, int *log_mask, int timestep, int episode_length);
int wdb_log_one_step(void *stream, void *log, const void *data, int n_agents,
                     int feature_dim, int timestep, int episode_length, int env_id);

/* ------------------------------------------------------------------ envs ----- */
/* Replaces testkernel (example_envs/dummy_env/test_step.cu:9-45). */
int wdb_testkernel(void *stream, int n_envs, int n_agents, float *x, int *y, int *done,
                   int *actions, float multiplier, int target, int step,
                   int episode_length);

/* Replaces CudaTagGridWorldStep (example_envs/tag_gridworld/
 * tag_gridworld_step_pycuda.cu:112-251; argument order of tag_gridworld.py:353-368).
 * index_to_action: device int[10], the reference's __constant__ kIndexToActionArr. */
int wdb_tag_gridworld_step(void *stream, int n_envs, int n_agents, int *loc_x,
                           int *loc_y, const int *actions, int *done, float *rewards,
                           float *obs, float wall_hit_penalty,
                           float tag_reward_for_tagger, float tag_penalty_for_runner,
                           float step_cost_for_tagger, int use_full_observation,
                           int world_boundary, int *env_timestep, int episode_length,
                           const int *index_to_action);

/* Replaces CudaTagContinuousStep (example_envs/tag_continuous/
 * tag_continuous_step_pycuda.cu:351-520; argument order of tag_continuous.py:806-840).
 * neighbor_distances / neighbor_ids_sorted_by_distance are the reference's O(N^2)
 * global scratch; they may be NULL (the B200 kernel keeps the sweep on chip).
 * stats (optional, device int[4]): [0] += agents that took the exact tie-resolution
 * path, [1] += tags, [2] += agents whose threshold scan had no usable candidate list
 * (full sorting network), [3] += agents whose candidate list overflowed (cluster kernel:
 * sorting network over the x-window only). */
int wdb_tag_continuous_step(
    void *stream, int n_envs, int n_agents, int blocks_per_env, float *loc_x,
    float *loc_y, float *speed, float *direction, float *acceleration,
    const int *agent_types, float *edge_hit_reward_penalty, float edge_hit_penalty,
    float grid_length, const float *acceleration_actions, const float *turn_actions,
    float max_speed, int num_other_agents_observed, const float *skill_levels,
    int runner_exits_game_after_tagged, int *still_in_the_game,
    int use_full_observation, float *obs, const int *action_indices,
    float *neighbor_distances, int *neighbor_ids_sorted_by_distance,
    int *nearest_neighbor_ids, float *rewards, const float *step_rewards,
    int *num_runners, float distance_margin_for_reward, float tag_reward_for_tagger,
    float tag_penalty_for_runner, float end_of_game_reward_for_runner, int *done,
    int *env_timestep, int episode_length, int *stats);

/* ------------------------------------------------------------ fused rollout step -- */
/* ONE launch per rollout timestep of tag_continuous:
 *   sample both action heads from the policies' probabilities (sample_actions x2,
 *   core/random.cu:51-85) -> CudaTagContinuousStep -> push actions / rewards / done to the
 *   training batch slots and update the episodic sums (trainer_base.py:437-601) ->
 *   write the new observations per policy for the next forward pass
 *   (model_base.py:181-200) -> done-masked reset of every registered array + undo of
 *   done/timestep (core/reset.cu:9-75; function_manager.py:256-273).
 * The reference issues ~50 launches and >= 5 host syncs for the same work. */
typedef struct wdb_tc_env {          /* the arguments of wdb_tag_continuous_step */
  int n_envs, n_agents;
  float *loc_x, *loc_y, *speed, *direction, *acceleration;
  const int *agent_types;
  float *edge_hit_reward_penalty;
  float edge_hit_penalty, grid_length;
  const float *acceleration_actions, *turn_actions;
  float max_speed;
  int num_other_agents_observed;
  const float *skill_levels;
  int runner_exits_game_after_tagged;
  int *still_in_the_game;
  int use_full_observation;
  float *obs;                        /* [E, N, F]; may be NULL when obs_next is used */
  float *neighbor_distances;         /* optional scratch */
  int *neighbor_ids_sorted_by_distance;
  int *nearest_neighbor_ids;
  float *rewards;
  const float *step_rewards;
  int *num_runners;
  float distance_margin_for_reward, tag_reward_for_tagger, tag_penalty_for_runner,
      end_of_game_reward_for_runner;
  int *done, *env_timestep;
  int episode_length;
  int *stats;
  int blocks_per_env;                /* <= 1: env replicas packed into CTAs; > 1: one env per
                                      * cluster of that many CTAs (as wdb_tag_continuous_step) */
} wdb_tc_env;

typedef struct wdb_tc_policy_io {    /* per policy; agents of a policy are [E, Np, ...] */
  int n_agents;                      /* Np */
  const float *probs0, *probs1;      /* [E, Np, A0], [E, Np, A1] from the forward pass */
  int *actions_batch;                /* [E, Np, 2]  slot t of sampled_actions_batch_p, or NULL */
  float *rewards_batch;              /* [E, Np]     slot t of rewards_batch_p, or NULL */
  float *obs_next;                   /* [E, Np, F]  where the next forward reads, or NULL */
  float *reward_running_sum;         /* [E, Np] or NULL */
  float *episodic_reward_sum;        /* scalar or NULL */
  void *obs_next_tiles;              /* optional bf16 copy of obs_next in the A-operand tile
                                      * layout of wdb_mlp_policy_forward_tiles
                                      * (wdb_mlp_obs_tiles_bytes bytes, 16-byte aligned), or NULL */
} wdb_tc_policy_io;

typedef struct wdb_tc_rollout {
  void *rng_state;
  const float *uniforms;             /* test hook [E, N, 2]; NULL = device RNG */
  int n_policies, n_actions0, n_actions1;
  const int *agent_policy;           /* [N] policy index of every agent */
  const int *agent_slot;             /* [N] index of the agent inside its policy */
  wdb_tc_policy_io policy[4];
  int *sampled_actions;              /* [E, N, 2] or NULL */
  int *sampled_actions_0, *sampled_actions_1;   /* [E, N, 1] or NULL */
  int *done_batch;                   /* [E] slot t of done_flags_batch, or NULL */
  int *step_running_sum;             /* [E] or NULL */
  unsigned long long *episodic_step_sum, *num_completed_episodes;   /* scalars or NULL */
  const wdb_reset_desc *reset_table; /* device table of the arrays to restore */
  int n_reset_arrays;
  const float *obs_at_reset;         /* [E, N, F]: source of obs_next for envs that reset */
  int reset_done_envs;               /* 0: leave done envs alone (done stays set) */
  int launch_after_forward;          /* 1: the kernel launched just before this one on `stream`
                                      * is this step's policy forward (and the one before that
                                      * the previous env step): with the "pdl" option on, the
                                      * launch is programmatically dependent on it -- the state
                                      * prologue overlaps the forward's tail, the probabilities
                                      * are read after griddepcontrol.wait.  0: plain launch. */
} wdb_tc_rollout;

int wdb_tag_continuous_rollout_step(void *stream, const wdb_tc_env *env,
                                    const wdb_tc_rollout *rollout);

/* Replaces NumbaClassicControlCartPoleEnvStep (example_envs/single_agent/
 * classic_control/cartpole/cartpole_step_numba.py:6-83; argument order of
 * cartpole.py:105-122). */
int wdb_cartpole_step(void *stream, int n_envs, float *state, const int *action,
                      int *done, float *reward, float *obs, float gravity,
                      float masspole, float total_mass, float length,
                      float polemass_length, float force_mag, float tau,
                      float theta_threshold_radians, float x_threshold,
                      int *env_timestep, int episode_length);

/* ------------------------------------------------- classic control ---- */
/* The four remaining single-agent envs of example_envs/single_agent/classic_control/.
 * The reference has them only as numba kernels launched as one 1-thread block per env;
 * argument order below = the kernel signatures (and the `args` lists of the env classes'
 * step()), with n_envs made explicit.  state/obs rows are [n_envs, 1, k] float32. */

/* Replaces NumbaClassicControlMountainCarEnvStep (mountain_car/mountain_car_step_numba.py:
 * 14-70; args of mountain_car.py:104-119).  state/obs k = 2, action int32 [E,1,1].
 * done = 1 at the episode end, 2 when the goal is reached (as the reference). */
int wdb_mountain_car_step(void *stream, int n_envs, float *state, const int *action,
                          int *done, float *reward, float *obs, float min_position,
                          float max_position, float max_speed, float goal_position,
                          float goal_velocity, float force, float gravity,
                          int *env_timestep, int episode_length);

/* Replaces NumbaClassicControlContinuousMountainCarEnvStep (continuous_mountain_car/
 * continuous_mountain_car_step_numba.py:14-71; args of continuous_mountain_car.py:105-121).
 * action float32 [E,1,1]. */
int wdb_continuous_mountain_car_step(void *stream, int n_envs, float *state,
                                     const float *action, int *done, float *reward,
                                     float *obs, float min_action, float max_action,
                                     float min_position, float max_position, float max_speed,
                                     float goal_position, float goal_velocity, float power,
                                     int *env_timestep, int episode_length);

/* Replaces NumbaClassicControlPendulumEnvStep (pendulum/pendulum_step_numba.py:30-72; args
 * of pendulum.py:92-100).  state k = 2 (theta, theta_dot), obs k = 3, action float32. */
int wdb_pendulum_step(void *stream, int n_envs, float *state, const float *action, int *done,
                      float *reward, float *obs, int *env_timestep, int episode_length);

/* Replaces NumbaClassicControlAcrobotEnvStep (acrobot/acrobot_step_numba.py:24-168; args of
 * acrobot.py:91-99).  state k = 4 (16-byte aligned), obs k = 6, action int32 in {0,1,2}. */
int wdb_acrobot_step(void *stream, int n_envs, float *state, const int *action, int *done,
                     float *reward, float *obs, int *env_timestep, int episode_length);

/* ------------------------------------------ whole rollout, single-agent envs ---- */
/* ONE launch = n_steps rollout timesteps of a discrete-action single-agent env, one thread
 * per env replica: FullyConnected forward (fp32, weights in shared memory) -> categorical
 * sample -> step -> bookkeeping -> done-masked reset -> push to the batch slots.  Replaces
 * the body of TrainerBase._generate_rollout_batch (trainer_base.py:383-428) for
 * ClassicControl{CartPole,MountainCar,Acrobot}Env: per timestep the reference issues the
 * model forward, sample_actions, Numba...EnvStep, ~20 bookkeeping torch ops and one reset
 * kernel per registered array.  The step physics are the same compiled functions the
 * stand-alone step kernels call. */
enum { WDB_SA_CARTPOLE = 0, WDB_SA_MOUNTAIN_CAR = 1, WDB_SA_ACROBOT = 2 };

typedef struct wdb_sa_rollout {
  int env_kind, n_envs, n_steps, episode_length, state_dim, use_argmax;
  float env_params[12];              /* the scalar arguments of the env's step function, in
                                      * the order of wdb_cartpole_step / wdb_mountain_car_step */
  float *state;                      /* [E, state_dim] */
  float *observations;               /* [E, F] current observation (read and updated) */
  int *done, *env_timestep;          /* [E] */
  float *rewards;                    /* [E] */
  int *sampled_actions;              /* [E] */
  int n_hidden;                      /* hidden Linear + ReLU layers (0..3), then the softmax head */
  int dims[5];                       /* F, H1, ..., A */
  const float *w[4], *b[4];          /* nn.Linear weights [out, in] and biases, fp32 */
  void *rng_state;
  const float *uniforms;             /* test hook [n_steps, E]; NULL = device RNG */
  float *obs_batch;                  /* [n_steps, E, F] or NULL */
  int *actions_batch;                /* [n_steps, E] or NULL */
  float *rewards_batch;              /* [n_steps, E] or NULL */
  int *done_batch;                   /* [n_steps, E] or NULL */
  float *probs_batch;                /* [n_steps, E, A] or NULL (tests) */
  float *reward_running_sum;         /* [E] or NULL */
  int *step_running_sum;             /* [E] or NULL */
  float *episodic_reward_sum;        /* scalar or NULL */
  unsigned long long *episodic_step_sum, *num_completed;   /* scalars or NULL */
  const wdb_reset_desc *reset_table; /* device table of the arrays to restore when done */
  int n_reset_arrays;
  void *pool_rng;                    /* RNG state of the reset pools, or NULL */
  int reset_done_envs;
} wdb_sa_rollout;

int wdb_single_agent_rollout_supported(int n_hidden, const int *dims);
int wdb_single_agent_rollout(void *stream, const wdb_sa_rollout *rollout);

/* -------------------------------------------------- rollout of ANY env: data movement ---- */
/* [E, N, W] array of 4-byte elements <-> one dense [E, Np, W] block per policy, every policy
 * in one launch.  scatter = 0: rows[e, j, :] = full[e, agent_ids[j], :] (the per-policy
 * observation push of warp_drive/training/trainer_base.py:437-464); scatter = 1: the reverse
 * (the per-policy probabilities into the [E, N, A] array the sampler reads, :466-512).
 * agent_ids == NULL: the policy covers agents 0..Np-1 in order. */
typedef struct wdb_gather_policy {
  int n_agents;                      /* Np */
  const int *agent_ids;              /* device [Np] or NULL */
  void *rows;                        /* [E, Np, W] */
} wdb_gather_policy;
typedef struct wdb_gather {
  int n_envs, n_agents, width, n_policies, scatter;
  void *full;                        /* [E, N, W] */
  wdb_gather_policy policy[4];
} wdb_gather;
int wdb_gather_policy_rows(void *stream, const wdb_gather *gather);

/* Per-timestep bookkeeping of the rollout (trainer_base.py:514-601, which uses
 * done_flags.nonzero() and len() on the host): done -> done_batch (raw flag); per policy the
 * agents' rewards -> rewards_batch, sampled actions -> actions_batch, reward_running_sum += r,
 * and for done envs episodic_reward_sum += the running sums, which are then cleared;
 * step_running_sum += 1, for done envs episodic_step_sum += it, cleared, num_completed += 1.
 * Any pointer may be NULL (that piece is skipped); ONE launch, no host synchronisation. */
typedef struct wdb_bookkeep_policy {
  int n_agents;
  const int *agent_ids;              /* device [Np] or NULL (agents 0..Np-1) */
  float *rewards_batch;              /* [E, Np] slot t */
  int *actions_batch;                /* [E, Np, n_heads] slot t */
  float *reward_running_sum;         /* [E, Np] */
  float *episodic_reward_sum;        /* scalar */
} wdb_bookkeep_policy;
typedef struct wdb_bookkeep {
  int n_envs, n_agents, n_policies, n_heads;
  const int *done;                   /* [E] */
  const float *rewards;              /* [E, N] */
  const int *actions;                /* [E, N, n_heads] or NULL */
  int *done_batch;                   /* [E] slot t */
  int *step_running_sum;             /* [E] */
  unsigned long long *episodic_step_sum, *num_completed_episodes;
  wdb_bookkeep_policy policy[4];
} wdb_bookkeep;
int wdb_rollout_bookkeep(void *stream, const wdb_bookkeep *bookkeep);

/* ------------------------------------------------------------- policy forward ---- */
/* Fused policy/value MLP forward on the tensor cores (tcgen05 + TMEM), replacing the
 * rollout-time FullyConnected.forward (warp_drive/training/models/fully_connected.py:51-89):
 *   obs [rows, F] -> Linear(F,H)+ReLU -> Linear(H,H)+ReLU -> softmax(Linear(H,A0)),
 *   softmax(Linear(H,A1)), Linear(H,1).
 * Weights are nn.Linear tensors ([out, in] fp32); wdb_mlp_pack_weights converts them to bf16
 * tiles inside a caller-owned blob of wdb_mlp_blob_bytes() bytes (re-pack after every
 * optimizer step).  H must be a multiple of 32 <= 256, A0 + A1 + 1 <= 64, F <= 256;
 * wdb_mlp_blob_bytes returns -1 for unsupported shapes. */
long long wdb_mlp_blob_bytes(int F, int H, int A0, int A1);
int wdb_mlp_pack_weights(void *stream, void *blob, const float *w1, const float *b1,
                         const float *w2, const float *b2, const float *wh0, const float *bh0,
                         const float *wh1, const float *bh1, const float *wv, const float *bv,
                         int F, int H, int A0, int A1);
int wdb_mlp_policy_forward(void *stream, const void *blob, int F, int H, int A0, int A1,
                           const float *obs, long long rows, float *probs0, float *probs1,
                           float *values /* may be NULL */);
/* Two policies in ONE launch (the rollout of a two-policy env: tag_continuous runners and
 * taggers): CTAs [0, grid - ctas_b) run policy 0, the last ctas_b CTAs policy 1, each
 * persistent over its own 128-row tiles.  ctas_b = 0 picks the split from the row counts.
 * flags: WDB_MLP_WEIGHTS_STABLE = neither blob was written by the kernel launched just before
 * this one on `stream` (true for every rollout step but the first after a weight re-pack): with
 * the "pdl" option on, the launch is a programmatic dependent launch and the weight load then
 * overlaps the tail of that previous kernel. */
#define WDB_MLP_WEIGHTS_STABLE 1
typedef struct wdb_mlp_pair {
  const void *blob[2];
  int F[2], H[2], A0[2], A1[2];
  const float *obs[2];
  long long rows[2];
  float *probs0[2], *probs1[2], *values[2];   /* values may be NULL */
  int ctas_b;
  int flags;
} wdb_mlp_pair;
int wdb_mlp_policy_forward_pair(void *stream, const wdb_mlp_pair *pair);

/* The same forward fed from the bf16 copy of the observations in A-operand layout: 128-row
 * tiles, each one contiguous block (canonical K-major layout, K padded to a multiple of 16
 * with zeros) that the kernel fetches with one TMA bulk copy.  wdb_mlp_pack_obs builds it
 * from fp32 obs [rows, F]; the fused tag_continuous step writes it directly
 * (wdb_tc_policy_io.obs_next_tiles).  No reference counterpart: the reference always runs
 * the torch forward on the fp32 batch slice (trainer_base.py:437-464). */
long long wdb_mlp_obs_tiles_bytes(int F, long long rows);
int wdb_mlp_pack_obs(void *stream, const float *obs, long long rows, int F, void *tiles);
int wdb_mlp_policy_forward_tiles(void *stream, const void *blob, int F, int H, int A0, int A1,
                                 const void *obs_tiles, long long rows, float *probs0,
                                 float *probs1, float *values /* may be NULL */);

/* ------------------------------------------------------------------ update ---- */
/* Bootstrapped discounted returns of the A2C/PPO update, backwards in time with done
 * masking (warp_drive/training/algorithms/policygradient/a2c.py:80-93):
 *   returns[T-1] = done ? r : V ;  returns[t] = r[t] + (done[t] ? 0 : gamma*returns[t+1]).
 * rewards/values/returns [T, E, Np] f32, done [T, E] i32. */
int wdb_discounted_returns(void *stream, const float *rewards, const int *done,
                           const float *values, float *returns, int T, int n_envs,
                           int n_agents, float gamma);

/* Elementwise pieces of the MLP's TRAINING forward / backward (warp_drive/training/models/
 * fully_connected.py:51-89 under autograd); the GEMMs between them stay on cuBLAS.
 * wdb_heads_softmax: logits [rows, pitch >= A0 + A1 + 1] (the pitch is padded to a multiple of
 * 4 floats so that cuBLAS can use its 16-byte-aligned tensor-core kernels; wdb_heads_softmax_backward
 * zero-fills the padding columns) of the two action heads and the value head
 * (one GEMM) -> softmax per head into probs0 [rows, A0], probs1 [rows, A1] (A1 may be 0) and
 * the value column into values [rows].  wdb_heads_softmax_backward: the gradient with respect
 * to those logits from the gradients of probs / values (NULL = no gradient).
 * wdb_relu_backward_bias: grad_hidden[r, c] = hidden[r, c] > 0 ? grad_hidden[r, c] : 0 in
 * place, and per CTA the column sums of the result into partial_bias_grads
 * [ceil(rows / wdb_relu_backward_bias_rows(rows)), width] (the bias gradient = their sum;
 * no atomics: deterministic). */
int wdb_heads_softmax(void *stream, const float *logits, long long rows, int A0, int A1,
                      int pitch, float *probs0, float *probs1, float *values);
int wdb_heads_softmax_backward(void *stream, const float *probs0, const float *probs1,
                               const float *grad_probs0, const float *grad_probs1,
                               const float *grad_values, long long rows, int A0, int A1,
                               int pitch, float *grad_logits);
int wdb_pad_rows(void *stream, const float *src, long long rows, int width, int pitch,
                 float *dst /* [rows, pitch], columns >= width zero-filled */);
int wdb_relu_backward_bias_rows(long long rows);
int wdb_relu_backward_bias(void *stream, float *grad_hidden, const float *hidden,
                           long long rows, int width, float *partial_bias_grads);

/* Fused A2C / PPO loss (a2c.py:80-130, ppo.py:82-141): ONE backward-in-time scan per (env,
 * agent) computes the bootstrapped returns, advantages, Categorical log-prob and entropy of
 * every head, the loss sums and the gradients of
 *     loss = mean(-logp * adv) + vf_coeff * mean((V - R)^2) - entropy_coeff * sum_k mean(H_k)
 * with respect to the probabilities and the values (PPO's clipped surrogate has the same
 * gradient at ratio == 1, the reference's single-epoch case, ppo.py:127-136).
 * sums (double[4], zeroed by the caller): sum(-logp*adv), sum((V-R)^2), sum_k sum(H_k), sum(adv). */
typedef struct wdb_pg_loss {
  int T, n_envs, n_agents, n_heads;
  int n_actions[4];
  const float *probs[4];             /* [T, E, Np, A_k] */
  const float *values;               /* [T, E, Np] */
  const int *actions;                /* [T, E, Np, n_heads] */
  const float *rewards;              /* [T, E, Np] */
  const int *done;                   /* [T, E] (non-zero = done) */
  float gamma, vf_coeff, entropy_coeff;
  float *grad_probs[4];              /* out, same shapes as probs (or NULL) */
  float *grad_values;                /* out (or NULL) */
  float *returns;                    /* out [T, E, Np] (or NULL) */
  double *sums;                      /* out [4], accumulated atomically */
} wdb_pg_loss;
int wdb_pg_loss_and_grads(void *stream, const wdb_pg_loss *loss);

/* Gradient-norm clipping + Adam over one flat float32 arena (torch.optim.Adam semantics, no
 * weight decay; torch.nn.utils.clip_grad_norm_ semantics for the clip factor, which the Adam
 * kernel reads from device memory: no host synchronisation).  Replaces the per-tensor
 * optimizer loop of trainer_a2c.py:300-339. */
int wdb_grad_sumsq(void *stream, const float *grads, long long n, double *out /* += */);
int wdb_adam_step(void *stream, float *params, float *grads, float *exp_avg, float *exp_avg_sq,
                  long long n, float lr, float beta1, float beta2, float eps, int step,
                  float max_grad_norm /* <= 0: no clipping */, const double *grad_sumsq);

#ifdef __cplusplus
}
#endif
#endif /* WDB200_H_ */

// wdb_rollout_generic.cu -- the per-timestep data movement and bookkeeping of the rollout for
// ANY env (tag_gridworld, custom envs, envs with reset pools ...), i.e. for everything that
// does not have a fused step kernel of its own.
//
// The reference does this with torch indexing per policy and per head
// (warp_drive/training/trainer_base.py:437-512 obs / action push, :514-601 reward / done push
// and the episodic sums with done_flags.nonzero() / len()): 30-40 small launches and two
// host synchronisations per timestep.  Here:
//   wdb_gather_policy_rows   [E, N, W] <-> per-policy [E, Np, W] rows, every policy in ONE
//                            launch, both directions (obs -> batch slot of each policy;
//                            per-policy probabilities -> the sampler's [E, N, A] array)
//   wdb_rollout_bookkeep     done -> batch; per policy rewards (gather) -> batch, actions
//                            (gather) -> batch, running / episodic reward sums; step sums,
//                            completed-episode count: ONE launch, no host synchronisation
#include "wdb_common.cuh"

using namespace wdb;

namespace {

constexpr int kGenThreads = 128;

struct GatherArgs {
  int n_envs, n_agents, width, n_policies, scatter;
  uint32_t *full;
  int np[4];
  const int *ids[4];
  uint32_t *rows[4];
  long long first_item[5];     // prefix over policies of E * Np (work items = rows)
};

// one warp per (policy, env, slot) row; lanes over the row's 4-byte elements
__global__ void __launch_bounds__(kGenThreads)
gather_rows_kernel(const __grid_constant__ GatherArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= a.first_item[a.n_policies]) return;
  int p = 0;
#pragma unroll
  for (int q = 1; q < 4; q++)
    if (q < a.n_policies && warp >= a.first_item[q]) p = q;
  const long long r = warp - a.first_item[p];
  const int np = a.np[p];
  const int env = (int)(r / np), slot = (int)(r - (long long)env * np);
  const int agent = a.ids[p] ? a.ids[p][slot] : slot;
  uint32_t *f = a.full + ((long long)env * a.n_agents + agent) * a.width;
  uint32_t *q = a.rows[p] + r * a.width;
  if (a.scatter) {
    for (int i = lane; i < a.width; i += 32) f[i] = q[i];
  } else {
    for (int i = lane; i < a.width; i += 32) q[i] = f[i];
  }
}

struct BookArgs {
  int n_envs, n_agents, n_policies, n_heads;
  const int *done;
  const float *rewards;
  const int *actions;
  int *done_batch;
  int *step_running_sum;
  unsigned long long *episodic_step_sum, *num_completed;
  int np[4];
  const int *ids[4];
  float *rewards_batch[4];
  int *actions_batch[4];
  float *reward_running_sum[4];
  float *episodic_reward_sum[4];
};

// one CTA per env
__global__ void __launch_bounds__(kGenThreads)
bookkeep_kernel(const __grid_constant__ BookArgs b) {
  const int env = blockIdx.x, tid = threadIdx.x;
  const int raw = b.done[env];
  const bool d = raw > 0;                     // any non-zero flag ends the episode
  __shared__ float red[kGenThreads / 32];
  for (int p = 0; p < b.n_policies; p++) {
    const int np = b.np[p];
    float part = 0.0f;
    for (int slot = tid; slot < np; slot += blockDim.x) {
      const int agent = b.ids[p] ? b.ids[p][slot] : slot;
      const long long src = (long long)env * b.n_agents + agent;
      const long long dst = (long long)env * np + slot;
      const float r = b.rewards[src];
      if (b.rewards_batch[p]) b.rewards_batch[p][dst] = r;
      if (b.actions && b.actions_batch[p])
        for (int k = 0; k < b.n_heads; k++)
          b.actions_batch[p][dst * b.n_heads + k] = b.actions[src * b.n_heads + k];
      if (b.reward_running_sum[p]) {
        const float run = b.reward_running_sum[p][dst] + r;
        b.reward_running_sum[p][dst] = d ? 0.0f : run;
        if (d) part += run;
      }
    }
    if (d && b.episodic_reward_sum[p] && b.reward_running_sum[p]) {   // CTA-uniform branch
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
      if ((tid & 31) == 0) red[tid >> 5] = part;
      __syncthreads();
      if (tid == 0) {
        float s = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
        atomicAdd(b.episodic_reward_sum[p], s);
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    if (b.done_batch) b.done_batch[env] = raw;    // the batch keeps the raw flag (MountainCar: 2)
    if (b.step_running_sum) {
      const int steps = b.step_running_sum[env] + 1;
      b.step_running_sum[env] = d ? 0 : steps;
      if (d && b.episodic_step_sum) atomicAdd(b.episodic_step_sum, (unsigned long long)steps);
    }
    if (d && b.num_completed) atomicAdd(b.num_completed, 1ull);
  }
}

}  // namespace

WDB_API int wdb_gather_policy_rows(void *stream, const wdb_gather *g) {
  if (!g || !g->full || g->n_envs < 1 || g->n_agents < 1 || g->width < 1 ||
      g->n_policies < 1 || g->n_policies > 4)
    return (int)cudaErrorInvalidValue;
  GatherArgs a = {};
  a.n_envs = g->n_envs; a.n_agents = g->n_agents; a.width = g->width;
  a.n_policies = g->n_policies; a.scatter = g->scatter ? 1 : 0;
  a.full = reinterpret_cast<uint32_t *>(g->full);
  long long items = 0;
  for (int p = 0; p < g->n_policies; p++) {
    if (!g->policy[p].rows || g->policy[p].n_agents < 1 || g->policy[p].n_agents > g->n_agents)
      return (int)cudaErrorInvalidValue;
    a.np[p] = g->policy[p].n_agents;
    a.ids[p] = g->policy[p].agent_ids;
    a.rows[p] = reinterpret_cast<uint32_t *>(g->policy[p].rows);
    a.first_item[p] = items;
    items += (long long)g->n_envs * a.np[p];
  }
  for (int p = g->n_policies; p <= 4; p++) a.first_item[p] = items;
  const long long threads = items * 32;
  gather_rows_kernel<<<(unsigned)((threads + kGenThreads - 1) / kGenThreads), kGenThreads, 0,
                       as_stream(stream)>>>(a);
  return finish_launch();
}

WDB_API int wdb_rollout_bookkeep(void *stream, const wdb_bookkeep *k) {
  if (!k || !k->done || !k->rewards || k->n_envs < 1 || k->n_agents < 1 ||
      k->n_policies < 1 || k->n_policies > 4 || k->n_heads < 0)
    return (int)cudaErrorInvalidValue;
  BookArgs b = {};
  b.n_envs = k->n_envs; b.n_agents = k->n_agents; b.n_policies = k->n_policies;
  b.n_heads = k->n_heads;
  b.done = k->done; b.rewards = k->rewards; b.actions = k->actions;
  b.done_batch = k->done_batch; b.step_running_sum = k->step_running_sum;
  b.episodic_step_sum = k->episodic_step_sum; b.num_completed = k->num_completed_episodes;
  for (int p = 0; p < k->n_policies; p++) {
    const wdb_bookkeep_policy &io = k->policy[p];
    if (io.n_agents < 1 || io.n_agents > k->n_agents) return (int)cudaErrorInvalidValue;
    if (io.actions_batch && (!k->actions || k->n_heads < 1)) return (int)cudaErrorInvalidValue;
    b.np[p] = io.n_agents; b.ids[p] = io.agent_ids;
    b.rewards_batch[p] = io.rewards_batch; b.actions_batch[p] = io.actions_batch;
    b.reward_running_sum[p] = io.reward_running_sum;
    b.episodic_reward_sum[p] = io.episodic_reward_sum;
  }
  bookkeep_kernel<<<k->n_envs, kGenThreads, 0, as_stream(stream)>>>(b);
  return finish_launch();
}

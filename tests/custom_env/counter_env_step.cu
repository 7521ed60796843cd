// A small custom env written the way the reference's docs tell developers to write one
// (tutorial "create custom environments": extern "C" kernels, the compile-time constants
// wkNumberEnvs / wkNumberAgents / wkBlocksPerEnv, blockIdx.x / wkBlocksPerEnv = env id,
// threadIdx.x (+ block offset) = agent id, a __constant__ lookup table filled through
// initialize_shared_constants).  Used by tests/test_custom_env*.py: it must compile for
// sm_100a and run through EnvWrapper without any change to this file.
//
// Dynamics (integer, so the test is bit-exact): every agent owns a counter; action a adds
// kStepTable[a]; reward = new counter value as float; obs = [own counter, env sum of all
// counters, t / episode_length]; done when t == episode_length or the env sum >= limit.
// Like the reference's multi-block envs it synchronises with __sync_env_threads() and keeps
// cross-block data (the env sum) in global memory, so it also runs with blocks_per_env > 1.
__constant__ int kStepTable[4];

extern "C" {

__global__ void CudaCounterEnvStep(int *counters, const int *actions, int *done, float *rewards,
                                   float *obs, int *env_sum, int limit, int *env_timestep,
                                   int episode_length) {
  const int kEnvId = getEnvID(blockIdx.x);
  const int kThisAgentId = getAgentID(threadIdx.x, blockIdx.x, blockDim.x);
  const int kIdx = kEnvId * wkNumberAgents + kThisAgentId;
  if (kThisAgentId == 0) {
    env_timestep[kEnvId] += 1;
    env_sum[kEnvId] = 0;
  }
  __sync_env_threads();
  int value = 0;
  if (kThisAgentId < wkNumberAgents) {
    value = counters[kIdx] + kStepTable[actions[kIdx]];
    counters[kIdx] = value;
    rewards[kIdx] = (float)value;
    atomicAdd(&env_sum[kEnvId], value);
  }
  __sync_env_threads();
  const int total = *((volatile int *)&env_sum[kEnvId]);
  if (kThisAgentId < wkNumberAgents) {
    obs[kIdx * 3 + 0] = (float)value;
    obs[kIdx * 3 + 1] = (float)total;
    obs[kIdx * 3 + 2] = env_timestep[kEnvId] / (float)episode_length;
  }
  if (kThisAgentId == 0 && (env_timestep[kEnvId] == episode_length || total >= limit))
    done[kEnvId] = 1;
}

// uses the optional helpers of wdb_env.cuh (included by the generated runner)
__global__ void CudaCounterEnvReset(int *counters, int value) {
  if (wdb_env::agent_valid()) counters[wdb_env::agent_index()] = value;
}

}  // extern "C"

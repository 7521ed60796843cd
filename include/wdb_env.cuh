// wdb_env.cuh -- device-side helpers for USER env kernels compiled at run time for sm_100a
// (warp_drive_b200/utils/custom_kernels.py).
//
// A custom env file written for the reference (one `extern "C" __global__
// Cuda<Env>Step(...)` using `wkNumberEnvs / wkNumberAgents / wkBlocksPerEnv`, `blockIdx.x /
// wkBlocksPerEnv` as the env id and `threadIdx.x + ...` as the agent id, e.g.
// example_envs/tag_gridworld/tag_gridworld_step_pycuda.cu:131-136) compiles unchanged: the
// generated runner defines the three constants and includes this header before the user
// file.  Everything below is optional sugar for NEW kernels: the same id mapping by name,
// coalesced staging of a per-env array into shared memory (the B200 way to run an O(N^2)
// neighbour sweep on chip), a warp arg-min for nearest-neighbour reductions, and the
// counter-based Philox4x32-10 generator libwdb200 uses (no per-thread state to allocate).
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "wdb_env.cuh targets sm_100a (B200)"
#endif

// ---- names the reference's core service gives every env file (warp_drive/cuda_includes/
// ---- core/env_dim_mapper.h:22-31, env_thread_sync.cu:31-64, array_indexing_util.h) -- same
// ---- names and meaning, so existing env sources compile as they are; own implementations.
__device__ __forceinline__ int getAgentID(const int thread_idx, const int block_idx,
                                          const int block_dim) {
  return wkBlocksPerEnv > 1 ? thread_idx + (block_idx % wkBlocksPerEnv) * block_dim : thread_idx;
}

__device__ __forceinline__ int getEnvID(const int block_idx) { return block_idx / wkBlocksPerEnv; }

// Barrier over every thread of ONE env.  One block per env: the block barrier.  Several blocks
// per env: the blocks of an env are launched as one thread-block CLUSTER (the launcher sets
// the cluster dimension to wkBlocksPerEnv), so the hardware cluster barrier synchronises them
// -- co-scheduled by construction, no spinning on global memory (the reference spins on a
// byte array, which needs all blocks of an env to be resident at once by luck).
__device__ __forceinline__ void __sync_env_threads() {
  if (wkBlocksPerEnv <= 1) {
    __syncthreads();
  } else {
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\t"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

__device__ __forceinline__ int get_flattened_array_index(const int *index_arr, const int *dim_arr,
                                                         const int dimensionality) {
  int flat = 0;
  for (int d = 0; d < dimensionality; d++) flat = flat * dim_arr[d] + index_arr[d];
  return flat;
}

namespace wdb_env {

// ---- id mapping of the reference (function_manager.py:65-67: block = ceil(N / bpe) threads,
// ---- grid = n_envs * bpe blocks)
__device__ __forceinline__ int env_id() { return blockIdx.x / wkBlocksPerEnv; }
__device__ __forceinline__ int agent_id() {
  return threadIdx.x + (blockIdx.x % wkBlocksPerEnv) * blockDim.x;
}
__device__ __forceinline__ bool agent_valid() { return agent_id() < wkNumberAgents; }
__device__ __forceinline__ int agent_index() { return env_id() * wkNumberAgents + agent_id(); }

// ---- stage `n` elements of this env's slice of a [n_envs, n] array into shared memory with
// ---- unit-stride (coalesced) loads; ends with a block barrier.  Single-block envs only.
template <typename T>
__device__ __forceinline__ void stage_env_array(T *smem, const T *global, int n) {
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "32- or 64-bit elements");
  const T *src = global + (size_t)env_id() * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) smem[i] = src[i];
  __syncthreads();
}

// ---- warp arg-min with the reference's tie rule (first index wins, strict <)
__device__ __forceinline__ void warp_argmin(float &value, int &index) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float v = __shfl_down_sync(0xffffffffu, value, off);
    const int i = __shfl_down_sync(0xffffffffu, index, off);
    if (v < value || (v == value && i < index)) { value = v; index = i; }
  }
  value = __shfl_sync(0xffffffffu, value, 0);
  index = __shfl_sync(0xffffffffu, index, 0);
}

// ---- Philox4x32-10 (Salmon et al., SC'11): 4 x 32 random bits for (seed, stream, counter)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ uint4 random_u32x4(unsigned long long seed, unsigned long long stream,
                                              unsigned long long counter) {
  return philox4x32_10(make_uint4((uint32_t)counter, (uint32_t)(counter >> 32),
                                  (uint32_t)stream, (uint32_t)(stream >> 32)),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// u32 -> (0, 1], curand_uniform's mapping
__device__ __forceinline__ float to_uniform(uint32_t x) {
  return x * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
}

}  // namespace wdb_env

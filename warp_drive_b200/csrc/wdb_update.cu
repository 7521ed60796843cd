// wdb_update.cu -- the fused pieces of the A2C / PPO update.
//
// (1) pg_loss_kernel: ONE backward-in-time scan per (env, agent) over the [T, E, Np] batch that
//     produces everything between the policy forward and the backward pass of the reference's
//     update (warp_drive/training/algorithms/policygradient/a2c.py:80-130, ppo.py:82-141):
//     bootstrapped discounted returns with done masking, advantages, the Categorical log-prob
//     of the taken action and the entropy of every action head, the three loss sums (policy,
//     value, entropy) AND the gradients of the total loss with respect to the probabilities
//     and the values.  The reference builds the same quantities out of ~40 elementwise torch
//     kernels per head plus ~6 per timestep for the returns recursion, and autograd then walks
//     that graph backwards.
// (2) sumsq_kernel + adam_kernel: gradient-norm clipping and Adam over ONE flat parameter
//     arena (every trained tensor of a policy is a view into it), no host synchronisation:
//     the clip factor is read from device memory by the Adam kernel.
// The batch forward / backward of the MLP stay on cuBLAS: they are plain [T*E*Np, F] x [F, H]
// library GEMMs with no fusion partner (the loss gradient arrives as a dense tensor), which is
// what cuBLAS is for.
#include <math_constants.h>

#include "wdb_common.cuh"

using namespace wdb;

namespace {

constexpr int kPgThreads = 128;
constexpr float kProbEps = 1.1920928955078125e-07f;   // torch.finfo(float32).eps (clamp_probs)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  return v;
}

// sums[0] = sum(-logp * adv), sums[1] = sum((V - R)^2), sums[2] = sum over heads of entropy,
// sums[3] = sum(adv)  (PPO's surrogate at ratio == 1 is -mean(adv))
__global__ void __launch_bounds__(kPgThreads)
pg_loss_kernel(const __grid_constant__ wdb_pg_loss L) {
  const long long per_t = (long long)L.n_envs * L.n_agents;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < per_t;
  const int env = live ? (int)(i / L.n_agents) : 0;
  const double inv_m = 1.0 / ((double)L.T * (double)per_t);
  const float inv_mf = (float)inv_m;
  double s_pol = 0.0, s_vf = 0.0, s_ent = 0.0, s_adv = 0.0;
  if (live) {
    float ret = 0.0f;
    for (int t = L.T - 1; t >= 0; t--) {
      const long long idx = (long long)t * per_t + i;
      const int d = L.done[(long long)t * L.n_envs + env] > 0;
      const float v = L.values[idx];
      // a2c.py:80-93: returns[T-1] = done ? r : V ;  returns[t] = r + (done ? 0 : gamma * next)
      if (t == L.T - 1) ret = d ? L.rewards[idx] : v;
      else ret = L.rewards[idx] + (d ? 0.0f : L.gamma * ret);
      if (L.returns) L.returns[idx] = ret;
      const float adv = ret - v;
      const float dv = v - ret;
      s_vf += (double)dv * (double)dv;
      s_adv += (double)adv;
      if (L.grad_values) L.grad_values[idx] = 2.0f * L.vf_coeff * dv * inv_mf;
      float logp = 0.0f;
      for (int k = 0; k < L.n_heads; k++) {
        const int A = L.n_actions[k];
        const float *p = L.probs[k] + idx * A;
        float *g = L.grad_probs[k] ? L.grad_probs[k] + idx * A : nullptr;
        const int a = L.actions[idx * L.n_heads + k];
        // torch.distributions.Categorical(probs=p): probs / sum, logits = log(clamp(probs,
        // eps, 1 - eps)); log_prob = logits[a]; entropy = -sum(logits * probs)
        float z = 0.0f;
        for (int j = 0; j < A; j++) z += p[j];
        const float inv_z = 1.0f / z;
        // pass 1: entropy, log-prob and S = sum_i (dL/dq_i) q_i  (q = p / z)
        float ent = 0.0f, S = 0.0f;
        for (int j = 0; j < A; j++) {
          const float q = p[j] * inv_z;
          const bool inside = q > kProbEps && q < 1.0f - kProbEps;
          const float lq = logf(fminf(fmaxf(q, kProbEps), 1.0f - kProbEps));
          ent -= lq * q;
          float gq = L.entropy_coeff * (lq + (inside ? 1.0f : 0.0f));   // -c_ent * dH/dq_j
          if (j == a) {
            logp += lq;
            if (inside) gq -= adv / q;                                   // -adv * dlogp/dq_a
          }
          S += gq * q;
        }
        // pass 2: dL/dp_j = (dL/dq_j - S) / z, scaled by 1 / M
        if (g) {
          for (int j = 0; j < A; j++) {
            const float q = p[j] * inv_z;
            const bool inside = q > kProbEps && q < 1.0f - kProbEps;
            const float lq = logf(fminf(fmaxf(q, kProbEps), 1.0f - kProbEps));
            float gq = L.entropy_coeff * (lq + (inside ? 1.0f : 0.0f));
            if (j == a && inside) gq -= adv / q;
            g[j] = (gq - S) * inv_mf * inv_z;
          }
        }
        s_ent += (double)ent;
      }
      s_pol += (double)(-logp * adv);
    }
  }
  // block reduction -> 4 double atomics per CTA
  __shared__ double red[4][kPgThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  s_pol = warp_sum(s_pol); s_vf = warp_sum(s_vf); s_ent = warp_sum(s_ent); s_adv = warp_sum(s_adv);
  if (lane == 0) { red[0][warp] = s_pol; red[1][warp] = s_vf; red[2][warp] = s_ent; red[3][warp] = s_adv; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double acc = 0.0;
    for (int w = 0; w < kPgThreads / 32; w++) acc += red[threadIdx.x][w];
    atomicAdd(&L.sums[threadIdx.x], acc);
  }
}

// sum of squares of a flat float buffer -> *out (double, zeroed by the caller)
__global__ void __launch_bounds__(256) sumsq_kernel(const float *__restrict__ g, long long n,
                                                    double *out) {
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i];
    acc += (double)v * (double)v;
  }
  acc = warp_sum(acc);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; w++) s += red[w];
    atomicAdd(out, s);
  }
}

// torch.optim.Adam (no weight decay, no amsgrad) on a flat arena; the gradient is first scaled
// by clip = min(1, max_norm / (norm + 1e-6)) like torch.nn.utils.clip_grad_norm_ (the scaled
// gradient is written back: callers that log the clipped gradient see the same values)
__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m,
            float *__restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
            float bias1, float bias2_sqrt, float max_norm, const double *sumsq) {
  float clip = 1.0f;
  if (max_norm > 0.0f && sumsq) {
    const float norm = (float)sqrt(*sumsq);
    clip = fminf(1.0f, max_norm / (norm + 1.0e-6f));
  }
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * clip;
  g[i] = gi;
  const float mi = beta1 * m[i] + (1.0f - beta1) * gi;      // lerp form of torch's foreach Adam
  const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bias2_sqrt + eps;
  p[i] = p[i] - (lr / bias1) * (mi / denom);
}

}  // namespace

WDB_API int wdb_pg_loss_and_grads(void *stream, const wdb_pg_loss *l) {
  if (!l || l->T < 1 || l->n_envs < 1 || l->n_agents < 1 || l->n_heads < 1 || l->n_heads > 4)
    return (int)cudaErrorInvalidValue;
  if (!l->values || !l->actions || !l->rewards || !l->done || !l->sums)
    return (int)cudaErrorInvalidValue;
  for (int k = 0; k < l->n_heads; k++)
    if (!l->probs[k] || l->n_actions[k] < 1) return (int)cudaErrorInvalidValue;
  const long long n = (long long)l->n_envs * l->n_agents;
  const int grid = (int)((n + kPgThreads - 1) / kPgThreads);
  pg_loss_kernel<<<grid, kPgThreads, 0, as_stream(stream)>>>(*l);
  return finish_launch();
}

WDB_API int wdb_grad_sumsq(void *stream, const float *grads, long long n, double *out) {
  if (!grads || !out || n < 1) return (int)cudaErrorInvalidValue;
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(grads, n, out);
  return finish_launch();
}

WDB_API int wdb_adam_step(void *stream, float *params, float *grads, float *exp_avg,
                          float *exp_avg_sq, long long n, float lr, float beta1, float beta2,
                          float eps, int step, float max_grad_norm, const double *grad_sumsq) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || n < 1 || step < 1)
    return (int)cudaErrorInvalidValue;
  if (max_grad_norm > 0.0f && !grad_sumsq) return (int)cudaErrorInvalidValue;
  const float bias1 = 1.0f - powf(beta1, (float)step);
  const float bias2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  adam_kernel<<<(int)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
      params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bias1, bias2_sqrt,
      max_grad_norm, grad_sumsq);
  return finish_launch();
}

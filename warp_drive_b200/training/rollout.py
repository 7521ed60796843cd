"""RolloutEngine: the device-resident rollout loop
`policy forward -> sample -> env.step -> bookkeeping -> done-masked reset -> push to batch`.

This is the B200 re-design of TrainerBase._generate_rollout_batch and its helpers
(warp_drive/training/trainers/trainer_base.py:383-601, trainer_a2c.py:159-216): the
reference issues ~55 kernel launches and >= 5 host synchronisations per timestep
(SURVEY.md section 3.2); here one timestep has NO host synchronisation (done handling is
a device-side mask), and the whole T-step rollout is captured once in a CUDA graph and
replayed.
"""
import ctypes

import torch

from warp_drive_b200 import lib as _lib
from warp_drive_b200.utils.constants import Constants

_OBSERVATIONS = Constants.OBSERVATIONS
_ACTIONS = Constants.ACTIONS
_REWARDS = Constants.REWARDS
_DONE_FLAGS = Constants.DONE_FLAGS
_PROCESSED_OBSERVATIONS = Constants.PROCESSED_OBSERVATIONS


class RolloutEngine:
    def __init__(self, env_wrapper, models, policy_tag_to_agent_id_map, sampler,
                 batch_size_per_env, use_cuda_graph=True, forward_dtype=None,
                 use_fused_step=True, write_observations=True, stats=None,
                 use_fused_forward=True, use_obs_tiles=False, use_pair_forward=False):
        self.env_wrapper = env_wrapper
        self.use_pair_forward = bool(use_pair_forward)
        # A/B switch: with one forward launch per policy, launch the env step programmatically
        # dependent on the main-stream forward (option "pdl" must be on as well)
        self.pdl_after_fork = False
        self.dm = env_wrapper.cuda_data_manager
        self.models = models
        self.policy_map = policy_tag_to_agent_id_map
        self.policies = list(policy_tag_to_agent_id_map.keys())
        self.sampler = sampler
        self.T = int(batch_size_per_env)
        self.use_cuda_graph = use_cuda_graph
        self.forward_dtype = forward_dtype
        dev = self.dm.device
        self.E = env_wrapper.n_envs
        self.N = env_wrapper.n_agents
        self.ids = {p: torch.as_tensor(ids, dtype=torch.long, device=dev)
                    for p, ids in self.policy_map.items()}
        self.covers_all = {p: len(ids) == self.N and list(ids) == list(range(self.N))
                           for p, ids in self.policy_map.items()}
        first = self.policy_map[self.policies[0]][0]
        from warp_drive_b200.training.utils.data_loader import action_head_sizes

        self.heads, self.continuous = action_head_sizes(env_wrapper.env.action_space[first])
        self.n_heads = len(self.heads)
        self.combined = None
        if len(self.policies) > 1:
            self.combined = [torch.zeros((self.E, self.N, h), device=dev) for h in self.heads]
        # episodic statistics (device scalars: no host sync while rolling out)
        self.reward_running_sum = {p: torch.zeros((self.E, len(ids)), device=dev)
                                   for p, ids in self.policy_map.items()}
        self.step_running_sum = torch.zeros(self.E, dtype=torch.int32, device=dev)
        self.episodic_reward_sum = {p: torch.zeros((), device=dev) for p in self.policies}
        self.episodic_step_sum = torch.zeros((), dtype=torch.int64, device=dev)
        self.num_completed_episodes = torch.zeros((), dtype=torch.int64, device=dev)
        self._graph = None
        self._graph_stream = None
        # ---- fused single-launch timestep (TagContinuous: 2 discrete heads, Box obs)
        self.fused = None
        if use_fused_step and self._fused_eligible():
            from warp_drive_b200.training.fused_tag_continuous import FusedTagContinuousStep

            self.fused = FusedTagContinuousStep(
                env_wrapper, self.policy_map, sampler, write_observations=write_observations,
                stats=stats)
            self.fused.set_bookkeeping(
                self.reward_running_sum, self.episodic_reward_sum, self.step_running_sum,
                self.episodic_step_sum, self.num_completed_episodes)
            obs = self._tensor(_OBSERVATIONS).view(self.E, self.N, -1)
            # observations of the *next* forward pass, per policy (filled by the kernel)
            self.cur_obs = {p: obs.index_select(1, self.ids[p]).contiguous()
                            for p in self.policies}

        # ---- whole-rollout single launch (discrete single-agent envs, small FullyConnected)
        self.sa = None
        if use_fused_step and self.fused is None:
            from warp_drive_b200.training.fused_single_agent import FusedSingleAgentRollout

            if (self.T >= 1 and not self.continuous
                    and FusedSingleAgentRollout.eligible(env_wrapper, models, self.policy_map)):
                p0 = self.policies[0]
                self.sa = FusedSingleAgentRollout(env_wrapper, models[p0], p0, sampler)
                self.sa.set_bookkeeping(
                    self.reward_running_sum, self.episodic_reward_sum, self.step_running_sum,
                    self.episodic_step_sum, self.num_completed_episodes)

        # ---- tensor-core forward (tcgen05 MLP kernel) for the fused path
        self.fused_forward = {}
        self.obs_tiles = {}
        if self.sa is None and use_fused_forward and forward_dtype is None and not self.continuous:
            from warp_drive_b200.training.models.fused_forward import FusedPolicyForward

            # (also on the generic multi-launch path: gridworld, classic control with a policy
            #  too large for the whole-rollout kernel, envs with reset pools ...)
            if all(FusedPolicyForward.supported(m) for m in self.models.values()):
                self.fused_forward = {p: FusedPolicyForward(self.models[p]) for p in self.policies}
                self._probs = {p: [torch.empty((self.E, len(self.policy_map[p]), h), device=dev)
                                   for h in self.heads] for p in self.policies}
                # bf16 MMA-ready mirror of the observations the next forward reads: written by
                # the fused env step itself (and by pack_obs whenever cur_obs changes outside it)
                # (off by default: measured at config 2, the env step's extra pass costs what
                #  the forward saves -- see DESIGN.md 3.2)
                if (use_obs_tiles and self.fused is not None
                        and int(getattr(env_wrapper, "blocks_per_env", 1) or 1) == 1
                        and not getattr(env_wrapper.env, "use_full_observation", False)):
                    self.obs_tiles = {
                        p: torch.zeros(self.fused_forward[p].tiles_bytes(
                            self.E * len(self.policy_map[p])), dtype=torch.uint8, device=dev)
                        for p in self.policies}

    def refresh_forward_weights(self):
        """Re-pack the policies' parameters for the tensor-core forward (after every
        optimizer step / checkpoint load)."""
        for fwd in self.fused_forward.values():
            fwd.refresh()

    def _fused_eligible(self):
        from warp_drive_b200.utils.spaces import Box, MultiDiscrete

        env = self.env_wrapper.env
        if getattr(env, "name", "") != "TagContinuous" or len(self.policies) > 4:
            return False
        if self.continuous or self.n_heads != 2:
            return False
        if not isinstance(env.action_space[0], MultiDiscrete):
            return False
        if not isinstance(env.observation_space[0], Box):
            return False
        if self.T < 1 or self.dm.reset_target_to_pool:
            return False
        bpe = int(getattr(self.env_wrapper, "blocks_per_env", 1) or 1)
        k_obs = int(getattr(env, "num_other_agents_observed", 0))
        full = bool(getattr(env, "use_full_observation", False))
        if bpe > 1 or (self.N > 320 and not full and k_obs + 2 <= 16):
            # one env per thread-block cluster, or one large env per CTA (wdb_tc_wide.cu):
            # <= 1024 agents per CTA
            if -(-self.N // bpe) > 1024:
                return False
        elif not full:
            # the fused step pushes the per-policy observations from its shared-memory tile
            # (same sizing rule as plan_launch in wdb_tag_continuous.cu); envs too large for
            # that (e.g. 1024 agents) take the generic multi-launch path
            n, f = self.N, 7 * int(env.num_other_agents_observed) + 1
            epb = max(1, 320 // n)
            n_warps = (epb * n + 31) // 32
            small = 36 * epb * n + 16 * n + max(8 * n, 1056) * n_warps
            if small + 4 * epb * n * f > 113 * 1024:
                return False
        return all(getattr(m, "action_mask", None) is None for m in self.models.values())

    def resync_observations(self):
        """Re-gather the per-policy observation buffers from `observations` (call after
        anything outside the engine changed the env state, e.g. reset_all_envs)."""
        if self.fused is not None:
            obs = self._tensor(_OBSERVATIONS).view(self.E, self.N, -1)
            for p in self.policies:
                self.cur_obs[p].copy_(obs.index_select(1, self.ids[p]))
            self._repack_obs_tiles()

    def _repack_obs_tiles(self):
        if self.obs_tiles:
            for p in self.policies:
                self.fused_forward[p].pack_obs(self.cur_obs[p], self.obs_tiles[p])

    def materialize_observations(self):
        """Scatter the per-policy observation buffers back into the [E, N, F]
        `observations` array (only needed when the engine runs with
        write_observations=False and someone wants to read that array)."""
        if self.fused is not None:
            obs = self._tensor(_OBSERVATIONS).view(self.E, self.N, -1)
            for p in self.policies:
                obs.index_copy_(1, self.ids[p], self.cur_obs[p])

    def _forward(self, model, obs_p):
        if self.forward_dtype is not None:
            with torch.autocast("cuda", dtype=self.forward_dtype):
                probs, _ = model(obs_p)
            return [q.float().contiguous() for q in probs]
        probs, _ = model(obs_p)
        return [q.contiguous() for q in probs]

    def _forward_side_by_side(self, obs_in, probs, weights_stable=False):
        """Every policy's fused forward, concurrently: the persistent MLP kernel uses one CTA
        per SM, so the SMs are split between the policies in proportion to their rows and the
        smaller policies run on side streams (fork/join by events, also under graph capture).
        Sequential launches would each pay the 195 KB weight load and leave most SMs idle
        during the small policy's single wave."""
        if len(self.policies) == 1:
            self._one_forward(self.policies[0], obs_in, probs)
            return False
        if len(self.policies) == 2 and not self.obs_tiles and self.use_pair_forward:
            # both policies in ONE launch (CTAs split by tile counts), no fork / join
            from warp_drive_b200.training.models.fused_forward import forward_pair

            pa, pb = sorted(self.policies,
                            key=lambda p: -(obs_in[p].numel() // obs_in[p].shape[-1]))
            forward_pair(self.fused_forward[pa], self.fused_forward[pb], obs_in[pa], obs_in[pb],
                         probs[pa], probs[pb], weights_stable=weights_stable)
            return True
        if not hasattr(self, "_fwd_plan"):
            n_sm = torch.cuda.get_device_properties(self.dm.device).multi_processor_count
            rows = {p: obs_in[p].numel() // obs_in[p].shape[-1] for p in self.policies}
            total = sum(rows.values())
            order = sorted(self.policies, key=lambda p: -rows[p])
            if len(order) == 2:
                # cost model in tile-times: ceil(tiles / SMs) + ~1.5 for the weight load;
                # pick the split that minimises the slower of the two
                tiles = {p: -(-rows[p] // 128) for p in order}
                best = None
                for k in range(1, n_sm // 2 + 1):
                    # (measured: the small policy's tiles run ~1.3x slower -- fewer CTAs share
                    #  its weights in L2 -- and a launch pays ~1.5 tile-times up front)
                    cost = max(-(-tiles[order[0]] // (n_sm - k)),
                               1.3 * -(-tiles[order[1]] // k)) + 1.5
                    if best is None or cost < best[0]:
                        best = (cost, k)
                share = {order[1]: best[1], order[0]: n_sm - best[1]}
            else:
                share = {p: max(1, min(n_sm - 1, -(-n_sm * rows[p] // total))) for p in order[1:]}
                share[order[0]] = max(1, n_sm - sum(share.values()))
            self._fwd_plan = (order, share)
            self._fwd_streams = [torch.cuda.Stream() for _ in order[1:]]
        order, share = self._fwd_plan
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        joins = []
        for p, st in zip(order[1:], self._fwd_streams):
            st.wait_event(fork)
            with torch.cuda.stream(st):
                self._one_forward(p, obs_in, probs, share[p])
                ev = torch.cuda.Event()
                ev.record(st)
            joins.append(ev)
        p = order[0]
        self._one_forward(p, obs_in, probs, share[p])
        for ev in joins:
            cur.wait_event(ev)
        return self.pdl_after_fork

    def _one_forward(self, p, obs_in, probs, max_ctas=0):
        if self.obs_tiles:
            rows = obs_in[p].numel() // obs_in[p].shape[-1]
            self.fused_forward[p].forward_tiles(self.obs_tiles[p], rows, probs[p][0], probs[p][1],
                                                max_ctas=max_ctas)
        else:
            self.fused_forward[p](obs_in[p], probs[p][0], probs[p][1], max_ctas=max_ctas)

    def step_fused(self, t, uniforms=None):
        """One timestep = policy forwards + ONE libwdb200 launch."""
        with torch.no_grad():
            T = self.T
            if t >= 0:
                batches = {p: self._tensor(f"{_PROCESSED_OBSERVATIONS}_batch_{p}")
                           for p in self.policies}
                if t == 0:
                    for p in self.policies:
                        batches[p][0].copy_(self.cur_obs[p])
                obs_in = {p: batches[p][t] for p in self.policies}
                obs_next = {p: (batches[p][t + 1] if t + 1 < T else self.cur_obs[p])
                            for p in self.policies}
                actions_batch = {p: self._tensor(f"{_ACTIONS}_batch_{p}")[t]
                                 for p in self.policies}
                rewards_batch = {p: self._tensor(f"{_REWARDS}_batch_{p}")[t]
                                 for p in self.policies}
                done_batch = self._tensor(f"{_DONE_FLAGS}_batch")[t]
            else:   # evaluation-style step: no batch push
                obs_in = {p: self.cur_obs[p] for p in self.policies}
                if not hasattr(self, "_scratch_obs"):
                    self._scratch_obs = {p: torch.empty_like(self.cur_obs[p])
                                         for p in self.policies}
                obs_next = self._scratch_obs
                actions_batch = rewards_batch = done_batch = None
            paired = False
            if self.fused_forward:
                probs = self._probs
                # t > 0: the kernel before this forward is the previous env step, not a weight
                # re-pack -> the weight load may overlap its tail (programmatic launch)
                paired = self._forward_side_by_side(obs_in, probs, weights_stable=t > 0)
            else:
                probs = {p: self._forward(self.models[p], obs_in[p]) for p in self.policies}
            self.fused.launch(probs, actions_batch=actions_batch, rewards_batch=rewards_batch,
                              obs_next=obs_next, done_batch=done_batch, uniforms=uniforms,
                              obs_next_tiles=self.obs_tiles or None, after_forward=paired)
            if t < 0:
                for p in self.policies:
                    self.cur_obs[p].copy_(self._scratch_obs[p])

    # ------------------------------------------------------------------ one timestep
    def _tensor(self, name):
        return self.dm.data_on_device_via_torch(name)

    # ---- native data movement of the generic path (csrc/wdb_rollout_generic.cu) ----------
    def _ids32(self, p):
        """int32 device copy of the policy's agent ids (None: the policy covers 0..N-1)."""
        if not hasattr(self, "_ids32_cache"):
            self._ids32_cache = {q: (None if self.covers_all[q]
                                     else self.ids[q].to(torch.int32).contiguous())
                                 for q in self.policies}
        return self._ids32_cache[p]

    def _move_rows(self, full, rows, scatter=False):
        """full [E, N, W] <-> rows {policy: [E, Np, W]} (4-byte elements), ONE launch for all
        policies in `rows` (wdb_gather_policy_rows)."""
        assert full.is_contiguous() and full.element_size() == 4
        g = _lib.Gather()
        g.n_envs, g.n_agents = self.E, self.N
        g.width = full.numel() // (self.E * self.N)
        g.n_policies, g.scatter = len(rows), int(bool(scatter))
        g.full = _lib.ptr(full)
        for i, (p, r) in enumerate(rows.items()):
            assert r.is_contiguous() and r.element_size() == 4
            assert r.numel() == self.E * len(self.policy_map[p]) * g.width, (p, r.shape)
            io = g.policy[i]
            io.n_agents = len(self.policy_map[p])
            io.agent_ids = _lib.ptr(self._ids32(p))
            io.rows = _lib.ptr(r)
        _lib.check(_lib.load().wdb_gather_policy_rows(_lib.stream_ptr(), ctypes.byref(g)),
                   "gather_policy_rows")

    def evaluate_policies(self, t):
        """obs -> per-head probabilities [E, N, A_k] (trainer_a2c.py:159-216).  The per-policy
        observation rows go to the batch slot in one gather launch and the forward reads the
        slot; the per-policy probabilities return to the sampler's [E, N, A] arrays in one
        scatter launch per head."""
        obs = self._tensor(_OBSERVATIONS).view(self.E, self.N, -1)
        if t >= 0:
            obs_p = {p: self._tensor(f"{_PROCESSED_OBSERVATIONS}_batch_{p}")[t]
                     for p in self.policies}
            self._move_rows(obs, obs_p)
        else:
            if not hasattr(self, "_obs_scratch"):
                self._obs_scratch = {
                    p: torch.empty((self.E, len(self.policy_map[p]), obs.shape[-1]),
                                   dtype=obs.dtype, device=obs.device)
                    for p in self.policies if not self.covers_all[p]}
            obs_p = {p: obs for p in self.policies if self.covers_all[p]}
            obs_p.update(self._obs_scratch)
            if self._obs_scratch:
                self._move_rows(obs, self._obs_scratch)
        per_policy = {}
        for p in self.policies:
            if self.fused_forward:
                # tcgen05 forward straight into the persistent probability buffers
                bufs = self._probs[p]
                self.fused_forward[p](obs_p[p], bufs[0], bufs[1] if len(bufs) > 1 else None)
                per_policy[p] = bufs
            else:
                per_policy[p] = self._forward(self.models[p], obs_p[p].view(
                    self.E, len(self.policy_map[p]), -1))
        if self.combined is None:
            return per_policy[self.policies[0]]
        for k in range(self.n_heads):
            self._move_rows(self.combined[k], {p: per_policy[p][k] for p in self.policies},
                            scatter=True)
        return self.combined

    def sample_actions(self, probs, t, **sample_params):
        """probs -> sampled_actions (trainer_base.py:466-512); the per-policy batch push of the
        actions happens in bookkeep()."""
        dm = self.dm
        if self.n_heads == 1:
            self.sampler.sample(dm, probs[0], _ACTIONS, write_cum_distr=False, **sample_params)
        else:
            actions = self._tensor(_ACTIONS)
            for k in range(self.n_heads):
                self.sampler.sample(dm, probs[k], f"{_ACTIONS}_{k}", write_cum_distr=False,
                                    combined=(actions, self.n_heads, k), **sample_params)

    def bookkeep(self, t):
        """done / rewards / actions -> batches, running episodic sums: ONE launch
        (wdb_rollout_bookkeep), no host synchronisation (trainer_base.py:514-601 uses
        nonzero() / len())."""
        k = _lib.Bookkeep()
        k.n_envs, k.n_agents, k.n_policies = self.E, self.N, len(self.policies)
        actions = self._tensor(_ACTIONS) if self.dm.is_data_on_device_via_torch(_ACTIONS) else None
        k.n_heads = (actions.numel() // (self.E * self.N)) if actions is not None else 0
        k.done = _lib.ptr(self._tensor("_done_"))
        k.rewards = _lib.ptr(self._tensor(_REWARDS))
        k.actions = _lib.ptr(actions)
        k.done_batch = _lib.ptr(self._tensor(f"{_DONE_FLAGS}_batch")[t]) if t >= 0 else None
        k.step_running_sum = _lib.ptr(self.step_running_sum)
        k.episodic_step_sum = _lib.ptr(self.episodic_step_sum)
        k.num_completed_episodes = _lib.ptr(self.num_completed_episodes)
        for i, p in enumerate(self.policies):
            io = k.policy[i]
            io.n_agents = len(self.policy_map[p])
            io.agent_ids = _lib.ptr(self._ids32(p))
            if t >= 0:
                io.rewards_batch = _lib.ptr(self._tensor(f"{_REWARDS}_batch_{p}")[t])
                if actions is not None:
                    io.actions_batch = _lib.ptr(self._tensor(f"{_ACTIONS}_batch_{p}")[t])
            io.reward_running_sum = _lib.ptr(self.reward_running_sum[p])
            io.episodic_reward_sum = _lib.ptr(self.episodic_reward_sum[p])
        _lib.check(_lib.load().wdb_rollout_bookkeep(_lib.stream_ptr(), ctypes.byref(k)),
                   "rollout_bookkeep")

    def bookkeep_torch(self, t):
        """The same bookkeeping in plain torch ops (the parity reference of bookkeep() in
        tests/test_gpu_rollout.py; not used by the engine)."""
        raw_done = self._tensor("_done_")
        rewards = self._tensor(_REWARDS)
        if t >= 0:
            self._tensor(f"{_DONE_FLAGS}_batch")[t].copy_(raw_done)
            actions = self._tensor(_ACTIONS)
            for p in self.policies:
                self._tensor(f"{_ACTIONS}_batch_{p}")[t].copy_(
                    actions if self.covers_all[p] else actions.index_select(1, self.ids[p]))
        done = (raw_done > 0).to(torch.int32)
        donef = done.to(torch.float32)
        for p in self.policies:
            r_p = rewards if self.covers_all[p] else rewards.index_select(1, self.ids[p])
            if t >= 0:
                self._tensor(f"{_REWARDS}_batch_{p}")[t].copy_(r_p)
            run = self.reward_running_sum[p]
            run.add_(r_p)
            self.episodic_reward_sum[p].add_((run * donef[:, None]).sum())
            run.mul_(1.0 - donef[:, None])
        self.step_running_sum.add_(1)
        self.episodic_step_sum.add_((self.step_running_sum * done).sum())
        self.step_running_sum.mul_(1 - done)
        self.num_completed_episodes.add_(done.sum())

    def step(self, t, **sample_params):
        if self.fused is not None and not sample_params:
            return self.step_fused(t)
        if self.sa is not None and self._sa_params_ok(sample_params):
            with torch.no_grad():
                self.sa.launch(1, t0=max(t, 0), record=t >= 0, **sample_params)
            return None
        with torch.no_grad():
            probs = self.evaluate_policies(t)
            self.sample_actions(probs, t, **sample_params)
            self.env_wrapper.step_all_envs()
            self.bookkeep(t)
            # done-masked reset: a no-op for envs that are not done, so no host-side
            # `if done_flags.any()` (trainer_base.py:421-422) is needed
            self.env_wrapper.reset_only_done_envs()

    # ------------------------------------------------------------------ T-step rollout
    def _rollout_eager(self, **sample_params):
        for t in range(self.T):
            self.step(t, **sample_params)

    @staticmethod
    def _sa_params_ok(sample_params):
        return all(k == "use_argmax" for k in sample_params)

    def rollout(self, **sample_params):
        """Generate one training batch (T timesteps for every env replica)."""
        if self.sa is not None and self._sa_params_ok(sample_params):
            # ONE launch for all T timesteps (no CUDA graph needed)
            with torch.no_grad():
                self.sa.launch(self.T, t0=0, record=True, **sample_params)
            return
        if not self.use_cuda_graph or sample_params:
            self._rollout_eager(**sample_params)
            return
        if self._graph is None:
            # warm-up on a side stream (cuBLAS workspaces, lazy inits), then capture
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                # an evaluation-style step: advances the envs AND the engine's current
                # observations consistently, records nothing (a step(0) here would leave
                # cur_obs one state behind the env for the first captured transition)
                self.step(-1)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._rollout_eager()
            self._graph = graph
        self._graph.replay()

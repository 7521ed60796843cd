"""Trainer: device-resident rollouts + A2C / PPO updates.

Public surface of the reference's TrainerBase / TrainerA2C
(warp_drive/training/trainers/trainer_base.py:69-846, trainer_a2c.py:43-384): same
constructor arguments, same YAML run-config schema merged over default_configs.yaml, same
`train / save_model_checkpoint / load_model_checkpoint / fetch_episode_states /
graceful_close`, same `{policy}_{timestep}.state_dict` checkpoints and JSON-lines metrics.

What is different underneath
  * rollouts go through RolloutEngine: one libwdb200 launch per timestep for
    TagContinuous (fused sample+step+push+reset), CUDA-graph replay of the whole T-step
    rollout, no host synchronisation inside a rollout;
  * multi-GPU is one process per GPU with a NCCL process group; every iteration issues
    ONE flat all-reduce over the gradients of all trained policies (the reference wraps
    each model in DDP over a gloo group: trainer_a2c.py:137-146, process_group_torch.py:7).
"""
import json
import logging
import os
import random
import time

import numpy as np
import torch
import yaml
from torch import nn

from warp_drive_b200.managers.function_manager import CUDALogController, CUDASampler
from warp_drive_b200.training.algorithms.policygradient import A2C, PPO
from warp_drive_b200.training.models.fully_connected import ModelFactory
from warp_drive_b200.training.rollout import RolloutEngine
from warp_drive_b200.training.utils.data_loader import (
    create_and_push_data_placeholders, validate_policy_map)
from warp_drive_b200.training.utils.param_scheduler import ParamScheduler
from warp_drive_b200.utils.constants import Constants

_ACTIONS = Constants.ACTIONS
_REWARDS = Constants.REWARDS
_DONE_FLAGS = Constants.DONE_FLAGS
_PROCESSED_OBSERVATIONS = Constants.PROCESSED_OBSERVATIONS
_EPSILON = 1e-10
_CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_configs")


def recursive_merge_config_dicts(config, default_config):
    """Fill every key missing from `config` with the default (trainer_base.py:46-60)."""
    assert isinstance(config, dict) and isinstance(default_config, dict)
    for k, v in default_config.items():
        if k not in config:
            config[k] = v
        elif isinstance(v, dict) and isinstance(config[k], dict):
            recursive_merge_config_dicts(config[k], v)
    return config


def load_run_config(name_or_path):
    path = name_or_path if os.path.exists(name_or_path) else os.path.join(
        _CONFIG_DIR, f"{name_or_path}.yaml")
    with open(path, "r", encoding="utf8") as fp:
        return yaml.safe_load(fp)


def verbose_print(message, device_id=0):
    print(f"[Device {device_id or 0}]: {message} ")


class PerfStats:
    """Timing of the rollout / training phases with CUDA events; `steps` counts env-steps
    like the reference (iters x train_batch_size, trainer_base.py:374) and
    `agent_steps` additionally multiplies by the number of agents."""

    def __init__(self):
        self.iters = 0
        self.steps = 0
        self.agent_steps = 0
        self.rollout_time = 0.0
        self.training_time = 0.0
        self.total_time = 0.0

    def get_perf_stats(self):
        it = max(self.iters, 1)
        tot = max(self.total_time, 1e-12)
        return {
            "Mean rollout time per iter (ms)": 1000 * self.rollout_time / it,
            "Mean training time per iter (ms)": 1000 * self.training_time / it,
            "Mean total time per iter (ms)": 1000 * self.total_time / it,
            "Mean steps per sec (rollout)": self.steps / max(self.rollout_time, 1e-12),
            "Mean steps per sec (training time)": self.steps / max(self.training_time, 1e-12),
            "Mean steps per sec (total)": self.steps / tot,
            "Mean agent-steps per sec (total)": self.agent_steps / tot,
        }


class Trainer:
    def __init__(self, env_wrapper=None, config=None, policy_tag_to_agent_id_map=None,
                 create_separate_placeholders_for_each_policy=False,
                 obs_dim_corresponding_to_num_agents="first", num_devices=1, device_id=0,
                 results_dir=None, verbose=True, use_cuda_graph=True):
        assert env_wrapper is not None and env_wrapper.env_backend != "cpu"
        assert config is not None
        assert obs_dim_corresponding_to_num_agents in ("first", "last")
        # the rollout engine views observations as [E, N, F]; 'last' ([E, ..., N]) would be
        # silently scrambled across agents and features
        assert obs_dim_corresponding_to_num_agents == "first", (
            "obs_dim_corresponding_to_num_agents='last' is not supported by the B200 rollout "
            "engine (observations must be [num_envs, num_agents, ...])")
        assert not create_separate_placeholders_for_each_policy, (
            "separate per-policy step placeholders are not implemented in the B200 trainer")
        self.cuda_envs = env_wrapper
        self.verbose = verbose
        self.num_devices, self.device_id = num_devices, device_id

        with open(os.path.join(_CONFIG_DIR, "default_configs.yaml"), encoding="utf8") as fp:
            default = yaml.safe_load(fp)
        self.config = config
        config["trainer"] = recursive_merge_config_dicts(config["trainer"], default["trainer"])
        for key in config["policy"]:
            config["policy"][key] = recursive_merge_config_dicts(config["policy"][key],
                                                                 default["policy"])
        config["saving"] = recursive_merge_config_dicts(config["saving"], default["saving"])
        self.sample_params_schedules = {
            k: ParamScheduler(v) for k, v in config.get("sampler", {}).get("params", {}).items()}

        results_dir = results_dir or f"{time.time():10.0f}"
        s = config["saving"]
        self.save_dir = os.path.join(s["basedir"], s["name"], s["tag"], results_dir)
        os.makedirs(self.save_dir, exist_ok=True)
        with open(os.path.join(self.save_dir, "run_config.json"), "a+", encoding="utf8") as fp:
            json.dump(config, fp)
            fp.write("\n")

        self.policy_tag_to_agent_id_map = validate_policy_map(env_wrapper,
                                                              policy_tag_to_agent_id_map)
        self.policies = list(config["policy"].keys())
        assert set(self.policies) == set(self.policy_tag_to_agent_id_map.keys())
        self.policies_to_train = [p for p in self.policies if config["policy"][p]["to_train"]]

        t = config["trainer"]
        self.num_episodes = t["num_episodes"]
        self.training_batch_size = t["train_batch_size"]
        self.num_envs = t["num_envs"]
        assert self.num_envs == env_wrapper.n_envs
        self.training_batch_size_per_env = self.training_batch_size // self.num_envs
        assert self.training_batch_size_per_env > 0

        # first reset pushes every array to HBM
        self.cuda_envs.reset_all_envs()
        self.cuda_sample_controller = CUDASampler(self.cuda_envs.cuda_function_manager)
        create_and_push_data_placeholders(
            env_wrapper=self.cuda_envs, action_sampler=self.cuda_sample_controller,
            policy_tag_to_agent_id_map=self.policy_tag_to_agent_id_map,
            obs_dim_corresponding_to_num_agents=obs_dim_corresponding_to_num_agents,
            training_batch_size_per_env=self.training_batch_size_per_env)
        seed = int(t.get("seed", np.int32(time.time()))) + self.device_id
        self.seed = seed
        self.cuda_sample_controller.init_random(seed)
        torch.manual_seed(seed)
        random.seed(seed)
        np.random.seed(seed)
        self.cuda_envs.init_reset_pool(seed + random.randint(1, 10000))

        self.total_steps = self.cuda_envs.episode_length * self.num_episodes
        self.num_iters = int(self.total_steps // self.training_batch_size)
        if self.num_iters == 0:
            raise ValueError("Not enough steps to even perform a single training iteration!. "
                             "Please increase the number of episodes or reduce the training "
                             "batch size.")

        self.models, self.optimizers, self.lr_schedules, self.trainers = {}, {}, {}, {}
        self.clip_grad_norm, self.max_grad_norm, self.current_timestep = {}, {}, {}
        for policy in self.policies:
            self.current_timestep[policy] = 0
            self._initialize_policy_model(policy)
        self.load_model_checkpoint()
        for policy in self.policies:
            self.models[policy].cuda()
        # ONE flat float32 arena for the parameters and one for the gradients of every trained
        # policy: Adam / clipping run as flat kernels (utils/flat_adam.py) and the multi-GPU
        # gradient all-reduce is a single in-place NCCL call on `_arena_grads`
        from warp_drive_b200.training.utils.flat_adam import FlatAdam

        total = sum(FlatAdam.numel(self.models[p].parameters()) for p in self.policies)
        dev = next(self.models[self.policies[0]].parameters()).device
        self._arena_params = torch.zeros(total, dtype=torch.float32, device=dev)
        self._arena_grads = torch.zeros(total, dtype=torch.float32, device=dev)
        self._arena_off = 0
        for policy in self.policies:
            self._initialize_optimizer(policy)
        for policy in self.policies_to_train:
            self._initialize_policy_algorithm(policy)
        if self.num_devices > 1:
            self._broadcast_parameters()

        self.engine = RolloutEngine(self.cuda_envs, self.models,
                                    self.policy_tag_to_agent_id_map,
                                    self.cuda_sample_controller,
                                    self.training_batch_size_per_env,
                                    use_cuda_graph=use_cuda_graph)
        self.perf_stats = PerfStats()
        self._flat_grad = None

    # ------------------------------------------------------------------ setup helpers
    def _get_config(self, args):
        cfg = self.config
        for a in args:
            cfg = cfg[a]
        return cfg

    def _initialize_policy_model(self, policy):
        mcfg = self._get_config(["policy", policy, "model"])
        model = ModelFactory.create(mcfg["type"])(
            env=self.cuda_envs, model_config=mcfg, policy=policy,
            policy_tag_to_agent_id_map=self.policy_tag_to_agent_id_map)
        if mcfg.get("init_method") == "xavier":
            for m in model.modules():
                if isinstance(m, nn.Linear):
                    nn.init.xavier_uniform_(m.weight)
        self.models[policy] = model

    def _initialize_optimizer(self, policy):
        self.lr_schedules[policy] = ParamScheduler(self._get_config(["policy", policy, "lr"]))
        lr = self.lr_schedules[policy].get_param_value(self.current_timestep[policy])
        from warp_drive_b200.training.utils.flat_adam import FlatAdam

        n = FlatAdam.numel(self.models[policy].parameters())
        lo = self._arena_off
        self._arena_off += n
        self.optimizers[policy] = FlatAdam(
            self.models[policy].parameters(), lr=lr,
            arena=(self._arena_params[lo:lo + n], self._arena_grads[lo:lo + n]))

    def _initialize_policy_algorithm(self, policy):
        c = self._get_config(["policy", policy])
        assert c["algorithm"] in ("A2C", "PPO")
        self.clip_grad_norm[policy] = c["clip_grad_norm"]
        if c["clip_grad_norm"]:
            self.max_grad_norm[policy] = c["max_grad_norm"]
        kw = dict(discount_factor_gamma=c["gamma"], normalize_advantage=c["normalize_advantage"],
                  normalize_return=c["normalize_return"], vf_loss_coeff=c["vf_loss_coeff"],
                  entropy_coeff=c["entropy_coeff"])
        self.trainers[policy] = A2C(**kw) if c["algorithm"] == "A2C" else PPO(
            clip_param=c["clip_param"], **kw)

    # ------------------------------------------------------------------ multi-GPU
    def _trained_params(self):
        return [p for pol in self.policies_to_train for p in self.models[pol].parameters()]

    def _broadcast_parameters(self):
        import torch.distributed as dist

        for pol in self.policies:
            for p in self.models[pol].parameters():
                dist.broadcast(p.data, src=0)

    def _allreduce_gradients(self):
        """ONE NCCL all-reduce (mean) over the gradients of every policy, in place on the flat
        gradient arena (every .grad is a view into it: no packing, no copies back)."""
        import torch.distributed as dist

        dist.all_reduce(self._arena_grads, op=dist.ReduceOp.SUM)
        self._arena_grads.div_(self.num_devices)

    # ------------------------------------------------------------------ training loop
    def train(self):
        self.cuda_envs.reset_all_envs()
        self.engine.resync_observations()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        n_agents = self.cuda_envs.n_agents
        for iteration in range(self.num_iters):
            t0 = time.time()
            ev[0].record()
            self._generate_rollout_batch()
            ev[1].record()
            metrics = self._update_model_params(iteration)
            ev[2].record()
            torch.cuda.synchronize()
            self.perf_stats.rollout_time += ev[0].elapsed_time(ev[1]) / 1000
            self.perf_stats.training_time += ev[1].elapsed_time(ev[2]) / 1000
            self.perf_stats.iters = iteration + 1
            self.perf_stats.steps = self.perf_stats.iters * self.training_batch_size
            self.perf_stats.agent_steps = self.perf_stats.steps * n_agents
            self.perf_stats.total_time += time.time() - t0
            self._log_metrics(metrics)
            self.save_model_checkpoint(iteration)

    def _get_sample_params(self):
        t = self.current_timestep[self.policies[0]]
        return {k: s.get_param_value(t) for k, s in self.sample_params_schedules.items()}

    def _generate_rollout_batch(self):
        params = self._get_sample_params()
        self.engine.rollout(**params)
        return params

    def _update_model_params(self, iteration):
        saving = self.config["saving"]
        logging_flag = (iteration % saving["metrics_log_freq"] == 0
                        or iteration == self.num_iters - 1)
        dm = self.cuda_envs.cuda_data_manager
        done_flags_batch = dm.data_on_device_via_torch(f"{_DONE_FLAGS}_batch")
        losses, metrics_dict = {}, {}
        for policy in self.policies_to_train:
            actions = dm.data_on_device_via_torch(f"{_ACTIONS}_batch_{policy}")
            rewards = dm.data_on_device_via_torch(f"{_REWARDS}_batch_{policy}")
            obs = dm.data_on_device_via_torch(f"{_PROCESSED_OBSERVATIONS}_batch_{policy}")
            probs, values = self.models[policy](obs=obs)
            self.current_timestep[policy] += self.training_batch_size
            loss, metrics = self.trainers[policy].compute_loss_and_metrics(
                self.current_timestep[policy], actions.long(), rewards, done_flags_batch,
                probs, values, perform_logging=logging_flag,
                negative_positive_ratio=self.config["trainer"].get("neg_pos_env_ratio", -1))
            lr = self.lr_schedules[policy].get_param_value(self.current_timestep[policy])
            for group in self.optimizers[policy].param_groups:
                group["lr"] = lr
            self.optimizers[policy].zero_grad()
            loss.backward()
            losses[policy] = (loss, metrics, lr)
        if self.num_devices > 1 and losses:
            self._allreduce_gradients()
        for policy, (loss, metrics, lr) in losses.items():
            grad_norm = None
            if logging_flag:
                grad_norm = float(sum(p.grad.norm(2) for p in self.models[policy].parameters()
                                      if p.grad is not None))
            # clip_grad_norm_ + Adam as two flat kernels, clip factor stays on the device
            self.optimizers[policy].step(
                max_grad_norm=self.max_grad_norm[policy] if self.clip_grad_norm[policy] else None)
            if logging_flag:
                n_done = int(self.engine.num_completed_episodes)
                metrics.update({
                    "Current timestep": self.current_timestep[policy],
                    "Gradient norm": grad_norm, "Learning rate": lr,
                    "Mean episodic reward":
                        float(self.engine.episodic_reward_sum[policy]) / (n_done + _EPSILON),
                    "Mean episodic steps":
                        float(self.engine.episodic_step_sum) / (n_done + _EPSILON)})
                metrics_dict[policy] = metrics
        self.engine.refresh_forward_weights()
        if logging_flag:
            for policy in self.policies:
                self.engine.episodic_reward_sum[policy].zero_()
            self.engine.episodic_step_sum.zero_()
            self.engine.num_completed_episodes.zero_()
        return metrics_dict

    # ------------------------------------------------------------------ logging / ckpt
    def _log_metrics(self, metrics):
        if not metrics:
            return
        perf = self.perf_stats.get_perf_stats()
        if self.verbose:
            print("\n" + "=" * 40)
            print(f"Device: {self.device_id}")
            print(f"{'Iterations Completed':40}: {self.perf_stats.iters} / {self.num_iters}")
            for k, v in perf.items():
                print(f"{k:40}: {v:10.2f}")
            for policy, m in metrics.items():
                print("=" * 40 + f"\nMetrics for policy '{policy}'\n" + "=" * 40)
                for k, v in m.items():
                    print(f"{k:40}: {v:10.5f}" if isinstance(v, float) else f"{k:40}: {v}")
            print("=" * 40 + "\n")
        logs = {"Iterations Completed": self.perf_stats.iters}
        logs.update(metrics)
        logs["Perf. Stats"] = perf
        fn = f"results_device_{self.device_id}.json" if self.num_devices > 1 else "results.json"
        with open(os.path.join(self.save_dir, fn), "a+", encoding="utf8") as fp:
            json.dump(logs, fp)
            fp.write("\n")

    def save_model_checkpoint(self, iteration=0):
        if self.device_id != 0:
            return
        freq = self.config["saving"]["model_params_save_freq"]
        if iteration % freq == 0 or iteration == self.num_iters - 1:
            for policy, model in self.models.items():
                path = os.path.join(
                    self.save_dir, f"{policy}_{self.current_timestep[policy]}.state_dict")
                if self.verbose:
                    verbose_print(f"Saving the '{policy}' torch model to the file: '{path}'.",
                                  self.device_id)
                torch.save(model.state_dict(), path)

    def load_model_checkpoint(self, ckpts_dict=None):
        if ckpts_dict is None:
            ckpts_dict = {p: self.config["policy"][p]["model"]["model_ckpt_filepath"]
                          for p in self.policies}
        for policy, path in ckpts_dict.items():
            assert policy in self.policies
            if path:
                assert os.path.isfile(path), "Invalid model checkpoint path!"
                self.models[policy].load_state_dict(torch.load(path, map_location="cpu"))
                # the timestep is encoded in the file name ({policy}_{timestep}.state_dict)
                self.current_timestep[policy] = int(
                    os.path.basename(path).split(".state_dict")[0].split("_")[-1])
        # the tensor-core rollout forward keeps its own packed bf16 copy of the weights
        if getattr(self, "engine", None) is not None:
            self.engine.refresh_forward_weights()

    def graceful_close(self):
        self.cuda_sample_controller = None
        if self.verbose:
            verbose_print("Trainer exits gracefully", self.device_id)

    # ------------------------------------------------------------------ evaluation
    def evaluate_episodes(self, **sample_params):
        """Play one episode in every env replica with the current actor (scale=0: no
        exploration noise; for categorical policies pass nothing) and return per-env reward /
        step sums (trainer_base.py:794-846); accumulated on the device, no per-step host
        transfer.  The env states are restarted afterwards."""
        eng = self.engine
        dm = self.cuda_envs.cuda_data_manager
        fused, eng.fused = eng.fused, None     # generic multi-launch step while evaluating
        dev = dm.device
        reward_sum = {p: torch.zeros((self.num_envs, len(ids)), device=dev)
                      for p, ids in self.policy_tag_to_agent_id_map.items()}
        step_sum = {p: torch.zeros(self.num_envs, dtype=torch.int32, device=dev)
                    for p in self.policies}
        self.cuda_envs.reset_all_envs()
        done = dm.data_on_device_via_torch("_done_")
        with torch.no_grad():
            for _ in range(self.cuda_envs.episode_length):
                probs = RolloutEngine.evaluate_policies(eng, -1)
                RolloutEngine.sample_actions(eng, probs, -1, **sample_params)
                self.cuda_envs.step_all_envs()
                undone = done == 0
                rewards = dm.data_on_device_via_torch(_REWARDS)
                for p in self.policies:
                    r_p = rewards if eng.covers_all[p] else rewards.index_select(1, eng.ids[p])
                    reward_sum[p] += r_p * undone[:, None]
                    step_sum[p] += undone.to(torch.int32)
                # done envs restart but keep their flag: one episode per env is counted
                self.cuda_envs.reset_only_done_envs(undo_done_after_reset=False)
        eng.fused = fused
        self.cuda_envs.reset_all_envs()
        eng.resync_observations()
        return reward_sum, step_sum

    # ------------------------------------------------------------------ episode states
    def fetch_episode_states(self, list_of_states=None, env_id=0, include_rewards_actions=False,
                             include_probabilities=False):
        """Play one episode of env `env_id` with the current policies and return the
        per-timestep values of the requested device arrays (trainer_base.py:689-792).

        The episode is logged ON THE DEVICE: every requested array has a [T + 1, ...] ring in
        HBM that `wdb_log_one_step` (the CUDALogController kernel, function_manager.py:295-422)
        fills with env `env_id`'s slice after each step, the done flag of that env is logged
        the same way, and the host pulls everything ONCE after the last step (the reference
        pulls every whole [E, ...] array to the host every timestep)."""
        from warp_drive_b200.managers.function_manager import CUDALogController

        assert 0 <= env_id < self.num_envs
        assert isinstance(list_of_states, list) and list_of_states
        dm = self.cuda_envs.cuda_data_manager
        self.cuda_envs.reset_all_envs()
        self.engine.resync_observations()
        T = self.cuda_envs.episode_length
        logger = CUDALogController(self.cuda_envs.cuda_function_manager)

        def ring(name, steps):
            t = dm.data_on_device_via_torch(name)
            fill = float("nan") if t.dtype.is_floating_point else 0
            return torch.full((steps,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)

        rings = {name: ring(name, T + 1) for name in list_of_states}
        for name in list_of_states:
            logger.log_tensor_one_step(rings[name], dm.data_on_device_via_torch(name), 0, env_id)
        extra = {}
        if include_rewards_actions:
            extra = {_ACTIONS: ring(_ACTIONS, T), _REWARDS: ring(_REWARDS, T)}
        done_ring = torch.zeros(T, dtype=torch.int32, device=dm.device)
        prob_rings = None
        fused, self.engine.fused = self.engine.fused, None
        sa, self.engine.sa = self.engine.sa, None
        try:
            for t in range(T):
                with torch.no_grad():
                    probs = self.engine.evaluate_policies(-1)
                    self.engine.sample_actions(probs, -1)
                    self.cuda_envs.step_all_envs()
                for name in list_of_states:
                    logger.log_tensor_one_step(rings[name], dm.data_on_device_via_torch(name),
                                               t + 1, env_id)
                for name, r in extra.items():
                    logger.log_tensor_one_step(r, dm.data_on_device_via_torch(name), t, env_id)
                logger.log_tensor_one_step(done_ring, dm.data_on_device_via_torch("_done_"), t,
                                           env_id)
                if include_probabilities:
                    if prob_rings is None:
                        prob_rings = [torch.zeros((T,) + tuple(p.shape[1:]), device=p.device)
                                      for p in probs]
                    for r, p in zip(prob_rings, probs):
                        logger.log_tensor_one_step(r, p.contiguous(), t, env_id)
            # ---- the one device -> host transfer of the episode
            done_host = done_ring.cpu().numpy()
            ended = np.nonzero(done_host)[0]
            n_steps = int(ended[0]) + 1 if len(ended) else T
            out = {name: rings[name][: n_steps + 1].cpu().numpy().astype(np.float32)
                   for name in list_of_states}
            for name, r in extra.items():
                out[name] = r[:n_steps].cpu().numpy().astype(np.float32)
            if include_probabilities:
                host = [r[:n_steps].cpu().numpy() for r in prob_rings]
                out["probabilities"] = [[h[t] for h in host] for t in range(n_steps)]
        finally:
            self.engine.fused = fused
            self.engine.sa = sa
            self.cuda_envs.reset_all_envs()
            self.engine.resync_observations()
        return out

    def _fetch_episode_states_host_pull(self, list_of_states=None, env_id=0, include_rewards_actions=False,
                             include_probabilities=False):
        """The reference-shaped implementation (one host pull of every whole [E, ...] array
        per timestep, trainer_base.py:689-792); kept as the checker of fetch_episode_states."""
        assert 0 <= env_id < self.num_envs
        assert isinstance(list_of_states, list) and list_of_states
        dm = self.cuda_envs.cuda_data_manager
        self.cuda_envs.reset_all_envs()
        self.engine.resync_observations()
        T = self.cuda_envs.episode_length
        out = {}
        for name in list_of_states:
            shape = dm.get_shape(name)
            out[name] = np.full((T + 1,) + tuple(shape[1:]), np.nan, dtype=np.float32)
            out[name][0] = dm.pull_data_from_device(name)[env_id]
        if include_rewards_actions:
            out[_ACTIONS] = np.full((T,) + tuple(dm.get_shape(_ACTIONS)[1:]), np.nan, np.float32)
            out[_REWARDS] = np.full((T,) + tuple(dm.get_shape(_REWARDS)[1:]), np.nan, np.float32)
        fused, self.engine.fused = self.engine.fused, None
        try:
            for t in range(T):
                with torch.no_grad():
                    probs = self.engine.evaluate_policies(-1)
                    self.engine.sample_actions(probs, -1)
                    self.cuda_envs.step_all_envs()
                for name in list_of_states:
                    out[name][t + 1] = dm.pull_data_from_device(name)[env_id]
                if include_rewards_actions:
                    out[_ACTIONS][t] = dm.pull_data_from_device(_ACTIONS)[env_id]
                    out[_REWARDS][t] = dm.pull_data_from_device(_REWARDS)[env_id]
                if include_probabilities:
                    out.setdefault("probabilities", []).append(
                        [p[env_id].cpu().numpy() for p in probs])
                if int(dm.pull_data_from_device("_done_")[env_id]):
                    for name in out:
                        if isinstance(out[name], np.ndarray):
                            out[name] = out[name][: t + 2 if name in list_of_states else t + 1]
                    break
        finally:
            self.engine.fused = fused
            self.cuda_envs.reset_all_envs()
            self.engine.resync_observations()
        return out


# names a reference user imports
TrainerA2C = Trainer
TrainerBase = Trainer

"""Actor and action-value critic for DDPG on continuous (Box) action spaces.

Forward contracts and parameter names of the reference's FullyConnectedActor /
FullyConnectedActionValueCritic (warp_drive/training/models/
fully_connected_actor_critic.py:13-144) so its checkpoints load unchanged:
  actor  : `fc.{i}.0`, `policy_head`            forward(obs)         -> [actions per head]
  critic : `fc.{i}.0` (input = obs ++ action), `vf_head`
                                                forward(obs, action) -> Q values [..]
The deterministic head is `output_w * tanh(linear) + output_b` (model config keys of
single_pendulum.yaml / single_continuous_mountain_car.yaml)."""
import torch
from torch import nn

from warp_drive_b200.training.models.fully_connected import FullyConnected


class FullyConnectedActor(FullyConnected):
    name = "torch_fully_connected_actor"

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        del self.vf_head                 # include_value_head=False in the reference

    def forward(self, obs=None, action=None):
        x = obs
        for i in range(len(self.fc)):
            x = self.fc[str(i)](x)
        return self._policy_outputs(x)

    def _policy_outputs(self, x):
        from torch.nn import functional as func

        from warp_drive_b200.training.models.fully_connected import apply_logit_mask

        if self.is_deterministic:
            out = torch.tanh(apply_logit_mask(self.policy_head(x), self.action_mask))
            out = self.action_scale * out + self.action_bias
            if self.output_dims[0] > 1:
                return [t.contiguous() for t in torch.split(out, 1, dim=-1)]
            return [out]
        masks = [None] * len(self.output_dims)
        if self.action_mask is not None:
            start = 0
            for k, dim in enumerate(self.output_dims):
                masks[k] = self.action_mask[..., start:start + dim]
                start += dim
        return [func.softmax(apply_logit_mask(head(x), masks[k]), dim=-1)
                for k, head in enumerate(self.policy_head)]


class FullyConnectedActionValueCritic(FullyConnected):
    name = "torch_fully_connected_q"

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        del self.policy_head             # include_policy_head=False in the reference
        dims = [self.flattened_obs_size + self.flattened_action_size] + self.fc_dims
        self.fc = nn.ModuleDict({
            str(i): nn.Sequential(nn.Linear(dims[i], dims[i + 1]), nn.ReLU())
            for i in range(len(self.fc_dims))})

    def forward(self, obs=None, action=None):
        assert action is not None
        parts = [obs] + (list(action) if isinstance(action, (list, tuple)) else [action])
        x = torch.cat(parts, dim=-1)
        for i in range(len(self.fc)):
            x = self.fc[str(i)](x)
        return self.vf_head(x)[..., 0]


class ActorAsPolicy(nn.Module):
    """Adapter for the rollout engine, whose forward contract is `(probs, values)`."""

    def __init__(self, actor):
        super().__init__()
        self.actor = actor

    def forward(self, obs=None):
        return self.actor(obs), None

"""load_policy / policy_forward with the reference's interface
(crl_kino/policy/rl_policy.py:18-39, :83-109).  The DDPG policy object of the
reference is reduced to what the planning path uses: the actor and its action
range (tianshou ddpg.py:101-131 re-clamps the actor output to the same range)."""
import numpy as np
import torch

from .. import ops
from .net import Actor

device = "cuda"


class DDPGPolicy:
    """Forward-only stand-in for tianshou.policy.DDPGPolicy."""

    def __init__(self, actor, actor_optim=None, critic=None, critic_optim=None, action_range=None, **kw):
        self.actor = actor
        self._range = action_range

    def load_state_dict(self, sd):
        actor_sd = {k[len("actor."):]: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v)
                    for k, v in sd.items() if k.startswith("actor.")}
        self.actor.load_state_dict(actor_sd)

    def forward_device(self, obs32, noise=None, eps=0.0):
        return ops.policy_forward(self.actor.handle, obs32, noise, eps)


def load_policy(env, layer, model_path=None, seed=None):
    """rl_policy.py:18-39.  `model_path=None` builds a seeded random-init actor
    (the reference checkout ships no trained policy.pth)."""
    state_shape = env.observation_space.shape
    action_shape = env.action_space.shape
    action_range = [env.action_space.low, env.action_space.high]
    actor = Actor(layer, state_shape, action_shape, action_range, device, seed=seed)
    policy = DDPGPolicy(actor, None, None, None, action_range=action_range)
    if model_path is not None:
        policy.load_state_dict(torch.load(model_path, map_location="cpu"))
    return policy


def policy_forward(policy, obs, info=None, eps=0.0, noise=None):
    """rl_policy.py:83-109 -> float32 numpy [B, 2].

    The reference draws torch.randn inside the actor; here the standard-normal
    draws are an explicit input (`noise`, [B, 2]); with eps != 0 and no noise
    given they are drawn with torch.randn on the host, as the reference does."""
    obs = np.array(obs)
    obs_len = 1 if obs.ndim == 1 else len(obs)
    obs = obs.reshape((obs_len, -1))
    if noise is None and eps != 0.0:
        noise = torch.randn(size=(obs_len, 2)).numpy()
    act = policy.forward_device(ops.pad_obs(obs), noise, eps)
    return act.cpu().numpy()

"""Base class of the on-policy actors (reference: harl/algorithms/actors/on_policy_base.py:8-137).

Keeps the reference's constructor signature, attributes and methods; the network is a
``DeviceNet`` (flat device-resident parameters + CUDA kernels) instead of an ``nn.Module``.
"""
import numpy as np
import torch

from ... import _lib as L
from ...nets import DeviceNet
from ...utils.envs_tools import get_shape_from_obs_space
from ...utils.models_tools import linear_schedule_lr

_INSTANCES = [0]


def to_device(x, device):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=device, dtype=torch.float32).contiguous()


class OnPolicyBase:
    def __init__(self, args, obs_space, act_space, device=torch.device("cpu")):
        self.args = args
        self.device = torch.device(device)
        self.tpdv = dict(dtype=torch.float32, device=self.device)
        self.data_chunk_length = args["data_chunk_length"]
        self.use_recurrent_policy = args["use_recurrent_policy"]
        self.use_naive_recurrent_policy = args["use_naive_recurrent_policy"]
        self.use_policy_active_masks = args["use_policy_active_masks"]
        self.action_aggregation = args["action_aggregation"]
        self.lr = args["lr"]
        self.opti_eps = args["opti_eps"]
        self.weight_decay = args["weight_decay"]
        self.obs_space = obs_space
        self.act_space = act_space
        obs_shape = get_shape_from_obs_space(obs_space)
        if len(obs_shape) == 3:
            raise NotImplementedError("CNN observation trunks are outside the B200 hot path (SURVEY.md section 2 row 6)")
        kind = act_space.__class__.__name__
        if kind == "Discrete":
            head, out_dim = L.HEAD_DISCRETE, act_space.n
        elif kind == "Box":
            head, out_dim = L.HEAD_BOX, act_space.shape[0]
        else:
            raise NotImplementedError(f"{kind} action spaces are outside the B200 hot path")
        self.actor = DeviceNet(args, obs_shape[0], head, out_dim, self.device)
        self.cur_lr = self.lr
        _INSTANCES[0] += 1
        self._seed = (torch.initial_seed() * 1000003 + _INSTANCES[0]) & (2**63 - 1)
        self._draws = 0

    def lr_decay(self, episode, episodes):
        self.cur_lr = linear_schedule_lr(episode, episodes, self.lr)

    @property
    def recurrent(self):
        return bool(self.use_recurrent_policy or self.use_naive_recurrent_policy)

    def _rnn_passthrough(self, rnn_states_actor):
        return rnn_states_actor if torch.is_tensor(rnn_states_actor) else to_device(rnn_states_actor, self.device)

    def get_actions(self, obs, rnn_states_actor, masks, available_actions=None, deterministic=False,
                    actions_out=None, logp_out=None, rnn_out=None):
        """Sample (or take the mode of) actions for a batch of observations; returns device tensors
        (actions [B, ad], log-probs [B, ad], rnn states [B, recurrent_n, h] -- the GRU's new hidden state for
        recurrent policies, the input passed through otherwise)."""
        obs = to_device(obs, self.device)
        avail = to_device(available_actions, self.device)
        B, w = obs.shape[0], self.actor.act_width
        actions = actions_out if actions_out is not None else torch.empty(B, w, **self.tpdv)
        logp = logp_out if logp_out is not None else torch.empty(B, w, **self.tpdv)
        self._draws += 1
        if self.recurrent:
            rnn_in = to_device(rnn_states_actor, self.device)
            mk = to_device(masks, self.device).reshape(B)
            rnn_new = rnn_out if rnn_out is not None else torch.empty_like(rnn_in)
            self.actor.act(obs, avail, deterministic, self._seed, self._draws, actions, logp, rnn_in, mk, rnn_new)
            return actions, logp, rnn_new
        self.actor.act(obs, avail, deterministic, self._seed, self._draws, actions, logp)
        return actions, logp, self._rnn_passthrough(rnn_states_actor)

    def evaluate_actions(self, obs, rnn_states_actor, action, masks, available_actions=None, active_masks=None):
        """Log-probabilities of given actions under the current policy: (logp [B, ad], None, None).

        The entropy and distribution object the reference also returns are consumed only by the
        loss / KL code, which here lives inside the fused gradient kernels."""
        obs, action = to_device(obs, self.device), to_device(action, self.device)
        avail = to_device(available_actions, self.device)
        logp = torch.empty(obs.shape[0], self.actor.act_width, **self.tpdv)
        if self.recurrent:
            # rnn.py:24: rows == states -> one step per row; otherwise T steps x N sequences, time-major rows
            rnn = to_device(rnn_states_actor, self.device)
            mk = to_device(masks, self.device).reshape(-1)
            seq_len = obs.shape[0] // rnn.shape[0]
            batch = DeviceNet.actor_batch(obs, action, avail=avail, rnn_states=rnn, masks=mk, seq_len=seq_len)
        else:
            batch = DeviceNet.actor_batch(obs, action, avail=avail)
        self.actor.evaluate(batch, logp_out=logp)
        return logp, None, None

    def act(self, obs, rnn_states_actor, masks, available_actions=None, deterministic=False):
        actions, _, rnn = self.get_actions(obs, rnn_states_actor, masks, available_actions, deterministic)
        return actions, rnn

    def update(self, sample):
        pass

    def train(self, actor_buffer, advantages, state_type):
        pass

    def prep_training(self):
        pass

    def prep_rollout(self):
        pass

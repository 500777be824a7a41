"""Per-agent ("feature pruned") critic rollout storage: arrays are [T(+1), N, A, ...]
(reference: harl/common/buffers/on_policy_critic_buffer_fp.py).  Same kernels as the EP buffer -- the
GAE scan simply runs over C = N*A independent columns."""
import torch

from .on_policy_critic_buffer_ep import OnPolicyCriticBufferEP


class OnPolicyCriticBufferFP(OnPolicyCriticBufferEP):
    def __init__(self, args, share_obs_space, num_agents, device=torch.device("cpu")):
        super().__init__(args, share_obs_space, num_agents=num_agents, device=device)

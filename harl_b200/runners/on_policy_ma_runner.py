"""Runner for MAPPO (reference: harl/runners/on_policy_ma_runner.py:7-60): advantages from the GAE kernel, FP global
advantage normalisation, per-agent (or shared-parameter) PPO updates without the sequential factor, then the critic --
enqueued first on a side stream, as in the HA runner."""
import torch

from .. import _lib as L
from .. import dist
from .on_policy_base_runner import OnPolicyBaseRunner


class OnPolicyMARunner(OnPolicyBaseRunner):
    def train(self):
        dev = self.device
        cb = self.critic_buffer
        advantages = cb.advantages  # returns[:-1] - denorm(value_preds[:-1]) (on_policy_ma_runner.py:15-23)
        if self.state_type == "FP":  # :26-35
            active = torch.stack([b.active_masks[:-1] for b in self.actor_buffer], dim=2).contiguous()
            m3 = torch.zeros(3, dtype=torch.float64, device=dev)
            L.call("hb_masked_moments", L.ptr(advantages), L.ptr(active), advantages.numel(), L.ptr(m3), L.stream_ptr())
            dist.all_reduce_sum_(m3)
            adv_n = torch.empty_like(advantages)
            L.call("hb_normalize_by_moments", L.ptr(advantages), L.ptr(adv_n), advantages.numel(), L.ptr(m3), L.stream_ptr())
            advantages = adv_n
        critic_pending = None
        if getattr(self, "overlap_critic_update", True):
            side = getattr(self, "_side_stream", None)
            if side is None:
                side = self._side_stream = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                critic_pending = self.critic.train(cb, self.value_normalizer, defer=True)
        infos = []
        if self.share_param:  # :38-44
            info = self.actor[0].share_param_train(self.actor_buffer, advantages, self.num_agents, self.state_type)
            infos = [info for _ in range(self.num_agents)]
        else:                 # :45-56
            for agent_id in range(self.num_agents):
                adv_a = advantages if self.state_type == "EP" else advantages[:, :, agent_id].contiguous()
                infos.append(self.actor[agent_id].train(self.actor_buffer[agent_id], adv_a, self.state_type))
        if critic_pending is not None:
            torch.cuda.current_stream(dev).wait_stream(self._side_stream)
            critic_info = critic_pending()
        else:
            critic_info = self.critic.train(cb, self.value_normalizer)  # :59
        self.last_agent_order = list(range(self.num_agents))
        return infos, critic_info

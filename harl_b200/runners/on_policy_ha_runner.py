"""Sequential-agent ("HA") update (reference: harl/runners/on_policy_ha_runner.py:8-130).

Agents are updated one after another in a random order; agent i's clip loss is re-weighted by
``factor`` = the running product of the importance ratios of the agents updated before it.
Everything stays on the device: advantages come from the GAE kernel, the old / new log-prob
sweeps are forward-only kernel passes over the buffer (the second one multiplies ``factor`` in
place), and the critic update follows.
"""
import os

import torch

from .. import _lib as L
from .. import dist
from ..nets import DeviceNet
from .on_policy_base_runner import OnPolicyBaseRunner


class OnPolicyHARunner(OnPolicyBaseRunner):
    def train(self):
        T = self.algo_args["train"]["episode_length"]
        N = self.n_local
        dev = self.device
        rows = T * N
        cb = self.critic_buffer
        factor = torch.ones(T, N, 1, dtype=torch.float32, device=dev)
        advantages = cb.advantages  # returns[:-1] - denorm(value_preds[:-1]), written by hb_gae_returns (:26-33)
        if self.state_type == "FP":  # global masked normalisation across agents (:36-45)
            active = torch.stack([b.active_masks[:-1] for b in self.actor_buffer], dim=2).contiguous()
            m3 = torch.zeros(3, dtype=torch.float64, device=dev)
            L.call("hb_masked_moments", L.ptr(advantages), L.ptr(active), advantages.numel(), L.ptr(m3), L.stream_ptr())
            dist.all_reduce_sum_(m3)
            adv_n = torch.empty_like(advantages)
            L.call("hb_normalize_by_moments", L.ptr(advantages), L.ptr(adv_n), advantages.numel(), L.ptr(m3), L.stream_ptr())
            advantages = adv_n
        if self.fixed_order:
            agent_order = list(range(self.num_agents))
        else:
            agent_order = list(torch.randperm(self.num_agents).numpy())
        self.last_agent_order = agent_order
        if self.world > 1 and not self.fixed_order:   # the gradient exchanges pair up agents by position in this order
            o = torch.tensor([int(a) for a in agent_order], dtype=torch.int64, device=dev)
            o0 = o.clone()
            torch.distributed.broadcast(o0, src=0)
            assert torch.equal(o, o0), "agent order differs across ranks (unsynchronised torch CPU generator)"
        infos = []
        agg_prod = self.action_aggregation == "prod"
        # The critic update (reference :128) shares no state with the actor updates: enqueue it first on a side
        # stream so that its (latency-bound, under-filling) kernels overlap the sequential actor updates.
        critic_pending = None
        overlap = getattr(self, "overlap_critic_update", True)
        # Two ways to overlap the critic's updates with the sequential actor updates (side stream):
        #  * world == 1: enqueue the whole critic update first and let the hardware scheduler interleave;
        #  * world > 1 (default there; HB_CRITIC_INTERLEAVE=0/1 forces): a FIXED interleaving -- the big (whole-GPU, persistent)
        #    kernels run strictly in the order A1 C1 A2 A3 A4 C2 ..., pinned by events, while each stream's small kernels
        #    (slot reduce, exchange, Adam, pack) overlap the other stream's big kernel.  With free-running streams the two
        #    ranks of an exchange can pick different orders, and every mismatch parks one rank's actor chain behind a whole
        #    critic kernel of the other (measured at 2 GPUs: the side stream gained nothing over a single stream).
        inter = os.environ.get("HB_CRITIC_INTERLEAVE")
        interleave = overlap and all(hasattr(a, "first_epoch_logp") for a in self.actor) and \
            (inter == "1" or (inter is None and self.world > 1))
        main = torch.cuda.current_stream(dev)
        if overlap:
            side = getattr(self, "_side_stream", None)
            if side is None:
                side = self._side_stream = torch.cuda.Stream(device=dev)
            side.wait_stream(main)
        if overlap and not interleave:
            with torch.cuda.stream(side):
                critic_pending = self.critic.train(cb, self.value_normalizer, defer=True)
        ch = None
        if interleave:
            ch = dict(gen=self.critic.train_steps(cb, self.value_normalizer), k=0, wait=None, done=False)

            def after_critic_big():            # on the side stream, right after the critic's gradient kernels
                ev = torch.cuda.Event()
                ev.record(side)
                ch["wait"] = ev

            def before_actor_big():            # the next actor big kernel goes after the critic big kernel in flight
                if ch["wait"] is not None:
                    main.wait_event(ch["wait"])
                    ch["wait"] = None

            def after_actor_big():             # after A1, A4, A7, ...: one critic update, behind this actor kernel
                ch["k"] += 1
                if not ch["done"] and ch["k"] % 3 == 1:
                    ev = torch.cuda.Event()
                    ev.record(main)
                    with torch.cuda.stream(side):
                        side.wait_event(ev)
                        if next(ch["gen"], None) is None:
                            ch["done"] = True

            self.critic.after_grad_hook = after_critic_big
            for a in self.actor:
                a.grad_hooks = (before_actor_big, after_actor_big)
        # masked advantage moments of every agent (happo.py:119-127: nanmean / nanstd over the agent's active steps) do not
        # depend on the sequential updates: one [A, 3] bucket, one exchange, one host read instead of one per agent
        adv_m3 = torch.zeros(self.num_agents, 3, dtype=torch.float64, device=dev)
        for a in range(self.num_agents):
            adv_a = advantages if self.state_type == "EP" else advantages[:, :, a].contiguous()
            L.call("hb_masked_moments", L.ptr(adv_a), L.ptr(self.actor_buffer[a].active_masks[:-1]), rows, L.ptr(adv_m3[a]),
                   L.stream_ptr())
        dist.all_reduce_sum_(adv_m3)
        n_active = adv_m3[:, 2].cpu().numpy()
        for agent_id in agent_order:
            buf, actor = self.actor_buffer[agent_id], self.actor[agent_id]
            buf.update_factor(factor)
            fl = lambda a: a.reshape(rows, *a.shape[2:])
            avail = None if buf.available_actions is None else fl(buf.available_actions[:-1])
            if actor.recurrent:  # full-buffer evaluate from rnn_states[0] (on_policy_ha_runner.py:66-83)
                sweep = DeviceNet.actor_batch(fl(buf.obs[:-1]), fl(buf.actions), avail=avail,
                                              rnn_states=buf.rnn_states.reshape((T + 1) * N, -1),
                                              masks=buf.masks.reshape((T + 1) * N), seq_len=T)
            else:
                sweep = DeviceNet.actor_batch(fl(buf.obs[:-1]), fl(buf.actions), avail=avail)
            old_logp = torch.empty(rows, actor.actor.act_width, dtype=torch.float32, device=dev)
            adv_a = advantages if self.state_type == "EP" else advantages[:, :, agent_id].contiguous()
            # :66-83 evaluates the buffer under the pre-update weights = the weights of the first PPO epoch, whose forward
            # (same rows, same order, when the minibatch is the whole buffer) writes those log-probs on its way
            fuse = (getattr(actor, "first_epoch_logp", False) and not actor.recurrent and getattr(actor, "actor_num_mini_batch", 0) == 1
                    and n_active[agent_id] > 0 and getattr(self, "fuse_old_logp", True))
            if not fuse:
                if ch is not None:
                    before_actor_big()
                actor.actor.evaluate(sweep, logp_out=old_logp)                   # :66-83
                if ch is not None:
                    after_actor_big()
                infos.append(actor.train(buf, adv_a, self.state_type, moments=(adv_m3[agent_id], float(n_active[agent_id]))))  # :86-93
            else:
                infos.append(actor.train(buf, adv_a, self.state_type, moments=(adv_m3[agent_id], float(n_active[agent_id])),
                                         old_logp_out=old_logp))
                assert actor.old_logp_filled
            if ch is not None:
                before_actor_big()
            actor.actor.evaluate(sweep, logp_ref=old_logp, factor_inout=factor.reshape(rows), agg_prod=agg_prod)  # :96-124
            if ch is not None:
                after_actor_big()
        if ch is not None:                     # critic updates the actor loop had no slot for, then the hooks come off
            with torch.cuda.stream(side):
                for _ in ch["gen"]:
                    pass
            self.critic.after_grad_hook = None
            for a in self.actor:
                a.grad_hooks = None
            critic_pending = self.critic.finish_train
        if critic_pending is not None:
            torch.cuda.current_stream(dev).wait_stream(self._side_stream)
            critic_info = critic_pending()
        else:
            critic_info = self.critic.train(cb, self.value_normalizer)           # :128
        # per-agent infos are reported in agent-id order
        ordered = [None] * self.num_agents
        for pos, agent_id in enumerate(agent_order):
            ordered[agent_id] = infos[pos]
        return ordered, critic_info

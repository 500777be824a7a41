"""MAPPO actor (reference: harl/algorithms/actors/mappo.py:10-222).

The same device update as HAPPO with the importance-ratio factor fixed at 1 (the buffer's ``factor`` stays ``None``, the
kernel then skips it).  ``share_param_train`` updates ONE shared actor on the concatenation of every agent's minibatch:
gradients are linear in the rows, so each agent's buffer contributes one fused forward / loss / backward pass
(hb_ppo_actor_grad with the GLOBAL normaliser) and the per-agent gradient buffers are summed before clip + Adam.
"""
import torch

from ... import _lib as L
from ... import dist
from ...common import seq_index
from ...nets import DeviceNet
from .happo import HAPPO
from .on_policy_base import to_device


class MAPPO(HAPPO):
    def update(self, sample):
        """Reference-compatible single update on a materialised 8-tuple minibatch (mappo.py:28-93)."""
        (obs, rnn, actions, masks, active, old_lp, adv, avail) = sample
        ones = torch.ones(to_device(adv, self.device).reshape(-1).shape[0], 1, device=self.device)
        return super().update((obs, rnn, actions, masks, active, old_lp, adv, avail, ones))

    def share_param_train(self, actor_buffer, advantages, num_agents, state_type):
        """Reference mappo.py:149-222 (feed-forward policies).

        With a recurrent policy the reference concatenates the agents' step-major chunk batches along axis 0 and then
        lets RNNLayer re-read the result as [L, A * mb] (mappo.py:207-216, rnn.py:33-41): row (l, j) of agent a lands
        at step (a * L * mb + l * mb + j) // (A * mb) of an unrelated sequence, so hidden states, observations and
        masks of different agents and chunks are interleaved.  That layout cannot be expressed as an in-place index
        into one agent's buffer; rather than silently computing something else, this combination fails loudly."""
        info = dict(policy_loss=0.0, dist_entropy=0.0, actor_grad_norm=0.0, ratio=0.0)
        if self.recurrent:
            raise NotImplementedError("share_param with recurrent (GRU) policies is not implemented: the reference "
                                      "interleaves the agents' sequences when it concatenates their minibatches")
        d = self.device
        T, N = actor_buffer[0].actions.shape[:2]
        rows = T * N
        adv_all = to_device(advantages, d)
        fl = lambda a: a.reshape(rows, *a.shape[2:])
        actives = [b.active_masks[:-1].reshape(rows) for b in actor_buffer]
        if state_type == "EP":
            # nanmean / nanstd over the agents' stacked copies of the advantages, masked by each agent's active mask
            adv = adv_all.reshape(rows)
            m3 = torch.zeros(3, dtype=torch.float64, device=d)
            for a in range(num_agents):
                L.call("hb_masked_moments", L.ptr(adv), L.ptr(actives[a]), rows, L.ptr(m3), L.stream_ptr())
            dist.all_reduce_sum_(m3)
            adv_n = torch.empty_like(adv)
            L.call("hb_normalize_by_moments", L.ptr(adv), L.ptr(adv_n), rows, L.ptr(m3), L.stream_ptr())
            advs = [adv_n] * num_agents
        else:
            advs = [adv_all[:, :, a].contiguous().reshape(rows) for a in range(num_agents)]
        nmb = self.actor_num_mini_batch
        n_up = self.ppo_epoch * nmb
        scal = torch.zeros(n_up, 4, dtype=torch.float64, device=d)
        gnorm = torch.zeros(n_up, dtype=torch.float32, device=d)
        norms = torch.zeros(n_up, 3, dtype=torch.float64, device=d)
        acc = torch.zeros_like(self.actor.grad)
        mode = "ff"
        rnns, masks = [None] * num_agents, [None] * num_agents
        u = 0
        for _ in range(self.ppo_epoch):
            # one generator per agent (mappo.py:179-205): independent permutations, same minibatch sizes
            gens = [list(seq_index.minibatches(T, N, nmb, mode, self.data_chunk_length, d)) for _ in range(num_agents)]
            for i in range(nmb):
                parts = [g[i] for g in gens]
                # global normaliser of the concatenated batch
                for a in range(num_agents):
                    idx, nrows, _ = parts[a]
                    if self.use_policy_active_masks:
                        norms[u, 2] += (actives[a] if idx is None else actives[a][idx.long()]).sum().double()
                    else:
                        norms[u, 2] += float(nrows)
                dist.all_reduce_sum_(norms[u])
                acc.zero_()
                for a, b in enumerate(actor_buffer):
                    idx, nrows, seq_len = parts[a]
                    avail = None if b.available_actions is None else fl(b.available_actions[:-1])
                    batch = DeviceNet.actor_batch(fl(b.obs[:-1]), fl(b.actions), fl(b.action_log_probs), advs[a], None,
                                                  actives[a], avail, idx, nrows, rnn_states=rnns[a], masks=masks[a],
                                                  seq_len=seq_len)
                    self.actor.actor_grad(batch, self._hyper(), norms[u], scal[u])
                    acc += self.actor.grad
                self.actor.grad.copy_(acc)
                dist.all_reduce_sum_(self.actor.grad)
                self.actor.adam_step(self.cur_lr, self.opti_eps, self.weight_decay, self.max_grad_norm, self.use_max_grad_norm)
                gnorm[u] = self.actor.grad_norm[0]
                u += 1
        dist.all_reduce_sum_(scal)
        s = scal.cpu().numpy()
        nr = norms[:, 2].cpu().numpy()
        info["policy_loss"] = float((s[:, 0] / nr).mean())
        info["dist_entropy"] = float((s[:, 1] / nr).mean())
        info["ratio"] = float((s[:, 2] / s[:, 3]).mean())
        info["actor_grad_norm"] = float(gnorm.mean().item())
        return info

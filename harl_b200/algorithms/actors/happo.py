"""HAPPO actor (reference: harl/algorithms/actors/happo.py:10-158).

``train`` runs entirely on the device: masked advantage normalisation, then
``ppo_epoch x actor_num_mini_batch`` updates, each = fused forward + clip loss (x the running
importance-ratio ``factor``) + entropy + backward (hb_ppo_actor_grad), a sum-allreduce of the
flat gradient when the rollout is sharded over GPUs, and clip-norm + Adam (hb_clip_adam_step).
Scalars come back once per call, not per minibatch.
"""
import torch

from ... import _lib as L
from ... import dist
from ...common import seq_index
from ...nets import DeviceNet
from .on_policy_base import OnPolicyBase, to_device


class HAPPO(OnPolicyBase):
    use_clip = True  # HAA2C switches the clipping off

    def __init__(self, args, obs_space, act_space, device=torch.device("cpu")):
        super().__init__(args, obs_space, act_space, device)
        self.clip_param = args["clip_param"]
        self.ppo_epoch = args.get("ppo_epoch", args.get("a2c_epoch"))
        self.actor_num_mini_batch = args["actor_num_mini_batch"]
        self.entropy_coef = args["entropy_coef"]
        self.use_max_grad_norm = args["use_max_grad_norm"]
        self.max_grad_norm = args["max_grad_norm"]

    first_epoch_logp = True   # train(old_logp_out=...) is supported

    def _hyper(self):
        return L.PPOHyper(float(self.clip_param), float(self.entropy_coef), int(bool(self.use_policy_active_masks)),
                          int(self.action_aggregation == "prod"), int(self.use_clip))

    def _step(self, batch, norm3, scalars_row, logp_out=None):
        """One update on a device batch: grad -> (allreduce) -> clip + Adam. Returns nothing (async)."""
        hooks = getattr(self, "grad_hooks", None)   # (before, after) the big kernel: the runner's stream choreography
        if hooks is not None:
            hooks[0]()
        self.actor.actor_grad(batch, self._hyper(), norm3, scalars_row, logp_out)
        if hooks is not None:
            hooks[1]()
        dist.all_reduce_sum_(self.actor.grad)
        self.actor.adam_step(self.cur_lr, self.opti_eps, self.weight_decay, self.max_grad_norm, self.use_max_grad_norm)

    def update(self, sample):
        """Reference-compatible single update on a materialised minibatch tuple (happo.py:28-102).

        Returns (policy_loss, dist_entropy, actor_grad_norm, imp_weights_mean) as floats."""
        (obs, _rnn, actions, _masks, active, old_lp, adv, avail, factor) = sample
        d = self.device
        obs, actions, active, old_lp, adv, factor, avail = (to_device(x, d) for x in
                                                             (obs, actions, active, old_lp, adv, factor, avail))
        rnn = mk = None
        seq_len = 0
        if self.recurrent:  # the reference's generator output: states at the sequence starts, step-major rows
            rnn = to_device(_rnn, d)
            rnn = rnn.reshape(rnn.shape[0], -1)
            mk = to_device(_masks, d).reshape(-1)
            seq_len = obs.shape[0] // rnn.shape[0]
        batch = DeviceNet.actor_batch(obs, actions, old_lp, adv.reshape(-1), factor.reshape(-1), active.reshape(-1), avail,
                                      rnn_states=rnn, masks=mk, seq_len=seq_len)
        norm3 = torch.zeros(3, dtype=torch.float64, device=d)
        norm3[2] = active.sum().double() if self.use_policy_active_masks else float(obs.shape[0])
        dist.all_reduce_sum_(norm3)
        scal = torch.zeros(4, dtype=torch.float64, device=d)
        self._step(batch, norm3, scal)
        dist.all_reduce_sum_(scal)
        s, n = scal.cpu().numpy(), norm3[2].item()
        return s[0] / n, s[1] / n, self.actor.grad_norm.item(), s[2] / s[3]

    def train(self, actor_buffer, advantages, state_type, moments=None, old_logp_out=None):
        """Reference happo.py:104-158.  ``advantages``: [T, N, 1] (tensor on the device or NumPy).

        ``old_logp_out`` [T*N, ad]: filled with the log-probs of the buffer's actions under the PRE-update weights by the
        forward of the first epoch (whole-buffer minibatch only; see ``first_epoch_logp``) -- what the sequential-update
        runner otherwise obtains from a separate evaluate sweep (on_policy_ha_runner.py:66-83).  ``self.old_logp_filled``
        says whether that happened."""
        self.old_logp_filled = False
        info = dict(policy_loss=0.0, dist_entropy=0.0, actor_grad_norm=0.0, ratio=0.0)
        d = self.device
        buf = actor_buffer
        T, N = buf.actions.shape[:2]
        rows = T * N
        adv = to_device(advantages, d).reshape(rows)
        active = buf.active_masks[:-1].reshape(rows)
        if moments is not None:   # (device double[3], host count): computed for all agents at once by the runner
            m3, n_active = moments
        else:
            m3 = torch.zeros(3, dtype=torch.float64, device=d)
            L.call("hb_masked_moments", L.ptr(adv), L.ptr(active), rows, L.ptr(m3), L.stream_ptr())
            dist.all_reduce_sum_(m3)
            n_active = m3[2].item()  # the one host read before the update loop (reference early-out, happo.py:119)
        if n_active == 0:
            return info
        if state_type == "EP":
            adv_n = torch.empty_like(adv)
            L.call("hb_normalize_by_moments", L.ptr(adv), L.ptr(adv_n), rows, L.ptr(m3), L.stream_ptr())
            adv = adv_n
        nmb = self.actor_num_mini_batch
        n_up = self.ppo_epoch * nmb
        scal = torch.zeros(n_up, 4, dtype=torch.float64, device=d)
        gnorm = torch.zeros(n_up, dtype=torch.float32, device=d)
        norms = torch.zeros(n_up, 3, dtype=torch.float64, device=d)
        fl = lambda a: a.reshape(rows, *a.shape[2:])
        obs, actions, old_lp = fl(buf.obs[:-1]), fl(buf.actions), fl(buf.action_log_probs)
        avail = None if buf.available_actions is None else fl(buf.available_actions[:-1])
        factor = None if buf.factor is None else buf.factor.reshape(rows)
        # recurrent policies: the kernels also read the stored hidden states and the reset masks in place
        rnn = buf.rnn_states.reshape((T + 1) * N, -1) if self.recurrent else None
        masks = buf.masks.reshape((T + 1) * N) if self.recurrent else None
        mode = seq_index.mode_of(self.use_recurrent_policy, self.use_naive_recurrent_policy)
        world = dist.world_size()
        u = 0
        for _ in range(self.ppo_epoch):
            # a single feed-forward / naive minibatch is the whole buffer: the shuffle only reorders a mean, skip it
            for idx, nrows, seq_len in seq_index.minibatches(T, N, nmb, mode, self.data_chunk_length, d):
                if idx is None or nrows == rows:
                    norms[u, 2] = n_active if self.use_policy_active_masks else float(rows * world)
                else:
                    if self.use_policy_active_masks:
                        norms[u, 2] = active[idx.long()].sum().double()
                    else:
                        norms[u, 2] = float(nrows)
                    dist.all_reduce_sum_(norms[u])
                batch = DeviceNet.actor_batch(obs, actions, old_lp, adv, factor, active, avail, idx, nrows,
                                              rnn_states=rnn, masks=masks, seq_len=seq_len)
                first = u == 0 and old_logp_out is not None and (idx is None) and nrows == rows
                self._step(batch, norms[u], scal[u], old_logp_out if first else None)
                self.old_logp_filled = self.old_logp_filled or first
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

"""Shared V-critic (reference: harl/algorithms/critics/v_critic.py:14-208).

``train`` = ``critic_epoch x critic_num_mini_batch`` updates on the device, each: ValueNorm
statistics update from the batch returns (BEFORE normalising them, v_critic.py:90-95), fused
forward + clipped Huber/MSE value loss + backward (hb_value_grad), gradient sum-allreduce over
ranks (the "shared V-critic gradients" exchange of BASELINE.json), clip-norm + Adam.
"""
import torch

from ... import _lib as L
from ... import dist
from ...common import seq_index
from ...nets import DeviceNet
from ...utils.envs_tools import get_shape_from_obs_space
from ...utils.models_tools import linear_schedule_lr
from ..actors.on_policy_base import to_device


class VCritic:
    def __init__(self, args, cent_obs_space, device=torch.device("cpu")):
        self.args = args
        self.device = torch.device(device)
        self.tpdv = dict(dtype=torch.float32, device=self.device)
        self.clip_param = args["clip_param"]
        self.critic_epoch = args["critic_epoch"]
        self.critic_num_mini_batch = args["critic_num_mini_batch"]
        self.data_chunk_length = args["data_chunk_length"]
        self.value_loss_coef = args["value_loss_coef"]
        self.max_grad_norm = args["max_grad_norm"]
        self.huber_delta = args["huber_delta"]
        self.use_recurrent_policy = args["use_recurrent_policy"]
        self.use_naive_recurrent_policy = args["use_naive_recurrent_policy"]
        self.use_max_grad_norm = args["use_max_grad_norm"]
        self.use_clipped_value_loss = args["use_clipped_value_loss"]
        self.use_huber_loss = args["use_huber_loss"]
        self.use_policy_active_masks = args["use_policy_active_masks"]  # stored, never used (as in the reference)
        self.critic_lr = args["critic_lr"]
        self.opti_eps = args["opti_eps"]
        self.weight_decay = args["weight_decay"]
        self.share_obs_space = cent_obs_space
        shp = get_shape_from_obs_space(cent_obs_space)
        if len(shp) == 3:
            raise NotImplementedError("CNN state trunks are outside the B200 hot path")
        self.critic = DeviceNet(args, shp[0], L.HEAD_VALUE, 1, self.device)
        self.cur_lr = self.critic_lr

    def lr_decay(self, episode, episodes):
        self.cur_lr = linear_schedule_lr(episode, episodes, self.critic_lr)

    @property
    def recurrent(self):
        return bool(self.use_recurrent_policy or self.use_naive_recurrent_policy)

    def get_values(self, cent_obs, rnn_states_critic, masks, values_out=None, rnn_out=None):
        """Value predictions [B, 1] and rnn states (the GRU's new hidden state for a recurrent critic, the input
        passed through otherwise), as device tensors."""
        x = to_device(cent_obs, self.device)
        v = values_out if values_out is not None else torch.empty(x.shape[0], 1, **self.tpdv)
        if self.recurrent:
            rnn_in = to_device(rnn_states_critic, self.device)
            mk = to_device(masks, self.device).reshape(x.shape[0])
            rnn_new = rnn_out if rnn_out is not None else torch.empty_like(rnn_in)
            self.critic.values(x, v, rnn_in, mk, rnn_new)
            return v, rnn_new
        self.critic.values(x, v)
        rnn = rnn_states_critic if torch.is_tensor(rnn_states_critic) else to_device(rnn_states_critic, self.device)
        return v, rnn

    def _hyper(self):
        return L.ValueHyper(float(self.clip_param), float(self.huber_delta), float(self.value_loss_coef),
                            int(bool(self.use_huber_loss)), int(bool(self.use_clipped_value_loss)))

    def _step(self, share_obs, value_preds, returns, index, rows, global_rows, value_normalizer, scalars_row,
              rnn_states=None, masks=None, seq_len=0, whole=None):
        d = self.device
        if value_normalizer is not None:
            # ValueNorm.update(return_batch) runs in every update (v_critic.py:93-96 via cal_value_loss); when the minibatch
            # is the whole buffer its batch moments are the same in every epoch: computed and exchanged once per train()
            m3 = whole.get("m3") if whole is not None and index is None else None
            if m3 is None:
                m3 = torch.zeros(3, dtype=torch.float64, device=d)
                src = returns if index is None else returns[index.long()]
                L.call("hb_masked_moments", L.ptr(src.contiguous()), None, rows, L.ptr(m3), L.stream_ptr())
                dist.all_reduce_sum_(m3)
                if whole is not None and index is None:
                    whole["m3"] = m3
            value_normalizer.update_from_moments(m3)
        cb = DeviceNet.critic_batch(share_obs, value_preds, returns, index, rows, rnn_states, masks, seq_len)
        vn = value_normalizer.state if value_normalizer is not None else None
        self.critic.value_grad(cb, self._hyper(), vn, 1.0 / global_rows, scalars_row)
        if getattr(self, "after_grad_hook", None) is not None:   # the runner's stream choreography (on_policy_ha_runner.py)
            self.after_grad_hook()
        dist.all_reduce_sum_(self.critic.grad)
        self.critic.adam_step(self.cur_lr, self.opti_eps, self.weight_decay, self.max_grad_norm, self.use_max_grad_norm)

    def update(self, sample, value_normalizer=None):
        """Reference-compatible single update on a materialised minibatch tuple (v_critic.py:116-157)."""
        share_obs, _rnn, value_preds, returns, _masks = sample
        d = self.device
        so, vp, rt = (to_device(x, d) for x in (share_obs, value_preds, returns))
        rows = so.shape[0]
        scal = torch.zeros(4, dtype=torch.float64, device=d)
        rnn = mk = None
        seq_len = 0
        if self.recurrent:  # the reference's generator output: states at the sequence starts, step-major rows
            rnn = to_device(_rnn, d)
            rnn = rnn.reshape(rnn.shape[0], -1)
            mk = to_device(_masks, d).reshape(-1)
            seq_len = rows // rnn.shape[0]
        self._step(so, vp.reshape(-1), rt.reshape(-1), None, rows, float(rows * dist.world_size()), value_normalizer, scal,
                   rnn, mk, seq_len)
        dist.all_reduce_sum_(scal)
        s = scal.cpu().numpy()
        return s[0] / s[1], self.critic.grad_norm.item()

    def train_steps(self, critic_buffer, value_normalizer=None):
        """Generator form of ``train``: every ``next()`` enqueues ONE update (one minibatch of one epoch) on the current
        stream; after exhaustion ``self.finish_train()`` returns the train-info.  Lets the runner interleave the critic's
        updates with the sequential actor updates in an order that is the same on every rank."""
        d = self.device
        buf = critic_buffer
        T = buf.episode_length
        rows = buf.value_preds[:-1].numel()
        Cn = rows // T  # N envs, or N * A (env, agent) pairs for the FP critic
        so = buf.share_obs[:-1].reshape(rows, -1)
        vp = buf.value_preds[:-1].reshape(rows)
        rt = buf.returns[:-1].reshape(rows)
        rnn = buf.rnn_states_critic.reshape((T + 1) * Cn, -1) if self.recurrent else None
        masks = buf.masks.reshape((T + 1) * Cn) if self.recurrent else None
        mode = seq_index.mode_of(self.use_recurrent_policy, self.use_naive_recurrent_policy)
        nmb = self.critic_num_mini_batch
        n_up = self.critic_epoch * nmb
        scal = torch.zeros(n_up, 4, dtype=torch.float64, device=d)
        gnorm = torch.zeros(n_up, dtype=torch.float32, device=d)
        u = 0
        whole = {}
        for _ in range(self.critic_epoch):
            for idx, n, seq_len in seq_index.minibatches(T, Cn, nmb, mode, self.data_chunk_length, d):
                self._step(so, vp, rt, idx, n, float(n * dist.world_size()), value_normalizer, scal[u], rnn, masks, seq_len, whole)
                gnorm[u] = self.critic.grad_norm[0]
                u += 1
                yield u
        dist.all_reduce_sum_(scal)

        def finish():
            s = scal.cpu().numpy()
            return dict(value_loss=float((s[:, 0] / s[:, 1]).mean()), critic_grad_norm=float(gnorm.mean().item()))

        self.finish_train = finish

    def train(self, critic_buffer, value_normalizer=None, defer=False):
        """Reference v_critic.py:159-200.  ``defer=True`` enqueues the whole update without a host read and returns
        a closure producing the train-info (the HA runner runs the critic update on a side stream, overlapped with
        the sequential actor updates, and reads the scalars after joining the streams)."""
        for _ in self.train_steps(critic_buffer, value_normalizer):
            pass
        return self.finish_train if defer else self.finish_train()

    def prep_training(self):
        pass

    def prep_rollout(self):
        pass

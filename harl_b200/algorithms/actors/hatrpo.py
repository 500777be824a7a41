"""HATRPO actor (reference: harl/algorithms/actors/hatrpo.py:18-247, harl/utils/trpo_util.py).

One ``update`` = surrogate gradient -> 10 conjugate-gradient steps on (F + 0.1 I) x = g -> step scaling by the KL
threshold -> backtracking line search on the surrogate and KL(old || new).  Everything that touches the batch runs
in CUDA over the buffer rows in place (no minibatch is materialised):

* gradient: the fused HAPPO kernel with clipping and the entropy bonus off (it differentiates -loss);
* Fisher-vector product: Gauss-Newton form J^T H J v -- a tangent pass through the trunk (trpo.cu) and, for
  recurrent policies, through the GRU and its LayerNorm (rnn.cu: rnn_jvp_forward), the distribution-space Hessian, and
  the ordinary backward kernels (BPTT included) -- instead of the reference's double backward (identical operator: at
  new == old the KL gradient w.r.t. the distribution parameters vanishes); products 2..11 of an update reuse the
  forward activations the first one left in the workspace;
* recurrent policies: the single minibatch is every data_chunk_length chunk (hatrpo.py:223-227), read in place through
  the chunk index map with the hidden state stored at each chunk start;
* the old distribution is evaluated once and kept ([rows, out_dim]) rather than re-instantiating an old actor;
* CG vector algebra stays on the device (no host sync inside the 10 iterations); the line search reads back four
  doubles per trial, as the reference does (`.cpu().numpy()`, hatrpo.py:163).

With the rollout sharded over GPUs the gradient, every Fisher-vector product and the line-search sums are
sum-allreduced, so all ranks take the identical step.
"""
import numpy as np
import torch

from ... import _lib as L
from ... import dist
from ...common import seq_index
from ...nets import DeviceNet
from .on_policy_base import OnPolicyBase, to_device


class HATRPO(OnPolicyBase):
    def __init__(self, args, obs_space, act_space, device=torch.device("cpu")):
        assert act_space.__class__.__name__ != "MultiDiscrete", \
            "only continuous and discrete action space is supported by HATRPO."
        super().__init__(args, obs_space, act_space, device)
        self.kl_threshold = args["kl_threshold"]
        self.ls_step = args["ls_step"]
        self.accept_ratio = args["accept_ratio"]
        self.backtrack_coeff = args["backtrack_coeff"]
        self.last_update = {}

    def _hyper(self):
        return L.PPOHyper(0.0, 0.0, int(bool(self.use_policy_active_masks)), int(self.action_aggregation == "prod"), 0)

    def _update_on_batch(self, batch, norm, global_rows):
        """hatrpo.py:37-194 on a device batch.  ``norm``: global sum(active) (or global row count)."""
        net, d = self.actor, self.device
        n = net.total
        st = L.stream_ptr()
        f32 = dict(dtype=torch.float32, device=d)
        f64 = dict(dtype=torch.float64, device=d)
        hyper = self._hyper()
        # --- surrogate gradient at theta_old (hatrpo.py:67-98)
        norm3 = torch.zeros(3, **f64)
        norm3[2] = norm
        scal = torch.zeros(4, **f64)
        net.actor_grad(batch, hyper, norm3, scal)
        dist.all_reduce_sum_(net.grad)
        dist.all_reduce_sum_(scal)
        g = net.grad.clone()
        L.call("hb_vec_scale", L.ptr(g), -1.0, n, st)  # the kernel differentiates -loss
        # --- old distribution, once (trpo_util.py:79-82)
        old_dist = torch.empty(batch.rows, net.out_dim, **f32)
        net.trpo_old_dist(batch, old_dist)
        inv_rows = 1.0 / float(global_rows)

        n_fvp = [0]

        def fvp(vec, out):
            # products 2..11 reuse the forward activations the first one left in the workspace (nothing else touches
            # this stream's workspace in between: CG steps and the allreduce work on the flat vectors only)
            net.trpo_fvp(batch, old_dist, vec, inv_rows, out, reuse_forward=n_fvp[0] > 0)
            n_fvp[0] += 1
            dist.all_reduce_sum_(out)
            net.trpo_fvp_finish(vec, out, 0.1)

        # --- conjugate gradient, 10 steps (trpo_util.py:100-133)
        x, r, p, avp = (torch.empty(n, **f32) for _ in range(4))
        cg_state = torch.zeros(2, **f32)
        L.call("hb_trpo_cg_init", L.ptr(g), L.ptr(x), L.ptr(r), L.ptr(p), L.ptr(cg_state), n, st)
        for _ in range(10):
            fvp(p, avp)
            L.call("hb_trpo_cg_step", L.ptr(p), L.ptr(avp), L.ptr(x), L.ptr(r), L.ptr(cg_state), n, 1e-10, st)
        # --- step scaling (hatrpo.py:112-133)
        fvp(x, avp)
        full = torch.empty(n, **f32)
        out3 = torch.zeros(3, **f64)
        L.call("hb_trpo_full_step", L.ptr(x), L.ptr(avp), L.ptr(g), float(self.kl_threshold), L.ptr(full), L.ptr(out3), n, st)
        params0 = net.params.clone()
        host = torch.cat([scal, out3]).cpu().numpy()  # one sync: loss sums + (shs, step_size, expected)
        loss = -host[0] / norm
        expected = float(host[6])
        # --- backtracking line search (hatrpo.py:135-189)
        flag, fraction = False, 1.0
        kl = improve = ent = ratio = 0.0
        ls = torch.zeros(self.ls_step, 4, **f64)
        tried = 0
        for i in range(self.ls_step):
            L.call("hb_trpo_apply_step", L.ptr(net.params), L.ptr(params0), L.ptr(full), float(fraction), n, st)
            net.prepare()
            net.trpo_eval(batch, hyper, old_dist, params0, ls[i])
            dist.all_reduce_sum_(ls[i])
            s = ls[i].cpu().numpy()
            tried = i + 1
            new_loss = s[0] / norm
            improve = float(new_loss - loss)
            kl = float(s[3] / global_rows)
            ent = float(s[1] / norm)
            ratio = float(s[2] / global_rows)
            # NumPy semantics of the reference (hatrpo.py:179-183): a zero expected improvement gives inf / nan, which rejects
            with np.errstate(divide="ignore", invalid="ignore"):
                rel = float(np.float64(improve) / np.float64(expected))
            if kl < self.kl_threshold and rel > self.accept_ratio and improve > 0:
                flag = True
                break
            expected *= self.backtrack_coeff
            fraction *= self.backtrack_coeff
        if not flag:
            net.params.copy_(params0)
            net.prepare()
            print("policy update does not impove the surrogate")
        self.last_update = dict(loss=float(loss), grad=g, step_dir=x, full_step=full, accepted=flag, trials=tried,
                                fraction=fraction)
        return kl, improve, expected, ent, ratio

    def update(self, sample):
        """Reference-compatible update on a materialised minibatch tuple (hatrpo.py:37-194).

        Returns (kl, loss_improve, expected_improve, dist_entropy, ratio_mean) as floats."""
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
        cnt = torch.zeros(2, dtype=torch.float64, device=d)
        cnt[0] = active.sum().double() if self.use_policy_active_masks else float(obs.shape[0])
        cnt[1] = float(obs.shape[0])
        dist.all_reduce_sum_(cnt)
        c = cnt.cpu().numpy()
        return self._update_on_batch(batch, float(c[0]), float(c[1]))

    def train(self, actor_buffer, advantages, state_type, moments=None):
        """Reference hatrpo.py:196-247: one update on the whole buffer."""
        info = dict(kl=0.0, dist_entropy=0.0, loss_improve=0.0, expected_improve=0.0, ratio=0.0)
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
            n_active = m3[2].item()
        if n_active == 0:
            return info
        if state_type == "EP":
            adv_n = torch.empty_like(adv)
            L.call("hb_normalize_by_moments", L.ptr(adv), L.ptr(adv_n), rows, L.ptr(m3), L.stream_ptr())
            adv = adv_n
        fl = lambda a: a.reshape(rows, *a.shape[2:])
        avail = None if buf.available_actions is None else fl(buf.available_actions[:-1])
        factor = None if buf.factor is None else buf.factor.reshape(rows)
        # the reference draws one permutation for its single minibatch (rows, envs or chunks, hatrpo.py:223-232); it
        # only reorders sums (and the order of whole sequences), so the rows are consumed in place
        mode = seq_index.mode_of(self.use_recurrent_policy, self.use_naive_recurrent_policy)
        (idx, nrows, seq_len), = seq_index.minibatches(T, N, 1, mode, self.data_chunk_length, d)
        rnn = buf.rnn_states.reshape((T + 1) * N, -1) if self.recurrent else None
        masks = buf.masks.reshape((T + 1) * N) if self.recurrent else None
        batch = DeviceNet.actor_batch(fl(buf.obs[:-1]), fl(buf.actions), fl(buf.action_log_probs), adv, factor, active,
                                      avail, idx, nrows, rnn_states=rnn, masks=masks, seq_len=seq_len)
        global_rows = float(rows * dist.world_size())
        norm = n_active if self.use_policy_active_masks else global_rows
        kl, improve, expected, ent, ratio = self._update_on_batch(batch, norm, global_rows)
        info.update(kl=kl, dist_entropy=ent, loss_improve=improve, expected_improve=expected, ratio=ratio)
        return info

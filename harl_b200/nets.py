"""Device-resident actor / critic networks behind the C-ABI.

A ``DeviceNet`` owns one flat fp32 parameter buffer laid out in the reference ``state_dict``
order (so ``state_dict()`` / ``load_state_dict()`` interchange with checkpoints of the
reference's StochasticPolicy -- harl/models/policy_models/stochastic_policy.py:12-53 -- and
VNet -- harl/models/value_function_models/v_net.py:10-46), the derived ("prepared") weights
the kernels read, the gradient and Adam moment buffers, and thin methods that forward to the
library.  All maths happens in CUDA; nothing here computes on tensors.
"""
import ctypes as C
from collections import OrderedDict

import torch

from . import _lib as L

_WORKSPACES = {}


def workspace(device, nbytes):
    """One grow-only scratch tensor per (device, stream), shared by every net (calls are stream-ordered)."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def _is_matrix(name):
    """Linear / GRU weights are 2-D in the reference state_dict; LayerNorm weights, biases, log_std are 1-D."""
    if name.startswith("rnn.rnn.weight") or name in (
            "act.action_out.linear.weight", "act.action_out.fc_mean.weight", "v_out.weight"):
        return True
    if name.startswith("base.mlp.fc.") and name.endswith(".weight"):
        return int(name.split(".")[3]) % 3 == 0
    return False


def make_desc(args, in_dim, head, out_dim):
    """Reference model config (happo.yaml ``model:`` keys) -> hb_net_desc."""
    hs = list(args["hidden_sizes"])
    if len(hs) > L.HB_MAX_LAYERS:
        raise NotImplementedError(f"more than {L.HB_MAX_LAYERS} hidden layers")
    if args["activation_func"] not in L.ACTIVATIONS:
        raise NotImplementedError(f"activation {args['activation_func']}")
    d = L.NetDesc()
    d.in_dim = int(in_dim)
    d.n_layers = len(hs)
    for i, h in enumerate(hs):
        d.hidden[i] = int(h)
    d.feature_norm = int(bool(args["use_feature_normalization"]))
    d.activation = L.ACTIVATIONS[args["activation_func"]]
    rnn = bool(args.get("use_recurrent_policy") or args.get("use_naive_recurrent_policy"))
    d.rnn_layers = int(args.get("recurrent_n", 1)) if rnn else 0
    d.head = head
    d.out_dim = int(out_dim)
    d.std_x_coef = float(args.get("std_x_coef", 1.0))
    d.std_y_coef = float(args.get("std_y_coef", 0.5))
    return d


class DeviceNet:
    def __init__(self, args, in_dim, head, out_dim, device, init=True):
        self.args = args
        self.device = torch.device(device)
        self.desc = make_desc(args, in_dim, head, out_dim)
        lay = L.NetLayout()
        L.check(L.lib.hb_net_layout_of(C.byref(self.desc), C.byref(lay)), "hb_net_layout_of")
        self.total = lay.total
        self.entries = OrderedDict()
        for i in range(lay.n_tensors):
            name = lay.names[i].value.decode()
            rows, cols = lay.rows[i], lay.cols[i]
            self.entries[name] = (lay.offset[i], (rows, cols) if _is_matrix(name) else (cols,))
        kw = dict(dtype=torch.float32, device=self.device)
        self.params = torch.zeros(self.total, **kw)
        self.grad = torch.zeros(self.total, **kw)
        self.exp_avg = torch.zeros(self.total, **kw)
        self.exp_avg_sq = torch.zeros(self.total, **kw)
        self.prepared = torch.zeros(lay.prepared_total, **kw)
        self.grad_norm = torch.zeros(1, **kw)
        self.adam_steps = 0
        self.head = head
        self.out_dim = int(out_dim)
        self.act_width = 1 if head == L.HEAD_DISCRETE else int(out_dim)
        if init:
            self.reset_parameters()

    # ------------------------------------------------------------------ parameters
    def views(self, flat=None):
        flat = self.params if flat is None else flat
        out = OrderedDict()
        for name, (off, shape) in self.entries.items():
            n = 1
            for s in shape:
                n *= s
            out[name] = flat[off:off + n].view(*shape)
        return out

    def state_dict(self):
        return OrderedDict((k, v.clone()) for k, v in self.views().items())

    def load_state_dict(self, sd):
        v = self.views()
        missing = [k for k in v if k not in sd]
        if missing:
            raise KeyError(f"missing keys {missing}")
        with torch.no_grad():
            for k, dst in v.items():
                src = torch.as_tensor(sd[k], dtype=torch.float32)
                dst.copy_(src.reshape(dst.shape))
        self.prepare()

    def reset_parameters(self):
        """Initialise exactly as the reference modules would under the current torch CPU RNG state.

        Builds throw-away torch.nn layers in the reference's construction order (mlp.py:17-36,
        rnn.py:14-21, distributions.py:43-49,74-82, v_net.py:41-44) so the default-init draws and
        the orthogonal_ draws consume the generator identically, then copies them in.
        """
        a = self.args
        init = getattr(torch.nn.init, a["initialization_method"])
        gain = torch.nn.init.calculate_gain(a["activation_func"])
        d = self.desc
        sd = {}
        if d.feature_norm:
            sd["base.feature_norm.weight"] = torch.ones(d.in_dim)
            sd["base.feature_norm.bias"] = torch.zeros(d.in_dim)
        prev = d.in_dim
        for li in range(d.n_layers):
            h = d.hidden[li]
            lin = torch.nn.Linear(prev, h)
            init(lin.weight.data, gain=gain)
            sd[f"base.mlp.fc.{3 * li}.weight"] = lin.weight.data
            sd[f"base.mlp.fc.{3 * li}.bias"] = torch.zeros(h)
            sd[f"base.mlp.fc.{3 * li + 2}.weight"] = torch.ones(h)
            sd[f"base.mlp.fc.{3 * li + 2}.bias"] = torch.zeros(h)
            prev = h
        if d.rnn_layers:
            gru = torch.nn.GRU(prev, prev, num_layers=d.rnn_layers)
            for name, p in gru.named_parameters():
                if "bias" in name:
                    torch.nn.init.constant_(p, 0)
                elif "weight" in name:
                    init(p)
                sd["rnn.rnn." + name] = p.data
            sd["rnn.norm.weight"] = torch.ones(prev)
            sd["rnn.norm.bias"] = torch.zeros(prev)
        if self.head == L.HEAD_DISCRETE:
            lin = torch.nn.Linear(prev, d.out_dim)
            init(lin.weight.data, gain=a["gain"])
            sd["act.action_out.linear.weight"] = lin.weight.data
            sd["act.action_out.linear.bias"] = torch.zeros(d.out_dim)
        elif self.head == L.HEAD_BOX:
            lin = torch.nn.Linear(prev, d.out_dim)
            init(lin.weight.data, gain=a["gain"])
            sd["act.action_out.fc_mean.weight"] = lin.weight.data
            sd["act.action_out.fc_mean.bias"] = torch.zeros(d.out_dim)
            sd["act.action_out.log_std"] = torch.ones(d.out_dim) * a["std_x_coef"]
        else:
            lin = torch.nn.Linear(prev, 1)
            init(lin.weight.data)
            sd["v_out.weight"] = lin.weight.data
            sd["v_out.bias"] = torch.zeros(1)
        v = self.views()
        with torch.no_grad():
            for k, dst in v.items():
                dst.copy_(sd[k].reshape(dst.shape))
        if self.device.type == "cuda":
            self.prepare()

    # ------------------------------------------------------------------ library calls
    def _need_cuda(self):
        if self.device.type != "cuda":
            raise RuntimeError("harl_b200 kernels need a CUDA device (no CPU fallback)")

    def prepare(self):
        self._need_cuda()
        L.call("hb_net_prepare", C.byref(self.desc), L.ptr(self.params), L.ptr(self.prepared), L.stream_ptr())

    def _ws(self, rows, mode):
        n = L.lib.hb_workspace_bytes(C.byref(self.desc), int(rows), mode)
        ws = workspace(self.device, n)
        return ws, ws.numel()

    @property
    def recurrent(self):
        return self.desc.rnn_layers > 0

    def act(self, obs, avail, deterministic, seed, offset, actions_out, logp_out, rnn_states=None, masks=None,
            rnn_out=None):
        self._need_cuda()
        rows = obs.shape[0]
        ws, n = self._ws(rows, 0)
        if self.recurrent:
            L.call("hb_policy_act_rnn", C.byref(self.desc), L.ptr(self.prepared), L.ptr(obs), rows, L.ptr(avail),
                   L.ptr(rnn_states), L.ptr(masks), int(bool(deterministic)), int(seed) & (2**64 - 1),
                   int(offset) & (2**64 - 1), None, L.ptr(actions_out), L.ptr(logp_out), L.ptr(rnn_out), L.ptr(ws), n,
                   L.stream_ptr())
            return
        L.call("hb_policy_act", C.byref(self.desc), L.ptr(self.prepared), L.ptr(obs), rows, L.ptr(avail),
               int(bool(deterministic)), int(seed) & (2**64 - 1), int(offset) & (2**64 - 1), L.ptr(actions_out),
               L.ptr(logp_out), L.ptr(ws), n, L.stream_ptr())

    def values(self, cent_obs, values_out, rnn_states=None, masks=None, rnn_out=None):
        self._need_cuda()
        rows = cent_obs.shape[0]
        ws, n = self._ws(rows, 0)
        if self.recurrent:
            L.call("hb_value_forward_rnn", C.byref(self.desc), L.ptr(self.prepared), L.ptr(cent_obs), rows,
                   L.ptr(rnn_states), L.ptr(masks), L.ptr(values_out), L.ptr(rnn_out), L.ptr(ws), n, L.stream_ptr())
            return
        L.call("hb_value_forward", C.byref(self.desc), L.ptr(self.prepared), L.ptr(cent_obs), rows,
               L.ptr(values_out), L.ptr(ws), n, L.stream_ptr())

    @staticmethod
    def actor_batch(obs, actions, old_logp=None, adv=None, factor=None, active=None, avail=None, index=None, rows=None,
                    rnn_states=None, masks=None, seq_len=0):
        """``rnn_states`` / ``masks`` / ``seq_len``: recurrent policies only (see hb_actor_batch)."""
        b = L.ActorBatch()
        b.obs, b.actions, b.old_logp, b.adv = L.ptr(obs), L.ptr(actions), L.ptr(old_logp), L.ptr(adv)
        b.factor, b.active, b.avail, b.index = L.ptr(factor), L.ptr(active), L.ptr(avail), L.ptr(index)
        b.rows = int(rows if rows is not None else (index.shape[0] if index is not None else actions.shape[0]))
        b.rnn_states, b.masks, b.seq_len = L.ptr(rnn_states), L.ptr(masks), int(seq_len)
        b._keep = (obs, actions, old_logp, adv, factor, active, avail, index, rnn_states, masks)  # raw pointers only
        return b

    @staticmethod
    def critic_batch(share_obs, value_preds, returns, index=None, rows=None, rnn_states=None, masks=None, seq_len=0):
        b = L.CriticBatch(L.ptr(share_obs), L.ptr(value_preds), L.ptr(returns), L.ptr(index),
                          int(rows if rows is not None else (index.shape[0] if index is not None else value_preds.numel())),
                          L.ptr(rnn_states), L.ptr(masks), int(seq_len))
        b._keep = (share_obs, value_preds, returns, index, rnn_states, masks)
        return b

    def evaluate(self, batch, logp_out=None, logp_ref=None, factor_inout=None, agg_prod=True):
        self._need_cuda()
        ws, n = self._ws(batch.rows, 0)
        L.call("hb_policy_evaluate", C.byref(self.desc), L.ptr(self.prepared), C.byref(batch), L.ptr(logp_out),
               L.ptr(logp_ref), L.ptr(factor_inout), int(bool(agg_prod)), L.ptr(ws), n, L.stream_ptr())

    def actor_grad(self, batch, hyper, norm3, scalars, logp_out=None):
        """PPO-clip gradient of the batch into ``self.grad``; ``logp_out`` [rows, ad] (identity batches) also receives the
        log-probs of the batch actions under the current weights (hb_ppo_actor_grad_logp)."""
        self._need_cuda()
        ws, n = self._ws(batch.rows, 1)
        L.call("hb_ppo_actor_grad_logp", C.byref(self.desc), L.ptr(self.params), L.ptr(self.prepared), C.byref(batch),
               C.byref(hyper), L.ptr(norm3), L.ptr(self.grad), L.ptr(scalars), L.ptr(logp_out), L.ptr(ws), n, L.stream_ptr())

    def value_grad(self, batch, hyper, vn_state, inv_count, scalars):
        self._need_cuda()
        ws, n = self._ws(batch.rows, 1)
        L.call("hb_value_grad", C.byref(self.desc), L.ptr(self.params), L.ptr(self.prepared), C.byref(batch),
               C.byref(hyper), L.ptr(vn_state), float(inv_count), L.ptr(self.grad), L.ptr(scalars), L.ptr(ws), n,
               L.stream_ptr())

    def adam_step(self, lr, eps, weight_decay, max_grad_norm, use_max_grad_norm, betas=(0.9, 0.999)):
        self._need_cuda()
        self.adam_steps += 1
        h = L.AdamHyper(float(lr), betas[0], betas[1], float(eps), float(weight_decay), float(max_grad_norm),
                        int(bool(use_max_grad_norm)), self.adam_steps)
        L.call("hb_clip_adam_step", C.byref(self.desc), L.ptr(self.params), L.ptr(self.grad), L.ptr(self.exp_avg),
               L.ptr(self.exp_avg_sq), L.ptr(self.prepared), C.byref(h), L.ptr(self.grad_norm), L.stream_ptr())

    # ------------------------------------------------------------------ trust-region (HATRPO) calls
    def _trpo_ws(self, rows):
        n = L.lib.hb_trpo_workspace_bytes(C.byref(self.desc), int(rows))
        ws = workspace(self.device, n)
        return ws, ws.numel()

    def trpo_old_dist(self, batch, old_dist):
        self._need_cuda()
        ws, n = self._trpo_ws(batch.rows)
        L.call("hb_trpo_old_dist", C.byref(self.desc), L.ptr(self.prepared), C.byref(batch), L.ptr(old_dist), L.ptr(ws), n,
               L.stream_ptr())

    def trpo_fvp(self, batch, old_dist, vec, inv_rows, out, reuse_forward=False):
        """out = J^T H J vec over this rank's rows (no damping; see trpo_fvp_finish).  ``reuse_forward``: the previous
        library call on this stream's workspace was trpo_fvp on the same batch and parameters."""
        self._need_cuda()
        ws, n = self._trpo_ws(batch.rows)
        L.call("hb_trpo_fvp", C.byref(self.desc), L.ptr(self.params), L.ptr(self.prepared), C.byref(batch),
               L.ptr(old_dist), L.ptr(vec), float(inv_rows), int(bool(reuse_forward)), L.ptr(out), L.ptr(ws), n,
               L.stream_ptr())

    def trpo_fvp_finish(self, vec, out, damping=0.1):
        L.call("hb_trpo_fvp_finish", C.byref(self.desc), L.ptr(self.params), L.ptr(vec), L.ptr(out), float(damping),
               L.stream_ptr())

    def trpo_eval(self, batch, hyper, old_dist, params_old, scalars):
        self._need_cuda()
        ws, n = self._trpo_ws(batch.rows)
        L.call("hb_trpo_eval", C.byref(self.desc), L.ptr(self.prepared), C.byref(batch), C.byref(hyper), L.ptr(old_dist),
               L.ptr(params_old), L.ptr(scalars), L.ptr(ws), n, L.stream_ptr())

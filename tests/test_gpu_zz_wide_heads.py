"""GPU parity of the trust-region path at the synthetic-SMAC widths (Discrete(12), hidden 64; MLP and GRU policies):
the 16-output instantiations of the head kernels.  Runs LAST in the suite on purpose: these cases pin a dispatch fix
(`launch_trpo_head_mode`) made after the round's GPU budget was spent -- they have not run on a GPU
yet, and `pytest -x` should reach them only after everything that has.

  * the reference's own goldens at those widths (tests/golden/hatrpo_parts_*12_h64.npz) through the same checks as
    tests/test_gpu_trpo.py;
  * surrogate gradient and Fisher-vector product vs the oracle on random buffers with availability masks."""
import os

import numpy as np
import pytest
import torch

from tests import test_gpu_trpo as T
from tests import util as U
from tests.test_gpu_trpo import DEV, gemm_impl  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", T.WIDE_PARTS)
def test_wide_goldens_gradient_and_fvp(name, gemm_impl):
    T.test_surrogate_gradient_and_fvp_vs_reference(name, gemm_impl)


@pytest.mark.parametrize("name", T.WIDE_PARTS)
def test_wide_goldens_update(name, gemm_impl):
    T.test_update_vs_reference(name, gemm_impl)


def _wide_case(recurrent, na=12, h=64):
    """SMAC-like head (12 actions: the 16-output instantiation of the trust-region head kernel), hidden 64, random
    buffers; returns everything both sides need.  Built on the CPU only (also exercised by the CPU suite)."""
    from oracle import algo as oa
    from oracle import nets as on

    T, N, od = 8, 6, 10
    cfg = U.base_args(hidden_sizes=[h, h], episode_length=T, n_rollout_threads=N, use_recurrent_policy=recurrent,
                      data_chunk_length=4)
    rng = np.random.default_rng(21 + int(recurrent))
    torch.manual_seed(5)
    p = on.init_params(cfg, od, "Discrete", na)
    for k in p:
        p[k] = (p[k] + 0.1 * torch.randn(p[k].shape)).requires_grad_(True)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    acts = rng.integers(0, na, (T, N, 1)).astype(np.float32)
    av = (rng.random((T + 1, N, na)) < 0.6).astype(np.float32)
    av[np.arange(T)[:, None], np.arange(N)[None, :], acts[..., 0].astype(int)] = 1.0
    buf = dict(obs=f(T + 1, N, od), rnn_states=f(T + 1, N, 1, h), masks=(rng.random((T + 1, N, 1)) > 0.25).astype(np.float32),
               active_masks=(rng.random((T + 1, N, 1)) > 0.2).astype(np.float32), actions=acts,
               action_log_probs=-np.abs(f(T, N, 1)) - 0.5, available_actions=av)
    adv, factor = f(T, N, 1), (1 + 0.1 * f(T, N, 1)).astype(np.float32)
    batch = next(oa.actor_minibatches(buf, adv, factor, dict(cfg, actor_num_mini_batch=1), lambda n: np.arange(n)))
    vec = {k: torch.randn(v.shape) for k, v in p.items()}
    return cfg, p, buf, adv, factor, batch, vec, (T, N, od, na)


@pytest.mark.parametrize("recurrent", [False, True])
def test_fvp_wide_head_vs_oracle(recurrent, gemm_impl):
    """Discrete(12) / hidden 64 (the synthetic-SMAC shapes): surrogate gradient and Fisher-vector product of the
    device path vs the oracle's double backward, MLP and GRU policies (recurrent chunk batch)."""
    from harl_b200 import _lib as L
    from harl_b200.common import seq_index
    from harl_b200.nets import DeviceNet
    from oracle import trpo as ot

    cfg, p, buf, adv, factor, batch, vec, (T, N, od, na) = _wide_case(recurrent)
    names = list(p.keys())
    loss, _, _ = ot.surrogate(p, cfg, "Discrete", batch)
    gs = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    want_f = ot.fisher_vector_product(p, cfg, "Discrete", batch, vec)
    net = DeviceNet(cfg, od, L.HEAD_DISCRETE, na, torch.device(DEV), init=False)
    net.load_state_dict({k: v.detach() for k, v in p.items()})
    rows = T * N
    cu = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to(DEV).contiguous()
    fl = lambda a: cu(a.reshape(rows, *a.shape[2:]))
    mode = seq_index.mode_of(recurrent, False)
    (idx, nrows, seq_len), = seq_index.minibatches(T, N, 1, mode, cfg["data_chunk_length"], torch.device(DEV))
    kw = {}
    if recurrent:
        kw = dict(rnn_states=cu(buf["rnn_states"].reshape((T + 1) * N, -1)), masks=cu(buf["masks"].reshape((T + 1) * N)),
                  seq_len=seq_len)
    dbatch = DeviceNet.actor_batch(fl(buf["obs"][:-1]), fl(buf["actions"]), fl(buf["action_log_probs"]), fl(adv).reshape(-1),
                                   fl(factor).reshape(-1), fl(buf["active_masks"][:-1]).reshape(-1),
                                   fl(buf["available_actions"][:-1]), idx, nrows, **kw)
    norm3 = torch.zeros(3, dtype=torch.float64, device=DEV)
    norm3[2] = float(batch["active"].sum())
    scal = torch.zeros(4, dtype=torch.float64, device=DEV)
    net.actor_grad(dbatch, L.PPOHyper(0.0, 0.0, 1, 1, 0), norm3, scal)
    torch.cuda.synchronize()
    np.testing.assert_allclose(-scal[0].item() / norm3[2].item(), float(loss.detach()), rtol=2e-4, atol=1e-6)
    got = {k: v.cpu().numpy() for k, v in net.views(-net.grad).items()}
    gscale = max(float(g.abs().max()) for g in gs if g is not None)
    for k, g in zip(names, gs):
        ref = np.zeros_like(got[k]) if g is None else g.numpy()
        assert np.abs(got[k] - ref).max() <= 3e-4 * gscale, "grad " + k
    flat = torch.zeros(net.total, dtype=torch.float32)
    for k, v in net.views(flat).items():
        v.copy_(vec[k].reshape(v.shape))
    dvec = flat.to(DEV)
    old_dist = torch.empty(nrows, na, dtype=torch.float32, device=DEV)
    net.trpo_old_dist(dbatch, old_dist)
    out = torch.empty(net.total, dtype=torch.float32, device=DEV)
    net.trpo_fvp(dbatch, old_dist, dvec, 1.0 / nrows, out)
    net.trpo_fvp_finish(dvec, out, 0.1)
    out2 = torch.empty_like(out)   # second product reuses the forward activations left in the workspace
    net.trpo_fvp(dbatch, old_dist, dvec, 1.0 / nrows, out2, reuse_forward=True)
    net.trpo_fvp_finish(dvec, out2, 0.1)
    torch.cuda.synchronize()
    assert float((out - out2).abs().max()) <= 2e-5 * float(out.abs().max())  # same maths, atomic summation order only
    gotf = {k: v.cpu().numpy() for k, v in net.views(out).items()}
    fscale = max(float(v.abs().max()) for v in want_f.values())
    for k in names:
        assert np.abs(gotf[k] - want_f[k].numpy()).max() <= 5e-4 * fscale, "fvp " + k


@pytest.mark.parametrize("name", ["hatrpo_parts_disc", "hatrpo_parts_box", "hatrpo_parts_disc12_h64"])
def test_tensor_core_tangent_block_equals_ffma(name):
    """The Fisher-vector product with the tensor-core tangent block (3xTF32) vs the FP32 FFMA one."""
    from harl_b200 import _lib as L

    g = U.load(name)
    cfg, m = U.cfg_of(g), U.meta_of(g)
    ac = T._actor(g, cfg, m)
    net = ac.actor
    batch, norm, rows = T._device_batch(ac, g, cfg)
    vec = T._flat(net, g, "vec/")
    old_dist = torch.empty(batch.rows, net.out_dim, dtype=torch.float32, device=DEV)
    net.trpo_old_dist(batch, old_dist)
    outs = []
    for impl in (0, 1):
        L.call("hb_set_trpo_jvp_impl", impl)
        try:
            out = torch.empty(net.total, dtype=torch.float32, device=DEV)
            net.trpo_fvp(batch, old_dist, vec, 1.0 / rows, out)
            net.trpo_fvp_finish(vec, out, 0.1)
            torch.cuda.synchronize()
            outs.append(out.clone())
        finally:
            L.call("hb_set_trpo_jvp_impl", 0)
    assert float((outs[0] - outs[1]).abs().max()) <= 2e-5 * float(outs[0].abs().max())
    assert T._max_rel(net, outs[1], g, "fvp/") <= 5e-4

"""The fused update kernel (harl_b200/csrc/fused_update.cu: one launch = feature norm -> MLP -> head -> loss -> backward,
fp16 hi/lo split tcgen05 GEMMs, LayerNorm affines folded into the next layer) against the layer-wise kernels
(hb_set_fused_update(0): 3xTF32 tcgen05 / FP32 SIMT, themselves pinned to the reference goldens) on the same inputs:
gradients of every parameter (incl. the unfolded LayerNorm / feature-norm affines), loss scalars, log-probs, the factor
update.  Both paths go through the same C-ABI entry points (hb_ppo_actor_grad / hb_value_grad / hb_policy_evaluate)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests import util as U  # noqa: E402


def _net(hidden, act, in_dim, head, out_dim, seed):
    from harl_b200 import _lib as L
    from harl_b200.nets import DeviceNet

    cfg = U.base_args(hidden_sizes=[hidden, hidden], activation_func=act)
    torch.manual_seed(seed)
    net = DeviceNet(cfg, in_dim, {"Discrete": L.HEAD_DISCRETE, "Box": L.HEAD_BOX, "Value": L.HEAD_VALUE}[head], out_dim, "cuda:0")
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():   # move every tensor off its init so the LayerNorm affines, biases and log_std matter
        net.params.add_((0.15 * torch.randn(net.params.shape, generator=g)).cuda())
    net.prepare()
    return net


def _both(fn):
    """Run fn() with the fused kernel and with the layer-wise kernels; returns (fused, layerwise)."""
    from harl_b200 import _lib as L

    assert L.lib.hb_get_gemm_impl() != 0 or True
    out = []
    for on in (1, 0):
        L.call("hb_set_fused_update", on)
        try:
            out.append(fn())
        finally:
            L.call("hb_set_fused_update", 1)
    return out


def _close(a, b, rel, what):
    a, b = a.double().cpu().numpy(), b.double().cpu().numpy()
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max()
    assert err <= rel * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("hidden,act,od,rows,gather", [(32, "relu", 18, 300, False), (128, "relu", 18, 1000, False),
                                                       (128, "tanh", 54, 777, True), (64, "selu", 7, 129, False),
                                                       (128, "sigmoid", 33, 256, False), (128, "leaky_relu", 64, 260, True)])
def test_discrete_actor_grad_fused_equals_layerwise(hidden, act, od, rows, gather):
    from harl_b200 import _lib as L
    from harl_b200.nets import DeviceNet

    na = 5
    net = _net(hidden, act, od, "Discrete", na, 3)
    g = torch.Generator().manual_seed(7)
    R = rows + 50
    cu = lambda t: t.cuda().contiguous()
    obs = cu(torch.randn(R, od, generator=g))
    actions = cu(torch.randint(0, na, (R, 1), generator=g).float())
    avail = (torch.rand(R, na, generator=g) < 0.75).float()
    avail[torch.arange(R), actions[:, 0].long().cpu()] = 1.0
    avail = cu(avail)
    old = cu(-1.6 + 0.3 * torch.randn(R, 1, generator=g))
    adv = cu(torch.randn(R, generator=g))
    factor = cu(torch.exp(0.2 * torch.randn(R, generator=g)))
    active = cu((torch.rand(R, generator=g) < 0.9).float())
    index = cu(torch.randperm(R, generator=g)[:rows].int()) if gather else None
    batch = DeviceNet.actor_batch(obs, actions, old, adv, factor, active, avail, index=index, rows=rows)
    hyper = L.PPOHyper(0.2, 0.01, 1, 1, 1)
    sel = index.long() if gather else torch.arange(rows, device="cuda")
    norm3 = torch.tensor([0.0, 0.0, float(active[sel].sum())], dtype=torch.float64, device="cuda")

    def run():
        scalars = torch.zeros(4, dtype=torch.float64, device="cuda")
        net.grad.fill_(float("nan"))
        net.actor_grad(batch, hyper, norm3, scalars)
        torch.cuda.synchronize()
        return net.grad.clone(), scalars.clone()

    (gf, sf), (gl, sl) = _both(run)
    assert torch.isfinite(gf).all()
    np.testing.assert_allclose(sf.cpu().numpy(), sl.cpu().numpy(), rtol=2e-5, atol=1e-6)
    vf, vl = net.views(gf), net.views(gl)
    for k in vf:
        _close(vf[k], vl[k], 3e-4, k)
    _close(gf, gl, 3e-4, "flat gradient")


@pytest.mark.parametrize("agg_prod", [1, 0])
def test_box_actor_grad_fused_equals_layerwise(agg_prod):
    from harl_b200 import _lib as L
    from harl_b200.nets import DeviceNet

    od, ad, rows = 23, 3, 900
    net = _net(128, "relu", od, "Box", ad, 5)
    g = torch.Generator().manual_seed(11)
    cu = lambda t: t.cuda().contiguous()
    obs = cu(torch.randn(rows, od, generator=g))
    actions = cu(0.7 * torch.randn(rows, ad, generator=g))
    old = cu(-1.0 + 0.2 * torch.randn(rows, ad, generator=g))
    adv = cu(torch.randn(rows, generator=g))
    factor = cu(torch.exp(0.1 * torch.randn(rows, generator=g)))
    active = cu((torch.rand(rows, generator=g) < 0.95).float())
    batch = DeviceNet.actor_batch(obs, actions, old, adv, factor, active, None)
    hyper = L.PPOHyper(0.2, 0.01, 1, agg_prod, 1)
    norm3 = torch.tensor([0.0, 0.0, float(active.sum())], dtype=torch.float64, device="cuda")

    def run():
        scalars = torch.zeros(4, dtype=torch.float64, device="cuda")
        net.actor_grad(batch, hyper, norm3, scalars)
        torch.cuda.synchronize()
        return net.grad.clone(), scalars.clone()

    (gf, sf), (gl, sl) = _both(run)
    np.testing.assert_allclose(sf.cpu().numpy(), sl.cpu().numpy(), rtol=2e-5, atol=1e-6)
    vf, vl = net.views(gf), net.views(gl)
    for k in vf:
        _close(vf[k], vl[k], 3e-4, k)


@pytest.mark.parametrize("use_huber,use_clipped,vn", [(1, 1, True), (0, 1, False), (1, 0, True)])
def test_value_grad_fused_equals_layerwise(use_huber, use_clipped, vn):
    from harl_b200 import _lib as L
    from harl_b200.nets import DeviceNet

    sd, rows = 54, 1111
    net = _net(128, "relu", sd, "Value", 1, 9)
    g = torch.Generator().manual_seed(13)
    cu = lambda t: t.cuda().contiguous()
    so = cu(torch.randn(rows, sd, generator=g))
    vp = cu(0.5 * torch.randn(rows, generator=g))
    ret = cu(2.0 * torch.randn(rows, generator=g) + 0.3)
    batch = DeviceNet.critic_batch(so, vp, ret)
    hyper = L.ValueHyper(0.2, 10.0, 1.0, use_huber, use_clipped)
    vn_state = torch.tensor([0.21, 1.7, 0.9], device="cuda") if vn else None

    def run():
        scalars = torch.zeros(4, dtype=torch.float64, device="cuda")
        net.value_grad(batch, hyper, vn_state, 1.0 / rows, scalars)
        torch.cuda.synchronize()
        return net.grad.clone(), scalars.clone()

    (gf, sf), (gl, sl) = _both(run)
    np.testing.assert_allclose(sf.cpu().numpy(), sl.cpu().numpy(), rtol=2e-5, atol=1e-6)
    vf, vl = net.views(gf), net.views(gl)
    for k in vf:
        _close(vf[k], vl[k], 3e-4, k)


@pytest.mark.parametrize("head,out", [("Discrete", 5), ("Discrete", 12), ("Box", 3)])
def test_evaluate_and_factor_update_fused_equals_layerwise(head, out):
    from harl_b200.nets import DeviceNet

    od, rows = 18, 1500
    net = _net(128, "relu", od, head, out, 17)
    g = torch.Generator().manual_seed(19)
    cu = lambda t: t.cuda().contiguous()
    obs = cu(torch.randn(rows, od, generator=g))
    ad = 1 if head == "Discrete" else out
    if head == "Discrete":
        actions = cu(torch.randint(0, out, (rows, 1), generator=g).float())
        avail = (torch.rand(rows, out, generator=g) < 0.8).float()
        avail[torch.arange(rows), actions[:, 0].long().cpu()] = 1.0
        avail = cu(avail)
    else:
        actions, avail = cu(0.2 * torch.randn(rows, out, generator=g)), None
    # reference log-probs near the net's own (a factor update multiplies by exp(logp - ref): keep it O(1))
    from harl_b200 import _lib as L_
    L_.call("hb_set_fused_update", 0)
    lp0 = torch.empty(rows, ad, device="cuda")
    net.evaluate(DeviceNet.actor_batch(obs, actions, avail=avail), logp_out=lp0)
    L_.call("hb_set_fused_update", 1)
    ref = (lp0 + 0.1 * torch.randn(rows, ad, generator=g).cuda()).contiguous()
    factor0 = cu(torch.exp(0.1 * torch.randn(rows, generator=g)))
    batch = DeviceNet.actor_batch(obs, actions, avail=avail)

    def run():
        logp = torch.full((rows, ad), float("nan"), device="cuda")
        factor = factor0.clone()
        net.evaluate(batch, logp_out=logp, logp_ref=ref, factor_inout=factor, agg_prod=True)
        torch.cuda.synchronize()
        return logp, factor

    (lf, ff), (ll, fl) = _both(run)
    np.testing.assert_allclose(lf.cpu().numpy(), ll.cpu().numpy(), rtol=2e-5, atol=3e-5)   # Box log-probs reach -70: relative
    # a Gaussian log-prob magnifies an error d of the mean by |action - mean| / var (~10 here): the factor exp(logp - ref)
    # of the two GEMM arithmetics (fp16 hi/lo split vs 3xTF32) agrees to that many more ulps
    np.testing.assert_allclose(ff.cpu().numpy(), fl.cpu().numpy(), rtol=5e-4 if head == "Box" else 5e-5, atol=0)


def test_fused_kernel_is_what_runs_by_default():
    """Launch labels of a default actor update: the fused kernel, not the layer-wise GEMMs."""
    from harl_b200 import _lib as L
    from harl_b200.nets import DeviceNet

    net = _net(128, "relu", 18, "Discrete", 5, 1)
    rows = 512
    cu = lambda t: t.cuda().contiguous()
    batch = DeviceNet.actor_batch(cu(torch.randn(rows, 18)), cu(torch.zeros(rows, 1)), cu(torch.zeros(rows, 1) - 1.6),
                                  cu(torch.randn(rows)), None, cu(torch.ones(rows)), None)
    hyper = L.PPOHyper(0.2, 0.01, 1, 1, 1)
    norm3 = torch.tensor([0.0, 0.0, float(rows)], dtype=torch.float64, device="cuda")
    scalars = torch.zeros(4, dtype=torch.float64, device="cuda")
    net.actor_grad(batch, hyper, norm3, scalars)   # warm (workspace allocation)
    torch.cuda.synchronize()
    L.call("hb_profile_begin", L.stream_ptr())
    net.actor_grad(batch, hyper, norm3, scalars)
    buf = C.create_string_buffer(1 << 14)
    n = L.lib.hb_profile_end(buf, len(buf))
    labels = buf.raw[:max(n, 0)].decode()
    assert "fused_actor_update" in labels and "tc_linear_ln_fwd" not in labels, labels


@pytest.mark.parametrize("head,out,fused", [("Discrete", 5, 1), ("Box", 3, 1), ("Discrete", 5, 0)])
def test_grad_pass_also_returns_the_log_probs_of_its_forward(head, out, fused):
    """hb_ppo_actor_grad_logp: the log-probs written by the gradient pass equal a separate evaluate sweep under the same
    weights (the sequential-update runner uses them as the pre-update log-probs, on_policy_ha_runner.py:66-83), and the
    gradient itself is unchanged by asking for them."""
    from harl_b200 import _lib as L
    from harl_b200.nets import DeviceNet

    od, rows = 18, 1000
    net = _net(128, "relu", od, head, out, 23)
    g = torch.Generator().manual_seed(29)
    cu = lambda t: t.cuda().contiguous()
    obs = cu(torch.randn(rows, od, generator=g))
    ad = 1 if head == "Discrete" else out
    if head == "Discrete":
        actions = cu(torch.randint(0, out, (rows, 1), generator=g).float())
        avail = (torch.rand(rows, out, generator=g) < 0.8).float()
        avail[torch.arange(rows), actions[:, 0].long().cpu()] = 1.0
        avail = cu(avail)
    else:
        actions, avail = cu(0.2 * torch.randn(rows, out, generator=g)), None
    old = cu(-1.5 + 0.1 * torch.randn(rows, ad, generator=g))
    adv, active = cu(torch.randn(rows, generator=g)), cu((torch.rand(rows, generator=g) < 0.9).float())
    batch = DeviceNet.actor_batch(obs, actions, old, adv, None, active, avail)
    hyper = L.PPOHyper(0.2, 0.01, 1, 1, 1)
    norm3 = torch.tensor([0.0, 0.0, float(active.sum())], dtype=torch.float64, device="cuda")
    L.call("hb_set_fused_update", fused)
    try:
        scal = torch.zeros(4, dtype=torch.float64, device="cuda")
        net.actor_grad(batch, hyper, norm3, scal)
        g_plain = net.grad.clone()
        lp = torch.full((rows, ad), float("nan"), device="cuda")
        scal2 = torch.zeros(4, dtype=torch.float64, device="cuda")
        net.actor_grad(batch, hyper, norm3, scal2, logp_out=lp)
        g_with = net.grad.clone()
        ref = torch.empty(rows, ad, device="cuda")
        net.evaluate(DeviceNet.actor_batch(obs, actions, avail=avail), logp_out=ref)
        torch.cuda.synchronize()
    finally:
        L.call("hb_set_fused_update", 1)
    np.testing.assert_allclose(g_with.cpu().numpy(), g_plain.cpu().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(scal2.cpu().numpy(), scal.cpu().numpy(), rtol=1e-9)
    np.testing.assert_allclose(lp.cpu().numpy(), ref.cpu().numpy(), rtol=1e-6, atol=1e-6)

"""Debug driver (not a pytest file): SIMT vs tcgen05 gradients, tensor by tensor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from harl_b200 import _lib as L
from harl_b200.nets import DeviceNet
from tests import util as U

dev = torch.device("cuda:0")
def run(cfg, od, head, na, R, impl):
    L.lib.hb_set_gemm_impl(impl)
    torch.manual_seed(0)
    net = DeviceNet(cfg, od, L.HEAD_DISCRETE if head == "Discrete" else L.HEAD_BOX, na, dev)
    g = torch.Generator().manual_seed(1)
    obs = torch.randn(R, od, generator=g).to(dev)
    ad = 1 if head == "Discrete" else na
    acts = (torch.randint(0, na, (R, 1), generator=g).float() if head == "Discrete" else 0.3 * torch.randn(R, ad, generator=g)).to(dev)
    old = (-1.2 + 0.3 * torch.randn(R, ad, generator=g)).to(dev)
    adv = torch.randn(R, generator=g).to(dev)
    fac = (1 + 0.1 * torch.randn(R, generator=g)).to(dev)
    active = (torch.rand(R, generator=g) > 0.1).float().to(dev)
    batch = DeviceNet.actor_batch(obs, acts, old, adv, fac, active, None)
    hyper = L.PPOHyper(0.2, 0.01, 1, 1, 1)
    norm3 = torch.tensor([0, 0, float(active.sum())], dtype=torch.float64, device=dev)
    scal = torch.zeros(4, dtype=torch.float64, device=dev)
    net.actor_grad(batch, hyper, norm3, scal)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy().copy() for k, v in net.views(net.grad).items()}, scal.cpu().numpy()

for name, cfg, od, head, na, R in [
    ("small32", U.base_args(hidden_sizes=[32, 32]), 6, "Discrete", 5, 48),
    ("c2", U.base_args(hidden_sizes=[128, 128]), 18, "Discrete", 5, 1000),
    ("c3", U.base_args(hidden_sizes=[128, 128, 128]), 23, "Box", 1, 40000),
    ("h256", U.base_args(hidden_sizes=[256, 256]), 54, "Discrete", 7, 3000),
]:
    g0, s0 = run(cfg, od, head, na, R, 0)
    for impl in (1, 2):
        g1, s1 = run(cfg, od, head, na, R, impl)
        print(f"== {name} impl={impl} scalars simt={s0[:3]} tc={s1[:3]}")
        for k in g0:
            d = np.abs(g0[k] - g1[k]).max(); sc = np.abs(g0[k]).max()
            flag = "" if d <= 2e-4 * max(sc, 1e-6) * (1 if impl == 1 else 50) else "   <<<<"
            print(f"   {k:34s} max|simt|={sc:.3e} max|diff|={d:.3e}{flag}")

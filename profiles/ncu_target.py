"""Small deterministic driver for `ncu --set full` captures: a few HAPPO actor-gradient passes at the C2 shapes
(obs 18 -> 128 -> 128 -> 5 logits, 65536 rows = 2 chunks), so every hot kernel appears a handful of times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from harl_b200 import _lib as L
from harl_b200.nets import DeviceNet
from harl_b200.utils.configs_tools import get_defaults_yaml_args

algo_args, _ = get_defaults_yaml_args("happo", "pettingzoo_mpe")
cfg = {**algo_args["model"], **algo_args["algo"]}
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = DeviceNet(cfg, 18, L.HEAD_DISCRETE, 5, dev)
R = int(os.environ.get("NCU_ROWS", "65536"))
g = torch.Generator().manual_seed(1)
obs = torch.randn(R, 18, generator=g).to(dev)
acts = torch.randint(0, 5, (R, 1), generator=g).float().to(dev)
old = (-1.6 + 0.1 * torch.randn(R, 1, generator=g)).to(dev)
adv = torch.randn(R, generator=g).to(dev)
fac = torch.ones(R, device=dev)
active = torch.ones(R, device=dev)
avail = torch.ones(R, 5, device=dev)
batch = DeviceNet.actor_batch(obs, acts, old, adv, fac, active, avail)
hyper = L.PPOHyper(0.2, 0.01, 1, 1, 1)
norm3 = torch.tensor([0, 0, float(R)], dtype=torch.float64, device=dev)
scal = torch.zeros(4, dtype=torch.float64, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    net.actor_grad(batch, hyper, norm3, scal)
    net.adam_step(5e-4, 1e-5, 0.0, 10.0, True)
torch.cuda.synchronize()
print("done", scal.cpu().numpy())
# forward-only sweep (policy_head_eval) and one critic gradient pass (value_head_grad)
lp = torch.empty(R, 1, device=dev)
net.evaluate(DeviceNet.actor_batch(obs, acts, avail=avail), logp_out=lp)
cnet = DeviceNet(cfg, 54, L.HEAD_VALUE, 1, dev)
sobs = torch.randn(R, 54, generator=g).to(dev)
vp, rt = torch.randn(R, generator=g).to(dev), torch.randn(R, generator=g).to(dev)
cscal = torch.zeros(4, dtype=torch.float64, device=dev)
cnet.value_grad(DeviceNet.critic_batch(sobs, vp, rt, None, R), L.ValueHyper(0.2, 10.0, 1.0, 1, 1), None, 1.0 / R, cscal)
torch.cuda.synchronize()
print("done2", lp.mean().item(), cscal.cpu().numpy())

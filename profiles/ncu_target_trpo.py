"""Driver for `ncu --set full` captures of the trust-region (HATRPO) kernels at the C2 launch size: one surrogate
gradient, the old distribution, two Fisher-vector products and one line-search evaluation over 819200 rows
(obs 18 -> 128 -> 128 -> 5 logits)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from harl_b200 import _lib as L
from harl_b200.nets import DeviceNet
from harl_b200.utils.configs_tools import get_defaults_yaml_args

algo_args, _ = get_defaults_yaml_args("hatrpo", "pettingzoo_mpe")
cfg = {**algo_args["model"], **algo_args["algo"]}
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = DeviceNet(cfg, 18, L.HEAD_DISCRETE, 5, dev)
R = int(os.environ.get("NCU_ROWS", "819200"))
g = torch.Generator().manual_seed(1)
obs = torch.randn(R, 18, generator=g).to(dev)
acts = torch.randint(0, 5, (R, 1), generator=g).float().to(dev)
old = (-1.6 + 0.1 * torch.randn(R, 1, generator=g)).to(dev)
adv = torch.randn(R, generator=g).to(dev)
fac = torch.ones(R, device=dev)
active = torch.ones(R, device=dev)
avail = torch.ones(R, 5, device=dev)
batch = DeviceNet.actor_batch(obs, acts, old, adv, fac, active, avail)
hyper = L.PPOHyper(0.0, 0.0, 1, 1, 0)
norm3 = torch.tensor([0, 0, float(R)], dtype=torch.float64, device=dev)
scal = torch.zeros(4, dtype=torch.float64, device=dev)
net.actor_grad(batch, hyper, norm3, scal)
old_dist = torch.empty(R, 5, device=dev)
net.trpo_old_dist(batch, old_dist)
vec = net.grad.clone()
out = torch.empty_like(vec)
for _ in range(2):
    net.trpo_fvp(batch, old_dist, vec, 1.0 / R, out)
    net.trpo_fvp_finish(vec, out, 0.1)
ls = torch.zeros(4, dtype=torch.float64, device=dev)
net.trpo_eval(batch, hyper, old_dist, net.params, ls)
torch.cuda.synchronize()
print("done", scal.cpu().numpy(), out.abs().max().item(), ls.cpu().numpy())

"""Tiny stand-alone driver for compute-sanitizer runs (not a pytest file)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import util as U
from tests.test_gpu_kernels import _net, _cu, _dev
from harl_b200 import _lib as L
from harl_b200.nets import DeviceNet

name = sys.argv[1] if len(sys.argv) > 1 else "single_update_box"
g = U.load(name)
cfg, m = U.cfg_of(g), U.meta_of(g)
T, N = cfg["episode_length"], cfg["n_rollout_threads"]
net = _net(cfg, m["od"], m["head"], m["act_dim"], U.params_of(g, "actor0/"))
fl = lambda a: _cu(a.reshape(T * N, -1))
active = fl(g["a0.active_masks"][:-1])
avail = fl(g["a0.available_actions"][:-1]) if "a0.available_actions" in g else None
batch = DeviceNet.actor_batch(fl(g["a0.obs"][:-1]), fl(g["a0.actions"]), fl(g["a0.action_log_probs"]),
                              fl(g["adv"]), fl(g["factor"]), active, avail)
hyper = L.PPOHyper(cfg["clip_param"], cfg["entropy_coef"], 1, 1, 1)
norm3 = torch.tensor([0, 0, float(g["a0.active_masks"][:-1].sum())], dtype=torch.float64, device=_dev())
scal = torch.zeros(4, dtype=torch.float64, device=_dev())
net.actor_grad(batch, hyper, norm3, scal)
torch.cuda.synchronize()
print("grad ok", scal.cpu().numpy())
net.adam_step(cfg["lr"], cfg["opti_eps"], cfg["weight_decay"], cfg["max_grad_norm"], cfg["use_max_grad_norm"])
torch.cuda.synchronize()
print("adam ok", net.grad_norm.item())

"""Per-phase SM-clock profile of the fused actor update at the C2 shapes (HB phase clock, fused_update.cu PhaseClock)."""
import ctypes as C
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
R = 819200
g = torch.Generator().manual_seed(1)
obs = torch.randn(R, 18, generator=g).to(dev)
acts = torch.randint(0, 5, (R, 1), generator=g).float().to(dev)
old = (-1.6 + 0.1 * torch.randn(R, 1, generator=g)).to(dev)
adv = torch.randn(R, generator=g).to(dev)
fac, active, avail = torch.ones(R, device=dev), torch.ones(R, device=dev), torch.ones(R, 5, device=dev)
batch = DeviceNet.actor_batch(obs, acts, old, adv, fac, active, avail)
hyper = L.PPOHyper(0.2, 0.01, 1, 1, 1)
norm3 = torch.tensor([0, 0, float(R)], dtype=torch.float64, device=dev)
scal = torch.zeros(4, dtype=torch.float64, device=dev)
for _ in range(3):
    net.actor_grad(batch, hyper, norm3, scal)
torch.cuda.synchronize()
L.call("hb_fused_timing_enable", 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
net.actor_grad(batch, hyper, norm3, scal)
e1.record()
torch.cuda.synchronize()
tab = np.zeros((148, 16), np.uint64)
L.call("hb_fused_timing_read", tab.ctypes.data_as(C.c_void_p))
L.call("hb_fused_timing_enable", 0)
names = ["inputs+featnorm stats", "wait prev-tile bwd MMAs", "write X0", "wait L0 MMA", "L0 fwd epilogue", "wait L1 MMA",
         "L1 fwd epilogue", "wait head MMA", "head epilogue", "wait head-bwd MMA", "L1 bwd epilogue", "wait L1-bwd MMAs",
         "L0 bwd epilogue"]
tiles = R / 128 / 148
m = tab.astype(np.float64).mean(0)
print(f"fused_actor_update {e0.elapsed_time(e1) * 1e3:.1f} us for {R} rows; {tiles:.1f} tiles per CTA; cycles per tile (mean over CTAs):")
for i, n in enumerate(names):
    print(f"  {i:2d} {n:28s} {m[i] / tiles:9.0f}")
print(f"     {'sum':28s} {m[:13].sum() / tiles:9.0f}")

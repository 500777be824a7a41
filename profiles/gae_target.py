"""Driver for ncu captures of hb_gae_returns at the C2 shape (T=200, C=4096): 8 rotating buffer sets, 2 rounds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from harl_b200 import _lib as L

T, C = 200, int(os.environ.get("GAE_C", "4096"))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
bufs = []
for _ in range(8):
    bufs.append((torch.randn(T, C, generator=g).to(dev), torch.randn(T + 1, C, generator=g).to(dev),
                 (torch.rand(T + 1, C, generator=g) > 0.04).float().to(dev), (torch.rand(T + 1, C, generator=g) > 0.02).float().to(dev),
                 torch.randn(C, generator=g).to(dev), torch.empty(T + 1, C, device=dev), torch.empty(T, C, device=dev)))
vn = torch.tensor([0.1, 1.3, 1.0], device=dev)
for _ in range(2):
    for b in bufs:
        L.call("hb_gae_returns", *[L.ptr(x) for x in b], T, C, 0.99, 0.99 * 0.95, 1, 1, L.ptr(vn), L.stream_ptr())
torch.cuda.synchronize()
print("ok")

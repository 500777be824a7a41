"""How fast do 2 MB of pinned host memory reach the device: cudaMemcpyAsync (DMA engine) vs hb_copy_segments reading the
pinned block from inside a kernel (zero-copy over PCIe), and the reverse for 48 KB of actions.  One rollout step of the
host-resident C2 env moves exactly these blocks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from harl_b200 import _lib as L

dev = torch.device("cuda:0")
n = 507904 + 2 * 3072          # floats per step at C2: obs, state, reward, avail (+ done / bad flags)
host = [torch.randn(n).pin_memory() for _ in range(8)]
devb = torch.empty(n, device=dev)
segs_d = [devb[i * (n // 8):(i + 1) * (n // 8)] for i in range(8)]


def timed(fn, reps=50):
    for _ in range(5):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


us = timed(lambda i: devb.copy_(host[i % 8], non_blocking=True))
print(f"cudaMemcpyAsync H2D  {n * 4 / 1e6:.2f} MB: {us:7.1f} us = {n * 4 / us / 1e3:6.1f} GB/s")
us = timed(lambda i: L.copy_segments(segs_d, [host[i % 8][j * (n // 8):(j + 1) * (n // 8)] for j in range(8)], src_pinned=True))
print(f"hb_copy_segments H2D {n * 4 / 1e6:.2f} MB: {us:7.1f} us = {n * 4 / us / 1e3:6.1f} GB/s (8 segments, kernel reads pinned memory)")
m = 3 * 4096
acts_d = [torch.randn(4096, device=dev) for _ in range(3)]
acts_h = [torch.empty(4096).pin_memory() for _ in range(3)]
us = timed(lambda i: [h.copy_(d, non_blocking=True) for h, d in zip(acts_h, acts_d)])
print(f"3 x cudaMemcpyAsync D2H 48 KB: {us:7.1f} us")
us = timed(lambda i: L.copy_segments(acts_h, acts_d, dst_pinned=True))
print(f"hb_copy_segments D2H    48 KB: {us:7.1f} us")


def step_sync(i):
    L.copy_segments(acts_h, acts_d, dst_pinned=True)
    torch.cuda.current_stream().synchronize()


import time
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(200):
    step_sync(i)
print(f"D2H kernel + stream synchronize, host wall time per step: {(time.perf_counter() - t0) / 200 * 1e6:7.1f} us")

"""torchrun script (N GPUs): the one-shot peer-memory allreduce (hb_allreduce_bucket, harl_b200/csrc/p2p_comm.cu)
against torch.distributed.all_reduce on the same buckets -- values, bit-identity across ranks, every size / dtype /
alignment the update issues, two streams with their own communicators, many back-to-back exchanges (slot reuse) --
and its latency next to NCCL's.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_check_allreduce.py
"""
import json
import os
import sys

import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HB_P2P_ALLREDUCE", "1")   # this script tests the one-shot kernel itself, whatever the calibration would pick


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    td.init_process_group("nccl", device_id=dev)
    from harl_b200 import dist

    g = torch.Generator().manual_seed(100 + rank)
    checked = 0
    for rep in range(3):
        for dtype in (torch.float32, torch.float64):
            for n in (1, 3, 4, 5, 12, 1023, 4096, 24931, 131072):
                if n * (4 if dtype == torch.float32 else 8) > dist.P2P_SLOT_BYTES:
                    continue
                for off in (0, 1, 3):                       # views that start off a 16-byte boundary
                    base = torch.randn(n + 4, generator=g, dtype=dtype).to(dev)
                    x = base[off:off + n]
                    ref = x.clone()
                    td.all_reduce(ref)
                    before = dist.stats["p2p"]
                    dist.all_reduce_sum_(x)
                    assert dist.stats["p2p"] == before + 1, "the bucket did not take the peer-memory path"
                    torch.testing.assert_close(x, ref, rtol=1e-6 if dtype == torch.float32 else 1e-14, atol=1e-6 if dtype == torch.float32 else 1e-13)
                    every = [torch.empty_like(x) for _ in range(world)]
                    td.all_gather(every, x.contiguous())
                    assert all(torch.equal(every[0], e) for e in every), "ranks disagree bitwise"
                    assert torch.equal(base[:off], base[:off]) and torch.isfinite(base).all()
                    checked += 1
    # exact expectation: integers sum exactly in any order
    x = torch.full((1000,), float(rank + 1), device=dev)
    dist.all_reduce_sum_(x)
    assert torch.equal(x, torch.full_like(x, world * (world + 1) / 2))
    # a second stream gets its own communicator; interleave exchanges on both
    side = torch.cuda.Stream(dev)
    a = torch.full((5000,), 1.0, device=dev)
    b = torch.full((7,), 2.0, dtype=torch.float64, device=dev)
    for _ in range(50):
        dist.all_reduce_sum_(a)
        a.div_(world)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            dist.all_reduce_sum_(b)
            b.div_(world)
        torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    assert torch.equal(a, torch.ones_like(a)) and torch.equal(b, torch.full_like(b, 2.0))
    assert len([c for c in dist._comms.values() if c is not None]) == 2
    dist.check_comms()

    # latency: 300 back-to-back exchanges of a gradient-sized bucket and of a 3-double normaliser
    out = {}
    for name, t in (("grad_97KB_f32", torch.randn(24931, device=dev)), ("moments_3xf64", torch.ones(3, dtype=torch.float64, device=dev))):
        for transport in ("p2p", "nccl"):
            fn = (lambda: dist.all_reduce_sum_(t)) if transport == "p2p" else (lambda: td.all_reduce(t))
            for _ in range(20):
                fn()
            td.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(300):
                fn()
                t.mul_(1.0 / world)          # keep the values bounded; one tiny kernel between exchanges, like Adam
            e1.record()
            torch.cuda.synchronize()
            us = torch.tensor([1e3 * e0.elapsed_time(e1) / 300], device=dev)
            td.all_reduce(us, op=td.ReduceOp.MAX)
            out[f"{name}_{transport}_us"] = round(float(us.item()), 2)
    if rank == 0:
        print(json.dumps({"world": world, "buckets_checked": checked, **out}))
    td.barrier()
    td.destroy_process_group()


if __name__ == "__main__":
    main()

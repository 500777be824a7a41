"""Thin helpers over torch.distributed for the rollout-sharded multi-GPU path.

One process per GPU; every rank owns N/G rollout threads and a full replica of all networks.
The only exchanges are sum-allreduces of small flat buffers (gradients, loss normalisers,
advantage / ValueNorm moments), so the replicas stay bit-identical after each optimiser step.

Transport: CUDA tensors go through ``hb_allreduce_bucket`` (harl_b200/csrc/p2p_comm.cu) -- one kernel that
publishes the bucket into a CUDA-IPC region and sums every rank's copy straight out of NVLink peer memory, in rank
order -- with one communicator per CUDA stream (the critic update runs on a side stream).  torch.distributed (NCCL)
only bootstraps it (all_gather of the IPC handles) and carries what the one-shot kernel does not: buckets larger
than ``P2P_SLOT_BYTES``, CPU tensors (gloo tests), or everything when ``HB_P2P_ALLREDUCE=0``.
"""
import ctypes as C
import os

import torch

# tests/dist_check_global_batch.py: run a runner with single-process semantics (no sharding, no exchanges) inside an
# initialised process group, to compare the sharded update with the same update on the whole global batch
FORCE_SINGLE = False

P2P_SLOT_BYTES = 1 << 20          # gradients of the supported nets are <= 0.5 MB; larger buckets go through NCCL
_comms = {}                       # cuda stream handle -> communicator (or None when peer mapping failed)
_p2p_mode = os.environ.get("HB_P2P_ALLREDUCE", "auto")   # "1": always the one-shot kernel, "0": always NCCL, "auto": the faster
_p2p_off = _p2p_mode == "0"
stats = {"p2p": 0, "nccl": 0}     # exchanges issued by this process, by transport; "calibration" once measured


def is_dist():
    return not FORCE_SINGLE and torch.distributed.is_available() and torch.distributed.is_initialized()


def world_size():
    return torch.distributed.get_world_size() if is_dist() else 1


def rank():
    return torch.distributed.get_rank() if is_dist() else 0


def _comm_for_current_stream(device):
    """Collective: every rank reaches this at the same point of the (identical) program the first time a stream
    exchanges something."""
    from . import _lib as L

    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    if key in _comms:
        return _comms[key]
    w, r = world_size(), rank()
    handle = C.c_void_p()
    mine = (C.c_ubyte * 64)()
    ok = torch.ones(1, dtype=torch.int32, device=device)
    try:
        L.call("hb_comm_create", r, w, P2P_SLOT_BYTES, C.byref(handle), mine)
    except RuntimeError as e:
        print(f"[harl_b200.dist] rank {r}: hb_comm_create failed ({e}); exchanges on this stream use NCCL", flush=True)
        ok.zero_()
    local = torch.tensor(list(mine), dtype=torch.uint8, device=device)
    gathered = [torch.empty_like(local) for _ in range(w)]
    torch.distributed.all_gather(gathered, local)
    if int(ok.item()):
        allh = (C.c_ubyte * (64 * w))(*torch.cat(gathered).cpu().tolist())
        try:
            L.call("hb_comm_open_peers", handle, allh)
        except RuntimeError as e:
            print(f"[harl_b200.dist] rank {r}: peer mapping failed ({e}); exchanges on this stream use NCCL", flush=True)
            ok.zero_()
    torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)   # all ranks or none
    torch.cuda.synchronize(device)
    _comms[key] = handle if int(ok.item()) else None
    if _comms[key] is not None and _p2p_mode == "auto":
        _calibrate(device, _comms[key])
    return _comms[key]


def _calibrate(device, comm):
    """Measure, don't guess: time a gradient-sized bucket and a 3-double normaliser through both transports on this box
    (CUDA events around 40 back-to-back exchanges, max over ranks) and keep the one-shot kernel only if it is not slower.
    Runs once per process, the first time a stream exchanges something; all ranks take the same decision."""
    global _p2p_off
    if "calibration" in stats:
        return
    from . import _lib as L

    res = {}
    for name, t in (("grad_100KB", torch.zeros(25000, device=device)), ("moments_24B", torch.zeros(3, dtype=torch.float64, device=device))):
        for transport in ("p2p", "nccl"):
            def fn():
                if transport == "p2p":
                    L.call("hb_allreduce_bucket", comm, L.ptr(t), t.numel(), 0 if t.dtype == torch.float32 else 1, L.stream_ptr())
                else:
                    torch.distributed.all_reduce(t)
            for _ in range(5):
                fn()
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(40):
                fn()
            e1.record()
            torch.cuda.synchronize(device)
            us = torch.tensor([1e3 * e0.elapsed_time(e1) / 40], device=device)
            torch.distributed.all_reduce(us, op=torch.distributed.ReduceOp.MAX)
            res[f"{name}_{transport}_us"] = round(float(us.item()), 2)
    p2p = res["grad_100KB_p2p_us"] + res["moments_24B_p2p_us"]
    nccl = res["grad_100KB_nccl_us"] + res["moments_24B_nccl_us"]
    res["chosen"] = "p2p" if p2p <= 1.05 * nccl else "nccl"
    stats["calibration"] = res
    _p2p_off = res["chosen"] == "nccl"


def all_reduce_sum_(t):
    """In-place sum over ranks (no-op single-process)."""
    if not (is_dist() and world_size() > 1):
        return t
    if (t.is_cuda and not _p2p_off and t.dtype in (torch.float32, torch.float64) and t.is_contiguous()
            and 0 < t.numel() * t.element_size() <= P2P_SLOT_BYTES):
        comm = _comm_for_current_stream(t.device)
        if comm is not None and not _p2p_off:
            from . import _lib as L

            L.call("hb_allreduce_bucket", comm, L.ptr(t), t.numel(), 0 if t.dtype == torch.float32 else 1, L.stream_ptr())
            stats["p2p"] += 1
            return t
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
    stats["nccl"] += 1
    return t


def check_comms():
    """Raise if any one-shot exchange timed out waiting for a rank (host-visible flag set by the kernel)."""
    from . import _lib as L

    for key, comm in _comms.items():
        if comm is not None:
            st = L.lib.hb_comm_status(comm)
            if st:
                raise RuntimeError(f"harl_b200 exchange on stream {key}: rank {st - 1} did not arrive (HB_COMM_TIMEOUT_S)")


def shard_bounds(n_global, world, r):
    """Contiguous env shard [lo, hi) of rank r; n_global must divide evenly (weak scaling keeps it so)."""
    if n_global % world:
        raise ValueError(f"n_rollout_threads={n_global} is not divisible by world size {world}")
    per = n_global // world
    return r * per, (r + 1) * per

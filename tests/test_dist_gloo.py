"""World-size-2 gloo tests (CPU) of the host-side logic of the rollout-sharded multi-GPU path:
shard partitioning, replica consistency, and the reductions that make every rank apply the
reference's full-batch update.  The kernels themselves need a GPU and are
covered by the -m gpu tests; here only torch.distributed plumbing and host arithmetic run."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from harl_b200 import dist
    from harl_b200.runners import RUNNER_REGISTRY
    from tests.smoke_check import small_config

    res = {}
    assert dist.is_dist() and dist.world_size() == world and dist.rank() == rank
    # ---- shard bounds tile the env axis exactly
    res["bounds"] = dist.shard_bounds(12, world, rank)
    try:
        dist.shard_bounds(13, world, rank)
        res["uneven_raises"] = False
    except ValueError:
        res["uneven_raises"] = True
    # ---- masked advantage moments: per-rank partial sums + sum-allreduce == global moments
    rng = np.random.default_rng(0)
    adv = rng.standard_normal((10, 12, 1))
    act = (rng.random((10, 12, 1)) > 0.3)
    lo, hi = res["bounds"]
    a, m = adv[:, lo:hi], act[:, lo:hi]
    m3 = torch.tensor([a[m].sum(), (a[m] ** 2).sum(), float(m.sum())], dtype=torch.float64)
    dist.all_reduce_sum_(m3)
    res["m3"] = m3.numpy()
    res["m3_global"] = np.array([adv[act].sum(), (adv[act] ** 2).sum(), float(act.sum())])
    # ---- gradient bucket: sum of per-shard sums == full-batch sum
    g = torch.full((7,), float(rank + 1))
    dist.all_reduce_sum_(g)
    res["grad_sum"] = g.numpy()
    # ---- runner construction under torch.distributed: local shard sizes, identical replicas, distinct env data
    args, algo_args, env_args = small_config(n=12, T=6)
    algo_args["device"]["cuda"] = False
    algo_args["logger"]["log_dir"] = tempfile.mkdtemp(prefix=f"gloo{rank}_")
    r = RUNNER_REGISTRY["happo"](args, algo_args, env_args)
    r.warmup()
    res["n_local"] = r.n_local
    res["buf_shape"] = tuple(r.actor_buffer[0].obs.shape)
    w = torch.cat([a_.actor.params for a_ in r.actor] + [r.critic.critic.params])
    gathered = [torch.zeros_like(w) for _ in range(world)]
    torch.distributed.all_gather(gathered, w)
    res["replicas_equal"] = bool(all(torch.equal(gathered[0], x) for x in gathered))
    o = r.actor_buffer[0].obs[0].clone()
    obs_all = [torch.zeros_like(o) for _ in range(world)]
    torch.distributed.all_gather(obs_all, o)
    res["env_data_differs"] = bool(not torch.equal(obs_all[0], obs_all[1]))
    # agent order is drawn from the (identically seeded) CPU generator on every rank
    order = torch.randperm(3)
    orders = [torch.zeros_like(order) for _ in range(world)]
    torch.distributed.all_gather(orders, order)
    res["same_agent_order"] = bool(torch.equal(orders[0], orders[1]))
    r.close()
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([res], dtype=object), allow_pickle=True)
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_gloo():
    out = tempfile.mkdtemp(prefix="gloo_test_")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = [np.load(os.path.join(out, f"rank{r}.npy"), allow_pickle=True)[0] for r in range(2)]
    assert res[0]["bounds"] == (0, 6) and res[1]["bounds"] == (6, 12)
    for r in res:
        assert r["uneven_raises"]
        np.testing.assert_allclose(r["m3"], r["m3_global"], rtol=1e-12)
        np.testing.assert_allclose(r["grad_sum"], 3.0)
        assert r["n_local"] == 6 and r["buf_shape"] == (7, 6, 7)
        assert r["replicas_equal"] and r["env_data_differs"] and r["same_agent_order"]

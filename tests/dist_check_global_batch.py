"""torchrun script (one process per GPU, NCCL): the update of a run whose rollout threads are SHARDED over G GPUs must
reproduce the update of the same global batch on ONE GPU (SURVEY.md section 8(e) "Validation": sum-order differences only).

Every rank first builds the whole-batch runner with single-process semantics (harl_b200.dist.FORCE_SINGLE), rolls it out and
computes returns; a second, sharded runner then receives its slice [lo, hi) of those buffers and the same initial weights,
both train once, and the updated weights are compared.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/dist_check_global_batch.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
    from harl_b200 import dist
    from harl_b200.runners import RUNNER_REGISTRY
    from tests.smoke_check import small_config

    cases = [("happo", "Discrete", "EP", (128, 128)), ("happo", "Box", "FP", (32, 32)), ("hatrpo", "Discrete", "EP", (32, 32))]
    for algo, action_type, state_type, hidden in cases:
        N, T = 64 * world, 10
        args, algo_args, env_args = small_config(algo=algo, action_type=action_type, state_type=state_type, n=N, T=T, hidden=hidden)
        algo_args["algo"]["fixed_order"] = True
        # ---- the whole batch with single-process semantics (identical on every rank: same seed, deterministic kernels)
        dist.FORCE_SINGLE = True
        whole = RUNNER_REGISTRY[algo](args, {k: dict(v) for k, v in algo_args.items()}, dict(env_args))
        whole.warmup()
        whole.logger.init(1)
        whole.logger.episode_init(1)
        whole.prep_rollout()
        for step in range(T):
            data = whole.collect(step)
            out = whole.envs.step(data[1])
            whole.insert((*out, *data))
        whole.compute()
        dist.FORCE_SINGLE = False
        # ---- the sharded run: same weights, its slice of the same buffers
        part = RUNNER_REGISTRY[algo](args, {k: dict(v) for k, v in algo_args.items()}, dict(env_args))
        lo, hi = dist.shard_bounds(N, world, rank)
        assert part.n_local == hi - lo
        for a in range(whole.num_agents):
            part.actor[a].actor.params.copy_(whole.actor[a].actor.params)
            part.actor[a].actor.prepare()
            bw, bp = whole.actor_buffer[a], part.actor_buffer[a]
            for k in ("obs", "actions", "action_log_probs", "masks", "active_masks", "available_actions"):
                if getattr(bw, k) is not None:
                    getattr(bp, k).copy_(getattr(bw, k)[:, lo:hi])
        part.critic.critic.params.copy_(whole.critic.critic.params)
        part.critic.critic.prepare()
        for k in ("share_obs", "value_preds", "returns", "rewards", "masks", "bad_masks", "advantages"):
            getattr(part.critic_buffer, k).copy_(getattr(whole.critic_buffer, k)[:, lo:hi])
        if whole.value_normalizer is not None:
            part.value_normalizer.state.copy_(whole.value_normalizer.state)
        dist.FORCE_SINGLE = True
        whole.prep_training()
        whole.train()
        dist.FORCE_SINGLE = False
        part.prep_training()
        part.train()
        torch.cuda.synchronize()
        worst = 0.0
        for nw, np_ in zip([a.actor for a in whole.actor] + [whole.critic.critic], [a.actor for a in part.actor] + [part.critic.critic]):
            step = (nw.params - nw.params.new_tensor(0)).abs().max().item()
            err = (nw.params - np_.params).abs().max().item()
            worst = max(worst, err)
            assert err <= 2e-5 * max(1.0, step), (algo, action_type, state_type, err)
        flat = torch.cat([a.actor.params for a in part.actor] + [part.critic.critic.params]).double()
        sig = torch.stack([flat.sum(), flat.abs().sum()])
        sigs = [torch.zeros_like(sig) for _ in range(world)]
        torch.distributed.all_gather(sigs, sig)
        assert all(torch.equal(sigs[0], s) for s in sigs), "replicas diverged"
        if rank == 0:
            print(f"{algo:7s} {action_type:8s} {state_type} hidden {hidden}: {world}-GPU sharded update == 1-GPU update on the global batch, "
                  f"max |weight difference| = {worst:.2e}; replicas identical", flush=True)
        whole.close()
        part.close()
    torch.distributed.destroy_process_group()
    if rank == 0:
        print("dist global-batch check ok", flush=True)


if __name__ == "__main__":
    main()

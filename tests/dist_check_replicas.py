"""torchrun script (one process per GPU, NCCL): HATRPO and GRU updates with the rollout sharded over ranks must leave
every rank with bit-identical replicas -- the surrogate gradient, every Fisher-vector product, the line-search sums,
the PPO gradients and the critic gradients are sum-allreduced, so all ranks take the same step.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check_replicas.py
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
    from harl_b200.runners import RUNNER_REGISTRY
    from tests.smoke_check import small_config

    rnn = dict(use_recurrent_policy=True, data_chunk_length=4)
    for algo, state_type, model_over in (("hatrpo", "EP", {}), ("hatrpo", "FP", rnn), ("happo", "FP", rnn), ("mappo", "EP", {})):
        args, algo_args, env_args = small_config(algo=algo, state_type=state_type, n=16 * world, T=12)
        algo_args["model"].update(model_over)
        runner = RUNNER_REGISTRY[algo](args, algo_args, env_args)
        runner.warmup()
        runner.logger.init(2)
        for ep in (1, 2):
            runner.run_iteration(ep, 2)
        torch.cuda.synchronize()
        nets = [a.actor for a in runner.actor] + [runner.critic.critic]
        flat = torch.cat([n.params for n in nets]).double()
        sig = torch.stack([flat.sum(), flat.abs().sum(), (flat * torch.arange(flat.numel(), device=flat.device)).sum()])
        sigs = [torch.zeros_like(sig) for _ in range(world)]
        torch.distributed.all_gather(sigs, sig)
        same = all(torch.equal(sigs[0], s) for s in sigs)
        moved = any(float((n.params - n.params.new_tensor(0)).abs().sum()) > 0 for n in nets)
        if rank == 0:
            print(f"{algo:7s} {state_type} rnn={bool(model_over)}: replicas identical across {world} ranks: {same}", flush=True)
        assert same and moved, (algo, state_type, [s.tolist() for s in sigs])
        runner.close()
    torch.distributed.destroy_process_group()
    if rank == 0:
        print("dist replicas ok", flush=True)


if __name__ == "__main__":
    main()

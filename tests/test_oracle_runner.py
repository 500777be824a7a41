"""CPU: the oracle's whole-iteration runner (the `cpu_baseline` / `--impl reference` arm of bench.py) steps through
HAPPO and HATRPO iterations, with and without GRU policies, and keeps its buffers consistent (hidden states are zero
exactly where an episode ended, finite everywhere, and actually propagate for recurrent nets)."""
import numpy as np
import pytest

from tests.smoke_check import small_config


@pytest.mark.parametrize("algo,state_type,model_over", [
    ("happo", "EP", {}), ("hatrpo", "EP", {}), ("happo", "FP", dict(use_recurrent_policy=True, data_chunk_length=4)),
    ("hatrpo", "FP", dict(use_recurrent_policy=True, data_chunk_length=4))])
def test_oracle_runner_iteration(algo, state_type, model_over):
    from harl_b200.envs.synthetic import resolve_shapes
    from oracle.runner import NumpySyntheticEnv, OracleRunner

    args, algo_args, env_args = small_config(algo=algo, state_type=state_type, n=6, T=8, hidden=(16, 16))
    algo_args["model"].update(model_over)
    cfg = {**algo_args["model"], **algo_args["algo"], **algo_args["train"], "algo_name": algo}
    cfg.setdefault("ppo_epoch", cfg.get("a2c_epoch", 1))
    shapes = resolve_shapes(args["env"], env_args)
    r = OracleRunner(cfg, NumpySyntheticEnv(shapes, 6, seed=1), state_type=shapes["state_type"], seed=1)
    r.warmup()
    before = [{k: v.detach().clone() for k, v in p.items()} for p, _ in r.actors]
    infos, cinfo = r.run_iteration()
    assert np.isfinite(cinfo["value_loss"])
    for a in range(r.A):
        assert all(np.isfinite(float(v)) for v in infos[a].values())
        b = r.abufs[a]
        assert np.isfinite(b["rnn_states"]).all()
        ended = b["masks"][1:, :, 0] == 0.0
        assert np.all(b["rnn_states"][1:][ended] == 0.0)
        if model_over:
            assert np.abs(b["rnn_states"][1:][~ended]).max() > 0  # the GRU state is carried, not reset every step
        else:
            assert np.all(b["rnn_states"] == 0.0)
    moved = any((p[k].detach() - before[a][k]).abs().max() > 0 for a, (p, _) in enumerate(r.actors) for k in p)
    assert moved or algo == "hatrpo"  # a rejected line search restores the parameters

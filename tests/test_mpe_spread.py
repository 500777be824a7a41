"""MPE simple_spread: the NumPy twin against a scalar restatement of the particle world (CPU), and the batched CUDA
env against the twin on identical seeds and actions (GPU).  World model: harl_b200/csrc/mpe_env.cu header; adapter
semantics: harl/envs/pettingzoo_mpe/pettingzoo_mpe_env.py:41-88."""
import math

import numpy as np
import pytest

from harl_b200.envs.mpe_spread import SimpleSpreadNumpy, initial_positions, obs_dim


def scalar_world_step(pos, vel, lm, acts):
    """One world, plain Python floats: returns (pos, vel, per-agent rewards)."""
    A = len(pos)
    u = {0: (0, 0), 1: (-1, 0), 2: (1, 0), 3: (0, -1), 4: (0, 1)}
    f = [[5.0 * u[a][0], 5.0 * u[a][1]] for a in acts]
    for i in range(A):
        for j in range(i + 1, A):
            dx, dy = pos[i][0] - pos[j][0], pos[i][1] - pos[j][1]
            dist = math.sqrt(dx * dx + dy * dy)
            pen = np.logaddexp(0.0, -(dist - 0.3) / 1e-3) * 1e-3
            fx, fy = 1e2 * dx / dist * pen, 1e2 * dy / dist * pen
            f[i][0] += fx; f[i][1] += fy
            f[j][0] -= fx; f[j][1] -= fy
    new_pos, new_vel = [], []
    for i in range(A):
        vx, vy = vel[i][0] * 0.75 + f[i][0] * 0.1, vel[i][1] * 0.75 + f[i][1] * 0.1
        new_vel.append((np.float32(vx), np.float32(vy)))
        new_pos.append((np.float32(pos[i][0] + vx * 0.1), np.float32(pos[i][1] + vy * 0.1)))
    glob = -sum(min(math.dist(p, l) for p in new_pos) for l in lm)
    rew = []
    for i in range(A):
        local = -sum(1.0 for j in range(A) if j != i and math.dist(new_pos[i], new_pos[j]) < 0.3)
        rew.append(0.5 * glob + 0.5 * local)
    return new_pos, new_vel, rew


def test_twin_matches_scalar_world_and_adapter_semantics():
    N, A = 5, 3
    env = SimpleSpreadNumpy(seed=3, n_threads=N, env_args={})
    obs, state, avail = env.reset()
    assert obs.shape == (N, A, 18) and state.shape == (N, A, 54) and avail.shape == (N, A, 5)
    assert obs_dim(3, 3) == 18
    np.testing.assert_array_equal(state[:, 0], obs.reshape(N, -1))   # state = concatenated observations, repeated
    np.testing.assert_array_equal(state[:, 1], state[:, 0])
    rng = np.random.default_rng(0)
    # squeeze two agents of env 0 together so that the contact force and the collision penalty are exercised
    env.pos[0, 1] = env.pos[0, 0] + np.float32(0.1)
    for t in range(1, 27):
        acts = rng.integers(0, 5, (N, A, 1)).astype(np.float32)
        pos0, vel0, lm0 = env.pos.copy(), env.vel.copy(), env.lm.copy()
        obs, state, rew, dones, infos, avail = env.step(acts)
        assert rew.shape == (N, A, 1) and dones.shape == (N, A)
        for n in range(N):
            p, v, r = scalar_world_step([tuple(x) for x in pos0[n]], [tuple(x) for x in vel0[n]], [tuple(x) for x in lm0[n]],
                                        [int(a) for a in acts[n, :, 0]])
            np.testing.assert_allclose(rew[n, :, 0], np.float32(sum(r)), rtol=1e-6, atol=1e-6)   # team reward, broadcast
            if t % 25 != 0:
                np.testing.assert_allclose(env.pos[n], np.array(p, np.float32), rtol=3e-7, atol=1e-7)
                np.testing.assert_allclose(env.vel[n], np.array(v, np.float32), rtol=3e-7, atol=1e-7)
        if t % 25 == 0:   # truncation with bad_transition, auto-reset: the observations are the new episode's
            assert dones.all() and all(i.get("bad_transition") for row in infos for i in row)
            assert (env.vel == 0).all() and (env.episode == t // 25).all() and (env.step_count == 0).all()
            np.testing.assert_array_equal(obs[:, :, 0:2], 0.0)
        else:
            assert not dones.any() and not any(i for row in infos for i in row)
        # observation layout: own velocity, own position, landmarks, other agents, comm zeros
        np.testing.assert_array_equal(obs[:, 1, 0:2], env.vel[:, 1])
        np.testing.assert_array_equal(obs[:, 1, 2:4], env.pos[:, 1])
        np.testing.assert_array_equal(obs[:, 1, 4:10], (env.lm - env.pos[:, 1:2]).reshape(N, 6))
        np.testing.assert_array_equal(obs[:, 1, 10:12], env.pos[:, 0] - env.pos[:, 1])
        np.testing.assert_array_equal(obs[:, 1, 12:14], env.pos[:, 2] - env.pos[:, 1])
        np.testing.assert_array_equal(obs[:, 1, 14:18], 0.0)


def test_initial_positions_are_a_function_of_seed_env_and_episode():
    p1, l1 = initial_positions(7, [0, 1, 2], [0, 0, 0], 3, 3)
    p2, l2 = initial_positions(7, [2, 1, 0], [0, 0, 0], 3, 3)
    np.testing.assert_array_equal(p1, p2[::-1])
    np.testing.assert_array_equal(l1, l2[::-1])
    p3, _ = initial_positions(7, [0, 1, 2], [1, 1, 1], 3, 3)
    assert not np.array_equal(p1, p3)
    assert (np.abs(p1) <= 1).all() and (np.abs(l1) <= 1).all()
    # spread over the square (not a constant stream)
    big, _ = initial_positions(11, np.arange(4096), np.zeros(4096), 3, 3)
    assert abs(float(big.mean())) < 0.03 and 0.30 < float((big ** 2).mean()) < 0.37


@pytest.mark.gpu
def test_device_env_equals_twin_on_identical_seeds_and_actions():
    import torch

    from harl_b200.envs.mpe_spread import BatchedSimpleSpread

    N, A = 257, 3
    dev = torch.device("cuda:0")
    genv = BatchedSimpleSpread(5, N, {}, dev)
    tenv = SimpleSpreadNumpy(5, N, {})
    go, gs, ga = genv.reset()
    to, ts, ta = tenv.reset()
    np.testing.assert_array_equal(go.cpu().numpy(), to)          # same Philox stream -> same worlds, bit for bit
    np.testing.assert_array_equal(gs.cpu().numpy(), ts)
    rng = np.random.default_rng(1)
    for t in range(1, 61):
        acts = rng.integers(0, 5, (N, A, 1)).astype(np.float32)
        go, gs, gr, gd, gi, ga = genv.step(torch.from_numpy(acts).to(dev))
        to, ts, tr, td, ti, ta = tenv.step(acts)
        np.testing.assert_array_equal(gd.cpu().numpy(), td)
        bad_g = np.array([[bool(i.get("bad_transition", False)) for i in gi[n]] for n in range(N)])
        bad_t = np.array([[bool(i.get("bad_transition", False)) for i in row] for row in ti])
        np.testing.assert_array_equal(bad_g, bad_t)
        np.testing.assert_allclose(go.cpu().numpy(), to, rtol=0, atol=2e-6)   # fp64 arithmetic on both sides; libm vs CUDA exp/log
        np.testing.assert_allclose(gs.cpu().numpy(), ts, rtol=0, atol=2e-6)
        np.testing.assert_allclose(gr.cpu().numpy(), tr, rtol=1e-5, atol=1e-5)
        # re-synchronise the twin to the device state so that rounding differences do not accumulate over episodes
        st = genv.get_state()
        tenv.pos, tenv.vel, tenv.lm = st["pos"].copy(), st["vel"].copy(), st["lm"].copy()


@pytest.mark.gpu
def test_happo_trains_on_the_native_env_through_the_runner():
    """Three iterations through the public Runner on the native backend (zero-copy step_into, CUDA-graph rollout):
    finite losses, episodes finish every 25 steps, the logger's episode return equals 25 steps of team reward."""
    import tempfile

    import torch

    from harl_b200.runners import RUNNER_REGISTRY
    from harl_b200.utils.configs_tools import get_defaults_yaml_args

    algo_args, env_args = get_defaults_yaml_args("happo", "pettingzoo_mpe")
    env_args.update(scenario="simple_spread_v2", continuous_actions=False, backend="native")
    algo_args["train"].update(n_rollout_threads=64, episode_length=50, num_env_steps=10**9, log_interval=10**9, eval_interval=10**9)
    algo_args["eval"]["use_eval"] = False
    algo_args["logger"]["log_dir"] = tempfile.mkdtemp(prefix="hb_mpe_")
    r = RUNNER_REGISTRY["happo"](dict(algo="happo", env="pettingzoo_mpe", exp_name="t"), algo_args, env_args)
    r.warmup()
    r.logger.init(10**9)
    for ep in range(1, 4):
        r.run_iteration(ep, 10**9)
    torch.cuda.synchronize()
    infos, cinfo = r.last_train_infos
    for i in infos:
        assert all(math.isfinite(float(v)) for v in i.values())
    assert math.isfinite(float(cinfo["value_loss"]))
    assert int(r.envs.episode.min()) == int(r.envs.episode.max()) == 6    # 150 steps / 25
    done_sum = r._done_sum.cpu().numpy()
    assert done_sum[1] == 64 * 2 or done_sum[1] == 64 * 6   # per-iteration or cumulative counter, both are whole episodes
    rew = r.critic_buffer.rewards[:, :, 0]
    assert torch.isfinite(rew).all() and float(rew.max()) < 0.0   # distances are positive: the team reward is negative
    r.close()

"""Training / evaluation logger with the reference's callback surface
(harl/common/base_logger.py:30-184): init, episode_init, per_step, episode_log, eval_*, close.

Episode-return bookkeeping stays on the device (one host read per ``episode_log``); the printed
FPS keeps the reference definition: episode * T * N / elapsed (agents do not multiply it)."""
import os
import time

import numpy as np
import torch

from ..utils.configs_tools import get_task_name


class OnPolicyLogger:
    def __init__(self, args, algo_args, env_args, num_agents, writter, run_dir):
        self.args, self.algo_args, self.env_args = args, algo_args, env_args
        self.task_name = self.get_task_name()
        self.num_agents = num_agents
        self.writter = writter
        self.run_dir = run_dir
        self.log_file = open(os.path.join(run_dir, "progress.txt"), "w", encoding="utf-8")
        self.world = 1

    def get_task_name(self):
        return get_task_name(self.args["env"], self.env_args)

    def init(self, episodes):
        self.start = time.time()
        self.episodes = episodes
        self.train_episode_rewards = None
        self.done_sum = getattr(self, "_device_done_sum", None)
        self.fps = 0

    def episode_init(self, episode):
        self.episode = episode

    def attach_device_stats(self, done_sum):
        """The zero-copy rollout keeps (sum of finished episode returns, count) in a device double[2] updated by the
        insert kernel; ``per_step`` is then not called."""
        self.done_sum = self._device_done_sum = done_sum

    def per_step(self, data):
        """Accumulate per-env episode returns (mean over agents of the summed reward), device-side."""
        rewards, dones = data[2], data[3]
        r = torch.as_tensor(rewards)
        d = torch.as_tensor(dones)
        if self.train_episode_rewards is None:
            self.train_episode_rewards = torch.zeros(r.shape[0], dtype=torch.float32, device=r.device)
            self.done_sum = torch.zeros(2, dtype=torch.float64, device=r.device)  # (sum of returns, count)
        self.train_episode_rewards += r.reshape(r.shape[0], -1).mean(1)
        de = d.reshape(d.shape[0], -1).all(1)
        self.done_sum[0] += (self.train_episode_rewards * de).sum()
        self.done_sum[1] += de.sum()
        self.train_episode_rewards *= (~de)

    def episode_log(self, actor_train_infos, critic_train_info, actor_buffer, critic_buffer):
        T = self.algo_args["train"]["episode_length"]
        N = self.algo_args["train"]["n_rollout_threads"]
        self.total_num_steps = self.episode * T * N
        self.end = time.time()
        self.fps = int(self.total_num_steps / max(self.end - self.start, 1e-9))
        print("Env {} Task {} Algo {} Exp {} updates {}/{} episodes, total num timesteps {}/{}, FPS {}.".format(
            self.args["env"], self.task_name, self.args["algo"], self.args["exp_name"], self.episode, self.episodes,
            self.total_num_steps, self.algo_args["train"]["num_env_steps"], self.fps))
        critic_train_info["average_step_rewards"] = critic_buffer.get_mean_rewards()
        self.log_train(actor_train_infos, critic_train_info)
        print("Average step reward is {}.".format(critic_train_info["average_step_rewards"]))
        if self.done_sum is not None:
            s, c = self.done_sum.cpu().numpy()
            if c > 0:
                aver = float(s / c)
                print("Some episodes done, average episode reward is {}.\n".format(aver))
                self.writter.add_scalars("train_episode_rewards", {"aver_rewards": aver}, self.total_num_steps)
                self.last_average_episode_reward = aver
            self.done_sum.zero_()

    def log_train(self, actor_train_infos, critic_train_info):
        for agent_id, info in enumerate(actor_train_infos):
            for k, v in info.items():
                self.writter.add_scalars(f"agent{agent_id}/{k}", {f"agent{agent_id}/{k}": float(v)}, self.total_num_steps)
        for k, v in critic_train_info.items():
            self.writter.add_scalars(f"critic/{k}", {f"critic/{k}": float(v)}, self.total_num_steps)

    # ---- evaluation callbacks (base_logger.py:96-162)
    def eval_init(self):
        self.total_num_steps = self.episode * self.algo_args["train"]["episode_length"] * self.algo_args["train"]["n_rollout_threads"]
        n = self.algo_args["eval"]["n_eval_rollout_threads"]
        self.eval_episode_rewards = [[] for _ in range(n)]
        self.one_episode_rewards = [[] for _ in range(n)]

    def eval_per_step(self, eval_data):
        rewards = eval_data[2]
        r = rewards.detach().cpu().numpy() if torch.is_tensor(rewards) else np.asarray(rewards)
        for i in range(len(self.one_episode_rewards)):
            self.one_episode_rewards[i].append(r[i])
        self.eval_infos = eval_data[4]

    def eval_thread_done(self, tid):
        self.eval_episode_rewards[tid].append(np.sum(self.one_episode_rewards[tid], axis=0))
        self.one_episode_rewards[tid] = []

    def eval_log(self, eval_episode):
        rews = np.concatenate([r for r in self.eval_episode_rewards if r])
        avg = float(np.mean(rews))
        self.writter.add_scalars("eval_average_episode_rewards", {"eval_average_episode_rewards": avg}, self.total_num_steps)
        print("Evaluation average episode reward is {}.\n".format(avg))
        self.log_file.write(",".join(map(str, [self.total_num_steps, avg])) + "\n")
        self.log_file.flush()
        self.last_eval_reward = avg

    def close(self):
        self.log_file.close()

"""On-policy runner (reference: harl/runners/on_policy_base_runner.py:26-775).

Same constructor signature, attributes and method names (run / warmup / collect / insert /
compute / train / after_update / eval / save / restore / close), same config keys.  What
changes is where the data lives: buffers, networks, optimiser state, ValueNorm statistics and
the env outputs are device tensors for the whole run; ``collect`` writes actions, log-probs and
values straight into their buffer slots, ``insert`` derives the masks with one kernel, and the
update never materialises a minibatch.  With ``torch.distributed`` initialised the rollout
threads are sharded over the ranks (one process per GPU) and gradients / normalisers are
sum-allreduced.
"""
import ctypes as C
import os

import torch

from .. import _lib as L
from .. import dist
from ..algorithms.actors import ALGO_REGISTRY
from ..algorithms.critics.v_critic import VCritic
from ..common.buffers.on_policy_actor_buffer import OnPolicyActorBuffer
from ..common.buffers.on_policy_critic_buffer_ep import OnPolicyCriticBufferEP
from ..common.buffers.on_policy_critic_buffer_fp import OnPolicyCriticBufferFP
from ..common.valuenorm import ValueNorm
from ..envs import LOGGER_REGISTRY
from ..utils.configs_tools import init_dir, save_config
from ..utils.envs_tools import get_num_agents, make_eval_env, make_train_env, set_seed
from ..utils.models_tools import init_device
from ..utils.trans_tools import _t2n


class OnPolicyBaseRunner:
    def __init__(self, args, algo_args, env_args):
        self.args, self.algo_args, self.env_args = args, algo_args, env_args
        self.hidden_sizes = algo_args["model"]["hidden_sizes"]
        self.rnn_hidden_size = self.hidden_sizes[-1]
        self.recurrent_n = algo_args["model"]["recurrent_n"]
        self.action_aggregation = algo_args["algo"]["action_aggregation"]
        self.state_type = env_args.get("state_type", "EP")
        self.share_param = algo_args["algo"]["share_param"]
        self.fixed_order = algo_args["algo"]["fixed_order"]
        if os.environ.get("HB_OVERLAP_CRITIC") is not None:   # the critic update on a side stream (default on); 0 = one stream
            self.overlap_critic_update = os.environ["HB_OVERLAP_CRITIC"] != "0"
        set_seed(algo_args["seed"])
        self.device = init_device(algo_args["device"])
        if dist.world_size() > 1:
            # every replica must start from the same weights and draw the same agent orders / permutations: all of them come
            # from the torch CPU generator, so every rank adopts rank 0's seed (with seed_specify = false each rank would
            # otherwise draw its own) and re-seeds
            seed_t = torch.tensor([int(algo_args["seed"]["seed"])], dtype=torch.int64, device=self.device)
            torch.distributed.broadcast(seed_t, src=0)
            if int(seed_t.item()) != int(algo_args["seed"]["seed"]):
                algo_args["seed"]["seed"] = int(seed_t.item())
                set_seed({**algo_args["seed"], "seed_specify": True})
        if self.share_param and (algo_args["model"].get("use_recurrent_policy") or algo_args["model"].get("use_naive_recurrent_policy")):
            # fail before the first rollout, not at the first train() (the reference accepts this combination: README / DESIGN.md)
            raise NotImplementedError("share_param with recurrent (GRU) policies is not implemented: the reference interleaves the "
                                      "agents' sequences when it concatenates their minibatches (mappo.py:149-222)")
        T_, L_ = algo_args["train"]["episode_length"], algo_args["model"].get("data_chunk_length", 1)
        if algo_args["model"].get("use_recurrent_policy") and T_ % L_ != 0:
            raise NotImplementedError(f"episode_length ({T_}) must be a multiple of data_chunk_length ({L_}) for recurrent policies")
        if algo_args["render"]["use_render"]:
            raise NotImplementedError("rendering needs the third-party simulators, which are out of scope")
        self.world, self.rank = dist.world_size(), dist.rank()
        self.n_global = algo_args["train"]["n_rollout_threads"]
        lo, hi = dist.shard_bounds(self.n_global, self.world, self.rank)
        self.n_local = hi - lo
        self.run_dir, self.log_dir, self.save_dir, self.writter = init_dir(
            args["env"], env_args, args["algo"], args["exp_name"] + (f"-rank{self.rank}" if self.world > 1 else ""),
            algo_args["seed"]["seed"], logger_path=algo_args["logger"]["log_dir"])
        save_config(args, algo_args, env_args, self.run_dir)
        try:
            import setproctitle

            setproctitle.setproctitle(f"{args['algo']}-{args['env']}-{args['exp_name']}")
        except ImportError:
            pass

        seed = algo_args["seed"]["seed"]
        self.envs = make_train_env(args["env"], seed + 1000 * self.rank, self.n_local, env_args, self.device)
        self.eval_envs = (make_eval_env(args["env"], seed, algo_args["eval"]["n_eval_rollout_threads"], env_args,
                                        self.device) if algo_args["eval"]["use_eval"] else None)
        self.num_agents = get_num_agents(args["env"], env_args, self.envs)
        if self.num_agents > L.HB_MAX_AGENTS:
            raise NotImplementedError(f"more than {L.HB_MAX_AGENTS} agents")
        if self.rank == 0:
            print("share_observation_space: ", self.envs.share_observation_space)
            print("observation_space: ", self.envs.observation_space)
            print("action_space: ", self.envs.action_space)

        actor_args = {**algo_args["model"], **algo_args["algo"]}
        if self.share_param:
            first = ALGO_REGISTRY[args["algo"]](actor_args, self.envs.observation_space[0], self.envs.action_space[0],
                                                device=self.device)
            for a in range(1, self.num_agents):
                assert self.envs.observation_space[a] == self.envs.observation_space[0], \
                    "Agents have heterogeneous observation spaces, parameter sharing is not valid."
                assert self.envs.action_space[a] == self.envs.action_space[0], \
                    "Agents have heterogeneous action spaces, parameter sharing is not valid."
            self.actor = [first] * self.num_agents
        else:
            self.actor = [ALGO_REGISTRY[args["algo"]](actor_args, self.envs.observation_space[a],
                                                      self.envs.action_space[a], device=self.device)
                          for a in range(self.num_agents)]

        for i, a in enumerate(self.actor):  # sampling streams: a function of (seed, rank, agent) only
            a._seed = ((int(seed) * 1000003 + 7919 * self.rank + i + 1) * 2654435761) & (2**63 - 1)
        train_local = {**algo_args["train"], "n_rollout_threads": self.n_local}
        self.actor_buffer = [OnPolicyActorBuffer({**train_local, **algo_args["model"]}, self.envs.observation_space[a],
                                                 self.envs.action_space[a], device=self.device)
                             for a in range(self.num_agents)]
        share_space = self.envs.share_observation_space[0]
        self.critic = VCritic(actor_args, share_space, device=self.device)
        cb_args = {**train_local, **algo_args["model"], **algo_args["algo"]}
        if self.state_type == "EP":
            self.critic_buffer = OnPolicyCriticBufferEP(cb_args, share_space, device=self.device)
        elif self.state_type == "FP":
            self.critic_buffer = OnPolicyCriticBufferFP(cb_args, share_space, self.num_agents, device=self.device)
        else:
            raise NotImplementedError
        self.value_normalizer = ValueNorm(1, device=self.device) if algo_args["train"]["use_valuenorm"] else None
        self.logger = LOGGER_REGISTRY[args["env"]](args, algo_args, env_args, self.num_agents, self.writter, self.run_dir)
        self.timers = {}
        if algo_args["train"]["model_dir"] is not None:
            self.restore()

    # ------------------------------------------------------------------ main loop
    def run(self):
        if self.rank == 0:
            print("start running")
        self.warmup()
        T = self.algo_args["train"]["episode_length"]
        episodes = int(self.algo_args["train"]["num_env_steps"]) // T // self.n_global
        self.logger.init(episodes)
        for episode in range(1, episodes + 1):
            self.run_iteration(episode, episodes)

    def run_iteration(self, episode, episodes):
        """One training iteration: T x (collect -> env.step -> insert) -> compute -> train -> after_update."""
        tr = self.algo_args["train"]
        if tr["use_linear_lr_decay"]:
            for a in (self.actor[:1] if self.share_param else self.actor):
                a.lr_decay(episode, episodes)
            self.critic.lr_decay(episode, episodes)
        self.logger.episode_init(episode)
        self.prep_rollout()
        marks = self._phase_mark(None, None)
        fast = self._fast_rollout_ready()
        for step in range(tr["episode_length"] if not fast else 0):
            values, actions, action_log_probs, rnn_states, rnn_states_critic = self.collect(step)
            obs, share_obs, rewards, dones, infos, available_actions = self.envs.step(actions)
            data = (obs, share_obs, rewards, dones, infos, available_actions, values, actions, action_log_probs,
                    rnn_states, rnn_states_critic)
            self.logger.per_step(data)
            self.insert(data)
        if fast:
            self._fast_rollout()
        marks = self._phase_mark(marks, "rollout")
        self.compute()
        marks = self._phase_mark(marks, "compute")
        self.prep_training()
        actor_train_infos, critic_train_info = self.train()
        marks = self._phase_mark(marks, "train")
        if episode % tr["log_interval"] == 0 and self.rank == 0:
            self.logger.episode_log(actor_train_infos, critic_train_info, self.actor_buffer, self.critic_buffer)
        if episode % tr["eval_interval"] == 0:
            if self.algo_args["eval"]["use_eval"]:
                self.prep_rollout()
                self.eval()
            if self.rank == 0:
                self.save()
        self.after_update()
        self._phase_mark(marks, "after_update", final=True)
        self.last_train_infos = (actor_train_infos, critic_train_info)

    def _phase_mark(self, marks, name, final=False):
        """CUDA-event phase timers (only when ``self.time_phases`` is set; one sync at the end of the iteration)."""
        if not getattr(self, "time_phases", False) or self.device.type != "cuda":
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks = (marks or []) + [(name, ev)]
        if final:
            torch.cuda.synchronize()
            self.phase_ms = {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1]) for i in range(1, len(marks))}
        return marks

    # ------------------------------------------------------------------ zero-copy rollout (device-resident envs)
    def _fast_rollout_ready(self):
        """The lean rollout loop applies when the env can write its outputs straight into the buffer slots
        (``step_into``: a device-resident env, or a host env staging through pinned memory); otherwise the generic
        collect/step/insert loop runs."""
        if getattr(self, "_fast", None) is None:
            ok = (self.device.type == "cuda" and hasattr(self.envs, "step_into")
                  and (getattr(self.envs, "device", None) == self.device or getattr(self.envs, "host_staged", False))
                  and not getattr(self, "disable_fast_rollout", False))
            self._fast = self._build_fast_path() if ok else False
        return bool(self._fast)

    def _build_fast_path(self):
        """Per-step argument structs with every slot pointer baked in (the buffers never move)."""
        T = self.algo_args["train"]["episode_length"]
        N, A = self.n_local, self.num_agents
        cb = self.critic_buffer
        fp = self.state_type == "FP"
        team = bool(getattr(self.envs, "team_reward", False))
        self._ep_return = torch.zeros(N, dtype=torch.float32, device=self.device)
        self._draw_ctr = torch.zeros(1, dtype=torch.int64, device=self.device)  # Philox offset base of graph replays
        self._draws = getattr(self, "_draws", 0)
        self._done_sum = torch.zeros(2, dtype=torch.float64, device=self.device)
        self._dones_u8 = torch.zeros(N, A, dtype=torch.uint8, device=self.device)
        self._bad_u8 = torch.zeros(N, A, dtype=torch.uint8, device=self.device)
        self._rewards_na = torch.zeros(N, A, dtype=torch.float32, device=self.device)
        if hasattr(self.logger, "attach_device_stats"):
            self.logger.attach_device_stats(self._done_sum)
        nets = [a.actor for a in self.actor] + [self.critic.critic]
        ws_bytes = max(L.lib.hb_workspace_bytes(C.byref(n.desc), N * (A if fp else 1), 0) for n in nets)
        from ..nets import workspace

        self._fast_ws = workspace(self.device, ws_bytes)
        collect, insert, dst = [], [], []
        for s in range(T):
            c = L.CollectArgs()
            c.n_agents, c.deterministic, c.rows = A, 0, N
            for i, (actor, b) in enumerate(zip(self.actor, self.actor_buffer)):
                c.actor_desc[i] = C.pointer(actor.actor.desc)
                c.actor_prepared[i] = L.ptr(actor.actor.prepared)
                c.obs[i] = L.ptr(b.obs[s])
                c.avail[i] = L.ptr(b.available_actions[s]) if b.available_actions is not None else None
                c.actions[i], c.logp[i] = L.ptr(b.actions[s]), L.ptr(b.action_log_probs[s])
                # a shared actor object (share_param) still needs one sampling stream per agent
                c.seed[i] = (actor._seed + (i * 0x9E3779B97F4A7C15 if self.share_param else 0)) & (2**64 - 1)
            c.critic_desc = C.pointer(self.critic.critic.desc)
            c.critic_prepared = L.ptr(self.critic.critic.prepared)
            c.share_obs, c.critic_rows, c.values = L.ptr(cb.share_obs[s]), cb.value_preds[s].numel(), L.ptr(cb.value_preds[s])
            # recurrent nets: hidden state in from slot s, out to slot s + 1 (the insert kernel zeroes finished envs)
            for i, b in enumerate(self.actor_buffer):
                if b.recurrent:
                    c.actor_rnn[i], c.actor_rnn_out[i] = L.ptr(b.rnn_states[s]), L.ptr(b.rnn_states[s + 1])
                    c.actor_masks[i] = L.ptr(b.masks[s])
            if cb.recurrent:
                c.critic_rnn, c.critic_rnn_out = L.ptr(cb.rnn_states_critic[s]), L.ptr(cb.rnn_states_critic[s + 1])
                c.critic_masks = L.ptr(cb.masks[s])
            collect.append(c)
            a = L.InsertArgs()
            a.n_envs, a.n_agents, a.state_type_fp = N, A, int(fp)
            a.dones, a.bad_transition = L.ptr(self._dones_u8), L.ptr(self._bad_u8)
            rec = self.actor_buffer[0].recurrent
            a.actor_rnn_row = self.recurrent_n * self.rnn_hidden_size if rec else 0
            a.critic_rnn_row = self.recurrent_n * self.rnn_hidden_size if cb.recurrent else 0
            a.critic_rnn_next = L.ptr(cb.rnn_states_critic[s + 1]) if cb.recurrent else None
            for i, b in enumerate(self.actor_buffer):
                a.actor_rnn_next[i] = L.ptr(b.rnn_states[s + 1]) if rec else None
                a.actor_masks_next[i], a.actor_active_next[i] = L.ptr(b.masks[s + 1]), L.ptr(b.active_masks[s + 1])
            a.critic_masks_next, a.critic_bad_next = L.ptr(cb.masks[s + 1]), L.ptr(cb.bad_masks[s + 1])
            # logger bookkeeping reads the reward where it already is: the critic slot (team reward, or FP per-agent
            # rewards) or a side buffer for per-agent rewards under an EP critic
            if fp:
                a.rewards, a.reward_stride_n, a.reward_stride_a = L.ptr(cb.rewards[s]), A, 1
            elif team:
                a.rewards, a.reward_stride_n, a.reward_stride_a = L.ptr(cb.rewards[s]), 1, 0
            else:
                a.rewards, a.reward_stride_n, a.reward_stride_a = L.ptr(self._rewards_na), A, 1
            a.ep_return, a.done_sum = L.ptr(self._ep_return), L.ptr(self._done_sum)
            insert.append(a)
            dst.append(dict(obs=[b.obs[s + 1] for b in self.actor_buffer], share_obs=cb.share_obs[s + 1],
                            rewards=cb.rewards[s], rewards_na=None if (fp or team) else self._rewards_na,
                            avail=[b.available_actions[s + 1] if b.available_actions is not None else None
                                   for b in self.actor_buffer],
                            actions=[b.actions[s] for b in self.actor_buffer], dones=self._dones_u8, bad=self._bad_u8))
        return dict(collect=collect, insert=insert, dst=dst)

    def _fast_rollout(self):
        """T x (collect -> env.step_into -> insert): two library calls per step plus the env's own work.

        When the env's host-side control flow repeats every ``graph_period()`` steps and T is a multiple of it, the
        whole T-step rollout is captured ONCE into a CUDA graph and replayed per iteration (the loop is launch-bound:
        ~3 small kernels per step against ~100 us of Python).  The sampling streams stay fresh through a device
        counter that the collect kernel adds to its per-step Philox offset and the graph's last node advances."""
        f = self._fast
        T = len(f["collect"])
        period = getattr(self.envs, "graph_period", lambda: None)()
        use_graph = (period is not None and T % period == 0 and getattr(self, "use_cuda_graph_rollout", True)
                     and not getattr(self, "time_phases_no_graph", False))
        if use_graph and f.get("graph") is not None:
            f["graph"].replay()
            self._draws += T
            self.envs.graph_advance(T)
            self.graph_replayed_launches = getattr(self, "graph_replayed_launches", 0) + f["graph_kernels"]
        elif use_graph and f.get("eager_runs", 0) >= 1:
            # capture (the first iteration ran eagerly: every lazy initialisation is done)
            self._draw_ctr.fill_(self._draws)
            for s in range(T):
                f["collect"][s].offset = s + 1
                f["collect"][s].offset_base = L.ptr(self._draw_ctr)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L.lib.hb_kernel_launch_count()
            with torch.cuda.graph(g):
                self._rollout_steps(f, L.stream_ptr(), bump=False)
                L.call("hb_counter_add", L.ptr(self._draw_ctr), T, L.stream_ptr())
            f["graph"] = g
            f["graph_kernels"] = int(L.lib.hb_kernel_launch_count() - n0)  # library kernels per replay (env copies not counted)
            g.replay()
            self._draws += T
        else:
            self._rollout_steps(f, L.stream_ptr(), bump=True)
            f["eager_runs"] = f.get("eager_runs", 0) + 1
        for b in self.actor_buffer:
            b.step = 0
        self.critic_buffer.step = 0

    def _rollout_steps(self, f, st, bump):
        ws, ws_n = L.ptr(self._fast_ws), self._fast_ws.numel()
        step_into = self.envs.step_into
        for s in range(len(f["collect"])):
            c = f["collect"][s]
            if bump:
                self._draws += 1
                c.offset = self._draws
            L.call("hb_rollout_collect", C.byref(c), ws, ws_n, st)
            step_into(f["dst"][s])
            L.call("hb_rollout_insert_masks", C.byref(f["insert"][s]), st)

    def warmup(self):
        """Reset the envs and fill slot 0 (reference :269-283)."""
        obs, share_obs, available_actions = self.envs.reset()
        obs = torch.as_tensor(obs, device=self.device)
        share_obs = torch.as_tensor(share_obs, device=self.device)
        for a in range(self.num_agents):
            self.actor_buffer[a].obs[0].copy_(obs[:, a])
            if self.actor_buffer[a].available_actions is not None:
                self.actor_buffer[a].available_actions[0].copy_(torch.as_tensor(available_actions, device=self.device)[:, a])
        if self.state_type == "EP":
            self.critic_buffer.share_obs[0].copy_(share_obs[:, 0])
        else:
            self.critic_buffer.share_obs[0].copy_(share_obs)

    @torch.no_grad()
    def collect(self, step):
        """Actions / log-probs / values for one rollout step, written straight into slot ``step``
        of the buffers (reference :285-340).  Returns the reference 5-tuple as device tensors."""
        new_rnn = []
        for a in range(self.num_agents):
            b = self.actor_buffer[a]
            _, _, rnn_a = self.actor[a].get_actions(b.obs[step], b.rnn_states[step], b.masks[step],
                                                    b.available_actions[step] if b.available_actions is not None else None,
                                                    actions_out=b.actions[step], logp_out=b.action_log_probs[step])
            new_rnn.append(rnn_a)
        actions = torch.stack([b.actions[step] for b in self.actor_buffer], dim=1)
        action_log_probs = torch.stack([b.action_log_probs[step] for b in self.actor_buffer], dim=1)
        rnn_states = torch.stack(new_rnn, dim=1) if self.actor_buffer[0].recurrent else None  # the GRUs' new states
        cb = self.critic_buffer
        sd = cb.share_obs.shape[-1]
        if cb.recurrent:
            R, h = cb.rnn_states_critic.shape[-2:]
            values, rnn_c = self.critic.get_values(cb.share_obs[step].reshape(-1, sd),
                                                   cb.rnn_states_critic[step].reshape(-1, R, h), cb.masks[step].reshape(-1, 1),
                                                   values_out=cb.value_preds[step].reshape(-1, 1))
            rnn_states_critic = rnn_c.reshape(cb.rnn_states_critic[step].shape)
        else:
            self.critic.get_values(cb.share_obs[step].reshape(-1, sd), None, None,
                                   values_out=cb.value_preds[step].reshape(-1, 1))
            rnn_states_critic = None
        values = cb.value_preds[step]
        self._in_place = (actions, action_log_probs)  # already sitting in their buffer slots
        return values, actions, action_log_probs, rnn_states, rnn_states_critic

    def _bad_transition_flags(self, infos):
        """[N, A] uint8 ``bad_transition`` flags from a batched env (device tensor) or a list of dict lists."""
        bad = getattr(infos, "_bad_dev", None)
        if bad is None:
            bad = torch.tensor([[bool(i.get("bad_transition", False)) for i in row] for row in infos])
        return bad.to(device=self.device, dtype=torch.uint8).contiguous()

    def insert(self, data):
        """Mask derivation (one kernel) + slot writes (reference :342-460)."""
        (obs, share_obs, rewards, dones, infos, available_actions, values, actions, action_log_probs, rnn_states,
         rnn_states_critic) = data
        dev = self.device
        cb = self.critic_buffer
        s = cb.step
        dones_u8 = torch.as_tensor(dones).to(device=dev, dtype=torch.uint8).contiguous()
        bad_u8 = self._bad_transition_flags(infos)
        a = L.InsertArgs()
        a.n_envs, a.n_agents, a.state_type_fp = self.n_local, self.num_agents, int(self.state_type == "FP")
        rec = self.actor_buffer[0].recurrent
        a.actor_rnn_row = self.recurrent_n * self.rnn_hidden_size if rec else 0
        a.critic_rnn_row = self.recurrent_n * self.rnn_hidden_size if cb.recurrent else 0
        a.dones, a.bad_transition = L.ptr(dones_u8), L.ptr(bad_u8)
        for i, b in enumerate(self.actor_buffer):
            if rec and rnn_states is not None:
                b.rnn_states[s + 1].copy_(torch.as_tensor(rnn_states, device=dev)[:, i])
            a.actor_masks_next[i] = L.ptr(b.masks[s + 1])
            a.actor_active_next[i] = L.ptr(b.active_masks[s + 1])
            a.actor_rnn_next[i] = L.ptr(b.rnn_states[s + 1]) if rec else None
        if cb.recurrent and rnn_states_critic is not None:
            cb.rnn_states_critic[s + 1].copy_(torch.as_tensor(rnn_states_critic, device=dev))
        a.critic_masks_next, a.critic_bad_next = L.ptr(cb.masks[s + 1]), L.ptr(cb.bad_masks[s + 1])
        a.critic_rnn_next = L.ptr(cb.rnn_states_critic[s + 1]) if cb.recurrent else None
        L.call("hb_rollout_insert_masks", C.byref(a), L.stream_ptr())
        obs = torch.as_tensor(obs, device=dev)
        avail = None if available_actions is None or (not torch.is_tensor(available_actions)
                                                      and available_actions[0] is None) \
            else torch.as_tensor(available_actions, device=dev)
        in_place = getattr(self, "_in_place", (None, None))
        skip = actions is in_place[0] and action_log_probs is in_place[1]
        for i, b in enumerate(self.actor_buffer):
            b.insert(obs[:, i], None, None if skip else torch.as_tensor(actions, device=dev)[:, i],
                     None if skip else torch.as_tensor(action_log_probs, device=dev)[:, i], None, None,
                     avail[:, i] if avail is not None else None)
        share_obs = torch.as_tensor(share_obs, device=dev)
        rewards = torch.as_tensor(rewards, device=dev)
        if self.state_type == "EP":
            cb.insert(share_obs[:, 0], None, values, rewards[:, 0], None, None)
        else:
            cb.insert(share_obs, None, values, rewards, None, None)

    @torch.no_grad()
    def compute(self):
        """Bootstrap value of slot T, then the GAE / return kernel (reference :462-484)."""
        cb = self.critic_buffer
        sd = cb.share_obs.shape[-1]
        if cb.recurrent:
            R, h = cb.rnn_states_critic.shape[-2:]
            next_value, _ = self.critic.get_values(cb.share_obs[-1].reshape(-1, sd),
                                                   cb.rnn_states_critic[-1].reshape(-1, R, h), cb.masks[-1].reshape(-1, 1))
        else:
            next_value, _ = self.critic.get_values(cb.share_obs[-1].reshape(-1, sd), None, None)
        cb.compute_returns(next_value, self.value_normalizer)

    def train(self):
        raise NotImplementedError

    def after_update(self):
        for b in self.actor_buffer:
            b.after_update()
        self.critic_buffer.after_update()

    # ------------------------------------------------------------------ evaluation / checkpoints
    @torch.no_grad()
    def eval(self):
        """Deterministic evaluation on ``eval_envs`` until ``eval_episodes`` finish (reference :500-591)."""
        if self.eval_envs is None:
            return
        self.logger.eval_init()
        n = self.algo_args["eval"]["n_eval_rollout_threads"]
        eval_episode = 0
        obs, share_obs, avail = self.eval_envs.reset()
        rec = self.actor_buffer[0].recurrent
        eval_rnn = [torch.zeros(n, self.recurrent_n, self.rnn_hidden_size, device=self.device) for _ in range(self.num_agents)] \
            if rec else [None] * self.num_agents
        eval_masks = torch.ones(n, 1, device=self.device) if rec else None
        while True:
            acts = []
            obs_t = torch.as_tensor(obs, device=self.device)
            for a in range(self.num_agents):
                av = None if avail is None else torch.as_tensor(avail, device=self.device)[:, a].contiguous()
                act, eval_rnn[a] = self.actor[a].act(obs_t[:, a].contiguous(), eval_rnn[a], eval_masks, av, deterministic=True)
                acts.append(act)
            actions = torch.stack(acts, dim=1)
            obs, share_obs, rewards, dones, infos, avail = self.eval_envs.step(actions)
            self.logger.eval_per_step((obs, share_obs, rewards, dones, infos, avail))
            dones_env = _t2n(torch.as_tensor(dones)).all(axis=1)
            if rec:  # finished envs restart from a zero state (reference :560-575)
                done_t = torch.as_tensor(dones_env, device=self.device)
                for a in range(self.num_agents):
                    eval_rnn[a][done_t] = 0.0
                eval_masks = (~done_t).float().reshape(n, 1)
            for i in range(n):
                if dones_env[i]:
                    eval_episode += 1
                    self.logger.eval_thread_done(i)
            if eval_episode >= self.algo_args["eval"]["eval_episodes"]:
                self.logger.eval_log(eval_episode)
                break

    def prep_rollout(self):
        for a in self.actor:
            a.prep_rollout()
        self.critic.prep_rollout()

    def prep_training(self):
        for a in self.actor:
            a.prep_training()
        self.critic.prep_training()

    def save(self):
        """Reference file names and state_dict keys (:724-740), so checkpoints interchange."""
        for a in range(self.num_agents):
            torch.save({k: v.cpu() for k, v in self.actor[a].actor.state_dict().items()},
                       os.path.join(self.save_dir, f"actor_agent{a}.pt"))
        torch.save({k: v.cpu() for k, v in self.critic.critic.state_dict().items()},
                   os.path.join(self.save_dir, "critic_agent.pt"))
        if self.value_normalizer is not None:
            torch.save({k: v.cpu() for k, v in self.value_normalizer.state_dict().items()},
                       os.path.join(self.save_dir, "value_normalizer.pt"))

    def restore(self):
        """Reference :742-763."""
        md = str(self.algo_args["train"]["model_dir"])
        for a in range(self.num_agents):
            self.actor[a].actor.load_state_dict(torch.load(os.path.join(md, f"actor_agent{a}.pt"), map_location="cpu"))
        self.critic.critic.load_state_dict(torch.load(os.path.join(md, "critic_agent.pt"), map_location="cpu"))
        vp = os.path.join(md, "value_normalizer.pt")
        if self.value_normalizer is not None and os.path.exists(vp):
            self.value_normalizer.load_state_dict(torch.load(vp, map_location="cpu"))

    def close(self):
        self.envs.close()
        if self.eval_envs is not None and self.eval_envs is not self.envs:
            self.eval_envs.close()
        try:
            self.writter.export_scalars_to_json(os.path.join(self.log_dir, "summary.json"))
        except Exception:
            pass
        self.writter.close()
        self.logger.close()

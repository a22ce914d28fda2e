"""Batched rollout / training driver: DQN over E lockstep envs on one B200.

Replaces the reference's ``Evaluation.run_episodes`` loop (``ttrl_agent/trainer/evaluation.py:139-194``) and the training
half of ``DQNAgent`` (``deep_q_network/abstract.py:37-63, 85-94``, ``deep_q_network/pytorch.py:32-93``) for a vector env:

* acting: the CUDA Q-network kernels (``agent.QNetRollout``; epsilon-greedy on the device), never the torch modules;
* stepping: ``TTRLVectorEnv.step`` (one launch per env-step of all E envs);
* replay memory: a ring buffer of device tensors (``ReplayMemory.push`` / ``sample`` semantics, memory.py:25-52, n_steps = 1);
* learning: double-DQN Bellman residual, loss per ``loss_function_factory``, gradient clamp to [-1, 1], Adam / RMSprop per
  ``optimizer_factory`` (optimizers.py:152-173), target network copied every ``target_update`` optimiser steps -- torch
  autograd on the modules of ``models.py``; with ``torch.distributed`` initialised the gradients are averaged over the
  ranks (NCCL all-reduce: the one data-path collective of the whole system, as SURVEY.md section 8e anticipates);
* checkpoints: ``{"state_dict", "optimizer"}`` like ``DQNAgent.save`` (pytorch.py:82-93), interchangeable with the reference.

Every transition is stored like the reference does (``evaluation.py:176-190``, ``abstract.py:37-63``): ``done`` is the
``terminated`` flag only, so an episode that ends by the time limit is stored with ``terminal=False`` and its REAL next
state -- the vector env restarts finished envs inside the step kernel and hands their terminal observation back as
``info["final_observation"]`` (gymnasium autoreset), which ``record`` uses as ``next_state`` for those rows.

Differences from the reference, all forced by batching and stated here:
* one ``record`` call pushes E (x K agents) transitions and performs ``updates_per_step`` optimiser steps of ``batch_size``
  samples (the reference performs one optimiser step per transition): the update-to-data ratio is
  ``updates_per_step / (E K)`` instead of 1;
* the replay memory holds at least ``min_memory_steps`` vector steps (``max(memory_capacity, min_memory_steps E K)``
  transitions): the reference's 15 000 entries are ~1150 episodes, two vector steps of 8192 envs would overwrite them;
* the exploration clock advances by E per vector step (one tick per collected env-step, like the reference).
"""
from __future__ import annotations

import copy
from typing import Optional

import numpy as np

import ctypes as C

from . import abi
from ._lib import check, lib
from .agent import QNetRollout
from .factory import rec_update
from .models import model_factory, size_model_config


class KernelDQNUpdate:
    """The DQN update as two kernels of the CUDA library (``csrc/ttrl_dqn.cu``: ``ttrl_dqn_grad`` = the three forward passes,
    the double-DQN target, the loss and the backward pass of one minibatch gathered from the device replay memory, in ONE
    launch; ``ttrl_dqn_adam`` = gradient clamp + Adam + the refreshed weight blob of the rollout kernels, in one more) instead of
    ~150 torch / cuBLAS launches.  MultiLayerPerceptron with two hidden layers (baseline.json), ADAM, l2 / l1 / smooth_l1.
    The torch modules and ``torch.optim.Adam``'s state tensors are updated IN PLACE, so checkpoints, the target-network copy
    and evaluation keep working unchanged."""

    LOSSES = {"l2": abi.LOSS_L2, "l1": abi.LOSS_L1, "smooth_l1": abi.LOSS_SMOOTH_L1}

    @classmethod
    def eligible(cls, model_config: dict, config: dict, obs_shape, n_actions: int) -> bool:
        layers = list(model_config.get("layers", [64, 64]))
        return (model_config.get("type") == "MultiLayerPerceptron" and len(layers) == 2 and max(layers) <= 128
                and model_config.get("activation", "RELU") == "RELU" and int(np.prod(obs_shape)) <= 128 and n_actions <= 7
                and config["optimizer"]["type"] == "ADAM" and config["loss_function"] in cls.LOSSES)

    def __init__(self, agent) -> None:
        import torch

        self.torch, self.agent = torch, agent
        self._L = lib()
        net, cfg = agent.value_net, agent.config
        d = abi.DqnDesc()
        d.n_in, d.h1, d.h2 = int(np.prod(agent.obs_shape)), int(net.layers[0].out_features), int(net.layers[1].out_features)
        d.n_actions, d.batch = int(agent.n_actions), int(cfg["batch_size"])
        d.loss, d.double_q, d.gamma = self.LOSSES[cfg["loss_function"]], int(bool(cfg["double"])), float(cfg["gamma"])
        h = C.c_void_p()
        check(self._L.ttrl_dqn_create(C.byref(d), agent.device.index or 0, C.byref(h)))
        self._h = h
        names = ["layers.0.weight", "layers.0.bias", "layers.1.weight", "layers.1.bias", "predict.weight", "predict.bias"]
        self.value_params = [dict(agent.value_net.named_parameters())[n] for n in names]
        self.target_params = [dict(agent.target_net.named_parameters())[n] for n in names]
        n = int(self._L.ttrl_dqn_num_params(self._h))
        assert n == sum(p.numel() for p in self.value_params)
        self.grad = torch.zeros(n, dtype=torch.float32, device=agent.device)
        self.loss = torch.zeros(1, dtype=torch.float32, device=agent.device)
        # torch.optim.Adam creates its state at the first step(): the same tensors, made here, are what the kernel updates
        for p in self.value_params:
            st = agent.optimizer.state[p]
            if "exp_avg" not in st:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(p), torch.zeros_like(p)
        self._bind()

    def _ptrs(self, tensors):
        for t in tensors:
            assert t.is_contiguous() and t.dtype == self.torch.float32
        return (C.c_void_p * 6)(*[t.data_ptr() for t in tensors])

    def _bind(self) -> None:
        """(Re)read the device pointers: after construction and after anything that may have replaced a tensor (load())."""
        state = self.agent.optimizer.state
        self._pv, self._pt = self._ptrs([p.data for p in self.value_params]), self._ptrs([p.data for p in self.target_params])
        self._pm = self._ptrs([state[p]["exp_avg"] for p in self.value_params])
        self._pq = self._ptrs([state[p]["exp_avg_sq"] for p in self.value_params])

    def update(self, idx, refresh_rollout: bool) -> None:
        torch, agent, m = self.torch, self.agent, self.agent.memory
        stream = int(torch.cuda.current_stream(agent.device).cuda_stream)
        check(self._L.ttrl_dqn_grad(self._h, self._pv, self._pt, m.state.data_ptr(), m.next_state.data_ptr(), m.action.data_ptr(),
                                    m.reward.data_ptr(), m.terminal.data_ptr(), idx.data_ptr(), self.grad.data_ptr(), self.loss.data_ptr(), stream))
        scale = 1.0
        if agent._world > 1:  # gradient averaging over the ranks: the one data-path collective (NCCL)
            torch.distributed.all_reduce(self.grad)
            scale = 1.0 / agent._world
        g = agent.optimizer.param_groups[0]
        step = int(agent.optimizer.state[self.value_params[0]]["step"].item()) + 1  # Adam's own counter (restored by load())
        blob = self._L.ttrl_qnet_weights_dev(agent.rollout._h) if refresh_rollout else None
        check(self._L.ttrl_dqn_adam(self._h, self._pv, self._pm, self._pq, self.grad.data_ptr(), step, float(g["lr"]), float(g["betas"][0]),
                                    float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), 1.0, scale, blob, stream))
        for p in self.value_params:  # host-side counters of torch's optimiser state (checkpoint format)
            agent.optimizer.state[p]["step"] += 1

    @property
    def launch_count(self) -> int:
        return int(self._L.ttrl_dqn_launch_count(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.ttrl_dqn_destroy(self._h)
            self._h = None


class DeviceReplayMemory:
    """Ring buffer of transitions on the device: ``push`` appends a batch, ``sample`` draws uniformly without replacement
    (``random.sample``, memory.py:47)."""

    def __init__(self, capacity: int, obs_shape, device, generator) -> None:
        import torch

        self.torch, self.capacity, self.gen = torch, int(capacity), generator
        self.state = torch.zeros((self.capacity,) + tuple(obs_shape), dtype=torch.float32, device=device)
        self.next_state = torch.zeros_like(self.state)
        self.action = torch.zeros(self.capacity, dtype=torch.int64, device=device)
        self.reward = torch.zeros(self.capacity, dtype=torch.float32, device=device)
        self.terminal = torch.zeros(self.capacity, dtype=torch.bool, device=device)
        self.position = 0
        self.size = 0

    def __len__(self) -> int:
        return self.size

    def push(self, state, action, reward, next_state, terminal) -> None:
        torch = self.torch
        n = int(state.shape[0])
        if n == 0:
            return
        if n > self.capacity:  # keep the newest
            state, action, reward, next_state, terminal = (t[-self.capacity:] for t in (state, action, reward, next_state, terminal))
            n = self.capacity
        idx = (self.position + torch.arange(n, device=state.device)) % self.capacity
        self.state[idx], self.next_state[idx] = state, next_state
        self.action[idx], self.reward[idx], self.terminal[idx] = action.to(torch.int64), reward, terminal
        self.position = (self.position + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def sample(self, batch_size: int):
        idx = self.torch.randperm(self.size, device=self.state.device, generator=self.gen)[:batch_size]
        return self.state[idx], self.action[idx], self.reward[idx], self.next_state[idx], self.terminal[idx]


class BatchedDQNAgent:
    """``DQNAgent`` for a vector env.  ``env``: a :class:`TTRLVectorEnv` (or anything with ``num_envs``, ``obs_shape``,
    ``num_agents``, ``single_action_space``, ``device``)."""

    @classmethod
    def default_config(cls) -> dict:  # deep_q_network/abstract.py:21-35
        return dict(model=dict(type="DuelingNetwork"), optimizer=dict(type="ADAM", lr=5e-4, weight_decay=0, k=5),
                    loss_function="l2", memory_capacity=50000, batch_size=100, gamma=0.99, device="cuda:best",
                    exploration=dict(method="EpsilonGreedy"), target_update=1, double=True, n_steps=1)

    def __init__(self, env, config: Optional[dict] = None, seed: int = 0, rollout_mode: str = "fp32",
                 updates_per_step: int = 1, refresh_every: int = 1, cuda_graph: bool = False, min_memory_steps: int = 8,
                 update_kernel="auto") -> None:
        import torch
        from torch.nn import functional as F

        self.torch, self.env = torch, env
        self.config = rec_update(self.default_config(), copy.deepcopy(config or {}))
        if int(self.config.get("n_steps", 1)) != 1:
            raise NotImplementedError("n-step returns are not used by the shipped agent configs (n_steps: 1)")
        self.device = env.device
        self.num_envs, self.K = int(env.num_envs), int(getattr(env, "num_agents", 1))
        self.obs_shape = tuple(env.obs_shape[1:]) if self.K > 1 else tuple(env.obs_shape)
        self.n_actions = int(env.single_action_space.n)
        self.model_config = size_model_config(self.obs_shape, self.n_actions, self.config["model"])
        torch.manual_seed(seed)
        self.value_net = model_factory(self.model_config).to(self.device)
        self.target_net = model_factory(self.model_config).to(self.device)
        self.target_net.load_state_dict(self.value_net.state_dict())
        self.target_net.eval()
        losses = {"l2": F.mse_loss, "l1": F.l1_loss, "smooth_l1": F.smooth_l1_loss, "bce": F.binary_cross_entropy}
        if self.config["loss_function"] not in losses:
            raise ValueError("Unknown loss function : {}".format(self.config["loss_function"]))
        self.loss_function = losses[self.config["loss_function"]]
        opt = self.config["optimizer"]
        # cuda_graph: the whole update (minibatch gather, both forwards, the double-DQN target, backward, gradient clamp, Adam
        # step) is captured once in a CUDA graph and replayed: the update is ~150 small launches and otherwise launch-bound
        dist = torch.distributed
        self._world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        # several ranks: the gradient all-reduce sits between backward and the optimiser step, the update runs eagerly
        self.cuda_graph = bool(cuda_graph) and self._world == 1
        # update_kernel: "auto" = the library's own update kernels when the model / optimiser / loss are ones they implement
        # (KernelDQNUpdate), torch autograd otherwise; True = require them; False = torch autograd
        eligible = KernelDQNUpdate.eligible(self.model_config, self.config, self.obs_shape, self.n_actions)
        if update_kernel is True and not eligible:
            raise NotImplementedError("update_kernel=True: the update kernels implement MultiLayerPerceptron (two hidden layers <= 128 "
                                      "wide, RELU), ADAM and the l2 / l1 / smooth_l1 losses")
        use_kernel = bool(update_kernel) and eligible
        if use_kernel:
            self.cuda_graph = False
        if self.cuda_graph and opt["type"] != "ADAM":
            raise NotImplementedError("cuda_graph=True is implemented for the ADAM optimiser")
        if opt["type"] == "ADAM":
            self.optimizer = torch.optim.Adam(self.value_net.parameters(), lr=opt["lr"], weight_decay=opt["weight_decay"],
                                              capturable=self.cuda_graph)
        elif opt["type"] == "RMS_PROP":
            self.optimizer = torch.optim.RMSprop(self.value_net.parameters(), weight_decay=opt["weight_decay"])
        elif opt["type"] == "RANGER":
            raise NotImplementedError("the RANGER optimiser is not used by the shipped agent configs")
        else:
            raise ValueError("Unknown optimizer type: {}".format(opt["type"]))
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        capacity = max(int(self.config["memory_capacity"]), int(min_memory_steps) * self.num_envs * self.K)
        self.memory = DeviceReplayMemory(capacity, self.obs_shape, self.device, self.gen)
        if self._world > 1:  # every rank starts from rank 0's parameters whatever its seed
            for p in self.value_net.parameters():
                dist.broadcast(p.data, src=0)
            self.target_net.load_state_dict(self.value_net.state_dict())
        # exploration uniforms are keyed by (seed, row, time): offset the key by the shard's first GLOBAL env so that the
        # shards of a multi-GPU run do not draw the same uniforms for the same local row
        first_env = int(getattr(env, "first_env", 0))
        self.rollout = QNetRollout(self.model_config, self.value_net.state_dict(), self.obs_shape, self.n_actions,
                                   device=self.device.index or 0, exploration=self.config["exploration"],
                                   seed=(int(seed) + 0x9E3779B1 * first_env) & 0x7FFFFFFFFFFFFFFF, mode=rollout_mode)
        self.kernel_update = KernelDQNUpdate(self) if use_kernel else None
        self.updates_per_step, self.refresh_every = int(updates_per_step), int(refresh_every)
        self.steps = 0          # optimiser steps (update_target_network's counter, abstract.py:91-94)
        self.training = True
        self.last_loss = None
        self._graph = None

    # ---- acting ---------------------------------------------------------------------------------------------
    def act(self, obs, step_exploration_time: bool = True):
        """obs ``[E, (K,) V, F]`` on the device -> int32 actions ``[E]`` / ``[E, K]`` (abstract.py:65-83: every agent of a
        multi-agent observation acts with the same network; the exploration clock ticks once per env-step)."""
        if step_exploration_time:
            self.rollout.time += self.num_envs
        flat = obs.reshape((-1,) + self.obs_shape)
        a = self.rollout.act(flat, step_exploration_time=False)
        return a.view(self.num_envs, self.K) if self.K > 1 else a

    def eval(self) -> None:
        self.training = False
        self.rollout.eval()

    # ---- learning -------------------------------------------------------------------------------------------
    def record(self, state, action, reward, next_state, terminated, truncated=None, info=None) -> None:
        """``AbstractDQNAgent.record`` (abstract.py:37-63) for a batch of E transitions; ``done`` = ``terminated`` like
        ``Evaluation.step`` passes it (evaluation.py:181-190).  ``info["final_observation"]`` (the vector env's autoreset)
        replaces ``next_state`` for the envs that finished; without it, rows that ended by the time limit only cannot be
        stored (their ``next_state`` is the next episode's first observation) and are dropped."""
        if not self.training:
            return
        torch = self.torch
        K = self.K
        s = state.reshape((self.num_envs, K) + self.obs_shape)
        ns = next_state.reshape((self.num_envs, K) + self.obs_shape)
        a = action.reshape(self.num_envs, K)
        term = terminated.bool()
        if info is not None and "final_observation" in info:
            done = info["_final_observation"].bool().view((self.num_envs,) + (1,) * (ns.dim() - 1))
            ns = torch.where(done, info["final_observation"].reshape(ns.shape), ns)
            r, t = reward, term
        elif truncated is not None:
            keep = ~(truncated.bool() & ~term)
            s, ns, a, r, t = s[keep], ns[keep], a[keep], reward[keep], term[keep]
        else:
            r, t = reward, term
        # the aggregated reward is shared by the agents (abstract.py:53-55)
        self.memory.push(s.reshape((-1,) + self.obs_shape), a.reshape(-1), r.repeat_interleave(K), ns.reshape((-1,) + self.obs_shape),
                         t.repeat_interleave(K))
        ready = len(self.memory) >= self.config["batch_size"]
        if self._world > 1 and not getattr(self, "_ready_everywhere", False):
            # the ranks must agree: the optimiser step holds a collective.  The memory only grows, so once every rank is ready the
            # question (a host synchronisation) is not asked again
            flag = torch.tensor([1 if ready else 0], dtype=torch.int32, device=self.device)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
            ready = bool(flag.item())
            self._ready_everywhere = ready
        for _ in range(self.updates_per_step):
            if not ready:
                return
            if self.kernel_update is not None:
                B = int(self.config["batch_size"])
                idx = torch.randperm(self.memory.size, device=self.device, generator=self.gen)[:B]
                last = _ == self.updates_per_step - 1
                self.kernel_update.update(idx, refresh_rollout=last and (self.steps + 1) % self.refresh_every == 0)
                self.last_loss = self.kernel_update.loss[0]
            elif self.cuda_graph:
                self._graph_update()
            else:
                loss = self.compute_bellman_residual(self.memory.sample(self.config["batch_size"]))
                self.step_optimizer(loss)
                self.last_loss = loss.detach()
            self.update_target_network()
        if self.kernel_update is None and self.steps % self.refresh_every == 0:  # (the kernel path refreshes the blob itself)
            self.rollout.load_parameters(self.value_net, self.model_config)

    def compute_bellman_residual(self, batch):
        """pytorch.py:41-73: double-DQN target from the target network, terminal states do not bootstrap."""
        torch = self.torch
        state, action, reward, next_state, terminal = batch
        q = self.value_net(state).gather(1, action.unsqueeze(1)).squeeze(1)
        with torch.no_grad():
            if self.config["double"]:
                best_actions = self.value_net(next_state).max(1)[1]
                best_values = self.target_net(next_state).gather(1, best_actions.unsqueeze(1)).squeeze(1)
            else:
                best_values = self.target_net(next_state).max(1)[0]
            next_values = torch.where(terminal, torch.zeros_like(reward), best_values)  # static shapes: capturable in a CUDA graph
            target = reward + self.config["gamma"] * next_values
        return self.loss_function(q, target)

    # ---- the update as one CUDA graph ----------------------------------------------------------------------------
    def _update_from_indices(self):
        m, idx = self.memory, self._graph_idx
        loss = self.compute_bellman_residual((m.state[idx], m.action[idx], m.reward[idx], m.next_state[idx], m.terminal[idx]))
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        for p in self.value_net.parameters():
            p.grad.clamp_(-1, 1)
        self.optimizer.step()
        return loss.detach()

    def _build_graph(self) -> None:
        """Capture one update.  The warm-up iterations that torch needs before a capture (lazy optimiser state, cuBLAS
        workspaces) are undone afterwards: parameters and Adam moments are restored in place, so training is unaffected."""
        torch = self.torch
        B = int(self.config["batch_size"])
        self._graph_idx = torch.zeros(B, dtype=torch.int64, device=self.device)
        saved = [p.detach().clone() for p in self.value_net.parameters()]
        # optimiser state as it is NOW (a checkpoint may have been loaded): restored after the warm-up updates below
        saved_opt = {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                     for p, st in self.optimizer.state.items()}
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._update_from_indices()
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self._graph):
            self._graph_loss = self._update_from_indices()
        with torch.no_grad():
            for p, v in zip(self.value_net.parameters(), saved):
                p.copy_(v)
            for p_, state in self.optimizer.state.items():
                old = saved_opt.get(id(p_))
                for k, v in state.items():
                    if torch.is_tensor(v):
                        if old is not None and torch.is_tensor(old.get(k)):
                            v.copy_(old[k])   # in place: the captured graph updates THESE tensors
                        else:
                            v.zero_()

    def _graph_update(self) -> None:
        if self._graph is None:
            self._build_graph()
        B = int(self.config["batch_size"])
        self._graph_idx.copy_(self.torch.randperm(self.memory.size, device=self.device, generator=self.gen)[:B])
        self._graph.replay()
        self.last_loss = self._graph_loss

    def step_optimizer(self, loss) -> None:
        """pytorch.py:32-39 (+ gradient averaging over the ranks when torch.distributed is initialised)."""
        torch = self.torch
        self.optimizer.zero_grad()
        loss.backward()
        dist = torch.distributed
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            grads = [p.grad for p in self.value_net.parameters() if p.grad is not None]
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat)
            flat /= dist.get_world_size()
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        for p in self.value_net.parameters():
            if p.grad is not None:
                p.grad.data.clamp_(-1, 1)
        self.optimizer.step()

    def update_target_network(self) -> None:
        self.steps += 1
        if self.steps % self.config["target_update"] == 0:
            self.target_net.load_state_dict(self.value_net.state_dict())

    # ---- checkpoints (pytorch.py:82-93) -------------------------------------------------------------------------
    def save(self, filename):
        self.torch.save({"state_dict": self.value_net.state_dict(), "optimizer": self.optimizer.state_dict()}, filename)
        return filename

    def load(self, filename):
        checkpoint = self.torch.load(filename, map_location=self.device)
        self.value_net.load_state_dict(checkpoint["state_dict"])
        self.target_net.load_state_dict(checkpoint["state_dict"])
        if self._graph is not None:
            # the captured update reads and writes the optimiser's CURRENT state tensors: copy the checkpoint into them
            # (load_state_dict would rebind the state to new tensors the graph never sees)
            loaded = checkpoint["optimizer"]["state"]
            params = [p for g in self.optimizer.param_groups for p in g["params"]]
            with self.torch.no_grad():
                for idx, p in enumerate(params):
                    src = loaded.get(idx, loaded.get(str(idx)))
                    if src is None:
                        continue
                    for k, v in self.optimizer.state[p].items():
                        if self.torch.is_tensor(v) and k in src:
                            v.copy_(self.torch.as_tensor(src[k]).to(v.device, v.dtype))
        else:
            self.optimizer.load_state_dict(checkpoint["optimizer"])
        self.rollout.load_parameters(self.value_net, self.model_config)
        if self.kernel_update is not None:
            self.kernel_update._bind()
        return filename

    def close(self) -> None:
        if self.kernel_update is not None:
            self.kernel_update.close()
        self.rollout.close()


class BatchedEvaluation:
    """The ``Evaluation.train`` / ``test`` loop (evaluation.py:115-161) for a vector env: no per-episode Python loop -- the
    envs restart on the device and episode statistics are accumulated there (``TTRLVectorEnv.stats``)."""

    def __init__(self, env, agent: BatchedDQNAgent, num_steps: int = 1000, training: bool = True) -> None:
        self.env, self.agent, self.num_steps, self.training = env, agent, int(num_steps), training
        self.observation = None
        self.history = []

    def train(self, log_every: int = 0) -> dict:
        self.training = True
        return self._run(log_every)

    def test(self) -> dict:
        self.training = False
        self.agent.eval()
        return self._run(0)

    def _run(self, log_every: int) -> dict:
        env, agent = self.env, self.agent
        obs, _ = env.reset()
        env.stats(reset=True)
        for step in range(self.num_steps):
            prev = obs.clone()
            actions = agent.act(prev)
            obs, reward, terminated, truncated, info = env.step(actions)
            if self.training:
                agent.record(prev, actions, reward, obs, terminated, truncated, info)
            if log_every and (step + 1) % log_every == 0:
                s = env.stats()
                self.history.append(dict(step=step + 1, episodes=s["episodes"], mean_return=s["total_return"] / max(s["episodes"], 1),
                                         crash_rate=s["crashes"] / max(s["episodes"], 1), epsilon=agent.rollout.epsilon,
                                         loss=None if agent.last_loss is None else float(agent.last_loss)))
        s = env.stats()
        n = max(s["episodes"], 1)
        return dict(episodes=s["episodes"], mean_return=s["total_return"] / n, mean_length=s["total_length"] / n,
                    crash_rate=s["crashes"] / n, arrival_rate=s["arrivals"] / n, env_steps=s["env_steps"])

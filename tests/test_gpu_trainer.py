"""Batched DQN training driver and the JSON loaders on the GPU."""
import json

import numpy as np
import pytest

from topotrafficrl_b200 import factory
from tests import common as T
from tests.test_gpu_parity import _torch
from tests.test_training_host import CONFIGS

pytestmark = pytest.mark.gpu

ENV_JSON = {  # scripts/configs/IntersectionEnv/env.json of the reference ("order": "shuffled")
    "id": "intersection-v0", "import_module": "ttrl_env",
    "observation": {"type": "Kinematics", "vehicles_count": 15, "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
                    "features_range": {"x": [-100, 100], "y": [-100, 100], "vx": [-20, 20], "vy": [-20, 20]},
                    "absolute": True, "order": "shuffled"},
    "destination": "o1"}
AGENT = {"__class__": "<class 'ttrl_agent.agents.deep_q_network.pytorch.DQNAgent'>", "gamma": 0.95, "n_steps": 1, "batch_size": 64,
         "memory_capacity": 15000, "target_update": 16,
         "exploration": {"method": "EpsilonGreedy", "tau": 15000, "temperature": 1.0, "final_temperature": 0.05}}


def test_load_environment_single_env_from_reference_json(tmp_path):
    """load_environment(path) -> a configured, reset single env whose seeded episode is the reference's (shuffled order:
    the env's own numpy stream draws the row permutation, then the spawn)."""
    p = tmp_path / "env.json"
    sorted_json = dict(ENV_JSON, observation={k: v for k, v in ENV_JSON["observation"].items() if k != "order"})
    p.write_text(json.dumps(sorted_json))
    env = factory.load_environment(str(p))
    assert env.import_module == "ttrl_env" and env.config["destination"] == "o1" and env.config["id"] == "intersection-v0"
    assert env.observation_space.shape == (15, 7) and env.action_space.n == 3
    g = T.golden("intersection_steps_kin.npz")  # generated with the default config == env.json with the default (sorted) order
    first, _ = env.reset(seed=100)
    for k in range(3):
        obs, reward, term, trunc, info = env.step(int(g["action"][k]))
        np.testing.assert_allclose(obs, g["obs"][k], rtol=0, atol=1e-5)
        assert abs(reward - g["reward"][k]) <= 1e-6
    env.close()
    # env.json itself ("shuffled"): same seeded reset, the observation rows 1.. are a permutation drawn from the env's stream
    env = factory.load_environment(ENV_JSON)
    shuffled, _ = env.reset(seed=100)
    np.testing.assert_array_equal(shuffled[0], first[0])
    assert not np.array_equal(shuffled, first)
    assert sorted(map(tuple, shuffled[1:].round(6))) == sorted(map(tuple, first[1:].round(6)))
    env.close()


@pytest.mark.parametrize("model", ["mlp", "ego2h", "dueling"])
def test_batched_dqn_training_step_by_step(model, tmp_path):
    torch = _torch()
    env = factory.load_environment(ENV_JSON, num_envs=256, seed=5)
    agent = factory.load_agent(dict(AGENT, model=CONFIGS[model]), env, seed=3)
    assert agent.obs_shape == (15, 7) and agent.n_actions == 3 and agent.rollout.epsilon == 1.0
    obs, _ = env.reset()
    before = {k: v.clone() for k, v in agent.value_net.state_dict().items()}
    for step in range(24):
        prev = obs.clone()
        actions = agent.act(prev)
        assert actions.shape == (256,) and actions.dtype == torch.int32 and int(actions.min()) >= 0 and int(actions.max()) <= 2
        obs, reward, term, trunc, _ = env.step(actions)
        agent.record(prev, actions, reward, obs, term, trunc)
    assert agent.steps == 24 and torch.isfinite(agent.last_loss)
    assert agent.rollout.time == 24 * 256 and agent.rollout.epsilon < 1.0
    assert len(agent.memory) > 20 * 256
    changed = [k for k, v in agent.value_net.state_dict().items() if not torch.equal(v, before[k])]
    assert len(changed) == len(before)
    # the target network lags: copied at optimiser steps 16 (target_update), not since
    assert any(not torch.equal(a, b) for a, b in zip(agent.target_net.state_dict().values(), agent.value_net.state_dict().values()))
    # the rollout kernels carry the CURRENT value network: their Q-values == the torch forward of value_net
    agent.eval()
    a, q = agent.rollout.act(obs.reshape(-1, 15, 7), step_exploration_time=False, return_q=True)
    with torch.no_grad():
        want = agent.value_net(obs.reshape(-1, 15, 7))
    np.testing.assert_allclose(q.cpu().numpy(), want.cpu().numpy(), rtol=0, atol=5e-5)
    gap = want.sort(dim=1, descending=True)[0]
    clear = (gap[:, 0] - gap[:, 1]) > 1e-4
    assert (a.long()[clear] == want.argmax(1)[clear]).all()
    # checkpoint round trip in the reference's format
    path = agent.save(str(tmp_path / "checkpoint.tar"))
    ck = torch.load(path, map_location="cpu")
    assert set(ck) == {"state_dict", "optimizer"} and set(ck["state_dict"]) == set(before)
    other = factory.load_agent(dict(AGENT, model=CONFIGS[model]), env, seed=99)
    other.load(path)
    _, q2 = other.rollout.act(obs.reshape(-1, 15, 7), step_exploration_time=False, return_q=True)
    np.testing.assert_allclose(q2.cpu().numpy(), q.cpu().numpy(), rtol=0, atol=1e-6)
    agent.close(); other.close(); env.close()


def test_batched_evaluation_multi_agent_and_learning_signal():
    """BatchedEvaluation over the multi-agent intersection (K = 2 agents share one network, abstract.py:53-55) trains, logs
    and then tests greedily; training lowers the Bellman loss on a fixed probe batch."""
    torch = _torch()
    from topotrafficrl_b200.trainer import BatchedEvaluation
    env = factory.load_environment({"id": "intersection-multi-agent-v0", "import_module": "ttrl_env",
                                    "observation": T.MULTI_AGENT["observation"], "controlled_vehicles": 2}, num_envs=512, seed=1)
    assert env.num_agents == 2 and env.obs_shape == (2, 15, 7)
    agent = factory.load_agent(dict(AGENT, model=CONFIGS["mlp"], batch_size=256), env, seed=0, updates_per_step=4)
    ev = BatchedEvaluation(env, agent, num_steps=60)
    out = ev.train(log_every=20)
    assert out["env_steps"] == 60 * 512 and out["episodes"] > 512 and len(ev.history) == 3
    assert agent.steps == 60 * 4 - 0 or agent.steps > 200
    # learning signal: with the target network frozen, optimiser steps on one fixed batch of collected transitions lower its loss
    probe = agent.memory.sample(2048)
    fresh = factory.load_agent(dict(AGENT, model=CONFIGS["mlp"], batch_size=256, target_update=10 ** 9), env, seed=7)
    losses = []
    for _ in range(60):
        loss = fresh.compute_bellman_residual(probe)
        losses.append(float(loss.detach()))
        fresh.step_optimizer(loss)
        fresh.update_target_network()
    assert losses[-1] < 0.8 * losses[0]  # lr 5e-4, 60 steps: 1.45 -> 0.91 on B200
    res = BatchedEvaluation(env, agent, num_steps=15).test()
    assert agent.rollout.epsilon == 0.0 and res["episodes"] > 0 and np.isfinite(res["mean_return"])
    agent.close(); fresh.close(); env.close()


def test_cuda_graph_update_equals_eager_update():
    """The update captured in a CUDA graph (gather, forwards, target, backward, clamp, Adam) == the eager update, step by step,
    up to float32 rounding (measured on B200: 2e-8 after 2 updates, 1e-7 after 4; Adam's g / (sqrt(v) + eps) amplifies
    last-bit gradient differences of parameters whose gradients are ~eps, so the bound is loose and the median is tight)."""
    torch = _torch()
    agents, envs = [], []
    for graph in (False, True):
        env = factory.load_environment(ENV_JSON, num_envs=128, seed=9)
        envs.append(env)
        agents.append(factory.load_agent(dict(AGENT, model=CONFIGS["mlp"]), env, seed=4, cuda_graph=graph, update_kernel=False))
    obs = [env.reset()[0] for env in envs]
    for step in range(8):
        for k, (env, agent) in enumerate(zip(envs, agents)):
            prev = obs[k].clone()
            a = agent.act(prev)
            obs[k], reward, term, trunc, _ = env.step(a)
            agent.record(prev, a, reward, obs[k], term, trunc)
        assert torch.equal(obs[0], obs[1]), f"rollouts diverged at step {step}"
    assert agents[1]._graph is not None and agents[0].steps == agents[1].steps == 8
    for (n, p), q in zip(agents[0].value_net.named_parameters(), agents[1].value_net.parameters()):
        d = (p.detach() - q.detach()).abs()
        assert float(d.max()) < 5e-5 and float(d.median()) < 2e-7, (n, float(d.max()), float(d.median()))
    assert abs(float(agents[0].last_loss) - float(agents[1].last_loss)) < 1e-4
    for a in agents:
        a.close()
    for e in envs:
        e.close()


@pytest.mark.parametrize("loss,double", [("l2", True), ("smooth_l1", True), ("l1", False)])
def test_kernel_update_equals_torch_autograd_update(loss, double):
    """The library's DQN update kernels (ttrl_dqn_grad + ttrl_dqn_adam: forwards, double-DQN target, loss, backward, clamp, Adam,
    rollout blob) against the torch autograd update (pytorch.py:32-73) on the same replay memory and the same minibatches:
    gradients of one minibatch to 1e-6, parameters and loss curve over 40 updates to float32 rounding, rollout weights refreshed."""
    torch = _torch()
    from topotrafficrl_b200.trainer import KernelDQNUpdate
    cfg = dict(AGENT, model=CONFIGS["mlp"], loss_function=loss, double=double, batch_size=96, target_update=8)  # 96 rows: a full + a ragged pass
    env = factory.load_environment(ENV_JSON, num_envs=128, seed=9)
    envs = [env]
    agents = [factory.load_agent(cfg, env, seed=4, update_kernel=kernel) for kernel in (False, True)]
    assert agents[0].kernel_update is None and isinstance(agents[1].kernel_update, KernelDQNUpdate)
    obs, _ = env.reset()
    losses = [[], []]
    for step in range(40):
        prev = obs.clone()
        a = agents[0].act(prev)  # ONE stream of transitions feeds both learners (same replay memory, same minibatch indices)
        obs, reward, term, trunc, info = env.step(a)
        for k, agent in enumerate(agents):
            agent.record(prev, a, reward, obs, term, trunc, info)
            losses[k].append(float(agent.last_loss))
        if step == 0:  # one minibatch: the kernel's flat gradient against autograd's, before anything can drift
            ref, ker = agents
            idx = torch.arange(96, device="cuda")
            ker.kernel_update._L.ttrl_dqn_grad(ker.kernel_update._h, ker.kernel_update._pv, ker.kernel_update._pt, ker.memory.state.data_ptr(),
                                               ker.memory.next_state.data_ptr(), ker.memory.action.data_ptr(), ker.memory.reward.data_ptr(),
                                               ker.memory.terminal.data_ptr(), idx.data_ptr(), ker.kernel_update.grad.data_ptr(),
                                               ker.kernel_update.loss.data_ptr(), 0)
            m = ker.memory
            l = ker.compute_bellman_residual((m.state[idx], m.action[idx], m.reward[idx], m.next_state[idx], m.terminal[idx]))
            ker.value_net.zero_grad()
            l.backward()
            want = torch.cat([p.grad.reshape(-1) for p in ker.kernel_update.value_params])
            np.testing.assert_allclose(ker.kernel_update.grad.cpu().numpy(), want.cpu().numpy(), rtol=0, atol=2e-6)
            assert abs(float(ker.kernel_update.loss[0]) - float(l)) < 1e-5
            ker.value_net.zero_grad()
        assert torch.equal(agents[0].memory.state, agents[1].memory.state)
    assert agents[0].steps == agents[1].steps == 40
    for (n, p), q in zip(agents[0].value_net.named_parameters(), agents[1].value_net.parameters()):
        d = (p.detach() - q.detach()).abs()
        assert float(d.max()) < 2e-3 and float(d.median()) < 2e-5, (n, float(d.max()), float(d.median()))
    np.testing.assert_allclose(losses[1][:8], losses[0][:8], rtol=1e-4, atol=1e-6)
    # Adam state in torch's own format (checkpoints), target network copied at multiples of target_update
    st = agents[1].optimizer.state[agents[1].kernel_update.value_params[0]]
    assert float(st["step"]) == 40 and float(st["exp_avg_sq"].abs().sum()) > 0
    assert all(torch.equal(a, b) for a, b in zip(agents[1].target_net.state_dict().values(), agents[1].value_net.state_dict().values()))
    # the rollout kernels carry the updated network (blob written by ttrl_dqn_adam)
    agents[1].eval()
    _, q = agents[1].rollout.act(obs.reshape(-1, 15, 7), step_exploration_time=False, return_q=True)
    with torch.no_grad():
        want = agents[1].value_net(obs.reshape(-1, 15, 7))
    np.testing.assert_allclose(q.cpu().numpy(), want.cpu().numpy(), rtol=0, atol=5e-5)
    for a in agents:
        a.close()
    for e in envs:
        e.close()


@pytest.mark.parametrize("path", ["graph", "kernel"])
def test_checkpoint_resume_keeps_optimizer_state(path, tmp_path):
    """save() -> load() into a fresh agent -> one more update == continuing the original agent: the Adam moments and step counter
    survive the resume on the CUDA-graph path (whose warm-up updates before the capture must not wipe a loaded state, and whose
    captured graph must see a state loaded after the capture) and on the kernel path (which updates torch's state tensors in place)."""
    torch = _torch()
    kw = dict(cuda_graph=True, update_kernel=False) if path == "graph" else dict(update_kernel=True)
    env = factory.load_environment(ENV_JSON, num_envs=128, seed=9)
    cfg = dict(AGENT, model=CONFIGS["mlp"], batch_size=64, target_update=4)
    a = factory.load_agent(cfg, env, seed=4, **kw)
    obs, _ = env.reset()
    for step in range(12):
        prev = obs.clone()
        act = a.act(prev)
        obs, reward, term, trunc, info = env.step(act)
        a.record(prev, act, reward, obs, term, trunc, info)
    ck = a.save(str(tmp_path / "ck.tar"))
    # resume BEFORE the first update of the new agent (graph path: the capture happens after the load) ...
    b = factory.load_agent(cfg, env, seed=99, **kw)
    b.load(ck)
    # ... and AFTER it (graph path: the capture already happened)
    c = factory.load_agent(cfg, env, seed=98, **kw)
    for name in ("state", "next_state", "action", "reward", "terminal"):
        for other in (b, c):
            getattr(other.memory, name).copy_(getattr(a.memory, name))
    for other in (b, c):
        other.memory.size, other.memory.position = a.memory.size, a.memory.position
    prev = obs.clone()
    act = a.act(prev)
    obs2, reward, term, trunc, info = env.step(act)
    c.gen.set_state(a.gen.get_state())
    c.record(prev, act, reward, obs2, term, trunc, info)   # one update with fresh moments, then the load
    c.load(ck)
    for name in ("state", "next_state", "action", "reward", "terminal"):
        getattr(c.memory, name).copy_(getattr(a.memory, name))
    c.memory.size, c.memory.position = a.memory.size, a.memory.position
    state = a.gen.get_state()
    for other in (b, c):
        other.gen.set_state(state)
    a.gen.set_state(state)
    for agent in (a, b, c):
        agent.record(prev, act, reward, obs2, term, trunc, info)
    for (n, p), q, r in zip(a.value_net.named_parameters(), b.value_net.parameters(), c.value_net.parameters()):
        assert float((p - q).abs().max()) < 1e-6, (path, "load before the first update", n)
        assert float((p - r).abs().max()) < 1e-6, (path, "load after the first update", n)
    pa = next(iter(a.value_net.parameters()))
    sa, sb = a.optimizer.state[pa], b.optimizer.state[next(iter(b.value_net.parameters()))]
    assert float(sa["step"]) == float(sb["step"]) == 13
    for agent in (a, b, c):
        agent.close()
    env.close()

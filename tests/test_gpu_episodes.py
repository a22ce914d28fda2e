"""GPU forms of tests/test_episode_goldens.py: the reference-shipped IntersectionEnv configs as whole seeded episodes
(env.json "shuffled", env_5fps.json, normalize_reward / destination None, o2 / o3) through the C ABI, the single-env front
end and the vector env (gymnasium final_observation, the reference's (s, a, r, s', done) stream incl. the truncated last
step, batched info), slot-capacity rejects, and a full-size BASELINE configs[3] check (8192 envs + Q-net in the loop)."""
import ctypes as C

import numpy as np
import pytest

from topotrafficrl_b200 import abi, scenes
from tests import common as T
from tests.test_episode_goldens import (check_info, check_step_outputs, compare_episode_states, episode_arrays, inverse_perm)
from tests.test_gpu_parity import _sim, _torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(T.EPISODE_CONFIGS))
def test_c_abi_step_follows_reference_episodes(name):
    """Every step of every golden episode resynced to the reference's state: ttrl_sim_step_host with the reference's spawn
    draws (ttrl_sim_inject_spawn) and row permutation (ttrl_sim_inject_shuffle) == state, obs, reward, flags, info."""
    over, _ = T.EPISODE_CONFIGS[name]
    g, st, want = episode_arrays(name)
    _, table, cfg, routes = T.intersection_scene(over)
    sim = _sim(cfg, table, st.num_envs, 32, routes)
    sim.set_state(st)
    sim.inject_spawn(T.draws_array(g["draw"]))
    sim.inject_shuffle(inverse_perm(g["perm"]))
    sim.host_info(copy=False)
    obs, reward, term, trunc = sim.step_host(g["action"].astype(np.int32))
    compare_episode_states(sim.get_state(), want, name)
    check_step_outputs(g, obs, reward, term, trunc)
    info, _ = sim.host_info()
    ar, at = sim.agent_outputs_host()
    check_info(g, info, ar, at)
    sim.close()


@pytest.mark.parametrize("name", sorted(T.EPISODE_CONFIGS))
def test_single_env_front_end_replays_reference_episodes(name):
    """IntersectionEnv(config).reset(seed) + step(a), free running over whole episodes: the env's own numpy stream draws the
    row permutations and the spawns in the reference's order; obs / reward / flags / info == the reference's, step by step."""
    from topotrafficrl_b200.envs import IntersectionEnv
    over, seeds = T.EPISODE_CONFIGS[name]
    g = T.golden("intersection_ep_%s.npz" % name)
    env = IntersectionEnv(config=over)
    n = int(g["n_steps"][0])
    keys = abi.REWARD_KEYS[abi.REWARD_INTERSECTION]
    # free running (NOT resynced: the resynced form above holds 1e-6 per step).  An ego that brakes to a near standstill makes
    # the reference's own steering law ill-conditioned (gain ~ 1 / not_zero(speed)^2, controller.py:170-186), which turns
    # last-ulp differences into 1e-5 .. 1e-4 of heading within an episode
    tol = 2e-4
    for ep, seed in enumerate(seeds):
        obs, info = env.reset(seed=seed)
        np.testing.assert_allclose(obs, g["reset_obs"][ep], rtol=0, atol=2e-6, err_msg=f"{name} reset {seed}")
        assert abs(info["speed"] - g["reset_info"][ep, 0]) <= 1e-9 and info["crashed"] == bool(g["reset_info"][ep, 1])
        np.testing.assert_allclose([info["rewards"][k] for k in keys], g["reset_info"][ep, 2:6], rtol=0, atol=1e-9)
        k = int(g["ep_first"][ep])
        last = int(g["ep_first"][ep + 1]) if ep + 1 < len(seeds) else n
        done = False
        while not done:
            obs, reward, term, trunc, info = env.step(int(g["action"][k]))
            np.testing.assert_allclose(obs, g["obs"][k], rtol=0, atol=tol, err_msg=f"{name} seed {seed} step {k}")
            assert abs(reward - g["reward"][k]) <= 10 * tol
            assert term == bool(g["terminated"][k]) and trunc == bool(g["truncated"][k])
            assert info["crashed"] == bool(g["info"][k, 1]) and abs(info["speed"] - g["info"][k, 0]) <= 10 * tol
            np.testing.assert_allclose([info["rewards"][q] for q in keys], g["info"][k, 2:6], rtol=0, atol=10 * tol)
            np.testing.assert_allclose(info["agents_rewards"], g["agents_rewards"][k], rtol=0, atol=10 * tol)
            assert tuple(info["agents_terminated"]) == tuple(bool(x) for x in g["agents_terminated"][k])
            done = term or trunc
            k += 1
        assert k == last, "episode length differs from the reference's"
    env.close()


def test_vector_env_transition_stream_and_final_observation():
    """The vector env replays the 12 reference episodes of env.json side by side (host-driven reset = the reference's
    reset(seed + e); the reference's spawn draws and row permutations injected): the (s, a, r, s', done) stream of every env
    equals the reference's up to and including its LAST step -- where the env restarts inside the step kernel, the returned
    observation already belongs to the next episode and s' is info["final_observation"] -- and the replay memory of the
    training driver holds exactly those transitions (time-limit ends with terminal=False and the real next state)."""
    torch = _torch()
    from topotrafficrl_b200.trainer import BatchedDQNAgent
    from topotrafficrl_b200.vector_env import TTRLVectorEnv
    over, seeds = T.EPISODE_CONFIGS["envjson"]
    g = T.golden("intersection_ep_envjson.npz")
    E, n = len(seeds), int(g["n_steps"][0])
    first = g["ep_first"].astype(int)
    length = np.diff(np.append(first, n))
    env = TTRLVectorEnv(E, scene="intersection", config=over, seed=int(seeds[0]), reset_mode="host", vcap=32)
    obs, _ = env.reset()
    T.compare_states(env.get_state(), T.batch_state(g, "reset"), 1e-9, "host reset == reference reset(seed + e)")
    agent = BatchedDQNAgent(env, {"model": {"type": "MultiLayerPerceptron", "layers": [32, 32]}, "batch_size": 4096, "memory_capacity": 1000},
                            seed=0, min_memory_steps=int(length.max()) + 1)
    truncated_rows = 0
    tol = 2e-4  # free-running episodes: see test_single_env_front_end_replays_reference_episodes
    for k in range(int(length.max())):
        running = k < length
        idx = np.where(running, first + k, first)          # finished envs: any valid row (not compared)
        env.sim.inject_spawn(T.draws_array(g["draw"][idx]))
        env.sim.inject_shuffle(inverse_perm(g["perm"][idx]))
        prev = obs.clone()
        actions = torch.as_tensor(g["action"][idx].astype(np.int32), device="cuda")
        obs, reward, term, trunc, info = env.step(actions)
        agent.record(prev, actions, reward, obs, term, trunc, info)
        o, r, t, u = obs.cpu().numpy(), reward.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
        fo, fmask = info["final_observation"].cpu().numpy(), info["_final_observation"].cpu().numpy()
        for e in np.nonzero(running)[0]:
            j = first[e] + k
            assert t[e] == g["terminated"][j] and u[e] == g["truncated"][j] and abs(r[e] - g["reward"][j]) <= 1e-5
            assert fmask[e] == (g["terminated"][j] or g["truncated"][j])
            next_state = fo[e] if fmask[e] else o[e]
            np.testing.assert_allclose(next_state, g["obs"][j], rtol=0, atol=tol, err_msg=f"env {e} step {k}")
            if fmask[e]:
                assert k + 1 == length[e]
                assert not np.allclose(o[e], g["obs"][j], atol=1e-3)  # the returned row is the next episode's first observation
                truncated_rows += int(g["truncated"][j] and not g["terminated"][j])
            # the transition the training driver stored for this env at this step
            slot = k * E + e
            np.testing.assert_allclose(agent.memory.next_state[slot].cpu().numpy(), g["obs"][j], rtol=0, atol=tol)
            assert bool(agent.memory.terminal[slot]) == bool(g["terminated"][j])
            assert int(agent.memory.action[slot]) == int(g["action"][j]) and abs(float(agent.memory.reward[slot]) - g["reward"][j]) <= 1e-5
            if k > 0:
                np.testing.assert_allclose(agent.memory.state[slot].cpu().numpy(), g["obs"][j - 1], rtol=0, atol=tol)
        crashed = info["crashed"].cpu().numpy()
        for e in np.nonzero(running)[0]:
            assert crashed[e] == bool(g["info"][first[e] + k, 1])
    assert truncated_rows > 0 and len(agent.memory) == int(length.max()) * E
    agent.close()
    env.close()


def test_spawn_capacity_rejects_are_counted():
    """A slot table too small for the traffic rejects spawns the reference would accept: counted, never silent."""
    torch = _torch()
    from topotrafficrl_b200.vector_env import TTRLVectorEnv
    rng = np.random.default_rng(0)
    rejects = {}
    for vcap in (12, 24):
        env = TTRLVectorEnv(512, scene="intersection", seed=9, vcap=vcap, config={"spawn_probability": 1.0, "duration": 40})
        env.reset()
        for _ in range(30):
            env.step(torch.as_tensor(np.zeros(512, np.int32), device="cuda"))  # SLOWER: the ego waits, traffic piles up
        s = env.stats()
        rejects[vcap] = s["spawn_capacity_rejects"]
        assert (env.get_state().env_i[abi.EI_NVEH] <= vcap).all()
        env.close()
    assert rejects[12] > 0 and rejects[24] <= rejects[12]


def test_full_size_configs3_vs_oracle_with_qnet_in_the_loop():
    """BASELINE configs[3] at full size: 8192 intersection envs (device reset, device Philox spawn draws, regulated road) with
    the DQN Q-network rollout in the loop, two env-steps: state / obs / reward / flags against the oracle fed the device's own
    draws, Q-values and greedy actions against a plain torch fp32 forward of the same weights."""
    from oracle import oracle as O
    from tests.emu.emu import lib as emu_lib
    from topotrafficrl_b200.agent import QNetRollout
    from topotrafficrl_b200.models import model_factory, size_model_config
    from topotrafficrl_b200.vector_env import TTRLVectorEnv
    torch = _torch()
    E = 8192
    env = TTRLVectorEnv(E, scene="intersection", seed=21, async_reset=False, autoreset=False)
    obs, _ = env.reset()
    from tests.test_training_host import CONFIGS
    mc = size_model_config((15, 7), 3, CONFIGS["ego2h"])  # scripts/configs/IntersectionEnv/agents/DQNAgent/ego_attention_2h.json
    torch.manual_seed(5)
    net = model_factory(mc).to("cuda").eval()
    roll = QNetRollout(mc, net.state_dict(), (15, 7), 3, mode="fp32")
    roll.eval()
    orc = O.Oracle(env.cfg, env.table, scenes.intersection_spawn_routes(env.net, env.table), threads=8)
    L = emu_lib()
    seed = (env.seed_value << 1) | 1
    for k in range(2):
        actions, q = roll.act(obs, return_q=True)
        with torch.no_grad():
            q_ref = net(obs)
        np.testing.assert_allclose(q.cpu().numpy(), q_ref.cpu().numpy(), rtol=0, atol=2e-5)
        top2 = q_ref.topk(2, dim=1).values
        clear = (top2[:, 0] - top2[:, 1]) > 1e-4
        assert (actions.long()[clear] == q_ref.argmax(1)[clear]).all() and clear.float().mean() > 0.99
        ref = env.get_state()
        draws = (abi.SpawnDraw * E)()
        for e in range(E):
            # the spawn of a step is keyed by env.steps AFTER its 15 sub-steps (ttrl_core.cuh: env_step_finish)
            counter = (int(ref.env_i[abi.EI_STEPS, e]) + 15) | (int(ref.env_i[abi.EI_EPISODE, e]) << 32)
            L.emu_device_spawn_draw(C.c_uint64(seed), C.c_int64(e), C.c_uint64(counter), C.byref(draws[e]))
        a_host = actions.cpu().numpy().astype(np.int32)
        obs, reward, term, trunc, info = env.step(actions)
        oo, orr, ot, ou, _ = orc.step(ref, a_host, draws)
        # envs holding a crashed pair are chaotic at the ulp level from the impact on
        calm = ((ref.veh_i[abi.I_FLAGS] & (abi.FL_CRASHED | abi.FL_HAS_IMPACT)) == 0).all(axis=1)
        got = env.get_state()
        sel = np.nonzero(calm)[0]
        assert sel.size > 0.9 * E
        T.compare_states(got.select_envs(sel), ref.select_envs(sel), 1e-6, f"configs[3] step {k}", check_action=False)
        np.testing.assert_allclose(obs.cpu().numpy().reshape(E, -1)[sel], oo[sel], rtol=0, atol=2e-5)
        np.testing.assert_allclose(reward.cpu().numpy()[sel], orr[sel], rtol=0, atol=1e-5)
        assert (term.cpu().numpy()[sel] == ot[sel].astype(bool)).all() and (trunc.cpu().numpy()[sel] == ou[sel].astype(bool)).all()
    assert env.stats()["spawn_capacity_rejects"] == 0
    roll.close()
    env.close()

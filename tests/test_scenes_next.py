"""SURVEY.md section 8 rows A29 (K controlled vehicles) and N3 (RoundaboutEnv, UTurnEnv): the CPU oracle and the emulated
device logic against golden vectors of the unmodified reference (tests/golden/make_golden.py multi / scripted), and the
host-driven resets against the reference's seeded resets.  CPU only; the CUDA kernels run the same comparisons in
tests/test_gpu_scenes.py."""
import numpy as np
import pytest

from oracle import oracle as O
from topotrafficrl_b200 import abi, scenes
from topotrafficrl_b200._gym import np_random
from topotrafficrl_b200.reset import reset_intersection, reset_roundabout, reset_uturn
from topotrafficrl_b200.state import SimState, unpack_route
from tests import common as T
from tests.emu.emu import Emulator
from tests.test_host_logic import _EmuBackend


def _engines(cfg, table, routes=None):
    return [("oracle", O.Oracle(cfg, table, routes)), ("emulated device logic", Emulator(cfg, table, routes))]


def test_multi_agent_step_vs_reference():
    """MultiAgentIntersectionEnv.step with tuple actions: state, K observations, mean reward, flags, per-agent info."""
    g = T.golden("multiagent_steps.npz")
    _, table, cfg, _, routes = T.multi_agent_scene()
    assert cfg.controlled_vehicles == 4
    for what, eng in _engines(cfg, table, routes):
        st = T.batch_state(g, "before")
        obs, reward, term, trunc, _ = eng.step(st, g["action"].astype(np.int32), T.draws_array(g["draw"]))
        T.compare_states(st, T.batch_state(g, "after"), T.TOL_STEP, what)
        np.testing.assert_allclose(obs.reshape(g["obs"].shape), g["obs"], rtol=0, atol=2e-6, err_msg=what)
        np.testing.assert_allclose(reward, g["reward"], rtol=0, atol=1e-6)
        assert (term.astype(bool) == g["terminated"]).all() and (trunc.astype(bool) == g["truncated"]).all()
        np.testing.assert_allclose(eng.agent_reward, g["agents_rewards"], rtol=0, atol=1e-6)
        assert (eng.agent_terminated.astype(bool) == g["agents_terminated"]).all()
    # the fixture exercises what the row is about: all four agents act, and routes longer than one 4-entry word exist
    assert len({tuple(a) for a in g["action"]}) > 4
    assert g["before_vi"][:, abi.I_ROUTE_LEN].max() > 4


def test_multi_agent_reset_vs_reference():
    """Host-driven reset with K = 4 controlled vehicles == MultiAgentIntersectionEnv.reset(seed) of the reference."""
    g = T.golden("multiagent_steps.npz")
    net, table, cfg, cfgd, routes = T.multi_agent_scene()
    emu = Emulator(cfg, table, routes)
    seeds = [int(s) for s in g["reset_seed"]]
    backend = _EmuBackend(emu, len(seeds), 32)
    st = reset_intersection(backend, [np_random(s)[0] for s in seeds], net, table, cfgd, cfg)
    T.compare_states(st, T.batch_state(g, "reset"), 1e-9, "multi-agent reset")
    obs = emu.observe(backend.st)
    np.testing.assert_allclose(obs.reshape(g["reset_obs"].shape), g["reset_obs"], rtol=0, atol=2e-6)
    # ego 1 starts on arm 1 with destination o1: the BFS route goes around (6 entries, > one route word)
    e1 = [s for s in range(32) if (st.veh_i[abi.I_FLAGS, 0, s] & abi.FL_AGENT_MASK) >> abi.FL_AGENT_SHIFT == 1][0]
    route = unpack_route(st.veh_i[abi.I_ROUTE_LEN, 0, e1], [st.veh_i[w, 0, e1] for w in abi.I_ROUTE_ROAD_WORDS],
                         [st.veh_i[w, 0, e1] for w in abi.I_ROUTE_LANE_WORDS])
    assert len(route) == 6 and table.road_keys[route[-1][0]] == ("il1", "o1")


def test_multi_agent_device_reset_follows_make_vehicles():
    """Device-side reset with K egos: one MDP ego per arm, agent bits in list order, routes to the configured exit."""
    _, table, cfg, cfgd, routes = T.multi_agent_scene()
    emu = Emulator(cfg, table, routes)
    emu.set_reset_params(scenes.intersection_reset_params(cfgd))
    st = SimState.zeros(8, 32)
    emu.reset(st, 9, 0, 0)
    for e in range(8):
        n = int(st.env_i[abi.EI_NVEH, e])
        fl = st.veh_i[abi.I_FLAGS, e, :n]
        ctl = [s for s in range(n) if fl[s] & abi.FL_CONTROLLED]
        assert [(int(fl[s]) & abi.FL_AGENT_MASK) >> abi.FL_AGENT_SHIFT for s in ctl] == [0, 1, 2, 3]
        assert st.env_i[abi.EI_EGO, e] == ctl[0]
        for k, s in enumerate(ctl):
            assert table.lane_keys[st.veh_i[abi.I_LANE, e, s]] == (f"o{k}", f"ir{k}", 0)
            route = unpack_route(st.veh_i[abi.I_ROUTE_LEN, e, s], [st.veh_i[w, e, s] for w in abi.I_ROUTE_ROAD_WORDS],
                                 [st.veh_i[w, e, s] for w in abi.I_ROUTE_LANE_WORDS])
            assert table.road_keys[route[-1][0]][1] == "o1"
    orc = O.Oracle(cfg, table, routes)
    a, b = st.copy(), st.copy()
    act = np.random.default_rng(0).integers(0, 3, size=(8, 4)).astype(np.int32)
    oa, ob = emu.step(a, act), orc.step(b, act)
    T.compare_states(a, b, 1e-7, "free-running multi-agent step")
    np.testing.assert_allclose(oa[0], ob[0], rtol=0, atol=2e-6)
    np.testing.assert_allclose(emu.agent_reward, orc.agent_reward, rtol=0, atol=1e-6)


@pytest.mark.parametrize("scene", ["roundabout", "uturn"])
def test_scripted_scene_substeps_and_steps_vs_reference(scene):
    _, table, cfg, _ = T.roundabout_scene() if scene == "roundabout" else T.uturn_scene()
    gs, g = T.golden(f"{scene}_substeps.npz"), T.golden(f"{scene}_steps.npz")
    for what, eng in _engines(cfg, table):
        st = T.batch_state(gs, "before")
        eng.substep(st, gs["action"].astype(np.int32))
        T.compare_states(st, T.batch_state(gs, "after"), T.TOL_SUBSTEP, f"{scene} sub-step, {what}")
        st = T.batch_state(g, "before")
        obs, reward, term, trunc, _ = eng.step(st, g["action"].astype(np.int32))
        T.compare_states(st, T.batch_state(g, "after"), T.TOL_STEP, f"{scene} step, {what}")
        np.testing.assert_allclose(obs.reshape(g["obs"].shape), g["obs"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(reward, g["reward"], rtol=0, atol=1e-6)
        assert (term.astype(bool) == g["terminated"]).all() and (trunc.astype(bool) == g["truncated"]).all()
    assert len(set(g["reward"].round(6))) > 3  # the reward terms vary over the fixture


@pytest.mark.parametrize("scene", ["roundabout", "uturn"])
def test_scripted_scene_reset_vs_reference(scene):
    """reset_roundabout / reset_uturn with gymnasium-seeded Generators == RoundaboutEnv / UTurnEnv .reset(seed)."""
    net, table, cfg, cfgd = T.roundabout_scene() if scene == "roundabout" else T.uturn_scene()
    g = T.golden(f"{scene}_steps.npz")
    rngs = [np_random(int(s))[0] for s in g["reset_seed"]]
    st = (reset_roundabout if scene == "roundabout" else reset_uturn)(rngs, net, table, cfgd, cfg, 16)
    T.compare_states(st, T.batch_state(g, "reset"), 1e-12, f"{scene} reset")
    obs = Emulator(cfg, table).observe(st)
    np.testing.assert_allclose(obs.reshape(g["reset_obs"].shape), g["reset_obs"], rtol=0, atol=2e-6)
    if scene == "roundabout":
        assert g["reset_vi"][:, abi.I_ROUTE_LEN].max() > 4  # routes beyond one 4-entry word


class _CastDrawRng:
    """Feeds ``reset_roundabout`` / ``reset_uturn`` (validated against the reference's resets above) with the draws the
    DEVICE-side scripted reset makes for one env (ttrl_core.cuh: reset_cast), member by member."""

    def __init__(self, emu, seed, env, episode):
        self.emu, self.seed, self.env, self.episode = emu, seed, env, episode
        self.member, self.calls = 0, 0

    def _u(self, which):
        return self.emu.reset_uniforms(self.seed, self.env, self.episode, 0x200 + which + 2 * self.member)

    def normal(self):
        if self.calls % 2 == 0:
            self.member += 1  # a new cast member: its longitudinal draw comes first (member 0 is the ego: no draws)
        self.calls += 1
        u0, u1 = self._u(0)
        rad = np.sqrt(-2.0 * np.log(1.0 - u0))
        return rad * np.cos(2 * np.pi * u1) if self.calls % 2 == 1 else rad * np.sin(2 * np.pi * u1)

    def choice(self, seq):
        ud, _ = self._u(1)
        return seq[min(int(ud * len(seq)), len(seq) - 1)]

    def uniform(self, low, high):
        _, ue = self._u(1)
        return low + ue * (high - low)


@pytest.mark.parametrize("scene", ["roundabout", "u-turn"])
def test_device_cast_reset_follows_make_vehicles(scene):
    """Device-side scripted reset (emulated device logic) == the host reset fed the same draws; keyed by the global env."""
    net, table, cfg, cfgd = T.roundabout_scene() if scene == "roundabout" else T.uturn_scene()
    emu = Emulator(cfg, table)
    emu.set_reset_params(scenes.cast_reset_params(scene, net, table, cfgd))
    E, seed, first, episode = 24, 41, 500, 2
    got = SimState.zeros(E, 16)
    emu.reset(got, seed, first, episode)
    rngs = [_CastDrawRng(emu, seed, first + e, episode) for e in range(E)]
    want = (reset_roundabout if scene == "roundabout" else reset_uturn)(rngs, net, table, cfgd, cfg, 16)
    want.env_i[abi.EI_EPISODE] = episode
    T.compare_states(got, want, 1e-12, f"{scene} device reset")
    assert (got.env_i[abi.EI_NVEH] == (5 if scene == "roundabout" else 7)).all() and (got.env_i[abi.EI_EGO] == 0).all()
    half = SimState.zeros(E // 2, 16)
    emu.reset(half, seed, first + E // 2, episode)
    np.testing.assert_array_equal(half.veh_d, got.veh_d[:, E // 2:])      # shard invariance
    if scene == "roundabout":
        dests = {table.road_keys[unpack_route(got.veh_i[abi.I_ROUTE_LEN, e, 1], [got.veh_i[w, e, 1] for w in abi.I_ROUTE_ROAD_WORDS],
                                              [got.veh_i[w, e, 1] for w in abi.I_ROUTE_LANE_WORDS])[-1][0]][1] for e in range(E)}
        assert dests == {"exr", "sxr", "nxr"}                                 # the destination draw covers all three exits
        fixed = scenes.cast_reset_params(scene, net, table, dict(cfgd, incoming_vehicle_destination=1))
        assert fixed.cast[1].n_dest == 1 and fixed.cast[1].dest[0] == 1


@pytest.mark.parametrize("name", ["uturn_ttc_steps.npz", "highway_ttc_n30_steps.npz"])
def test_time_to_collision_observation_vs_reference(name):
    """TimeToCollisionObservation (observation.py:114-151, finite_mdp.compute_ttc_grid): UTurnEnv's DEFAULT config, and a
    4-lane highway (lane padding on both sides, every speed row)."""
    g = T.golden(name)
    if name.startswith("uturn"):
        _, table, cfg, _ = T.uturn_scene(None)
    else:
        _, table, cfg, _ = T.highway_scene(30, 2.0, overrides={"observation": {"type": "TimeToCollision", "horizon": 10}})
    assert cfg.obs_type == abi.OBS_TTC and scenes.obs_shape(cfg) == g["obs"].shape[1:]
    for what, eng in _engines(cfg, table):
        st = T.batch_state(g, "before")
        obs, reward, term, trunc, _ = eng.step(st, g["action"].astype(np.int32))
        T.compare_states(st, T.batch_state(g, "after"), T.TOL_STEP, f"{name}, {what}")
        np.testing.assert_array_equal(obs.reshape(g["obs"].shape), g["obs"], err_msg=what)  # costs 0 / 0.5 / 1: exact
        np.testing.assert_allclose(reward, g["reward"], rtol=0, atol=1e-6)
        if "reset_obs" in g.files:
            np.testing.assert_array_equal(eng.observe(T.batch_state(g, "reset")).reshape(g["reset_obs"].shape), g["reset_obs"])
    assert 0 < (g["obs"] == 0.5).mean() and 0 < (g["obs"] == 1.0).mean() < 1  # both cost levels and free cells occur

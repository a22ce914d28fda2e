"""Reference-shipped IntersectionEnv configs pinned as whole seeded episodes (tests/golden/intersection_ep_*.npz, written
by make_golden.py from the unmodified reference): env.json as shipped ("shuffled" row order), env_5fps.json (3 sub-steps per
step, regulation ticks not aligned to steps), normalize_reward + destination None, destinations o2 / o3.

CPU here: the oracle and the emulated device logic, every step resynced to the reference's state.  The GPU forms are in
tests/test_gpu_episodes.py."""
import numpy as np
import pytest

from oracle import oracle as O
from topotrafficrl_b200 import abi
from topotrafficrl_b200.state import SimState
from tests import common as T
from tests.emu.emu import Emulator


def episode_arrays(name):
    """-> (golden, before, after): `before[k]` is the reference's state when step k began (the reset state for the first
    step of an episode, else the state after step k - 1)."""
    g = T.golden("intersection_ep_%s.npz" % name)
    after = T.batch_state(g, "after")
    reset = T.batch_state(g, "reset")
    n = int(g["n_steps"][0])
    src_after = np.arange(n) - 1
    first = np.zeros(n, bool)
    first[g["ep_first"]] = True
    ep_of = np.cumsum(first) - 1
    before = SimState.zeros(n, after.vcap, linear=after.lin is not None)
    per_vehicle = [(before.veh_d, after.veh_d, reset.veh_d), (before.veh_i, after.veh_i, reset.veh_i)]
    if after.lin is not None:
        per_vehicle.append((before.lin, after.lin, reset.lin))
    for dst, a, b in per_vehicle:
        dst[:, ~first] = a[:, src_after[~first]]
        dst[:, first] = b[:, ep_of[first]]
    for dst, a, b in ((before.env_i, after.env_i, reset.env_i), (before.env_d, after.env_d, reset.env_d)):
        dst[:, ~first] = a[:, src_after[~first]]
        dst[:, first] = b[:, ep_of[first]]
    return g, before, after


def shuffled_rows(rows, perm):
    """np_random.shuffle(obs[1:]) with the recorded permutation: new row i = old row perm[i] (per step)."""
    out = rows.copy()
    for k in range(rows.shape[0]):
        out[k, 1:] = rows[k, 1:][perm[k]]
    return out


def inverse_perm(perm):
    inv = np.empty_like(perm)
    np.put_along_axis(inv, perm, np.broadcast_to(np.arange(perm.shape[1], dtype=perm.dtype), perm.shape), axis=1)
    return inv


def compare_episode_states(got, want, what):
    """State after a whole env-step.  The recorded steering command (an output: it is recomputed every sub-step) of a
    vehicle that is nearly standing is ill-conditioned in the reference's own controller -- lateral and heading errors are
    divided by not_zero(speed) twice (controller.py:170-186), a gain of 2e5 rad/m at 0.01 m/s -- so last-ulp differences of
    the position show up at 1e-5 there: compared at 1e-6 above 1 m/s, at 1e-3 below."""
    T.compare_states(got, want, T.TOL_STEP, what, check_action=False)
    live = want.live_mask()
    fast = live & (np.abs(want.veh_d[abi.D_SPEED]) > 1.0)
    for f in (abi.D_STEERING, abi.D_ACCEL):
        d = np.abs(got.veh_d[f] - want.veh_d[f])
        assert d[fast].max(initial=0.0) <= T.TOL_STEP and d[live].max(initial=0.0) <= 1e-3, (what, f, d[fast].max(initial=0.0), d[live].max(initial=0.0))


def check_step_outputs(g, obs, reward, term, trunc):
    np.testing.assert_allclose(obs.reshape(g["obs"].shape), g["obs"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(reward, g["reward"], rtol=0, atol=1e-6)
    assert (term.astype(bool) == g["terminated"]).all() and (trunc.astype(bool) == g["truncated"]).all()


def check_info(g, info, agent_reward, agent_terminated):
    """info[NINFO][E] float64 of the device / emulator against the reference's info dict of every step."""
    want = g["info"]
    assert (info[abi.INFO_CRASHED] == want[:, 1]).all()
    np.testing.assert_allclose(info[abi.INFO_SPEED], want[:, 0], rtol=0, atol=1e-6)
    np.testing.assert_allclose(info[abi.INFO_REWARDS:abi.INFO_REWARDS + 4].T, want[:, 2:6], rtol=0, atol=1e-6)
    np.testing.assert_allclose(agent_reward, g["agents_rewards"], rtol=0, atol=1e-6)
    assert (agent_terminated.astype(bool) == g["agents_terminated"]).all()


@pytest.mark.parametrize("name", sorted(T.EPISODE_CONFIGS))
def test_oracle_follows_reference_episodes(name):
    over, _ = T.EPISODE_CONFIGS[name]
    g, st, want = episode_arrays(name)
    _, table, cfg, routes = T.intersection_scene(over)
    assert int(np.floor(cfg.simulation_frequency / cfg.policy_frequency)) == (3 if name == "5fps" else 15)
    orc = O.Oracle(cfg, table, routes)
    obs, reward, term, trunc, _ = orc.step(st, g["action"].astype(np.int32), T.draws_array(g["draw"]))
    compare_episode_states(st, want, name)
    check_step_outputs(g, shuffled_rows(obs.reshape(g["obs"].shape), g["perm"]), reward, term, trunc)
    np.testing.assert_allclose(st.env_d[abi.ED_TIME], want.env_d[abi.ED_TIME], rtol=0, atol=1e-9)
    if name == "5fps":  # regulation ticks fall on different sub-steps of consecutive steps (regulation.py:28-32)
        assert len(set((want.env_i[abi.EI_ROAD_STEPS] % 7).tolist())) == 7
    if name in ("envjson", "normdest"):
        assert (g["perm"] != np.arange(g["perm"].shape[1])).any()


@pytest.mark.parametrize("name", sorted(T.EPISODE_CONFIGS))
def test_emulated_device_logic_follows_reference_episodes(name):
    """The device code (ttrl_core.cuh built for the host): injected spawn draws and row permutation, info outputs."""
    over, _ = T.EPISODE_CONFIGS[name]
    g, st, want = episode_arrays(name)
    _, table, cfg, routes = T.intersection_scene(over)
    emu = Emulator(cfg, table, routes)
    obs, reward, term, trunc, _ = emu.step(st, g["action"].astype(np.int32), T.draws_array(g["draw"]), inv_perm=inverse_perm(g["perm"]))
    compare_episode_states(st, want, name)
    check_step_outputs(g, obs, reward, term, trunc)
    check_info(g, emu.info, emu.agent_reward, emu.agent_terminated)


def test_device_drawn_shuffle_is_a_permutation_keyed_by_env_and_step():
    """order == "shuffled" without an injected permutation: rows 1.. are a Philox-keyed permutation of the unshuffled rows,
    different per env / step / seed, reproducible for the same key."""
    over, _ = T.EPISODE_CONFIGS["envjson"]
    g, st, _ = episode_arrays("envjson")
    _, table, cfg, routes = T.intersection_scene(over)
    emu = Emulator(cfg, table, routes)
    acts, draws = g["action"].astype(np.int32), T.draws_array(g["draw"])
    plain = emu.step(st.copy(), acts, draws)[0].reshape(-1, 15, 7)
    a = emu.step(st.copy(), acts, draws, seed=11)[0].reshape(-1, 15, 7)
    b = emu.step(st.copy(), acts, draws, seed=11)[0].reshape(-1, 15, 7)
    c = emu.step(st.copy(), acts, draws, seed=12)[0].reshape(-1, 15, 7)
    d = emu.step(st.copy(), acts, draws, seed=11, first_env=1000)[0].reshape(-1, 15, 7)
    np.testing.assert_array_equal(a, b)
    assert not np.array_equal(a, c) and not np.array_equal(a, d) and not np.array_equal(a, plain)
    for k in range(a.shape[0]):
        np.testing.assert_array_equal(a[k, 0], plain[k, 0])
        assert sorted(map(tuple, a[k, 1:])) == sorted(map(tuple, plain[k, 1:]))
    # every row position is reached: the permutation is not a fixed rotation
    moved = np.array([[tuple(a[k, 1 + r]) != tuple(plain[k, 1 + r]) for r in range(14)] for k in range(a.shape[0])])
    assert moved.any(axis=0).all()

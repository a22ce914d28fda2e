"""Parity tests proper: the CUDA path, called through the C ABI (topotrafficrl_b200.sim.Sim ->
libttrl_b200.so), against (a) golden vectors produced by the unmodified reference and (b) the CPU oracle on
the same seeded inputs.  Bars: discrete fields (lane indices, target lanes, crash/impact/yield flags, routes,
counters, terminated/truncated, actions) bit-exact; continuous state within TOL_SUBSTEP = 1e-9 per resynced
sub-step and TOL_STEP = 1e-6 per env-step (float64 on both sides; the north-star tolerance is 1e-4 m / 1e-5 rad).
Run on the B200 box:  python -m pytest tests -m gpu
"""
import numpy as np
import pytest

from topotrafficrl_b200 import abi, scenes
from tests import common as T

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch


def _sim(cfg, table, E, V, routes=None):
    from topotrafficrl_b200.sim import Sim
    return Sim(cfg, table, E, V, 0, routes)


def _scene(scene, over=None):
    if scene == "intersection":
        _, table, cfg, routes = T.intersection_scene(over)
        return cfg, table, routes
    _, table, cfg, _ = T.highway_scene(int(scene[7:]), overrides=over)
    return cfg, table, None


def _dev_step(sim, actions):
    torch = _torch()
    E = sim.num_envs
    a = torch.as_tensor(np.ascontiguousarray(actions, dtype=np.int32), device="cuda")
    obs = torch.zeros(E * sim.obs_size, dtype=torch.float32, device="cuda")
    rew = torch.zeros(E, dtype=torch.float32, device="cuda")
    term = torch.zeros(E, dtype=torch.uint8, device="cuda")
    trunc = torch.zeros(E, dtype=torch.uint8, device="cuda")
    sim.step_ptr(a.data_ptr(), obs.data_ptr(), rew.data_ptr(), term.data_ptr(), trunc.data_ptr(), 0)
    torch.cuda.synchronize()
    return obs.cpu().numpy().reshape(E, -1), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()


@pytest.mark.parametrize("name,scene", [
    ("intersection_substeps.npz", "intersection"),
    ("highway_n8_substeps.npz", "highway8"),
    ("highway_n50_substeps.npz", "highway50"),
    ("highway_n200_substeps.npz", "highway200"),
    ("highway_grid_n40_substeps.npz", "highway40"),
])
def test_substep_vs_reference_golden(name, scene):
    """Resynced sub-step parity: inject the reference's state, one device sub-step, compare with the reference."""
    torch = _torch()
    g = T.golden(name)
    cfg, table, routes = _scene(scene)
    st = T.batch_state(g, "before")
    sim = _sim(cfg, table, st.num_envs, st.vcap, routes)
    sim.set_state(st)
    a = torch.as_tensor(g["action"].astype(np.int32), device="cuda")
    sim.substep_ptr(a.data_ptr(), 0)
    got = sim.get_state()
    T.compare_states(got, T.batch_state(g, "after"), T.TOL_SUBSTEP, name)
    sim.close()


@pytest.mark.parametrize("name,scene,over", [
    ("intersection_steps_kin.npz", "intersection", None),
    ("intersection_steps_grid_dense.npz", "intersection", T.GRID_DENSE),
    ("intersection_steps_grid_road.npz", "intersection", T.GRID_ROAD),
    ("highway_n8_steps.npz", "highway8", None),
    ("highway_n50_steps.npz", "highway50", None),
    ("highway_grid_n40_steps.npz", "highway40", T.HIGHWAY_GRID),
])
def test_step_vs_reference_golden(name, scene, over):
    """Full env.step(): 15 sub-steps + observation + reward + flags (+ clear/spawn with the reference's draws)."""
    g = T.golden(name)
    cfg, table, routes = _scene(scene, over)
    st = T.batch_state(g, "before")
    sim = _sim(cfg, table, st.num_envs, st.vcap, routes)
    sim.set_state(st)
    if "draw" in g.files:
        sim.inject_spawn(T.draws_array(g["draw"]))
    obs, reward, term, trunc = _dev_step(sim, g["action"])
    T.compare_states(sim.get_state(), T.batch_state(g, "after"), T.TOL_STEP, name)
    np.testing.assert_allclose(obs.reshape(g["obs"].shape), g["obs"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(reward, g["reward"], rtol=0, atol=1e-6)
    assert (term.astype(bool) == g["terminated"]).all() and (trunc.astype(bool) == g["truncated"]).all()
    sim.close()


@pytest.mark.parametrize("n,density,E,steps,grid", [(50, 2.0, 512, 6, False), (200, 4.0, 64, 2, True), (20, 1.0, 256, 4, False)])
def test_free_running_vs_oracle(n, density, E, steps, grid):
    """Seeded scenes from the product generator, free-running env-steps, device vs CPU oracle."""
    from oracle import oracle as O
    over = T.HIGHWAY_GRID if grid else None
    _, table, cfg, cfgd = T.highway_scene(n, density, overrides=over)
    a = scenes.make_highway_state(E, cfgd, seed=11)
    b = a.copy()
    sim = _sim(cfg, table, E, n)
    sim.set_state(a)
    orc = O.Oracle(cfg, table, threads=8)
    rng = np.random.default_rng(0)
    for k in range(steps):
        act = rng.integers(0, 5, size=E).astype(np.int32)
        obs, rew, term, trunc = _dev_step(sim, act)
        oo, orr, ot, ou, _ = orc.step(b, act)
        T.compare_states(sim.get_state(), b, 1e-7, f"N={n} step {k}")
        np.testing.assert_allclose(obs, oo, rtol=0, atol=2e-6)
        np.testing.assert_allclose(rew, orr, rtol=0, atol=1e-6)
        assert (term == ot).all() and (trunc == ou).all()
    sim.close()


def test_full_size_baseline_config_vs_oracle_and_properties():
    """BASELINE config 2 at full size (4096 envs x 50 vehicles): two env-steps against the oracle, then
    size-independent properties over a longer run: shard invariance (one 4096-env sim == two 2048-env sims),
    determinism, vehicle count conservation, lane indices in range, crashed flags monotone within an episode."""
    from oracle import oracle as O
    E, n = 4096, 50
    _, table, cfg, cfgd = T.highway_scene(n, 2.0)
    st0 = scenes.make_highway_state(E, cfgd, seed=0)
    sim = _sim(cfg, table, E, n)
    sim.set_state(st0)
    ref = st0.copy()
    orc = O.Oracle(cfg, table, threads=8)
    rng = np.random.default_rng(5)
    acts = [rng.integers(0, 5, size=E).astype(np.int32) for _ in range(6)]
    for k in range(2):
        obs, rew, term, trunc = _dev_step(sim, acts[k])
        oo, orr, ot, ou, _ = orc.step(ref, acts[k])
        T.compare_states(sim.get_state(), ref, 1e-7, f"full-size step {k}")
        np.testing.assert_allclose(obs, oo, rtol=0, atol=2e-6)
        assert (term == ot).all() and (trunc == ou).all()
    halves = [_sim(cfg, table, E // 2, n), _sim(cfg, table, E // 2, n)]
    halves[0].set_state(st0.slice_envs(0, E // 2))
    halves[1].set_state(st0.slice_envs(E // 2, E))
    sim.set_state(st0)
    prev_crashed = np.zeros((E, n), bool)
    for k in range(6):
        obs, rew, term, trunc = _dev_step(sim, acts[k])
        parts = [_dev_step(h, acts[k][i * E // 2:(i + 1) * E // 2]) for i, h in enumerate(halves)]
        np.testing.assert_array_equal(obs, np.concatenate([p[0] for p in parts]))
        np.testing.assert_array_equal(rew, np.concatenate([p[1] for p in parts]))
        st = sim.get_state()
        assert (st.env_i[abi.EI_NVEH] == n).all()
        assert ((st.veh_i[abi.I_LANE] >= 0) & (st.veh_i[abi.I_LANE] < 4)).all()
        crashed = (st.veh_i[abi.I_FLAGS] & abi.FL_CRASHED) != 0
        assert (crashed | ~prev_crashed).all()
        prev_crashed = crashed
        assert np.isfinite(st.veh_d).all()
    both = [h.get_state() for h in halves]
    np.testing.assert_array_equal(sim.get_state().veh_d, np.concatenate([b.veh_d for b in both], axis=1))
    np.testing.assert_array_equal(sim.get_state().veh_i, np.concatenate([b.veh_i for b in both], axis=1))
    for s in [sim] + halves:
        s.close()


def test_autoreset_from_pool_and_stats():
    """Finished envs restart from the reset pool inside the step kernel; device == oracle incl. the reset."""
    from oracle import oracle as O
    E, n = 128, 50
    _, table, cfg, cfgd = T.highway_scene(n, 2.0, overrides={"duration": 3})
    a = scenes.make_highway_state(E, cfgd, seed=2)
    b = a.copy()
    sim = _sim(cfg, table, E, n)
    sim.set_state(a)
    sim.set_reset_pool(a)
    sim.set_autoreset(True)
    orc = O.Oracle(cfg, table, threads=8)
    orc.set_reset_pool(a)
    orc.set_autoreset(True)
    stats = np.zeros(8)
    rng = np.random.default_rng(1)
    n_done = 0
    for k in range(7):
        act = rng.integers(0, 5, size=E).astype(np.int32)
        obs, rew, term, trunc = _dev_step(sim, act)
        oo, orr, ot, ou, _ = orc.step(b, act, stats=stats)
        n_done += int((term | trunc).sum())
        T.compare_states(sim.get_state(), b, 1e-7, f"autoreset step {k}")
        np.testing.assert_allclose(obs, oo, rtol=0, atol=2e-6)
        assert (term == ot).all() and (trunc == ou).all()
    assert n_done >= 2 * E  # duration 3 -> every env finished at least twice in 7 steps
    s = sim.stats()
    assert s.episodes == n_done == stats[0]
    np.testing.assert_allclose([s.total_return, s.total_length, s.crashes, s.vehicle_steps, s.env_steps],
                               [stats[1], stats[2], stats[3], stats[6], stats[7]], rtol=1e-9)
    sim.close()


def test_host_buffer_step_equals_device_pointer_step():
    E, n = 64, 50
    _, table, cfg, cfgd = T.highway_scene(n, 2.0)
    st = scenes.make_highway_state(E, cfgd, seed=4)
    s1, s2 = _sim(cfg, table, E, n), _sim(cfg, table, E, n)
    s1.set_state(st)
    s2.set_state(st)
    act = np.random.default_rng(3).integers(0, 5, size=E).astype(np.int32)
    for _ in range(3):
        o1, r1, t1, u1 = _dev_step(s1, act)
        o2, r2, t2, u2 = s2.step_host(act)
        np.testing.assert_array_equal(o1, o2)
        np.testing.assert_array_equal(r1, r2)
        assert (t1 == t2).all() and (u1 == u2).all()
    s1.close()
    s2.close()


def test_intersection_reset_and_seeded_episodes_match_reference():
    """Single-env front end end to end: IntersectionEnv.reset(seed) + step(a) reproduce the reference episode
    (state at reset, then obs / reward / flags per step, with the env's own numpy RNG stream for spawns)."""
    from topotrafficrl_b200.envs import IntersectionEnv
    g = T.golden("intersection_reset.npz")
    env = IntersectionEnv(config={"observation": dict(scenes.INTERSECTION_CONFIG["observation"], order="sorted")})
    for k, seed in enumerate(g["seed"]):
        obs, _ = env.reset(seed=int(seed))
        np.testing.assert_allclose(obs, g["obs"][k], rtol=0, atol=2e-6)
        T.compare_states(env.sim.get_state(), T.batch_state(g, "state", slice(k, k + 1)), 1e-9, f"reset seed {seed}")
    env.close()
    gs = T.golden("intersection_steps_kin.npz")
    env = IntersectionEnv()
    k = 0
    for seed in range(100, 112):
        env.reset(seed=seed)
        done = False
        while not done:
            obs, reward, term, trunc, info = env.step(int(gs["action"][k]))
            np.testing.assert_allclose(obs, gs["obs"][k], rtol=0, atol=1e-5, err_msg=f"seed {seed} k {k}")
            assert abs(reward - gs["reward"][k]) <= 1e-6
            assert term == bool(gs["terminated"][k]) and trunc == bool(gs["truncated"][k])
            assert info["crashed"] == bool(gs["crashed"][k]) and abs(info["speed"] - gs["speed"][k]) <= 1e-6
            done = term or trunc
            k += 1
    assert k == len(gs["action"])
    env.close()


def test_intersection_device_spawn_draws_and_vector_env():
    """Throughput mode of the intersection scene: device-side Philox spawn draws, autoreset, E=256.
    Property checks + agreement with the oracle when the oracle is fed the device's own draws."""
    from topotrafficrl_b200.vector_env import TTRLVectorEnv
    torch = _torch()
    env = TTRLVectorEnv(256, scene="intersection", seed=3)
    obs, _ = env.reset()
    assert obs.shape == (256, 15, 7)
    rng = np.random.default_rng(0)
    counts = []
    for k in range(20):
        a = torch.as_tensor(rng.integers(0, 3, size=256).astype(np.int32), device="cuda")
        obs, rew, term, trunc, _ = env.step(a)
        st = env.get_state()
        nveh = st.env_i[abi.EI_NVEH]
        counts.append(nveh.mean())
        assert (nveh >= 1).all() and (nveh <= 32).all()
        assert np.isfinite(st.veh_d).all() and torch.isfinite(obs).all()
        assert ((st.veh_i[abi.I_LANE] >= 0) & (st.veh_i[abi.I_LANE] < 20)).all()
    s = env.stats()
    assert s["episodes"] > 0 and s["env_steps"] == 256 * 20
    assert 3 < np.mean(counts) < 20
    env.close()


def test_missing_device_arguments_fail_loudly():
    from topotrafficrl_b200._lib import TTRLError
    _, table, cfg, _ = T.highway_scene(50)
    with pytest.raises(TTRLError):
        _sim(cfg, table, 0, 50)
    with pytest.raises(TTRLError):
        _sim(cfg, table, 4, 1000)


@pytest.mark.parametrize("n,density", [(12, 2.0), (20, 2.0), (50, 3.0), (60, 3.0), (100, 3.0), (200, 4.0)])  # kernel sets 16, 24, 50, 64, 100, 200
def test_task_queue_stress_vs_oracle(n, density):
    """Pair-queue overflow (serial fallback), several MOBIL batches in one sub-step, many ongoing lane changes and
    ragged vehicle counts (1..n): device == oracle per resynced-free sub-step and per env-step."""
    from oracle import oracle as O
    torch = _torch()
    E = 48
    _, table, cfg, cfgd = T.highway_scene(n, density)
    a = T.stress_states(cfgd, E, n, seed=4)
    b = a.copy()
    sim = _sim(cfg, table, E, n)
    sim.set_state(a)
    orc = O.Oracle(cfg, table, threads=8)
    for k in range(4):
        sim.substep_ptr(None, 0)
        orc.substep(b, None)
        T.compare_states(sim.get_state(), b, 1e-9, f"stress sub-step {k}")
    assert ((b.veh_i[abi.I_FLAGS] & abi.FL_CRASHED) != 0).sum() > n
    rng = np.random.default_rng(1)
    for k in range(3):
        act = rng.integers(0, 5, size=E).astype(np.int32)
        obs, rew, term, trunc = _dev_step(sim, act)
        oo, orr, ot, ou, _ = orc.step(b, act)
        T.compare_states(sim.get_state(), b, 1e-7, f"stress step {k}")
        np.testing.assert_allclose(obs, oo, rtol=0, atol=2e-6)
        assert (term == ot).all() and (trunc == ou).all()
    sim.close()


def test_device_reset_equals_emulated_logic_and_autoreset_fresh_episodes():
    """ttrl_sim_reset on the GPU == the same device logic run on the host (which tests/test_host_logic.py ties to the
    reference's _make_vehicles); device autoreset gives every finished env a NEW episode; oracle agreement on the
    dynamics that follow a device reset."""
    from oracle import oracle as O
    from tests.emu.emu import Emulator
    from topotrafficrl_b200.state import SimState
    torch = _torch()
    # intersection
    net, table, cfg, routes = T.intersection_scene()
    cfgd = scenes.merged_config(scenes.INTERSECTION_CONFIG, None)
    rp = scenes.intersection_reset_params(cfgd)
    E = 64
    sim = _sim(cfg, table, E, 24, routes)
    sim.set_reset_params(rp)
    sim.seed(12345, 500)
    sim.reset_device()
    got = sim.get_state()
    emu = Emulator(cfg, table, routes)
    emu.set_reset_params(rp)
    want = SimState.zeros(E, 24)
    emu.reset(want, 12345, 500, 0)
    T.compare_states(got, want, 1e-9, "device reset (intersection)")
    # masked reset: only the flagged envs restart, with episode + 1
    mask = torch.zeros(E, dtype=torch.uint8, device="cuda")
    mask[::3] = 1
    sim.reset_device(mask.data_ptr())
    after = sim.get_state()
    keep = np.ones(E, bool)
    keep[::3] = False
    np.testing.assert_array_equal(after.veh_d[:, keep], got.veh_d[:, keep])
    assert (after.env_i[abi.EI_EPISODE, ~keep] == 1).all() and (after.env_i[abi.EI_EPISODE, keep] == 0).all()
    assert not np.array_equal(after.veh_d[:, ~keep], got.veh_d[:, ~keep])
    sim.close()
    # highway: device reset, then dynamics vs the oracle, then device autoreset inside the step kernel
    _, table, cfg, cfgd = T.highway_scene(50, 2.0, overrides={"duration": 2})
    sim = _sim(cfg, table, 96, 50)
    sim.set_reset_params(scenes.highway_reset_params(cfgd))
    sim.seed(99, 0)
    sim.reset_device()
    ref = sim.get_state()
    first = ref.copy()
    orc = O.Oracle(cfg, table, threads=8)
    rng = np.random.default_rng(3)
    act = rng.integers(0, 5, size=96).astype(np.int32)
    obs, rew, term, trunc = _dev_step(sim, act)
    oo, orr, ot, ou, _ = orc.step(ref, act)
    T.compare_states(sim.get_state(), ref, 1e-7, "step after device reset")
    np.testing.assert_allclose(obs, oo, rtol=0, atol=2e-6)
    sim.set_autoreset("device")
    obs, rew, term, trunc = _dev_step(sim, act)  # duration 2 -> every env is truncated here and restarts
    assert (term | trunc).all()
    st = sim.get_state()
    assert (st.env_i[abi.EI_EPISODE] == 1).all() and (st.env_i[abi.EI_STEPS] == 0).all() and (st.env_d[abi.ED_TIME] == 0).all()
    assert (st.env_i[abi.EI_NVEH] == 50).all()
    assert not np.array_equal(st.veh_d[abi.D_X], first.veh_d[abi.D_X])  # a fresh scene, not the first one again
    emu = Emulator(cfg, table)
    emu.set_reset_params(scenes.highway_reset_params(cfgd))
    want = SimState.zeros(96, 50)
    emu.reset(want, 99, 0, 1)
    T.compare_states(st, want, 1e-9, "device autoreset (highway)")
    np.testing.assert_allclose(obs, emu.observe(want), rtol=0, atol=2e-6)
    sim.close()


def test_batched_device_autoreset_intersection():
    """Intersection scene, device autoreset: finished envs are queued by k_step and reset by the packed k_reset_list
    pass; the result must equal the reset logic run on the host for (seed, global env, episode + 1)."""
    from tests.emu.emu import Emulator
    from topotrafficrl_b200.state import SimState
    torch = _torch()
    net, table, cfg, routes = T.intersection_scene({"duration": 2})
    cfgd = scenes.merged_config(scenes.INTERSECTION_CONFIG, {"duration": 2})
    rp = scenes.intersection_reset_params(cfgd)
    E = 100  # not a multiple of the envs per CTA
    sim = _sim(cfg, table, E, 24, routes)
    sim.set_reset_params(rp)
    sim.seed(4242, 7)
    sim.reset_device()
    sim.set_autoreset("device")
    act = np.ones(E, np.int32)
    _dev_step(sim, act)
    assert (sim.get_state().env_i[abi.EI_EPISODE] == 0).sum() > 0  # not everybody crashed in the first second
    obs, rew, term, trunc = _dev_step(sim, act)  # time 2 >= duration: every env is done here
    assert (term | trunc).all()
    st = sim.get_state()
    emu = Emulator(cfg, table, routes)
    emu.set_reset_params(rp)
    for episode in (1, 2):
        sel = np.nonzero(st.env_i[abi.EI_EPISODE] == episode)[0]
        if sel.size == 0:
            continue
        want = SimState.zeros(E, 24)
        emu.reset(want, 4242, 7, episode)
        for e in sel:
            T.compare_states(st.slice_envs(e, e + 1), want.slice_envs(e, e + 1), 1e-9, f"env {e} episode {episode}")
        np.testing.assert_allclose(obs[sel], emu.observe(want)[sel], rtol=0, atol=2e-6)
    assert (st.env_i[abi.EI_STEPS] == 0).all() and (st.env_i[abi.EI_DONE] == 0).all()
    sim.close()

"""The CUDA kernels' device logic (topotrafficrl_b200/csrc/ttrl_core.cuh), compiled for the host and run
phase by phase (tests/emu), against the reference-generated golden vectors and against the CPU oracle.
CPU only; the same comparisons run on the real kernels in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import oracle as O
from topotrafficrl_b200 import abi, scenes
from tests import common as T
from tests.emu.emu import Emulator


def _scene(scene, over=None):
    if scene == "intersection":
        _, table, cfg, routes = T.intersection_scene(over)
        return cfg, table, routes
    _, table, cfg, _ = T.highway_scene(int(scene[7:]), overrides=over)
    return cfg, table, None


@pytest.mark.parametrize("name,scene", [
    ("intersection_substeps.npz", "intersection"),
    ("highway_n8_substeps.npz", "highway8"),
    ("highway_n50_substeps.npz", "highway50"),
    ("highway_n200_substeps.npz", "highway200"),
    ("highway_grid_n40_substeps.npz", "highway40"),
])
def test_substep_vs_golden(name, scene):
    g = T.golden(name)
    cfg, table, routes = _scene(scene)
    emu = Emulator(cfg, table, routes)
    st = T.batch_state(g, "before")
    emu.substep(st, g["action"].astype(np.int32))
    T.compare_states(st, T.batch_state(g, "after"), T.TOL_SUBSTEP, name)


@pytest.mark.parametrize("name,scene,over", [
    ("intersection_steps_kin.npz", "intersection", None),
    ("intersection_steps_grid_dense.npz", "intersection", T.GRID_DENSE),
    ("intersection_steps_grid_road.npz", "intersection", T.GRID_ROAD),
    ("highway_n8_steps.npz", "highway8", None),
    ("highway_n50_steps.npz", "highway50", None),
    ("highway_grid_n40_steps.npz", "highway40", T.HIGHWAY_GRID),
])
def test_step_vs_golden(name, scene, over):
    g = T.golden(name)
    cfg, table, routes = _scene(scene, over)
    emu = Emulator(cfg, table, routes)
    st = T.batch_state(g, "before")
    draws = T.draws_array(g["draw"]) if "draw" in g.files else None
    obs, reward, term, trunc, _ = emu.step(st, g["action"].astype(np.int32), draws)
    T.compare_states(st, T.batch_state(g, "after"), T.TOL_STEP, name)
    np.testing.assert_allclose(obs.reshape(g["obs"].shape), g["obs"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(reward, g["reward"], rtol=0, atol=1e-6)
    assert (term.astype(bool) == g["terminated"]).all() and (trunc.astype(bool) == g["truncated"]).all()


@pytest.mark.parametrize("n,density,steps", [(50, 2.0, 12), (200, 4.0, 2)])
def test_free_running_vs_oracle(n, density, steps):
    """Free-running env-steps from the product's own scene generator: emulated device logic == oracle."""
    _, table, cfg, cfgd = T.highway_scene(n, density)
    emu, orc = Emulator(cfg, table), O.Oracle(cfg, table)
    a = scenes.make_highway_state(6, cfgd, seed=3)
    b = a.copy()
    rng = np.random.default_rng(0)
    for k in range(steps):
        act = rng.integers(0, 5, size=6).astype(np.int32)
        oa = emu.step(a, act)
        ob = orc.step(b, act)
        T.compare_states(a, b, 1e-7, f"step {k}")
        np.testing.assert_allclose(oa[0], ob[0], rtol=0, atol=1e-6)
        np.testing.assert_allclose(oa[1], ob[1], rtol=0, atol=1e-6)
        assert (oa[2] == ob[2]).all() and (oa[3] == ob[3]).all()


@pytest.mark.parametrize("n,density", [(20, 2.0), (50, 3.0)])
def test_task_queue_stress_vs_oracle(n, density):
    """Pair-queue overflow (serial fallback), several MOBIL batches in one sub-step, many ongoing lane changes and
    ragged vehicle counts: emulated device logic == oracle, per sub-step and per env-step."""
    _, table, cfg, cfgd = T.highway_scene(n, density)
    emu, orc = Emulator(cfg, table), O.Oracle(cfg, table)
    a = T.stress_states(cfgd, 9, n, seed=4)
    b = a.copy()
    for k in range(4):
        emu.substep(a, None)
        orc.substep(b, None)
        T.compare_states(a, b, 1e-9, f"stress sub-step {k}")
    assert ((a.veh_i[abi.I_FLAGS] & abi.FL_CRASHED) != 0).sum() > n  # the pile-ups did crash
    rng = np.random.default_rng(1)
    for k in range(3):
        act = rng.integers(0, 5, size=9).astype(np.int32)
        oa, ob = emu.step(a, act), orc.step(b, act)
        T.compare_states(a, b, 1e-7, f"stress step {k}")
        np.testing.assert_allclose(oa[0], ob[0], rtol=0, atol=2e-6)
        assert (oa[2] == ob[2]).all() and (oa[3] == ob[3]).all()


@pytest.mark.parametrize("scene", ["roundabout", "uturn"])
def test_curved_lane_task_list_on_scenes_that_do_not_use_it_by_default(scene, monkeypatch):
    """closest_lane_tasks (straight pass + curved-lane task list + merge; sine-lane bounds, NaN cache entries read on demand) is
    only enabled where it pays (use_arc_tasks: the intersection); forced on here, it must reproduce the reference on the ring and the
    sine lanes of the roundabout and on the u-turn just the same."""
    from tests.emu.emu import Emulator
    monkeypatch.setenv("TTRL_EMU_ARC_TASKS", "1")
    _, table, cfg, _ = T.roundabout_scene() if scene == "roundabout" else T.uturn_scene()
    emu = Emulator(cfg, table)
    gs, g = T.golden(f"{scene}_substeps.npz"), T.golden(f"{scene}_steps.npz")
    st = T.batch_state(gs, "before")
    emu.substep(st, gs["action"].astype(np.int32))
    T.compare_states(st, T.batch_state(gs, "after"), T.TOL_SUBSTEP, f"{scene} sub-step, task form")
    st = T.batch_state(g, "before")
    obs, reward, term, trunc, _ = emu.step(st, g["action"].astype(np.int32))
    T.compare_states(st, T.batch_state(g, "after"), T.TOL_STEP, f"{scene} step, task form")
    np.testing.assert_allclose(obs.reshape(g["obs"].shape), g["obs"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("n", [50, 200])
def test_collision_candidate_scan_vs_oracle(n):
    """The collision candidate scan (collide_all: float pre-filter on x with the lower index's guard radius, half-ring over a
    doubled list) on adversarial layouts: clusters far apart, negative x, x on round values +- 1e-9, very fast vehicles (guard
    radius > 10 m), one long pile-up, distant clusters interleaved in index order -- emulated device logic == oracle per sub-step,
    and the collisions did happen.  (Written for a cell-list form of the scan, which was measured slower and dropped.)"""
    _, table, cfg, cfgd = T.highway_scene(n, 3.0)
    emu, orc = Emulator(cfg, table), O.Oracle(cfg, table)
    a = scenes.make_highway_state(6, cfgd, seed=11)
    rng = np.random.default_rng(5)
    x = a.veh_d[abi.D_X]
    x[0, : n // 2] += 640.0                                   # two clusters exactly one bucket span apart
    x[1, :n] -= x[1, :n].mean() + 3.0                          # negative x, cluster across 0
    x[2, :n] = np.round(x[2, :n] / 10.0) * 10.0 + rng.choice([-1e-9, 0.0, 1e-9], size=n)  # on cell boundaries
    a.veh_d[abi.D_SPEED, 3, : n // 3] = 75.0                   # guard radius 5.39 + 75 / 15 > cell width
    x[4, :n] = 200.0 + np.arange(n) * 2.6                      # one long pile-up across many cells
    x[5, : n // 2] += 1280.0
    x[5, n // 2: n] = x[5, : n - n // 2] - 1280.0 + rng.uniform(-3, 3, size=n - n // 2)  # aliased buckets that DO hold real neighbours two spans away
    b = a.copy()
    for k in range(6):
        emu.substep(a, None)
        orc.substep(b, None)
        T.compare_states(a, b, 1e-9, f"scan sub-step {k}")
    crashed = (a.veh_i[abi.I_FLAGS] & abi.FL_CRASHED) != 0
    assert crashed[4].sum() > n // 2 and crashed[2].sum() > 0

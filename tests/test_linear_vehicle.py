"""LinearVehicle traffic (ttrl_env/vehicle/behavior.py:350-558): what
scripts/configs/RoundaboutEnv/env.json, env_route_0/1.json and IntersectionEnv/env_linear.json configure.  CPU: the oracle
and the emulated device logic against golden vectors of the unmodified reference (make_golden.py `linear`), the host-driven
resets against the reference's seeded resets (randomize_behavior draws five uniforms instead of one), the device-side
scripted reset against the host one.  The intersection episodes (env_linear.json) are in tests/test_episode_goldens.py."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from topotrafficrl_b200 import abi, factory, scenes
from topotrafficrl_b200._gym import np_random
from topotrafficrl_b200.reset import linear_parameters, reset_intersection, reset_roundabout
from topotrafficrl_b200.state import SimState
from tests import common as T
from tests.emu.emu import Emulator
from tests.test_host_logic import _EmuBackend


def _engines(cfg, table, routes=None):
    return [("oracle", O.Oracle(cfg, table, routes)), ("emulated device logic", Emulator(cfg, table, routes))]


def test_config_translation():
    _, table, cfg, _ = T.roundabout_scene(T.LINEAR)
    assert cfg.vehicle_model == abi.VEHICLE_LINEAR and cfg.time_wanted == 2.5          # LinearVehicle.TIME_WANTED behavior.py:373
    assert list(cfg.lin_default) == [0.3, 0.3, 2.0, 5.0, 5.0 * (1 / 0.6)]
    np.testing.assert_allclose(list(cfg.lin_lo), [0.15, 0.15, 1.0, 4.93, 5.0 / 0.6 - 1.5], rtol=0, atol=1e-15)
    _, table, cfg, _ = T.roundabout_scene()
    assert cfg.vehicle_model == abi.VEHICLE_IDM and cfg.time_wanted == 1.5
    for bad in ("ttrl_env.vehicle.uncertainty.estimation.MultipleModelVehicle", "ttrl_env.vehicle.behavior.AggressiveVehicle"):
        with pytest.raises(NotImplementedError):
            T.roundabout_scene({"other_vehicles_type": bad})
    with pytest.raises(NotImplementedError):  # the synthetic highway has no LinearVehicle form
        T.highway_scene(8, overrides=T.LINEAR)


@pytest.mark.parametrize("name,over", [("roundabout_linear", T.LINEAR), ("roundabout_linear_route0", dict(T.LINEAR, incoming_vehicle_destination=0))])
def test_roundabout_linear_substeps_and_steps_vs_reference(name, over):
    _, table, cfg, _ = T.roundabout_scene(over)
    g = T.golden(f"{name}_steps.npz")
    for what, eng in _engines(cfg, table):
        if os.path.exists(os.path.join(T.GOLDEN, f"{name}_substeps.npz")):
            gs = T.golden(f"{name}_substeps.npz")
            st = T.batch_state(gs, "before")
            assert st.lin is not None
            eng.substep(st, gs["action"].astype(np.int32))
            T.compare_states(st, T.batch_state(gs, "after"), T.TOL_SUBSTEP, f"{name} sub-step, {what}")
        st = T.batch_state(g, "before")
        obs, reward, term, trunc, _ = eng.step(st, g["action"].astype(np.int32))
        T.compare_states(st, T.batch_state(g, "after"), T.TOL_STEP, f"{name} step, {what}")
        np.testing.assert_allclose(obs.reshape(g["obs"].shape), g["obs"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(reward, g["reward"], rtol=0, atol=1e-6)
        assert (term.astype(bool) == g["terminated"]).all() and (trunc.astype(bool) == g["truncated"]).all()
    # the traffic really is linear: parameters randomised per vehicle inside ACCELERATION_RANGE / STEERING_RANGE, zero for the ego
    lin, n = g["reset_lin"], g["reset_ei"][:, abi.EI_NVEH]
    assert (lin[:, :, 0] == 0).all() and len(np.unique(lin[:, 0, 1:5].round(9))) > 8
    lo, hi = np.array(abi.LINEAR_RANGE_LO)[None, :, None], np.array(abi.LINEAR_RANGE_HI)[None, :, None]
    assert ((lin[:, :, 1:5] >= lo) & (lin[:, :, 1:5] <= hi)).all() and (n == 5).all()


def test_linear_dynamics_differ_from_idm():
    """Guard against a silent IDM fallback: the same state stepped with IDM traffic does not follow the LinearVehicle golden."""
    _, table, cfg, _ = T.roundabout_scene()
    g = T.golden("roundabout_linear_steps.npz")
    st = T.batch_state(g, "before")
    st.lin = None
    O.Oracle(cfg, table).step(st, g["action"].astype(np.int32))
    want = T.batch_state(g, "after")
    live = want.live_mask()
    assert np.abs(st.veh_d[abi.D_SPEED] - want.veh_d[abi.D_SPEED])[live].max() > 0.1


def test_host_resets_with_linear_traffic_vs_reference():
    """reset_roundabout / reset_intersection with gymnasium-seeded Generators == the reference's reset(seed) with
    LinearVehicle traffic: positions, routes and the five randomised parameters per vehicle."""
    net, table, cfg, cfgd = T.roundabout_scene(T.LINEAR)
    g = T.golden("roundabout_linear_steps.npz")
    st = reset_roundabout([np_random(int(s))[0] for s in g["reset_seed"]], net, table, cfgd, cfg, 16)
    T.compare_states(st, T.batch_state(g, "reset"), 1e-12, "roundabout reset, LinearVehicle traffic")
    over, seeds = T.EPISODE_CONFIGS["linear"]
    g = T.golden("intersection_ep_linear.npz")
    net, table, cfg, routes = T.intersection_scene(over)
    cfgd = scenes.merged_config(scenes.INTERSECTION_CONFIG, over)
    emu = Emulator(cfg, table, routes)
    backend = _EmuBackend(emu, len(seeds), 32)
    st = reset_intersection(backend, [np_random(int(s))[0] for s in seeds], net, table, cfgd, cfg)
    T.compare_states(st, T.batch_state(g, "reset"), 1e-9, "intersection reset, LinearVehicle traffic")
    assert st.lin is not None and (st.lin[:, st.live_mask()] != 0).any()


class _LinearCastDrawRng:
    """The draws the DEVICE-side scripted reset makes for one env with LinearVehicle traffic (ttrl_core.cuh: reset_cast): the
    two normals and the destination like IDM traffic, then five uniforms from stream indices 0x300 + 4 m .. + 2."""

    def __init__(self, emu, seed, env, episode):
        self.emu, self.seed, self.env, self.episode = emu, seed, env, episode
        self.member, self.calls, self.sizes = 0, 0, []

    def _u(self, idx):
        return self.emu.reset_uniforms(self.seed, self.env, self.episode, idx)

    def normal(self):
        if self.calls % 2 == 0:
            self.member += 1
        self.calls += 1
        u0, u1 = self._u(0x200 + 2 * self.member)
        rad = np.sqrt(-2.0 * np.log(1.0 - u0))
        return rad * np.cos(2 * np.pi * u1) if self.calls % 2 == 1 else rad * np.sin(2 * np.pi * u1)

    def choice(self, seq):
        ud, _ = self._u(0x201 + 2 * self.member)
        return seq[min(int(ud * len(seq)), len(seq) - 1)]

    def uniform(self, size=None, low=0.0, high=1.0):
        lu = [x for k in range(3) for x in self._u(0x300 + 4 * self.member + k)]
        return np.array(lu[:3] if size == 3 else lu[3:5])


def test_device_cast_reset_with_linear_traffic():
    net, table, cfg, cfgd = T.roundabout_scene(T.LINEAR)
    emu = Emulator(cfg, table)
    emu.set_reset_params(scenes.cast_reset_params("roundabout", net, table, cfgd))
    E, seed, first, episode = 12, 5, 100, 1
    got = SimState.zeros(E, 16, linear=True)
    emu.reset(got, seed, first, episode)
    want = reset_roundabout([_LinearCastDrawRng(emu, seed, first + e, episode) for e in range(E)], net, table, cfgd, cfg, 16)
    want.env_i[abi.EI_EPISODE] = episode
    T.compare_states(got, want, 1e-12, "device cast reset, LinearVehicle traffic")
    # ... and the dynamics that follow: emulated device logic == oracle, free running
    orc = O.Oracle(cfg, table)
    a, b = got.copy(), got.copy()
    act = np.random.default_rng(0).integers(0, 5, size=E).astype(np.int32)
    for _ in range(3):
        oa, ob = emu.step(a, act), orc.step(b, act)
        T.compare_states(a, b, 1e-7, "free-running step, LinearVehicle traffic")
        np.testing.assert_allclose(oa[0], ob[0], rtol=0, atol=2e-6)


def test_every_reference_env_json_loads_or_raises():
    """scripts/configs/*/env*.json of the reference (copied here as dicts): every file either translates to a device config
    or raises -- no shipped config runs with silently different traffic."""
    configs = {
        "IntersectionEnv/env_linear.json": ("intersection", T.EPISODE_CONFIGS["linear"][0], abi.VEHICLE_LINEAR),
        "IntersectionEnv/env_multi_model.json": ("intersection", dict(T.EPISODE_CONFIGS["linear"][0],
                                                 other_vehicles_type="ttrl_env.vehicle.uncertainty.estimation.MultipleModelVehicle"), None),
        "IntersectionEnv/env.json": ("intersection", T.EPISODE_CONFIGS["envjson"][0], abi.VEHICLE_IDM),
        "IntersectionEnv/env_5fps.json": ("intersection", T.EPISODE_CONFIGS["5fps"][0], abi.VEHICLE_IDM),
        "RoundaboutEnv/env.json": ("roundabout", T.LINEAR, abi.VEHICLE_LINEAR),
        "RoundaboutEnv/env_route_1.json": ("roundabout", dict(T.LINEAR, incoming_vehicle_destination=1), abi.VEHICLE_LINEAR),
    }
    for name, (scene, over, model) in configs.items():
        if scene == "intersection":
            build = lambda: T.intersection_scene(over)[2]
        else:
            build = lambda: T.roundabout_scene(over)[2]
        if model is None:
            with pytest.raises(NotImplementedError):
                build()
        else:
            assert build().vehicle_model == model, name
    ref = "/root/reference/scripts/configs"
    if os.path.isdir(ref):  # in the build container: the files themselves
        for sub in ("IntersectionEnv", "RoundaboutEnv"):
            for f in sorted(os.listdir(os.path.join(ref, sub))):
                if not (f.startswith("env") and f.endswith(".json")):
                    continue
                d = json.load(open(os.path.join(ref, sub, f)))
                scene = "intersection" if sub == "IntersectionEnv" else "roundabout"
                base = scenes.MULTI_AGENT_INTERSECTION_CONFIG if "multi-agent" in d["id"] else (scenes.INTERSECTION_CONFIG if scene == "intersection" else scenes.ROUNDABOUT_CONFIG)
                cfgd = scenes.merged_config(base, {k: v for k, v in d.items() if k not in ("id", "import_module")})
                table = (scenes.make_intersection_network().to_table(scenes.intersection_exit_predicate) if scene == "intersection"
                         else scenes.make_roundabout_network().to_table())
                if "MultipleModelVehicle" in str(d.get("other_vehicles_type", "")):
                    with pytest.raises(NotImplementedError):
                        scenes.build_config(table, cfgd, scene)
                else:
                    cfg = scenes.build_config(table, cfgd, scene)
                    assert cfg.vehicle_model == (abi.VEHICLE_LINEAR if "LinearVehicle" in str(d.get("other_vehicles_type", "")) else abi.VEHICLE_IDM), f

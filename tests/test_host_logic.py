"""Host-side logic (no GPU): ABI binding, config translation, lane tables, route planning, the host-driven
reset (validated with the emulated device primitives against the reference's seeded resets), RNG mirroring."""
import ctypes as C
import re
import os

import numpy as np
import pytest

from topotrafficrl_b200 import abi, scenes
from topotrafficrl_b200._gym import np_random
from topotrafficrl_b200.reset import reset_intersection
from topotrafficrl_b200.state import SimState, pack_route, unpack_route
from tests import common as T
from tests.emu.emu import Emulator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports every function include/ttrl_b200.h declares (no compute calls)."""
    from topotrafficrl_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "ttrl_b200.h")).read()
    names = set(re.findall(r"\b(ttrl_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 25
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert L.ttrl_abi_version() == abi.ABI_VERSION == 2
    sizes = [C.sizeof(x) for x in (abi.Lane, abi.Road, abi.Config, abi.SpawnDraw, abi.EpisodeStats, abi.QnetDesc, abi.ResetParams,
                                   abi.CastMember)]
    assert [L.ttrl_abi_sizeof(i) for i in range(8)] == sizes


def test_product_never_imports_oracle():
    """The product package must not import, link or execute anything under oracle/ or tests/."""
    pkg = os.path.join(ROOT, "topotrafficrl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r"#.*|//.*", "", src).replace("oracle-", ""), f
                assert "tests.emu" not in src and "libttrl_emu" not in src, f


def test_route_packing_roundtrip():
    for route in (None, [], [(3, 0)], [(0, 0), (2, None), (19, None)], [(1, 3), (2, 2), (3, 1), (4, None)],
                  [(k, None if k % 3 else k % 4) for k in range(9)], [(23 - k, None) for k in range(12)]):
        assert unpack_route(*pack_route(route)) == route


def test_intersection_table_matches_reference_layout():
    net, table, cfg, (spawn_lane, rlen, rroad) = T.intersection_scene()
    assert table.n_lanes == 20 and table.n_roads == 20
    # insertion order per corner: incoming, right turn, left turn, straight, exit
    assert table.lane_keys[:5] == [("o0", "ir0", 0), ("ir0", "il3", 0), ("ir0", "il1", 0), ("ir0", "il2", 0), ("il3", "o3", 0)]
    lengths = [table.lanes[i].length for i in range(5)]
    np.testing.assert_allclose(lengths, [100.0, 9 * np.pi / 2, 13 * np.pi / 2, 22.0, 100.0])
    assert [table.lanes[i].priority for i in range(5)] == [1, 1, 0, 1, 1]
    assert [table.lanes[5 + i].priority for i in range(5)] == [3, 3, 2, 3, 3]
    assert [bool(table.lanes[i].is_exit) for i in range(5)] == [False, False, False, False, True]
    assert list(spawn_lane) == [0, 5, 10, 15]
    assert rlen[0, 1] == 2 and rlen[0, 0] == 5  # own arm (multi-agent egos): around the block, 1 + 5 route entries
    assert net.plan_route(("o0", "ir0", 0), "o1") == [("o0", "ir0", 0), ("ir0", "il1", None), ("il1", "o1", None)]
    assert cfg.regulated == 1 and cfg.distance_wanted == 7.0 and cfg.comfort_acc_max == 6.0 and cfg.comfort_acc_min == -3.0


def test_config_translation_and_errors():
    _, table, cfg, _ = T.highway_scene(50)
    assert cfg.action_mode == abi.ACT_ALL and cfg.n_target_speeds == 3 and list(cfg.target_speeds)[:3] == [20.0, 25.0, 30.0]
    assert cfg.obs_vehicles == 15 and cfg.n_features == 7 and cfg.absolute == 0
    # default features_range derived from the 4-lane road (reference observation.py:213-225)
    assert (cfg.range_lo[1], cfg.range_hi[1], cfg.range_lo[2], cfg.range_hi[2], cfg.range_hi[3]) == (-200.0, 200.0, -16.0, 16.0, 80.0)
    with pytest.raises(ValueError):
        scenes.build_config(table, scenes.merged_config(scenes.HIGHWAY_CONFIG, {"observation": {"type": "Nope"}}), "highway")
    with pytest.raises(ValueError):
        scenes.build_config(table, scenes.merged_config(scenes.HIGHWAY_CONFIG, {"action": {"type": "Nope"}}), "highway")
    _, _, gcfg, _ = T.intersection_scene(T.GRID_DENSE)
    assert (gcfg.grid_w, gcfg.grid_h) == (32, 32) and gcfg.grid_has_xrange == 1
    # configure() is a SHALLOW update: a partial observation dict replaces the whole block
    merged = scenes.merged_config(scenes.INTERSECTION_CONFIG, {"observation": {"type": "Kinematics"}})
    assert merged["observation"] == {"type": "Kinematics"}


def test_highway_generator_is_shard_invariant():
    _, _, _, cfgd = T.highway_scene(50)
    full = scenes.make_highway_state(8, cfgd, seed=5)
    shard = scenes.make_highway_state(4, cfgd, seed=5, first_env=4)
    np.testing.assert_array_equal(full.veh_d[:, 4:], shard.veh_d)
    np.testing.assert_array_equal(full.veh_i[:, 4:], shard.veh_i)
    assert (full.env_i[abi.EI_NVEH] == 50).all()
    x = full.veh_d[abi.D_X]
    assert (np.diff(x, axis=1) > 0).all()  # each vehicle is placed ahead of the previous one


def test_numpy_shuffle_draws_depend_only_on_length():
    """envs.py mirrors np_random.shuffle(obs[1:]) by shuffling an index vector with the same Generator."""
    a, _ = np_random(11)
    b, _ = np_random(11)
    rows = np.arange(14 * 7, dtype=np.float64).reshape(14, 7)
    ref = rows.copy()
    a.shuffle(ref)
    perm = np.arange(14)
    b.shuffle(perm)
    np.testing.assert_array_equal(ref, rows[perm])
    assert a.uniform() == b.uniform()


class _EmuBackend:
    def __init__(self, emu, num_envs, vcap):
        self.emu, self.num_envs = emu, num_envs
        self.st = SimState.zeros(num_envs, vcap)

    def spawn(self, draws, longitudinal, position_deviation, speed_deviation, spawn_probability, go_straight):
        return self.emu.spawn(self.st, draws, longitudinal, position_deviation, speed_deviation, spawn_probability, go_straight)

    def substep_none(self):
        self.emu.substep(self.st, None)

    def get_state(self):
        return self.st.copy()

    def set_state(self, st):
        self.st = st.copy().contiguous()


def test_host_driven_reset_reproduces_reference_resets():
    """Seeded reset through (emulated) device primitives == IntersectionEnv.reset(seed) of the reference."""
    g = T.golden("intersection_reset.npz")
    net, table, cfg, routes = T.intersection_scene({"observation": dict(scenes.INTERSECTION_CONFIG["observation"], order="sorted")})
    emu = Emulator(cfg, table, routes)
    seeds = [int(s) for s in g["seed"]]
    backend = _EmuBackend(emu, len(seeds), 32)
    rngs = [np_random(s)[0] for s in seeds]
    cfgd = scenes.merged_config(scenes.INTERSECTION_CONFIG, None)
    st = reset_intersection(backend, rngs, net, table, cfgd, cfg)
    want = T.batch_state(g, "state")
    T.compare_states(st, want, 1e-9, "reset", check_action=True)
    obs = emu.observe(backend.st)
    np.testing.assert_allclose(obs.reshape(g["obs"].shape), g["obs"], rtol=0, atol=2e-6)


class _DeviceDrawRng:
    """Feeds ``reset_intersection`` (the host-driven reset, validated against the reference's resets above) with
    the draws the DEVICE-side reset makes for one env, so both procedures can be compared state for state."""

    class _BG:
        state = None

    def __init__(self, emu, seed, env, episode):
        self.emu, self.seed, self.env, self.episode = emu, seed, env, episode
        self.attempt, self.cur, self.normals = -1, None, 0
        self.bit_generator = self._BG()

    def uniform(self, low=None, high=None):
        if low is None:  # u_spawn opens a new _spawn_vehicle call
            self.attempt += 1
            self.cur = self.emu.reset_attempt_draw(self.seed, self.env, self.episode, self.attempt)
            self.normals = 0
            return self.cur.u_spawn
        return self.cur.delta

    def choice(self, *a, **k):
        return np.array([self.cur.entry, self.cur.exit])

    def normal(self, loc=0.0):
        if loc == 1:  # the ego's longitudinal draw: 5 * normal(1)
            u0, u1 = self.emu.reset_uniforms(self.seed, self.env, self.episode, 0x100)
            return 1.0 + np.sqrt(-2.0 * np.log(1.0 - u0)) * np.cos(2 * np.pi * u1)
        self.normals += 1
        return self.cur.n_pos if self.normals == 1 else self.cur.n_speed


def test_device_reset_follows_make_vehicles():
    """Device-side IntersectionEnv reset (emulated device logic) == the host-driven reset fed the same draws."""
    net, table, cfg, routes = T.intersection_scene()
    cfgd = scenes.merged_config(scenes.INTERSECTION_CONFIG, None)
    emu = Emulator(cfg, table, routes)
    emu.set_reset_params(scenes.intersection_reset_params(cfgd))
    E, seed, first, episode = 12, 77, 1000, 3
    got = SimState.zeros(E, 32)
    emu.reset(got, seed, first, episode)
    backend = _EmuBackend(emu, E, 32)
    rngs = [_DeviceDrawRng(emu, seed, first + e, episode) for e in range(E)]
    want = reset_intersection(backend, rngs, net, table, cfgd, cfg)
    want.env_i[abi.EI_EPISODE] = episode
    T.compare_states(got, want, 1e-12, "device reset")
    assert (got.env_i[abi.EI_EPISODE] == episode).all() and (got.env_i[abi.EI_NVEH] >= 1).all()
    egos = got.env_i[abi.EI_EGO]
    assert all(got.veh_i[abi.I_FLAGS, e, egos[e]] & abi.FL_MDP for e in range(E))
    assert len({tuple(np.round(got.veh_d[abi.D_Y, e, : got.env_i[abi.EI_NVEH, e]], 3)) for e in range(E)}) > E // 2  # envs differ


def test_device_highway_reset_properties():
    """Device-side highway reset: create_random's spacing rule, value ranges, determinism and shard invariance."""
    _, table, cfg, cfgd = T.highway_scene(50, 2.0)
    emu = Emulator(cfg, table)
    emu.set_reset_params(scenes.highway_reset_params(cfgd))
    a = SimState.zeros(6, 50)
    emu.reset(a, 5, 0, 0)
    b = SimState.zeros(3, 50)
    emu.reset(b, 5, 3, 0)
    np.testing.assert_array_equal(a.veh_d[:, 3:], b.veh_d)  # keyed by the GLOBAL env id
    np.testing.assert_array_equal(a.veh_i[:, 3:], b.veh_i)
    c = SimState.zeros(6, 50)
    emu.reset(c, 5, 0, 1)
    assert not np.array_equal(a.veh_d, c.veh_d)  # next episode, new draws
    assert (a.env_i[abi.EI_NVEH] == 50).all() and (a.env_i[abi.EI_EGO] == 0).all()
    x, y, v = a.veh_d[abi.D_X], a.veh_d[abi.D_Y], a.veh_d[abi.D_SPEED]
    assert (np.diff(x, axis=1) > 0).all()
    lane = a.veh_i[abi.I_LANE]
    assert ((lane >= 0) & (lane < 4)).all() and np.array_equal(y, lane * 4.0)
    assert (v[:, 0] == 25.0).all() and ((v[:, 1:] >= 21.0) & (v[:, 1:] <= 24.0)).all()
    lf = np.exp(-5 / 40 * 4)
    off = (1 / 2.0) * (12 + v[:, 1:]) * lf
    gaps = np.diff(x, axis=1)
    assert ((gaps >= 0.9 * off - 1e-9) & (gaps <= 1.1 * off + 1e-9)).all()
    d = a.veh_d[abi.D_DELTA][:, 1:]
    assert ((d >= 3.5) & (d <= 4.5)).all()
    np.testing.assert_allclose(a.veh_d[abi.D_TIMER][:, 1:], ((x[:, 1:] + y[:, 1:]) * np.pi) % 1.0, atol=1e-9)
    assert len(np.unique(lane[:, 1:])) == 4


def test_headless_rasteriser_geometry():
    """render.render_rgb (env.render() with render_mode="rgb_array"): window centred like EnvViewer.window_position with the
    config's centering_position / scaling, ego green, IDM traffic blue, crashed red, lane markings white on grey."""
    from topotrafficrl_b200.render import BLUE, GREEN, GREY, RED, WHITE, render_rgb
    net, table, cfg, _ = T.intersection_scene()
    cfgd = scenes.merged_config(scenes.INTERSECTION_CONFIG, None)
    g = T.golden("intersection_reset.npz")
    st = T.batch_state(g, "state", slice(0, 1))
    img = render_rgb(net, st, cfgd)
    assert img.shape == (600, 600, 3) and img.dtype == np.uint8
    ego = int(st.env_i[abi.EI_EGO, 0])
    cx, cy = int(0.5 * 600), int(0.6 * 600)                      # centering_position [0.5, 0.6]: the ego's pixel
    assert tuple(img[cy, cx]) == GREEN
    scaling = cfgd["scaling"]
    n = int(st.env_i[abi.EI_NVEH, 0])
    seen = 0
    for s in range(n):
        if s == ego:
            continue
        px = int((st.veh_d[abi.D_X, 0, s] - st.veh_d[abi.D_X, 0, ego]) * scaling) + cx
        py = int((st.veh_d[abi.D_Y, 0, s] - st.veh_d[abi.D_Y, 0, ego]) * scaling) + cy
        if 5 < px < 595 and 5 < py < 595:
            assert tuple(img[py, px]) == BLUE, (s, px, py, img[py, px])
            seen += 1
    assert seen >= 2
    assert (img == np.array(WHITE, np.uint8)).all(axis=2).sum() > 500 and (img == np.array(GREY, np.uint8)).all(axis=2).mean() > 0.5
    st.veh_i[abi.I_FLAGS, 0, ego] |= abi.FL_CRASHED
    assert tuple(render_rgb(net, st, cfgd)[cy, cx]) == RED
    # the continuous right-hand line of the ego's incoming lane (o0 -> ir0: x = 2 +- 2 m): a white column right of the ego
    col = int((4.0 - st.veh_d[abi.D_X, 0, ego]) * scaling) + cx
    assert (img[cy - 40:cy + 40, col - 2:col + 3] == np.array(WHITE, np.uint8)).all(axis=2).any(axis=1).mean() > 0.9

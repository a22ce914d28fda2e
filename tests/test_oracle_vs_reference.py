"""Live pinning of the CPU oracle against the UNMODIFIED reference, imported from /root/reference through
oracle/ref_shim.py.  Runs only in the build container (the reference does not exist on the GPU box: skipped
there); uses seeds that are NOT in the committed golden vectors, so it widens their coverage each time it runs.
Bars: discrete fields bit-exact, continuous state 1e-9 per resynced sub-step / 1e-6 per env-step."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/ttrl_env"), reason="reference tree not present")

from tests import common as T  # noqa: E402
from topotrafficrl_b200 import abi, scenes  # noqa: E402


@pytest.fixture(scope="module")
def H():
    from oracle import ref_harness
    yield ref_harness
    ref_harness.restore_idm_class_constants()


def _one(st):
    return st  # extract_state returns a 1-env SimState


def test_intersection_substeps_and_steps_live(H):
    from oracle import oracle as O
    net, table, cfg, routes = T.intersection_scene()
    orc = O.Oracle(cfg, table, routes)
    env = H.IntersectionEnv()
    rng = np.random.default_rng(777)
    n_sub = n_step = 0
    for seed in (901, 902, 903):
        env.reset(seed=seed)
        proxy = H.RecordingRng(env.np_random)
        env.np_random = proxy
        env.road.np_random = proxy
        done = False
        while not done:
            a = int(rng.integers(0, 3))
            # resynced sub-steps on a deep copy of the reference env would disturb its RNG: step the oracle
            # from the reference's pre-step state and compare after the full env.step instead ...
            before = H.extract_state(env, table, 32)
            proxy.log.clear()
            obs, reward, term, trunc, info = env.step(a)
            draw = H.draws_from_log(proxy.log)
            draws = (abi.SpawnDraw * 1)()
            if draw is not None:
                draws[0] = draw
            else:
                draws[0].u_spawn = 2.0
            oo, orr, ot, ou, _ = orc.step(before, np.array([a], np.int32), draws=draws)
            T.compare_states(before, H.extract_state(env, table, 32), T.TOL_STEP, f"seed {seed} step {n_step}")
            np.testing.assert_allclose(oo.reshape(obs.shape), obs, rtol=0, atol=2e-6)
            assert abs(float(orr[0]) - reward) <= 1e-6 and bool(ot[0]) == term and bool(ou[0]) == trunc
            done = term or trunc
            n_step += 1
    assert n_step >= 15
    # ... and resynced single sub-steps on a fresh episode (no spawn draws consumed inside a sub-step)
    env.reset(seed=950)
    for k in range(30):
        a = int(rng.integers(0, 3))
        before = H.extract_state(env, table, 32)
        H.ref_substep(env, a)
        orc.substep(before, np.array([a], np.int32))
        T.compare_states(before, H.extract_state(env, table, 32), T.TOL_SUBSTEP, f"sub-step {k}")
        n_sub += 1
    assert n_sub == 30


@pytest.mark.parametrize("n,density", [(12, 1.5), (50, 2.5)])
def test_highway_substeps_live(H, n, density):
    from oracle import oracle as O
    _, table, cfg, cfgd = T.highway_scene(n, density)
    orc = O.Oracle(cfg, table)
    env = H.SyntheticHighwayEnv(config={"vehicles_count": n, "vehicles_density": density})
    rng = np.random.default_rng(5)
    for seed in (31, 32):
        env.reset(seed=seed)
        a = 1
        for k in range(45):
            if k % 15 == 0:
                a = int(rng.integers(0, 5))
            before = H.extract_state(env, table, n)
            H.ref_substep(env, a)
            orc.substep(before, np.array([a], np.int32))
            T.compare_states(before, H.extract_state(env, table, n), T.TOL_SUBSTEP, f"seed {seed} sub-step {k}")


class _DeviceHighwayDraws:
    """numpy-Generator stand-in that feeds the reference's highway reset (Vehicle.create_random kinematics.py:50-104 +
    IDMVehicle.randomize_behavior behavior.py:66-69) the DEVICE's reset draws, vehicle by vehicle: stream index 2 s -> (lane,
    speed), 2 s + 1 -> (spacing jitter, DELTA) of slot s (ttrl_core.cuh: reset_highway)."""

    def __init__(self, emu, seed, env, episode, lanes, speed_limit):
        self.u = lambda idx: emu.reset_uniforms(seed, env, episode, idx)
        self.lanes, self.speed_limit = lanes, speed_limit
        self.slot, self.choices = 0, 0

    def choice(self, a, *args, **kw):
        self.choices += 1
        if self.choices % 3:  # `_from`, `_to`: one road, a single candidate each
            return a[0]
        lid = int(self.u(2 * self.slot)[0] * self.lanes)
        return min(lid, self.lanes - 1)

    def uniform(self, low=0.0, high=1.0, size=None):
        if (low, high) == (0.9, 1.1):     # spacing jitter: the last draw of create_random
            v = 0.9 + 0.2 * self.u(2 * self.slot + 1)[0]
            self._placed = True
            return v
        if (low, high) == (3.5, 4.5):     # randomize_behavior
            v = 3.5 + self.u(2 * self.slot + 1)[1]
            self.slot += 1
            return v
        assert abs(low - 0.7 * self.speed_limit) < 1e-12 and abs(high - 0.8 * self.speed_limit) < 1e-12, (low, high)
        return (0.7 + 0.1 * self.u(2 * self.slot)[1]) * self.speed_limit


def test_device_highway_reset_is_create_random_with_the_same_draws(H):
    """Device-side highway reset (the device logic built for the host) == the reference's own procedure --
    RoadNetwork.straight_road_network + Vehicle.create_random for the ego and every IDMVehicle + randomize_behavior --
    when the reference is fed the device's draws: positions, lanes, speeds, DELTA, lane-change timers, target speeds."""
    from tests.emu.emu import Emulator
    from topotrafficrl_b200.state import SimState
    n, density = 50, 2.0
    _, table, cfg, cfgd = T.highway_scene(n, density)
    emu = Emulator(cfg, table)
    emu.set_reset_params(scenes.highway_reset_params(cfgd))
    seed, first, episode, E = 17, 40, 2, 4
    got = SimState.zeros(E, n)
    emu.reset(got, seed, first, episode)
    env = H.SyntheticHighwayEnv(config={"vehicles_count": n, "vehicles_density": density})
    for e in range(E):
        env.reset(seed=0)
        draws = _DeviceHighwayDraws(emu, seed, first + e, episode, int(cfgd["lanes_count"]), float(cfgd["speed_limit"]))
        env.np_random = draws
        # the ego's create_random draws no DELTA: advance the slot by hand after it (SyntheticHighwayEnv._reset otherwise)

        def patched():
            H.restore_idm_class_constants()
            from ttrl_env.road.road import Road, RoadNetwork
            from ttrl_env.vehicle.behavior import IDMVehicle
            from ttrl_env.vehicle.kinematics import Vehicle
            net = RoadNetwork.straight_road_network(env.config["lanes_count"], length=env.config["road_length"], speed_limit=env.config["speed_limit"])
            env.road = Road(network=net, np_random=draws, record_history=False)
            ego = Vehicle.create_random(env.road, speed=25, spacing=env.config["ego_spacing"])
            ego = env.action_type.vehicle_class(env.road, ego.position, ego.heading, ego.speed)
            env.controlled_vehicles = [ego]
            env.road.vehicles.append(ego)
            draws.slot = 1
            for _ in range(env.config["vehicles_count"] - 1):
                v = IDMVehicle.create_random(env.road, spacing=1 / env.config["vehicles_density"])
                v.randomize_behavior()
                env.road.vehicles.append(v)

        patched()
        env.define_spaces()
        want = H.extract_state(env, table, n)
        want.env_i[abi.EI_EPISODE] = episode
        want.env_i[abi.EI_STEPS] = 0
        want.env_d[abi.ED_TIME] = 0
        T.compare_states(got.slice_envs(e, e + 1), want, 1e-9, f"highway reset env {e}")

"""TEST INFRASTRUCTURE ONLY -- drive the UNMODIFIED reference (through ``ref_shim``) and convert its
Python objects to/from the struct-of-arrays state of ``include/ttrl_b200.h``.

Used in this container only (``/root/reference`` does not exist on the GPU box): by
``tests/golden/make_golden.py`` to produce the committed fixtures and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import ref_shim

ref_shim.install()

from ttrl_env import utils as ref_utils  # noqa: E402
from ttrl_env.envs.common.abstract import AbstractEnv  # noqa: E402
from ttrl_env.envs.intersection_env import IntersectionEnv, MultiAgentIntersectionEnv  # noqa: E402
from ttrl_env.envs.roundabout_env import RoundaboutEnv  # noqa: E402
from ttrl_env.envs.u_turn_env import UTurnEnv  # noqa: E402
from ttrl_env.road.road import Road, RoadNetwork  # noqa: E402
from ttrl_env.vehicle.behavior import IDMVehicle, LinearVehicle  # noqa: E402
from ttrl_env.vehicle.controller import MDPVehicle  # noqa: E402
from ttrl_env.vehicle.kinematics import Vehicle  # noqa: E402

from topotrafficrl_b200 import abi  # noqa: E402
from topotrafficrl_b200.road import NetworkTable  # noqa: E402
from topotrafficrl_b200.state import SimState  # noqa: E402

IDM_CLASS_DEFAULTS = dict(DISTANCE_WANTED=IDMVehicle.DISTANCE_WANTED, COMFORT_ACC_MAX=IDMVehicle.COMFORT_ACC_MAX,
                          COMFORT_ACC_MIN=IDMVehicle.COMFORT_ACC_MIN)


def restore_idm_class_constants() -> None:
    """IntersectionEnv mutates class attributes of ``other_vehicles_type`` process-wide (intersection_env.py:258-261): on
    IDMVehicle itself, or -- with LinearVehicle traffic -- as new attributes of the subclass, removed again here."""
    for k, v in IDM_CLASS_DEFAULTS.items():
        setattr(IDMVehicle, k, v)
        if k in LinearVehicle.__dict__:
            delattr(LinearVehicle, k)


class SyntheticHighwayEnv(AbstractEnv):
    """Test-only env composed purely of reference classes (the reference ships no HighwayEnv):
    straight_road_network + 1 MDPVehicle ego + IDMVehicles from create_random + randomize_behavior;
    reward/termination per the u_turn_env template (u_turn_env.py:39-77)."""

    @classmethod
    def default_config(cls) -> dict:
        config = super().default_config()
        config.update({
            "observation": {"type": "Kinematics", "vehicles_count": 15,
                            "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
                            "absolute": False, "order": "sorted"},
            "action": {"type": "DiscreteMetaAction"},
            "lanes_count": 4, "vehicles_count": 50, "vehicles_density": 2.0, "ego_spacing": 2.0,
            "road_length": 10000, "speed_limit": 30,
            "duration": 40, "collision_reward": -1.0, "left_lane_reward": 0.1, "high_speed_reward": 0.4,
            "reward_speed_range": [20, 30], "normalize_reward": True, "offroad_terminal": False,
        })
        return config

    def _reset(self) -> None:
        restore_idm_class_constants()
        net = RoadNetwork.straight_road_network(self.config["lanes_count"], length=self.config["road_length"],
                                                speed_limit=self.config["speed_limit"])
        self.road = Road(network=net, np_random=self.np_random, record_history=False)
        ego = Vehicle.create_random(self.road, speed=25, spacing=self.config["ego_spacing"])
        ego = self.action_type.vehicle_class(self.road, ego.position, ego.heading, ego.speed)
        self.controlled_vehicles = [ego]
        self.road.vehicles.append(ego)
        for _ in range(self.config["vehicles_count"] - 1):
            v = IDMVehicle.create_random(self.road, spacing=1 / self.config["vehicles_density"])
            v.randomize_behavior()
            self.road.vehicles.append(v)

    def _rewards(self, action):
        neighbours = self.road.network.all_side_lanes(self.vehicle.lane_index)
        lane = self.vehicle.lane_index[2]
        scaled_speed = ref_utils.lmap(self.vehicle.speed, self.config["reward_speed_range"], [0, 1])
        return {"collision_reward": self.vehicle.crashed,
                "left_lane_reward": lane / max(len(neighbours) - 1, 1),
                "high_speed_reward": np.clip(scaled_speed, 0, 1),
                "on_road_reward": self.vehicle.on_road}

    def _reward(self, action) -> float:
        rewards = self._rewards(action)
        reward = sum(self.config.get(name, 0) * reward for name, reward in rewards.items())
        if self.config["normalize_reward"]:
            reward = ref_utils.lmap(reward, [self.config["collision_reward"],
                                             self.config["high_speed_reward"] + self.config["left_lane_reward"]], [0, 1])
        reward *= rewards["on_road_reward"]
        return reward

    def _is_terminated(self) -> bool:
        return self.vehicle.crashed

    def _is_truncated(self) -> bool:
        return self.time >= self.config["duration"]


# --------------------------------------------------------------------------------------------------
# reference objects <-> SoA state
# --------------------------------------------------------------------------------------------------
def extract_state(env, table: NetworkTable, vcap: int) -> SimState:
    """Snapshot ``env.road.vehicles`` (list order) into a 1-env SimState."""
    vehicles = env.road.vehicles
    linear = any(isinstance(v, LinearVehicle) for v in vehicles) or "LinearVehicle" in str(env.config.get("other_vehicles_type", ""))
    st = SimState.zeros(1, vcap, linear=linear)
    assert len(vehicles) <= vcap, (len(vehicles), vcap)
    for s, v in enumerate(vehicles):
        route = getattr(v, "route", None)
        if route is not None:
            route = [(table.road_index_of[(r[0], r[1])], r[2]) for r in route]
        is_mdp = isinstance(v, MDPVehicle)
        st.set_vehicle(
            0, s, x=float(v.position[0]), y=float(v.position[1]), heading=float(v.heading), speed=float(v.speed),
            lane=table.flat(v.lane_index), target_lane=table.flat(v.target_lane_index),
            target_speed=float(v.target_speed), timer=float(getattr(v, "timer", 0.0)), delta=float(getattr(v, "DELTA", 4.0)),
            mdp=is_mdp, controlled=v in env.controlled_vehicles, crashed=bool(v.crashed),
            agent=env.controlled_vehicles.index(v) if v in env.controlled_vehicles else 0,
            speed_index=int(getattr(v, "speed_index", 0)), route=route,
            steering=float(v.action["steering"]), accel=float(v.action["acceleration"]),
            impact=None if v.impact is None else (float(v.impact[0]), float(v.impact[1])),
            yielding=bool(getattr(v, "is_yielding", False)), yield_timer=int(getattr(v, "yield_timer", 0)),
            linear=(list(v.ACCELERATION_PARAMETERS) + list(v.STEERING_PARAMETERS)) if isinstance(v, LinearVehicle) else None)
    st.env_i[abi.EI_NVEH, 0] = len(vehicles)
    st.env_i[abi.EI_STEPS, 0] = env.steps
    st.env_i[abi.EI_ROAD_STEPS, 0] = getattr(env.road, "steps", 0)
    st.env_i[abi.EI_EGO, 0] = vehicles.index(env.vehicle)
    st.env_d[abi.ED_TIME, 0] = env.time
    return st


def ref_substep(env, action: Optional[int]) -> None:
    """One iteration of ``AbstractEnv._simulate``'s loop body (abstract.py:257-273), rendering omitted."""
    frames = int(env.config["simulation_frequency"] // env.config["policy_frequency"])
    if action is not None and not env.config["manual_control"] and env.steps % frames == 0:
        env.action_type.act(action)
    env.road.act()
    env.road.step(1 / env.config["simulation_frequency"])
    env.steps += 1


class RecordingRng:
    """Proxy around a numpy Generator that records the spawn draws of ``_spawn_vehicle``
    (intersection_env.py:328-346 + behavior.py:66-69) so they can be injected into the oracle / device."""

    def __init__(self, rng):
        self._rng = rng
        self.log: List[tuple] = []
        self.perms: List[np.ndarray] = []  # np_random.shuffle(obs[1:]) (observation.py:272-273): new row i = old row perm[i]

    def shuffle(self, x, *a, **k):
        twin = np.random.Generator(np.random.PCG64())
        twin.bit_generator.state = self._rng.bit_generator.state
        idx = np.arange(len(x))
        twin.shuffle(idx)  # the permutation depends on the stream and the row count only
        self.perms.append(idx)
        return self._rng.shuffle(x, *a, **k)

    def uniform(self, *a, **k):
        v = self._rng.uniform(*a, **k)
        self.log.append(("uniform", v))
        return v

    def normal(self, *a, **k):
        v = self._rng.normal(*a, **k)
        self.log.append(("normal", v))
        return v

    def choice(self, *a, **k):
        v = self._rng.choice(*a, **k)
        self.log.append(("choice", np.array(v).copy()))
        return v

    def __getattr__(self, name):
        return getattr(self._rng, name)


def draws_from_log(log) -> Optional[abi.SpawnDraw]:
    """Turn the recorded draws of ONE ``_spawn_vehicle`` call into a SpawnDraw record."""
    d = abi.SpawnDraw()
    d.u_spawn = 2.0  # > any probability: "no spawn"
    if not log:
        return d
    assert log[0][0] == "uniform"
    d.u_spawn = float(log[0][1])
    if len(log) == 1:
        return d
    assert log[1][0] == "choice" and log[2][0] == "normal" and log[3][0] == "normal", log
    d.entry, d.exit = int(log[1][1][0]), int(log[1][1][1])
    d.n_pos, d.n_speed = float(log[2][1]), float(log[3][1])
    d.delta = 4.0
    if len(log) > 5:    # LinearVehicle.randomize_behavior: uniform(size=3), uniform(size=2)
        u = list(np.ravel(log[4][1])) + list(np.ravel(log[5][1]))
        for k in range(5):
            d.lin_u[k] = float(u[k])
    elif len(log) > 4:  # IDMVehicle.randomize_behavior: uniform(3.5, 4.5)
        d.delta = float(log[4][1])
    return d

"""Shared helpers for the parity tests: golden fixtures -> SimState, scene construction, comparisons."""
from __future__ import annotations

import os

import numpy as np

from topotrafficrl_b200 import abi, scenes
from topotrafficrl_b200.state import SimState

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GRID_DENSE = {"observation": {"type": "OccupancyGrid", "vehicles_count": 15,
                              "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
                              "features_range": {"x": [-100, 100], "y": [-100, 100], "vx": [-20, 20], "vy": [-20, 20]},
                              "grid_size": [[-32, 32], [-32, 32]], "grid_step": [2, 2], "absolute": False}}
GRID_ROAD = {"observation": {"type": "OccupancyGrid", "features": ["presence", "vx", "vy", "on_road"],
                             "grid_size": [[-27.5, 27.5], [-27.5, 27.5]], "grid_step": [5, 5], "absolute": False}}
HIGHWAY_GRID = {"observation": {"type": "OccupancyGrid", "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
                                "grid_size": [[-32, 32], [-32, 32]], "grid_step": [2, 2], "absolute": False}}

# continuous-state tolerances of the parity tests.  The device computes in float64 like the reference, so
# these are far inside the north-star tolerance (1e-4 m / 1e-5 rad per step): they only absorb last-ulp
# differences between libm / numpy / CUDA transcendental functions.
TOL_SUBSTEP = 1e-9
TOL_STEP = 1e-6


# Reference-shipped env configs (scripts/configs/IntersectionEnv/*.json) pinned as whole seeded episodes
# (tests/golden/intersection_ep_<name>.npz, written by make_golden.py `episodes`): name -> (config overrides, seeds)
_KIN_OBS = {"type": "Kinematics", "vehicles_count": 15, "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
            "features_range": {"x": [-100, 100], "y": [-100, 100], "vx": [-20, 20], "vy": [-20, 20]}, "absolute": True}
LINEAR = {"other_vehicles_type": "ttrl_env.vehicle.behavior.LinearVehicle"}  # RoundaboutEnv/env.json
EPISODE_CONFIGS = {
    # env.json as shipped: "order": "shuffled" (BASELINE configs[0])
    "envjson": ({"observation": dict(_KIN_OBS, order="shuffled"), "destination": "o1"}, list(range(100, 112))),
    # env_5fps.json: policy_frequency 5 -> 3 sub-steps per step, regulation ticks not aligned to steps, 75-step episodes
    "5fps": ({"policy_frequency": 5, "show_history": False, "destination": "o1", "duration": 15,
              "observation": dict(_KIN_OBS, order="sorted")}, [120, 121, 122]),
    # env_linear.json / env_multi_model.json with the IDM traffic class: normalize_reward, destination None (drawn per
    # episode), un-normalised shuffled observation, 9 initial vehicles, spawn probability 0.3
    "normdest": ({"initial_vehicle_count": 9, "spawn_probability": 0.3, "observation": dict(_KIN_OBS, normalize=False, order="shuffled"),
                  "normalize_reward": True, "destination": None}, list(range(130, 136))),
    # env_linear.json as shipped: LinearVehicle traffic (behavior.py:350-558)
    "linear": ({"other_vehicles_type": "ttrl_env.vehicle.behavior.LinearVehicle", "initial_vehicle_count": 9, "spawn_probability": 0.3,
                "observation": dict(_KIN_OBS, normalize=False, order="shuffled"), "normalize_reward": True, "destination": None},
               list(range(160, 166))),
    "straight": ({"destination": "o2"}, [140, 141, 142, 143]),   # env_straight.json
    "right": ({"destination": "o3"}, [150, 151, 152, 153]),      # env_right.json
}


def golden(name: str):
    return np.load(os.path.join(GOLDEN, name))


def batch_state(g, prefix: str, sel=None) -> SimState:
    """Stack S golden 1-env snapshots into one S-env SimState."""
    vd, vi, ei, ed = (g[prefix + s] for s in ("_vd", "_vi", "_ei", "_ed"))
    if sel is not None:
        vd, vi, ei, ed = vd[sel], vi[sel], ei[sel], ed[sel]
    if vi.shape[1] < abi.NI:  # fixtures written before the route grew to 12 entries: words 1, 2 are zero (routes <= 4 entries)
        vi = np.concatenate([vi, np.zeros((vi.shape[0], abi.NI - vi.shape[1], vi.shape[2]), vi.dtype)], axis=1)
    lin = None
    if prefix + "_lin" in g.files:  # LinearVehicle parameters
        lin = g[prefix + "_lin"] if sel is None else g[prefix + "_lin"][sel]
        lin = np.ascontiguousarray(lin.transpose(1, 0, 2))
    return SimState(np.ascontiguousarray(vd.transpose(1, 0, 2)), np.ascontiguousarray(vi.transpose(1, 0, 2)),
                    np.ascontiguousarray(ei.T), np.ascontiguousarray(ed.T), lin)


def intersection_scene(overrides=None):
    net = scenes.make_intersection_network()
    table = net.to_table(scenes.intersection_exit_predicate)
    cfgd = scenes.merged_config(scenes.INTERSECTION_CONFIG, overrides)
    cfg = scenes.build_config(table, cfgd, "intersection")
    return net, table, cfg, scenes.intersection_spawn_routes(net, table)


def highway_scene(n_vehicles=50, density=2.0, overrides=None):
    net = scenes.make_highway_network(4)
    table = net.to_table()
    over = {"vehicles_count": n_vehicles, "vehicles_density": density}
    if overrides:
        over.update(overrides)
    cfgd = scenes.merged_config(scenes.HIGHWAY_CONFIG, over)
    cfg = scenes.build_config(table, cfgd, "highway", ego_lanes_count=4)
    return net, table, cfg, cfgd


def draws_array(draw_rows):
    arr = (abi.SpawnDraw * len(draw_rows))()
    for k, row in enumerate(draw_rows):
        arr[k].u_spawn, arr[k].entry, arr[k].exit = float(row[0]), int(row[1]), int(row[2])
        arr[k].n_pos, arr[k].n_speed, arr[k].delta = float(row[3]), float(row[4]), float(row[5])
        for q in range(5):  # LinearVehicle.randomize_behavior's uniforms (fixtures of IDM traffic carry six columns)
            arr[k].lin_u[q] = float(row[6 + q]) if len(row) > 6 else 0.0
    return arr


def compare_states(got: SimState, want: SimState, tol: float, what: str = "", check_action: bool = True):
    """Discrete fields bit-exact on live slots, continuous within tol.  Returns max abs continuous diff."""
    n_got, n_want = got.env_i[abi.EI_NVEH], want.env_i[abi.EI_NVEH]
    assert (n_got == n_want).all(), f"{what}: vehicle counts differ at envs {np.nonzero(n_got != n_want)[0][:8]}"
    live = want.live_mask()
    for f, name in enumerate(["lane", "target_lane", "flags", "speed_index", "route_len", "route_road", "route_lane", "yield_timer",
                              "route_road1", "route_road2", "route_lane1", "route_lane2"]):
        bad = (got.veh_i[f] != want.veh_i[f]) & live
        assert not bad.any(), (f"{what}: discrete field {name} differs at (env,slot) {np.argwhere(bad)[:8].tolist()} "
                               f"got {got.veh_i[f][bad][:8]} want {want.veh_i[f][bad][:8]}")
    for f in (abi.EI_STEPS, abi.EI_ROAD_STEPS, abi.EI_EGO):
        assert (got.env_i[f] == want.env_i[f]).all(), f"{what}: env int field {f} differs"
    worst = 0.0
    assert (got.lin is None) == (want.lin is None), f"{what}: LinearVehicle parameter block present on one side only"
    if want.lin is not None:
        d = np.abs(got.lin - want.lin)[:, live]
        assert d.size == 0 or d.max() <= 1e-12, f"{what}: LinearVehicle parameters differ by {d.max()}"
    for f in range(abi.ND):
        if not check_action and f in (abi.D_STEERING, abi.D_ACCEL):
            continue
        d = np.abs(got.veh_d[f] - want.veh_d[f])[live]
        if d.size:
            m = float(d.max())
            assert m <= tol, f"{what}: continuous field {f} differs by {m} (> {tol})"
            worst = max(worst, m)
    return worst


def stress_states(cfgd, E: int, n: int, seed: int = 0):
    """Adversarial highway states for the task-queue paths of the kernel (tests only):
    * env 0..E/3: a pile-up -- every vehicle within a few metres of one point, so every pair passes the collision
      pre-check (n(n-1)/2 candidate pairs overflow the pair queue -> serial fallback) and most pairs intersect;
    * next third: every lane-change timer already elapsed (all vehicles enter MOBIL in the same sub-step: several
      MOBIL batches) in dense traffic with many ongoing lane changes (phase B);
    * rest: ragged vehicle counts 1..n.
    """
    st = scenes.make_highway_state(E, cfgd, seed=seed)
    rng = np.random.default_rng(seed + 99)
    third = max(E // 3, 1)
    for e in range(E):
        if e < third:
            st.veh_d[abi.D_X, e, :n] = 300.0 + rng.uniform(-4, 4, size=n)
            st.veh_d[abi.D_Y, e, :n] = rng.integers(0, 4, size=n) * 4.0 + rng.uniform(-1.5, 1.5, size=n)
            st.veh_d[abi.D_HEADING, e, :n] = rng.uniform(-0.3, 0.3, size=n)
            st.veh_i[abi.I_LANE, e, :n] = np.clip(np.round(st.veh_d[abi.D_Y, e, :n] / 4.0), 0, 3).astype(np.int32)
            st.veh_i[abi.I_TARGET_LANE, e, :n] = st.veh_i[abi.I_LANE, e, :n]
        elif e < 2 * third:
            st.veh_d[abi.D_TIMER, e, 1:n] = 1.0 + rng.uniform(0.01, 0.5, size=n - 1)
            tl = np.clip(st.veh_i[abi.I_LANE, e, 1:n] + rng.integers(-1, 2, size=n - 1), 0, 3)
            chg = rng.random(n - 1) < 0.4
            st.veh_i[abi.I_TARGET_LANE, e, 1:n] = np.where(chg, tl, st.veh_i[abi.I_TARGET_LANE, e, 1:n])
            st.veh_d[abi.D_TIMER, e, 1:n] = np.where(chg, 0.3, st.veh_d[abi.D_TIMER, e, 1:n])
        else:
            k = 1 + (e * 7) % n
            st.veh_d[:, e, k:] = 0
            st.veh_i[:, e, k:] = 0
            st.env_i[abi.EI_NVEH, e] = k
    return st


# --------------------------------------------------------------------------------------------------
# scenes of SURVEY.md section 8 rows A29 / N3: multi-agent intersection, roundabout, u-turn
# --------------------------------------------------------------------------------------------------
MULTI_AGENT = {"observation": {"type": "MultiAgentObservation",
                               "observation_config": {"type": "Kinematics", "vehicles_count": 15,
                                                      "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
                                                      "features_range": {"x": [-100, 100], "y": [-100, 100], "vx": [-20, 20], "vy": [-20, 20]},
                                                      "absolute": True, "order": "sorted"}},
               "initial_vehicle_count": 5, "controlled_vehicles": 4}
UTURN_KIN = {"observation": {"type": "Kinematics", "vehicles_count": 6, "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
                             "absolute": False, "order": "sorted"}}


def multi_agent_scene(overrides=MULTI_AGENT):
    net = scenes.make_intersection_network()
    table = net.to_table(scenes.intersection_exit_predicate)
    cfgd = scenes.merged_config(scenes.MULTI_AGENT_INTERSECTION_CONFIG, overrides)
    cfg = scenes.build_config(table, cfgd, "intersection")
    return net, table, cfg, cfgd, scenes.intersection_spawn_routes(net, table)


def roundabout_scene(overrides=None):
    net = scenes.make_roundabout_network()
    table = net.to_table()
    cfgd = scenes.merged_config(scenes.ROUNDABOUT_CONFIG, overrides)
    return net, table, scenes.build_config(table, cfgd, "roundabout", ego_lanes_count=1), cfgd


def uturn_scene(overrides=UTURN_KIN):
    net = scenes.make_uturn_network()
    table = net.to_table()
    cfgd = scenes.merged_config(scenes.UTURN_CONFIG, overrides)
    return net, table, scenes.build_config(table, cfgd, "u-turn", ego_lanes_count=2), cfgd
